# round 2, call 3: why is <0,1,0> slower at 1e7 than at 1.25e6 per byte?  array stagger, size sweep, S=2 early loads
set -x
B="python bench.py --no-e2e --no-cpu-baseline --no-parity"
for st in 0 4352 69888 266240; do
FC_ALLOC_STAGGER=$st $B --workload C4 --diag 0 > gpurun_out/r2_03_c4_nodiag_st$st.json 2>>gpurun_out/r2_03.err; cut -c1-200 gpurun_out/r2_03_c4_nodiag_st$st.json
done
FC_ALLOC_STAGGER=4352 $B --workload C4 > gpurun_out/r2_03_c4_st4352.json 2>>gpurun_out/r2_03.err; cut -c1-200 gpurun_out/r2_03_c4_st4352.json
FC_ALLOC_STAGGER=4352 $B --workload C3 --cells 10000000 > gpurun_out/r2_03_c3_1e7_st4352.json 2>>gpurun_out/r2_03.err; cut -c1-200 gpurun_out/r2_03_c3_1e7_st4352.json
FC_ALLOC_STAGGER=4352 $B --workload C5 > gpurun_out/r2_03_c5_st4352.json 2>>gpurun_out/r2_03.err; cut -c1-200 gpurun_out/r2_03_c5_st4352.json
$B --workload C5 --opt early_loads=0 > gpurun_out/r2_03_c5_noearly.json 2>>gpurun_out/r2_03.err; cut -c1-200 gpurun_out/r2_03_c5_noearly.json
for n in 2500000 5000000 20000000; do
$B --workload C4 --diag 0 --cells $n > gpurun_out/r2_03_c4_nodiag_n$n.json 2>>gpurun_out/r2_03.err; cut -c1-200 gpurun_out/r2_03_c4_nodiag_n$n.json
$B --workload C4 --cells $n > gpurun_out/r2_03_c4_n$n.json 2>>gpurun_out/r2_03.err; cut -c1-200 gpurun_out/r2_03_c4_n$n.json
done
tail -3 gpurun_out/r2_03.err
