# round 2, call 1: baseline numbers of the weak instantiations + ncu captures to work from
set -x
B="python bench.py --no-e2e --no-cpu-baseline --no-parity"
$B --workload C4 --cells 1250000 --steps 400 --warmup 20 > gpurun_out/r2_01_c4_shard.json 2>gpurun_out/r2_01.err; cut -c1-260 gpurun_out/r2_01_c4_shard.json
$B --workload C4 --cells 1250000 --diag 0 --steps 400 --warmup 20 > gpurun_out/r2_01_c4_shard_nodiag.json 2>>gpurun_out/r2_01.err; cut -c1-260 gpurun_out/r2_01_c4_shard_nodiag.json
$B --workload C4 --diag 0 > gpurun_out/r2_01_c4_nodiag.json 2>>gpurun_out/r2_01.err; cut -c1-260 gpurun_out/r2_01_c4_nodiag.json
$B --workload C4 > gpurun_out/r2_01_c4.json 2>>gpurun_out/r2_01.err; cut -c1-260 gpurun_out/r2_01_c4.json
$B --workload C3 --steps 400 --warmup 20 > gpurun_out/r2_01_c3.json 2>>gpurun_out/r2_01.err; cut -c1-260 gpurun_out/r2_01_c3.json
$B --workload C3 --cells 10000000 > gpurun_out/r2_01_c3_1e7.json 2>>gpurun_out/r2_01.err; cut -c1-260 gpurun_out/r2_01_c3_1e7.json
$B --workload C5 > gpurun_out/r2_01_c5.json 2>>gpurun_out/r2_01.err; cut -c1-260 gpurun_out/r2_01_c5.json
N="ncu --set full --clock-control none --import-source on -k regex:flux_spec_kernel -s 3 -c 1 -f"
C="$B --workload C4 --cells 1250000 --steps 3 --warmup 3"
$C > gpurun_out/plain_a.log 2>&1 && $N -o gpurun_out/r2_01_prof_c4_shard $C > gpurun_out/ncu_a.log 2>&1
C="$B --workload C4 --diag 0 --steps 3 --warmup 3"
$C > gpurun_out/plain_b.log 2>&1 && $N -o gpurun_out/r2_01_prof_c4_nodiag $C > gpurun_out/ncu_b.log 2>&1
C="$B --workload C3 --cells 10000000 --steps 3 --warmup 3"
$C > gpurun_out/plain_c.log 2>&1 && $N -o gpurun_out/r2_01_prof_c3_1e7 $C > gpurun_out/ncu_c.log 2>&1
C="$B --workload C3 --steps 3 --warmup 3"
$C > gpurun_out/plain_d.log 2>&1 && $N -o gpurun_out/r2_01_prof_c3 $C > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/r2_01*
