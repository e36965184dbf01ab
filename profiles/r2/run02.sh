# round 2, call 2: deferred fold + early loads: parity suite, then A/B of the knobs
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --no-e2e --no-cpu-baseline --no-parity"
S="--workload C4 --cells 1250000 --steps 1000 --warmup 50"
for rep in 1 2; do
$B $S > gpurun_out/r2_02_shard_$rep.json 2>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_shard_$rep.json
$B $S --opt early_loads=0 > gpurun_out/r2_02_shard_noearly_$rep.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_shard_noearly_$rep.json
$B $S --diag 0 > gpurun_out/r2_02_shard_nodiag_$rep.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_shard_nodiag_$rep.json
$B $S --diag 0 --opt early_loads=0 > gpurun_out/r2_02_shard_nodiag_noearly_$rep.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_shard_nodiag_noearly_$rep.json
done
$B --workload C4 > gpurun_out/r2_02_c4.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_c4.json
$B --workload C4 --diag 0 > gpurun_out/r2_02_c4_nodiag.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_c4_nodiag.json
$B --workload C4 --diag 0 --opt early_loads=0 > gpurun_out/r2_02_c4_nodiag_noearly.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_c4_nodiag_noearly.json
FC_NO_PDL=1 $B --workload C4 --diag 0 > gpurun_out/r2_02_c4_nodiag_nopdl.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_c4_nodiag_nopdl.json
FC_NO_PDL=1 $B --workload C4 > gpurun_out/r2_02_c4_nopdl.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_c4_nopdl.json
$B --workload C3 --cells 10000000 > gpurun_out/r2_02_c3_1e7.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_c3_1e7.json
FC_NO_PDL=1 $B --workload C3 --cells 10000000 > gpurun_out/r2_02_c3_1e7_nopdl.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_c3_1e7_nopdl.json
$B --workload C5 > gpurun_out/r2_02_c5.json 2>>gpurun_out/r2_02.err; cut -c1-200 gpurun_out/r2_02_c5.json
tail -5 gpurun_out/r2_02.err
