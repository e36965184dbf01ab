# round 2, call 12: byte-balanced static schedule; targeted tests; final-kernel ncu captures
set -x
export COLUMNS=200
timeout 600 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_full_size.py tests/test_step_golden.py tests/test_standalone.py -m gpu -q -rf --tb=short --timeout 120 -p no:cacheprovider > gpurun_out/r2_12_pytest.log 2>&1
tail -8 gpurun_out/r2_12_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-parity"
run() { name=$1; shift; "$@" > gpurun_out/r2_12_$name.json 2>>gpurun_out/r2_12.err; cut -c1-160 gpurun_out/r2_12_$name.json; }
run shard $B --workload C4 --cells 1250000 --steps 2000 --warmup 50
run shard_b $B --workload C4 --cells 1250000 --steps 2000 --warmup 50
run c4 $B --workload C4
N="ncu --set full --clock-control none --import-source on -k regex:flux_spec_kernel -s 3 -c 1 -f"
C="$B --workload C4 --cells 1250000 --steps 3 --warmup 3"
$C > gpurun_out/plain_a.log 2>&1 && $N -o gpurun_out/r2_12_prof_c4_shard $C > gpurun_out/ncu_a.log 2>&1
C="$B --workload C4 --diag 0 --steps 3 --warmup 3"
$C > gpurun_out/plain_b.log 2>&1 && $N -o gpurun_out/r2_12_prof_c4_nodiag_dyn $C > gpurun_out/ncu_b.log 2>&1
C="$B --workload C3 --cells 10000000 --steps 3 --warmup 3"
$C > gpurun_out/plain_c.log 2>&1 && $N -o gpurun_out/r2_12_prof_c3_1e7_dyn $C > gpurun_out/ncu_c.log 2>&1
C="$B --workload C4 --steps 3 --warmup 3"
$C > gpurun_out/plain_d.log 2>&1 && $N -o gpurun_out/r2_12_prof_c4 $C > gpurun_out/ncu_d.log 2>&1
C="$B --workload C4 --steps 2 --warmup 3"
$C > gpurun_out/plain_e.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_12_launches_c4.csv $C > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/r2_12*; tail -3 gpurun_out/r2_12.err
