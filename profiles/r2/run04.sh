# round 2, call 4: dynamic tile schedule (no-diagnostics instantiations), area prefetch, two-type geometry 3 teams x 5 warps
set -x
V=components/flux_calculator_b200/csrc/build_variants
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
FLUXCALC_LIB=$V/libfluxcalc_t3w5.so timeout 600 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-parity"
run() { name=$1; shift; "$@" > gpurun_out/r2_04_$name.json 2>>gpurun_out/r2_04.err; cut -c1-160 gpurun_out/r2_04_$name.json; }
run shard $B --workload C4 --cells 1250000 --steps 1000 --warmup 50
run shard_nodiag $B --workload C4 --cells 1250000 --diag 0 --steps 1000 --warmup 50
run c4 $B --workload C4
run c4_nodiag $B --workload C4 --diag 0
run c3 $B --workload C3 --steps 1000 --warmup 50
run c3_1e7 $B --workload C3 --cells 10000000
run c5 $B --workload C5
run c2 $B --workload C2 --steps 2000 --warmup 50
export FLUXCALC_LIB=$V/libfluxcalc_static.so
run static_shard_nodiag $B --workload C4 --cells 1250000 --diag 0 --steps 1000 --warmup 50
run static_c4_nodiag $B --workload C4 --diag 0
run static_c3 $B --workload C3 --steps 1000 --warmup 50
run static_c3_1e7 $B --workload C3 --cells 10000000
run static_c5 $B --workload C5
run static_c2 $B --workload C2 --steps 2000 --warmup 50
export FLUXCALC_LIB=$V/libfluxcalc_t3w5.so
run t3w5_c5 $B --workload C5
run t3w5_c5_diag $B --workload C5 --diag 1
unset FLUXCALC_LIB
run c5_diag $B --workload C5 --diag 1
tail -5 gpurun_out/r2_04.err
