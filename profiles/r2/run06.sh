# round 2, call 6: fixed claim accounting, register caps 96 / 112; full tests with readable failures; new bench.py
set -x
export COLUMNS=220
V=components/flux_calculator_b200/csrc/build_variants
timeout 1700 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -60
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-parity"
run() { name=$1; shift; "$@" > gpurun_out/r2_06_$name.json 2>>gpurun_out/r2_06.err; cut -c1-160 gpurun_out/r2_06_$name.json; }
run shard $B --workload C4 --cells 1250000 --steps 1000 --warmup 50
run shard_nodiag $B --workload C4 --cells 1250000 --diag 0 --steps 1000 --warmup 50
run c4_nodiag $B --workload C4 --diag 0
run c3 $B --workload C3 --steps 1000 --warmup 50
run c3_1e7 $B --workload C3 --cells 10000000
run c5 $B --workload C5
run c5_diag $B --workload C5 --diag 1
run c2 $B --workload C2 --steps 2000 --warmup 50
FLUXCALC_LIB=$V/libfluxcalc_static.so run static_c5 $B --workload C5
run c4_full timeout 600 python bench.py
run ref_c4 timeout 600 python bench.py --impl reference --steps 10 --warmup 2
tail -5 gpurun_out/r2_06.err
