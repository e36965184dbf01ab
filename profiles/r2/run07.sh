# round 2, call 7: two-type kernel back at 96 registers; full GPU suite; C5 batched through fc_run_steps; e2e
set -x
export COLUMNS=220
V=components/flux_calculator_b200/csrc/build_variants
timeout 1700 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -40
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-parity"
run() { name=$1; shift; "$@" > gpurun_out/r2_07_$name.json 2>>gpurun_out/r2_07.err; cut -c1-160 gpurun_out/r2_07_$name.json; }
run c5 $B --workload C5
run c5_nographs $B --workload C5 --opt graphs=0
run c5_diag $B --workload C5 --diag 1
FLUXCALC_LIB=$V/libfluxcalc_static.so run static_c5 $B --workload C5
run c2 $B --workload C2 --steps 2000 --warmup 50
run c5_full timeout 900 python bench.py --workload C5 --no-cpu-baseline
tail -5 gpurun_out/r2_07.err
