# round 2, call 5: runtime-selected dynamic schedule (work units, 2 claims in flight), __maxnreg__, area one tile ahead on small shards,
# fc_run_steps as CUDA graphs; new tests
set -x
V=components/flux_calculator_b200/csrc/build_variants
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -25
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-parity"
run() { name=$1; shift; "$@" > gpurun_out/r2_05_$name.json 2>>gpurun_out/r2_05.err; cut -c1-160 gpurun_out/r2_05_$name.json; }
run shard $B --workload C4 --cells 1250000 --steps 1000 --warmup 50
run shard_nodiag $B --workload C4 --cells 1250000 --diag 0 --steps 1000 --warmup 50
run c4 $B --workload C4
run c4_nodiag $B --workload C4 --diag 0
run c3 $B --workload C3 --steps 1000 --warmup 50
run c3_1e7 $B --workload C3 --cells 10000000
run c5 $B --workload C5
run c5_diag $B --workload C5 --diag 1
run c2 $B --workload C2 --steps 2000 --warmup 50
export FLUXCALC_LIB=$V/libfluxcalc_regs96.so
run regs96_shard $B --workload C4 --cells 1250000 --steps 1000 --warmup 50
run regs96_c4 $B --workload C4
run regs96_c4_nodiag $B --workload C4 --diag 0
run regs96_c3_1e7 $B --workload C3 --cells 10000000
export FLUXCALC_LIB=$V/libfluxcalc_static.so
run static_c5 $B --workload C5
run static_c3_1e7 $B --workload C3 --cells 10000000
unset FLUXCALC_LIB
tail -5 gpurun_out/r2_05.err
