# round 2, call 14 (1 GPU): full -m gpu suite, the default bench lines that the index names (C4, C5 at one GPU), A/B of the
# ring-bypass builds (FC_SPEC_BYPASS, csrc/Makefile `variant`) against the shipped library, short-run behaviour of the
# 8-GPU shard (the driver times 20 steps after 5 warm-up steps), CPU arm lines, parity table
set -x
export COLUMNS=200
O=gpurun_out
V=components/flux_calculator_b200/csrc/build_variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_14_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rfs --tb=short --timeout 300 -p no:cacheprovider > $O/r2_14_pytest.log 2>&1
tail -6 $O/r2_14_pytest.log
timeout 400 python bench.py > $O/r2_14_c4_n1.json 2> $O/r2_14_c4_n1.err; cut -c1-250 $O/r2_14_c4_n1.json; tail -2 $O/r2_14_c4_n1.err
timeout 400 python bench.py --workload C5 --no-cpu-baseline > $O/r2_14_c5_n1.json 2> $O/r2_14_c5_n1.err; cut -c1-250 $O/r2_14_c5_n1.json; tail -2 $O/r2_14_c5_n1.err

B="timeout 200 python bench.py --no-e2e --no-cpu-baseline"
ab() { name=$1; lib=$2; shift 2; FLUXCALC_LIB=$lib "$@" > $O/r2_14_ab_$name.json 2>>$O/r2_14_ab.err; python - $O/r2_14_ab_$name.json $name <<'E'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print("AB %-28s ms/step %.4f  kernel_ms %.4f  frac %.3f  frac_step %.3f  parity %s" % (sys.argv[2], d["ms_per_step"], r["kernel_ms"], r["frac"], r["frac_from_ms_per_step"], (d.get("parity") or {}).get("worst_error_over_tolerance")))
except Exception as e:
    print("AB %-28s FAILED %s" % (sys.argv[2], e))
E
}
BASE=components/flux_calculator_b200/libfluxcalc_b200.so
# two surface types (C5's kernel), 300 single steps: shipped (dynamic schedule), static schedule, bypass builds
for rep in 1 2; do
ab c5_base_dyn_$rep      $BASE $B --workload C5 --steps 300
ab c5_base_static_$rep   $BASE $B --workload C5 --steps 300 --opt dyn_min_tiles=100000000
ab c5_bypass_dyn_$rep    $V/libfluxcalc_bypass2.so $B --workload C5 --steps 300
ab c5_bypass_static_$rep $V/libfluxcalc_bypass2.so $B --workload C5 --steps 300 --opt dyn_min_tiles=100000000
done
# one surface type: C4 and its 8-GPU shard
ab c4_base               $BASE $B --workload C4 --no-parity
ab c4_bypass             $V/libfluxcalc_bypass3.so $B --workload C4
for rep in 1 2; do
ab shard_base_$rep       $BASE $B --workload C4 --cells 1250000 --steps 2000 --warmup 50 --no-parity
ab shard_bypass_$rep     $V/libfluxcalc_bypass3.so $B --workload C4 --cells 1250000 --steps 2000 --warmup 50 --no-parity
done
# what the driver's scaling run times: 20 steps after 5 warm-up steps
for rep in 1 2 3; do
ab shard_short_$rep      $BASE $B --workload C4 --cells 1250000 --steps 20 --warmup 5 --no-parity
done
ab shard_short_bypass    $V/libfluxcalc_bypass3.so $B --workload C4 --cells 1250000 --steps 20 --warmup 5 --no-parity
# the bypass builds through the parity suites
FLUXCALC_LIB=$V/libfluxcalc_bypass3.so timeout 600 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_full_size.py tests/test_step_golden.py -m gpu -q -rf --tb=short --timeout 300 -p no:cacheprovider > $O/r2_14_pytest_bypass3.log 2>&1
tail -4 $O/r2_14_pytest_bypass3.log
# CPU arm (the reference-side lines) and the parity table
timeout 500 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_14_c4_reference_arm.json 2> $O/r2_14_ref.err; cut -c1-250 $O/r2_14_c4_reference_arm.json
timeout 200 python bench.py --impl reference --workload C1 --steps 200 --warmup 5 > $O/r2_14_c1_reference.json 2>> $O/r2_14_ref.err; cut -c1-250 $O/r2_14_c1_reference.json
timeout 400 python profiles/parity_table.py > $O/r2_14_parity_table.md 2> $O/r2_14_parity.err; tail -5 $O/r2_14_parity_table.md
ls -la $O | tail -50
