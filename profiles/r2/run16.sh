# round 2, call 16 (1 GPU): generic fused kernel moved to the flag-and-recompute scheme (no by-value cold call): full -m gpu
# suite, three surface types at 10^7 cells (before: profiles/r2_bench_s3_generic_kernel.json, 1.88 ms) and C4 on the generic kernel
set -x
export COLUMNS=200
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rfs --tb=short --timeout 300 -p no:cacheprovider > $O/r2_16_pytest.log 2>&1
tail -6 $O/r2_16_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline"
$B --workload S3 > $O/r2_16_s3.json 2> $O/r2_16_s3.err; cut -c1-200 $O/r2_16_s3.json; tail -2 $O/r2_16_s3.err
$B --workload C4 --staged 0 --diag 0 > $O/r2_16_c4_generic.json 2> $O/r2_16_c4_generic.err; cut -c1-200 $O/r2_16_c4_generic.json; tail -2 $O/r2_16_c4_generic.err
python - <<'EOF'
import json
for f in ("r2_16_s3", "r2_16_c4_generic"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); r = d["roofline"]
        print(f, "ms/step %.4f kernel_ms %.4f B/cell %d frac %.3f frac_step %.3f kernel %s parity %s" % (d["ms_per_step"], r["kernel_ms"], r["algorithmic_bytes_per_cell"], r["frac"], r["frac_from_ms_per_step"], r["kernel"], (d.get("parity") or {}).get("worst_error_over_tolerance")))
    except Exception as e:
        print(f, "FAILED", e)
EOF
