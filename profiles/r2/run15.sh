# round 2, call 15 (1 GPU): the library as shipped (barriers armed in parallel, static schedule for two surface types) through the
# full -m gpu suite and smoke(); final bench lines with the profiling loop OUTSIDE the timed region; what the driver's scaling
# run times (20 steps after 5 warm-up steps) on the 8-GPU shard, with and without a preheat; ncu capture of the two-type kernel
set -x
export COLUMNS=200
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rfs --tb=short --timeout 300 -p no:cacheprovider > $O/r2_15_pytest.log 2>&1
tail -6 $O/r2_15_pytest.log
timeout 120 python __graft_entry__.py smoke > $O/r2_15_smoke.log 2>&1; tail -3 $O/r2_15_smoke.log
timeout 400 python bench.py > $O/r2_15_c4_n1.json 2> $O/r2_15_c4_n1.err; cut -c1-250 $O/r2_15_c4_n1.json; tail -2 $O/r2_15_c4_n1.err
timeout 400 python bench.py --workload C5 --no-cpu-baseline > $O/r2_15_c5_n1.json 2> $O/r2_15_c5_n1.err; cut -c1-250 $O/r2_15_c5_n1.json; tail -2 $O/r2_15_c5_n1.err
B="timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-parity"
ab() { name=$1; shift; "$@" > $O/r2_15_ab_$name.json 2>>$O/r2_15_ab.err; python - $O/r2_15_ab_$name.json $name <<'EOF'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print("AB %-28s steps %4d ms/step %.4f  kernel_ms %.4f  frac %.3f  frac_step %.3f" % (sys.argv[2], d["steps"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["frac_from_ms_per_step"]))
except Exception as e:
    print("AB %-28s FAILED %s" % (sys.argv[2], e))
EOF
}
for rep in 1 2 3; do
ab shard_short_$rep         $B --workload C4 --cells 1250000 --steps 20 --warmup 5
ab shard_short_preheat_$rep $B --workload C4 --cells 1250000 --steps 20 --warmup 5 --preheat-ms 200
done
ab shard_long               $B --workload C4 --cells 1250000 --steps 2000 --warmup 50
ab c4_short_1               $B --workload C4 --steps 20 --warmup 5
ab c4_short_2               $B --workload C4 --steps 20 --warmup 5
ab c4_short_preheat         $B --workload C4 --steps 20 --warmup 5 --preheat-ms 200
ab c5_short                 $B --workload C5 --steps 20 --warmup 5
ab c5_300                   $B --workload C5 --steps 300
N="ncu --set full --clock-control none --import-source on -k regex:flux_spec_kernel -s 3 -c 1 -f"
C="$B --workload C5 --steps 3 --warmup 3 --profile-steps 0 --opt graphs=0"
$C > $O/plain_c5.log 2>&1 && $N -o $O/r2_15_prof_c5_static $C > $O/ncu_c5.log 2>&1
tail -3 $O/ncu_c5.log
ls -la $O | tail -30
