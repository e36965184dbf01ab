# round 2, call 10: month slabs 16-byte aligned (odd grids kept the fast kernel only in January), full suite with per-test timeouts,
# regridding and generic-kernel measurements, bench lines of the small workloads
set -x
export COLUMNS=200
timeout 900 python -m pytest tests -m gpu -q -rf --tb=short --timeout 120 --durations=8 -p no:cacheprovider > gpurun_out/r2_10_pytest.log 2>&1
tail -25 gpurun_out/r2_10_pytest.log
timeout 200 python profiles/regrid_bench.py > gpurun_out/r2_10_regrid.json 2>>gpurun_out/r2_10.err; cat gpurun_out/r2_10_regrid.json
B="timeout 300 python bench.py --no-parity"
run() { name=$1; shift; "$@" > gpurun_out/r2_10_$name.json 2>>gpurun_out/r2_10.err; cut -c1-160 gpurun_out/r2_10_$name.json; }
run s3 $B --workload S3 --no-e2e --no-cpu-baseline
run c2 $B --workload C2 --steps 2000 --warmup 50
run c3 $B --workload C3 --steps 1000 --warmup 50
run shard $B --workload C4 --cells 1250000 --steps 1000 --warmup 50 --no-e2e --no-cpu-baseline
tail -5 gpurun_out/r2_10.err
