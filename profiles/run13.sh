set -x
for cells in 200000 1000000 3000000 10000000; do
  timeout 120 python bench.py --workload C5 --cells $cells --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -2 | cut -c1-300
done
for cells in 200000 1000000; do
  timeout 120 python bench.py --workload C5 --cells $cells --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --profile-stride 1 2>&1 | tail -2 | cut -c1-300
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --workload C5 --cells 1000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>&1 | grep -v "^{" | head -60
