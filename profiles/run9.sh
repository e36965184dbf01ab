set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1s2.json 2> gpurun_out/bench_r1s2.err; tail -3 gpurun_out/bench_r1s2.err; cat gpurun_out/bench_r1s2.json
python bench.py --steps 100 --warmup 5 --workload C4 --cells 1251456 --no-e2e --no-cpu-baseline --no-parity | cut -c1-400
