set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -2 gpurun_out/bench_final_n1.err; cat gpurun_out/bench_final_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_ref.json 2>gpurun_out/bench_final_ref.err; cat gpurun_out/bench_final_ref.json | cut -c1-600
