set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err; tail -2 gpurun_out/bench_final2_n1.err; cut -c1-400 gpurun_out/bench_final2_n1.json
python bench.py --workload C5 --no-cpu-baseline > gpurun_out/bench_final2_c5.json 2>/dev/null; cut -c1-300 gpurun_out/bench_final2_c5.json
CMD5="python bench.py --workload C5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
$CMD5 > gpurun_out/plain24.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flux_spec_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1s2_c5_spec_final $CMD5 > gpurun_out/ncu24.log 2>&1
