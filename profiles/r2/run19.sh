# round 2, call 19 (1 GPU): the library rebuilt from a clean tree at HEAD: smoke() and the driver's N=1 arguments
set -x
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 60 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_19_c4_20steps.json 2> gpurun_out/r2_19.err; cut -c1-330 gpurun_out/r2_19_c4_20steps.json; tail -2 gpurun_out/r2_19.err
