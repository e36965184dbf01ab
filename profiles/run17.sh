set -x
CMD5="python bench.py --workload C5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
$CMD5 > gpurun_out/plain17a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flux_spec_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1s2_c5_spec $CMD5 > gpurun_out/ncu17a.log 2>&1
CMD4="python bench.py --workload C4 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
$CMD4 > gpurun_out/plain17b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flux_spec_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1s2_c4_spec_final $CMD4 > gpurun_out/ncu17b.log 2>&1
$CMD4 > gpurun_out/plain17c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1s2_final.csv $CMD4 > gpurun_out/ncu17c.log 2>&1
ls -la gpurun_out/*.ncu-rep
