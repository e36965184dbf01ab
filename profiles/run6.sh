# round-1 session-2: parity incl. the new spec-kernel tests, bench sweep, ncu launch list + full capture of the spec kernel
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for a in "--workload C4" "--workload C4 --diag 0" "--workload C3 --cells 10000000" "--workload C4 --cells 1250000" "--workload C4 --cells 1250000 --diag 0" "--workload C2"; do
   timeout 300 python bench.py $a --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$a', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'launches/step', d['gpu_launches']/d['steps'])
    else: print(l.rstrip()[:300])
"
done
CMD="python bench.py --workload C4 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
$CMD > gpurun_out/plain6.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1s2.csv $CMD > gpurun_out/ncu6a.log 2>&1
$CMD > gpurun_out/plain6b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flux_spec_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1s2_c4_spec $CMD > gpurun_out/ncu6b.log 2>&1
tail -3 gpurun_out/ncu6a.log gpurun_out/ncu6b.log
