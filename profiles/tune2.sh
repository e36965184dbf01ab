timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for a in "--workload C4 --diag 0" "--workload C4" "--workload C3 --cells 10000000" "--workload C5" "--workload C2"; do
   timeout 300 python bench.py $a --steps 30 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['config']['workload'][:3], 'diag', d['config']['diagnostics'], 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'parity', d['parity'])
    else: print(l.rstrip()[:200])
"
done
