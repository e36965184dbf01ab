for st in 1 0; do
for a in "--workload C4 --diag 0" "--workload C4" "--workload C3 --cells 10000000"; do
   timeout 300 python bench.py $a --staged $st --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('staged $st', d['config']['workload'][:3], 'diag', d['config']['diagnostics'], 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3))
    else: print(l.rstrip()[:200])
"
done; done
