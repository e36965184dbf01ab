for cells in 1250000 2500000 5000000; do for st in 1 0; do
   timeout 300 python bench.py --workload C4 --cells $cells --staged $st --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('cells $cells staged $st', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'launches/step', d['gpu_launches']/d['steps'])
    else: print(l.rstrip()[:200])
"
done; done
