set -x
for a in "--cells 1250000 --profile-stride 1" "--cells 1250000 --profile-stride 8" "--cells 1250000 --profile-stride 1000000" "--profile-stride 8" "--workload C3 --profile-stride 8" "--workload C2 --staged 2"; do
   timeout 300 python bench.py --workload C4 $a --steps 500 --warmup 10 --no-e2e --no-cpu-baseline --no-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$a', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'launches/step', d['gpu_launches']/d['steps'], r['kernel'], d['clocks'])
    else: print(l.rstrip()[:300])
"
done
