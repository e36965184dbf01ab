set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for a in "--workload C5" "--workload C5 --cells 1250000" "--workload C4" "--workload C3"; do
   timeout 300 python bench.py $a --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$a', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'launches/step', d['gpu_launches']/d['steps'], r['kernel'], r['algorithmic_bytes_per_cell'], d['parity'], d['exact_path_calls'])
    else: print(l.rstrip()[:300])
"
done
