set -x
timeout 1200 python -m pytest tests/test_gpu_step_parity.py -m gpu -x -q -k "diag or two_surface or cold or surface_types" 2>&1 | tail -4
for a in "--workload C5 --diag 1" "--workload C5 --diag 1 --staged 0"; do
   timeout 300 python bench.py $a --steps 100 --warmup 10 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$a', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), r['kernel'], d['parity']['max_rel_err'], d['diagnostics_sample'])
    else: print(l.rstrip()[:300])
"
done
