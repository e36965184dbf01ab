python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for lib in "" _mb3 _mb4; do
 for pf in 0 592; do
  for a in "--workload C4 --diag 0" "--workload C4" "--workload C5" "--workload C3"; do
   FLUXCALC_LIB=$PWD/components/flux_calculator_b200/libfluxcalc_b200$lib.so python bench.py $a --cells 4000000 --prefetch $pf --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('lib$lib pf$pf', d['config']['workload'][:3], 'diag', d['config']['diagnostics'], 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3))
    else: print(l.rstrip()[:200])
"
  done
 done
done
