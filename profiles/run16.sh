set -x
timeout 900 python -m pytest tests/test_gpu_step_parity.py -m gpu -x -q -k "spec or surface_types or diag" 2>&1 | tail -4
V2=components/flux_calculator_b200/csrc/build_v2/libfluxcalc_t2.so
for lib in "" "$V2"; do
for a in "--workload C5" "--workload C5 --cells 1250000" "--workload C4"; do
   FLUXCALC_LIB=$lib timeout 300 python bench.py $a --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --no-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('lib=$lib $a', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3))
    else: print(l.rstrip()[:300])
"
done; done
