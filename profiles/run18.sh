# A/B on one box: two-surface-type geometry (2 teams x 8 warps x 512-cell tiles | 4 teams x 4 warps x 256-cell tiles) x producer (one lane | one lane per slot)
for rep in 1 2 3; do
for v in t2p0 t2p1 t4p0 t4p1; do
for a in "--workload C5" "--workload C5 --cells 1250000" "--workload C4"; do
   FLUXCALC_LIB=components/flux_calculator_b200/csrc/build_v2/lib_$v.so timeout 300 python bench.py $a --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --no-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$v rep$rep $a', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3))
    else: print(l.rstrip()[:300])
"
done; done; done
