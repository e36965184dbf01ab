for ch in 0 16 8; do
python - <<PY
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, components.flux_calculator_b200 as m
from components.flux_calculator_b200 import pinned_empty
sc = bench.build_scenario("C4")
g_in, g_out = sc.clone()
memo = {}
def pin(a):
    if id(a) not in memo:
        p = pinned_empty(a.size); p[:] = a; memo[id(a)] = p
    return memo[id(a)]
h_in = {k: pin(a) for k, a in g_in.items()}; h_out = {k: pin(a) for k, a in g_out.items()}
fh = m.FluxCalculator(sc.n, sc.S)
sc.apply(fh, h_in, h_out)
for g in (1, 2, 3): fh.set_area(g, sc.area[g])
fh.set_option("diagnostics", 1)
if $ch: fh.set_option("h2d_chunks", $ch)
fh.prepare(); fh.step_all(0); fh.synchronize()
t0 = time.perf_counter()
for k in range(5): fh.step_all(600 * (k + 1))
fh.synchronize()
print("chunks $ch (0 = auto): %.2f ms/step" % ((time.perf_counter() - t0) / 5 * 1e3))
PY
done
timeout 300 python -m pytest tests/test_gpu_step_parity.py -m gpu -x -q -k "chunked" 2>&1 | tail -2
