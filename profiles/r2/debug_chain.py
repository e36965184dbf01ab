import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import components.flux_calculator_b200 as m
from components.flux_calculator_b200 import DeviceArray
from synthetic import Scenario

def run(label, n, alt, dynopt, nsteps=6, sync_each=False):
    sc = Scenario("CCLM", n=n, S=1, bias=True, init_date=19610101)
    g_in, g_out = sc.clone()
    fc = m.FluxCalculator(sc.n, sc.S)
    wrapped = sc.apply(fc, g_in, g_out, wrap=lambda a: DeviceArray.from_numpy(a))
    if dynopt:
        fc.set_option("dyn_min_tiles", 1 << 30)
    fc.prepare()
    t0 = time.perf_counter()
    for k in range(nsteps):
        fc.step_all((86400 * 45 if k % 2 else 0) if alt else 600 * k)
        if sync_each:
            fc.synchronize()
    fc.synchronize()
    print("%-40s %.3f s for %d steps" % (label, time.perf_counter() - t0, nsteps), flush=True)
    fc.close()
    for w in wrapped.values():
        w.free()

N1 = (512 * 900 + 3, 512 * 900, 512 * 901)
run("test sizes, alternating months, dynopt", N1, True, True)
run("test sizes, same month, dynopt", N1, False, True)
run("test sizes, alternating, no dynopt", N1, True, False)
run("test sizes, same month, no dynopt", N1, False, False)
run("shard sizes, same month, no dynopt", (1250000,) * 3, False, False)
run("shard sizes, alternating, no dynopt", (1250000,) * 3, True, False)
run("test sizes, alternating, sync each", N1, True, True, sync_each=True)
