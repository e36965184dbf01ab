"""Per-field error table of the CUDA path against the CPU oracle on 2 x 10^6 cells per grid and formula set:
cells not bit-equal, max ulp distance, max |err|, max relative error, and the worst error in units of the stated
tolerance (tests/tolerances.py).  Fields with no transcendental upstream must be bit-exact; the others carry the
documented libm-vs-CUDA exp/pow last-place differences, amplified where the flux is a small difference of large terms.
Run on the GPU box:  python profiles/parity_table.py > gpurun_out/parity_table.md"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import components.flux_calculator_b200 as m  # noqa: E402
from components.flux_calculator_b200 import DeviceArray  # noqa: E402
from synthetic import Scenario  # noqa: E402
from oracle_py import Oracle, ulp_diff  # noqa: E402
from tolerances import K_ULP, ULP, Scales  # noqa: E402

N = 2_000_000
print("| set | S | field | cells differing | max ulp | max abs err | max rel err | worst err / tolerance |")
print("|---|---|---|---|---|---|---|---|")
for fset, S in (("CCLM", 1), ("MOM5", 1), ("RCO", 1), ("CCLM", 2)):
    sc = Scenario(fset, n=(N, N, N), S=S, bias=True, averaging=True)
    o_in, o_out = sc.clone()
    orc = Oracle(sc.n, sc.S)
    sc.apply(orc, o_in, o_out)
    orc.step_all(0)
    g_in, g_out = sc.clone()
    fc = m.FluxCalculator(sc.n, sc.S)
    wrapped = sc.apply(fc, g_in, g_out, wrap=lambda a: DeviceArray.from_numpy(a))
    fc.prepare()
    assert fc.info("spec_kernel") == 1
    fc.step_all(0)
    fc.synchronize()
    scales = Scales(sc.inputs, o_out, sc.methods, sc.S)
    for k in sorted(o_out):
        got = wrapped[id(g_out[k])].download()
        ref = o_out[k]
        err = np.abs(got - ref)
        tol = K_ULP * ULP * (np.abs(ref) + scales.of(k))      # per cell: tests/tolerances.py
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.where(ref != 0, err / np.abs(ref), 0.0)
            frac = np.where(tol > 0, err / tol, np.where(err > 0, np.inf, 0.0))
        print("| %s | %d | %s type %d grid %s | %d | %d | %.3g | %.3g | %.3g |" % (
            fset, S, k[2], k[0], "tuv"[k[1] - 1], int((got != ref).sum()), int(ulp_diff(got, ref).max()), err.max(), rel.max(), frac.max()))
    fc.close()
    for w in wrapped.values():
        w.free()
