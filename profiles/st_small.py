import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import components.flux_calculator_b200 as m
from synthetic import Scenario
from oracle_py import Oracle
from tolerances import check_field
fset, staged, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sc = Scenario(fset, n=(n,n,n), S=1, bias=True)
o_in, o_out = sc.clone(); orc = Oracle(sc.n, 1); sc.apply(orc, o_in, o_out); orc.step_all(0)
g_in, g_out = sc.clone()
fc = m.FluxCalculator(sc.n, sc.S)
w = sc.apply(fc, g_in, g_out, wrap=lambda a: m.DeviceArray.from_numpy(a))
fc.set_option("staged", staged)
fc.prepare(); fc.step_all(0); fc.synchronize()
for k in g_out: w[id(g_out[k])].download(g_out[k])
for k in sorted(o_out): check_field(k[2], g_out[k], o_out[k], fset)
print(fset, "staged", staged, "n", n, "ok; exact_path_calls", fc.info("exact_path_calls"))
