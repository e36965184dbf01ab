"""-m gpu: the fused / generic per-step path through the C ABI against the CPU oracle on seeded fields."""
import numpy as np
import pytest

from oracle_py import Oracle, ulp_diff
from tolerances import check_field, check_scenario

pytestmark = pytest.mark.gpu


def run_both(fcmod, sc, mode="host", t=0, force_generic=False, phases="all", chunks=None, diagnostics=0, staged=None):
    from components.flux_calculator_b200 import DeviceArray
    # oracle
    o_in, o_out = sc.clone()
    orc = Oracle(sc.n, sc.S)
    sc.apply(orc, o_in, o_out)
    if phases == "all":
        orc.step_all(t)
    else:
        orc.step_early(t)
        orc.step_normal(t)
    # CUDA
    g_in, g_out = sc.clone()
    fc = fcmod.FluxCalculator(sc.n, sc.S)
    if force_generic:
        fc.set_option("force_generic", 1)
    if chunks:
        fc.set_option("h2d_chunks", chunks)
    if staged is not None:      # 0: direct-load kernel only, 2: staged (cp.async.bulk + mbarrier) kernel whenever possible
        fc.set_option("staged", staged)
    if mode == "host":
        sc.apply(fc, g_in, g_out)
        wrapped = None
    else:
        wrapped = sc.apply(fc, g_in, g_out, wrap=lambda a: DeviceArray.from_numpy(a))
    if diagnostics:
        for g in (1, 2, 3):
            fc.set_area(g, sc.area[g])
        fc.set_option("diagnostics", int(diagnostics))
    fc.prepare()
    if phases == "all":
        fc.step_all(t)
    else:
        fc.step_early(t)
        fc.step_normal(t)
    fc.synchronize()
    if wrapped is not None:
        for d in (g_in, g_out):
            for k, a in d.items():
                wrapped[id(a)].download(a)
    return fc, o_out, g_out, o_in, g_in


def compare(sc, o_out, g_out):
    return check_scenario(sc, g_out, o_out)


@pytest.mark.parametrize("fset", ["CCLM", "MOM5", "RCO"])
@pytest.mark.parametrize("mode", ["host", "device"])
@pytest.mark.parametrize("staged", [0, 2])
def test_step_all_s1(fcmod, fset, mode, staged):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario(fset, n=(20000, 20000, 20000), S=1, bias=(fset == "MOM5"))
    fc, o_out, g_out, o_in, g_in = run_both(fcmod, sc, mode, staged=staged)
    assert fc.info("fused") == 1
    compare(sc, o_out, g_out)
    for k in o_in:      # inputs untouched
        assert np.array_equal(o_in[k], g_in[k], equal_nan=True)


@pytest.mark.parametrize("fset", ["CCLM", "MOM5", "RCO"])
def test_generic_path_matches(fcmod, fset):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario(fset, n=(5003, 4999, 5001), S=2, bias=True, averaging=True)
    fc, o_out, g_out, _, _ = run_both(fcmod, sc, "host", force_generic=True)
    assert fc.info("fused") == 0
    compare(sc, o_out, g_out)
    fc2, _, f_out, _, _ = run_both(fcmod, sc, "device")
    assert fc2.info("fused") == 1
    # fused (lock-step exp / exp(c*log x)) and generic (libdevice exp / pow) kernels: identical wherever no
    # transcendental is upstream, within the stated tolerance elsewhere
    check_scenario(sc, f_out, g_out)


@pytest.mark.parametrize("S", [2, 3, 5])
def test_surface_types_and_averaging(fcmod, S):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=(7001, 7003, 6999), S=S, bias=True, averaging=True)
    for mode in ("host", "device"):
        fc, o_out, g_out, _, _ = run_both(fcmod, sc, mode, t=40 * 86400)
        assert fc.info("fused") == 1
        compare(sc, o_out, g_out)
    # averaging itself is exact arithmetic on the per-type values
    for (i, g, name) in sc.send:
        acc = np.zeros(sc.n[g - 1])
        for t in range(1, S + 1):
            acc = acc + g_out[(t, g, name)] * sc.inputs[(t, g, "FARE")]
        assert np.array_equal(acc, g_out[(0, g, name)])


def test_split_phases_equal_fused_all(fcmod):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=(4096, 4096, 4096), S=2, bias=True, averaging=True, passthrough_avg=True)
    _, o_out, a_out, _, _ = run_both(fcmod, sc, "device", phases="all")
    _, _, s_out, _, _ = run_both(fcmod, sc, "device", phases="split")
    compare(sc, o_out, a_out)
    for k in a_out:
        assert np.array_equal(a_out[k], s_out[k], equal_nan=True), k


@pytest.mark.parametrize("n", [(0, 0, 0), (1, 1, 1), (3, 2, 1), (511, 513, 1025), (33, 0, 7), (5121, 1024, 2047)])
@pytest.mark.parametrize("staged", [0, 2])
def test_ragged_and_tiny_grids(fcmod, n, staged):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=n, S=1, bias=True)
    for mode in ("host", "device"):
        _, o_out, g_out, _, _ = run_both(fcmod, sc, mode, staged=staged)
        compare(sc, o_out, g_out)


def test_staged_and_direct_kernels_agree_bitwise(fcmod):
    """same formula templates, same arithmetic policy: the shared-memory staged kernel and the direct-load kernel
    must produce identical bits (incl. a persistent grid that wraps: 700 tiles > 2 x 148 CTAs)"""
    from components.flux_calculator_b200.synthetic import Scenario
    for fset in ("CCLM", "RCO"):
        sc = Scenario(fset, n=(358400, 358400 + 77, 358400 - 513), S=1, bias=True)
        _, o_out, d_out, _, _ = run_both(fcmod, sc, "device", staged=0)
        _, _, s_out, _, _ = run_both(fcmod, sc, "device", staged=2)
        compare(sc, o_out, d_out)
        for k in d_out:
            assert np.array_equal(d_out[k], s_out[k], equal_nan=True), k


def test_extreme_operands_take_the_exact_path(fcmod):
    """operands outside [2^-500, 2^500] (or -0 / Inf where it matters) leave the lock-step fast path: the thread
    recomputes its cells with the IEEE routines, results stay bit-exact where no transcendental is involved"""
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("RCO", n=(4096, 4096, 4096), S=1)
    before = None
    for staged in (0, 2):
        sc.inputs[(0, 2, "UATM")][5] = 1e200      # u*u overflows -> sqrt(inf)
        sc.inputs[(0, 2, "UATM")][6] = 1e-200     # u*u underflows to a tiny sqrt argument
        sc.inputs[(0, 3, "VATM")][9] = -0.0
        fc, o_out, g_out, _, _ = run_both(fcmod, sc, "device", staged=staged)
        calls = fc.info("exact_path_calls")
        assert calls > (before or 0)
        before = calls
        for k in ((1, 2, "UMOM"), (1, 3, "VMOM")):
            a, b = g_out[k], o_out[k]
            assert np.array_equal(a, b, equal_nan=True), k


def test_month_rollover_bias(fcmod):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("MOM5", n=(3000, 3000, 3000), S=1, bias=True, init_date=19611231)
    outs = []
    for t in (0, 86399, 86400, 31 * 86400 + 86400):     # Dec, Dec, Jan, Feb
        _, o_out, g_out, _, _ = run_both(fcmod, sc, "host", t=t)
        compare(sc, o_out, g_out)
        outs.append(g_out[(1, 1, "MEVA")].copy())
    assert np.array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[1], outs[2])
    assert not np.array_equal(outs[2], outs[3])


def test_chunked_host_pipeline(fcmod):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=(300001, 299999, 300003), S=1, bias=True)
    _, o_out, g1, _, _ = run_both(fcmod, sc, "host", chunks=1)
    _, _, g7, _, _ = run_both(fcmod, sc, "host", chunks=7)
    compare(sc, o_out, g1)
    for k in g1:
        assert np.array_equal(g1[k], g7[k], equal_nan=True), k


@pytest.mark.parametrize("level", [1, 2])
@pytest.mark.parametrize("staged", [0, 2])
@pytest.mark.parametrize("S,n", [(2, (10007, 10009, 10011)), (1, (200000, 150001, 99999)), (1, (3, 700, 0))])
def test_diagnostics(fcmod, level, S, n, staged):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=n, S=S, bias=True, averaging=True)
    for mode in ("device", "host"):
        fc, o_out, g_out, _, _ = run_both(fcmod, sc, mode, diagnostics=level, staged=staged)
        compare(sc, o_out, g_out)
        for (i, g, name), arr in g_out.items():
            if arr.size == 0:
                continue
            s, mn, mx = fc.diagnostics(i, g, name)
            if level == 2:      # reproduces the reference's debug "range =" lines (flux_calculator.F90:1013)
                assert mn == arr.min() and mx == arr.max(), (i, g, name)
            else:
                assert np.isnan(mn) and np.isnan(mx)
            ref = float(np.sum(sc.area[g] * arr))
            assert abs(s - ref) <= 1e-11 * float(np.sum(np.abs(sc.area[g] * arr))), (i, g, name)


def _diag_check(fc, sc, g_out, level):
    for (i, g, name), arr in g_out.items():
        if arr.size == 0:
            continue
        s, mn, mx = fc.diagnostics(i, g, name)
        if level == 2:
            assert mn == arr.min() and mx == arr.max(), (i, g, name)
        ref = float(np.sum(sc.area[g] * arr))
        assert abs(s - ref) <= 1e-11 * float(np.sum(np.abs(sc.area[g] * arr))), (i, g, name)


@pytest.mark.parametrize("fset", ["CCLM", "MOM5", "RCO"])
@pytest.mark.parametrize("level", [0, 1, 2])
def test_spec_kernel_cold_epilogue(fcmod, fset, level):
    """specialised persistent kernel: tiles with operands outside the proven range are recomputed by the cold
    epilogue (IEEE routines, from global memory) and the warp's diagnostics are rebuilt from the stored outputs;
    every other cell keeps the bits of the lock-step path (same results as the direct-load kernel)"""
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario(fset, n=(300000 + 37, 200000, 250000 + 511), S=1, bias=True)
    a_t = {"CCLM": "AMOI", "MOM5": "CMOI", "RCO": "QATM"}[fset]
    t_key = [k for k in sc.inputs if k[1] == 1 and k[2] == a_t][0]
    sc.inputs[t_key][3] = 1e-310                     # denormal operand (t grid, first tile)
    sc.inputs[t_key][123457] = 1e-310                # ... and a tile deep inside another CTA's schedule
    for idx in (5, 199999):                          # u*u + v*v denormal: sqrt argument outside the proven range
        sc.inputs[(0, 2, "UATM")][idx] = 1e-160
        sc.inputs[(0, 2, "VATM")][idx] = 1e-160
    sc.inputs[(0, 3, "UATM")][77777] = 1e-160
    sc.inputs[(0, 3, "VATM")][77777] = -1e-160
    outs = {}
    for staged in (0, 2):
        fc, o_out, g_out, _, _ = run_both(fcmod, sc, "device", diagnostics=level, staged=staged)
        compare(sc, o_out, g_out)
        if level:
            _diag_check(fc, sc, g_out, level)
        outs[staged] = g_out
    assert fc.info("exact_path_calls") > 0
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[2][k], equal_nan=True), k


def test_spec_kernel_flag_overflow(fcmod):
    """more flagged tiles per warp than the epilogue remembers individually -> every tile of that warp is redone"""
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=(512 * 296 * 17 + 512 * 40, 1024, 512), S=1, bias=True)
    t_key = [k for k in sc.inputs if k[1] == 1 and k[2] == "AMOI"][0]
    sc.inputs[t_key][::512] = 1e-310                 # one denormal operand in every tile
    fc, o_out, g_out, _, _ = run_both(fcmod, sc, "device", diagnostics=1, staged=2)
    compare(sc, o_out, g_out)
    _diag_check(fc, sc, g_out, 1)


@pytest.mark.parametrize("n", [(512, 300000, 1024), (150000, 512, 0), (1024, 1024, 400000), (152064, 152064, 152064)])
@pytest.mark.parametrize("fset", ["MOM5", "RCO"])
def test_spec_kernel_schedules(fcmod, n, fset):
    """persistent schedule corner cases: CTAs without t tiles, empty grids, t->u/v ring hand-over with few tiles,
    exactly one tile per CTA"""
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario(fset, n=n, S=1, bias=True)
    _, o_out, d_out, _, _ = run_both(fcmod, sc, "device", staged=0)
    fc, _, s_out, _, _ = run_both(fcmod, sc, "device", staged=2, diagnostics=2)
    compare(sc, o_out, s_out)
    _diag_check(fc, sc, s_out, 2)
    for k in d_out:
        assert np.array_equal(d_out[k], s_out[k], equal_nan=True), k


def test_repeated_steps_reuse_diagnostics_buffers(fcmod):
    """the in-kernel last-CTA reduction resets its counter: many steps in a row give the same diagnostics"""
    from components.flux_calculator_b200 import DeviceArray
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=(200000, 200000, 200000 + 3), S=1, bias=True)
    g_in, g_out = sc.clone()
    fc = fcmod.FluxCalculator(sc.n, sc.S)
    fc.set_option("staged", 2)
    sc.apply(fc, g_in, g_out, wrap=lambda a: DeviceArray.from_numpy(a))
    for g in (1, 2, 3):
        fc.set_area(g, sc.area[g])
    fc.set_option("diagnostics", 2)
    fc.prepare()
    first = None
    for k in range(7):
        fc.step_all(0)
        d = fc.diagnostics(1, 3, "VMOM")
        if first is None:
            first = d
        assert d == first


@pytest.mark.parametrize("staged", [0, 2])
@pytest.mark.parametrize("S", [1, 2])
def test_chunked_host_pipeline_with_diagnostics(fcmod, S, staged):
    """host-pointer pipeline in several chunks: every chunk reduces its own diagnostics vector (concurrently, on
    three streams), a combine kernel folds them in chunk order"""
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=(300001, 299999, 300003), S=S, bias=True, averaging=True)
    fc1, o_out, g1, _, _ = run_both(fcmod, sc, "host", chunks=1, diagnostics=2, staged=staged)
    fc7, _, g7, _, _ = run_both(fcmod, sc, "host", chunks=7, diagnostics=2, staged=staged)
    compare(sc, o_out, g7)
    _diag_check(fc7, sc, g7, 2)
    for k in g1:
        assert np.array_equal(g1[k], g7[k], equal_nan=True), k
        if g1[k].size:
            d1, d7 = fc1.diagnostics(*k), fc7.diagnostics(*k)
            assert d1[1:] == d7[1:] and abs(d1[0] - d7[0]) <= 1e-12 * abs(d1[0]) + 1e-300, k


@pytest.mark.parametrize("fset", ["CCLM", "MOM5", "RCO"])
@pytest.mark.parametrize("level", [0, 1, 2])
def test_spec_kernel_two_surface_types_long_schedule(fcmod, fset, level):
    """two surface types: one CTA per SM, two consumer teams sharing the ring (one barrier per (team, stage) pair);
    several tiles per team and per stage so that the barrier phases wrap, ragged remainders, averaging of the sent
    fluxes; bitwise equal to the direct-load kernel, within tolerance of the oracle"""
    from components.flux_calculator_b200.synthetic import Scenario
    n = (512 * 148 * 9 + 300, 512 * 148 * 5 + 1, 512 * 148 * 6 + 511)
    sc = Scenario(fset, n=n, S=2, bias=True, averaging=True)
    _, o_out, d_out, _, _ = run_both(fcmod, sc, "device", staged=0, diagnostics=level)
    fc, _, s_out, _, _ = run_both(fcmod, sc, "device", staged=2, diagnostics=level)
    assert fc.info("spec_kernel") == 1
    compare(sc, o_out, s_out)
    if level:
        _diag_check(fc, sc, s_out, level)
    for k in d_out:
        assert np.array_equal(d_out[k], s_out[k], equal_nan=True), k


def test_spec_kernel_two_surface_types_without_averaging(fcmod):
    from components.flux_calculator_b200.synthetic import Scenario
    sc = Scenario("CCLM", n=(200000, 100001, 99999), S=2, bias=True, averaging=False)
    _, o_out, d_out, _, _ = run_both(fcmod, sc, "device", staged=0)
    fc, _, s_out, _, _ = run_both(fcmod, sc, "device", staged=2)
    assert fc.info("spec_kernel") == 1
    compare(sc, o_out, s_out)
    for k in d_out:
        assert np.array_equal(d_out[k], s_out[k], equal_nan=True), k
