"""-m gpu: Level 1 of the C ABI (the 14 flux_library routines in array form) and the nine unfused calculators
against the oracle and the committed golden vectors."""
import ctypes as C

import numpy as np
import pytest

import oracle_py
from oracle_py import Oracle, ulp_diff
from tolerances import check_field, check_scenario, routine_scale

pytestmark = pytest.mark.gpu

# routine -> (n_out, input field makers, optional constants, oracle driver, transcendental?)
ROUTINES = {
    "spec_vapor_surface_cclm": (1, ["FICE", "PSUR", "TSUR"], [287.058, 461.495], True),
    "flux_mass_evap_cclm": (1, ["AMOI", "PSUR", "QATM", "QSUR", "TATM", "UATM", "VATM"], [0.02, 287.058, 461.495], False),
    "flux_mass_evap_mom5": (1, ["CMOI", "PSUR", "QATM", "QSUR", "TATM", "UATM", "VATM"], [], False),
    "flux_mass_evap_rco": (1, ["QATM", "TSUR", "UATM", "VATM"], [], True),
    "flux_heat_latent_ice": (1, ["MEVA"], [2.834e6], False),
    "flux_heat_latent_water": (1, ["MEVA"], [2.5008e6], False),
    "flux_heat_sensible_cclm": (1, ["AMOI", "PATM", "PSUR", "QATM", "TATM", "TSUR", "UATM", "VATM"],
                                [1004.64, 0.02, 287.058, 461.495], True),
    "flux_heat_sensible_mom5": (1, ["CHEA", "PATM", "PSUR", "QATM", "TATM", "TSUR", "UATM", "VATM"], [], True),
    "flux_heat_sensible_rco": (1, ["TATM", "TSUR", "UATM", "VATM"], [], False),
    "flux_momentum_cclm": (2, ["AMOM", "PSUR", "QSUR", "TSUR", "UATM", "VATM"], [287.058, 461.495], False),
    "flux_momentum_mom5": (2, ["CMOM", "PSUR", "QSUR", "TSUR", "UATM", "VATM"], [], False),
    "flux_momentum_rco": (2, ["UATM", "VATM"], [], False),
    "flux_radiation_blackbody_StBo": (1, ["TSUR"], [5.670374419e-8], False),
    "distribute_radiation_flux": (1, ["RSDD", "ALBA", "ALBE"], [], False),
}


def fields(n, seed=1):
    from synthetic import make_field
    f = {}
    f["TSUR"] = make_field("TSUR", n, seed=seed)
    f["PSUR"] = make_field("PSUR", n, seed=seed)
    f["TATM"] = make_field("TATM", n, seed=seed, base=f["TSUR"])
    f["PATM"] = make_field("PATM", n, seed=seed, base=f["PSUR"])
    for v in ("QATM", "UATM", "VATM", "AMOI", "AMOM", "CMOI", "CHEA", "CMOM", "RSDD", "ALBA", "ALBE"):
        f[v] = make_field(v, n, seed=seed)
    f["FICE"] = (np.arange(n) % 2).astype(np.float64)
    f["QSUR"] = make_field("QATM", n, seed=seed + 1) * 1.3
    f["MEVA"] = 1e-4 * (make_field("FARE", n, seed=seed) - 0.3)
    return f


def oracle_call(name, ins, opts, n_out, n):
    lib = oracle_py.load()
    fn = getattr(lib, "orc_v_" + name)
    fn.restype = None
    outs = [np.full(n, np.nan) for _ in range(n_out)]
    args = [o.ctypes.data_as(C.c_void_p) for o in outs] + [a.ctypes.data_as(C.c_void_p) for a in ins] + [C.c_int64(n)]
    nopt = len(ROUTINES[name][2])
    args += [(C.byref(C.c_double(x)) if opts else None) for x in ROUTINES[name][2]][:nopt]
    fn(*args)
    return outs


@pytest.mark.parametrize("name", sorted(ROUTINES))
@pytest.mark.parametrize("mode", ["host", "device"])
@pytest.mark.parametrize("with_opt", [False, True])
def test_flux_library_routine(fcmod, name, mode, with_opt):
    n_out, in_names, optvals, transcendental = ROUTINES[name]
    if with_opt and not optvals:
        pytest.skip("routine has no OPTIONAL constants")
    n = 10007
    f = fields(n)
    ins = [f[v] for v in in_names]
    ref = oracle_call(name, ins, with_opt, n_out, n)
    fn = getattr(fcmod.flux_library, name)
    if mode == "host":
        outs = [np.full(n, np.nan) for _ in range(n_out)]
        fn(*outs, *ins, *(optvals if with_opt else []))
        got = outs
    else:
        d_out = [fcmod.DeviceArray(n) for _ in range(n_out)]
        d_in = [fcmod.DeviceArray.from_numpy(a) for a in ins]
        fn(*d_out, *d_in, *(optvals if with_opt else []))
        got = [d.download() for d in d_out]
    for g, r in zip(got, ref):
        if transcendental:
            check_field(name, g, r, exact=False, scale=routine_scale(name, ins))
        else:
            assert ulp_diff(g, r).max() == 0, name      # IEEE ops only: bit exact


def test_momentum_dummy_results_may_be_null(fcmod):
    n = 513
    f = fields(n)
    e = np.full(n, np.nan)
    nn = np.full(n, np.nan)
    both = [np.full(n, np.nan), np.full(n, np.nan)]
    L = fcmod.flux_library
    L.flux_momentum_cclm(both[0], both[1], f["AMOM"], f["PSUR"], f["QSUR"], f["TSUR"], f["UATM"], f["VATM"])
    L.flux_momentum_cclm(e, None, f["AMOM"], f["PSUR"], f["QSUR"], f["TSUR"], f["UATM"], f["VATM"])   # east on the u grid
    L.flux_momentum_cclm(None, nn, f["AMOM"], f["PSUR"], f["QSUR"], f["TSUR"], f["UATM"], f["VATM"])  # north on the v grid
    assert np.array_equal(e, both[0]) and np.array_equal(nn, both[1])


def test_golden_vectors_through_the_c_abi(fcmod, golden):
    """committed, source-interpreted reference vectors fed straight through the Level-1 ABI"""
    fh = float.fromhex
    for name, g in golden["level0"].items():
        fn = getattr(fcmod.flux_library, name.replace("stbo", "StBo"))
        cases = [c for c in g["cases"] if c["opt"] is None]
        ins = [np.array([fh(c["in"][k]) for c in cases]) for k in range(len(g["in_names"]))]
        n_out = len(cases[0]["out"])
        outs = [np.full(len(cases), np.nan) for _ in range(n_out)]
        fn(*outs, *ins)
        for k, o in enumerate(outs):
            ref = np.array([fh(c["out"][k]) for c in cases])
            check_field(name, o, ref, exact=False, scale=routine_scale(name, ins))


def test_properties(fcmod):
    """SURVEY App. E properties that do not need the oracle"""
    L = fcmod.flux_library
    n = 4096
    f = fields(n, seed=5)
    a, b = np.full(n, np.nan), np.full(n, np.nan)
    # MOM5 routine == CCLM routine bit for bit
    L.flux_mass_evap_cclm(a, f["AMOI"], f["PSUR"], f["QATM"], f["QSUR"], f["TATM"], f["UATM"], f["VATM"])
    L.flux_mass_evap_mom5(b, f["AMOI"], f["PSUR"], f["QATM"], f["QSUR"], f["TATM"], f["UATM"], f["VATM"])
    assert np.array_equal(a, b)
    L.flux_heat_sensible_cclm(a, f["AMOI"], f["PATM"], f["PSUR"], f["QATM"], f["TATM"], f["TSUR"], f["UATM"], f["VATM"])
    L.flux_heat_sensible_mom5(b, f["AMOI"], f["PATM"], f["PSUR"], f["QATM"], f["TATM"], f["TSUR"], f["UATM"], f["VATM"])
    assert np.array_equal(a, b)
    # calm cells: momentum exactly +-0, RCO fluxes 0, CCLM evaporation uses u_min
    z = np.zeros(n)
    L.flux_momentum_cclm(a, b, f["AMOM"], f["PSUR"], f["QSUR"], f["TSUR"], z, z)
    assert np.all(a == 0.0) and np.all(b == 0.0)
    L.flux_mass_evap_rco(a, f["QATM"], f["TSUR"], z, z)
    assert np.all(a == 0.0)
    small = np.full(n, 0.003)
    L.flux_mass_evap_cclm(a, f["AMOI"], f["PSUR"], f["QATM"], f["QSUR"], f["TATM"], small, np.full(n, 0.004))
    L.flux_mass_evap_cclm(b, f["AMOI"], f["PSUR"], f["QATM"], f["QSUR"], f["TATM"], np.full(n, 0.01), z)
    assert np.array_equal(a, b)
    # TATM == TSUR: RCO sensible heat is exactly 0 (stable branch); vel == 11 exactly: linear drag branch
    L.flux_heat_sensible_rco(a, f["TSUR"], f["TSUR"], f["UATM"], f["VATM"])
    assert np.all(a == 0.0)
    L.flux_momentum_rco(a, b, np.full(n, 11.0), z)
    assert np.all(a == -(1.225 * (0.49E-03 + 0.065E-03 * 11.0) * 11.0 * 11.0))
    # HLAT / MEVA == L exactly; blackbody is sigma*(T*T)*(T*T)
    L.flux_heat_latent_water(a, f["MEVA"])
    assert np.array_equal(a, f["MEVA"] * 2.501e6)
    L.flux_heat_latent_ice(a, f["MEVA"])
    assert np.array_equal(a, f["MEVA"] * 2.835e6)
    L.flux_radiation_blackbody_StBo(a, f["TSUR"])
    t2 = f["TSUR"] * f["TSUR"]
    assert np.array_equal(a, 5.67e-8 * (t2 * t2))
    # fractional ice interpolates the Magnus coefficients linearly: between pure water and pure ice
    qw, qi, qf = np.empty(n), np.empty(n), np.empty(n)
    L.spec_vapor_surface_cclm(qw, z, f["PSUR"], f["TSUR"])
    L.spec_vapor_surface_cclm(qi, np.ones(n), f["PSUR"], f["TSUR"])
    L.spec_vapor_surface_cclm(qf, np.full(n, 0.5), f["PSUR"], f["TSUR"])
    assert np.all((qf >= np.minimum(qw, qi)) & (qf <= np.maximum(qw, qi)))


def test_unfused_calculators_match_oracle_pass_by_pass(fcmod):
    """the nine calc_* entry points, called one by one like the reference's time loop"""
    from synthetic import Scenario
    sc = Scenario("MOM5", n=(3001, 2999, 3003), S=2, bias=True, averaging=True)
    o_in, o_out = sc.clone()
    orc = Oracle(sc.n, sc.S)
    sc.apply(orc, o_in, o_out)
    g_in, g_out = sc.clone()
    fc = fcmod.FluxCalculator(sc.n, sc.S)
    sc.apply(fc, g_in, g_out)
    for tgt in (orc, fc):
        tgt.set_time(70 * 86400)
        tgt.calc_flux_radiation_blackbody()
        for g in (1, 2, 3):
            tgt.calc_spec_vapor_surface(g)
        tgt.calc_flux_mass_evap()
        tgt.calc_flux_heat_latent()
        tgt.calc_flux_heat_sensible()
        tgt.calc_flux_momentum_east(2)
        tgt.calc_flux_momentum_north(3)
        tgt.distribute_shortwave_radiation_flux()
        for (i, g, name) in sc.send:
            tgt.average_across_surface_types(g, name)
    check_scenario(sc, g_out, o_out)


def test_copy_and_zero_methods_follow_reference_aliasing(fcmod):
    """'copy' aliases surface type 1's array (prepare.F90:36-38); the bias is then added once per surface type to
    the SAME array (calculate.F90:112-116) -- the generic path reproduces it, the fused path declines"""
    from synthetic import Scenario
    sc = Scenario("CCLM", n=(2001, 2001, 2001), S=3, bias=True, averaging=True)
    sc.methods[("which_flux_mass_evap", 2)] = "copy"
    sc.methods[("which_flux_heat_sensible", 3)] = "zero"
    sc.methods[("which_flux_radiation_blackbody", 2)] = "copy"
    res = []
    for cls in ("oracle", "cuda"):
        ins, outs = sc.clone()
        outs[(2, 1, "MEVA")] = outs[(1, 1, "MEVA")]
        outs[(2, 1, "RBBR")] = outs[(1, 1, "RBBR")]
        tgt = Oracle(sc.n, sc.S) if cls == "oracle" else fcmod.FluxCalculator(sc.n, sc.S)
        sc.apply(tgt, ins, outs)
        tgt.step_all(0)
        if cls == "cuda":
            assert tgt.info("fused") == 0
        res.append(outs)
    check_scenario(sc, res[1], res[0])
    assert np.all(res[1][(3, 1, "HSEN")] == 0.0)


def test_validation_errors(fcmod):
    fc = fcmod.FluxCalculator((16, 16, 16), 1)
    with pytest.raises(fcmod.FluxCalcError) as e:
        fc.set_method("which_flux_mass_evap", 1, "COARE")
    assert e.value.code == 2 and "is not known" in e.value.message          # prepare.F90:110-113
    with pytest.raises(fcmod.FluxCalcError) as e:
        fc.set_method("which_flux_heat_latent", 1, "CCLM")
    assert e.value.code == 2
    fc.set_method("which_flux_mass_evap", 1, "CCLM")
    fc.bind_field(1, 1, "MEVA", np.zeros(16))
    fc.bind_field(1, 1, "PSUR", np.zeros(16))
    with pytest.raises(fcmod.FluxCalcError) as e:
        fc.prepare()
    assert e.value.code == 3
    for v in ("AMOI", "QATM", "QSUR", "TATM", "UATM", "VATM"):
        assert v in e.value.message                                         # prepare.F90:90-96
    assert "PSUR" not in e.value.message.split("variables:")[1]
    with pytest.raises(fcmod.FluxCalcError) as e:
        fc.bind_field(1, 1, "TSUR", np.zeros(15))
    assert e.value.code == 1


def test_strict_validation_reproduces_prepare_quirks(fcmod):
    """prepare_flux_mass_evap('RCO') tests QSUR and reports TSUR (prepare.F90:107)"""
    n = 8
    fc = fcmod.FluxCalculator((n, n, n), 1)
    fc.set_method("which_flux_mass_evap", 1, "RCO")
    for v in ("MEVA", "QATM", "TSUR", "UATM", "VATM"):
        fc.bind_field(1, 1, v, np.ones(n))
    fc.prepare(strict=False)                       # what the formula really reads is all there
    with pytest.raises(fcmod.FluxCalcError) as e:
        fc.prepare(strict=True)
    assert e.value.code == 3 and "TSUR" in e.value.message
    fc.bind_field(1, 1, "QSUR", np.ones(n))
    fc.prepare(strict=True)


def test_regrid_matches_reference_accumulation_order(fcmod):
    rng = np.random.default_rng(5)
    nt, nu = 3000, 2900
    fc = fcmod.FluxCalculator((nt, nu, nu), 1)
    nnz = 20000
    s = rng.integers(1, nu + 1, nnz).astype(np.int32)
    d = rng.integers(1, nt + 1, nnz).astype(np.int32)
    w = rng.random(nnz)
    src = rng.random(nu)
    fc.set_regrid_matrix(0, s, d, w)               # u -> t
    dst = np.full(nt, np.nan)
    fc.regrid(0, dst, src)
    ref = np.zeros(nt)
    for k in range(nnz):                           # basic.F90:483-486
        ref[d[k] - 1] = ref[d[k] - 1] + src[s[k] - 1] * w[k]
    assert np.array_equal(dst, ref)
    dd = fcmod.DeviceArray(nt)
    fc.regrid(0, dd, fcmod.DeviceArray.from_numpy(src))
    fc.synchronize()
    assert np.array_equal(dd.download(), ref)
