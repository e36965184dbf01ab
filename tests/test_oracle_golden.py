"""not-gpu: pins the CPU oracle (oracle/flux_oracle.c) against
  (1) tests/golden/flux_lib_golden.json -- produced by tests/golden/make_golden.py, which interprets the
      reference's Fortran source text (formulas, call-site wiring, calculator order), and
  (2) 50-digit mpmath known-answer values (recomputed in tests/golden/kat_mpmath.json).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_py
from oracle_py import Oracle, ulp_diff

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = oracle_py.load()


def fh(x):
    return float.fromhex(x)


# Fortran dummy order of every flux_lib routine -> oracle array driver
DRIVERS = {
    "spec_vapor_surface_cclm": ("orc_v_spec_vapor_surface_cclm", 1, 3, 2),
    "flux_mass_evap_cclm": ("orc_v_flux_mass_evap_cclm", 1, 7, 3),
    "flux_mass_evap_mom5": ("orc_v_flux_mass_evap_mom5", 1, 7, 0),
    "flux_mass_evap_rco": ("orc_v_flux_mass_evap_rco", 1, 4, 0),
    "flux_heat_latent_ice": ("orc_v_flux_heat_latent_ice", 1, 1, 1),
    "flux_heat_latent_water": ("orc_v_flux_heat_latent_water", 1, 1, 1),
    "flux_heat_sensible_cclm": ("orc_v_flux_heat_sensible_cclm", 1, 8, 4),
    "flux_heat_sensible_mom5": ("orc_v_flux_heat_sensible_mom5", 1, 8, 0),
    "flux_heat_sensible_rco": ("orc_v_flux_heat_sensible_rco", 1, 4, 0),
    "flux_momentum_cclm": ("orc_v_flux_momentum_cclm", 2, 6, 2),
    "flux_momentum_mom5": ("orc_v_flux_momentum_mom5", 2, 6, 0),
    "flux_momentum_rco": ("orc_v_flux_momentum_rco", 2, 2, 0),
    "flux_radiation_blackbody_stbo": ("orc_v_flux_radiation_blackbody_StBo", 1, 1, 1),
    "distribute_radiation_flux": ("orc_v_distribute_radiation_flux", 1, 3, 0),
}


def call_driver(name, ins, opts):
    fn_name, n_out, n_in, n_opt = DRIVERS[name]
    fn = getattr(LIB, fn_name)
    n = len(ins[0])
    outs = [np.full(n, np.nan) for _ in range(n_out)]
    args = [o.ctypes.data_as(C.c_void_p) for o in outs]
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in ins]
    args += [a.ctypes.data_as(C.c_void_p) for a in keep]
    args.append(C.c_int64(n))
    for k in range(n_opt):
        args.append(C.byref(C.c_double(opts[k])) if opts is not None else None)
    fn.restype = None
    fn(*args)
    return outs


def test_golden_covers_all_14_routines(golden):
    assert sorted(golden["level0"]) == sorted(DRIVERS)
    assert len(golden["reference_files"]) == 9   # 8 formula files + the flux_library facade


@pytest.mark.parametrize("name", sorted(DRIVERS))
def test_level0_bit_exact_against_source_interpreted_reference(golden, name):
    g = golden["level0"][name]
    assert len(g["in_names"]) == DRIVERS[name][2]
    assert len(g["opt_names"]) == DRIVERS[name][3]
    for with_opt in (False, True):
        cases = [c for c in g["cases"] if (c["opt"] is not None) == with_opt]
        if not cases:
            continue
        ins = [np.array([fh(c["in"][k]) for c in cases]) for k in range(len(g["in_names"]))]
        opts = [fh(x) for x in cases[0]["opt"]] if with_opt else None
        outs = call_driver(name, ins, opts)
        for k, o in enumerate(outs):
            ref = np.array([fh(c["out"][k]) for c in cases])
            assert ulp_diff(o, ref).max() == 0, (name, k, with_opt)


def test_wiring_extracted_from_reference_matches_oracle_calls(golden):
    """the argument order the oracle's calc_* pass to the scalar routines == the reference's call sites"""
    w = golden["wiring"]
    get = lambda calc, m: [a["var"] for a in w[calc][m]["args"]]
    assert get("calc_flux_mass_evap", "CCLM") == ["MEVA", "AMOI", "PSUR", "QATM", "QSUR", "TATM", "UATM", "VATM"]
    assert get("calc_flux_mass_evap", "MOM5")[1] == "CMOI" and get("calc_flux_mass_evap", "MOM5")[5] == "TATM"
    assert get("calc_flux_heat_sensible", "CCLM") == ["HSEN", "AMOI", "PATM", "PSUR", "QATM", "TATM", "TSUR", "UATM", "VATM"]
    assert get("calc_flux_heat_sensible", "MOM5")[1] == "CHEA"
    assert get("calc_flux_momentum_east", "CCLM")[:3] == ["UMOM", "dummy", "AMOM"]
    assert get("calc_flux_momentum_north", "MOM5")[:3] == ["dummy", "VMOM", "CMOM"]
    assert get("calc_flux_mass_evap", "RCO") == ["MEVA", "QATM", "TSUR", "UATM", "VATM"]
    seq = [(s["calc"], s["grid"]) for s in golden["step_sequence"]]
    assert seq == [("calc_flux_radiation_blackbody", None), ("calc_spec_vapor_surface", 1), ("calc_spec_vapor_surface", 2),
                   ("calc_spec_vapor_surface", 3), ("calc_flux_mass_evap", None), ("calc_flux_heat_latent", None),
                   ("calc_flux_heat_sensible", None), ("calc_flux_momentum_east", 2), ("calc_flux_momentum_north", 3),
                   ("distribute_shortwave_radiation_flux", None)]


@pytest.mark.parametrize("key", ["CCLM/water", "MOM5/water", "RCO/water", "CCLM/ice"])
def test_chain_through_calc_routines_bit_exact(golden, key):
    """whole per-cell chain through the oracle's calc_* drivers == the source-interpreted chain"""
    fset, hl = key.split("/")
    cells = golden["cells"]
    n = len(cells)
    col = lambda v: np.array([fh(c[v]) for c in cells])
    orc = Oracle((n, n, n), 1)
    arrays = {}
    for g in (1, 2, 3):
        for v in ("FICE", "PSUR", "TSUR", "QATM", "TATM", "PATM", "UATM", "VATM", "AMOI", "AMOM", "CMOI", "CHEA", "CMOM"):
            arrays[(g, v)] = col(v)
            orc.bind_field(1, g, v, arrays[(g, v)])
    rsdd = col("RSDD")
    orc.bind_field(0, 1, "RSDD", rsdd)
    orc.bind_field(0, 1, "ALBA", col("ALBA"))
    orc.bind_field(1, 1, "ALBE", col("ALBE"))
    out = {}
    for g, names in ((1, ["QSUR", "MEVA", "HLAT", "HSEN", "RBBR", "RSDR"]), (2, ["QSUR", "UMOM"]), (3, ["QSUR", "VMOM"])):
        for v in names:
            out[(g, v)] = np.full(n, np.nan)
            orc.bind_field(1, g, v, out[(g, v)])
    for w in ("which_spec_vapor_surface_t", "which_spec_vapor_surface_u", "which_spec_vapor_surface_v"):
        orc.set_method(w, 1, "CCLM")
    orc.set_method("which_flux_mass_evap", 1, fset)
    orc.set_method("which_flux_heat_latent", 1, hl)
    orc.set_method("which_flux_heat_sensible", 1, fset)
    orc.set_method("which_flux_momentum", 1, fset)
    orc.set_method("which_flux_radiation_blackbody", 1, "StBo")
    orc.set_distribute_shortwave(True)
    corr = np.zeros((n, 12))
    corr[:, 0] = col("CORR")
    rows = golden["chains"][key]
    ref = lambda v: np.array([fh(r[v]) for r in rows])
    # without bias
    orc.step_all(0)
    for (g, v), r in (((1, "QSUR"), "QSUR"), ((1, "MEVA"), "MEVA_nobias"), ((1, "HLAT"), "HLAT_nobias"), ((1, "HSEN"), "HSEN"),
                      ((2, "UMOM"), "UMOM"), ((3, "VMOM"), "VMOM"), ((1, "RBBR"), "RBBR"), ((1, "RSDR"), "RSDR")):
        assert ulp_diff(out[(g, v)], ref(r)).max() == 0, (key, v)
    # with the January bias slab (calculate.F90:112-116); HLAT sees the corrected MEVA
    orc.set_corrections(corr, True, 20000101)
    orc.step_all(0)
    assert ulp_diff(out[(1, "MEVA")], ref("MEVA")).max() == 0
    assert ulp_diff(out[(1, "HLAT")], ref("HLAT")).max() == 0


def test_mpmath_known_answers():
    with open(os.path.join(ROOT, "tests", "golden", "kat_mpmath.json")) as f:
        kat = json.load(f)
    for case in kat["cases"]:
        ins = [np.array([float.fromhex(x)]) for x in case["in"]]
        outs = call_driver(case["routine"], ins, None)
        for o, ref, tol, scale in zip(outs, case["out"], case["tol_ulp"], case["scale"]):
            r = float(ref)
            # ulps of max(|result|, scale): cancellation (T_s - T_a*EF, q_s - q_a) turns an ulp of a term
            # into many ulps of a small difference
            err_ulp = abs(o[0] - r) / np.spacing(max(abs(r), scale)) if max(abs(r), scale) != 0 else abs(o[0])
            assert err_ulp <= tol, (case["routine"], case["in"], o[0], ref, err_ulp)


def test_month_function_known_answers():
    kat = [((20000101, 0), 1), ((20000101, 31 * 86400), 2), ((20000201, 29 * 86400), 3), ((19000201, 28 * 86400), 3),
           ((19991231, 86399), 12), ((19991231, 86400), 1), ((19610101, 0), 1), ((21000228, 86400), 3)]
    for (d, s), m in kat:
        assert LIB.orc_current_month(d, s) == m


def test_month_function_against_python_datetime():
    from datetime import datetime, timedelta
    rng = np.random.default_rng(7)
    for _ in range(3000):
        y, mo, d = int(rng.integers(1850, 2200)), int(rng.integers(1, 13)), int(rng.integers(1, 29))
        secs = int(rng.integers(0, 200 * 365 * 86400))
        init = y * 10000 + mo * 100 + d
        ref = (datetime.strptime(str(init), "%Y%m%d") + timedelta(seconds=secs)).month   # datetime_helpers.py:7-8
        assert LIB.orc_current_month(init, secs) == ref


def test_regrid_sequential_order():
    rng = np.random.default_rng(3)
    n_src, n_dst, nnz = 50, 40, 400
    s = rng.integers(1, n_src + 1, nnz).astype(np.int32)
    d = rng.integers(1, n_dst + 1, nnz).astype(np.int32)
    w = rng.random(nnz)
    src = rng.random(n_src)
    dst = np.full(n_dst, np.nan)
    LIB.orc_v_regrid.restype = None
    LIB.orc_v_regrid(dst.ctypes.data_as(C.c_void_p), C.c_int64(n_dst), src.ctypes.data_as(C.c_void_p), C.c_int64(nnz),
                     s.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
    ref = np.zeros(n_dst)
    for k in range(nnz):                       # basic.F90:483-486
        ref[d[k] - 1] = ref[d[k] - 1] + src[s[k] - 1] * w[k]
    assert np.array_equal(dst, ref)


def test_decomp_apple_rule():
    for n, R in ((100, 4), (1003, 8), (7, 8), (10**7, 8)):
        tot, prev_end = 0, 0
        for r in range(R):
            off, size = C.c_int64(), C.c_int64()
            LIB.orc_decomp_apple(n, r, R, C.byref(off), C.byref(size))
            assert off.value == prev_end == r * (n // R)          # decomp_def.F90:24,28
            prev_end = off.value + size.value
            tot += size.value
        assert tot == n and prev_end == n
