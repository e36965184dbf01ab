"""The oracle (and, with -m gpu, the CUDA library) against tests/golden/step_golden.json: what the REFERENCE'S OWN HOST
SOURCE sends when the interpreter of tests/golden/fortran_interp.py executes it (set-up from a namelist, prepare_*,
the nine calculators with the bias statement / 'zero' / average_across_surface_types incl. its %allocated guard, the
time loop with its send loops, do_regridding).  This closes the holes of the round-1 golden file, which re-typed
those loops."""
import numpy as np
import pytest

import step_replay
from oracle_py import Oracle

SCEN = {s["name"]: s for s in step_replay.scenarios()}
RUNNABLE = [n for n, s in SCEN.items() if "sent" in s and s["sent"]]


def test_golden_file_is_what_the_generator_says():
    import json
    with open(step_replay.GOLDEN) as f:
        d = json.load(f)
    assert "flux_calculator.F90" in d["interpreted"] and d["interpreted"]["flux_calculator.F90"] == [[340, 768], [859, 1028]]
    assert set(d["interpreted"]["sha256"]) >= {"flux_calculator_calculate.F90", "flux_calculator_prepare.F90", "flux_calculator_basic.F90"}
    assert len(RUNNABLE) >= 7


@pytest.mark.parametrize("name", RUNNABLE)
@pytest.mark.parametrize("explicit", [True, False])
def test_oracle_sends_what_the_reference_sends(name, explicit):
    """bit for bit (same libm); explicit = the %allocated flags are passed / inferred from the aliasing"""
    s = SCEN[name]
    orc = Oracle(s["grid_size"], s["num_surface_types"])
    rp = step_replay.Replay(s, orc, explicit_allocated=explicit)
    checked = [0, 0]

    def compare(put, got, key):
        ref = step_replay.arr(put["values"])
        known = ~np.isnan(ref)      # NaN = memory the reference never wrote before sending it (e.g. FICE at the first early phase)
        assert np.array_equal(got[known], ref[known]), (name, put["name"], put["time"])
        checked[0] += int(known.sum())
        checked[1] += 1
    rp.run(compare)
    assert checked[1] == len(s["sent"]) and checked[0] > 0


EXTRA = {s["name"]: s for s in step_replay.scenarios(step_replay.GOLDEN_EXTRA)}


@pytest.mark.parametrize("name", sorted(EXTRA))
@pytest.mark.parametrize("explicit", [True, False])
def test_oracle_sends_what_the_reference_sends_extra(name, explicit):
    """tests/golden/step_golden_extra.json (make_golden_step.py --set extra): RCO with two surface types, MOM5 coefficients
    per type with ice and the bias across the turn of the year, five surface types -- the interpreted reference against the
    oracle only, bit for bit (these registry shapes are not replayed through the CUDA library)"""
    s = EXTRA[name]
    assert s.get("sent"), name
    orc = Oracle(s["grid_size"], s["num_surface_types"])
    rp = step_replay.Replay(s, orc, explicit_allocated=explicit)
    n = [0, 0]

    def compare(put, got, key):
        ref = step_replay.arr(put["values"])
        known = ~np.isnan(ref)
        assert np.array_equal(got[known], ref[known]), (name, put["name"], put["time"])
        n[0] += int(known.sum())
        n[1] += 1
    rp.run(compare)
    assert n[1] == len(s["sent"]) and n[0] > 0


def test_bias_month_comes_from_the_reference_python_helper():
    """the months the interpreted reference obtained from pyfort/datetime_helpers.py == the oracle's calendar"""
    import ctypes as C
    from oracle_py import load
    lib = load()
    lib.orc_current_month.restype = C.c_int
    for s in list(SCEN.values()) + list(EXTRA.values()):
        for k, m in enumerate(s.get("months") or []):
            assert lib.orc_current_month(s["init_date"], k * s["timestep"]) == m


def test_reference_error_and_undefined_cases_are_recorded():
    assert SCEN["missing_input_stops"]["stops_in_setup"] and "AMOI VATM" in SCEN["missing_input_stops"]["log_tail"][1]
    assert "distribute_shortwave_radiation_flux" in SCEN["no_rsdr_is_undefined"]["undefined_in_reference"]


def test_allocated_flags_follow_the_aliasing_rule():
    """the inference the C ABI uses when the host does not state %allocated (a type-0 slot owns its storage iff no surface
    type >= 1 shares it) reproduces the reference's flags for every type-0 slot the send loops can average"""
    for s in SCEN.values():
        reg = s.get("registry_after_setup") or []
        for r in reg:
            if r["type"] != 0:
                continue
            shared = any(q["type"] >= 1 and q["storage"] == r["storage"] for q in reg)
            sent = any(o["type"] == 0 and o["grid"] == r["grid"] and o["var"] == r["var"] for o in s.get("output_fields", []))
            if sent:
                assert r["allocated"] == (not shared), (s["name"], r["var"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", RUNNABLE)
def test_cuda_sends_what_the_reference_sends(fcmod, name):
    from tolerances import Scales, check_field
    s = SCEN[name]
    fc = fcmod.FluxCalculator(s["grid_size"], s["num_surface_types"])
    rp = step_replay.Replay(s, fc)
    methods = {(w, int(i)): m for w, per in s["methods"].items() for i, m in per.items()}
    exact_vars = {"RBBR", "RSDR", "TSUR", "FICE", "ALBE", "PATM", "MPRE"}

    def compare(put, got, key):
        ref = step_replay.arr(put["values"])
        known = ~np.isnan(ref)
        outs = {k: v for k, v in rp.slot.items()}
        sc = Scales(rp.slot, outs, methods, s["num_surface_types"])
        scale = sc.of(key)
        scale = scale[known] if np.ndim(scale) else scale
        check_field(key[2], got[known], ref[known], exact=key[2] in exact_vars, scale=scale)
    rp.run(compare)
    fc.close()
