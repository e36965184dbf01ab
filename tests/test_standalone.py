"""fc_create_from_namelist / fc_namelist_registry: the library's own set-up from flux_calculator.nml against what the
REFERENCE'S main program builds from the same namelist (tests/golden/step_golden.json: flux_calculator.F90 STEP 1.4-1.7
executed by the interpreter of tests/golden/) -- every slot, which slots alias which, %allocated, constants, regridding
flags, the received and the sent field lists with their OASIS names and early flags, and the reference's stop messages."""
import numpy as np
import pytest

import step_replay

SCEN = {s["name"]: s for s in step_replay.scenarios()}
EXTRA = {s["name"]: s for s in step_replay.scenarios(step_replay.GOLDEN_EXTRA)}      # set-up only (no device needed)


def _nml(tmp_path, s):
    p = tmp_path / "flux_calculator.nml"
    p.write_text(s["namelist"])
    return p


@pytest.mark.parametrize("name", [n for n, s in SCEN.items() if "registry_after_setup" in s] + sorted(EXTRA))
def test_registry_matches_the_reference(fcmod, tmp_path, name):
    s = SCEN.get(name) or EXTRA[name]
    got = fcmod.namelist_registry(_nml(tmp_path, s), 1, s["grid_size"])
    assert got["num_surface_types"] == s["num_surface_types"]
    ref = {(r["type"], r["grid"], r["var"]): r for r in s["registry_after_setup"]}
    mine = {(r["type"], r["grid"], r["var"]): r for r in got["registry"]}
    assert set(mine) == set(ref), (sorted(set(mine) - set(ref)), sorted(set(ref) - set(mine)))
    # the same partition into storage groups (aliases), the same flags
    ga, gb = {}, {}
    for k in ref:
        ga.setdefault(ref[k]["storage"], set()).add(k)
        gb.setdefault(mine[k]["storage"], set()).add(k)
        assert mine[k]["allocated"] == ref[k]["allocated"], k
        assert sorted(mine[k]["regrid_to"]) == sorted(ref[k].get("regrid_to", [])), k
        vals = step_replay.arr(ref[k]["values"])
        if mine[k]["fill"] is None:
            assert np.isnan(vals).all(), k      # nothing written by the set-up
        else:
            assert np.all(vals == mine[k]["fill"]), k      # val_* constant / default value of a flux nobody computes
    assert sorted(map(sorted, ga.values())) == sorted(map(sorted, gb.values()))
    strip = lambda lst: [{k: f[k] for k in ("name", "grid", "early", "type", "var")} for f in lst]      # noqa: E731
    assert strip(got["input_fields"]) == strip(s["input_fields"])
    assert strip(got["output_fields"]) == strip(s["output_fields"])


def test_reference_stop_is_an_error_with_its_message(fcmod, tmp_path):
    s = SCEN["missing_input_stops"]
    with pytest.raises(fcmod.FluxCalcError) as e:
        fcmod.namelist_registry(_nml(tmp_path, s), 1, s["grid_size"])
    assert e.value.code == 3      # FC_ERR_MISSING
    assert "Error calculating MEVA" in e.value.message and "For method CCLM we are lacking the following variables:" in e.value.message \
        and e.value.message.rstrip().endswith("AMOI VATM")      # the list prepare_flux_mass_evap accumulates (prepare.F90:93-99)


def test_overfull_send_list_is_refused(fcmod, tmp_path):
    """a uniform flux sent to the atmosphere and to the bottom model: the reference counts one output field and adds two
    (flux_calculator.F90:655-660 vs :694-712) -- out of bounds there, an error here"""
    s = SCEN["cclm_s1_bias"]
    text = s["namelist"].replace("send_to_atmos_t(1) = .FALSE.", "send_to_atmos_t(1) = .TRUE.")
    for j in range(2, 6):
        text = text.replace("send_to_atmos_t(%d) = .FALSE." % j, "send_to_atmos_t(%d) = .TRUE." % j)
    p = tmp_path / "f.nml"
    p.write_text(text)
    with pytest.raises(fcmod.FluxCalcError) as e:
        fcmod.namelist_registry(p, 1, s["grid_size"])
    assert "more fields than the reference allocates" in e.value.message


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n, s in SCEN.items() if s.get("sent")])
def test_namelist_drives_the_time_loop(fcmod, tmp_path, name):
    """flux_calculator.nml -> context -> the reference's time loop (receive, regrid, early phase, send, receive, regrid,
    normal phase, send): everything the reference sends, within the stated tolerance"""
    from tolerances import Scales, check_field
    s = SCEN[name]
    fc = fcmod.NamelistCalculator(_nml(tmp_path, s), s["grid_size"], bottom_model=1)
    assert [f["name"] for f in fc.inputs] == [f["name"] for f in s["input_fields"]]
    if "corrections_month_major" in s:
        corr = np.array([[float.fromhex(x) for x in row] for row in s["corrections_month_major"]])
        fc.set_corrections(np.ascontiguousarray(corr.T), True, s["init_date"])
    for which, m in s.get("regrid_matrices", {}).items():
        fc.set_regrid_matrix({"u_to_t": 0, "v_to_t": 1, "t_to_u": 2, "t_to_v": 3}[which], m["src_index"], m["dst_index"], step_replay.arr(m["weight"]))
    received = {(r["name"], r["grid"], r["time"]): step_replay.arr(r["values"]) for r in s["received"]}
    methods = {(w, int(i)): m for w, per in s["methods"].items() for i, m in per.items()}
    slots = {(r["type"], r["grid"], r["var"]): fc.array(r["type"], r["grid"], r["var"]) for r in s["registry_after_setup"]}
    exact_vars = {"RBBR", "RSDR", "TSUR", "FICE", "ALBE", "PATM", "MPRE"}
    puts, k = s["sent"], 0
    for n in range(s["num_timesteps"]):
        t = n * s["timestep"]
        for early in (True, False):
            for f in fc.inputs:
                if f["early"] == early:
                    f["array"][:] = received[(f["name"], f["grid"], t)]
            (fc.step_early if early else fc.step_normal)(t)
            for g in (1, 2, 3):
                for o in fc.outputs:
                    if o["grid"] == g and o["early"] == early:
                        put = puts[k]
                        k += 1
                        assert put["name"] == o["name"] and put["time"] == t
                        ref = step_replay.arr(put["values"])
                        known = ~np.isnan(ref)
                        key = (o["type"], g, o["var"])
                        scale = Scales(slots, slots, methods, s["num_surface_types"]).of(key)
                        scale = scale[known] if np.ndim(scale) else scale
                        check_field(o["var"], o["array"][known], ref[known], exact=o["var"] in exact_vars, scale=scale)
    assert k == len(puts)
    fc.close()


def test_compact_namelist_syntax_builds_the_same_registry(fcmod, tmp_path):
    """the namelist written the way a person writes it -- array sections with value lists, repeat counts, several assignments
    per line, a one-line &correctionsctl group -- against the one-assignment-per-line text of the golden
    scenario with the same content: identical registry, field lists and OASIS names"""
    s = SCEN["cclm_s1_bias"]
    compact = """&input
  timestep=43200, num_timesteps=3, name_atmos_model='CCLM', name_bottom_model(1)='MOM5', letter_bottom_model(1)='M',
  num_tasks_per_model(1)=1,
  name_atmos_var_t(1:9)='PSUR','PATM','QATM','TATM','UATM','VATM','RSDD','ALBA','AMOI',
  name_atmos_var_u(1:5)='PSUR','UATM','VATM','AMOM','TATM',  name_atmos_var_v(1:5)='PSUR','UATM','VATM','AMOM','TATM',
  name_bottom_var_t(1,1,1:3)='TSUR','FICE','ALBE', val_bottom_var_t(1,1,2)=0.0,
  name_bottom_var_u(1,1,1:2)='TSUR','FICE', val_bottom_var_u(1,1,2)=0.0,
  name_bottom_var_v(1,1,1:2)='TSUR','FICE', val_bottom_var_v(1,1,2)=0.0,
  which_spec_vapor_surface_t(1,1)='CCLM', which_spec_vapor_surface_u(1,1)='CCLM', which_spec_vapor_surface_v(1,1)='CCLM',
  which_flux_mass_evap(1,1)='CCLM', which_flux_heat_latent(1,1)='water', which_flux_heat_sensible(1,1)='CCLM',
  which_flux_momentum(1,1)='CCLM', which_flux_radiation_blackbody(1,1)='StBo',   ! one surface type: open water
  name_send_t(1:5)='MEVA','HLAT','HSEN','RBBR','RSDR', name_send_u(1)='UMOM', name_send_v(1)='VMOM',
  send_uniform_t(1,1:5)=5*.TRUE., send_uniform_u(1,1)=.TRUE., send_uniform_v(1,1)=.TRUE.,
  send_to_atmos_t(1:5)=5*.FALSE., send_to_bottom_u(1,1)=.FALSE., send_to_bottom_v(1,1)=.FALSE.
/
&correctionsctl  init_date=19610131, lcorrections(1)=.TRUE. /
"""
    p = tmp_path / "compact.nml"
    p.write_text(compact)
    a = fcmod.namelist_registry(p, 1, s["grid_size"])
    b = fcmod.namelist_registry(_nml(tmp_path, s), 1, s["grid_size"])
    key = lambda r: (r["type"], r["grid"], r["var"])      # noqa: E731
    assert sorted(map(key, a["registry"])) == sorted(map(key, b["registry"]))
    ra, rb = {key(r): r for r in a["registry"]}, {key(r): r for r in b["registry"]}
    for k in ra:
        assert (ra[k]["allocated"], ra[k]["fill"], sorted(ra[k]["regrid_to"])) == (rb[k]["allocated"], rb[k]["fill"], sorted(rb[k]["regrid_to"])), k
    groups = lambda reg: sorted(sorted(k for k in reg if reg[k]["storage"] == st) for st in {r["storage"] for r in reg.values()})      # noqa: E731
    assert groups(ra) == groups(rb)
    strip = lambda lst: [{k: f[k] for k in ("name", "grid", "early", "type", "var")} for f in lst]      # noqa: E731
    assert strip(a["input_fields"]) == strip(b["input_fields"]) and strip(a["output_fields"]) == strip(b["output_fields"])
