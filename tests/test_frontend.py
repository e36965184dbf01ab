"""'Next' rows 3-4: the reference's flux_calculator.nml and corrections/*.nc feed the context.
The two parsers are host code and are tested here without a GPU; the end-to-end path is in the -m gpu test below."""
import os

import numpy as np
import pytest

NML = """
! flux_calculator.nml -- written after the declarations in flux_calculator.F90:60-130
&input
  timestep = 600, num_timesteps = 144
  name_bottom_model(1) = 'MOM5_Baltic', letter_bottom_model = 'M'      ! trailing comment
  which_spec_vapor_surface_t(1,1) = 'CCLM', 'ignored-model-2'
  which_spec_vapor_surface_t(1,2) = "CCLM"
  which_spec_vapor_surface_u(1,:) = 2*'CCLM'
  which_spec_vapor_surface_v(1,1:2) = 'CCLM' , 'CCLM'
  which_flux_mass_evap(1,1) = 'MOM5'
  which_flux_mass_evap(1,2) = 'MOM5'
  which_flux_heat_latent = 'water', 9*'none', 'ice'       ! array element order: (1,1), (2,1) ... (10,1), (1,2)
  which_flux_heat_sensible(1,1) = 'MOM5'  which_flux_heat_sensible(1,2) = 'MOM5'
  which_flux_momentum(1, 1) = 'MOM5',
  which_flux_momentum(1, 2) = 'MOM5'
  which_flux_radiation_blackbody(1,1:2) = 'StBo', 'StBo'
  val_flux_t = 3*0.0, , 1.5e0
/
&correctionsctl
  init_date = 19611231
  lcorrections = .TRUE.
/
"""


@pytest.fixture()
def nml(tmp_path):
    p = tmp_path / "flux_calculator.nml"
    p.write_text(NML)
    return p


def test_namelist_parser(nml):
    import components.flux_calculator_b200 as m
    g = lambda name, i, j: m.namelist_get(nml, "input", name, (10, 10), (i, j))      # noqa: E731
    assert g("which_spec_vapor_surface_t", 1, 1) == "CCLM" and g("which_spec_vapor_surface_t", 2, 1) == "ignored-model-2"
    assert g("which_spec_vapor_surface_t", 1, 2) == "CCLM" and g("which_spec_vapor_surface_t", 1, 3) is None
    assert [g("which_spec_vapor_surface_u", 1, j) for j in (1, 2, 3)] == ["CCLM", "CCLM", None]
    assert [g("which_spec_vapor_surface_v", 1, j) for j in (1, 2, 3)] == ["CCLM", "CCLM", None]
    assert g("which_flux_heat_latent", 1, 1) == "water" and g("which_flux_heat_latent", 1, 2) == "ice"
    assert g("which_flux_heat_latent", 5, 1) == "none"
    assert g("WHICH_FLUX_MOMENTUM", 1, 2) == "MOM5"                       # names are case-insensitive
    assert m.namelist_get(nml, "input", "timestep") == "600"
    assert m.namelist_get(nml, "input", "letter_bottom_model", (10,), (1,)) == "M"
    assert [m.namelist_get(nml, "input", "val_flux_t", (100,), (k,)) for k in (1, 3, 4, 5, 6)] == ["0.0", "0.0", None, "1.5e0", None]
    assert m.namelist_get(nml, "correctionsctl", "init_date") == "19611231"
    assert m.namelist_get(nml, "correctionsctl", "lcorrections", (1,), (1,)) == ".TRUE."
    assert m.namelist_get(nml, "nosuchgroup", "x") is None


def test_namelist_errors(tmp_path):
    import components.flux_calculator_b200 as m
    bad = tmp_path / "bad.nml"
    bad.write_text("&input\n which_flux_momentum(1,1) = 'CCLM\n/\n")
    with pytest.raises(m.FluxCalcError):
        m.namelist_get(bad, "input", "which_flux_momentum", (10, 10), (1, 1))
    bad.write_text("&input\n which_flux_momentum(11,1) = 'CCLM'\n/\n")
    with pytest.raises(m.FluxCalcError):
        m.namelist_get(bad, "input", "which_flux_momentum", (10, 10), (1, 1))
    with pytest.raises(m.FluxCalcError):
        m.namelist_get(tmp_path / "missing.nml", "input", "x")


def _write_corrections(root, n, months=range(1, 13), version=1, dtype="f8", fill=-1.0e20, with_fill=True, record=False, seed=3):
    from scipy.io import netcdf_file
    os.makedirs(root / "corrections", exist_ok=True)
    rng = np.random.default_rng(seed)
    data = {}
    for mth in months:
        a = rng.normal(0.0, 1e-6, n)
        a[rng.random(n) < 0.1] = fill
        if dtype == "f4":
            a = a.astype(np.float32).astype(np.float64)
            fill_w = np.float32(fill)
        else:
            fill_w = np.float64(fill)
        f = netcdf_file(str(root / "corrections" / ("mass_evap-%02d.nc" % mth)), "w", version=version)
        if record:
            f.createDimension("time", None)
            f.createDimension("cell", n)
            other = f.createVariable("other", "f8", ("time", "cell"))      # a second record variable: records interleave
            v = f.createVariable("mass_evap", dtype, ("time", "cell"))
            v[0, :] = a
            other[0, :] = 7.0
        else:
            f.createDimension("cell", n)
            f.createVariable("lon", "f4", ("cell",))[:] = 1.0
            v = f.createVariable("mass_evap", dtype, ("cell",))
            v[:] = a
        if with_fill:
            v._FillValue = fill_w
        v.units = "kg m-2 s-1"
        f.close()
        data[mth] = (a, float(np.float64(fill_w)))
    return data


@pytest.mark.parametrize("version,dtype,record", [(1, "f8", False), (2, "f4", False), (1, "f8", True)])
def test_netcdf_classic_reader(tmp_path, version, dtype, record):
    import components.flux_calculator_b200 as m
    n = 1000
    data = _write_corrections(tmp_path, n, months=[3], version=version, dtype=dtype, record=record)
    a, fill = data[3]
    path = tmp_path / "corrections" / "mass_evap-03.nc"
    got, f = m.nc_read_var(path, "mass_evap", 0, n)
    assert np.array_equal(got, a) and f == fill
    got, _ = m.nc_read_var(path, "mass_evap", 137, 300)
    assert np.array_equal(got, a[137:437])
    with pytest.raises(OSError) as e:
        m.nc_read_var(path, "mass_evap", 900, 200)
    assert e.value.errno == 104
    with pytest.raises(OSError) as e:
        m.nc_read_var(path, "nosuchvar", 0, 1)
    assert e.value.errno == 103
    with pytest.raises(OSError) as e:
        m.nc_read_var(tmp_path / "nofile.nc", "mass_evap", 0, 1)
    assert e.value.errno == 101
    hdf = tmp_path / "h.nc"
    hdf.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(OSError) as e:
        m.nc_read_var(hdf, "mass_evap", 0, 1)
    assert e.value.errno == 102


@pytest.mark.gpu
def test_namelist_and_corrections_drive_a_step(fcmod, nml, tmp_path):
    """flux_calculator.nml + corrections/*.nc -> context -> fused step == oracle configured by hand (MOM5 set, S = 2,
    month from init_date 19611231 + one day = January); a shard (grid_offset) reads its own slice; a month without file
    stays zero with a warning; the reference's start-index quirk (App. F-8) is reproduced on request"""
    from synthetic import Scenario
    from oracle_py import Oracle
    from tolerances import check_scenario
    n_glob, off, n = 5000, 1024, 3000
    data = _write_corrections(tmp_path, n_glob, months=[m_ for m_ in range(1, 13) if m_ != 7])
    full = np.zeros((n_glob, 12))
    for mth, (a, fill) in data.items():
        full[:, mth - 1] = np.where(a == fill, 0.0, a)
    sc = Scenario("MOM5", n=(n, n, n), S=2, bias=True, init_date=19611231, offset=(off, off, off))
    for quirk in (False, True):
        want = full[off - 1:off - 1 + n] if quirk else full[off:off + n]
        sc.corrections = np.ascontiguousarray(want)
        o_in, o_out = sc.clone()
        orc = Oracle(sc.n, sc.S)
        sc.apply(orc, o_in, o_out)
        orc.step_all(86400)
        g_in, g_out = sc.clone()
        fc = fcmod.FluxCalculator(sc.n, sc.S)
        for (i, g, name), a in list(g_in.items()) + list(g_out.items()):
            fc.bind_field(i, g, name, a)
        for (i, g, name) in sc.send:
            fc.add_output_field(i, g, name)
        fc.set_distribute_shortwave(True)
        fc.configure_from_namelist(nml, bottom_model=1)          # methods + init_date + lcorrections
        warn = fc.load_corrections(tmp_path, grid_offset=off, reference_start_quirk=quirk)
        assert "mass_evap-07.nc" in warn and "Unset correction" in warn and warn.count("\n") == 1
        fc.prepare()
        fc.step_all(86400)
        fc.synchronize()
        check_scenario(sc, g_out, o_out)
        fc.close()
    # rank 0 under the quirk: start 0 is invalid -> every month unset -> no correction at all
    fc = fcmod.FluxCalculator(sc.n, sc.S)
    fc.configure_from_namelist(nml, bottom_model=1)
    warn = fc.load_corrections(tmp_path, grid_offset=0, reference_start_quirk=True)
    assert warn.count("Unset correction") == 12
    fc.close()


def test_namelist_parser_fuzz(tmp_path):
    """random assignments in every supported form (whole array, element + value list, sections, r*c repeats, null
    values, comments, either quote) against a direct model of Fortran's array-element-order semantics"""
    import random
    import components.flux_calculator_b200 as m
    rnd = random.Random(12345)
    words = ["CCLM", "MOM5", "RCO", "none", "zero", "copy", "water", "ice", "it''s", "a b"]
    for trial in range(25):
        model = {}
        lines = ["&input"]
        for _ in range(rnd.randint(1, 8)):
            form = rnd.choice(["element", "section_row", "section_col", "whole"])
            if form == "element":
                i, j = rnd.randint(1, 10), rnd.randint(1, 10)
                targets = list(range((i - 1) + 10 * (j - 1), 100))
                lhs = "a(%d,%d)" % (i, j)
            elif form == "section_row":
                i, lo, hi = rnd.randint(1, 10), rnd.randint(1, 5), rnd.randint(5, 10)
                targets = [(i - 1) + 10 * (j - 1) for j in range(lo, hi + 1)]
                lhs = "A(%d, %d:%d)" % (i, lo, hi) if rnd.random() < 0.7 else ("a(%d,:)" % i)
                if lhs.endswith(":)"):
                    targets = [(i - 1) + 10 * (j - 1) for j in range(1, 11)]
            elif form == "section_col":
                j = rnd.randint(1, 10)
                targets = [(i - 1) + 10 * (j - 1) for i in range(1, 11)]
                lhs = "a(:,%d)" % j
            else:
                targets, lhs = list(range(100)), "a"
            nvals = rnd.randint(1, min(6, len(targets)))
            parts, k = [], 0
            while k < nvals:
                kind = rnd.choice(["value", "null", "repeat"])
                w = rnd.choice(words)
                q = rnd.choice("'\"")
                lit = q + (w if q == "'" else w.replace("''", "'")) + q
                val = w.replace("''", "'")
                if kind == "value":
                    parts.append(lit)
                    model[targets[k]] = val
                    k += 1
                elif kind == "null" and parts:
                    parts.append("")
                    k += 1
                elif kind == "repeat":
                    r = rnd.randint(1, max(1, min(3, nvals - k)))
                    parts.append("%d*%s" % (r, lit))
                    for _ in range(r):
                        model[targets[k]] = val
                        k += 1
            sep = rnd.choice([", ", " , ", ","])
            lines.append("  %s = %s%s" % (lhs, sep.join(parts), rnd.choice(["", "   ! comment with 'quotes' and = signs", ","])))
        lines.append("/")
        path = tmp_path / ("f%d.nml" % trial)
        path.write_text("\n".join(lines) + "\n")
        for lin in rnd.sample(range(100), 30) + list(model)[:20]:
            got = m.namelist_get(path, "input", "a", (10, 10), (lin % 10 + 1, lin // 10 + 1))
            assert got == model.get(lin), (trial, lin, got, model.get(lin), path.read_text())
