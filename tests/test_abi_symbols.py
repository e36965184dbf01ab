"""not-gpu: the C-ABI library loads and exports every symbol include/fluxcalc.h declares; the pure host
entry points (no device needed) behave; compute entry points fail LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fluxcalc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fc_[a-zA-Z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(fcmod):
    syms = declared_symbols()
    assert len(syms) >= 60
    for s in syms:
        assert hasattr(fcmod.lib, s), "libfluxcalc_b200.so does not export " + s
    assert sorted(fcmod.SIGNATURES) == syms, "ctypes table and header disagree"


def test_header_mirrors_the_14_flux_library_routines(golden):
    syms = set(declared_symbols())
    for routine in golden["routine_signatures"]:
        name = "fc_" + routine.replace("stbo", "StBo")
        assert name in syms, name


def test_header_mirrors_the_9_calculators():
    ref = open(os.path.join(ROOT, "tests", "golden", "flux_lib_golden.json")).read()
    syms = set(declared_symbols())
    for calc in ("calc_spec_vapor_surface", "calc_flux_mass_evap", "calc_flux_heat_latent", "calc_flux_heat_sensible",
                 "calc_flux_momentum_east", "calc_flux_momentum_north", "calc_flux_radiation_blackbody",
                 "distribute_shortwave_radiation_flux", "average_across_surface_types"):
        assert "fc_" + calc in syms
        if calc != "average_across_surface_types":
            assert calc in ref      # the name was found in the reference's time loop / calculate module


def test_variable_table_matches_reference_order(fcmod):
    names = ("ALBE ALBA AMOI AMOM FARE FICE PATM PSUR QATM TATM TSUR UATM VATM U10M V10M CMOM CMOI CHEA QSUR HLAT HSEN "
             "MEVA MPRE MRAI MSNO RBBR RLWD RLWU RSID RSIU RSIN RSDD RSDR UMOM VMOM").split()   # basic.F90:43-51
    assert fcmod.VARNAMES == names
    for i, n in enumerate(names, 1):
        assert fcmod.lib.fc_var_index(n.encode()) == i
        assert fcmod.lib.fc_var_name(i).decode() == n
    assert fcmod.lib.fc_var_index(b"XXXX") == 0


def test_month_function_matches_python_datetime(fcmod):
    from datetime import datetime, timedelta
    rng = np.random.default_rng(11)
    for _ in range(3000):
        y, mo, d = int(rng.integers(1850, 2200)), int(rng.integers(1, 13)), int(rng.integers(1, 29))
        secs = int(rng.integers(0, 200 * 365 * 86400))
        init = y * 10000 + mo * 100 + d
        ref = (datetime.strptime(str(init), "%Y%m%d") + timedelta(seconds=secs)).month
        assert fcmod.current_month(init, secs) == ref
    for (d, s), m in (((20000101, 0), 1), ((20000201, 29 * 86400), 3), ((19000201, 28 * 86400), 3),
                      ((19991231, 86399), 12), ((19991231, 86400), 1)):
        assert fcmod.current_month(d, s) == m
    assert fcmod.current_month(20001301, 0) == 0       # not a date


def test_shard_range_is_a_partition(fcmod):
    for n in (0, 1, 31, 512, 20_000, 1_000_003, 10**7):
        for R in (1, 2, 3, 4, 8):
            end = 0
            for r in range(R):
                off, size = fcmod.shard_range(n, r, R, 512)
                assert off == end and size >= 0
                if n // R >= 512:
                    assert off % 512 == 0              # every shard start is a tile start (4 KB) whenever the shards are that large
                end = off + size
            assert end == n
    with pytest.raises(fcmod.FluxCalcError):
        fcmod.shard_range(10, 4, 4 - 1)


def test_shard_range_reduces_to_decomp_def_apple_rule(fcmod):
    import oracle_py
    lib = oracle_py.load()
    for n, R in ((1000, 4), (10**7, 8), (77, 5)):
        for r in range(R):
            off, size = C.c_int64(), C.c_int64()
            lib.orc_decomp_apple(n, r, R, C.byref(off), C.byref(size))
            assert fcmod.shard_range(n, r, R, 1) == (off.value, size.value)     # decomp_def.F90:14-31


def test_no_cpu_fallback(fcmod):
    """on a box without a GPU every compute entry point must fail with FC_ERR_CUDA, never compute on the host"""
    if fcmod.lib.fc_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(fcmod.FluxCalcError) as e:
        fcmod.FluxCalculator((8, 8, 8), 1)
    assert e.value.code == 4 and "no CPU fallback" in e.value.message
    out = np.full(4, np.nan)
    with pytest.raises(fcmod.FluxCalcError) as e:
        fcmod.flux_library.flux_radiation_blackbody_StBo(out, np.full(4, 280.0))
    assert e.value.code == 4
    assert np.isnan(out).all()


def test_product_never_references_the_oracle():
    """the package and its C/CUDA sources must not include, import or link anything under oracle/"""
    pkg = os.path.join(ROOT, "components")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".F90", "Makefile")):
                text = open(os.path.join(dp, fn), errors="replace").read()
                assert "oracle_py" not in text and "flux_oracle" not in text and "liboracle" not in text, os.path.join(dp, fn)


def test_header_is_strict_c99_and_a_c_host_links(tmp_path):
    """the boundary is a C ABI: the header must compile as plain C (no C++-isms) and a C host must link against the
    library and call a non-compute entry point without a GPU"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "host.c"
    src.write_text('#include "fluxcalc.h"\n#include <string.h>\n'
                   'int main(void) {\n'
                   '  int64_t off = 0, size = 0;\n'
                   '  if (fc_version() != FC_VERSION) return 1;\n'
                   '  if (fc_shard_range(1000000, 3, 8, 512, &off, &size) != FC_OK) return 2;\n'
                   '  if (off != 3 * 124928 || size != 124928) return 3;\n'
                   '  if (fc_current_month(19611231, 86400) != 1) return 4;\n'
                   '  if (strcmp(fc_var_name(fc_var_index("TSUR")), "TSUR") != 0) return 5;\n'
                   '  return 0;\n}\n')
    libdir = os.path.join(root, "components", "flux_calculator_b200")
    exe = tmp_path / "host"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lfluxcalc_b200", "-Wl,-rpath," + libdir])
    assert subprocess.run([str(exe)]).returncode == 0


def _c_prototypes():
    import re
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fluxcalc.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char \*|fc_stream_t)\s*\**\s*(fc_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = [a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        protos[m.group(1)] = len(args)
    return protos


def test_fortran_shims_declare_the_boundary():
    """no Fortran compiler here, so at least: every bind(c) interface of the shim sources names an exported symbol
    of the header and has as many dummies as the C prototype has parameters; all 14 Level-1 routines, the nine
    calculators and the fused steps are there; the generated Level-1 files are current"""
    import re
    import subprocess
    fdir = os.path.join(ROOT, "components", "flux_calculator_b200", "fortran")
    protos = _c_prototypes()
    seen = {}
    for fn in sorted(os.listdir(fdir)):
        if not fn.endswith(".F90"):
            continue
        text = open(os.path.join(fdir, fn)).read()
        text = re.sub(r"&\s*\n\s*", " ", text)      # join continuation lines
        for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*(?:result\(\w+\)\s*)?bind\(c,\s*name='(\w+)'\)(?:\s*result\(\w+\))?", text, flags=re.I):
            name, args, cname = m.group(1), [a for a in m.group(2).split(",") if a.strip()], m.group(3)
            assert name == cname
            assert cname in protos, "%s: %s is not in include/fluxcalc.h" % (fn, cname)
            assert len(args) == protos[cname], "%s: %s has %d dummies, the C prototype %d parameters" % (fn, cname, len(args), protos[cname])
            seen[cname] = fn
    level1 = ["fc_spec_vapor_surface_cclm", "fc_flux_mass_evap_cclm", "fc_flux_mass_evap_mom5", "fc_flux_mass_evap_rco",
              "fc_flux_heat_latent_ice", "fc_flux_heat_latent_water", "fc_flux_heat_sensible_cclm", "fc_flux_heat_sensible_mom5",
              "fc_flux_heat_sensible_rco", "fc_flux_momentum_cclm", "fc_flux_momentum_mom5", "fc_flux_momentum_rco",
              "fc_flux_radiation_blackbody_StBo", "fc_distribute_radiation_flux"]
    level2 = ["fc_create", "fc_bind_field", "fc_set_allocated", "fc_set_method", "fc_set_corrections", "fc_add_output_field",
              "fc_prepare", "fc_calc_spec_vapor_surface", "fc_calc_flux_mass_evap", "fc_calc_flux_heat_latent",
              "fc_calc_flux_heat_sensible", "fc_calc_flux_momentum_east", "fc_calc_flux_momentum_north",
              "fc_calc_flux_radiation_blackbody", "fc_distribute_shortwave_radiation_flux", "fc_average_across_surface_types",
              "fc_step_early", "fc_step_normal", "fc_step_all", "fc_run_steps", "fc_regrid", "fc_set_regrid_matrix"]
    for n in level1 + level2:
        assert n in seen, n + " has no Fortran interface"
    # MODULE flux_library: the reference's 14 public names (flux_lib/flux_library.F90:32-45), each with a scalar and an array form
    lib = open(os.path.join(fdir, "flux_library_gpu.F90")).read()
    for n in level1:
        r = n[3:]
        assert re.search(r"public %s\b" % r, lib) and ("subroutine %s_scalar(" % r) in lib and ("subroutine %s_array(" % r) in lib
    # generated files are what the generator writes from the current header
    before = {fn: open(os.path.join(fdir, fn)).read() for fn in ("fluxcalc_level1_api.F90", "flux_library_gpu.F90")}
    subprocess.check_call([sys.executable, os.path.join(fdir, "gen_fortran_api.py")], stdout=subprocess.DEVNULL)
    for fn, text in before.items():
        assert open(os.path.join(fdir, fn)).read() == text, fn + " is stale: run gen_fortran_api.py"
    for fn in os.listdir(fdir):
        if fn.endswith(".F90"):
            for no, line in enumerate(open(os.path.join(fdir, fn)), 1):
                assert len(line.rstrip("\n")) <= 132, "%s:%d exceeds the free-form line length" % (fn, no)


def test_current_month_rejects_dates_that_do_not_exist(fcmod):
    """datetime.strptime(init_date, '%Y%m%d') (pyfort/datetime_helpers.py:7) raises on 19610231; the C routine returns 0"""
    import datetime
    lib = fcmod.lib
    for date in (19610231, 19610431, 19000229, 20230229, 19611301, 19610100, 19610132, 99, 0, -19610101, 100000101):
        try:
            datetime.datetime.strptime(str(date), "%Y%m%d")
            valid = True
        except ValueError:
            valid = False
        if date in (99, 0, -19610101, 100000101):
            valid = False      # (strptime would read '99' as year 9, month 9: not a yyyymmdd integer of the namelist)
        assert (lib.fc_current_month(date, 0) != 0) == valid, date
    for date, secs in ((20000229, 0), (20000229, 86400), (20240131, 86400 * 30), (19611231, 86399), (19611231, 86400), (19610101, -1)):
        d = datetime.datetime.strptime(str(date), "%Y%m%d") + datetime.timedelta(seconds=secs)
        assert lib.fc_current_month(date, secs) == d.month, (date, secs)


def test_shard_range_leaves_no_rank_empty(fcmod):
    """n < nranks * align: the alignment gives way (decomp_def.F90:23-30 never produces an empty rank for n >= nranks)"""
    for n, R in ((1000, 8), (4095, 8), (513, 2), (8, 8), (100000, 7), (3000, 4)):
        cover = 0
        for r in range(R):
            off, size = fcmod.shard_range(n, r, R, 512)
            assert size > 0 and off == cover, (n, R, r, off, size)
            cover += size
        assert cover == n
    assert fcmod.shard_range(10_000_000, 3, 8, 512) == (3 * 1249792, 1249792)      # large grids keep the tile alignment
