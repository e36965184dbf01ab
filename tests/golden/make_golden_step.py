#!/usr/bin/env python3
"""Generate tests/golden/step_golden.json by EXECUTING THE REFERENCE'S HOST SOURCE with tests/golden/fortran_interp.py.

What is interpreted, straight from /root/reference/src (nothing below is re-typed here):
  * flux_calculator.F90 STEP 1.4 - 1.7 (lines 340-768): allocation of the received fields, 'val_*' constants, pointer
    distribution of atmosphere fields over the surface types, prepare_regridding, the prepare_* calls, add_output_field;
  * flux_calculator_basic.F90 (allocate_localvar, init_localvar, distribute_input_field, add_input_field,
    add_output_field, prepare_regridding, do_regridding, init_varname_idx, nullify_localvars),
    flux_calculator_prepare.F90 (all), flux_calculator_calculate.F90 (all nine calculators incl. the bias statement
    :112-116, 'zero' :79, average_across_surface_types :368-385 with its %allocated guard);
  * flux_calculator.F90 STEP 2 (lines 859-1028): the time loop with its send loops and averaging triggers;
  * the scalar flux_lib routines through make_golden.FluxLib (the statement-level interpreter of round 1);
  * pyfort/datetime_helpers.py is IMPORTED and called for the month (as the reference does through call_python).
Stubbed (no arithmetic on the path): OASIS (oasis_get fills the received fields from this script's seeded data,
oasis_put records what is sent), MPI, WRITE.

Output per scenario: the namelist as text (what a user would put into flux_calculator.nml), grid sizes, received data per
time step, corrections, regridding matrices -- and what the reference sends (every oasis_put of every step, bit exact as
C99 hex floats), the final registry (which slots share storage, %allocated flags) or the error it stops with.

Usage: python tests/golden/make_golden_step.py [--ref /root/reference]      (build container only)
"""
import argparse
import hashlib
import importlib.util
import json
import math
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from fortran_interp import FArr, FObj, FortranStop, Interp, Ref      # noqa: E402
import make_golden      # noqa: E402  (FluxLib: the scalar routines of flux_lib)

MAIN_SETUP = (340, 768)      # STEP 1.4 .. 1.7 of flux_calculator.F90
MAIN_LOOP = (859, 1028)      # STEP 2


def hx(v):
    return float(v).hex()


class Reference:
    """the interpreted reference, one instance per scenario"""

    def __init__(self, ref):
        self.ref = ref
        src = os.path.join(ref, "src")
        self.I = I = Interp()
        self.files = []
        for f in ("flux_calculator_basic.F90", "bias_corrections.F90", "flux_calculator_prepare.F90",
                  "flux_calculator_calculate.F90"):
            I.load_module(os.path.join(src, f))
            self.files.append(f)
        self.main = I.load_program(os.path.join(src, "flux_calculator.F90"))
        self.files.append("flux_calculator.F90")
        self.lib = make_golden.FluxLib(ref)
        self.log = []
        self.puts = []
        self.state = {}          # call_python's STATE
        self.months = []         # current_month of every call of get_current_date
        self.feed = None         # callable(name, time) -> list of floats
        spec = importlib.util.spec_from_file_location("datetime_helpers", os.path.join(src, "pyfort", "datetime_helpers.py"))
        self.dth = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(self.dth)
        self._externals()
        I.call("init_varname_idx", [], I.glob)      # flux_calculator.F90:183
        self.scope = I.new_scope(self.main)
        self.scope.update(oasis_ok=0, oasis_recvd=3, comp_id=1, comp_name="flxcalc")

    def _externals(self):
        I, ext = self.I, self.I.externals
        for name, r in self.lib.routines.items():
            def f(interp, refs, name=name, r=r):
                actual = [None if (r.intent.get(formal) == "out") else float(ref.get()) for formal, ref in zip(r.args, refs)]
                res = self.lib.call(name, actual)
                for formal, ref in zip(r.args, refs):
                    if r.intent.get(formal) == "out":
                        ref.set(res[formal])
            ext[name] = f
        ext["mpi_barrier"] = lambda i, refs: None
        ext["flush"] = lambda i, refs: None

        def stop(interp, refs):
            raise FortranStop("mpi_finalize / oasis_abort")
        ext["mpi_finalize"] = stop
        ext["oasis_abort"] = stop

        def py_set(interp, refs):
            import numpy as np
            self.state[str(refs[0].get())] = np.array([int(refs[1].get())], dtype=np.int32)
        def py_get(interp, refs):
            arr = refs[1].get()
            v = self.state[str(refs[0].get())]
            arr.data[:] = [int(v)] * len(arr.data)
        def py_call(interp, refs):
            assert str(refs[0].get()) == "datetime_helpers"
            getattr(self.dth, str(refs[1].get()))(self.state)
            self.months.append(int(self.state["current_month"]))
        ext["set"], ext["get"], ext["call_function"] = py_set, py_get, py_call

        def oasis_get(interp, refs):
            fid, t, field = refs[0].get(), refs[1].get(), refs[2].get()
            name, grid = self.in_names[fid]
            vals = [float(x) for x in self.feed(name, grid, int(t))]
            assert len(vals) == len(field.data)
            field.data[:] = vals
            refs[3].set(0)
        def oasis_put(interp, refs):
            fid, t, field = refs[0].get(), refs[1].get(), refs[2].get()
            name, grid = self.out_names[fid]
            self.puts.append({"time": int(t), "name": name.strip(), "grid": grid, "values": [hx(x) for x in field.data]})
            refs[3].set(0)
        ext["oasis_get"], ext["oasis_put"] = oasis_get, oasis_put

    # ---- namelist ------------------------------------------------------------------------------
    def apply_namelist(self, nml):
        """nml: {name: value | {index tuple (1-based): value}}"""
        for name, val in nml.items():
            tgt = self.scope[name.lower()] if name.lower() in self.scope else self.I.glob[name.lower()]
            if isinstance(tgt, FArr):
                assert isinstance(val, dict), name
                for idx, v in val.items():
                    idx = idx if isinstance(idx, tuple) else (idx,)
                    tgt.data[tgt.flat(idx)] = float(v) if isinstance(tgt.data[0], float) else v
            else:
                if name.lower() in self.scope:
                    self.scope[name.lower()] = val
                else:
                    self.I.glob[name.lower()] = val

    def setup(self, grid_size, bottom_model=1, letter="M"):
        sc = self.scope
        sc["my_bottom_model"] = bottom_model          # STEP 1.3 (rank -> model), not interpreted
        sc["my_bottom_letter"] = letter
        sc["grid_size"].data[:] = list(grid_size)
        # the reference never nullifies local_field(0,:) (flux_calculator.F90:351 starts at 1): treated as disassociated
        self.I.run_main_range(sc, *MAIN_SETUP)
        # oasis_def_var (STEP 1.9) hands out the ids
        self.in_names, self.out_names = {}, {}
        n_in, n_out = sc["num_input_fields"], sc["num_output_fields"]
        for j in range(1, n_in + 1):
            f = sc["input_field"].data[j - 1]
            f.c["id"] = j
            self.in_names[j] = (f.c["name"], f.c["which_grid"])
        for j in range(1, n_out + 1):
            f = sc["output_field"].data[j - 1]
            f.c["id"] = 1000 + j
            self.out_names[1000 + j] = (f.c["name"], f.c["which_grid"])

    def set_corrections(self, init_date, corr):
        """bias_corrections.F90:26-33 module state; corr[month-1][j]"""
        g = self.I.glob
        g["init_date"] = int(init_date)
        g["lcorrections"].data[0] = True
        n = len(corr[0])
        arr = FArr([(1, 1), (1, 12), (1, n)], 0.0)
        for m in range(12):
            for j in range(n):
                arr.data[arr.flat((1, m + 1, j + 1))] = float(corr[m][j])
        g["corrections"] = arr

    def set_matrix(self, which, src, dst, w):
        """regrid_<which>_matrix (flux_calculator_io.F90:118-198 reads these from a file)"""
        M = self.scope["regrid_%s_matrix" % which]
        M.c["num_elements"] = len(w)
        for comp, vals, conv in (("src_index", src, int), ("dst_index", dst, int), ("weight", w, float)):
            a = FArr([(1, len(vals))], 0)
            a.data[:] = [conv(x) for x in vals]
            M.c[comp].c["field"] = a
            M.c[comp].c["allocated"] = True

    def time_loop(self, num_timesteps, timestep):
        sc = self.scope
        sc["num_timesteps"], sc["timestep"] = int(num_timesteps), int(timestep)
        for w in ("u_to_t", "v_to_t", "t_to_u", "t_to_v"):
            M = sc["regrid_%s_matrix" % w]
            if not isinstance(M.c.get("num_elements"), int) or M.c["src_index"].c["field"] is None:
                M.c["num_elements"] = 0
        self.I.run_main_range(sc, *MAIN_LOOP)

    # ---- inspection ----------------------------------------------------------------------------
    def registry(self):
        """every associated slot of local_field: storage group (aliases share one), %allocated"""
        lf = self.scope["local_field"]
        names = self.I.glob["varnames"].data
        groups, out = {}, []
        for i in range(0, 11):
            for g in (1, 2, 3):
                o = lf.data[lf.flat((i, g))]
                for k, v in enumerate(o.c["var"].data):
                    arr = v.c["field"]
                    if arr is None:
                        continue
                    gid = groups.setdefault(id(arr), len(groups))
                    out.append({"type": i, "grid": g, "var": names[k], "storage": gid, "allocated": bool(v.c["allocated"]),
                                "values": [hx(x) for x in arr.data]})
                    for tg, comp in ((1, "put_to_t_grid"), (2, "put_to_u_grid"), (3, "put_to_v_grid")):
                        if v.c[comp] is True:
                            out[-1].setdefault("regrid_to", []).append(tg)
        return out

    def io_lists(self):
        sc = self.scope
        names = self.I.glob["varnames"].data
        def lst(which, n):
            res = []
            for j in range(n):
                f = sc[which].data[j].c
                res.append({"name": f["name"].strip(), "grid": f["which_grid"], "early": bool(f["early"]),
                            "type": f["surface_type"], "var": names[f["idx"] - 1]})
            return res
        return lst("input_field", sc["num_input_fields"]), lst("output_field", sc["num_output_fields"])


# --------------------------------------------------------------------------------------------
# namelist text (what the user writes) from the same dict
# --------------------------------------------------------------------------------------------
def nml_text(nml, extra=""):
    lines = ["&input"]
    for name, val in nml.items():
        if isinstance(val, dict):
            for idx, v in val.items():
                idx = idx if isinstance(idx, tuple) else (idx,)
                lines.append("  %s(%s) = %s" % (name, ",".join(str(i) for i in idx), fmt(v)))
        else:
            lines.append("  %s = %s" % (name, fmt(val)))
    lines.append("/")
    return "\n".join(lines) + "\n" + extra


def fmt(v):
    if isinstance(v, bool):
        return ".TRUE." if v else ".FALSE."
    if isinstance(v, str):
        return "'%s'" % v
    if isinstance(v, float):
        return repr(v)
    return str(v)


# --------------------------------------------------------------------------------------------
# seeded input data (SURVEY 8(d) distributions)
# --------------------------------------------------------------------------------------------
def field_values(rng, var, n, ice=False):
    U = lambda a, b: [rng.uniform(a, b) for _ in range(n)]      # noqa: E731
    if var == "TSUR":
        return U(243.15, 273.15) if ice else U(271.35, 303.15)
    if var == "TATM":
        return U(268.0, 305.0)
    if var == "PSUR":
        return U(9.8e4, 1.04e5)
    if var == "PATM":
        return U(9.65e4, 9.79e4)
    if var == "QATM":
        return U(1e-3, 1.5e-2)
    if var in ("UATM", "VATM"):
        v = [max(-35.0, min(35.0, rng.gauss(0, 6))) for _ in range(n)]
        if n > 3:
            v[1] = 0.0                      # calm cell (both components are zeroed at index 1)
            v[2] = 11.0 if var == "UATM" else 0.0      # RCO drag threshold exactly
        return v
    if var in ("AMOI", "AMOM", "CMOI", "CHEA", "CMOM"):
        return U(8e-4, 2.5e-3)
    if var == "FICE":
        return [1.0 if ice else 0.0] * n
    if var == "FARE":
        return U(0.0, 1.0)
    if var == "RSDD":
        return [-x for x in U(0.0, 900.0)]
    if var in ("ALBA", "ALBE"):
        return U(0.05, 0.8)
    if var == "QSUR":
        return U(2e-3, 2e-2)
    return U(-1.0, 1.0)


class Feeder:
    """oasis_get: values per (field name, time); FARE of the last surface type completes the others to 1"""

    def __init__(self, seed, sizes, S):
        self.seed, self.sizes, self.S = seed, sizes, S
        self.cache = {}

    def __call__(self, name, grid, t):
        key = (name, grid, t)
        if key not in self.cache:
            var, st = name[2:6], int(name[6:8])
            n = self.sizes[grid - 1]
            rng = random.Random("%d/%s/%d/%d" % (self.seed, name, grid, t))
            vals = field_values(rng, var, n, ice=st >= 2)
            if var == "FARE" and self.S >= 2 and st == self.S:
                vals = [1.0] * n
                for i in range(1, self.S):
                    other = self("%sFARE%02d" % (name[:2], i), grid, t)
                    vals = [a - b / (self.S - 1) * 1.0 for a, b in zip(vals, other)]
            self.cache[key] = vals
        return self.cache[key]


# --------------------------------------------------------------------------------------------
# scenarios
# --------------------------------------------------------------------------------------------
def names(lst, first=1):
    return {first + k: v for k, v in enumerate(lst)}


def bottom(model, per_type):
    """{type: [names]} -> {(model, type, j): name}"""
    out = {}
    for i, lst in per_type.items():
        for j, v in enumerate(lst, 1):
            out[(model, i, j)] = v
    return out


def methods(model, per_type):
    return {(model, i): m for i, m in per_type.items()}


def scenario_defs():
    S = []
    atm_t = ["PSUR", "PATM", "QATM", "TATM", "UATM", "VATM", "RSDD", "ALBA"]
    # 1: CCLM set, one surface type, bias over a month boundary (C1/C2 shape)
    S.append(dict(
        name="cclm_s1_bias", uniform=True, sizes=(20, 18, 19), steps=3, timestep=43200, init_date=19610131, bias=True,
        nml={
            "name_atmos_var_t": names(atm_t + ["AMOI"]), "name_atmos_var_u": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_atmos_var_v": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_bottom_var_t": bottom(1, {1: ["TSUR", "FICE", "ALBE"]}),
            "name_bottom_var_u": bottom(1, {1: ["TSUR", "FICE"]}), "name_bottom_var_v": bottom(1, {1: ["TSUR", "FICE"]}),
            "val_bottom_var_t": {(1, 1, 2): 0.0}, "val_bottom_var_u": {(1, 1, 2): 0.0}, "val_bottom_var_v": {(1, 1, 2): 0.0},
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM"}), "which_spec_vapor_surface_u": methods(1, {1: "CCLM"}),
            "which_spec_vapor_surface_v": methods(1, {1: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "CCLM"}), "which_flux_heat_latent": methods(1, {1: "water"}),
            "which_flux_heat_sensible": methods(1, {1: "CCLM"}), "which_flux_momentum": methods(1, {1: "CCLM"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR"]), "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # 2: MOM5 coefficients from the ocean model (bottom fields, early), bias
    S.append(dict(
        name="mom5_s1_bias", uniform=True, sizes=(17, 17, 17), steps=2, timestep=600, init_date=19611231, bias=True, t0_note="month 12",
        nml={
            "name_atmos_var_t": names(atm_t), "name_atmos_var_u": names(["PSUR", "UATM", "VATM", "TATM"]),
            "name_atmos_var_v": names(["PSUR", "UATM", "VATM", "TATM"]),
            "name_bottom_var_t": bottom(1, {1: ["TSUR", "FICE", "ALBE", "CMOI", "CHEA"]}),
            "name_bottom_var_u": bottom(1, {1: ["TSUR", "FICE", "CMOM"]}), "name_bottom_var_v": bottom(1, {1: ["TSUR", "FICE", "CMOM"]}),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM"}), "which_spec_vapor_surface_u": methods(1, {1: "CCLM"}),
            "which_spec_vapor_surface_v": methods(1, {1: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "MOM5"}), "which_flux_heat_latent": methods(1, {1: "water"}),
            "which_flux_heat_sensible": methods(1, {1: "MOM5"}), "which_flux_momentum": methods(1, {1: "MOM5"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR"]), "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # 3: RCO set (QSUR on t is required by prepare, App. F-1)
    S.append(dict(
        name="rco_s1", uniform=True, sizes=(16, 16, 16), steps=1, timestep=600, init_date=19610101, bias=False,
        nml={
            "name_atmos_var_t": names(atm_t), "name_atmos_var_u": names(["UATM", "VATM"]), "name_atmos_var_v": names(["UATM", "VATM"]),
            "name_bottom_var_t": bottom(1, {1: ["TSUR", "FICE", "ALBE"]}),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "RCO"}), "which_flux_heat_latent": methods(1, {1: "water"}),
            "which_flux_heat_sensible": methods(1, {1: "RCO"}), "which_flux_momentum": methods(1, {1: "RCO"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR"]), "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # 4: open water + ice, area-fraction averages of the sent fluxes AND of pass-through variables (C5 shape), bias
    two = {1: ["TSUR", "FICE", "ALBE", "FARE"], 2: ["TSUR", "FICE", "ALBE", "FARE"]}
    twouv = {1: ["TSUR", "FICE", "FARE"], 2: ["TSUR", "FICE", "FARE"]}
    S.append(dict(
        name="cclm_s2_avg_bias", sizes=(21, 21, 21), steps=2, timestep=600, init_date=19610228, bias=True,
        nml={
            "name_atmos_var_t": names(atm_t + ["AMOI"]), "name_atmos_var_u": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_atmos_var_v": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_bottom_var_t": bottom(1, two), "name_bottom_var_u": bottom(1, twouv), "name_bottom_var_v": bottom(1, twouv),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM", 2: "CCLM"}), "which_spec_vapor_surface_u": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_spec_vapor_surface_v": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "CCLM", 2: "CCLM"}), "which_flux_heat_latent": methods(1, {1: "water", 2: "ice"}),
            "which_flux_heat_sensible": methods(1, {1: "CCLM", 2: "CCLM"}), "which_flux_momentum": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo", 2: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR", "TSUR", "FICE", "ALBE"]),
            "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # 5: three surface types with method mixes: 'zero', 'copy', 'none'; a flux nobody computes is sent with its default value
    three = {1: ["TSUR", "FICE", "ALBE", "FARE"], 2: ["TSUR", "FICE", "ALBE", "FARE"], 3: ["TSUR", "FICE", "ALBE", "FARE"]}
    threeuv = {1: ["TSUR", "FICE", "FARE"], 2: ["TSUR", "FICE", "FARE"], 3: ["TSUR", "FICE", "FARE"]}
    S.append(dict(
        name="mixed_s3_zero_copy", sizes=(15, 15, 15), steps=2, timestep=600, init_date=19610615, bias=True,
        nml={
            "name_atmos_var_t": names(atm_t + ["AMOI"]), "name_atmos_var_u": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_atmos_var_v": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_bottom_var_t": bottom(1, three), "name_bottom_var_u": bottom(1, threeuv), "name_bottom_var_v": bottom(1, threeuv),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM", 2: "CCLM", 3: "copy"}),
            "which_spec_vapor_surface_u": methods(1, {1: "CCLM", 2: "CCLM", 3: "CCLM"}),
            "which_spec_vapor_surface_v": methods(1, {1: "CCLM", 2: "CCLM", 3: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "CCLM", 2: "zero", 3: "CCLM"}),
            "which_flux_heat_latent": methods(1, {1: "water", 2: "ice", 3: "zero"}),
            "which_flux_heat_sensible": methods(1, {1: "CCLM", 2: "RCO", 3: "copy"}),
            "which_flux_momentum": methods(1, {1: "CCLM", 2: "RCO", 3: "zero"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo", 2: "StBo", 3: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR", "MPRE"]), "val_flux_t": {6: -1.5e-5},
            "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # 6: an atmosphere field listed as sent: on surface type 0 with two types the reference averages it onto itself
    S.append(dict(
        name="atmos_field_sent_s2", sizes=(12, 12, 12), steps=1, timestep=600, init_date=19610101, bias=False,
        nml={
            "name_atmos_var_t": names(atm_t + ["AMOI"]), "name_atmos_var_u": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_atmos_var_v": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_bottom_var_t": bottom(1, two), "name_bottom_var_u": bottom(1, twouv), "name_bottom_var_v": bottom(1, twouv),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM", 2: "CCLM"}), "which_spec_vapor_surface_u": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_spec_vapor_surface_v": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "CCLM", 2: "CCLM"}), "which_flux_heat_latent": methods(1, {1: "water", 2: "ice"}),
            "which_flux_heat_sensible": methods(1, {1: "CCLM", 2: "CCLM"}), "which_flux_momentum": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo", 2: "StBo"}),
            "name_send_t": names(["HSEN", "PATM", "MEVA", "RSDR"]), "send_uniform_t": {(1, 3): True},
            "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # 7: transfer coefficient and surface state received on the t grid only, regridded t -> u / t -> v by the sparse
    #    matrices (do_regridding), QSUR computed on u / v from the regridded fields
    S.append(dict(
        name="regrid_t_to_uv", uniform=True, sizes=(14, 11, 12), steps=2, timestep=600, init_date=19610101, bias=False, regrid=True,
        nml={
            "name_atmos_var_t": names(atm_t + ["AMOI", "AMOM"]), "name_atmos_var_u": names(["UATM", "VATM", "TATM"]),
            "name_atmos_var_v": names(["UATM", "VATM", "TATM"]),
            "name_bottom_var_t": bottom(1, {1: ["TSUR", "FICE", "ALBE"]}),
            "regrid_t_to_u": bottom(1, {1: ["TSUR", "FICE"]}) | {(1, 1, 3): "PSUR", (1, 1, 4): "AMOM"},
            "regrid_t_to_v": bottom(1, {1: ["TSUR", "FICE"]}) | {(1, 1, 3): "PSUR", (1, 1, 4): "AMOM"},
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM"}), "which_spec_vapor_surface_u": methods(1, {1: "CCLM"}),
            "which_spec_vapor_surface_v": methods(1, {1: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "CCLM"}), "which_flux_heat_latent": methods(1, {1: "water"}),
            "which_flux_heat_sensible": methods(1, {1: "CCLM"}), "which_flux_momentum": methods(1, {1: "CCLM"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR"]), "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # 7b: nobody lists RSDR: distribute_shortwave_radiation_flux (flux_calculator.F90:991, unconditional) writes through a
    #     disassociated pointer -- undefined in the reference; the new ABI skips the routine instead
    S.append(dict(
        name="no_rsdr_is_undefined", uniform=True, sizes=(9, 9, 9), steps=1, timestep=600, init_date=19610101, bias=False,
        nml={
            "name_atmos_var_t": names(atm_t + ["AMOI"]), "name_atmos_var_u": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_atmos_var_v": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_bottom_var_t": bottom(1, {1: ["TSUR", "FICE", "ALBE"]}),
            "name_bottom_var_u": bottom(1, {1: ["TSUR", "FICE"]}), "name_bottom_var_v": bottom(1, {1: ["TSUR", "FICE"]}),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM"}), "which_spec_vapor_surface_u": methods(1, {1: "CCLM"}),
            "which_spec_vapor_surface_v": methods(1, {1: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "CCLM"}), "which_flux_heat_latent": methods(1, {1: "water"}),
            "which_flux_heat_sensible": methods(1, {1: "CCLM"}), "which_flux_momentum": methods(1, {1: "CCLM"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR"]), "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # 8: a method whose inputs are missing: the reference stops in prepare_* (error behaviour of the boundary)
    S.append(dict(
        name="missing_input_stops", sizes=(8, 8, 8), steps=1, timestep=600, init_date=19610101, bias=False,
        nml={
            "name_atmos_var_t": names(["PSUR", "QATM", "TATM", "UATM"]),
            "name_bottom_var_t": bottom(1, {1: ["TSUR", "FICE"]}),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM"}), "which_flux_mass_evap": methods(1, {1: "CCLM"}),
            "name_send_t": names(["MEVA"]),
        }))
    return S


def scenario_defs_extra():
    """More formula-set x surface-type combinations of the reference's host code, for the ORACLE only (CPU tests): they widen
    what pins the oracle without adding untried registry shapes to the GPU suite (tests/golden/step_golden_extra.json)."""
    S = []
    atm_t = ["PSUR", "PATM", "QATM", "TATM", "UATM", "VATM", "RSDD", "ALBA"]
    two = {1: ["TSUR", "FICE", "ALBE", "FARE"], 2: ["TSUR", "FICE", "ALBE", "FARE"]}
    # e1: RCO set (Meier et al. 1999), open water + ice, area-fraction averages of every sent flux
    S.append(dict(
        name="rco_s2_avg", sizes=(13, 13, 13), steps=2, timestep=600, init_date=19610301, bias=False,
        nml={
            "name_atmos_var_t": names(atm_t), "name_atmos_var_u": names(["UATM", "VATM"]), "name_atmos_var_v": names(["UATM", "VATM"]),
            "name_bottom_var_t": bottom(1, two),
            "name_bottom_var_u": bottom(1, {1: ["FARE"], 2: ["FARE"]}), "name_bottom_var_v": bottom(1, {1: ["FARE"], 2: ["FARE"]}),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "RCO", 2: "RCO"}), "which_flux_heat_latent": methods(1, {1: "water", 2: "ice"}),
            "which_flux_heat_sensible": methods(1, {1: "RCO", 2: "RCO"}), "which_flux_momentum": methods(1, {1: "RCO", 2: "RCO"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo", 2: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR"]), "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # e2: MOM5 transfer coefficients per surface type (CMOI != CHEA), open water + ice, bias across the turn of the year
    mom_t = {1: ["TSUR", "FICE", "ALBE", "FARE", "CMOI", "CHEA"], 2: ["TSUR", "FICE", "ALBE", "FARE", "CMOI", "CHEA"]}
    mom_uv = {1: ["TSUR", "FICE", "FARE", "CMOM"], 2: ["TSUR", "FICE", "FARE", "CMOM"]}
    S.append(dict(
        name="mom5_s2_ice_bias", sizes=(11, 12, 13), steps=3, timestep=43200, init_date=19611231, bias=True,
        nml={
            "name_atmos_var_t": names(atm_t), "name_atmos_var_u": names(["PSUR", "UATM", "VATM", "TATM"]),
            "name_atmos_var_v": names(["PSUR", "UATM", "VATM", "TATM"]),
            "name_bottom_var_t": bottom(1, mom_t), "name_bottom_var_u": bottom(1, mom_uv), "name_bottom_var_v": bottom(1, mom_uv),
            "which_spec_vapor_surface_t": methods(1, {1: "CCLM", 2: "CCLM"}), "which_spec_vapor_surface_u": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_spec_vapor_surface_v": methods(1, {1: "CCLM", 2: "CCLM"}),
            "which_flux_mass_evap": methods(1, {1: "MOM5", 2: "MOM5"}), "which_flux_heat_latent": methods(1, {1: "water", 2: "ice"}),
            "which_flux_heat_sensible": methods(1, {1: "MOM5", 2: "MOM5"}), "which_flux_momentum": methods(1, {1: "MOM5", 2: "MOM5"}),
            "which_flux_radiation_blackbody": methods(1, {1: "StBo", 2: "StBo"}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR"]), "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    # e3: five surface types (open water + four ice classes), CCLM set, averages of the fluxes and of a pass-through variable
    five = {i: ["TSUR", "FICE", "ALBE", "FARE"] for i in range(1, 6)}
    fiveuv = {i: ["TSUR", "FICE", "FARE"] for i in range(1, 6)}
    allc = {i: "CCLM" for i in range(1, 6)}
    S.append(dict(
        name="cclm_s5_avg_bias", sizes=(10, 10, 10), steps=2, timestep=600, init_date=19610831, bias=True,
        nml={
            "name_atmos_var_t": names(atm_t + ["AMOI"]), "name_atmos_var_u": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_atmos_var_v": names(["PSUR", "UATM", "VATM", "AMOM", "TATM"]),
            "name_bottom_var_t": bottom(1, five), "name_bottom_var_u": bottom(1, fiveuv), "name_bottom_var_v": bottom(1, fiveuv),
            "which_spec_vapor_surface_t": methods(1, allc), "which_spec_vapor_surface_u": methods(1, allc),
            "which_spec_vapor_surface_v": methods(1, allc),
            "which_flux_mass_evap": methods(1, allc), "which_flux_heat_latent": methods(1, {1: "water", 2: "ice", 3: "ice", 4: "ice", 5: "ice"}),
            "which_flux_heat_sensible": methods(1, allc), "which_flux_momentum": methods(1, allc),
            "which_flux_radiation_blackbody": methods(1, {i: "StBo" for i in range(1, 6)}),
            "name_send_t": names(["MEVA", "HLAT", "HSEN", "RBBR", "RSDR", "TSUR"]),
            "name_send_u": names(["UMOM"]), "name_send_v": names(["VMOM"]),
        }))
    return S


def run_scenario(ref, sd, seed):
    R = Reference(ref)
    nml = dict(sd["nml"])
    if sd.get("uniform"):      # one surface type: every flux is "the same for each surface_type" (no FARE needed)
        for g in "tuv":
            if "name_send_" + g in nml:
                nml["send_uniform_" + g] = {(1, j): True for j in nml["name_send_" + g]}
                # the reference sizes output_field with ONE entry per uniform flux (flux_calculator.F90:655-660) but adds two
                # when it goes to the atmosphere and to the bottom model (:694-712): out-of-bounds there, so one receiver each
                if g == "t":
                    nml["send_to_atmos_t"] = {j: False for j in nml["name_send_t"]}
                else:
                    nml["send_to_bottom_" + g] = {(1, j): False for j in nml["name_send_" + g]}
    nml.setdefault("letter_bottom_model", {1: "M"})
    R.apply_namelist({k: v for k, v in nml.items() if k != "letter_bottom_model"})
    out = {"name": sd["name"], "grid_size": list(sd["sizes"]), "num_timesteps": sd["steps"], "timestep": sd["timestep"],
           "init_date": sd["init_date"], "bottom_model": 1, "bottom_letter": "M"}
    corr_txt = "&correctionsctl\n  init_date = %d\n  lcorrections(1) = %s\n/\n" % (sd["init_date"], ".TRUE." if sd["bias"] else ".FALSE.")
    out["namelist"] = nml_text({k: v for k, v in nml.items()}, corr_txt)
    try:
        R.setup(sd["sizes"])
    except FortranStop as e:
        out["stops_in_setup"] = True
        out["log_tail"] = R.I.log[-2:]
        return out
    S = R.scope["num_surface_types"]
    out["num_surface_types"] = S
    ins, outs = R.io_lists()
    out["input_fields"], out["output_fields"] = ins, outs
    feeder = Feeder(seed, sd["sizes"], S)
    R.feed = feeder
    if sd["bias"]:
        rng = random.Random("%d/corr" % seed)
        corr = [[(0.0 if rng.random() < 0.05 else rng.gauss(0, 1e-6)) for _ in range(sd["sizes"][0])] for _ in range(12)]
        R.set_corrections(sd["init_date"], corr)
        out["corrections_month_major"] = [[hx(x) for x in row] for row in corr]
    else:
        R.I.glob["init_date"] = sd["init_date"]
    if sd.get("regrid"):
        rng = random.Random("%d/regrid" % seed)
        mats = {}
        for which, ns, nd in (("t_to_u", sd["sizes"][0], sd["sizes"][1]), ("t_to_v", sd["sizes"][0], sd["sizes"][2])):
            src, dst, w = [], [], []
            for d in range(1, nd + 1):      # 1-3 sources per destination, deliberately unsorted by destination
                for _ in range(rng.randint(1, 3)):
                    src.append(rng.randint(1, ns)); dst.append(d); w.append(rng.uniform(0.1, 0.9))
            order = list(range(len(w)))
            rng.shuffle(order)
            src, dst, w = [src[k] for k in order], [dst[k] for k in order], [w[k] for k in order]
            R.set_matrix(which, src, dst, w)
            mats[which] = {"src_index": src, "dst_index": dst, "weight": [hx(x) for x in w]}
        out["regrid_matrices"] = mats
    out["registry_after_setup"] = R.registry()      # values: 'nan' = never written (uninitialised memory in the reference)
    out["methods"] = {k: {str(i): m for (mod, i), m in v.items() if mod == 1} for k, v in nml.items() if k.startswith("which_")}
    try:
        R.time_loop(sd["steps"], sd["timestep"])
    except TypeError as e:      # a disassociated pointer is dereferenced: undefined behaviour in the reference
        out["undefined_in_reference"] = "%s (call stack %s)" % (str(e)[:40] + " ...", [x[0] for x in R.I.stack])
        out["sent_before"] = R.puts
        return out, R
    out["received"] = [{"name": k[0], "grid": k[1], "time": k[2], "values": [hx(x) for x in v]} for k, v in sorted(feeder.cache.items())]
    out["sent"] = R.puts
    out["registry_final"] = R.registry()
    out["months"] = R.months
    out["warnings"] = [l for l in R.I.log if "WARNING" in l]
    return out, R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--set", default="main", choices=["main", "extra"],
                    help="main: step_golden.json (oracle on CPU + CUDA library with -m gpu); extra: step_golden_extra.json (oracle only)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.out is None:
        a.out = os.path.join(HERE, "step_golden.json" if a.set == "main" else "step_golden_extra.json")
    scen = []
    files = None
    defs, seed0 = (scenario_defs(), 0x5EEDF1C5) if a.set == "main" else (scenario_defs_extra(), 0x5EEDF1C5 + 100)
    for k, sd in enumerate(defs):
        res = run_scenario(a.ref, sd, seed0 + k)
        if isinstance(res, tuple):
            res, R = res
            files = R.files
        scen.append(res)
        print("  %-24s S=%s inputs=%s outputs=%s sent=%s %s" % (
            res["name"], res.get("num_surface_types"), len(res.get("input_fields", [])), len(res.get("output_fields", [])),
            len(res.get("sent", [])), "STOPS in set-up" if res.get("stops_in_setup") else ""))
    sha = {}
    for f in files or []:
        sha[f] = hashlib.sha256(open(os.path.join(a.ref, "src", f), "rb").read()).hexdigest()
    out = {"generator": "tests/golden/make_golden_step.py: the reference's host source executed by tests/golden/fortran_interp.py",
           "interpreted": {"flux_calculator.F90": [list(MAIN_SETUP), list(MAIN_LOOP)], "modules": files, "sha256": sha},
           "scenarios": scen}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote", a.out, os.path.getsize(a.out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
