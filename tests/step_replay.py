"""Replays the scenarios of tests/golden/step_golden.json -- the reference's own host source executed by the interpreter
of tests/golden/ (set-up of the registry from a namelist, time loop, send loops) -- through an object with the
FluxCalculator interface (the CPU oracle or the CUDA library) and compares everything the reference sends.

The registry is rebuilt from the golden file exactly as the reference left it after its set-up: which slots share
storage (pointer aliases), the %allocated flags, constants written by init_localvar / default-valued outputs.  Per time
step the received fields are written into the bound arrays, regridded where the namelist says so, the early and the
normal phase run, and every oasis_put of the reference is compared with the bound array of that output."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_golden.json")
fh = float.fromhex
DIRECTION = {(2, 1): 0, (3, 1): 1, (1, 2): 2, (1, 3): 3}      # (from grid, to grid) -> fc_set_regrid_matrix direction


GOLDEN_EXTRA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_golden_extra.json")      # oracle only


def scenarios(path=GOLDEN):
    with open(path) as f:
        return json.load(f)["scenarios"]


def arr(hexes):
    return np.array([fh(x) for x in hexes], dtype=np.float64)


class Replay:
    def __init__(self, scen, target, explicit_allocated=True):
        self.s, self.t = scen, target
        self.store = {}      # storage id -> array (aliases share it)
        self.slot = {}       # (type, grid, var) -> array
        self.reg = {}
        for r in scen["registry_after_setup"]:
            if r["storage"] not in self.store:
                self.store[r["storage"]] = arr(r["values"])
            a = self.store[r["storage"]]
            key = (r["type"], r["grid"], r["var"])
            self.slot[key], self.reg[key] = a, r
            target.bind_field(r["type"], r["grid"], r["var"], a)
        if explicit_allocated:
            for key, r in self.reg.items():
                target.set_allocated(key[0], key[1], key[2], r["allocated"])
        for which, per_type in scen["methods"].items():
            for i, m in per_type.items():
                target.set_method(which, int(i), m)
        if "corrections_month_major" in scen:
            corr = np.array([[fh(x) for x in row] for row in scen["corrections_month_major"]])      # [12][n]
            target.set_corrections(np.ascontiguousarray(corr.T), True, scen["init_date"])          # Fortran (1,12,n) == C (n,12)
        for o in scen["output_fields"]:
            target.add_output_field(o["type"], o["grid"], o["var"])
        # the reference runs distribute_shortwave_radiation_flux unconditionally and is undefined without RSDR (App. F-7);
        # the scenarios replayed here all have RSDD(0) and RSDR(i)
        target.set_distribute_shortwave(all((i, 1, "RSDR") in self.slot for i in range(1, scen["num_surface_types"] + 1))
                                        and (0, 1, "RSDD") in self.slot)
        for which, m in scen.get("regrid_matrices", {}).items():
            d = {"u_to_t": 0, "v_to_t": 1, "t_to_u": 2, "t_to_v": 3}[which]
            target.set_regrid_matrix(d, m["src_index"], m["dst_index"], arr(m["weight"]))
        self.received = {(r["name"], r["grid"], r["time"]): arr(r["values"]) for r in scen.get("received", [])}

    def receive(self, time, early):
        """oasis_get of one phase + do_regridding of the received fields (flux_calculator.F90:873-897 / :943-967)"""
        ins = [f for f in self.s["input_fields"] if f["early"] == early]
        for g in (1, 2, 3):
            for f in ins:
                if f["grid"] == g:
                    self.slot[(f["type"], g, f["var"])][:] = self.received[(f["name"], g, time)]
        for f in ins:      # do_regridding(idx, surface_type): basic.F90:463-522, order u->t, v->t, t->u, t->v per surface type
            for j in range(1, 11):
                if not (j == f["type"] or f["type"] == 0):
                    continue
                for (frm, to) in ((2, 1), (3, 1), (1, 2), (1, 3)):
                    r = self.reg.get((j, frm, f["var"]))
                    if r and to in r.get("regrid_to", []):
                        self.t.regrid(DIRECTION[(frm, to)], self.slot[(j, to, f["var"])], self.slot[(j, frm, f["var"])])

    def run(self, compare):
        """compare(put, got_array, key): called for every oasis_put of the reference, in the reference's order"""
        puts = list(self.s["sent"])
        k = 0
        for n in range(self.s["num_timesteps"]):
            time = n * self.s["timestep"]
            for early in (True, False):
                self.receive(time, early)
                (self.t.step_early if early else self.t.step_normal)(time)
                if hasattr(self.t, "synchronize"):
                    self.t.synchronize()
                for g in (1, 2, 3):
                    for o in self.s["output_fields"]:
                        if o["grid"] == g and o["early"] == early:
                            put = puts[k]
                            k += 1
                            assert put["name"] == o["name"] and put["time"] == time and put["grid"] == g, (put["name"], o["name"])
                            compare(put, self.slot[(o["type"], g, o["var"])], (o["type"], g, o["var"]))
        assert k == len(puts)
