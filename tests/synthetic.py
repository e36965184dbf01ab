"""Synthetic exchange grids for tests and benchmarks (SURVEY 8(d) distributions).

Counter-based RNG keyed by (seed, field name, surface type, grid, GLOBAL cell index): a shard
[offset, offset+n) of a larger grid sees exactly the values the unsharded grid has at those cells, so
results can be compared across GPU counts.  Host-side data generation only -- no flux arithmetic.

Test / benchmark infrastructure: lives beside the tests, imports nothing of the product (the CPU arm of bench.py uses
it without mapping libfluxcalc_b200.so).
"""
import zlib

import numpy as np


SEED = 0x5EEDF1C5
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform01(name, n, surface_type=0, grid=1, offset=0, seed=SEED, stream=0):
    """n doubles in [0,1) for global cells offset .. offset+n-1"""
    key = np.uint64(zlib.crc32(("%s/%d/%d/%d" % (name, surface_type, grid, stream)).encode())) ^ (np.uint64(seed) << np.uint64(32))
    with np.errstate(over="ignore"):
        ctr = np.arange(offset, offset + n, dtype=np.uint64)
        h = _splitmix64(_splitmix64(ctr ^ _splitmix64(np.full(1, key, dtype=np.uint64))[0]))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(name, n, **kw):
    u1 = uniform01(name, n, stream=1, **kw)
    u2 = uniform01(name, n, stream=2, **kw)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def make_field(name, n, surface_type=0, grid=1, offset=0, seed=SEED, ice=False, base=None):
    kw = dict(surface_type=surface_type, grid=grid, offset=offset, seed=seed)
    U = lambda lo, hi, nm=name: lo + (hi - lo) * uniform01(nm, n, **kw)   # noqa: E731
    if name == "TSUR":
        return U(243.15, 273.15) if ice else U(271.35, 303.15)
    if name == "TATM":      # base = TSUR of surface type 1; 1 % exactly equal to it
        t = base + U(-5.0, 5.0)
        eq = uniform01("TATM_eq", n, **kw) < 0.01
        return np.where(eq, base, t)
    if name == "PSUR":
        return U(9.8e4, 1.04e5)
    if name == "PATM":      # base = PSUR
        return base - U(100.0, 1500.0)
    if name == "QATM":
        return U(1e-3, 1.5e-2)
    if name in ("UATM", "VATM"):
        w = np.clip(6.0 * normal(name, n, **kw), -35.0, 35.0)
        sel = uniform01("WIND_special", n, **kw)      # same selector for U and V
        calm = sel < 0.01
        thr = (sel >= 0.01) & (sel < 0.02)
        exact = (sel >= 0.02) & (sel < 0.021)
        if name == "UATM":
            w = np.where(calm, 0.0, w)
            w = np.where(thr, 11.0 + 2e-9 * (uniform01("WIND_eps", n, **kw) - 0.5), w)
            w = np.where(exact, 11.0, w)
        else:
            w = np.where(calm | thr | exact, 0.0, w)
        return w
    if name in ("AMOI", "AMOM", "CMOI", "CHEA", "CMOM"):
        return U(8e-4, 2.5e-3)
    if name == "FICE":
        return np.full(n, 1.0 if ice else 0.0)
    if name == "FARE":
        return U(0.0, 1.0)
    if name == "RSDD":
        return -U(0.0, 900.0)
    if name in ("ALBA", "ALBE"):
        return U(0.05, 0.8)
    if name == "AREA":
        return U(1e7, 4e8)
    if name == "CORR":      # (n, 12) == Fortran corrections(1,12,n)
        out = np.empty((n, 12))
        for m in range(12):
            c = 1e-6 * normal("CORR%02d" % m, n, **kw)
            out[:, m] = np.where(uniform01("CORR0_%02d" % m, n, **kw) < 0.05, 0.0, c)
        return out
    raise KeyError(name)


ATMOS_T = ["PSUR", "PATM", "QATM", "TATM", "UATM", "VATM", "RSDD", "ALBA"]
ATMOS_UV = ["PSUR", "UATM", "VATM"]
COEF = {"CCLM": ("AMOI", "AMOI", "AMOM"), "MOM5": ("CMOI", "CHEA", "CMOM"), "RCO": (None, None, None)}
T_OUTPUTS = ["QSUR", "MEVA", "HLAT", "HSEN", "RBBR", "RSDR"]


class Scenario:
    """One flux_calculator instance worth of fields + namelist choices.

    formula_set: 'CCLM' | 'MOM5' | 'RCO'; S surface types (type 1 open water, type >= 2 ice);
    averaging: bind FARE and type-0 arrays for the sent fluxes (needs S >= 2);
    n = (N_t, N_u, N_v) local cells, offset = global index of the first local cell on each grid.
    """

    def __init__(self, formula_set="CCLM", n=(1000, 1000, 1000), S=1, bias=False, averaging=False,
                 offset=(0, 0, 0), seed=SEED, shortwave=True, init_date=19610101, fractional_ice=False,
                 passthrough_avg=False):
        self.formula_set, self.n, self.S, self.bias = formula_set, tuple(int(x) for x in n), S, bias
        self.averaging = averaging and S >= 2
        self.offset, self.seed, self.shortwave, self.init_date = tuple(offset), seed, shortwave, init_date
        self.inputs = {}     # (type, grid, var) -> array; aliased arrays are the SAME object
        self.outputs = {}
        self.methods = {}
        self.send = []
        a_evap, a_sens, a_mom = COEF[formula_set]
        for g in (1, 2, 3):
            ng, off = self.n[g - 1], self.offset[g - 1]
            mk = lambda name, st=0, **kw: make_field(name, ng, surface_type=st, grid=g, offset=off, seed=seed, **kw)  # noqa: E731
            # bottom-model fields per surface type
            for i in range(1, S + 1):
                ice = i >= 2
                self.inputs[(i, g, "TSUR")] = mk("TSUR", i, ice=ice)
                fice = mk("FICE", i, ice=ice)
                if fractional_ice:
                    fice = make_field("FARE", ng, surface_type=i, grid=g, offset=off, seed=seed + 7)
                self.inputs[(i, g, "FICE")] = fice
                if self.averaging or passthrough_avg:
                    if S == 2 and i == 2:
                        self.inputs[(i, g, "FARE")] = 1.0 - self.inputs[(1, g, "FARE")]
                    else:
                        self.inputs[(i, g, "FARE")] = mk("FARE", i)
            # atmosphere fields: type 0 owns them, every surface type aliases them (basic.F90:334-358)
            names = list(ATMOS_T if g == 1 else ATMOS_UV)
            coefs = [a_evap, a_sens] if g == 1 else [a_mom]
            for cname in coefs:
                if cname and cname not in names:
                    names.append(cname)
            made = {}
            for name in names:
                if name == "TATM":
                    arr = mk("TATM", 0, base=self.inputs[(1, g, "TSUR")])
                elif name == "PATM":
                    arr = mk("PATM", 0, base=made["PSUR"])
                else:
                    arr = mk(name, 0)
                made[name] = arr
                if name in ("RSDD", "ALBA") and not shortwave:
                    continue
                self.inputs[(0, g, name)] = arr
                if name in ("CMOI", "CHEA", "CMOM"):      # ocean coefficients are bottom fields of type 1, shared
                    del self.inputs[(0, g, name)]
                for i in range(1, S + 1):
                    if name not in ("RSDD", "ALBA"):
                        self.inputs[(i, g, name)] = arr
            if g == 1 and shortwave:
                for i in range(1, S + 1):
                    self.inputs[(i, g, "ALBE")] = mk("ALBE", i)
            # outputs
            outs = T_OUTPUTS if g == 1 else (["QSUR", "UMOM"] if g == 2 else ["QSUR", "VMOM"])
            for i in range(1, S + 1):
                for name in outs:
                    if name == "RSDR" and not shortwave:
                        continue
                    self.outputs[(i, g, name)] = np.full(ng, np.nan)
            if self.averaging:
                for name in outs:
                    if name == "QSUR" or (name == "RSDR" and not shortwave):
                        continue
                    self.outputs[(0, g, name)] = np.full(ng, np.nan)
                    self.send.append((0, g, name))
            if passthrough_avg and g == 1:
                for name in ("TSUR", "FICE"):
                    self.outputs[(0, g, name)] = np.full(ng, np.nan)
                    self.send.append((0, g, name))
        for i in range(1, S + 1):
            self.methods[("which_spec_vapor_surface_t", i)] = "CCLM"
            self.methods[("which_spec_vapor_surface_u", i)] = "CCLM" if formula_set != "RCO" else "none"
            self.methods[("which_spec_vapor_surface_v", i)] = "CCLM" if formula_set != "RCO" else "none"
            self.methods[("which_flux_mass_evap", i)] = formula_set
            self.methods[("which_flux_heat_latent", i)] = "ice" if i >= 2 else "water"
            self.methods[("which_flux_heat_sensible", i)] = formula_set
            self.methods[("which_flux_momentum", i)] = formula_set
            self.methods[("which_flux_radiation_blackbody", i)] = "StBo"
        if formula_set == "RCO":
            for i in range(1, S + 1):
                self.outputs.pop((i, 2, "QSUR"), None)
                self.outputs.pop((i, 3, "QSUR"), None)
        self.corrections = make_field("CORR", self.n[0], grid=1, offset=self.offset[0], seed=seed) if bias else None
        self.area = {g: make_field("AREA", self.n[g - 1], grid=g, offset=self.offset[g - 1], seed=seed) for g in (1, 2, 3)}

    # ---------------------------------------------------------------------------------------
    def distinct_arrays(self):
        seen, res = set(), []
        for d in (self.inputs, self.outputs):
            for k, a in d.items():
                if id(a) not in seen:
                    seen.add(id(a))
                    res.append((k, a))
        return res

    def clone(self):
        """deep copy preserving aliasing; returns (inputs, outputs) dicts"""
        memo = {}
        def cp(a):
            if id(a) not in memo:
                memo[id(a)] = a.copy()
            return memo[id(a)]
        return {k: cp(a) for k, a in self.inputs.items()}, {k: cp(a) for k, a in self.outputs.items()}

    def apply(self, target, inputs=None, outputs=None, wrap=lambda a: a):
        """bind into a FluxCalculator-like object (bind_field/set_method/set_corrections/add_output_field)"""
        inputs = self.inputs if inputs is None else inputs
        outputs = self.outputs if outputs is None else outputs
        wrapped = {}
        def w(a):
            if id(a) not in wrapped:
                wrapped[id(a)] = wrap(a)
            return wrapped[id(a)]
        for (i, g, name), a in list(inputs.items()) + list(outputs.items()):
            target.bind_field(i, g, name, w(a))
        for (which, i), m in self.methods.items():
            target.set_method(which, i, m)
        if self.corrections is not None:
            target.set_corrections(self.corrections, True, self.init_date)
        for (i, g, name) in self.send:
            target.add_output_field(i, g, name)
        target.set_distribute_shortwave(self.shortwave)
        return wrapped
