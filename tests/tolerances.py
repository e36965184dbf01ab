"""Stated FP64 tolerance of the CUDA path against the oracle.

Every +,-,*,/,sqrt in the kernels is the same IEEE-754 operation the oracle performs (no FMA contraction), so fields
with no transcendental upstream must be BIT-EXACT.  The kernels' exp() (<= 1 ulp) and x**c (<= 2 ulp) may differ from
glibc's (<= 1 ulp each) in the last place; a field downstream of them is compared PER CELL with

    |x - ref| <= K_ULP * 2^-52 * ( |ref| + C )

where C is the magnitude of the cancelling terms of THAT cell -- the quantities an ulp of exp / pow is an ulp OF:

    MEVA (CCLM/MOM5) = F (q_s - q_a)              C = |F| (|q_s| + |q_a|),   F = a max(vel, u_min) p_s / (R_d T~)
    MEVA (RCO)       = rho c vel (q_w - q_a)      C = rho c vel (|q_w| + |q_a|)
    HLAT             = L * MEVA                   C = L * C(MEVA)
    HSEN (CCLM/MOM5) = F c_p (T_s - T_a EF)       C = |F| c_p (|T_s| + |T_a EF|),   EF = (p_s/p_a)**(R_d/c_p)
    QSUR, UMOM, VMOM                              C = 0 (no cancellation: a few ulp of the result itself)
    type-0 averages  = sum_i X_i FARE_i           C = sum_i |FARE_i| (|X_i| + C_i)

K_ULP = 8: three last-place units from the transcendentals (ours vs glibc's), the rest for their propagation through
the following roundings.  north_star's "relative <= 1e-12" is 4500 ulp, i.e. this is ~560 times tighter wherever
nothing cancels; where HSEN or MEVA are themselves 1e-9 of their terms no relative bound on the RESULT can hold for
any two correctly working libms, which is why the bound is stated on the terms.

The magnitudes C are formed here in plain numpy from the inputs (a few percent accuracy is all a bound needs); they
are not a second implementation of the formulae -- the reference values always come from the oracle / golden files.
"""
import numpy as np

ULP = 2.0 ** -52
K_ULP = 8.0
RTOL = K_ULP * ULP      # what "relative" means for the fields without cancellation

R_D, R_V, C_P, U_MIN = 287.05, 461.51, 1005.0, 0.01
L_V, L_S = 2.501e6, 2.835e6
EXACT_ALWAYS = {"RBBR", "RSDR"}
EXACT_RCO = {"HSEN", "UMOM", "VMOM"}


def _vel(f):
    return np.sqrt(f("UATM") ** 2 + f("VATM") ** 2)


def _flux_air(a, vel, ps, T, q):
    return np.abs(a * np.maximum(vel, U_MIN) * ps / (R_D * T * (1.0 + (R_V / R_D - 1.0) * q)))


class Scales:
    """per-cell magnitudes C of the cancelling terms for every output of one scenario.

    inputs: {(type, grid, var): array}; outputs_ref: {(type, grid, var): reference array}; methods: {(which, type): str}
    (tests/synthetic.Scenario.methods / the golden files' method tables)."""

    def __init__(self, inputs, outputs_ref, methods, num_surface_types):
        self.i, self.o, self.m, self.S = inputs, outputs_ref, methods, num_surface_types
        self.cache = {}

    def method(self, which, i):
        return self.m.get((which, i), "none")

    def of(self, key):
        if key not in self.cache:
            self.cache[key] = self._of(key)
        return self.cache[key]

    def _of(self, key):
        i, g, name = key
        if i == 0:      # area-fraction average of the surface types (average_across_surface_types)
            tot = 0.0
            for t in range(1, self.S + 1):
                if (t, g, name) in self.o and (t, g, "FARE") in self.i:
                    tot = tot + np.abs(self.i[(t, g, "FARE")]) * (np.abs(self.o[(t, g, name)]) + self.of((t, g, name)))
            return tot
        f = lambda v: self.i[(i, g, v)]      # noqa: E731
        which = {"MEVA": "which_flux_mass_evap", "HLAT": "which_flux_heat_latent", "HSEN": "which_flux_heat_sensible"}.get(name)
        if which and self.method(which, i) == "copy" and i != 1:      # a pointer alias of surface type 1's array (prepare.F90:36-38)
            return self.of((1, g, name))
        if name == "MEVA":
            m = self.method("which_flux_mass_evap", i)
            qa = f("QATM")
            if m in ("CCLM", "MOM5"):
                qs = self.o[(i, 1, "QSUR")] if (i, 1, "QSUR") in self.o else f("QSUR")
                F = _flux_air(f("AMOI" if m == "CCLM" else "CMOI"), _vel(f), f("PSUR"), f("TATM"), qs)
                return F * (np.abs(qs) + np.abs(qa))
            if m == "RCO":
                T = f("TSUR")
                qw = 0.62197 * 610.78 * np.exp(17.269 * (T - 273.15) / (T - 35.86)) / 1.013e5
                return 1.225 * 1.15e-3 * _vel(f) * (np.abs(qw) + np.abs(qa))
            return 0.0
        if name == "HLAT":
            m = self.method("which_flux_heat_latent", i)
            L = L_V if m == "water" else (L_S if m == "ice" else 0.0)
            return L * self.of((i, g, "MEVA"))
        if name == "HSEN":
            m = self.method("which_flux_heat_sensible", i)
            if m in ("CCLM", "MOM5"):
                F = _flux_air(f("AMOI" if m == "CCLM" else "CHEA"), _vel(f), f("PSUR"), f("TSUR"), f("QATM"))
                EF = (f("PSUR") / f("PATM")) ** (R_D / C_P)
                return F * C_P * (np.abs(f("TSUR")) + np.abs(f("TATM") * EF))
            return 0.0
        return 0.0


def is_exact(name, formula_set):
    return name in EXACT_ALWAYS or (formula_set == "RCO" and name in EXACT_RCO)


def check_field(name, got, ref, formula_set="CCLM", exact=None, scale=None):
    """assert got == ref bit for bit (exact fields) or within K_ULP last-place units of |ref| + scale per cell;
    returns the worst error as a fraction of the tolerance (0 for exact fields)"""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, name
    if exact is None:
        exact = is_exact(name, formula_set)
    if exact:
        bad = ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
        assert not bad.any(), "%s: %d of %d cells not bit-exact, first at %d: %r vs %r" % (
            name, bad.sum(), got.size, np.argmax(bad), got[np.argmax(bad)], ref[np.argmax(bad)])
        return 0.0
    if scale is None:
        scale = 0.0
    err = np.abs(got - ref)
    tol = K_ULP * ULP * (np.abs(ref) + scale)
    both_nan = np.isnan(got) & np.isnan(ref)
    bad = ~((err <= tol) | both_nan)
    if bad.any():
        w = int(np.nanargmax(np.where(bad, err / np.maximum(tol, 1e-300), 0.0)))
        raise AssertionError("%s: %d of %d cells outside %g ulp of the cancelling terms, worst at %d: ref %r got %r, |err| %g = %.2f x tolerance"
                             % (name, bad.sum(), got.size, K_ULP, w, ref[w], got[w], err[w], err[w] / max(tol[w] if np.ndim(tol) else tol, 1e-300)))
    with np.errstate(divide="ignore", invalid="ignore"):
        frac = np.where(tol > 0, err / tol, 0.0)
    return float(np.nanmax(frac)) if frac.size else 0.0


def check_outputs(got, ref, inputs, methods, num_surface_types, formula_set="CCLM"):
    """all outputs of a scenario: {key: array} against {key: array}; returns {key: worst fraction of the tolerance}"""
    sc = Scales(inputs, ref, methods, num_surface_types)
    worst = {}
    for k in sorted(ref):
        worst[k] = check_field(k[2], got[k], ref[k], formula_set, scale=sc.of(k))
    return worst


def check_scenario(sc, got, ref, keys=None, cells=None):
    """all outputs (or `keys`) of a synthetic.Scenario; cells: optional index array -- got/ref hold only those cells"""
    inputs = sc.inputs if cells is None else {k: a[cells] for k, a in sc.inputs.items()}
    scales = Scales(inputs, ref, sc.methods, sc.S)
    worst = {}
    for k in sorted(keys if keys is not None else ref):
        worst[k] = check_field(k[2], got[k], ref[k], sc.formula_set, scale=scales.of(k))
    return worst


def routine_scale(name, ins):
    """cancelling-term magnitude of ONE flux_lib routine called on its own; ins = input arrays in the routine's dummy order"""
    if name in ("flux_heat_sensible_cclm", "flux_heat_sensible_mom5"):
        a, pa, ps, q, Ta, Ts, u, v = ins[:8]
        F = _flux_air(a, np.sqrt(u * u + v * v), ps, Ts, q)
        return F * C_P * (np.abs(Ts) + np.abs(Ta * (ps / pa) ** (R_D / C_P)))
    if name == "flux_mass_evap_rco":
        qa, T, u, v = ins[:4]
        qw = 0.62197 * 610.78 * np.exp(17.269 * (T - 273.15) / (T - 35.86)) / 1.013e5
        return 1.225 * 1.15e-3 * np.sqrt(u * u + v * v) * (np.abs(qw) + np.abs(qa))
    return 0.0
