"""Stated FP64 tolerance of the CUDA path against the oracle (north_star: relative <= 1e-12).

Every +,-,*,/,sqrt in the kernels is the same IEEE-754 operation the oracle performs (no FMA contraction),
so fields that involve no transcendental must be BIT-EXACT.  CUDA's exp() (<= 1 ulp) and pow() (<= 2 ulp)
may differ from glibc's by a last-place unit; fields downstream of them are compared with
    |x - ref| <= RTOL*|ref| + RTOL*SCALE[field]
where SCALE is the magnitude of the terms whose difference forms the flux (MEVA ~ flux_air*q,
HSEN ~ flux_air*c_p*T: cancellation turns an ulp of a term into many ulps of a small difference).
"""
import numpy as np

RTOL = 1e-12
SCALE = {"QSUR": 0.0, "MEVA": 1e-3, "HLAT": 3e3, "HSEN": 2e4, "UMOM": 0.0, "VMOM": 0.0, "RBBR": 0.0, "RSDR": 0.0}
# exact == no transcendental anywhere upstream, per formula set
EXACT_ALWAYS = {"RBBR", "RSDR"}
EXACT_RCO = {"HSEN", "UMOM", "VMOM"}


def check_field(name, got, ref, formula_set="CCLM", exact=None):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, name
    if exact is None:
        exact = name in EXACT_ALWAYS or (formula_set == "RCO" and name in EXACT_RCO)
    if exact:
        bad = ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
        assert not bad.any(), "%s: %d of %d cells not bit-exact, first at %d: %r vs %r" % (
            name, bad.sum(), got.size, np.argmax(bad), got[np.argmax(bad)], ref[np.argmax(bad)])
        return 0.0
    err = np.abs(got - ref)
    tol = RTOL * np.abs(ref) + RTOL * SCALE.get(name, 0.0)
    bad = ~(err <= tol)
    assert not bad.any(), "%s: %d of %d cells outside tolerance, worst |err|=%g at %d (ref %r got %r)" % (
        name, bad.sum(), got.size, np.nanmax(err), np.nanargmax(err), ref[np.nanargmax(err)], got[np.nanargmax(err)])
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(ref != 0, err / np.abs(ref), 0.0)
    return float(np.nanmax(rel)) if rel.size else 0.0
