#!/usr/bin/env python3
"""50-digit mpmath known-answer values for the flux_lib formulas (independent of the C oracle and of the
source interpreter): constants and inputs are rounded to binary64 first, the formula is then evaluated
exactly; tol_ulp is the rounding-error budget of the binary64 evaluation order (HSEN/MEVA: cancellation).
Writes tests/golden/kat_mpmath.json.   Usage: python tests/golden/make_kat.py"""
import json
import os

from mpmath import mp, mpf, exp, sqrt

mp.dps = 50
D = lambda x: mpf(float(x))          # binary64-rounded literal
c_p, L_v, L_s, R_d, R_v, sigma, u_min = D(1005.0), D(2.501e6), D(2.835e6), D(287.05), D(461.51), D(5.67e-8), D(0.01)


def qsur(f, p, T):
    a = D(17.2693882) + (D(21.8745584) - D(17.2693882)) * f
    T2 = D(35.86) + (D(7.66) - D(35.86)) * f
    e = D(610.78) * exp(a * (T - D(273.16)) / (T - T2))
    return (R_d / R_v) * e / (p - (1 - R_d / R_v) * e)


def flux_air(a, vel, p, T, q, clip=True):
    v = max(vel, u_min) if clip else vel
    return a * v * p / (R_d * T * (1 + (R_v / R_d - 1) * q))


cases = []
def add(routine, ins, outs, tol, scale=None):
    """tol is in ulps of max(|result|, scale): scale = magnitude of the terms whose difference forms the result"""
    cases.append({"routine": routine, "in": [float(x).hex() for x in ins], "out": [mp.nstr(o, 40) for o in outs],
                  "tol_ulp": tol, "scale": [float(abs(x)) for x in (scale or [0] * len(outs))]})

pts = [dict(T=283.15, p=101325.0, Ta=281.15, pa=100800.0, qa=0.005, u=5.0, v=-3.0, am=1.2e-3, amom=1.3e-3),
       dict(T=275.4, p=99321.7, Ta=279.9, pa=98011.2, qa=0.0031, u=-11.2, v=7.9, am=2.1e-3, amom=9.5e-4),
       dict(T=300.2, p=103400.0, Ta=299.95, pa=103100.0, qa=0.0142, u=0.001, v=0.002, am=8.8e-4, amom=2.2e-3)]
for P in pts:
    T, p, Ta, pa, qa, u, v, am, amom = [D(P[k]) for k in ("T", "p", "Ta", "pa", "qa", "u", "v", "am", "amom")]
    vel = sqrt(u * u + v * v)
    for f in (0.0, 1.0):
        add("spec_vapor_surface_cclm", [f, p, T], [qsur(D(f), p, T)], [6])
    q0 = mpf(float(qsur(D(0.0), p, T)))          # the binary64 QSUR that is fed on
    # call-site wiring: T slot <- TATM (calculate.F90:87); the routine itself just takes "temperature_surface"
    add("flux_mass_evap_cclm", [am, p, qa, q0, Ta, u, v], [flux_air(am, vel, p, Ta, q0) * (q0 - qa)], [8],
        [flux_air(am, vel, p, Ta, q0) * q0])
    add("flux_mass_evap_mom5", [am, p, qa, q0, Ta, u, v], [flux_air(am, vel, p, Ta, q0) * (q0 - qa)], [8],
        [flux_air(am, vel, p, Ta, q0) * q0])
    ew = D(6.1078E+02) * exp(D(17.269) * (T - D(273.15)) / (T - D(35.86)))
    add("flux_mass_evap_rco", [qa, T, u, v], [D(1.225) * D(1.15E-03) * vel * (D(0.62197) * ew / D(1.013E+05) - qa)], [8],
        [D(1.225) * D(1.15E-03) * vel * D(0.62197) * ew / D(1.013E+05)])
    m0 = mpf(float(flux_air(am, vel, p, Ta, q0) * (q0 - qa)))
    add("flux_heat_latent_water", [m0], [m0 * L_v], [1])
    add("flux_heat_latent_ice", [m0], [m0 * L_s], [1])
    EF = (p / pa) ** (R_d / c_p)
    add("flux_heat_sensible_cclm", [am, pa, p, qa, Ta, T, u, v], [flux_air(am, vel, p, T, qa) * c_p * (T - Ta * EF)], [8],
        [flux_air(am, vel, p, T, qa) * c_p * T])
    add("flux_heat_sensible_mom5", [am, pa, p, qa, Ta, T, u, v], [flux_air(am, vel, p, T, qa) * c_p * (T - Ta * EF)], [8],
        [flux_air(am, vel, p, T, qa) * c_p * T])
    caw = D(1.13E-03) if Ta < T else D(0.66E-03)
    add("flux_heat_sensible_rco", [Ta, T, u, v], [D(1.225) * D(1.008E+03) * caw * vel * (T - Ta)], [4])
    fa = flux_air(amom, vel, p, T, q0, clip=False)
    add("flux_momentum_cclm", [amom, p, q0, T, u, v], [-fa * u, -fa * v], [6, 6])
    add("flux_momentum_mom5", [amom, p, q0, T, u, v], [-fa * u, -fa * v], [6, 6])
    cd = D(1.2E-03) if vel < 11 else D(0.49E-03) + D(0.065E-03) * vel
    add("flux_momentum_rco", [u, v], [-D(1.225) * cd * vel * u, -D(1.225) * cd * vel * v], [4, 4])
    add("flux_radiation_blackbody_stbo", [T], [sigma * T ** 4], [3])
    add("distribute_radiation_flux", [D(-612.5), D(0.1), D(0.2)], [D(-612.5)], [0])

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_mpmath.json")
json.dump({"generator": "tests/golden/make_kat.py (mpmath %d digits)" % mp.dps, "cases": cases}, open(out, "w"), indent=0)
print("wrote", out, len(cases), "cases")
