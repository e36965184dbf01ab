// formulas.cuh -- per-cell flux formulae as __device__ inlines (binary64).
//
// Each function follows one routine of the reference's flux_lib in the Fortran evaluation order
// (left to right within equal precedence, parentheses as written).  Products and sums use the
// *_rn intrinsics, which the compiler never contracts into FMA, so every +,-,*,/,sqrt is the same
// IEEE-754 operation the reference's `-fp-model precise` build performs; only exp() and pow()
// (CUDA libdevice: <=1 ulp / <=2 ulp) can differ from the host libm by a last-place unit.
// Reference paths relative to /root/reference/src/flux_lib.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace fc {

// default_values, constants/flux_constants.F90:13-32 (binary64 literals under -r8)
struct Consts {
    double heat_capacity_air;         // c_p
    double latent_heat_vaporization;  // L_v
    double latent_heat_sublimation;   // L_s
    double gas_constant_air;          // R_d
    double gas_constant_vapor;        // R_v
    double stefan_boltzmann_constant; // sigma
    double u_min_evap;
    // derived (computed on the host with IEEE double ops; identical to evaluating them per cell)
    double rv_over_rd_m1;  // R_v/R_d - 1.0
    double rd_over_rv;     // R_d/R_v
    double one_m_rd_rv;    // 1.0 - R_d/R_v
    double rd_over_cp;     // R_d/c_p
};

__host__ __device__ inline Consts make_consts(double c_p = 1005.0, double L_v = 2.501e6, double L_s = 2.835e6,
                                              double R_d = 287.05, double R_v = 461.51, double sigma = 5.67e-8,
                                              double u_min = 0.01)
{
    Consts c;
    c.heat_capacity_air = c_p;
    c.latent_heat_vaporization = L_v;
    c.latent_heat_sublimation = L_s;
    c.gas_constant_air = R_d;
    c.gas_constant_vapor = R_v;
    c.stefan_boltzmann_constant = sigma;
    c.u_min_evap = u_min;
    c.rv_over_rd_m1 = R_v / R_d - 1.0;
    c.rd_over_rv = R_d / R_v;
    c.one_m_rd_rv = 1.0 - R_d / R_v;
    c.rd_over_cp = R_d / c_p;
    return c;
}

#define FC_DI __device__ __forceinline__

FC_DI double mul(double a, double b) { return __dmul_rn(a, b); }
FC_DI double add(double a, double b) { return __dadd_rn(a, b); }
FC_DI double sub(double a, double b) { return __dsub_rn(a, b); }
FC_DI double dvd(double a, double b) { return __ddiv_rn(a, b); }

// vel = sqrt(u*u + v*v)   (e.g. mass/flux_mass_evap.F90:76)
FC_DI double wind_speed(double u, double v) { return __dsqrt_rn(add(mul(u, u), mul(v, v))); }

// auxiliaries/flux_aux_vapor.F90:60-68
FC_DI double spec_vapor_surface_cclm(double f_ice, double p_s, double T_s, const Consts &c)
{
    const double alpha_water = 17.2693882, alpha_ice = 21.8745584;      // :39-40
    const double T_1 = 273.16, T_2_water = 35.86, T_2_ice = 7.66;       // :41-44
    const double p_0 = 610.78;                                          // :45
    const double alpha = add(alpha_water, mul(alpha_ice - alpha_water, f_ice));   // :60
    const double T_2 = add(T_2_water, mul(T_2_ice - T_2_water, f_ice));           // :61
    const double e_sat = mul(p_0, exp(dvd(mul(alpha, sub(T_s, T_1)), sub(T_s, T_2))));  // :63-64
    return dvd(mul(c.rd_over_rv, e_sat), sub(p_s, mul(c.one_m_rd_rv, e_sat)));          // :66-68
}

// T_tilde = T * (1.0 + (R_v/R_d - 1.0) * q)    (mass/flux_mass_evap.F90:72-74 and siblings)
FC_DI double t_tilde(double T, double q, const Consts &c) { return mul(T, add(1.0, mul(c.rv_over_rd_m1, q))); }

// mass/flux_mass_evap.F90:72-83; `vel` passed in so the caller can share it with the sensible heat
FC_DI double flux_mass_evap_cclm(double a_moisture, double p_s, double q_a, double q_s, double T_s, double vel,
                                 const Consts &c)
{
    const double T_tilde = t_tilde(T_s, q_s, c);
    const double flux_air = dvd(mul(mul(a_moisture, fmax(vel, c.u_min_evap)), p_s),
                                mul(c.gas_constant_air, T_tilde));     // :78-80
    return mul(flux_air, sub(q_s, q_a));                                // :82-83
}

// mass/flux_mass_evap.F90:148-156 (Meier et al. 1999)
FC_DI double flux_mass_evap_rco(double q_a, double T_s, double vel)
{
    const double rho_a = 1.225, c_aw = 1.15E-03, epsilon = 0.62197, P_0 = 1.013E+05;   // :135-138
    const double r = 6.1078E+02, c_1 = 17.269, c_2 = 35.86;                            // :143-145
    const double e_w = mul(r, exp(dvd(mul(c_1, sub(T_s, 273.15)), sub(T_s, c_2))));    // :148
    const double q_w = dvd(mul(epsilon, e_w), P_0);                                    // :151
    return mul(mul(rho_a * c_aw, vel), sub(q_w, q_a));                                 // :156
}

// heat/flux_heat_latent.F90:41 (ice: L_s) and :65 (water: L_v)
FC_DI double flux_heat_latent(double evap, double latent_heat) { return mul(evap, latent_heat); }

// heat/flux_heat_sensible.F90:84-98
FC_DI double flux_heat_sensible_cclm(double a_moisture, double p_a, double p_s, double q_s, double T_a, double T_s,
                                     double vel, const Consts &c)
{
    const double T_tilde = t_tilde(T_s, q_s, c);                        // :84-86
    const double flux_air = dvd(mul(mul(a_moisture, fmax(vel, c.u_min_evap)), p_s),
                                mul(c.gas_constant_air, T_tilde));     // :90-92
    const double EF = pow(dvd(p_s, p_a), c.rd_over_cp);                 // :94-95
    return mul(mul(flux_air, c.heat_capacity_air), sub(T_s, mul(T_a, EF)));   // :97-98
}

// heat/flux_heat_sensible.F90:157-165
FC_DI double flux_heat_sensible_rco(double T_a, double T_s, double vel)
{
    const double rho_a = 1.225, c_pa = 1.008E+03;                       // :151-152
    const double c_aw = (T_a < T_s) ? 1.13E-03 : 0.66E-03;              // :157-161
    return mul(mul(mul(rho_a * c_pa, c_aw), vel), sub(T_s, T_a));       // :165
}

// momentum/flux_momentum.F90:63-73; returns flux_air, the caller forms -flux_air*u / -flux_air*v
FC_DI double momentum_flux_air_cclm(double a_momentum, double p_s, double q_s, double T_s, double vel,
                                    const Consts &c)
{
    const double T_tilde = t_tilde(T_s, q_s, c);                        // :63-65
    return dvd(mul(mul(a_momentum, vel), p_s), mul(c.gas_constant_air, T_tilde));   // :69-70
}

// momentum/flux_momentum.F90:126-136; returns rho_a*c_aw*vel
FC_DI double momentum_flux_air_rco(double vel)
{
    const double rho_a = 1.225;                                         // :122
    const double c_aw = (vel < 11.0) ? 1.2E-03 : add(0.49E-03, mul(0.065E-03, vel));   // :129-133
    return mul(mul(rho_a, c_aw), vel);
}

FC_DI double momentum_component(double flux_air, double wind) { return -mul(flux_air, wind); }   // :72-73 / :135-136

// radiation/flux_radiation_blackbody.F90:40: sigma * T**4, integer power == (T*T)*(T*T)
FC_DI double flux_radiation_blackbody_StBo(double T_s, double sigma)
{
    const double T2 = mul(T_s, T_s);
    return mul(sigma, mul(T2, T2));
}

}  // namespace fc
