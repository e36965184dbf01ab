// formulas.cuh -- the per-cell flux formulae, written ONCE as templates over an arithmetic policy
// (vmath.cuh): ExactScalar for one cell with CUDA's IEEE routines, FastVec<V> for V cells in lock step.
//
// Each function follows one routine of the reference's flux_lib in the Fortran evaluation order (left to
// right within equal precedence, parentheses as written).  The policies implement mul/add/sub with the
// *_rn intrinsics (never contracted into FMA) and div/sqrt correctly rounded, so every +,-,*,/,sqrt is the
// same IEEE-754 operation the reference's `-fp-model precise` build performs; only exp() and pow() can
// differ from the host libm in the last place.  Reference paths relative to /root/reference/src/flux_lib.
#pragma once

#include "vmath.cuh"

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

// plain scalar helpers (averaging, diagnostics)
FC_DI double mul(double a, double b) { return __dmul_rn(a, b); }
FC_DI double add(double a, double b) { return __dadd_rn(a, b); }

// vel = sqrt(u*u + v*v)   (e.g. mass/flux_mass_evap.F90:76)
template <class M>
FC_DI typename M::T wind_speed(M &m, const typename M::T &u, const typename M::T &v)
{
    return m.sqrt(M::add(M::mul(u, u), M::mul(v, v)));
}

// auxiliaries/flux_aux_vapor.F90:60-68
template <class M>
FC_DI typename M::T spec_vapor_surface_cclm(M &m, const typename M::T &f_ice, const typename M::T &p_s,
                                            const typename M::T &T_s, const Consts &c)
{
    const double alpha_water = 17.2693882, alpha_ice = 21.8745584;      // :39-40
    const double T_1 = 273.16, T_2_water = 35.86, T_2_ice = 7.66;       // :41-44
    const double p_0 = 610.78;                                          // :45
    const auto alpha = M::add(M::bc(alpha_water), M::mul(M::bc(alpha_ice - alpha_water), f_ice));   // :60
    const auto T_2 = M::add(M::bc(T_2_water), M::mul(M::bc(T_2_ice - T_2_water), f_ice));           // :61
    const auto arg = m.div(M::mul(alpha, M::sub(T_s, M::bc(T_1))), M::sub(T_s, T_2));
    const auto e_sat = M::mul(M::bc(p_0), m.exp(arg));                                              // :63-64
    return m.div(M::mul(M::bc(c.rd_over_rv), e_sat), M::sub(p_s, M::mul(M::bc(c.one_m_rd_rv), e_sat)));   // :66-68
}

// T_tilde = T * (1.0 + (R_v/R_d - 1.0) * q)    (mass/flux_mass_evap.F90:72-74 and siblings)
template <class M>
FC_DI typename M::T t_tilde(const typename M::T &T, const typename M::T &q, const Consts &c)
{
    return M::mul(T, M::add(M::bc(1.0), M::mul(M::bc(c.rv_over_rd_m1), q)));
}

// mass/flux_mass_evap.F90:72-83; `vel` passed in so the caller can share it with the sensible heat
template <class M>
FC_DI typename M::T flux_mass_evap_cclm(M &m, const typename M::T &a_moisture, const typename M::T &p_s,
                                        const typename M::T &q_a, const typename M::T &q_s, const typename M::T &T_s,
                                        const typename M::T &vel, const Consts &c)
{
    const auto T_tilde = t_tilde<M>(T_s, q_s, c);
    const auto flux_air = m.div(M::mul(M::mul(a_moisture, M::max(vel, M::bc(c.u_min_evap))), p_s),
                                M::mul(M::bc(c.gas_constant_air), T_tilde));     // :78-80
    return M::mul(flux_air, M::sub(q_s, q_a));                                    // :82-83
}

// mass/flux_mass_evap.F90:148-156 (Meier et al. 1999)
template <class M>
FC_DI typename M::T flux_mass_evap_rco(M &m, const typename M::T &q_a, const typename M::T &T_s,
                                       const typename M::T &vel)
{
    const double rho_a = 1.225, c_aw = 1.15E-03, epsilon = 0.62197, P_0 = 1.013E+05;   // :135-138
    const double r = 6.1078E+02, c_1 = 17.269, c_2 = 35.86;                            // :143-145
    const auto e_w = M::mul(M::bc(r), m.exp(m.div(M::mul(M::bc(c_1), M::sub(T_s, M::bc(273.15))),
                                                  M::sub(T_s, M::bc(c_2)))));          // :148
    const auto q_w = m.div(M::mul(M::bc(epsilon), e_w), M::bc(P_0));                   // :151
    return M::mul(M::mul(M::bc(rho_a * c_aw), vel), M::sub(q_w, q_a));                 // :156
}

// heat/flux_heat_sensible.F90:84-98
template <class M>
FC_DI typename M::T flux_heat_sensible_cclm(M &m, const typename M::T &a_moisture, const typename M::T &p_a,
                                            const typename M::T &p_s, const typename M::T &q_s,
                                            const typename M::T &T_a, const typename M::T &T_s,
                                            const typename M::T &vel, const Consts &c)
{
    const auto T_tilde = t_tilde<M>(T_s, q_s, c);                                 // :84-86
    const auto flux_air = m.div(M::mul(M::mul(a_moisture, M::max(vel, M::bc(c.u_min_evap))), p_s),
                                M::mul(M::bc(c.gas_constant_air), T_tilde));     // :90-92
    const auto EF = m.powc(m.div(p_s, p_a), c.rd_over_cp);                        // :94-95
    return M::mul(M::mul(flux_air, M::bc(c.heat_capacity_air)), M::sub(T_s, M::mul(T_a, EF)));   // :97-98
}

// heat/flux_heat_sensible.F90:157-165
template <class M>
FC_DI typename M::T flux_heat_sensible_rco(const typename M::T &T_a, const typename M::T &T_s,
                                           const typename M::T &vel)
{
    const double rho_a = 1.225, c_pa = 1.008E+03;                                 // :151-152
    const auto c_aw = M::sel_lt(T_a, T_s, M::bc(1.13E-03), M::bc(0.66E-03));      // :157-161
    return M::mul(M::mul(M::mul(M::bc(rho_a * c_pa), c_aw), vel), M::sub(T_s, T_a));   // :165
}

// momentum/flux_momentum.F90:63-73; returns flux_air, the caller forms -flux_air*u / -flux_air*v
template <class M>
FC_DI typename M::T momentum_flux_air_cclm(M &m, const typename M::T &a_momentum, const typename M::T &p_s,
                                           const typename M::T &q_s, const typename M::T &T_s,
                                           const typename M::T &vel, const Consts &c)
{
    const auto T_tilde = t_tilde<M>(T_s, q_s, c);                                 // :63-65
    return m.div(M::mul(M::mul(a_momentum, vel), p_s), M::mul(M::bc(c.gas_constant_air), T_tilde));   // :69-70
}

// momentum/flux_momentum.F90:126-136; returns rho_a*c_aw*vel
template <class M>
FC_DI typename M::T momentum_flux_air_rco(const typename M::T &vel)
{
    const double rho_a = 1.225;                                                   // :122
    const auto c_aw = M::sel_lt(vel, M::bc(11.0), M::bc(1.2E-03),
                                M::add(M::bc(0.49E-03), M::mul(M::bc(0.065E-03), vel)));   // :129-133
    return M::mul(M::mul(M::bc(rho_a), c_aw), vel);
}

template <class M>
FC_DI typename M::T momentum_component(const typename M::T &flux_air, const typename M::T &wind)
{
    return M::neg(M::mul(flux_air, wind));                                        // :72-73 / :135-136
}

// radiation/flux_radiation_blackbody.F90:40: sigma * T**4, integer power == (T*T)*(T*T)
template <class M>
FC_DI typename M::T flux_radiation_blackbody_StBo(const typename M::T &T_s, double sigma)
{
    const auto T2 = M::mul(T_s, T_s);
    return M::mul(M::bc(sigma), M::mul(T2, T2));
}

}  // namespace fc
