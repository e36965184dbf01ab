/*
 * flux_oracle.c -- CPU restatement of the IOW-ESM flux_calculator hot path.
 * TEST INFRASTRUCTURE ONLY (see flux_oracle.h): only tests/, __graft_entry__.smoke() and the
 * CPU legs of bench.py may load it; the product library never does.
 * PARITY UNPINNED by any reference binary or reference-shipped vector: the reference (Fortran +
 * MPI + netCDF + OASIS3-MCT) cannot be compiled in this image and ships no tests.  What pins this
 * file instead: tests/golden/flux_lib_golden.json, produced by tests/golden/make_golden.py, which
 * parses and evaluates the reference's own Fortran text of flux_lib/ and the call-site wiring of
 * flux_calculator_calculate.F90 (bit-for-bit agreement), plus 50-digit mpmath known answers
 * (tests/golden/kat_mpmath.json).  See DESIGN.md section 3.
 *
 * All arithmetic is binary64 in the Fortran evaluation order (left to right within
 * equal precedence, parentheses as written in the reference).  Compile with
 * -ffp-contract=off so no multiply-add is fused (== ifort -fp-model precise).
 * Reference paths are relative to /root/reference/src.
 */
#include "flux_oracle.h"

#include <math.h>
#include <string.h>

/* ---- default_values: flux_lib/constants/flux_constants.F90:13-32 ---- */
static const double DEF_heat_capacity_air         = 1005.0;
static const double DEF_latent_heat_vaporization  = 2.501e6;
static const double DEF_latent_heat_sublimation   = 2.835e6;
static const double DEF_gas_constant_air          = 287.05;
static const double DEF_gas_constant_vapor        = 461.51;
static const double DEF_stefan_boltzmann_constant = 5.67e-8;
static const double DEF_u_min_evap                = 0.01;

#define OPT(ptr, dflt) ((ptr) ? *(ptr) : (dflt))   /* IF (PRESENT(x)) ... ELSE default */

/* flux_lib/auxiliaries/flux_aux_vapor.F90:20-70 */
void orc_spec_vapor_surface_cclm(double *q_s, const double *f_ice, const double *p_s, const double *T_s,
                                 const double *R_d_new, const double *R_v_new)
{
    const double alpha_water = 17.2693882, alpha_ice = 21.8745584;   /* :39-40 */
    const double T_1 = 273.16, T_2_water = 35.86, T_2_ice = 7.66;     /* :41-44 */
    const double p_0 = 610.78;                                        /* :45 */
    double R_d = OPT(R_d_new, DEF_gas_constant_air);                  /* :48-52 */
    double R_v = OPT(R_v_new, DEF_gas_constant_vapor);                /* :53-57 */
    double alpha = alpha_water + (alpha_ice - alpha_water) * *f_ice;  /* :60 */
    double T_2   = T_2_water + (T_2_ice - T_2_water) * *f_ice;        /* :61 */
    double e_sat = p_0 * exp(alpha * (*T_s - T_1) / (*T_s - T_2));    /* :63-64 */
    *q_s = (R_d / R_v) * e_sat / (*p_s - (1.0 - R_d / R_v) * e_sat);  /* :66-68 */
}

/* flux_lib/mass/flux_mass_evap.F90:22-85 */
void orc_flux_mass_evap_cclm(double *evap, const double *a_moisture, const double *p_s, const double *q_a,
                             const double *q_s, const double *T_s, const double *u_a, const double *v_a,
                             const double *u_min_new, const double *R_d_new, const double *R_v_new)
{
    double u_min = OPT(u_min_new, DEF_u_min_evap);                    /* :55-59 */
    double R_d = OPT(R_d_new, DEF_gas_constant_air);                  /* :60-64 */
    double R_v = OPT(R_v_new, DEF_gas_constant_vapor);                /* :65-69 */
    double T_tilde = *T_s * (1.0 + (R_v / R_d - 1.0) * *q_s);         /* :72-74 */
    double vel = sqrt(*u_a * *u_a + *v_a * *v_a);                     /* :76 */
    double flux_air = *a_moisture * fmax(vel, u_min) * *p_s / (R_d * T_tilde); /* :78-80 */
    *evap = flux_air * (*q_s - *q_a);                                 /* :82-83 */
}

/* flux_lib/mass/flux_mass_evap.F90:87-118 (forwards to _cclm, no optionals) */
void orc_flux_mass_evap_mom5(double *evap, const double *a_moisture, const double *p_s, const double *q_a,
                             const double *q_s, const double *T_s, const double *u_a, const double *v_a)
{
    orc_flux_mass_evap_cclm(evap, a_moisture, p_s, q_a, q_s, T_s, u_a, v_a, 0, 0, 0);
}

/* flux_lib/mass/flux_mass_evap.F90:120-158 (Meier et al. 1999) */
void orc_flux_mass_evap_rco(double *evap, const double *q_a, const double *T_s, const double *u_a,
                            const double *v_a)
{
    const double rho_a = 1.225, c_aw = 1.15E-03, epsilon = 0.62197, P_0 = 1.013E+05; /* :135-138 */
    const double r = 6.1078E+02, c_1 = 17.269, c_2 = 35.86;                           /* :143-145 */
    double e_w = r * exp(c_1 * (*T_s - 273.15) / (*T_s - c_2));       /* :148 */
    double q_w = epsilon * e_w / P_0;                                 /* :151 */
    double vel = sqrt(*u_a * *u_a + *v_a * *v_a);                     /* :153 */
    *evap = rho_a * c_aw * vel * (q_w - *q_a);                        /* :156 */
}

/* flux_lib/heat/flux_heat_latent.F90:23-43 */
void orc_flux_heat_latent_ice(double *hlat, const double *evap, const double *L_s_new)
{
    double L_s = OPT(L_s_new, DEF_latent_heat_sublimation);           /* :35-39 */
    *hlat = *evap * L_s;                                              /* :41 */
}

/* flux_lib/heat/flux_heat_latent.F90:47-67 */
void orc_flux_heat_latent_water(double *hlat, const double *evap, const double *L_v_new)
{
    double L_v = OPT(L_v_new, DEF_latent_heat_vaporization);          /* :59-63 */
    *hlat = *evap * L_v;                                              /* :65 */
}

/* flux_lib/heat/flux_heat_sensible.F90:24-99 */
void orc_flux_heat_sensible_cclm(double *hsen, const double *a_moisture, const double *p_a, const double *p_s,
                                 const double *q_s, const double *T_a, const double *T_s, const double *u_a,
                                 const double *v_a, const double *c_p_new, const double *u_min_new,
                                 const double *R_d_new, const double *R_v_new)
{
    double c_p   = OPT(c_p_new, DEF_heat_capacity_air);               /* :63-67 */
    double u_min = OPT(u_min_new, DEF_u_min_evap);                    /* :68-72 */
    double R_d   = OPT(R_d_new, DEF_gas_constant_air);                /* :73-77 */
    double R_v   = OPT(R_v_new, DEF_gas_constant_vapor);              /* :78-82 */
    double T_tilde = *T_s * (1.0 + (R_v / R_d - 1.0) * *q_s);         /* :84-86 */
    double vel = sqrt(*u_a * *u_a + *v_a * *v_a);                     /* :88 */
    double flux_air = *a_moisture * fmax(vel, u_min) * *p_s / (R_d * T_tilde); /* :90-92 */
    double EF = pow(*p_s / *p_a, R_d / c_p);                          /* :94-95 */
    *hsen = flux_air * c_p * (*T_s - *T_a * EF);                      /* :97-98 */
}

/* flux_lib/heat/flux_heat_sensible.F90:101-135 */
void orc_flux_heat_sensible_mom5(double *hsen, const double *a_moisture, const double *p_a, const double *p_s,
                                 const double *q_s, const double *T_a, const double *T_s, const double *u_a,
                                 const double *v_a)
{
    orc_flux_heat_sensible_cclm(hsen, a_moisture, p_a, p_s, q_s, T_a, T_s, u_a, v_a, 0, 0, 0, 0);
}

/* flux_lib/heat/flux_heat_sensible.F90:137-167 */
void orc_flux_heat_sensible_rco(double *hsen, const double *T_a, const double *T_s, const double *u_a,
                                const double *v_a)
{
    const double rho_a = 1.225, c_pa = 1.008E+03;                     /* :151-152 */
    double c_aw, vel;
    if (*T_a < *T_s) c_aw = 1.13E-03;                                 /* :157-158 unstable */
    else             c_aw = 0.66E-03;                                 /* :159-160 stable */
    vel = sqrt(*u_a * *u_a + *v_a * *v_a);                            /* :163 */
    *hsen = rho_a * c_pa * c_aw * vel * (*T_s - *T_a);                /* :165 */
}

/* flux_lib/momentum/flux_momentum.F90:22-76 */
void orc_flux_momentum_cclm(double *tau_e, double *tau_n, const double *a_momentum, const double *p_s,
                            const double *q_s, const double *T_s, const double *u_a, const double *v_a,
                            const double *R_d_new, const double *R_v_new)
{
    double R_d = OPT(R_d_new, DEF_gas_constant_air);                  /* :51-55 */
    double R_v = OPT(R_v_new, DEF_gas_constant_vapor);                /* :56-60 */
    double T_tilde = *T_s * (1.0 + (R_v / R_d - 1.0) * *q_s);         /* :63-65 */
    double vel = sqrt(*u_a * *u_a + *v_a * *v_a);                     /* :67 */
    double flux_air = *a_momentum * vel * *p_s / (R_d * T_tilde);     /* :69-70 */
    *tau_e = -flux_air * *u_a;                                        /* :72 */
    *tau_n = -flux_air * *v_a;                                        /* :73 */
}

/* flux_lib/momentum/flux_momentum.F90:78-108 */
void orc_flux_momentum_mom5(double *tau_e, double *tau_n, const double *a_momentum, const double *p_s,
                            const double *q_s, const double *T_s, const double *u_a, const double *v_a)
{
    orc_flux_momentum_cclm(tau_e, tau_n, a_momentum, p_s, q_s, T_s, u_a, v_a, 0, 0);
}

/* flux_lib/momentum/flux_momentum.F90:110-138 */
void orc_flux_momentum_rco(double *tau_e, double *tau_n, const double *u_a, const double *v_a)
{
    const double rho_a = 1.225;                                       /* :122 */
    double c_aw;
    double vel = sqrt(*u_a * *u_a + *v_a * *v_a);                     /* :126 */
    if (vel < 11.0) c_aw = 1.2E-03;                                   /* :129-130 */
    else            c_aw = 0.49E-03 + 0.065E-03 * vel;                /* :131-132 */
    *tau_e = -rho_a * c_aw * vel * *u_a;                              /* :135 */
    *tau_n = -rho_a * c_aw * vel * *v_a;                              /* :136 */
}

/* flux_lib/radiation/flux_radiation_blackbody.F90:22-42.  T**4 is an INTEGER power:
 * Fortran compilers expand it by repeated squaring, (T*T)*(T*T). */
void orc_flux_radiation_blackbody_StBo(double *rbbr, const double *T_s, const double *sigma_new)
{
    double sigma = OPT(sigma_new, DEF_stefan_boltzmann_constant);     /* :34-38 */
    double T2 = *T_s * *T_s;
    *rbbr = sigma * (T2 * T2);                                        /* :40 */
}

/* flux_lib/radiation/distribute_radiation_flux.F90:12-26 (albedo factors commented out) */
void orc_distribute_radiation_flux(double *out, const double *flux_avg, const double *albedo_avg,
                                   const double *albedo_type)
{
    (void)albedo_avg; (void)albedo_type;
    *out = *flux_avg;                                                 /* :24 */
}

/* ------------------------------------------------------------------------------------------ */
/* pyfort/datetime_helpers.py:4-13: datetime.strptime(init_date,"%Y%m%d") + timedelta(seconds)  */
static int64_t days_from_civil(int64_t y, int m, int d)
{
    y -= m <= 2;
    int64_t era = (y >= 0 ? y : y - 399) / 400;
    int64_t yoe = y - era * 400;
    int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}

int orc_current_month(int init_date, int64_t seconds)
{
    int y = init_date / 10000, m = (init_date / 100) % 100, d = init_date % 100;
    int64_t day_shift = seconds >= 0 ? seconds / 86400 : -((-seconds + 86399) / 86400);
    int64_t z = days_from_civil(y, m, d) + day_shift + 719468;
    int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    int64_t doe = z - era * 146097;
    int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    int64_t mp = (5 * doy + 2) / 153;
    return (int)(mp < 10 ? mp + 3 : mp - 9);
}

/* ------------------------------------------------------------------------------------------ */
/* per-step calculators: flux_calculator_calculate.F90                                          */

void orc_state_init(orc_state *s)
{
    memset(s, 0, sizeof *s);
    for (int g = 0; g < 4; ++g)
        for (int i = 0; i <= ORC_MAX_SURFACE_TYPES; ++i)
            strcpy(s->which_spec_vapor_surface[g][i], "none");        /* flux_calculator.F90:99-107 */
    for (int i = 0; i <= ORC_MAX_SURFACE_TYPES; ++i) {
        strcpy(s->which_flux_mass_evap[i], "none");
        strcpy(s->which_flux_heat_latent[i], "none");
        strcpy(s->which_flux_heat_sensible[i], "none");
        strcpy(s->which_flux_momentum[i], "none");
        strcpy(s->which_flux_radiation_blackbody[i], "none");
    }
}

#define F(i, g, idx) (s->local_field[i][g].var[idx].field)
#define IS(m, lit) (strcmp((m), (lit)) == 0)

/* flux_calculator_calculate.F90:25-50 */
void orc_calc_spec_vapor_surface(orc_state *s, int which_grid)
{
    const int g = which_grid;
    for (int i = 1; i <= s->num_surface_types; ++i) {
        const char *method = s->which_spec_vapor_surface[g][i];
        if (!IS(method, "none")) {
            if (IS(method, "CCLM")) {
                for (int64_t j = 0; j < s->grid_size[g]; ++j)         /* :41-46 */
                    orc_spec_vapor_surface_cclm(&F(i, g, ORC_QSUR)[j], &F(i, g, ORC_FICE)[j],
                                                &F(i, g, ORC_PSUR)[j], &F(i, g, ORC_TSUR)[j], 0, 0);
            }
        }
    }
}

/* flux_calculator_calculate.F90:54-120 */
void orc_calc_flux_mass_evap(orc_state *s)
{
    int current_month = 1;
    if (s->lcorrections)                                              /* :66-73 */
        current_month = orc_current_month(s->init_date, s->current_step_time);
    for (int i = 1; i <= s->num_surface_types; ++i) {
        const char *method = s->which_flux_mass_evap[i];
        const int64_t n = s->grid_size[1];
        if (!IS(method, "none")) {
            if (IS(method, "zero")) {
                for (int64_t j = 0; j < n; ++j) F(i, 1, ORC_MEVA)[j] = 0.0;   /* :79 */
            } else if (IS(method, "CCLM")) {
                for (int64_t j = 0; j < n; ++j)                       /* :81-90; T slot <- TATM */
                    orc_flux_mass_evap_cclm(&F(i, 1, ORC_MEVA)[j], &F(i, 1, ORC_AMOI)[j], &F(i, 1, ORC_PSUR)[j],
                                            &F(i, 1, ORC_QATM)[j], &F(i, 1, ORC_QSUR)[j], &F(i, 1, ORC_TATM)[j],
                                            &F(i, 1, ORC_UATM)[j], &F(i, 1, ORC_VATM)[j], 0, 0, 0);
            } else if (IS(method, "MOM5")) {
                for (int64_t j = 0; j < n; ++j)                       /* :92-101 */
                    orc_flux_mass_evap_mom5(&F(i, 1, ORC_MEVA)[j], &F(i, 1, ORC_CMOI)[j], &F(i, 1, ORC_PSUR)[j],
                                            &F(i, 1, ORC_QATM)[j], &F(i, 1, ORC_QSUR)[j], &F(i, 1, ORC_TATM)[j],
                                            &F(i, 1, ORC_UATM)[j], &F(i, 1, ORC_VATM)[j]);
            } else if (IS(method, "RCO")) {
                for (int64_t j = 0; j < n; ++j)                       /* :103-109 */
                    orc_flux_mass_evap_rco(&F(i, 1, ORC_MEVA)[j], &F(i, 1, ORC_QATM)[j], &F(i, 1, ORC_TSUR)[j],
                                           &F(i, 1, ORC_UATM)[j], &F(i, 1, ORC_VATM)[j]);
            }
            if (s->lcorrections) {                                    /* :112-116, stride 12 in j */
                for (int64_t j = 0; j < n; ++j)
                    F(i, 1, ORC_MEVA)[j] = F(i, 1, ORC_MEVA)[j] + s->corrections[(current_month - 1) + 12 * j];
            }
        }
    }
}

/* flux_calculator_calculate.F90:124-154 */
void orc_calc_flux_heat_latent(orc_state *s)
{
    for (int i = 1; i <= s->num_surface_types; ++i) {
        const char *method = s->which_flux_heat_latent[i];
        const int64_t n = s->grid_size[1];
        if (!IS(method, "none")) {
            if (IS(method, "zero")) {
                for (int64_t j = 0; j < n; ++j) F(i, 1, ORC_HLAT)[j] = 0.0;   /* :139 */
            } else if (IS(method, "water")) {
                for (int64_t j = 0; j < n; ++j)                       /* :141-144 */
                    orc_flux_heat_latent_water(&F(i, 1, ORC_HLAT)[j], &F(i, 1, ORC_MEVA)[j], 0);
            } else if (IS(method, "ice")) {
                for (int64_t j = 0; j < n; ++j)                       /* :146-149 */
                    orc_flux_heat_latent_ice(&F(i, 1, ORC_HLAT)[j], &F(i, 1, ORC_MEVA)[j], 0);
            }
        }
    }
}

/* flux_calculator_calculate.F90:156-208 */
void orc_calc_flux_heat_sensible(orc_state *s)
{
    for (int i = 1; i <= s->num_surface_types; ++i) {
        const char *method = s->which_flux_heat_sensible[i];
        const int64_t n = s->grid_size[1];
        if (!IS(method, "none")) {
            if (IS(method, "zero")) {
                for (int64_t j = 0; j < n; ++j) F(i, 1, ORC_HSEN)[j] = 0.0;   /* :171 */
            } else if (IS(method, "CCLM")) {
                for (int64_t j = 0; j < n; ++j)                       /* :173-183; q_s slot <- QATM */
                    orc_flux_heat_sensible_cclm(&F(i, 1, ORC_HSEN)[j], &F(i, 1, ORC_AMOI)[j], &F(i, 1, ORC_PATM)[j],
                                                &F(i, 1, ORC_PSUR)[j], &F(i, 1, ORC_QATM)[j], &F(i, 1, ORC_TATM)[j],
                                                &F(i, 1, ORC_TSUR)[j], &F(i, 1, ORC_UATM)[j], &F(i, 1, ORC_VATM)[j],
                                                0, 0, 0, 0);
            } else if (IS(method, "MOM5")) {
                for (int64_t j = 0; j < n; ++j)                       /* :185-195 */
                    orc_flux_heat_sensible_mom5(&F(i, 1, ORC_HSEN)[j], &F(i, 1, ORC_CHEA)[j], &F(i, 1, ORC_PATM)[j],
                                                &F(i, 1, ORC_PSUR)[j], &F(i, 1, ORC_QATM)[j], &F(i, 1, ORC_TATM)[j],
                                                &F(i, 1, ORC_TSUR)[j], &F(i, 1, ORC_UATM)[j], &F(i, 1, ORC_VATM)[j]);
            } else if (IS(method, "RCO")) {
                for (int64_t j = 0; j < n; ++j)                       /* :197-203 */
                    orc_flux_heat_sensible_rco(&F(i, 1, ORC_HSEN)[j], &F(i, 1, ORC_TATM)[j], &F(i, 1, ORC_TSUR)[j],
                                               &F(i, 1, ORC_UATM)[j], &F(i, 1, ORC_VATM)[j]);
            }
        }
    }
}

/* flux_calculator_calculate.F90:212-263 (east) and :265-316 (north) share this body; the
 * component that is not wanted lands in the local `dummy` (:223, :276). */
static void calc_flux_momentum(orc_state *s, int which_grid, int north)
{
    const int g = which_grid;
    const int out_idx = north ? ORC_VMOM : ORC_UMOM;
    double dummy;
    for (int i = 1; i <= s->num_surface_types; ++i) {
        const char *method = s->which_flux_momentum[i];
        const int64_t n = s->grid_size[g];
        if (!IS(method, "none")) {
            if (IS(method, "zero")) {
                for (int64_t j = 0; j < n; ++j) F(i, g, out_idx)[j] = 0.0;    /* :229 / :282 */
            } else if (IS(method, "CCLM")) {
                for (int64_t j = 0; j < n; ++j) {                     /* :231-240 / :284-293 */
                    double *o = &F(i, g, out_idx)[j];
                    orc_flux_momentum_cclm(north ? &dummy : o, north ? o : &dummy, &F(i, g, ORC_AMOM)[j],
                                           &F(i, g, ORC_PSUR)[j], &F(i, g, ORC_QSUR)[j], &F(i, g, ORC_TSUR)[j],
                                           &F(i, g, ORC_UATM)[j], &F(i, g, ORC_VATM)[j], 0, 0);
                }
            } else if (IS(method, "MOM5")) {
                for (int64_t j = 0; j < n; ++j) {                     /* :242-251 / :295-304 */
                    double *o = &F(i, g, out_idx)[j];
                    orc_flux_momentum_mom5(north ? &dummy : o, north ? o : &dummy, &F(i, g, ORC_CMOM)[j],
                                           &F(i, g, ORC_PSUR)[j], &F(i, g, ORC_QSUR)[j], &F(i, g, ORC_TSUR)[j],
                                           &F(i, g, ORC_UATM)[j], &F(i, g, ORC_VATM)[j]);
                }
            } else if (IS(method, "RCO")) {
                for (int64_t j = 0; j < n; ++j) {                     /* :253-258 / :306-311 */
                    double *o = &F(i, g, out_idx)[j];
                    orc_flux_momentum_rco(north ? &dummy : o, north ? o : &dummy,
                                          &F(i, g, ORC_UATM)[j], &F(i, g, ORC_VATM)[j]);
                }
            }
        }
    }
}

void orc_calc_flux_momentum_east(orc_state *s, int which_grid)  { calc_flux_momentum(s, which_grid, 0); }
void orc_calc_flux_momentum_north(orc_state *s, int which_grid) { calc_flux_momentum(s, which_grid, 1); }

/* flux_calculator_calculate.F90:320-345 */
void orc_calc_flux_radiation_blackbody(orc_state *s)
{
    for (int i = 1; i <= s->num_surface_types; ++i) {
        const char *method = s->which_flux_radiation_blackbody[i];
        const int64_t n = s->grid_size[1];
        if (!IS(method, "none")) {
            if (IS(method, "zero")) {
                for (int64_t j = 0; j < n; ++j) F(i, 1, ORC_RBBR)[j] = 0.0;   /* :335 */
            } else if (IS(method, "StBo")) {
                for (int64_t j = 0; j < n; ++j)                       /* :337-340 */
                    orc_flux_radiation_blackbody_StBo(&F(i, 1, ORC_RBBR)[j], &F(i, 1, ORC_TSUR)[j], 0);
            }
        }
    }
}

/* flux_calculator_calculate.F90:347-364.  The reference runs this unconditionally
 * (flux_calculator.F90:991) and is undefined when the arrays do not exist;
 * the oracle runs it only when state.distribute_shortwave is set. */
void orc_distribute_shortwave_radiation_flux(orc_state *s)
{
    for (int i = 1; i <= s->num_surface_types; ++i)
        for (int64_t j = 0; j < s->grid_size[1]; ++j)                 /* :355-360 */
            orc_distribute_radiation_flux(&F(i, 1, ORC_RSDR)[j], &F(0, 1, ORC_RSDD)[j],
                                          &F(0, 1, ORC_ALBA)[j], &F(i, 1, ORC_ALBE)[j]);
}

/* flux_calculator_calculate.F90:368-385 */
void orc_average_across_surface_types(orc_state *s, int which_grid, int my_idx)
{
    const int g = which_grid;
    if (s->local_field[0][g].var[my_idx].allocated) {                 /* :376 */
        double *acc = F(0, g, my_idx);
        for (int64_t j = 0; j < s->grid_size[g]; ++j) acc[j] = 0.0;   /* :377 */
        for (int i = 1; i <= s->num_surface_types; ++i)
            for (int64_t j = 0; j < s->grid_size[g]; ++j)             /* :379-382: separate * and + */
                acc[j] = acc[j] + F(i, g, my_idx)[j] * F(i, g, ORC_FARE)[j];
    }
}

/* send loop: flux_calculator.F90:909-936 (early) / :999-1026 (normal); only the averaging
 * trigger (:912-919 / :1002-1009) is arithmetic, oasis_put is host transport. */
static void send_loop(orc_state *s, int early)
{
    for (int g = 1; g <= 3; ++g)
        for (int j = 0; j < s->num_output_fields; ++j) {
            const orc_output_field *o = &s->output_field[j];
            if (o->which_grid == g && o->early == early && o->surface_type == 0)
                if (F(0, g, o->idx) && F(2, g, o->idx))
                    orc_average_across_surface_types(s, g, o->idx);
        }
}

void orc_step_early(orc_state *s)
{
    orc_calc_flux_radiation_blackbody(s);                             /* flux_calculator.F90:902 */
    send_loop(s, 1);
}

void orc_step_normal(orc_state *s)
{
    orc_calc_spec_vapor_surface(s, 1);                                /* :972 */
    orc_calc_spec_vapor_surface(s, 2);                                /* :973 */
    orc_calc_spec_vapor_surface(s, 3);                                /* :974 */
    orc_calc_flux_mass_evap(s);                                       /* :977 */
    orc_calc_flux_heat_latent(s);                                     /* :980 */
    orc_calc_flux_heat_sensible(s);                                   /* :983 */
    orc_calc_flux_momentum_east(s, 2);                                /* :986 */
    orc_calc_flux_momentum_north(s, 3);                               /* :988 */
    if (s->distribute_shortwave)
        orc_distribute_shortwave_radiation_flux(s);                   /* :991 */
    send_loop(s, 0);
}

/* flux_calculator_basic.F90:476-486: dst = 0; sequential COO accumulation in element order */
void orc_regrid(double *dst, int64_t n_dst, const double *src, const orc_sparse_matrix *m)
{
    for (int64_t j = 0; j < n_dst; ++j) dst[j] = 0.0;
    for (int64_t k = 0; k < m->num_elements; ++k)
        dst[m->dst_index[k] - 1] = dst[m->dst_index[k] - 1] + src[m->src_index[k] - 1] * m->weight[k];
}

/* decomp_def.F90:14-31 with id_im = 1 (1-D exchange grid), id_jm = n */
void orc_decomp_apple(int64_t n, int rank, int npes, int64_t *offset, int64_t *size)
{
    int64_t part = n / npes;                                          /* il_partj, :15 */
    *offset = rank * part;                                            /* :24 / :28 */
    *size = rank < npes - 1 ? part : n - rank * part;                 /* :25 / :29 */
}
