/*
 * cpu_baseline.c -- handle-style setters for the oracle state (so Python/ctypes never mirrors the
 * struct) and the multi-rank CPU arm: P independent "flux_calculator instances", each owning one
 * contiguous range of every exchange grid, no communication -- the reference's only parallelism
 * (flux_calculator_io.F90:77-107, one MPI rank per range).  TEST / BENCH INFRASTRUCTURE ONLY.
 */
#include "flux_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

orc_state *orc_state_new(int num_surface_types, const int64_t grid_size[3])
{
    orc_state *s = (orc_state *)malloc(sizeof *s);
    if (!s) return 0;
    orc_state_init(s);
    s->num_surface_types = num_surface_types;
    for (int g = 1; g <= 3; ++g) s->grid_size[g] = grid_size[g - 1];
    return s;
}

void orc_state_free(orc_state *s) { free(s); }

/* allocated != 0: own storage (ALLOCATE); 0: pointer alias (=>) */
int orc_bind(orc_state *s, int surface_type, int grid, int idx, double *p, int allocated)
{
    if (surface_type < 0 || surface_type > ORC_MAX_SURFACE_TYPES || grid < 1 || grid > 3 || idx < 1 ||
        idx > ORC_MAX_VARNAMES)
        return 1;
    s->local_field[surface_type][grid].var[idx].field = p;
    s->local_field[surface_type][grid].var[idx].allocated = allocated;
    return 0;
}

/* quantity: "QSUR" (needs grid 1..3), "MEVA", "HLAT", "HSEN", "MOM", "RBBR" */
int orc_set_method(orc_state *s, const char *quantity, int grid, int surface_type, const char *method)
{
    char *dst = 0;
    if (surface_type < 1 || surface_type > ORC_MAX_SURFACE_TYPES || strlen(method) > ORC_METHOD_LEN) return 1;
    if (!strcmp(quantity, "QSUR")) {
        if (grid < 1 || grid > 3) return 1;
        dst = s->which_spec_vapor_surface[grid][surface_type];
    } else if (!strcmp(quantity, "MEVA")) dst = s->which_flux_mass_evap[surface_type];
    else if (!strcmp(quantity, "HLAT")) dst = s->which_flux_heat_latent[surface_type];
    else if (!strcmp(quantity, "HSEN")) dst = s->which_flux_heat_sensible[surface_type];
    else if (!strcmp(quantity, "MOM"))  dst = s->which_flux_momentum[surface_type];
    else if (!strcmp(quantity, "RBBR")) dst = s->which_flux_radiation_blackbody[surface_type];
    else return 1;
    strcpy(dst, method);
    return 0;
}

void orc_set_corrections(orc_state *s, const double *corr_fortran_1_12_n, int enabled, int init_date)
{
    s->corrections = corr_fortran_1_12_n;
    s->lcorrections = enabled;
    s->init_date = init_date;
}

void orc_set_time(orc_state *s, int64_t seconds) { s->current_step_time = seconds; }
void orc_set_distribute_shortwave(orc_state *s, int on) { s->distribute_shortwave = on; }

int orc_add_output(orc_state *s, int surface_type, int grid, int idx)
{
    if (s->num_output_fields >= 256) return 1;
    orc_output_field *o = &s->output_field[s->num_output_fields++];
    o->surface_type = surface_type;
    o->which_grid = grid;
    o->idx = idx;
    /* basic.F90:271-273 */
    o->early = (idx == ORC_RBBR || idx == ORC_TSUR || idx == ORC_FICE || idx == ORC_ALBE);
    return 0;
}

/* view of rank `rank` of `npes`: every bound array and the corrections slab are offset by the
 * rank's (offset,size) on that grid (decomp_def.F90 APPLE rule) */
static void shard_view(const orc_state *full, orc_state *v, int rank, int npes)
{
    *v = *full;
    for (int g = 1; g <= 3; ++g) {
        int64_t off, size;
        orc_decomp_apple(full->grid_size[g], rank, npes, &off, &size);
        v->grid_size[g] = size;
        for (int i = 0; i <= ORC_MAX_SURFACE_TYPES; ++i)
            for (int k = 1; k <= ORC_MAX_VARNAMES; ++k)
                if (v->local_field[i][g].var[k].field) v->local_field[i][g].var[k].field += off;
        if (g == 1 && v->corrections) v->corrections += 12 * off;
    }
}

/* nsteps coupling steps (early + normal phase each) on P independent ranks (one thread each) */
typedef struct { const orc_state *full; int rank, npes, nsteps; int64_t timestep; } rank_arg;

static void *rank_main(void *p)
{
    rank_arg *a = (rank_arg *)p;
    orc_state *v = (orc_state *)malloc(sizeof *v);
    shard_view(a->full, v, a->rank, a->npes);
    for (int n = 0; n < a->nsteps; ++n) {
        v->current_step_time = a->full->current_step_time + (int64_t)n * a->timestep;
        orc_step_early(v);
        orc_step_normal(v);
    }
    free(v);
    return 0;
}

int orc_run_ranks(const orc_state *full, int npes, int nsteps, int64_t timestep)
{
    if (npes < 1) return 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * npes);
    rank_arg *args = (rank_arg *)malloc(sizeof(rank_arg) * npes);
    for (int r = 0; r < npes; ++r) {
        args[r] = (rank_arg){full, r, npes, nsteps, timestep};
        if (r > 0 && pthread_create(&th[r], 0, rank_main, &args[r])) return 2;
    }
    rank_main(&args[0]);
    for (int r = 1; r < npes; ++r) pthread_join(th[r], 0);
    free(th);
    free(args);
    return 0;
}

int orc_max_threads(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }

/* ---- array drivers of the level-0 routines (one scalar call per cell, like the reference's loops);
 *      used by the tests to check the Level-1 C ABI ---- */
#define LOOP for (int64_t j = 0; j < n; ++j)
void orc_v_spec_vapor_surface_cclm(double *q, const double *f, const double *p, const double *T, int64_t n,
                                   const double *Rd, const double *Rv)
{ LOOP orc_spec_vapor_surface_cclm(q + j, f + j, p + j, T + j, Rd, Rv); }
void orc_v_flux_mass_evap_cclm(double *e, const double *a, const double *ps, const double *qa, const double *qs,
                               const double *T, const double *u, const double *v, int64_t n, const double *umin,
                               const double *Rd, const double *Rv)
{ LOOP orc_flux_mass_evap_cclm(e + j, a + j, ps + j, qa + j, qs + j, T + j, u + j, v + j, umin, Rd, Rv); }
void orc_v_flux_mass_evap_mom5(double *e, const double *a, const double *ps, const double *qa, const double *qs,
                               const double *T, const double *u, const double *v, int64_t n)
{ LOOP orc_flux_mass_evap_mom5(e + j, a + j, ps + j, qa + j, qs + j, T + j, u + j, v + j); }
void orc_v_flux_mass_evap_rco(double *e, const double *qa, const double *T, const double *u, const double *v, int64_t n)
{ LOOP orc_flux_mass_evap_rco(e + j, qa + j, T + j, u + j, v + j); }
void orc_v_flux_heat_latent_ice(double *h, const double *e, int64_t n, const double *L)
{ LOOP orc_flux_heat_latent_ice(h + j, e + j, L); }
void orc_v_flux_heat_latent_water(double *h, const double *e, int64_t n, const double *L)
{ LOOP orc_flux_heat_latent_water(h + j, e + j, L); }
void orc_v_flux_heat_sensible_cclm(double *h, const double *a, const double *pa, const double *ps, const double *q,
                                   const double *Ta, const double *Ts, const double *u, const double *v, int64_t n,
                                   const double *cp, const double *umin, const double *Rd, const double *Rv)
{ LOOP orc_flux_heat_sensible_cclm(h + j, a + j, pa + j, ps + j, q + j, Ta + j, Ts + j, u + j, v + j, cp, umin, Rd, Rv); }
void orc_v_flux_heat_sensible_mom5(double *h, const double *a, const double *pa, const double *ps, const double *q,
                                   const double *Ta, const double *Ts, const double *u, const double *v, int64_t n)
{ LOOP orc_flux_heat_sensible_mom5(h + j, a + j, pa + j, ps + j, q + j, Ta + j, Ts + j, u + j, v + j); }
void orc_v_flux_heat_sensible_rco(double *h, const double *Ta, const double *Ts, const double *u, const double *v, int64_t n)
{ LOOP orc_flux_heat_sensible_rco(h + j, Ta + j, Ts + j, u + j, v + j); }
void orc_v_flux_momentum_cclm(double *te, double *tn, const double *a, const double *ps, const double *q, const double *T,
                              const double *u, const double *v, int64_t n, const double *Rd, const double *Rv)
{ double d; LOOP orc_flux_momentum_cclm(te ? te + j : &d, tn ? tn + j : &d, a + j, ps + j, q + j, T + j, u + j, v + j, Rd, Rv); }
void orc_v_flux_momentum_mom5(double *te, double *tn, const double *a, const double *ps, const double *q, const double *T,
                              const double *u, const double *v, int64_t n)
{ double d; LOOP orc_flux_momentum_mom5(te ? te + j : &d, tn ? tn + j : &d, a + j, ps + j, q + j, T + j, u + j, v + j); }
void orc_v_flux_momentum_rco(double *te, double *tn, const double *u, const double *v, int64_t n)
{ double d; LOOP orc_flux_momentum_rco(te ? te + j : &d, tn ? tn + j : &d, u + j, v + j); }
void orc_v_flux_radiation_blackbody_StBo(double *r, const double *T, int64_t n, const double *sigma)
{ LOOP orc_flux_radiation_blackbody_StBo(r + j, T + j, sigma); }
void orc_v_distribute_radiation_flux(double *o, const double *f, const double *a1, const double *a2, int64_t n)
{ double z = 0.0; LOOP orc_distribute_radiation_flux(o + j, f + j, a1 ? a1 + j : &z, a2 ? a2 + j : &z); }
void orc_v_regrid(double *dst, int64_t n_dst, const double *src, int64_t nnz, const int32_t *s, const int32_t *d,
                  const double *w)
{ orc_sparse_matrix m = {nnz, s, d, w}; orc_regrid(dst, n_dst, src, &m); }
