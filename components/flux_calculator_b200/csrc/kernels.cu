// kernels.cu -- hand-written sm_100a kernels of the flux calculator hot path.
//
//  fused_step_kernel   one launch per coupling-step phase: t-, u- and v-grid chains of every surface
//                      type in one pass over SoA fields, all intermediates (QSUR, MEVA, vel, T~,
//                      flux_air) in registers, 128-bit coalesced loads/stores, optional diagnostics
//                      epilogue (warp shuffle -> shared -> per-block partials).
//  oplist_kernel       interpreter for the reference's pass sequence (exact semantics for aliased
//                      outputs / unfused calc_* calls / the Level-1 flux_lib array routines).
//  diag_finalize, transpose_corrections, regrid_csr: small helpers.
//
// The path is elementwise FP64 streaming: HBM-bound, no reuse across cells, so no tensor cores, no
// shared-memory staging (ncu: see profiles/).
#include "plan.h"

#include <cuda_runtime.h>
#include <float.h>

namespace fc {

// ---------------------------------------------------------------------------------------------
// vector helpers
// ---------------------------------------------------------------------------------------------
struct V2 {
    double v[kFusedVec];
};

template <bool AL>
__device__ __forceinline__ V2 ldv(const double *__restrict__ p, int64_t j, int nv)
{
    V2 r;
    if (AL && nv == kFusedVec) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p + j));
        r.v[0] = t.x;
        r.v[1] = t.y;
    } else {
#pragma unroll
        for (int k = 0; k < kFusedVec; ++k) r.v[k] = (k < nv) ? __ldg(p + j + k) : 1.0;
    }
    return r;
}

template <bool AL>
__device__ __forceinline__ void stv(double *p, int64_t j, int nv, const V2 &x)
{
    if (AL && nv == kFusedVec) {
        *reinterpret_cast<double2 *>(p + j) = make_double2(x.v[0], x.v[1]);
    } else {
#pragma unroll
        for (int k = 0; k < kFusedVec; ++k)
            if (k < nv) p[j + k] = x.v[k];
    }
}

// a field loaded at most once per distinct pointer: atmosphere fields are aliased into every
// surface type (distribute_input_field, basic.F90:334-358), so consecutive types usually share them
template <bool AL>
struct Cached {
    const double *ptr = nullptr;
    V2 val;
    __device__ __forceinline__ const V2 &get(const double *p, int64_t j, int nv)
    {
        if (p != ptr) {
            ptr = p;
            if (p) val = ldv<AL>(p, j, nv);
        }
        return val;
    }
};

// ---------------------------------------------------------------------------------------------
// diagnostics accumulation: (sum area*x, min, max) per slot
// ---------------------------------------------------------------------------------------------
struct DiagAcc {
    double s, mn, mx;
};

__device__ __forceinline__ DiagAcc diag_cells(const V2 &x, const V2 &area, int nv)
{
    DiagAcc a{0.0, DBL_MAX, -DBL_MAX};
#pragma unroll
    for (int k = 0; k < kFusedVec; ++k)
        if (k < nv) {
            a.s = add(a.s, mul(area.v[k], x.v[k]));
            a.mn = fmin(a.mn, x.v[k]);
            a.mx = fmax(a.mx, x.v[k]);
        }
    return a;
}

__device__ __forceinline__ void diag_warp_commit(DiagAcc a, double *smem_slot /* [warps][3] base of this slot */)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a.s = add(a.s, __shfl_down_sync(0xffffffffu, a.s, off));
        a.mn = fmin(a.mn, __shfl_down_sync(0xffffffffu, a.mn, off));
        a.mx = fmax(a.mx, __shfl_down_sync(0xffffffffu, a.mx, off));
    }
    if ((threadIdx.x & 31) == 0) {
        double *d = smem_slot + (threadIdx.x >> 5) * 3;
        d[0] = a.s;
        d[1] = a.mn;
        d[2] = a.mx;
    }
}

constexpr int kWarps = kFusedThreads / 32;
// shared layout: [compact slot][warp][3]; p.diag_map[slot] = compact index of an active slot
#define DIAG_SMEM(slot) (diag_smem + (size_t)(p.diag_map[(slot)]) * kWarps * 3)

// ---------------------------------------------------------------------------------------------
// fused chains
// ---------------------------------------------------------------------------------------------
template <int SS, bool DIAG, bool AL>
__device__ __forceinline__ void t_chain(const FusedPlan &p, int blk, double *diag_smem)
{
    const FusedT &t = p.t;
    const Consts &c = p.c;
    const int64_t end = p.cell0[0] + p.cells[0];
    const int64_t j = p.cell0[0] + ((int64_t)blk * kFusedThreads + threadIdx.x) * kFusedVec;
    const int nv = (j >= end) ? 0 : (end - j >= kFusedVec ? kFusedVec : (int)(end - j));
    if (!DIAG && nv == 0) return;
    const int S = SS ? SS : p.S;

    Cached<AL> cPSUR, cQATM, cTATM, cPATM, cUATM, cVATM, cAEV, cASE, cFICE, cTSUR;
    V2 bias, rsdd, area;
    const bool has_bias = p.do_normal && t.bias != nullptr;
    const bool has_rsdr = p.do_normal && t.rsdd != nullptr;
    if (nv) {
        if (has_bias) bias = ldv<AL>(t.bias, j, nv);
        if (has_rsdr) rsdd = ldv<AL>(t.rsdd, j, nv);
        if (DIAG) area = ldv<AL>(t.area, j, nv);
    }

    V2 aQ, aM, aL, aH, aR;   // type-0 averages, sequential from 0.0 (calculate.F90:377-383)
#pragma unroll
    for (int k = 0; k < kFusedVec; ++k) aQ.v[k] = aM.v[k] = aL.v[k] = aH.v[k] = aR.v[k] = 0.0;
    V2 aS = aQ;

    auto per_type = [&](const int i) {
        const FusedTType &ty = t.ty[i];
        V2 qsur, meva, hlat, hsen, rbbr, fare;
        if (nv) {
            const V2 &tsur = cTSUR.get(ty.tsur, j, nv);
            if (ty.fare) fare = ldv<AL>(ty.fare, j, nv);
            if (p.do_normal) {
                const V2 &psur = cPSUR.get(ty.psur, j, nv);
                const V2 &qatm = cQATM.get(ty.qatm, j, nv);
                const V2 &tatm = cTATM.get(ty.tatm, j, nv);
                const V2 &uatm = cUATM.get(ty.uatm, j, nv);
                const V2 &vatm = cVATM.get(ty.vatm, j, nv);
                // --- QSUR: calc_spec_vapor_surface (calculate.F90:25-50)
                if (ty.m_qsur == M_CCLM) {
                    const V2 &fice = cFICE.get(ty.fice, j, nv);
#pragma unroll
                    for (int k = 0; k < kFusedVec; ++k)
                        qsur.v[k] = spec_vapor_surface_cclm(fice.v[k], psur.v[k], tsur.v[k], c);
                    stv<AL>(ty.qsur, j, nv, qsur);
                } else if (ty.qsur_in) {
                    qsur = ldv<AL>(ty.qsur_in, j, nv);
                }
                V2 vel;
#pragma unroll
                for (int k = 0; k < kFusedVec; ++k) vel.v[k] = wind_speed(uatm.v[k], vatm.v[k]);
                // --- MEVA: calc_flux_mass_evap (calculate.F90:54-120); T slot <- TATM (:87,:98)
                if (ty.m_meva != M_NONE) {
                    if (ty.m_meva == M_CCLM || ty.m_meva == M_MOM5) {
                        const V2 &a = cAEV.get(ty.a_evap, j, nv);
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k)
                            meva.v[k] = flux_mass_evap_cclm(a.v[k], psur.v[k], qatm.v[k], qsur.v[k], tatm.v[k],
                                                            vel.v[k], c);
                    } else if (ty.m_meva == M_RCO) {
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k)
                            meva.v[k] = flux_mass_evap_rco(qatm.v[k], tsur.v[k], vel.v[k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k) meva.v[k] = 0.0;   // 'zero' (:79)
                    }
                    if (has_bias) {                                            // :112-116
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k) meva.v[k] = add(meva.v[k], bias.v[k]);
                    }
                    stv<AL>(ty.meva, j, nv, meva);
                }
                // --- HLAT: calc_flux_heat_latent (calculate.F90:124-154), sees the corrected MEVA
                if (ty.m_hlat != M_NONE) {
#pragma unroll
                    for (int k = 0; k < kFusedVec; ++k)
                        hlat.v[k] = (ty.m_hlat == M_ZERO) ? 0.0 : flux_heat_latent(meva.v[k], ty.latent_heat);
                    stv<AL>(ty.hlat, j, nv, hlat);
                }
                // --- HSEN: calc_flux_heat_sensible (calculate.F90:156-208); q_s slot <- QATM (:178,:190)
                if (ty.m_hsen != M_NONE) {
                    if (ty.m_hsen == M_CCLM || ty.m_hsen == M_MOM5) {
                        const V2 &a = cASE.get(ty.a_sens, j, nv);
                        const V2 &patm = cPATM.get(ty.patm, j, nv);
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k)
                            hsen.v[k] = flux_heat_sensible_cclm(a.v[k], patm.v[k], psur.v[k], qatm.v[k], tatm.v[k],
                                                                tsur.v[k], vel.v[k], c);
                    } else if (ty.m_hsen == M_RCO) {
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k)
                            hsen.v[k] = flux_heat_sensible_rco(tatm.v[k], tsur.v[k], vel.v[k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k) hsen.v[k] = 0.0;
                    }
                    stv<AL>(ty.hsen, j, nv, hsen);
                }
                // --- RSDR: distribute_shortwave_radiation_flux (calculate.F90:347-364)
                if (has_rsdr) stv<AL>(ty.rsdr, j, nv, rsdd);
            }
            // --- RBBR: calc_flux_radiation_blackbody (calculate.F90:320-345), early phase
            if (p.do_early && ty.m_rbbr != M_NONE) {
#pragma unroll
                for (int k = 0; k < kFusedVec; ++k)
                    rbbr.v[k] = (ty.m_rbbr == M_ZERO)
                                    ? 0.0
                                    : flux_radiation_blackbody_StBo(tsur.v[k], c.stefan_boltzmann_constant);
                stv<AL>(ty.rbbr, j, nv, rbbr);
            }
            // --- average_across_surface_types (calculate.F90:368-385): acc = acc + X(i)*FARE(i)
#pragma unroll
            for (int k = 0; k < kFusedVec; ++k) {
                if (t.avg_qsur) aQ.v[k] = add(aQ.v[k], mul(qsur.v[k], fare.v[k]));
                if (t.avg_meva) aM.v[k] = add(aM.v[k], mul(meva.v[k], fare.v[k]));
                if (t.avg_hlat) aL.v[k] = add(aL.v[k], mul(hlat.v[k], fare.v[k]));
                if (t.avg_hsen) aH.v[k] = add(aH.v[k], mul(hsen.v[k], fare.v[k]));
                if (t.avg_rbbr) aR.v[k] = add(aR.v[k], mul(rbbr.v[k], fare.v[k]));
                if (t.avg_rsdr) aS.v[k] = add(aS.v[k], mul(rsdd.v[k], fare.v[k]));
            }
        }
        if (DIAG) {
            const int base = (i + 1) * DQ_COUNT;
            if (p.do_normal) {
                if (ty.m_qsur == M_CCLM) diag_warp_commit(diag_cells(qsur, area, nv), DIAG_SMEM(base + DQ_QSUR_T));
                if (ty.m_meva != M_NONE) diag_warp_commit(diag_cells(meva, area, nv), DIAG_SMEM(base + DQ_MEVA));
                if (ty.m_hlat != M_NONE) diag_warp_commit(diag_cells(hlat, area, nv), DIAG_SMEM(base + DQ_HLAT));
                if (ty.m_hsen != M_NONE) diag_warp_commit(diag_cells(hsen, area, nv), DIAG_SMEM(base + DQ_HSEN));
                if (has_rsdr) diag_warp_commit(diag_cells(rsdd, area, nv), DIAG_SMEM(base + DQ_RSDR));
            }
            if (p.do_early && ty.m_rbbr != M_NONE)
                diag_warp_commit(diag_cells(rbbr, area, nv), DIAG_SMEM(base + DQ_RBBR));
        }
    };
    if constexpr (SS > 0) {
#pragma unroll
        for (int i = 0; i < SS; ++i) per_type(i);
    } else {
#pragma unroll 1
        for (int i = 0; i < S; ++i) per_type(i);
    }
    if (nv) {
        if (t.avg_qsur) stv<AL>(t.avg_qsur, j, nv, aQ);
        if (t.avg_meva) stv<AL>(t.avg_meva, j, nv, aM);
        if (t.avg_hlat) stv<AL>(t.avg_hlat, j, nv, aL);
        if (t.avg_hsen) stv<AL>(t.avg_hsen, j, nv, aH);
        if (t.avg_rbbr) stv<AL>(t.avg_rbbr, j, nv, aR);
        if (t.avg_rsdr) stv<AL>(t.avg_rsdr, j, nv, aS);
    }
    if (DIAG) {
        if (t.avg_qsur) diag_warp_commit(diag_cells(aQ, area, nv), DIAG_SMEM(DQ_QSUR_T));
        if (t.avg_meva) diag_warp_commit(diag_cells(aM, area, nv), DIAG_SMEM(DQ_MEVA));
        if (t.avg_hlat) diag_warp_commit(diag_cells(aL, area, nv), DIAG_SMEM(DQ_HLAT));
        if (t.avg_hsen) diag_warp_commit(diag_cells(aH, area, nv), DIAG_SMEM(DQ_HSEN));
        if (t.avg_rbbr) diag_warp_commit(diag_cells(aR, area, nv), DIAG_SMEM(DQ_RBBR));
        if (t.avg_rsdr) diag_warp_commit(diag_cells(aS, area, nv), DIAG_SMEM(DQ_RSDR));
    }
}

template <int SS, bool DIAG, bool AL>
__device__ __forceinline__ void uv_chain(const FusedPlan &p, const FusedUV &g, int which, int blk, double *diag_smem)
{
    const Consts &c = p.c;
    const int64_t end = p.cell0[which] + p.cells[which];
    const int64_t j = p.cell0[which] + ((int64_t)blk * kFusedThreads + threadIdx.x) * kFusedVec;
    const int nv = (j >= end) ? 0 : (end - j >= kFusedVec ? kFusedVec : (int)(end - j));
    if (!DIAG && nv == 0) return;
    const int S = SS ? SS : p.S;
    const int dq_q = (which == 1) ? DQ_QSUR_U : DQ_QSUR_V;
    const int dq_m = (which == 1) ? DQ_UMOM : DQ_VMOM;

    Cached<AL> cPSUR, cUATM, cVATM, cAMOM, cFICE, cTSUR;
    V2 area;
    if (DIAG && nv) area = ldv<AL>(g.area, j, nv);
    V2 aQ, aM;
#pragma unroll
    for (int k = 0; k < kFusedVec; ++k) aQ.v[k] = aM.v[k] = 0.0;

    auto per_type = [&](const int i) {
        const FusedUVType &ty = g.ty[i];
        V2 qsur, mom, fare;
        if (nv) {
            if (ty.fare) fare = ldv<AL>(ty.fare, j, nv);
            // --- QSUR on this grid (calculate.F90:25-50, called for grids 2 and 3)
            if (ty.m_qsur == M_CCLM) {
                const V2 &fice = cFICE.get(ty.fice, j, nv);
                const V2 &psur = cPSUR.get(ty.psur, j, nv);
                const V2 &tsur = cTSUR.get(ty.tsur, j, nv);
#pragma unroll
                for (int k = 0; k < kFusedVec; ++k)
                    qsur.v[k] = spec_vapor_surface_cclm(fice.v[k], psur.v[k], tsur.v[k], c);
                stv<AL>(ty.qsur, j, nv, qsur);
            } else if (ty.qsur_in) {
                qsur = ldv<AL>(ty.qsur_in, j, nv);
            }
            // --- momentum: calc_flux_momentum_east / _north (calculate.F90:212-316)
            if (ty.m_mom != M_NONE) {
                if (ty.m_mom == M_ZERO) {
#pragma unroll
                    for (int k = 0; k < kFusedVec; ++k) mom.v[k] = 0.0;
                } else {
                    const V2 &uatm = cUATM.get(ty.uatm, j, nv);
                    const V2 &vatm = cVATM.get(ty.vatm, j, nv);
                    if (ty.m_mom == M_RCO) {
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k) {
                            const double vel = wind_speed(uatm.v[k], vatm.v[k]);
                            mom.v[k] = momentum_component(momentum_flux_air_rco(vel), g.north ? vatm.v[k] : uatm.v[k]);
                        }
                    } else {
                        const V2 &a = cAMOM.get(ty.a_mom, j, nv);
                        const V2 &psur = cPSUR.get(ty.psur, j, nv);
                        const V2 &tsur = cTSUR.get(ty.tsur, j, nv);
#pragma unroll
                        for (int k = 0; k < kFusedVec; ++k) {
                            const double vel = wind_speed(uatm.v[k], vatm.v[k]);
                            const double fa = momentum_flux_air_cclm(a.v[k], psur.v[k], qsur.v[k], tsur.v[k], vel, c);
                            mom.v[k] = momentum_component(fa, g.north ? vatm.v[k] : uatm.v[k]);
                        }
                    }
                }
                stv<AL>(ty.mom, j, nv, mom);
            }
#pragma unroll
            for (int k = 0; k < kFusedVec; ++k) {
                if (g.avg_qsur) aQ.v[k] = add(aQ.v[k], mul(qsur.v[k], fare.v[k]));
                if (g.avg_mom) aM.v[k] = add(aM.v[k], mul(mom.v[k], fare.v[k]));
            }
        }
        if (DIAG) {
            const int base = (i + 1) * DQ_COUNT;
            if (ty.m_qsur == M_CCLM) diag_warp_commit(diag_cells(qsur, area, nv), DIAG_SMEM(base + dq_q));
            if (ty.m_mom != M_NONE) diag_warp_commit(diag_cells(mom, area, nv), DIAG_SMEM(base + dq_m));
        }
    };
    if constexpr (SS > 0) {
#pragma unroll
        for (int i = 0; i < SS; ++i) per_type(i);
    } else {
#pragma unroll 1
        for (int i = 0; i < S; ++i) per_type(i);
    }
    if (nv) {
        if (g.avg_qsur) stv<AL>(g.avg_qsur, j, nv, aQ);
        if (g.avg_mom) stv<AL>(g.avg_mom, j, nv, aM);
    }
    if (DIAG) {
        if (g.avg_qsur) diag_warp_commit(diag_cells(aQ, area, nv), DIAG_SMEM(dq_q));
        if (g.avg_mom) diag_warp_commit(diag_cells(aM, area, nv), DIAG_SMEM(dq_m));
    }
}

template <int SS, bool DIAG, bool AL>
__global__ void __launch_bounds__(kFusedThreads)
fused_step_kernel(const __grid_constant__ FusedPlan p, int nb_t, int nb_u)
{
    extern __shared__ double diag_smem[];
    const int b = blockIdx.x;
    if (DIAG) {
        // neutral elements for every slot this block does not touch
        for (int k = threadIdx.x; k < p.diag_n * kWarps; k += kFusedThreads) {
            diag_smem[k * 3 + 0] = 0.0;
            diag_smem[k * 3 + 1] = DBL_MAX;
            diag_smem[k * 3 + 2] = -DBL_MAX;
        }
        __syncthreads();
    }
    if (b < nb_t) {
        if (p.do_normal || p.do_early) t_chain<SS, DIAG, AL>(p, b, diag_smem);
    } else if (b < nb_t + nb_u) {
        uv_chain<SS, DIAG, AL>(p, p.uv[0], 1, b - nb_t, diag_smem);
    } else {
        uv_chain<SS, DIAG, AL>(p, p.uv[1], 2, b - nb_t - nb_u, diag_smem);
    }
    if (DIAG) {
        __syncthreads();
        // fixed-order combine over the block's warps, one thread per slot
        for (int s = threadIdx.x; s < p.diag_n; s += kFusedThreads) {
            const double *w = diag_smem + (size_t)s * kWarps * 3;
            double sum = 0.0, mn = DBL_MAX, mx = -DBL_MAX;
#pragma unroll
            for (int k = 0; k < kWarps; ++k) {
                sum = add(sum, w[k * 3 + 0]);
                mn = fmin(mn, w[k * 3 + 1]);
                mx = fmax(mx, w[k * 3 + 2]);
            }
            double *o = p.diag_partials + ((size_t)b * p.diag_n + s) * 3;
            o[0] = sum;
            o[1] = mn;
            o[2] = mx;
        }
    }
}

// partials[blocks][slots][3] -> out[slots][3]; blocks are combined in index order (deterministic)
__global__ void diag_finalize_kernel(const double *__restrict__ partials, int nblocks, int nslots,
                                     double *__restrict__ out)
{
    __shared__ double sh[3][256];
    const int slot = blockIdx.x;
    double sum = 0.0, mn = DBL_MAX, mx = -DBL_MAX;
    // each thread takes a contiguous range of blocks so the combine order is fixed
    const int per = (nblocks + blockDim.x - 1) / blockDim.x;
    const int b0 = threadIdx.x * per;
    const int b1 = min(nblocks, b0 + per);
    for (int b = b0; b < b1; ++b) {
        const double *q = partials + ((size_t)b * nslots + slot) * 3;
        sum = add(sum, q[0]);
        mn = fmin(mn, q[1]);
        mx = fmax(mx, q[2]);
    }
    sh[0][threadIdx.x] = sum;
    sh[1][threadIdx.x] = mn;
    sh[2][threadIdx.x] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        sum = 0.0, mn = DBL_MAX, mx = -DBL_MAX;
        for (int k = 0; k < (int)blockDim.x; ++k) {
            sum = add(sum, sh[0][k]);
            mn = fmin(mn, sh[1][k]);
            mx = fmax(mx, sh[2][k]);
        }
        out[slot * 3 + 0] = sum;
        out[slot * 3 + 1] = mn;
        out[slot * 3 + 2] = mx;
    }
}

// ---------------------------------------------------------------------------------------------
// op-list interpreter: one cell per thread, ops in the reference's order, every op reads and
// writes global memory (a thread always observes its own earlier writes), so aliased arrays
// behave exactly like the Fortran pointer aliases.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) oplist_kernel(const __grid_constant__ OpList L, const __grid_constant__ Consts c,
                                                     int64_t n)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        for (int o = 0; o < L.n; ++o) {
            const Op &op = L.ops[o];
            switch (op.code) {
                case OP_ZERO: op.out[j] = 0.0; break;
                case OP_COPY: op.out[j] = op.in[0][j]; break;
                case OP_QSUR_CCLM:
                    op.out[j] = spec_vapor_surface_cclm(op.in[0][j], op.in[1][j], op.in[2][j], c);
                    break;
                case OP_MEVA_CCLM:
                    op.out[j] = flux_mass_evap_cclm(op.in[0][j], op.in[1][j], op.in[2][j], op.in[3][j], op.in[4][j],
                                                    wind_speed(op.in[5][j], op.in[6][j]), c);
                    break;
                case OP_MEVA_RCO:
                    op.out[j] = flux_mass_evap_rco(op.in[0][j], op.in[1][j], wind_speed(op.in[2][j], op.in[3][j]));
                    break;
                case OP_ADD: op.out[j] = add(op.out[j], op.in[0][j]); break;
                case OP_SCALE: op.out[j] = mul(op.in[0][j], op.cst); break;
                case OP_HSEN_CCLM:
                    op.out[j] = flux_heat_sensible_cclm(op.in[0][j], op.in[1][j], op.in[2][j], op.in[3][j], op.in[4][j],
                                                        op.in[5][j], wind_speed(op.in[6][j], op.in[7][j]), c);
                    break;
                case OP_HSEN_RCO:
                    op.out[j] = flux_heat_sensible_rco(op.in[0][j], op.in[1][j], wind_speed(op.in[2][j], op.in[3][j]));
                    break;
                case OP_MOM_CCLM: {
                    const double u = op.in[4][j], v = op.in[5][j];
                    const double fa = momentum_flux_air_cclm(op.in[0][j], op.in[1][j], op.in[2][j], op.in[3][j],
                                                             wind_speed(u, v), c);
                    if (op.out) op.out[j] = momentum_component(fa, u);
                    if (op.out2) op.out2[j] = momentum_component(fa, v);
                    break;
                }
                case OP_MOM_RCO: {
                    const double u = op.in[0][j], v = op.in[1][j];
                    const double fa = momentum_flux_air_rco(wind_speed(u, v));
                    if (op.out) op.out[j] = momentum_component(fa, u);
                    if (op.out2) op.out2[j] = momentum_component(fa, v);
                    break;
                }
                case OP_RBBR: op.out[j] = flux_radiation_blackbody_StBo(op.in[0][j], op.cst); break;
                case OP_MULADD: op.out[j] = add(op.out[j], mul(op.in[0][j], op.in[1][j])); break;
                default: break;
            }
        }
    }
}

// corrections(1,12,n) Fortran order (month fastest) -> [12][n] month-major (bias_corrections.F90:29-30,191)
__global__ void transpose_corrections_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t n)
{
    __shared__ double tile[12][65];
    const int64_t j0 = (int64_t)blockIdx.x * 64;
    for (int e = threadIdx.x; e < 12 * 64; e += blockDim.x) {
        const int64_t g = j0 * 12 + e;   // contiguous read
        if (g < n * 12) tile[e % 12][e / 12] = src[g];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 12 * 64; e += blockDim.x) {
        const int m = e / 64, jj = e % 64;
        if (j0 + jj < n) dst[(int64_t)m * n + j0 + jj] = tile[m][jj];
    }
}

// do_regridding (basic.F90:476-486) with the COO elements grouped by destination (stable), one thread
// per destination cell, accumulation in the reference's element order: acc = acc + src*w
__global__ void regrid_csr_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ src_idx,
                                  const double *__restrict__ weight, const double *__restrict__ src,
                                  double *__restrict__ dst, int64_t n_dst)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_dst) return;
    double acc = 0.0;
    for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) acc = add(acc, mul(src[src_idx[k]], weight[k]));
    dst[r] = acc;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline int blocks_for(int64_t cells) { return (int)((cells + kFusedCellsPerBlock - 1) / kFusedCellsPerBlock); }

int fused_grid_blocks(const FusedPlan &p)
{
    const int nb_t = (p.do_early || p.do_normal) ? blocks_for(p.cells[0]) : 0;
    const int nb_u = p.do_normal ? blocks_for(p.cells[1]) : 0;
    const int nb_v = p.do_normal ? blocks_for(p.cells[2]) : 0;
    return nb_t + nb_u + nb_v;
}

static bool plan_aligned(const FusedPlan &p)
{
    auto ok = [](const void *q, int64_t cell0) { return q == nullptr || ((reinterpret_cast<uintptr_t>(q) + cell0 * 8) & 15) == 0; };
    bool a = true;
    const FusedT &t = p.t;
    const int64_t c0 = p.cell0[0];
    a = a && ok(t.rsdd, c0) && ok(t.bias, c0) && ok(t.area, c0) && ok(t.avg_qsur, c0) && ok(t.avg_meva, c0) &&
        ok(t.avg_hlat, c0) && ok(t.avg_hsen, c0) && ok(t.avg_rbbr, c0) && ok(t.avg_rsdr, c0);
    for (int i = 0; i < p.S; ++i) {
        const FusedTType &y = t.ty[i];
        const void *ps[] = {y.fice, y.psur, y.tsur, y.qatm, y.tatm, y.patm, y.uatm, y.vatm, y.a_evap, y.a_sens,
                            y.qsur_in, y.fare, y.qsur, y.meva, y.hlat, y.hsen, y.rbbr, y.rsdr};
        for (const void *q : ps) a = a && ok(q, c0);
    }
    for (int g = 0; g < 2; ++g) {
        const FusedUV &u = p.uv[g];
        const int64_t cg = p.cell0[g + 1];
        a = a && ok(u.area, cg) && ok(u.avg_qsur, cg) && ok(u.avg_mom, cg);
        for (int i = 0; i < p.S; ++i) {
            const FusedUVType &y = u.ty[i];
            const void *ps[] = {y.fice, y.psur, y.tsur, y.a_mom, y.uatm, y.vatm, y.qsur_in, y.fare, y.qsur, y.mom};
            for (const void *q : ps) a = a && ok(q, cg);
        }
    }
    return a;
}

template <int SS, bool DIAG, bool AL>
static cudaError_t launch_fused_t(const FusedPlan &p, int nb_t, int nb_u, int nb, cudaStream_t stream)
{
    const size_t smem = DIAG ? sizeof(double) * p.diag_n * kWarps * 3 : 0;   // <= 110*8*3*8 = 21 KB
    fused_step_kernel<SS, DIAG, AL><<<nb, kFusedThreads, smem, stream>>>(p, nb_t, nb_u);
    return cudaGetLastError();
}

template <int SS>
static cudaError_t launch_fused_s(const FusedPlan &p, int nb_t, int nb_u, int nb, bool al, cudaStream_t stream)
{
    if (p.diag)
        return al ? launch_fused_t<SS, true, true>(p, nb_t, nb_u, nb, stream)
                  : launch_fused_t<SS, true, false>(p, nb_t, nb_u, nb, stream);
    return al ? launch_fused_t<SS, false, true>(p, nb_t, nb_u, nb, stream)
              : launch_fused_t<SS, false, false>(p, nb_t, nb_u, nb, stream);
}

int launch_fused(const FusedPlan &p, cudaStream_t stream, int *launches)
{
    const int nb_t = (p.do_early || p.do_normal) ? blocks_for(p.cells[0]) : 0;
    const int nb_u = p.do_normal ? blocks_for(p.cells[1]) : 0;
    const int nb_v = p.do_normal ? blocks_for(p.cells[2]) : 0;
    const int nb = nb_t + nb_u + nb_v;
    if (nb == 0) return 0;
    const bool al = plan_aligned(p);
    cudaError_t e;
    switch (p.S) {
        case 1: e = launch_fused_s<1>(p, nb_t, nb_u, nb, al, stream); break;
        case 2: e = launch_fused_s<2>(p, nb_t, nb_u, nb, al, stream); break;
        default: e = launch_fused_s<0>(p, nb_t, nb_u, nb, al, stream); break;
    }
    if (launches) *launches += 1;
    return (int)e;
}

int launch_diag_finalize(const double *partials, int nblocks, int nslots, double *diag_out, cudaStream_t stream)
{
    if (nslots <= 0) return 0;
    diag_finalize_kernel<<<nslots, 256, 0, stream>>>(partials, nblocks, nslots, diag_out);
    return (int)cudaGetLastError();
}

int launch_oplist(const OpList &ops, const Consts &c, int64_t n, cudaStream_t stream)
{
    if (n <= 0 || ops.n == 0) return 0;
    int64_t nb = (n + 255) / 256;
    if (nb > 148 * 64) nb = 148 * 64;
    oplist_kernel<<<(int)nb, 256, 0, stream>>>(ops, c, n);
    return (int)cudaGetLastError();
}

int launch_transpose_corrections(const double *src, double *dst, int64_t n, cudaStream_t stream)
{
    if (n <= 0) return 0;
    transpose_corrections_kernel<<<(int)((n + 63) / 64), 256, 0, stream>>>(src, dst, n);
    return (int)cudaGetLastError();
}

int launch_regrid_csr(const int64_t *row_ptr, const int32_t *src_idx, const double *weight, const double *src,
                      double *dst, int64_t n_dst, cudaStream_t stream)
{
    if (n_dst <= 0) return 0;
    regrid_csr_kernel<<<(int)((n_dst + 255) / 256), 256, 0, stream>>>(row_ptr, src_idx, weight, src, dst, n_dst);
    return (int)cudaGetLastError();
}

}  // namespace fc
