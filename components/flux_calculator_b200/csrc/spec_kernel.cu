// spec_kernel.cu -- the fast path of the fused coupling step: one persistent, formula-set-specialised kernel
// for single-surface-type canonical configurations (every configuration BASELINE.json benchmarks with S = 1).
//
//   * persistent CTAs (two per SM), each one producer warp + kSpecCW consumer warps;
//   * the producer streams the input arrays of 512-cell tiles into a shared-memory ring with cp.async.bulk,
//     completion counted on mbarriers; the ring is carved in UNITS: a t-grid tile (11-12 arrays) takes
//     t_units consecutive units, a u/v-grid tile (2-6 arrays) takes one, so the light u/v tiles get twice (bulk
//     set) or five times (RCO set) as many stages as the t tiles out of the same bytes, and the switch from t
//     tiles to u/v tiles needs no drain (the first use of a unit waits for the last t tile that covered it);
//   * consumers read operands from shared memory at the point of use (LDS.128 with an immediate offset: the
//     slot of every array is a compile-time constant of the formula set), two cells per thread in lock step,
//     results leave through 128-bit global stores; nothing but the running quantities lives in registers;
//   * the formula set is a template parameter -- no method dispatch inside the kernel;
//   * diagnostics: per-thread running sums (and min/max at level 2) over all tiles of a thread, one warp tree per
//     quantity per kernel, the CTA's warps combined in shared memory into one row per CTA, and the LAST CTA to
//     finish (atomic counter) folds all rows -- and those of the ragged-remainder launch, which runs first --
//     into the result vector: no follow-up kernel.  Static schedule + fixed trees: reproducible sums;
//   * cells whose operands leave the range in which the lock-step division / sqrt / exp / log sequences are
//     proven (vmath.cuh) are NOT handled inline: the warp notes the tile, and a cold, out-of-line epilogue
//     recomputes those tiles with the IEEE routines from global memory and rebuilds the warp's diagnostics
//     from the stored outputs.  The hot loop carries no call, no stack frame and no spill.
//
//   * the ragged remainder of a grid (cells mod 512) is one more tile of the schedule, taken by its CTA after
//     the ring tiles with guarded global loads/stores -- same chain code, no second launch;
//   * launched with programmatic stream serialisation: the next step's CTAs become resident and set up their
//     barriers while this step drains, then wait (griddepcontrol.wait) before touching global memory.
//
// Everything else (S > 1, averaging, 'zero'/'none' mixes, misaligned arrays, early-only phase) runs on the
// generic kernels of kernels.cu, instantiated from the same formula templates.
#include "plan.h"

#include <cuda_runtime.h>
#include <float.h>
#include <string.h>

namespace fc {

constexpr int kSpecV = 2;
constexpr int kSpecCW = kFusedThreads / 32;              // consumer warps: one 512-cell tile per pass
constexpr int kSpecConsumers = kSpecCW * 32;
constexpr int kSpecThreads = kSpecConsumers + 32;        // + producer warp
constexpr int kSpecTile = kSpecConsumers * kSpecV;
constexpr int kSpecSlotBytes = kSpecTile * 8;            // one array of one tile
constexpr int kSpecMaxUnits = 16;
constexpr int kSpecMaxSlots = 12;
constexpr int kSpecBadCap = 16;                          // flagged tiles remembered per warp and phase; more -> all
static_assert(kSpecTile == kFusedCellsPerBlock, "the ragged-remainder launch assumes the same tile size");

using S2 = Vd<kSpecV>;

// slots of the shared-memory stage, per formula set
namespace bulk {    // CCLM / MOM5 formulae (the MOM5 routines forward to the CCLM ones, flux_mass_evap.F90:107-115)
enum { FICE = 0, PSUR, TSUR, QATM, TATM, UATM, VATM, AEV, PATM, RSDD, BIAS, ASE /* MOM5: CHEA != CMOI */, NT };
enum { U_FICE = 0, U_PSUR, U_TSUR, U_UATM, U_VATM, U_AMOM, NUV };
constexpr int kUnitSlots = 6, kTUnits = 2;
}  // namespace bulk
namespace rco {     // Meier et al. 1999 formulae; QSUR on the t grid is still the CCLM routine (App. F-1)
enum { FICE = 0, PSUR, TSUR, QATM, TATM, UATM, VATM, RSDD, BIAS, NT };
enum { U_UATM = 0, U_VATM, NUV };
constexpr int kUnitSlots = 2, kTUnits = 5;
}  // namespace rco
static_assert(bulk::NT <= bulk::kUnitSlots * bulk::kTUnits && bulk::NUV <= bulk::kUnitSlots, "");
static_assert(rco::NT <= rco::kUnitSlots * rco::kTUnits && rco::NUV <= rco::kUnitSlots, "");
static_assert(bulk::NT <= kSpecMaxSlots && rco::NT <= kSpecMaxSlots, "");

enum SpecSet { SET_BULK = 0, SET_RCO = 1 };

struct SpecPlan {
    Consts c;
    int64_t first[3];                 // first cell of the range on each grid
    int64_t end[3];                   // one past its last cell
    int ntiles[3];                    // ceil(cells / tile): the last tile of a grid may be partial (guarded path)
    int units;                        // ring size in units (a multiple of t_units)
    int unit_bytes;
    int t_units;
    int do_early;                     // RBBR in this launch
    int has_bias, has_rsdr;
    int ase_slot;                     // bulk: slot of the sensible-heat transfer coefficient (AEV for CCLM, ASE for MOM5)
    int diag;
    double latent_heat;
    const double *src[3][kSpecMaxSlots];   // per grid: source array of each stage slot (null: slot unused)
    uint32_t tx_bytes[3];             // bytes one tile of that grid brings in
    const double *area[3];
    double *outq[DQ_COUNT];           // output array per diagnostics quantity (null: not produced)
    double *partials;                 // [plane][compact slot][row]; rows [0, grid) are this kernel's (one per CTA)
    int64_t rows, plane;
    double *diag_out;                 // [sum|min|max][kDiagSlots]
    unsigned int *counter;            // CTAs done
    PeerPost post;                    // peer mailboxes (multi-GPU): the last CTA also posts the result there
    signed char dmap[DQ_COUNT];       // quantity -> compact diagnostics slot (-1: inactive)
};

struct WarpSums {    // per-CTA staging of the consumer warps' diagnostics
    double v[3][DQ_COUNT][kSpecCW];
};

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred P_OUT;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P_OUT, [%1], %2;\n\t"
                 "selp.b32 %0, 1, 0, P_OUT;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// operand sources, result sinks, diagnostics accumulators
// ---------------------------------------------------------------------------------------------
// hot path: this thread's 16 bytes of each slot of the current stage
struct LdStage {
    const char *base;
    __device__ __forceinline__ S2 operator()(int slot) const
    {
        const double2 t = *reinterpret_cast<const double2 *>(base + slot * kSpecSlotBytes);
        S2 r;
        r.v[0] = t.x;
        r.v[1] = t.y;
        return r;
    }
};
struct StPair {
    int64_t j;
    __device__ __forceinline__ void operator()(double *p, const S2 &x) const
    {
        *reinterpret_cast<double2 *>(p + j) = make_double2(x.v[0], x.v[1]);
    }
};
// partial tile: guarded global accesses, nv = 0, 1 or 2 valid cells
struct LdPairGuard {
    const double *const *src;
    int64_t j;
    int nv;
    __device__ __forceinline__ S2 operator()(int slot) const
    {
        S2 r;
#pragma unroll
        for (int k = 0; k < kSpecV; ++k) r.v[k] = (k < nv) ? __ldg(src[slot] + j + k) : 1.0;
        return r;
    }
};
struct StPairGuard {
    int64_t j;
    int nv;
    __device__ __forceinline__ void operator()(double *p, const S2 &x) const
    {
#pragma unroll
        for (int k = 0; k < kSpecV; ++k)
            if (k < nv) p[j + k] = x.v[k];
    }
};
// cold path: one cell from global memory
struct LdCell {
    const double *const *src;
    int64_t j;
    __device__ __forceinline__ Vd<1> operator()(int slot) const
    {
        Vd<1> r;
        r.v[0] = src[slot][j];
        return r;
    }
};
struct StCell {
    int64_t j;
    __device__ __forceinline__ void operator()(double *p, const Vd<1> &x) const { p[j] = x.v[0]; }
};
struct NoDiag {
    template <class T>
    __device__ __forceinline__ void operator()(int, const T &) const {}
};

// acc = acc + (area0*x0 + area1*x1): the one definition both the hot loop and the rebuild of the cold epilogue use
__device__ __forceinline__ void diag_pair(double &s, double &mn, double &mx, int level, double a0, double a1, double x0, double x1)
{
    s = add(s, add(mul(a0, x0), mul(a1, x1)));
    if (level >= 2) {
        mn = fmin(fmin(mn, x0), x1);
        mx = fmax(fmax(mx, x0), x1);
    }
}

// nv valid cells (partial tile); identical to diag_pair for nv == 2
__device__ __forceinline__ void diag_cells(double &s, double &mn, double &mx, int level, double a0, double a1, double x0, double x1, int nv)
{
    if (nv >= 2) {
        diag_pair(s, mn, mx, level, a0, a1, x0, x1);
    } else if (nv == 1) {
        s = add(s, mul(a0, x0));
        if (level >= 2) {
            mn = fmin(mn, x0);
            mx = fmax(mx, x0);
        }
    }
}

template <int DIAG, int NQ, int Q0>
struct DiagAcc {    // running diagnostics of the NQ quantities Q0.. of one phase
    double s[NQ], mn[DIAG >= 2 ? NQ : 1], mx[DIAG >= 2 ? NQ : 1];
    S2 area;
    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            s[q] = 0.0;
            if (DIAG >= 2) {
                mn[q] = DBL_MAX;
                mx[q] = -DBL_MAX;
            }
        }
    }
    __device__ __forceinline__ void operator()(int q, const S2 &x)
    {
        if (DIAG == 0) return;
        const int k = q - Q0;
        if constexpr (DIAG >= 2) {
            diag_pair(s[k], mn[k], mx[k], DIAG, area.v[0], area.v[1], x.v[0], x.v[1]);
        } else {
            double dm = 0.0, dM = 0.0;
            diag_pair(s[k], dm, dM, DIAG, area.v[0], area.v[1], x.v[0], x.v[1]);
        }
    }
};

template <int DIAG, int NQ, int Q0>
struct DiagAccGuard {    // the same accumulators fed from a partial tile
    DiagAcc<DIAG, NQ, Q0> &a;
    int nv;
    __device__ __forceinline__ void operator()(int q, const S2 &x)
    {
        if (DIAG == 0) return;
        const int k = q - Q0;
        if constexpr (DIAG >= 2) {
            diag_cells(a.s[k], a.mn[k], a.mx[k], DIAG, a.area.v[0], a.area.v[1], x.v[0], x.v[1], nv);
        } else {
            double dm = 0.0, dM = 0.0;
            diag_cells(a.s[k], dm, dM, DIAG, a.area.v[0], a.area.v[1], x.v[0], x.v[1], nv);
        }
    }
};

// warp tree of one quantity; lane 0 leaves the warp's value in the CTA's staging area
template <int DIAG>
__device__ __forceinline__ void diag_flush_one(WarpSums &ws, int q, double s, double mn, double mx)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s = add(s, __shfl_down_sync(0xffffffffu, s, off));
        if (DIAG >= 2) {
            mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off));
            mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        ws.v[0][q][w] = s;
        if (DIAG >= 2) {
            ws.v[1][q][w] = mn;
            ws.v[2][q][w] = mx;
        }
    }
}

__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kSpecConsumers) : "memory"); }

// end of kernel, consumer warps only: warps -> one row per CTA -> (last CTA) rows of all CTAs + remainder rows -> result
template <int DIAG>
__device__ __forceinline__ void diag_finish(const SpecPlan &p, WarpSums &ws, int *is_last)
{
    const int tid = threadIdx.x, G = gridDim.x;
    consumer_barrier();
    if (tid < DQ_COUNT) {
        const int cs = p.dmap[tid];
        if (cs >= 0) {
            double s = 0.0, mn = DBL_MAX, mx = -DBL_MAX;
#pragma unroll
            for (int w = 0; w < kSpecCW; ++w) {
                s = add(s, ws.v[0][tid][w]);
                if (DIAG >= 2) {
                    mn = fmin(mn, ws.v[1][tid][w]);
                    mx = fmax(mx, ws.v[2][tid][w]);
                }
            }
            double *o = p.partials + (int64_t)cs * p.rows + blockIdx.x;
            o[0] = s;
            if (DIAG >= 2) {
                o[p.plane] = mn;
                o[2 * p.plane] = mx;
            }
        }
        __threadfence();
    }
    consumer_barrier();
    if (tid == 0) *is_last = (atomicAdd(p.counter, 1u) == (unsigned)(G - 1));
    consumer_barrier();
    if (!*is_last) return;
    __threadfence();
    const int warp = tid >> 5, lane = tid & 31;
    for (int q = warp; q < DQ_COUNT; q += kSpecCW) {      // one warp per quantity, fixed order: lane-strided rows, then a tree
        const int cs = p.dmap[q];
        if (cs < 0) continue;
        const double *col = p.partials + (int64_t)cs * p.rows;
        double s = 0.0, mn = DBL_MAX, mx = -DBL_MAX;
        constexpr int kPer = 10;      // rows per lane fetched at once (2 CTAs x 148 SMs = 296 rows -> one round)
        for (int r0 = 0; r0 < G; r0 += 32 * kPer) {
            double vs[kPer], vn[kPer], vx[kPer];
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int r = r0 + k * 32 + lane;
                vs[k] = (r < G) ? __ldcg(col + r) : 0.0;
                if (DIAG >= 2) {
                    vn[k] = (r < G) ? __ldcg(col + p.plane + r) : DBL_MAX;
                    vx[k] = (r < G) ? __ldcg(col + 2 * p.plane + r) : -DBL_MAX;
                }
            }
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                s = add(s, vs[k]);
                if (DIAG >= 2) {
                    mn = fmin(mn, vn[k]);
                    mx = fmax(mx, vx[k]);
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s = add(s, __shfl_down_sync(0xffffffffu, s, off));
            if (DIAG >= 2) {
                mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off));
                mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
            }
        }
        if (lane == 0) {
            p.diag_out[0 * kDiagSlots + cs] = s;
            p.diag_out[1 * kDiagSlots + cs] = mn;
            p.diag_out[2 * kDiagSlots + cs] = mx;
            // compute -> exchange in one kernel: the result goes straight into every rank's mailbox over NVLink
            for (int r = 0; r < p.post.nranks; ++r) {
                DiagMail *m = p.post.mail[r] + (size_t)p.post.parity * p.post.nranks + p.post.rank;
                m->v[0][cs] = s;
                m->v[1][cs] = mn;
                m->v[2][cs] = mx;
            }
        }
    }
    if (p.post.nranks > 1) {
        __threadfence_system();          // the records are visible system-wide before ...
        consumer_barrier();
        if (tid == 0)
            for (int r = 0; r < p.post.nranks; ++r) {      // ... the sequence number that publishes them
                DiagMail *m = p.post.mail[r] + (size_t)p.post.parity * p.post.nranks + p.post.rank;
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&m->seq), "l"(p.post.seq) : "memory");
            }
    }
    if (tid == 0) *p.counter = 0u;      // ready for the next launch (same stream)
}

// ---------------------------------------------------------------------------------------------
// the chains, written once over (arithmetic policy, operand source, result sink, diagnostics sink)
// ---------------------------------------------------------------------------------------------
// t grid: QSUR -> MEVA (+bias) -> HLAT, HSEN, RBBR, RSDR.  Call-site wiring of calculate.F90 (App. A.2): the evaporation
// routine gets TATM in its temperature slot (:87), the sensible-heat routine gets QATM in its q_s slot (:178).
template <int SET, class M, class LD, class ST, class DG>
__device__ __forceinline__ void spec_t_chain(M &m, const SpecPlan &p, const LD &ld, const ST &st, DG &dg)
{
    using T = typename M::T;
    const Consts &c = p.c;
    constexpr int FICE = SET == SET_BULK ? (int)bulk::FICE : (int)rco::FICE, PSUR = SET == SET_BULK ? (int)bulk::PSUR : (int)rco::PSUR,
                  TSUR = SET == SET_BULK ? (int)bulk::TSUR : (int)rco::TSUR, QATM = SET == SET_BULK ? (int)bulk::QATM : (int)rco::QATM,
                  TATM = SET == SET_BULK ? (int)bulk::TATM : (int)rco::TATM, UATM = SET == SET_BULK ? (int)bulk::UATM : (int)rco::UATM,
                  VATM = SET == SET_BULK ? (int)bulk::VATM : (int)rco::VATM, RSDD = SET == SET_BULK ? (int)bulk::RSDD : (int)rco::RSDD,
                  BIAS = SET == SET_BULK ? (int)bulk::BIAS : (int)rco::BIAS;
    // calc_spec_vapor_surface (calculate.F90:25-50)
    const T qsur = spec_vapor_surface_cclm(m, ld(FICE), ld(PSUR), ld(TSUR), c);
    st(p.outq[DQ_QSUR_T], qsur);
    dg(DQ_QSUR_T, qsur);
    const T vel = wind_speed(m, ld(UATM), ld(VATM));
    // calc_flux_mass_evap (calculate.F90:54-120) + bias (:112-116)
    T meva;
    if (SET == SET_BULK) meva = flux_mass_evap_cclm(m, ld(bulk::AEV), ld(PSUR), ld(QATM), qsur, ld(TATM), vel, c);
    else meva = flux_mass_evap_rco(m, ld(QATM), ld(TSUR), vel);
    if (p.has_bias) meva = M::add(meva, ld(BIAS));
    st(p.outq[DQ_MEVA], meva);
    dg(DQ_MEVA, meva);
    // calc_flux_heat_latent (calculate.F90:124-154): the corrected MEVA
    const T hlat = M::mul(meva, M::bc(p.latent_heat));
    st(p.outq[DQ_HLAT], hlat);
    dg(DQ_HLAT, hlat);
    // calc_flux_heat_sensible (calculate.F90:156-208)
    T hsen;
    if (SET == SET_BULK) hsen = flux_heat_sensible_cclm(m, ld(p.ase_slot), ld(bulk::PATM), ld(PSUR), ld(QATM), ld(TATM), ld(TSUR), vel, c);
    else hsen = flux_heat_sensible_rco<M>(ld(TATM), ld(TSUR), vel);
    st(p.outq[DQ_HSEN], hsen);
    dg(DQ_HSEN, hsen);
    // calc_flux_radiation_blackbody (calculate.F90:320-345), early phase
    if (p.do_early) {
        const T rbbr = flux_radiation_blackbody_StBo<M>(ld(TSUR), c.stefan_boltzmann_constant);
        st(p.outq[DQ_RBBR], rbbr);
        dg(DQ_RBBR, rbbr);
    }
    // distribute_shortwave_radiation_flux (calculate.F90:347-364): a copy
    if (p.has_rsdr) {
        const T rsdr = ld(RSDD);
        st(p.outq[DQ_RSDR], rsdr);
        dg(DQ_RSDR, rsdr);
    }
}

// u / v grid: QSUR on that grid (bulk sets) -> momentum flux, east component on the u grid, north on the v grid
template <int SET, class M, class LD, class ST, class DG>
__device__ __forceinline__ void spec_uv_chain(M &m, const SpecPlan &p, int north, const LD &ld, const ST &st, DG &dg)
{
    using T = typename M::T;
    const Consts &c = p.c;
    const int qQ = north ? DQ_QSUR_V : DQ_QSUR_U, qM = north ? DQ_VMOM : DQ_UMOM;
    if (SET == SET_BULK) {
        const T qsur = spec_vapor_surface_cclm(m, ld(bulk::U_FICE), ld(bulk::U_PSUR), ld(bulk::U_TSUR), c);
        st(p.outq[qQ], qsur);
        dg(DQ_QSUR_U, qsur);      // phase-local index: DiagAcc of the u/v phase starts at DQ_QSUR_U for both grids
        const T vel = wind_speed(m, ld(bulk::U_UATM), ld(bulk::U_VATM));
        const T fa = momentum_flux_air_cclm(m, ld(bulk::U_AMOM), ld(bulk::U_PSUR), qsur, ld(bulk::U_TSUR), vel, c);
        const T mom = momentum_component<M>(fa, ld(north ? (int)bulk::U_VATM : (int)bulk::U_UATM));
        st(p.outq[qM], mom);
        dg(DQ_UMOM, mom);
    } else {
        const T vel = wind_speed(m, ld(rco::U_UATM), ld(rco::U_VATM));
        const T fa = momentum_flux_air_rco<M>(vel);
        const T mom = momentum_component<M>(fa, ld(north ? (int)rco::U_VATM : (int)rco::U_UATM));
        st(p.outq[qM], mom);
        dg(DQ_UMOM, mom);
    }
}

// ---------------------------------------------------------------------------------------------
// cold epilogue of one phase of one warp: recompute the flagged tiles with the IEEE routines (scalar, from global
// memory), then rebuild this warp's diagnostics of the phase from the stored outputs in the hot loop's order.
// ---------------------------------------------------------------------------------------------
__device__ unsigned long long g_spec_exact_calls = 0ull;

unsigned long long read_spec_exact_calls()
{
    unsigned long long v = 0;
    cudaMemcpyFromSymbol(&v, g_spec_exact_calls, sizeof v);
    return v;
}

template <int SET, int DIAG>
__device__ __noinline__ void spec_cold_phase(const SpecPlan &p, int ph, int64_t j0, int64_t jstride, int ntiles, const int *flagged,
                                             int nflag, WarpSums &ws)
{
    const bool all = nflag > kSpecBadCap;
    for (int i = 0; i < ntiles; ++i) {
        bool f = all;
        for (int e = 0; e < nflag && e < kSpecBadCap; ++e) f = f || (flagged[e] == i);
        if (!f) continue;
        atomicAdd(&g_spec_exact_calls, 1ull);
        for (int k = 0; k < kSpecV; ++k) {
            const int64_t j = j0 + (int64_t)i * jstride + k;
            if (j >= p.end[ph]) break;      // partial tile
            ExactVec<1> m;
            const LdCell ld{p.src[ph], j};
            const StCell st{j};
            NoDiag nd;
            if (ph == 0) spec_t_chain<SET>(m, p, ld, st, nd);
            else spec_uv_chain<SET>(m, p, ph - 1, ld, st, nd);
        }
    }
    if (DIAG == 0) return;
    const int q0 = (ph == 0) ? DQ_QSUR_T : (ph == 1 ? DQ_QSUR_U : DQ_QSUR_V);
    const int nq = (ph == 0) ? 6 : 2;
    for (int q = q0; q < q0 + nq; ++q) {
        const double *x = p.outq[q];
        if (p.dmap[q] < 0 || x == nullptr) continue;      // uniform
        double s = 0.0, mn = DBL_MAX, mx = -DBL_MAX;
        // the hot loop's order: the partial tile (if this CTA has it: always its last) first, then the ring tiles
        const int64_t last_start = j0 - (int64_t)threadIdx.x * kSpecV + (int64_t)(ntiles - 1) * jstride;
        const bool has_partial = ntiles > 0 && last_start + kSpecTile > p.end[ph];
        for (int ii = 0; ii < ntiles; ++ii) {
            const int i = has_partial ? (ii == 0 ? ntiles - 1 : ii - 1) : ii;
            const int64_t j = j0 + (int64_t)i * jstride;
            const int64_t left = p.end[ph] - j;
            const int nv = left >= 2 ? 2 : (left > 0 ? (int)left : 0);
            if (nv == 2) diag_pair(s, mn, mx, DIAG, p.area[ph][j], p.area[ph][j + 1], x[j], x[j + 1]);
            else if (nv == 1) diag_cells(s, mn, mx, DIAG, p.area[ph][j], 0.0, x[j], 0.0, 1);
        }
        diag_flush_one<DIAG>(ws, q, s, mn, mx);
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int SET, int DIAG>
__global__ void __launch_bounds__(kSpecThreads, 2) flux_spec_kernel(const __grid_constant__ SpecPlan p)
{
    extern __shared__ __align__(128) char ring[];
    __shared__ uint64_t fullT[kSpecMaxUnits], emptyT[kSpecMaxUnits], fullU[kSpecMaxUnits], emptyU[kSpecMaxUnits];
    __shared__ int flagged[kSpecCW][kSpecBadCap];
    __shared__ int nflagged[kSpecCW];
    __shared__ WarpSums ws;
    __shared__ int is_last;

    const int NU = p.units, A = p.t_units, GT = NU / A;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NU; ++s) {
            mbar_init(&fullT[s], 1);             // one expect_tx arrival + the bytes
            mbar_init(&emptyT[s], kSpecCW);      // one arrival per consumer warp
            mbar_init(&fullU[s], 1);
            mbar_init(&emptyU[s], kSpecCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < kSpecCW) nflagged[threadIdx.x] = 0;
    if (DIAG)
        for (int e = threadIdx.x; e < DQ_COUNT * kSpecCW; e += blockDim.x) {
            (&ws.v[0][0][0])[e] = 0.0;
            (&ws.v[1][0][0])[e] = DBL_MAX;
            (&ws.v[2][0][0])[e] = -DBL_MAX;
        }
    __syncthreads();
    // programmatic dependent launch: everything above overlapped the previous kernel of the stream; from here on
    // global memory is touched, so wait for that kernel to complete (no-op without the launch attribute)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // static schedule: this CTA takes positions b, b+G, b+2G, ... of the tile list [t tiles | u tiles | v tiles]
    const int G = gridDim.x, b = blockIdx.x;
    // (scalars, not arrays: a runtime-indexed local array would live in local memory)
    int cnt0, cnt1, cnt2;
    int64_t tl0, tl1, tl2;      // first tile of this CTA inside each grid
    {
        auto sched = [&](int64_t off, int64_t n, int &cnt, int64_t &tl) {
            int64_t x0 = b;
            if (off > b) x0 = b + ((off - b + G - 1) / G) * G;
            cnt = (x0 < off + n) ? (int)((off + n - x0 + G - 1) / G) : 0;
            tl = x0 - off;
        };
        sched(0, p.ntiles[0], cnt0, tl0);
        sched(p.ntiles[0], p.ntiles[1], cnt1, tl1);
        sched((int64_t)p.ntiles[0] + p.ntiles[1], p.ntiles[2], cnt2, tl2);
    }
    // the last tile of a grid may be partial: its CTA takes it after the ring tiles, through guarded global accesses
    auto partial = [&](int ph, int cnt, int64_t tl) {
        return cnt > 0 && tl + (int64_t)(cnt - 1) * G == p.ntiles[ph] - 1 && (p.end[ph] - p.first[ph]) % kSpecTile != 0;
    };
    const int part0 = partial(0, cnt0, tl0), part1 = partial(1, cnt1, tl1), part2 = partial(2, cnt2, tl2);
    const int ring0 = cnt0 - part0, ring1 = cnt1 - part1, ring2 = cnt2 - part2;      // tiles that travel through the ring
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == kSpecCW) {
        // ---------------- producer warp: one lane issues the bulk copies ----------------
        if (lane != 0) return;
        {   // t tiles: group g = i mod GT covers units [g*A, g*A + A)
            int g = 0, use = 0;
            for (int i = 0; i < ring0; ++i) {
                if (use > 0) mbar_wait(&emptyT[g], (use - 1) & 1);
                const int64_t cell = p.first[0] + (tl0 + (int64_t)i * G) * kSpecTile;
                mbar_expect_tx(&fullT[g], p.tx_bytes[0]);
                char *dst = ring + (size_t)g * A * p.unit_bytes;
#pragma unroll 1
                for (int a = 0; a < kSpecMaxSlots; ++a)
                    if (p.src[0][a]) bulk_g2s(dst + a * kSpecSlotBytes, p.src[0][a] + cell, kSpecSlotBytes, &fullT[g]);
                if (++g == GT) {
                    g = 0;
                    ++use;
                }
            }
        }
        {   // u then v tiles: one unit each, k counts across both grids
            int h = 0, use = 0;
#pragma unroll 1
            for (int ph = 1; ph < 3; ++ph) {
                const int cnt_ph = ph == 1 ? ring1 : ring2;
                const int64_t tl_ph = ph == 1 ? tl1 : tl2;
                for (int i = 0; i < cnt_ph; ++i) {
                    if (use > 0) {
                        mbar_wait(&emptyU[h], (use - 1) & 1);
                    } else {      // first use of this unit: the last t tile of the group that covered it must be done
                        const int g = h / A;
                        const int uses_t = (ring0 > g) ? (ring0 - g + GT - 1) / GT : 0;
                        if (uses_t > 0) mbar_wait(&emptyT[g], (uses_t - 1) & 1);
                    }
                    const int64_t cell = p.first[ph] + (tl_ph + (int64_t)i * G) * kSpecTile;
                    mbar_expect_tx(&fullU[h], p.tx_bytes[ph]);
                    char *dst = ring + (size_t)h * p.unit_bytes;
#pragma unroll 1
                    for (int a = 0; a < (SET == SET_BULK ? (int)bulk::NUV : (int)rco::NUV); ++a)
                        if (p.src[ph][a]) bulk_g2s(dst + a * kSpecSlotBytes, p.src[ph][a] + cell, kSpecSlotBytes, &fullU[h]);
                    if (++h == NU) {
                        h = 0;
                        ++use;
                    }
                }
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int toff = threadIdx.x * (kSpecV * 8);
    {   // t phase
        DiagAcc<DIAG, 6, DQ_QSUR_T> dg;
        dg.reset();
        const int64_t j0 = p.first[0] + tl0 * kSpecTile + threadIdx.x * kSpecV, jstride = (int64_t)G * kSpecTile;
        if (part0) {      // the partial tile first: its (slow, guarded) global loads overlap the filling of the ring
            const int64_t j = j0 + (int64_t)(cnt0 - 1) * jstride;
            const int64_t left = p.end[0] - j;
            const int nv = left >= 2 ? 2 : (left > 0 ? (int)left : 0);
            if (DIAG) {
                dg.area.v[0] = nv > 0 ? __ldg(p.area[0] + j) : 0.0;
                dg.area.v[1] = nv > 1 ? __ldg(p.area[0] + j + 1) : 0.0;
            }
            FastVec<kSpecV> m;
            const LdPairGuard ld{p.src[0], j, nv};
            const StPairGuard st{j, nv};
            DiagAccGuard<DIAG, 6, DQ_QSUR_T> dgg{dg, nv};
            spec_t_chain<SET>(m, p, ld, st, dgg);
            if (__any_sync(0xffffffffu, nv > 0 && m.bad()) && lane == 0) {
                const int n = nflagged[warp];
                if (n < kSpecBadCap) flagged[warp][n] = cnt0 - 1;
                nflagged[warp] = n + 1;
            }
        }
        int g = 0, use = 0;
        int64_t j = j0;
        for (int i = 0; i < ring0; ++i, j += jstride) {
            if (DIAG) {
                const double2 a = __ldg(reinterpret_cast<const double2 *>(p.area[0] + j));
                dg.area.v[0] = a.x;
                dg.area.v[1] = a.y;
            }
            mbar_wait(&fullT[g], use & 1);
            FastVec<kSpecV> m;
            const LdStage ld{ring + (size_t)g * A * p.unit_bytes + toff};
            const StPair st{j};
            spec_t_chain<SET>(m, p, ld, st, dg);
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyT[g]);
            if (__any_sync(0xffffffffu, m.bad()) && lane == 0) {
                const int n = nflagged[warp];
                if (n < kSpecBadCap) flagged[warp][n] = i;
                nflagged[warp] = n + 1;
            }
            if (++g == GT) {
                g = 0;
                ++use;
            }
        }
        __syncwarp();
        const int nf = nflagged[warp];
        if (nf) {
            spec_cold_phase<SET, DIAG>(p, 0, j0, jstride, cnt0, flagged[warp], nf, ws);
            __syncwarp();
            if (lane == 0) nflagged[warp] = 0;
        } else if (DIAG) {
#pragma unroll
            for (int q = 0; q < 6; ++q) diag_flush_one<DIAG>(ws, DQ_QSUR_T + q, dg.s[q], DIAG >= 2 ? dg.mn[q] : 0.0, DIAG >= 2 ? dg.mx[q] : 0.0);
        }
    }
    {   // u phase, then v phase: same code, same ring
        int h = 0, use = 0;
#pragma unroll 1
        for (int ph = 1; ph < 3; ++ph) {
            DiagAcc<DIAG, 2, DQ_QSUR_U> dg;
            dg.reset();
            const int north = ph - 1;
            const int cnt_ph = ph == 1 ? cnt1 : cnt2, ring_ph = ph == 1 ? ring1 : ring2;
            const int64_t j0 = p.first[ph] + (ph == 1 ? tl1 : tl2) * kSpecTile + threadIdx.x * kSpecV, jstride = (int64_t)G * kSpecTile;
            if (ring_ph != cnt_ph) {      // partial tile of this grid, before its ring tiles
                const int64_t j = j0 + (int64_t)(cnt_ph - 1) * jstride;
                const int64_t left = p.end[ph] - j;
                const int nv = left >= 2 ? 2 : (left > 0 ? (int)left : 0);
                if (DIAG) {
                    dg.area.v[0] = nv > 0 ? __ldg(p.area[ph] + j) : 0.0;
                    dg.area.v[1] = nv > 1 ? __ldg(p.area[ph] + j + 1) : 0.0;
                }
                FastVec<kSpecV> m;
                const LdPairGuard ld{p.src[ph], j, nv};
                const StPairGuard st{j, nv};
                DiagAccGuard<DIAG, 2, DQ_QSUR_U> dgg{dg, nv};
                spec_uv_chain<SET>(m, p, north, ld, st, dgg);
                if (__any_sync(0xffffffffu, nv > 0 && m.bad()) && lane == 0) {
                    const int n = nflagged[warp];
                    if (n < kSpecBadCap) flagged[warp][n] = cnt_ph - 1;
                    nflagged[warp] = n + 1;
                }
            }
            int64_t j = j0;
            for (int i = 0; i < ring_ph; ++i, j += jstride) {
                if (DIAG) {
                    const double2 a = __ldg(reinterpret_cast<const double2 *>(p.area[ph] + j));
                    dg.area.v[0] = a.x;
                    dg.area.v[1] = a.y;
                }
                mbar_wait(&fullU[h], use & 1);
                FastVec<kSpecV> m;
                const LdStage ld{ring + (size_t)h * p.unit_bytes + toff};
                const StPair st{j};
                spec_uv_chain<SET>(m, p, north, ld, st, dg);
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyU[h]);
                if (__any_sync(0xffffffffu, m.bad()) && lane == 0) {
                    const int n = nflagged[warp];
                    if (n < kSpecBadCap) flagged[warp][n] = i;
                    nflagged[warp] = n + 1;
                }
                if (++h == NU) {
                    h = 0;
                    ++use;
                }
            }
            __syncwarp();
            const int nf = nflagged[warp];
            if (nf) {
                spec_cold_phase<SET, DIAG>(p, ph, j0, jstride, cnt_ph, flagged[warp], nf, ws);
                __syncwarp();
                if (lane == 0) nflagged[warp] = 0;
                __syncwarp();
            } else if (DIAG) {
                const int q0 = north ? DQ_QSUR_V : DQ_QSUR_U;
                diag_flush_one<DIAG>(ws, q0, dg.s[0], DIAG >= 2 ? dg.mn[0] : 0.0, DIAG >= 2 ? dg.mx[0] : 0.0);
                diag_flush_one<DIAG>(ws, q0 + 1, dg.s[1], DIAG >= 2 ? dg.mn[1] : 0.0, DIAG >= 2 ? dg.mx[1] : 0.0);
            }
        }
    }
    if (DIAG) diag_finish<DIAG>(p, ws, &is_last);
}

// ---------------------------------------------------------------------------------------------
// host side: does the fused plan fit the specialised kernel?  If so, build its plan and launch.
// ---------------------------------------------------------------------------------------------
static int spec_num_sms()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

static bool is_bulk(int m) { return m == M_CCLM || m == M_MOM5; }

// cells[g] cells starting at first[g] (every bound array 16-byte aligned there: decided by the caller)
static bool spec_build(const FusedPlan &p, const int64_t first[3], const int64_t cells[3], SpecPlan &sp, int *set_out)
{
    if (p.S != 1 || !p.do_normal) return false;
    const FusedTType &T = p.t.ty[0];
    const FusedUVType &U = p.uv[0].ty[0], &W = p.uv[1].ty[0];
    const FusedT &t = p.t;
    if (t.avg_qsur || t.avg_meva || t.avg_hlat || t.avg_hsen || t.avg_rbbr || t.avg_rsdr) return false;
    if (p.uv[0].avg_qsur || p.uv[0].avg_mom || p.uv[1].avg_qsur || p.uv[1].avg_mom) return false;
    if (T.m_qsur != M_CCLM || T.qsur_in || U.qsur_in || W.qsur_in) return false;
    if (T.m_hlat != M_WATER && T.m_hlat != M_ICE) return false;
    if (p.do_early ? (T.m_rbbr != M_STBO) : (T.m_rbbr != M_NONE)) return false;
    int set;
    if (is_bulk(T.m_meva) && is_bulk(T.m_hsen) && U.m_qsur == M_CCLM && W.m_qsur == M_CCLM && is_bulk(U.m_mom) && is_bulk(W.m_mom))
        set = SET_BULK;
    else if (T.m_meva == M_RCO && T.m_hsen == M_RCO && U.m_qsur == M_NONE && W.m_qsur == M_NONE && U.m_mom == M_RCO && W.m_mom == M_RCO)
        set = SET_RCO;
    else
        return false;
    if (p.diag && !(t.area && p.uv[0].area && p.uv[1].area)) return false;

    memset(&sp, 0, sizeof sp);
    sp.c = p.c;
    for (int g = 0; g < 3; ++g) {
        sp.first[g] = first[g];
        sp.end[g] = first[g] + cells[g];
        sp.ntiles[g] = (int)((cells[g] + kSpecTile - 1) / kSpecTile);
    }
    sp.do_early = p.do_early;
    sp.has_bias = t.bias != nullptr;
    sp.has_rsdr = t.rsdd != nullptr;
    sp.latent_heat = T.latent_heat;
    sp.diag = p.diag;
    if (set == SET_BULK) {
        const double **s = sp.src[0];
        s[bulk::FICE] = T.fice; s[bulk::PSUR] = T.psur; s[bulk::TSUR] = T.tsur; s[bulk::QATM] = T.qatm; s[bulk::TATM] = T.tatm;
        s[bulk::UATM] = T.uatm; s[bulk::VATM] = T.vatm; s[bulk::AEV] = T.a_evap; s[bulk::PATM] = T.patm;
        s[bulk::RSDD] = t.rsdd; s[bulk::BIAS] = t.bias;
        if (T.a_sens == T.a_evap) {
            sp.ase_slot = bulk::AEV;
        } else {
            s[bulk::ASE] = T.a_sens;
            sp.ase_slot = bulk::ASE;
        }
        for (int k = 0; k <= bulk::PATM; ++k)
            if (!s[k]) return false;
        for (int g = 0; g < 2; ++g) {
            const FusedUVType &Y = p.uv[g].ty[0];
            const double **u = sp.src[g + 1];
            u[bulk::U_FICE] = Y.fice; u[bulk::U_PSUR] = Y.psur; u[bulk::U_TSUR] = Y.tsur; u[bulk::U_UATM] = Y.uatm;
            u[bulk::U_VATM] = Y.vatm; u[bulk::U_AMOM] = Y.a_mom;
            for (int k = 0; k < bulk::NUV; ++k)
                if (!u[k]) return false;
        }
        sp.t_units = bulk::kTUnits;
        sp.unit_bytes = bulk::kUnitSlots * kSpecSlotBytes;
    } else {
        const double **s = sp.src[0];
        s[rco::FICE] = T.fice; s[rco::PSUR] = T.psur; s[rco::TSUR] = T.tsur; s[rco::QATM] = T.qatm; s[rco::TATM] = T.tatm;
        s[rco::UATM] = T.uatm; s[rco::VATM] = T.vatm; s[rco::RSDD] = t.rsdd; s[rco::BIAS] = t.bias;
        for (int k = 0; k <= rco::VATM; ++k)
            if (!s[k]) return false;
        for (int g = 0; g < 2; ++g) {
            const FusedUVType &Y = p.uv[g].ty[0];
            sp.src[g + 1][rco::U_UATM] = Y.uatm;
            sp.src[g + 1][rco::U_VATM] = Y.vatm;
            if (!Y.uatm || !Y.vatm) return false;
        }
        sp.t_units = rco::kTUnits;
        sp.unit_bytes = rco::kUnitSlots * kSpecSlotBytes;
    }
    for (int g = 0; g < 3; ++g) {
        int n = 0;
        for (int k = 0; k < kSpecMaxSlots; ++k) n += sp.src[g][k] != nullptr;
        sp.tx_bytes[g] = (uint32_t)n * kSpecSlotBytes;
    }
    sp.area[0] = t.area;
    sp.area[1] = p.uv[0].area;
    sp.area[2] = p.uv[1].area;
    sp.outq[DQ_QSUR_T] = T.qsur; sp.outq[DQ_MEVA] = T.meva; sp.outq[DQ_HLAT] = T.hlat; sp.outq[DQ_HSEN] = T.hsen;
    sp.outq[DQ_RBBR] = p.do_early ? T.rbbr : nullptr;
    sp.outq[DQ_RSDR] = sp.has_rsdr ? T.rsdr : nullptr;
    sp.outq[DQ_QSUR_U] = U.qsur; sp.outq[DQ_UMOM] = U.mom; sp.outq[DQ_QSUR_V] = W.qsur; sp.outq[DQ_VMOM] = W.mom;
    if (!T.qsur || !T.meva || !T.hlat || !T.hsen || !U.mom || !W.mom) return false;
    if (set == SET_BULK && (!U.qsur || !W.qsur)) return false;
    if (p.do_early && !T.rbbr) return false;
    if (sp.has_rsdr && !T.rsdr) return false;
    // ring: as many units as fit next to a second CTA on the SM (227 KB - 1 KB reserved per CTA - static), a multiple of t_units
    const int budget = 112 * 1024;
    int units = budget / sp.unit_bytes;
    if (units > kSpecMaxUnits) units = kSpecMaxUnits;
    units -= units % sp.t_units;
    if (units < 2 * sp.t_units) return false;
    sp.units = units;
    for (int q = 0; q < DQ_COUNT; ++q) sp.dmap[q] = p.diag ? p.diag_map[DQ_COUNT + q] : (signed char)-1;    // surface type 1
    sp.partials = p.diag_partials;
    sp.rows = p.diag_rows;
    sp.plane = (int64_t)p.diag_n * p.diag_rows;
    sp.diag_out = p.diag_out;
    sp.counter = p.diag_counter;
    sp.post = p.post;
    if (!p.diag || sp.post.nranks <= 1) sp.post.nranks = 0;
    *set_out = set;
    return true;
}

static int spec_grid(const SpecPlan &sp)
{
    const int64_t total = (int64_t)sp.ntiles[0] + sp.ntiles[1] + sp.ntiles[2];
    const int cap = 2 * spec_num_sms();
    return (int)(total < cap ? total : cap);
}

// > 0: the plan fits the specialised kernel, value = its grid size (= diagnostics rows); 0: it does not
int spec_applicable(const FusedPlan &p, const int64_t first[3], const int64_t cells[3])
{
    SpecPlan sp;
    int set = 0;
    if (cells[0] + cells[1] + cells[2] <= 0 || !spec_build(p, first, cells, sp, &set)) return 0;
    return spec_grid(sp);
}

template <int SET, int DIAG>
static cudaError_t spec_launch_t(const SpecPlan &sp, int grid, cudaStream_t stream)
{
    const size_t smem = (size_t)sp.units * sp.unit_bytes;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(flux_spec_kernel<SET, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(flux_spec_kernel<SET, DIAG>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    // programmatic stream serialisation: this launch may become resident while the previous kernel of the stream
    // drains; the kernel itself waits (griddepcontrol.wait) before its first global access
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kSpecThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, flux_spec_kernel<SET, DIAG>, sp);
}

int spec_launch(const FusedPlan &p, const int64_t first[3], const int64_t cells[3], cudaStream_t stream)
{
    SpecPlan sp;
    int set = 0;
    if (!spec_build(p, first, cells, sp, &set)) return (int)cudaErrorInvalidValue;
    if (sp.diag && (!sp.partials || !sp.diag_out || !sp.counter)) return (int)cudaErrorInvalidValue;
    const int grid = spec_grid(sp);
    cudaError_t e;
    if (set == SET_BULK) {
        if (sp.diag >= 2) e = spec_launch_t<SET_BULK, 2>(sp, grid, stream);
        else if (sp.diag == 1) e = spec_launch_t<SET_BULK, 1>(sp, grid, stream);
        else e = spec_launch_t<SET_BULK, 0>(sp, grid, stream);
    } else {
        if (sp.diag >= 2) e = spec_launch_t<SET_RCO, 2>(sp, grid, stream);
        else if (sp.diag == 1) e = spec_launch_t<SET_RCO, 1>(sp, grid, stream);
        else e = spec_launch_t<SET_RCO, 0>(sp, grid, stream);
    }
    return (int)e;
}

}  // namespace fc
