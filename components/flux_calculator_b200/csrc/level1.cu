// level1.cu -- Level 1 of the C ABI: the 14 public routines of module flux_library
// (flux_lib/flux_library.F90:32-45) in array form.  Each call is one op of the interpreter kernel.
// Host arrays are staged through temporary device buffers; device arrays are used in place and the
// call is asynchronous on `stream`.
#include "context.h"

#include <vector>

using namespace fc;

namespace {

struct Staged {
    const double *user;
    double *dev;
    bool temp;
    bool is_out;
};

int run_level1(int code, double cst, const Consts &consts, double *out, double *out2,
               std::initializer_list<const double *> ins, int64_t n, fc_stream_t stream_)
{
    if (n < 0) return fail(nullptr, FC_ERR_ARG, "flux_lib array routine: n < 0");
    if (n == 0) return FC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    std::vector<Staged> st;
    int dev = -1;
    auto stage = [&](const double *p, bool is_out) -> int {
        bool is_dev = false, is_pin = false;
        int d = -1;
        classify_pointer(p, &is_dev, &is_pin, &d);
        if (is_dev) {
            if (dev >= 0 && d != dev) return fail(nullptr, FC_ERR_ARG, "flux_lib array routine: arrays live on different devices");
            if (dev < 0) {
                dev = d;
                CUDA_TRY(nullptr, cudaSetDevice(dev));
            }
            st.push_back({p, const_cast<double *>(p), false, is_out});
        } else {
            st.push_back({p, nullptr, true, is_out});
        }
        return FC_OK;
    };
    for (const double *p : ins) {
        if (!p) return fail(nullptr, FC_ERR_ARG, "flux_lib array routine: NULL input array");
        if (int rc = stage(p, false)) return rc;
    }
    const size_t n_in = st.size();
    for (double *p : {out, out2})
        if (p)
            if (int rc = stage(p, true)) return rc;
    if (st.size() == n_in) return fail(nullptr, FC_ERR_ARG, "flux_lib array routine: no result array");
    bool any_temp = false;
    int rc = FC_OK;
    for (Staged &s : st)
        if (s.temp) {
            any_temp = true;
            cudaError_t e = cudaMalloc(&s.dev, (size_t)n * sizeof(double));
            if (e == cudaSuccess && !s.is_out) e = cudaMemcpyAsync(s.dev, s.user, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) {
                rc = fail(nullptr, FC_ERR_CUDA, "flux_lib array routine: %s (no CPU fallback exists)", cudaGetErrorString(e));
                break;
            }
        }
    if (rc == FC_OK) {
        OpList L;
        L.n = 1;
        L.pad = 0;
        Op &o = L.ops[0];
        memset(&o, 0, sizeof o);
        o.code = code;
        o.cst = cst;
        size_t k = 0;
        for (; k < n_in; ++k) o.in[k] = st[k].dev;
        if (out) o.out = st[k++].dev;
        if (out2) o.out2 = st[k++].dev;
        if (launch_oplist(L, consts, n, stream)) rc = fail(nullptr, FC_ERR_CUDA, "flux_lib array routine: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == FC_OK)
        for (Staged &s : st)
            if (s.temp && s.is_out) {
                cudaError_t e = cudaMemcpyAsync(const_cast<double *>(s.user), s.dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream);
                if (e != cudaSuccess) rc = fail(nullptr, FC_ERR_CUDA, "flux_lib array routine: %s", cudaGetErrorString(e));
            }
    if (any_temp) {
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess && rc == FC_OK) rc = fail(nullptr, FC_ERR_CUDA, "flux_lib array routine: %s", cudaGetErrorString(e));
        for (Staged &s : st)
            if (s.temp && s.dev) cudaFree(s.dev);
    }
    return rc;
}

inline double opt(const double *p, double dflt) { return p ? *p : dflt; }   // IF (PRESENT(x)) ... ELSE default

}  // namespace

extern "C" {

int fc_spec_vapor_surface_cclm(double *q_s, const double *f_ice, const double *p_s, const double *T_s, int64_t n,
                               const double *R_d_new, const double *R_v_new, fc_stream_t stream)
{
    const Consts d = make_consts();
    const Consts c = make_consts(d.heat_capacity_air, d.latent_heat_vaporization, d.latent_heat_sublimation,
                                 opt(R_d_new, d.gas_constant_air), opt(R_v_new, d.gas_constant_vapor));
    return run_level1(OP_QSUR_CCLM, 0.0, c, q_s, nullptr, {f_ice, p_s, T_s}, n, stream);
}

int fc_flux_mass_evap_cclm(double *evap, const double *a_moisture, const double *p_s, const double *q_a,
                           const double *q_s, const double *T_s, const double *u_a, const double *v_a, int64_t n,
                           const double *u_min_new, const double *R_d_new, const double *R_v_new, fc_stream_t stream)
{
    const Consts d = make_consts();
    const Consts c = make_consts(d.heat_capacity_air, d.latent_heat_vaporization, d.latent_heat_sublimation,
                                 opt(R_d_new, d.gas_constant_air), opt(R_v_new, d.gas_constant_vapor),
                                 d.stefan_boltzmann_constant, opt(u_min_new, d.u_min_evap));
    return run_level1(OP_MEVA_CCLM, 0.0, c, evap, nullptr, {a_moisture, p_s, q_a, q_s, T_s, u_a, v_a}, n, stream);
}

int fc_flux_mass_evap_mom5(double *evap, const double *a_moisture, const double *p_s, const double *q_a,
                           const double *q_s, const double *T_s, const double *u_a, const double *v_a, int64_t n,
                           fc_stream_t stream)
{
    return fc_flux_mass_evap_cclm(evap, a_moisture, p_s, q_a, q_s, T_s, u_a, v_a, n, nullptr, nullptr, nullptr, stream);
}

int fc_flux_mass_evap_rco(double *evap, const double *q_a, const double *T_s, const double *u_a, const double *v_a,
                          int64_t n, fc_stream_t stream)
{
    return run_level1(OP_MEVA_RCO, 0.0, make_consts(), evap, nullptr, {q_a, T_s, u_a, v_a}, n, stream);
}

int fc_flux_heat_latent_ice(double *hlat, const double *evap, int64_t n, const double *L_s_new, fc_stream_t stream)
{
    const Consts d = make_consts();
    return run_level1(OP_SCALE, opt(L_s_new, d.latent_heat_sublimation), d, hlat, nullptr, {evap}, n, stream);
}

int fc_flux_heat_latent_water(double *hlat, const double *evap, int64_t n, const double *L_v_new, fc_stream_t stream)
{
    const Consts d = make_consts();
    return run_level1(OP_SCALE, opt(L_v_new, d.latent_heat_vaporization), d, hlat, nullptr, {evap}, n, stream);
}

int fc_flux_heat_sensible_cclm(double *hsen, const double *a_moisture, const double *p_a, const double *p_s,
                               const double *q_s, const double *T_a, const double *T_s, const double *u_a,
                               const double *v_a, int64_t n, const double *c_p_new, const double *u_min_new,
                               const double *R_d_new, const double *R_v_new, fc_stream_t stream)
{
    const Consts d = make_consts();
    const Consts c = make_consts(opt(c_p_new, d.heat_capacity_air), d.latent_heat_vaporization, d.latent_heat_sublimation,
                                 opt(R_d_new, d.gas_constant_air), opt(R_v_new, d.gas_constant_vapor),
                                 d.stefan_boltzmann_constant, opt(u_min_new, d.u_min_evap));
    return run_level1(OP_HSEN_CCLM, 0.0, c, hsen, nullptr, {a_moisture, p_a, p_s, q_s, T_a, T_s, u_a, v_a}, n, stream);
}

int fc_flux_heat_sensible_mom5(double *hsen, const double *a_moisture, const double *p_a, const double *p_s,
                               const double *q_s, const double *T_a, const double *T_s, const double *u_a,
                               const double *v_a, int64_t n, fc_stream_t stream)
{
    return fc_flux_heat_sensible_cclm(hsen, a_moisture, p_a, p_s, q_s, T_a, T_s, u_a, v_a, n, nullptr, nullptr, nullptr,
                                      nullptr, stream);
}

int fc_flux_heat_sensible_rco(double *hsen, const double *T_a, const double *T_s, const double *u_a, const double *v_a,
                              int64_t n, fc_stream_t stream)
{
    return run_level1(OP_HSEN_RCO, 0.0, make_consts(), hsen, nullptr, {T_a, T_s, u_a, v_a}, n, stream);
}

int fc_flux_momentum_cclm(double *tau_e, double *tau_n, const double *a_momentum, const double *p_s, const double *q_s,
                          const double *T_s, const double *u_a, const double *v_a, int64_t n, const double *R_d_new,
                          const double *R_v_new, fc_stream_t stream)
{
    const Consts d = make_consts();
    const Consts c = make_consts(d.heat_capacity_air, d.latent_heat_vaporization, d.latent_heat_sublimation,
                                 opt(R_d_new, d.gas_constant_air), opt(R_v_new, d.gas_constant_vapor));
    return run_level1(OP_MOM_CCLM, 0.0, c, tau_e, tau_n, {a_momentum, p_s, q_s, T_s, u_a, v_a}, n, stream);
}

int fc_flux_momentum_mom5(double *tau_e, double *tau_n, const double *a_momentum, const double *p_s, const double *q_s,
                          const double *T_s, const double *u_a, const double *v_a, int64_t n, fc_stream_t stream)
{
    return fc_flux_momentum_cclm(tau_e, tau_n, a_momentum, p_s, q_s, T_s, u_a, v_a, n, nullptr, nullptr, stream);
}

int fc_flux_momentum_rco(double *tau_e, double *tau_n, const double *u_a, const double *v_a, int64_t n,
                         fc_stream_t stream)
{
    return run_level1(OP_MOM_RCO, 0.0, make_consts(), tau_e, tau_n, {u_a, v_a}, n, stream);
}

int fc_flux_radiation_blackbody_StBo(double *rbbr, const double *T_s, int64_t n, const double *sigma_new,
                                     fc_stream_t stream)
{
    const Consts d = make_consts();
    return run_level1(OP_RBBR, opt(sigma_new, d.stefan_boltzmann_constant), d, rbbr, nullptr, {T_s}, n, stream);
}

int fc_distribute_radiation_flux(double *out, const double *flux_avg, const double *albedo_avg, const double *albedo_type,
                                 int64_t n, fc_stream_t stream)
{
    (void)albedo_avg;    // distribute_radiation_flux.F90:24: the albedo factors are commented out
    (void)albedo_type;
    return run_level1(OP_COPY, 0.0, make_consts(), out, nullptr, {flux_avg}, n, stream);
}

}  // extern "C"
