// nccl_dyn.cu -- NCCL binding for the global flux diagnostics (the path's only collective).
//
// libnccl.so.2 is dlopen'ed on first use: a Fortran host gets the system NCCL, a Python process that
// already imported torch gets torch's bundled copy (same SONAME -> same handle), and a single-GPU run
// never needs NCCL at all.  Per step the traffic is < 3 KB (sum / min / max vectors), so on NVSwitch
// only the ~10-20 us latency matters; it runs on a side stream, overlapped with the next step's kernel.
#include "context.h"

#include <dlfcn.h>
#include <nccl.h>

using namespace fc;

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &api()
{
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return a;
#define LOAD(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym))
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
}

}  // namespace

namespace fc {

void nccl_destroy(fc_context *c)
{
    if (c->nccl_comm && api().ok) api().CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
}

// The diagnostics of the step just issued live plane-major ([sum|min|max][kDiagSlots]) in diag_buf[cur]; each
// plane is reduced in place by one ncclAllReduce on a SIDE stream, so the main stream goes straight on to the
// next step's kernel.  The buffer is reused two steps later (finalize_diag waits for ev_comm then).
int nccl_allreduce_diag(fc_context *c)
{
    NcclApi &a = api();
    if (!a.ok) return fail(c, FC_ERR_NCCL, "libnccl.so.2 could not be loaded");
    cudaSetDevice(c->device);
    const int n = (int)c->diag_active.size();
    const int b = c->diag_cur;
    if (!c->comm_stream) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaEventRecord(c->ev_fin[b], c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->comm_stream, c->ev_fin[b], 0));
    double *buf = c->diag_buf[b];
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    ncclResult_t r = ncclSuccess;
    if (c->diag_level >= 2) {
        r = a.GroupStart();
        if (r == ncclSuccess) r = a.AllReduce(buf, buf, n, ncclDouble, ncclSum, comm, c->comm_stream);
        if (r == ncclSuccess) r = a.AllReduce(buf + kDiagSlots, buf + kDiagSlots, n, ncclDouble, ncclMin, comm, c->comm_stream);
        if (r == ncclSuccess) r = a.AllReduce(buf + 2 * kDiagSlots, buf + 2 * kDiagSlots, n, ncclDouble, ncclMax, comm, c->comm_stream);
        const ncclResult_t r2 = a.GroupEnd();
        if (r == ncclSuccess) r = r2;
    } else {
        r = a.AllReduce(buf, buf, n, ncclDouble, ncclSum, comm, c->comm_stream);
    }
    if (r != ncclSuccess) return fail(c, FC_ERR_NCCL, "ncclAllReduce failed: %s", a.GetErrorString(r));
    CUDA_TRY(c, cudaEventRecord(c->ev_comm[b], c->comm_stream));
    c->comm_busy[b] = true;
    c->launches += 1;
    return FC_OK;
}

}  // namespace fc

extern "C" int fc_comm_get_unique_id(char id[FC_UNIQUE_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) <= FC_UNIQUE_ID_BYTES, "unique id size");
    NcclApi &a = api();
    if (!a.ok) return fail(nullptr, FC_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
    ncclUniqueId u;
    ncclResult_t r = a.GetUniqueId(&u);
    if (r != ncclSuccess) return fail(nullptr, FC_ERR_NCCL, "ncclGetUniqueId failed: %s", a.GetErrorString(r));
    memset(id, 0, FC_UNIQUE_ID_BYTES);
    memcpy(id, &u, sizeof u);
    return FC_OK;
}

extern "C" int fc_comm_init(fc_context *c, const char id[FC_UNIQUE_ID_BYTES], int rank, int nranks)
{
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(c, FC_ERR_ARG, "fc_comm_init: bad argument");
    c->rank = rank;
    c->nranks = nranks;
    if (nranks == 1) return FC_OK;
    NcclApi &a = api();
    if (!a.ok) return fail(c, FC_ERR_NCCL, "libnccl.so.2 could not be loaded");
    cudaSetDevice(c->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm;
    ncclResult_t r = a.CommInitRank(&comm, nranks, u, rank);
    if (r != ncclSuccess) return fail(c, FC_ERR_NCCL, "ncclCommInitRank failed: %s", a.GetErrorString(r));
    c->nccl_comm = comm;
    return FC_OK;
}
