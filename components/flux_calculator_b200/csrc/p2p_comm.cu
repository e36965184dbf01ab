// p2p_comm.cu -- multi-GPU exchange of the diagnostics vector over NVLink peer memory, fused into the step.
//
// One process per GPU.  Every rank allocates a mailbox (kMailDepth slots x nranks DiagMail records), exports it as a CUDA
// IPC handle; the host exchanges the handles (MPI_Allgather in the Fortran host, torch.distributed/gloo in bench.py)
// and every rank maps all mailboxes.  From then on the kernel that folds a step's diagnostics rows into its result
// vector -- the producer warps of the NEXT step's kernel, or the small fold kernel when the host asks first
// (spec_kernel.cu: diag_fold_slot; other paths: diag_post_kernel) -- stores that vector into the mailbox of every rank
// with plain peer stores of self-validating 8-byte words (32 data bits + the exchange's 32-bit sequence number; no
// fence, no flag store), one lane per (rank, plane).  No collective launch, nothing competing with the persistent
// kernel for an SM.  fc_get_diagnostics folds the nranks records in rank order on the host, so all ranks see
// bit-identical global sums.  NCCL (nccl_dyn.cu) remains as the fallback when IPC is unavailable.
//
// Records are keyed by EXCHANGE, not by step: fc_allreduce_diagnostics (a collective: every rank calls it for the same
// steps) numbers the exchanges, and the record of exchange e lives in slot e mod kMailDepth of every mailbox.  Steps
// whose global values nobody asked for post nothing.  Ranks that all read every exchange keep each other in lock step
// (a reader waits for every rank's record of e before it can start e + 1), so a slot is never reused under a reader;
// a rank that does not read may run at most kMailDepth - 1 exchanges ahead of one that does -- beyond that the late
// reader gets an error, never mixed data (every word carries its exchange's tag).
#include "context.h"

#include <string.h>
#include <unistd.h>

#include <chrono>
#include <vector>

using namespace fc;

extern "C" int fc_comm_p2p_handle(fc_context *c, char handle[FC_P2P_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) <= FC_P2P_HANDLE_BYTES, "IPC handle size");
    if (!c || !handle) return fail(c, FC_ERR_ARG, "fc_comm_p2p_handle: NULL argument");
    cudaSetDevice(c->device);
    if (!c->mailbox) {
        const size_t bytes = sizeof(DiagMail) * kMailDepth * kMaxPeers;
        CUDA_TRY(c, cudaMalloc(&c->mailbox, bytes));
        CUDA_TRY(c, cudaMemset(c->mailbox, 0, bytes));
    }
    cudaIpcMemHandle_t h;
    CUDA_TRY(c, cudaIpcGetMemHandle(&h, c->mailbox));
    memset(handle, 0, FC_P2P_HANDLE_BYTES);
    memcpy(handle, &h, sizeof h);
    return FC_OK;
}

extern "C" int fc_comm_p2p_connect(fc_context *c, const char *handles, int rank, int nranks)
{
    if (!c || !handles || nranks < 1 || nranks > kMaxPeers || rank < 0 || rank >= nranks)
        return fail(c, FC_ERR_ARG, "fc_comm_p2p_connect: bad argument (at most %d ranks)", kMaxPeers);
    if (!c->mailbox) return fail(c, FC_ERR_STATE, "fc_comm_p2p_connect: call fc_comm_p2p_handle first");
    cudaSetDevice(c->device);
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            c->peer_mail[r] = c->mailbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * FC_P2P_HANDLE_BYTES, sizeof h);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (q != rank && c->peer_mail[q]) {
                    cudaIpcCloseMemHandle(c->peer_mail[q]);
                    c->peer_mail[q] = nullptr;
                }
            return fail(c, FC_ERR_NCCL, "cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
        }
        c->peer_mail[r] = (DiagMail *)ptr;
    }
    c->rank = rank;
    c->nranks = nranks;
    c->p2p = nranks > 1;
    return FC_OK;
}

namespace fc {

void p2p_destroy(fc_context *c)
{
    for (int r = 0; r < kMaxPeers; ++r)
        if (c->peer_mail[r] && c->peer_mail[r] != c->mailbox) cudaIpcCloseMemHandle(c->peer_mail[r]);
    if (c->mailbox) cudaFree(c->mailbox);
    c->mailbox = nullptr;
    c->p2p = false;
}

// the PeerPost block of the next exchange
void p2p_make_post(fc_context *c, PeerPost &post)
{
    memset(&post, 0, sizeof post);
    if (!c->p2p) return;
    c->diag_seq += 1;
    post.nranks = c->nranks;
    post.rank = c->rank;
    post.seq = c->diag_seq;
    post.slot = (int)(c->diag_seq % (unsigned long long)kMailDepth);
    for (int r = 0; r < c->nranks; ++r) post.mail[r] = c->peer_mail[r];
}

// global diagnostics of the last step: wait until every needed word of every rank carries that step's tag, fold in rank order
int p2p_fetch(fc_context *c, double *planes /* [3][kDiagSlots] */, int n_active, int level)
{
    const int R = c->nranks, slot = (int)(c->diag_seq % (unsigned long long)kMailDepth);
    const unsigned int want = (unsigned int)c->diag_seq;
    std::vector<DiagMail> host((size_t)R);
    const DiagMail *src = c->mailbox + (size_t)slot * R;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));      // our own record is posted
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        CUDA_TRY(c, cudaMemcpy(host.data(), src, sizeof(DiagMail) * R, cudaMemcpyDeviceToHost));
        bool all = true;
        for (int r = 0; r < R && all; ++r)
            for (int pl = 0; pl < 3 && all; ++pl)
                for (int k = 0; k < n_active && all; ++k)
                    for (int h = 0; h < 2; ++h) {
                        const unsigned int tag = (unsigned int)(host[r].w[pl][k][h] >> 32);
                        if (tag == want) continue;
                        if ((int)(tag - want) > 0)
                            return fail(c, FC_ERR_STATE, "diagnostics of exchange %llu were overwritten: rank %d is already %d exchange(s) "
                                        "further (a rank that does not read may run at most %d exchanges ahead of one that does)",
                                        c->diag_seq, r, (int)(tag - want), kMailDepth - 1);
                        all = false;      // not arrived yet
                        break;
                    }
        if (all) break;
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0)
            return fail(c, FC_ERR_NCCL, "peer diagnostics of exchange %llu did not arrive within 30 s (every rank must call fc_allreduce_diagnostics for it and then issue another step, read or synchronize)", c->diag_seq);
        usleep(20);
    }
    // every word of this snapshot validated itself: fold
    auto val = [&](int r, int pl, int k) {
        const unsigned long long b = (host[r].w[pl][k][0] & 0xffffffffull) | (host[r].w[pl][k][1] << 32);
        double x;
        memcpy(&x, &b, 8);
        return x;
    };
    for (int k = 0; k < n_active; ++k) {
        double s = 0.0, mn = val(0, 1, k), mx = val(0, 2, k);
        for (int r = 0; r < R; ++r) {
            s += val(r, 0, k);
            if (level >= 2) {
                const double a = val(r, 1, k), b = val(r, 2, k);
                mn = a < mn ? a : mn;
                mx = b > mx ? b : mx;
            }
        }
        planes[0 * kDiagSlots + k] = s;
        planes[1 * kDiagSlots + k] = mn;
        planes[2 * kDiagSlots + k] = mx;
    }
    return FC_OK;
}

}  // namespace fc
