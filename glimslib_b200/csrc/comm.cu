// Multi-GPU plumbing: ghost-vertex halo exchange (grouped ncclSend/ncclRecv) and the Krylov
// dot-product allreduce.  One context per rank; nothing here runs on a single GPU.
#include "common.h"
#include <nccl.h>

#define GL_NCCL(call)                                                                  \
    do {                                                                               \
        ncclResult_t r_ = (call);                                                      \
        if (r_ != ncclSuccess) {                                                       \
            char b_[256];                                                              \
            snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
            throw GlError(GLIMS_ERR_NCCL, b_);                                         \
        }                                                                              \
    } while (0)

namespace {
__global__ void k_pack(const double* __restrict__ x, const int* __restrict__ idx, i64 n, int bs, double* buf) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n * bs) return;
    i64 v = t / bs; int k = (int)(t - v * bs);
    buf[t] = x[(i64)idx[v] * bs + k];
}
}  // namespace

void halo_exchange(glims_ctx* c, double* xb, int bs) {
    Halo& h = c->halo;
    if (!h.active || !h.comm) return;
    ncclComm_t comm = (ncclComm_t)h.comm;
    if (h.n_send > 0) {
        i64 n = h.n_send * bs;
        k_pack<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(xb, h.send_idx, h.n_send, bs, h.send_buf);
        c->launches++;
    }
    GL_NCCL(ncclGroupStart());
    for (size_t p = 0; p < h.peers.size(); ++p) {
        i64 ns = h.send_ptr[p + 1] - h.send_ptr[p], nr = h.recv_ptr[p + 1] - h.recv_ptr[p];
        if (ns > 0) GL_NCCL(ncclSend(h.send_buf + h.send_ptr[p] * bs, ns * bs, ncclDouble, h.peers[p], comm, c->stream));
        if (nr > 0) GL_NCCL(ncclRecv(xb + (h.n_owned + h.recv_ptr[p]) * bs, nr * bs, ncclDouble, h.peers[p], comm, c->stream));
    }
    GL_NCCL(ncclGroupEnd());
}

namespace {
__global__ void k_pack32(const float* __restrict__ x, const int* __restrict__ idx, i64 n, int bs, float* buf) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n * bs) return;
    i64 v = t / bs; int k = (int)(t - v * bs);
    buf[t] = x[(i64)idx[v] * bs + k];
}
}  // namespace

void halo_exchange_f32(glims_ctx* c, float* xb, int bs) {
    Halo& h = c->halo;
    if (!h.active || !h.comm) return;
    ncclComm_t comm = (ncclComm_t)h.comm;
    float* sb = (float*)h.send_buf;
    if (h.n_send > 0) {
        i64 n = h.n_send * bs;
        k_pack32<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(xb, h.send_idx, h.n_send, bs, sb);
        c->launches++;
    }
    GL_NCCL(ncclGroupStart());
    for (size_t p = 0; p < h.peers.size(); ++p) {
        i64 ns = h.send_ptr[p + 1] - h.send_ptr[p], nr = h.recv_ptr[p + 1] - h.recv_ptr[p];
        if (ns > 0) GL_NCCL(ncclSend(sb + h.send_ptr[p] * bs, ns * bs, ncclFloat, h.peers[p], comm, c->stream));
        if (nr > 0) GL_NCCL(ncclRecv(xb + (h.n_owned + h.recv_ptr[p]) * bs, nr * bs, ncclFloat, h.peers[p], comm, c->stream));
    }
    GL_NCCL(ncclGroupEnd());
}

void allreduce_scalars(glims_ctx* c, int slot0, int n) {
    Halo& h = c->halo;
    if (!h.active || !h.comm) return;
    GL_NCCL(ncclAllReduce(c->scal + slot0, c->scal + slot0, n, ncclDouble, ncclSum, (ncclComm_t)h.comm, c->stream));
}

extern "C" {

int glims_nccl_unique_id(void* id128) {
    if (!id128) return GLIMS_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    return ncclGetUniqueId((ncclUniqueId*)id128) == ncclSuccess ? GLIMS_OK : GLIMS_ERR_NCCL;
}

int glims_comm_init(glims_ctx* c, int32_t n_ranks, int32_t rank, const void* id128) {
    if (!c || !id128) return GLIMS_ERR_ARG;
    try {
        GL_CUDA(cudaSetDevice(c->device));
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        ncclComm_t comm;
        GL_NCCL(ncclCommInitRank(&comm, n_ranks, id, rank));
        c->halo.comm = comm; c->halo.n_ranks = n_ranks; c->halo.rank = rank;
    } catch (const GlError& e) { c->err = e.msg; return e.code; }
    return GLIMS_OK;
}

int glims_set_halo(glims_ctx* c, int32_t n_peers, const int32_t* peers, const int64_t* send_ptr,
                   const int32_t* send_idx, const int64_t* recv_ptr) {
    if (!c || n_peers < 0) return GLIMS_ERR_ARG;
    try {
        GL_CUDA(cudaSetDevice(c->device));
        Halo& h = c->halo;
        h.peers.assign(peers, peers + n_peers);
        h.send_ptr.assign(send_ptr, send_ptr + n_peers + 1);
        h.recv_ptr.assign(recv_ptr, recv_ptr + n_peers + 1);
        h.n_send = n_peers ? send_ptr[n_peers] : 0;
        if (h.send_idx) cudaFree(h.send_idx);
        if (h.send_buf) cudaFree(h.send_buf);
        GL_CUDA(cudaMalloc(&h.send_idx, sizeof(int) * (h.n_send > 0 ? h.n_send : 1)));
        GL_CUDA(cudaMalloc(&h.send_buf, sizeof(double) * (h.n_send > 0 ? h.n_send : 1) * 8));
        if (h.n_send > 0) GL_CUDA(cudaMemcpy(h.send_idx, send_idx, sizeof(int) * h.n_send, cudaMemcpyHostToDevice));
        if (h.n_owned + (n_peers ? recv_ptr[n_peers] : 0) != c->n_v)
            throw GlError(GLIMS_ERR_ARG, "set_halo: owned + ghosts != local vertices");
    } catch (const GlError& e) { c->err = e.msg; return e.code; }
    return GLIMS_OK;
}

}  // extern "C"
