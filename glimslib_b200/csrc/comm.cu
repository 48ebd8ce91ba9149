// Multi-GPU data path: ghost-vertex halo exchange before every operator application and the allreduce of the
// Krylov scalars.  One context per rank; nothing here runs on a single GPU.
//
// Two transports:
//  * peer memory over NVLink / NVSwitch (default): every rank owns a "window" that all ranks of the box map through
//    CUDA IPC.  A halo exchange is one kernel per rank (k_halo): gather the owned boundary values, store them
//    straight into the neighbours' windows, publish a sequence number, wait for the neighbours' sequence numbers
//    and copy the received values into the vector's ghost block.  The allreduce is one single-CTA
//    kernel: all-to-all stores of the partial sums, sequence flags, sum in rank order (bitwise identical on every
//    rank).  No host involvement, no proxy thread, graph-capturable; latency is a few NVLink round trips instead
//    of an NCCL kernel + rendezvous per call.
//  * NCCL grouped send/recv + ncclAllReduce: bootstrap (IPC handle exchange) and fallback (GLIMS_NO_P2P=1, or
//    peers that cannot map each other's memory).
//
// Protocol.  Every rank executes the same sequence of exchanges, so the k-th exchange has the same sequence number
// everywhere.  Receive buffers are double-buffered by the parity of the sequence number: a neighbour can start
// exchange k+1 before I have consumed exchange k (other buffer), but it cannot start k+2 before it has consumed my
// k+1, which I only send after consuming k (stream order) -- so a buffer is never overwritten while it is being read.
// Waits are bounded (about thirty seconds of SM clock); a timeout raises a device flag that glims_step turns into
// GLIMS_ERR_NCCL instead of hanging the GPU.
#include "common.h"
#include <nccl.h>
#include <cstdlib>
#include <cstring>

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

constexpr int P2P_MAX_RANKS = 8;
constexpr int RED_MAX = 64;                 // scalars per allreduce (S_GM0 batches use up to 32)
constexpr long long SPIN_LIMIT = 60000000000LL;  // clock64 ticks (~30 s: ranks may be seconds apart during setup)

// device-visible description of the windows (lives in device memory, one copy per rank)
struct P2PDev {
    int n_ranks, rank, n_peers;
    int peer_rank[P2P_MAX_RANKS];                 // halo neighbours
    long long send_ptr[P2P_MAX_RANKS + 1];        // my send list is grouped by neighbour
    long long dst_off[P2P_MAX_RANKS];             // first ghost slot (in vertices) of my values in that neighbour's ghost block
    long long recv_cnt[P2P_MAX_RANKS];            // ghosts I receive from that neighbour (0: nothing to wait for)
    unsigned char* win[P2P_MAX_RANKS];            // window base of every rank (own one included), indexed by rank
    long long halo_bytes;                         // bytes of ONE parity buffer of a window's halo area -- per rank:
    long long halo_bytes_of[P2P_MAX_RANKS];       //   ... of every rank's window (sizes differ with the ghost count)
};
// window layout (byte offsets, identical on every rank except for the size of the halo area)
constexpr size_t OFF_FLAG_HALO = 0;                                   // u64 [P2P_MAX_RANKS]
constexpr size_t OFF_FLAG_RED = 64;                                   // u64 [P2P_MAX_RANKS]
constexpr size_t OFF_SEQ = 128;                                       // u64 seq_halo, seq_red ; u32 pushed ; i32 error ; u32 left
constexpr size_t OFF_RED = 256;                                       // double [2][P2P_MAX_RANKS][RED_MAX]
constexpr size_t OFF_HALO = OFF_RED + 2 * P2P_MAX_RANKS * RED_MAX * 8;  // 2 parity buffers of halo_bytes each

struct P2P {
    bool on = false;
    unsigned char* win = nullptr;
    size_t win_bytes = 0;
    std::vector<void*> opened;        // peer mappings to close
    P2PDev* dev = nullptr;            // device copy
    P2PDev host;
};

__device__ inline unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ inline void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ inline bool spin_until(const unsigned long long* flag, unsigned long long seq, int* err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < seq) {
        if (clock64() - t0 > SPIN_LIMIT) { *err = 1; return false; }
        __nanosleep(64);
    }
    return true;
}

// One kernel per halo exchange.  Every CTA (1) gathers its share of my boundary values and stores them straight into
// the neighbours' windows, (2) the CTA that finishes last publishes the sequence number to the neighbours, (3) every CTA
// waits for the neighbours' sequence numbers and copies its share of the received values into the ghost block of x.
// A rank's flag is published by the LAST of its CTAs to have pushed, and every CTA waits for the neighbours' flags: the
// grid must therefore be co-resident (the launcher caps it with the occupancy API; the stream runs alone on the GPU).
template <typename T>
__global__ void __launch_bounds__(256)
k_halo(const P2PDev* __restrict__ P, T* __restrict__ x, const int* __restrict__ idx, long long n_send, int bs,
       long long n_owned, long long n_ghost) {
    unsigned char* me = P->win[P->rank];
    unsigned long long* seqp = (unsigned long long*)(me + OFF_SEQ);
    int* err = (int*)(me + OFF_SEQ + 20);
    const unsigned long long seq = *seqp + 1;            // incremented by the last CTA on its way out; all CTAs of this
    const int par = (int)(seq & 1);                      // grid read the same value (see the exit protocol below)
    const long long n = n_send * bs;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long v = t / bs;
        const int k = (int)(t - v * bs);
        int p = 0;
        while (p + 1 < P->n_peers && v >= P->send_ptr[p + 1]) ++p;
        const int r = P->peer_rank[p];
        T* dst = (T*)(P->win[r] + OFF_HALO + (size_t)par * P->halo_bytes_of[r]);
        dst[(P->dst_off[p] + (v - P->send_ptr[p])) * bs + k] = x[(long long)idx[v] * bs + k];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    unsigned* done = (unsigned*)(me + OFF_SEQ + 16);
    if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < P->n_peers)
            st_release_sys((unsigned long long*)(P->win[P->peer_rank[threadIdx.x]] + OFF_FLAG_HALO) + P->rank, seq);
    }
    if (threadIdx.x < P->n_peers)
        spin_until((const unsigned long long*)(me + OFF_FLAG_HALO) + P->peer_rank[threadIdx.x], seq, err);
    __syncthreads();
    const T* src = (const T*)(me + OFF_HALO + (size_t)par * P->halo_bytes);
    T* dst = x + n_owned * bs;
    const long long ng = n_ghost * bs;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ng; t += (long long)gridDim.x * blockDim.x)
        dst[t] = __ldcv(&src[t]);
    // exit protocol: the sequence number advances only after every CTA has read it
    __syncthreads();
    unsigned* gone = (unsigned*)(me + OFF_SEQ + 24);     // separate counter: a fast CTA may leave before a slow one has pushed
    if (threadIdx.x == 0 && atomicAdd(gone, 1u) == gridDim.x - 1) { *done = 0; *gone = 0; __threadfence(); *seqp = seq; }
}

// scal[slot0 .. slot0+n) <- sum over ranks, in rank order on every rank
__global__ void __launch_bounds__(256)
k_allreduce(const P2PDev* __restrict__ P, double* __restrict__ scal, int slot0, int n) {
    unsigned char* me = P->win[P->rank];
    unsigned long long* seqp = (unsigned long long*)(me + OFF_SEQ + 8);
    int* err = (int*)(me + OFF_SEQ + 20);
    const unsigned long long seq = *seqp + 1;
    const int par = (int)(seq & 1), R = P->n_ranks;
    for (int t = threadIdx.x; t < R * n; t += blockDim.x) {
        const int r = t / n, k = t - r * n;
        double* dst = (double*)(P->win[r] + OFF_RED) + ((size_t)par * P2P_MAX_RANKS + P->rank) * RED_MAX + k;
        *dst = scal[slot0 + k];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < R) st_release_sys((unsigned long long*)(P->win[threadIdx.x] + OFF_FLAG_RED) + P->rank, seq);
    if (threadIdx.x < R) spin_until((const unsigned long long*)(me + OFF_FLAG_RED) + threadIdx.x, seq, err);
    __syncthreads();
    if (threadIdx.x < n) {
        const double* src = (const double*)(me + OFF_RED) + (size_t)par * P2P_MAX_RANKS * RED_MAX + threadIdx.x;
        double s = 0.0;
        for (int r = 0; r < R; ++r) s += __ldcv(&src[(size_t)r * RED_MAX]);
        scal[slot0 + threadIdx.x] = s;
    }
    if (threadIdx.x == 0) *seqp = seq;
}

// ---- symmetric buffers: one allocation per rank that every rank maps (AMG level-1 vectors, amg.cu) ----------------------
// layout: [0,64) u64 flag[rank] | [64,72) u64 seq | [72,76) u32 pushed | [76,80) i32 error | [80,84) u32 left | data from 256
constexpr size_t SYM_DATA = SYM_DATA_OFFSET;
typedef SymView SymDev;
static_assert(P2P_MAX_RANKS == 8, "SymView holds eight ranks");

// In-place allgather of [n_ranks][seg] floats at byte offset `off` of the symmetric buffer: every rank pushes its own
// segment into the same place of every peer's buffer, publishes a sequence number and waits for the peers' numbers.
// Same protocol and exit discipline as k_halo; no copy-out: the data land where the consumer kernels read them.
__global__ void __launch_bounds__(256)
k_allgather32(const SymDev* __restrict__ S, size_t off, long long seg) {
    unsigned char* me = S->base[S->rank];
    unsigned long long* seqp = (unsigned long long*)(me + 64);
    int* err = (int*)(me + 76);
    const unsigned long long seq = *seqp + 1;
    const int R = S->n_ranks, rank = S->rank;
    const float* src = (const float*)(me + SYM_DATA + off) + (size_t)rank * seg;
    const long long n4 = seg >> 2;               // seg is a multiple of 32 floats
    const long long total = n4 * (R - 1);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(t / n4);
        const long long k = t - q * n4;
        const int r = q < rank ? q : q + 1;
        float4* dst = (float4*)((float*)(S->base[r] + SYM_DATA + off) + (size_t)rank * seg);
        dst[k] = ((const float4*)src)[k];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    unsigned* done = (unsigned*)(me + 72);
    if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < R && threadIdx.x != rank)
            st_release_sys((unsigned long long*)(S->base[threadIdx.x]) + rank, seq);
    }
    if (threadIdx.x < R && threadIdx.x != rank)
        spin_until((const unsigned long long*)me + threadIdx.x, seq, err);
    __syncthreads();
    unsigned* gone = (unsigned*)(me + 80);
    if (threadIdx.x == 0 && atomicAdd(gone, 1u) == gridDim.x - 1) { *done = 0; *gone = 0; __threadfence(); *seqp = seq; }
}

__global__ void k_pack(const double* __restrict__ x, const int* __restrict__ idx, i64 n, int bs, double* buf) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n * bs) return;
    i64 v = t / bs; int k = (int)(t - v * bs);
    buf[t] = x[(i64)idx[v] * bs + k];
}
__global__ void k_pack32(const float* __restrict__ x, const int* __restrict__ idx, i64 n, int bs, float* buf) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n * bs) return;
    i64 v = t / bs; int k = (int)(t - v * bs);
    buf[t] = x[(i64)idx[v] * bs + k];
}

inline P2P* p2p_of(glims_ctx* c) {
    P2P* p = (P2P*)c->halo.p2p;
    return (p && p->on && c->halo.p2p_enabled) ? p : nullptr;
}

template <typename T>
void exchange_p2p(glims_ctx* c, P2P* p, T* xb, int bs) {
    Halo& h = c->halo;
    const i64 n_ghost = c->n_v - h.n_owned;
    if ((size_t)n_ghost * bs * sizeof(T) > (size_t)p->host.halo_bytes) throw GlError(GLIMS_ERR_ARG, "halo exchange: block too wide for the window");
    const i64 n = std::max(h.n_send, n_ghost) * bs;
    // every CTA waits for the neighbours' flags, and a rank publishes its flag only after ALL of its CTAs have pushed:
    // the grid must be co-resident, or two ranks whose resident CTAs spin while the rest cannot start would wait for each
    // other until the time-out.  Cap it with what the device can hold (occupancy API), not with a hard-coded SM count.
    static int cap = 0;
    if (cap == 0) {
        int per_sm = 1, sms = 1, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_halo<T>, 256, 0);
        cap = std::max(1, std::min(per_sm, 2) * sms);
    }
    int g = (int)std::min<i64>((n + 255) / 256, cap);
    k_halo<T><<<g > 0 ? g : 1, 256, 0, c->stream>>>(p->dev, xb, h.send_idx, h.n_send, bs, h.n_owned, n_ghost);
    c->launches++;
}

// Map every rank's window into this process.  Collective over the NCCL communicator; all ranks agree on the outcome.
void p2p_setup(glims_ctx* c) {
    Halo& h = c->halo;
    if (h.p2p || !h.comm || !h.active) return;
    if (std::getenv("GLIMS_NO_P2P")) return;
    const int R = h.n_ranks;
    if (R > P2P_MAX_RANKS || (int)h.peers.size() > P2P_MAX_RANKS) return;
    ncclComm_t comm = (ncclComm_t)h.comm;
    P2P* p = new P2P();
    h.p2p = p;
    const i64 n_ghost = c->n_v - h.n_owned;
    const size_t halo_bytes = (((size_t)n_ghost * 8 * sizeof(double)) + 255) & ~(size_t)255;     // up to 8 doubles per vertex
    p->win_bytes = OFF_HALO + 2 * halo_bytes;
    GL_CUDA(cudaMalloc(&p->win, p->win_bytes));
    GL_CUDA(cudaMemset(p->win, 0, p->win_bytes));
    // gather: IPC handle (64 B), halo_bytes, and for every rank the ghost offset at which it receives from each rank
    struct Rec { cudaIpcMemHandle_t h; long long halo_bytes; long long recv_off[P2P_MAX_RANKS]; long long recv_cnt[P2P_MAX_RANKS]; };
    Rec mine;
    memset(&mine, 0, sizeof mine);
    cudaError_t e = cudaIpcGetMemHandle(&mine.h, p->win);
    int ok = (e == cudaSuccess);
    if (!ok) cudaGetLastError();
    mine.halo_bytes = (long long)halo_bytes;
    for (int r = 0; r < P2P_MAX_RANKS; ++r) { mine.recv_off[r] = -1; mine.recv_cnt[r] = 0; }
    for (size_t q = 0; q < h.peers.size(); ++q) { mine.recv_off[h.peers[q]] = h.recv_ptr[q]; mine.recv_cnt[h.peers[q]] = h.recv_ptr[q + 1] - h.recv_ptr[q]; }
    Rec* d_all = nullptr;
    GL_CUDA(cudaMalloc(&d_all, sizeof(Rec) * R));
    GL_CUDA(cudaMemcpy(d_all + h.rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice));
    GL_NCCL(ncclAllGather(d_all + h.rank, d_all, sizeof(Rec), ncclChar, comm, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<Rec> all(R);
    GL_CUDA(cudaMemcpy(all.data(), d_all, sizeof(Rec) * R, cudaMemcpyDeviceToHost));
    P2PDev& D = p->host;
    memset(&D, 0, sizeof D);
    D.n_ranks = R; D.rank = h.rank; D.n_peers = (int)h.peers.size();
    D.halo_bytes = (long long)halo_bytes;
    for (int r = 0; r < R && ok; ++r) {
        D.halo_bytes_of[r] = all[r].halo_bytes;
        if (r == h.rank) { D.win[r] = p->win; continue; }
        void* q = nullptr;
        e = cudaIpcOpenMemHandle(&q, all[r].h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        p->opened.push_back(q);
        D.win[r] = (unsigned char*)q;
    }
    for (size_t q = 0; q < h.peers.size(); ++q) {
        D.peer_rank[q] = h.peers[q];
        D.send_ptr[q] = h.send_ptr[q];
        D.dst_off[q] = all[h.peers[q]].recv_off[h.rank];
        D.recv_cnt[q] = h.recv_ptr[q + 1] - h.recv_ptr[q];
        // what I send to q must be what q expects from me
        if ((h.send_ptr[q + 1] - h.send_ptr[q]) != all[h.peers[q]].recv_cnt[h.rank]) ok = 0;
        if ((h.send_ptr[q + 1] - h.send_ptr[q]) > 0 && D.dst_off[q] < 0) ok = 0;
    }
    D.send_ptr[h.peers.size()] = h.n_send;
    // all ranks must agree
    int* d_ok = (int*)d_all;
    GL_CUDA(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    GL_NCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    GL_CUDA(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    if (!ok) {
        if (h.rank == 0) fprintf(stderr, "glims: peer-memory windows unavailable, using NCCL send/recv for the halo exchange\n");
        return;
    }
    GL_CUDA(cudaMalloc(&p->dev, sizeof(P2PDev)));
    GL_CUDA(cudaMemcpy(p->dev, &D, sizeof(P2PDev), cudaMemcpyHostToDevice));
    // nobody may push into a window before its owner has zeroed it: one more rendezvous
    int* d_sync = nullptr;
    GL_CUDA(cudaMalloc(&d_sync, sizeof(int)));
    GL_CUDA(cudaMemset(d_sync, 0, sizeof(int)));
    GL_NCCL(ncclAllReduce(d_sync, d_sync, 1, ncclInt, ncclSum, comm, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_sync);
    p->on = true;
}

}  // namespace

void comm_free(glims_ctx* c) {
    P2P* p = (P2P*)c->halo.p2p;
    if (!p) return;
    for (void* q : p->opened) cudaIpcCloseMemHandle(q);
    if (p->dev) cudaFree(p->dev);
    if (p->win) cudaFree(p->win);
    delete p;
    c->halo.p2p = nullptr;
}

void comm_check(glims_ctx* c) {
    P2P* p = (P2P*)c->halo.p2p;
    if (!p || !p->on) return;
    int err = 0;
    GL_CUDA(cudaMemcpy(&err, p->win + OFF_SEQ + 20, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) throw GlError(GLIMS_ERR_NCCL, "peer-memory exchange timed out waiting for a neighbour rank");
}

void halo_exchange(glims_ctx* c, double* xb, int bs) {
    Halo& h = c->halo;
    if (!h.active || !h.comm) return;
    if (P2P* p = p2p_of(c)) { exchange_p2p<double>(c, p, xb, bs); return; }
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

void halo_exchange_f32(glims_ctx* c, float* xb, int bs) {
    Halo& h = c->halo;
    if (!h.active || !h.comm) return;
    if (P2P* p = p2p_of(c)) { exchange_p2p<float>(c, p, xb, bs); return; }
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
    if (P2P* p = p2p_of(c)) {
        if (n > RED_MAX) throw GlError(GLIMS_ERR_ARG, "allreduce: too many scalars");
        k_allreduce<<<1, 256, 0, c->stream>>>(p->dev, c->scal, slot0, n);
        c->launches++;
        return;
    }
    GL_NCCL(ncclAllReduce(c->scal + slot0, c->scal + slot0, n, ncclDouble, ncclSum, (ncclComm_t)h.comm, c->stream));
}

// ---- symmetric buffers (host side) ------------------------------------------------------------------------------------
struct SymBuf {
    unsigned char* mine = nullptr;
    size_t bytes = 0;
    std::vector<void*> opened;
    SymDev host;
    SymDev* dev = nullptr;
    bool p2p = false;       // every rank could map every other rank's buffer
};

// Collective over the context's NCCL communicator.  `bytes` of payload (zeroed).
void* sym_alloc(glims_ctx* c, size_t bytes) {
    Halo& h = c->halo;
    SymBuf* sb = new SymBuf();
    sb->bytes = SYM_DATA + ((bytes + 255) & ~(size_t)255);
    GL_CUDA(cudaMalloc(&sb->mine, sb->bytes));
    GL_CUDA(cudaMemset(sb->mine, 0, sb->bytes));
    memset(&sb->host, 0, sizeof sb->host);
    sb->host.n_ranks = h.n_ranks; sb->host.rank = h.rank;
    sb->host.base[h.rank] = sb->mine;
    const int R = h.n_ranks;
    if (!h.comm || R <= 1 || R > P2P_MAX_RANKS) return sb;
    ncclComm_t comm = (ncclComm_t)h.comm;
    int ok = (p2p_of(c) != nullptr);            // only where the halo windows could be mapped as well
    cudaIpcMemHandle_t mine_h;
    memset(&mine_h, 0, sizeof mine_h);
    if (ok && cudaIpcGetMemHandle(&mine_h, sb->mine) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    cudaIpcMemHandle_t* d_all = nullptr;
    GL_CUDA(cudaMalloc(&d_all, sizeof(cudaIpcMemHandle_t) * R));
    GL_CUDA(cudaMemcpy(d_all + h.rank, &mine_h, sizeof mine_h, cudaMemcpyHostToDevice));
    GL_NCCL(ncclAllGather(d_all + h.rank, d_all, sizeof(cudaIpcMemHandle_t), ncclChar, comm, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<cudaIpcMemHandle_t> all(R);
    GL_CUDA(cudaMemcpy(all.data(), d_all, sizeof(cudaIpcMemHandle_t) * R, cudaMemcpyDeviceToHost));
    for (int r = 0; r < R && ok; ++r) {
        if (r == h.rank) continue;
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        sb->opened.push_back(q);
        sb->host.base[r] = (unsigned char*)q;
    }
    int* d_ok = (int*)d_all;
    GL_CUDA(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    GL_NCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, c->stream));      // also the "everybody has zeroed" rendezvous
    GL_CUDA(cudaStreamSynchronize(c->stream));
    GL_CUDA(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    sb->p2p = ok != 0;
    if (sb->p2p) {
        GL_CUDA(cudaMalloc(&sb->dev, sizeof(SymDev)));
        GL_CUDA(cudaMemcpy(sb->dev, &sb->host, sizeof(SymDev), cudaMemcpyHostToDevice));
    }
    return sb;
}

void sym_free(void* p) {
    SymBuf* sb = (SymBuf*)p;
    if (!sb) return;
    for (void* q : sb->opened) cudaIpcCloseMemHandle(q);
    if (sb->dev) cudaFree(sb->dev);
    if (sb->mine) cudaFree(sb->mine);
    delete sb;
}

bool sym_is_p2p(void* p) { SymBuf* sb = (SymBuf*)p; return sb && sb->p2p; }
const SymView* sym_view(void* p) { return p ? &((SymBuf*)p)->host : nullptr; }

void* sym_data(void* p) { return p ? ((SymBuf*)p)->mine + SYM_DATA : nullptr; }

bool sym_check(void* p) {           // true: a wait timed out
    SymBuf* sb = (SymBuf*)p;
    if (!sb || !sb->p2p) return false;
    int err = 0;
    cudaMemcpy(&err, sb->mine + 76, sizeof(int), cudaMemcpyDeviceToHost);
    return err != 0;
}

// buf = float vector of n_ranks * seg entries living at `buf` inside the symmetric buffer; every rank owns segment `rank`
void allgather_f32(glims_ctx* c, void* p, float* buf, i64 seg) {
    Halo& h = c->halo;
    SymBuf* sb = (SymBuf*)p;
    if (!sb || !h.comm || h.n_ranks <= 1) return;
    if (sb->p2p && h.p2p_enabled) {
        const size_t off = (size_t)((unsigned char*)buf - (sb->mine + SYM_DATA));
        const i64 total = (seg >> 2) * (h.n_ranks - 1);
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int g = (int)std::min<i64>(std::max<i64>((total + 255) / 256, 1), sms);      // co-resident (see k_halo)
        k_allgather32<<<g, 256, 0, c->stream>>>(sb->dev, off, seg);
        c->launches++;
        return;
    }
    GL_NCCL(ncclAllGather(buf + (size_t)h.rank * seg, buf, seg, ncclFloat, (ncclComm_t)h.comm, c->stream));
}

// setup-time helpers for amg.cu (collective, synchronous)
void comm_allreduce_max_i64(glims_ctx* c, long long* v, int n) {
    Halo& h = c->halo;
    if (!h.comm || h.n_ranks <= 1) return;
    long long* d = nullptr;
    GL_CUDA(cudaMalloc(&d, sizeof(long long) * n));
    GL_CUDA(cudaMemcpy(d, v, sizeof(long long) * n, cudaMemcpyHostToDevice));
    GL_NCCL(ncclAllReduce(d, d, n, ncclInt64, ncclMax, (ncclComm_t)h.comm, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    GL_CUDA(cudaMemcpy(v, d, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
}
void comm_allreduce_max_f64(glims_ctx* c, double* v, int n) {
    Halo& h = c->halo;
    if (!h.comm || h.n_ranks <= 1) return;
    double* d = nullptr;
    GL_CUDA(cudaMalloc(&d, sizeof(double) * n));
    GL_CUDA(cudaMemcpy(d, v, sizeof(double) * n, cudaMemcpyHostToDevice));
    GL_NCCL(ncclAllReduce(d, d, n, ncclDouble, ncclMax, (ncclComm_t)h.comm, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    GL_CUDA(cudaMemcpy(v, d, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
}
// counts[r] = number of `elem_bytes`-sized items rank r contributes; dst (device) receives all items in rank order
void comm_allgatherv(glims_ctx* c, const void* src_dev, void* dst_dev, const std::vector<long long>& counts, size_t elem_bytes) {
    Halo& h = c->halo;
    ncclComm_t comm = (ncclComm_t)h.comm;
    size_t off = 0;
    GL_NCCL(ncclGroupStart());
    for (int r = 0; r < h.n_ranks; ++r) {
        const size_t nbytes = (size_t)counts[r] * elem_bytes;
        if (nbytes > 0)
            GL_NCCL(ncclBroadcast(r == h.rank ? src_dev : (const void*)((unsigned char*)dst_dev + off), (unsigned char*)dst_dev + off,
                                  nbytes, ncclChar, r, comm, c->stream));
        off += nbytes;
    }
    GL_NCCL(ncclGroupEnd());
    GL_CUDA(cudaStreamSynchronize(c->stream));
}
void comm_allgather_i64(glims_ctx* c, long long mine, std::vector<long long>& all) {
    Halo& h = c->halo;
    all.assign(h.n_ranks, 0);
    long long* d = nullptr;
    GL_CUDA(cudaMalloc(&d, sizeof(long long) * h.n_ranks));
    GL_CUDA(cudaMemcpy(d + h.rank, &mine, sizeof(long long), cudaMemcpyHostToDevice));
    GL_NCCL(ncclAllGather(d + h.rank, d, 1, ncclInt64, (ncclComm_t)h.comm, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    GL_CUDA(cudaMemcpy(all.data(), d, sizeof(long long) * h.n_ranks, cudaMemcpyDeviceToHost));
    cudaFree(d);
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
        // every rank of a multi-rank run takes part in every collective, also one whose share happens to have no ghost
        // vertices (it would otherwise skip the allreduces and leave the others waiting)
        if (n_ranks > 1 && !c->halo.active) { c->halo.active = true; c->halo.n_owned = c->n_v; }
        if (c->halo.send_idx) p2p_setup(c);      // halo plan known (glims_set_halo comes first): map the peer windows
    } catch (const GlError& e) { c->err = e.msg; return e.code; }
    return GLIMS_OK;
}

int glims_set_halo(glims_ctx* c, int32_t n_peers, const int32_t* peers, const int64_t* send_ptr,
                   const int32_t* send_idx, const int64_t* recv_ptr) {
    if (!c || n_peers < 0) return GLIMS_ERR_ARG;
    try {
        GL_CUDA(cudaSetDevice(c->device));
        Halo& h = c->halo;
        if (h.p2p) throw GlError(GLIMS_ERR_STATE, "set_halo: the halo plan cannot change after glims_comm_init");
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

/* Switch between the peer-memory transport (1, default when the windows could be mapped) and NCCL (0) at run time;
   must be called with the same value on every rank.  Returns 1 if peer memory is in use afterwards, 0 if not. */
int glims_set_p2p(glims_ctx* c, int32_t on) {
    if (!c) return GLIMS_ERR_ARG;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->halo.p2p_enabled = on != 0;
    solver_free_graphs(c);          // captured PCG iterations contain the old transport's calls
    return p2p_of(c) ? 1 : 0;
}

/* Average device time (microseconds, CUDA events on the context stream) of `reps` back-to-back collectives:
   kind 0 = halo exchange of the state vector (FP64, dim+1 values per vertex), 1 = FP32 halo exchange with dim values
   per vertex (the V-cycle's), 2 = allreduce of two scalars. */
int glims_comm_bench(glims_ctx* c, int32_t kind, int32_t reps, float* us_avg) {
    if (!c || !us_avg || reps <= 0) return GLIMS_ERR_ARG;
    try {
        GL_CUDA(cudaSetDevice(c->device));
        float* tmp32 = nullptr;
        if (kind == 1) { GL_CUDA(cudaMalloc(&tmp32, sizeof(float) * c->n_v * c->dim)); GL_CUDA(cudaMemset(tmp32, 0, sizeof(float) * c->n_v * c->dim)); }
        auto run = [&]() {
            if (kind == 0) halo_exchange(c, c->dx, c->nb);
            else if (kind == 1) halo_exchange_f32(c, tmp32, c->dim);
            else allreduce_scalars(c, S_TMP2, 2);
        };
        for (int w = 0; w < 5; ++w) run();
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, c->stream);
        for (int r = 0; r < reps; ++r) run();
        cudaEventRecord(b, c->stream);
        GL_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        cudaEventDestroy(a); cudaEventDestroy(b);
        if (tmp32) cudaFree(tmp32);
        *us_avg = 1e3f * ms / reps;
        comm_check(c);
    } catch (const GlError& e) { c->err = e.msg; return e.code; }
    return GLIMS_OK;
}

}  // extern "C"
