// Sparsity pattern and scatter maps, built on the device (replaces DOLFIN's SparsityPatternBuilder
// and the per-call MatSetValues searches with maps that are reused by every assembly).
#include "common.h"
#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>
#include <thrust/sort.h>
#include <thrust/unique.h>
#include <thrust/scan.h>
#include <thrust/binary_search.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/reduce.h>
#include <thrust/extrema.h>
#include <algorithm>
#include <vector>

namespace {

__global__ void k_gen_keys(const int* __restrict__ cells, i64 n_c, int nb, i64 n_own, unsigned long long* keys) {
    i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    i64 np = n_c * nb * nb;
    if (p >= np) return;
    i64 e = p / (nb * nb);
    int ab = (int)(p - e * nb * nb);
    int a = ab / nb, b = ab - a * nb;
    unsigned va = (unsigned)cells[e * nb + a], vb = (unsigned)cells[e * nb + b];
    // rows exist only for owned vertices; ghost rows get the sentinel key (sorted last, removed)
    keys[p] = (va < (unsigned long long)n_own) ? (((unsigned long long)va << 32) | vb) : ~0ULL;
}

struct RowStart {
    __host__ __device__ unsigned long long operator()(i64 r) const { return (unsigned long long)r << 32; }
};

__global__ void k_slice_width(const i64* __restrict__ rowptr, int n_rows, int n_slices, int* slice_w) {
    int S = blockIdx.x * blockDim.x + threadIdx.x;
    if (S >= n_slices) return;
    int w = 0;
    for (int l = 0; l < SLICE; ++l) {
        int r = S * SLICE + l;
        if (r < n_rows) w = max(w, (int)(rowptr[r + 1] - rowptr[r]));
    }
    slice_w[S] = w;
}

__global__ void k_slice_off(const int* __restrict__ slice_w, int n_slices, i64* off_in_out) {
    // off_in_out holds the exclusive scan of widths; convert to slots
    int S = blockIdx.x * blockDim.x + threadIdx.x;
    if (S > n_slices) return;
    off_in_out[S] *= SLICE;
}

__global__ void k_init_col(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, int n_rows,
                           int n_slices, int* col) {
    int S = blockIdx.x;
    i64 base = slice_off[S];
    int w = slice_w[S];
    for (int t = threadIdx.x; t < w * SLICE; t += blockDim.x) {
        int r = S * SLICE + (t & 31);
        col[base + t] = r < n_rows ? r : 0;
    }
}

__global__ void k_fill_col(const unsigned long long* __restrict__ ukeys, i64 nnzb, const i64* __restrict__ rowptr,
                           const i64* __restrict__ slice_off, int* col, int* diag) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= nnzb) return;
    unsigned long long k = ukeys[t];
    int r = (int)(k >> 32), cidx = (int)(k & 0xffffffffu);
    int j = (int)(t - rowptr[r]);
    i64 s = slice_off[r >> 5] + (i64)j * SLICE + (r & 31);
    col[s] = cidx;
    if (cidx == r) diag[r] = (int)s;
}

__global__ void k_eslot(const int* __restrict__ cells, i64 n_c, int nb, i64 n_own,
                        const unsigned long long* __restrict__ ukeys, i64 nnzb, const i64* __restrict__ rowptr,
                        const i64* __restrict__ slice_off, int* eslot) {
    i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    i64 np = n_c * nb * nb;
    if (p >= np) return;
    i64 e = p / (nb * nb);
    int ab = (int)(p - e * nb * nb);
    int a = ab / nb, b = ab - a * nb;
    int va = cells[e * nb + a], vb = cells[e * nb + b];
    if (va >= n_own) { eslot[p] = -1; return; }
    unsigned long long key = ((unsigned long long)(unsigned)va << 32) | (unsigned)vb;
    i64 lo = rowptr[va], hi = rowptr[va + 1];
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if (ukeys[mid] < key) lo = mid + 1; else hi = mid;
    }
    int j = (int)(lo - rowptr[va]);
    eslot[p] = (int)(slice_off[va >> 5] + (i64)j * SLICE + (va & 31));
}

__global__ void k_gather_keys(const int* __restrict__ eslot, i64 np, unsigned long long* keys, i64 n_slots) {
    i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (p >= np) return;
    int s = eslot[p];
    // key = slot (high) | pair id (low 36 bits fits 2^36 pairs)
    keys[p] = s < 0 ? ~0ULL : (((unsigned long long)s << 36) | (unsigned long long)p);
}

__global__ void k_gather_unpack(const unsigned long long* __restrict__ keys, i64 n, int nbnb, int nb, int* gent) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned long long p = keys[t] & ((1ULL << 36) - 1);
    i64 e = (i64)(p / nbnb);
    int ab = (int)(p - (unsigned long long)e * nbnb);
    int a = ab / nb, b = ab - a * nb;
    gent[t] = (int)((e << 4) | (a << 2) | b);
}

// ---- slice-local assembly maps ---------------------------------------------------------------------
__global__ void k_slice_elem_keys(const int* __restrict__ cells, i64 n_c, int nb, i64 n_own, unsigned long long* keys) {
    i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (p >= n_c * nb) return;
    i64 e = p / nb;
    int v = cells[p];
    keys[p] = (v < n_own) ? (((unsigned long long)(v >> 5) << 32) | (unsigned long long)e) : ~0ULL;
}
struct SliceStart {
    __host__ __device__ unsigned long long operator()(i64 S) const { return (unsigned long long)S << 32; }
};
__global__ void k_slice_elem_unpack(const unsigned long long* __restrict__ keys, i64 n, int* out) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t < n) out[t] = (int)(keys[t] & 0xffffffffu);
}
// contributor entries re-expressed with the element's index inside its slice's element list
__global__ void k_local_entries(const i64* __restrict__ slice_off, const i64* __restrict__ gptr,
                                const int* __restrict__ gent, const i64* __restrict__ sl_ptr,
                                const int* __restrict__ sl_elem, unsigned short* __restrict__ lent) {
    const int S = blockIdx.x;
    const i64 t0 = gptr[slice_off[S]], t1 = gptr[slice_off[S + 1]];
    const i64 p0 = sl_ptr[S], p1 = sl_ptr[S + 1];
    for (i64 t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const int ent = gent[t];
        const int e = ent >> 4;
        i64 lo = p0, hi = p1;
        while (lo < hi) { i64 mid = (lo + hi) >> 1; if (sl_elem[mid] < e) lo = mid + 1; else hi = mid; }
        lent[t] = (unsigned short)(((lo - p0) << 4) | (ent & 15));
    }
}

struct SlotStart {
    __host__ __device__ unsigned long long operator()(i64 s) const { return (unsigned long long)s << 36; }
};

inline int nblk(i64 n, int t = 256) { return (int)((n + t - 1) / t); }

}  // namespace

void build_pattern_from_keys(cudaStream_t st, unsigned long long* kp, i64 np, i64 n_own, SellPattern& P,
                             unsigned long long** ukeys_out) {
    auto pol = thrust::cuda::par.on(st);
    thrust::device_ptr<unsigned long long> keys(kp);
    thrust::sort(pol, keys, keys + np);
    auto uend = thrust::unique(pol, keys, keys + np);
    i64 nu = uend - keys;
    if (nu > 0) {   // drop the sentinel if present
        unsigned long long last;
        GL_CUDA(cudaMemcpyAsync(&last, kp + nu - 1, 8, cudaMemcpyDeviceToHost, st));
        GL_CUDA(cudaStreamSynchronize(st));
        if (last == ~0ULL) nu -= 1;
    }
    P.n_rows = (int)n_own;
    P.nnzb = nu;
    P.n_slices = (int)((n_own + SLICE - 1) / SLICE);
    GL_CUDA(cudaMalloc(&P.rowptr, sizeof(i64) * (n_own + 1)));
    thrust::device_ptr<i64> rp(P.rowptr);
    thrust::lower_bound(pol, keys, keys + nu,
                        thrust::make_transform_iterator(thrust::counting_iterator<i64>(0), RowStart()),
                        thrust::make_transform_iterator(thrust::counting_iterator<i64>(n_own + 1), RowStart()), rp);
    GL_CUDA(cudaMalloc(&P.slice_w, sizeof(int) * P.n_slices));
    GL_CUDA(cudaMalloc(&P.slice_off, sizeof(i64) * (P.n_slices + 1)));
    k_slice_width<<<nblk(P.n_slices), 256, 0, st>>>(P.rowptr, P.n_rows, P.n_slices, P.slice_w);
    thrust::device_ptr<int> swp(P.slice_w);
    thrust::device_ptr<i64> sop(P.slice_off);
    GL_CUDA(cudaMemsetAsync(P.slice_off, 0, sizeof(i64) * (P.n_slices + 1), st));
    thrust::inclusive_scan(pol, swp, swp + P.n_slices, sop + 1, thrust::plus<i64>());
    P.max_w = thrust::reduce(pol, swp, swp + P.n_slices, 0, thrust::maximum<int>());
    k_slice_off<<<nblk(P.n_slices + 1), 256, 0, st>>>(P.slice_w, P.n_slices, P.slice_off);
    GL_CUDA(cudaMemcpyAsync(&P.n_slots, P.slice_off + P.n_slices, sizeof(i64), cudaMemcpyDeviceToHost, st));
    GL_CUDA(cudaStreamSynchronize(st));
    GL_CUDA(cudaMalloc(&P.col, sizeof(int) * (P.n_slots > 0 ? P.n_slots : 1)));
    GL_CUDA(cudaMalloc(&P.diag, sizeof(int) * n_own));
    k_init_col<<<P.n_slices, 128, 0, st>>>(P.slice_off, P.slice_w, P.n_rows, P.n_slices, P.col);
    k_fill_col<<<nblk(nu), 256, 0, st>>>(kp, nu, P.rowptr, P.slice_off, P.col, P.diag);
    if (ukeys_out) {
        GL_CUDA(cudaMalloc(ukeys_out, sizeof(unsigned long long) * (nu > 0 ? nu : 1)));
        GL_CUDA(cudaMemcpyAsync(*ukeys_out, kp, sizeof(unsigned long long) * nu, cudaMemcpyDeviceToDevice, st));
    }
    GL_CUDA(cudaStreamSynchronize(st));
}

void free_pattern(SellPattern& P) {
    for (void* q : {(void*)P.slice_off, (void*)P.slice_w, (void*)P.col, (void*)P.rowptr, (void*)P.diag})
        if (q) cudaFree(q);
    P = SellPattern();
}

void build_pattern_generic(cudaStream_t st, const int* cells, i64 n_c, int nb, i64 n_own, SellPattern& P, int** eslot_out) {
    i64 np = n_c * nb * nb;
    thrust::device_vector<unsigned long long> keys(np);
    unsigned long long* kp = thrust::raw_pointer_cast(keys.data());
    k_gen_keys<<<nblk(np), 256, 0, st>>>(cells, n_c, nb, n_own, kp);
    build_pattern_from_keys(st, kp, np, n_own, P, nullptr);
    if (eslot_out) {
        GL_CUDA(cudaMalloc(eslot_out, sizeof(int) * np));
        k_eslot<<<nblk(np), 256, 0, st>>>(cells, n_c, nb, n_own, kp, P.nnzb, P.rowptr, P.slice_off, *eslot_out);
    }
    GL_CUDA(cudaStreamSynchronize(st));
}

void build_pattern(glims_ctx* c) {
    i64 n_own = c->halo.active ? c->halo.n_owned : c->n_v;
    build_pattern_generic(c->stream, c->cells, c->n_c, c->nb, n_own, c->pat, &c->eslot);
}

// Transposed scatter map: for every matrix slot the list of (element, a, b) that contribute to it.
void build_gather_map(glims_ctx* c) {
    if (c->have_gather) return;
    auto pol = thrust::cuda::par.on(c->stream);
    int nb = c->nb;
    i64 np = c->n_c * nb * nb;
    if ((c->n_c << 4) > 0x7fffffffLL) throw GlError(GLIMS_ERR_ARG, "gather map: too many cells for packed entries");
    thrust::device_vector<unsigned long long> keys(np);
    unsigned long long* kp = thrust::raw_pointer_cast(keys.data());
    k_gather_keys<<<nblk(np), 256, 0, c->stream>>>(c->eslot, np, kp, c->pat.n_slots);
    thrust::sort(pol, keys.begin(), keys.end());
    // entries with sentinel keys (ghost rows) sort last
    i64 nvalid = thrust::lower_bound(pol, keys.begin(), keys.end(), ~0ULL) - keys.begin();
    GL_CUDA(cudaMalloc(&c->gptr, sizeof(i64) * (c->pat.n_slots + 1)));
    GL_CUDA(cudaMalloc(&c->gent, sizeof(int) * (nvalid > 0 ? nvalid : 1)));
    thrust::device_ptr<i64> gp(c->gptr);
    thrust::lower_bound(pol, keys.begin(), keys.begin() + nvalid,
                        thrust::make_transform_iterator(thrust::counting_iterator<i64>(0), SlotStart()),
                        thrust::make_transform_iterator(thrust::counting_iterator<i64>(c->pat.n_slots + 1), SlotStart()),
                        gp);
    k_gather_unpack<<<nblk(nvalid), 256, 0, c->stream>>>(kp, nvalid, nb * nb, nb, c->gent);
    GL_CUDA(cudaStreamSynchronize(c->stream));
    c->have_gather = true;
}

// Per SELL slice: the sorted list of elements touching one of its 32 rows, and the contributor entries of the
// gather map rewritten against that list (12-bit local element index | a | b) for the slice-local kernel.
void build_slice_map(glims_ctx* c) {
    if (c->have_slice) return;
    build_gather_map(c);
    auto pol = thrust::cuda::par.on(c->stream);
    const int nb = c->nb;
    const i64 n_own = c->pat.n_rows, np = c->n_c * nb;
    const int n_slices = c->pat.n_slices;
    thrust::device_vector<unsigned long long> keys(np);
    unsigned long long* kp = thrust::raw_pointer_cast(keys.data());
    k_slice_elem_keys<<<nblk(np), 256, 0, c->stream>>>(c->cells, c->n_c, nb, n_own, kp);
    thrust::sort(pol, keys.begin(), keys.end());
    auto uend = thrust::unique(pol, keys.begin(), keys.end());
    i64 nu = uend - keys.begin();
    if (nu > 0) {
        unsigned long long last;
        GL_CUDA(cudaMemcpyAsync(&last, kp + nu - 1, 8, cudaMemcpyDeviceToHost, c->stream));
        GL_CUDA(cudaStreamSynchronize(c->stream));
        if (last == ~0ULL) nu -= 1;
    }
    GL_CUDA(cudaMalloc(&c->sl_ptr, sizeof(i64) * (n_slices + 1)));
    GL_CUDA(cudaMalloc(&c->sl_elem, sizeof(int) * (nu > 0 ? nu : 1)));
    thrust::device_ptr<i64> sp(c->sl_ptr);
    thrust::lower_bound(pol, keys.begin(), keys.begin() + nu,
                        thrust::make_transform_iterator(thrust::counting_iterator<i64>(0), SliceStart()),
                        thrust::make_transform_iterator(thrust::counting_iterator<i64>(n_slices + 1), SliceStart()), sp);
    k_slice_elem_unpack<<<nblk(nu), 256, 0, c->stream>>>(kp, nu, c->sl_elem);
    // largest element list decides the shared-memory footprint
    std::vector<i64> h(n_slices + 1);
    GL_CUDA(cudaMemcpyAsync(h.data(), c->sl_ptr, sizeof(i64) * (n_slices + 1), cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    i64 mx = 0;
    for (int S = 0; S < n_slices; ++S) mx = std::max(mx, h[S + 1] - h[S]);
    c->sl_max = (int)mx;
    c->sl_total = nu;
    i64 n_ent;
    GL_CUDA(cudaMemcpy(&n_ent, c->gptr + c->pat.n_slots, sizeof(i64), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMalloc(&c->lent, sizeof(unsigned short) * (n_ent > 0 ? n_ent : 1)));
    if (mx < 4096)
        k_local_entries<<<n_slices, 256, 0, c->stream>>>(c->pat.slice_off, c->gptr, c->gent, c->sl_ptr, c->sl_elem, c->lent);
    GL_CUDA(cudaStreamSynchronize(c->stream));
    c->have_slice = true;
}
