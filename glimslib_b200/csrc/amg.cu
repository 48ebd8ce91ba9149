// Aggregation AMG preconditioner for the constant elasticity block K_uu (K5).
//
// Hierarchy: greedy vertex aggregation (host, on the vertex graph); tentative prolongator built from the
// rigid-body modes of each aggregate about its centroid -- d translations + (1 | 3) rotations, so coarse
// levels carry 3 (2D) or 6 (3D) unknowns per aggregate; Galerkin operators P^T A P accumulated on the
// device into SELL-32 block matrices; Chebyshev(D^-1 A) smoothing with block-Jacobi D; dense inverse on the
// coarsest level.  Everything in the V-cycle is a device kernel on the context stream: no host sync.
// Fully constrained (Dirichlet) vertices and ghost vertices are left out of the coarse space, so on a
// partitioned mesh the preconditioner is rank-local (block-Jacobi over sub-domains) with no communication.
#include "common.h"
#include "tma.cuh"
#include <cuda_fp16.h>
#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>
#include <thrust/transform_reduce.h>
#include <thrust/functional.h>
#include <thrust/execution_policy.h>
#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>

namespace {

constexpr int TPB = 256;
inline int nblk(i64 n, int t = TPB) { return (int)((n + t - 1) / t); }

struct Level {
    int bs = 0;                 // unknowns per node on this level
    int n = 0;                  // nodes (block rows)
    i64 n_cols = 0;             // length of x in nodes (level 0: local vertices incl. ghosts)
    SellPattern pat;
    bool owns_pat = false;
    double* A = nullptr;        // [n_slots][bs*bs] SELL value layout
    bool owns_A = false;
    double* dinv = nullptr;     // [n][bs*bs]
    double lmax = 0;
    // transfer to the next (coarser) level
    int bsc = 0, nc = 0;
    int* agg = nullptr;         // [n_cols] aggregate of node, -1 = not in the coarse space
    double* rvec = nullptr;     // [n][dim] node position minus aggregate centroid
    unsigned char* free_mask = nullptr;   // level 0: bit k set <=> dof k is free
    int *mem_ptr = nullptr, *mem_idx = nullptr;   // members of each aggregate
    double* X = nullptr;        // [n][dim] node positions
    // work vectors
    double *x = nullptr, *b = nullptr, *r = nullptr, *d = nullptr;
    // FP32 copies for the mixed-precision V-cycle (the outer PCG stays FP64)
    float *A32 = nullptr, *dinv32 = nullptr, *x32 = nullptr, *b32 = nullptr, *r32 = nullptr, *d32 = nullptr;
    __half* A16 = nullptr;      // fine level only (optional): matrix values in FP16, scaled by a power of two
    float a16_unscale = 1.f;    // multiply row sums by this to undo the scaling
    float* y32 = nullptr;       // second iterate buffer: the fused smoother steps ping-pong between x32 and y32
    // Partitioned mesh (one rank per GPU): level 0 is distributed by rows (ghost halo exchange); level 1 is the GLOBAL
    // Galerkin operator, replicated on every rank, of which a rank applies only the rows of its own aggregates
    // [row0, row0 + rows) and all-gathers the vector segments; levels >= 2 are small and run redundantly on every rank.
    bool dist_rows = false;     // level 1 of a distributed hierarchy
    int row0 = 0, rows = 0;     // own row range (multiples of 32)
    int nc_off = 0;             // level 0 of a distributed hierarchy: first global id of this rank's aggregates
    int nc_local = 0;           //   ... and their number (l.nc is the size of the coarse level)
    void* sym = nullptr;        // dist_rows: symmetric buffer holding x32 | y32 | r32 (comm.cu)
};

}  // namespace

struct Amg {
    int dim = 0;
    std::vector<Level> L;
    double* coarse_inv = nullptr;
    float* coarse_inv32 = nullptr;
    int coarse_m = 0;
    int cheb_degree = 2;
    int coarse_degree = 2;      // smoother degree on levels >= 1 (GLIMS_AMG_COARSE_DEGREE)
    double cheb_ratio = 0.1;
    bool dist = false;          // hierarchy of a partitioned mesh (see Level::dist_rows)
    int n_ranks = 1, rank = 0;
    void* fused = nullptr;      // FusedPlan: levels >= 2 of the FP32 V-cycle in one persistent kernel (amg_fused.cuh)
};

namespace {

// ---- small dense helpers on the device ------------------------------------------------------------
// P_i (bs_f x bsc): level 0: diag(free) [I_d | R(r)] ; higher levels: [[I_d, R(r)], [0, I_rot]]
template <int D>
__device__ inline void build_P(bool level0, const double* r, unsigned mask, double (&P)[6][6], int& bsf, int& bsc) {
    constexpr int NR = (D == 2) ? 1 : 3;
    bsc = D + NR;
    bsf = level0 ? D : bsc;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) P[i][j] = 0.0;
    for (int i = 0; i < D; ++i) P[i][i] = 1.0;
    if (D == 2) { P[0][2] = -r[1]; P[1][2] = r[0]; }
    else {
        P[0][4] = r[2];  P[0][5] = -r[1];
        P[1][3] = -r[2]; P[1][5] = r[0];
        P[2][3] = r[1];  P[2][4] = -r[0];
    }
    if (level0) {
        for (int i = 0; i < D; ++i)
            if (!((mask >> i) & 1u)) for (int j = 0; j < 6; ++j) P[i][j] = 0.0;
    } else {
        for (int k = 0; k < NR; ++k) P[D + k][D + k] = 1.0;
    }
}

// keys of the coarse pattern: one per fine slot
__global__ void k_coarse_keys(const i64* __restrict__ slice_off, const int* __restrict__ slice_w,
                              const int* __restrict__ col, int n_rows, const int* __restrict__ agg,
                              unsigned long long* keys) {
    const int S = blockIdx.x;
    const i64 base = slice_off[S];
    const int w = slice_w[S];
    for (int t = threadIdx.x; t < w * 32; t += blockDim.x) {
        const int r = S * 32 + (t & 31);
        unsigned long long key = ~0ULL;
        if (r < n_rows) {
            int I = agg[r], J = agg[col[base + t]];
            if (I >= 0 && J >= 0) key = ((unsigned long long)(unsigned)I << 32) | (unsigned)J;
        }
        keys[base + t] = key;
    }
}

// diagonal keys (I, I) for I in [first, first + n): every node of a rank's segment gets a diagonal block, padding included
__global__ void k_diag_keys(unsigned long long* keys, int first, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((unsigned long long)(unsigned)(first + i) << 32) | (unsigned)(first + i);
}
// SELL values <-> CSR-ordered blocks [t][nc]
__global__ void k_sell_to_csr(const i64* __restrict__ rowptr, const i64* __restrict__ slice_off, int n_rows, int nc,
                              const double* __restrict__ A, double* __restrict__ out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const i64 base = slice_off[r >> 5];
    for (i64 t = rowptr[r]; t < rowptr[r + 1]; ++t) {
        const i64 s = base + (t - rowptr[r]) * 32 + (r & 31);
        for (int k = 0; k < nc; ++k) out[t * nc + k] = A[vidx(s, k, nc)];
    }
}
__global__ void k_csr_to_sell(const i64* __restrict__ rowptr, const i64* __restrict__ slice_off, int n_rows, int nc,
                              const double* __restrict__ in, double* __restrict__ A) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const i64 base = slice_off[r >> 5];
    for (i64 t = rowptr[r]; t < rowptr[r + 1]; ++t) {
        const i64 s = base + (t - rowptr[r]) * 32 + (r & 31);
        for (int k = 0; k < nc; ++k) A[vidx(s, k, nc)] = in[t * nc + k];
    }
}

template <int D>
__global__ void k_galerkin(const i64* __restrict__ slice_off, const int* __restrict__ slice_w,
                           const int* __restrict__ col, int n_rows, bool level0, const double* __restrict__ A,
                           const int* __restrict__ agg, const double* __restrict__ rvec,
                           const unsigned char* __restrict__ free_mask, const unsigned long long* __restrict__ ukeys,
                           const i64* __restrict__ c_rowptr, const i64* __restrict__ c_slice_off, double* Ac) {
    const int S = blockIdx.x;
    const i64 base = slice_off[S];
    const int w = slice_w[S];
    for (int t = threadIdx.x; t < w * 32; t += blockDim.x) {
        const int r = S * 32 + (t & 31);
        if (r >= n_rows) continue;
        const i64 s = base + t;
        const int cj = col[s];
        const int I = agg[r], J = agg[cj];
        if (I < 0 || J < 0) continue;
        double Pi[6][6], Pj[6][6];
        int bsf, bsc;
        build_P<D>(level0, rvec + (i64)r * D, level0 ? free_mask[r] : 0xffu, Pi, bsf, bsc);
        build_P<D>(level0, rvec + (i64)cj * D, level0 ? free_mask[cj] : 0xffu, Pj, bsf, bsc);
        double a[6][6];
        bool nz = false;
        for (int i = 0; i < bsf; ++i)
            for (int j = 0; j < bsf; ++j) { a[i][j] = A[vidx(s, i * bsf + j, bsf * bsf)]; nz |= (a[i][j] != 0.0); }
        if (!nz) continue;
        double B[6][6];   // A Pj
        for (int i = 0; i < bsf; ++i)
            for (int j = 0; j < bsc; ++j) {
                double v = 0;
                for (int k = 0; k < bsf; ++k) v += a[i][k] * Pj[k][j];
                B[i][j] = v;
            }
        // coarse slot
        unsigned long long key = ((unsigned long long)(unsigned)I << 32) | (unsigned)J;
        i64 lo = c_rowptr[I], hi = c_rowptr[I + 1];
        while (lo < hi) { i64 mid = (lo + hi) >> 1; if (ukeys[mid] < key) lo = mid + 1; else hi = mid; }
        i64 cs = c_slice_off[I >> 5] + (lo - c_rowptr[I]) * 32 + (I & 31);
        for (int i = 0; i < bsc; ++i)
            for (int j = 0; j < bsc; ++j) {
                double v = 0;
                for (int k = 0; k < bsf; ++k) v += Pi[k][i] * B[k][j];
                if (v != 0.0) atomicAdd(&Ac[vidx(cs, i * bsc + j, bsc * bsc)], v);
            }
    }
}

// r_c[I] = sum_{i in I} P_i^T r_f[i]   (one thread per aggregate, fixed order: deterministic)
template <int D>
__global__ void k_restrict(int nc, const int* __restrict__ mem_ptr, const int* __restrict__ mem_idx, bool level0,
                           const double* __restrict__ rvec, const unsigned char* __restrict__ free_mask,
                           const double* __restrict__ rf, double* __restrict__ rc) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc) return;
    constexpr int NR = (D == 2) ? 1 : 3;
    const int bsc = D + NR, bsf = level0 ? D : bsc;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int m = mem_ptr[I]; m < mem_ptr[I + 1]; ++m) {
        const int i = mem_idx[m];
        double P[6][6];
        int a_, b_;
        build_P<D>(level0, rvec + (i64)i * D, level0 ? free_mask[i] : 0xffu, P, a_, b_);
        for (int k = 0; k < bsf; ++k) {
            double v = rf[(i64)i * bsf + k];
            for (int j = 0; j < bsc; ++j) acc[j] += P[k][j] * v;
        }
    }
    for (int j = 0; j < bsc; ++j) rc[(i64)I * bsc + j] = acc[j];
}

// x_f[i] += P_i x_c[agg[i]]
template <int D>
__global__ void k_prolong_add(int n, const int* __restrict__ agg, bool level0, const double* __restrict__ rvec,
                              const unsigned char* __restrict__ free_mask, const double* __restrict__ xc,
                              double* __restrict__ xf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int I = agg[i];
    if (I < 0) return;
    constexpr int NR = (D == 2) ? 1 : 3;
    const int bsc = D + NR, bsf = level0 ? D : bsc;
    double P[6][6];
    int a_, b_;
    build_P<D>(level0, rvec + (i64)i * D, level0 ? free_mask[i] : 0xffu, P, a_, b_);
    for (int k = 0; k < bsf; ++k) {
        double v = 0;
        for (int j = 0; j < bsc; ++j) v += P[k][j] * xc[(i64)I * bsc + j];
        xf[(i64)i * bsf + k] += v;
    }
}

// Chebyshev step: z = Dinv r ; d = c1 d + c2 z ; x += d      (first step: c1 = 0)
template <int BS>
__global__ void k_cheb_update(const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ d,
                              double* __restrict__ x, int n, double c1, double c2) {
    for (i64 row = blockIdx.x * (i64)TPB + threadIdx.x; row < n; row += (i64)gridDim.x * TPB) {
        double rv[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) rv[i] = r[row * BS + i];
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            double z = 0;
#pragma unroll
            for (int j = 0; j < BS; ++j) z += dinv[row * BS * BS + i * BS + j] * rv[j];
            double dn = c2 * z + (c1 != 0.0 ? c1 * d[row * BS + i] : 0.0);
            d[row * BS + i] = dn;
            x[row * BS + i] += dn;
        }
    }
}

template <int BS>
__global__ void k_block_diag_inverse(const int* __restrict__ diag, int n, const double* __restrict__ A, double* out,
                                     double reg) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    i64 s = diag[r];
    double a[BS][BS], inv[BS][BS];
    double scale = 0;
    for (int i = 0; i < BS; ++i)
        for (int j = 0; j < BS; ++j) { a[i][j] = A[vidx(s, i * BS + j, BS * BS)]; inv[i][j] = i == j; }
    for (int i = 0; i < BS; ++i) scale = fmax(scale, fabs(a[i][i]));
    // rows with an empty diagonal (node outside the coarse space / fully constrained) -> identity
    for (int i = 0; i < BS; ++i) if (a[i][i] == 0.0) a[i][i] = scale > 0 ? scale : 1.0;
    // coarse blocks of tiny aggregates can be rank deficient (a rotation that moves no member): Tikhonov shift
    for (int i = 0; i < BS; ++i) a[i][i] += reg * scale;
    for (int p = 0; p < BS; ++p) {     // Gauss-Jordan with partial pivoting
        int piv = p; double best = fabs(a[p][p]);
        for (int i = p + 1; i < BS; ++i) if (fabs(a[i][p]) > best) { best = fabs(a[i][p]); piv = i; }
        if (piv != p) for (int j = 0; j < BS; ++j) { double t = a[p][j]; a[p][j] = a[piv][j]; a[piv][j] = t; t = inv[p][j]; inv[p][j] = inv[piv][j]; inv[piv][j] = t; }
        double ip = 1.0 / a[p][p];
        for (int j = 0; j < BS; ++j) { a[p][j] *= ip; inv[p][j] *= ip; }
        for (int i = 0; i < BS; ++i) {
            if (i == p) continue;
            double f = a[i][p];
            for (int j = 0; j < BS; ++j) { a[i][j] -= f * a[p][j]; inv[i][j] -= f * inv[p][j]; }
        }
    }
    for (int i = 0; i < BS; ++i) for (int j = 0; j < BS; ++j) out[(i64)r * BS * BS + i * BS + j] = inv[i][j];
}

// In-place Gauss-Jordan inverse of a (regularised) SPD matrix on the device: M (m x m, row major) is reduced to the
// identity while Inv (identity on entry) becomes its inverse.  SPD => no pivoting.  Two small kernels per pivot (scale the
// pivot row + save the pivot column; rank-one update of both halves with a full grid): ~10 us per pivot.  Only used for
// the coarsest level when it is too large for the host loop (m up to ~1500).
__global__ void k_gj_pivot(double* M, double* Inv, int m, int p, double* fcol) {
    __shared__ double ip;
    if (threadIdx.x == 0) ip = 1.0 / M[(i64)p * m + p];
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) fcol[i] = (i == p) ? 0.0 : M[(i64)i * m + p];
    for (int j = threadIdx.x; j < m; j += blockDim.x) { M[(i64)p * m + j] *= ip; Inv[(i64)p * m + j] *= ip; }
}
__global__ void k_gj_update(double* M, double* Inv, int m, int p, const double* __restrict__ fcol) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;       // column of [M | Inv]
    const int i = blockIdx.y;                                  // row
    if (j >= 2 * m) return;
    const double f = fcol[i];
    if (f == 0.0) return;
    double* T = j < m ? M : Inv;
    const int jj = j < m ? j : j - m;
    T[(i64)i * m + jj] -= f * T[(i64)p * m + jj];
}

__global__ void k_dense_matvec(const double* __restrict__ M, const double* __restrict__ x, double* __restrict__ y, int m) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= m) return;
    double s = 0;
    for (int j = lane; j < m; j += 32) s += M[(i64)row * m + j] * x[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
}

__global__ void k_fill_pseudo_random(double* v, i64 n) {
    i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 29;
    v[i] = (double)(z & 0xffffff) / (double)0xffffff - 0.5;
}

// ---- host helpers -----------------------------------------------------------------------------------
struct HostGraph { std::vector<i64> rowptr; std::vector<int> col; };

HostGraph download_graph(const SellPattern& p) {
    HostGraph g;
    g.rowptr.resize(p.n_rows + 1);
    std::vector<i64> so(p.n_slices + 1);
    std::vector<int> scol(p.n_slots);
    GL_CUDA(cudaMemcpy(g.rowptr.data(), p.rowptr, sizeof(i64) * (p.n_rows + 1), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(so.data(), p.slice_off, sizeof(i64) * (p.n_slices + 1), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(scol.data(), p.col, sizeof(int) * p.n_slots, cudaMemcpyDeviceToHost));
    g.col.resize(p.nnzb);
    for (i64 r = 0; r < p.n_rows; ++r)
        for (i64 t = g.rowptr[r]; t < g.rowptr[r + 1]; ++t)
            g.col[t] = scol[so[r >> 5] + (t - g.rowptr[r]) * 32 + (r & 31)];
    return g;
}

// Greedy (Vanek-style) aggregation; excluded[i] nodes stay out (-1). Returns number of aggregates.
int aggregate(const HostGraph& g, int n, const std::vector<char>& excluded, std::vector<int>& agg) {
    agg.assign(n, -1);
    int na = 0;
    auto nbr_ok = [&](int j) { return j < n && !excluded[j]; };
    // pass 1: a node whose whole (eligible) neighbourhood is free becomes a root
    for (int i = 0; i < n; ++i) {
        if (excluded[i] || agg[i] >= 0) continue;
        bool all_free = true;
        int cnt = 0;
        for (i64 t = g.rowptr[i]; t < g.rowptr[i + 1]; ++t) {
            int j = g.col[t];
            if (j == i || !nbr_ok(j)) continue;
            ++cnt;
            if (agg[j] >= 0) { all_free = false; break; }
        }
        if (!all_free || cnt == 0) continue;
        agg[i] = na;
        for (i64 t = g.rowptr[i]; t < g.rowptr[i + 1]; ++t) { int j = g.col[t]; if (j != i && nbr_ok(j)) agg[j] = na; }
        ++na;
    }
    // pass 2: attach leftovers to a neighbouring aggregate (as decided after pass 1)
    std::vector<int> agg1 = agg;
    for (int i = 0; i < n; ++i) {
        if (excluded[i] || agg1[i] >= 0) continue;
        for (i64 t = g.rowptr[i]; t < g.rowptr[i + 1]; ++t) {
            int j = g.col[t];
            if (j != i && nbr_ok(j) && agg1[j] >= 0) { agg[i] = agg1[j]; break; }
        }
    }
    // pass 3: remaining nodes form aggregates with their remaining neighbours
    for (int i = 0; i < n; ++i) {
        if (excluded[i] || agg[i] >= 0) continue;
        agg[i] = na;
        for (i64 t = g.rowptr[i]; t < g.rowptr[i + 1]; ++t) { int j = g.col[t]; if (j != i && nbr_ok(j) && agg[j] < 0) agg[j] = na; }
        ++na;
    }
    return na;
}

// Renumber the aggregates so that coarse rows of similar length share a SELL slice: within windows of `window` consecutive
// aggregates (the sweep order of `aggregate` is spatially coherent, so a window keeps the gather locality) sort by the number of
// distinct neighbouring aggregates.  At C4 the level-1 operator carried 23 % padding slots (1.30 M blocks in 1.59 M slots), all
// of them streamed by four passes per V-cycle.  GLIMS_AMG_SORT=<window> enables it (default off, see the call site).
void relabel_by_degree(const HostGraph& g, int n, std::vector<int>& agg, int na, int window) {
    if (na <= 32) return;
    std::vector<int> mptr(na + 1, 0), midx;
    for (int i = 0; i < n; ++i) if (agg[i] >= 0) mptr[agg[i] + 1]++;
    for (int I = 0; I < na; ++I) mptr[I + 1] += mptr[I];
    midx.resize(mptr[na]);
    { std::vector<int> pos(mptr.begin(), mptr.end() - 1); for (int i = 0; i < n; ++i) if (agg[i] >= 0) midx[pos[agg[i]]++] = i; }
    std::vector<int> deg(na, 0), mark(na, -1);
    for (int I = 0; I < na; ++I)
        for (int t = mptr[I]; t < mptr[I + 1]; ++t) {
            const int i = midx[t];
            for (i64 e = g.rowptr[i]; e < g.rowptr[i + 1]; ++e) {
                const int j = g.col[e];
                if (j >= n || agg[j] < 0) continue;
                if (mark[agg[j]] != I) { mark[agg[j]] = I; deg[I]++; }
            }
        }
    std::vector<int> order(na), newid(na);
    for (int I = 0; I < na; ++I) order[I] = I;
    for (int w0 = 0; w0 < na; w0 += window) {
        const int w1 = std::min(na, w0 + window);
        std::stable_sort(order.begin() + w0, order.begin() + w1, [&](int a, int b) { return deg[a] > deg[b]; });
    }
    for (int k = 0; k < na; ++k) newid[order[k]] = k;
    for (int i = 0; i < n; ++i) if (agg[i] >= 0) agg[i] = newid[agg[i]];
}

void alloc_work(Level& l) {
    i64 n = std::max<i64>(l.n_cols, l.n) * l.bs;
    for (double** v : {&l.x, &l.b, &l.r, &l.d}) {
        GL_CUDA(cudaMalloc(v, sizeof(double) * (n > 0 ? n : 1)));
        GL_CUDA(cudaMemset(*v, 0, sizeof(double) * (n > 0 ? n : 1)));
    }
}

void diag_inverse(glims_ctx* c, Level& l, double reg) {
    GL_CUDA(cudaMalloc(&l.dinv, sizeof(double) * (i64)l.n * l.bs * l.bs));
    int g = nblk(l.n);
    if (l.bs == 2) k_block_diag_inverse<2><<<g, TPB, 0, c->stream>>>(l.pat.diag, l.n, l.A, l.dinv, reg);
    else if (l.bs == 3) k_block_diag_inverse<3><<<g, TPB, 0, c->stream>>>(l.pat.diag, l.n, l.A, l.dinv, reg);
    else k_block_diag_inverse<6><<<g, TPB, 0, c->stream>>>(l.pat.diag, l.n, l.A, l.dinv, reg);
}

void estimate_lmax(glims_ctx* c, Level& l) {
    i64 n = (i64)l.n * l.bs;
    k_fill_pseudo_random<<<nblk(n), TPB, 0, c->stream>>>(l.x, n);
    double lam = 1.0;
    for (int it = 0; it < 15; ++it) {
        launch_spmv_generic(c, l.pat, l.A, l.bs, l.x, l.r);
        launch_block_jacobi(c, l.dinv, l.bs, l.r, l.d, l.n, -1);
        launch_dot(c, l.d, l.d, n, S_TMP0);
        launch_dot(c, l.x, l.x, n, S_TMP1);
        double v[2];
        read_scalars(c, S_TMP0, 2, v);
        lam = std::sqrt(v[0] / (v[1] > 0 ? v[1] : 1.0));
        launch_copy(c, l.d, l.x, n);
        launch_scale(c, 1.0 / std::sqrt(v[0] > 0 ? v[0] : 1.0), l.x, n);
    }
    l.lmax = 1.1 * lam;
    launch_zero(c, l.x, std::max<i64>(l.n_cols, l.n) * l.bs);
    launch_zero(c, l.d, std::max<i64>(l.n_cols, l.n) * l.bs);
    launch_zero(c, l.r, std::max<i64>(l.n_cols, l.n) * l.bs);
}

template <int BS>
void cheb_update(glims_ctx* c, Level& l, double* x, double c1, double c2) {
    int g = (int)std::min<i64>((l.n + TPB - 1) / TPB, 148 * 8);
    k_cheb_update<BS><<<g > 0 ? g : 1, TPB, 0, c->stream>>>(l.dinv, l.r, l.d, x, l.n, c1, c2);
    c->launches++;
}

// x <- x + p(D^-1 A) D^-1 (b - A x), Chebyshev polynomial of the given degree on [ratio*lmax, lmax]
void smooth(glims_ctx* c, Amg* amg, Level& l, const double* b, double* x, bool zero_guess, bool fine) {
    const double lmax = l.lmax, lmin = amg->cheb_ratio * lmax;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    for (int k = 0; k < amg->cheb_degree; ++k) {
        if (k == 0 && zero_guess) {
            GL_CUDA(cudaMemcpyAsync(l.r, b, sizeof(double) * (i64)l.n * l.bs, cudaMemcpyDeviceToDevice, c->stream));
        } else {
            if (fine) halo_exchange(c, x, l.bs);      // partitioned mesh: the fine-level smoother is the global one
            launch_spmv_generic(c, l.pat, l.A, l.bs, x, l.r, b);
        }
        double c1, c2;
        if (k == 0) { c1 = 0.0; c2 = 1.0 / theta; }
        else {
            double rho_new = 1.0 / (2.0 * sigma - rho);
            c1 = rho_new * rho;
            c2 = 2.0 * rho_new / delta;
            rho = rho_new;
        }
        if (l.bs == 2) cheb_update<2>(c, l, x, c1, c2);
        else if (l.bs == 3) cheb_update<3>(c, l, x, c1, c2);
        else cheb_update<6>(c, l, x, c1, c2);
    }
}

void vcycle(glims_ctx* c, Amg* amg, int li, const double* b, double* x) {
    Level& l = amg->L[li];
    const int D = amg->dim;
    if (li == (int)amg->L.size() - 1) {
        int m = amg->coarse_m;
        k_dense_matvec<<<(m + 7) / 8, 256, 0, c->stream>>>(amg->coarse_inv, b, x, m);
        c->launches++;
        return;
    }
    Level& lc = amg->L[li + 1];
    launch_zero(c, x, (i64)l.n * l.bs);
    const bool l0 = (li == 0);
    smooth(c, amg, l, b, x, true, l0);
    if (l0) halo_exchange(c, x, l.bs);
    launch_spmv_generic(c, l.pat, l.A, l.bs, x, l.r, b);
    if (D == 2) k_restrict<2><<<nblk(l.nc), TPB, 0, c->stream>>>(l.nc, l.mem_ptr, l.mem_idx, l0, l.rvec, l.free_mask, l.r, lc.b);
    else k_restrict<3><<<nblk(l.nc), TPB, 0, c->stream>>>(l.nc, l.mem_ptr, l.mem_idx, l0, l.rvec, l.free_mask, l.r, lc.b);
    c->launches++;
    vcycle(c, amg, li + 1, lc.b, lc.x);
    if (D == 2) k_prolong_add<2><<<nblk(l.n), TPB, 0, c->stream>>>(l.n, l.agg, l0, l.rvec, l.free_mask, lc.x, x);
    else k_prolong_add<3><<<nblk(l.n), TPB, 0, c->stream>>>(l.n, l.agg, l0, l.rvec, l.free_mask, lc.x, x);
    c->launches++;
    smooth(c, amg, l, b, x, false, l0);
}


// ================================================================================================
// Mixed-precision V-cycle: matrices, block-diagonal inverses and level vectors in FP32 (half the bytes of
// the bandwidth-bound smoother), applied as a fixed linear preconditioner inside the FP64 PCG.
__global__ void k_to_float(const double* __restrict__ a, float* __restrict__ b, i64 n) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) b[i] = (float)a[i];
}
__global__ void k_to_half(const double* __restrict__ a, __half* __restrict__ b, i64 n, double scale) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) b[i] = __double2half(a[i] * scale);
}
struct AbsD { __host__ __device__ double operator()(double v) const { return v < 0 ? -v : v; } };
__global__ void k_to_double(const float* __restrict__ a, double* __restrict__ b, i64 n) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) b[i] = (double)a[i];
}

__device__ inline float ld_mat(const float* p) { return __ldcs(p); }
__device__ inline float ld_mat(const __half* p) { return __half2float(__ldcs(p)); }

// Fine-level FP16 matrix, slot-pair layout.  Slots 2q and 2q+1 of a slice are interleaved as __half2 -- entry k of both blocks
// for the 32 rows of the slice is one 128-byte line, [q][k][lane] -- and an odd last slot keeps the scalar [k][lane] form, so the
// array has exactly the size and slice offsets of the scalar layout.  A block pair then costs BS*BS 4-byte loads per lane
// instead of 2*BS*BS 2-byte ones (64-byte warp requests), and the column indices of the next pair are fetched before the
// gathers of the current one are consumed: twice the bytes in flight per warp for a kernel that waits on the long scoreboard
// 64 cycles per issue (profiles/r02_ncu_summary.md).
template <int BSQ>
__global__ void k_to_half_paired(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, int n_slices,
                                 const double* __restrict__ A, __half* __restrict__ A16, double scale) {
    for (int S = blockIdx.x; S < n_slices; S += gridDim.x) {
        const i64 base = slice_off[S] * BSQ;
        const int w = slice_w[S], np = w >> 1, n = w * BSQ * 32;
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const int j = t / (BSQ * 32), rem = t - j * (BSQ * 32);
            const i64 dst = j < 2 * np ? base + (i64)(j >> 1) * (2 * BSQ * 32) + rem * 2 + (j & 1) : base + t;
            A16[dst] = __double2half(A[base + t] * scale);
        }
    }
}

// matrix stream: read once, keep it out of L1 (which then holds the gathered x blocks)
__device__ inline __half2 ld_stream_h2(const __half2* p) {
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return *reinterpret_cast<__half2*>(&v);
}
template <int BS>
__device__ inline void row_dot(const i64 base, const int w, const int* __restrict__ col, const __half* __restrict__ A,
                               const float* __restrict__ x, const int lane, float (&acc)[BS]) {
    constexpr int BSQ = BS * BS;
    const int np = w >> 1;
    const __half2* Ap = reinterpret_cast<const __half2*>(A + base * BSQ) + lane;
    const int* cp = col + base + lane;
    // software pipeline: the matrix pair and the column indices of step q+1 are requested before the gathers of step q are used
    int c0 = 0, c1 = 0;
    __half2 an[BSQ];
    if (np > 0) {
        c0 = __ldg(cp); c1 = __ldg(cp + 32);
#pragma unroll
        for (int k = 0; k < BSQ; ++k) an[k] = ld_stream_h2(&Ap[k * 32]);
    }
    for (int q = 0; q < np; ++q) {
        float x0[BS], x1[BS];
#pragma unroll
        for (int b = 0; b < BS; ++b) { x0[b] = __ldg(&x[(i64)c0 * BS + b]); x1[b] = __ldg(&x[(i64)c1 * BS + b]); }
        __half2 a[BSQ];
#pragma unroll
        for (int k = 0; k < BSQ; ++k) a[k] = an[k];
        if (q + 1 < np) {
            c0 = __ldg(cp + (2 * q + 2) * 32); c1 = __ldg(cp + (2 * q + 3) * 32);
            const __half2* Aq = Ap + (i64)(q + 1) * (BSQ * 32);
#pragma unroll
            for (int k = 0; k < BSQ; ++k) an[k] = ld_stream_h2(&Aq[k * 32]);
        }
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int b = 0; b < BS; ++b) {
                const float2 f = __half22float2(a[i * BS + b]);
                acc[i] = fmaf(f.x, x0[b], acc[i]); acc[i] = fmaf(f.y, x1[b], acc[i]);
            }
    }
    if (w & 1) {
        const i64 g = base + (i64)(w - 1) * 32;
        const int cidx = __ldg(&col[g + lane]);
        const __half* Ag = A + g * BSQ + lane;
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int b = 0; b < BS; ++b) acc[i] = fmaf(__half2float(__ldcs(&Ag[(i * BS + b) * 32])), __ldg(&x[(i64)cidx * BS + b]), acc[i]);
    }
}
template <int BS>
__device__ inline void row_dot(const i64 base, const int w, const int* __restrict__ col, const float* __restrict__ A,
                               const float* __restrict__ x, const int lane, float (&acc)[BS]) {
    for (int j = 0; j < w; ++j) {
        const i64 g = base + (i64)j * 32;
        const int cidx = __ldg(&col[g + lane]);
        float xv[BS];
#pragma unroll
        for (int b = 0; b < BS; ++b) xv[b] = __ldg(&x[(i64)cidx * BS + b]);
        const float* Ag = A + g * (BS * BS) + lane;
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int b = 0; b < BS; ++b) acc[i] += ld_mat(&Ag[(i * BS + b) * 32]) * xv[b];
    }
}

// thread per block row (fine level, BS = 2|3): y = A x  or  y = rhs - A x.  MT = float, or __half (slot-pair layout) with the
// values scaled by a power of two (unscale restores the row sums; FP32 accumulation either way)
template <int BS, bool RESID, typename MT>
__global__ void __launch_bounds__(TPB)
k_spmv32_row(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
             const MT* __restrict__ A, const float* __restrict__ x, float* __restrict__ y, int n_rows,
             const float* __restrict__ rhs, float unscale) {
    const int n_tiles = (n_rows + TPB - 1) / TPB;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r = tile * TPB + threadIdx.x;
        const int S = r >> 5, lane = r & 31;
        if (S * 32 >= n_rows) continue;
        float acc[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) acc[i] = 0.f;
        row_dot<BS>(slice_off[S], slice_w[S], col, A, x, lane, acc);
        if (r < n_rows) {
#pragma unroll
            for (int i = 0; i < BS; ++i) y[(i64)r * BS + i] = RESID ? rhs[(i64)r * BS + i] - unscale * acc[i] : unscale * acc[i];
        }
    }
}

// One component (block-row component i) of y = A x for the slice row `lane`: four block columns in flight, so that a
// row of ~24 blocks costs ~6 dependent rounds of loads instead of ~12 (the coarse levels are latency bound).
template <int BS, typename MT>
__device__ inline float split_row_dot(const i64 base, const int w, const int* __restrict__ col, const MT* __restrict__ A,
                                      const float* __restrict__ x, const int i, const int lane) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // (a predicated whole-round form -- clamp the column, weight by zero -- was tried: on level 1, which streams 230 MB per
    // pass, the clamped columns are real loads and the kernel got 10 % slower, 70 -> 77 us; the fused kernel of the
    // latency-bound levels below uses that form, amg_fused.cuh)
    int j = 0;
    for (; j + 3 < w; j += 4) {
        int cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) cc[u] = __ldg(&col[base + (i64)(j + u) * 32 + lane]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const MT* Au = A + (base + (i64)(j + u) * 32) * (BS * BS) + (i * BS) * 32 + lane;
#pragma unroll
            for (int b = 0; b < BS; ++b) acc[u] += ld_mat(&Au[b * 32]) * __ldg(&x[(i64)cc[u] * BS + b]);
        }
    }
    for (; j < w; ++j) {
        const int c0 = __ldg(&col[base + (i64)j * 32 + lane]);
        const MT* A0 = A + (base + (i64)j * 32) * (BS * BS) + (i * BS) * 32 + lane;
#pragma unroll
        for (int b = 0; b < BS; ++b) acc[0] += ld_mat(&A0[b * 32]) * __ldg(&x[(i64)c0 * BS + b]);
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// The same product with the gathered x blocks staged in shared memory.  In the form above each of the BS component warps of a
// slice gathers all BS components of every column block itself: BS*BS scattered 4-byte gathers per slot and CTA, and on level 1
// (2 MB vector, 230 MB matrix per pass) the L1 tag stage of those gathers, not HBM, set the pace -- an FP16 matrix (half the
// bytes) and a degree-sorted numbering (23 % fewer slots) both left the pass at 75-82 us.  Here warp i fetches the column index
// and the whole x block of slots i, i+BS, ... once (8-byte loads when the block is 24 bytes) into xs[slot][component][lane];
// after one barrier every warp streams its BS matrix values per slot against shared memory: the scattered work drops by BS to
// 2 BS, and the matrix loads no longer wait behind a column -> gather chain.
constexpr int SPLIT_CH = 24;      // slots staged per round (level-1 slices are <= 23 wide at C4)
template <int BS, typename MT>
__device__ inline float split_row_dot_staged(float (*xs)[BS][32], const i64 base, const int w, const int* __restrict__ col,
                                             const MT* __restrict__ A, const float* __restrict__ x, const int i, const int lane) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 < w; j0 += SPLIT_CH) {
        const int nj = min(SPLIT_CH, w - j0);
        if (j0 > 0) __syncthreads();
#pragma unroll 4
        for (int j = i; j < nj; j += BS) {
            const int cidx = __ldg(&col[base + (i64)(j0 + j) * 32 + lane]);
            if constexpr (BS % 2 == 0) {
                const float2* xp = reinterpret_cast<const float2*>(x + (i64)cidx * BS);
#pragma unroll
                for (int b = 0; b < BS / 2; ++b) { const float2 v = __ldg(&xp[b]); xs[j][2 * b][lane] = v.x; xs[j][2 * b + 1][lane] = v.y; }
            } else {
#pragma unroll
                for (int b = 0; b < BS; ++b) xs[j][b][lane] = __ldg(&x[(i64)cidx * BS + b]);
            }
        }
        __syncthreads();
        const MT* Ab = A + (base + (i64)j0 * 32) * (BS * BS) + (i * BS) * 32 + lane;
        int j = 0;
        for (; j + 3 < nj; j += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const MT* Au = Ab + (i64)(j + u) * 32 * (BS * BS);
#pragma unroll
                for (int b = 0; b < BS; ++b) acc[u] += ld_mat(&Au[b * 32]) * xs[j + u][b][lane];
            }
        }
        for (; j < nj; ++j) {
            const MT* A0 = Ab + (i64)j * 32 * (BS * BS);
#pragma unroll
            for (int b = 0; b < BS; ++b) acc[0] += ld_mat(&A0[b * 32]) * xs[j][b][lane];
        }
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// CTA per slice, warp per block-row component (coarse levels, BS = 3|6)
template <int BS, bool RESID, typename MT, bool STAGED>
__global__ void __launch_bounds__(32 * BS)
k_spmv32_split(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
               const MT* __restrict__ A, const float* __restrict__ x, float* __restrict__ y, int n_rows,
               int slice0, int n_slices, const float* __restrict__ rhs, float unscale) {
    __shared__ float xs[STAGED ? SPLIT_CH : 1][BS][32];
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
    for (int S = slice0 + blockIdx.x; S < slice0 + n_slices; S += gridDim.x) {
        const int r = S * 32 + lane;
        const i64 base = slice_off[S];
        const int w = slice_w[S];
        const float dotv = STAGED ? split_row_dot_staged<BS, MT>(xs, base, w, col, A, x, i, lane) : split_row_dot<BS, MT>(base, w, col, A, x, i, lane);
        if (r < n_rows) {
            const float v = unscale * dotv;
            y[(i64)r * BS + i] = RESID ? rhs[(i64)r * BS + i] - v : v;
        }
        if (STAGED) __syncthreads();      // xs is free for the next slice
    }
}

// ---- fused smoother steps (FP32 V-cycle) ------------------------------------------------------------
// First Chebyshev step from a zero guess: x = d = c2 * Dinv * b  (replaces memset + copy + update).
template <int BS>
__global__ void k_cheb_first32(const float* __restrict__ dinv, const float* __restrict__ b, float* __restrict__ d,
                               float* __restrict__ x, int n, float c2) {
    for (i64 row = blockIdx.x * (i64)TPB + threadIdx.x; row < n; row += (i64)gridDim.x * TPB) {
        float rv[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) rv[i] = b[row * BS + i];
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            float z = 0.f;
#pragma unroll
            for (int j = 0; j < BS; ++j) z += dinv[row * BS * BS + i * BS + j] * rv[j];
            const float dn = c2 * z;
            d[row * BS + i] = dn;
            x[row * BS + i] = dn;
        }
    }
}

// One Chebyshev step in one pass over the matrix: r = b - A x (row-local), d = c1 d + c2 Dinv r, xn = x + d.
// The new iterate goes to a second buffer because other rows still gather the old one.  Thread per block row.
template <int BS, typename MT>
__global__ void __launch_bounds__(TPB)
k_spmv32_row_cheb(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
                  const MT* __restrict__ A, const float* __restrict__ x, float* __restrict__ xn, int n_rows,
                  const float* __restrict__ rhs, const float* __restrict__ dinv, float* __restrict__ d, float c1, float c2,
                  float unscale) {
    const int n_tiles = (n_rows + TPB - 1) / TPB;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r = tile * TPB + threadIdx.x;
        const int S = r >> 5, lane = r & 31;
        if (S * 32 >= n_rows) continue;
        float acc[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) acc[i] = 0.f;
        row_dot<BS>(slice_off[S], slice_w[S], col, A, x, lane, acc);
        if (r < n_rows) {
            float rv[BS];
#pragma unroll
            for (int i = 0; i < BS; ++i) rv[i] = rhs[(i64)r * BS + i] - unscale * acc[i];
#pragma unroll
            for (int i = 0; i < BS; ++i) {
                float z = 0.f;
#pragma unroll
                for (int j = 0; j < BS; ++j) z += dinv[(i64)r * BS * BS + i * BS + j] * rv[j];
                const float dn = c2 * z + (c1 != 0.f ? c1 * d[(i64)r * BS + i] : 0.f);
                d[(i64)r * BS + i] = dn;
                xn[(i64)r * BS + i] = x[(i64)r * BS + i] + dn;
            }
        }
    }
}

// Same step for the coarse levels: CTA per slice, warp i computes component i of the slice's 32 block rows; the residual
// block of a row is exchanged through shared memory so that every thread can apply its row of Dinv.
template <int BS, typename MT, bool STAGED>
__global__ void __launch_bounds__(32 * BS)
k_spmv32_split_cheb(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
                    const MT* __restrict__ A, const float* __restrict__ x, float* __restrict__ xn, int n_rows,
                    int slice0, int n_slices, const float* __restrict__ rhs, const float* __restrict__ dinv, float* __restrict__ d,
                    float c1, float c2, float unscale) {
    __shared__ float rs[BS][32];
    __shared__ float xs[STAGED ? SPLIT_CH : 1][BS][32];
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
    for (int S = slice0 + blockIdx.x; S < slice0 + n_slices; S += gridDim.x) {
        const int r = S * 32 + lane;
        const i64 base = slice_off[S];
        const int w = slice_w[S];
        const float dotv = STAGED ? split_row_dot_staged<BS, MT>(xs, base, w, col, A, x, i, lane) : split_row_dot<BS, MT>(base, w, col, A, x, i, lane);
        rs[i][lane] = r < n_rows ? rhs[(i64)r * BS + i] - unscale * dotv : 0.f;
        __syncthreads();
        if (r < n_rows) {
            float z = 0.f;
#pragma unroll
            for (int jj = 0; jj < BS; ++jj) z += dinv[(i64)r * BS * BS + i * BS + jj] * rs[jj][lane];
            const float dn = c2 * z + (c1 != 0.f ? c1 * d[(i64)r * BS + i] : 0.f);
            d[(i64)r * BS + i] = dn;
            xn[(i64)r * BS + i] = x[(i64)r * BS + i] + dn;
        }
        __syncthreads();
    }
}

// ---- coarse-level products fed by the bulk-copy engine ---------------------------------------------------------------------
// The staged kernels above are bound by their chain of dependent loads: a slice is slice_off -> column -> gather -> w/4 rounds of
// matrix loads -> epilogue operands, about ten memory latencies, and level 1 of C4 is only 2625 slices -- two waves, each CTA
// idle at the memory system for most of its life (45 % of HBM bandwidth, 62 cycles of long-scoreboard stall per issue).  The
// matrix of a slice is one contiguous run (w slots x BS*BS*128 B), so here thread 0 hands it to the bulk-copy engine in chunks
// of TMA_SL slots through a TMA_NST-deep ring of shared-memory stages (mbarrier complete_tx), the warps gather the x blocks
// meanwhile, and the products read shared memory only.  Persistent CTAs (grid = resident CTAs); ~74 KB of matrix in flight per
// CTA independent of what the warps wait for.
template <int BS, int TMA_SL, int TMA_NST, int CH>
constexpr size_t split_tma_smem() { return (size_t)TMA_NST * TMA_SL * BS * BS * 128 + (size_t)CH * BS * 128 + (size_t)BS * 128; }

// MODE 0: y = A x, 1: y = rhs - A x, 2: Chebyshev step (xn = x + d_new, d_new = c1 d + c2 Dinv (rhs - A x))
template <int BS, int MODE, int TMA_SL, int TMA_NST, int CH>
__global__ void __launch_bounds__(32 * BS)
k_split_tma(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
            const float* __restrict__ A, const float* __restrict__ x, float* __restrict__ out, int n_rows, int slice0, int n_slices,
            const float* __restrict__ rhs, const float* __restrict__ dinv, float* __restrict__ d, float c1, float c2) {
    constexpr int BSQ = BS * BS;
    constexpr unsigned SLOT_BYTES = BSQ * 128;
    extern __shared__ __align__(128) unsigned char smem[];
    float* As = reinterpret_cast<float*>(smem);                                                  // [NST][SL][BSQ][32]
    float (*xs)[BS][32] = reinterpret_cast<float (*)[BS][32]>(smem + (size_t)TMA_NST * TMA_SL * SLOT_BYTES);      // [CH][BS][32]
    float (*rs)[32] = reinterpret_cast<float (*)[32]>(smem + (size_t)TMA_NST * TMA_SL * SLOT_BYTES + (size_t)CH * BS * 128);
    __shared__ unsigned long long full[TMA_NST];
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        for (int st = 0; st < TMA_NST; ++st) mbar_init(&full[st], 1);
    __syncthreads();
    unsigned kc = 0;          // chunks consumed so far by this CTA: stage = kc % NST, parity = (kc / NST) & 1
    const int S_end = slice0 + n_slices;
    // producer cursor (thread 0): the CTA's chunks form one sequence over all of its slices; chunk number k lives in stage k % NST
    // and is requested as soon as chunk k - NST has been consumed, across slice boundaries
    int pS = slice0 + blockIdx.x, pq = 0, p_w = 0, p_nch = 0;
    i64 p_base = 0;
    unsigned pk = 0;
    int S = slice0 + blockIdx.x, nx_S = S, nx_w = S < S_end ? slice_w[S] : 0;      // header of the slice after the one being consumed
    i64 nx_base = S < S_end ? slice_off[S] : 0;
    auto advance = [&]() {
        if (pS >= S_end) return;
        const unsigned st = pk % TMA_NST;
        const unsigned bytes = (unsigned)min(TMA_SL, p_w - pq * TMA_SL) * SLOT_BYTES;
        mbar_expect_tx(&full[st], bytes);
        bulk_g2s(As + (size_t)st * TMA_SL * BSQ * 32, A + (p_base + (i64)pq * TMA_SL * 32) * BSQ, bytes, &full[st]);
        ++pk;
        if (++pq == p_nch) {
            pS += gridDim.x; pq = 0;
            while (pS < S_end) {          // skip empty slices
                if (pS == nx_S) { p_base = nx_base; p_w = nx_w; }        // normally the slice whose header is already here
                else { p_base = slice_off[pS]; p_w = slice_w[pS]; }
                p_nch = (p_w + TMA_SL - 1) / TMA_SL;
                if (p_nch > 0) break;
                pS += gridDim.x;
            }
        }
    };
    if (threadIdx.x == 0) {
        while (pS < S_end) {
            if (pS == nx_S) { p_base = nx_base; p_w = nx_w; } else { p_base = slice_off[pS]; p_w = slice_w[pS]; }
            p_nch = (p_w + TMA_SL - 1) / TMA_SL;
            if (p_nch > 0) break;
            pS += gridDim.x;
        }
        for (int q = 0; q < TMA_NST; ++q) advance();
    }
    for (; S < S_end; S += gridDim.x) {
        const int r = S * 32 + lane;
        const i64 base = nx_base;
        const int w = nx_w;
        nx_S = S + (int)gridDim.x;
        if (nx_S < S_end) { nx_base = slice_off[nx_S]; nx_w = slice_w[nx_S]; }      // next slice's header, early
        const int nch = (w + TMA_SL - 1) / TMA_SL;
        // epilogue operands do not depend on the product: request them now
        float rhs_v = 0.f, d_v = 0.f, x_v = 0.f;
        if (MODE >= 1 && r < n_rows) rhs_v = rhs[(i64)r * BS + i];
        if (MODE == 2 && r < n_rows) { d_v = c1 != 0.f ? d[(i64)r * BS + i] : 0.f; x_v = x[(i64)r * BS + i]; }
        float acc[TMA_SL];
#pragma unroll
        for (int u = 0; u < TMA_SL; ++u) acc[u] = 0.f;
        for (int q = 0; q < nch; ++q) {
            const int j0 = q * TMA_SL, jx = j0 % CH;
            if (jx == 0) {            // (re)fill xs: warp i stages the x blocks of slots j0 + i, j0 + i + BS, ...
                const int nj = min(CH, w - j0);
#pragma unroll 4
                for (int j = i; j < nj; j += BS) {
                    const int cidx = __ldg(&col[base + (i64)(j0 + j) * 32 + lane]);
                    if constexpr (BS % 2 == 0) {
                        const float2* xp = reinterpret_cast<const float2*>(x + (i64)cidx * BS);
#pragma unroll
                        for (int b = 0; b < BS / 2; ++b) { const float2 v = __ldg(&xp[b]); xs[j][2 * b][lane] = v.x; xs[j][2 * b + 1][lane] = v.y; }
                    } else {
#pragma unroll
                        for (int b = 0; b < BS; ++b) xs[j][b][lane] = __ldg(&x[(i64)cidx * BS + b]);
                    }
                }
                __syncthreads();
            }
            const unsigned st = (kc + q) % TMA_NST;
            mbar_wait(&full[st], ((kc + q) / TMA_NST) & 1u);
            const int nsl = min(TMA_SL, w - j0);
            const float* Ast = As + (size_t)st * TMA_SL * BSQ * 32 + (i * BS) * 32 + lane;
#pragma unroll
            for (int u = 0; u < TMA_SL; ++u)
                if (u < nsl) {
#pragma unroll
                    for (int b = 0; b < BS; ++b) acc[u] += Ast[u * BSQ * 32 + b * 32] * xs[jx + u][b][lane];
                }
            __syncthreads();          // every warp is done with stage st (and, at the end of an xs round, with xs)
            if (threadIdx.x == 0) advance();
        }
        kc += nch;
        float dotv = 0.f;
#pragma unroll
        for (int u = 0; u < TMA_SL; ++u) dotv += acc[u];
        if (MODE < 2) {
            if (r < n_rows) out[(i64)r * BS + i] = MODE == 1 ? rhs_v - dotv : dotv;
        } else {
            rs[i][lane] = rhs_v - dotv;
            __syncthreads();
            if (r < n_rows) {
                float z = 0.f;
#pragma unroll
                for (int jj = 0; jj < BS; ++jj) z += dinv[(i64)r * BSQ + i * BS + jj] * rs[jj][lane];
                const float dn = c2 * z + c1 * d_v;
                d[(i64)r * BS + i] = dn;
                out[(i64)r * BS + i] = x_v + dn;
            }
            __syncthreads();
        }
    }
}

template <int BS, int MODE, int SL, int NST, int CH>
void launch_split_tma_cfg(glims_ctx* c, const SellPattern& p, const float* A, const float* x, float* out, int s0, int ns,
                          const float* rhs, const float* dinv, float* d, float c1, float c2) {
    constexpr size_t smem = split_tma_smem<BS, SL, NST, CH>();
    static int grid_caps[64] = {0};          // per device: the shared-memory attribute is a per-device property of the function
    int dev = 0;
    GL_CUDA(cudaGetDevice(&dev));
    int& grid_cap = grid_caps[dev & 63];
    if (!grid_cap) {
        GL_CUDA(cudaFuncSetAttribute(k_split_tma<BS, MODE, SL, NST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        grid_cap = resident_grid((const void*)k_split_tma<BS, MODE, SL, NST, CH>, 32 * BS, smem);
    }
    const int g = ns < grid_cap ? ns : grid_cap;
    k_split_tma<BS, MODE, SL, NST, CH><<<g, 32 * BS, smem, c->stream>>>(p.slice_off, p.slice_w, p.col, A, x, out, p.n_rows, s0, ns, rhs, dinv, d, c1, c2);
}
// GLIMS_SPLIT_TMA: 0 off, else the ring shape (slots per chunk, stages, x slots staged): 1 = 4,3,24 (75 KB per CTA), 2 = 2,4,24 (56 KB),
// 3 = 2,3,12 (38 KB, five CTAs per SM; default).  One level-1 smoother step at C4 / at an eighth of C4, L2 flushed: register-staged
// kernel 75 / 37 us, shared-memory-staged x 76 / 28 us, bulk-copy ring 1: 73 / 16.7, 2: 69 / 16.4, 3: 63 / 17.8 us.
template <int BS, int MODE>
bool launch_split_tma(glims_ctx* c, const SellPattern& p, const float* A, const float* x, float* out, int s0, int ns,
                      const float* rhs, const float* dinv, float* d, float c1, float c2) {
    static const int cfg = [] { const char* e = std::getenv("GLIMS_SPLIT_TMA"); return e ? atoi(e) : 3; }();
    if (!cfg || ns <= 0) return false;
    if (cfg == 2) launch_split_tma_cfg<BS, MODE, 2, 4, 24>(c, p, A, x, out, s0, ns, rhs, dinv, d, c1, c2);
    else if (cfg == 3) launch_split_tma_cfg<BS, MODE, 2, 3, 12>(c, p, A, x, out, s0, ns, rhs, dinv, d, c1, c2);
    else launch_split_tma_cfg<BS, MODE, 4, 3, 24>(c, p, A, x, out, s0, ns, rhs, dinv, d, c1, c2);
    return true;
}

// thread per aggregate: better for the large fine-level restriction (tens of thousands of aggregates)
template <int D>
__global__ void k_restrict32_serial(int nc, const int* __restrict__ mem_ptr, const int* __restrict__ mem_idx, bool level0,
                             const double* __restrict__ rvec, const unsigned char* __restrict__ free_mask,
                             const float* __restrict__ rf, float* __restrict__ rc) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc) return;
    constexpr int NR = (D == 2) ? 1 : 3;
    const int bsc = D + NR, bsf = level0 ? D : bsc;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int m = mem_ptr[I]; m < mem_ptr[I + 1]; ++m) {
        const int i = mem_idx[m];
        double P[6][6];
        int a_, b_;
        build_P<D>(level0, rvec + (i64)i * D, level0 ? free_mask[i] : 0xffu, P, a_, b_);
        for (int k = 0; k < bsf; ++k) {
            double v = rf[(i64)i * bsf + k];
            for (int j = 0; j < bsc; ++j) acc[j] += P[k][j] * v;
        }
    }
    for (int j = 0; j < bsc; ++j) rc[(i64)I * bsc + j] = (float)acc[j];
}

// r_c = P^T r_f.  One warp per aggregate, lanes over its members (about twenty), shuffle reduction: the aggregate's
// members are visited in one round instead of a serial loop -- these kernels are pure latency on the small levels.
template <int D>
__global__ void __launch_bounds__(TPB)
k_restrict32(int nc, const int* __restrict__ mem_ptr, const int* __restrict__ mem_idx, bool level0,
             const double* __restrict__ rvec, const unsigned char* __restrict__ free_mask,
             const float* __restrict__ rf, float* __restrict__ rc) {
    const int I = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (I >= nc) return;
    constexpr int NR = (D == 2) ? 1 : 3;
    const int bsc = D + NR, bsf = level0 ? D : bsc;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int m = mem_ptr[I] + lane; m < mem_ptr[I + 1]; m += 32) {
        const int i = mem_idx[m];
        double P[6][6];
        int a_, b_;
        build_P<D>(level0, rvec + (i64)i * D, level0 ? free_mask[i] : 0xffu, P, a_, b_);
        for (int k = 0; k < bsf; ++k) {
            double v = rf[(i64)i * bsf + k];
            for (int j = 0; j < bsc; ++j) acc[j] += P[k][j] * v;
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
#pragma unroll
    for (int j = 0; j < 6; ++j)
        if (lane == j && j < bsc) rc[(i64)I * bsc + j] = (float)acc[j];
}

template <int D>
__global__ void k_prolong_add32(int n, const int* __restrict__ agg, bool level0, const double* __restrict__ rvec,
                                const unsigned char* __restrict__ free_mask, const float* __restrict__ xc,
                                float* __restrict__ xf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int I = agg[i];
    if (I < 0) return;
    constexpr int NR = (D == 2) ? 1 : 3;
    const int bsc = D + NR, bsf = level0 ? D : bsc;
    double P[6][6];
    int a_, b_;
    build_P<D>(level0, rvec + (i64)i * D, level0 ? free_mask[i] : 0xffu, P, a_, b_);
    for (int k = 0; k < bsf; ++k) {
        double v = 0;
        for (int j = 0; j < bsc; ++j) v += P[k][j] * (double)xc[(i64)I * bsc + j];
        xf[(i64)i * bsf + k] += (float)v;
    }
}

__global__ void k_dense_matvec32(const float* __restrict__ M, const float* __restrict__ x, float* __restrict__ y, int m) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= m) return;
    float s = 0.f;
    for (int j = lane; j < m; j += 32) s += M[(i64)row * m + j] * x[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
}

inline bool split_staged() {      // GLIMS_SPLIT_STAGED=0|1: x blocks of the coarse-level products staged in shared memory
    static const int v = [] { const char* e = std::getenv("GLIMS_SPLIT_STAGED"); return e ? atoi(e) : 1; }();
    return v != 0;
}
inline int sgrid(i64 n) { i64 g = (n + TPB - 1) / TPB; return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g)); }

void spmv32(glims_ctx* c, Level& l, const float* x, float* y, const float* rhs) {
    const SellPattern& p = l.pat;
    if (l.owns_A) {      // coarse levels (6x6 blocks in 3D, 3x3 in 2D): few rows, use the split kernel
        const int s0 = l.dist_rows ? l.row0 / 32 : 0, ns = l.dist_rows ? l.rows / 32 : p.n_slices;     // own rows only
        int g = ns < 148 * 16 ? ns : 148 * 16;
        if (g < 1) g = 1;
#define SPLIT_GO(BS, RES, MT, AP, US) do { if (split_staged()) k_spmv32_split<BS, RES, MT, true><<<g, 32 * BS, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, AP, x, y, p.n_rows, s0, ns, rhs, US); \
                                             else k_spmv32_split<BS, RES, MT, false><<<g, 32 * BS, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, AP, x, y, p.n_rows, s0, ns, rhs, US); } while (0)
        if (l.A16) {
            if (l.bs == 6) { if (rhs) SPLIT_GO(6, true, __half, l.A16, l.a16_unscale); else SPLIT_GO(6, false, __half, l.A16, l.a16_unscale); }
            else { if (rhs) SPLIT_GO(3, true, __half, l.A16, l.a16_unscale); else SPLIT_GO(3, false, __half, l.A16, l.a16_unscale); }
        } else {
            bool done;
            if (l.bs == 6) done = rhs ? launch_split_tma<6, 1>(c, p, l.A32, x, y, s0, ns, rhs, nullptr, nullptr, 0.f, 0.f) : launch_split_tma<6, 0>(c, p, l.A32, x, y, s0, ns, nullptr, nullptr, nullptr, 0.f, 0.f);
            else done = rhs ? launch_split_tma<3, 1>(c, p, l.A32, x, y, s0, ns, rhs, nullptr, nullptr, 0.f, 0.f) : launch_split_tma<3, 0>(c, p, l.A32, x, y, s0, ns, nullptr, nullptr, nullptr, 0.f, 0.f);
            if (!done) {
                if (l.bs == 6) { if (rhs) SPLIT_GO(6, true, float, l.A32, 1.f); else SPLIT_GO(6, false, float, l.A32, 1.f); }
                else { if (rhs) SPLIT_GO(3, true, float, l.A32, 1.f); else SPLIT_GO(3, false, float, l.A32, 1.f); }
            }
        }
#undef SPLIT_GO
    } else {
        // one 256-row tile per CTA: the hardware scheduler balances tiles of different widths, and the grid does not depend on
        // how many CTAs the register count lets an SM hold (a capped grid of 8 per SM ran a second, thin wave at 6 resident)
        int g = (p.n_rows + TPB - 1) / TPB;
        if (l.A16) {
            const float us = l.a16_unscale;
            if (l.bs == 2) {
                if (rhs) k_spmv32_row<2, true, __half><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A16, x, y, p.n_rows, rhs, us);
                else k_spmv32_row<2, false, __half><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A16, x, y, p.n_rows, nullptr, us);
            } else {
                if (rhs) k_spmv32_row<3, true, __half><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A16, x, y, p.n_rows, rhs, us);
                else k_spmv32_row<3, false, __half><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A16, x, y, p.n_rows, nullptr, us);
            }
        } else if (l.bs == 2) {
            if (rhs) k_spmv32_row<2, true, float><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A32, x, y, p.n_rows, rhs, 1.f);
            else k_spmv32_row<2, false, float><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A32, x, y, p.n_rows, nullptr, 1.f);
        } else {
            if (rhs) k_spmv32_row<3, true, float><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A32, x, y, p.n_rows, rhs, 1.f);
            else k_spmv32_row<3, false, float><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A32, x, y, p.n_rows, nullptr, 1.f);
        }
    }
    c->launches++;
}

// one fused Chebyshev step: xn = x + d_new  with  d_new = c1 d + c2 Dinv (b - A x)
void cheb_step32(glims_ctx* c, Level& l, const float* b, const float* x, float* xn, float c1, float c2) {
    auto& p = l.pat;
    if (l.owns_A) {      // coarse levels: split kernel, like spmv32
        const int s0 = l.dist_rows ? l.row0 / 32 : 0, ns = l.dist_rows ? l.rows / 32 : p.n_slices;     // own rows only
        int g = ns < 148 * 16 ? ns : 148 * 16;
        if (g < 1) g = 1;
#define SPLITC_GO(BS, MT, AP, US) do { if (split_staged()) k_spmv32_split_cheb<BS, MT, true><<<g, 32 * BS, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, AP, x, xn, p.n_rows, s0, ns, b, l.dinv32, l.d32, c1, c2, US); \
                                        else k_spmv32_split_cheb<BS, MT, false><<<g, 32 * BS, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, AP, x, xn, p.n_rows, s0, ns, b, l.dinv32, l.d32, c1, c2, US); } while (0)
        if (l.A16) { if (l.bs == 6) SPLITC_GO(6, __half, l.A16, l.a16_unscale); else SPLITC_GO(3, __half, l.A16, l.a16_unscale); }
        else {
            const bool done = l.bs == 6 ? launch_split_tma<6, 2>(c, p, l.A32, x, xn, s0, ns, b, l.dinv32, l.d32, c1, c2)
                                        : launch_split_tma<3, 2>(c, p, l.A32, x, xn, s0, ns, b, l.dinv32, l.d32, c1, c2);
            if (!done) { if (l.bs == 6) SPLITC_GO(6, float, l.A32, 1.f); else SPLITC_GO(3, float, l.A32, 1.f); }
        }
#undef SPLITC_GO
    } else {
        // one 256-row tile per CTA: the hardware scheduler balances tiles of different widths, and the grid does not depend on
        // how many CTAs the register count lets an SM hold (a capped grid of 8 per SM ran a second, thin wave at 6 resident)
        int g = (p.n_rows + TPB - 1) / TPB;
        if (l.A16) {
            if (l.bs == 2) k_spmv32_row_cheb<2, __half><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A16, x, xn, p.n_rows, b, l.dinv32, l.d32, c1, c2, l.a16_unscale);
            else k_spmv32_row_cheb<3, __half><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A16, x, xn, p.n_rows, b, l.dinv32, l.d32, c1, c2, l.a16_unscale);
        } else if (l.bs == 2) k_spmv32_row_cheb<2, float><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A32, x, xn, p.n_rows, b, l.dinv32, l.d32, c1, c2, 1.f);
        else k_spmv32_row_cheb<3, float><<<g, TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, l.A32, x, xn, p.n_rows, b, l.dinv32, l.d32, c1, c2, 1.f);
    }
    c->launches++;
}

struct ChebCoef {                 // Chebyshev recurrence for the interval [ratio * lmax, lmax] of Dinv A
    double theta, delta, sigma, rho;
    ChebCoef(double lmax, double ratio) {
        const double lmin = ratio * lmax;
        theta = 0.5 * (lmax + lmin); delta = 0.5 * (lmax - lmin); sigma = theta / delta; rho = 1.0 / sigma;
    }
    void step(int k, double& c1, double& c2) {
        if (k == 0) { c1 = 0.0; c2 = 1.0 / theta; return; }
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        c1 = rho_new * rho; c2 = 2.0 * rho_new / delta; rho = rho_new;
    }
};

// V-cycle in FP32 with fused smoother steps.  The iterate ping-pongs between l.y32 and x; with `degree` pre- and
// post-smoothing steps there are 2*degree-1 fused steps after the first (zero-guess) one, an odd number, so starting
// in l.y32 leaves the result in x.
#include "amg_fused.cuh"

void vcycle32(glims_ctx* c, Amg* amg, int li, float* b, float* x);

// cur(level li) += P * V-cycle(levels li+1 ..)(P^T r32(level li)): the coarse-grid correction below level li
void coarse_correction32(glims_ctx* c, Amg* amg, int li, float* cur) {
    Level& l = amg->L[li];
    Level& lc = amg->L[li + 1];
    const int D = amg->dim;
    const bool l0 = (li == 0);
    if (li >= 1 && amg_fused_run(c, amg, li, cur)) {
        // levels li+1 .. coarsest ran in one persistent kernel: restriction of r32, the coarse V-cycle and the
        // prolongation into `cur` included (amg_fused.cuh)
        return;
    }
    float* rc = lc.b32 + (i64)l.nc_off * lc.bs;
    const int nca = l.nc_local;
    if (nca >= 16384) {      // many aggregates: thread per aggregate; few: warp per aggregate (latency)
        if (D == 2) k_restrict32_serial<2><<<nblk(nca), TPB, 0, c->stream>>>(nca, l.mem_ptr, l.mem_idx, l0, l.rvec, l.free_mask, l.r32, rc);
        else k_restrict32_serial<3><<<nblk(nca), TPB, 0, c->stream>>>(nca, l.mem_ptr, l.mem_idx, l0, l.rvec, l.free_mask, l.r32, rc);
    } else if (nca > 0) {
        if (D == 2) k_restrict32<2><<<nblk(nca, TPB / 32), TPB, 0, c->stream>>>(nca, l.mem_ptr, l.mem_idx, l0, l.rvec, l.free_mask, l.r32, rc);
        else k_restrict32<3><<<nblk(nca, TPB / 32), TPB, 0, c->stream>>>(nca, l.mem_ptr, l.mem_idx, l0, l.rvec, l.free_mask, l.r32, rc);
    }
    c->launches++;
    vcycle32(c, amg, li + 1, lc.b32, lc.x32);
    // level 1 of a partitioned mesh: every rank prolongs ALL nodes (its copy of `cur` is complete after the gather
    // in vcycle32 and x of level 2 is identical everywhere), so the first post-smoothing step needs no exchange
    if (D == 2) k_prolong_add32<2><<<nblk(l.n), TPB, 0, c->stream>>>(l.n, l.agg, l0, l.rvec, l.free_mask, lc.x32, cur);
    else k_prolong_add32<3><<<nblk(l.n), TPB, 0, c->stream>>>(l.n, l.agg, l0, l.rvec, l.free_mask, lc.x32, cur);
    c->launches++;
}

void vcycle32(glims_ctx* c, Amg* amg, int li, float* b, float* x) {
    Level& l = amg->L[li];
    const int D = amg->dim;
    // own row range: everything, except on level 1 of a partitioned mesh (Level::dist_rows)
    const bool dr = l.dist_rows;
    const i64 o = dr ? (i64)l.row0 * l.bs : 0;                  // first own entry of a level vector
    const int n_own = dr ? l.rows : l.n;
    const i64 seg = (i64)n_own * l.bs;
    auto gather = [&](float* v) { if (dr) allgather_f32(c, l.sym, v, seg); };
    if (li == 1 && amg_fused_run_full(c, amg, li, b, x)) return;
    if (li == (int)amg->L.size() - 1) {
        int m = amg->coarse_m;
        gather(b);
        k_dense_matvec32<<<(m + 7) / 8, 256, 0, c->stream>>>(amg->coarse_inv32, b, x, m);
        c->launches++;
        return;
    }
    Level& lc = amg->L[li + 1];
    const bool l0 = (li == 0);
    const int degree = l0 ? amg->cheb_degree : amg->coarse_degree;
    float *cur = l.y32, *oth = x;
    double c1, c2;
    {   // pre-smoothing from a zero guess
        ChebCoef cc(l.lmax, amg->cheb_ratio);
        cc.step(0, c1, c2);
        const int g = sgrid(n_own);
        if (l.bs == 2) k_cheb_first32<2><<<g, TPB, 0, c->stream>>>(l.dinv32, b, l.d32, cur, n_own, (float)c2);
        else if (l.bs == 3) k_cheb_first32<3><<<g, TPB, 0, c->stream>>>(l.dinv32 + o * l.bs, b + o, l.d32 + o, cur + o, n_own, (float)c2);
        else k_cheb_first32<6><<<g, TPB, 0, c->stream>>>(l.dinv32 + o * l.bs, b + o, l.d32 + o, cur + o, n_own, (float)c2);
        c->launches++;
        for (int k = 1; k < degree; ++k) {
            cc.step(k, c1, c2);
            if (l0) halo_exchange_f32(c, cur, l.bs);
            gather(cur);
            cheb_step32(c, l, b, cur, oth, (float)c1, (float)c2);
            std::swap(cur, oth);
        }
    }
    if (l0) halo_exchange_f32(c, cur, l.bs);
    gather(cur);
    spmv32(c, l, cur, l.r32, b);
    gather(l.r32);                    // level 1 of a partitioned mesh: what follows runs redundantly on every rank
    coarse_correction32(c, amg, li, cur);
    {   // post-smoothing
        ChebCoef cc(l.lmax, amg->cheb_ratio);
        for (int k = 0; k < degree; ++k) {
            cc.step(k, c1, c2);
            if (l0) halo_exchange_f32(c, cur, l.bs);
            if (k > 0) gather(cur);
            cheb_step32(c, l, b, cur, oth, (float)c1, (float)c2);
            std::swap(cur, oth);
        }
    }
    if (cur != x) GL_CUDA(cudaMemcpyAsync(x + o, cur + o, sizeof(float) * seg, cudaMemcpyDeviceToDevice, c->stream));
}

void build_fp32(glims_ctx* c, Amg* amg) {
    // Fine level in FP16 (GLIMS_AMG_FP16=0 keeps FP32): the smoother streams this matrix four times per V-cycle and is
    // bandwidth bound; the V-cycle is only a preconditioner (a fixed SPD operator as long as the rounded matrix is
    // symmetric, which it is: K(r,c) and K(c,r)^T round identically), the outer PCG stays FP64.
    const char* e16 = std::getenv("GLIMS_AMG_FP16");
    const bool want16 = !(e16 && atoi(e16) == 0);
    for (auto& l : amg->L) {
        i64 na = l.pat.n_slots * l.bs * l.bs, nd = (i64)l.n * l.bs * l.bs, nv = std::max<i64>(l.n_cols, l.n) * l.bs;
        // default: the fine level only (the coarse-level kernels are latency bound: measured no gain); GLIMS_AMG_FP16=2: every smoothed level
        // GLIMS_AMG_FP16: 0 none, 1 (default) level 0, 3 levels 0 and 1 (level 1 streams 230 MB per pass at C4 on one GPU;
        // the levels below it run in the fused kernel from FP32 copies), 2 every smoothed level (unfused coarse path)
        const int m16 = e16 ? atoi(e16) : 1;
        const bool fine16 = want16 && &l != &amg->L.back() && na > 0 &&
                            (&l == &amg->L[0] || m16 == 2 || (m16 == 3 && amg->L.size() > 1 && &l == &amg->L[1]));
        if (fine16) {
            thrust::device_ptr<const double> ap(l.A);
            const double amax = thrust::transform_reduce(thrust::cuda::par.on(c->stream), ap, ap + na, AbsD(), 0.0, thrust::maximum<double>());
            int ex = 0;
            if (amax > 0) { std::frexp(amax, &ex); }
            const double scale = std::ldexp(1.0, 14 - ex);       // largest magnitude lands in [2^13, 2^14)
            GL_CUDA(cudaMalloc(&l.A16, sizeof(__half) * na));
            if (l.owns_A) k_to_half<<<sgrid(na), TPB, 0, c->stream>>>(l.A, l.A16, na, scale);          // split kernels: scalar layout
            else if (l.bs == 2) k_to_half_paired<4><<<148 * 8, TPB, 0, c->stream>>>(l.pat.slice_off, l.pat.slice_w, l.pat.n_slices, l.A, l.A16, scale);
            else k_to_half_paired<9><<<148 * 8, TPB, 0, c->stream>>>(l.pat.slice_off, l.pat.slice_w, l.pat.n_slices, l.A, l.A16, scale);
            l.a16_unscale = (float)(1.0 / scale);
        } else {
            GL_CUDA(cudaMalloc(&l.A32, sizeof(float) * std::max<i64>(na, 1)));
            k_to_float<<<sgrid(na), TPB, 0, c->stream>>>(l.A, l.A32, na);
        }
        GL_CUDA(cudaMalloc(&l.dinv32, sizeof(float) * std::max<i64>(nd, 1)));
        k_to_float<<<sgrid(nd), TPB, 0, c->stream>>>(l.dinv, l.dinv32, nd);
        if (l.dist_rows) {
            // vectors that are all-gathered live in memory every rank maps (x32 | y32 | r32 | b32)
            const size_t vb = (sizeof(float) * (size_t)std::max<i64>(nv, 1) + 255) & ~(size_t)255;
            l.sym = sym_alloc(c, 4 * vb);
            unsigned char* base = (unsigned char*)sym_data(l.sym);
            l.x32 = (float*)base; l.y32 = (float*)(base + vb); l.r32 = (float*)(base + 2 * vb); l.b32 = (float*)(base + 3 * vb);
            GL_CUDA(cudaMalloc(&l.d32, sizeof(float) * std::max<i64>(nv, 1)));
            GL_CUDA(cudaMemsetAsync(l.d32, 0, sizeof(float) * std::max<i64>(nv, 1), c->stream));
        } else
        for (float** v : {&l.x32, &l.b32, &l.r32, &l.d32, &l.y32}) {
            GL_CUDA(cudaMalloc(v, sizeof(float) * std::max<i64>(nv, 1)));
            GL_CUDA(cudaMemsetAsync(*v, 0, sizeof(float) * std::max<i64>(nv, 1), c->stream));
        }
    }
    i64 m2 = (i64)amg->coarse_m * amg->coarse_m;
    GL_CUDA(cudaMalloc(&amg->coarse_inv32, sizeof(float) * std::max<i64>(m2, 1)));
    k_to_float<<<sgrid(m2), TPB, 0, c->stream>>>(amg->coarse_inv, amg->coarse_inv32, m2);
    GL_CUDA(cudaStreamSynchronize(c->stream));
}

void free_level(Level& l) {
    if (l.owns_pat) free_pattern(l.pat);
    if (l.owns_A && l.A) cudaFree(l.A);
    if (l.sym) { sym_free(l.sym); l.sym = nullptr; l.x32 = l.y32 = l.r32 = l.b32 = nullptr; }
    for (void* q : {(void*)l.dinv, (void*)l.agg, (void*)l.rvec, (void*)l.free_mask, (void*)l.mem_ptr, (void*)l.mem_idx,
                    (void*)l.X, (void*)l.x, (void*)l.b, (void*)l.r, (void*)l.d, (void*)l.A32, (void*)l.dinv32,
                    (void*)l.x32, (void*)l.b32, (void*)l.r32, (void*)l.d32, (void*)l.y32, (void*)l.A16})
        if (q) cudaFree(q);
}

}  // namespace

void amg_free(glims_ctx* c) {
    if (!c->amg) return;
    amg_fused_free(c->amg);
    for (auto& l : c->amg->L) free_level(l);
    if (c->amg->coarse_inv) cudaFree(c->amg->coarse_inv);
    if (c->amg->coarse_inv32) cudaFree(c->amg->coarse_inv32);
    delete c->amg;
    c->amg = nullptr;
}

void amg_setup(glims_ctx* c) {
    const bool verbose_t = std::getenv("GLIMS_VERBOSE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_levels = 0, t_dense = 0, t_lmax = 0;
    amg_free(c);
    Amg* amg = new Amg();
    c->amg = amg;
    const int D = c->dim;
    amg->dim = D;
    if (const char* e = std::getenv("GLIMS_AMG_DEGREE")) amg->cheb_degree = std::max(1, atoi(e));
    if (const char* e = std::getenv("GLIMS_AMG_RATIO")) amg->cheb_ratio = atof(e);
    amg->coarse_degree = amg->cheb_degree;
    if (const char* e = std::getenv("GLIMS_AMG_COARSE_DEGREE")) amg->coarse_degree = std::max(1, atoi(e));
    const int bsc = (D == 2) ? 3 : 6;
    GL_CUDA(cudaStreamSynchronize(c->stream));

    // ---- level 0 wraps K_uu ---------------------------------------------------------------------
    Level l0;
    l0.bs = D; l0.n = c->pat.n_rows; l0.n_cols = c->n_v; l0.pat = c->pat; l0.A = c->Kuu;
    std::vector<double> X(c->n_v * D);
    GL_CUDA(cudaMemcpy(X.data(), c->coords, sizeof(double) * c->n_v * D, cudaMemcpyDeviceToHost));
    std::vector<unsigned char> bcm(c->n_v);
    GL_CUDA(cudaMemcpy(bcm.data(), c->bcmask, c->n_v, cudaMemcpyDeviceToHost));
    std::vector<unsigned char> freem(c->n_v);
    const unsigned umask = (1u << D) - 1u;
    for (i64 i = 0; i < c->n_v; ++i) freem[i] = (unsigned char)(~bcm[i] & umask);
    GL_CUDA(cudaMalloc(&l0.free_mask, c->n_v));
    GL_CUDA(cudaMemcpy(l0.free_mask, freem.data(), c->n_v, cudaMemcpyHostToDevice));
    amg->L.push_back(l0);

    std::vector<double> Xl = X;              // positions of the current level's nodes
    const bool dist = c->halo.active && c->halo.comm && c->halo.n_ranks > 1;
    const int R = dist ? c->halo.n_ranks : 1, rank = dist ? c->halo.rank : 0;
    amg->dist = dist; amg->n_ranks = R; amg->rank = rank;
    std::vector<char> pad1;                  // dist: padding nodes of the global level 1
    int level = 0;
    while (true) {
        Level& l = amg->L[level];
        alloc_work(l);
        diag_inverse(c, l, level == 0 ? 0.0 : 1e-8);
        const int n = l.n;
        // coarsest level = dense inverse: up to ~256 nodes (one matrix-vector product instead of another smoothed level:
        // every level costs seven latency-bound phases per V-cycle)
        const bool last = ((n * l.bs <= 1536) && !(dist && level == 0)) || level >= 12;
        if (last) break;
        // ---- aggregation on the host ---------------------------------------------------------------
        HostGraph g = download_graph(l.pat);
        std::vector<char> excl(n, 0);
        if (level == 0) for (int i = 0; i < n; ++i) excl[i] = (freem[i] == 0);
        if (level == 1 && dist) for (int i = 0; i < n; ++i) excl[i] = pad1[i];      // padding nodes of the rank segments
        std::vector<int> agg;
        int na = aggregate(g, n, excl, agg);
        {
            const char* es = std::getenv("GLIMS_AMG_SORT");
            const int window = es ? atoi(es) : 0;      // measured at C4: 23 % fewer level-1 slots, but the smoother step got slower (75 -> 82 us: gather locality) and PCG needed 5 % more iterations -- off by default
            if (window > 0) relabel_by_degree(g, n, agg, na, window);
        }
        if (!(dist && level == 0) && (na == 0 || na >= n)) break;
        // centroids, offsets, member lists
        std::vector<double> Xc((i64)std::max(na, 1) * D, 0.0);
        std::vector<int> cnt(std::max(na, 1), 0);
        for (int i = 0; i < n; ++i) if (agg[i] >= 0) { cnt[agg[i]]++; for (int k = 0; k < D; ++k) Xc[(i64)agg[i] * D + k] += Xl[(i64)i * D + k]; }
        for (int I = 0; I < na; ++I) for (int k = 0; k < D; ++k) Xc[(i64)I * D + k] /= cnt[I];
        std::vector<double> rv((i64)std::max<i64>(l.n_cols, n) * D, 0.0);
        for (int i = 0; i < n; ++i) if (agg[i] >= 0) for (int k = 0; k < D; ++k) rv[(i64)i * D + k] = Xl[(i64)i * D + k] - Xc[(i64)agg[i] * D + k];
        std::vector<int> mptr(na + 1, 0), midx;
        for (int I = 0; I < na; ++I) mptr[I + 1] = mptr[I] + cnt[I];
        midx.resize(mptr[na]);
        { std::vector<int> pos(mptr.begin(), mptr.end() - 1); for (int i = 0; i < n; ++i) if (agg[i] >= 0) midx[pos[agg[i]]++] = i; }
        std::vector<int> aggfull(std::max<i64>(l.n_cols, n), -1);
        std::copy(agg.begin(), agg.end(), aggfull.begin());
        int n_coarse = na;            // nodes of the next level
        i64 seg = 0;
        if (dist && level == 0) {
            // Global numbering of the aggregates: rank r owns ids [r*seg, r*seg + na_r), seg = the largest rank count
            // rounded up to a whole SELL slice (ids na_r .. seg-1 are padding nodes: identity rows, never aggregated).
            long long m = ((long long)std::max(na, 1) + 31) / 32 * 32;
            comm_allreduce_max_i64(c, &m, 1);
            seg = m;
            n_coarse = (int)(seg * R);
            const int off = (int)(seg * rank);
            for (int i = 0; i < n; ++i) if (aggfull[i] >= 0) aggfull[i] += off;
            // ghost vertices: aggregate id and centroid of the owner, through the level-0 halo exchange
            {
                const int bsx = 1 + D;
                std::vector<double> hx((i64)c->n_v * bsx, 0.0);
                for (int i = 0; i < n; ++i) {
                    hx[(i64)i * bsx] = (double)aggfull[i];
                    if (agg[i] >= 0) for (int k = 0; k < D; ++k) hx[(i64)i * bsx + 1 + k] = Xc[(i64)agg[i] * D + k];
                }
                double* dx = nullptr;
                GL_CUDA(cudaMalloc(&dx, sizeof(double) * hx.size()));
                GL_CUDA(cudaMemcpy(dx, hx.data(), sizeof(double) * hx.size(), cudaMemcpyHostToDevice));
                halo_exchange(c, dx, bsx);
                GL_CUDA(cudaStreamSynchronize(c->stream));
                GL_CUDA(cudaMemcpy(hx.data(), dx, sizeof(double) * hx.size(), cudaMemcpyDeviceToHost));
                cudaFree(dx);
                for (i64 j = n; j < c->n_v; ++j) {
                    const int gid = (int)std::llround(hx[j * bsx]);
                    aggfull[j] = gid;
                    if (gid >= 0) for (int k = 0; k < D; ++k) rv[j * D + k] = X[j * D + k] - hx[j * bsx + 1 + k];
                }
            }
            l.nc_off = off; l.nc_local = na;
        }
        l.bsc = bsc; l.nc = n_coarse;
        if (!(dist && level == 0)) { l.nc_off = 0; l.nc_local = na; }
        GL_CUDA(cudaMalloc(&l.agg, sizeof(int) * aggfull.size()));
        GL_CUDA(cudaMemcpy(l.agg, aggfull.data(), sizeof(int) * aggfull.size(), cudaMemcpyHostToDevice));
        GL_CUDA(cudaMalloc(&l.rvec, sizeof(double) * rv.size()));
        GL_CUDA(cudaMemcpy(l.rvec, rv.data(), sizeof(double) * rv.size(), cudaMemcpyHostToDevice));
        GL_CUDA(cudaMalloc(&l.mem_ptr, sizeof(int) * (na + 1)));
        GL_CUDA(cudaMemcpy(l.mem_ptr, mptr.data(), sizeof(int) * (na + 1), cudaMemcpyHostToDevice));
        GL_CUDA(cudaMalloc(&l.mem_idx, sizeof(int) * std::max<size_t>(midx.size(), 1)));
        GL_CUDA(cudaMemcpy(l.mem_idx, midx.data(), sizeof(int) * midx.size(), cudaMemcpyHostToDevice));
        // ---- Galerkin operator on the device -------------------------------------------------------
        Level lc;
        lc.bs = bsc; lc.n = n_coarse; lc.n_cols = n_coarse; lc.owns_pat = true; lc.owns_A = true;
        unsigned long long *keys = nullptr, *ukeys = nullptr;
        const i64 n_extra = (dist && level == 0) ? seg : 0;      // diagonal keys of this rank's whole segment
        GL_CUDA(cudaMalloc(&keys, sizeof(unsigned long long) * std::max<i64>(l.pat.n_slots + n_extra, 1)));
        k_coarse_keys<<<l.pat.n_slices, 128, 0, c->stream>>>(l.pat.slice_off, l.pat.slice_w, l.pat.col, l.pat.n_rows, l.agg, keys);
        if (n_extra > 0) k_diag_keys<<<nblk(n_extra), TPB, 0, c->stream>>>(keys + l.pat.n_slots, (int)(seg * rank), (int)n_extra);
        if (!(dist && level == 0)) {
            build_pattern_from_keys(c->stream, keys, l.pat.n_slots, n_coarse, lc.pat, &ukeys);
            cudaFree(keys);
            GL_CUDA(cudaMalloc(&lc.A, sizeof(double) * lc.pat.n_slots * bsc * bsc));
            GL_CUDA(cudaMemsetAsync(lc.A, 0, sizeof(double) * lc.pat.n_slots * bsc * bsc, c->stream));
            if (D == 2) k_galerkin<2><<<l.pat.n_slices, 128, 0, c->stream>>>(l.pat.slice_off, l.pat.slice_w, l.pat.col, l.pat.n_rows, level == 0, l.A, l.agg, l.rvec, l.free_mask, ukeys, lc.pat.rowptr, lc.pat.slice_off, lc.A);
            else k_galerkin<3><<<l.pat.n_slices, 128, 0, c->stream>>>(l.pat.slice_off, l.pat.slice_w, l.pat.col, l.pat.n_rows, level == 0, l.A, l.agg, l.rvec, l.free_mask, ukeys, lc.pat.rowptr, lc.pat.slice_off, lc.A);
            GL_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(ukeys);
            if (dist) {
                // levels >= 2 are built redundantly; Galerkin sums use atomics, so make the bits identical everywhere
                std::vector<long long> cntv(R, 0);
                cntv[0] = (long long)lc.pat.n_slots * bsc * bsc;
                comm_allgatherv(c, lc.A, lc.A, cntv, sizeof(double));
            }
        } else {
            // this rank's rows of the global operator (rows of the other segments stay empty) ...
            SellPattern ploc;
            build_pattern_from_keys(c->stream, keys, l.pat.n_slots + n_extra, n_coarse, ploc, &ukeys);
            cudaFree(keys);
            double* Aloc = nullptr;
            const int bb = bsc * bsc;
            GL_CUDA(cudaMalloc(&Aloc, sizeof(double) * std::max<i64>(ploc.n_slots, 1) * bb));
            GL_CUDA(cudaMemsetAsync(Aloc, 0, sizeof(double) * std::max<i64>(ploc.n_slots, 1) * bb, c->stream));
            if (D == 2) k_galerkin<2><<<l.pat.n_slices, 128, 0, c->stream>>>(l.pat.slice_off, l.pat.slice_w, l.pat.col, l.pat.n_rows, true, l.A, l.agg, l.rvec, l.free_mask, ukeys, ploc.rowptr, ploc.slice_off, Aloc);
            else k_galerkin<3><<<l.pat.n_slices, 128, 0, c->stream>>>(l.pat.slice_off, l.pat.slice_w, l.pat.col, l.pat.n_rows, true, l.A, l.agg, l.rvec, l.free_mask, ukeys, ploc.rowptr, ploc.slice_off, Aloc);
            // ... gathered from all ranks: unique keys are in CSR order, rank segments are contiguous row ranges, so the
            // concatenation in rank order is the CSR order of the global matrix
            const i64 nnz_loc = ploc.nnzb;
            double* vloc = nullptr;
            GL_CUDA(cudaMalloc(&vloc, sizeof(double) * std::max<i64>(nnz_loc, 1) * bb));
            k_sell_to_csr<<<nblk(n_coarse), TPB, 0, c->stream>>>(ploc.rowptr, ploc.slice_off, n_coarse, bb, Aloc, vloc);
            GL_CUDA(cudaStreamSynchronize(c->stream));
            std::vector<long long> counts;
            comm_allgather_i64(c, nnz_loc, counts);
            i64 nnz_tot = 0;
            for (long long v : counts) nnz_tot += v;
            unsigned long long* gkeys = nullptr;
            double* gvals = nullptr;
            GL_CUDA(cudaMalloc(&gkeys, sizeof(unsigned long long) * std::max<i64>(nnz_tot, 1)));
            GL_CUDA(cudaMalloc(&gvals, sizeof(double) * std::max<i64>(nnz_tot, 1) * bb));
            comm_allgatherv(c, ukeys, gkeys, counts, sizeof(unsigned long long));
            comm_allgatherv(c, vloc, gvals, counts, sizeof(double) * bb);
            cudaFree(ukeys); cudaFree(vloc); cudaFree(Aloc);
            free_pattern(ploc);
            build_pattern_from_keys(c->stream, gkeys, nnz_tot, n_coarse, lc.pat, nullptr);       // consumes gkeys (already sorted + unique)
            cudaFree(gkeys);
            if (lc.pat.nnzb != nnz_tot) throw GlError(GLIMS_ERR_NCCL, "amg: gathered level-1 operator has duplicate blocks");
            GL_CUDA(cudaMalloc(&lc.A, sizeof(double) * lc.pat.n_slots * bb));
            GL_CUDA(cudaMemsetAsync(lc.A, 0, sizeof(double) * lc.pat.n_slots * bb, c->stream));
            k_csr_to_sell<<<nblk(n_coarse), TPB, 0, c->stream>>>(lc.pat.rowptr, lc.pat.slice_off, n_coarse, bb, gvals, lc.A);
            GL_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(gvals);
            lc.dist_rows = true; lc.row0 = (int)(seg * rank); lc.rows = (int)seg;
            // node positions and the padding mask of the global level (identical on every rank)
            std::vector<double> Xg((i64)n_coarse * D, 0.0);
            {
                double *dl = nullptr, *dg = nullptr;
                std::vector<double> mine((i64)seg * D, 0.0);
                std::copy(Xc.begin(), Xc.begin() + (i64)na * D, mine.begin());
                GL_CUDA(cudaMalloc(&dl, sizeof(double) * mine.size()));
                GL_CUDA(cudaMalloc(&dg, sizeof(double) * Xg.size()));
                GL_CUDA(cudaMemcpy(dl, mine.data(), sizeof(double) * mine.size(), cudaMemcpyHostToDevice));
                std::vector<long long> eq(R, (long long)seg * D);
                comm_allgatherv(c, dl, dg, eq, sizeof(double));
                GL_CUDA(cudaMemcpy(Xg.data(), dg, sizeof(double) * Xg.size(), cudaMemcpyDeviceToHost));
                cudaFree(dl); cudaFree(dg);
            }
            std::vector<long long> nas;
            comm_allgather_i64(c, na, nas);
            pad1.assign(n_coarse, 0);
            for (int r = 0; r < R; ++r) for (i64 i = nas[r]; i < seg; ++i) pad1[(i64)r * seg + i] = 1;
            Xc.swap(Xg);
        }
        amg->L.push_back(lc);
        Xl.swap(Xc);
        ++level;
    }
    t_levels = now();
    // ---- coarsest level: dense (regularised) inverse on the host --------------------------------------
    {
        Level& l = amg->L.back();
        const int bs = l.bs, n = l.n, m = n * bs;
        // aggregation that stalls on a large level must not end in an O(m^2) dense inverse: fail loudly
        if (m > 8192) throw GlError(GLIMS_ERR_STATE, "amg: coarsening stalled at " + std::to_string(n) + " nodes on level " + std::to_string(amg->L.size() - 1) +
                                                     "; use GLIMS_PC_JACOBI for this mesh");
        amg->coarse_m = m;
        HostGraph g = download_graph(l.pat);
        std::vector<double> Av((i64)l.pat.n_slots * bs * bs);
        std::vector<i64> so(l.pat.n_slices + 1);
        GL_CUDA(cudaMemcpy(Av.data(), l.A, sizeof(double) * Av.size(), cudaMemcpyDeviceToHost));
        GL_CUDA(cudaMemcpy(so.data(), l.pat.slice_off, sizeof(i64) * so.size(), cudaMemcpyDeviceToHost));
        std::vector<double> M((i64)m * m, 0.0), Inv((i64)m * m, 0.0);
        for (int r = 0; r < n; ++r)
            for (i64 t = g.rowptr[r]; t < g.rowptr[r + 1]; ++t) {
                i64 s = so[r >> 5] + (t - g.rowptr[r]) * 32 + (r & 31);
                int cj = g.col[t];
                if (cj >= n) continue;
                for (int i = 0; i < bs; ++i) for (int j = 0; j < bs; ++j)
                    M[(i64)(r * bs + i) * m + cj * bs + j] = Av[vidx(s, i * bs + j, bs * bs)];
            }
        double dmax = 0;
        for (int i = 0; i < m; ++i) dmax = std::max(dmax, std::fabs(M[(i64)i * m + i]));
        if (dmax == 0) dmax = 1.0;
        for (int i = 0; i < m; ++i) {
            if (M[(i64)i * m + i] == 0.0) M[(i64)i * m + i] = dmax;     // empty rows (constrained nodes on a one-level hierarchy)
            M[(i64)i * m + i] += 1e-10 * dmax;                          // floating sub-domains: regularise the rigid modes
            Inv[(i64)i * m + i] = 1.0;
        }
        GL_CUDA(cudaMalloc(&amg->coarse_inv, sizeof(double) * std::max<i64>((i64)m * m, 1)));
        if (m > 600) {
            // large coarsest level: invert on the device (the O(m^3) host loop would take seconds)
            double *dM = nullptr, *df = nullptr;
            GL_CUDA(cudaMalloc(&dM, sizeof(double) * (i64)m * m));
            GL_CUDA(cudaMalloc(&df, sizeof(double) * m));
            GL_CUDA(cudaMemcpy(dM, M.data(), sizeof(double) * (i64)m * m, cudaMemcpyHostToDevice));
            GL_CUDA(cudaMemcpy(amg->coarse_inv, Inv.data(), sizeof(double) * (i64)m * m, cudaMemcpyHostToDevice));
            for (int p = 0; p < m; ++p) {
                k_gj_pivot<<<1, 1024, 0, c->stream>>>(dM, amg->coarse_inv, m, p, df);
                k_gj_update<<<dim3((2 * m + 255) / 256, m), 256, 0, c->stream>>>(dM, amg->coarse_inv, m, p, df);
            }
            GL_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(dM); cudaFree(df);
        } else {
        for (int p = 0; p < m; ++p) {
            int piv = p; double best = std::fabs(M[(i64)p * m + p]);
            for (int i = p + 1; i < m; ++i) if (std::fabs(M[(i64)i * m + p]) > best) { best = std::fabs(M[(i64)i * m + p]); piv = i; }
            if (piv != p) for (int j = 0; j < m; ++j) { std::swap(M[(i64)p * m + j], M[(i64)piv * m + j]); std::swap(Inv[(i64)p * m + j], Inv[(i64)piv * m + j]); }
            double ip = 1.0 / M[(i64)p * m + p];
            for (int j = 0; j < m; ++j) { M[(i64)p * m + j] *= ip; Inv[(i64)p * m + j] *= ip; }
            for (int i = 0; i < m; ++i) {
                if (i == p) continue;
                double f = M[(i64)i * m + p];
                if (f == 0.0) continue;
                for (int j = 0; j < m; ++j) { M[(i64)i * m + j] -= f * M[(i64)p * m + j]; Inv[(i64)i * m + j] -= f * Inv[(i64)p * m + j]; }
            }
        }
        GL_CUDA(cudaMemcpy(amg->coarse_inv, Inv.data(), sizeof(double) * (i64)m * m, cudaMemcpyHostToDevice));
        }
    }
    t_dense = now();
    // ---- smoother spectra ----------------------------------------------------------------------------
    for (size_t li = 0; li + 1 < amg->L.size(); ++li) estimate_lmax(c, amg->L[li]);
    GL_CUDA(cudaStreamSynchronize(c->stream));
    if (dist) {
        // one smoother for the whole mesh: the level-0 spectrum estimate is rank-local, take the largest; the other
        // levels are replicated (identical data, deterministic kernels), agree on them anyway
        std::vector<double> lm(amg->L.size(), 0.0);
        for (size_t li = 0; li < amg->L.size(); ++li) lm[li] = amg->L[li].lmax;
        comm_allreduce_max_f64(c, lm.data(), (int)lm.size());
        for (size_t li = 0; li < amg->L.size(); ++li) amg->L[li].lmax = lm[li];
    }
    t_lmax = now();
    build_fp32(c, amg);
    if (verbose_t) fprintf(stderr, "glims amg setup: levels %.3f s, coarsest inverse %.3f s, spectra %.3f s, fp32 copies %.3f s\n", t_levels - t_begin, t_dense - t_levels, t_lmax - t_dense, now() - t_lmax);
    if (std::getenv("GLIMS_VERBOSE"))
        for (size_t li = 0; li < amg->L.size(); ++li) {
            const Level& l = amg->L[li];
            fprintf(stderr, "glims amg level %zu: %d nodes x %d dofs, %lld blocks (%lld slots, max row %d), lmax %.3f\n", li, l.n, l.bs,
                    (long long)l.pat.nnzb, (long long)l.pat.n_slots, l.pat.max_w, l.lmax);
        }
}

// throws if a device-side wait of the preconditioner timed out (fused coarse kernel barrier, level-1 all-gather)
void amg_check(glims_ctx* c) {
    Amg* amg = c->amg;
    if (!amg) return;
    if (amg_fused_failed(amg)) throw GlError(GLIMS_ERR_CUDA, "amg: the fused coarse-level kernel timed out at a grid barrier (grid not co-resident)");
    for (auto& l : amg->L) if (l.sym && sym_check(l.sym)) throw GlError(GLIMS_ERR_NCCL, "amg: level-1 all-gather timed out waiting for a peer rank");
}

// coarse-grid correction below level 1 with the hierarchy's own buffers (glims_time_kernel 9): the fused persistent kernel,
// or the launch sequence it replaces with GLIMS_AMG_FUSED=0
bool amg_time_coarse(glims_ctx* c) {
    Amg* amg = c->amg;
    if (!amg || amg->L.size() < 3 || !amg->L[1].y32) return false;
    // the iterate buffer the V-cycle holds at this point (after `coarse_degree` - 1 ping-pong steps from y32)
    float* cur = ((amg->coarse_degree - 1) & 1) ? amg->L[1].x32 : amg->L[1].y32;
    coarse_correction32(c, amg, 1, cur);
    if (std::getenv("GLIMS_VERBOSE")) { cudaStreamSynchronize(c->stream); amg_fused_print_phases(amg); }
    return true;
}

// one fused smoother step on level 1 (glims_time_kernel 10)
bool amg_time_level1_step(glims_ctx* c) {
    Amg* amg = c->amg;
    if (!amg || amg->L.size() < 3 || !amg->L[1].y32) return false;
    Level& l = amg->L[1];
    cheb_step32(c, l, l.b32, l.x32, l.y32, 0.3f, 0.5f);
    return true;
}

// one fused smoother step on the fine level with the hierarchy's own buffers (roofline bench, glims_time_kernel 6)
bool amg_time_fine_step(glims_ctx* c) {
    Amg* amg = c->amg;
    if (!amg || amg->L.size() < 2 || !amg->L[0].dinv32) return false;
    Level& l = amg->L[0];
    cheb_step32(c, l, l.b32, l.x32, l.y32, 0.3f, 0.5f);
    return true;
}

void amg_vcycle(glims_ctx* c, const double* r, double* z, bool fp32) {
    Amg* amg = c->amg;
    if (amg->dist) fp32 = true;      // partitioned mesh: the hierarchy with the replicated level 1 exists in FP32 only
    if (!fp32) {
        if (amg->L.size() == 1) {     // tiny problem: the dense inverse is the whole hierarchy
            k_dense_matvec<<<(amg->coarse_m + 7) / 8, 256, 0, c->stream>>>(amg->coarse_inv, r, z, amg->coarse_m);
            c->launches++;
            return;
        }
        vcycle(c, amg, 0, r, z);
        return;
    }
    Level& l0 = amg->L[0];
    const i64 n = (i64)l0.n * l0.bs;
    k_to_float<<<sgrid(n), TPB, 0, c->stream>>>(r, l0.b32, n);
    if (amg->L.size() == 1) k_dense_matvec32<<<(amg->coarse_m + 7) / 8, 256, 0, c->stream>>>(amg->coarse_inv32, l0.b32, l0.x32, amg->coarse_m);
    else vcycle32(c, amg, 0, l0.b32, l0.x32);
    k_to_double<<<sgrid(n), TPB, 0, c->stream>>>(l0.x32, z, n);
    c->launches += 3;
}
