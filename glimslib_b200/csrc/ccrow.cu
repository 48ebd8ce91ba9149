// Row-walk assembly (variant GLIMS_ASMK_ROWS): the per-Newton-iteration pass of the block-triangular solver,
// deterministic and atomics-free.
//
// What changes between Newton iterations is only the concentration block (stg:115-120): K_uu and K_uc depend on
// the mesh and the materials alone.  With per-slot constants assembled once per (mesh, materials, dt)
//   Klin(a,b) = sum_e |K| [ m(1+d_ab)(1 - dt rho_e) + dt D_e grad(l_a).grad(l_b) ]      (linear part of K_cc)
//   M(a,b)    = sum_e |K| m(1+d_ab)                                                      (mass matrix)
// and the per-element weight rvol_e = rho_e |K_e|, the state-dependent part is
//   K_cc(a,b) = Klin(a,b) + 2 dt kappa * A(a,b),   A(a,b) = sum_{e contains a,b} rvol_e * (a==b ? 4c_a + 2S_e : c_a + c_b + S_e)
//   F_c[a]    = sum_b (Klin + dt kappa A)(a,b) c_b - (M c_prev)[a] - f_ext,c[a]           (R = J_r c / 2)
//   F_u       = K_uu u + K_uc c (+ lift of eliminated Dirichlet columns) - f_ext,u        (linear: one SpMV, k_fu)
//
// k_cc_rows: one warp per SELL-32 slice, lane = block row.  The row walks its incident elements (a padded
// [iteration][lane] list of (element, own local index, the columns of the element's vertices inside the row)) and adds
// into a lane-private strip of shared memory (address = column*32 + lane: conflict-free whatever the columns are);
// then one pass over the row's slots writes K_cc and forms F_c.  Every value is written exactly once, the summation
// order is the list order (elements ascending): bitwise reproducible.  No element->slot map is read (the positions
// travel in the list: 8 bytes per (row, element) pair instead of 64 bytes per element of eslot + RED.ADD traffic).
#include "common.h"
#include "geom.cuh"
#include "tma.cuh"
#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>
#include <thrust/sort.h>
#include <thrust/scan.h>
#include <thrust/binary_search.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/reduce.h>
#include <algorithm>
#include <cstdlib>

struct CcMap {
    bool ok = false;
    std::string why;
    i64* pl_off = nullptr;      // [n_slices+1] first list entry of a slice (multiple of 32)
    int* pl_w = nullptr;        // [n_slices] list iterations of the slice (longest row)
    unsigned* pe = nullptr;     // [total] element | own local vertex << 30 ; 0xFFFFFFFF = padding
    unsigned* pp = nullptr;     // [total] column (inside the row) of the element's local vertex q in byte q
    double* prv = nullptr;      // [total] rho_e |K_e| of the entry's element (0 for padding): what k_cc_rows streams
    unsigned* ppa = nullptr;    // [total] four 7-bit columns | own local vertex << 28 (0 for padding)
    i64 total = 0;
    int max_pw = 0;
    double* rvol = nullptr;     // [n_c] rho_e |K_e|
    double* Klin = nullptr;     // [n_slots]
    double* Mass = nullptr;     // [n_slots]
    double* mcp = nullptr;      // [n_rows] M c_prev of the current step
    double* lift = nullptr;     // [n_rows][dim] K_raw(:, Dirichlet columns) g  (null: all Dirichlet values are zero)
    bool const_valid = false;   // Klin / Mass / rvol match materials and dt
    size_t map_bytes = 0;
};

namespace {

constexpr int CC_TPB = 128;                 // 4 warps = 4 slices per CTA
constexpr unsigned CC_PAD = 0xFFFFFFFFu;
constexpr int CC_MAX_W = 96;                // widest row (block columns): 7-bit columns in the stream, shared-memory strips
inline int nblk(i64 n, int t = 256) { return (int)((n + t - 1) / t); }

__global__ void k_pair_keys(const int* __restrict__ cells, i64 n_c, int nb, i64 n_own, unsigned long long* keys) {
    i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (p >= n_c * nb) return;
    const i64 e = p / nb;
    const int a = (int)(p - e * nb);
    const int v = cells[p];
    keys[p] = v < n_own ? (((unsigned long long)v << 36) | ((unsigned long long)e << 2) | (unsigned)a) : ~0ULL;
}
struct PairRowStart {
    __host__ __device__ unsigned long long operator()(i64 r) const { return (unsigned long long)r << 36; }
};
__global__ void k_pair_width(const i64* __restrict__ pstart, int n_rows, int n_slices, int* w) {
    int S = blockIdx.x * blockDim.x + threadIdx.x;
    if (S >= n_slices) return;
    int m = 0;
    for (int l = 0; l < 32; ++l) {
        int r = S * 32 + l;
        if (r < n_rows) m = max(m, (int)(pstart[r + 1] - pstart[r]));
    }
    w[S] = (m + 3) & ~3;      // k_cc_rows walks the list four iterations at a time
}
__global__ void k_times32(i64* v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= 32;
}
__global__ void k_pair_fill(const unsigned long long* __restrict__ keys, i64 n, int nb, const i64* __restrict__ pstart,
                            const i64* __restrict__ pl_off, const i64* __restrict__ slice_off,
                            const int* __restrict__ eslot, unsigned* __restrict__ pe, unsigned* __restrict__ pp) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned long long k = keys[t];
    const int v = (int)(k >> 36), a = (int)(k & 3);
    const i64 e = (i64)((k >> 2) & ((1ULL << 34) - 1));
    const i64 dest = pl_off[v >> 5] + (t - pstart[v]) * 32 + (v & 31);
    unsigned pos = 0;
    for (int b = 0; b < nb; ++b) {
        const i64 s = eslot[e * nb * nb + a * nb + b];
        pos |= (unsigned)((s - slice_off[v >> 5] - (v & 31)) >> 5) << (8 * b);
    }
    pe[dest] = (unsigned)e | ((unsigned)a << 30);
    pp[dest] = pos;
}

template <int D>
__global__ void k_rvol(const double* __restrict__ coords, const int* __restrict__ cells, const int* __restrict__ cell_mat,
                       const double* __restrict__ mat, i64 n_c, double* __restrict__ rvol) {
    constexpr int NB = D + 1;
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= n_c) return;
    double X[NB][D];
#pragma unroll
    for (int a = 0; a < NB; ++a) {
        const i64 v = cells[e * NB + a];
#pragma unroll
        for (int k = 0; k < D; ++k) X[a][k] = coords[v * D + k];
    }
    Geo<D> G;
    geometry(X, G);
    rvol[e] = mat[cell_mat[e] * MAT_STRIDE + 3] * G.vol;
}

// One-time pass: Klin and Mass per slot, same row-walk as k_cc_rows (deterministic).
template <int D>
__global__ void __launch_bounds__(CC_TPB)
k_cc_setup(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, int n_slices,
           const i64* __restrict__ pl_off, const int* __restrict__ pl_w, const unsigned* __restrict__ pe,
           const unsigned* __restrict__ pp, const double* __restrict__ coords, const int* __restrict__ cells,
           const int* __restrict__ cell_mat, const double* __restrict__ mat, double dt, int max_w,
           double* __restrict__ Klin, double* __restrict__ Mass) {
    constexpr int NB = D + 1;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* aK = sm + (size_t)warp * 2 * max_w * 32;
    double* aM = aK + max_w * 32;
    const int S = blockIdx.x * (CC_TPB / 32) + warp;
    if (S >= n_slices) return;
    const i64 base = slice_off[S];
    const int w = slice_w[S];
    for (int j = 0; j < w; ++j) { aK[j * 32 + lane] = 0.0; aM[j * 32 + lane] = 0.0; }
    const i64 po = pl_off[S];
    const int pw = pl_w[S];
    for (int t = 0; t < pw; ++t) {
        const unsigned ea = pe[po + (i64)t * 32 + lane];
        if (ea == CC_PAD) continue;
        const unsigned ps = pp[po + (i64)t * 32 + lane];
        const i64 e = ea & 0x3FFFFFFFu;
        const int a = (int)(ea >> 30);
        double X[NB][D];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const i64 v = cells[e * NB + q];
#pragma unroll
            for (int k = 0; k < D; ++k) X[q][k] = coords[v * D + k];
        }
        Geo<D> G;
        geometry(X, G);
        const double* m = mat + cell_mat[e] * MAT_STRIDE;
        const double Dc = m[2], rho = m[3];
        double ga[D];
#pragma unroll
        for (int q = 0; q < NB; ++q)
            if (q == a) {
#pragma unroll
                for (int k = 0; k < D; ++k) ga[k] = G.g[q][k];
            }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            double gg = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) gg += ga[k] * G.g[q][k];
            const double mf = Consts<D>::mass * (q == a ? 2.0 : 1.0);
            const int p = (int)((ps >> (8 * q)) & 255u) * 32 + lane;
            aK[p] += G.vol * (mf * (1.0 - dt * rho) + dt * Dc * gg);
            aM[p] += G.vol * mf;
        }
    }
    for (int j = 0; j < w; ++j) {
        const i64 s = base + (i64)j * 32 + lane;
        Klin[s] = aK[j * 32 + lane];
        Mass[s] = aM[j * 32 + lane];
    }
}

// mcp = M c_prev (once per time step; c_prev read from the vertex-blocked u_previous)
template <int D>
__global__ void __launch_bounds__(256)
k_mass_cprev(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
             const double* __restrict__ Mass, const double* __restrict__ xprev, int n_rows, double* __restrict__ mcp) {
    constexpr int NB = D + 1;
    const int r = blockIdx.x * 256 + threadIdx.x;
    const int S = r >> 5, lane = r & 31;
    if (S * 32 >= n_rows) return;
    const i64 base = slice_off[S];
    const int w = slice_w[S];
    double acc = 0.0;
    for (int j = 0; j < w; ++j) {
        const i64 s = base + (i64)j * 32 + lane;
        acc += __ldcs(&Mass[s]) * __ldg(&xprev[(i64)__ldg(&col[s]) * NB + D]);
    }
    if (r < n_rows) mcp[r] = acc;
}

// (rho|K|, columns) stream of the main kernel from the element ids (after every change of the materials)
__global__ void k_pair_stream(const unsigned* __restrict__ pe, const unsigned* __restrict__ pp, i64 n,
                              const double* __restrict__ rvol, double* __restrict__ prv, unsigned* __restrict__ ppa) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned ea = pe[t];
    if (ea == CC_PAD) { prv[t] = 0.0; ppa[t] = 0u; return; }
    const unsigned ps = pp[t];
    prv[t] = rvol[ea & 0x3FFFFFFFu];
    ppa[t] = (ps & 127u) | (((ps >> 8) & 127u) << 7) | (((ps >> 16) & 127u) << 14) | (((ps >> 24) & 127u) << 21) | ((ea >> 30) << 28);
}

// The per-Newton-iteration pass: K_cc (WK) and F_c from the current concentration.  Branch-free: padding entries carry
// rho|K| = 0 and column 0, i.e. they add 0.0; the list is streamed four iterations ahead of the shared-memory updates.
template <int D, bool WK>
__global__ void __launch_bounds__(CC_TPB)
k_cc_rows(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col, int n_rows,
          int n_slices, const i64* __restrict__ pl_off, const int* __restrict__ pl_w, const double* __restrict__ prv,
          const unsigned* __restrict__ ppa, const double* __restrict__ Klin,
          const double* __restrict__ mcp, const double* __restrict__ fext, const double* __restrict__ x, double dt,
          int max_w, double* __restrict__ Kcc, double* __restrict__ F) {
    constexpr int NB = D + 1;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* __restrict__ cc = sm + (size_t)warp * 2 * max_w * 32 + lane;     // c of the row's columns   [column*32]
    double* __restrict__ acc = cc + max_w * 32;                              // A(a, column)             [column*32]
    const int S = blockIdx.x * (CC_TPB / 32) + warp;
    if (S >= n_slices) return;
    const i64 base = slice_off[S] + lane;
    const int w = slice_w[S];
#pragma unroll 4
    for (int j = 0; j < w; ++j) {
        const int cidx = __ldg(&col[base + (i64)j * 32]);
        cc[j * 32] = __ldg(&x[(i64)cidx * NB + D]);
        acc[j * 32] = 0.0;
    }
    const double* __restrict__ lrv = prv + pl_off[S] + lane;
    const unsigned* __restrict__ lpa = ppa + pl_off[S] + lane;
    const int pw = pl_w[S];            // multiple of 4
    double rv[4];
    unsigned ps[4];
    if (pw > 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { rv[u] = __ldcs(&lrv[u * 32]); ps[u] = __ldcs(&lpa[u * 32]); }
    }
    for (int t = 0; t < pw; t += 4) {
        double rvn[4] = {0.0, 0.0, 0.0, 0.0};
        unsigned psn[4] = {0u, 0u, 0u, 0u};
        if (t + 4 < pw) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { rvn[u] = __ldcs(&lrv[(t + 4 + u) * 32]); psn[u] = __ldcs(&lpa[(t + 4 + u) * 32]); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = (int)(ps[u] >> 28);
            int p[NB];
            double cq[NB], av[NB], Ssum = 0.0, ca = 0.0;
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                p[q] = (int)((ps[u] >> (7 * q)) & 127u) * 32;
                cq[q] = cc[p[q]];
                av[q] = acc[p[q]];
                Ssum += cq[q];
                if (q == a) ca = cq[q];
            }
            // the NB columns of a real entry are distinct; a padding entry hits column 0 NB times with +0.0
#pragma unroll
            for (int q = 0; q < NB; ++q)
                acc[p[q]] = av[q] + rv[u] * (q == a ? 4.0 * ca + 2.0 * Ssum : ca + cq[q] + Ssum);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { rv[u] = rvn[u]; ps[u] = psn[u]; }
    }
    const double kf = 2.0 * dt * Consts<D>::kappa;
    double fc = 0.0;
#pragma unroll 4
    for (int j = 0; j < w; ++j) {
        const i64 s = base + (i64)j * 32;
        const double kl = __ldcs(&Klin[s]);
        const double jr = kf * acc[j * 32];
        if (WK) __stcs(&Kcc[s], kl + jr);
        fc += (kl + 0.5 * jr) * cc[j * 32];
    }
    const int r = S * 32 + lane;
    if (F && r < n_rows) F[(i64)r * NB + D] = fc - mcp[r] - (fext ? fext[(i64)r * NB + D] : 0.0);
}

// ---- bulk-copy (TMA) variant ---------------------------------------------------------------------------------------------
// The (rho|K|, columns) list of a slice is two contiguous runs of global memory (pw*256 B and pw*128 B).  Lane 0 hands both to
// the bulk-copy engine (cp.async.bulk ... mbarrier::complete_tx) and the warp gathers the slice's column values while the
// copies fly; the pair loop then reads only shared memory.  No register staging, no long-scoreboard stall inside the loop,
// and ~10 KB in flight per warp instead of what fits in registers.
template <int D, bool WK>
__global__ void __launch_bounds__(CC_TPB)
k_cc_rows_tma(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col, int n_rows,
              int n_slices, const i64* __restrict__ pl_off, const int* __restrict__ pl_w, const double* __restrict__ prv,
              const unsigned* __restrict__ ppa, const double* __restrict__ Klin,
              const double* __restrict__ mcp, const double* __restrict__ fext, const double* __restrict__ x, double dt,
              int max_w, int max_pw, double* __restrict__ Kcc, double* __restrict__ F) {
    constexpr int NB = D + 1;
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long bars[CC_TPB / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t strip_b = (size_t)max_w * 256, wbytes = 2 * strip_b + (size_t)max_pw * 384;
    unsigned char* wbase = smraw + (size_t)warp * wbytes;
    double* __restrict__ cc = (double*)wbase + lane;                         // [column*32]
    double* __restrict__ acc = (double*)(wbase + strip_b) + lane;
    const double* __restrict__ srv = (const double*)(wbase + 2 * strip_b) + lane;                               // [t*32]
    const unsigned* __restrict__ spa = (const unsigned*)(wbase + 2 * strip_b + (size_t)max_pw * 256) + lane;    // [t*32]
    const int S = blockIdx.x * (CC_TPB / 32) + warp;
    if (S >= n_slices) return;
    const int pw = pl_w[S];
    const i64 po = pl_off[S];
    if (lane == 0) {
        mbar_init(&bars[warp], 1);
        if (pw > 0) {
            mbar_expect_tx(&bars[warp], (unsigned)pw * 384u);
            bulk_g2s(wbase + 2 * strip_b, prv + po, (unsigned)pw * 256u, &bars[warp]);
            bulk_g2s(wbase + 2 * strip_b + (size_t)max_pw * 256, ppa + po, (unsigned)pw * 128u, &bars[warp]);
        }
    }
    __syncwarp();
    const i64 base = slice_off[S] + lane;
    const int w = slice_w[S];
#pragma unroll 4
    for (int j = 0; j < w; ++j) {
        const int cidx = __ldg(&col[base + (i64)j * 32]);
        cc[j * 32] = __ldg(&x[(i64)cidx * NB + D]);
        acc[j * 32] = 0.0;
    }
    if (pw > 0) mbar_wait(&bars[warp], 0);
#pragma unroll 2
    for (int t = 0; t < pw; ++t) {
        const double rv = srv[t * 32];
        const unsigned ps = spa[t * 32];
        const int a = (int)(ps >> 28);
        int p[NB];
        double cq[NB], av[NB], Ssum = 0.0, ca = 0.0;
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            p[q] = (int)((ps >> (7 * q)) & 127u) * 32;
            cq[q] = cc[p[q]];
            av[q] = acc[p[q]];
            Ssum += cq[q];
            if (q == a) ca = cq[q];
        }
#pragma unroll
        for (int q = 0; q < NB; ++q)
            acc[p[q]] = av[q] + rv * (q == a ? 4.0 * ca + 2.0 * Ssum : ca + cq[q] + Ssum);
    }
    const double kf = 2.0 * dt * Consts<D>::kappa;
    double fc = 0.0;
#pragma unroll 4
    for (int j = 0; j < w; ++j) {
        const i64 s = base + (i64)j * 32;
        const double kl = __ldcs(&Klin[s]);
        const double jr = kf * acc[j * 32];
        if (WK) __stcs(&Kcc[s], kl + jr);
        fc += (kl + 0.5 * jr) * cc[j * 32];
    }
    const int r = S * 32 + lane;
    if (F && r < n_rows) F[(i64)r * NB + D] = fc - mcp[r] - (fext ? fext[(i64)r * NB + D] : 0.0);
}

// F_u = K_uu u + K_uc c + lift - f_ext,u on vertex-blocked x; one thread per block row (as k_spmv_mono without the
// concentration row).  xmask != null: multiply x by the Dirichlet mask first (used to compute the lift itself).
template <int D>
__global__ void __launch_bounds__(256)
k_fu(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
     const double* __restrict__ Kuu, const double* __restrict__ Kuc, const double* __restrict__ x, int n_rows,
     const double* __restrict__ lift, const double* __restrict__ fext, double* __restrict__ out, int out_stride) {
    constexpr int NB = D + 1;
    const int n_tiles = (n_rows + 255) / 256;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r = tile * 256 + threadIdx.x;
        const int S = r >> 5, lane = r & 31;
        if (S * 32 >= n_rows) continue;
        double acc[D];
#pragma unroll
        for (int i = 0; i < D; ++i) acc[i] = 0.0;
        const i64 base = slice_off[S];
        const int w = slice_w[S];
        for (int j = 0; j < w; ++j) {
            const i64 g = base + (i64)j * 32;
            const int cidx = __ldg(&col[g + lane]);
            double xv[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) xv[b] = __ldg(&x[(i64)cidx * NB + b]);
            const double* Au = Kuu + g * (D * D) + lane;
            const double* Ac = Kuc + g * D + lane;
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int b = 0; b < D; ++b) acc[i] += __ldcs(&Au[(i * D + b) * 32]) * xv[b];
                acc[i] += __ldcs(&Ac[i * 32]) * xv[D];
            }
        }
        if (r < n_rows) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double v = acc[i];
                if (lift) v += lift[(i64)r * D + i];
                if (fext) v -= fext[(i64)r * NB + i];
                out[(i64)r * out_stride + i] = v;
            }
        }
    }
}

__global__ void k_scatter_bc(const i64* __restrict__ dofs, const double* __restrict__ vals, i64 n, double* x) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t < n) x[dofs[t]] = vals[t];
}

template <typename T>
T* dalloc(i64 n, size_t& bytes) {
    T* p = nullptr;
    GL_CUDA(cudaMalloc(&p, sizeof(T) * (n > 0 ? n : 1)));
    bytes += sizeof(T) * (n > 0 ? n : 1);
    return p;
}

}  // namespace

void cc_free(glims_ctx* c) {
    CcMap* m = (CcMap*)c->ccmap;
    if (!m) return;
    for (void* q : {(void*)m->pl_off, (void*)m->pl_w, (void*)m->pe, (void*)m->pp, (void*)m->prv, (void*)m->ppa, (void*)m->rvol, (void*)m->Klin,
                    (void*)m->Mass, (void*)m->mcp, (void*)m->lift})
        if (q) cudaFree(q);
    delete m;
    c->ccmap = nullptr;
}

// Build the (row, element) pair lists on the device from the connectivity and the element->slot map.
static CcMap* cc_ensure(glims_ctx* c) {
    if (c->ccmap) return (CcMap*)c->ccmap;
    CcMap* m = new CcMap();
    c->ccmap = m;
    const auto& p = c->pat;
    const int nb = c->nb;
    if (p.max_w > CC_MAX_W) { m->why = "a row has more than 96 block columns"; return m; }
    if (c->n_c >= (1LL << 30) - 1 || c->n_v >= (1LL << 28)) { m->why = "mesh too large for the packed pair entries"; return m; }
    auto pol = thrust::cuda::par.on(c->stream);
    const i64 np = c->n_c * nb, n_rows = p.n_rows;
    thrust::device_vector<unsigned long long> keys(np);
    unsigned long long* kp = thrust::raw_pointer_cast(keys.data());
    k_pair_keys<<<nblk(np), 256, 0, c->stream>>>(c->cells, c->n_c, nb, n_rows, kp);
    thrust::sort(pol, keys.begin(), keys.end());
    const i64 nvalid = thrust::lower_bound(pol, keys.begin(), keys.end(), ~0ULL) - keys.begin();
    thrust::device_vector<i64> pstart(n_rows + 1);
    thrust::lower_bound(pol, keys.begin(), keys.begin() + nvalid,
                        thrust::make_transform_iterator(thrust::counting_iterator<i64>(0), PairRowStart()),
                        thrust::make_transform_iterator(thrust::counting_iterator<i64>(n_rows + 1), PairRowStart()),
                        pstart.begin());
    m->pl_w = dalloc<int>(p.n_slices, m->map_bytes);
    m->pl_off = dalloc<i64>(p.n_slices + 1, m->map_bytes);
    k_pair_width<<<nblk(p.n_slices), 256, 0, c->stream>>>(thrust::raw_pointer_cast(pstart.data()), p.n_rows, p.n_slices, m->pl_w);
    thrust::device_ptr<int> wp(m->pl_w);
    thrust::device_ptr<i64> op(m->pl_off);
    GL_CUDA(cudaMemsetAsync(m->pl_off, 0, sizeof(i64) * (p.n_slices + 1), c->stream));
    thrust::inclusive_scan(pol, wp, wp + p.n_slices, op + 1, thrust::plus<i64>());
    m->max_pw = thrust::reduce(pol, wp, wp + p.n_slices, 0, thrust::maximum<int>());
    k_times32<<<nblk(p.n_slices + 1), 256, 0, c->stream>>>(m->pl_off, p.n_slices + 1);
    GL_CUDA(cudaMemcpyAsync(&m->total, m->pl_off + p.n_slices, sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    m->pe = dalloc<unsigned>(m->total, m->map_bytes);
    m->pp = dalloc<unsigned>(m->total, m->map_bytes);
    GL_CUDA(cudaMemsetAsync(m->pe, 0xFF, sizeof(unsigned) * std::max<i64>(m->total, 1), c->stream));
    GL_CUDA(cudaMemsetAsync(m->pp, 0, sizeof(unsigned) * std::max<i64>(m->total, 1), c->stream));
    k_pair_fill<<<nblk(nvalid), 256, 0, c->stream>>>(kp, nvalid, nb, thrust::raw_pointer_cast(pstart.data()), m->pl_off,
                                                     p.slice_off, c->eslot, m->pe, m->pp);
    m->prv = dalloc<double>(m->total, m->map_bytes);
    m->ppa = dalloc<unsigned>(m->total, m->map_bytes);
    m->rvol = dalloc<double>(c->n_c, m->map_bytes);
    m->Klin = dalloc<double>(p.n_slots, m->map_bytes);
    m->Mass = dalloc<double>(p.n_slots, m->map_bytes);
    m->mcp = dalloc<double>(n_rows, m->map_bytes);
    GL_CUDA(cudaMemsetAsync(m->mcp, 0, sizeof(double) * std::max<i64>(n_rows, 1), c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    GL_CUDA(cudaGetLastError());
    m->ok = true;
    return m;
}

static size_t cc_smem(const glims_ctx* c) { return (size_t)(CC_TPB / 32) * 2 * std::max(c->pat.max_w, 1) * 32 * sizeof(double); }

bool cc_available(glims_ctx* c) { return cc_ensure(c)->ok; }
const char* cc_status(glims_ctx* c) { CcMap* m = (CcMap*)c->ccmap; return !m ? "not built" : m->ok ? "ok" : m->why.c_str(); }
void cc_invalidate_consts(glims_ctx* c) { if (c->ccmap) ((CcMap*)c->ccmap)->const_valid = false; }

// Klin, Mass, rvol for the current materials and dt
template <int D>
static void cc_consts_dim(glims_ctx* c, CcMap* m) {
    const auto& p = c->pat;
    const size_t smem = cc_smem(c);
    auto kfn = k_cc_setup<D>;
    GL_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rvol<D><<<nblk(c->n_c), 256, 0, c->stream>>>(c->coords, c->cells, c->cell_mat, c->mat, c->n_c, m->rvol);
    kfn<<<nblk(p.n_slices, CC_TPB / 32), CC_TPB, smem, c->stream>>>(p.slice_off, p.slice_w, p.n_slices, m->pl_off, m->pl_w,
        m->pe, m->pp, c->coords, c->cells, c->cell_mat, c->mat, c->dt, std::max(p.max_w, 1), m->Klin, m->Mass);
    k_pair_stream<<<nblk(m->total), 256, 0, c->stream>>>(m->pe, m->pp, m->total, m->rvol, m->prv, m->ppa);
    c->launches += 3;
}
static void cc_ensure_consts(glims_ctx* c, CcMap* m) {
    if (m->const_valid) return;
    if (c->dim == 2) cc_consts_dim<2>(c, m); else cc_consts_dim<3>(c, m);
    GL_CUDA(cudaGetLastError());
    m->const_valid = true;
}

// M c_prev of the step that is about to be solved (call after u_previous changed)
void cc_mass_cprev(glims_ctx* c) {
    CcMap* m = cc_ensure(c);
    if (!m->ok) return;
    cc_ensure_consts(c, m);
    const auto& p = c->pat;
    if (c->dim == 2) k_mass_cprev<2><<<nblk(p.n_slices * 32), 256, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, m->Mass, c->xprev, p.n_rows, m->mcp);
    else k_mass_cprev<3><<<nblk(p.n_slices * 32), 256, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, m->Mass, c->xprev, p.n_rows, m->mcp);
    c->launches++;
}

// F_c (with_res) and K_cc (with_kcc) from the current state; with_res needs cc_mass_cprev for the current u_previous
template <int D>
static void cc_rows_dim(glims_ctx* c, CcMap* m, bool with_kcc, bool with_res) {
    const auto& p = c->pat;
    const int g = nblk(p.n_slices, CC_TPB / 32);
    const double* fext = c->have_load ? c->fext : nullptr;
    const int mw = std::max(p.max_w, 1), mpw = std::max(m->max_pw, 4);
    // Bulk-copy variant (GLIMS_CC_TMA=1): measured SLOWER than the register-prefetch kernel at C4 (0.58 vs 0.33 ms): its
    // 74 KB per CTA leave 12 warps per SM, too few to cover the latency of the column / Klin / K_cc accesses that still go
    // through ordinary loads (profiles/r02_ncu_summary.md).  Kept selectable; the default is the register-prefetch kernel.
    const size_t smem_tma = (size_t)(CC_TPB / 32) * (2 * (size_t)mw * 256 + (size_t)mpw * 384);
    static const bool use_tma = [] { const char* e = std::getenv("GLIMS_CC_TMA"); return e && atoi(e) != 0; }();
    if (use_tma && smem_tma <= 100 * 1024) {
#define CC_TMA(WK) do { auto kfn = k_cc_rows_tma<D, WK>; \
        GL_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma)); \
        kfn<<<g, CC_TPB, smem_tma, c->stream>>>(p.slice_off, p.slice_w, p.col, p.n_rows, p.n_slices, m->pl_off, m->pl_w, m->prv, \
            m->ppa, m->Klin, m->mcp, fext, c->x, c->dt, mw, mpw, c->Kcc, with_res ? c->F : nullptr); } while (0)
        if (with_kcc) CC_TMA(true); else CC_TMA(false);
#undef CC_TMA
        c->launches++;
        return;
    }
    const size_t smem = cc_smem(c);
#define CC_GO(WK) do { auto kfn = k_cc_rows<D, WK>; \
        GL_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kfn<<<g, CC_TPB, smem, c->stream>>>(p.slice_off, p.slice_w, p.col, p.n_rows, p.n_slices, m->pl_off, m->pl_w, m->prv, \
            m->ppa, m->Klin, m->mcp, fext, c->x, c->dt, mw, c->Kcc, with_res ? c->F : nullptr); } while (0)
    if (with_kcc) CC_GO(true); else CC_GO(false);
#undef CC_GO
    c->launches++;
}
bool launch_cc_rows(glims_ctx* c, bool with_kcc, bool with_res) {
    CcMap* m = cc_ensure(c);
    if (!m->ok) return false;
    cc_ensure_consts(c, m);
    if (c->dim == 2) cc_rows_dim<2>(c, m, with_kcc, with_res); else cc_rows_dim<3>(c, m, with_kcc, with_res);
    GL_CUDA(cudaGetLastError());
    return true;
}

// F_u rows of c->F from the stored K_uu / K_uc (raw, or symmetrically eliminated + lift)
static void fu_launch(glims_ctx* c, const double* x, const double* lift, const double* fext, double* out, int stride) {
    const auto& p = c->pat;
    i64 need = (p.n_rows + 255) / 256;
    const int g = c->dim == 2 ? fit_grid(k_fu<2>, need, 256, 148 * 8) : fit_grid(k_fu<3>, need, 256, 148 * 8);
    if (c->dim == 2) k_fu<2><<<g, 256, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, c->Kuu, c->Kuc, x, p.n_rows, lift, fext, out, stride);
    else k_fu<3><<<g, 256, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, c->Kuu, c->Kuc, x, p.n_rows, lift, fext, out, stride);
    c->launches++;
    GL_CUDA(cudaGetLastError());
}
void launch_fu(glims_ctx* c, bool eliminated) {
    CcMap* m = (CcMap*)c->ccmap;
    fu_launch(c, c->x, (eliminated && m) ? m->lift : nullptr, c->have_load ? c->fext : nullptr, c->F, c->nb);
}

// lift = K_raw (Dirichlet values scattered into a zero vector), u rows.  Call while K_uu / K_uc are still raw.
void cc_compute_lift(glims_ctx* c, bool any_nonzero) {
    CcMap* m = cc_ensure(c);
    if (!any_nonzero || c->n_bc == 0) {
        if (m->lift) { cudaFree(m->lift); m->lift = nullptr; }
        return;
    }
    if (!m->lift) GL_CUDA(cudaMalloc(&m->lift, sizeof(double) * (i64)c->pat.n_rows * c->dim));
    double* tmp = c->dx;                       // scratch: the Newton update vector is dead between solves
    GL_CUDA(cudaMemsetAsync(tmp, 0, sizeof(double) * c->ndof, c->stream));
    k_scatter_bc<<<nblk(c->n_bc), 256, 0, c->stream>>>(c->bc_dofs, c->bc_vals, c->n_bc, tmp);
    c->launches++;
    fu_launch(c, tmp, nullptr, nullptr, m->lift, c->dim);
}

// the P1 mass matrix in the SELL slots of the vertex pattern (null when the pair lists cannot represent the mesh)
const double* cc_mass_matrix(glims_ctx* c) {
    CcMap* m = cc_ensure(c);
    if (!m->ok) return nullptr;
    cc_ensure_consts(c, m);
    return m->Mass;
}

i64 cc_map_bytes(glims_ctx* c) { CcMap* m = (CcMap*)c->ccmap; return m ? (i64)m->map_bytes : 0; }
