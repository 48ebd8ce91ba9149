// Device kernels of the hot path: element residual/Jacobian (K1/K2), Dirichlet elimination (K3),
// SELL-32 block SpMV (K4), block-Jacobi (K5), fused Krylov vector updates with deterministic
// shuffle reductions (K6).  sm_100a; FP64 throughout; no tensor cores (nothing here is a dense
// contraction) -- every kernel is bound by HBM traffic or FP64 issue.
#include "common.h"
#include "geom.cuh"

namespace {

constexpr int TPB = 256;
inline int nblk(i64 n, int t = TPB) { return (int)((n + t - 1) / t); }
inline int red_grid(glims_ctx* c, i64 n) {   // grid for grid-stride reduction kernels
    i64 need = (n + TPB - 1) / TPB;
    i64 cap = 148 * 8;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}
constexpr int MAXBLK = 148 * 8;

// ------------------------------------------------------------------------------------------------
// deterministic grid reduction: per-block shuffle reduce -> partials[blk]; the last block to arrive
// sums the partials in a fixed order and publishes the scalar.
template <int NV>
__device__ inline void grid_reduce(double (&v)[NV], double* partials, unsigned* tickets, double* scal, int slot0) {
    __shared__ double sm[NV][TPB / 32];
    __shared__ bool last;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[k][wid] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < TPB / 32; ++w) s += sm[threadIdx.x][w];
        partials[(i64)(slot0 + threadIdx.x) * MAXBLK + blockIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicInc(&tickets[slot0], gridDim.x - 1);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        for (int k = 0; k < NV; ++k) {
            double s = 0;
            for (int b = threadIdx.x; b < (int)gridDim.x; b += TPB)
                s += __ldcg(&partials[(i64)(slot0 + k) * MAXBLK + b]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            __syncthreads();
            if (lane == 0) sm[0][wid] = s;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0;
                for (int w = 0; w < TPB / 32; ++w) t += sm[0][w];
                scal[slot0 + k] = t;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 + K2, element-parallel variant: one thread per element, scatter through the precomputed
// element->slot map with RED.ADD.F64.  RES: residual; KCONST: K_uu, K_uc; KCC: K_cc.
template <int D, bool RES, bool KCONST, bool KCC>
__global__ void __launch_bounds__(128)
k_assemble_atomic(const double* __restrict__ coords, const int* __restrict__ cells, const int* __restrict__ cell_mat,
                  const double* __restrict__ mat_g, int n_mat, i64 n_c, i64 n_own, double dt,
                  const double* __restrict__ x, const double* __restrict__ xprev, const int* __restrict__ eslot,
                  double* __restrict__ F, double* __restrict__ Kuu, double* __restrict__ Kuc, double* __restrict__ Kcc) {
    constexpr int NB = D + 1;
    __shared__ double smat[MAX_MAT * MAT_STRIDE];
    for (int t = threadIdx.x; t < n_mat * MAT_STRIDE; t += blockDim.x) smat[t] = mat_g[t];
    __syncthreads();
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= n_c) return;
    int v[NB];
#pragma unroll
    for (int a = 0; a < NB; ++a) v[a] = cells[e * NB + a];
    double X[NB][D];
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int k = 0; k < D; ++k) X[a][k] = coords[(i64)v[a] * D + k];
    Geo<D> G;
    geometry(X, G);
    const double* m = &smat[cell_mat[e] * MAT_STRIDE];
    const double mu = m[0], lam = m[1], Dc = m[2], rho = m[3], beta = m[5];
    const double vol = G.vol;
    double cv[NB], S = 0, Q = 0;
#pragma unroll
    for (int a = 0; a < NB; ++a) { cv[a] = x[(i64)v[a] * NB + D]; S += cv[a]; Q += cv[a] * cv[a]; }

    if (RES) {
        double u[NB][D], cp[NB], Sp = 0;
#pragma unroll
        for (int a = 0; a < NB; ++a) {
#pragma unroll
            for (int k = 0; k < D; ++k) u[a][k] = x[(i64)v[a] * NB + k];
            cp[a] = xprev[(i64)v[a] * NB + D];
            Sp += cp[a];
        }
        double gu[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double s = 0;
#pragma unroll
                for (int b = 0; b < NB; ++b) s += u[b][i] * G.g[b][j];
                gu[i][j] = s;
            }
        double tr = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) tr += gu[i][i];
        double sig[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) sig[i][j] = mu * (gu[i][j] + gu[j][i]) + (i == j ? lam * tr : 0.0);
        const double cbar = S / NB;
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            if (v[a] >= n_own) continue;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double s = 0;
#pragma unroll
                for (int j = 0; j < D; ++j) s += sig[i][j] * G.g[a][j];
                atomicAdd(&F[(i64)v[a] * NB + i], vol * (s - beta * cbar * G.g[a][i]));
            }
            double diff = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                double gg = 0;
#pragma unroll
                for (int k = 0; k < D; ++k) gg += G.g[a][k] * G.g[b][k];
                diff += gg * cv[b];
            }
            double massc = Consts<D>::mass * (cv[a] + S), massp = Consts<D>::mass * (cp[a] + Sp);
            double trip = Consts<D>::kappa * (2.0 * cv[a] * cv[a] + 2.0 * cv[a] * S + S * S + Q);
            atomicAdd(&F[(i64)v[a] * NB + D], vol * (massc - massp + dt * Dc * diff - dt * rho * (massc - trip)));
        }
    }
    if (KCONST || KCC) {
#pragma unroll
        for (int a = 0; a < NB; ++a) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int s = eslot[e * (NB * NB) + a * NB + b];
                if (s < 0) continue;
                double gg = 0;
#pragma unroll
                for (int k = 0; k < D; ++k) gg += G.g[a][k] * G.g[b][k];
                if (KCONST) {
#pragma unroll
                    for (int i = 0; i < D; ++i) {
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            double val = vol * (mu * ((i == j ? gg : 0.0) + G.g[a][j] * G.g[b][i]) + lam * G.g[a][i] * G.g[b][j]);
                            atomicAdd(&Kuu[vidx(s, i * D + j, D * D)], val);
                        }
                        atomicAdd(&Kuc[vidx(s, i, D)], -beta * vol * (1.0 / NB) * G.g[a][i]);
                    }
                }
                if (KCC) {
                    double ml = Consts<D>::mass * (a == b ? 2.0 : 1.0);
                    double tl = 2.0 * Consts<D>::kappa * (a == b ? 4.0 * cv[a] + 2.0 * S : cv[a] + cv[b] + S);
                    atomicAdd(&Kcc[s], vol * (ml * (1.0 - dt * rho) + dt * Dc * gg + dt * rho * tl));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2, row-parallel gather variant: one thread per matrix slot walks the transposed scatter map
// (the (element, a, b) triples that contribute to the slot), recomputes each element's gradients
// from L2-resident coordinates and writes every matrix value exactly once -- no atomics, no
// pre-zeroing, deterministic summation order.
template <int D, bool KCONST, bool KCC>
__global__ void __launch_bounds__(128)
k_assemble_gather(const double* __restrict__ coords, const int* __restrict__ cells, const int* __restrict__ cell_mat,
                  const double* __restrict__ mat_g, int n_mat, double dt, const double* __restrict__ x,
                  const i64* __restrict__ gptr, const int* __restrict__ gent, i64 n_slots,
                  double* __restrict__ Kuu, double* __restrict__ Kuc, double* __restrict__ Kcc) {
    constexpr int NB = D + 1;
    __shared__ double smat[MAX_MAT * MAT_STRIDE];
    for (int t = threadIdx.x; t < n_mat * MAT_STRIDE; t += blockDim.x) smat[t] = mat_g[t];
    __syncthreads();
    i64 s = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    double auu[D * D], auc[D], acc = 0;
#pragma unroll
    for (int k = 0; k < D * D; ++k) auu[k] = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) auc[k] = 0;
    const i64 t0 = gptr[s], t1 = gptr[s + 1];
    for (i64 t = t0; t < t1; ++t) {
        const int ent = gent[t];
        const i64 e = ent >> 4;
        const int a = (ent >> 2) & 3, b = ent & 3;
        int v[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) v[q] = cells[e * NB + q];
        double X[NB][D];
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int k = 0; k < D; ++k) X[q][k] = coords[(i64)v[q] * D + k];
        Geo<D> G;
        geometry(X, G);
        const double* m = &smat[cell_mat[e] * MAT_STRIDE];
        const double mu = m[0], lam = m[1], Dc = m[2], rho = m[3], beta = m[5];
        double ga[D], gb[D], gg = 0;
#pragma unroll
        for (int q = 0; q < NB; ++q) {   // select without dynamic register indexing
            if (q == a) {
#pragma unroll
                for (int k = 0; k < D; ++k) ga[k] = G.g[q][k];
            }
            if (q == b) {
#pragma unroll
                for (int k = 0; k < D; ++k) gb[k] = G.g[q][k];
            }
        }
#pragma unroll
        for (int k = 0; k < D; ++k) gg += ga[k] * gb[k];
        if (KCONST) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int j = 0; j < D; ++j)
                    auu[i * D + j] += G.vol * (mu * ((i == j ? gg : 0.0) + ga[j] * gb[i]) + lam * ga[i] * gb[j]);
                auc[i] += -beta * G.vol * (1.0 / NB) * ga[i];
            }
        }
        if (KCC) {
            double S = 0, ca = 0, cb = 0;
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                double cq = x[(i64)v[q] * NB + D];
                S += cq;
                if (q == a) ca = cq;
                if (q == b) cb = cq;
            }
            double ml = Consts<D>::mass * (a == b ? 2.0 : 1.0);
            double tl = 2.0 * Consts<D>::kappa * (a == b ? 4.0 * ca + 2.0 * S : ca + cb + S);
            acc += G.vol * (ml * (1.0 - dt * rho) + dt * Dc * gg + dt * rho * tl);
        }
    }
    if (KCONST) {
#pragma unroll
        for (int k = 0; k < D * D; ++k) Kuu[vidx(s, k, D * D)] = auu[k];
#pragma unroll
        for (int k = 0; k < D; ++k) Kuc[vidx(s, k, D)] = auc[k];
    }
    if (KCC) Kcc[s] = acc;
}

// K2, slice-local variant: one CTA per SELL slice (32 block rows).  Phase A stages, in shared memory, the
// gradients / volume / material / c-sum of every element touching the slice (each element's geometry is
// computed once per slice it touches, ~2x overall instead of 16x in the plain gather kernel).  Phase B: one
// thread per matrix slot walks its contributors -- 16-bit entries (local element, a, b) -- reading only shared
// memory, and writes each value exactly once.  No atomics, no zero-fill, deterministic summation order.
template <int D, bool KCONST, bool KCC>
__global__ void __launch_bounds__(256)
k_assemble_slice(const double* __restrict__ coords, const int* __restrict__ cells, const int* __restrict__ cell_mat,
                 const double* __restrict__ mat_g, int n_mat, double dt, const double* __restrict__ x,
                 const i64* __restrict__ sl_ptr, const int* __restrict__ sl_elem, const i64* __restrict__ gptr,
                 const unsigned short* __restrict__ lent, const i64* __restrict__ slice_off,
                 const int* __restrict__ slice_w, const int* __restrict__ col, int n_rows,
                 double* __restrict__ Kuu, double* __restrict__ Kuc, double* __restrict__ Kcc) {
    constexpr int NB = D + 1, REC = NB * D + 3;     // gradients, |K|, material index, sum of c
    extern __shared__ double sm[];
    __shared__ double smat[MAX_MAT * MAT_STRIDE];
    for (int t = threadIdx.x; t < n_mat * MAT_STRIDE; t += blockDim.x) smat[t] = mat_g[t];
    const int S = blockIdx.x;
    const i64 p0 = sl_ptr[S];
    const int n_el = (int)(sl_ptr[S + 1] - p0);
    for (int i = threadIdx.x; i < n_el; i += blockDim.x) {
        const i64 e = sl_elem[p0 + i];
        int v[NB];
#pragma unroll
        for (int a = 0; a < NB; ++a) v[a] = cells[e * NB + a];
        double X[NB][D];
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int k = 0; k < D; ++k) X[a][k] = coords[(i64)v[a] * D + k];
        Geo<D> G;
        geometry(X, G);
        double* o = sm + i * REC;
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int k = 0; k < D; ++k) o[a * D + k] = G.g[a][k];
        o[NB * D] = G.vol;
        o[NB * D + 1] = (double)cell_mat[e];
        if (KCC) {
            double sc = 0;
#pragma unroll
            for (int a = 0; a < NB; ++a) sc += x[(i64)v[a] * NB + D];
            o[NB * D + 2] = sc;
        }
    }
    __syncthreads();
    const i64 base = slice_off[S];
    const int w = slice_w[S];
    for (int tt = threadIdx.x; tt < w * 32; tt += blockDim.x) {
        const i64 s = base + tt;
        double auu[D * D], auc[D], acc = 0, wsum = 0;
#pragma unroll
        for (int k = 0; k < D * D; ++k) auu[k] = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) auc[k] = 0;
        const i64 t0 = gptr[s], t1 = gptr[s + 1];
        bool diag = false;
        for (i64 t = t0; t < t1; ++t) {
            const int ent = lent[t];
            const int a = (ent >> 2) & 3, b = ent & 3;
            const double* g = sm + (ent >> 4) * REC;
            diag = (a == b);
            double ga[D], gb[D], gg = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) { ga[k] = g[a * D + k]; gb[k] = g[b * D + k]; gg += ga[k] * gb[k]; }
            const double vol = g[NB * D];
            const double* m = &smat[(int)g[NB * D + 1] * MAT_STRIDE];
            if (KCONST) {
                const double mu = m[0] * vol, lam = m[1] * vol, bv = -m[5] * vol * (1.0 / NB);
#pragma unroll
                for (int i = 0; i < D; ++i) {
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        auu[i * D + j] += mu * ((i == j ? gg : 0.0) + ga[j] * gb[i]) + lam * ga[i] * gb[j];
                    auc[i] += bv * ga[i];
                }
            }
            if (KCC) {
                const double rho = m[3], Dc = m[2];
                acc += vol * (Consts<D>::mass * (a == b ? 2.0 : 1.0) * (1.0 - dt * rho) + dt * Dc * gg);
                const double wr = 2.0 * dt * rho * Consts<D>::kappa * vol;
                wsum += wr;
                acc += wr * (a == b ? 2.0 : 1.0) * g[NB * D + 2];
            }
        }
        if (KCONST) {
#pragma unroll
            for (int k = 0; k < D * D; ++k) Kuu[vidx(s, k, D * D)] = auu[k];
#pragma unroll
            for (int k = 0; k < D; ++k) Kuc[vidx(s, k, D)] = auc[k];
        }
        if (KCC) {
            const int r = S * 32 + (tt & 31);
            if (t1 > t0 && r < n_rows) {
                const double ca = x[(i64)r * NB + D], cb = x[(i64)col[s] * NB + D];
                acc += wsum * (diag ? 4.0 * ca : ca + cb);
            }
            Kcc[s] = acc;
        }
    }
}

// N2 derived fields, one thread per cell (P1 displacement => strain and stress are constant per cell):
// out[e][*] = strain (d*d), stress (d*d), pressure = tr(sigma)/3, von Mises, det(I + grad u),
// det(I + cbar*gamma*I), logistic growth rho*cbar*(1-cbar)   (math_linear_elasticity.py:12-46, math_reaction_diffusion.py:2-3)
template <int D>
__global__ void k_cell_fields(const double* __restrict__ coords, const int* __restrict__ cells,
                              const int* __restrict__ cell_mat, const double* __restrict__ mat_g, int n_mat, i64 n_c,
                              const double* __restrict__ x, double* __restrict__ out, double* __restrict__ vol_out) {
    constexpr int NB = D + 1, NF = 2 * D * D + 5;
    __shared__ double smat[MAX_MAT * MAT_STRIDE];
    for (int t = threadIdx.x; t < n_mat * MAT_STRIDE; t += blockDim.x) smat[t] = mat_g[t];
    __syncthreads();
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= n_c) return;
    int v[NB];
    double X[NB][D];
#pragma unroll
    for (int a = 0; a < NB; ++a) {
        v[a] = cells[e * NB + a];
#pragma unroll
        for (int k = 0; k < D; ++k) X[a][k] = coords[(i64)v[a] * D + k];
    }
    Geo<D> G;
    geometry(X, G);
    const double* m = &smat[cell_mat[e] * MAT_STRIDE];
    const double mu = m[0], lam = m[1], rho = m[3], gam = m[4];
    double gu[D][D], cbar = 0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) gu[i][j] = 0;
#pragma unroll
    for (int a = 0; a < NB; ++a) {
        cbar += x[(i64)v[a] * NB + D] * (1.0 / NB);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double ui = x[(i64)v[a] * NB + i];
#pragma unroll
            for (int j = 0; j < D; ++j) gu[i][j] += ui * G.g[a][j];
        }
    }
    double* o = out + e * NF;
    double tr = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) tr += gu[i][i];
    double sig[D][D], trs = 0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double eps = 0.5 * (gu[i][j] + gu[j][i]);
            o[i * D + j] = eps;
            sig[i][j] = 2.0 * mu * eps + (i == j ? lam * tr : 0.0);
            o[D * D + i * D + j] = sig[i][j];
        }
#pragma unroll
    for (int i = 0; i < D; ++i) trs += sig[i][i];
    o[2 * D * D] = trs * (1.0 / 3.0);                                  // compute_pressure_from_stress_tensor: 1/3 tr
    double dev2 = 0;                                                   // deviator uses 1/3 tr I_d (mle:35-36) in 2D too
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double dv = sig[i][j] - (i == j ? trs * (1.0 / 3.0) : 0.0);
            dev2 += dv * dv;
        }
    o[2 * D * D + 1] = sqrt(1.5 * dev2);
    double F[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) F[i][j] = gu[i][j] + (i == j ? 1.0 : 0.0);
    double detF;
    if (D == 2) detF = F[0][0] * F[1][1] - F[0][1] * F[1][0];
    else detF = F[0][0] * (F[1][1] * F[2 % D][2 % D] - F[1][2 % D] * F[2 % D][1]) - F[0][1] * (F[1][0] * F[2 % D][2 % D] - F[1][2 % D] * F[2 % D][0])
              + F[0][2 % D] * (F[1][0] * F[2 % D][1] - F[1][1] * F[2 % D][0]);
    o[2 * D * D + 2] = detF;
    double jg = 1.0;
#pragma unroll
    for (int i = 0; i < D; ++i) jg *= (1.0 + cbar * gam);
    o[2 * D * D + 3] = jg;
    o[2 * D * D + 4] = rho * cbar * (1.0 - cbar);
    vol_out[e] = G.vol;
}

// Right-hand side of the consistent-mass L2 projection onto P1 of the derived fields (the reference's `project(expr, V)`,
// helper_classes.py:1560-1618): load[v][f] = int f(x) phi_v dx.  Fields that are constant per cell (strain, stress, pressure,
// von Mises, det(I + grad u)) give |K|/(d+1) f; the two that are polynomials of the P1 concentration -- det(I + c gamma I) =
// (1 + gamma c)^d and rho c (1 - c) -- are integrated exactly with int lambda^alpha = |K| d! alpha! / (|alpha| + d)!.
__device__ inline double fact_prod4(int a, int b, int c, int e) {        // prod_i (multiplicity of i)!  over the given indices (-1: unused)
    const int idx[4] = {a, b, c, e};
    double f = 1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (idx[i] < 0) continue;
        int m = 1;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < i && idx[j] == idx[i]) ++m;       // this occurrence is the m-th of its value: contributes the factor m
        f *= (double)m;
    }
    return f;
}
template <int D>
__global__ void k_project_load(const int* __restrict__ cells, i64 n_c, const double* __restrict__ q, const double* __restrict__ vol,
                               const int* __restrict__ cell_mat, const double* __restrict__ mat_g, const double* __restrict__ x,
                               double* load) {
    constexpr int NB = D + 1, NF = 2 * D * D + 5;
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= n_c) return;
    int v[NB];
    double cv[NB];
#pragma unroll
    for (int a = 0; a < NB; ++a) { v[a] = cells[e * NB + a]; cv[a] = x[(i64)v[a] * NB + D]; }
    const double w = vol[e], rho = mat_g[cell_mat[e] * MAT_STRIDE + 3], gam = mat_g[cell_mat[e] * MAT_STRIDE + 4];
    // d! / (k + 1 + d)!  for k = 0..3 (k = polynomial degree in c, +1 for the test function)
    double fac[4];
    {
        double dfact = (D == 2) ? 2.0 : 6.0, den = dfact;           // d!
#pragma unroll
        for (int k = 0; k < 4; ++k) { den *= (double)(k + 1 + D); fac[k] = dfact / den; }
    }
#pragma unroll
    for (int a = 0; a < NB; ++a) {
        double I0 = w * fac[0], I1 = 0, I2 = 0, I3 = 0;               // int lambda_a c^k
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            I1 += cv[b] * fact_prod4(a, b, -1, -1);
#pragma unroll
            for (int c2 = 0; c2 < NB; ++c2) {
                I2 += cv[b] * cv[c2] * fact_prod4(a, b, c2, -1);
                if (D == 3) {
#pragma unroll
                    for (int e3 = 0; e3 < NB; ++e3) I3 += cv[b] * cv[c2] * cv[e3] * fact_prod4(a, b, c2, e3);
                }
            }
        }
        I1 *= w * fac[1]; I2 *= w * fac[2]; I3 *= w * fac[3];
        double* o = load + (i64)v[a] * NF;
        for (int f = 0; f < 2 * D * D + 3; ++f) atomicAdd(&o[f], I0 * q[e * NF + f]);
        const double jg = (D == 2) ? I0 + 2.0 * gam * I1 + gam * gam * I2
                                   : I0 + 3.0 * gam * I1 + 3.0 * gam * gam * I2 + gam * gam * gam * I3;
        atomicAdd(&o[2 * D * D + 3], jg);
        atomicAdd(&o[2 * D * D + 4], rho * (I1 - I2));
    }
}
__global__ void k_strided_copy(const double* __restrict__ src, i64 n, int stride_src, int off_src, double* __restrict__ dst,
                               int stride_dst, int off_dst) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB)
        dst[i * stride_dst + off_dst] = src[i * stride_src + off_src];
}
__global__ void k_scalar_diag_inverse(const int* __restrict__ diag, int n_rows, const double* __restrict__ A, double* out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) out[r] = 1.0 / A[diag[r]];
}

// volume-weighted nodal average of per-cell fields: num[v][f] += vol*q, den[v] += vol
__global__ void k_cell_to_vertex(const int* __restrict__ cells, int nb, i64 n_c, int nf, const double* __restrict__ q,
                                 const double* __restrict__ vol, double* num, double* den) {
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= n_c) return;
    const double w = vol[e];
    for (int a = 0; a < nb; ++a) {
        const i64 v = cells[e * nb + a];
        atomicAdd(&den[v], w);
        for (int f = 0; f < nf; ++f) atomicAdd(&num[v * nf + f], w * q[e * nf + f]);
    }
}
__global__ void k_divide_rows(double* num, const double* __restrict__ den, i64 n_v, int nf) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n_v * nf) return;
    const double d = den[t / nf];
    num[t] = d > 0 ? num[t] / d : 0.0;
}

// ------------------------------------------------------------------------------------------------
// K3 Dirichlet
__global__ void k_bc_values(const i64* __restrict__ dofs, const double* __restrict__ vals, i64 n, double* x) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t < n) x[dofs[t]] = vals[t];
}
__global__ void k_bc_residual(const i64* __restrict__ dofs, const double* __restrict__ vals, i64 n,
                              const double* __restrict__ x, double* F) {
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t < n) F[dofs[t]] = x[dofs[t]] - vals[t];
}
// one block per slice: zero Dirichlet rows (and, if sym, columns) of the requested matrices.
// bcmask[v] has bit k set when dof (v,k) is constrained.
template <int D>
__global__ void k_bc_matrix(const unsigned char* __restrict__ bcmask, const i64* __restrict__ slice_off,
                            const int* __restrict__ slice_w, const int* __restrict__ col, int n_rows, int what, bool sym,
                            double* Kuu, double* Kuc, double* Kcc) {
    const int S = blockIdx.x;
    const i64 base = slice_off[S];
    const int w = slice_w[S];
    for (int t = threadIdx.x; t < w * 32; t += blockDim.x) {
        const int r = S * 32 + (t & 31);
        if (r >= n_rows) continue;
        const i64 s = base + t;
        const unsigned mr = bcmask[r];
        const unsigned mc = sym ? bcmask[col[s]] : 0u;
        if ((mr | mc) == 0u) continue;
        if (what & GLIMS_ASM_KCONST) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (((mr >> i) | (mc >> j)) & 1u) Kuu[vidx(s, i * D + j, D * D)] = 0.0;
                if (((mr >> i) | (mc >> D)) & 1u) Kuc[vidx(s, i, D)] = 0.0;
            }
        }
        if ((what & GLIMS_ASM_KCC) && (((mr | mc) >> D) & 1u)) Kcc[s] = 0.0;
    }
}
template <int D>
__global__ void k_bc_diag(const i64* __restrict__ dofs, i64 n, i64 n_own, const int* __restrict__ diag, int what,
                          double* Kuu, double* Kcc) {
    constexpr int NB = D + 1;
    i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (t >= n) return;
    i64 dof = dofs[t];
    int v = (int)(dof / NB), k = (int)(dof - (i64)v * NB);
    if (v >= n_own) return;
    i64 s = diag[v];
    if (k < D) { if (what & GLIMS_ASM_KCONST) Kuu[vidx(s, k * D + k, D * D)] = 1.0; }
    else if (what & GLIMS_ASM_KCC) Kcc[s] = 1.0;
}

// ------------------------------------------------------------------------------------------------
// K4 SpMV, SELL-32: one thread per block row, slice width uniform per warp, value loads coalesced
// (component-major inside each 32-slot group), x gathered through L1/L2.
template <int BR, int BC, bool DOT, bool RESID = false>
__global__ void __launch_bounds__(TPB)
k_spmv_block(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
             const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y, int n_rows,
             const double* __restrict__ wdot, double* partials, unsigned* tickets, double* scal, int slot) {
    // RESID: y = wdot - A x  (wdot doubles as the right-hand side; no dot product in that mode)
    double local[1] = {0.0};
    const int n_tiles = (n_rows + TPB - 1) / TPB;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r = tile * TPB + threadIdx.x;
        const int S = r >> 5, lane = r & 31;
        double acc[BR];
#pragma unroll
        for (int i = 0; i < BR; ++i) acc[i] = 0.0;
        if (S * 32 < n_rows) {
            const i64 base = slice_off[S];
            const int w = slice_w[S];
            for (int j = 0; j < w; ++j) {
                const i64 g = base + (i64)j * 32;
                const int cidx = __ldg(&col[g + lane]);
                double xv[BC];
#pragma unroll
                for (int b = 0; b < BC; ++b) xv[b] = __ldg(&x[(i64)cidx * BC + b]);
                const double* Ag = A + g * (BR * BC) + lane;
#pragma unroll
                for (int i = 0; i < BR; ++i)
#pragma unroll
                    for (int b = 0; b < BC; ++b) acc[i] += __ldcs(&Ag[(i * BC + b) * 32]) * xv[b];
            }
        }
        if (r < n_rows) {
            if (RESID) {
#pragma unroll
                for (int i = 0; i < BR; ++i) y[(i64)r * BR + i] = wdot[(i64)r * BR + i] - acc[i];
            } else {
#pragma unroll
                for (int i = 0; i < BR; ++i) y[(i64)r * BR + i] = acc[i];
            }
            if (DOT) {
#pragma unroll
                for (int i = 0; i < BR; ++i) local[0] += acc[i] * wdot[(i64)r * BR + i];
            }
        }
    }
    if (DOT) grid_reduce<1>(local, partials, tickets, scal, slot);
}

// Coarse AMG levels: few rows, fat blocks. One CTA per slice, warp i computes block-row component i of the
// slice's 32 rows, so a 6x6 level exposes 6x the parallelism of the thread-per-row kernel while every value
// load stays a coalesced 256-byte warp access.  RESID: y = rhs - A x.
template <int BS, bool RESID>
__global__ void __launch_bounds__(32 * BS)
k_spmv_split(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
             const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y, int n_rows,
             int n_slices, const double* __restrict__ rhs) {
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
    for (int S = blockIdx.x; S < n_slices; S += gridDim.x) {
        const int r = S * 32 + lane;
        const i64 base = slice_off[S];
        const int w = slice_w[S];
        double acc0 = 0.0, acc1 = 0.0;
        int j = 0;
        for (; j + 1 < w; j += 2) {      // two blocks in flight per thread
            const i64 g0 = base + (i64)j * 32, g1 = g0 + 32;
            const int c0 = __ldg(&col[g0 + lane]), c1 = __ldg(&col[g1 + lane]);
            const double* A0 = A + g0 * (BS * BS) + (i * BS) * 32 + lane;
            const double* A1 = A + g1 * (BS * BS) + (i * BS) * 32 + lane;
#pragma unroll
            for (int b = 0; b < BS; ++b) {
                acc0 += __ldcs(&A0[b * 32]) * __ldg(&x[(i64)c0 * BS + b]);
                acc1 += __ldcs(&A1[b * 32]) * __ldg(&x[(i64)c1 * BS + b]);
            }
        }
        if (j < w) {
            const i64 g0 = base + (i64)j * 32;
            const int c0 = __ldg(&col[g0 + lane]);
            const double* A0 = A + g0 * (BS * BS) + (i * BS) * 32 + lane;
#pragma unroll
            for (int b = 0; b < BS; ++b) acc0 += __ldcs(&A0[b * 32]) * __ldg(&x[(i64)c0 * BS + b]);
        }
        if (r < n_rows) {
            const double v = acc0 + acc1;
            y[(i64)r * BS + i] = RESID ? rhs[(i64)r * BS + i] - v : v;
        }
    }
}

// monolithic Jacobian J = [[K_uu, K_uc], [0, K_cc]] on vertex-blocked vectors [n_v][D+1]
template <int D, bool DOT>
__global__ void __launch_bounds__(TPB)
k_spmv_mono(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
            const double* __restrict__ Kuu, const double* __restrict__ Kuc, const double* __restrict__ Kcc,
            const double* __restrict__ x, double* __restrict__ y, int n_rows, const double* __restrict__ wdot,
            double* partials, unsigned* tickets, double* scal, int slot) {
    constexpr int NB = D + 1;
    double local[1] = {0.0};
    const int n_tiles = (n_rows + TPB - 1) / TPB;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r = tile * TPB + threadIdx.x;
        const int S = r >> 5, lane = r & 31;
        double acc[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) acc[i] = 0.0;
        if (S * 32 < n_rows) {
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
                acc[D] += __ldcs(&Kcc[g + lane]) * xv[D];
            }
        }
        if (r < n_rows) {
#pragma unroll
            for (int i = 0; i < NB; ++i) y[(i64)r * NB + i] = acc[i];
            if (DOT) {
#pragma unroll
                for (int i = 0; i < NB; ++i) local[0] += acc[i] * wdot[(i64)r * NB + i];
            }
        }
    }
    if (DOT) grid_reduce<1>(local, partials, tickets, scal, slot);
}

// y_u = K_uc x_c  (D x 1 blocks)
template <int D>
__global__ void __launch_bounds__(TPB)
k_spmv_uc(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
          const double* __restrict__ Kuc, const double* __restrict__ xc, double* __restrict__ yu, int n_rows) {
    const int r = blockIdx.x * TPB + threadIdx.x;
    const int S = r >> 5, lane = r & 31;
    if (S * 32 >= n_rows) return;
    double acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.0;
    const i64 base = slice_off[S];
    const int w = slice_w[S];
    for (int j = 0; j < w; ++j) {
        const i64 g = base + (i64)j * 32;
        const double xv = __ldg(&xc[col[g + lane]]);
#pragma unroll
        for (int i = 0; i < D; ++i) acc[i] += Kuc[g * D + i * 32 + lane] * xv;
    }
    if (r < n_rows)
#pragma unroll
        for (int i = 0; i < D; ++i) yu[(i64)r * D + i] = acc[i];
}

// ------------------------------------------------------------------------------------------------
// diagonal-block inverses (K5 block-Jacobi data)
template <int N>
__device__ inline void invert_small(double (&a)[N][N]) {   // Gauss-Jordan without pivoting (SPD / diagonally safe blocks)
    double inv[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) inv[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int p = 0; p < N; ++p) {
        double ip = 1.0 / a[p][p];
#pragma unroll
        for (int j = 0; j < N; ++j) { a[p][j] *= ip; inv[p][j] *= ip; }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i == p) continue;
            double f = a[i][p];
#pragma unroll
            for (int j = 0; j < N; ++j) { a[i][j] -= f * a[p][j]; inv[i][j] -= f * inv[p][j]; }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) a[i][j] = inv[i][j];
}

template <int D>
__global__ void k_diag_inverse(const int* __restrict__ diag, int n_rows, int which, const double* __restrict__ Kuu,
                               const double* __restrict__ Kuc, const double* __restrict__ Kcc, double* out) {
    constexpr int NB = D + 1;
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    i64 s = diag[r];
    if (which == 2) { out[r] = 1.0 / Kcc[s]; return; }
    if (which == 1) {
        double a[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) a[i][j] = Kuu[vidx(s, i * D + j, D * D)];
        invert_small<D>(a);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) out[(i64)r * D * D + i * D + j] = a[i][j];
        return;
    }
    double a[NB][NB];
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < D; ++j) a[i][j] = Kuu[vidx(s, i * D + j, D * D)];
        a[i][D] = Kuc[vidx(s, i, D)];
        a[D][i] = 0.0;
    }
    a[D][D] = Kcc[s];
    invert_small<NB>(a);
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) out[(i64)r * NB * NB + i * NB + j] = a[i][j];
}

template <int BS>
__global__ void __launch_bounds__(TPB)
k_block_jacobi(const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ z, i64 n_rows,
               double* partials, unsigned* tickets, double* scal, int slot) {
    double local[1] = {0.0};
    for (i64 row = blockIdx.x * (i64)TPB + threadIdx.x; row < n_rows; row += (i64)gridDim.x * TPB) {
        double rv[BS], zv[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) rv[i] = r[row * BS + i];
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            double s = 0;
#pragma unroll
            for (int j = 0; j < BS; ++j) s += dinv[row * BS * BS + i * BS + j] * rv[j];
            zv[i] = s;
            z[row * BS + i] = s;
        }
#pragma unroll
        for (int i = 0; i < BS; ++i) local[0] += rv[i] * zv[i];
    }
    if (slot >= 0) grid_reduce<1>(local, partials, tickets, scal, slot);
}

// ------------------------------------------------------------------------------------------------
// K6 vector kernels
__global__ void __launch_bounds__(TPB)
k_dot(const double* __restrict__ a, const double* __restrict__ b, i64 n, double* partials, unsigned* tickets,
      double* scal, int slot) {
    double local[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) local[0] += a[i] * b[i];
    grid_reduce<1>(local, partials, tickets, scal, slot);
}
// h[i] = V_i . w for i < k (k <= 32), one pass over w
__global__ void __launch_bounds__(TPB)
k_multi_dot(const double* __restrict__ V, i64 ld, int k, const double* __restrict__ w, i64 n, double* partials,
            unsigned* tickets, double* scal, int slot0) {
    for (int i0 = 0; i0 < k; i0 += 4) {
        double local[4] = {0, 0, 0, 0};
        for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) {
            double wi = w[i];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (i0 + q < k) local[q] += V[(i64)(i0 + q) * ld + i] * wi;
        }
        grid_reduce<4>(local, partials, tickets, scal, slot0 + i0);
        __syncthreads();
    }
}
__global__ void k_multi_axpy(const double* __restrict__ V, i64 ld, int k, const double* __restrict__ coef, double sign,
                             double* __restrict__ w, i64 n) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) {
        double s = w[i];
        for (int q = 0; q < k; ++q) s += sign * coef[q] * V[(i64)q * ld + i];
        w[i] = s;
    }
}
__global__ void k_axpy(double alpha, const double* __restrict__ x, double* __restrict__ y, i64 n) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) y[i] += alpha * x[i];
}
__global__ void k_scale(double alpha, double* __restrict__ x, i64 n) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) x[i] *= alpha;
}
// x += a p ; r -= a Ap ; rr = r.r   with a = scal[num]/scal[den]
__global__ void __launch_bounds__(TPB)
k_cg_update_xr(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
               const double* __restrict__ Ap, i64 n, int s_num, int s_den, int s_rr, double* partials,
               unsigned* tickets, double* scal) {
    const double den = scal[s_den];
    const double a = den != 0.0 ? scal[s_num] / den : 0.0;
    double local[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) {
        x[i] += a * p[i];
        double ri = r[i] - a * Ap[i];
        r[i] = ri;
        local[0] += ri * ri;
    }
    grid_reduce<1>(local, partials, tickets, scal, s_rr);
}
// p = z + (scal[num]/scal[den]) p
__global__ void k_cg_update_p(double* __restrict__ p, const double* __restrict__ z, i64 n, int s_num, int s_den,
                              const double* __restrict__ scal) {
    const double den = scal[s_den];
    const double b = den != 0.0 ? scal[s_num] / den : 0.0;
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) p[i] = z[i] + b * p[i];
}
template <int D>
__global__ void k_extract(const double* __restrict__ xb, double* xu, double* xc, i64 n_v) {
    constexpr int NB = D + 1;
    for (i64 v = blockIdx.x * (i64)TPB + threadIdx.x; v < n_v; v += (i64)gridDim.x * TPB) {
        if (xu)
#pragma unroll
            for (int k = 0; k < D; ++k) xu[v * D + k] = xb[v * NB + k];
        if (xc) xc[v] = xb[v * NB + D];
    }
}
template <int D>
__global__ void k_insert_add(double* xb, const double* du, const double* dc, double alpha, i64 n_v) {
    constexpr int NB = D + 1;
    for (i64 v = blockIdx.x * (i64)TPB + threadIdx.x; v < n_v; v += (i64)gridDim.x * TPB) {
        if (du)
#pragma unroll
            for (int k = 0; k < D; ++k) xb[v * NB + k] += alpha * du[v * D + k];
        if (dc) xb[v * NB + D] += alpha * dc[v];
    }
}
// |F_u|^2 and |F_c|^2 of a blocked vector -> slots s0, s0+1
template <int D>
__global__ void __launch_bounds__(TPB)
k_split_norms(const double* __restrict__ F, i64 n_v, double* partials, unsigned* tickets, double* scal, int s0) {
    constexpr int NB = D + 1;
    double local[2] = {0.0, 0.0};
    for (i64 v = blockIdx.x * (i64)TPB + threadIdx.x; v < n_v; v += (i64)gridDim.x * TPB) {
#pragma unroll
        for (int k = 0; k < D; ++k) local[0] += F[v * NB + k] * F[v * NB + k];
        local[1] += F[v * NB + D] * F[v * NB + D];
    }
    grid_reduce<2>(local, partials, tickets, scal, s0);
}
// F -= f_ext
__global__ void k_sub(double* __restrict__ F, const double* __restrict__ f, i64 n) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) F[i] -= f[i];
}
// SELL -> CSR order gather of values for export
template <int NC>
__global__ void k_export(const i64* __restrict__ rowptr, const i64* __restrict__ slice_off, int n_rows,
                         const double* __restrict__ A, double* out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    i64 base = slice_off[r >> 5];
    for (i64 t = rowptr[r]; t < rowptr[r + 1]; ++t) {
        i64 s = base + (t - rowptr[r]) * 32 + (r & 31);
        for (int k = 0; k < NC; ++k) out[t * NC + k] = A[vidx(s, k, NC)];
    }
}
template <int D>
__global__ void k_extrapolate_c(double* __restrict__ x, const double* __restrict__ xold, i64 n_v) {
    constexpr int NB = D + 1;
    for (i64 v = blockIdx.x * (i64)TPB + threadIdx.x; v < n_v; v += (i64)gridDim.x * TPB)
        x[v * NB + D] = 2.0 * x[v * NB + D] - xold[v * NB + D];
}
// end of a PCG iteration: r.z <- new r.z, append r.r to the ring, bump the iteration counter
__global__ void k_pcg_shift(double* scal, double* ring) {
    const int it = (int)ring[64];
    ring[it & 63] = scal[S_RR];
    ring[64] = (double)(it + 1);
    scal[S_RZ] = scal[S_RZNEW];
}
// condition of the PCG iteration's IF node: run the body while the last r.r (ring) is above tol^2 (ring[66]);
// NaN compares false and stops the loop.  ring[64] counts the iterations that really ran.
__global__ void k_pcg_cond(cudaGraphConditionalHandle handle, const double* __restrict__ ring) {
    const int it = (int)ring[64];
    unsigned go = 1u;
    if (it > 0) go = ring[(it - 1) & 63] > ring[66] ? 1u : 0u;
    cudaGraphSetConditional(handle, go);
}
__global__ void k_permute(const double* __restrict__ src, double* __restrict__ dst, const i64* __restrict__ perm, i64 n, bool scatter) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) {
        if (scatter) dst[perm[i]] = src[i]; else dst[i] = src[perm[i]];
    }
}
__global__ void k_flush(double* buf, i64 n) {
    for (i64 i = blockIdx.x * (i64)TPB + threadIdx.x; i < n; i += (i64)gridDim.x * TPB) buf[i] = (double)i;
}

}  // namespace

// ================================================================================================
// launch wrappers
#define LAUNCHED(c) ((c)->launches++)

template <int D>
static void assemble_dim(glims_ctx* c, int what, int variant) {
    constexpr int NB = D + 1;
    i64 n_own = c->pat.n_rows;
    const bool res = what & GLIMS_ASM_RESIDUAL, kconst = what & GLIMS_ASM_KCONST, kcc = what & GLIMS_ASM_KCC;
    i64 ns = c->pat.n_slots;
    if (variant == GLIMS_ASMK_ROWS) {
        // K_uu / K_uc from the tile kernel (state independent); K_cc and F_c by the row-walk kernel; F_u by SpMV over
        // the stored blocks, which therefore have to be the raw ones here (glims_step uses the eliminated blocks plus
        // the lift of the eliminated columns instead, solver.cu)
        if (!cc_available(c)) {
            static bool warned = false;
            if (!warned) { fprintf(stderr, "glims: row-walk assembly unavailable (%s); using the atomic kernel\n", cc_status(c)); warned = true; }
            variant = GLIMS_ASMK_ATOMIC;
        } else {
            if (kconst || (res && c->kuu_state != 1)) {
                assemble_dim<D>(c, GLIMS_ASM_KCONST, GLIMS_ASMK_TILE);
                if (!kconst) c->kconst_valid = false;      // the solver's eliminated copy is gone
            }
            if (res || kcc) {
                if (res) {
                    if (c->halo.active)
                        GL_CUDA(cudaMemsetAsync(c->F + n_own * NB, 0, sizeof(double) * (c->n_v - n_own) * NB, c->stream));
                    cc_mass_cprev(c);
                }
                launch_cc_rows(c, kcc, res);
                if (res) launch_fu(c, false);
            }
            return;
        }
    }
    if (kconst) c->kuu_state = 1;
    if (variant == GLIMS_ASMK_TILE) {
        // fused residual + Jacobian, one CTA per slice; writes every owned row of F and every slot (no zero-fill)
        if (res && c->halo.active)
            GL_CUDA(cudaMemsetAsync(c->F + n_own * NB, 0, sizeof(double) * (c->n_v - n_own) * NB, c->stream));
        if (launch_assemble_tile(c, what)) return;
        static bool warned = false;
        if (!warned) { fprintf(stderr, "glims: tile assembly unavailable (%s); using the slice/atomic kernels\n", tile_status(c, nullptr)); warned = true; }
        variant = GLIMS_ASMK_SLICE;
    }
    if (res) {
        GL_CUDA(cudaMemsetAsync(c->F, 0, sizeof(double) * c->ndof, c->stream));
    }
    bool slice = (variant == GLIMS_ASMK_SLICE) && (kconst || kcc);
    if (slice) {
        build_slice_map(c);
        constexpr int REC = NB * D + 3;
        size_t smem = (size_t)c->sl_max * REC * sizeof(double);
        if (c->sl_max >= 4096 || smem > 200 * 1024) slice = false;      // falls back to the plain gather kernel
        else {
            auto& p = c->pat;
#define SLICEK(KC, KK) do { auto kfn = k_assemble_slice<D, KC, KK>; \
            GL_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            kfn<<<p.n_slices, 256, smem, c->stream>>>(c->coords, c->cells, c->cell_mat, c->mat, c->n_mat, c->dt, c->x, \
                c->sl_ptr, c->sl_elem, c->gptr, c->lent, p.slice_off, p.slice_w, p.col, p.n_rows, c->Kuu, c->Kuc, c->Kcc); } while (0)
            if (kconst && kcc) SLICEK(true, true); else if (kconst) SLICEK(true, false); else SLICEK(false, true);
#undef SLICEK
            LAUNCHED(c);
            if (!res) return;
        }
    }
    bool gather = ((variant == GLIMS_ASMK_GATHER) || (variant == GLIMS_ASMK_SLICE && !slice)) && (kconst || kcc);
    if (gather) {
        build_gather_map(c);
        int g = nblk(ns, 128);
#define GATHER(KC, KK) k_assemble_gather<D, KC, KK><<<g, 128, 0, c->stream>>>(c->coords, c->cells, c->cell_mat, c->mat, \
            c->n_mat, c->dt, c->x, c->gptr, c->gent, ns, c->Kuu, c->Kuc, c->Kcc)
        if (kconst && kcc) GATHER(true, true); else if (kconst) GATHER(true, false); else GATHER(false, true);
#undef GATHER
        LAUNCHED(c);
        if (!res) return;
    } else if (!slice) {
        if (kconst) {
            GL_CUDA(cudaMemsetAsync(c->Kuu, 0, sizeof(double) * ns * D * D, c->stream));
            GL_CUDA(cudaMemsetAsync(c->Kuc, 0, sizeof(double) * ns * D, c->stream));
        }
        if (kcc) GL_CUDA(cudaMemsetAsync(c->Kcc, 0, sizeof(double) * ns, c->stream));
    }
    int g = nblk(c->n_c, 128);
#define ATOMIC(R, KC, KK) k_assemble_atomic<D, R, KC, KK><<<g, 128, 0, c->stream>>>(c->coords, c->cells, c->cell_mat, \
        c->mat, c->n_mat, c->n_c, n_own, c->dt, c->x, c->xprev, c->eslot, c->F, c->Kuu, c->Kuc, c->Kcc)
    bool akc = kconst && !gather && !slice, akk = kcc && !gather && !slice;
    if (res && akc && akk) ATOMIC(true, true, true);
    else if (res && akc) ATOMIC(true, true, false);
    else if (res && akk) ATOMIC(true, false, true);
    else if (res) ATOMIC(true, false, false);
    else if (akc && akk) ATOMIC(false, true, true);
    else if (akc) ATOMIC(false, true, false);
    else if (akk) ATOMIC(false, false, true);
#undef ATOMIC
    LAUNCHED(c);
    if (res && c->have_load) {
        k_sub<<<red_grid(c, c->ndof), TPB, 0, c->stream>>>(c->F, c->fext, (i64)c->pat.n_rows * NB);
        LAUNCHED(c);
    }
    (void)NB;
}

void launch_assemble(glims_ctx* c, int what, int variant) {
    if (c->dim == 2) assemble_dim<2>(c, what, variant); else assemble_dim<3>(c, what, variant);
    GL_CUDA(cudaGetLastError());
}

void launch_bc_values(glims_ctx* c, double* x) {
    if (!c->n_bc) return;
    k_bc_values<<<nblk(c->n_bc), TPB, 0, c->stream>>>(c->bc_dofs, c->bc_vals, c->n_bc, x);
    LAUNCHED(c);
}
void launch_bc_residual(glims_ctx* c, double* F, const double* x) {
    if (!c->n_bc) return;
    k_bc_residual<<<nblk(c->n_bc), TPB, 0, c->stream>>>(c->bc_dofs, c->bc_vals, c->n_bc, x, F);
    LAUNCHED(c);
}
void launch_bc_matrix(glims_ctx* c, int what, bool sym) {
    if (!c->n_bc) return;
    // skip the launch when no Dirichlet dof touches the requested blocks
    if (!(what & GLIMS_ASM_KCONST) && c->n_bc_c == 0) return;
    auto& p = c->pat;
    if (c->dim == 2) {
        k_bc_matrix<2><<<p.n_slices, 128, 0, c->stream>>>(c->bcmask, p.slice_off, p.slice_w, p.col, p.n_rows, what, sym, c->Kuu, c->Kuc, c->Kcc);
        k_bc_diag<2><<<nblk(c->n_bc), TPB, 0, c->stream>>>(c->bc_dofs, c->n_bc, p.n_rows, p.diag, what, c->Kuu, c->Kcc);
    } else {
        k_bc_matrix<3><<<p.n_slices, 128, 0, c->stream>>>(c->bcmask, p.slice_off, p.slice_w, p.col, p.n_rows, what, sym, c->Kuu, c->Kuc, c->Kcc);
        k_bc_diag<3><<<nblk(c->n_bc), TPB, 0, c->stream>>>(c->bc_dofs, c->n_bc, p.n_rows, p.diag, what, c->Kuu, c->Kcc);
    }
    c->launches += 2;
}
void launch_diag_inverse(glims_ctx* c, int which) {
    auto& p = c->pat;
    double* out = which == 0 ? c->dinv_mono : which == 1 ? c->dinv_uu : c->dinv_cc;
    if (c->dim == 2) k_diag_inverse<2><<<nblk(p.n_rows), TPB, 0, c->stream>>>(p.diag, p.n_rows, which, c->Kuu, c->Kuc, c->Kcc, out);
    else k_diag_inverse<3><<<nblk(p.n_rows), TPB, 0, c->stream>>>(p.diag, p.n_rows, which, c->Kuu, c->Kuc, c->Kcc, out);
    LAUNCHED(c);
}
void launch_export_values(glims_ctx* c, double* Kuu, double* Kuc, double* Kcc) {
    auto& p = c->pat;
    int g = nblk(p.n_rows);
    if (c->dim == 2) {
        k_export<4><<<g, TPB, 0, c->stream>>>(p.rowptr, p.slice_off, p.n_rows, c->Kuu, Kuu);
        k_export<2><<<g, TPB, 0, c->stream>>>(p.rowptr, p.slice_off, p.n_rows, c->Kuc, Kuc);
    } else {
        k_export<9><<<g, TPB, 0, c->stream>>>(p.rowptr, p.slice_off, p.n_rows, c->Kuu, Kuu);
        k_export<3><<<g, TPB, 0, c->stream>>>(p.rowptr, p.slice_off, p.n_rows, c->Kuc, Kuc);
    }
    k_export<1><<<g, TPB, 0, c->stream>>>(p.rowptr, p.slice_off, p.n_rows, c->Kcc, Kcc);
    c->launches += 3;
}

void launch_spmv(glims_ctx* c, int which, const double* x, double* y, SpmvDot dot) {
    auto& p = c->pat;
    const i64 tiles = (p.n_rows + TPB - 1) / TPB;
    bool d = dot.w != nullptr;
#define ARGS p.n_rows, dot.w, c->partials, c->tickets, c->scal, dot.slot
#define GO(K, ...) K<<<fit_grid(K, tiles, TPB, MAXBLK), TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, __VA_ARGS__, x, y, ARGS)
    if (which == 0) {
        if (c->dim == 2) { if (d) GO((k_spmv_mono<2, true>), c->Kuu, c->Kuc, c->Kcc); else GO((k_spmv_mono<2, false>), c->Kuu, c->Kuc, c->Kcc); }
        else             { if (d) GO((k_spmv_mono<3, true>), c->Kuu, c->Kuc, c->Kcc); else GO((k_spmv_mono<3, false>), c->Kuu, c->Kuc, c->Kcc); }
    } else if (which == 1) {
        if (c->dim == 2) { if (d) GO((k_spmv_block<2, 2, true>), c->Kuu); else GO((k_spmv_block<2, 2, false>), c->Kuu); }
        else             { if (d) GO((k_spmv_block<3, 3, true>), c->Kuu); else GO((k_spmv_block<3, 3, false>), c->Kuu); }
    } else {
        const double* A = which == 3 ? cc_mass_matrix(c) : c->Kcc;      // 3: the P1 mass matrix (L2 projections)
        if (d) GO((k_spmv_block<1, 1, true>), A); else GO((k_spmv_block<1, 1, false>), A);
    }
#undef GO
#undef ARGS
    LAUNCHED(c);
}
void launch_spmv_uc(glims_ctx* c, const double* xc, double* yu) {
    auto& p = c->pat;
    if (c->dim == 2) k_spmv_uc<2><<<nblk(p.n_slices * 32), TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, c->Kuc, xc, yu, p.n_rows);
    else k_spmv_uc<3><<<nblk(p.n_slices * 32), TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, c->Kuc, xc, yu, p.n_rows);
    LAUNCHED(c);
}
void launch_spmv_generic(glims_ctx* c, const SellPattern& p, const double* A, int bs, const double* x, double* y,
                         const double* rhs) {
    // rhs == nullptr: y = A x ; else y = rhs - A x
    if (bs == 6) {     // coarse AMG levels: warp-per-component kernel
        int g = p.n_slices < 148 * 16 ? p.n_slices : 148 * 16;
        if (g < 1) g = 1;
        if (rhs) k_spmv_split<6, true><<<g, 192, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, A, x, y, p.n_rows, p.n_slices, rhs);
        else k_spmv_split<6, false><<<g, 192, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, A, x, y, p.n_rows, p.n_slices, nullptr);
        LAUNCHED(c);
        return;
    }
    const i64 tiles = (p.n_rows + TPB - 1) / TPB;
#define GEN(B) do { if (rhs) k_spmv_block<B, B, false, true><<<fit_grid(k_spmv_block<B, B, false, true>, tiles, TPB, MAXBLK), TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, A, x, y, p.n_rows, rhs, c->partials, c->tickets, c->scal, 0); \
                    else k_spmv_block<B, B, false, false><<<fit_grid(k_spmv_block<B, B, false, false>, tiles, TPB, MAXBLK), TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, A, x, y, p.n_rows, nullptr, c->partials, c->tickets, c->scal, 0); } while (0)
    if (bs == 1) GEN(1); else if (bs == 2) GEN(2); else if (bs == 3) GEN(3);
    else throw GlError(GLIMS_ERR_ARG, "spmv_generic: unsupported block size");
#undef GEN
    LAUNCHED(c);
}

void launch_dot(glims_ctx* c, const double* a, const double* b, i64 n, int slot) {
    k_dot<<<red_grid(c, n), TPB, 0, c->stream>>>(a, b, n, c->partials, c->tickets, c->scal, slot);
    LAUNCHED(c);
}
void launch_multi_dot(glims_ctx* c, const double* V, i64 ld, int k, const double* w, i64 n, int slot0) {
    k_multi_dot<<<red_grid(c, n), TPB, 0, c->stream>>>(V, ld, k, w, n, c->partials, c->tickets, c->scal, slot0);
    LAUNCHED(c);
}
void launch_multi_axpy(glims_ctx* c, const double* V, i64 ld, int k, const double* coef_dev, double sign, double* w, i64 n) {
    k_multi_axpy<<<red_grid(c, n), TPB, 0, c->stream>>>(V, ld, k, coef_dev, sign, w, n);
    LAUNCHED(c);
}
void launch_axpy(glims_ctx* c, double alpha, const double* x, double* y, i64 n) {
    k_axpy<<<red_grid(c, n), TPB, 0, c->stream>>>(alpha, x, y, n);
    LAUNCHED(c);
}
void launch_scale(glims_ctx* c, double alpha, double* x, i64 n) {
    k_scale<<<red_grid(c, n), TPB, 0, c->stream>>>(alpha, x, n);
    LAUNCHED(c);
}
void launch_copy(glims_ctx* c, const double* x, double* y, i64 n) {
    GL_CUDA(cudaMemcpyAsync(y, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
}
void launch_zero(glims_ctx* c, double* x, i64 n) { GL_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n, c->stream)); }
void launch_cg_update_xr(glims_ctx* c, double* x, double* r, const double* p, const double* Ap, i64 n, int s_num,
                         int s_den, int s_rr) {
    k_cg_update_xr<<<red_grid(c, n), TPB, 0, c->stream>>>(x, r, p, Ap, n, s_num, s_den, s_rr, c->partials, c->tickets, c->scal);
    LAUNCHED(c);
}
void launch_cg_update_p(glims_ctx* c, double* p, const double* z, i64 n, int s_num, int s_den) {
    k_cg_update_p<<<red_grid(c, n), TPB, 0, c->stream>>>(p, z, n, s_num, s_den, c->scal);
    LAUNCHED(c);
}
void launch_block_jacobi(glims_ctx* c, const double* dinv, int bs, const double* r, double* z, i64 n_rows, int s_rz) {
    int g = red_grid(c, n_rows);
#define BJ(B) k_block_jacobi<B><<<g, TPB, 0, c->stream>>>(dinv, r, z, n_rows, c->partials, c->tickets, c->scal, s_rz)
    if (bs == 1) BJ(1); else if (bs == 2) BJ(2); else if (bs == 3) BJ(3); else if (bs == 4) BJ(4); else if (bs == 6) BJ(6);
    else throw GlError(GLIMS_ERR_ARG, "block_jacobi: unsupported block size");
#undef BJ
    LAUNCHED(c);
}
void launch_extract(glims_ctx* c, const double* xb, double* xu, double* xc) {
    i64 n = c->pat.n_rows;
    if (c->dim == 2) k_extract<2><<<red_grid(c, n), TPB, 0, c->stream>>>(xb, xu, xc, n);
    else k_extract<3><<<red_grid(c, n), TPB, 0, c->stream>>>(xb, xu, xc, n);
    LAUNCHED(c);
}
void launch_insert_add(glims_ctx* c, double* xb, const double* du, const double* dc, double alpha) {
    i64 n = c->pat.n_rows;
    if (c->dim == 2) k_insert_add<2><<<red_grid(c, n), TPB, 0, c->stream>>>(xb, du, dc, alpha, n);
    else k_insert_add<3><<<red_grid(c, n), TPB, 0, c->stream>>>(xb, du, dc, alpha, n);
    LAUNCHED(c);
}
void launch_split_norms(glims_ctx* c, const double* F, int s0) {
    i64 n = c->pat.n_rows;
    if (c->dim == 2) k_split_norms<2><<<red_grid(c, n), TPB, 0, c->stream>>>(F, n, c->partials, c->tickets, c->scal, s0);
    else k_split_norms<3><<<red_grid(c, n), TPB, 0, c->stream>>>(F, n, c->partials, c->tickets, c->scal, s0);
    LAUNCHED(c);
}
void read_scalars(glims_ctx* c, int slot0, int n, double* out) {
    GL_CUDA(cudaMemcpyAsync(c->h_scal + slot0, c->scal + slot0, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) out[i] = c->h_scal[slot0 + i];
}
void launch_cell_fields(glims_ctx* c, double* out, double* vol) {
    int g = nblk(c->n_c, 128);
    if (c->dim == 2) k_cell_fields<2><<<g, 128, 0, c->stream>>>(c->coords, c->cells, c->cell_mat, c->mat, c->n_mat, c->n_c, c->x, out, vol);
    else k_cell_fields<3><<<g, 128, 0, c->stream>>>(c->coords, c->cells, c->cell_mat, c->mat, c->n_mat, c->n_c, c->x, out, vol);
    LAUNCHED(c);
}
void launch_cell_to_vertex(glims_ctx* c, int nf, const double* q, const double* vol, double* num, double* den) {
    GL_CUDA(cudaMemsetAsync(num, 0, sizeof(double) * c->n_v * nf, c->stream));
    GL_CUDA(cudaMemsetAsync(den, 0, sizeof(double) * c->n_v, c->stream));
    k_cell_to_vertex<<<nblk(c->n_c), TPB, 0, c->stream>>>(c->cells, c->nb, c->n_c, nf, q, vol, num, den);
    k_divide_rows<<<nblk(c->n_v * nf), TPB, 0, c->stream>>>(num, den, c->n_v, nf);
    c->launches += 2;
}
void launch_project_load(glims_ctx* c, const double* q, const double* vol, double* load) {
    const int nf = 2 * c->dim * c->dim + 5;
    GL_CUDA(cudaMemsetAsync(load, 0, sizeof(double) * c->n_v * nf, c->stream));
    if (c->dim == 2) k_project_load<2><<<nblk(c->n_c, 128), 128, 0, c->stream>>>(c->cells, c->n_c, q, vol, c->cell_mat, c->mat, c->x, load);
    else k_project_load<3><<<nblk(c->n_c, 128), 128, 0, c->stream>>>(c->cells, c->n_c, q, vol, c->cell_mat, c->mat, c->x, load);
    LAUNCHED(c);
}
void launch_strided_copy(glims_ctx* c, const double* src, i64 n, int ss, int os, double* dst, int sd, int od) {
    k_strided_copy<<<red_grid(c, n), TPB, 0, c->stream>>>(src, n, ss, os, dst, sd, od);
    LAUNCHED(c);
}
void launch_scalar_diag_inverse(glims_ctx* c, const double* A, double* out) {
    k_scalar_diag_inverse<<<nblk(c->pat.n_rows), TPB, 0, c->stream>>>(c->pat.diag, c->pat.n_rows, A, out);
    LAUNCHED(c);
}
void launch_extrapolate_c(glims_ctx* c, double* x, const double* xold) {
    if (c->dim == 2) k_extrapolate_c<2><<<red_grid(c, c->n_v), TPB, 0, c->stream>>>(x, xold, c->n_v);
    else k_extrapolate_c<3><<<red_grid(c, c->n_v), TPB, 0, c->stream>>>(x, xold, c->n_v);
    LAUNCHED(c);
}
void launch_pcg_shift(glims_ctx* c, double* ring) {
    k_pcg_shift<<<1, 1, 0, c->stream>>>(c->scal, ring);
    LAUNCHED(c);
}
void launch_pcg_cond(glims_ctx* c, unsigned long long handle, const double* ring) {
    k_pcg_cond<<<1, 1, 0, c->stream>>>((cudaGraphConditionalHandle)handle, ring);
    LAUNCHED(c);
}
void launch_permute(glims_ctx* c, const double* src, double* dst, const i64* perm, i64 n, bool scatter) {
    k_permute<<<red_grid(c, n), TPB, 0, c->stream>>>(src, dst, perm, n, scatter);
    LAUNCHED(c);
}
void flush_l2(glims_ctx* c) {
    if (!c->flush_buf) {
        c->flush_bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
        GL_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
    }
    k_flush<<<148 * 8, TPB, 0, c->stream>>>((double*)c->flush_buf, (i64)(c->flush_bytes / 8));
}
