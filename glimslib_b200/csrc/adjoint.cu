// Discrete adjoint of the backward-Euler time loop on the device.
//
// The reference's production caller is an inverse problem: `run_for_adjoint` re-runs the forward model for new values of
// (D_WM, D_GM, rho_WM, rho_GM, coupling) (simulation_tumor_growth_brain.py:127-145) and dolfin-adjoint differentiates a
// misfit of the final state -- smoothly thresholded concentration at several levels plus displacement -- with respect to
// those controls (image_based_optimization.py:660-767; threshold 0.5*(tanh((c - level)/0.01) + 1), :1404-1407).
//
//   forward   R_n(x_n, x_{n-1}; p) = 0,  n = 1..N                       (glims_step; the trajectory stays on the device)
//   misfit    J(x_N) = sum_l (th_l(c_N) - t_l)^T M (th_l(c_N) - t_l) + (u_N - u_t)^T (M x I_d) (u_N - u_t)
//   adjoint   (dR_N/dx_N)^T l_N = -dJ/dx_N ;  (dR_n/dx_n)^T l_n = -(dR_{n+1}/dx_n)^T l_{n+1} = M l^c_{n+1}  (concentration rows)
//   gradient  dJ/dp = sum_n l_n^T dR_n/dp,  p = (D_m, rho_m, gamma_m) of every material m
//
// The Jacobian is block upper triangular, J = [[K_uu, K_uc], [0, K_cc]], so its transpose is solved by the forward step's two
// SPD solves in the opposite order: K_uu l_u = rhs_u (PCG + AMG, the hierarchy of the forward run), then
// K_cc(c_n) l_c = rhs_c - K_uc^T l_u (Jacobi-PCG; K_cc re-assembled at the stored state by the row-walk kernel).  Dirichlet rows
// are identities, their multipliers never enter the gradient (dR/dp vanishes there).  The controls enter R linearly:
//   dR_c/dD_m   . l_c =  dt  sum_{e in m} |K| (grad l_c . grad c)
//   dR_c/drho_m . l_c = -dt  sum_{e in m} |K| sum_a l_c,a [ m(c_a + S) - kappa(2c_a^2 + 2c_a S + S^2 + Q) ]
//   dR_u/dgam_m . l_u = -    sum_{e in m} (2mu + d lambda) |K| (S/(d+1)) div(l_u)
// one element-parallel kernel with a per-material reduction.  Same statement as oracle/adjoint.py (which is checked against
// central finite differences of the forward run); tests/test_gpu_adjoint.py compares the two.
#include "common.h"
#include "geom.cuh"
#include <cmath>
#include <vector>

namespace {

constexpr int TPB = 256;
inline int nblk(i64 n, int t = TPB) { return (int)((n + t - 1) / t); }

// r = th(c) - target, dth = th'(c)   (image_based_optimization.py:1404-1407, width 0.01)
__global__ void k_threshold(const double* __restrict__ c, const double* __restrict__ tgt, double level, double width, i64 n,
                            double* __restrict__ r, double* __restrict__ dth) {
    i64 i = blockIdx.x * (i64)TPB + threadIdx.x;
    if (i >= n) return;
    const double t = tanh((c[i] - level) / width);
    r[i] = 0.5 * (t + 1.0) - tgt[i];
    dth[i] = 0.5 * (1.0 - t * t) / width;
}
// g[i*sg + og] += s * a[i] * (b ? b[i] : 1)
__global__ void k_acc_prod(double* __restrict__ g, int sg, int og, double s, const double* __restrict__ a,
                           const double* __restrict__ b, i64 n) {
    i64 i = blockIdx.x * (i64)TPB + threadIdx.x;
    if (i < n) g[i * sg + og] += s * a[i] * (b ? b[i] : 1.0);
}
__global__ void k_diff_strided(const double* __restrict__ a, const double* __restrict__ b, int stride, int off, i64 n, double* __restrict__ r) {
    i64 i = blockIdx.x * (i64)TPB + threadIdx.x;
    if (i < n) r[i] = a[i * stride + off] - b[i * stride + off];
}
// zero the entries of split (u | c) vectors at Dirichlet dofs
template <int D>
__global__ void k_zero_bc_split(const i64* __restrict__ dofs, i64 n, i64 n_rows, double* ru, double* rc) {
    i64 t = blockIdx.x * (i64)TPB + threadIdx.x;
    if (t >= n) return;
    const i64 v = dofs[t] / (D + 1);
    const int k = (int)(dofs[t] - v * (D + 1));
    if (v >= n_rows) return;
    if (k < D) ru[v * D + k] = 0.0; else rc[v] = 0.0;
}
// out[col] += K_uc(row, col)^T . l_u[row]   (transposed D x 1 block product; out zeroed by the caller)
template <int D>
__global__ void __launch_bounds__(TPB)
k_spmv_uc_T(const i64* __restrict__ slice_off, const int* __restrict__ slice_w, const int* __restrict__ col,
            const double* __restrict__ Kuc, const double* __restrict__ lu, double* out, int n_rows) {
    const int r = blockIdx.x * TPB + threadIdx.x;
    const int S = r >> 5, lane = r & 31;
    if (S * 32 >= n_rows || r >= n_rows) return;
    double l[D];
#pragma unroll
    for (int i = 0; i < D; ++i) l[i] = lu[(i64)r * D + i];
    const i64 base = slice_off[S];
    const int w = slice_w[S];
    for (int j = 0; j < w; ++j) {
        const i64 g = base + (i64)j * 32;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s += Kuc[g * D + i * 32 + lane] * l[i];
        if (s != 0.0) atomicAdd(&out[col[g + lane]], s);
    }
}
// per-material gradient contributions of one adjoint step: grad[m][0..2] += (D, rho, gamma) terms
template <int D>
__global__ void __launch_bounds__(128)
k_adjoint_grad(const double* __restrict__ coords, const int* __restrict__ cells, const int* __restrict__ cell_mat,
               const double* __restrict__ mat_g, int n_mat, i64 n_c, double dt, const double* __restrict__ x,
               const double* __restrict__ lu, const double* __restrict__ lc, double* grad) {
    constexpr int NB = D + 1;
    __shared__ double sg[MAX_MAT * 3];
    for (int t = threadIdx.x; t < n_mat * 3; t += blockDim.x) sg[t] = 0.0;
    __syncthreads();
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e < n_c) {
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
        const int m = cell_mat[e];
        const double* mt = mat_g + m * MAT_STRIDE;
        const double mu = mt[0], lam = mt[1];
        double cv[NB], lcv[NB], S = 0, Q = 0, gl[D], gc[D], divl = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) { gl[k] = 0.0; gc[k] = 0.0; }
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            cv[a] = x[(i64)v[a] * NB + D];
            lcv[a] = lc[v[a]];
            S += cv[a]; Q += cv[a] * cv[a];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                gl[k] += lcv[a] * G.g[a][k];
                gc[k] += cv[a] * G.g[a][k];
                divl += lu[(i64)v[a] * D + k] * G.g[a][k];
            }
        }
        double dD = 0.0, dR = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) dD += gl[k] * gc[k];
#pragma unroll
        for (int a = 0; a < NB; ++a)
            dR += lcv[a] * (Consts<D>::mass * (cv[a] + S) - Consts<D>::kappa * (2.0 * cv[a] * cv[a] + 2.0 * cv[a] * S + S * S + Q));
        atomicAdd(&sg[m * 3 + 0], dt * G.vol * dD);
        atomicAdd(&sg[m * 3 + 1], -dt * G.vol * dR);
        atomicAdd(&sg[m * 3 + 2], -(2.0 * mu + D * lam) * G.vol * (S / NB) * divl);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_mat * 3; t += blockDim.x)
        if (sg[t] != 0.0) atomicAdd(&grad[t], sg[t]);
}

}  // namespace

void launch_spmv_uc_T(glims_ctx* c, const double* lu, double* out) {
    auto& p = c->pat;
    GL_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * c->n_v, c->stream));
    if (c->dim == 2) k_spmv_uc_T<2><<<nblk(p.n_slices * 32), TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, c->Kuc, lu, out, p.n_rows);
    else k_spmv_uc_T<3><<<nblk(p.n_slices * 32), TPB, 0, c->stream>>>(p.slice_off, p.slice_w, p.col, c->Kuc, lu, out, p.n_rows);
    c->launches++;
}
void launch_adjoint_grad(glims_ctx* c, const double* x, const double* lu, const double* lc, double* grad) {
    const int g = nblk(c->n_c, 128);
    if (c->dim == 2) k_adjoint_grad<2><<<g, 128, 0, c->stream>>>(c->coords, c->cells, c->cell_mat, c->mat, c->n_mat, c->n_c, c->dt, x, lu, lc, grad);
    else k_adjoint_grad<3><<<g, 128, 0, c->stream>>>(c->coords, c->cells, c->cell_mat, c->mat, c->n_mat, c->n_c, c->dt, x, lu, lc, grad);
    c->launches++;
}
void launch_threshold(glims_ctx* c, const double* cvec, const double* tgt, double level, double* r, double* dth) {
    k_threshold<<<nblk(c->n_v), TPB, 0, c->stream>>>(cvec, tgt, level, 0.01, c->n_v, r, dth);
    c->launches++;
}
void launch_acc_prod(glims_ctx* c, double* g, int sg, int og, double s, const double* a, const double* b, i64 n) {
    k_acc_prod<<<nblk(n), TPB, 0, c->stream>>>(g, sg, og, s, a, b, n);
    c->launches++;
}
void launch_diff_strided(glims_ctx* c, const double* a, const double* b, int stride, int off, i64 n, double* r) {
    k_diff_strided<<<nblk(n), TPB, 0, c->stream>>>(a, b, stride, off, n, r);
    c->launches++;
}
void launch_zero_bc_split(glims_ctx* c, double* ru, double* rc) {
    if (!c->n_bc) return;
    if (c->dim == 2) k_zero_bc_split<2><<<nblk(c->n_bc), TPB, 0, c->stream>>>(c->bc_dofs, c->n_bc, c->pat.n_rows, ru, rc);
    else k_zero_bc_split<3><<<nblk(c->n_bc), TPB, 0, c->stream>>>(c->bc_dofs, c->n_bc, c->pat.n_rows, ru, rc);
    c->launches++;
}
