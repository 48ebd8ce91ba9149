// Tile assembly (K1 + K2 fused, variant GLIMS_ASMK_TILE): data structures and the per-thread work of every
// phase as __host__ __device__ functions.  The CUDA kernel (tile.cu) and the host emulator used by the CPU
// tests (tests/emu/tile_emu.cpp) both call exactly these functions, so index arithmetic, formulas and maps are
// checked on the CPU against the oracle before the kernel ever runs on a GPU.
//
// One CTA owns one tile = 16 consecutive block rows (half a SELL-32 slice).  Shared memory holds
//   sv   : the tile's local vertices (own 16 rows first, then every other vertex of a touching element):
//          coords, u, c, c_prev                                    -- gathered from global memory once
//   rec  : per touching element sqrt|K|*grad(lambda_a), sqrt|K|, |K|, |K|*sum(c); structure of arrays
//          [a][element][3] + [3][element] so that the bank of an access depends on (element mod 16) only
//   sent : the tile's contributor entries (u16: local element | a<<12 | b<<14), ELL layout [iteration][lane]
// Phase B: a warp takes one "item" = two SELL columns of the tile (lanes 0-15 / 16-31; lane & 15 = row), or the two
// halves of one long column such as the diagonal (partial sums combined with a shuffle).  Every contributor costs 9 shared-memory doubles and 14 FMAs: the raw sums
//   S = sum g~a (x) g~b,  t = sum sqrt|K| g~a,  V = sum |K|,  W = sum |K| sum(c)
// are material-free; the material enters once per slot (chunks in which a lane mixes tissues take a
// per-contributor path).  The residual is formed from the same sums: F_u = K_uu u + K_uc c, F_c = (K_lin + J_r/2) c
// - M c_prev (exact identities of the element formulas), so the fused kernel writes K and F
// in one pass with no atomics and a fixed summation order.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstddef>
#include <string>
#include <vector>

#ifdef __CUDACC__
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

typedef long long tl_i64;

constexpr int TILE_ROWS = 16;

struct TileHdr {        // one per tile (tile T = rows [16 T, 16 T + 16) = half (T & 1) of slice T >> 1)
    tl_i64 v_off;       // into tv
    tl_i64 e_off;       // into te
    tl_i64 ent_off;     // into ent (u16 units, multiple of 8)
    int item_off;       // into items
    int n_lv;           // local vertices (>= 16, <= 1024)
    int n_el;           // element records incl. bucket padding; records n_el .. n_el+15 are all-zero (padding entries)
    int n_items;
    int n_ent;          // u16 entries of this tile (multiple of 8)
    int pad;
};
struct TileItem {       // 12 bytes; h = lane >> 4 selects the half-warp
    uint16_t col_j[2];  // SELL column handled by each half-warp (equal for a split item)
    uint16_t L;         // contributor iterations (max of the two halves; the shorter one is padded)
    uint8_t flags;      // TILE_MIXED | TILE_SPLIT | TILE_NULLB
    uint8_t pad;
    uint32_t ent_off;   // offset (u16 units) inside the tile's entry block
};
// MIXED: some lane mixes materials in this column -> per-contributor weights.  SPLIT: both half-warps work on
// the same (long) column, first / second half of its contributors; the partial sums are combined with a shuffle
// and half 0 writes.  NULLB: half 1 has no column.
enum { TILE_MIXED = 1, TILE_SPLIT = 2, TILE_NULLB = 4, TILE_NULLITEM = 8 /* padding of the warp schedule: skip */ };
// lcol entry of a slot: local vertex of the column (10 bits) | material of the slot's contributors << 10
constexpr int TILE_LCOL_BITS = 10;
constexpr unsigned long long TILE_NOELEM = ~0ULL;
constexpr int TILE_MAT_STRIDE = 6;   // mu, lambda, D, rho, gamma, beta (same table as common.h)

template <int D> struct TileC {
    static constexpr int NB = D + 1;
    static constexpr int DD = D * D;
    static constexpr int GS = 3;               // doubles per gradient slot (odd: distinct (element mod 16) => distinct banks)
    static constexpr int REC = NB * GS + 3;    // doubles per element: NB gradient slots + sqrt|K|, |K|, |K| sum(c)
    static constexpr int VS = 2 * D + 3;       // coords, u, c, c_prev (+1 pad): 9 / 7, odd
    static constexpr int KF = DD + D + 4;      // K_uu, K_uc, kl, Vrho, Wrho, Mv
    static constexpr double mass() { return D == 2 ? 1.0 / 12.0 : 1.0 / 20.0; }
    static constexpr double kappa() { return D == 2 ? 1.0 / 60.0 : 1.0 / 120.0; }
};

// ---- phase 0: one local vertex -> sv --------------------------------------------------------------
template <int D>
GL_HD void tile_stage_vertex(int v, const double* coords, const double* x, const double* xprev, double* o) {
    constexpr int NB = D + 1;
    if (v < 0) {
#pragma unroll
        for (int k = 0; k < TileC<D>::VS; ++k) o[k] = 0.0;
        return;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = coords[(tl_i64)v * D + k];
#pragma unroll
    for (int k = 0; k < NB; ++k) o[D + k] = x[(tl_i64)v * NB + k];
    o[D + NB] = xprev[(tl_i64)v * NB + D];
}

// ---- phase A: one element record -> rec, emat -----------------------------------------------------
GL_HD void tile_geometry(const double (&X)[3][2], double (&g)[3][2], double& vol) {
    double j00 = X[1][0] - X[0][0], j01 = X[2][0] - X[0][0];
    double j10 = X[1][1] - X[0][1], j11 = X[2][1] - X[0][1];
    double det = j00 * j11 - j01 * j10, id = 1.0 / det;
    g[1][0] = j11 * id;  g[1][1] = -j01 * id;
    g[2][0] = -j10 * id; g[2][1] = j00 * id;
    g[0][0] = -(g[1][0] + g[2][0]);
    g[0][1] = -(g[1][1] + g[2][1]);
    vol = 0.5 * fabs(det);
}
// 3D: returns the gradients already scaled by sqrt|K| (what the records store) with one reciprocal square root:
// grad = cof/det, |K| = |det|/6  =>  sqrt|K| * grad = cof * sign(det) / sqrt(6 |det|),  sqrt|K| = |det| / sqrt(6 |det|).
GL_HD void tile_geometry_scaled(const double (&X)[4][3], double (&g)[4][3], double& sq, double& vol) {
    double e1[3], e2[3], e3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { e1[k] = X[1][k] - X[0][k]; e2[k] = X[2][k] - X[0][k]; e3[k] = X[3][k] - X[0][k]; }
    double c1[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
    double c2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
    double c3[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double det = e1[0] * c1[0] + e1[1] * c1[1] + e1[2] * c1[2], ad = fabs(det);
#ifdef __CUDA_ARCH__
    const double r = rsqrt(6.0 * ad);
#else
    const double r = 1.0 / sqrt(6.0 * ad);
#endif
    const double s = det < 0.0 ? -r : r;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        g[1][k] = c1[k] * s; g[2][k] = c2[k] * s; g[3][k] = c3[k] * s;
        g[0][k] = -(g[1][k] + g[2][k] + g[3][k]);
    }
    sq = ad * r;
    vol = ad * (1.0 / 6.0);
}
GL_HD void tile_geometry_scaled(const double (&X)[3][2], double (&g)[3][2], double& sq, double& vol) {
    tile_geometry(X, g, vol);
    sq = sqrt(vol);
#pragma unroll
    for (int a = 0; a < 3; ++a) { g[a][0] *= sq; g[a][1] *= sq; }
}

// Element records in shared memory, structure of arrays over NE (a multiple of 16) element positions:
//   gradient a of element i at rec[(a*NE + i)*GS + k];  sqrt|K|, |K|, |K|*sum(c) at rec[NB*NE*GS + q*NE + i].
// te record: 4 x 12-bit local vertex | material << 48
template <int D>
GL_HD void tile_stage_element(unsigned long long rec64, const double* sv, double* rec, int NE, int i, unsigned char* emat) {
    constexpr int NB = D + 1, GS = TileC<D>::GS, VS = TileC<D>::VS;
    double* sc = rec + NB * NE * GS + i;
    if (rec64 == TILE_NOELEM) {
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int k = 0; k < D; ++k) rec[(a * NE + i) * GS + k] = 0.0;
        sc[0] = 0.0; sc[NE] = 0.0; sc[2 * NE] = 0.0;
        *emat = 255;
        return;
    }
    double X[NB][D], S = 0.0;
#pragma unroll
    for (int a = 0; a < NB; ++a) {
        const double* p = sv + (int)((rec64 >> (12 * a)) & 0xfffULL) * VS;
#pragma unroll
        for (int k = 0; k < D; ++k) X[a][k] = p[k];
        S += p[2 * D];
    }
    double g[NB][D], vol, sq;
    tile_geometry_scaled(X, g, sq, vol);
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int k = 0; k < D; ++k) rec[(a * NE + i) * GS + k] = g[a][k];
    sc[0] = sq;
    sc[NE] = vol;
    sc[2 * NE] = vol * S;
    *emat = (unsigned char)((rec64 >> 48) & 0xffULL);
}

// ---- phase B ---------------------------------------------------------------------------------------
// contributors of one slot chunk -> "K form": kf = [K_uu (DD) | K_uc (D) | kl | Vrho | Wrho | Mv]
//   kl   = m(1+d_ab) sum (1 - dt rho)|K| + dt sum D |K| g_a.g_b      (linear part of K_cc)
//   Vrho = sum rho |K|,  Wrho = sum rho |K| sum(c),  Mv = m(1+d_ab) sum |K|
template <int D>
GL_HD void tile_accumulate(const double* rec, int NE, const unsigned char* emat, const double* smat, const uint16_t* ent,
                           int L, int lane, bool mixed, int slot_mat, bool diag, double dt, double (&kf)[TileC<D>::KF]) {
    constexpr int NB = D + 1, DD = D * D, GS = TileC<D>::GS;
    const double* scal = rec + NB * NE * GS;
    const double mfac = TileC<D>::mass() * (diag ? 2.0 : 1.0);
    if (!mixed) {
        double Sm[DD], t[D], V = 0.0, W = 0.0;
#pragma unroll
        for (int k = 0; k < DD; ++k) Sm[k] = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) t[k] = 0.0;
        for (int j = 0; j < L; ++j) {
            const unsigned e = ent[j * 32 + lane];
            const int lel = e & 0xfff;
            const double* ga = rec + (((e >> 12) & 3) * NE + lel) * GS;
            const double* gb = rec + ((e >> 14) * NE + lel) * GS;
            double a_[D], b_[D];
#pragma unroll
            for (int k = 0; k < D; ++k) { a_[k] = ga[k]; b_[k] = gb[k]; }
            const double sq = scal[lel];
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int k = 0; k < D; ++k) Sm[i * D + k] = fma(a_[i], b_[k], Sm[i * D + k]);
                t[i] = fma(sq, a_[i], t[i]);
            }
            V += scal[NE + lel];
            W += scal[2 * NE + lel];
        }
        const double* mt = smat + slot_mat * TILE_MAT_STRIDE;
        const double mu = mt[0], lam = mt[1], Dc = mt[2], rho = mt[3], beta = mt[5];
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) tr += Sm[i * D + i];
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int k = 0; k < D; ++k) kf[i * D + k] = mu * (Sm[k * D + i] + (i == k ? tr : 0.0)) + lam * Sm[i * D + k];
            kf[DD + i] = -beta * (1.0 / NB) * t[i];
        }
        kf[DD + D] = mfac * (1.0 - dt * rho) * V + dt * Dc * tr;
        kf[DD + D + 1] = rho * V;
        kf[DD + D + 2] = rho * W;
        kf[DD + D + 3] = mfac * V;
    } else {
#pragma unroll
        for (int k = 0; k < TileC<D>::KF; ++k) kf[k] = 0.0;
        for (int j = 0; j < L; ++j) {
            const unsigned e = ent[j * 32 + lane];
            const int lel = e & 0xfff;
            const double* ga = rec + (((e >> 12) & 3) * NE + lel) * GS;
            const double* gb = rec + ((e >> 14) * NE + lel) * GS;
            double a_[D], b_[D], gg = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) { a_[k] = ga[k]; b_[k] = gb[k]; gg = fma(a_[k], b_[k], gg); }
            const double sq = scal[lel], vol = scal[NE + lel], vS = scal[2 * NE + lel];
            int m = emat[lel];
            if (m == 255) m = 0;
            const double* mt = smat + m * TILE_MAT_STRIDE;
            const double mu = mt[0], lam = mt[1], Dc = mt[2], rho = mt[3], beta = mt[5];
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int k = 0; k < D; ++k)
                    kf[i * D + k] += mu * ((i == k ? gg : 0.0) + a_[k] * b_[i]) + lam * a_[i] * b_[k];
                kf[DD + i] += -beta * (1.0 / NB) * sq * a_[i];
            }
            kf[DD + D] += mfac * (1.0 - dt * rho) * vol + dt * Dc * gg;
            kf[DD + D + 1] += rho * vol;
            kf[DD + D + 2] += rho * vS;
            kf[DD + D + 3] += mfac * vol;
        }
    }
}

// K form of one slot -> matrix values (optional) and the slot's share of the row residual.
//   xa: sv record of the row vertex, xb: sv record of the column vertex.
template <int D, bool WKCONST, bool WKCC, bool RES>
GL_HD void tile_finalize(const double (&kf)[TileC<D>::KF], bool diag, double dt, const double* xa, const double* xb,
                         tl_i64 g /* first slot of the column's 32-group */, int lane /* slot lane inside the group */,
                         double* Kuu, double* Kuc, double* Kcc,
                         double (&Facc)[D + 1]) {
    constexpr int DD = D * D;
    if (WKCONST) {
#pragma unroll
        for (int k = 0; k < DD; ++k) Kuu[g * DD + k * 32 + lane] = kf[k];
#pragma unroll
        for (int k = 0; k < D; ++k) Kuc[g * D + k * 32 + lane] = kf[DD + k];
    }
    const double ca = xa[2 * D], cb = xb[2 * D];
    const double kl = kf[DD + D];
    const double jr = 2.0 * dt * TileC<D>::kappa() *
                      ((diag ? 2.0 : 1.0) * kf[DD + D + 2] + (diag ? 4.0 * ca : ca + cb) * kf[DD + D + 1]);
    if (WKCC) Kcc[g + lane] = kl + jr;
    if (RES) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = kf[DD + i] * cb;
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(kf[i * D + k], xb[D + k], s);
            Facc[i] += s;
        }
        Facc[D] += (kl + 0.5 * jr) * cb - kf[DD + D + 3] * xb[2 * D + 1];
    }
}

// shared-memory carve-up (byte offsets, 16-byte aligned), identical on host and device
struct TileSmem {
    int lv_cap, el_cap, ent_cap, item_cap, w_cap, n_warps, n_mat;
    int ne;             // element positions of the record arrays (multiple of 16, > el_cap)
    size_t off_sv, off_rec, off_fw, off_mat, off_ent, off_items, off_emat, off_lcol, total;
};
template <int D>
inline TileSmem tile_smem_layout(int lv_cap, int el_cap, int ent_cap, int item_cap, int w_cap, int n_warps,
                                 int n_mat) {
    TileSmem s;
    s.lv_cap = lv_cap; s.el_cap = el_cap; s.ent_cap = ent_cap; s.item_cap = item_cap;
    s.w_cap = w_cap; s.n_warps = n_warps; s.n_mat = n_mat;
    auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t o = 0;
    s.off_sv = o;    o = up(o + (size_t)lv_cap * TileC<D>::VS * 8);
    s.ne = (el_cap + TILE_ROWS + 15) & ~15;
    s.off_rec = o;   o = up(o + (size_t)s.ne * TileC<D>::REC * 8);
    s.off_fw = o;    o = up(o + (size_t)n_warps * 32 * (D + 1) * 8);   // [warp][half][row][NB]
    s.off_mat = o;   o = up(o + (size_t)n_mat * TILE_MAT_STRIDE * 8);
    s.off_ent = o;   o = up(o + (size_t)ent_cap * 2);
    s.off_items = o; o = up(o + (size_t)item_cap * sizeof(TileItem));
    s.off_emat = o;  o = up(o + (size_t)(el_cap + TILE_ROWS));
    s.off_lcol = o;  o = up(o + (size_t)w_cap * TILE_ROWS * 2);
    s.total = o;
    return s;
}

// ---- host-side map (tilemap.cpp) -----------------------------------------------------------------
struct TileMapHost {
    std::vector<TileHdr> hdr;
    std::vector<int> tv;
    std::vector<unsigned long long> te;
    std::vector<TileItem> items;
    std::vector<uint16_t> ent;
    std::vector<uint16_t> lcol;        // [n_slots]
    int lv_cap = 0, el_cap = 0, ent_cap = 0, item_cap = 0, w_cap = 0, n_warps = 0, chunk = 0;
    bool ok = false;                   // false: some tile exceeds the 12-bit local index space
    std::string why;
};
// Pattern arrays are host copies of SellPattern; rows >= n_rows of the last slice are padding.
void tile_build_map(int dim, tl_i64 n_c, const int* cells, const int* cell_mat, int n_rows, int n_slices,
                    const tl_i64* slice_off, const int* slice_w, const int* col, const tl_i64* rowptr,
                    int n_warps, int chunk, int n_threads, TileMapHost& out);
