// CPU emulator of the tile-assembly kernel (test infrastructure, not product code): builds a SELL-32 pattern on
// the host, runs glimslib_b200/csrc/tilemap.cpp to get the tile maps, and then executes the very same per-thread
// phase functions the CUDA kernel calls (glimslib_b200/csrc/tile.h), one emulated CTA per slice.  The CPU tests
// compare its matrices and residual with the oracle, so maps, index arithmetic and formulas are verified
// without a GPU; what remains GPU-only is the launch configuration and the synchronisation.
#include <string>
#include "../../glimslib_b200/csrc/tile.h"
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cstdlib>

typedef tl_i64 i64;

namespace {
long long g_bank_wf = 0, g_bank_n = 0, g_hist[9] = {0}; int g_dumped = 0;

template <int D>
int emulate(const TileMapHost& M, int n_rows, int n_slices, const i64* slice_off, const int* slice_w,
            const double* coords, const double* x, const double* xprev, const double* mat6, int n_mat, double dt,
            const double* fext, double* Kuu, double* Kuc, double* Kcc, double* F) {
    constexpr int NB = D + 1, GS = TileC<D>::GS, VS = TileC<D>::VS, KF = TileC<D>::KF, TR = TILE_ROWS;
    const int NW = M.n_warps;
    TileSmem L = tile_smem_layout<D>(M.lv_cap, M.el_cap, M.ent_cap, M.item_cap, M.w_cap, NW, n_mat);
    std::vector<unsigned char> smem(L.total);
    for (int T = 0; T < 2 * n_slices; ++T) {
        const int S = T >> 1, hf = T & 1;
        std::fill(smem.begin(), smem.end(), 0xCD);     // poison: reading anything unstaged shows up as garbage
        double* sv = (double*)(smem.data() + L.off_sv);
        double* rec = (double*)(smem.data() + L.off_rec);
        double* fw = (double*)(smem.data() + L.off_fw);
        double* smat = (double*)(smem.data() + L.off_mat);
        uint16_t* sent = (uint16_t*)(smem.data() + L.off_ent);
        TileItem* sitems = (TileItem*)(smem.data() + L.off_items);
        unsigned char* emat = smem.data() + L.off_emat;
        uint16_t* slcol = (uint16_t*)(smem.data() + L.off_lcol);
        const TileHdr h = M.hdr[T];
        if (h.n_lv > L.lv_cap || h.n_el > L.el_cap || h.n_ent > L.ent_cap || h.n_items > L.item_cap) return -10;
        if (slice_w[S] > L.w_cap) return -15;
        std::memcpy(sent, M.ent.data() + h.ent_off, (size_t)h.n_ent * 2);
        std::memcpy(sitems, M.items.data() + h.item_off, (size_t)h.n_items * sizeof(TileItem));
        const i64 sbase = slice_off[S];
        for (int j = 0; j < slice_w[S]; ++j)
            std::memcpy(slcol + j * TR, M.lcol.data() + sbase + (i64)j * 32 + hf * TR, TR * 2);
        std::memcpy(smat, mat6, sizeof(double) * n_mat * TILE_MAT_STRIDE);
        for (int i = 0; i < h.n_lv; ++i) tile_stage_vertex<D>(M.tv[h.v_off + i], coords, x, xprev, sv + i * VS);
        for (int i = 0; i < h.n_el + TR; ++i)
            tile_stage_element<D>(i < h.n_el ? M.te[h.e_off + i] : TILE_NOELEM, sv, rec, L.ne, i, emat + i);
        std::vector<double> Facc((size_t)NW * 32 * NB, 0.0);
        static const bool bank_stats = getenv("TILE_EMU_BANKS") != nullptr;
        for (int idx = 0; idx < h.n_items; ++idx) {
            const int warp = idx % NW;
            const TileItem it = sitems[idx];
            if (it.flags & TILE_NULLITEM) continue;
            if (bank_stats && getenv("TILE_EMU_DUMP") && T == atoi(getenv("TILE_EMU_DUMP"))) {
                fprintf(stderr, "tile %d item %d cols %d,%d L %d flags %d\n", T, idx, it.col_j[0], it.col_j[1], it.L, it.flags);
                for (int j = 0; j < it.L; ++j) {
                    fprintf(stderr, "  j=%d:", j);
                    for (int lane = 0; lane < 32; ++lane) {
                        const unsigned e = sent[it.ent_off + j * 32 + lane];
                        fprintf(stderr, " %d.%d%d", (int)(e & 0xfff), (int)((e >> 12) & 3), (int)(e >> 14));
                        if (lane == 15) fprintf(stderr, " |");
                    }
                    fprintf(stderr, "\n");
                }
            }
            if (bank_stats) {       // wavefronts of the phase-B LDS.64 loads: per half-warp, max lanes per 8-byte bank
                for (int j = 0; j < it.L; ++j)
                    for (int hh = 0; hh < 2; ++hh)
                        for (int q = 0; q < 2 * D + 3; ++q) {
                            int cnt[16] = {0}, mx = 0;
                            std::vector<int> seen;
                            for (int row = 0; row < 16; ++row) {
                                const unsigned e = sent[it.ent_off + j * 32 + hh * 16 + row];
                                const int lel = e & 0xfff, a = (e >> 12) & 3, b = e >> 14;
                                int w = q < D ? (a * L.ne + lel) * GS + q : q < 2 * D ? (b * L.ne + lel) * GS + (q - D) : NB * L.ne * GS + (q - 2 * D) * L.ne + lel;
                                bool dup = false;
                                for (int s2 : seen) if (s2 == w) dup = true;   // same address: broadcast
                                if (dup) continue;
                                seen.push_back(w);
                                mx = std::max(mx, ++cnt[w & 15]);
                            }
                            g_bank_wf += mx; g_bank_n += 1; g_hist[std::min(mx, 8)]++;
                            if (mx == 2 && T >= 4000 && T < 4040 && q == 0 && g_dumped < 24 && getenv("TILE_EMU_WORST")) { g_dumped++; fprintf(stderr, "T=%d item=%d j=%d hh=%d q=%d mx=%d:", T, idx, j, hh, q, mx); for (int row = 0; row < 16; ++row) { unsigned e = sent[it.ent_off + j * 32 + hh * 16 + row]; fprintf(stderr, " %d.%d%d", (int)(e & 0xfff), (int)((e >> 12) & 3), (int)(e >> 14)); } fprintf(stderr, "\n"); }
                        }
            }
            double kfs[32][KF];
            int lcs[32];
            for (int lane = 0; lane < 32; ++lane) {
                const int hh = lane >> 4, row = lane & 15;
                const int cj = it.col_j[hh];
                if (cj >= slice_w[S]) return -11;
                const int lcw = slcol[cj * TR + row];
                const int lc = lcw & ((1 << TILE_LCOL_BITS) - 1);
                if (lc >= h.n_lv) return -14;
                lcs[lane] = lc;
                tile_accumulate<D>(rec, L.ne, emat, smat, sent + it.ent_off, it.L, lane, (it.flags & TILE_MIXED) != 0,
                                   lcw >> TILE_LCOL_BITS, lc == row, dt, kfs[lane]);
            }
            if (it.flags & TILE_SPLIT) {
                if (it.col_j[0] != it.col_j[1]) return -12;
                for (int lane = 0; lane < 16; ++lane)
                    for (int k = 0; k < KF; ++k) { double s2 = kfs[lane][k] + kfs[lane + 16][k]; kfs[lane][k] = s2; kfs[lane + 16][k] = s2; }
            }
            for (int lane = 0; lane < 32; ++lane) {
                const int hh = lane >> 4, row = lane & 15;
                const bool writer = hh == 0 || !(it.flags & (TILE_SPLIT | TILE_NULLB));
                if (!writer) continue;
                double (&fa)[NB] = *reinterpret_cast<double (*)[NB]>(&Facc[((size_t)warp * 32 + lane) * NB]);
                tile_finalize<D, true, true, true>(kfs[lane], lcs[lane] == row, dt, sv + row * VS, sv + lcs[lane] * VS,
                                                   sbase + (i64)it.col_j[hh] * 32, hf * TR + row, Kuu, Kuc, Kcc, fa);
            }
        }
        for (size_t t = 0; t < Facc.size(); ++t) fw[t] = Facc[t];
        for (int t = 0; t < TR * NB; ++t) {
            const int r = T * TR + t / NB;
            if (r >= n_rows) continue;
            double s = 0.0;
            for (int q = 0; q < 2 * NW; ++q) s += fw[q * TR * NB + t];
            F[(i64)r * NB + (t % NB)] = s - (fext ? fext[(i64)r * NB + (t % NB)] : 0.0);
        }
    }
    return 0;
}

}  // namespace

extern "C" int tile_emu_assemble(int dim, i64 n_v, i64 n_rows, const double* coords, i64 n_c, const int* cells,
                                 const int* cell_mat, int n_mat, const double* table5, double dt, const double* x,
                                 const double* xprev, const double* fext, int n_warps, int chunk, int n_threads,
                                 i64* rowptr /* n_rows+1 */, i64 nnzb_cap, int* colidx, double* Kuu, double* Kuc,
                                 double* Kcc, double* F, i64* info /* 8 */) {
    const int nb = dim + 1, DD = dim * dim;
    // host SELL-32 pattern of the vertex graph (same rules as csrc/pattern.cu)
    std::vector<std::vector<int>> rows(n_rows);
    for (i64 e = 0; e < n_c; ++e)
        for (int a = 0; a < nb; ++a) {
            int va = cells[e * nb + a];
            if (va >= n_rows) continue;
            for (int b = 0; b < nb; ++b) rows[va].push_back(cells[e * nb + b]);
        }
    rowptr[0] = 0;
    for (i64 r = 0; r < n_rows; ++r) {
        auto& v = rows[r];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        rowptr[r + 1] = rowptr[r] + (i64)v.size();
    }
    if (rowptr[n_rows] > nnzb_cap) return -1;
    const int n_slices = (int)((n_rows + 31) / 32);
    std::vector<int> slice_w(n_slices, 0);
    std::vector<i64> slice_off(n_slices + 1, 0);
    for (i64 r = 0; r < n_rows; ++r) slice_w[r >> 5] = std::max(slice_w[r >> 5], (int)rows[r].size());
    for (int S = 0; S < n_slices; ++S) slice_off[S + 1] = slice_off[S] + (i64)slice_w[S] * 32;
    const i64 n_slots = slice_off[n_slices];
    std::vector<int> col(n_slots);
    for (int S = 0; S < n_slices; ++S)
        for (int t = 0; t < slice_w[S] * 32; ++t) {
            i64 r = (i64)S * 32 + (t & 31);
            col[slice_off[S] + t] = r < n_rows ? (int)r : 0;
        }
    for (i64 r = 0; r < n_rows; ++r)
        for (size_t j = 0; j < rows[r].size(); ++j) {
            col[slice_off[r >> 5] + (i64)j * 32 + (r & 31)] = rows[r][j];
            colidx[rowptr[r] + (i64)j] = rows[r][j];
        }
    TileMapHost M;
    tile_build_map(dim, n_c, cells, cell_mat, (int)n_rows, n_slices, slice_off.data(), slice_w.data(), col.data(), rowptr,
                   n_warps, chunk, n_threads, M);
    if (!M.ok) { fprintf(stderr, "tile map: %s\n", M.why.c_str()); return -2; }
    std::vector<double> mat6((size_t)n_mat * TILE_MAT_STRIDE);
    for (int m = 0; m < n_mat; ++m) {
        for (int k = 0; k < 5; ++k) mat6[m * 6 + k] = table5[m * 5 + k];
        mat6[m * 6 + 5] = (2.0 * table5[m * 5] + dim * table5[m * 5 + 1]) * table5[m * 5 + 4];
    }
    std::vector<double> sKuu((size_t)n_slots * DD, 1e300), sKuc((size_t)n_slots * dim, 1e300), sKcc((size_t)n_slots, 1e300);
    int rc = dim == 2 ? emulate<2>(M, (int)n_rows, n_slices, slice_off.data(), slice_w.data(), coords, x, xprev, mat6.data(), n_mat, dt, fext, sKuu.data(), sKuc.data(), sKcc.data(), F)
                      : emulate<3>(M, (int)n_rows, n_slices, slice_off.data(), slice_w.data(), coords, x, xprev, mat6.data(), n_mat, dt, fext, sKuu.data(), sKuc.data(), sKcc.data(), F);
    if (rc) return rc;
    // every padded slot must have been written, and padding must be zero
    i64 bad = 0;
    for (i64 s = 0; s < n_slots; ++s) if (sKcc[s] == 1e300) bad++;
    for (i64 r = 0; r < n_rows; ++r) {
        i64 base = slice_off[r >> 5];
        int len = (int)rows[r].size();
        for (int j = 0; j < slice_w[r >> 5]; ++j) {
            i64 s = base + (i64)j * 32 + (r & 31);
            i64 g = s & ~(i64)31; int lane = (int)(s & 31);
            if (j < len) {
                i64 t = rowptr[r] + j;
                for (int k = 0; k < DD; ++k) Kuu[t * DD + k] = sKuu[g * DD + (i64)k * 32 + lane];
                for (int k = 0; k < dim; ++k) Kuc[t * dim + k] = sKuc[g * dim + (i64)k * 32 + lane];
                Kcc[t] = sKcc[s];
            } else if (sKcc[s] != 0.0 || sKuu[g * DD + lane] != 0.0) bad++;
        }
    }
    if (getenv("TILE_EMU_BANKS")) { fprintf(stderr, "wavefront histogram:"); for (int i = 1; i < 9; ++i) fprintf(stderr, " %d:%.3f", i, (double)g_hist[i] / std::max<long long>(1, g_bank_n)); fprintf(stderr, "\n"); }
    if (getenv("TILE_EMU_BANKS")) fprintf(stderr, "phase-B LDS.64 half-warp wavefronts: %.3f per access (1.0 = conflict-free)\n", (double)g_bank_wf / std::max<long long>(1, g_bank_n));
    if (info) {
        TileSmem L = dim == 2 ? tile_smem_layout<2>(M.lv_cap, M.el_cap, M.ent_cap, M.item_cap, M.w_cap, n_warps, n_mat)
                              : tile_smem_layout<3>(M.lv_cap, M.el_cap, M.ent_cap, M.item_cap, M.w_cap, n_warps, n_mat);
        info[0] = M.lv_cap; info[1] = M.el_cap; info[2] = M.ent_cap; info[3] = M.item_cap; info[4] = 0; for (auto& it : M.items) if (it.flags & TILE_SPLIT) info[4]++;
        info[5] = (i64)L.total; info[6] = bad; info[7] = (i64)M.ent.size();
        if (getenv("TILE_EMU_STATS")) {
            std::vector<int> hel(64, 0), hent(64, 0), hlv(64, 0), hsec(16, 0), hit(64, 0);
            for (auto& h : M.hdr) { hel[std::min(63, h.n_el / 16)]++; hent[std::min(63, h.n_ent / 256)]++; hlv[std::min(63, h.n_lv / 16)]++; hit[std::min(63, h.n_items)]++; }
            fprintf(stderr, "n_el/32 hist:"); for (int i = 0; i < 64; ++i) if (hel[i]) fprintf(stderr, " %d:%d", i * 16, hel[i]);
            fprintf(stderr, "\nn_ent/256 hist:"); for (int i = 0; i < 64; ++i) if (hent[i]) fprintf(stderr, " %d:%d", i * 256, hent[i]);
            fprintf(stderr, "\nn_lv/16 hist:"); for (int i = 0; i < 64; ++i) if (hlv[i]) fprintf(stderr, " %d:%d", i * 16, hlv[i]);
            fprintf(stderr, "\nn_sec hist:"); for (int i = 0; i < 16; ++i) if (hsec[i]) fprintf(stderr, " %d:%d", i, hsec[i]);
            fprintf(stderr, "\nn_items hist:"); for (int i = 0; i < 64; ++i) if (hit[i]) fprintf(stderr, " %d:%d", i, hit[i]);
            fprintf(stderr, "\n");
        }
    }
    (void)n_v;
    return bad ? -20 : 0;
}
