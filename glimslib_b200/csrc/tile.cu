// K1 + K2 fused "tile" assembly kernel (GLIMS_ASMK_TILE): one CTA per SELL slice, residual and Jacobian in one
// pass, no atomics, no zero-fill, fixed summation order.  The per-thread work lives in tile.h (shared with the CPU
// emulator of the tests); this file holds the kernel's orchestration, the device copies of the maps and the launch.
//
// Bounds (C4, per tile of 16 rows / ~240 slots / ~300 touching elements):
//   HBM  : 25 KB of matrix values out + ~7 KB of maps in   (algorithmic bytes: DESIGN.md section 5)
//   smem : 9 doubles per contributor (bucketed element order => lanes hit distinct banks)
//   FP64 : ~130 flops per staged element + 14 FMA per contributor + ~50 per slot
#include "common.h"
#include "tile.h"
#include <thread>
#include <cstdlib>

struct TileDev {
    TileHdr* hdr = nullptr;
    int* tv = nullptr;
    unsigned long long* te = nullptr;
    TileItem* items = nullptr;
    uint16_t* ent = nullptr;
    uint16_t* lcol = nullptr;
    int lv_cap = 0, el_cap = 0, ent_cap = 0, item_cap = 0, w_cap = 0, n_warps = 0, chunk = 0;
    bool ok = false;
    std::string why;
    size_t map_bytes = 0;
};

namespace {

struct TileArgs {
    const TileHdr* hdr; const int* tv; const unsigned long long* te; const TileItem* items;
    const uint16_t* ent; const uint16_t* lcol;
    const double* coords; const double* x; const double* xprev; const double* mat; const double* fext;
    const i64* slice_off; const int* slice_w;
    double *Kuu, *Kuc, *Kcc, *F;
    double dt;
    int n_mat, n_rows, what;
};

template <int D, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_assemble_tile(const TileArgs A, const TileSmem L) {
    constexpr int NB = D + 1, VS = TileC<D>::VS, KF = TileC<D>::KF, NW = NT / 32, TR = TILE_ROWS;
    const int NE = L.ne;
    extern __shared__ __align__(16) unsigned char smem[];
    double* sv = (double*)(smem + L.off_sv);
    double* rec = (double*)(smem + L.off_rec);
    double* fw = (double*)(smem + L.off_fw);
    double* smat = (double*)(smem + L.off_mat);
    uint16_t* sent = (uint16_t*)(smem + L.off_ent);
    TileItem* sitems = (TileItem*)(smem + L.off_items);
    unsigned char* emat = smem + L.off_emat;
    uint16_t* slcol = (uint16_t*)(smem + L.off_lcol);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hh = lane >> 4, row = lane & 15;
    const int T = blockIdx.x, S = T >> 1, hf = T & 1;
    const TileHdr h = A.hdr[T];
    const i64 sbase = A.slice_off[S];
    const int w = A.slice_w[S];
    const bool wkconst = A.what & GLIMS_ASM_KCONST, wkcc = A.what & GLIMS_ASM_KCC, res = A.what & GLIMS_ASM_RESIDUAL;

    // contributor entries and local columns: asynchronous 16-byte copies, consumed after phase A
    {
        const char* src = (const char*)(A.ent + h.ent_off);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sent);
        for (int i = tid; i < (h.n_ent >> 3); i += NT)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(src + (size_t)i * 16) : "memory");
        const uint16_t* src2 = A.lcol + sbase + hf * TR;
        const unsigned dst2 = (unsigned)__cvta_generic_to_shared(slcol);
        for (int i = tid; i < w * 2; i += NT)      // column j = i >> 1: 16 u16 = two 16-byte pieces
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst2 + i * 16),
                         "l"(src2 + (size_t)(i >> 1) * 32 + (i & 1) * 8) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    {
        const unsigned* src = (const unsigned*)(A.items + h.item_off);
        for (int i = tid; i < h.n_items * (int)(sizeof(TileItem) / 4); i += NT) ((unsigned*)sitems)[i] = src[i];
        for (int i = tid; i < A.n_mat * TILE_MAT_STRIDE; i += NT) smat[i] = A.mat[i];
    }
    // element records are fetched now (registers) so that their latency overlaps the vertex gathers
    constexpr int NPRE = (512 + NT - 1) / NT;
    unsigned long long tpre[NPRE];
#pragma unroll
    for (int q = 0; q < NPRE; ++q) {
        const int i = tid + q * NT;
        tpre[q] = i < h.n_el ? __ldg(&A.te[h.e_off + i]) : TILE_NOELEM;
    }
    // phase 0: local vertices
    for (int i = tid; i < h.n_lv; i += NT) tile_stage_vertex<D>(__ldg(&A.tv[h.v_off + i]), A.coords, A.x, A.xprev, sv + i * VS);
    __syncthreads();
    // phase A: element records (records n_el .. n_el+15 are the zero records the padding entries point at)
#pragma unroll
    for (int q = 0; q < NPRE; ++q) {
        const int i = tid + q * NT;
        if (i < h.n_el + TR) tile_stage_element<D>(tpre[q], sv, rec, NE, i, emat + i);
    }
    for (int i = tid + NPRE * NT; i < h.n_el + TR; i += NT)
        tile_stage_element<D>(i < h.n_el ? A.te[h.e_off + i] : TILE_NOELEM, sv, rec, NE, i, emat + i);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // phase B: warps loop over the items (two columns, or the two halves of one long column); no barriers
    double Facc[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) Facc[k] = 0.0;
    for (int idx = warp; idx < h.n_items; idx += NW) {
        const TileItem& it = sitems[idx];      // fields are read from shared memory (hh indexes col_j dynamically)
        const int flags = it.flags;
        if (flags & TILE_NULLITEM) continue;
        const int cj = it.col_j[hh];
        const int lcw = slcol[cj * TR + row];
        const int lc = lcw & ((1 << TILE_LCOL_BITS) - 1);
        const bool diag = lc == row;
        double kf[KF];
        tile_accumulate<D>(rec, NE, emat, smat, sent + it.ent_off, it.L, lane, (flags & TILE_MIXED) != 0, lcw >> TILE_LCOL_BITS,
                           diag, A.dt, kf);
        bool writer = hh == 0 || !(flags & (TILE_SPLIT | TILE_NULLB));
        if (flags & TILE_SPLIT) {
#pragma unroll
            for (int k = 0; k < KF; ++k) kf[k] += __shfl_xor_sync(0xffffffffu, kf[k], 16);
        }
        if (writer) {
            const double* xa = sv + row * VS;
            const double* xb = sv + lc * VS;
            const i64 g = sbase + (i64)cj * 32;
            const int sl = hf * TR + row;
            if (wkconst) {
                if (wkcc) { if (res) tile_finalize<D, true, true, true>(kf, diag, A.dt, xa, xb, g, sl, A.Kuu, A.Kuc, A.Kcc, Facc);
                            else tile_finalize<D, true, true, false>(kf, diag, A.dt, xa, xb, g, sl, A.Kuu, A.Kuc, A.Kcc, Facc); }
                else      { if (res) tile_finalize<D, true, false, true>(kf, diag, A.dt, xa, xb, g, sl, A.Kuu, A.Kuc, A.Kcc, Facc);
                            else tile_finalize<D, true, false, false>(kf, diag, A.dt, xa, xb, g, sl, A.Kuu, A.Kuc, A.Kcc, Facc); }
            } else {
                if (wkcc) { if (res) tile_finalize<D, false, true, true>(kf, diag, A.dt, xa, xb, g, sl, A.Kuu, A.Kuc, A.Kcc, Facc);
                            else tile_finalize<D, false, true, false>(kf, diag, A.dt, xa, xb, g, sl, A.Kuu, A.Kuc, A.Kcc, Facc); }
                else if (res) tile_finalize<D, false, false, true>(kf, diag, A.dt, xa, xb, g, sl, A.Kuu, A.Kuc, A.Kcc, Facc);
            }
        }
    }
    if (res) {
#pragma unroll
        for (int k = 0; k < NB; ++k) fw[(warp * 32 + lane) * NB + k] = Facc[k];
        __syncthreads();
        for (int t = tid; t < TR * NB; t += NT) {
            const int r = T * TR + t / NB;
            if (r < A.n_rows) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 2 * NW; ++q) s += fw[q * TR * NB + t];      // [warp][half] in order
                const i64 o = (i64)T * TR * NB + t;
                A.F[o] = A.fext ? s - A.fext[o] : s;
            }
        }
    }
}

template <typename T>
T* upload(const std::vector<T>& v, cudaStream_t st, size_t& bytes) {
    T* d = nullptr;
    size_t n = v.size() ? v.size() : 1;
    GL_CUDA(cudaMalloc(&d, sizeof(T) * n));
    if (v.size()) GL_CUDA(cudaMemcpyAsync(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
    bytes += sizeof(T) * n;
    return d;
}

int env_int(const char* name, int dflt) {
    const char* s = std::getenv(name);
    return s ? std::atoi(s) : dflt;
}

}  // namespace

void tile_free(glims_ctx* c) {
    TileDev* t = (TileDev*)c->tile;
    if (!t) return;
    for (void* q : {(void*)t->hdr, (void*)t->tv, (void*)t->te, (void*)t->items, (void*)t->ent, (void*)t->lcol})
        if (q) cudaFree(q);
    delete t;
    c->tile = nullptr;
}

// Build the maps on the host from the device pattern (downloaded once) and upload them.
static TileDev* tile_ensure(glims_ctx* c) {
    if (c->tile) return (TileDev*)c->tile;
    TileDev* t = new TileDev();
    c->tile = t;
    auto& p = c->pat;
    const int nb = c->nb;
    std::vector<i64> slice_off(p.n_slices + 1), rowptr(p.n_rows + 1);
    std::vector<int> slice_w(p.n_slices), col(p.n_slots), cells(c->n_c * nb), cell_mat(c->n_c);
    GL_CUDA(cudaStreamSynchronize(c->stream));
    GL_CUDA(cudaMemcpy(slice_off.data(), p.slice_off, sizeof(i64) * slice_off.size(), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(rowptr.data(), p.rowptr, sizeof(i64) * rowptr.size(), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(slice_w.data(), p.slice_w, sizeof(int) * slice_w.size(), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(col.data(), p.col, sizeof(int) * col.size(), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(cells.data(), c->cells, sizeof(int) * cells.size(), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(cell_mat.data(), c->cell_mat, sizeof(int) * cell_mat.size(), cudaMemcpyDeviceToHost));
    int nt = c->tile_nt ? c->tile_nt : env_int("GLIMS_TILE_NT", 192);
    if (nt != 128 && nt != 192 && nt != 256) nt = 192;
    const int chunk = std::max(1, c->tile_chunk ? c->tile_chunk : env_int("GLIMS_TILE_CH", 12));
    int hw = (int)std::thread::hardware_concurrency();
    hw = std::max(1, std::min(hw, 32));
    TileMapHost M;
    tile_build_map(c->dim, c->n_c, cells.data(), cell_mat.data(), p.n_rows, p.n_slices, slice_off.data(), slice_w.data(),
                   col.data(), rowptr.data(), nt / 32, chunk, hw, M);
    t->ok = M.ok; t->why = M.why;
    if (!M.ok) return t;
    t->lv_cap = M.lv_cap; t->el_cap = M.el_cap; t->ent_cap = M.ent_cap; t->item_cap = M.item_cap;
    t->w_cap = M.w_cap; t->n_warps = M.n_warps; t->chunk = M.chunk;
    // pad the entry array so the last slice's 16-byte copies stay inside the allocation
    M.ent.resize(M.ent.size() + 8, 0);
    M.lcol.resize(M.lcol.size() + 8, 0);
    t->hdr = upload(M.hdr, c->stream, t->map_bytes);
    t->tv = upload(M.tv, c->stream, t->map_bytes);
    t->te = upload(M.te, c->stream, t->map_bytes);
    t->items = upload(M.items, c->stream, t->map_bytes);
    t->ent = upload(M.ent, c->stream, t->map_bytes);
    t->lcol = upload(M.lcol, c->stream, t->map_bytes);
    GL_CUDA(cudaStreamSynchronize(c->stream));
    return t;
}

template <int D>
static bool tile_launch_dim(glims_ctx* c, TileDev* t, int what) {
    TileSmem L = tile_smem_layout<D>(t->lv_cap, t->el_cap, t->ent_cap, t->item_cap, t->w_cap, t->n_warps, c->n_mat);
    if (L.total > 227 * 1024) { t->ok = false; t->why = "tile: shared-memory footprint exceeds 227 KB"; return false; }
    TileArgs A;
    A.hdr = t->hdr; A.tv = t->tv; A.te = t->te; A.items = t->items; A.ent = t->ent; A.lcol = t->lcol;
    A.coords = c->coords; A.x = c->x; A.xprev = c->xprev; A.mat = c->mat; A.fext = c->have_load ? c->fext : nullptr;
    A.slice_off = c->pat.slice_off; A.slice_w = c->pat.slice_w;
    A.Kuu = c->Kuu; A.Kuc = c->Kuc; A.Kcc = c->Kcc; A.F = c->F;
    A.dt = c->dt; A.n_mat = c->n_mat; A.n_rows = c->pat.n_rows; A.what = what;
    const int nt = t->n_warps * 32;
#define TILE_GO(NT, MINB) do { auto kfn = k_assemble_tile<D, NT, MINB>; \
        GL_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total)); \
        kfn<<<2 * c->pat.n_slices, NT, L.total, c->stream>>>(A, L); } while (0)
    if (nt == 256) TILE_GO(256, 3);
    else if (nt == 192) TILE_GO(192, 3);
    else TILE_GO(128, 4);
#undef TILE_GO
    c->launches++;
    return true;
}

// returns false when the tile maps cannot represent this mesh (caller falls back to another variant)
bool launch_assemble_tile(glims_ctx* c, int what) {
    TileDev* t = tile_ensure(c);
    if (!t->ok) return false;
    bool ok = c->dim == 2 ? tile_launch_dim<2>(c, t, what) : tile_launch_dim<3>(c, t, what);
    if (ok) GL_CUDA(cudaGetLastError());
    return ok;
}

const char* tile_status(glims_ctx* c, long long* info) {
    TileDev* t = (TileDev*)c->tile;
    if (!t) return "not built";
    if (info) {
        TileSmem L = c->dim == 2 ? tile_smem_layout<2>(t->lv_cap, t->el_cap, t->ent_cap, t->item_cap, t->w_cap, t->n_warps, c->n_mat)
                                 : tile_smem_layout<3>(t->lv_cap, t->el_cap, t->ent_cap, t->item_cap, t->w_cap, t->n_warps, c->n_mat);
        info[0] = t->lv_cap; info[1] = t->el_cap; info[2] = t->ent_cap; info[3] = t->item_cap; info[4] = t->chunk;
        info[5] = (long long)L.total; info[6] = (long long)t->map_bytes; info[7] = t->n_warps * 32;
    }
    return t->ok ? "ok" : t->why.c_str();
}
