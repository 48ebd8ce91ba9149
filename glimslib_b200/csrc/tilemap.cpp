// Host-side construction of the tile-assembly maps (see tile.h).  Built once per mesh from the SELL pattern
// and the connectivity; replaces, for the tile kernel, both the element->slot scatter map and its transpose.
// Plain C++ (no CUDA) so that the CPU tests can link it into the emulator.
#include <string>
#include "tile.h"
#include <algorithm>
#include <thread>
#include <cstring>

namespace {

struct Chunk {       // per worker thread: concatenated outputs of a contiguous slice range
    std::vector<TileHdr> hdr;
    std::vector<int> tv;
    std::vector<unsigned long long> te;
    std::vector<TileItem> items;
    std::vector<uint16_t> ent;
    int lv_cap = 0, el_cap = 0, ent_cap = 0, item_cap = 0, sec_cap = 0;
    bool ok = true;
    std::string why;
};

void build_range(int dim, const int* cells, const int* cell_mat, int n_rows, const tl_i64* slice_off,
                 const int* slice_w, const int* col, const tl_i64* rowptr, const tl_i64* v2e_ptr, const int* v2e,
                 int n_warps, int chunk, int T0, int T1, uint16_t* lcol, Chunk& out) {
    const int nb = dim + 1;
    constexpr int TR = TILE_ROWS;
    std::vector<int> elems, lverts, bucket_of, pos_of;
    std::vector<std::vector<uint16_t>> lists;     // [j*TR + row] contributor entries
    std::vector<std::vector<unsigned char>> lmats;
    struct Part { int j, lo, hi, part, nparts; bool mixed; };
    std::vector<Part> parts;
    for (int T = T0; T < T1; ++T) {
        TileHdr h;
        std::memset(&h, 0, sizeof h);
        const int S = T >> 1, hf = T & 1;
        const int r0 = T * TR;
        const int nr = std::max(0, std::min(TR, n_rows - r0));
        const int w = slice_w[S];
        const tl_i64 base = slice_off[S] + hf * TR;     // slot of (column 0, row 0 of this tile)
        // elements touching the tile, ascending
        elems.clear();
        for (int l = 0; l < nr; ++l)
            for (tl_i64 t = v2e_ptr[r0 + l]; t < v2e_ptr[r0 + l + 1]; ++t) elems.push_back(v2e[t]);
        std::sort(elems.begin(), elems.end());
        elems.erase(std::unique(elems.begin(), elems.end()), elems.end());
        const int ne = (int)elems.size();
        // Element order inside the tile: shared-memory banks repeat every 16 doubles and a record has an odd
        // stride, so records whose positions differ mod 16 never collide.  Give every element a bucket
        // (= position mod 16) equal to one of its own rows -- the half-warp lanes that will read it -- choosing the
        // emptiest candidate so the buckets stay balanced; position = i*16 + bucket.
        bucket_of.assign(ne, 0);
        int cnt[TR] = {0};
        for (int i = 0; i < ne; ++i) {
            int best = -1;
            for (int a = 0; a < nb; ++a) {
                int v = cells[(tl_i64)elems[i] * nb + a];
                if (v >= r0 && v < r0 + nr) {
                    int m = v - r0;
                    if (best < 0 || cnt[m] < cnt[best]) best = m;
                }
            }
            bucket_of[i] = best;
            cnt[best]++;
        }
        int mxb = 0;
        for (int m = 0; m < TR; ++m) mxb = std::max(mxb, cnt[m]);
        pos_of.assign(ne, 0);
        int n_el;
        if (TR * mxb <= ne + ne / 4 + TR) {
            int fill[TR] = {0};
            for (int i = 0; i < ne; ++i) pos_of[i] = (fill[bucket_of[i]]++) * TR + bucket_of[i];
            n_el = TR * mxb;
        } else {                       // very uneven buckets: plain order (bank conflicts, but compact)
            for (int i = 0; i < ne; ++i) pos_of[i] = i;
            n_el = ne;
        }
        if (n_el + 1 > 4095) { out.ok = false; out.why = "a tile touches more than 4094 elements"; return; }
        // local vertices: own rows first, then the other vertices of the touching elements (ascending)
        lverts.clear();
        for (int i = 0; i < ne; ++i)
            for (int a = 0; a < nb; ++a) {
                int v = cells[(tl_i64)elems[i] * nb + a];
                if (!(v >= r0 && v < r0 + nr)) lverts.push_back(v);
            }
        std::sort(lverts.begin(), lverts.end());
        lverts.erase(std::unique(lverts.begin(), lverts.end()), lverts.end());
        const int n_lv = TR + (int)lverts.size();
        if (n_lv > 4095) { out.ok = false; out.why = "a tile references more than 4095 vertices"; return; }
        auto local_vertex = [&](int v) -> int {
            if (v >= r0 && v < r0 + nr) return v - r0;
            return TR + (int)(std::lower_bound(lverts.begin(), lverts.end(), v) - lverts.begin());
        };
        h.v_off = (tl_i64)out.tv.size();
        for (int l = 0; l < TR; ++l) out.tv.push_back(l < nr ? r0 + l : -1);
        out.tv.insert(out.tv.end(), lverts.begin(), lverts.end());
        h.n_lv = n_lv;
        // element records
        h.e_off = (tl_i64)out.te.size();
        out.te.resize(out.te.size() + n_el, TILE_NOELEM);
        for (int i = 0; i < ne; ++i) {
            unsigned long long r = 0;
            for (int a = 0; a < nb; ++a)
                r |= (unsigned long long)local_vertex(cells[(tl_i64)elems[i] * nb + a]) << (12 * a);
            r |= (unsigned long long)(cell_mat[elems[i]] & 0xff) << 48;
            out.te[h.e_off + pos_of[i]] = r;
        }
        h.n_el = n_el;
        // contributor lists per slot, and the local column of every slot (padding slots: own row)
        lists.assign((size_t)w * TR, std::vector<uint16_t>());
        lmats.assign((size_t)w * TR, std::vector<unsigned char>());
        for (int j = 0; j < w; ++j)
            for (int l = 0; l < TR; ++l) lcol[base + (tl_i64)j * 32 + l] = (uint16_t)l;
        for (int l = 0; l < nr; ++l) {
            const int r = r0 + l;
            const int len = (int)(rowptr[r + 1] - rowptr[r]);
            for (int j = 0; j < len; ++j) lcol[base + (tl_i64)j * 32 + l] = (uint16_t)local_vertex(col[base + (tl_i64)j * 32 + l]);
            for (tl_i64 t = v2e_ptr[r]; t < v2e_ptr[r + 1]; ++t) {
                const int e = v2e[t];
                const int le = pos_of[std::lower_bound(elems.begin(), elems.end(), e) - elems.begin()];
                int a = 0;
                for (int q = 0; q < nb; ++q) if (cells[(tl_i64)e * nb + q] == r) a = q;
                for (int b = 0; b < nb; ++b) {
                    const int cv = cells[(tl_i64)e * nb + b];
                    int lo = 0, hi = len;        // column index of cv in the row (SELL order = ascending columns)
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (col[base + (tl_i64)mid * 32 + l] < cv) lo = mid + 1; else hi = mid; }
                    lists[(size_t)lo * TR + l].push_back((uint16_t)(le | (a << 12) | (b << 14)));
                    lmats[(size_t)lo * TR + l].push_back((unsigned char)cell_mat[e]);
                }
            }
        }
        // parts: every column, long ones split into chunks of <= chunk iterations
        parts.clear();
        for (int j = 0; j < w; ++j) {
            int Lmax = 0;
            for (int l = 0; l < TR; ++l) Lmax = std::max(Lmax, (int)lists[(size_t)j * TR + l].size());
            int np = std::max(1, (Lmax + chunk - 1) / chunk);
            int step = (Lmax + np - 1) / np;
            for (int p = 0; p < np; ++p) {
                Part P{j, p * step, std::min(Lmax, (p + 1) * step), p, np, false};
                for (int l = 0; l < TR && !P.mixed; ++l) {
                    auto& lm = lmats[(size_t)j * TR + l];
                    int m0 = -1;
                    for (int q = P.lo; q < P.hi && q < (int)lm.size(); ++q) {
                        if (m0 < 0) m0 = lm[q];
                        else if (lm[q] != m0) { P.mixed = true; break; }
                    }
                }
                parts.push_back(P);
            }
        }
        // order of the half-items: secondaries, plain primaries (longest first, so paired halves have similar
        // lengths), split primaries -- those must start in a later round than the last secondary.
        std::vector<int> order;          // indices into parts, -1 = null half
        for (int i = 0; i < (int)parts.size(); ++i) if (parts[i].part > 0) order.push_back(i);
        const int n_secondary = (int)order.size();
        if (n_secondary > 255) { out.ok = false; out.why = "too many split columns in one tile"; return; }
        std::vector<int> first_sec(w, -1);
        for (int q = 0; q < n_secondary; ++q) { int j = parts[order[q]].j; if (first_sec[j] < 0) first_sec[j] = q; }
        {
            std::vector<int> plain;
            for (int i = 0; i < (int)parts.size(); ++i) if (parts[i].part == 0 && parts[i].nparts == 1) plain.push_back(i);
            std::stable_sort(plain.begin(), plain.end(), [&](int a, int b) { return parts[a].hi - parts[a].lo > parts[b].hi - parts[b].lo; });
            order.insert(order.end(), plain.begin(), plain.end());
        }
        bool any_split = false;
        for (auto& P : parts) any_split |= (P.part == 0 && P.nparts > 1);
        if (any_split) {
            const int last_sec_round = ((n_secondary - 1) / 2) / n_warps;
            while (((int)order.size() / 2) / n_warps <= last_sec_round) order.push_back(-1);
            for (int i = 0; i < (int)parts.size(); ++i) if (parts[i].part == 0 && parts[i].nparts > 1) order.push_back(i);
        }
        if (order.size() % 2) order.push_back(-1);
        h.item_off = (int)out.items.size();
        h.ent_off = (tl_i64)out.ent.size();
        for (size_t q = 0; q < order.size(); q += 2) {
            TileItem it;
            std::memset(&it, 0, sizeof it);
            int L = 0;
            for (int hh = 0; hh < 2; ++hh) {
                const int pi = order[q + hh];
                if (pi < 0) { it.kind[hh] = TILE_NULL; continue; }
                const Part& P = parts[pi];
                it.col_j[hh] = (uint16_t)P.j;
                L = std::max(L, P.hi - P.lo);
                if (P.mixed) it.mixed = 1;
                if (P.part > 0) {
                    it.kind[hh] = TILE_SECONDARY;
                    // buffer index = position among the secondaries (they lead `order`)
                    it.sec_idx[hh] = (uint8_t)(q + hh);
                } else if (P.nparts > 1) {
                    it.kind[hh] = TILE_PRIMARY_SPLIT;
                    it.sec_idx[hh] = (uint8_t)first_sec[P.j];
                    it.n_sec[hh] = (uint8_t)(P.nparts - 1);
                } else it.kind[hh] = TILE_PRIMARY;
            }
            it.L = (uint16_t)L;
            it.ent_off = (uint32_t)(out.ent.size() - (size_t)h.ent_off);
            for (int k = 0; k < L; ++k)
                for (int lane = 0; lane < 32; ++lane) {
                    const int pi = order[q + (lane >> 4)];
                    uint16_t e = (uint16_t)n_el;                 // sentinel: zero record, a = b = 0
                    if (pi >= 0) {
                        const Part& P = parts[pi];
                        auto& li = lists[(size_t)P.j * TR + (lane & 15)];
                        if (P.lo + k < P.hi && P.lo + k < (int)li.size()) e = li[P.lo + k];
                    }
                    out.ent.push_back(e);
                }
            out.items.push_back(it);
        }
        while (out.ent.size() % 8) out.ent.push_back((uint16_t)n_el);
        h.n_items = (int)out.items.size() - h.item_off;
        h.n_ent = (int)(out.ent.size() - (size_t)h.ent_off);
        h.n_sec = n_secondary;
        out.hdr.push_back(h);
        out.lv_cap = std::max(out.lv_cap, n_lv);
        out.el_cap = std::max(out.el_cap, n_el);
        out.ent_cap = std::max(out.ent_cap, h.n_ent);
        out.item_cap = std::max(out.item_cap, h.n_items);
        out.sec_cap = std::max(out.sec_cap, n_secondary);
    }
}

}  // namespace

void tile_build_map(int dim, tl_i64 n_c, const int* cells, const int* cell_mat, int n_rows, int n_slices,
                    const tl_i64* slice_off, const int* slice_w, const int* col, const tl_i64* rowptr,
                    int n_warps, int chunk, int n_threads, TileMapHost& out) {
    const int nb = dim + 1;
    // vertex -> element adjacency of the owned rows, elements ascending
    std::vector<tl_i64> v2e_ptr((size_t)n_rows + 1, 0);
    for (tl_i64 p = 0; p < n_c * nb; ++p) { int v = cells[p]; if (v < n_rows) v2e_ptr[v + 1]++; }
    for (int v = 0; v < n_rows; ++v) v2e_ptr[v + 1] += v2e_ptr[v];
    std::vector<int> v2e((size_t)v2e_ptr[n_rows]);
    {
        std::vector<tl_i64> fill(v2e_ptr.begin(), v2e_ptr.end() - 1);
        for (tl_i64 e = 0; e < n_c; ++e)
            for (int a = 0; a < nb; ++a) { int v = cells[e * nb + a]; if (v < n_rows) v2e[fill[v]++] = (int)e; }
    }
    out.lcol.assign((size_t)slice_off[n_slices], 0);
    const int n_tiles = 2 * n_slices;
    n_threads = std::max(1, std::min(n_threads, n_tiles));
    std::vector<Chunk> chunks(n_threads);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) {
        int T0 = (int)((tl_i64)n_tiles * t / n_threads), T1 = (int)((tl_i64)n_tiles * (t + 1) / n_threads);
        th.emplace_back(build_range, dim, cells, cell_mat, n_rows, slice_off, slice_w, col, rowptr, v2e_ptr.data(),
                        v2e.data(), n_warps, chunk, T0, T1, out.lcol.data(), std::ref(chunks[t]));
    }
    for (auto& t : th) t.join();
    out.ok = true;
    out.n_warps = n_warps; out.chunk = chunk;
    out.w_cap = 0;
    for (int S = 0; S < n_slices; ++S) out.w_cap = std::max(out.w_cap, slice_w[S]);
    out.hdr.clear(); out.tv.clear(); out.te.clear(); out.items.clear(); out.ent.clear();
    out.lv_cap = out.el_cap = out.ent_cap = out.item_cap = out.sec_cap = 0;
    for (auto& c : chunks) {
        if (!c.ok) { out.ok = false; out.why = c.why; return; }
        const tl_i64 bv = (tl_i64)out.tv.size(), be = (tl_i64)out.te.size(), bent = (tl_i64)out.ent.size();
        const int bi = (int)out.items.size();
        for (auto h : c.hdr) {
            h.v_off += bv; h.e_off += be; h.ent_off += bent; h.item_off += bi;
            out.hdr.push_back(h);
        }
        out.tv.insert(out.tv.end(), c.tv.begin(), c.tv.end());
        out.te.insert(out.te.end(), c.te.begin(), c.te.end());
        out.items.insert(out.items.end(), c.items.begin(), c.items.end());
        out.ent.insert(out.ent.end(), c.ent.begin(), c.ent.end());
        out.lv_cap = std::max(out.lv_cap, c.lv_cap); out.el_cap = std::max(out.el_cap, c.el_cap);
        out.ent_cap = std::max(out.ent_cap, c.ent_cap); out.item_cap = std::max(out.item_cap, c.item_cap);
        out.sec_cap = std::max(out.sec_cap, c.sec_cap);
        c = Chunk();
    }
}
