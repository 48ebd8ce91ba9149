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
    int lv_cap = 0, el_cap = 0, ent_cap = 0, item_cap = 0;
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
        if (n_lv > (1 << TILE_LCOL_BITS)) { out.ok = false; out.why = "a tile references more than 1024 vertices"; return; }
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
                const int le = (int)(std::lower_bound(elems.begin(), elems.end(), e) - elems.begin());   // index, not yet a position
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
        // columns: length, mixed-material flag, slot material (stored in the high bits of lcol)
        struct Col { int j, L; bool mixed, split; };
        std::vector<Col> cols;
        for (int j = 0; j < w; ++j) {
            Col C{j, 0, false, false};
            for (int l = 0; l < TR; ++l) {
                auto& lm = lmats[(size_t)j * TR + l];
                C.L = std::max(C.L, (int)lm.size());
                int m0 = lm.empty() ? 0 : lm[0];
                for (unsigned char m : lm) if (m != m0) C.mixed = true;
                lcol[base + (tl_i64)j * 32 + l] |= (uint16_t)(m0 << TILE_LCOL_BITS);
            }
            C.split = C.L > chunk;
            cols.push_back(C);
        }
        // items: a split column takes a whole warp (first / second half of its contributors per half-warp);
        // the others are paired by decreasing length so that the two halves of a warp have similar trip counts.
        std::vector<int> plain;
        for (auto& C : cols) if (!C.split) plain.push_back(C.j);
        std::stable_sort(plain.begin(), plain.end(), [&](int a, int b) { return cols[a].L > cols[b].L; });
        struct Item { int jA, jB, L; bool split; };
        std::vector<Item> its;
        for (auto& C : cols) if (C.split) its.push_back(Item{C.j, C.j, (C.L + 1) / 2, true});
        for (size_t q = 0; q < plain.size(); q += 2) {
            int jA = plain[q], jB = q + 1 < plain.size() ? plain[q + 1] : -1;
            its.push_back(Item{jA, jB, std::max(cols[jA].L, jB >= 0 ? cols[jB].L : 0), false});
        }
        std::stable_sort(its.begin(), its.end(), [](const Item& a, const Item& b) { return a.L > b.L; });
        {
            // Warp w processes items w, w + n_warps, ...: balance the warps (longest-processing-time first; an item
            // costs its iterations plus a fixed share for the conversion and the stores) and lay the list out so
            // that the stride-n_warps walk realises the assignment; short bins are padded with null items.
            const int fixed = 4;
            std::vector<std::vector<Item>> bins(n_warps);
            std::vector<int> load(n_warps, 0);
            for (auto& I : its) {
                int w = 0;
                for (int q = 1; q < n_warps; ++q) if (load[q] < load[w]) w = q;
                bins[w].push_back(I);
                load[w] += I.L + fixed;
            }
            size_t depth = 0;
            for (auto& b : bins) depth = std::max(depth, b.size());
            its.clear();
            for (size_t k = 0; k < depth; ++k)
                for (int w = 0; w < n_warps; ++w)
                    its.push_back(k < bins[w].size() ? bins[w][k] : Item{-1, -1, 0, false});
            while (!its.empty() && its.back().jA < 0) its.pop_back();
        }
        // Element order inside the tile.  Shared-memory banks repeat every 16 doubles and the record arrays have odd
        // strides, so two elements collide iff their positions agree mod 16.  The 16 lanes of a half-warp read, in
        // iteration k of an item, 16 (mostly different) elements: walk those access groups and give every element a
        // bucket (= position mod 16) no other member of its group has yet, preferring the emptiest; position =
        // rank*16 + bucket.  On structured meshes this is conflict-free; elsewhere it is a greedy best effort.
        const int cap = (ne + TR - 1) / TR + 5;
        bucket_of.assign(ne, -1);
        int cnt[TR] = {0};
        // preferred bucket: the smallest own row (translates of one element then get distinct buckets); elements of
        // row 0 that contain the vertex before the tile would belong to "row -1" and prefer the last bucket
        std::vector<int> pref(ne);
        for (int i = 0; i < ne; ++i) {
            int first = TR;
            bool has_prev = false;
            for (int a = 0; a < nb; ++a) {
                int v = cells[(tl_i64)elems[i] * nb + a];
                if (v >= r0 && v < r0 + nr) first = std::min(first, v - r0);
                if (v == r0 - 1) has_prev = true;
            }
            pref[i] = (first == 0 && has_prev) ? nr - 1 : first;
        }
        for (auto& I : its)
            for (int k = 0; k < I.L; ++k)
                for (int hh = 0; hh < 2; ++hh) {
                    const int j = hh == 0 ? I.jA : I.jB;   // null items have L == 0
                    if (j < 0) continue;
                    const int q = I.split ? hh * I.L + k : k;
                    unsigned used = 0;
                    for (int l = 0; l < TR; ++l) {
                        auto& li = lists[(size_t)j * TR + l];
                        if (q < (int)li.size() && bucket_of[li[q] & 0xfff] >= 0) used |= 1u << bucket_of[li[q] & 0xfff];
                    }
                    for (int l = 0; l < TR; ++l) {
                        auto& li = lists[(size_t)j * TR + l];
                        if (q >= (int)li.size()) continue;
                        const int e = li[q] & 0xfff;
                        if (bucket_of[e] >= 0) continue;
                        int pick = -1, any = 0;
                        if (!(used >> pref[e] & 1u) && cnt[pref[e]] < cap) pick = pref[e];
                        else for (int m = 0; m < TR; ++m) {
                            if (cnt[m] < cnt[any]) any = m;
                            if (!(used >> m & 1u) && cnt[m] < cap && (pick < 0 || cnt[m] < cnt[pick])) pick = m;
                        }
                        if (pick < 0) pick = any;
                        bucket_of[e] = pick;
                        cnt[pick]++;
                        used |= 1u << pick;
                    }
                }
        for (int i = 0; i < ne; ++i) if (bucket_of[i] < 0) { int any = 0; for (int m = 1; m < TR; ++m) if (cnt[m] < cnt[any]) any = m; bucket_of[i] = any; cnt[any]++; }
        int mxb = 0;
        for (int m = 0; m < TR; ++m) mxb = std::max(mxb, cnt[m]);
        pos_of.assign(ne, 0);
        {
            int fill[TR] = {0};
            for (int i = 0; i < ne; ++i) pos_of[i] = (fill[bucket_of[i]]++) * TR + bucket_of[i];
        }
        const int n_el = TR * mxb;
        if (n_el + TR > 4095) { out.ok = false; out.why = "a tile touches more than 4094 elements"; return; }
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
        h.item_off = (int)out.items.size();
        h.ent_off = (tl_i64)out.ent.size();
        for (auto& I : its) {
            TileItem it;
            std::memset(&it, 0, sizeof it);
            if (I.jA < 0) { it.flags = TILE_NULLITEM; out.items.push_back(it); continue; }
            it.col_j[0] = (uint16_t)I.jA;
            it.col_j[1] = (uint16_t)(I.jB >= 0 ? I.jB : I.jA);
            it.L = (uint16_t)I.L;
            it.flags = (uint8_t)((cols[I.jA].mixed || (I.jB >= 0 && cols[I.jB].mixed) ? TILE_MIXED : 0) |
                                 (I.split ? TILE_SPLIT : 0) | (I.jB < 0 ? TILE_NULLB : 0));
            it.ent_off = (uint32_t)(out.ent.size() - (size_t)h.ent_off);
            for (int k = 0; k < I.L; ++k) {
                uint16_t row_e[32];
                unsigned used[2] = {0, 0};
                for (int lane = 0; lane < 32; ++lane) {
                    const int hh = lane >> 4;
                    const int j = hh == 0 ? I.jA : I.jB;
                    row_e[lane] = 0xffff;
                    if (j >= 0) {
                        auto& li = lists[(size_t)j * TR + (lane & 15)];
                        const int q = I.split ? hh * I.L + k : k;
                        if (q < (int)li.size()) {
                            row_e[lane] = (uint16_t)((li[q] & 0xf000) | pos_of[li[q] & 0xfff]);
                            used[hh] |= 1u << (pos_of[li[q] & 0xfff] & 15);
                        }
                    }
                }
                // padding lanes read one of the 16 all-zero records n_el..n_el+15: the one in a bucket no real
                // element of this half-warp access uses (a = b = 0)
                for (int lane = 0; lane < 32; ++lane) {
                    if (row_e[lane] != 0xffff) continue;
                    const int hh = lane >> 4;
                    int m = 0;
                    while (m < TR - 1 && (used[hh] >> m & 1u)) ++m;
                    used[hh] |= 1u << m;
                    row_e[lane] = (uint16_t)(n_el + m);
                }
                for (int lane = 0; lane < 32; ++lane) out.ent.push_back(row_e[lane]);
            }
            out.items.push_back(it);
        }
        while (out.ent.size() % 8) out.ent.push_back((uint16_t)n_el);
        h.n_items = (int)out.items.size() - h.item_off;
        h.n_ent = (int)(out.ent.size() - (size_t)h.ent_off);
        out.hdr.push_back(h);
        out.lv_cap = std::max(out.lv_cap, n_lv);
        out.el_cap = std::max(out.el_cap, n_el);
        out.ent_cap = std::max(out.ent_cap, h.n_ent);
        out.item_cap = std::max(out.item_cap, h.n_items);
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
    out.lv_cap = out.el_cap = out.ent_cap = out.item_cap = 0;
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
        c = Chunk();
    }
}
