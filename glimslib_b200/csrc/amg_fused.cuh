// Levels >= 2 of the FP32 V-cycle in ONE persistent kernel (included by amg.cu).
//
// The coarse levels are a few thousand block rows: every kernel on them is launch / latency bound (about twenty launches
// per V-cycle at ~15 us each, 20 % of a PCG iteration at C4 on one GPU and most of it on eight).  Here the restriction from
// level 1, every smoother step, residual, restriction, the dense coarsest solve, every prolongation and the prolongation
// back into level 1 are phases of one kernel, separated by a software grid barrier (sense-reversing, self-resetting, so
// the kernel is a plain launch that can sit inside the captured PCG graphs and their conditional bodies).  The grid is
// sized to be co-resident (occupancy API); vectors written in one phase and read in a later one go through L2
// (ld.global.cg), matrices and transfer maps through the read-only path.
//
// CTA = 32*BS threads (BS = 6 in 3D, 3 in 2D): warp i handles component i of the 32 block rows of a SELL slice, exactly
// the mapping of k_spmv32_split_cheb; restriction = one warp per aggregate, prolongation = one thread per node.
#pragma once
// (included inside amg.cu's anonymous namespace)

enum { FOP_RESTRICT = 0, FOP_FIRST, FOP_STEP, FOP_RESID, FOP_DENSE, FOP_PROLONG, FOP_COPY, FOP_GATHER };
constexpr int FUSED_MAX_LEVELS = 12, FUSED_MAX_PHASES = 96;

struct FusedLevel {
    int n, n_slices, bs;        // n: nodes of the level (vector length / bs); n_slices: slices of the rows this rank applies
    int n_own, row0;            // rows held in the matrix arrays and their first global row (whole level: n, 0)
    const i64* slice_off; const int* slice_w; const int* col;
    const float* A; const float* dinv;
    float *x, *y, *b, *r, *d;
    // transfer to the next (coarser) level
    int nc; const int* mem_ptr; const int* mem_idx; const int* agg; const double* rvec;
};
struct FusedPhase { int op, lvl; float* src; float* dst; float c1, c2; };
struct FusedPlan {
    int n_levels, n_phases, coarse_m;
    const float* coarse_inv;
    FusedLevel L[FUSED_MAX_LEVELS];
    FusedPhase P[FUSED_MAX_PHASES];
    unsigned* bar;          // [0] arrivals, [1] generation, [2] error (a barrier wait timed out)
    unsigned long long* tstamp;     // [n_phases+1] %globaltimer of CTA 0 at the start and after every phase (diagnostics)
    // partitioned mesh, level 1 inside the kernel: its vectors live in a symmetric buffer, FOP_GATHER all-gathers one of them
    SymView sym;
    int n_gathers;
    long long seg;                  // floats per rank segment of a level-1 vector
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all CTAs of the (co-resident) grid; sense-reversing: leaves the arrival counter at zero.  `gen` is the generation this CTA
// expects (read once at kernel start, then tracked in a register: one L2 round trip less per barrier); the arrival is a
// release-atomic, the wait an acquire-load.  The wait is bounded (about two seconds of SM clock): a grid that is not
// co-resident after all raises bar[2] (amg_fused_failed) instead of hanging the GPU.
__device__ __forceinline__ void fused_grid_barrier(unsigned* bar, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned prev;
        asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
        if (prev == gridDim.x - 1) {
            bar[0] = 0u;
            st_release_gpu_u32(bar + 1, gen + 1u);
        } else {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(bar + 1) == gen) {
                if (clock64() - t0 > 4000000000LL) { atomicExch(bar + 2, 1u); break; }
            }
        }
        __threadfence();
    }
    gen += 1u;
    __syncthreads();
}

// Vectors written by other CTAs in an earlier phase: a plain (L1-cached) load is safe because the grid barrier's acquire
// (thread 0: ld.acquire.gpu + fence => CCTL.IVALL, then bar.sync) invalidates the SM's L1 -- the cooperative-groups
// grid.sync discipline.  Going through L1 matters: the six component warps of a slice and neighbouring rows gather the
// same x blocks, and with L2-only loads (ld.cg) the ~90 KB vector became an L2 hot spot (3 M sector requests per phase).
template <typename T> __device__ __forceinline__ T vload(const T* p) { return *p; }
constexpr int FUSED_CG = 4;      // column groups per block-row component: a slice's columns are split over CG warps
constexpr int FUSED_U = 6;       // block columns a warp holds in flight per round

// partial sum of component i of (A x) for row `lane` of a slice over the columns g*U + u + k*CG*U; x through L2.
// One round covers CG*U = 24 columns, i.e. a whole row of these levels: two dependent load latencies (column -> x).
template <int BS>
__device__ __forceinline__ float fused_row_dot(const i64 base, const int w, const int* __restrict__ col,
                                               const float* __restrict__ A, const float* x, const int i, const int g,
                                               const int lane, int* scol /* this warp's [U][32] staging area */) {
    constexpr int U = FUSED_U;
    float s = 0.f;
    for (int j0 = g * U; j0 < w; j0 += FUSED_CG * U) {
        int cc[U];
        float av[U][BS], xv[U][BS];
#pragma unroll
        for (int u = 0; u < U; ++u) cc[u] = __ldg(&col[base + (i64)min(j0 + u, w - 1) * 32 + lane]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float* Au = A + (base + (i64)min(j0 + u, w - 1) * 32) * (BS * BS) + (i * BS) * 32 + lane;
#pragma unroll
            for (int b = 0; b < BS; ++b) av[u][b] = __ldg(&Au[b * 32]);
        }
        // A warp issues in order and the x loads depend on the column indices.  Left alone, the scheduler interleaves
        // "column u, x of column u" and the chain becomes U load latencies long.  Passing the indices through shared
        // memory around a warp barrier pins the order: all U column loads first, then all x loads -- two latencies.
#pragma unroll
        for (int u = 0; u < U; ++u) scol[u * 32 + lane] = cc[u];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; ++u) cc[u] = scol[u * 32 + lane];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int b = 0; b < BS; ++b) xv[u][b] = vload(&x[(i64)cc[u] * BS + b]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float on = (j0 + u < w) ? 1.f : 0.f;
#pragma unroll
            for (int b = 0; b < BS; ++b) s += on * av[u][b] * xv[u][b];
        }
    }
    return s;
}

__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The plan travels as a kernel parameter (constant bank): phase and level descriptors cost no global-memory round trip.
template <int D>
__global__ void __launch_bounds__(32 * ((D == 2) ? 3 : 6) * FUSED_CG)
k_amg_fused(const __grid_constant__ FusedPlan plan) {
    constexpr int BS = (D == 2) ? 3 : 6, NW = BS * FUSED_CG, NT = 32 * NW;
    __shared__ float part[FUSED_CG][BS][32];
    __shared__ int scol_all[NW][FUSED_U * 32];
    __shared__ float rs[BS][32];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int ci = wi % BS, cg = wi / BS;                 // component / column group of this warp in the matrix phases
    const int gtid = blockIdx.x * NT + threadIdx.x, gthreads = gridDim.x * NT;
    const int gwarp = blockIdx.x * NW + wi, gwarps = gridDim.x * NW;
    const int n_phases = plan.n_phases;
    unsigned bar_gen = ld_acquire_gpu_u32(plan.bar + 1);
    unsigned long long gseq0 = 0;
    int gather_no = 0;
    if (plan.n_gathers > 0) gseq0 = *(volatile unsigned long long*)(plan.sym.base[plan.sym.rank] + 64);
    if (gtid == 0) plan.tstamp[0] = global_ns();
    // The matrices of these levels were evicted from L2 by the fine-level streams of the same PCG iteration and every
    // phase is a chain of dependent loads: pull them (and the transfer maps) back into L2 while phase 0 runs.
    for (int k = 0; k + 1 < plan.n_levels; ++k) {
        const FusedLevel& F = plan.L[k];
        if (F.A == nullptr || F.n_own > 8192) continue;        // top-only level, or too big to be worth pulling in
        const i64 n_slots = F.slice_off[F.n_slices];
        const char* a = (const char*)F.A;
        const i64 nb_a = n_slots * (i64)(BS * BS) * 4, nb_c = n_slots * 4, nb_d = (i64)F.n_own * BS * BS * 4;
        for (i64 o = (i64)gtid * 128; o < nb_a; o += (i64)gthreads * 128) prefetch_l2(a + o);
        for (i64 o = (i64)gtid * 128; o < nb_c; o += (i64)gthreads * 128) prefetch_l2((const char*)F.col + o);
        for (i64 o = (i64)gtid * 128; o < nb_d; o += (i64)gthreads * 128) prefetch_l2((const char*)F.dinv + o);
    }
    if (plan.coarse_m > 0) {
        const i64 nb_i = (i64)plan.coarse_m * plan.coarse_m * 4;
        for (i64 o = (i64)gtid * 128; o < nb_i; o += (i64)gthreads * 128) prefetch_l2((const char*)plan.coarse_inv + o);
    }
    for (int ph = 0; ph < n_phases; ++ph) {
        const FusedPhase& P = plan.P[ph];
        const FusedLevel& L = plan.L[P.lvl];
        if (P.op == FOP_RESTRICT) {
            // b(next) = P^T r: one warp per aggregate, lanes over its members, shuffle reduction (as k_restrict32)
            const FusedLevel& C = plan.L[P.lvl + 1];
            for (int I = gwarp; I < L.nc; I += gwarps) {
                double acc[6] = {0, 0, 0, 0, 0, 0};
                for (int m = L.mem_ptr[I] + lane; m < L.mem_ptr[I + 1]; m += 32) {
                    const int i = L.mem_idx[m];
                    double Pm[6][6];
                    int a_, b_;
                    build_P<D>(false, L.rvec + (i64)i * D, 0xffu, Pm, a_, b_);
#pragma unroll
                    for (int k = 0; k < BS; ++k) {
                        const double v = vload(&P.src[(i64)i * BS + k]);
#pragma unroll
                        for (int j = 0; j < BS; ++j) acc[j] += Pm[k][j] * v;
                    }
                }
#pragma unroll
                for (int j = 0; j < 6; ++j)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    if (lane == j && j < BS) C.b[(i64)I * BS + j] = (float)acc[j];
            }
        } else if (P.op == FOP_FIRST) {
            // x = d = c2 Dinv b
            for (int t = gtid; t < L.n_own * BS; t += gthreads) {
                const int rl = t / BS, i = t - rl * BS;
                const i64 row = (i64)L.row0 + rl;
                float z = 0.f;
#pragma unroll
                for (int j = 0; j < BS; ++j) z += __ldg(&L.dinv[(i64)rl * BS * BS + i * BS + j]) * vload(&L.b[row * BS + j]);
                const float dn = P.c2 * z;
                L.d[row * BS + i] = dn;
                P.dst[row * BS + i] = dn;
            }
        } else if (P.op == FOP_STEP || P.op == FOP_RESID) {
            for (int S = blockIdx.x; S < L.n_slices; S += gridDim.x) {
                const int rl = S * 32 + lane;              // row inside the matrix arrays
                const i64 r = (i64)L.row0 + rl;            // row of the level (vector index)
                const bool live = rl < L.n_own;
                const i64 base = L.slice_off[S];
                const int w = L.slice_w[S];
                // issue the row-local loads before the column walk: they overlap its two latencies
                float bv = 0.f, dv = 0.f, xo = 0.f, dinv_row[BS];
                if (cg == 0 && live) {
                    bv = vload(&L.b[r * BS + ci]);
                    if (P.op == FOP_STEP) {
                        dv = P.c1 != 0.f ? vload(&L.d[r * BS + ci]) : 0.f;
                        xo = vload(&P.src[r * BS + ci]);
#pragma unroll
                        for (int jj = 0; jj < BS; ++jj) dinv_row[jj] = __ldg(&L.dinv[(i64)rl * BS * BS + ci * BS + jj]);
                    }
                }
                if (gtid == 0) plan.tstamp[128 + ph * 4 + 0] = global_ns();
                part[cg][ci][lane] = fused_row_dot<BS>(base, w, L.col, L.A, P.src, ci, cg, lane, scol_all[wi]);
                if (gtid == 0) plan.tstamp[128 + ph * 4 + 1] = global_ns();
                __syncthreads();
                if (gtid == 0) plan.tstamp[128 + ph * 4 + 2] = global_ns();
                if (cg == 0) {
                    float dotv = part[0][ci][lane];
#pragma unroll
                    for (int q = 1; q < FUSED_CG; ++q) dotv += part[q][ci][lane];
                    const float res = live ? bv - dotv : 0.f;
                    if (P.op == FOP_RESID) {
                        if (live) L.r[r * BS + ci] = res;
                    } else rs[ci][lane] = res;
                }
                if (P.op == FOP_STEP) {
                    __syncthreads();
                    if (cg == 0 && live) {
                        float z = 0.f;
#pragma unroll
                        for (int jj = 0; jj < BS; ++jj) z += dinv_row[jj] * rs[jj][lane];
                        const float dn = P.c2 * z + P.c1 * dv;
                        L.d[r * BS + ci] = dn;
                        P.dst[r * BS + ci] = xo + dn;
                    }
                }
                __syncthreads();
            }
        } else if (P.op == FOP_DENSE) {
            const int m = plan.coarse_m;
            for (int row = gwarp; row < m; row += gwarps) {
                float s = 0.f;
                for (int j = lane; j < m; j += 32) s += __ldg(&plan.coarse_inv[(i64)row * m + j]) * vload(&L.b[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) L.x[row] = s;
            }
        } else if (P.op == FOP_PROLONG) {
            // dst(level) += P x(next)
            const FusedLevel& C = plan.L[P.lvl + 1];
            for (int i = gtid; i < L.n; i += gthreads) {
                const int I = L.agg[i];
                if (I < 0) continue;
                double Pm[6][6];
                int a_, b_;
                build_P<D>(false, L.rvec + (i64)i * D, 0xffu, Pm, a_, b_);
                double xc[BS];
#pragma unroll
                for (int j = 0; j < BS; ++j) xc[j] = (double)vload(&C.x[(i64)I * BS + j]);
#pragma unroll
                for (int k = 0; k < BS; ++k) {
                    double v = 0;
#pragma unroll
                    for (int j = 0; j < BS; ++j) v += Pm[k][j] * xc[j];
                    P.dst[(i64)i * BS + k] = vload(&P.dst[(i64)i * BS + k]) + (float)v;
                }
            }
        } else if (P.op == FOP_COPY) {
            const i64 o = (i64)L.row0 * BS;
            for (int t = gtid; t < L.n_own * BS; t += gthreads) P.dst[o + t] = vload(&P.src[o + t]);
        } else if (P.op == FOP_GATHER) {
            // in-place all-gather of a level-1 vector over peer memory: push the own segment into every peer's copy, then
            // (grid barrier: all pushes of this rank are done and fenced) publish the sequence number and wait for the peers'
            const int R = plan.sym.n_ranks, rank = plan.sym.rank;
            unsigned char* me = plan.sym.base[rank];
            const size_t off = (size_t)((unsigned char*)P.src - me);
            const long long seg = plan.seg, n4 = seg >> 2, total = n4 * (R - 1);
            const float4* src4 = (const float4*)(P.src + (size_t)rank * seg);
            for (long long t = gtid; t < total; t += gthreads) {
                const int q = (int)(t / n4);
                const long long k = t - q * n4;
                const int pr = q < rank ? q : q + 1;
                ((float4*)((float*)(plan.sym.base[pr] + off) + (size_t)rank * seg))[k] = src4[k];
            }
            __threadfence_system();
            fused_grid_barrier(plan.bar, bar_gen);
            ++gather_no;
            const unsigned long long seq = gseq0 + gather_no;
            if (blockIdx.x == 0 && threadIdx.x < R && threadIdx.x != rank)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"((unsigned long long*)plan.sym.base[threadIdx.x] + rank), "l"(seq) : "memory");
            if (threadIdx.x < R && threadIdx.x != rank && ld_acquire_gpu_u32(plan.bar + 2) == 0u) {
                const unsigned long long* flag = (const unsigned long long*)me + threadIdx.x;
                const long long t0 = clock64();
                while (true) {
                    unsigned long long v;
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
                    if (v >= seq) break;
                    if (clock64() - t0 > 8000000000LL) { atomicExch(plan.bar + 2, 2u); break; }
                }
                __threadfence_system();
            }
            __syncthreads();
        }
        if (gtid == 0) plan.tstamp[128 + ph * 4 + 3] = global_ns();
        if (ph == 2 && threadIdx.x == 0 && blockIdx.x < 512) { plan.tstamp[1024 + blockIdx.x] = global_ns(); unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm)); plan.tstamp[1536 + blockIdx.x] = sm; }
        if (ph + 1 < n_phases && P.op != FOP_GATHER) fused_grid_barrier(plan.bar, bar_gen);
        if (gtid == 0) plan.tstamp[ph + 1] = global_ns();
    }
    // the all-gather sequence number of the symmetric buffer advances once per launch, after every CTA has read it (each
    // has passed at least one grid barrier by now)
    if (plan.n_gathers > 0 && gtid == 0) *(volatile unsigned long long*)(plan.sym.base[plan.sym.rank] + 64) = gseq0 + plan.n_gathers;
}

struct FusedHost {
    FusedPlan* plan = nullptr;      // host copy: passed to the kernel by value (__grid_constant__)
    unsigned* bar = nullptr;
    unsigned long long* tstamp = nullptr;
    void* arena = nullptr;          // one allocation holding every matrix, map and vector of the fused levels
    size_t arena_bytes = 0;
    std::vector<int> ops, lvls;
    int grid = 0, top_level = 0;
    bool full = false;              // partitioned mesh: level `top_level` itself (own rows + all-gathers) runs inside the kernel
    float* top_cur = nullptr;       // the iterate buffer of the level above that the plan prolongs into
    int n_phases = 0;
};

static void amg_fused_free(Amg* amg) {
    FusedHost* f = (FusedHost*)amg->fused;
    if (!f) return;
    if (f->plan) delete f->plan;
    if (f->bar) cudaFree(f->bar);
    if (f->tstamp) cudaFree(f->tstamp);
    if (f->arena) cudaFree(f->arena);
    delete f;
    amg->fused = nullptr;
}

// Build the phase list for "restrict r32 of level `top`, V-cycle on levels top+1 .. coarsest, prolong into `cur`", or --
// `full`, partitioned mesh -- for the whole V-cycle from level `top` (= 1) down: b32(top) own segment in, x32(top) own
// segment out, the rank's own rows of the replicated level-1 operator and the all-gathers of its vectors inside the kernel.
static FusedHost* amg_fused_build(glims_ctx* c, Amg* amg, int top, float* cur, bool full) {
    const int nl = (int)amg->L.size();
    FusedHost* f = new FusedHost();
    f->top_level = top; f->top_cur = cur; f->full = full;
    FusedPlan* P = new FusedPlan();
    memset(P, 0, sizeof(FusedPlan));
    P->n_levels = nl - top;
    if (P->n_levels > FUSED_MAX_LEVELS) { delete P; return f; }
    // Everything the kernel reads or writes (except r32 / cur of the level above) lives in ONE allocation: the data of
    // these levels are ~20 MB, but spread over fifty separate allocations they cost fifty 2 MB pages, and after the
    // fine-level kernels of an iteration have streamed gigabytes through the TLBs every phase started with a storm of
    // page walks (measured: the slowest CTA of a matrix phase took 22 us for 5 us of work).
    std::vector<i64> h_so;      // slot counts
    size_t arena_bytes = 0;
    auto reserve = [&](size_t b) { size_t o = arena_bytes; arena_bytes += (b + 255) & ~(size_t)255; return o; };
    struct Off { size_t so, sw, col, A, dinv, x, y, b, r, d, agg, rvec, mptr, midx; i64 n_slots, slot0; int n_mem, s0, ns; };
    std::vector<i64> top_so;        // full: slice offsets of the top level (host copy)
    std::vector<Off> offs(P->n_levels);
    for (int k = 0; k < P->n_levels; ++k) {
        const Level& l = amg->L[top + k];
        Off& o = offs[k];
        o.n_slots = l.pat.n_slots; o.slot0 = 0; o.s0 = 0; o.ns = l.pat.n_slices;
        const i64 bb = (i64)l.bs * l.bs, nv = (i64)l.n * l.bs;
        const bool last = (k + 1 == P->n_levels);
        const bool smoothed = !last && (k > 0 || full);
        if (smoothed && !l.A32) { delete P; return f; }            // FP16 coarse levels: not fused
        if (k == 0 && full) {
            // the rank's own rows of the replicated operator only
            top_so.resize(l.pat.n_slices + 1);
            GL_CUDA(cudaMemcpy(top_so.data(), l.pat.slice_off, sizeof(i64) * top_so.size(), cudaMemcpyDeviceToHost));
            o.s0 = l.row0 / 32; o.ns = l.rows / 32;
            o.slot0 = top_so[o.s0]; o.n_slots = top_so[o.s0 + o.ns] - o.slot0;
        }
        if (smoothed) {
            o.so = reserve(sizeof(i64) * (o.ns + 1)); o.sw = reserve(sizeof(int) * std::max(o.ns, 1));
            o.col = reserve(sizeof(int) * std::max<i64>(o.n_slots, 1)); o.A = reserve(sizeof(float) * std::max<i64>(o.n_slots, 1) * bb);
            o.dinv = reserve(sizeof(float) * std::max<i64>((k == 0 ? l.rows : l.n), 1) * bb);
        }
        if (k > 0) { o.x = reserve(4 * nv); o.y = reserve(4 * nv); o.b = reserve(4 * nv); o.r = reserve(4 * nv); o.d = reserve(4 * nv); }
        else if (full) o.d = reserve(4 * nv);
        if (!last) {
            GL_CUDA(cudaMemcpy(&o.n_mem, l.mem_ptr + l.nc_local, sizeof(int), cudaMemcpyDeviceToHost));
            o.agg = reserve(sizeof(int) * l.n); o.rvec = reserve(sizeof(double) * (i64)l.n * amg->dim);
            o.mptr = reserve(sizeof(int) * (l.nc_local + 1)); o.midx = reserve(sizeof(int) * std::max(o.n_mem, 1));
        }
    }
    const size_t o_inv = reserve(sizeof(float) * (size_t)amg->coarse_m * amg->coarse_m);
    GL_CUDA(cudaMalloc(&f->arena, arena_bytes));
    GL_CUDA(cudaMemset(f->arena, 0, arena_bytes));
    unsigned char* ar = (unsigned char*)f->arena;
    auto put = [&](size_t off, const void* src, size_t bytes) {
        if (bytes) GL_CUDA(cudaMemcpy(ar + off, src, bytes, cudaMemcpyDeviceToDevice));
        return (void*)(ar + off);
    };
    for (int k = 0; k < P->n_levels; ++k) {
        const Level& l = amg->L[top + k];
        FusedLevel& F = P->L[k];
        const Off& o = offs[k];
        const i64 bb = (i64)l.bs * l.bs;
        const bool last = (k + 1 == P->n_levels);
        F.n = l.n; F.n_slices = o.ns; F.bs = l.bs;
        F.n_own = l.n; F.row0 = 0;
        F.nc = l.nc;
        const bool smoothed = !last && (k > 0 || full);
        if (smoothed) {
            if (k == 0) {
                // own rows: slice offsets rebased to the first own slot
                std::vector<i64> so(o.ns + 1);
                for (int i = 0; i <= o.ns; ++i) so[i] = top_so[o.s0 + i] - o.slot0;
                GL_CUDA(cudaMemcpy(ar + o.so, so.data(), sizeof(i64) * so.size(), cudaMemcpyHostToDevice));
                F.slice_off = (const i64*)(ar + o.so);
                F.n_own = l.rows; F.row0 = l.row0;
            } else F.slice_off = (const i64*)put(o.so, l.pat.slice_off, sizeof(i64) * (o.ns + 1));
            F.slice_w = (const int*)put(o.sw, l.pat.slice_w + o.s0, sizeof(int) * o.ns);
            F.col = (const int*)put(o.col, l.pat.col + o.slot0, sizeof(int) * o.n_slots);
            F.A = (const float*)put(o.A, l.A32 + o.slot0 * bb, sizeof(float) * o.n_slots * bb);
            F.dinv = (const float*)put(o.dinv, l.dinv32 + (i64)F.row0 * bb, sizeof(float) * (i64)F.n_own * bb);
        }
        if (k == 0) {
            // level above the fused block (or, full, level 1 itself): its vectors stay where the rest of the V-cycle and
            // the peers find them; only d32 is private
            F.r = l.r32; F.x = l.x32; F.y = l.y32; F.b = l.b32;
            F.d = full ? (float*)(ar + o.d) : l.d32;
        } else {
            F.x = (float*)(ar + o.x); F.y = (float*)(ar + o.y); F.b = (float*)(ar + o.b); F.r = (float*)(ar + o.r); F.d = (float*)(ar + o.d);
        }
        if (!last) {
            F.agg = (const int*)put(o.agg, l.agg, sizeof(int) * l.n);
            F.rvec = (const double*)put(o.rvec, l.rvec, sizeof(double) * (i64)l.n * amg->dim);
            F.mem_ptr = (const int*)put(o.mptr, l.mem_ptr, sizeof(int) * (l.nc_local + 1));
            F.mem_idx = (const int*)put(o.midx, l.mem_idx, sizeof(int) * o.n_mem);
        }
    }
    P->coarse_inv = (const float*)put(o_inv, amg->coarse_inv32, sizeof(float) * (size_t)amg->coarse_m * amg->coarse_m);
    f->arena_bytes = arena_bytes;
    P->coarse_m = amg->coarse_m;
    int np = 0;
    bool ok = true;
    auto push = [&](int op, int lvl, float* src, float* dst, double c1, double c2) {
        if (np >= FUSED_MAX_PHASES) { ok = false; return; }
        P->P[np++] = FusedPhase{op, lvl, src, dst, (float)c1, (float)c2};
    };
    const int deg = amg->coarse_degree;
    int n_gathers = 0;
    std::function<void(int)> emit = [&](int k) {
        FusedLevel& F = P->L[k];
        if (k == P->n_levels - 1) { push(FOP_DENSE, k, nullptr, nullptr, 0, 0); return; }
        const Level& l = amg->L[top + k];
        const bool dr = (k == 0 && full);              // own rows + all-gathers (same schedule as vcycle32)
        auto gather = [&](float* v) { if (dr) { push(FOP_GATHER, k, v, nullptr, 0, 0); ++n_gathers; } };
        float *cu = F.y, *ot = F.x;
        double c1, c2;
        {
            ChebCoef cc(l.lmax, amg->cheb_ratio);
            cc.step(0, c1, c2);
            push(FOP_FIRST, k, nullptr, cu, 0, c2);
            for (int j = 1; j < deg; ++j) { cc.step(j, c1, c2); gather(cu); push(FOP_STEP, k, cu, ot, c1, c2); std::swap(cu, ot); }
        }
        gather(cu);
        push(FOP_RESID, k, cu, nullptr, 0, 0);
        gather(F.r);
        push(FOP_RESTRICT, k, F.r, nullptr, 0, 0);
        emit(k + 1);
        push(FOP_PROLONG, k, nullptr, cu, 0, 0);
        {
            ChebCoef cc(l.lmax, amg->cheb_ratio);
            for (int j = 0; j < deg; ++j) { cc.step(j, c1, c2); if (j > 0) gather(cu); push(FOP_STEP, k, cu, ot, c1, c2); std::swap(cu, ot); }
        }
        if (cu != F.x) push(FOP_COPY, k, cu, F.x, 0, 0);
    };
    if (full) {
        emit(0);
        const Level& l1 = amg->L[top];
        P->sym = *sym_view(l1.sym);
        P->seg = (long long)l1.rows * l1.bs;
        P->n_gathers = n_gathers;
    } else {
        push(FOP_RESTRICT, 0, P->L[0].r, nullptr, 0, 0);
        emit(1);
        push(FOP_PROLONG, 0, nullptr, cur, 0, 0);
    }
    if (!ok) { delete P; return f; }
    P->n_phases = np;
    f->n_phases = np;
    GL_CUDA(cudaMalloc(&f->bar, 4 * sizeof(unsigned)));
    GL_CUDA(cudaMemset(f->bar, 0, 4 * sizeof(unsigned)));
    P->bar = f->bar;
    GL_CUDA(cudaMalloc(&f->tstamp, 2048 * sizeof(unsigned long long)));
    GL_CUDA(cudaMemset(f->tstamp, 0, 2048 * sizeof(unsigned long long)));
    P->tstamp = f->tstamp;
    for (int i = 0; i < np; ++i) { f->ops.push_back(P->P[i].op); f->lvls.push_back(P->P[i].lvl); }
    // grid: enough CTAs for the widest phase, never more than can be co-resident (the barrier spins)
    const int D = amg->dim, NT = 32 * (D == 2 ? 3 : 6) * FUSED_CG;
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (D == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_amg_fused<2>, NT, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_amg_fused<3>, NT, 0);
    int want = std::max(std::max(P->L[1].n_slices, full ? P->L[0].n_slices : 0), (P->L[0].n + NT * 2 - 1) / (NT * 2));
    want = std::max(want, 1);
    const char* eg = std::getenv("GLIMS_AMG_FUSED_GRID");
    if (eg) want = std::max(1, atoi(eg));
    // stay at one CTA per SM at most unless asked otherwise: the barrier cost grows with the CTA count
    f->grid = std::min(want, std::max(1, std::min(per_sm, 2) * sms));
    f->plan = P;
    return f;
}

static bool amg_fused_failed(Amg* amg) {
    FusedHost* f = amg ? (FusedHost*)amg->fused : nullptr;
    if (!f || !f->bar) return false;
    unsigned e = 0;
    cudaMemcpy(&e, f->bar + 2, sizeof(unsigned), cudaMemcpyDeviceToHost);
    return e != 0;
}

// per-phase durations of the last launch (GLIMS_VERBOSE diagnostics)
static void amg_fused_print_phases(Amg* amg) {
    FusedHost* f = amg ? (FusedHost*)amg->fused : nullptr;
    if (!f || !f->tstamp) return;
    std::vector<unsigned long long> t(f->n_phases + 1);
    cudaMemcpy(t.data(), f->tstamp, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost);
    static const char* names[] = {"restrict", "first", "step", "resid", "dense", "prolong", "copy", "gather"};
    fprintf(stderr, "glims amg fused kernel: grid %d, %d phases, %.1f us:", f->grid, f->n_phases, (t.back() - t[0]) * 1e-3);
    for (int i = 0; i < f->n_phases; ++i) fprintf(stderr, " %s@%d %.1f", names[f->ops[i]], f->lvls[i], (t[i + 1] - t[i]) * 1e-3);
    fprintf(stderr, "\n");
    std::vector<unsigned long long> t2(4 * f->n_phases);
    cudaMemcpy(t2.data(), f->tstamp + 128, sizeof(unsigned long long) * t2.size(), cudaMemcpyDeviceToHost);
    fprintf(stderr, "glims amg fused kernel, CTA 0 inside the matrix phases [start->dot, dot->sync, sync->end, end->barrier done]:");
    for (int i = 0; i < f->n_phases; ++i)
        if (f->ops[i] == FOP_STEP || f->ops[i] == FOP_RESID)
            fprintf(stderr, " %s@%d [%.1f %.1f %.1f %.1f | phase start->dot start %.1f]", names[f->ops[i]], f->lvls[i], (t2[4 * i + 1] - t2[4 * i]) * 1e-3,
                    (t2[4 * i + 2] - t2[4 * i + 1]) * 1e-3, (t2[4 * i + 3] - t2[4 * i + 2]) * 1e-3, (t[i + 1] - t2[4 * i + 3]) * 1e-3, (t2[4 * i] - t[i]) * 1e-3);
    fprintf(stderr, "\n");
    {
        const int g = std::min(f->grid, 512);
        std::vector<unsigned long long> tc(g), sm(g);
        cudaMemcpy(tc.data(), f->tstamp + 1024, sizeof(unsigned long long) * g, cudaMemcpyDeviceToHost);
        cudaMemcpy(sm.data(), f->tstamp + 1536, sizeof(unsigned long long) * g, cudaMemcpyDeviceToHost);
        fprintf(stderr, "glims amg fused kernel, phase 2: work-done time of every CTA after the phase start (us) [cta:sm]:");
        for (int i = 0; i < g; ++i) fprintf(stderr, " %d:%llu=%.1f", i, sm[i], ((double)tc[i] - (double)t[2]) * 1e-3);
        fprintf(stderr, "\n");
    }
}

// true: levels li+1.. ran fused (r32 of level li restricted, result prolonged into cur)
static bool amg_fused_run(glims_ctx* c, Amg* amg, int li, float* cur) {
    if (li != 1 || (int)amg->L.size() < 3) return false;
    static const bool enabled = [] { const char* e = std::getenv("GLIMS_AMG_FUSED"); return !(e && atoi(e) == 0); }();
    if (!enabled) return false;
    FusedHost* f = (FusedHost*)amg->fused;
    if (!f) { f = amg_fused_build(c, amg, li, cur, false); amg->fused = f; }
    if (!f->plan || f->full || f->top_cur != cur || f->top_level != li) return false;
    if (amg->dim == 2) k_amg_fused<2><<<f->grid, 96 * FUSED_CG, 0, c->stream>>>(*f->plan);
    else k_amg_fused<3><<<f->grid, 192 * FUSED_CG, 0, c->stream>>>(*f->plan);
    c->launches++;
    return true;
}

// Partitioned mesh with a small per-rank share of level 1: the whole V-cycle from level 1 down as one launch (b32 own
// segment in, x32 own segment out).  true: handled.
static bool amg_fused_run_full(glims_ctx* c, Amg* amg, int li, float* b, float* x) {
    if (li != 1 || !amg->dist || (int)amg->L.size() < 3) return false;
    const char* e = std::getenv("GLIMS_AMG_FUSED");
    if (e && (atoi(e) == 0 || atoi(e) == 1)) return false;       // 0: nothing fused, 1: levels >= 2 only, default: level 1 too
    Level& l = amg->L[li];
    // only where the rank's share of level 1 is small enough to be latency bound (a persistent grid of ~150 CTAs cannot
    // stream a bandwidth-bound level); GLIMS_AMG_FUSED_ROWS overrides the limit
    static const int max_rows = [] { const char* r = std::getenv("GLIMS_AMG_FUSED_ROWS"); return r ? atoi(r) : 32768; }();
    if (!l.dist_rows || !l.sym || !sym_is_p2p(l.sym) || !c->halo.p2p_enabled || !l.A32 || l.rows > max_rows) return false;
    if (b != l.b32 || x != l.x32) return false;
    FusedHost* f = (FusedHost*)amg->fused;
    if (!f) { f = amg_fused_build(c, amg, li, nullptr, true); amg->fused = f; }
    if (!f->plan || !f->full) return false;
    if (amg->dim == 2) k_amg_fused<2><<<f->grid, 96 * FUSED_CG, 0, c->stream>>>(*f->plan);
    else k_amg_fused<3><<<f->grid, 192 * FUSED_CG, 0, c->stream>>>(*f->plan);
    c->launches++;
    return true;
}
