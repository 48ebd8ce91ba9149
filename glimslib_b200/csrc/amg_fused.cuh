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

enum { FOP_RESTRICT = 0, FOP_FIRST, FOP_STEP, FOP_RESID, FOP_DENSE, FOP_PROLONG, FOP_COPY };
constexpr int FUSED_MAX_LEVELS = 12, FUSED_MAX_PHASES = 96;

struct FusedLevel {
    int n, n_slices, bs;
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
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all CTAs of the (co-resident) grid; sense-reversing: leaves the arrival counter at zero.  The wait is bounded (about two
// seconds of SM clock): a grid that is not co-resident after all raises bar[2] (amg_fused_check) instead of hanging the GPU.
__device__ __forceinline__ void fused_grid_barrier(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = ld_acquire_gpu_u32(bar + 1);
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            bar[0] = 0u;
            __threadfence();
            st_release_gpu_u32(bar + 1, gen + 1u);
        } else if (ld_acquire_gpu_u32(bar + 2) == 0u) {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(bar + 1) == gen) {
                if (clock64() - t0 > 4000000000LL) { atomicExch(bar + 2, 1u); break; }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// component i of (A x) for row `lane` of a slice; x through L2.  Eight block columns in flight: a row of ~24 blocks is
// three rounds of dependent loads (column index -> x), which is what a phase costs on these latency-bound levels.
template <int BS>
__device__ __forceinline__ float fused_row_dot(const i64 base, const int w, const int* __restrict__ col,
                                               const float* __restrict__ A, const float* x, const int i, const int lane) {
    constexpr int U = 8;
    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.f;
    int j = 0;
    for (; j + U - 1 < w; j += U) {
        int cc[U];
        float av[U][BS];
#pragma unroll
        for (int u = 0; u < U; ++u) cc[u] = __ldg(&col[base + (i64)(j + u) * 32 + lane]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float* Au = A + (base + (i64)(j + u) * 32) * (BS * BS) + (i * BS) * 32 + lane;
#pragma unroll
            for (int b = 0; b < BS; ++b) av[u][b] = __ldg(&Au[b * 32]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int b = 0; b < BS; ++b) acc[u] += av[u][b] * __ldcg(&x[(i64)cc[u] * BS + b]);
    }
    for (; j < w; ++j) {
        const int c0 = __ldg(&col[base + (i64)j * 32 + lane]);
        const float* A0 = A + (base + (i64)j * 32) * (BS * BS) + (i * BS) * 32 + lane;
#pragma unroll
        for (int b = 0; b < BS; ++b) acc[j & (U - 1)] += __ldg(&A0[b * 32]) * __ldcg(&x[(i64)c0 * BS + b]);
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) s += acc[u];
    return s;
}

__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int D>
__global__ void __launch_bounds__(32 * ((D == 2) ? 3 : 6))
k_amg_fused(const FusedPlan* __restrict__ plan) {
    constexpr int BS = (D == 2) ? 3 : 6, NT = 32 * BS;
    __shared__ float rs[BS][32];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int gtid = blockIdx.x * NT + threadIdx.x, gthreads = gridDim.x * NT;
    const int gwarp = blockIdx.x * BS + wi, gwarps = gridDim.x * BS;
    const int n_phases = plan->n_phases;
    if (gtid == 0) plan->tstamp[0] = global_ns();
    // The matrices of these levels were evicted from L2 by the fine-level streams of the same PCG iteration and every
    // phase is a chain of dependent loads: pull them (and the transfer maps) back into L2 while phase 0 runs.
    for (int k = 1; k + 1 < plan->n_levels; ++k) {
        const FusedLevel& F = plan->L[k];
        const i64 n_slots = F.slice_off[F.n_slices];
        const char* a = (const char*)F.A;
        const i64 nb_a = n_slots * (i64)(BS * BS) * 4, nb_c = n_slots * 4, nb_d = (i64)F.n * BS * BS * 4;
        for (i64 o = (i64)gtid * 128; o < nb_a; o += (i64)gthreads * 128) prefetch_l2(a + o);
        for (i64 o = (i64)gtid * 128; o < nb_c; o += (i64)gthreads * 128) prefetch_l2((const char*)F.col + o);
        for (i64 o = (i64)gtid * 128; o < nb_d; o += (i64)gthreads * 128) prefetch_l2((const char*)F.dinv + o);
    }
    for (int ph = 0; ph < n_phases; ++ph) {
        const FusedPhase P = plan->P[ph];
        const FusedLevel& L = plan->L[P.lvl];
        if (P.op == FOP_RESTRICT) {
            // b(next) = P^T r: one warp per aggregate, lanes over its members, shuffle reduction (as k_restrict32)
            const FusedLevel& C = plan->L[P.lvl + 1];
            for (int I = gwarp; I < L.nc; I += gwarps) {
                double acc[6] = {0, 0, 0, 0, 0, 0};
                for (int m = L.mem_ptr[I] + lane; m < L.mem_ptr[I + 1]; m += 32) {
                    const int i = L.mem_idx[m];
                    double Pm[6][6];
                    int a_, b_;
                    build_P<D>(false, L.rvec + (i64)i * D, 0xffu, Pm, a_, b_);
                    for (int k = 0; k < BS; ++k) {
                        const double v = __ldcg(&P.src[(i64)i * BS + k]);
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
            for (int t = gtid; t < L.n * BS; t += gthreads) {
                const int row = t / BS, i = t - row * BS;
                float z = 0.f;
#pragma unroll
                for (int j = 0; j < BS; ++j) z += __ldg(&L.dinv[(i64)row * BS * BS + i * BS + j]) * __ldcg(&L.b[(i64)row * BS + j]);
                const float dn = P.c2 * z;
                L.d[t] = dn;
                P.dst[t] = dn;
            }
        } else if (P.op == FOP_STEP || P.op == FOP_RESID) {
            for (int S = blockIdx.x; S < L.n_slices; S += gridDim.x) {
                const int r = S * 32 + lane;
                const i64 base = L.slice_off[S];
                const int w = L.slice_w[S];
                const float dotv = fused_row_dot<BS>(base, w, L.col, L.A, P.src, wi, lane);
                const float res = r < L.n ? __ldcg(&L.b[(i64)r * BS + wi]) - dotv : 0.f;
                if (P.op == FOP_RESID) {
                    if (r < L.n) L.r[(i64)r * BS + wi] = res;
                } else {
                    rs[wi][lane] = res;
                    __syncthreads();
                    if (r < L.n) {
                        float z = 0.f;
#pragma unroll
                        for (int jj = 0; jj < BS; ++jj) z += __ldg(&L.dinv[(i64)r * BS * BS + wi * BS + jj]) * rs[jj][lane];
                        const float dn = P.c2 * z + (P.c1 != 0.f ? P.c1 * __ldcg(&L.d[(i64)r * BS + wi]) : 0.f);
                        L.d[(i64)r * BS + wi] = dn;
                        P.dst[(i64)r * BS + wi] = __ldcg(&P.src[(i64)r * BS + wi]) + dn;
                    }
                    __syncthreads();
                }
            }
        } else if (P.op == FOP_DENSE) {
            const int m = plan->coarse_m;
            for (int row = gwarp; row < m; row += gwarps) {
                float s = 0.f;
                for (int j = lane; j < m; j += 32) s += __ldg(&plan->coarse_inv[(i64)row * m + j]) * __ldcg(&L.b[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) L.x[row] = s;
            }
        } else if (P.op == FOP_PROLONG) {
            // dst(level) += P x(next)
            const FusedLevel& C = plan->L[P.lvl + 1];
            for (int i = gtid; i < L.n; i += gthreads) {
                const int I = L.agg[i];
                if (I < 0) continue;
                double Pm[6][6];
                int a_, b_;
                build_P<D>(false, L.rvec + (i64)i * D, 0xffu, Pm, a_, b_);
                for (int k = 0; k < BS; ++k) {
                    double v = 0;
                    for (int j = 0; j < BS; ++j) v += Pm[k][j] * (double)__ldcg(&C.x[(i64)I * BS + j]);
                    P.dst[(i64)i * BS + k] = __ldcg(&P.dst[(i64)i * BS + k]) + (float)v;
                }
            }
        } else if (P.op == FOP_COPY) {
            for (int t = gtid; t < L.n * BS; t += gthreads) P.dst[t] = __ldcg(&P.src[t]);
        }
        if (ph + 1 < n_phases) fused_grid_barrier(plan->bar);
        if (gtid == 0) plan->tstamp[ph + 1] = global_ns();
    }
}

struct FusedHost {
    FusedPlan* dev = nullptr;
    unsigned* bar = nullptr;
    unsigned long long* tstamp = nullptr;
    std::vector<int> ops, lvls;
    int grid = 0, top_level = 0;
    float* top_cur = nullptr;       // the iterate buffer of the level above that the plan prolongs into
    int n_phases = 0;
};

static void amg_fused_free(Amg* amg) {
    FusedHost* f = (FusedHost*)amg->fused;
    if (!f) return;
    if (f->dev) cudaFree(f->dev);
    if (f->bar) cudaFree(f->bar);
    if (f->tstamp) cudaFree(f->tstamp);
    delete f;
    amg->fused = nullptr;
}

// Build the phase list for "restrict r32 of level `top`, V-cycle on levels top+1 .. coarsest, prolong into `cur`".
static FusedHost* amg_fused_build(glims_ctx* c, Amg* amg, int top, float* cur) {
    const int nl = (int)amg->L.size();
    FusedHost* f = new FusedHost();
    f->top_level = top; f->top_cur = cur;
    FusedPlan* P = new FusedPlan();
    memset(P, 0, sizeof(FusedPlan));
    P->n_levels = nl - top;
    if (P->n_levels > FUSED_MAX_LEVELS) { delete P; return f; }
    for (int k = 0; k < P->n_levels; ++k) {
        const Level& l = amg->L[top + k];
        FusedLevel& F = P->L[k];
        F.n = l.n; F.n_slices = l.pat.n_slices; F.bs = l.bs;
        F.slice_off = l.pat.slice_off; F.slice_w = l.pat.slice_w; F.col = l.pat.col;
        F.A = l.A32; F.dinv = l.dinv32;
        F.x = l.x32; F.y = l.y32; F.b = l.b32; F.r = l.r32; F.d = l.d32;
        F.nc = l.nc; F.mem_ptr = l.mem_ptr; F.mem_idx = l.mem_idx; F.agg = l.agg; F.rvec = l.rvec;
        if (k > 0 && k + 1 < P->n_levels && !l.A32) { delete P; return f; }      // FP16 coarse levels: not fused
    }
    P->coarse_m = amg->coarse_m; P->coarse_inv = amg->coarse_inv32;
    int np = 0;
    bool ok = true;
    auto push = [&](int op, int lvl, float* src, float* dst, double c1, double c2) {
        if (np >= FUSED_MAX_PHASES) { ok = false; return; }
        P->P[np++] = FusedPhase{op, lvl, src, dst, (float)c1, (float)c2};
    };
    const int deg = amg->coarse_degree;
    std::function<void(int)> emit = [&](int k) {
        FusedLevel& F = P->L[k];
        if (k == P->n_levels - 1) { push(FOP_DENSE, k, nullptr, nullptr, 0, 0); return; }
        const Level& l = amg->L[top + k];
        float *cu = F.y, *ot = F.x;
        double c1, c2;
        {
            ChebCoef cc(l.lmax, amg->cheb_ratio);
            cc.step(0, c1, c2);
            push(FOP_FIRST, k, nullptr, cu, 0, c2);
            for (int j = 1; j < deg; ++j) { cc.step(j, c1, c2); push(FOP_STEP, k, cu, ot, c1, c2); std::swap(cu, ot); }
        }
        push(FOP_RESID, k, cu, nullptr, 0, 0);
        push(FOP_RESTRICT, k, F.r, nullptr, 0, 0);
        emit(k + 1);
        push(FOP_PROLONG, k, nullptr, cu, 0, 0);
        {
            ChebCoef cc(l.lmax, amg->cheb_ratio);
            for (int j = 0; j < deg; ++j) { cc.step(j, c1, c2); push(FOP_STEP, k, cu, ot, c1, c2); std::swap(cu, ot); }
        }
        if (cu != F.x) push(FOP_COPY, k, cu, F.x, 0, 0);
    };
    push(FOP_RESTRICT, 0, P->L[0].r, nullptr, 0, 0);
    emit(1);
    push(FOP_PROLONG, 0, nullptr, cur, 0, 0);
    if (!ok) { delete P; return f; }
    P->n_phases = np;
    f->n_phases = np;
    GL_CUDA(cudaMalloc(&f->bar, 4 * sizeof(unsigned)));
    GL_CUDA(cudaMemset(f->bar, 0, 4 * sizeof(unsigned)));
    P->bar = f->bar;
    GL_CUDA(cudaMalloc(&f->tstamp, (FUSED_MAX_PHASES + 1) * sizeof(unsigned long long)));
    GL_CUDA(cudaMemset(f->tstamp, 0, (FUSED_MAX_PHASES + 1) * sizeof(unsigned long long)));
    P->tstamp = f->tstamp;
    for (int i = 0; i < np; ++i) { f->ops.push_back(P->P[i].op); f->lvls.push_back(P->P[i].lvl); }
    GL_CUDA(cudaMalloc(&f->dev, sizeof(FusedPlan)));
    GL_CUDA(cudaMemcpy(f->dev, P, sizeof(FusedPlan), cudaMemcpyHostToDevice));
    // grid: enough CTAs for the widest phase, never more than can be co-resident (the barrier spins)
    const int D = amg->dim, NT = 32 * (D == 2 ? 3 : 6);
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (D == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_amg_fused<2>, NT, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_amg_fused<3>, NT, 0);
    int want = std::max(P->L[1].n_slices, (P->L[0].n + NT * 4 - 1) / (NT * 4));
    want = std::max(want, 1);
    const char* eg = std::getenv("GLIMS_AMG_FUSED_GRID");
    if (eg) want = std::max(1, atoi(eg));
    // stay at one CTA per SM at most unless asked otherwise: the barrier cost grows with the CTA count
    f->grid = std::min(want, std::max(1, std::min(per_sm, 2) * sms));
    delete P;
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
    static const char* names[] = {"restrict", "first", "step", "resid", "dense", "prolong", "copy"};
    fprintf(stderr, "glims amg fused kernel: grid %d, %d phases, %.1f us:", f->grid, f->n_phases, (t.back() - t[0]) * 1e-3);
    for (int i = 0; i < f->n_phases; ++i) fprintf(stderr, " %s@%d %.1f", names[f->ops[i]], f->lvls[i], (t[i + 1] - t[i]) * 1e-3);
    fprintf(stderr, "\n");
}

// true: levels li+1.. ran fused (r32 of level li restricted, result prolonged into cur)
static bool amg_fused_run(glims_ctx* c, Amg* amg, int li, float* cur) {
    if (li != 1 || (int)amg->L.size() < 3) return false;
    static const bool enabled = [] { const char* e = std::getenv("GLIMS_AMG_FUSED"); return !(e && atoi(e) == 0); }();
    if (!enabled) return false;
    FusedHost* f = (FusedHost*)amg->fused;
    if (!f) { f = amg_fused_build(c, amg, li, cur); amg->fused = f; }
    if (!f->dev || f->top_cur != cur || f->top_level != li) return false;
    if (amg->dim == 2) k_amg_fused<2><<<f->grid, 96, 0, c->stream>>>(f->dev);
    else k_amg_fused<3><<<f->grid, 192, 0, c->stream>>>(f->dev);
    c->launches++;
    return true;
}
