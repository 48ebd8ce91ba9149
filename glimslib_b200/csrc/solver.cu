// Host drivers: context life cycle, C ABI, Krylov solvers with device-resident scalars,
// Newton loop (SNES newtonls/basic semantics) and the backward-Euler step.
#include "common.h"
#include <cmath>
#include <cstring>
#include <algorithm>
#include <map>
#include <tuple>
#include <cstdlib>
#include <chrono>

void launch_split_norms(glims_ctx* c, const double* F, int s0);
void launch_pcg_shift(glims_ctx* c, double* ring);
void launch_pcg_cond(glims_ctx* c, unsigned long long handle, const double* ring);
void launch_extrapolate_c(glims_ctx* c, double* x, const double* xold);
void launch_cell_fields(glims_ctx* c, double* out, double* vol);
void launch_cell_to_vertex(glims_ctx* c, int nf, const double* q, const double* vol, double* num, double* den);

namespace {

struct GraphKey {
    int which, pc; void* amg; void* x; void* r; int bs;
    bool operator<(const GraphKey& o) const {
        return std::tie(which, pc, amg, x, r, bs) < std::tie(o.which, o.pc, o.amg, o.x, o.r, o.bs);
    }
};
struct PcgGraph { cudaGraphExec_t exec = nullptr; i64 launches = 0; bool failed = false; bool conditional = false; int its = 1; };

// Everything the host drivers keep per context (no process-global state: contexts may live on different threads)
struct SolverState {
    std::map<std::string, std::pair<double*, i64>> pool;     // named device work vectors
    std::vector<std::vector<double>> rec_hist;               // a[j][k]: coefficient of U_k in the j-th most recent solution
    std::map<GraphKey, PcgGraph> graphs;                     // captured PCG iterations
};
SolverState& state(glims_ctx* c) {
    if (!c->solver_state) c->solver_state = new SolverState();
    return *static_cast<SolverState*>(c->solver_state);
}

double* ws(glims_ctx* c, const char* name, i64 n) {
    auto& e = state(c).pool[name];
    if (e.second < n) {
        if (e.first) cudaFree(e.first);
        GL_CUDA(cudaMalloc(&e.first, sizeof(double) * (n > 0 ? n : 1)));
        GL_CUDA(cudaMemsetAsync(e.first, 0, sizeof(double) * (n > 0 ? n : 1), c->stream));
        e.second = n;
    }
    return e.first;
}
void free_graphs(glims_ctx* c) {
    for (auto& kv : state(c).graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    state(c).graphs.clear();
}
void free_pool(glims_ctx* c) {
    if (!c->solver_state) return;
    for (auto& kv : state(c).pool) cudaFree(kv.second.first);
    free_graphs(c);
    delete static_cast<SolverState*>(c->solver_state);
    c->solver_state = nullptr;
}

struct EvTimer {           // accumulates device time of bracketed segments on the stream
    cudaStream_t st;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> segs[3];   // 0 total, 1 assembly, 2 krylov
    explicit EvTimer(cudaStream_t s) : st(s) {}
    void begin(int k) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
        segs[k].push_back({a, b});
    }
    void end(int k) { cudaEventRecord(segs[k].back().second, st); }
    float total(int k) {
        float t = 0;
        for (auto& p : segs[k]) { float ms = 0; cudaEventSynchronize(p.second); cudaEventElapsedTime(&ms, p.first, p.second); t += ms; }
        return t;
    }
    ~EvTimer() { for (auto& s : segs) for (auto& p : s) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); } }
};

template <class T> T* dev_upload(const T* h, i64 n, cudaStream_t st) {
    T* d = nullptr;
    GL_CUDA(cudaMalloc(&d, sizeof(T) * (n > 0 ? n : 1)));
    if (n > 0) GL_CUDA(cudaMemcpyAsync(d, h, sizeof(T) * n, cudaMemcpyHostToDevice, st));
    return d;
}

i64 nrows(glims_ctx* c) { return c->pat.n_rows; }

// ------------------------------------------------------------------------------------------------
// preconditioner application z = M^-1 r, rz -> slot
void apply_pc(glims_ctx* c, int which, int pc, const double* r, double* z, int s_rz) {
    if (which == 2) { launch_block_jacobi(c, c->dinv_cc, 1, r, z, nrows(c), s_rz); return; }
    if (which == 3) { launch_block_jacobi(c, c->dinv_mass, 1, r, z, nrows(c), s_rz); return; }
    if (which == 1) {
        if ((pc == GLIMS_PC_AMG || pc == GLIMS_PC_AMG_FP64) && c->amg) {
            amg_vcycle(c, r, z, pc == GLIMS_PC_AMG);
            launch_dot(c, r, z, nrows(c) * c->dim, s_rz);
        } else launch_block_jacobi(c, c->dinv_uu, c->dim, r, z, nrows(c), s_rz);
        return;
    }
    launch_block_jacobi(c, c->dinv_mono, c->nb, r, z, nrows(c), s_rz);
}

// One PCG iteration as an executable graph.  Preferred form: [k_pcg_cond] -> IF(body = captured iteration), so that a
// launch after convergence is a no-op; if the conditional node cannot be built (driver, or a library call that refuses
// to be captured into a body graph) the iteration is captured as a plain graph, as before.
template <typename F>
bool build_pcg_graph(glims_ctx* c, PcgGraph* G, double* ring, F&& iteration, int unroll) {
    const bool want_cond = std::getenv("GLIMS_NO_COND_GRAPH") == nullptr;
    if (want_cond) {
        cudaGraph_t g = nullptr, tmp = nullptr;
        bool ok = cudaGraphCreate(&g, 0) == cudaSuccess;
        cudaGraphConditionalHandle h = 0;
        ok = ok && cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
        if (ok) {
            ok = cudaStreamBeginCaptureToGraph(c->stream, g, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                launch_pcg_cond(c, (unsigned long long)h, ring);
                ok = cudaStreamEndCapture(c->stream, &tmp) == cudaSuccess;
            }
        }
        cudaGraphNode_t knode = nullptr, cnode = nullptr;
        size_t nn = 1;
        ok = ok && cudaGraphGetNodes(g, &knode, &nn) == cudaSuccess && nn == 1;
        cudaGraph_t body = nullptr;
        if (ok) {
            cudaGraphNodeParams cp = {};
            cp.type = cudaGraphNodeTypeConditional;
            cp.conditional.handle = h;
            cp.conditional.type = cudaGraphCondTypeIf;
            cp.conditional.size = 1;
            ok = cudaGraphAddNode(&cnode, g, &knode, 1, &cp) == cudaSuccess;
            if (ok) body = cp.conditional.phGraph_out[0];
        }
        if (ok) {
            ok = cudaStreamBeginCaptureToGraph(c->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                try { for (int u = 0; u < unroll; ++u) iteration(); } catch (const GlError&) { ok = false; }
                if (cudaStreamEndCapture(c->stream, &tmp) != cudaSuccess) ok = false;
            }
        }
        if (ok && cudaGraphInstantiate(&G->exec, g, 0) != cudaSuccess) { ok = false; G->exec = nullptr; }
        if (g) cudaGraphDestroy(g);
        if (ok) { G->conditional = true; G->its = unroll; c->launches += 1; return true; }
        cudaGetLastError();
        // make sure the stream is not left in capture mode
        cudaStreamCaptureStatus st;
        if (cudaStreamIsCapturing(c->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) cudaStreamEndCapture(c->stream, &tmp);
        cudaGetLastError();
    }
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
        try { iteration(); } catch (const GlError&) { ok = false; }
        if (cudaStreamEndCapture(c->stream, &graph) != cudaSuccess || !graph) ok = false;
    }
    if (ok && cudaGraphInstantiate(&G->exec, graph, 0) != cudaSuccess) { ok = false; G->exec = nullptr; }
    if (graph) cudaGraphDestroy(graph);
    G->conditional = false;
    G->its = 1;
    return ok;
}

// PCG on block `which` (1: K_uu, 2: K_cc), zero initial guess. Returns iterations, or -1 if not converged.
// solutions whose span survives a compression of the projection basis (default 10; GLIMS_REC_KEEP)
inline int rec_keep() { const char* e = std::getenv("GLIMS_REC_KEEP"); int v = e ? atoi(e) : 10; return v < 2 ? 2 : (v > 28 ? 28 : v); }

constexpr int REC_M = 30;    // directions kept for the successive-RHS projection (k_multi_dot handles <= 32)

int pcg(glims_ctx* c, int which, int pc, const double* b, double* x, double tol_rel, double scale, double tol_abs,
        int maxit, double* res_out, bool recycle = false) {
    const int bs = which == 1 ? c->dim : 1;
    const i64 n = nrows(c) * bs, nl = c->n_v * bs;
    const char* tag = which == 1 ? "u" : which == 3 ? "m" : "c";
    char nm[32];
    auto W = [&](const char* base) { snprintf(nm, sizeof nm, "cg_%s_%s", base, tag); return ws(c, nm, nl); };
    double *r = W("r"), *p = W("p"), *Ap = W("Ap"), *z = W("z");
    launch_copy(c, b, r, n);
    launch_zero(c, x, n);
    // Successive right-hand sides, same SPD matrix (K_uu is constant over Newton iterations and time steps):
    // Galerkin-project b onto the A-orthonormal corrections U of earlier solves (Fischer 1998), then let PCG
    // solve only for the remainder.  x_bar = U (U^T b), r0 = b - (A U)(U^T b).
    double *U = nullptr, *AU = nullptr, *r0 = nullptr, *coef = nullptr;
    recycle = recycle && which == 1;
    if (recycle) {
        U = ws(c, "rec_U", nl * REC_M); AU = ws(c, "rec_AU", nl * REC_M);
        r0 = ws(c, "rec_r0", nl); coef = ws(c, "rec_coef", 64);
        if (c->rec_n > 0) {
            launch_multi_dot(c, U, nl, c->rec_n, b, n, S_GM0);
            allreduce_scalars(c, S_GM0, c->rec_n);
            GL_CUDA(cudaMemcpyAsync(coef, c->scal + S_GM0, sizeof(double) * c->rec_n, cudaMemcpyDeviceToDevice, c->stream));
            launch_multi_axpy(c, U, nl, c->rec_n, coef, 1.0, x, n);
            launch_multi_axpy(c, AU, nl, c->rec_n, coef, -1.0, r, n);
        }
        if (c->rec_n == 0) state(c).rec_hist.clear();
        launch_copy(c, r, r0, n);
    }
    static_assert(S_TMP0 == S_BN + 1, "|b|^2 and |r|^2 travel in one message");
    launch_dot(c, b, b, n, S_BN);
    launch_dot(c, r, r, n, S_TMP0);
    allreduce_scalars(c, S_BN, 2);
    double nrm[2];
    read_scalars(c, S_BN, 2, nrm);
    const double bn2 = nrm[0], rn2 = nrm[1];
    double bn = std::sqrt(bn2), rn0 = std::sqrt(rn2);
    if (which == 1 && std::getenv("GLIMS_VERBOSE")) fprintf(stderr, "glims pcg(u): basis %d, |b| %.3e, |r| after projection %.3e (%.2e of |b|)\n", c->rec_n, bn, rn0, bn > 0 ? rn0 / bn : 0.0);
    if (scale <= 0) scale = bn;
    double tol = std::max(tol_rel * scale, tol_abs);
    if (res_out) *res_out = rn0;
    if (rn0 <= tol || bn == 0.0) return 0;
    double* xbar = nullptr;
    if (recycle) {      // PCG iterates on the correction only; keep x_bar aside
        xbar = ws(c, "rec_xbar", nl);
        launch_copy(c, x, xbar, n);
        launch_zero(c, x, n);
    }
    // Every iteration is the same launch sequence (r.z lives in S_RZ, the new one in S_RZNEW and is shifted down
    // by k_pcg_shift at the end, which also appends r.r to a device ring), so one iteration is captured once into
    // a CUDA graph and replayed: ~50 launches + NCCL calls per iteration collapse to one cudaGraphLaunch.
    apply_pc(c, which, pc, r, z, S_RZ);
    allreduce_scalars(c, S_RZ, 1);
    launch_copy(c, z, p, n);
    double* ring = ws(c, "pcg_ring", 72);            // [64] r.r history, [64] iteration counter
    GL_CUDA(cudaMemsetAsync(ring, 0, sizeof(double) * 72, c->stream));
    if (!c->h_ring) GL_CUDA(cudaMallocHost(&c->h_ring, sizeof(double) * 72));
    auto iteration = [&]() {
        halo_exchange(c, p, bs);
        SpmvDot d; d.w = p; d.slot = S_PAP;
        launch_spmv(c, which, p, Ap, d);
        allreduce_scalars(c, S_PAP, 1);
        launch_cg_update_xr(c, x, r, p, Ap, n, S_RZ, S_PAP, S_RR);
        apply_pc(c, which, pc, r, z, S_RZNEW);
        allreduce_scalars(c, S_RR, 2);               // S_RR and S_RZNEW are adjacent: one message
        launch_cg_update_p(c, p, z, n, S_RZNEW, S_RZ);
        launch_pcg_shift(c, ring);
        GL_CUDA(cudaMemcpyAsync(c->h_ring, ring, sizeof(double) * 65, cudaMemcpyDeviceToHost, c->stream));
    };
    // Convergence test on the device: ring[66] holds tol^2; the captured iteration is the body of a conditional (IF)
    // graph node whose condition is "the last r.r is still above tol^2" (k_pcg_cond).  The host can therefore queue
    // several iterations ahead -- no GPU idling on the host's convergence check -- and iterations queued past
    // convergence cost one tiny kernel instead of a full iteration.
    const double tol2 = tol * tol;
    GL_CUDA(cudaMemcpyAsync(ring + 66, &tol2, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GraphKey key{which, pc, (void*)c->amg, (void*)x, (void*)r, bs};
    PcgGraph* G = c->use_graphs ? &state(c).graphs[key] : nullptr;
    constexpr int NEV = 8;
    cudaEvent_t ev[NEV];
    for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    // cheap iterations (K_cc: one scalar SpMV) are captured several to a body, so that the per-launch cost of the
    // conditional node is amortised; at most unroll-1 extra iterations run after convergence (harmless: alpha and
    // beta guard against 0/0)
    const char* eu = std::getenv("GLIMS_PCG_UNROLL");
    const int unroll_c = eu ? std::max(1, atoi(eu)) : 4;
    const int unroll = which == 2 ? unroll_c : 1;
    const int lookahead = (G && !G->failed) ? (which == 1 ? 2 : 3) : 1;      // launches in flight beyond the one inspected
    int result = -1, launched = 0, checked = 0, seen = 0, cond_launched = 0, plain_its = 0, its_queued = 0;
    std::vector<int> cum;                 // iterations queued up to and including launch i
    bool bad = false;
    while (result < 0 && !bad && (checked < launched || its_queued < maxit)) {
        while (its_queued < maxit && launched < checked + 1 + ((G && G->exec && G->conditional) ? lookahead : 1)) {
            if (G && G->exec) {
                GL_CUDA(cudaGraphLaunch(G->exec, c->stream));
                if (G->conditional) { c->launches += 1; cond_launched++; }      // body kernels are counted once we know they ran
                else c->launches += G->launches;
                its_queued += G->its;
            } else if (G && !G->failed && its_queued >= 1) {
                // second iteration of the first solve with this configuration: all buffers exist -> capture
                i64 l0 = c->launches;
                bool ok = build_pcg_graph(c, G, ring, iteration, unroll);
                G->launches = c->launches - l0;
                c->launches = l0;
                if (!ok) { G->failed = true; cudaGetLastError(); iteration(); plain_its++; its_queued++; }
                else {
                    GL_CUDA(cudaGraphLaunch(G->exec, c->stream));
                    if (G->conditional) { c->launches += 1; cond_launched++; } else c->launches += G->launches;
                    its_queued += G->its;
                }
            } else { iteration(); plain_its++; its_queued++; }
            GL_CUDA(cudaEventRecord(ev[launched % NEV], c->stream));
            cum.push_back(its_queued);
            launched++;
        }
        GL_CUDA(cudaEventSynchronize(ev[checked % NEV]));
        const int expected = cum[checked];
        checked++;
        // Iterations that really ran (skipped bodies do not advance the device counter), but never beyond the launch
        // whose event was just waited for: what a later, still running launch may already have mirrored must not enter
        // the decision, or ranks whose hosts poll at different moments would stop at different iterations and leave
        // unmatched halo exchanges / allreduces behind.  r.r is all-reduced, so with this bound every rank sees the
        // same sequence and queues the same number of launches.
        const int n_done = std::min((int)c->h_ring[64], expected);
        for (int k = seen + 1; k <= n_done && result < 0; ++k) {
            const double rr = c->h_ring[(k - 1) & 63];
            if (res_out) *res_out = std::sqrt(rr);
            if (!(rr == rr)) { bad = true; break; }
            if (rr <= tol2) result = k;
        }
        seen = n_done;
        if (n_done < expected && result < 0) bad = true;      // device stopped but the host test did not fire: NaN
    }
    if (result > maxit) result = maxit;
    for (auto& e : ev) cudaEventDestroy(e);
    if (G && G->conditional && cond_launched > 0)      // kernels of the conditional bodies that really executed
        c->launches += (i64)std::max(0, std::min((seen - plain_its + G->its - 1) / G->its, cond_launched)) * (G->launches - 1);
    if (recycle) {
        if (result >= 0) {
            // A w = r0 - r_final for the correction w = x (PCG started from zero on r0)
            double* Aw = r0;
            launch_axpy(c, -1.0, r, Aw, n);
            // coefficients of this solve's projection part in the current basis (host copy for the bookkeeping)
            std::vector<double> acur(REC_M + 1, 0.0);
            if (c->rec_n > 0) {
                GL_CUDA(cudaMemcpyAsync(c->h_scal + S_GM0, coef, sizeof(double) * c->rec_n, cudaMemcpyDeviceToHost, c->stream));
                GL_CUDA(cudaStreamSynchronize(c->stream));
                for (int k = 0; k < c->rec_n; ++k) acur[k] = c->h_scal[S_GM0 + k];
            }
            auto& Ha = state(c).rec_hist;
            if (c->rec_n >= REC_M) {
                // Basis full: compress it to an A-orthonormal basis of span{last rec_keep()-1 solutions, the
                // projection part of this one}.  U is A-orthonormal, so Gram-Schmidt on the small coefficient
                // vectors is Gram-Schmidt in the A-inner product; the new vectors are U * q_i.
                const int m = c->rec_n;
                std::vector<std::vector<double>> cand;
                cand.push_back(std::vector<double>(acur.begin(), acur.begin() + m));
                for (size_t j = 0; j < Ha.size() && (int)cand.size() < rec_keep(); ++j)
                    cand.push_back(std::vector<double>(Ha[j].begin(), Ha[j].begin() + m));
                std::vector<std::vector<double>> Q;
                for (auto v : cand) {
                    double n0 = 0; for (double t : v) n0 += t * t;
                    for (int pass = 0; pass < 2; ++pass)
                        for (auto& q : Q) { double d = 0; for (int k = 0; k < m; ++k) d += q[k] * v[k]; for (int k = 0; k < m; ++k) v[k] -= d * q[k]; }
                    double n1 = 0; for (double t : v) n1 += t * t;
                    if (n1 > 1e-20 * n0 && n1 > 0) { double inv = 1.0 / std::sqrt(n1); for (auto& t : v) t *= inv; Q.push_back(v); }
                }
                const int kq = (int)Q.size();
                double *T = ws(c, "rec_T", nl * rec_keep()), *AT = ws(c, "rec_AT", nl * rec_keep()), *qd = ws(c, "rec_q", 64);
                for (int i = 0; i < kq; ++i) {
                    GL_CUDA(cudaMemcpyAsync(qd, Q[i].data(), sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
                    launch_zero(c, T + (i64)i * nl, n);
                    launch_zero(c, AT + (i64)i * nl, n);
                    launch_multi_axpy(c, U, nl, m, qd, 1.0, T + (i64)i * nl, n);
                    launch_multi_axpy(c, AU, nl, m, qd, 1.0, AT + (i64)i * nl, n);
                    GL_CUDA(cudaStreamSynchronize(c->stream));      // qd is reused
                }
                for (int i = 0; i < kq; ++i) {
                    launch_copy(c, T + (i64)i * nl, U + (i64)i * nl, n);
                    launch_copy(c, AT + (i64)i * nl, AU + (i64)i * nl, n);
                }
                // re-express the history and the current projection in the new basis: a' = Q^T a
                auto reproject = [&](const std::vector<double>& v) {
                    std::vector<double> o(REC_M + 1, 0.0);
                    for (int i = 0; i < kq; ++i) { double d = 0; for (int k = 0; k < m; ++k) d += Q[i][k] * v[k]; o[i] = d; }
                    return o;
                };
                for (auto& h : Ha) h = reproject(h);
                if ((int)Ha.size() > rec_keep()) Ha.resize(rec_keep());
                acur = reproject(acur);
                c->rec_n = kq;
            }
            {
                // append w, A-orthonormalised against the kept directions (it already is in exact arithmetic)
                const int slot = c->rec_n;
                double *Un = U + (i64)slot * nl, *AUn = AU + (i64)slot * nl;
                launch_copy(c, x, Un, n);
                launch_copy(c, Aw, AUn, n);
                if (slot > 0) {
                    launch_multi_dot(c, AU, nl, slot, Un, n, S_GM0);
                    allreduce_scalars(c, S_GM0, slot);
                    double* coef2 = ws(c, "rec_coef2", 64);
                    GL_CUDA(cudaMemcpyAsync(coef2, c->scal + S_GM0, sizeof(double) * slot, cudaMemcpyDeviceToDevice, c->stream));
                    launch_multi_axpy(c, U, nl, slot, coef2, -1.0, Un, n);
                    launch_multi_axpy(c, AU, nl, slot, coef2, -1.0, AUn, n);
                }
                launch_dot(c, Un, AUn, n, S_TMP0);
                allreduce_scalars(c, S_TMP0, 1);
                double nn; read_scalars(c, S_TMP0, 1, &nn);
                if (nn > 0 && nn == nn) {
                    launch_scale(c, 1.0 / std::sqrt(nn), Un, n);
                    launch_scale(c, 1.0 / std::sqrt(nn), AUn, n);
                    c->rec_n = slot + 1;
                    acur[slot] = std::sqrt(nn);
                }
                Ha.insert(Ha.begin(), acur);            // newest first
                if ((int)Ha.size() > rec_keep()) Ha.resize(rec_keep());
            }
        }
        launch_axpy(c, 1.0, xbar, x, n);      // x = x_bar + correction
    }
    return result;
}

// Right-preconditioned restarted GMRES(m) on the monolithic Jacobian, block-Jacobi PC, zero initial guess.
int gmres_mono(glims_ctx* c, const double* b, double* x, double tol, int maxit, double* res_out) {
    const int m = 30;
    const i64 n = nrows(c) * c->nb, nl = c->n_v * c->nb;
    double* V = ws(c, "gm_V", nl * (m + 1));
    double* w = ws(c, "gm_w", nl);
    double* z = ws(c, "gm_z", nl);
    double* coef = ws(c, "gm_coef", 64);
    launch_zero(c, x, n);
    std::vector<double> H((m + 1) * m), cs(m), sn(m), g(m + 1), y(m);
    int total = 0;
    bool first = true;
    double beta = 0;
    while (total < maxit) {
        // r = b - A x
        if (first) launch_copy(c, b, w, n);
        else {
            halo_exchange(c, x, c->nb);
            launch_spmv(c, 0, x, w);
            launch_scale(c, -1.0, w, n);
            launch_axpy(c, 1.0, b, w, n);
        }
        launch_dot(c, w, w, n, S_TMP0);
        allreduce_scalars(c, S_TMP0, 1);
        double b2; read_scalars(c, S_TMP0, 1, &b2);
        beta = std::sqrt(b2);
        if (res_out) *res_out = beta;
        if (beta <= tol) return total;
        first = false;
        launch_copy(c, w, V, n);
        launch_scale(c, 1.0 / beta, V, n);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int j = 0;
        for (; j < m && total < maxit; ++j, ++total) {
            launch_block_jacobi(c, c->dinv_mono, c->nb, V + (i64)j * nl, z, nrows(c), -1);
            halo_exchange(c, z, c->nb);
            launch_spmv(c, 0, z, w);
            // classical Gram-Schmidt, twice (CGS2): batched dots then one fused update
            std::vector<double> h(j + 2, 0.0), h2(j + 1);
            for (int pass = 0; pass < 2; ++pass) {
                launch_multi_dot(c, V, nl, j + 1, w, n, S_GM0);
                allreduce_scalars(c, S_GM0, j + 1);
                GL_CUDA(cudaMemcpyAsync(coef, c->scal + S_GM0, sizeof(double) * (j + 1), cudaMemcpyDeviceToDevice, c->stream));
                launch_multi_axpy(c, V, nl, j + 1, coef, -1.0, w, n);
                read_scalars(c, S_GM0, j + 1, h2.data());
                for (int i = 0; i <= j; ++i) h[i] += h2[i];
            }
            launch_dot(c, w, w, n, S_TMP0);
            allreduce_scalars(c, S_TMP0, 1);
            double w2; read_scalars(c, S_TMP0, 1, &w2);
            h[j + 1] = std::sqrt(w2);
            launch_copy(c, w, V + (i64)(j + 1) * nl, n);
            if (h[j + 1] > 0) launch_scale(c, 1.0 / h[j + 1], V + (i64)(j + 1) * nl, n);
            for (int i = 0; i < j; ++i) {
                double t = cs[i] * h[i] + sn[i] * h[i + 1];
                h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1];
                h[i] = t;
            }
            double den = std::hypot(h[j], h[j + 1]);
            cs[j] = den > 0 ? h[j] / den : 1.0;
            sn[j] = den > 0 ? h[j + 1] / den : 0.0;
            h[j] = den;
            g[j + 1] = -sn[j] * g[j];
            g[j] = cs[j] * g[j];
            for (int i = 0; i <= j; ++i) H[i * m + j] = h[i];
            if (res_out) *res_out = std::fabs(g[j + 1]);
            if (std::fabs(g[j + 1]) <= tol) { ++j; ++total; break; }
        }
        // y = H^-1 g ; x += M^-1 (V y)
        for (int i = j - 1; i >= 0; --i) {
            double s = g[i];
            for (int k = i + 1; k < j; ++k) s -= H[i * m + k] * y[k];
            y[i] = s / H[i * m + i];
        }
        GL_CUDA(cudaMemcpyAsync(coef, y.data(), sizeof(double) * j, cudaMemcpyHostToDevice, c->stream));
        GL_CUDA(cudaStreamSynchronize(c->stream));
        launch_zero(c, w, n);
        launch_multi_axpy(c, V, nl, j, coef, 1.0, w, n);
        launch_block_jacobi(c, c->dinv_mono, c->nb, w, z, nrows(c), -1);
        launch_axpy(c, 1.0, z, x, n);
        if (std::fabs(g[j]) <= tol && j <= m) {
            if (res_out) *res_out = std::fabs(g[j]);
            return total;
        }
    }
    return -1;
}

// The row-walk assembly (GLIMS_ASMK_ROWS) is the default of the block-triangular solver; it needs the pair lists
// (cc_available) and, for F_u by SpMV, K_uu / K_uc assembled.  Anything else falls back to the element kernels.
bool use_rows(glims_ctx* c, const glims_solver_opts* o) {
    return o->asm_kernel == GLIMS_ASMK_ROWS && cc_available(c);
}
int kconst_kernel(glims_ctx* c, const glims_solver_opts* o) {
    return o->asm_kernel == GLIMS_ASMK_ROWS ? GLIMS_ASMK_TILE : o->asm_kernel;
}

void ensure_kconst(glims_ctx* c, const glims_solver_opts* o) {
    bool need_amg = (o->pc == GLIMS_PC_AMG || o->pc == GLIMS_PC_AMG_FP64) && o->solver == GLIMS_SOLVER_BLOCK_TRI;
    const bool rows = use_rows(c, o);
    if (!c->kconst_valid) {
        launch_assemble(c, GLIMS_ASM_KCONST, kconst_kernel(c, o));
        if (rows) cc_compute_lift(c, c->bc_nonzero);
        launch_bc_matrix(c, GLIMS_ASM_KCONST, true);
        launch_diag_inverse(c, 1);
        c->kconst_valid = true;
        c->kuu_state = 3;
        c->lift_dirty = false;
        c->rec_n = c->rec_head = 0;
        free_graphs(c);
        if (c->amg) amg_free(c);
    } else if (c->lift_dirty) {
        // Dirichlet VALUES changed on an unchanged dof set (time-dependent boundary data, helper_classes.py
        // time_update_bcs): the eliminated matrices, the AMG hierarchy, the captured graphs and the projection basis all
        // stay valid (symmetric elimination depends on the constrained set only).  Only the lift K_raw(:, bc) g of the
        // SpMV residual needs the raw columns once more.
        if (rows && (c->bc_nonzero || c->kuu_state != 3)) {
            launch_assemble(c, GLIMS_ASM_KCONST, kconst_kernel(c, o));
            cc_compute_lift(c, c->bc_nonzero);
            launch_bc_matrix(c, GLIMS_ASM_KCONST, true);
            c->kuu_state = 3;
        }
        c->lift_dirty = false;
    }
    if (need_amg && !c->amg) amg_setup(c);
}

void newton_step(glims_ctx* c, const glims_solver_opts* o, glims_step_stats* st) {
    EvTimer tm(c->stream);
    tm.begin(0);
    const int D = c->dim, NB = c->nb;
    const i64 nr = nrows(c);
    ensure_kconst(c, o);
    const bool rows = use_rows(c, o);
    const bool mono = o->solver == GLIMS_SOLVER_MONO_GMRES;
    const int ek = rows ? GLIMS_ASMK_ATOMIC : o->asm_kernel;      // element kernel for the rare K_cc-only re-assembly
    launch_bc_values(c, c->x);
    double *Fu = ws(c, "Fu", c->n_v * D), *Fc = ws(c, "Fc", c->n_v);
    double *du = ws(c, "du", c->n_v * D), *dc = ws(c, "dc", c->n_v), *tmpu = ws(c, "tmpu", c->n_v * D);
    glims_step_stats s;
    memset(&s, 0, sizeof s);
    double f0 = -1, fu_scale = 0, T = 0;
    double fu2 = 0.0;                   // |F_u|^2 of the last evaluation
    bool fu_current = false;            // F's displacement rows belong to the current iterate
    bool converged = false, c_was_done = false;
    if (rows) {
        tm.begin(1);
        cc_mass_cprev(c);               // M c_prev: constant during the step
        tm.end(1);
    }
    // displacement rows of the residual, row-walk path: one SpMV over the stored (eliminated) blocks + the lift
    auto eval_fu = [&]() {
        launch_fu(c, true);
        launch_bc_residual(c, c->F, c->x);
        launch_split_norms(c, c->F, S_TMP0);
        allreduce_scalars(c, S_TMP0, 2);
        double nn[2];
        read_scalars(c, S_TMP0, 2, nn);
        fu2 = nn[0];
        fu_current = true;
        c->fu2_cached = fu2; c->fu_cache_valid = true;
    };
    for (int k = 0; k <= o->max_newton; ++k) {
        tm.begin(1);
        halo_exchange(c, c->x, NB);
        // one pass: residual, plus K_cc while the concentration block is still iterating
        const bool with_kcc = mono || !(o->lag_mechanics && c_was_done);
        double fu, fc, fn;
        bool fu_known = true;           // fn contains the displacement rows of THIS iterate (evaluated, or cached at k = 0)
        if (rows) {
            // concentration rows first (F_c, K_cc); the displacement rows cost an SpMV over K_uu/K_uc and are evaluated
            // only when the decision needs them: at the first iterate (unless the previous step's final residual still
            // describes this state), when |F_c| is within the tolerance, and whenever mechanics is not lagged
            launch_cc_rows(c, with_kcc, true);
            launch_bc_residual(c, c->F, c->x);
            launch_split_norms(c, c->F, S_TMP0);
            allreduce_scalars(c, S_TMP0, 2);
            double nn[2];
            read_scalars(c, S_TMP0, 2, nn);
            fc = std::sqrt(nn[1]);
            fu_current = false;
            const bool reuse = (k == 0 && c->fu_cache_valid && !mono && o->lag_mechanics);
            const bool need_fu = mono || !o->lag_mechanics || (k == 0 ? !reuse : fc <= T);
            if (need_fu) eval_fu();
            else if (reuse) fu2 = c->fu2_cached;
            fu_known = need_fu || reuse;
            fu = std::sqrt(fu2);
            fn = std::sqrt(fu2 + nn[1]);
        } else {
            launch_assemble(c, GLIMS_ASM_RESIDUAL | (with_kcc ? GLIMS_ASM_KCC : 0), o->asm_kernel);
            launch_bc_residual(c, c->F, c->x);
            launch_split_norms(c, c->F, S_TMP0);
            allreduce_scalars(c, S_TMP0, 2);
            double nn[2];
            read_scalars(c, S_TMP0, 2, nn);
            fu = std::sqrt(nn[0]); fc = std::sqrt(nn[1]); fn = std::sqrt(nn[0] + nn[1]);
            fu_current = true;
        }
        tm.end(1);
        if (!(fn == fn)) throw GlError(GLIMS_ERR_NOT_CONVERGED, "Newton: residual is NaN");
        if (k == 0) {
            f0 = fn; s.fnorm0 = fn;
            T = std::max(o->snes_rtol * f0, o->snes_atol);
            // scale for the mechanics tolerance: |f_ext_u - K_uc c| or the first |F_u|, whichever is larger
            launch_extract(c, c->x, nullptr, dc);
            halo_exchange(c, dc, 1);
            launch_spmv_uc(c, dc, tmpu);
            launch_dot(c, tmpu, tmpu, nr * D, S_TMP2);
            allreduce_scalars(c, S_TMP2, 1);
            double bu2; read_scalars(c, S_TMP2, 1, &bu2);
            fu_scale = std::max(std::sqrt(bu2), fu);
        }
        s.fnorm = fn;
        s.newton_its = k;
        // the stopping test is on the whole residual; an iterate whose displacement rows were skipped has |F_c| > T
        if (fu_known && (fn < o->snes_atol || fn <= o->snes_rtol * f0)) { converged = true; break; }
        if (k == o->max_newton) break;

        tm.begin(2);
        // monolithic update: GMRES(30) on J with block-Jacobi; also the fallback when a block solve breaks down
        // (e.g. K_cc indefinite for dt*rho > 1, where PCG is not applicable but the reference's GMRES still is)
        auto mono_update = [&](bool kcc_eliminated) {
            if (rows && !fu_current) eval_fu();
            if (!kcc_eliminated) launch_bc_matrix(c, GLIMS_ASM_KCC, true);
            launch_diag_inverse(c, 0);
            double* rhs = ws(c, "rhs_mono", c->ndof);
            launch_copy(c, c->F, rhs, nr * NB);
            launch_scale(c, -1.0, rhs, nr * NB);
            double res = 0;
            int its = gmres_mono(c, rhs, c->dx, std::max(o->ksp_rtol * fn, o->ksp_atol), o->max_krylov, &res);
            if (its < 0) throw GlError(GLIMS_ERR_NOT_CONVERGED, "GMRES did not converge");
            s.krylov_its_mono += its;
            launch_axpy(c, 1.0, c->dx, c->x, nr * NB);
        };
        auto assemble_kcc_only = [&]() {
            if (rows) launch_cc_rows(c, true, false); else launch_assemble(c, GLIMS_ASM_KCC, ek);
        };
        c->fu_cache_valid = false;          // the iterate is about to change
        if (mono) {
            mono_update(false);
        } else {
            const bool c_done = fc <= 0.5 * T;
            c_was_done = c_was_done || c_done;
            const bool solve_c = fc > 0.0 && (!c_done || !o->lag_mechanics);
            const bool solve_u_now = c_done || !o->lag_mechanics;
            if (rows && solve_u_now && !fu_current) { eval_fu(); fu = std::sqrt(fu2); }
            launch_extract(c, c->F, solve_u_now ? Fu : nullptr, Fc);
            bool broke = false, kcc_elim = false;
            if (solve_c && !with_kcc) assemble_kcc_only();
            if (solve_c) {
                kcc_elim = true;
                launch_bc_matrix(c, GLIMS_ASM_KCC, true);
                launch_diag_inverse(c, 2);
                launch_scale(c, -1.0, Fc, nr);
                double res = 0;
                int its = pcg(c, 2, GLIMS_PC_JACOBI, Fc, dc, o->ksp_rtol, fc, o->ksp_atol, o->max_krylov, &res);
                if (its < 0) broke = true; else s.krylov_its_c += its;
            } else launch_zero(c, dc, nr);
            if (broke) {
                if (!with_kcc && !solve_c) assemble_kcc_only();
                mono_update(kcc_elim);
            } else if (solve_u_now && fu > 0.0) {
                // rhs_u = -F_u - K_uc dc
                if (solve_c) {
                    halo_exchange(c, dc, 1);
                    launch_spmv_uc(c, dc, tmpu);
                    launch_axpy(c, 1.0, tmpu, Fu, nr * D);
                }
                launch_scale(c, -1.0, Fu, nr * D);
                double res = 0;
                int its = pcg(c, 1, o->pc, Fu, du, o->ksp_rtol, fu_scale, o->ksp_atol, o->max_krylov, &res, o->recycle != 0);
                if (its < 0) {
                    // K_uu PCG broke down: take the monolithic update instead (needs the current K_cc)
                    if (!with_kcc && !solve_c) assemble_kcc_only();
                    mono_update(kcc_elim);
                } else {
                    s.krylov_its_u += its;
                    launch_insert_add(c, c->x, du, dc, 1.0);
                }
            } else launch_insert_add(c, c->x, nullptr, dc, 1.0);
        }
        tm.end(2);
    }
    tm.end(0);
    s.converged = converged;
    s.ms_total = tm.total(0);
    s.ms_assembly = tm.total(1);
    s.ms_krylov = tm.total(2);
    if (st) *st = s;
    if (!converged) { c->fu_cache_valid = false; throw GlError(GLIMS_ERR_NOT_CONVERGED, "Newton did not converge"); }
}

}  // namespace

// ================================================================================================
void solver_free_graphs(glims_ctx* c) { free_graphs(c); }

#define API_BEGIN if (!c) return GLIMS_ERR_ARG; try { GL_CUDA(cudaSetDevice(c->device));
#define API_END } catch (const GlError& e) { c->err = e.msg; return e.code; } catch (const std::exception& e) { c->err = e.what(); return GLIMS_ERR_CUDA; } return GLIMS_OK;

extern "C" {

void glims_default_opts(glims_solver_opts* o) {
    o->snes_rtol = 1e-9; o->snes_atol = 1e-10; o->snes_stol = 1e-16; o->max_newton = 50;
    o->ksp_rtol = 1e-10; o->ksp_atol = 1e-300; o->max_krylov = 20000;
    o->solver = GLIMS_SOLVER_BLOCK_TRI; o->pc = GLIMS_PC_AMG; o->asm_kernel = GLIMS_ASMK_ROWS;
    o->lag_mechanics = 1;
    o->recycle = 1;
    o->extrapolate = 0;
}

int glims_create(glims_ctx** out, int32_t dim, int64_t n_vertices, const double* coords, int64_t n_cells,
                 const int32_t* cells, const int32_t* cell_mat, int64_t n_owned, int32_t device) {
    if (!out || (dim != 2 && dim != 3) || n_vertices <= 0 || n_cells <= 0 || !coords || !cells || !cell_mat)
        return GLIMS_ERR_ARG;
    glims_ctx* c = new glims_ctx();
    *out = c;
    c->device = device;
    c->use_graphs = std::getenv("GLIMS_NO_GRAPH") == nullptr;
    API_BEGIN
    c->dim = dim; c->nb = dim + 1; c->n_v = n_vertices; c->n_c = n_cells; c->ndof = n_vertices * c->nb;
    if (n_owned >= 0 && n_owned < n_vertices) { c->halo.active = true; c->halo.n_owned = n_owned; }
    GL_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->coords = dev_upload(coords, n_vertices * dim, c->stream);
    c->cells = dev_upload(cells, n_cells * c->nb, c->stream);
    c->cell_mat = dev_upload(cell_mat, n_cells, c->stream);
    GL_CUDA(cudaMalloc(&c->scal, sizeof(double) * S_COUNT));
    GL_CUDA(cudaMemsetAsync(c->scal, 0, sizeof(double) * S_COUNT, c->stream));
    GL_CUDA(cudaMalloc(&c->partials, sizeof(double) * S_COUNT * 148 * 8));
    GL_CUDA(cudaMalloc(&c->tickets, sizeof(unsigned) * S_COUNT));
    GL_CUDA(cudaMemsetAsync(c->tickets, 0, sizeof(unsigned) * S_COUNT, c->stream));
    GL_CUDA(cudaMallocHost(&c->h_scal, sizeof(double) * (S_COUNT + 8)));
    build_pattern(c);
    i64 ns = c->pat.n_slots;
    GL_CUDA(cudaMalloc(&c->Kuu, sizeof(double) * ns * dim * dim));
    GL_CUDA(cudaMalloc(&c->Kuc, sizeof(double) * ns * dim));
    GL_CUDA(cudaMalloc(&c->Kcc, sizeof(double) * ns));
    GL_CUDA(cudaMemsetAsync(c->Kuu, 0, sizeof(double) * ns * dim * dim, c->stream));
    GL_CUDA(cudaMemsetAsync(c->Kuc, 0, sizeof(double) * ns * dim, c->stream));
    GL_CUDA(cudaMemsetAsync(c->Kcc, 0, sizeof(double) * ns, c->stream));
    i64 nr = c->pat.n_rows;
    GL_CUDA(cudaMalloc(&c->dinv_uu, sizeof(double) * nr * dim * dim));
    GL_CUDA(cudaMalloc(&c->dinv_cc, sizeof(double) * nr));
    GL_CUDA(cudaMalloc(&c->dinv_mono, sizeof(double) * nr * c->nb * c->nb));
    for (double** v : {&c->x, &c->xprev, &c->F, &c->fext, &c->dx}) {
        GL_CUDA(cudaMalloc(v, sizeof(double) * c->ndof));
        GL_CUDA(cudaMemsetAsync(*v, 0, sizeof(double) * c->ndof, c->stream));
    }
    GL_CUDA(cudaMalloc(&c->bcmask, c->n_v));
    GL_CUDA(cudaMemsetAsync(c->bcmask, 0, c->n_v, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}

int glims_destroy(glims_ctx* c) {
    if (!c) return GLIMS_ERR_ARG;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->amg) amg_free(c);
    tile_free(c);
    cc_free(c);
    comm_free(c);
    free_pool(c);
    auto& p = c->pat;
    for (void* q : {(void*)c->coords, (void*)c->cells, (void*)c->cell_mat, (void*)c->mat, (void*)p.slice_off,
                    (void*)p.slice_w, (void*)p.col, (void*)p.rowptr, (void*)p.diag, (void*)c->eslot, (void*)c->gptr,
                    (void*)c->gent, (void*)c->Kuu, (void*)c->Kuc, (void*)c->Kcc, (void*)c->dinv_uu, (void*)c->dinv_cc,
                    (void*)c->dinv_mono, (void*)c->x, (void*)c->xprev, (void*)c->F, (void*)c->fext, (void*)c->dx,
                    (void*)c->bc_dofs, (void*)c->bc_vals, (void*)c->bcmask, (void*)c->scal, (void*)c->partials,
                    (void*)c->tickets, (void*)c->flush_buf, (void*)c->sl_ptr, (void*)c->sl_elem, (void*)c->lent, (void*)c->halo.send_idx, (void*)c->halo.send_buf, (void*)c->dof_perm, (void*)c->dinv_mass})
        if (q) cudaFree(q);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    if (c->h_ring) cudaFreeHost(c->h_ring);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return GLIMS_OK;
}

const char* glims_last_error(const glims_ctx* c) { return c ? c->err.c_str() : "null context"; }

int glims_set_materials(glims_ctx* c, int32_t n_mat, const double* table) {
    API_BEGIN
    if (n_mat <= 0 || n_mat > MAX_MAT || !table) throw GlError(GLIMS_ERR_ARG, "set_materials: need 1..64 rows");
    std::vector<double> t(n_mat * MAT_STRIDE);
    for (int m = 0; m < n_mat; ++m) {
        for (int k = 0; k < 5; ++k) t[m * MAT_STRIDE + k] = table[m * 5 + k];
        t[m * MAT_STRIDE + 5] = (2.0 * table[m * 5 + 0] + c->dim * table[m * 5 + 1]) * table[m * 5 + 4];
    }
    if (c->mat) cudaFree(c->mat);
    c->mat = dev_upload(t.data(), (i64)t.size(), c->stream);
    GL_CUDA(cudaStreamSynchronize(c->stream));
    c->n_mat = n_mat; c->have_mat = true; c->kconst_valid = false;
    c->kuu_state = 0; c->fu_cache_valid = false;
    cc_invalidate_consts(c);
    API_END
}

int glims_set_dt(glims_ctx* c, double dt) {
    if (!c) return GLIMS_ERR_ARG;
    if (dt != c->dt) cc_invalidate_consts(c);      // Klin = (1 - dt rho) M + dt D K depends on dt
    c->dt = dt;
    return GLIMS_OK;
}

int glims_set_dirichlet(glims_ctx* c, int64_t n, const int64_t* dofs, const double* vals) {
    API_BEGIN
    if (n < 0 || (n > 0 && (!dofs || !vals))) throw GlError(GLIMS_ERR_ARG, "set_dirichlet: bad arguments");
    for (i64 t = 0; t < n; ++t)
        if (dofs[t] < 0 || dofs[t] >= c->ndof) throw GlError(GLIMS_ERR_ARG, "set_dirichlet: dof out of range");
    std::vector<int64_t> mapped;
    if (!c->h_dof_perm.empty()) {           // caller numbering -> vertex-blocked
        mapped.resize(n);
        for (i64 t = 0; t < n; ++t) mapped[t] = c->h_dof_perm[dofs[t]];
        dofs = mapped.data();
    }
    bool nonzero = false;
    for (i64 t = 0; t < n; ++t) nonzero = nonzero || vals[t] != 0.0;
    c->fu_cache_valid = false;
    // Same constrained set as before: only the values change (time-dependent boundary data).  Symmetric elimination
    // depends on the set alone, so K_uu / K_uc, the AMG hierarchy, the captured graphs and the projection basis stay.
    const bool same_set = (i64)c->h_bc_dofs.size() == n && c->n_bc == n &&
                          std::equal(c->h_bc_dofs.begin(), c->h_bc_dofs.end(), (const long long*)dofs);
    if (same_set && n > 0) {
        GL_CUDA(cudaMemcpyAsync(c->bc_vals, vals, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        GL_CUDA(cudaStreamSynchronize(c->stream));
        if (nonzero || c->bc_nonzero) c->lift_dirty = true;
        c->bc_nonzero = nonzero;
        return GLIMS_OK;
    }
    if (same_set) return GLIMS_OK;      // still no Dirichlet condition
    if (c->bc_dofs) { cudaFree(c->bc_dofs); c->bc_dofs = nullptr; }
    if (c->bc_vals) { cudaFree(c->bc_vals); c->bc_vals = nullptr; }
    std::vector<unsigned char> mask(c->n_v, 0);
    c->n_bc_u = c->n_bc_c = 0;
    for (i64 t = 0; t < n; ++t) {
        i64 v = dofs[t] / c->nb; int k = (int)(dofs[t] - v * c->nb);
        mask[v] |= (unsigned char)(1u << k);
        if (k < c->dim) c->n_bc_u++; else c->n_bc_c++;
    }
    c->n_bc = n;
    c->h_bc_dofs.assign((const long long*)dofs, (const long long*)dofs + n);
    c->bc_nonzero = nonzero;
    if (n > 0) {
        c->bc_dofs = dev_upload((const i64*)dofs, n, c->stream);
        c->bc_vals = dev_upload(vals, n, c->stream);
    }
    GL_CUDA(cudaMemcpyAsync(c->bcmask, mask.data(), c->n_v, cudaMemcpyHostToDevice, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    c->kconst_valid = false;
    API_END
}

static int copy_vec(glims_ctx* c, double* dev, double* host_out, const double* host_in);

int glims_set_load(glims_ctx* c, const double* f_ext) {
    API_BEGIN
    c->have_load = f_ext != nullptr;
    c->fu_cache_valid = false;
    if (f_ext) {
        int rc = copy_vec(c, c->fext, nullptr, f_ext);
        if (rc != GLIMS_OK) return rc;
    }
    API_END
}

// Host <-> device copy of an ndof vector in the CALLER's dof numbering (glims_set_dof_permutation; identity by default)
static int copy_vec(glims_ctx* c, double* dev, double* host_out, const double* host_in) {
    API_BEGIN
    const size_t nbytes = sizeof(double) * c->ndof;
    if (!c->dof_perm) {
        if (host_in) GL_CUDA(cudaMemcpyAsync(dev, host_in, nbytes, cudaMemcpyHostToDevice, c->stream));
        if (host_out) GL_CUDA(cudaMemcpyAsync(host_out, dev, nbytes, cudaMemcpyDeviceToHost, c->stream));
    } else {
        double* stage = ws(c, "perm_stage", c->ndof);
        if (host_in) {
            GL_CUDA(cudaMemcpyAsync(stage, host_in, nbytes, cudaMemcpyHostToDevice, c->stream));
            launch_permute(c, stage, dev, c->dof_perm, c->ndof, true);       // dev[perm[i]] = stage[i]
        }
        if (host_out) {
            launch_permute(c, dev, stage, c->dof_perm, c->ndof, false);      // stage[i] = dev[perm[i]]
            GL_CUDA(cudaMemcpyAsync(host_out, stage, nbytes, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    GL_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}
int glims_set_state(glims_ctx* c, const double* x) { if (c) { c->have_hist = false; c->fu_cache_valid = false; } return x ? copy_vec(c, c ? c->x : nullptr, nullptr, x) : GLIMS_ERR_ARG; }
int glims_get_state(glims_ctx* c, double* x) { return x ? copy_vec(c, c ? c->x : nullptr, x, nullptr) : GLIMS_ERR_ARG; }
int glims_set_prev(glims_ctx* c, const double* x) { if (c) c->have_hist = false; return x ? copy_vec(c, c ? c->xprev : nullptr, nullptr, x) : GLIMS_ERR_ARG; }
int glims_get_prev(glims_ctx* c, double* x) { return x ? copy_vec(c, c ? c->xprev : nullptr, x, nullptr) : GLIMS_ERR_ARG; }
int64_t glims_ndof(const glims_ctx* c) { return c ? c->ndof : 0; }
int64_t glims_nnzb(const glims_ctx* c) { return c ? c->pat.nnzb : 0; }
int64_t glims_nslots(const glims_ctx* c) { return c ? c->pat.n_slots : 0; }
void* glims_state_dev(glims_ctx* c) { return c ? c->x : nullptr; }
void* glims_stream(glims_ctx* c) { return c ? (void*)c->stream : nullptr; }
int64_t glims_launch_count(const glims_ctx* c) { return c ? c->launches : 0; }

int glims_step(glims_ctx* c, int32_t n_steps, const glims_solver_opts* o, glims_step_stats* stats) {
    API_BEGIN
    if (!c->have_mat) throw GlError(GLIMS_ERR_STATE, "glims_step before glims_set_materials");
    glims_solver_opts od;
    if (!o) { glims_default_opts(&od); o = &od; }
    double* backup = ws(c, "step_backup", c->ndof);
    for (int s = 0; s < n_steps; ++s) {
        // keep the last good iterate: a failed step leaves the state where the previous step ended
        launch_copy(c, c->x, backup, c->ndof);
        if (o->extrapolate && c->have_hist) {
            // first Newton guess: c_n + (c_n - c_{n-1}); only the start of the iteration changes, not its fixed point
            launch_extrapolate_c(c, c->x, ws(c, "x_hist", c->ndof));
            c->fu_cache_valid = false;
        }
        try {
            newton_step(c, o, stats ? &stats[s] : nullptr);
        } catch (const GlError&) {
            launch_copy(c, backup, c->x, c->ndof);
            cudaStreamSynchronize(c->stream);
            c->fu_cache_valid = false;
            throw;
        }
        // u_previous.assign(solution)  (simulation_base.py:312); ghosts refreshed first so that the next
        // step's residual sees the converged c_prev on the overlap cells
        halo_exchange(c, c->x, c->nb);
        launch_copy(c, c->xprev, ws(c, "x_hist", c->ndof), c->ndof);     // solution of the step before this one
        c->have_hist = true;
        launch_copy(c, c->x, c->xprev, c->ndof);
    }
    GL_CUDA(cudaStreamSynchronize(c->stream));
    comm_check(c);
    amg_check(c);
    API_END
}

int glims_prepare(glims_ctx* c, const glims_solver_opts* o) {
    API_BEGIN
    if (!c->have_mat) throw GlError(GLIMS_ERR_STATE, "glims_prepare before glims_set_materials");
    glims_solver_opts od;
    if (!o) { glims_default_opts(&od); o = &od; }
    const bool verbose = std::getenv("GLIMS_VERBOSE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    const bool rows = use_rows(c, o);            // builds the (row, element) pair lists on first use
    GL_CUDA(cudaStreamSynchronize(c->stream));
    double t1 = now();
    ensure_kconst(c, o);
    GL_CUDA(cudaStreamSynchronize(c->stream));
    double t2 = now();
    if (rows) cc_mass_cprev(c);                  // also builds the per-slot constants
    GL_CUDA(cudaStreamSynchronize(c->stream));
    double t3 = now();
    if (verbose) fprintf(stderr, "glims prepare: pair lists %.3f s, K_uu/K_uc + elimination + AMG %.3f s, per-slot constants %.3f s\n", t1 - t0, t2 - t1, t3 - t2);
    API_END
}

int glims_reset_history(glims_ctx* c) {
    API_BEGIN
    GL_CUDA(cudaStreamSynchronize(c->stream));
    c->rec_n = c->rec_head = 0;
    if (c->solver_state) state(c).rec_hist.clear();
    c->have_hist = false;
    c->fu_cache_valid = false;
    API_END
}

int glims_set_dof_permutation(glims_ctx* c, const int64_t* perm) {
    API_BEGIN
    GL_CUDA(cudaStreamSynchronize(c->stream));
    if (c->dof_perm) { cudaFree(c->dof_perm); c->dof_perm = nullptr; }
    c->h_dof_perm.clear();
    if (perm) {
        std::vector<char> seen(c->ndof, 0);
        for (i64 i = 0; i < c->ndof; ++i) {
            if (perm[i] < 0 || perm[i] >= c->ndof || seen[perm[i]]) throw GlError(GLIMS_ERR_ARG, "set_dof_permutation: not a permutation of [0, ndof)");
            seen[perm[i]] = 1;
        }
        if (c->n_bc > 0) throw GlError(GLIMS_ERR_STATE, "set_dof_permutation: call before glims_set_dirichlet");
        c->h_dof_perm.assign(perm, perm + c->ndof);
        c->dof_perm = dev_upload((const i64*)perm, c->ndof, c->stream);
        GL_CUDA(cudaStreamSynchronize(c->stream));
    }
    API_END
}

int glims_get_dof_permutation(glims_ctx* c, int64_t* perm) {
    if (!c || !perm) return GLIMS_ERR_ARG;
    for (i64 i = 0; i < c->ndof; ++i) perm[i] = c->h_dof_perm.empty() ? i : c->h_dof_perm[i];
    return GLIMS_OK;
}

int glims_assemble(glims_ctx* c, int32_t what, int32_t kernel, int32_t apply_bc) {
    API_BEGIN
    if (!c->have_mat) throw GlError(GLIMS_ERR_STATE, "glims_assemble before glims_set_materials");
    halo_exchange(c, c->x, c->nb);
    launch_assemble(c, what, kernel);
    if (apply_bc) {
        if (what & GLIMS_ASM_RESIDUAL) launch_bc_residual(c, c->F, c->x);
        if (what & GLIMS_ASM_JACOBIAN) launch_bc_matrix(c, what & GLIMS_ASM_JACOBIAN, apply_bc == 2);
    }
    if (what & GLIMS_ASM_KCONST) {
        c->kconst_valid = (apply_bc == 2);
        if (apply_bc) c->kuu_state = apply_bc == 2 ? 3 : 2;
        c->lift_dirty = true;           // no lift was computed for these matrices: glims_step rebuilds it if it needs one
        if (c->kconst_valid) launch_diag_inverse(c, 1);
        free_graphs(c);
        c->rec_n = 0;
        if (c->amg) amg_free(c);
    }
    GL_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}

int glims_get_residual(glims_ctx* c, double* F) { return F ? copy_vec(c, c ? c->F : nullptr, F, nullptr) : GLIMS_ERR_ARG; }

int glims_export_pattern(glims_ctx* c, int64_t* rowptr, int32_t* colidx) {
    API_BEGIN
    auto& p = c->pat;
    std::vector<i64> so(p.n_slices + 1);
    std::vector<int> col(p.n_slots);
    GL_CUDA(cudaMemcpy(rowptr, p.rowptr, sizeof(i64) * (p.n_rows + 1), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(so.data(), p.slice_off, sizeof(i64) * (p.n_slices + 1), cudaMemcpyDeviceToHost));
    GL_CUDA(cudaMemcpy(col.data(), p.col, sizeof(int) * p.n_slots, cudaMemcpyDeviceToHost));
    for (i64 r = 0; r < p.n_rows; ++r)
        for (i64 t = rowptr[r]; t < rowptr[r + 1]; ++t)
            colidx[t] = col[so[r >> 5] + (t - rowptr[r]) * 32 + (r & 31)];
    API_END
}

int glims_export_values(glims_ctx* c, double* Kuu, double* Kuc, double* Kcc) {
    API_BEGIN
    i64 nz = c->pat.nnzb; int D = c->dim;
    double *a = ws(c, "exp_uu", nz * D * D), *b = ws(c, "exp_uc", nz * D), *d = ws(c, "exp_cc", nz);
    launch_export_values(c, a, b, d);
    if (Kuu) GL_CUDA(cudaMemcpyAsync(Kuu, a, sizeof(double) * nz * D * D, cudaMemcpyDeviceToHost, c->stream));
    if (Kuc) GL_CUDA(cudaMemcpyAsync(Kuc, b, sizeof(double) * nz * D, cudaMemcpyDeviceToHost, c->stream));
    if (Kcc) GL_CUDA(cudaMemcpyAsync(Kcc, d, sizeof(double) * nz, cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}

int glims_spmv(glims_ctx* c, int32_t which, const double* x, double* y) {
    API_BEGIN
    int bs = which == 0 ? c->nb : which == 1 ? c->dim : 1;
    if (which < 0 || which > 2 || !x || !y) throw GlError(GLIMS_ERR_ARG, "glims_spmv: bad arguments");
    double *dx = ws(c, "spmv_x", c->n_v * bs), *dy = ws(c, "spmv_y", c->n_v * bs);
    GL_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * c->n_v * bs, cudaMemcpyHostToDevice, c->stream));
    halo_exchange(c, dx, bs);
    launch_spmv(c, which, dx, dy);
    GL_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * c->pat.n_rows * bs, cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}

int glims_cell_fields(glims_ctx* c, double* cell_out, double* vertex_out) {
    API_BEGIN
    if (!c->have_mat) throw GlError(GLIMS_ERR_STATE, "glims_cell_fields before glims_set_materials");
    const int nf = 2 * c->dim * c->dim + 5;
    double *q = ws(c, "pp_q", c->n_c * nf), *vol = ws(c, "pp_vol", c->n_c);
    launch_cell_fields(c, q, vol);
    if (cell_out) GL_CUDA(cudaMemcpyAsync(cell_out, q, sizeof(double) * c->n_c * nf, cudaMemcpyDeviceToHost, c->stream));
    if (vertex_out) {
        double *num = ws(c, "pp_num", c->n_v * nf), *den = ws(c, "pp_den", c->n_v);
        launch_cell_to_vertex(c, nf, q, vol, num, den);
        GL_CUDA(cudaMemcpyAsync(vertex_out, num, sizeof(double) * c->n_v * nf, cudaMemcpyDeviceToHost, c->stream));
    }
    GL_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}

int glims_project_fields(glims_ctx* c, double* vertex_out) {
    API_BEGIN
    if (!c->have_mat) throw GlError(GLIMS_ERR_STATE, "glims_project_fields before glims_set_materials");
    if (!vertex_out) throw GlError(GLIMS_ERR_ARG, "glims_project_fields: null output");
    const double* M = cc_mass_matrix(c);
    if (!M) throw GlError(GLIMS_ERR_STATE, std::string("glims_project_fields: mass matrix unavailable: ") + cc_status(c));
    const int nf = 2 * c->dim * c->dim + 5;
    const i64 nr = nrows(c);
    double *q = ws(c, "pp_q", c->n_c * nf), *vol = ws(c, "pp_vol", c->n_c), *load = ws(c, "pp_load", c->n_v * nf);
    double *out = ws(c, "pp_proj", c->n_v * nf), *b = ws(c, "pp_b", c->n_v), *xq = ws(c, "pp_x", c->n_v);
    halo_exchange(c, c->x, c->nb);
    launch_cell_fields(c, q, vol);
    launch_project_load(c, q, vol, load);
    if (!c->dinv_mass) GL_CUDA(cudaMalloc(&c->dinv_mass, sizeof(double) * nr));
    launch_scalar_diag_inverse(c, M, c->dinv_mass);
    launch_zero(c, out, c->n_v * nf);
    for (int f = 0; f < nf; ++f) {
        launch_strided_copy(c, load, nr, nf, f, b, 1, 0);
        double res = 0;
        int its = pcg(c, 3, GLIMS_PC_JACOBI, b, xq, 1e-13, 0.0, 1e-300, 2000, &res);
        if (its < 0) throw GlError(GLIMS_ERR_NOT_CONVERGED, "glims_project_fields: mass-matrix CG did not converge");
        launch_strided_copy(c, xq, nr, 1, 0, out, nf, f);
    }
    GL_CUDA(cudaMemcpyAsync(vertex_out, out, sizeof(double) * c->n_v * nf, cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}

int glims_mass_solve(glims_ctx* c, int32_t nf, const double* load, double* out) {
    API_BEGIN
    if (!c->have_mat) throw GlError(GLIMS_ERR_STATE, "glims_mass_solve before glims_set_materials");
    if (nf <= 0 || !load || !out) throw GlError(GLIMS_ERR_ARG, "glims_mass_solve: bad arguments");
    const double* M = cc_mass_matrix(c);
    if (!M) throw GlError(GLIMS_ERR_STATE, std::string("glims_mass_solve: mass matrix unavailable: ") + cc_status(c));
    const i64 nr = nrows(c);
    double *dl = ws(c, "ms_load", c->n_v * nf), *dout = ws(c, "ms_out", c->n_v * nf), *b = ws(c, "pp_b", c->n_v), *xq = ws(c, "pp_x", c->n_v);
    GL_CUDA(cudaMemcpyAsync(dl, load, sizeof(double) * c->n_v * nf, cudaMemcpyHostToDevice, c->stream));
    if (!c->dinv_mass) GL_CUDA(cudaMalloc(&c->dinv_mass, sizeof(double) * nr));
    launch_scalar_diag_inverse(c, M, c->dinv_mass);
    launch_zero(c, dout, c->n_v * nf);
    for (int f = 0; f < nf; ++f) {
        launch_strided_copy(c, dl, nr, nf, f, b, 1, 0);
        double res = 0;
        int its = pcg(c, 3, GLIMS_PC_JACOBI, b, xq, 1e-13, 0.0, 1e-300, 2000, &res);
        if (its < 0) throw GlError(GLIMS_ERR_NOT_CONVERGED, "glims_mass_solve: mass-matrix CG did not converge");
        launch_strided_copy(c, xq, nr, 1, 0, dout, nf, f);
    }
    GL_CUDA(cudaMemcpyAsync(out, dout, sizeof(double) * c->n_v * nf, cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}

int glims_adjoint(glims_ctx* c, int32_t n_steps, const glims_solver_opts* o, int32_t n_levels, const double* levels,
                  const double* level_targets, const double* u_target, double* J_out, double* grad) {
    API_BEGIN
    if (!c->have_mat) throw GlError(GLIMS_ERR_STATE, "glims_adjoint before glims_set_materials");
    if (n_steps <= 0 || n_levels < 0 || (n_levels > 0 && (!levels || !level_targets)) || !J_out || !grad)
        throw GlError(GLIMS_ERR_ARG, "glims_adjoint: bad arguments");
    if (c->halo.active) throw GlError(GLIMS_ERR_STATE, "glims_adjoint: one GPU only (the adjoint sweep is not partitioned yet)");
    if (c->dof_perm) throw GlError(GLIMS_ERR_STATE, "glims_adjoint: not available with a caller dof permutation");
    glims_solver_opts od;
    if (!o) { glims_default_opts(&od); o = &od; }
    if (o->solver != GLIMS_SOLVER_BLOCK_TRI) throw GlError(GLIMS_ERR_ARG, "glims_adjoint: needs the block-triangular solver");
    if (!cc_available(c)) throw GlError(GLIMS_ERR_STATE, std::string("glims_adjoint: row-walk maps unavailable: ") + cc_status(c));
    const int D = c->dim, NB = c->nb;
    const i64 nv = c->n_v, nr = nrows(c), nd = c->ndof;
    // ---- forward sweep, trajectory on the device ------------------------------------------------------------------
    double* traj = ws(c, "adj_traj", (i64)(n_steps + 1) * nd);
    launch_copy(c, c->xprev, traj, nd);
    for (int s = 0; s < n_steps; ++s) {
        newton_step(c, o, nullptr);
        launch_copy(c, c->x, c->xprev, nd);
        launch_copy(c, c->x, traj + (i64)(s + 1) * nd, nd);
    }
    c->have_hist = false;
    // ---- misfit of the final state and -dJ/dx_N ---------------------------------------------------------------------
    double *xu = ws(c, "adj_xu", nv * D), *xc = ws(c, "adj_xc", nv), *ru = ws(c, "adj_ru", nv * D), *rc = ws(c, "adj_rc", nv);
    double *r = ws(c, "adj_r", nv), *dth = ws(c, "adj_dth", nv), *Mr = ws(c, "adj_Mr", nv), *tg = ws(c, "adj_tg", nv * std::max(D, 1));
    double *lu = ws(c, "adj_lu", nv * D), *lc = ws(c, "adj_lc", nv), *tmpc = ws(c, "adj_tmpc", nv);
    double* dgrad = ws(c, "adj_grad", MAX_MAT * 3);
    launch_extract(c, c->x, xu, xc);
    launch_zero(c, ru, nv * D);
    launch_zero(c, rc, nv);
    double J = 0.0;
    (void)cc_mass_matrix(c);         // builds the mass matrix on first use
    for (int l = 0; l < n_levels; ++l) {
        GL_CUDA(cudaMemcpyAsync(tg, level_targets + (i64)l * nv, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
        launch_threshold(c, xc, tg, levels[l], r, dth);
        launch_spmv(c, 3, r, Mr);
        launch_dot(c, r, Mr, nr, S_TMP0);
        double v; read_scalars(c, S_TMP0, 1, &v);
        J += v;
        launch_acc_prod(c, rc, 1, 0, -2.0, Mr, dth, nr);            // rhs_c = -dJ/dc
    }
    if (u_target) {
        GL_CUDA(cudaMemcpyAsync(tg, u_target, sizeof(double) * nv * D, cudaMemcpyHostToDevice, c->stream));
        for (int k = 0; k < D; ++k) {
            launch_diff_strided(c, xu, tg, D, k, nr, r);
            launch_spmv(c, 3, r, Mr);
            launch_dot(c, r, Mr, nr, S_TMP0);
            double v; read_scalars(c, S_TMP0, 1, &v);
            J += v;
            launch_acc_prod(c, ru, D, k, -2.0, Mr, nullptr, nr);    // rhs_u = -dJ/du
        }
    }
    *J_out = J;
    // ---- adjoint sweep ------------------------------------------------------------------------------------------------
    GL_CUDA(cudaMemsetAsync(dgrad, 0, sizeof(double) * MAX_MAT * 3, c->stream));
    ensure_kconst(c, o);
    for (int n = n_steps; n >= 1; --n) {
        // state of step n: K_cc(c_n), Dirichlet-eliminated like the forward solves
        launch_copy(c, traj + (i64)n * nd, c->x, nd);
        launch_cc_rows(c, true, false);
        launch_bc_matrix(c, GLIMS_ASM_KCC, true);
        launch_diag_inverse(c, 2);
        launch_zero_bc_split(c, ru, rc);
        // K_uu l_u = rhs_u
        double res = 0;
        launch_dot(c, ru, ru, nr * D, S_TMP0);
        double bu2; read_scalars(c, S_TMP0, 1, &bu2);
        if (bu2 > 0.0) {
            int its = pcg(c, 1, o->pc, ru, lu, o->ksp_rtol, 0.0, o->ksp_atol, o->max_krylov, &res, false);
            if (its < 0) throw GlError(GLIMS_ERR_NOT_CONVERGED, "glims_adjoint: K_uu solve did not converge");
        } else launch_zero(c, lu, nv * D);
        // K_cc l_c = rhs_c - K_uc^T l_u
        launch_spmv_uc_T(c, lu, tmpc);
        launch_axpy(c, -1.0, tmpc, rc, nr);
        launch_zero_bc_split(c, ru, rc);
        launch_dot(c, rc, rc, nr, S_TMP0);
        double bc2; read_scalars(c, S_TMP0, 1, &bc2);
        if (bc2 > 0.0) {
            int its = pcg(c, 2, GLIMS_PC_JACOBI, rc, lc, o->ksp_rtol, 0.0, o->ksp_atol, o->max_krylov, &res);
            if (its < 0) throw GlError(GLIMS_ERR_NOT_CONVERGED, "glims_adjoint: K_cc solve did not converge");
        } else launch_zero(c, lc, nv);
        // gradient contributions of step n
        launch_adjoint_grad(c, c->x, lu, lc, dgrad);
        // rhs of step n-1: -(dR_n/dx_{n-1})^T l_n = M l_c on the concentration rows, nothing on the displacement rows
        launch_spmv(c, 3, lc, rc);
        launch_zero(c, ru, nv * D);
    }
    std::vector<double> hg(MAX_MAT * 3);
    GL_CUDA(cudaMemcpyAsync(hg.data(), dgrad, sizeof(double) * MAX_MAT * 3, cudaMemcpyDeviceToHost, c->stream));
    // leave the context at the end of the forward run
    launch_copy(c, traj + (i64)n_steps * nd, c->x, nd);
    launch_copy(c, c->x, c->xprev, nd);
    GL_CUDA(cudaStreamSynchronize(c->stream));
    for (int m = 0; m < c->n_mat * 3; ++m) grad[m] = hg[m];
    c->fu_cache_valid = false;
    API_END
}

int glims_tile_config(glims_ctx* c, int32_t threads_per_cta, int32_t chunk) {
    API_BEGIN
    if (threads_per_cta != 0 && threads_per_cta != 128 && threads_per_cta != 192 && threads_per_cta != 256)
        throw GlError(GLIMS_ERR_ARG, "glims_tile_config: threads_per_cta must be 0, 128, 192 or 256");
    if (chunk < 0) throw GlError(GLIMS_ERR_ARG, "glims_tile_config: chunk must be >= 0");
    GL_CUDA(cudaStreamSynchronize(c->stream));
    tile_free(c);
    c->tile_nt = threads_per_cta; c->tile_chunk = chunk;
    API_END
}

int glims_tile_info(glims_ctx* c, int64_t* info8) {
    API_BEGIN
    long long tmp[8] = {0};
    const char* st = tile_status(c, tmp);
    if (std::string(st) != "ok") throw GlError(GLIMS_ERR_STATE, std::string("tile maps: ") + st);
    for (int i = 0; i < 8; ++i) info8[i] = tmp[i];
    API_END
}

int glims_time_kernel(glims_ctx* c, int32_t kernel, int32_t variant, int32_t reps, int32_t do_flush, float* ms_avg) {
    API_BEGIN
    if (!c->have_mat || reps <= 0 || !ms_avg) throw GlError(GLIMS_ERR_ARG, "glims_time_kernel: bad arguments/state");
    int bs = kernel == 1 ? c->nb : kernel == 2 ? c->dim : 1;
    double *dx = ws(c, "tk_x", c->ndof), *dy = ws(c, "tk_y", c->ndof);
    if (kernel >= 1 && kernel <= 3) launch_copy(c, c->x, dx, c->n_v * bs);
    auto run = [&]() {
        switch (kernel) {
            case 0: launch_assemble(c, GLIMS_ASM_ALL, variant); break;
            case 1: launch_spmv(c, 0, dx, dy); break;
            case 2: launch_spmv(c, 1, dx, dy); break;
            case 3: launch_spmv(c, 2, dx, dy); break;
            case 4: launch_assemble(c, GLIMS_ASM_RESIDUAL, variant); break;
            case 5: launch_assemble(c, GLIMS_ASM_RESIDUAL | GLIMS_ASM_KCC, variant); break;
            case 10: if (!amg_time_level1_step(c)) throw GlError(GLIMS_ERR_STATE, "glims_time_kernel 10: no three-level FP32 hierarchy"); break;
            case 9: if (!amg_time_coarse(c)) throw GlError(GLIMS_ERR_STATE, "glims_time_kernel 9: no three-level FP32 hierarchy"); break;
            case 7: launch_cc_rows(c, true, true); break;
            case 8: launch_fu(c, false); break;
            case 6: if (!amg_time_fine_step(c)) throw GlError(GLIMS_ERR_STATE, "glims_time_kernel 6: no AMG hierarchy yet (run a step with pc = GLIMS_PC_AMG first)"); break;
            default: throw GlError(GLIMS_ERR_ARG, "glims_time_kernel: unknown kernel");
        }
    };
    if (kernel == 7) { if (!cc_available(c)) throw GlError(GLIMS_ERR_STATE, std::string("glims_time_kernel 7: ") + cc_status(c)); cc_mass_cprev(c); }
    if (kernel == 8 && c->kuu_state == 0) launch_assemble(c, GLIMS_ASM_KCONST, GLIMS_ASMK_TILE);
    for (int w = 0; w < 3; ++w) run();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float tot = 0;
    for (int r = 0; r < reps; ++r) {
        if (do_flush) flush_l2(c);
        cudaEventRecord(a, c->stream);
        run();
        cudaEventRecord(b, c->stream);
        GL_CUDA(cudaEventSynchronize(b));
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        tot += ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_avg = tot / reps;
    if (kernel == 0 || kernel == 8) c->kconst_valid = false;   // raw matrices: no Dirichlet elimination applied
    c->fu_cache_valid = false;
    API_END
}

}  // extern "C"
