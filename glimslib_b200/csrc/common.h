// Internal declarations shared by the translation units of libglims_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/glims_b200.h"

typedef long long i64;

#define GL_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            char b_[512];                                                               \
            snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call,          \
                     cudaGetErrorString(e_));                                           \
            throw GlError(GLIMS_ERR_CUDA, b_);                                          \
        }                                                                               \
    } while (0)

struct GlError {
    int code;
    std::string msg;
    GlError(int c, const std::string& m) : code(c), msg(m) {}
};

// Grid of a grid-stride kernel = what the device holds at once (resident CTAs per SM x SMs), from the occupancy calculator and
// cached per kernel.  A fixed 8-per-SM grid runs a second, thin wave for every kernel that needs more than 32 registers
// (6 resident CTAs of 256 threads at 40): the fine-level smoother step lost 16 % to that tail.
inline int resident_grid(const void* kernel, int threads, size_t dyn_smem = 0) {
    static std::mutex mu;
    static std::unordered_map<const void*, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(kernel);
    if (it != cache.end()) return it->second;
    int per_sm = 0, dev = 0, sms = 148;
    GL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem));
    GL_CUDA(cudaGetDevice(&dev));
    GL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int g = (per_sm > 0 ? per_sm : 1) * sms;
    cache[kernel] = g;
    return g;
}
template <typename K>
inline int fit_grid(K kernel, i64 tiles, int threads, int cap) {
    i64 g = resident_grid((const void*)kernel, threads);
    if (g > cap) g = cap;
    if (g > tiles) g = tiles;
    return (int)(g > 0 ? g : 1);
}

constexpr int SLICE = 32;          // SELL-32: one warp lane per block row
constexpr int MAX_MAT = 64;        // material table rows staged in shared memory
constexpr int MAT_STRIDE = 6;      // mu, lambda, D, rho, gamma, beta=(2mu+d*lambda)*gamma

// Sliced-ELL block pattern of one vertex graph. Slot s of row r (block j of the row):
//   s = slice_off[r/32] + j*32 + r%32 ; value k of an ncomp-block: (s & ~31)*ncomp + k*32 + (s & 31).
// Padding slots carry col = own row and zero values.
struct SellPattern {
    int n_rows = 0;
    int n_slices = 0;
    i64 n_slots = 0;       // padded
    i64 nnzb = 0;          // true blocks
    i64* slice_off = nullptr;   // [n_slices+1]
    int* slice_w = nullptr;     // [n_slices]
    int* col = nullptr;         // [n_slots]
    i64* rowptr = nullptr;      // [n_rows+1] CSR (true entries, columns ascending)
    int* diag = nullptr;        // [n_rows] slot of the diagonal block
    int max_w = 0;
};

__host__ __device__ inline i64 vidx(i64 s, int k, int ncomp) {
    return (s & ~(i64)31) * ncomp + (i64)k * 32 + (s & 31);
}

struct AmgLevel;   // amg.cu
struct Amg;

struct Halo {       // multi-GPU ghost exchange plan (device copies)
    bool active = false;
    int n_ranks = 1, rank = 0;
    void* comm = nullptr;          // ncclComm_t
    i64 n_owned = 0;
    std::vector<int> peers;
    std::vector<i64> send_ptr, recv_ptr;
    int* send_idx = nullptr;       // device
    double* send_buf = nullptr;    // device, send_ptr.back() * nb doubles (max)
    i64 n_send = 0;
    void* p2p = nullptr;           // comm.cu: peer-memory windows (NVLink P2P halo exchange / allreduce), null = NCCL path
    bool p2p_enabled = true;       // runtime switch (glims_set_p2p); GLIMS_NO_P2P=1 disables at setup
};

struct glims_ctx {
    int dim = 0, nb = 0, device = 0;
    i64 n_v = 0, n_c = 0, ndof = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    i64 launches = 0;

    // mesh
    double* coords = nullptr;   // [n_v][dim]
    int* cells = nullptr;       // [n_c][nb]
    int* cell_mat = nullptr;    // [n_c]
    int n_mat = 0;
    double* mat = nullptr;      // [n_mat][MAT_STRIDE] device
    double dt = 1.0;
    bool have_mat = false, kconst_valid = false;

    // pattern + scatter maps
    SellPattern pat;
    int* eslot = nullptr;       // [n_c][nb*nb] slot per (a,b) pair: element -> matrix scatter map
    // transposed map for the gather kernel: contributors of every slot, CSR over slots
    i64* gptr = nullptr;        // [n_slots+1]
    int* gent = nullptr;        // [n_c*nb*nb] packed: element<<4 | a<<2 | b
    bool have_gather = false;
    // slice-local assembly: per SELL slice the elements touching its rows + contributor entries against that list
    i64* sl_ptr = nullptr;      // [n_slices+1]
    int* sl_elem = nullptr;     // [sl_total]
    unsigned short* lent = nullptr;   // [#contributors] (local element index << 4) | a << 2 | b
    int sl_max = 0;
    i64 sl_total = 0;
    bool have_slice = false;
    int tile_nt = 0, tile_chunk = 0;   // 0: GLIMS_TILE_NT / GLIMS_TILE_CH or the defaults (glims_tile_config)
    void* tile = nullptr;       // tile.cu: TileDev (maps of the fused tile-assembly kernel), built on first use
    void* ccmap = nullptr;      // ccrow.cu: CcMap ((row, element) pair lists + per-slot constants of the row-walk assembly)
    int kuu_state = 0;          // K_uu / K_uc contents: 0 not assembled for the current materials, 1 raw, 2 Dirichlet-eliminated
    bool bc_nonzero = false;    // some Dirichlet value is non-zero (F_u by SpMV then needs the lift of the eliminated columns)
    bool lift_dirty = false;    // Dirichlet values changed on an unchanged dof set: K_uu/K_uc must be re-eliminated for the lift
    bool fu_cache_valid = false;// fu2_cached is |F_u|^2 of the current (x, c, BCs, load): reused as the first residual of the next step
    double fu2_cached = 0.0;
    i64* dof_perm = nullptr;            // caller dof numbering -> vertex-blocked (device; null = identity)
    std::vector<int64_t> h_dof_perm;
    std::vector<long long> h_bc_dofs;   // host copy of the Dirichlet dof set (to tell a value update from a set change)

    // matrices (SELL value layout, see vidx)
    double *Kuu = nullptr, *Kuc = nullptr, *Kcc = nullptr;
    // preconditioner data
    double* dinv_uu = nullptr;  // [n_v][dim*dim] inverse diagonal blocks of K_uu
    double* dinv_cc = nullptr;  // [n_v]
    double* dinv_mono = nullptr;// [n_v][nb*nb]
    double* dinv_mass = nullptr;// [n_v] inverse diagonal of the P1 mass matrix (L2 projections)

    // vectors, vertex-blocked [n_v][nb]
    double *x = nullptr, *xprev = nullptr, *F = nullptr, *fext = nullptr, *dx = nullptr;
    bool have_load = false;
    // Dirichlet
    i64 n_bc = 0;
    i64* bc_dofs = nullptr;
    double* bc_vals = nullptr;
    i64 n_bc_u = 0, n_bc_c = 0;
    unsigned char* bcmask = nullptr;   // [n_v] bit k set: dof (v,k) constrained

    // Krylov workspace (allocated lazily)
    double* work = nullptr;
    i64 work_size = 0;
    double* scal = nullptr;         // device scalars
    double* partials = nullptr;     // block partial sums
    unsigned* tickets = nullptr;
    double* h_scal = nullptr;       // pinned host mirror
    double* h_ring = nullptr;       // pinned mirror of the PCG r.r ring
    bool use_graphs = true;         // replay PCG iterations as CUDA graphs (GLIMS_NO_GRAPH=1 disables)
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;

    Amg* amg = nullptr;
    Halo halo;
    bool first_step_done = false;
    // successive-right-hand-side projection for the constant K_uu (solver.cu: pcg)
    int rec_n = 0, rec_head = 0;
    void* solver_state = nullptr;   // solver.cu: work-vector pool, projection history, graph cache
    bool have_hist = false;     // x_hist holds the solution two steps back (extrapolated Newton start)
};

// ---------------- pattern.cu
void build_pattern(glims_ctx* c);
void build_gather_map(glims_ctx* c);
void build_slice_map(glims_ctx* c);
// SELL pattern from unsorted 64-bit (row<<32|col) keys on the device; ~0 keys are dropped. keys is consumed.
// ukeys_out (optional) receives the sorted unique keys (device, caller frees), in CSR order.
void build_pattern_from_keys(cudaStream_t st, unsigned long long* keys, i64 n_keys, i64 n_rows, SellPattern& P,
                             unsigned long long** ukeys_out);
void free_pattern(SellPattern& P);

// ---------------- tile.cu
bool launch_assemble_tile(glims_ctx* c, int what);   // false: maps cannot represent this mesh (caller falls back)
void tile_free(glims_ctx* c);
const char* tile_status(glims_ctx* c, long long* info8);

// ---------------- ccrow.cu (row-walk assembly, GLIMS_ASMK_ROWS)
bool cc_available(glims_ctx* c);                     // builds the pair lists on first use; false: mesh not representable
const char* cc_status(glims_ctx* c);
void cc_free(glims_ctx* c);
void cc_invalidate_consts(glims_ctx* c);             // materials or dt changed
void cc_mass_cprev(glims_ctx* c);                    // M c_prev of the current u_previous
bool launch_cc_rows(glims_ctx* c, bool with_kcc, bool with_res);   // K_cc and/or F_c from the current state
void launch_fu(glims_ctx* c, bool eliminated);       // F_u = K_uu u + K_uc c (+ lift) - f_ext from the stored blocks
void cc_compute_lift(glims_ctx* c, bool any_nonzero);// while K_uu / K_uc are raw
i64 cc_map_bytes(glims_ctx* c);
const double* cc_mass_matrix(glims_ctx* c);          // P1 mass matrix over the vertex pattern (scalar SELL values)

// ---------------- kernels.cu (launch wrappers; all on c->stream)
void launch_assemble(glims_ctx* c, int what, int variant);
void launch_bc_values(glims_ctx* c, double* x);
void launch_bc_residual(glims_ctx* c, double* F, const double* x);
void launch_bc_matrix(glims_ctx* c, int what, bool sym);
void launch_diag_inverse(glims_ctx* c, int which);   // 0 mono, 1 uu, 2 cc
void launch_export_values(glims_ctx* c, double* Kuu, double* Kuc, double* Kcc);  // device CSR-order outputs

struct SpmvDot {            // optional fused dot: out = sum_r y[r] . w[r]
    const double* w = nullptr;
    int slot = -1;          // scalar slot index in c->scal
};
// y = A x for block sizes: which 0 mono (nb x nb over K_uu,K_uc,K_cc), 1 K_uu (dim x dim), 2 K_cc (1x1)
void launch_spmv(glims_ctx* c, int which, const double* x, double* y, SpmvDot dot = SpmvDot());
// y_u[n_v][dim] = K_uc x_c
void launch_spmv_uc(glims_ctx* c, const double* xc, double* yu);

// generic SELL block SpMV used by AMG levels: y = A x, A blocks BRxBC stored [slot][BR*BC]
void launch_spmv_generic(glims_ctx* c, const SellPattern& p, const double* A, int bs, const double* x, double* y,
                         const double* rhs = nullptr);   // rhs: y = rhs - A x

// vector kernels with device-resident scalars; scal indices refer to c->scal
enum { S_RZ = 0, S_PAP, S_RR, S_RZNEW, S_BN, S_TMP0, S_TMP1, S_TMP2, S_TMP3, S_GM0 /* 64 slots from here */, S_COUNT = 128 };
void launch_dot(glims_ctx* c, const double* a, const double* b, i64 n, int slot);
void launch_multi_dot(glims_ctx* c, const double* V, i64 ld, int k, const double* w, i64 n, int slot0);
void launch_axpy(glims_ctx* c, double alpha, const double* x, double* y, i64 n);               // y += alpha x
void launch_scale(glims_ctx* c, double alpha, double* x, i64 n);
void launch_copy(glims_ctx* c, const double* x, double* y, i64 n);
void launch_zero(glims_ctx* c, double* x, i64 n);
// CG fused updates (alpha = scal[num]/scal[den] computed on device)
void launch_cg_update_xr(glims_ctx* c, double* x, double* r, const double* p, const double* Ap, i64 n,
                         int s_num, int s_den, int s_rr);
void launch_cg_update_p(glims_ctx* c, double* p, const double* z, i64 n, int s_num, int s_den);
// z = Dinv r (block size bs), fused dot(r,z) -> slot
void launch_block_jacobi(glims_ctx* c, const double* dinv, int bs, const double* r, double* z, i64 n_rows, int s_rz);
// blocked <-> split
void launch_extract(glims_ctx* c, const double* xb, double* xu, double* xc);   // either may be null
void launch_insert_add(glims_ctx* c, double* xb, const double* du, const double* dc, double alpha);
void launch_split_norms(glims_ctx* c, const double* F, int s0);   // |F_u|^2, |F_c|^2 -> slots s0, s0+1
void launch_multi_axpy(glims_ctx* c, const double* V, i64 ld, int k, const double* coef_dev, double sign, double* w, i64 n);
void read_scalars(glims_ctx* c, int slot0, int n, double* out);   // sync
void launch_project_load(glims_ctx* c, const double* q, const double* vol, double* load);   // load[n_v][nf] = int f phi
void launch_strided_copy(glims_ctx* c, const double* src, i64 n, int ss, int os, double* dst, int sd, int od);
void launch_scalar_diag_inverse(glims_ctx* c, const double* A, double* out);
void flush_l2(glims_ctx* c);
void launch_permute(glims_ctx* c, const double* src, double* dst, const i64* perm, i64 n, bool scatter);  // scatter: dst[perm[i]] = src[i]

// ---------------- adjoint.cu (kernels of the discrete adjoint; the driver is glims_adjoint in solver.cu)
void launch_spmv_uc_T(glims_ctx* c, const double* lu, double* out);                       // out[n_v] = K_uc^T l_u
void launch_adjoint_grad(glims_ctx* c, const double* x, const double* lu, const double* lc, double* grad);   // grad[n_mat][3] +=
void launch_threshold(glims_ctx* c, const double* cvec, const double* tgt, double level, double* r, double* dth);
void launch_acc_prod(glims_ctx* c, double* g, int sg, int og, double s, const double* a, const double* b, i64 n);
void launch_diff_strided(glims_ctx* c, const double* a, const double* b, int stride, int off, i64 n, double* r);
void launch_zero_bc_split(glims_ctx* c, double* ru, double* rc);

// ---------------- amg.cu
void amg_setup(glims_ctx* c);
void amg_free(glims_ctx* c);
void amg_vcycle(glims_ctx* c, const double* r, double* z, bool fp32);
void amg_check(glims_ctx* c);              // throws if a device-side wait inside the V-cycle timed out
bool amg_time_coarse(glims_ctx* c);        // false: hierarchy has fewer than three levels
bool amg_time_level1_step(glims_ctx* c);   // false: hierarchy has fewer than three levels
bool amg_time_fine_step(glims_ctx* c);     // false: no FP32 hierarchy yet   // z = M^-1 r on K_uu ([n_v][dim] vectors)

// ---------------- comm.cu
void halo_exchange(glims_ctx* c, double* xb, int bs);        // fill ghost values of a blocked vector
void halo_exchange_f32(glims_ctx* c, float* xb, int bs);
void allreduce_scalars(glims_ctx* c, int slot0, int n);      // in-place sum over ranks of c->scal slots
void comm_free(glims_ctx* c);                                // peer windows (NCCL communicator is left to process exit)
void comm_check(glims_ctx* c);                               // throws if a peer-memory wait timed out
void solver_free_graphs(glims_ctx* c);                       // solver.cu: drop captured PCG graphs (transport changed)
// device-visible description of a symmetric buffer: base address of every rank's copy (own one included), indexed by rank.
// Layout of a copy: [0,64) u64 flag[rank] | [64,72) u64 seq | [72,76) u32 pushed | [76,80) i32 error | [80,84) u32 left | data from 256
struct SymView { int n_ranks, rank; unsigned char* base[8]; };
constexpr size_t SYM_DATA_OFFSET = 256;
// symmetric buffers: one allocation per rank mapped by every rank (collective alloc); allgather_f32 gathers, in place, a
// float vector [n_ranks][seg] living inside such a buffer (peer-memory pushes, or ncclAllGather when peer memory is off)
void* sym_alloc(glims_ctx* c, size_t bytes);
void sym_free(void* sb);
void* sym_data(void* sb);
bool sym_check(void* sb);
bool sym_is_p2p(void* sb);                    // every rank maps every other rank's copy (peer-memory path in use)
const SymView* sym_view(void* sb);          // host copy of the view (valid while the buffer lives)
void allgather_f32(glims_ctx* c, void* sb, float* buf, i64 seg);
// setup-time collectives (synchronous)
void comm_allreduce_max_i64(glims_ctx* c, long long* v, int n);
void comm_allreduce_max_f64(glims_ctx* c, double* v, int n);
void comm_allgatherv(glims_ctx* c, const void* src_dev, void* dst_dev, const std::vector<long long>& counts, size_t elem_bytes);
void comm_allgather_i64(glims_ctx* c, long long mine, std::vector<long long>& all);
