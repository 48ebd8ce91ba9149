/* glims_b200.h -- C ABI of the B200-native GlimSLib forward-simulation hot path.
 *
 * The reference (danielabler/glimslib) has no FFI of its own: every hot-path
 * module reaches its backend through one Python seam, glimslib/fenics_local.py:3-10
 * (`from dolfin import *`).  The functions below are what a ctypes binding in that
 * seam calls instead of DOLFIN/PETSc; each cites the reference call it replaces.
 * INTEGRATION.md shows the binding.
 *
 * Conventions: every function returns 0 on success and a negative glims_status on
 * failure (message via glims_last_error); no exception crosses the ABI.  All
 * pointers are HOST pointers to C-contiguous caller-owned arrays unless the name
 * says `_dev`; the library owns all device memory.  One host thread per context.
 *
 * Unknown layout ("vertex-blocked"): dof(v, k) = v*(dim+1) + k, k < dim the
 * displacement components (sub-space 0), k == dim the concentration (sub-space 1)
 * -- the mixed element of simulation_tumor_growth.py:67-72.
 */
#ifndef GLIMS_B200_H
#define GLIMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct glims_ctx glims_ctx;

typedef enum {
    GLIMS_OK = 0,
    GLIMS_ERR_ARG = -1,          /* bad argument */
    GLIMS_ERR_CUDA = -2,         /* CUDA runtime error (see glims_last_error) */
    GLIMS_ERR_NOT_CONVERGED = -3,/* Newton or Krylov did not converge: the Python shim raises, so that
                                    simulation_base.py:301-305's `except:` path keeps working */
    GLIMS_ERR_STATE = -4,        /* call order violated (e.g. step before set_materials) */
    GLIMS_ERR_NCCL = -5
} glims_status;

/* what to assemble (bit mask) */
enum { GLIMS_ASM_RESIDUAL = 1, GLIMS_ASM_KCONST = 2 /* K_uu, K_uc */, GLIMS_ASM_KCC = 4,
       GLIMS_ASM_JACOBIAN = 6, GLIMS_ASM_ALL = 7 };

/* linear-solver strategy for the Newton update  J dx = -F */
enum { GLIMS_SOLVER_BLOCK_TRI = 0,   /* exact block back-substitution of J = [[K_uu,K_uc],[0,K_cc]]: PCG on each block */
       GLIMS_SOLVER_MONO_GMRES = 1   /* GMRES(30) on the monolithic J, block-Jacobi left PC (PETSc-default-like) */ };
enum { GLIMS_PC_JACOBI = 0,          /* (block-)Jacobi on every block */
       GLIMS_PC_AMG = 1,             /* aggregation AMG V-cycle on K_uu (FP32 storage inside the V-cycle,
                                        FP64 outer PCG), Jacobi on K_cc */
       GLIMS_PC_AMG_FP64 = 2         /* same hierarchy, V-cycle entirely in FP64 */ };
/* assembly kernel variant for the Jacobian */
enum { GLIMS_ASMK_ATOMIC = 0,        /* element-parallel, scatter map + RED.ADD.F64 */
       GLIMS_ASMK_GATHER = 1,        /* row-parallel gather through the transposed scatter map: no atomics, deterministic */
       GLIMS_ASMK_SLICE = 2,         /* gather with the element geometry of each 32-row slice staged in shared memory */
       GLIMS_ASMK_TILE = 3,          /* fused residual+Jacobian, one CTA per 16-row tile: vertices, element gradients and
                                        contributor lists staged in shared memory, material-free raw sums per slot, residual
                                        from the same sums; no atomics, deterministic (csrc/tile.h, tile.cu) */
       GLIMS_ASMK_ROWS = 4           /* default of glims_step: the per-Newton-iteration pass without atomics or scatter map.
                                        K_uu/K_uc (state independent, stg:110-114) come from the tile kernel once; per iteration
                                        one row-walk kernel forms K_cc and F_c from per-slot constants + (row, element) pair lists,
                                        and F_u = K_uu u + K_uc c is one SpMV over the stored blocks (csrc/ccrow.cu) */ };

typedef struct {
    /* SNES-like controls; defaults mirror DOLFIN's PETScSNESSolver defaults that
       simulation_tumor_growth.py:126-130 leaves untouched */
    double snes_rtol;        /* 1e-9  */
    double snes_atol;        /* 1e-10 */
    double snes_stol;        /* 1e-16; accepted for parity with DOLFIN's parameter set, not used as a stopping test */
    int32_t max_newton;      /* 50    */
    double ksp_rtol;         /* relative to |rhs| of each linear solve */
    double ksp_atol;         /* absolute floor for the linear residual */
    int32_t max_krylov;      /* per linear solve */
    int32_t solver;          /* GLIMS_SOLVER_* */
    int32_t pc;              /* GLIMS_PC_* */
    int32_t asm_kernel;      /* GLIMS_ASMK_* */
    int32_t lag_mechanics;   /* 1: skip the K_uu solve until the concentration block has converged
                                (same fixed point; the displacement does not feed back, stg:110-120) */
    int32_t recycle;         /* 1: project every K_uu solve onto the A-orthonormalised corrections of the previous
                                (up to 8) solves before PCG starts -- K_uu is constant, the loads vary smoothly in time */
    int32_t extrapolate;     /* 1: first Newton guess of a step = c_n + (c_n - c_{n-1}) once two solutions exist
                                (the reference starts from c_n; default 0: measured at C4 it lowers the final |F| but not the
                                Newton iteration count) */
} glims_solver_opts;

typedef struct {
    int32_t newton_its;
    int32_t krylov_its_c;    /* summed over the step's Newton iterations */
    int32_t krylov_its_u;
    int32_t krylov_its_mono;
    int32_t converged;       /* 1 / 0 */
    double fnorm0;           /* |F| at the first Newton iterate of the step */
    double fnorm;            /* |F| at exit (monolithic residual with Dirichlet rows) */
    float ms_total;          /* device time of the step, CUDA events on the context stream */
    float ms_assembly;
    float ms_krylov;
} glims_step_stats;

void glims_default_opts(glims_solver_opts* o);

/* ---- problem definition ------------------------------------------------------------------ */

/* Mesh + cell labels -> device, sparsity pattern, element->slot scatter map.
   Replaces fenics.FunctionSpace(mesh, MixedElement) + DOLFIN SparsityPatternBuilder
   (helper_classes.py:271-282) and SubDomains' cell MeshFunction (helper_classes.py:431-444).
   cells: n_cells x (dim+1) vertex ids; cell_mat: compact material index per cell.
   n_owned: -1 (or n_vertices) on one GPU; on a partitioned mesh the first n_owned local vertices are
   owned by this rank, the rest are ghosts, and cells must be every cell touching an owned vertex. */
int glims_create(glims_ctx** out, int32_t dim, int64_t n_vertices, const double* coords,
                 int64_t n_cells, const int32_t* cells, const int32_t* cell_mat, int64_t n_owned,
                 int32_t device);
int glims_destroy(glims_ctx* c);
const char* glims_last_error(const glims_ctx* c);

/* Per-label coefficients, table[m] = {mu, lambda, D, rho, gamma}: replaces the Python
   DiscontinuousScalar.eval_cell callbacks (helper_classes.py:47-58, 564-603) and
   compute_mu/compute_lambda (math_linear_elasticity.py:6-10). Invalidates cached K_uu/K_uc. */
int glims_set_materials(glims_ctx* c, int32_t n_mat, const double* table);
int glims_set_dt(glims_ctx* c, double dt);                       /* stg:108 */
/* fenics.DirichletBC list (helper_classes.py:673-723): dof indices (vertex-blocked) + values */
int glims_set_dirichlet(glims_ctx* c, int64_t n, const int64_t* dofs, const double* vals);
/* body force, RD source and von-Neumann facet terms (stg:112-113,119-120; helper_classes.py:861-908)
   pre-integrated into one load vector f_ext[ndof]; F = F_int(x) - f_ext.  NULL clears it. */
int glims_set_load(glims_ctx* c, const double* f_ext);

/* Dof numbering of the caller.  perm[i] = vertex-blocked index (see top) of the caller's dof i, e.g. built from
   DOLFIN's vertex_to_dof_map(V) / dofmap().dofs() (helper_classes.py:271-282, data_io.py:242-252) so that
   `sim.solution.vector()` arrays travel unchanged.  Applies to every ndof-sized HOST vector of this API (state, prev,
   load, residual) and to the dof indices of glims_set_dirichlet; NULL restores the identity.  Must be called before
   glims_set_dirichlet.  The reference's own numbering comes from DOLFIN/SCOTCH and cannot be generated offline
   (DESIGN.md section 3): this entry point is where it is plugged in. */
int glims_set_dof_permutation(glims_ctx* c, const int64_t* perm);
int glims_get_dof_permutation(glims_ctx* c, int64_t* perm);          /* ndof entries */

/* ---- state ------------------------------------------------------------------------------- */
int glims_set_state(glims_ctx* c, const double* x);      /* current Newton iterate / solution  */
int glims_get_state(glims_ctx* c, double* x);
int glims_set_prev(glims_ctx* c, const double* x_prev);  /* u_previous (simulation_base.py:253,312) */
int glims_get_prev(glims_ctx* c, double* x_prev);
int64_t glims_ndof(const glims_ctx* c);
int64_t glims_nnzb(const glims_ctx* c);                  /* vertex-graph blocks (true, unpadded) */
int64_t glims_nslots(const glims_ctx* c);                /* SELL-32 padded slots */
void* glims_state_dev(glims_ctx* c);                     /* device pointer of the iterate (torch interop) */
void* glims_stream(glims_ctx* c);                        /* cudaStream_t the library launches on */

/* ---- the hot path ------------------------------------------------------------------------ */

/* n_steps backward-Euler steps, each one `self.solver.solve()` (simulation_base.py:302) followed by
   `u_previous.assign(solution)` (:312).  stats may be NULL, else n_steps entries.
   Returns GLIMS_ERR_NOT_CONVERGED at the first failing step (state left at the last good step). */
int glims_step(glims_ctx* c, int32_t n_steps, const glims_solver_opts* o, glims_step_stats* stats);

/* One-time work of the first step, callable on its own so that it can be timed (bench.py `setup_s`): K_uu / K_uc assembly
   and Dirichlet elimination (DOLFIN assembles these inside every solve, stg:124), the AMG hierarchy, the row-walk maps.
   o == NULL: defaults. */
int glims_prepare(glims_ctx* c, const glims_solver_opts* o);
/* Forget what the solver learnt from earlier steps (successive-right-hand-side projection basis, extrapolation history,
   cached residual norm) without touching matrices, hierarchy or captured graphs: a run restarted from the same initial
   state then repeats the same iteration counts. */
int glims_reset_history(glims_ctx* c);

/* ---- building blocks (parity tests, roofline benches) -------------------------------------- */

/* Assemble on the device from the current state / prev; raw = before Dirichlet elimination.
   `what` is a GLIMS_ASM_* mask. apply_bc: 0 none, 1 rows only (DOLFIN DirichletBC::apply), 2 rows+cols. */
int glims_assemble(glims_ctx* c, int32_t what, int32_t kernel, int32_t apply_bc);
int glims_get_residual(glims_ctx* c, double* F);          /* ndof */
/* Block pattern in CSR order (rowptr n_v+1, colidx nnzb) and the three value arrays gathered to that
   order: Kuu[nnzb][dim][dim], Kuc[nnzb][dim], Kcc[nnzb] (K_cu is structurally zero, stg:115-120). */
int glims_export_pattern(glims_ctx* c, int64_t* rowptr, int32_t* colidx);
int glims_export_values(glims_ctx* c, double* Kuu, double* Kuc, double* Kcc);
/* y = J x with the assembled matrices; which: 0 monolithic (ndof), 1 K_uu (n_v*dim), 2 K_cc (n_v) */
int glims_spmv(glims_ctx* c, int32_t which, const double* x, double* y);
/* Average device time (ms, CUDA events on the context stream) of `reps` back-to-back launches of
   one kernel on resident data: kernel 0 = full assembly (residual+Jacobian, `variant` = GLIMS_ASMK_*),
   1 = monolithic SpMV, 2 = K_uu SpMV, 3 = K_cc SpMV, 4 = residual only, 5 = residual + K_cc (the per-Newton-iteration
   pass of the block-triangular solver), 6 = one fused Chebyshev smoother step of the V-cycle's fine level (FP16 matrix,
   FP32 vectors; needs a step with GLIMS_PC_AMG first), 7 = the row-walk K_cc + F_c kernel alone (what glims_step launches
   per Newton iteration), 8 = F_u = K_uu u + K_uc c by SpMV alone, 9 = the coarse-grid correction below level 1 of the
   V-cycle (levels >= 2: one persistent kernel, or its launch sequence with GLIMS_AMG_FUSED=0), 10 = one fused smoother step
   on level 1 of the V-cycle. flush_l2 != 0 writes a
   >L2-sized buffer between launches (outside the timed events). */
int glims_time_kernel(glims_ctx* c, int32_t kernel, int32_t variant, int32_t reps, int32_t flush_l2,
                      float* ms_avg);
/* Derived fields of the current state (PostProcessTumorGrowth, helper_classes.py:1560-1618,1736-1786), evaluated per
   cell on the device instead of by L2 projections: for every cell nf = 2*dim*dim + 5 values
   [strain (dim*dim), stress (dim*dim), pressure = tr(sigma)/3, von Mises, det(I + grad u), det(I + cbar*gamma*I),
   rho*cbar*(1-cbar)].  cell_out[n_cells][nf] and/or vertex_out[n_vertices][nf] (volume-weighted nodal average);
   either may be NULL. */
int glims_cell_fields(glims_ctx* c, double* cell_out, double* vertex_out);
/* The same fields L2-projected onto P1 with the CONSISTENT mass matrix -- what the reference's `project(expr, V)` returns
   (helper_classes.py:1560-1618, fenics_local project: M q = int f phi, here by Jacobi-PCG on the device to 1e-13).  Fields
   constant per cell are integrated against the hat functions exactly; det(I + c gamma I) and rho c (1 - c) are polynomials
   of the P1 concentration and are integrated exactly as well.  vertex_out[n_vertices][nf], vertex order of glims_create. */
int glims_project_fields(glims_ctx* c, double* vertex_out);
/* Generic consistent-mass L2 solve: out[:, f] = M^-1 load[:, f] for nf right-hand sides load[n_vertices][nf] (host), M the P1
   mass matrix of the mesh -- the linear solve inside every `fenics.project(expr, V)` of the reference's post-processing
   (helper_classes.py:1566-1618); the caller integrates `expr` against the hat functions. */
int glims_mass_solve(glims_ctx* c, int32_t nf, const double* load, double* out);
/* Discrete adjoint of the time loop (the reference's production caller: image_based_optimization.py:660-767 differentiates
   a misfit of the final state with dolfin-adjoint w.r.t. the controls of run_for_adjoint,
   simulation_tumor_growth_brain.py:127-145).  Runs n_steps forward steps from the current (prev, state) keeping the
   trajectory on the device, evaluates
       J = sum_l (th_l(c_N) - t_l)^T M (th_l(c_N) - t_l) + (u_N - u_t)^T (M x I) (u_N - u_t),
   th_l(c) = 0.5 (tanh((c - level_l)/0.01) + 1)  (:1404-1407), M the P1 mass matrix, and returns J and
   grad[n_mat][3] = dJ/d(D_m), dJ/d(rho_m), dJ/d(gamma_m) for every material row by one backward sweep of transposed block
   solves (K_uu with the forward AMG hierarchy, then K_cc).  level_targets[n_levels][n_vertices], u_target[n_vertices][dim]
   (NULL: no displacement term).  One GPU, block-triangular solver. */
int glims_adjoint(glims_ctx* c, int32_t n_steps, const glims_solver_opts* o, int32_t n_levels, const double* levels,
                  const double* level_targets, const double* u_target, double* J_out, double* grad);
/* Tile-assembly map statistics (after the first GLIMS_ASMK_TILE assembly): info[0..7] = max local vertices, max element
   records, max contributor entries, max items, max partial buffers per slice, shared-memory bytes per CTA, device bytes of
   the maps, threads per CTA.  Returns GLIMS_ERR_STATE when the maps were not built or cannot represent the mesh. */
int glims_tile_info(glims_ctx* c, int64_t* info8);
/* Tuning knobs of the tile kernel: threads per CTA (128, 192 or 256; 0 = default 192 / env GLIMS_TILE_NT) and the longest
   contributor chunk one warp handles before a column is split (0 = default 12 / env GLIMS_TILE_CH).  Drops the maps;
   they are rebuilt by the next GLIMS_ASMK_TILE assembly. */
int glims_tile_config(glims_ctx* c, int32_t threads_per_cta, int32_t chunk);
/* number of kernels this context has launched so far */
int64_t glims_launch_count(const glims_ctx* c);

/* ---- multi-GPU (one context per rank) ------------------------------------------------------ */

/* Create the rank's NCCL communicator (unique_id: 128 bytes from glims_nccl_unique_id on rank 0) and, when the halo
   plan is already set (call glims_set_halo first), map every rank's peer-memory window. */
int glims_nccl_unique_id(void* id128);
int glims_comm_init(glims_ctx* c, int32_t n_ranks, int32_t rank, const void* id128);
/* Halo plan: this rank's local vertices are [owned | ghost] (see glims_create).  For each peer p,
   send_idx[send_ptr[p]..send_ptr[p+1]) lists local owned vertices whose values p needs, and the ghosts
   received from p are the contiguous range [recv_ptr[p], recv_ptr[p+1]) of the ghost block. */
int glims_set_halo(glims_ctx* c, int32_t n_peers, const int32_t* peers, const int64_t* send_ptr,
                   const int32_t* send_idx, const int64_t* recv_ptr);

/* Transport of the halo exchange and of the scalar allreduce: 1 = peer-memory windows over NVLink (CUDA IPC; default
   whenever every rank can map every other rank's window, csrc/comm.cu), 0 = NCCL send/recv + ncclAllReduce.  Same value on
   every rank.  Returns 1 if peer memory is in use afterwards, 0 if not, negative on error. */
int glims_set_p2p(glims_ctx* c, int32_t on);
/* Average device time (microseconds, CUDA events on the context stream) of `reps` back-to-back collectives on every rank:
   kind 0 = halo exchange of a state-sized vector (FP64, dim+1 values per vertex), 1 = FP32 halo exchange with dim values
   per vertex (the one inside the V-cycle), 2 = allreduce of two scalars. */
int glims_comm_bench(glims_ctx* c, int32_t kind, int32_t reps, float* us_avg);

#ifdef __cplusplus
}
#endif
#endif
