/*
 * hgrid.h -- C ABI of libhgrid.so: B200 (sm_100a) grid-diffusion solver.
 *
 * This is the drop-in boundary for the hot path of yig/harmonic_interpolation.  The
 * reference has no FFI seam: its seam is the Python module `recovery` (and
 * `fill_alpha_harmonic`).  Each entry point below replaces a block of that module;
 * the citation is file:line in the reference tree.  The Python side
 * (harmonic_interpolation_b200/recovery.py) keeps the reference's call signatures and
 * binds these symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success and a negative hg_status on failure;
 *     hg_last_error() returns a thread-local, human-readable description;
 *   - plain pointers and sizes only, all host buffers caller-owned and C-contiguous;
 *   - a grid has `rows` x `cols` pixels, row-major index p = i*cols + j
 *     (recovery.py:254-256); value arrays are (rows, cols, K) with K innermost,
 *     exactly the layout the reference returns (recovery.py:1078, 1329);
 *   - one solver handle may be used by one thread at a time.
 */
#ifndef HGRID_H
#define HGRID_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGRID_VERSION 100

typedef struct hg_solver* hg_handle;

/* per-pixel flag byte (hg_set_flags) */
enum {
    HG_FLAG_HARD      = 1,  /* hard value constraint: recovery.py:647-777, 783-836            */
    HG_FLAG_CUT_DOWN  = 2,  /* edge (i,j)-(i+1,j) is cut: recovery.py:309-319, 422-439        */
    HG_FLAG_CUT_RIGHT = 4,  /* edge (i,j)-(i,j+1) is cut                                      */
    HG_FLAG_PEN_DOWN  = 8,  /* dx constraint x[i+1,j]-x[i,j] = d: recovery.py:549-552, 926-927 */
    HG_FLAG_PEN_RIGHT = 16, /* dy constraint x[i,j+1]-x[i,j] = d: recovery.py:554-557          */
    HG_FLAG_SOFT      = 32  /* soft value constraint: recovery.py:559-562, 1280-1283          */
};

/* operator: recovery.py:909-913 */
enum { HG_OP_LAPLACIAN = 0, HG_OP_BILAPLACIAN = 1 };

/* arithmetic: fp32 storage with fp64 reductions, or fp64 throughout */
enum { HG_PREC_FP32 = 0, HG_PREC_FP64 = 1 };

/* preconditioner of the conjugate-gradient solve */
enum { HG_PRECOND_JACOBI = 0, HG_PRECOND_MULTIGRID = 1 };

typedef enum {
    HG_OK = 0,
    HG_ERR_INVALID = -1,       /* bad argument / call order                                         */
    HG_ERR_CUDA = -2,          /* CUDA runtime failure                                              */
    HG_ERR_NO_DEVICE = -3,     /* no CUDA device: there is no CPU fallback                          */
    HG_ERR_UNREFERENCED = -4,  /* bilaplacian: a pixel is referenced by no row (recovery.py:495)    */
    HG_ERR_EMPTY_ROWS = -5,    /* bilaplacian without cuts: wrong number of empty rows (:493)       */
    HG_ERR_NOT_CONVERGED = -6, /* maxiter reached (the reference asserts info == 0, recovery.py:972) */
    HG_ERR_BREAKDOWN = -7,     /* NaN / non-positive curvature: singular or under-constrained system */
    HG_ERR_NOMEM = -8,
    HG_ERR_COMM = -9,          /* slab mode: NCCL failure or a peer that stopped; the communicator is aborted */
    HG_ERR_SINGULAR = -10      /* a connected component of free pixels carries no value constraint: the reference's
                                  direct solve raises ArithmeticError there (cvxopt.cholmod, recovery.py:993)          */
} hg_status;

typedef struct {
    int levels;          /* multigrid levels built (1 = none)              */
    int empty_rows;      /* bilaplacian: rows of L_refl that are all zero  */
    int unreferenced;    /* bilaplacian: pixels referenced by no row       */
    long long n_free;    /* free (unknown) pixels                          */
    long long n_hard;
    double setup_ms;
    long long unconstrained; /* free pixels in components without any value constraint (HG_ERR_SINGULAR if != 0) */
} hg_setup_info;

/* host value arrays of one solve; each is (rows, cols, K) float64 or NULL (= zeros) */
typedef struct {
    const double* hard_vals; /* read where HG_FLAG_HARD   (recovery.py:756, 835)            */
    const double* soft_vals; /* read where HG_FLAG_SOFT   (recovery.py:1321)                */
    const double* d_down;    /* read where HG_FLAG_PEN_DOWN  (recovery.py:552)              */
    const double* d_right;   /* read where HG_FLAG_PEN_RIGHT (recovery.py:557)              */
    const double* x0;        /* optional initial guess on the free pixels                   */
    const double* rhs;       /* optional explicit right-hand side, added at the free pixels
                                (G.T * target_gradients of poisson.py:73)                    */
} hg_values;

typedef struct {
    int iterations;
    int status;            /* hg_status of the solve                                        */
    double relres;         /* max over channels of ||b_f - A_ff x_f|| / ||b_f||, recomputed  */
    double relres_cg;      /* same, from the CG recurrence at exit                           */
    double solve_ms;       /* device time of the solve proper (CUDA events)                  */
    double upload_ms;      /* host -> device of the value arrays + conversion                */
    double download_ms;    /* device -> host of the result                                   */
    double bytes_moved;    /* algorithmic bytes of the solve (SURVEY 8(d) formulas)          */
    long long launches;    /* kernels launched by the solve                                  */
} hg_stats;

int hg_version(void);
const char* hg_last_error(void);
int hg_device_count(void);

/* Creates a solver for one rows x cols grid with K channels on the current CUDA device.
 * Replaces the operator construction gen_symmetric_grid_laplacian (recovery.py:232-347)
 * / gen_grid_laplacian2_with_boundary_reflection + L.T*L (recovery.py:349-500, 911). */
int hg_create(int rows, int cols, int channels, int op, int precision, hg_handle* out);
int hg_destroy(hg_handle h);
/* Solvers take their device memory from the device's stream-ordered pool and hg_destroy returns it to the POOL, not to the
 * driver (mapping / unmapping tens of GB costs more than a solve); hg_trim_memory() releases what the pool holds. */
int hg_trim_memory(void);

/* rows*cols flag bytes.  Replaces the sparse constraint matrices of gen_constraint_system
 * (recovery.py:504-573) and the cut-edge mask (recovery.py:309-319). */
int hg_set_flags(hg_handle h, const uint8_t* flags);
/* Per-pixel soft weights (rows*cols float64, read where HG_FLAG_SOFT) or NULL to use
 * `scalar_w` for every HG_FLAG_SOFT pixel.  Replaces W = diag(w_lsq), C.T*W*C
 * (recovery.py:1273-1283). */
int hg_set_soft_weights(hg_handle h, const double* weights, double scalar_w);
/* weight of the dx/dy constraints, w_gradients (recovery.py:875-884, 926-927) */
int hg_set_grad_weight(hg_handle h, double w_g);
int hg_set_preconditioner(hg_handle h, int precond);
/* Edge weights of the Laplacian: HG_EDGES_REFERENCE = gen_symmetric_grid_laplacian2 (1/4, and 1/8 along the
 * first / last row and column, recovery.py:263-299); HG_EDGES_UNIFORM = 1/4 on every edge, i.e. one quarter of
 * the graph Laplacian G.T*G of poisson.py:55-56 (gradient_operator, poisson.py:204-251). */
enum { HG_EDGES_REFERENCE = 0, HG_EDGES_UNIFORM = 1 };
int hg_set_edge_weights(hg_handle h, int mode);
/* Validates the domain (recovery.py:484-495) and builds the device-resident operator
 * and preconditioner.  Replaces the factorisation cvxopt.umfpack.numeric
 * (recovery.py:1304); call once, then hg_solve any number of times. */
int hg_setup(hg_handle h, hg_setup_info* info);

/* y = system * x at free pixels, 0 at hard pixels; x, y are (rows, cols, K) float64.
 * The matrix-free equivalent of `system * x` for the system of recovery.py:926, 1283. */
int hg_apply(hg_handle h, const double* x, double* y);

/* Solves for the grid.  Replaces rhs assembly + cvxopt.cholmod.linsolve / umfpack.solve
 * (recovery.py:926-943, 993, 1059-1075, 1319-1327).  x_out is (rows, cols, K) float64. */
int hg_solve(hg_handle h, const hg_values* vals, double* x_out, double rtol, int maxiter, hg_stats* stats);

/* Composed systems of the reference's gradient-domain integrator, poisson.solve_poisson_simple (poisson.py:6-99):
 *
 *     ( L^m + weight A^T A ) x = L^(m-1) g + weight A^T c        m = 1, or 2 for `squared` (bilaplacian=True, poisson.py:65-67)
 *
 * L is the handle's Laplacian WITHOUT its soft weights (with HG_EDGES_UNIFORM: G.T*G / 4 of poisson.py:55-56, 204-251);
 * g = vals->rhs (G.T * target_gradients of poisson.py:73, scaled like L); A holds the rows of generate_constraint_matrix
 * (poisson.py:329-366) in CSR form, any number of pixels per row, c their right-hand sides, (n_constraints, K).  The
 * handle's soft weights D only shape the preconditioner M = (L + D)^m: set D = diag(weight A^T A) for m = 1 and its
 * square root for m = 2.  fp64, Laplacian + multigrid handles on one GPU; constraint rows must not touch hard pixels, and
 * for m = 2 hard pixels must be separated from the free ones by cut edges (what a mask border of poisson.py is), so that
 * eliminating them from L^2 equals squaring the reduced L.  vals->hard_vals and vals->x0 as in hg_solve. */
typedef struct {
    int squared;             /* 0: m = 1, 1: m = 2                                            */
    int n_constraints;       /* rows of A                                                     */
    const int* row_ptr;      /* n_constraints + 1 offsets into pixel / coeff (row_ptr[0] = 0) */
    const int* pixel;        /* row-major pixel index i * cols + j of every entry             */
    const double* coeff;     /* coefficient of every entry                                    */
    const double* rhs;       /* (n_constraints, K) right-hand sides c                         */
    double weight;           /* linear_constraints_weight (poisson.py:88-89), scaled like L^m */
} hg_linear_constraints;
int hg_solve_composed(hg_handle h, const hg_values* vals, const hg_linear_constraints* lc, double* x_out, double rtol,
                      int maxiter, hg_stats* stats);

/* Image form of hg_solve for hard-constraint fills (the CLI's pipeline, fill_alpha_harmonic.py:27-45,
 * 72-76): hard values arrive as uint8 (rows, cols, K) and mean u8 / 255; the result is returned as
 * uint8 = clip(round_half_even(255 x), 0, 255).  Soft / gradient value arrays are taken as zero.
 * Moves 1 byte per pixel and channel each way instead of 8. */
int hg_solve_u8(hg_handle h, const uint8_t* hard_vals, uint8_t* out, double rtol, int maxiter, hg_stats* stats);

/* The same in three steps, so that a benchmark can time the device-resident solve. */
int hg_upload_values(hg_handle h, const hg_values* vals, hg_stats* stats);
/* List form of hg_upload_values for the factor-once closure (recovery.py:1304-1336), which is called with the VALUES of the
 * hard / soft constraints whose positions were fixed at construction: row-major pixel indices i * cols + j (unique) and
 * (n, K) values; the library scatters them on the device instead of the caller filling (rows, cols, K) planes.  Gradient
 * values, initial guess and explicit right-hand side are taken as absent.  Follow with hg_solve_device + hg_download. */
int hg_upload_values_sparse(hg_handle h, long long n_hard, const long long* hard_idx, const double* hard_vals,
                            long long n_soft, const long long* soft_idx, const double* soft_vals, hg_stats* stats);
int hg_solve_device(hg_handle h, double rtol, int maxiter, hg_stats* stats);
int hg_download(hg_handle h, double* x_out, hg_stats* stats);
/* Sets the device-resident iterate back to the zero guess on the free pixels (hard values
 * stay), so that hg_solve_device can be timed repeatedly on the same uploaded values. */
int hg_reset_guess(hg_handle h);

/* ---- multi-GPU: row slabs of one grid, one process per GPU ------------------------------------
 * The global rows are cut into bands at multiples of HG_GHOST_ROWS.  Each process creates its
 * solver on the band it owns PLUS a ghost band of HG_GHOST_ROWS rows towards each neighbouring
 * slab; the ghost rows MIRROR the neighbour (same flags / soft weights; values are refreshed by
 * the library).  After hg_comm_init, hg_setup builds the local hierarchy down to level 5 and a
 * replicated global hierarchy below it; during a solve the library refreshes ghost bands over
 * NCCL send / recv (NVLink) once per level and leg of the V-cycle and before every operator
 * application, gathers the level-5 right-hand side, and all-reduces the K conjugate-gradient
 * scalars.  The V-cycle computed this way is the single-GPU V-cycle: iteration counts do not
 * depend on the number of slabs.  hg_comm_unique_id: HG_COMM_ID_BYTES bytes from rank 0, to be
 * broadcast by the caller.  row_starts: first owned global row of every rank (world entries).
 * Laplacian + multigrid only. */
#define HG_COMM_ID_BYTES 128
#define HG_GHOST_ROWS 128
int hg_comm_unique_id(uint8_t* id);
int hg_comm_init(hg_handle h, const uint8_t* id, int rank, int world, int ghost_top, int ghost_bottom,
                 const int* row_starts, int global_rows);

/* z = M^-1 r: one application of the preconditioner (Jacobi or one multigrid V-cycle) to a
 * host array; r is masked to the free pixels first.  For tests of symmetry / definiteness. */
int hg_precond_apply(hg_handle h, const double* r, double* z);
/* Multigrid hierarchy introspection: number of levels (level 0 = the flag-encoded operator),
 * and the explicit 9-point Galerkin stencil of level >= 1 as (rows, cols, 9) float64,
 * entry c = (di+1)*3 + (dj+1).  coef9 may be NULL to query the size only. */
int hg_mg_levels(hg_handle h);
int hg_mg_get_level(hg_handle h, int level, int* rows, int* cols, double* coef9);

/* x and b of hierarchy level >= 1 as the last V-cycle left them, each (rows, cols, K) float64 or NULL: lets a test compare
 * the legs of the V-cycle level by level (tile legs / streaming legs / the unfused reference kernels). */
int hg_mg_get_level_vectors(hg_handle h, int level, double* x, double* b);

/* Times `iters` launches of the stencil kernel on device-resident data (CUDA events on
 * the solver's stream); returns mean milliseconds per launch. */
int hg_bench_apply(hg_handle h, int iters, double* ms_per_launch);

/* Per-phase device time of the MG-PCG loop (Laplacian + multigrid path): after hg_set_phase_timing(h, 1) every
 * hg_solve_device records CUDA events on the solver's stream between its launches; hg_get_phase_ms returns, per
 * phase, the milliseconds (and the number of intervals) of the last solve.  The events cost a few microseconds
 * each, so a timed benchmark run leaves this off.  No reference counterpart (tictoc.py prints host wall time). */
enum {
    HG_PH_OTHER = 0,
    HG_PH_STENCIL = 1,   /* q = A p, p.q, alpha                                  */
    HG_PH_UPDATE = 2,    /* x += alpha p, r -= alpha q, r.r, convergence check   */
    HG_PH_FINE_DOWN = 3, /* V-cycle level 0, downward leg (+ ghost exchange of r) */
    HG_PH_L1_DOWN = 4,   /* level 1, downward leg                                 */
    HG_PH_COARSE = 5,    /* levels >= 2, both legs, coarsest solve                */
    HG_PH_L1_UP = 6,     /* level 1, upward leg                                   */
    HG_PH_FINE_UP = 7,   /* level 0, upward leg, r.z                              */
    HG_PH_PUPDATE = 8,   /* beta, p = z + beta p                                  */
    HG_PH_OUTER = 9,     /* true residual in fp64 (outer refinement loop)         */
    HG_N_PHASES = 10
};
int hg_set_phase_timing(hg_handle h, int enable);
int hg_get_phase_ms(hg_handle h, double* ms, int* counts, int n);

#ifdef __cplusplus
}
#endif
#endif /* HGRID_H */
