// solver.cuh -- the solver handle behind the C ABI of libhgrid.so (include/hgrid.h): state, small inline helpers and the
// launch templates.  The member functions live in six translation units, compiled in parallel:
//   solver_setup.cu   handle life cycle, pattern upload, hg_setup (domain checks, unconstrained components), value upload /
//                     download, operator application, introspection
//   solver_mg.cu      Laplacian path: code bytes / activity / task lists, the Galerkin hierarchy, the V-cycle and the
//                     launches of the streaming and tile kernels
//   solver_mg25.cu    bilaplacian path: the 25-point hierarchy and its V-cycle
//   solver_slab.cu    multi-GPU slab mode: communicator, gather of the replicated levels, failure handling
//   solver_pcg.cu     the scalar stages and the MG-PCG / refinement driver
//   solver_composed.cu  the composed systems of poisson.py (squared Laplacian, constraint rows over several pixels)
// capi.cu holds the extern "C" entry points, the error string and the run-time binding of NCCL.
#pragma once
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "coarse_fused.cuh"
#include "coarse_stream.cuh"
#include "common.cuh"
#include "fine.cuh"
#include "fine_stream.cuh"
#include "fine_vec.cuh"
#include "mg.cuh"
#include "mg25.cuh"
#include "mg25_clusters.cuh"
#include "pcg.cuh"
#include "stencil.cuh"

namespace hg {

extern thread_local std::string g_err;

inline int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return fail(e_ == cudaErrorMemoryAllocation ? HG_ERR_NOMEM : HG_ERR_CUDA, buf_);        \
        }                                                                                          \
    } while (0)

#define HGTRY(call)                 \
    do {                            \
        int rc_ = (call);           \
        if (rc_ != HG_OK) return rc_; \
    } while (0)

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- NCCL, bound at run time -------------------------------------------------------------------
// Only the multi-GPU slab mode needs NCCL; it is opened with dlopen on first use so that the
// library has no hard dependency on it (in a torch process this resolves to the libnccl.so.2 that
// torch has already loaded).
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
extern Nccl g_nccl;

// Communicators are kept for the life of the process, one per (world, rank, device): creating one costs about a second
// (bootstrap, channel set-up on the first send / recv to every peer), which would otherwise be paid by every solver that
// is constructed -- a one-shot solve of a 16384^2 slab takes a tenth of that.  All ranks construct their solvers in the same
// order, so either every rank finds its communicator here or none does.  An aborted communicator leaves the cache.
struct CommKey { int world, rank, dev; bool operator<(const CommKey& o) const { return world != o.world ? world < o.world : (rank != o.rank ? rank < o.rank : dev < o.dev); } };
extern std::map<CommKey, ncclComm_t> g_comms;
extern std::mutex g_comms_mu;

int nccl_load();

#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return fail(HG_ERR_CUDA, buf_);                                                        \
        }                                                                                          \
    } while (0)

struct SolverBase {
    virtual ~SolverBase() {}
    virtual int set_flags(const uint8_t* f) = 0;
    virtual int set_soft(const double* w, double scalar) = 0;
    virtual int set_grad_weight(double w) = 0;
    virtual int set_precond(int p) = 0;
    virtual int set_edges(int mode) = 0;
    virtual int setup(hg_setup_info* info) = 0;
    virtual int apply(const double* x, double* y) = 0;
    virtual int upload(const hg_values* v, hg_stats* st) = 0;
    virtual int upload_sparse(long long nh, const long long* hidx, const double* hv, long long ns, const long long* sidx, const double* sv, hg_stats* st) = 0;
    virtual int solve_device(double rtol, int maxiter, hg_stats* st) = 0;
    virtual int solve_composed(const hg_linear_constraints* lc, double rtol, int maxiter, hg_stats* st) = 0;
    virtual int download(double* x, hg_stats* st) = 0;
    virtual int bench_apply(int iters, double* ms) = 0;
    virtual int mg_levels() = 0;
    virtual int mg_get_level(int level, int* rows, int* cols, double* coef9) = 0;
    virtual int precond_apply(const double* r, double* z) = 0;
    virtual int reset_guess() = 0;
    virtual int upload_u8(const uint8_t* vals, hg_stats* st) = 0;
    virtual int download_u8(uint8_t* out, hg_stats* st) = 0;
    virtual int comm_init(const uint8_t* id, int rank, int world, int ghost_top, int ghost_bottom, const int* row_starts, int global_rows) = 0;
    virtual int mg_get_level_vectors(int level, double* x, double* b) = 0;
    virtual int set_phase_timing(int on) = 0;
    virtual int get_phase_ms(double* ms, int* counts, int n) = 0;
};

template <typename T>
struct Solver : SolverBase {
    int rows, cols, K, op, pitch;
    long long plane;
    int precond = HG_PRECOND_JACOBI;
    bool flags_set = false, is_setup = false, have_values = false;
    double soft_scalar = 0.0, w_g = 0.0;
    bool soft_plane_set = false;

    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // Device memory comes from the device's default stream-ordered pool (cudaMallocAsync), whose release threshold init()
    // raises to "never": a solver's tens of GB are mapped once per process, later solvers of the process reuse them, and
    // destruction does not pay for unmapping (cudaMalloc / cudaFree of a 16384^2 solver cost 0.4 s + 0.2 .. 1 s).
    // hg_trim_memory() hands the cached memory back to the driver.
    template <typename P> cudaError_t dmalloc(P** p, size_t n) { return cudaMallocAsync(reinterpret_cast<void**>(p), n ? n : 1, stream); }
    cudaError_t dfree(void* p) { return p ? cudaFreeAsync(p, stream) : cudaSuccess; }

    uint8_t* d_flags = nullptr;
    T* d_soft = nullptr;
    T* d_dinv = nullptr;
    T *d_x = nullptr, *d_b = nullptr, *d_r = nullptr, *d_p = nullptr, *d_q = nullptr;
    double* d_stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // staged host arrays
    int edge_uniform = 0;
    double* d_out = nullptr;
    double *d_part0 = nullptr, *d_part1 = nullptr;
    int part_cap = 0;
    CgState* d_state = nullptr;
    double* d_presum = nullptr;
    unsigned long long* d_counts = nullptr;
    long long launches = 0;

    // fp32 mode keeps the iterate and the right-hand side in fp64 ("master") for iterative
    // refinement; in fp64 mode these alias d_x / d_b
    double *d_xm = nullptr, *d_bm = nullptr;
    static constexpr bool kRefine = !std::is_same<T, double>::value;

    // multi-GPU slab mode: this solver owns a band of rows of a larger grid; the rows directly
    // above / below (ghost rows, flagged hard) mirror the neighbouring slabs
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, ghost_top = 0, ghost_bottom = 0, global_rows = 0;
    std::vector<int> row_starts;             // first owned global row of every rank (+ global_rows at the end)
    std::vector<Coarse<T>> gcoarse;          // replicated global hierarchy from level HG_GATHER_LEVEL down
    int* d_gl_rows = nullptr;                // [2 q] first row, [2 q + 1] row count of rank q at the gather level
    double* d_gdense = nullptr;

    // multigrid hierarchy (levels >= 1; level 0 is the flag-encoded operator)
    T* d_z = nullptr;
    std::vector<Coarse<T>> coarse;
    double* d_dense = nullptr;   // [A | A^-1] of the coarsest level
    // bilaplacian: hierarchy of explicit 25-point stencils, level 0 included (mg25.cuh)
    std::vector<Lev25<T>> lev25;
    double* d_dense25 = nullptr;
    int mg25_sweeps = getenv("HGRID_MG25_SWEEPS") ? atoi(getenv("HGRID_MG25_SWEEPS")) : 1;
    // tile path (Laplacian + multigrid): activity map and per-tile partial sums
    uint8_t* d_tile_active = nullptr;
    double *d_tpart0 = nullptr, *d_tpart1 = nullptr;
    int tiles_x = 0, tiles_y = 0;
    uint8_t* d_down_active = nullptr;   // tiles with a free pixel within one pixel (restriction reaches that far)
    uint8_t* d_code = nullptr;          // one byte of edge-weight codes per pixel (fine.cuh)
    unsigned* d_qmask = nullptr;        // lean mask words: free-pixel bits of 8 rows x 4 pixels (fine_stream.cuh: sl_qmask_kernel)
    double* d_lpart = nullptr;          // partial sums of the stencil kernel (two blocks per tile)
    // streaming legs (fine_stream.cuh): chunk activity (halo 0 / halo 1) and per-chunk partial sums of r.z
    uint8_t *d_sl_up = nullptr, *d_sl_down = nullptr;
    double* d_spart = nullptr;
    double* d_spart2 = nullptr;        // per-chunk partial sums of p.q (streaming stencil kernel)
    // Phase timing (hg_set_phase_timing, or HGRID_PHASES=1 which also prints to stderr): CUDA events on the solver's
    // stream between the launches of the MG-PCG loop; totals per phase of the last solve are kept for hg_get_phase_ms.
    bool phase_print = getenv("HGRID_PHASES") != nullptr;
    bool phase_timing = phase_print;
    std::vector<cudaEvent_t> phase_ev;
    std::vector<int> phase_id;
    double phase_tot[HG_N_PHASES] = {0};
    int phase_cnt[HG_N_PHASES] = {0};
    void mark(int id);
    void report_phases();
    int set_phase_timing(int on) override { phase_timing = on != 0 || phase_print; return HG_OK; }
    int get_phase_ms(double* ms, int* counts, int n) override {
        for (int k = 0; k < n && k < HG_N_PHASES; ++k) { if (ms) ms[k] = phase_tot[k]; if (counts) counts[k] = phase_cnt[k]; }
        return HG_OK;
    }
    bool exchange_p = getenv("HGRID_EXCHANGE_P") != nullptr;
    // Ghost-band refreshes inside the V-cycle.  What a leg computes on ghost rows is wrong only near the OUTER edge of the
    // band, and the wrong region grows slowly: going down, b_{l+1} is wrong within floor((w + 2) / 2) + 1 rows of the edge if
    // b_l is wrong within w (sweep 1 row, residual 1, restriction 1): 2, 3, 3, 3, 3 rows at levels 1 .. 5 against bands of
    // 64 .. 4 rows -- no refresh is needed on the way down.  Going up a level doubles the region and the reversed sweep adds 2
    // rows: 6 of 8 rows at level 4 without any refresh, and 14 / 30 / 62 / 126 of 16 / 32 / 64 / 128 above it -- valid, but
    // only just.  ONE refresh of x at level 4 (8 rows of a 1/16-width plane) resets that to 6 / 14 / 30 / 62.  So a V-cycle
    // exchanges r (level 0, 128 rows), gathers level 5 and refreshes level 4 once; HGRID_EXCHANGE_ALL=1 restores a refresh per
    // level and leg (8 more send / recv groups per cycle).  tests/dist_check.py holds both to the single-GPU answer.
    bool exchange_all = getenv("HGRID_EXCHANGE_ALL") != nullptr;
    // HGRID_OVERLAP=1 (slab mode; off by default): the 128-row exchange of r at the head of the V-cycle travels on a second
    // stream while the downward leg works on the chunks that read no ghost row; the chunks next to the ghost bands follow
    // once the rows have arrived.  Run in round 2 on 4 GPUs: exact (profiles/r2/dist_check_world4_overlap.log), 74.9 vs
    // 74.5 ms -- the exchange costs less than the split launch saves, so it stays an option.
    bool overlap = getenv("HGRID_OVERLAP") != nullptr;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_arrived = nullptr;
    int sl_n_int[2] = {0, 0};          // down lists (lean, general): number of interior tasks, which come first
    int sl_ns = 0, sl_nt = 0;
    int* d_sl_tasks = nullptr;
    int* d_task_counts = nullptr;
    bool task_counts_pending = false;
    int finish_task_lists();
    int sl_off[4] = {0, 0, 0, 0}, sl_n_down_lean = 0, sl_n_down_gen = 0, sl_n_up_lean = 0, sl_n_up_gen = 0;
    bool tile_legs = getenv("HGRID_TILE_LEGS") != nullptr;       // debugging aid: the shared-memory legs of fine_vec.cuh
    bool scalar_legs = getenv("HGRID_SCALAR_LEGS") != nullptr;
    bool unfused_coarse = getenv("HGRID_UNFUSED_COARSE") != nullptr;   // debugging aid: one kernel per colour / transfer (mg.cuh)   // debugging aid: the scalar fused legs of fine.cuh
    bool b_is_zero = false;             // no soft / gradient constraints: the right-hand side is identically 0
    bool has_pen = false;

    Solver(int r, int c, int k, int o) : rows(r), cols(c), K(k), op(o) {
        pitch = round_up(cols, 32);
        plane = (long long)rows * pitch;
    }
    ~Solver() override;
    int init();

    Grid<T> grid() const {
        Grid<T> g;
        g.rows = rows; g.cols = cols; g.pitch = pitch; g.plane = plane; g.K = K;
        g.flags = d_flags; g.soft = d_soft; g.soft_scalar = (T)soft_scalar; g.w_g = (T)w_g;
        g.uniform = edge_uniform;
        return g;
    }

    // ---- grid geometry of the operator launch ------------------------------------------
    dim3 op_grid() const {
        if (op == HG_OP_LAPLACIAN) return dim3((pitch / 4 + LAP_TX - 1) / LAP_TX, (rows + LAP_ROWS - 1) / LAP_ROWS);
        return dim3((pitch + BilapTile<T>::TX - 1) / BilapTile<T>::TX, (rows + BilapTile<T>::TY - 1) / BilapTile<T>::TY);
    }
    int op_blocks() const { dim3 g = op_grid(); return g.x * g.y; }
    int vec_blocks() const {
        long long n4 = plane / 4;
        long long nb = (n4 + VEC_THREADS - 1) / VEC_THREADS;
        return (int)std::min<long long>(nb, 148 * 8);
    }

    template <int MODE, int DOT>
    void launch_op(const T* x, const T* b, T* y, double* partial, const int* done) {
        const Grid<T> g = grid();
        if (op == HG_OP_LAPLACIAN)
            lap_apply_kernel<T, MODE, DOT><<<op_grid(), LAP_TX, 0, stream>>>(g, x, b, y, partial, done);
        else
            bilap_apply_kernel<T, MODE, DOT><<<op_grid(), BILAP_THREADS, 0, stream>>>(g, x, b, y, partial, done);
        ++launches;
    }

    uint8_t* d_raw = nullptr;     // rows * cols bytes as the caller holds them (flags; also the uint8 image of hg_solve_u8)
    int set_flags(const uint8_t* f) override;
    int set_soft(const double* w, double scalar) override;

    int set_grad_weight(double w) override {
        w_g = w;
        is_setup = false;
        return HG_OK;
    }
    int set_edges(int mode) override;
    int set_precond(int p) override;
    int ensure_vectors();


    // ---- multigrid ----------------------------------------------------------------------------
    void free_hierarchy();

    static int chain(int n, int levels) { for (int l = 0; l < levels; ++l) n = (n + 1) / 2; return n; }
    int alloc_level(Coarse<T>& L, int r, int c);

    // Interior stencil of every level: depends on T and the level only.  Taken from the hierarchy of
    // a 256^2 grid without constraints (centre pixel of each level; the same kernel, so bit-identical
    // to what the real hierarchy holds wherever the neighbourhood is regular).
    static constexpr int REF_LEVELS = 5;
    T ref_stencil[REF_LEVELS + 1][9];
    bool ref_ready = false;
    // (process-wide cache per storage type: the values depend on nothing else)
    struct RefCache { T st[REF_LEVELS + 1][9]; bool ready; };
    static RefCache& ref_cache() { static RefCache c = {}; return c; }
    int build_ref_stencils();

    // code bytes and tile activity of a level whose coefficients are final (`level` >= 1)
    int finish_level(Coarse<T>& L, int level);
    bool coarse_tiles = getenv("HGRID_COARSE_TILES") != nullptr;     // A/B aid: the shared-memory tile legs of coarse_fused.cuh on every level
    template <bool UP>
    void launch_cs(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
        constexpr int WPB = CsCfg<T, UP>::WPB;
        const int nchunks = L.cs_ns * L.cs_nt;
        const unsigned blocks = (unsigned)(((long long)nchunks * K + WPB - 1) / WPB);
        const size_t sm = (size_t)(UP ? 2 : 1) * WPB * CS_RING * 128 * sizeof(T) + 32 * WPB * sizeof(float);   // rings + scratch words
        cs_leg_kernel<T, UP><<<blocks, 32 * WPB, sm, stream>>>(L, Lc, 1, L.cs_act, L.cs_cr, L.cs_ns, nchunks, done);
    }

    template <int KC>
    int set_fused_attrs() {
        CU(cudaFuncSetAttribute(cdown_kernel<T, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CFSmem<T, KC, false>)));
        CU(cudaFuncSetAttribute(cup_kernel<T, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CFSmem<T, KC, true>)));
        CU(cudaFuncSetAttribute(cs_leg_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * CsCfg<T, true>::WPB * CS_RING * 128 * sizeof(T) + 32 * CsCfg<T, true>::WPB * sizeof(float))));
        CU(cudaFuncSetAttribute(cs_leg_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CsCfg<T, false>::WPB * CS_RING * 128 * sizeof(T) + 32 * CsCfg<T, false>::WPB * sizeof(float))));
        return HG_OK;
    }
    template <int KC>
    void launch_cdown(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
        cdown_kernel<T, KC><<<dim3(L.tiles_x, L.tiles_y, (K + KC - 1) / KC), CF_THREADS, sizeof(CFSmem<T, KC, false>), stream>>>(L, Lc, done);
    }
    template <int KC>
    void launch_cup(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
        cup_kernel<T, KC><<<dim3(L.tiles_x, L.tiles_y, (K + KC - 1) / KC), CF_THREADS, sizeof(CFSmem<T, KC, true>), stream>>>(L, Lc, done);
    }
    // fused legs of a 9-point level: pre-smoothing + residual + restriction / prolongation + post-smoothing
    void coarse_down(const Coarse<T>& L, const Coarse<T>& Lc, const int* done);
    // (the fused upward legs write x2; afterwards x2 IS the level's x: the host-side structs are swapped, every later
    // launch -- the level above, the next V-cycle -- takes the struct by value and sees the current pointers)
    void coarse_up(Coarse<T>& L, const Coarse<T>& Lc, const int* done);
    int build_dense(const Coarse<T>& L, double** out);

    // slab mode: rows of level HG_GATHER_LEVEL owned by rank q, in the replicated global level
    int gl_start(int q) const { return row_starts[q] >> HG_GATHER_LEVEL; }
    int gl_end(int q) const { return q == world - 1 ? chain(global_rows, HG_GATHER_LEVEL) : row_starts[q + 1] >> HG_GATHER_LEVEL; }

    // every rank's owned rows of `nplanes` planes of the local gather level -> the global level: packed into one
    // equal-sized block per rank, ONE all-gather, unpacked (a broadcast per plane and rank cost 24 collectives per
    // V-cycle on 8 GPUs)
    T *d_gsend = nullptr, *d_grecv = nullptr;
    size_t gbuf_elems = 0;
    int gather_level(const T* local, long long lplane, T* global, long long gplane, int pitch_, int nplanes);
    int build_global_hierarchy();

    bool use_mg() const { return precond == HG_PRECOND_MULTIGRID && op == HG_OP_LAPLACIAN; }
    bool use_mg25() const { return precond == HG_PRECOND_MULTIGRID && op == HG_OP_BILAPLACIAN; }

    // ---- bilaplacian hierarchy --------------------------------------------------------------------
    int build_hierarchy25();
    // cluster-flat transfers of level 0 for stiff gradient constraints (mg25_clusters.cuh); HGRID_MG25_CLUSTERS=0 switches them off
    Clusters cl25 = {nullptr, nullptr, 0};
    T *cl_sum = nullptr, *cl_tmp = nullptr;      // K fine planes each: cluster sums; the prolongated correction
    int build_clusters25();
    void cluster_average(T* v, int nplanes);
    int probe_level1(const Lev25<T>& L0, Lev25<T>& L1);
    void gs25(Lev25<T>& L, bool forward, const int* done);
    // z = M^-1 r: one symmetric V-cycle from the zero guess on the 25-point hierarchy (level 0 works in place on z, r)
    int vcycle25(T* z, T* r, const int* done);

    dim3 tile_grid() const { return dim3(tiles_x, tiles_y); }

    int ntiles() const { return tiles_x * tiles_y; }
    int build_tiles();
    int build_hierarchy();
    void coarse_pre(const Coarse<T>& L, const int* done);
    void coarse_post(const Coarse<T>& L, const int* done);
    void coarse_restrict(const Coarse<T>& L, const Coarse<T>& Lc, const int* done);
    void coarse_prolong(const Coarse<T>& L, const Coarse<T>& Lc, const int* done);
    void coarse_solve(const Coarse<T>& L, const double* dense, const int* done);
    // V-cycle over levels[from..] of a hierarchy whose first level already holds b (and x = 0)
    void coarse_vcycle(std::vector<Coarse<T>>& lv, const double* dense, const int* done);

    // z = M^-1 r : one V(1,1) cycle from a zero initial guess; also leaves the per-tile partial
    // sums of r.z in part_rz.  Level 0 is two fused tile kernels (fine.cuh).  In slab mode the
    // ghost bands are refreshed once per level and leg and levels >= HG_GATHER_LEVEL run replicated.
    // Experiment aid (HGRID_GENERIC_VCYCLE=nu): a V(nu,nu) cycle assembled from the separate, unfused kernels
    // of mg.cuh on every level -- slow, but any nu; used to measure what more smoothing buys in iterations.
    int generic_nu = getenv("HGRID_GENERIC_VCYCLE") ? atoi(getenv("HGRID_GENERIC_VCYCLE")) : 0;
    int vcycle_generic(T* z, T* r, const int* done);
    int vcycle(T* z, T* r, const int* done, double* part_rz);

    // the streaming legs serve every configuration without gradient-penalty edges
    bool stream_legs() const { return !has_pen && !tile_legs && !scalar_legs && plane < (1LL << 31) - (1LL << 20); }   // 32-bit offsets on the lean paths
    bool has_soft() const { return d_soft != nullptr || soft_scalar != 0.0; }
    dim3 sl_grid(int ntasks) const { return dim3((unsigned)(((long long)ntasks * K + SL_WPB - 1) / SL_WPB)); }
    dim3 sl_block() const { return dim3(32 * SL_WPB); }
    size_t sl_smem() const { return (size_t)SL_WPB * SL_RING * 128 * sizeof(T); }   // lean kernels: one ring of rows per warp
    // part 0: every task; 1: the interior tasks (front of the lists); 2: the boundary tasks (HGRID_OVERLAP)
    void launch_sl_down(const Grid<T>& g, const T* r, const Coarse<T>& c1, const int* done, int part = 0);
    void launch_sl_up(const Grid<T>& g, const T* r, T* z, const Coarse<T>& c1, int have_coarse, const int* done);
    // q = A p with the streaming kernel (no soft weights, no gradient-penalty edges), partial sums of p.q per chunk
    bool stream_lap() const { return stream_legs() && !has_soft() && !getenv("HGRID_TILE_LAP"); }
    void launch_sl_lap(const T* p, T* q, const int* done);
    // r = -A x from the fp64 master iterate (the true residual when the right-hand side is zero), partial sums of r.r
    void launch_sl_resid(const double* x, T* r);
    // the stencil launch of the CG loop + the scalar stage that forms alpha
    int lap_and_alpha(const T* p, T* q, const int* done);
    // scalar stage after the V-cycle: r.z from the partial sums the upward leg left (owned rows only)
    int stage_beta_rz(int init);

    int mg_levels() override { return 1 + (int)coarse.size(); }
    int mg_get_level(int level, int* r_out, int* c_out, double* coef9) override;

    // x and b of hierarchy level >= 1 as the last V-cycle left them, (rows, cols, K) float64 (debugging / tests of the legs)
    int mg_get_level_vectors(int level, double* xo, double* bo) override;

    // z = M^-1 r for host arrays (parity tests of the preconditioner: symmetry, definiteness)
    int precond_apply(const double* r, double* z) override;

    // Free pixels of connected components that no value constraint touches (pcg.cuh).  One slab cannot decide this (a
    // component may be anchored in a neighbouring slab): not checked in slab mode.  HGRID_ALLOW_SINGULAR=1 skips the check.
    unsigned long long n_unconstrained = 0;
    int find_unconstrained(unsigned long long* out);
    int setup(hg_setup_info* info) override;
    int stage(int slot, const double* host);

    dim3 lap_grid() const { return dim3(tiles_x, (rows + LT_ROWS - 1) / LT_ROWS); }
    int n_lap() const { dim3 l = lap_grid(); return l.x * l.y; }

    template <typename TX, int MODE, int DOT, bool PEN>
    void launch_lap_tile_kc(const TX* x, const TX* b, T* y, double* partial, const int* done) {
        const Grid<T> g = grid();
        const dim3 lg = lap_grid();
        // as many channels per block as fit in registers (fp64 operands: two)
        const int kc = std::min(K, sizeof(TX) == 8 ? 2 : 4);
        dim3 lgz = lg;
        lgz.z = (K + kc - 1) / kc;
        switch (kc) {
            case 1: lap_tile_kernel<T, TX, MODE, DOT, 1, PEN><<<lgz, FT_THREADS, 0, stream>>>(g, d_code, x, b, y, d_tile_active, partial, done); break;
            case 2: lap_tile_kernel<T, TX, MODE, DOT, 2, PEN><<<lgz, FT_THREADS, 0, stream>>>(g, d_code, x, b, y, d_tile_active, partial, done); break;
            case 3: lap_tile_kernel<T, TX, MODE, DOT, 3, PEN><<<lgz, FT_THREADS, 0, stream>>>(g, d_code, x, b, y, d_tile_active, partial, done); break;
            default: lap_tile_kernel<T, TX, MODE, DOT, 4, PEN><<<lgz, FT_THREADS, 0, stream>>>(g, d_code, x, b, y, d_tile_active, partial, done); break;
        }
        ++launches;
    }
    template <typename TX, int MODE, int DOT>
    void launch_lap_tile(const TX* x, const TX* b, T* y, double* partial, const int* done) {
        if (has_pen) launch_lap_tile_kc<TX, MODE, DOT, true>(x, b, y, partial, done);
        else launch_lap_tile_kc<TX, MODE, DOT, false>(x, b, y, partial, done);
    }

    // y = F(x) with the kernel the solver itself uses for this configuration
    void apply_op(const T* x, T* y);
    int apply(const double* x, double* y) override;
    int upload(const hg_values* v, hg_stats* st) override;
    int upload_sparse(long long nh, const long long* hidx, const double* hv, long long ns, const long long* sidx, const double* sv, hg_stats* st) override;

    // Slab mode.  The local grid is [ghost band | owned rows | ghost band]; ghost bands are
    // HG_GHOST_ROWS deep MIRRORS of the neighbouring slabs (same flags, same values after a
    // refresh).  Every kernel simply runs on the whole local grid: what it computes on ghost rows
    // is redundant with the neighbour and valid up to a few rows from the outer edge, so a whole
    // smoothing sweep needs no communication and its result on the owned rows is exactly the
    // single-GPU one.  Ghost bands are refreshed (send / recv) once per level and leg; from level
    // HG_GATHER_LEVEL down the hierarchy is replicated on every rank.  Dot products run over the
    // owned tile rows only and are all-reduced.
    int comm_init(const uint8_t* id, int rank_, int world_, int gt, int gb, const int* starts, int grows) override;

    // refresh `depth_t` / `depth_b` ghost rows of the planes of `v` from the neighbours' owned rows
    // (level geometry given by pitch_, plane_, rows_, and the ghost-band depths gt_, gb_ at that level)
    template <typename V>
    int exchange_rows(V* v, int nplanes, int pitch_, long long plane_, int rows_, int gt_, int gb_, int depth, cudaStream_t xs = nullptr) {
        if (world <= 1) return HG_OK;
        cudaStream_t stream = xs ? xs : this->stream;
        const ncclDataType_t dt = sizeof(V) == 8 ? ncclDouble : ncclFloat;
        NC(g_nccl.GroupStart());
        ncclResult_t bad = ncclSuccess;
        auto note = [&](ncclResult_t r) { if (r != ncclSuccess && bad == ncclSuccess) bad = r; };
        for (int k = 0; k < nplanes && bad == ncclSuccess; ++k) {
            V* base = v + (long long)k * plane_;
            if (gt_ > 0) {
                const int d = std::min(depth, gt_);
                note(g_nccl.Send(base + (long long)gt_ * pitch_, (size_t)d * pitch_, dt, rank - 1, comm, stream));
                note(g_nccl.Recv(base + (long long)(gt_ - d) * pitch_, (size_t)d * pitch_, dt, rank - 1, comm, stream));
            }
            if (gb_ > 0) {
                const int d = std::min(depth, gb_);
                note(g_nccl.Send(base + (long long)(rows_ - gb_ - d) * pitch_, (size_t)d * pitch_, dt, rank + 1, comm, stream));
                note(g_nccl.Recv(base + (long long)(rows_ - gb_) * pitch_, (size_t)d * pitch_, dt, rank + 1, comm, stream));
            }
        }
        note(g_nccl.GroupEnd());       // the group is closed whatever happened inside it
        NC(bad);
        return HG_OK;
    }
    // level-0 vector, `depth` ghost rows
    template <typename V>
    int exchange(V* v, int depth, cudaStream_t xs = nullptr) { return exchange_rows(v, K, pitch, plane, rows, ghost_top, ghost_bottom, depth, xs); }
    template <typename V>
    int exchange_level(const Coarse<T>& L, int level, V* v, int nplanes) {
        return exchange_rows(v, nplanes, L.pitch, L.plane, L.rows, ghost_top >> level, ghost_bottom >> level, HG_GHOST_ROWS);
    }

    // owned range of the per-tile partial sums (dot products skip the ghost tile rows)
    int own_ty0() const { return ghost_top / FT_Y; }
    int own_ty1() const { return ghost_bottom ? (rows - ghost_bottom) / FT_Y : tiles_y; }
    const double* tile_part(const double* part) const { return part + (long long)own_ty0() * tiles_x; }
    int n_own_tiles() const { return (own_ty1() - own_ty0()) * tiles_x; }
    // the stencil kernel has two blocks (of LT_ROWS rows) per tile
    const double* lap_part(const double* part) const { return part + (long long)(ghost_top / LT_ROWS) * tiles_x; }
    int n_own_lap() const {
        const int y1 = ghost_bottom ? (rows - ghost_bottom) / LT_ROWS : (rows + LT_ROWS - 1) / LT_ROWS;
        return (y1 - ghost_top / LT_ROWS) * tiles_x;
    }
    int allreduce_tmp();

    // the four scalar stages; with several slabs each is split around an allreduce of the local sums.
    // `part` points at the first owned partial of channel 0, `n` owned partials, `stride` per channel.
    // many partial sums: add them in chunks first (presum_kernel), the scalar stage adds the chunk sums
    void presum(const double*& part, int& n, long long& stride, const int* done);
    int stage_alpha(const double* part, int n, long long stride);
    int stage_check(const double* part, int n, long long stride);
    int stage_beta(const double* part, int n, long long stride, int init);
    int stage_outer(const double* part, int n, long long stride, int first, double red2);

    // Solve on the device-resident master iterate (hard values + initial guess) and b.
    //
    // Outer loop (iterative refinement / verified restart): the TRUE residual r = b - F(x) is
    // evaluated in the master precision (fp64 even in fp32 mode) and decides convergence; the
    // inner loop is preconditioned CG on A e = r in the storage precision.  In fp64 mode the
    // inner CG updates x in place and the outer loop only verifies (normally one pass).
    //
    // Two kernel sets: the tile path (Laplacian + multigrid, fine.cuh: fused V-cycle legs, tile
    // skipping) and the flat path (bilaplacian, or Jacobi on request).
    // ---- slab mode: failure handling -----------------------------------------------------------------
    // A rank that fails must not leave its peers inside a collective for ever.  (1) Host syncs of the solve
    // poll the stream and the communicator: an asynchronous NCCL error, or HGRID_COMM_TIMEOUT seconds (default
    // 60; an iteration takes milliseconds) without progress, aborts the communicator and returns HG_ERR_COMM.
    // (2) Any hard failure of a solve in slab mode aborts the communicator, which releases the peers' NCCL
    // kernels; they then fail the same way within their own timeout.  A dead handle refuses further solves.
    bool comm_dead = false;
    CommKey comm_key = {0, 0, 0};
    void abort_comm();
    int sync_watchful();
    int solve_device(double rtol, int maxiter, hg_stats* st) override;
    int solve_device_impl(double rtol, int maxiter, hg_stats* st);
    // (L^m + w A^T A) x = L^(m-1) g + w A^T c of poisson.py (solver_composed.cu)
    int solve_composed(const hg_linear_constraints* lc, double rtol, int maxiter, hg_stats* st) override;
    int upload_u8(const uint8_t* vals, hg_stats* st) override;
    int download_u8(uint8_t* out, hg_stats* st) override;
    int reset_guess() override;
    int download(double* x, hg_stats* st) override;
    int bench_apply(int iters, double* ms_out) override;
};

}  // namespace hg
