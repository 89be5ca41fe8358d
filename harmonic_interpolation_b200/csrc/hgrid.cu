// hgrid.cu -- solver handle, PCG driver and the C ABI of libhgrid.so (see include/hgrid.h).
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
#include "pcg.cuh"
#include "stencil.cuh"

namespace hg {

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
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
static Nccl g_nccl;

// Communicators are kept for the life of the process, one per (world, rank, device): creating one costs about a second
// (bootstrap, channel set-up on the first send / recv to every peer), which would otherwise be paid by every solver that
// is constructed -- a one-shot solve of a 16384^2 slab takes a tenth of that.  All ranks construct their solvers in the same
// order, so either every rank finds its communicator here or none does.  An aborted communicator leaves the cache.
struct CommKey { int world, rank, dev; bool operator<(const CommKey& o) const { return world != o.world ? world < o.world : (rank != o.rank ? rank < o.rank : dev < o.dev); } };
static std::map<CommKey, ncclComm_t> g_comms;
static std::mutex g_comms_mu;

static int nccl_load() {
    if (g_nccl.ok) return HG_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail(HG_ERR_INVALID, std::string("multi-GPU mode needs NCCL: ") + dlerror());
#define HG_SYM(field, name)                                                           \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                               \
    if (!g_nccl.field) return fail(HG_ERR_INVALID, std::string("NCCL symbol missing: ") + name)
    HG_SYM(GetUniqueId, "ncclGetUniqueId");
    HG_SYM(CommInitRank, "ncclCommInitRank");
    HG_SYM(CommDestroy, "ncclCommDestroy");
    HG_SYM(CommAbort, "ncclCommAbort");
    HG_SYM(CommGetAsyncError, "ncclCommGetAsyncError");
    HG_SYM(AllReduce, "ncclAllReduce");
    HG_SYM(Send, "ncclSend");
    HG_SYM(Recv, "ncclRecv");
    HG_SYM(Broadcast, "ncclBroadcast");
    HG_SYM(AllGather, "ncclAllGather");
    HG_SYM(GroupStart, "ncclGroupStart");
    HG_SYM(GroupEnd, "ncclGroupEnd");
    HG_SYM(GetErrorString, "ncclGetErrorString");
#undef HG_SYM
    g_nccl.ok = true;
    return HG_OK;
}

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
    virtual int solve_device(double rtol, int maxiter, hg_stats* st) = 0;
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
    void mark(int id) {
        if (!phase_timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, stream);
        phase_ev.push_back(e);
        phase_id.push_back(id);
    }
    void report_phases() {
        for (int k = 0; k < HG_N_PHASES; ++k) { phase_tot[k] = 0.0; phase_cnt[k] = 0; }
        if (!phase_timing || phase_ev.empty()) return;
        for (size_t i = 1; i < phase_ev.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, phase_ev[i - 1], phase_ev[i]);
            phase_tot[phase_id[i]] += ms;
            phase_cnt[phase_id[i]] += 1;
        }
        if (phase_print) {
            static const char* names[HG_N_PHASES] = {"(other)", "stencil + alpha", "x/r update + check", "fine down", "level-1 down", "levels >= 2",
                                                     "level-1 up", "fine up", "beta + p update", "outer residual"};
            fprintf(stderr, "[hgrid rank %d] phases (ms):", rank);
            for (int k = 1; k < HG_N_PHASES; ++k) fprintf(stderr, "  %s %.2f (%d) |", names[k], phase_tot[k], phase_cnt[k]);
            fprintf(stderr, "\n");
        }
        for (auto e : phase_ev) cudaEventDestroy(e);
        phase_ev.clear(); phase_id.clear();
    }
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
    // HGRID_OVERLAP=1 (slab mode; written after round 1's GPU budget was spent, NOT YET RUN on a GPU -- off by default):
    // the 128-row exchange of r at the head of the V-cycle travels on a second stream while the downward leg works
    // on the chunks that read no ghost row; the chunks next to the ghost bands follow once the rows have arrived.
    bool overlap = getenv("HGRID_OVERLAP") != nullptr;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_arrived = nullptr;
    int sl_n_int[2] = {0, 0};          // down lists (lean, general): number of interior tasks, which come first
    int sl_ns = 0, sl_nt = 0;
    int* d_sl_tasks = nullptr;
    int* d_task_counts = nullptr;
    bool task_counts_pending = false;
    int finish_task_lists() {      // after a stream sync: the list lengths the launches need
        if (!task_counts_pending) return HG_OK;
        int c[8];
        CU(cudaMemcpy(c, d_task_counts, sizeof c, cudaMemcpyDeviceToHost));
        sl_n_down_lean = c[0]; sl_n_down_gen = c[1]; sl_n_up_lean = c[2]; sl_n_up_gen = c[3];
        sl_n_int[0] = c[4]; sl_n_int[1] = c[5];
        task_counts_pending = false;
        return HG_OK;
    }
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

    ~Solver() override {
        dfree(d_flags); dfree(d_soft); dfree(d_dinv); dfree(d_raw);
        dfree(d_x); dfree(d_b); dfree(d_r); dfree(d_p); dfree(d_q);
        for (auto& p : d_stage) dfree(p);
        dfree(d_out); dfree(d_part0); dfree(d_part1); dfree(d_state); dfree(d_presum); dfree(d_counts);
        free_hierarchy();
        dfree(d_z);
        if (kRefine) { dfree(d_xm); dfree(d_bm); }
        dfree(d_tile_active); dfree(d_tpart0); dfree(d_tpart1);
        dfree(d_down_active); dfree(d_code); dfree(d_lpart);
        dfree(d_sl_up); dfree(d_sl_down); dfree(d_spart); dfree(d_spart2); dfree(d_sl_tasks); dfree(d_task_counts);
        
        dfree(d_gl_rows); dfree(d_gsend);
        // (the communicator stays in the process-wide cache)
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream2) cudaStreamDestroy(stream2);
        if (ev_ready) cudaEventDestroy(ev_ready);
        if (ev_arrived) cudaEventDestroy(ev_arrived);
        if (stream) cudaStreamDestroy(stream);
    }

    int init() {
        {
            int dev = 0;
            CU(cudaGetDevice(&dev));
            cudaMemPool_t pool = nullptr;
            CU(cudaDeviceGetDefaultMemPool(&pool, dev));
            unsigned long long keep = ~0ULL;
            CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&ev0));
        CU(cudaEventCreate(&ev1));
        if (overlap) {
            CU(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ev_arrived, cudaEventDisableTiming));
        }
        CU(dmalloc(&d_flags, plane));
        CU(dmalloc(&d_state, sizeof(CgState)));
        CU(dmalloc(&d_presum, (size_t)HG_MAXK * PRESUM_BLOCKS * sizeof(double)));
        CU(dmalloc(&d_counts, 8 * sizeof(unsigned long long)));
        return HG_OK;
    }

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
    int set_flags(const uint8_t* f) override {
        if (!f) return fail(HG_ERR_INVALID, "hg_set_flags: NULL flags");
        // one contiguous upload; padding (hard) and the sanitising -- edge bits that point outside the grid are dropped -- on the device
        const size_t n = (size_t)rows * cols;
        if (!d_raw) CU(dmalloc(&d_raw, n * (size_t)std::max(K, 1)));
        CU(cudaMemcpyAsync(d_raw, f, n, cudaMemcpyHostToDevice, stream));
        pad_flags_kernel<<<dim3((pitch + 255) / 256, rows), 256, 0, stream>>>(d_raw, d_flags, rows, cols, pitch);
        CU(cudaStreamSynchronize(stream));       // (the caller may reuse its buffer)
        flags_set = true;
        is_setup = false;
        return HG_OK;
    }

    int set_soft(const double* w, double scalar) override {
        soft_scalar = scalar;
        if (!w) {
            dfree(d_soft);
            d_soft = nullptr;
            soft_plane_set = false;
        } else {
            std::vector<T> tmp((size_t)plane, T(0));
            for (int i = 0; i < rows; ++i)
                for (int j = 0; j < cols; ++j) tmp[(size_t)i * pitch + j] = (T)w[(size_t)i * cols + j];
            if (!d_soft) CU(dmalloc(&d_soft, plane * sizeof(T)));
            CU(cudaMemcpyAsync(d_soft, tmp.data(), plane * sizeof(T), cudaMemcpyHostToDevice, stream));
            CU(cudaStreamSynchronize(stream));
            soft_plane_set = true;
        }
        is_setup = false;
        return HG_OK;
    }

    int set_grad_weight(double w) override {
        w_g = w;
        is_setup = false;
        return HG_OK;
    }

    int set_edges(int mode) override {
        if (mode != HG_EDGES_REFERENCE && mode != HG_EDGES_UNIFORM) return fail(HG_ERR_INVALID, "unknown edge-weight mode");
        if (mode == HG_EDGES_UNIFORM && op != HG_OP_LAPLACIAN) return fail(HG_ERR_INVALID, "uniform edge weights apply to the Laplacian operator only");
        edge_uniform = mode;
        is_setup = false;
        return HG_OK;
    }
    int set_precond(int p) override {
        if (p != HG_PRECOND_JACOBI && p != HG_PRECOND_MULTIGRID) return fail(HG_ERR_INVALID, "unknown preconditioner");
        precond = p;
        is_setup = false;
        return HG_OK;
    }

    int ensure_vectors() {
        const size_t bytes = (size_t)plane * K * sizeof(T);
        T** v[] = {&d_x, &d_b, &d_r, &d_p, &d_q};
        for (T** p : v)
            if (!*p) {
                CU(dmalloc(p, bytes));
                CU(cudaMemsetAsync(*p, 0, bytes, stream));
            }
        if (!d_dinv) CU(dmalloc(&d_dinv, plane * sizeof(T)));
        if (kRefine) {
            if (!d_xm) CU(dmalloc(&d_xm, (size_t)plane * K * sizeof(double)));
            if (!d_bm) CU(dmalloc(&d_bm, (size_t)plane * K * sizeof(double)));
        } else {
            d_xm = reinterpret_cast<double*>(d_x);
            d_bm = reinterpret_cast<double*>(d_b);
        }
        const int need = std::max(std::max(op_blocks(), vec_blocks()), ((pitch + 255) / 256) * rows) * K;
        if (need > part_cap) {
            dfree(d_part0); dfree(d_part1);
            d_part0 = d_part1 = nullptr;
            CU(dmalloc(&d_part0, (size_t)need * sizeof(double)));
            CU(dmalloc(&d_part1, (size_t)need * sizeof(double)));
            part_cap = need;
        }
        return HG_OK;
    }


    // ---- multigrid ----------------------------------------------------------------------------
    void free_hierarchy() {
        for (auto* lv : {&coarse, &gcoarse}) {
            for (auto& c : *lv) { dfree(c.coef); dfree(c.x); dfree(c.x2); dfree(c.b); dfree(c.code); dfree(c.act_up); dfree(c.act_down); dfree(c.cs_act); dfree(c.refq); }
            lv->clear();
        }
        dfree(d_dense); dfree(d_gdense);
        d_dense = d_gdense = nullptr;
        for (size_t l = 0; l < lev25.size(); ++l) {
            dfree(lev25[l].coef);
            if (l > 0) { dfree(lev25[l].x); dfree(lev25[l].b); }
        }
        lev25.clear();
        dfree(d_dense25);
        d_dense25 = nullptr;
    }

    static int chain(int n, int levels) { for (int l = 0; l < levels; ++l) n = (n + 1) / 2; return n; }

    int alloc_level(Coarse<T>& L, int r, int c) {
        L.rows = r; L.cols = c;
        L.pitch = round_up(c, 32);
        L.plane = (long long)L.rows * L.pitch;
        L.K = K;
        L.coef = L.x = L.b = L.x2 = nullptr;
        L.code = L.act_up = L.act_down = nullptr;
        L.cs_act = nullptr; L.cs_cr = L.cs_ns = L.cs_nt = 0;
        L.refq = nullptr;
        L.tiles_x = (L.pitch + CF_TX - 1) / CF_TX;
        L.tiles_y = (L.rows + CF_TY - 1) / CF_TY;
        L.has_ref = 0;
        for (T& v : L.ref) v = T(0);
        CU(dmalloc(&L.code, (size_t)L.plane));
        CU(dmalloc(&L.act_up, (size_t)L.tiles_x * L.tiles_y));
        CU(dmalloc(&L.act_down, (size_t)L.tiles_x * L.tiles_y));
        CU(dmalloc(&L.coef, (size_t)L.plane * 9 * sizeof(T)));
        CU(dmalloc(&L.x, (size_t)L.plane * K * sizeof(T)));
        CU(dmalloc(&L.b, (size_t)L.plane * K * sizeof(T)));
        CU(dmalloc(&L.x2, (size_t)L.plane * K * sizeof(T)));
        CU(cudaMemsetAsync(L.x2, 0, (size_t)L.plane * K * sizeof(T), stream));
        CU(cudaMemsetAsync(L.coef, 0, (size_t)L.plane * 9 * sizeof(T), stream));
        CU(cudaMemsetAsync(L.x, 0, (size_t)L.plane * K * sizeof(T), stream));
        CU(cudaMemsetAsync(L.b, 0, (size_t)L.plane * K * sizeof(T), stream));
        return HG_OK;
    }

    // Interior stencil of every level: depends on T and the level only.  Taken from the hierarchy of
    // a 256^2 grid without constraints (centre pixel of each level; the same kernel, so bit-identical
    // to what the real hierarchy holds wherever the neighbourhood is regular).
    static constexpr int REF_LEVELS = 5;
    T ref_stencil[REF_LEVELS + 1][9];
    bool ref_ready = false;
    // (process-wide cache per storage type: the values depend on nothing else)
    struct RefCache { T st[REF_LEVELS + 1][9]; bool ready; };
    static RefCache& ref_cache() { static RefCache c = {}; return c; }
    int build_ref_stencils() {
        if (ref_ready) return HG_OK;
        if (ref_cache().ready) {
            memcpy(ref_stencil, ref_cache().st, sizeof ref_stencil);
            ref_ready = true;
            return HG_OK;
        }
        const int n = 256;
        uint8_t* f = nullptr;
        CU(dmalloc(&f, (size_t)n * n));
        cudaMemsetAsync(f, 0, (size_t)n * n, stream);
        Grid<T> g;
        g.rows = g.cols = g.pitch = n; g.plane = (long long)n * n; g.K = 1;
        g.flags = f; g.soft = nullptr; g.soft_scalar = T(0); g.w_g = T(0); g.uniform = 0;
        std::vector<Coarse<T>> lv;
        cudaError_t err = cudaSuccess;
        for (int l = 1; l <= REF_LEVELS; ++l) {
            Coarse<T> L;
            memset(&L, 0, sizeof L);
            const int m = n >> l;
            L.rows = L.cols = m; L.pitch = round_up(m, 32); L.plane = (long long)m * L.pitch; L.K = 1;
            err = dmalloc(&L.coef, (size_t)L.plane * 9 * sizeof(T));
            if (err != cudaSuccess) break;
            dim3 cg((L.pitch + 127) / 128, L.rows);
            if (l == 1) galerkin_kernel<T, FineAcc<T>><<<cg, 128, 0, stream>>>(FineAcc<T>{g}, L);
            else galerkin_kernel<T, CoarseAcc<T>><<<cg, 128, 0, stream>>>(CoarseAcc<T>{lv.back()}, L);
            lv.push_back(L);
            for (int c = 0; c < 9; ++c)
                cudaMemcpyAsync(&ref_stencil[l][c], L.coef + cidx(c, (long long)(m / 2) * L.pitch + m / 2), sizeof(T),
                                cudaMemcpyDeviceToHost, stream);
        }
        if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
        if (err == cudaSuccess) err = cudaGetLastError();
        for (auto& L : lv) dfree(L.coef);
        dfree(f);
        CU(err);
        ref_ready = true;
        memcpy(ref_cache().st, ref_stencil, sizeof ref_stencil);
        ref_cache().ready = true;
        return HG_OK;
    }

    // code bytes and tile activity of a level whose coefficients are final (`level` >= 1)
    int finish_level(Coarse<T>& L, int level) {
        L.has_ref = level <= REF_LEVELS;
        if (L.has_ref)
            for (int c = 0; c < 9; ++c) L.ref[c] = ref_stencil[level][c];
        ccode_kernel<T><<<dim3((L.pitch + 255) / 256, L.rows), 256, 0, stream>>>(L);
        ctile_activity_kernel<<<dim3(L.tiles_x, L.tiles_y), 256, 0, stream>>>(L.code, L.rows, L.cols, L.pitch, 0, L.act_up);
        ctile_activity_kernel<<<dim3(L.tiles_x, L.tiles_y), 256, 0, stream>>>(L.code, L.rows, L.cols, L.pitch, 1, L.act_down);
        // streaming legs (coarse_stream.cuh) on every level that is large enough to feed warps with 120-column strips;
        // chunks of 64 rows where that still gives every SM several blocks, 16 rows on the smaller levels
        L.cs_cr = 0;
        if (!coarse_tiles && (long long)L.rows * L.cols > 256 && L.plane < (1LL << 31)) {
            L.cs_ns = (L.pitch + CS_OUT - 1) / CS_OUT;
            const long long warps64 = (long long)L.cs_ns * ((L.rows + 63) / 64) * K;
            L.cs_cr = warps64 >= 148LL * 8 * 2 ? 64 : 16;
            L.cs_nt = (L.rows + L.cs_cr - 1) / L.cs_cr;
            CU(dmalloc(&L.cs_act, (size_t)L.cs_ns * L.cs_nt));
            cs_activity_kernel<<<dim3(L.cs_ns, L.cs_nt), 256, 0, stream>>>(L.code, L.rows, L.cols, L.pitch, L.cs_cr, L.cs_ns, L.cs_act);
            CU(dmalloc(&L.refq, 36 * sizeof(T)));
            cs_refq_kernel<T><<<1, 64, 0, stream>>>(L);
        }
        return HG_OK;
    }
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
    void coarse_down(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
        if (unfused_coarse) { coarse_pre(L, done); coarse_restrict(L, Lc, done); return; }
        if (L.cs_cr) { launch_cs<false>(L, Lc, done); ++launches; return; }
        switch (std::min(K, 4)) {
            case 1: launch_cdown<1>(L, Lc, done); break;
            case 2: launch_cdown<2>(L, Lc, done); break;
            case 3: launch_cdown<3>(L, Lc, done); break;
            default: launch_cdown<4>(L, Lc, done); break;
        }
        ++launches;
    }
    // (the fused upward legs write x2; afterwards x2 IS the level's x: the host-side structs are swapped, every later
    // launch -- the level above, the next V-cycle -- takes the struct by value and sees the current pointers)
    void coarse_up(Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
        if (unfused_coarse) { coarse_prolong(L, Lc, done); coarse_post(L, done); return; }
        if (L.cs_cr) launch_cs<true>(L, Lc, done);
        else switch (std::min(K, 4)) {
            case 1: launch_cup<1>(L, Lc, done); break;
            case 2: launch_cup<2>(L, Lc, done); break;
            case 3: launch_cup<3>(L, Lc, done); break;
            default: launch_cup<4>(L, Lc, done); break;
        }
        std::swap(L.x, L.x2);
        ++launches;
    }

    int build_dense(const Coarse<T>& L, double** out) {
        const int n = L.rows * L.cols;
        CU(dmalloc(out, (size_t)n * 2 * n * sizeof(double)));
        dense_from_stencil_kernel<T><<<64, 256, 0, stream>>>(L, *out);
        dense_fill_kernel<T><<<(n + 127) / 128, 128, 0, stream>>>(L, *out);
        gauss_jordan_kernel<<<1, 1024, 2 * n * sizeof(double), stream>>>(*out, n);
        return HG_OK;
    }

    // slab mode: rows of level HG_GATHER_LEVEL owned by rank q, in the replicated global level
    int gl_start(int q) const { return row_starts[q] >> HG_GATHER_LEVEL; }
    int gl_end(int q) const { return q == world - 1 ? chain(global_rows, HG_GATHER_LEVEL) : row_starts[q + 1] >> HG_GATHER_LEVEL; }

    // every rank's owned rows of `nplanes` planes of the local gather level -> the global level: packed into one
    // equal-sized block per rank, ONE all-gather, unpacked (a broadcast per plane and rank cost 24 collectives per
    // V-cycle on 8 GPUs)
    T *d_gsend = nullptr, *d_grecv = nullptr;
    size_t gbuf_elems = 0;
    int gather_level(const T* local, long long lplane, T* global, long long gplane, int pitch_, int nplanes) {
        const ncclDataType_t dt = sizeof(T) == 8 ? ncclDouble : ncclFloat;
        const int gtL = ghost_top >> HG_GATHER_LEVEL;
        int maxrows = 0;
        for (int q = 0; q < world; ++q) maxrows = std::max(maxrows, gl_end(q) - gl_start(q));
        const size_t block = (size_t)nplanes * maxrows * pitch_;
        if (block * (world + 1) > gbuf_elems) {
            dfree(d_gsend);
            d_gsend = nullptr;
            gbuf_elems = block * (world + 1);
            CU(dmalloc(&d_gsend, gbuf_elems * sizeof(T)));
            d_grecv = d_gsend + block;
        } else {
            d_grecv = d_gsend + block;
        }
        const int myrows = gl_end(rank) - gl_start(rank);
        for (int c = 0; c < nplanes; ++c)
            CU(cudaMemcpyAsync(d_gsend + (size_t)c * maxrows * pitch_, local + (long long)c * lplane + (long long)gtL * pitch_,
                               (size_t)myrows * pitch_ * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        NC(g_nccl.AllGather(d_gsend, d_grecv, block, dt, comm, stream));
        gather_unpack_kernel<T><<<dim3((pitch_ + 255) / 256, maxrows, world * nplanes), 256, 0, stream>>>(
            d_grecv, global, gplane, pitch_, maxrows, nplanes, d_gl_rows);
        return HG_OK;
    }

    int build_global_hierarchy() {
        {
            std::vector<int> gl(2 * world);
            for (int q = 0; q < world; ++q) { gl[2 * q] = gl_start(q); gl[2 * q + 1] = gl_end(q) - gl_start(q); }
            if (!d_gl_rows) CU(dmalloc(&d_gl_rows, 2 * world * sizeof(int)));
            CU(cudaMemcpyAsync(d_gl_rows, gl.data(), 2 * world * sizeof(int), cudaMemcpyHostToDevice, stream));
            HGTRY(sync_watchful());
        }
        const Coarse<T>& Ll = coarse.back();      // local level HG_GATHER_LEVEL
        Coarse<T> G;
        HGTRY(alloc_level(G, chain(global_rows, HG_GATHER_LEVEL), Ll.cols));
        if (G.pitch != Ll.pitch) return fail(HG_ERR_INVALID, "slab mode: pitch mismatch at the gather level");
        HGTRY(gather_level(Ll.coef, 0, G.coef, 0, 9 * G.pitch, 1));     // (quad-interleaved: a row of the level is 9 * pitch contiguous values)
        HGTRY(finish_level(G, HG_GATHER_LEVEL));
        gcoarse.push_back(G);
        int r = G.rows, c = G.cols;
        while ((long long)r * c > 256) {
            Coarse<T> L;
            HGTRY(alloc_level(L, (r + 1) / 2, (c + 1) / 2));
            dim3 cg((L.pitch + 127) / 128, L.rows);
            CoarseAcc<T> ca{gcoarse.back()};
            galerkin_kernel<T, CoarseAcc<T>><<<cg, 128, 0, stream>>>(ca, L);
            HGTRY(finish_level(L, HG_GATHER_LEVEL + (int)gcoarse.size()));
            gcoarse.push_back(L);
            r = L.rows; c = L.cols;
        }
        HGTRY(build_dense(gcoarse.back(), &d_gdense));
        return HG_OK;
    }

    bool use_mg() const { return precond == HG_PRECOND_MULTIGRID && op == HG_OP_LAPLACIAN; }
    bool use_mg25() const { return precond == HG_PRECOND_MULTIGRID && op == HG_OP_BILAPLACIAN; }

    // ---- bilaplacian hierarchy --------------------------------------------------------------------
    int build_hierarchy25() {
        free_hierarchy();
        if (!d_z) {
            CU(dmalloc(&d_z, (size_t)plane * K * sizeof(T)));
            CU(cudaMemsetAsync(d_z, 0, (size_t)plane * K * sizeof(T), stream));
        }
        int r = rows, c = cols;
        for (int level = 0;; ++level) {
            Lev25<T> L;
            L.rows = r; L.cols = c;
            L.pitch = level == 0 ? pitch : round_up(c, 32);
            L.plane = (long long)L.rows * L.pitch;
            L.K = K;
            L.coef = L.x = L.b = nullptr;
            CU(dmalloc(&L.coef, (size_t)L.plane * 25 * sizeof(T)));
            if (level > 0) {
                CU(dmalloc(&L.x, (size_t)L.plane * K * sizeof(T)));
                CU(dmalloc(&L.b, (size_t)L.plane * K * sizeof(T)));
                CU(cudaMemsetAsync(L.x, 0, (size_t)L.plane * K * sizeof(T), stream));
                CU(cudaMemsetAsync(L.b, 0, (size_t)L.plane * K * sizeof(T), stream));
            }
            dim3 cg((L.pitch + 127) / 128, L.rows);
            if (level == 0) assemble25_kernel<T><<<cg, 128, 0, stream>>>(grid(), L);
            else galerkin25_kernel<T><<<cg, 128, 0, stream>>>(Acc25<T>{lev25.back()}, L);
            lev25.push_back(L);
            if ((long long)r * c <= 256) break;
            r = (r + 1) / 2; c = (c + 1) / 2;
        }
        const Lev25<T>& L = lev25.back();
        const int n = L.rows * L.cols;
        CU(dmalloc(&d_dense25, (size_t)n * 2 * n * sizeof(double)));
        CU(cudaMemsetAsync(d_dense25, 0, (size_t)n * 2 * n * sizeof(double), stream));
        dense_fill25_kernel<T><<<(n + 127) / 128, 128, 0, stream>>>(L, d_dense25);
        gauss_jordan_kernel<<<1, 1024, 2 * n * sizeof(double), stream>>>(d_dense25, n);
        CU(cudaGetLastError());
        return HG_OK;
    }
    void gs25(Lev25<T>& L, bool forward, const int* done) {
        dim3 gg(((L.cols + 2) / 3 + 127) / 128, (L.rows + 2) / 3);
        for (int s = 0; s < mg25_sweeps; ++s)
            for (int c = 0; c < 9; ++c) gs25_kernel<T><<<gg, 128, 0, stream>>>(L, forward ? c : 8 - c, done);
        launches += 9 * mg25_sweeps;
    }
    // z = M^-1 r: one symmetric V-cycle from the zero guess on the 25-point hierarchy (level 0 works in place on z, r)
    int vcycle25(T* z, T* r, const int* done) {
        lev25[0].x = z; lev25[0].b = r;
        CU(cudaMemsetAsync(z, 0, (size_t)plane * K * sizeof(T), stream));
        const int n = (int)lev25.size();
        for (int l = 0; l + 1 < n; ++l) {
            gs25(lev25[l], true, done);
            const Lev25<T>& Lc = lev25[l + 1];
            resid_restrict25_kernel<T><<<dim3((Lc.pitch + 127) / 128, Lc.rows), 128, 0, stream>>>(lev25[l], Lc, done);
        }
        const Lev25<T>& Lb = lev25[n - 1];
        coarsest_dense25_kernel<T><<<K, 256, Lb.rows * Lb.cols * sizeof(double), stream>>>(Lb, d_dense25, done);
        for (int l = n - 2; l >= 0; --l) {
            prolong_add25_kernel<T><<<dim3((lev25[l].cols + 255) / 256, lev25[l].rows), 256, 0, stream>>>(lev25[l], lev25[l + 1], done);
            gs25(lev25[l], false, done);
        }
        launches += 2 * (n - 1) + 2;
        return HG_OK;
    }

    dim3 tile_grid() const { return dim3(tiles_x, tiles_y); }

    int ntiles() const { return tiles_x * tiles_y; }

    int build_tiles() {
        tiles_x = (pitch + FT_X - 1) / FT_X;
        tiles_y = (rows + FT_Y - 1) / FT_Y;
        dfree(d_tile_active); dfree(d_tpart0); dfree(d_tpart1);
        d_tile_active = nullptr; d_tpart0 = d_tpart1 = nullptr;
        CU(dmalloc(&d_tile_active, ntiles()));
        CU(dmalloc(&d_tpart0, (size_t)ntiles() * K * sizeof(double)));
        CU(dmalloc(&d_tpart1, (size_t)ntiles() * K * sizeof(double)));
        // inactive tiles never write their partial sums: they must read as zero
        CU(cudaMemsetAsync(d_tpart0, 0, (size_t)ntiles() * K * sizeof(double), stream));
        CU(cudaMemsetAsync(d_tpart1, 0, (size_t)ntiles() * K * sizeof(double), stream));
        dfree(d_down_active); dfree(d_code); dfree(d_lpart);
        d_down_active = d_code = nullptr; d_lpart = nullptr;
        CU(dmalloc(&d_down_active, ntiles()));
        CU(dmalloc(&d_code, plane));
        CU(dmalloc(&d_lpart, (size_t)n_lap() * K * sizeof(double)));
        CU(cudaMemsetAsync(d_lpart, 0, (size_t)n_lap() * K * sizeof(double), stream));
        tile_activity_kernel<<<tile_grid(), 256, 0, stream>>>(d_flags, rows, cols, pitch, 0, d_tile_active);
        tile_activity_kernel<<<tile_grid(), 256, 0, stream>>>(d_flags, rows, cols, pitch, 1, d_down_active);
        opcode_kernel<T><<<dim3((pitch + 255) / 256, rows), 256, 0, stream>>>(grid(), d_code);
        sl_ns = (pitch + SL_OUT - 1) / SL_OUT;
        sl_nt = (rows + SL_R - 1) / SL_R;
        dfree(d_sl_up); dfree(d_sl_down); dfree(d_spart);
        d_sl_up = d_sl_down = nullptr; d_spart = nullptr;
        CU(dmalloc(&d_sl_up, (size_t)sl_ns * sl_nt));
        CU(dmalloc(&d_sl_down, (size_t)sl_ns * sl_nt));
        CU(dmalloc(&d_spart, (size_t)sl_ns * sl_nt * K * sizeof(double)));
        CU(cudaMemsetAsync(d_spart, 0, (size_t)sl_ns * sl_nt * K * sizeof(double), stream));
        dfree(d_spart2);
        d_spart2 = nullptr;
        CU(dmalloc(&d_spart2, (size_t)sl_ns * sl_nt * K * sizeof(double)));
        CU(cudaMemsetAsync(d_spart2, 0, (size_t)sl_ns * sl_nt * K * sizeof(double), stream));
        sl_activity_kernel<<<dim3(sl_ns, sl_nt), 256, 0, stream>>>(d_flags, d_code, rows, cols, pitch, 0, sl_ns, d_sl_up);
        sl_activity_kernel<<<dim3(sl_ns, sl_nt), 256, 0, stream>>>(d_flags, d_code, rows, cols, pitch, 1, sl_ns, d_sl_down);
        {   // task lists: [down lean | down general | up lean | up general], in row-major chunk order, compacted on the
            // device; the counts come back with the other counters at the end of hg_setup (finish_task_lists)
            const int n = sl_ns * sl_nt;
            dfree(d_sl_tasks);
            d_sl_tasks = nullptr;
            CU(dmalloc(&d_sl_tasks, (size_t)4 * n * sizeof(int)));
            if (!d_task_counts) CU(dmalloc(&d_task_counts, 8 * sizeof(int)));
            const int split = overlap && world > 1;
            sl_task_lists_kernel<<<4, 1024, 0, stream>>>(d_sl_up, d_sl_down, n, sl_ns, rows, ghost_top, ghost_bottom, split, d_sl_tasks, d_task_counts);
            for (int which = 0; which < 4; ++which) sl_off[which] = which * n;
            task_counts_pending = true;
        }
        CU(cudaFuncSetAttribute(mg_down0_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RegionSmem<T, DOWN_H, false>)));
        CU(cudaFuncSetAttribute(mg_down0_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RegionSmem<T, DOWN_H, true>)));
        CU(cudaFuncSetAttribute(mg_up0_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RegionSmem<T, UP_H, false>)));
        CU(cudaFuncSetAttribute(mg_up0_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RegionSmem<T, UP_H, true>)));
        CU(cudaFuncSetAttribute(sl_down_kernel<T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SL_WPB * SL_RING * 128 * sizeof(T))));
        CU(cudaFuncSetAttribute(sl_up_kernel<T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SL_WPB * SL_RING * 128 * sizeof(T))));
        CU(cudaFuncSetAttribute(sl_lap_kernel<T, T, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SL_WPB * SL_RING * 128 * sizeof(T))));
        CU(cudaFuncSetAttribute(sl_lap_kernel<T, T, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SL_WPB * SL_RING * 128 * sizeof(T))));
        CU(cudaFuncSetAttribute(sl_lap_kernel<double, T, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SL_WPB * SL_RING * 128 * sizeof(double))));
        CU(cudaFuncSetAttribute(sl_lap_kernel<double, T, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SL_WPB * SL_RING * 128 * sizeof(double))));
        CU(cudaFuncSetAttribute(mg_down0v_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VRegionSmem<T>)));
        CU(cudaFuncSetAttribute(mg_up0v_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VRegionSmem<T>)));
        return HG_OK;
    }

    int build_hierarchy() {
        free_hierarchy();
        HGTRY(build_ref_stencils());
        HGTRY(set_fused_attrs<1>()); HGTRY(set_fused_attrs<2>()); HGTRY(set_fused_attrs<3>()); HGTRY(set_fused_attrs<4>());
        HGTRY(build_tiles());
        if (!d_z) {
            CU(dmalloc(&d_z, (size_t)plane * K * sizeof(T)));
            CU(cudaMemsetAsync(d_z, 0, (size_t)plane * K * sizeof(T), stream));
        }
        int r = rows, c = cols, level = 0;
        while ((long long)r * c > 256 && (world == 1 || level < HG_GATHER_LEVEL)) {
            Coarse<T> L;
            HGTRY(alloc_level(L, (r + 1) / 2, (c + 1) / 2));
            dim3 cg((L.pitch + 127) / 128, L.rows);
            if (coarse.empty()) {
                FineAcc<T> fa{grid()};
                galerkin_kernel<T, FineAcc<T>><<<cg, 128, 0, stream>>>(fa, L);
            } else {
                CoarseAcc<T> ca{coarse.back()};
                galerkin_kernel<T, CoarseAcc<T>><<<cg, 128, 0, stream>>>(ca, L);
            }
            HGTRY(finish_level(L, level + 1));
            coarse.push_back(L);
            r = L.rows; c = L.cols;
            ++level;
        }
        if (world > 1) {
            if (level < HG_GATHER_LEVEL) return fail(HG_ERR_INVALID, "slab mode: the grid is too small (fewer than HG_GATHER_LEVEL coarse levels)");
            HGTRY(build_global_hierarchy());
        } else if (!coarse.empty()) {
            HGTRY(build_dense(coarse.back(), &d_dense));
        }
        CU(cudaGetLastError());
        return HG_OK;
    }

    void coarse_pre(const Coarse<T>& L, const int* done) {
        dim3 mg((L.cols / 2 + 1 + 255) / 256, (L.rows + 1) / 2);
        for (int c = 0; c < 4; ++c) mcgs_kernel<T><<<mg, 256, 0, stream>>>(L, c, done);
        launches += 4;
    }
    void coarse_post(const Coarse<T>& L, const int* done) {
        dim3 mg((L.cols / 2 + 1 + 255) / 256, (L.rows + 1) / 2);
        for (int c = 3; c >= 0; --c) mcgs_kernel<T><<<mg, 256, 0, stream>>>(L, c, done);
        launches += 4;
    }
    void coarse_restrict(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
        dim3 gg((Lc.pitch + 255) / 256, Lc.rows);
        CoarseAcc<T> ca{L};
        resid_restrict_kernel<T, CoarseAcc<T>><<<gg, 256, 0, stream>>>(ca, L.x, L.b, Lc, K, done);
        ++launches;
    }
    void coarse_prolong(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
        CoarseAcc<T> ca{L};
        dim3 gg((L.cols + 255) / 256, L.rows);
        prolong_add_kernel<T, CoarseAcc<T>><<<gg, 256, 0, stream>>>(ca, L.x, Lc, K, done);
        ++launches;
    }
    void coarse_solve(const Coarse<T>& L, const double* dense, const int* done) {
        const int n = L.rows * L.cols;
        coarsest_dense_kernel<T><<<K, 256, n * sizeof(double), stream>>>(L, dense, done);
        ++launches;
    }
    // V-cycle over levels[from..] of a hierarchy whose first level already holds b (and x = 0)
    void coarse_vcycle(std::vector<Coarse<T>>& lv, const double* dense, const int* done) {
        const int n = (int)lv.size();
        const bool top = &lv == &coarse;      // (marks: level 1 of the local hierarchy on its own)
        for (int l = 0; l + 1 < n; ++l) {
            coarse_down(lv[l], lv[l + 1], done);
            if (top && l == 0) mark(HG_PH_L1_DOWN);
        }
        coarse_solve(lv[n - 1], dense, done);
        for (int l = n - 2; l >= 0; --l) {
            if (top && l == 0) mark(HG_PH_COARSE);
            coarse_up(lv[l], lv[l + 1], done);
            if (top && l == 0) mark(HG_PH_L1_UP);
        }
    }

    // z = M^-1 r : one V(1,1) cycle from a zero initial guess; also leaves the per-tile partial
    // sums of r.z in part_rz.  Level 0 is two fused tile kernels (fine.cuh).  In slab mode the
    // ghost bands are refreshed once per level and leg and levels >= HG_GATHER_LEVEL run replicated.
    // Experiment aid (HGRID_GENERIC_VCYCLE=nu): a V(nu,nu) cycle assembled from the separate, unfused kernels
    // of mg.cuh on every level -- slow, but any nu; used to measure what more smoothing buys in iterations.
    int generic_nu = getenv("HGRID_GENERIC_VCYCLE") ? atoi(getenv("HGRID_GENERIC_VCYCLE")) : 0;
    int vcycle_generic(T* z, T* r, const int* done) {
        const Grid<T> g = grid();
        FineAcc<T> fa{g};
        const int nl = (int)coarse.size(), nu = generic_nu;
        CU(cudaMemsetAsync(z, 0, (size_t)plane * K * sizeof(T), stream));
        dim3 rb((cols / 2 + 1 + 255) / 256, rows);
        for (int s = 0; s < nu; ++s)
            for (int c = 0; c < 2; ++c) rbgs0_kernel<T><<<rb, 256, 0, stream>>>(fa, z, r, c, done);
        if (nl > 0) {
            resid_restrict_kernel<T, FineAcc<T>><<<dim3((coarse[0].pitch + 255) / 256, coarse[0].rows), 256, 0, stream>>>(fa, z, r, coarse[0], K, done);
            for (int l = 0; l + 1 < nl; ++l) {
                for (int s = 0; s < nu; ++s) coarse_pre(coarse[l], done);
                coarse_restrict(coarse[l], coarse[l + 1], done);
            }
            coarse_solve(coarse[nl - 1], d_dense, done);
            for (int l = nl - 2; l >= 0; --l) {
                coarse_prolong(coarse[l], coarse[l + 1], done);
                for (int s = 0; s < nu; ++s) coarse_post(coarse[l], done);
            }
            prolong_add_kernel<T, FineAcc<T>><<<dim3((cols + 255) / 256, rows), 256, 0, stream>>>(fa, z, coarse[0], K, done);
        }
        for (int s = 0; s < nu; ++s)
            for (int c = 1; c >= 0; --c) rbgs0_kernel<T><<<rb, 256, 0, stream>>>(fa, z, r, c, done);
        cg_dot_kernel<T><<<vec_blocks(), VEC_THREADS, 0, stream>>>(nullptr, plane, K, r, z, d_part1);
        return HG_OK;
    }

    int vcycle(T* z, T* r, const int* done, double* part_rz) {
        if (generic_nu > 0 && world == 1) return vcycle_generic(z, r, done);
        const Grid<T> g = grid();
        const int nl = (int)coarse.size();
        if (nl > 0) {
            if (world > 1 && overlap && stream_legs()) {
                CU(cudaEventRecord(ev_ready, stream));               // r is final on the owned rows
                CU(cudaStreamWaitEvent(stream2, ev_ready, 0));
                HGTRY(exchange(r, HG_GHOST_ROWS, stream2));
                CU(cudaEventRecord(ev_arrived, stream2));
                launch_sl_down(g, r, coarse[0], done, 1);            // chunks that read no ghost row
                CU(cudaStreamWaitEvent(stream, ev_arrived, 0));
                launch_sl_down(g, r, coarse[0], done, 2);            // chunks next to the ghost bands
                mark(HG_PH_FINE_DOWN);
            } else {
            if (world > 1) HGTRY(exchange(r, HG_GHOST_ROWS));
            if (stream_legs()) { launch_sl_down(g, r, coarse[0], done); mark(HG_PH_FINE_DOWN); }
            else if (has_pen) mg_down0_kernel<T, true><<<tile_grid(), FT_THREADS, sizeof(RegionSmem<T, DOWN_H, true>), stream>>>(g, d_code, r, coarse[0], d_down_active, done);
            else if (scalar_legs) mg_down0_kernel<T, false><<<tile_grid(), FT_THREADS, sizeof(RegionSmem<T, DOWN_H, false>), stream>>>(g, d_code, r, coarse[0], d_down_active, done);
            else mg_down0v_kernel<T><<<tile_grid(), VR_THREADS, sizeof(VRegionSmem<T>), stream>>>(g, d_code, r, coarse[0], d_down_active, done);
            }
            ++launches;
            if (world == 1) {
                coarse_vcycle(coarse, d_dense, done);
            } else {
                const int Lg = HG_GATHER_LEVEL;
                for (int l = 0; l + 1 < Lg; ++l) {       // local levels 1 .. Lg-1
                    if (exchange_all) HGTRY(exchange_level(coarse[l], l + 1, coarse[l].b, K));
                    coarse_down(coarse[l], coarse[l + 1], done);
                    if (l == 0) mark(HG_PH_L1_DOWN);
                }
                // level Lg: gather the owned rows of b, solve replicated, take back the local window of x
                const Coarse<T>& Ll = coarse[Lg - 1];
                const Coarse<T>& G = gcoarse[0];
                CU(cudaMemsetAsync(G.x, 0, (size_t)G.plane * K * sizeof(T), stream));
                HGTRY(gather_level(Ll.b, Ll.plane, G.b, G.plane, G.pitch, K));
                coarse_vcycle(gcoarse, d_gdense, done);
                const long long off = (long long)((row_starts[rank] - ghost_top) >> Lg) * G.pitch;
                for (int k = 0; k < K; ++k)
                    CU(cudaMemcpyAsync(Ll.x + (long long)k * Ll.plane, G.x + (long long)k * G.plane + off,
                                       (size_t)Ll.plane * sizeof(T), cudaMemcpyDeviceToDevice, stream));
                for (int l = Lg - 2; l >= 0; --l) {
                    if (exchange_all ? l + 2 < Lg : l + 2 == Lg - 1) HGTRY(exchange_level(coarse[l + 1], l + 2, coarse[l + 1].x, K));
                    if (l == 0) mark(HG_PH_COARSE);
                    coarse_up(coarse[l], coarse[l + 1], done);
                    if (l == 0) mark(HG_PH_L1_UP);
                }
                if (exchange_all) HGTRY(exchange_level(coarse[0], 1, coarse[0].x, K));
            }
        }
        Coarse<T> c0;
        memset(&c0, 0, sizeof c0);
        if (nl > 0) c0 = coarse[0];
        mark(HG_PH_COARSE);
        if (stream_legs()) { launch_sl_up(g, r, z, c0, nl > 0, done); mark(HG_PH_FINE_UP); }
        else if (has_pen) mg_up0_kernel<T, true><<<tile_grid(), FT_THREADS, sizeof(RegionSmem<T, UP_H, true>), stream>>>(g, d_code, r, z, c0, nl > 0, d_tile_active, part_rz, done);
        else if (scalar_legs) mg_up0_kernel<T, false><<<tile_grid(), FT_THREADS, sizeof(RegionSmem<T, UP_H, false>), stream>>>(g, d_code, r, z, c0, nl > 0, d_tile_active, part_rz, done);
        else mg_up0v_kernel<T><<<tile_grid(), VR_THREADS, sizeof(VRegionSmem<T>), stream>>>(g, d_code, r, z, c0, nl > 0, d_tile_active, part_rz, done);
        ++launches;
        return HG_OK;
    }

    // the streaming legs serve every configuration without gradient-penalty edges
    bool stream_legs() const { return !has_pen && !tile_legs && !scalar_legs && plane < (1LL << 31) - (1LL << 20); }   // 32-bit offsets on the lean paths
    bool has_soft() const { return d_soft != nullptr || soft_scalar != 0.0; }
    dim3 sl_grid(int ntasks) const { return dim3((unsigned)(((long long)ntasks * K + SL_WPB - 1) / SL_WPB)); }
    dim3 sl_block() const { return dim3(32 * SL_WPB); }
    size_t sl_smem() const { return (size_t)SL_WPB * SL_RING * 128 * sizeof(T); }   // lean kernels: one ring of rows per warp
    // part 0: every task; 1: the interior tasks (front of the lists); 2: the boundary tasks (HGRID_OVERLAP)
    void launch_sl_down(const Grid<T>& g, const T* r, const Coarse<T>& c1, const int* done, int part = 0) {
        const int zx = unfused_coarse ? 1 : 0;
        const int o_l = part == 2 ? sl_n_int[0] : 0, n_l = part == 1 ? sl_n_int[0] : sl_n_down_lean - o_l;
        const int o_g = part == 2 ? sl_n_int[1] : 0, n_g = part == 1 ? sl_n_int[1] : sl_n_down_gen - o_g;
        if (n_l > 0)
            sl_down_kernel<T, false, true><<<sl_grid(n_l), sl_block(), sl_smem(), stream>>>(g, d_code, r, c1, d_sl_tasks + sl_off[0] + o_l, n_l, sl_ns, zx, done);
        if (n_g > 0) {
            if (has_soft()) sl_down_kernel<T, true, false><<<sl_grid(n_g), sl_block(), 0, stream>>>(g, d_code, r, c1, d_sl_tasks + sl_off[1] + o_g, n_g, sl_ns, zx, done);
            else sl_down_kernel<T, false, false><<<sl_grid(n_g), sl_block(), 0, stream>>>(g, d_code, r, c1, d_sl_tasks + sl_off[1] + o_g, n_g, sl_ns, zx, done);
        }
        launches += (n_l > 0) + (n_g > 0) - (part == 2 ? 0 : 1);
    }
    void launch_sl_up(const Grid<T>& g, const T* r, T* z, const Coarse<T>& c1, int have_coarse, const int* done) {
        if (sl_n_up_lean)
            sl_up_kernel<T, false, true><<<sl_grid(sl_n_up_lean), sl_block(), sl_smem(), stream>>>(g, d_code, r, z, c1, have_coarse, d_sl_tasks + sl_off[2], sl_n_up_lean, sl_ns, sl_nt, d_spart, done);
        if (sl_n_up_gen) {
            if (has_soft()) sl_up_kernel<T, true, false><<<sl_grid(sl_n_up_gen), sl_block(), 0, stream>>>(g, d_code, r, z, c1, have_coarse, d_sl_tasks + sl_off[3], sl_n_up_gen, sl_ns, sl_nt, d_spart, done);
            else sl_up_kernel<T, false, false><<<sl_grid(sl_n_up_gen), sl_block(), 0, stream>>>(g, d_code, r, z, c1, have_coarse, d_sl_tasks + sl_off[3], sl_n_up_gen, sl_ns, sl_nt, d_spart, done);
        }
        launches += (sl_n_up_lean > 0) + (sl_n_up_gen > 0) - 1;
    }
    // q = A p with the streaming kernel (no soft weights, no gradient-penalty edges), partial sums of p.q per chunk
    bool stream_lap() const { return stream_legs() && !has_soft() && !getenv("HGRID_TILE_LAP"); }
    void launch_sl_lap(const T* p, T* q, const int* done) {
        const Grid<T> g = grid();
        if (sl_n_up_lean)
            sl_lap_kernel<T, T, false, 0><<<sl_grid(sl_n_up_lean), sl_block(), sl_smem(), stream>>>(g, d_code, p, q, d_sl_tasks + sl_off[2], sl_n_up_lean, sl_ns, sl_nt, d_spart2, done);
        if (sl_n_up_gen)
            sl_lap_kernel<T, T, true, 0><<<sl_grid(sl_n_up_gen), sl_block(), sl_smem(), stream>>>(g, d_code, p, q, d_sl_tasks + sl_off[3], sl_n_up_gen, sl_ns, sl_nt, d_spart2, done);
        launches += (sl_n_up_lean > 0) + (sl_n_up_gen > 0);
    }
    // r = -A x from the fp64 master iterate (the true residual when the right-hand side is zero), partial sums of r.r
    void launch_sl_resid(const double* x, T* r) {
        const Grid<T> g = grid();
        const size_t sm = (size_t)SL_WPB * SL_RING * 128 * sizeof(double);
        if (sl_n_up_lean)
            sl_lap_kernel<double, T, false, 1><<<sl_grid(sl_n_up_lean), sl_block(), sm, stream>>>(g, d_code, x, r, d_sl_tasks + sl_off[2], sl_n_up_lean, sl_ns, sl_nt, d_spart2, nullptr);
        if (sl_n_up_gen)
            sl_lap_kernel<double, T, true, 1><<<sl_grid(sl_n_up_gen), sl_block(), sm, stream>>>(g, d_code, x, r, d_sl_tasks + sl_off[3], sl_n_up_gen, sl_ns, sl_nt, d_spart2, nullptr);
        launches += (sl_n_up_lean > 0) + (sl_n_up_gen > 0);
    }
    // the stencil launch of the CG loop + the scalar stage that forms alpha
    int lap_and_alpha(const T* p, T* q, const int* done) {
        if (stream_lap()) {
            launch_sl_lap(p, q, done);
            const int t0 = ghost_top / SL_R, t1 = ghost_bottom ? (rows - ghost_bottom) / SL_R : sl_nt;
            return stage_alpha(d_spart2 + (long long)t0 * sl_ns, (t1 - t0) * sl_ns, (long long)sl_ns * sl_nt);
        }
        launch_lap_tile<T, 0, 1>(p, nullptr, q, d_lpart, done);
        return stage_alpha(lap_part(d_lpart), n_own_lap(), n_lap());
    }
    // scalar stage after the V-cycle: r.z from the partial sums the upward leg left (owned rows only)
    int stage_beta_rz(int init) {
        if (generic_nu > 0 && world == 1) return stage_beta(d_part1, vec_blocks(), vec_blocks(), init);
        if (stream_legs()) {
            const int t0 = ghost_top / SL_R, t1 = ghost_bottom ? (rows - ghost_bottom) / SL_R : sl_nt;
            return stage_beta(d_spart + (long long)t0 * sl_ns, (t1 - t0) * sl_ns, (long long)sl_ns * sl_nt, init);
        }
        return stage_beta(tile_part(d_tpart1), n_own_tiles(), ntiles(), init);
    }

    int mg_levels() override { return 1 + (int)coarse.size(); }

    int mg_get_level(int level, int* r_out, int* c_out, double* coef9) override {
        if (level < 1 || level > (int)coarse.size()) return fail(HG_ERR_INVALID, "hg_mg_get_level: no such level");
        const Coarse<T>& L = coarse[level - 1];
        if (r_out) *r_out = L.rows;
        if (c_out) *c_out = L.cols;
        if (coef9) {
            std::vector<T> tmp((size_t)L.plane * 9);
            CU(cudaMemcpyAsync(tmp.data(), L.coef, tmp.size() * sizeof(T), cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            for (int c = 0; c < 9; ++c)
                for (int i = 0; i < L.rows; ++i)
                    for (int j = 0; j < L.cols; ++j)
                        coef9[((size_t)i * L.cols + j) * 9 + c] = (double)tmp[(size_t)cidx(c, (long long)i * L.pitch + j)];
        }
        return HG_OK;
    }

    // x and b of hierarchy level >= 1 as the last V-cycle left them, (rows, cols, K) float64 (debugging / tests of the legs)
    int mg_get_level_vectors(int level, double* xo, double* bo) override {
        if (level < 1 || level > (int)coarse.size()) return fail(HG_ERR_INVALID, "hg_mg_get_level_vectors: no such level");
        const Coarse<T>& L = coarse[level - 1];
        std::vector<T> tmp((size_t)L.plane * K);
        for (int which = 0; which < 2; ++which) {
            double* out = which ? bo : xo;
            if (!out) continue;
            CU(cudaMemcpyAsync(tmp.data(), which ? L.b : L.x, tmp.size() * sizeof(T), cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            for (int k = 0; k < K; ++k)
                for (int i = 0; i < L.rows; ++i)
                    for (int j = 0; j < L.cols; ++j)
                        out[((size_t)i * L.cols + j) * K + k] = (double)tmp[(size_t)k * L.plane + (size_t)i * L.pitch + j];
        }
        return HG_OK;
    }

    // z = M^-1 r for host arrays (parity tests of the preconditioner: symmetry, definiteness)
    int precond_apply(const double* r, double* z) override {
        if (!is_setup) return fail(HG_ERR_INVALID, "hg_precond_apply: hg_setup has not been called");
        HGTRY(stage(0, r));
        const size_t bytes = (size_t)rows * cols * K * sizeof(double);
        if (!d_out) CU(dmalloc(&d_out, bytes));
        dim3 pg((pitch + 255) / 256, rows);
        import_kernel<T><<<pg, 256, 0, stream>>>(grid(), d_stage[0], d_r);
        mask_hard_kernel<T><<<pg, 256, 0, stream>>>(grid(), d_r);
        if (use_mg()) {
            HGTRY(vcycle(d_z, d_r, nullptr, d_tpart1));
            export_kernel<T, T><<<pg, 256, 0, stream>>>(grid(), d_z, d_out);
        } else if (use_mg25()) {
            HGTRY(vcycle25(d_z, d_r, nullptr));
            export_kernel<T, T><<<pg, 256, 0, stream>>>(grid(), d_z, d_out);
        } else {
            jacobi_apply_kernel<T><<<vec_blocks(), VEC_THREADS, 0, stream>>>(plane, K, d_q, d_r, d_dinv);
            export_kernel<T, T><<<pg, 256, 0, stream>>>(grid(), d_q, d_out);
        }
        CU(cudaMemcpyAsync(z, d_out, bytes, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        have_values = false;
        return HG_OK;
    }

    // Free pixels of connected components that no value constraint touches (pcg.cuh).  One slab cannot decide this (a
    // component may be anchored in a neighbouring slab): not checked in slab mode.  HGRID_ALLOW_SINGULAR=1 skips the check.
    unsigned long long n_unconstrained = 0;
    int find_unconstrained(unsigned long long* out) {
        *out = 0;
        if (world > 1 || getenv("HGRID_ALLOW_SINGULAR")) return HG_OK;
        uint8_t* anc = nullptr;
        int* d_ch = nullptr;
        unsigned long long* d_cnt = nullptr;
        CU(dmalloc(&anc, (size_t)plane));
        CU(dmalloc(&d_ch, sizeof(int)));
        CU(dmalloc(&d_cnt, sizeof(unsigned long long)));
        dim3 pg((pitch + 255) / 256, rows), tg((cols + ANC_T - 1) / ANC_T, (rows + ANC_T - 1) / ANC_T);
        const int pen = w_g != 0.0 ? 1 : 0;
        anchor_init_kernel<<<pg, 256, 0, stream>>>(d_flags, rows, cols, pitch, pen, anc);
        const int max_rounds = 2 * ((rows + cols) / ANC_T + 2) * ANC_T;     // (a serpentine component one pixel wide)
        int changed = 1;
        for (int r = 0; r < max_rounds && changed; ++r) {
            CU(cudaMemsetAsync(d_ch, 0, sizeof(int), stream));
            anchor_spread_kernel<<<tg, 256, 0, stream>>>(d_flags, rows, cols, pitch, pen, anc, d_ch);
            CU(cudaMemcpyAsync(&changed, d_ch, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
        }
        CU(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), stream));
        anchor_count_kernel<<<pg, 256, 0, stream>>>(anc, rows, cols, pitch, d_cnt);
        CU(cudaMemcpyAsync(out, d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        dfree(anc); dfree(d_ch); dfree(d_cnt);
        return HG_OK;
    }

    int setup(hg_setup_info* info) override {
        if (!flags_set) return fail(HG_ERR_INVALID, "hg_setup: hg_set_flags has not been called");
        CU(cudaEventRecord(ev0, stream));
        HGTRY(ensure_vectors());
        CU(cudaMemsetAsync(d_counts, 0, 8 * sizeof(unsigned long long), stream));
        dim3 pg((pitch + 255) / 256, rows);
        domain_check_kernel<<<pg, 256, 0, stream>>>(d_flags, rows, cols, pitch, op == HG_OP_BILAPLACIAN, d_counts);
        diag_kernel<T><<<pg, 256, 0, stream>>>(grid(), op == HG_OP_BILAPLACIAN, d_dinv);
        has_pen = w_g != 0.0;   // refined below once the flag counts are known
        HGTRY(find_unconstrained(&n_unconstrained));
        if (use_mg()) HGTRY(build_hierarchy());
        else if (use_mg25()) HGTRY(build_hierarchy25());
        else free_hierarchy();
        unsigned long long counts[8];
        CU(cudaMemcpyAsync(counts, d_counts, sizeof counts, cudaMemcpyDeviceToHost, stream));
        CU(cudaEventRecord(ev1, stream));
        HGTRY(sync_watchful());
        HGTRY(finish_task_lists());
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev0, ev1));
        if (info) {
            info->levels = use_mg25() ? (int)lev25.size() : 1 + (int)coarse.size();
            info->empty_rows = (int)counts[0];
            info->unreferenced = (int)counts[1];
            info->n_hard = (long long)counts[2];
            info->n_free = (long long)rows * cols - (long long)counts[2];
            info->setup_ms = ms;
        }
        if (op == HG_OP_BILAPLACIAN) {
            // recovery.py:493 and :495
            const int expect = (rows == 1 || cols == 1) ? 2 : 4;
            if ((int)counts[0] != expect && counts[3] == 0)
                return fail(HG_ERR_EMPTY_ROWS, "bilaplacian: unexpected number of empty rows in the reflected Laplacian (recovery.py:493)");
            if (counts[1] != 0)
                return fail(HG_ERR_UNREFERENCED, "bilaplacian: a pixel is referenced by no row of the reflected Laplacian (recovery.py:495)");
        }
        if (info) info->unconstrained = (long long)n_unconstrained;
        if (n_unconstrained != 0) {
            char buf[256];
            snprintf(buf, sizeof buf, "singular system: %llu free pixels lie in connected components that no hard or soft value constraint touches "
                                      "(the reference's direct solver fails here, recovery.py:993)", n_unconstrained);
            return fail(HG_ERR_SINGULAR, buf);
        }
        is_setup = true;
        return HG_OK;
    }

    int stage(int slot, const double* host) {
        const size_t bytes = (size_t)rows * cols * K * sizeof(double);
        if (!host) return HG_OK;
        if (!d_stage[slot]) CU(dmalloc(&d_stage[slot], bytes));
        CU(cudaMemcpyAsync(d_stage[slot], host, bytes, cudaMemcpyHostToDevice, stream));
        return HG_OK;
    }

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
    void apply_op(const T* x, T* y) {
        if (use_mg()) {
            launch_lap_tile<T, 0, 0>(x, nullptr, y, nullptr, nullptr);
        } else {
            launch_op<0, 0>(x, nullptr, y, nullptr, nullptr);
        }
    }

    int apply(const double* x, double* y) override {
        if (!is_setup) return fail(HG_ERR_INVALID, "hg_apply: hg_setup has not been called");
        if (!x || !y) return fail(HG_ERR_INVALID, "hg_apply: NULL buffer");
        HGTRY(stage(0, x));
        const size_t bytes = (size_t)rows * cols * K * sizeof(double);
        if (!d_out) CU(dmalloc(&d_out, bytes));
        dim3 pg((pitch + 255) / 256, rows);
        import_kernel<T><<<pg, 256, 0, stream>>>(grid(), d_stage[0], d_p);
        CU(cudaMemsetAsync(d_q, 0, (size_t)plane * K * sizeof(T), stream));
        apply_op(d_p, d_q);
        export_kernel<T, T><<<pg, 256, 0, stream>>>(grid(), d_q, d_out);
        CU(cudaMemcpyAsync(y, d_out, bytes, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        have_values = false;
        return HG_OK;
    }

    int upload(const hg_values* v, hg_stats* st) override {
        if (!is_setup) return fail(HG_ERR_INVALID, "hg_upload_values: hg_setup has not been called");
        hg_values z;
        memset(&z, 0, sizeof z);
        if (!v) v = &z;
        CU(cudaEventRecord(ev0, stream));
        const double* host[6] = {v->hard_vals, v->soft_vals, v->d_down, v->d_right, v->x0, v->rhs};
        b_is_zero = !v->soft_vals && !v->d_down && !v->d_right && !v->rhs;
        for (int s = 0; s < 6; ++s) HGTRY(stage(s, host[s]));
        dim3 pg((pitch + 255) / 256, rows);
        build_problem_kernel<T, double><<<pg, 256, 0, stream>>>(grid(), host[0] ? d_stage[0] : nullptr, host[1] ? d_stage[1] : nullptr,
                                                               host[2] ? d_stage[2] : nullptr, host[3] ? d_stage[3] : nullptr,
                                                               host[4] ? d_stage[4] : nullptr, host[5] ? d_stage[5] : nullptr, d_xm, d_bm);
        CU(cudaEventRecord(ev1, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev0, ev1));
        if (st) st->upload_ms = ms;
        have_values = true;
        return HG_OK;
    }

    // Slab mode.  The local grid is [ghost band | owned rows | ghost band]; ghost bands are
    // HG_GHOST_ROWS deep MIRRORS of the neighbouring slabs (same flags, same values after a
    // refresh).  Every kernel simply runs on the whole local grid: what it computes on ghost rows
    // is redundant with the neighbour and valid up to a few rows from the outer edge, so a whole
    // smoothing sweep needs no communication and its result on the owned rows is exactly the
    // single-GPU one.  Ghost bands are refreshed (send / recv) once per level and leg; from level
    // HG_GATHER_LEVEL down the hierarchy is replicated on every rank.  Dot products run over the
    // owned tile rows only and are all-reduced.
    int comm_init(const uint8_t* id, int rank_, int world_, int gt, int gb, const int* starts, int grows) override {
        if (world_ < 1 || rank_ < 0 || rank_ >= world_) return fail(HG_ERR_INVALID, "hg_comm_init: bad rank / world");
        rank = rank_; world = world_; ghost_top = gt; ghost_bottom = gb; global_rows = grows;
        if (world == 1) return HG_OK;
        if (op != HG_OP_LAPLACIAN) return fail(HG_ERR_INVALID, "hg_comm_init: the slab mode supports the Laplacian operator only");
        if (!id || !starts) return fail(HG_ERR_INVALID, "hg_comm_init: NULL id / row_starts");
        row_starts.assign(starts, starts + world);
        row_starts.push_back(grows);
        const int want_t = rank == 0 ? 0 : HG_GHOST_ROWS, want_b = rank == world - 1 ? 0 : HG_GHOST_ROWS;
        if (gt != want_t || gb != want_b) return fail(HG_ERR_INVALID, "hg_comm_init: ghost bands must be HG_GHOST_ROWS deep towards every neighbour");
        for (int q = 0; q < world; ++q) {
            if (row_starts[q] % HG_GHOST_ROWS) return fail(HG_ERR_INVALID, "hg_comm_init: slab boundaries must be multiples of HG_GHOST_ROWS");
            if (row_starts[q + 1] - row_starts[q] < 2 * HG_GHOST_ROWS) return fail(HG_ERR_INVALID, "hg_comm_init: a slab needs at least 2 HG_GHOST_ROWS rows");
        }
        if (rows != row_starts[rank + 1] - row_starts[rank] + gt + gb) return fail(HG_ERR_INVALID, "hg_comm_init: local rows do not match the partition");
        HGTRY(nccl_load());
        int dev = 0;
        CU(cudaGetDevice(&dev));
        comm_key = CommKey{world, rank, dev};
        {
            std::lock_guard<std::mutex> lock(g_comms_mu);
            auto it = g_comms.find(comm_key);
            if (it != g_comms.end()) comm = it->second;
        }
        if (!comm) {
            ncclUniqueId uid;
            memcpy(&uid, id, sizeof uid);
            NC(g_nccl.CommInitRank(&comm, world, uid, rank));
            std::lock_guard<std::mutex> lock(g_comms_mu);
            g_comms[comm_key] = comm;
        }
        is_setup = false;
        return HG_OK;
    }

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

    int allreduce_tmp() {
        NC(g_nccl.AllReduce(d_state->tmp, d_state->tmp, K, ncclDouble, ncclSum, comm, stream));
        return HG_OK;
    }

    // the four scalar stages; with several slabs each is split around an allreduce of the local sums.
    // `part` points at the first owned partial of channel 0, `n` owned partials, `stride` per channel.
    // many partial sums: add them in chunks first (presum_kernel), the scalar stage adds the chunk sums
    void presum(const double*& part, int& n, long long& stride, const int* done) {
        if (n < PRESUM_MIN) return;
        presum_kernel<<<dim3(PRESUM_BLOCKS, K), 256, 0, stream>>>(done, part, n, stride, d_presum);
        part = d_presum; n = PRESUM_BLOCKS; stride = PRESUM_BLOCKS;
        ++launches;
    }
    int stage_alpha(const double* part, int n, long long stride) {
        presum(part, n, stride, &d_state->done);
        if (world <= 1) { cg_alpha_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, 0); return HG_OK; }
        cg_alpha_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, 1);
        HGTRY(allreduce_tmp());
        cg_alpha_kernel<<<1, 32, 0, stream>>>(d_state, part, n, stride, K, 2);
        return HG_OK;
    }
    int stage_check(const double* part, int n, long long stride) {
        presum(part, n, stride, &d_state->done);
        if (world <= 1) { cg_check_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, 0); return HG_OK; }
        cg_check_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, 1);
        HGTRY(allreduce_tmp());
        cg_check_kernel<<<1, 32, 0, stream>>>(d_state, part, n, stride, K, 2);
        return HG_OK;
    }
    int stage_beta(const double* part, int n, long long stride, int init) {
        presum(part, n, stride, &d_state->done);
        if (world <= 1) { cg_beta_rz_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, init, 0); return HG_OK; }
        cg_beta_rz_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, init, 1);
        HGTRY(allreduce_tmp());
        cg_beta_rz_kernel<<<1, 32, 0, stream>>>(d_state, part, n, stride, K, init, 2);
        return HG_OK;
    }
    int stage_outer(const double* part, int n, long long stride, int first, double red2) {
        presum(part, n, stride, nullptr);
        if (world <= 1) { cg_outer_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, first, red2, 0); return HG_OK; }
        cg_outer_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, first, red2, 1);
        HGTRY(allreduce_tmp());
        cg_outer_kernel<<<1, 32, 0, stream>>>(d_state, part, n, stride, K, first, red2, 2);
        return HG_OK;
    }

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
    void abort_comm() {
        if (comm) {
            std::lock_guard<std::mutex> lock(g_comms_mu);
            g_comms.erase(comm_key);
        }
        if (comm && g_nccl.ok) g_nccl.CommAbort(comm);
        comm = nullptr;
        comm_dead = true;
    }
    int sync_watchful() {
        if (world <= 1 || !comm) { CU(cudaStreamSynchronize(stream)); return HG_OK; }
        static const double limit = getenv("HGRID_COMM_TIMEOUT") ? atof(getenv("HGRID_COMM_TIMEOUT")) : 60.0;
        timespec t0;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (long spin = 0;; ++spin) {
            const cudaError_t q = cudaStreamQuery(stream);
            if (q == cudaSuccess) return HG_OK;
            if (q != cudaErrorNotReady) CU(q);
            if ((spin & 1023) == 1023) {
                ncclResult_t ar = ncclSuccess;
                if (g_nccl.CommGetAsyncError(comm, &ar) != ncclSuccess || (ar != ncclSuccess && ar != ncclInProgress))
                    return fail(HG_ERR_COMM, std::string("slab mode: asynchronous NCCL error: ") + g_nccl.GetErrorString(ar));
                timespec t1;
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > limit)
                    return fail(HG_ERR_COMM, "slab mode: no progress within HGRID_COMM_TIMEOUT seconds (a peer has failed or left the solve)");
            }
        }
    }
    int solve_device(double rtol, int maxiter, hg_stats* st) override {
        if (comm_dead) return fail(HG_ERR_COMM, "slab mode: the communicator of this solver was aborted after an earlier failure");
        const int rc = solve_device_impl(rtol, maxiter, st);
        // not converged / breakdown are collective outcomes (every rank sees the same all-reduced scalars)
        if (world > 1 && rc != HG_OK && rc != HG_ERR_NOT_CONVERGED && rc != HG_ERR_BREAKDOWN) {
            const std::string keep = g_err;
            abort_comm();
            g_err = keep;
        }
        return rc;
    }

    int solve_device_impl(double rtol, int maxiter, hg_stats* st) {
        if (!is_setup || !have_values) return fail(HG_ERR_INVALID, "hg_solve_device: values have not been uploaded");
        if (maxiter <= 0) maxiter = 100000;
        launches = 0;
        const bool mg = use_mg(), mg25 = use_mg25();
        if (world > 1 && !mg) return fail(HG_ERR_INVALID, "the slab mode needs the multigrid preconditioner");
        const int nb_op = op_blocks(), nb_vec = vec_blocks();
        const dim3 pg((pitch + 255) / 256, rows);
        const int nb_res = pg.x * pg.y;
        const size_t vbytes = (size_t)plane * K * sizeof(T);
        if (!d_z) {
            CU(dmalloc(&d_z, vbytes));
            CU(cudaMemsetAsync(d_z, 0, vbytes, stream));
        }
        CgState init;
        memset(&init, 0, sizeof init);
        init.tol2 = rtol * rtol;
        init.maxiter = maxiter;
        CU(cudaMemcpyAsync(d_state, &init, sizeof init, cudaMemcpyHostToDevice, stream));
        CU(cudaEventRecord(ev0, stream));
        int* done = &d_state->done;
        const Grid<T> g = grid();
        // what one inner CG run may promise: fp32 storage reduces a residual by ~1e-5 at best
        const double inner_red = kRefine ? 1e-5 : 0.0;
        const int max_outer = kRefine ? 8 : 4;
        T* e = kRefine ? d_x : reinterpret_cast<T*>(d_xm);   // the vector the inner CG updates
        // large grids: look at the state after every iteration (a sync costs microseconds, an
        // iteration milliseconds); small grids: enqueue batches, the surplus launches are no-ops
        // (the same decision on every slab: ranks must enqueue the same collectives)
        const long long npix = (long long)(world > 1 ? global_rows : rows) * cols;
        const int batch = npix >= (1LL << 22) ? 1 : ((mg || mg25) ? 4 : 32);
        const int nt = mg ? ntiles() : 0;

        CgState host;
        memset(&host, 0, sizeof host);
        double prev_ratio = 0.0;
        int outer = 0;
        for (;; ++outer) {
            // true residual in master precision -> d_r (storage type), ||r||^2 -> state
            if (mg) {
                // master-precision residual with the tile kernel (x, b in fp64 even in fp32 mode)
                HGTRY(exchange(d_xm, 1));
                if (stream_lap() && b_is_zero) {
                    launch_sl_resid(d_xm, d_r);
                    const int t0 = ghost_top / SL_R, t1 = ghost_bottom ? (rows - ghost_bottom) / SL_R : sl_nt;
                    HGTRY(stage_outer(d_spart2 + (long long)t0 * sl_ns, (t1 - t0) * sl_ns, (long long)sl_ns * sl_nt, outer == 0, inner_red * inner_red));
                    launches -= 2;
                } else {
                    launch_lap_tile<double, 1, 2>(d_xm, b_is_zero ? nullptr : d_bm, d_r, d_lpart, nullptr);
                    HGTRY(stage_outer(lap_part(d_lpart), n_own_lap(), n_lap(), outer == 0, inner_red * inner_red));
                    --launches;
                }
            } else if (kRefine) {
                resid_master_kernel<T><<<pg, 256, 0, stream>>>(g, op == HG_OP_BILAPLACIAN, d_xm, d_bm, d_r, d_part0);
                HGTRY(stage_outer(d_part0, nb_res, nb_res, outer == 0, inner_red * inner_red));
            } else {
                launch_op<1, 2>(e, d_b, d_r, d_part0, nullptr);
                HGTRY(stage_outer(d_part0, nb_op, nb_op, outer == 0, inner_red * inner_red));
            }
            launches += 2;
            mark(HG_PH_OUTER);
            CU(cudaMemcpyAsync(&host, d_state, sizeof host, cudaMemcpyDeviceToHost, stream));
            HGTRY(sync_watchful());
            if (host.done) break;                                   // converged (1), NaN (2) or out of iterations (4)
            if (outer >= max_outer) break;
            // a refinement step that did not gain a factor 4 has hit the attainable accuracy
            if (outer > 0 && !(host.outer_ratio < 0.0625 * prev_ratio)) break;
            prev_ratio = host.outer_ratio;

            // inner PCG on A e = r, e0 = 0 (fp32 mode) or continuing from x (fp64 mode)
            if (kRefine) CU(cudaMemsetAsync(d_x, 0, vbytes, stream));
            if (mg) {
                HGTRY(vcycle(d_z, d_r, done, d_tpart1));
                HGTRY(stage_beta_rz(1));
                cg_pupdate_tile_kernel<T><<<tile_grid(), FT_THREADS, 0, stream>>>(d_state, g, d_p, d_z, d_tile_active, 1);
                launches += 2;
            } else {
                if (mg25) HGTRY(vcycle25(d_z, d_r, done));
                else jacobi_apply_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(plane, K, d_z, d_r, d_dinv);
                cg_dot_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(d_state, plane, K, d_r, d_z, d_part1);
                HGTRY(stage_beta(d_part1, nb_vec, nb_vec, 1));
                CU(cudaMemcpyAsync(d_p, d_z, vbytes, cudaMemcpyDeviceToDevice, stream));
                launches += 4;
            }
            while (true) {
                for (int s = 0; s < batch; ++s) {
if (mg) {
                        mark(HG_PH_OTHER);
                        // (no exchange of p: its ghost rows are computed redundantly from z and the old p, and are
                        // valid next to the owned band just as the ghost rows of z are; HGRID_EXCHANGE_P=1 restores it)
                        if (exchange_p) HGTRY(exchange(d_p, 1));
                        HGTRY(lap_and_alpha(d_p, d_q, done));
                        launches -= 2;
                        mark(HG_PH_STENCIL);
                        cg_update_tile_kernel<T><<<tile_grid(), FT_THREADS, 0, stream>>>(d_state, g, e, d_r, d_p, d_q, d_tile_active, d_tpart0);
                        HGTRY(stage_check(tile_part(d_tpart0), n_own_tiles(), nt));
                        mark(HG_PH_UPDATE);
                        HGTRY(vcycle(d_z, d_r, done, d_tpart1));
                        HGTRY(stage_beta_rz(0));
                        cg_pupdate_tile_kernel<T><<<tile_grid(), FT_THREADS, 0, stream>>>(d_state, g, d_p, d_z, d_tile_active, 0);
                        mark(HG_PH_PUPDATE);
                        launches += 6;
                    } else {
                        launch_op<0, 1>(d_p, nullptr, d_q, d_part0, done);
                        HGTRY(stage_alpha(d_part0, nb_op, nb_op));
                        cg_update_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(d_state, plane, K, e, d_r, d_p, d_q, d_part0);
                        HGTRY(stage_check(d_part0, nb_vec, nb_vec));
                        if (mg25) HGTRY(vcycle25(d_z, d_r, done));
                        else jacobi_apply_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(plane, K, d_z, d_r, d_dinv);
                        cg_dot_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(d_state, plane, K, d_r, d_z, d_part1);
                        HGTRY(stage_beta(d_part1, nb_vec, nb_vec, 0));
                        cg_pupdate_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(d_state, plane, K, d_p, d_z);
                        launches += 7;
                    }
                }
                CU(cudaMemcpyAsync(&host, d_state, sizeof host, cudaMemcpyDeviceToHost, stream));
                HGTRY(sync_watchful());
                if (host.done) break;
            }
            if (kRefine) {
                if (mg) axpy_master_tile_kernel<T><<<tile_grid(), FT_THREADS, 0, stream>>>(g, d_xm, d_x, d_tile_active);
                else axpy_master_kernel<T><<<vec_blocks(), 256, 0, stream>>>(plane * K, d_xm, d_x);
                ++launches;
            }
            if (host.done == 2) break;
        }
        CU(cudaEventRecord(ev1, stream));
        HGTRY(sync_watchful());
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev0, ev1));
        report_phases();

        double relres = sqrt(host.outer_ratio), relres_cg = 0.0;
        for (int k = 0; k < K; ++k)
            if (host.bb[k] > 0.0) relres_cg = std::max(relres_cg, sqrt(host.rr[k] / host.bb[k]));
        int status = HG_OK;
        if (host.done == 2 || relres != relres) status = HG_ERR_BREAKDOWN;
        else if (!host.outer_done) status = HG_ERR_NOT_CONVERGED;   // maxiter, stagnation or attainable accuracy
        if (st) {
            st->iterations = host.iter;
            st->status = status;
            st->relres = relres;
            st->relres_cg = relres_cg;
            st->solve_ms = ms;
            st->launches = launches;
            const double W = sizeof(T);
            const double npx = (double)rows * cols;
            // un-preconditioned CG iteration: K*11W + 1 bytes per pixel; the
            // preconditioner's traffic is not credited
            st->bytes_moved = npx * ((double)host.iter * (K * 11.0 * W + 1.0) + (outer + 1.0) * (K * 3.0 * W + 1.0));
        }
        if (status == HG_ERR_BREAKDOWN) return fail(status, "conjugate gradients broke down (NaN or non-positive curvature): singular or under-constrained system");
        if (status == HG_ERR_NOT_CONVERGED) return fail(status, "the true residual did not reach the requested tolerance (iteration limit, stagnation, or attainable accuracy of the arithmetic)");
        return HG_OK;
    }

    int upload_u8(const uint8_t* vals, hg_stats* st) override {
        if (!is_setup) return fail(HG_ERR_INVALID, "hg_solve_u8: hg_setup has not been called");
        if (!vals) return fail(HG_ERR_INVALID, "hg_solve_u8: NULL values");
        const size_t bytes = (size_t)rows * cols * K;
        uint8_t* d_u8 = d_raw;       // (allocated by hg_set_flags: rows * cols * K bytes)
        CU(cudaEventRecord(ev0, stream));
        CU(cudaMemcpyAsync(d_u8, vals, bytes, cudaMemcpyHostToDevice, stream));
        dim3 pg((pitch + 255) / 256, rows);
        build_problem_u8_kernel<T><<<pg, 256, 0, stream>>>(grid(), d_u8, d_xm, d_bm, b_is_zero ? 0 : 1);
        b_is_zero = true;
        CU(cudaEventRecord(ev1, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev0, ev1));
        if (st) st->upload_ms = ms;
        have_values = true;
        return HG_OK;
    }

    int download_u8(uint8_t* out, hg_stats* st) override {
        if (!out) return fail(HG_ERR_INVALID, "hg_solve_u8: NULL output");
        const size_t bytes = (size_t)rows * cols * K;
        uint8_t* d_u8 = d_raw;
        CU(cudaEventRecord(ev0, stream));
        dim3 pg((pitch + 255) / 256, rows);
        export_u8_kernel<T><<<pg, 256, 0, stream>>>(grid(), d_xm, d_u8);
        CU(cudaMemcpyAsync(out, d_u8, bytes, cudaMemcpyDeviceToHost, stream));
        CU(cudaEventRecord(ev1, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev0, ev1));
        if (st) st->download_ms = ms;
        return HG_OK;
    }

    int reset_guess() override {
        if (!is_setup || !have_values) return fail(HG_ERR_INVALID, "hg_reset_guess: values have not been uploaded");
        dim3 pg((pitch + 255) / 256, rows);
        reset_guess_kernel<T><<<pg, 256, 0, stream>>>(grid(), d_xm);
        CU(cudaGetLastError());
        return HG_OK;
    }

    int download(double* x, hg_stats* st) override {
        if (!x) return fail(HG_ERR_INVALID, "hg_download: NULL buffer");
        const size_t bytes = (size_t)rows * cols * K * sizeof(double);
        if (!d_out) CU(dmalloc(&d_out, bytes));
        CU(cudaEventRecord(ev0, stream));
        dim3 pg((pitch + 255) / 256, rows);
        export_kernel<T, double><<<pg, 256, 0, stream>>>(grid(), d_xm, d_out);
        CU(cudaMemcpyAsync(x, d_out, bytes, cudaMemcpyDeviceToHost, stream));
        CU(cudaEventRecord(ev1, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev0, ev1));
        if (st) st->download_ms = ms;
        return HG_OK;
    }

    int bench_apply(int iters, double* ms_out) override {
        if (!is_setup) return fail(HG_ERR_INVALID, "hg_bench_apply: hg_setup has not been called");
        if (iters <= 0) iters = 1;
        // the launch the CG loop itself makes: q = A p with the partial sums of p.q (tile path), or the flat operator
        auto one = [&]() {
            if (use_mg() && stream_lap()) launch_sl_lap(d_p, d_q, nullptr);
            else if (use_mg()) launch_lap_tile<T, 0, 1>(d_p, nullptr, d_q, d_lpart, nullptr);
            else launch_op<0, 1>(d_p, nullptr, d_q, d_part0, nullptr);
        };
        for (int w = 0; w < 3; ++w) one();
        CU(cudaEventRecord(ev0, stream));
        for (int s = 0; s < iters; ++s) one();
        CU(cudaEventRecord(ev1, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev0, ev1));
        if (ms_out) *ms_out = ms / iters;
        return HG_OK;
    }
};

}  // namespace hg

using namespace hg;

struct hg_solver {
    SolverBase* impl;
};

extern "C" {

int hg_version(void) { return HGRID_VERSION; }

const char* hg_last_error(void) { return g_err.c_str(); }

int hg_trim_memory(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return HG_OK; }
    cudaMemPool_t pool = nullptr;
    CU(cudaDeviceGetDefaultMemPool(&pool, dev));
    CU(cudaDeviceSynchronize());
    CU(cudaMemPoolTrimTo(pool, 0));
    return HG_OK;
}

int hg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int hg_create(int rows, int cols, int channels, int op, int precision, hg_handle* out) {
    if (!out) return fail(HG_ERR_INVALID, "hg_create: NULL out");
    *out = nullptr;
    if (rows <= 0 || cols <= 0) return fail(HG_ERR_INVALID, "hg_create: rows and cols must be positive");
    if (channels <= 0 || channels > HG_MAXK) return fail(HG_ERR_INVALID, "hg_create: channels must be in 1..16");
    if (op != HG_OP_LAPLACIAN && op != HG_OP_BILAPLACIAN) return fail(HG_ERR_INVALID, "hg_create: unknown operator");
    if (precision != HG_PREC_FP32 && precision != HG_PREC_FP64) return fail(HG_ERR_INVALID, "hg_create: unknown precision");
    if (hg_device_count() <= 0) return fail(HG_ERR_NO_DEVICE, "hg_create: no CUDA device is visible; libhgrid has no CPU path");
    SolverBase* s = nullptr;
    int rc;
    if (precision == HG_PREC_FP32) {
        auto* p = new (std::nothrow) Solver<float>(rows, cols, channels, op);
        if (!p) return fail(HG_ERR_NOMEM, "out of host memory");
        rc = p->init();
        s = p;
    } else {
        auto* p = new (std::nothrow) Solver<double>(rows, cols, channels, op);
        if (!p) return fail(HG_ERR_NOMEM, "out of host memory");
        rc = p->init();
        s = p;
    }
    if (rc != HG_OK) {
        delete s;
        return rc;
    }
    hg_solver* h = new hg_solver;
    h->impl = s;
    *out = h;
    return HG_OK;
}

int hg_destroy(hg_handle h) {
    if (!h) return HG_OK;
    delete h->impl;
    delete h;
    return HG_OK;
}

#define CHECK_H(h) \
    if (!(h) || !(h)->impl) return fail(HG_ERR_INVALID, "NULL solver handle")

int hg_set_flags(hg_handle h, const uint8_t* flags) { CHECK_H(h); return h->impl->set_flags(flags); }
int hg_set_soft_weights(hg_handle h, const double* w, double scalar_w) { CHECK_H(h); return h->impl->set_soft(w, scalar_w); }
int hg_set_grad_weight(hg_handle h, double w_g) { CHECK_H(h); return h->impl->set_grad_weight(w_g); }
int hg_set_preconditioner(hg_handle h, int p) { CHECK_H(h); return h->impl->set_precond(p); }
int hg_set_edge_weights(hg_handle h, int mode) { CHECK_H(h); return h->impl->set_edges(mode); }
int hg_setup(hg_handle h, hg_setup_info* info) { CHECK_H(h); return h->impl->setup(info); }
int hg_apply(hg_handle h, const double* x, double* y) { CHECK_H(h); return h->impl->apply(x, y); }
int hg_upload_values(hg_handle h, const hg_values* v, hg_stats* st) { CHECK_H(h); return h->impl->upload(v, st); }
int hg_solve_device(hg_handle h, double rtol, int maxiter, hg_stats* st) { CHECK_H(h); return h->impl->solve_device(rtol, maxiter, st); }
int hg_download(hg_handle h, double* x, hg_stats* st) { CHECK_H(h); return h->impl->download(x, st); }
int hg_bench_apply(hg_handle h, int iters, double* ms) { CHECK_H(h); return h->impl->bench_apply(iters, ms); }
int hg_precond_apply(hg_handle h, const double* r, double* z) { CHECK_H(h); return h->impl->precond_apply(r, z); }
int hg_comm_unique_id(uint8_t* id) {
    if (!id) return fail(HG_ERR_INVALID, "hg_comm_unique_id: NULL id");
    HGTRY(nccl_load());
    ncclUniqueId uid;
    NC(g_nccl.GetUniqueId(&uid));
    memcpy(id, &uid, sizeof uid);
    return HG_OK;
}
int hg_comm_init(hg_handle h, const uint8_t* id, int rank, int world, int ghost_top, int ghost_bottom, const int* row_starts,
                 int global_rows) {
    CHECK_H(h);
    return h->impl->comm_init(id, rank, world, ghost_top, ghost_bottom, row_starts, global_rows);
}
int hg_reset_guess(hg_handle h) { CHECK_H(h); return h->impl->reset_guess(); }
int hg_mg_get_level_vectors(hg_handle h, int level, double* x, double* b) { CHECK_H(h); return h->impl->mg_get_level_vectors(level, x, b); }
int hg_set_phase_timing(hg_handle h, int enable) { CHECK_H(h); return h->impl->set_phase_timing(enable); }
int hg_get_phase_ms(hg_handle h, double* ms, int* counts, int n) { CHECK_H(h); return h->impl->get_phase_ms(ms, counts, n); }
int hg_mg_levels(hg_handle h) { CHECK_H(h); return h->impl->mg_levels(); }
int hg_mg_get_level(hg_handle h, int level, int* rows, int* cols, double* coef9) { CHECK_H(h); return h->impl->mg_get_level(level, rows, cols, coef9); }

int hg_solve_u8(hg_handle h, const uint8_t* hard_vals, uint8_t* out, double rtol, int maxiter, hg_stats* stats) {
    CHECK_H(h);
    hg_stats local;
    memset(&local, 0, sizeof local);
    hg_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof *st);
    HGTRY(h->impl->upload_u8(hard_vals, st));
    int rc = h->impl->solve_device(rtol, maxiter, st);
    if (rc != HG_OK && rc != HG_ERR_NOT_CONVERGED) return rc;
    HGTRY(h->impl->download_u8(out, st));
    return rc;
}

int hg_solve(hg_handle h, const hg_values* vals, double* x_out, double rtol, int maxiter, hg_stats* stats) {
    CHECK_H(h);
    hg_stats local;
    memset(&local, 0, sizeof local);
    hg_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof *st);
    HGTRY(h->impl->upload(vals, st));
    int rc = h->impl->solve_device(rtol, maxiter, st);
    // the iterate is returned even if the tolerance was not reached; the caller decides
    if (rc != HG_OK && rc != HG_ERR_NOT_CONVERGED) return rc;
    HGTRY(h->impl->download(x_out, st));
    return rc;
}

}  // extern "C"
