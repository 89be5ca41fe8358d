// capi.cu -- the extern "C" entry points of libhgrid.so (include/hgrid.h), the thread-local error string and the
// run-time binding of NCCL.
#include "solver.cuh"

namespace hg {

thread_local std::string g_err;
Nccl g_nccl;
std::map<CommKey, ncclComm_t> g_comms;
std::mutex g_comms_mu;

int nccl_load() {
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
int hg_upload_values_sparse(hg_handle h, long long n_hard, const long long* hard_idx, const double* hard_vals, long long n_soft,
                            const long long* soft_idx, const double* soft_vals, hg_stats* st) {
    CHECK_H(h);
    return h->impl->upload_sparse(n_hard, hard_idx, hard_vals, n_soft, soft_idx, soft_vals, st);
}
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

int hg_solve_composed(hg_handle h, const hg_values* vals, const hg_linear_constraints* lc, double* x_out, double rtol,
                      int maxiter, hg_stats* stats) {
    CHECK_H(h);
    hg_stats local;
    memset(&local, 0, sizeof local);
    hg_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof *st);
    HGTRY(h->impl->upload(vals, st));
    int rc = h->impl->solve_composed(lc, rtol, maxiter, st);
    if (rc != HG_OK && rc != HG_ERR_NOT_CONVERGED) return rc;
    HGTRY(h->impl->download(x_out, st));
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
