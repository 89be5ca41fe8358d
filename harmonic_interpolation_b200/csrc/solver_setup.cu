// solver_setup.cu -- handle life cycle, pattern upload, hg_setup, value upload / download, operator application, introspection (members of hg::Solver<T>, solver.cuh).
#include "solver.cuh"

namespace hg {

template <typename T>
void Solver<T>::mark(int id) {
    if (!phase_timing) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, stream);
    phase_ev.push_back(e);
    phase_id.push_back(id);
}

template <typename T>
void Solver<T>::report_phases() {
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

template <typename T>
int Solver<T>::finish_task_lists() {      // after a stream sync: the list lengths the launches need
    if (!task_counts_pending) return HG_OK;
    int c[8];
    CU(cudaMemcpy(c, d_task_counts, sizeof c, cudaMemcpyDeviceToHost));
    sl_n_down_lean = c[0]; sl_n_down_gen = c[1]; sl_n_up_lean = c[2]; sl_n_up_gen = c[3];
    sl_n_int[0] = c[4]; sl_n_int[1] = c[5];
    task_counts_pending = false;
    return HG_OK;
}

template <typename T>
Solver<T>::~Solver() {
    dfree(d_flags); dfree(d_soft); dfree(d_dinv); dfree(d_raw);
    dfree(d_x); dfree(d_b); dfree(d_r); dfree(d_p); dfree(d_q);
    for (auto& p : d_stage) dfree(p);
    dfree(d_out); dfree(d_part0); dfree(d_part1); dfree(d_state); dfree(d_presum); dfree(d_counts);
    free_hierarchy();
    dfree(d_z);
    if (kRefine) { dfree(d_xm); dfree(d_bm); }
    dfree(d_tile_active); dfree(d_tpart0); dfree(d_tpart1);
    dfree(d_down_active); dfree(d_code); dfree(d_qmask); dfree(d_lpart);
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

template <typename T>
int Solver<T>::init() {
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

 // rows * cols bytes as the caller holds them (flags; also the uint8 image of hg_solve_u8)
template <typename T>
int Solver<T>::set_flags(const uint8_t* f) {
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

template <typename T>
int Solver<T>::set_soft(const double* w, double scalar) {
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

template <typename T>
int Solver<T>::set_edges(int mode) {
    if (mode != HG_EDGES_REFERENCE && mode != HG_EDGES_UNIFORM) return fail(HG_ERR_INVALID, "unknown edge-weight mode");
    if (mode == HG_EDGES_UNIFORM && op != HG_OP_LAPLACIAN) return fail(HG_ERR_INVALID, "uniform edge weights apply to the Laplacian operator only");
    edge_uniform = mode;
    is_setup = false;
    return HG_OK;
}

template <typename T>
int Solver<T>::set_precond(int p) {
    if (p != HG_PRECOND_JACOBI && p != HG_PRECOND_MULTIGRID) return fail(HG_ERR_INVALID, "unknown preconditioner");
    precond = p;
    is_setup = false;
    return HG_OK;
}

template <typename T>
int Solver<T>::ensure_vectors() {
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

template <typename T>
int Solver<T>::mg_get_level(int level, int* r_out, int* c_out, double* coef9) {
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
template <typename T>
int Solver<T>::mg_get_level_vectors(int level, double* xo, double* bo) {
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
template <typename T>
int Solver<T>::precond_apply(const double* r, double* z) {
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

template <typename T>
int Solver<T>::find_unconstrained(unsigned long long* out) {
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

template <typename T>
int Solver<T>::setup(hg_setup_info* info) {
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

template <typename T>
int Solver<T>::stage(int slot, const double* host) {
    const size_t bytes = (size_t)rows * cols * K * sizeof(double);
    if (!host) return HG_OK;
    if (!d_stage[slot]) CU(dmalloc(&d_stage[slot], bytes));
    CU(cudaMemcpyAsync(d_stage[slot], host, bytes, cudaMemcpyHostToDevice, stream));
    return HG_OK;
}

// y = F(x) with the kernel the solver itself uses for this configuration
template <typename T>
void Solver<T>::apply_op(const T* x, T* y) {
    if (use_mg()) {
        launch_lap_tile<T, 0, 0>(x, nullptr, y, nullptr, nullptr);
    } else {
        launch_op<0, 0>(x, nullptr, y, nullptr, nullptr);
    }
}

template <typename T>
int Solver<T>::apply(const double* x, double* y) {
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

template <typename T>
int Solver<T>::upload(const hg_values* v, hg_stats* st) {
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

// values of the hard / soft constraints as compact lists: scattered into the staging planes on the device
static __global__ void scatter_values_kernel(const long long* __restrict__ idx, const double* __restrict__ vals, long long n, int K,
                                             double* __restrict__ dense) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * K) return;
    const long long e = t / K;
    dense[idx[e] * K + (t - e * K)] = vals[t];
}

template <typename T>
int Solver<T>::upload_sparse(long long nh, const long long* hidx, const double* hv, long long ns, const long long* sidx, const double* sv,
                             hg_stats* st) {
    if (!is_setup) return fail(HG_ERR_INVALID, "hg_upload_values_sparse: hg_setup has not been called");
    if (nh < 0 || ns < 0 || (nh > 0 && (!hidx || !hv)) || (ns > 0 && (!sidx || !sv))) return fail(HG_ERR_INVALID, "hg_upload_values_sparse: bad arguments");
    const long long npix = (long long)rows * cols;
    for (long long e = 0; e < nh; ++e) if (hidx[e] < 0 || hidx[e] >= npix) return fail(HG_ERR_INVALID, "hg_upload_values_sparse: pixel index outside the grid");
    for (long long e = 0; e < ns; ++e) if (sidx[e] < 0 || sidx[e] >= npix) return fail(HG_ERR_INVALID, "hg_upload_values_sparse: pixel index outside the grid");
    CU(cudaEventRecord(ev0, stream));
    const size_t bytes = (size_t)npix * K * sizeof(double);
    long long* d_idx = nullptr; double* d_val = nullptr;
    const long long nmax = std::max(nh, ns);
    if (nmax > 0) {
        CU(dmalloc(&d_idx, (size_t)nmax * sizeof(long long)));
        CU(dmalloc(&d_val, (size_t)nmax * K * sizeof(double)));
    }
    const long long n_of[2] = {nh, ns};
    const long long* idx_of[2] = {hidx, sidx};
    const double* val_of[2] = {hv, sv};
    for (int slot = 0; slot < 2; ++slot) {
        const long long n = n_of[slot];
        if (n == 0) continue;
        if (!d_stage[slot]) CU(dmalloc(&d_stage[slot], bytes));
        CU(cudaMemsetAsync(d_stage[slot], 0, bytes, stream));
        CU(cudaMemcpyAsync(d_idx, idx_of[slot], (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(d_val, val_of[slot], (size_t)n * K * sizeof(double), cudaMemcpyHostToDevice, stream));
        scatter_values_kernel<<<(unsigned)((n * K + 255) / 256), 256, 0, stream>>>(d_idx, d_val, n, K, d_stage[slot]);
    }
    b_is_zero = ns == 0;
    dim3 pg((pitch + 255) / 256, rows);
    build_problem_kernel<T, double><<<pg, 256, 0, stream>>>(grid(), nh ? d_stage[0] : nullptr, ns ? d_stage[1] : nullptr, nullptr, nullptr,
                                                           nullptr, nullptr, d_xm, d_bm);
    CU(cudaEventRecord(ev1, stream));
    CU(cudaStreamSynchronize(stream));
    CU(cudaGetLastError());
    dfree(d_idx); dfree(d_val);
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ev0, ev1));
    if (st) st->upload_ms = ms;
    have_values = true;
    return HG_OK;
}

template <typename T>
int Solver<T>::upload_u8(const uint8_t* vals, hg_stats* st) {
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

template <typename T>
int Solver<T>::download_u8(uint8_t* out, hg_stats* st) {
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

template <typename T>
int Solver<T>::reset_guess() {
    if (!is_setup || !have_values) return fail(HG_ERR_INVALID, "hg_reset_guess: values have not been uploaded");
    dim3 pg((pitch + 255) / 256, rows);
    reset_guess_kernel<T><<<pg, 256, 0, stream>>>(grid(), d_xm);
    CU(cudaGetLastError());
    return HG_OK;
}

template <typename T>
int Solver<T>::download(double* x, hg_stats* st) {
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

template <typename T>
int Solver<T>::bench_apply(int iters, double* ms_out) {
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

// the members defined above, for both storage types
template struct Solver<float>;
template struct Solver<double>;

}  // namespace hg
