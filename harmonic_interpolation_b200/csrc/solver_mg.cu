// solver_mg.cu -- Laplacian path: set-up of the level-0 code bytes / activity / task lists and of the Galerkin hierarchy, the V-cycle (members of hg::Solver<T>, solver.cuh).
#include "solver.cuh"

namespace hg {

// ---- multigrid ----------------------------------------------------------------------------
template <typename T>
void Solver<T>::free_hierarchy() {
    for (auto* lv : {&coarse, &gcoarse}) {
        for (auto& c : *lv) { dfree(c.coef); dfree(c.x); dfree(c.x2); dfree(c.b); dfree(c.code); dfree(c.act_up); dfree(c.act_down); dfree(c.cs_act); dfree(c.refq); }
        lv->clear();
    }
    dfree(d_dense); dfree(d_gdense);
    d_dense = d_gdense = nullptr;
    for (size_t l = 0; l < lev25.size(); ++l) {
        dfree(lev25[l].coef); dfree(lev25[l].coefc); dfree(lev25[l].res);
        if (l > 0) { dfree(lev25[l].x); dfree(lev25[l].b); }
    }
    lev25.clear();
    dfree(cl25.label); dfree(cl25.count); dfree(cl_sum); dfree(cl_tmp);
    cl25.label = cl25.count = nullptr; cl25.members = 0; cl_sum = cl_tmp = nullptr;
    dfree(d_dense25);
    d_dense25 = nullptr;
}

template <typename T>
int Solver<T>::alloc_level(Coarse<T>& L, int r, int c) {
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

template <typename T>
int Solver<T>::build_ref_stencils() {
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
template <typename T>
int Solver<T>::finish_level(Coarse<T>& L, int level) {
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

// fused legs of a 9-point level: pre-smoothing + residual + restriction / prolongation + post-smoothing
template <typename T>
void Solver<T>::coarse_down(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
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
template <typename T>
void Solver<T>::coarse_up(Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
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

template <typename T>
int Solver<T>::build_dense(const Coarse<T>& L, double** out) {
    const int n = L.rows * L.cols;
    CU(dmalloc(out, (size_t)n * 2 * n * sizeof(double)));
    dense_from_stencil_kernel<T><<<64, 256, 0, stream>>>(L, *out);
    dense_fill_kernel<T><<<(n + 127) / 128, 128, 0, stream>>>(L, *out);
    gauss_jordan_kernel<<<1, 1024, 2 * n * sizeof(double), stream>>>(*out, n);
    return HG_OK;
}

template <typename T>
int Solver<T>::build_tiles() {
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
    dfree(d_down_active); dfree(d_code); dfree(d_qmask); dfree(d_lpart);
    d_down_active = d_code = nullptr; d_qmask = nullptr; d_lpart = nullptr;
    CU(dmalloc(&d_down_active, ntiles()));
    CU(dmalloc(&d_code, plane));
    CU(dmalloc(&d_lpart, (size_t)n_lap() * K * sizeof(double)));
    CU(cudaMemsetAsync(d_lpart, 0, (size_t)n_lap() * K * sizeof(double), stream));
    tile_activity_kernel<<<tile_grid(), 256, 0, stream>>>(d_flags, rows, cols, pitch, 0, d_tile_active);
    tile_activity_kernel<<<tile_grid(), 256, 0, stream>>>(d_flags, rows, cols, pitch, 1, d_down_active);
    opcode_kernel<T><<<dim3((pitch + 255) / 256, rows), 256, 0, stream>>>(grid(), d_code);
    {
        const int qp = pitch >> 2, qg = (rows + 7) >> 3;
        CU(dmalloc(&d_qmask, (size_t)qp * qg * sizeof(unsigned)));
        sl_qmask_kernel<<<dim3((qp + 127) / 128, qg), 128, 0, stream>>>(d_code, rows, pitch, d_qmask);
    }
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

template <typename T>
int Solver<T>::build_hierarchy() {
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

template <typename T>
void Solver<T>::coarse_pre(const Coarse<T>& L, const int* done) {
    dim3 mg((L.cols / 2 + 1 + 255) / 256, (L.rows + 1) / 2);
    for (int c = 0; c < 4; ++c) mcgs_kernel<T><<<mg, 256, 0, stream>>>(L, c, done);
    launches += 4;
}

template <typename T>
void Solver<T>::coarse_post(const Coarse<T>& L, const int* done) {
    dim3 mg((L.cols / 2 + 1 + 255) / 256, (L.rows + 1) / 2);
    for (int c = 3; c >= 0; --c) mcgs_kernel<T><<<mg, 256, 0, stream>>>(L, c, done);
    launches += 4;
}

template <typename T>
void Solver<T>::coarse_restrict(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
    dim3 gg((Lc.pitch + 255) / 256, Lc.rows);
    CoarseAcc<T> ca{L};
    resid_restrict_kernel<T, CoarseAcc<T>><<<gg, 256, 0, stream>>>(ca, L.x, L.b, Lc, K, done);
    ++launches;
}

template <typename T>
void Solver<T>::coarse_prolong(const Coarse<T>& L, const Coarse<T>& Lc, const int* done) {
    CoarseAcc<T> ca{L};
    dim3 gg((L.cols + 255) / 256, L.rows);
    prolong_add_kernel<T, CoarseAcc<T>><<<gg, 256, 0, stream>>>(ca, L.x, Lc, K, done);
    ++launches;
}

template <typename T>
void Solver<T>::coarse_solve(const Coarse<T>& L, const double* dense, const int* done) {
    const int n = L.rows * L.cols;
    coarsest_dense_kernel<T><<<K, 256, n * sizeof(double), stream>>>(L, dense, done);
    ++launches;
}

// V-cycle over levels[from..] of a hierarchy whose first level already holds b (and x = 0)
template <typename T>
void Solver<T>::coarse_vcycle(std::vector<Coarse<T>>& lv, const double* dense, const int* done) {
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

template <typename T>
int Solver<T>::vcycle_generic(T* z, T* r, const int* done) {
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

template <typename T>
int Solver<T>::vcycle(T* z, T* r, const int* done, double* part_rz) {
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

   // lean kernels: one ring of rows per warp
// part 0: every task; 1: the interior tasks (front of the lists); 2: the boundary tasks (HGRID_OVERLAP)
template <typename T>
void Solver<T>::launch_sl_down(const Grid<T>& g, const T* r, const Coarse<T>& c1, const int* done, int part) {
    const int zx = unfused_coarse ? 1 : 0;
    const int o_l = part == 2 ? sl_n_int[0] : 0, n_l = part == 1 ? sl_n_int[0] : sl_n_down_lean - o_l;
    const int o_g = part == 2 ? sl_n_int[1] : 0, n_g = part == 1 ? sl_n_int[1] : sl_n_down_gen - o_g;
    if (n_l > 0)
        sl_down_kernel<T, false, true><<<sl_grid(n_l), sl_block(), sl_smem(), stream>>>(g, d_code, d_qmask, r, c1, d_sl_tasks + sl_off[0] + o_l, n_l, sl_ns, zx, done);
    if (n_g > 0) {
        if (has_soft()) sl_down_kernel<T, true, false><<<sl_grid(n_g), sl_block(), 0, stream>>>(g, d_code, d_qmask, r, c1, d_sl_tasks + sl_off[1] + o_g, n_g, sl_ns, zx, done);
        else sl_down_kernel<T, false, false><<<sl_grid(n_g), sl_block(), 0, stream>>>(g, d_code, d_qmask, r, c1, d_sl_tasks + sl_off[1] + o_g, n_g, sl_ns, zx, done);
    }
    launches += (n_l > 0) + (n_g > 0) - (part == 2 ? 0 : 1);
}

template <typename T>
void Solver<T>::launch_sl_up(const Grid<T>& g, const T* r, T* z, const Coarse<T>& c1, int have_coarse, const int* done) {
    if (sl_n_up_lean)
        sl_up_kernel<T, false, true><<<sl_grid(sl_n_up_lean), sl_block(), sl_smem(), stream>>>(g, d_code, d_qmask, r, z, c1, have_coarse, d_sl_tasks + sl_off[2], sl_n_up_lean, sl_ns, sl_nt, d_spart, done);
    if (sl_n_up_gen) {
        if (has_soft()) sl_up_kernel<T, true, false><<<sl_grid(sl_n_up_gen), sl_block(), 0, stream>>>(g, d_code, d_qmask, r, z, c1, have_coarse, d_sl_tasks + sl_off[3], sl_n_up_gen, sl_ns, sl_nt, d_spart, done);
        else sl_up_kernel<T, false, false><<<sl_grid(sl_n_up_gen), sl_block(), 0, stream>>>(g, d_code, d_qmask, r, z, c1, have_coarse, d_sl_tasks + sl_off[3], sl_n_up_gen, sl_ns, sl_nt, d_spart, done);
    }
    launches += (sl_n_up_lean > 0) + (sl_n_up_gen > 0) - 1;
}

template <typename T>
void Solver<T>::launch_sl_lap(const T* p, T* q, const int* done) {
    const Grid<T> g = grid();
    if (sl_n_up_lean)
        sl_lap_kernel<T, T, false, 0><<<sl_grid(sl_n_up_lean), sl_block(), sl_smem(), stream>>>(g, d_code, d_qmask, p, q, d_sl_tasks + sl_off[2], sl_n_up_lean, sl_ns, sl_nt, d_spart2, done);
    if (sl_n_up_gen)
        sl_lap_kernel<T, T, true, 0><<<sl_grid(sl_n_up_gen), sl_block(), sl_smem(), stream>>>(g, d_code, d_qmask, p, q, d_sl_tasks + sl_off[3], sl_n_up_gen, sl_ns, sl_nt, d_spart2, done);
    launches += (sl_n_up_lean > 0) + (sl_n_up_gen > 0);
}

// r = -A x from the fp64 master iterate (the true residual when the right-hand side is zero), partial sums of r.r
template <typename T>
void Solver<T>::launch_sl_resid(const double* x, T* r) {
    const Grid<T> g = grid();
    const size_t sm = (size_t)SL_WPB * SL_RING * 128 * sizeof(double);
    if (sl_n_up_lean)
        sl_lap_kernel<double, T, false, 1><<<sl_grid(sl_n_up_lean), sl_block(), sm, stream>>>(g, d_code, d_qmask, x, r, d_sl_tasks + sl_off[2], sl_n_up_lean, sl_ns, sl_nt, d_spart2, nullptr);
    if (sl_n_up_gen)
        sl_lap_kernel<double, T, true, 1><<<sl_grid(sl_n_up_gen), sl_block(), sm, stream>>>(g, d_code, d_qmask, x, r, d_sl_tasks + sl_off[3], sl_n_up_gen, sl_ns, sl_nt, d_spart2, nullptr);
    launches += (sl_n_up_lean > 0) + (sl_n_up_gen > 0);
}

// the stencil launch of the CG loop + the scalar stage that forms alpha
template <typename T>
int Solver<T>::lap_and_alpha(const T* p, T* q, const int* done) {
    if (stream_lap()) {
        launch_sl_lap(p, q, done);
        const int t0 = ghost_top / SL_R, t1 = ghost_bottom ? (rows - ghost_bottom) / SL_R : sl_nt;
        return stage_alpha(d_spart2 + (long long)t0 * sl_ns, (t1 - t0) * sl_ns, (long long)sl_ns * sl_nt);
    }
    launch_lap_tile<T, 0, 1>(p, nullptr, q, d_lpart, done);
    return stage_alpha(lap_part(d_lpart), n_own_lap(), n_lap());
}

// scalar stage after the V-cycle: r.z from the partial sums the upward leg left (owned rows only)
template <typename T>
int Solver<T>::stage_beta_rz(int init) {
    if (generic_nu > 0 && world == 1) return stage_beta(d_part1, vec_blocks(), vec_blocks(), init);
    if (stream_legs()) {
        const int t0 = ghost_top / SL_R, t1 = ghost_bottom ? (rows - ghost_bottom) / SL_R : sl_nt;
        return stage_beta(d_spart + (long long)t0 * sl_ns, (t1 - t0) * sl_ns, (long long)sl_ns * sl_nt, init);
    }
    return stage_beta(tile_part(d_tpart1), n_own_tiles(), ntiles(), init);
}

// the members defined above, for both storage types
template struct Solver<float>;
template struct Solver<double>;

}  // namespace hg
