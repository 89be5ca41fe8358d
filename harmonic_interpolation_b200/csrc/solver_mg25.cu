// solver_mg25.cu -- bilaplacian path: the 25-point hierarchy and its V-cycle (members of hg::Solver<T>, solver.cuh).
#include "solver.cuh"

namespace hg {

// ---- clusters of stiff gradient constraints (mg25_clusters.cuh) -----------------------------------
template <typename T>
int Solver<T>::build_clusters25() {
    const char* e = getenv("HGRID_MG25_CLUSTERS");          // on whenever there are gradient constraints; =0 switches it off
    if ((e && atoi(e) == 0) || w_g == 0.0) return HG_OK;
    if (plane >= (1LL << 31)) return HG_OK;                 // labels are 32-bit pixel offsets
    CU(dmalloc(&cl25.label, (size_t)plane * sizeof(int)));
    CU(dmalloc(&cl25.count, (size_t)plane * sizeof(int)));
    int* d_ch = nullptr;
    unsigned long long* d_mem = nullptr;
    CU(dmalloc(&d_ch, sizeof(int)));
    CU(dmalloc(&d_mem, sizeof(unsigned long long)));
    const dim3 pg((pitch + 255) / 256, rows);
    cl_init_kernel<<<pg, 256, 0, stream>>>(d_flags, rows, cols, pitch, cl25.label);
    int changed = 1;
    for (int r = 0; r < CL_ROUNDS && changed; r += 4) {
        CU(cudaMemsetAsync(d_ch, 0, sizeof(int), stream));
        for (int q = 0; q < 4; ++q) cl_spread_kernel<<<pg, 256, 0, stream>>>(d_flags, rows, cols, pitch, cl25.label, d_ch);
        CU(cudaMemcpyAsync(&changed, d_ch, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    CU(cudaMemsetAsync(cl25.count, 0, (size_t)plane * sizeof(int), stream));
    cl_bad_kernel<<<pg, 256, 0, stream>>>(d_flags, rows, cols, pitch, cl25.label, cl25.count);
    cl_drop_kernel<<<pg, 256, 0, stream>>>(rows, cols, pitch, cl25.label, cl25.count);
    CU(cudaMemsetAsync(cl25.count, 0, (size_t)plane * sizeof(int), stream));
    CU(cudaMemsetAsync(d_mem, 0, sizeof(unsigned long long), stream));
    cl_count_kernel<<<pg, 256, 0, stream>>>(rows, cols, pitch, cl25.label, cl25.count, d_mem);
    unsigned long long mem = 0;
    CU(cudaMemcpyAsync(&mem, d_mem, sizeof mem, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    dfree(d_ch); dfree(d_mem);
    cl25.members = (long long)mem;
    if (mem == 0) {
        dfree(cl25.label); dfree(cl25.count);
        cl25.label = cl25.count = nullptr;
        return HG_OK;
    }
    CU(dmalloc(&cl_sum, (size_t)plane * K * sizeof(T)));
    CU(dmalloc(&cl_tmp, (size_t)plane * K * sizeof(T)));
    return HG_OK;
}

// v <- C v on `nplanes` planes of a level-0 vector
template <typename T>
void Solver<T>::cluster_average(T* v, int nplanes) {
    const dim3 pg((cols + 255) / 256, rows);
    cudaMemsetAsync(cl_sum, 0, (size_t)plane * nplanes * sizeof(T), stream);
    cl_sum_kernel<T><<<pg, 256, 0, stream>>>(rows, cols, pitch, plane, nplanes, cl25.label, v, cl_sum);
    cl_mean_kernel<T><<<pg, 256, 0, stream>>>(rows, cols, pitch, plane, nplanes, cl25.label, cl25.count, cl_sum, v);
    launches += 2;
}

// level-1 operator P'^T A P' by probing (mg25_clusters.cuh), truncated to 5 x 5 with the cut entries lumped
template <typename T>
int Solver<T>::probe_level1(const Lev25<T>& L0, Lev25<T>& L1) {
    T *zero_f = nullptr, *vprobe = nullptr, *yprobe = nullptr, *raw = nullptr, *nrs = nullptr;
    CU(dmalloc(&zero_f, (size_t)L0.plane * sizeof(T)));
    CU(dmalloc(&vprobe, (size_t)L1.plane * sizeof(T)));
    CU(dmalloc(&yprobe, (size_t)L1.plane * sizeof(T)));
    CU(dmalloc(&nrs, (size_t)L1.plane * sizeof(T)));
    CU(dmalloc(&raw, (size_t)L1.plane * 25 * sizeof(T)));
    CU(cudaMemsetAsync(zero_f, 0, (size_t)L0.plane * sizeof(T), stream));
    CU(cudaMemsetAsync(raw, 0, (size_t)L1.plane * 25 * sizeof(T), stream));
    const dim3 cg((L1.pitch + 127) / 128, L1.rows), cg2((L1.cols + 127) / 128, L1.rows);
    // every next-coarser pixel counts as active while probing (restrict25_kernel looks at the centre coefficient)
    CU(cudaMemsetAsync(L1.coef, 0, (size_t)L1.plane * 25 * sizeof(T), stream));
    cl_probe_fill_kernel<T><<<cg, 128, 0, stream>>>(L1.rows, L1.cols, L1.pitch, 0, 0, 0, L1.coef + 12 * L1.plane);
    Lev25<T> F = L0;
    F.K = 1; F.x = cl_tmp; F.b = zero_f; F.res = L0.res;
    Lev25<T> Cc = L1;
    Cc.K = 1; Cc.x = vprobe; Cc.b = yprobe;
    const int s = CL_SPACING;
    for (int pr = -1; pr < s * s; ++pr) {
        const int u = pr < 0 ? 0 : pr / s, w = pr < 0 ? 0 : pr % s;
        cl_probe_fill_kernel<T><<<cg, 128, 0, stream>>>(L1.rows, L1.cols, L1.pitch, pr < 0 ? 0 : s, u, w, vprobe);
        CU(cudaMemsetAsync(cl_tmp, 0, (size_t)L0.plane * sizeof(T), stream));
        prolong_add25_kernel<T><<<dim3((L0.cols + 255) / 256, L0.rows), 256, 0, stream>>>(F, Cc, nullptr);
        cluster_average(cl_tmp, 1);
        resid25_kernel<T><<<dim3((L0.pitch + 127) / 128, L0.rows), 128, 0, stream>>>(F, nullptr);
        cluster_average(F.res, 1);
        restrict25_kernel<T><<<cg, 128, 0, stream>>>(F, Cc, nullptr);
        if (pr < 0) CU(cudaMemcpyAsync(nrs, yprobe, (size_t)L1.plane * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        else cl_probe_read_kernel<T><<<cg2, 128, 0, stream>>>(L1.rows, L1.cols, L1.pitch, L1.plane, s, u, w, yprobe, raw);
    }
    cl_probe_finish_kernel<T><<<cg, 128, 0, stream>>>(L1.rows, L1.cols, L1.pitch, L1.plane, raw, nrs, L1.coef);
    CU(cudaGetLastError());
    dfree(zero_f); dfree(vprobe); dfree(yprobe); dfree(raw); dfree(nrs);
    return HG_OK;
}

// ---- bilaplacian hierarchy --------------------------------------------------------------------
template <typename T>
int Solver<T>::build_hierarchy25() {
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
        L.coef = L.x = L.b = L.coefc = L.res = nullptr;
        L.pitch3 = round_up((L.cols + 2) / 3, 32);
        L.plane3 = (long long)((L.rows + 2) / 3) * L.pitch3;
        L.nz = 0u;
        for (int di = -2; di <= 2; ++di)
            for (int dj = -2; dj <= 2; ++dj)
                if (level > 0 || abs(di) + abs(dj) <= 2) L.nz |= 1u << ((di + 2) * 5 + (dj + 2));     // L^T L couples |di| + |dj| <= 2
        CU(dmalloc(&L.coef, (size_t)L.plane * 25 * sizeof(T)));
        CU(dmalloc(&L.coefc, (size_t)L.plane3 * 9 * 25 * sizeof(T)));
        if (level > 0) {
            CU(dmalloc(&L.x, (size_t)L.plane * K * sizeof(T)));
            CU(dmalloc(&L.b, (size_t)L.plane * K * sizeof(T)));
            CU(cudaMemsetAsync(L.x, 0, (size_t)L.plane * K * sizeof(T), stream));
            CU(cudaMemsetAsync(L.b, 0, (size_t)L.plane * K * sizeof(T), stream));
        }
        dim3 cg((L.pitch + 127) / 128, L.rows);
        const bool last = (long long)r * c <= 256;
        if (!last) CU(dmalloc(&L.res, (size_t)L.plane * K * sizeof(T)));
        if (level == 0) {
            assemble25_kernel<T><<<cg, 128, 0, stream>>>(grid(), L);
            if (!last) HGTRY(build_clusters25());
        } else if (level == 1 && cl25.members > 0) {
            HGTRY(probe_level1(lev25.back(), L));
        } else {
            galerkin25_kernel<T><<<cg, 128, 0, stream>>>(Acc25<T>{lev25.back()}, L);
        }
        colour_pack25_kernel<T><<<dim3((L.pitch3 + 127) / 128, (L.rows + 2) / 3, 9), 128, 0, stream>>>(L);
        lev25.push_back(L);
        if (last) break;
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

template <typename T>
void Solver<T>::gs25(Lev25<T>& L, bool forward, const int* done) {
    const int c3 = (L.cols + 2) / 3;
    if ((long long)L.rows * L.cols >= GS25_COLOUR_LAUNCH_PIXELS) {
        dim3 gg((c3 + 127) / 128, (L.rows + 2) / 3);
        for (int s = 0; s < mg25_sweeps; ++s)
            for (int c = 0; c < 9; ++c) gs25_kernel<T><<<gg, 128, 0, stream>>>(L, forward ? c : 8 - c, done);
        launches += 9 * mg25_sweeps;
        return;
    }
    const int threads = std::min(GS25_ROW_THREADS, round_up(c3, 32));
    for (int s = 0; s < mg25_sweeps; ++s)
        for (int q = 0; q < 3; ++q) gs25_rows_kernel<T><<<(L.rows + 2) / 3, threads, 0, stream>>>(L, forward ? q : 2 - q, forward ? 1 : 0, done);
    launches += 3 * mg25_sweeps;
}

// z = M^-1 r: one symmetric V-cycle from the zero guess on the 25-point hierarchy (level 0 works in place on z, r)
template <typename T>
int Solver<T>::vcycle25(T* z, T* r, const int* done) {
    lev25[0].x = z; lev25[0].b = r;
    CU(cudaMemsetAsync(z, 0, (size_t)plane * K * sizeof(T), stream));
    const int n = (int)lev25.size();
    for (int l = 0; l + 1 < n; ++l) {
        gs25(lev25[l], true, done);
        const Lev25<T>& Lc = lev25[l + 1];
        resid25_kernel<T><<<dim3((lev25[l].pitch + 127) / 128, lev25[l].rows), 128, 0, stream>>>(lev25[l], done);
        if (l == 0 && cl25.members > 0) cluster_average(lev25[0].res, K);       // R' = P^T C
        restrict25_kernel<T><<<dim3((Lc.pitch + 127) / 128, Lc.rows), 128, 0, stream>>>(lev25[l], Lc, done);
    }
    const Lev25<T>& Lb = lev25[n - 1];
    coarsest_dense25_kernel<T><<<K, 256, Lb.rows * Lb.cols * sizeof(double), stream>>>(Lb, d_dense25, done);
    for (int l = n - 2; l >= 0; --l) {
        if (l == 0 && cl25.members > 0) {
            // x += C P x_1  (P' = C P): the prolongated correction is averaged over the clusters before it is added
            Lev25<T> F = lev25[0];
            F.x = cl_tmp;
            CU(cudaMemsetAsync(cl_tmp, 0, (size_t)plane * K * sizeof(T), stream));
            prolong_add25_kernel<T><<<dim3((F.cols + 255) / 256, F.rows), 256, 0, stream>>>(F, lev25[1], done);
            cluster_average(cl_tmp, K);
            cl_add_kernel<T><<<vec_blocks(), 256, 0, stream>>>(plane * K, lev25[0].x, cl_tmp);
            launches += 2;
        } else {
            prolong_add25_kernel<T><<<dim3((lev25[l].cols + 255) / 256, lev25[l].rows), 256, 0, stream>>>(lev25[l], lev25[l + 1], done);
        }
        gs25(lev25[l], false, done);
    }
    launches += 3 * (n - 1) + 2;
    return HG_OK;
}

// the members defined above, for both storage types
template struct Solver<float>;
template struct Solver<double>;

}  // namespace hg
