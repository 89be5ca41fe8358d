// solver_mg25.cu -- bilaplacian path: the 25-point hierarchy and its V-cycle (members of hg::Solver<T>, solver.cuh).
#include "solver.cuh"

namespace hg {

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
        if (level == 0) assemble25_kernel<T><<<cg, 128, 0, stream>>>(grid(), L);
        else galerkin25_kernel<T><<<cg, 128, 0, stream>>>(Acc25<T>{lev25.back()}, L);
        colour_pack25_kernel<T><<<dim3((L.pitch3 + 127) / 128, (L.rows + 2) / 3, 9), 128, 0, stream>>>(L);
        const bool last = (long long)r * c <= 256;
        if (!last) CU(dmalloc(&L.res, (size_t)L.plane * K * sizeof(T)));
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
        restrict25_kernel<T><<<dim3((Lc.pitch + 127) / 128, Lc.rows), 128, 0, stream>>>(lev25[l], Lc, done);
    }
    const Lev25<T>& Lb = lev25[n - 1];
    coarsest_dense25_kernel<T><<<K, 256, Lb.rows * Lb.cols * sizeof(double), stream>>>(Lb, d_dense25, done);
    for (int l = n - 2; l >= 0; --l) {
        prolong_add25_kernel<T><<<dim3((lev25[l].cols + 255) / 256, lev25[l].rows), 256, 0, stream>>>(lev25[l], lev25[l + 1], done);
        gs25(lev25[l], false, done);
    }
    launches += 3 * (n - 1) + 2;
    return HG_OK;
}

// the members defined above, for both storage types
template struct Solver<float>;
template struct Solver<double>;

}  // namespace hg
