// mg25.cuh -- geometric multigrid V-cycle preconditioner for the bilaplacian systems.
//
// B = L_refl^T L_refl (recovery.py:911) is a 13-point operator; with bilinear transfer its Galerkin
// coarse operators P^T A P are 25-point (5 x 5) stencils, and 5 x 5 is closed under further
// coarsening.  So every level -- the finest included -- is stored as an EXPLICIT 25-plane stencil:
// level 0 is assembled once at setup from the flag bytes (reflected-Laplacian axis masks, cut edges,
// soft weights, gradient-penalty edges, hard pixels eliminated symmetrically), every coarser level
// is the Galerkin product of the one above.  That makes the cycle generic: nine-colour Gauss-Seidel
// (colour = 3 (i mod 3) + (j mod 3); two pixels of one colour are never coupled by a 5 x 5 stencil),
// colours reversed after the coarse correction, so the V-cycle is symmetric positive definite for
// any pattern of constraints, as PCG needs.  Transfer, far-edge extrapolation and the dense
// (generalised) inverse on the coarsest level are those of the Laplacian hierarchy (mg.cuh).
//
// Cost: the fine stencil is 25 W bytes per pixel (100 B fp32, 200 B fp64) against 1 flag byte for
// the matrix-free apply; this buys a preconditioner whose iteration count does not grow with the
// grid, where Jacobi-PCG needs O(n^2) iterations on an n x n thin-plate problem.
#pragma once
#include "common.cuh"
#include "mg.cuh"
#include "stencil.cuh"

namespace hg {

template <typename T>
struct Lev25 {
    int rows, cols, pitch;
    long long plane;
    int K;
    T* coef;   // 25 planes, plane (di + 2) * 5 + (dj + 2); a pixel whose centre coefficient is 0 is inactive
    T* x;      // K planes
    T* b;      // K planes
    T* x2;     // K planes: the other half of the double buffer of the row-coloured smoother (gs25_rows_kernel), else nullptr
    // aggregate correction for stiff couplings (finest levels only; nullptr elsewhere)
    int* label;      // cluster root (a pixel index) of every pixel that has a strong coupling, -1 otherwise
    double* aggD;    // at root pixels: sum of the cluster's block of A (v^T A v for the cluster's indicator v)
    double* aggS;    // K planes, at root pixels: sum of the residual over the cluster
};

template <typename T>
struct Acc25 {
    Lev25<T> c;
    __device__ __forceinline__ int rows() const { return c.rows; }
    __device__ __forceinline__ int cols() const { return c.cols; }
    __device__ __forceinline__ bool active(int i, int j) const {
        return i >= 0 && i < c.rows && j >= 0 && j < c.cols && c.coef[12 * c.plane + (long long)i * c.pitch + j] != T(0);
    }
    __device__ __forceinline__ double entry(int i, int j, int di, int dj) const {
        return (double)c.coef[((di + 2) * 5 + (dj + 2)) * c.plane + (long long)i * c.pitch + j];
    }
};

// ---- level 0: assemble the reduced bilaplacian system explicitly -------------------------------------
// entry (m, n) of the reflected Laplacian (recovery.py:349-500): 0 unless n is m or a 4-neighbour of m
__device__ __forceinline__ double lrefl_entry(const uint8_t* fl, int rows, int cols, int pitch, int mi, int mj, int ni, int nj) {
    if (mi < 0 || mi >= rows || mj < 0 || mj >= cols) return 0.0;
    const int di = ni - mi, dj = nj - mj;
    if (di == 0 && dj == 0) return 0.5 * ((int)axis_v(fl, rows, cols, pitch, mi, mj) + (int)axis_h(fl, rows, cols, pitch, mi, mj));
    if (dj == 0 && (di == 1 || di == -1)) return axis_v(fl, rows, cols, pitch, mi, mj) ? -0.25 : 0.0;
    if (di == 0 && (dj == 1 || dj == -1)) return axis_h(fl, rows, cols, pitch, mi, mj) ? -0.25 : 0.0;
    return 0.0;
}

template <typename T>
__global__ void __launch_bounds__(128)
assemble25_kernel(Grid<T> g, Lev25<T> L) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.pitch) return;
    const long long o = (long long)i * L.pitch + j;
    const bool free_p = j < g.cols && !(g.flags[o] & HG_FLAG_HARD);
    for (int di = -2; di <= 2; ++di)
        for (int dj = -2; dj <= 2; ++dj) {
            double a = 0.0;
            const int qi = i + di, qj = j + dj;
            if (free_p && abs(di) + abs(dj) <= 2 && qi >= 0 && qi < g.rows && qj >= 0 && qj < g.cols &&
                !(g.flags[(long long)qi * g.pitch + qj] & HG_FLAG_HARD)) {
                // (L^T L)[p, q] = sum over the rows m that reference both p and q
                const int mi[5] = {i, i - 1, i + 1, i, i}, mj[5] = {j, j, j, j - 1, j + 1};
                for (int m = 0; m < 5; ++m) {
                    const double lp = lrefl_entry(g.flags, g.rows, g.cols, g.pitch, mi[m], mj[m], i, j);
                    if (lp != 0.0) a += lp * lrefl_entry(g.flags, g.rows, g.cols, g.pitch, mi[m], mj[m], qi, qj);
                }
                const unsigned f = g.flags[o];
                if (di == 0 && dj == 0) {
                    // soft value constraints and gradient-penalty edges (recovery.py:926-927, 1280-1283)
                    a += (double)soft_w(g, o, f);
                    a += (double)pen_down(g, i, f) + (double)pen_right(g, j, f);
                    if (i > 0) a += (double)pen_down(g, i - 1, g.flags[o - g.pitch]);
                    if (j > 0) a += (double)pen_right(g, j - 1, g.flags[o - 1]);
                } else if (di == 1 && dj == 0) a -= (double)pen_down(g, i, f);
                else if (di == -1 && dj == 0) a -= (double)pen_down(g, i - 1, g.flags[o - g.pitch]);
                else if (di == 0 && dj == 1) a -= (double)pen_right(g, j, f);
                else if (di == 0 && dj == -1) a -= (double)pen_right(g, j - 1, g.flags[o - 1]);
            }
            L.coef[((di + 2) * 5 + (dj + 2)) * L.plane + o] = (T)a;
        }
}

// ---- Galerkin product  A_c = P^T A P  -----------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
galerkin25_kernel(Acc25<T> fine, Lev25<T> cl) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y;
    if (J >= cl.pitch) return;
    const long long oc = (long long)I * cl.pitch + J;
    double acc[5][5];
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) acc[a][b] = 0.0;
    if (J < cl.cols) {
        const int nfr = fine.rows(), nfc = fine.cols();
        for (int a = -1; a <= 1; ++a) {
            const double wa = w1d(a, I, nfr, cl.rows);
            if (wa == 0.0) continue;
            for (int b = -1; b <= 1; ++b) {
                const double wp = wa * w1d(b, J, nfc, cl.cols);
                if (wp == 0.0) continue;
                const int pi = 2 * I + a, pj = 2 * J + b;
                if (!fine.active(pi, pj)) continue;
                for (int di = -2; di <= 2; ++di)
                    for (int dj = -2; dj <= 2; ++dj) {
                        const int qi = pi + di, qj = pj + dj;
                        if (!fine.active(qi, qj)) continue;
                        const double apq = fine.entry(pi, pj, di, dj);
                        if (apq == 0.0) continue;
                        int Iq[2], Jq[2];
                        double wi[2], wj[2];
                        const int ni = parents(qi, cl.rows, Iq, wi);
                        const int nj = parents(qj, cl.cols, Jq, wj);
                        for (int u = 0; u < ni; ++u)
                            for (int v = 0; v < nj; ++v) acc[Iq[u] - I + 2][Jq[v] - J + 2] += wp * apq * wi[u] * wj[v];
                    }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) cl.coef[(a * 5 + b) * cl.plane + oc] = (T)acc[a][b];
}

// ---- nine-colour Gauss-Seidel, one colour per launch ------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
gs25_kernel(Lev25<T> L, int colour, const int* __restrict__ done) {
    if (done && *done) return;
    const int i = 3 * blockIdx.y + colour / 3;
    const int j = 3 * (blockIdx.x * blockDim.x + threadIdx.x) + colour % 3;
    if (i >= L.rows || j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    const T c12 = L.coef[12 * L.plane + o];
    if (!(c12 > T(0))) return;
    T cf[25];
#pragma unroll
    for (int c = 0; c < 25; ++c) cf[c] = c == 12 ? c12 : L.coef[c * L.plane + o];
    const T inv = T(1) / c12;
    for (int k = 0; k < L.K; ++k) {
        T* xk = L.x + k * L.plane;
        T v = L.b[k * L.plane + o];
#pragma unroll
        for (int di = -2; di <= 2; ++di)
#pragma unroll
            for (int dj = -2; dj <= 2; ++dj) {
                const int c = (di + 2) * 5 + (dj + 2);
                if (c != 12 && cf[c] != T(0)) v -= cf[c] * xk[o + (long long)di * L.pitch + dj];   // 0 outside the grid
            }
        xk[o] = v * inv;
    }
}

// ---- residual + restriction:  b_c = P^T (b_f - A_f x_f),  x_c = 0 ------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
resid_restrict25_kernel(Lev25<T> F, Lev25<T> cl, const int* __restrict__ done) {
    if (done && *done) return;
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y;
    if (J >= cl.pitch) return;
    const long long oc = (long long)I * cl.pitch + J;
    const bool live = J < cl.cols && cl.coef[12 * cl.plane + oc] != T(0);
    for (int k = 0; k < F.K; ++k) {
        double s = 0.0;
        if (live) {
            const T* xk = F.x + k * F.plane;
            const T* bk = F.b + k * F.plane;
            for (int a = -1; a <= 1; ++a) {
                const double wa = w1d(a, I, F.rows, cl.rows);
                if (wa == 0.0) continue;
                for (int b = -1; b <= 1; ++b) {
                    const double wp = wa * w1d(b, J, F.cols, cl.cols);
                    if (wp == 0.0) continue;
                    const int pi = 2 * I + a, pj = 2 * J + b;
                    if (pi < 0 || pi >= F.rows || pj < 0 || pj >= F.cols) continue;
                    const long long o = (long long)pi * F.pitch + pj;
                    if (F.coef[12 * F.plane + o] == T(0)) continue;
                    T res = bk[o];
#pragma unroll
                    for (int di = -2; di <= 2; ++di)
#pragma unroll
                        for (int dj = -2; dj <= 2; ++dj) {
                            const T cf = F.coef[((di + 2) * 5 + (dj + 2)) * F.plane + o];
                            if (cf != T(0)) res -= cf * xk[o + (long long)di * F.pitch + dj];
                        }
                    s += wp * (double)res;
                }
            }
        }
        cl.b[k * cl.plane + oc] = (T)s;
        cl.x[k * cl.plane + oc] = T(0);
    }
}

// ---- prolongation + correction:  x_f += P x_c -----------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
prolong_add25_kernel(Lev25<T> F, Lev25<T> cl, const int* __restrict__ done) {
    if (done && *done) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= F.cols) return;
    const long long o = (long long)i * F.pitch + j;
    if (F.coef[12 * F.plane + o] == T(0)) return;
    int Ip[2], Jp[2];
    double wi[2], wj[2];
    const int ni = parents(i, cl.rows, Ip, wi);
    const int nj = parents(j, cl.cols, Jp, wj);
    for (int k = 0; k < F.K; ++k) {
        const T* xc = cl.x + k * cl.plane;
        double s = 0.0;
        for (int u = 0; u < ni; ++u)
            for (int v = 0; v < nj; ++v) s += wi[u] * wj[v] * (double)xc[(long long)Ip[u] * cl.pitch + Jp[v]];
        F.x[k * F.plane + o] += (T)s;
    }
}

// ---- coarsest level: dense [A | I] for the Gauss-Jordan inverse of mg.cuh ---------------------------------
template <typename T>
__global__ void dense_fill25_kernel(Lev25<T> L, double* __restrict__ M) {
    const int n = L.rows * L.cols;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int i = e / L.cols, j = e % L.cols;
    const long long o = (long long)i * L.pitch + j;
    double* row = M + (long long)e * 2 * n;
    row[n + e] = 1.0;
    if (!(L.coef[12 * L.plane + o] > T(0))) { row[e] = 1.0; return; }
    for (int di = -2; di <= 2; ++di)
        for (int dj = -2; dj <= 2; ++dj) {
            const T a = L.coef[((di + 2) * 5 + (dj + 2)) * L.plane + o];
            if (a == T(0)) continue;
            const int qi = i + di, qj = j + dj;
            if (L.coef[12 * L.plane + (long long)qi * L.pitch + qj] > T(0)) row[qi * L.cols + qj] = (double)a;
        }
}
template <typename T>
__global__ void __launch_bounds__(256)
coarsest_dense25_kernel(Lev25<T> L, const double* __restrict__ M, const int* __restrict__ done) {
    if (done && *done) return;
    const int n = L.rows * L.cols;
    const int k = blockIdx.x;
    extern __shared__ double sb25[];
    for (int e = threadIdx.x; e < n; e += blockDim.x) sb25[e] = (double)L.b[k * L.plane + (long long)(e / L.cols) * L.pitch + e % L.cols];
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const double* row = M + (long long)e * 2 * n + n;
        double s = 0.0;
        for (int q = 0; q < n; ++q) s += row[q] * sb25[q];
        L.x[k * L.plane + (long long)(e / L.cols) * L.pitch + e % L.cols] = (T)s;
    }
}

// ---- aggregate correction for stiff couplings ----------------------------------------------------------------
// w_lsq = 1e5 gradient constraints tie pairs / small clusters of pixels together with couplings 10^5 times
// the stencil's.  A point smoother fixes the difference within such a cluster at once but moves the
// cluster's common value by O(1/w) per sweep, and bilinear interpolation cannot represent a two-pixel bump,
// so plain multigrid crawls (1 700 PCG iterations).  Remedy, as line relaxation is for anisotropy: after the
// Gauss-Seidel sweep, one Jacobi step in the span of the clusters' indicator vectors,
//     x += sum_C  v_C (v_C^T r) / (v_C^T A v_C),        r = b - A x,
// (before it on the way up, so the cycle stays symmetric).  Clusters are the connected components of the
// couplings |a_pq| > threshold, found at setup by label propagation; Galerkin coarsening carries the stiff
// couplings to the next levels, so the three finest levels get the correction.  175 instead of 1 157
// iterations in the numpy prototype (96^2, w = 1e5); no effect where no coupling is strong.
template <typename T>
__global__ void __launch_bounds__(256)
agg_label_init_kernel(Lev25<T> L, T thr) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.pitch) return;
    const long long o = (long long)i * L.pitch + j;
    int lab = -1;
    if (j < L.cols && L.coef[12 * L.plane + o] != T(0)) {
        for (int c = 0; c < 25; ++c)
            if (c != 12 && fabs((double)L.coef[c * L.plane + o]) > (double)thr) lab = (int)o;
    }
    L.label[o] = lab;
    L.aggD[o] = 0.0;
}
// one round of label propagation over the strong couplings (min root wins); sets *changed
template <typename T>
__global__ void __launch_bounds__(256)
agg_label_step_kernel(Lev25<T> L, T thr, int* __restrict__ changed) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    int lab = L.label[o];
    if (lab < 0) return;
    int best = lab;
    for (int di = -2; di <= 2; ++di)
        for (int dj = -2; dj <= 2; ++dj) {
            const int c = (di + 2) * 5 + (dj + 2);
            if (c == 12 || !(fabs((double)L.coef[c * L.plane + o]) > (double)thr)) continue;
            const int lq = L.label[o + (long long)di * L.pitch + dj];
            if (lq >= 0) {
                const int root = L.label[lq];          // pointer jumping: the neighbour's root's root
                best = min(best, root >= 0 ? min(lq, root) : lq);
            }
        }
    if (best < lab) { L.label[o] = best; *changed = 1; }
}
// cluster sizes (into aggD at the roots, as counts), then: clusters larger than `maxc` are dissolved -- near
// percolation they span the grid, their indicator is no longer a local mode, and thousands of atomic adds
// onto one root would serialise
template <typename T>
__global__ void __launch_bounds__(256)
agg_count_kernel(Lev25<T> L) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    const int lab = L.label[o];
    if (lab >= 0) atomicAdd(&L.aggD[lab], 1.0);
}
template <typename T>
__global__ void __launch_bounds__(256)
agg_prune_kernel(Lev25<T> L, double maxc, int* __restrict__ scratch, unsigned long long* __restrict__ kept) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    const int lab = L.label[o];
    int out = lab;
    if (lab >= 0) {
        const double n = L.aggD[lab];
        if (n < 2.0 || n > maxc) out = -1;
        else atomicAdd(kept, 1ull);
    }
    scratch[o] = out;        // labels are read by other threads until everybody has looked: write to a copy
}
template <typename T>
__global__ void __launch_bounds__(256)
agg_commit_kernel(Lev25<T> L, const int* __restrict__ scratch) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.pitch) return;
    const long long o = (long long)i * L.pitch + j;
    L.label[o] = j < L.cols ? scratch[o] : -1;
    L.aggD[o] = 0.0;
}
// aggD[root] = sum over the cluster's pixels p, q of a_pq
template <typename T>
__global__ void __launch_bounds__(256)
agg_diag_kernel(Lev25<T> L) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    const int lab = L.label[o];
    if (lab < 0) return;
    double s = 0.0;
    for (int di = -2; di <= 2; ++di)
        for (int dj = -2; dj <= 2; ++dj) {
            const T a = L.coef[((di + 2) * 5 + (dj + 2)) * L.plane + o];
            if (a != T(0) && L.label[o + (long long)di * L.pitch + dj] == lab) s += (double)a;
        }
    atomicAdd(&L.aggD[lab], s);
}
template <typename T>
__global__ void __launch_bounds__(256)
agg_zero_kernel(Lev25<T> L, const int* __restrict__ done) {
    if (done && *done) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    if (L.label[o] == (int)o)
        for (int k = 0; k < L.K; ++k) L.aggS[k * L.plane + o] = 0.0;
}
template <typename T>
__global__ void __launch_bounds__(256)
agg_resid_kernel(Lev25<T> L, const int* __restrict__ done) {
    if (done && *done) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    const int lab = L.label[o];
    if (lab < 0) return;
    T cf[25];
#pragma unroll
    for (int c = 0; c < 25; ++c) cf[c] = L.coef[c * L.plane + o];
    for (int k = 0; k < L.K; ++k) {
        const T* xk = L.x + k * L.plane;
        double res = (double)L.b[k * L.plane + o];
#pragma unroll
        for (int di = -2; di <= 2; ++di)
#pragma unroll
            for (int dj = -2; dj <= 2; ++dj) {
                const int c = (di + 2) * 5 + (dj + 2);
                if (cf[c] != T(0)) res -= (double)cf[c] * (double)xk[o + (long long)di * L.pitch + dj];
            }
        atomicAdd(&L.aggS[k * L.plane + lab], res);
    }
}
template <typename T>
__global__ void __launch_bounds__(256)
agg_apply_kernel(Lev25<T> L, const int* __restrict__ done) {
    if (done && *done) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    const int lab = L.label[o];
    if (lab < 0) return;
    const double d = L.aggD[lab];
    if (!(d > 0.0)) return;
    for (int k = 0; k < L.K; ++k) L.x[k * L.plane + o] += (T)(L.aggS[k * L.plane + lab] / d);
}

// ---- row-coloured Gauss-Seidel: one launch per ROW colour, the three column colours inside the block -------------
// (HGRID_MG25_ROWGS=1; written after round 1's GPU budget was spent: compiled, NOT YET RUN on a GPU, off by default.)
// The nine launches of gs25_kernel each touch every third pixel of every third row, so every 32-byte sector of the
// 25 coefficient planes is fetched three times per sweep.  Here a block owns a segment of one row i = ci (mod 3) and
// sweeps its three column colours in sequence, keeping the row in shared memory: the coefficients of the segment
// come from DRAM once and from L1 / L2 twice.  Same-row dependencies across segment borders are met by recomputing
// RG_H = 4 halo pixels on either side (first colour on s0-4 .. s1+4, second on s0-2 .. s1+2, third on the segment).
// To keep that recomputation race-free the iterate is double-buffered: a launch reads the rows it sweeps, and the
// rows not yet swept in this sweep, from `xin`, the rows already swept from `xout`, and writes `xout` only.  After
// the three launches `xout` holds the whole new iterate.  The ordering is exactly that of the nine-colour sweep.
#define RG_SEG 384
#define RG_THREADS 128
#define RG_H 4
template <typename T>
__global__ void __launch_bounds__(RG_THREADS)
gs25_rows_kernel(Lev25<T> L, const T* __restrict__ xin, T* __restrict__ xout, int ci, int forward, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ T xs[RG_SEG + 2 * RG_H + 4];            // columns s0 - RG_H - 2 .. s0 + RG_SEG + RG_H + 1 of the row
    const int i = 3 * blockIdx.y + ci;
    if (i >= L.rows) return;
    const int s0 = blockIdx.x * RG_SEG, s1 = min(s0 + RG_SEG, L.cols);
    const int lo = s0 - RG_H - 2, nx = RG_SEG + 2 * RG_H + 4;
    const long long orow = (long long)i * L.pitch;
    for (int k = 0; k < L.K; ++k) {
        const T* xi = xin + k * L.plane;
        T* xo = xout + k * L.plane;
        const T* bk = L.b + k * L.plane;
        __syncthreads();
        for (int e = threadIdx.x; e < nx; e += RG_THREADS) {
            const int j = lo + e;
            xs[e] = (j >= 0 && j < L.cols) ? xi[orow + j] : T(0);
        }
        __syncthreads();
        for (int step = 0; step < 3; ++step) {
            const int c = forward ? step : 2 - step;
            const int h = RG_H - 2 * step;              // this colour is needed (and valid) on s0 - h .. s1 + h - 1
            const int first = s0 - h;
            const int jstart = first + ((c - (first % 3 + 3) % 3) + 3) % 3;
            for (int j = jstart + 3 * (int)threadIdx.x; j < s1 + h; j += 3 * RG_THREADS) {
                if (j < 0 || j >= L.cols) continue;
                const long long o = orow + j;
                const T c12 = L.coef[12 * L.plane + o];
                if (!(c12 > T(0))) continue;
                T v = bk[o];
#pragma unroll
                for (int di = -2; di <= 2; ++di) {
                    // rows i + di, di != 0, have the other two row colours: swept already in this sweep?
                    const int cr = ((ci + di) % 3 + 3) % 3;
                    const bool swept = forward ? cr < ci : cr > ci;
                    const T* src = swept ? xo : xi;
#pragma unroll
                    for (int dj = -2; dj <= 2; ++dj) {
                        const int cc = (di + 2) * 5 + (dj + 2);
                        if (cc == 12) continue;
                        const T a = L.coef[cc * L.plane + o];
                        if (a == T(0)) continue;                     // also: neighbours outside the grid
                        v -= a * (di == 0 ? xs[j + dj - lo] : src[o + (long long)di * L.pitch + dj]);
                    }
                }
                xs[j - lo] = v / c12;
            }
            __syncthreads();
        }
        for (int j = s0 + (int)threadIdx.x; j < s1; j += RG_THREADS) xo[orow + j] = xs[j - lo];
    }
}

}  // namespace hg
