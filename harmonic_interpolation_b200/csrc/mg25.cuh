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
    // Colour-major copy of the coefficients for the smoother (round 2): colour (a, b) = (i mod 3, j mod 3) owns the compact
    // planes coefc[((3 a + b) * 25 + c) * plane3 + (i / 3) * pitch3 + j / 3].  A colour launch of the Gauss-Seidel sweep
    // reads its coefficients contiguously, every sector once per sweep; out of the pixel-major planes it touched every sector
    // that holds one of its columns, i.e. every sector in three of the nine launches (3.5x the bytes, ncu round 2).
    T* coefc;
    int pitch3;
    long long plane3;
    unsigned nz;   // bit c set: plane c holds a non-zero somewhere (level 0: the 13 entries with |di| + |dj| <= 2)
    T* res;    // K planes: the residual of the level (levels with a coarser one below)
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

// ---- colour-major copy of a level's coefficients (setup) ---------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
colour_pack25_kernel(Lev25<T> L) {
    const int j3 = blockIdx.x * blockDim.x + threadIdx.x, i3 = blockIdx.y, colour = blockIdx.z;
    if (j3 >= L.pitch3) return;
    const int i = 3 * i3 + colour / 3, j = 3 * j3 + colour % 3;
    const bool in = i < L.rows && j < L.cols;
    const long long o = (long long)i * L.pitch + j, oc = (long long)i3 * L.pitch3 + j3;
    for (int c = 0; c < 25; ++c)
        if ((L.nz >> c) & 1u) L.coefc[((long long)colour * 25 + c) * L.plane3 + oc] = in ? L.coef[c * L.plane + o] : T(0);
}

// ---- nine-colour Gauss-Seidel ---------------------------------------------------------------------------------
// update of pixel (3 i3 + colour / 3, 3 j3 + colour % 3), every channel
template <typename T>
__device__ __forceinline__ void gs25_pixel(const Lev25<T>& L, int colour, int i3, int j3) {
    const int i = 3 * i3 + colour / 3;
    const int j = 3 * j3 + colour % 3;
    if (i >= L.rows || j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    const T* __restrict__ cc = L.coefc + (long long)colour * 25 * L.plane3 + ((long long)i3 * L.pitch3 + j3);
    const T c12 = cc[12 * L.plane3];
    if (!(c12 > T(0))) return;
    T cf[25];
#pragma unroll
    for (int c = 0; c < 25; ++c) cf[c] = c == 12 ? c12 : (((L.nz >> c) & 1u) ? cc[c * L.plane3] : T(0));
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
// One launch per ROW colour a = i mod 3, one block per row: the three column colours of the row follow each other inside
// the block with a block barrier in between (rows 3 apart are not coupled by a 5 x 5 stencil, so blocks never wait for each
// other).  Same order as nine launches -- colours 3 a + 0, 1, 2, then the next a -- so the same iterates; but three launches
// per sweep instead of nine (the levels <= 512^2 are launch-latency-bound), and the row's b and the five rows of x under the
// stencil are fetched from DRAM once per row colour instead of once per colour (they were half of the smoother's traffic
// on level 0 once the coefficients were colour-major).
// one colour per launch, one pixel per thread: on the largest levels nine launches of many small blocks keep more loads in
// flight than three launches of one 512-thread block per row (ncu, level 0 of 4096^2: 0.77 ms against 0.85 ms per sweep,
// although the row kernel moves 2.4 GB instead of 3.4 GB) -- the row kernel wins where the launches themselves are the cost
template <typename T>
__global__ void __launch_bounds__(128)
gs25_kernel(Lev25<T> L, int colour, const int* __restrict__ done) {
    if (done && *done) return;
    gs25_pixel(L, colour, (int)blockIdx.y, (int)(blockIdx.x * blockDim.x + threadIdx.x));
}
#define GS25_COLOUR_LAUNCH_PIXELS (1LL << 22)     // levels of 2048^2 pixels and more: one launch per colour
#define GS25_ROW_THREADS 512
template <typename T>
__global__ void __launch_bounds__(GS25_ROW_THREADS)
gs25_rows_kernel(Lev25<T> L, int a, int forward, const int* __restrict__ done) {
    if (done && *done) return;
    const int i3 = blockIdx.x, c3 = (L.cols + 2) / 3;
    for (int q = 0; q < 3; ++q) {
        const int colour = 3 * a + (forward ? q : 2 - q);
        for (int j3 = threadIdx.x; j3 < c3; j3 += blockDim.x) gs25_pixel(L, colour, i3, j3);
        __syncthreads();          // (also orders the block's global writes for its own later reads)
    }
}

// ---- residual + restriction:  b_c = P^T (b_f - A_f x_f),  x_c = 0 ------------------------------------------
// Two passes (round 2): the residual once per fine pixel (coalesced reads of the coefficient planes that hold non-zeros),
// then full weighting of the stored residual.  The one-pass form evaluated the residual of a fine pixel for each of its
// up to four coarse parents (1.8x the coefficient bytes from DRAM at 4096^2).  Same arithmetic, same order: same bits.
template <typename T>
__global__ void __launch_bounds__(128)
resid25_kernel(Lev25<T> F, const int* __restrict__ done) {
    if (done && *done) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= F.pitch) return;
    const long long o = (long long)i * F.pitch + j;
    const bool live = j < F.cols && F.coef[12 * F.plane + o] != T(0);
    T cf[25];
#pragma unroll
    for (int c = 0; c < 25; ++c) cf[c] = (live && ((F.nz >> c) & 1u)) ? F.coef[c * F.plane + o] : T(0);
    for (int k = 0; k < F.K; ++k) {
        T res = T(0);
        if (live) {
            const T* xk = F.x + k * F.plane;
            res = F.b[k * F.plane + o];
#pragma unroll
            for (int di = -2; di <= 2; ++di)
#pragma unroll
                for (int dj = -2; dj <= 2; ++dj) {
                    const int c = (di + 2) * 5 + (dj + 2);
                    if (cf[c] != T(0)) res -= cf[c] * xk[o + (long long)di * F.pitch + dj];
                }
        }
        F.res[k * F.plane + o] = res;
    }
}
template <typename T>
__global__ void __launch_bounds__(128)
restrict25_kernel(Lev25<T> F, Lev25<T> cl, const int* __restrict__ done) {
    if (done && *done) return;
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y;
    if (J >= cl.pitch) return;
    const long long oc = (long long)I * cl.pitch + J;
    const bool live = J < cl.cols && cl.coef[12 * cl.plane + oc] != T(0);
    for (int k = 0; k < F.K; ++k) {
        double s = 0.0;
        if (live) {
            const T* rk = F.res + k * F.plane;
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
                    s += wp * (double)rk[o];
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

}  // namespace hg
