// mg.cuh -- geometric multigrid V-cycle preconditioner for the masked Laplacian system.
//
// Hierarchy: vertex-centred coarsening, coarse (I,J) <-> fine (2I,2J), coarse size ceil(n/2).
// Level 0 is the flag-encoded 5-point operator (stencil.cuh); every coarser level stores an
// explicit 9-point stencil A_c = P^T A P (Galerkin), which makes the coarse correction an
// A-orthogonal projection for ANY pattern of hard pixels, cut edges, soft weights and
// gradient-penalty edges -- the V-cycle is symmetric positive definite by construction, as
// PCG needs.
//
// Constraint-aware transfer: P is bilinear (1, 1/2, 1/4) times the mask of active fine pixels
// (hard pixels carry zero error and receive no correction), R = P^T.  For even fine sizes the
// last fine row / column lies beyond the last coarse one; P extrapolates it by a constant
// (weight 1 instead of 1/2), which keeps Neumann boundaries exact for constants.
//
// Smoothers: red-black Gauss-Seidel on level 0 (5-point), four-colour Gauss-Seidel on the
// 9-point levels; post-smoothing runs the colours in reverse so the cycle is symmetric.
#pragma once
#include "common.cuh"

namespace hg {

// explicit 9-point level; coefficient c = (di+1)*3 + (dj+1), centre 4.  Storage is interleaved by QUAD of four
// consecutive pixels of a row: [quad][9][4] -- the nine coefficient vectors of a quad are 144 contiguous bytes (fp32),
// which is what a lane of the streaming legs (coarse_stream.cuh) needs for a row, and a general quad costs 5 sectors
// of HBM traffic instead of 9.  cidx(c, o) is the element index of coefficient c of pixel offset o = i * pitch + j.
// A pixel whose centre coefficient is 0 is inactive (x stays 0 there).
__host__ __device__ __forceinline__ long long cidx(int c, long long o) { return (o >> 2) * 36 + c * 4 + (o & 3); }
template <typename T>
struct Coarse {
    int rows, cols, pitch;
    long long plane;
    int K;
    T* coef;   // 9 planes
    T* x;      // K planes
    T* b;      // K planes
    T* x2;     // K planes: the fused upward legs write here (x + P e and the post-smoothing read x of NEIGHBOURING tiles /
               // chunks, which must still hold the pre-smoothed values); the host swaps x and x2 after the launch
    // fused legs (coarse_fused.cuh): code byte per pixel (0 inactive, 1 regular = coefficients equal
    // `ref`, 2 general), activity of the 64 x 32 tiles (halo 0 / halo 1), the level's interior stencil
    uint8_t* code;
    uint8_t* act_up;
    uint8_t* act_down;
    int tiles_x, tiles_y;
    int has_ref;
    T ref[9];
    // streaming legs (coarse_stream.cuh): chunks of 120 columns x cs_cr rows, activity bits per chunk; cs_cr == 0: tile legs
    uint8_t* cs_act;
    int cs_cr, cs_ns, cs_nt;
    T* refq;   // ref[n] four times over (n * 4 + p): what a lane without a general pixel reads in place of the planes
};

// ---- transfer weights ---------------------------------------------------------------------
// weight of fine index 2I + a (a in {-1,0,1}) in the basis function of coarse index I
__device__ __forceinline__ double w1d(int a, int I, int nf, int nc) {
    if (a == 0) return 1.0;
    if (a < 0) return (2 * I - 1 >= 0) ? 0.5 : 0.0;
    if (2 * I + 1 >= nf) return 0.0;
    return (I + 1 < nc) ? 0.5 : 1.0;
}
// parents of fine index i: returns count (1 or 2), fills coarse indices and weights
__device__ __forceinline__ int parents(int i, int nc, int* I, double* w) {
    if ((i & 1) == 0) { I[0] = i >> 1; w[0] = 1.0; return 1; }
    const int I0 = (i - 1) >> 1;
    if (I0 + 1 < nc) { I[0] = I0; w[0] = 0.5; I[1] = I0 + 1; w[1] = 0.5; return 2; }
    I[0] = I0; w[0] = 1.0; return 1;
}

// ---- operator accessors ----------------------------------------------------------------------
template <typename T>
struct FineAcc {            // level 0: flag-encoded 5-point operator
    Grid<T> g;
    __device__ __forceinline__ int rows() const { return g.rows; }
    __device__ __forceinline__ int cols() const { return g.cols; }
    __device__ __forceinline__ int pitch() const { return g.pitch; }
    __device__ __forceinline__ long long plane() const { return g.plane; }
    __device__ __forceinline__ bool active(int i, int j) const {
        return i >= 0 && i < g.rows && j >= 0 && j < g.cols && !(g.flags[(long long)i * g.pitch + j] & HG_FLAG_HARD);
    }
    // the four edge weights and the diagonal of pixel (i,j)
    __device__ __forceinline__ void weights(int i, int j, T& wu, T& wd, T& wl, T& wr, T& diag) const {
        const long long o = (long long)i * g.pitch + j;
        const unsigned f = g.flags[o];
        wu = i > 0 ? w_down(g, i - 1, j, g.flags[o - g.pitch]) : T(0);
        wd = w_down(g, i, j, f);
        wl = j > 0 ? w_right(g, i, j - 1, g.flags[o - 1]) : T(0);
        wr = w_right(g, i, j, f);
        diag = wu + wd + wl + wr + soft_w(g, o, f);
    }
    // A[p, p+d]
    __device__ __forceinline__ double entry(int i, int j, int di, int dj) const {
        if (di != 0 && dj != 0) return 0.0;
        T wu, wd, wl, wr, diag;
        weights(i, j, wu, wd, wl, wr, diag);
        if (di == 0 && dj == 0) return (double)diag;
        if (di == -1) return -(double)wu;
        if (di == 1) return -(double)wd;
        if (dj == -1) return -(double)wl;
        return -(double)wr;
    }
    // (A x)_p for an active pixel; x is 0 at inactive pixels
    __device__ __forceinline__ T apply(const T* x, int i, int j) const {
        T wu, wd, wl, wr, diag;
        weights(i, j, wu, wd, wl, wr, diag);
        const long long o = (long long)i * g.pitch + j;
        T v = diag * x[o];
        if (wu != T(0)) v -= wu * x[o - g.pitch];
        if (wd != T(0)) v -= wd * x[o + g.pitch];
        if (wl != T(0)) v -= wl * x[o - 1];
        if (wr != T(0)) v -= wr * x[o + 1];
        return v;
    }
};

template <typename T>
struct CoarseAcc {          // explicit 9-point operator
    Coarse<T> c;
    __device__ __forceinline__ int rows() const { return c.rows; }
    __device__ __forceinline__ int cols() const { return c.cols; }
    __device__ __forceinline__ int pitch() const { return c.pitch; }
    __device__ __forceinline__ long long plane() const { return c.plane; }
    __device__ __forceinline__ bool active(int i, int j) const {
        return i >= 0 && i < c.rows && j >= 0 && j < c.cols && c.coef[cidx(4, (long long)i * c.pitch + j)] != T(0);
    }
    __device__ __forceinline__ double entry(int i, int j, int di, int dj) const {
        return (double)c.coef[cidx((di + 1) * 3 + (dj + 1), (long long)i * c.pitch + j)];
    }
    __device__ __forceinline__ T apply(const T* x, int i, int j) const {
        const long long o = (long long)i * c.pitch + j;
        T v = T(0);
#pragma unroll
        for (int di = -1; di <= 1; ++di)
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                const T a = c.coef[cidx((di + 1) * 3 + (dj + 1), o)];
                if (a != T(0)) v += a * x[o + (long long)di * c.pitch + dj];   // a == 0 outside the grid
            }
        return v;
    }
};

// ---- Galerkin product  A_c = P^T A P  (setup) --------------------------------------------------
template <typename T, typename Acc>
__global__ void galerkin_kernel(Acc fine, Coarse<T> cl) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y;
    if (J >= cl.pitch) return;
    const long long oc = (long long)I * cl.pitch + J;
    double acc[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[a][b] = 0.0;
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
                for (int di = -1; di <= 1; ++di)
                    for (int dj = -1; dj <= 1; ++dj) {
                        const int qi = pi + di, qj = pj + dj;
                        if (!fine.active(qi, qj)) continue;
                        const double apq = fine.entry(pi, pj, di, dj);
                        if (apq == 0.0) continue;
                        int Iq[2], Jq[2];
                        double wi[2], wj[2];
                        const int ni = parents(qi, cl.rows, Iq, wi);
                        const int nj = parents(qj, cl.cols, Jq, wj);
                        for (int u = 0; u < ni; ++u)
                            for (int v = 0; v < nj; ++v) acc[Iq[u] - I + 1][Jq[v] - J + 1] += wp * apq * wi[u] * wj[v];
                    }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) cl.coef[cidx(a * 3 + b, oc)] = (T)acc[a][b];
}

// ---- level-0 red-black Gauss-Seidel, one colour per launch -------------------------------------
// z_p <- (r_p + sum_q w_pq z_q) / diag_p for active p with (i + j) & 1 == colour
template <typename T>
__global__ void __launch_bounds__(256)
rbgs0_kernel(FineAcc<T> A, T* __restrict__ z, const T* __restrict__ r, int colour, const int* __restrict__ done) {
    if (done && *done) return;
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + ((i + colour) & 1);
    if (j >= A.g.cols) return;
    const long long o = (long long)i * A.g.pitch + j;
    if (A.g.flags[o] & HG_FLAG_HARD) return;
    T wu, wd, wl, wr, diag;
    A.weights(i, j, wu, wd, wl, wr, diag);
    if (!(diag > T(0))) return;
    const T inv = T(1) / diag;
    for (int k = 0; k < A.g.K; ++k) {
        T* zk = z + k * A.g.plane;
        T v = r[k * A.g.plane + o];
        if (wu != T(0)) v += wu * zk[o - A.g.pitch];
        if (wd != T(0)) v += wd * zk[o + A.g.pitch];
        if (wl != T(0)) v += wl * zk[o - 1];
        if (wr != T(0)) v += wr * zk[o + 1];
        zk[o] = v * inv;
    }
}

// ---- 9-point four-colour Gauss-Seidel, one colour per launch -----------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
mcgs_kernel(Coarse<T> L, int colour, const int* __restrict__ done) {
    if (done && *done) return;
    const int i = 2 * blockIdx.y + (colour >> 1);
    const int j = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + (colour & 1);
    if (i >= L.rows || j >= L.cols) return;
    const long long o = (long long)i * L.pitch + j;
    // inactive pixels (inside fully constrained areas) are decided on the centre coefficient alone,
    // before the other eight planes are touched
    const T c4 = L.coef[cidx(4, o)];
    if (!(c4 > T(0))) return;
    T cf[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) cf[c] = c == 4 ? c4 : L.coef[cidx(c, o)];
    const T inv = T(1) / c4;
    for (int k = 0; k < L.K; ++k) {
        T* xk = L.x + k * L.plane;
        T v = L.b[k * L.plane + o];
#pragma unroll
        for (int di = -1; di <= 1; ++di)
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                const int c = (di + 1) * 3 + (dj + 1);
                if (c != 4 && cf[c] != T(0)) v -= cf[c] * xk[o + (long long)di * L.pitch + dj];
            }
        xk[o] = v * inv;
    }
}

// ---- coarsest level: `sweeps` symmetric four-colour sweeps in one block --------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
coarsest_kernel(Coarse<T> L, int sweeps, const int* __restrict__ done) {
    if (done && *done) return;
    const int n = L.rows * L.cols;
    for (int k = 0; k < L.K; ++k) {
        T* xk = L.x + k * L.plane;
        const T* bk = L.b + k * L.plane;
        for (int e = threadIdx.x; e < L.rows * L.pitch; e += blockDim.x) xk[e] = T(0);
        __syncthreads();
        for (int s = 0; s < 2 * sweeps; ++s) {
            for (int cc = 0; cc < 4; ++cc) {
                const int colour = (s & 1) ? 3 - cc : cc;
                for (int e = threadIdx.x; e < n; e += blockDim.x) {
                    const int i = e / L.cols, j = e % L.cols;
                    if ((((i & 1) << 1) | (j & 1)) != colour) continue;
                    const long long o = (long long)i * L.pitch + j;
                    const T c4 = L.coef[cidx(4, o)];
                    if (!(c4 > T(0))) continue;
                    T v = bk[o];
                    for (int di = -1; di <= 1; ++di)
                        for (int dj = -1; dj <= 1; ++dj) {
                            const int c = (di + 1) * 3 + (dj + 1);
                            if (c == 4) continue;
                            const T a = L.coef[cidx(c, o)];
                            if (a != T(0)) v -= a * xk[o + (long long)di * L.pitch + dj];
                        }
                    xk[o] = v / c4;
                }
                __syncthreads();
            }
        }
    }
}

// ---- residual + restriction:  b_c = P^T (b_f - A_f x_f) ------------------------------------------
// one thread per coarse pixel gathers its 3x3 fine residuals (recomputed on the fly)
template <typename T, typename Acc>
__global__ void __launch_bounds__(256)
resid_restrict_kernel(Acc fine, const T* __restrict__ xf, const T* __restrict__ bf, Coarse<T> cl, int K,
                      const int* __restrict__ done) {
    if (done && *done) return;
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y;
    if (J >= cl.pitch) return;
    const long long oc = (long long)I * cl.pitch + J;
    const bool live = J < cl.cols && cl.coef[cidx(4, oc)] != T(0);
    const int nfr = fine.rows(), nfc = fine.cols();
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        if (live) {
            const T* xk = xf + k * fine.plane();
            const T* bk = bf + k * fine.plane();
            for (int a = -1; a <= 1; ++a) {
                const double wa = w1d(a, I, nfr, cl.rows);
                if (wa == 0.0) continue;
                for (int b = -1; b <= 1; ++b) {
                    const double wp = wa * w1d(b, J, nfc, cl.cols);
                    if (wp == 0.0) continue;
                    const int pi = 2 * I + a, pj = 2 * J + b;
                    if (!fine.active(pi, pj)) continue;
                    const T res = bk[(long long)pi * fine.pitch() + pj] - fine.apply(xk, pi, pj);
                    s += wp * (double)res;
                }
            }
        }
        cl.b[k * cl.plane + oc] = (T)s;
        cl.x[k * cl.plane + oc] = T(0);
    }
}

// ---- prolongation + correction:  x_f += P x_c -------------------------------------------------------
template <typename T, typename Acc>
__global__ void __launch_bounds__(256)
prolong_add_kernel(Acc fine, T* __restrict__ xf, Coarse<T> cl, int K, const int* __restrict__ done) {
    if (done && *done) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= fine.cols() || !fine.active(i, j)) return;
    int Ip[2], Jp[2];
    double wi[2], wj[2];
    const int ni = parents(i, cl.rows, Ip, wi);
    const int nj = parents(j, cl.cols, Jp, wj);
    const long long o = (long long)i * fine.pitch() + j;
    for (int k = 0; k < K; ++k) {
        const T* xc = cl.x + k * cl.plane;
        double s = 0.0;
        for (int u = 0; u < ni; ++u)
            for (int v = 0; v < nj; ++v) s += wi[u] * wj[v] * (double)xc[(long long)Ip[u] * cl.pitch + Jp[v]];
        xf[k * fine.plane() + o] += (T)s;
    }
}


// ---- exact coarsest solve: dense inverse built once at setup -------------------------------------
// M is n x 2n row-major, [A | I]; inactive pixels get identity rows.
template <typename T>
__global__ void dense_from_stencil_kernel(Coarse<T> L, double* __restrict__ M) {
    const int n = L.rows * L.cols;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * 2 * n; e += gridDim.x * blockDim.x) M[e] = 0.0;
}
template <typename T>
__global__ void dense_fill_kernel(Coarse<T> L, double* __restrict__ M) {
    const int n = L.rows * L.cols;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int i = e / L.cols, j = e % L.cols;
    const long long o = (long long)i * L.pitch + j;
    double* row = M + (long long)e * 2 * n;
    row[n + e] = 1.0;
    const T c4 = L.coef[cidx(4, o)];
    if (!(c4 > T(0))) { row[e] = 1.0; return; }
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            const T a = L.coef[cidx((di + 1) * 3 + (dj + 1), o)];
            if (a == T(0)) continue;
            const int qi = i + di, qj = j + dj;
            // couplings to inactive pixels are dropped: their value is held at 0
            if (L.coef[cidx(4, (long long)qi * L.pitch + qj)] > T(0)) row[qi * L.cols + qj] = (double)a;
        }
}
// Gauss-Jordan without pivoting, one block.  The matrix is symmetric positive SEMI-definite:
// when the masked interpolation P loses column rank (e.g. four coarse pixels that all
// interpolate one lone free fine pixel) A_c = P^T A P is singular but the coarse right-hand
// side P^T r is always consistent and null vectors are annihilated by P, so any solution
// will do.  A pivot that has collapsed (<= 1e-10 of its original diagonal) marks a dependent
// unknown: it is set to zero by clearing its row, which yields the generalised inverse on
// the independent set.
static __global__ void __launch_bounds__(1024) gauss_jordan_kernel(double* __restrict__ M, int n) {
    extern __shared__ double fac[];   // n factors + n original diagonals
    double* d0 = fac + n;
    const int w = 2 * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) d0[i] = M[(long long)i * w + i];
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        const double piv = M[(long long)k * w + k];
        const bool dead = !(piv > 1e-10 * d0[k]);
        __syncthreads();
        if (dead) {
            for (int j = threadIdx.x; j < w; j += blockDim.x) M[(long long)k * w + j] = 0.0;
            continue;
        }
        for (int j = threadIdx.x; j < w; j += blockDim.x) M[(long long)k * w + j] /= piv;
        for (int i = threadIdx.x; i < n; i += blockDim.x) fac[i] = (i == k) ? 0.0 : M[(long long)i * w + k];
        __syncthreads();
        for (long long e = threadIdx.x; e < (long long)n * w; e += blockDim.x) {
            const int i = (int)(e / w), j = (int)(e % w);
            const double f = fac[i];
            if (f != 0.0) M[e] -= f * M[(long long)k * w + j];
        }
    }
}
// x = A^-1 b on the coarsest level (right half of M), one block per channel
template <typename T>
__global__ void __launch_bounds__(256)
coarsest_dense_kernel(Coarse<T> L, const double* __restrict__ M, const int* __restrict__ done) {
    if (done && *done) return;
    const int n = L.rows * L.cols;
    const int k = blockIdx.x;
    extern __shared__ double sb[];
    for (int e = threadIdx.x; e < n; e += blockDim.x) sb[e] = (double)L.b[k * L.plane + (long long)(e / L.cols) * L.pitch + e % L.cols];
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const double* row = M + (long long)e * 2 * n + n;
        double s = 0.0;
        for (int q = 0; q < n; ++q) s += row[q] * sb[q];
        L.x[k * L.plane + (long long)(e / L.cols) * L.pitch + e % L.cols] = (T)s;
    }
}

// slab mode: rows of the all-gathered blocks ([rank][plane][maxrows x pitch]) -> the replicated global level
template <typename T>
__global__ void gather_unpack_kernel(const T* __restrict__ recv, T* __restrict__ global, long long gplane, int pitch, int maxrows,
                                     int nplanes, const int* __restrict__ gl_rows) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    const int q = blockIdx.z / nplanes, c = blockIdx.z - q * nplanes;
    if (j >= pitch || i >= gl_rows[2 * q + 1]) return;
    global[(long long)c * gplane + (long long)(gl_rows[2 * q] + i) * pitch + j] =
        recv[((long long)(q * nplanes + c) * maxrows + i) * pitch + j];
}

}  // namespace hg
