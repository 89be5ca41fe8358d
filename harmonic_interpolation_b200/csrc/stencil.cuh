// stencil.cuh -- matrix-free operator kernels.
//
//   lap_apply_kernel    masked 5-point symmetric Laplacian        (recovery.py:232-347)
//   bilap_apply_kernel  13-point bilaplacian  L_refl^T L_refl      (recovery.py:349-500, 911)
//
// Both add the soft-constraint diagonal and the gradient-penalty edges
// (recovery.py:926-927, 1280-1283) and apply the hard-constraint elimination
// (recovery.py:647-777) by producing 0 at hard pixels:
//
//   F(x)_p = sum_q w_pq (x_p - x_q)  [or (Bx)_p]  + s_p x_p + sum_pen w_g (x_p - x_q)   p free
//   F(x)_p = 0                                                                        p hard
//
// MODE 0: y = F(x).   MODE 1: y = b - F(x) at free pixels (residual), 0 at hard pixels.
// DOT 0: none.  DOT 1: per-block partial sums of x.y (p.Ap).  DOT 2: of y.y.
// Partial sums are fp64 and land in partial[k * nblocks + block]; a second kernel adds
// them in a fixed order, so results do not depend on scheduling.
#pragma once
#include "common.cuh"

namespace hg {

// --------------------------------------------------------------------------------------
// 5-point Laplacian.  One thread owns 4 consecutive columns (one float4) and marches down
// LAP_ROWS rows keeping the rows above / at / below in registers, so each row of x is
// fetched once per block (plus one halo row per LAP_ROWS).  Left / right neighbours come
// from the adjacent lanes by shuffle; only the two edge lanes of a warp issue an extra
// scalar load.
// --------------------------------------------------------------------------------------
#define LAP_TX 128
#define LAP_ROWS 32

template <typename T, int MODE, int DOT>
__global__ void __launch_bounds__(LAP_TX)
lap_apply_kernel(Grid<T> g, const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ y,
                 double* __restrict__ partial, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double red[32];
    const int lane = threadIdx.x & 31;
    const int j0 = (blockIdx.x * LAP_TX + threadIdx.x) * 4;
    const bool active = j0 < g.pitch;
    const int i0 = blockIdx.y * LAP_ROWS;
    const int i1 = min(i0 + LAP_ROWS, g.rows);
    const int pitch = g.pitch;
    const int nblk = gridDim.x * gridDim.y;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;

    for (int k = 0; k < g.K; ++k) {
        const T* xk = x + (long long)k * g.plane;
        const T* bk = (MODE == 1) ? b + (long long)k * g.plane : nullptr;
        T* yk = y + (long long)k * g.plane;
        double acc = 0.0;

        V4<T> up = zero4<T>(), cur = zero4<T>(), dn;
        uchar4 fup = make_uchar4(0, 0, 0, 0), fcur = make_uchar4(0, 0, 0, 0), fdn;
        if (active) {
            const long long o = (long long)i0 * pitch + j0;
            cur = ld4(xk + o);
            fcur = *reinterpret_cast<const uchar4*>(g.flags + o);
            if (i0 > 0) {
                up = ld4(xk + o - pitch);
                fup = *reinterpret_cast<const uchar4*>(g.flags + o - pitch);
            }
        }
        for (int i = i0; i < i1; ++i) {
            const long long o = (long long)i * pitch + j0;
            dn = zero4<T>();
            fdn = make_uchar4(0, 0, 0, 0);
            if (active && i + 1 < g.rows) {
                dn = ld4(xk + o + pitch);
                fdn = *reinterpret_cast<const uchar4*>(g.flags + o + pitch);
            }
            T xl = __shfl_up_sync(HG_FULL, cur.v[3], 1);
            T xr = __shfl_down_sync(HG_FULL, cur.v[0], 1);
            unsigned fl = __shfl_up_sync(HG_FULL, (unsigned)fcur.w, 1);
            if (active) {
                if (lane == 0) {
                    xl = j0 > 0 ? xk[o - 1] : T(0);
                    fl = j0 > 0 ? g.flags[o - 1] : 0u;
                }
                if (lane == 31) xr = (j0 + 4 < pitch) ? xk[o + 4] : T(0);
                const unsigned fc[4] = {fcur.x, fcur.y, fcur.z, fcur.w};
                const unsigned fu[4] = {fup.x, fup.y, fup.z, fup.w};
                V4<T> out;
                V4<T> bv = zero4<T>();
                if (MODE == 1) bv = ld4(bk + o);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = j0 + c;
                    const T xc = cur.v[c];
                    const T xL = c == 0 ? xl : cur.v[c - 1];
                    const T xR = c == 3 ? xr : cur.v[c + 1];
                    const unsigned fL = c == 0 ? fl : fc[c - 1];
                    const T wu = (i > 0) ? w_down(g, i - 1, j, fu[c]) : T(0);
                    const T wd = w_down(g, i, j, fc[c]);
                    const T wl = (j > 0) ? w_right(g, i, j - 1, fL) : T(0);
                    const T wr = w_right(g, i, j, fc[c]);
                    T v = wu * (xc - up.v[c]) + wd * (xc - dn.v[c]) + wl * (xc - xL) + wr * (xc - xR);
                    v += soft_w(g, o + c, fc[c]) * xc;
                    if (MODE == 1) v = bv.v[c] - v;
                    if (fc[c] & HG_FLAG_HARD) v = T(0);
                    out.v[c] = v;
                    if (DOT == 1) acc += (double)xc * (double)v;
                    if (DOT == 2) acc += (double)v * (double)v;
                }
                st4(yk + o, out);
            }
            up = cur; cur = dn; fup = fcur; fcur = fdn;
        }
        if (DOT != 0) {
            const double s = block_sum(acc, red);
            if (threadIdx.x == 0) partial[(long long)k * nblk + blin] = s;
        }
    }
}

// --------------------------------------------------------------------------------------
// 13-point bilaplacian, fused: a (TY+4) x (TX+4) tile of x (halo 2) is staged in shared
// memory, u = L_refl x is formed on the halo-1 ring it is needed on, and y = L_refl^T u
// is produced without u ever going to HBM.
//   u_p      = 1/4 sum_axis a_axis(p) (2 x_p - x_{p-e} - x_{p+e})
//   (L^T u)_p = 1/4 sum_axis [ 2 a(p) u_p - a(p-e) u_{p-e} - a(p+e) u_{p+e} ]
// a_axis(p) = 1 iff both neighbours of p along the axis exist and both edges are uncut.
// --------------------------------------------------------------------------------------
template <typename T> struct BilapTile { static constexpr int TX = 64, TY = 32; };
template <> struct BilapTile<double> { static constexpr int TX = 64, TY = 16; };
#define BILAP_THREADS 256

template <typename T, int MODE, int DOT>
__global__ void __launch_bounds__(BILAP_THREADS)
bilap_apply_kernel(Grid<T> g, const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ y,
                   double* __restrict__ partial, const int* __restrict__ done) {
    if (done && *done) return;
    constexpr int TX = BilapTile<T>::TX, TY = BilapTile<T>::TY;
    constexpr int XW = TX + 4, XH = TY + 4;
    __shared__ T sx[XH][XW];
    __shared__ T sv[TY + 2][TX];      // a_v u on rows bi-1 .. bi+TY
    __shared__ T sh[TY][TX + 2];      // a_h u on cols bj-1 .. bj+TX
    __shared__ uint8_t sf[XH][XW];
    __shared__ double red[32];

    const int tid = threadIdx.x;
    const int bi = blockIdx.y * TY, bj = blockIdx.x * TX;
    const int rows = g.rows, cols = g.cols, pitch = g.pitch;
    const int nblk = gridDim.x * gridDim.y;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;

    for (int e = tid; e < XH * XW; e += BILAP_THREADS) {
        const int li = e / XW, lj = e % XW;
        const int i = bi - 2 + li, j = bj - 2 + lj;
        sf[li][lj] = (i >= 0 && i < rows && j >= 0 && j < cols) ? g.flags[(long long)i * pitch + j] : 0;
    }
    __syncthreads();

    // axis masks in tile coordinates (li, lj) <-> (bi - 2 + li, bj - 2 + lj)
    auto a_v = [&](int li, int lj) -> bool {
        const int i = bi - 2 + li, j = bj - 2 + lj;
        return i >= 1 && i <= rows - 2 && j >= 0 && j < cols &&
               !(sf[li - 1][lj] & HG_FLAG_CUT_DOWN) && !(sf[li][lj] & HG_FLAG_CUT_DOWN);
    };
    auto a_h = [&](int li, int lj) -> bool {
        const int i = bi - 2 + li, j = bj - 2 + lj;
        return j >= 1 && j <= cols - 2 && i >= 0 && i < rows &&
               !(sf[li][lj - 1] & HG_FLAG_CUT_RIGHT) && !(sf[li][lj] & HG_FLAG_CUT_RIGHT);
    };

    for (int k = 0; k < g.K; ++k) {
        const T* xk = x + (long long)k * g.plane;
        T* yk = y + (long long)k * g.plane;
        __syncthreads();
        for (int e = tid; e < XH * XW; e += BILAP_THREADS) {
            const int li = e / XW, lj = e % XW;
            const int i = bi - 2 + li, j = bj - 2 + lj;
            sx[li][lj] = (i >= 0 && i < rows && j >= 0 && j < cols) ? xk[(long long)i * pitch + j] : T(0);
        }
        __syncthreads();
        auto u_at = [&](int li, int lj, bool av, bool ah) -> T {
            T u = T(0);
            const T c2 = T(2) * sx[li][lj];
            if (av) u += c2 - sx[li - 1][lj] - sx[li + 1][lj];
            if (ah) u += c2 - sx[li][lj - 1] - sx[li][lj + 1];
            return T(0.25) * u;
        };
        for (int e = tid; e < (TY + 2) * TX; e += BILAP_THREADS) {
            const int r = e / TX, c = e % TX;
            const int li = r + 1, lj = c + 2;
            const bool av = a_v(li, lj);
            sv[r][c] = av ? u_at(li, lj, true, a_h(li, lj)) : T(0);
        }
        for (int e = tid; e < TY * (TX + 2); e += BILAP_THREADS) {
            const int r = e / (TX + 2), c = e % (TX + 2);
            const int li = r + 2, lj = c + 1;
            const bool ah = a_h(li, lj);
            sh[r][c] = ah ? u_at(li, lj, a_v(li, lj), true) : T(0);
        }
        __syncthreads();
        double acc = 0.0;
        for (int e = tid; e < TY * TX; e += BILAP_THREADS) {
            const int r = e / TX, c = e % TX;
            const int i = bi + r, j = bj + c;
            if (i >= rows || j >= pitch) continue;
            const long long o = (long long)i * pitch + j;
            T v = T(0);
            if (j < cols) {
                const int li = r + 2, lj = c + 2;
                const unsigned f = sf[li][lj];
                if (!(f & HG_FLAG_HARD)) {
                    const T xc = sx[li][lj];
                    v = T(0.25) * (T(2) * sv[r + 1][c] - sv[r][c] - sv[r + 2][c] +
                                   T(2) * sh[r][c + 1] - sh[r][c] - sh[r][c + 2]);
                    v += soft_w(g, o, f) * xc;
                    v += pen_down(g, i, f) * (xc - sx[li + 1][lj]);
                    v += pen_right(g, j, f) * (xc - sx[li][lj + 1]);
                    if (i > 0) v += pen_down(g, i - 1, sf[li - 1][lj]) * (xc - sx[li - 1][lj]);
                    if (j > 0) v += pen_right(g, j - 1, sf[li][lj - 1]) * (xc - sx[li][lj - 1]);
                    if (MODE == 1) v = b[(long long)k * g.plane + o] - v;
                    if (DOT == 1) acc += (double)xc * (double)v;
                    if (DOT == 2) acc += (double)v * (double)v;
                }
            }
            yk[o] = v;
        }
        if (DOT != 0) {
            const double s = block_sum(acc, red);
            if (tid == 0) partial[(long long)k * nblk + blin] = s;
        }
    }
}

// --------------------------------------------------------------------------------------
// setup kernels: inverse diagonal of the system (Jacobi preconditioner) and the domain
// checks of the reflected operator (recovery.py:493, 495).
// --------------------------------------------------------------------------------------
__device__ __forceinline__ bool axis_v(const uint8_t* fl, int rows, int cols, int pitch, int i, int j) {
    return i >= 1 && i <= rows - 2 && j >= 0 && j < cols &&
           !(fl[(long long)(i - 1) * pitch + j] & HG_FLAG_CUT_DOWN) && !(fl[(long long)i * pitch + j] & HG_FLAG_CUT_DOWN);
}
__device__ __forceinline__ bool axis_h(const uint8_t* fl, int rows, int cols, int pitch, int i, int j) {
    return j >= 1 && j <= cols - 2 && i >= 0 && i < rows &&
           !(fl[(long long)i * pitch + j - 1] & HG_FLAG_CUT_RIGHT) && !(fl[(long long)i * pitch + j] & HG_FLAG_CUT_RIGHT);
}

template <typename T>
__global__ void diag_kernel(Grid<T> g, int bilap, T* __restrict__ dinv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.pitch) return;
    const long long o = (long long)i * g.pitch + j;
    T d = T(0);
    if (j < g.cols) {
        const unsigned f = g.flags[o];
        if (!(f & HG_FLAG_HARD)) {
            const unsigned fu = i > 0 ? g.flags[o - g.pitch] : 0u;
            const unsigned fl = j > 0 ? g.flags[o - 1] : 0u;
            if (!bilap) {
                d = w_down(g, i, j, f) + w_right(g, i, j, f);
                if (i > 0) d += w_down(g, i - 1, j, fu);
                if (j > 0) d += w_right(g, i, j - 1, fl);
            } else {
                const int av = axis_v(g.flags, g.rows, g.cols, g.pitch, i, j);
                const int ah = axis_h(g.flags, g.rows, g.cols, g.pitch, i, j);
                const int n = axis_v(g.flags, g.rows, g.cols, g.pitch, i - 1, j) + axis_v(g.flags, g.rows, g.cols, g.pitch, i + 1, j) +
                              axis_h(g.flags, g.rows, g.cols, g.pitch, i, j - 1) + axis_h(g.flags, g.rows, g.cols, g.pitch, i, j + 1);
                d = T(0.25) * T((av + ah) * (av + ah)) + T(0.0625) * T(n);
                d += pen_down(g, i, f) + pen_right(g, j, f);
                if (i > 0) d += pen_down(g, i - 1, fu);
                if (j > 0) d += pen_right(g, j - 1, fl);
            }
            d += soft_w(g, o, f);
            d = d > T(0) ? T(1) / d : T(0);
        }
    }
    dinv[o] = d;
}

// counts[0] = pixels with no axis term (all-zero rows of L_refl), counts[1] = pixels
// referenced by no row, counts[2] = hard pixels, counts[3] = any cut bit seen
static __global__ void domain_check_kernel(const uint8_t* __restrict__ fl, int rows, int cols, int pitch, int bilap,
                                    unsigned long long* __restrict__ counts) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= cols) return;
    const unsigned f = fl[(long long)i * pitch + j];
    if (f & HG_FLAG_HARD) atomicAdd(&counts[2], 1ull);
    if (f & (HG_FLAG_CUT_DOWN | HG_FLAG_CUT_RIGHT)) atomicAdd(&counts[3], 1ull);
    if (!bilap) return;
    const bool av = axis_v(fl, rows, cols, pitch, i, j), ah = axis_h(fl, rows, cols, pitch, i, j);
    if (!(av || ah)) {
        atomicAdd(&counts[0], 1ull);
        const bool ref = axis_v(fl, rows, cols, pitch, i - 1, j) || axis_v(fl, rows, cols, pitch, i + 1, j) ||
                         axis_h(fl, rows, cols, pitch, i, j - 1) || axis_h(fl, rows, cols, pitch, i, j + 1);
        if (!ref) atomicAdd(&counts[1], 1ull);
    }
}


// --------------------------------------------------------------------------------------
// Master-precision residual for the fp32 mode (iterative refinement): x and b are kept in
// fp64, r = b - F(x) is evaluated in fp64 and rounded to the storage type T only at the end.
// Runs once per outer refinement step (2-4 times per solve), so it is a plain
// one-thread-per-pixel kernel.  Also leaves per-block partial sums of r.r.
// --------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double u_master(const Grid<T>& g, const double* x, int i, int j) {
    // (L_refl x)_(i,j); 0 outside the grid
    if (i < 0 || i >= g.rows || j < 0 || j >= g.cols) return 0.0;
    const long long o = (long long)i * g.pitch + j;
    double u = 0.0;
    if (axis_v(g.flags, g.rows, g.cols, g.pitch, i, j)) u += 2.0 * x[o] - x[o - g.pitch] - x[o + g.pitch];
    if (axis_h(g.flags, g.rows, g.cols, g.pitch, i, j)) u += 2.0 * x[o] - x[o - 1] - x[o + 1];
    return 0.25 * u;
}

template <typename T>
__global__ void __launch_bounds__(256)
resid_master_kernel(Grid<T> g, int bilap, const double* __restrict__ x, const double* __restrict__ b,
                    T* __restrict__ r, double* __restrict__ partial) {
    __shared__ double red[32];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    const int nblk = gridDim.x * gridDim.y;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    const bool inside = j < g.pitch;
    const long long o = (long long)i * g.pitch + j;
    bool free_px = false;
    unsigned f = 0, fu = 0, fl = 0;
    if (inside && j < g.cols) {
        f = g.flags[o];
        free_px = !(f & HG_FLAG_HARD);
        fu = i > 0 ? g.flags[o - g.pitch] : 0u;
        fl = j > 0 ? g.flags[o - 1] : 0u;
    }
    for (int k = 0; k < g.K; ++k) {
        const double* xk = x + k * g.plane;
        double res = 0.0;
        if (free_px) {
            const double xc = xk[o];
            double v = (double)soft_w(g, o, f) * xc;
            if (!bilap) {
                if (i > 0) v += (double)w_down(g, i - 1, j, fu) * (xc - xk[o - g.pitch]);
                if (i + 1 < g.rows) v += (double)w_down(g, i, j, f) * (xc - xk[o + g.pitch]);
                if (j > 0) v += (double)w_right(g, i, j - 1, fl) * (xc - xk[o - 1]);
                if (j + 1 < g.cols) v += (double)w_right(g, i, j, f) * (xc - xk[o + 1]);
            } else {
                const int av = axis_v(g.flags, g.rows, g.cols, g.pitch, i, j), ah = axis_h(g.flags, g.rows, g.cols, g.pitch, i, j);
                double t = 0.0;
                if (av || ah) t += 2.0 * (av + ah) * u_master(g, xk, i, j);
                if (axis_v(g.flags, g.rows, g.cols, g.pitch, i - 1, j)) t -= u_master(g, xk, i - 1, j);
                if (axis_v(g.flags, g.rows, g.cols, g.pitch, i + 1, j)) t -= u_master(g, xk, i + 1, j);
                if (axis_h(g.flags, g.rows, g.cols, g.pitch, i, j - 1)) t -= u_master(g, xk, i, j - 1);
                if (axis_h(g.flags, g.rows, g.cols, g.pitch, i, j + 1)) t -= u_master(g, xk, i, j + 1);
                v += 0.25 * t;
                if (i > 0) v += (double)pen_down(g, i - 1, fu) * (xc - xk[o - g.pitch]);
                if (i + 1 < g.rows) v += (double)pen_down(g, i, f) * (xc - xk[o + g.pitch]);
                if (j > 0) v += (double)pen_right(g, j - 1, fl) * (xc - xk[o - 1]);
                if (j + 1 < g.cols) v += (double)pen_right(g, j, f) * (xc - xk[o + 1]);
            }
            res = b[k * g.plane + o] - v;
        }
        if (inside) r[k * g.plane + o] = (T)res;
        const double s = block_sum(res * res, red);
        if (threadIdx.x == 0) partial[(long long)k * nblk + blin] = s;
    }
}

// x (fp64 master) += e (storage type); hard pixels and padding have e == 0
template <typename T>
__global__ void __launch_bounds__(256)
axpy_master_kernel(long long n, double* __restrict__ x, const T* __restrict__ e) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x)
        x[v] += (double)e[v];
}

}  // namespace hg
