// fine.cuh -- tile kernels of the finest level (Laplacian + multigrid path).
//
// Every kernel here works on FT_X x FT_Y pixel tiles, one 256-thread block per tile, and
// consults a per-tile activity byte computed at setup: a tile whose pixels (and 4-pixel halo)
// are all hard constraints is skipped outright -- it holds no unknowns, so nothing in it ever
// changes.  On the 50 %-opaque fill workload that halves every pass.
//
//   lap_tile_kernel      q = A p (+ p.q), or r = b - A x (+ r.r)
//   cg_update_tile_kernel   x += alpha p, r -= alpha q (+ r.r)
//   cg_pupdate_tile_kernel  p = z + beta p
//   mg_down0_kernel      V-cycle, downward leg on level 0, fused:
//                           red-black Gauss-Seidel from the zero guess + residual + restriction
//   mg_up0_kernel        V-cycle, upward leg on level 0, fused:
//                           prolongation + black-red Gauss-Seidel (+ r.z)
//
// Why the two V-cycle legs need so little traffic.  With one red-then-black sweep from z = 0:
//   z_R = r_R / d_R,   z_B = (r_B + sum_R w z_R) / d_B,
// the residual afterwards vanishes at black pixels and equals sum_B w z_B at red pixels, so
// restriction needs r only (read once, halo 3) and z need not be stored.  On the way up the
// black-then-red sweep overwrites every pixel and looks only at the prolongated correction at
// red pixels, z_R + (P e)_R with z_R = r_R / d_R recomputed: the leg reads r and the coarse
// correction, and writes z once.  Fine-level V-cycle cost: 12 bytes per pixel and channel
// (fp32) instead of ~70 for separate smoothing / residual / transfer passes.
#pragma once
#include "common.cuh"
#include "mg.cuh"
#include "pcg.cuh"

namespace hg {

#define FT_X 64
#define FT_Y 32
#define FT_H 4                       // halo of the shared-memory tiles
#define FT_SW (FT_X + 2 * FT_H)      // 72
#define FT_SH (FT_Y + 2 * FT_H)      // 40
#define FT_THREADS 256

// ---- setup: per-tile activity ------------------------------------------------------------------
__global__ void tile_activity_kernel(const uint8_t* __restrict__ flags, int rows, int cols, int pitch,
                                     uint8_t* __restrict__ active) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    const int bi = blockIdx.y * FT_Y - FT_H, bj = blockIdx.x * FT_X - FT_H;
    int mine = 0;
    for (int e = threadIdx.x; e < FT_SH * FT_SW; e += blockDim.x) {
        const int i = bi + e / FT_SW, j = bj + e % FT_SW;
        if (i >= 0 && i < rows && j >= 0 && j < cols && !(flags[(long long)i * pitch + j] & HG_FLAG_HARD)) mine = 1;
    }
    if (mine) any = 1;
    __syncthreads();
    if (threadIdx.x == 0) active[blockIdx.y * gridDim.x + blockIdx.x] = (uint8_t)any;
}

// weights of pixel (i,j) from the flag tile; (li,lj) are its tile coordinates
template <typename T>
__device__ __forceinline__ void tile_weights(const Grid<T>& g, const uint8_t (*sf)[FT_SW], int li, int lj, int i, int j,
                                             T& wu, T& wd, T& wl, T& wr, T& diag) {
    const unsigned f = sf[li][lj];
    wu = i > 0 ? w_down(g, i - 1, j, sf[li - 1][lj]) : T(0);
    wd = w_down(g, i, j, f);
    wl = j > 0 ? w_right(g, i, j - 1, sf[li][lj - 1]) : T(0);
    wr = w_right(g, i, j, f);
    diag = wu + wd + wl + wr + soft_w(g, (long long)i * g.pitch + j, f);
}

__device__ __forceinline__ void load_flag_tile(const uint8_t* __restrict__ flags, int rows, int cols, int pitch, int bi, int bj,
                                               uint8_t (*sf)[FT_SW]) {
    for (int e = threadIdx.x; e < FT_SH * FT_SW; e += FT_THREADS) {
        const int li = e / FT_SW, lj = e % FT_SW;
        const int i = bi - FT_H + li, j = bj - FT_H + lj;
        sf[li][lj] = (i >= 0 && i < rows && j >= 0 && j < cols) ? flags[(long long)i * pitch + j] : (uint8_t)HG_FLAG_HARD;
    }
}

template <typename T>
__device__ __forceinline__ void load_value_tile(const T* __restrict__ v, int rows, int cols, int pitch, int bi, int bj, int halo,
                                                T (*sv)[FT_SW]) {
    const int h0 = FT_H - halo, hh = FT_Y + 2 * halo, ww = FT_X + 2 * halo;
    for (int e = threadIdx.x; e < hh * ww; e += FT_THREADS) {
        const int li = h0 + e / ww, lj = h0 + e % ww;
        const int i = bi - FT_H + li, j = bj - FT_H + lj;
        sv[li][lj] = (i >= 0 && i < rows && j >= 0 && j < cols) ? v[(long long)i * pitch + j] : T(0);
    }
}

// Per-tile operator cache, shared by all channels and all stages of a leg: inverse diagonal
// (0 at hard pixels) and the four edge weights packed as 2-bit codes, weight = 0.125 * code
// (codes 0, 1, 2 <-> 0, 1/8, 1/4).  With gradient-penalty edges (PEN) the weights are not
// representable that way and are recomputed from the flag tile instead.
struct TileOp {
    template <typename T, bool PEN>
    static __device__ __forceinline__ void weights(const Grid<T>& g, const uint8_t (*sf)[FT_SW], const uint8_t (*sw)[FT_SW],
                                                   int li, int lj, int i, int j, T& wu, T& wd, T& wl, T& wr) {
        if (PEN) {
            T d;
            tile_weights(g, sf, li, lj, i, j, wu, wd, wl, wr, d);
        } else {
            const unsigned c = sw[li][lj];
            wu = T(0.125) * T(c & 3u);
            wd = T(0.125) * T((c >> 2) & 3u);
            wl = T(0.125) * T((c >> 4) & 3u);
            wr = T(0.125) * T((c >> 6) & 3u);
        }
    }
};

template <typename T>
__device__ __forceinline__ void build_tile_op(const Grid<T>& g, const uint8_t (*sf)[FT_SW], int bi, int bj,
                                              T (*sd)[FT_SW], uint8_t (*sw)[FT_SW]) {
    // halo 3 (rows / cols 1 .. S-2 of the tile arrays)
    for (int e = threadIdx.x; e < (FT_Y + 6) * (FT_X + 6); e += FT_THREADS) {
        const int li = 1 + e / (FT_X + 6), lj = 1 + e % (FT_X + 6);
        const int i = bi - FT_H + li, j = bj - FT_H + lj;
        T dinv = T(0);
        unsigned code = 0;
        if (!(sf[li][lj] & HG_FLAG_HARD)) {
            T wu, wd, wl, wr, d;
            tile_weights(g, sf, li, lj, i, j, wu, wd, wl, wr, d);
            if (d > T(0)) dinv = T(1) / d;
            code = (unsigned)(wu * T(8)) | ((unsigned)(wd * T(8)) << 2) | ((unsigned)(wl * T(8)) << 4) | ((unsigned)(wr * T(8)) << 6);
        }
        sd[li][lj] = dinv;
        sw[li][lj] = (uint8_t)code;
    }
}

#define FT_SMEM(T) (3 * sizeof(T) * FT_SH * FT_SW + 2 * FT_SH * FT_SW)

// ---- V-cycle, downward leg ------------------------------------------------------------------------
template <typename T, bool PEN>
__global__ void __launch_bounds__(FT_THREADS)
mg_down0_kernel(Grid<T> g, const T* __restrict__ r, Coarse<T> cl, const uint8_t* __restrict__ tile_active,
                const int* __restrict__ done) {
    if (done && *done) return;
    if (!tile_active[blockIdx.y * gridDim.x + blockIdx.x]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T (*sr)[FT_SW] = reinterpret_cast<T (*)[FT_SW]>(smem_raw);
    T (*sz)[FT_SW] = reinterpret_cast<T (*)[FT_SW]>(smem_raw + sizeof(T) * FT_SH * FT_SW);
    T (*sd)[FT_SW] = reinterpret_cast<T (*)[FT_SW]>(smem_raw + 2 * sizeof(T) * FT_SH * FT_SW);
    uint8_t (*sf)[FT_SW] = reinterpret_cast<uint8_t (*)[FT_SW]>(smem_raw + 3 * sizeof(T) * FT_SH * FT_SW);
    uint8_t (*sw)[FT_SW] = reinterpret_cast<uint8_t (*)[FT_SW]>(smem_raw + 3 * sizeof(T) * FT_SH * FT_SW + FT_SH * FT_SW);
    const int bi = blockIdx.y * FT_Y, bj = blockIdx.x * FT_X;
    const int rows = g.rows, cols = g.cols;
    load_flag_tile(g.flags, rows, cols, g.pitch, bi, bj, sf);
    __syncthreads();
    build_tile_op(g, sf, bi, bj, sd, sw);

    for (int k = 0; k < g.K; ++k) {
        __syncthreads();
        load_value_tile(r + k * g.plane, rows, cols, g.pitch, bi, bj, 3, sr);
        __syncthreads();
        // red pixels, halo 3: z = r / d
        for (int e = threadIdx.x; e < (FT_Y + 6) * (FT_X + 6); e += FT_THREADS) {
            const int li = 1 + e / (FT_X + 6), lj = 1 + e % (FT_X + 6);
            if (!((bi + bj + li + lj) & 1)) sz[li][lj] = sr[li][lj] * sd[li][lj];
        }
        __syncthreads();
        // black pixels, halo 2: z = (r + sum w z_red) / d
        for (int e = threadIdx.x; e < (FT_Y + 4) * (FT_X + 4); e += FT_THREADS) {
            const int li = 2 + e / (FT_X + 4), lj = 2 + e % (FT_X + 4);
            if (!((bi + bj + li + lj) & 1)) continue;
            T z = T(0);
            const T dinv = sd[li][lj];
            if (dinv != T(0)) {
                T wu, wd, wl, wr;
                TileOp::weights<T, PEN>(g, sf, sw, li, lj, bi - FT_H + li, bj - FT_H + lj, wu, wd, wl, wr);
                z = (sr[li][lj] + wu * sz[li - 1][lj] + wd * sz[li + 1][lj] + wl * sz[li][lj - 1] + wr * sz[li][lj + 1]) * dinv;
            }
            sz[li][lj] = z;
        }
        __syncthreads();
        // residual after the sweep, rows / cols [-1, T-1]: red: sum w z_black, black: 0  (into sr)
        for (int e = threadIdx.x; e < (FT_Y + 1) * (FT_X + 1); e += FT_THREADS) {
            const int li = 3 + e / (FT_X + 1), lj = 3 + e % (FT_X + 1);
            T res = T(0);
            if (!((bi + bj + li + lj) & 1) && sd[li][lj] != T(0)) {
                T wu, wd, wl, wr;
                TileOp::weights<T, PEN>(g, sf, sw, li, lj, bi - FT_H + li, bj - FT_H + lj, wu, wd, wl, wr);
                res = wu * sz[li - 1][lj] + wd * sz[li + 1][lj] + wl * sz[li][lj - 1] + wr * sz[li][lj + 1];
            }
            sr[li][lj] = res;
        }
        __syncthreads();
        // restriction to the coarse pixels owned by this tile: centre + the four diagonals (the
        // edge midpoints are black, their residual is 0)
        for (int e = threadIdx.x; e < (FT_Y / 2) * (FT_X / 2); e += FT_THREADS) {
            const int ci = e / (FT_X / 2), cj = e % (FT_X / 2);
            const int I = bi / 2 + ci, J = bj / 2 + cj;
            if (I >= cl.rows || J >= cl.cols) continue;
            const int li = FT_H + 2 * ci, lj = FT_H + 2 * cj;
            const double wim = w1d(-1, I, rows, cl.rows), wip = w1d(1, I, rows, cl.rows);
            const double wjm = w1d(-1, J, cols, cl.cols), wjp = w1d(1, J, cols, cl.cols);
            const double s = (double)sr[li][lj] + wim * (wjm * (double)sr[li - 1][lj - 1] + wjp * (double)sr[li - 1][lj + 1]) +
                             wip * (wjm * (double)sr[li + 1][lj - 1] + wjp * (double)sr[li + 1][lj + 1]);
            const long long oc = (long long)I * cl.pitch + J;
            cl.b[k * cl.plane + oc] = (T)s;
            cl.x[k * cl.plane + oc] = T(0);
        }
    }
}

// ---- V-cycle, upward leg -----------------------------------------------------------------------------
template <typename T, bool PEN>
__global__ void __launch_bounds__(FT_THREADS)
mg_up0_kernel(Grid<T> g, const T* __restrict__ r, T* __restrict__ z, Coarse<T> cl, int have_coarse,
              const uint8_t* __restrict__ tile_active, double* __restrict__ part_rz, const int* __restrict__ done) {
    if (done && *done) return;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    if (!tile_active[blin]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T (*sr)[FT_SW] = reinterpret_cast<T (*)[FT_SW]>(smem_raw);
    T (*sz)[FT_SW] = reinterpret_cast<T (*)[FT_SW]>(smem_raw + sizeof(T) * FT_SH * FT_SW);
    T (*sd)[FT_SW] = reinterpret_cast<T (*)[FT_SW]>(smem_raw + 2 * sizeof(T) * FT_SH * FT_SW);
    uint8_t (*sf)[FT_SW] = reinterpret_cast<uint8_t (*)[FT_SW]>(smem_raw + 3 * sizeof(T) * FT_SH * FT_SW);
    uint8_t (*sw)[FT_SW] = reinterpret_cast<uint8_t (*)[FT_SW]>(smem_raw + 3 * sizeof(T) * FT_SH * FT_SW + FT_SH * FT_SW);
    __shared__ double red[32];
    const int bi = blockIdx.y * FT_Y, bj = blockIdx.x * FT_X;
    const int rows = g.rows, cols = g.cols;
    const int ntiles = gridDim.x * gridDim.y;
    load_flag_tile(g.flags, rows, cols, g.pitch, bi, bj, sf);
    __syncthreads();
    build_tile_op(g, sf, bi, bj, sd, sw);

    for (int k = 0; k < g.K; ++k) {
        __syncthreads();
        load_value_tile(r + k * g.plane, rows, cols, g.pitch, bi, bj, 2, sr);
        __syncthreads();
        const T* xc = cl.x + k * cl.plane;
        // red pixels, halo 2: z = r / d + (P e)
        for (int e = threadIdx.x; e < (FT_Y + 4) * (FT_X + 4); e += FT_THREADS) {
            const int li = 2 + e / (FT_X + 4), lj = 2 + e % (FT_X + 4);
            if ((bi + bj + li + lj) & 1) continue;
            const T dinv = sd[li][lj];
            T zz = T(0);
            if (dinv != T(0)) {
                zz = sr[li][lj] * dinv;
                if (have_coarse) {
                    const int i = bi - FT_H + li, j = bj - FT_H + lj;
                    int Ip[2], Jp[2];
                    double wi[2], wj[2];
                    const int ni = parents(i, cl.rows, Ip, wi), nj = parents(j, cl.cols, Jp, wj);
                    double s = 0.0;
                    for (int u = 0; u < ni; ++u)
                        for (int v = 0; v < nj; ++v) s += wi[u] * wj[v] * (double)xc[(long long)Ip[u] * cl.pitch + Jp[v]];
                    zz += (T)s;
                }
            }
            sz[li][lj] = zz;
        }
        __syncthreads();
        // black pixels, halo 1
        for (int e = threadIdx.x; e < (FT_Y + 2) * (FT_X + 2); e += FT_THREADS) {
            const int li = 3 + e / (FT_X + 2), lj = 3 + e % (FT_X + 2);
            if (!((bi + bj + li + lj) & 1)) continue;
            const T dinv = sd[li][lj];
            T zz = T(0);
            if (dinv != T(0)) {
                T wu, wd, wl, wr;
                TileOp::weights<T, PEN>(g, sf, sw, li, lj, bi - FT_H + li, bj - FT_H + lj, wu, wd, wl, wr);
                zz = (sr[li][lj] + wu * sz[li - 1][lj] + wd * sz[li + 1][lj] + wl * sz[li][lj - 1] + wr * sz[li][lj + 1]) * dinv;
            }
            sz[li][lj] = zz;
        }
        __syncthreads();
        // red pixels of the tile
        for (int e = threadIdx.x; e < FT_Y * FT_X; e += FT_THREADS) {
            const int li = FT_H + e / FT_X, lj = FT_H + e % FT_X;
            if ((bi + bj + li + lj) & 1) continue;
            const T dinv = sd[li][lj];
            T zz = T(0);
            if (dinv != T(0)) {
                T wu, wd, wl, wr;
                TileOp::weights<T, PEN>(g, sf, sw, li, lj, bi - FT_H + li, bj - FT_H + lj, wu, wd, wl, wr);
                zz = (sr[li][lj] + wu * sz[li - 1][lj] + wd * sz[li + 1][lj] + wl * sz[li][lj - 1] + wr * sz[li][lj + 1]) * dinv;
            }
            sz[li][lj] = zz;
        }
        __syncthreads();
        // write the tile, accumulate r.z
        double acc = 0.0;
        T* zk = z + k * g.plane;
        for (int e = threadIdx.x; e < FT_Y * FT_X; e += FT_THREADS) {
            const int li = FT_H + e / FT_X, lj = FT_H + e % FT_X;
            const int i = bi + e / FT_X, j = bj + e % FT_X;
            if (i < rows && j < g.pitch) {
                const T zz = j < cols ? sz[li][lj] : T(0);
                zk[(long long)i * g.pitch + j] = zz;
                acc += (double)zz * (double)sr[li][lj];
            }
        }
        const double s = block_sum(acc, red);
        if (threadIdx.x == 0) part_rz[(long long)k * ntiles + blin] = s;
    }
}

// ---- q = A p (+ p.q)   or   r = b - A x (+ r.r) -----------------------------------------------------------
// thread -> 4 consecutive columns of rows ty and ty + 16 of the tile; all loads are independent
// (rows above / below come straight from L1 / L2), which keeps many requests in flight.
template <typename T, int MODE, int DOT>
__global__ void __launch_bounds__(FT_THREADS)
lap_tile_kernel(Grid<T> g, const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ y,
                const uint8_t* __restrict__ tile_active, double* __restrict__ partial, const int* __restrict__ done) {
    if (done && *done) return;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    if (!tile_active[blin]) return;
    __shared__ double red[32];
    const int ntiles = gridDim.x * gridDim.y;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int j0 = blockIdx.x * FT_X + 4 * tx;
    const int pitch = g.pitch;
    // operator of this thread's 2 x 4 pixels, shared by all channels
    T wu[2][4], wd[2][4], wl[2][4], wr[2][4], dg[2][4];
    bool hard[2][4], live[2], work[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = blockIdx.y * FT_Y + ty + 16 * h;
        live[h] = i < g.rows && j0 < pitch;
        work[h] = false;
        if (!live[h]) continue;
        const long long o = (long long)i * pitch + j0;
        const uchar4 f4 = *reinterpret_cast<const uchar4*>(g.flags + o);
        const unsigned fc[4] = {f4.x, f4.y, f4.z, f4.w};
        unsigned fu[4] = {0, 0, 0, 0};
        if (i > 0) {
            const uchar4 u4 = *reinterpret_cast<const uchar4*>(g.flags + o - pitch);
            fu[0] = u4.x; fu[1] = u4.y; fu[2] = u4.z; fu[3] = u4.w;
        }
        const unsigned fl = j0 > 0 ? g.flags[o - 1] : 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + c;
            wu[h][c] = i > 0 ? w_down(g, i - 1, j, fu[c]) : T(0);
            wd[h][c] = w_down(g, i, j, fc[c]);
            wl[h][c] = j > 0 ? w_right(g, i, j - 1, c == 0 ? fl : fc[c - 1]) : T(0);
            wr[h][c] = w_right(g, i, j, fc[c]);
            dg[h][c] = wu[h][c] + wd[h][c] + wl[h][c] + wr[h][c] + soft_w(g, o + c, fc[c]);
            hard[h][c] = (fc[c] & HG_FLAG_HARD) != 0;
            work[h] = work[h] || !hard[h][c];
        }
    }
    for (int k = 0; k < g.K; ++k) {
        const T* xk = x + k * g.plane;
        double acc = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!live[h]) continue;
            const int i = blockIdx.y * FT_Y + ty + 16 * h;
            const long long o = (long long)i * pitch + j0;
            V4<T> out = zero4<T>();
            if (work[h]) {
                const V4<T> cur = ld4(xk + o);
                const V4<T> up = i > 0 ? ld4(xk + o - pitch) : zero4<T>();
                const V4<T> dn = i + 1 < g.rows ? ld4(xk + o + pitch) : zero4<T>();
                const T xl = j0 > 0 ? xk[o - 1] : T(0);
                const T xr = j0 + 4 < pitch ? xk[o + 4] : T(0);
                V4<T> bv = zero4<T>();
                if (MODE == 1 && b) bv = ld4(b + k * g.plane + o);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const T xL = c == 0 ? xl : cur.v[c - 1];
                    const T xR = c == 3 ? xr : cur.v[c + 1];
                    T v = dg[h][c] * cur.v[c] - wu[h][c] * up.v[c] - wd[h][c] * dn.v[c] - wl[h][c] * xL - wr[h][c] * xR;
                    if (MODE == 1) v = bv.v[c] - v;
                    if (hard[h][c]) v = T(0);
                    out.v[c] = v;
                    if (DOT == 1) acc += (double)cur.v[c] * (double)v;
                    if (DOT == 2) acc += (double)v * (double)v;
                }
            }
            st4(y + k * g.plane + o, out);
        }
        if (DOT != 0) {
            const double s = block_sum(acc, red);
            if (threadIdx.x == 0) partial[(long long)k * ntiles + blin] = s;
        }
    }
}

// ---- x += alpha p ; r -= alpha q ; r.r ------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(FT_THREADS)
cg_update_tile_kernel(const CgState* __restrict__ st, Grid<T> g, T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p,
                      const T* __restrict__ q, const uint8_t* __restrict__ tile_active, double* __restrict__ part_rr) {
    if (st->done) return;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    if (!tile_active[blin]) return;
    __shared__ double red[32];
    const int ntiles = gridDim.x * gridDim.y;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int j0 = blockIdx.x * FT_X + 4 * tx;
    for (int k = 0; k < g.K; ++k) {
        const T a = (T)st->alpha[k];
        double srr = 0.0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = blockIdx.y * FT_Y + ty + 16 * half;
            if (i < g.rows && j0 < g.pitch) {
                const long long o = k * g.plane + (long long)i * g.pitch + j0;
                V4<T> xv = ld4(x + o), rv = ld4(r + o);
                const V4<T> pv = ld4(p + o), qv = ld4(q + o);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    xv.v[c] += a * pv.v[c];
                    rv.v[c] -= a * qv.v[c];
                    srr += (double)rv.v[c] * (double)rv.v[c];
                }
                st4(x + o, xv);
                st4(r + o, rv);
            }
        }
        const double s = block_sum(srr, red);
        if (threadIdx.x == 0) part_rr[(long long)k * ntiles + blin] = s;
    }
}

// ---- p = z + beta p -----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(FT_THREADS)
cg_pupdate_tile_kernel(const CgState* __restrict__ st, Grid<T> g, T* __restrict__ p, const T* __restrict__ z,
                       const uint8_t* __restrict__ tile_active, int init) {
    if (st->done) return;
    if (!tile_active[blockIdx.y * gridDim.x + blockIdx.x]) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int j0 = blockIdx.x * FT_X + 4 * tx;
    for (int k = 0; k < g.K; ++k) {
        const T bt = init ? T(0) : (T)st->beta[k];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = blockIdx.y * FT_Y + ty + 16 * half;
            if (i < g.rows && j0 < g.pitch) {
                const long long o = k * g.plane + (long long)i * g.pitch + j0;
                const V4<T> zv = ld4(z + o);
                V4<T> pv = zv;
                if (!init) {
                    pv = ld4(p + o);
#pragma unroll
                    for (int c = 0; c < 4; ++c) pv.v[c] = zv.v[c] + bt * pv.v[c];
                }
                st4(p + o, pv);
            }
        }
    }
}

}  // namespace hg
