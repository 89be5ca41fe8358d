// fine.cuh -- tile kernels of the finest level (Laplacian + multigrid path).
//
// Every kernel here works on aligned FT_X x FT_Y pixel tiles and consults a per-tile activity
// byte computed at setup: a tile without a single free pixel holds no unknowns -- every vector of
// the solver is identically zero there and stays so -- and is skipped outright.  On the
// 50 %-opaque fill workload that halves every pass.
//
// The level-0 operator is cached at setup as ONE BYTE per pixel: four 2-bit codes, one per edge
// (up, down, left, right), weight = code / 8, i.e. 0, 1/8 or 1/4 (recovery.py:263-299, 309-319).
// A pixel whose byte is 0xAA (all four 1/4) and that is neither hard nor soft takes a fast path.
// Gradient-penalty edges (weight w_g) do not fit the code; then (PEN) the weights are recomputed
// from the flag bytes.
//
//   lap_tile_kernel          q = A p (+ p.q), or r = b - A x (+ r.r); x, b may be fp64 while the
//                            result is fp32 (master-precision residual of the fp32 mode)
//   cg_update_tile_kernel    x += alpha p, r -= alpha q (+ r.r)
//   cg_pupdate_tile_kernel   p = z + beta p
//   mg_down0_kernel          V-cycle, downward leg on level 0, fused:
//                               red-black Gauss-Seidel from the zero guess + residual + restriction
//   mg_up0_kernel            V-cycle, upward leg on level 0, fused:
//                               prolongation + black-red Gauss-Seidel (+ r.z)
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
#define FT_THREADS 256
#define HG_HS4 0x21212121u           // HG_FLAG_HARD | HG_FLAG_SOFT in each byte of a uchar4

// ---- setup ------------------------------------------------------------------------------------------
// activity of the aligned tiles: any free pixel within the tile grown by `halo`
static __global__ void tile_activity_kernel(const uint8_t* __restrict__ flags, int rows, int cols, int pitch, int halo,
                                     uint8_t* __restrict__ active) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    const int bi = blockIdx.y * FT_Y - halo, bj = blockIdx.x * FT_X - halo;
    const int w = FT_X + 2 * halo, h = FT_Y + 2 * halo;
    int mine = 0;
    for (int e = threadIdx.x; e < w * h; e += blockDim.x) {
        const int i = bi + e / w, j = bj + e % w;
        if (i >= 0 && i < rows && j >= 0 && j < cols && !(flags[(long long)i * pitch + j] & HG_FLAG_HARD)) mine = 1;
    }
    if (mine) any = 1;
    __syncthreads();
    if (threadIdx.x == 0) active[blockIdx.y * gridDim.x + blockIdx.x] = (uint8_t)any;
}

// one byte of edge-weight codes per pixel (0 in the padding and at hard pixels)
template <typename T>
__global__ void opcode_kernel(Grid<T> g, uint8_t* __restrict__ code) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.pitch) return;
    const long long o = (long long)i * g.pitch + j;
    unsigned c = 0;
    if (j < g.cols) {
        const unsigned f = g.flags[o];
        if (!(f & HG_FLAG_HARD)) {
            Grid<T> h = g;
            h.w_g = T(0);     // codes describe the Laplacian edges only
            const T wu = i > 0 ? w_down(h, i - 1, j, g.flags[o - g.pitch]) : T(0);
            const T wd = w_down(h, i, j, f);
            const T wl = j > 0 ? w_right(h, i, j - 1, g.flags[o - 1]) : T(0);
            const T wr = w_right(h, i, j, f);
            c = (unsigned)(wu * T(8)) | ((unsigned)(wd * T(8)) << 2) | ((unsigned)(wl * T(8)) << 4) | ((unsigned)(wr * T(8)) << 6);
        }
    }
    code[o] = (uint8_t)c;
}

template <typename T>
__device__ __forceinline__ void decode(unsigned c, T& wu, T& wd, T& wl, T& wr) {
    wu = T(0.125) * T(c & 3u);
    wd = T(0.125) * T((c >> 2) & 3u);
    wl = T(0.125) * T((c >> 4) & 3u);
    wr = T(0.125) * T((c >> 6) & 3u);
}

// the four edge weights of pixel (i,j) incl. gradient-penalty edges, from the flag bytes
template <typename T>
__device__ __forceinline__ void flag_weights(const Grid<T>& g, int i, int j, unsigned f, unsigned fu, unsigned fl,
                                             T& wu, T& wd, T& wl, T& wr) {
    wu = i > 0 ? w_down(g, i - 1, j, fu) : T(0);
    wd = w_down(g, i, j, f);
    wl = j > 0 ? w_right(g, i, j - 1, fl) : T(0);
    wr = w_right(g, i, j, f);
}

// ==================================================================================================
// Fused V-cycle legs.  A block owns one aligned 64 x 32 tile; its shared-memory REGION is the tile
// plus a halo H on every side.  Every stage sweeps the whole region with a fixed slot -> pixel
// mapping and the valid part shrinks by one ring per stage, ending exactly on the tile.  Values
// live in colour-split arrays (red = (i+j) even): consecutive lanes touch consecutive words, and a
// slot owns one red and one black pixel of a row, so no lane idles in a single-colour stage.
//
//   slot (li, t):   red pixel  lj = 2 t + (li & 1)      -> index t      (s = li & 1)
//                   black pixel lj = 2 t + 1 - (li & 1)  -> index t
//   neighbours of black (li, t): red[li-1][t], red[li+1][t], red[li][t - s], red[li][t + 1 - s]
//   neighbours of red   (li, t): blk[li-1][t], blk[li+1][t], blk[li][t - 1 + s], blk[li][t + s]
// ==================================================================================================
template <int H> struct Region {
    static constexpr int W = FT_X + 2 * H, RH = FT_Y + 2 * H;   // region size (W even)
    static constexpr int PW = W / 2;                            // red / black pairs per row
    static constexpr int NS = PW * RH;                          // slots
    static constexpr int SW = PW + 2, SH = RH + 2;              // split arrays with one guard ring
};

template <typename T, int H, bool PEN>
struct RegionSmem {
    using R = Region<H>;
    T rr[R::SH][R::SW], rb[R::SH][R::SW];        // right-hand side, red / black
    T zr[R::SH][R::SW], zb[R::SH][R::SW];        // iterate
    T dr[R::SH][R::SW], db[R::SH][R::SW];        // inverse diagonal (0 = hard / dead)
    uint8_t cr[R::SH][R::SW], cb[R::SH][R::SW];  // weight codes
    T wgt[PEN ? 2 : 1][PEN ? 4 : 1][PEN ? R::SH : 1][PEN ? R::SW : 1];   // explicit weights (PEN only)
};

// inverse diagonal and weight codes of every region pixel, from the code / flag planes
template <typename T, int H, bool PEN>
__device__ __forceinline__ void region_build_op(const Grid<T>& g, const uint8_t* __restrict__ code, int gi0, int gj0,
                                                RegionSmem<T, H, PEN>& S) {
    using R = Region<H>;
    for (int e = threadIdx.x; e < R::RH * R::W; e += FT_THREADS) {
        const int li = e / R::W, lj = e - li * R::W;
        const int i = gi0 + li, j = gj0 + lj;
        T dinv = T(0);
        unsigned c = 0;
        T wu = T(0), wd = T(0), wl = T(0), wr = T(0);
        if (i >= 0 && i < g.rows && j >= 0 && j < g.cols) {
            const long long o = (long long)i * g.pitch + j;
            const unsigned f = g.flags[o];
            if (!(f & HG_FLAG_HARD)) {
                T d;
                if (PEN) {
                    flag_weights(g, i, j, f, i > 0 ? g.flags[o - g.pitch] : 0u, j > 0 ? g.flags[o - 1] : 0u, wu, wd, wl, wr);
                    d = wu + wd + wl + wr;
                } else {
                    c = code[o];
                    d = T(0.125) * T((c & 3u) + ((c >> 2) & 3u) + ((c >> 4) & 3u) + ((c >> 6) & 3u));
                }
                if (f & HG_FLAG_SOFT) d += soft_w(g, o, f);
                if (d > T(0)) dinv = T(1) / d;
            }
        }
        const int a = li + 1, t = (lj >> 1) + 1;
        if ((li + lj) & 1) { S.db[a][t] = dinv; S.cb[a][t] = (uint8_t)c; }
        else               { S.dr[a][t] = dinv; S.cr[a][t] = (uint8_t)c; }
        if (PEN) {
            const int col = (li + lj) & 1;
            S.wgt[col][0][a][t] = wu; S.wgt[col][1][a][t] = wd; S.wgt[col][2][a][t] = wl; S.wgt[col][3][a][t] = wr;
        }
    }
}

// loads one channel of a value plane into the colour-split arrays (zeros outside the grid)
template <typename T, int H>
__device__ __forceinline__ void region_load(const T* __restrict__ v, int rows, int cols, int pitch, int gi0, int gj0,
                                            T (*red)[Region<H>::SW], T (*blk)[Region<H>::SW]) {
    using R = Region<H>;
    for (int e = threadIdx.x; e < R::RH * R::W; e += FT_THREADS) {
        const int li = e / R::W, lj = e - li * R::W;
        const int i = gi0 + li, j = gj0 + lj;
        const T val = (i >= 0 && i < rows && j >= 0 && j < cols) ? v[(long long)i * pitch + j] : T(0);
        if ((li + lj) & 1) blk[li + 1][(lj >> 1) + 1] = val;
        else red[li + 1][(lj >> 1) + 1] = val;
    }
}

// sum over the four neighbours of  weight * value  for the pixel (row a, index c) of colour `col`
// (0 red, 1 black); `other` is the array of the opposite colour, s = (a - 1) & 1
template <typename T, int H, bool PEN>
__device__ __forceinline__ T region_nsum(const RegionSmem<T, H, PEN>& S, int col, const T (*other)[Region<H>::SW],
                                         int a, int c, int s, unsigned code) {
    const int l = col == 0 ? c - 1 + s : c - s;
    const T vu = other[a - 1][c], vd = other[a + 1][c], vl = other[a][l], vr = other[a][l + 1];
    if (PEN) return S.wgt[col][0][a][c] * vu + S.wgt[col][1][a][c] * vd + S.wgt[col][2][a][c] * vl + S.wgt[col][3][a][c] * vr;
    if (code == 0xAAu) return T(0.25) * ((vu + vd) + (vl + vr));
    T wu, wd, wl, wr;
    decode(code, wu, wd, wl, wr);
    return wu * vu + wd * vd + wl * vl + wr * vr;
}

// ---- V-cycle, downward leg: RB Gauss-Seidel from zero + residual + restriction -----------------------
#define DOWN_H 3
template <typename T, bool PEN>
__global__ void __launch_bounds__(FT_THREADS)
mg_down0_kernel(Grid<T> g, const uint8_t* __restrict__ code, const T* __restrict__ r, Coarse<T> cl,
                const uint8_t* __restrict__ active, const int* __restrict__ done) {
    if (done && *done) return;
    if (!active[blockIdx.y * gridDim.x + blockIdx.x]) return;
    using R = Region<DOWN_H>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RegionSmem<T, DOWN_H, PEN>& S = *reinterpret_cast<RegionSmem<T, DOWN_H, PEN>*>(smem_raw);
    const int gi0 = blockIdx.y * FT_Y - DOWN_H, gj0 = blockIdx.x * FT_X - DOWN_H;
    region_build_op<T, DOWN_H, PEN>(g, code, gi0, gj0, S);

    for (int k = 0; k < g.K; ++k) {
        __syncthreads();
        region_load<T, DOWN_H>(r + k * g.plane, g.rows, g.cols, g.pitch, gi0, gj0, S.rr, S.rb);
        __syncthreads();
        // red: z = r / d
        for (int sl = threadIdx.x; sl < R::NS; sl += FT_THREADS) {
            const int li = sl / R::PW, a = li + 1, c = sl - li * R::PW + 1;
            S.zr[a][c] = S.rr[a][c] * S.dr[a][c];
        }
        __syncthreads();
        // black: z = (r + sum w z_red) / d
        for (int sl = threadIdx.x; sl < R::NS; sl += FT_THREADS) {
            const int li = sl / R::PW, a = li + 1, c = sl - li * R::PW + 1;
            const T dinv = S.db[a][c];
            T z = T(0);
            if (dinv != T(0)) z = (S.rb[a][c] + region_nsum<T, DOWN_H, PEN>(S, 1, S.zr, a, c, li & 1, S.cb[a][c])) * dinv;
            S.zb[a][c] = z;
        }
        __syncthreads();
        // residual after the sweep: red: sum w z_black (black: 0); overwrites rr
        for (int sl = threadIdx.x; sl < R::NS; sl += FT_THREADS) {
            const int li = sl / R::PW, a = li + 1, c = sl - li * R::PW + 1;
            T res = T(0);
            if (S.dr[a][c] != T(0)) res = region_nsum<T, DOWN_H, PEN>(S, 0, S.zb, a, c, li & 1, S.cr[a][c]);
            S.rr[a][c] = res;
        }
        __syncthreads();
        // restriction to the 16 x 32 coarse pixels of the tile: centre + the four diagonals (all red)
        for (int e = threadIdx.x; e < (FT_Y / 2) * (FT_X / 2); e += FT_THREADS) {
            const int ci = e >> 5, cj = e & 31;
            const int I = blockIdx.y * (FT_Y / 2) + ci, J = blockIdx.x * (FT_X / 2) + cj;
            if (I >= cl.rows || J >= cl.cols) continue;
            const int li = DOWN_H + 2 * ci, lj = DOWN_H + 2 * cj;     // region coordinates of fine (2I, 2J): both odd
            const int a = li + 1, cc = (lj >> 1) + 1;
            const int cl_ = ((lj - 1) >> 1) + 1, cr_ = ((lj + 1) >> 1) + 1;
            const T wim = (T)w1d(-1, I, g.rows, cl.rows), wip = (T)w1d(1, I, g.rows, cl.rows);
            const T wjm = (T)w1d(-1, J, g.cols, cl.cols), wjp = (T)w1d(1, J, g.cols, cl.cols);
            const T sres = S.rr[a][cc] + wim * (wjm * S.rr[a - 1][cl_] + wjp * S.rr[a - 1][cr_]) +
                           wip * (wjm * S.rr[a + 1][cl_] + wjp * S.rr[a + 1][cr_]);
            const long long oc = (long long)I * cl.pitch + J;
            cl.b[k * cl.plane + oc] = sres;
            cl.x[k * cl.plane + oc] = T(0);
        }
    }
}

// ---- V-cycle, upward leg: prolongation + BR Gauss-Seidel + r.z ----------------------------------------------
#define UP_H 2
template <typename T, bool PEN>
__global__ void __launch_bounds__(FT_THREADS)
mg_up0_kernel(Grid<T> g, const uint8_t* __restrict__ code, const T* __restrict__ r, T* __restrict__ z, Coarse<T> cl,
              int have_coarse, const uint8_t* __restrict__ active, double* __restrict__ part_rz, const int* __restrict__ done) {
    if (done && *done) return;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    if (!active[blin]) return;
    using R = Region<UP_H>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RegionSmem<T, UP_H, PEN>& S = *reinterpret_cast<RegionSmem<T, UP_H, PEN>*>(smem_raw);
    __shared__ T se[R::RH / 2 + 1][R::PW + 1];    // coarse correction under the region (replicated at the far edge)
    __shared__ double red[32];
    const int gi0 = blockIdx.y * FT_Y - UP_H, gj0 = blockIdx.x * FT_X - UP_H;     // both even
    const int ntiles = gridDim.x * gridDim.y;
    region_build_op<T, UP_H, PEN>(g, code, gi0, gj0, S);

    for (int k = 0; k < g.K; ++k) {
        __syncthreads();
        region_load<T, UP_H>(r + k * g.plane, g.rows, g.cols, g.pitch, gi0, gj0, S.rr, S.rb);
        if (have_coarse) {
            const T* xc = cl.x + k * cl.plane;
            for (int e = threadIdx.x; e < (R::RH / 2 + 1) * (R::PW + 1); e += FT_THREADS) {
                const int a = e / (R::PW + 1), b = e - a * (R::PW + 1);
                const int I = min(max(gi0 / 2 + a, 0), cl.rows - 1), J = min(max(gj0 / 2 + b, 0), cl.cols - 1);
                se[a][b] = xc[(long long)I * cl.pitch + J];
            }
        }
        __syncthreads();
        // red: z = r / d + (P e)        (red pixels are (even, even) or (odd, odd))
        for (int sl = threadIdx.x; sl < R::NS; sl += FT_THREADS) {
            const int li = sl / R::PW, t = sl - li * R::PW, a = li + 1, c = t + 1;
            const T dinv = S.dr[a][c];
            T zz = T(0);
            if (dinv != T(0)) {
                zz = S.rr[a][c] * dinv;
                if (have_coarse) {
                    const int ci = li >> 1;
                    zz += (li & 1) ? T(0.25) * ((se[ci][t] + se[ci][t + 1]) + (se[ci + 1][t] + se[ci + 1][t + 1])) : se[ci][t];
                }
            }
            S.zr[a][c] = zz;
        }
        __syncthreads();
        // black: z = (r + sum w z_red) / d
        for (int sl = threadIdx.x; sl < R::NS; sl += FT_THREADS) {
            const int li = sl / R::PW, a = li + 1, c = sl - li * R::PW + 1;
            const T dinv = S.db[a][c];
            T zz = T(0);
            if (dinv != T(0)) zz = (S.rb[a][c] + region_nsum<T, UP_H, PEN>(S, 1, S.zr, a, c, li & 1, S.cb[a][c])) * dinv;
            S.zb[a][c] = zz;
        }
        __syncthreads();
        // red again: z = (r + sum w z_black) / d
        for (int sl = threadIdx.x; sl < R::NS; sl += FT_THREADS) {
            const int li = sl / R::PW, a = li + 1, c = sl - li * R::PW + 1;
            const T dinv = S.dr[a][c];
            T zz = T(0);
            if (dinv != T(0)) zz = (S.rr[a][c] + region_nsum<T, UP_H, PEN>(S, 0, S.zb, a, c, li & 1, S.cr[a][c])) * dinv;
            S.zr[a][c] = zz;
        }
        __syncthreads();
        // write the tile, accumulate r.z
        double acc = 0.0;
        T* zk = z + k * g.plane;
        for (int e = threadIdx.x; e < FT_Y * FT_X; e += FT_THREADS) {
            const int ii = e >> 6, jj = e & 63;
            const int i = blockIdx.y * FT_Y + ii, j = blockIdx.x * FT_X + jj;
            if (i >= g.rows || j >= g.cols) continue;
            const int li = ii + UP_H, lj = jj + UP_H;
            const bool blk = (li + lj) & 1;
            const T zz = blk ? S.zb[li + 1][(lj >> 1) + 1] : S.zr[li + 1][(lj >> 1) + 1];
            const T rv = blk ? S.rb[li + 1][(lj >> 1) + 1] : S.rr[li + 1][(lj >> 1) + 1];
            zk[(long long)i * g.pitch + j] = zz;
            acc += (double)zz * (double)rv;
        }
        const double sdot = block_sum(acc, red);
        if (threadIdx.x == 0) part_rz[(long long)k * ntiles + blin] = sdot;
    }
}

// ---- q = A p (+ p.q)   or   r = b - A x (+ r.r) -----------------------------------------------------------
// One block = the upper or lower 64 x 16 half of a tile for KC channels (blockIdx.z selects the
// channel group when K > KC), one thread = 4 consecutive columns of one row.  (Measured: one channel
// per block, 48 registers and 3x the blocks, is 40 % SLOWER -- per-block latency dominates, so a block
// should carry as many bytes in flight as its registers allow.)  All loads -- flag / code bytes and the five value loads per
// channel -- are issued together, before any arithmetic, so a block pays ONE memory latency; the
// block reduction runs once, after the store.  Partial sums: partial[k * nblocks + block].
// TX is the type of x (and b): T, or double for the master-precision residual of the fp32 mode.
#define LT_ROWS 16
template <typename T, typename TX, int MODE, int DOT, int KC, bool PEN>
__global__ void __launch_bounds__(FT_THREADS, sizeof(TX) == 8 ? 2 : 3)
lap_tile_kernel(Grid<T> g, const uint8_t* __restrict__ code, const TX* __restrict__ x, const TX* __restrict__ b,
                T* __restrict__ y, const uint8_t* __restrict__ tile_active, double* __restrict__ partial,
                const int* __restrict__ done) {
    if (done && *done) return;
    if (!tile_active[(blockIdx.y >> 1) * gridDim.x + blockIdx.x]) return;
    __shared__ double red[8 * KC];
    const int nblk = gridDim.x * gridDim.y;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int j0 = blockIdx.x * FT_X + 4 * tx;
    const int i = blockIdx.y * LT_ROWS + ty;
    const int pitch = g.pitch;
    const bool live = i < g.rows && j0 < pitch;
    const long long o = (long long)i * pitch + j0;
    unsigned f4 = 0x01010101u, c4 = 0;
    if (live) {
        f4 = *reinterpret_cast<const unsigned*>(g.flags + o);
        if (!PEN) c4 = *reinterpret_cast<const unsigned*>(code + o);
    }
    double acc[KC];
    {
        const int k0 = blockIdx.z * KC;      // this block's channel group
        V4<TX> cur[KC], up[KC], dn[KC], bv[KC];
        TX xl[KC], xr[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            cur[q] = up[q] = dn[q] = bv[q] = zero4<TX>();
            xl[q] = xr[q] = TX(0);
            if (live && k0 + q < g.K) {
                const TX* xk = x + (k0 + q) * g.plane;
                cur[q] = ld4(xk + o);
                if (i > 0) up[q] = ld4(xk + o - pitch);
                if (i + 1 < g.rows) dn[q] = ld4(xk + o + pitch);
                if (j0 > 0) xl[q] = xk[o - 1];
                if (j0 + 4 < pitch) xr[q] = xk[o + 4];
                if (MODE == 1 && b) bv[q] = ld4(b + (k0 + q) * g.plane + o);
            }
        }
        // operator of the four pixels (after the value loads have been issued: one memory latency, not two)
        TX wu[4], wd[4], wl[4], wr[4], dg[4];
        const bool work = (f4 & 0x01010101u) != 0x01010101u;              // some pixel is not hard
        const bool fast = !PEN && c4 == 0xAAAAAAAAu && (f4 & HG_HS4) == 0; // four regular interior pixels
        if (live && work && !fast) {
            unsigned fu4 = 0, fl = 0;
            if (PEN) {
                if (i > 0) fu4 = *reinterpret_cast<const unsigned*>(g.flags + o - pitch);
                if (j0 > 0) fl = g.flags[o - 1];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const unsigned f = (f4 >> (8 * c)) & 0xffu;
                T a, bb, cc, dd;
                if (PEN) flag_weights(g, i, j0 + c, f, (fu4 >> (8 * c)) & 0xffu, c == 0 ? fl : (f4 >> (8 * (c - 1))) & 0xffu, a, bb, cc, dd);
                else decode((c4 >> (8 * c)) & 0xffu, a, bb, cc, dd);
                wu[c] = (TX)a; wd[c] = (TX)bb; wl[c] = (TX)cc; wr[c] = (TX)dd;
                dg[c] = wu[c] + wd[c] + wl[c] + wr[c];
                if (f & HG_FLAG_SOFT) dg[c] += (TX)soft_w(g, o + c, f);
            }
        }
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            acc[q] = 0.0;
            if (live && k0 + q < g.K) {
                V4<T> out = zero4<T>();
                TX part = TX(0);     // the thread's four products in the operand type, then fp64 (conversions are slow)
                if (work) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const TX xL = c == 0 ? xl[q] : cur[q].v[c - 1];
                        const TX xR = c == 3 ? xr[q] : cur[q].v[c + 1];
                        TX v;
                        if (fast) v = cur[q].v[c] - TX(0.25) * ((up[q].v[c] + dn[q].v[c]) + (xL + xR));
                        else v = dg[c] * cur[q].v[c] - wu[c] * up[q].v[c] - wd[c] * dn[q].v[c] - wl[c] * xL - wr[c] * xR;
                        if (MODE == 1) v = bv[q].v[c] - v;
                        if ((f4 >> (8 * c)) & HG_FLAG_HARD) v = TX(0);
                        out.v[c] = (T)v;
                        if (DOT == 1) part += cur[q].v[c] * v;
                        if (DOT == 2) part += v * v;
                    }
                }
                acc[q] = (double)part;
                st4(y + (k0 + q) * g.plane + o, out);
            }
        }
        if (DOT != 0) {
            block_sum_n<KC>(acc, red);
            if (threadIdx.x == 0) {
#pragma unroll
                for (int q = 0; q < KC; ++q)
                    if (k0 + q < g.K) partial[(long long)(k0 + q) * nblk + blin] = acc[q];
            }
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

// x (fp64 master) += e (storage type), tile-wise
template <typename T>
__global__ void __launch_bounds__(FT_THREADS)
axpy_master_tile_kernel(Grid<T> g, double* __restrict__ x, const T* __restrict__ e, const uint8_t* __restrict__ tile_active) {
    if (!tile_active[blockIdx.y * gridDim.x + blockIdx.x]) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int j0 = blockIdx.x * FT_X + 4 * tx;
    for (int k = 0; k < g.K; ++k) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = blockIdx.y * FT_Y + ty + 16 * half;
            if (i < g.rows && j0 < g.pitch) {
                const long long o = k * g.plane + (long long)i * g.pitch + j0;
                V4<double> xv = ld4(x + o);
                const V4<T> ev = ld4(e + o);
#pragma unroll
                for (int c = 0; c < 4; ++c) xv.v[c] += (double)ev.v[c];
                st4(x + o, xv);
            }
        }
    }
}

}  // namespace hg
