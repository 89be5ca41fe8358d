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

// ==================================================================================================
// Fused V-cycle legs.  Lean formulation: every block owns a fixed 64 x 32 pixel REGION of shared
// memory (interior tile + halo H on every side), every stage sweeps the whole region with a
// fixed thread -> pixel mapping (no index division), and the valid part shrinks by one ring per
// stage, ending exactly on the interior.  Values live in colour-split arrays (red = (i+j) even),
// so the 32 lanes of a warp touch 32 consecutive words, and every lane owns one red and one black
// pixel per row: no lane idles in a single-colour stage.
//
//   thread (tx, ty), row li = ty + 8 m:   red pixel  lj = 2 tx + (li & 1)      -> index tx
//                                         black pixel lj = 2 tx + 1 - (li & 1)  -> index tx
//   neighbours of black (li, tx): red[li-1][tx], red[li+1][tx], red[li][tx - s], red[li][tx + 1 - s]
//   neighbours of red   (li, tx): blk[li-1][tx], blk[li+1][tx], blk[li][tx - 1 + s], blk[li][tx + s]    (s = li & 1)
//
// The operator is cached per pixel as an inverse diagonal (0 = hard / dead) and a byte of four
// 2-bit weight codes (weight = code / 8); code 0xAA (all four 1/4) takes the fast path.  With
// gradient-penalty edges (PEN) the four weights are stored as values instead.
// ==================================================================================================
#define RG_W 64
#define RG_H 32
#define RG_SW 34          // split arrays: 32 + one guard column on each side
#define RG_SH 34          //               32 + one guard row on each side

template <typename T, bool PEN>
struct RegionSmem {
    T rr[RG_SH][RG_SW], rb[RG_SH][RG_SW];        // right-hand side, red / black
    T zr[RG_SH][RG_SW], zb[RG_SH][RG_SW];        // iterate
    T dr[RG_SH][RG_SW], db[RG_SH][RG_SW];        // inverse diagonal
    uint8_t cr[RG_SH][RG_SW], cb[RG_SH][RG_SW];  // weight codes
    uint8_t sf[RG_H + 2][RG_W + 2];              // flags, region + one ring
    T wgt[PEN ? 2 : 1][PEN ? 4 : 1][PEN ? RG_SH : 1][PEN ? RG_SW : 1];   // explicit weights (PEN only)
};

// tiling of a leg with halo H: interior (64 - 2H) x (32 - 2H)
template <int H> struct LegTiling {
    static constexpr int IW = RG_W - 2 * H, IH = RG_H - 2 * H;
    static __host__ __device__ int tiles_x(int cols) { return (cols + IW - 1) / IW; }
    static __host__ __device__ int tiles_y(int rows) { return (rows + IH - 1) / IH; }
};

// activity of the regions of a leg tiling: any free pixel in the region
template <int H>
__global__ void region_activity_kernel(const uint8_t* __restrict__ flags, int rows, int cols, int pitch,
                                       uint8_t* __restrict__ active) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    const int gi0 = blockIdx.y * LegTiling<H>::IH - H, gj0 = blockIdx.x * LegTiling<H>::IW - H;
    int mine = 0;
    for (int e = threadIdx.x; e < RG_H * RG_W; e += blockDim.x) {
        const int i = gi0 + (e >> 6), j = gj0 + (e & 63);
        if (i >= 0 && i < rows && j >= 0 && j < cols && !(flags[(long long)i * pitch + j] & HG_FLAG_HARD)) mine = 1;
    }
    if (mine) any = 1;
    __syncthreads();
    if (threadIdx.x == 0) active[blockIdx.y * gridDim.x + blockIdx.x] = (uint8_t)any;
}

// flags of the region + one ring, then inverse diagonal and weight codes of every region pixel
template <typename T, bool PEN>
__device__ __forceinline__ void region_build_op(const Grid<T>& g, int gi0, int gj0, RegionSmem<T, PEN>& S) {
    const int tid = threadIdx.x;
    for (int e = tid; e < (RG_H + 2) * (RG_W + 2); e += FT_THREADS) {
        const int a = e / (RG_W + 2), b = e % (RG_W + 2);
        const int i = gi0 - 1 + a, j = gj0 - 1 + b;
        S.sf[a][b] = (i >= 0 && i < g.rows && j >= 0 && j < g.cols) ? g.flags[(long long)i * g.pitch + j] : (uint8_t)HG_FLAG_HARD;
    }
    __syncthreads();
    const int tx = tid & 31, ty = tid >> 5;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int li = ty + 8 * m, s = li & 1;
#pragma unroll
        for (int col = 0; col < 2; ++col) {          // col 0: the red pixel of this lane, col 1: the black one
            const int lj = 2 * tx + (col == 0 ? s : 1 - s);
            const int i = gi0 + li, j = gj0 + lj;
            const unsigned f = S.sf[li + 1][lj + 1];
            T dinv = T(0);
            unsigned code = 0;
            T wu = T(0), wd = T(0), wl = T(0), wr = T(0);
            if (!(f & HG_FLAG_HARD)) {
                wu = i > 0 ? w_down(g, i - 1, j, S.sf[li][lj + 1]) : T(0);
                wd = w_down(g, i, j, f);
                wl = j > 0 ? w_right(g, i, j - 1, S.sf[li + 1][lj]) : T(0);
                wr = w_right(g, i, j, f);
                const T d = wu + wd + wl + wr + soft_w(g, (long long)i * g.pitch + j, f);
                if (d > T(0)) dinv = T(1) / d;
                if (!PEN) code = (unsigned)(wu * T(8)) | ((unsigned)(wd * T(8)) << 2) | ((unsigned)(wl * T(8)) << 4) | ((unsigned)(wr * T(8)) << 6);
            }
            if (col == 0) { S.dr[li + 1][tx + 1] = dinv; S.cr[li + 1][tx + 1] = (uint8_t)code; }
            else          { S.db[li + 1][tx + 1] = dinv; S.cb[li + 1][tx + 1] = (uint8_t)code; }
            if (PEN) {
                S.wgt[col][0][li + 1][tx + 1] = wu; S.wgt[col][1][li + 1][tx + 1] = wd;
                S.wgt[col][2][li + 1][tx + 1] = wl; S.wgt[col][3][li + 1][tx + 1] = wr;
            }
        }
    }
}

// loads one channel of a value plane into the colour-split arrays (zeros outside the grid)
template <typename T>
__device__ __forceinline__ void region_load(const T* __restrict__ v, int rows, int cols, int pitch, int gi0, int gj0,
                                            T (*red)[RG_SW], T (*blk)[RG_SW]) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int li = ty + 8 * m, i = gi0 + li;
        const bool rowok = i >= 0 && i < rows;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int lj = tx + 32 * hh, j = gj0 + lj;
            const T val = (rowok && j >= 0 && j < cols) ? v[(long long)i * pitch + j] : T(0);
            if ((li + lj) & 1) blk[li + 1][(lj >> 1) + 1] = val;
            else red[li + 1][(lj >> 1) + 1] = val;
        }
    }
}

// sum over the four neighbours of  weight * value  for the pixel (row a = li + 1, index c = tx + 1)
// of colour `col` (0 red, 1 black); `other` is the array of the opposite colour
template <typename T, bool PEN>
__device__ __forceinline__ T region_nsum(const RegionSmem<T, PEN>& S, int col, const T (*other)[RG_SW], int a, int c, int s,
                                         unsigned code) {
    // red:  left idx c - 1 + s, right idx c + s;   black: left idx c - s, right idx c + 1 - s
    const int l = col == 0 ? c - 1 + s : c - s;
    const T vu = other[a - 1][c], vd = other[a + 1][c], vl = other[a][l], vr = other[a][l + 1];
    if (PEN) return S.wgt[col][0][a][c] * vu + S.wgt[col][1][a][c] * vd + S.wgt[col][2][a][c] * vl + S.wgt[col][3][a][c] * vr;
    if (code == 0xAAu) return T(0.25) * ((vu + vd) + (vl + vr));
    return T(0.125) * (T(code & 3u) * vu + T((code >> 2) & 3u) * vd + T((code >> 4) & 3u) * vl + T((code >> 6) & 3u) * vr);
}

// ---- V-cycle, downward leg: RB Gauss-Seidel from zero + residual + restriction -----------------------
#define DOWN_H 3
template <typename T, bool PEN>
__global__ void __launch_bounds__(FT_THREADS)
mg_down0_kernel(Grid<T> g, const T* __restrict__ r, Coarse<T> cl, const uint8_t* __restrict__ region_active,
                const int* __restrict__ done) {
    if (done && *done) return;
    if (!region_active[blockIdx.y * gridDim.x + blockIdx.x]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RegionSmem<T, PEN>& S = *reinterpret_cast<RegionSmem<T, PEN>*>(smem_raw);
    using Tl = LegTiling<DOWN_H>;
    const int gi0 = blockIdx.y * Tl::IH - DOWN_H, gj0 = blockIdx.x * Tl::IW - DOWN_H;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = tx + 1;
    region_build_op<T, PEN>(g, gi0, gj0, S);

    for (int k = 0; k < g.K; ++k) {
        __syncthreads();
        region_load(r + k * g.plane, g.rows, g.cols, g.pitch, gi0, gj0, S.rr, S.rb);
        __syncthreads();
        // red: z = r / d
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int a = ty + 8 * m + 1;
            S.zr[a][c] = S.rr[a][c] * S.dr[a][c];
        }
        __syncthreads();
        // black: z = (r + sum w z_red) / d
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int a = ty + 8 * m + 1, s = (a - 1) & 1;
            const T dinv = S.db[a][c];
            T z = T(0);
            if (dinv != T(0)) z = (S.rb[a][c] + region_nsum<T, PEN>(S, 1, S.zr, a, c, s, S.cb[a][c])) * dinv;
            S.zb[a][c] = z;
        }
        __syncthreads();
        // residual after the sweep: red: sum w z_black (black: 0); overwrites rr
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int a = ty + 8 * m + 1, s = (a - 1) & 1;
            T res = T(0);
            if (S.dr[a][c] != T(0)) res = region_nsum<T, PEN>(S, 0, S.zb, a, c, s, S.cr[a][c]);
            S.rr[a][c] = res;
        }
        __syncthreads();
        // restriction to the coarse pixels of the interior: centre + the four diagonals (all red)
        const int I0 = (blockIdx.y * Tl::IH) >> 1, J0 = (blockIdx.x * Tl::IW) >> 1;
        for (int e = threadIdx.x; e < (Tl::IH / 2) * (Tl::IW / 2); e += FT_THREADS) {
            const int ci = e / (Tl::IW / 2), cj = e % (Tl::IW / 2);
            const int I = I0 + ci, J = J0 + cj;
            if (I >= cl.rows || J >= cl.cols) continue;
            const int li = DOWN_H + 2 * ci, lj = DOWN_H + 2 * cj;     // region coordinates of fine (2I, 2J)
            const int a = li + 1, s = li & 1, cc = (lj >> 1) + 1;
            // diagonal neighbours live in rows a -+ 1 at indices (lj -+ 1) >> 1
            const int cl_ = ((lj - 1) >> 1) + 1, cr_ = ((lj + 1) >> 1) + 1;
            const T wim = (T)w1d(-1, I, g.rows, cl.rows), wip = (T)w1d(1, I, g.rows, cl.rows);
            const T wjm = (T)w1d(-1, J, g.cols, cl.cols), wjp = (T)w1d(1, J, g.cols, cl.cols);
            const T sres = S.rr[a][cc] + wim * (wjm * S.rr[a - 1][cl_] + wjp * S.rr[a - 1][cr_]) +
                           wip * (wjm * S.rr[a + 1][cl_] + wjp * S.rr[a + 1][cr_]);
            (void)s;
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
mg_up0_kernel(Grid<T> g, const T* __restrict__ r, T* __restrict__ z, Coarse<T> cl, int have_coarse,
              const uint8_t* __restrict__ region_active, double* __restrict__ part_rz, const int* __restrict__ done) {
    if (done && *done) return;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    if (!region_active[blin]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RegionSmem<T, PEN>& S = *reinterpret_cast<RegionSmem<T, PEN>*>(smem_raw);
    __shared__ T se[RG_H / 2 + 1][RG_W / 2 + 1];    // coarse correction under the region (replicated at the far edge)
    __shared__ double red[32];
    using Tl = LegTiling<UP_H>;
    const int gi0 = blockIdx.y * Tl::IH - UP_H, gj0 = blockIdx.x * Tl::IW - UP_H;     // both even
    const int ntiles = gridDim.x * gridDim.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = tx + 1;
    region_build_op<T, PEN>(g, gi0, gj0, S);

    for (int k = 0; k < g.K; ++k) {
        __syncthreads();
        region_load(r + k * g.plane, g.rows, g.cols, g.pitch, gi0, gj0, S.rr, S.rb);
        if (have_coarse) {
            const T* xc = cl.x + k * cl.plane;
            for (int e = threadIdx.x; e < (RG_H / 2 + 1) * (RG_W / 2 + 1); e += FT_THREADS) {
                const int a = e / (RG_W / 2 + 1), b = e % (RG_W / 2 + 1);
                const int I = min(max(gi0 / 2 + a, 0), cl.rows - 1), J = min(max(gj0 / 2 + b, 0), cl.cols - 1);
                se[a][b] = xc[(long long)I * cl.pitch + J];
            }
        }
        __syncthreads();
        // red: z = r / d + (P e)        (red pixels are (even, even) or (odd, odd))
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int li = ty + 8 * m, a = li + 1, s = li & 1;
            const T dinv = S.dr[a][c];
            T zz = T(0);
            if (dinv != T(0)) {
                zz = S.rr[a][c] * dinv;
                if (have_coarse) {
                    const int ci = li >> 1;      // lj = 2 tx + s  ->  coarse column tx (+1 when odd)
                    zz += s ? T(0.25) * ((se[ci][tx] + se[ci][tx + 1]) + (se[ci + 1][tx] + se[ci + 1][tx + 1])) : se[ci][tx];
                }
            }
            S.zr[a][c] = zz;
        }
        __syncthreads();
        // black: z = (r + sum w z_red) / d
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int a = ty + 8 * m + 1, s = (a - 1) & 1;
            const T dinv = S.db[a][c];
            T zz = T(0);
            if (dinv != T(0)) zz = (S.rb[a][c] + region_nsum<T, PEN>(S, 1, S.zr, a, c, s, S.cb[a][c])) * dinv;
            S.zb[a][c] = zz;
        }
        __syncthreads();
        // red again: z = (r + sum w z_black) / d
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int a = ty + 8 * m + 1, s = (a - 1) & 1;
            const T dinv = S.dr[a][c];
            T zz = T(0);
            if (dinv != T(0)) zz = (S.rr[a][c] + region_nsum<T, PEN>(S, 0, S.zb, a, c, s, S.cr[a][c])) * dinv;
            S.zr[a][c] = zz;
        }
        __syncthreads();
        // write the interior, accumulate r.z
        double acc = 0.0;
        T* zk = z + k * g.plane;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int li = ty + 8 * m, i = gi0 + li;
            if (li < UP_H || li >= RG_H - UP_H || i >= g.rows) continue;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int lj = tx + 32 * hh, j = gj0 + lj;
                if (lj < UP_H || lj >= RG_W - UP_H || j >= g.cols) continue;
                const bool blk = (li + lj) & 1;
                const T zz = blk ? S.zb[li + 1][(lj >> 1) + 1] : S.zr[li + 1][(lj >> 1) + 1];
                const T rv = blk ? S.rb[li + 1][(lj >> 1) + 1] : S.rr[li + 1][(lj >> 1) + 1];
                zk[(long long)i * g.pitch + j] = zz;
                acc += (double)zz * (double)rv;
            }
        }
        const double sdot = block_sum(acc, red);
        if (threadIdx.x == 0) part_rz[(long long)k * ntiles + blin] = sdot;
    }
}

// ---- q = A p (+ p.q)   or   r = b - A x (+ r.r) -----------------------------------------------------------
// One block = the upper or lower 64 x 16 half of a tile, one thread = 4 consecutive columns of one
// row.  All loads of up to KC channels are issued before any arithmetic and are mutually independent
// (the rows above / below come straight from L1 / L2), so a warp keeps ~5 KC requests in flight; the
// block reduction runs once, after the last store.  Partial sums: partial[k * 2 ntiles + block].
#define LT_ROWS 16
template <typename T, int MODE, int DOT, int KC>
__global__ void __launch_bounds__(FT_THREADS, 3)
lap_tile_kernel(Grid<T> g, const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ y,
                const uint8_t* __restrict__ tile_active, double* __restrict__ partial, const int* __restrict__ done) {
    if (done && *done) return;
    if (!tile_active[(blockIdx.y >> 1) * gridDim.x + blockIdx.x]) return;
    __shared__ double red[32];
    const int nblk = gridDim.x * gridDim.y;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int j0 = blockIdx.x * FT_X + 4 * tx;
    const int i = blockIdx.y * LT_ROWS + ty;
    const int pitch = g.pitch;
    const bool live = i < g.rows && j0 < pitch;
    const long long o = (long long)i * pitch + j0;
    T wu[4], wd[4], wl[4], wr[4], dg[4];
    bool hard[4], work = false;
    if (live) {
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
            wu[c] = i > 0 ? w_down(g, i - 1, j, fu[c]) : T(0);
            wd[c] = w_down(g, i, j, fc[c]);
            wl[c] = j > 0 ? w_right(g, i, j - 1, c == 0 ? fl : fc[c - 1]) : T(0);
            wr[c] = w_right(g, i, j, fc[c]);
            dg[c] = wu[c] + wd[c] + wl[c] + wr[c] + soft_w(g, o + c, fc[c]);
            hard[c] = (fc[c] & HG_FLAG_HARD) != 0;
            work = work || !hard[c];
        }
    }
    double acc[KC];
    for (int k0 = 0; k0 < g.K; k0 += KC) {
        V4<T> cur[KC], up[KC], dn[KC], bv[KC];
        T xl[KC], xr[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            cur[q] = up[q] = dn[q] = bv[q] = zero4<T>();
            xl[q] = xr[q] = T(0);
            if (live && work && k0 + q < g.K) {
                const T* xk = x + (k0 + q) * g.plane;
                cur[q] = ld4(xk + o);
                if (i > 0) up[q] = ld4(xk + o - pitch);
                if (i + 1 < g.rows) dn[q] = ld4(xk + o + pitch);
                if (j0 > 0) xl[q] = xk[o - 1];
                if (j0 + 4 < pitch) xr[q] = xk[o + 4];
                if (MODE == 1 && b) bv[q] = ld4(b + (k0 + q) * g.plane + o);
            }
        }
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            acc[q] = 0.0;
            if (live && k0 + q < g.K) {
                V4<T> out = zero4<T>();
                if (work) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const T xL = c == 0 ? xl[q] : cur[q].v[c - 1];
                        const T xR = c == 3 ? xr[q] : cur[q].v[c + 1];
                        T v = dg[c] * cur[q].v[c] - wu[c] * up[q].v[c] - wd[c] * dn[q].v[c] - wl[c] * xL - wr[c] * xR;
                        if (MODE == 1) v = bv[q].v[c] - v;
                        if (hard[c]) v = T(0);
                        out.v[c] = v;
                        if (DOT == 1) acc[q] += (double)cur[q].v[c] * (double)v;
                        if (DOT == 2) acc[q] += (double)v * (double)v;
                    }
                }
                st4(y + (k0 + q) * g.plane + o, out);
            }
        }
        if (DOT != 0) {
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                const double s = block_sum(acc[q], red);
                if (threadIdx.x == 0 && k0 + q < g.K) partial[(long long)(k0 + q) * nblk + blin] = s;
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

}  // namespace hg
