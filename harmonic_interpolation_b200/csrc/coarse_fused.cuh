// coarse_fused.cuh -- fused V-cycle legs of the 9-point (Galerkin) levels.
//
// Same mathematics as the separate kernels of mg.cuh (four-colour Gauss-Seidel, residual,
// restriction; prolongation, reversed four-colour Gauss-Seidel) but ONE kernel per level and leg,
// working tile by tile in shared memory, so a level's planes cross HBM once per leg instead of
// once per colour:
//
//   cdown_kernel   b_l  ->  x_l = S_pre(0, b_l),   b_{l+1} = R (b_l - A_l x_l)
//   cup_kernel     x_l  <-  S_post(x_l + P x_{l+1}, b_l)
//
// Colours: c = 2 (i & 1) + (j & 1), swept 0,1,2,3 on the way down and 3,2,1,0 on the way up.
//
// Geometry.  A block owns an aligned 64 x 32 tile.  One full four-colour sweep reaches 1 row and
// 3 columns far (even rows only look along their own row), the residual one more, the restriction
// needs the residual one pixel above / left of the tile: the REGION is rows [i0-2, i0+34) and
// columns [j0-4, j0+68), 36 x 72 pixels (18 % more rows/columns read than written).  Every stage
// sweeps the whole region; what it computes near the region's edge is incomplete but is never
// consumed by a value that leaves the block.
//
// Shared memory holds only the iterate, colour-split: X[k][c][r][q] is pixel (2r + ci, 2q + cj) of
// the region, so the eight neighbours of a pixel are plain [r or r+-1][q or q+-1] accesses of the
// other three colour arrays and consecutive lanes touch consecutive words.  The right-hand side
// and the nine coefficients of a pixel are needed exactly once per leg -- when that pixel is
// updated (down leg: the residual of colour c uses exactly the coefficients its update skipped) --
// and are read straight from global memory.  On the way down the right-hand side is loaded INTO
// the pixel's own slot (the guess is zero and not-yet-updated neighbours are skipped, not read).
//
// Coefficient compression: at setup every pixel of a level gets a code byte -- 0 inactive,
// 1 regular (its nine coefficients are bit-identical to the level's interior stencil `ref`, which
// depends on the level only), 2 general.  Regular pixels take their coefficients from registers:
// on the fill workload ~90 % of level 1, and the 36 (72) bytes per pixel of coefficient traffic
// all but disappear.  Tiles without an active pixel are skipped (their x stays 0 for ever).
#pragma once
#include "common.cuh"
#include "mg.cuh"

namespace hg {

#define CF_TX 64
#define CF_TY 32
#define CF_HT 2                       // halo rows above the tile (2 more rows below)
#define CF_HL 4                       // halo columns left of the tile (4 more on the right)
#define CF_RW (CF_TX + 2 * CF_HL)     // 72
#define CF_RH (CF_TY + 2 * CF_HT)     // 36
#define CF_NR (CF_RH / 2)             // 18 rows per colour
#define CF_NQ (CF_RW / 2)             // 36 columns per colour
#define CF_SR (CF_NR + 2)             // array rows: one guard row above and below
#define CF_SQ (CF_NQ + 4)             // array columns: two guard columns on each side (keeps pairs 8/16-byte aligned)
#define CF_THREADS 256
#define CF_NGUARD (2 * CF_SQ + 4 * CF_NR)   // guard cells of one array

#define CF_ALL8 0x1EF
#define CF_ROW 0x028                  // left, right
#define CF_COLDIAG 0x1C7              // up, down and the four diagonals

template <typename T, int KC, bool UP>
struct CFSmem {
    T X[KC][4][CF_SR][CF_SQ];
    T SE[UP ? KC : 1][UP ? CF_NR + 1 : 1][UP ? CF_SQ : 1];   // next-coarser correction under the region (19 x 37 used)
    uint8_t C[4][CF_SR][CF_SQ];
};

__device__ __forceinline__ void cf_sts2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void cf_sts2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// ---- setup: code bytes and tile activity of a level ------------------------------------------------
template <typename T>
__global__ void ccode_kernel(Coarse<T> L) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= L.pitch) return;
    const long long o = (long long)i * L.pitch + j;
    unsigned c = 0;
    if (j < L.cols && L.coef[cidx(4, o)] != T(0)) {
        c = 2;
        if (L.has_ref) {
            bool same = true;
#pragma unroll
            for (int n = 0; n < 9; ++n) same = same && (L.coef[cidx(n, o)] == L.ref[n]);
            if (same) c = 1;
        }
    }
    L.code[o] = (uint8_t)c;
}

// any active pixel within the tile grown by `halo`
static __global__ void ctile_activity_kernel(const uint8_t* __restrict__ code, int rows, int cols, int pitch, int halo,
                                      uint8_t* __restrict__ active) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    const int bi = blockIdx.y * CF_TY - halo, bj = blockIdx.x * CF_TX - halo;
    const int w = CF_TX + 2 * halo, h = CF_TY + 2 * halo;
    int mine = 0;
    for (int e = threadIdx.x; e < w * h; e += blockDim.x) {
        const int i = bi + e / w, j = bj + e % w;
        if (i >= 0 && i < rows && j >= 0 && j < cols && code[(long long)i * pitch + j]) mine = 1;
    }
    if (mine) any = 1;
    __syncthreads();
    if (threadIdx.x == 0) active[blockIdx.y * gridDim.x + blockIdx.x] = (uint8_t)any;
}

// ---- pieces -----------------------------------------------------------------------------------------
template <typename T, int KC, bool UP>
__device__ __forceinline__ void cf_zero_guards(CFSmem<T, KC, UP>& S) {
    for (int e = threadIdx.x; e < KC * 4 * CF_NGUARD; e += CF_THREADS) {
        const int arr = e / CF_NGUARD, gq = e - arr * CF_NGUARD;
        T* base = &S.X[0][0][0][0] + arr * (CF_SR * CF_SQ);
        int idx;
        if (gq < CF_SQ) idx = gq;
        else if (gq < 2 * CF_SQ) idx = (CF_SR - 1) * CF_SQ + (gq - CF_SQ);
        else {
            const int h = gq - 2 * CF_SQ, row = 1 + (h >> 2), cc = h & 3;
            idx = row * CF_SQ + (cc < 2 ? cc : CF_SQ - 4 + cc);
        }
        base[idx] = T(0);
    }
}

template <typename T, int KC, bool UP>
__device__ __forceinline__ void cf_load_codes(const Coarse<T>& L, int gi0, int gj0, CFSmem<T, KC, UP>& S) {
    for (int e = threadIdx.x; e < CF_RH * (CF_RW / 4); e += CF_THREADS) {
        const int li = e / (CF_RW / 4), u = e - li * (CF_RW / 4);
        const int i = gi0 + li, j = gj0 + 4 * u;
        unsigned c4 = 0;
        if (i >= 0 && i < L.rows && j >= 0 && j < L.pitch) c4 = *reinterpret_cast<const unsigned*>(L.code + (long long)i * L.pitch + j);
        const int ca = (li & 1) * 2, a = (li >> 1) + 1, t = 2 * u + 2;
        S.C[ca][a][t] = (uint8_t)(c4 & 0xffu);
        S.C[ca + 1][a][t] = (uint8_t)((c4 >> 8) & 0xffu);
        S.C[ca][a][t + 1] = (uint8_t)((c4 >> 16) & 0xffu);
        S.C[ca + 1][a][t + 1] = (uint8_t)(c4 >> 24);
    }
}

// the nine coefficients of a pixel: registers (regular) or the planes (general); only those in MASK (+ centre)
template <typename T, int MASK>
__device__ __forceinline__ void cf_coefs(const Coarse<T>& L, unsigned code, long long o, T (&cf)[9]) {
    if (code == 1u) {
#pragma unroll
        for (int n = 0; n < 9; ++n) cf[n] = L.ref[n];
    } else {
#pragma unroll
        for (int n = 0; n < 9; ++n) cf[n] = ((MASK >> n) & 1) ? L.coef[cidx(n, o)] : T(0);
    }
}

// sum over the neighbours in MASK of  coefficient * value  for pixel (r, q) of colour c, channel slot k
template <typename T, int KC, bool UP, int MASK>
__device__ __forceinline__ T cf_nsum(const CFSmem<T, KC, UP>& S, int k, int c, int r, int q, const T (&cf)[9]) {
    const int ci = c >> 1, cj = c & 1;
    const int a = r + 1, aU = r + ci, aD = r + ci + 1;          // array rows: own, above, below
    const int t = q + 2, tL = q + 1 + cj, tR = q + 2 + cj;      // array columns: own, left, right
    T s = T(0);
    if (MASK & 0x008) s += cf[3] * S.X[k][c ^ 1][a][tL];
    if (MASK & 0x020) s += cf[5] * S.X[k][c ^ 1][a][tR];
    if (MASK & 0x002) s += cf[1] * S.X[k][c ^ 2][aU][t];
    if (MASK & 0x080) s += cf[7] * S.X[k][c ^ 2][aD][t];
    if (MASK & 0x001) s += cf[0] * S.X[k][c ^ 3][aU][tL];
    if (MASK & 0x004) s += cf[2] * S.X[k][c ^ 3][aU][tR];
    if (MASK & 0x040) s += cf[6] * S.X[k][c ^ 3][aD][tL];
    if (MASK & 0x100) s += cf[8] * S.X[k][c ^ 3][aD][tR];
    return s;
}

// Gauss-Seidel update of all region pixels of colour c.  MASK: the neighbours that take part (the
// others are known to hold zero).  RHS_IN_X: the right-hand side sits in the pixel's own slot
// (down leg), otherwise it is read from the planes `b`.
template <typename T, int KC, bool UP, int MASK, bool RHS_IN_X>
__device__ __forceinline__ void cf_sweep(const Coarse<T>& L, CFSmem<T, KC, UP>& S, int c, int gi0, int gj0, int k0, int kn,
                                         const T* __restrict__ b) {
    const int ci = c >> 1, cj = c & 1;
    for (int e = threadIdx.x; e < CF_NR * CF_NQ; e += CF_THREADS) {
        const int r = e / CF_NQ, q = e - r * CF_NQ;
        const unsigned code = S.C[c][r + 1][q + 2];
        if (code == 0u) {
            if (RHS_IN_X) {
#pragma unroll
                for (int k = 0; k < KC; ++k) S.X[k][c][r + 1][q + 2] = T(0);
            }
            continue;
        }
        const long long o = (long long)(gi0 + 2 * r + ci) * L.pitch + (gj0 + 2 * q + cj);
        T rhs[KC];
        if (!RHS_IN_X) {
#pragma unroll
            for (int k = 0; k < KC; ++k) rhs[k] = k < kn ? b[(k0 + k) * L.plane + o] : T(0);
        }
        T cf[9];
        cf_coefs<T, MASK | 0x010>(L, code, o, cf);
        const T inv = T(1) / cf[4];
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const T v = (RHS_IN_X ? S.X[k][c][r + 1][q + 2] : rhs[k]) - cf_nsum<T, KC, UP, MASK>(S, k, c, r, q, cf);
            S.X[k][c][r + 1][q + 2] = v * inv;
        }
    }
}

// residual of pixel (r, q) of colour c after the downward sweep: minus the couplings to the colours
// updated after it (MASK); zero at inactive pixels
template <typename T, int KC, int MASK>
__device__ __forceinline__ void cf_resid(const Coarse<T>& L, const CFSmem<T, KC, false>& S, int c, int r, int q, int gi0, int gj0,
                                         T (&out)[KC]) {
#pragma unroll
    for (int k = 0; k < KC; ++k) out[k] = T(0);
    const unsigned code = S.C[c][r + 1][q + 2];
    if (code == 0u) return;
    const long long o = (long long)(gi0 + 2 * r + (c >> 1)) * L.pitch + (gj0 + 2 * q + (c & 1));
    T cf[9];
    cf_coefs<T, MASK>(L, code, o, cf);
#pragma unroll
    for (int k = 0; k < KC; ++k) out[k] = -cf_nsum<T, KC, false, MASK>(S, k, c, r, q, cf);
}

// the tile of X (re-interleaving the colours) -> the planes `x`
template <typename T, int KC, bool UP>
__device__ __forceinline__ void cf_store_tile(const Coarse<T>& L, const CFSmem<T, KC, UP>& S, int ti0, int tj0, int k0, int kn,
                                              T* __restrict__ x) {
    for (int e = threadIdx.x; e < CF_TY * (CF_TX / 4); e += CF_THREADS) {
        const int ii = e >> 4, u = e & 15;
        const int i = ti0 + ii, j = tj0 + 4 * u;
        if (i >= L.rows || j >= L.pitch) continue;
        const int li = ii + CF_HT;
        const int ca = (li & 1) * 2, a = (li >> 1) + 1, t = 2 * u + (CF_HL >> 1) + 2;
        // a quad without an active pixel is zero in memory already and is never written
        if (!(S.C[ca][a][t] | S.C[ca + 1][a][t] | S.C[ca][a][t + 1] | S.C[ca + 1][a][t + 1])) continue;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (k >= kn) break;
            V4<T> out;
            out.v[0] = S.X[k][ca][a][t]; out.v[1] = S.X[k][ca + 1][a][t];
            out.v[2] = S.X[k][ca][a][t + 1]; out.v[3] = S.X[k][ca + 1][a][t + 1];
            st4(x + (k0 + k) * L.plane + (long long)i * L.pitch + j, out);
        }
    }
}

// ---- downward leg --------------------------------------------------------------------------------------
template <typename T, int KC>
__global__ void __launch_bounds__(CF_THREADS)
cdown_kernel(Coarse<T> L, Coarse<T> Lc, const int* __restrict__ done) {
    if (done && *done) return;
    if (!L.act_down[blockIdx.y * gridDim.x + blockIdx.x]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CFSmem<T, KC, false>& S = *reinterpret_cast<CFSmem<T, KC, false>*>(smem_raw);
    const int ti0 = blockIdx.y * CF_TY, tj0 = blockIdx.x * CF_TX;
    const int gi0 = ti0 - CF_HT, gj0 = tj0 - CF_HL;
    const int k0 = blockIdx.z * KC, kn = min(KC, L.K - k0);

    cf_zero_guards(S);
    // codes, and the right-hand side -> the pixels' own slots (quads without an active pixel are not read: b is 0 there)
    for (int e = threadIdx.x; e < CF_RH * (CF_RW / 4); e += CF_THREADS) {
        const int li = e / (CF_RW / 4), u = e - li * (CF_RW / 4);
        const int i = gi0 + li, j = gj0 + 4 * u;
        const bool in = i >= 0 && i < L.rows && j >= 0 && j < L.pitch;
        const int ca = (li & 1) * 2, a = (li >> 1) + 1, t = 2 * u + 2;
        unsigned c4 = 0;
        if (in) c4 = *reinterpret_cast<const unsigned*>(L.code + (long long)i * L.pitch + j);
        S.C[ca][a][t] = (uint8_t)(c4 & 0xffu);
        S.C[ca + 1][a][t] = (uint8_t)((c4 >> 8) & 0xffu);
        S.C[ca][a][t + 1] = (uint8_t)((c4 >> 16) & 0xffu);
        S.C[ca + 1][a][t + 1] = (uint8_t)(c4 >> 24);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            V4<T> v = zero4<T>();
            if (c4 != 0u && k < kn) v = ld4(L.b + (k0 + k) * L.plane + (long long)i * L.pitch + j);
            cf_sts2(&S.X[k][ca][a][t], v.v[0], v.v[2]);
            cf_sts2(&S.X[k][ca + 1][a][t], v.v[1], v.v[3]);
        }
    }
    __syncthreads();
    cf_sweep<T, KC, false, 0, true>(L, S, 0, gi0, gj0, k0, kn, nullptr);
    __syncthreads();
    cf_sweep<T, KC, false, CF_ROW, true>(L, S, 1, gi0, gj0, k0, kn, nullptr);
    __syncthreads();
    cf_sweep<T, KC, false, CF_COLDIAG, true>(L, S, 2, gi0, gj0, k0, kn, nullptr);
    __syncthreads();
    cf_sweep<T, KC, false, CF_ALL8, true>(L, S, 3, gi0, gj0, k0, kn, nullptr);
    __syncthreads();
    cf_store_tile(L, S, ti0, tj0, k0, kn, L.x);

    // residual + restriction to the 32 x 16 next-coarser pixels under the tile.  Next-coarser (I, J)
    // sits on pixel (2I, 2J) of this level, colour 0; its diagonal neighbours are colour 3, whose
    // residual vanishes, so five residuals make up the restricted value.
    for (int e = threadIdx.x; e < (CF_TY / 2) * (CF_TX / 2); e += CF_THREADS) {
        const int cI = e >> 5, cJ = e & 31;
        const int I = (ti0 >> 1) + cI, J = (tj0 >> 1) + cJ;
        if (I >= Lc.rows || J >= Lc.pitch) continue;
        const int r = (CF_HT >> 1) + cI, q = (CF_HL >> 1) + cJ;      // colour-0 position of (2I, 2J) in the region
        T r0[KC], rl[KC], rr[KC], ru[KC], rd[KC];
        cf_resid<T, KC, CF_ALL8>(L, S, 0, r, q, gi0, gj0, r0);
        cf_resid<T, KC, CF_COLDIAG>(L, S, 1, r, q - 1, gi0, gj0, rl);
        cf_resid<T, KC, CF_COLDIAG>(L, S, 1, r, q, gi0, gj0, rr);
        cf_resid<T, KC, CF_ROW>(L, S, 2, r - 1, q, gi0, gj0, ru);
        cf_resid<T, KC, CF_ROW>(L, S, 2, r, q, gi0, gj0, rd);
        const T wim = (T)w1d(-1, I, L.rows, Lc.rows), wip = (T)w1d(1, I, L.rows, Lc.rows);
        const T wjm = (T)w1d(-1, J, L.cols, Lc.cols), wjp = (T)w1d(1, J, L.cols, Lc.cols);
        const long long oc = (long long)I * Lc.pitch + J;
        if (!Lc.code[oc]) continue;       // inactive next-coarser pixels keep the zeros they were created with
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (k >= kn) break;
            Lc.b[(k0 + k) * Lc.plane + oc] = r0[k] + (wjm * rl[k] + wjp * rr[k]) + (wim * ru[k] + wip * rd[k]);
        }
    }
}

// ---- upward leg -----------------------------------------------------------------------------------------
template <typename T, int KC>
__global__ void __launch_bounds__(CF_THREADS)
cup_kernel(Coarse<T> L, Coarse<T> Lc, const int* __restrict__ done) {
    if (done && *done) return;
    if (!L.act_up[blockIdx.y * gridDim.x + blockIdx.x]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CFSmem<T, KC, true>& S = *reinterpret_cast<CFSmem<T, KC, true>*>(smem_raw);
    const int ti0 = blockIdx.y * CF_TY, tj0 = blockIdx.x * CF_TX;
    const int gi0 = ti0 - CF_HT, gj0 = tj0 - CF_HL;       // both even
    const int k0 = blockIdx.z * KC, kn = min(KC, L.K - k0);

    cf_zero_guards(S);
    // next-coarser correction under the region: rows gi0/2 .. +18, columns gj0/2 .. +36, replicated at
    // the far edge of the grid (constant extrapolation of P), clamped at the near edge
    for (int e = threadIdx.x; e < (CF_NR + 1) * (CF_NQ + 1); e += CF_THREADS) {
        const int a = e / (CF_NQ + 1), bq = e - a * (CF_NQ + 1);
        const int I = min(max((gi0 >> 1) + a, 0), Lc.rows - 1), J = min(max((gj0 >> 1) + bq, 0), Lc.cols - 1);
#pragma unroll
        for (int k = 0; k < KC; ++k) S.SE[k][a][bq] = k < kn ? Lc.x[(k0 + k) * Lc.plane + (long long)I * Lc.pitch + J] : T(0);
    }
    __syncthreads();
    // x + P e -> colour-split arrays (active pixels only; the codes ride along)
    for (int e = threadIdx.x; e < CF_RH * (CF_RW / 4); e += CF_THREADS) {
        const int li = e / (CF_RW / 4), u = e - li * (CF_RW / 4);
        const int i = gi0 + li, j = gj0 + 4 * u;
        const bool in = i >= 0 && i < L.rows && j >= 0 && j < L.pitch;
        unsigned c4 = 0;
        if (in) c4 = *reinterpret_cast<const unsigned*>(L.code + (long long)i * L.pitch + j);
        const int ci = li & 1, ca = ci * 2, r = li >> 1, a = r + 1, t = 2 * u + 2;
        S.C[ca][a][t] = (uint8_t)(c4 & 0xffu);
        S.C[ca + 1][a][t] = (uint8_t)((c4 >> 8) & 0xffu);
        S.C[ca][a][t + 1] = (uint8_t)((c4 >> 16) & 0xffu);
        S.C[ca + 1][a][t + 1] = (uint8_t)(c4 >> 24);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            V4<T> v = zero4<T>();
            if (c4 != 0u && k < kn) {      // quads without an active pixel hold zeros and stay so
                v = ld4(L.x + (k0 + k) * L.plane + (long long)i * L.pitch + j);
                // pixels lj = 4u + p: next-coarser column 2u + (p >> 1) (+1 for odd lj)
                const T e00 = S.SE[k][r][2 * u], e01 = S.SE[k][r][2 * u + 1], e02 = S.SE[k][r][2 * u + 2];
                T p0, p1, p2, p3;
                if (ci) {
                    const T e10 = S.SE[k][r + 1][2 * u], e11 = S.SE[k][r + 1][2 * u + 1], e12 = S.SE[k][r + 1][2 * u + 2];
                    p0 = T(0.5) * (e00 + e10);
                    p1 = T(0.25) * ((e00 + e01) + (e10 + e11));
                    p2 = T(0.5) * (e01 + e11);
                    p3 = T(0.25) * ((e01 + e02) + (e11 + e12));
                } else {
                    p0 = e00;
                    p1 = T(0.5) * (e00 + e01);
                    p2 = e01;
                    p3 = T(0.5) * (e01 + e02);
                }
                if (c4 & 0x000000ffu) v.v[0] += p0;
                if (c4 & 0x0000ff00u) v.v[1] += p1;
                if (c4 & 0x00ff0000u) v.v[2] += p2;
                if (c4 & 0xff000000u) v.v[3] += p3;
            }
            cf_sts2(&S.X[k][ca][a][t], v.v[0], v.v[2]);
            cf_sts2(&S.X[k][ca + 1][a][t], v.v[1], v.v[3]);
        }
    }
    __syncthreads();
    cf_sweep<T, KC, true, CF_ALL8, false>(L, S, 3, gi0, gj0, k0, kn, L.b);
    __syncthreads();
    cf_sweep<T, KC, true, CF_ALL8, false>(L, S, 2, gi0, gj0, k0, kn, L.b);
    __syncthreads();
    cf_sweep<T, KC, true, CF_ALL8, false>(L, S, 1, gi0, gj0, k0, kn, L.b);
    __syncthreads();
    cf_sweep<T, KC, true, CF_ALL8, false>(L, S, 0, gi0, gj0, k0, kn, L.b);
    __syncthreads();
    // (to the second buffer: neighbouring tiles read this tile's pre-smoothed x as their halo)
    cf_store_tile(L, S, ti0, tj0, k0, kn, L.x2);
}

}  // namespace hg
