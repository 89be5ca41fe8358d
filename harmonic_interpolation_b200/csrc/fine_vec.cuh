// fine_vec.cuh -- quad-vectorised fused V-cycle legs of level 0 (no gradient-penalty edges).
//
// Same algorithm, region logic and colour-split layout as the scalar legs in fine.cuh (read that
// header first), but every shared-memory and global access moves FOUR values: a slot owns four
// consecutive same-colour pixels of a row, the rows above / below are one 16-byte load each, the
// left / right neighbours one 16-byte load plus one scalar, and the tile of r (and of the code /
// flag bytes) is fetched with aligned float4 / uchar4 loads.  The scalar version executed ~100
// instructions per pixel and stage and was instruction-bound (ncu: 4.1e9 warp instructions, 73 %
// issue-active, 13 % of DRAM peak); this one needs ~12.
//
// Region: the aligned 64 x 32 tile plus a halo of 4 -> 72 x 40 pixels, 36 red/black pairs per row,
// i.e. 9 quads.  Arrays are [42][44]: data at rows 1..40, columns 4..39 (16-byte aligned quads),
// guards around.  Block = 384 threads: 360 quad slots per stage.
#pragma once
#include "fine.cuh"

namespace hg {

#define VR_H 4
#define VR_W (FT_X + 2 * VR_H)        // 72
#define VR_RH (FT_Y + 2 * VR_H)       // 40
#define VR_PW (VR_W / 2)              // 36
#define VR_Q (VR_PW / 4)              // 9 quads per row
#define VR_SW (VR_PW + 8)             // 44: 4 guard columns on each side keep quads 16-byte aligned
#define VR_SH (VR_RH + 2)
#define VR_THREADS 384
#define VR_CW 40                      // row stride of the coarse-correction tile

template <typename T>
struct VRegionSmem {
    T rr[VR_SH][VR_SW], rb[VR_SH][VR_SW];
    T zr[VR_SH][VR_SW], zb[VR_SH][VR_SW];
    T dr[VR_SH][VR_SW], db[VR_SH][VR_SW];
    uint8_t cr[VR_SH][VR_SW], cb[VR_SH][VR_SW];
    T se[VR_RH / 2 + 2][VR_CW];       // coarse correction under the region (upward leg)
};

template <typename T> struct Q4 { T v[4]; };

template <typename T> __device__ __forceinline__ Q4<T> lds4(const T* p);
template <> __device__ __forceinline__ Q4<float> lds4<float>(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    Q4<float> q; q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w; return q;
}
template <> __device__ __forceinline__ Q4<double> lds4<double>(const double* p) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    Q4<double> q; q.v[0] = a.x; q.v[1] = a.y; q.v[2] = b.x; q.v[3] = b.y; return q;
}
__device__ __forceinline__ void sts4(float* p, const Q4<float>& q) { *reinterpret_cast<float4*>(p) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]); }
__device__ __forceinline__ void sts4(double* p, const Q4<double>& q) {
    *reinterpret_cast<double2*>(p) = make_double2(q.v[0], q.v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(q.v[2], q.v[3]);
}
// four values from an address that is only 8-byte aligned (fp32) -- two 64-bit loads
__device__ __forceinline__ Q4<float> lds4u(const float* p) {
    const float2 a = *reinterpret_cast<const float2*>(p), b = *reinterpret_cast<const float2*>(p + 2);
    Q4<float> q; q.v[0] = a.x; q.v[1] = a.y; q.v[2] = b.x; q.v[3] = b.y; return q;
}
__device__ __forceinline__ Q4<double> lds4u(const double* p) { return lds4<double>(p); }
__device__ __forceinline__ void sts2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void sts2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// inverse diagonal of a free pixel from its code byte (and soft weight)
template <typename T>
__device__ __forceinline__ T dinv_of(const Grid<T>& g, unsigned c, unsigned f, long long o) {
    if (f & HG_FLAG_HARD) return T(0);
    if (c == 0xAAu && !(f & HG_FLAG_SOFT)) return T(1);
    T d = T(0.125) * T((c & 3u) + ((c >> 2) & 3u) + ((c >> 4) & 3u) + ((c >> 6) & 3u));
    if (f & HG_FLAG_SOFT) d += soft_w(g, o, f);
    return d > T(0) ? T(1) / d : T(0);
}

// operator cache of the region: aligned uchar4 loads of the code and flag planes
template <typename T>
__device__ __forceinline__ void vregion_build_op(const Grid<T>& g, const uint8_t* __restrict__ code, int gi0, int gj0,
                                                 VRegionSmem<T>& S) {
    for (int e = threadIdx.x; e < VR_RH * (VR_W / 4); e += VR_THREADS) {
        const int li = e / (VR_W / 4), u = e - li * (VR_W / 4);
        const int i = gi0 + li, j = gj0 + 4 * u;
        unsigned c4 = 0, f4 = 0x01010101u;
        const long long o = (long long)i * g.pitch + j;
        if (i >= 0 && i < g.rows && j >= 0 && j < g.pitch) {
            c4 = *reinterpret_cast<const unsigned*>(code + o);
            f4 = *reinterpret_cast<const unsigned*>(g.flags + o);
        }
        T d[4];
        unsigned cc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            cc[q] = (c4 >> (8 * q)) & 0xffu;
            d[q] = dinv_of(g, cc[q], (f4 >> (8 * q)) & 0xffu, o + q);
        }
        const int a = li + 1, t = 2 * u + 4;
        // pixels 4u .. 4u+3: colours alternate, starting with red iff (li + 4u) is even, i.e. li even
        if (li & 1) {
            sts2(&S.db[a][t], d[0], d[2]); sts2(&S.dr[a][t], d[1], d[3]);
            S.cb[a][t] = (uint8_t)cc[0]; S.cb[a][t + 1] = (uint8_t)cc[2];
            S.cr[a][t] = (uint8_t)cc[1]; S.cr[a][t + 1] = (uint8_t)cc[3];
        } else {
            sts2(&S.dr[a][t], d[0], d[2]); sts2(&S.db[a][t], d[1], d[3]);
            S.cr[a][t] = (uint8_t)cc[0]; S.cr[a][t + 1] = (uint8_t)cc[2];
            S.cb[a][t] = (uint8_t)cc[1]; S.cb[a][t + 1] = (uint8_t)cc[3];
        }
    }
}

// Load geometry of a thread, computed once per tile: the (up to) two float4 it fetches per channel.
struct VLoadSlot {
    long long off[2];     // element offset in the plane, -1 = outside the grid (zeros)
    int first[2];         // smem element index (row * VR_SW + col) of the quad's first-colour pair
    int odd[2];           // row parity: which colour comes first
    int n;
};
__device__ __forceinline__ VLoadSlot vregion_load_slots(int rows, int pitch, int gi0, int gj0) {
    VLoadSlot L;
    L.n = 0;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int e = threadIdx.x + it * VR_THREADS;
        L.off[it] = -1; L.first[it] = 0; L.odd[it] = 0;
        if (e < VR_RH * (VR_W / 4)) {
            const int li = e / (VR_W / 4), u = e - li * (VR_W / 4);
            const int i = gi0 + li, j = gj0 + 4 * u;
            if (i >= 0 && i < rows && j >= 0 && j < pitch) L.off[it] = (long long)i * pitch + j;
            L.first[it] = (li + 1) * VR_SW + 2 * u + 4;
            L.odd[it] = li & 1;
            L.n = it + 1;
        }
    }
    return L;
}
// one channel of a value plane -> colour-split arrays, aligned float4 loads
template <typename T>
__device__ __forceinline__ void vregion_load(const T* __restrict__ v, const VLoadSlot& L, T* red, T* blk) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        if (it < L.n) {
            V4<T> x = zero4<T>();
            if (L.off[it] >= 0) x = ld4(v + L.off[it]);
            T* p0 = (L.odd[it] ? blk : red) + L.first[it];
            T* p1 = (L.odd[it] ? red : blk) + L.first[it];
            sts2(p0, x.v[0], x.v[2]);
            sts2(p1, x.v[1], x.v[3]);
        }
    }
}

// weighted neighbour sums of the quad (row a, first index c) of colour `col` (0 red, 1 black)
template <typename T>
__device__ __forceinline__ Q4<T> vregion_nsum(const T (*other)[VR_SW], const uint8_t (*codes)[VR_SW], int col, int a, int c) {
    const int s = (a - 1) & 1;
    const int d = col == 0 ? s : 1 - s;           // left index = c - 1 + d, right index = c + d
    const Q4<T> up = lds4<T>(&other[a - 1][c]), dn = lds4<T>(&other[a + 1][c]), mid = lds4<T>(&other[a][c]);
    const T ex = d ? other[a][c + 4] : other[a][c - 1];
    T lf[4], rt[4];
    if (d) { lf[0] = mid.v[0]; lf[1] = mid.v[1]; lf[2] = mid.v[2]; lf[3] = mid.v[3]; rt[0] = mid.v[1]; rt[1] = mid.v[2]; rt[2] = mid.v[3]; rt[3] = ex; }
    else   { lf[0] = ex; lf[1] = mid.v[0]; lf[2] = mid.v[1]; lf[3] = mid.v[2]; rt[0] = mid.v[0]; rt[1] = mid.v[1]; rt[2] = mid.v[2]; rt[3] = mid.v[3]; }
    const unsigned c4 = *reinterpret_cast<const unsigned*>(&codes[a][c]);
    Q4<T> out;
    // fast path: every pixel of the quad is regular (code 0xAA) or dead (code 0x00: its value is
    // forced to 0 by the caller through dinv == 0, whatever is computed here)
    const unsigned hi = c4 & 0xAAAAAAAAu;
    if ((c4 & 0x55555555u) == 0u && hi == ((hi >> 7) & 0x01010101u) * 0xAAu) {
#pragma unroll
        for (int q = 0; q < 4; ++q) out.v[q] = T(0.25) * ((up.v[q] + dn.v[q]) + (lf[q] + rt[q]));
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            T wu, wd, wl, wr;
            decode((c4 >> (8 * q)) & 0xffu, wu, wd, wl, wr);
            out.v[q] = wu * up.v[q] + wd * dn.v[q] + wl * lf[q] + wr * rt[q];
        }
    }
    return out;
}

// dst = (rhs + nsum) * dinv   (Gauss-Seidel update of one colour), or dst = nsum masked by dinv != 0
template <typename T, bool RESID>
__device__ __forceinline__ void vregion_sweep(VRegionSmem<T>& S, int col, T (*dst)[VR_SW], const T (*rhs)[VR_SW],
                                              const T (*other)[VR_SW]) {
    const T (*dinv)[VR_SW] = col == 0 ? S.dr : S.db;
    const uint8_t (*codes)[VR_SW] = col == 0 ? S.cr : S.cb;
    for (int sl = threadIdx.x; sl < VR_Q * VR_RH; sl += VR_THREADS) {
        const int li = sl / VR_Q, a = li + 1, c = 4 * (sl - li * VR_Q) + 4;
        const Q4<T> di = lds4<T>(&dinv[a][c]);
        Q4<T> out;
        if (di.v[0] == T(0) && di.v[1] == T(0) && di.v[2] == T(0) && di.v[3] == T(0)) {
            out.v[0] = out.v[1] = out.v[2] = out.v[3] = T(0);
        } else {
            const Q4<T> ns = vregion_nsum<T>(other, codes, col, a, c);
            if (RESID) {
#pragma unroll
                for (int q = 0; q < 4; ++q) out.v[q] = di.v[q] != T(0) ? ns.v[q] : T(0);
            } else {
                const Q4<T> rh = lds4<T>(&rhs[a][c]);
#pragma unroll
                for (int q = 0; q < 4; ++q) out.v[q] = (rh.v[q] + ns.v[q]) * di.v[q];
            }
        }
        sts4(&dst[a][c], out);
    }
}

// ---- downward leg ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(VR_THREADS)
mg_down0v_kernel(Grid<T> g, const uint8_t* __restrict__ code, const T* __restrict__ r, Coarse<T> cl,
                 const uint8_t* __restrict__ active, const int* __restrict__ done) {
    if (done && *done) return;
    if (!active[blockIdx.y * gridDim.x + blockIdx.x]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VRegionSmem<T>& S = *reinterpret_cast<VRegionSmem<T>*>(smem_raw);
    const int gi0 = blockIdx.y * FT_Y - VR_H, gj0 = blockIdx.x * FT_X - VR_H;
    vregion_build_op<T>(g, code, gi0, gj0, S);
    const VLoadSlot LS = vregion_load_slots(g.rows, g.pitch, gi0, gj0);

    for (int k = 0; k < g.K; ++k) {
        __syncthreads();
        vregion_load<T>(r + k * g.plane, LS, &S.rr[0][0], &S.rb[0][0]);
        __syncthreads();
        // red: z = r / d
        for (int sl = threadIdx.x; sl < VR_Q * VR_RH; sl += VR_THREADS) {
            const int li = sl / VR_Q, a = li + 1, c = 4 * (sl - li * VR_Q) + 4;
            const Q4<T> rh = lds4<T>(&S.rr[a][c]), di = lds4<T>(&S.dr[a][c]);
            Q4<T> z;
#pragma unroll
            for (int q = 0; q < 4; ++q) z.v[q] = rh.v[q] * di.v[q];
            sts4(&S.zr[a][c], z);
        }
        __syncthreads();
        vregion_sweep<T, false>(S, 1, S.zb, S.rb, S.zr);      // black: z = (r + sum w z_red) / d
        __syncthreads();
        vregion_sweep<T, true>(S, 0, S.rr, nullptr, S.zb);    // residual on red pixels -> rr
        __syncthreads();
        // restriction: coarse (ci, cj) of the tile <-> region pixel (4 + 2 ci, 4 + 2 cj): centre + four diagonals
        for (int e = threadIdx.x; e < (FT_Y / 2) * (FT_X / 8); e += VR_THREADS) {
            const int ci = e >> 3, cq = e & 7;
            const int I = blockIdx.y * (FT_Y / 2) + ci, J0 = blockIdx.x * (FT_X / 2) + 4 * cq;
            if (I >= cl.rows || J0 >= cl.pitch) continue;
            const int a = VR_H + 2 * ci + 1;          // row a is even-li: red pixels at even lj, index lj / 2
            const int c = (VR_H >> 1) + 4 * cq + 4;   // index of lj = 4 + 2 cj for cj = 4 cq
            const Q4<T> ce = lds4u(&S.rr[a][c]);              // c = 4 cq + 6: 8-byte aligned only
            // diagonal rows a -+ 1 (odd li: red pixels at odd lj): lj - 1 -> index c - 1, lj + 1 -> index c
            const Q4<T> um = lds4u(&S.rr[a - 1][c]), dm = lds4u(&S.rr[a + 1][c]);
            const T ul = S.rr[a - 1][c - 1], dl = S.rr[a + 1][c - 1];
            const T wim = (T)w1d(-1, I, g.rows, cl.rows), wip = (T)w1d(1, I, g.rows, cl.rows);
            V4<T> out;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int J = J0 + q;
                const T wjm = (T)w1d(-1, J, g.cols, cl.cols), wjp = (T)w1d(1, J, g.cols, cl.cols);
                const T upl = q == 0 ? ul : um.v[q - 1], dnl = q == 0 ? dl : dm.v[q - 1];
                out.v[q] = J < cl.cols ? ce.v[q] + wim * (wjm * upl + wjp * um.v[q]) + wip * (wjm * dnl + wjp * dm.v[q]) : T(0);
            }
            const long long oc = k * cl.plane + (long long)I * cl.pitch + J0;
            st4(cl.b + oc, out);
            st4(cl.x + oc, zero4<T>());
        }
    }
}

// ---- upward leg --------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(VR_THREADS)
mg_up0v_kernel(Grid<T> g, const uint8_t* __restrict__ code, const T* __restrict__ r, T* __restrict__ z, Coarse<T> cl,
               int have_coarse, const uint8_t* __restrict__ active, double* __restrict__ part_rz, const int* __restrict__ done) {
    if (done && *done) return;
    const int blin = blockIdx.y * gridDim.x + blockIdx.x;
    if (!active[blin]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VRegionSmem<T>& S = *reinterpret_cast<VRegionSmem<T>*>(smem_raw);
    __shared__ double red[32];
    const int gi0 = blockIdx.y * FT_Y - VR_H, gj0 = blockIdx.x * FT_X - VR_H;     // both even (multiples of 4)
    const int ntiles = gridDim.x * gridDim.y;
    vregion_build_op<T>(g, code, gi0, gj0, S);
    const VLoadSlot LS = vregion_load_slots(g.rows, g.pitch, gi0, gj0);

    for (int k = 0; k < g.K; ++k) {
        __syncthreads();
        vregion_load<T>(r + k * g.plane, LS, &S.rr[0][0], &S.rb[0][0]);
        if (have_coarse) {
            // coarse correction under the region: rows gi0/2 .. gi0/2 + 20, columns gj0/2 .. gj0/2 + 36 (+ guards),
            // replicated at the far edge of the grid (constant extrapolation of P), clamped at the near edge
            const T* xc = cl.x + k * cl.plane;
            for (int e = threadIdx.x; e < (VR_RH / 2 + 2) * VR_CW; e += VR_THREADS) {
                const int a = e / VR_CW, b = e - a * VR_CW;
                const int I = min(max(gi0 / 2 + a, 0), cl.rows - 1), J = min(max(gj0 / 2 + b, 0), cl.cols - 1);
                S.se[a][b] = xc[(long long)I * cl.pitch + J];
            }
        }
        __syncthreads();
        // red: z = r / d + (P e); red pixels are (even, even) [injection] or (odd, odd) [mean of 4]
        for (int sl = threadIdx.x; sl < VR_Q * VR_RH; sl += VR_THREADS) {
            const int li = sl / VR_Q, t0 = 4 * (sl - li * VR_Q), a = li + 1, c = t0 + 4;
            const Q4<T> rh = lds4<T>(&S.rr[a][c]), di = lds4<T>(&S.dr[a][c]);
            Q4<T> zz;
#pragma unroll
            for (int q = 0; q < 4; ++q) zz.v[q] = rh.v[q] * di.v[q];
            if (have_coarse) {
                const int ci = li >> 1;
                const Q4<T> e0 = lds4<T>(&S.se[ci][t0]);
                if (li & 1) {
                    const Q4<T> e1 = lds4<T>(&S.se[ci + 1][t0]);
                    const T x0 = S.se[ci][t0 + 4], x1 = S.se[ci + 1][t0 + 4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const T r0 = q == 3 ? x0 : e0.v[q + 1], r1 = q == 3 ? x1 : e1.v[q + 1];
                        if (di.v[q] != T(0)) zz.v[q] += T(0.25) * ((e0.v[q] + r0) + (e1.v[q] + r1));
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (di.v[q] != T(0)) zz.v[q] += e0.v[q];
                }
            }
            sts4(&S.zr[a][c], zz);
        }
        __syncthreads();
        vregion_sweep<T, false>(S, 1, S.zb, S.rb, S.zr);      // black
        __syncthreads();
        vregion_sweep<T, false>(S, 0, S.zr, S.rr, S.zb);      // red again
        __syncthreads();
        // write the 64 x 32 tile as float4 rows (re-interleaving the colours), accumulate r.z
        double acc = 0.0;
        T* zk = z + k * g.plane;
        for (int e = threadIdx.x; e < FT_Y * (FT_X / 4); e += VR_THREADS) {
            const int ii = e >> 4, u = e & 15;
            const int i = blockIdx.y * FT_Y + ii, j = blockIdx.x * FT_X + 4 * u;
            if (i >= g.rows || j >= g.pitch) continue;
            const int li = ii + VR_H, a = li + 1, t = 2 * u + (VR_H >> 1) + 4;     // lj = 4 u + 4 -> index lj / 2
            const T* pr = &S.zr[a][t]; const T* pb = &S.zb[a][t];
            const T* qr = &S.rr[a][t]; const T* qb = &S.rb[a][t];
            V4<T> out;
            T rv[4];
            if (li & 1) { out.v[0] = pb[0]; out.v[1] = pr[0]; out.v[2] = pb[1]; out.v[3] = pr[1]; rv[0] = qb[0]; rv[1] = qr[0]; rv[2] = qb[1]; rv[3] = qr[1]; }
            else        { out.v[0] = pr[0]; out.v[1] = pb[0]; out.v[2] = pr[1]; out.v[3] = pb[1]; rv[0] = qr[0]; rv[1] = qb[0]; rv[2] = qr[1]; rv[3] = qb[1]; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (j + q >= g.cols) out.v[q] = T(0);
                acc += (double)out.v[q] * (double)rv[q];
            }
            st4(zk + (long long)i * g.pitch + j, out);
        }
        const double sdot = block_sum(acc, red);
        if (threadIdx.x == 0) part_rz[(long long)k * ntiles + blin] = sdot;
    }
}

}  // namespace hg
