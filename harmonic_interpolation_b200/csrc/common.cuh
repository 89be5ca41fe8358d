// common.cuh -- shared device helpers of libhgrid (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hgrid.h"

#define HG_MAXK 16          // channels solved together; the Python layer chunks beyond this
#define HG_FULL 0xffffffffu
#define HG_GATHER_LEVEL 5    // from this level down the hierarchy is replicated on every rank

namespace hg {

// One level of the grid hierarchy as the kernels see it.  Planes are pitch-strided,
// `pitch` is a multiple of 32 elements so every row starts on a 128-byte boundary
// (fp32) and float4 / uchar4 accesses are aligned.  Padded columns carry HG_FLAG_HARD
// and zeros, so elementwise kernels may ignore `cols`.
template <typename T>
struct Grid {
    int rows, cols, pitch;
    long long plane;       // rows * pitch, distance between channel planes
    int K;
    const uint8_t* flags;  // one byte per pixel
    const T* soft;         // per-pixel soft weights or nullptr (then soft_scalar where HG_FLAG_SOFT)
    T soft_scalar;
    T w_g;                 // weight of HG_FLAG_PEN_* edges
    int uniform;           // 1: every edge weighs 1/4 (poisson.py's G.T*G / 4); 0: 1/8 along the grid border (recovery.py:263-299)
};

template <typename T> struct V4 { T v[4]; };

__device__ __forceinline__ V4<float> ld4(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    V4<float> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
__device__ __forceinline__ V4<double> ld4(const double* p) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    V4<double> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; return r;
}
__device__ __forceinline__ void st4(float* p, const V4<float>& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
__device__ __forceinline__ void st4(double* p, const V4<double>& r) {
    *reinterpret_cast<double2*>(p) = make_double2(r.v[0], r.v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(r.v[2], r.v[3]);
}
template <typename T> __device__ __forceinline__ V4<T> zero4() {
    V4<T> r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = T(0); return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HG_FULL, v, o);
    return v;
}

// Sum over the block in a fixed order; result valid in thread 0.  `red` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (tid < 32) {
        s = tid < nw ? red[tid] : 0.0;
        s = warp_sum(s);
    }
    return s;
}

// N sums at once (one pair of barriers instead of N); results valid in thread 0.  Blocks of at most
// 8 warps; `red` holds >= 8 N doubles.
template <int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double* red) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = warp_sum(v[q]);
    __syncthreads();
    if ((tid & 31) == 0) {
#pragma unroll
        for (int q = 0; q < N; ++q) red[q * 8 + (tid >> 5)] = v[q];
    }
    __syncthreads();
    if (tid < 32) {
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = warp_sum(tid < nw ? red[q * 8 + tid] : 0.0);
    }
}

// ---- symmetric Laplacian edge weights (recovery.py:263-299, 309-319) ------------------
// weight of the edge (i,j)-(i+1,j), given the flag byte of (i,j)
template <typename T>
__device__ __forceinline__ T w_down(const Grid<T>& g, int i, int j, unsigned f) {
    T w = T(0);
    if (i + 1 < g.rows) {
        if (!(f & HG_FLAG_CUT_DOWN)) w = (!g.uniform && g.cols > 1 && (j == 0 || j == g.cols - 1)) ? T(0.125) : T(0.25);
        if (f & HG_FLAG_PEN_DOWN) w += g.w_g;
    }
    return w;
}
// weight of the edge (i,j)-(i,j+1), given the flag byte of (i,j)
template <typename T>
__device__ __forceinline__ T w_right(const Grid<T>& g, int i, int j, unsigned f) {
    T w = T(0);
    if (j + 1 < g.cols) {
        if (!(f & HG_FLAG_CUT_RIGHT)) w = (!g.uniform && g.rows > 1 && (i == 0 || i == g.rows - 1)) ? T(0.125) : T(0.25);
        if (f & HG_FLAG_PEN_RIGHT) w += g.w_g;
    }
    return w;
}
// gradient-penalty part only (used on top of the bilaplacian)
template <typename T>
__device__ __forceinline__ T pen_down(const Grid<T>& g, int i, unsigned f) {
    return (i + 1 < g.rows && (f & HG_FLAG_PEN_DOWN)) ? g.w_g : T(0);
}
template <typename T>
__device__ __forceinline__ T pen_right(const Grid<T>& g, int j, unsigned f) {
    return (j + 1 < g.cols && (f & HG_FLAG_PEN_RIGHT)) ? g.w_g : T(0);
}
template <typename T>
__device__ __forceinline__ T soft_w(const Grid<T>& g, long long p, unsigned f) {
    return (f & HG_FLAG_SOFT) ? (g.soft ? g.soft[p] : g.soft_scalar) : T(0);
}

}  // namespace hg
