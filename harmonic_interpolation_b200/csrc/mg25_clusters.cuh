// mg25_clusters.cuh -- cluster-flat transfers for the stiff gradient-constraint problems (solve_grid_linear, default
// w_lsq = 1e5, recovery.py:875-884, 926-927) on top of the 25-point hierarchy of mg25.cuh.
//
// A dx / dy row of weight w pins the DIFFERENCE of its two pixels: the error of any reasonable iterate is constant on every
// connected cluster of such edges.  Bilinear interpolation cannot represent "smooth, but flat on thousands of three-pixel
// clusters", the Galerkin coarse operators inherit a penalty w (difference of interpolation weights)^2 at every cluster and
// the coarse correction degenerates (1 200 - 2 600 PCG iterations, growing with the grid; tools/stiff_proto.py).  With the
// level-0 transfers composed with the cluster-averaging projector C -- P' = C P, R' = P^T C -- and the level-1 operator the
// Galerkin product P'^T A P', the stiff term drops out of every coarse operator: 19 / 26 iterations at 128^2 / 256^2 in the
// prototype, and truncating P'^T A P' to the 5 x 5 stencil the hierarchy stores (cut entries, 1e-4 of the operator, lumped
// onto the diagonal) does not change the count.  Clusters whose labels have not settled after CL_ROUNDS rounds (dense
// constraints: one giant cluster, a Poisson-like region where the plain interpolation works) are left alone.
//
// The level-1 operator is found by PROBING: A_1 v = R' A P' v for v = the indicator of every CL_SPACING-th coarse pixel in
// both directions, run through the transfer and residual kernels the V-cycle itself uses, so what is stored is exactly the
// operator the cycle composes (entries of neighbouring probes further than 2 pixels out pollute each other by what is cut
// anyway; the result is symmetrised).
#pragma once
#include "mg25.cuh"

namespace hg {

#define CL_ROUNDS 48
#define CL_SPACING 6

struct Clusters {
    int* label;        // per pixel: the cluster's label (the smallest pixel offset in it), -1 = not in a cluster
    int* count;        // per label: pixels in the cluster
    long long members; // pixels in clusters
};

// stiff link between (i, j) and its down / right neighbour: both free, the edge carries a gradient constraint
__device__ __forceinline__ bool cl_link_down(const uint8_t* fl, int rows, int pitch, int i, long long o) {
    return i + 1 < rows && (fl[o] & HG_FLAG_PEN_DOWN) && !(fl[o] & HG_FLAG_HARD) && !(fl[o + pitch] & HG_FLAG_HARD);
}
__device__ __forceinline__ bool cl_link_right(const uint8_t* fl, int cols, int j, long long o) {
    return j + 1 < cols && (fl[o] & HG_FLAG_PEN_RIGHT) && !(fl[o] & HG_FLAG_HARD) && !(fl[o + 1] & HG_FLAG_HARD);
}

static __global__ void cl_init_kernel(const uint8_t* __restrict__ fl, int rows, int cols, int pitch, int* __restrict__ label) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= pitch) return;
    const long long o = (long long)i * pitch + j;
    bool m = false;
    if (j < cols && !(fl[o] & HG_FLAG_HARD)) {
        m = cl_link_down(fl, rows, pitch, i, o) || cl_link_right(fl, cols, j, o);
        if (i > 0 && cl_link_down(fl, rows, pitch, i - 1, o - pitch)) m = true;
        if (j > 0 && cl_link_right(fl, cols, j - 1, o - 1)) m = true;
    }
    label[o] = m ? (int)o : -1;
}
// one round of min-label propagation over the stiff links
static __global__ void cl_spread_kernel(const uint8_t* __restrict__ fl, int rows, int cols, int pitch, int* __restrict__ label, int* __restrict__ changed) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const long long o = (long long)i * pitch + j;
    const int mine = label[o];
    if (mine < 0) return;
    int best = mine;
    if (cl_link_down(fl, rows, pitch, i, o)) best = min(best, label[o + pitch]);
    if (cl_link_right(fl, cols, j, o)) best = min(best, label[o + 1]);
    if (i > 0 && cl_link_down(fl, rows, pitch, i - 1, o - pitch)) best = min(best, label[o - pitch]);
    if (j > 0 && cl_link_right(fl, cols, j - 1, o - 1)) best = min(best, label[o - 1]);
    if (best < mine) { atomicMin(label + o, best); *changed = 1; }
}
// a piece whose neighbour over a stiff link carries another label belongs to a cluster that has not settled: bad[label] = 1
static __global__ void cl_bad_kernel(const uint8_t* __restrict__ fl, int rows, int cols, int pitch, const int* __restrict__ label, int* __restrict__ bad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const long long o = (long long)i * pitch + j;
    const int mine = label[o];
    if (mine < 0) return;
    bool b = false;
    if (cl_link_down(fl, rows, pitch, i, o) && label[o + pitch] != mine) b = true;
    if (cl_link_right(fl, cols, j, o) && label[o + 1] != mine) b = true;
    if (i > 0 && cl_link_down(fl, rows, pitch, i - 1, o - pitch) && label[o - pitch] != mine) b = true;
    if (j > 0 && cl_link_right(fl, cols, j - 1, o - 1) && label[o - 1] != mine) b = true;
    if (b) bad[mine] = 1;
}
// drops the unsettled pieces, counts the members of the others (count and bad share one plane: bad is read before count is built)
static __global__ void cl_drop_kernel(int rows, int cols, int pitch, int* __restrict__ label, const int* __restrict__ bad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const long long o = (long long)i * pitch + j;
    const int mine = label[o];
    if (mine >= 0 && bad[mine]) label[o] = -1;
}
static __global__ void cl_count_kernel(int rows, int cols, int pitch, const int* __restrict__ label, int* __restrict__ count, unsigned long long* __restrict__ members) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const int mine = label[(long long)i * pitch + j];
    if (mine < 0) return;
    atomicAdd(count + mine, 1);
    atomicAdd(members, 1ULL);
}

// v <- C v: every member of a cluster takes the cluster's mean (two passes; sum is a zeroed K-plane scratch)
template <typename T>
__global__ void cl_sum_kernel(int rows, int cols, int pitch, long long plane, int K, const int* __restrict__ label, const T* __restrict__ v, T* __restrict__ sum) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const long long o = (long long)i * pitch + j;
    const int mine = label[o];
    if (mine < 0) return;
    for (int k = 0; k < K; ++k) atomicAdd(sum + k * plane + mine, v[k * plane + o]);
}
template <typename T>
__global__ void cl_mean_kernel(int rows, int cols, int pitch, long long plane, int K, const int* __restrict__ label, const int* __restrict__ count,
                               const T* __restrict__ sum, T* __restrict__ v) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const long long o = (long long)i * pitch + j;
    const int mine = label[o];
    if (mine < 0) return;
    const T inv = T(1) / (T)count[mine];
    for (int k = 0; k < K; ++k) v[k * plane + o] = sum[k * plane + mine] * inv;
}
// x += d
template <typename T>
__global__ void cl_add_kernel(long long n, T* __restrict__ x, const T* __restrict__ d) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] += d[i];
}

// ---- probing of the level-1 operator --------------------------------------------------------------------------------
// v = 1 at the coarse pixels (I, J) with I mod s == u and J mod s == w (s <= 0: everywhere), 0 elsewhere
template <typename T>
__global__ void cl_probe_fill_kernel(int rows, int cols, int pitch, int s, int u, int w, T* __restrict__ v) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x, I = blockIdx.y;
    if (J >= pitch) return;
    v[(long long)I * pitch + J] = (J < cols && (s <= 0 || (I % s == u && J % s == w))) ? T(1) : T(0);
}
// y = -(A_1 v): entry (a, b) of row (I, J) is read off where the probed neighbour (I + a, J + b) carried the 1
template <typename T>
__global__ void cl_probe_read_kernel(int rows, int cols, int pitch, long long plane, int s, int u, int w, const T* __restrict__ y, T* __restrict__ raw) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x, I = blockIdx.y;
    if (J >= cols) return;
    const long long o = (long long)I * pitch + J;
    for (int a = -2; a <= 2; ++a)
        for (int b = -2; b <= 2; ++b) {
            const int Ia = I + a, Jb = J + b;
            if (Ia < 0 || Ia >= rows || Jb < 0 || Jb >= cols) continue;
            if (Ia % s == u && Jb % s == w) raw[((a + 2) * 5 + (b + 2)) * plane + o] = -y[o];
        }
}
// symmetrise, lump what the 5 x 5 stencil cuts onto the diagonal (rowsum = -(A_1 1)), clear rows without a diagonal
template <typename T>
__global__ void cl_probe_finish_kernel(int rows, int cols, int pitch, long long plane, const T* __restrict__ raw, const T* __restrict__ negrowsum,
                                       T* __restrict__ coef) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x, I = blockIdx.y;
    if (J >= pitch) return;
    const long long o = (long long)I * pitch + J;
    const bool live = J < cols && raw[12 * plane + o] > T(0);
    T off = T(0);
    for (int a = -2; a <= 2; ++a)
        for (int b = -2; b <= 2; ++b) {
            const int c = (a + 2) * 5 + (b + 2);
            if (c == 12) continue;
            T v = T(0);
            const int Ia = I + a, Jb = J + b;
            if (live && Ia >= 0 && Ia < rows && Jb >= 0 && Jb < cols) {
                const long long on = (long long)Ia * pitch + Jb;
                if (raw[12 * plane + on] > T(0)) v = T(0.5) * (raw[c * plane + o] + raw[(24 - c) * plane + on]);
            }
            coef[c * plane + o] = v;
            off += v;
        }
    coef[12 * plane + o] = live ? -negrowsum[o] - off : T(0);
}

}  // namespace hg
