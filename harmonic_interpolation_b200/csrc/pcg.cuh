// pcg.cuh -- conjugate-gradient vector kernels with device-resident scalars.
//
// The K channels are K independent systems sharing one operator; they are advanced in
// lock step by the same launches, each with its own alpha / beta.  No scalar ever visits
// the host inside the iteration: reductions leave per-block fp64 partial sums, a
// one-block "scalar stage" adds them in a fixed order and forms alpha / beta, and every
// kernel returns immediately once CgState::done is set, so the host can enqueue
// iterations in batches and look at the state only between batches.
#pragma once
#include "common.cuh"

namespace hg {

struct CgState {
    double rz[HG_MAXK];     // r.z of the current iteration
    double pq[HG_MAXK];     // p.Ap
    double rr[HG_MAXK];     // r.r
    double bb[HG_MAXK];     // ||b_f - A_fc x_c||^2, the reference norm of the stopping rule
    double alpha[HG_MAXK];
    double beta[HG_MAXK];
    double tmp[HG_MAXK];    // local sums awaiting the allreduce over slabs (multi-GPU)
    double tol2;            // rtol^2: the stopping rule on the TRUE residual (outer check)
    double tol2_eff;        // stopping rule of the current inner CG run (recurrence residual)
    double outer_ratio;     // max_k rr/bb of the true residual at the last outer check
    int outer_done;         // 1: the true residual meets rtol
    int pad0;
    double best;            // smallest max_k rr/bb seen so far (stagnation detection)
    int best_iter;
    int iter;
    int done;               // inner run: 1 converged, 2 breakdown, 3 stagnated, 4 maxiter
    int maxiter;
};

// Stagnation: no new best residual for as long as it took to reach the best one plus 500
// iterations.  CG residuals are not monotone, so the window is deliberately generous.
__device__ __forceinline__ void track_stagnation(CgState* st, double worst_ratio) {
    if (worst_ratio < 0.99 * st->best) {
        st->best = worst_ratio;
        st->best_iter = st->iter;
    } else if (st->iter - st->best_iter > 500 + st->best_iter / 2 && st->done == 0) {
        st->done = 3;
    }
}

// Adds `n` partial sums per channel in a fixed order (one block).
__device__ __forceinline__ double ordered_sum(const double* part, int n, double* red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    return block_sum(s, red);
}

// First level of the ordered sum when there are many partial sums (one per tile of a 16384^2 grid is
// a quarter of a million): block (c, k) adds a contiguous chunk of channel k's partials; the
// one-block scalar stage then adds the PRESUM_BLOCKS chunk sums.  Fixed order, deterministic.
#define PRESUM_BLOCKS 128
#define PRESUM_MIN 4096
static __global__ void __launch_bounds__(256)
presum_kernel(const int* __restrict__ done, const double* __restrict__ part, int n, long long stride, double* __restrict__ out) {
    if (done && *done) return;
    __shared__ double red[32];
    const int chunk = (n + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
    const double* pk = part + (long long)blockIdx.y * stride;
    double s = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) s += pk[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[blockIdx.y * gridDim.x + blockIdx.x] = s;
}

// Scalar stages.  `phase` 0: sum the partial sums and finish (single GPU).  Multi-GPU splits a
// stage around an allreduce of CgState::tmp over the slabs: phase 1 only leaves the local sums
// in tmp, phase 2 finishes from tmp.
__device__ __forceinline__ double stage_value(CgState* st, const double* part, int nblk, long long stride, int k, int phase, double* red) {
    if (phase == 2) return st->tmp[k];
    const double v = ordered_sum(part + (long long)k * stride, nblk, red);
    if (phase == 1 && threadIdx.x == 0) st->tmp[k] = v;
    return v;
}

// after q = A p:  alpha = rz / pq
static __global__ void cg_alpha_kernel(CgState* st, const double* __restrict__ part_pq, int nblk, long long stride, int K, int phase) {
    if (st->done) return;
    __shared__ double red[32];
    for (int k = 0; k < K; ++k) {
        const double pq = stage_value(st, part_pq, nblk, stride, k, phase, red);
        if (phase == 1) { __syncthreads(); continue; }
        if (threadIdx.x == 0) {
            st->pq[k] = pq;
            double a = 0.0;
            if (st->rz[k] != 0.0) {
                if (!(pq > 0.0)) st->done = 2;  // NaN or non-positive curvature
                else a = st->rz[k] / pq;
            }
            st->alpha[k] = a;
        }
        __syncthreads();
    }
}

#define VEC_THREADS 256

// ---- unconstrained components (setup) ---------------------------------------------------------------------------------
// The reference's direct solver fails on a singular system (cvxopt raises ArithmeticError, recovery.py:993); an iterative
// solver would quietly return one of the solutions.  A connected component of free pixels whose constant is in the null
// space -- no soft pixel in it, no hard pixel next to it -- is found by spreading an "anchored" mark from the pixels that
// ARE tied to a value.  Two pixels are connected if their edge is uncut OR carries a gradient (dx / dy) constraint of
// non-zero weight: `open` = the flag bits that open a DOWN edge, shifted by one for RIGHT edges; an edge is open if it is
// not cut or has its penalty bit set while penalties count (pen != 0).
// anchor byte: 2 hard, 1 free and anchored, 0 free and not (yet) anchored.
__device__ __forceinline__ bool edge_open_down(unsigned f, int pen) { return !(f & HG_FLAG_CUT_DOWN) || (pen && (f & HG_FLAG_PEN_DOWN)); }
__device__ __forceinline__ bool edge_open_right(unsigned f, int pen) { return !(f & HG_FLAG_CUT_RIGHT) || (pen && (f & HG_FLAG_PEN_RIGHT)); }
static __global__ void __launch_bounds__(256)
anchor_init_kernel(const uint8_t* __restrict__ flags, int rows, int cols, int pitch, int pen, uint8_t* __restrict__ anc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= pitch) return;
    const long long o = (long long)i * pitch + j;
    const unsigned f = flags[o];
    unsigned a = 2;
    if (j < cols && !(f & HG_FLAG_HARD)) {
        a = (f & HG_FLAG_SOFT) ? 1u : 0u;
        if (i > 0 && (flags[o - pitch] & HG_FLAG_HARD) && edge_open_down(flags[o - pitch], pen)) a = 1;
        if (j > 0 && (flags[o - 1] & HG_FLAG_HARD) && edge_open_right(flags[o - 1], pen)) a = 1;
        if (i + 1 < rows && (flags[o + pitch] & HG_FLAG_HARD) && edge_open_down(f, pen)) a = 1;
        if (j + 1 < cols && (flags[o + 1] & HG_FLAG_HARD) && edge_open_right(f, pen)) a = 1;
    }
    anc[o] = (uint8_t)a;
}
// one round: every 32 x 32 tile (+ a one-pixel halo as it stands) is iterated to its local fixed point in shared memory
#define ANC_T 32
static __global__ void __launch_bounds__(256)
anchor_spread_kernel(const uint8_t* __restrict__ flags, int rows, int cols, int pitch, int pen, uint8_t* __restrict__ anc, int* __restrict__ changed) {
    __shared__ uint8_t A[ANC_T + 2][ANC_T + 2];
    __shared__ uint8_t F[ANC_T + 2][ANC_T + 2];
    const int i0 = blockIdx.y * ANC_T - 1, j0 = blockIdx.x * ANC_T - 1;
    for (int e = threadIdx.x; e < (ANC_T + 2) * (ANC_T + 2); e += blockDim.x) {
        const int li = e / (ANC_T + 2), lj = e - li * (ANC_T + 2);
        const int i = i0 + li, j = j0 + lj;
        const bool in = i >= 0 && i < rows && j >= 0 && j < cols;
        A[li][lj] = in ? anc[(long long)i * pitch + j] : (uint8_t)2;
        F[li][lj] = in ? flags[(long long)i * pitch + j] : (uint8_t)(HG_FLAG_HARD | HG_FLAG_CUT_DOWN | HG_FLAG_CUT_RIGHT);
    }
    __syncthreads();
    int mine = 0;
    for (int round = 0; round < 4 * ANC_T; ++round) {
        int ch = 0;
        for (int e = threadIdx.x; e < ANC_T * ANC_T; e += blockDim.x) {
            const int li = 1 + e / ANC_T, lj = 1 + e % ANC_T;
            if (A[li][lj] != 0) continue;
            const unsigned f = F[li][lj];
            bool a = false;
            a = a || (A[li - 1][lj] == 1 && edge_open_down(F[li - 1][lj], pen));
            a = a || (A[li][lj - 1] == 1 && edge_open_right(F[li][lj - 1], pen));
            a = a || (A[li + 1][lj] == 1 && edge_open_down(f, pen));
            a = a || (A[li][lj + 1] == 1 && edge_open_right(f, pen));
            if (a) { A[li][lj] = 1; ch = 1; }
        }
        mine |= ch;
        if (!__syncthreads_or(ch)) break;
    }
    for (int e = threadIdx.x; e < ANC_T * ANC_T; e += blockDim.x) {
        const int li = 1 + e / ANC_T, lj = 1 + e % ANC_T;
        const int i = i0 + li, j = j0 + lj;
        if (i < rows && j < cols && A[li][lj] == 1) anc[(long long)i * pitch + j] = 1;
    }
    if (mine) *changed = 1;
}
static __global__ void __launch_bounds__(256)
anchor_count_kernel(const uint8_t* __restrict__ anc, int rows, int cols, int pitch, unsigned long long* __restrict__ count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    const bool loose = j < cols && anc[(long long)i * pitch + j] == 0;
    const unsigned b = __ballot_sync(HG_FULL, loose);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}

// caller's rows x cols flag bytes -> the pitch-padded plane: padding columns are hard, edge bits that point outside the grid are dropped
static __global__ void __launch_bounds__(256)
pad_flags_kernel(const uint8_t* __restrict__ raw, uint8_t* __restrict__ flags, int rows, int cols, int pitch) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= pitch) return;
    unsigned f = HG_FLAG_HARD;
    if (j < cols) {
        f = raw[(long long)i * cols + j];
        if (i == rows - 1) f &= ~(unsigned)(HG_FLAG_CUT_DOWN | HG_FLAG_PEN_DOWN);
        if (j == cols - 1) f &= ~(unsigned)(HG_FLAG_CUT_RIGHT | HG_FLAG_PEN_RIGHT);
    }
    flags[(long long)i * pitch + j] = (uint8_t)f;
}

// ---- generic (any preconditioner) CG kernels ----------------------------------------------------
// convergence check right after the x / r update, so that the preconditioner launches that follow
// turn into no-ops once the tolerance is met
static __global__ void cg_check_kernel(CgState* st, const double* __restrict__ part_rr, int nblk, long long stride, int K, int phase) {
    if (st->done) return;
    __shared__ double red[32];
    __shared__ int all_conv;
    __shared__ double worst;
    if (threadIdx.x == 0) { all_conv = 1; worst = 0.0; }
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const double rr = stage_value(st, part_rr, nblk, stride, k, phase, red);
        if (phase == 1) { __syncthreads(); continue; }
        if (threadIdx.x == 0) {
            st->rr[k] = rr;
            if (!(rr == rr)) st->done = 2;
            if (rr > st->tol2_eff * st->bb[k]) all_conv = 0;
            if (st->bb[k] > 0.0) worst = fmax(worst, rr / st->bb[k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && phase != 1) {
        st->iter += 1;
        if (all_conv && st->done == 0) st->done = 1;
        if (st->iter >= st->maxiter && st->done == 0) st->done = 4;
        track_stagnation(st, worst);
    }
}

// Outer check on the TRUE residual r = b - A x (computed in the master precision): fixes the
// reference norm bb at the first call, decides final convergence, and arms the next inner CG
// run, which only has to reduce the residual by `inner_red` (fp32 storage cannot do better).
static __global__ void cg_outer_kernel(CgState* st, const double* __restrict__ part_rr, int nblk, long long stride, int K, int first,
                                double inner_red2, int phase) {
    __shared__ double red[32];
    __shared__ int all_conv;
    __shared__ double worst;
    if (threadIdx.x == 0) { all_conv = 1; worst = 0.0; }
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const double rr = stage_value(st, part_rr, nblk, stride, k, phase, red);
        if (phase == 1) { __syncthreads(); continue; }
        if (threadIdx.x == 0) {
            if (first) st->bb[k] = rr;
            st->rr[k] = rr;
            st->rz[k] = 0.0;
            if (!(rr == rr)) st->done = 2;
            if (rr > st->tol2 * st->bb[k]) all_conv = 0;
            if (st->bb[k] > 0.0) worst = fmax(worst, rr / st->bb[k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && phase != 1) {
        st->outer_ratio = worst;
        if (st->done != 2) {
            st->outer_done = all_conv;
            st->done = all_conv ? 1 : (st->iter >= st->maxiter ? 4 : 0);
            st->tol2_eff = fmax(0.25 * st->tol2, inner_red2 * worst);
            st->best = worst;
            st->best_iter = st->iter;
        }
    }
}

// after z = M^-1 r:  rz_new -> beta
static __global__ void cg_beta_rz_kernel(CgState* st, const double* __restrict__ part_rz, int nblk, long long stride, int K, int init, int phase) {
    if (st->done) return;
    __shared__ double red[32];
    for (int k = 0; k < K; ++k) {
        const double rz = stage_value(st, part_rz, nblk, stride, k, phase, red);
        if (phase == 1) { __syncthreads(); continue; }
        if (threadIdx.x == 0) {
            st->beta[k] = (!init && st->rz[k] != 0.0) ? rz / st->rz[k] : 0.0;
            st->rz[k] = rz;
            if (!(rz == rz) || rz < 0.0) st->done = 2;   // the preconditioner must be positive definite
        }
        __syncthreads();
    }
}

// x += alpha p ; r -= alpha q ; partial sums of r.r
template <typename T>
__global__ void __launch_bounds__(VEC_THREADS)
cg_update_kernel(const CgState* __restrict__ st, long long plane, int K, T* __restrict__ x, T* __restrict__ r,
                 const T* __restrict__ p, const T* __restrict__ q, double* __restrict__ part_rr) {
    if (st->done) return;
    __shared__ double red[32];
    const long long n4 = plane >> 2;
    for (int k = 0; k < K; ++k) {
        const T a = (T)st->alpha[k];
        T* xk = x + k * plane; T* rk = r + k * plane;
        const T* pk = p + k * plane; const T* qk = q + k * plane;
        double srr = 0.0;
        for (long long v = (long long)blockIdx.x * VEC_THREADS + threadIdx.x; v < n4; v += (long long)gridDim.x * VEC_THREADS) {
            const long long o = v << 2;
            V4<T> xv = ld4(xk + o), rv = ld4(rk + o);
            const V4<T> pv = ld4(pk + o), qv = ld4(qk + o);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                xv.v[c] += a * pv.v[c];
                rv.v[c] -= a * qv.v[c];
                srr += (double)rv.v[c] * (double)rv.v[c];
            }
            st4(xk + o, xv);
            st4(rk + o, rv);
        }
        const double s1 = block_sum(srr, red);
        if (threadIdx.x == 0) part_rr[(long long)k * gridDim.x + blockIdx.x] = s1;
    }
}

// partial sums of a.b  (a == b gives the squared norm)
template <typename T>
__global__ void __launch_bounds__(VEC_THREADS)
cg_dot_kernel(const CgState* __restrict__ st, long long plane, int K, const T* __restrict__ a, const T* __restrict__ b,
              double* __restrict__ part) {
    if (st && st->done) return;
    __shared__ double red[32];
    const long long n4 = plane >> 2;
    for (int k = 0; k < K; ++k) {
        const T* ak = a + k * plane; const T* bk = b + k * plane;
        double s = 0.0;
        for (long long v = (long long)blockIdx.x * VEC_THREADS + threadIdx.x; v < n4; v += (long long)gridDim.x * VEC_THREADS) {
            const long long o = v << 2;
            const V4<T> av = ld4(ak + o), bv = ld4(bk + o);
#pragma unroll
            for (int c = 0; c < 4; ++c) s += (double)av.v[c] * (double)bv.v[c];
        }
        const double s1 = block_sum(s, red);
        if (threadIdx.x == 0) part[(long long)k * gridDim.x + blockIdx.x] = s1;
    }
}

// p = z + beta p
template <typename T>
__global__ void __launch_bounds__(VEC_THREADS)
cg_pupdate_kernel(const CgState* __restrict__ st, long long plane, int K, T* __restrict__ p, const T* __restrict__ z) {
    if (st->done) return;
    const long long n4 = plane >> 2;
    for (int k = 0; k < K; ++k) {
        const T bt = (T)st->beta[k];
        T* pk = p + k * plane; const T* zk = z + k * plane;
        for (long long v = (long long)blockIdx.x * VEC_THREADS + threadIdx.x; v < n4; v += (long long)gridDim.x * VEC_THREADS) {
            const long long o = v << 2;
            V4<T> pv = ld4(pk + o);
            const V4<T> zv = ld4(zk + o);
#pragma unroll
            for (int c = 0; c < 4; ++c) pv.v[c] = zv.v[c] + bt * pv.v[c];
            st4(pk + o, pv);
        }
    }
}

// z = dinv r   (Jacobi preconditioner as a stand-alone operator)
template <typename T>
__global__ void __launch_bounds__(VEC_THREADS)
jacobi_apply_kernel(long long plane, int K, T* __restrict__ z, const T* __restrict__ r, const T* __restrict__ dinv) {
    const long long n4 = plane >> 2;
    for (int k = 0; k < K; ++k)
        for (long long v = (long long)blockIdx.x * VEC_THREADS + threadIdx.x; v < n4; v += (long long)gridDim.x * VEC_THREADS) {
            const long long o = v << 2;
            const V4<T> rv = ld4(r + k * plane + o), dv = ld4(dinv + o);
            V4<T> zv;
#pragma unroll
            for (int c = 0; c < 4; ++c) zv.v[c] = dv.v[c] * rv.v[c];
            st4(z + k * plane + o, zv);
        }
}

// v = 0 at hard pixels and in the padding
template <typename T>
__global__ void mask_hard_kernel(Grid<T> g, T* __restrict__ v) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.pitch) return;
    const long long o = (long long)i * g.pitch + j;
    if (j >= g.cols || (g.flags[o] & HG_FLAG_HARD))
        for (int k = 0; k < g.K; ++k) v[k * g.plane + o] = T(0);
}

// x = 0 at free pixels (hard values stay): restart from the zero guess
template <typename T>
__global__ void reset_guess_kernel(Grid<T> g, double* __restrict__ x) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.cols) return;
    const long long o = (long long)i * g.pitch + j;
    if (!(g.flags[o] & HG_FLAG_HARD))
        for (int k = 0; k < g.K; ++k) x[k * g.plane + o] = 0.0;
}

// ---- host <-> device value conversion ---------------------------------------------------
// Host arrays are (rows, cols, K) float64 with K innermost; device planes are K planes of
// pitch-strided T.  build_problem_kernel forms, in one pass over the staged host arrays,
//   x   = hard value at hard pixels, x0 (or 0) at free pixels
//   b   = s_p v_p + w_g [ d_down(up) - d_down(p) + d_right(left) - d_right(p) ]  at free pixels
// (C^T W Crhs of recovery.py:927, 1321; +w d at the far endpoint, -w d at (i,j)).
template <typename T, typename TO>
__global__ void build_problem_kernel(Grid<T> g, const double* __restrict__ hard_vals, const double* __restrict__ soft_vals,
                                     const double* __restrict__ d_down, const double* __restrict__ d_right,
                                     const double* __restrict__ x0, const double* __restrict__ rhs, TO* __restrict__ x, TO* __restrict__ b) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.pitch) return;
    const long long o = (long long)i * g.pitch + j;
    const int K = g.K;
    if (j >= g.cols) {
        for (int k = 0; k < K; ++k) { x[k * g.plane + o] = TO(0); b[k * g.plane + o] = TO(0); }
        return;
    }
    const long long h = ((long long)i * g.cols + j) * K;  // host offset
    const unsigned f = g.flags[o];
    if (f & HG_FLAG_HARD) {
        for (int k = 0; k < K; ++k) {
            x[k * g.plane + o] = hard_vals ? (TO)hard_vals[h + k] : TO(0);
            b[k * g.plane + o] = TO(0);
        }
        return;
    }
    const unsigned fu = i > 0 ? g.flags[o - g.pitch] : 0u;
    const unsigned fl = j > 0 ? g.flags[o - 1] : 0u;
    const double sw = (double)soft_w(g, o, f);
    const double wg = (double)g.w_g;
    const bool pd = d_down && (f & HG_FLAG_PEN_DOWN) && i + 1 < g.rows;
    const bool pr = d_right && (f & HG_FLAG_PEN_RIGHT) && j + 1 < g.cols;
    const bool pu = d_down && i > 0 && (fu & HG_FLAG_PEN_DOWN);
    const bool pl = d_right && j > 0 && (fl & HG_FLAG_PEN_RIGHT);
    for (int k = 0; k < K; ++k) {
        double v = rhs ? rhs[h + k] : 0.0;
        if (sw != 0.0 && soft_vals) v += sw * soft_vals[h + k];
        if (pd) v -= wg * d_down[h + k];
        if (pr) v -= wg * d_right[h + k];
        if (pu) v += wg * d_down[h - (long long)g.cols * K + k];
        if (pl) v += wg * d_right[h - K + k];
        b[k * g.plane + o] = (TO)v;
        x[k * g.plane + o] = x0 ? (TO)x0[h + k] : TO(0);
    }
}

// uint8 images (the CLI's data, fill_alpha_harmonic.py:27-45): x = u8 / 255 at hard pixels, 0 elsewhere,
// b = 0;  and back: u8 = clip(round_half_even(255 x), 0, 255)
template <typename T>
__global__ void build_problem_u8_kernel(Grid<T> g, const uint8_t* __restrict__ vals, double* __restrict__ x, double* __restrict__ b,
                                        int zero_b) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.pitch) return;
    const long long o = (long long)i * g.pitch + j;
    const bool hard = j < g.cols && (g.flags[o] & HG_FLAG_HARD);
    const long long h = ((long long)i * g.cols + j) * g.K;
    for (int k = 0; k < g.K; ++k) {
        x[k * g.plane + o] = hard ? (double)vals[h + k] / 255.0 : 0.0;
        if (zero_b) b[k * g.plane + o] = 0.0;
    }
}
template <typename T>
__global__ void export_u8_kernel(Grid<T> g, const double* __restrict__ x, uint8_t* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.cols) return;
    const long long o = (long long)i * g.pitch + j;
    const long long h = ((long long)i * g.cols + j) * g.K;
    for (int k = 0; k < g.K; ++k) out[h + k] = (uint8_t)fmin(fmax(rint(x[k * g.plane + o] * 255.0), 0.0), 255.0);
}

// planar -> (rows, cols, K) float64
template <typename T, typename TI>
__global__ void export_kernel(Grid<T> g, const TI* __restrict__ x, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.cols) return;
    const long long o = (long long)i * g.pitch + j;
    const long long h = ((long long)i * g.cols + j) * g.K;
    for (int k = 0; k < g.K; ++k) out[h + k] = (double)x[k * g.plane + o];
}

// (rows, cols, K) float64 -> planar T (padding zeroed)
template <typename T>
__global__ void import_kernel(Grid<T> g, const double* __restrict__ in, T* __restrict__ x) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= g.pitch) return;
    const long long o = (long long)i * g.pitch + j;
    const long long h = ((long long)i * g.cols + j) * g.K;
    for (int k = 0; k < g.K; ++k) x[k * g.plane + o] = j < g.cols ? (T)in[h + k] : T(0);
}

}  // namespace hg
