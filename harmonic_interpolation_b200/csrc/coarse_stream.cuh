// coarse_stream.cuh -- streaming V-cycle legs of the 9-point (Galerkin) levels: registers and warp shuffles only.
//
// Same mathematics as cdown_kernel / cup_kernel of coarse_fused.cuh (read its header first):
//
//   down:  x_l = S_pre(0, b_l)  (four-colour Gauss-Seidel, colours 0,1,2,3),   b_{l+1} = R (b_l - A_l x_l)
//   up:    x_l <- S_post(x_l + P x_{l+1}, b_l)                                   (colours 3,2,1,0)
//
// but in the style of fine_stream.cuh: a WARP owns a strip of 120 columns (lane = 4 consecutive pixels, 4 halo
// columns on either side, neighbours along a row by shuffle) and marches DOWN its chunk of rows two rows at a
// time.  Colours are c = 2 (i & 1) + (j & 1): even rows hold colours 0 / 1, odd rows 2 / 3, and a colour only
// looks at rows i-1 .. i+1, so a whole leg is a software pipeline over a window of four rows in registers:
//
//   down, step m:   colours 0, 1 of even row E = 2m+2  (own row only)
//                   colours 2, 3 of odd  row O = 2m+1  (rows P = 2m and E)           -> x rows O, E are final: stored
//                   residual of row O (colour 2) and of row P (colours 0, 1; rows Q = 2m-1, P, O)
//                   next-coarser row m = full weighting of the residuals of rows Q, P, O  -> b_{l+1}
//   up,   step m:   rows O, E arrive: x + P x_{l+1}
//                   colours 3, 2 of row O  (rows P, E still un-swept)
//                   colours 1, 0 of row P  (rows Q, O swept)                          -> x rows O, P are final: stored
//
// After a sweep from zero the residual of a pixel is minus its couplings to the colours swept AFTER it, so the
// downward leg needs the right-hand side exactly once, when a pixel is updated.  Every plane of a level crosses
// HBM once per leg, rows arrive through a per-lane cp.async FIFO in shared memory two steps ahead (predicated on the
// quad holding an active pixel), ~25 instructions per pixel and channel instead of ~150 for the tile legs.
//
// Coefficients: code byte per pixel -- 0 inactive, 1 regular (the level's interior stencil `ref`, held in
// registers), 2 general (nine planes in global memory).  A colour step votes: if no lane holds a general pixel it
// runs with the register stencil, otherwise lanes load what they need (predicated).
//
// Chunks: 120 columns x `cr` rows (64 on large levels, 16 on small ones -- more warps); one warp per chunk and
// channel, chunks without an active pixel (grown by one pixel for the downward leg: restriction reaches that far)
// exit at once.
#pragma once
#include "coarse_fused.cuh"
#include "fine_stream.cuh"

namespace hg {

#define CS_OUT 120
#define CS_RING 8                       // rows of the shared-memory FIFO (row <-> slot row & 7)
// warps per block (one block per SM): the pipeline state of a warp is ~165 registers (fp32); 16 warps at 128 registers with
// a few hundred bytes of spills measured faster than 12 warps without (the legs are latency-bound: occupancy wins)
#ifndef CS_WPB32
#define CS_WPB32 16
#endif
template <typename T, bool UP> struct CsCfg { static constexpr int WPB = sizeof(T) == 4 ? CS_WPB32 : 8; };
// Brings a line into L1 without a register destination (a load into a dummy register makes the next writer of that
// register wait for HBM; prefetch.global.L1 gave no measurable hit rate): a 4-byte cp.async.ca into a scratch word.
__device__ __forceinline__ void cs_touch(const void* gmem, void* scratch) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(scratch);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
// 1 / x: fp32 by the approximate reciprocal and one Newton step (no slow path, no branch); fp64 by division
__device__ __forceinline__ float cs_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * (2.0f - x * r);
}
__device__ __forceinline__ double cs_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ V4<float> ldg4(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    V4<float> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
__device__ __forceinline__ V4<double> ldg4(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p + 2));
    V4<double> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; return r;
}
// the broadcast copy of a level's interior stencil: refq[n * 4 + p] = ref[n]
template <typename T>
__global__ void cs_refq_kernel(Coarse<T> L) {
    if (threadIdx.x < 36) L.refq[threadIdx.x] = L.ref[threadIdx.x >> 2];
}

// activity of the chunks of a level: bit 0 any active pixel in the chunk, bit 1 any within the chunk grown by one pixel
static __global__ void cs_activity_kernel(const uint8_t* __restrict__ code, int rows, int cols, int pitch, int cr, int ns, uint8_t* __restrict__ act) {
    __shared__ int a0, a1;
    if (threadIdx.x == 0) a0 = a1 = 0;
    __syncthreads();
    const int bi = blockIdx.y * cr - 1, bj = blockIdx.x * CS_OUT - 1;
    const int w = CS_OUT + 2, h = cr + 2;
    int m0 = 0, m1 = 0;
    for (int e = threadIdx.x; e < w * h; e += blockDim.x) {
        const int li = e / w, lj = e - li * w;
        const int i = bi + li, j = bj + lj;
        if (i >= 0 && i < rows && j >= 0 && j < cols && code[(long long)i * pitch + j]) {
            m1 = 1;
            if (li >= 1 && li <= cr && lj >= 1 && lj <= CS_OUT) m0 = 1;
        }
    }
    if (m0) a0 = 1;
    if (m1) a1 = 1;
    __syncthreads();
    if (threadIdx.x == 0) act[blockIdx.y * ns + blockIdx.x] = (uint8_t)(a0 | (a1 << 1));
}

// value at column q (-1 .. 4) of a lane's quad; lf / rt are the neighbouring lanes' pixels 3 / 0
template <int Q, typename T>
__device__ __forceinline__ T cs_at(const V4<T>& v, T lf, T rt) {
    return Q < 0 ? lf : (Q > 3 ? rt : v.v[Q < 0 ? 0 : (Q > 3 ? 3 : Q)]);
}

template <typename T, bool UP>
struct CsWarp {
    // level
    const uint8_t* code;
    const T* coef;
    long long plane;
    int rows, pitch;
    T ref[9];
    T inv_ref;
    // strip / chunk
    int c0, i0, cr, rbase;          // rbase = i0 - 4: row of relative index 0
    bool colin, owner;
    // operand planes of this channel
    const T* pa;                    // down: b;  up: x
    const T* pb;                    // up: b
    T* xk;                          // x (stored)
    T* ringA; T* ringB;             // this lane's quad in slot 0 of the warp's rings
    float* scratch;                 // this lane's scratch word (target of the L1 touches)
    // next-coarser level
    const uint8_t* ccode;
    T* cb;                          // down: b_{l+1} of this channel
    const T* ce;                    // up: x_{l+1} of this channel
    int crows, ccols, cpitch;
    T wjm[2], wjp[2];
    // registers of the pipeline, indexed by (relative row) mod 12
    V4<T> X[12], R[12];
    unsigned C[12];
    T L3[12], R0[12];               // neighbouring lanes' pixel 3 (left) / pixel 0 (right) of a row: down: final; up: un-swept
    T L3u[12], R0u[12];             // up: the same after the row's sweep
    T EA[4][3];                     // up: next-coarser rows (slot I & 3): the lane's two columns and the next lane's first

    unsigned ccnext;
    __device__ __forceinline__ unsigned load_ccode(int I) const {
        const int Ja = c0 >> 1;
        return (colin && owner && (unsigned)I < (unsigned)crows && Ja < ccols) ? *reinterpret_cast<const unsigned short*>(ccode + ((long long)I * cpitch + Ja)) : 0u;
    }
    __device__ __forceinline__ unsigned load_code(int rho) const {
        const int r = rbase + rho;
        return (colin && (unsigned)r < (unsigned)rows) ? *reinterpret_cast<const unsigned*>(code + ((long long)r * pitch + c0)) : 0u;
    }
    // request relative rows rho, rho + 1 (one commit group); quads without an active pixel are zero-filled, not read
    __device__ __forceinline__ void request(int rho, unsigned m0, unsigned m1) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned m = q ? m1 : m0;
            const int r = rbase + rho + q;
            const long long off = m ? (long long)r * pitch + c0 : 0;
            const int slot = (rho + q) & (CS_RING - 1);
            T* dst = ringA + slot * 128;
            sl_cp16(dst, pa + off, m != 0u);
            if (sizeof(T) == 8) sl_cp16(dst + 2, pa + off + 2, m != 0u);
            if (UP) {
                T* dst2 = ringB + slot * 128;
                sl_cp16(dst2, pb + off, m != 0u);
                if (sizeof(T) == 8) sl_cp16(dst2 + 2, pb + off + 2, m != 0u);
            }
        }
        sl_cp_commit();
    }
    // the coefficients of a quad with a general pixel (144 contiguous bytes, two lines) -> L1 one step before they are
    // needed: the loads inside a step sit on the dependent chain of the sweep, an L2 / HBM round trip there stalls the warp
    __device__ __forceinline__ void prefetch_coefs(int rho, unsigned cw) const {
        if (cw & 0x02020202u) {
            const T* base = coef + (((long long)(rbase + rho) * pitch + c0) >> 2) * 36;
            cs_touch(base, scratch);
            cs_touch(base + 32, scratch);
        }
    }
    __device__ __forceinline__ V4<T> rowA(int rho) const { return ld4(ringA + (rho & (CS_RING - 1)) * 128); }
    __device__ __forceinline__ V4<T> rowB(int rho) const { return ld4(ringB + (rho & (CS_RING - 1)) * 128); }

    // ---- coefficients of a row ------------------------------------------------------------------------------------
    // CfRef: the level's interior stencil, from registers (no lane of the warp holds a general pixel in this row).
    // CfQuad: the quad's nine coefficient vectors from memory -- EVERY active pixel's coefficients are stored, regular ones
    // included, so no per-pixel choice is needed; a lane whose quad holds no general pixel reads the level's broadcast
    // copy of the interior stencil instead (`refq`, same [9][4] layout, always in L1): only general quads cost HBM traffic.
    struct CfRef {
        const T* r; T inv;
        template <int N, int P> __device__ __forceinline__ T get() const { return r[N]; }
        template <int P> __device__ __forceinline__ T invc() const { return inv; }
    };
    struct CfQuad {
        V4<T> v[9];
        template <int N, int P> __device__ __forceinline__ T get() const { return v[N].v[P]; }
        template <int P> __device__ __forceinline__ T invc() const { return cs_rcp(v[4].v[P]); }
    };
    const T* refq;
    template <int SET>
    __device__ __forceinline__ void load_quads(CfQuad& q, int rho, unsigned cw) const {
        const bool gen = (cw & 0x02020202u) != 0u;
        const T* base = gen ? coef + (((long long)(rbase + rho) * pitch + c0) >> 2) * 36 : refq;
#pragma unroll
        for (int n = 0; n < 9; ++n)
            if ((SET >> n) & 1) q.v[n] = ldg4(base + 4 * n);
    }
    // sum over the neighbours in MASK of coefficient * value for pixel P; U / M / D: rows above / own / below with their lane-edge values
    template <int MASK, int P, typename CF>
    __device__ __forceinline__ static T nsum(const CF& cf, const V4<T>& U, T ul, T ur, const V4<T>& M, T ml, T mr, const V4<T>& D, T dl, T dr) {
        T s = T(0);
        if (MASK & 0x008) s += cf.template get<3, P>() * cs_at<P - 1>(M, ml, mr);
        if (MASK & 0x020) s += cf.template get<5, P>() * cs_at<P + 1>(M, ml, mr);
        if (MASK & 0x002) s += cf.template get<1, P>() * cs_at<P>(U, ul, ur);
        if (MASK & 0x080) s += cf.template get<7, P>() * cs_at<P>(D, dl, dr);
        if (MASK & 0x001) s += cf.template get<0, P>() * cs_at<P - 1>(U, ul, ur);
        if (MASK & 0x004) s += cf.template get<2, P>() * cs_at<P + 1>(U, ul, ur);
        if (MASK & 0x040) s += cf.template get<6, P>() * cs_at<P - 1>(D, dl, dr);
        if (MASK & 0x100) s += cf.template get<8, P>() * cs_at<P + 1>(D, dl, dr);
        return s;
    }
    // Gauss-Seidel update of pixels PAR, PAR + 2 of the row M (code word cw), right-hand side rhs
    template <int MASK, int PAR, typename CF>
    __device__ __forceinline__ static void update(const CF& cf, unsigned cw, V4<T>& M, const V4<T>& rhs, const V4<T>& U, T ul, T ur, T ml, T mr,
                                                  const V4<T>& D, T dl, T dr) {
        const T sa = nsum<MASK, PAR>(cf, U, ul, ur, M, ml, mr, D, dl, dr);
        const T sb = nsum<MASK, PAR + 2>(cf, U, ul, ur, M, ml, mr, D, dl, dr);
        M.v[PAR] = (cw & (0xffu << (8 * PAR))) ? (rhs.v[PAR] - sa) * cf.template invc<PAR>() : T(0);
        M.v[PAR + 2] = (cw & (0xffu << (8 * (PAR + 2)))) ? (rhs.v[PAR + 2] - sb) * cf.template invc<PAR + 2>() : T(0);
    }
    // residual (minus the couplings in MASK) of pixels PAR, PAR + 2 of the row M into Rm
    template <int MASK, int PAR, typename CF>
    __device__ __forceinline__ static void resid(const CF& cf, unsigned cw, V4<T>& Rm, const V4<T>& U, T ul, T ur, const V4<T>& M, T ml, T mr,
                                                 const V4<T>& D, T dl, T dr) {
        const T sa = nsum<MASK, PAR>(cf, U, ul, ur, M, ml, mr, D, dl, dr);
        const T sb = nsum<MASK, PAR + 2>(cf, U, ul, ur, M, ml, mr, D, dl, dr);
        Rm.v[PAR] = (cw & (0xffu << (8 * PAR))) ? -sa : T(0);
        Rm.v[PAR + 2] = (cw & (0xffu << (8 * (PAR + 2)))) ? -sb : T(0);
    }
    __device__ __forceinline__ bool any_general(unsigned cw) const { return __any_sync(HG_FULL, (cw & 0x02020202u) != 0u) != 0; }
    __device__ __forceinline__ bool own_row(int r) const { return r >= i0 && r < i0 + cr && r < rows && owner && colin; }

    // ---- downward leg ---------------------------------------------------------------------------------
    template <int SE, typename CF> __device__ __forceinline__ void down_even(const CF& cf, const V4<T>& bE) {
        const T zero = T(0);
        const V4<T> none = zero4<T>();
        X[SE] = zero4<T>();
        update<0, 0>(cf, C[SE], X[SE], bE, none, zero, zero, zero, zero, none, zero, zero);
        R0[SE] = sl_dn1(X[SE].v[0]);
        update<CF_ROW, 1>(cf, C[SE], X[SE], bE, none, zero, zero, zero, R0[SE], none, zero, zero);
        L3[SE] = sl_up1(X[SE].v[3]);
    }
    template <int SO, int SP, int SE, typename CF> __device__ __forceinline__ void down_odd(const CF& cf, const V4<T>& bO) {
        const T zero = T(0);
        const V4<T> none = zero4<T>();
        X[SO] = zero4<T>();
        update<CF_COLDIAG, 0>(cf, C[SO], X[SO], bO, X[SP], L3[SP], R0[SP], zero, zero, X[SE], L3[SE], R0[SE]);
        R0[SO] = sl_dn1(X[SO].v[0]);
        update<CF_ALL8, 1>(cf, C[SO], X[SO], bO, X[SP], L3[SP], R0[SP], zero, R0[SO], X[SE], L3[SE], R0[SE]);
        L3[SO] = sl_up1(X[SO].v[3]);
        R[SO] = zero4<T>();
        resid<CF_ROW, 0>(cf, C[SO], R[SO], none, zero, zero, X[SO], L3[SO], R0[SO], none, zero, zero);
    }
    template <int SP, int SQ, int SO, typename CF> __device__ __forceinline__ void down_resid_even(const CF& cf) {
        R[SP] = zero4<T>();
        resid<CF_ALL8, 0>(cf, C[SP], R[SP], X[SQ], L3[SQ], R0[SQ], X[SP], L3[SP], R0[SP], X[SO], L3[SO], R0[SO]);
        resid<CF_COLDIAG, 1>(cf, C[SP], R[SP], X[SQ], L3[SQ], R0[SQ], X[SP], L3[SP], R0[SP], X[SO], L3[SO], R0[SO]);
    }
    template <int U> __device__ __forceinline__ void step_down(int j, int nsteps) {
        constexpr int SE = (2 * U + 2) % 12, SO = (2 * U + 1) % 12, SP = (2 * U) % 12, SQ = (2 * U + 11) % 12;
        constexpr int A5 = (2 * U + 5) % 12, A6 = (2 * U + 6) % 12, A7 = (2 * U + 7) % 12, A8 = (2 * U + 8) % 12;
        if (j >= nsteps) return;
        C[A7] = load_code(2 * j + 7);
        C[A8] = load_code(2 * j + 8);
        {
            const bool more = j + 2 < nsteps;       // (rows beyond the last step are not needed)
            request(2 * j + 5, more ? C[A5] : 0u, more ? C[A6] : 0u);
        }
        prefetch_coefs(2 * j + 3, C[(2 * U + 3) % 12]);
        prefetch_coefs(2 * j + 4, C[(2 * U + 4) % 12]);
        sl_cp_wait<2>();
        const int rhoO = 2 * j + 1, rhoE = 2 * j + 2, rhoP = 2 * j;
        const V4<T> bO = rowA(rhoO), bE = rowA(rhoE);
        // coefficients of rows E and O first (row O's are in flight while row E is swept), then the sweeps
        const bool gE = any_general(C[SE]), gO = any_general(C[SO]);
        if (gE | gO) {
            CfQuad qE, qO;
            load_quads<0x038>(qE, rhoE, C[SE]);
            load_quads<0x1FF>(qO, rhoO, C[SO]);
            down_even<SE>(qE, bE);                       // even row E: colours 0 and 1 (own row only)
            down_odd<SO, SP, SE>(qO, bO);                // odd row O: colours 2 and 3, then its residual (colour 2)
        } else {
            down_even<SE>(CfRef{ref, inv_ref}, bE);
            down_odd<SO, SP, SE>(CfRef{ref, inv_ref}, bO);
        }
        // rows O and E are final
        const int rO = rbase + rhoO, rE = rbase + rhoE;
        if (own_row(rO) && C[SO]) st4(xk + ((long long)rO * pitch + c0), X[SO]);
        if (own_row(rE) && C[SE]) st4(xk + ((long long)rE * pitch + c0), X[SE]);
        // residual of row P (colours 0 and 1)
        if (any_general(C[SP])) { CfQuad q; load_quads<0x1EF>(q, rhoP, C[SP]); down_resid_even<SP, SQ, SO>(q); }
        else down_resid_even<SP, SQ, SO>(CfRef{ref, inv_ref});
        // next-coarser row I on row P: full weighting (the diagonal neighbours are colour 3: zero residual)
        const T rl = sl_up1(R[SP].v[3]);
        const int rP = rbase + rhoP, I = rP >> 1;
        const unsigned ccw = ccnext;                 // code bytes of the lane's two next-coarser pixels, loaded a step ago
        ccnext = load_ccode(I + 1);
        if (rP >= i0 && rP < i0 + cr && I < crows && owner && colin) {
            const int Ja = c0 >> 1;
            if (Ja < ccols) {
                const T wim = (T)w1d(-1, I, rows, crows), wip = (T)w1d(1, I, rows, crows);
                const T va = R[SP].v[0] + (wjm[0] * rl + wjp[0] * R[SP].v[1]) + (wim * R[SQ].v[0] + wip * R[SO].v[0]);
                T vb = R[SP].v[2] + (wjm[1] * R[SP].v[1] + wjp[1] * R[SP].v[3]) + (wim * R[SQ].v[2] + wip * R[SO].v[2]);
                if (Ja + 1 >= ccols) vb = T(0);
                const long long oc = (long long)I * cpitch + Ja;
                // inactive next-coarser pixels keep the zeros they were created with
                if (ccw) sts2(cb + oc, va, vb);
            }
        }
    }
    __device__ __forceinline__ void run_down() {
        const int nsteps = cr / 2 + 2;
#pragma unroll
        for (int q = 0; q < 12; ++q) { X[q] = zero4<T>(); R[q] = zero4<T>(); C[q] = 0u; L3[q] = R0[q] = T(0); }
        // codes of the rows of steps 0, 1, 2 (relative rows 1 .. 6), data of steps 0 and 1
#pragma unroll
        for (int q = 1; q <= 6; ++q) C[q] = load_code(q);
        request(1, C[1], C[2]);
        request(3, C[3], C[4]);
        ccnext = load_ccode(rbase >> 1);             // step 0's row P is relative row 0
#pragma unroll 1
        for (int jb = 0; jb < nsteps; jb += 6) {
            step_down<0>(jb, nsteps); step_down<1>(jb + 1, nsteps); step_down<2>(jb + 2, nsteps);
            step_down<3>(jb + 3, nsteps); step_down<4>(jb + 4, nsteps); step_down<5>(jb + 5, nsteps);
        }
        sl_cp_wait<0>();
    }

    // ---- upward leg --------------------------------------------------------------------------------------
    // next-coarser row I (clamped: replicated at the far edge, as P extrapolates by a constant): the lane's two columns + the next lane's first
    template <int SLOT> __device__ __forceinline__ void load_e(int I) {
        T a = T(0), b = T(0);
        if (ce) {
            const int Ic = min(max(I, 0), crows - 1);
            const int J0 = min(max(c0 >> 1, 0), ccols - 1), J1 = min(max((c0 >> 1) + 1, 0), ccols - 1);
            const T* row = ce + (long long)Ic * cpitch;
            a = row[J0]; b = row[J1];
        }
        EA[SLOT][0] = a; EA[SLOT][1] = b;
    }
    // the next lane's first column, fetched a step after the load (a shuffle of a value just loaded waits for HBM)
    template <int SLOT> __device__ __forceinline__ void share_e() { EA[SLOT][2] = sl_dn1(EA[SLOT][0]); }
    template <int SO, int SP, int SE, typename CF> __device__ __forceinline__ void up_odd(const CF& cf, const V4<T>& bO) {
        update<CF_ALL8, 1>(cf, C[SO], X[SO], bO, X[SP], L3[SP], R0[SP], L3[SO], R0[SO], X[SE], L3[SE], R0[SE]);
        L3u[SO] = sl_up1(X[SO].v[3]);
        update<CF_ALL8, 0>(cf, C[SO], X[SO], bO, X[SP], L3[SP], R0[SP], L3u[SO], R0[SO], X[SE], L3[SE], R0[SE]);
        R0u[SO] = sl_dn1(X[SO].v[0]);
    }
    template <int SP, int SQ, int SO, typename CF> __device__ __forceinline__ void up_even(const CF& cf, const V4<T>& bP) {
        update<CF_ALL8, 1>(cf, C[SP], X[SP], bP, X[SQ], L3u[SQ], R0u[SQ], L3[SP], R0[SP], X[SO], L3u[SO], R0u[SO]);
        L3u[SP] = sl_up1(X[SP].v[3]);
        update<CF_ALL8, 0>(cf, C[SP], X[SP], bP, X[SQ], L3u[SQ], R0u[SQ], L3u[SP], R0[SP], X[SO], L3u[SO], R0u[SO]);
    }
    template <int U> __device__ __forceinline__ void step_up(int j, int nsteps) {
        constexpr int SE = (2 * U + 2) % 12, SO = (2 * U + 1) % 12, SP = (2 * U) % 12, SQ = (2 * U + 11) % 12;
        constexpr int A5 = (2 * U + 5) % 12, A6 = (2 * U + 6) % 12, A7 = (2 * U + 7) % 12, A8 = (2 * U + 8) % 12;
        // next-coarser rows: step j's rows O, E lie between / on rows m, m + 1 with m = (rbase >> 1) + j; their slots are relative
        // to the chunk (slot = j & 3): E0 = row m, E1 = row m + 1, EN = row m + 2 (loaded now, used by the next step)
        constexpr int E0 = U % 4;
        constexpr int E1 = (E0 + 1) & 3, EN = (E0 + 2) & 3;
        if (j >= nsteps) return;
        C[A7] = load_code(2 * j + 7);
        C[A8] = load_code(2 * j + 8);
        {
            const bool more = j + 2 < nsteps;
            request(2 * j + 5, more ? C[A5] : 0u, more ? C[A6] : 0u);
        }
        const int m = (rbase >> 1) + j;
        share_e<E1>();                             // loaded by the previous step
        load_e<EN>(m + 2);
        prefetch_coefs(2 * j + 3, C[(2 * U + 3) % 12]);
        prefetch_coefs(2 * j + 4, C[(2 * U + 4) % 12]);
        sl_cp_wait<2>();
        const int rhoO = 2 * j + 1, rhoE = 2 * j + 2, rhoP = 2 * j;
        // rows O, E arrive: x + P e at the active pixels
        {
            const unsigned co = C[SO], ce_ = C[SE];
            V4<T> xo = rowA(rhoO), xe = rowA(rhoE);
            const T a0 = EA[E0][0], a1 = EA[E0][1], a2 = EA[E0][2], b0 = EA[E1][0], b1 = EA[E1][1], b2 = EA[E1][2];
            if (co & 0x000000ffu) xo.v[0] += T(0.5) * (a0 + b0);
            if (co & 0x0000ff00u) xo.v[1] += T(0.25) * ((a0 + a1) + (b0 + b1));
            if (co & 0x00ff0000u) xo.v[2] += T(0.5) * (a1 + b1);
            if (co & 0xff000000u) xo.v[3] += T(0.25) * ((a1 + a2) + (b1 + b2));
            if (ce_ & 0x000000ffu) xe.v[0] += b0;
            if (ce_ & 0x0000ff00u) xe.v[1] += T(0.5) * (b0 + b1);
            if (ce_ & 0x00ff0000u) xe.v[2] += b1;
            if (ce_ & 0xff000000u) xe.v[3] += T(0.5) * (b1 + b2);
            X[SO] = xo; X[SE] = xe;
        }
        const V4<T> bO = rowB(rhoO), bP = rowB(rhoP);
        R0[SO] = sl_dn1(X[SO].v[0]); R0[SE] = sl_dn1(X[SE].v[0]);
        L3[SO] = sl_up1(X[SO].v[3]); L3[SE] = sl_up1(X[SE].v[3]);
        // odd row O: colours 3 then 2; rows P and E are still un-swept
        if (any_general(C[SO])) { CfQuad q; load_quads<0x1FF>(q, rhoO, C[SO]); up_odd<SO, SP, SE>(q, bO); }
        else up_odd<SO, SP, SE>(CfRef{ref, inv_ref}, bO);
        // even row P: colours 1 then 0; rows Q and O are swept
        if (any_general(C[SP])) { CfQuad q; load_quads<0x1FF>(q, rhoP, C[SP]); up_even<SP, SQ, SO>(q, bP); }
        else up_even<SP, SQ, SO>(CfRef{ref, inv_ref}, bP);
        const int rO = rbase + rhoO, rP = rbase + rhoP;
        if (own_row(rP) && C[SP]) st4(xk + ((long long)rP * pitch + c0), X[SP]);
        if (own_row(rO) && C[SO]) st4(xk + ((long long)rO * pitch + c0), X[SO]);
    }
    __device__ __forceinline__ void run_up() {
        const int nsteps = cr / 2 + 2;
#pragma unroll
        for (int q = 0; q < 12; ++q) { X[q] = zero4<T>(); C[q] = 0u; L3[q] = R0[q] = L3u[q] = R0u[q] = T(0); }
#pragma unroll
        for (int q = 1; q <= 6; ++q) C[q] = load_code(q);
        request(1, C[1], C[2]);
        request(3, C[3], C[4]);
        load_e<0>(rbase >> 1);
        load_e<1>((rbase >> 1) + 1);
        share_e<0>();
#pragma unroll 1
        for (int jb = 0; jb < nsteps; jb += 12) {
            step_up<0>(jb, nsteps); step_up<1>(jb + 1, nsteps); step_up<2>(jb + 2, nsteps); step_up<3>(jb + 3, nsteps);
            step_up<4>(jb + 4, nsteps); step_up<5>(jb + 5, nsteps); step_up<6>(jb + 6, nsteps); step_up<7>(jb + 7, nsteps);
            step_up<8>(jb + 8, nsteps); step_up<9>(jb + 9, nsteps); step_up<10>(jb + 10, nsteps); step_up<11>(jb + 11, nsteps);
        }
        sl_cp_wait<0>();
    }
};

// this lane's slot-0 quad in ring `which` of the warp (dynamic shared memory: [which][warp][CS_RING][128] T)
template <typename T, bool UP> __device__ __forceinline__ T* cs_ring(int which) {
    extern __shared__ __align__(16) unsigned char sl_smem[];
    return reinterpret_cast<T*>(sl_smem) + (((which * CsCfg<T, UP>::WPB + (threadIdx.x >> 5)) * CS_RING) * 32 + (threadIdx.x & 31)) * 4;
}

template <typename T, bool UP>
__global__ void __launch_bounds__(32 * CsCfg<T, UP>::WPB, 1)
cs_leg_kernel(Coarse<T> L, Coarse<T> Lc, int have_coarse, const uint8_t* __restrict__ act, int cr, int ns, int nchunks, const int* __restrict__ done) {
    if (done && *done) return;
    const int wid = blockIdx.x * CsCfg<T, UP>::WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;     // warp <-> (chunk, channel)
    const int idx = wid / L.K, k = wid - idx * L.K;
    if (idx >= nchunks) return;
    if (!(act[idx] & (UP ? 1 : 2))) return;
    const int t = idx / ns, s = idx - t * ns;
    CsWarp<T, UP> w;
    w.code = L.code; w.coef = L.coef; w.plane = L.plane; w.rows = L.rows; w.pitch = L.pitch;
#pragma unroll
    for (int n = 0; n < 9; ++n) w.ref[n] = L.ref[n];
    w.inv_ref = L.has_ref ? cs_rcp(L.ref[4]) : T(0);
    w.refq = L.refq;
    w.c0 = s * CS_OUT - 4 + 4 * lane;
    w.i0 = t * cr; w.cr = cr; w.rbase = w.i0 - 4;
    w.colin = w.c0 >= 0 && w.c0 < L.pitch;
    w.owner = lane >= 1 && lane <= 30;
    w.xk = (UP ? L.x2 : L.x) + k * L.plane;      // (up: to the second buffer -- neighbouring chunks read the pre-smoothed x as halo)
    w.ringA = cs_ring<T, UP>(0);
    {
        extern __shared__ __align__(16) unsigned char sl_smem[];
        w.scratch = reinterpret_cast<float*>(sl_smem + (size_t)(UP ? 2 : 1) * CsCfg<T, UP>::WPB * CS_RING * 128 * sizeof(T)) + threadIdx.x;
    }
    w.ccode = Lc.code; w.crows = Lc.rows; w.ccols = Lc.cols; w.cpitch = Lc.pitch;
    if (UP) {
        w.pa = L.x + k * L.plane; w.pb = L.b + k * L.plane;
        w.ringB = cs_ring<T, UP>(1);
        w.ce = have_coarse ? Lc.x + k * Lc.plane : nullptr;
        w.cb = nullptr;
        w.run_up();
    } else {
        w.pa = L.b + k * L.plane; w.pb = nullptr; w.ringB = nullptr;
        w.cb = Lc.b + k * Lc.plane; w.ce = nullptr;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int J = (w.c0 >> 1) + q;
            w.wjm[q] = (T)w1d(-1, J, L.cols, Lc.cols);
            w.wjp[q] = (T)w1d(1, J, L.cols, Lc.cols);
        }
        w.run_down();
    }
}

}  // namespace hg
