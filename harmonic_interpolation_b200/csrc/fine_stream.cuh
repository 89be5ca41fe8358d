// fine_stream.cuh -- streaming fused V-cycle legs of level 0: registers and warp shuffles only.
//
// Same mathematics as the tile legs of fine.cuh / fine_vec.cuh (read the header of fine.cuh first):
//
//   down:  z_R = r_R / d,  z_B = (r_B + W z_R) / d,  res_R = W z_B,  b_1 = P^T res,  x_1 = 0
//   up:    z_R = r_R / d + (P x_1)_R,  z_B = (r_B + W z_R) / d,  z_R = (r_R + W z_B) / d,  r.z
//
// but no shared memory and no block barriers.  A WARP owns a column strip: lane l holds the four
// pixels of columns c0 .. c0+3, c0 = j0 - 4 + 4 l, of the current row, so the warp spans 128 columns
// of which the middle 120 are its own (4 halo columns on either side); left / right neighbours of a
// lane's outer pixels come from the adjacent lane by shuffle.  The warp marches DOWN the rows of its
// chunk (SL_R rows plus 3 rows of run-in and run-out).  Because a red-black sweep only looks one row
// up and down, the whole leg is a software pipeline over a window of four rows held in registers:
//
//   step i:   load r(i+3), code(i+4)                                  [prefetch, see below]
//             red   pixels of row i      (first update)
//             black pixels of row i - 1  (needs red of rows i-2, i-1, i)
//             red   pixels of row i - 2  (residual / second update; needs black of rows i-3, i-2, i-1)
//             down: every other step the restricted residual of coarse row (i-3)/2 is complete
//             up:   row i - 2 is final -> one float4 store per lane, r.z accumulated
//
// All three stages of a step touch the same two of the lane's four columns (p = parity of i, and
// p + 2), which keeps the code small: the step is instantiated for even and odd i.
//
// Memory behaviour: every load / store is a full 16-byte quad per lane, 512 contiguous bytes per
// warp and row; r is fetched two rows ahead and the code bytes three rows ahead (the r load is
// predicated on the quad having a free pixel, so constrained areas cost 1 byte per pixel, not 5).
// Traffic per pixel and channel: down r (4 B) + code (1 B) + 1/4 (b_1, x_1); up r + z + code + 1/4 x_1.
// Instruction count: ~10 per pixel and channel against ~85 for the shared-memory legs.
//
// Chunks (120 columns x SL_R rows) without a free pixel are skipped (activity bytes at setup; the
// downward leg uses activity grown by one pixel because restriction reaches that far).
#pragma once
#include "fine_vec.cuh"

namespace hg {

#define SL_OUT 120        // columns a warp owns
#define SL_R 64           // rows per chunk
#define SL_WPB 8          // warps per block; a warp serves one (chunk, channel) pair

// activity of the chunks: 0 = no free pixel within the chunk grown by `halo` (skipped), 1 = simple
// (every pixel the warp reads is hard or a regular interior pixel -- code 0xAA, no soft weight: the
// lean code path), 2 = general
static __global__ void sl_activity_kernel(const uint8_t* __restrict__ flags, const uint8_t* __restrict__ code, int rows, int cols, int pitch,
                                   int halo, int ns, uint8_t* __restrict__ active) {
    __shared__ int any, general;
    if (threadIdx.x == 0) any = general = 0;
    __syncthreads();
    {
        const int bi = blockIdx.y * SL_R - halo, bj = blockIdx.x * SL_OUT - halo;
        const int w = SL_OUT + 2 * halo, h = SL_R + 2 * halo;
        int mine = 0;
        for (int e = threadIdx.x; e < w * h; e += blockDim.x) {
            const int i = bi + e / w, j = bj + e % w;
            if (i >= 0 && i < rows && j >= 0 && j < cols && !(flags[(long long)i * pitch + j] & HG_FLAG_HARD)) mine = 1;
        }
        if (mine) any = 1;
    }
    {   // everything the warp may read: rows i0 - 3 .. i0 + SL_R + 8, columns j0 - 4 .. j0 + 123
        const int bi = blockIdx.y * SL_R - 3, bj = blockIdx.x * SL_OUT - 4;
        const int w = SL_OUT + 8, h = SL_R + 12;
        int mine = 0;
        for (int e = threadIdx.x; e < w * h; e += blockDim.x) {
            const int i = bi + e / w, j = bj + e % w;
            if (i >= 0 && i < rows && j >= 0 && j < cols) {
                const unsigned f = flags[(long long)i * pitch + j], c = code[(long long)i * pitch + j];
                if (!(f & HG_FLAG_HARD) && (c != 0xAAu || (f & HG_FLAG_SOFT))) mine = 1;
            }
        }
        if (mine) general = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) active[blockIdx.y * ns + blockIdx.x] = (uint8_t)(any ? (general ? 2 : 1) : 0);
}

// Task lists of the streaming kernels, built on the device (one block per list, ordered compaction: chunk order is kept, so
// the lists -- and with them every launch -- are the same from run to run): list `which` = [down lean | down general |
// up lean | up general] at lists + which * n; counts[which] = its length.  split (slab mode with overlap): the down
// lists hold the interior chunks first, then those whose rows touch a ghost band; counts[4 + which] = number of interior ones.
static __global__ void __launch_bounds__(1024)
sl_task_lists_kernel(const uint8_t* __restrict__ up, const uint8_t* __restrict__ down, int n, int ns, int rows, int ghost_top,
                     int ghost_bottom, int split, int* __restrict__ lists, int* __restrict__ counts) {
    const int which = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* a = which < 2 ? down : up;
    const unsigned want = (which & 1) ? 2u : 1u;
    int* out = lists + (long long)which * n;
    __shared__ int wsum[32];
    __shared__ int base, total;
    if (tid == 0) base = 0;
    __syncthreads();
    const int passes = (split && which < 2) ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
        for (int e0 = 0; e0 < n; e0 += 1024) {
            const int e = e0 + tid;
            bool sel = e < n && a[e] == want;
            if (sel && passes == 2) {
                // a down task is a boundary task if its rows 64 t - 3 .. 64 t + 66 touch a ghost band
                const int t = e / ns;
                const bool bnd = (ghost_top > 0 && SL_R * t - 3 < ghost_top) || (ghost_bottom > 0 && SL_R * t + SL_R + 2 >= rows - ghost_bottom);
                sel = bnd == (pass == 1);
            }
            const unsigned bal = __ballot_sync(HG_FULL, sel);
            if (lane == 0) wsum[warp] = __popc(bal);
            __syncthreads();
            if (warp == 0) {
                const int v = wsum[lane];
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(HG_FULL, inc, o);
                    if (lane >= o) inc += u;
                }
                wsum[lane] = inc - v;
                if (lane == 31) total = inc;
            }
            __syncthreads();
            if (sel) out[base + wsum[warp] + __popc(bal & ((1u << lane) - 1u))] = e;
            __syncthreads();
            if (tid == 0) base += total;
            __syncthreads();
        }
        if (pass == 0 && tid == 0 && which < 2) counts[4 + which] = base;
    }
    if (tid == 0) counts[which] = base;
}

// Lean mask words: word (g, qc) describes rows 8g .. 8g+7 of the quad of pixels 4qc .. 4qc+3: bit 4 (row & 7) + p is set where the
// pixel is free (code byte != 0).  The lean kernels read ONE word per lane and eight rows instead of a code word per row: the
// word of a row group is loaded a whole group ahead, so no request of the row FIFO waits for a load (the code-word loads were
// the long-scoreboard stalls of round 1), and the masks cost 1/8 byte of HBM traffic per pixel instead of 1.
static __global__ void sl_qmask_kernel(const uint8_t* __restrict__ code, int rows, int pitch, unsigned* __restrict__ qmask) {
    const int qp = pitch >> 2, qc = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (qc >= qp) return;
    unsigned w = 0u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = 8 * g + r;
        if (i < rows) {
            const unsigned cw = *reinterpret_cast<const unsigned*>(code + ((long long)i * pitch + 4 * qc));
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if ((cw >> (8 * q)) & 0xffu) w |= 1u << (4 * r + q);
        }
    }
    qmask[(long long)g * qp + qc] = w;
}

template <typename T> __device__ __forceinline__ T sl_up1(T v) { return __shfl_up_sync(HG_FULL, v, 1); }
template <typename T> __device__ __forceinline__ T sl_dn1(T v) { return __shfl_down_sync(HG_FULL, v, 1); }

// weighted sum of the four neighbours of a pixel with code byte cb
template <typename T>
__device__ __forceinline__ T sl_nsum(unsigned cb, T up, T dn, T lf, T rt) {
    // (a code of 0 -- hard, or isolated -- also takes this branch: its result is masked by dinv = 0 or
    // multiplies neighbours that do not exist; this keeps warps with hard pixels from diverging)
    if (cb == 0xAAu || cb == 0u) return cb ? T(0.25) * ((up + dn) + (lf + rt)) : T(0);
    T wu, wd, wl, wr;
    decode(cb, wu, wd, wl, wr);
    return wu * up + wd * dn + wl * lf + wr * rt;
}

// what a lane knows about the operator on its quad of one row
template <bool SOFT> struct SlRowOp {
    unsigned cw;      // four code bytes (0 outside the grid)
    unsigned fw;      // four flag bytes (SOFT only)
};

template <typename T, bool SOFT>
struct SlCtx {
    Grid<T> g;
    const uint8_t* code;
    int c0;           // first column of the lane's quad
    bool colin;       // the quad lies inside the pitch
    __device__ __forceinline__ SlRowOp<SOFT> load_op(int i) const {
        SlRowOp<SOFT> o;
        o.cw = 0u; o.fw = 0x01010101u;
        if (colin && i >= 0 && i < g.rows) {
            const long long off = (long long)i * g.pitch + c0;
            o.cw = *reinterpret_cast<const unsigned*>(code + off);
            if (SOFT) o.fw = *reinterpret_cast<const unsigned*>(g.flags + off);
        }
        return o;
    }
    // does the quad hold a pixel that is not hard (so that r can be non-zero there)?
    __device__ __forceinline__ bool any_free(const SlRowOp<SOFT>& o) const {
        return SOFT ? (o.fw & 0x01010101u) != 0x01010101u : o.cw != 0u;
    }
    __device__ __forceinline__ V4<T> load_row(const T* __restrict__ v, int i, const SlRowOp<SOFT>& o) const {
        if (any_free(o)) return ld4(v + (long long)i * g.pitch + c0);
        return zero4<T>();
    }
    __device__ __forceinline__ T dinv(const SlRowOp<SOFT>& o, int i, int p) const {
        const unsigned cb = (o.cw >> (8 * p)) & 0xffu;
        if (!SOFT) {
            if (cb == 0xAAu) return T(1);
            if (cb == 0u) return T(0);
            return T(8) / T((cb & 3u) + ((cb >> 2) & 3u) + ((cb >> 4) & 3u) + ((cb >> 6) & 3u));
        }
        return dinv_of(g, cb, (o.fw >> (8 * p)) & 0xffu, (long long)i * g.pitch + c0 + p);
    }
};

// Gauss-Seidel update of the two pixels p = PAR, PAR + 2 of row `row` (quad zc, right-hand side rc),
// given the quads of the rows above (zu) and below (zd):  zc[p] = (rc[p] + W z) * dinv.
// RESID: zc[p] = W z where the pixel is free, 0 elsewhere (the residual after a sweep from zero).
template <typename T, bool SOFT, int PAR, bool RESID>
__device__ __forceinline__ void sl_update(const SlCtx<T, SOFT>& cx, const SlRowOp<SOFT>& op, int row, V4<T>& zc, const V4<T>& zsrc,
                                          const V4<T>& rc, const V4<T>& zu, const V4<T>& zd) {
    // neighbours along the row come from zsrc (the same row: values of the other colour)
    const T fromL = sl_up1(zsrc.v[3]), fromR = sl_dn1(zsrc.v[0]);
    const T lfA = PAR == 0 ? fromL : zsrc.v[0], rtA = PAR == 0 ? zsrc.v[1] : zsrc.v[2];
    const T lfB = PAR == 0 ? zsrc.v[1] : zsrc.v[2], rtB = PAR == 0 ? zsrc.v[3] : fromR;
    const int pa = PAR, pb = PAR + 2;
    const unsigned cbA = (op.cw >> (8 * pa)) & 0xffu, cbB = (op.cw >> (8 * pb)) & 0xffu;
    const T diA = cx.dinv(op, row, pa), diB = cx.dinv(op, row, pb);
    const T nsA = sl_nsum<T>(cbA, zu.v[pa], zd.v[pa], lfA, rtA), nsB = sl_nsum<T>(cbB, zu.v[pb], zd.v[pb], lfB, rtB);
    if (RESID) {
        zc.v[pa] = diA != T(0) ? nsA : T(0);
        zc.v[pb] = diB != T(0) ? nsB : T(0);
    } else {
        zc.v[pa] = (rc.v[pa] + nsA) * diA;
        zc.v[pb] = (rc.v[pb] + nsB) * diB;
    }
}

// ---- downward leg -------------------------------------------------------------------------------------
template <typename T, bool SOFT>
struct SlDown {
    SlCtx<T, SOFT> cx;
    const T* rk;                 // this channel's plane of r
    Coarse<T> cl;
    T* bk; T* xk;                // this channel's planes of the next level
    int zero_x;                  // also clear x of the next level (only the unfused coarse path needs it)
    int i0, lane;
    bool owner;                  // lanes 1..30 own their columns
    T wjm[2], wjp[2];            // restriction weights of the lane's two coarse columns
    // pipeline registers
    V4<T> z0, z1, z2, z3, r1, s0, s1, s2;
    SlRowOp<SOFT> o0, o1, o2;
    V4<T> rq0, rq1;              // r of rows i+1, i+2
    SlRowOp<SOFT> q0, q1, q2;    // operator of rows i+1, i+2, i+3

    template <int PAR>
    __device__ __forceinline__ void step(int i, const V4<T>& rc) {
        const int pa = PAR, pb = PAR + 2;
        // red pixels of row i
        z0 = zero4<T>();
        z0.v[pa] = rc.v[pa] * cx.dinv(o0, i, pa);
        z0.v[pb] = rc.v[pb] * cx.dinv(o0, i, pb);
        // black pixels of row i - 1
        sl_update<T, SOFT, PAR, false>(cx, o1, i - 1, z1, z1, r1, z2, z0);
        // residual at the red pixels of row i - 2
        s0 = zero4<T>();
        sl_update<T, SOFT, PAR, true>(cx, o2, i - 2, s0, z2, r1, z3, z1);
        if (PAR == 1) {
            // coarse row I sits on row i - 3 (even): centre s1, diagonals in s2 (above) and s0 (below)
            const int ic = i - 3, I = ic >> 1;
            const T upL = sl_up1(s2.v[3]), dnL = sl_up1(s0.v[3]);
            if (ic >= i0 && ic < i0 + SL_R && I < cl.rows && owner && cx.colin) {
                const int Ja = cx.c0 >> 1;
                if (Ja < cl.cols) {
                    const T wim = (T)w1d(-1, I, cx.g.rows, cl.rows), wip = (T)w1d(1, I, cx.g.rows, cl.rows);
                    const T va = s1.v[0] + wim * (wjm[0] * upL + wjp[0] * s2.v[1]) + wip * (wjm[0] * dnL + wjp[0] * s0.v[1]);
                    T vb = s1.v[2] + wim * (wjm[1] * s2.v[1] + wjp[1] * s2.v[3]) + wip * (wjm[1] * s0.v[1] + wjp[1] * s0.v[3]);
                    if (Ja + 1 >= cl.cols) vb = T(0);
                    const long long oc = (long long)I * cl.pitch + Ja;
                    // inactive coarse pixels keep the zeros they were created with
                    if (*reinterpret_cast<const unsigned short*>(cl.code + oc)) {
                        sts2(bk + oc, va, vb);
                        if (zero_x) sts2(xk + oc, T(0), T(0));
                    }
                }
            }
        }
    }

    __device__ __forceinline__ void run() {
        const int ibeg = i0 - 3;
        z1 = z2 = z3 = r1 = s1 = s2 = zero4<T>();
        o1.cw = o2.cw = 0u; o1.fw = o2.fw = 0x01010101u;
        o0 = cx.load_op(ibeg); q0 = cx.load_op(ibeg + 1); q1 = cx.load_op(ibeg + 2); q2 = cx.load_op(ibeg + 3);
        V4<T> rc = fetch(ibeg, o0);
        rq0 = fetch(ibeg + 1, q0);
        rq1 = fetch(ibeg + 2, q1);
#pragma unroll 1
        for (int i = ibeg; i < i0 + SL_R + 3; i += 2) {
            adv<1>(i, rc);          // ibeg is odd
            adv<0>(i + 1, rc);
        }
    }
    __device__ __forceinline__ V4<T> fetch(int i, const SlRowOp<SOFT>& o) { return cx.load_row(rk, i, o); }
    template <int PAR>
    __device__ __forceinline__ void adv(int i, V4<T>& rc) {
        const V4<T> rnew = fetch(i + 3, q2);
        const SlRowOp<SOFT> onew = cx.load_op(i + 4);
        step<PAR>(i, rc);
        z3 = z2; z2 = z1; z1 = z0; r1 = rc; s2 = s1; s1 = s0; o2 = o1; o1 = o0;
        rc = rq0; rq0 = rq1; rq1 = rnew; o0 = q0; q0 = q1; q1 = q2; q2 = onew;
    }
};


// ---- lean paths for simple chunks ----------------------------------------------------------------------
// All weights are 1/4 and the inverse diagonal is 1 at free pixels (code byte != 0), 0 elsewhere.  The
// row window lives in rings of eight slots (row i <-> slot i & 7) and the loop is unrolled eight
// steps, so every slot index is a compile-time constant and the rotation costs no register moves.
// cp.async (LDGSTS) helpers: a lane's private FIFO in shared memory takes the place of prefetch registers
__device__ __forceinline__ void sl_cp16(void* smem, const void* gmem, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;        // 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(bytes) : "memory");
}
// the same with the `ignore-src` predicate of cp.async (zeros are written, the source is not accessed): no size arithmetic
__device__ __forceinline__ void sl_cp16p(void* smem, const void* gmem, unsigned valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("{ .reg .pred p; setp.eq.u32 p, %2, 0; cp.async.cg.shared.global [%0], [%1], 16, p; }" ::"r"(sa), "l"(gmem), "r"(valid) : "memory");
}
__device__ __forceinline__ void sl_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sl_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#define SL_RING 8         // slots of the row rings (row i <-> slot i & 7)
#define SL_PD 3           // r is requested this many rows ahead (codes two rows earlier still)

template <typename T>
struct SlLean {
    const uint8_t* code;  // code bytes (the GENERAL stencil kernel only)
    const unsigned* qm;   // lean mask words (sl_qmask_kernel)
    const T* rk;
    T* ring;              // this lane's quad in slot 0 of the warp's ring of r rows (shared memory); slots are 128 T apart
    int rows, pitch, c0, i0;
    int qp, qg;           // mask words per row group, number of row groups
    bool colin, owner;
    V4<T> Z[8];
    unsigned M[8];        // code words of the rows in flight (the GENERAL stencil kernel only)
    // Mask words of the row groups G0 - 1 .. G0 + 2 and (prefetched) G0 + 3, G0 = (first row of the current block of eight
    // steps) >> 3.  A row is named by its offset O from row 8 G0 (-8 <= O < 24): its ring slot is O & 7, its word W[(O + 8) >> 3].
    unsigned W[4], Wn;
    __device__ __forceinline__ unsigned load_code(int i) const {
        // (32-bit element offsets on the lean paths: the host only takes them when rows * pitch < 2^31)
        return (colin && (unsigned)i < (unsigned)rows) ? *reinterpret_cast<const unsigned*>(code + (i * pitch + c0)) : 0u;
    }
    __device__ __forceinline__ unsigned load_word(int g) const {
        return (colin && (unsigned)g < (unsigned)qg) ? __ldg(qm + (g * qp + (c0 >> 2))) : 0u;
    }
    __device__ __forceinline__ void words_init(int ib) {
#pragma unroll
        for (int k = 0; k < 4; ++k) W[k] = load_word((ib >> 3) - 1 + k);
    }
    __device__ __forceinline__ void words_prefetch(int ib) { Wn = load_word((ib >> 3) + 3); }          // top of a block of eight steps
    __device__ __forceinline__ void words_rotate() { W[0] = W[1]; W[1] = W[2]; W[2] = W[3]; W[3] = Wn; }   // end of the block
    // non-zero if the quad of row O holds a free pixel / if its pixel P is free
    template <int O> __device__ __forceinline__ unsigned quad() const {
        static_assert(O >= -8 && O < 24, "row offset outside the mask words held");
        return W[(O + 8) >> 3] & (0xFu << (4 * ((O + 8) & 7)));
    }
    template <int O, int P> __device__ __forceinline__ unsigned pix() const {
        static_assert(O >= -8 && O < 24, "row offset outside the mask words held");
        return W[(O + 8) >> 3] & (1u << (4 * ((O + 8) & 7) + P));
    }
    // request row i of r into slot `slot` (zeros where the quad has no free pixel); one commit group per row
    __device__ __forceinline__ void request(int i, int slot, unsigned m) {
        // (a caller's mask may come from the NEIGHBOURING rows -- SlLap MODE 1 -- so the row itself is checked here:
        // rows -1 and `rows` lie outside the plane, for the first / last channel outside the allocation)
        if ((unsigned)i >= (unsigned)rows || !colin) m = 0u;
        const T* src = rk + (m ? i * pitch + c0 : 0);
        T* dst = ring + slot * 128;
        sl_cp16(dst, src, m != 0u);
        if (sizeof(T) == 8) sl_cp16(dst + 2, src + 2, m != 0u);
        sl_cp_commit();
    }
    // The lean form: rows are requested in order, so the source is a running pointer (rq: this lane's quad of the next row to
    // request), and the mask words are already zero outside the grid and the pitch -- no row test, no address arithmetic.
    const T* rq;
    __device__ __forceinline__ void request_start(int i) { rq = rk + ((long long)i * pitch + c0); }
    __device__ __forceinline__ void request_next(int slot, unsigned m) {
        T* dst = ring + slot * 128;
        sl_cp16p(dst, rq, m);
        if (sizeof(T) == 8) sl_cp16p(dst + 2, rq + 2, m);
        sl_cp_commit();
        rq += pitch;
    }
    __device__ __forceinline__ V4<T> row(int slot) const { return ld4(ring + slot * 128); }
    template <int S0, int Q> __device__ __forceinline__ void init_requests(int ibeg) {
        if constexpr (Q < SL_PD) {
            request_next((S0 + Q) & 7, quad<S0 + Q>());
            init_requests<S0, Q + 1>(ibeg);
        }
    }
    // S0 = ibeg & 7 (a compile-time fact of the caller): mask words of the first block, r of rows ibeg .. ibeg + SL_PD - 1
    template <int S0> __device__ __forceinline__ void init(int ibeg) {
#pragma unroll
        for (int q = 0; q < 8; ++q) Z[q] = zero4<T>();
        words_init(ibeg);
        request_start(ibeg);
        init_requests<S0, 0>(ibeg);
    }
    // at the top of the step of row i (offset O): request for row i + SL_PD; returns row i of r
    template <int O> __device__ __forceinline__ V4<T> advance(int i) {
        request_next((O + SL_PD) & 7, quad<O + SL_PD>());
        sl_cp_wait<SL_PD>();
        return row(O & 7);
    }
    // red pixels of row O: z = r (+ add) where free
    template <int O, int PAR> __device__ __forceinline__ void red_first(const V4<T>& rc, const T addA, const T addB) {
        constexpr int C = (O + 8) & 7;
        Z[C] = zero4<T>();
        Z[C].v[PAR] = pix<O, PAR>() ? rc.v[PAR] + addA : T(0);
        Z[C].v[PAR + 2] = pix<O, PAR + 2>() ? rc.v[PAR + 2] + addB : T(0);
    }
    // 1/4 of the four neighbours of the pixels PAR, PAR + 2 of slot C (rows above / below: slots CU, CD)
    template <int C, int CU, int CD, int PAR> __device__ __forceinline__ void nsum(T& nA, T& nB) const {
        T lfA, rtA, lfB, rtB;
        if (PAR == 0) { lfA = sl_up1(Z[C].v[3]); rtA = Z[C].v[1]; lfB = Z[C].v[1]; rtB = Z[C].v[3]; }
        else          { lfA = Z[C].v[0]; rtA = Z[C].v[2]; lfB = Z[C].v[2]; rtB = sl_dn1(Z[C].v[0]); }
        nA = T(0.25) * ((Z[CU].v[PAR] + Z[CD].v[PAR]) + (lfA + rtA));
        nB = T(0.25) * ((Z[CU].v[PAR + 2] + Z[CD].v[PAR + 2]) + (lfB + rtB));
    }
    // Gauss-Seidel update of the pixels PAR, PAR + 2 of row O with right-hand side rc
    template <int O, int PAR> __device__ __forceinline__ void update(const V4<T>& rc) {
        constexpr int C = (O + 8) & 7, CU = (O + 7) & 7, CD = (O + 9) & 7;
        T nA, nB;
        nsum<C, CU, CD, PAR>(nA, nB);
        Z[C].v[PAR] = pix<O, PAR>() ? rc.v[PAR] + nA : T(0);
        Z[C].v[PAR + 2] = pix<O, PAR + 2>() ? rc.v[PAR + 2] + nB : T(0);
    }
};

template <typename T>
struct SlLeanDown : SlLean<T> {
    using B = SlLean<T>;
    T* bk; T* xk;
    const uint8_t* ccode;
    int zero_x;
    int grows, crows, ccols, cpitch;
    T wjm[2], wjp[2];
    V4<T> S[8];

    template <int U> __device__ __forceinline__ void step(int ib) {     // ib = i0 - 3 (mod 8: 5)
        constexpr int O = 5 + U;
        constexpr int C = O & 7, C1 = (C + 7) & 7, C2 = (C + 6) & 7, C3 = (C + 5) & 7, C4 = (C + 4) & 7;
        constexpr int PAR = (5 + U) & 1;
        const int i = ib + U;
        const V4<T> rc = B::template advance<O>(i);
        B::template red_first<O, PAR>(rc, T(0), T(0));
        B::template update<O - 1, PAR>(B::row(C1));            // black pixels of row i - 1
        {   // residual at the red pixels of row i - 2
            T nA, nB;
            B::template nsum<C2, C3, C1, PAR>(nA, nB);
            S[C2] = zero4<T>();
            S[C2].v[PAR] = B::template pix<O - 2, PAR>() ? nA : T(0);
            S[C2].v[PAR + 2] = B::template pix<O - 2, PAR + 2>() ? nB : T(0);
        }
        if (PAR == 1) {
            const int ic = i - 3, I = ic >> 1;
            const T upL = sl_up1(S[C4].v[3]), dnL = sl_up1(S[C2].v[3]);
            if (ic >= B::i0 && ic < B::i0 + SL_R && I < crows && B::owner && B::colin) {
                const int Ja = B::c0 >> 1;
                if (Ja < ccols) {
                    const T wim = (T)w1d(-1, I, grows, crows), wip = (T)w1d(1, I, grows, crows);
                    const T va = S[C3].v[0] + wim * (wjm[0] * upL + wjp[0] * S[C4].v[1]) + wip * (wjm[0] * dnL + wjp[0] * S[C2].v[1]);
                    T vb = S[C3].v[2] + wim * (wjm[1] * S[C4].v[1] + wjp[1] * S[C4].v[3]) + wip * (wjm[1] * S[C2].v[1] + wjp[1] * S[C2].v[3]);
                    if (Ja + 1 >= ccols) vb = T(0);
                    const int oc = I * cpitch + Ja;
                    if (*reinterpret_cast<const unsigned short*>(ccode + oc)) {
                        sts2(bk + oc, va, vb);
                        if (zero_x) sts2(xk + oc, T(0), T(0));
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void run() {
#pragma unroll
        for (int q = 0; q < 8; ++q) S[q] = zero4<T>();
        B::template init<5>(B::i0 - 3);
#pragma unroll 1
        for (int ib = B::i0 - 3; ib < B::i0 + SL_R + 3; ib += 8) {
            B::words_prefetch(ib);
            step<0>(ib); step<1>(ib); step<2>(ib); step<3>(ib); step<4>(ib); step<5>(ib); step<6>(ib); step<7>(ib);
            B::words_rotate();
        }
        sl_cp_wait<0>();
    }
};

template <typename T>
struct SlLeanUp : SlLean<T> {
    using B = SlLean<T>;
    T* zk;
    const T* ek;
    int crows, ccols, cpitch;
    double acc;
    T E[4][3];        // coarse rows, slot I & 3: the lane's two columns and the next lane's first

    // (the next lane's first column is NOT fetched here: a shuffle of a value just loaded waits for HBM -- share_e, two steps later)
    template <int SLOT> __device__ __forceinline__ void load_e(int I) {
        T a = T(0), b = T(0);
        if (ek) {
            const int Ic = min(max(I, 0), crows - 1);
            const int J0 = min(max(B::c0 >> 1, 0), ccols - 1), J1 = min(max((B::c0 >> 1) + 1, 0), ccols - 1);
            const T* row = ek + Ic * cpitch;
            a = row[J0]; b = row[J1];
        }
        E[SLOT][0] = a; E[SLOT][1] = b;
    }
    template <int SLOT> __device__ __forceinline__ void share_e() { E[SLOT][2] = sl_dn1(E[SLOT][0]); }
    template <int U> __device__ __forceinline__ void step(int ib) {     // ib = i0 - 2 (mod 8: 6)
        constexpr int O = 6 + U;
        constexpr int C = O & 7, C1 = (C + 7) & 7, C2 = (C + 6) & 7;
        constexpr int PAR = U & 1;
        constexpr int EA = (3 + U / 2) & 3, EB = (EA + 1) & 3, EN = (EA + 2) & 3;      // (ib >> 1) & 3 == 3
        const int i = ib + U;
        const V4<T> rc = B::template advance<O>(i);
        if (PAR == 0) {
            share_e<EB>();                      // loaded two steps ago, first needed by the next (odd) step
            load_e<EN>((i >> 1) + 2);
            B::template red_first<O, 0>(rc, E[EA][0], E[EA][1]);
        } else {
            B::template red_first<O, 1>(rc, T(0.25) * ((E[EA][0] + E[EA][1]) + (E[EB][0] + E[EB][1])),
                                        T(0.25) * ((E[EA][1] + E[EA][2]) + (E[EB][1] + E[EB][2])));
        }
        B::template update<O - 1, PAR>(B::row(C1));            // black pixels of row i - 1
        const V4<T> r2 = B::row(C2);
        B::template update<O - 2, PAR>(r2);                    // red pixels of row i - 2 again
        const int io = i - 2;
        if (io >= B::i0 && io < B::i0 + SL_R && B::owner && B::template quad<O - 2>()) {
            // the row's four products in T, then fp64 (conversions and fp64 adds are the slow instructions here)
            const T part = (B::Z[C2].v[0] * r2.v[0] + B::Z[C2].v[1] * r2.v[1]) + (B::Z[C2].v[2] * r2.v[2] + B::Z[C2].v[3] * r2.v[3]);
            acc += (double)part;
            st4(zo, B::Z[C2]);
        }
        zo += B::pitch;
    }
    T* zo;            // this lane's quad of z in the row that becomes final in the current step (running pointer)
    __device__ __forceinline__ void run() {
        acc = 0.0;
        zo = zk + ((long long)(B::i0 - 4) * B::pitch + B::c0);
        B::template init<6>(B::i0 - 2);
        load_e<3>((B::i0 - 2) >> 1);
        load_e<0>(((B::i0 - 2) >> 1) + 1);
        share_e<3>();
#pragma unroll 1
        for (int ib = B::i0 - 2; ib < B::i0 + SL_R + 2; ib += 8) {
            B::words_prefetch(ib);
            step<0>(ib); step<1>(ib); step<2>(ib); step<3>(ib); step<4>(ib); step<5>(ib); step<6>(ib); step<7>(ib);
            B::words_rotate();
        }
        sl_cp_wait<0>();
    }
};

// this lane's slot-0 quad in the warp's shared-memory ring (dynamic shared memory: SL_RING x 128 T per warp)
template <typename T> __device__ __forceinline__ T* sl_ring() {
    extern __shared__ __align__(16) unsigned char sl_smem[];
    return reinterpret_cast<T*>(sl_smem) + ((threadIdx.x >> 5) * SL_RING * 32 + (threadIdx.x & 31)) * 4;
}

// One warp per task (chunk) and channel; `tasks` lists the chunks of one kind (lean or general).
template <typename T, bool SOFT, bool LEAN>
__global__ void __launch_bounds__(32 * SL_WPB, LEAN ? (sizeof(T) == 4 ? 4 : 2) : 1)
sl_down_kernel(Grid<T> g, const uint8_t* __restrict__ code, const unsigned* __restrict__ qmask, const T* __restrict__ r, Coarse<T> cl,
               const int* __restrict__ tasks, int ntasks, int ns, int zero_x, const int* __restrict__ done) {
    if (done && *done) return;
    const int wid = blockIdx.x * SL_WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;     // warp <-> (task, channel)
    const int idx = wid / g.K, k = wid - idx * g.K;
    if (idx >= ntasks) return;
    const int task = tasks[idx], t = task / ns, s = task - t * ns;
    const int c0 = s * SL_OUT - 4 + 4 * lane;
    const bool colin = c0 >= 0 && c0 < g.pitch, owner = lane >= 1 && lane <= 30;
    T wjm[2], wjp[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int J = (c0 >> 1) + q;
        wjm[q] = (T)w1d(-1, J, g.cols, cl.cols);
        wjp[q] = (T)w1d(1, J, g.cols, cl.cols);
    }
    if (LEAN) {
        SlLeanDown<T> f;
        f.ring = sl_ring<T>();
        f.code = code; f.qm = qmask; f.qp = g.pitch >> 2; f.qg = (g.rows + 7) >> 3; f.rk = r + k * g.plane;
        f.rows = g.rows; f.pitch = g.pitch; f.i0 = t * SL_R;
        f.c0 = c0; f.colin = colin; f.owner = owner;
        f.bk = cl.b + k * cl.plane; f.xk = cl.x + k * cl.plane; f.ccode = cl.code; f.zero_x = zero_x;
        f.grows = g.rows; f.crows = cl.rows; f.ccols = cl.cols; f.cpitch = cl.pitch;
        f.wjm[0] = wjm[0]; f.wjm[1] = wjm[1]; f.wjp[0] = wjp[0]; f.wjp[1] = wjp[1];
        f.run();
    } else {
        SlDown<T, SOFT> d;
        d.lane = lane;
        d.cx.g = g; d.cx.code = code; d.cx.c0 = c0; d.cx.colin = colin; d.owner = owner;
        d.rk = r + k * g.plane;
        d.cl = cl; d.zero_x = zero_x;
        d.bk = cl.b + k * cl.plane; d.xk = cl.x + k * cl.plane;
        d.i0 = t * SL_R;
        d.wjm[0] = wjm[0]; d.wjm[1] = wjm[1]; d.wjp[0] = wjp[0]; d.wjp[1] = wjp[1];
        d.run();
    }
}

// ---- upward leg -----------------------------------------------------------------------------------------
template <typename T, bool SOFT>
struct SlUp {
    SlCtx<T, SOFT> cx;
    const T* rk; T* zk;
    const T* ek;                 // this channel's plane of the next level's x (nullptr: no coarse level)
    int crows, ccols, cpitch;
    int i0, lane;
    bool owner;
    double acc;
    V4<T> z0, z1, z2, z3, r1, r2;
    SlRowOp<SOFT> o0, o1, o2;
    V4<T> rq0, rq1;
    SlRowOp<SOFT> q0, q1, q2;
    T ea[3], eb[3];              // coarse correction under the quad: rows floor(i/2) and floor(i/2) + 1, columns c0/2 .. c0/2 + 2

    // coarse row I (clamped: replicated at the far edge, as P extrapolates by a constant), the lane's two columns
    __device__ __forceinline__ void load_e(int I, T (&e)[3]) const {
        e[0] = e[1] = T(0);
        if (ek) {     // also for lanes beyond the last column: their clamped value is the neighbour's replicated one
            const int Ic = min(max(I, 0), crows - 1);
            const int J0 = min(max(cx.c0 >> 1, 0), ccols - 1), J1 = min(max((cx.c0 >> 1) + 1, 0), ccols - 1);
            const T* row = ek + (long long)Ic * cpitch;
            e[0] = row[J0]; e[1] = row[J1];
        }
    }

    template <int PAR>
    __device__ __forceinline__ void step(int i, const V4<T>& rc) {
        const int pa = PAR, pb = PAR + 2;
        // red pixels of row i: r / d + (P e)
        {
            const T dA = cx.dinv(o0, i, pa), dB = cx.dinv(o0, i, pb);
            T zA = rc.v[pa] * dA, zB = rc.v[pb] * dB;
            if (PAR == 0) {          // even row, even columns: injection
                if (dA != T(0)) zA += ea[0];
                if (dB != T(0)) zB += ea[1];
            } else {                 // odd row, odd columns: mean of four
                if (dA != T(0)) zA += T(0.25) * ((ea[0] + ea[1]) + (eb[0] + eb[1]));
                if (dB != T(0)) zB += T(0.25) * ((ea[1] + ea[2]) + (eb[1] + eb[2]));
            }
            z0 = zero4<T>();
            z0.v[pa] = zA; z0.v[pb] = zB;
        }
        // black pixels of row i - 1
        sl_update<T, SOFT, PAR, false>(cx, o1, i - 1, z1, z1, r1, z2, z0);
        // red pixels of row i - 2 again
        sl_update<T, SOFT, PAR, false>(cx, o2, i - 2, z2, z2, r2, z3, z1);
        // row i - 2 is final
        const int io = i - 2;
        if (io >= i0 && io < i0 + SL_R && io < cx.g.rows && owner && cx.colin) {
            if (cx.any_free(o2)) {
#pragma unroll
                for (int p = 0; p < 4; ++p) acc += (double)(z2.v[p] * r2.v[p]);
                st4(zk + (long long)io * cx.g.pitch + cx.c0, z2);
            }
        }
    }

    template <int PAR>
    __device__ __forceinline__ void adv(int i, V4<T>& rc) {
        const V4<T> rnew = cx.load_row(rk, i + 3, q2);
        const SlRowOp<SOFT> onew = cx.load_op(i + 4);
        if (PAR == 0) {
            // a new pair of coarse rows: floor(i/2) (already held in eb) and floor(i/2) + 1
            ea[0] = eb[0]; ea[1] = eb[1]; ea[2] = eb[2];
            load_e((i >> 1) + 1, eb);
            eb[2] = sl_dn1(eb[0]);
        }
        step<PAR>(i, rc);
        z3 = z2; z2 = z1; z1 = z0; r2 = r1; r1 = rc; o2 = o1; o1 = o0;
        rc = rq0; rq0 = rq1; rq1 = rnew; o0 = q0; q0 = q1; q1 = q2; q2 = onew;
    }

    __device__ __forceinline__ void run() {
        const int ibeg = i0 - 2;     // even
        z1 = z2 = z3 = r1 = r2 = zero4<T>();
        o1.cw = o2.cw = 0u; o1.fw = o2.fw = 0x01010101u;
        o0 = cx.load_op(ibeg); q0 = cx.load_op(ibeg + 1); q1 = cx.load_op(ibeg + 2); q2 = cx.load_op(ibeg + 3);
        V4<T> rc = cx.load_row(rk, ibeg, o0);
        rq0 = cx.load_row(rk, ibeg + 1, q0);
        rq1 = cx.load_row(rk, ibeg + 2, q1);
        load_e(ibeg >> 1, eb);
        eb[2] = sl_dn1(eb[0]);
        acc = 0.0;
#pragma unroll 1
        for (int i = ibeg; i < i0 + SL_R + 2; i += 2) {
            adv<0>(i, rc);
            adv<1>(i + 1, rc);
        }
    }
};

template <typename T, bool SOFT, bool LEAN>
__global__ void __launch_bounds__(32 * SL_WPB, LEAN ? (sizeof(T) == 4 ? 3 : 1) : 1)
sl_up_kernel(Grid<T> g, const uint8_t* __restrict__ code, const unsigned* __restrict__ qmask, const T* __restrict__ r, T* __restrict__ z, Coarse<T> cl,
             int have_coarse, const int* __restrict__ tasks, int ntasks, int ns, int nt, double* __restrict__ part_rz, const int* __restrict__ done) {
    if (done && *done) return;
    const int wid = blockIdx.x * SL_WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;     // warp <-> (task, channel)
    const int idx = wid / g.K, k = wid - idx * g.K;
    if (idx >= ntasks) return;
    const int task = tasks[idx], t = task / ns, s = task - t * ns;
    const int c0 = s * SL_OUT - 4 + 4 * lane;
    const bool colin = c0 >= 0 && c0 < g.pitch, owner = lane >= 1 && lane <= 30;
    double acc;
    if (LEAN) {
        SlLeanUp<T> f;
        f.ring = sl_ring<T>();
        f.code = code; f.qm = qmask; f.qp = g.pitch >> 2; f.qg = (g.rows + 7) >> 3; f.rk = r + k * g.plane; f.zk = z + k * g.plane;
        f.rows = g.rows; f.pitch = g.pitch; f.i0 = t * SL_R;
        f.c0 = c0; f.colin = colin; f.owner = owner;
        f.ek = have_coarse ? cl.x + k * cl.plane : nullptr;
        f.crows = cl.rows; f.ccols = cl.cols; f.cpitch = cl.pitch;
        f.run();
        acc = f.acc;
    } else {
        SlUp<T, SOFT> u;
        u.lane = lane;
        u.cx.g = g; u.cx.code = code; u.cx.c0 = c0; u.cx.colin = colin; u.owner = owner;
        u.rk = r + k * g.plane; u.zk = z + k * g.plane;
        u.ek = have_coarse ? cl.x + k * cl.plane : nullptr;
        u.crows = cl.rows; u.ccols = cl.cols; u.cpitch = cl.pitch;
        u.i0 = t * SL_R;
        u.run();
        acc = u.acc;
    }
    const double sdot = warp_sum(acc);
    if (lane == 0) part_rz[(long long)k * ((long long)ns * nt) + task] = sdot;
}

// ---- q = A p (+ p.q), streaming ------------------------------------------------------------------------------
// The stencil launch of the CG loop in the same style: a warp marches down its chunk with a window of
// three rows of p in registers, rows arriving through the cp.async ring; every row of p is read once
// (no re-read of the rows above / below a tile), quads without a free pixel are neither read (p is 0
// there) nor written, left / right neighbours come by shuffle.  GENERAL: weights decoded from the code
// bytes (grid borders, cut edges); otherwise every free pixel is a regular interior pixel.  Soft
// weights and gradient-penalty edges stay with lap_tile_kernel.
// MODE 0: q = A p, partial sums of p.q  (p is 0 at hard pixels: quads without a free pixel are not read).
// MODE 1: r = -A x, partial sums of r.r  (the true residual of a system whose right-hand side is zero --
//         hard constraints only; x carries the hard values, so a quad is read whenever it or a
//         neighbouring quad holds a free pixel).  TX is the operand type (double for the master
//         iterate of the fp32 mode), TO the result type.
template <typename TX, typename TO, bool GENERAL, int MODE>
struct SlLap : SlLean<TX> {
    using B = SlLean<TX>;
    static constexpr int PD = MODE ? 3 : 4;     // rows requested ahead
    TO* qk;
    double acc;
    V4<TX> P[8];

    // GENERAL (code words): does the lane's quad of a row matter, given the code words of the row above, the row itself, the row below?
    __device__ __forceinline__ unsigned need(unsigned up, unsigned mid, unsigned dn) const {
        if (MODE == 0) return mid;
        return up | mid | dn | sl_up1(mid >> 24) | sl_dn1(mid & 0xffu);
    }
    // lean (mask words): the same for the row at offset R
    template <int R> __device__ __forceinline__ unsigned needq() const {
        if (MODE == 0) return B::template quad<R>();
        return B::template quad<R - 1>() | B::template quad<R>() | B::template quad<R + 1>() | sl_up1(B::template pix<R, 3>()) | sl_dn1(B::template pix<R, 0>());
    }
    // (MODE 1 takes its mask from the neighbouring rows and lanes too: the row and the lane themselves are checked here --
    // rows -1 and `rows` lie outside the plane, for the first / last channel outside the allocation)
    template <int R> __device__ __forceinline__ void request_lean(int i) {
        unsigned m = needq<R>();
        if (MODE == 1 && ((unsigned)i >= (unsigned)B::rows || !B::colin)) m = 0u;
        B::request_next(R & 7, m);
    }
    template <int U> __device__ __forceinline__ void step(int ib) {     // ib = i0 (mod 8: 0)
        constexpr int C = U & 7, CU = (C + 7) & 7, CD = (C + 1) & 7;
        const int i = ib + U;
        // request for row i + 1 + PD, row i + 1 arrives
        if constexpr (GENERAL) {
            B::M[(C + PD + 3) & 7] = B::load_code(i + PD + 3);
            B::request(i + 1 + PD, (C + 1 + PD) & 7, need(B::M[(C + PD) & 7], B::M[(C + 1 + PD) & 7], B::M[(C + 2 + PD) & 7]));
        } else {
            request_lean<U + 1 + PD>(i + 1 + PD);
        }
        sl_cp_wait<PD>();
        P[CD] = B::row(CD);
        unsigned m;
        if constexpr (GENERAL) m = B::M[C]; else m = B::template quad<U>();
        const TX lf0 = sl_up1(P[C].v[3]), rt3 = sl_dn1(P[C].v[0]);
        if (i < B::i0 + SL_R && B::owner && m) {
            V4<TO> out;
            TX part = TX(0);
            if constexpr (GENERAL) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const unsigned cb = (m >> (8 * p)) & 0xffu;
                    const TX lf = p == 0 ? lf0 : P[C].v[p - 1], rt = p == 3 ? rt3 : P[C].v[p + 1];
                    TX v;
                    if (cb == 0xAAu || cb == 0u) {
                        v = cb ? P[C].v[p] - TX(0.25) * ((P[CU].v[p] + P[CD].v[p]) + (lf + rt)) : TX(0);
                    } else {
                        TX wu, wd, wl, wr;
                        decode(cb, wu, wd, wl, wr);
                        v = (wu + wd + wl + wr) * P[C].v[p] - wu * P[CU].v[p] - wd * P[CD].v[p] - wl * lf - wr * rt;
                    }
                    if (MODE == 1) v = -v;
                    out.v[p] = (TO)v;
                    part += (MODE == 0 ? P[C].v[p] : v) * v;
                }
            } else {
                const unsigned free4[4] = {B::template pix<U, 0>(), B::template pix<U, 1>(), B::template pix<U, 2>(), B::template pix<U, 3>()};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const TX lf = p == 0 ? lf0 : P[C].v[p - 1], rt = p == 3 ? rt3 : P[C].v[p + 1];
                    TX v = free4[p] ? P[C].v[p] - TX(0.25) * ((P[CU].v[p] + P[CD].v[p]) + (lf + rt)) : TX(0);
                    if (MODE == 1) v = -v;
                    out.v[p] = (TO)v;
                    part += (MODE == 0 ? P[C].v[p] : v) * v;
                }
            }
            acc += (double)part;
            st4(qo, out);
        }
        qo += B::pitch;
    }
    TO* qo;           // this lane's quad of the result in the current row (running pointer)
    template <int Q> __device__ __forceinline__ void first_requests(int i0) {     // rows i0 - 1 .. i0 + PD
        if constexpr (Q <= PD) {
            request_lean<Q>(i0 + Q);
            first_requests<Q + 1>(i0);
        }
    }
    __device__ __forceinline__ void run() {
        acc = 0.0;
        const int i0 = B::i0;
        qo = qk + ((long long)i0 * B::pitch + B::c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) { P[q] = zero4<TX>(); B::M[q] = 0u; }
        if constexpr (GENERAL) {
            // codes of rows i0 - 2 .. i0 + PD + 2; the ring keeps rows i0 - 1 .. (slots 7, 0, 1, ...); operand rows i0 - 1 .. i0 + PD
            unsigned cc[PD + 5];
#pragma unroll
            for (int q = 0; q < PD + 5; ++q) cc[q] = B::load_code(i0 - 2 + q);
#pragma unroll
            for (int q = -1; q < PD + 3; ++q) B::M[q & 7] = cc[q + 2];
#pragma unroll
            for (int q = -1; q <= PD; ++q) B::request(i0 + q, q & 7, need(cc[q + 1], cc[q + 2], cc[q + 3]));
        } else {
            B::words_init(i0);
            B::request_start(i0 - 1);
            first_requests<-1>(i0);
        }
        sl_cp_wait<PD>();          // rows i0 - 1 and i0 have landed
        P[7] = B::row(7);
        P[0] = B::row(0);
#pragma unroll 1
        for (int ib = i0; ib < i0 + SL_R; ib += 8) {
            if constexpr (!GENERAL) B::words_prefetch(ib);
            step<0>(ib); step<1>(ib); step<2>(ib); step<3>(ib); step<4>(ib); step<5>(ib); step<6>(ib); step<7>(ib);
            if constexpr (!GENERAL) B::words_rotate();
        }
        sl_cp_wait<0>();
    }
};

template <typename TX, typename TO, bool GENERAL, int MODE>
__global__ void __launch_bounds__(32 * SL_WPB, sizeof(TX) == 4 ? 4 : 2)
sl_lap_kernel(Grid<TO> g, const uint8_t* __restrict__ code, const unsigned* __restrict__ qmask, const TX* __restrict__ p, TO* __restrict__ q,
              const int* __restrict__ tasks, int ntasks, int ns, int nt, double* __restrict__ part_pq, const int* __restrict__ done) {
    if (done && *done) return;
    const int wid = blockIdx.x * SL_WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int idx = wid / g.K, k = wid - idx * g.K;
    if (idx >= ntasks) return;
    const int task = tasks[idx], t = task / ns, s = task - t * ns;
    SlLap<TX, TO, GENERAL, MODE> f;
    f.ring = sl_ring<TX>();
    f.code = code; f.qm = qmask; f.qp = g.pitch >> 2; f.qg = (g.rows + 7) >> 3; f.rk = p + k * g.plane; f.qk = q + k * g.plane;
    f.rows = g.rows; f.pitch = g.pitch; f.i0 = t * SL_R;
    f.c0 = s * SL_OUT - 4 + 4 * lane;
    f.colin = f.c0 >= 0 && f.c0 < g.pitch;
    f.owner = lane >= 1 && lane <= 30;
    f.run();
    const double sdot = warp_sum(f.acc);
    if (lane == 0) part_pq[(long long)k * ((long long)ns * nt) + task] = sdot;
}

}  // namespace hg
