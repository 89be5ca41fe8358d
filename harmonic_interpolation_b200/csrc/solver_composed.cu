// solver_composed.cu -- composed systems of poisson.py (member Solver<T>::solve_composed, solver.cuh):
//
//     ( L^m + w A^T A ) x = L^(m-1) g + w A^T c,        m = 1 (poisson.py:55-56) or 2 (bilaplacian=True, poisson.py:65-67)
//
// L = the handle's level-0 Laplacian WITHOUT its soft weights (uniform edges: G.T*G / 4 of poisson.py:204-251, cut at the mask
// borders), g = the explicit right-hand side of hg_values.rhs (G.T * target_gradients / 4), A = the sparse rows of
// generate_constraint_matrix (poisson.py:329-366; any number of pixels per row), c their right-hand sides.
//
// Preconditioned CG in fp64 on the free pixels.  Operator: one or two launches of the tile stencil kernel (with the soft
// weights switched off in the Grid it is handed) plus a scatter kernel for the low-rank constraint term.  Preconditioner:
// M^-1 = V^m, V = the multigrid V-cycle of the handle for L + D, D = the handle's soft weights -- the caller sets
// D = diag(w A^T A) for m = 1 and its square root for m = 2, so that (L + D)^2 = L^2 + D^2 + (L D + D L) differs from the
// system by a term of low rank (CG takes a few extra iterations per constraint).  V is a symmetric positive definite
// operator (symmetric V(1,1) cycle), so V^2 is one too.
//
// This is the small-grid path of the reference's gradient-domain tools (poisson.py's own test is 250 x 250): scalars live
// on the host, one stream synchronisation per dot product.
#include "solver.cuh"

namespace hg {

struct Scal { double v[HG_MAXK]; };

// entries of the constraint rows that sit on hard pixels (not allowed: the caller turns such pixels into unknowns)
static __global__ void con_check_kernel(const uint8_t* __restrict__ flags, const long long* __restrict__ off, int nnz, int* __restrict__ bad) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnz && (flags[off[e]] & HG_FLAG_HARD)) atomicAdd(bad, 1);
}

// y += w A^T (A x)   (rhs == nullptr)     or     y += w A^T c   (rhs given); one thread per (constraint row, channel)
static __global__ void con_term_kernel(int n, int K, long long plane, const int* __restrict__ ptr, const long long* __restrict__ off,
                                       const double* __restrict__ coef, double w, const double* __restrict__ x,
                                       const double* __restrict__ rhs, double* __restrict__ y) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * K) return;
    const int c = t / K, k = t - c * K;
    double s;
    if (rhs) s = rhs[(long long)c * K + k];
    else {
        s = 0.0;
        for (int e = ptr[c]; e < ptr[c + 1]; ++e) s += coef[e] * x[k * plane + off[e]];
    }
    s *= w;
    for (int e = ptr[c]; e < ptr[c + 1]; ++e) atomicAdd(y + k * plane + off[e], coef[e] * s);
}

// partial sums of a.b per channel: partial[k * gridDim.x + block]
static __global__ void __launch_bounds__(256) cdot_kernel(long long plane, const double* __restrict__ a, const double* __restrict__ b,
                                                          double* __restrict__ partial) {
    __shared__ double red[32];
    const int k = blockIdx.y;
    const double* ak = a + k * plane; const double* bk = b + k * plane;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) s += ak[i] * bk[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[(long long)k * gridDim.x + blockIdx.x] = s;
}

// x += alpha p, r -= alpha q, partial sums of r.r
static __global__ void __launch_bounds__(256) cupdate_kernel(long long plane, Scal alpha, double* __restrict__ x, double* __restrict__ r,
                                                             const double* __restrict__ p, const double* __restrict__ q,
                                                             double* __restrict__ partial) {
    __shared__ double red[32];
    const int k = blockIdx.y;
    const double a = alpha.v[k];
    const long long o = k * plane;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) {
        x[o + i] += a * p[o + i];
        const double rv = r[o + i] - a * q[o + i];
        r[o + i] = rv;
        s += rv * rv;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[(long long)k * gridDim.x + blockIdx.x] = s;
}

// p = z + beta p
static __global__ void cpupdate_kernel(long long plane, Scal beta, double* __restrict__ p, const double* __restrict__ z) {
    const int k = blockIdx.y;
    const double b = beta.v[k];
    const long long o = k * plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) p[o + i] = z[o + i] + b * p[o + i];
}

// r = b - q, partial sums of r.r
static __global__ void __launch_bounds__(256) cresid_kernel(long long plane, const double* __restrict__ b, const double* __restrict__ q,
                                                            double* __restrict__ r, double* __restrict__ partial) {
    __shared__ double red[32];
    const int k = blockIdx.y;
    const long long o = k * plane;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) {
        const double rv = b[o + i] - q[o + i];
        r[o + i] = rv;
        s += rv * rv;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[(long long)k * gridDim.x + blockIdx.x] = s;
}

template <typename T>
int Solver<T>::solve_composed(const hg_linear_constraints* lc, double rtol, int maxiter, hg_stats* st) {
    if constexpr (kRefine) {
        return fail(HG_ERR_INVALID, "hg_solve_composed: fp64 solvers only (create the handle with HG_PREC_FP64)");
    } else {
        if (!is_setup || !have_values) return fail(HG_ERR_INVALID, "hg_solve_composed: values have not been uploaded");
        if (!use_mg() || world > 1 || w_g != 0.0)
            return fail(HG_ERR_INVALID, "hg_solve_composed: needs a single-GPU Laplacian handle with the multigrid preconditioner and no gradient-penalty edges");
        if (!lc || lc->n_constraints < 0 || (lc->squared != 0 && lc->squared != 1)) return fail(HG_ERR_INVALID, "hg_solve_composed: bad constraint block");
        const int n = lc->n_constraints;
        if (n > 0 && (!lc->row_ptr || !lc->pixel || !lc->coeff || !lc->rhs)) return fail(HG_ERR_INVALID, "hg_solve_composed: NULL constraint arrays");
        const int nnz = n > 0 ? lc->row_ptr[n] : 0;
        if (n > 0 && lc->row_ptr[0] != 0) return fail(HG_ERR_INVALID, "hg_solve_composed: row_ptr[0] must be 0");
        std::vector<long long> off((size_t)std::max(nnz, 1));
        for (int c = 0; c < n; ++c)
            if (lc->row_ptr[c + 1] < lc->row_ptr[c]) return fail(HG_ERR_INVALID, "hg_solve_composed: row_ptr must not decrease");
        for (int e = 0; e < nnz; ++e) {
            const long long p = lc->pixel[e];
            if (p < 0 || p >= (long long)rows * cols) return fail(HG_ERR_INVALID, "hg_solve_composed: pixel index outside the grid");
            off[e] = (p / cols) * pitch + p % cols;
        }
        if (maxiter <= 0) maxiter = 100000;
        const bool squared = lc->squared != 0;
        const double w = lc->weight;
        launches = 0;

        const size_t vbytes = (size_t)plane * K * sizeof(double);
        double *t1 = nullptr, *z1 = nullptr, *cpart = nullptr, *d_coef = nullptr, *d_crhs = nullptr;
        long long* d_off = nullptr;
        int *d_ptr = nullptr, *d_bad = nullptr;
        const int NB = (int)std::min<long long>((plane + 255) / 256, 256);
        std::vector<double> hpart((size_t)NB * K);
        auto cleanup = [&]() {
            dfree(t1); dfree(z1); dfree(cpart); dfree(d_coef); dfree(d_crhs); dfree(d_off); dfree(d_ptr); dfree(d_bad);
        };
        // (the macros return on failure; the few buffers of a failed solve go back to the pool with the handle's stream)
        CU(dmalloc(&t1, vbytes)); CU(dmalloc(&z1, vbytes)); CU(dmalloc(&cpart, (size_t)NB * K * sizeof(double)));
        CU(cudaMemsetAsync(t1, 0, vbytes, stream)); CU(cudaMemsetAsync(z1, 0, vbytes, stream));
        if (!d_z) { CU(dmalloc(&d_z, vbytes)); }
        CU(cudaMemsetAsync(d_z, 0, vbytes, stream));
        CU(cudaMemsetAsync(d_q, 0, vbytes, stream));
        CU(dmalloc(&d_bad, sizeof(int)));
        CU(cudaMemsetAsync(d_bad, 0, sizeof(int), stream));
        if (n > 0) {
            CU(dmalloc(&d_coef, (size_t)nnz * sizeof(double))); CU(dmalloc(&d_crhs, (size_t)n * K * sizeof(double)));
            CU(dmalloc(&d_off, (size_t)nnz * sizeof(long long))); CU(dmalloc(&d_ptr, (size_t)(n + 1) * sizeof(int)));
            CU(cudaMemcpyAsync(d_coef, lc->coeff, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, stream));
            CU(cudaMemcpyAsync(d_crhs, lc->rhs, (size_t)n * K * sizeof(double), cudaMemcpyHostToDevice, stream));
            CU(cudaMemcpyAsync(d_off, off.data(), (size_t)nnz * sizeof(long long), cudaMemcpyHostToDevice, stream));
            CU(cudaMemcpyAsync(d_ptr, lc->row_ptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
            if (nnz > 0) con_check_kernel<<<(nnz + 255) / 256, 256, 0, stream>>>(d_flags, d_off, nnz, d_bad);
        }
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        if (bad) { cleanup(); return fail(HG_ERR_INVALID, "hg_solve_composed: a linear constraint touches a hard pixel (make it an unknown: no HG_FLAG_HARD, all four edges cut)"); }

        CU(cudaEventRecord(ev0, stream));
        Grid<T> gp = grid();           // the Laplacian without the soft weights
        gp.soft = nullptr; gp.soft_scalar = T(0);
        const dim3 lg = lap_grid();
        const int kc = std::min(K, 2);
        dim3 lgz = lg; lgz.z = (K + kc - 1) / kc;
        auto lap_pure = [&](const double* x, double* y) {
            if (kc == 1) lap_tile_kernel<T, T, 0, 0, 1, false><<<lgz, FT_THREADS, 0, stream>>>(gp, d_code, x, nullptr, y, d_tile_active, nullptr, nullptr);
            else lap_tile_kernel<T, T, 0, 0, 2, false><<<lgz, FT_THREADS, 0, stream>>>(gp, d_code, x, nullptr, y, d_tile_active, nullptr, nullptr);
            ++launches;
        };
        const int cblocks = (n * K + 127) / 128;
        // y = (L^m + w A^T A) x
        auto op_apply = [&](const double* x, double* y) {
            if (squared) { lap_pure(x, t1); lap_pure(t1, y); }
            else lap_pure(x, y);
            if (n > 0) { con_term_kernel<<<cblocks, 128, 0, stream>>>(n, K, plane, d_ptr, d_off, d_coef, w, x, nullptr, y); ++launches; }
        };
        const dim3 pg((pitch + 255) / 256, rows);
        // z = V^m r
        auto precond = [&](double* r, double* z) -> int {
            if (squared) {
                HGTRY(vcycle(z1, r, nullptr, d_tpart1));
                mask_hard_kernel<T><<<pg, 256, 0, stream>>>(grid(), z1);
                HGTRY(vcycle(z, z1, nullptr, d_tpart1));
            } else {
                HGTRY(vcycle(z, r, nullptr, d_tpart1));
            }
            mask_hard_kernel<T><<<pg, 256, 0, stream>>>(grid(), z);
            launches += 2;
            return HG_OK;
        };
        const dim3 vg(NB, K);
        auto fetch = [&](double* out) -> int {      // per-channel sums of the partials the last kernel wrote
            CU(cudaMemcpyAsync(hpart.data(), cpart, (size_t)NB * K * sizeof(double), cudaMemcpyDeviceToHost, stream));
            HGTRY(sync_watchful());
            for (int k = 0; k < K; ++k) {
                double s = 0.0;
                for (int b = 0; b < NB; ++b) s += hpart[(size_t)k * NB + b];
                out[k] = s;
            }
            return HG_OK;
        };
        auto dot = [&](const double* a, const double* b, double* out) -> int {
            cdot_kernel<<<vg, 256, 0, stream>>>(plane, a, b, cpart);
            ++launches;
            return fetch(out);
        };

        double* x = d_xm;
        double* b = d_bm;
        // right-hand side: L^(m-1) g + w A^T c at the free pixels
        if (squared) {
            lap_pure(b, t1);
            CU(cudaMemcpyAsync(b, t1, vbytes, cudaMemcpyDeviceToDevice, stream));
        }
        if (n > 0) { con_term_kernel<<<cblocks, 128, 0, stream>>>(n, K, plane, d_ptr, d_off, d_coef, w, nullptr, d_crhs, b); ++launches; }

        double bb[HG_MAXK], rr[HG_MAXK], rz[HG_MAXK], pq[HG_MAXK], tmp[HG_MAXK];
        int rc = dot(b, b, bb);
        if (rc != HG_OK) { cleanup(); return rc; }
        const double tol2 = rtol * rtol;
        int iter = 0, status = HG_ERR_NOT_CONVERGED;
        double relres = 0.0, relres_cg = 0.0, prev_true = 1e300;
        auto worst = [&](const double* v) {
            double m = 0.0;
            for (int k = 0; k < K; ++k) if (bb[k] > 0.0) m = std::max(m, v[k] / bb[k]);
            return m;
        };
        auto any_nan = [&](const double* v) { for (int k = 0; k < K; ++k) if (!(v[k] == v[k])) return true; return false; };

        const int max_pass = 6;      // (the last pass only evaluates the true residual of what the one before it left)
        for (int pass = 0; pass <= max_pass && status == HG_ERR_NOT_CONVERGED; ++pass) {
            // true residual r = b - A x
            op_apply(x, d_q);
            cresid_kernel<<<vg, 256, 0, stream>>>(plane, b, d_q, d_r, cpart);
            ++launches;
            if ((rc = fetch(rr)) != HG_OK) break;
            if (any_nan(rr)) { status = HG_ERR_BREAKDOWN; break; }
            const double now = worst(rr);
            relres = sqrt(now);
            if (now <= tol2) { status = HG_OK; break; }
            if (pass > 0 && !(now < 0.0625 * prev_true)) break;      // a restart that gains less than a factor 4: attainable accuracy
            if (iter >= maxiter || pass == max_pass) break;
            prev_true = now;
            if ((rc = precond(d_r, d_z)) != HG_OK) break;
            if ((rc = dot(d_r, d_z, rz)) != HG_OK) break;
            CU(cudaMemcpyAsync(d_p, d_z, vbytes, cudaMemcpyDeviceToDevice, stream));
            bool inner_done = false;
            double best = worst(rr);
            int best_iter = iter;
            while (!inner_done && iter < maxiter) {
                op_apply(d_p, d_q);
                if ((rc = dot(d_p, d_q, pq)) != HG_OK) break;
                Scal alpha;
                bool broke = any_nan(pq) || any_nan(rz);
                for (int k = 0; k < K; ++k) {
                    const bool live = rr[k] > tol2 * bb[k] * 1e-4 && rz[k] > 0.0;   // a channel far below its tolerance rests
                    if (live && !(pq[k] > 0.0)) broke = true;
                    alpha.v[k] = live && pq[k] > 0.0 ? rz[k] / pq[k] : 0.0;
                }
                if (broke) { status = HG_ERR_BREAKDOWN; break; }
                cupdate_kernel<<<vg, 256, 0, stream>>>(plane, alpha, x, d_r, d_p, d_q, cpart);
                ++launches;
                if ((rc = fetch(rr)) != HG_OK) break;
                ++iter;
                if (any_nan(rr)) { status = HG_ERR_BREAKDOWN; break; }
                relres_cg = sqrt(worst(rr));
                const double wr = worst(rr);
                if (wr <= 0.25 * tol2) { inner_done = true; break; }
                if (wr < best) { best = wr; best_iter = iter; }
                else if (iter - best_iter > 200) { inner_done = true; break; }      // stagnation: let the true residual decide
                if ((rc = precond(d_r, d_z)) != HG_OK) break;
                if ((rc = dot(d_r, d_z, tmp)) != HG_OK) break;
                Scal beta;
                for (int k = 0; k < K; ++k) { beta.v[k] = rz[k] > 0.0 ? tmp[k] / rz[k] : 0.0; rz[k] = tmp[k]; }
                cpupdate_kernel<<<vg, 256, 0, stream>>>(plane, beta, d_p, d_z);
                ++launches;
            }
            if (rc != HG_OK || status == HG_ERR_BREAKDOWN) break;
        }
        CU(cudaEventRecord(ev1, stream));
        const int rs = sync_watchful();
        cleanup();
        if (rc != HG_OK) return rc;
        if (rs != HG_OK) return rs;
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev0, ev1));
        if (st) {
            st->iterations = iter;
            st->status = status;
            st->relres = relres;
            st->relres_cg = relres_cg;
            st->solve_ms = ms;
            st->launches = launches;
            st->bytes_moved = 0.0;
        }
        if (status == HG_ERR_BREAKDOWN) return fail(status, "conjugate gradients broke down (NaN or non-positive curvature): singular or under-constrained system");
        if (status == HG_ERR_NOT_CONVERGED) return fail(status, "the true residual did not reach the requested tolerance (iteration limit, stagnation, or attainable accuracy of the arithmetic)");
        return HG_OK;
    }
}

template struct Solver<float>;
template struct Solver<double>;

}  // namespace hg
