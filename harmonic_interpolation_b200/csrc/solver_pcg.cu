// solver_pcg.cu -- scalar stages and the MG-PCG / iterative-refinement driver (members of hg::Solver<T>, solver.cuh).
#include "solver.cuh"

namespace hg {

// the four scalar stages; with several slabs each is split around an allreduce of the local sums.
// `part` points at the first owned partial of channel 0, `n` owned partials, `stride` per channel.
// many partial sums: add them in chunks first (presum_kernel), the scalar stage adds the chunk sums
template <typename T>
void Solver<T>::presum(const double*& part, int& n, long long& stride, const int* done) {
    if (n < PRESUM_MIN) return;
    presum_kernel<<<dim3(PRESUM_BLOCKS, K), 256, 0, stream>>>(done, part, n, stride, d_presum);
    part = d_presum; n = PRESUM_BLOCKS; stride = PRESUM_BLOCKS;
    ++launches;
}

template <typename T>
int Solver<T>::stage_alpha(const double* part, int n, long long stride) {
    presum(part, n, stride, &d_state->done);
    if (world <= 1) { cg_alpha_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, 0); return HG_OK; }
    cg_alpha_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, 1);
    HGTRY(allreduce_tmp());
    cg_alpha_kernel<<<1, 32, 0, stream>>>(d_state, part, n, stride, K, 2);
    return HG_OK;
}

template <typename T>
int Solver<T>::stage_check(const double* part, int n, long long stride) {
    presum(part, n, stride, &d_state->done);
    if (world <= 1) { cg_check_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, 0); return HG_OK; }
    cg_check_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, 1);
    HGTRY(allreduce_tmp());
    cg_check_kernel<<<1, 32, 0, stream>>>(d_state, part, n, stride, K, 2);
    return HG_OK;
}

template <typename T>
int Solver<T>::stage_beta(const double* part, int n, long long stride, int init) {
    presum(part, n, stride, &d_state->done);
    if (world <= 1) { cg_beta_rz_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, init, 0); return HG_OK; }
    cg_beta_rz_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, init, 1);
    HGTRY(allreduce_tmp());
    cg_beta_rz_kernel<<<1, 32, 0, stream>>>(d_state, part, n, stride, K, init, 2);
    return HG_OK;
}

template <typename T>
int Solver<T>::stage_outer(const double* part, int n, long long stride, int first, double red2) {
    presum(part, n, stride, nullptr);
    if (world <= 1) { cg_outer_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, first, red2, 0); return HG_OK; }
    cg_outer_kernel<<<1, 1024, 0, stream>>>(d_state, part, n, stride, K, first, red2, 1);
    HGTRY(allreduce_tmp());
    cg_outer_kernel<<<1, 32, 0, stream>>>(d_state, part, n, stride, K, first, red2, 2);
    return HG_OK;
}

template <typename T>
int Solver<T>::solve_device(double rtol, int maxiter, hg_stats* st) {
    if (comm_dead) return fail(HG_ERR_COMM, "slab mode: the communicator of this solver was aborted after an earlier failure");
    const int rc = solve_device_impl(rtol, maxiter, st);
    // not converged / breakdown are collective outcomes (every rank sees the same all-reduced scalars)
    if (world > 1 && rc != HG_OK && rc != HG_ERR_NOT_CONVERGED && rc != HG_ERR_BREAKDOWN) {
        const std::string keep = g_err;
        abort_comm();
        g_err = keep;
    }
    return rc;
}

template <typename T>
int Solver<T>::solve_device_impl(double rtol, int maxiter, hg_stats* st) {
    if (!is_setup || !have_values) return fail(HG_ERR_INVALID, "hg_solve_device: values have not been uploaded");
    if (maxiter <= 0) maxiter = 100000;
    launches = 0;
    const bool mg = use_mg(), mg25 = use_mg25();
    if (world > 1 && !mg) return fail(HG_ERR_INVALID, "the slab mode needs the multigrid preconditioner");
    const int nb_op = op_blocks(), nb_vec = vec_blocks();
    const dim3 pg((pitch + 255) / 256, rows);
    const int nb_res = pg.x * pg.y;
    const size_t vbytes = (size_t)plane * K * sizeof(T);
    if (!d_z) {
        CU(dmalloc(&d_z, vbytes));
        CU(cudaMemsetAsync(d_z, 0, vbytes, stream));
    }
    CgState init;
    memset(&init, 0, sizeof init);
    init.tol2 = rtol * rtol;
    init.maxiter = maxiter;
    CU(cudaMemcpyAsync(d_state, &init, sizeof init, cudaMemcpyHostToDevice, stream));
    CU(cudaEventRecord(ev0, stream));
    int* done = &d_state->done;
    const Grid<T> g = grid();
    // what one inner CG run may promise: fp32 storage reduces a residual by ~1e-5 at best
    const double inner_red = kRefine ? 1e-5 : 0.0;
    const int max_outer = kRefine ? 8 : 4;
    T* e = kRefine ? d_x : reinterpret_cast<T*>(d_xm);   // the vector the inner CG updates
    // large grids: look at the state after every iteration (a sync costs microseconds, an
    // iteration milliseconds); small grids: enqueue batches, the surplus launches are no-ops
    // (the same decision on every slab: ranks must enqueue the same collectives)
    const long long npix = (long long)(world > 1 ? global_rows : rows) * cols;
    const int batch = npix >= (1LL << 22) ? 1 : ((mg || mg25) ? 4 : 32);
    const int nt = mg ? ntiles() : 0;

    CgState host;
    memset(&host, 0, sizeof host);
    double prev_ratio = 0.0;
    int outer = 0;
    for (;; ++outer) {
        // true residual in master precision -> d_r (storage type), ||r||^2 -> state
        if (mg) {
            // master-precision residual with the tile kernel (x, b in fp64 even in fp32 mode)
            HGTRY(exchange(d_xm, 1));
            if (stream_lap() && b_is_zero) {
                launch_sl_resid(d_xm, d_r);
                const int t0 = ghost_top / SL_R, t1 = ghost_bottom ? (rows - ghost_bottom) / SL_R : sl_nt;
                HGTRY(stage_outer(d_spart2 + (long long)t0 * sl_ns, (t1 - t0) * sl_ns, (long long)sl_ns * sl_nt, outer == 0, inner_red * inner_red));
                launches -= 2;
            } else {
                launch_lap_tile<double, 1, 2>(d_xm, b_is_zero ? nullptr : d_bm, d_r, d_lpart, nullptr);
                HGTRY(stage_outer(lap_part(d_lpart), n_own_lap(), n_lap(), outer == 0, inner_red * inner_red));
                --launches;
            }
        } else if (kRefine) {
            resid_master_kernel<T><<<pg, 256, 0, stream>>>(g, op == HG_OP_BILAPLACIAN, d_xm, d_bm, d_r, d_part0);
            HGTRY(stage_outer(d_part0, nb_res, nb_res, outer == 0, inner_red * inner_red));
        } else {
            launch_op<1, 2>(e, d_b, d_r, d_part0, nullptr);
            HGTRY(stage_outer(d_part0, nb_op, nb_op, outer == 0, inner_red * inner_red));
        }
        launches += 2;
        mark(HG_PH_OUTER);
        CU(cudaMemcpyAsync(&host, d_state, sizeof host, cudaMemcpyDeviceToHost, stream));
        HGTRY(sync_watchful());
        if (phase_print) fprintf(stderr, "[hgrid] outer pass %d: true relres %.3e after %d iterations\n", outer, sqrt(host.outer_ratio), host.iter);
        if (host.done) break;                                   // converged (1), NaN (2) or out of iterations (4)
        if (outer >= max_outer) break;
        // a refinement step that did not gain a factor 4 has hit the attainable accuracy
        if (outer > 0 && !(host.outer_ratio < 0.0625 * prev_ratio)) break;
        prev_ratio = host.outer_ratio;

        // inner PCG on A e = r, e0 = 0 (fp32 mode) or continuing from x (fp64 mode)
        if (kRefine) CU(cudaMemsetAsync(d_x, 0, vbytes, stream));
        if (mg) {
            HGTRY(vcycle(d_z, d_r, done, d_tpart1));
            HGTRY(stage_beta_rz(1));
            cg_pupdate_tile_kernel<T><<<tile_grid(), FT_THREADS, 0, stream>>>(d_state, g, d_p, d_z, d_tile_active, 1);
            launches += 2;
        } else {
            if (mg25) HGTRY(vcycle25(d_z, d_r, done));
            else jacobi_apply_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(plane, K, d_z, d_r, d_dinv);
            cg_dot_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(d_state, plane, K, d_r, d_z, d_part1);
            HGTRY(stage_beta(d_part1, nb_vec, nb_vec, 1));
            CU(cudaMemcpyAsync(d_p, d_z, vbytes, cudaMemcpyDeviceToDevice, stream));
            launches += 4;
        }
        while (true) {
            for (int s = 0; s < batch; ++s) {
if (mg) {
                    mark(HG_PH_OTHER);
                    // (no exchange of p: its ghost rows are computed redundantly from z and the old p, and are
                    // valid next to the owned band just as the ghost rows of z are; HGRID_EXCHANGE_P=1 restores it)
                    if (exchange_p) HGTRY(exchange(d_p, 1));
                    HGTRY(lap_and_alpha(d_p, d_q, done));
                    launches -= 2;
                    mark(HG_PH_STENCIL);
                    cg_update_tile_kernel<T><<<tile_grid(), FT_THREADS, 0, stream>>>(d_state, g, e, d_r, d_p, d_q, d_tile_active, d_tpart0);
                    HGTRY(stage_check(tile_part(d_tpart0), n_own_tiles(), nt));
                    mark(HG_PH_UPDATE);
                    HGTRY(vcycle(d_z, d_r, done, d_tpart1));
                    HGTRY(stage_beta_rz(0));
                    cg_pupdate_tile_kernel<T><<<tile_grid(), FT_THREADS, 0, stream>>>(d_state, g, d_p, d_z, d_tile_active, 0);
                    mark(HG_PH_PUPDATE);
                    launches += 6;
                } else {
                    launch_op<0, 1>(d_p, nullptr, d_q, d_part0, done);
                    HGTRY(stage_alpha(d_part0, nb_op, nb_op));
                    cg_update_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(d_state, plane, K, e, d_r, d_p, d_q, d_part0);
                    HGTRY(stage_check(d_part0, nb_vec, nb_vec));
                    if (mg25) HGTRY(vcycle25(d_z, d_r, done));
                    else jacobi_apply_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(plane, K, d_z, d_r, d_dinv);
                    cg_dot_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(d_state, plane, K, d_r, d_z, d_part1);
                    HGTRY(stage_beta(d_part1, nb_vec, nb_vec, 0));
                    cg_pupdate_kernel<T><<<nb_vec, VEC_THREADS, 0, stream>>>(d_state, plane, K, d_p, d_z);
                    launches += 7;
                }
            }
            CU(cudaMemcpyAsync(&host, d_state, sizeof host, cudaMemcpyDeviceToHost, stream));
            HGTRY(sync_watchful());
            if (host.done) break;
        }
        if (kRefine) {
            if (mg) axpy_master_tile_kernel<T><<<tile_grid(), FT_THREADS, 0, stream>>>(g, d_xm, d_x, d_tile_active);
            else axpy_master_kernel<T><<<vec_blocks(), 256, 0, stream>>>(plane * K, d_xm, d_x);
            ++launches;
        }
        if (host.done == 2) break;
    }
    CU(cudaEventRecord(ev1, stream));
    HGTRY(sync_watchful());
    CU(cudaGetLastError());
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ev0, ev1));
    report_phases();

    double relres = sqrt(host.outer_ratio), relres_cg = 0.0;
    for (int k = 0; k < K; ++k)
        if (host.bb[k] > 0.0) relres_cg = std::max(relres_cg, sqrt(host.rr[k] / host.bb[k]));
    int status = HG_OK;
    if (host.done == 2 || relres != relres) status = HG_ERR_BREAKDOWN;
    else if (!host.outer_done) status = HG_ERR_NOT_CONVERGED;   // maxiter, stagnation or attainable accuracy
    if (st) {
        st->iterations = host.iter;
        st->status = status;
        st->relres = relres;
        st->relres_cg = relres_cg;
        st->solve_ms = ms;
        st->launches = launches;
        const double W = sizeof(T);
        const double npx = (double)rows * cols;
        // un-preconditioned CG iteration: K*11W + 1 bytes per pixel; the
        // preconditioner's traffic is not credited
        st->bytes_moved = npx * ((double)host.iter * (K * 11.0 * W + 1.0) + (outer + 1.0) * (K * 3.0 * W + 1.0));
    }
    if (status == HG_ERR_BREAKDOWN) return fail(status, "conjugate gradients broke down (NaN or non-positive curvature): singular or under-constrained system");
    if (status == HG_ERR_NOT_CONVERGED) return fail(status, "the true residual did not reach the requested tolerance (iteration limit, stagnation, or attainable accuracy of the arithmetic)");
    return HG_OK;
}

// the members defined above, for both storage types
template struct Solver<float>;
template struct Solver<double>;

}  // namespace hg
