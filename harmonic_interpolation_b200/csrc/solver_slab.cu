// solver_slab.cu -- multi-GPU slab mode: communicator, gather of the replicated levels, failure handling (members of hg::Solver<T>, solver.cuh).
#include "solver.cuh"

namespace hg {

template <typename T>
int Solver<T>::gather_level(const T* local, long long lplane, T* global, long long gplane, int pitch_, int nplanes) {
    const ncclDataType_t dt = sizeof(T) == 8 ? ncclDouble : ncclFloat;
    const int gtL = ghost_top >> HG_GATHER_LEVEL;
    int maxrows = 0;
    for (int q = 0; q < world; ++q) maxrows = std::max(maxrows, gl_end(q) - gl_start(q));
    const size_t block = (size_t)nplanes * maxrows * pitch_;
    if (block * (world + 1) > gbuf_elems) {
        dfree(d_gsend);
        d_gsend = nullptr;
        gbuf_elems = block * (world + 1);
        CU(dmalloc(&d_gsend, gbuf_elems * sizeof(T)));
        d_grecv = d_gsend + block;
    } else {
        d_grecv = d_gsend + block;
    }
    const int myrows = gl_end(rank) - gl_start(rank);
    for (int c = 0; c < nplanes; ++c)
        CU(cudaMemcpyAsync(d_gsend + (size_t)c * maxrows * pitch_, local + (long long)c * lplane + (long long)gtL * pitch_,
                           (size_t)myrows * pitch_ * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    NC(g_nccl.AllGather(d_gsend, d_grecv, block, dt, comm, stream));
    gather_unpack_kernel<T><<<dim3((pitch_ + 255) / 256, maxrows, world * nplanes), 256, 0, stream>>>(
        d_grecv, global, gplane, pitch_, maxrows, nplanes, d_gl_rows);
    return HG_OK;
}

template <typename T>
int Solver<T>::build_global_hierarchy() {
    {
        std::vector<int> gl(2 * world);
        for (int q = 0; q < world; ++q) { gl[2 * q] = gl_start(q); gl[2 * q + 1] = gl_end(q) - gl_start(q); }
        if (!d_gl_rows) CU(dmalloc(&d_gl_rows, 2 * world * sizeof(int)));
        CU(cudaMemcpyAsync(d_gl_rows, gl.data(), 2 * world * sizeof(int), cudaMemcpyHostToDevice, stream));
        HGTRY(sync_watchful());
    }
    const Coarse<T>& Ll = coarse.back();      // local level HG_GATHER_LEVEL
    Coarse<T> G;
    HGTRY(alloc_level(G, chain(global_rows, HG_GATHER_LEVEL), Ll.cols));
    if (G.pitch != Ll.pitch) return fail(HG_ERR_INVALID, "slab mode: pitch mismatch at the gather level");
    HGTRY(gather_level(Ll.coef, 0, G.coef, 0, 9 * G.pitch, 1));     // (quad-interleaved: a row of the level is 9 * pitch contiguous values)
    HGTRY(finish_level(G, HG_GATHER_LEVEL));
    gcoarse.push_back(G);
    int r = G.rows, c = G.cols;
    while ((long long)r * c > 256) {
        Coarse<T> L;
        HGTRY(alloc_level(L, (r + 1) / 2, (c + 1) / 2));
        dim3 cg((L.pitch + 127) / 128, L.rows);
        CoarseAcc<T> ca{gcoarse.back()};
        galerkin_kernel<T, CoarseAcc<T>><<<cg, 128, 0, stream>>>(ca, L);
        HGTRY(finish_level(L, HG_GATHER_LEVEL + (int)gcoarse.size()));
        gcoarse.push_back(L);
        r = L.rows; c = L.cols;
    }
    HGTRY(build_dense(gcoarse.back(), &d_gdense));
    return HG_OK;
}

// Slab mode.  The local grid is [ghost band | owned rows | ghost band]; ghost bands are
// HG_GHOST_ROWS deep MIRRORS of the neighbouring slabs (same flags, same values after a
// refresh).  Every kernel simply runs on the whole local grid: what it computes on ghost rows
// is redundant with the neighbour and valid up to a few rows from the outer edge, so a whole
// smoothing sweep needs no communication and its result on the owned rows is exactly the
// single-GPU one.  Ghost bands are refreshed (send / recv) once per level and leg; from level
// HG_GATHER_LEVEL down the hierarchy is replicated on every rank.  Dot products run over the
// owned tile rows only and are all-reduced.
template <typename T>
int Solver<T>::comm_init(const uint8_t* id, int rank_, int world_, int gt, int gb, const int* starts, int grows) {
    if (world_ < 1 || rank_ < 0 || rank_ >= world_) return fail(HG_ERR_INVALID, "hg_comm_init: bad rank / world");
    rank = rank_; world = world_; ghost_top = gt; ghost_bottom = gb; global_rows = grows;
    if (world == 1) return HG_OK;
    if (op != HG_OP_LAPLACIAN) return fail(HG_ERR_INVALID, "hg_comm_init: the slab mode supports the Laplacian operator only");
    if (!id || !starts) return fail(HG_ERR_INVALID, "hg_comm_init: NULL id / row_starts");
    row_starts.assign(starts, starts + world);
    row_starts.push_back(grows);
    const int want_t = rank == 0 ? 0 : HG_GHOST_ROWS, want_b = rank == world - 1 ? 0 : HG_GHOST_ROWS;
    if (gt != want_t || gb != want_b) return fail(HG_ERR_INVALID, "hg_comm_init: ghost bands must be HG_GHOST_ROWS deep towards every neighbour");
    for (int q = 0; q < world; ++q) {
        if (row_starts[q] % HG_GHOST_ROWS) return fail(HG_ERR_INVALID, "hg_comm_init: slab boundaries must be multiples of HG_GHOST_ROWS");
        if (row_starts[q + 1] - row_starts[q] < 2 * HG_GHOST_ROWS) return fail(HG_ERR_INVALID, "hg_comm_init: a slab needs at least 2 HG_GHOST_ROWS rows");
    }
    if (rows != row_starts[rank + 1] - row_starts[rank] + gt + gb) return fail(HG_ERR_INVALID, "hg_comm_init: local rows do not match the partition");
    HGTRY(nccl_load());
    int dev = 0;
    CU(cudaGetDevice(&dev));
    comm_key = CommKey{world, rank, dev};
    {
        std::lock_guard<std::mutex> lock(g_comms_mu);
        auto it = g_comms.find(comm_key);
        if (it != g_comms.end()) comm = it->second;
    }
    if (!comm) {
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof uid);
        NC(g_nccl.CommInitRank(&comm, world, uid, rank));
        std::lock_guard<std::mutex> lock(g_comms_mu);
        g_comms[comm_key] = comm;
    }
    is_setup = false;
    return HG_OK;
}

template <typename T>
int Solver<T>::allreduce_tmp() {
    NC(g_nccl.AllReduce(d_state->tmp, d_state->tmp, K, ncclDouble, ncclSum, comm, stream));
    return HG_OK;
}

template <typename T>
void Solver<T>::abort_comm() {
    if (comm) {
        std::lock_guard<std::mutex> lock(g_comms_mu);
        g_comms.erase(comm_key);
    }
    if (comm && g_nccl.ok) g_nccl.CommAbort(comm);
    comm = nullptr;
    comm_dead = true;
}

template <typename T>
int Solver<T>::sync_watchful() {
    if (world <= 1 || !comm) { CU(cudaStreamSynchronize(stream)); return HG_OK; }
    static const double limit = getenv("HGRID_COMM_TIMEOUT") ? atof(getenv("HGRID_COMM_TIMEOUT")) : 60.0;
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (long spin = 0;; ++spin) {
        const cudaError_t q = cudaStreamQuery(stream);
        if (q == cudaSuccess) return HG_OK;
        if (q != cudaErrorNotReady) CU(q);
        if ((spin & 1023) == 1023) {
            ncclResult_t ar = ncclSuccess;
            if (g_nccl.CommGetAsyncError(comm, &ar) != ncclSuccess || (ar != ncclSuccess && ar != ncclInProgress))
                return fail(HG_ERR_COMM, std::string("slab mode: asynchronous NCCL error: ") + g_nccl.GetErrorString(ar));
            timespec t1;
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > limit)
                return fail(HG_ERR_COMM, "slab mode: no progress within HGRID_COMM_TIMEOUT seconds (a peer has failed or left the solve)");
        }
    }
}

// the members defined above, for both storage types
template struct Solver<float>;
template struct Solver<double>;

}  // namespace hg
