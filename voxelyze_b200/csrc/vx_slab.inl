// vx_slab.inl -- z-slab support of the C-ABI: the peer-memory halo (CUDA IPC mappings of the neighbours' ghost layers,
// arrival counters, the step loop that overlaps the push with the interior part).  Included by vx_capi.cu inside
// its extern "C" block; not a translation unit of its own.

// ---- peer-memory halo ---------------------------------------------------------------------------
struct PeerDescWire {                       // what vx_peer_export writes into vx_peer_desc::bytes
    uint64_t magic; int64_t pid; int32_t device, side; uint64_t first, count;
    cudaIpcMemHandle_t mem[2]; cudaIpcMemHandle_t flag;      // the pose allocation of generation 0 and 1 (pose0, then pose1 at +pose1_off); flag array
    uint64_t raw[2]; uint64_t raw_flag;                      // same-process peers use the addresses directly
    uint64_t pose1_off;                                      // in double4 elements
    int32_t has_ps, pad;                                     // Poisson models: the two pStrain generations travel too
    cudaIpcMemHandle_t mem_ps[2]; uint64_t raw_ps[2];
};
static_assert(sizeof(PeerDescWire) <= VX_PEER_DESC_BYTES, "vx_peer_desc too small");

static int plane_range(vx_sim* s, int iz, size_t& first, size_t& count)
{
    int64_t key = (int64_t)(iz + 32768);
    auto lo = std::lower_bound(s->sort_key.begin(), s->sort_key.end(), key);
    auto hi = std::upper_bound(s->sort_key.begin(), s->sort_key.end(), key);
    first = lo - s->sort_key.begin(); count = hi - lo;
    return count ? VX_OK : VX_ERR_ARG;
}

static int ensure_peer_state(vx_sim* s)
{
    CK(cudaSetDevice(s->device));
    if (!s->peer_flags.p) { CK(s->peer_flags.alloc(4)); CK(cudaMemset(s->peer_flags.p, 0, 4 * sizeof(int))); }
    if (!s->comm_stream) {             // highest priority: its few blocks must not queue behind the interior part's
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, greatest));
    }
    if (!s->ev_boundary) CK(cudaEventCreateWithFlags(&s->ev_boundary, cudaEventDisableTiming));
    if (!s->ev_comm) CK(cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
    return VX_OK;
}

int vx_peer_export(vx_sim* s, int ghost_iz, int from_above, vx_peer_desc* out)
{
    if (!s || !out || !s->lattice || s->n_members != 1) return VX_ERR_ARG;
    int rc = ensure_peer_state(s); if (rc != VX_OK) return rc;
    size_t first, count;
    if (plane_range(s, ghost_iz, first, count) != VX_OK) return fail(s, VX_ERR_ARG, "vx_peer_export: empty layer");
    PeerDescWire w{}; w.magic = 0x56585045455231ULL; w.pid = (int64_t)getpid(); w.device = s->device; w.side = from_above ? 1 : 0;
    w.first = first; w.count = count;
    double4* base[2] = {s->pose0[0].p, s->pose0[1].p};
    for (int k = 0; k < 2; k++) { CK(cudaIpcGetMemHandle(&w.mem[k], base[k])); w.raw[k] = (uint64_t)(uintptr_t)base[k]; }
    w.pose1_off = (uint64_t)(s->pose1[0].p - s->pose0[0].p);
    if ((uint64_t)(s->pose1[1].p - s->pose0[1].p) != w.pose1_off) return fail(s, VX_ERR_CUDA, "vx_peer_export: generations laid out differently");
    CK(cudaIpcGetMemHandle(&w.flag, s->peer_flags.p)); w.raw_flag = (uint64_t)(uintptr_t)s->peer_flags.p;
    w.has_ps = s->any_poisson && s->ps[0].p && s->ps[1].p ? 1 : 0;
    for (int k = 0; k < 2 && w.has_ps; k++) { CK(cudaIpcGetMemHandle(&w.mem_ps[k], s->ps[k].p)); w.raw_ps[k] = (uint64_t)(uintptr_t)s->ps[k].p; }
    memset(out->bytes, 0, VX_PEER_DESC_BYTES); memcpy(out->bytes, &w, sizeof(w));
    s->expect_side[w.side] = true;                                   // a neighbour will write here
    return VX_OK;
}

int vx_peer_attach(vx_sim* s, int send_iz, const vx_peer_desc* peer_ghost)
{
    if (!s || !peer_ghost || !s->lattice || s->n_members != 1) return VX_ERR_ARG;
    int rc = ensure_peer_state(s); if (rc != VX_OK) return rc;
    PeerDescWire w; memcpy(&w, peer_ghost->bytes, sizeof(w));
    if (w.magic != 0x56585045455231ULL) return fail(s, VX_ERR_ARG, "vx_peer_attach: not a peer descriptor");
    vx_sim::PeerLink pl;
    if (plane_range(s, send_iz, pl.src_first, pl.count) != VX_OK || pl.count != w.count) return fail(s, VX_ERR_ARG, "vx_peer_attach: layer size mismatch");
    if ((w.has_ps != 0) != (s->any_poisson && s->ps[0].p)) return fail(s, VX_ERR_ARG, "vx_peer_attach: only one side carries Poisson strains");
    void* base[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (w.pid == (int64_t)getpid()) {                                 // same process (tests, vx_slabbed): plain addresses
        for (int k = 0; k < 2; k++) base[k] = (void*)(uintptr_t)w.raw[k];
        base[2] = (void*)(uintptr_t)w.raw_flag;
        for (int k = 0; k < 2 && w.has_ps; k++) base[3 + k] = (void*)(uintptr_t)w.raw_ps[k];
        if (w.device != s->device) { int can = 0; cudaDeviceCanAccessPeer(&can, s->device, w.device); if (!can) return fail(s, VX_ERR_UNSUPPORTED, "no peer access"); cudaError_t e = cudaDeviceEnablePeerAccess(w.device, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(s, e, "cudaDeviceEnablePeerAccess"); cudaGetLastError(); }
    } else {
        for (int k = 0; k < 2; k++) {
            cudaError_t e = cudaIpcOpenMemHandle(&base[k], w.mem[k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { cudaGetLastError(); return fail(s, VX_ERR_UNSUPPORTED, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
            pl.opened[k] = base[k];
        }
        cudaError_t e = cudaIpcOpenMemHandle(&base[2], w.flag, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(s, VX_ERR_UNSUPPORTED, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
        pl.opened[2] = base[2];
        for (int k = 0; k < 2 && w.has_ps; k++) {
            e = cudaIpcOpenMemHandle(&base[3 + k], w.mem_ps[k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { cudaGetLastError(); return fail(s, VX_ERR_UNSUPPORTED, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
            pl.opened[3 + k] = base[3 + k];
        }
    }
    if (w.has_ps) { pl.dst_ps[0] = (float4*)base[3] + w.first; pl.dst_ps[1] = (float4*)base[4] + w.first; }
    pl.dst0[0] = (double4*)base[0] + w.first; pl.dst0[1] = (double4*)base[1] + w.first;
    pl.dst1[0] = (double4*)base[0] + w.pose1_off + w.first; pl.dst1[1] = (double4*)base[1] + w.pose1_off + w.first;
    pl.dst_flag = (int*)base[2] + w.side;
    s->peers.push_back(pl);
    return VX_OK;
}

int vx_peer_detach(vx_sim* s)
{
    if (!s) return VX_ERR_ARG;
    cudaSetDevice(s->device);
    if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
    for (auto& pl : s->peers) for (void* q : pl.opened) if (q) cudaIpcCloseMemHandle(q);
    s->peers.clear(); s->expect_side[0] = s->expect_side[1] = false;
    // exchange counting restarts with the next attachment: the slabs that meet then need not have the same history (a model
    // that is cut again may use a slab that sat idle before)
    s->xseq = 0;
    if (s->peer_flags.p) cudaMemset(s->peer_flags.p, 0, 4 * sizeof(int));
    return VX_OK;
}

// queue: wait for every exchange so far (compute stream)
static void peer_wait(vx_sim* s, cudaStream_t st)
{
    if (s->xseq == 0 || (!s->expect_side[0] && !s->expect_side[1])) return;
    static long long limit = 0;                            // VX_PEER_TIMEOUT_S (default 30 s) at ~2 GHz
    if (limit == 0) { const char* e = getenv("VX_PEER_TIMEOUT_S"); double sec = e ? atof(e) : 30.0; limit = (long long)(std::max(sec, 0.1) * 2.0e9); }
    k_peer_wait<<<1, 1, 0, st>>>(s->peer_flags.p, s->expect_side[0] ? s->xseq : 0, s->expect_side[1] ? s->xseq : 0, s->peer_flags.p + 2, limit);
    s->launches++;
}
// queue on the comm stream: ship generation g of my boundary layers (unless the step kernel already stored
// them into the neighbours' ghost layers itself) and signal
static void peer_push(vx_sim* s, int g, bool already_stored)
{
    s->xseq++;
    for (auto& pl : s->peers) {
        if (!already_stored) {
            k_halo_push<<<blocks_for((long long)pl.count), TPB, 0, s->comm_stream>>>(s->pose0[g].p + pl.src_first, s->pose1[g].p + pl.src_first,
                                                                                     pl.dst0[g], pl.dst1[g], (int)pl.count);
            s->launches++;
            if (pl.dst_ps[g]) cudaMemcpyAsync(pl.dst_ps[g], s->ps[g].p + pl.src_first, pl.count * sizeof(float4), cudaMemcpyDefault, s->comm_stream);
        }
        k_peer_signal<<<1, 1, 0, s->comm_stream>>>(pl.dst_flag, s->xseq);
        s->launches++;
    }
}
static int peer_check(vx_sim* s)
{
    int t = 0;
    CK(cudaMemcpy(&t, s->peer_flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost));
    return t ? fail(s, VX_ERR_CUDA, "peer halo: a neighbouring slab did not deliver in time") : VX_OK;
}

// Poisson strains travel with the poses: peers attached while the model had no Poisson material do not know where to put them
static int peers_carry_ps(vx_sim* s)
{
    if (s->any_poisson) for (auto& pl : s->peers) if (!pl.dst_ps[0] || !pl.dst_ps[1])
        return fail(s, VX_ERR_ARG, "Poisson's ratio was switched on after vx_peer_attach: detach, export and attach again");
    return VX_OK;
}

int vx_slab_exchange(vx_sim* s)
{
    if (!s || !s->lattice || s->call_active) return VX_ERR_ARG;
    int rc = ensure_peer_state(s); if (rc != VX_OK) return rc;
    rc = peers_carry_ps(s); if (rc != VX_OK) return rc;
    rc = flush_ambient(s); if (rc != VX_OK) return rc;
    CK(cudaEventRecord(s->ev_boundary, s->stream));
    CK(cudaStreamWaitEvent(s->comm_stream, s->ev_boundary, 0));
    peer_push(s, s->gen, false);
    CK(cudaStreamSynchronize(s->comm_stream));       // delivered; the neighbours' deliveries are awaited by the next vx_slab_step
    return VX_OK;
}

int vx_slab_step_begin(vx_sim* s, float dt, int n_steps)
{
    if (!s || n_steps < 1) return VX_ERR_ARG;
    int rc = ensure_peer_state(s); if (rc != VX_OK) return rc;
    rc = peers_carry_ps(s); if (rc != VX_OK) return rc;
    NvtxRange nvtx("vx_slab_step_begin");
    rc = vx_step_begin(s, dt); if (rc != VX_OK) return rc;
    auto abandon = [&](int code) { s->call_active = false; s->call_half = false; s->push_in_kernel = false; return code; };   // leave no call open behind an error
    for (int k = 0; k < n_steps; k++) {
        peer_wait(s, s->stream);                            // the boundary part reads the ghosts of the previous exchange
        s->push_in_kernel = s->peers.size() <= 2;          // the boundary kernels store into the neighbours' ghost layers themselves
        const bool fused = s->push_in_kernel;
        rc = vx_step_enqueue(s, VX_PART_Z_BOUNDARY);
        s->push_in_kernel = false;
        if (rc != VX_OK) return abandon(rc);
        if (cudaEventRecord(s->ev_boundary, s->stream) != cudaSuccess) return abandon(cuda_fail(s, cudaGetLastError(), "cudaEventRecord"));
        rc = vx_step_enqueue(s, VX_PART_Z_INTERIOR); if (rc != VX_OK) return abandon(rc);
        if (cudaStreamWaitEvent(s->comm_stream, s->ev_boundary, 0) != cudaSuccess) return abandon(cuda_fail(s, cudaGetLastError(), "cudaStreamWaitEvent"));
        peer_push(s, s->newest_gen(), fused);
    }
    if (cudaEventRecord(s->ev_comm, s->comm_stream) != cudaSuccess || cudaStreamWaitEvent(s->stream, s->ev_comm, 0) != cudaSuccess)
        return abandon(cuda_fail(s, cudaGetLastError(), "cudaEventRecord"));
    return VX_OK;
}

int vx_slab_step_finish(vx_sim* s, int* diverged_step)
{
    if (!s || !s->call_active) return VX_ERR_ARG;
    NvtxRange nvtx("vx_slab_step_finish");
    CK(cudaSetDevice(s->device));
    int rc = vx_step_end(s, diverged_step);
    if (rc != VX_OK && rc != VX_DIVERGED) return rc;
    int rc2 = peer_check(s);
    return rc2 != VX_OK ? rc2 : rc;
}

int vx_slab_step(vx_sim* s, float dt, int n_steps, int* diverged_step)
{
    if (!s || n_steps < 0) return VX_ERR_ARG;
    if (n_steps == 0) return VX_OK;
    NvtxRange nvtx("vx_slab_step");
    int rc = vx_slab_step_begin(s, dt, n_steps); if (rc != VX_OK) return rc;
    return vx_slab_step_finish(s, diverged_step);
}
