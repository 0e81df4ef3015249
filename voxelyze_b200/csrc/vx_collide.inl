// vx_collide.inl -- host side of the collision pipeline (kernels: vx_collide.cuh): surface tables, the rebuild chain, the
// conditional graph node, per-call bookkeeping.  Included by vx_capi.cu; not a translation unit of its own.

// ------------------------------------------------------------------------------------------------
// collisions
// Surface voxels, their 5-hop exclusion masks (CVX_Voxel::generateNearby, src/VX_Voxel.cpp:395-418:
// breadth-first over links, depth (int)(2*2.5) = 5) and the device tables.  Topology is static
// between vx_set_voxels calls, so this runs once on the host.
static int build_collision_tables(vx_sim* s)
{
    const int N = s->N;
    std::vector<int> surf_vox, surf_orig, surf_member, slot(std::max(N, 1), -1);
    std::vector<short4> surf_ijk;
    for (int i = 0; i < N; i++) {                       // internal order
        int e = s->v_i2e[i];
        if (s->linkmask[e] == 0x3F) continue;
        if (!s->vflags.empty() && (s->vflags[e] & VF_FILL)) continue;      // the inert cells that fill a box with holes
        slot[i] = (int)surf_vox.size();
        surf_vox.push_back(i); surf_orig.push_back(e); surf_member.push_back(s->member[e]);
        surf_ijk.push_back(make_short4((short)s->ijk[3 * e], (short)s->ijk[3 * e + 1], (short)s->ijk[3 * e + 2], 0));
    }
    const int S = s->n_surf = (int)surf_vox.size();
    std::vector<uint32_t> nearby((size_t)std::max(S, 1) * VX_NEARBY_WORDS, 0u);
    std::vector<int> frontier, next_frontier, visited_list;
    std::vector<char> visited(N, 0);
    for (int k = 0; k < S; k++) {
        const int root = surf_orig[k];
        frontier.assign(1, root); visited_list.assign(1, root); visited[root] = 1;
        for (int depth = 0; depth < 5 && !frontier.empty(); depth++) {
            next_frontier.clear();
            for (int v : frontier)
                for (int d = 0; d < 6; d++) {
                    if (!(s->linkmask[v] & (1u << d))) continue;
                    int o = s->nbr[(size_t)v * 6 + d];
                    if (o < 0 || visited[o]) continue;
                    visited[o] = 1; visited_list.push_back(o); next_frontier.push_back(o);
                }
            frontier.swap(next_frontier);
        }
        uint32_t* mask = &nearby[(size_t)k * VX_NEARBY_WORDS];
        for (int o : visited_list) {
            visited[o] = 0;
            int ox = s->ijk[3 * o] - s->ijk[3 * root], oy = s->ijk[3 * o + 1] - s->ijk[3 * root + 1], oz = s->ijk[3 * o + 2] - s->ijk[3 * root + 2];
            int bit = ((oz + 5) * 11 + (oy + 5)) * 11 + (ox + 5);
            mask[bit >> 5] |= 1u << (bit & 31);
        }
    }
    s->hash_size = 1024; while (s->hash_size < S) s->hash_size <<= 1;
    size_t s1 = std::max(S, 1);
    CK(s->c_surf_vox.alloc(s1)); CK(s->c_surf_orig.alloc(s1)); CK(s->c_surf_member.alloc(s1)); CK(s->c_surf_ijk.alloc(s1));
    CK(s->c_nearby.alloc(s1 * VX_NEARBY_WORDS)); CK(s->c_slot.alloc(std::max(N, 1)));
    CK(s->c_last_watch.alloc(s1)); CK(s->c_cell_count.alloc(s->hash_size)); CK(s->c_cell_start.alloc((size_t)s->hash_size + 1)); CK(s->c_sorted.alloc(s1)); CK(s->c_cell.alloc(s1));
    CK(s->c_counters.alloc(CC_COUNT)); CK(s->c_deg.alloc(s1)); CK(s->c_ref_start.alloc(s1 + 1)); CK(s->c_ref_fill.alloc(s1));
    if (!s->counters_host) CK(cudaMallocHost((void**)&s->counters_host, CC_COUNT * sizeof(int)));
    if (!s->aux_stream) CK(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
    memset(s->counters_host, 0, CC_COUNT * sizeof(int));
    CK(cudaStreamSynchronize(s->stream));
    if (S) {
        CK(cudaMemcpy(s->c_surf_vox.p, surf_vox.data(), (size_t)S * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s->c_surf_orig.p, surf_orig.data(), (size_t)S * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s->c_surf_member.p, surf_member.data(), (size_t)S * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s->c_surf_ijk.p, surf_ijk.data(), (size_t)S * sizeof(short4), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s->c_nearby.p, nearby.data(), nearby.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (N) CK(cudaMemcpy(s->c_slot.p, slot.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(s->c_last_watch.p, 0, s1 * sizeof(float4)));       // new Vec3D<float>() in CVX_Voxel::enableCollisions
    CK(cudaMemset(s->c_ref_start.p, 0, (s1 + 1) * sizeof(int)));
    // the pair list cannot grow while steps are queued: generous to begin with, doubled between calls when half full
    s->col_cap = std::max(s->col_cap, std::max(4096, 16 * S));
    CK(s->c_pairs.alloc(s->col_cap)); CK(s->c_pair_kc.alloc(s->col_cap)); CK(s->c_pair_force.alloc(s->col_cap)); CK(s->c_refs.alloc((size_t)2 * s->col_cap));
    const int init[CC_COUNT] = {1, 0, 0, 0};                         // stale: the first step builds the lists
    CK(cudaMemcpy(s->c_counters.p, init, sizeof(init), cudaMemcpyHostToDevice));
    s->n_pairs = 0; s->col_rebuilds = 0; s->col_tables = true; s->col_stale_host = false;
    s->drop_graph();
    return VX_OK;
}

// CVoxelyze::regenerateCollisions (src/Voxelyze.cpp:725-750): the rebuild chain, every kernel predicated on the stale flag
static void launch_collision_rebuild(vx_sim* s, const Frame& f, const ColFrame& c, cudaStream_t st)
{
    const int S = s->n_surf;
    const int gs = std::min(blocks_for(S), 148 * 16), gh = std::min(blocks_for(std::max(S, s->hash_size)), 148 * 16);
    k_col_clear<<<gh, TPB, 0, st>>>(c);
    k_col_keys<<<gs, TPB, 0, st>>>(f, c);
    k_col_scan<<<1, 1024, 0, st>>>(c, c.cell_count, c.cell_start, s->hash_size);
    k_col_scatter<<<gs, TPB, 0, st>>>(c);
    k_col_pairs<<<gs, TPB, 0, st>>>(f, c);
    k_col_scan<<<1, 1024, 0, st>>>(c, c.deg, s->c_ref_start.p, S);
    k_col_fill<<<std::min(blocks_for(s->col_cap), 148 * 16), TPB, 0, st>>>(c);
    k_col_sort<<<gs, TPB, 0, st>>>(c);
    k_col_done<<<1, 1, 0, st>>>(c);
    s->launches += 9;
}

// CVoxelyze::updateCollisions (src/Voxelyze.cpp:670-710) queued on the handle's stream: stale test, rebuild if stale, contact
// forces.  Nothing here waits for the device.  capturing: the stream is being captured into a step graph; the rebuild chain
// then becomes the body of a conditional IF node (it is not even launched on the steps that keep their lists).
static int enqueue_collision_step(vx_sim* s, bool capturing)
{
    if (s->n_surf == 0) return VX_OK;
    const Frame f = s->frame();
    const ColFrame c = s->col_frame();
    cudaStream_t st = s->stream;
    const int gs = std::min(blocks_for(s->n_surf), 148 * 16);
    k_col_stale<<<gs, TPB, 0, st>>>(f, c); s->launches++;
    bool chained = false;
    if (capturing && s->cond_nodes) {
        // IF node: condition set by k_col_decide, body = the rebuild chain captured on a second stream
        cudaStreamCaptureStatus status; cudaGraph_t graph = nullptr; const cudaGraphNode_t* deps = nullptr; size_t n_deps = 0;
        cudaGraphConditionalHandle handle;
        bool ok = cudaStreamGetCaptureInfo_v2(st, &status, nullptr, &graph, &deps, &n_deps) == cudaSuccess && status == cudaStreamCaptureStatusActive &&
                  cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault) == cudaSuccess;
        if (ok) {
            k_col_decide<<<1, 1, 0, st>>>(handle, c); s->launches++;
            ok = cudaStreamGetCaptureInfo_v2(st, &status, nullptr, &graph, &deps, &n_deps) == cudaSuccess;
        }
        cudaGraphNode_t node = nullptr;
        cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
        if (ok) {
            np.type = cudaGraphNodeTypeConditional;
            np.conditional.handle = handle; np.conditional.type = cudaGraphCondTypeIf; np.conditional.size = 1;
            ok = cudaGraphAddNode(&node, graph, deps, n_deps, &np) == cudaSuccess && np.conditional.phGraph_out && np.conditional.phGraph_out[0];
        }
        if (ok) {
            ok = s->aux_stream != nullptr;               // created by build_collision_tables (not while a capture is open)
            if (ok) ok = cudaStreamBeginCaptureToGraph(s->aux_stream, np.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                const int64_t before = s->launches;
                launch_collision_rebuild(s, f, c, s->aux_stream);
                s->launches = before;                     // body launches are conditional: not counted as launches of the step
                ok = cudaStreamEndCapture(s->aux_stream, nullptr) == cudaSuccess;
            }
            if (ok) ok = cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies) == cudaSuccess;
        }
        if (!ok) { cudaGetLastError(); return fail(s, VX_ERR_CUDA, "conditional graph node for the collision rebuild could not be built"); }
        chained = true;
    }
    if (!chained) launch_collision_rebuild(s, f, c, st);
    k_col_narrow<<<std::min(blocks_for(s->col_cap), 148 * 16), TPB, 0, st>>>(f, c); s->launches++;
    return VX_OK;
}

// start of a stepping call: events the host knows about (reset, new externals, loaded state, ...) raise the device's stale flag
static void collision_call_begin(vx_sim* s)
{
    if (!s->collisions || !s->col_tables || !s->col_stale_host) return;
    k_col_mark_stale<<<1, 1, 0, s->stream>>>(s->c_counters.p); s->launches++;
    s->col_stale_host = false;
}
// end of a stepping call (the stream has been synchronised and counters_host holds the device counters): mirror the
// pair count, grow the pair list for the next call when it is half full, report an overflow
static int collision_call_end(vx_sim* s)
{
    if (!s->collisions || !s->col_tables) return VX_OK;
    const int* c = s->counters_host;
    s->n_pairs = std::min(c[CC_PAIRS], s->col_cap);
    s->col_rebuilds = c[CC_REBUILDS];
    const bool overflow = c[CC_OVERFLOW] != 0;
    if (overflow || 2LL * c[CC_PAIRS] > s->col_cap) {
        s->col_cap = (int)std::min<long long>(std::max(2LL * s->col_cap, 2LL * c[CC_PAIRS]), 1LL << 30);
        s->c_pairs.release(); s->c_pair_kc.release(); s->c_pair_force.release(); s->c_refs.release();
        CK(s->c_pairs.alloc(s->col_cap)); CK(s->c_pair_kc.alloc(s->col_cap)); CK(s->c_pair_force.alloc(s->col_cap)); CK(s->c_refs.alloc((size_t)2 * s->col_cap));
        const int init[CC_COUNT] = {1, 0, 0, c[CC_REBUILDS]};      // the lists are rebuilt into the new arrays by the next step
        CK(cudaMemcpy(s->c_counters.p, init, sizeof(init), cudaMemcpyHostToDevice));
        s->n_pairs = 0;
        s->drop_graph();
    }
    if (overflow) return fail(s, VX_ERR_ALLOC, "the watched-pair list overflowed during this call (its capacity has been doubled): contacts were missed, reload or reset the state");
    return VX_OK;
}
