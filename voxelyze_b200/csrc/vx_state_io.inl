// vx_state_io.inl -- state in and out of a handle: dynamic-state checkpoint files, contact forces, packed link state.
// Included by vx_capi.cu inside its extern "C" block; not a translation unit of its own.

// ---- dynamic-state checkpoint (the reference has none) -----------------
namespace {
struct StateHeader {
    char magic[8]; int32_t abi, lattice, N, L, nx, ny, nz, n_members, gen, have_prev, collisions, header_bytes;       // header_bytes: sizeof(StateHeader) of the writer (layout check)
    float last_prev_dt, prev_dt_host, time_host, ambient; uint64_t topo_hash; DevParams params;
};
struct Chunk { void* p; size_t bytes; };
static uint64_t topo_hash(const vx_sim* s)
{
    uint64_t h = 1469598103934665603ULL;                                   // FNV-1a over the model the arrays belong to
    auto mix = [&](const void* d, size_t n) { const unsigned char* b = (const unsigned char*)d; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; } };
    mix(s->ijk.data(), s->ijk.size() * sizeof(int32_t)); mix(s->vmat_id.data(), s->vmat_id.size() * sizeof(uint16_t));
    mix(s->member.data(), s->member.size() * sizeof(int32_t)); mix(s->vflags.data(), s->vflags.size() * sizeof(uint32_t));
    // ... and everything else the restored arrays are only meaningful with: voxel size, every material parameter and data
    // curve, the externals, gravity and floor (a state resumed under other physics would be silently wrong)
    mix(&s->vox_size, sizeof(double)); mix(&s->grav, sizeof(float));
    const int env[2] = {s->floor_on ? 1 : 0, (int)s->descs.size()}; mix(env, sizeof(env));
    for (size_t i = 0; i < s->descs.size(); i++) {
        vx_material_desc d = s->descs[i]; d.strain = d.stress = nullptr;       // the curve arrays are hashed by content
        mix(&d.model, sizeof(d.model)); mix(&d.youngs_modulus, sizeof(float)); mix(&d.plastic_modulus, sizeof(float)); mix(&d.yield_stress, sizeof(float));
        mix(&d.fail_stress, sizeof(float)); mix(&d.n_points, sizeof(d.n_points)); mix(&d.density, sizeof(float)); mix(&d.poissons_ratio, sizeof(float));
        mix(&d.cte, sizeof(float)); mix(&d.mu_static, sizeof(float)); mix(&d.mu_kinetic, sizeof(float)); mix(&d.zeta_internal, sizeof(float));
        mix(&d.zeta_global, sizeof(float)); mix(&d.zeta_collision, sizeof(float)); mix(d.ext_scale, sizeof(d.ext_scale));
        if (i < s->d_eps.size()) { mix(s->d_eps[i].data(), s->d_eps[i].size() * sizeof(float)); mix(s->d_sig[i].data(), s->d_sig[i].size() * sizeof(float)); }
    }
    mix(s->ext_vox.data(), s->ext_vox.size() * sizeof(int32_t));
    for (const DevExt& e : s->ext_rows) {
        mix(e.nominal, sizeof(e.nominal)); mix(e.translation, sizeof(e.translation)); mix(e.rot_q, sizeof(e.rot_q));
        mix(e.force, sizeof(e.force)); mix(e.moment, sizeof(e.moment)); mix(&e.dof, sizeof(e.dof));
    }
    return h;
}
static std::vector<Chunk> state_chunks(vx_sim* s)
{
    std::vector<Chunk> c;
    const size_t N = s->N, L = s->L;
    if (s->lattice) {
        for (int g = 0; g < 2; g++) {
            c.push_back({s->pose0[g].p, N * sizeof(double4)}); c.push_back({s->pose1[g].p, N * sizeof(double4)});
            c.push_back({s->mom0[g].p, N * sizeof(double4)}); c.push_back({s->mom1[g].p, N * sizeof(double2)});
            c.push_back({s->rec[g].p, N * VX_REC_PARTS * sizeof(double2)});
            if (s->ps[g].p) c.push_back({s->ps[g].p, N * sizeof(float4)});
        }
    } else {
        c.push_back({s->pose0[0].p, N * sizeof(double4)}); c.push_back({s->pose1[0].p, N * sizeof(double4)});
        c.push_back({s->mom0[0].p, N * sizeof(double4)}); c.push_back({s->mom1[0].p, N * sizeof(double2)});
        c.push_back({s->slots.p, N * 36 * sizeof(double)}); c.push_back({s->slot_strain.p, N * 6 * sizeof(float)});
        c.push_back({s->pstrain.p, N * sizeof(float4)});
        c.push_back({s->lstA.p, L * sizeof(double4)}); c.push_back({s->lstB.p, L * sizeof(double4)}); c.push_back({s->lstC.p, L * sizeof(double)});
        c.push_back({s->lstrain.p, L * sizeof(float4)}); c.push_back({s->lmeta.p, L * sizeof(uint32_t)});
    }
    return c;
}
} // namespace

int vx_save_state(vx_sim* s, const char* path)
{
    if (!s || !path || s->call_active) return VX_ERR_ARG;
    { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(s, VX_ERR_ARG, std::string("cannot write ") + path);
    StateHeader h{};
    memcpy(h.magic, "VXB2ST02", 8);
    h.abi = VX_ABI_VERSION; h.lattice = s->lattice; h.N = s->N; h.L = s->L; h.nx = s->nx; h.ny = s->ny; h.nz = s->nz; h.n_members = s->n_members;
    h.gen = s->gen; h.have_prev = s->have_prev; h.collisions = s->collisions; h.header_bytes = (int32_t)sizeof(StateHeader);
    h.last_prev_dt = s->last_prev_dt; h.prev_dt_host = s->prev_dt_host; h.time_host = s->time_host; h.ambient = s->ambient;
    h.topo_hash = topo_hash(s);
    cudaError_t e = cudaMemcpy(&h.params, s->params.p, sizeof(DevParams), cudaMemcpyDeviceToHost);
    bool ok = e == cudaSuccess && fwrite(&h, sizeof(h), 1, fp) == 1;
    std::vector<unsigned char> bounce(64u << 20);
    for (const Chunk& c : state_chunks(s)) {
        for (size_t off = 0; ok && off < c.bytes; off += bounce.size()) {
            const size_t n = std::min(bounce.size(), c.bytes - off);
            ok = cudaMemcpy(bounce.data(), (const unsigned char*)c.p + off, n, cudaMemcpyDeviceToHost) == cudaSuccess && fwrite(bounce.data(), 1, n, fp) == n;
        }
    }
    ok = fclose(fp) == 0 && ok;
    return ok ? VX_OK : fail(s, VX_ERR_CUDA, std::string("writing ") + path + " failed");
}

int vx_load_state(vx_sim* s, const char* path)
{
    if (!s || !path || s->call_active) return VX_ERR_ARG;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(s, VX_ERR_ARG, std::string("cannot read ") + path);
    StateHeader h{};
    bool ok = fread(&h, sizeof(h), 1, fp) == 1 && memcmp(h.magic, "VXB2ST02", 8) == 0 && h.abi == VX_ABI_VERSION && h.header_bytes == (int32_t)sizeof(StateHeader);
    if (ok && (h.lattice != (int)s->lattice || h.N != s->N || h.L != s->L || h.nx != s->nx || h.ny != s->ny || h.nz != s->nz ||
               h.n_members != s->n_members || h.collisions != (int)s->collisions || h.topo_hash != topo_hash(s))) {
        fclose(fp);
        return fail(s, VX_ERR_ARG, "vx_load_state: the file belongs to a different model (voxels, voxel size, materials, externals, environment or options differ)");
    }
    std::vector<unsigned char> bounce(64u << 20);
    for (const Chunk& c : state_chunks(s)) {
        for (size_t off = 0; ok && off < c.bytes; off += bounce.size()) {
            const size_t n = std::min(bounce.size(), c.bytes - off);
            ok = fread(bounce.data(), 1, n, fp) == n && cudaMemcpy((unsigned char*)c.p + off, bounce.data(), n, cudaMemcpyHostToDevice) == cudaSuccess;
        }
    }
    fclose(fp);
    if (!ok) return fail(s, VX_ERR_ARG, std::string("reading ") + path + " failed (truncated or not a state file)");
    h.params.col_stale = 1;                                  // watch lists are rebuilt from the restored positions at the next step
    h.params.div_now = h.params.div_latched = h.params.pending = 0; h.params.div_flag[0] = h.params.div_flag[1] = 0;   // step bookkeeping restarts (vx_step's k_begin does the same)
    s->amb_pending = false; s->last_amb = false; s->n_pairs = 0;
    CK(cudaMemcpy(s->params.p, &h.params, sizeof(DevParams), cudaMemcpyHostToDevice));
    s->gen = h.gen; s->have_prev = h.have_prev != 0; s->last_prev_dt = h.last_prev_dt; s->prev_dt_host = h.prev_dt_host;
    s->time_host = h.time_host; s->ambient = h.ambient; s->col_stale_host = true;
    return VX_OK;
}

int vx_collision_forces(vx_sim* s, int32_t* pairs, float* forces, int cap, int* n_pairs)
{
    if (!s) return VX_ERR_ARG;
    const int P = (s->collisions && s->col_tables) ? s->n_pairs : 0;
    if (n_pairs) *n_pairs = P;
    if ((!pairs && !forces) || P == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    std::vector<int2> raw(P); std::vector<float4> fr(P); std::vector<int> orig(s->n_surf);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(raw.data(), s->c_pairs.p, (size_t)P * sizeof(int2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fr.data(), s->c_pair_force.p, (size_t)P * sizeof(float4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(orig.data(), s->c_surf_orig.p, (size_t)s->n_surf * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> order(P);
    for (int k = 0; k < P; k++) order[k] = k;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        return std::make_pair(orig[raw[a].x], orig[raw[a].y]) < std::make_pair(orig[raw[b].x], orig[raw[b].y]); });
    for (int k = 0; k < P && k < cap; k++) {
        const int q = order[k];
        if (pairs) { pairs[2 * k] = orig[raw[q].x]; pairs[2 * k + 1] = orig[raw[q].y]; }
        if (forces) { forces[3 * k] = fr[q].x; forces[3 * k + 1] = fr[q].y; forces[3 * k + 2] = fr[q].z; }
    }
    return VX_OK;
}

// ---- packed link state (topology edits and layout changes keep the state of surviving links) -------
int vx_download_link_state(vx_sim* s, int first, int count, vx_link_state* dst)
{
    if (!s || !dst || first < 0 || count < 0 || (long long)first + count > s->L) return VX_ERR_ARG;
    if (count == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    std::vector<double> p2(3 * (size_t)count), a1(3 * (size_t)count), a2(3 * (size_t)count);
    std::vector<float> e(count), em(count), eo(count), sg(count); std::vector<uint32_t> fl(count);
    int rc = vx_download(s, VX_F_POS2, first, count, p2.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_ANGLE1V, first, count, a1.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_ANGLE2V, first, count, a2.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_STRAIN, first, count, e.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_MAXSTRAIN, first, count, em.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_STRAINOFFSET, first, count, eo.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_STRESS, first, count, sg.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_LINKFLAGS, first, count, fl.data());
    if (rc != VX_OK) return rc;
    for (int k = 0; k < count; k++) {
        vx_link_state& r = dst[k];
        for (int c = 0; c < 3; c++) { r.pos2[c] = p2[3 * (size_t)k + c]; r.angle1v[c] = a1[3 * (size_t)k + c]; r.angle2v[c] = a2[3 * (size_t)k + c]; }
        r.strain = e[k]; r.max_strain = em[k]; r.strain_offset = eo[k]; r.stress = sg[k]; r.flags = fl[k]; r.reserved = 0;
    }
    return VX_OK;
}

int vx_upload_link_state(vx_sim* s, int first, int count, const vx_link_state* src)
{
    static_assert(sizeof(vx_link_state) == sizeof(LinkStateRec), "vx_link_state layout");
    if (!s || !src || first < 0 || count < 0 || (long long)first + count > s->L || s->call_active) return VX_ERR_ARG;
    if (count == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    const size_t bytes = (size_t)count * sizeof(vx_link_state);
    CK(s->staging.alloc(bytes));
    CK(cudaMemcpyAsync(s->staging.p, src, bytes, cudaMemcpyHostToDevice, s->stream));
    if (!s->lattice) {
        k_scatter_link_state<<<blocks_for(count), TPB, 0, s->stream>>>(s->frame(), s->link_e2i_dev.p, first, count, (const LinkStateRec*)s->staging.p, s->axis_first[1], s->axis_first[2]);
    } else {
        std::vector<int> of((size_t)3 * s->N, -1);              // (axis, owner voxel) -> caller link index
        for (int e = 0; e < s->L; e++) of[(size_t)s->lk_axis[e] * s->N + s->v_e2i[s->lk_vn[e]]] = e;
        DevBuf<int> of_dev;
        CK(of_dev.alloc(of.size()));
        CK(cudaMemcpyAsync(of_dev.p, of.data(), of.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
        const int g = s->gen;
        k_lattice_scatter_link_state<<<blocks_for(s->N), TPB, 0, s->stream>>>(s->pose1[g].p, s->rec[g].p, of_dev.p, s->N,
                                                                              (const LinkStateRec*)s->staging.p, first, count);
        CK(cudaStreamSynchronize(s->stream));
        of_dev.release();
        { int rc = refresh_lattice_ps(s); if (rc != VX_OK) return rc; }       // Poisson: pStrain follows the new link strains
        s->have_prev = false;                                    // link forces are recomputed from the previous generation, which no longer matches
    }
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s->stream));
    return VX_OK;
}

