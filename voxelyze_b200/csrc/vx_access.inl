// vx_access.inl -- state access of the C-ABI: vx_download / vx_upload in caller order, the one-call voxel record, collision
// pair lists, CVoxelyze::stateInfo reductions (src/Voxelyze.cpp:752-800).  Included by vx_capi.cu inside its extern "C" block; not a
// translation unit of its own.

static bool field_info(int field, int& what, int& comps, int& esize, bool& is_link)
{
    switch (field) {
    case VX_F_POS: what = G_POS; comps = 3; esize = 8; is_link = false; return true;
    case VX_F_ORIENT: what = G_ORIENT; comps = 4; esize = 8; is_link = false; return true;
    case VX_F_LINMOM: what = G_LINMOM; comps = 3; esize = 8; is_link = false; return true;
    case VX_F_ANGMOM: what = G_ANGMOM; comps = 3; esize = 8; is_link = false; return true;
    case VX_F_TEMP: what = G_TEMP; comps = 1; esize = 4; is_link = false; return true;
    case VX_F_VOXFLAGS: what = G_VOXFLAGS; comps = 1; esize = 4; is_link = false; return true;
    case VX_F_PSTRAIN: what = G_PSTRAIN; comps = 3; esize = 4; is_link = false; return true;
    case VX_F_FORCE_NEG: what = G_FORCE_NEG; comps = 3; esize = 8; is_link = true; return true;
    case VX_F_FORCE_POS: what = G_FORCE_POS; comps = 3; esize = 8; is_link = true; return true;
    case VX_F_MOMENT_NEG: what = G_MOMENT_NEG; comps = 3; esize = 8; is_link = true; return true;
    case VX_F_MOMENT_POS: what = G_MOMENT_POS; comps = 3; esize = 8; is_link = true; return true;
    case VX_F_POS2: what = G_POS2; comps = 3; esize = 8; is_link = true; return true;
    case VX_F_ANGLE1V: what = G_ANGLE1V; comps = 3; esize = 8; is_link = true; return true;
    case VX_F_ANGLE2V: what = G_ANGLE2V; comps = 3; esize = 8; is_link = true; return true;
    case VX_F_STRAIN: what = G_STRAIN; comps = 1; esize = 4; is_link = true; return true;
    case VX_F_MAXSTRAIN: what = G_MAXSTRAIN; comps = 1; esize = 4; is_link = true; return true;
    case VX_F_STRAINOFFSET: what = G_STRAINOFFSET; comps = 1; esize = 4; is_link = true; return true;
    case VX_F_STRESS: what = G_STRESS; comps = 1; esize = 4; is_link = true; return true;
    case VX_F_LINKFLAGS: what = G_LINKFLAGS; comps = 1; esize = 4; is_link = true; return true;
    }
    return false;
}

// lattice mode: link index -> (owner voxel, axis), built on the first link download
static int ensure_link_refs(vx_sim* s)
{
    if (s->link_owner.p || s->L == 0) return VX_OK;
    std::vector<int> owner(s->L); std::vector<unsigned char> axis(s->L);
    for (int i = 0; i < s->L; i++) { int e = s->l_i2e[i]; owner[i] = s->v_e2i[s->lk_vn[e]]; axis[i] = s->lk_axis[e]; }
    CK(s->link_owner.alloc(s->L)); CK(s->link_axis_dev.alloc(s->L));
    CK(cudaMemcpy(s->link_owner.p, owner.data(), (size_t)s->L * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->link_axis_dev.p, axis.data(), (size_t)s->L, cudaMemcpyHostToDevice));
    return VX_OK;
}

int vx_download(vx_sim* s, int field, int first, int count, void* dst)
{
    int what, comps, esize; bool is_link;
    if (!s || !dst || first < 0 || count < 0 || !field_info(field, what, comps, esize, is_link)) return VX_ERR_ARG;
    if ((long long)first + count > (is_link ? s->L : s->N_user)) return VX_ERR_ARG;
    if (count == 0) return VX_OK;
    { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    CK(cudaSetDevice(s->device));
    size_t bytes = (size_t)count * comps * esize;
    CK(cudaStreamSynchronize(s->stream));
    CK(s->staging.alloc(bytes));
    if (s->lattice && is_link) {
        int rc = ensure_link_refs(s);
        if (rc != VX_OK) return rc;
        LatLinkRef ref{s->link_owner.p, s->link_axis_dev.p};
        k_lattice_gather_links<<<blocks_for(count), TPB, 0, s->stream>>>(s->lat_frame(s->gen), prev_frame(s), s->have_prev ? 1 : 0,
                                                                         s->last_prev_dt, what, s->link_e2i_dev.p, ref, first, count, s->staging.p);
    } else {
        k_gather<<<blocks_for(count), TPB, 0, s->stream>>>(s->frame(), what, is_link ? s->link_e2i_dev.p : s->vox_e2i_dev.p, first, count,
                                                           s->staging.p, s->axis_first[1], s->axis_first[2]);
    }
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, s->staging.p, bytes, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return VX_OK;
}

int vx_download_voxel_state(vx_sim* s, int first, int count, vx_voxel_state* dst)
{
    static_assert(sizeof(vx_voxel_state) == sizeof(VoxelStateRec), "vx_voxel_state layout");
    if (!s || !dst || first < 0 || count < 0 || first + (long long)count > s->N_user) return VX_ERR_ARG;
    if (count == 0) return VX_OK;
    { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    CK(cudaSetDevice(s->device));
    if (count <= 32) {                                  // the common case (a caller polling a few voxels per step): the kernel writes
        if (!s->probe_host) CK(cudaHostAlloc((void**)&s->probe_host, 32 * sizeof(VoxelStateRec), cudaHostAllocMapped));     // straight into mapped pinned memory
        VoxelStateRec* dev = nullptr;
        CK(cudaHostGetDevicePointer((void**)&dev, s->probe_host, 0));
        k_gather_voxel_state<<<1, 32, 0, s->stream>>>(s->frame(), s->vox_e2i_dev.p, first, count, dev); s->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(s->stream));
        memcpy(dst, s->probe_host, (size_t)count * sizeof(VoxelStateRec));
        return VX_OK;
    }
    const size_t bytes = (size_t)count * sizeof(VoxelStateRec);
    CK(cudaStreamSynchronize(s->stream));
    CK(s->staging.alloc(bytes));
    k_gather_voxel_state<<<blocks_for(count), TPB, 0, s->stream>>>(s->frame(), s->vox_e2i_dev.p, first, count, (VoxelStateRec*)s->staging.p); s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, s->staging.p, bytes, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return VX_OK;
}

int vx_upload(vx_sim* s, int field, int first, int count, const void* src)
{
    int what, comps, esize; bool is_link;
    if (!s || !src || first < 0 || count < 0 || !field_info(field, what, comps, esize, is_link)) return VX_ERR_ARG;
    if (is_link) return fail(s, VX_ERR_UNSUPPORTED, "only voxel state can be uploaded");
    if (what == G_PSTRAIN && !(s->lattice && s->any_poisson)) return fail(s, VX_ERR_UNSUPPORTED, "Poisson strains are uploaded into the ghost voxels of a z-slab on the fused layout only");
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_upload inside vx_step_begin .. vx_step_end");
    if ((long long)first + count > s->N_user) return VX_ERR_ARG;
    if (count == 0) return VX_OK;
    { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    CK(cudaSetDevice(s->device));
    size_t bytes = (size_t)count * comps * esize;
    CK(cudaStreamSynchronize(s->stream));
    CK(s->staging.alloc(bytes));
    CK(cudaMemcpyAsync(s->staging.p, src, bytes, cudaMemcpyHostToDevice, s->stream));
    k_scatter<<<blocks_for(count), TPB, 0, s->stream>>>(s->frame(), what, s->vox_e2i_dev.p, first, count, s->staging.p);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s->stream));
    return VX_OK;
}

int vx_collision_pairs(vx_sim* s, int32_t* pairs, int cap, int* n_pairs)
{
    if (!s) return VX_ERR_ARG;
    const int P = (s->collisions && s->col_tables) ? s->n_pairs : 0;
    if (n_pairs) *n_pairs = P;
    if (!pairs || P == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    std::vector<int2> raw(P); std::vector<int> orig(s->n_surf);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(raw.data(), s->c_pairs.p, (size_t)P * sizeof(int2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(orig.data(), s->c_surf_orig.p, (size_t)s->n_surf * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<std::pair<int, int>> out(P);
    for (int k = 0; k < P; k++) out[k] = {orig[raw[k].x], orig[raw[k].y]};
    std::sort(out.begin(), out.end());                              // creation order of the reference: i ascending, then j
    for (int k = 0; k < P && k < cap; k++) { pairs[2 * k] = out[k].first; pairs[2 * k + 1] = out[k].second; }
    return VX_OK;
}
int vx_collision_stats(vx_sim* s, int* n_pairs, int* n_rebuilds)
{
    if (!s) return VX_ERR_ARG;
    const bool on = s->collisions && s->col_tables;
    if (n_pairs) *n_pairs = on ? s->n_pairs : 0;
    if (n_rebuilds) *n_rebuilds = on ? s->col_rebuilds : 0;
    return VX_OK;
}
// fills `dst` (device) with one link field for all links, caller order
static int gather_link_field(vx_sim* s, int what, void* dst)
{
    if (s->lattice) {
        int rc = ensure_link_refs(s);
        if (rc != VX_OK) return rc;
        LatLinkRef ref{s->link_owner.p, s->link_axis_dev.p};
        k_lattice_gather_links<<<blocks_for(s->L), TPB, 0, s->stream>>>(s->lat_frame(s->gen), prev_frame(s), s->have_prev ? 1 : 0,
                                                                        s->last_prev_dt, what, s->link_e2i_dev.p, ref, 0, s->L, dst);
    } else {
        k_gather<<<blocks_for(s->L), TPB, 0, s->stream>>>(s->frame(), what, s->link_e2i_dev.p, 0, s->L, dst, s->axis_first[1], s->axis_first[2]);
    }
    s->launches++;
    return VX_OK;
}

// halo copies among the caller's voxels (z-slab handles): how many
static int user_ghosts(const vx_sim* s)
{
    int n = 0;
    if (!s->vflags.empty()) for (int v = 0; v < s->N_user; v++) n += (s->vflags[v] & VX_VF_GHOST) ? 1 : 0;
    return n;
}

// per voxel (caller order) the caller index of its link in each of the six directions, the strain ratio of every link and
// {E, nu} of every voxel: shared by the pressure reduction and the surface mesh
static int ensure_vlinks(vx_sim* s)
{
    if (s->si_pressure_ok) return VX_OK;
    const size_t nu = (size_t)s->N_user;
    std::vector<int> vl(6 * std::max<size_t>(nu, 1), -1); std::vector<float> ratio(std::max(s->L, 1)); std::vector<float2> en(std::max<size_t>(nu, 1));
    for (int l = 0; l < s->L; l++) {
        vl[(size_t)(2 * s->lk_axis[l]) * nu + s->lk_vn[l]] = l;            // +axis slot of the negative-end voxel
        vl[(size_t)(2 * s->lk_axis[l] + 1) * nu + s->lk_vp[l]] = l;        // -axis slot of the positive-end voxel
        ratio[l] = s->mats[s->vmat_id[s->lk_vp[l]]].E / s->mats[s->vmat_id[s->lk_vn[l]]].E;
    }
    for (size_t v = 0; v < nu; v++) en[v] = make_float2(s->mats[s->vmat_id[v]].E, s->mats[s->vmat_id[v]].nu);
    CK(s->si_vlinks.alloc(vl.size())); CK(s->si_ratio.alloc(ratio.size())); CK(s->si_en.alloc(en.size()));
    CK(cudaMemcpy(s->si_vlinks.p, vl.data(), vl.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->si_ratio.p, ratio.data(), ratio.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->si_en.p, en.data(), en.size() * sizeof(float2), cudaMemcpyHostToDevice));
    s->si_skip.release();
    if (user_ghosts(s)) {                                    // z-slab: halo copies are left out of the reductions
        std::vector<unsigned char> sk(nu, 0);
        for (size_t v = 0; v < nu; v++) sk[v] = (s->vflags[v] & VX_VF_GHOST) ? 1 : 0;
        CK(s->si_skip.alloc(nu));
        CK(cudaMemcpy(s->si_skip.p, sk.data(), nu, cudaMemcpyHostToDevice));
    }
    s->si_pressure_ok = true;
    return VX_OK;
}

static int state_info_impl(vx_sim* s, int info, int type, float* out, float* vals);
int vx_state_info(vx_sim* s, int info, int type, float* out) { return state_info_impl(s, info, type, out, nullptr); }

// vals (device, optional): the value of every element -- voxels in internal order (pressure: caller order), links in caller order
static int state_info_impl(vx_sim* s, int info, int type, float* out, float* vals)
{
    if (!s || !out || info < 0 || info > SI_MASS || type < 0 || type > SI_AVERAGE) return VX_ERR_ARG;
    *out = 0.0f;
    const bool link_info = info == SI_STRAIN_ENERGY || info == SI_ENG_STRESS || info == SI_ENG_STRAIN;
    const int count = link_info ? s->L : s->N_user - user_ghosts(s);       // fill cells of a box with holes and halo copies are not voxels
    if (count == 0) return VX_OK;                                  // src/Voxelyze.cpp:759,777
    { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    CK(cudaSetDevice(s->device));
    CK(s->si_minmax.alloc(2)); CK(s->si_sum.alloc(1));
    const float init[2] = {3.402823466e38f, -3.402823466e38f};
    CK(cudaMemcpyAsync(s->si_minmax.p, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemsetAsync(s->si_sum.p, 0, sizeof(double), s->stream));
    const int grid = std::min(blocks_for(count, 256), 148 * 8);
    if (info == SI_PRESSURE) {
        { int rc = ensure_vlinks(s); if (rc != VX_OK) return rc; }
        CK(s->si_buf.alloc((size_t)std::max(s->L, 1) * sizeof(float)));
        if (s->L) { int rc = gather_link_field(s, G_STRAIN, s->si_buf.p); if (rc != VX_OK) return rc; }
        k_state_pressure<<<grid, 256, 0, s->stream>>>(s->N_user, s->si_vlinks.p, (const float*)s->si_buf.p, s->si_ratio.p, s->si_en.p,
                                                      s->si_minmax.p, s->si_minmax.p + 1, s->si_sum.p, vals, s->si_skip.p);
        s->launches++;
    } else if (!link_info) {
        if (info == SI_DISPLACEMENT && !s->si_nominal_ok) {
            std::vector<double4> nom(s->N);
            for (int i = 0; i < s->N; i++) { int e = s->v_i2e[i]; nom[i] = make_double4(s->ijk[3 * e] * s->vox_size, s->ijk[3 * e + 1] * s->vox_size, s->ijk[3 * e + 2] * s->vox_size, 0.0); }
            CK(s->si_nominal.alloc(s->N));
            CK(cudaMemcpy(s->si_nominal.p, nom.data(), (size_t)s->N * sizeof(double4), cudaMemcpyHostToDevice));
            s->si_nominal_ok = true;
        }
        k_state_voxels<<<grid, 256, 0, s->stream>>>(s->frame(), info, s->si_nominal.p, s->si_minmax.p, s->si_minmax.p + 1, s->si_sum.p, vals);
        s->launches++;
    } else if (info != SI_STRAIN_ENERGY) {
        CK(s->si_buf.alloc((size_t)s->L * sizeof(float)));
        int rc = gather_link_field(s, info == SI_ENG_STRESS ? G_STRESS : G_STRAIN, s->si_buf.p);
        if (rc != VX_OK) return rc;
        k_state_links<<<grid, 256, 0, s->stream>>>(s->L, (const float*)s->si_buf.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                   s->si_minmax.p, s->si_minmax.p + 1, s->si_sum.p, vals);
        s->launches++;
    } else {
        if (!s->si_consts_ok) {                                    // a1, a2, b3 of every link's material (caller order)
            std::vector<float> c(3 * (size_t)s->L);
            for (int l = 0; l < s->L; l++) {
                vxm::BeamConsts k = vxm::beam_consts(s->lmats[link_material(s, s->vmat_id[s->lk_vn[l]], s->vmat_id[s->lk_vp[l]])].mat, s->vox_size);
                c[l] = k.a1; c[(size_t)s->L + l] = k.a2; c[2 * (size_t)s->L + l] = k.b3;
            }
            CK(s->si_consts.alloc(c.size()));
            CK(cudaMemcpy(s->si_consts.p, c.data(), c.size() * sizeof(float), cudaMemcpyHostToDevice));
            s->si_consts_ok = true;
        }
        const size_t stride = (size_t)s->L * 3 * sizeof(double);
        CK(s->si_buf.alloc(3 * stride));
        double* fneg = (double*)s->si_buf.p; double* mneg = (double*)(s->si_buf.p + stride); double* mpos = (double*)(s->si_buf.p + 2 * stride);
        int rc = gather_link_field(s, G_FORCE_NEG, fneg);
        if (rc == VX_OK) rc = gather_link_field(s, G_MOMENT_NEG, mneg);
        if (rc == VX_OK) rc = gather_link_field(s, G_MOMENT_POS, mpos);
        if (rc != VX_OK) return rc;
        k_state_links<<<grid, 256, 0, s->stream>>>(s->L, nullptr, fneg, mneg, mpos, s->si_consts.p, s->si_consts.p + s->L, s->si_consts.p + 2 * (size_t)s->L,
                                                   s->si_minmax.p, s->si_minmax.p + 1, s->si_sum.p, vals);
        s->launches++;
    }
    CK(cudaGetLastError());
    float mm[2]; double sum;
    CK(cudaMemcpyAsync(mm, s->si_minmax.p, sizeof(mm), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(&sum, s->si_sum.p, sizeof(sum), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    switch (type) {
    case SI_MIN: *out = mm[0]; break;
    case SI_MAX: *out = mm[1]; break;
    case SI_TOTAL: *out = (float)sum; break;
    default: *out = (float)sum / count; break;
    }
    return VX_OK;
}
