// vx_mesh.inl -- surface mesh entry points of the C-ABI: host-built topology in the reference's
// numbering (CVX_MeshRender::generateMesh, src/VX_MeshRender.cpp:49-145), device-side vertex / normal / colour update
// (vx_mesh.cuh), buffers that stay in HBM.  Included by vx_capi.cu inside its extern "C" block.

int vx_mesh_set_material_colors(vx_sim* s, int n, const unsigned char* rgba)
{
    if (!s || n < 0 || (n && !rgba)) return VX_ERR_ARG;
    s->mesh.rgb_host.assign(3 * (size_t)std::max(n, 1), 1.0f);
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) s->mesh.rgb_host[3 * i + k] = ((float)rgba[4 * i + k]) / 255.0f;   // src/VX_MeshRender.cpp:193-195
    if (s->mesh.built) {
        CK(cudaSetDevice(s->device));
        CK(s->mesh.mat_rgb.alloc(s->mesh.rgb_host.size()));
        CK(cudaMemcpy(s->mesh.mat_rgb.p, s->mesh.rgb_host.data(), s->mesh.rgb_host.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return VX_OK;
}

int vx_mesh_build(vx_sim* s, int* n_vertices, int* n_quads)
{
    if (!s) return VX_ERR_ARG;
    if (s->n_members != 1) return fail(s, VX_ERR_UNSUPPORTED, "the surface mesh is built per simulation, not for ensembles");
    CK(cudaSetDevice(s->device));
    const int n = s->N_user;
    vx_sim::Mesh& m = s->mesh;
    // clockwise corner table of the six faces, CVX_Voxel::voxelCorner codes (bit 2: +x, bit 1: +y, bit 0: +z)
    static const int cw[6][4] = {{4, 6, 7, 5}, {0, 1, 3, 2}, {2, 3, 7, 6}, {0, 4, 5, 1}, {1, 5, 7, 3}, {0, 2, 6, 4}};
    int lo[3] = {0, 0, 0}, hi[3] = {-1, -1, -1};
    for (int k = 0; k < n; k++)
        for (int a = 0; a < 3; a++) {
            const int c = s->ijk[3 * k + a];
            if (k == 0) { lo[a] = hi[a] = c; } else { lo[a] = std::min(lo[a], c); hi[a] = std::max(hi[a], c); }
        }
    const long long ex = n ? hi[0] - lo[0] + 2 : 0, ey = n ? hi[1] - lo[1] + 2 : 0, ez = n ? hi[2] - lo[2] + 2 : 0;    // vertex lattice: one more than the voxels
    if (ex * ey * ez > (1LL << 31)) return fail(s, VX_ERR_ALLOC, "vertex lattice too large");
    std::vector<int> vmap((size_t)(ex * ey * ez), -1), vox_at((size_t)(ex * ey * ez), -1);
    auto cell = [&](int x, int y, int z) -> long long {
        if (x < lo[0] || x > hi[0] + 1 || y < lo[1] || y > hi[1] + 1 || z < lo[2] || z > hi[2] + 1) return -1;
        return ((long long)(z - lo[2]) * ey + (y - lo[1])) * ex + (x - lo[0]);
    };
    for (int k = 0; k < n; k++) vox_at[(size_t)cell(s->ijk[3 * k], s->ijk[3 * k + 1], s->ijk[3 * k + 2])] = k;
    std::vector<int> quads, quad_vox; std::vector<long long> vert_cell;
    for (int k = 0; k < n; k++) {                                   // voxelsList order, then direction, then corner: the reference's numbering
        const int x = s->ijk[3 * k], y = s->ijk[3 * k + 1], z = s->ijk[3 * k + 2];
        for (int d = 0; d < 6; d++) {
            if (s->linkmask[k] & (1u << d)) continue;               // adjacentVoxel(d): a face is exposed where no link leaves
            for (int j = 0; j < 4; j++) {
                const int c = cw[d][j];
                const long long vc = cell(x + ((c >> 2) & 1), y + ((c >> 1) & 1), z + (c & 1));
                int& ind = vmap[(size_t)vc];
                if (ind == -1) { ind = (int)vert_cell.size(); vert_cell.push_back(vc); }
                quads.push_back(ind);
            }
            quad_vox.push_back(k);
        }
    }
    const int nv = (int)vert_cell.size(), nq = (int)quad_vox.size();
    std::vector<int> vert_vox(8 * (size_t)std::max(nv, 1), -1);
    for (int i = 0; i < nv; i++) {                                  // the (up to) eight voxels around a vertex see it as their corner j
        long long r = vert_cell[i];
        const int x = (int)(r % ex) + lo[0]; r /= ex; const int y = (int)(r % ey) + lo[1]; const int z = (int)(r / ey) + lo[2];
        for (int j = 0; j < 8; j++) {
            const long long c = cell(x - ((j >> 2) & 1), y - ((j >> 1) & 1), z - (j & 1));
            if (c >= 0) vert_vox[8 * (size_t)i + j] = vox_at[(size_t)c];
        }
    }
    std::vector<float> ef(std::max(s->L, 1), -1.0f), eyld(std::max(s->L, 1), -1.0f);
    for (int l = 0; l < s->L; l++) {
        const vxm::Material& lm = s->lmats[link_material(s, s->vmat_id[s->lk_vn[l]], s->vmat_id[s->lk_vp[l]])].mat;
        ef[l] = lm.eps_fail; eyld[l] = lm.eps_yield;
    }
    if (m.rgb_host.size() < 3 * s->mats.size()) m.rgb_host.resize(3 * std::max<size_t>(s->mats.size(), 1), 1.0f);
    { int rc = ensure_vlinks(s); if (rc != VX_OK) return rc; }
    CK(m.vert_vox.alloc(vert_vox.size())); CK(m.quads.alloc(std::max<size_t>(quads.size(), 1))); CK(m.quad_vox.alloc(std::max(nq, 1)));
    CK(m.vertices.alloc(3 * (size_t)std::max(nv, 1))); CK(m.normals.alloc(3 * (size_t)std::max(nq, 1))); CK(m.colors.alloc(3 * (size_t)std::max(nq, 1)));
    CK(m.strain.alloc(std::max(s->L, 1))); CK(m.max_strain.alloc(std::max(s->L, 1))); CK(m.eps_fail.alloc(ef.size())); CK(m.eps_yield.alloc(eyld.size()));
    CK(m.mat_rgb.alloc(m.rgb_host.size())); CK(m.vals.alloc(std::max(std::max(s->N, s->L), 1)));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(m.vert_vox.p, vert_vox.data(), vert_vox.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (nq) {
        CK(cudaMemcpy(m.quads.p, quads.data(), quads.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(m.quad_vox.p, quad_vox.data(), (size_t)nq * sizeof(int), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(m.eps_fail.p, ef.data(), ef.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.eps_yield.p, eyld.data(), eyld.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.mat_rgb.p, m.rgb_host.data(), m.rgb_host.size() * sizeof(float), cudaMemcpyHostToDevice));
    m.n_vert = nv; m.n_quad = nq; m.built = true;
    if (n_vertices) *n_vertices = nv;
    if (n_quads) *n_quads = nq;
    return VX_OK;
}

int vx_mesh_update(vx_sim* s, int coloring, int state_type)
{
    if (!s || coloring < MESH_MATERIAL || coloring > MESH_STATE_INFO || state_type < 0 || state_type > SI_MASS) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_mesh_update inside vx_step_begin .. vx_step_end");
    vx_sim::Mesh& m = s->mesh;
    if (!m.built) { int rc = vx_mesh_build(s, nullptr, nullptr); if (rc != VX_OK) return rc; }
    if (m.n_vert == 0) return VX_OK;
    NvtxRange nvtx("vx_mesh_update");
    { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    CK(cudaSetDevice(s->device));
    float max_val = 0.0f;
    if (coloring == MESH_STATE_INFO) {                              // src/VX_MeshRender.cpp:166-175
        int rc = state_info_impl(s, state_type, SI_MAX, &max_val, m.vals.p);
        if (rc != VX_OK) return rc;
        if (state_type == SI_PRESSURE) {
            float min_val = 0.0f;
            rc = state_info_impl(s, state_type, SI_MIN, &min_val, nullptr);
            if (rc != VX_OK) return rc;
            max_val = max_val > -min_val ? max_val : -min_val;
        }
    }
    if (s->L) {
        int rc = gather_link_field(s, G_STRAIN, m.strain.p);
        if (rc == VX_OK) rc = gather_link_field(s, G_MAXSTRAIN, m.max_strain.p);
        if (rc != VX_OK) return rc;
    }
    MeshFrame mf{};
    mf.n_vert = m.n_vert; mf.n_quad = m.n_quad; mf.vert_vox = m.vert_vox.p; mf.quads = m.quads.p; mf.quad_vox = m.quad_vox.p;
    mf.vertices = m.vertices.p; mf.normals = m.normals.p; mf.colors = m.colors.p; mf.e2i = s->vox_e2i_dev.p; mf.vlinks = s->si_vlinks.p;
    mf.strain = m.strain.p; mf.max_strain = m.max_strain.p; mf.ratio = s->si_ratio.p; mf.eps_fail = m.eps_fail.p; mf.eps_yield = m.eps_yield.p;
    mf.mat_rgb = m.mat_rgb.p; mf.vox_val = m.vals.p; mf.link_val = m.vals.p; mf.n_user = s->N_user;
    const Frame f = s->frame();
    k_mesh_vertices<<<blocks_for(m.n_vert), TPB, 0, s->stream>>>(f, mf);
    k_mesh_quads<<<blocks_for(m.n_quad), TPB, 0, s->stream>>>(f, mf, coloring, state_type, max_val);
    s->launches += 2;
    CK(cudaGetLastError());
    return VX_OK;
}

int vx_mesh_counts(vx_sim* s, int* n_vertices, int* n_quads)
{
    if (!s) return VX_ERR_ARG;
    if (n_vertices) *n_vertices = s->mesh.built ? s->mesh.n_vert : 0;
    if (n_quads) *n_quads = s->mesh.built ? s->mesh.n_quad : 0;
    return VX_OK;
}

int vx_mesh_download(vx_sim* s, float* vertices, int32_t* quads, float* normals, float* colors, int32_t* quad_voxel)
{
    if (!s || !s->mesh.built) return VX_ERR_ARG;
    vx_sim::Mesh& m = s->mesh;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    if (vertices && m.n_vert) CK(cudaMemcpy(vertices, m.vertices.p, 3 * (size_t)m.n_vert * sizeof(float), cudaMemcpyDeviceToHost));
    if (quads && m.n_quad) CK(cudaMemcpy(quads, m.quads.p, 4 * (size_t)m.n_quad * sizeof(int), cudaMemcpyDeviceToHost));
    if (normals && m.n_quad) CK(cudaMemcpy(normals, m.normals.p, 3 * (size_t)m.n_quad * sizeof(float), cudaMemcpyDeviceToHost));
    if (colors && m.n_quad) CK(cudaMemcpy(colors, m.colors.p, 3 * (size_t)m.n_quad * sizeof(float), cudaMemcpyDeviceToHost));
    if (quad_voxel && m.n_quad) CK(cudaMemcpy(quad_voxel, m.quad_vox.p, (size_t)m.n_quad * sizeof(int), cudaMemcpyDeviceToHost));
    return VX_OK;
}

int vx_mesh_device(vx_sim* s, uint64_t* vertices, uint64_t* quads, uint64_t* normals, uint64_t* colors)
{
    if (!s || !s->mesh.built) return VX_ERR_ARG;
    if (vertices) *vertices = (uint64_t)(uintptr_t)s->mesh.vertices.p;
    if (quads) *quads = (uint64_t)(uintptr_t)s->mesh.quads.p;
    if (normals) *normals = (uint64_t)(uintptr_t)s->mesh.normals.p;
    if (colors) *colors = (uint64_t)(uintptr_t)s->mesh.colors.p;
    return VX_OK;
}
