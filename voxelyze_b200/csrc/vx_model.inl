// vx_model.inl -- model definition through the C-ABI: materials (CVX_Material setters, src/VX_Material.cpp:283-523), the setVoxel
// sequence with the layout decision (fused lattice / general, src/Voxelyze.cpp:422-461, 508-539), externals (CVX_External).
// Included by vx_capi.cu inside its extern "C" block; not a translation unit of its own.

static int relayout_fresh(vx_sim* s);
static int relayout_keep_state(vx_sim* s);

int vx_set_materials(vx_sim* s, int n, const vx_material_desc* d)
{
    if (!s || n < 0 || (n && !d)) return VX_ERR_ARG;
    if (n > VX_MAX_VOXMATS) return fail(s, VX_ERR_ARG, "too many materials");
    if (s->N > 0 && n != (int)s->mats.size()) return fail(s, VX_ERR_ARG, "material count changed after voxels were set");
    std::vector<vxm::Material> nm(n);
    std::vector<std::vector<float>> ne(n), ns(n);
    for (int i = 0; i < n; i++) {
        if (d[i].model == VX_MODEL_DATA) {
            if (d[i].n_points < 0 || !d[i].strain || !d[i].stress) return fail(s, VX_ERR_ARG, "data model without points");
            ne[i].assign(d[i].strain, d[i].strain + d[i].n_points);
            ns[i].assign(d[i].stress, d[i].stress + d[i].n_points);
        }
        if (!vxm::from_desc(nm[i], d[i], ne[i].data(), ns[i].data())) return fail(s, VX_ERR_MATERIAL, nm[i].error);
    }
    s->descs.assign(d, d + n);
    for (auto& x : s->descs) x.strain = x.stress = nullptr;
    s->d_eps.swap(ne); s->d_sig.swap(ns); s->mats.swap(nm);
    // Poisson's ratio may be switched on or off at any time (the reference allows it, src/VX_Link.cpp:160-166): on the fused
    // layout the per-voxel pStrain arrays are created from the current link strains the first time it becomes non-zero
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_set_materials inside vx_step_begin .. vx_step_end");
    const bool had_poisson = s->any_poisson;
    int rc = upload_tables(s);
    if (rc != VX_OK) return rc;
    if (s->lattice && had_poisson != s->any_poisson) { find_boundary_layers(s); s->drop_graph(); }      // z-slabs: the ghost-skipping kernel has no Poisson variant
    if (!s->lattice && s->any_poisson && s->state_ready && s->L > 0) {
        k_refresh_slot_strain<<<blocks_for(s->L), TPB, 0, s->stream>>>(s->frame(), s->axis_first[1], s->axis_first[2]); s->launches++;
        CK(cudaGetLastError());
    }
    return refresh_lattice_ps(s);
}

int vx_get_voxmat(const vx_sim* s, int i, vx_voxmat_row* o)
{
    if (!s || !o || i < 0 || i >= (int)s->mats.size()) return VX_ERR_ARG;
    vxm::fill_row(*o, s->mats[i], s->vox_size);
    return VX_OK;
}
int vx_get_linkmat(vx_sim* s, int a, int b, vx_linkmat_row* o)
{
    if (!s || !o || a < 0 || b < 0 || a >= (int)s->mats.size() || b >= (int)s->mats.size()) return VX_ERR_ARG;
    vxm::Material m = vxm::combine(s->mats[a], s->mats[b]);
    vxm::fill_row(*o, m, s->vox_size, a, b);
    return VX_OK;
}
int vx_get_linkmat_curve(vx_sim* s, int a, int b, float* eps, float* sig, int cap)
{
    if (!s || a < 0 || b < 0 || a >= (int)s->mats.size() || b >= (int)s->mats.size()) return VX_ERR_ARG;
    vxm::Material m = vxm::combine(s->mats[a], s->mats[b]);
    int n = (int)m.eps.size();
    if (eps && sig) for (int i = 0; i < n && i < cap; i++) { eps[i] = m.eps[i]; sig[i] = m.sig[i]; }
    return n;
}

static int set_voxels_impl(vx_sim* s, int n, const int32_t* ijk, const uint16_t* mat, const int32_t* sim_id, const uint32_t* flags, int n_user);

// Small models are stepped by one thread-block cluster that runs a whole vx_step call in a single launch
// (k_small_steps, general layout): chosen by vx_set_path(3), or by default up to VX_SMALL_MAX voxels when the model has no
// halo / per-voxel flags and self-collisions are off at this point (with collisions the fused path's captured graphs win).
constexpr int VX_SMALL_MAX = 700;       // ~2 000 links: one pass of a 16 x 128-thread cluster; beyond, the fused kernel's graphs are as fast or faster
static bool small_model(const vx_sim* s, int n, const uint32_t* flags)
{
    if (s->path != 3 && !(s->path == 0 && n <= VX_SMALL_MAX && !s->collisions && !getenv("VX_NO_SMALL"))) return false;
    if (flags) for (int i = 0; i < n; i++) if (flags[i] & VX_VF_GHOST) return false;
    return n > 0;
}

// A box with holes still runs on the fused lattice path: the missing cells are appended as inert voxels (never
// integrated, no links, invisible to the caller) when that costs at most 60 % more cells.
int vx_set_voxels(vx_sim* s, int n, const int32_t* ijk, const uint16_t* mat, const int32_t* sim_id, const uint32_t* flags)
{
    if (!s || n < 0 || (n && (!ijk || !mat))) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_set_voxels inside vx_step_begin .. vx_step_end");
    if (n == 0 || s->path == 1 || small_model(s, n, flags))
        return set_voxels_impl(s, n, ijk, mat, sim_id, flags, n);
    int lo[3] = {32767, 32767, 32767}, hi[3] = {-32768, -32768, -32768}, members = 1;
    for (int i = 0; i < n; i++) {
        for (int a = 0; a < 3; a++) {
            int c = ijk[3 * i + a];
            if (c < -32768 || c > 32767) return fail(s, VX_ERR_ARG, "lattice index does not fit a short");
            lo[a] = std::min(lo[a], c); hi[a] = std::max(hi[a], c);
        }
        if (sim_id) { if (sim_id[i] < 0 || sim_id[i] > 65535) return fail(s, VX_ERR_ARG, "bad member id"); members = std::max(members, sim_id[i] + 1); }
    }
    const long long ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1, cells = ex * ey * ez * members;
    // up to 60 % more cells for any model; a single body without halo voxels may be as sparse as one voxel in eight cells
    // (only its occupied 8x8x4 brick groups are launched, LatFrame::groups), as long as the padded arrays stay below ~60 GB
    bool user_flags = false;
    if (flags) for (int i = 0; i < n && !user_flags; i++) user_flags = flags[i] != 0;
    const double max_ratio = (members == 1 && !user_flags && cells <= 100000000LL) ? 8.0 : 1.6;
    if (cells == n || cells > (long long)(max_ratio * n) || cells > 2000000000LL) return set_voxels_impl(s, n, ijk, mat, sim_id, flags, n);
    std::vector<char> used((size_t)cells, 0);
    for (int i = 0; i < n; i++) {
        const long long c = ((((long long)(sim_id ? sim_id[i] : 0) * ez + (ijk[3 * i + 2] - lo[2])) * ey + (ijk[3 * i + 1] - lo[1])) * ex + (ijk[3 * i] - lo[0]));
        if (used[(size_t)c]) return fail(s, VX_ERR_TOPOLOGY, "duplicate voxel");
        used[(size_t)c] = 1;
    }
    std::vector<int32_t> ijk2(ijk, ijk + 3 * (size_t)n), sim2; std::vector<uint16_t> mat2(mat, mat + n); std::vector<uint32_t> fl2(n, 0u);
    if (flags) fl2.assign(flags, flags + n);
    if (sim_id || members > 1) sim2.assign(sim_id, sim_id + n);
    for (long long c = 0; c < cells; c++) {
        if (used[(size_t)c]) continue;
        long long r = c; const int x = (int)(r % ex); r /= ex; const int y = (int)(r % ey); r /= ey; const int z = (int)(r % ez); const int m = (int)(r / ez);
        ijk2.push_back(lo[0] + x); ijk2.push_back(lo[1] + y); ijk2.push_back(lo[2] + z);
        mat2.push_back(mat[0]); fl2.push_back(VX_VF_GHOST | VF_FILL);
        if (!sim2.empty()) sim2.push_back(m);
    }
    return set_voxels_impl(s, (int)cells, ijk2.data(), mat2.data(), sim2.empty() ? nullptr : sim2.data(), fl2.data(), n);
}

static int set_voxels_impl(vx_sim* s, int n, const int32_t* ijk, const uint16_t* mat, const int32_t* sim_id, const uint32_t* flags, int n_user)
{
    if (!s || n < 0 || (n && (!ijk || !mat))) return VX_ERR_ARG;
    CK(cudaSetDevice(s->device));
    // ---- validate + bounding box
    int max_member = 0;
    int lo[3] = {32767, 32767, 32767}, hi[3] = {-32768, -32768, -32768};
    for (int i = 0; i < n; i++) {
        if (mat[i] >= s->mats.size()) return fail(s, VX_ERR_ARG, "material index out of range");
        for (int a = 0; a < 3; a++) {
            int c = ijk[3 * i + a];
            if (c < -32768 || c > 32767) return fail(s, VX_ERR_ARG, "lattice index does not fit a short");
            lo[a] = std::min(lo[a], c); hi[a] = std::max(hi[a], c);
        }
        int m = sim_id ? sim_id[i] : 0;
        if (m < 0 || m > 65535) return fail(s, VX_ERR_ARG, "bad member id");
        max_member = std::max(max_member, m);
    }
    s->N = n; s->N_user = n_user; s->n_members = max_member + 1;
    s->ijk.assign(ijk, ijk + 3 * (size_t)n);
    s->vmat_id.assign(mat, mat + n);
    s->member.assign(n, 0); if (sim_id) s->member.assign(sim_id, sim_id + n);
    s->vflags.clear(); if (flags) s->vflags.assign(flags, flags + n);
    s->ext_vox.clear(); s->ext_rows.clear();
    if (!s->relayout) { s->ext_raw_vox.clear(); s->ext_raw_dof.clear(); s->ext_raw_f.clear(); s->ext_raw_m.clear(); s->ext_raw_t.clear(); s->ext_raw_r.clear(); }
    s->lmats.clear(); s->lmat_of.clear();
    s->drop_graph();

    // ---- occupancy lookup: dense grid over the common bounding box when affordable, else hash
    long long ext3[3] = {n ? hi[0] - lo[0] + 1 : 0, n ? hi[1] - lo[1] + 1 : 0, n ? hi[2] - lo[2] + 1 : 0};
    long long cells = ext3[0] * ext3[1] * ext3[2] * (long long)s->n_members;
    bool dense = n > 0 && cells <= std::max<long long>(8LL * n, 1 << 20);
    std::vector<int32_t> grid;
    std::unordered_map<uint64_t, int32_t> hash;
    auto cell_of = [&](int m, int x, int y, int z) -> long long {
        if (x < lo[0] || x > hi[0] || y < lo[1] || y > hi[1] || z < lo[2] || z > hi[2]) return -1;
        return (((long long)m * ext3[2] + (z - lo[2])) * ext3[1] + (y - lo[1])) * ext3[0] + (x - lo[0]);
    };
    auto key_of = [](int m, int x, int y, int z) -> uint64_t {
        return ((uint64_t)(uint32_t)m << 48) | ((uint64_t)(uint16_t)(int16_t)x << 32) | ((uint64_t)(uint16_t)(int16_t)y << 16) | (uint64_t)(uint16_t)(int16_t)z;
    };
    if (dense) grid.assign((size_t)cells, -1); else hash.reserve((size_t)n * 2);
    auto lookup = [&](int m, int x, int y, int z) -> int {
        if (dense) { long long c = cell_of(m, x, y, z); return c < 0 ? -1 : grid[(size_t)c]; }
        auto it = hash.find(key_of(m, x, y, z)); return it == hash.end() ? -1 : it->second;
    };

    // ---- links in the reference's creation order (src/Voxelyze.cpp:453-455, 508-539)
    s->lk_vn.clear(); s->lk_vp.clear(); s->lk_axis.clear();
    s->linkmask.assign(n, 0);
    s->nbr.clear(); s->col_tables = false; s->col_stale_host = true; s->n_surf = 0; s->n_pairs = 0;
    if (s->collisions) s->nbr.assign((size_t)n * 6, -1);
    std::vector<int32_t> plus_link((size_t)n * 3, -1);        // caller link index of the +axis link of each voxel
    static const int dx[6] = {1, -1, 0, 0, 0, 0}, dy[6] = {0, 0, 1, -1, 0, 0}, dz[6] = {0, 0, 0, 0, 1, -1};
    for (int i = 0; i < n; i++) {
        int m = s->member[i], x = ijk[3 * i], y = ijk[3 * i + 1], z = ijk[3 * i + 2];
        if (lookup(m, x, y, z) >= 0) return fail(s, VX_ERR_TOPOLOGY, "duplicate voxel");
        if (dense) grid[(size_t)cell_of(m, x, y, z)] = i; else hash[key_of(m, x, y, z)] = i;
        bool gi = flags && (flags[i] & VX_VF_GHOST);
        if (flags && (flags[i] & VF_FILL)) continue;            // a fill cell has no links
        for (int d = 0; d < 6; d++) {
            int o = lookup(m, x + dx[d], y + dy[d], z + dz[d]);
            if (o < 0) continue;
            if (flags && (flags[o] & VF_FILL)) continue;
            if (gi && (flags[o] & VX_VF_GHOST)) continue;      // halo-halo links are never needed
            bool this_neg = (d % 2) == 0;                      // src/VX_Link.cpp:31-53
            int vn = this_neg ? i : o, vp = this_neg ? o : i;
            int li = (int)s->lk_vn.size();
            s->lk_vn.push_back(vn); s->lk_vp.push_back(vp); s->lk_axis.push_back((uint8_t)(d / 2));
            s->linkmask[i] |= (uint8_t)(1u << d); s->linkmask[o] |= (uint8_t)(1u << (d ^ 1));
            if (s->collisions) { s->nbr[(size_t)i * 6 + d] = o; s->nbr[(size_t)o * 6 + (d ^ 1)] = i; }
            plus_link[(size_t)vn * 3 + d / 2] = li;
        }
    }
    const int L = s->L = (int)s->lk_vn.size();

    // ---- layout: a completely filled box (per member) runs fused.  The members of an ensemble are tiled side by side into
    // one device lattice when that fills the 8 x 8 x 4 voxel tiles of the fused kernel better than one box per member
    // (4096 robots of 10^3: 4 x 4 x 256 robots = a 40 x 40 x 2560 lattice without a single idle lane, against 69 % lane
    // use for 10^3 boxes on their own).  Members never link: every link bit comes from a per-member neighbour lookup.
    bool poisson = false, halo = false;
    for (auto& m : s->mats) if (m.nu != 0.0f) poisson = true;
    if (flags) for (int i = 0; i < n && !halo; i++) halo = (flags[i] & VX_VF_GHOST) && !(flags[i] & VF_FILL);
    // fused layout: a completely filled box; Poisson materials too (k_lattice_tma<.., POISSON>; on a z-slab the halo carries
    // the ghosts' Poisson strains along with their poses)
    s->small = n == n_user && small_model(s, n, flags);
    s->lattice = n > 0 && cells == (long long)n && s->path != 1 && !s->small;
    s->state_ready = false;
    s->pack[0] = s->pack[1] = 1; s->pack[2] = s->n_members; s->lat_members = s->n_members;
    if (s->lattice && s->n_members > 1 && !getenv("VX_NO_PACK")) {
        auto up = [](long long v, long long q) { return (v + q - 1) / q * q; };
        double best = 0; int bx_ = 1, by_ = 1;
        for (int px = 1; px <= 8; px *= 2)
            for (int py = 1; py <= px; py *= 2) {
                if (s->n_members % (px * py) || px * ext3[0] > 30000 || py * ext3[1] > 30000) continue;
                const long long pz = s->n_members / (px * py);
                const double cost = (double)up(px * ext3[0], 2 * VX_WB_X) * up(py * ext3[1], 2 * VX_WB_Y) * up(pz * ext3[2], 2 * VX_WB_Z);
                if (best == 0 || cost < best * 0.999) { best = cost; bx_ = px; by_ = py; }
            }
        if (bx_ * by_ > 1) { s->pack[0] = bx_; s->pack[1] = by_; s->pack[2] = s->n_members / (bx_ * by_); s->lat_members = 1; }
    }
    const bool packed = s->lat_members == 1 && s->n_members > 1;

    // ---- internal voxel order: (member, z, y, x), or (Z, Y, X) in the tiled lattice of a packed ensemble
    s->v_i2e.resize(n);
    for (int i = 0; i < n; i++) s->v_i2e[i] = i;
    auto vkey = [&](int e) -> int64_t {
        if (packed) {
            const int m = s->member[e], mi = m % s->pack[0], mj = (m / s->pack[0]) % s->pack[1], mk = m / (s->pack[0] * s->pack[1]);
            const int64_t X = (int64_t)mi * ext3[0] + (ijk[3 * e] - lo[0]), Y = (int64_t)mj * ext3[1] + (ijk[3 * e + 1] - lo[1]), Z = (int64_t)mk * ext3[2] + (ijk[3 * e + 2] - lo[2]);
            return (Z << 40) | (Y << 20) | X;
        }
        return ((int64_t)s->member[e] << 48) | ((int64_t)(ijk[3 * e + 2] + 32768) << 32) | ((int64_t)(ijk[3 * e + 1] + 32768) << 16) | (int64_t)(ijk[3 * e] + 32768);
    };
    {
        bool sorted = true;
        for (int i = 1; i < n && sorted; i++) if (vkey(i - 1) > vkey(i)) sorted = false;
        if (!sorted) std::sort(s->v_i2e.begin(), s->v_i2e.end(), [&](int a, int b) { return vkey(a) < vkey(b); });
    }
    s->v_e2i.resize(n); s->sort_key.resize(n);
    for (int i = 0; i < n; i++) { s->v_e2i[s->v_i2e[i]] = i; s->sort_key[i] = vkey(s->v_i2e[i]) >> 32; }

    // ---- internal link order: by axis, then by internal index of the negative-end voxel
    s->l_i2e.clear(); s->l_i2e.reserve(L);
    for (int a = 0; a < 3; a++) {
        s->axis_first[a] = (int)s->l_i2e.size();
        for (int i = 0; i < n; i++) { int li = plus_link[(size_t)s->v_i2e[i] * 3 + a]; if (li >= 0) s->l_i2e.push_back(li); }
    }
    s->axis_first[3] = (int)s->l_i2e.size();
    s->l_e2i.resize(L);
    for (int i = 0; i < L; i++) s->l_e2i[s->l_i2e[i]] = i;

    s->nx = (int)ext3[0]; s->ny = (int)ext3[1]; s->nz = (int)ext3[2];
    if (packed) { s->nx *= s->pack[0]; s->ny *= s->pack[1]; s->nz *= s->pack[2]; }
    s->link_owner.release(); s->link_axis_dev.release();
    s->si_nominal_ok = false; s->si_consts_ok = false; s->si_pressure_ok = false; s->mesh.built = false;

    for (int i = 0; i < L; i++) {
        int id = link_material(s, s->vmat_id[s->lk_vn[i]], s->vmat_id[s->lk_vp[i]]);
        if (id > 0xFFFF) return fail(s, VX_ERR_ARG, "too many link materials");
    }

    // ---- device memory
    size_t n1 = std::max(n, 1), l1 = std::max(L, 1);
    CK(s->ext_idx.alloc(n1)); CK(s->vox_e2i_dev.alloc(n1)); CK(s->link_e2i_dev.alloc(l1)); CK(s->member_dev.alloc(n1));
    if (s->lattice) {
        for (int g = 0; g < 2; g++) {
            CK(s->alloc_pose(g, n1)); CK(s->mom0[g].alloc(n1)); CK(s->mom1[g].alloc(n1));
            CK(s->rec[g].alloc(n1 * VX_REC_PARTS));
            if (poisson) CK(s->ps[g].alloc(n1)); else s->ps[g].release();
        }
        s->slots.release(); s->lends.release(); s->lmeta.release(); s->lstA.release(); s->lstB.release(); s->lstC.release();
        s->lstrain.release(); s->pstrain.release(); s->slot_strain.release();
        s->lk_mat.clear();
        s->tmaps.release();                           // describe the old arrays
    } else {
        CK(s->alloc_pose(0, n1)); CK(s->mom0[0].alloc(n1)); CK(s->mom1[0].alloc(n1));
        for (int g = 0; g < 2; g++) { s->rec[g].release(); s->ps[g].release(); }
        s->release_pose(1); s->mom0[1].release(); s->mom1[1].release();
        CK(s->slots.alloc(n1 * 36));
        CK(s->lends.alloc(l1)); CK(s->lmeta.alloc(l1)); CK(s->lstA.alloc(l1)); CK(s->lstB.alloc(l1)); CK(s->lstC.alloc(l1)); CK(s->lstrain.alloc(l1));
        CK(s->pstrain.alloc(n1)); CK(s->slot_strain.alloc(n1 * 6));
        s->lk_mat.resize(L);
        std::vector<int2> ends(L);
        for (int i = 0; i < L; i++) {
            int e = s->l_i2e[i];
            s->lk_mat[i] = (uint16_t)link_material(s, s->vmat_id[s->lk_vn[e]], s->vmat_id[s->lk_vp[e]]);
            ends[i] = make_int2(s->v_e2i[s->lk_vn[e]], s->v_e2i[s->lk_vp[e]]);
        }
        CK(cudaStreamSynchronize(s->stream));
        if (L) CK(cudaMemcpy(s->lends.p, ends.data(), (size_t)L * sizeof(int2), cudaMemcpyHostToDevice));
    }
    CK(cudaStreamSynchronize(s->stream));
    if (n) {
        CK(cudaMemcpy(s->vox_e2i_dev.p, s->v_e2i.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        std::vector<int> mem_internal(n);
        for (int i = 0; i < n; i++) mem_internal[i] = s->member[s->v_i2e[i]];
        CK(cudaMemcpy(s->member_dev.p, mem_internal.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (L) CK(cudaMemcpy(s->link_e2i_dev.p, s->l_e2i.data(), (size_t)L * sizeof(int), cudaMemcpyHostToDevice));
    int rc = upload_tables(s);                     // new link materials may have appeared
    if (rc != VX_OK) return rc;
    if (s->collisions) { rc = build_collision_tables(s); if (rc != VX_OK) return rc; }
    find_boundary_layers(s);
    // sparse body on the fused layout: list of the 8x8x4 brick groups that hold at least one real voxel
    s->n_groups = 0;
    if (s->lattice && s->N_user < s->N && s->lat_members == 1 && s->zb_layers.empty()) {
        const int gx = (s->nx + 2 * VX_WB_X - 1) / (2 * VX_WB_X), gy = (s->ny + 2 * VX_WB_Y - 1) / (2 * VX_WB_Y), gz = (s->nz + 2 * VX_WB_Z - 1) / (2 * VX_WB_Z);
        if (gx <= 1024 && gy <= 1024 && gz <= 2048) {
            std::vector<char> occupied((size_t)gx * gy * gz, 0);
            for (int e = 0; e < s->N_user; e++) {
                const int x = ijk[3 * e] - lo[0], y = ijk[3 * e + 1] - lo[1], z = ijk[3 * e + 2] - lo[2];
                occupied[((size_t)(z / (2 * VX_WB_Z)) * gy + y / (2 * VX_WB_Y)) * gx + x / (2 * VX_WB_X)] = 1;
            }
            std::vector<int> list;
            for (int z = 0; z < gz; z++) for (int y = 0; y < gy; y++) for (int x = 0; x < gx; x++)
                if (occupied[((size_t)z * gy + y) * gx + x]) list.push_back(x | (y << 10) | (z << 20));
            if (list.size() * 10 <= occupied.size() * 9) {               // worth it from 10 % empty groups on
                CK(s->group_list.alloc(list.size()));
                CK(cudaMemcpy(s->group_list.p, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice));
                s->n_groups = (int)list.size();
            }
        }
    }
    // a voxel set replaced in the middle of a run: simulation time and CVX_Voxel::previousDt go on (setVoxel does not
    // touch them, src/Voxelyze.cpp:422-498); vx_reset is what rewinds them
    const float time = s->time_host, prev_dt = s->prev_dt_host;
    rc = upload_initial_state(s, s->ambient);      // new voxels start at ambient temperature, src/Voxelyze.cpp:449
    if (rc != VX_OK || (time == 0.f && prev_dt == 0.f)) return rc;
    DevParams p{}; p.col_stale = 1; p.time = time; p.prev_dt = prev_dt;
    CK(cudaMemcpy(s->params.p, &p, sizeof(p), cudaMemcpyHostToDevice));
    s->time_host = time; s->prev_dt_host = prev_dt;
    return VX_OK;
}

int vx_voxel_count(const vx_sim* s) { return s ? s->N_user : 0; }
int vx_link_count(const vx_sim* s) { return s ? s->L : 0; }
int vx_get_links(const vx_sim* s, int32_t* vn, int32_t* vp, uint8_t* ax)
{
    if (!s) return VX_ERR_ARG;
    if (vn) memcpy(vn, s->lk_vn.data(), (size_t)s->L * sizeof(int32_t));
    if (vp) memcpy(vp, s->lk_vp.data(), (size_t)s->L * sizeof(int32_t));
    if (ax) memcpy(ax, s->lk_axis.data(), (size_t)s->L);
    return VX_OK;
}

int vx_set_externals(vx_sim* s, int n, const int32_t* voxel, const uint8_t* dof, const float* force, const float* moment,
                     const double* tr, const double* rot)
{
    if (!s || n < 0 || (n && (!voxel || !dof))) return VX_ERR_ARG;
    for (int k = 0; k < n; k++) if (voxel[k] < 0 || voxel[k] >= s->N_user) return fail(s, VX_ERR_ARG, "external voxel index out of range");
    CK(cudaSetDevice(s->device));
    if (voxel != s->ext_raw_vox.data()) {
        s->ext_raw_vox.assign(voxel, voxel + n); s->ext_raw_dof.assign(dof, dof + n);
        s->ext_raw_f.clear(); s->ext_raw_m.clear(); s->ext_raw_t.clear(); s->ext_raw_r.clear();
        if (force) s->ext_raw_f.assign(force, force + 3 * (size_t)n);
        if (moment) s->ext_raw_m.assign(moment, moment + 3 * (size_t)n);
        if (tr) s->ext_raw_t.assign(tr, tr + 3 * (size_t)n);
        if (rot) s->ext_raw_r.assign(rot, rot + 3 * (size_t)n);
    }
    s->ext_vox.assign(voxel, voxel + n);
    s->ext_rows.assign(n, DevExt{});
    for (int k = 0; k < n; k++) {
        DevExt& e = s->ext_rows[k];
        int v = voxel[k];
        for (int a = 0; a < 3; a++) {
            e.nominal[a] = s->ijk[3 * v + a] * s->vox_size;
            e.translation[a] = tr ? tr[3 * k + a] : 0.0;
            e.force[a] = force ? force[3 * k + a] : 0.0f;
            e.moment[a] = moment ? moment[3 * k + a] : 0.0f;
        }
        e.rot_q[0] = 1.0; e.rot_q[1] = e.rot_q[2] = e.rot_q[3] = 0.0;
        if (rot && (rot[3 * k] != 0 || rot[3 * k + 1] != 0 || rot[3 * k + 2] != 0)) {
            // Quat3D::FromRotationVector on the host (include/Quat3D.h:124-139, src/VX_External.cpp:100-109)
            double hx = 0.5 * rot[3 * k], hy = 0.5 * rot[3 * k + 1], hz = 0.5 * rot[3 * k + 2];
            double m2 = hx * hx + hy * hy + hz * hz, w, sc;
            if (m2 * m2 < 5.328e-15) { w = 1.0 - 0.5 * m2; sc = 1.0 - m2 / 6.0; }
            else { double m = std::sqrt(m2); w = std::cos(m); sc = std::sin(m) / m; }
            e.rot_q[0] = w; e.rot_q[1] = hx * sc; e.rot_q[2] = hy * sc; e.rot_q[3] = hz * sc;
        }
        e.dof = dof[k] & 0x3F;
    }
    return upload_externals(s);
}
