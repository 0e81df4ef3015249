// vx_lattice.cuh -- fused single-pass step for dense box lattices (the headline path).
//
// One kernel launch per doTimeStep: the (up to) six links that touch a voxel are evaluated, their forces summed in
// the reference's slot order X+,X-,Y+,Y-,Z+,Z- and the voxel integrated -- link forces never travel through HBM.
//   * A voxel update reads only OLD poses (its own and its six face neighbours'), which is exactly what the
//     reference's "all links, then all voxels" order does (src/Voxelyze.cpp:251-284), so state is kept in
//     ping-pong buffers: read generation g, write g^1, swap per step.
//   * A link is owned by its negative-end voxel (its +X/+Y/+Z link); the owner advances the link's persistent
//     state.  Where the two ends of a link are handled by different warps/blocks the link is re-evaluated
//     from the same old inputs (identical bits) instead of shipping its force.
//   * Implicit indexing: neighbour = v +- {1, nx, nx*ny}; no index arrays are read.
//
// Kernels (both produce identical bits, and the same bits as the general path; vx_set_path selects,
// tests/test_gpu_parity.py compares them):
//   k_lattice_tma     7, and what 0 picks on large lattices: one warp per 4x4x2 brick, staged by TMA
//   k_lattice_warp    5, and what 0 picks for ensembles of small boxes: same, staged by cp.async
// (Round 1 also carried four slower formulations -- block bricks with barriers, one thread per voxel, z-marching
// warps, marching warp bricks; they were ablations, are documented in profiles/r1_*_ncu_summary.md and no longer ship.)
//
// HBM traffic per voxel per step (nu = 0):  read 64 B pose + 48 B momenta + 3 x 64 B link records,
// write the same = 608 B, vs. 216 + 3 x 268 = 1020 B "algorithmic" bytes of SURVEY.md section 8d (the
// 96 B/link force write of the reference layout is not needed: forces of the last step are
// recomputed on demand from the two buffer generations, see k_lattice_gather_links).
//
// Link record (64 B): six doubles + float4 {strain, maxStrain, strainOffset, stress}.
// The nine doubles pos2/angle1v/angle2v of the reference collapse to six without loss:
//   small-angle mode: angle1v == 0            -> { pos2.x, pos2.y, pos2.z, angle2v.xyz }
//   large-angle mode: pos2.y == pos2.z == 0, angle1v.x == 0 -> { pos2.x, angle1v.y, angle1v.z, angle2v.xyz }
// The two mode bits per link live in the owner voxel's meta word (bits 26..31).
#pragma once
#include "vx_kernels.cuh"

namespace vxd {

#define VM_LFLAG_SHIFT 26     // 2 bits per owned link: bit 0 small-angle, bit 1 local-velocity-valid
#define VX_REC_PARTS 12       // sixteen-byte parts of the link records of one voxel: axis a -> parts 4a..4a+2 (double2) and 4a+3 (float4)

struct LatFrame {
    int nx, ny, nz, nxy, n_vox, n_mat, n_lmat;
    // voxel state, current (read) and next (write) generation
    const double4* c_pose0; const double4* c_pose1; const double4* c_mom0; const double2* c_mom1;
    double4* n_pose0; double4* n_pose1; double4* n_mom0; double2* n_mom1;
    // link records owned by voxel v for axis a: rec[a][0..2][v] (double2 x3) + recf[a][v] (float4); all twelve arrays of a
    // generation are one allocation, part (4a + k) at c_rec[0][0] + (4a + k) * n_vox, recf[a] being part 4a + 3
    const double2* c_rec[3][3]; const float4* c_recf[3];
    double2* n_rec[3][3]; float4* n_recf[3];
    const int* ext_idx;
    const DevVoxMat* vmat; const DevLinkMat* lmat; const float* curve_e; const float* curve_s;
    const uint16_t* pair_lmat;      // [n_mat][n_mat] -> link material
    const DevExt* ext;
    DevParams* params;
    DevVoxMat vm0; DevLinkMat lm0;  // single-material models: rows in the constant bank (UNI)
    // collisions (k_lattice_warp / k_lattice_tma): per surface voxel CSR of signed contact references, as in Frame
    const int* col_slot; const int* col_start; const int* col_ref; const float4* col_force;
    // z-slab runs (k_lattice_warp only): voxels of plane push_z[k] also store their new pose into the ghost
    // plane of the neighbouring slab, push0/1[k] = that plane in the neighbour's pose0/pose1 arrays (peer memory)
    int push_z[2]; double4* push0[2]; double4* push1[2];
    // CVoxelyze::setAmbientTemperature(t, true) applied by the step itself: every voxel reads temperature `amb` instead of
    // its stored one in this launch and carries it into the new generation (no separate pass over the voxels)
    int amb_set; float amb;
    // sparse bodies (k_lattice_tma, grouped grids): the occupied 8x8x4-voxel brick groups, gx | gy << 10 | gz << 20, one CTA each
    // (null: every group of the bounding box is launched)
    const int* groups;
    // Poisson coupling (nu != 0, k_lattice_tma<.., POISSON>): CVX_Voxel::pStrain of every voxel as of the state the step reads
    // (= computed from the link strains of the previous step), double-buffered like the voxel state
    const float4* c_ps; float4* n_ps;
    bool stream_out;                    // k_lattice_tma: the state does not fit the L2, results are stored with st.global.cs (st_out)
    float4* push_ps[2];                 // with PUSH: the neighbours' ghost planes in their pStrain arrays of the generation being written
    // z-slab runs on k_lattice_tma<.., GSKIP>: bricks cover the planes [z_lo, z_hi) only -- the all-ghost planes below and above
    // are data, not work (brick origins are shifted by z_lo; 0 / nz everywhere else)
    int z_lo, z_hi;
};

// where a kernel reads the material tables of a multi-material model: global memory (through L1), or a copy the CTA staged
// in its shared memory (k_lattice_tma when the tables are small, which they are for every BASELINE config) together with the
// internal-damping factor 2 sqrt(m) zeta / previousDt of every voxel material, divided once per CTA instead of per link end
struct MatView { const DevVoxMat* vmat; const DevLinkMat* lmat; const uint16_t* pair; const float* damp; int n_mat; };
__device__ __forceinline__ MatView mat_view(const LatFrame& f) { MatView m; m.vmat = f.vmat; m.lmat = f.lmat; m.pair = f.pair_lmat; m.damp = nullptr; m.n_mat = f.n_mat; return m; }

__device__ __forceinline__ void lat_decode(double2 a, double2 b, double2 c, float4 s, uint32_t lflags, LinkState& st)
{
    st.small_angle = (lflags & 1u) != 0;
    st.vel_valid = (lflags & 2u) != 0;
    if (st.small_angle) { st.pos2 = mk3(a.x, a.y, b.x); st.a1v = mk3(0.0, 0.0, 0.0); }
    else { st.pos2 = mk3(a.x, 0.0, 0.0); st.a1v = mk3(0.0, a.y, b.x); }
    st.a2v = mk3(b.y, c.x, c.y);
    st.strain = s.x; st.max_strain = s.y; st.strain_offset = s.z; st.stress = s.w;
}
__device__ __forceinline__ void lat_encode(const LinkState& st, double2& a, double2& b, double2& c, float4& s, uint32_t& lflags)
{
    if (st.small_angle) { a = make_double2(st.pos2.x, st.pos2.y); b.x = st.pos2.z; }
    else { a = make_double2(st.pos2.x, st.a1v.y); b.x = st.a1v.z; }
    b.y = st.a2v.x; c = make_double2(st.a2v.y, st.a2v.z);
    s = make_float4(st.strain, st.max_strain, st.strain_offset, st.stress);
    lflags = (st.small_angle ? 1u : 0u) | (st.vel_valid ? 2u : 0u);
}

// evaluates link (owner, axis) between negative-end voxel N and positive-end voxel P from the
// current generation; returns forces on both ends and the advanced link state
template <bool UNI>
__device__ __forceinline__ void lat_eval_link_rec(const LatFrame& f, int axis, uint32_t owner_bits,
                                                  double2 ra, double2 rb, double2 rc, float4 rs,
                                                  double4 n0, double4 n1, double4 p0, double4 p1, float prev_dt,
                                                  LinkState& st, d3& fN, d3& mN, d3& fP, d3& mP, float damp_uni = -1.0f,
                                                  const float4* psn = nullptr, const float4* psp = nullptr, float* end_strain = nullptr,
                                                  const MatView* tables = nullptr)
{
    // psn/psp: Poisson strains of the two end voxels (nu != 0 models); end_strain[0/1]: axial strain of the half of the link
    // inside the negative / positive end voxel (CVX_Link::axialStrain(bool), src/VX_Link.cpp:121-124), input of the next pStrain
    // damp_uni: single-material models may pass 2*sqrtMass*zeta/previousDt computed once per kernel (same float division)
    const uint32_t hn = meta_hi(n1.w), hp = meta_hi(p1.w);
    const MatView mv = tables ? *tables : mat_view(f);
    const DevVoxMat& vmn = UNI ? f.vm0 : mv.vmat[hn & VM_MAT_MASK];
    const DevVoxMat& vmp = UNI ? f.vm0 : mv.vmat[hp & VM_MAT_MASK];
    const DevLinkMat& lm = UNI ? f.lm0 : mv.lmat[mv.pair[(hn & VM_MAT_MASK) * mv.n_mat + (hp & VM_MAT_MASK)]];
    lat_decode(ra, rb, rc, rs, (owner_bits >> (VM_LFLAG_SHIFT + 2 * axis)) & 3u, st);
    // CVX_Link::updateRestLength (src/VX_Link.cpp:137-140)
    const float tn = f.amb_set ? f.amb : meta_temp(n1.w), tp = f.amb_set ? f.amb : meta_temp(p1.w);
    double rest = 0.5 * (vmn.size[axis] * (1 + tn * vmn.cte) + vmp.size[axis] * (1 + tp * vmp.cte));
    float t_area, t_sum = 0.0f;
    if (psn) {                                                              // CVX_Link::updateTransverseInfo, src/VX_Link.cpp:142-147
        t_area = 0.5f * (transverse_area(vmn, axis, *psn) + transverse_area(vmp, axis, *psp));
        t_sum = 0.5f * (transverse_strain_sum(vmn, axis, *psn) + transverse_strain_sum(vmp, axis, *psp));
    } else t_area = 0.5f * (vmn.nom_f * vmn.nom_f + vmp.nom_f * vmp.nom_f);
    float damp_n, damp_p;
    if (UNI && damp_uni >= 0.0f) damp_n = damp_p = damp_uni;
    else if (!UNI && mv.damp) { damp_n = mv.damp[hn & VM_MAT_MASK]; damp_p = mv.damp[hp & VM_MAT_MASK]; }
    else { damp_n = vmn.two_sqrtm_zeta / prev_dt; damp_p = vmp.two_sqrtm_zeta / prev_dt; }
    q4 on, op;
    on.w = n0.w; on.x = n1.x; on.y = n1.y; on.z = n1.z;
    op.w = p0.w; op.x = p1.x; op.y = p1.y; op.z = p1.z;
    link_forces(axis, mk3(n0.x, n0.y, n0.z), on, mk3(p0.x, p0.y, p0.z), op, rest, t_area, t_sum,
                damp_n, damp_p, lm, f.curve_e, f.curve_s, st, fN, mN, fP, mP);
    if (end_strain) {
        const float ratio = vmp.E / vmn.E;
        end_strain[0] = 2.0f * st.strain / (1.0f + ratio);
        end_strain[1] = 2.0f * st.strain * ratio / (1.0f + ratio);
    }
}
template <bool UNI>
__device__ __forceinline__ void lat_eval_link(const LatFrame& f, int axis, int owner, uint32_t owner_bits,
                                              double4 n0, double4 n1, double4 p0, double4 p1, float prev_dt,
                                              LinkState& st, d3& fN, d3& mN, d3& fP, d3& mP)
{
    double2 ra = __ldg(f.c_rec[axis][0] + owner), rb = __ldg(f.c_rec[axis][1] + owner), rc = __ldg(f.c_rec[axis][2] + owner);
    float4 rs = __ldg(f.c_recf[axis] + owner);
    if (f.c_ps) {                                       // Poisson models: the strains of both end voxels as the step saw them
        const float4 psn = __ldg(f.c_ps + owner), psp = __ldg(f.c_ps + owner + (axis == 0 ? 1 : (axis == 1 ? f.nx : f.nxy)));
        lat_eval_link_rec<UNI>(f, axis, owner_bits, ra, rb, rc, rs, n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP, -1.0f, &psn, &psp);
        return;
    }
    lat_eval_link_rec<UNI>(f, axis, owner_bits, ra, rb, rc, rs, n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP);
}

// closes the bookkeeping of the last step of a vx_step call
__global__ void k_lattice_finish(DevParams* p, int parity_of_last)
{
    if (p->div_flag[parity_of_last] | p->div_latched) p->div_latched = 1;
    else if (p->pending) { p->steps_done += 1; p->time += p->dt; }
    p->last_prev = p->prev_dt;
    if (p->steps_done > 0) p->prev_dt = p->dt;
    p->pending = 0;
}

// copies voxel state between generations (used to re-align generations after a diverged step,
// where links advance but voxels do not; src/Voxelyze.cpp:263-269)
__global__ void k_lattice_copy_voxels(int n, const double4* a0, const double4* a1, const double4* am0, const double2* am1,
                                      double4* b0, double4* b1, double4* bm0, double2* bm1, int keep_dst_lflags)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    double4 q1 = a1[v];
    if (keep_dst_lflags) {      // link mode bits belong to the (advanced) link generation
        uint32_t src = meta_hi(q1.w), dst = meta_hi(b1[v].w);
        q1.w = meta_pack(meta_temp(q1.w), (src & ~(0x3Fu << VM_LFLAG_SHIFT)) | (dst & (0x3Fu << VM_LFLAG_SHIFT)));
    }
    b0[v] = a0[v]; b1[v] = q1; bm0[v] = am0[v]; bm1[v] = am1[v];
}

// state access in lattice mode: link i (internal order) is (owner voxel, axis)
struct LatLinkRef { const int* owner; const unsigned char* axis; };

// what: G_* link field.  Force/moment fields are recomputed from the PREVIOUS generation
// (`prev` = inputs of the last executed step), everything else is read from the current one.
__global__ void k_lattice_gather_links(LatFrame cur, LatFrame prev, int have_prev, float prev_dt_of_last, int what,
                                       const int* e2i, LatLinkRef ref, int first, int count, void* out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int i = e2i[first + k];
    int owner = ref.owner[i], axis = ref.axis[i];
    double* d = (double*)out; float* fl = (float*)out; uint32_t* u32 = (uint32_t*)out;
    if (what == G_FORCE_NEG || what == G_FORCE_POS || what == G_MOMENT_NEG || what == G_MOMENT_POS) {
        d3 r = mk3(0.0, 0.0, 0.0);
        if (have_prev) {
            const int stride = axis == 0 ? 1 : (axis == 1 ? prev.nx : prev.nxy);
            const int pv = owner + stride;
            double4 n0 = prev.c_pose0[owner], n1 = prev.c_pose1[owner], p0 = prev.c_pose0[pv], p1 = prev.c_pose1[pv];
            LinkState st; d3 fN, mN, fP, mP;
            lat_eval_link<false>(prev, axis, owner, meta_hi(n1.w), n0, n1, p0, p1, prev_dt_of_last, st, fN, mN, fP, mP);
            r = what == G_FORCE_NEG ? fN : what == G_FORCE_POS ? fP : what == G_MOMENT_NEG ? mN : mP;
        }
        d[3 * k] = r.x; d[3 * k + 1] = r.y; d[3 * k + 2] = r.z;
        return;
    }
    LinkState st;
    uint32_t obits = meta_hi(cur.c_pose1[owner].w);
    lat_decode(cur.c_rec[axis][0][owner], cur.c_rec[axis][1][owner], cur.c_rec[axis][2][owner], cur.c_recf[axis][owner],
               (obits >> (VM_LFLAG_SHIFT + 2 * axis)) & 3u, st);
    switch (what) {
    case G_POS2: d[3 * k] = st.pos2.x; d[3 * k + 1] = st.pos2.y; d[3 * k + 2] = st.pos2.z; break;
    case G_ANGLE1V: d[3 * k] = st.a1v.x; d[3 * k + 1] = st.a1v.y; d[3 * k + 2] = st.a1v.z; break;
    case G_ANGLE2V: d[3 * k] = st.a2v.x; d[3 * k + 1] = st.a2v.y; d[3 * k + 2] = st.a2v.z; break;
    case G_STRAIN: fl[k] = st.strain; break;
    case G_MAXSTRAIN: fl[k] = st.max_strain; break;
    case G_STRAINOFFSET: fl[k] = st.strain_offset; break;
    case G_STRESS: fl[k] = st.stress; break;
    case G_LINKFLAGS: {
        const int stride = axis == 0 ? 1 : (axis == 1 ? cur.nx : cur.nxy);
        uint32_t pbits = meta_hi(cur.c_pose1[owner + stride].w);
        const DevLinkMat& lm = cur.lmat[cur.pair_lmat[(obits & VM_MAT_MASK) * cur.n_mat + (pbits & VM_MAT_MASK)]];
        u32[k] = (st.small_angle ? 1u : 0u) | (st.vel_valid ? 2u : 0u) | (mat_yielded(lm, st.max_strain) ? 4u : 0u) | (mat_failed(lm, st.max_strain) ? 8u : 0u);
        break; }
    }
}

// persistent link state <- packed records (see k_scatter_link_state); the mode bits go to the owner's meta word.
// One thread per OWNER VOXEL and axis pass (no two threads touch the same meta word).
__global__ void k_lattice_scatter_link_state(double4* pose1, double2* rec, const int* link_of_owner, int n_vox,
                                             const LinkStateRec* src, int first, int count)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox) return;
    uint32_t bits = meta_hi(pose1[v].w);
    bool touched = false;
    for (int axis = 0; axis < 3; axis++) {
        const int ext = link_of_owner[axis * n_vox + v];          // caller index of the link this voxel owns on that axis, or -1
        if (ext < first || ext >= first + count) continue;
        const LinkStateRec r = src[ext - first];
        LinkState st;
        st.small_angle = (r.flags & 1u) != 0; st.vel_valid = (r.flags & 2u) != 0;
        st.pos2 = mk3(r.pos2[0], r.pos2[1], r.pos2[2]); st.a1v = mk3(r.a1v[0], r.a1v[1], r.a1v[2]); st.a2v = mk3(r.a2v[0], r.a2v[1], r.a2v[2]);
        st.strain = r.strain; st.max_strain = r.max_strain; st.strain_offset = r.strain_offset; st.stress = r.stress;
        double2 a, b, c; float4 s4; uint32_t lf;
        lat_encode(st, a, b, c, s4, lf);
        rec[(size_t)(axis * 4 + 0) * n_vox + v] = a; rec[(size_t)(axis * 4 + 1) * n_vox + v] = b; rec[(size_t)(axis * 4 + 2) * n_vox + v] = c;
        reinterpret_cast<float4*>(rec + (size_t)(axis * 4 + 3) * n_vox)[v] = s4;
        bits = (bits & ~(3u << (VM_LFLAG_SHIFT + 2 * axis))) | (lf << (VM_LFLAG_SHIFT + 2 * axis));
        touched = true;
    }
    if (touched) reinterpret_cast<uint32_t*>(&pose1[v].w)[1] = bits;
}

// max over links of a1/min(m1,m2) for the dense lattice (nu = 0 only), same reduction as k_max_freq
__global__ void __launch_bounds__(256) k_lattice_max_freq(LatFrame f, unsigned int* out)
{
    float best = 0.0f;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.n_vox; v += gridDim.x * blockDim.x) {
        uint32_t bits = meta_hi(f.c_pose1[v].w);
        const uint32_t mask = (bits >> VM_LINK_SHIFT) & 0x3Fu;
        const DevVoxMat& vmn = f.vmat[bits & VM_MAT_MASK];
        for (int axis = 0; axis < 3; axis++) {
            if (!(mask & (1u << (2 * axis)))) continue;
            const int stride = axis == 0 ? 1 : (axis == 1 ? f.nx : f.nxy);
            uint32_t pb = meta_hi(f.c_pose1[v + stride].w);
            const DevVoxMat& vmp = f.vmat[pb & VM_MAT_MASK];
            const DevLinkMat& lm = f.lmat[f.pair_lmat[(bits & VM_MAT_MASK) * f.n_mat + (pb & VM_MAT_MASK)]];
            float m1 = vmn.mass, m2 = vmp.mass;
            float stiff = lm.a1;
            if (lm.nu != 0.0f && f.c_ps) {                // CVX_Link::axialStiffness with Poisson coupling, src/VX_Link.cpp:259-267
                const double w = f.c_pose1[v].w, wp = f.c_pose1[v + stride].w;
                const double rest = 0.5 * (vmn.size[axis] * (1 + meta_temp(w) * vmn.cte) + vmp.size[axis] * (1 + meta_temp(wp) * vmp.cte));
                const float area = 0.5f * (transverse_area(vmn, axis, f.c_ps[v]) + transverse_area(vmp, axis, f.c_ps[v + stride]));
                stiff = (float)(lm.e_hat * area / ((f.c_recf[axis][v].x + 1) * rest));
            }
            float f2 = stiff / (m1 < m2 ? m1 : m2);
            if (f2 > best) best = f2;
        }
    }
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best > 0.0f) atomicMax(out, __float_as_uint(best));
}

// CVX_Voxel::pStrain of every voxel from the link strains currently in the records of generation `f.c_*` (Poisson models:
// after a reset, after Poisson's ratio was switched on, after externals changed); written to `out`
__global__ void __launch_bounds__(128) k_lattice_pstrain_init(LatFrame f, float4* out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= f.n_vox) return;
    const uint32_t bits = meta_hi(f.c_pose1[v].w);
    const uint32_t mask = (bits >> VM_LINK_SHIFT) & 0x3Fu;
    const DevVoxMat& vm = f.vmat[bits & VM_MAT_MASK];
    float r[3] = {0.f, 0.f, 0.f}; int nb[3] = {0, 0, 0};
    for (int k = 0; k < 6; k++) {
        if (!(mask & (1u << k))) continue;
        const int axis = k >> 1, stride = axis == 0 ? 1 : (axis == 1 ? f.nx : f.nxy);
        const int owner = (k & 1) ? v - stride : v, other = (k & 1) ? v - stride : v + stride;      // slot odd: this voxel is the positive end
        const float strain = f.c_recf[axis][owner].x;
        const float e_other = f.vmat[meta_hi(f.c_pose1[other].w) & VM_MAT_MASK].E;
        const float ratio = (k & 1) ? vm.E / e_other : e_other / vm.E;                               // E_pos / E_neg
        r[axis] += (k & 1) ? 2.0f * strain * ratio / (1.0f + ratio) : 2.0f * strain / (1.0f + ratio);
        nb[axis]++;
    }
    const DevExt* ext = (bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
    out[v] = voxel_pstrain(vm, ext, r, nb);
}

// =================================================================================================
// k_lattice_warp -- fused step, one WARP per 4 x 4 x 2 brick (32 voxels = 32 lanes).
//
// No block-wide barrier, no idle phase, no dependent global load in the arithmetic: each warp first
// requests everything it will read (link records, the poses just outside the brick, later the
// momenta) with cp.async into its private shared-memory window, then runs four rounds of 32 link
// evaluations and one round of 32 voxel integrations out of shared memory and registers.
//   round H      the 32 links that ENTER the brick through its three negative faces (8 through -X,
//                8 through -Y, 16 through -Z), evaluated from the positive end; the force on the
//                in-brick voxel is parked in shared memory (hslot)
//   rounds 0..2  lane = voxel, evaluates its own +X / +Y / +Z link (owner: stores the new record).
//                The force on the partner goes to the lane that holds it by warp shuffle, so the six contributions of a voxel
//                are accumulated in registers in the reference's order X+ X- Y+ Y- Z+ Z-.
//   last         lane = voxel: integrate, store
// 4.0 link evaluations per voxel (3 + 32/32), all 32 lanes busy in every round of a full brick.
// Shared memory per warp: two record windows 2x4x32x16 B + 72 poses (32 of the brick, 40 just
// outside it) x 64 B + hslot 6x32x8 B = 10 240 B; requests run one round ahead of their use; the round-H windows are re-used for round 2 and the round-0 window for the momenta.
// Requests are predicated on GEOMETRY only (is there a cell on the other side?), which lets them go out before the
// first byte of voxel state has arrived; whether the link exists is decided later from the voxel's link mask
// (a box with holes has cells without voxels: their data is fetched and ignored).
// =================================================================================================
#define VX_WB_X 4
#define VX_WB_Y 4
#define VX_WB_Z 2
#ifndef VX_WB_WARPS
#define VX_WB_WARPS 8                                   // bricks per CTA (consecutive brick ids; ids walk 2x2x2 groups of bricks)
#endif
#define VX_WB_POSES 72
#define VX_WB_WARP_BYTES (2 * 4 * 32 * 16 + 4 * VX_WB_POSES * 16 + 6 * 32 * 8)
#define VX_WB_SMEM (VX_WB_WARPS * VX_WB_WARP_BYTES)
#ifndef VX_WB_MINBLOCKS
#define VX_WB_MINBLOCKS 2
#endif

__device__ __forceinline__ double4 shfl_d4(double4 v, int src)
{
    return make_double4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src),
                        __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool UNI, bool PUSH>
__global__ void __launch_bounds__(32 * VX_WB_WARPS, VX_WB_MINBLOCKS)
k_lattice_warp(LatFrame f, int parity, int first_of_call, int floor_on, int nbx, int nby, int nbz, int gz_off, int book, int grouped)
{
    // grouped: nbx x nby x nbz counts 2x2x2 GROUPS of bricks, nbz layers of them starting at layer gz_off (a whole
    //          step: all layers, gz_off = 0); good L1/L2 locality on large lattices, pads odd brick counts
    // else:    nbx x nby x nbz counts bricks, x fastest, no padding (ensembles of small boxes)
    // book:    this launch does the step bookkeeping (exactly one launch per step does)
    extern __shared__ __align__(16) unsigned char wb_smem[];
    DevParams* p = f.params;
    const int frozen = p->div_flag[parity ^ 1] | p->div_latched;
    const float dt = p->dt;
    const float prev_dt = first_of_call ? p->prev_dt : dt;
    if (book && blockIdx.x == 0 && threadIdx.x == 0) {
        if (frozen) p->div_latched = 1;
        else if (p->pending) { p->steps_done += 1; p->time += dt; }
        if (!frozen) p->pending = 1;
    }
    if (frozen) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = wb_smem + (size_t)warp * VX_WB_WARP_BYTES;
    uint4 (*rec_sh)[4][32] = reinterpret_cast<uint4 (*)[4][32]>(wbase);                       // [window][part][lane]: 0 = H, round 1, momenta; 1 = round 0, round 2
    uint4 (*pose_sh)[VX_WB_POSES] = reinterpret_cast<uint4 (*)[VX_WB_POSES]>(wbase + 2 * 4 * 32 * 16);   // [part][entry]: 0..31 brick, 32..63 H then round 2 (32..47) and round 1 (48..55), 64..71 round 0
    double (*hslot)[32] = reinterpret_cast<double (*)[32]>(wbase + 2 * 4 * 32 * 16 + 4 * VX_WB_POSES * 16);   // [comp][entering link]

    // warp -> brick: consecutive ids walk a 2x2x2 group of bricks, groups x-fastest
    int b = blockIdx.x * VX_WB_WARPS + warp;
    const int w8 = grouped ? b & 7 : 0;
    if (grouped) b >>= 3;
    const int gx = b % nbx; b /= nbx;
    const int gy = b % nby; b /= nby;
    const int gz = gz_off + b % nbz; const int member = b / nbz;
    const int x0 = grouped ? (gx * 2 + (w8 & 1)) * VX_WB_X : gx * VX_WB_X, y0 = grouped ? (gy * 2 + ((w8 >> 1) & 1)) * VX_WB_Y : gy * VX_WB_Y,
              z0 = grouped ? (gz * 2 + (w8 >> 2)) * VX_WB_Z : gz * VX_WB_Z;
    if (member * f.nz * f.nxy >= f.n_vox || x0 >= f.nx || y0 >= f.ny || z0 >= f.nz) return;            // whole warp
    const int vbase = member * f.nz * f.nxy;

    const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
    const int x = x0 + lx, y = y0 + ly, z = z0 + lz;
    const bool has_voxel = x < f.nx && y < f.ny && z < f.nz;
    const int v = vbase + (min(z, f.nz - 1) * f.ny + min(y, f.ny - 1)) * f.nx + min(x, f.nx - 1);   // clamped: idle lanes load valid memory

    // ---- entering link of this lane (round H): axis, in-brick target lane, negative-end voxel
    int h_axis, h_tl;
    if (lane < 8) { h_axis = 0; h_tl = ((lane >> 2) << 4) | ((lane & 3) << 2); }              // target lx = 0: (ly, lz) = (lane&3, lane>>2)
    else if (lane < 16) { h_axis = 1; h_tl = (((lane - 8) >> 2) << 4) | ((lane - 8) & 3); }  // target ly = 0: (lx, lz)
    else { h_axis = 2; h_tl = lane - 16; }                                                    // target lz = 0: (lx, ly)
    const int h_tx = x0 + (h_tl & 3), h_ty = y0 + ((h_tl >> 2) & 3), h_tz = z0 + (h_tl >> 4);
    const int h_stride = h_axis == 0 ? 1 : (h_axis == 1 ? f.nx : f.nxy);
    const bool h_geo = h_tx < f.nx && h_ty < f.ny && h_tz < f.nz && (h_axis == 0 ? x0 : (h_axis == 1 ? y0 : z0)) > 0;
    const int h_vn = vbase + (h_tz * f.ny + h_ty) * f.nx + h_tx - h_stride;

    // ---- requests, one round ahead.  cp.async groups in issue order: H (+ the brick's own poses), round 0,
    //      [after H] round 1 and the round-2 poses, [after round 0] round 2, [after round 1] momenta
    // record arrays of one generation are one allocation: [axis][part][voxel] (vx_capi.cu lat_frame)
    auto c_rec = [&](int a, int k) { return f.c_rec[0][0] + (size_t)(a * 4 + k) * f.n_vox; };
    auto c_recf = [&](int a) { return f.c_rec[0][0] + (size_t)(a * 4 + 3) * f.n_vox; };
    auto request_pose = [&](int entry, int vox) {
        cp_async16(&pose_sh[0][entry], reinterpret_cast<const uint4*>(f.c_pose0 + vox));
        cp_async16(&pose_sh[1][entry], reinterpret_cast<const uint4*>(f.c_pose0 + vox) + 1);
        cp_async16(&pose_sh[2][entry], reinterpret_cast<const uint4*>(f.c_pose1 + vox));
        cp_async16(&pose_sh[3][entry], reinterpret_cast<const uint4*>(f.c_pose1 + vox) + 1);
    };
    auto load_pose = [&](int entry, double4& a, double4& c) {
        const uint4 e0 = pose_sh[0][entry], e1 = pose_sh[1][entry], e2 = pose_sh[2][entry], e3 = pose_sh[3][entry];
        a = make_double4(__hiloint2double(e0.y, e0.x), __hiloint2double(e0.w, e0.z), __hiloint2double(e1.y, e1.x), __hiloint2double(e1.w, e1.z));
        c = make_double4(__hiloint2double(e2.y, e2.x), __hiloint2double(e2.w, e2.z), __hiloint2double(e3.y, e3.x), __hiloint2double(e3.w, e3.z));
    };
    auto ext_entry = [&](int a) { return a == 0 ? 64 + ly + 4 * lz : (a == 1 ? 48 + lx + 4 * lz : 32 + lx + 4 * ly); };
    auto request_link = [&](int a, bool records, bool poses) {
        const int coord = a == 0 ? x : (a == 1 ? y : z), nn = a == 0 ? f.nx : (a == 1 ? f.ny : f.nz);
        const bool inside = a == 0 ? lx < VX_WB_X - 1 : (a == 1 ? ly < VX_WB_Y - 1 : lz < VX_WB_Z - 1);
        if (has_voxel && coord + 1 < nn) {
            if (records) {
#pragma unroll
                for (int k = 0; k < 3; k++) cp_async16(&rec_sh[(a & 1) ^ 1][k][lane], c_rec(a, k) + v);
                cp_async16(&rec_sh[(a & 1) ^ 1][3][lane], c_recf(a) + v);
            }
            if (poses && !inside) request_pose(ext_entry(a), v + (a == 0 ? 1 : (a == 1 ? f.nx : f.nxy)));
        }
    };
    request_pose(lane, v);
    if (h_geo) {
#pragma unroll
        for (int k = 0; k < 3; k++) cp_async16(&rec_sh[0][k][lane], c_rec(h_axis, k) + h_vn);
        cp_async16(&rec_sh[0][3][lane], c_recf(h_axis) + h_vn);
        request_pose(32 + lane, h_vn);
    }
    cp_async_commit();
    request_link(0, true, true);
    cp_async_commit();

    // ---- round H
    cp_async_wait<1>();
    __syncwarp();                     // the brick's poses were requested by other lanes
    const uint32_t bits = pose_sh[3][lane].w;                      // high word of pose1.w: the meta word
    const uint32_t mask = has_voxel ? ((bits >> VM_LINK_SHIFT) & 0x3Fu) : 0u;
    uint32_t new_bits = bits;
    if (h_geo && ((pose_sh[3][h_tl].w >> (VM_LINK_SHIFT + 2 * h_axis + 1)) & 1u)) {
        double4 n0, n1, p0, p1;
        load_pose(32 + lane, n0, n1);
        load_pose(h_tl, p0, p1);
        const uint4 r0 = rec_sh[0][0][lane], r1 = rec_sh[0][1][lane], r2 = rec_sh[0][2][lane], r3 = rec_sh[0][3][lane];
        LinkState st; d3 fN, mN, fP, mP;
        lat_eval_link_rec<UNI>(f, h_axis, meta_hi(n1.w),
                               make_double2(__hiloint2double(r0.y, r0.x), __hiloint2double(r0.w, r0.z)),
                               make_double2(__hiloint2double(r1.y, r1.x), __hiloint2double(r1.w, r1.z)),
                               make_double2(__hiloint2double(r2.y, r2.x), __hiloint2double(r2.w, r2.z)),
                               make_float4(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w)),
                               n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP);
        hslot[0][lane] = fP.x; hslot[1][lane] = fP.y; hslot[2][lane] = fP.z; hslot[3][lane] = mP.x; hslot[4][lane] = mP.y; hslot[5][lane] = mP.z;
    }
    __syncwarp();
    request_link(1, true, true);      // into the windows round H has left
    request_link(2, false, true);
    cp_async_commit();

    // ---- rounds 0..2: own links, forces accumulated in reference order
    d3 F = mk3(0.0, 0.0, 0.0), M = mk3(0.0, 0.0, 0.0);
#pragma unroll 1
    for (int a = 0; a < 3; a++) {
        cp_async_wait<1>();
        const int win = (a & 1) ^ 1;
        const bool inside = a == 0 ? lx < VX_WB_X - 1 : (a == 1 ? ly < VX_WB_Y - 1 : lz < VX_WB_Z - 1);
        const bool first = a == 0 ? lx == 0 : (a == 1 ? ly == 0 : lz == 0);
        const int dl = a == 0 ? 1 : (a == 1 ? 4 : 16);
        d3 fN = mk3(0.0, 0.0, 0.0), mN = fN, fP = fN, mP = fN;
        if ((mask >> (2 * a)) & 1u) {
            double4 n0, n1, p0, p1;
            load_pose(lane, n0, n1);
            load_pose(inside ? lane + dl : ext_entry(a), p0, p1);
            const uint4 r0 = rec_sh[win][0][lane], r1 = rec_sh[win][1][lane], r2 = rec_sh[win][2][lane], r3 = rec_sh[win][3][lane];
            LinkState st;
            lat_eval_link_rec<UNI>(f, a, bits,
                                   make_double2(__hiloint2double(r0.y, r0.x), __hiloint2double(r0.w, r0.z)),
                                   make_double2(__hiloint2double(r1.y, r1.x), __hiloint2double(r1.w, r1.z)),
                                   make_double2(__hiloint2double(r2.y, r2.x), __hiloint2double(r2.w, r2.z)),
                                   make_float4(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w)),
                                   n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP);
            double2 wa, wb, wc; float4 ws; uint32_t lf;
            lat_encode(st, wa, wb, wc, ws, lf);
            double2* nr = f.n_rec[0][0] + (size_t)(a * 4) * f.n_vox + v;       // the twelve record arrays are one allocation
            nr[0] = wa; nr[f.n_vox] = wb; nr[2 * (size_t)f.n_vox] = wc; *reinterpret_cast<float4*>(nr + 3 * (size_t)f.n_vox) = ws;
            new_bits = (new_bits & ~(3u << (VM_LFLAG_SHIFT + 2 * a))) | (lf << (VM_LFLAG_SHIFT + 2 * a));
            if (st.strain > 100) p->div_flag[parity] = 1;          // src/Voxelyze.cpp:265
            F = F + fN; M = M + mN;
        }
        // force on the positive end travels to the lane that holds that voxel
        const int src = (lane - dl) & 31;
        d3 inF = mk3(__shfl_sync(0xffffffffu, fP.x, src), __shfl_sync(0xffffffffu, fP.y, src), __shfl_sync(0xffffffffu, fP.z, src));
        d3 inM = mk3(__shfl_sync(0xffffffffu, mP.x, src), __shfl_sync(0xffffffffu, mP.y, src), __shfl_sync(0xffffffffu, mP.z, src));
        if ((mask >> (2 * a + 1)) & 1u) {
            if (first) {
                const int hl = a == 0 ? ly + 4 * lz : (a == 1 ? 8 + lx + 4 * lz : 16 + lx + 4 * ly);
                inF = mk3(hslot[0][hl], hslot[1][hl], hslot[2][hl]);
                inM = mk3(hslot[3][hl], hslot[4][hl], hslot[5][hl]);
            }
            F = F + inF; M = M + inM;
        }
        // next request into the window this round has left (a lane only ever touches its own column of a window)
        if (a == 0) request_link(2, true, false);
        else if (a == 1 && has_voxel) {
            cp_async16(&rec_sh[0][0][lane], reinterpret_cast<const uint4*>(f.c_mom0 + v));
            cp_async16(&rec_sh[0][1][lane], reinterpret_cast<const uint4*>(f.c_mom0 + v) + 1);
            cp_async16(&rec_sh[0][2][lane], reinterpret_cast<const uint4*>(f.c_mom1 + v));
        }
        if (a < 2) cp_async_commit();
    }

    // ---- last round: one lane per voxel
    cp_async_wait<0>();
    if (!has_voxel) return;
    const uint4 q0 = rec_sh[0][0][lane], q1 = rec_sh[0][1][lane], q2 = rec_sh[0][2][lane];
    double4 s0, s1;
    load_pose(lane, s0, s1);
    VoxelState vs;
    vs.bits = new_bits; vs.temp = f.amb_set ? f.amb : meta_temp(s1.w);
    vs.pos = mk3(s0.x, s0.y, s0.z);
    vs.orient.w = s0.w; vs.orient.x = s1.x; vs.orient.y = s1.y; vs.orient.z = s1.z;
    vs.lin = mk3(__hiloint2double(q0.y, q0.x), __hiloint2double(q0.w, q0.z), __hiloint2double(q1.y, q1.x));
    vs.ang = mk3(__hiloint2double(q1.w, q1.z), __hiloint2double(q2.y, q2.x), __hiloint2double(q2.w, q2.z));
    if (vs.bits & VM_GHOST) {
        // a ghost's pose and temperature arrive from the slab that owns the voxel (vx_halo_import, or the
        // peer's k_halo_push straight into this array, possibly while this kernel runs): only the flag word,
        // which carries the mode bits of the links this ghost owns, is ours to write
        reinterpret_cast<uint32_t*>(&f.n_pose1[v].w)[1] = vs.bits;
        return;
    }
    {
        const DevVoxMat& vm = UNI ? f.vm0 : f.vmat[vs.bits & VM_MAT_MASK];
        const DevExt* ext = (vs.bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
        const int* refs = nullptr; int n_refs = 0;
        if (f.col_slot && ((vs.bits >> VM_LINK_SHIFT) & 0x3Fu) != 0x3Fu) {        // only surface voxels are ever watched
            const int cs = f.col_slot[v];
            refs = f.col_ref + f.col_start[cs];
            n_refs = f.col_start[cs + 1] - f.col_start[cs];
        }
        voxel_integrate(vs, F, M, refs, n_refs, f.col_force, vm, ext, dt, floor_on != 0);
    }
    f.n_pose0[v] = make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w);
    f.n_pose1[v] = make_double4(vs.orient.x, vs.orient.y, vs.orient.z, meta_pack(vs.temp, vs.bits));
    f.n_mom0[v] = make_double4(vs.lin.x, vs.lin.y, vs.lin.z, vs.ang.x);
    f.n_mom1[v] = make_double2(vs.ang.y, vs.ang.z);
    // halo push fused into the step: posted stores over NVLink; the receiver owns the upper half of pose1.w
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (PUSH && z == f.push_z[k]) {
            const int q = y * f.nx + x;
            f.push0[k][q] = make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w);
            double* d = reinterpret_cast<double*>(f.push1[k] + q);
            d[0] = vs.orient.x; d[1] = vs.orient.y; d[2] = vs.orient.z;
            reinterpret_cast<float*>(d + 3)[0] = vs.temp;
        }
    }
}


// GSKIP (see k_lattice_tma): no brick rewrites the flag words (upper half of pose1.w) of the ghost planes any more, so the
// generation a stepping call writes first gets them copied once, from the generation it starts on
__global__ void k_lattice_ghost_words(const double4* __restrict__ c_pose1, double4* __restrict__ n_pose1, int n_lo, int hi_first, int n_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lo + n_hi) return;
    const int v = i < n_lo ? i : hi_first + (i - n_lo);
    reinterpret_cast<uint32_t*>(&n_pose1[v].w)[1] = reinterpret_cast<const uint32_t*>(&c_pose1[v].w)[1];
}

// =================================================================================================
// k_lattice_tma -- the warp-brick step with its staging done by the Tensor Memory Accelerator.
//
// Same arithmetic, same rounds as k_lattice_warp.  What changes is how a brick's inputs reach shared
// memory: instead of ~70 cp.async instructions per lane, ONE lane issues 13 bulk tensor copies
// (cp.async.bulk.tensor, boxes of the lattice arrays described by CUtensorMap descriptors; both pose
// records of a voxel set, and all four parts of a link record, are one 4-D box because their arrays
// share an allocation):
//   group 0 (mbarrier 0)  the brick's own poses; the three faces just outside -X/-Y/-Z (poses and the
//                         records of the links entering through them)                    -> round H
//   group 1 (mbarrier 1)  all twelve record parts of the brick's own links in one box; the +X/+Y faces
//   group 2 (mbarrier 2)  issued after round H into the space its inputs leave: momenta, the +Z face
// Boxes that poke out of the lattice are zero-filled by the hardware, so the requests need no
// predicates and no index arithmetic.  Waiting is mbarrier.try_wait.parity with a bounded spin.
// Shared memory per warp 13 440 B:
//   0      own pose0 [32]x32      1024   | 9216  -X pose0/pose1 [8]x32  2x256  \
//   1024   own pose1              1024   | 9728  -Y pose0/pose1         2x256   > after H: hslot 1536, +Z pose0 512
//   2048   own rec  [12][32]x16   6144   | 10240 -Z pose0/pose1 [16]x32 2x512  /
//          (part 4a+k, k = 3: float4)    | 11264 -X rec [4][8]x16  512 \  after H: +Z pose1 512,
//   8192   +X pose0/pose1 [8]x32  2x256  | 11776 -Y rec            512  > momenta mom0 1024, mom1 512
//   8704   +Y pose0/pose1         2x256  | 12288 -Z rec [4][16]x16 1024 /
//   13312  three mbarriers
// Tensor maps (built on the host, vx_capi.cu build_tensor_maps), u64 elements, per generation:
enum { TM_P_OWN, TM_P_XF, TM_P_YF, TM_P_ZF, TM_M0, TM_M1, TM_REC_OWN, TM_REC_XF, TM_REC_YF, TM_REC_ZF, TM_COUNT };
// =================================================================================================
#define VX_TMA_WARP_BYTES 13440
#define VX_TMA_SMEM (VX_WB_WARPS * VX_TMA_WARP_BYTES)
#define VX_TMA_TABLE_BYTES 6144                         // room for staged material tables behind the warp windows (2 CTAs/SM still fit)

// stores of a step's results: nothing reads them before the next step.  On lattices whose state does not fit the L2 (stream:
// LatFrame::stream_out) that is ~10 GB of other traffic later, and st.global.cs keeps them from pushing the face data that
// neighbouring bricks are about to read out of L2 (256^3: 2.644 -> 2.618 ms/step, profiles/r2_ablation_cache_hints.log);
// smaller lattices keep plain stores, their next step finds the results in L2
__device__ __forceinline__ void st_out(double2* p, const double2& v, bool stream) { if (stream) __stcs(p, v); else *p = v; }
__device__ __forceinline__ void st_out(float4* p, const float4& v, bool stream) { if (stream) __stcs(p, v); else *p = v; }
__device__ __forceinline__ void st_out(double4* p, const double4& v, bool stream)
{
    if (stream) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); __stcs(reinterpret_cast<double2*>(p) + 1, make_double2(v.z, v.w)); }
    else *p = v;
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
// A bulk copy that never lands must fail the launch, not hang the GPU -- but a healthy copy can take arbitrarily long
// under a debugger, compute-sanitizer or an MPS time slice, so the bound is wall time (globaltimer, 20 s), not a poll count.
__device__ __forceinline__ void mbar_wait(uint32_t bar)
{
    unsigned long long t0 = 0;
#pragma unroll 1
    for (unsigned spin = 0;; spin++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(bar) : "memory");
        if (done) return;
        if ((spin & 0xFFFu) == 0xFFFu) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 20000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_4d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// the same copies with an L2 eviction hint: a brick's own link records and momenta are read exactly once per step, so they
// need not stay in L2 behind the poses, which neighbouring bricks read again as faces (VX_REC_EVICT_FIRST)
__device__ __forceinline__ uint64_t l2_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_3d_hint(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_4d_hint(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol) : "memory");
}

// Grid: grouped (large lattices) -> 3-D, one CTA per 2x2x2 group of bricks: blockIdx = (group x, group y, member * nbz + group layer);
//       else (ensembles of small boxes) -> 1-D, eight consecutive bricks per CTA, bricks x-fastest, no padding.
// GSKIP (z-slabs): the ghost planes at the two ends of the slab are not covered by bricks.  The top one never needed any: its
// voxels are not integrated and the links that reach it are owned from below.  The bottom one owns the +Z links into the first
// owned plane; those are evaluated in round H of the bricks above it anyway (as entering links), so with GSKIP that lane also
// stores the link's new record and mode bits on the ghost's behalf -- the same values the ghost's own brick would have stored.
template <bool UNI, bool PUSH, bool POISSON = false, bool GSKIP = false>
__global__ void __launch_bounds__(32 * VX_WB_WARPS, VX_WB_MINBLOCKS)
k_lattice_tma(LatFrame f, const unsigned char* tmaps, int parity, int first_of_call, int floor_on, int nbx, int nby, int nbz, int gz_off, int book, int grouped, int stage_tables)
{
    // nbx/nby/nbz, gz_off, book, grouped, PUSH: as in k_lattice_warp
    // stage_tables (multi-material models): the CTA copies the material tables into its shared memory behind the warp windows
    extern __shared__ __align__(128) unsigned char tma_smem[];
    DevParams* p = f.params;
    // the step scalars are requested first and looked at only after the bulk copies are on their way
    const int div_prev = p->div_flag[parity ^ 1], div_latched = p->div_latched;
    const float dt = p->dt, prev_dt_call = p->prev_dt;

    // the warp index through a shuffle: the compiler then knows it (and the brick origin, the shared-memory window, the
    // tensor coordinates) to be warp-uniform and issues the bulk copies from uniform registers without a per-lane loop
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    unsigned char* wbase = tma_smem + (size_t)warp * VX_TMA_WARP_BYTES;
    const uint32_t sb = smem_u32(wbase);
    const uint32_t bar0 = sb + 13312, bar1 = sb + 13320, bar2 = sb + 13328;

    int x0, y0, z0, member;
    if (grouped && f.groups) {                          // brick-group list of a sparse body
        const int gid = __ldg(f.groups + blockIdx.x);
        member = 0;
        x0 = ((gid & 1023) * 2 + (warp & 1)) * VX_WB_X; y0 = (((gid >> 10) & 1023) * 2 + ((warp >> 1) & 1)) * VX_WB_Y; z0 = ((gid >> 20) * 2 + (warp >> 2)) * VX_WB_Z;
    } else if (grouped) {
        const unsigned zz = blockIdx.z;
        member = gridDim.z > (unsigned)nbz ? (int)(zz / (unsigned)nbz) : 0;
        const int gz = gz_off + (int)zz - member * nbz;
        x0 = ((int)blockIdx.x * 2 + (warp & 1)) * VX_WB_X; y0 = ((int)blockIdx.y * 2 + ((warp >> 1) & 1)) * VX_WB_Y; z0 = (gz * 2 + (warp >> 2)) * VX_WB_Z + (GSKIP ? f.z_lo : 0);
    } else {
        unsigned b = blockIdx.x * VX_WB_WARPS + warp;
        const unsigned gx = b % (unsigned)nbx; b /= (unsigned)nbx;
        const unsigned gy = b % (unsigned)nby; b /= (unsigned)nby;
        const unsigned gz = b % (unsigned)nbz; member = (int)(b / (unsigned)nbz);
        x0 = (int)gx * VX_WB_X; y0 = (int)gy * VX_WB_Y; z0 = (gz_off + (int)gz) * VX_WB_Z + (GSKIP ? f.z_lo : 0);
    }
    const bool no_brick = member * f.nz * f.nxy >= f.n_vox || x0 >= f.nx || y0 >= f.ny || z0 >= (GSKIP ? f.z_hi : f.nz);            // whole warp
    const int vbase = member * f.nz * f.nxy;
    const int Z0 = member * f.nz + z0;             // the tensors see the members stacked along z

    // ---- one lane arms the barriers and issues every copy of groups 0 and 1
    if (!no_brick && elect_one()) {
        mbar_init(bar0); mbar_init(bar1); mbar_init(bar2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const unsigned char* tm = tmaps + (size_t)parity * TM_COUNT * 128;
        auto map = [&](int k) { return tm + k * 128; };
        mbar_expect(bar0, 2048 + 2048 + 2048);
        tma_4d(sb + 0, map(TM_P_OWN), bar0, 4 * x0, y0, Z0, 0);
        tma_4d(sb + 9216, map(TM_P_XF), bar0, 4 * (x0 - 1), y0, Z0, 0);
        tma_4d(sb + 9728, map(TM_P_YF), bar0, 4 * x0, y0 - 1, Z0, 0);
        tma_4d(sb + 10240, map(TM_P_ZF), bar0, 4 * x0, y0, Z0 - 1, 0);
        tma_4d(sb + 11264, map(TM_REC_XF), bar0, 2 * (x0 - 1), y0, Z0, 0);
        tma_4d(sb + 11776, map(TM_REC_YF), bar0, 2 * x0, y0 - 1, Z0, 4);
        tma_4d(sb + 12288, map(TM_REC_ZF), bar0, 2 * x0, y0, Z0 - 1, 8);
        mbar_expect(bar1, 6144 + 1024);
#ifdef VX_REC_EVICT_FIRST
        tma_4d_hint(sb + 2048, map(TM_REC_OWN), bar1, 2 * x0, y0, Z0, 0, l2_evict_first());
#else
        tma_4d(sb + 2048, map(TM_REC_OWN), bar1, 2 * x0, y0, Z0, 0);
#endif
        tma_4d(sb + 8192, map(TM_P_XF), bar1, 4 * (x0 + VX_WB_X), y0, Z0, 0);
        tma_4d(sb + 8704, map(TM_P_YF), bar1, 4 * x0, y0 + VX_WB_Y, Z0, 0);
    }
    __syncwarp();
    const int frozen = div_prev | div_latched;
    const float prev_dt = first_of_call ? prev_dt_call : dt;
    const float damp_u = UNI ? f.vm0.two_sqrtm_zeta / prev_dt : -1.0f;
    // ---- multi-material models: material tables into shared memory (all 256 threads, while the bulk copies are in flight)
    MatView tables = mat_view(f);
    if (!UNI && stage_tables) {
        unsigned char* tb = tma_smem + VX_TMA_SMEM;
        const int nv = f.n_mat, nl = f.n_lmat;
        DevVoxMat* s_vm = reinterpret_cast<DevVoxMat*>(tb);
        DevLinkMat* s_lm = reinterpret_cast<DevLinkMat*>(tb + nv * sizeof(DevVoxMat));
        float* s_damp = reinterpret_cast<float*>(tb + nv * sizeof(DevVoxMat) + nl * sizeof(DevLinkMat));
        uint16_t* s_pair = reinterpret_cast<uint16_t*>(s_damp + nv);
        const int w_vm = nv * (int)(sizeof(DevVoxMat) / 8), w_lm = nl * (int)(sizeof(DevLinkMat) / 8);
        for (int i = threadIdx.x; i < w_vm; i += blockDim.x) reinterpret_cast<double*>(s_vm)[i] = reinterpret_cast<const double*>(f.vmat)[i];
        for (int i = threadIdx.x; i < w_lm; i += blockDim.x) reinterpret_cast<double*>(s_lm)[i] = reinterpret_cast<const double*>(f.lmat)[i];
        for (int i = threadIdx.x; i < nv * nv; i += blockDim.x) s_pair[i] = f.pair_lmat[i];
        for (int i = threadIdx.x; i < nv; i += blockDim.x) s_damp[i] = f.vmat[i].two_sqrtm_zeta / prev_dt;
        __syncthreads();                                // once per CTA, before any warp leaves
        tables.vmat = s_vm; tables.lmat = s_lm; tables.pair = s_pair; tables.damp = s_damp;
    }
    if (book && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
        if (frozen) p->div_latched = 1;
        else if (p->pending) { p->steps_done += 1; p->time += dt; }
        if (!frozen) p->pending = 1;
    }
    if (no_brick) return;
    if (frozen) { mbar_wait(bar0); mbar_wait(bar1); return; }      // the copies must have landed before the CTA's shared memory is released

    const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
    const int x = x0 + lx, y = y0 + ly, z = z0 + lz;
    const bool has_voxel = x < f.nx && y < f.ny && z < f.nz;
    const int v = vbase + (min(z, f.nz - 1) * f.ny + min(y, f.ny - 1)) * f.nx + min(x, f.nx - 1);

    // POISSON: the Poisson strains the step reads travel outside the staged boxes (plain loads, L2-resident neighbours): this
    // lane's voxel now, partners outside the brick when their link is evaluated; pe[k] collects the per-end axial strain of
    // the voxel's link in slot k for the pStrain of the next step
    float4 ps_own = make_float4(0.f, 0.f, 0.f, 0.f);
    float pe[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (POISSON && has_voxel) ps_own = __ldg(f.c_ps + v);

    auto load_pose = [&](int off0, int off1, int entry, double4& a, double4& c) {        // two 32-byte records of one voxel
        const uint4* q0 = reinterpret_cast<const uint4*>(wbase + off0 + entry * 32);
        const uint4* q1 = reinterpret_cast<const uint4*>(wbase + off1 + entry * 32);
        const uint4 e0 = q0[0], e1 = q0[1], e2 = q1[0], e3 = q1[1];
        a = make_double4(__hiloint2double(e0.y, e0.x), __hiloint2double(e0.w, e0.z), __hiloint2double(e1.y, e1.x), __hiloint2double(e1.w, e1.z));
        c = make_double4(__hiloint2double(e2.y, e2.x), __hiloint2double(e2.w, e2.z), __hiloint2double(e3.y, e3.x), __hiloint2double(e3.w, e3.z));
    };
    auto meta_of = [&](int entry) { return reinterpret_cast<const uint32_t*>(wbase + 1024 + entry * 32)[7]; };   // high word of own pose1.w

    // ---- entering link of this lane (round H)
    int h_axis, h_tl, h_pose, h_rec, h_part;                     // face regions: poses, records (part stride)
    if (lane < 8) { h_axis = 0; h_tl = ((lane >> 2) << 4) | ((lane & 3) << 2); h_pose = 9216 + lane * 32; h_rec = 11264 + lane * 16; h_part = 128; }
    else if (lane < 16) { h_axis = 1; h_tl = (((lane - 8) >> 2) << 4) | ((lane - 8) & 3); h_pose = 9728 + (lane - 8) * 32; h_rec = 11776 + (lane - 8) * 16; h_part = 128; }
    else { h_axis = 2; h_tl = lane - 16; h_pose = 10240 + (lane - 16) * 32; h_rec = 12288 + (lane - 16) * 16; h_part = 256; }
    const int h_face = h_axis == 2 ? 512 : 256;                  // pose1 of a face follows its pose0

    mbar_wait(bar0);
    const uint32_t bits = meta_of(lane);
    const uint32_t mask = has_voxel ? ((bits >> VM_LINK_SHIFT) & 0x3Fu) : 0u;
    uint32_t new_bits = bits;
    d3 hF = mk3(0.0, 0.0, 0.0), hM = hF;
    float h_end[2] = {0.f, 0.f};
    // zero-filled (out of range) voxels carry no link bits; the z test keeps the next ensemble member's planes out
    const bool h_active = x0 + (h_tl & 3) < f.nx && y0 + ((h_tl >> 2) & 3) < f.ny && z0 + (h_tl >> 4) < f.nz &&
                          ((meta_of(h_tl) >> (VM_LINK_SHIFT + 2 * h_axis + 1)) & 1u);
    float4 h_psp = make_float4(0.f, 0.f, 0.f, 0.f), h_psn = h_psp;
    if (POISSON) {                                     // positive end: the in-brick voxel h_tl (another lane's ps_own); negative end: outside
        h_psp = make_float4(__shfl_sync(0xffffffffu, ps_own.x, h_tl), __shfl_sync(0xffffffffu, ps_own.y, h_tl), __shfl_sync(0xffffffffu, ps_own.z, h_tl), 0.f);
        if (h_active) {
            const int h_stride = h_axis == 0 ? 1 : (h_axis == 1 ? f.nx : f.nxy);
            h_psn = __ldg(f.c_ps + vbase + ((z0 + (h_tl >> 4)) * f.ny + y0 + ((h_tl >> 2) & 3)) * f.nx + x0 + (h_tl & 3) - h_stride);
        }
    }
    if (h_active) {
        double4 n0, n1, p0, p1;
        load_pose(h_pose, h_pose + h_face, 0, n0, n1);
        load_pose(0, 1024, h_tl, p0, p1);
        const uint4 r0 = *reinterpret_cast<const uint4*>(wbase + h_rec), r1 = *reinterpret_cast<const uint4*>(wbase + h_rec + h_part),
                    r2 = *reinterpret_cast<const uint4*>(wbase + h_rec + 2 * h_part), r3 = *reinterpret_cast<const uint4*>(wbase + h_rec + 3 * h_part);
        LinkState st; d3 fN, mN;
        lat_eval_link_rec<UNI>(f, h_axis, meta_hi(n1.w),
                               make_double2(__hiloint2double(r0.y, r0.x), __hiloint2double(r0.w, r0.z)),
                               make_double2(__hiloint2double(r1.y, r1.x), __hiloint2double(r1.w, r1.z)),
                               make_double2(__hiloint2double(r2.y, r2.x), __hiloint2double(r2.w, r2.z)),
                               make_float4(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w)),
                               n0, n1, p0, p1, prev_dt, st, fN, mN, hF, hM, damp_u,
                               POISSON ? &h_psn : nullptr, POISSON ? &h_psp : nullptr, POISSON ? h_end : nullptr, &tables);
        if (GSKIP && h_axis == 2 && (meta_hi(n1.w) & VM_GHOST)) {     // the link's owner is a ghost below the bricks: keep its record
            const int vg = vbase + ((z0 - 1) * f.ny + y0 + ((h_tl >> 2) & 3)) * f.nx + x0 + (h_tl & 3);
            double2 wa, wb, wc; float4 ws; uint32_t lf;
            lat_encode(st, wa, wb, wc, ws, lf);
            double2* nr = f.n_rec[0][0] + (size_t)8 * f.n_vox + vg;
            nr[0] = wa; nr[f.n_vox] = wb; nr[2 * (size_t)f.n_vox] = wc; *reinterpret_cast<float4*>(nr + 3 * (size_t)f.n_vox) = ws;
            reinterpret_cast<uint32_t*>(&f.n_pose1[vg].w)[1] = (meta_hi(n1.w) & ~(3u << (VM_LFLAG_SHIFT + 4))) | (lf << (VM_LFLAG_SHIFT + 4));
            if (st.strain > 100) p->div_flag[parity] = 1;
        }
    }
    __syncwarp();                          // every lane has read its round-H inputs: their space is re-used now
    double (*hslot)[32] = reinterpret_cast<double (*)[32]>(wbase + 9216);                 // [comp][entering link]
    if (h_active) { hslot[0][lane] = hF.x; hslot[1][lane] = hF.y; hslot[2][lane] = hF.z; hslot[3][lane] = hM.x; hslot[4][lane] = hM.y; hslot[5][lane] = hM.z; }
    // POISSON: the entering link's strain at its positive end stays in a register of the lane that evaluated it (h_end[1], zero
    // where there is no such link) and is fetched by shuffle when the voxel it enters adds its X- / Y- / Z- slot
    if (elect_one()) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const unsigned char* tm = tmaps + (size_t)parity * TM_COUNT * 128;
        mbar_expect(bar2, 1024 + 1024 + 512);
        tma_4d(sb + 10752, tm + TM_P_ZF * 128, bar2, 4 * x0, y0, Z0 + VX_WB_Z, 0);
#ifdef VX_REC_EVICT_FIRST
        const uint64_t once = l2_evict_first();
        tma_3d_hint(sb + 11776, tm + TM_M0 * 128, bar2, 4 * x0, y0, Z0, once);
        tma_3d_hint(sb + 12800, tm + TM_M1 * 128, bar2, 2 * x0, y0, Z0, once);
#else
        tma_3d(sb + 11776, tm + TM_M0 * 128, bar2, 4 * x0, y0, Z0);
        tma_3d(sb + 12800, tm + TM_M1 * 128, bar2, 2 * x0, y0, Z0);
#endif
    }
    __syncwarp();
    mbar_wait(bar1);

    // ---- rounds 0..2: own links, forces accumulated in reference order
    d3 F = mk3(0.0, 0.0, 0.0), M = mk3(0.0, 0.0, 0.0);
#pragma unroll 1
    for (int a = 0; a < 3; a++) {
        if (a == 2) mbar_wait(bar2);
        const bool inside = a == 0 ? lx < VX_WB_X - 1 : (a == 1 ? ly < VX_WB_Y - 1 : lz < VX_WB_Z - 1);
        const bool first = a == 0 ? lx == 0 : (a == 1 ? ly == 0 : lz == 0);
        const int dl = a == 0 ? 1 : (a == 1 ? 4 : 16);
#ifdef VX_ZERO_PARTNER_FORCE
        d3 fN = mk3(0.0, 0.0, 0.0), mN = fN, fP = fN, mP = fN;
#else
        // a lane without this link leaves fP/mP undefined: the lane that would receive them tests its own link bit first
        d3 fN, mN, fP, mP;
        asm volatile("" : "=d"(fP.x), "=d"(fP.y), "=d"(fP.z), "=d"(mP.x), "=d"(mP.y), "=d"(mP.z));
#endif
        float4 psp = make_float4(0.f, 0.f, 0.f, 0.f);
        float end_s[2] = {0.f, 0.f};
        if (POISSON) {                                 // partner's Poisson strains: another lane's, or a plain load across the + face
            const int pl = (lane + dl) & 31;
            psp = make_float4(__shfl_sync(0xffffffffu, ps_own.x, pl), __shfl_sync(0xffffffffu, ps_own.y, pl), __shfl_sync(0xffffffffu, ps_own.z, pl), 0.f);
            if (!inside && ((mask >> (2 * a)) & 1u)) psp = __ldg(f.c_ps + v + (a == 0 ? 1 : (a == 1 ? f.nx : f.nxy)));
        }
        if ((mask >> (2 * a)) & 1u) {
            double4 n0, n1, p0, p1;
            load_pose(0, 1024, lane, n0, n1);
            {   // partner pose: one address computation, one set of loads (in the brick, or in the +X / +Y / +Z face region)
                const int pose0_at = inside ? (lane + dl) * 32 : (a == 0 ? 8192 + (lz * 4 + ly) * 32 : (a == 1 ? 8704 + (lz * 4 + lx) * 32 : 10752 + (ly * 4 + lx) * 32));
                const int pose1_by = inside ? 1024 : (a == 2 ? 512 : 256);
                load_pose(pose0_at, pose0_at + pose1_by, 0, p0, p1);
            }
            const uint4* rr = reinterpret_cast<const uint4*>(wbase + 2048 + a * 4 * 512) + lane;
            const uint4 r0 = rr[0], r1 = rr[32], r2 = rr[64], r3 = rr[96];
            LinkState st;
            lat_eval_link_rec<UNI>(f, a, bits,
                                   make_double2(__hiloint2double(r0.y, r0.x), __hiloint2double(r0.w, r0.z)),
                                   make_double2(__hiloint2double(r1.y, r1.x), __hiloint2double(r1.w, r1.z)),
                                   make_double2(__hiloint2double(r2.y, r2.x), __hiloint2double(r2.w, r2.z)),
                                   make_float4(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w)),
                                   n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP, damp_u,
                                   POISSON ? &ps_own : nullptr, POISSON ? &psp : nullptr, POISSON ? end_s : nullptr, &tables);
            double2 wa, wb, wc; float4 ws; uint32_t lf;
            lat_encode(st, wa, wb, wc, ws, lf);
            double2* nr = f.n_rec[0][0] + (size_t)(a * 4) * f.n_vox + v;
            st_out(nr, wa, f.stream_out); st_out(nr + f.n_vox, wb, f.stream_out); st_out(nr + 2 * (size_t)f.n_vox, wc, f.stream_out); st_out(reinterpret_cast<float4*>(nr + 3 * (size_t)f.n_vox), ws, f.stream_out);
            new_bits = (new_bits & ~(3u << (VM_LFLAG_SHIFT + 2 * a))) | (lf << (VM_LFLAG_SHIFT + 2 * a));
            if (st.strain > 100) p->div_flag[parity] = 1;          // src/Voxelyze.cpp:265
            F = F + fN; M = M + mN;
        }
        const int src = (lane - dl) & 31;
        d3 inF = mk3(__shfl_sync(0xffffffffu, fP.x, src), __shfl_sync(0xffffffffu, fP.y, src), __shfl_sync(0xffffffffu, fP.z, src));
        d3 inM = mk3(__shfl_sync(0xffffffffu, mP.x, src), __shfl_sync(0xffffffffu, mP.y, src), __shfl_sync(0xffffffffu, mP.z, src));
        float in_e = 0.f;
        if (POISSON) {
            const int hl_ = a == 0 ? ly + 4 * lz : (a == 1 ? 8 + lx + 4 * lz : 16 + lx + 4 * ly);
            const float from_brick = __shfl_sync(0xffffffffu, end_s[1], src), from_h = __shfl_sync(0xffffffffu, h_end[1], hl_);
            in_e = first ? from_h : from_brick;
            pe[2 * a] = end_s[0];
        }
        if ((mask >> (2 * a + 1)) & 1u) {
            if (first) {
                const int hl = a == 0 ? ly + 4 * lz : (a == 1 ? 8 + lx + 4 * lz : 16 + lx + 4 * ly);
                inF = mk3(hslot[0][hl], hslot[1][hl], hslot[2][hl]);
                inM = mk3(hslot[3][hl], hslot[4][hl], hslot[5][hl]);
            }
            F = F + inF; M = M + inM;
            if (POISSON) pe[2 * a + 1] = in_e;
        }
    }

    // ---- last round: one lane per voxel
    if (!has_voxel) return;
    double4 s0, s1;
    load_pose(0, 1024, lane, s0, s1);
    const uint4* mq = reinterpret_cast<const uint4*>(wbase + 11776 + lane * 32);
    const uint4 q0 = mq[0], q1 = mq[1], q2 = *reinterpret_cast<const uint4*>(wbase + 12800 + lane * 16);
    VoxelState vs;
    vs.bits = new_bits; vs.temp = f.amb_set ? f.amb : meta_temp(s1.w);
    vs.pos = mk3(s0.x, s0.y, s0.z);
    vs.orient.w = s0.w; vs.orient.x = s1.x; vs.orient.y = s1.y; vs.orient.z = s1.z;
    vs.lin = mk3(__hiloint2double(q0.y, q0.x), __hiloint2double(q0.w, q0.z), __hiloint2double(q1.y, q1.x));
    vs.ang = mk3(__hiloint2double(q1.w, q1.z), __hiloint2double(q2.y, q2.x), __hiloint2double(q2.w, q2.z));
    if (vs.bits & VM_GHOST) { reinterpret_cast<uint32_t*>(&f.n_pose1[v].w)[1] = vs.bits; return; }   // see k_lattice_warp
    {
        const DevVoxMat& vm = UNI ? f.vm0 : tables.vmat[vs.bits & VM_MAT_MASK];
        const DevExt* ext = (vs.bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
        const int* refs = nullptr; int n_refs = 0;
        if (f.col_slot && ((vs.bits >> VM_LINK_SHIFT) & 0x3Fu) != 0x3Fu) {        // only surface voxels are ever watched
            const int cs = f.col_slot[v];
            refs = f.col_ref + f.col_start[cs];
            n_refs = f.col_start[cs + 1] - f.col_start[cs];
        }
        if (POISSON) {                                  // CVX_Voxel::pStrain for the NEXT step, from this step's link strains
            // a fully fixed voxel returns from CVX_Voxel::timeStep before its cache is invalidated (src/VX_Voxel.cpp:167-172,
            // 231): it keeps the Poisson strain it had
            if (!(ext && (ext->dof & 0x3Fu) == 0x3Fu)) {
                float r[3] = {0.f, 0.f, 0.f}; int nb[3] = {0, 0, 0};
                const uint32_t lm6 = (vs.bits >> VM_LINK_SHIFT) & 0x3Fu;
#pragma unroll
                for (int k = 0; k < 6; k++) if (lm6 & (1u << k)) { r[k >> 1] += pe[k]; nb[k >> 1]++; }
                ps_own = voxel_pstrain(vm, ext, r, nb);
            }
            f.n_ps[v] = ps_own;
        }
        voxel_integrate(vs, F, M, refs, n_refs, f.col_force, vm, ext, dt, floor_on != 0);
    }
    st_out(f.n_pose0 + v, make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w), f.stream_out);
    st_out(f.n_pose1 + v, make_double4(vs.orient.x, vs.orient.y, vs.orient.z, meta_pack(vs.temp, vs.bits)), f.stream_out);
    st_out(f.n_mom0 + v, make_double4(vs.lin.x, vs.lin.y, vs.lin.z, vs.ang.x), f.stream_out);
    st_out(f.n_mom1 + v, make_double2(vs.ang.y, vs.ang.z), f.stream_out);
    // halo push fused into the step (z-slab runs): posted stores over NVLink; the receiver owns the upper half of pose1.w
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (PUSH && z == f.push_z[k]) {
            const int q = y * f.nx + x;
            f.push0[k][q] = make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w);
            double* d = reinterpret_cast<double*>(f.push1[k] + q);
            d[0] = vs.orient.x; d[1] = vs.orient.y; d[2] = vs.orient.z;
            reinterpret_cast<float*>(d + 3)[0] = vs.temp;
            if (POISSON) f.push_ps[k][q] = ps_own;       // the ghost copy's Poisson strain for the next step
        }
    }
}

} // namespace vxd

