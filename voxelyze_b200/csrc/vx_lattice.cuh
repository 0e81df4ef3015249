// vx_lattice.cuh -- fused single-pass step for dense box lattices (the headline path).
//
// One kernel per doTimeStep: every thread owns one voxel of a full nx*ny*nz box (x fastest),
// evaluates the (up to) six links that touch it, sums their forces in the reference's slot order
// X+,X-,Y+,Y-,Z+,Z- and integrates the voxel -- link forces never travel through HBM.
//   * The voxel reads only OLD poses (its own and its six face neighbours'), which is exactly
//     what the reference's "all links, then all voxels" order does (src/Voxelyze.cpp:251-284),
//     so state is kept in ping-pong buffers: read `cur`, write `nxt`, swap per step.
//   * A link is owned by its negative-end voxel (its +X/+Y/+Z link).  The owner thread advances
//     the link's persistent state; the positive-end voxel's thread re-evaluates the same link
//     from the same old inputs (identical bits) just to obtain its own force -- 2x the FP64 work
//     in exchange for ~45 % of the memory traffic of the two-kernel path.
//   * Implicit indexing: neighbour = v +- {1, nx, nx*ny}; no index arrays are read.
//
// HBM traffic per voxel per step (nu = 0):  read 64 B pose + 48 B momenta + 3 x 64 B link records,
// write the same = 608 B, vs. 216 + 3 x 268 = 1020 B "algorithmic" bytes of SURVEY.md section 8d (the
// 96 B/link force write of the reference layout is not needed: forces of the last step can be
// recomputed on demand from the two buffer generations, see k_lattice_link_forces).
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

struct LatFrame {
    int nx, ny, nz, nxy, n_vox, n_mat;
    // voxel state, current (read) and next (write) generation
    const double4* c_pose0; const double4* c_pose1; const double4* c_mom0; const double2* c_mom1;
    double4* n_pose0; double4* n_pose1; double4* n_mom0; double2* n_mom1;
    // link records owned by voxel v for axis a: rec[a][0..2][v] (double2 x3) + recf[a][v] (float4)
    const double2* c_rec[3][3]; const float4* c_recf[3];
    double2* n_rec[3][3]; float4* n_recf[3];
    const int* ext_idx;
    const DevVoxMat* vmat; const DevLinkMat* lmat; const float* curve_e; const float* curve_s;
    const uint16_t* pair_lmat;      // [n_mat][n_mat] -> link material
    const DevExt* ext;
    DevParams* params;
};

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
__device__ __forceinline__ void lat_eval_link(const LatFrame& f, int axis, int owner, uint32_t owner_bits,
                                              double4 n0, double4 n1, double4 p0, double4 p1, float prev_dt,
                                              LinkState& st, d3& fN, d3& mN, d3& fP, d3& mP)
{
    const uint32_t hn = meta_hi(n1.w), hp = meta_hi(p1.w);
    const DevVoxMat& vmn = f.vmat[hn & VM_MAT_MASK];
    const DevVoxMat& vmp = f.vmat[hp & VM_MAT_MASK];
    const DevLinkMat lm = f.lmat[f.pair_lmat[(hn & VM_MAT_MASK) * f.n_mat + (hp & VM_MAT_MASK)]];
    double2 ra = __ldg(f.c_rec[axis][0] + owner), rb = __ldg(f.c_rec[axis][1] + owner), rc = __ldg(f.c_rec[axis][2] + owner);
    float4 rs = __ldg(f.c_recf[axis] + owner);
    lat_decode(ra, rb, rc, rs, (owner_bits >> (VM_LFLAG_SHIFT + 2 * axis)) & 3u, st);
    // CVX_Link::updateRestLength (src/VX_Link.cpp:137-140)
    double rest = 0.5 * (vmn.size[axis] * (1 + meta_temp(n1.w) * vmn.cte) + vmp.size[axis] * (1 + meta_temp(p1.w) * vmp.cte));
    float t_area = 0.5f * (vmn.nom_f * vmn.nom_f + vmp.nom_f * vmp.nom_f);
    float damp_n = vmn.two_sqrtm_zeta / prev_dt, damp_p = vmp.two_sqrtm_zeta / prev_dt;
    q4 on, op;
    on.w = n0.w; on.x = n1.x; on.y = n1.y; on.z = n1.z;
    op.w = p0.w; op.x = p1.x; op.y = p1.y; op.z = p1.z;
    link_forces(axis, mk3(n0.x, n0.y, n0.z), on, mk3(p0.x, p0.y, p0.z), op, rest, t_area, 0.0f,
                damp_n, damp_p, lm, f.curve_e, f.curve_s, st, fN, mN, fP, mP);
}

// dt lives in device memory (p->dt) so that captured graphs survive a change of time step.
// first_of_call: the first step of a vx_step call damps with the previous call's dt
// (CVX_Voxel::previousDt), all later steps of the call with dt itself.
__global__ void __launch_bounds__(128) k_lattice_step(LatFrame f, int parity, int first_of_call, int floor_on)
{
    DevParams* p = f.params;
    const int frozen = p->div_flag[parity ^ 1] | p->div_latched;   // did the previous step diverge?
    const float dt = p->dt;
    const float prev_dt = first_of_call ? p->prev_dt : dt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {                     // bookkeeping of the previous step
        if (frozen) p->div_latched = 1;
        else if (p->pending) { p->steps_done += 1; p->time += dt; }
        if (!frozen) p->pending = 1;
    }
    if (frozen) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= f.n_vox) return;

    const double4 s0 = ld4(f.c_pose0 + v), s1 = ld4(f.c_pose1 + v);
    VoxelState vs;
    vs.bits = meta_hi(s1.w);
    vs.temp = meta_temp(s1.w);
    const uint32_t mask = (vs.bits >> VM_LINK_SHIFT) & 0x3Fu;
    uint32_t new_bits = vs.bits;

    d3 F = mk3(0.0, 0.0, 0.0), M = mk3(0.0, 0.0, 0.0);
#pragma unroll 1
    for (int k = 0; k < 6; k++) {                                  // slot order = reference summation order
        if (!(mask & (1u << k))) continue;
        const int axis = k >> 1;
        const bool i_am_neg = (k & 1) == 0;
        const int stride = axis == 0 ? 1 : (axis == 1 ? f.nx : f.nxy);
        const int u = i_am_neg ? v + stride : v - stride;
        const double4 u0 = ld4(f.c_pose0 + u), u1 = ld4(f.c_pose1 + u);
        LinkState st;
        d3 fN, mN, fP, mP;
        // one inlined copy of the link physics serves both roles: select the operands first
        const double4 n0 = i_am_neg ? s0 : u0, n1 = i_am_neg ? s1 : u1;
        const double4 p0 = i_am_neg ? u0 : s0, p1 = i_am_neg ? u1 : s1;
        lat_eval_link(f, axis, i_am_neg ? v : u, i_am_neg ? vs.bits : meta_hi(u1.w), n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP);
        if (i_am_neg) {
            F = F + fN; M = M + mN;
            double2 ra, rb, rc; float4 rs; uint32_t lf;
            lat_encode(st, ra, rb, rc, rs, lf);
            f.n_rec[axis][0][v] = ra; f.n_rec[axis][1][v] = rb; f.n_rec[axis][2][v] = rc; f.n_recf[axis][v] = rs;
            new_bits = (new_bits & ~(3u << (VM_LFLAG_SHIFT + 2 * axis))) | (lf << (VM_LFLAG_SHIFT + 2 * axis));
            if (st.strain > 100) p->div_flag[parity] = 1;          // src/Voxelyze.cpp:265
        } else {
            F = F + fP; M = M + mP;
        }
    }

    double4 m0 = f.c_mom0[v]; double2 m1 = f.c_mom1[v];
    vs.bits = new_bits;
    vs.pos = mk3(s0.x, s0.y, s0.z);
    vs.orient.w = s0.w; vs.orient.x = s1.x; vs.orient.y = s1.y; vs.orient.z = s1.z;
    vs.lin = mk3(m0.x, m0.y, m0.z);
    vs.ang = mk3(m0.w, m1.x, m1.y);
    if (!(vs.bits & VM_GHOST)) {
        const DevVoxMat& vm = f.vmat[vs.bits & VM_MAT_MASK];
        const DevExt* ext = (vs.bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
        voxel_integrate(vs, F, M, mk3(0.0, 0.0, 0.0), false, vm, ext, dt, floor_on != 0);
    }
    f.n_pose0[v] = make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w);
    f.n_pose1[v] = make_double4(vs.orient.x, vs.orient.y, vs.orient.z, meta_pack(vs.temp, vs.bits));
    f.n_mom0[v] = make_double4(vs.lin.x, vs.lin.y, vs.lin.z, vs.ang.x);
    f.n_mom1[v] = make_double2(vs.ang.y, vs.ang.z);
}

// closes the bookkeeping of the last step of a vx_step call
__global__ void k_lattice_finish(DevParams* p, int parity_of_last)
{
    if (p->div_flag[parity_of_last] | p->div_latched) p->div_latched = 1;
    else if (p->pending) { p->steps_done += 1; p->time += p->dt; }
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
            lat_eval_link(prev, axis, owner, meta_hi(n1.w), n0, n1, p0, p1, prev_dt_of_last, st, fN, mN, fP, mP);
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
            float f2 = lm.a1 / (m1 < m2 ? m1 : m2);
            if (f2 > best) best = f2;
        }
    }
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best > 0.0f) atomicMax(out, __float_as_uint(best));
}

} // namespace vxd
