// vx_kernels.cuh -- __global__ kernels of the general (arbitrary lattice) path.
//
//   k_link<AXIS,POISSON>   one thread per link, replaces CVX_Link::updateForces
//                          (src/VX_Link.cpp:149-217) for all links of one axis.
//   k_voxel                one thread per voxel, replaces CVX_Voxel::timeStep
//                          (src/VX_Voxel.cpp:162-232): 6-slot gather, no atomics.
//   k_pstrain              per-voxel Poisson strain pre-pass (src/VX_Voxel.cpp:300-343).
//   k_max_freq             warp-shuffle + grid max reduction of recommendedTimeStep
//                          (src/Voxelyze.cpp:286-311).
// Memory layout: vx_types.h.  All are HBM-bound streaming kernels: no shared memory reuse
// exists between threads, so the design rules are coalesced 128-bit accesses, enough
// resident warps to cover DRAM latency and grids of many waves over the 148 SMs.
#pragma once
#include <cooperative_groups.h>
#include "vx_physics.cuh"

namespace vxd {

struct Frame {
    int n_vox, n_link;
    // voxel arrays
    double4* pose0; double4* pose1; double4* mom0; double2* mom1;
    const int* ext_idx;
    float4* pstrain;
    double* slots;          // [6][n_vox][6]
    float* slot_strain;     // [6][n_vox] per-end axial strain (Poisson only)
    // link arrays
    const int2* lends; uint32_t* lmeta;
    double4* lstA; double4* lstB; double* lstC; float4* lstrain;
    // tables
    const DevVoxMat* vmat; const DevLinkMat* lmat; const float* curve_e; const float* curve_s;
    const DevExt* ext;
    DevParams* params;
    // single-material models: the two table rows travel in the kernel parameters (constant bank),
    // so material constants cost no load instructions (template parameter UNI)
    DevVoxMat vm0; DevLinkMat lm0;
    // collisions: per surface voxel CSR of signed contact references (vx_collide.cuh)
    const int* col_slot; const int* col_start; const int* col_ref; const float4* col_force;
};

__device__ __forceinline__ uint32_t meta_hi(double w) { return (uint32_t)(((unsigned long long)__double_as_longlong(w)) >> 32); }
__device__ __forceinline__ float meta_temp(double w) { return __uint_as_float((uint32_t)((unsigned long long)__double_as_longlong(w))); }
__device__ __forceinline__ double meta_pack(float temp, uint32_t hi)
{ return __longlong_as_double((long long)(((unsigned long long)hi << 32) | (unsigned long long)__float_as_uint(temp))); }

// 128-bit loads of the read-only pose records
__device__ __forceinline__ double4 ld4(const double4* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// Loads of state that another CTA of the same (persistent, multi-step) kernel may have written since this SM last read it
// (k_small_steps): COH = true bypasses L1 (ld.global.cg); COH = false is the plain / read-only path of the one-step kernels.
template <bool COH> __device__ __forceinline__ double4 ldv4(const double4* p)
{
    if (!COH) return ld4(p);
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldcg(q), b = __ldcg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
template <bool COH> __device__ __forceinline__ double4 ldw4(const double4* p)       // arrays the one-step kernels read with plain loads
{
    if (!COH) return *p;
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldcg(q), b = __ldcg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
template <bool COH, typename T> __device__ __forceinline__ T ldv(const T* p) { return COH ? __ldcg(p) : *p; }

// transverse area of a voxel cross-section seen along `axis` (src/VX_Voxel.cpp:361-374)
__device__ __forceinline__ float transverse_area(const DevVoxMat& m, int axis, float4 ps)
{
    float s = m.nom_f;
    if (m.nu == 0) return s * s;
    double px = ps.x, py = ps.y, pz = ps.z;
    if (axis == 0) return (float)(s * s * (1 + py) * (1 + pz));
    if (axis == 1) return (float)(s * s * (1 + px) * (1 + pz));
    return (float)(s * s * (1 + px) * (1 + py));
}
// (src/VX_Voxel.cpp:346-359)
__device__ __forceinline__ float transverse_strain_sum(const DevVoxMat& m, int axis, float4 ps)
{
    if (m.nu == 0) return 0.0f;
    if (axis == 0) return ps.y + ps.z;
    if (axis == 1) return ps.x + ps.z;
    return ps.x + ps.y;
}

// CVX_Link::updateForces of link l (src/VX_Link.cpp:149-217) on the general layout
template <int AXIS, bool POISSON, bool UNI, bool COH>
__device__ __forceinline__ void link_body(const int l, const Frame& f)
{
    if (ldv<COH>(&f.params->div_latched)) return;
    const float prev_dt = ldv<COH>(&f.params->prev_dt);

    int2 e = f.lends[l];
    double4 n0 = ldv4<COH>(f.pose0 + e.x), n1 = ldv4<COH>(f.pose1 + e.x);
    double4 p0 = ldv4<COH>(f.pose0 + e.y), p1 = ldv4<COH>(f.pose1 + e.y);
    double4 sa = ldw4<COH>(f.lstA + l), sb = ldw4<COH>(f.lstB + l);
    double sc = ldv<COH>(f.lstC + l);
    float4 sm = ldv<COH>(f.lstrain + l);
    uint32_t lm_bits = ldv<COH>(f.lmeta + l);

    const uint32_t hn = meta_hi(n1.w), hp = meta_hi(p1.w);
    const float tn = meta_temp(n1.w), tp = meta_temp(p1.w);
    const DevVoxMat& vmn = UNI ? f.vm0 : f.vmat[hn & VM_MAT_MASK];
    const DevVoxMat& vmp = UNI ? f.vm0 : f.vmat[hp & VM_MAT_MASK];
    const DevLinkMat& lm = UNI ? f.lm0 : f.lmat[lm_bits & LM_MAT_MASK];

    // CVX_Link::updateRestLength (src/VX_Link.cpp:137-140), (1+temp*cte) in float
    double rest = 0.5 * (vmn.size[AXIS] * (1 + tn * vmn.cte) + vmp.size[AXIS] * (1 + tp * vmp.cte));

    float t_area, t_sum = 0.0f;
    if (POISSON) {
        float4 psn = ldv<COH>(f.pstrain + e.x), psp = ldv<COH>(f.pstrain + e.y);
        t_area = 0.5f * (transverse_area(vmn, AXIS, psn) + transverse_area(vmp, AXIS, psp));
        t_sum = 0.5f * (transverse_strain_sum(vmn, AXIS, psn) + transverse_strain_sum(vmp, AXIS, psp));
    } else {
        t_area = 0.5f * (vmn.nom_f * vmn.nom_f + vmp.nom_f * vmp.nom_f);
    }

    LinkState st;
    st.pos2 = mk3(sa.x, sa.y, sa.z);
    st.a1v = mk3(sa.w, sb.x, sb.y);
    st.a2v = mk3(sb.z, sb.w, sc);
    st.strain = sm.x; st.max_strain = sm.y; st.strain_offset = sm.z; st.stress = sm.w;
    st.small_angle = (lm_bits & LM_SMALL_ANGLE) != 0;
    st.vel_valid = (lm_bits & LM_VEL_VALID) != 0;

    float damp_n = vmn.two_sqrtm_zeta / prev_dt;          // CVX_Voxel::dampingMultiplier, VX_Voxel.h:130
    float damp_p = vmp.two_sqrtm_zeta / prev_dt;

    q4 on, op;
    on.w = n0.w; on.x = n1.x; on.y = n1.y; on.z = n1.z;
    op.w = p0.w; op.x = p1.x; op.y = p1.y; op.z = p1.z;
    d3 fN, mN, fP, mP;
    link_forces(AXIS, mk3(n0.x, n0.y, n0.z), on, mk3(p0.x, p0.y, p0.z), op, rest, t_area, t_sum,
                      damp_n, damp_p, lm, f.curve_e, f.curve_s, st, fN, mN, fP, mP);

    f.lstA[l] = make_double4(st.pos2.x, st.pos2.y, st.pos2.z, st.a1v.x);
    f.lstB[l] = make_double4(st.a1v.y, st.a1v.z, st.a2v.x, st.a2v.y);
    f.lstC[l] = st.a2v.z;
    f.lstrain[l] = make_float4(st.strain, st.max_strain, st.strain_offset, st.stress);
    f.lmeta[l] = (lm_bits & LM_MAT_MASK) | (st.small_angle ? LM_SMALL_ANGLE : 0u) | (st.vel_valid ? LM_VEL_VALID : 0u);
    if (st.strain > 100) f.params->div_now = 1;            // src/Voxelyze.cpp:265

    const size_t nv = (size_t)f.n_vox;
    double2* sn = reinterpret_cast<double2*>(f.slots + ((size_t)(2 * AXIS) * nv + e.x) * 6);
    double2* sp = reinterpret_cast<double2*>(f.slots + ((size_t)(2 * AXIS + 1) * nv + e.y) * 6);
    sn[0] = make_double2(fN.x, fN.y); sn[1] = make_double2(fN.z, mN.x); sn[2] = make_double2(mN.y, mN.z);
    sp[0] = make_double2(fP.x, fP.y); sp[1] = make_double2(fP.z, mP.x); sp[2] = make_double2(mP.y, mP.z);

    if (POISSON) {                                          // CVX_Link::axialStrain(bool), src/VX_Link.cpp:121-124
        float ratio = vmp.E / vmn.E;
        f.slot_strain[(size_t)(2 * AXIS) * nv + e.x] = 2.0f * st.strain / (1.0f + ratio);
        f.slot_strain[(size_t)(2 * AXIS + 1) * nv + e.y] = 2.0f * st.strain * ratio / (1.0f + ratio);
    }
}

template <int AXIS, bool POISSON, bool UNI>
__global__ void __launch_bounds__(128) k_link(int first, int count, Frame f)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    link_body<AXIS, POISSON, UNI, false>(first + t, f);
}

// CVX_Voxel::timeStep of voxel v (src/VX_Voxel.cpp:162-232) on the general layout
template <bool UNI, bool COH>
__device__ __forceinline__ void voxel_body(const int v, const Frame& f, int floor_on, int collisions)
{
    DevParams* p = f.params;
    if (ldv<COH>(&p->div_now) | ldv<COH>(&p->div_latched)) { if (v == 0) p->div_latched = 1; return; }
    const float dt = p->dt;
    if (v == 0) { p->prev_dt = dt; p->time = ldv<COH>(&p->time) + dt; p->steps_done = ldv<COH>(&p->steps_done) + 1; }

    double4 q1 = ldw4<COH>(f.pose1 + v);
    VoxelState s;
    s.bits = meta_hi(q1.w);
    if (s.bits & VM_GHOST) return;
    s.temp = meta_temp(q1.w);
    double4 q0 = ldw4<COH>(f.pose0 + v), m0 = ldw4<COH>(f.mom0 + v);
    double2 m1 = ldv<COH>(f.mom1 + v);
    s.pos = mk3(q0.x, q0.y, q0.z);
    s.orient.w = q0.w; s.orient.x = q1.x; s.orient.y = q1.y; s.orient.z = q1.z;
    s.lin = mk3(m0.x, m0.y, m0.z);
    s.ang = mk3(m0.w, m1.x, m1.y);

    // gather link forces in slot order X+,X-,Y+,Y-,Z+,Z- (src/VX_Voxel.cpp:238-240, 262-264)
    d3 F = mk3(0.0, 0.0, 0.0), M = mk3(0.0, 0.0, 0.0);
    const uint32_t mask = (s.bits >> VM_LINK_SHIFT) & 0x3Fu;
    const size_t nv = (size_t)f.n_vox;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        if (mask & (1u << k)) {
            const double2* sl = reinterpret_cast<const double2*>(f.slots + ((size_t)k * nv + v) * 6);
            double2 a = ldv<COH>(sl), b = ldv<COH>(sl + 1), c = ldv<COH>(sl + 2);
            F = F + mk3(a.x, a.y, b.x);
            M = M + mk3(b.y, c.x, c.y);
        }
    }
    const DevVoxMat& vm = UNI ? f.vm0 : f.vmat[s.bits & VM_MAT_MASK];
    const DevExt* ext = (s.bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;

    const int* refs = nullptr; int n_refs = 0;
    if (collisions && mask != 0x3Fu) {                     // only surface voxels are ever watched
        const int slot = f.col_slot[v];
        refs = f.col_ref + f.col_start[slot];
        n_refs = f.col_start[slot + 1] - f.col_start[slot];
    }
    voxel_integrate(s, F, M, refs, n_refs, f.col_force, vm, ext, dt, floor_on != 0);

    f.pose0[v] = make_double4(s.pos.x, s.pos.y, s.pos.z, s.orient.w);
    f.pose1[v] = make_double4(s.orient.x, s.orient.y, s.orient.z, meta_pack(s.temp, s.bits));
    f.mom0[v] = make_double4(s.lin.x, s.lin.y, s.lin.z, s.ang.x);
    f.mom1[v] = make_double2(s.ang.y, s.ang.z);
}

template <bool UNI>
__global__ void __launch_bounds__(128) k_voxel(Frame f, int floor_on, int collisions)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= f.n_vox) return;
    voxel_body<UNI, false>(v, f, floor_on, collisions);
}

// CVX_Voxel::strain(true) (src/VX_Voxel.cpp:300-343) from the sums r[] and counts nb[] of the per-end axial strains of the
// voxel's links along each axis
__device__ __forceinline__ float4 voxel_pstrain(const DevVoxMat& vm, const DevExt* ext, float r[3], const int nb[3])
{
    const uint32_t dof = ext ? ext->dof : 0u;
    bool tension[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (nb[i] == 2) r[i] *= 0.5f;
        tension[i] = (nb[i] == 2) || (ext && nb[i] == 1 && ((dof & (1u << i)) || ext->force[i] != 0));
    }
    if (!(tension[0] && tension[1] && tension[2])) {
        float add = 0;
        for (int i = 0; i < 3; i++) if (tension[i]) add += r[i];
        // powf of the reference, evaluated in double and rounded once (glibc powf is < 1 ulp)
        float value = (float)pow((double)(1.0f + add), (double)(-vm.nu)) - 1.0f;
        for (int i = 0; i < 3; i++) if (!tension[i]) r[i] = value;
    }
    return make_float4(r[0], r[1], r[2], 0.0f);
}

// CVX_Voxel::strain(true) for voxels whose cache is stale (src/VX_Voxel.cpp:300-343).
// The reference fills this cache lazily inside the link loop; every link strain it reads is
// still the previous step's at that point, so a pre-pass over stale voxels is equivalent.
template <bool COH>
__device__ __forceinline__ void pstrain_body(const int v, const Frame& f)
{
    if (ldv<COH>(&f.params->div_latched)) return;
    double w = ldv<COH>(&f.pose1[v].w);
    uint32_t bits = meta_hi(w);
    if (!(bits & VM_PSTRAIN_STALE)) return;
    const DevVoxMat& vm = f.vmat[bits & VM_MAT_MASK];
    const uint32_t mask = (bits >> VM_LINK_SHIFT) & 0x3Fu;
    const size_t nv = (size_t)f.n_vox;
    float r[3] = {0.0f, 0.0f, 0.0f};
    int nb[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (mask & (1u << k)) { r[k >> 1] += ldv<COH>(f.slot_strain + (size_t)k * nv + v); nb[k >> 1]++; }
    const DevExt* ext = (bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
    f.pstrain[v] = voxel_pstrain(vm, ext, r, nb);
    f.pose1[v].w = meta_pack(meta_temp(w), bits & ~VM_PSTRAIN_STALE);
}
__global__ void __launch_bounds__(128) k_pstrain(Frame f)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= f.n_vox) return;
    pstrain_body<false>(v, f);
}

// ---- small models: ALL steps of a vx_step call in ONE launch -------------------------------------------------------
// A model of a few hundred voxels (320 voxels, 784 links) cannot fill the GPU; what it pays per step is launch
// latency and the dependent FP64 chain of one link and one voxel update.  k_small_steps runs the whole call as one
// thread-block cluster (<= 16 CTAs of 64 or 128 threads on as many SMs: one or two warps per SM sub-partition): per step a
// link phase (thread per link, axis by range), a cluster barrier, a voxel phase, a cluster barrier -- the barrier
// (barrier.cluster, release/acquire) replaces the kernel boundary of the general path, the state stays in L2 (state loads
// bypass L1, ld.global.cg, because another SM wrote them one phase earlier).  Same device functions, same per-voxel
// summation order: bit-identical to k_link / k_voxel.
template <bool POISSON, bool UNI>
__global__ void __launch_bounds__(256) k_small_steps(Frame f, int af1, int af2, int n_steps, int floor_on)
{
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int step = 0; step < n_steps; step++) {
        if (POISSON) {
            for (int v = tid; v < f.n_vox; v += nthr) pstrain_body<true>(v, f);
            cluster.sync();
        }
#ifndef VX_SMALL_PHASES
#define VX_SMALL_PHASES 7      // ablation only (profiles/r2_small_model.md): 1 link phase, 2 voxel phase, 4 cluster barriers
#endif
        if (VX_SMALL_PHASES & 1)
        for (int l = tid; l < f.n_link; l += nthr) {
            if (l < af1) link_body<0, POISSON, UNI, true>(l, f);
            else if (l < af2) link_body<1, POISSON, UNI, true>(l, f);
            else link_body<2, POISSON, UNI, true>(l, f);
        }
        if (VX_SMALL_PHASES & 4) cluster.sync();
        if (VX_SMALL_PHASES & 2)
        for (int v = tid; v < f.n_vox; v += nthr) voxel_body<UNI, true>(v, f, floor_on, 0);
        if (VX_SMALL_PHASES & 4) cluster.sync();
    }
}

// max over links of axialStiffness/min(m1,m2) (src/Voxelyze.cpp:291-299, src/VX_Link.cpp:259-267)
__global__ void __launch_bounds__(256) k_max_freq(Frame f, int axis_first1, int axis_first2, unsigned int* out)
{
    float best = 0.0f;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < f.n_link; l += gridDim.x * blockDim.x) {
        int2 e = f.lends[l];
        double wn = f.pose1[e.x].w, wp = f.pose1[e.y].w;
        const DevVoxMat& vmn = f.vmat[meta_hi(wn) & VM_MAT_MASK];
        const DevVoxMat& vmp = f.vmat[meta_hi(wp) & VM_MAT_MASK];
        const DevLinkMat& lm = f.lmat[f.lmeta[l] & LM_MAT_MASK];
        float stiff;
        if (lm.nu == 0.0f) stiff = lm.a1;
        else {
            int axis = l >= axis_first2 ? 2 : (l >= axis_first1 ? 1 : 0);
            double rest = 0.5 * (vmn.size[axis] * (1 + meta_temp(wn) * vmn.cte) + vmp.size[axis] * (1 + meta_temp(wp) * vmp.cte));
            float area = 0.5f * (transverse_area(vmn, axis, f.pstrain[e.x]) + transverse_area(vmp, axis, f.pstrain[e.y]));
            stiff = (float)(lm.e_hat * area / ((f.lstrain[l].x + 1) * rest));
        }
        float m1 = vmn.mass, m2 = vmp.mass;
        float f2 = stiff / (m1 < m2 ? m1 : m2);
        if (f2 > best) best = f2;
    }
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    __shared__ float warp_best[8];
    if ((threadIdx.x & 31) == 0) warp_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        float b = threadIdx.x < (blockDim.x >> 5) ? warp_best[threadIdx.x] : 0.0f;
        for (int o = 4; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
        if (threadIdx.x == 0 && b > 0.0f) atomicMax(out, __float_as_uint(b));   // positive floats order like uints
    }
}

// fallback of recommendedTimeStep when there are no links (src/Voxelyze.cpp:302-307)
__global__ void __launch_bounds__(256) k_max_freq_voxels(Frame f, unsigned int* out)
{
    float best = 0.0f;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.n_vox; v += gridDim.x * blockDim.x) {
        const DevVoxMat& vm = f.vmat[meta_hi(f.pose1[v].w) & VM_MAT_MASK];
        float f2 = vm.E * vm.nom / vm.mass;
        if (f2 > best) best = f2;
    }
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best > 0.0f) atomicMax(out, __float_as_uint(best));
}

// dt = 1/(2*pi*sqrt(maxFreq2)) written to the device-resident step parameters
// roll (fused path, calls that re-evaluate dt every step): the step before this one ran in the same call with another dt --
// it is counted here, with ITS dt, before that is replaced (the fused kernels otherwise count a step when the next one
// starts), and its dt becomes CVX_Voxel::previousDt; parity_prev: generation parity whose divergence flag that step wrote
__global__ void k_dt_from_freq(const unsigned int* freq2, DevParams* p, int roll, int parity_prev = 0)
{
    float m = __uint_as_float(*freq2);
    if (roll) {
        if (p->div_flag[parity_prev] | p->div_latched) p->div_latched = 1;
        else if (p->pending) { p->steps_done += 1; p->time += p->dt; p->pending = 0; }
        p->prev_dt = p->dt;
    }
    p->dt = (m <= 0.0f) ? 0.0f : 1.0f / (6.283185f * sqrtf(m));
}

// ------------------------------------------------------------------ state access helpers
// caller-order <-> internal-order gather/scatter, so one cudaMemcpy moves any field.
enum { G_POS, G_ORIENT, G_LINMOM, G_ANGMOM, G_TEMP, G_VOXFLAGS, G_PSTRAIN,
       G_FORCE_NEG, G_FORCE_POS, G_MOMENT_NEG, G_MOMENT_POS, G_POS2, G_ANGLE1V, G_ANGLE2V,
       G_STRAIN, G_MAXSTRAIN, G_STRAINOFFSET, G_STRESS, G_LINKFLAGS };

__global__ void k_gather(Frame f, int what, const int* e2i, int first, int count, void* out, int axis_first1, int axis_first2)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int i = e2i[first + k];
    double* d = (double*)out; float* fl = (float*)out; uint32_t* u = (uint32_t*)out;
    const size_t nv = (size_t)f.n_vox;
    switch (what) {
    case G_POS: { double4 a = f.pose0[i]; d[3 * k] = a.x; d[3 * k + 1] = a.y; d[3 * k + 2] = a.z; break; }
    case G_ORIENT: { double4 a = f.pose0[i], b = f.pose1[i]; d[4 * k] = a.w; d[4 * k + 1] = b.x; d[4 * k + 2] = b.y; d[4 * k + 3] = b.z; break; }
    case G_LINMOM: { double4 a = f.mom0[i]; d[3 * k] = a.x; d[3 * k + 1] = a.y; d[3 * k + 2] = a.z; break; }
    case G_ANGMOM: { double4 a = f.mom0[i]; double2 b = f.mom1[i]; d[3 * k] = a.w; d[3 * k + 1] = b.x; d[3 * k + 2] = b.y; break; }
    case G_TEMP: fl[k] = meta_temp(f.pose1[i].w); break;
    case G_VOXFLAGS: {
        uint32_t b = meta_hi(f.pose1[i].w);
        u[k] = ((b & VM_STATIC_FRIC) ? 1u : 0u) | ((((b >> VM_LINK_SHIFT) & 0x3Fu) != 0x3Fu) ? 2u : 0u) | ((b & VM_GHOST) ? 4u : 0u) |
               ((b & VM_FLOOR_OFF) ? 8u : 0u) | ((b & VM_FLOOR_ON) ? 16u : 0u);
        break; }
    case G_PSTRAIN: { float4 a = f.pstrain ? f.pstrain[i] : make_float4(0, 0, 0, 0); fl[3 * k] = a.x; fl[3 * k + 1] = a.y; fl[3 * k + 2] = a.z; break; }
    case G_FORCE_NEG: case G_MOMENT_NEG: case G_FORCE_POS: case G_MOMENT_POS: {
        int axis = i >= axis_first2 ? 2 : (i >= axis_first1 ? 1 : 0);
        int2 e = f.lends[i];
        bool pos_end = (what == G_FORCE_POS || what == G_MOMENT_POS);
        const double* s = f.slots + ((size_t)(2 * axis + (pos_end ? 1 : 0)) * nv + (pos_end ? e.y : e.x)) * 6;
        int off = (what == G_MOMENT_NEG || what == G_MOMENT_POS) ? 3 : 0;
        d[3 * k] = s[off]; d[3 * k + 1] = s[off + 1]; d[3 * k + 2] = s[off + 2];
        break; }
    case G_POS2: { double4 a = f.lstA[i]; d[3 * k] = a.x; d[3 * k + 1] = a.y; d[3 * k + 2] = a.z; break; }
    case G_ANGLE1V: { double4 a = f.lstA[i], b = f.lstB[i]; d[3 * k] = a.w; d[3 * k + 1] = b.x; d[3 * k + 2] = b.y; break; }
    case G_ANGLE2V: { double4 b = f.lstB[i]; d[3 * k] = b.z; d[3 * k + 1] = b.w; d[3 * k + 2] = f.lstC[i]; break; }
    case G_STRAIN: fl[k] = f.lstrain[i].x; break;
    case G_MAXSTRAIN: fl[k] = f.lstrain[i].y; break;
    case G_STRAINOFFSET: fl[k] = f.lstrain[i].z; break;
    case G_STRESS: fl[k] = f.lstrain[i].w; break;
    case G_LINKFLAGS: {
        uint32_t b = f.lmeta[i]; const DevLinkMat& lm = f.lmat[b & LM_MAT_MASK]; float mx = f.lstrain[i].y;
        u[k] = ((b & LM_SMALL_ANGLE) ? 1u : 0u) | ((b & LM_VEL_VALID) ? 2u : 0u) | (mat_yielded(lm, mx) ? 4u : 0u) | (mat_failed(lm, mx) ? 8u : 0u);
        break; }
    }
}

struct VoxelStateRec { double pos[3], orient[4], linmom[3], angmom[3]; float temp; uint32_t flags; };       // = vx_voxel_state
__global__ void k_gather_voxel_state(Frame f, const int* e2i, int first, int count, VoxelStateRec* out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = e2i[first + k];
    const double4 a = f.pose0[i], b = f.pose1[i], c = f.mom0[i]; const double2 d = f.mom1[i];
    VoxelStateRec r;
    r.pos[0] = a.x; r.pos[1] = a.y; r.pos[2] = a.z; r.orient[0] = a.w; r.orient[1] = b.x; r.orient[2] = b.y; r.orient[3] = b.z;
    r.linmom[0] = c.x; r.linmom[1] = c.y; r.linmom[2] = c.z; r.angmom[0] = c.w; r.angmom[1] = d.x; r.angmom[2] = d.y;
    r.temp = meta_temp(b.w);
    const uint32_t m = meta_hi(b.w);
    r.flags = ((m & VM_STATIC_FRIC) ? 1u : 0u) | ((((m >> VM_LINK_SHIFT) & 0x3Fu) != 0x3Fu) ? 2u : 0u) | ((m & VM_GHOST) ? 4u : 0u) |
              ((m & VM_FLOOR_OFF) ? 8u : 0u) | ((m & VM_FLOOR_ON) ? 16u : 0u);
    out[k] = r;
}

__global__ void k_scatter(Frame f, int what, const int* e2i, int first, int count, const void* in)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int i = e2i[first + k];
    const double* d = (const double*)in; const float* fl = (const float*)in; const uint32_t* u = (const uint32_t*)in;
    switch (what) {
    case G_POS: { double4 a = f.pose0[i]; a.x = d[3 * k]; a.y = d[3 * k + 1]; a.z = d[3 * k + 2]; f.pose0[i] = a; break; }
    case G_ORIENT: { double4 a = f.pose0[i], b = f.pose1[i]; a.w = d[4 * k]; b.x = d[4 * k + 1]; b.y = d[4 * k + 2]; b.z = d[4 * k + 3]; f.pose0[i] = a; f.pose1[i] = b; break; }
    case G_LINMOM: { double4 a = f.mom0[i]; a.x = d[3 * k]; a.y = d[3 * k + 1]; a.z = d[3 * k + 2]; f.mom0[i] = a; break; }
    case G_ANGMOM: { double4 a = f.mom0[i]; a.w = d[3 * k]; f.mom0[i] = a; f.mom1[i] = make_double2(d[3 * k + 1], d[3 * k + 2]); break; }
    case G_TEMP: { double w = f.pose1[i].w; f.pose1[i].w = meta_pack(fl[k], meta_hi(w)); break; }
    case G_PSTRAIN: if (f.pstrain) f.pstrain[i] = make_float4(fl[3 * k], fl[3 * k + 1], fl[3 * k + 2], 0.f); break;     // a ghost's Poisson strain comes from its owner
    case G_VOXFLAGS: { double w = f.pose1[i].w; uint32_t b = meta_hi(w); b = (b & ~(VM_STATIC_FRIC | VM_FLOOR_OFF | VM_FLOOR_ON)) | ((u[k] & 1u) ? VM_STATIC_FRIC : 0u) | ((u[k] & 8u) ? VM_FLOOR_OFF : 0u) | ((u[k] & 16u) ? VM_FLOOR_ON : 0u); f.pose1[i].w = meta_pack(meta_temp(w), b); break; }
    }
}

// persistent state of links [first, first+count) in caller order <- packed vx_link_state records (96 B each)
struct LinkStateRec { double pos2[3], a1v[3], a2v[3]; float strain, max_strain, strain_offset, stress; uint32_t flags, pad; };
__global__ void k_scatter_link_state(Frame f, const int* e2i, int first, int count, const LinkStateRec* src, int axis_first1, int axis_first2)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int l = e2i[first + k];
    const LinkStateRec r = src[k];
    {   // the per-end axial strains the Poisson pre-pass reads (CVX_Link::axialStrain(bool), src/VX_Link.cpp:121-124)
        const int2 e = f.lends[l];
        const int axis = l >= axis_first2 ? 2 : (l >= axis_first1 ? 1 : 0);
        const float ratio = f.vmat[meta_hi(f.pose1[e.y].w) & VM_MAT_MASK].E / f.vmat[meta_hi(f.pose1[e.x].w) & VM_MAT_MASK].E;
        const size_t nv = (size_t)f.n_vox;
        f.slot_strain[(size_t)(2 * axis) * nv + e.x] = 2.0f * r.strain / (1.0f + ratio);
        f.slot_strain[(size_t)(2 * axis + 1) * nv + e.y] = 2.0f * r.strain * ratio / (1.0f + ratio);
    }
    f.lstA[l] = make_double4(r.pos2[0], r.pos2[1], r.pos2[2], r.a1v[0]);
    f.lstB[l] = make_double4(r.a1v[1], r.a1v[2], r.a2v[0], r.a2v[1]);
    f.lstC[l] = r.a2v[2];
    f.lstrain[l] = make_float4(r.strain, r.max_strain, r.strain_offset, r.stress);
    f.lmeta[l] = (f.lmeta[l] & LM_MAT_MASK) | ((r.flags & 1u) ? LM_SMALL_ANGLE : 0u) | ((r.flags & 2u) ? LM_VEL_VALID : 0u);
}

// Poisson's ratio switched on mid-run on the general layout: the per-end axial strains the Poisson pre-pass reads are only kept
// up to date by the POISSON link kernels, so they are rebuilt here from the current link strains (src/VX_Link.cpp:121-124)
__global__ void k_refresh_slot_strain(Frame f, int axis_first1, int axis_first2)
{
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= f.n_link) return;
    const int2 e = f.lends[l];
    const int axis = l >= axis_first2 ? 2 : (l >= axis_first1 ? 1 : 0);
    const float strain = f.lstrain[l].x;
    const float ratio = f.vmat[meta_hi(f.pose1[e.y].w) & VM_MAT_MASK].E / f.vmat[meta_hi(f.pose1[e.x].w) & VM_MAT_MASK].E;
    const size_t nv = (size_t)f.n_vox;
    f.slot_strain[(size_t)(2 * axis) * nv + e.x] = 2.0f * strain / (1.0f + ratio);
    f.slot_strain[(size_t)(2 * axis + 1) * nv + e.y] = 2.0f * strain * ratio / (1.0f + ratio);
}

__global__ void k_fill_temp(Frame f, float t, const float* member_t, const int* member_of)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= f.n_vox) return;
    double w = f.pose1[v].w;
    f.pose1[v].w = meta_pack(member_t ? member_t[member_of[v]] : t, meta_hi(w));
}

__global__ void k_clear_ext_bits(Frame f)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= f.n_vox) return;
    double w = f.pose1[v].w;
    uint32_t b = meta_hi(w);
    if (b & VM_HAS_EXT) f.pose1[v].w = meta_pack(meta_temp(w), b & ~VM_HAS_EXT);
}

__global__ void k_set_ext_bits(Frame f, int n, const int* vox, int* ext_idx)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int v = vox[k];
    ext_idx[v] = k;
    double w = f.pose1[v].w;
    f.pose1[v].w = meta_pack(meta_temp(w), meta_hi(w) | VM_HAS_EXT);
}

} // namespace vxd
