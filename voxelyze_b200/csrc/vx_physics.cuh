// vx_physics.cuh -- device functions of the dynamics step (FP64 state, FP32 material math).
//
// The functions here are the per-link and per-voxel physics shared by every kernel variant
// (general two-kernel path, fused lattice path).  They follow the reference's evaluation
// order exactly and the translation unit is compiled with
// -fmad=false, so results differ from the x86-64 reference only where libm and CUDA's
// sin/cos/acos/pow differ (<= 2 ulp, large-angle links and large rotation increments only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "vx_types.h"

namespace vxd {

struct d3 { double x, y, z; };
struct q4 { double w, x, y, z; };

__device__ __forceinline__ d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator-(d3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ d3 operator*(double f, d3 a) { return mk3(f * a.x, f * a.y, f * a.z); }
__device__ __forceinline__ double norm2(d3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

// a / b, bit for bit.  CUDA's double division leaves its inline path whenever the numerator is zero (a ~70-instruction
// subroutine), and zero numerators are the rule wherever a lattice is still at rest (transverse offsets, strains and
// rotation increments are exactly 0 there).  0 * b has the sign of 0 / b for every finite non-zero b.
__device__ __forceinline__ double ddiv(double a, double b)
{
#ifndef VX_NO_DDIV            // ablation switch (tools/build_variant.sh)
    if (a == 0.0 && b != 0.0 && fabs(b) <= 1.7976931348623157e308) return a * b;
#endif
    return a / b;
}

// quaternion product (include/Quat3D.h:83)
__device__ __forceinline__ q4 qmul(const q4& a, const q4& b)
{
    q4 r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return r;
}
__device__ __forceinline__ q4 qconj(const q4& a) { q4 r; r.w = a.w; r.x = -a.x; r.y = -a.y; r.z = -a.z; return r; }
__device__ __forceinline__ q4 qident() { q4 r; r.w = 1.0; r.x = r.y = r.z = 0.0; return r; }

// v rotated by q (include/Quat3D.h:170-177)
__device__ __forceinline__ d3 qrot(const q4& q, d3 f)
{
    double tw = f.x * q.x + f.y * q.y + f.z * q.z;
    double tx = f.x * q.w - f.y * q.z + f.z * q.y;
    double ty = f.x * q.z + f.y * q.w - f.z * q.x;
    double tz = -f.x * q.y + f.y * q.x + f.z * q.w;
    return mk3(q.w * tx + q.x * tw + q.y * tz - q.z * ty,
               q.w * ty - q.x * tz + q.y * tw + q.z * tx,
               q.w * tz + q.x * ty - q.y * tx + q.z * tw);
}
// v rotated by the inverse of q (include/Quat3D.h:188-195)
__device__ __forceinline__ d3 qrot_inv(const q4& q, d3 f)
{
    double tw = q.x * f.x + q.y * f.y + q.z * f.z;
    double tx = q.w * f.x - q.y * f.z + q.z * f.y;
    double ty = q.w * f.y + q.x * f.z - q.z * f.x;
    double tz = q.w * f.z - q.x * f.y + q.y * f.x;
    return mk3(tw * q.x + tx * q.w + ty * q.z - tz * q.y,
               tw * q.y - tx * q.z + ty * q.w + tz * q.x,
               tw * q.z + tx * q.y - ty * q.x + tz * q.w);
}
// Rare branches (rotations beyond a few degrees, non-linear material curves) are kept out of line: the hot path of a
// link evaluation then is a third of the code, which matters because a fused kernel holds several copies of it and
// the instruction cache is 128 KB.  Same arithmetic either way.
__device__ __noinline__ void q_to_rotvec_acos(double w, double sl, double& scale_u, double& inv)
{
    scale_u = acos(w);
    inv = 1.0 / sqrt(sl);
}
// quaternion -> rotation vector (include/Quat3D.h:117-122)
__device__ __forceinline__ d3 q_to_rotvec(const q4& q)
{
    if (q.w >= 1.0 || q.w <= -1.0) return mk3(0.0, 0.0, 0.0);
    double sl = 1.0 - q.w * q.w;
    d3 v = mk3(2.0 * q.x, 2.0 * q.y, 2.0 * q.z);
    if (sl < 2.4e-3) return sqrt((2 - 2 * q.w) / sl) * v;
    double a, inv;
    q_to_rotvec_acos(q.w, sl, a, inv);
    d3 u = a * v;
    return inv * u;
}
__device__ __noinline__ void q_from_rotvec_trig(double m2, double& w, double& s)
{
    double m = sqrt(m2); w = cos(m); s = sin(m) / m;
}
// rotation vector -> quaternion (include/Quat3D.h:124-139)
__device__ __forceinline__ q4 q_from_rotvec(d3 v)
{
    d3 h = 0.5 * v;
    double m2 = norm2(h), w, s;
    if (m2 * m2 < 5.328e-15) { w = 1.0 - 0.5 * m2; s = 1.0 - ddiv(m2, 6.0); }
    else q_from_rotvec_trig(m2, w, s);
    q4 r; r.w = w; r.x = h.x * s; r.y = h.y * s; r.z = h.z * s;
    return r;
}
__device__ __noinline__ void q_align_to_x_large(double fx, double fy, double fz, double& qw, double& qy, double& qz)
{
    double l = sqrt(fx * fx + fy * fy + fz * fz);
    double nx = fx, ny = fy, nz = fz;
    if (l > 0) { double li = 1.0 / l; nx *= li; ny *= li; nz *= li; }
    // The reference computes theta = acos(nx), then cos(theta/2) and sin(theta/2) (include/Quat3D.h:154-163).  Those two are
    // sqrt((1 + nx)/2) and sqrt((1 - nx)/2): 1 -+ nx is exact for nx in [-1, -1/2] / [1/2, 1] and the square root is correctly
    // rounded, so the half-angle form is at least as close to the true value as acos followed by sin / cos (each good to
    // 1-2 ulp, here and in glibc) -- at a tenth of the instructions.  Only the reference's 180-degree special case needs the
    // angle itself, and only within 1e-7 rad of it.
    if (nx < -0.99999999) {
        double theta = acos(nx);
        if (theta > 3.14159265358979 - 1e-7) { qw = 0.0; qy = 1.0; qz = 0.0; return; }
    }
    double ami = 1.0 / sqrt(nz * nz + ny * ny);
    double s = sqrt(0.5 * (1.0 - nx));
    qw = sqrt(0.5 * (1.0 + nx)); qy = nz * ami * s; qz = -ny * ami * s;
}
// rotation that takes `from` onto +X (include/Quat3D.h:141-166), starting from identity
__device__ __forceinline__ q4 q_align_to_x(d3 from)
{
    q4 q = qident();
    if (from.x == 0.0 && from.y == 0.0 && from.z == 0.0) return q;
    double yox = from.y / from.x, zox = from.z / from.x;
    const double sa = 1.732e-2;
    if (yox < sa && yox > -sa && zox < sa && zox > -sa) {
        q.x = 0.0; q.y = 0.5 * zox; q.z = -0.5 * yox;
        q.w = 1 + 0.5 * (-q.y * q.y - q.z * q.z);
        return q;
    }
    q_align_to_x_large(from.x, from.y, from.z, q.w, q.y, q.z);
    q.x = 0.0;
    return q;
}

// link-axis swizzles: express vectors as if the link pointed along +X (include/VX_Link.h:112-117).
// `axis` is a compile-time constant in the per-axis kernels and a loop variable in the fused one.
__device__ __forceinline__ d3 to_axis_x(int axis, d3 v)
{
    if (axis == 1) return mk3(v.y, -v.x, v.z);
    if (axis == 2) return mk3(v.z, v.y, -v.x);
    return v;
}
__device__ __forceinline__ q4 to_axis_x(int axis, q4 q)
{
    q4 r = q;
    if (axis == 1) { r.x = q.y; r.y = -q.x; }
    if (axis == 2) { r.x = q.z; r.z = -q.x; }
    return r;
}
__device__ __forceinline__ d3 to_axis_original(int axis, d3 v)
{
    if (axis == 1) return mk3(-v.y, v.x, v.z);
    if (axis == 2) return mk3(-v.z, v.y, v.x);
    return v;
}

// ---------------------------------------------------------------------------------------------
// material model on the device: CVX_Material::stress (src/VX_Material.cpp:165-195), float math
__device__ __forceinline__ bool mat_failed(const DevLinkMat& m, float strain) { return m.eps_fail != -1.0f && strain > m.eps_fail; }
__device__ __forceinline__ bool mat_yielded(const DevLinkMat& m, float strain) { return m.eps_yield != -1.0f && strain > m.eps_yield; }

// beyond the first segment of a data curve (src/VX_Material.cpp:176-194)
__device__ __noinline__ float mat_stress_curve(const float* __restrict__ e, const float* __restrict__ s, int n, float nu, float strain, float tss)
{
    for (int i = 2; i < n; i++) {
        float ei = __ldg(e + i);
        if (strain <= ei || i == n - 1) {
            float e0 = __ldg(e + i - 1), s0 = __ldg(s + i - 1), s1 = __ldg(s + i);
            float perc = (strain - e0) / (ei - e0);
            float basic = s0 + perc * (s1 - s0);
            if (nu == 0.0f) return basic;
            float modulus = (s1 - s0) / (ei - e0);
            float mod_hat = modulus / ((1 - 2 * nu) * (1 + nu));
            float eff = basic / modulus;
            float eff_tss = tss * (eff / strain);
            return mod_hat * ((1 - nu) * eff + nu * eff_tss);
        }
    }
    return 0.0f;
}
__device__ __forceinline__ float mat_stress(const DevLinkMat& m, const float* __restrict__ ce, const float* __restrict__ cs,
                                            float strain, float tss, bool force_linear)
{
    if (mat_failed(m, strain)) return 0.0f;
    const float* e = ce + m.curve_off; const float* s = cs + m.curve_off;
    if (m.linear || force_linear || strain <= __ldg(e + 1)) {       // same predicate as src/VX_Material.cpp:170, ordered so linear models never touch the curve
        if (m.nu == 0.0f) return m.E * strain;
        return m.e_hat * ((1 - m.nu) * strain + m.nu * tss);
    }
    return mat_stress_curve(e, s, m.curve_n, m.nu, strain, tss);
}

// persistent state of one link (include/VX_Link.h:74-107)
struct LinkState {
    d3 pos2, a1v, a2v;
    float strain, max_strain, strain_offset, stress;
    bool small_angle, vel_valid;
};

// CVX_Link::updateStrain (src/VX_Link.cpp:220-249)
__device__ __forceinline__ float link_update_strain(LinkState& st, const DevLinkMat& m, const float* __restrict__ ce,
                                                    const float* __restrict__ cs, float axial, float tss)
{
    st.strain = axial;
    if (m.linear) {
        if (axial > st.max_strain) st.max_strain = axial;
        return mat_stress(m, ce, cs, axial, tss, false);
    }
    float ret;
    if (axial > st.max_strain) {
        st.max_strain = axial;
        ret = mat_stress(m, ce, cs, axial, tss, false);
        if (m.nu != 0.0f) st.strain_offset = st.max_strain - mat_stress(m, ce, cs, axial, 0.0f, false) / (m.e_hat * (1 - m.nu));
        else st.strain_offset = st.max_strain - ret / m.E;
    } else {
        float rel = axial - st.strain_offset;
        if (m.nu != 0.0f) ret = mat_stress(m, ce, cs, rel, tss, true);
        else ret = m.E * rel;
    }
    return ret;
}

// CVX_Link::updateForces + orientLink (src/VX_Link.cpp:77-119, 149-217).
// In:  poses of both end voxels, rest length, transverse area / strain sum, damping multipliers
//      (2*sqrtMass*zeta/previousDt of each end, float), link material.
// I/O: st (old pos2/angle1v/angle2v in, new out; strain memory; flags).
// Out: force/moment on the negative and positive end voxel in that voxel's local frame.
__device__ __forceinline__ void link_forces(const int axis, d3 pN, q4 oN, d3 pP, q4 oP, double rest_len, float t_area, float t_sum,
                                            float damp_n, float damp_p, const DevLinkMat& m,
                                            const float* __restrict__ ce, const float* __restrict__ cs,
                                            LinkState& st, d3& fN, d3& mN, d3& fP, d3& mP)
{
    d3 old_pos2 = st.pos2, old_a1 = st.a1v, old_a2 = st.a2v;

    // --- orientLink
    d3 pos2 = to_axis_x(axis, pP - pN);
    q4 ang1 = to_axis_x(axis, oN);
    q4 ang2 = to_axis_x(axis, oP);
    q4 total = qconj(ang1);
    pos2 = qrot(total, pos2);
    ang2 = qmul(total, ang2);
    ang1 = qident();

    float small_turn = (float)ddiv(fabs(pos2.z) + fabs(pos2.y), pos2.x);
    float extend = (float)(fabs(1 - pos2.x / rest_len));
    const float HYST = 1.2f, BEND = 0.05f, EXT = 0.50f;                     // src/VX_Link.cpp:21-23
    if (!st.small_angle && small_turn < BEND && extend < EXT) { st.small_angle = true; st.vel_valid = false; }
    else if (st.small_angle && (small_turn > HYST * BEND || extend > HYST * EXT)) { st.small_angle = false; st.vel_valid = false; }

    d3 a1v;
    if (st.small_angle) { pos2.x -= rest_len; a1v = mk3(0.0, 0.0, 0.0); }
    else {
        ang1 = q_align_to_x(pos2);
        ang2 = qmul(ang1, ang2);
        pos2 = mk3(sqrt(norm2(pos2)) - rest_len, 0.0, 0.0);
        a1v = q_to_rotvec(ang1);
    }
    d3 a2v = q_to_rotvec(ang2);
    st.pos2 = pos2; st.a1v = a1v; st.a2v = a2v;

    // --- updateForces
    d3 d_pos2 = 0.5 * (pos2 - old_pos2);
    d3 d_a1 = 0.5 * (a1v - old_a1);
    d3 d_a2 = 0.5 * (a2v - old_a2);

    st.stress = link_update_strain(st, m, ce, cs, (float)ddiv(pos2.x, rest_len), t_sum);
    if (mat_failed(m, st.max_strain)) {
        fN = mN = fP = mP = mk3(0.0, 0.0, 0.0);
        return;
    }

    const double b1 = m.b1, b2 = m.b2, b3 = m.b3, a2 = m.a2;
    fN = mk3(st.stress * t_area,
             b1 * pos2.y - b2 * (a1v.z + a2v.z),
             b1 * pos2.z + b2 * (a1v.y + a2v.y));
    fP = -fN;
    mN = mk3(a2 * (a2v.x - a1v.x),
             -b2 * pos2.z - b3 * (2 * a1v.y + a2v.y),
             b2 * pos2.y - b3 * (2 * a1v.z + a2v.z));
    mP = mk3(a2 * (a1v.x - a2v.x),
             -b2 * pos2.z - b3 * (a1v.y + 2 * a2v.y),
             b2 * pos2.y - b3 * (a1v.z + 2 * a2v.z));

    if (st.vel_valid) {
        const double sqA1 = m.sq_a1, sqA2 = m.sq_a2_ip, sqB1 = m.sq_b1, sqB2 = m.sq_b2_fmp, sqB3 = m.sq_b3_ip;
        d3 pc = mk3(sqA1 * d_pos2.x,
                    sqB1 * d_pos2.y - sqB2 * (d_a1.z + d_a2.z),
                    sqB1 * d_pos2.z + sqB2 * (d_a1.y + d_a2.y));
        fN = fN + (double)damp_n * pc;
        fP = fP - (double)damp_p * pc;
        d3 rn = mk3(-sqA2 * (d_a2.x - d_a1.x),
                    sqB2 * d_pos2.z + sqB3 * (2 * d_a1.y + d_a2.y),
                    -sqB2 * d_pos2.y + sqB3 * (2 * d_a1.z + d_a2.z));
        d3 rp = mk3(sqA2 * (d_a2.x - d_a1.x),
                    sqB2 * d_pos2.z + sqB3 * (d_a1.y + 2 * d_a2.y),
                    -sqB2 * d_pos2.y + sqB3 * (d_a1.z + 2 * d_a2.z));
        mN = mN - (0.5 * damp_n) * rn;
        mP = mP - (0.5 * damp_p) * rp;
    } else st.vel_valid = true;

    if (!st.small_angle) { fN = qrot_inv(ang1, fN); mN = qrot_inv(ang1, mN); }
    fP = qrot_inv(ang2, fP);
    mP = qrot_inv(ang2, mP);
    fN = to_axis_original(axis, fN); fP = to_axis_original(axis, fP);
    mN = to_axis_original(axis, mN); mP = to_axis_original(axis, mP);
}

// ---------------------------------------------------------------------------------------------
// voxel integration: CVX_Voxel::timeStep / force / moment / floorForce (src/VX_Voxel.cpp:162-298)
struct VoxelState { d3 pos; q4 orient; d3 lin, ang; float temp; uint32_t bits; };

__device__ __forceinline__ double base_size(const DevVoxMat& m, int axis, float temp) { return m.size[axis] * (1 + temp * m.cte); }

// F, M: link force / moment sums in the voxel's local frame (slot order already applied);
// contact_refs: this voxel's watched collisions in creation order, ref = 2*pair + (1 if this voxel
// is the pair's second voxel); each float force is subtracted on its own like the reference's
// loop over colWatch (src/VX_Voxel.cpp:249-253, src/VX_Collision.cpp:34-39).
__device__ __forceinline__ void voxel_integrate(VoxelState& v, d3 F, d3 M, const int* __restrict__ contact_refs, int n_contacts,
                                                const float4* __restrict__ contact_force,
                                                const DevVoxMat& m, const DevExt* __restrict__ ext,
                                                float dt, bool floor_on)
{
    floor_on = (floor_on && !(v.bits & VM_FLOOR_OFF)) || (v.bits & VM_FLOOR_ON);      // per-voxel FLOOR_ENABLED, include/VX_Voxel.h:119-120
    const uint32_t dof = ext ? (ext->dof & 0x3Fu) : 0u;
    if (dof == 0x3Fu) {                                       // fully fixed: pose prescribed
        v.pos = mk3(ext->nominal[0] + ext->translation[0], ext->nominal[1] + ext->translation[1], ext->nominal[2] + ext->translation[2]);
        v.orient.w = ext->rot_q[0]; v.orient.x = ext->rot_q[1]; v.orient.y = ext->rot_q[2]; v.orient.z = ext->rot_q[3];
        v.lin = v.ang = mk3(0.0, 0.0, 0.0);
        return;
    }
    // force()
    d3 tot = qrot(v.orient, F);
    if (ext) { tot.x += ext->force[0]; tot.y += ext->force[1]; tot.z += ext->force[2]; }
    d3 vel = m.mass_inv_d * v.lin;
    tot = tot - m.glob_damp_t_d * vel;
    tot.z += m.gravity_force_d;
    for (int k = 0; k < n_contacts; k++) {
        const int ref = contact_refs[k];
        const float4 cf = contact_force[ref >> 1];
        if (ref & 1) { tot.x -= -cf.x; tot.y -= -cf.y; tot.z -= -cf.z; }
        else { tot.x -= cf.x; tot.y -= cf.y; tot.z -= cf.z; }
    }

    d3 fric = tot;
    bool static_fric = (v.bits & VM_STATIC_FRIC) != 0;
    float pen = 0.0f;
    if (floor_on) {                                                  // floorForce()
        double bs_avg = (base_size(m, 0, v.temp) + base_size(m, 1, v.temp) + base_size(m, 2, v.temp)) / 3.0f;
        pen = (float)(bs_avg / 2 - m.nom / 2 - v.pos.z);
        if (pen >= 0) {
            float normal = m.pen_stiff * pen;
            tot.z += normal - m.coll_damp_t_d * vel.z;
            if (static_fric) {
                float surf = (float)(tot.x * tot.x + tot.y * tot.y);
                float lim = (m.mu_s * normal) * (m.mu_s * normal);
                if (surf > lim) static_fric = false;
            } else {
                double hl = sqrt(vel.x * vel.x + vel.y * vel.y + 0.0 * 0.0);
                d3 hn = mk3(vel.x, vel.y, 0.0);
                if (hl > 0) { double inv = 1.0 / hl; hn = inv * hn; }
                float mk = m.mu_k * normal;
                tot = tot - (double)mk * hn;
            }
        } else static_fric = false;
    }
    fric = tot - fric;

    v.lin = v.lin + (double)dt * tot;
    d3 tr = (double)(dt * m.mass_inv) * v.lin;
    if (floor_on && pen >= 0) {
        double work = fric.x * tr.x + fric.y * tr.y;
        double hke = 0.5 * m.mass_inv_d * (v.lin.x * v.lin.x + v.lin.y * v.lin.y);
        if (hke + work <= 0) static_fric = true;
        if (static_fric) { v.lin.x = v.lin.y = 0.0; tr.x = tr.y = 0.0; }
    } else static_fric = false;
    v.pos = v.pos + tr;

    // moment()
    d3 mom = qrot(v.orient, M);
    if (ext) { mom.x += ext->moment[0]; mom.y += ext->moment[1]; mom.z += ext->moment[2]; }
    d3 avel = m.inertia_inv_d * v.ang;
    mom = mom - m.glob_damp_r_d * avel;
    v.ang = v.ang + (double)dt * mom;
    v.orient = qmul(q_from_rotvec((double)(dt * m.inertia_inv) * v.ang), v.orient);

    if (ext && dof) {
        if (dof & 0x01) { v.pos.x = ext->nominal[0] + ext->translation[0]; v.lin.x = 0.0; }
        if (dof & 0x02) { v.pos.y = ext->nominal[1] + ext->translation[1]; v.lin.y = 0.0; }
        if (dof & 0x04) { v.pos.z = ext->nominal[2] + ext->translation[2]; v.lin.z = 0.0; }
        if (dof & 0x38) {
            if ((dof & 0x38) == 0x38) {
                v.orient.w = ext->rot_q[0]; v.orient.x = ext->rot_q[1]; v.orient.y = ext->rot_q[2]; v.orient.z = ext->rot_q[3];
                v.ang = mk3(0.0, 0.0, 0.0);
            } else {
                d3 rv = q_to_rotvec(v.orient);
                if (dof & 0x08) { rv.x = 0.0; v.ang.x = 0.0; }
                if (dof & 0x10) { rv.y = 0.0; v.ang.y = 0.0; }
                if (dof & 0x20) { rv.z = 0.0; v.ang.z = 0.0; }
                v.orient = q_from_rotvec(rv);
            }
        }
    }
    v.bits = (v.bits & ~VM_STATIC_FRIC) | (static_fric ? VM_STATIC_FRIC : 0u) | VM_PSTRAIN_STALE;
}

} // namespace vxd
