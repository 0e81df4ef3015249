// vx_types.h -- data layout shared by host and device code of libvoxelyze_b200.so.
//
// HBM layout (structure of arrays, internal order = voxels sorted by (member, z, y, x);
// links sorted by axis then by negative-end voxel):
//
//   voxel  pose0[v]  double4  { pos.x, pos.y, pos.z, orient.w }             32 B
//          pose1[v]  double4  { orient.x, orient.y, orient.z, meta }        32 B
//                    meta = 64 bits: low 32 = temperature (float bits),
//                                    high 32 = material | link mask | flags (VM_*)
//          mom0[v]   double4  { linMom.x, linMom.y, linMom.z, angMom.x }     32 B
//          mom1[v]   double2  { angMom.y, angMom.z }                         16 B
//          => 112 B read + 112 B written per voxel per step; a link reads both 64 B pose
//             records of its end voxels with four 128-bit loads each.
//   link   lstA[l]   double4  { pos2.x, pos2.y, pos2.z, angle1v.x }          32 B
//          lstB[l]   double4  { angle1v.y, angle1v.z, angle2v.x, angle2v.y } 32 B
//          lstC[l]   double   { angle2v.z }                                   8 B
//          lstrain[l] float4  { strain, maxStrain, strainOffset, stress }    16 B
//          lmeta[l]  uint32   link material | small-angle | local-velocity-valid bits
//          lends[l]  int2     { negative-end voxel, positive-end voxel }
//   force  slot[s][v] 6 doubles { force.xyz, moment.xyz } acting on voxel v through its
//          link slot s (0 X+,1 X-,2 Y+,3 Y-,4 Z+,5 Z-): the link kernel scatters, the voxel
//          kernel gathers by fixed index in reference summation order, no atomics.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VX_HD __host__ __device__
#else
#define VX_HD
#endif

// bits of the high word of the voxel meta
#define VM_MAT_MASK      0x3FFu          // bits 0-9   voxel material (max 1024)
#define VM_FLOOR_OFF     (1u << 10)      // CVX_Voxel::enableFloor(false) on this voxel while the simulation's floor is on
#define VM_FLOOR_ON      (1u << 11)      // CVX_Voxel::enableFloor(true) on this voxel while the simulation's floor is off
#define VM_LINK_SHIFT    16              // bits 16-21 link slot present
#define VM_STATIC_FRIC   (1u << 22)      // FLOOR_STATIC_FRICTION
#define VM_HAS_EXT       (1u << 23)      // entry in the externals table
#define VM_GHOST         (1u << 24)      // halo copy, never integrated
#define VM_PSTRAIN_STALE (1u << 25)      // poissonsStrainInvalid
#define VX_MAX_VOXMATS   1024

// link meta bits
#define LM_MAT_MASK      0xFFFFu
#define LM_SMALL_ANGLE   (1u << 16)
#define LM_VEL_VALID     (1u << 17)

// per voxel material, everything the kernels read (floats already combined on the host in the
// reference's evaluation order)
struct DevVoxMat {
    double size[3];        // nominal size * external scale                       VX_MaterialVoxel.h:34
    double nom;            // nominal size
    float  cte;
    float  mass, mass_inv, inertia_inv;
    float  E, nu;
    float  two_sqrtm_zeta; // (2*_sqrtMass)*zetaInternal                          VX_Voxel.h:130
    float  glob_damp_t;    // zetaGlobal*_2xSqMxExS                               VX_MaterialVoxel.h:43
    float  glob_damp_r;    // zetaGlobal*_2xSqIxExSxSxS                           VX_MaterialVoxel.h:44
    float  coll_damp_t;    // zetaCollision*_2xSqMxExS                            VX_MaterialVoxel.h:45
    float  pen_stiff;      // (float)(2*E*nomSize)                                VX_MaterialVoxel.h:49
    float  mu_s, mu_k;
    float  gravity_force;  // -_mass*9.80665f*gravMult                            VX_MaterialVoxel.h:57
    float  nom_f;          // (float)nomSize
    float  pad;
    // exact double copies of the floats that are only ever used after promotion to double
    double mass_inv_d, inertia_inv_d, glob_damp_t_d, glob_damp_r_d, coll_damp_t_d, gravity_force_d;
};

// per link material
struct DevLinkMat {
    int32_t linear;
    int32_t curve_off, curve_n;   // into the shared curve arrays (incl. the (0,0) point)
    float   E, nu, e_hat;
    float   eps_yield, eps_fail;
    float   a1;
    float   pad;
    // beam constants: float values of the reference, stored promoted (they are only used in
    // double expressions, src/VX_Link.cpp:165-195)
    double  a2, b1, b2, b3;
    double  sq_a1, sq_a2_ip, sq_b1, sq_b2_fmp, sq_b3_ip;
};

// externals table entry (sparse: only voxels with a CVX_External)
struct DevExt {
    double nominal[3];     // ix*size, iy*size, iz*size
    double translation[3];
    double rot_q[4];       // w,x,y,z of the prescribed rotation
    float  force[3];
    float  moment[3];
    uint32_t dof;          // VX_DOF_* bits (dofObject, include/VX_External.h:18)
    uint32_t pad;
};

// scalars that live in device memory so that captured graphs stay valid when they change
struct DevParams {
    float dt;              // time step of the step being executed
    float prev_dt;         // CVX_Voxel::previousDt
    float time;            // CVoxelyze::currentTime (float accumulation)
    int   div_now;         // a link of the current step exceeded strain 100
    int   div_latched;     // a previous step diverged: everything is frozen
    int   steps_done;      // completed steps since the last vx_step call
    int   col_stale;       // collision watch list must be rebuilt
    int   pending;         // fused path: the previous step still has to be counted
    int   div_flag[2];     // fused path: divergence flag of the step with that generation parity
    float last_prev;       // fused path: DevParams::prev_dt as the last step of the finished call saw it (calls that change dt every step)
};
