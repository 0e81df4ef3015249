// VX_Collision.h -- drop-in CVX_Collision handle (reference include/VX_Collision.h:26-45): one
// watched pair of voxels.  Watch lists and contact forces are produced on the GPU
// (csrc/vx_collide.cuh); collisionList() materialises these handles on demand.
#ifndef VXB200_VX_COLLISION_H
#define VXB200_VX_COLLISION_H

#include "Vec3D.h"
class CVX_Voxel;

class CVX_Collision {
public:
    CVX_Collision(CVX_Voxel* v1, CVX_Voxel* v2) : pV1(v1), pV2(v2) {}
    CVX_Voxel* voxel1() const { return pV1; }
    CVX_Voxel* voxel2() const { return pV2; }
    Vec3D<float> force() const { return f_; }                       // contact force of the last step on voxel1
    Vec3D<float> contactForce(CVX_Voxel* pVoxel) const              // src/VX_Collision.cpp:34-39
    { return pVoxel == pV1 ? f_ : (pVoxel == pV2 ? -f_ : Vec3D<float>(0, 0, 0)); }
    static float envelopeRadius;    // collision envelope radius in voxel edge lengths (default 0.625)
private:
    CVX_Voxel* pV1; CVX_Voxel* pV2;
    Vec3D<float> f_;
    friend class CVoxelyze;
};

#endif // VXB200_VX_COLLISION_H
