// VX_Link.h -- drop-in CVX_Link handle (reference include/VX_Link.h:33-61).
// The beam computation itself runs in the CUDA kernels (csrc/vx_physics.cuh: link_forces); this
// object is a stable handle whose accessors read the link's device state through the C-ABI.
#ifndef VXB200_VX_LINK_H
#define VXB200_VX_LINK_H

#include "Vec3D.h"
#include "Quat3D.h"

class CVX_Voxel;
class CVX_MaterialLink;
class CVoxelyze;

class CVX_Link {
public:
    enum linkAxis { X_AXIS = 0, Y_AXIS = 1, Z_AXIS = 2 };

    CVX_Voxel* voxel(bool positiveEnd) const { return positiveEnd ? pVPos : pVNeg; }
    Vec3D<> force(bool positiveEnd) const;
    Vec3D<> moment(bool positiveEnd) const;
    float axialStrain() const;
    float axialStrain(bool positiveEnd) const;
    float axialStress() const;
    bool isSmallAngle() const;
    bool isYielded() const;
    bool isFailed() const;
    float strainEnergy() const;
    float axialStiffness();
    linkAxis axisOf() const { return axis; }

private:
    CVX_Link(CVoxelyze* owner, CVX_Voxel* neg, CVX_Voxel* pos, linkAxis ax) : sim(owner), pVNeg(neg), pVPos(pos), axis(ax) {}
    CVoxelyze* sim;
    CVX_Voxel* pVNeg; CVX_Voxel* pVPos;
    linkAxis axis;
    int index = -1;             // link index of the C-ABI (creation order)
    CVX_MaterialLink* mat = nullptr;
    friend class CVoxelyze;
    friend class CVX_Voxel;
};

#endif // VXB200_VX_LINK_H
