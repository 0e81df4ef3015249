// VX_MaterialLink.h -- drop-in CVX_MaterialLink (reference include/VX_MaterialLink.h:24-50): the
// series combination of two voxel materials and the beam constants of the link between them.
#ifndef VXB200_VX_MATERIALLINK_H
#define VXB200_VX_MATERIALLINK_H

#include "VX_MaterialVoxel.h"

class CVX_MaterialLink : public CVX_MaterialVoxel {
public:
    CVX_MaterialLink(CVX_MaterialVoxel* mat1, CVX_MaterialVoxel* mat2);
    CVX_MaterialLink(const CVX_MaterialLink& o) : CVX_MaterialVoxel(o) { *this = o; }
    virtual CVX_MaterialLink& operator=(const CVX_MaterialLink& o);

    // beam constants (protected members _a1.._b3 in the reference, exposed read-only here)
    float a1() const { return k_.a1; }
    float a2() const { return k_.a2; }
    float b1() const { return k_.b1; }
    float b2() const { return k_.b2; }
    float b3() const { return k_.b3; }
    bool updateAll();       // re-derive from the two constituent materials

protected:
    virtual bool updateDerived();
    CVX_MaterialVoxel* vox1Mat;
    CVX_MaterialVoxel* vox2Mat;
    vxm::BeamConsts k_;
    friend class CVoxelyze;
    friend class CVX_Link;
};

#endif // VXB200_VX_MATERIALLINK_H
