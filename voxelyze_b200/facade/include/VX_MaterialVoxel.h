// VX_MaterialVoxel.h -- drop-in CVX_MaterialVoxel (reference include/VX_MaterialVoxel.h:23-76):
// a material plus the nominal voxel size, giving mass and the cached damping/contact constants.
#ifndef VXB200_VX_MATERIALVOXEL_H
#define VXB200_VX_MATERIALVOXEL_H

#include "VX_Material.h"

class CVX_MaterialVoxel : public CVX_Material {
public:
    CVX_MaterialVoxel(float youngsModulus = 1e6f, float density = 1e3f, double nominalSize = 0.001);
    CVX_MaterialVoxel(const CVX_Material& mat, double nominalSize = 0.001);
    CVX_MaterialVoxel(const CVX_MaterialVoxel& o) : CVX_Material(o) { *this = o; }
    virtual CVX_MaterialVoxel& operator=(const CVX_MaterialVoxel& o);

    bool setNominalSize(double size);
    double nominalSize() { return nom_; }
    Vec3D<double> size() { return Vec3D<double>(nom_ * m_.ext_scale[0], nom_ * m_.ext_scale[1], nom_ * m_.ext_scale[2]); }

    float mass() { return p_.mass; }
    float momentInertia() { return p_.inertia; }

    float internalDampingTranslateC() const { return m_.zeta_int * p_.two_sq_mes; }
    float internalDampingRotateC() const { return m_.zeta_int * p_.two_sq_ies3; }
    float globalDampingTranslateC() const { return m_.zeta_glob * p_.two_sq_mes; }
    float globalDampingRotateC() const { return m_.zeta_glob * p_.two_sq_ies3; }
    float collisionDampingTranslateC() const { return m_.zeta_coll * p_.two_sq_mes; }
    float collisionDampingRotateC() const { return m_.zeta_coll * p_.two_sq_ies3; }
    float penetrationStiffness() const { return (float)(2 * m_.E * nom_); }

    const vxm::MassProps& massProps() const { return p_; }

protected:
    virtual bool updateDerived();
    double nom_ = 0.001;
    float gravMult_ = 0.0f;
    vxm::MassProps p_;
    friend class CVoxelyze;
    friend class CVX_Voxel;
    friend class CVX_MaterialLink;
};

#endif // VXB200_VX_MATERIALVOXEL_H
