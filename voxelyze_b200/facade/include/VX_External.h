// VX_External.h -- drop-in CVX_External (reference include/VX_External.h:18-100): fixed degrees of
// freedom, prescribed displacement, applied force/moment of one voxel.
// Callers mutate this object through the pointer CVX_Voxel::external() returns, with no call into
// CVoxelyze; every setter therefore bumps a counter owned by the simulation (if attached) so that
// the next doTimeStep re-uploads the externals table.
#ifndef VXB200_VX_EXTERNAL_H
#define VXB200_VX_EXTERNAL_H

#include <cstdint>
#include "Vec3D.h"
#include "Quat3D.h"

typedef unsigned char dofObject;
enum dofComponent {
    X_TRANSLATE = 1 << 0,
    Y_TRANSLATE = 1 << 1,
    Z_TRANSLATE = 1 << 2,
    X_ROTATE = 1 << 3,
    Y_ROTATE = 1 << 4,
    Z_ROTATE = 1 << 5
};
inline void dofSet(dofObject& obj, dofComponent dof, bool set) { if (set) obj |= dof; else obj &= ~dof; }
inline void dofSetAll(dofObject& obj, bool set) { if (set) obj |= 0x3F; else obj &= ~0x3F; }
inline bool dofIsSet(dofObject obj, dofComponent dof) { return (dof & obj) != 0; }
inline bool dofIsAllSet(dofObject obj) { return (obj & 0x3F) == 0x3F; }
inline bool dofIsNoneSet(dofObject obj) { return !(obj & 0x3F); }
inline dofObject dof(bool tx, bool ty, bool tz, bool rx, bool ry, bool rz)
{
    dofObject r = 0;
    dofSet(r, X_TRANSLATE, tx); dofSet(r, Y_TRANSLATE, ty); dofSet(r, Z_TRANSLATE, tz);
    dofSet(r, X_ROTATE, rx); dofSet(r, Y_ROTATE, ry); dofSet(r, Z_ROTATE, rz);
    return r;
}

class CVX_External {
public:
    CVX_External() { reset(); }
    CVX_External(const CVX_External& e) { *this = e; }
    CVX_External& operator=(const CVX_External& e);
    bool operator==(const CVX_External& b) const
    { return dofFixed == b.dofFixed && extForce == b.extForce && extMoment == b.extMoment && extTranslation == b.extTranslation && extRotation == b.extRotation; }

    void reset();
    bool isEmpty() { return dofFixed == 0 && extForce == Vec3D<float>() && extMoment == Vec3D<float>(); }

    bool isFixed(dofComponent d) const { return dofIsSet(dofFixed, d); }
    bool isFixedAll() const { return dofIsAllSet(dofFixed); }
    bool isFixedAllTranslation() const { return (dofFixed & 0x07) == 0x07; }
    bool isFixedAllRotation() const { return (dofFixed & 0x38) == 0x38; }
    bool isFixedAny() const { return dofFixed != 0; }
    bool isFixedAnyTranslation() const { return (dofFixed & 0x07) != 0; }
    bool isFixedAnyRotation() const { return (dofFixed & 0x38) != 0; }

    Vec3D<double> translation() const { return extTranslation; }
    Vec3D<double> rotation() const { return extRotation; }
    Quat3D<double> rotationQuat() const { return rotQ; }

    void setFixed(bool xTranslate, bool yTranslate, bool zTranslate, bool xRotate, bool yRotate, bool zRotate);
    void setFixed(dofComponent d, bool fixed = true) { if (fixed) setDisplacement(d); else clearDisplacement(d); }
    void setFixedAll(bool fixed = true) { if (fixed) setDisplacementAll(); else clearDisplacementAll(); }
    void setDisplacement(dofComponent d, double displacement = 0.0);
    void setDisplacementAll(const Vec3D<double>& translation = Vec3D<double>(0, 0, 0), const Vec3D<double>& rotation = Vec3D<double>(0, 0, 0));
    void clearDisplacement(dofComponent d);
    void clearDisplacementAll();

    Vec3D<float> force() const { return extForce; }
    Vec3D<float> moment() const { return extMoment; }
    void setForce(const float x, const float y, const float z) { extForce = Vec3D<float>(x, y, z); touch(); }
    void setForce(const Vec3D<float>& f) { extForce = f; touch(); }
    void setMoment(const float x, const float y, const float z) { extMoment = Vec3D<float>(x, y, z); touch(); }
    void setMoment(const Vec3D<float>& m) { extMoment = m; touch(); }
    void addForce(const float x, const float y, const float z) { extForce += Vec3D<float>(x, y, z); touch(); }
    void addForce(const Vec3D<float>& f) { extForce += f; touch(); }
    void addMoment(const float x, const float y, const float z) { extMoment += Vec3D<float>(x, y, z); touch(); }
    void addMoment(const Vec3D<float>& m) { extMoment += m; touch(); }
    void clearForce() { extForce = Vec3D<float>(); touch(); }
    void clearMoment() { extMoment = Vec3D<float>(); touch(); }

    dofObject dofMask() const { return dofFixed; }
    void attach(uint64_t* counter) { changeCounter = counter; touch(); }

private:
    void touch() { if (changeCounter) ++*changeCounter; }
    void rotationChanged();
    dofObject dofFixed;
    Vec3D<float> extForce, extMoment;
    Vec3D<double> extTranslation, extRotation;
    Quat3D<double> rotQ;
    uint64_t* changeCounter = nullptr;
};

#endif // VXB200_VX_EXTERNAL_H
