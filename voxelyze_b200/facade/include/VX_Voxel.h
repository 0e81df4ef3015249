// VX_Voxel.h -- drop-in CVX_Voxel handle (reference include/VX_Voxel.h:35-139).
// Holds the model-side data of a voxel (lattice index, material, lazily created CVX_External);
// the dynamic state (pose, momenta, temperature) lives in HBM and is mirrored on demand when an
// accessor is called (CVoxelyze keeps the mirror fresh per step).  The integrator itself is
// csrc/vx_physics.cuh: voxel_integrate.
#ifndef VXB200_VX_VOXEL_H
#define VXB200_VX_VOXEL_H

#include "Vec3D.h"
#include "Quat3D.h"
#include "VX_Link.h"
#include "VX_External.h"
#include "VX_MaterialVoxel.h"
#include "VX_Collision.h"

class CVoxelyze;

class CVX_Voxel {
public:
    enum linkDirection { X_POS = 0, X_NEG = 1, Y_POS = 2, Y_NEG = 3, Z_POS = 4, Z_NEG = 5 };
    enum voxelCorner { NNN = 0, NNP = 1, NPN = 2, NPP = 3, PNN = 4, PNP = 5, PPN = 6, PPP = 7 };

    CVX_Voxel(CVX_MaterialVoxel* material, short indexX, short indexY, short indexZ);
    ~CVX_Voxel();

    CVX_Link* link(linkDirection direction) const { return links[direction]; }
    int linkCount() const { int n = 0; for (int i = 0; i < 6; i++) if (links[i]) n++; return n; }
    CVX_Voxel* adjacentVoxel(linkDirection direction) const;
    short indexX() { return ix; }
    short indexY() { return iy; }
    short indexZ() { return iz; }
    CVX_MaterialVoxel* material() { return mat; }

    bool externalExists() { return ext != nullptr; }
    CVX_External* external();

    Vec3D<double> position() const;
    Vec3D<double> originalPosition() const { double s = mat->nominalSize(); return Vec3D<double>(ix * s, iy * s, iz * s); }
    Vec3D<double> displacement() const { return position() - originalPosition(); }
    Vec3D<float> size() const { return cornerOffset(PPP) - cornerOffset(NNN); }       // VX_Voxel.h:82
    Vec3D<float> cornerPosition(voxelCorner corner) const;                               // VX_Voxel.cpp:141-144
    Vec3D<float> cornerOffset(voxelCorner corner) const;                                 // VX_Voxel.cpp:146-159
    bool isInterior() const { return linkCount() == 6; }
    bool isSurface() const { return !isInterior(); }

    Vec3D<double> baseSize() const { return mat->size() * (1 + temperatureValue() * mat->cte()); }
    double baseSize(CVX_Link::linkAxis axis) const { return mat->size()[axis] * (1 + temperatureValue() * mat->cte()); }
    double baseSizeAverage() const { Vec3D<double> b = baseSize(); return (b.x + b.y + b.z) / 3.0f; }

    Quat3D<double> orientation() const;
    float orientationAngle() const { return (float)orientation().Angle(); }
    Vec3D<double> orientationAxis() const { return orientation().Axis(); }
    float displacementMagnitude() const { return (float)displacement().Length(); }
    float angularDisplacementMagnitude() const { return (float)orientation().Angle(); }
    Vec3D<double> linearMomentum() const;
    Vec3D<double> angularMomentum() const;
    Vec3D<double> velocity() const { return linearMomentum() * (double)mat->massProps().mass_inv; }
    float velocityMagnitude() const { return (float)(linearMomentum().Length() * mat->massProps().mass_inv); }
    Vec3D<double> angularVelocity() const { return angularMomentum() * (double)mat->massProps().inertia_inv; }
    float angularVelocityMagnitude() const { return (float)(angularMomentum().Length() * mat->massProps().inertia_inv); }
    float kineticEnergy() const
    { return (float)(0.5 * (mat->massProps().mass_inv * linearMomentum().Length2() + mat->massProps().inertia_inv * angularMomentum().Length2())); }

    float volumetricStrain() const { Vec3D<float> s = strain(false); return (float)(s.x + s.y + s.z); }     // VX_Voxel.h:103
    float pressure() const { return -mat->youngsModulus() * volumetricStrain() / (3 * (1 - 2 * mat->poissonsRatio())); }

    bool isYielded() const;
    bool isFailed() const;

    float temperature() { return temperatureValue(); }
    void setTemperature(float temperature);
    void haltMotion();

    Vec3D<float> externalForce();                  // applied force, or the reaction on fixed degrees of freedom (VX_Voxel.cpp:115-125)
    Vec3D<float> externalMoment();
    Vec3D<double> force();                         // sum of the forces on this voxel, GCS (VX_Voxel.cpp:234-256)
    Vec3D<double> moment();
    float transverseArea(CVX_Link::linkAxis axis);
    float transverseStrainSum(CVX_Link::linkAxis axis);
    Vec3D<float> strain(bool poissonsStrain) const;            // LCS voxel strain (VX_Voxel.cpp:300-334; private in the reference)
    void enableFloor(bool enabled);                // include/VX_Voxel.h:119: this voxel only; CVoxelyze::enableFloor sets every voxel again
    bool isFloorEnabled() const;                   // include/VX_Voxel.h:120
    float dampingMultiplier();                     // include/VX_Voxel.h:130: 2 sqrt(m) zeta_internal / previousDt

    bool isFloorStaticFriction() const;
    float floorPenetration() const { return (float)(baseSizeAverage() / 2 - mat->nominalSize() / 2 - position().z); }

    static inline CVX_Link::linkAxis toAxis(linkDirection d) { return (CVX_Link::linkAxis)((int)d / 2); }
    static inline linkDirection toDirection(CVX_Link::linkAxis axis, bool positiveDirection) { return (linkDirection)(2 * (int)axis + (positiveDirection ? 0 : 1)); }
    static inline bool isNegative(linkDirection d) { return d % 2 == 1; }
    static inline bool isPositive(linkDirection d) { return d % 2 == 0; }
    static inline linkDirection toOpposite(linkDirection d) { return (linkDirection)(d - d % 2 + (d + 1) % 2); }

private:
    float temperatureValue() const;
    CVoxelyze* sim = nullptr;       // null for a stand-alone voxel (reference test/tVX_Voxel.h)
    int index = -1;                 // voxel index of the C-ABI (creation order)
    int floorOverride = -1;         // enableFloor() on this voxel: 0 / 1, -1 = follows the simulation
    CVX_MaterialVoxel* mat;
    short ix, iy, iz;
    CVX_External* ext = nullptr;
    CVX_Link* links[6];
    // stand-alone state (no simulation attached)
    Vec3D<double> pos0; float temp0 = 0.0f;
    friend class CVoxelyze;
    friend class CVX_Link;
};

#endif // VXB200_VX_VOXEL_H
