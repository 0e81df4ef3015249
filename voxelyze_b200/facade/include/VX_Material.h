// VX_Material.h -- drop-in CVX_Material of the voxelyze_b200 facade.
//
// Mirrors the public interface of the reference's include/VX_Material.h:33-103 (same method
// names, argument meaning, clamping and lastError() strings).  The numerical content lives in
// vxm::Material (csrc/vx_material.hpp), which is also what the CUDA library derives its device
// tables from, so an accessor here and a kernel there can never disagree.  JSON constructors of
// the reference are out of scope.
#ifndef VXB200_VX_MATERIAL_H
#define VXB200_VX_MATERIAL_H

#include <string>
#include <vector>
#include <cstdint>
#include "Vec3D.h"
#include "vx_material.hpp"

class CVoxelyze;

class CVX_Material {
public:
    CVX_Material(float youngsModulus = 1e6f, float density = 1e3f);
    virtual ~CVX_Material() {}
    CVX_Material(const CVX_Material& o) { *this = o; }
    virtual CVX_Material& operator=(const CVX_Material& o);

    void clear();
    const char* lastError() const { return m_.error.c_str(); }
    void setName(const char* name) { name_ = name; }
    const char* name() const { return name_.c_str(); }

    float stress(float strain, float transverseStrainSum = 0.0f, bool forceLinear = false) { return m_.stress(strain, transverseStrainSum, forceLinear); }
    float modulus(float strain) { return m_.modulus(strain); }
    bool isYielded(float strain) { return m_.yielded(strain); }
    bool isFailed(float strain) { return m_.failed(strain); }

    void setColor(int red, int green, int blue, int alpha = 255) { setRed(red); setGreen(green); setBlue(blue); setAlpha(alpha); }
    void setRed(int v) { r_ = clamp255(v); }
    void setGreen(int v) { g_ = clamp255(v); }
    void setBlue(int v) { b_ = clamp255(v); }
    void setAlpha(int v) { a_ = clamp255(v); }
    int red() const { return r_; }
    int green() const { return g_; }
    int blue() const { return b_; }
    int alpha() const { return a_; }

    bool setModel(int dataPointCount, float* pStrainValues, float* pStressValues);
    bool setModelLinear(float youngsModulus, float failureStress = -1);
    bool setModelBilinear(float youngsModulus, float plasticModulus, float yieldStress, float failureStress = -1);
    bool isModelLinear() const { return m_.linear; }

    float youngsModulus() const { return m_.E; }
    float yieldStress() const { return m_.sigma_yield; }
    float failureStress() const { return m_.sigma_fail; }
    int modelDataPoints() const { return (int)m_.eps.size(); }
    const float* modelDataStrain() const { return &m_.eps[0]; }
    const float* modelDataStress() const { return &m_.sig[0]; }

    void setPoissonsRatio(float poissonsRatio);
    float poissonsRatio() const { return m_.nu; }
    float bulkModulus() const { return m_.E / (3 * (1 - 2 * m_.nu)); }
    float lamesFirstParameter() const { return (m_.E * m_.nu) / ((1 + m_.nu) * (1 - 2 * m_.nu)); }
    float shearModulus() const { return m_.E / (2 * (1 + m_.nu)); }
    bool isXyzIndependent() const { return m_.nu == 0.0f; }

    void setDensity(float density);
    float density() const { return m_.rho; }
    void setStaticFriction(float c) { m_.mu_s = c <= 0 ? 0.0f : c; changed(); }
    float staticFriction() const { return m_.mu_s; }
    void setKineticFriction(float c) { m_.mu_k = c <= 0 ? 0.0f : c; changed(); }
    float kineticFriction() const { return m_.mu_k; }
    void setInternalDamping(float zeta) { m_.zeta_int = zeta <= 0 ? 0.0f : zeta; changed(); }
    float internalDamping() const { return m_.zeta_int; }
    void setGlobalDamping(float zeta) { m_.zeta_glob = zeta <= 0 ? 0.0f : zeta; changed(); }
    float globalDamping() const { return m_.zeta_glob; }
    void setCollisionDamping(float zeta) { m_.zeta_coll = zeta <= 0 ? 0.0f : zeta; changed(); }
    float collisionDamping() const { return m_.zeta_coll; }

    void setExternalScaleFactor(Vec3D<double> factor);
    void setExternalScaleFactor(double factor) { setExternalScaleFactor(Vec3D<double>(factor, factor, factor)); }
    Vec3D<double> externalScaleFactor() { return Vec3D<double>(m_.ext_scale[0], m_.ext_scale[1], m_.ext_scale[2]); }

    void setCte(float cte) { m_.cte = cte; changed(); }
    float cte() const { return m_.cte; }

    // facade plumbing: the flat description the C-ABI takes, and a change counter the owning
    // CVoxelyze compares at doTimeStep entry (material setters carry no other notification)
    const vxm::Material& model() const { return m_; }
    uint64_t changeCount() const { return changes_; }

protected:
    virtual void changed() { changes_++; updateDerived(); }
    virtual bool updateDerived() { m_.refresh_e_hat(); return true; }
    static int clamp255(int v) { return v > 255 ? 255 : (v < 0 ? 0 : v); }

    vxm::Material m_;
    std::string name_;
    int r_ = -1, g_ = -1, b_ = -1, a_ = -1;
    uint64_t changes_ = 0;
    friend class CVoxelyze;
};

#endif // VXB200_VX_MATERIAL_H
