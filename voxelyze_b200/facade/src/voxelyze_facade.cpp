// voxelyze_facade.cpp -- host side of the drop-in C++ API (see facade/include/Voxelyze.h).
//
// Everything here is bookkeeping: it records the model the caller builds through the reference's
// class API, keeps stable CVX_Voxel / CVX_Link / CVX_Material handles, and translates to the flat
// C-ABI of include/voxelyze_b200.h.  No dynamics are computed on the host: doTimeStep is one
// vx_step call, accessors are vx_download calls.  There is no CPU fallback; if the CUDA library
// cannot create a device handle the process stops with a message.
#include "Voxelyze.h"
#include "voxelyze_b200.h"

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// the same call on the single-device handle or on the slabbed one (include/voxelyze_b200.h: vx_slabbed_* mirror their namesakes)
#define VXH(sim, fn, ...) ((sim)->hm ? vx_slabbed_##fn((sim)->hm, __VA_ARGS__) : vx_##fn((sim)->h, __VA_ARGS__))
#define VXH0(sim, fn) ((sim)->hm ? vx_slabbed_##fn((sim)->hm) : vx_##fn((sim)->h))

float CVX_Collision::envelopeRadius = 0.625f;

// ================================================================================================ CVX_Material
CVX_Material::CVX_Material(float youngsModulus, float density)
{
    clear();
    m_.rho = density;
    setModelLinear(youngsModulus);
}

CVX_Material& CVX_Material::operator=(const CVX_Material& o)
{
    m_ = o.m_; name_ = o.name_; r_ = o.r_; g_ = o.g_; b_ = o.b_; a_ = o.a_;
    changes_++;
    return *this;
}

void CVX_Material::clear()
{
    r_ = g_ = b_ = a_ = -1;
    vxm::Material fresh;            // defaults of the reference's clear(): nu 0, rho 1, zeta_internal 1, ...
    m_ = fresh;
    m_.model_linear(1.0f);
    changed();
}

bool CVX_Material::setModel(int n, float* strain, float* stress) { bool ok = m_.model_data(n, strain, stress); if (ok) changed(); return ok; }
bool CVX_Material::setModelLinear(float E, float fail) { bool ok = m_.model_linear(E, fail); if (ok) changed(); return ok; }
bool CVX_Material::setModelBilinear(float E, float plastic, float yield, float fail) { bool ok = m_.model_bilinear(E, plastic, yield, fail); if (ok) changed(); return ok; }

void CVX_Material::setPoissonsRatio(float nu)
{
    if (nu < 0) nu = 0;
    if (nu >= 0.5) nu = 0.5 - FLT_EPSILON * 2;
    m_.nu = nu;
    changed();
}
void CVX_Material::setDensity(float density) { m_.rho = density <= 0 ? FLT_MIN : density; changed(); }
void CVX_Material::setExternalScaleFactor(Vec3D<double> f)
{
    m_.ext_scale[0] = f.x <= 0 ? (double)FLT_MIN : f.x;
    m_.ext_scale[1] = f.y <= 0 ? (double)FLT_MIN : f.y;
    m_.ext_scale[2] = f.z <= 0 ? (double)FLT_MIN : f.z;
    changed();
}

// ================================================================================================ CVX_MaterialVoxel
CVX_MaterialVoxel::CVX_MaterialVoxel(float E, float rho, double nominalSize) : CVX_Material(E, rho) { nom_ = nominalSize; updateDerived(); }
CVX_MaterialVoxel::CVX_MaterialVoxel(const CVX_Material& mat, double nominalSize) : CVX_Material(mat) { nom_ = nominalSize; updateDerived(); }
CVX_MaterialVoxel& CVX_MaterialVoxel::operator=(const CVX_MaterialVoxel& o)
{
    CVX_Material::operator=(o);
    nom_ = o.nom_; gravMult_ = o.gravMult_;
    updateDerived();
    return *this;
}
bool CVX_MaterialVoxel::setNominalSize(double size) { nom_ = size <= 0 ? (double)FLT_MIN : size; changes_++; return updateDerived(); }
bool CVX_MaterialVoxel::updateDerived()
{
    CVX_Material::updateDerived();
    p_ = vxm::mass_props(m_, nom_);
    return p_.mass_inv != 0.0f;
}

// ================================================================================================ CVX_MaterialLink
CVX_MaterialLink::CVX_MaterialLink(CVX_MaterialVoxel* a, CVX_MaterialVoxel* b) : vox1Mat(a), vox2Mat(b) { updateAll(); }
CVX_MaterialLink& CVX_MaterialLink::operator=(const CVX_MaterialLink& o)
{
    CVX_MaterialVoxel::operator=(o);
    vox1Mat = o.vox1Mat; vox2Mat = o.vox2Mat; k_ = o.k_;
    return *this;
}
bool CVX_MaterialLink::updateAll()
{
    nom_ = 0.5 * (vox1Mat->nom_ + vox2Mat->nom_);
    m_ = vxm::combine(vox1Mat->model(), vox2Mat->model());
    r_ = (int)(0.5 * (vox1Mat->r_ + vox2Mat->r_)); g_ = (int)(0.5 * (vox1Mat->g_ + vox2Mat->g_));
    b_ = (int)(0.5 * (vox1Mat->b_ + vox2Mat->b_)); a_ = (int)(0.5 * (vox1Mat->a_ + vox2Mat->a_));
    return updateDerived();
}
bool CVX_MaterialLink::updateDerived()
{
    CVX_MaterialVoxel::updateDerived();
    k_ = vxm::beam_consts(m_, nom_);
    return true;
}

// ================================================================================================ CVX_External
CVX_External& CVX_External::operator=(const CVX_External& e)
{
    dofFixed = e.dofFixed; extForce = e.extForce; extMoment = e.extMoment;
    extTranslation = e.extTranslation; extRotation = e.extRotation;
    rotationChanged();
    return *this;
}
void CVX_External::reset()
{
    dofFixed = 0;
    extForce = extMoment = Vec3D<float>();
    extTranslation = extRotation = Vec3D<double>();
    rotationChanged();
}
void CVX_External::setFixed(bool tx, bool ty, bool tz, bool rx, bool ry, bool rz)
{
    dofFixed = dof(tx, ty, tz, rx, ry, rz);
    extTranslation = extRotation = Vec3D<double>();      // like the reference, the cached quaternion is left as it was
    touch();
}
void CVX_External::setDisplacement(dofComponent d, double displacement)
{
    dofSet(dofFixed, d, true);
    if (displacement != 0.0f) {
        if (d & X_TRANSLATE) extTranslation.x = displacement;
        if (d & Y_TRANSLATE) extTranslation.y = displacement;
        if (d & Z_TRANSLATE) extTranslation.z = displacement;
        if (d & X_ROTATE) extRotation.x = displacement;
        if (d & Y_ROTATE) extRotation.y = displacement;
        if (d & Z_ROTATE) extRotation.z = displacement;
    }
    rotationChanged();
}
void CVX_External::setDisplacementAll(const Vec3D<double>& t, const Vec3D<double>& r)
{
    dofSetAll(dofFixed, true);
    extTranslation = t; extRotation = r;
    rotationChanged();
}
void CVX_External::clearDisplacement(dofComponent d)
{
    dofSet(dofFixed, d, false);
    if (d & X_TRANSLATE) extTranslation.x = 0.0;
    if (d & Y_TRANSLATE) extTranslation.y = 0.0;
    if (d & Z_TRANSLATE) extTranslation.z = 0.0;
    if (d & X_ROTATE) extRotation.x = 0.0;
    if (d & Y_ROTATE) extRotation.y = 0.0;
    if (d & Z_ROTATE) extRotation.z = 0.0;
    rotationChanged();
}
void CVX_External::clearDisplacementAll()
{
    dofSetAll(dofFixed, false);
    extTranslation = extRotation = Vec3D<double>();
    rotationChanged();
}
void CVX_External::rotationChanged()
{
    if (extRotation != Vec3D<double>()) rotQ = Quat3D<double>(extRotation);
    else rotQ = Quat3D<double>();
    touch();
}

// ================================================================================================ CVX_Voxel
CVX_Voxel::CVX_Voxel(CVX_MaterialVoxel* material, short x, short y, short z) : mat(material), ix(x), iy(y), iz(z)
{
    for (int i = 0; i < 6; i++) links[i] = nullptr;
    pos0 = originalPosition();
}
CVX_Voxel::~CVX_Voxel() { delete ext; }

CVX_External* CVX_Voxel::external()
{
    if (!ext) { ext = new CVX_External(); if (sim) ext->attach(&sim->extChanges); }
    return ext;
}
CVX_Voxel* CVX_Voxel::adjacentVoxel(linkDirection d) const
{
    CVX_Link* l = links[d];
    if (!l) return nullptr;
    return l->voxel(true) == this ? l->voxel(false) : l->voxel(true);
}
Vec3D<double> CVX_Voxel::position() const
{
    if (!sim) return pos0;
    sim->fetchVoxel(index);
    return Vec3D<double>(sim->mPos[3 * index], sim->mPos[3 * index + 1], sim->mPos[3 * index + 2]);
}
Quat3D<double> CVX_Voxel::orientation() const
{
    if (!sim) return Quat3D<double>();
    sim->fetchVoxel(index);
    return Quat3D<double>(sim->mOrient[4 * index], sim->mOrient[4 * index + 1], sim->mOrient[4 * index + 2], sim->mOrient[4 * index + 3]);
}
Vec3D<double> CVX_Voxel::linearMomentum() const
{
    if (!sim) return Vec3D<double>();
    sim->fetchVoxel(index);
    return Vec3D<double>(sim->mLin[3 * index], sim->mLin[3 * index + 1], sim->mLin[3 * index + 2]);
}
Vec3D<double> CVX_Voxel::angularMomentum() const
{
    if (!sim) return Vec3D<double>();
    sim->fetchVoxel(index);
    return Vec3D<double>(sim->mAng[3 * index], sim->mAng[3 * index + 1], sim->mAng[3 * index + 2]);
}
float CVX_Voxel::temperatureValue() const
{
    if (!sim) return temp0;
    sim->fetchVoxel(index);
    return sim->mTemp[index];
}
bool CVX_Voxel::isFloorStaticFriction() const
{
    if (!sim) return true;
    sim->fetchVoxel(index);
    return (sim->mFlags[index] & VX_VF_STATIC_FRICTION) != 0;
}
void CVX_Voxel::setTemperature(float t)
{
    if (!sim) { temp0 = t; return; }
    sim->sync();
    VXH(sim, upload, VX_F_TEMP, index, 1, &t);
    sim->epoch++;
}
void CVX_Voxel::haltMotion()
{
    if (!sim) return;
    sim->sync();
    const double zero[3] = {0, 0, 0};
    VXH(sim, upload, VX_F_LINMOM, index, 1, zero);
    VXH(sim, upload, VX_F_ANGMOM, index, 1, zero);
    sim->epoch++;
}
// ---- derived quantities, computed on the host from the mirrored state exactly as the reference does
Vec3D<float> CVX_Voxel::cornerOffset(voxelCorner corner) const
{
    Vec3D<> strains;
    for (int i = 0; i < 3; i++) {
        bool posLink = (corner & (1 << (2 - i))) != 0;
        CVX_Link* pL = links[2 * i + (posLink ? 0 : 1)];
        if (pL && !pL->isFailed()) strains[i] = (1 + pL->axialStrain(posLink)) * (posLink ? 1 : -1);
        else strains[i] = posLink ? 1.0 : -1.0;
    }
    return (Vec3D<float>)((0.5 * baseSize()).Scale(strains));
}
Vec3D<float> CVX_Voxel::cornerPosition(voxelCorner corner) const
{
    return (Vec3D<float>)position() + (Vec3D<float>)orientation().RotateVec3D((Vec3D<double>)cornerOffset(corner));
}
Vec3D<float> CVX_Voxel::strain(bool poissonsStrain) const
{
    Vec3D<float> ret(0, 0, 0);
    int n[3] = {0, 0, 0};
    bool tension[3] = {false, false, false};
    for (int i = 0; i < 6; i++) {
        if (!links[i]) continue;
        int axis = toAxis((linkDirection)i);
        ret[axis] += links[i]->axialStrain(isNegative((linkDirection)i));
        n[axis]++;
    }
    for (int i = 0; i < 3; i++) {
        if (n[i] == 2) ret[i] *= 0.5f;
        if (poissonsStrain) tension[i] = (n[i] == 2) || (ext && (n[i] == 1 && (ext->isFixed((dofComponent)(1 << i)) || ext->force()[i] != 0)));
    }
    if (poissonsStrain && !(tension[0] && tension[1] && tension[2])) {
        float add = 0;
        for (int i = 0; i < 3; i++) if (tension[i]) add += ret[i];
        float value = (float)pow(1.0f + add, -mat->poissonsRatio()) - 1.0f;
        for (int i = 0; i < 3; i++) if (!tension[i]) ret[i] = value;
    }
    return ret;
}
float CVX_Voxel::transverseStrainSum(CVX_Link::linkAxis axis)
{
    if (mat->poissonsRatio() == 0) return 0;
    Vec3D<float> ps = strain(true);
    switch (axis) {
    case CVX_Link::X_AXIS: return ps.y + ps.z;
    case CVX_Link::Y_AXIS: return ps.x + ps.z;
    case CVX_Link::Z_AXIS: return ps.x + ps.y;
    default: return 0.0f;
    }
}
float CVX_Voxel::transverseArea(CVX_Link::linkAxis axis)
{
    float size = (float)mat->nominalSize();
    if (mat->poissonsRatio() == 0) return size * size;
    Vec3D<> ps = (Vec3D<>)strain(true);
    switch (axis) {
    case CVX_Link::X_AXIS: return (float)(size * size * (1 + ps.y) * (1 + ps.z));
    case CVX_Link::Y_AXIS: return (float)(size * size * (1 + ps.x) * (1 + ps.z));
    case CVX_Link::Z_AXIS: return (float)(size * size * (1 + ps.x) * (1 + ps.y));
    default: return size * size;
    }
}
bool CVX_Voxel::isFloorEnabled() const { return floorOverride >= 0 ? floorOverride != 0 : (sim ? sim->isFloorEnabled() : false); }
void CVX_Voxel::enableFloor(bool enabled)
{
    floorOverride = enabled ? 1 : 0;
    if (sim) sim->floorEdits.push_back(this);
}
float CVX_Voxel::dampingMultiplier()
{
    vxm::MassProps p = vxm::mass_props(mat->m_, mat->nominalSize());
    return 2 * p.sqrt_mass * mat->m_.zeta_int / (sim ? sim->previousDt : 0.0f);
}
Vec3D<double> CVX_Voxel::force()
{
    Vec3D<double> total(0, 0, 0);
    for (int i = 0; i < 6; i++) if (links[i]) total += links[i]->force(isNegative((linkDirection)i));     // LCS
    total = orientation().RotateVec3D(total);
    if (externalExists()) total += (Vec3D<double>)external()->force();
    total -= velocity() * (double)mat->globalDampingTranslateC();
    total.z += -mat->mass() * 9.80665f * mat->gravMult_;                                                       // gravityForce(), VX_MaterialVoxel.h:44
    if (sim && sim->isCollisionsEnabled()) {
        for (CVX_Collision* c : *sim->collisionList()) total -= (Vec3D<double>)c->contactForce(this);
    }
    return total;
}
Vec3D<double> CVX_Voxel::moment()
{
    Vec3D<double> total(0, 0, 0);
    for (int i = 0; i < 6; i++) if (links[i]) total += links[i]->moment(isNegative((linkDirection)i));
    total = orientation().RotateVec3D(total);
    if (externalExists()) total += (Vec3D<double>)external()->moment();
    total -= angularVelocity() * (double)mat->globalDampingRotateC();
    return total;
}
Vec3D<float> CVX_Voxel::externalForce()
{
    Vec3D<float> ret(external()->force());
    if (ext->isFixed(X_TRANSLATE) || ext->isFixed(Y_TRANSLATE) || ext->isFixed(Z_TRANSLATE)) {
        Vec3D<float> reaction = (Vec3D<float>)(-force());
        if (ext->isFixed(X_TRANSLATE)) ret.x = reaction.x;
        if (ext->isFixed(Y_TRANSLATE)) ret.y = reaction.y;
        if (ext->isFixed(Z_TRANSLATE)) ret.z = reaction.z;
    }
    return ret;
}
Vec3D<float> CVX_Voxel::externalMoment()
{
    Vec3D<float> ret(external()->moment());
    if (ext->isFixed(X_ROTATE) || ext->isFixed(Y_ROTATE) || ext->isFixed(Z_ROTATE)) {
        Vec3D<float> reaction = (Vec3D<float>)(-moment());
        if (ext->isFixed(X_ROTATE)) ret.x = reaction.x;
        if (ext->isFixed(Y_ROTATE)) ret.y = reaction.y;
        if (ext->isFixed(Z_ROTATE)) ret.z = reaction.z;
    }
    return ret;
}
bool CVX_Voxel::isYielded() const { for (int i = 0; i < 6; i++) if (links[i] && links[i]->isYielded()) return true; return false; }
bool CVX_Voxel::isFailed() const { for (int i = 0; i < 6; i++) if (links[i] && links[i]->isFailed()) return true; return false; }

// ================================================================================================ CVX_Link
struct LinkRead {         // single link fields, from the device that owns the link
    static Vec3D<> vec(const CVoxelyze* sim, int field, int index) { double v[3] = {0, 0, 0}; VXH(sim, download, field, index, 1, v); return Vec3D<>(v[0], v[1], v[2]); }
    static float f32(const CVoxelyze* sim, int field, int index) { float v = 0; VXH(sim, download, field, index, 1, &v); return v; }
    static uint32_t flags(const CVoxelyze* sim, int index) { uint32_t v = 0; VXH(sim, download, VX_F_LINKFLAGS, index, 1, &v); return v; }
};

Vec3D<> CVX_Link::force(bool positiveEnd) const { sim->sync(); return LinkRead::vec(sim, positiveEnd ? VX_F_FORCE_POS : VX_F_FORCE_NEG, index); }
Vec3D<> CVX_Link::moment(bool positiveEnd) const { sim->sync(); return LinkRead::vec(sim, positiveEnd ? VX_F_MOMENT_POS : VX_F_MOMENT_NEG, index); }
float CVX_Link::axialStrain() const { sim->sync(); return LinkRead::f32(sim, VX_F_STRAIN, index); }
float CVX_Link::axialStrain(bool positiveEnd) const
{
    float strain = axialStrain();
    float ratio = pVPos->material()->youngsModulus() / pVNeg->material()->youngsModulus();      // strainRatio, src/VX_Link.cpp:67
    return positiveEnd ? 2.0f * strain * ratio / (1.0f + ratio) : 2.0f * strain / (1.0f + ratio);
}
float CVX_Link::axialStress() const { sim->sync(); return LinkRead::f32(sim, VX_F_STRESS, index); }
bool CVX_Link::isSmallAngle() const { sim->sync(); return (LinkRead::flags(sim, index) & VX_LF_SMALL_ANGLE) != 0; }
bool CVX_Link::isYielded() const { sim->sync(); return (LinkRead::flags(sim, index) & VX_LF_YIELDED) != 0; }
bool CVX_Link::isFailed() const { sim->sync(); return (LinkRead::flags(sim, index) & VX_LF_FAILED) != 0; }
float CVX_Link::strainEnergy() const                                                              // src/VX_Link.cpp:251-257
{
    Vec3D<> fN = force(false), mN = moment(false), mP = moment(true);
    return fN.x * fN.x / (2.0f * mat->a1()) + mN.x * mN.x / (2.0 * mat->a2()) +
           (mN.z * mN.z - mN.z * mP.z + mP.z * mP.z) / (3.0 * mat->b3()) +
           (mN.y * mN.y - mN.y * mP.y + mP.y * mP.y) / (3.0 * mat->b3());
}
float CVX_Link::axialStiffness() { return mat->a1(); }      // nu = 0 value (src/VX_Link.cpp:260)

// ================================================================================================ CVoxelyze
CVoxelyze::CVoxelyze(double voxelSize) : voxSize(voxelSize)
{
    if (const char* e = getenv("VX_DEVICES")) {               // "0,1,2,3": an unmodified caller of the class API runs slabbed over these
        std::vector<int> d;
        for (const char* q = e; *q;) { char* end = nullptr; long v = strtol(q, &end, 10); if (end == q) break; d.push_back((int)v); q = *end == ',' ? end + 1 : end; }
        if (!d.empty()) { devices = d; device = d[0]; devicesFromEnv = d.size() > 1; }
    }
}
CVoxelyze::~CVoxelyze() { clear(); destroyHandle(); }

CVoxelyze& CVoxelyze::operator=(CVoxelyze& VIn)             // src/Voxelyze.cpp:39-58
{
    setVoxelSize(VIn.voxSize);
    setAmbientTemperature(VIn.ambientTemperature(), true);
    setGravity(VIn.gravity());
    enableFloor(VIn.isFloorEnabled());
    enableCollisions(VIn.isCollisionsEnabled());
    std::unordered_map<CVX_Material*, CVX_Material*> matMap;
    for (int i = 0; i < VIn.materialCount(); i++) matMap[VIn.material(i)] = addMaterial(*(VIn.material(i)));
    for (int i = 0; i < VIn.voxelCount(); i++) {
        CVX_Voxel* pVIn = VIn.voxel(i);
        CVX_Voxel* pVOut = setVoxel(matMap[pVIn->material()], pVIn->indexX(), pVIn->indexY(), pVIn->indexZ());
        *pVOut->external() = *pVIn->external();
    }
    return *this;
}

void CVoxelyze::setVoxelSize(double voxelSize)
{
    const double scale = voxelSize / voxSize;
    const bool had = stepped && (h || hm) && !voxelsList.empty();
    if (had) { fetchAll(); clockTime = VXH0(this, time); clockPending = true; }
    voxSize = voxelSize;
    for (CVX_MaterialVoxel* m : voxelMats) m->setNominalSize(voxelSize);
    for (CVX_MaterialLink* m : linkMats) delete m;            // combined materials depend on the size: rebuilt on demand
    linkMats.clear();
    destroyHandle();                                          // the voxel size is a property of the device handle
    topologyDirty = envDirty = true; matChangesSeen = ~0ull; extChangesSeen = ~0ull;
    linkStateFetched = false; linkStateMirror.clear(); editedVoxels.clear();        // links restart (CVX_Link::reset)
    if (had) {
        for (double& p : mPos) p *= scale;
        std::fill(mLin.begin(), mLin.end(), 0.0); std::fill(mAng.begin(), mAng.end(), 0.0);     // haltMotion()
        for (uint32_t& fl : mFlags) fl &= ~VX_VF_STATIC_FRICTION;
    }
    epoch++;
}

void CVoxelyze::die(const char* what) const
{
    fprintf(stderr, "voxelyze_b200: %s: %s\n", what, hm ? vx_slabbed_last_error(hm) : h ? vx_last_error(h) : "no CUDA device handle (this build has no CPU fallback)");
    abort();
}

void CVoxelyze::clear()
{
    for (auto& kv : linkPool) delete kv.second;
    linkPool.clear(); linksList.clear();
    for (CVX_Voxel* v : voxelsList) delete v;
    voxelsList.clear(); cells.clear();
    for (CVX_MaterialVoxel* m : voxelMats) delete m;
    voxelMats.clear();
    for (CVX_MaterialLink* m : linkMats) delete m;
    linkMats.clear();
    for (CVX_Collision* c : collisionsList) delete c;
    collisionsList.clear();
    ambientTemp = 0.0f; grav = 0.0f; floor = false; collisions = false;
    topologyDirty = envDirty = true; stepped = false; clockPending = false; epoch++;
    destroyHandle();
}

CVX_Material* CVoxelyze::addMaterial(float E, float rho)
{
    CVX_MaterialVoxel* m = new CVX_MaterialVoxel(E, rho, voxSize);
    m->gravMult_ = grav;
    voxelMats.push_back(m);
    topologyDirty = true;           // the material table grew: the device model is rebuilt at the next sync
    return m;
}
CVX_Material* CVoxelyze::addMaterial(const CVX_Material& mat)
{
    CVX_MaterialVoxel* m = new CVX_MaterialVoxel(mat, voxSize);
    m->gravMult_ = grav;
    voxelMats.push_back(m);
    topologyDirty = true;
    return m;
}
bool CVoxelyze::removeMaterial(CVX_Material* toRemove)
{
    auto it = std::find(voxelMats.begin(), voxelMats.end(), (CVX_MaterialVoxel*)toRemove);
    if (it == voxelMats.end()) return false;
    std::vector<CVX_Voxel*> doomed;
    for (CVX_Voxel* v : voxelsList) if (v->mat == *it) doomed.push_back(v);
    for (CVX_Voxel* v : doomed) removeVoxel(v->ix, v->iy, v->iz);
    delete *it;
    voxelMats.erase(it);
    topologyDirty = true;
    return true;
}
bool CVoxelyze::replaceMaterial(CVX_Material* replaceMe, CVX_Material* replaceWith)
{
    auto has = [&](CVX_Material* m) { return std::find(voxelMats.begin(), voxelMats.end(), (CVX_MaterialVoxel*)m) != voxelMats.end(); };
    if (!has(replaceMe) || !has(replaceWith)) return false;
    for (CVX_Voxel* v : std::vector<CVX_Voxel*>(voxelsList)) if (v->mat == (CVX_MaterialVoxel*)replaceMe) setVoxel(replaceWith, v->ix, v->iy, v->iz);
    return true;
}

CVX_Voxel* CVoxelyze::voxel(int x, int y, int z) const
{
    auto it = cells.find(key(x, y, z));
    return it == cells.end() ? nullptr : it->second;
}

CVX_Voxel* CVoxelyze::setVoxel(CVX_Material* material, int x, int y, int z)
{
    if (material == nullptr) { removeVoxel(x, y, z); return nullptr; }
    CVX_MaterialVoxel* m = (CVX_MaterialVoxel*)material;
    CVX_Voxel* v = voxel(x, y, z);
    if (v) {                                        // replaceVoxel (src/Voxelyze.cpp:485-498)
        if (v->mat != m) {
            if (stepped) {                          // keep velocity across the material change (src/VX_Voxel.cpp:78-90)
                fetchAll(); fetchLinkState();
                editedVoxels.push_back(v);
                double ls = m->p_.mass / v->mat->p_.mass, as = m->p_.inertia / v->mat->p_.inertia;
                for (int a = 0; a < 3; a++) { mLin[3 * v->index + a] *= ls; mAng[3 * v->index + a] *= as; }
                mFlags[v->index] &= ~VX_VF_STATIC_FRICTION;
                pendingStateEdit.push_back(v->index);
            }
            v->mat = m;
            topologyDirty = true;
        }
        return v;
    }
    if (stepped) { fetchAll(); fetchLinkState(); }   // existing voxels and links keep their state across the re-layout
    v = new CVX_Voxel(m, (short)x, (short)y, (short)z);
    v->sim = this; v->index = (int)voxelsList.size();
    voxelsList.push_back(v);
    cells[key(x, y, z)] = v;
    topologyDirty = true;
    return v;
}

void CVoxelyze::removeVoxel(int x, int y, int z)
{
    CVX_Voxel* v = voxel(x, y, z);
    if (!v) return;
    if (stepped) { fetchAll(); fetchLinkState(); }
    cells.erase(key(x, y, z));
    removedIndices.push_back(v->index);
    voxelsList.erase(voxelsList.begin() + v->index);
    for (size_t i = 0; i < voxelsList.size(); i++) voxelsList[i]->index = (int)i;
    for (auto it = linkPool.begin(); it != linkPool.end();) {
        if (it->second->pVNeg == v || it->second->pVPos == v) { delete it->second; it = linkPool.erase(it); } else ++it;
    }
    delete v;
    topologyDirty = true;
}

int CVoxelyze::bound(int axis, bool max) const
{
    if (voxelsList.empty()) return 0;
    int best = max ? -32768 : 32767;
    for (CVX_Voxel* v : voxelsList) {
        int c = axis == 0 ? v->ix : (axis == 1 ? v->iy : v->iz);
        best = max ? std::max(best, c) : std::min(best, c);
    }
    return best;
}

CVX_MaterialLink* CVoxelyze::combinedMaterial(CVX_MaterialVoxel* a, CVX_MaterialVoxel* b) const
{
    for (CVX_MaterialLink* m : linkMats)
        if ((m->vox1Mat == a && m->vox2Mat == b) || (m->vox1Mat == b && m->vox2Mat == a)) return m;
    CVX_MaterialLink* m = new CVX_MaterialLink(a, b);
    linkMats.push_back(m);
    return m;
}

// ---- devices ---------------------------------------------------------------------------------------
void CVoxelyze::destroyHandle() const
{
    if (h) { vx_destroy(h); h = nullptr; }
    if (hm) { vx_slabbed_destroy(hm); hm = nullptr; }
}
void CVoxelyze::notSlabbed(const char* what) const
{
    if (devicesFromEnv && !syncing) {                 // the device list came from VX_DEVICES, not from the caller: move the model to one device and carry on
        const_cast<CVoxelyze*>(this)->setDevices(std::vector<int>(1, devices[0]));
        sync();
        return;
    }
    fprintf(stderr, "voxelyze_b200: %s is not available while the model is cut into slabs over %d devices; call setDevices() with one device first\n", what, (int)devices.size());
    abort();
}
// the model moves to another set of devices with its dynamic state: voxel mirrors and link records are fetched from the old
// handle(s) and uploaded by the rebuild, like after a topology edit
void CVoxelyze::setDevices(const std::vector<int>& cudaDevices)
{
    if (cudaDevices.empty() || cudaDevices == devices) return;
    const bool had = stepped && (h || hm) && !voxelsList.empty();
    if (had) { sync(); fetchAll(); fetchLinkState(); clockTime = VXH0(this, time); clockPending = true; }      // (sync: pending edits are applied where the state is)
    destroyHandle();
    devices = cudaDevices; device = devices[0]; devicesFromEnv = false;
    topologyDirty = envDirty = true; matChangesSeen = ~0ull; extChangesSeen = ~0ull; envelopeSeen = 0.0f;
    if (ambientTemp != 0.0f && !had) tempAllDirty = true;
    epoch++;
}

// ---- device synchronisation -------------------------------------------------------------------
void CVoxelyze::uploadMaterials() const
{
    std::vector<vx_material_desc> d(voxelMats.size());
    for (size_t i = 0; i < d.size(); i++) {
        const vxm::Material& m = voxelMats[i]->model();
        vx_material_desc& o = d[i];
        memset(&o, 0, sizeof(o));
        // the model is handed over as data points: that is what every setModel* call produces, so all
        // three model kinds round-trip exactly (linear models keep their special flag below)
        if (m.linear) { o.model = VX_MODEL_LINEAR; o.youngs_modulus = m.E; o.fail_stress = m.sigma_fail; }
        else { o.model = VX_MODEL_DATA; o.n_points = (int)m.eps.size() - 1; o.strain = m.eps.data() + 1; o.stress = m.sig.data() + 1; }
        o.density = m.rho; o.poissons_ratio = m.nu; o.cte = m.cte; o.mu_static = m.mu_s; o.mu_kinetic = m.mu_k;
        o.zeta_internal = m.zeta_int; o.zeta_global = m.zeta_glob; o.zeta_collision = m.zeta_coll;
        for (int a = 0; a < 3; a++) o.ext_scale[a] = m.ext_scale[a];
    }
    if (VXH(this, set_materials, (int)d.size(), d.data()) != VX_OK) die("vx_set_materials");
    for (CVX_MaterialLink* lm : linkMats) lm->updateAll();
}

void CVoxelyze::uploadExternals() const
{
    std::vector<int32_t> vox; std::vector<uint8_t> dof; std::vector<float> f, m; std::vector<double> t, r;
    for (CVX_Voxel* v : voxelsList) {
        if (!v->ext) continue;
        CVX_External* e = v->ext;
        vox.push_back(v->index); dof.push_back(e->dofMask());
        Vec3D<float> ef = e->force(), em = e->moment(); Vec3D<double> et = e->translation(), er = e->rotation();
        f.insert(f.end(), {ef.x, ef.y, ef.z}); m.insert(m.end(), {em.x, em.y, em.z});
        t.insert(t.end(), {et.x, et.y, et.z}); r.insert(r.end(), {er.x, er.y, er.z});
    }
    if (VXH(this, set_externals, (int)vox.size(), vox.data(), dof.data(), f.data(), m.data(), t.data(), r.data()) != VX_OK) die("vx_set_externals");
}

// state of every link of the model the device currently holds, fetched once before the first edit of a batch
void CVoxelyze::fetchLinkState() const
{
    if (!stepped || linkStateFetched || (!h && !hm) || topologyDirty) return;
    const int L = VXH0(this, link_count);
    linkStateMirror.resize((size_t)L * sizeof(vx_link_state));
    if (L && VXH(this, download_link_state, 0, L, (vx_link_state*)linkStateMirror.data()) != VX_OK) die("vx_download_link_state");
    linkStateFetched = true;
}

void CVoxelyze::rebuildTopology() const
{
    const int n = (int)voxelsList.size();
    std::vector<int32_t> ijk(3 * (size_t)n); std::vector<uint16_t> mat(n);
    for (int i = 0; i < n; i++) {
        CVX_Voxel* v = voxelsList[i];
        ijk[3 * i] = v->ix; ijk[3 * i + 1] = v->iy; ijk[3 * i + 2] = v->iz;
        mat[i] = (uint16_t)(std::find(voxelMats.begin(), voxelMats.end(), v->mat) - voxelMats.begin());
    }
    const bool keepState = stepped;                 // mirrors were made current by setVoxel/removeVoxel before the edit
    if (hm) {
        if (collisions) notSlabbed("enableCollisions");
        const int rc = vx_slabbed_set_voxels(hm, n, ijk.data(), mat.data());
        if (rc == VX_ERR_UNSUPPORTED && devicesFromEnv) {             // a model that cannot be cut (too sparse for its bounding box): one device after all
            destroyHandle();
            devices.resize(1);
            if (vx_create(voxSize, devices[0], &h) != VX_OK) die("vx_create");
            if (vx_set_gravity(h, grav) != VX_OK || vx_enable_floor(h, floor) != VX_OK) die("environment");
            uploadMaterials();
        } else if (rc != VX_OK) die("vx_slabbed_set_voxels");
    }
    if (h) {
        if (vx_enable_collisions(h, 0) != VX_OK) die("vx_enable_collisions");
        if (collisions && vx_enable_collisions(h, 1) != VX_OK) die("vx_enable_collisions");
        if (vx_set_voxels(h, n, ijk.data(), mat.data(), nullptr, nullptr) != VX_OK) die("vx_set_voxels");
    }

    // link handles in C-ABI link order; surviving links keep their handle
    const int L = VXH0(this, link_count);
    std::vector<int32_t> vn(L), vp(L); std::vector<uint8_t> ax(L);
    VXH(this, get_links, vn.data(), vp.data(), ax.data());
    std::map<std::pair<CVX_Voxel*, int>, CVX_Link*> pool;
    // links that survive the edit keep their state; the links of a voxel whose material was swapped restart, like
    // every new link (the reference destroys and recreates exactly those, src/Voxelyze.cpp:485-498)
    std::vector<vx_link_state> carried;
    const vx_link_state* oldState = (const vx_link_state*)linkStateMirror.data();
    const int oldL = (int)(linkStateMirror.size() / sizeof(vx_link_state));
    if (keepState && linkStateFetched) {
        vx_link_state fresh; memset(&fresh, 0, sizeof(fresh)); fresh.flags = VX_LF_SMALL_ANGLE;     // CVX_Link::reset, src/VX_Link.cpp:61-75
        carried.assign(L, fresh);
    }
    linksList.assign(L, nullptr);
    for (CVX_Voxel* v : voxelsList) for (int d = 0; d < 6; d++) v->links[d] = nullptr;
    for (int i = 0; i < L; i++) {
        CVX_Voxel* a = voxelsList[vn[i]]; CVX_Voxel* b = voxelsList[vp[i]];
        std::pair<CVX_Voxel*, int> k(a, ax[i]);
        CVX_Link* l;
        auto it = linkPool.find(k);
        if (it != linkPool.end() && it->second->pVPos == b) {
            l = it->second; linkPool.erase(it);
            const bool restarted = std::find(editedVoxels.begin(), editedVoxels.end(), a) != editedVoxels.end() ||
                                   std::find(editedVoxels.begin(), editedVoxels.end(), b) != editedVoxels.end();
            if (!carried.empty() && !restarted && l->index >= 0 && l->index < oldL) carried[i] = oldState[l->index];
        } else l = new CVX_Link(const_cast<CVoxelyze*>(this), a, b, (CVX_Link::linkAxis)ax[i]);
        l->index = i; l->mat = combinedMaterial(a->mat, b->mat);
        pool[k] = l; linksList[i] = l;
        a->links[2 * ax[i]] = l; b->links[2 * ax[i] + 1] = l;
    }
    for (auto& kv : linkPool) delete kv.second;      // links that no longer exist
    linkPool.swap(pool);

    if (keepState && n) {                            // voxel state survives a topology edit
        // mirrors are indexed by the OLD voxel order minus removed entries: compact them first
        std::sort(removedIndices.begin(), removedIndices.end());
        for (int k = (int)removedIndices.size() - 1; k >= 0; k--) {
            int idx = removedIndices[k];
            if (idx < (int)mTemp.size()) {
                mPos.erase(mPos.begin() + 3 * idx, mPos.begin() + 3 * idx + 3); mOrient.erase(mOrient.begin() + 4 * idx, mOrient.begin() + 4 * idx + 4);
                mLin.erase(mLin.begin() + 3 * idx, mLin.begin() + 3 * idx + 3); mAng.erase(mAng.begin() + 3 * idx, mAng.begin() + 3 * idx + 3);
                mTemp.erase(mTemp.begin() + idx); mFlags.erase(mFlags.begin() + idx);
            }
        }
        int old = (int)mTemp.size();                 // voxels [0, old) existed before; the rest are new and start fresh
        if (old > n) old = n;
        if (old) {
            VXH(this, upload, VX_F_POS, 0, old, mPos.data()); VXH(this, upload, VX_F_ORIENT, 0, old, mOrient.data());
            VXH(this, upload, VX_F_LINMOM, 0, old, mLin.data()); VXH(this, upload, VX_F_ANGMOM, 0, old, mAng.data());
            VXH(this, upload, VX_F_TEMP, 0, old, mTemp.data()); VXH(this, upload, VX_F_VOXFLAGS, 0, old, mFlags.data());
        }
    }
    if (!carried.empty() && VXH(this, upload_link_state, 0, L, carried.data()) != VX_OK) die("vx_upload_link_state");
    removedIndices.clear(); pendingStateEdit.clear(); editedVoxels.clear();
    linkStateMirror.clear(); linkStateFetched = false;
    mirrorEpoch.assign(n, 0);
    mPos.resize(3 * (size_t)n); mOrient.resize(4 * (size_t)n); mLin.resize(3 * (size_t)n); mAng.resize(3 * (size_t)n);
    mTemp.resize(n); mFlags.resize(n);
    extChangesSeen = ~0ull;
    epoch++;
    if (!keepState) for (CVX_Voxel* v : voxelsList) if (v->floorOverride >= 0) floorEdits.push_back(v);     // a fresh device model knows no per-voxel flags
}

void CVoxelyze::sync() const
{
    struct Busy { bool& b; bool was; Busy(bool& x) : b(x), was(x) { b = true; } ~Busy() { b = was; } } busy(syncing);
    if (!h && !hm) {
        if (devices.size() > 1 && devicesFromEnv) {                                          // VX_DEVICES is a wish, setDevices an order
            const bool cut = !collisions && !voxelsList.empty() && bound(2, true) - bound(2, false) + 1 >= 4;      // at least two planes per slab
            if (!cut) devices.resize(1);
        }
        if (devices.size() > 1) { if (vx_slabbed_create(voxSize, (int)devices.size(), devices.data(), &hm) != VX_OK) die("vx_slabbed_create"); }
        else if (vx_create(voxSize, devices.empty() ? device : devices[0], &h) != VX_OK) die("vx_create");
        topologyDirty = envDirty = true; matChangesSeen = ~0ull;
    }
    uint64_t matChanges = voxelMats.size();
    for (CVX_MaterialVoxel* m : voxelMats) matChanges += m->changeCount() * 1315423911ull;
    if (h && CVX_Collision::envelopeRadius != envelopeSeen) { vx_set_collision_envelope(h, CVX_Collision::envelopeRadius); envelopeSeen = CVX_Collision::envelopeRadius; }
    if (envDirty) {
        if (VXH(this, set_gravity, grav) != VX_OK || VXH(this, enable_floor, floor) != VX_OK) die("environment");
    }
    if (matChanges != matChangesSeen || topologyDirty) { uploadMaterials(); matChangesSeen = matChanges; epoch++; }
    if (topologyDirty) { rebuildTopology(); topologyDirty = false; }
    if (clockPending) {                               // the model came from another handle mid-run: time and CVX_Voxel::previousDt go on
        if (stepped && VXH(this, set_clock, clockTime, previousDt) != VX_OK) die("vx_set_clock");
        clockPending = false;
    }
    if (envDirty) {
        if (hm && collisions) notSlabbed("enableCollisions");
        if (h && vx_enable_collisions(h, collisions) != VX_OK) die("vx_enable_collisions");
        envDirty = false;
    }
    if (extChanges != extChangesSeen) { uploadExternals(); extChangesSeen = extChanges; epoch++; }
    if (tempAllDirty) { VXH(this, set_temperature_all, ambientTemp); tempAllDirty = false; epoch++; }
    if (!floorEdits.empty()) applyFloorEdits();
}

// per-voxel CVX_Voxel::enableFloor: the voxel's VX_VF_FLOOR_OFF / VX_VF_FLOOR_ON bits, read-modify-write (its friction bit stays)
void CVoxelyze::applyFloorEdits() const
{
    std::vector<CVX_Voxel*> edits;
    edits.swap(floorEdits);
    for (CVX_Voxel* v : edits) {
        if (v->index < 0 || v->index >= (int)voxelsList.size() || voxelsList[v->index] != v) continue;
        uint32_t fl = 0;
        if (VXH(this, download, VX_F_VOXFLAGS, v->index, 1, &fl) != VX_OK) die("vx_download");
        fl &= ~(VX_VF_FLOOR_OFF | VX_VF_FLOOR_ON);
        if (v->floorOverride == 0 && floor) fl |= VX_VF_FLOOR_OFF;
        if (v->floorOverride == 1 && !floor) fl |= VX_VF_FLOOR_ON;
        if (VXH(this, upload, VX_F_VOXFLAGS, v->index, 1, &fl) != VX_OK) die("vx_upload");
    }
    epoch++;
}

void CVoxelyze::fetchAll() const
{
    if ((!h && !hm) || voxelsList.empty() || topologyDirty) return;
    const int n = (int)mTemp.size();
    if (n == 0) return;
    VXH(this, download, VX_F_POS, 0, n, mPos.data()); VXH(this, download, VX_F_ORIENT, 0, n, mOrient.data());
    VXH(this, download, VX_F_LINMOM, 0, n, mLin.data()); VXH(this, download, VX_F_ANGMOM, 0, n, mAng.data());
    VXH(this, download, VX_F_TEMP, 0, n, mTemp.data()); VXH(this, download, VX_F_VOXFLAGS, 0, n, mFlags.data());
    std::fill(mirrorEpoch.begin(), mirrorEpoch.end(), epoch);
    singleFetches = 0;
}

void CVoxelyze::fetchVoxel(int i) const
{
    sync();
    if (mirrorEpoch[i] == epoch) return;
    // a caller polling a handful of voxels per step pays a few 100-byte copies; a caller walking the
    // whole list gets one bulk download
    if (++singleFetches > 32) { fetchAll(); return; }
    vx_voxel_state r;                                   // one call, one tiny kernel writing into mapped pinned memory
    if (VXH(this, download_voxel_state, i, 1, &r) != VX_OK) die("vx_download_voxel_state");
    for (int k = 0; k < 3; k++) { mPos[3 * i + k] = r.pos[k]; mLin[3 * i + k] = r.linmom[k]; mAng[3 * i + k] = r.angmom[k]; }
    for (int k = 0; k < 4; k++) mOrient[4 * i + k] = r.orient[k];
    mTemp[i] = r.temp; mFlags[i] = r.flags;
    mirrorEpoch[i] = epoch;
}

// ---- the hot path -----------------------------------------------------------------------------
bool CVoxelyze::doTimeStep(float dt)
{
    if (dt == 0) return true;
    sync();
    if (voxelsList.empty()) return true;
    const float timeBefore = VXH0(this, time);
    int rc = VXH(this, step, dt, 1, nullptr);
    if (rc != VX_OK && rc != VX_DIVERGED) die("vx_step");
    stepped = true; epoch++; singleFetches = 0;
    if (rc == VX_OK) previousDt = dt > 0 ? dt : VXH0(this, time) - timeBefore;
    return rc == VX_OK;
}

// the static solve writes new poses and zero momenta on the device: every host mirror is stale afterwards, like after a step
bool CVoxelyze::staticSolve(double relTol, int maxIter, int* iterations, double* residual, std::string* error)
{
    sync();
    if (hm) notSlabbed("doLinearSolve");
    int rc = vx_linear_solve(h, relTol, maxIter, iterations, residual);
    if (rc == VX_OK) { stepped = true; epoch++; singleFetches = 0; return true; }
    if (error) *error = vx_last_error(h);
    if (rc != VX_ERR_SOLVER && rc != VX_ERR_ARG) die("vx_linear_solve");
    return false;
}
bool CVoxelyze::doLinearSolve() { staticSolve(0.0, 0, nullptr, nullptr, nullptr); return true; }

float CVoxelyze::recommendedTimeStep() const
{
    sync();
    float dt = 0.0f;
    if (VXH(this, recommended_dt, &dt) != VX_OK) die("vx_recommended_dt");
    return dt;
}

void CVoxelyze::resetTime()
{
    sync();
    if (VXH0(this, reset) != VX_OK) die("vx_reset");
    stepped = false; epoch++; previousDt = 0.0f;
    for (CVX_Voxel* v : voxelsList) if (v->floorOverride >= 0) floorEdits.push_back(v);      // vx_reset keeps no per-voxel flag
}

void CVoxelyze::setAmbientTemperature(float t, bool allVoxels) { ambientTemp = t; if (allVoxels) { tempAllDirty = true; } }
void CVoxelyze::setGravity(float g) { grav = g; for (CVX_MaterialVoxel* m : voxelMats) m->gravMult_ = g; envDirty = true; }
void CVoxelyze::enableFloor(bool e)         // src/Voxelyze.cpp:604-610: every voxel follows
{
    floor = e; envDirty = true;
    for (CVX_Voxel* v : voxelsList) if (v->floorOverride >= 0) { v->floorOverride = -1; floorEdits.push_back(v); }
}
void CVoxelyze::enableCollisions(bool e) { if (collisions == e) return; if (e && hm && devicesFromEnv) notSlabbed("enableCollisions"); collisions = e; envDirty = true; if (!stepped) topologyDirty = true; }

// ---- links / collisions -------------------------------------------------------------------------
int CVoxelyze::linkCount() const { sync(); return (int)linksList.size(); }
CVX_Link* CVoxelyze::link(int i) { sync(); return linksList[i]; }
const std::vector<CVX_Link*>* CVoxelyze::linkList() const { sync(); return &linksList; }
CVX_Link* CVoxelyze::link(int x, int y, int z, CVX_Voxel::linkDirection d) const
{
    sync();
    CVX_Voxel* v = voxel(x, y, z);
    return v ? v->links[d] : nullptr;
}
const std::vector<CVX_Collision*>* CVoxelyze::collisionList() const
{
    sync();
    for (CVX_Collision* c : collisionsList) delete c;
    collisionsList.clear();
    int n = 0;
    if (hm) return &collisionsList;                   // no self-collisions on a slabbed run
    vx_collision_forces(h, nullptr, nullptr, 0, &n);
    std::vector<int32_t> p(2 * (size_t)n); std::vector<float> fr(3 * (size_t)n);
    if (n) vx_collision_forces(h, p.data(), fr.data(), n, &n);
    for (int k = 0; k < n; k++) {
        CVX_Collision* c = new CVX_Collision(voxelsList[p[2 * k]], voxelsList[p[2 * k + 1]]);
        c->f_ = Vec3D<float>(fr[3 * k], fr[3 * k + 1], fr[3 * k + 2]);
        collisionsList.push_back(c);
    }
    return &collisionsList;
}

// CVoxelyze::stateInfo (src/Voxelyze.cpp:752-800): min / max / sum reductions run on the device
float CVoxelyze::stateInfo(stateInfoType info, valueType type)
{
    sync();
    float v = 0.0f;
    int rc = VXH(this, state_info, (int)info, (int)type, &v);
    if (rc != VX_OK && rc != VX_ERR_UNSUPPORTED) die("vx_state_info");
    return v;
}
