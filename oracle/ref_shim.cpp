// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Wraps the UNMODIFIED reference (jonhiller/Voxelyze, compiled from its own sources
// where they lie under /root/reference by oracle/Makefile) behind the C-ABI of
// include/voxelyze_b200.h, so that tests can drive "the reference itself" and the
// CUDA product with identical calls.  Output: oracle/_ref/libvxref.so (+ _omp).
// Nothing in the product path may link, load or call this file.
//
// The only liberty taken: the reference's headers are included with `private` and
// `protected` opened up so that link/voxel internals (pos2, maxStrain, ...) can be
// read for parity checks.  No reference source is copied or modified.

#include <vector>
#include <list>
#include <string>
#include <map>
#include <unordered_map>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <sstream>
#include <iostream>
#include <fstream>

#define private public
#define protected public
#include "Voxelyze.h"
#include "VX_Voxel.h"
#include "VX_Link.h"
#include "VX_Material.h"
#include "VX_MaterialVoxel.h"
#include "VX_MaterialLink.h"
#include "VX_External.h"
#include "VX_Collision.h"
#include "VX_MeshRender.h"
#include "VX_LinearSolver.h"
#undef private
#undef protected

#include "voxelyze_b200.h"

namespace {

struct MatDesc {
    vx_material_desc d;
    std::vector<float> strain, stress;
};

struct Member {
    CVoxelyze* sim = nullptr;
    std::vector<CVX_Material*> mats;
    std::vector<int> voxels;   // global voxel indices, member creation order
};

} // namespace

struct vx_sim {
    double voxel_size = 0.001;
    std::vector<MatDesc> descs;
    std::vector<Member> members;
    std::vector<CVX_Voxel*> vox;          // by global voxel index
    std::vector<int> vox_member;
    std::vector<CVX_Link*> links;         // by global link index
    std::unordered_map<const CVX_Voxel*, int> vox_index;
    float grav = 0.f;
    bool floor_on = false, collisions = false;
    std::string err;
    CVX_MeshRender* mesh = nullptr;
    std::vector<unsigned char> colors;     // rgba per material
};

static int fail(vx_sim* s, int code, const char* msg) { if (s) s->err = msg; return code; }

static bool apply_desc(CVX_Material* m, const MatDesc& md, std::string& err)
{
    const vx_material_desc& d = md.d;
    bool ok = true;
    switch (d.model) {
    case VX_MODEL_LINEAR:   ok = m->setModelLinear(d.youngs_modulus, d.fail_stress); break;
    case VX_MODEL_BILINEAR: ok = m->setModelBilinear(d.youngs_modulus, d.plastic_modulus, d.yield_stress, d.fail_stress); break;
    case VX_MODEL_DATA: {
        std::vector<float> a = md.strain, b = md.stress;
        ok = m->setModel((int)a.size(), a.data(), b.data());
        break; }
    default: err = "unknown model"; return false;
    }
    if (!ok) { err = m->lastError(); return false; }
    m->setDensity(d.density);
    m->setPoissonsRatio(d.poissons_ratio);
    m->setCte(d.cte);
    m->setStaticFriction(d.mu_static);
    m->setKineticFriction(d.mu_kinetic);
    m->setInternalDamping(d.zeta_internal);
    m->setGlobalDamping(d.zeta_global);
    m->setCollisionDamping(d.zeta_collision);
    m->setExternalScaleFactor(Vec3D<double>(d.ext_scale[0], d.ext_scale[1], d.ext_scale[2]));
    return true;
}

static void free_members(vx_sim* s)
{
    for (auto& m : s->members) delete m.sim;
    s->members.clear(); s->vox.clear(); s->vox_member.clear(); s->links.clear(); s->vox_index.clear();
}

static bool make_member(vx_sim* s, Member& m)
{
    m.sim = new CVoxelyze(s->voxel_size);
    m.sim->setGravity(s->grav);
    m.sim->enableFloor(s->floor_on);
    for (auto& md : s->descs) {
        CVX_Material* pm = m.sim->addMaterial(1e6f, 1e3f);
        if (!apply_desc(pm, md, s->err)) return false;
        m.mats.push_back(pm);
    }
    return true;
}

extern "C" {

int vx_abi_version(void) { return VX_ABI_VERSION; }
const char* vx_backend(void) {
#ifdef USE_OMP
    return "reference-omp";
#else
    return "reference";
#endif
}

int vx_create(double voxel_size, int, vx_sim** out)
{
    if (!out) return VX_ERR_ARG;
    vx_sim* s = new vx_sim; s->voxel_size = voxel_size; *out = s; return VX_OK;
}
void vx_destroy(vx_sim* s) { if (!s) return; free_members(s); delete s; }
const char* vx_last_error(const vx_sim* s) { return s ? s->err.c_str() : "null handle"; }

int vx_set_materials(vx_sim* s, int n, const vx_material_desc* descs)
{
    if (!s || n < 0 || (n && !descs)) return VX_ERR_ARG;
    std::vector<MatDesc> nd(n);
    for (int i = 0; i < n; i++) {
        nd[i].d = descs[i];
        if (descs[i].model == VX_MODEL_DATA) {
            nd[i].strain.assign(descs[i].strain, descs[i].strain + descs[i].n_points);
            nd[i].stress.assign(descs[i].stress, descs[i].stress + descs[i].n_points);
        }
        nd[i].d.strain = nd[i].d.stress = nullptr;
    }
    if (!s->members.empty()) {
        if ((int)s->descs.size() != n) return fail(s, VX_ERR_ARG, "material count changed after voxels were set");
        for (auto& m : s->members)
            for (int i = 0; i < n; i++)
                if (!apply_desc(m.mats[i], nd[i], s->err)) return VX_ERR_MATERIAL;
    } else {
        // validate on a scratch material so that errors surface here
        for (int i = 0; i < n; i++) { CVX_MaterialVoxel tmp(1e6f, 1e3f, s->voxel_size); if (!apply_desc(&tmp, nd[i], s->err)) return VX_ERR_MATERIAL; }
    }
    s->descs.swap(nd);
    return VX_OK;
}

int vx_get_voxmat(const vx_sim* s, int mat, vx_voxmat_row* o)
{
    if (!s || !o || mat < 0 || mat >= (int)s->descs.size()) return VX_ERR_ARG;
    CVX_MaterialVoxel tmp(1e6f, 1e3f, s->voxel_size);
    CVX_MaterialVoxel* m = &tmp;
    std::string e;
    if (!s->members.empty()) m = (CVX_MaterialVoxel*)s->members[0].mats[mat];
    else if (!apply_desc(&tmp, s->descs[mat], e)) return VX_ERR_MATERIAL;
    o->nom_size = m->nomSize;
    Vec3D<double> sz = m->size(); o->size[0] = sz.x; o->size[1] = sz.y; o->size[2] = sz.z;
    o->E = m->E; o->nu = m->nu; o->rho = m->rho; o->cte = m->alphaCTE; o->mu_static = m->muStatic; o->mu_kinetic = m->muKinetic;
    o->zeta_internal = m->zetaInternal; o->zeta_global = m->zetaGlobal; o->zeta_collision = m->zetaCollision;
    o->e_hat = m->_eHat; o->mass = m->_mass; o->mass_inv = m->_massInverse; o->sqrt_mass = m->_sqrtMass; o->first_moment = m->_firstMoment;
    o->moment_inertia = m->_momentInertia; o->moment_inertia_inv = m->_momentInertiaInverse;
    o->two_sq_m_e_s = m->_2xSqMxExS; o->two_sq_i_e_s3 = m->_2xSqIxExSxSxS;
    o->eps_yield = m->epsilonYield; o->eps_fail = m->epsilonFail; o->sigma_yield = m->sigmaYield; o->sigma_fail = m->sigmaFail;
    o->linear = m->linear ? 1 : 0; o->n_curve = (int)m->strainData.size();
    return VX_OK;
}

static int with_linkmat(vx_sim* s, int a, int b, vx_linkmat_row* o, float* strain, float* stress, int cap)
{
    if (!s || a < 0 || b < 0 || a >= (int)s->descs.size() || b >= (int)s->descs.size()) return VX_ERR_ARG;
    CVX_MaterialVoxel ma(1e6f, 1e3f, s->voxel_size), mb(1e6f, 1e3f, s->voxel_size);
    if (!apply_desc(&ma, s->descs[a], s->err) || !apply_desc(&mb, s->descs[b], s->err)) return VX_ERR_MATERIAL;
    CVX_MaterialLink* l = (a == b) ? new CVX_MaterialLink(&ma, &ma) : new CVX_MaterialLink(&ma, &mb);
    if (o) {
        o->mat_a = std::min(a, b); o->mat_b = std::max(a, b);
        o->linear = l->linear ? 1 : 0; o->n_curve = (int)l->strainData.size();
        o->E = l->E; o->nu = l->nu; o->e_hat = l->_eHat;
        o->eps_yield = l->epsilonYield; o->eps_fail = l->epsilonFail; o->sigma_yield = l->sigmaYield; o->sigma_fail = l->sigmaFail;
        o->a1 = l->_a1; o->a2 = l->_a2; o->b1 = l->_b1; o->b2 = l->_b2; o->b3 = l->_b3;
        o->sq_a1 = l->_sqA1; o->sq_a2_ip = l->_sqA2xIp; o->sq_b1 = l->_sqB1; o->sq_b2_fmp = l->_sqB2xFMp; o->sq_b3_ip = l->_sqB3xIp;
    }
    int n = (int)l->strainData.size();
    if (strain && stress) for (int i = 0; i < n && i < cap; i++) { strain[i] = l->strainData[i]; stress[i] = l->stressData[i]; }
    delete l;
    return n;
}
int vx_get_linkmat(vx_sim* s, int a, int b, vx_linkmat_row* o) { int r = with_linkmat(s, a, b, o, nullptr, nullptr, 0); return r < 0 ? r : VX_OK; }
int vx_get_linkmat_curve(vx_sim* s, int a, int b, float* strain, float* stress, int cap) { return with_linkmat(s, a, b, nullptr, strain, stress, cap); }

int vx_set_voxels(vx_sim* s, int n, const int32_t* ijk, const uint16_t* mat, const int32_t* sim_id, const uint32_t* flags)
{
    if (!s || n < 0 || (n && (!ijk || !mat))) return VX_ERR_ARG;
    if (flags) for (int i = 0; i < n; i++) if (flags[i] & VX_VF_GHOST) return fail(s, VX_ERR_UNSUPPORTED, "ghost voxels are not a reference concept");
    free_members(s);
    std::map<int, int> member_of;   // sim id -> member slot, in order of first appearance
    s->vox.assign(n, nullptr); s->vox_member.assign(n, 0);
    for (int i = 0; i < n; i++) {
        int sid = sim_id ? sim_id[i] : 0;
        auto it = member_of.find(sid);
        int slot;
        if (it == member_of.end()) {
            slot = (int)s->members.size(); member_of[sid] = slot; s->members.emplace_back();
            if (!make_member(s, s->members.back())) return VX_ERR_MATERIAL;
        } else slot = it->second;
        if (mat[i] >= s->descs.size()) return fail(s, VX_ERR_ARG, "material index out of range");
        Member& m = s->members[slot];
        if (m.sim->voxel(ijk[3*i], ijk[3*i+1], ijk[3*i+2])) return fail(s, VX_ERR_TOPOLOGY, "duplicate voxel");
        CVX_Voxel* v = m.sim->setVoxel(m.mats[mat[i]], ijk[3*i], ijk[3*i+1], ijk[3*i+2]);
        s->vox[i] = v; s->vox_member[i] = slot; m.voxels.push_back(i); s->vox_index[v] = i;
    }
    // global link order: member after member is NOT creation order when members interleave;
    // define it as: replay voxel creation order, and for each voxel the links it created.
    // The reference appends to linksList at creation, so per member linksList is already in
    // that order; merge by the creating voxel (= the later-created end of the link).
    std::vector<std::pair<long long, CVX_Link*>> keyed;
    for (auto& m : s->members) {
        const std::vector<CVX_Link*>* ll = m.sim->linkList();
        long long seq = 0;
        for (CVX_Link* l : *ll) {
            int a = s->vox_index[l->pVNeg], b = s->vox_index[l->pVPos];
            keyed.push_back({ (long long)std::max(a, b) * (1LL << 20) + (seq++ & 0xFFFFF), l });
        }
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<long long, CVX_Link*>& x, const std::pair<long long, CVX_Link*>& y) { return (x.first >> 20) < (y.first >> 20); });
    for (auto& k : keyed) s->links.push_back(k.second);
    if (s->collisions) for (auto& m : s->members) m.sim->enableCollisions(true);
    return VX_OK;
}

int vx_voxel_count(const vx_sim* s) { return s ? (int)s->vox.size() : 0; }
int vx_link_count(const vx_sim* s) { return s ? (int)s->links.size() : 0; }
int vx_get_links(const vx_sim* s, int32_t* vn, int32_t* vp, uint8_t* ax)
{
    if (!s) return VX_ERR_ARG;
    for (size_t i = 0; i < s->links.size(); i++) {
        CVX_Link* l = s->links[i];
        if (vn) vn[i] = s->vox_index.at(l->pVNeg);
        if (vp) vp[i] = s->vox_index.at(l->pVPos);
        if (ax) ax[i] = (uint8_t)l->axis;
    }
    return VX_OK;
}

int vx_set_externals(vx_sim* s, int n, const int32_t* voxel, const uint8_t* dof, const float* force, const float* moment, const double* tr, const double* rot)
{
    if (!s || n < 0 || (n && (!voxel || !dof))) return VX_ERR_ARG;
    for (CVX_Voxel* v : s->vox) if (v->externalExists()) v->external()->reset();
    for (int i = 0; i < n; i++) {
        if (voxel[i] < 0 || voxel[i] >= (int)s->vox.size()) return fail(s, VX_ERR_ARG, "external voxel index out of range");
        CVX_External* e = s->vox[voxel[i]]->external();
        e->reset();
        e->dofFixed = dof[i] & 0x3F;
        if (force)  e->extForce  = Vec3D<float>(force[3*i], force[3*i+1], force[3*i+2]);
        if (moment) e->extMoment = Vec3D<float>(moment[3*i], moment[3*i+1], moment[3*i+2]);
        if (tr)  e->extTranslation = Vec3D<double>(tr[3*i], tr[3*i+1], tr[3*i+2]);
        if (rot) e->extRotation    = Vec3D<double>(rot[3*i], rot[3*i+1], rot[3*i+2]);
        e->rotationChanged();
    }
    return VX_OK;
}

int vx_set_gravity(vx_sim* s, float g) { if (!s) return VX_ERR_ARG; s->grav = g; for (auto& m : s->members) m.sim->setGravity(g); return VX_OK; }
int vx_enable_floor(vx_sim* s, int e) { if (!s) return VX_ERR_ARG; s->floor_on = e != 0; for (auto& m : s->members) m.sim->enableFloor(e != 0); return VX_OK; }
int vx_enable_collisions(vx_sim* s, int e) { if (!s) return VX_ERR_ARG; s->collisions = e != 0; for (auto& m : s->members) m.sim->enableCollisions(e != 0); return VX_OK; }
int vx_set_collision_envelope(vx_sim* s, float r) { if (!s) return VX_ERR_ARG; CVX_Collision::envelopeRadius = r; return VX_OK; }

int vx_set_temperature_all(vx_sim* s, float t) { if (!s) return VX_ERR_ARG; for (auto& m : s->members) m.sim->setAmbientTemperature(t, true); return VX_OK; }
int vx_set_temperature_members(vx_sim* s, int n, const float* t)
{
    if (!s || !t || n != (int)s->members.size()) return VX_ERR_ARG;
    for (int i = 0; i < n; i++) s->members[i].sim->setAmbientTemperature(t[i], true);
    return VX_OK;
}
int vx_set_temperature(vx_sim* s, int n, const float* t)
{
    if (!s || !t || n != (int)s->vox.size()) return VX_ERR_ARG;
    for (int i = 0; i < n; i++) s->vox[i]->setTemperature(t[i]);
    return VX_OK;
}

int vx_step(vx_sim* s, float dt, int n_steps, int* diverged_step)
{
    if (!s || n_steps < 0) return VX_ERR_ARG;
    for (int k = 0; k < n_steps; k++) {
        bool ok = true;
        for (auto& m : s->members) ok = m.sim->doTimeStep(dt) && ok;
        if (!ok) { if (diverged_step) *diverged_step = k; return VX_DIVERGED; }
    }
    return VX_OK;
}

int vx_recommended_dt(vx_sim* s, float* dt)
{
    if (!s || !dt) return VX_ERR_ARG;
    float best = 0.f; bool any = false;
    for (auto& m : s->members) { float d = m.sim->recommendedTimeStep(); if (!any || d < best) best = d; any = true; }
    *dt = best; return VX_OK;
}
int vx_reset(vx_sim* s) { if (!s) return VX_ERR_ARG; for (auto& m : s->members) m.sim->resetTime(); return VX_OK; }
float vx_time(const vx_sim* s) { return (s && !s->members.empty()) ? s->members[0].sim->currentTime : 0.f; }
int vx_set_clock(vx_sim* s, float time, float previous_dt)
{
    if (!s || !(time >= 0.f) || !(previous_dt >= 0.f)) return VX_ERR_ARG;
    for (auto& m : s->members) { m.sim->currentTime = time; for (int i = 0; i < m.sim->voxelCount(); i++) m.sim->voxel(i)->previousDt = previous_dt; }
    return VX_OK;
}

static void put3(double* d, const Vec3D<double>& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
static Vec3D<double> get3(const double* d) { return Vec3D<double>(d[0], d[1], d[2]); }

int vx_download(vx_sim* s, int field, int first, int count, void* dst)
{
    if (!s || !dst || first < 0 || count < 0) return VX_ERR_ARG;
    bool is_link = field >= 16;
    int total = is_link ? (int)s->links.size() : (int)s->vox.size();
    if (first + count > total) return VX_ERR_ARG;
    double* d = (double*)dst; float* f = (float*)dst; uint32_t* u = (uint32_t*)dst;
    for (int k = 0; k < count; k++) {
        if (!is_link) {
            CVX_Voxel* v = s->vox[first + k];
            switch (field) {
            case VX_F_POS: put3(d + 3*k, v->pos); break;
            case VX_F_ORIENT: d[4*k] = v->orient.w; d[4*k+1] = v->orient.x; d[4*k+2] = v->orient.y; d[4*k+3] = v->orient.z; break;
            case VX_F_LINMOM: put3(d + 3*k, v->linMom); break;
            case VX_F_ANGMOM: put3(d + 3*k, v->angMom); break;
            case VX_F_TEMP: f[k] = v->temp; break;
            case VX_F_VOXFLAGS: u[k] = (v->isFloorStaticFriction() ? VX_VF_STATIC_FRICTION : 0) | (v->isSurface() ? VX_VF_SURFACE : 0) |
                                       (s->floor_on && !v->isFloorEnabled() ? VX_VF_FLOOR_OFF : 0) | (!s->floor_on && v->isFloorEnabled() ? VX_VF_FLOOR_ON : 0); break;
            case VX_F_PSTRAIN: { const Vec3D<float> p = v->poissonsStrainInvalid ? v->strain(true) : v->pStrain;      // what the links of the next step will read, cache untouched
                                 f[3*k] = p.x; f[3*k+1] = p.y; f[3*k+2] = p.z; break; }
            default: return VX_ERR_ARG;
            }
        } else {
            CVX_Link* l = s->links[first + k];
            switch (field) {
            case VX_F_FORCE_NEG: put3(d + 3*k, l->forceNeg); break;
            case VX_F_FORCE_POS: put3(d + 3*k, l->forcePos); break;
            case VX_F_MOMENT_NEG: put3(d + 3*k, l->momentNeg); break;
            case VX_F_MOMENT_POS: put3(d + 3*k, l->momentPos); break;
            case VX_F_POS2: put3(d + 3*k, l->pos2); break;
            case VX_F_ANGLE1V: put3(d + 3*k, l->angle1v); break;
            case VX_F_ANGLE2V: put3(d + 3*k, l->angle2v); break;
            case VX_F_STRAIN: f[k] = l->strain; break;
            case VX_F_MAXSTRAIN: f[k] = l->maxStrain; break;
            case VX_F_STRAINOFFSET: f[k] = l->strainOffset; break;
            case VX_F_STRESS: f[k] = l->_stress; break;
            case VX_F_LINKFLAGS: u[k] = (l->smallAngle ? VX_LF_SMALL_ANGLE : 0) | (l->isLocalVelocityValid() ? VX_LF_LOCAL_VEL_VALID : 0)
                                       | (l->isYielded() ? VX_LF_YIELDED : 0) | (l->isFailed() ? VX_LF_FAILED : 0); break;
            default: return VX_ERR_ARG;
            }
        }
    }
    return VX_OK;
}

int vx_upload(vx_sim* s, int field, int first, int count, const void* src)
{
    if (!s || !src || first < 0 || count < 0) return VX_ERR_ARG;
    if (field >= 16) return fail(s, VX_ERR_UNSUPPORTED, "link state upload");
    if (first + count > (int)s->vox.size()) return VX_ERR_ARG;
    const double* d = (const double*)src; const float* f = (const float*)src; const uint32_t* u = (const uint32_t*)src;
    for (int k = 0; k < count; k++) {
        CVX_Voxel* v = s->vox[first + k];
        switch (field) {
        case VX_F_POS: v->pos = get3(d + 3*k); break;
        case VX_F_ORIENT: v->orient = Quat3D<double>(d[4*k], d[4*k+1], d[4*k+2], d[4*k+3]); break;
        case VX_F_LINMOM: v->linMom = get3(d + 3*k); break;
        case VX_F_ANGMOM: v->angMom = get3(d + 3*k); break;
        case VX_F_TEMP: v->setTemperature(f[k]); break;
        case VX_F_PSTRAIN: v->pStrain = Vec3D<float>(f[3*k], f[3*k+1], f[3*k+2]); v->poissonsStrainInvalid = false; break;
        case VX_F_VOXFLAGS: v->setFloorStaticFriction((u[k] & VX_VF_STATIC_FRICTION) != 0);
                            v->enableFloor((u[k] & VX_VF_FLOOR_OFF) ? false : ((u[k] & VX_VF_FLOOR_ON) ? true : s->floor_on)); break;
        default: return VX_ERR_ARG;
        }
    }
    return VX_OK;
}

int vx_collision_pairs(vx_sim* s, int32_t* pairs, int cap, int* n_pairs)
{
    if (!s) return VX_ERR_ARG;
    int n = 0;
    for (auto& m : s->members) {
        const std::vector<CVX_Collision*>* cl = m.sim->collisionList();
        for (CVX_Collision* c : *cl) {
            if (pairs && n < cap) { pairs[2*n] = s->vox_index.at(c->voxel1()); pairs[2*n+1] = s->vox_index.at(c->voxel2()); }
            n++;
        }
    }
    if (n_pairs) *n_pairs = n;
    return VX_OK;
}

int vx_state_info(vx_sim* s, int info, int type, float* out)
{
    if (!s || !out || s->members.size() != 1) return VX_ERR_ARG;
    *out = s->members[0].sim->stateInfo((CVoxelyze::stateInfoType)info, (CVoxelyze::valueType)type);
    return VX_OK;
}

int vx_set_stream(vx_sim*, uint64_t) { return VX_ERR_UNSUPPORTED; }
int vx_pose_plane(vx_sim*, int, uint64_t*, uint64_t*, int*, int*) { return VX_ERR_UNSUPPORTED; }
int vx_halo_import(vx_sim*, int, uint64_t, uint64_t, int) { return VX_ERR_UNSUPPORTED; }
int vx_halo_import_on(vx_sim*, int, uint64_t, uint64_t, int, uint64_t) { return VX_ERR_UNSUPPORTED; }
int vx_step_begin(vx_sim*, float) { return VX_ERR_UNSUPPORTED; }
int vx_step_enqueue(vx_sim*, int) { return VX_ERR_UNSUPPORTED; }
int vx_step_end(vx_sim*, int*) { return VX_ERR_UNSUPPORTED; }
int vx_peer_export(vx_sim*, int, int, vx_peer_desc*) { return VX_ERR_UNSUPPORTED; }
int vx_peer_attach(vx_sim*, int, const vx_peer_desc*) { return VX_ERR_UNSUPPORTED; }
int vx_peer_detach(vx_sim*) { return VX_ERR_UNSUPPORTED; }
int vx_slab_step(vx_sim*, float, int, int*) { return VX_ERR_UNSUPPORTED; }
int vx_step_ambient(vx_sim* s, float dt, int n_steps, const float* ambient, int* diverged_step)
{
    if (!s || n_steps < 0 || (n_steps && !ambient)) return VX_ERR_ARG;
    if (diverged_step) *diverged_step = -1;
    for (int k = 0; k < n_steps; k++) {                 // the calls it stands for, one by one
        int rc = vx_set_temperature_all(s, ambient[k]); if (rc != VX_OK) return rc;
        int d = -1;
        rc = vx_step(s, dt, 1, &d);
        if (rc == VX_DIVERGED && diverged_step) *diverged_step = k;
        if (rc != VX_OK) return rc;
    }
    return VX_OK;
}
int vx_slab_step_begin(vx_sim*, float, int) { return VX_ERR_UNSUPPORTED; }
int vx_slab_step_finish(vx_sim*, int*) { return VX_ERR_UNSUPPORTED; }
int vx_slab_exchange(vx_sim*) { return VX_ERR_UNSUPPORTED; }
int vx_save_state(vx_sim*, const char*) { return VX_ERR_UNSUPPORTED; }
int vx_load_state(vx_sim*, const char*) { return VX_ERR_UNSUPPORTED; }
int vx_collision_forces(vx_sim*, int32_t*, float*, int, int*) { return VX_ERR_UNSUPPORTED; }
// what CVX_Link keeps between steps (include/VX_Link.h:74-107), read from and written into the reference's own objects: with these
// records (plus the voxel fields, the cached Poisson strains and the clock) a run of the UNMODIFIED reference can be moved into a
// freshly built CVoxelyze and goes on bit for bit (tests/test_oracle.py) -- which is what the C-ABI's state-carrying calls rest on
int vx_download_link_state(vx_sim* s, int first, int count, vx_link_state* dst)
{
    if (!s || !dst || first < 0 || count < 0 || first + count > (int)s->links.size()) return VX_ERR_ARG;
    for (int k = 0; k < count; k++) {
        const CVX_Link* l = s->links[first + k];
        vx_link_state& r = dst[k];
        memset(&r, 0, sizeof(r));
        put3(r.pos2, l->pos2); put3(r.angle1v, l->angle1v); put3(r.angle2v, l->angle2v);
        r.strain = l->strain; r.max_strain = l->maxStrain; r.strain_offset = l->strainOffset; r.stress = l->_stress;
        r.flags = (l->smallAngle ? VX_LF_SMALL_ANGLE : 0) | (l->isLocalVelocityValid() ? VX_LF_LOCAL_VEL_VALID : 0)
                | (l->isYielded() ? VX_LF_YIELDED : 0) | (l->isFailed() ? VX_LF_FAILED : 0);
    }
    return VX_OK;
}
int vx_upload_link_state(vx_sim* s, int first, int count, const vx_link_state* src)
{
    if (!s || !src || first < 0 || count < 0 || first + count > (int)s->links.size()) return VX_ERR_ARG;
    for (int k = 0; k < count; k++) {
        CVX_Link* l = s->links[first + k]; const vx_link_state& r = src[k];
        l->pos2 = get3(r.pos2); l->angle1v = get3(r.angle1v); l->angle2v = get3(r.angle2v);
        l->strain = r.strain; l->maxStrain = r.max_strain; l->strainOffset = r.strain_offset; l->_stress = r.stress;
        l->smallAngle = (r.flags & VX_LF_SMALL_ANGLE) != 0;
        l->setBoolState(CVX_Link::LOCAL_VELOCITY_VALID, (r.flags & VX_LF_LOCAL_VEL_VALID) != 0);
    }
    for (CVX_Voxel* v : s->vox) v->poissonsStrainInvalid = true;
    return VX_OK;
}
int64_t vx_launch_count(const vx_sim*) { return 0; }
int vx_sync(vx_sim*) { return VX_OK; }
int vx_set_path(vx_sim*, int) { return VX_OK; }
int vx_active_path(const vx_sim*) { return 0; }
const char* vx_kernel_name(const vx_sim*) { return "cpu (reference CVoxelyze::doTimeStep)"; }
int vx_step_profile(vx_sim*, float, int, float*, int*) { return VX_ERR_UNSUPPORTED; }
int vx_prepare(vx_sim*) { return VX_OK; }
int vx_download_voxel_state(vx_sim* s, int first, int count, vx_voxel_state* dst)
{
    if (!s || !dst || first < 0 || count < 0) return VX_ERR_ARG;
    for (int k = 0; k < count; k++) {
        vx_voxel_state& r = dst[k];
        int rc = vx_download(s, VX_F_POS, first + k, 1, r.pos);
        if (rc == VX_OK) rc = vx_download(s, VX_F_ORIENT, first + k, 1, r.orient);
        if (rc == VX_OK) rc = vx_download(s, VX_F_LINMOM, first + k, 1, r.linmom);
        if (rc == VX_OK) rc = vx_download(s, VX_F_ANGMOM, first + k, 1, r.angmom);
        if (rc == VX_OK) rc = vx_download(s, VX_F_TEMP, first + k, 1, &r.temp);
        if (rc == VX_OK) rc = vx_download(s, VX_F_VOXFLAGS, first + k, 1, &r.flags);
        if (rc != VX_OK) return rc;
    }
    return VX_OK;
}

// ---- surface mesh: the reference's own CVX_MeshRender (src/VX_MeshRender.cpp), members read through the opened-up header
static void apply_colors(vx_sim* s)
{
    if (s->members.size() != 1) return;
    for (size_t i = 0; i < s->members[0].mats.size() && 4 * i + 3 < s->colors.size(); i++)
        s->members[0].mats[i]->setColor(s->colors[4 * i], s->colors[4 * i + 1], s->colors[4 * i + 2], s->colors[4 * i + 3]);
}
int vx_mesh_set_material_colors(vx_sim* s, int n, const unsigned char* rgba)
{
    if (!s || n < 0 || (n && !rgba)) return VX_ERR_ARG;
    s->colors.assign(rgba, rgba + 4 * (size_t)n);
    apply_colors(s);
    return VX_OK;
}
int vx_mesh_build(vx_sim* s, int* nv, int* nq)
{
    if (!s || s->members.size() != 1) return VX_ERR_UNSUPPORTED;
    delete s->mesh;
    apply_colors(s);
    s->mesh = new CVX_MeshRender(s->members[0].sim);
    if (nv) *nv = (int)s->mesh->vertices.size() / 3;
    if (nq) *nq = (int)s->mesh->quads.size() / 4;
    return VX_OK;
}
int vx_mesh_update(vx_sim* s, int coloring, int state_type)
{
    if (!s) return VX_ERR_ARG;
    if (!s->mesh) { int rc = vx_mesh_build(s, nullptr, nullptr); if (rc != VX_OK) return rc; }
    s->mesh->updateMesh((CVX_MeshRender::viewColoring)coloring, (CVoxelyze::stateInfoType)state_type);
    return VX_OK;
}
int vx_mesh_counts(vx_sim* s, int* nv, int* nq)
{
    if (!s) return VX_ERR_ARG;
    if (nv) *nv = s->mesh ? (int)s->mesh->vertices.size() / 3 : 0;
    if (nq) *nq = s->mesh ? (int)s->mesh->quads.size() / 4 : 0;
    return VX_OK;
}
int vx_mesh_download(vx_sim* s, float* vertices, int32_t* quads, float* normals, float* colors, int32_t* quad_voxel)
{
    if (!s || !s->mesh) return VX_ERR_ARG;
    CVX_MeshRender& m = *s->mesh;
    if (vertices) std::copy(m.vertices.begin(), m.vertices.end(), vertices);
    if (quads) std::copy(m.quads.begin(), m.quads.end(), quads);
    if (normals) std::copy(m.quadNormals.begin(), m.quadNormals.end(), normals);
    if (colors) std::copy(m.quadColors.begin(), m.quadColors.end(), colors);
    if (quad_voxel) for (size_t i = 0; i < m.quadVoxIndices.size(); i++) quad_voxel[i] = s->members[0].voxels[m.quadVoxIndices[i]];
    return VX_OK;
}
int vx_mesh_device(vx_sim*, uint64_t*, uint64_t*, uint64_t*, uint64_t*) { return VX_ERR_UNSUPPORTED; }
int vx_collision_stats(vx_sim* s, int* n_pairs, int* n_rebuilds)
{
    if (n_rebuilds) *n_rebuilds = -1;                  // not counted here
    return vx_collision_pairs(s, nullptr, 0, n_pairs);
}

// ---- static solve: the reference's own CVX_LinearSolver::solve (calculateA, applyBX, postResults run unmodified).  It is
// compiled with PARDISO_5 defined (oracle/Makefile) and the two entry points of the closed-source PARDISO library it
// declares (include/VX_LinearSolver.h:26-27) are provided below by a plain banded Cholesky factorisation of the
// upper-triangular CSR matrix it passes -- enough for the model sizes the tests use.
extern "C" void pardisoinit(void*, int*, int*, int* iparm, double*, int* error) { for (int i = 0; i < 64; i++) iparm[i] = 0; *error = 0; }
extern "C" void pardiso(void*, int*, int*, int*, int* phase, int* n_, double* a, int* ia, int* ja, int*, int*, int*, int*, double* b, double* x, int* error, double*)
{
    *error = 0;
    if (*phase != 33) return;                          // analysis / factorisation / release: all work is done in the solve phase
    const int n = *n_;
    int bw = 0;
    for (int i = 0; i < n; i++) for (int k = ia[i] - 1; k < ia[i + 1] - 1; k++) bw = std::max(bw, ja[k] - 1 - i);
    const size_t W = (size_t)bw + 1;
    if ((double)n * W > 4e8) { *error = -2; return; }
    std::vector<double> U((size_t)n * W, 0.0);
    for (int i = 0; i < n; i++) for (int k = ia[i] - 1; k < ia[i + 1] - 1; k++) U[(size_t)i * W + (ja[k] - 1 - i)] += a[k];
    std::vector<double> diag0(n);
    for (int k = 0; k < n; k++) diag0[k] = U[(size_t)k * W];
    for (int k = 0; k < n; k++) {
        double* row = &U[(size_t)k * W];
        if (!(row[0] > 1e-11 * diag0[k])) { *error = -4; return; }      // pivot lost to cancellation: singular to working precision
        const double piv = std::sqrt(row[0]);
        const int m = std::min(bw, n - 1 - k);
        for (int j = 0; j <= m; j++) row[j] /= piv;
        for (int i = 1; i <= m; i++) {
            const double f = row[i];
            if (f == 0.0) continue;
            double* ri = &U[(size_t)(k + i) * W];
            for (int j = i; j <= m; j++) ri[j - i] -= f * row[j];
        }
    }
    std::vector<double> y(b, b + n);
    for (int k = 0; k < n; k++) {
        const double* row = &U[(size_t)k * W];
        y[k] /= row[0];
        const int m = std::min(bw, n - 1 - k);
        for (int j = 1; j <= m; j++) y[k + j] -= row[j] * y[k];
    }
    for (int k = n - 1; k >= 0; k--) {
        const double* row = &U[(size_t)k * W];
        const int m = std::min(bw, n - 1 - k);
        double t = y[k];
        for (int j = 1; j <= m; j++) t -= row[j] * y[k + j];
        y[k] = t / row[0];
    }
    std::copy(y.begin(), y.end(), x);
}

int vx_linear_solve(vx_sim* s, double, int, int* iterations, double* rel_residual)
{
    if (iterations) *iterations = 0;
    if (rel_residual) *rel_residual = 0.0;
    if (!s) return VX_ERR_ARG;
    if (s->vox.empty()) return fail(s, VX_ERR_ARG, "vx_linear_solve: no voxels");
    for (auto& m : s->members) {
        std::streambuf* keep = std::cout.rdbuf(nullptr);          // solve() announces itself on stdout
        CVX_LinearSolver solver(m.sim);
        solver.msglvl = 0;
        const bool ok = solver.solve();
        std::cout.rdbuf(keep);
        if (!ok) return fail(s, VX_ERR_SOLVER, solver.errorMsg.c_str());
    }
    return VX_OK;
}

} // extern "C"

// vx_slabbed_*: host-side composition over the entry points above (shared with the product library: the partition logic under test)
#include "../voxelyze_b200/csrc/vx_slabbed.hpp"
