// oracle_port.cpp -- TEST INFRASTRUCTURE ONLY.  PARITY PINNED (see below).
//
// A plain, scalar CPU restatement of the reference's explicit dynamics step
// (jonhiller/Voxelyze, CVoxelyze::doTimeStep and everything it calls) behind the C-ABI of
// include/voxelyze_b200.h.  It is the checker for the CUDA product on machines where the
// reference sources are absent (the GPU box).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it; the product never does.
//
// Pinning: tests/test_oracle_vs_reference.py drives this file and oracle/_ref/libvxref.so
// (the unmodified reference) with identical calls and requires BIT-EXACT equality of every
// voxel and link field on all scenarios (both are x86-64 SSE2 + glibc libm, no FMA), and
// tests/golden/ holds reference outputs for the GPU box where _ref may be rebuilt but
// /root/reference is not available.
//
// Every function cites the reference file:line it restates.  Rounding points (where the
// reference computes in float before promoting) follow SURVEY.md Appendix A.

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <unordered_map>
#include <algorithm>

#include "voxelyze_b200.h"

namespace {

// ------------------------------------------------------------------ small math (Vec3D.h / Quat3D.h)
struct V3 { double x = 0, y = 0, z = 0; };
struct V3f { float x = 0, y = 0, z = 0; };
struct Q4 { double w = 1, x = 0, y = 0, z = 0; };

inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 scale(double f, V3 a) { return {f * a.x, f * a.y, f * a.z}; }           // Vec3D.h:53,62
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline double len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }             // Vec3D.h:98

// Quat3D.h:83 quaternion product
inline Q4 qmul(const Q4& a, const Q4& f)
{
    return { a.w * f.w - a.x * f.x - a.y * f.y - a.z * f.z,
             a.w * f.x + a.x * f.w + a.y * f.z - a.z * f.y,
             a.w * f.y - a.x * f.z + a.y * f.w + a.z * f.x,
             a.w * f.z + a.x * f.y - a.y * f.x + a.z * f.w };
}
inline Q4 conj(const Q4& q) { return {q.w, -q.x, -q.y, -q.z}; }                    // Quat3D.h:94

// Quat3D.h:170-177
inline V3 rotate(const Q4& q, V3 f)
{
    double tw = f.x * q.x + f.y * q.y + f.z * q.z;
    double tx = f.x * q.w - f.y * q.z + f.z * q.y;
    double ty = f.x * q.z + f.y * q.w - f.z * q.x;
    double tz = -f.x * q.y + f.y * q.x + f.z * q.w;
    return { q.w * tx + q.x * tw + q.y * tz - q.z * ty,
             q.w * ty - q.x * tz + q.y * tw + q.z * tx,
             q.w * tz + q.x * ty - q.y * tx + q.z * tw };
}
// Quat3D.h:188-195
inline V3 rotate_inv(const Q4& q, V3 f)
{
    double tw = q.x * f.x + q.y * f.y + q.z * f.z;
    double tx = q.w * f.x - q.y * f.z + q.z * f.y;
    double ty = q.w * f.y + q.x * f.z - q.z * f.x;
    double tz = q.w * f.z - q.x * f.y + q.y * f.x;
    return { tw * q.x + tx * q.w + ty * q.z - tz * q.y,
             tw * q.y - tx * q.z + ty * q.w + tz * q.x,
             tw * q.z + tx * q.y - ty * q.x + tz * q.w };
}
// Quat3D.h:117-122
inline V3 to_rotation_vector(const Q4& q)
{
    if (q.w >= 1.0 || q.w <= -1.0) return {};
    double sl = 1.0 - q.w * q.w;
    V3 v2 = scale(2.0, V3{q.x, q.y, q.z});
    if (sl < 2.4e-3) return scale(std::sqrt((2 - 2 * q.w) / sl), v2);
    V3 v3 = scale(std::acos(q.w), v2);
    double inv = 1.0 / std::sqrt(sl);                                               // Vec3D.h:65
    return scale(inv, v3);
}
// Quat3D.h:124-139
inline Q4 from_rotation_vector(V3 v)
{
    V3 th = scale(0.5, v);                                                          // VecIn/2 -> Inv = 1.0/2
    double m2 = len2(th), w, s;
    if (m2 * m2 < 5.328e-15) { w = 1.0 - 0.5 * m2; s = 1.0 - m2 / 6.0; }
    else { double m = std::sqrt(m2); w = std::cos(m); s = std::sin(m) / m; }
    return {w, th.x * s, th.y * s, th.z * s};
}
// Quat3D.h:141-166 (applied to a quaternion that is the identity on entry, VX_Link.cpp:87,105)
inline Q4 from_angle_to_pos_x(V3 from)
{
    Q4 q;
    if (from.x == 0 && from.y == 0 && from.z == 0) return q;
    double yox = from.y / from.x, zox = from.z / from.x;
    const double SA = 1.732e-2;
    if (yox < SA && yox > -SA && zox < SA && zox > -SA) {
        q.x = 0; q.y = 0.5 * zox; q.z = -0.5 * yox;
        q.w = 1 + 0.5 * (-q.y * q.y - q.z * q.z);
        return q;
    }
    V3 n = from;
    double l = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    if (l > 0) { double li = 1.0 / l; n.x *= li; n.y *= li; n.z *= li; }
    double theta = std::acos(n.x);
    if (theta > 3.14159265358979 - 1e-7) { q.w = 0; q.x = 0; q.y = 1; q.z = 0; return q; }
    double ami = 1.0 / std::sqrt(n.z * n.z + n.y * n.y);
    double a = 0.5 * theta, s = std::sin(a);
    q.w = std::cos(a); q.x = 0; q.y = n.z * ami * s; q.z = -n.y * ami * s;
    return q;
}

// ------------------------------------------------------------------ material model (VX_Material.*)
struct Model {
    bool linear = true;
    float E = 1, sigmaYield = -1, sigmaFail = -1, epsYield = -1, epsFail = -1;
    std::vector<float> strain, stress;
    float nu = 0, rho = 1, cte = 0, muS = 0, muK = 0, zI = 1, zG = 0, zC = 0;
    double ext[3] = {1, 1, 1};
    float eHat = 1;
    std::string err;

    bool isYielded(float s) const { return epsYield != -1.0f && s > epsYield; }    // VX_Material.h:48
    bool isFailed(float s) const { return epsFail != -1.0f && s > epsFail; }       // VX_Material.h:49
    void updateEHat() { eHat = E / ((1 - 2 * nu) * (1 + nu)); }                    // VX_Material.cpp:527

    // VX_Material.cpp:336-367
    bool setLinear(float ym, float fail)
    {
        if (ym <= 0) { err = "Young's modulus must be positive"; return false; }
        if (fail != -1.0f && fail <= 0) { err = "Failure stress must be positive"; return false; }
        float tf = fail; if (tf == -1) tf = 1000000;
        float ts = tf / ym;
        strain = {0, ts}; stress = {0, tf};
        linear = true; E = ym; sigmaYield = fail; sigmaFail = fail;
        epsYield = (fail == -1) ? -1 : ts; epsFail = (fail == -1) ? -1 : ts;
        updateEHat(); return true;
    }
    // VX_Material.cpp:372-419
    bool setBilinear(float ym, float pm, float ys, float fail)
    {
        if (ym <= 0) { err = "Young's modulus must be positive"; return false; }
        if (pm <= 0 || pm >= ym) { err = "Plastic modulus must be positive but less than Young's modulus"; return false; }
        if (ys <= 0) { err = "Yield stress must be positive"; return false; }
        if (fail != -1.0f && fail <= ys) { err = "Failure stress must be positive and greater than the yield stress"; return false; }
        float yStrain = ys / ym;
        float tf = fail; if (tf == -1) tf = 3 * ys;
        float tM = pm, tB = ys - tM * yStrain;
        float tfs = (tf - tB) / tM;
        strain = {0, yStrain, tfs}; stress = {0, ys, tf};
        linear = false; E = ym; sigmaYield = ys; sigmaFail = fail; epsYield = yStrain;
        epsFail = fail == -1.0f ? -1.0f : tfs;
        updateEHat(); return true;
    }
    // VX_Material.cpp:437-470
    bool yieldFromData(float pct = 0.2f)
    {
        sigmaYield = -1.0f; epsYield = -1.0f;
        float oM = E, oB = (-pct / 100 * oM);
        int dp = (int)strain.size() - 1;
        for (int i = 1; i < dp - 1; i++) {
            float x1 = strain[i], x2 = strain[i + 1], y1 = stress[i], y2 = stress[i + 1];
            float tM = (y2 - y1) / (x2 - x1), tB = y1 - tM * x1;
            if (oM != tM) {
                float xi = (tB - oB) / (oM - tM);
                if (xi > x1 && xi < x2) {
                    float p = (xi - x1) / (x2 - x1);
                    sigmaYield = y1 + p * (y2 - y1); epsYield = xi; return true;
                }
            }
        }
        sigmaYield = sigmaFail; epsYield = epsFail; return false;
    }
    // VX_Material.cpp:283-331
    bool setData(int n, const float* ps, const float* pt)
    {
        if (n > 0 && *ps == 0 && *pt == 0) { ps++; pt++; n--; }
        if (n <= 0) { err = "Not enough data points"; return false; }
        if (*ps <= 0 || *pt <= 0) { err = "First stress and strain data points negative or zero"; return false; }
        std::vector<float> a{0}, b{0};
        float sa = 0, sb = 0;
        for (int i = 0; i < n; i++) {
            float x = ps[i], y = pt[i];
            if (x <= sa) { err = "Out of order strain data"; return false; }
            if (y <= sb) err = "Stress data is not monotonically increasing";
            if (i > 0 && (y - sb) / (x - sa) > b[0] / a[0]) {   // 0/0 = NaN: comparison is false (VX_Material.cpp:320)
                err = "Slope of stress/strain curve should never exceed that of the first line segment (youngs modulus)"; return false;
            }
            sa = x; sb = y; a.push_back(x); b.push_back(y);
        }
        strain = a; stress = b;
        E = stress[1] / strain[1];
        sigmaFail = stress.back(); epsFail = strain.back();
        linear = (n == 1);
        if (n == 1 || n == 2) { sigmaYield = stress[1]; epsYield = strain[1]; }
        else yieldFromData();
        updateEHat(); return true;
    }
    // VX_Material.cpp:165-195
    float stressAt(float s, float tss = 0.0f, bool forceLinear = false) const
    {
        if (isFailed(s)) return 0.0f;
        if (s <= strain[1] || linear || forceLinear) {
            if (nu == 0.0f) return E * s;
            return eHat * ((1 - nu) * s + nu * tss);
        }
        int n = (int)strain.size();
        for (int i = 2; i < n; i++) {
            if (s <= strain[i] || i == n - 1) {
                float perc = (s - strain[i - 1]) / (strain[i] - strain[i - 1]);
                float basic = stress[i - 1] + perc * (stress[i] - stress[i - 1]);
                if (nu == 0.0f) return basic;
                float modulus = (stress[i] - stress[i - 1]) / (strain[i] - strain[i - 1]);
                float modHat = modulus / ((1 - 2 * nu) * (1 + nu));
                float effStrain = basic / modulus;
                float effTss = tss * (effStrain / s);
                return modHat * ((1 - nu) * effStrain + nu * effTss);
            }
        }
        return 0.0f;
    }
    // VX_Material.cpp:197-211
    float strainAt(float st) const
    {
        if (st <= stress[1] || linear) return st / E;
        int n = (int)strain.size();
        for (int i = 2; i < n; i++)
            if (st <= stress[i] || i == n - 1) {
                float perc = (st - stress[i - 1]) / (stress[i] - stress[i - 1]);
                return strain[i - 1] + perc * (strain[i] - strain[i - 1]);
            }
        return 0.0f;
    }
    // VX_Material.cpp:214-226
    float modulusAt(float s) const
    {
        if (isFailed(s)) return 0.0f;
        if (s <= strain[1] || linear) return E;
        int n = (int)strain.size();
        for (int i = 2; i < n; i++)
            if (s <= strain[i] || i == n - 1) return (stress[i] - stress[i - 1]) / (strain[i] - strain[i - 1]);
        return 0.0f;
    }
};

// CVX_MaterialVoxel cached quantities (VX_MaterialVoxel.cpp:57-79, VX_MaterialVoxel.h:41-57)
struct VoxMat : Model {
    double nom = 0.001;
    float mass = 0, massInv = 0, sqrtMass = 0, firstMoment = 0, inertia = 0, inertiaInv = 0, sqMES = 0, sqIES3 = 0;
    void derive()
    {
        updateEHat();
        double volume = nom * nom * nom;
        mass = (float)(volume * rho);
        inertia = (float)(mass * nom * nom / 6.0f);
        firstMoment = (float)(mass * nom / 2.0f);
        if (volume == 0 || mass == 0 || inertia == 0) { massInv = sqrtMass = inertiaInv = sqMES = sqIES3 = 0.0f; return; }
        massInv = 1.0f / mass;
        sqrtMass = std::sqrt(mass);
        inertiaInv = 1.0f / inertia;
        sqMES = (float)(2.0f * std::sqrt(mass * E * nom));
        sqIES3 = (float)(2.0f * std::sqrt(inertia * E * nom * nom * nom));
    }
    float globalDampT() const { return zG * sqMES; }
    float globalDampR() const { return zG * sqIES3; }
    float collDampT() const { return zC * sqMES; }
    float penetrationStiffness() const { return (float)(2 * E * nom); }
    double size(int a) const { return nom * ext[a]; }
};

// CVX_MaterialLink (VX_MaterialLink.cpp:45-141)
struct LinkMat : VoxMat {
    int ma = 0, mb = 0;
    float a1 = 0, a2 = 0, b1 = 0, b2 = 0, b3 = 0, sqA1 = 0, sqA2xIp = 0, sqB1 = 0, sqB2xFMp = 0, sqB3xIp = 0;
    void deriveLink()
    {
        derive();
        float L = (float)nom;
        a1 = E * L;
        a2 = E * L * L * L / (12.0f * (1 + nu));
        b1 = E * L;
        b2 = E * L * L / 2.0f;
        b3 = E * L * L * L / 6.0f;
        sqA1 = std::sqrt(a1);
        sqA2xIp = std::sqrt(a2 * L * L / 6.0f);
        sqB1 = std::sqrt(b1);
        sqB2xFMp = std::sqrt(b2 * L / 2.0f);
        sqB3xIp = std::sqrt(b3 * L * L / 6.0f);
    }
    void combine(const VoxMat& m1, const VoxMat& m2)
    {
        nom = 0.5 * (m1.nom + m2.nom);
        rho = 0.5f * (m1.rho + m2.rho);
        cte = 0.5f * (m1.cte + m2.cte);
        muS = 0.5f * (m1.muS + m2.muS);
        muK = 0.5f * (m1.muK + m2.muK);
        zI = 0.5f * (m1.zI + m2.zI);
        zG = 0.5f * (m1.zG + m2.zG);
        zC = 0.5f * (m1.zC + m2.zC);
        ext[0] = ext[1] = ext[2] = 1.0;
        float stressFail = -1.0f, f1 = m1.sigmaFail, f2 = m2.sigmaFail;
        if (f1 == -1.0f) stressFail = f2; else if (f2 == -1.0f) stressFail = f1; else stressFail = f1 < f2 ? f1 : f2;
        if (m1.linear && m2.linear) setLinear(2.0f * m1.E * m2.E / (m1.E + m2.E), stressFail);
        else {
            std::vector<float> ns{0.0f}, nt{0.0f};
            int i1 = 1, i2 = 1;
            while (i1 < (int)m1.strain.size() && i2 < (int)m2.strain.size()) {
                float s = FLT_MAX;
                if (i1 < (int)m1.strain.size()) s = m1.strain[i1];
                if (i2 < (int)m2.strain.size() && m2.strain[i2] < s) s = m2.strain[i2];
                if (s == m1.strain[i1]) i1++;
                if (i2 < (int)m2.strain.size() && s == m2.strain[i2]) i2++;
                float mod1 = m1.modulusAt(s - FLT_EPSILON), mod2 = m2.modulusAt(s - FLT_EPSILON);
                float mod = 2.0f * mod1 * mod2 / (mod1 + mod2);
                int last = (int)ns.size() - 1;
                ns.push_back(s);
                nt.push_back(nt[last] + mod * (s - ns[last]));
            }
            setData((int)ns.size(), ns.data(), nt.data());
            sigmaFail = stressFail;
            epsFail = stressFail == -1.0f ? -1.0f : strainAt(stressFail);
        }
        if (m1.nu == 0 && m2.nu == 0) nu = 0;
        else {
            float tmpEHat = 2 * m1.eHat * m2.eHat / (m1.eHat + m2.eHat);
            float tmpE = E;
            float c2 = (tmpEHat - tmpE) / (2 * tmpEHat) + 0.0625;
            nu = std::sqrt(c2) - 0.25;
        }
        deriveLink();
    }
};

bool apply_desc(VoxMat& m, const vx_material_desc& d, const std::vector<float>& cs, const std::vector<float>& ct)
{
    bool ok;
    switch (d.model) {
    case VX_MODEL_LINEAR: ok = m.setLinear(d.youngs_modulus, d.fail_stress); break;
    case VX_MODEL_BILINEAR: ok = m.setBilinear(d.youngs_modulus, d.plastic_modulus, d.yield_stress, d.fail_stress); break;
    case VX_MODEL_DATA: ok = m.setData((int)cs.size(), cs.data(), ct.data()); break;
    default: m.err = "unknown model"; ok = false;
    }
    if (!ok) return false;
    // clamps of the setters, VX_Material.cpp:472-523
    float rho = d.density; if (rho <= 0) rho = FLT_MIN; m.rho = rho;
    float nu = d.poissons_ratio; if (nu < 0) nu = 0; if (nu >= 0.5) nu = 0.5 - FLT_EPSILON * 2; m.nu = nu;
    m.cte = d.cte;
    m.muS = d.mu_static <= 0 ? 0 : d.mu_static;
    m.muK = d.mu_kinetic <= 0 ? 0 : d.mu_kinetic;
    m.zI = d.zeta_internal <= 0 ? 0 : d.zeta_internal;
    m.zG = d.zeta_global <= 0 ? 0 : d.zeta_global;
    m.zC = d.zeta_collision <= 0 ? 0 : d.zeta_collision;
    for (int a = 0; a < 3; a++) m.ext[a] = d.ext_scale[a] <= 0 ? (double)FLT_MIN : d.ext_scale[a];
    m.derive();
    return true;
}

// ------------------------------------------------------------------ elements
struct Ext {                         // CVX_External, VX_External.h:90-96
    uint8_t dof = 0;
    float f[3] = {0, 0, 0}, m[3] = {0, 0, 0};
    V3 t, r;
    Q4 q;
};

struct Voxel {                       // CVX_Voxel private state, VX_Voxel.h:150-192
    int ix = 0, iy = 0, iz = 0, mat = 0, member = 0;
    int link[6] = {-1, -1, -1, -1, -1, -1};
    int ext = -1;
    bool ghost = false;
    V3 pos, linMom, angMom; Q4 orient;
    float temp = 0;
    bool floorStatic = true;
    int floorOverride = -1;            // CVX_Voxel::enableFloor on this voxel: -1 follows the simulation
    V3f pStrain; bool pInvalid = true;
    V3f lastWatch;
    std::vector<int> colWatch, nearby;
};

struct Link {                        // CVX_Link private state, VX_Link.h:69-107
    int vn = 0, vp = 0, axis = 0, lmat = 0;
    V3 fN, fP, mN, mP;
    float strain = 0, maxStrain = 0, strainOffset = 0, stress = 0, strainRatio = 1;
    bool smallAngle = true, velValid = false;
    V3 pos2, a1v, a2v; Q4 angle1, angle2;
    double restLen = 0;
    float tArea = 0, tStrainSum = 0;
};

struct Collision { int v1, v2; float k, c; V3f force; };   // CVX_Collision, VX_Collision.h:41-44

inline V3 to_axis_x(int axis, V3 v) { if (axis == 1) return {v.y, -v.x, v.z}; if (axis == 2) return {v.z, v.y, -v.x}; return v; }   // VX_Link.h:114
inline Q4 to_axis_x(int axis, Q4 q) { if (axis == 1) return {q.w, q.y, -q.x, q.z}; if (axis == 2) return {q.w, q.z, q.y, -q.x}; return q; }   // VX_Link.h:115
inline void to_axis_original(int axis, V3& v)                                                                                    // VX_Link.h:116
{
    if (axis == 1) { double t = v.y; v.y = v.x; v.x = -t; }
    else if (axis == 2) { double t = v.z; v.z = v.x; v.x = -t; }
}

} // namespace

struct vx_sim {
    double voxSize = 0.001;
    std::vector<vx_material_desc> descs; std::vector<std::vector<float>> dStrain, dStress;
    std::vector<VoxMat> vmats;
    std::vector<LinkMat> lmats;
    std::vector<Ext> exts;
    std::vector<Voxel> vox;
    std::vector<Link> links;
    std::vector<Collision> cols;
    int nMembers = 1;
    float grav = 0, ambient = 0, curTime = 0, prevDt = 0;
    bool floorOn = false, collisions = false, colStale = true, nearbyStale = true;
    float envelope = 0.625f;
    std::string err;
    std::unordered_map<uint64_t, int> cell;

    static uint64_t key(int member, int x, int y, int z)
    { return ((uint64_t)(uint32_t)member << 48) | ((uint64_t)(uint16_t)(int16_t)x << 32) | ((uint64_t)(uint16_t)(int16_t)y << 16) | (uint64_t)(uint16_t)(int16_t)z; }
    int at(int member, int x, int y, int z) const { auto it = cell.find(key(member, x, y, z)); return it == cell.end() ? -1 : it->second; }

    int linkMat(int a, int b)        // CVoxelyze::combinedMaterial, Voxelyze.cpp:626-640
    {
        for (size_t i = 0; i < lmats.size(); i++)
            if ((lmats[i].ma == a && lmats[i].mb == b) || (lmats[i].ma == b && lmats[i].mb == a)) return (int)i;
        LinkMat lm; lm.ma = a; lm.mb = b; lm.combine(vmats[a], vmats[b]);
        lmats.push_back(lm); return (int)lmats.size() - 1;
    }

    // ---- voxel helpers
    double baseSize(const Voxel& v, int axis) const { const VoxMat& m = vmats[v.mat]; return m.size(axis) * (1 + v.temp * m.cte); }   // VX_Voxel.h:89
    double baseSizeAvg(const Voxel& v) const { return (baseSize(v, 0) + baseSize(v, 1) + baseSize(v, 2)) / 3.0f; }                     // VX_Voxel.h:90
    V3 velocity(const Voxel& v) const { return scale(vmats[v.mat].massInv, v.linMom); }                                                // VX_Voxel.h:98
    V3 angVelocity(const Voxel& v) const { return scale(vmats[v.mat].inertiaInv, v.angMom); }                                          // VX_Voxel.h:100
    float dampingMultiplier(const Voxel& v) const { const VoxMat& m = vmats[v.mat]; return 2 * m.sqrtMass * m.zI / prevDt; }            // VX_Voxel.h:130
    float floorPenetration(const Voxel& v) const { return (float)(baseSizeAvg(v) / 2 - vmats[v.mat].nom / 2 - v.pos.z); }              // VX_Voxel.h:122
    bool isSurface(const Voxel& v) const { for (int i = 0; i < 6; i++) if (v.link[i] < 0) return true; return false; }                 // VX_Voxel.cpp:376-381

    // CVX_Link::axialStrain(bool positiveEnd), VX_Link.cpp:121-124
    float axialStrainEnd(const Link& l, bool positiveEnd) const
    { return positiveEnd ? 2.0f * l.strain * l.strainRatio / (1.0f + l.strainRatio) : 2.0f * l.strain / (1.0f + l.strainRatio); }

    // CVX_Voxel::strain(bool), VX_Voxel.cpp:300-334
    V3f voxelStrain(const Voxel& v, bool poissons) const
    {
        float r[3] = {0, 0, 0}; int nb[3] = {0, 0, 0}; bool tension[3] = {false, false, false};
        for (int i = 0; i < 6; i++) if (v.link[i] >= 0) { int ax = i / 2; r[ax] += axialStrainEnd(links[v.link[i]], (i % 2) == 1); nb[ax]++; }
        for (int i = 0; i < 3; i++) {
            if (nb[i] == 2) r[i] *= 0.5f;
            if (poissons) {
                const Ext* e = v.ext >= 0 ? &exts[v.ext] : nullptr;
                tension[i] = ((nb[i] == 2) || (e && (nb[i] == 1 && ((e->dof & (1 << i)) || e->f[i] != 0))));
            }
        }
        if (poissons && !(tension[0] && tension[1] && tension[2])) {
            float addv = 0;
            for (int i = 0; i < 3; i++) if (tension[i]) addv += r[i];
            float value = std::pow(1.0f + addv, -vmats[v.mat].nu) - 1.0f;
            for (int i = 0; i < 3; i++) if (!tension[i]) r[i] = value;
        }
        return {r[0], r[1], r[2]};
    }
    V3f poissonsStrain(Voxel& v) { if (v.pInvalid) { v.pStrain = voxelStrain(v, true); v.pInvalid = false; } return v.pStrain; }         // VX_Voxel.cpp:336-343
    float transverseStrainSum(Voxel& v, int axis)                                                                                       // VX_Voxel.cpp:346-359
    {
        if (vmats[v.mat].nu == 0) return 0;
        V3f p = poissonsStrain(v);
        return axis == 0 ? p.y + p.z : axis == 1 ? p.x + p.z : p.x + p.y;
    }
    float transverseArea(Voxel& v, int axis)                                                                                            // VX_Voxel.cpp:361-374
    {
        float size = (float)vmats[v.mat].nom;
        if (vmats[v.mat].nu == 0) return size * size;
        V3f pf = poissonsStrain(v); double px = pf.x, py = pf.y, pz = pf.z;
        if (axis == 0) return (float)(size * size * (1 + py) * (1 + pz));
        if (axis == 1) return (float)(size * size * (1 + px) * (1 + pz));
        return (float)(size * size * (1 + px) * (1 + py));
    }

    // ---- link
    void updateRestLength(Link& l) { l.restLen = 0.5 * (baseSize(vox[l.vn], l.axis) + baseSize(vox[l.vp], l.axis)); }                    // VX_Link.cpp:137-140
    void updateTransverseInfo(Link& l)                                                                                                  // VX_Link.cpp:142-147
    {
        l.tArea = 0.5f * (transverseArea(vox[l.vn], l.axis) + transverseArea(vox[l.vp], l.axis));
        l.tStrainSum = 0.5f * (transverseStrainSum(vox[l.vn], l.axis) + transverseStrainSum(vox[l.vp], l.axis));
    }
    void resetLink(Link& l)                                                                                                             // VX_Link.cpp:61-75
    {
        l.pos2 = l.a1v = l.a2v = V3{}; l.angle1 = l.angle2 = Q4{};
        l.fN = l.fP = l.mN = l.mP = V3{};
        l.strain = l.maxStrain = l.strainOffset = l.stress = 0.0f;
        l.strainRatio = vmats[vox[l.vp].mat].E / vmats[vox[l.vn].mat].E;
        l.smallAngle = true; l.velValid = false;
        updateRestLength(l); updateTransverseInfo(l);
    }
    // CVX_Link::orientLink, VX_Link.cpp:77-119
    void orientLink(Link& l)
    {
        const Voxel& vn = vox[l.vn]; const Voxel& vp = vox[l.vp];
        l.pos2 = to_axis_x(l.axis, sub(vp.pos, vn.pos));
        l.angle1 = to_axis_x(l.axis, vn.orient);
        l.angle2 = to_axis_x(l.axis, vp.orient);
        Q4 totalRot = conj(l.angle1);
        l.pos2 = rotate(totalRot, l.pos2);
        l.angle2 = qmul(totalRot, l.angle2);
        l.angle1 = Q4{};
        float smallTurn = (float)((std::fabs(l.pos2.z) + std::fabs(l.pos2.y)) / l.pos2.x);
        float extendPerc = (float)(std::fabs(1 - l.pos2.x / l.restLen));
        const float HYST = 1.2f, BEND = 0.05f, EXTP = 0.50f;                         // VX_Link.cpp:21-23
        if (!l.smallAngle && smallTurn < BEND && extendPerc < EXTP) { l.smallAngle = true; l.velValid = false; }
        else if (l.smallAngle && (smallTurn > HYST * BEND || extendPerc > HYST * EXTP)) { l.smallAngle = false; l.velValid = false; }
        if (l.smallAngle) l.pos2.x -= l.restLen;
        else {
            l.angle1 = from_angle_to_pos_x(l.pos2);
            totalRot = qmul(l.angle1, totalRot);
            l.angle2 = qmul(l.angle1, l.angle2);
            l.pos2 = V3{std::sqrt(len2(l.pos2)) - l.restLen, 0, 0};
        }
        l.a1v = to_rotation_vector(l.angle1);
        l.a2v = to_rotation_vector(l.angle2);
    }
    // CVX_Link::updateStrain, VX_Link.cpp:220-249
    float updateStrain(Link& l, float axial)
    {
        const LinkMat& m = lmats[l.lmat];
        l.strain = axial;
        if (m.linear) {
            if (axial > l.maxStrain) l.maxStrain = axial;
            return m.stressAt(axial, l.tStrainSum);
        }
        float ret;
        if (axial > l.maxStrain) {
            l.maxStrain = axial;
            ret = m.stressAt(axial, l.tStrainSum);
            if (m.nu != 0.0f) l.strainOffset = l.maxStrain - m.stressAt(axial) / (m.eHat * (1 - m.nu));
            else l.strainOffset = l.maxStrain - ret / m.E;
        } else {
            float rel = axial - l.strainOffset;
            if (m.nu != 0.0f) ret = m.stressAt(rel, l.tStrainSum, true);
            else ret = m.E * rel;
        }
        return ret;
    }
    // CVX_Link::updateForces, VX_Link.cpp:149-217
    void updateForces(Link& l)
    {
        const LinkMat& m = lmats[l.lmat];
        V3 oldPos2 = l.pos2, oldA1 = l.a1v, oldA2 = l.a2v;
        orientLink(l);
        V3 dPos2 = scale(0.5, sub(l.pos2, oldPos2));
        V3 dA1 = scale(0.5, sub(l.a1v, oldA1));
        V3 dA2 = scale(0.5, sub(l.a2v, oldA2));
        if (!(m.nu == 0.0f) || l.tStrainSum != 0) updateTransverseInfo(l);
        l.stress = updateStrain(l, (float)(l.pos2.x / l.restLen));
        if (m.isFailed(l.maxStrain)) { l.fN = l.fP = l.mN = l.mP = V3{}; return; }
        float b1 = m.b1, b2 = m.b2, b3 = m.b3, a2 = m.a2;
        const V3& p = l.pos2; const V3& a = l.a1v; const V3& b = l.a2v;
        l.fN = V3{ l.stress * l.tArea,
                   b1 * p.y - b2 * (a.z + b.z),
                   b1 * p.z + b2 * (a.y + b.y) };
        l.fP = neg(l.fN);
        l.mN = V3{ a2 * (b.x - a.x), -b2 * p.z - b3 * (2 * a.y + b.y), b2 * p.y - b3 * (2 * a.z + b.z) };
        l.mP = V3{ a2 * (a.x - b.x), -b2 * p.z - b3 * (a.y + 2 * b.y), b2 * p.y - b3 * (a.z + 2 * b.z) };
        if (l.velValid) {
            float sqA1 = m.sqA1, sqA2xIp = m.sqA2xIp, sqB1 = m.sqB1, sqB2xFMp = m.sqB2xFMp, sqB3xIp = m.sqB3xIp;
            V3 posCalc{ sqA1 * dPos2.x,
                        sqB1 * dPos2.y - sqB2xFMp * (dA1.z + dA2.z),
                        sqB1 * dPos2.z + sqB2xFMp * (dA1.y + dA2.y) };
            l.fN = add(l.fN, scale(dampingMultiplier(vox[l.vn]), posCalc));
            l.fP = sub(l.fP, scale(dampingMultiplier(vox[l.vp]), posCalc));
            V3 mn{ -sqA2xIp * (dA2.x - dA1.x),
                   sqB2xFMp * dPos2.z + sqB3xIp * (2 * dA1.y + dA2.y),
                   -sqB2xFMp * dPos2.y + sqB3xIp * (2 * dA1.z + dA2.z) };
            V3 mp{ sqA2xIp * (dA2.x - dA1.x),
                   sqB2xFMp * dPos2.z + sqB3xIp * (dA1.y + 2 * dA2.y),
                   -sqB2xFMp * dPos2.y + sqB3xIp * (dA1.z + 2 * dA2.z) };
            l.mN = sub(l.mN, scale(0.5 * dampingMultiplier(vox[l.vn]), mn));
            l.mP = sub(l.mP, scale(0.5 * dampingMultiplier(vox[l.vp]), mp));
        } else l.velValid = true;
        if (!l.smallAngle) { l.fN = rotate_inv(l.angle1, l.fN); l.mN = rotate_inv(l.angle1, l.mN); }
        l.fP = rotate_inv(l.angle2, l.fP);
        l.mP = rotate_inv(l.angle2, l.mP);
        to_axis_original(l.axis, l.fN); to_axis_original(l.axis, l.fP);
        to_axis_original(l.axis, l.mN); to_axis_original(l.axis, l.mP);
    }
    // CVX_Link::axialStiffness, VX_Link.cpp:259-267
    float axialStiffness(Link& l)
    {
        const LinkMat& m = lmats[l.lmat];
        if (m.nu == 0.0f) return m.a1;
        updateRestLength(l); updateTransverseInfo(l);
        return (float)(m.eHat * l.tArea / ((l.strain + 1) * l.restLen));
    }

    // ---- voxel dynamics
    V3f contactForce(const Collision& c, int v) const                                // VX_Collision.cpp:34-39
    { if (v == c.v1) return c.force; if (v == c.v2) return {-c.force.x, -c.force.y, -c.force.z}; return {}; }

    V3 voxelForce(const Voxel& v, int self)                                            // VX_Voxel.cpp:234-256
    {
        const VoxMat& m = vmats[v.mat];
        V3 tot;
        for (int i = 0; i < 6; i++) if (v.link[i] >= 0) { const Link& l = links[v.link[i]]; tot = add(tot, (i % 2) == 1 ? l.fP : l.fN); }
        tot = rotate(v.orient, tot);
        if (v.ext >= 0) { const Ext& e = exts[v.ext]; tot.x += e.f[0]; tot.y += e.f[1]; tot.z += e.f[2]; }
        V3 vel = velocity(v); float c = m.globalDampT();
        tot = sub(tot, scale(c, vel));
        tot.z += -m.mass * 9.80665f * grav;                                          // VX_MaterialVoxel.h:57
        if (collisions) for (int ci : v.colWatch) { V3f f = contactForce(cols[ci], self); tot.x -= f.x; tot.y -= f.y; tot.z -= f.z; }
        return tot;
    }
    V3 voxelMoment(const Voxel& v)                                                     // VX_Voxel.cpp:258-271
    {
        const VoxMat& m = vmats[v.mat];
        V3 tot;
        for (int i = 0; i < 6; i++) if (v.link[i] >= 0) { const Link& l = links[v.link[i]]; tot = add(tot, (i % 2) == 1 ? l.mP : l.mN); }
        tot = rotate(v.orient, tot);
        if (v.ext >= 0) { const Ext& e = exts[v.ext]; tot.x += e.m[0]; tot.y += e.m[1]; tot.z += e.m[2]; }
        tot = sub(tot, scale(m.globalDampR(), angVelocity(v)));
        return tot;
    }
    void floorForce(Voxel& v, V3& F)                                                   // VX_Voxel.cpp:274-298
    {
        const VoxMat& m = vmats[v.mat];
        float pen = floorPenetration(v);
        if (pen >= 0) {
            V3 vel = velocity(v); V3 hv{vel.x, vel.y, 0};
            float normalForce = m.penetrationStiffness() * pen;
            F.z += normalForce - m.collDampT() * vel.z;
            if (v.floorStatic) {
                float surf = (float)(F.x * F.x + F.y * F.y);
                float fric = (m.muS * normalForce) * (m.muS * normalForce);
                if (surf > fric) v.floorStatic = false;
            } else {
                double l = std::sqrt(hv.x * hv.x + hv.y * hv.y + hv.z * hv.z);       // Vec3D.h:95 Normalized()
                V3 n = hv; if (l > 0) { double inv = 1.0 / l; n = scale(inv, hv); }
                float mk = m.muK * normalForce;
                F = sub(F, scale(mk, n));
            }
        } else v.floorStatic = false;
    }
    void timeStep(Voxel& v, int self, float dt)                                        // VX_Voxel.cpp:162-232
    {
        if (dt == 0.0f) return;
        const VoxMat& m = vmats[v.mat];
        const Ext* e = v.ext >= 0 ? &exts[v.ext] : nullptr;
        double s = m.nom;
        if (e && (e->dof & 0x3F) == 0x3F) {
            v.pos = add(V3{v.ix * s, v.iy * s, v.iz * s}, e->t);
            v.orient = e->q; v.linMom = v.angMom = V3{};
            return;
        }
        V3 cur = voxelForce(v, self), fric = cur;
        const bool floorOn = v.floorOverride < 0 ? this->floorOn : v.floorOverride != 0;     // include/VX_Voxel.h:119-120
        if (floorOn) floorForce(v, cur);
        fric = sub(cur, fric);
        v.linMom = add(v.linMom, scale(dt, cur));
        V3 tr = scale(dt * m.massInv, v.linMom);
        if (floorOn && floorPenetration(v) >= 0) {
            double work = fric.x * tr.x + fric.y * tr.y;
            double hKe = 0.5 * m.massInv * (v.linMom.x * v.linMom.x + v.linMom.y * v.linMom.y);
            if (hKe + work <= 0) v.floorStatic = true;
            if (v.floorStatic) { v.linMom.x = v.linMom.y = 0; tr.x = tr.y = 0; }
        } else v.floorStatic = false;
        v.pos = add(v.pos, tr);
        V3 mom = voxelMoment(v);
        v.angMom = add(v.angMom, scale(dt, mom));
        v.orient = qmul(from_rotation_vector(scale(dt * m.inertiaInv, v.angMom)), v.orient);
        if (e) {
            if (e->dof & VX_DOF_TX) { v.pos.x = v.ix * s + e->t.x; v.linMom.x = 0; }
            if (e->dof & VX_DOF_TY) { v.pos.y = v.iy * s + e->t.y; v.linMom.y = 0; }
            if (e->dof & VX_DOF_TZ) { v.pos.z = v.iz * s + e->t.z; v.linMom.z = 0; }
            if (e->dof & (VX_DOF_RX | VX_DOF_RY | VX_DOF_RZ)) {
                if ((e->dof & 0x38) == 0x38) { v.orient = e->q; v.angMom = V3{}; }
                else {
                    V3 rv = to_rotation_vector(v.orient);
                    if (e->dof & VX_DOF_RX) { rv.x = 0; v.angMom.x = 0; }
                    if (e->dof & VX_DOF_RY) { rv.y = 0; v.angMom.y = 0; }
                    if (e->dof & VX_DOF_RZ) { rv.z = 0; v.angMom.z = 0; }
                    v.orient = from_rotation_vector(rv);
                }
            }
        }
        v.pInvalid = true;
    }

    // ---- collisions
    void generateNearby(int vi, int depth)                                             // VX_Voxel.cpp:395-418
    {
        std::vector<int> all{vi}; size_t cur = 0;
        for (int k = 0; k < depth; k++) {
            size_t passEnd = all.size();
            while (cur != passEnd) {
                const Voxel& pv = vox[all[cur++]];
                for (int i = 0; i < 6; i++) {
                    if (pv.link[i] < 0) continue;
                    const Link& l = links[pv.link[i]];
                    int other = (&vox[l.vp] == &pv) ? l.vn : l.vp;
                    if (std::find(all.begin(), all.end(), other) == all.end()) all.push_back(other);
                }
            }
        }
        Voxel& v = vox[vi]; v.nearby.clear();
        for (int o : all) if (o != vi && isSurface(vox[o])) v.nearby.push_back(o);
    }
    void clearCollisions() { cols.clear(); for (auto& v : vox) v.colWatch.clear(); }   // Voxelyze.cpp:712-722
    void regenerateCollisions(float threshSq)                                          // Voxelyze.cpp:725-750
    {
        clearCollisions();
        int n = (int)vox.size();
        for (int i = 0; i < n; i++) {
            Voxel& a = vox[i];
            if (!isSurface(a)) continue;
            a.lastWatch = {(float)a.pos.x, (float)a.pos.y, (float)a.pos.z};
            for (int j = i + 1; j < n; j++) {
                Voxel& b = vox[j];
                if (b.member != a.member) continue;      // ensemble members are separate CVoxelyze objects
                if (!isSurface(b) || len2(sub(a.pos, b.pos)) > threshSq ||
                    std::find(a.nearby.begin(), a.nearby.end(), j) != a.nearby.end()) continue;
                Collision c; c.v1 = i; c.v2 = j;                                       // VX_Collision.cpp:17-23
                const VoxMat& m1 = vmats[a.mat]; const VoxMat& m2 = vmats[b.mat];
                c.k = 2.0f / (1.0f / m1.penetrationStiffness() + 1.0f / m2.penetrationStiffness());
                c.c = 0.5f * (m1.collDampT() + m2.collDampT());
                cols.push_back(c);
                a.colWatch.push_back((int)cols.size() - 1); b.colWatch.push_back((int)cols.size() - 1);
            }
        }
        colStale = false;
    }
    void updateContactForce(Collision& c)                                              // VX_Collision.cpp:42-56
    {
        const Voxel& a = vox[c.v1]; const Voxel& b = vox[c.v2];
        V3 d = sub(b.pos, a.pos);
        V3f off{(float)d.x, (float)d.y, (float)d.z};
        float nomDist = (float)((baseSizeAvg(a) + baseSizeAvg(b)) * envelope);
        float length = std::sqrt(off.x * off.x + off.y * off.y + off.z * off.z);
        float rel = nomDist - length;
        if (rel > 0) {
            V3f unit = off;
            if (length > 0) { float inv = 1.0f / length; unit = {inv * off.x, inv * off.y, inv * off.z}; }
            V3 ud{unit.x, unit.y, unit.z};
            V3 va = velocity(a), vb = velocity(b);
            float relVel = (float)((va.x * ud.x + va.y * ud.y + va.z * ud.z) - (vb.x * ud.x + vb.y * ud.y + vb.z * ud.z));
            float mag = c.k * rel + c.c * relVel;
            c.force = {mag * unit.x, mag * unit.y, mag * unit.z};
        } else c.force = {};
    }
    void updateCollisions()                                                            // Voxelyze.cpp:670-710
    {
        float watchRadiusVx = 2 * 0.75f + 1.0f;
        float watchRadiusMm = (float)(voxSize * watchRadiusVx);
        float recalcDist = (float)(voxSize * 1.0f / 2);
        if (nearbyStale) {
            for (int i = 0; i < (int)vox.size(); i++) generateNearby(i, (int)(watchRadiusVx * 2));
            nearbyStale = false; colStale = true;
        }
        for (auto& v : vox) {
            if (!isSurface(v)) continue;
            V3 d{v.pos.x - v.lastWatch.x, v.pos.y - v.lastWatch.y, v.pos.z - v.lastWatch.z};
            if (len2(d) > recalcDist * recalcDist) colStale = true;
        }
        if (colStale) regenerateCollisions(watchRadiusMm * watchRadiusMm);
        for (auto& c : cols) updateContactForce(c);
    }

    // CVoxelyze::doTimeStep, Voxelyze.cpp:251-284
    bool doTimeStep(float dt)
    {
        if (dt == 0) return true;
        else if (dt < 0) dt = recommendedTimeStep();
        bool diverged = false;
        for (auto& l : links) { updateForces(l); if (l.strain > 100) diverged = true; }
        if (diverged) return false;
        if (collisions) updateCollisions();
        for (int i = 0; i < (int)vox.size(); i++) if (!vox[i].ghost) timeStep(vox[i], i, dt);
        prevDt = dt;                       // CVX_Voxel::previousDt, VX_Voxel.cpp:164 (same value for every voxel)
        curTime += dt;
        return true;
    }
    // CVoxelyze::recommendedTimeStep, Voxelyze.cpp:286-311
    float recommendedTimeStep()
    {
        float maxFreq2 = 0.0f;
        for (auto& l : links) {
            float m1 = vmats[vox[l.vn].mat].mass, m2 = vmats[vox[l.vp].mat].mass;
            float f2 = axialStiffness(l) / (m1 < m2 ? m1 : m2);
            if (f2 > maxFreq2) maxFreq2 = f2;
        }
        if (maxFreq2 <= 0.0f)
            for (auto& v : vox) { const VoxMat& m = vmats[v.mat]; float f2 = m.E * m.nom / m.mass; if (f2 > maxFreq2) maxFreq2 = f2; }
        if (maxFreq2 <= 0.0f) return 0.0f;
        return 1.0f / (6.283185f * std::sqrt(maxFreq2));
    }
    // CVoxelyze::resetTime, Voxelyze.cpp:313-321 + CVX_Voxel::reset VX_Voxel.cpp:47-56
    void resetTime()
    {
        curTime = 0.0f; colStale = true; nearbyStale = true;
        for (auto& v : vox) {
            double s = vmats[v.mat].nom;
            v.pos = V3{v.ix * s, v.iy * s, v.iz * s}; v.orient = Q4{}; v.linMom = v.angMom = V3{};
            v.floorStatic = true; v.temp = 0.0f; v.pInvalid = true;
        }
        prevDt = 0.0f;
        for (auto& l : links) resetLink(l);
    }
    void setTemperature(Voxel& v, float t) { v.temp = t; for (int i = 0; i < 6; i++) if (v.link[i] >= 0) updateRestLength(links[v.link[i]]); }   // VX_Voxel.cpp:108-114
};

static int fail(vx_sim* s, int code, const std::string& msg) { if (s) s->err = msg; return code; }

static int derive_materials(vx_sim* s)
{
    std::vector<VoxMat> nv(s->descs.size());
    for (size_t i = 0; i < nv.size(); i++) {
        nv[i].nom = s->voxSize;
        if (!apply_desc(nv[i], s->descs[i], s->dStrain[i], s->dStress[i])) return fail(s, VX_ERR_MATERIAL, nv[i].err);
    }
    s->vmats.swap(nv);
    for (auto& lm : s->lmats) lm.combine(s->vmats[lm.ma], s->vmats[lm.mb]);    // dependentMaterials->updateAll, VX_Material.cpp:529
    return VX_OK;
}

extern "C" {

int vx_abi_version(void) { return VX_ABI_VERSION; }
const char* vx_backend(void) { return "oracle-port"; }
int vx_create(double voxel_size, int, vx_sim** out) { if (!out) return VX_ERR_ARG; vx_sim* s = new vx_sim; s->voxSize = voxel_size; *out = s; return VX_OK; }
void vx_destroy(vx_sim* s) { delete s; }
const char* vx_last_error(const vx_sim* s) { return s ? s->err.c_str() : "null handle"; }

int vx_set_materials(vx_sim* s, int n, const vx_material_desc* d)
{
    if (!s || n < 0 || (n && !d)) return VX_ERR_ARG;
    if (!s->vox.empty() && n != (int)s->descs.size()) return fail(s, VX_ERR_ARG, "material count changed after voxels were set");
    auto od = s->descs; auto os = s->dStrain; auto ot = s->dStress; auto ov = s->vmats;
    s->descs.assign(d, d + n); s->dStrain.assign(n, {}); s->dStress.assign(n, {});
    for (int i = 0; i < n; i++) {
        if (d[i].model == VX_MODEL_DATA) {
            if (d[i].n_points < 0 || !d[i].strain || !d[i].stress) return VX_ERR_ARG;
            s->dStrain[i].assign(d[i].strain, d[i].strain + d[i].n_points);
            s->dStress[i].assign(d[i].stress, d[i].stress + d[i].n_points);
        }
        s->descs[i].strain = s->descs[i].stress = nullptr;
    }
    int rc = derive_materials(s);
    if (rc != VX_OK) { s->descs = od; s->dStrain = os; s->dStress = ot; s->vmats = ov; return rc; }
    // a material change refreshes what the reference refreshes lazily (documented deviation:
    // rest lengths follow ext_scale immediately)
    for (auto& l : s->links) { s->updateRestLength(l); }
    return VX_OK;
}

int vx_get_voxmat(const vx_sim* s, int i, vx_voxmat_row* o)
{
    if (!s || !o || i < 0 || i >= (int)s->vmats.size()) return VX_ERR_ARG;
    const VoxMat& m = s->vmats[i];
    o->nom_size = m.nom; for (int a = 0; a < 3; a++) o->size[a] = m.size(a);
    o->E = m.E; o->nu = m.nu; o->rho = m.rho; o->cte = m.cte; o->mu_static = m.muS; o->mu_kinetic = m.muK;
    o->zeta_internal = m.zI; o->zeta_global = m.zG; o->zeta_collision = m.zC; o->e_hat = m.eHat;
    o->mass = m.mass; o->mass_inv = m.massInv; o->sqrt_mass = m.sqrtMass; o->first_moment = m.firstMoment;
    o->moment_inertia = m.inertia; o->moment_inertia_inv = m.inertiaInv; o->two_sq_m_e_s = m.sqMES; o->two_sq_i_e_s3 = m.sqIES3;
    o->eps_yield = m.epsYield; o->eps_fail = m.epsFail; o->sigma_yield = m.sigmaYield; o->sigma_fail = m.sigmaFail;
    o->linear = m.linear; o->n_curve = (int)m.strain.size();
    return VX_OK;
}
static const LinkMat* find_linkmat(vx_sim* s, int a, int b, LinkMat& scratch)
{
    if (!s || a < 0 || b < 0 || a >= (int)s->vmats.size() || b >= (int)s->vmats.size()) return nullptr;
    scratch.ma = a; scratch.mb = b; scratch.combine(s->vmats[a], s->vmats[b]);
    return &scratch;
}
int vx_get_linkmat(vx_sim* s, int a, int b, vx_linkmat_row* o)
{
    LinkMat tmp; const LinkMat* m = find_linkmat(s, a, b, tmp);
    if (!m || !o) return VX_ERR_ARG;
    o->mat_a = std::min(a, b); o->mat_b = std::max(a, b); o->linear = m->linear; o->n_curve = (int)m->strain.size();
    o->E = m->E; o->nu = m->nu; o->e_hat = m->eHat; o->eps_yield = m->epsYield; o->eps_fail = m->epsFail;
    o->sigma_yield = m->sigmaYield; o->sigma_fail = m->sigmaFail;
    o->a1 = m->a1; o->a2 = m->a2; o->b1 = m->b1; o->b2 = m->b2; o->b3 = m->b3;
    o->sq_a1 = m->sqA1; o->sq_a2_ip = m->sqA2xIp; o->sq_b1 = m->sqB1; o->sq_b2_fmp = m->sqB2xFMp; o->sq_b3_ip = m->sqB3xIp;
    return VX_OK;
}
int vx_get_linkmat_curve(vx_sim* s, int a, int b, float* st, float* ss, int cap)
{
    LinkMat tmp; const LinkMat* m = find_linkmat(s, a, b, tmp);
    if (!m) return VX_ERR_ARG;
    int n = (int)m->strain.size();
    if (st && ss) for (int i = 0; i < n && i < cap; i++) { st[i] = m->strain[i]; ss[i] = m->stress[i]; }
    return n;
}

int vx_set_voxels(vx_sim* s, int n, const int32_t* ijk, const uint16_t* mat, const int32_t* sim_id, const uint32_t* flags)
{
    if (!s || n < 0 || (n && (!ijk || !mat))) return VX_ERR_ARG;
    s->vox.clear(); s->links.clear(); s->cols.clear(); s->cell.clear(); s->lmats.clear(); s->exts.clear();
    s->curTime = 0; s->prevDt = 0; s->colStale = s->nearbyStale = true;
    s->vox.reserve(n);
    int maxMember = 0;
    static const int dx[6] = {1, -1, 0, 0, 0, 0}, dy[6] = {0, 0, 1, -1, 0, 0}, dz[6] = {0, 0, 0, 0, 1, -1};
    for (int i = 0; i < n; i++) {
        if (mat[i] >= s->vmats.size()) return fail(s, VX_ERR_ARG, "material index out of range");
        Voxel v; v.ix = ijk[3 * i]; v.iy = ijk[3 * i + 1]; v.iz = ijk[3 * i + 2]; v.mat = mat[i];
        for (int a = 0; a < 3; a++) if (ijk[3 * i + a] < -32768 || ijk[3 * i + a] > 32767) return fail(s, VX_ERR_ARG, "lattice index does not fit a short");
        v.member = sim_id ? sim_id[i] : 0; if (v.member < 0 || v.member > 65535) return fail(s, VX_ERR_ARG, "bad member id");
        maxMember = std::max(maxMember, v.member);
        v.ghost = flags && (flags[i] & VX_VF_GHOST);
        uint64_t k = vx_sim::key(v.member, v.ix, v.iy, v.iz);
        if (s->cell.count(k)) return fail(s, VX_ERR_TOPOLOGY, "duplicate voxel");
        double sz = s->vmats[v.mat].nom;
        v.pos = V3{v.ix * sz, v.iy * sz, v.iz * sz};                                   // Voxelyze.cpp:447
        v.temp = s->ambient;                                                          // Voxelyze.cpp:449
        s->cell[k] = i; s->vox.push_back(v);
        for (int d = 0; d < 6; d++) {                                                 // Voxelyze.cpp:453-455, 508-539
            int o = s->at(v.member, v.ix + dx[d], v.iy + dy[d], v.iz + dz[d]);
            if (o < 0) continue;
            if (s->vox[i].ghost && s->vox[o].ghost) continue;                         // halo-halo links are never needed
            Link l; l.axis = d / 2;
            bool thisIsNeg = (d % 2) == 0;                                            // VX_Link.cpp:31-53
            l.vn = thisIsNeg ? i : o; l.vp = thisIsNeg ? o : i;
            l.lmat = s->linkMat(s->vox[i].mat, s->vox[o].mat);
            s->links.push_back(l);
            int li = (int)s->links.size() - 1;
            s->resetLink(s->links[li]);               // CVX_Link ctor runs before addLinkInfo (Voxelyze.cpp:525,536)
            s->vox[i].link[d] = li; s->vox[o].link[d ^ 1] = li;
        }
    }
    s->nMembers = maxMember + 1;
    return VX_OK;
}
int vx_voxel_count(const vx_sim* s) { return s ? (int)s->vox.size() : 0; }
int vx_link_count(const vx_sim* s) { return s ? (int)s->links.size() : 0; }
int vx_get_links(const vx_sim* s, int32_t* vn, int32_t* vp, uint8_t* ax)
{
    if (!s) return VX_ERR_ARG;
    for (size_t i = 0; i < s->links.size(); i++) { if (vn) vn[i] = s->links[i].vn; if (vp) vp[i] = s->links[i].vp; if (ax) ax[i] = (uint8_t)s->links[i].axis; }
    return VX_OK;
}

int vx_set_externals(vx_sim* s, int n, const int32_t* voxel, const uint8_t* dof, const float* f, const float* m, const double* t, const double* r)
{
    if (!s || n < 0 || (n && (!voxel || !dof))) return VX_ERR_ARG;
    for (int i = 0; i < n; i++) if (voxel[i] < 0 || voxel[i] >= (int)s->vox.size()) return fail(s, VX_ERR_ARG, "external voxel index out of range");
    for (auto& v : s->vox) v.ext = -1;
    s->exts.assign(n, Ext{});
    for (int i = 0; i < n; i++) {
        Ext& e = s->exts[i]; e.dof = dof[i] & 0x3F;
        for (int a = 0; a < 3; a++) { if (f) e.f[a] = f[3 * i + a]; if (m) e.m[a] = m[3 * i + a]; }
        if (t) e.t = V3{t[3 * i], t[3 * i + 1], t[3 * i + 2]};
        if (r) e.r = V3{r[3 * i], r[3 * i + 1], r[3 * i + 2]};
        if (e.r.x != 0 || e.r.y != 0 || e.r.z != 0) e.q = from_rotation_vector(e.r);   // VX_External.cpp:100-109
        s->vox[voxel[i]].ext = i;
    }
    return VX_OK;
}
int vx_set_gravity(vx_sim* s, float g) { if (!s) return VX_ERR_ARG; s->grav = g; return VX_OK; }
int vx_enable_floor(vx_sim* s, int e) { if (!s) return VX_ERR_ARG; s->floorOn = e != 0; return VX_OK; }
int vx_enable_collisions(vx_sim* s, int e)                                             // Voxelyze.cpp:612-622
{
    if (!s) return VX_ERR_ARG;
    if (s->collisions == (e != 0)) return VX_OK;
    s->collisions = e != 0; if (!s->collisions) s->clearCollisions(); s->colStale = true; return VX_OK;
}
int vx_set_collision_envelope(vx_sim* s, float r) { if (!s) return VX_ERR_ARG; s->envelope = r; return VX_OK; }
int vx_set_temperature_all(vx_sim* s, float t) { if (!s) return VX_ERR_ARG; s->ambient = t; for (auto& v : s->vox) s->setTemperature(v, t); return VX_OK; }
int vx_set_temperature_members(vx_sim* s, int n, const float* t)
{
    if (!s || !t || n != s->nMembers) return VX_ERR_ARG;
    for (auto& v : s->vox) s->setTemperature(v, t[v.member]);
    return VX_OK;
}
int vx_set_temperature(vx_sim* s, int n, const float* t)
{
    if (!s || !t || n != (int)s->vox.size()) return VX_ERR_ARG;
    for (int i = 0; i < n; i++) s->setTemperature(s->vox[i], t[i]);
    return VX_OK;
}

int vx_step(vx_sim* s, float dt, int n_steps, int* diverged_step)
{
    if (!s || n_steps < 0) return VX_ERR_ARG;
    for (int k = 0; k < n_steps; k++) if (!s->doTimeStep(dt)) { if (diverged_step) *diverged_step = k; return VX_DIVERGED; }
    return VX_OK;
}
int vx_recommended_dt(vx_sim* s, float* dt) { if (!s || !dt) return VX_ERR_ARG; *dt = s->recommendedTimeStep(); return VX_OK; }
int vx_reset(vx_sim* s) { if (!s) return VX_ERR_ARG; s->resetTime(); return VX_OK; }
float vx_time(const vx_sim* s) { return s ? s->curTime : 0.f; }
int vx_set_clock(vx_sim* s, float time, float previous_dt) { if (!s || !(time >= 0.f) || !(previous_dt >= 0.f)) return VX_ERR_ARG; s->curTime = time; s->prevDt = previous_dt; return VX_OK; }

int vx_download(vx_sim* s, int field, int first, int count, void* dst)
{
    if (!s || !dst || first < 0 || count < 0) return VX_ERR_ARG;
    bool isLink = field >= 16;
    if (first + count > (isLink ? (int)s->links.size() : (int)s->vox.size())) return VX_ERR_ARG;
    double* d = (double*)dst; float* f = (float*)dst; uint32_t* u = (uint32_t*)dst;
    auto put = [&](int k, const V3& v) { d[3 * k] = v.x; d[3 * k + 1] = v.y; d[3 * k + 2] = v.z; };
    for (int k = 0; k < count; k++) {
        if (!isLink) {
            const Voxel& v = s->vox[first + k];
            switch (field) {
            case VX_F_POS: put(k, v.pos); break;
            case VX_F_ORIENT: d[4 * k] = v.orient.w; d[4 * k + 1] = v.orient.x; d[4 * k + 2] = v.orient.y; d[4 * k + 3] = v.orient.z; break;
            case VX_F_LINMOM: put(k, v.linMom); break;
            case VX_F_ANGMOM: put(k, v.angMom); break;
            case VX_F_TEMP: f[k] = v.temp; break;
            case VX_F_VOXFLAGS: u[k] = (v.floorStatic ? VX_VF_STATIC_FRICTION : 0) | (s->isSurface(v) ? VX_VF_SURFACE : 0) | (v.ghost ? VX_VF_GHOST : 0) |
                                       (v.floorOverride == 0 ? VX_VF_FLOOR_OFF : 0) | (v.floorOverride == 1 ? VX_VF_FLOOR_ON : 0); break;
            case VX_F_PSTRAIN: {       // what the links of the next step will read (VX_Voxel.cpp:336-343), without touching the cache: reading must not change when it is filled
                                 const V3f p = (v.pInvalid && !v.ghost) ? s->voxelStrain(const_cast<Voxel&>(v), true) : v.pStrain;
                                 f[3 * k] = p.x; f[3 * k + 1] = p.y; f[3 * k + 2] = p.z; break; }
            default: return VX_ERR_ARG;
            }
        } else {
            const Link& l = s->links[first + k]; const LinkMat& m = s->lmats[l.lmat];
            switch (field) {
            case VX_F_FORCE_NEG: put(k, l.fN); break;
            case VX_F_FORCE_POS: put(k, l.fP); break;
            case VX_F_MOMENT_NEG: put(k, l.mN); break;
            case VX_F_MOMENT_POS: put(k, l.mP); break;
            case VX_F_POS2: put(k, l.pos2); break;
            case VX_F_ANGLE1V: put(k, l.a1v); break;
            case VX_F_ANGLE2V: put(k, l.a2v); break;
            case VX_F_STRAIN: f[k] = l.strain; break;
            case VX_F_MAXSTRAIN: f[k] = l.maxStrain; break;
            case VX_F_STRAINOFFSET: f[k] = l.strainOffset; break;
            case VX_F_STRESS: f[k] = l.stress; break;
            case VX_F_LINKFLAGS: u[k] = (l.smallAngle ? VX_LF_SMALL_ANGLE : 0) | (l.velValid ? VX_LF_LOCAL_VEL_VALID : 0)
                                       | (m.isYielded(l.maxStrain) ? VX_LF_YIELDED : 0) | (m.isFailed(l.maxStrain) ? VX_LF_FAILED : 0); break;
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
        Voxel& v = s->vox[first + k];
        switch (field) {
        case VX_F_POS: v.pos = V3{d[3 * k], d[3 * k + 1], d[3 * k + 2]}; break;
        case VX_F_ORIENT: v.orient = Q4{d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]}; break;
        case VX_F_LINMOM: v.linMom = V3{d[3 * k], d[3 * k + 1], d[3 * k + 2]}; break;
        case VX_F_ANGMOM: v.angMom = V3{d[3 * k], d[3 * k + 1], d[3 * k + 2]}; break;
        case VX_F_TEMP: s->setTemperature(v, f[k]); break;
        case VX_F_PSTRAIN: v.pStrain = V3f{f[3 * k], f[3 * k + 1], f[3 * k + 2]}; v.pInvalid = false; break;      // a halo voxel's Poisson strain comes from its owner
        case VX_F_VOXFLAGS: v.floorStatic = (u[k] & VX_VF_STATIC_FRICTION) != 0;
                            v.floorOverride = (u[k] & VX_VF_FLOOR_OFF) ? 0 : ((u[k] & VX_VF_FLOOR_ON) ? 1 : -1); break;
        default: return VX_ERR_ARG;
        }
    }
    return VX_OK;
}
int vx_collision_pairs(vx_sim* s, int32_t* pairs, int cap, int* n_pairs)
{
    if (!s) return VX_ERR_ARG;
    int n = (int)s->cols.size();
    if (pairs) for (int i = 0; i < n && i < cap; i++) { pairs[2 * i] = s->cols[i].v1; pairs[2 * i + 1] = s->cols[i].v2; }
    if (n_pairs) *n_pairs = n;
    return VX_OK;
}

// CVoxelyze::stateInfo, Voxelyze.cpp:752-800 (enum values of include/Voxelyze.h:48-67)
int vx_state_info(vx_sim* s, int info, int type, float* out)
{
    if (!s || !out) return VX_ERR_ARG;
    enum { DISPLACEMENT, VELOCITY, KINETIC_ENERGY, ANGULAR_DISPLACEMENT, ANGULAR_VELOCITY, ENG_STRESS, ENG_STRAIN, STRAIN_ENERGY, PRESSURE, MASS };
    enum { MIN, MAX, TOTAL, AVERAGE };
    float ret = 0; if (type == MAX) ret = -FLT_MAX; else if (type == MIN) ret = FLT_MAX;
    auto acc = [&](float v) { switch (type) { case MIN: if (v < ret) ret = v; break; case MAX: if (v > ret) ret = v; break; default: ret += v; } };
    if (info == STRAIN_ENERGY || info == ENG_STRESS || info == ENG_STRAIN) {
        if (s->links.empty()) { *out = 0; return VX_OK; }
        for (auto& l : s->links) {
            const LinkMat& m = s->lmats[l.lmat];
            float v = 0;
            if (info == STRAIN_ENERGY)                                                 // VX_Link.cpp:251-257
                v = l.fN.x * l.fN.x / (2.0f * m.a1) + l.mN.x * l.mN.x / (2.0 * m.a2)
                  + (l.mN.z * l.mN.z - l.mN.z * l.mP.z + l.mP.z * l.mP.z) / (3.0 * m.b3)
                  + (l.mN.y * l.mN.y - l.mN.y * l.mP.y + l.mP.y * l.mP.y) / (3.0 * m.b3);
            else if (info == ENG_STRESS) v = l.stress; else v = l.strain;
            acc(v);
        }
        if (type == AVERAGE) ret /= (int)s->links.size();
    } else {
        if (s->vox.empty()) { *out = 0; return VX_OK; }
        int counted = 0;
        for (auto& v : s->vox) {
            if (v.ghost) continue;                       // halo copies are not voxels of this model (their owner counts them)
            counted++;
            const VoxMat& m = s->vmats[v.mat];
            double sz = m.nom; V3 disp = sub(v.pos, V3{v.ix * sz, v.iy * sz, v.iz * sz});
            float val = 0;
            switch (info) {
            case DISPLACEMENT: val = (float)std::sqrt(len2(disp)); break;
            case VELOCITY: val = (float)(std::sqrt(len2(v.linMom)) * m.massInv); break;
            case KINETIC_ENERGY: val = (float)(0.5 * (m.massInv * len2(v.linMom) + m.inertiaInv * len2(v.angMom))); break;
            case ANGULAR_DISPLACEMENT: val = (float)(2.0 * std::acos(v.orient.w > 1 ? 1 : v.orient.w)); break;
            case ANGULAR_VELOCITY: val = (float)(std::sqrt(len2(v.angMom)) * m.inertiaInv); break;
            case PRESSURE: { V3f st = s->voxelStrain(v, false); float vol = (float)(st.x + st.y + st.z); val = -m.E * vol / (3 * (1 - 2 * m.nu)); break; }
            case MASS: val = m.mass; break;
            default: val = 0;
            }
            acc(val);
        }
        if (type == AVERAGE && counted) ret /= counted;
    }
    *out = ret; return VX_OK;
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
// what CVX_Link keeps between steps (VX_Link.h:74-107): everything else a link holds (forces, the two quaternions, rest length,
// transverse strain sums) is recomputed by the next updateForces before it is read
int vx_download_link_state(vx_sim* s, int first, int count, vx_link_state* dst)
{
    if (!s || !dst || first < 0 || count < 0 || first + count > (int)s->links.size()) return VX_ERR_ARG;
    for (int k = 0; k < count; k++) {
        const Link& l = s->links[first + k]; const LinkMat& m = s->lmats[l.lmat];
        vx_link_state& r = dst[k];
        memset(&r, 0, sizeof(r));
        r.pos2[0] = l.pos2.x; r.pos2[1] = l.pos2.y; r.pos2[2] = l.pos2.z;
        r.angle1v[0] = l.a1v.x; r.angle1v[1] = l.a1v.y; r.angle1v[2] = l.a1v.z;
        r.angle2v[0] = l.a2v.x; r.angle2v[1] = l.a2v.y; r.angle2v[2] = l.a2v.z;
        r.strain = l.strain; r.max_strain = l.maxStrain; r.strain_offset = l.strainOffset; r.stress = l.stress;
        uint32_t fl = 0;
        vx_download(s, VX_F_LINKFLAGS, first + k, 1, &fl);
        r.flags = fl; (void)m;
    }
    return VX_OK;
}
int vx_upload_link_state(vx_sim* s, int first, int count, const vx_link_state* src)
{
    if (!s || !src || first < 0 || count < 0 || first + count > (int)s->links.size()) return VX_ERR_ARG;
    for (int k = 0; k < count; k++) {
        Link& l = s->links[first + k]; const vx_link_state& r = src[k];
        l.pos2 = V3{r.pos2[0], r.pos2[1], r.pos2[2]};
        l.a1v = V3{r.angle1v[0], r.angle1v[1], r.angle1v[2]};
        l.a2v = V3{r.angle2v[0], r.angle2v[1], r.angle2v[2]};
        l.strain = r.strain; l.maxStrain = r.max_strain; l.strainOffset = r.strain_offset; l.stress = r.stress;
        l.smallAngle = (r.flags & VX_LF_SMALL_ANGLE) != 0; l.velValid = (r.flags & VX_LF_LOCAL_VEL_VALID) != 0;
    }
    for (auto& v : s->vox) if (!v.ghost) v.pInvalid = true;          // Poisson strains follow the link strains (halo copies keep what their owner sent)
    return VX_OK;
}
int64_t vx_launch_count(const vx_sim*) { return 0; }
int vx_sync(vx_sim*) { return VX_OK; }
int vx_set_path(vx_sim*, int) { return VX_OK; }
int vx_active_path(const vx_sim*) { return 0; }
const char* vx_kernel_name(const vx_sim*) { return "cpu (oracle port)"; }
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
// surface mesh: pinned directly against the reference's CVX_MeshRender (oracle/ref_shim.cpp), not restated here
int vx_mesh_set_material_colors(vx_sim*, int, const unsigned char*) { return VX_ERR_UNSUPPORTED; }
int vx_mesh_build(vx_sim*, int*, int*) { return VX_ERR_UNSUPPORTED; }
int vx_mesh_update(vx_sim*, int, int) { return VX_ERR_UNSUPPORTED; }
int vx_mesh_counts(vx_sim*, int*, int*) { return VX_ERR_UNSUPPORTED; }
int vx_mesh_download(vx_sim*, float*, int32_t*, float*, float*, int32_t*) { return VX_ERR_UNSUPPORTED; }
int vx_mesh_device(vx_sim*, uint64_t*, uint64_t*, uint64_t*, uint64_t*) { return VX_ERR_UNSUPPORTED; }
int vx_collision_stats(vx_sim* s, int* n_pairs, int* n_rebuilds)
{
    if (n_rebuilds) *n_rebuilds = -1;                  // not counted here
    return vx_collision_pairs(s, nullptr, 0, n_pairs);
}

// ---- static solve: CVX_LinearSolver (src/VX_LinearSolver.cpp), restated with a banded Cholesky factorisation in place of
// PARDISO.  Element matrices are written out per link (12 x 12, "voxel 1" = the end with the lower voxelsList index,
// :173), scattered into a symmetric band, the fixed dofs eliminated as applyBX does (:302-327), results posted as
// postResults does (:336-347).  rel_tol / max_iter do not apply to a direct solve.
int vx_linear_solve(vx_sim* s, double, int, int* iterations, double* rel_residual)
{
    if (iterations) *iterations = 0;
    if (rel_residual) *rel_residual = 0.0;
    if (!s) return VX_ERR_ARG;
    const int nv = (int)s->vox.size(), dof = 6 * nv;
    if (dof == 0) return fail(s, VX_ERR_ARG, "vx_linear_solve: no voxels");                        // :60
    int bw = 5;
    for (const Link& l : s->links) bw = std::max(bw, 6 * std::abs(l.vn - l.vp) + 5);
    if ((double)dof * (bw + 1) > 4e8) return fail(s, VX_ERR_ALLOC, "vx_linear_solve (oracle): band too large for the direct solve");
    const size_t W = (size_t)bw + 1;
    std::vector<double> U((size_t)dof * W, 0.0);                                                   // U[i*W + (j-i)], j >= i
    auto at = [&](int i, int j) -> double& { return i <= j ? U[(size_t)i * W + (j - i)] : U[(size_t)j * W + (i - j)]; };
    for (const Link& l : s->links) {
        const LinkMat& m = s->lmats[l.lmat];
        const int i1 = std::min(l.vn, l.vp), i2 = std::max(l.vn, l.vp), ax = l.axis;            // :171-173
        double Ke[12][12] = {};
        auto sym = [&](int r, int c, double v) { Ke[r][c] += v; if (r != c) Ke[c][r] += v; };
        for (int j = 0; j < 3; j++) {
            const float dt = (ax == j) ? m.a1 : m.b1;                                              // :182-187
            sym(j, j, dt); sym(j, 6 + j, -dt); sym(6 + j, 6 + j, dt);
            const float dr = (ax == j) ? m.a2 : 2 * m.b3, orr = (ax == j) ? -m.a2 : m.b3;          // :189-195
            sym(3 + j, 3 + j, dr); sym(3 + j, 9 + j, orr); sym(9 + j, 9 + j, dr);
        }
        int R1, C1, R2, C2; float val;                                                             // :199-217
        if (ax == 0) { R1 = 1; C1 = 5; R2 = 2; C2 = 4; val = m.b2; }
        else if (ax == 1) { R1 = 0; C1 = 5; R2 = 2; C2 = 3; val = -m.b2; }
        else { R1 = 0; C1 = 4; R2 = 1; C2 = 3; val = m.b2; }
        sym(R1, C1, val); sym(R1, 6 + C1, val); sym(C1, 6 + R1, -val); sym(6 + R1, 6 + C1, -val);  // :219-222
        sym(R2, C2, -val); sym(R2, 6 + C2, -val); sym(C2, 6 + R2, val); sym(6 + R2, 6 + C2, val);  // :224-227
        for (int r = 0; r < 12; r++)
            for (int c = r; c < 12; c++) {
                if (Ke[r][c] == 0.0) continue;
                const int gr = (r < 6 ? 6 * i1 + r : 6 * i2 + r - 6), gc = (c < 6 ? 6 * i1 + c : 6 * i2 + c - 6);
                at(gr, gc) += Ke[r][c];
            }
    }
    std::vector<double> x(dof, 0.0), b(dof, 0.0); std::vector<char> fixed(dof, 0);
    for (int i = 0; i < nv; i++) {                                                                 // :281-300
        const Voxel& v = s->vox[i];
        const double sz = s->vmats[v.mat].nom;
        const V3 disp = sub(v.pos, V3{v.ix * sz, v.iy * sz, v.iz * sz});
        const V3 ang = v.orient.w == 1 ? V3{} : to_rotation_vector(v.orient);
        const double u[6] = {disp.x, disp.y, disp.z, ang.x, ang.y, ang.z};
        const Ext* e = v.ext >= 0 ? &s->exts[v.ext] : nullptr;
        for (int j = 0; j < 6; j++) {
            const int d = 6 * i + j;
            x[d] = u[j];
            fixed[d] = e ? ((e->dof >> j) & 1) : 0;
            if (!fixed[d] && e) b[d] = j < 3 ? e->f[j] : e->m[j - 3];
        }
    }
    for (int d = 0; d < dof; d++) {                                                                // :302-327
        if (!fixed[d]) continue;
        const int lo = std::max(0, d - bw), hi = std::min(dof - 1, d + bw);
        for (int i = lo; i <= hi; i++) if (i != d) { b[i] -= x[d] * at(i, d); }
    }
    for (int d = 0; d < dof; d++) {
        if (!fixed[d]) continue;
        const int lo = std::max(0, d - bw), hi = std::min(dof - 1, d + bw);
        for (int i = lo; i <= hi; i++) at(i, d) = 0.0;
        at(d, d) = 1.0; b[d] = x[d];
    }
    // U^T U factorisation in the band, then two triangular solves
    std::vector<double> diag0(dof);
    for (int k = 0; k < dof; k++) diag0[k] = U[(size_t)k * W];
    for (int k = 0; k < dof; k++) {
        double* row = &U[(size_t)k * W];
        if (!(row[0] > 1e-11 * diag0[k])) return fail(s, VX_ERR_SOLVER, "vx_linear_solve: the stiffness matrix is singular (a part of the model is not held) or not positive definite");      // pivot lost to cancellation: singular to working precision
        const double piv = std::sqrt(row[0]);
        const int m = std::min(bw, dof - 1 - k);
        for (int j = 0; j <= m; j++) row[j] /= piv;
        for (int i = 1; i <= m; i++) {
            const double f = row[i];
            if (f == 0.0) continue;
            double* ri = &U[(size_t)(k + i) * W];
            for (int j = i; j <= m; j++) ri[j - i] -= f * row[j];
        }
    }
    for (int k = 0; k < dof; k++) {                                                                // U^T y = b
        const double* row = &U[(size_t)k * W];
        b[k] /= row[0];
        const int m = std::min(bw, dof - 1 - k);
        for (int j = 1; j <= m; j++) b[k + j] -= row[j] * b[k];
    }
    for (int k = dof - 1; k >= 0; k--) {                                                           // U x = y
        const double* row = &U[(size_t)k * W];
        const int m = std::min(bw, dof - 1 - k);
        double t = b[k];
        for (int j = 1; j <= m; j++) t -= row[j] * b[k + j];
        b[k] = t / row[0];
    }
    for (int i = 0; i < nv; i++) {                                                                 // :336-347
        Voxel& v = s->vox[i];
        const double sz = s->vmats[v.mat].nom;
        v.pos = add(V3{v.ix * sz, v.iy * sz, v.iz * sz}, V3{b[6 * i], b[6 * i + 1], b[6 * i + 2]});
        v.linMom = V3{};
        v.orient = from_rotation_vector(V3{b[6 * i + 3], b[6 * i + 4], b[6 * i + 5]});
        v.angMom = V3{};
    }
    return VX_OK;
}

} // extern "C"

// vx_slabbed_*: host-side composition over the entry points above (shared with the product library: the partition logic under test)
#include "../voxelyze_b200/csrc/vx_slabbed.hpp"
