// voxelyze_json.cpp -- *.vxl.json load / save of the facade.
//
// Same file format as the reference (CVoxelyze::readJSON / writeJSON src/Voxelyze.cpp:95-241,
// CVX_Material::readJSON / writeJSON src/VX_Material.cpp:75-163), with its own small parser in
// place of the vendored RapidJSON.  Compatibility rules, all taken from the reference code:
//   * a numeric member only counts where the reference tests IsDouble(): the literal must have a
//     fraction or an exponent ("1000000" is ignored exactly like the reference ignores it;
//     its own writer always emits "1000000.0");
//   * environment members (gravityAcceleration, floorEnabled, collisionsEnabled,
//     relativeAmbientTemperature) are written but NOT read back (src/Voxelyze.cpp:168-171 vs
//     :95-161) -- a caller sets them after loading, as with the reference;
//   * the reference only closes the root object when externals exist (:211,237): files without
//     externals end unterminated.  The loader accepts such files, the writer always closes;
//   * the reference's data-curve reader indexes the array itself instead of element i
//     (src/VX_Material.cpp:136-137, asserts at run time); the evident intent is implemented.
#include <cctype>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "Voxelyze.h"

namespace {

struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0.0; bool isDouble = false; long long inum = 0;   // isDouble: the literal had '.', 'e' or 'E'
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;

    const JVal* get(const char* key) const
    {
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const JVal* getDouble(const char* key) const { const JVal* v = get(key); return v && v->kind == Num && v->isDouble ? v : nullptr; }
    const JVal* getInt(const char* key) const { const JVal* v = get(key); return v && v->kind == Num && !v->isDouble && v->inum >= INT_MIN && v->inum <= INT_MAX ? v : nullptr; }
    const JVal* getArray(const char* key) const { const JVal* v = get(key); return v && v->kind == Arr ? v : nullptr; }
};

struct Parser {
    const char* p; const char* end; bool ok = true;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool lit(const char* s) { size_t n = strlen(s); if ((size_t)(end - p) >= n && memcmp(p, s, n) == 0) { p += n; return true; } return false; }
    JVal value(int depth)
    {
        JVal v; ws();
        if (p >= end || depth > 64) { ok = false; return v; }
        if (*p == '{') {
            p++; v.kind = JVal::Obj; ws();
            if (p < end && *p == '}') { p++; return v; }
            while (ok) {
                ws();
                if (p >= end || *p != '"') { ok = false; break; }
                JVal k = value(depth + 1);
                ws();
                if (!ok || p >= end || *p != ':') { ok = false; break; }
                p++;
                v.obj.emplace_back(k.str, value(depth + 1));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                if (p >= end && depth == 0) break;         // the reference leaves the root open when there are no externals
                ok = false;
            }
        } else if (*p == '[') {
            p++; v.kind = JVal::Arr; ws();
            if (p < end && *p == ']') { p++; return v; }
            while (ok) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                ok = false;
            }
        } else if (*p == '"') {
            p++; v.kind = JVal::Str;
            while (p < end && *p != '"') {
                if (*p == '\\' && p + 1 < end) {
                    p++;
                    switch (*p) {
                    case 'n': v.str += '\n'; break; case 't': v.str += '\t'; break; case 'r': v.str += '\r'; break;
                    case 'b': v.str += '\b'; break; case 'f': v.str += '\f'; break;
                    case 'u': if (end - p >= 5) { v.str += (char)strtol(std::string(p + 1, p + 5).c_str(), nullptr, 16); p += 4; } break;
                    default: v.str += *p;
                    }
                    p++;
                } else v.str += *p++;
            }
            if (p >= end) ok = false; else p++;
        } else if (lit("true")) { v.kind = JVal::Bool; v.b = true; }
        else if (lit("false")) { v.kind = JVal::Bool; v.b = false; }
        else if (lit("null")) { v.kind = JVal::Null; }
        else {
            const char* s = p;
            if (p < end && (*p == '-' || *p == '+')) p++;
            while (p < end && (isdigit((unsigned char)*p) || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) p++;
            if (p == s) { ok = false; return v; }
            std::string t(s, p);
            v.kind = JVal::Num;
            v.isDouble = t.find_first_of(".eE") != std::string::npos;
            v.num = strtod(t.c_str(), nullptr);
            if (!v.isDouble) {
                char* e2 = nullptr; errno = 0;
                v.inum = strtoll(t.c_str(), &e2, 10);
                if (errno == ERANGE) v.isDouble = true;    // RapidJSON falls back to double beyond 64 bits
            }
        }
        return v;
    }
};

struct Writer {
    std::ostringstream o; int indent = 0; std::vector<bool> first; bool afterKey = false;
    void nl() { o << '\n'; for (int i = 0; i < indent; i++) o << "    "; }
    void prefix()
    {
        if (afterKey) { afterKey = false; return; }
        if (!first.empty()) { if (!first.back()) o << ','; first.back() = false; nl(); }
    }
    void open(char c) { prefix(); o << c; indent++; first.push_back(true); }
    void close(char c) { indent--; bool empty = first.back(); first.pop_back(); if (!empty) nl(); o << c; }
    void key(const char* k) { prefix(); o << '"' << k << "\": "; afterKey = true; }
    void num(double d)
    {
        prefix();
        char buf[40]; snprintf(buf, sizeof(buf), "%.17g", d);
        for (int prec = 1; prec < 17; prec++) { char t[40]; snprintf(t, sizeof(t), "%.*g", prec, d); if (strtod(t, nullptr) == d) { strcpy(buf, t); break; } }
        o << buf;
        if (!strpbrk(buf, ".eEn")) o << ".0";               // always reads back as a double, like RapidJSON's writer
    }
    void integer(long long i) { prefix(); o << i; }
    void boolean(bool b) { prefix(); o << (b ? "true" : "false"); }
    void str(const std::string& s)
    {
        prefix(); o << '"';
        for (char c : s) { if (c == '"' || c == '\\') o << '\\' << c; else if (c == '\n') o << "\\n"; else if (c == '\t') o << "\\t"; else o << c; }
        o << '"';
    }
};

} // namespace

// ---- materials: CVX_Material::readJSON / writeJSON (src/VX_Material.cpp:75-163)
static bool materialFromJSON(CVX_Material& mat, const JVal& m, vxm::Material& raw);

bool CVoxelyze::loadJSON(const char* jsonFilePath)
{
    std::ifstream t(jsonFilePath);
    if (!t) return false;
    std::stringstream buffer; buffer << t.rdbuf();
    const std::string text = buffer.str();
    Parser ps{text.data(), text.data() + text.size()};
    JVal vxl = ps.value(0);
    // like the reference, a file that opens is reported as loaded whatever readJSON made of it (src/Voxelyze.cpp:61-76)
    clear();
    if (!ps.ok || vxl.kind != JVal::Obj) return true;
    const JVal* vs = vxl.getDouble("voxelSize");
    if (!vs) return true;
    voxSize = vs->num;
    const JVal* mats = vxl.getArray("materials");
    if (!mats) return true;
    for (const JVal& m : mats->arr) {
        CVX_MaterialVoxel* mv = new CVX_MaterialVoxel(1e6f, 1e3f, voxSize);
        vxm::Material raw;
        materialFromJSON(*mv, m, raw);                      // an invalid entry still takes its slot (src/Voxelyze.cpp:390-402)
        mv->gravMult_ = grav;
        voxelMats.push_back(mv);
    }
    topologyDirty = true;

    const JVal* v = vxl.getArray("voxels");
    if (v && v->arr.size() % 4 == 0) {
        for (size_t i = 0; i < v->arr.size() / 4; i++) {
            const long long mi = v->arr[4 * i + 3].inum;
            if (mi < 0 || mi >= (long long)voxelMats.size()) continue;
            setVoxel(voxelMats[(size_t)mi], (int)v->arr[4 * i].inum, (int)v->arr[4 * i + 1].inum, (int)v->arr[4 * i + 2].inum);
        }
    }
    if (const JVal* exts = vxl.getArray("externals")) {
        for (const JVal& ext : exts->arr) {
            const JVal* idx = ext.getArray("voxelIndices");
            if (!idx) continue;                              // invalid external
            bool dof[6] = {false, false, false, false, false, false};
            double disp[6] = {0, 0, 0, 0, 0, 0};
            Vec3D<float> force, moment;
            const JVal* a;
            if ((a = ext.getArray("fixed")) && a->arr.size() == 6) for (int j = 0; j < 6; j++) dof[j] = a->arr[j].kind == JVal::Bool && a->arr[j].b;
            if ((a = ext.getArray("translate")) && a->arr.size() == 3) for (int j = 0; j < 3; j++) disp[j] = a->arr[j].num;
            if ((a = ext.getArray("rotate")) && a->arr.size() == 3) for (int j = 0; j < 3; j++) disp[3 + j] = a->arr[j].num;
            if ((a = ext.getArray("force")) && a->arr.size() == 3) for (int j = 0; j < 3; j++) force[j] = (float)a->arr[j].num;
            if ((a = ext.getArray("moment")) && a->arr.size() == 3) for (int j = 0; j < 3; j++) moment[j] = (float)a->arr[j].num;
            for (const JVal& vi : idx->arr) {
                if (vi.inum < 0 || vi.inum >= (long long)voxelsList.size()) continue;
                CVX_External* pE = voxelsList[(size_t)vi.inum]->external();
                for (int k = 0; k < 6; k++) if (dof[k]) pE->setDisplacement((dofComponent)(1 << k), disp[k]);
                pE->addForce(force);
                pE->addMoment(moment);
            }
        }
    }
    return true;
}

static bool materialFromJSON(CVX_Material& mat, const JVal& m, vxm::Material& raw)
{
    mat.clear();
    if (m.kind != JVal::Obj) return false;
    const JVal* E = m.getDouble("youngsModulus");
    const JVal* sd = m.getArray("strainData");
    const JVal* ss = m.getArray("stressData");
    if (E) {
        float failStress = -1.0f;
        if (const JVal* ef = m.getDouble("epsilonFail")) failStress = ef->num * E->num;
        mat.setModelLinear(E->num, failStress);
    } else if (sd && ss && sd->arr.size() == ss->arr.size()) {
        std::vector<float> stress, strain;
        for (size_t i = 0; i < sd->arr.size(); i++) { stress.push_back((float)ss->arr[i].num); strain.push_back((float)sd->arr[i].num); }
        if (strain.empty() || !mat.setModel((int)strain.size(), &strain[0], &stress[0])) return false;
    } else return false;                                    // no valid model

    struct Access : CVX_Material { static vxm::Material& model(CVX_Material& c) { return static_cast<Access&>(c).m_; }
                                   static std::string& name(CVX_Material& c) { return static_cast<Access&>(c).name_; }
                                   static int& col(CVX_Material& c, int k) { Access& a = static_cast<Access&>(c); return k == 0 ? a.r_ : k == 1 ? a.g_ : k == 2 ? a.b_ : a.a_; }
                                   static void touch(CVX_Material& c) { static_cast<Access&>(c).changed(); } };
    vxm::Material& mm = Access::model(mat);
    const JVal* v;                                            // assigned without the setters' clamping, like the reference
    if ((v = m.getDouble("density"))) mm.rho = v->num;
    if ((v = m.get("name")) && v->kind == JVal::Str) Access::name(mat) = v->str;
    const char* colours[4] = {"red", "green", "blue", "alpha"};
    for (int k = 0; k < 4; k++) if ((v = m.getInt(colours[k]))) Access::col(mat, k) = (int)v->inum;
    if ((v = m.getDouble("poissonsRatio"))) mm.nu = v->num;
    if ((v = m.getDouble("CTE"))) mm.cte = v->num;
    if ((v = m.getDouble("staticFriction"))) mm.mu_s = v->num;
    if ((v = m.getDouble("kineticFriction"))) mm.mu_k = v->num;
    if ((v = m.getDouble("internalDamping"))) mm.zeta_int = v->num;
    if ((v = m.getDouble("globalDamping"))) mm.zeta_glob = v->num;
    if ((v = m.getDouble("collisionDamping"))) mm.zeta_coll = v->num;
    if ((v = m.getArray("externalScaleFactor")) && v->arr.size() == 3) for (int i = 0; i < 3; i++) mm.ext_scale[i] = v->arr[i].num;
    Access::touch(mat);                                      // updateDerived()
    raw = mm;
    return true;
}

static void materialToJSON(CVX_Material& mat, Writer& w)
{
    w.open('{');
    if (mat.isModelLinear()) {
        w.key("youngsModulus"); w.num((double)mat.youngsModulus());
        const vxm::Material& mm = mat.model();
        if (mm.eps_fail != -1) { w.key("epsilonFail"); w.num((double)mm.eps_fail); }
    } else {
        w.key("strainData"); w.open('[');
        for (int i = 0; i < mat.modelDataPoints(); i++) w.num((double)mat.modelDataStrain()[i]);
        w.close(']');
        w.key("stressData"); w.open('[');
        for (int i = 0; i < mat.modelDataPoints(); i++) w.num((double)mat.modelDataStress()[i]);
        w.close(']');
    }
    if (mat.density() != 1.0f) { w.key("density"); w.num(mat.density()); }
    if (std::string(mat.name()) != "") { w.key("name"); w.str(mat.name()); }
    if (mat.red() != -1) { w.key("red"); w.integer(mat.red()); }
    if (mat.green() != -1) { w.key("green"); w.integer(mat.green()); }
    if (mat.blue() != -1) { w.key("blue"); w.integer(mat.blue()); }
    if (mat.alpha() != -1) { w.key("alpha"); w.integer(mat.alpha()); }
    if (mat.poissonsRatio() != 0) { w.key("poissonsRatio"); w.num(mat.poissonsRatio()); }
    if (mat.cte() != 0) { w.key("CTE"); w.num(mat.cte()); }
    if (mat.staticFriction() != 0) { w.key("staticFriction"); w.num(mat.staticFriction()); }
    if (mat.kineticFriction() != 0) { w.key("kineticFriction"); w.num(mat.kineticFriction()); }
    if (mat.internalDamping() != 1) { w.key("internalDamping"); w.num(mat.internalDamping()); }
    if (mat.globalDamping() != 0) { w.key("globalDamping"); w.num(mat.globalDamping()); }
    if (mat.collisionDamping() != 1) { w.key("collisionDamping"); w.num(mat.collisionDamping()); }
    Vec3D<double> es = mat.externalScaleFactor();
    if (es.x != 1 || es.y != 1 || es.z != 1) {
        w.key("externalScaleFactor"); w.open('[');
        for (int i = 0; i < 3; i++) w.num(es[i]);
        w.close(']');
    }
    w.close('}');
}

bool CVoxelyze::saveJSON(const char* jsonFilePath)
{
    std::ofstream t(jsonFilePath);
    if (!t) return false;
    Writer w;
    w.open('{');
    w.key("voxelSize"); w.num(voxSize);
    if (ambientTemp != 0) { w.key("relativeAmbientTemperature"); w.num((double)ambientTemp); }
    if (grav != 0) { w.key("gravityAcceleration"); w.num((double)grav); }
    if (floor) { w.key("floorEnabled"); w.boolean(floor); }
    if (collisions) { w.key("collisionsEnabled"); w.boolean(collisions); }

    std::unordered_map<CVX_Material*, int> m2i;
    w.key("materials"); w.open('[');
    for (int i = 0; i < materialCount(); i++) { m2i[material(i)] = i; materialToJSON(*material(i), w); }
    w.close(']');

    std::vector<CVX_External*> exts;                          // catalogue of distinct externals, first-seen order
    std::vector<std::vector<int>> extVoxIndices;
    w.key("voxels"); w.open('[');
    for (int i = 0; i < voxelCount(); i++) {
        CVX_Voxel* pVox = voxel(i);
        w.integer(pVox->indexX()); w.integer(pVox->indexY()); w.integer(pVox->indexZ()); w.integer(m2i[pVox->material()]);
        if (pVox->externalExists() && !pVox->external()->isEmpty()) {
            bool match = false;
            for (size_t j = 0; j < exts.size() && !match; j++) if (*pVox->external() == *exts[j]) { extVoxIndices[j].push_back(i); match = true; }
            if (!match) { exts.push_back(pVox->external()); extVoxIndices.push_back(std::vector<int>(1, i)); }
        }
    }
    w.close(']');

    if (!exts.empty()) {
        w.key("externals"); w.open('[');
        for (size_t i = 0; i < exts.size(); i++) {
            CVX_External* e = exts[i];
            w.open('{');
            if (e->isFixedAny()) {
                w.key("fixed"); w.open('[');
                w.boolean(e->isFixed(X_TRANSLATE)); w.boolean(e->isFixed(Y_TRANSLATE)); w.boolean(e->isFixed(Z_TRANSLATE));
                w.boolean(e->isFixed(X_ROTATE)); w.boolean(e->isFixed(Y_ROTATE)); w.boolean(e->isFixed(Z_ROTATE));
                w.close(']');
            }
            if (e->isFixedAnyTranslation() && !(e->translation() == Vec3D<double>())) { w.key("translate"); w.open('['); for (int j = 0; j < 3; j++) w.num(e->translation()[j]); w.close(']'); }
            if (e->isFixedAnyRotation() && !(e->rotation() == Vec3D<double>())) { w.key("rotate"); w.open('['); for (int j = 0; j < 3; j++) w.num(e->rotation()[j]); w.close(']'); }
            if (!e->isFixedAllTranslation() && !(e->force() == Vec3D<float>())) { w.key("force"); w.open('['); for (int j = 0; j < 3; j++) w.num(e->force()[j]); w.close(']'); }
            if (!e->isFixedAllRotation() && !(e->moment() == Vec3D<float>())) { w.key("moment"); w.open('['); for (int j = 0; j < 3; j++) w.num(e->moment()[j]); w.close(']'); }
            w.key("voxelIndices"); w.open('[');
            for (int vi : extVoxIndices[i]) w.integer(vi);
            w.close(']');
            w.close('}');
        }
        w.close(']');
    }
    w.close('}');
    t << w.o.str() << '\n';
    t.close();
    return true;
}

// ---- dynamic state (additive; the reference cannot checkpoint)
bool CVoxelyze::saveState(const char* path) { sync(); return hm ? vx_slabbed_save_state(hm, path) == VX_OK : (h && vx_save_state(h, path) == VX_OK); }
bool CVoxelyze::loadState(const char* path)
{
    sync();
    if (hm ? vx_slabbed_load_state(hm, path) != VX_OK : (!h || vx_load_state(h, path) != VX_OK)) return false;       // slabbed: one file per slab, same slab count
    stepped = true; epoch++;
    return true;
}
