// dropin_tests.cpp -- one test source, two implementations of the same public C++ API.
//
//   -DDROPIN_REFERENCE + reference headers/sources  -> tests/cpp/_build/dropin_ref   (CPU, unmodified reference)
//   facade headers + libvoxelyze_facade/b200        -> tests/cpp/_build/dropin_b200  (B200)
//
// Only the public API of include/Voxelyze.h is used, so passing both builds is the drop-in
// proof.  The scenarios restate the reference's gtest cases (test/tVoxelyze.h, test/tVX_Material.h,
// test/tVX_MaterialLink.h, test/tVX_Voxel.h; line cited per case) with the reference's own expected
// values and tolerances; gtest itself is not available in this image, hence the tiny harness.
#include "Voxelyze.h"
#include "VX_Voxel.h"
#include "VX_Link.h"
#include "VX_MaterialLink.h"
#include "VX_MeshRender.h"
#ifndef DROPIN_REFERENCE
#include "VX_LinearSolver.h"
#endif

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

static int g_failures = 0, g_checks = 0;
static const char* g_current = "";
#define CHECK(cond) do { g_checks++; if (!(cond)) { g_failures++; printf("  FAIL %s:%d  %s   [%s]\n", __FILE__, __LINE__, #cond, g_current); } } while (0)
#define CHECK_NEAR(a, b, tol) do { g_checks++; double a_ = (a), b_ = (b); if (!(std::fabs(a_ - b_) <= (tol))) { g_failures++; \
    printf("  FAIL %s:%d  |%s - %s| = |%.10g - %.10g| > %g   [%s]\n", __FILE__, __LINE__, #a, #b, a_, b_, (double)(tol), g_current); } } while (0)
// gtest's EXPECT_FLOAT_EQ: within 4 ulp of float
static bool float_eq(float a, float b)
{
    if (a == b) return true;
    int ia, ib; memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
    if ((ia < 0) != (ib < 0)) return false;
    return std::abs(ia - ib) <= 4;
}
#define CHECK_FLOAT_EQ(a, b) do { g_checks++; float a_ = (float)(a), b_ = (float)(b); if (!float_eq(a_, b_)) { g_failures++; \
    printf("  FAIL %s:%d  %s = %.9g != %s = %.9g   [%s]\n", __FILE__, __LINE__, #a, a_, #b, b_, g_current); } } while (0)

static Vec3D<> offsetOf(CVX_Voxel::linkDirection d)
{
    switch (d) {
    case CVX_Voxel::X_POS: return Vec3D<>(1, 0, 0);
    case CVX_Voxel::X_NEG: return Vec3D<>(-1, 0, 0);
    case CVX_Voxel::Y_POS: return Vec3D<>(0, 1, 0);
    case CVX_Voxel::Y_NEG: return Vec3D<>(0, -1, 0);
    case CVX_Voxel::Z_POS: return Vec3D<>(0, 0, 1);
    default: return Vec3D<>(0, 0, -1);
    }
}

// tVoxelyze.h:63-113: two voxels, one fixed (or mirrored load), returns steps until converged
static int twoVoxels(bool firstFixed, CVX_Voxel::linkDirection dir, Vec3D<float> force, Vec3D<float> moment, dofObject dofs,
                     int maxSteps, float expected, int component)
{
    const double sz = 0.001;
    CVoxelyze sim(sz);
    CVX_Material* mat = sim.addMaterial(1e6, 1e3);
    mat->setInternalDamping(1.0);
    mat->setGlobalDamping(0.2f);
    Vec3D<> off = offsetOf(dir);
    CVX_Voxel* a = sim.setVoxel(mat, 0, 0, 0);
    auto fix = [&](CVX_Voxel* v) {
        v->external()->setFixed(dofIsSet(dofs, X_TRANSLATE), dofIsSet(dofs, Y_TRANSLATE), dofIsSet(dofs, Z_TRANSLATE),
                                dofIsSet(dofs, X_ROTATE), dofIsSet(dofs, Y_ROTATE), dofIsSet(dofs, Z_ROTATE));
    };
    if (firstFixed) a->external()->setFixedAll();
    else { a->external()->setForce(-force); a->external()->setMoment(-moment); fix(a); }
    CVX_Voxel* b = sim.setVoxel(mat, (int)off.x, (int)off.y, (int)off.z);
    b->external()->setForce(force); b->external()->setMoment(moment); fix(b);

    float ts = sim.recommendedTimeStep();
    int streak = 0, k;
    for (k = 0; k < maxSteps; k++) {
        sim.doTimeStep(ts);
        double value = component < 3 ? (b->position() - off * sz)[component] : b->orientation().ToRotationVector()[component % 3];
        if (std::fabs(value - expected) < std::fabs(expected) * 1e-5 || (float)value == (float)expected) streak++; else streak = 0;
        if (streak == 10) break;
    }
    return k;
}

// ------------------------------------------------------------------------------------------------
static void simpleSetup()               // tVoxelyze.h:117-139
{
    CVoxelyze sim(0.001f);
    CVX_Material* m = sim.addMaterial();
    sim.setVoxel(m, 0, 0, 0);
    sim.setVoxel(m, 1, 0, 0);
    CHECK(sim.indexMinX() == 0 && sim.indexMaxX() == 1 && sim.indexMinY() == 0 && sim.indexMaxY() == 0 && sim.indexMinZ() == 0 && sim.indexMaxZ() == 0);
    CHECK(sim.voxelCount() == 2);
    CHECK(sim.voxel(0, 0, 0)->material() == m && sim.voxel(1, 0, 0)->material() == m);
    CHECK(sim.voxel(2, 0, 0) == NULL);
    CHECK(sim.linkCount() == 1);
    CHECK(sim.link(0, 0, 0, CVX_Voxel::X_POS) != NULL && sim.link(0, 0, 0, CVX_Voxel::X_POS) == sim.link(1, 0, 0, CVX_Voxel::X_NEG));
    CHECK(sim.link(0, 0, 0, CVX_Voxel::Y_POS) == NULL);
}

static void singleBondFixedFree()       // tVoxelyze.h:142-198 (one case per load type and axis family)
{
    const Vec3D<float> none(0, 0, 0);
    CHECK(150 > twoVoxels(true, CVX_Voxel::X_POS, Vec3D<float>(1e-3f, 0, 0), none, dof(false, true, true, true, true, true), 1000, 1e-6f, 0));
    CHECK(150 > twoVoxels(true, CVX_Voxel::X_POS, Vec3D<float>(0, -1e-3f, 0), none, dof(true, false, true, true, true, true), 1000, -1e-6f, 1));
    CHECK(150 > twoVoxels(true, CVX_Voxel::Y_NEG, Vec3D<float>(1e-3f, 0, 0), none, dof(false, true, true, true, true, true), 1000, 1e-6f, 0));
    CHECK(150 > twoVoxels(true, CVX_Voxel::Z_POS, Vec3D<float>(0, 0, 1e-3f), none, dof(true, true, false, true, true, true), 1000, 1e-6f, 2));
    CHECK(200 > twoVoxels(true, CVX_Voxel::X_POS, none, Vec3D<float>(1e-9f, 0, 0), dof(true, true, true, false, true, true), 1000, 1.2e-5f, 3));
    CHECK(200 > twoVoxels(true, CVX_Voxel::X_POS, none, Vec3D<float>(0, 1e-9f, 0), dof(true, true, true, true, false, true), 1000, 3e-6f, 4));
    CHECK(200 > twoVoxels(true, CVX_Voxel::Z_NEG, none, Vec3D<float>(0, 0, 1e-9f), dof(true, true, true, true, true, false), 1000, 1.2e-5f, 5));
    CHECK(300 > twoVoxels(true, CVX_Voxel::X_POS, none, Vec3D<float>(0, 1e-9f, 0), dof(true, true, false, true, false, true), 1000, -6e-9f, 2));
    CHECK(300 > twoVoxels(true, CVX_Voxel::X_POS, Vec3D<float>(0, 1e-3f, 0), none, dof(true, false, true, true, true, false), 1000, 6e-3f, 5));
}

static void singleBondFreeFree()        // tVoxelyze.h:202-244
{
    const Vec3D<float> none(0, 0, 0);
    CHECK(100 > twoVoxels(false, CVX_Voxel::X_POS, Vec3D<float>(1e-3f, 0, 0), none, dof(false, true, true, true, true, true), 1000, 5e-7f, 0));
    CHECK(100 > twoVoxels(false, CVX_Voxel::Y_POS, Vec3D<float>(0, 1e-3f, 0), none, dof(true, false, true, true, true, true), 1000, 5e-7f, 1));
    CHECK(150 > twoVoxels(false, CVX_Voxel::X_POS, none, Vec3D<float>(0, 0, 1e-9f), dof(true, true, true, true, true, false), 1000, 6e-6f, 5));
    CHECK(150 > twoVoxels(false, CVX_Voxel::Z_NEG, none, Vec3D<float>(1e-9f, 0, 0), dof(true, true, true, false, true, true), 1000, 6e-6f, 3));
}

static void resetTime()                 // tVoxelyze.h:246-259
{
    CVoxelyze sim(0.001);
    CVX_Material* m = sim.addMaterial(1e6, 1e3);
    CVX_Voxel* a = sim.setVoxel(m, 0, 0, 0);
    CVX_Voxel* b = sim.setVoxel(m, 1, 0, 0);
    a->external()->setFixedAll();
    b->external()->setForce(1e-3f, 0, 0);
    for (int i = 0; i < 100; i++) sim.doTimeStep();
    CHECK(b->position().x > 0.001);
    sim.resetTime();
    CHECK_FLOAT_EQ(1e-3f, (float)b->position().x);
}

static void dampingInternalAndGlobal()  // tVoxelyze.h:261-334
{
    for (int variant = 0; variant < 2; variant++) {
        CVoxelyze sim(0.001);
        CVX_Material* m = sim.addMaterial(1e6, 1e3);
        if (variant == 0) m->setInternalDamping(1.0);
        else { m->setGlobalDamping(1.0); m->setInternalDamping(0); }
        CVX_Voxel* c = sim.setVoxel(m, 0, 0, 0);
        c->external()->setForce(1e-6f, 1e-6f, 1e-6f);
        for (int i = 0; i < 6; i++) {
            Vec3D<> o = offsetOf((CVX_Voxel::linkDirection)i);
            sim.setVoxel(m, (int)o.x, (int)o.y, (int)o.z)->external()->setFixedAll();
        }
        float ts = sim.recommendedTimeStep();
        for (int k = 0; k < 100; k++) sim.doTimeStep(ts);
        CHECK_FLOAT_EQ(1e-9f / 6, (float)c->position().y);
        if (variant == 1) { CHECK_FLOAT_EQ(1e-9f / 6, (float)c->position().x); CHECK_FLOAT_EQ(1e-9f / 6, (float)c->position().z); }
    }
    CVoxelyze sim2(0.001);              // two-voxel cantilever
    CVX_Material* m2 = sim2.addMaterial(1e6, 1e3);
    m2->setGlobalDamping(0.25f); m2->setInternalDamping(0);
    sim2.setVoxel(m2, 0, 0, 0)->external()->setFixedAll();
    CVX_Voxel* tip = sim2.setVoxel(m2, 1, 0, 0);
    tip->external()->setForce(1e-6f, 1e-6f, 1e-6f);
    float ts = sim2.recommendedTimeStep();
    for (int k = 0; k < 300; k++) sim2.doTimeStep(ts);
    CHECK_FLOAT_EQ(4e-9f, (float)tip->position().y);
}

static void combinedDamping()           // tVoxelyze.h:336-379
{
    CVoxelyze sim(0.001);
    CVX_Material* m = sim.addMaterial(1e6, 1e3);
    m->setInternalDamping(1.0f);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
        CVX_Voxel* v = sim.setVoxel(m, i, j, k);
        if (i == 0) v->external()->setFixedAll();
        else if (i == 3) v->external()->setForce(0, 0, 1e-6f);
    }
    for (int i = 0; i < 2; i++) {
        m->setGlobalDamping(0.05f + 0.05f * i);     // material changed between runs through the handle
        float ts = sim.recommendedTimeStep();
        sim.resetTime();
        for (int k = 0; k < 1000; k++) sim.doTimeStep(ts);
        CHECK_NEAR(1.742e-8, sim.voxel(3, 0, 0)->position().z, 1e-10);
    }
}

static void scaleOfForce()              // tVoxelyze.h:382-412
{
    CVoxelyze sim(0.001);
    CVX_Material* m = sim.addMaterial(1e6, 1e3);
    m->setInternalDamping(1.0); m->setGlobalDamping(0.2f);
    sim.setVoxel(m, 0, 0, 0)->external()->setFixedAll();
    CVX_Voxel* b = sim.setVoxel(m, 1, 0, 0);
    for (int i = 3; i < 8; i++) {
        float force = 1 / std::pow(10.0f, i), result = 4 / std::pow(10.0f, i + 3);
        b->external()->setForce(force, force, force);
        float ts = sim.recommendedTimeStep();
        sim.resetTime();
        for (int k = 0; k < 240; k++) sim.doTimeStep(ts);
        CHECK_NEAR(result, (float)b->position().y, result / 1000);
    }
}

static void axialFrequency()            // tVoxelyze.h:415-443
{
    CVoxelyze sim(0.001f);
    CVX_Material* m = sim.addMaterial(1e6f, 1e3f);
    m->setInternalDamping(0);
    sim.setVoxel(m, 0, 0, 0)->external()->setFixedAll();
    CVX_Voxel* b = sim.setVoxel(m, 1, 0, 0);
    b->external()->setFixed(false, true, true, true, true, true);
    b->external()->setForce(1e-3f, 0, 0);
    float ts = sim.recommendedTimeStep() / 10;
    std::vector<double> data;
    for (int i = 0; i < 1000; i++) { sim.doTimeStep(ts); data.push_back(b->position().x - 0.001001); }
    std::vector<double> crossings;
    for (size_t i = 1; i < data.size(); i++)
        if ((data[i - 1] <= 0 && data[i] > 0) || (data[i - 1] >= 0 && data[i] < 0)) crossings.push_back(ts * ((double)i - 1 + data[i - 1] / (data[i - 1] - data[i])));
    CHECK(crossings.size() >= 2);
    double acc = 0; for (size_t i = 1; i < crossings.size(); i++) acc += crossings[i] - crossings[i - 1];
    double period = (float)(acc / (crossings.size() - 1) * 2), expected = 2 * 3.1415926 / std::sqrt(1e9);
    CHECK_NEAR(expected, period, expected / 1000);
}

static void largeDeformation()          // tVoxelyze.h:445-522
{
    CVoxelyze sim(0.001);
    CVX_Material* m = sim.addMaterial(1e6, 1e3);
    m->setInternalDamping(1.0); m->setGlobalDamping(0.2f);
    sim.setVoxel(m, 0, 0, 0)->external()->setFixedAll();
    CVX_Voxel* b = sim.setVoxel(m, 1, 0, 0);
    float ts = sim.recommendedTimeStep();
    CVX_Link* l = sim.link(0, 0, 0, CVX_Voxel::X_POS);
    CHECK(l != NULL && l->isSmallAngle());
    b->external()->setForce(-0.2f, 0.0f, 0.2f);
    for (int k = 0; k < 200; k++) sim.doTimeStep(ts);
    CHECK_NEAR(9.5587e-4, b->position().z, 1e-7);
    CHECK(!l->isSmallAngle());
}

static void doubleBondCantilever()      // tVoxelyze.h:525-575 (z direction)
{
    CVoxelyze sim(0.001);
    CVX_Material* m = sim.addMaterial(1e6, 1e3);
    m->setInternalDamping(1.0); m->setGlobalDamping(0.1f);
    sim.setVoxel(m, 0, 0, 0)->external()->setFixedAll();
    sim.setVoxel(m, 1, 0, 0);
    CVX_Voxel* c = sim.setVoxel(m, 2, 0, 0);
    float ts = sim.recommendedTimeStep();
    for (int i = 0; i < 3; i += 2) {
        c->external()->setFixed(!(i == 0), !(i == 1), !(i == 2), true, !(i == 2), !(i == 1));
        Vec3D<float> f(0, 0, 0); f[i] = 5e-6f;
        c->external()->setForce(f);
        float expected = i == 0 ? 1e-8f : 1.6e-7f;
        sim.resetTime();
        int streak = 0, k;
        for (k = 0; k < 3000; k++) {
            sim.doTimeStep(ts);
            float value = (float)((c->position() - Vec3D<double>(0.002, 0, 0))[i]);
            if (std::fabs(value - expected) < std::fabs(expected) * 1e-4) streak++; else streak = 0;
            if (streak == 10) break;
        }
        CHECK(k < (i == 0 ? 200 : 500));
    }
}

static void impulse()                   // tVoxelyze.h:577-610
{
    CVoxelyze sim(0.001);
    CVX_Material* m = sim.addMaterial(1e6, 1e3);
    m->setInternalDamping(1.0); m->setGlobalDamping(0.05f);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 2; j++) { CVX_Voxel* v = sim.setVoxel(m, i, j, 0); if (i == 0) v->external()->setFixedAll(); }
    float ts = sim.recommendedTimeStep();
    for (int k = 0; k < 1000; k++) {
        CVX_Voxel* v = sim.voxel(3, 0, 0);
        if (k == 10) v->external()->setForce(0, 0, 100);
        else if (k == 11) v->external()->setForce(0, 0, 0);
        sim.doTimeStep(ts);
    }
    CHECK_NEAR(0.0f, (float)sim.voxel(3, 0, 0)->position().z, 1e-5);
}

static void multiMaterial()             // tVoxelyze.h:614-684
{
    {
        CVoxelyze sim(0.001);
        CVX_Material* soft = sim.addMaterial(1e6, 1e3); CVX_Material* stiff = sim.addMaterial(1e9, 1e3);
        soft->setGlobalDamping(0.03f); stiff->setGlobalDamping(0.03f); soft->setInternalDamping(0.01f); stiff->setInternalDamping(0.01f);
        for (int i = 0; i < 2; i++) {
            CVX_Voxel* a = sim.setVoxel(soft, 2 * i, 0, 0); CVX_Voxel* b = sim.setVoxel(stiff, 2 * i + 1, 0, 0);
            if (i == 0) a->external()->setFixedAll();
            if (i == 1) b->external()->setForce(1e-3f, 0, 0);
        }
        float ts = sim.recommendedTimeStep();
        for (int i = 0; i < 800; i++) sim.doTimeStep(ts);
        CHECK_FLOAT_EQ(1.5015e-6f, (float)(sim.voxel(3, 0, 0)->position().x - 0.003));
    }
    {
        CVoxelyze sim(0.001);
        CVX_Material* soft = sim.addMaterial(1e6, 1e3); CVX_Material* stiff = sim.addMaterial(1e9, 1e3);
        soft->setGlobalDamping(0.01f); stiff->setGlobalDamping(0.01f); soft->setInternalDamping(1.0f); stiff->setInternalDamping(1.0f);
        for (int i = 0; i < 8; i++) {
            CVX_Voxel* v = sim.setVoxel((i / 2) % 2 == 0 ? soft : stiff, i, 0, 0);
            if (i == 0) v->external()->setFixedAll();
            if (i == 7) v->external()->setForce(1e-3f, 0, 0);
        }
        float ts = sim.recommendedTimeStep();
        for (int i = 0; i < 10000; i++) sim.doTimeStep(ts);
        CHECK_NEAR(3.5035e-6f, (float)(sim.voxel(7, 0, 0)->position().x - 0.007), 1e-10);
    }
}

static void deformableMaterial()        // tVoxelyze.h:845-889
{
    CVoxelyze sim(0.001);
    CVX_Material* m = sim.addMaterial(1e6, 1e3);
    m->setModelBilinear(1e6, 5e5, 1e5);
    m->setInternalDamping(1.0f); m->setGlobalDamping(0.2f);
    for (int i = 0; i < 5; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
        CVX_Voxel* v = sim.setVoxel(m, i, j, k);
        if (i == 0) v->external()->setFixedAll();
        if (i == 4) v->external()->setForce(Vec3D<float>(0.2f, 0.0f, 0.0f));
    }
    float ts = sim.recommendedTimeStep();
    for (int i = 0; i < 400; i++) sim.doTimeStep(ts);
    CHECK(sim.link(1, 1, 1, CVX_Voxel::X_POS)->isYielded());
    for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) sim.voxel(4, j, k)->external()->setForce(Vec3D<float>(0, 0, 0));
    for (int i = 0; i < 250; i++) sim.doTimeStep(ts);
    CHECK_NEAR(4e-4, (float)(sim.voxel(4, 1, 1)->position().x - 0.004), 1e-7);
}

static void replaceMaterialMidRun()     // tVoxelyze.h:946-986
{
    CVoxelyze sim(0.001);
    CVX_Material* a = sim.addMaterial(1e6, 1e3); a->setInternalDamping(1.0f); a->setGlobalDamping(0.08f);
    CVX_Material* b = sim.addMaterial(1e7, 1e3); b->setInternalDamping(1.0f); b->setGlobalDamping(0.08f);
    for (int i = 0; i < 5; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        CVX_Voxel* v = sim.setVoxel(a, i, j, k);
        if (i == 0) v->external()->setFixedAll();
        else if (i == 4) v->external()->setForce(0, 0, 1e-6f);
    }
    float ts = sim.recommendedTimeStep();
    for (int l = 0; l < 2000; l++) {
        sim.doTimeStep(ts);
        if (l == 150) {
            for (int i = 0; i < 5; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) if (i % 2 == 1) sim.setVoxel(b, i, j, k);
            ts = sim.recommendedTimeStep();
        }
    }
    CHECK_NEAR(3.8167e-8, sim.voxel(4, 0, 0)->position().z, 1e-10);
}

static void temperatureBimorph()        // tVoxelyze.h:989-1032
{
    CVoxelyze sim(0.001);
    CVX_Material* a = sim.addMaterial(1e6, 1e3); a->setInternalDamping(1.0f); a->setGlobalDamping(0.15f); a->setCte(0.01f);
    CVX_Material* b = sim.addMaterial(1e7f, 1e3f); b->setInternalDamping(1.0f); b->setGlobalDamping(0.15f);
    for (int i = 0; i < 3; i++) {
        CVX_Voxel* v1 = sim.setVoxel(a, i, 0, 0); CVX_Voxel* v2 = sim.setVoxel(b, i, 0, 1);
        if (i == 0) { v1->external()->setFixedAll(); v2->external()->setFixedAll(); }
    }
    sim.setAmbientTemperature(5, true);
    float ts = sim.recommendedTimeStep();
    for (int l = 0; l < 500; l++) sim.doTimeStep(ts);
    CHECK_NEAR(2.55e-5, sim.voxel(2, 0, 0)->position().z, 1e-8);
    sim.resetTime();
    sim.setAmbientTemperature(-5, true);
    for (int l = 0; l < 500; l++) sim.doTimeStep(ts);
    CHECK_NEAR(-2.591e-5, sim.voxel(2, 0, 0)->position().z, 1e-8);
}

static void staticFriction()            // tVoxelyze.h:1034-1112
{
    const double vSize = 0.001; const float density = 1e3f;
    CVoxelyze sim(vSize);
    sim.enableFloor(true);
    sim.setGravity();
    CVX_Material* m = sim.addMaterial(1e6, density);
    const float normalForce = (float)(density * vSize * vSize * vSize * 9.80665);
    CVX_Voxel* v = sim.setVoxel(m, 0, 0, 0);
    m->setStaticFriction(1.0f); m->setKineticFriction(0.1f); m->setGlobalDamping(1.0f);
    float ts = sim.recommendedTimeStep();
    struct Trial { float g, mu, push; bool moves; };
    const Trial trials[] = {{1, 1, 0.9f, false}, {1, 1, 1.1f, true}, {2, 1, 1.9f, false}, {2, 1, 2.1f, true}, {1, 2, 1.9f, false}, {1, 2, 2.1f, true}};
    for (const Trial& t : trials) {
        sim.setGravity(t.g); m->setStaticFriction(t.mu);
        for (int l = 0; l < 50; l++) sim.doTimeStep(ts);
        v->external()->setForce(t.push * normalForce, 0.0f, 0.0f);
        for (int l = 0; l < 10; l++) sim.doTimeStep(ts);
        if (t.moves) CHECK(v->position().x != 0.0); else CHECK(v->position().x == 0.0);
        v->external()->setForce(0.0f, 0.0f, 0.0f);
        sim.resetTime();
    }
}

static void kineticFriction()           // tVoxelyze.h:1116-1167
{
    const double vSize = 0.001; const float density = 1e3f;
    CVoxelyze sim(vSize);
    sim.enableFloor(true); sim.setGravity();
    CVX_Material* m = sim.addMaterial(1e6, density);
    const double mass = density * vSize * vSize * vSize; const float normalForce = (float)(mass * 9.80665);
    CVX_Voxel* v = sim.setVoxel(m, 0, 0, 0);
    m->setStaticFriction(1.0f); m->setKineticFriction(0.1f); m->setGlobalDamping(1.0f);
    float ts = sim.recommendedTimeStep();
    for (int l = 0; l < 50; l++) sim.doTimeStep(ts);
    m->setGlobalDamping(0.0001f);
    const float push = 2.0f * normalForce;
    v->external()->setForce(push, 0.0f, 0.0f);
    double last = 0, vel = 0, energy = 0;
    for (int l = 0; l < 10; l++) {
        sim.doTimeStep(ts);
        double cur = v->position().x;
        vel = (cur - last) / ts;
        energy += (l == 9 ? 0.5f : 1.0f) * (push - m->kineticFriction() * normalForce) * (cur - last);
        last = cur;
    }
    CHECK_NEAR(energy, 0.5 * mass * vel * vel, 2e-16);
}

static void collisionsHoldUp()          // tVoxelyze.h:1169-1197
{
    CVoxelyze sim(0.001);
    sim.enableFloor(true); sim.setGravity();
    CVX_Material* m = sim.addMaterial(1e6, 1e6f);
    m->setGlobalDamping(0.0f);
    sim.setVoxel(m, 0, 0, 0)->external()->setFixedAll();
    sim.setVoxel(m, 0, 0, 2);
    sim.enableCollisions();
    float ts = sim.recommendedTimeStep();
    for (int l = 0; l < 150; l++) sim.doTimeStep(ts);
    CHECK(sim.voxel(0, 0, 2)->position().z > 0.001);
    CHECK(sim.collisionList()->size() == 1);
}

static void stateInfoBasics()           // Voxelyze.cpp:752-800 through the public call
{
    CVoxelyze sim(0.001);
    CVX_Material* m = sim.addMaterial(1e6, 1e3);
    sim.setVoxel(m, 0, 0, 0)->external()->setFixedAll();
    CVX_Voxel* b = sim.setVoxel(m, 1, 0, 0);
    b->external()->setForce(1e-3f, 0, 0);
    float ts = sim.recommendedTimeStep();
    for (int l = 0; l < 300; l++) sim.doTimeStep(ts);
    CHECK_NEAR(sim.stateInfo(CVoxelyze::DISPLACEMENT, CVoxelyze::MAX), (float)b->displacement().Length(), 1e-12);
    CHECK_NEAR(sim.stateInfo(CVoxelyze::ENG_STRAIN, CVoxelyze::MAX), 1e-3, 1e-6);
    CHECK_NEAR(sim.stateInfo(CVoxelyze::ENG_STRESS, CVoxelyze::AVERAGE), 1e3, 1.0);
    CHECK_NEAR(sim.stateInfo(CVoxelyze::MASS, CVoxelyze::TOTAL), 2e-6, 1e-12);
    CHECK_NEAR(sim.stateInfo(CVoxelyze::STRAIN_ENERGY, CVoxelyze::TOTAL), 0.5 * 1e-3 * 1e-6, 1e-11);
}

// ---- material classes (stand-alone objects, no simulation / no device) ---------------------------
static void materialModels()            // tVX_Material.h, tVX_MaterialLink.h:3-69
{
    CVX_Material mat;
    CHECK(!mat.setModelLinear(-1.0f));
    CHECK(std::string(mat.lastError()).find("Young") != std::string::npos);
    CHECK(mat.setModelBilinear(1e6f, 5e5f, 1e5f, 2e5f));
    CHECK(!mat.isModelLinear());
    CHECK_FLOAT_EQ(0.1f, mat.yieldStress() / mat.youngsModulus());
    CHECK_FLOAT_EQ(5e4f, mat.stress(0.05f));
    CHECK_FLOAT_EQ(1.5e5f, mat.stress(0.2f));
    CHECK_FLOAT_EQ(5e5f, mat.modulus(0.2f));
    CHECK(mat.isYielded(0.11f) && !mat.isYielded(0.09f) && mat.isFailed(0.31f) && !mat.isFailed(0.29f));
    mat.setPoissonsRatio(0.7f);
    CHECK(mat.poissonsRatio() < 0.5f);
    mat.setDensity(-5.0f);
    CHECK(mat.density() > 0);

    CVX_MaterialVoxel a, b;
    CHECK(a.setModelLinear(1.0f)); CHECK(b.setModelLinear(10.0f));
    CVX_MaterialLink ab(&a, &b);
    CHECK_FLOAT_EQ(20.0f / 11.0f, ab.youngsModulus());
    CHECK(!ab.isFailed(1));
    CVX_MaterialVoxel c, d;
    CHECK(c.setModelLinear(10.0f, 30.0f)); CHECK(d.setModelLinear(1.0f, 2.0f));
    CVX_MaterialLink cd(&c, &d);
    CHECK_FLOAT_EQ(20.0f / 11.0f, cd.youngsModulus());
    CHECK(!cd.isFailed(1.0f) && cd.isFailed(1.2f));
    CVX_MaterialVoxel e, f;
    CHECK(e.setModelBilinear(1.0f, 0.5f, 1.0f, 2.0f)); CHECK(f.setModelBilinear(2.0f, 1.0f, 4.0f, 6.0f));
    CVX_MaterialLink ef(&e, &f);
    CHECK_FLOAT_EQ(4.0f / 3.0f, ef.youngsModulus());
    CHECK_FLOAT_EQ(4.0f / 3.0f, ef.modulus(0.5));
    CHECK_FLOAT_EQ(2.0f / 2.5f, ef.modulus(1.5));
    CHECK_FLOAT_EQ(0.0f, ef.modulus(2.5));
    CHECK(!ef.isFailed(1.8) && ef.isFailed(1.9));
}

static void voxelExternals()            // tVX_Voxel.h:3-68
{
    CVX_MaterialVoxel mat;
    CVX_Voxel vox(&mat, 0, 0, 0);
    CHECK(!vox.externalExists());
    CHECK(!vox.external()->isFixed(X_TRANSLATE) && !vox.external()->isFixed(Z_ROTATE));
    CHECK(vox.externalExists());
    vox.external()->setFixedAll();
    CHECK(vox.external()->isFixedAll() && vox.external()->isFixed(Y_TRANSLATE) && vox.external()->isFixed(X_ROTATE));
    vox.external()->setFixedAll(false);
    CHECK(!vox.external()->isFixedAny());
    vox.external()->setFixed(X_TRANSLATE);
    CHECK(vox.external()->isFixed(X_TRANSLATE) && !vox.external()->isFixed(Y_TRANSLATE));
    vox.external()->setDisplacement(Z_ROTATE, 0.25);
    CHECK(vox.external()->isFixed(Z_ROTATE));
    CHECK_NEAR(vox.external()->rotation().z, 0.25, 0);
    CHECK_NEAR(vox.external()->rotationQuat().w, std::cos(0.125), 1e-15);
}

static void largeDeformationDamping()    // tVoxelyze.h:445-473
{
    CVoxelyze Sim(0.001);
    CVX_Material* pMat1 = Sim.addMaterial(1e6f, 1e3f);
    pMat1->setInternalDamping(1.0);
    pMat1->setGlobalDamping(0.2f);
    CVX_Voxel* pV1 = Sim.setVoxel(pMat1, 0, 0, 0);
    CVX_Voxel* pV2 = Sim.setVoxel(pMat1, 1, 0, 0);
    float ts = Sim.recommendedTimeStep();
    pV1->external()->setFixedAll();
    pV2->external()->setForce(-0.2f, 0.0f, 0.2f);
    for (int k = 0; k < 200; k++) Sim.doTimeStep(ts);
    CHECK_NEAR(9.5587e-4, pV2->position().z, 1e-7);
}

static void poissonsLarge()              // tVoxelyze.h:715-769, 9x2x2, nu = 0.3: corner offsets, size(), reaction forces
{
    CVoxelyze Sim(0.001);
    CVX_Material* pMat1 = Sim.addMaterial(1e6f, 1e3f);
    float mu = 0.3f;
    pMat1->setPoissonsRatio(mu);
    pMat1->setInternalDamping(1.0f);
    pMat1->setGlobalDamping(0.2f);
    for (int i = 0; i < 9; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        CVX_Voxel* pV = Sim.setVoxel(pMat1, i, j, k);
        if (i == 0) pV->external()->setFixedAll();
        if (i == 8) pV->external()->setDisplacementAll(Vec3D<>(1e-3f, 0, 0));
    }
    float ts = Sim.recommendedTimeStep();
    for (int i = 0; i < 300; i++) Sim.doTimeStep(ts);
    CHECK_NEAR(5e-4, (float)(Sim.voxel(4, 0, 0)->position().x - 0.004), 1e-7);
    Vec3D<float> curSize = 2 * Sim.voxel(4, 0, 0)->cornerOffset(CVX_Voxel::PPP);
    Vec3D<float> curStrain = curSize - Vec3D<float>(0.001f, 0.001f, 0.001f);
    curStrain /= 0.001f;
    CHECK_NEAR(1.306e-1, curStrain.x, 1e-3);
    CHECK_NEAR(-4.048e-2, curStrain.y, 1e-5);
    CHECK_NEAR(-4.048e-2, curStrain.z, 1e-5);
    float eHat = pMat1->youngsModulus() / ((1 - 2 * mu) * (1 + mu));
    float sigmaX = eHat * ((1 - mu) * curStrain.x + mu * (curStrain.y + curStrain.z));
    Vec3D<float> curSize2 = Sim.voxel(4, 0, 0)->size();
    float fX = sigmaX * 4 * curSize2.y * curSize2.z;
    float totalForce = 0;
    for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) totalForce += Sim.voxel(8, j, k)->externalForce().x;
    CHECK_NEAR(totalForce, fX, 0.01);
    CHECK(totalForce > 0.05f);                                    // the reaction really is there (about 0.5 N)
}

static void poissonsHigh()               // tVoxelyze.h:771-803, 5x3x3, nu = 0.495
{
    CVoxelyze Sim(0.001);
    CVX_Material* pMat1 = Sim.addMaterial(1e6f, 1e3f);
    pMat1->setPoissonsRatio(0.495f);
    pMat1->setInternalDamping(1.0f);
    pMat1->setGlobalDamping(2.0f);
    for (int i = 0; i < 5; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
        CVX_Voxel* pV = Sim.setVoxel(pMat1, i, j, k);
        if (i == 0) pV->external()->setFixedAll();
        if (i == 4) pV->external()->setDisplacementAll(Vec3D<>(1e-5f, 0, 0));
    }
    float ts = Sim.recommendedTimeStep();
    for (int i = 0; i < 300; i++) Sim.doTimeStep(ts);
    CHECK_NEAR(5e-6, (float)(Sim.voxel(2, 1, 1)->position().x - 0.002), 1e-9);
}

static void poissonsMixed()              // tVoxelyze.h:806-842, 7x3x3, nu = 0.3 around a nu = 0 core
{
    CVoxelyze Sim(0.001);
    CVX_Material* pMat1 = Sim.addMaterial(1e6f, 1e3f);
    pMat1->setPoissonsRatio(0.3f); pMat1->setInternalDamping(1.0f); pMat1->setGlobalDamping(0.3f);
    CVX_Material* pMat2 = Sim.addMaterial(1e6f, 1e3f);
    pMat2->setPoissonsRatio(0.0f); pMat2->setInternalDamping(1.0f); pMat2->setGlobalDamping(0.3f);
    for (int i = 0; i < 7; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
        CVX_Voxel* pV = Sim.setVoxel((i > 1 && i < 6) ? pMat2 : pMat1, i, j, k);
        if (i == 0) pV->external()->setFixedAll();
        if (i == 6) pV->external()->setDisplacementAll(Vec3D<>(1e-3f, 0, 0));
    }
    float ts = Sim.recommendedTimeStep();
    for (int i = 0; i < 300; i++) Sim.doTimeStep(ts);
    CHECK_NEAR(5e-4, (float)(Sim.voxel(3, 1, 1)->position().x - 0.003), 5e-6);
}

// ---- *.vxl.json (Voxelyze.cpp:61-241, VX_Material.cpp:75-163): a model with two materials and three kinds of
// externals; written by one implementation, it must load into the same model in either.
static void buildJsonModel(CVoxelyze& Vx)
{
    CVX_Material* a = Vx.addMaterial(1e6f, 1000.0f);
    a->setName("soft"); a->setColor(200, 30, 40); a->setStaticFriction(0.7f); a->setKineticFriction(0.3f); a->setGlobalDamping(0.02f);
    a->setCte(0.01f); a->setModelLinear(1e6f, 2e5f);
    CVX_Material* b = Vx.addMaterial(5e6f, 1500.0f);
    b->setPoissonsRatio(0.3f); b->setInternalDamping(0.8f); b->setCollisionDamping(0.5f); b->setExternalScaleFactor(Vec3D<double>(1.0, 1.25, 0.75));
    for (int k = 0; k < 2; k++) for (int j = 0; j < 2; j++) for (int i = 0; i < 5; i++) Vx.setVoxel((i + j + k) % 3 == 0 ? b : a, i, j - 1, k + 3);
    for (int k = 0; k < 2; k++) for (int j = 0; j < 2; j++) Vx.voxel(0, j - 1, k + 3)->external()->setFixedAll();
    for (int k = 0; k < 2; k++) for (int j = 0; j < 2; j++) Vx.voxel(4, j - 1, k + 3)->external()->setForce(0.0f, 0.001f, -0.002f);
    Vx.voxel(2, 0, 4)->external()->setDisplacement(Z_TRANSLATE, 1e-4);
    Vx.voxel(2, 0, 4)->external()->setMoment(1e-6f, 0.0f, 0.0f);
    Vx.setGravity(1.0f); Vx.enableFloor(true);
}
static void printJsonDigest(CVoxelyze& Vx)
{
    printf("voxelSize %.17g materials %d voxels %d\n", Vx.voxelSize(), Vx.materialCount(), Vx.voxelCount());
    printf("env gravity %g floor %d collisions %d ambient %g\n", Vx.gravity(), (int)Vx.isFloorEnabled(), (int)Vx.isCollisionsEnabled(), Vx.ambientTemperature());
    for (int i = 0; i < Vx.materialCount(); i++) {
        CVX_Material* m = Vx.material(i);
        printf("mat %d '%s' rgba %d %d %d %d linear %d E %.9g fail %.9g rho %.9g nu %.9g cte %.9g mu %.9g %.9g zeta %.9g %.9g %.9g scale %.17g %.17g %.17g\n", i, m->name(),
               m->red(), m->green(), m->blue(), m->alpha(), (int)m->isModelLinear(), m->youngsModulus(), m->failureStress(), m->density(), m->poissonsRatio(), m->cte(),
               m->staticFriction(), m->kineticFriction(), m->internalDamping(), m->globalDamping(), m->collisionDamping(),
               m->externalScaleFactor().x, m->externalScaleFactor().y, m->externalScaleFactor().z);
    }
    for (int i = 0; i < Vx.voxelCount(); i++) {
        CVX_Voxel* v = Vx.voxel(i);
        int mi = -1;
        for (int k = 0; k < Vx.materialCount(); k++) if (Vx.material(k) == v->material()) mi = k;
        printf("vox %d at %d %d %d mat %d", i, v->indexX(), v->indexY(), v->indexZ(), mi);
        if (v->externalExists() && !v->external()->isEmpty()) {
            CVX_External* e = v->external();
            printf(" fixed %d%d%d%d%d%d T %.17g %.17g %.17g R %.17g %.17g %.17g F %.9g %.9g %.9g M %.9g %.9g %.9g", (int)e->isFixed(X_TRANSLATE), (int)e->isFixed(Y_TRANSLATE),
                   (int)e->isFixed(Z_TRANSLATE), (int)e->isFixed(X_ROTATE), (int)e->isFixed(Y_ROTATE), (int)e->isFixed(Z_ROTATE), e->translation().x, e->translation().y, e->translation().z,
                   e->rotation().x, e->rotation().y, e->rotation().z, e->force().x, e->force().y, e->force().z, e->moment().x, e->moment().y, e->moment().z);
        }
        printf("\n");
    }
}
static std::string g_tmpdir = "/tmp";
static void jsonRoundTrip()             // Voxelyze.h:70,77-78: save, load into a second object, same model and same motion
{
    CVoxelyze A(0.002);
    buildJsonModel(A);
    std::string path = g_tmpdir + "/dropin_roundtrip.vxl.json";
    CHECK(A.saveJSON(path.c_str()));
    CVoxelyze B(path.c_str());
    CHECK(B.voxelSize() == 0.002);
    CHECK(B.materialCount() == 2); CHECK(B.voxelCount() == 20);
    if (B.materialCount() != 2 || B.voxelCount() != 20) return;
    CHECK(std::string(B.material(0)->name()) == "soft"); CHECK(B.material(0)->red() == 200);
    CHECK_FLOAT_EQ(B.material(0)->failureStress(), 2e5f); CHECK_FLOAT_EQ(B.material(0)->cte(), 0.01f);
    CHECK_FLOAT_EQ(B.material(1)->youngsModulus(), 5e6f); CHECK_FLOAT_EQ(B.material(1)->density(), 1500.0f);
    CHECK_FLOAT_EQ(B.material(1)->poissonsRatio(), 0.3f); CHECK(B.material(1)->externalScaleFactor().y == 1.25);
    for (int i = 0; i < 20; i++) {
        CHECK(B.voxel(i)->indexX() == A.voxel(i)->indexX() && B.voxel(i)->indexY() == A.voxel(i)->indexY() && B.voxel(i)->indexZ() == A.voxel(i)->indexZ());
        CHECK((B.voxel(i)->material() == B.material(0)) == (A.voxel(i)->material() == A.material(0)));
    }
    CHECK(B.voxel(0, -1, 3)->external()->isFixedAll());
    CHECK_FLOAT_EQ(B.voxel(4, 0, 4)->external()->force().z, -0.002f);
    CHECK(B.voxel(2, 0, 4)->external()->isFixed(Z_TRANSLATE) && !B.voxel(2, 0, 4)->external()->isFixed(X_TRANSLATE));
    CHECK(B.voxel(2, 0, 4)->external()->translation().z == 1e-4);
    CHECK_FLOAT_EQ(B.voxel(2, 0, 4)->external()->moment().x, 1e-6f);
    CHECK(B.gravity() == 0.0f && !B.isFloorEnabled());      // written, never read back (Voxelyze.cpp:168-171 vs :95-161)
    B.setGravity(1.0f); B.enableFloor(true);
    float dt = A.recommendedTimeStep();
    CHECK_FLOAT_EQ(B.recommendedTimeStep(), dt);
    for (int i = 0; i < 300; i++) { A.doTimeStep(dt); B.doTimeStep(dt); }
    for (int i = 0; i < 20; i++) {
        CHECK(A.voxel(i)->position().x == B.voxel(i)->position().x && A.voxel(i)->position().y == B.voxel(i)->position().y && A.voxel(i)->position().z == B.voxel(i)->position().z);
    }
    CHECK(std::fabs(A.voxel(19)->position().z - 4 * 0.002) > 1e-7);    // it did move
}

// A plastically bent beam whose middle voxel gets a new material mid-run: only that voxel's links restart
// (src/Voxelyze.cpp:485-498), every other link keeps its plastic memory; then collisions are switched on mid-run
// (a change of device layout for the facade).  Prints the final state for a cross-implementation comparison.
static void printEditScenario()
{
    CVoxelyze Vx(0.001);
    CVX_Material* m = Vx.addMaterial(1e6f, 1000.0f);
    m->setModelBilinear(1e6f, 1e5f, 2e4f); m->setGlobalDamping(0.05f);
    CVX_Material* m2 = Vx.addMaterial(1e6f, 1000.0f);
    m2->setModelBilinear(1e6f, 1e5f, 2e4f); m2->setGlobalDamping(0.05f);
    for (int k = 0; k < 2; k++) for (int j = 0; j < 2; j++) for (int i = 0; i < 8; i++) Vx.setVoxel(m, i, j, k);
    for (int k = 0; k < 2; k++) for (int j = 0; j < 2; j++) {
        Vx.voxel(0, j, k)->external()->setFixedAll();
        Vx.voxel(7, j, k)->external()->setForce(0.0f, 0.0f, -0.004f);
    }
    float dt = Vx.recommendedTimeStep();
    for (int i = 0; i < 1500; i++) Vx.doTimeStep(dt);
    int yielded = 0;
    for (int i = 0; i < Vx.linkCount(); i++) if (Vx.link(i)->isYielded()) yielded++;
    printf("after loading: yielded links %d tip z %.12e\n", yielded, Vx.voxel(7, 0, 0)->position().z);
    Vx.setVoxel(m2, 3, 0, 1);                                        // swap one voxel's material mid-run
    for (int k = 0; k < 2; k++) for (int j = 0; j < 2; j++) Vx.voxel(7, j, k)->external()->setForce(0.0f, 0.0f, 0.0f);
    for (int i = 0; i < 800; i++) Vx.doTimeStep(dt);
    Vx.enableCollisions(true);                                       // mid-run
    for (int i = 0; i < 700; i++) Vx.doTimeStep(dt);
    yielded = 0;
    for (int i = 0; i < Vx.linkCount(); i++) if (Vx.link(i)->isYielded()) yielded++;
    printf("after unloading: yielded links %d\n", yielded);
    for (int i = 0; i < Vx.voxelCount(); i++) {
        Vec3D<double> p = Vx.voxel(i)->position();
        printf("vox %d %.12e %.12e %.12e\n", i, p.x, p.y, p.z);
    }
    // a new voxel size scales positions, halts motion and restarts the links (Voxelyze.cpp:643-668)
    Vx.setVoxelSize(0.0015);
    float dt2 = Vx.recommendedTimeStep();
    for (int i = 0; i < 200; i++) Vx.doTimeStep(dt2);
    for (int i = 0; i < Vx.voxelCount(); i += 5) {
        Vec3D<double> q = Vx.voxel(i)->position();
        printf("vox %d %.12e %.12e %.12e\n", 200 + i, q.x, q.y, q.z);
    }
}

// include/VX_Voxel.h:119-120, 130: the floor can be switched per voxel; dampingMultiplier follows the last time step
static void perVoxelFloorAndDampingMultiplier()
{
    CVoxelyze Vx(0.001);
    CVX_Material* m = Vx.addMaterial(1e6f, 1e3f);
    m->setGlobalDamping(0.002f); m->setCollisionDamping(1.0f); m->setInternalDamping(0.7f);      // (more global damping = a terminal velocity of a few mm/s)
    CVX_Voxel* a = Vx.setVoxel(m, 0, 0, 0);
    CVX_Voxel* b = Vx.setVoxel(m, 4, 0, 0);
    CVX_Voxel* c = Vx.setVoxel(m, 8, 0, 0);
    Vx.setGravity(1.0f); Vx.enableFloor(true);
    b->enableFloor(false);
    CHECK(a->isFloorEnabled() && !b->isFloorEnabled() && c->isFloorEnabled());
    float dt = Vx.recommendedTimeStep();
    for (int i = 0; i < 8000; i++) Vx.doTimeStep(dt);
    CHECK_FLOAT_EQ(a->dampingMultiplier(), 2 * sqrtf((float)(1e-9 * 1e3)) * 0.7f / dt);
    CHECK(a->position().z > -1e-4 && c->position().z > -1e-4);         // held up by the floor
    CHECK(b->position().z < -5e-4);                                     // fell through it
    CHECK_NEAR(a->position().z, c->position().z, 1e-15);
    const double zb = b->position().z;
    Vx.enableFloor(true);                                               // src/Voxelyze.cpp:604-610: every voxel follows the simulation again
    CHECK(b->isFloorEnabled());
    c->enableFloor(false);
    for (int i = 0; i < 8000; i++) Vx.doTimeStep(dt);
    CHECK(b->position().z > zb);                                        // pushed back up by the floor spring
    CHECK(c->position().z < -5e-4);
}

#ifndef DROPIN_REFERENCE
// include/Voxelyze.h:80, include/VX_LinearSolver.h:42-57: static solve of a cantilever.  (The reference's needs PARDISO and is a
// no-op without it, so this one runs on the facade only; tests/test_static_solve.py pins the algebra to the reference's.)
static void linearSolveCantilever()
{
    CVoxelyze Vx(0.001);
    CVX_Material* m = Vx.addMaterial(1e6f, 1e3f);
    m->setGlobalDamping(0.05f);
    const int n = 12;
    for (int i = 0; i < n; i++) Vx.setVoxel(m, i, 0, 0);
    Vx.voxel(0)->external()->setFixedAll();
    const float F = 1e-6f;
    Vx.voxel(n - 1)->external()->setForce(0, 0, -F);
    CVX_LinearSolver solver(&Vx);
    solver.relTolerance = 1e-13;
    CHECK(solver.solve());
    CHECK(solver.iterations > 0 && solver.residual <= 1e-13 && solver.progressTick == 90);
    // a chain of Euler-Bernoulli beam elements is exact at the nodes: F L^3 / (3 E I) with I = h^4 / 12, slope F L^2 / (2 E I) -- up to
    // the float rounding of the beam constants (b1 * 2 b3 - b2^2 cancels a third of its digits): 1e-4
    const double L = (n - 1) * 1e-3, EI = 1e6 * 1e-12 / 12.0;
    CHECK_NEAR(Vx.voxel(n - 1)->displacement().z, -F * L * L * L / (3 * EI), 1e-4 * F * L * L * L / (3 * EI));
    CHECK_NEAR(Vx.voxel(n - 1)->orientation().ToRotationVector().y, F * L * L / (2 * EI), 1e-4 * F * L * L / (2 * EI));
    CHECK(Vx.voxel(n - 1)->velocity().Length2() == 0 && Vx.voxel(3)->angularVelocity().Length2() == 0);
    // it is a rest state of the time stepper (small deflection: 0.5 % of the length)
    const double z0 = Vx.voxel(n - 1)->position().z;
    float dt = Vx.recommendedTimeStep();
    for (int i = 0; i < 300; i++) Vx.doTimeStep(dt);
    CHECK_NEAR(Vx.voxel(n - 1)->position().z, z0, 2e-4 * fabs(Vx.voxel(n - 1)->displacement().z));
    // doLinearSolve: same thing through CVoxelyze; a model that is not held reports through the solver object
    Vx.resetTime();
    CHECK(Vx.doLinearSolve());
    CHECK_NEAR(Vx.voxel(n - 1)->position().z, z0, 1e-9 * fabs(z0));
    Vx.voxel(0)->external()->setFixedAll(false);
    CVX_LinearSolver loose(&Vx);
    loose.maxIterations = 2000;
    CHECK(!loose.solve() && !loose.errorMsg.empty());
    CHECK_NEAR(Vx.voxel(n - 1)->position().z, z0, 1e-9 * fabs(z0));     // state untouched
}
#endif

#ifndef DROPIN_REFERENCE
static void copyTakesTheModel()         // Voxelyze.cpp:39-58; in the reference the copied materials come out broken (_sqrtMass negated,
{                                       // VX_MaterialVoxel.cpp:47) and the copy diverges at once, so this can only be checked on the facade
    CVoxelyze A(0.002);
    buildJsonModel(A);
    float dt = A.recommendedTimeStep();
    for (int i = 0; i < 50; i++) A.doTimeStep(dt);
    CVoxelyze B(0.001);
    B = A;                                                           // model only: B starts at rest
    CVoxelyze C(0.002);
    buildJsonModel(C);
    CHECK(B.voxelSize() == 0.002 && B.voxelCount() == C.voxelCount() && B.materialCount() == C.materialCount());
    CHECK(B.gravity() == 1.0f && B.isFloorEnabled());
    for (int i = 0; i < 120; i++) { B.doTimeStep(dt); C.doTimeStep(dt); }
    for (int i = 0; i < B.voxelCount(); i++) CHECK(B.voxel(i)->position().z == C.voxel(i)->position().z && B.voxel(i)->position().x == C.voxel(i)->position().x);
}
static void stateCheckpoint()           // facade extra: saveState / loadState (the reference cannot checkpoint, Voxelyze.h:78)
{
    std::string path = g_tmpdir + "/dropin_state.bin";
    CVoxelyze A(0.002), B(0.002);
    buildJsonModel(A); buildJsonModel(B);
    float dt = A.recommendedTimeStep();
    for (int i = 0; i < 150; i++) A.doTimeStep(dt);
    CHECK(A.saveState(path.c_str()));
    for (int i = 0; i < 100; i++) A.doTimeStep(dt);
    CHECK(B.loadState(path.c_str()));
    for (int i = 0; i < 100; i++) B.doTimeStep(dt);
    for (int i = 0; i < 20; i++) {
        CHECK(A.voxel(i)->position().x == B.voxel(i)->position().x && A.voxel(i)->position().z == B.voxel(i)->position().z);
        CHECK(A.voxel(i)->velocity().z == B.voxel(i)->velocity().z);
    }
    CVoxelyze C(0.002);
    C.setVoxel(C.addMaterial(), 0, 0, 0);
    CHECK(!C.loadState(path.c_str()));                      // another model: refused
}
static void slabbedModel(CVoxelyze& Vx)
{
    CVX_Material* soft = Vx.addMaterial(1e6f, 1e3f); soft->setGlobalDamping(0.01f); soft->setCte(0.01f); soft->setStaticFriction(1.0f); soft->setKineticFriction(0.5f);
    CVX_Material* hard = Vx.addMaterial(4e6f, 2e3f); hard->setGlobalDamping(0.01f);
    Vx.setGravity(1.0f); Vx.enableFloor(true);
    for (int z = 0; z < 12; z++) for (int y = 0; y < 4; y++) for (int x = 0; x < 6; x++) Vx.setVoxel(((x + z) & 1) ? hard : soft, x, y, z);
    for (int z = 4; z < 12; z++) for (int y = 0; y < 4; y++) Vx.voxel(0, y, z)->external()->setFixedAll();
    for (int y = 0; y < 4; y++) Vx.voxel(5, y, 11)->external()->setForce(0.0f, 0.002f, -0.01f);
}
static void slabbedDevices()            // facade extra: setDevices -- the class API on several devices of one process (vx_slabbed_*), bits as on one
{
    CVoxelyze A(0.005), B(0.005);
    slabbedModel(A); slabbedModel(B);
    A.setDevice(0);                                                 // whatever VX_DEVICES says
    B.setDevices(std::vector<int>(3, 0));                           // three slabs (on one GPU here; gpurun --gpus N: one each)
    CHECK(B.isSlabbed() && !A.isSlabbed() && B.linkCount() == A.linkCount());
    float dt = A.recommendedTimeStep();
    CHECK(dt == B.recommendedTimeStep());
    auto same = [&]() {
        bool ok = true;
        for (int i = 0; i < A.voxelCount(); i++) {
            CVX_Voxel *a = A.voxel(i), *b = B.voxel(i);
            ok = ok && a->position() == b->position() && a->orientation() == b->orientation() && a->velocity() == b->velocity() && a->angularVelocity() == b->angularVelocity()
                    && a->temperature() == b->temperature() && a->isFloorStaticFriction() == b->isFloorStaticFriction();
        }
        for (int i = 0; i < A.linkCount(); i += 7) {
            CVX_Link *a = A.link(i), *b = B.link(i);
            ok = ok && a->axialStrain() == b->axialStrain() && a->force(true) == b->force(true) && a->moment(false) == b->moment(false) && a->isSmallAngle() == b->isSmallAngle();
        }
        return ok;
    };
    for (int i = 0; i < 150; i++) {
        float t = 3.0f * sinf(i / 10.0f);
        A.setAmbientTemperature(t); B.setAmbientTemperature(t);
        CHECK(A.doTimeStep(dt) && B.doTimeStep(dt));
    }
    CHECK(same());
    // stateInfo of the whole model while slabbed: voxel quantities reduced per slab and combined, link quantities gathered
    CHECK(A.stateInfo(CVoxelyze::DISPLACEMENT, CVoxelyze::MAX) == B.stateInfo(CVoxelyze::DISPLACEMENT, CVoxelyze::MAX));
    CHECK(A.stateInfo(CVoxelyze::ENG_STRAIN, CVoxelyze::MIN) == B.stateInfo(CVoxelyze::ENG_STRAIN, CVoxelyze::MIN));
    const CVoxelyze::stateInfoType sums[4] = {CVoxelyze::STRAIN_ENERGY, CVoxelyze::KINETIC_ENERGY, CVoxelyze::MASS, CVoxelyze::ENG_STRESS};
    for (int k = 0; k < 4; k++) {
        float a = A.stateInfo(sums[k], CVoxelyze::TOTAL), b = B.stateInfo(sums[k], CVoxelyze::TOTAL);
        CHECK(a != 0 && fabsf(a - b) <= 2e-5f * fabsf(a));
        a = A.stateInfo(sums[k], CVoxelyze::AVERAGE); b = B.stateInfo(sums[k], CVoxelyze::AVERAGE);
        CHECK(fabsf(a - b) <= 2e-5f * fabsf(a));
    }
    CHECK(B.isSlabbed());
    // per-voxel edits reach the owner and the ghost copies across the cuts (voxels of the planes next to a cut: z = 3, 4, 7, 8)
    CVoxelyze* both[2] = {&A, &B};
    for (CVoxelyze* V : both) {
        V->voxel(3, 1, 3)->setTemperature(9.0f); V->voxel(2, 2, 4)->setTemperature(-4.0f); V->voxel(4, 0, 8)->haltMotion();
        V->voxel(1, 1, 0)->enableFloor(false);
        for (int i = 0; i < 60; i++) V->doTimeStep(dt);
    }
    CHECK(same());
    // back to one device with the dynamic state; what a slabbed run cannot answer is available again
    B.setDevices(std::vector<int>(1, 0));
    CHECK(!B.isSlabbed());
    CHECK(float_eq(A.stateInfo(CVoxelyze::DISPLACEMENT, CVoxelyze::MAX), B.stateInfo(CVoxelyze::DISPLACEMENT, CVoxelyze::MAX)));
    for (int i = 0; i < 60; i++) { A.doTimeStep(dt); B.doTimeStep(dt); }
    CHECK(same());
    A.voxel(2, 2, 6)->external()->setForce(0.0f, 0.0f, 0.004f);    // an edit that is still pending when the model moves
    B.voxel(2, 2, 6)->external()->setForce(0.0f, 0.0f, 0.004f);
    B.setDevices(std::vector<int>(2, 0));                           // and out again, mid-run
    for (int i = 0; i < 60; i++) { A.doTimeStep(dt); B.doTimeStep(dt); }
    CHECK(B.isSlabbed() && same());
    B.resetTime(); A.resetTime();
    for (int i = 0; i < 30; i++) { A.doTimeStep(dt); B.doTimeStep(dt); }
    CHECK(same());
    // checkpoint of the slabbed object (one file per slab) into a freshly built one with the same cut
    std::string path = g_tmpdir + "/dropin_slabbed_state";
    CHECK(B.saveState(path.c_str()));
    CVoxelyze C(0.005);
    slabbedModel(C);
    C.voxel(2, 2, 6)->external()->setForce(0.0f, 0.0f, 0.004f);    // the same model, edit included: a state file of another model is refused
    C.setDevices(std::vector<int>(2, 0));
    CHECK(C.loadState(path.c_str()) && C.isSlabbed());
    for (int i = 0; i < 25; i++) { B.doTimeStep(dt); C.doTimeStep(dt); }
    bool resumed = true;
    for (int i = 0; i < B.voxelCount(); i++) resumed = resumed && B.voxel(i)->position() == C.voxel(i)->position() && B.voxel(i)->velocity() == C.voxel(i)->velocity();
    CHECK(resumed);
}
#endif

// the deformed surface mesh of a stepped model written by CVX_MeshRender::saveObj (src/VX_MeshRender.cpp:238-251): the file of
// the facade must equal the reference's line for line
static int meshObj(const char* path)
{
    CVoxelyze Vx(0.002);
    buildJsonModel(Vx);
    float dt = Vx.recommendedTimeStep();
    for (int i = 0; i < 200; i++) Vx.doTimeStep(dt);
    CVX_MeshRender mesh(&Vx);
    mesh.updateMesh(CVX_MeshRender::STATE_INFO, CVoxelyze::KINETIC_ENERGY);
    mesh.saveObj(path);
    return 0;
}
// Poisson's ratio switched on in the middle of a run through the material handle (this used to abort the facade)
static void printPoissonScenario()
{
    CVoxelyze Vx(0.001);
    CVX_Material* m = Vx.addMaterial(1e6f, 1e3f);
    m->setGlobalDamping(0.05f);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) Vx.setVoxel(m, i, j, k);
    for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) { Vx.voxel(0, j, k)->external()->setFixedAll(); Vx.voxel(5, j, k)->external()->setForce(2e-3f, 0, -1e-3f); }
    float dt = Vx.recommendedTimeStep();
    for (int i = 0; i < 300; i++) Vx.doTimeStep(dt);
    m->setPoissonsRatio(0.3f);
    float dt2 = Vx.recommendedTimeStep();
    bool ok = true;
    for (int i = 0; i < 300; i++) ok = Vx.doTimeStep(0.5f * dt2) && ok;
    m->setPoissonsRatio(0.0f);
    for (int i = 0; i < 100; i++) ok = Vx.doTimeStep(0.5f * dt2) && ok;
    printf("ok %d dt %.9e dt2 %.9e\n", (int)ok, dt, dt2);
    for (int i = 0; i < Vx.voxelCount(); i++) { Vec3D<double> q = Vx.voxel(i)->position(); printf("vox %d %.12e %.12e %.12e\n", i, q.x, q.y, q.z); }
}

int main(int argc, char** argv)
{
    if (const char* t = getenv("TMPDIR")) g_tmpdir = t;
    if (argc > 2 && std::string(argv[1]) == "--mesh-obj") return meshObj(argv[2]);
    if (argc > 1 && std::string(argv[1]) == "--poisson-scenario") { printPoissonScenario(); return 0; }
    if (argc > 2 && std::string(argv[1]) == "--json-save") { CVoxelyze Vx(0.002); buildJsonModel(Vx); return Vx.saveJSON(argv[2]) ? 0 : 1; }
    if (argc > 1 && std::string(argv[1]) == "--edit-scenario") { printEditScenario(); return 0; }
    if (argc > 2 && std::string(argv[1]) == "--json-digest") { CVoxelyze Vx(argv[2]); printJsonDigest(Vx); return 0; }
    struct T { const char* name; void (*fn)(); bool device; };
    const T tests[] = {
        {"materialModels", materialModels, false}, {"voxelExternals", voxelExternals, false},
        {"simpleSetup", simpleSetup, true}, {"singleBondFixedFree", singleBondFixedFree, true}, {"singleBondFreeFree", singleBondFreeFree, true},
        {"resetTime", resetTime, true}, {"dampingInternalAndGlobal", dampingInternalAndGlobal, true}, {"combinedDamping", combinedDamping, true},
        {"scaleOfForce", scaleOfForce, true}, {"axialFrequency", axialFrequency, true}, {"largeDeformation", largeDeformation, true},
        {"doubleBondCantilever", doubleBondCantilever, true}, {"impulse", impulse, true}, {"multiMaterial", multiMaterial, true},
        {"deformableMaterial", deformableMaterial, true}, {"replaceMaterialMidRun", replaceMaterialMidRun, true},
        {"temperatureBimorph", temperatureBimorph, true}, {"staticFriction", staticFriction, true}, {"kineticFriction", kineticFriction, true},
        {"collisionsHoldUp", collisionsHoldUp, true}, {"stateInfoBasics", stateInfoBasics, true}, {"jsonRoundTrip", jsonRoundTrip, true},
        {"largeDeformationDamping", largeDeformationDamping, true}, {"poissonsLarge", poissonsLarge, true}, {"poissonsHigh", poissonsHigh, true},
        {"poissonsMixed", poissonsMixed, true}, {"perVoxelFloorAndDampingMultiplier", perVoxelFloorAndDampingMultiplier, true},
#ifndef DROPIN_REFERENCE
        {"stateCheckpoint", stateCheckpoint, true}, {"copyTakesTheModel", copyTakesTheModel, true},
        {"linearSolveCantilever", linearSolveCantilever, true}, {"slabbedDevices", slabbedDevices, true},
#endif
    };
    bool host_only = argc > 1 && std::string(argv[1]) == "--host-only";
    int ran = 0;
    for (const T& t : tests) {
        if (host_only && t.device) continue;
        if (argc > 1 && !host_only && std::string(argv[1]) != t.name) continue;
        g_current = t.name;
        int before = g_failures;
        t.fn();
        printf("%s %s\n", g_failures == before ? "PASS" : "FAILED", t.name);
        ran++;
    }
    printf("%d tests, %d checks, %d failures\n", ran, g_checks, g_failures);
    return g_failures ? 1 : 0;
}
