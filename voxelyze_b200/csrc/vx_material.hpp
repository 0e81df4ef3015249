// vx_material.hpp -- host-side material tables of the product (header only).
//
// Everything the kernels need from a material is a handful of float constants that the
// reference caches inside CVX_Material / CVX_MaterialVoxel / CVX_MaterialLink.  This file
// derives those constants with the same float/double evaluation order as the reference so
// that the uploaded tables are bit-identical to what the reference would hold:
//   stress-strain model        src/VX_Material.cpp:165-226, 283-470
//   voxel mass properties      src/VX_MaterialVoxel.cpp:57-79, include/VX_MaterialVoxel.h:41-57
//   link (pair) material       src/VX_MaterialLink.cpp:45-141
// Used by the C-ABI library (vx_capi.cu) and by the C++ facade (facade/).  Not used by oracle/.
#pragma once

#include <cfloat>
#include <cmath>
#include <string>
#include <vector>

#include "voxelyze_b200.h"

namespace vxm {

// A piecewise-linear true stress/strain model plus the scalar properties of CVX_Material.
struct Material {
    // model (VX_Material.h:107-116)
    bool  linear = true;
    float E = 1.0f;
    float sigma_yield = -1.0f, sigma_fail = -1.0f, eps_yield = -1.0f, eps_fail = -1.0f;
    std::vector<float> eps, sig;          // data points, [0] is always (0,0)
    // scalars (VX_Material.h:115-123)
    float nu = 0.0f, rho = 1.0f, cte = 0.0f, mu_s = 0.0f, mu_k = 0.0f;
    float zeta_int = 1.0f, zeta_glob = 0.0f, zeta_coll = 0.0f;
    double ext_scale[3] = {1.0, 1.0, 1.0};
    // derived
    float e_hat = 1.0f;
    std::string error;

    bool yielded(float s) const { return eps_yield != -1.0f && s > eps_yield; }
    bool failed(float s) const { return eps_fail != -1.0f && s > eps_fail; }
    void refresh_e_hat() { e_hat = E / ((1 - 2 * nu) * (1 + nu)); }

    bool model_linear(float youngs, float fail_stress = -1.0f)
    {
        if (youngs <= 0) return reject("Young's modulus must be positive");
        if (fail_stress != -1.0f && fail_stress <= 0) return reject("Failure stress must be positive");
        float top_sig = fail_stress;
        if (top_sig == -1) top_sig = 1000000;          // arbitrary point on the line when no failure given
        float top_eps = top_sig / youngs;
        eps.assign({0.0f, top_eps});
        sig.assign({0.0f, top_sig});
        linear = true; E = youngs;
        sigma_yield = sigma_fail = fail_stress;
        eps_yield = eps_fail = (fail_stress == -1) ? -1 : top_eps;
        refresh_e_hat();
        return true;
    }

    bool model_bilinear(float youngs, float plastic, float yield_stress, float fail_stress = -1.0f)
    {
        if (youngs <= 0) return reject("Young's modulus must be positive");
        if (plastic <= 0 || plastic >= youngs) return reject("Plastic modulus must be positive but less than Young's modulus");
        if (yield_stress <= 0) return reject("Yield stress must be positive");
        if (fail_stress != -1.0f && fail_stress <= yield_stress) return reject("Failure stress must be positive and greater than the yield stress");
        float y_eps = yield_stress / youngs;
        float top_sig = fail_stress;
        if (top_sig == -1) top_sig = 3 * yield_stress;
        float slope = plastic;
        float icpt = yield_stress - slope * y_eps;
        float top_eps = (top_sig - icpt) / slope;
        eps.assign({0.0f, y_eps, top_eps});
        sig.assign({0.0f, yield_stress, top_sig});
        linear = false; E = youngs;
        sigma_yield = yield_stress; sigma_fail = fail_stress;
        eps_yield = y_eps;
        eps_fail = (fail_stress == -1.0f) ? -1.0f : top_eps;
        refresh_e_hat();
        return true;
    }

    bool model_data(int count, const float* e, const float* s)
    {
        if (count > 0 && e[0] == 0 && s[0] == 0) { e++; s++; count--; }
        if (count <= 0) return reject("Not enough data points");
        if (e[0] <= 0 || s[0] <= 0) return reject("First stress and strain data points negative or zero");
        std::vector<float> te(1, 0.0f), ts(1, 0.0f);
        float pe = 0.0f, ps = 0.0f;
        for (int i = 0; i < count; i++) {
            if (e[i] <= pe) return reject("Out of order strain data");
            if (s[i] <= ps) error = "Stress data is not monotonically increasing";   // the reference only records this one
            // the reference compares against te[0]/ts[0] = 0/0 (NaN): the test can never fire, keep it that way
            if (i > 0 && (s[i] - ps) / (e[i] - pe) > ts[0] / te[0])
                return reject("Slope of stress/strain curve should never exceed that of the first line segment (youngs modulus)");
            pe = e[i]; ps = s[i];
            te.push_back(pe); ts.push_back(ps);
        }
        eps.swap(te); sig.swap(ts);
        E = sig[1] / eps[1];
        sigma_fail = sig.back(); eps_fail = eps.back();
        linear = (count == 1);
        if (count <= 2) { sigma_yield = sig[1]; eps_yield = eps[1]; }
        else offset_yield(0.2f);
        refresh_e_hat();
        return true;
    }

    // 0.2 % offset yield point (VX_Material.cpp:437-470)
    bool offset_yield(float percent)
    {
        sigma_yield = eps_yield = -1.0f;
        float om = E, ob = (-percent / 100 * om);
        int last = (int)eps.size() - 1;
        for (int i = 1; i < last - 1; i++) {
            float x1 = eps[i], x2 = eps[i + 1], y1 = sig[i], y2 = sig[i + 1];
            float m = (y2 - y1) / (x2 - x1);
            float b = y1 - m * x1;
            if (om != m) {
                float xi = (b - ob) / (om - m);
                if (xi > x1 && xi < x2) {
                    float frac = (xi - x1) / (x2 - x1);
                    sigma_yield = y1 + frac * (y2 - y1);
                    eps_yield = xi;
                    return true;
                }
            }
        }
        sigma_yield = sigma_fail; eps_yield = eps_fail;
        return false;
    }

    // segment index used by stress()/modulus() for strains beyond the first segment
    int segment(float strain) const
    {
        int n = (int)eps.size();
        for (int i = 2; i < n; i++) if (strain <= eps[i] || i == n - 1) return i;
        return -1;
    }

    float modulus(float strain) const
    {
        if (failed(strain)) return 0.0f;
        if (strain <= eps[1] || linear) return E;
        int i = segment(strain);
        return i < 0 ? 0.0f : (sig[i] - sig[i - 1]) / (eps[i] - eps[i - 1]);
    }

    // host copy of the stress law the kernels evaluate (VX_Material.cpp:165-195); used by the facade's
    // CVX_Material::stress() accessor
    float stress(float strain, float transverse_sum = 0.0f, bool force_linear = false) const
    {
        if (failed(strain)) return 0.0f;
        if (strain <= eps[1] || linear || force_linear) {
            if (nu == 0.0f) return E * strain;
            return e_hat * ((1 - nu) * strain + nu * transverse_sum);
        }
        int i = segment(strain);
        if (i < 0) return 0.0f;
        float frac = (strain - eps[i - 1]) / (eps[i] - eps[i - 1]);
        float basic = sig[i - 1] + frac * (sig[i] - sig[i - 1]);
        if (nu == 0.0f) return basic;
        float seg_mod = (sig[i] - sig[i - 1]) / (eps[i] - eps[i - 1]);
        float seg_hat = seg_mod / ((1 - 2 * nu) * (1 + nu));
        float eff = basic / seg_mod;
        float eff_sum = transverse_sum * (eff / strain);
        return seg_hat * ((1 - nu) * eff + nu * eff_sum);
    }

    // inverse lookup (VX_Material.cpp:197-211)
    float strain_at(float stress) const
    {
        if (stress <= sig[1] || linear) return stress / E;
        int n = (int)eps.size();
        for (int i = 2; i < n; i++)
            if (stress <= sig[i] || i == n - 1) {
                float frac = (stress - sig[i - 1]) / (sig[i] - sig[i - 1]);
                return eps[i - 1] + frac * (eps[i] - eps[i - 1]);
            }
        return 0.0f;
    }

    bool reject(const char* why) { error = why; return false; }
};

// applies a vx_material_desc with the clamping of the reference's setters (VX_Material.cpp:472-523)
inline bool from_desc(Material& m, const vx_material_desc& d, const float* eps, const float* sig)
{
    bool ok = false;
    switch (d.model) {
    case VX_MODEL_LINEAR:   ok = m.model_linear(d.youngs_modulus, d.fail_stress); break;
    case VX_MODEL_BILINEAR: ok = m.model_bilinear(d.youngs_modulus, d.plastic_modulus, d.yield_stress, d.fail_stress); break;
    case VX_MODEL_DATA:     ok = m.model_data(d.n_points, eps, sig); break;
    default: m.error = "unknown material model";
    }
    if (!ok) return false;
    m.rho = d.density <= 0 ? FLT_MIN : d.density;
    float nu = d.poissons_ratio;
    if (nu < 0) nu = 0;
    if (nu >= 0.5) nu = 0.5 - FLT_EPSILON * 2;
    m.nu = nu;
    m.cte = d.cte;
    m.mu_s = d.mu_static <= 0 ? 0.0f : d.mu_static;
    m.mu_k = d.mu_kinetic <= 0 ? 0.0f : d.mu_kinetic;
    m.zeta_int = d.zeta_internal <= 0 ? 0.0f : d.zeta_internal;
    m.zeta_glob = d.zeta_global <= 0 ? 0.0f : d.zeta_global;
    m.zeta_coll = d.zeta_collision <= 0 ? 0.0f : d.zeta_collision;
    for (int a = 0; a < 3; a++) m.ext_scale[a] = d.ext_scale[a] <= 0 ? (double)FLT_MIN : d.ext_scale[a];
    m.refresh_e_hat();
    return true;
}

// mass properties of one voxel of nominal edge `nom` (CVX_MaterialVoxel::updateDerived)
struct MassProps {
    float mass = 0, mass_inv = 0, sqrt_mass = 0, first_moment = 0, inertia = 0, inertia_inv = 0;
    float two_sq_mes = 0, two_sq_ies3 = 0;
};
inline MassProps mass_props(const Material& m, double nom)
{
    MassProps p;
    double vol = nom * nom * nom;
    p.mass = (float)(vol * m.rho);
    p.inertia = (float)(p.mass * nom * nom / 6.0f);
    p.first_moment = (float)(p.mass * nom / 2.0f);
    if (vol == 0 || p.mass == 0 || p.inertia == 0) return MassProps{p.mass, 0, 0, p.first_moment, p.inertia, 0, 0, 0};
    p.mass_inv = 1.0f / p.mass;
    p.sqrt_mass = std::sqrt(p.mass);
    p.inertia_inv = 1.0f / p.inertia;
    p.two_sq_mes = (float)(2.0f * std::sqrt(p.mass * m.E * nom));
    p.two_sq_ies3 = (float)(2.0f * std::sqrt(p.inertia * m.E * nom * nom * nom));
    return p;
}

// combined material of the link between voxels of materials a and b (series springs)
inline Material combine(const Material& a, const Material& b)
{
    Material c;
    c.rho = 0.5f * (a.rho + b.rho);
    c.cte = 0.5f * (a.cte + b.cte);
    c.mu_s = 0.5f * (a.mu_s + b.mu_s);
    c.mu_k = 0.5f * (a.mu_k + b.mu_k);
    c.zeta_int = 0.5f * (a.zeta_int + b.zeta_int);
    c.zeta_glob = 0.5f * (a.zeta_glob + b.zeta_glob);
    c.zeta_coll = 0.5f * (a.zeta_coll + b.zeta_coll);

    float fail = -1.0f;                       // weaker of the two failure stresses, -1 = none
    if (a.sigma_fail == -1.0f) fail = b.sigma_fail;
    else if (b.sigma_fail == -1.0f) fail = a.sigma_fail;
    else fail = a.sigma_fail < b.sigma_fail ? a.sigma_fail : b.sigma_fail;

    if (a.linear && b.linear) c.model_linear(2.0f * a.E * b.E / (a.E + b.E), fail);
    else {
        std::vector<float> ce(1, 0.0f), cs(1, 0.0f);
        size_t ia = 1, ib = 1;
        while (ia < a.eps.size() && ib < b.eps.size()) {
            float x = a.eps[ia];
            if (b.eps[ib] < x) x = b.eps[ib];
            if (x == a.eps[ia]) ia++;
            if (ib < b.eps.size() && x == b.eps[ib]) ib++;
            float ma = a.modulus(x - FLT_EPSILON), mb = b.modulus(x - FLT_EPSILON);
            float series = 2.0f * ma * mb / (ma + mb);
            float prev_e = ce.back(), prev_s = cs.back();
            ce.push_back(x);
            cs.push_back(prev_s + series * (x - prev_e));
        }
        c.model_data((int)ce.size(), ce.data(), cs.data());
        c.sigma_fail = fail;
        c.eps_fail = (fail == -1.0f) ? -1.0f : c.strain_at(fail);
    }

    if (a.nu == 0 && b.nu == 0) c.nu = 0;
    else {                                    // nu such that eHat is the series combination of both eHats
        float eh = 2 * a.e_hat * b.e_hat / (a.e_hat + b.e_hat);
        float ee = c.E;
        float sq = (eh - ee) / (2 * eh) + 0.0625;
        c.nu = std::sqrt(sq) - 0.25;
    }
    c.refresh_e_hat();
    return c;
}

// beam constants of a link of length L = (float)nominal size (CVX_MaterialLink::updateDerived)
struct BeamConsts { float a1, a2, b1, b2, b3, sq_a1, sq_a2_ip, sq_b1, sq_b2_fmp, sq_b3_ip; };
inline BeamConsts beam_consts(const Material& m, double nom)
{
    BeamConsts k;
    float L = (float)nom, E = m.E;
    k.a1 = E * L;
    k.a2 = E * L * L * L / (12.0f * (1 + m.nu));
    k.b1 = E * L;
    k.b2 = E * L * L / 2.0f;
    k.b3 = E * L * L * L / 6.0f;
    k.sq_a1 = std::sqrt(k.a1);
    k.sq_a2_ip = std::sqrt(k.a2 * L * L / 6.0f);
    k.sq_b1 = std::sqrt(k.b1);
    k.sq_b2_fmp = std::sqrt(k.b2 * L / 2.0f);
    k.sq_b3_ip = std::sqrt(k.b3 * L * L / 6.0f);
    return k;
}

inline void fill_row(vx_voxmat_row& r, const Material& m, double nom)
{
    MassProps p = mass_props(m, nom);
    r.nom_size = nom;
    for (int a = 0; a < 3; a++) r.size[a] = nom * m.ext_scale[a];
    r.E = m.E; r.nu = m.nu; r.rho = m.rho; r.cte = m.cte; r.mu_static = m.mu_s; r.mu_kinetic = m.mu_k;
    r.zeta_internal = m.zeta_int; r.zeta_global = m.zeta_glob; r.zeta_collision = m.zeta_coll; r.e_hat = m.e_hat;
    r.mass = p.mass; r.mass_inv = p.mass_inv; r.sqrt_mass = p.sqrt_mass; r.first_moment = p.first_moment;
    r.moment_inertia = p.inertia; r.moment_inertia_inv = p.inertia_inv;
    r.two_sq_m_e_s = p.two_sq_mes; r.two_sq_i_e_s3 = p.two_sq_ies3;
    r.eps_yield = m.eps_yield; r.eps_fail = m.eps_fail; r.sigma_yield = m.sigma_yield; r.sigma_fail = m.sigma_fail;
    r.linear = m.linear ? 1 : 0; r.n_curve = (int)m.eps.size();
}

inline void fill_row(vx_linkmat_row& r, const Material& m, double nom, int a, int b)
{
    BeamConsts k = beam_consts(m, nom);
    r.mat_a = a < b ? a : b; r.mat_b = a < b ? b : a;
    r.linear = m.linear ? 1 : 0; r.n_curve = (int)m.eps.size();
    r.E = m.E; r.nu = m.nu; r.e_hat = m.e_hat;
    r.eps_yield = m.eps_yield; r.eps_fail = m.eps_fail; r.sigma_yield = m.sigma_yield; r.sigma_fail = m.sigma_fail;
    r.a1 = k.a1; r.a2 = k.a2; r.b1 = k.b1; r.b2 = k.b2; r.b3 = k.b3;
    r.sq_a1 = k.sq_a1; r.sq_a2_ip = k.sq_a2_ip; r.sq_b1 = k.sq_b1; r.sq_b2_fmp = k.sq_b2_fmp; r.sq_b3_ip = k.sq_b3_ip;
}

} // namespace vxm
