// vx_slabbed.hpp -- ONE lattice on several devices of ONE process (include/voxelyze_b200.h, "vx_slabbed_*").
//
// Host-side composition over the per-device C-ABI of the same header and nothing else: the whole model a caller of
// CVoxelyze hands over (src/Voxelyze.cpp:422-461) is cut into z-slabs, every slab is an ordinary vx_sim handle on its own
// device that stores its planes plus one ghost plane per cut (VX_VF_GHOST), and the calls of the hot path
// (CVoxelyze::doTimeStep, src/Voxelyze.cpp:251-284) and of the state accessors are fanned out / gathered in the caller's
// voxel and link numbering.  Cut-crossing links are evaluated on both sides from identical inputs, so a slabbed run has
// the bits of the unsplit run.  Halo transport, best first:
//   2  peer stores: the step kernel of a slab writes its boundary poses into the neighbours' ghost planes
//      (vx_peer_export / vx_peer_attach, same-process mappings with cudaDeviceEnablePeerAccess between devices);
//      all slabs are queued (vx_slab_step_begin) before any is waited for (vx_slab_step_finish), the host does nothing
//      per step.  Slabs that share a device (tests on one GPU) are stepped in lock step instead, so that a queued wait
//      can never sit in front of the work it waits for.
//   1  host copies of the two pose fields through vx_download / vx_upload after every step (any implementation of
//      the header; what the CPU checkers run, and the fall-back when peer mappings cannot be made).
// Written against the public header only (C++11), so the same file is compiled into every library that implements it.
#ifndef VX_SLABBED_HPP
#define VX_SLABBED_HPP

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct vx_slabbed {
    struct Part {
        int z0 = 0, z1 = 0, lo = 0, hi = 0;             // owned planes [z0, z1), stored planes [lo, hi) (lattice z minus the model's lowest)
        std::vector<int32_t> l2g;                         // local voxel -> caller index of the whole model
        std::vector<int> plane_first;                     // local index of the first voxel of stored plane lo + k (hi - lo + 1 entries)
        std::vector<int32_t> link_l2g;                    // local link -> link index of the whole model
        int owned_first() const { return plane_first[z0 - lo]; }
        int owned_count() const { return plane_first[z1 - lo] - plane_first[z0 - lo]; }
    };
    double voxel_size = 0;
    std::vector<vx_sim*> slab;                            // one handle per listed device
    std::vector<int> device;
    bool shared_device = false;
    int active = 0, halo = 0;
    std::string err;
    bool poisson = false;
    int N = 0, L = 0, z_origin = 0;
    std::vector<uint16_t> mat;                            // material of every voxel (link quantities of vx_slabbed_state_info)
    std::vector<Part> part;
    std::vector<int32_t> owner, local;                    // per voxel of the model: owning slab, index there
    std::vector<int32_t> lneg, lpos; std::vector<uint8_t> laxis;
    std::vector<int32_t> link_owner, link_local, link_slab2, link_local2;     // owner copy; the copy above a cut (-1: none)
};

namespace vxs {

static int fail(vx_slabbed* m, int code, const std::string& what) { m->err = what; return code; }
static int fail_from(vx_slabbed* m, int k, int code, const char* call)
{
    m->err = std::string(call) + " on slab " + std::to_string(k) + ": " + vx_last_error(m->slab[k]);
    return code;
}

static size_t field_bytes(int field)
{
    switch (field) {
        case VX_F_POS: case VX_F_LINMOM: case VX_F_ANGMOM: return 24;
        case VX_F_ORIENT: return 32;
        case VX_F_TEMP: case VX_F_VOXFLAGS: return 4;
        case VX_F_PSTRAIN: return 12;
        case VX_F_FORCE_NEG: case VX_F_FORCE_POS: case VX_F_MOMENT_NEG: case VX_F_MOMENT_POS: case VX_F_POS2: case VX_F_ANGLE1V: case VX_F_ANGLE2V: return 24;
        case VX_F_STRAIN: case VX_F_MAXSTRAIN: case VX_F_STRAINOFFSET: case VX_F_STRESS: case VX_F_LINKFLAGS: return 4;
        default: return 0;
    }
}
static bool is_link_field(int field) { return field >= 16; }

// planes are dealt as evenly as possible, like voxelyze_b200/slab.py::slab_range
static void slab_range(int nz, int k, int g, int& z0, int& z1)
{
    const int base = nz / g, rem = nz % g;
    z0 = k * base + std::min(k, rem); z1 = z0 + base + (k < rem ? 1 : 0);
}

// host copies of the boundary poses into the ghost planes across every cut
static int host_exchange(vx_slabbed* m)
{
    std::vector<double> buf;
    for (int k = 0; k + 1 < m->active; k++) {
        const vx_slabbed::Part &a = m->part[k], &b = m->part[k + 1];
        const int fields[3] = {VX_F_POS, VX_F_ORIENT, VX_F_PSTRAIN};
        const int n_fields = m->poisson ? 3 : 2;
        for (int dir = 0; dir < 2; dir++) {
            const int z = dir == 0 ? a.z1 - 1 : b.z0;                          // plane that travels: up out of a, down out of b
            const vx_slabbed::Part &src = dir == 0 ? a : b, &dst = dir == 0 ? b : a;
            const int ks = dir == 0 ? k : k + 1, kd = dir == 0 ? k + 1 : k;
            const int sf = src.plane_first[z - src.lo], sn = src.plane_first[z - src.lo + 1] - sf;
            const int df = dst.plane_first[z - dst.lo], dn = dst.plane_first[z - dst.lo + 1] - df;
            if (sn != dn) return fail(m, VX_ERR_TOPOLOGY, "slab planes of different size");
            if (sn == 0) continue;                                             // the body has no voxels in this plane
            for (int f = 0; f < n_fields; f++) {
                buf.resize((size_t)sn * 4);
                int rc = vx_download(m->slab[ks], fields[f], sf, sn, buf.data()); if (rc != VX_OK) return fail_from(m, ks, rc, "vx_download");
                rc = vx_upload(m->slab[kd], fields[f], df, dn, buf.data()); if (rc != VX_OK) return fail_from(m, kd, rc, "vx_upload");
            }
        }
    }
    return VX_OK;
}

static int exchange_all(vx_slabbed* m)
{
    if (m->active < 2) return VX_OK;
    if (m->halo == 2) {
        for (int k = 0; k < m->active; k++) { int rc = vx_slab_exchange(m->slab[k]); if (rc != VX_OK) return fail_from(m, k, rc, "vx_slab_exchange"); }
        return VX_OK;
    }
    return host_exchange(m);
}

// where voxel g of the model is stored: its owner first, then the ghost copies in the slabs across the cuts it borders
static int copies(const vx_slabbed* m, int g, int where[3][2])
{
    const int k = m->owner[g], j = m->local[g];
    const vx_slabbed::Part& p = m->part[k];
    const int pl = (int)(std::upper_bound(p.plane_first.begin(), p.plane_first.end(), j) - p.plane_first.begin()) - 1;     // stored plane of j
    const int z = p.lo + pl, off = j - p.plane_first[pl];
    int n = 0;
    where[n][0] = k; where[n][1] = j; n++;
    if (z == p.z1 - 1 && k + 1 < m->active) { const vx_slabbed::Part& q = m->part[k + 1]; where[n][0] = k + 1; where[n][1] = q.plane_first[z - q.lo] + off; n++; }
    if (z == p.z0 && k > 0) { const vx_slabbed::Part& q = m->part[k - 1]; where[n][0] = k - 1; where[n][1] = q.plane_first[z - q.lo] + off; n++; }
    return n;
}

// VX_SLABBED_THREADS=0: the calling thread drives every device itself
static bool use_threads()
{
    static int on = -1;
    if (on < 0) { const char* e = getenv("VX_SLABBED_THREADS"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}

static void detach_all(vx_slabbed* m) { for (size_t k = 0; k < m->slab.size(); k++) vx_peer_detach(m->slab[k]); }

// every slab's boundary plane mirrored into its neighbour's ghost plane by the step kernels themselves, if the library can
static bool connect_peers(vx_slabbed* m)
{
    for (int k = 0; k < m->active; k++) if (vx_active_path(m->slab[k]) != 2) return false;
    for (int k = 0; k + 1 < m->active; k++) {
        const vx_slabbed::Part &a = m->part[k], &b = m->part[k + 1];
        vx_peer_desc d;
        if (vx_peer_export(m->slab[k], m->z_origin + a.z1, 1, &d) != VX_OK || vx_peer_attach(m->slab[k + 1], m->z_origin + b.z0, &d) != VX_OK ||
            vx_peer_export(m->slab[k + 1], m->z_origin + b.z0 - 1, 0, &d) != VX_OK || vx_peer_attach(m->slab[k], m->z_origin + a.z1 - 1, &d) != VX_OK) {
            detach_all(m);
            return false;
        }
    }
    return true;
}

} // namespace vxs

extern "C" {

int vx_slabbed_create(double voxel_size, int n_slabs, const int* devices, vx_slabbed** out)
{
    if (!out || n_slabs < 1 || n_slabs > 64) return VX_ERR_ARG;
    *out = nullptr;
    vx_slabbed* m = new vx_slabbed();
    m->voxel_size = voxel_size;
    for (int k = 0; k < n_slabs; k++) {
        const int dev = devices ? devices[k] : k;
        vx_sim* s = nullptr;
        const int rc = vx_create(voxel_size, dev, &s);
        if (rc != VX_OK) { for (size_t j = 0; j < m->slab.size(); j++) vx_destroy(m->slab[j]); delete m; return rc; }
        for (size_t j = 0; j < m->device.size(); j++) if (m->device[j] == dev) m->shared_device = true;
        m->slab.push_back(s); m->device.push_back(dev);
    }
    *out = m;
    return VX_OK;
}

void vx_slabbed_destroy(vx_slabbed* m)
{
    if (!m) return;
    if (m->halo == 2) vxs::detach_all(m);
    for (size_t k = 0; k < m->slab.size(); k++) vx_destroy(m->slab[k]);
    delete m;
}

const char* vx_slabbed_last_error(const vx_slabbed* m) { return m ? m->err.c_str() : "null handle"; }
int vx_slabbed_slab_count(const vx_slabbed* m) { return m ? m->active : 0; }
vx_sim* vx_slabbed_slab(vx_slabbed* m, int k) { return m && k >= 0 && k < (int)m->slab.size() ? m->slab[k] : nullptr; }
int vx_slabbed_halo_mode(const vx_slabbed* m) { return m ? (m->active > 1 ? m->halo : 0) : 0; }
int vx_slabbed_voxel_count(const vx_slabbed* m) { return m ? m->N : 0; }
int vx_slabbed_link_count(const vx_slabbed* m) { return m ? m->L : 0; }

int vx_slabbed_set_materials(vx_slabbed* m, int n, const vx_material_desc* descs)
{
    if (!m || n < 0 || (n && !descs)) return VX_ERR_ARG;
    bool poisson = false;
    for (int i = 0; i < n; i++) poisson = poisson || descs[i].poissons_ratio != 0.0f;
    for (size_t k = 0; k < m->slab.size(); k++) { int rc = vx_set_materials(m->slab[k], n, descs); if (rc != VX_OK) return vxs::fail_from(m, (int)k, rc, "vx_set_materials"); }
    const bool changed = poisson != m->poisson;
    m->poisson = poisson;
    // a voxel's Poisson strain needs all of its links (src/VX_Voxel.cpp:300-374); a ghost copy does not have them and takes
    // its owner's value through the halo -- also right after Poisson's ratio was switched on mid-run
    if (m->active < 2) return VX_OK;
    if (changed && m->halo == 2) { vxs::detach_all(m); m->halo = vxs::connect_peers(m) ? 2 : 1; }      // the peer mappings cover the Poisson strain arrays only when they exist
    return poisson ? vxs::exchange_all(m) : VX_OK;           // every slab has just rebuilt its Poisson strains from its own links
}

int vx_slabbed_set_gravity(vx_slabbed* m, float g)
{
    if (!m) return VX_ERR_ARG;
    for (size_t k = 0; k < m->slab.size(); k++) { int rc = vx_set_gravity(m->slab[k], g); if (rc != VX_OK) return vxs::fail_from(m, (int)k, rc, "vx_set_gravity"); }
    return VX_OK;
}

int vx_slabbed_enable_floor(vx_slabbed* m, int enabled)
{
    if (!m) return VX_ERR_ARG;
    for (size_t k = 0; k < m->slab.size(); k++) { int rc = vx_enable_floor(m->slab[k], enabled); if (rc != VX_OK) return vxs::fail_from(m, (int)k, rc, "vx_enable_floor"); }
    return VX_OK;
}

int vx_slabbed_set_voxels(vx_slabbed* m, int n, const int32_t* ijk, const uint16_t* mat)
{
    if (!m || n < 0 || (n && (!ijk || !mat))) return VX_ERR_ARG;
    if (m->halo == 2) vxs::detach_all(m);
    m->halo = 0; m->active = 0; m->N = 0; m->L = 0; m->part.clear();
    m->owner.clear(); m->local.clear(); m->lneg.clear(); m->lpos.clear(); m->laxis.clear();
    m->link_owner.clear(); m->link_local.clear(); m->link_slab2.clear(); m->link_local2.clear();
    const int G = (int)m->slab.size();
    if (n == 0) {
        for (int k = 0; k < G; k++) { int rc = vx_set_voxels(m->slab[k], 0, nullptr, nullptr, nullptr, nullptr); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_set_voxels"); }
        return VX_OK;
    }
    int lo[3] = {32767, 32767, 32767}, hi[3] = {-32768, -32768, -32768};
    for (int i = 0; i < n; i++) for (int a = 0; a < 3; a++) {
        const int c = ijk[3 * i + a];
        if (c < -32768 || c > 32767) return vxs::fail(m, VX_ERR_ARG, "lattice index does not fit a short");
        lo[a] = std::min(lo[a], c); hi[a] = std::max(hi[a], c);
    }
    const long long ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1;
    if (ex * ey * ez > std::max(16LL * n, 1LL << 22)) return vxs::fail(m, VX_ERR_UNSUPPORTED, "body too sparse for its bounding box to be cut into slabs");
    const int act = (int)std::max(1LL, std::min((long long)G, ez / 2));       // at least two owned planes per slab
    // occupancy grid over the bounding box: cell -> voxel
    std::vector<int32_t> grid((size_t)(ex * ey * ez), -1);
    auto cell = [&](int x, int y, int z) -> size_t { return ((size_t)(z - lo[2]) * ey + (y - lo[1])) * ex + (x - lo[0]); };
    for (int i = 0; i < n; i++) {
        int32_t& c = grid[cell(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2])];
        if (c >= 0) return vxs::fail(m, VX_ERR_TOPOLOGY, "duplicate voxel");
        c = i;
    }
    // links of the whole model in the reference's creation order (src/Voxelyze.cpp:453-455, 508-539): for every voxel in
    // order, for X+ X- Y+ Y- Z+ Z-, a link to a neighbour that already exists
    std::vector<int32_t> link_of((size_t)3 * n, -1);                         // (negative-end voxel, axis) -> link
    for (int i = 0; i < n; i++) {
        const int x = ijk[3 * i], y = ijk[3 * i + 1], z = ijk[3 * i + 2];
        for (int d = 0; d < 6; d++) {
            const int axis = d >> 1, sgn = (d & 1) ? -1 : 1;
            int q[3] = {x, y, z}; q[axis] += sgn;
            if (q[axis] < lo[axis] || q[axis] > hi[axis]) continue;
            const int32_t o = grid[cell(q[0], q[1], q[2])];
            if (o < 0 || o > i) continue;
            const int32_t neg = sgn > 0 ? i : o, pos = sgn > 0 ? o : i;
            link_of[(size_t)3 * neg + axis] = (int32_t)m->lneg.size();
            m->lneg.push_back(neg); m->lpos.push_back(pos); m->laxis.push_back((uint8_t)axis);
        }
    }
    m->N = n; m->L = (int)m->lneg.size(); m->z_origin = lo[2]; m->active = act;
    m->mat.assign(mat, mat + n);
    // voxels plane by plane, x fastest: every stored plane of a slab is one contiguous index range on both sides of a cut
    std::vector<int32_t> order; order.reserve(n);
    std::vector<int> plane_start((size_t)ez + 1, 0);
    for (long long z = 0; z < ez; z++) {
        plane_start[(size_t)z] = (int)order.size();
        for (long long c = z * ex * ey; c < (z + 1) * ex * ey; c++) if (grid[(size_t)c] >= 0) order.push_back(grid[(size_t)c]);
    }
    plane_start[(size_t)ez] = (int)order.size();
    m->part.assign(act, vx_slabbed::Part());
    m->owner.assign(n, -1); m->local.assign(n, -1);
    m->link_owner.assign(m->L, -1); m->link_local.assign(m->L, -1); m->link_slab2.assign(m->L, -1); m->link_local2.assign(m->L, -1);
    // pass 1 (host only): which planes each slab stores, who owns which voxel
    for (int k = 0; k < act; k++) {
        vx_slabbed::Part& p = m->part[k];
        vxs::slab_range((int)ez, k, act, p.z0, p.z1);
        p.lo = k > 0 ? p.z0 - 1 : p.z0; p.hi = k < act - 1 ? p.z1 + 1 : p.z1;
        p.l2g.assign(order.begin() + plane_start[p.lo], order.begin() + plane_start[p.hi]);
        p.plane_first.resize(p.hi - p.lo + 1);
        for (int z = p.lo; z <= p.hi; z++) p.plane_first[z - p.lo] = plane_start[z] - plane_start[p.lo];
        if (p.owned_count() == 0) return vxs::fail(m, VX_ERR_TOPOLOGY, "a slab without voxels (the body has an empty z range)");
        for (int j = p.owned_first(); j < p.owned_first() + p.owned_count(); j++) { m->owner[p.l2g[j]] = k; m->local[p.l2g[j]] = j; }
    }
    // pass 2: every slab's device model and its links in the numbering of the whole model -- the slabs are independent
    // handles on different devices, so each is built by a host thread of its own (big models: minutes otherwise)
    std::vector<int> rcs(G, VX_OK); std::vector<const char*> where(G, "");
    auto build_slab = [&](int k) {
        if (k >= act) { rcs[k] = vx_set_voxels(m->slab[k], 0, nullptr, nullptr, nullptr, nullptr); where[k] = "vx_set_voxels"; return; }
        vx_slabbed::Part& p = m->part[k];
        const int cnt = (int)p.l2g.size();
        std::vector<int32_t> lijk((size_t)3 * cnt); std::vector<uint16_t> lmat(cnt); std::vector<uint32_t> lflags(cnt, 0u);
        for (int j = 0; j < cnt; j++) {
            const int32_t g = p.l2g[j];
            lijk[3 * j] = ijk[3 * g]; lijk[3 * j + 1] = ijk[3 * g + 1]; lijk[3 * j + 2] = ijk[3 * g + 2]; lmat[j] = mat[g];
            if (m->owner[g] != k) lflags[j] = VX_VF_GHOST;
        }
        rcs[k] = vx_set_voxels(m->slab[k], cnt, lijk.data(), lmat.data(), nullptr, act > 1 ? lflags.data() : nullptr); where[k] = "vx_set_voxels";
        if (rcs[k] != VX_OK) return;
        const int lk = vx_link_count(m->slab[k]);
        std::vector<int32_t> vn(lk), vp(lk); std::vector<uint8_t> ax(lk);
        if (lk) { rcs[k] = vx_get_links(m->slab[k], vn.data(), vp.data(), ax.data()); where[k] = "vx_get_links"; if (rcs[k] != VX_OK) return; }
        p.link_l2g.assign(lk, -1);
        for (int j = 0; j < lk; j++) {
            const int32_t gneg = p.l2g[vn[j]], gl = link_of[(size_t)3 * gneg + ax[j]];
            if (gl < 0 || m->lpos[gl] != p.l2g[vp[j]]) { rcs[k] = VX_ERR_TOPOLOGY; where[k] = nullptr; return; }
            p.link_l2g[j] = gl;
            if (m->owner[gneg] == k) { m->link_owner[gl] = k; m->link_local[gl] = j; }
            else if (m->owner[m->lpos[gl]] == k) { m->link_slab2[gl] = k; m->link_local2[gl] = j; }       // crosses the cut below this slab
        }
    };
    if (act > 1 && vxs::use_threads()) {
        std::vector<std::thread> workers;
        for (int k = 0; k < G; k++) workers.emplace_back(build_slab, k);
        for (size_t k = 0; k < workers.size(); k++) workers[k].join();
    } else {
        for (int k = 0; k < G; k++) build_slab(k);
    }
    for (int k = 0; k < G; k++) {
        if (rcs[k] == VX_OK) continue;
        return where[k] ? vxs::fail_from(m, k, rcs[k], where[k]) : vxs::fail(m, VX_ERR_TOPOLOGY, "slab link without a counterpart in the whole model");
    }
    for (int g = 0; g < m->L; g++) if (m->link_owner[g] < 0) return vxs::fail(m, VX_ERR_TOPOLOGY, "a link of the model is in no slab");
    if (act > 1) {
        m->halo = vxs::connect_peers(m) ? 2 : 1;
        return vxs::exchange_all(m);
    }
    return VX_OK;
}

int vx_slabbed_get_links(const vx_slabbed* m, int32_t* v_neg, int32_t* v_pos, uint8_t* axis)
{
    if (!m) return VX_ERR_ARG;
    if (v_neg && m->L) memcpy(v_neg, m->lneg.data(), sizeof(int32_t) * m->L);
    if (v_pos && m->L) memcpy(v_pos, m->lpos.data(), sizeof(int32_t) * m->L);
    if (axis && m->L) memcpy(axis, m->laxis.data(), m->L);
    return VX_OK;
}

int vx_slabbed_set_externals(vx_slabbed* m, int n, const int32_t* voxel, const uint8_t* dof, const float* force, const float* moment,
                             const double* translation, const double* rotation)
{
    if (!m || n < 0 || (n && (!voxel || !dof))) return VX_ERR_ARG;
    for (int i = 0; i < n; i++) if (voxel[i] < 0 || voxel[i] >= m->N) return vxs::fail(m, VX_ERR_ARG, "external on a voxel that does not exist");
    for (int k = 0; k < m->active; k++) {                 // a ghost's pose comes from its owner: externals go to the owner only
        std::vector<int32_t> v; std::vector<uint8_t> d; std::vector<float> f, mo; std::vector<double> t, r;
        for (int i = 0; i < n; i++) {
            if (m->owner[voxel[i]] != k) continue;
            v.push_back(m->local[voxel[i]]); d.push_back(dof[i]);
            for (int c = 0; c < 3; c++) {
                f.push_back(force ? force[3 * i + c] : 0.f); mo.push_back(moment ? moment[3 * i + c] : 0.f);
                t.push_back(translation ? translation[3 * i + c] : 0.0); r.push_back(rotation ? rotation[3 * i + c] : 0.0);
            }
        }
        int rc = vx_set_externals(m->slab[k], (int)v.size(), v.data(), d.data(), f.data(), mo.data(), t.data(), r.data());
        if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_set_externals");
    }
    return m->poisson ? vxs::exchange_all(m) : VX_OK;         // fixed degrees of freedom enter a voxel's Poisson strain
}

int vx_slabbed_set_temperature_all(vx_slabbed* m, float t)
{
    if (!m) return VX_ERR_ARG;
    for (int k = 0; k < m->active; k++) { int rc = vx_set_temperature_all(m->slab[k], t); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_set_temperature_all"); }
    return VX_OK;
}

int vx_slabbed_set_temperature(vx_slabbed* m, int n, const float* t)
{
    if (!m || n != m->N || (n && !t)) return VX_ERR_ARG;
    std::vector<float> tl;
    for (int k = 0; k < m->active; k++) {                 // ghost copies take the temperature of the voxel they mirror
        const vx_slabbed::Part& p = m->part[k];
        tl.resize(p.l2g.size());
        for (size_t j = 0; j < p.l2g.size(); j++) tl[j] = t[p.l2g[j]];
        int rc = vx_set_temperature(m->slab[k], (int)tl.size(), tl.data()); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_set_temperature");
    }
    return VX_OK;
}

int vx_slabbed_recommended_dt(vx_slabbed* m, float* dt)
{
    if (!m || !dt) return VX_ERR_ARG;
    *dt = 0.f;
    for (int k = 0; k < m->active; k++) {                 // the stiffest link / lightest voxel of any slab decides (src/Voxelyze.cpp:286-311)
        float d = 0.f;
        int rc = vx_recommended_dt(m->slab[k], &d); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_recommended_dt");
        if (k == 0 || d < *dt) *dt = d;
    }
    return VX_OK;
}

int vx_slabbed_step(vx_slabbed* m, float dt, int n_steps, int* diverged_step)
{
    if (!m || n_steps < 0) return VX_ERR_ARG;
    if (m->active == 0 || n_steps == 0 || dt == 0) return VX_OK;
    if (m->active == 1) {
        int rc = vx_step(m->slab[0], dt, n_steps, diverged_step);
        return rc == VX_OK || rc == VX_DIVERGED ? rc : vxs::fail_from(m, 0, rc, "vx_step");
    }
    if (dt < 0 && m->poisson && n_steps > 1) {            // with Poisson coupling the stable step follows the state (src/VX_Link.cpp:259-267): step by step
        for (int s = 0; s < n_steps; s++) {
            int d = -1;
            int rc = vx_slabbed_step(m, -1.0f, 1, &d);
            if (rc == VX_DIVERGED) { if (diverged_step) *diverged_step = s; return rc; }
            if (rc != VX_OK) return rc;
        }
        return VX_OK;
    }
    if (dt < 0) { int rc = vx_slabbed_recommended_dt(m, &dt); if (rc != VX_OK) return rc; if (dt <= 0) return VX_OK; }      // state independent without Poisson coupling
    int first_div = -1, err = VX_OK;
    if (m->halo == 2 && !m->shared_device && n_steps >= 4 && vxs::use_threads()) {
        // one host thread per device for the length of the call: queueing a step costs the host a handful of launches per
        // slab, which one thread feeding eight devices cannot hide behind a 2.7 ms step (2.87 against 2.68 ms measured)
        std::vector<int> rc_begin(m->active, VX_OK), rc_end(m->active, VX_OK), div(m->active, -1);
        std::vector<std::thread> workers;
        for (int k = 0; k < m->active; k++)
            workers.emplace_back([m, k, dt, n_steps, &rc_begin, &rc_end, &div]() {
                rc_begin[k] = vx_slab_step_begin(m->slab[k], dt, n_steps);
                if (rc_begin[k] == VX_OK) rc_end[k] = vx_slab_step_finish(m->slab[k], &div[k]);
            });
        for (size_t k = 0; k < workers.size(); k++) workers[k].join();
        for (int k = 0; k < m->active; k++) {
            if (rc_begin[k] != VX_OK) { if (err == VX_OK) err = vxs::fail_from(m, k, rc_begin[k], "vx_slab_step_begin"); }
            else if (rc_end[k] == VX_DIVERGED) { if (first_div < 0 || div[k] < first_div) first_div = div[k]; }
            else if (rc_end[k] != VX_OK && err == VX_OK) err = vxs::fail_from(m, k, rc_end[k], "vx_slab_step_finish");
        }
    } else if (m->halo == 2 && !m->shared_device) {
        // every device gets all n steps queued before the host waits for any of them
        int queued = 0;
        for (; queued < m->active; queued++) {
            int rc = vx_slab_step_begin(m->slab[queued], dt, n_steps);
            if (rc != VX_OK) { err = vxs::fail_from(m, queued, rc, "vx_slab_step_begin"); break; }
        }
        for (int k = 0; k < queued; k++) {
            int d = -1;
            int rc = vx_slab_step_finish(m->slab[k], &d);
            if (rc == VX_DIVERGED) { if (first_div < 0 || d < first_div) first_div = d; }
            else if (rc != VX_OK && err == VX_OK) err = vxs::fail_from(m, k, rc, "vx_slab_step_finish");
        }
    } else {
        for (int s = 0; s < n_steps && first_div < 0 && err == VX_OK; s++) {
            for (int k = 0; k < m->active; k++) {
                int d = -1;
                int rc = m->halo == 2 ? vx_slab_step(m->slab[k], dt, 1, &d) : vx_step(m->slab[k], dt, 1, &d);
                if (rc == VX_DIVERGED) first_div = s;
                else if (rc != VX_OK && err == VX_OK) { err = vxs::fail_from(m, k, rc, m->halo == 2 ? "vx_slab_step" : "vx_step"); break; }
            }
            if (err == VX_OK && first_div < 0 && m->halo == 1) err = vxs::host_exchange(m);
        }
    }
    if (err != VX_OK) return err;
    // doTimeStep returns false as soon as any link anywhere diverged (src/Voxelyze.cpp:265-269); the slabs that did not
    // diverge themselves may have advanced past that step -- a diverged simulation is not continued
    if (first_div >= 0) { if (diverged_step) *diverged_step = first_div; return VX_DIVERGED; }
    return VX_OK;
}

int vx_slabbed_reset(vx_slabbed* m)
{
    if (!m) return VX_ERR_ARG;
    for (int k = 0; k < m->active; k++) { int rc = vx_reset(m->slab[k]); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_reset"); }
    return vxs::exchange_all(m);
}

float vx_slabbed_time(const vx_slabbed* m) { return m && m->active ? vx_time(m->slab[0]) : 0.f; }

int vx_slabbed_set_clock(vx_slabbed* m, float time, float previous_dt)
{
    if (!m) return VX_ERR_ARG;
    for (int k = 0; k < m->active; k++) { int rc = vx_set_clock(m->slab[k], time, previous_dt); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_set_clock"); }
    return VX_OK;
}

int vx_slabbed_download(vx_slabbed* m, int field, int first, int count, void* dst)
{
    const size_t eb = vxs::field_bytes(field);
    if (!m || !eb || first < 0 || count < 0 || (count && !dst)) return VX_ERR_ARG;
    const bool link = vxs::is_link_field(field);
    if ((long long)first + count > (link ? m->L : m->N)) return vxs::fail(m, VX_ERR_ARG, "vx_slabbed_download: range");
    unsigned char* out = (unsigned char*)dst;
    if (count <= 4) {
        for (int i = first; i < first + count; i++) {
            const int k = link ? m->link_owner[i] : m->owner[i], j = link ? m->link_local[i] : m->local[i];
            int rc = vx_download(m->slab[k], field, j, 1, out + (size_t)(i - first) * eb); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_download");
        }
        return VX_OK;
    }
    std::vector<unsigned char> buf;
    for (int k = 0; k < m->active; k++) {
        const vx_slabbed::Part& p = m->part[k];
        const int lf = link ? 0 : p.owned_first(), ln = link ? (int)p.link_l2g.size() : p.owned_count();
        if (!ln) continue;
        buf.resize((size_t)ln * eb);
        int rc = vx_download(m->slab[k], field, lf, ln, buf.data()); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_download");
        for (int j = 0; j < ln; j++) {
            const int g = link ? p.link_l2g[j] : p.l2g[lf + j];
            if (g < first || g >= first + count || (link && (m->link_owner[g] != k || m->link_local[g] != j))) continue;
            memcpy(out + (size_t)(g - first) * eb, buf.data() + (size_t)j * eb, eb);
        }
    }
    return VX_OK;
}

int vx_slabbed_upload(vx_slabbed* m, int field, int first, int count, const void* src)
{
    const size_t eb = vxs::field_bytes(field);
    if (!m || !eb || first < 0 || count < 0 || (count && !src)) return VX_ERR_ARG;
    if (vxs::is_link_field(field)) return vxs::fail(m, VX_ERR_UNSUPPORTED, "link fields are written through vx_slabbed_upload_link_state");
    if ((long long)first + count > m->N) return vxs::fail(m, VX_ERR_ARG, "vx_slabbed_upload: range");
    const unsigned char* in = (const unsigned char*)src;
    // every stored copy of a voxel takes the value: its owner's and the ghost copies across the cuts
    if (count <= 16) {
        for (int i = first; i < first + count; i++) {
            int where[3][2];
            const int nc = vxs::copies(m, i, where);
            for (int c = 0; c < nc; c++) {
                unsigned char v[32]; memcpy(v, in + (size_t)(i - first) * eb, eb);
                if (field == VX_F_VOXFLAGS && c > 0) { uint32_t w; memcpy(&w, v, 4); w |= VX_VF_GHOST; memcpy(v, &w, 4); }
                int rc = vx_upload(m->slab[where[c][0]], field, where[c][1], 1, v); if (rc != VX_OK) return vxs::fail_from(m, where[c][0], rc, "vx_upload");
            }
        }
        return VX_OK;
    }
    std::vector<unsigned char> buf;
    for (int k = 0; k < m->active; k++) {
        const vx_slabbed::Part& p = m->part[k];
        int j = 0;
        const int cnt = (int)p.l2g.size();
        while (j < cnt) {
            while (j < cnt && (p.l2g[j] < first || p.l2g[j] >= first + count)) j++;
            int e = j;
            while (e < cnt && p.l2g[e] >= first && p.l2g[e] < first + count) e++;
            if (e == j) break;
            buf.resize((size_t)(e - j) * eb);
            for (int q = j; q < e; q++) {
                memcpy(buf.data() + (size_t)(q - j) * eb, in + (size_t)(p.l2g[q] - first) * eb, eb);
                if (field == VX_F_VOXFLAGS && m->owner[p.l2g[q]] != k) { uint32_t w; memcpy(&w, buf.data() + (size_t)(q - j) * eb, 4); w |= VX_VF_GHOST; memcpy(buf.data() + (size_t)(q - j) * eb, &w, 4); }
            }
            int rc = vx_upload(m->slab[k], field, j, e - j, buf.data()); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_upload");
            j = e;
        }
    }
    return VX_OK;
}

int vx_slabbed_download_voxel_state(vx_slabbed* m, int first, int count, vx_voxel_state* dst)
{
    if (!m || first < 0 || count < 0 || (count && !dst) || (long long)first + count > m->N) return VX_ERR_ARG;
    if (count <= 4) {
        for (int i = first; i < first + count; i++) {
            int rc = vx_download_voxel_state(m->slab[m->owner[i]], m->local[i], 1, dst + (i - first)); if (rc != VX_OK) return vxs::fail_from(m, m->owner[i], rc, "vx_download_voxel_state");
        }
        return VX_OK;
    }
    std::vector<vx_voxel_state> buf;
    for (int k = 0; k < m->active; k++) {
        const vx_slabbed::Part& p = m->part[k];
        const int lf = p.owned_first(), ln = p.owned_count();
        buf.resize(ln);
        int rc = vx_download_voxel_state(m->slab[k], lf, ln, buf.data()); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_download_voxel_state");
        for (int j = 0; j < ln; j++) { const int g = p.l2g[lf + j]; if (g >= first && g < first + count) dst[g - first] = buf[j]; }
    }
    return VX_OK;
}

int vx_slabbed_download_link_state(vx_slabbed* m, int first, int count, vx_link_state* dst)
{
    if (!m || first < 0 || count < 0 || (count && !dst) || (long long)first + count > m->L) return VX_ERR_ARG;
    if (count <= 4) {
        for (int i = first; i < first + count; i++) {
            int rc = vx_download_link_state(m->slab[m->link_owner[i]], m->link_local[i], 1, dst + (i - first)); if (rc != VX_OK) return vxs::fail_from(m, m->link_owner[i], rc, "vx_download_link_state");
        }
        return VX_OK;
    }
    std::vector<vx_link_state> buf;
    for (int k = 0; k < m->active; k++) {
        const vx_slabbed::Part& p = m->part[k];
        const int ln = (int)p.link_l2g.size();
        if (!ln) continue;
        buf.resize(ln);
        int rc = vx_download_link_state(m->slab[k], 0, ln, buf.data()); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_download_link_state");
        for (int j = 0; j < ln; j++) { const int g = p.link_l2g[j]; if (g >= first && g < first + count && m->link_owner[g] == k && m->link_local[g] == j) dst[g - first] = buf[j]; }
    }
    return VX_OK;
}

int vx_slabbed_upload_link_state(vx_slabbed* m, int first, int count, const vx_link_state* src)
{
    if (!m || first < 0 || count < 0 || (count && !src) || (long long)first + count > m->L) return VX_ERR_ARG;
    if (first == 0 && count == m->L) {                    // every copy of every link, one call per slab
        std::vector<vx_link_state> buf;
        for (int k = 0; k < m->active; k++) {
            const vx_slabbed::Part& p = m->part[k];
            const int ln = (int)p.link_l2g.size();
            if (!ln) continue;
            buf.resize(ln);
            for (int j = 0; j < ln; j++) buf[j] = src[p.link_l2g[j]];
            int rc = vx_upload_link_state(m->slab[k], 0, ln, buf.data()); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_upload_link_state");
        }
        return m->poisson ? vxs::exchange_all(m) : VX_OK;      // the slabs rebuilt their Poisson strains from their own links: ghosts take their owners'
    }
    for (int i = first; i < first + count; i++) {         // the owner's copy and the copy above the cut the link crosses
        int rc = vx_upload_link_state(m->slab[m->link_owner[i]], m->link_local[i], 1, src + (i - first)); if (rc != VX_OK) return vxs::fail_from(m, m->link_owner[i], rc, "vx_upload_link_state");
        if (m->link_slab2[i] >= 0) { rc = vx_upload_link_state(m->slab[m->link_slab2[i]], m->link_local2[i], 1, src + (i - first)); if (rc != VX_OK) return vxs::fail_from(m, m->link_slab2[i], rc, "vx_upload_link_state"); }
    }
    return m->poisson ? vxs::exchange_all(m) : VX_OK;
}

// CVoxelyze::stateInfo (src/Voxelyze.cpp:752-800) of the whole model.  Voxel quantities: every slab reduces its own voxels on
// its device (ghost copies are skipped there) and the results are combined.  Link quantities: a link that crosses a cut lives
// in two slabs, so the per-link values are gathered in the numbering of the whole model (the owner's copy) and reduced here.
int vx_slabbed_state_info(vx_slabbed* m, int info, int type, float* out)
{
    enum { DISPLACEMENT, VELOCITY, KINETIC_ENERGY, ANGULAR_DISPLACEMENT, ANGULAR_VELOCITY, ENG_STRESS, ENG_STRAIN, STRAIN_ENERGY, PRESSURE, MASS };
    enum { MIN, MAX, TOTAL, AVERAGE };
    if (!m || !out || info < 0 || info > MASS || type < 0 || type > AVERAGE) return VX_ERR_ARG;
    *out = 0.f;
    if (m->active == 0) return VX_OK;
    if (m->active == 1) { int rc = vx_state_info(m->slab[0], info, type, out); return rc == VX_OK ? rc : vxs::fail_from(m, 0, rc, "vx_state_info"); }
    float ret = type == MAX ? -3.402823466e38f : (type == MIN ? 3.402823466e38f : 0.f);
    if (info == STRAIN_ENERGY || info == ENG_STRESS || info == ENG_STRAIN) {
        if (m->L == 0) return VX_OK;
        std::vector<float> val(m->L);
        if (info != STRAIN_ENERGY) {
            int rc = vx_slabbed_download(m, info == ENG_STRESS ? VX_F_STRESS : VX_F_STRAIN, 0, m->L, val.data()); if (rc != VX_OK) return rc;
        } else {
            std::vector<double> fn((size_t)3 * m->L), mn((size_t)3 * m->L), mp((size_t)3 * m->L);
            int rc = vx_slabbed_download(m, VX_F_FORCE_NEG, 0, m->L, fn.data());
            if (rc == VX_OK) rc = vx_slabbed_download(m, VX_F_MOMENT_NEG, 0, m->L, mn.data());
            if (rc == VX_OK) rc = vx_slabbed_download(m, VX_F_MOMENT_POS, 0, m->L, mp.data());
            if (rc != VX_OK) return rc;
            int n_mat = 0;
            for (int i = 0; i < m->N; i++) n_mat = std::max(n_mat, (int)m->mat[i] + 1);
            std::vector<vx_linkmat_row> rows((size_t)n_mat * n_mat);
            std::vector<char> have((size_t)n_mat * n_mat, 0);
            for (int l = 0; l < m->L; l++) {                      // CVX_Link::strainEnergy, src/VX_Link.cpp:251-257
                const int a = m->mat[m->lneg[l]], b = m->mat[m->lpos[l]];
                const size_t k = (size_t)std::min(a, b) * n_mat + std::max(a, b);
                if (!have[k]) { rc = vx_get_linkmat(m->slab[0], a, b, &rows[k]); if (rc != VX_OK) return vxs::fail_from(m, 0, rc, "vx_get_linkmat"); have[k] = 1; }
                const vx_linkmat_row& r = rows[k];
                const double fx = fn[3 * (size_t)l], mnx = mn[3 * (size_t)l], mny = mn[3 * (size_t)l + 1], mnz = mn[3 * (size_t)l + 2], mpy = mp[3 * (size_t)l + 1], mpz = mp[3 * (size_t)l + 2];
                val[l] = (float)(fx * fx / (2.0f * r.a1) + mnx * mnx / (2.0 * r.a2) + (mnz * mnz - mnz * mpz + mpz * mpz) / (3.0 * r.b3) + (mny * mny - mny * mpy + mpy * mpy) / (3.0 * r.b3));
            }
        }
        double sum = 0.0;                                         // totals in double, like the one-device library (vx_collide.cuh)
        for (int l = 0; l < m->L; l++) {
            if (type == MIN) { if (val[l] < ret) ret = val[l]; } else if (type == MAX) { if (val[l] > ret) ret = val[l]; } else sum += (double)val[l];
        }
        *out = type == TOTAL ? (float)sum : (type == AVERAGE ? (float)sum / m->L : ret);
        return VX_OK;
    }
    double total = 0.0;
    for (int k = 0; k < m->active; k++) {
        float v = 0.f;
        int rc = vx_state_info(m->slab[k], info, type == AVERAGE ? TOTAL : type, &v); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_state_info");
        if (type == MIN) ret = std::min(ret, v); else if (type == MAX) ret = std::max(ret, v); else total += v;
    }
    *out = type == TOTAL ? (float)total : (type == AVERAGE ? (float)total / m->N : ret);
    return VX_OK;
}

// dynamic-state checkpoint of a slabbed run: one vx_save_state file per slab, "<path>.<k>of<n>" (ghost planes and both
// generations included, so a restored run continues without an exchange).  Restores into a handle built from the same model
// with the same number of slabs.
int vx_slabbed_save_state(vx_slabbed* m, const char* path)
{
    if (!m || !path) return VX_ERR_ARG;
    for (int k = 0; k < m->active; k++) {
        const std::string file = std::string(path) + "." + std::to_string(k) + "of" + std::to_string(m->active);
        int rc = vx_save_state(m->slab[k], file.c_str()); if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_save_state");
    }
    return VX_OK;
}

int vx_slabbed_load_state(vx_slabbed* m, const char* path)
{
    if (!m || !path) return VX_ERR_ARG;
    for (int k = 0; k < m->active; k++) {
        const std::string file = std::string(path) + "." + std::to_string(k) + "of" + std::to_string(m->active);
        int rc = vx_load_state(m->slab[k], file.c_str());
        if (rc != VX_OK) return vxs::fail_from(m, k, rc, "vx_load_state");      // (slabs before k are restored already: reload or reset)
    }
    return VX_OK;
}

int64_t vx_slabbed_launch_count(const vx_slabbed* m)
{
    int64_t n = 0;
    if (m) for (int k = 0; k < m->active; k++) n += vx_launch_count(m->slab[k]);
    return n;
}

} // extern "C"

#endif // VX_SLABBED_HPP
