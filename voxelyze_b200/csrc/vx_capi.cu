// vx_capi.cu -- implementation of include/voxelyze_b200.h for sm_100a (the product).
//
// Host side of the drop-in boundary: owns the device-resident structure-of-arrays state of
// one simulation, translates the flat model description into it, and drives the kernels.
// Two device layouts exist behind the same ABI:
//   lattice mode  bodies that fill >= 62.5 % of their bounding box (holes padded with inert cells), nu = 0,
//                 self-collisions included: fused single-pass kernel on ping-pong generations,
//                 vx_lattice.cuh  (headline path)
//   general mode  any voxel set, Poisson materials: link kernels + voxel kernel, vx_kernels.cuh
// There is deliberately no CPU code path for stepping: without a usable CUDA device
// vx_create fails with VX_ERR_NO_DEVICE.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "voxelyze_b200.h"
#include "vx_material.hpp"
#include "vx_lattice.cuh"
#include "vx_collide.cuh"
#include "vx_mesh.cuh"
#include "vx_linsolve.cuh"

using namespace vxd;

namespace {

template <typename T> struct DevView { T* p = nullptr; };      // a part of another buffer's allocation
template <typename T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct LinkMatEntry { int a, b; vxm::Material mat; };

#define VF_FILL 0x80000000u               // internal voxel flag: inert cell that fills a hole of the bounding box
constexpr int GRAPH_STEPS = 16;     // steps per captured CUDA graph (even: generations re-align)
constexpr int TPB = 128;

inline int blocks_for(long long n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

// NVTX range around the entry points a timeline should show (nsys / ncu --nvtx); costs nothing without a tool attached
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

} // namespace

struct vx_sim {
    int device = 0, sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    double vox_size = 0.001;
    std::string err;

    // ---- model (host)
    std::vector<vx_material_desc> descs; std::vector<std::vector<float>> d_eps, d_sig;
    std::vector<vxm::Material> mats;
    std::vector<LinkMatEntry> lmats; std::map<std::pair<int, int>, int> lmat_of;
    bool any_poisson = false;

    int N = 0, L = 0, n_members = 1;
    // lattice mode: ensemble members are tiled pack[0] x pack[1] x pack[2] into ONE device lattice (their link masks keep
    // them apart), lat_members = 1; unpacked ensembles are stacked along z as lat_members separate boxes
    int pack[3] = {1, 1, 1}, lat_members = 1;
    int N_user = 0;                     // voxels the caller created; [N_user, N) are inert fill cells of a box with holes (lattice mode)
    std::vector<int32_t> ijk; std::vector<uint16_t> vmat_id; std::vector<int32_t> member; std::vector<uint32_t> vflags;
    std::vector<int32_t> lk_vn, lk_vp; std::vector<uint8_t> lk_axis;   // caller (creation) order, caller voxel indices
    std::vector<int32_t> v_e2i, v_i2e, l_e2i, l_i2e;
    std::vector<uint8_t> linkmask;                                     // by caller voxel index
    std::vector<uint16_t> lk_mat;                                      // by internal link index (general mode)
    int axis_first[4] = {0, 0, 0, 0};
    std::vector<int64_t> sort_key;                                     // by internal voxel index: (member,z)

    std::vector<int32_t> ext_vox; std::vector<DevExt> ext_rows;        // externals, caller voxel indices

    float grav = 0.f, ambient = 0.f, envelope = 0.625f;
    // lattice mode: vx_set_temperature_all is applied by the first fused kernel of the next vx_step (LatFrame::amb) instead
    // of by a pass of its own; flush_ambient() materialises it for every other reader.  last_amb: the last executed
    // step was such a step (its inputs, generation gen^1, still hold the old temperatures)
    bool amb_pending = false, last_amb = false; float amb_value = 0.f, last_amb_value = 0.f;
    bool amb_each_step = false;         // vx_step_ambient is queueing: every step applies the ambient temperature it is launched with
    bool floor_on = false, collisions = false;
    float time_host = 0.f;
    int path = 0;                       // vx_set_path: 0 auto, 1 general, 3 small-model cluster kernel, 5 / 7 fused lattice with cp.async / TMA staging
    bool small = false;                 // general layout stepped by k_small_steps (one launch per vx_step call)
    bool relayout = false;

    // ---- lattice mode
    bool lattice = false;
    int nx = 0, ny = 0, nz = 0;
    int gen = 0;                        // generation that holds the current state
    int gen_view = -1;                  // inside a multi-step lattice call: the generation frame() shows (collision kernels)
    bool have_prev = false;             // gen^1 holds the inputs of the last executed step
    float last_prev_dt = 0.f;           // previousDt that step used
    bool call_per_step_dt = false;      // the running / last stepping call re-evaluated dt before every step (Poisson models, dt < 0)
    float prev_dt_host = 0.f;           // mirror of DevParams::prev_dt
    // asynchronous call (vx_step_begin .. vx_step_end), used by z-slab runs to overlap the halo exchange
    bool call_active = false, call_half = false;
    int call_g0 = 0, call_done = 0;     // starting generation, steps whose boundary part has been enqueued
    std::vector<int> zb_layers;         // brick-group layers (4 planes each) that hold ghost planes or their neighbours
    // z-slabs on the TMA-staged kernel: the all-ghost planes at the ends are not covered by bricks (k_lattice_tma<.., GSKIP>);
    // bricks, brick-group layers and zb_layers then count from plane z_lo
    bool ghost_skip = false; int z_lo = 0, z_hi = 0;
    // peer-memory halo (vx_peer_*): my boundary layer -> the ghost layer of the neighbouring slab, over NVLink
    struct PeerLink {
        size_t src_first = 0, count = 0;                  // my layer (internal voxel range)
        double4* dst0[2] = {nullptr, nullptr};            // the peer's ghost layer in its pose0/pose1 arrays, per generation
        double4* dst1[2] = {nullptr, nullptr};
        float4* dst_ps[2] = {nullptr, nullptr};           // Poisson models: the peer's ghost layer in its pStrain arrays, per generation
        int* dst_flag = nullptr;                          // the peer's arrival counter for messages from me
        void* opened[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};     // cudaIpcOpenMemHandle results (other-process peers)
    };
    std::vector<PeerLink> peers;
    bool wb_opted_in = false, capturing = false;
    bool launch_failed = false;                     // a step kernel could not be launched at all (reported by the call that queued it)
    bool state_ready = false;                       // the device arrays hold a valid state (false while vx_set_voxels rebuilds them)
    DevBuf<int> group_list; int n_groups = 0;   // sparse bodies: the occupied brick groups (LatFrame::groups), 0 = launch the whole bounding box
    DevBuf<unsigned char> tmaps;        // CUtensorMap descriptors of the lattice arrays (k_lattice_tma), rebuilt with the arrays
    bool push_in_kernel = false;        // set around the boundary launches of vx_slab_step
    DevBuf<int> peer_flags;             // [0] arrivals from the slab below, [1] from the slab above, [2] time-out marker
    int n_expect[2] = {0, 0};           // exchanges a neighbour on that side takes part in (0 or 1 per exchange)
    bool expect_side[2] = {false, false};
    int xseq = 0;                       // exchanges completed or queued so far (same on all slabs)
    cudaStream_t comm_stream = nullptr; cudaEvent_t ev_boundary = nullptr, ev_comm = nullptr;
    int newest_gen() const { return call_active ? (call_g0 + call_done) & 1 : gen; }

    // ---- device
    // pose0 and pose1 of a generation are ONE allocation ([2][N] double4, pose1 = pose0 + N) so that a single 4-D TMA
    // box brings both records of a voxel set; general mode uses [0] only
    DevBuf<double4> pose0[2], mom0[2]; DevView<double4> pose1[2]; DevBuf<double2> mom1[2];
    // lattice link records, one allocation of VX_REC_PARTS x N sixteen-byte parts per generation: for axis a the
    // parts 4a..4a+2 are the three double2 of the record, part 4a+3 its float4 {strain, maxStrain, strainOffset, stress}
    DevBuf<double2> rec[2];
    DevBuf<float4> ps[2];                               // lattice mode with Poisson materials: CVX_Voxel::pStrain per generation
    cudaError_t alloc_pose(int g, size_t n1) { cudaError_t e = pose0[g].alloc(2 * n1); pose1[g].p = e == cudaSuccess ? pose0[g].p + n1 : nullptr; return e; }
    void release_pose(int g) { pose0[g].release(); pose1[g].p = nullptr; }
    DevBuf<uint16_t> pair_lmat; DevBuf<int> link_owner; DevBuf<unsigned char> link_axis_dev;
    DevBuf<int> ext_idx, ext_vox_dev, vox_e2i_dev, link_e2i_dev, member_dev;
    DevBuf<float4> pstrain; DevBuf<double> slots; DevBuf<float> slot_strain;
    DevBuf<int2> lends; DevBuf<uint32_t> lmeta; DevBuf<double4> lstA, lstB; DevBuf<double> lstC; DevBuf<float4> lstrain;
    DevBuf<DevVoxMat> vmat_dev; DevBuf<DevLinkMat> lmat_dev; DevBuf<float> curve_e, curve_s;
    DevBuf<DevExt> ext_dev;
    DevBuf<DevParams> params; DevBuf<unsigned int> freq2; DevBuf<float> member_t;
    DevBuf<unsigned char> staging;
    DevParams* params_host = nullptr;     // pinned mirror
    unsigned int* freq_host = nullptr;    // pinned
    VoxelStateRec* probe_host = nullptr;  // pinned + mapped: vx_download_voxel_state of a few voxels lands here without a copy

    // ---- collisions (tables indexed by internal voxel index: both layouts)
    std::vector<int32_t> nbr;                                      // [N][6] neighbour voxel (caller index) or -1
    std::vector<int> ext_raw_vox; std::vector<uint8_t> ext_raw_dof; std::vector<float> ext_raw_f, ext_raw_m; std::vector<double> ext_raw_t, ext_raw_r;
    int n_surf = 0, n_pairs = 0, col_cap = 0, hash_size = 0;
    bool col_tables = false, col_stale_host = true;
    DevBuf<int> c_surf_vox, c_surf_orig, c_surf_member, c_slot; DevBuf<short4> c_surf_ijk; DevBuf<uint32_t> c_nearby;
    DevBuf<float4> c_last_watch; DevBuf<int> c_cell_count, c_cell_start, c_sorted; DevBuf<int4> c_cell;
    DevBuf<int2> c_pairs; DevBuf<float2> c_pair_kc; DevBuf<float4> c_pair_force;
    DevBuf<int> c_counters, c_deg, c_ref_start, c_ref_fill, c_refs;
    int* counters_host = nullptr;                                  // pinned, CC_COUNT ints: mirror of c_counters after every step call
    int col_rebuilds = 0;                                          // watch-list rebuilds since vx_set_voxels
    cudaStream_t aux_stream = nullptr;                             // capture stream of conditional-node bodies
    int cond_nodes = 1;                                            // 1: rebuild chain inside a conditional IF node of the step graphs; 0: predicated kernels only
    // stateInfo reductions
    DevBuf<float> si_minmax; DevBuf<double> si_sum; DevBuf<double4> si_nominal; DevBuf<float> si_consts; DevBuf<unsigned char> si_buf;
    bool si_nominal_ok = false, si_consts_ok = false, si_pressure_ok = false;
    DevBuf<int> si_vlinks; DevBuf<float> si_ratio; DevBuf<float2> si_en; DevBuf<unsigned char> si_skip;
    // surface mesh (vx_mesh.inl): topology tables built by vx_mesh_build, float buffers refreshed by vx_mesh_update
    struct Mesh {
        bool built = false; int n_vert = 0, n_quad = 0;
        DevBuf<int> vert_vox, quads, quad_vox; DevBuf<float> vertices, normals, colors, strain, max_strain, eps_fail, eps_yield, mat_rgb, vals;
        std::vector<float> rgb_host;
        void release() { vert_vox.release(); quads.release(); quad_vox.release(); vertices.release(); normals.release(); colors.release(); strain.release();
                         max_strain.release(); eps_fail.release(); eps_yield.release(); mat_rgb.release(); vals.release(); built = false; n_vert = n_quad = 0; }
    } mesh;

    bool uni = false; DevVoxMat vm0{}; DevLinkMat lm0{};          // single-material model: rows passed by value
    cudaGraphExec_t graph = nullptr; int graph_kernels = 0;      // general mode
    cudaGraphExec_t lgraph[2] = {nullptr, nullptr};              // lattice mode, keyed by starting generation
    int lgraph_kernels = GRAPH_STEPS;                            // launches one lattice graph stands for
    int64_t launches = 0;

    Frame frame() const
    {
        Frame f{};
        const int g = lattice ? (gen_view >= 0 ? gen_view : gen) : 0;
        f.n_vox = N; f.n_link = L;
        f.pose0 = pose0[g].p; f.pose1 = pose1[g].p; f.mom0 = mom0[g].p; f.mom1 = mom1[g].p;
        f.ext_idx = ext_idx.p; f.pstrain = lattice ? ps[g].p : pstrain.p; f.slots = slots.p; f.slot_strain = slot_strain.p;
        f.lends = lends.p; f.lmeta = lmeta.p; f.lstA = lstA.p; f.lstB = lstB.p; f.lstC = lstC.p; f.lstrain = lstrain.p;
        f.vmat = vmat_dev.p; f.lmat = lmat_dev.p; f.curve_e = curve_e.p; f.curve_s = curve_s.p;
        f.ext = ext_dev.p; f.params = params.p;
        f.col_slot = c_slot.p; f.col_start = c_ref_start.p; f.col_ref = c_refs.p; f.col_force = c_pair_force.p;
        f.vm0 = vm0; f.lm0 = lm0;
        return f;
    }
    ColFrame col_frame() const
    {
        ColFrame c{};
        c.n_surf = n_surf; c.surf_vox = c_surf_vox.p; c.surf_orig = c_surf_orig.p; c.surf_member = c_surf_member.p;
        c.surf_ijk = c_surf_ijk.p; c.nearby = c_nearby.p; c.last_watch = c_last_watch.p;
        c.cell_count = c_cell_count.p; c.cell_start = c_cell_start.p; c.sorted = c_sorted.p; c.cell = c_cell.p; c.hash_mask = hash_size - 1;
        c.params = params.p; c.parity = lattice ? (gen_view >= 0 ? gen_view : gen) : -1;
        c.pairs = c_pairs.p; c.pair_kc = c_pair_kc.p; c.pair_force = c_pair_force.p; c.cap = col_cap;
        c.counters = c_counters.p; c.deg = c_deg.p; c.ref_start = c_ref_start.p; c.ref_fill = c_ref_fill.p; c.refs = c_refs.p;
        // watch radius and re-watch distance of CVoxelyze::updateCollisions (src/Voxelyze.cpp:672-674)
        const float watch_vx = 2 * 0.75f + 1.0f;
        const float watch_mm = (float)(vox_size * watch_vx);
        const float recalc = (float)(vox_size * 1.0f / 2);
        c.inv_cell = 1.0 / (double)watch_mm;
        c.thresh_sq = watch_mm * watch_mm; c.recalc_sq = recalc * recalc; c.envelope = envelope;
        return c;
    }
    // lattice frame reading generation g and writing generation g^1
    LatFrame lat_frame(int g) const
    {
        LatFrame f{};
        f.nx = nx; f.ny = ny; f.nz = nz; f.nxy = nx * ny; f.n_vox = N; f.n_mat = (int)mats.size(); f.n_lmat = (int)lmats.size();
        f.c_pose0 = pose0[g].p; f.c_pose1 = pose1[g].p; f.c_mom0 = mom0[g].p; f.c_mom1 = mom1[g].p;
        f.n_pose0 = pose0[g ^ 1].p; f.n_pose1 = pose1[g ^ 1].p; f.n_mom0 = mom0[g ^ 1].p; f.n_mom1 = mom1[g ^ 1].p;
        for (int a = 0; a < 3; a++) {
            for (int k = 0; k < 3; k++) {
                f.c_rec[a][k] = rec[g].p + (size_t)(a * 4 + k) * N;
                f.n_rec[a][k] = rec[g ^ 1].p + (size_t)(a * 4 + k) * N;
            }
            f.c_recf[a] = reinterpret_cast<const float4*>(rec[g].p + (size_t)(a * 4 + 3) * N);
            f.n_recf[a] = reinterpret_cast<float4*>(rec[g ^ 1].p + (size_t)(a * 4 + 3) * N);
        }
        f.ext_idx = ext_idx.p; f.vmat = vmat_dev.p; f.lmat = lmat_dev.p; f.curve_e = curve_e.p; f.curve_s = curve_s.p;
        f.pair_lmat = pair_lmat.p; f.ext = ext_dev.p; f.params = params.p;
        f.vm0 = vm0; f.lm0 = lm0;
        f.col_slot = (collisions && col_tables) ? c_slot.p : nullptr;
        f.col_start = c_ref_start.p; f.col_ref = c_refs.p; f.col_force = c_pair_force.p;
        f.amb_set = 0; f.amb = 0.f;
        f.groups = n_groups > 0 ? group_list.p : nullptr;
        f.c_ps = any_poisson ? ps[g].p : nullptr; f.n_ps = any_poisson ? ps[g ^ 1].p : nullptr;
        f.z_lo = ghost_skip ? z_lo : 0; f.z_hi = ghost_skip ? z_hi : nz;
        f.stream_out = (size_t)N * 608 > ((size_t)96 << 20) && !getenv("VX_NO_STREAM_STORES");      // 608 B of state per voxel against the 126 MB L2
        f.push_z[0] = f.push_z[1] = -1; f.push_ps[0] = f.push_ps[1] = nullptr;
        if (push_in_kernel) {
            for (size_t k = 0; k < peers.size() && k < 2; k++) {
                f.push_z[k] = (int)(peers[k].src_first / ((size_t)nx * ny));
                f.push0[k] = peers[k].dst0[g ^ 1]; f.push1[k] = peers[k].dst1[g ^ 1]; f.push_ps[k] = peers[k].dst_ps[g ^ 1];
            }
        }
        return f;
    }
    void drop_graph()
    {
        if (graph) { cudaGraphExecDestroy(graph); graph = nullptr; }
        for (int g = 0; g < 2; g++) if (lgraph[g]) { cudaGraphExecDestroy(lgraph[g]); lgraph[g] = nullptr; }
    }
};

static int fail(vx_sim* s, int code, const std::string& msg) { if (s) s->err = msg; return code; }
static int cuda_fail(vx_sim* s, cudaError_t e, const char* what)
{
    if (s) s->err = std::string(what) + ": " + cudaGetErrorString(e);
    return VX_ERR_CUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(s, e_, #call); } while (0)

// ------------------------------------------------------------------------------------------------
// material tables
static int link_material(vx_sim* s, int a, int b)
{
    std::pair<int, int> key(std::min(a, b), std::max(a, b));
    auto it = s->lmat_of.find(key);
    if (it != s->lmat_of.end()) return it->second;
    LinkMatEntry e; e.a = key.first; e.b = key.second;
    e.mat = vxm::combine(s->mats[key.first], s->mats[key.second]);
    s->lmats.push_back(e);
    int id = (int)s->lmats.size() - 1;
    s->lmat_of[key] = id;
    return id;
}

static int upload_tables(vx_sim* s)
{
    CK(cudaSetDevice(s->device));
    std::vector<DevVoxMat> vm(s->mats.size());
    s->any_poisson = false;
    for (size_t i = 0; i < vm.size(); i++) {
        const vxm::Material& m = s->mats[i];
        vxm::MassProps p = vxm::mass_props(m, s->vox_size);
        DevVoxMat& d = vm[i];
        for (int a = 0; a < 3; a++) d.size[a] = s->vox_size * m.ext_scale[a];
        d.nom = s->vox_size;
        d.cte = m.cte; d.mass = p.mass; d.mass_inv = p.mass_inv; d.inertia_inv = p.inertia_inv;
        d.E = m.E; d.nu = m.nu;
        d.two_sqrtm_zeta = 2 * p.sqrt_mass * m.zeta_int;
        d.glob_damp_t = m.zeta_glob * p.two_sq_mes;
        d.glob_damp_r = m.zeta_glob * p.two_sq_ies3;
        d.coll_damp_t = m.zeta_coll * p.two_sq_mes;
        d.pen_stiff = (float)(2 * m.E * s->vox_size);
        d.mu_s = m.mu_s; d.mu_k = m.mu_k;
        d.gravity_force = -p.mass * 9.80665f * s->grav;
        d.nom_f = (float)s->vox_size;
        d.pad = 0;
        d.mass_inv_d = d.mass_inv; d.inertia_inv_d = d.inertia_inv; d.glob_damp_t_d = d.glob_damp_t;
        d.glob_damp_r_d = d.glob_damp_r; d.coll_damp_t_d = d.coll_damp_t; d.gravity_force_d = d.gravity_force;
        if (m.nu != 0.0f) s->any_poisson = true;
    }
    std::vector<DevLinkMat> lm(s->lmats.size());
    std::vector<float> ce, cs;
    const int nm = (int)s->mats.size();
    std::vector<uint16_t> pair((size_t)std::max(nm * nm, 1), 0);
    for (size_t i = 0; i < lm.size(); i++) {
        LinkMatEntry& e = s->lmats[i];
        e.mat = vxm::combine(s->mats[e.a], s->mats[e.b]);
        const vxm::Material& m = e.mat;
        vxm::BeamConsts k = vxm::beam_consts(m, s->vox_size);
        DevLinkMat& d = lm[i];
        d.linear = m.linear ? 1 : 0;
        d.curve_off = (int)ce.size(); d.curve_n = (int)m.eps.size();
        ce.insert(ce.end(), m.eps.begin(), m.eps.end());
        cs.insert(cs.end(), m.sig.begin(), m.sig.end());
        d.E = m.E; d.nu = m.nu; d.e_hat = m.e_hat; d.eps_yield = m.eps_yield; d.eps_fail = m.eps_fail;
        d.pad = 0; d.a1 = k.a1; d.a2 = k.a2; d.b1 = k.b1; d.b2 = k.b2; d.b3 = k.b3;
        d.sq_a1 = k.sq_a1; d.sq_a2_ip = k.sq_a2_ip; d.sq_b1 = k.sq_b1; d.sq_b2_fmp = k.sq_b2_fmp; d.sq_b3_ip = k.sq_b3_ip;
        if (m.nu != 0.0f) s->any_poisson = true;
        pair[(size_t)e.a * nm + e.b] = pair[(size_t)e.b * nm + e.a] = (uint16_t)i;
    }
    CK(s->vmat_dev.alloc(std::max<size_t>(vm.size(), 1)));
    CK(s->lmat_dev.alloc(std::max<size_t>(lm.size(), 1)));
    CK(s->curve_e.alloc(std::max<size_t>(ce.size(), 2)));
    CK(s->curve_s.alloc(std::max<size_t>(cs.size(), 2)));
    CK(s->pair_lmat.alloc(pair.size()));
    CK(cudaStreamSynchronize(s->stream));
    if (!vm.empty()) CK(cudaMemcpy(s->vmat_dev.p, vm.data(), vm.size() * sizeof(DevVoxMat), cudaMemcpyHostToDevice));
    if (!lm.empty()) CK(cudaMemcpy(s->lmat_dev.p, lm.data(), lm.size() * sizeof(DevLinkMat), cudaMemcpyHostToDevice));
    if (!ce.empty()) {
        CK(cudaMemcpy(s->curve_e.p, ce.data(), ce.size() * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s->curve_s.p, cs.data(), cs.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(s->pair_lmat.p, pair.data(), pair.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    s->uni = vm.size() == 1 && lm.size() == 1 && lm[0].linear && !s->any_poisson;
    if (s->uni) { s->vm0 = vm[0]; s->lm0 = lm[0]; }
    s->si_consts_ok = false;
    s->drop_graph();
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------
// initial dynamic state (fresh CVoxelyze / resetTime)
static uint32_t initial_bits(const vx_sim* s, int caller_voxel)
{
    uint32_t b = s->vmat_id[caller_voxel] & VM_MAT_MASK;
    b |= (uint32_t)s->linkmask[caller_voxel] << VM_LINK_SHIFT;
    b |= VM_STATIC_FRIC | VM_PSTRAIN_STALE;                         // CVX_Voxel::reset, src/VX_Voxel.cpp:47-56
    if (!s->vflags.empty() && (s->vflags[caller_voxel] & VX_VF_GHOST)) b |= VM_GHOST;
    if (s->lattice) b |= 0x15u << VM_LFLAG_SHIFT;                   // three owned links start in small-angle mode
    return b;
}

static int upload_initial_state(vx_sim* s, float temp)
{
    const int N = s->N, L = s->L;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    s->gen = 0; s->have_prev = false; s->prev_dt_host = 0.f; s->col_stale_host = true;
    s->amb_pending = false; s->last_amb = false;
    if (N) {
        std::vector<double4> p0(N), p1(N);
        std::vector<char> has_ext(N, 0);
        for (int v : s->ext_vox) has_ext[v] = 1;
        for (int i = 0; i < N; i++) {
            int e = s->v_i2e[i];
            double sz = s->vox_size;
            p0[i] = make_double4(s->ijk[3 * e] * sz, s->ijk[3 * e + 1] * sz, s->ijk[3 * e + 2] * sz, 1.0);
            uint32_t bits = initial_bits(s, e) | (has_ext[e] ? VM_HAS_EXT : 0u);
            unsigned long long w = ((unsigned long long)bits << 32);
            uint32_t tb; memcpy(&tb, &temp, 4); w |= tb;
            double wd; memcpy(&wd, &w, 8);
            p1[i] = make_double4(0.0, 0.0, 0.0, wd);
        }
        const int gens = s->lattice ? 2 : 1;
        for (int g = 0; g < gens; g++) {
            CK(cudaMemcpy(s->pose0[g].p, p0.data(), (size_t)N * sizeof(double4), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(s->pose1[g].p, p1.data(), (size_t)N * sizeof(double4), cudaMemcpyHostToDevice));
            CK(cudaMemset(s->mom0[g].p, 0, (size_t)N * sizeof(double4)));
            CK(cudaMemset(s->mom1[g].p, 0, (size_t)N * sizeof(double2)));
            if (s->lattice) {
                CK(cudaMemset(s->rec[g].p, 0, (size_t)N * VX_REC_PARTS * sizeof(double2)));
            }
        }
        if (!s->lattice) {
            CK(cudaMemset(s->slots.p, 0, (size_t)N * 36 * sizeof(double)));
            CK(cudaMemset(s->pstrain.p, 0, (size_t)N * sizeof(float4)));
            CK(cudaMemset(s->slot_strain.p, 0, (size_t)N * 6 * sizeof(float)));
        }
    }
    if (L && !s->lattice) {
        CK(cudaMemset(s->lstA.p, 0, (size_t)L * sizeof(double4)));
        CK(cudaMemset(s->lstB.p, 0, (size_t)L * sizeof(double4)));
        CK(cudaMemset(s->lstC.p, 0, (size_t)L * sizeof(double)));
        CK(cudaMemset(s->lstrain.p, 0, (size_t)L * sizeof(float4)));
        std::vector<uint32_t> lm(L);
        for (int i = 0; i < L; i++) lm[i] = s->lk_mat[i] | LM_SMALL_ANGLE;    // CVX_Link::reset, src/VX_Link.cpp:61-75
        CK(cudaMemcpy(s->lmeta.p, lm.data(), (size_t)L * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    for (int g = 0; g < 2; g++) if (s->lattice && s->ps[g].p && N) CK(cudaMemset(s->ps[g].p, 0, (size_t)N * sizeof(float4)));   // zero strains: zero pStrain
    DevParams p{}; p.col_stale = 1;
    CK(cudaMemcpy(s->params.p, &p, sizeof(p), cudaMemcpyHostToDevice));
    s->time_host = 0.f;
    s->drop_graph();
    s->state_ready = true;
    return VX_OK;
}

// lattice mode with Poisson materials: (re)computes CVX_Voxel::pStrain of the current generation from the link strains in
// its records -- after Poisson's ratio was switched on, after externals or link state were replaced
static int refresh_lattice_ps(vx_sim* s)
{
    if (!s->lattice || !s->any_poisson || s->N == 0 || !s->state_ready) return VX_OK;
    CK(cudaSetDevice(s->device));
    const size_t n1 = (size_t)s->N;
    for (int g = 0; g < 2; g++)
        if (!s->ps[g].p) { CK(s->ps[g].alloc(n1)); CK(cudaMemsetAsync(s->ps[g].p, 0, n1 * sizeof(float4), s->stream)); }
    k_lattice_pstrain_init<<<blocks_for(s->N), TPB, 0, s->stream>>>(s->lat_frame(s->gen), s->ps[s->gen].p); s->launches++;
    CK(cudaGetLastError());
    s->drop_graph();                                    // frames now carry the pStrain arrays
    return VX_OK;
}

// a pending vx_set_temperature_all written into the voxel records (for every reader that is not the fused step)
static int flush_ambient(vx_sim* s)
{
    if (!s->amb_pending) return VX_OK;
    s->amb_pending = false;
    if (s->N == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    k_fill_temp<<<blocks_for(s->N), TPB, 0, s->stream>>>(s->frame(), s->amb_value, nullptr, nullptr); s->launches++;
    CK(cudaGetLastError());
    return VX_OK;
}
// the inputs of the last executed step, as that step saw them (k_lattice_gather_links recomputes its link forces)
static LatFrame prev_frame(const vx_sim* s)
{
    LatFrame f = s->lat_frame(s->gen ^ 1);
    if (s->last_amb) { f.amb_set = 1; f.amb = s->last_amb_value; }
    return f;
}

static int upload_externals(vx_sim* s)
{
    if (s->N == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    Frame f = s->frame();
    k_clear_ext_bits<<<blocks_for(s->N), TPB, 0, s->stream>>>(f); s->launches++;
    int n = (int)s->ext_vox.size();
    if (n) {
        std::vector<int> internal(n);
        for (int k = 0; k < n; k++) internal[k] = s->v_e2i[s->ext_vox[k]];
        CK(s->ext_dev.alloc(n)); CK(s->ext_vox_dev.alloc(n));
        CK(cudaStreamSynchronize(s->stream));
        CK(cudaMemcpy(s->ext_dev.p, s->ext_rows.data(), (size_t)n * sizeof(DevExt), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s->ext_vox_dev.p, internal.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        f = s->frame();
        k_set_ext_bits<<<blocks_for(n), TPB, 0, s->stream>>>(f, n, s->ext_vox_dev.p, s->ext_idx.p); s->launches++;
    }
    CK(cudaGetLastError());
    s->drop_graph();      // the ext table pointer may have moved
    return refresh_lattice_ps(s);                       // Poisson: which axes count as "in tension" depends on the externals
}

// ------------------------------------------------------------------------------------------------
// stepping, general mode
static void launch_links(vx_sim* s, const Frame& f)
{
    const int* af = s->axis_first;
    const bool P = s->any_poisson;
    for (int a = 0; a < 3; a++) {
        int cnt = af[a + 1] - af[a];
        if (cnt <= 0) continue;
        int g = blocks_for(cnt);
        const bool U = s->uni;
#define VX_LAUNCH_LINK(AX) do { \
        if (P) k_link<AX, true, false><<<g, TPB, 0, s->stream>>>(af[a], cnt, f); \
        else if (U) k_link<AX, false, true><<<g, TPB, 0, s->stream>>>(af[a], cnt, f); \
        else k_link<AX, false, false><<<g, TPB, 0, s->stream>>>(af[a], cnt, f); } while (0)
        if (a == 0) VX_LAUNCH_LINK(0);
        if (a == 1) VX_LAUNCH_LINK(1);
        if (a == 2) VX_LAUNCH_LINK(2);
#undef VX_LAUNCH_LINK
        s->launches++;
    }
}

static void launch_recommended_dt(vx_sim* s, int gen = -1)
{
    cudaMemsetAsync(s->freq2.p, 0, sizeof(unsigned int), s->stream);
    if (s->lattice) {
        int g = std::min(blocks_for(s->N, 256), 148 * 8);
        if (s->L > 0) k_lattice_max_freq<<<g, 256, 0, s->stream>>>(s->lat_frame(gen >= 0 ? gen : s->gen), s->freq2.p);
        else k_max_freq_voxels<<<g, 256, 0, s->stream>>>(s->frame(), s->freq2.p);
    } else if (s->L > 0) {
        int g = std::min(blocks_for(s->L, 256), 148 * 8);
        k_max_freq<<<g, 256, 0, s->stream>>>(s->frame(), s->axis_first[1], s->axis_first[2], s->freq2.p);
    } else {
        int g = std::min(blocks_for(s->N, 256), 148 * 8);
        k_max_freq_voxels<<<g, 256, 0, s->stream>>>(s->frame(), s->freq2.p);
    }
    s->launches++;
}

#include "vx_collide.inl"

static void launch_voxel(vx_sim* s, const Frame& f)
{
    if (s->uni) k_voxel<true><<<blocks_for(s->N), TPB, 0, s->stream>>>(f, s->floor_on ? 1 : 0, s->collisions ? 1 : 0);
    else k_voxel<false><<<blocks_for(s->N), TPB, 0, s->stream>>>(f, s->floor_on ? 1 : 0, s->collisions ? 1 : 0);
    s->launches++;
}

// one doTimeStep (src/Voxelyze.cpp:251-284): links, [divergence test inside the kernels], collisions, voxels;
// per_step_dt: dt < 0 with Poisson materials; capturing: the stream is being captured into a step graph
static int launch_step(vx_sim* s, const Frame& f, bool per_step_dt, bool capturing = false)
{
    if (s->any_poisson) { k_pstrain<<<blocks_for(s->N), TPB, 0, s->stream>>>(f); s->launches++; }
    if (per_step_dt) { launch_recommended_dt(s); k_dt_from_freq<<<1, 1, 0, s->stream>>>(s->freq2.p, s->params.p, 0); s->launches++; }
    launch_links(s, f);
    if (s->collisions) { int rc = enqueue_collision_step(s, capturing); if (rc != VX_OK) return rc; }
    launch_voxel(s, f);
    return VX_OK;
}

// k_small_steps: one cluster of up to 16 CTAs, all n steps
static int launch_small_steps(vx_sim* s, const Frame& f, int n_steps)
{
    // launch shape (profiles/r2_small_model.md): few warps per SM -- a step is the dependent FP64 chain of one link plus one voxel
    // update, so spreading the threads over up to 16 SMs (non-portable cluster size, checked once) beats filling 4 of them
    const int work = std::max(s->N, s->L);
    const int tpb = work <= 1024 ? 64 : 128;
    static int max_ctas = 0;
    if (!max_ctas) {
        max_ctas = 8;
        bool ok = true;
        ok &= cudaFuncSetAttribute(k_small_steps<true, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        ok &= cudaFuncSetAttribute(k_small_steps<false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        ok &= cudaFuncSetAttribute(k_small_steps<false, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        cudaLaunchConfig_t probe = {};
        probe.gridDim = dim3(16); probe.blockDim = dim3(128);
        cudaLaunchAttribute pa[1];
        pa[0].id = cudaLaunchAttributeClusterDimension; pa[0].val.clusterDim.x = 16; pa[0].val.clusterDim.y = 1; pa[0].val.clusterDim.z = 1;
        probe.attrs = pa; probe.numAttrs = 1;
        int clusters = 0;
        if (ok && cudaOccupancyMaxActiveClusters(&clusters, k_small_steps<false, false>, &probe) == cudaSuccess && clusters > 0) max_ctas = 16;
        cudaGetLastError();
    }
    int ctas = std::max(1, std::min(max_ctas, (work + tpb - 1) / tpb));
    while (ctas & (ctas - 1)) ctas++;                               // 1, 2, 4, 8 or 16
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3((unsigned)tpb); cfg.dynamicSmemBytes = 0; cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const int af1 = s->axis_first[1], af2 = s->axis_first[2], fl = s->floor_on ? 1 : 0;
    cudaError_t e;
    if (s->any_poisson) e = cudaLaunchKernelEx(&cfg, k_small_steps<true, false>, f, af1, af2, n_steps, fl);
    else if (s->uni) e = cudaLaunchKernelEx(&cfg, k_small_steps<false, true>, f, af1, af2, n_steps, fl);
    else e = cudaLaunchKernelEx(&cfg, k_small_steps<false, false>, f, af1, af2, n_steps, fl);
    s->launches++;
    if (e != cudaSuccess) return cuda_fail(s, e, "k_small_steps");
    return VX_OK;
}

__global__ void k_begin(DevParams* p, float dt, int set_dt)
{
    p->div_now = 0; p->div_latched = 0; p->steps_done = 0;
    p->pending = 0; p->div_flag[0] = 0; p->div_flag[1] = 0;
    if (set_dt) p->dt = dt;
}

static int ensure_graph(vx_sim* s)
{
    if (s->graph) return VX_OK;
    Frame f = s->frame();
    cudaGraph_t g = nullptr;
    int64_t before = s->launches;
    // capture on the library's own stream (the caller's stream may be the legacy default stream,
    // which cannot be captured); the instantiated graph is launched into the caller's stream
    cudaStream_t user = s->stream; s->stream = s->own_stream;
    cudaError_t e0 = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal);
    if (e0 != cudaSuccess) { s->stream = user; return cuda_fail(s, e0, "cudaStreamBeginCapture"); }
    int rc = VX_OK;
    for (int k = 0; k < GRAPH_STEPS && rc == VX_OK; k++) rc = launch_step(s, f, false, true);
    cudaError_t e = cudaStreamEndCapture(s->stream, &g);
    s->stream = user;
    s->graph_kernels = (int)(s->launches - before);
    s->launches = before;
    if (rc != VX_OK) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
    if (e != cudaSuccess) return cuda_fail(s, e, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(&s->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { s->graph = nullptr; return cuda_fail(s, e, "cudaGraphInstantiate"); }
    return VX_OK;
}

// ------------------------------------------------------------------------------------------------
// stepping, lattice mode
#ifndef VX_TMA_L2PROMO
#define VX_TMA_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_256B      // 256 B against 128 B: 2.618 -> 2.615 ms/step at 256^3 (profiles/r2_ablation_cache_hints.log)
#endif
// ---- tensor maps of the lattice arrays for k_lattice_tma (u64 elements; members stacked along z) --------------
static int build_tensor_maps(vx_sim* s)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return fail(s, VX_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
        encode = (EncodeFn)fn;
    }
    const cuuint64_t nx = s->nx, ny = s->ny, NZ = (cuuint64_t)s->nz * s->lat_members, N = s->N;
    std::vector<CUtensorMap> maps(2 * TM_COUNT);
    // box of bx x by x bz voxels (x bp parts) of an array with per_voxel_u64 eight-byte words per voxel and `parts` sub-arrays
    auto make = [&](CUtensorMap* m, void* base, int per_voxel_u64, int parts, cuuint32_t bx, cuuint32_t by, cuuint32_t bz, cuuint32_t bp) -> bool {
        const cuuint64_t rec = (cuuint64_t)per_voxel_u64 * 8;                    // bytes per voxel
        cuuint64_t dims[4] = {nx * per_voxel_u64, ny, NZ, (cuuint64_t)parts};
        cuuint64_t strides[3] = {nx * rec, nx * ny * rec, N * rec};
        cuuint32_t box[4] = {bx * per_voxel_u64, by, bz, bp}, es[4] = {1, 1, 1, 1};
        const cuuint32_t rank = parts ? 4 : 3;
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      VX_TMA_L2PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool ok = true;
    for (int g = 0; g < 2 && ok; g++) {
        CUtensorMap* m = maps.data() + g * TM_COUNT;
        void* po = s->pose0[g].p;                      // [2][N] double4: part 0 = pose0, part 1 = pose1
        void* rc = s->rec[g].p;                        // [12][N] sixteen-byte parts, four per axis
        ok = make(m + TM_P_OWN, po, 4, 2, 4, 4, 2, 2) && make(m + TM_P_XF, po, 4, 2, 1, 4, 2, 2) && make(m + TM_P_YF, po, 4, 2, 4, 1, 2, 2) && make(m + TM_P_ZF, po, 4, 2, 4, 4, 1, 2) &&
             make(m + TM_M0, s->mom0[g].p, 4, 0, 4, 4, 2, 0) && make(m + TM_M1, s->mom1[g].p, 2, 0, 4, 4, 2, 0) &&
             make(m + TM_REC_OWN, rc, 2, VX_REC_PARTS, 4, 4, 2, VX_REC_PARTS) && make(m + TM_REC_XF, rc, 2, VX_REC_PARTS, 1, 4, 2, 4) &&
             make(m + TM_REC_YF, rc, 2, VX_REC_PARTS, 4, 1, 2, 4) && make(m + TM_REC_ZF, rc, 2, VX_REC_PARTS, 4, 4, 1, 4);
    }
    if (!ok) return fail(s, VX_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    CK(s->tmaps.alloc(maps.size() * sizeof(CUtensorMap)));
    CK(cudaMemcpy(s->tmaps.p, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    return VX_OK;
}

// > 48 KB of dynamic shared memory needs a one-time opt-in per function and device
static void lattice_opt_in(vx_sim* s)
{
    if (s->wb_opted_in) return;
    cudaFuncSetAttribute(k_lattice_warp<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_WB_SMEM);
    cudaFuncSetAttribute(k_lattice_warp<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_WB_SMEM);
    cudaFuncSetAttribute(k_lattice_warp<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_WB_SMEM);
    cudaFuncSetAttribute(k_lattice_warp<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_WB_SMEM);
    cudaFuncSetAttribute(k_lattice_tma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM);
    cudaFuncSetAttribute(k_lattice_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM + VX_TMA_TABLE_BYTES);
    cudaFuncSetAttribute(k_lattice_tma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM);
    cudaFuncSetAttribute(k_lattice_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM + VX_TMA_TABLE_BYTES);
    cudaFuncSetAttribute(k_lattice_tma<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM + VX_TMA_TABLE_BYTES);
    cudaFuncSetAttribute(k_lattice_tma<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM + VX_TMA_TABLE_BYTES);
    cudaFuncSetAttribute(k_lattice_tma<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM);
    cudaFuncSetAttribute(k_lattice_tma<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM + VX_TMA_TABLE_BYTES);
    cudaFuncSetAttribute(k_lattice_tma<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM);
    cudaFuncSetAttribute(k_lattice_tma<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VX_TMA_SMEM + VX_TMA_TABLE_BYTES);
    s->wb_opted_in = true;
}

// default fused kernel over the brick-group layers [gz_off, gz_off + ngz) (ngz < 0: all)
static void launch_lattice_warp(vx_sim* s, int g, int first_of_call, int gz_off, int ngz, int book)
{
    const int planes = s->ghost_skip ? s->z_hi - s->z_lo : s->nz;    // z-slabs: the ghost planes at the ends are data, not bricks
    const int bx = (s->nx + VX_WB_X - 1) / VX_WB_X, by = (s->ny + VX_WB_Y - 1) / VX_WB_Y, bz = (planes + VX_WB_Z - 1) / VX_WB_Z;
    const int gx = (bx + 1) / 2, gy = (by + 1) / 2, gz = (bz + 1) / 2;
    // 2x2x2 groups of bricks unless their padding would waste more than a tenth of the warps (small boxes)
    const bool grouped = ngz >= 0 || (double)gx * gy * gz * 8 <= 1.1 * (double)bx * by * bz;
    const int nbx = grouped ? gx : bx, nby = grouped ? gy : by, nbz = ngz >= 0 ? ngz : (grouped ? gz : bz);
    const long long bricks = (long long)nbx * nby * nbz * (grouped ? 8 : 1) * s->lat_members;
    const long long grid = (bricks + VX_WB_WARPS - 1) / VX_WB_WARPS;
    // staging: TMA bulk tensor copies (7, and what 0 picks on large lattices) or per-lane cp.async (5, and what 0 picks for
    // ensembles of small boxes, where whole-box copies fetch too much padding: 1.15 against 1.19 ms on 4096 robots of 10^3)
    const bool listed = s->n_groups > 0 && ngz < 0;                 // sparse body: only its occupied brick groups
    const bool want_tma = s->path == 7 || (s->path != 5 && grouped) || s->any_poisson || listed || s->ghost_skip;     // Poisson coupling and group lists live in the TMA-staged kernel only
    // (tensor maps are built outside stream capture: ensure_lattice_graph launches nothing before they exist)
    const bool tma = want_tma && (s->tmaps.p || (!s->capturing && build_tensor_maps(s) == VX_OK));
    lattice_opt_in(s);
    if (grid > 0) {
        LatFrame f = s->lat_frame(g);
        if ((first_of_call || s->amb_each_step) && s->amb_pending) { f.amb_set = 1; f.amb = s->amb_value; }
        const int fl = s->floor_on ? 1 : 0;
        const dim3 bl(32 * VX_WB_WARPS);
        const unsigned char* tm = s->tmaps.p;
        if (tma) {
            // grouped: one CTA per 2x2x2 group of bricks on a 3-D grid (no index divisions in the kernel); a grid too tall for
            // blockIdx.y/z falls back to the 1-D brick enumeration, which covers the same bricks
            const long long zdim = (long long)nbz * s->lat_members;
            const bool g3 = (grouped || listed) && VX_WB_WARPS == 8 && nby <= 65535 && zdim <= 65535;
            if (!listed) f.groups = nullptr;
            const dim3 gr = listed ? dim3((unsigned)s->n_groups) : g3 ? dim3((unsigned)nbx, (unsigned)nby, (unsigned)zdim) : dim3((unsigned)(((long long)bx * by * (ngz >= 0 ? 2 * ngz : bz) * s->lat_members + VX_WB_WARPS - 1) / VX_WB_WARPS));
            const int kx = g3 ? nbx : bx, ky = g3 ? nby : by, kz = g3 ? nbz : (ngz >= 0 ? 2 * ngz : bz), koff = g3 ? gz_off : 2 * gz_off, gr_ = g3 ? 1 : 0;
            // multi-material models: tables staged in the CTA's shared memory when they fit behind the warp windows
            const size_t tab = s->mats.size() * sizeof(DevVoxMat) + s->lmats.size() * sizeof(DevLinkMat) + s->mats.size() * 4 + s->mats.size() * s->mats.size() * 2;
            const int stage = (!s->uni && tab <= VX_TMA_TABLE_BYTES - 16) ? 1 : 0;
            const size_t tma_smem = VX_TMA_SMEM + (stage ? VX_TMA_TABLE_BYTES : 0);
            if (s->ghost_skip) {                                  // z-slab: bricks over the owned planes only
                if (s->push_in_kernel) {
                    if (s->uni) k_lattice_tma<true, true, false, true><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
                    else k_lattice_tma<false, true, false, true><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
                } else {
                    if (s->uni) k_lattice_tma<true, false, false, true><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
                    else k_lattice_tma<false, false, false, true><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
                }
            }
            else if (s->any_poisson && s->push_in_kernel) k_lattice_tma<false, true, true><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
            else if (s->any_poisson) k_lattice_tma<false, false, true><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
            else if (s->push_in_kernel) {     // boundary part of vx_slab_step: new poses also go to the neighbours' ghost layers
                if (s->uni) k_lattice_tma<true, true><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
                else k_lattice_tma<false, true><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
            } else {
                if (s->uni) k_lattice_tma<true, false><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
                else k_lattice_tma<false, false><<<gr, bl, tma_smem, s->stream>>>(f, tm, g, first_of_call, fl, kx, ky, kz, koff, book, gr_, stage);
            }
        } else if (s->any_poisson || s->ghost_skip) {
            s->err = "this model needs the TMA-staged kernel (tensor maps could not be built)"; s->launch_failed = true;
        } else {
            const dim3 gr((unsigned)grid);
            const int gr_ = grouped ? 1 : 0;
            if (s->push_in_kernel) {
                if (s->uni) k_lattice_warp<true, true><<<gr, bl, VX_WB_SMEM, s->stream>>>(f, g, first_of_call, fl, nbx, nby, nbz, gz_off, book, gr_);
                else k_lattice_warp<false, true><<<gr, bl, VX_WB_SMEM, s->stream>>>(f, g, first_of_call, fl, nbx, nby, nbz, gz_off, book, gr_);
            } else {
                if (s->uni) k_lattice_warp<true, false><<<gr, bl, VX_WB_SMEM, s->stream>>>(f, g, first_of_call, fl, nbx, nby, nbz, gz_off, book, gr_);
                else k_lattice_warp<false, false><<<gr, bl, VX_WB_SMEM, s->stream>>>(f, g, first_of_call, fl, nbx, nby, nbz, gz_off, book, gr_);
            }
        }
        s->launches++;
    }
}

// one doTimeStep on the fused path: contact forces from the OLD state (generation g), then the fused kernel
static int launch_lattice(vx_sim* s, int g, int first_of_call, bool capturing = false)
{
    if (s->collisions) {
        s->gen_view = g;
        int rc = enqueue_collision_step(s, capturing);
        s->gen_view = -1;
        if (rc != VX_OK) return rc;
    }
    launch_lattice_warp(s, g, first_of_call, 0, -1, 1);
    return VX_OK;
}

static int ensure_lattice_graph(vx_sim* s, int g0)
{
    if (s->lgraph[g0]) return VX_OK;
    if (!s->tmaps.p) build_tensor_maps(s);            // allocates and copies: not allowed while a capture is open (a failure falls back to cp.async staging)
    lattice_opt_in(s);
    cudaGraph_t g = nullptr;
    int64_t before = s->launches;
    cudaStream_t user = s->stream; s->stream = s->own_stream;       // see ensure_graph
    cudaError_t e0 = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal);
    if (e0 != cudaSuccess) { s->stream = user; return cuda_fail(s, e0, "cudaStreamBeginCapture"); }
    int rc = VX_OK;
    s->capturing = true;
    for (int k = 0; k < GRAPH_STEPS && rc == VX_OK; k++) rc = launch_lattice(s, (g0 + k) & 1, 0, true);
    s->capturing = false;
    cudaError_t e = cudaStreamEndCapture(s->stream, &g);
    s->stream = user;
    s->lgraph_kernels = (int)(s->launches - before);
    s->launches = before;
    if (rc != VX_OK) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
    if (e != cudaSuccess) return cuda_fail(s, e, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(&s->lgraph[g0], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { s->lgraph[g0] = nullptr; return cuda_fail(s, e, "cudaGraphInstantiate"); }
    return VX_OK;
}

// after the queued steps: read back the step parameters and re-align generations
static int finish_lattice_call(vx_sim* s, int g_start, int launched, int* diverged_step)
{
    k_lattice_finish<<<1, 1, 0, s->stream>>>(s->params.p, (g_start + launched - 1) & 1); s->launches++;
    CK(cudaGetLastError());
    if (s->launch_failed) { s->launch_failed = false; cudaStreamSynchronize(s->stream); return VX_ERR_UNSUPPORTED; }
    CK(cudaMemcpyAsync(s->params_host, s->params.p, sizeof(DevParams), cudaMemcpyDeviceToHost, s->stream));
    if (s->collisions && s->col_tables) CK(cudaMemcpyAsync(s->counters_host, s->c_counters.p, CC_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    const DevParams& p = *s->params_host;
    s->time_host = p.time;
    const bool amb_used = s->amb_pending;            // the first step of this call applied a pending ambient temperature
    s->amb_pending = false;
    const int col_rc = collision_call_end(s);
    if (col_rc != VX_OK) return col_rc;
    if (!p.div_latched) {
        s->gen = (g_start + launched) & 1;
        s->have_prev = true;
        s->last_amb = amb_used && launched == 1; s->last_amb_value = s->amb_value;
        s->last_prev_dt = launched > 1 ? (s->call_per_step_dt ? p.last_prev : p.dt) : s->prev_dt_host;
        s->prev_dt_host = p.prev_dt;
        return VX_OK;
    }
    // step number steps_done (0-based) diverged: it read generation g, wrote g^1.  Links keep the
    // advanced state, voxels are not advanced (src/Voxelyze.cpp:263-269): copy them over.
    const int g = (g_start + p.steps_done) & 1;
    k_lattice_copy_voxels<<<blocks_for(s->N), TPB, 0, s->stream>>>(s->N, s->pose0[g].p, s->pose1[g].p, s->mom0[g].p, s->mom1[g].p,
                                                                  s->pose0[g ^ 1].p, s->pose1[g ^ 1].p, s->mom0[g ^ 1].p, s->mom1[g ^ 1].p, 1);
    s->launches++;
    s->gen = g ^ 1;
    if (amb_used && p.steps_done == 0) {             // the voxels that were not advanced still got their new temperature
        k_fill_temp<<<blocks_for(s->N), TPB, 0, s->stream>>>(s->frame(), s->amb_value, nullptr, nullptr); s->launches++;
    }
    CK(cudaStreamSynchronize(s->stream));
    s->last_amb = amb_used && p.steps_done == 0; s->last_amb_value = s->amb_value;
    s->have_prev = true;
    s->last_prev_dt = p.steps_done > 0 ? (s->call_per_step_dt ? p.last_prev : p.dt) : s->prev_dt_host;
    s->prev_dt_host = p.prev_dt;
    if (diverged_step) *diverged_step = p.steps_done;
    return VX_DIVERGED;
}

// z-slabs with skipped ghost planes: their flag words into the generation the call writes first
static void ghost_words_begin(vx_sim* s)
{
    if (!s->ghost_skip) return;
    const int plane = s->nx * s->ny, n_lo = s->z_lo * plane, n_hi = (s->nz - s->z_hi) * plane;
    if (n_lo + n_hi == 0) return;
    k_lattice_ghost_words<<<blocks_for(n_lo + n_hi), TPB, 0, s->stream>>>(s->pose1[s->gen].p, s->pose1[s->gen ^ 1].p, n_lo, s->z_hi * plane, n_hi);
    s->launches++;
}

static int lattice_step(vx_sim* s, float dt, int n_steps, int* diverged_step)
{
    const bool per_step_dt = dt < 0 && s->any_poisson;      // with Poisson coupling the stable step depends on the state (src/VX_Link.cpp:259-267)
    if (dt < 0 && !per_step_dt) {                   // nu = 0: the recommended step is a constant of the model
        int rc = vx_recommended_dt(s, &dt);
        if (rc != VX_OK) return rc;
        if (dt <= 0) return VX_OK;
    }
    k_begin<<<1, 1, 0, s->stream>>>(s->params.p, dt, per_step_dt ? 0 : 1); s->launches++;
    collision_call_begin(s);
    ghost_words_begin(s);
    const int g0 = s->gen;
    int done = 0;
    s->call_per_step_dt = per_step_dt;
    if (per_step_dt) {
        for (; done < n_steps; done++) {
            launch_recommended_dt(s, (g0 + done) & 1);
            k_dt_from_freq<<<1, 1, 0, s->stream>>>(s->freq2.p, s->params.p, done > 0 ? 1 : 0, ((g0 + done) & 1) ^ 1); s->launches++;
            int rc = launch_lattice(s, (g0 + done) & 1, 1); if (rc != VX_OK) return rc;       // every step damps with the dt of the step before it (DevParams::prev_dt)
        }
        return finish_lattice_call(s, g0, n_steps, diverged_step);
    }
    { int rc = launch_lattice(s, g0, 1); if (rc != VX_OK) return rc; }
    done++;
    while (n_steps - done >= GRAPH_STEPS) {
        const int g = (g0 + done) & 1;
        int rc = ensure_lattice_graph(s, g);                 // both generations at once: a later call may start on the other one
        if (rc == VX_OK) rc = ensure_lattice_graph(s, g ^ 1);
        if (rc != VX_OK) return rc;
        CK(cudaGraphLaunch(s->lgraph[g], s->stream));
        s->launches += s->lgraph_kernels; done += GRAPH_STEPS;
    }
    for (; done < n_steps; done++) { int rc = launch_lattice(s, (g0 + done) & 1, 0); if (rc != VX_OK) return rc; }
    return finish_lattice_call(s, g0, n_steps, diverged_step);
}

#include "vx_slab_kernels.cuh"

// brick-group layers that a halo exchange touches: those holding an all-ghost plane or a plane next to one
static void find_boundary_layers(vx_sim* s)
{
    s->zb_layers.clear(); s->ghost_skip = false; s->z_lo = 0; s->z_hi = s->nz;
    if (!s->lattice || s->n_members != 1 || s->vflags.empty()) return;
    const size_t plane = (size_t)s->nx * s->ny;
    std::vector<char> ghost(s->nz, 0);
    for (int z = 0; z < s->nz; z++) {
        bool all = true;
        for (size_t k = 0; k < plane && all; k++) all = (s->vflags[s->v_i2e[(size_t)z * plane + k]] & VX_VF_GHOST) != 0;
        ghost[z] = all;
    }
    const int per = 2 * VX_WB_Z;
    // ghost planes only at the two ends, single-material-or-not but no Poisson coupling, TMA staging allowed: skip them
    s->z_lo = ghost[0] ? 1 : 0; s->z_hi = s->nz - (s->nz > 1 && ghost[s->nz - 1] ? 1 : 0);
    bool inner = false;
    for (int z = s->z_lo; z < s->z_hi; z++) inner = inner || ghost[z];
    s->ghost_skip = (s->z_lo > 0 || s->z_hi < s->nz) && !inner && s->z_hi > s->z_lo && !s->any_poisson && !s->collisions && s->path != 5 && !getenv("VX_NO_GSKIP");
    if (s->ghost_skip) {
        const int last = (s->z_hi - 1 - s->z_lo) / per;
        if (s->z_lo > 0) s->zb_layers.push_back(0);
        if (s->z_hi < s->nz && (s->zb_layers.empty() || s->zb_layers.back() != last)) s->zb_layers.push_back(last);
        return;
    }
    for (int z = 0; z < s->nz; z++) {
        bool b = ghost[z] || (z > 0 && ghost[z - 1]) || (z + 1 < s->nz && ghost[z + 1]);
        if (b && (s->zb_layers.empty() || s->zb_layers.back() != z / per)) s->zb_layers.push_back(z / per);
    }
}

extern "C" {

int vx_step_begin(vx_sim* s, float dt)
{
    if (!s) return VX_ERR_ARG;
    if (!s->lattice) return fail(s, VX_ERR_UNSUPPORTED, "asynchronous stepping needs the fused lattice path");
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_step_begin: a call is already open");
    if (dt <= 0) return fail(s, VX_ERR_ARG, "vx_step_begin needs an explicit dt");
    CK(cudaSetDevice(s->device));
    k_begin<<<1, 1, 0, s->stream>>>(s->params.p, dt, 1); s->launches++;
    ghost_words_begin(s);
    s->call_active = true; s->call_half = false; s->call_g0 = s->gen; s->call_done = 0; s->call_per_step_dt = false;
    return VX_OK;
}

int vx_step_enqueue(vx_sim* s, int part)
{
    if (!s || !s->call_active || part < 0 || part > 2) return VX_ERR_ARG;
    CK(cudaSetDevice(s->device));
    const int ngz = ((s->ghost_skip ? s->z_hi - s->z_lo : s->nz) + 2 * VX_WB_Z - 1) / (2 * VX_WB_Z);
    const bool split = !s->zb_layers.empty() && (int)s->zb_layers.size() < ngz;
    if (part == VX_PART_Z_INTERIOR) {
        if (!s->call_half) return fail(s, VX_ERR_ARG, "vx_step_enqueue: interior part without its boundary part");
        s->call_half = false;
        if (!split) return VX_OK;                    // the boundary part was the whole step
        const int g = (s->call_g0 + s->call_done - 1) & 1, first = s->call_done == 1;
        int from = 0;                                // the complement of zb_layers, as contiguous ranges
        for (size_t k = 0; k <= s->zb_layers.size(); k++) {
            const int to = k < s->zb_layers.size() ? s->zb_layers[k] : ngz;
            if (to > from) launch_lattice_warp(s, g, first, from, to - from, 0);
            from = to + 1;
        }
        CK(cudaGetLastError());
        return VX_OK;
    }
    if (s->call_half) return fail(s, VX_ERR_ARG, "vx_step_enqueue: the previous step still lacks its interior part");
    const int g = (s->call_g0 + s->call_done) & 1, first = s->call_done == 0;
    if (part == VX_PART_ALL || !split) launch_lattice_warp(s, g, first, 0, -1, 1);
    else for (size_t k = 0; k < s->zb_layers.size(); k++) launch_lattice_warp(s, g, first, s->zb_layers[k], 1, k == 0);
    s->call_done++;
    s->call_half = part == VX_PART_Z_BOUNDARY;
    CK(cudaGetLastError());
    return VX_OK;
}

int vx_step_end(vx_sim* s, int* diverged_step)
{
    if (!s || !s->call_active) return VX_ERR_ARG;
    if (s->call_half) return fail(s, VX_ERR_ARG, "vx_step_end: the last step lacks its interior part");
    CK(cudaSetDevice(s->device));
    s->call_active = false;
    if (diverged_step) *diverged_step = -1;
    if (s->call_done == 0) return VX_OK;
    return finish_lattice_call(s, s->call_g0, s->call_done, diverged_step);
}

#include "vx_slab.inl"
#include "vx_state_io.inl"

int vx_abi_version(void) { return VX_ABI_VERSION; }
const char* vx_backend(void) { return "cuda-sm100a"; }

int vx_create(double voxel_size, int device, vx_sim** out)
{
    if (!out) return VX_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return VX_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return VX_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return VX_ERR_NO_DEVICE;   // sm_100a code only
    vx_sim* s = new vx_sim;
    s->device = device; s->vox_size = voxel_size; s->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return VX_ERR_CUDA; }
    s->stream = s->own_stream;
    if (s->params.alloc(1) != cudaSuccess || s->freq2.alloc(1) != cudaSuccess ||
        cudaMallocHost((void**)&s->params_host, sizeof(DevParams)) != cudaSuccess ||
        cudaMallocHost((void**)&s->freq_host, sizeof(unsigned int)) != cudaSuccess) { vx_destroy(s); return VX_ERR_ALLOC; }
    DevParams p{}; cudaMemcpy(s->params.p, &p, sizeof(p), cudaMemcpyHostToDevice);
    *out = s;
    return VX_OK;
}

void vx_destroy(vx_sim* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    vx_peer_detach(s);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    if (s->ev_boundary) cudaEventDestroy(s->ev_boundary);
    if (s->ev_comm) cudaEventDestroy(s->ev_comm);
    s->peer_flags.release(); s->tmaps.release(); s->group_list.release();
    s->drop_graph();
    for (int g = 0; g < 2; g++) { s->release_pose(g); s->mom0[g].release(); s->mom1[g].release(); s->rec[g].release(); s->ps[g].release(); }
    s->pair_lmat.release(); s->link_owner.release(); s->link_axis_dev.release();
    s->c_surf_vox.release(); s->c_surf_orig.release(); s->c_surf_member.release(); s->c_slot.release(); s->c_surf_ijk.release(); s->c_nearby.release();
    s->c_last_watch.release(); s->c_cell_count.release(); s->c_cell_start.release(); s->c_sorted.release(); s->c_cell.release();
    if (s->aux_stream) cudaStreamDestroy(s->aux_stream); s->c_pairs.release(); s->c_pair_kc.release();
    s->c_pair_force.release(); s->c_counters.release(); s->c_deg.release(); s->c_ref_start.release(); s->c_ref_fill.release(); s->c_refs.release();
    if (s->counters_host) cudaFreeHost(s->counters_host);
    s->si_minmax.release(); s->si_sum.release(); s->si_nominal.release(); s->si_consts.release(); s->si_buf.release();
    s->si_vlinks.release(); s->si_ratio.release(); s->si_en.release(); s->si_skip.release(); s->mesh.release();
    s->ext_idx.release(); s->ext_vox_dev.release(); s->vox_e2i_dev.release(); s->link_e2i_dev.release(); s->member_dev.release();
    s->pstrain.release(); s->slots.release(); s->slot_strain.release();
    s->lends.release(); s->lmeta.release(); s->lstA.release(); s->lstB.release(); s->lstC.release(); s->lstrain.release();
    s->vmat_dev.release(); s->lmat_dev.release(); s->curve_e.release(); s->curve_s.release(); s->ext_dev.release();
    s->params.release(); s->freq2.release(); s->member_t.release(); s->staging.release();
    if (s->params_host) cudaFreeHost(s->params_host);
    if (s->freq_host) cudaFreeHost(s->freq_host);
    if (s->probe_host) cudaFreeHost(s->probe_host);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

const char* vx_last_error(const vx_sim* s) { return s ? s->err.c_str() : "null handle"; }

#include "vx_model.inl"

int vx_set_gravity(vx_sim* s, float g) { if (!s) return VX_ERR_ARG; s->grav = g; return s->mats.empty() ? VX_OK : upload_tables(s); }
int vx_enable_floor(vx_sim* s, int e) { if (!s) return VX_ERR_ARG; s->floor_on = e != 0; s->drop_graph(); return VX_OK; }
// the device layout depends on whether collisions are watched (collisions run on the general layout),
// so switching them on re-lays out a simulation that has not been stepped yet
static int relayout_fresh(vx_sim* s)
{
    const int nu = s->N_user;                                      // only the caller's voxels; fill cells are derived again (or dropped)
    std::vector<int32_t> ijk(s->ijk.begin(), s->ijk.begin() + 3 * (size_t)nu), member(s->member.begin(), s->member.begin() + nu);
    std::vector<uint16_t> mat(s->vmat_id.begin(), s->vmat_id.begin() + nu); std::vector<uint32_t> fl;
    if (!s->vflags.empty()) fl.assign(s->vflags.begin(), s->vflags.begin() + nu);
    bool any_flag = false;
    for (uint32_t f : fl) if (f) any_flag = true;
    if (!any_flag) fl.clear();
    s->relayout = true;
    int rc = vx_set_voxels(s, nu, ijk.data(), mat.data(), s->n_members > 1 ? member.data() : nullptr, fl.empty() ? nullptr : fl.data());
    s->relayout = false;
    if (rc != VX_OK) return rc;
    if (!s->ext_raw_vox.empty()) {
        int n = (int)s->ext_raw_vox.size();
        rc = vx_set_externals(s, n, s->ext_raw_vox.data(), s->ext_raw_dof.data(), s->ext_raw_f.empty() ? nullptr : s->ext_raw_f.data(),
                              s->ext_raw_m.empty() ? nullptr : s->ext_raw_m.data(), s->ext_raw_t.empty() ? nullptr : s->ext_raw_t.data(),
                              s->ext_raw_r.empty() ? nullptr : s->ext_raw_r.data());
    }
    return rc;
}

// change of layout in the middle of a run (collisions switched on: lattice -> general): every voxel and link keeps its state
static int relayout_keep_state(vx_sim* s)
{
    const int N = s->N_user, L = s->L;
    std::vector<double> pos(3 * (size_t)N), ori(4 * (size_t)N), lin(3 * (size_t)N), ang(3 * (size_t)N);
    std::vector<float> temp(N); std::vector<uint32_t> vfl(N); std::vector<vx_link_state> ls(L);
    int rc = vx_download(s, VX_F_POS, 0, N, pos.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_ORIENT, 0, N, ori.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_LINMOM, 0, N, lin.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_ANGMOM, 0, N, ang.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_TEMP, 0, N, temp.data());
    if (rc == VX_OK) rc = vx_download(s, VX_F_VOXFLAGS, 0, N, vfl.data());
    if (rc == VX_OK && L) rc = vx_download_link_state(s, 0, L, ls.data());
    if (rc != VX_OK) return rc;
    const float time = s->time_host, prev_dt = s->prev_dt_host, ambient = s->ambient;
    rc = relayout_fresh(s);
    if (rc != VX_OK) return rc;
    if (s->L != L || s->N_user != N) return fail(s, VX_ERR_CUDA, "relayout changed the model");
    rc = vx_upload(s, VX_F_POS, 0, N, pos.data());
    if (rc == VX_OK) rc = vx_upload(s, VX_F_ORIENT, 0, N, ori.data());
    if (rc == VX_OK) rc = vx_upload(s, VX_F_LINMOM, 0, N, lin.data());
    if (rc == VX_OK) rc = vx_upload(s, VX_F_ANGMOM, 0, N, ang.data());
    if (rc == VX_OK) rc = vx_upload(s, VX_F_TEMP, 0, N, temp.data());
    if (rc == VX_OK) rc = vx_upload(s, VX_F_VOXFLAGS, 0, N, vfl.data());
    if (rc == VX_OK && L) rc = vx_upload_link_state(s, 0, L, ls.data());
    if (rc != VX_OK) return rc;
    DevParams p{}; p.col_stale = 1; p.time = time; p.prev_dt = prev_dt;
    CK(cudaMemcpy(s->params.p, &p, sizeof(p), cudaMemcpyHostToDevice));
    s->time_host = time; s->prev_dt_host = prev_dt; s->ambient = ambient;
    return VX_OK;
}

int vx_enable_collisions(vx_sim* s, int e)                         // src/Voxelyze.cpp:612-622
{
    if (!s) return VX_ERR_ARG;
    if (s->collisions == (e != 0)) return VX_OK;
    { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    s->collisions = e != 0;
    s->col_stale_host = true;
    s->drop_graph();
    if (!s->collisions) { s->n_pairs = 0; return VX_OK; }           // clearCollisions()
    if (s->N == 0 || s->col_tables) return VX_OK;
    if (s->time_host != 0.f || s->have_prev) return relayout_keep_state(s);   // mid-run: the layout changes, the state does not
    return relayout_fresh(s);
}
int vx_set_collision_envelope(vx_sim* s, float r) { if (!s) return VX_ERR_ARG; s->envelope = r; return VX_OK; }

int vx_set_temperature_all(vx_sim* s, float t)
{
    if (!s) return VX_ERR_ARG;
    s->ambient = t;
    if (s->N == 0) return VX_OK;
    s->amb_pending = true; s->amb_value = t;
    if (s->lattice && !s->collisions && !s->call_active) return VX_OK;     // the next fused step applies it (LatFrame::amb)
    return flush_ambient(s);
}
int vx_set_temperature_members(vx_sim* s, int n, const float* t)
{
    if (!s || !t || n != s->n_members) return VX_ERR_ARG;
    if (s->N == 0) return VX_OK;
    s->amb_pending = false;                                          // every voxel is overwritten
    CK(cudaSetDevice(s->device));
    CK(s->member_t.alloc(n));
    CK(cudaMemcpyAsync(s->member_t.p, t, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    k_fill_temp<<<blocks_for(s->N), TPB, 0, s->stream>>>(s->frame(), 0.f, s->member_t.p, s->member_dev.p); s->launches++;
    CK(cudaGetLastError());
    return VX_OK;
}
int vx_set_temperature(vx_sim* s, int n, const float* t)
{
    if (!s || !t || n != s->N_user) return VX_ERR_ARG;
    return vx_upload(s, VX_F_TEMP, 0, n, t);
}

int vx_step(vx_sim* s, float dt, int n_steps, int* diverged_step)
{
    if (!s || n_steps < 0) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_step inside vx_step_begin .. vx_step_end");
    if (n_steps == 0 || dt == 0 || s->N == 0) return VX_OK;       // dt == 0: src/Voxelyze.cpp:253
    NvtxRange nvtx("vx_step");
    CK(cudaSetDevice(s->device));
    if (s->lattice) return lattice_step(s, dt, n_steps, diverged_step);
    Frame f = s->frame();
    const bool per_step_dt = dt < 0 && s->any_poisson;
    k_begin<<<1, 1, 0, s->stream>>>(s->params.p, dt, dt > 0 ? 1 : 0); s->launches++;
    if (dt < 0 && !per_step_dt) {                                  // constant recommended dt
        launch_recommended_dt(s);
        k_dt_from_freq<<<1, 1, 0, s->stream>>>(s->freq2.p, s->params.p, 0); s->launches++;
    }
    int left = n_steps;
    collision_call_begin(s);
    if (s->small && !s->collisions && !per_step_dt) {              // the whole call in one launch of one thread-block cluster
        int rc = launch_small_steps(s, f, left);
        if (rc != VX_OK) return rc;
        left = 0;
    }
    if (!per_step_dt && left >= GRAPH_STEPS) {
        int rc = ensure_graph(s);
        if (rc != VX_OK) return rc;
        while (left >= GRAPH_STEPS) { CK(cudaGraphLaunch(s->graph, s->stream)); s->launches += s->graph_kernels; left -= GRAPH_STEPS; }
    }
    for (; left > 0; left--) { int rc = launch_step(s, f, per_step_dt); if (rc != VX_OK) return rc; }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(s->params_host, s->params.p, sizeof(DevParams), cudaMemcpyDeviceToHost, s->stream));
    if (s->collisions && s->col_tables) CK(cudaMemcpyAsync(s->counters_host, s->c_counters.p, CC_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->time_host = s->params_host->time;
    s->prev_dt_host = s->params_host->prev_dt;
    { int rc = collision_call_end(s); if (rc != VX_OK) return rc; }
    if (s->params_host->div_latched) {
        if (diverged_step) *diverged_step = s->params_host->steps_done;
        return VX_DIVERGED;
    }
    return VX_OK;
}

int vx_step_ambient(vx_sim* s, float dt, int n_steps, const float* ambient, int* diverged_step)
{
    if (!s || n_steps < 0 || (n_steps && !ambient)) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_step_ambient inside vx_step_begin .. vx_step_end");
    if (diverged_step) *diverged_step = -1;
    if (n_steps == 0 || dt == 0 || s->N == 0) return VX_OK;
    if (!s->lattice || s->collisions || (dt < 0 && s->any_poisson)) {         // other layouts: the calls it stands for, one by one
        for (int k = 0; k < n_steps; k++) {
            int rc = vx_set_temperature_all(s, ambient[k]); if (rc != VX_OK) return rc;
            int d = -1;
            rc = vx_step(s, dt, 1, &d);
            if (rc == VX_DIVERGED && diverged_step) *diverged_step = k;
            if (rc != VX_OK) return rc;
        }
        return VX_OK;
    }
    NvtxRange nvtx("vx_step_ambient");
    CK(cudaSetDevice(s->device));
    // fused lattice path: the temperature is a launch parameter of the step kernel (LatFrame::amb), so a program of n
    // temperatures is n launches queued back to back -- no pass over the voxels, no host synchronisation until the end
    if (dt < 0) { int rc = vx_recommended_dt(s, &dt); if (rc != VX_OK) return rc; if (dt <= 0) return VX_OK; }
    k_begin<<<1, 1, 0, s->stream>>>(s->params.p, dt, 1); s->launches++;
    collision_call_begin(s);
    ghost_words_begin(s);
    const int g0 = s->gen;
    s->call_per_step_dt = false; s->amb_each_step = true;
    int rc = VX_OK;
    for (int k = 0; k < n_steps && rc == VX_OK; k++) {
        s->amb_pending = true; s->amb_value = ambient[k];
        rc = launch_lattice(s, (g0 + k) & 1, k == 0 ? 1 : 0);
    }
    s->amb_each_step = false; s->amb_pending = false;      // (finish_lattice_call's own ambient bookkeeping is for single pending values)
    if (rc != VX_OK) { cudaStreamSynchronize(s->stream); return rc; }
    int d = -1;
    rc = finish_lattice_call(s, g0, n_steps, &d);
    if (rc != VX_OK && rc != VX_DIVERGED) return rc;
    const int last = rc == VX_DIVERGED ? d : n_steps - 1;         // the last step that ran (a diverging step updates its links)
    s->ambient = ambient[last]; s->amb_value = ambient[last];
    if (rc == VX_DIVERGED) {               // its voxels were not advanced, but they had been given their new temperature (src/Voxelyze.cpp:585-594 precedes doTimeStep)
        k_fill_temp<<<blocks_for(s->N), TPB, 0, s->stream>>>(s->frame(), ambient[last], nullptr, nullptr); s->launches++;
        CK(cudaStreamSynchronize(s->stream));
        if (diverged_step) *diverged_step = d;
    }
    s->last_amb = true; s->last_amb_value = ambient[last];       // the inputs of that step (generation gen^1) still hold the temperatures before it
    return rc;
}

int vx_step_profile(vx_sim* s, float dt, int n_steps, float* ms, int* launches)
{
    if (!s || n_steps < 0 || !ms) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_step_profile inside vx_step_begin .. vx_step_end");
    ms[0] = ms[1] = ms[2] = ms[3] = 0.f;
    if (launches) launches[0] = launches[1] = launches[2] = 0;
    if (n_steps == 0 || dt == 0 || s->N == 0) return VX_OK;
    NvtxRange nvtx("vx_step_profile");
    CK(cudaSetDevice(s->device));
    struct Events {                                                // destroyed on every return path
        std::vector<cudaEvent_t> v;
        ~Events() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
    } evs;
    evs.v.assign((size_t)n_steps * 4, nullptr);
    std::vector<cudaEvent_t>& ev = evs.v;
    for (auto& e : ev) CK(cudaEventCreate(&e));
    int rc = VX_OK;
    if (s->lattice) {
        // one fused kernel per step: reported as the "link" group (it is the dominant kernel)
        if (dt < 0) { rc = vx_recommended_dt(s, &dt); if (rc != VX_OK || dt <= 0) return rc; }
        s->call_per_step_dt = false;
        k_begin<<<1, 1, 0, s->stream>>>(s->params.p, dt, 1); s->launches++;
        collision_call_begin(s);
        ghost_words_begin(s);
        const int g0 = s->gen;
        for (int k = 0; k < n_steps; k++) {
            CK(cudaEventRecord(ev[4 * k + 0], s->stream));
            rc = launch_lattice(s, (g0 + k) & 1, k == 0 ? 1 : 0); if (rc != VX_OK) return rc;
            CK(cudaEventRecord(ev[4 * k + 3], s->stream));
            if (launches) launches[0] += 1;
        }
        rc = finish_lattice_call(s, g0, n_steps, nullptr);
        for (int k = 0; k < n_steps; k++) { float d = 0; cudaEventElapsedTime(&d, ev[4 * k + 0], ev[4 * k + 3]); ms[0] += d; ms[3] += d; }
    } else {
        Frame f = s->frame();
        const bool per_step_dt = dt < 0 && s->any_poisson;
        k_begin<<<1, 1, 0, s->stream>>>(s->params.p, dt, dt > 0 ? 1 : 0); s->launches++;
        if (dt < 0 && !per_step_dt) { launch_recommended_dt(s); k_dt_from_freq<<<1, 1, 0, s->stream>>>(s->freq2.p, s->params.p, 0); s->launches++; }
        collision_call_begin(s);
        for (int k = 0; k < n_steps; k++) {
            int64_t l0 = s->launches;
            CK(cudaEventRecord(ev[4 * k + 0], s->stream));
            if (s->any_poisson) { k_pstrain<<<blocks_for(s->N), TPB, 0, s->stream>>>(f); s->launches++; }
            if (per_step_dt) { launch_recommended_dt(s); k_dt_from_freq<<<1, 1, 0, s->stream>>>(s->freq2.p, s->params.p, 0); s->launches++; }
            int64_t l1 = s->launches;
            CK(cudaEventRecord(ev[4 * k + 1], s->stream));
            launch_links(s, f);
            int64_t l2 = s->launches;
            CK(cudaEventRecord(ev[4 * k + 2], s->stream));
            if (s->collisions) { rc = enqueue_collision_step(s, false); if (rc != VX_OK) return rc; }
            launch_voxel(s, f);
            CK(cudaEventRecord(ev[4 * k + 3], s->stream));
            if (launches) { launches[2] += (int)(l1 - l0); launches[0] += (int)(l2 - l1); launches[1] += 1; }
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(s->params_host, s->params.p, sizeof(DevParams), cudaMemcpyDeviceToHost, s->stream));
        if (s->collisions && s->col_tables) CK(cudaMemcpyAsync(s->counters_host, s->c_counters.p, CC_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        { int rc2 = collision_call_end(s); if (rc2 != VX_OK) return rc2; }
        for (int k = 0; k < n_steps; k++) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime(&a, ev[4 * k + 0], ev[4 * k + 1]);
            cudaEventElapsedTime(&b, ev[4 * k + 1], ev[4 * k + 2]);
            cudaEventElapsedTime(&c, ev[4 * k + 2], ev[4 * k + 3]);
            cudaEventElapsedTime(&d, ev[4 * k + 0], ev[4 * k + 3]);
            ms[2] += a; ms[0] += b; ms[1] += c; ms[3] += d;
        }
        s->time_host = s->params_host->time;
        s->prev_dt_host = s->params_host->prev_dt;
        rc = s->params_host->div_latched ? VX_DIVERGED : VX_OK;
    }
    return rc;
}

int vx_prepare(vx_sim* s)
{
    if (!s) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_prepare inside vx_step_begin .. vx_step_end");
    if (s->N == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    if (s->lattice) {
        int rc = ensure_lattice_graph(s, 0);
        if (rc == VX_OK) rc = ensure_lattice_graph(s, 1);
        return rc;
    }
    return s->any_poisson ? VX_OK : ensure_graph(s);
}

int vx_recommended_dt(vx_sim* s, float* dt)
{
    if (!s || !dt) return VX_ERR_ARG;
    *dt = 0.f;
    if (s->N == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    if (s->any_poisson && !s->lattice) { k_pstrain<<<blocks_for(s->N), TPB, 0, s->stream>>>(s->frame()); s->launches++; }
    launch_recommended_dt(s);
    CK(cudaMemcpyAsync(s->freq_host, s->freq2.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    float m; memcpy(&m, s->freq_host, 4);
    *dt = (m <= 0.0f) ? 0.0f : 1.0f / (6.283185f * std::sqrt(m));
    return VX_OK;
}

int vx_reset(vx_sim* s)
{
    if (!s) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_reset inside vx_step_begin .. vx_step_end");
    return upload_initial_state(s, 0.0f);          // CVX_Voxel::reset zeroes the temperature, src/VX_Voxel.cpp:53
}
float vx_time(const vx_sim* s) { return s ? s->time_host : 0.f; }
int vx_set_clock(vx_sim* s, float time, float previous_dt)
{
    if (!s || !(time >= 0.f) || !(previous_dt >= 0.f)) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_set_clock inside vx_step_begin .. vx_step_end");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(&s->params.p->time, &time, sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(&s->params.p->prev_dt, &previous_dt, sizeof(float), cudaMemcpyHostToDevice));
    s->time_host = time; s->prev_dt_host = previous_dt; s->last_prev_dt = previous_dt;
    return VX_OK;
}

#include "vx_access.inl"

int vx_set_stream(vx_sim* s, uint64_t stream)
{
    if (!s) return VX_ERR_ARG;
    cudaStreamSynchronize(s->stream);
    s->stream = stream == VX_OWN_STREAM ? s->own_stream : (cudaStream_t)(uintptr_t)stream;
    s->drop_graph();
    return VX_OK;
}

int vx_pose_plane(vx_sim* s, int iz, uint64_t* p0, uint64_t* p1, int* count, int* rec_bytes)
{
    if (!s || s->n_members != 1) return VX_ERR_ARG;
    if (!s->call_active) { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    int64_t key = (int64_t)(iz + 32768);
    auto lo = std::lower_bound(s->sort_key.begin(), s->sort_key.end(), key);
    auto hi = std::upper_bound(s->sort_key.begin(), s->sort_key.end(), key);
    size_t first = lo - s->sort_key.begin();
    const int g = s->lattice ? s->newest_gen() : 0;  // lattice mode ping-pongs generations; inside an asynchronous call the
                                                     // newest one is the output of the last enqueued boundary part
    if (p0) *p0 = (uint64_t)(uintptr_t)(s->pose0[g].p + first);
    if (p1) *p1 = (uint64_t)(uintptr_t)(s->pose1[g].p + first);
    if (count) *count = (int)(hi - lo);
    if (rec_bytes) *rec_bytes = (int)sizeof(double4);
    return VX_OK;
}

__global__ void k_halo_import(double4* pose0, double4* pose1, const double4* src0, const double4* src1, int count)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    double4 a = src0[k], b = src1[k];
    double keep = pose1[k].w;
    pose0[k] = a;
    pose1[k] = make_double4(b.x, b.y, b.z, meta_pack(meta_temp(b.w), meta_hi(keep)));
}

int vx_halo_import(vx_sim* s, int iz, uint64_t src0, uint64_t src1, int count)
{
    return vx_halo_import_on(s, iz, src0, src1, count, s ? (uint64_t)(uintptr_t)s->stream : 0);
}

int vx_halo_import_on(vx_sim* s, int iz, uint64_t src0, uint64_t src1, int count, uint64_t stream)
{
    if (!s || !src0 || !src1) return VX_ERR_ARG;
    uint64_t p0, p1; int n, rb;
    int rc = vx_pose_plane(s, iz, &p0, &p1, &n, &rb);
    if (rc != VX_OK) return rc;
    if (n != count) return fail(s, VX_ERR_ARG, "halo layer size mismatch");
    if (n == 0) return VX_OK;
    CK(cudaSetDevice(s->device));
    k_halo_import<<<blocks_for(n), TPB, 0, (cudaStream_t)(uintptr_t)stream>>>((double4*)(uintptr_t)p0, (double4*)(uintptr_t)p1,
                                                       (const double4*)(uintptr_t)src0, (const double4*)(uintptr_t)src1, n);
    s->launches++;
    CK(cudaGetLastError());
    return VX_OK;
}

int64_t vx_launch_count(const vx_sim* s) { return s ? s->launches : 0; }
int vx_sync(vx_sim* s) { if (!s) return VX_ERR_ARG; CK(cudaSetDevice(s->device)); CK(cudaStreamSynchronize(s->stream)); return VX_OK; }
/* takes effect at the next vx_set_voxels (the device layout depends on it); 5 <-> 7 <-> 0 on a lattice handle switch at once */
int vx_set_path(vx_sim* s, int path)
{
    if (!s) return VX_ERR_ARG;
    if (path != 0 && path != 1 && path != 3 && path != 5 && path != 7) return fail(s, VX_ERR_ARG, "vx_set_path: 0 (auto), 1 (general), 3 (small-model cluster kernel), 5 (fused, cp.async) or 7 (fused, TMA)");
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_set_path inside vx_step_begin .. vx_step_end");
    if (path != s->path) s->drop_graph();            // captured graphs hold the kernel variant
    s->path = path;
    return VX_OK;
}
int vx_active_path(const vx_sim* s) { return s && s->lattice ? 2 : 1; }
const char* vx_kernel_name(const vx_sim* s)
{
    if (s && !s->lattice && s->small && !s->collisions) return "k_small_steps (general layout, one thread-block cluster runs all steps of a call in one launch)";
    if (!s || !s->lattice) return "k_link<AXIS> (3 launches per step, one per link axis)";
    bool tma = s->path == 7;
    if (s->path == 0) {
        const int bx = (s->nx + VX_WB_X - 1) / VX_WB_X, by = (s->ny + VX_WB_Y - 1) / VX_WB_Y, bz = (s->nz + VX_WB_Z - 1) / VX_WB_Z;
        tma = (double)((bx + 1) / 2) * ((by + 1) / 2) * ((bz + 1) / 2) * 8 <= 1.1 * (double)bx * by * bz;
    }
    if (s->any_poisson || s->n_groups > 0) tma = true;            // Poisson coupling and brick-group lists live in the TMA-staged kernel only
    if (s->ghost_skip) return "k_lattice_tma<GSKIP> (fused link+voxel, 4x4x2 brick per warp, TMA staging, z-slab: no bricks on the ghost planes, 1 launch per step part)";
    return tma ? "k_lattice_tma (fused link+voxel, 4x4x2 brick per warp, TMA staging, 1 launch per step)"
               : "k_lattice_warp (fused link+voxel, 4x4x2 brick per warp, cp.async staging, 1 launch per step)";
}

#include "vx_mesh.inl"
#include "vx_linsolve.inl"

} // extern "C"

#include "vx_slabbed.hpp"      // vx_slabbed_*: one lattice on several devices of one process, composed from the entry points above

