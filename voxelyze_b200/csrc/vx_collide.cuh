// vx_collide.cuh -- self-collision: uniform-grid broadphase with a sort-by-cell pass + narrowphase, device resident.
//
// Replaces the reference's watch-list machinery:
//   CVoxelyze::updateCollisions / regenerateCollisions   src/Voxelyze.cpp:670-750   (serial O(N*S))
//   CVX_Voxel::generateNearby (5-hop exclusion lists)     src/VX_Voxel.cpp:395-418
//   CVX_Collision::updateContactForce                     src/VX_Collision.cpp:42-56
// Semantics kept exactly: a pair of surface voxels (i < j in voxelsList order) is watched iff
// |p_i - p_j|^2 <= (2.5 voxel sizes)^2 (double compare against the float threshold) and j is not
// within 5 link hops of i; the list is rebuilt when any surface voxel moved more than half a voxel
// from where it was at the last rebuild.  Contact forces are float like the reference and are
// subtracted from each voxel one by one in ascending partner index (= the reference's colWatch
// creation order), so the per-voxel sum has the reference's bit pattern.
//
// Everything is decided and done on the device; the host only enqueues (no synchronisation per step, so colliding
// runs are captured in CUDA graphs like all others):
//   every step      k_col_stale    any surface voxel further than recalcDist from its last watch position -> stale flag
//   if stale        k_col_clear .. k_col_done: the rebuild.  In a captured graph this chain is the body of a conditional IF node
//                   whose condition k_col_decide sets (cudaGraphSetConditional); launched directly, every kernel of the chain
//                   returns at once unless the flag is up.
//                     k_col_keys     cell of every surface voxel (cell edge = watch radius), hashed into a power-of-two table;
//                                    histogram of the buckets
//                     k_col_scan     exclusive prefix sum of the histogram -> bucket ranges          (one block, shuffles)
//                     k_col_scatter  counting sort: surface voxels ordered by bucket
//                     k_col_pairs    27-cell scan per surface voxel over the sorted ranges; distance, member and 5-hop tests
//                                    (the 5-hop exclusion is a precomputed 11^3-bit mask per surface voxel, tested by lattice
//                                    offset); appends the pair, counts it for both voxels
//                     k_col_scan     prefix sum of the per-voxel pair counts -> CSR of contact references
//                     k_col_fill, k_col_sort   the references of every voxel, ordered by the partner's caller index
//   every step      k_col_narrow   contact force of every watched pair from the OLD state
#pragma once
#include "vx_kernels.cuh"

namespace vxd {

#define VX_NEARBY_WORDS 42           // 11*11*11 = 1331 bits
enum { CC_STALE, CC_PAIRS, CC_OVERFLOW, CC_REBUILDS, CC_COUNT };

struct ColFrame {
    int n_surf;
    const int* surf_vox;             // internal voxel index
    const int* surf_orig;            // caller voxel index (ordering of the reference's lists)
    const int* surf_member;
    const short4* surf_ijk;          // lattice index
    const uint32_t* nearby;          // [n_surf][VX_NEARBY_WORDS]
    float4* last_watch;              // CVX_Voxel::lastColWatchPosition (float copy of pos)
    int* cell_count; int* cell_start; int* sorted; int4* cell; int hash_mask;      // buckets: histogram, ranges, surface voxels by bucket
    int2* pairs; float2* pair_kc; float4* pair_force; int cap;
    int* counters;                   // CC_*: rebuild wanted, number of pairs, list overflowed (sticky), rebuilds so far
    int* deg; int* ref_start; int* ref_fill; int* refs;
    double inv_cell;
    float thresh_sq, recalc_sq, envelope;
    const DevParams* params; int parity;     // divergence test: general path (parity < 0) or the fused step that reads generation `parity`
};

__device__ __forceinline__ int col_hash(int cx, int cy, int cz, int mask)
{
    return (int)(((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u) ^ ((unsigned)cz * 83492791u)) & mask;
}
// a diverged step returns before updateCollisions (src/Voxelyze.cpp:265-271)
__device__ __forceinline__ bool col_frozen(const ColFrame& c)
{
    const DevParams* p = c.params;
    return c.parity < 0 ? (p->div_now | p->div_latched) != 0 : (p->div_flag[c.parity ^ 1] | p->div_latched) != 0;
}
__device__ __forceinline__ bool col_rebuilding(const ColFrame& c) { return c.counters[CC_STALE] != 0 && !col_frozen(c); }
__device__ __forceinline__ int col_pair_count(const ColFrame& c) { const int n = c.counters[CC_PAIRS]; return n < c.cap ? n : c.cap; }

// any surface voxel moved further than recalcDist since the last rebuild?  (src/Voxelyze.cpp:688-696)
__global__ void k_col_stale(Frame f, ColFrame c)
{
    if (col_frozen(c)) return;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < c.n_surf; s += gridDim.x * blockDim.x) {
        double4 p = f.pose0[c.surf_vox[s]];
        float4 w = c.last_watch[s];
        double dx = p.x - w.x, dy = p.y - w.y, dz = p.z - w.z;
        if (dx * dx + dy * dy + dz * dz > c.recalc_sq) c.counters[CC_STALE] = 1;
    }
}
// condition of the graph's IF node around the rebuild chain
__global__ void k_col_decide(cudaGraphConditionalHandle handle, ColFrame c)
{
    cudaGraphSetConditional(handle, col_rebuilding(c) ? 1u : 0u);
}
__global__ void k_col_mark_stale(int* counters) { counters[CC_STALE] = 1; }

__global__ void k_col_clear(ColFrame c)
{
    if (!col_rebuilding(c)) return;
    const int n = max(c.hash_mask + 1, c.n_surf);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i <= c.hash_mask) c.cell_count[i] = 0;
        if (i < c.n_surf) { c.deg[i] = 0; c.ref_fill[i] = 0; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { c.counters[CC_PAIRS] = 0; c.counters[CC_REBUILDS] += 1; }
}

__global__ void k_col_keys(Frame f, ColFrame c)
{
    if (!col_rebuilding(c)) return;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < c.n_surf; s += gridDim.x * blockDim.x) {
        double4 p = f.pose0[c.surf_vox[s]];
        c.last_watch[s] = make_float4((float)p.x, (float)p.y, (float)p.z, 0.0f);        // src/Voxelyze.cpp:733
        int cx = (int)floor(p.x * c.inv_cell), cy = (int)floor(p.y * c.inv_cell), cz = (int)floor(p.z * c.inv_cell);
        const int key = col_hash(cx, cy, cz, c.hash_mask);
        c.cell[s] = make_int4(cx, cy, cz, key);
        atomicAdd(&c.cell_count[key], 1);
    }
}

// out[i] = in[0] + .. + in[i-1] for i in [0, n]; ONE block of 1024 threads (rebuilds are rare and n is a few 10^5 at most)
__global__ void __launch_bounds__(1024) k_col_scan(ColFrame c, const int* in, int* out, int n)
{
    if (!col_rebuilding(c)) return;
    __shared__ int warp_sum[32];
    __shared__ int carry_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sum[lane], t = w;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            warp_sum[lane] = t - w;                                   // exclusive over the warps
        }
        __syncthreads();
        const int carry = carry_sh;
        if (i < n) out[i] = carry + warp_sum[warp] + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_sh = carry + warp_sum[warp] + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_sh;
}

// counting sort by bucket (the histogram is consumed)
__global__ void k_col_scatter(ColFrame c)
{
    if (!col_rebuilding(c)) return;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < c.n_surf; s += gridDim.x * blockDim.x) {
        const int key = c.cell[s].w;
        c.sorted[c.cell_start[key] + atomicSub(&c.cell_count[key], 1) - 1] = s;
    }
}

// watched pairs found by the voxel with the smaller caller index (src/Voxelyze.cpp:730-747)
__global__ void k_col_pairs(Frame f, ColFrame c)
{
    if (!col_rebuilding(c)) return;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < c.n_surf; a += gridDim.x * blockDim.x) {
        const double4 pa = f.pose0[c.surf_vox[a]];
        const int4 ca = c.cell[a];
        const short4 ia = c.surf_ijk[a];
        const int oa = c.surf_orig[a], ma = c.surf_member[a];
        const uint32_t* near_a = c.nearby + (size_t)a * VX_NEARBY_WORDS;
        for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            const int cx = ca.x + dx, cy = ca.y + dy, cz = ca.z + dz;
            const int key = col_hash(cx, cy, cz, c.hash_mask);
            for (int j = c.cell_start[key], je = c.cell_start[key + 1]; j < je; j++) {
                const int b = c.sorted[j];
                const int4 cb = c.cell[b];
                if (cb.x != cx || cb.y != cy || cb.z != cz) continue;      // another cell sharing the bucket
                if (c.surf_orig[b] <= oa || c.surf_member[b] != ma) continue;
                const double4 pb = f.pose0[c.surf_vox[b]];
                const double ex = pa.x - pb.x, ey = pa.y - pb.y, ez = pa.z - pb.z;
                if (ex * ex + ey * ey + ez * ez > c.thresh_sq) continue;    // double > float, like the reference
                const short4 ib = c.surf_ijk[b];
                const int ox = ib.x - ia.x, oy = ib.y - ia.y, oz = ib.z - ia.z;
                if (ox >= -5 && ox <= 5 && oy >= -5 && oy <= 5 && oz >= -5 && oz <= 5) {
                    const int bit = ((oz + 5) * 11 + (oy + 5)) * 11 + (ox + 5);
                    if (near_a[bit >> 5] & (1u << (bit & 31))) continue;    // within 5 link hops
                }
                const int slot = atomicAdd(&c.counters[CC_PAIRS], 1);
                if (slot >= c.cap) { c.counters[CC_OVERFLOW] = 1; continue; }
                c.pairs[slot] = make_int2(a, b);
                atomicAdd(&c.deg[a], 1); atomicAdd(&c.deg[b], 1);
                // CVX_Collision constructor (src/VX_Collision.cpp:17-23)
                const DevVoxMat& m1 = f.vmat[meta_hi(f.pose1[c.surf_vox[a]].w) & VM_MAT_MASK];
                const DevVoxMat& m2 = f.vmat[meta_hi(f.pose1[c.surf_vox[b]].w) & VM_MAT_MASK];
                c.pair_kc[slot] = make_float2(2.0f / (1.0f / m1.pen_stiff + 1.0f / m2.pen_stiff), 0.5f * (m1.coll_damp_t + m2.coll_damp_t));
                c.pair_force[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

__global__ void k_col_fill(ColFrame c)
{
    if (!col_rebuilding(c)) return;
    const int n_pairs = col_pair_count(c);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += gridDim.x * blockDim.x) {
        int2 ab = c.pairs[p];
        c.refs[c.ref_start[ab.x] + atomicAdd(&c.ref_fill[ab.x], 1)] = p * 2;          // this voxel is voxel1: +force
        c.refs[c.ref_start[ab.y] + atomicAdd(&c.ref_fill[ab.y], 1)] = p * 2 + 1;      // voxel2: -force
    }
}

// order every voxel's contact list by the partner's caller index = colWatch creation order
__global__ void k_col_sort(ColFrame c)
{
    if (!col_rebuilding(c)) return;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < c.n_surf; s += gridDim.x * blockDim.x) {
        const int b = c.ref_start[s], e = c.ref_start[s + 1];
        auto partner = [&](int ref) { int2 ab = c.pairs[ref >> 1]; return c.surf_orig[(ref & 1) ? ab.x : ab.y]; };
        for (int i = b + 1; i < e; i++) {
            int r = c.refs[i], key = partner(r), j = i - 1;
            while (j >= b && partner(c.refs[j]) > key) { c.refs[j + 1] = c.refs[j]; j--; }
            c.refs[j + 1] = r;
        }
    }
}
// last kernel of the rebuild chain
__global__ void k_col_done(ColFrame c)
{
    if (!col_rebuilding(c)) return;
    c.counters[CC_STALE] = 0;
}

// CVX_Collision::updateContactForce (src/VX_Collision.cpp:42-56), float arithmetic
__global__ void k_col_narrow(Frame f, ColFrame c)
{
    if (col_frozen(c)) return;
    const int n_pairs = col_pair_count(c);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += gridDim.x * blockDim.x) {
        const int2 ab = c.pairs[p];
        const int v1 = c.surf_vox[ab.x], v2 = c.surf_vox[ab.y];
        const double4 a0 = f.pose0[v1], b0 = f.pose0[v2];
        const double w1 = f.pose1[v1].w, w2 = f.pose1[v2].w;
        const DevVoxMat& m1 = f.vmat[meta_hi(w1) & VM_MAT_MASK];
        const DevVoxMat& m2 = f.vmat[meta_hi(w2) & VM_MAT_MASK];
        const float ox = (float)(b0.x - a0.x), oy = (float)(b0.y - a0.y), oz = (float)(b0.z - a0.z);
        const float t1 = meta_temp(w1), t2 = meta_temp(w2);
        const double bs1 = (base_size(m1, 0, t1) + base_size(m1, 1, t1) + base_size(m1, 2, t1)) / 3.0f;
        const double bs2 = (base_size(m2, 0, t2) + base_size(m2, 1, t2) + base_size(m2, 2, t2)) / 3.0f;
        const float nom_dist = (float)((bs1 + bs2) * c.envelope);
        const float len = sqrtf(ox * ox + oy * oy + oz * oz);
        const float rel = nom_dist - len;
        float4 force = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rel > 0) {
            float ux = ox, uy = oy, uz = oz;
            if (len > 0) { float inv = 1.0f / len; ux = inv * ox; uy = inv * oy; uz = inv * oz; }
            const double4 ma = f.mom0[v1], mb = f.mom0[v2];
            const double dux = ux, duy = uy, duz = uz;
            const double va = (m1.mass_inv_d * ma.x) * dux + (m1.mass_inv_d * ma.y) * duy + (m1.mass_inv_d * ma.z) * duz;
            const double vb = (m2.mass_inv_d * mb.x) * dux + (m2.mass_inv_d * mb.y) * duy + (m2.mass_inv_d * mb.z) * duz;
            const float rel_vel = (float)(va - vb);
            const float2 kc = c.pair_kc[p];
            const float mag = kc.x * rel + kc.y * rel_vel;
            force = make_float4(mag * ux, mag * uy, mag * uz, 0.f);
        }
        c.pair_force[p] = force;
    }
}

} // namespace vxd

// =================================================================================================
// stateInfo reductions (CVoxelyze::stateInfo, src/Voxelyze.cpp:752-800): per-element float value
// exactly as the reference computes it, reduced with warp shuffles + one atomic per block.
// MIN/MAX use ordered-integer atomics on the float bit pattern, TOTAL accumulates in double
// (the reference adds floats in list order; the difference is below 1e-6 relative).
// =================================================================================================
namespace vxd {

enum { SI_DISPLACEMENT, SI_VELOCITY, SI_KINETIC_ENERGY, SI_ANGULAR_DISPLACEMENT, SI_ANGULAR_VELOCITY,
       SI_ENG_STRESS, SI_ENG_STRAIN, SI_STRAIN_ENERGY, SI_PRESSURE, SI_MASS };
enum { SI_MIN, SI_MAX, SI_TOTAL, SI_AVERAGE };

struct StateAcc { float mn, mx; double sum; };

__device__ __forceinline__ void si_block_reduce(StateAcc a, float* out_min, float* out_max, double* out_sum)
{
    for (int o = 16; o > 0; o >>= 1) {
        a.mn = fminf(a.mn, __shfl_xor_sync(0xffffffffu, a.mn, o));
        a.mx = fmaxf(a.mx, __shfl_xor_sync(0xffffffffu, a.mx, o));
        a.sum += __shfl_xor_sync(0xffffffffu, a.sum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        // float atomic min/max through the monotone int mapping of IEEE floats
        int imn = __float_as_int(a.mn), imx = __float_as_int(a.mx);
        if (imn >= 0) atomicMin((int*)out_min, imn); else atomicMax((unsigned int*)out_min, (unsigned int)imn);
        if (imx >= 0) atomicMax((int*)out_max, imx); else atomicMin((unsigned int*)out_max, (unsigned int)imx);
        atomicAdd(out_sum, a.sum);
    }
}

// voxel quantities; ijk_size = nominal position source: displacement needs the lattice index, passed as a
// per-voxel nominal position array only when info == DISPLACEMENT
// vals (optional): the value of every voxel, internal order (the mesh colouring reads it)
__global__ void __launch_bounds__(256) k_state_voxels(Frame f, int info, const double4* nominal, float* out_min, float* out_max, double* out_sum, float* vals = nullptr)
{
    StateAcc a; a.mn = 3.402823466e38f; a.mx = -3.402823466e38f; a.sum = 0.0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.n_vox; v += gridDim.x * blockDim.x) {
        const double4 p0 = f.pose0[v], p1 = f.pose1[v];
        if (meta_hi(p1.w) & VM_GHOST) continue;           // halo copies and the fill cells of a box with holes are not voxels of this model
        const DevVoxMat& m = f.vmat[meta_hi(p1.w) & VM_MAT_MASK];
        float val = 0.0f;
        if (info == SI_DISPLACEMENT) { double4 n = nominal[v]; double dx = p0.x - n.x, dy = p0.y - n.y, dz = p0.z - n.z; val = (float)sqrt(dx * dx + dy * dy + dz * dz); }
        else if (info == SI_VELOCITY) { double4 l = f.mom0[v]; val = (float)(sqrt(l.x * l.x + l.y * l.y + l.z * l.z) * m.mass_inv); }
        else if (info == SI_ANGULAR_VELOCITY) { double4 l = f.mom0[v]; double2 l1 = f.mom1[v]; val = (float)(sqrt(l.w * l.w + l1.x * l1.x + l1.y * l1.y) * m.inertia_inv); }
        else if (info == SI_KINETIC_ENERGY) {
            double4 l = f.mom0[v]; double2 l1 = f.mom1[v];
            val = (float)(0.5 * (m.mass_inv * (l.x * l.x + l.y * l.y + l.z * l.z) + m.inertia_inv * (l.w * l.w + l1.x * l1.x + l1.y * l1.y)));
        }
        else if (info == SI_ANGULAR_DISPLACEMENT) { double w = p0.w; val = (float)(2.0 * acos(w > 1 ? 1.0 : w)); }
        else if (info == SI_MASS) val = m.mass;
        if (vals) vals[v] = val;
        a.mn = fminf(a.mn, val); a.mx = fmaxf(a.mx, val); a.sum += (double)val;
    }
    si_block_reduce(a, out_min, out_max, out_sum);
}

// CVX_Voxel::pressure (include/VX_Voxel.h:103-104): -E * volumetricStrain / (3 (1 - 2 nu)), volumetric strain = sum over the
// axes of the voxel's strain (src/VX_Voxel.cpp:300-317: half-link strains, averaged when links exist on both sides), float.
// vlinks[d * n + v]: caller index of the link of voxel v (caller order) in direction d or -1; strain: per-link axial strain;
// ratio: CVX_Link::strainRatio (src/VX_Link.cpp:67); en: per voxel {E, nu}
// skip (optional): 1 for the halo copies of a z-slab -- they are not voxels of this model (their owner counts them)
__global__ void __launch_bounds__(256) k_state_pressure(int n, const int* vlinks, const float* strain, const float* ratio, const float2* en,
                                                        float* out_min, float* out_max, double* out_sum, float* vals = nullptr, const unsigned char* skip = nullptr)
{
    StateAcc a; a.mn = 3.402823466e38f; a.mx = -3.402823466e38f; a.sum = 0.0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        if (skip && skip[v]) { if (vals) vals[v] = 0.f; continue; }
        float s3[3] = {0.0f, 0.0f, 0.0f}; int cnt[3] = {0, 0, 0};
        for (int d = 0; d < 6; d++) {
            const int l = vlinks[(size_t)d * n + v];
            if (l < 0) continue;
            const float e = strain[l], r = ratio[l];
            // direction d even: this voxel is the link's negative end (src/VX_Link.cpp:121-124)
            s3[d >> 1] += (d & 1) ? 2.0f * e * r / (1.0f + r) : 2.0f * e / (1.0f + r);
            cnt[d >> 1]++;
        }
        for (int k = 0; k < 3; k++) if (cnt[k] == 2) s3[k] *= 0.5f;
        const float vol = (float)(s3[0] + s3[1] + s3[2]);
        const float2 m = en[v];
        const float val = -m.x * vol / (3 * (1 - 2 * m.y));
        if (vals) vals[v] = val;
        a.mn = fminf(a.mn, val); a.mx = fmaxf(a.mx, val); a.sum += (double)val;
    }
    si_block_reduce(a, out_min, out_max, out_sum);
}

// link quantities from a flat value array produced by the gather kernels (strain / stress) or from
// force+moment triples (strain energy, src/VX_Link.cpp:251-257)
__global__ void __launch_bounds__(256) k_state_links(int n, const float* scalar, const double* fneg, const double* mneg, const double* mpos,
                                                     const float* a1, const float* a2, const float* b3, float* out_min, float* out_max, double* out_sum, float* vals = nullptr)
{
    StateAcc a; a.mn = 3.402823466e38f; a.mx = -3.402823466e38f; a.sum = 0.0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n; l += gridDim.x * blockDim.x) {
        float val;
        if (scalar) val = scalar[l];
        else {
            const double fx = fneg[3 * l], nx = mneg[3 * l], ny = mneg[3 * l + 1], nz = mneg[3 * l + 2], py = mpos[3 * l + 1], pz = mpos[3 * l + 2];
            val = fx * fx / (2.0f * a1[l]) + nx * nx / (2.0 * a2[l]) + (nz * nz - nz * pz + pz * pz) / (3.0 * b3[l]) + (ny * ny - ny * py + py * py) / (3.0 * b3[l]);
        }
        if (vals) vals[l] = val;
        a.mn = fminf(a.mn, val); a.mx = fmaxf(a.mx, val); a.sum += (double)val;
    }
    si_block_reduce(a, out_min, out_max, out_sum);
}

} // namespace vxd
