// vx_lattice_variants.cuh -- the fused lattice step in the formulations that lost to k_lattice_warp / k_lattice_tma.
// They stay selectable (vx_set_path 2, 3, 4, 6) because they are the ablation the design rests on and because every one of
// them must produce the same bits as the general path (tests/test_gpu_parity.py).  Included at the end of vx_lattice.cuh.
#pragma once

namespace vxd {

// dt lives in device memory (p->dt) so that captured graphs survive a change of time step.
// first_of_call: the first step of a vx_step call damps with the previous call's dt
// (CVX_Voxel::previousDt), all later steps of the call with dt itself.
#ifndef VX_LAT_MINBLOCKS
#define VX_LAT_MINBLOCKS 1
#endif
#ifndef VX_LAT_UNROLL
#define VX_LAT_UNROLL 1
#endif
template <bool UNI>
__global__ void __launch_bounds__(128, VX_LAT_MINBLOCKS) k_lattice_step(LatFrame f, int parity, int first_of_call, int floor_on)
{
    DevParams* p = f.params;
    const int frozen = p->div_flag[parity ^ 1] | p->div_latched;   // did the previous step diverge?
    const float dt = p->dt;
    const float prev_dt = first_of_call ? p->prev_dt : dt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {                     // bookkeeping of the previous step
        if (frozen) p->div_latched = 1;
        else if (p->pending) { p->steps_done += 1; p->time += dt; }
        if (!frozen) p->pending = 1;
    }
    if (frozen) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= f.n_vox) return;

    const double4 s0 = ld4(f.c_pose0 + v), s1 = ld4(f.c_pose1 + v);
    VoxelState vs;
    vs.bits = meta_hi(s1.w);
    vs.temp = meta_temp(s1.w);
    const uint32_t mask = (vs.bits >> VM_LINK_SHIFT) & 0x3Fu;
    uint32_t new_bits = vs.bits;

    d3 F = mk3(0.0, 0.0, 0.0), M = mk3(0.0, 0.0, 0.0);
    constexpr int kUnroll = VX_LAT_UNROLL;
#pragma unroll kUnroll
    for (int k = 0; k < 6; k++) {                                  // slot order = reference summation order
        if (!(mask & (1u << k))) continue;
        const int axis = k >> 1;
        const bool i_am_neg = (k & 1) == 0;
        const int stride = axis == 0 ? 1 : (axis == 1 ? f.nx : f.nxy);
        const int u = i_am_neg ? v + stride : v - stride;
        const double4 u0 = ld4(f.c_pose0 + u), u1 = ld4(f.c_pose1 + u);
        LinkState st;
        d3 fN, mN, fP, mP;
        // one inlined copy of the link physics serves both roles: select the operands first
        const double4 n0 = i_am_neg ? s0 : u0, n1 = i_am_neg ? s1 : u1;
        const double4 p0 = i_am_neg ? u0 : s0, p1 = i_am_neg ? u1 : s1;
        lat_eval_link<UNI>(f, axis, i_am_neg ? v : u, i_am_neg ? vs.bits : meta_hi(u1.w), n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP);
        if (i_am_neg) {
            F = F + fN; M = M + mN;
            double2 ra, rb, rc; float4 rs; uint32_t lf;
            lat_encode(st, ra, rb, rc, rs, lf);
            f.n_rec[axis][0][v] = ra; f.n_rec[axis][1][v] = rb; f.n_rec[axis][2][v] = rc; f.n_recf[axis][v] = rs;
            new_bits = (new_bits & ~(3u << (VM_LFLAG_SHIFT + 2 * axis))) | (lf << (VM_LFLAG_SHIFT + 2 * axis));
            if (st.strain > 100) p->div_flag[parity] = 1;          // src/Voxelyze.cpp:265
        } else {
            F = F + fP; M = M + mP;
        }
    }

    double4 m0 = f.c_mom0[v]; double2 m1 = f.c_mom1[v];
    vs.bits = new_bits;
    vs.pos = mk3(s0.x, s0.y, s0.z);
    vs.orient.w = s0.w; vs.orient.x = s1.x; vs.orient.y = s1.y; vs.orient.z = s1.z;
    vs.lin = mk3(m0.x, m0.y, m0.z);
    vs.ang = mk3(m0.w, m1.x, m1.y);
    if (!(vs.bits & VM_GHOST)) {
        const DevVoxMat& vm = UNI ? f.vm0 : f.vmat[vs.bits & VM_MAT_MASK];
        const DevExt* ext = (vs.bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
        voxel_integrate(vs, F, M, nullptr, 0, nullptr, vm, ext, dt, floor_on != 0);
    }
    f.n_pose0[v] = make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w);
    f.n_pose1[v] = make_double4(vs.orient.x, vs.orient.y, vs.orient.z, meta_pack(vs.temp, vs.bits));
    f.n_mom0[v] = make_double4(vs.lin.x, vs.lin.y, vs.lin.z, vs.ang.x);
    f.n_mom1[v] = make_double2(vs.ang.y, vs.ang.z);
}


// =================================================================================================
// k_lattice_march -- the fused step with the redundant link evaluations removed where the hardware
// offers a free exchange path:
//   X: a warp covers 31 consecutive voxels of one x-row plus one overlap lane (lane 0 = last voxel
//      of the previous segment).  Every lane evaluates the +X link of its voxel once; the force on
//      the positive end travels one lane up with __shfl_up_sync.
//   Z: the warp marches along z over a chunk of ZL planes; the force its +Z link exerts on the
//      voxel above is carried in registers to the next iteration (where that voxel is "me"), and
//      the pose of the voxel above is loaded once and becomes the own pose of the next iteration.
//   Y: the -Y link is still re-evaluated by the positive-end voxel (identical inputs, identical bits).
// => 1 + 1/31 (X) + 2 (Y) + 1 + 1/ZL (Z) = ~4.1 link evaluations per voxel instead of 6, and the
//    long-distance (one z-plane) re-read of neighbour poses and link records disappears.
// Summation order per voxel is unchanged: X+, X-, Y+, Y-, Z+, Z- (src/VX_Voxel.cpp:238-240).
// =================================================================================================
struct Pose { double4 a, b; };          // a = {pos.xyz, orient.w}, b = {orient.xyz, meta}

__device__ __forceinline__ Pose lat_load_pose(const LatFrame& f, int v) { Pose p; p.a = ld4(f.c_pose0 + v); p.b = ld4(f.c_pose1 + v); return p; }
__device__ __forceinline__ double shfl_down_d(double x, int d) { return __shfl_down_sync(0xffffffffu, x, d); }
__device__ __forceinline__ double shfl_up_d(double x, int d) { return __shfl_up_sync(0xffffffffu, x, d); }
__device__ __forceinline__ Pose shfl_down_pose(const Pose& p)
{
    Pose r;
    r.a = make_double4(shfl_down_d(p.a.x, 1), shfl_down_d(p.a.y, 1), shfl_down_d(p.a.z, 1), shfl_down_d(p.a.w, 1));
    r.b = make_double4(shfl_down_d(p.b.x, 1), shfl_down_d(p.b.y, 1), shfl_down_d(p.b.z, 1), shfl_down_d(p.b.w, 1));
    return r;
}
__device__ __forceinline__ d3 shfl_up_d3(d3 v) { return mk3(shfl_up_d(v.x, 1), shfl_up_d(v.y, 1), shfl_up_d(v.z, 1)); }

// evaluates link (owner, AXIS); OWNER: also stores the advanced record and returns the new mode bits
template <int AXIS, bool OWNER, bool UNI>
__device__ __forceinline__ void lat_link(const LatFrame& f, int owner, uint32_t owner_bits, const Pose& N, const Pose& P,
                                         float prev_dt, int parity, uint32_t& new_bits, d3& fN, d3& mN, d3& fP, d3& mP)
{
    LinkState st;
    lat_eval_link<UNI>(f, AXIS, owner, owner_bits, N.a, N.b, P.a, P.b, prev_dt, st, fN, mN, fP, mP);
    if (OWNER) {
        double2 ra, rb, rc; float4 rs; uint32_t lf;
        lat_encode(st, ra, rb, rc, rs, lf);
        f.n_rec[AXIS][0][owner] = ra; f.n_rec[AXIS][1][owner] = rb; f.n_rec[AXIS][2][owner] = rc; f.n_recf[AXIS][owner] = rs;
        new_bits = (new_bits & ~(3u << (VM_LFLAG_SHIFT + 2 * AXIS))) | (lf << (VM_LFLAG_SHIFT + 2 * AXIS));
        if (st.strain > 100) f.params->div_flag[parity] = 1;      // src/Voxelyze.cpp:265
    }
}

#ifndef VX_MARCH_ZL
#define VX_MARCH_ZL 32
#endif
#ifndef VX_MARCH_MINBLOCKS
#define VX_MARCH_MINBLOCKS 3
#endif
#define VX_MARCH_ROWS 4                 // warps (y-rows) per CTA

template <bool UNI>
__global__ void __launch_bounds__(32 * VX_MARCH_ROWS, VX_MARCH_MINBLOCKS)
k_lattice_march(LatFrame f, int parity, int first_of_call, int floor_on, int n_seg, int n_yg, int n_zc)
{
    DevParams* p = f.params;
    const int frozen = p->div_flag[parity ^ 1] | p->div_latched;
    const float dt = p->dt;
    const float prev_dt = first_of_call ? p->prev_dt : dt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (frozen) p->div_latched = 1;
        else if (p->pending) { p->steps_done += 1; p->time += dt; }
        if (!frozen) p->pending = 1;
    }
    if (frozen) return;

    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    int b = blockIdx.x;
    const int seg = b % n_seg; b /= n_seg;
    const int yg = b % n_yg; b /= n_yg;
    const int zc = b % n_zc; const int member = b / n_zc;
    const int y = yg * VX_MARCH_ROWS + row;
    if (y >= f.ny) return;                                         // whole warp
    const int x = seg * 31 - 1 + lane;
    const bool has_voxel = x >= 0 && x < f.nx;
    const bool real = has_voxel && lane >= 1;                      // lane 0 only feeds the X- force of lane 1
    const int z0 = zc * VX_MARCH_ZL;
    const int z1 = min(z0 + VX_MARCH_ZL, f.nz);
    const int xs = has_voxel ? x : (x < 0 ? 0 : f.nx - 1);         // clamp so that idle lanes load valid memory
    int v = ((member * f.nz + z0) * f.ny + y) * f.nx + xs;

    Pose S = lat_load_pose(f, v);
    d3 zf = mk3(0.0, 0.0, 0.0), zm = mk3(0.0, 0.0, 0.0);           // force/moment of the -Z link on this voxel
    if (real && z0 > 0 && ((meta_hi(S.b.w) >> VM_LINK_SHIFT) & 0x20u)) {   // chunk start: re-evaluate the link from below
        Pose D = lat_load_pose(f, v - f.nxy);
        d3 fN, mN; uint32_t nb = 0;
        lat_link<2, false, UNI>(f, v - f.nxy, meta_hi(D.b.w), D, S, prev_dt, parity, nb, fN, mN, zf, zm);
    }

    for (int z = z0; z < z1; z++, v += f.nxy) {
        const uint32_t bits = meta_hi(S.b.w);
        const uint32_t mask = has_voxel ? ((bits >> VM_LINK_SHIFT) & 0x3Fu) : 0u;
        uint32_t new_bits = bits;
        d3 F = mk3(0.0, 0.0, 0.0), M = mk3(0.0, 0.0, 0.0);
        d3 fN, mN, fP, mP;

        // ---- X: one evaluation per link, positive-end force shuffled one lane up
        Pose XP = shfl_down_pose(S);
        if (lane == 31 && (mask & 0x01u)) XP = lat_load_pose(f, v + 1);
        d3 xf = mk3(0.0, 0.0, 0.0), xm = mk3(0.0, 0.0, 0.0);
        if (mask & 0x01u) {
            if (real) lat_link<0, true, UNI>(f, v, bits, S, XP, prev_dt, parity, new_bits, fN, mN, xf, xm);
            else lat_link<0, false, UNI>(f, v, bits, S, XP, prev_dt, parity, new_bits, fN, mN, xf, xm);
            F = F + fN; M = M + mN;
        }
        xf = shfl_up_d3(xf); xm = shfl_up_d3(xm);
        if (mask & 0x02u) { F = F + xf; M = M + xm; }

        // the voxel above becomes "me" in the next iteration
        Pose U = S;
        if (z + 1 < f.nz && has_voxel) U = lat_load_pose(f, v + f.nxy);

        if (real) {
            // ---- Y: own +Y link, then the -Y link re-evaluated from the positive end
            if (mask & 0x04u) {
                Pose YP = lat_load_pose(f, v + f.nx);
                lat_link<1, true, UNI>(f, v, bits, S, YP, prev_dt, parity, new_bits, fN, mN, fP, mP);
                F = F + fN; M = M + mN;
            }
            if (mask & 0x08u) {
                Pose YM = lat_load_pose(f, v - f.nx);
                uint32_t nb = 0;
                lat_link<1, false, UNI>(f, v - f.nx, meta_hi(YM.b.w), YM, S, prev_dt, parity, nb, fN, mN, fP, mP);
                F = F + fP; M = M + mP;
            }
            // ---- Z: own +Z link (its force on the voxel above is carried), then the carried -Z force
            d3 nzf = mk3(0.0, 0.0, 0.0), nzm = mk3(0.0, 0.0, 0.0);
            if (mask & 0x10u) {
                lat_link<2, true, UNI>(f, v, bits, S, U, prev_dt, parity, new_bits, fN, mN, nzf, nzm);
                F = F + fN; M = M + mN;
            }
            if (mask & 0x20u) { F = F + zf; M = M + zm; }
            zf = nzf; zm = nzm;

            // ---- integrate and store the next generation
            double4 m0 = ld4(f.c_mom0 + v); double2 m1 = __ldg(f.c_mom1 + v);
            VoxelState vs;
            vs.bits = new_bits; vs.temp = meta_temp(S.b.w);
            vs.pos = mk3(S.a.x, S.a.y, S.a.z);
            vs.orient.w = S.a.w; vs.orient.x = S.b.x; vs.orient.y = S.b.y; vs.orient.z = S.b.z;
            vs.lin = mk3(m0.x, m0.y, m0.z);
            vs.ang = mk3(m0.w, m1.x, m1.y);
            if (!(vs.bits & VM_GHOST)) {
                const DevVoxMat& vm = UNI ? f.vm0 : f.vmat[vs.bits & VM_MAT_MASK];
                const DevExt* ext = (vs.bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
                voxel_integrate(vs, F, M, nullptr, 0, nullptr, vm, ext, dt, floor_on != 0);
            }
            f.n_pose0[v] = make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w);
            f.n_pose1[v] = make_double4(vs.orient.x, vs.orient.y, vs.orient.z, meta_pack(vs.temp, vs.bits));
            f.n_mom0[v] = make_double4(vs.lin.x, vs.lin.y, vs.lin.z, vs.ang.x);
            f.n_mom1[v] = make_double2(vs.ang.y, vs.ang.z);
        }
        S = U;
    }
}


// =================================================================================================
// k_lattice_tile -- fused step, tile formulation (the default lattice kernel).
//
// A CTA owns a TX x TY x TZ = 8 x 4 x 4 brick of voxels.  Phase 1 is "one thread per link
// evaluation" exactly like the general link kernel (short threads, high occupancy): the 464 links
// that touch the brick (304 inside it, 160 crossing one of its faces) are evaluated once each and
// the force/moment on each end that lies inside the brick is written to a shared-memory slot
// slot[link direction][component][voxel].  Links crossing a face are evaluated by both bricks from
// identical inputs; only the brick holding the negative end (the owner) stores the advanced record.
// Phase 2 is "one thread per voxel": gather the six slots in reference order, integrate, store.
// => 3.6 link evaluations per voxel instead of 6, forces never leave the SM, and both phases keep
//    the simple, wide thread shape that runs at DRAM speed in the general path.
// =================================================================================================
#ifndef VX_TILE_X
#define VX_TILE_X 8
#endif
#ifndef VX_TILE_Y
#define VX_TILE_Y 4
#endif
#ifndef VX_TILE_Z
#define VX_TILE_Z 4
#endif
#define VX_TILE_VOX (VX_TILE_X * VX_TILE_Y * VX_TILE_Z)                       // 128
#define VX_TILE_EX ((VX_TILE_X + 1) * VX_TILE_Y * VX_TILE_Z)                  // 144 x-links
#define VX_TILE_EY (VX_TILE_X * (VX_TILE_Y + 1) * VX_TILE_Z)                  // 160 y-links
#define VX_TILE_EZ (VX_TILE_X * VX_TILE_Y * (VX_TILE_Z + 1))                  // 160 z-links
#define VX_TILE_EVALS (VX_TILE_EX + VX_TILE_EY + VX_TILE_EZ)                  // 464
#ifndef VX_TILE_THREADS
#define VX_TILE_THREADS 256
#endif
#define VX_TILE_ROUNDS ((VX_TILE_EVALS + VX_TILE_THREADS - 1) / VX_TILE_THREADS)
#ifndef VX_TILE_MINBLOCKS
#define VX_TILE_MINBLOCKS 2
#endif

#define VX_TILE_HX (VX_TILE_X + 2)
#define VX_TILE_HY (VX_TILE_Y + 2)
#define VX_TILE_HZ (VX_TILE_Z + 2)
#define VX_TILE_HALO (VX_TILE_HX * VX_TILE_HY * VX_TILE_HZ)                    // 360 staged poses

#define VX_TILE_SMEM (2 * VX_TILE_HALO * 32 + 6 * 6 * VX_TILE_VOX * 8 + 3 * VX_TILE_VOX)   // 60 288 B

struct TileEval { int axis, lx, ly, lz, vn; bool ok; };

__device__ __forceinline__ TileEval tile_decode(const LatFrame& f, int e, int tx0, int ty0, int tz0, int vbase)
{
    TileEval t;
    if (e < VX_TILE_EX) { t.axis = 0; t.lx = e % (VX_TILE_X + 1) - 1; int r = e / (VX_TILE_X + 1); t.ly = r % VX_TILE_Y; t.lz = r / VX_TILE_Y; }
    else if (e < VX_TILE_EX + VX_TILE_EY) { int q = e - VX_TILE_EX; t.axis = 1; t.lx = q % VX_TILE_X; int r = q / VX_TILE_X; t.ly = r % (VX_TILE_Y + 1) - 1; t.lz = r / (VX_TILE_Y + 1); }
    else { int q = e - VX_TILE_EX - VX_TILE_EY; t.axis = 2; t.lx = q % VX_TILE_X; int r = q / VX_TILE_X; t.ly = r % VX_TILE_Y; t.lz = r / VX_TILE_Y - 1; }
    const int gx = tx0 + t.lx, gy = ty0 + t.ly, gz = tz0 + t.lz;
    const int px = gx + (t.axis == 0), py = gy + (t.axis == 1), pz = gz + (t.axis == 2);
    t.ok = e < VX_TILE_EVALS && gx >= 0 && gy >= 0 && gz >= 0 && px < f.nx && py < f.ny && pz < f.nz;
    t.vn = vbase + (gz * f.ny + gy) * f.nx + gx;
    return t;
}

template <bool UNI>
__global__ void __launch_bounds__(VX_TILE_THREADS, VX_TILE_MINBLOCKS)
k_lattice_tile(LatFrame f, int parity, int first_of_call, int floor_on, int ntx, int nty, int ntz)
{
    // dynamic shared memory (VX_TILE_SMEM bytes, > 48 KB so it is opted in by the host):
    extern __shared__ __align__(16) unsigned char tile_smem[];
    double4* sp0 = reinterpret_cast<double4*>(tile_smem);                        // staged poses, brick + halo  11.5 KB
    double4* sp1 = sp0 + VX_TILE_HALO;                                           //                             11.5 KB
    double (*slot)[6][VX_TILE_VOX] = reinterpret_cast<double (*)[6][VX_TILE_VOX]>(sp1 + VX_TILE_HALO);   // [dir][comp][voxel] 36 KB
    unsigned char (*lflag_sh)[VX_TILE_VOX] = reinterpret_cast<unsigned char (*)[VX_TILE_VOX]>(slot + 6); // new mode bits of owned links
    DevParams* p = f.params;
    const int frozen = p->div_flag[parity ^ 1] | p->div_latched;
    const float dt = p->dt;
    const float prev_dt = first_of_call ? p->prev_dt : dt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (frozen) p->div_latched = 1;
        else if (p->pending) { p->steps_done += 1; p->time += dt; }
        if (!frozen) p->pending = 1;
    }
    if (frozen) return;

    int b = blockIdx.x;
    const int tx0 = (b % ntx) * VX_TILE_X; b /= ntx;
    const int ty0 = (b % nty) * VX_TILE_Y; b /= nty;
    const int tz0 = (b % ntz) * VX_TILE_Z; const int member = b / ntz;
    const int vbase = member * f.nz * f.nxy;

    // ---- phase 0: stage the poses the brick needs (its own voxels + face neighbours) in shared memory;
    //      every thread has its loads in flight at once, the evaluations below never wait on a pose
    for (int c = threadIdx.x; c < VX_TILE_HALO; c += VX_TILE_THREADS) {
        const int hx = c % VX_TILE_HX, hy = (c / VX_TILE_HX) % VX_TILE_HY, hz = c / (VX_TILE_HX * VX_TILE_HY);
        const int gx = tx0 + hx - 1, gy = ty0 + hy - 1, gz = tz0 + hz - 1;
        const int edge = (hx == 0 || hx == VX_TILE_HX - 1) + (hy == 0 || hy == VX_TILE_HY - 1) + (hz == 0 || hz == VX_TILE_HZ - 1);
        if (edge <= 1 && gx >= 0 && gy >= 0 && gz >= 0 && gx < f.nx && gy < f.ny && gz < f.nz) {
            const int v = vbase + (gz * f.ny + gy) * f.nx + gx;
            sp0[c] = ld4(f.c_pose0 + v); sp1[c] = ld4(f.c_pose1 + v);
        }
    }
    // link records of this thread's (up to two) evaluations: issue the loads before the barrier
    TileEval ev[VX_TILE_ROUNDS];
    double2 ra[VX_TILE_ROUNDS], rb[VX_TILE_ROUNDS], rc[VX_TILE_ROUNDS]; float4 rs[VX_TILE_ROUNDS];
#pragma unroll
    for (int r = 0; r < VX_TILE_ROUNDS; r++) {
        ev[r] = tile_decode(f, threadIdx.x + r * VX_TILE_THREADS, tx0, ty0, tz0, vbase);
        if (ev[r].ok) {
            ra[r] = __ldg(f.c_rec[ev[r].axis][0] + ev[r].vn); rb[r] = __ldg(f.c_rec[ev[r].axis][1] + ev[r].vn);
            rc[r] = __ldg(f.c_rec[ev[r].axis][2] + ev[r].vn); rs[r] = __ldg(f.c_recf[ev[r].axis] + ev[r].vn);
        }
    }
    __syncthreads();

    // ---- phase 1: one thread per link evaluation
#pragma unroll
    for (int r = 0; r < VX_TILE_ROUNDS; r++) {
        if (!ev[r].ok) continue;
        const int axis = ev[r].axis, lx = ev[r].lx, ly = ev[r].ly, lz = ev[r].lz, vn = ev[r].vn;
        const int cn = ((lz + 1) * VX_TILE_HY + (ly + 1)) * VX_TILE_HX + (lx + 1);
        const int cp = cn + (axis == 0 ? 1 : (axis == 1 ? VX_TILE_HX : VX_TILE_HX * VX_TILE_HY));
        const double4 n0 = sp0[cn], n1 = sp1[cn];
        const uint32_t nbits = meta_hi(n1.w);
        if (!((nbits >> (VM_LINK_SHIFT + 2 * axis)) & 1u)) continue;        // no +axis link at this voxel
        const double4 p0 = sp0[cp], p1 = sp1[cp];

        const uint32_t hn = nbits, hp = meta_hi(p1.w);
        const DevVoxMat& vmn = UNI ? f.vm0 : f.vmat[hn & VM_MAT_MASK];
        const DevVoxMat& vmp = UNI ? f.vm0 : f.vmat[hp & VM_MAT_MASK];
        const DevLinkMat& lm = UNI ? f.lm0 : f.lmat[f.pair_lmat[(hn & VM_MAT_MASK) * f.n_mat + (hp & VM_MAT_MASK)]];
        LinkState st;
        lat_decode(ra[r], rb[r], rc[r], rs[r], (nbits >> (VM_LFLAG_SHIFT + 2 * axis)) & 3u, st);
        double rest = 0.5 * (vmn.size[axis] * (1 + meta_temp(n1.w) * vmn.cte) + vmp.size[axis] * (1 + meta_temp(p1.w) * vmp.cte));
        float t_area = 0.5f * (vmn.nom_f * vmn.nom_f + vmp.nom_f * vmp.nom_f);
        float damp_n = vmn.two_sqrtm_zeta / prev_dt, damp_p = vmp.two_sqrtm_zeta / prev_dt;
        q4 on, op;
        on.w = n0.w; on.x = n1.x; on.y = n1.y; on.z = n1.z;
        op.w = p0.w; op.x = p1.x; op.y = p1.y; op.z = p1.z;
        d3 fN, mN, fP, mP;
        link_forces(axis, mk3(n0.x, n0.y, n0.z), on, mk3(p0.x, p0.y, p0.z), op, rest, t_area, 0.0f,
                    damp_n, damp_p, lm, f.curve_e, f.curve_s, st, fN, mN, fP, mP);

        const bool n_in = lx >= 0 && ly >= 0 && lz >= 0;                     // negative end inside the brick: owner
        const int px = lx + (axis == 0), py = ly + (axis == 1), pz = lz + (axis == 2);
        const bool p_in = px < VX_TILE_X && py < VX_TILE_Y && pz < VX_TILE_Z;
        if (n_in) {
            double2 wa, wb, wc; float4 ws; uint32_t lf;
            lat_encode(st, wa, wb, wc, ws, lf);
            f.n_rec[axis][0][vn] = wa; f.n_rec[axis][1][vn] = wb; f.n_rec[axis][2][vn] = wc; f.n_recf[axis][vn] = ws;
            if (st.strain > 100) p->div_flag[parity] = 1;                    // src/Voxelyze.cpp:265
            const int li = (lz * VX_TILE_Y + ly) * VX_TILE_X + lx;
            double (*s)[VX_TILE_VOX] = slot[2 * axis];
            s[0][li] = fN.x; s[1][li] = fN.y; s[2][li] = fN.z; s[3][li] = mN.x; s[4][li] = mN.y; s[5][li] = mN.z;
            lflag_sh[axis][li] = (unsigned char)lf;
        }
        if (p_in) {
            const int li = (pz * VX_TILE_Y + py) * VX_TILE_X + px;
            double (*s)[VX_TILE_VOX] = slot[2 * axis + 1];
            s[0][li] = fP.x; s[1][li] = fP.y; s[2][li] = fP.z; s[3][li] = mP.x; s[4][li] = mP.y; s[5][li] = mP.z;
        }
    }

    // ---- phase 2: one thread per voxel (momenta requested before the barrier to overlap their latency)
    const int li = threadIdx.x;
    const int vlx = li % VX_TILE_X, vly = (li / VX_TILE_X) % VX_TILE_Y, vlz = li / (VX_TILE_X * VX_TILE_Y);
    const int gx = tx0 + vlx, gy = ty0 + vly, gz = tz0 + vlz;
    const bool voxel_thread = li < VX_TILE_VOX && gx < f.nx && gy < f.ny && gz < f.nz;
    const int v = vbase + (gz * f.ny + gy) * f.nx + gx;
    double4 m0 = make_double4(0.0, 0.0, 0.0, 0.0); double2 m1 = make_double2(0.0, 0.0);
    if (voxel_thread) { m0 = ld4(f.c_mom0 + v); m1 = __ldg(f.c_mom1 + v); }
    __syncthreads();
    if (!voxel_thread) return;
    const int cs = ((vlz + 1) * VX_TILE_HY + (vly + 1)) * VX_TILE_HX + (vlx + 1);
    const double4 s0 = sp0[cs], s1 = sp1[cs];
    VoxelState vs;
    vs.bits = meta_hi(s1.w);
    vs.temp = meta_temp(s1.w);
    const uint32_t mask = (vs.bits >> VM_LINK_SHIFT) & 0x3Fu;
    d3 F = mk3(0.0, 0.0, 0.0), M = mk3(0.0, 0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 6; k++) {                                            // reference summation order
        if (mask & (1u << k)) {
            F = F + mk3(slot[k][0][li], slot[k][1][li], slot[k][2][li]);
            M = M + mk3(slot[k][3][li], slot[k][4][li], slot[k][5][li]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
        if (mask & (1u << (2 * a)))
            vs.bits = (vs.bits & ~(3u << (VM_LFLAG_SHIFT + 2 * a))) | ((uint32_t)lflag_sh[a][li] << (VM_LFLAG_SHIFT + 2 * a));
    vs.pos = mk3(s0.x, s0.y, s0.z);
    vs.orient.w = s0.w; vs.orient.x = s1.x; vs.orient.y = s1.y; vs.orient.z = s1.z;
    vs.lin = mk3(m0.x, m0.y, m0.z);
    vs.ang = mk3(m0.w, m1.x, m1.y);
    if (!(vs.bits & VM_GHOST)) {
        const DevVoxMat& vm = UNI ? f.vm0 : f.vmat[vs.bits & VM_MAT_MASK];
        const DevExt* ext = (vs.bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
        voxel_integrate(vs, F, M, nullptr, 0, nullptr, vm, ext, dt, floor_on != 0);
    }
    f.n_pose0[v] = make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w);
    f.n_pose1[v] = make_double4(vs.orient.x, vs.orient.y, vs.orient.z, meta_pack(vs.temp, vs.bits));
    f.n_mom0[v] = make_double4(vs.lin.x, vs.lin.y, vs.lin.z, vs.ang.x);
    f.n_mom1[v] = make_double2(vs.ang.y, vs.ang.z);
}


// =================================================================================================
// k_lattice_zmarch -- the warp-brick step, marching.  One warp owns a 4 x 4 column of the lattice and walks
// up a chunk of it two bricks (A below, B above; 4 planes) at a time.
//   * The force a brick's top layer exerts on the layer above is computed once, by the brick below, and
//     carried over in shared memory (zslot): no -Z entering links are re-evaluated (only once per chunk).
//   * The -X / -Y entering links of A and of B (2 x 16) fill ONE round H of 32 lanes.
//   -> 7 rounds of link evaluations per 64 voxels = 3.5 per voxel (warp bricks: 4.0), all lanes busy.
//   * Everything is requested with cp.async one round before it is read: the payload of round k+1 (link
//     records or momenta) goes into the window round k-1 has left; poses of brick B, of the next A and
//     of the next B's -X/-Y faces go into regions of the pose table as they fall free.  Only the 32 outside
//     poses of round H are waited for (once per 64 voxels).
// Shared memory per warp: 2 windows x 2 KB + 96 poses x 64 B + hslot 1.5 KB + zslot 2 x 768 B = 13 312 B.
// Pose table regions: OA 0..31 own poses of A | HB 32..63 outside poses of round H, then own poses of B |
//                     EB 64..79 B's x=0 / y=0 voxels for round H, then the poses beyond B's +X/+Y faces |
//                     EA 80..95 the poses beyond A's +X/+Y faces.
// =================================================================================================
#define VX_ZM_WARPS 8
#define VX_ZM_POSES 96
#ifndef VX_ZM_PAIRS
#define VX_ZM_PAIRS 8                                   // brick pairs per warp: 32 planes
#endif
#define VX_ZM_WARP_BYTES (2 * 4 * 32 * 16 + 4 * VX_ZM_POSES * 16 + 6 * 32 * 8 + 2 * 6 * 16 * 8)
#define VX_ZM_SMEM (VX_ZM_WARPS * VX_ZM_WARP_BYTES)

template <bool UNI>
__global__ void __launch_bounds__(32 * VX_ZM_WARPS, 2)
k_lattice_zmarch(LatFrame f, int parity, int first_of_call, int floor_on, int ncx, int ncy, int ncz)
{
    extern __shared__ __align__(16) unsigned char wb_smem[];
    DevParams* p = f.params;
    const int frozen = p->div_flag[parity ^ 1] | p->div_latched;
    const float dt = p->dt;
    const float prev_dt = first_of_call ? p->prev_dt : dt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (frozen) p->div_latched = 1;
        else if (p->pending) { p->steps_done += 1; p->time += dt; }
        if (!frozen) p->pending = 1;
    }
    if (frozen) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = wb_smem + (size_t)warp * VX_ZM_WARP_BYTES;
    uint4 (*win)[4][32] = reinterpret_cast<uint4 (*)[4][32]>(wbase);                                      // [window][part][lane]
    uint4 (*pose_sh)[VX_ZM_POSES] = reinterpret_cast<uint4 (*)[VX_ZM_POSES]>(wbase + 4096);               // [part][entry]
    double (*hslot)[32] = reinterpret_cast<double (*)[32]>(wbase + 4096 + 4 * VX_ZM_POSES * 16);          // [comp][entering link]
    double (*zslot)[6][16] = reinterpret_cast<double (*)[6][16]>(wbase + 4096 + 4 * VX_ZM_POSES * 16 + 1536);   // [ping][comp][x + 4 y]
    constexpr int OA = 0, HB = 32, EB = 64, EA = 80;

    int b = blockIdx.x * VX_ZM_WARPS + warp;
    const int cx = b % ncx; b /= ncx;
    const int cy = b % ncy; b /= ncy;
    const int cz = b % ncz; const int member = b / ncz;
    const int x0 = cx * VX_WB_X, y0 = cy * VX_WB_Y;
    const int z_begin = cz * (4 * VX_ZM_PAIRS), z_end = min(f.nz, z_begin + 4 * VX_ZM_PAIRS);
    if ((size_t)member * f.nz * f.nxy >= (size_t)f.n_vox) return;
    const int vbase = member * f.nz * f.nxy;

    const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
    const int x = x0 + lx, y = y0 + ly;
    const bool xy_ok = x < f.nx && y < f.ny;
    const int vxy = vbase + min(y, f.ny - 1) * f.nx + min(x, f.nx - 1);
    auto vox = [&](int z) { return vxy + min(z, f.nz - 1) * f.nxy; };                // clamped: idle lanes address valid memory

    auto c_rec = [&](int a, int k) { return f.c_rec[0][0] + (size_t)(a * 3 + k) * f.n_vox; };
    auto c_recf = [&](int a) { return f.c_recf[0] + (size_t)a * f.n_vox; };
    auto request_pose = [&](int entry, int v) {
        cp_async16(&pose_sh[0][entry], reinterpret_cast<const uint4*>(f.c_pose0 + v));
        cp_async16(&pose_sh[1][entry], reinterpret_cast<const uint4*>(f.c_pose0 + v) + 1);
        cp_async16(&pose_sh[2][entry], reinterpret_cast<const uint4*>(f.c_pose1 + v));
        cp_async16(&pose_sh[3][entry], reinterpret_cast<const uint4*>(f.c_pose1 + v) + 1);
    };
    auto load_pose = [&](int entry, double4& a, double4& c) {
        const uint4 e0 = pose_sh[0][entry], e1 = pose_sh[1][entry], e2 = pose_sh[2][entry], e3 = pose_sh[3][entry];
        a = make_double4(__hiloint2double(e0.y, e0.x), __hiloint2double(e0.w, e0.z), __hiloint2double(e1.y, e1.x), __hiloint2double(e1.w, e1.z));
        c = make_double4(__hiloint2double(e2.y, e2.x), __hiloint2double(e2.w, e2.z), __hiloint2double(e3.y, e3.x), __hiloint2double(e3.w, e3.z));
    };
    auto request_rec = [&](int w, int a, int owner) {
#pragma unroll
        for (int k = 0; k < 3; k++) cp_async16(&win[w][k][lane], c_rec(a, k) + owner);
        cp_async16(&win[w][3][lane], c_recf(a) + owner);
    };
    // own poses of the brick at z0 -> region `own`; the poses beyond its +X / +Y faces -> region `ext`
    auto request_own = [&](int own, int z0) { if (z0 < f.nz) request_pose(own + lane, vox(z0 + lz)); };
    auto request_ext = [&](int ext, int z0) {
        if (!xy_ok || z0 + lz >= f.nz) return;
        if (lx == VX_WB_X - 1 && x + 1 < f.nx) request_pose(ext + ly + 4 * lz, vox(z0 + lz) + 1);
        if (ly == VX_WB_Y - 1 && y + 1 < f.ny) request_pose(ext + 8 + lx + 4 * lz, vox(z0 + lz) + f.nx);
    };
    // round H: lanes 0..15 serve brick A, 16..31 brick B; q < 8: the link through -X into voxel (0, q&3, q>>2), else through -Y into ((q-8)&3, 0, (q-8)>>2)
    const int hq = lane & 15, hb = lane >> 4;
    const int h_axis = hq < 8 ? 0 : 1;
    const int h_tl = hq < 8 ? ((hq >> 2) << 4) | ((hq & 3) << 2) : (((hq - 8) >> 2) << 4) | ((hq - 8) & 3);
    const bool h_xy = x0 + (h_tl & 3) < f.nx && y0 + ((h_tl >> 2) & 3) < f.ny && (h_axis == 0 ? x0 : y0) > 0;
    const int h_vxy = vbase + (y0 + ((h_tl >> 2) & 3)) * f.nx + x0 + (h_tl & 3) - (h_axis == 0 ? 1 : f.nx);   // negative-end voxel, plane 0
    // B's x = 0 / y = 0 voxels (targets of B's entering links) -> EB 0..15; requested by lanes 16..31 for themselves
    auto request_h_targets = [&](int zB) {
        if (hb == 1 && h_xy && zB + (h_tl >> 4) < f.nz) request_pose(EB + hq, h_vxy + (h_axis == 0 ? 1 : f.nx) + (zB + (h_tl >> 4)) * f.nxy);
    };
    auto request_h = [&](int w, int zp) {                              // outside poses -> HB, records -> window w
        const int z = zp + 2 * hb + (h_tl >> 4);
        if (h_xy && z < f.nz) { request_pose(HB + lane, h_vxy + z * f.nxy); request_rec(w, h_axis, h_vxy + z * f.nxy); }
    };
    auto request_h_rec = [&](int w, int zp) {
        const int z = zp + 2 * hb + (h_tl >> 4);
        if (h_xy && z < f.nz) request_rec(w, h_axis, h_vxy + z * f.nxy);
    };

    // ---- chunk prologue
    int r = 0;                                   // round counter: round r reads window r & 1
    int kb = 0;                                  // bricks done: brick kb takes its -Z forces from zslot[kb & 1], leaves its +Z forces in zslot[(kb + 1) & 1]
    int zp = z_begin;
    int mode = z_begin > 0 ? 0 : 1;              // 0: the 16 links entering the chunk from below still have to be evaluated
    request_own(OA, zp); request_ext(EA, zp); request_h_targets(zp + 2);
    if (mode == 0) {
        if (lane < 16 && xy_ok) { request_pose(HB + lane, vox(zp) - f.nxy); request_rec(0, 2, vox(zp) - f.nxy); }   // lane = (lx, ly) of the bottom layer
    } else request_h_rec(0, zp);
    cp_async_commit();

#pragma unroll 1
    while (zp < z_end) {
        // ================= entering links: round ZH (chunk start only) or round H of this pair
        if (mode == 0) request_h_rec(1, zp);
        else {
            const int z = zp + 2 * hb + (h_tl >> 4);
            if (h_xy && z < f.nz) request_pose(HB + lane, h_vxy + z * f.nxy);
            if (xy_ok && zp + lz < f.nz && x + 1 < f.nx) request_rec((r + 1) & 1, 0, vox(zp + lz));      // payload of A's round 0
        }
        cp_async_commit();
        if (mode == 0) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        {
            int axis, tgt_entry; bool act; double* dst; int dst_stride;
            if (mode == 0) {
                axis = 2; tgt_entry = OA + lane; act = lane < 16 && xy_ok;
                dst = &zslot[0][0][lane & 15]; dst_stride = 16;
            } else {
                axis = h_axis; tgt_entry = hb ? EB + hq : OA + h_tl; act = h_xy && zp + 2 * hb + (h_tl >> 4) < f.nz;
                dst = &hslot[0][lane]; dst_stride = 32;
            }
            if (act && ((pose_sh[3][tgt_entry].w >> (VM_LINK_SHIFT + 2 * axis + 1)) & 1u)) {
                double4 n0, n1, p0, p1;
                load_pose(HB + lane, n0, n1);
                load_pose(tgt_entry, p0, p1);
                const uint4 r0 = win[r & 1][0][lane], r1 = win[r & 1][1][lane], r2 = win[r & 1][2][lane], r3 = win[r & 1][3][lane];
                LinkState st; d3 fN, mN, fP, mP;
                lat_eval_link_rec<UNI>(f, axis, meta_hi(n1.w),
                                       make_double2(__hiloint2double(r0.y, r0.x), __hiloint2double(r0.w, r0.z)),
                                       make_double2(__hiloint2double(r1.y, r1.x), __hiloint2double(r1.w, r1.z)),
                                       make_double2(__hiloint2double(r2.y, r2.x), __hiloint2double(r2.w, r2.z)),
                                       make_float4(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w)),
                                       n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP);
                dst[0] = fP.x; dst[dst_stride] = fP.y; dst[2 * dst_stride] = fP.z; dst[3 * dst_stride] = mP.x; dst[4 * dst_stride] = mP.y; dst[5 * dst_stride] = mP.z;
            }
        }
        __syncwarp();
        r++;
        if (mode == 0) { mode = 1; continue; }

        // ================= bricks A (h = 0) and B (h = 1): rounds 0..2 = own +X/+Y/+Z links, round 3 = voxels
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int z0 = zp + 2 * h;
            const int own = h ? HB : OA, ext = h ? EB : EA;
            const int z = z0 + lz;
            const bool has_voxel = xy_ok && z < f.nz;
            const int v = vox(z);
            uint32_t bits = 0, mask = 0, new_bits = 0;
            d3 F = mk3(0.0, 0.0, 0.0), M = mk3(0.0, 0.0, 0.0);
#pragma unroll 1
            for (int a = 0; a < 4; a++) {
                // ---- requests: the payload of the next round, and poses into regions that have just fallen free
                const int wn = (r + 1) & 1;
                if (a < 2) { if (has_voxel && (a == 0 ? y + 1 < f.ny : z + 1 < f.nz)) request_rec(wn, a + 1, v); }
                else if (a == 2) {
                    if (has_voxel) {
                        cp_async16(&win[wn][0][lane], reinterpret_cast<const uint4*>(f.c_mom0 + v));
                        cp_async16(&win[wn][1][lane], reinterpret_cast<const uint4*>(f.c_mom0 + v) + 1);
                        cp_async16(&win[wn][2][lane], reinterpret_cast<const uint4*>(f.c_mom1 + v));
                    }
                } else if (h == 0) { if (xy_ok && z + 2 < f.nz && x + 1 < f.nx) request_rec(wn, 0, vox(z + 2)); }     // B's round 0
                else if (zp + 4 < z_end) request_h_rec(wn, zp + 4);                                              // the next pair's round H
                if (h == 0 && a == 0) { request_own(HB, zp + 2); request_ext(EB, zp + 2); }                    // round H has left HB and EB
                if (h == 1 && a == 0 && zp + 4 < f.nz) { request_own(OA, zp + 4); request_ext(EA, zp + 4); }   // A is done with OA and EA
                if (h == 1 && a == 2 && zp + 4 < z_end) request_h_targets(zp + 6);                               // B's round 1 was the last reader of EB
                cp_async_commit();
                cp_async_wait<1>();
                __syncwarp();

                if (a == 0) {
                    bits = pose_sh[3][own + lane].w;
                    mask = has_voxel ? ((bits >> VM_LINK_SHIFT) & 0x3Fu) : 0u;
                    new_bits = bits;
                }
                if (a < 3) {
                    const bool inside = a == 0 ? lx < VX_WB_X - 1 : (a == 1 ? ly < VX_WB_Y - 1 : lz < VX_WB_Z - 1);
                    const bool first = a == 0 ? lx == 0 : (a == 1 ? ly == 0 : lz == 0);
                    const int dl = a == 0 ? 1 : (a == 1 ? 4 : 16);
                    d3 fN = mk3(0.0, 0.0, 0.0), mN = fN, fP = fN, mP = fN;
                    if ((mask >> (2 * a)) & 1u) {
                        // partner: inside the brick, beyond its +X/+Y face, or (a == 2, top layer) the bottom layer of the brick above
                        const int pe = inside ? own + lane + dl : (a == 0 ? ext + ly + 4 * lz : (a == 1 ? ext + 8 + lx + 4 * lz : (h ? OA : HB) + lane - 16));
                        double4 n0, n1, p0, p1;
                        load_pose(own + lane, n0, n1);
                        load_pose(pe, p0, p1);
                        const uint4 r0 = win[r & 1][0][lane], r1 = win[r & 1][1][lane], r2 = win[r & 1][2][lane], r3 = win[r & 1][3][lane];
                        LinkState st;
                        lat_eval_link_rec<UNI>(f, a, bits,
                                               make_double2(__hiloint2double(r0.y, r0.x), __hiloint2double(r0.w, r0.z)),
                                               make_double2(__hiloint2double(r1.y, r1.x), __hiloint2double(r1.w, r1.z)),
                                               make_double2(__hiloint2double(r2.y, r2.x), __hiloint2double(r2.w, r2.z)),
                                               make_float4(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w)),
                                               n0, n1, p0, p1, prev_dt, st, fN, mN, fP, mP);
                        double2 wa, wb, wc; float4 ws; uint32_t lf;
                        lat_encode(st, wa, wb, wc, ws, lf);
                        double2* nr = f.n_rec[0][0] + (size_t)(a * 3) * f.n_vox + v;
                        nr[0] = wa; nr[f.n_vox] = wb; nr[2 * (size_t)f.n_vox] = wc; (f.n_recf[0] + (size_t)a * f.n_vox)[v] = ws;
                        new_bits = (new_bits & ~(3u << (VM_LFLAG_SHIFT + 2 * a))) | (lf << (VM_LFLAG_SHIFT + 2 * a));
                        if (st.strain > 100) p->div_flag[parity] = 1;          // src/Voxelyze.cpp:265
                        F = F + fN; M = M + mN;
                        if (a == 2 && lz == 1) {                               // carried to the brick above
                            double (*zs)[16] = zslot[(kb + 1) & 1];
                            const int q = lane - 16;
                            zs[0][q] = fP.x; zs[1][q] = fP.y; zs[2][q] = fP.z; zs[3][q] = mP.x; zs[4][q] = mP.y; zs[5][q] = mP.z;
                        }
                    }
                    const int src = (lane - dl) & 31;
                    d3 inF = mk3(__shfl_sync(0xffffffffu, fP.x, src), __shfl_sync(0xffffffffu, fP.y, src), __shfl_sync(0xffffffffu, fP.z, src));
                    d3 inM = mk3(__shfl_sync(0xffffffffu, mP.x, src), __shfl_sync(0xffffffffu, mP.y, src), __shfl_sync(0xffffffffu, mP.z, src));
                    if ((mask >> (2 * a + 1)) & 1u) {
                        if (first) {
                            if (a < 2) {
                                const int hl = 16 * h + (a == 0 ? ly + 4 * lz : 8 + lx + 4 * lz);
                                inF = mk3(hslot[0][hl], hslot[1][hl], hslot[2][hl]);
                                inM = mk3(hslot[3][hl], hslot[4][hl], hslot[5][hl]);
                            } else {
                                double (*zs)[16] = zslot[kb & 1];
                                inF = mk3(zs[0][lane], zs[1][lane], zs[2][lane]);
                                inM = mk3(zs[3][lane], zs[4][lane], zs[5][lane]);
                            }
                        }
                        F = F + inF; M = M + inM;
                    }
                } else if (has_voxel) {
                    // ---- round 3: one lane per voxel
                    const uint4 q0 = win[r & 1][0][lane], q1 = win[r & 1][1][lane], q2 = win[r & 1][2][lane];
                    double4 s0, s1;
                    load_pose(own + lane, s0, s1);
                    VoxelState vs;
                    vs.bits = new_bits; vs.temp = meta_temp(s1.w);
                    if (vs.bits & VM_GHOST) reinterpret_cast<uint32_t*>(&f.n_pose1[v].w)[1] = vs.bits;      // see k_lattice_warp
                    else {
                        vs.pos = mk3(s0.x, s0.y, s0.z);
                        vs.orient.w = s0.w; vs.orient.x = s1.x; vs.orient.y = s1.y; vs.orient.z = s1.z;
                        vs.lin = mk3(__hiloint2double(q0.y, q0.x), __hiloint2double(q0.w, q0.z), __hiloint2double(q1.y, q1.x));
                        vs.ang = mk3(__hiloint2double(q1.w, q1.z), __hiloint2double(q2.y, q2.x), __hiloint2double(q2.w, q2.z));
                        const DevVoxMat& vm = UNI ? f.vm0 : f.vmat[vs.bits & VM_MAT_MASK];
                        const DevExt* ext_row = (vs.bits & VM_HAS_EXT) ? f.ext + f.ext_idx[v] : nullptr;
                        voxel_integrate(vs, F, M, nullptr, 0, nullptr, vm, ext_row, dt, floor_on != 0);
                        f.n_pose0[v] = make_double4(vs.pos.x, vs.pos.y, vs.pos.z, vs.orient.w);
                        f.n_pose1[v] = make_double4(vs.orient.x, vs.orient.y, vs.orient.z, meta_pack(vs.temp, vs.bits));
                        f.n_mom0[v] = make_double4(vs.lin.x, vs.lin.y, vs.lin.z, vs.ang.x);
                        f.n_mom1[v] = make_double2(vs.ang.y, vs.ang.z);
                    }
                }
                __syncwarp();
                r++;
            }
            kb++;
        }
        zp += 4;
    }
    cp_async_wait<0>();
}


} // namespace vxd

