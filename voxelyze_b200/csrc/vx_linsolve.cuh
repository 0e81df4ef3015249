// vx_linsolve.cuh -- the static (linearised) solve of CVX_LinearSolver (src/VX_LinearSolver.cpp) on the device.
//
// The reference assembles the upper triangle of the 6N x 6N direct-stiffness matrix in CSR form (calculateA, :116-233),
// eliminates the fixed degrees of freedom (applyBX, :273-328), hands the system to PARDISO and writes the displacements
// back as poses (postResults, :336-347).  Here the matrix is never formed: every entry of a voxel's six rows follows from
// the beam constants a1, a2, b1, b2, b3 of its (at most six) links, so K*u is a 7-point gather over the voxel graph, and
// the system is solved by a Jacobi-preconditioned conjugate-gradient iteration in FP64:
//
//   k_lin_step_a   p = z + beta*p (recomputed for the six neighbours instead of a pass of its own), y = P K p, <p,y>
//   k_lin_step_b   x += alpha p, r -= alpha y, z = r / diag, <r,z>, <r,r>
//
// two launches per iteration, ~0.6 kB of HBM traffic per voxel per iteration, scalars (alpha, beta, convergence) never leave
// the device: every block reduces the previous kernel's per-block partial sums in the same fixed order, so all blocks take
// the same decision and two runs give the same bits.  P projects out the prescribed degrees of freedom; their values
// enter through the first residual r0 = P (f - K x0).
//
// Row/column conventions of the reference that are kept (they are visible in the result):
//  * of the two voxels of a link the one with the LOWER voxelsList index plays "voxel 1" of the element matrix, whichever
//    end of the link it is (VX_LinearSolver.cpp:173: "swap to keep i1 lower than i2") -- the sign of the b2 coupling terms
//    follows that role, not the geometry;
//  * matrix entries are the float beam constants, added up in double (addAValue takes a float, :236).
#pragma once
#include "vx_physics.cuh"

namespace vxd {

constexpr int VX_LIN_TPB = 256;
constexpr int VX_LIN_MAX_GRID = 148 * 8;

struct LinScalars {
    double bb;          // |r0|^2, the reference of the relative residual
    double rr;          // |r|^2 at the last check
    double tol2;        // rel_tol^2
    int done;           // 0 running, 1 converged, 2 breakdown (<p,Kp> <= 0: the system is singular or not positive definite)
    int iters;          // iterations completed when `done` was set
};

struct LinFrame {
    int n, grid;                          // caller voxels; blocks of every launch (partials per sum)
    const int* nbr;                       // [6][n] neighbour (caller index) through link slot s (0 X+ .. 5 Z-), -1: no link
    const uint16_t* mat;                  // [n] voxel material
    const uint16_t* pair_lmat; int n_mat; // link material of a pair of voxel materials
    const DevLinkMat* lmat;
    unsigned char* fixed;                 // [n] prescribed dof bits (CVX_External::isFixed, plus free dofs nothing holds)
    const int* ijk;                       // [n][3] lattice indices (originalPosition, include/VX_Voxel.h:80)
    const int* e2i;                       // caller -> internal voxel index
    double voxel_size;
    // vectors: 3 double2 parts per voxel, part k of voxel v at [k*n + v]  (dofs x y | z rx | ry rz)
    double2* x; double2* r; double2* z; double2* y; double2* minv; double2* p[2];
    double* part_pap;                     // [grid]
    double* part_rz[2];                   // [grid] per iteration parity
    double* part_rr[2];
    LinScalars* sc;
};

struct Dof6 { double v[6]; };

__device__ __forceinline__ Dof6 lin_load(const double2* __restrict__ a, int n, int i)
{
    Dof6 d; const double2 p0 = a[i], p1 = a[(size_t)n + i], p2 = a[2 * (size_t)n + i];
    d.v[0] = p0.x; d.v[1] = p0.y; d.v[2] = p1.x; d.v[3] = p1.y; d.v[4] = p2.x; d.v[5] = p2.y;
    return d;
}
__device__ __forceinline__ void lin_store(double2* __restrict__ a, int n, int i, const Dof6& d)
{
    a[i] = make_double2(d.v[0], d.v[1]); a[(size_t)n + i] = make_double2(d.v[2], d.v[3]); a[2 * (size_t)n + i] = make_double2(d.v[4], d.v[5]);
}

// the six rows of one link's element matrix that belong to voxel "me" times (u_me, u_nb); first: me is the link's "voxel 1"
// (VX_LinearSolver.cpp:178-227).  AX is the link axis.
template <int AX>
__device__ __forceinline__ void lin_link_rows(const DevLinkMat& m, bool first, const Dof6& um, const Dof6& un, Dof6& y)
{
    const double a1 = (double)m.a1, b3x2 = 2.0 * m.b3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const double d = (AX == j) ? a1 : m.b1;                       // :182-187
        y.v[j] += d * (um.v[j] - un.v[j]);
        const double dd = (AX == j) ? m.a2 : b3x2, dof = (AX == j) ? -m.a2 : m.b3;   // :189-195
        y.v[3 + j] += dd * um.v[3 + j] + dof * un.v[3 + j];
    }
    constexpr int R1 = AX == 0 ? 1 : 0, C1 = AX == 2 ? 4 : 5, R2 = AX == 2 ? 1 : 2, C2 = AX == 0 ? 4 : 3;   // :201-217
    const double val = (AX == 1 ? -m.b2 : m.b2) * (first ? 1.0 : -1.0);
    y.v[R1] += val * (um.v[C1] + un.v[C1]);                          // :219-222 and their transposes
    y.v[C1] += val * (um.v[R1] - un.v[R1]);
    y.v[R2] -= val * (um.v[C2] + un.v[C2]);                          // :224-227
    y.v[C2] -= val * (um.v[R2] - un.v[R2]);
}

__device__ __forceinline__ const DevLinkMat& lin_mat(const LinFrame& f, int ma, int mb) { return f.lmat[f.pair_lmat[ma * f.n_mat + mb]]; }

// diagonal of K at voxel v (the Jacobi preconditioner and the test for dofs nothing holds)
__device__ __forceinline__ Dof6 lin_diag(const LinFrame& f, int v)
{
    Dof6 d; for (int j = 0; j < 6; j++) d.v[j] = 0.0;
    const int mv = f.mat[v];
    for (int s = 0; s < 6; s++) {
        const int nb = f.nbr[(size_t)s * f.n + v];
        if (nb < 0) continue;
        const DevLinkMat& m = lin_mat(f, mv, f.mat[nb]);
        const int ax = s >> 1;
        for (int j = 0; j < 3; j++) {
            d.v[j] += (ax == j) ? (double)m.a1 : m.b1;
            d.v[3 + j] += (ax == j) ? m.a2 : 2.0 * m.b3;
        }
    }
    return d;
}

// sums of up to three arrays of per-block partials, same order in every block; result valid in all threads
template <int NA>
__device__ __forceinline__ void lin_totals(const double* const (&arr)[NA], int count, double (&out)[NA])
{
    __shared__ double sh[NA][VX_LIN_TPB / 32];
    double acc[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) {
        acc[a] = 0.0;
        for (int i = threadIdx.x; i < count; i += VX_LIN_TPB) acc[a] += arr[a][i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
        if ((threadIdx.x & 31) == 0) sh[a][threadIdx.x >> 5] = acc[a];
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NA; a++) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < VX_LIN_TPB / 32; w++) t += sh[a][w];
        out[a] = t;
    }
    __syncthreads();
}

// per-block sum of NA per-thread values -> part[a][blockIdx.x]
template <int NA>
__device__ __forceinline__ void lin_block_sums(double (&acc)[NA], double* const (&part)[NA])
{
    __shared__ double sh[NA][VX_LIN_TPB / 32];
#pragma unroll
    for (int a = 0; a < NA; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
        if ((threadIdx.x & 31) == 0) sh[a][threadIdx.x >> 5] = acc[a];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int a = 0; a < NA; a++) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < VX_LIN_TPB / 32; w++) t += sh[a][w];
            part[a][blockIdx.x] = t;
        }
    }
}

// rows of voxel v of P K u, u given by a functor (so that step A can form p on the fly)
template <typename U>
__device__ __forceinline__ Dof6 lin_apply_rows(const LinFrame& f, int v, const Dof6& um, U&& u_of)
{
    Dof6 y; for (int j = 0; j < 6; j++) y.v[j] = 0.0;
    const int mv = f.mat[v];
    int nbs[6];                                                      // all six indices, then all six gathers, before any of them is used:
#pragma unroll                                                       // the iteration is latency-bound otherwise (4.5 against 3.7 TB/s at 6 M voxels)
    for (int s = 0; s < 6; s++) nbs[s] = f.nbr[(size_t)s * f.n + v];
#pragma unroll
    for (int s = 0; s < 6; s++) {
        const int nb = nbs[s];
        const int j = nb < 0 ? v : nb;                               // no link: the loads go to the voxel itself (cached), nothing is added
        const DevLinkMat& m = lin_mat(f, mv, f.mat[j]);
        const Dof6 un = u_of(j);
        const bool first = v < nb;
        if (nb >= 0) {
            if ((s >> 1) == 0) lin_link_rows<0>(m, first, um, un, y);
            else if ((s >> 1) == 1) lin_link_rows<1>(m, first, um, un, y);
            else lin_link_rows<2>(m, first, um, un, y);
        }
    }
    const unsigned fx = f.fixed[v];
#pragma unroll
    for (int j = 0; j < 6; j++) if (fx & (1u << j)) y.v[j] = 0.0;
    return y;
}

// externals -> prescribed-dof bits and the load vector (applyBX :283-300: forces only on dofs that are not fixed)
__global__ void k_lin_externals(LinFrame f, int n_ext, const int* __restrict__ ext_vox, const DevExt* __restrict__ ext, double2* __restrict__ load)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ext) return;
    const int v = ext_vox[k];
    const DevExt& e = ext[k];
    const unsigned fx = e.dof & 0x3Fu;
    f.fixed[v] = (unsigned char)fx;
    Dof6 b;
    for (int j = 0; j < 3; j++) { b.v[j] = (fx & (1u << j)) ? 0.0 : (double)e.force[j]; b.v[3 + j] = (fx & (8u << j)) ? 0.0 : (double)e.moment[j]; }
    lin_store(load, f.n, v, b);
}

// x0 = the current displacement and rotation vector on the prescribed dofs (applyBX :288-296), 0 elsewhere; 1/diag
__global__ void k_lin_start(LinFrame f, const double4* __restrict__ pose0, const double4* __restrict__ pose1)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.n; v += gridDim.x * blockDim.x) {
        const Dof6 d = lin_diag(f, v);
        unsigned fx = f.fixed[v];
        Dof6 mi, x0;
        for (int j = 0; j < 6; j++) {
            if (!(d.v[j] > 0.0)) fx |= 1u << j;                       // nothing holds this dof: keep it where it is
            mi.v[j] = (fx & (1u << j)) ? 0.0 : 1.0 / d.v[j];
            x0.v[j] = 0.0;
        }
        f.fixed[v] = (unsigned char)fx;
        if (fx) {
            const int i = f.e2i[v];
            const double4 a = pose0[i], b = pose1[i];
            const double s = f.voxel_size;
            const double disp[3] = {a.x - f.ijk[3 * v] * s, a.y - f.ijk[3 * v + 1] * s, a.z - f.ijk[3 * v + 2] * s};
            d3 ang = mk3(0.0, 0.0, 0.0);
            if (a.w != 1.0) { q4 q; q.w = a.w; q.x = b.x; q.y = b.y; q.z = b.z; ang = q_to_rotvec(q); }
            const double an[3] = {ang.x, ang.y, ang.z};
            for (int j = 0; j < 3; j++) { if (fx & (1u << j)) x0.v[j] = disp[j]; if (fx & (8u << j)) x0.v[3 + j] = an[j]; }
        }
        lin_store(f.minv, f.n, v, mi);
        lin_store(f.x, f.n, v, x0);
    }
}

// r0 = P (f - K x0), z0 = r0/diag, <r0,z0>, <r0,r0>
__global__ void __launch_bounds__(VX_LIN_TPB) k_lin_residual0(LinFrame f, const double2* __restrict__ load)
{
    double acc[2] = {0.0, 0.0};
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.n; v += gridDim.x * blockDim.x) {
        const Dof6 um = lin_load(f.x, f.n, v);
        const Dof6 kx = lin_apply_rows(f, v, um, [&](int nb) { return lin_load(f.x, f.n, nb); });
        const Dof6 b = lin_load(load, f.n, v), mi = lin_load(f.minv, f.n, v);
        const unsigned fx = f.fixed[v];
        Dof6 r, z;
        for (int j = 0; j < 6; j++) {
            r.v[j] = (fx & (1u << j)) ? 0.0 : b.v[j] - kx.v[j];
            z.v[j] = mi.v[j] * r.v[j];
            acc[0] += r.v[j] * z.v[j]; acc[1] += r.v[j] * r.v[j];
        }
        lin_store(f.r, f.n, v, r);
        lin_store(f.z, f.n, v, z);
    }
    double* const part[2] = {f.part_rz[0], f.part_rr[0]};
    lin_block_sums<2>(acc, part);
}

__global__ void __launch_bounds__(VX_LIN_TPB) k_lin_begin(LinFrame f, double rel_tol)
{
    const double* const arr[1] = {f.part_rr[0]};
    double t[1];
    lin_totals<1>(arr, f.grid, t);
    if (threadIdx.x == 0) {
        f.sc->bb = t[0]; f.sc->rr = t[0]; f.sc->tol2 = rel_tol * rel_tol; f.sc->iters = 0;
        f.sc->done = t[0] == 0.0 ? 1 : 0;
    }
}

// convergence test on the residual left by iteration k - 1 (also the last thing a batch of iterations runs)
__global__ void __launch_bounds__(VX_LIN_TPB) k_lin_status(LinFrame f, int k)
{
    if (f.sc->done) return;
    const double* const arr[1] = {f.part_rr[k & 1]};
    double t[1];
    lin_totals<1>(arr, f.grid, t);
    if (threadIdx.x == 0) {
        f.sc->rr = t[0]; f.sc->iters = k;
        if (t[0] <= f.sc->tol2 * f.sc->bb) f.sc->done = 1;
    }
}

// iteration k, first half
__global__ void __launch_bounds__(VX_LIN_TPB, 4) k_lin_step_a(LinFrame f, int k)
{
    if (f.sc->done) return;
    const double* const arr[3] = {f.part_rr[k & 1], f.part_rz[k & 1], f.part_rz[(k + 1) & 1]};
    double t[3];
    lin_totals<3>(arr, f.grid, t);
    if (t[0] <= f.sc->tol2 * f.sc->bb) {                              // every block sees the same sums: same decision
        if (blockIdx.x == 0 && threadIdx.x == 0) { f.sc->rr = t[0]; f.sc->iters = k; f.sc->done = 1; }
        return;
    }
    const bool restart = k == 0;
    const double beta = restart ? 0.0 : t[1] / t[2];
    const double2* __restrict__ po = f.p[(k + 1) & 1];
    double2* __restrict__ pn = f.p[k & 1];
    auto p_of = [&](int i) {
        Dof6 z = lin_load(f.z, f.n, i);
        if (!restart) { const Dof6 o = lin_load(po, f.n, i); for (int j = 0; j < 6; j++) z.v[j] += beta * o.v[j]; }
        return z;
    };
    double acc[1] = {0.0};
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.n; v += gridDim.x * blockDim.x) {
        const Dof6 pm = p_of(v);
        const Dof6 y = lin_apply_rows(f, v, pm, p_of);
        lin_store(pn, f.n, v, pm);
        lin_store(f.y, f.n, v, y);
        for (int j = 0; j < 6; j++) acc[0] += pm.v[j] * y.v[j];
    }
    double* const part[1] = {f.part_pap};
    lin_block_sums<1>(acc, part);
}

// iteration k, second half
__global__ void __launch_bounds__(VX_LIN_TPB) k_lin_step_b(LinFrame f, int k)
{
    if (f.sc->done) return;
    const double* const arr[2] = {f.part_rz[k & 1], f.part_pap};
    double t[2];
    lin_totals<2>(arr, f.grid, t);
    if (!(t[1] > 0.0)) {                                              // <p,Kp> <= 0 or NaN: no descent direction left
        if (blockIdx.x == 0 && threadIdx.x == 0) { f.sc->iters = k; f.sc->done = 2; }
        return;
    }
    const double alpha = t[0] / t[1];
    const double2* __restrict__ p = f.p[k & 1];
    double acc[2] = {0.0, 0.0};
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.n; v += gridDim.x * blockDim.x) {
        Dof6 x = lin_load(f.x, f.n, v), r = lin_load(f.r, f.n, v), z;
        const Dof6 pv = lin_load(p, f.n, v), y = lin_load(f.y, f.n, v), mi = lin_load(f.minv, f.n, v);
        for (int j = 0; j < 6; j++) {
            x.v[j] += alpha * pv.v[j];
            r.v[j] -= alpha * y.v[j];
            z.v[j] = mi.v[j] * r.v[j];
            acc[0] += r.v[j] * z.v[j]; acc[1] += r.v[j] * r.v[j];
        }
        lin_store(f.x, f.n, v, x);
        lin_store(f.r, f.n, v, r);
        lin_store(f.z, f.n, v, z);
    }
    double* const part[2] = {f.part_rz[(k + 1) & 1], f.part_rr[(k + 1) & 1]};
    lin_block_sums<2>(acc, part);
}

// postResults (:336-347): pos = originalPosition + u, orient = Quat3D(rotation vector), both momenta zero
__global__ void k_lin_post(LinFrame f, double4* __restrict__ pose0, double4* __restrict__ pose1, double4* __restrict__ mom0, double2* __restrict__ mom1)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.n; v += gridDim.x * blockDim.x) {
        const Dof6 x = lin_load(f.x, f.n, v);
        const int i = f.e2i[v];
        const double s = f.voxel_size;
        const q4 q = q_from_rotvec(mk3(x.v[3], x.v[4], x.v[5]));
        const double meta = pose1[i].w;
        pose0[i] = make_double4(f.ijk[3 * v] * s + x.v[0], f.ijk[3 * v + 1] * s + x.v[1], f.ijk[3 * v + 2] * s + x.v[2], q.w);
        pose1[i] = make_double4(q.x, q.y, q.z, meta);
        mom0[i] = make_double4(0.0, 0.0, 0.0, 0.0);
        mom1[i] = make_double2(0.0, 0.0);
    }
}

} // namespace vxd
