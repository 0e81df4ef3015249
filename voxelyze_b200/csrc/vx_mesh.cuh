// vx_mesh.cuh -- deformed surface mesh on the device.
//
// Replaces CVX_MeshRender::updateMesh (src/VX_MeshRender.cpp:148-218) and what it calls per vertex,
// CVX_Voxel::cornerPosition / cornerOffset (src/VX_Voxel.cpp:141-159): every mesh vertex is the average of the
// deformed corner positions of the (up to eight) voxels that share it, every quad gets a normal and a colour.
// The topology (which vertices, which quads, which voxels share a vertex; CVX_MeshRender::generateMesh,
// src/VX_MeshRender.cpp:49-145) is static between vx_set_voxels calls and is built once on the host in the
// reference's numbering, so an OBJ file written from these buffers equals the reference's line for line.
// The buffers stay in HBM (vx_mesh_device hands out their addresses, e.g. for graphics interop); vx_mesh_download copies.
// Arithmetic follows the reference's float/double mix exactly (noted per line), so with equal voxel and link state the
// vertices are bit-identical.
#pragma once
#include "vx_kernels.cuh"

namespace vxd {

enum { MESH_MATERIAL, MESH_FAILURE, MESH_STATE_INFO };        // CVX_MeshRender::viewColoring (include/VX_MeshRender.h:30-34)

struct MeshFrame {
    int n_vert, n_quad;
    const int* vert_vox;         // [n_vert][8] caller voxel index sharing the vertex through its corner j (voxelCorner order), or -1
    const int* quads;            // [n_quad][4]
    const int* quad_vox;         // [n_quad] caller voxel index
    float* vertices; float* normals; float* colors;
    const int* e2i;              // caller voxel index -> internal
    const int* vlinks;           // [6][n_vox_user] caller link index of the voxel's link in direction d, or -1
    const float* strain; const float* max_strain;      // per link (caller order), gathered for this update
    const float* ratio;          // per link: CVX_Link::strainRatio = E_pos / E_neg (src/VX_Link.cpp:67)
    const float* eps_fail; const float* eps_yield;      // per link material limits (-1: none)
    const float* mat_rgb;        // [n_mat][3] material colour / 255.0f
    const float* vox_val; const float* link_val;        // STATE_INFO: per voxel (internal order, or caller order for pressure) / per link value
    int n_user;
};

// CVX_Voxel::cornerPosition(corner) of caller voxel e (src/VX_Voxel.cpp:141-159)
__device__ __forceinline__ void mesh_corner(const Frame& f, const MeshFrame& m, int e, int corner, float& ox, float& oy, float& oz)
{
    const int v = m.e2i[e];
    const double4 p0 = f.pose0[v], p1 = f.pose1[v];
    const DevVoxMat& vm = f.vmat[meta_hi(p1.w) & VM_MAT_MASK];
    const float scale = 1 + meta_temp(p1.w) * vm.cte;                       // baseSize(): size * (1 + temp * cte), the factor in float
    double strains[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const bool pos_link = (corner & (1 << (2 - i))) != 0;
        const int l = m.vlinks[(size_t)(2 * i + (pos_link ? 0 : 1)) * m.n_user + e];
        float sgn = pos_link ? 1.0f : -1.0f, val = sgn;
        if (l >= 0 && !(m.eps_fail[l] != -1.0f && m.max_strain[l] > m.eps_fail[l])) {
            const float r = m.ratio[l], st = m.strain[l];
            const float half = pos_link ? 2.0f * st * r / (1.0f + r) : 2.0f * st / (1.0f + r);      // CVX_Link::axialStrain(positiveEnd), src/VX_Link.cpp:121-124
            val = (1 + half) * sgn;                                          // float, then stored to double
        }
        strains[i] = val;
    }
    // (0.5 * baseSize()).Scale(strains) in double, returned as Vec3D<float>
    const float cx = (float)((0.5 * (vm.size[0] * scale)) * strains[0]);
    const float cy = (float)((0.5 * (vm.size[1] * scale)) * strains[1]);
    const float cz = (float)((0.5 * (vm.size[2] * scale)) * strains[2]);
    // orient.RotateVec3D<float>(offset): products with the double quaternion, every intermediate rounded to float (include/Quat3D.h:179-186)
    const double qw = p0.w, qx = p1.x, qy = p1.y, qz = p1.z;
    const float tw = (float)(cx * qx + cy * qy + cz * qz);
    const float tx = (float)(cx * qw - cy * qz + cz * qy);
    const float ty = (float)(cx * qz + cy * qw - cz * qx);
    const float tz = (float)(-cx * qy + cy * qx + cz * qw);
    const float rx = (float)(qw * tx + qx * tw + qy * tz - qz * ty);
    const float ry = (float)(qw * ty - qx * tz + qy * tw + qz * tx);
    const float rz = (float)(qw * tz + qx * ty - qy * tx + qz * tw);
    ox = (float)p0.x + rx; oy = (float)p0.y + ry; oz = (float)p0.z + rz;    // (Vec3D<float>)pos + ...
}

__global__ void __launch_bounds__(128) k_mesh_vertices(Frame f, MeshFrame m)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.n_vert) return;
    float ax = 0.f, ay = 0.f, az = 0.f; int n = 0;
    for (int j = 0; j < 8; j++) {
        const int e = m.vert_vox[8 * (size_t)i + j];
        if (e < 0) continue;
        float x, y, z;
        mesh_corner(f, m, e, j, x, y, z);
        ax += x; ay += y; az += z; n++;
    }
    const float inv = 1.0f / (float)n;                                       // Vec3D<float>::operator/= multiplies by the reciprocal (include/Vec3D.h:71)
    m.vertices[3 * (size_t)i] = ax * inv; m.vertices[3 * (size_t)i + 1] = ay * inv; m.vertices[3 * (size_t)i + 2] = az * inv;
}

__device__ __forceinline__ float jet_r(float v) { return v < 0.5f ? 0.0f : (v > 0.75f ? 1.0f : v * 4 - 2); }        // include/VX_MeshRender.h:58-60
__device__ __forceinline__ float jet_g(float v) { return v < 0.25f ? v * 4 : (v > 0.75f ? 4 - v * 4 : 1.0f); }
__device__ __forceinline__ float jet_b(float v) { return v > 0.5f ? 0.0f : (v < 0.25f ? 1.0f : 2 - v * 4); }

// normals and colours (src/VX_MeshRender.cpp:177-217); state_type: CVoxelyze::stateInfoType, max_val: stateInfo(type, MAX)
// (for pressure: max(|min|, |max|))
__global__ void __launch_bounds__(128) k_mesh_quads(Frame f, MeshFrame m, int scheme, int state_type, float max_val)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m.n_quad) return;
    float vx[4][3];
    for (int j = 0; j < 4; j++) { const int vi = m.quads[4 * (size_t)q + j]; for (int k = 0; k < 3; k++) vx[j][k] = m.vertices[3 * (size_t)vi + k]; }
    const float ax = vx[1][0] - vx[0][0], ay = vx[1][1] - vx[0][1], az = vx[1][2] - vx[0][2];
    const float bx = vx[3][0] - vx[0][0], by = vx[3][1] - vx[0][1], bz = vx[3][2] - vx[0][2];
    float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const float len = sqrtf(nx * nx + ny * ny + nz * nz);
    if (len > 0) { nx /= len; ny /= len; nz /= len; }                        // Vec3D::Normalize divides (include/Vec3D.h:84)
    m.normals[3 * (size_t)q] = nx; m.normals[3 * (size_t)q + 1] = ny; m.normals[3 * (size_t)q + 2] = nz;

    const int e = m.quad_vox[q];
    float r = 1.0f, g = 1.0f, b = 1.0f, jet = -1.0f;
    if (scheme == MESH_MATERIAL) {
        const int mat = meta_hi(f.pose1[m.e2i[e]].w) & VM_MAT_MASK;
        r = m.mat_rgb[3 * mat]; g = m.mat_rgb[3 * mat + 1]; b = m.mat_rgb[3 * mat + 2];
    } else if (scheme == MESH_FAILURE) {                                     // any link of the voxel failed: red; yielded: yellow
        bool failed = false, yielded = false;
        for (int d = 0; d < 6; d++) {
            const int l = m.vlinks[(size_t)d * m.n_user + e];
            if (l < 0) continue;
            const float ms = m.max_strain[l];
            if (m.eps_fail[l] != -1.0f && ms > m.eps_fail[l]) failed = true;
            if (m.eps_yield[l] != -1.0f && ms > m.eps_yield[l]) yielded = true;
        }
        if (failed) { g = 0.0f; b = 0.0f; } else if (yielded) b = 0.0f;
    } else {
        if (state_type == SI_KINETIC_ENERGY || state_type == SI_DISPLACEMENT) jet = m.vox_val[m.e2i[e]] / max_val;
        else if (state_type == SI_PRESSURE) jet = 0.5 - m.vox_val[e] / (2 * max_val);
        else if (state_type == SI_STRAIN_ENERGY || state_type == SI_ENG_STRAIN || state_type == SI_ENG_STRESS) {
            float best = -3.402823466e38f;                                   // CVX_MeshRender::linkMaxColorValue
            for (int d = 0; d < 6; d++) {
                const int l = m.vlinks[(size_t)d * m.n_user + e];
                const float v = l >= 0 ? m.link_val[l] : -3.402823466e38f;
                if (v > best) best = v;
            }
            jet = best / max_val;
        } else jet = 0;
    }
    if (jet != -1.0f) { r = jet_r(jet); g = jet_g(jet); b = jet_b(jet); }
    m.colors[3 * (size_t)q] = r; m.colors[3 * (size_t)q + 1] = g; m.colors[3 * (size_t)q + 2] = b;
}

} // namespace vxd
