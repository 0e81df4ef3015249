// VX_MeshRender.h -- drop-in CVX_MeshRender of the voxelyze_b200 facade (reference: include/VX_MeshRender.h:26-62).
//
// Same public interface: construct on a CVoxelyze, generateMesh() after voxels were added or removed, updateMesh(colouring,
// state type) after the state changed, saveObj(path).  The mesh lives on the device (vx_mesh_* of include/voxelyze_b200.h:
// one thread per vertex averages the deformed corners of the voxels around it, one thread per quad computes normal and
// colour); saveObj / the accessors below copy it to the host.  glDraw() is a no-op: there is no OpenGL in this build (the
// reference compiles its drawing code out as well unless USE_OPEN_GL is defined).
#ifndef VXB200_MESH_H
#define VXB200_MESH_H

#include <vector>
#include <cstdint>
#include "Voxelyze.h"

class CVX_MeshRender {
public:
    enum viewColoring { MATERIAL, FAILURE, STATE_INFO };        // include/VX_MeshRender.h:30-34

    CVX_MeshRender(CVoxelyze* voxelyzeInstance);
    void generateMesh();
    void updateMesh(viewColoring colorScheme = MATERIAL, CVoxelyze::stateInfoType stateType = CVoxelyze::DISPLACEMENT);
    void saveObj(const char* filePath);
    void glDraw() {}

    // facade extras (additive): sizes, host copies of the buffers, their device addresses (graphics interop without a host round trip)
    int vertexCount() const { return nVert; }
    int quadCount() const { return nQuad; }
    const std::vector<float>& vertexData();         // x1 y1 z1 x2 ...
    const std::vector<int>& quadData();             // four vertex indices per quad
    const std::vector<float>& quadNormalData();
    const std::vector<float>& quadColorData();
    bool deviceBuffers(uint64_t* vertices, uint64_t* quads, uint64_t* normals, uint64_t* colors) const;

private:
    CVoxelyze* vx;
    int nVert = 0, nQuad = 0;
    bool hostCurrent = false;
    std::vector<float> vertices, quadNormals, quadColors;
    std::vector<int> quads, quadVoxIndices;
    void fetch();
};
#endif
