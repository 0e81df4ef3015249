// The reference's README example (README.md:20-45) plus a few accessor reads, unchanged source:
//   reference:      g++ -O2 -I$VOXELYZE/include cantilever.cpp -L$VOXELYZE/lib -lvoxelyze.0.9
//   voxelyze-b200:  make -C examples            (facade headers, libvoxelyze_facade + libvoxelyze_b200; needs a B200 to run)
#include "Voxelyze.h"
#include <cstdio>

int main()
{
    CVoxelyze Vx(0.005);                                   // 5 mm voxels
    CVX_Material* pMaterial = Vx.addMaterial(1000000, 1000);
    CVX_Voxel* Voxel1 = Vx.setVoxel(pMaterial, 0, 0, 0);
    CVX_Voxel* Voxel2 = Vx.setVoxel(pMaterial, 1, 0, 0);
    CVX_Voxel* Voxel3 = Vx.setVoxel(pMaterial, 2, 0, 0);
    Voxel1->external()->setFixedAll();
    Voxel3->external()->setForce(0, 0, -1);
    for (int i = 0; i < 100; i++) Vx.doTimeStep();
    std::printf("tip z after 100 steps: %.9e  (middle voxel %.9e)\n", Voxel3->position().z, Voxel2->position().z);
    return 0;
}
