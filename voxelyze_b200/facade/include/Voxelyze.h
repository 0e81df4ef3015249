// Voxelyze.h -- drop-in CVoxelyze of the voxelyze_b200 facade.
//
// Same public C++ API as the reference (include/Voxelyze.h:45-134) for the dynamics path:
// materials, voxels, externals, environment, doTimeStep, recommendedTimeStep, resetTime, state
// accessors and stateInfo.  Existing callers switch their include path and link
// libvoxelyze_facade.so + libvoxelyze_b200.so; nothing else changes.
//
// Behind the API the object only records the model; every time-dependent quantity lives in HBM
// and is produced by the CUDA kernels through the C-ABI of include/voxelyze_b200.h.  Changes a
// caller makes through returned handles (material setters, external()->set*, setVoxel mid-run)
// are detected by change counters at the next doTimeStep / accessor and uploaded then.
// *.vxl.json files load and save like the reference's (facade/src/voxelyze_json.cpp; the RapidJSON-typed
// overloads are not provided).  doLinearSolve / CVX_LinearSolver (VX_LinearSolver.h) and CVX_MeshRender (VX_MeshRender.h) run on
// the device as well.
#ifndef VXB200_VOXELYZE_H
#define VXB200_VOXELYZE_H

#include <vector>
#include <string>
#include <list>
#include <algorithm>
#include <map>
#include <unordered_map>
#include <cstdint>

#include "VX_Material.h"
#include "VX_MaterialVoxel.h"
#include "VX_MaterialLink.h"
#include "VX_Voxel.h"
#include "VX_Link.h"
#include "VX_Collision.h"

#define DEFAULT_VOXEL_SIZE 0.001

struct vx_sim;
struct vx_slabbed;

class CVoxelyze {
public:
    enum stateInfoType { DISPLACEMENT, VELOCITY, KINETIC_ENERGY, ANGULAR_DISPLACEMENT, ANGULAR_VELOCITY, ENG_STRESS, ENG_STRAIN, STRAIN_ENERGY, PRESSURE, MASS };
    enum valueType { MIN, MAX, TOTAL, AVERAGE };

    CVoxelyze(double voxelSize = DEFAULT_VOXEL_SIZE);
    CVoxelyze(const char* jsonFilePath) : voxSize(DEFAULT_VOXEL_SIZE) { loadJSON(jsonFilePath); }   // include/Voxelyze.h:70
    ~CVoxelyze();
    CVoxelyze(CVoxelyze& VIn) : voxSize(DEFAULT_VOXEL_SIZE) { *this = VIn; }     // include/Voxelyze.h:73: a copy of the MODEL (materials, voxels,
    CVoxelyze& operator=(CVoxelyze& VIn);                                           // externals, environment), not of the dynamic state

    bool loadJSON(const char* jsonFilePath);    // include/Voxelyze.h:77
    bool saveJSON(const char* jsonFilePath);    // include/Voxelyze.h:78 (initial configuration only, like the reference)

    void clear();

    bool doLinearSolve();                       // include/Voxelyze.h:80, src/Voxelyze.cpp:243-249 (always true, like the reference; CVX_LinearSolver::solve reports)
    bool doTimeStep(float dt = -1.0f);
    float recommendedTimeStep() const;
    void resetTime();

    CVX_Material* addMaterial(float youngsModulus = 1e6f, float density = 1e3f);
    CVX_Material* addMaterial(const CVX_Material& mat);
    bool removeMaterial(CVX_Material* toRemove);
    bool replaceMaterial(CVX_Material* replaceMe, CVX_Material* replaceWith);
    int materialCount() { return (int)voxelMats.size(); }
    CVX_Material* material(int materialIndex) { return (CVX_Material*)voxelMats[materialIndex]; }

    CVX_Voxel* setVoxel(CVX_Material* material, int xIndex, int yIndex, int zIndex);
    CVX_Voxel* voxel(int xIndex, int yIndex, int zIndex) const;
    int voxelCount() const { return (int)voxelsList.size(); }
    CVX_Voxel* voxel(int voxelIndex) const { return voxelsList[voxelIndex]; }
    const std::vector<CVX_Voxel*>* voxelList() const { return &voxelsList; }

    int indexMinX() const { return bound(0, false); }
    int indexMaxX() const { return bound(0, true); }
    int indexMinY() const { return bound(1, false); }
    int indexMaxY() const { return bound(1, true); }
    int indexMinZ() const { return bound(2, false); }
    int indexMaxZ() const { return bound(2, true); }

    CVX_Link* link(int xIndex, int yIndex, int zIndex, CVX_Voxel::linkDirection direction) const;
    int linkCount() const;
    CVX_Link* link(int linkIndex);
    const std::vector<CVX_Link*>* linkList() const;
    const std::vector<CVX_Collision*>* collisionList() const;

    void setVoxelSize(double voxelSize);            // src/Voxelyze.cpp:643-668: positions scale, motion halts, links restart
    double voxelSize() const { return voxSize; }

    void setAmbientTemperature(float relativeTemperature, bool allVoxels = false);
    float ambientTemperature() const { return ambientTemp; }
    void setGravity(float g = 1.0f);
    float gravity() const { return grav; }
    void enableFloor(bool enabled = true);
    bool isFloorEnabled() const { return floor; }
    void enableCollisions(bool enabled = true);
    bool isCollisionsEnabled() const { return collisions; }

    float stateInfo(stateInfoType info, valueType type);

    // facade extras (additive): the dynamic state, which saveJSON does not capture (vx_save_state / vx_load_state);
    // loadState needs an object built with the same materials, voxels and externals
    bool saveState(const char* path);
    bool loadState(const char* path);
    // device ordinal before the first step; the raw C-ABI handle
    void setDevice(int cudaDevice) { setDevices(std::vector<int>(1, cudaDevice)); }
    // several GPUs of this process: the lattice is cut into z-slabs, one per listed device (vx_slabbed_*, include/voxelyze_b200.h),
    // doTimeStep and the voxel / link accessors work in the numbering of the whole model, bits as on one device.  May be called
    // at any time: dynamic state moves with the model.  Not available while slabbed (the call aborts with a message; setDevices
    // with one device first): self-collisions, the mesh, the static solve, handle().  saveState / loadState of a
    // slabbed object write / read one file per slab ("<path>.<k>of<n>").
    // The environment variable VX_DEVICES ("0,1,2,3") sets the initial list, so that an unmodified caller of the class API
    // reaches every GPU; a list that came from there is a wish: a model or a call that cannot run slabbed moves the model to
    // the first listed device instead of aborting.
    void setDevices(const std::vector<int>& cudaDevices);
    int deviceCount() const { return (int)devices.size(); }
    bool isSlabbed() const { sync(); return hm != nullptr; }
    vx_sim* handle() const { sync(); if (hm) notSlabbed("handle()"); return h; }

private:
    double voxSize;
    float ambientTemp = 0.0f, grav = 0.0f;
    bool floor = false, collisions = false;
    int device = 0;
    mutable std::vector<int> devices;       // more than one entry: slabbed over these devices

    std::vector<CVX_MaterialVoxel*> voxelMats;
    std::vector<CVX_Voxel*> voxelsList;
    std::unordered_map<uint64_t, CVX_Voxel*> cells;

    // ---- device side, maintained lazily (logically const: accessors may have to upload first)
    mutable vx_sim* h = nullptr;
    mutable vx_slabbed* hm = nullptr;       // instead of h when the model runs on several devices
    void destroyHandle() const;
    mutable bool devicesFromEnv = false, syncing = false, clockPending = false;
    mutable float clockTime = 0.0f;
    void notSlabbed(const char* what) const;
    mutable bool topologyDirty = true, envDirty = true, tempAllDirty = false;
    mutable uint64_t matChangesSeen = ~0ull, extChanges = 0, extChangesSeen = ~0ull;
    mutable std::vector<CVX_Link*> linksList;
    mutable std::map<std::pair<CVX_Voxel*, int>, CVX_Link*> linkPool;
    mutable std::list<CVX_MaterialLink*> linkMats;
    mutable std::vector<CVX_Collision*> collisionsList;
    mutable bool stepped = false;           // dynamic state exists on the device
    mutable std::vector<CVX_Voxel*> floorEdits;  // voxels whose per-voxel floor flag still has to reach the device
    float previousDt = 0.0f;                // CVX_Voxel::previousDt (include/VX_Voxel.h:171), the same for every voxel
    void applyFloorEdits() const;
    mutable float envelopeSeen = 0.0f;

    // host mirror of the voxel state, refreshed per step on demand
    mutable uint64_t epoch = 1;             // bumped whenever device state changes
    mutable std::vector<uint64_t> mirrorEpoch;
    mutable std::vector<double> mPos, mOrient, mLin, mAng;
    mutable std::vector<float> mTemp; mutable std::vector<uint32_t> mFlags;
    mutable int singleFetches = 0;
    mutable std::vector<int> removedIndices, pendingStateEdit;   // topology edits since the last device rebuild
    mutable std::vector<unsigned char> linkStateMirror;          // vx_link_state records fetched before the first edit of a batch
    mutable bool linkStateFetched = false;
    mutable std::vector<CVX_Voxel*> editedVoxels;                // voxels whose links restart (material swapped), src/Voxelyze.cpp:485-498
    void fetchLinkState() const;

    static uint64_t key(int x, int y, int z) { return ((uint64_t)(uint16_t)(int16_t)x << 32) | ((uint64_t)(uint16_t)(int16_t)y << 16) | (uint64_t)(uint16_t)(int16_t)z; }
    int bound(int axis, bool max) const;
    void sync() const;                      // bring the device model up to date
    void rebuildTopology() const;
    void uploadMaterials() const;
    void uploadExternals() const;
    void fetchVoxel(int index) const;       // make the mirror of one voxel current
    void fetchAll() const;
    void removeVoxel(int x, int y, int z);
    CVX_MaterialLink* combinedMaterial(CVX_MaterialVoxel* a, CVX_MaterialVoxel* b) const;
    [[noreturn]] void die(const char* what) const;

    bool staticSolve(double relTol, int maxIter, int* iterations, double* residual, std::string* error);

    friend class CVX_Voxel;
    friend class CVX_Link;
    friend class CVX_LinearSolver;
    friend struct LinkRead;
};

#endif // VXB200_VOXELYZE_H
