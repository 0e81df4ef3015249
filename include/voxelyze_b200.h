/*
 * voxelyze_b200.h -- C-ABI of the B200-native explicit dynamics step.
 *
 * This is the drop-in boundary for ONE hot path of jonhiller/Voxelyze:
 *     CVoxelyze::doTimeStep            (reference src/Voxelyze.cpp:251-284)
 * and the calls that feed it / read from it.  The reference has no FFI layer of
 * its own: its contract is the public C++ class API.  The C++
 * facade in voxelyze_b200/facade/ re-creates that class API (CVoxelyze,
 * CVX_Material, CVX_Voxel, CVX_Link, CVX_External, CVX_Collision) on top of the
 * entry points declared here; nothing else crosses the host/device boundary.
 *
 * Rules of the boundary
 *   - extern "C", plain pointers and sizes only.  No C++/torch types.
 *   - every function returns an int status (VX_OK == 0) unless noted.
 *   - the caller owns all host buffers; the library owns all device memory.
 *   - one caller thread per vx_sim handle (the reference is not thread safe either,
 *     include/Voxelyze.h:134).
 *   - there is NO CPU fallback: the CUDA build of this library fails with
 *     VX_ERR_NO_DEVICE when no sm_100 device is usable.
 *
 * Three shared objects implement exactly this header:
 *   voxelyze_b200/lib/libvoxelyze_b200.so   the product (hand written sm_100a CUDA)
 *   oracle/liboracle_port.so                CPU restatement, test infrastructure only
 *   oracle/_ref/libvxref.so                 the unmodified reference behind a shim,
 *                                           test infrastructure only
 * so a parity test is "same calls, two libraries, compare downloads".
 *
 * Index conventions
 *   voxel index  = position in the array handed to vx_set_voxels (the reference's
 *                  voxelsList order, i.e. setVoxel call order; Voxelyze.cpp:446).
 *   link index   = the order in which the reference would have created the links for
 *                  that setVoxel sequence (Voxelyze.cpp:453-455,508-539): for every
 *                  voxel in order, for dir = X+,X-,Y+,Y-,Z+,Z-, a link to an already
 *                  existing neighbour.  Query it with vx_get_links.
 *   link direction / slot numbering follows CVX_Voxel::linkDirection
 *                  (include/VX_Voxel.h:39-46): 0 X+, 1 X-, 2 Y+, 3 Y-, 4 Z+, 5 Z-.
 *   quaternions are stored (w, x, y, z) like Quat3D (include/Quat3D.h:44-49).
 */
#ifndef VOXELYZE_B200_H
#define VOXELYZE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX_ABI_VERSION 4

/* ---- status codes ------------------------------------------------------- */
#define VX_OK              0
#define VX_DIVERGED        1  /* vx_step: a link strain exceeded 100 (Voxelyze.cpp:265-269) */
#define VX_ERR_ARG        -1
#define VX_ERR_NO_DEVICE  -2
#define VX_ERR_CUDA       -3
#define VX_ERR_MATERIAL   -4  /* material model rejected (VX_Material.cpp:283-434) */
#define VX_ERR_TOPOLOGY   -5
#define VX_ERR_UNSUPPORTED -6
#define VX_ERR_ALLOC      -7
#define VX_ERR_SOLVER     -8  /* vx_linear_solve: singular system or no convergence (CVX_LinearSolver::solve returning false) */

typedef struct vx_sim vx_sim;

/* ---- materials ---------------------------------------------------------- */
/* Raw, user level description of one voxel material == what a caller sets through
 * CVX_Material (include/VX_Material.h:33-103).  The library derives the cached
 * constants of CVX_MaterialVoxel::updateDerived (src/VX_MaterialVoxel.cpp:57-79)
 * and builds one combined link material per adjacent pair exactly like
 * CVoxelyze::combinedMaterial / CVX_MaterialLink::updateAll
 * (src/Voxelyze.cpp:626-640, src/VX_MaterialLink.cpp:45-141).                     */
#define VX_MODEL_LINEAR   0   /* setModelLinear(E, fail_stress)                    */
#define VX_MODEL_BILINEAR 1   /* setModelBilinear(E, plastic_modulus, yield, fail) */
#define VX_MODEL_DATA     2   /* setModel(n_points, strain, stress)                */

typedef struct vx_material_desc {
    int32_t model;            /* VX_MODEL_*                                            */
    float   youngs_modulus;   /* Pa; LINEAR / BILINEAR                                  */
    float   plastic_modulus;  /* Pa; BILINEAR                                           */
    float   yield_stress;     /* Pa; BILINEAR                                           */
    float   fail_stress;      /* Pa; LINEAR / BILINEAR; -1 = no failure                 */
    int32_t n_points;         /* DATA: number of (strain, stress) points               */
    const float* strain;      /* DATA                                                   */
    const float* stress;      /* DATA                                                   */
    float   density;          /* kg/m^3                                                 */
    float   poissons_ratio;
    float   cte;              /* coefficient of thermal expansion, 1/degC               */
    float   mu_static;
    float   mu_kinetic;
    float   zeta_internal;    /* default 1 (VX_Material.cpp:65)                         */
    float   zeta_global;
    float   zeta_collision;
    double  ext_scale[3];     /* setExternalScaleFactor, default (1,1,1)                */
} vx_material_desc;

/* Derived voxel-material constants, for accessors and tests (all float like the
 * reference, except the two sizes).                                                */
typedef struct vx_voxmat_row {
    double nom_size;
    double size[3];           /* nom_size * ext_scale                                   */
    float  E, nu, rho, cte, mu_static, mu_kinetic;
    float  zeta_internal, zeta_global, zeta_collision;
    float  e_hat;
    float  mass, mass_inv, sqrt_mass, first_moment;
    float  moment_inertia, moment_inertia_inv;
    float  two_sq_m_e_s;      /* _2xSqMxExS                                              */
    float  two_sq_i_e_s3;     /* _2xSqIxExSxSxS                                          */
    float  eps_yield, eps_fail, sigma_yield, sigma_fail;
    int32_t linear;
    int32_t n_curve;          /* number of model data points incl. the (0,0) point     */
} vx_voxmat_row;

/* Derived link-material constants (CVX_MaterialLink, include/VX_MaterialLink.h:34-43). */
typedef struct vx_linkmat_row {
    int32_t mat_a, mat_b;     /* constituent voxel materials (a <= b)                   */
    int32_t linear;
    int32_t n_curve;
    float  E, nu, e_hat;
    float  eps_yield, eps_fail, sigma_yield, sigma_fail;
    float  a1, a2, b1, b2, b3;
    float  sq_a1, sq_a2_ip, sq_b1, sq_b2_fmp, sq_b3_ip;
} vx_linkmat_row;

/* ---- state fields for upload / download --------------------------------- */
/* Voxel fields are indexed by voxel index, link fields by link index.             */
enum vx_field {
    /* voxel, double */
    VX_F_POS = 0,            /* 3 doubles / voxel  CVX_Voxel::pos     VX_Voxel.h:163  */
    VX_F_ORIENT = 1,         /* 4 doubles (w,x,y,z)          orient   VX_Voxel.h:165  */
    VX_F_LINMOM = 2,         /* 3 doubles                    linMom   VX_Voxel.h:164  */
    VX_F_ANGMOM = 3,         /* 3 doubles                    angMom   VX_Voxel.h:166  */
    /* voxel, 32 bit */
    VX_F_TEMP = 4,           /* 1 float                      temp     VX_Voxel.h:171  */
    VX_F_VOXFLAGS = 5,       /* 1 uint32, VX_VF_* bits                                 */
    VX_F_PSTRAIN = 6,        /* 3 floats, cached poissons strain      VX_Voxel.h:179  */
    /* link, double */
    VX_F_FORCE_NEG = 16,     /* 3 doubles / link             forceNeg VX_Link.h:71    */
    VX_F_FORCE_POS = 17,
    VX_F_MOMENT_NEG = 18,
    VX_F_MOMENT_POS = 19,
    VX_F_POS2 = 20,          /* 3 doubles                    pos2     VX_Link.h:101   */
    VX_F_ANGLE1V = 21,
    VX_F_ANGLE2V = 22,
    /* link, 32 bit */
    VX_F_STRAIN = 24,        /* float                        strain   VX_Link.h:74    */
    VX_F_MAXSTRAIN = 25,
    VX_F_STRAINOFFSET = 26,
    VX_F_STRESS = 27,        /* float                        _stress  VX_Link.h:107   */
    VX_F_LINKFLAGS = 28      /* uint32, VX_LF_* bits                                   */
};

/* voxel flag bits (download) */
#define VX_VF_STATIC_FRICTION 0x1u  /* FLOOR_STATIC_FRICTION, VX_Voxel.h:146          */
#define VX_VF_SURFACE         0x2u  /* fewer than 6 links, VX_Voxel.cpp:376-381       */
#define VX_VF_GHOST           0x4u  /* z-slab halo copy: pose is imported, never integrated */
#define VX_VF_FLOOR_OFF       0x8u  /* CVX_Voxel::enableFloor(false) on this voxel (include/VX_Voxel.h:119) while the floor is on */
#define VX_VF_FLOOR_ON        0x10u /* CVX_Voxel::enableFloor(true) on this voxel while the simulation's floor is off            */
/* link flag bits (download) */
#define VX_LF_SMALL_ANGLE     0x1u  /* CVX_Link::smallAngle, VX_Link.h:103            */
#define VX_LF_LOCAL_VEL_VALID 0x2u  /* LOCAL_VELOCITY_VALID, VX_Link.h:80             */
#define VX_LF_YIELDED         0x4u  /* CVX_Link::isYielded, VX_Link.cpp:127-130       */
#define VX_LF_FAILED          0x8u  /* CVX_Link::isFailed,  VX_Link.cpp:132-135       */

/* degrees of freedom, identical to dofComponent (include/VX_External.h:18-26) */
#define VX_DOF_TX 0x01
#define VX_DOF_TY 0x02
#define VX_DOF_TZ 0x04
#define VX_DOF_RX 0x08
#define VX_DOF_RY 0x10
#define VX_DOF_RZ 0x20
#define VX_DOF_ALL 0x3F

/* ---- life cycle ---------------------------------------------------------- */
/* replaces CVoxelyze::CVoxelyze(double voxelSize)      include/Voxelyze.h:69.
 * device = CUDA ordinal (ignored by the CPU oracles).                               */
int  vx_create(double voxel_size, int device, vx_sim** out);
/* replaces CVoxelyze::~CVoxelyze / clear()             src/Voxelyze.cpp:323-355     */
void vx_destroy(vx_sim* s);
/* human readable text of the last failure on this handle (never NULL)              */
const char* vx_last_error(const vx_sim* s);
/* "cuda-sm100a", "oracle-port" or "reference" */
const char* vx_backend(void);
int  vx_abi_version(void);

/* ---- model definition ---------------------------------------------------- */
/* replaces CVoxelyze::addMaterial + CVX_Material setters (src/Voxelyze.cpp:357-382,
 * src/VX_Material.cpp:283-523).  May be called again between steps with the same
 * count to change properties mid-run (test/tVoxelyze.h:356); dynamic state is kept. */
int  vx_set_materials(vx_sim* s, int n, const vx_material_desc* descs);
int  vx_get_voxmat(const vx_sim* s, int mat, vx_voxmat_row* out);
/* link material for the unordered pair (mat_a, mat_b); derived on demand.           */
int  vx_get_linkmat(vx_sim* s, int mat_a, int mat_b, vx_linkmat_row* out);
/* model data points of the link material (strain/stress incl. the zero point)       */
int  vx_get_linkmat_curve(vx_sim* s, int mat_a, int mat_b, float* strain, float* stress, int cap);

/* replaces the setVoxel() sequence                     src/Voxelyze.cpp:422-461.
 * ijk: 3*n lattice indices (each must fit a short like the reference,
 * include/VX_Voxel.h:151); mat: material index per voxel; sim_id: NULL, or an
 * ensemble member id per voxel -- voxels of different members never link and each
 * member lives in its own lattice frame (SURVEY.md section 8e "ensemble"); flags: NULL or
 * VX_VF_GHOST per voxel.  Resets all dynamic state like a fresh CVoxelyze.          */
int  vx_set_voxels(vx_sim* s, int n, const int32_t* ijk, const uint16_t* mat,
                   const int32_t* sim_id, const uint32_t* flags);
int  vx_voxel_count(const vx_sim* s);
int  vx_link_count(const vx_sim* s);
/* link list in link-index order: negative-end voxel, positive-end voxel, axis 0..2
 * (CVX_Link::pVNeg/pVPos/axis, include/VX_Link.h:70,88).  Any pointer may be NULL.  */
int  vx_get_links(const vx_sim* s, int32_t* v_neg, int32_t* v_pos, uint8_t* axis);

/* replaces CVX_External (include/VX_External.h:45-88) for the listed voxels; voxels
 * not listed have no external.  dof: VX_DOF_* bits; force/moment: 3 floats each;
 * translation/rotation: 3 doubles each (rotation is a rotation vector; the library
 * caches the quaternion like CVX_External::rotationChanged, VX_External.cpp:100-109).
 * Any of force/moment/translation/rotation may be NULL (= zeros).                    */
int  vx_set_externals(vx_sim* s, int n, const int32_t* voxel, const uint8_t* dof,
                      const float* force, const float* moment,
                      const double* translation, const double* rotation);

/* replaces setGravity / enableFloor / enableCollisions (src/Voxelyze.cpp:596-622)   */
int  vx_set_gravity(vx_sim* s, float g);
int  vx_enable_floor(vx_sim* s, int enabled);
int  vx_enable_collisions(vx_sim* s, int enabled);
/* CVX_Collision::envelopeRadius (static, src/VX_Collision.cpp:15)                   */
int  vx_set_collision_envelope(vx_sim* s, float envelope_radius);

/* replaces CVoxelyze::setAmbientTemperature(t, true) (src/Voxelyze.cpp:585-594):
 * every voxel takes temperature t.                                                   */
int  vx_set_temperature_all(vx_sim* s, float t);
/* per ensemble member: member m takes t[m] (n_members values).                       */
int  vx_set_temperature_members(vx_sim* s, int n_members, const float* t);
/* replaces CVX_Voxel::setTemperature per voxel (src/VX_Voxel.cpp:108-114)            */
int  vx_set_temperature(vx_sim* s, int n, const float* t);

/* ---- the hot path --------------------------------------------------------- */
/* replaces CVoxelyze::doTimeStep(dt) called n_steps times (src/Voxelyze.cpp:251-284).
 * dt < 0: use vx_recommended_dt() each step (reference default argument).
 * Returns VX_OK, or VX_DIVERGED with *diverged_step (may be NULL) = number of steps
 * completed before the diverging one; like the reference, on the diverging step the
 * links are updated but the voxels are not advanced, and no later step is run.       */
int  vx_step(vx_sim* s, float dt, int n_steps, int* diverged_step);
/* vx_step with an ambient temperature program: before step k every voxel takes temperature ambient[k] -- what a caller does
 * who calls CVoxelyze::setAmbientTemperature(t_k, true) and doTimeStep(dt) in turn (src/Voxelyze.cpp:585-594, 251-284; the
 * thermally actuated robots of BASELINE config C4), n_steps of it without a host round trip per step.  Same bits as the
 * n_steps pairs of vx_set_temperature_all + vx_step(s, dt, 1).  On divergence the voxels keep the temperature of the
 * diverging step, like there.                                                                                          */
int  vx_step_ambient(vx_sim* s, float dt, int n_steps, const float* ambient, int* diverged_step);
/* optional: builds everything vx_step would otherwise build lazily on its first long call (the captured
 * CUDA graphs of 16 steps, tensor maps) without touching the state, so that a caller who times
 * steps does not time the set-up.  No reference counterpart.                          */
int  vx_prepare(vx_sim* s);
/* replaces CVoxelyze::recommendedTimeStep()           src/Voxelyze.cpp:286-311      */
int  vx_recommended_dt(vx_sim* s, float* dt);
/* replaces CVoxelyze::resetTime()                     src/Voxelyze.cpp:313-321      */
int  vx_reset(vx_sim* s);
/* simulated time (float accumulation like currentTime, Voxelyze.cpp:282)            */
float vx_time(const vx_sim* s);
/* sets CVoxelyze::currentTime and CVX_Voxel::previousDt (include/VX_Voxel.h:171; the same for every voxel): what a model
 * that moves to another handle mid-run (other devices, other voxel size) takes along besides its voxel and link state,
 * so that the next step damps with the dt of the step before it, as if nothing had happened.                          */
int  vx_set_clock(vx_sim* s, float time, float previous_dt);

/* ---- state access ---------------------------------------------------------- */
/* Copies elements [first, first+count) of a field to/from host memory.  Element
 * sizes are given at enum vx_field.  Download waits for all queued steps.            */
int  vx_download(vx_sim* s, int field, int first, int count, void* dst);
int  vx_upload(vx_sim* s, int field, int first, int count, const void* src);

/* replaces CVoxelyze::collisionList() (include/Voxelyze.h:115): watched pairs as
 * (voxel1, voxel2) voxel indices with voxel1 < voxel2, in creation order
 * (src/Voxelyze.cpp:730-747).  pairs may be NULL to query the count.                 */
int  vx_collision_pairs(vx_sim* s, int32_t* pairs, int cap, int* n_pairs);

/* number of watched pairs and of watch-list rebuilds (CVoxelyze::regenerateCollisions, src/Voxelyze.cpp:725-750) since
 * the voxels were set, as of the end of the last step call; -1 rebuilds where an implementation does not count them   */
int  vx_collision_stats(vx_sim* s, int* n_pairs, int* n_rebuilds);

/* replaces CVoxelyze::stateInfo (src/Voxelyze.cpp:752-800); info/type use the
 * reference's enum values (include/Voxelyze.h:48-67).                                */
int  vx_state_info(vx_sim* s, int info, int type, float* out);

/* ---- static solve -------------------------------------------------------
 * replaces CVoxelyze::doLinearSolve (src/Voxelyze.cpp:243-249) = CVX_LinearSolver::solve (src/VX_LinearSolver.cpp:49-113):
 * the model is linearised about the nominal lattice (only the beam constants a1, a2, b1, b2, b3 of the link materials
 * enter, calculateA :116-233), degrees of freedom fixed through CVX_External keep their CURRENT displacement / rotation
 * vector, the others carry the external force / moment (applyBX :273-328), and the solution overwrites the voxel poses:
 * pos = originalPosition + u, orient = Quat3D(rotation vector), momenta zero; link state is left as it was
 * (postResults :336-347).  The reference factorises with PARDISO; here the system is solved matrix-free by a
 * preconditioned conjugate-gradient iteration in FP64 on the device, to the relative residual |r| <= rel_tol * |r0|
 * (rel_tol <= 0: 1e-10; max_iter <= 0: 200 000).  Returns VX_OK (poses written), or VX_ERR_SOLVER with the state
 * untouched when a part of the model is not held (singular matrix; PARDISO error -4) or max_iter was not enough.
 * iterations / rel_residual may be NULL.  Ensembles are solved as one block-diagonal system.                        */
int  vx_linear_solve(vx_sim* s, double rel_tol, int max_iter, int* iterations, double* rel_residual);

/* ---- deformed surface mesh ---------------------------------------------
 * replaces CVX_MeshRender (include/VX_MeshRender.h:26-62, src/VX_MeshRender.cpp:49-218) and what it calls per vertex,
 * CVX_Voxel::cornerPosition / cornerOffset (src/VX_Voxel.cpp:141-159).  vx_mesh_build = generateMesh (exposed faces as
 * quads over shared vertices, numbered like the reference so that saveObj output is identical), vx_mesh_update =
 * updateMesh(colorScheme, stateType) (coloring: 0 material, 1 failure, 2 state info with a CVoxelyze::stateInfoType);
 * the float buffers (3 per vertex, 3 normal + 3 colour components per quad) stay on the device: vx_mesh_device returns
 * their addresses, vx_mesh_download copies any of them (NULL = skip).  One simulation per mesh (no ensembles).
 * The reference shim drives the real CVX_MeshRender; the oracle port returns VX_ERR_UNSUPPORTED.                  */
int  vx_mesh_set_material_colors(vx_sim* s, int n_materials, const unsigned char* rgba);
int  vx_mesh_build(vx_sim* s, int* n_vertices, int* n_quads);
int  vx_mesh_update(vx_sim* s, int coloring, int state_type);
int  vx_mesh_counts(vx_sim* s, int* n_vertices, int* n_quads);
int  vx_mesh_download(vx_sim* s, float* vertices, int32_t* quads, float* normals, float* colors, int32_t* quad_voxel);
int  vx_mesh_device(vx_sim* s, uint64_t* vertices, uint64_t* quads, uint64_t* normals, uint64_t* colors);

/* ---- device side hooks (CUDA build only; oracles return VX_ERR_UNSUPPORTED) -- */
/* run all work of this handle on the given cudaStream_t (passed as an integer; 0 is CUDA's
 * legacy default stream and is used as such).  VX_OWN_STREAM selects the library's own
 * non-blocking stream again (the initial state).  Lets a caller order NCCL halo traffic
 * and event timing with the step.                                                     */
#define VX_OWN_STREAM (~(uint64_t)0)
int  vx_set_stream(vx_sim* s, uint64_t cuda_stream);
/* z-slab halo exchange support: device address and element
 * range of the two packed pose record arrays (pose0: pos.xyz+orient.w, pose1:
 * orient.xyz+meta) of all voxels with lattice z == iz, so that one plane can be
 * sent/received as two contiguous messages.  rec_bytes is the record size (32).                                                                         */
int  vx_pose_plane(vx_sim* s, int iz, uint64_t* dev_ptr0, uint64_t* dev_ptr1,
                   int* count, int* rec_bytes);
/* stores `count` received pose records (two arrays in the vx_pose_plane layout, device
 * addresses) into the ghost voxels of layer iz: position and orientation are replaced,
 * the receiving voxel keeps its own flag word (it stays a ghost) and takes the sender's
 * temperature.                                                                        */
int  vx_halo_import(vx_sim* s, int iz, uint64_t src_ptr0, uint64_t src_ptr1, int count);
/* same, queued on the given cudaStream_t instead of the handle's stream (the stream the
 * halo messages arrive on)                                                            */
int  vx_halo_import_on(vx_sim* s, int iz, uint64_t src_ptr0, uint64_t src_ptr1, int count,
                       uint64_t cuda_stream);
/* asynchronous stepping, for z-slab runs that overlap the halo exchange with the step
 * ("boundary-plane voxels integrated first -> halo push -> interior").
 * A call is vx_step_begin, any number of steps queued with vx_step_enqueue, and
 * vx_step_end, which blocks and reports exactly like vx_step.  Nothing in between blocks
 * the host.  One step is either vx_step_enqueue(VX_PART_ALL) or VX_PART_Z_BOUNDARY followed
 * by VX_PART_Z_INTERIOR: the boundary part advances the layers that hold ghost planes
 * (VX_VF_GHOST) and their neighbours, i.e. everything an exchange sends or receives, the
 * interior part the rest.  After a boundary part vx_pose_plane/vx_halo_import* address the
 * NEW state, so the caller can ship the fresh boundary poses on a second stream (ordered
 * by events: after the boundary part, before the next step) while the interior part runs.
 * Same arithmetic, same bits as vx_step.  Fused lattice path only; no other call on the
 * handle between begin and end except vx_pose_plane, vx_halo_import*, vx_launch_count.      */
#define VX_PART_ALL        0
#define VX_PART_Z_BOUNDARY 1
#define VX_PART_Z_INTERIOR 2
int  vx_step_begin(vx_sim* s, float dt);
int  vx_step_enqueue(vx_sim* s, int part);
int  vx_step_end(vx_sim* s, int* diverged_step);
/* peer-memory halo for z-slab runs, one process per GPU on one NVLink/NVSwitch node: each
 * slab stores its fresh boundary poses straight into its neighbours' ghost layers (CUDA IPC
 * mappings) from its own kernel and bumps an arrival counter there; no NCCL call and no host
 * work per step.  Setup, after vx_set_voxels on every slab:
 *   owner of a ghost layer:  vx_peer_export(s, ghost_iz, from_above, &desc)  (from_above: the
 *       writer is the slab above, i.e. this is the top ghost layer) and hands the descriptor
 *       to that neighbour by any means (bench.py: torch.distributed all_gather of the bytes);
 *   the neighbour:           vx_peer_attach(s, send_iz, &desc)  -- its layer send_iz is the
 *       one mirrored into that ghost layer.
 * vx_slab_step(s, dt, n, &div) then runs n steps like vx_step, each as: wait for the
 * neighbours' previous delivery -> boundary part -> [second stream: push + signal] overlapped
 * with the interior part.  All slabs must make the same sequence of vx_slab_step /
 * vx_slab_exchange calls.  vx_slab_exchange ships the CURRENT boundary poses (needed once after
 * state was changed by vx_upload or vx_reset); the neighbours' deliveries are awaited by the next
 * vx_slab_step.  A neighbour that does not deliver within 30 s (environment variable
 * VX_PEER_TIMEOUT_S) makes vx_slab_step fail (VX_ERR_CUDA) instead of hanging the GPU.         */
#define VX_PEER_DESC_BYTES 512
typedef struct vx_peer_desc { unsigned char bytes[VX_PEER_DESC_BYTES]; } vx_peer_desc;
int  vx_peer_export(vx_sim* s, int ghost_iz, int from_above, vx_peer_desc* out);
int  vx_peer_attach(vx_sim* s, int send_iz, const vx_peer_desc* peer_ghost);
int  vx_peer_detach(vx_sim* s);
int  vx_slab_step(vx_sim* s, float dt, int n_steps, int* diverged_step);
int  vx_slab_exchange(vx_sim* s);
/* vx_slab_step in two halves for a caller that drives several slabs from ONE thread: vx_slab_step_begin queues all n steps
 * (with their waits, pushes and signals) and returns without blocking, vx_slab_step_finish blocks and reports like
 * vx_slab_step.  Begin every slab before finishing any; each slab must be on a device of its own.                      */
int  vx_slab_step_begin(vx_sim* s, float dt, int n_steps);
int  vx_slab_step_finish(vx_sim* s, int* diverged_step);
/* number of kernels this handle has launched so far (bench.py "gpu_launches").       */
int64_t vx_launch_count(const vx_sim* s);
/* block until all queued work of this handle is done.                                */
int  vx_sync(vx_sim* s);
/* measurement hook for bench.py: runs n_steps steps one launch at a time with CUDA events
 * around each kernel group on the launching stream and returns accumulated device
 * milliseconds: ms[0] link-force kernels, ms[1] voxel-integrate kernel, ms[2] everything
 * else (Poisson pre-pass, collisions), ms[3] whole steps; launches[0..2] = kernel launches
 * per group.  Same arithmetic as vx_step.                                                 */
int  vx_step_profile(vx_sim* s, float dt, int n_steps, float* ms, int* launches);
/* the watched pairs like vx_collision_pairs plus the contact force of the last step on each
 * (CVX_Collision::contactForce, src/VX_Collision.cpp:34-56: `force` acts on the first voxel, -force
 * on the second); forces: 3 floats per pair, same order as pairs.                              */
int  vx_collision_forces(vx_sim* s, int32_t* pairs, float* forces, int cap, int* n_pairs);
/* persistent state of links [first, first+count) (caller link order, vx_get_links) as packed records:
 * what CVX_Link keeps between steps (include/VX_Link.h:74-107).  Upload is how a caller carries
 * the state of surviving links across vx_set_voxels (the reference's setVoxel only recreates the
 * links of the edited voxel, src/Voxelyze.cpp:485-498) or across a change of layout
 * (vx_enable_collisions mid-run).  flags: VX_LF_SMALL_ANGLE | VX_LF_LOCAL_VEL_VALID are stored,
 * yielded/failed are derived from max_strain.                                               */
/* all six voxel fields of a range of voxels in ONE call (one kernel, one copy): what CVX_Voxel::position() / orientation() /
 * velocity() / ... of a single voxel needs between steps (SURVEY.md section 3.5, 8b hazard 5).                            */
typedef struct vx_voxel_state {
    double pos[3], orient[4], linmom[3], angmom[3];     /* orient: w, x, y, z */
    float temp; uint32_t flags;                          /* VX_VF_* */
} vx_voxel_state;
int  vx_download_voxel_state(vx_sim* s, int first, int count, vx_voxel_state* dst);
typedef struct vx_link_state {
    double pos2[3], angle1v[3], angle2v[3];
    float strain, max_strain, strain_offset, stress;
    uint32_t flags, reserved;
} vx_link_state;
int  vx_download_link_state(vx_sim* s, int first, int count, vx_link_state* dst);
int  vx_upload_link_state(vx_sim* s, int first, int count, const vx_link_state* src);
/* dynamic-state checkpoint (the reference has none: saveJSON stores the initial configuration only,
 * include/Voxelyze.h:78).  vx_save_state writes every time-dependent device array of the handle
 * (voxel poses and momenta, link state, step bookkeeping) to a file; vx_load_state restores it
 * into a handle that was built with the same materials, voxels, externals and options
 * (VX_ERR_ARG otherwise) so that continuing is bit-identical to never having stopped.
 * Collision watch lists are rebuilt at the next step.                                        */
int  vx_save_state(vx_sim* s, const char* path);
int  vx_load_state(vx_sim* s, const char* path);
/* select kernel variant (tests; the layout part takes effect at the next vx_set_voxels):
 *   0 auto: models of at most 700 voxels without halo flags and without self-collisions: small-model kernel (3);
 *           otherwise the fused lattice kernel for bodies whose bounding box is at most 8x their voxel count (holes
 *           are padded with inert cells, sparse bodies launch only their occupied brick groups), general path beyond
 *   1 general layout, one-step kernels (k_link<AXIS> x3 + k_voxel per step), any topology
 *   3 general layout stepped by k_small_steps: ONE thread-block cluster (<= 16 CTAs) runs all steps of a vx_step call in
 *     a single launch, cluster barriers between the link and voxel phases (with self-collisions on, or dt < 0 on
 *     Poisson models, the one-step kernels of path 1 run instead)
 *   fused lattice kernel (one warp per 4x4x2 brick), with a fixed staging flavour:
 *   5 cp.async staging (k_lattice_warp; what 0 picks for ensembles of small boxes)
 *   7 TMA staging (k_lattice_tma; what 0 picks on large lattices)
 *   all of them produce bit-identical state.
 * any other value: VX_ERR_ARG */
int  vx_set_path(vx_sim* s, int path);
/* which layout the handle runs: 1 general, 2 fused lattice (decided by vx_set_voxels)  */
int  vx_active_path(const vx_sim* s);
/* name of the kernel that dominates a step of this handle (static string, for reports)  */
const char* vx_kernel_name(const vx_sim* s);

/* ---- one lattice on several GPUs of ONE process ---------------------------------
 * What a caller of the C++ class API needs to reach more than one GPU: a vx_slabbed handle takes the WHOLE model in the
 * caller's numbering, cuts it into z-slabs (one vx_sim per listed device, one ghost plane per cut), and fans the calls of
 * the hot path out / gathers state back, in the voxel and link numbering of the whole model.  The functions mirror their
 * vx_* namesakes (same argument meaning, same status codes); bits are those of the unsplit run.  Halo transport:
 * vx_slabbed_halo_mode = 2: the step kernels store boundary poses into the neighbours' ghost planes themselves (peer
 * memory over NVLink, all devices queued before any is waited for); 1: host copies after every step (implementations
 * without peer memory); 0: the model runs on one slab (fewer than four z planes, or one device listed).
 * Poisson materials: a ghost copy lacks the links its Poisson strain needs (src/VX_Voxel.cpp:300-374), so the owners' values
 * travel with the halo.  Restrictions: one body (no ensemble ids), no self-collisions.  devices == NULL: devices 0 .. n_slabs-1; the
 * same device may be listed more than once (tests on one GPU).  vx_slabbed_slab exposes slab k for reports
 * (vx_kernel_name, vx_launch_count, vx_active_path); do not step it directly.                                        */
typedef struct vx_slabbed vx_slabbed;
int  vx_slabbed_create(double voxel_size, int n_slabs, const int* devices, vx_slabbed** out);
void vx_slabbed_destroy(vx_slabbed* m);
const char* vx_slabbed_last_error(const vx_slabbed* m);
int  vx_slabbed_slab_count(const vx_slabbed* m);                 /* slabs the current model uses */
vx_sim* vx_slabbed_slab(vx_slabbed* m, int k);
int  vx_slabbed_halo_mode(const vx_slabbed* m);
int  vx_slabbed_set_materials(vx_slabbed* m, int n, const vx_material_desc* descs);
int  vx_slabbed_set_gravity(vx_slabbed* m, float g);
int  vx_slabbed_enable_floor(vx_slabbed* m, int enabled);
int  vx_slabbed_set_voxels(vx_slabbed* m, int n, const int32_t* ijk, const uint16_t* mat);
int  vx_slabbed_voxel_count(const vx_slabbed* m);
int  vx_slabbed_link_count(const vx_slabbed* m);
int  vx_slabbed_get_links(const vx_slabbed* m, int32_t* v_neg, int32_t* v_pos, uint8_t* axis);
int  vx_slabbed_set_externals(vx_slabbed* m, int n, const int32_t* voxel, const uint8_t* dof,
                              const float* force, const float* moment,
                              const double* translation, const double* rotation);
int  vx_slabbed_set_temperature_all(vx_slabbed* m, float t);
int  vx_slabbed_set_temperature(vx_slabbed* m, int n, const float* t);
int  vx_slabbed_step(vx_slabbed* m, float dt, int n_steps, int* diverged_step);
int  vx_slabbed_recommended_dt(vx_slabbed* m, float* dt);
int  vx_slabbed_reset(vx_slabbed* m);
float vx_slabbed_time(const vx_slabbed* m);
int  vx_slabbed_set_clock(vx_slabbed* m, float time, float previous_dt);
int  vx_slabbed_download(vx_slabbed* m, int field, int first, int count, void* dst);
int  vx_slabbed_upload(vx_slabbed* m, int field, int first, int count, const void* src);      /* voxel fields */
int  vx_slabbed_download_voxel_state(vx_slabbed* m, int first, int count, vx_voxel_state* dst);
int  vx_slabbed_download_link_state(vx_slabbed* m, int first, int count, vx_link_state* dst);
int  vx_slabbed_upload_link_state(vx_slabbed* m, int first, int count, const vx_link_state* src);
/* CVoxelyze::stateInfo of the whole model: voxel quantities reduced per slab on its device and combined, link quantities
 * gathered in the whole model's numbering (a cut-crossing link lives in two slabs) and reduced on the host.            */
int  vx_slabbed_state_info(vx_slabbed* m, int info, int type, float* out);
/* checkpoint of a slabbed run: one vx_save_state file per slab, "<path>.<k>of<n>"; restores into a handle built from the
 * same model with the same number of slabs, and the run continues bit-identically.                                    */
int  vx_slabbed_save_state(vx_slabbed* m, const char* path);
int  vx_slabbed_load_state(vx_slabbed* m, const char* path);
int64_t vx_slabbed_launch_count(const vx_slabbed* m);

#ifdef __cplusplus
}
#endif
#endif /* VOXELYZE_B200_H */
