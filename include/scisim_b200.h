/* include/scisim_b200.h
 *
 * C ABI of libscisim_b200.so -- the B200 (sm_100a) collision-detection + unconstrained-flow back end
 * that drops in under SCISim's Ball2DSim / RigidBody2DSim / RigidBody3DSim.
 *
 * The reference has no FFI layer: its seam is three C++ virtual interfaces plus one free function
 * (SURVEY.md 8b). Each entry point below names the reference interface it replaces (paths relative
 * to the SCISim checkout); INTEGRATION.md shows the C++14 shim a maintainer adds on the SCISim side.
 *
 * Conventions
 *   - plain pointers and sizes only; all floating point is IEEE binary64, indices are uint32_t
 *   - vectors use SCISim's own layouts (ball2d: q = [x0,y0,x1,y1,...]; rb3d: q = [3N pos | 9N row-major R])
 *   - every call returns SG_OK (0) or an error code; text via sg_last_error(). The reference convention
 *     (print to std::cerr and std::exit(EXIT_FAILURE), e.g. rigidbody3d/RigidBody3DSim.cpp:908-909) is
 *     applied by the host shim, not by this library
 *   - single-threaded, non-reentrant per context, like the reference's sims
 *   - input pointers are HOST pointers unless a function says "device"; outputs returned through
 *     sg_contacts / sg_pairs point into library-owned pinned host memory valid until the next call on
 *     the same context
 *   - there is no CPU fallback: if no CUDA device is usable sg_create fails
 */
#ifndef SCISIM_B200_H
#define SCISIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_OK 0
#define SG_ERR_INVALID 1     /* bad argument / call order */
#define SG_ERR_CUDA 2        /* CUDA runtime error */
#define SG_ERR_UNSUPPORTED 3 /* pair of geometry types the reference itself aborts on */
#define SG_ERR_INTERNAL 4
#define SG_ERR_REBALANCE 5   /* multi-GPU slabs: a body left its slab's neighbourhood -- re-partition the scene and repeat the step */

/* UnconstrainedMap implementations (created by name in ball2d/Ball2DUtilities.cpp:37,
 * rigidbody2d/RigidBody2DUtilities.cpp:39-43, rigidbody3d/RigidBody3DUtilities.cpp:40-44) */
#define SG_MAP_SYMPLECTIC_EULER 0 /* ball2d/SymplecticEulerMap.cpp:21-32, rigidbody2d/SymplecticEulerMap.cpp:15-38 */
#define SG_MAP_VERLET 1           /* ball2d/VerletMap.cpp:15-41, rigidbody2d/VerletMap.cpp:27-55 */
#define SG_MAP_SPLIT_HAM 2        /* rigidbody3d/UnconstrainedMaps/SplitHamMap.cpp:17-182 */
#define SG_MAP_DMV 3              /* rigidbody3d/UnconstrainedMaps/DMVMap.cpp:103-207 */
#define SG_MAP_EXPONENTIAL_EULER 4 /* rigidbody3d/UnconstrainedMaps/ExponentialEulerMap.cpp:13-91 (orientation = orthogonal polar factor, 1e-12 of the reference's JacobiSVD route) */
/* OR into the map kind of sg_rb3d_flow / sg_rb3d_step for every flow after a simulation's first.  Both maps start with
 * v1 = fsys.M() * v0 (SplitHamMap.cpp:44, DMVMap.cpp:130).  RigidBody3DState's constructor stores the world-space inertia block
 * transposed, M(r,c) = I(c,r) (formWorldSpaceMassMatrix, RigidBody3DState.cpp:165-182); RigidBody3DSim::flow then calls
 * updateMandMinv (RigidBody3DSim.cpp:442,517,592), which assigns I = R I0 R^T through a column-major map over the same values
 * (RigidBody3DState.cpp:444-446): from then on M(r,c) = I(r,c).  I is symmetric only up to rounding, so the two differ in the last
 * bit of the angular momentum when the angular velocity is not zero.  Without the flag: M as constructed. */
#define SG_MAP_M_UPDATED 0x100
#define SG_MAP_NONE (-1)           /* slab calls only: q1 was supplied by the caller (sg_ball2d_slab_upload_q1), nothing is integrated */

/* contact types, in the order the reference appends them to active_set */
#define SG_BALL_BALL 0   /* ball2d/Constraints/BallBallConstraint */
#define SG_BALL_DRUM 1   /* ball2d/Constraints/BallStaticDrumConstraint */
#define SG_BALL_PLANE 2  /* ball2d/Constraints/BallStaticPlaneConstraint */
/* with portals (ball2d/Ball2DSim.cpp:653-728): after the regular ball-ball contacts, ascending (i,j) */
#define SG_BALL_BALL_TELEPORTED 3      /* BallBallConstraint{ i, j, x0, x1, ri, rj, teleported = true } */
#define SG_BALL_BALL_KICK_TELEPORTED 4 /* KinematicKickBallBallConstraint{ i, j, x0, x1, ri, rj, kick, true } (Lees-Edwards portal) */
/* rigidbody3d (rigidbody3d/Constraints/): body-body in ascending (i,j) candidate order, then planes plane-major */
#define SG_SPHERE_SPHERE 10           /* SphereSphereConstraint{ i, j, n, p, ri, rj } */
#define SG_KINEMATIC_SPHERE_SPHERE 11 /* KinematicSphereSphereConstraint: i = free sphere, j = kinematic one, p = its centre at q0 */
#define SG_BODY_BODY 12               /* BodyBodyConstraint{ i, j, p, n, q0 } (box-box, mesh-mesh) */
#define SG_KINEMATIC_BODY_BODY 13     /* KinematicObjectBodyConstraint{ i, j, p, n, q0 } */
#define SG_PLANE_SPHERE 14            /* StaticPlaneSphereConstraint: j = plane */
#define SG_PLANE_BOX 15               /* StaticPlaneBoxConstraint: j = plane, aux = corner number, p = x0 + R0*corner */
#define SG_PLANE_BODY 16              /* StaticPlaneBodyConstraint: j = plane, aux = convex hull vertex, p = collision point at q0 */
#define SG_CYLINDER_SPHERE 17         /* StaticCylinderSphereConstraint{ i, r_i, staticCylinder(j), j }: n = computeN( q0 ), p = x0 - r n, depth = min( 0, R - dist( q1 ) - r ) */
#define SG_CYLINDER_BODY 18           /* StaticCylinderBodyConstraint{ i, p, staticCylinder(j), j, q0 }: aux = hull vertex, n from the centre of mass */
/* with portals (rigidbody3d/RigidBody3DSim.cpp:1338-1397): after the contacts of the un-teleported pairs, ascending body pair */
#define SG_SPHERE_SPHERE_TELEPORTED 19           /* TeleportedSphereSphereConstraint{ i, j, x0, x1, ri, rj }: p = q0_i + ri/(ri+rj) (x1 - x0) */
#define SG_KINEMATIC_OBJECT_SPHERE_TELEPORTED 30 /* KinematicObjectSphereConstraint{ i, ri, n, j, X, 0, 0 }: i = free sphere, j = kinematic one, p = X = its teleported centre */

/* rigidbody2d (the constraint classes under rigidbody2d/) */
#define SG_CIRCLE_CIRCLE 20    /* CircleCircleConstraint{ i, j, n, p, ri, rj } */
#define SG_KINEMATIC_CIRCLE 21 /* KinematicObjectCircleConstraint: i = free circle, j = kinematic body, p = its position at q0 */
#define SG_BODY_BODY_2D 22     /* BodyBodyConstraint{ i, j, p, n, q0 } (box-box, circle-box) */
#define SG_PLANE_CIRCLE 23     /* StaticPlaneCircleConstraint: j = plane */
#define SG_PLANE_BODY_2D 24    /* StaticPlaneBodyConstraint: j = plane, aux = corner 0..3, p = body-space arm of the corner */
/* with portals (rigidbody2d/RigidBody2DSim.cpp:477-636): after the contacts of the un-teleported pairs, ascending (i,j) */
#define SG_CIRCLE_CIRCLE_TELEPORTED 25      /* TeleportedCircleCircleConstraint{ i, j, x0, x1, ri, rj, delta0, delta1, ri, rj }: p = q0_i + ri/(ri+rj) (x1 - x0) */
#define SG_CIRCLE_CIRCLE_KICK_TELEPORTED 26 /* KinematicKickCircleCircleConstraint{ i, j, x0, x1, ri, rj, kick } (Lees-Edwards portal) */
/* RigidBody2DGeometryType */
#define SG_GEO2_CIRCLE_TYPE 0
#define SG_GEO2_BOX_TYPE 1

/* RigidBodyGeometryType (rigidbody3d/Geometry/RigidBodyGeometry.h:13-19) */
#define SG_GEO_BOX 0
#define SG_GEO_SPHERE 1
#define SG_GEO_STAPLE 2 /* not supported (not on the north-star path) */
#define SG_GEO_MESH 3
/* rigidbody2d geometry types: values of RigidBody2DGeometryType (rigidbody2d/RigidBody2DGeometry.h) */
#define SG_GEO2D_CIRCLE 0
#define SG_GEO2D_BOX 1

/* which optional arrays sg_*_active_set copies back to the host */
#define SG_OUT_NORMALS 1u
#define SG_OUT_POINTS 2u
#define SG_OUT_DEPTHS 4u
#define SG_OUT_CANDIDATES 8u
#define SG_OUT_ALL 15u
/* Input flag for sg_ball2d_active_set / sg_rb2d_active_set / sg_rb3d_active_set (or-ed into out_flags): q0 and q1 are
   exactly the input and the output of the last sg_*_flow on this context -- as in ImpactMap::flow, which passes the unconstrained map's (q0, q1) straight
   to computeActiveSet (scisim/ConstrainedMaps/ImpactMaps/ImpactMap.cpp:54-58) -- so they are not uploaded again.
   The pointers may then be NULL.  SG_ERR_INVALID if no such flow preceded the call. */
#define SG_IN_RESIDENT 256u

typedef struct sg_ctx sg_ctx;

/* Sorted candidate pairs: drop-in for std::set<std::pair<unsigned,unsigned>> filled by
 * SpatialGridDetector::getPotentialOverlaps (ball2d/SpatialGridDetector.h:39). ij = [i0,j0,i1,j1,...],
 * i<j, ascending lexicographic -- the std::set iteration order. */
typedef struct sg_pairs
{
  uint64_t n;
  const uint32_t* ij;
} sg_pairs;

/* Active set in REFERENCE ORDER (one entry per Constraint the reference would emplace_back):
 * ball-ball ascending (i,j) | drums drum-major | planes plane-major. dim = 2 or 3. */
typedef struct sg_contacts
{
  uint32_t dim;
  uint64_t n_candidates; /* size of the broad-phase pair set (P_c) */
  uint64_t n_active;     /* = n_body_body + n_drum + n_plane */
  uint64_t n_body_body;
  uint64_t n_drum;
  uint64_t n_plane;
  const uint32_t* type;  /* n_active: SG_BALL_BALL, ... */
  const uint32_t* i;     /* n_active: (first) body index */
  const uint32_t* j;     /* n_active: second body / drum / plane index */
  const double* n;       /* dim*n_active constraint normal (body-body, drum: from q0; plane: plane normal) */
  const double* p;       /* dim*n_active getWorldSpaceContactPoint( q0 ) */
  const double* depth;   /* n_active penetrationDepth( q1 ) (NaN where the reference has no override) */
  const uint32_t* cand_ij; /* 2*n_candidates, only with SG_OUT_CANDIDATES */
  const uint32_t* aux;   /* n_active (rigidbody3d only): corner / hull vertex number for plane-box / plane-body */
} sg_contacts;

/* What the portal path of ball2d adds to an active set (sg_ball2d_teleported).  Portal words hold the portal index, with
 * bit 31 set when the ball went through plane B (TeleportedBall::planeIndex() == 1); 0xffffffff = that ball was not
 * teleported.  The teleported contacts are entries [n_regular, n_regular + n_teleported) of the sg_contacts arrays; for
 * each, x0 / x1 are the teleported centres at q0 and kick the kinematic kick -- the arguments generateTeleportedBallBallCollision
 * (ball2d/Ball2DSim.cpp:653-728) passes to the constraint constructors.  Candidate pairs of that active set index the
 * extended box list: i >= n is teleported box i - n of this table. */
typedef struct sg_teleported
{
  uint64_t n_boxes;           /* teleported AABBs appended after the n real ones, portal-major (Ball2DSim.cpp:390-413) */
  const uint32_t* box_body;   /* n_boxes: TeleportedBall::bodyIndex() */
  const uint32_t* box_portal; /* n_boxes: portal word */
  uint64_t n_regular;         /* ball-ball contacts between un-teleported balls */
  uint64_t n_teleported;      /* teleported contacts (n_body_body = n_regular + n_teleported) */
  const uint32_t* portal0;    /* n_teleported: portal word of body i */
  const uint32_t* portal1;    /* n_teleported: portal word of body j */
  const double* x0;           /* dim * n_teleported (dim = 2; 3 for rigidbody3d) */
  const double* x1;           /* dim * n_teleported */
  const double* kick;         /* 2 n_teleported (0 for SG_BALL_BALL_TELEPORTED / SG_CIRCLE_CIRCLE_TELEPORTED) */
  const double* delta0;       /* 2 n_teleported, rigidbody2d only (NULL for ball2d): x0 - q0_i, TeleportedCircleCircleConstraint's delta0 */
  const double* delta1;       /* 2 n_teleported, rigidbody2d only: x1 - q0_j (both NaN for kinematic-kick contacts, as the reference stores) */
} sg_teleported;

/* ---- context ------------------------------------------------------------------------------------- */
int sg_create( sg_ctx** ctx, int device );
void sg_destroy( sg_ctx* ctx );
const char* sg_last_error( const sg_ctx* ctx ); /* ctx may be NULL: error of a failed sg_create */
int sg_synchronize( sg_ctx* ctx );
/* pinned host memory for the caller's q/v vectors (optional; pageable pointers work, slower) */
int sg_host_alloc( sg_ctx* ctx, uint64_t bytes, void** ptr );
int sg_host_free( sg_ctx* ctx, void* ptr );
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on */
void* sg_stream( sg_ctx* ctx );

/* per-kernel device timing (CUDA events on the context's stream; off by default) */
int sg_profile_enable( sg_ctx* ctx, int on );
int sg_profile_reset( sg_ctx* ctx );
int sg_profile_count( sg_ctx* ctx );
/* entry k: kernel name, number of launches, summed device milliseconds, summed algorithmic bytes */
int sg_profile_get( sg_ctx* ctx, int k, const char** name, uint64_t* launches, double* ms, double* bytes );
/* number of kernels this context has launched since creation */
uint64_t sg_launch_count( const sg_ctx* ctx );
/* CUDA-event stopwatch on the context's stream (begin records, end records + waits + returns milliseconds) */
int sg_timer_begin( sg_ctx* ctx );
int sg_timer_end( sg_ctx* ctx, double* ms );
/* overwrite a 384 MB scratch buffer on the context's stream so the next kernels start with a cold L2 */
int sg_flush_l2( sg_ctx* ctx );

/* ---- broad phase alone ---------------------------------------------------------------------------
 * Replaces SpatialGridDetector::getPotentialOverlaps( aabbs, overlaps ) -- ball2d/SpatialGridDetector.cpp:106-133,
 * rigidbody2d/SpatialGrid.cpp:114-141 (dim 2), rigidbody3d/SpatialGridDetector.cpp:110-137 (dim 3).
 * aabbs: n boxes, each [lo(dim), hi(dim)]. */
int sg_candidate_pairs( sg_ctx* ctx, int dim, uint32_t n, const double* aabbs, sg_pairs* out );

/* ---- ball2d ----------------------------------------------------------------------------------------
 * Static data of Ball2DState (ball2d/Ball2DState.h): radii, per-ball mass (Minv = 1.0/m as
 * Ball2DState.cpp:54-66), gravity (Ball2DGravityForce), static planes (normal is normalised here exactly as
 * ball2d/StaticGeometry/StaticPlane.cpp:10-14 does) and static drums. */
int sg_ball2d_set_bodies( sg_ctx* ctx, uint32_t n, const double* r, const double* m );
int sg_ball2d_set_gravity( sg_ctx* ctx, const double* g /* 2 */ );
int sg_ball2d_set_planes( sg_ctx* ctx, uint32_t n, const double* x /* 2n */, const double* nrm /* 2n */ );
int sg_ball2d_set_drums( sg_ctx* ctx, uint32_t n, const double* x /* 2n */, const double* r /* n */ );

/* Planar and Lees-Edwards portals (ball2d/Portals/PlanarPortal.h; at most 8).  Per portal: the two StaticPlanes (point and
 * normal, normalised here as StaticPlane's constructor does), the tangential velocity v (0 = plain periodic portal) and the
 * bounds of the periodic tangential coordinate (0 for plain portals), as PlanarPortal( plane_a, plane_b, velocity, bounds ).
 * Once portals are set, sg_ball2d_active_set / sg_ball2d_step follow Ball2DSim::computeBallBallActiveSetSpatialGridWithPortals
 * (ball2d/Ball2DSim.cpp:368-546): boxes at q1 only, BallBallConstraint::isActive at q1 instead of CCD, teleported copies of the
 * balls that touch a portal plane, SG_BALL_BALL_TELEPORTED / SG_BALL_BALL_KICK_TELEPORTED contacts after the regular ones.
 * SG_ERR_UNSUPPORTED where the reference exits (a ball touching both planes of one portal, PlanarPortal.cpp:117-121).
 * n = 0 removes the portals.  Not available in slab mode.
 *   update_portals   Ball2DSim::updatePeriodicBoundaryConditionsStartOfStep (Ball2DSim.cpp:327-334) with t = next_iteration * dt;
 *                    dx_out (optional, n values) receives each portal's tangential offset
 *   enforce_portals  Ball2DSim::enforcePeriodicBoundaryConditions (Ball2DSim.cpp:336-366) on host vectors q, v (2n each), in place
 *   teleported       details of the last active set computed with portals (pinned memory, valid until the next call) */
int sg_ball2d_set_portals( sg_ctx* ctx, uint32_t n, const double* plane_a_x /* 2n */, const double* plane_a_n /* 2n */, const double* plane_b_x /* 2n */, const double* plane_b_n /* 2n */,
                           const double* v /* n */, const double* bounds /* n */ );
int sg_ball2d_update_portals( sg_ctx* ctx, double t, double* dx_out );
int sg_ball2d_enforce_portals( sg_ctx* ctx, double* q, double* v );
int sg_ball2d_teleported( sg_ctx* ctx, sg_teleported* out );

/* UnconstrainedMap::flow( q0, v0, fsys, iteration, dt, q1, v1 ) -- scisim/UnconstrainedMaps/UnconstrainedMap.h:33
 * for ball2d's SymplecticEulerMap / VerletMap with the gravity force set above. q1/v1 are caller-sized (2n). */
int sg_ball2d_flow( sg_ctx* ctx, int map_kind, const double* q0, const double* v0, double dt, double* q1, double* v1 );

/* ConstrainedSystem::computeActiveSet( q0, qp, v, active_set ) -- scisim/Constraints/ConstrainedSystem.h:20 as
 * implemented by Ball2DSim::computeActiveSet (ball2d/Ball2DSim.cpp:151-173; with portals set: the portal branch, see above). */
int sg_ball2d_active_set( sg_ctx* ctx, const double* q0, const double* q1, uint32_t out_flags, sg_contacts* out );

/* Resident variants: state stays in HBM between calls (what `value` in bench.py times).
 * upload: q,v -> device q0,v0.  step: flow(q0,v0)->(q1,v1) then active set on (q0,q1); nothing is copied
 * to the host except the counts in *out (array pointers are NULL).  fetch: copy the last step's lists.
 * The step reports n_candidates, n_body_body and n_active only: splitting the static contacts into n_drum / n_plane needs one
 * more device read-back, which the step avoids -- sg_ball2d_step / sg_ball2d_slab_detect leave both at 0, sg_ball2d_fetch
 * fills them (n_active = n_body_body + n_drum + n_plane holds for fetch and for sg_ball2d_active_set). */
int sg_ball2d_upload( sg_ctx* ctx, const double* q, const double* v );
int sg_ball2d_step( sg_ctx* ctx, int map_kind, double dt, sg_contacts* out );
int sg_ball2d_fetch( sg_ctx* ctx, uint32_t out_flags, double* q1, double* v1, sg_contacts* out );
/* q1, v1 (either may be NULL) as the last flow / step / sg_ball2d_slab_flow on this context left them on the device (slab mode: the owned block) */
int sg_ball2d_fetch_state( sg_ctx* ctx, double* q1, double* v1 );

/* Device-side assembly for the solver's first step, from the contact arrays of the LAST active set on this context:
 *   N      ImpactOperatorUtilities::computeN (scisim/ConstrainedMaps/ImpactMaps/ImpactOperatorUtilities.cpp:10-48) with the constraints'
 *          evalgradg (ball2d/Constraints/BallBallConstraint.cpp:88-100, BallStaticPlaneConstraint.cpp:65-73, BallStaticDrumConstraint.cpp:59-66),
 *          zeros pruned; n_dofs x n_constraints, compressed column-major (outer n_constraints + 1, inner / values n_nnz)
 *   Q      N^T * Minv * N (ImpactMap.cpp:106-110): n_constraints^2, compressed column-major with sorted rows, Eigen's accumulation order
 *   bases  computeContactBases (ball2d/Ball2DSim.cpp:188-201): 4 doubles per constraint, the column-major 2x2 [ n | t ]
 * Pointers go into library-owned pinned host memory, valid until the next call on the context.  Contact types 0..2 only. */
#define SG_ASM_N 1u
#define SG_ASM_Q 2u
#define SG_ASM_BASES 4u
typedef struct sg_assembly
{
  uint64_t n_constraints;
  uint64_t n_dofs;
  uint64_t n_nnz;
  const int32_t* n_outer;
  const int32_t* n_inner;
  const double* n_values;
  uint64_t q_nnz;
  const int32_t* q_outer;
  const int32_t* q_inner;
  const double* q_values;
  const double* bases;
} sg_assembly;
int sg_ball2d_assemble( sg_ctx* ctx, uint32_t flags, sg_assembly* out );
/* ConstraintCache (ball2d/ConstraintCache.cpp:20-122) on the device, as a sorted-key join:
 *   store   cacheConstraint for every constraint of the current active set: r = ncomp values per constraint, active-set order (host)
 *   lookup  getCachedConstraintImpulse for every constraint of the current active set: the stored values where the key
 *           ( (i,j) | (plane,ball) | (drum,ball) ) was cached, zeros otherwise; *hits = constraints found
 *   clear   clearConstraintCache */
int sg_ball2d_cache_store( sg_ctx* ctx, uint32_t ncomp, const double* r );
int sg_ball2d_cache_lookup( sg_ctx* ctx, uint32_t ncomp, double* r_out, uint64_t* hits );
int sg_ball2d_cache_clear( sg_ctx* ctx );

/* State I/O at the seam: Ball2DState's binary snapshot (ball2d/Ball2DState.cpp:259-312, scisim/Utilities.h:43-94,
   scisim/Math/MathUtilities.h:42-60, MathUtilities.cpp:142-177), byte for byte, written from / read into the device-resident state.
   serialize   which = 0: ( q0, v0 ) as uploaded, 1: ( q1, v1 ) of the last flow / step.  buf = NULL: *bytes <- size needed
   deserialize configures bodies, gravity, planes, drums and portals from a snapshot and uploads ( q, v ) */
int sg_ball2d_state_serialize( sg_ctx* ctx, int which, void* buf, uint64_t cap, uint64_t* bytes );
int sg_ball2d_state_deserialize( sg_ctx* ctx, const void* buf, uint64_t bytes );

/* Peer-memory halo exchange (one process per GPU, neighbours' mailboxes mapped over NVLink with CUDA IPC; no
   collective, no host round trip).  No reference counterpart (SCISim is single-process); see DESIGN.md section 5.
   sg_ball2d_slab_mailbox    allocates this rank's mailbox (once) and returns its device address and/or its CUDA IPC
                             handle (64 opaque bytes) for the neighbours.
   sg_ball2d_slab_connect    maps the mailbox of the neighbour on `side` (0 = lower ranks, 1 = higher): either from
                             its IPC handle (other process) or from its address (same process; peer_device = its
                             device ordinal, -1 if the same device).  Exactly one of the two must be given.
   After both calls sg_ball2d_slab_flow also posts this rank's swept interval to the neighbours (interval_dev may
   then be NULL) and sg_ball2d_slab_exchange( phase ) replaces pack / send / recv / unpack:
     phase 1  wait for each neighbour's interval, pack the bodies that reach it straight into its mailbox, raise its flag
     phase 2  wait for the neighbours' halos, move them into the ghost slots
     phase 0  both.  Everything is asynchronous on the context's stream; waits give up after 10 s and
              sg_ball2d_slab_detect then returns SG_ERR_INTERNAL. */
int sg_ball2d_slab_mailbox( sg_ctx* ctx, void** mailbox_dev, void* ipc_handle_64 );
int sg_ball2d_slab_connect( sg_ctx* ctx, int side, const void* ipc_handle_64, void* same_process_mailbox, int peer_device );
int sg_ball2d_slab_exchange( sg_ctx* ctx, int phase );
/* drops the mailbox and the neighbour mappings: the context is back on the pack / unpack calls (collective transport) */
int sg_ball2d_slab_disconnect( sg_ctx* ctx );

/* Slab mode (multi-GPU): this context holds ONE spatial slab of a larger scene -- n_owned bodies of arbitrary global
 * numbering, stored in ascending global index -- plus per-step ghost copies of the neighbouring slabs' bodies.  The reference has no
 * distributed mode; these calls are driven by sg_multi (one process, N GPUs: below) or by scisim_b200/slab.py (one process per GPU).
 *   init      reserve ghost_cap slots on either side of the owned block; upload r, m of the owned bodies; the owned bodies get the
 *             global indices gid_first, gid_first + 1, ... until set_gids says otherwise
 *   set_gids  gid_owned (n_owned values, strictly ascending): global index of every owned body; x_limits (2 values, either may be
 *             NULL): the x-range the owned bodies' swept boxes must stay inside.  A body outside could touch a body of a
 *             NON-neighbouring slab, which no halo carries: detect then returns SG_ERR_REBALANCE (re-partition, repeat the step)
 *   (sg_ball2d_upload / sg_ball2d_fetch then address the owned block only; sg_ball2d_slab_upload_q1 supplies an externally
 *   computed q1 for the owned block, to be followed by flow( SG_MAP_NONE ))
 *   flow      integrate the owned bodies; *interval_dev (DEVICE pointer, 2 doubles) receives [min lo.x, max hi.x] of their
 *             swept AABBs -- what the other ranks need to select the ghosts they owe this one
 *   pack      list of the owned bodies whose swept AABB overlaps the x-interval at interval_dev (DEVICE) into
 *             send_dev (DEVICE, cap + 1 records of 48 bytes; record 0 is a header carrying the count, so the receiver learns
 *             it on the device); the count also goes to *count_dev (DEVICE).  send_dev = NULL: count only
 *   unpack    take a received buffer (same layout) as this step's ghosts from the neighbour on side 0 (lower x) or 1 (higher x)
 *   detect    broad + narrow phase over owned + ghosts; a pair is kept iff this rank owns the body with the smaller
 *             global index; planes / drums are tested for owned bodies only; indices in the lists are global and every list is
 *             ascending, so the per-rank lists merge into the reference's order (sg_slab_merge_dest).
 *             ghosts_out (optional, 2 values): how many ghosts arrived on each side.  Nothing before detect blocks the host. */
int sg_ball2d_slab_init( sg_ctx* ctx, uint32_t n_owned, uint32_t gid_first, uint32_t ghost_cap, const double* r, const double* m );
int sg_ball2d_slab_set_gids( sg_ctx* ctx, const uint32_t* gid_owned, const double* x_limits );
int sg_ball2d_slab_upload_q1( sg_ctx* ctx, const double* q1 );
int sg_ball2d_slab_flow( sg_ctx* ctx, int map_kind, double dt, double* interval_dev );
int sg_ball2d_slab_pack( sg_ctx* ctx, const double* interval_dev, void* send_dev, uint32_t cap, uint32_t* count_dev );
int sg_ball2d_slab_unpack( sg_ctx* ctx, int side, const void* recv_dev );
int sg_ball2d_slab_detect( sg_ctx* ctx, sg_contacts* out, uint32_t* ghosts_out );
/* diagnostics (4 values): ghosts held on side 0 / 1, steps since the mailbox was created in which the halo pack had to scan all bodies
   (the candidate band did not hold: the first step always, later ones only after a jump), capacity of the candidate lists */
int sg_ball2d_slab_stats( sg_ctx* ctx, uint32_t* out4 );

/* ---- multi-GPU: host helpers of the slab decomposition (no device work, no context) -----------------------------
 * sg_slab_partition   equal-count x-quantiles: bodies ranked by ( x, index ), slab k takes ranks
 *                     [ k n / world, (k+1) n / world ).  x[ i * x_stride ] = x of body i (ball2d: q, stride 2).
 *                     rank_of[n] <- slab of every body; cuts[world + 1] (may be NULL) <- scene extent and the world - 1 cut positions
 * sg_slab_limits      the x-range bodies of slab `rank` must stay inside (sg_ball2d_slab_set_gids): the slab widened by half the
 *                     width of each neighbouring slab, so that bodies of slabs k and k + 2 can never touch
 * sg_slab_merge_dest  per-slab lists, each ascending in its first index and with every body's entries in ONE slab (the owner of
 *                     the pair's lower index), merge body by body: dest[k][e] <- position of entry e of slab k in the merged list,
 *                     which is the reference's ascending (i,j) order (the std::set of ball2d/Ball2DSim.cpp:580).
 *                     first[k][e * stride] = first index of the entry (stride 2 for a candidate list, 1 for the i column)
 * sg_slab_merge_static_dest  the same for the static contacts behind them: ordered by ( type, j = geometry, i = body ),
 *                     i.e. drums drum-major then planes plane-major, body ascending (Ball2DSim.cpp:735-761) */
int sg_slab_partition( uint32_t n, const double* x, uint32_t x_stride, uint32_t world, uint32_t* rank_of, double* cuts );
int sg_slab_limits( uint32_t world, const double* cuts, uint32_t rank, double* limits /* 2 */ );
int sg_slab_merge_dest( uint32_t n_bodies, uint32_t n_parts, const uint32_t* const* first, uint32_t stride, const uint64_t* len, uint64_t* const* dest );
int sg_slab_merge_static_dest( uint32_t n_parts, const uint32_t* const* type, const uint32_t* const* i, const uint32_t* const* j, const uint64_t* len, uint64_t* const* dest );

/* ---- multi-GPU in ONE process (what a SCISim process links against to use several GPUs; SURVEY.md 8b sketched it as
 * sg_create( ctx, n_gpus, devices )).  sg_multi owns one context and one worker thread per GPU, partitions the scene into
 * x-slabs, and offers the calls of a single ball2d context over GLOBAL vectors and indices:
 *   set_bodies / set_gravity / set_planes / set_drums     as sg_ball2d_set_*
 *   flow        UnconstrainedMap::flow over host vectors (q0, v0 are partitioned and uploaded, q1, v1 come back)
 *   active_set  ConstrainedSystem::computeActiveSet: merged lists in the reference's order, pointers into memory owned by
 *               sg_multi, valid until the next call.  SG_IN_RESIDENT: (q0, q1) of the flow just done (ImpactMap.cpp:54-58)
 *   upload / step / fetch   the resident variant (step leaves everything on the GPUs, *out carries the counts only)
 * The scene is (re-)partitioned at the first upload, whenever a slab reports SG_ERR_REBALANCE or runs out of ghost slots (the
 * step is then repeated on the new partition, invisibly to the caller), and every `every_n_uploads` uploads if set.
 * devices = NULL: GPUs 0 .. n_gpus - 1.  The same device may be listed more than once (testing on one GPU). */
typedef struct sg_multi sg_multi;
int sg_create_multi( sg_multi** m, int n_gpus, const int* devices );
void sg_destroy_multi( sg_multi* m );
const char* sg_multi_last_error( const sg_multi* m ); /* m may be NULL: error of a failed sg_create_multi */
int sg_multi_n_gpus( const sg_multi* m );
sg_ctx* sg_multi_context( sg_multi* m, int k );       /* slab k's context (timers, profile, launch count) */
int sg_multi_set_rebalance( sg_multi* m, uint32_t every_n_uploads, uint32_t ghost_cap /* 0: chosen from the slab size */ );
int sg_multi_partition_info( sg_multi* m, double* cuts /* n_gpus + 1 */, uint32_t* n_owned /* n_gpus */, uint32_t* ghosts /* 2 n_gpus */, uint64_t* n_partitions );
int sg_multi_ball2d_set_bodies( sg_multi* m, uint32_t n, const double* r, const double* mass );
int sg_multi_ball2d_set_gravity( sg_multi* m, const double* g /* 2 */ );
int sg_multi_ball2d_set_planes( sg_multi* m, uint32_t n, const double* x /* 2n */, const double* nrm /* 2n */ );
int sg_multi_ball2d_set_drums( sg_multi* m, uint32_t n, const double* x /* 2n */, const double* r /* n */ );
int sg_multi_ball2d_flow( sg_multi* m, int map_kind, const double* q0, const double* v0, double dt, double* q1, double* v1 );
int sg_multi_ball2d_active_set( sg_multi* m, const double* q0, const double* q1, uint32_t out_flags, sg_contacts* out );
int sg_multi_ball2d_upload( sg_multi* m, const double* q, const double* v );
int sg_multi_ball2d_step( sg_multi* m, int map_kind, double dt, sg_contacts* out );
int sg_multi_ball2d_fetch( sg_multi* m, uint32_t out_flags, double* q1, double* v1, sg_contacts* out );

/* ---- rigidbody2d --------------------------------------------------------------------------------------
 * q = v-layout [x, y, theta] per body (rigidbody2d/RigidBody2DState.h). M = the 3N diagonal of the mass matrix
 * (m, m, I per body) exactly as FlowableSystem::M().valuePtr() holds it; Minv = 1.0 / M (RigidBody2DState.cpp:31-43).
 * Plane normals are used as given (RigidBody2DStaticPlane does not normalise). */
/* type: SG_GEO2D_CIRCLE (r) or SG_GEO2D_BOX (half-widths) per geometry entry */
int sg_rb2d_set_geometry( sg_ctx* ctx, uint32_t ngeo, const uint32_t* type, const double* r /* ngeo */, const double* half /* 2 ngeo */ );
int sg_rb2d_set_bodies( sg_ctx* ctx, uint32_t n, const uint32_t* geo_of_body, const uint8_t* fixed, const double* M /* 3n */ );
int sg_rb2d_set_gravity( sg_ctx* ctx, const double* g /* 2 */ );
int sg_rb2d_set_planes( sg_ctx* ctx, uint32_t n, const double* x /* 2n */, const double* nrm /* 2n */ );
/* rigidbody2d/SymplecticEulerMap.cpp:15-38, VerletMap.cpp:27-55 */
int sg_rb2d_flow( sg_ctx* ctx, int map_kind, const double* q0, const double* v0, double dt, double* q1, double* v1 );
/* RigidBody2DSim::computeActiveSet (rigidbody2d/RigidBody2DSim.cpp:696-714; with portals set: the portal branch, see below); SG_ERR_UNSUPPORTED where the
 * reference exits (kinematic box-box, kinematic circle vs box: RigidBody2DSim.cpp:186-190, 210-214) */
int sg_rb2d_active_set( sg_ctx* ctx, const double* q0, const double* q1, uint32_t out_flags, sg_contacts* out );
/* resident variants, as for ball2d and rigidbody3d: upload (q, v) once, step = flow + active set on the device copies ( only the
 * list sizes come back in *out ), fetch copies q1, v1 and the lists the flags ask for */
int sg_rb2d_upload( sg_ctx* ctx, const double* q, const double* v );
int sg_rb2d_step( sg_ctx* ctx, int map_kind, double dt, sg_contacts* out );
int sg_rb2d_fetch( sg_ctx* ctx, uint32_t out_flags, double* q1, double* v1, sg_contacts* out );
/* State I/O at the seam: RigidBody2DState's binary snapshot (rigidbody2d/RigidBody2DState.cpp:485-556, scisim/Utilities.h:43-94,
   scisim/Math/MathUtilities.h:42-60, MathUtilities.cpp:142-177), byte for byte, written from / read into the device-resident state.
   serialize   which = 0: ( q0, v0 ) as uploaded, 1: ( q1, v1 ) of the last flow / step.  Portals are written with the offset ( m_dx ) of the
               last sg_rb2d_update_portals.  buf = NULL: *bytes <- size needed
   deserialize configures geometry, bodies, gravity, planes and portals (frames and offsets as stored) from a snapshot and uploads ( q, v );
               SG_ERR_UNSUPPORTED for a moving plane or a plane whose stored tangent is not ( -n.y, n.x ) */
int sg_rb2d_state_serialize( sg_ctx* ctx, int which, void* buf, uint64_t cap, uint64_t* bytes );
int sg_rb2d_state_deserialize( sg_ctx* ctx, const void* buf, uint64_t bytes );

/* Planar and Lees-Edwards portals of the 2-D rigid-body sim (rigidbody2d/PlanarPortal.h; at most 8), arguments as
 * sg_ball2d_set_portals except that plane normals are used as given (RigidBody2DStaticPlane does not normalise).  With portals
 * set, sg_rb2d_active_set follows RigidBody2DSim::computeBodyBodyActiveSetSpatialGridWithPortals (RigidBody2DSim.cpp:876-1040):
 * boxes at q1 (computeAABB, not the swept computeCollisionAABB), a teleported box per body whose box reaches a portal plane
 * (plane A first; the both-planes check of aabbTouchesPortal is debug-only), un-teleported pairs through the regular narrow
 * phase, SG_CIRCLE_CIRCLE_TELEPORTED / SG_CIRCLE_CIRCLE_KICK_TELEPORTED contacts after them, then the planes.
 * SG_ERR_UNSUPPORTED where the reference exits: a box in a teleported candidate (RigidBody2DSim.cpp:374-385), a kinematic
 * body in a teleported collision (:481-485).  update / enforce / teleported as for ball2d (RigidBody2DSim.cpp:832-874);
 * q, v are [x, y, theta] per body. */
int sg_rb2d_set_portals( sg_ctx* ctx, uint32_t n, const double* plane_a_x /* 2n */, const double* plane_a_n /* 2n */, const double* plane_b_x /* 2n */, const double* plane_b_n /* 2n */,
                         const double* v /* n */, const double* bounds /* n */ );
int sg_rb2d_update_portals( sg_ctx* ctx, double t, double* dx_out );
int sg_rb2d_enforce_portals( sg_ctx* ctx, double* q, double* v );
int sg_rb2d_teleported( sg_ctx* ctx, sg_teleported* out );

/* ---- rigidbody3d --------------------------------------------------------------------------------------
 * Layouts are RigidBody3DState's (rigidbody3d/RigidBody3DState.cpp:70-240): q = [3N centres | 9N row-major R],
 * v = [3N linear | 3N angular]. Static data: the geometry list (as m_geometry) and, per body, its geometry index,
 * the kinematically-scripted flag, total mass and body-frame inertia (the diagonals of M0).  The world-space mass
 * matrix M = R diag(I0) R^T the reference keeps in sync with q (updateMandMinv) is recomputed on the device from
 * R(q0) with the same evaluation order. */
int sg_rb3d_set_geometry( sg_ctx* ctx, uint32_t ngeo, const uint32_t* type, const double* r /* ngeo */, const double* half /* 3 ngeo */, const uint32_t* mesh /* ngeo */ );
/* a triangle mesh as RigidBodyTriangleMesh holds it (rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp:54-102): all vertices (AABB),
 * surface samples, convex-hull vertices, and the signed distance grid (x fastest). Returns its index in *mesh_index. */
int sg_rb3d_add_mesh( sg_ctx* ctx, uint32_t nverts, const double* verts, uint32_t nsamples, const double* samples, uint32_t nhull, const double* hull,
                      const double* cell_delta, const uint32_t* dims, const double* origin, const double* sdf, uint32_t* mesh_index );
int sg_rb3d_set_bodies( sg_ctx* ctx, uint32_t n, const uint32_t* geo_of_body, const uint8_t* fixed, const double* m /* n */, const double* I0 /* 3n */ );
int sg_rb3d_set_gravity( sg_ctx* ctx, const double* g /* 3 */ );
/* normals are normalised as rigidbody3d/StaticGeometry/StaticPlane.cpp:10-15 does */
int sg_rb3d_set_planes( sg_ctx* ctx, uint32_t n, const double* x /* 3n */, const double* nrm /* 3n */ );
/* static cylinders (rigidbody3d/StaticGeometry/StaticCylinder.cpp:8-19: axis normalised; at most 8): bodies live INSIDE them.
 * Contacts follow the plane contacts, cylinder-major (RigidBody3DSim::computeBodyCylinderActiveSetAllPairs,
 * RigidBody3DSim.cpp:1504-1557): spheres and mesh convex-hull vertices; a non-kinematic box makes sg_rb3d_active_set /
 * sg_rb3d_step return SG_ERR_UNSUPPORTED where the reference prints and exits. */
int sg_rb3d_set_cylinders( sg_ctx* ctx, uint32_t n, const double* x /* 3n */, const double* axis /* 3n */, const double* r /* n */ );
/* Planar portals of the 3-D rigid-body sim (rigidbody3d/Portals/PlanarPortal.h; at most 8): per portal plane A and B as
 * (point, normal; the normal is normalised and the tangents t0, t1 are built as StaticPlane does, FromTwoVectors( UnitY, n ) * UnitX / UnitZ
 * -- a normal within 1e-12 of -UnitY takes Eigen's SVD branch and is rejected with SG_ERR_UNSUPPORTED) and the three integer
 * multipliers of PlanarPortal( plane_a, plane_b, portal_multiplier ).  With portals set, sg_rb3d_active_set / sg_rb3d_step follow
 * RigidBody3DSim::computeActiveSetBodyBodySpatialGrid with its portal loop (RigidBody3DSim.cpp:1072-1260): a teleported box per body
 * whose box reaches a portal plane, un-teleported pairs through the regular narrow phase, then SG_SPHERE_SPHERE_TELEPORTED /
 * SG_KINEMATIC_OBJECT_SPHERE_TELEPORTED contacts, then planes and cylinders.  The reference supports teleported collisions for
 * spheres only (everything else exits, RigidBody3DSim.cpp:1262-1292): this path is limited to all-sphere scenes and returns
 * SG_ERR_UNSUPPORTED otherwise.  enforce_portals: RigidBody3DSim::enforcePeriodicBoundaryConditions (:642-663) on a host q
 * ( centres of mass only ), in place; teleported: as sg_ball2d_teleported with 3 doubles per point ( kick, delta0, delta1 = NULL ). */
int sg_rb3d_set_portals( sg_ctx* ctx, uint32_t n, const double* plane_a_x /* 3n */, const double* plane_a_n /* 3n */, const double* plane_b_x /* 3n */, const double* plane_b_n /* 3n */,
                         const int32_t* multiplier /* 3n */ );
int sg_rb3d_enforce_portals( sg_ctx* ctx, double* q );
int sg_rb3d_teleported( sg_ctx* ctx, sg_teleported* out );
/* RigidBody3DState::updateMandMinv (rigidbody3d/RigidBody3DState.cpp:428-462), the last thing RigidBody3DSim::flow does to the
 * state: per body I = R I0 R^T and Iinv = R Iinv0 R^T, 9 doubles each, column-major -- exactly the runs
 * M.data().value( 3 N + 9 b ... ) / Minv.data().value( 3 N + 9 b ... ) of the reference's sparse matrices, so a shim passes
 * &M.data().value( 3 N ) and &Minv.data().value( 3 N ).  q: host vector [3N x | 9N R row-major], or NULL for the device copy of
 * q1 the last sg_rb3d_flow / sg_rb3d_step on this context left (no upload). */
int sg_rb3d_update_m_and_minv( sg_ctx* ctx, const double* q, double* m_blocks /* 9N */, double* minv_blocks /* 9N */ );
/* UnconstrainedMap::flow for SplitHamMap (SG_MAP_SPLIT_HAM) / DMVMap (SG_MAP_DMV) with NearEarthGravityForce; see SG_MAP_M_UPDATED */
int sg_rb3d_flow( sg_ctx* ctx, int map_kind, const double* q0, const double* v0, double dt, double* q1, double* v1 );
/* RigidBody3DSim::computeActiveSet (rigidbody3d/RigidBody3DSim.cpp:250-262; cylinders via sg_rb3d_set_cylinders; with portals set: the
 * portal loop of computeActiveSetBodyBodySpatialGrid, all-sphere scenes, see sg_rb3d_set_portals). Returns
 * SG_ERR_UNSUPPORTED where the reference exits on a pair of geometry types it cannot collide (RigidBody3DSim.cpp:960-961). */
int sg_rb3d_active_set( sg_ctx* ctx, const double* q0, const double* q1, uint32_t out_flags, sg_contacts* out );
/* resident variants, as for ball2d */
int sg_rb3d_upload( sg_ctx* ctx, const double* q, const double* v );
int sg_rb3d_step( sg_ctx* ctx, int map_kind, double dt, sg_contacts* out );
int sg_rb3d_fetch( sg_ctx* ctx, uint32_t out_flags, double* q1, double* v1, sg_contacts* out );
/* State I/O at the seam: RigidBody3DState's binary snapshot (rigidbody3d/RigidBody3DState.cpp:586-668, scisim/Utilities.h:43-94,
   scisim/Math/MathUtilities.h:42-60, MathUtilities.cpp:142-177), byte for byte, for scenes of spheres, boxes and triangle meshes (a mesh's snapshot holds its
   whole input file, RigidBodyTriangleMesh.cpp:215-232: attach it with sg_rb3d_set_mesh_snapshot; without it serialize returns SG_ERR_UNSUPPORTED).
   serialize   which = 0: ( q0, v0 ) as uploaded, 1: ( q1, v1 ) of the last flow / step.  m_updated = 0: the world-space blocks of M / Minv as
               RigidBody3DState's constructor stores them (a state that no flow has touched), 1: as updateMandMinv leaves them (see SG_MAP_M_UPDATED).
               buf = NULL: *bytes <- size needed
   deserialize configures geometry, bodies, gravity, planes, cylinders and portals from a snapshot and uploads ( q, v ) */
int sg_rb3d_state_serialize( sg_ctx* ctx, int which, int m_updated, void* buf, uint64_t cap, uint64_t* bytes );
/* A triangle mesh's own record for the snapshot: the bytes RigidBodyTriangleMesh::serialize writes (rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp:215-232, type byte
   included) for mesh `mesh_index` of sg_rb3d_add_mesh.  The record holds what only the caller has (file name, faces, volume, moments); the library checks that
   it is well formed and sized like the mesh, keeps it, and writes it back verbatim.  sg_rb3d_state_deserialize adds the meshes of a snapshot
   (RigidBodyTriangleMesh( std::istream& ), RigidBodyTriangleMesh.cpp:105-129) and keeps their records itself. */
int sg_rb3d_set_mesh_snapshot( sg_ctx* ctx, uint32_t mesh_index, const void* record, uint64_t bytes );
int sg_rb3d_state_deserialize( sg_ctx* ctx, const void* buf, uint64_t bytes );
/* Diagnostics for the mesh narrow phase: how many sample sweeps (one per direction of a mesh-mesh pair,
   MeshMeshUtilities.cpp:10-65) read the distance field from a TMA-staged shared-memory brick and how many read
   it directly from global memory, summed since the first mesh was added. */
int sg_rb3d_mesh_stats( sg_ctx* ctx, uint64_t* staged_sweeps, uint64_t* direct_sweeps );

/* Slab mode for rigidbody3d (multi-GPU, all-sphere scenes: BASELINE configs[3]) -- the ball2d slab calls for spheres, peer-memory exchange only.
 * Reference path being sharded: RigidBody3DSim::computeActiveSetBodyBodySpatialGrid (rigidbody3d/RigidBody3DSim.cpp:1072-1260) and
 * computeBodyPlaneActiveSetAllPairs (:1414-1502).
 *   init      after sg_rb3d_set_geometry: this rank's n_owned spheres (geometry index, mass, body-frame inertia, strictly ascending global
 *             indices), ghost_cap halo slots per neighbour, the x-range the owned spheres' boxes must stay inside (sg_slab_limits; NULL:
 *             unlimited).  Kinematically scripted bodies and portals are not supported in slab mode.  sg_rb3d_upload / sg_rb3d_fetch then
 *             address the owned bodies only ( q = [3 n_owned | 9 n_owned], v = [3 n_owned | 3 n_owned] )
 *   mailbox / connect / disconnect / exchange   as sg_ball2d_slab_*
 *   flow      integrate the owned bodies, post [min lo.x, max hi.x] of their boxes at q1 to the neighbours' mailboxes, list the halo candidates
 *   detect    broad + narrow phase over owned + ghosts; a pair is kept iff this rank owns the body with the smaller global index; planes and
 *             cylinders are tested for owned bodies only; indices are global, lists ascending.  SG_ERR_REBALANCE as for ball2d. */
int sg_rb3d_slab_init( sg_ctx* ctx, uint32_t n_owned, uint32_t ghost_cap, const uint32_t* geo_of_body, const double* m /* n_owned */, const double* I0 /* 3 n_owned */,
                       const uint32_t* gid_owned, const double* x_limits /* 2 or NULL */ );
int sg_rb3d_slab_mailbox( sg_ctx* ctx, void** mailbox_dev, void* ipc_handle_64 );
int sg_rb3d_slab_connect( sg_ctx* ctx, int side, const void* ipc_handle_64, void* same_process_mailbox, int peer_device );
int sg_rb3d_slab_disconnect( sg_ctx* ctx );
int sg_rb3d_slab_flow( sg_ctx* ctx, int map_kind, double dt );
int sg_rb3d_slab_exchange( sg_ctx* ctx, int phase );
int sg_rb3d_slab_detect( sg_ctx* ctx, sg_contacts* out, uint32_t* ghosts_out );

#ifdef __cplusplus
}
#endif

#endif
