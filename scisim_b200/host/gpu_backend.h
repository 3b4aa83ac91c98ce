// gpu_backend.h -- C++14 host shim over the C ABI (include/scisim_b200.h): the classes a SCISim maintainer links in so
// that Ball2DSim / RigidBody2DSim / RigidBody3DSim run their unconstrained flow and active-set computation on the GPU
// (INTEGRATION.md).
//
//   GpuBall2DBackend        owns the sg_ctx; static scene data in, contact records out (reference order)
//   GpuBall2DMultiBackend   the same calls over several GPUs of this process (sg_multi: x-slabs behind the interface)
//   GpuSymplecticEulerMap,
//   GpuVerletMap            UnconstrainedMap implementations registered beside ball2d's own maps
//                           (ball2d/Ball2DUtilities.cpp:37, ball2dutils/Ball2DSceneParser.cpp:624-631)
//   PairImpulseCache        flat sorted replacement for the three ConstraintCache classes (std::map keyed by pair)
//   GpuRigidBody3DBackend,  the same for rigidbody3d (SplitHam / DMV maps; spheres, boxes, meshes, planes) and
//   GpuRigidBody2DBackend   rigidbody2d (symplectic Euler / Verlet; circles, boxes, planes)
//   GpuSplitHamMap, GpuDMVMap, GpuRB2DSymplecticEulerMap, GpuRB2DVerletMap   their UnconstrainedMap faces
//
// Error convention of the reference: print to std::cerr and std::exit( EXIT_FAILURE ).
#ifndef SCISIM_B200_GPU_BACKEND_H
#define SCISIM_B200_GPU_BACKEND_H

#include "scisim_compat.h"

#include "../../include/scisim_b200.h"

#include <cstdint>
#include <utility>
#include <istream>
#include <ostream>
#include <vector>

// One entry per Constraint the reference would emplace_back, in the reference's order
struct GpuContact2D
{
  uint32_t type; // SG_BALL_BALL, SG_BALL_DRUM, SG_BALL_PLANE; with portals also SG_BALL_BALL_TELEPORTED, SG_BALL_BALL_KICK_TELEPORTED
  uint32_t i;    // ball (first ball for ball-ball)
  uint32_t j;    // second ball / drum / plane
  double n[2];
  double p[2];
  double depth;
};

// Constructor arguments of a teleported contact (ball2d/Ball2DSim.cpp:653-728): BallBallConstraint{ i, j, x0, x1, ri, rj, true }
// or KinematicKickBallBallConstraint{ i, j, x0, x1, ri, rj, kick, true }.  portal0 / portal1: portal index with bit 31 set
// when the ball went through plane B, 0xffffffff when that ball was not teleported.
struct GpuTeleportedContact2D
{
  uint32_t portal0, portal1;
  double x0[2], x1[2], kick[2];
  double delta0[2], delta1[2]; // rigidbody2d only (TeleportedCircleCircleConstraint's displacements); NaN otherwise
};

// The GPU maps integrate with the single gravity vector pushed to the device and never call fsys.computeForce.  The guard checks once
// per state -- on the first flow after the bodies or the gravity changed -- that the system's force IS that gravity:
// F = fsys.computeForce( q0, v0, t ) must equal 0 + m g on every translational degree of freedom and 0 on the rotational ones.
// Anything else (ball2d's PenaltyForce, a second force, another gravity) prints and exits, as the reference does for set-ups it does
// not support -- the scene must then keep the CPU map.
class GravityOnlyGuard final
{
public:
  enum Layout { BALL2D, RIGIDBODY2D, RIGIDBODY3D }; // [x y]*, [x y theta]*, [x y z]* then [wx wy wz]*
  void setMasses( const double* m, const unsigned n, const unsigned stride ) { m_mass.resize( n ); for( unsigned i = 0; i < n; ++i ) { m_mass[i] = m[std::size_t( i ) * stride]; } m_checked = false; }
  void setGravity( const double gx, const double gy, const double gz ) { m_g[0] = gx; m_g[1] = gy; m_g[2] = gz; m_checked = false; }
  void verify( FlowableSystem& fsys, const VectorXs& q0, const VectorXs& v0, const scalar& t, const Layout layout, const char* who );
  // masses and gravity from a state snapshot the library has accepted (the three *State::serialize layouts); returns the body count
  // *consumed <- the length of the state's snapshot inside buf (what follows -- the constraint cache in <Sim>::serialize's stream -- is not the state's)
  unsigned configureFromSnapshot( const Layout layout, const char* buf, const std::size_t bytes, const char* who, std::size_t* consumed = nullptr );
  const char* tryConfigureFromSnapshot( const Layout layout, const char* buf, const std::size_t bytes, unsigned& n, std::size_t* consumed );
private:
  std::vector<double> m_mass;
  double m_g[3] = { 0.0, 0.0, 0.0 };
  bool m_checked = false;
};

// what the deserializeState wrappers do around the library call: read the rest of the stream (its position goes to start), and afterwards put a seekable
// stream back to the first byte behind the state's `consumed` bytes
std::vector<char> sghReadRest( std::istream& input_stream, std::istream::pos_type& start );
void sghRewindBehindState( std::istream& input_stream, const std::istream::pos_type start, const std::size_t consumed );

class GpuBall2DBackend final
{
public:
  explicit GpuBall2DBackend( const int device = 0 );
  ~GpuBall2DBackend();
  GpuBall2DBackend( const GpuBall2DBackend& ) = delete;
  GpuBall2DBackend& operator=( const GpuBall2DBackend& ) = delete;

  // Ball2DState contents (ball2d/Ball2DState.h): radii, per-ball masses, gravity, planes (x, n), drums (X, R)
  void setBodies( const VectorXs& r, const VectorXs& m );
  void setGravity( const double gx, const double gy );
  void setPlanes( const std::vector<double>& x, const std::vector<double>& n );
  void setDrums( const std::vector<double>& x, const std::vector<double>& r );
  // Ball2DState::planarPortals() (ball2d/Portals/PlanarPortal.h): per portal plane A and B as (x, n), velocity, bounds.
  // With portals set computeActiveSet follows computeBallBallActiveSetSpatialGridWithPortals (ball2d/Ball2DSim.cpp:368-546).
  void setPortals( const std::vector<double>& plane_a_x, const std::vector<double>& plane_a_n, const std::vector<double>& plane_b_x, const std::vector<double>& plane_b_n,
                   const std::vector<double>& velocity, const std::vector<double>& bounds );
  // Ball2DSim::updatePeriodicBoundaryConditionsStartOfStep / enforcePeriodicBoundaryConditions (ball2d/Ball2DSim.cpp:327-366)
  void updatePeriodicBoundaryConditionsStartOfStep( const unsigned next_iteration, const scalar& dt );
  void enforcePeriodicBoundaryConditions( VectorXs& q, VectorXs& v );
  // After a computeActiveSet with portals: contacts [num_regular, num_regular + teleported.size()) are the teleported ones
  void teleportedContacts( std::vector<GpuTeleportedContact2D>& teleported, uint64_t* num_regular = nullptr );

  // UnconstrainedMap::flow for the two ball2d maps
  void flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 );
  // Ball2DSim::computeActiveSet (ball2d/Ball2DSim.cpp:151-173): contacts in active_set order
  // from_last_flow: (q0, q1) are the vectors the last flow() call read and wrote, unchanged since -- what
  // ImpactMap::flow hands to computeActiveSet (ImpactMap.cpp:54-58); they are then not uploaded a second time.
  void computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact2D>& contacts, uint64_t* num_candidates = nullptr, const bool from_last_flow = false );
  // SpatialGridDetector::getPotentialOverlaps (ball2d/SpatialGridDetector.h:39) on caller-built boxes [minx,miny,maxx,maxy]
  void getPotentialOverlaps( const std::vector<double>& aabbs, std::vector<std::pair<unsigned,unsigned>>& overlaps );

  // What the impact maps build first from the active set just computed (ImpactMap.cpp:100-110, Ball2DSim.cpp:188-201), assembled on the
  // device: N (zeros pruned), Q = N^T Minv N, contact bases -- compressed column-major, pointers valid until the next call (sg_ball2d_assemble)
  void assemble( const uint32_t flags, sg_assembly& out );
  // ConstraintCache (ball2d/ConstraintCache.cpp:20-122) for the whole active set at once: store the impulses (ncomp values per constraint,
  // active-set order), look up the warm start of the current active set (zeros where a constraint was not cached; returns the hits)
  void cacheStore( const unsigned ncomp, const VectorXs& r );
  uint64_t cacheLookup( const unsigned ncomp, VectorXs& r );
  void cacheClear();
  // Ball2DState::serialize / deserialize (ball2d/Ball2DState.cpp:259-312) from / into the device-resident state: what Ball2DSim::serialize writes for its
  // state.  from_last_flow: ( q1, v1 ) of the last flow / step, else ( q0, v0 ) as uploaded.
  void serializeState( std::ostream& output_stream, const bool from_last_flow = true ) const;
  void deserializeState( std::istream& input_stream );

  sg_ctx* context() { return m_ctx; }
  GravityOnlyGuard& forceGuard() { return m_guard; }

private:
  void check( const int rc, const char* what ) const;
  sg_ctx* m_ctx;
  unsigned m_nbodies;
  GravityOnlyGuard m_guard;
};

// Ball2DSim's hot path on several GPUs of ONE process (include/scisim_b200.h, sg_multi): same calls, global vectors and indices.
// The scene is cut into equal-count x-slabs behind the interface, re-cut when bodies migrate, and the per-slab lists come back merged
// in the reference's order -- the caller cannot tell it from GpuBall2DBackend except by the clock.  Portals are not supported here.
class GpuBall2DMultiBackend final
{
public:
  explicit GpuBall2DMultiBackend( const std::vector<int>& devices );
  ~GpuBall2DMultiBackend();
  GpuBall2DMultiBackend( const GpuBall2DMultiBackend& ) = delete;
  GpuBall2DMultiBackend& operator=( const GpuBall2DMultiBackend& ) = delete;
  void setBodies( const VectorXs& r, const VectorXs& m );
  void setGravity( const double gx, const double gy );
  void setPlanes( const std::vector<double>& x, const std::vector<double>& n );
  void setDrums( const std::vector<double>& x, const std::vector<double>& r );
  void flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 );
  void computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact2D>& contacts, uint64_t* num_candidates = nullptr, const bool from_last_flow = false );
  unsigned numGpus() const;
  uint64_t numPartitions(); // how often the scene has been (re-)partitioned so far
  sg_multi* handle() { return m_multi; }
  GravityOnlyGuard& forceGuard() { return m_guard; }
private:
  void check( const int rc, const char* what ) const;
  sg_multi* m_multi;
  GravityOnlyGuard m_guard;
};

class GpuSymplecticEulerMap final : public UnconstrainedMap
{
public:
  explicit GpuSymplecticEulerMap( GpuBall2DBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "symplectic_euler"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuBall2DBackend& m_backend;
};

class GpuVerletMap final : public UnconstrainedMap
{
public:
  explicit GpuVerletMap( GpuBall2DBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "verlet"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuBall2DBackend& m_backend;
};

// the same two maps over GpuBall2DMultiBackend
class GpuMultiSymplecticEulerMap final : public UnconstrainedMap
{
public:
  explicit GpuMultiSymplecticEulerMap( GpuBall2DMultiBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "symplectic_euler"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuBall2DMultiBackend& m_backend;
};

class GpuMultiVerletMap final : public UnconstrainedMap
{
public:
  explicit GpuMultiVerletMap( GpuBall2DMultiBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "verlet"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuBall2DMultiBackend& m_backend;
};

// ball2d/ConstraintCache.{h,cpp}: three std::map<std::pair<unsigned,unsigned>,VectorXs> (ball-ball, plane-ball,
// drum-ball).  The active set arrives sorted, so the cache is three flat arrays kept in key order: insertion appends
// (re-sorting only if a caller inserts out of order), lookup is a binary search, a miss zeroes r (ConstraintCache.cpp:122).
class PairImpulseCache final
{
public:
  void clear();
  bool empty() const;
  // kind selects one of the reference's keyed maps:
  //   ball2d       0 ball-ball (i,j)      1 plane-ball (plane, ball)    2 drum-ball (drum, ball)      (ball2d/ConstraintCache.cpp:20-122)
  //   rigidbody3d  0 sphere-sphere (i,j)  1 plane-sphere (plane, body)  2 cylinder-sphere (cyl, body)  3 kinematic sphere-sphere
  //                (rigidbody3d/ConstraintCache.cpp:14-72 -- like the reference, only sphere constraints can be cached)
  //   rigidbody2d  0 circle-circle (i,j)  1 plane-circle (plane, body)  2 body-body (i,j)  3 kinematic object - circle ( min, max of the two bodies )
  //                (rigidbody2d/ConstraintCache.cpp:14-60)
  // A key stored twice keeps its FIRST impulse, as std::map::insert does (box-box gives two body-body contacts per pair: both read the first one's).
  void cacheConstraint( const int kind, const unsigned a, const unsigned b, const VectorXs& r );
  void getCachedConstraint( const int kind, const unsigned a, const unsigned b, VectorXs& r ) const;
  // ConstraintCache::serialize / deserialize of the three sims, byte for byte (ball2d/ConstraintCache.cpp:125-174, rigidbody2d/ConstraintCache.cpp:140-191,
  // rigidbody3d/ConstraintCache.cpp:168-220): per keyed map, in the order the sim's class writes them, a size_t count and then, in ascending key order,
  // { unsigned first, unsigned second, Eigen::Index rows, rows doubles }.  <Sim>::serialize writes the state first and this after it.
  enum Sim { BALL2D = 0, RIGIDBODY2D = 1, RIGIDBODY3D = 2 };
  void serialize( const Sim sim, std::ostream& output_stream ) const;
  bool deserialize( const Sim sim, std::istream& input_stream ); // false on a truncated or malformed stream (the cache is then empty)
private:
  struct Table
  {
    std::vector<uint64_t> keys;
    std::vector<double> values;
    unsigned width = 0;
    bool sorted = true;
    void sortIfNeeded();
  };
  mutable Table m_tables[4];
};

// PairImpulseCache behind plain C calls (what the parity tests bind with ctypes to compare it with the reference's three ConstraintCache.cpp)
extern "C"
{
void* sgh_cache_create();
void sgh_cache_destroy( void* cache );
void sgh_cache_clear( void* cache );
int sgh_cache_empty( const void* cache );
void sgh_cache_store( void* cache, int kind, unsigned a, unsigned b, const double* r, unsigned ncomp );
void sgh_cache_lookup( const void* cache, int kind, unsigned a, unsigned b, double* r, unsigned ncomp );
uint64_t sgh_cache_serialize( const void* cache, int sim, void* buf, uint64_t cap );        /* returns the length; written when it fits cap */
int sgh_cache_deserialize( void* cache, int sim, const void* buf, uint64_t bytes );          /* 1 ok, 0 malformed */
/* the length of the state snapshot ( sim 0 Ball2DState, 1 RigidBody2DState, 2 RigidBody3DState ) at the head of buf -- where the constraint cache starts in a
   <Sim>::serialize stream; 0 = not a state snapshot */
uint64_t sgh_state_snapshot_length( int sim, const void* buf, uint64_t bytes );
}

// ---- rigidbody3d ---------------------------------------------------------------------------------------------------
// One entry per Constraint RigidBody3DSim::computeActiveSet would emplace_back (RigidBody3DSim.cpp:250-262), same order
struct GpuContact3D
{
  uint32_t type; // SG_SPHERE_SPHERE ... SG_PLANE_BODY (constructor table in INTEGRATION.md section 3)
  uint32_t i;    // (first) body
  uint32_t j;    // second body / plane
  uint32_t aux;  // box corner / convex-hull vertex number for plane-box / plane-body
  double n[3];
  double p[3];
  double depth;  // NaN where the reference's constraint has no penetrationDepth override
};

// Constructor arguments of a teleported rigidbody3d contact (RigidBody3DSim.cpp:1338-1397): TeleportedSphereSphereConstraint{ i, j, x0, x1, ri, rj }
// or KinematicObjectSphereConstraint{ free sphere, r, n, kinematic sphere, its teleported centre, 0, 0 }
struct GpuTeleportedContact3D
{
  uint32_t portal0, portal1;
  double x0[3], x1[3];
};

class GpuRigidBody3DBackend final
{
public:
  explicit GpuRigidBody3DBackend( const int device = 0 );
  ~GpuRigidBody3DBackend();
  GpuRigidBody3DBackend( const GpuRigidBody3DBackend& ) = delete;
  GpuRigidBody3DBackend& operator=( const GpuRigidBody3DBackend& ) = delete;

  // RigidBody3DState::geometry(): per entry its type (SG_GEO_BOX / SG_GEO_SPHERE / SG_GEO_MESH), sphere radius, box
  // half-widths (3 per entry) and, for meshes, the index returned by addMesh
  void setGeometry( const std::vector<uint32_t>& type, const std::vector<double>& r, const std::vector<double>& half_widths, const std::vector<uint32_t>& mesh );
  // RigidBodyTriangleMesh members: vertices, surface samples, convex-hull vertices (3 x n, column major as stored by the
  // reference), signed-distance grid (cell_delta, dimensions, origin, values with x fastest)
  uint32_t addMesh( const std::vector<double>& verts, const std::vector<double>& samples, const std::vector<double>& hull, const double cell_delta[3], const uint32_t dims[3],
                    const double origin[3], const std::vector<double>& sdf );
  // the mesh's own snapshot record: what RigidBodyTriangleMesh::serialize( stm ) writes for it (RigidBodyTriangleMesh.cpp:215-232); serializeState writes it back
  void setMeshSnapshot( const uint32_t mesh_index, const std::string& record );
  // per body: geometry index, isKinematicallyScripted, total mass and body-frame inertia (the diagonals of M0)
  void setBodies( const std::vector<uint32_t>& geo_of_body, const std::vector<uint8_t>& fixed, const VectorXs& m, const VectorXs& I0 );
  void setGravity( const double gx, const double gy, const double gz );
  void setPlanes( const std::vector<double>& x, const std::vector<double>& n );
  // RigidBody3DState::staticCylinders(): point on the axis, axis, radius per cylinder
  void setCylinders( const std::vector<double>& x, const std::vector<double>& axis, const std::vector<double>& r );

  // RigidBody3DState::planarPortals() (rigidbody3d/Portals/PlanarPortal.h): plane A and B as (x, n), the integer portal multipliers (3 per portal).
  // All-sphere scenes only (the reference's teleported collisions are sphere-only); computeActiveSet then lists the contacts of un-teleported
  // pairs | SG_SPHERE_SPHERE_TELEPORTED / SG_KINEMATIC_OBJECT_SPHERE_TELEPORTED | planes | cylinders (RigidBody3DSim.cpp:1072-1397)
  void setPortals( const std::vector<double>& plane_a_x, const std::vector<double>& plane_a_n, const std::vector<double>& plane_b_x, const std::vector<double>& plane_b_n,
                   const std::vector<int32_t>& multiplier );
  // RigidBody3DSim::enforcePeriodicBoundaryConditions (RigidBody3DSim.cpp:642-663): centres of mass in q, in place
  void enforcePeriodicBoundaryConditions( VectorXs& q );
  // after a computeActiveSet with portals: teleported centres ( x0, x1 at q0 ) of contacts [num_regular, num_regular + teleported.size())
  void teleportedContacts( std::vector<GpuTeleportedContact3D>& teleported, uint64_t* num_regular = nullptr );

  void flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 );
  // RigidBody3DState::updateMandMinv (rigidbody3d/RigidBody3DState.cpp:428-462) on the GPU: writes the 9 N values of the inertia
  // blocks of M and of Minv in place -- pass &M.data().value( 3 * nbodies ) and &Minv.data().value( 3 * nbodies ).  from_last_flow:
  // q is the q1 the last flow() wrote (still on the device).  Flows after this call read the updated M (SG_MAP_M_UPDATED).
  void updateMandMinv( const VectorXs& q, double* m_values, double* minv_values, const bool from_last_flow = false );
  // RigidBody3DState::serialize / deserialize (rigidbody3d/RigidBody3DState.cpp:586-668) from / into the device-resident state (spheres and boxes; the mass
  // matrices in the layout this backend's flows read: the constructor's until updateMandMinv has run, its own afterwards)
  void serializeState( std::ostream& output_stream, const bool from_last_flow = true ) const;
  void deserializeState( std::istream& input_stream, const bool from_running_simulation = true );
  // false + message on std::cerr where the reference would print and exit (unsupported geometry pairing): the caller exits
  void computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact3D>& contacts, uint64_t* num_candidates = nullptr, const bool from_last_flow = false );
  // rigidbody3d/SpatialGridDetector.h:37 on caller-built boxes [minx,miny,minz,maxx,maxy,maxz]
  void getPotentialOverlaps( const std::vector<double>& aabbs, std::vector<std::pair<unsigned,unsigned>>& overlaps );

  sg_ctx* context() { return m_ctx; }
  GravityOnlyGuard& forceGuard() { return m_guard; }

private:
  void check( const int rc, const char* what ) const;
  sg_ctx* m_ctx;
  unsigned m_nbodies;
  bool m_m_updated = false;
  GravityOnlyGuard m_guard;
};

class GpuSplitHamMap final : public UnconstrainedMap
{
public:
  explicit GpuSplitHamMap( GpuRigidBody3DBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "split_ham"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuRigidBody3DBackend& m_backend;
};

class GpuDMVMap final : public UnconstrainedMap
{
public:
  explicit GpuDMVMap( GpuRigidBody3DBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "dmv"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuRigidBody3DBackend& m_backend;
};

// ---- rigidbody2d ---------------------------------------------------------------------------------------------------
class GpuRigidBody2DBackend final
{
public:
  explicit GpuRigidBody2DBackend( const int device = 0 );
  ~GpuRigidBody2DBackend();
  GpuRigidBody2DBackend( const GpuRigidBody2DBackend& ) = delete;
  GpuRigidBody2DBackend& operator=( const GpuRigidBody2DBackend& ) = delete;

  // RigidBody2DState::geometry(): type (SG_GEO2D_CIRCLE / SG_GEO2D_BOX), circle radius, box half-widths (2 per entry)
  void setGeometry( const std::vector<uint32_t>& type, const std::vector<double>& r, const std::vector<double>& half_widths );
  // geometry index, fixed flag and the 3N diagonal of M() = [m, m, I] per body
  void setBodies( const std::vector<uint32_t>& geo_of_body, const std::vector<uint8_t>& fixed, const VectorXs& M );
  void setGravity( const double gx, const double gy );
  void setPlanes( const std::vector<double>& x, const std::vector<double>& n );

  // RigidBody2DState::planarPortals() (rigidbody2d/PlanarPortal.h): plane A and B as (x, n; n used as given), velocity, bounds.
  // With portals set computeActiveSet follows computeBodyBodyActiveSetSpatialGridWithPortals (rigidbody2d/RigidBody2DSim.cpp:876-1040).
  void setPortals( const std::vector<double>& plane_a_x, const std::vector<double>& plane_a_n, const std::vector<double>& plane_b_x, const std::vector<double>& plane_b_n,
                   const std::vector<double>& velocity, const std::vector<double>& bounds );
  // RigidBody2DSim::updatePeriodicBoundaryConditionsStartOfStep / enforcePeriodicBoundaryConditions (RigidBody2DSim.cpp:832-874)
  void updatePeriodicBoundaryConditionsStartOfStep( const unsigned next_iteration, const scalar& dt );
  void enforcePeriodicBoundaryConditions( VectorXs& q, VectorXs& v );
  // after a computeActiveSet with portals: constructor arguments of the SG_CIRCLE_CIRCLE_TELEPORTED / _KICK_TELEPORTED contacts,
  // which are entries [num_regular, num_regular + teleported.size()) of the contact list
  void teleportedContacts( std::vector<GpuTeleportedContact2D>& teleported, uint64_t* num_regular = nullptr );

  void flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 );
  // contacts use GpuContact2D with the rigidbody2d type codes (SG_CIRCLE_CIRCLE ... SG_PLANE_BODY_2D)
  void computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact2D>& contacts, uint64_t* num_candidates = nullptr, const bool from_last_flow = false );
  // RigidBody2DState::serialize / deserialize (rigidbody2d/RigidBody2DState.cpp:485-556) from / into the device-resident state
  void serializeState( std::ostream& output_stream, const bool from_last_flow = true ) const;
  void deserializeState( std::istream& input_stream );

  sg_ctx* context() { return m_ctx; }
  GravityOnlyGuard& forceGuard() { return m_guard; }

private:
  void check( const int rc, const char* what ) const;
  sg_ctx* m_ctx;
  unsigned m_nbodies;
  GravityOnlyGuard m_guard;
};

class GpuRB2DSymplecticEulerMap final : public UnconstrainedMap
{
public:
  explicit GpuRB2DSymplecticEulerMap( GpuRigidBody2DBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "symplectic_euler"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuRigidBody2DBackend& m_backend;
};

class GpuRB2DVerletMap final : public UnconstrainedMap
{
public:
  explicit GpuRB2DVerletMap( GpuRigidBody2DBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "verlet"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuRigidBody2DBackend& m_backend;
};

#endif
