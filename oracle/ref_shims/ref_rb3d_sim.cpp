// oracle/ref_shims/ref_rb3d_sim.cpp -- TEST INFRASTRUCTURE.
// Drives the reference's OWN RigidBody3DSim (rigidbody3d/RigidBody3DSim.cpp + RigidBody3DState.cpp + every geometry, constraint and utility file they use,
// compiled unchanged by oracle/Makefile.ref against the Eigen stand-in): RigidBody3DSim::computeActiveSet as a whole -- AABBs at q1, the spatial grid,
// dispatchNarrowPhaseCollision with its kinematic rules, the portal branch with its teleported collisions, planes, cylinders -- in the order the reference
// emits the constraints.  This pins the GLUE of oracle/rb3d.h and oracle/rb3d_portals.h.  Stubbed: ImpactMap's constructor and flow (the class needs the LCP
// solver stack; RigidBody3DSim holds one as a member and only a flow overload that is never called here uses it).
#include "rigidbody3d/RigidBody3DSim.h"
#include "rigidbody3d/RigidBody3DState.h"
#include "rigidbody3d/Geometry/RigidBodyBox.h"
#include "rigidbody3d/Geometry/RigidBodySphere.h"
#include "rigidbody3d/Geometry/RigidBodyTriangleMesh.h"
#include "rigidbody3d/Forces/NearEarthGravityForce.h"
#include "rigidbody3d/StaticGeometry/StaticPlane.h"
#include "rigidbody3d/StaticGeometry/StaticCylinder.h"
#include "rigidbody3d/Portals/PlanarPortal.h"
#include "scisim/Constraints/Constraint.h"
#include "scisim/ConstrainedMaps/ImpactMaps/ImpactMap.h"
#include "scisim/Math/Rational.h"
#include "rigidbody3d/PythonScripting.h"
#include "rigidbody3d/UnconstrainedMaps/SplitHamMap.h"
#include "rigidbody3d/UnconstrainedMaps/DMVMap.h"
#include "rigidbody3d/UnconstrainedMaps/ExponentialEulerMap.h"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <cstring>

// stubs: see the header comment (RigidBody3DSim holds an ImpactMap member)
ImpactMap::ImpactMap( const bool warm_start ) : m_warm_start( warm_start ) {}
void ImpactMap::flow( ScriptingCallback&, FlowableSystem&, ConstrainedSystem&, UnconstrainedMap&, ImpactOperator&, const unsigned, const scalar&, const scalar&, const VectorXs&, const VectorXs&, VectorXs&, VectorXs& )
{
  std::cerr << "oracle/ref_shims: ImpactMap::flow is not part of the compiled reference subset" << std::endl;
  std::abort();
}

extern "C"
{

// q: 12 n ( 3n positions | 9n row-major rotations ), v: 6 n, m: n, I0: 3 n.  Geometry table: geo_type 0 box ( geo_half ), 1 sphere ( geo_r ), 3 mesh
// ( geo_mesh[k] = a RigidBodyTriangleMesh* made by ref_rb3d_mesh_create; cloned ).  planes: x[3], n[3]; cylinders: x[3], axis[3], r; portals: plane A,
// plane B, integer multipliers[3] -- the arguments of the reference's constructors.
void* ref_rb3d_sim_create( const uint32_t n, const double* q, const double* v, const double* m, const double* I0, const uint8_t* fixed, const uint32_t* geo_of_body,
                           const uint32_t ngeo, const uint32_t* geo_type, const double* geo_r, const double* geo_half, void* const* geo_mesh, const double* g,
                           const uint32_t nplanes, const double* px, const double* pn, const uint32_t ncyl, const double* cx, const double* caxis, const double* cr,
                           const uint32_t nportals, const double* pax, const double* pan, const double* pbx, const double* pbn, const int* mult )
{
  std::vector<Vector3s> X( n ), V( n ), omega( n ), I0v( n );
  std::vector<scalar> M( n );
  std::vector<VectorXs> R( n );
  std::vector<bool> fx( n );
  std::vector<unsigned> geo_idx( n );
  for( uint32_t b = 0; b < n; ++b )
  {
    X[b] = Vector3s{ q[3 * b], q[3 * b + 1], q[3 * b + 2] };
    V[b] = Vector3s{ v[3 * b], v[3 * b + 1], v[3 * b + 2] };
    omega[b] = Vector3s{ v[3 * size_t( n ) + 3 * b], v[3 * size_t( n ) + 3 * b + 1], v[3 * size_t( n ) + 3 * b + 2] };
    M[b] = m[b];
    I0v[b] = Vector3s{ I0[3 * b], I0[3 * b + 1], I0[3 * b + 2] };
    R[b].resize( 9 );
    for( int k = 0; k < 9; ++k ) { R[b]( k ) = q[3 * size_t( n ) + 9 * size_t( b ) + k]; }
    fx[b] = fixed[b] != 0;
    geo_idx[b] = geo_of_body[b];
  }
  std::vector<std::unique_ptr<RigidBodyGeometry>> geometry;
  for( uint32_t k = 0; k < ngeo; ++k )
  {
    if( geo_type[k] == 0u ) { geometry.emplace_back( new RigidBodyBox{ Vector3s{ geo_half[3 * k], geo_half[3 * k + 1], geo_half[3 * k + 2] } } ); }
    else if( geo_type[k] == 1u ) { geometry.emplace_back( new RigidBodySphere{ geo_r[k] } ); }
    else { geometry.emplace_back( static_cast<const RigidBodyTriangleMesh*>( geo_mesh[k] )->clone() ); }
  }
  RigidBody3DSim* sim = new RigidBody3DSim;
  RigidBody3DState& s = sim->getState();
  s.setState( X, V, M, R, omega, I0v, fx, geo_idx, geometry );
  s.addForce( NearEarthGravityForce{ Vector3s{ g[0], g[1], g[2] } } );
  for( uint32_t k = 0; k < nplanes; ++k ) { s.addStaticPlane( StaticPlane{ Vector3s{ px[3 * k], px[3 * k + 1], px[3 * k + 2] }, Vector3s{ pn[3 * k], pn[3 * k + 1], pn[3 * k + 2] } } ); }
  for( uint32_t k = 0; k < ncyl; ++k ) { s.addStaticCylinder( StaticCylinder{ Vector3s{ cx[3 * k], cx[3 * k + 1], cx[3 * k + 2] }, Vector3s{ caxis[3 * k], caxis[3 * k + 1], caxis[3 * k + 2] }, cr[k] } ); }
  for( uint32_t k = 0; k < nportals; ++k )
  {
    const StaticPlane a{ Vector3s{ pax[3 * k], pax[3 * k + 1], pax[3 * k + 2] }, Vector3s{ pan[3 * k], pan[3 * k + 1], pan[3 * k + 2] } };
    const StaticPlane b{ Vector3s{ pbx[3 * k], pbx[3 * k + 1], pbx[3 * k + 2] }, Vector3s{ pbn[3 * k], pbn[3 * k + 1], pbn[3 * k + 2] } };
    s.addPlanarPortal( PlanarPortal{ a, b, Array3i{ mult[3 * k], mult[3 * k + 1], mult[3 * k + 2] } } );
  }
  return sim;
}

void ref_rb3d_sim_destroy( void* h ) { delete static_cast<RigidBody3DSim*>( h ); }

// RigidBody3DSim::computeActiveSet( q0, q1, v ) (rigidbody3d/RigidBody3DSim.cpp:250-262).  Per constraint, in the reference's order: the contact type code of
// include/scisim_b200.h (from name(); 99 = a name without a code), the body indices ( second = 0xffffffff where there is none ), the static object index
// ( 0xffffffff where the class has no accessor ), the world-space normal and contact point at q0 ( NaN where the class reports none ) and penetrationDepth( q1 ).
uint64_t ref_rb3d_sim_active_set( void* h, const double* q0, const double* q1, const uint64_t cap, uint32_t* type, uint32_t* ci, uint32_t* cj, uint32_t* cstatic,
                                  double* cn, double* cp, double* depth )
{
  RigidBody3DSim& sim = *static_cast<RigidBody3DSim*>( h );
  const int nb = int( sim.getState().nbodies() );
  VectorXs wq0{ 12 * nb }, wq1{ 12 * nb }, wv{ 6 * nb };
  for( int k = 0; k < 12 * nb; ++k ) { wq0( k ) = q0[k]; wq1( k ) = q1[k]; }
  wv.setZero();
  const VectorXs& vq0 = wq0; const VectorXs& vq1 = wq1;
  std::vector<std::unique_ptr<Constraint>> active_set;
  sim.computeActiveSet( vq0, vq1, wv, active_set );
  static const struct { const char* name; uint32_t code; bool is_static; } table[] = {
    { "sphere_sphere", 10u, false }, { "kinematic_sphere_sphere", 11u, false }, { "body_body", 12u, false }, { "kinematic_object_body", 13u, false },
    { "static_plane_sphere", 14u, true }, { "static_plane_box", 15u, true }, { "static_plane_body", 16u, true }, { "static_cylinder_sphere", 17u, true },
    { "static_cylinder_body", 18u, true }, { "teleported_sphere_sphere", 19u, false }, { "kinematic_object_sphere", 30u, false } };
  const double nan = std::numeric_limits<double>::quiet_NaN();
  uint64_t k = 0;
  for( const std::unique_ptr<Constraint>& con : active_set )
  {
    if( k < cap )
    {
      const std::string name{ con->name() };
      uint32_t code = 99u; bool is_static = false;
      for( const auto& e : table ) { if( name == e.name ) { code = e.code; is_static = e.is_static; } }
      type[k] = code;
      std::pair<int,int> bodies{ -1, -1 };
      if( code == 18u ) { con->getSimulatedBodyIndices( bodies ); } else { con->getBodyIndices( bodies ); }
      ci[k] = uint32_t( bodies.first ); cj[k] = uint32_t( bodies.second );
      cstatic[k] = ( is_static && code != 18u ) ? con->getStaticObjectIndex() : 0xffffffffu;
      if( code == 18u )
      {
        // no normal / point accessors (the base class exits): the normal is column 0 of the contact basis
        MatrixXXsc basis;
        con->computeBasis( vq0, wv, basis );
        for( int c = 0; c < 3; ++c ) { cn[3 * k + c] = basis( c, 0 ); cp[3 * k + c] = nan; }
      }
      else
      {
        VectorXs n, p;
        con->getWorldSpaceContactNormal( vq0, n );
        con->getWorldSpaceContactPoint( vq0, p );
        for( int c = 0; c < 3; ++c ) { cn[3 * k + c] = n( c ); cp[3 * k + c] = p( c ); }
      }
      depth[k] = con->penetrationDepth( vq1 );
    }
    ++k;
  }
  return k;
}

// RigidBody3DSim::flow( call_back, iteration, dt, umap ) (rigidbody3d/RigidBody3DSim.cpp:398-460): the unconstrained map on the simulation's own state,
// then updateMandMinv and the periodic boundaries; the state is advanced and returned.  kind 2: SplitHamMap, 3: DMVMap, 4: ExponentialEulerMap.
void ref_rb3d_sim_flow( void* h, const int kind, const unsigned iteration, const long long dt_num, const long long dt_den, double* q_out, double* v_out )
{
  RigidBody3DSim& sim = *static_cast<RigidBody3DSim*>( h );
  PythonScripting call_back;
  const Rational<std::intmax_t> dt{ std::intmax_t( dt_num ), std::intmax_t( dt_den ) };
  if( kind == 2 ) { SplitHamMap umap; sim.flow( call_back, iteration, dt, umap ); }
  else if( kind == 4 ) { ExponentialEulerMap umap; sim.flow( call_back, iteration, dt, umap ); }
  else { DMVMap umap; sim.flow( call_back, iteration, dt, umap ); }
  const int nb = int( sim.getState().nbodies() );
  for( int k = 0; k < 12 * nb; ++k ) { q_out[k] = sim.getState().q()( k ); }
  for( int k = 0; k < 6 * nb; ++k ) { v_out[k] = sim.getState().v()( k ); }
}

// RigidBody3DState::serialize (rigidbody3d/RigidBody3DState.cpp:586-612) of the simulation's current state: the reference's binary snapshot.  Returns its
// length; the bytes are written when they fit cap.  ( q, v ) may be replaced first (set != 0), and updateMandMinv run (update != 0), as RigidBody3DSim::flow does.
uint64_t ref_rb3d_sim_serialize_state( void* h, const int set, const double* q, const double* v, const int update, void* buf, const uint64_t cap )
{
  RigidBody3DSim& sim = *static_cast<RigidBody3DSim*>( h );
  const int nb = int( sim.getState().nbodies() );
  if( set != 0 )
  {
    for( int k = 0; k < 12 * nb; ++k ) { sim.getState().q()( k ) = q[k]; }
    for( int k = 0; k < 6 * nb; ++k ) { sim.getState().v()( k ) = v[k]; }
  }
  if( update != 0 ) { sim.getState().updateMandMinv(); }
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  sim.getState().serialize( stm );
  const std::string bytes = stm.str();
  if( bytes.size() <= cap ) { std::memcpy( buf, bytes.data(), bytes.size() ); }
  return bytes.size();
}

// RigidBody3DState::deserialize of a snapshot into a fresh RigidBody3DSim
void* ref_rb3d_sim_from_snapshot( const void* buf, const uint64_t bytes )
{
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  stm.write( static_cast<const char*>( buf ), std::streamsize( bytes ) );
  RigidBody3DSim* sim = new RigidBody3DSim;
  sim->getState().deserialize( stm );
  return sim;
}

// One benchmark step of the reference's own code, timed (profiles/config4_reference.py): the state is set to ( q0, v0 ) (untimed),
// RigidBody3DSim::flow( call_back, iteration, dt, umap ) is timed (the map, updateMandMinv, the boundary treatment), then
// RigidBody3DSim::computeActiveSet( q0, q1, v1 ) is timed -- one heap-allocated Constraint per contact, as the reference allocates them; freeing
// them is not timed.  Returns the seconds of the two calls; *n_active = active_set.size().  kind 2: SplitHamMap, 3: DMVMap.
double ref_rb3d_sim_step_timed( void* h, const double* q0, const double* v0, const int kind, const unsigned iteration, const long long dt_num, const long long dt_den,
                                uint64_t* n_active, double* seconds_flow )
{
  RigidBody3DSim& sim = *static_cast<RigidBody3DSim*>( h );
  const int nb = int( sim.getState().nbodies() );
  VectorXs wq0{ 12 * nb };
  for( int k = 0; k < 12 * nb; ++k ) { wq0( k ) = q0[k]; sim.getState().q()( k ) = q0[k]; }
  for( int k = 0; k < 6 * nb; ++k ) { sim.getState().v()( k ) = v0[k]; }
  PythonScripting call_back;
  const Rational<std::intmax_t> dt{ std::intmax_t( dt_num ), std::intmax_t( dt_den ) };
  SplitHamMap sh;
  DMVMap dmv;
  UnconstrainedMap& umap = ( kind == 2 ) ? static_cast<UnconstrainedMap&>( sh ) : static_cast<UnconstrainedMap&>( dmv );
  const auto t0 = std::chrono::steady_clock::now();
  sim.flow( call_back, iteration, dt, umap );
  const auto t1 = std::chrono::steady_clock::now();
  const VectorXs q1{ sim.getState().q() };
  const VectorXs v1{ sim.getState().v() };
  std::vector<std::unique_ptr<Constraint>> active_set;
  const auto t2 = std::chrono::steady_clock::now();
  sim.computeActiveSet( wq0, q1, v1, active_set );
  const auto t3 = std::chrono::steady_clock::now();
  *n_active = active_set.size();
  const double tf = std::chrono::duration<double>( t1 - t0 ).count();
  if( seconds_flow != nullptr ) { *seconds_flow = tf; }
  return tf + std::chrono::duration<double>( t3 - t2 ).count();
}

}
