// oracle/ref_shims/ref_rb2d_sim.cpp -- TEST INFRASTRUCTURE.
// Drives the reference's OWN RigidBody2DSim (rigidbody2d/RigidBody2DSim.cpp + RigidBody2DState.cpp + every geometry, constraint and utility file they use,
// compiled unchanged by oracle/Makefile.ref against the Eigen stand-in): RigidBody2DSim::computeActiveSet as a whole -- swept / rotated AABBs, the spatial
// grid, dispatchNarrowPhaseCollision (circle-circle CCD, circle-box, box-box, kinematic rules), the portal branch with its teleported collisions, planes --
// in the order the reference emits the constraints, and RigidBody2DSim::flow with an unconstrained map.  This pins the GLUE of oracle/rb2d.h and
// oracle/rb2d_portals.h.  Stubbed: ImpactMap::flow (the LCP solver stack; only a flow overload that is never called here refers to it).
#include "rigidbody2d/RigidBody2DSim.h"
#include "rigidbody2d/RigidBody2DState.h"
#include "rigidbody2d/CircleGeometry.h"
#include "rigidbody2d/BoxGeometry.h"
#include "rigidbody2d/NearEarthGravityForce.h"
#include "rigidbody2d/RigidBody2DStaticPlane.h"
#include "rigidbody2d/PlanarPortal.h"
#include "rigidbody2d/PythonScripting.h"
#include "rigidbody2d/SymplecticEulerMap.h"
#include "rigidbody2d/VerletMap.h"
#include "scisim/Constraints/Constraint.h"
#include "scisim/ConstrainedMaps/ImpactMaps/ImpactMap.h"
#include "scisim/Math/Rational.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>

// stub: see the header comment
void ImpactMap::flow( ScriptingCallback&, FlowableSystem&, ConstrainedSystem&, UnconstrainedMap&, ImpactOperator&, const unsigned, const scalar&, const scalar&, const VectorXs&, const VectorXs&, VectorXs&, VectorXs& )
{
  std::cerr << "oracle/ref_shims: ImpactMap::flow is not part of the compiled reference subset" << std::endl;
  std::abort();
}

extern "C"
{

// q, v: 3 n ( x, y, theta per body ); M: 3 n ( m, m, I per body ).  Geometry table: geo_type 0 circle ( geo_r ), 1 box ( geo_half ).  planes: x[2], n[2];
// portals: plane A, plane B, velocity, bounds -- the arguments of the reference's constructors.
void* ref_rb2d_sim_create( const uint32_t n, const double* q, const double* v, const double* M, const uint8_t* fixed, const uint32_t* geo_of_body,
                           const uint32_t ngeo, const uint32_t* geo_type, const double* geo_r, const double* geo_half, const double* g,
                           const uint32_t nplanes, const double* px, const double* pn,
                           const uint32_t nportals, const double* pax, const double* pan, const double* pbx, const double* pbn, const double* pv, const double* pbounds )
{
  VectorXs wq{ int( 3 * n ) }, wv{ int( 3 * n ) }, wm{ int( 3 * n ) };
  VectorXu gi;
  gi.resize( int( n ) );
  std::vector<bool> fx( n );
  for( uint32_t k = 0; k < 3 * n; ++k ) { wq( int( k ) ) = q[k]; wv( int( k ) ) = v[k]; wm( int( k ) ) = M[k]; }
  for( uint32_t b = 0; b < n; ++b ) { gi( int( b ) ) = geo_of_body[b]; fx[b] = fixed[b] != 0; }
  std::vector<std::unique_ptr<RigidBody2DGeometry>> geometry;
  for( uint32_t k = 0; k < ngeo; ++k )
  {
    if( geo_type[k] == 0u ) { geometry.emplace_back( new CircleGeometry{ geo_r[k] } ); }
    else { geometry.emplace_back( new BoxGeometry{ Vector2s{ geo_half[2 * k], geo_half[2 * k + 1] } } ); }
  }
  std::vector<std::unique_ptr<RigidBody2DForce>> forces;
  forces.emplace_back( new NearEarthGravityForce{ Vector2s{ g[0], g[1] } } );
  std::vector<RigidBody2DStaticPlane> planes;
  for( uint32_t k = 0; k < nplanes; ++k ) { planes.emplace_back( Vector2s{ px[2 * k], px[2 * k + 1] }, Vector2s{ pn[2 * k], pn[2 * k + 1] } ); }
  std::vector<PlanarPortal> portals;
  for( uint32_t k = 0; k < nportals; ++k )
  {
    const RigidBody2DStaticPlane a{ Vector2s{ pax[2 * k], pax[2 * k + 1] }, Vector2s{ pan[2 * k], pan[2 * k + 1] } };
    const RigidBody2DStaticPlane b{ Vector2s{ pbx[2 * k], pbx[2 * k + 1] }, Vector2s{ pbn[2 * k], pbn[2 * k + 1] } };
    portals.emplace_back( a, b, pv[k], pbounds[k] );
  }
  RigidBody2DSim* sim = new RigidBody2DSim;
  sim->state() = RigidBody2DState{ wq, wv, wm, fx, gi, geometry, forces, planes, portals };
  return sim;
}

void ref_rb2d_sim_destroy( void* h ) { delete static_cast<RigidBody2DSim*>( h ); }

// RigidBody2DSim::computeActiveSet( q0, q1, v ) (rigidbody2d/RigidBody2DSim.cpp:696-714).  Per constraint, in the reference's order: the contact type code of
// include/scisim_b200.h (from name(); 99 = a name without a code), the body indices ( second = 0xffffffff where there is none ), the static object index
// ( 0xffffffff where there is none ), the world-space normal and contact point at q0 and penetrationDepth( q1 ).
uint64_t ref_rb2d_sim_active_set( void* h, const double* q0, const double* q1, const uint64_t cap, uint32_t* type, uint32_t* ci, uint32_t* cj, uint32_t* cstatic,
                                  double* cn, double* cp, double* depth )
{
  RigidBody2DSim& sim = *static_cast<RigidBody2DSim*>( h );
  const int nq = int( sim.state().q().size() );
  VectorXs wq0{ nq }, wq1{ nq }, wv{ nq };
  for( int k = 0; k < nq; ++k ) { wq0( k ) = q0[k]; wq1( k ) = q1[k]; }
  wv.setZero();
  const VectorXs& vq0 = wq0; const VectorXs& vq1 = wq1;
  std::vector<std::unique_ptr<Constraint>> active_set;
  sim.computeActiveSet( vq0, vq1, wv, active_set );
  static const struct { const char* name; uint32_t code; bool is_static; bool has_normal; } table[] = {
    { "circle_circle", 20u, false, true }, { "kinematic_object_circle", 21u, false, true }, { "body_body", 22u, false, true }, { "static_plane_circle", 23u, true, true },
    { "static_plane_body", 24u, true, true }, { "teleported_circle_circle", 25u, false, false }, { "kinematic_kick_circle_circle", 26u, false, false } };
  uint64_t k = 0;
  for( const std::unique_ptr<Constraint>& con : active_set )
  {
    if( k < cap )
    {
      const std::string name{ con->name() };
      uint32_t code = 99u; bool is_static = false, has_normal = true;
      for( const auto& e : table ) { if( name == e.name ) { code = e.code; is_static = e.is_static; has_normal = e.has_normal; } }
      type[k] = code;
      std::pair<int,int> bodies{ -1, -1 };
      con->getBodyIndices( bodies );
      ci[k] = uint32_t( bodies.first ); cj[k] = uint32_t( bodies.second );
      cstatic[k] = is_static ? con->getStaticObjectIndex() : 0xffffffffu;
      VectorXs n, p;
      if( has_normal ) { con->getWorldSpaceContactNormal( vq0, n ); cn[2 * k] = n( 0 ); cn[2 * k + 1] = n( 1 ); }
      else
      {
        // the teleported classes implement no getWorldSpaceContactNormal (the base class exits): the normal is column 0 of the contact basis
        MatrixXXsc basis;
        con->computeBasis( vq0, wv, basis );
        cn[2 * k] = basis( 0, 0 ); cn[2 * k + 1] = basis( 1, 0 );
      }
      con->getWorldSpaceContactPoint( vq0, p );
      cp[2 * k] = p( 0 ); cp[2 * k + 1] = p( 1 );
      depth[k] = con->penetrationDepth( vq1 );
    }
    ++k;
  }
  return k;
}

// RigidBody2DSim::flow( call_back, iteration, dt, umap ): portals advanced to the step's time, the unconstrained map, the periodic boundary conditions;
// the simulation's state is advanced and returned.  kind 0: SymplecticEulerMap, 1: VerletMap; dt = dt_num / dt_den.
void ref_rb2d_sim_flow( void* h, const int kind, const unsigned iteration, const long long dt_num, const long long dt_den, double* q_out, double* v_out )
{
  RigidBody2DSim& sim = *static_cast<RigidBody2DSim*>( h );
  PythonScripting call_back;
  const Rational<std::intmax_t> dt{ std::intmax_t( dt_num ), std::intmax_t( dt_den ) };
  if( kind == 0 ) { SymplecticEulerMap umap; sim.flow( call_back, iteration, dt, umap ); }
  else { VerletMap umap; sim.flow( call_back, iteration, dt, umap ); }
  const int nq = int( sim.state().q().size() );
  for( int k = 0; k < nq; ++k ) { q_out[k] = sim.state().q()( k ); v_out[k] = sim.state().v()( k ); }
}

void ref_rb2d_sim_set_state( void* h, const double* q, const double* v )
{
  RigidBody2DSim& sim = *static_cast<RigidBody2DSim*>( h );
  const int nq = int( sim.state().q().size() );
  for( int k = 0; k < nq; ++k ) { sim.state().q()( k ) = q[k]; sim.state().v()( k ) = v[k]; }
}

// RigidBody2DState::serialize (rigidbody2d/RigidBody2DState.cpp:485-498) of the simulation's current state: the reference's binary snapshot.  Returns its
// length; the bytes are written when they fit cap.
uint64_t ref_rb2d_sim_serialize_state( void* h, void* buf, const uint64_t cap )
{
  RigidBody2DSim& sim = *static_cast<RigidBody2DSim*>( h );
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  sim.state().serialize( stm );
  const std::string bytes = stm.str();
  if( bytes.size() <= cap ) { std::memcpy( buf, bytes.data(), bytes.size() ); }
  return bytes.size();
}

// RigidBody2DState::deserialize (rigidbody2d/RigidBody2DState.cpp:542-556) of a snapshot (the product's sg_rb2d_state_serialize output, say) into a fresh
// RigidBody2DSim; *n_out = its body count
void* ref_rb2d_sim_from_snapshot( const void* buf, const uint64_t bytes, uint32_t* n_out )
{
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  stm.write( static_cast<const char*>( buf ), std::streamsize( bytes ) );
  RigidBody2DSim* sim = new RigidBody2DSim;
  sim->state().deserialize( stm );
  *n_out = sim->state().nbodies();
  return sim;
}

// PlanarPortal::updateMovingPortals( t ) on every portal of the state (what RigidBody2DSim::flow does before the map): a snapshot holds each portal's m_dx
void ref_rb2d_sim_update_portals( void* h, const double t )
{
  RigidBody2DSim& sim = *static_cast<RigidBody2DSim*>( h );
  for( PlanarPortal& p : sim.state().planarPortals() ) { p.updateMovingPortals( t ); }
}

void ref_rb2d_sim_get_state( void* h, double* q, double* v )
{
  RigidBody2DSim& sim = *static_cast<RigidBody2DSim*>( h );
  const int nq = int( sim.state().q().size() );
  for( int k = 0; k < nq; ++k ) { q[k] = sim.state().q()( k ); v[k] = sim.state().v()( k ); }
}

}
