// oracle/ref_shims/ref_rb2d.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own, unmodified
//   rigidbody2d/SpatialGrid.cpp     (2-D AABB grid of the rigid-body sim)
//   rigidbody2d/BoxBoxTools.cpp     (BoxBoxTools::isActive)
//   rigidbody2d/CircleBoxTools.cpp  (CircleBoxTools::isActive)
//   rigidbody2d/CircleGeometry.cpp, BoxGeometry.cpp (+ RigidBody2DGeometry.cpp)   (computeCollisionAABB, computeAABB)
//   rigidbody2d/SymplecticEulerMap.cpp, VerletMap.cpp, NearEarthGravityForce.cpp (+ RigidBody2DForce.cpp, scisim/UnconstrainedMaps/*.cpp)   (the maps)
// compiled from /root/reference against oracle/eigen_standin (oracle/Makefile.ref).
#include "rigidbody2d/SpatialGrid.h"
#include "rigidbody2d/BoxBoxTools.h"
#include "rigidbody2d/CircleBoxTools.h"

#include "rigidbody2d/PlanarPortal.h"
#include "rigidbody2d/CircleGeometry.h"
#include "rigidbody2d/BoxGeometry.h"
#include "rigidbody2d/SymplecticEulerMap.h"
#include "rigidbody2d/VerletMap.h"
#include "rigidbody2d/NearEarthGravityForce.h"
#include "scisim/UnconstrainedMaps/FlowableSystem.h"
#include "rigidbody2d/CircleCircleConstraint.h"
#include "rigidbody2d/StaticPlaneCircleConstraint.h"
#include "rigidbody2d/StaticPlaneBodyConstraint.h"
#include "rigidbody2d/BodyBodyConstraint.h"
#include "rigidbody2d/ConstraintCache.h"
#include "rigidbody2d/KinematicObjectCircleConstraint.h"
#include "rigidbody2d/RigidBody2DStaticPlane.h"
#include <memory>
#include <cstring>
#include <sstream>

#include <cstdint>

extern "C"
{

uint64_t ref_rb2d_overlaps( const uint32_t n, const double* boxes, uint32_t* ij, const uint64_t cap )
{
  std::vector<AABB> aabbs;
  aabbs.reserve( n );
  for( uint32_t i = 0; i < n; ++i ) { aabbs.emplace_back( Array2s{ boxes[4 * i], boxes[4 * i + 1] }, Array2s{ boxes[4 * i + 2], boxes[4 * i + 3] } ); }
  std::set<std::pair<unsigned,unsigned>> overlaps;
  SpatialGrid::getPotentialOverlaps( aabbs, overlaps );
  uint64_t k = 0;
  for( const std::pair<unsigned,unsigned>& p : overlaps ) { if( k < cap ) { ij[2 * k] = p.first; ij[2 * k + 1] = p.second; } ++k; }
  return k;
}

// returns the number of contact points (<= 2 written, 2 doubles each)
int ref_rb2d_box_box( const double* x0, const double theta0, const double* r0, const double* x1, const double theta1, const double* r1, double* n, double* points )
{
  Vector2s nn;
  nn.setZero();
  std::vector<Vector2s> pts;
  BoxBoxTools::isActive( Vector2s{ x0[0], x0[1] }, theta0, Vector2s{ r0[0], r0[1] }, Vector2s{ x1[0], x1[1] }, theta1, Vector2s{ r1[0], r1[1] }, nn, pts );
  n[0] = nn.x(); n[1] = nn.y();
  int k = 0;
  for( const Vector2s& p : pts ) { if( k < 2 ) { points[2 * k] = p.x(); points[2 * k + 1] = p.y(); } ++k; }
  return k;
}

int ref_rb2d_circle_box( const double* x0, const double r0, const double* x1, const double theta1, const double* r1, double* n, double* p )
{
  Vector2s nn, pp;
  nn.setZero(); pp.setZero();
  const bool hit = CircleBoxTools::isActive( Vector2s{ x0[0], x0[1] }, r0, Vector2s{ x1[0], x1[1] }, theta1, Vector2s{ r1[0], r1[1] }, nn, pp );
  n[0] = nn.x(); n[1] = nn.y(); p[0] = pp.x(); p[1] = pp.y();
  return hit ? 1 : 0;
}

// ---- PlanarPortal (rigidbody2d/PlanarPortal.cpp, RigidBody2DStaticPlane.cpp) ------------------------------------------
void* ref_rb2d_portal_create( const double* ax, const double* an, const double* bx, const double* bn, const double v, const double bounds )
{
  const RigidBody2DStaticPlane a{ Vector2s{ ax[0], ax[1] }, Vector2s{ an[0], an[1] } };
  const RigidBody2DStaticPlane b{ Vector2s{ bx[0], bx[1] }, Vector2s{ bn[0], bn[1] } };
  return new PlanarPortal{ a, b, v, bounds };
}
void ref_rb2d_portal_destroy( void* p ) { delete static_cast<PlanarPortal*>( p ); }
void ref_rb2d_portal_update( void* p, const double t ) { static_cast<PlanarPortal*>( p )->updateMovingPortals( t ); }
// same layout as orc_rb2d_portal_probe (oracle/capi.cpp)
uint32_t ref_rb2d_portal_probe( const void* pv, const double* box, const double* x, double* out )
{
  const PlanarPortal& p = *static_cast<const PlanarPortal*>( pv );
  const Array2s mn{ box[0], box[1] }, mx{ box[2], box[3] };
  bool plane_idx = false;
  const bool touches = p.aabbTouchesPortal( mn, mx, plane_idx );
  const uint32_t touch = touches ? ( plane_idx ? 2u : 1u ) : 0u;
  Vector2s xo;
  p.teleportPoint( Vector2s{ x[0], x[1] }, touch == 2u, xo );
  const Vector2s k{ p.getKinematicVelocityOfAABB( mn, mx ) };
  out[0] = xo.x(); out[1] = xo.y(); out[2] = k.x(); out[3] = k.y();
  return touch;
}

// type 0 circle ( r ), 1 box ( half ); swept != 0: computeCollisionAABB( x0, theta0, x1, theta1 ), else computeAABB( x1, theta1 ); out = min(2), max(2)
void ref_rb2d_aabb( const int type, const double r, const double* half, const double* q0b, const double* q1b, const int swept, double* out )
{
  Array2s mn, mx;
  const Vector2s x0{ q0b[0], q0b[1] }, x1{ q1b[0], q1b[1] };
  if( type == 0 )
  {
    const CircleGeometry g{ r };
    if( swept ) { g.computeCollisionAABB( x0, q0b[2], x1, q1b[2], mn, mx ); } else { g.computeAABB( x1, q1b[2], mn, mx ); }
  }
  else
  {
    const BoxGeometry g{ Vector2s{ half[0], half[1] } };
    if( swept ) { g.computeCollisionAABB( x0, q0b[2], x1, q1b[2], mn, mx ); } else { g.computeAABB( x1, q1b[2], mn, mx ); }
  }
  out[0] = mn( 0 ); out[1] = mn( 1 ); out[2] = mx( 0 ); out[3] = mx( 1 );
}

}

// ---- the unconstrained maps: a FlowableSystem with what they ask of RigidBody2DSim -- the diagonal M ( m, m, I per body ), Minv = 1 / M
// ( rigidbody2d/RigidBody2DState.cpp:31-43 ), the kinematic flags and computeForce = setZero + the forces ( RigidBody2DSim.cpp:104-113 ) -------
namespace
{
class ShimRB2DSystem final : public FlowableSystem
{
public:
  ShimRB2DSystem( const uint32_t n, const double* M, const uint8_t* fixed, const double* g )
  : m_fixed( fixed, fixed + n )
  , m_force( Vector2s{ g[0], g[1] } )
  {
    std::vector<double> mi( 3 * size_t( n ) );
    for( size_t k = 0; k < mi.size(); ++k ) { mi[k] = 1.0 / M[k]; }
    m_M.setDiagonal( M, int( 3 * n ) ); m_Minv.setDiagonal( mi.data(), int( 3 * n ) );
  }
  virtual int nqdofs() const override { return m_M.rows(); }
  virtual int nvdofs() const override { return m_M.rows(); }
  virtual unsigned numVelDoFsPerBody() const override { return 3; }
  virtual unsigned ambientSpaceDimensions() const override { return 2; }
  virtual bool isKinematicallyScripted( const int i ) const override { return m_fixed[size_t( i )] != 0; }
  virtual void computeForce( const VectorXs& q, const VectorXs& v, const scalar& t, VectorXs& F ) override
  {
    F.setZero();
    m_force.computeForce( q, v, m_M, F );
  }
  virtual void zeroOutForcesOnFixedBodies( VectorXs& ) const override { std::abort(); }
  virtual void linearInertialConfigurationUpdate( const VectorXs&, const VectorXs&, const scalar&, VectorXs& ) const override { std::abort(); }
  virtual const SparseMatrixsc& M() const override { return m_M; }
  virtual const SparseMatrixsc& Minv() const override { return m_Minv; }
  virtual const SparseMatrixsc& M0() const override { return m_M; }
  virtual const SparseMatrixsc& Minv0() const override { return m_Minv; }
  virtual void computeMomentum( const VectorXs&, VectorXs& ) const override { std::abort(); }
  virtual void computeAngularMomentum( const VectorXs&, VectorXs& ) const override { std::abort(); }
  virtual std::string name() const override { return "shim_rigid_body_2d"; }
private:
  SparseMatrixsc m_M, m_Minv;
  std::vector<uint8_t> m_fixed;
  NearEarthGravityForce m_force;
};
}

extern "C"
{

// kind 0: SymplecticEulerMap::flow, 1: VerletMap::flow; q, v: 3n doubles ( x, y, theta per body ); M: the 3n mass diagonal
void ref_rb2d_flow( const int kind, const uint32_t n, const double* M, const uint8_t* fixed, const double* g, const double* q0, const double* v0, const unsigned iteration, const double dt,
                    double* q1, double* v1 )
{
  ShimRB2DSystem sys{ n, M, fixed, g };
  VectorXs vq0( int( 3 * n ) ), vv0( int( 3 * n ) ), vq1( int( 3 * n ) ), vv1( int( 3 * n ) );
  for( uint32_t k = 0; k < 3 * n; ++k ) { vq0( int( k ) ) = q0[k]; vv0( int( k ) ) = v0[k]; }
  if( kind == 0 ) { SymplecticEulerMap map; map.flow( vq0, vv0, sys, iteration, dt, vq1, vv1 ); }
  else { VerletMap map; map.flow( vq0, vv0, sys, iteration, dt, vq1, vv1 ); }
  for( uint32_t k = 0; k < 3 * n; ++k ) { q1[k] = vq1( int( k ) ); v1[k] = vv1( int( k ) ); }
}

}


// ---- the constraint classes themselves (rigidbody2d/{CircleCircle,StaticPlaneCircle,StaticPlaneBody,BodyBody}Constraint.cpp + scisim/Constraints/
// Constraint.cpp, compiled unchanged): the constraint built as RigidBody2DSim builds it, its normal, world-space contact point at q0 and -- where the
// class overrides it -- penetrationDepth at q1 -----------------------------------------------------------------------------------------------------
extern "C"
{

// q0, q1: 3 n doubles ( x, y, theta per body ).
// kind 0  circle-circle ( i, j ), geo = r_i, r_j: n, p formed as RigidBody2DSim.cpp:299-303 forms them (two lines restated, marked)
// kind 1  plane-circle: geo = x[2], n[2], r;  out[0] = StaticPlaneCircleConstraint::isActive at q1
// kind 2  plane-body (box corner): geo = x[2], n[2], arm[2] (body-space arm of the corner)
// kind 3  body-body ( i, j ): geo = p[2], n[2] as the narrow phase produced them
// out[0] isActive (kinds 0, 1; 1 otherwise), out[1..2] normal, out[3..4] contact point at q0, out[5] penetrationDepth( q1 ) (NaN where the class has no override)
void ref_rb2d_constraint_probe( const int kind, const unsigned i, const unsigned j, const uint32_t nbodies, const double* q0, const double* q1, const double* geo, double* out )
{
  const int nq = 3 * int( nbodies );
  VectorXs wq0{ nq }, wq1{ nq };
  for( int k = 0; k < nq; ++k ) { wq0( k ) = q0[k]; wq1( k ) = q1[k]; }
  const VectorXs& vq0 = wq0; const VectorXs& vq1 = wq1;
  const RigidBody2DStaticPlane plane{ Vector2s{ geo[0], geo[1] }, ( kind == 1 || kind == 2 ) ? Vector2s{ geo[2], geo[3] } : Vector2s{ 0.0, 1.0 } }; // outlives the constraints
  std::unique_ptr<Constraint> con;
  bool active = true;
  if( kind == 0 )
  {
    const scalar ra = geo[0], rb = geo[1];
    const Vector2s q0a{ vq0.segment<2>( 3 * i ) }, q0b{ vq0.segment<2>( 3 * j ) };
    active = CircleCircleConstraint::isActive( vq1.segment<2>( 3 * i ), vq1.segment<2>( 3 * j ), ra, rb );
    // restated glue (RigidBody2DSim.cpp:299, 303):
    const Vector2s n{ ( q0a - q0b ).normalized() };
    const Vector2s p{ q0a + ( ra / ( ra + rb ) ) * ( q0b - q0a ) };
    con.reset( new CircleCircleConstraint{ i, j, n, p, ra, rb } );
  }
  else if( kind == 1 )
  {
    active = StaticPlaneCircleConstraint::isActive( vq1.segment<2>( 3 * i ), geo[4], plane );
    con.reset( new StaticPlaneCircleConstraint{ i, j, geo[4], plane } );
  }
  else if( kind == 2 )
  {
    con.reset( new StaticPlaneBodyConstraint{ i, Vector2s{ geo[4], geo[5] }, plane, j } );
  }
  else
  {
    con.reset( new BodyBodyConstraint{ i, j, Vector2s{ geo[0], geo[1] }, Vector2s{ geo[2], geo[3] }, vq0 } );
  }
  out[0] = active ? 1.0 : 0.0;
  VectorXs n, p;
  con->getWorldSpaceContactNormal( vq0, n );
  con->getWorldSpaceContactPoint( vq0, p );
  out[1] = n( 0 ); out[2] = n( 1 ); out[3] = p( 0 ); out[4] = p( 1 );
  out[5] = con->penetrationDepth( vq1 );
}

}

// ---- rigidbody2d/ConstraintCache.cpp compiled unchanged: cacheConstraint for a list of constraints, then getCachedConstraint for another list.
// Constraints are built with the reference's own classes (only their indices and names matter to the cache).  type = the contact type codes of
// include/scisim_b200.h; a = first body, b = second body / static object.  r: ncomp doubles per constraint.  Returns constraintCacheEmpty() after the stores.
// (20 circle-circle, 22 body-body, 21 kinematic circle, 23 plane-circle -- anything else exits, as in the reference)
extern "C"
{
int ref_rb2d_cache_roundtrip_ex( const uint32_t nstore, const uint32_t* stype, const uint32_t* sa, const uint32_t* sb, const uint32_t ncomp, const double* rstore,
                              const uint32_t nquery, const uint32_t* qtype, const uint32_t* qa, const uint32_t* qb, double* rout,
                                 void* ser_out, const uint64_t ser_cap, uint64_t* ser_bytes, const void* deser_in, const uint64_t deser_bytes )
{
  uint32_t nb = 1;
  for( uint32_t k = 0; k < nstore; ++k ) { nb = std::max( nb, std::max( sa[k], stype[k] != 23 ? sb[k] : 0u ) + 1u ); }
  for( uint32_t k = 0; k < nquery; ++k ) { nb = std::max( nb, std::max( qa[k], qtype[k] != 23 ? qb[k] : 0u ) + 1u ); }
  VectorXs wq{ int( 3 * nb ) };
  for( uint32_t b = 0; b < nb; ++b ) { wq( int( 3 * b ) ) = double( b ); wq( int( 3 * b + 1 ) ) = 0.5 * double( b ); wq( int( 3 * b + 2 ) ) = 0.0; }
  const VectorXs& q = wq;
  const RigidBody2DStaticPlane plane{ Vector2s{ 0.0, 0.0 }, Vector2s{ 0.0, 1.0 } };
  const Vector2s n{ 1.0, 0.0 }, p{ 0.0, 0.0 };
  const auto make = [&]( const uint32_t type, const uint32_t a, const uint32_t b ) -> std::unique_ptr<Constraint>
  {
    if( type == 20 ) { return std::unique_ptr<Constraint>{ new CircleCircleConstraint{ a, b, n, p, 0.5, 0.5 } }; }
    if( type == 22 ) { return std::unique_ptr<Constraint>{ new BodyBodyConstraint{ a, b, p, n, q } }; }
    if( type == 23 ) { return std::unique_ptr<Constraint>{ new StaticPlaneCircleConstraint{ a, b, 0.5, plane } }; }
    return std::unique_ptr<Constraint>{ new KinematicObjectCircleConstraint{ a, 0.5, n, b, p, Vector2s::Zero(), 0.0 } };
  };
  ConstraintCache cache;
  VectorXs r{ int( ncomp ) };
  for( uint32_t k = 0; k < nstore; ++k )
  {
    for( uint32_t c = 0; c < ncomp; ++c ) { r( int( c ) ) = rstore[size_t( k ) * ncomp + c]; }
    cache.cacheConstraint( *make( stype[k], sa[k], sb[k] ), r );
  }
  const int empty = cache.empty() ? 1 : 0;
  // optionally: ConstraintCache::serialize of what was stored, handed to the caller, and / or the queries answered by a second cache that
  // ConstraintCache::deserialize filled from the caller's bytes
  if( ser_bytes != nullptr )
  {
    std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
    cache.serialize( stm );
    const std::string bytes = stm.str();
    *ser_bytes = bytes.size();
    if( ser_out != nullptr && bytes.size() <= ser_cap ) { std::memcpy( ser_out, bytes.data(), bytes.size() ); }
  }
  ConstraintCache restored;
  if( deser_in != nullptr )
  {
    std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
    stm.write( static_cast<const char*>( deser_in ), std::streamsize( deser_bytes ) );
    restored.deserialize( stm );
  }
  ConstraintCache& qcache = ( deser_in != nullptr ) ? restored : cache;
  for( uint32_t k = 0; k < nquery; ++k )
  {
    for( uint32_t c = 0; c < ncomp; ++c ) { r( int( c ) ) = -7.0; }
    qcache.getCachedConstraint( *make( qtype[k], qa[k], qb[k] ), r );
    for( uint32_t c = 0; c < ncomp; ++c ) { rout[size_t( k ) * ncomp + c] = r( int( c ) ); }
  }
  return empty;
}
int ref_rb2d_cache_roundtrip( const uint32_t nstore, const uint32_t* stype, const uint32_t* sa, const uint32_t* sb, const uint32_t ncomp, const double* rstore,
                              const uint32_t nquery, const uint32_t* qtype, const uint32_t* qa, const uint32_t* qb, double* rout )
{
  return ref_rb2d_cache_roundtrip_ex( nstore, stype, sa, sb, ncomp, rstore, nquery, qtype, qa, qb, rout, nullptr, 0, nullptr, nullptr, 0 );
}
}
