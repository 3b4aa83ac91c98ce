// oracle/ref_shims/ref_ball2d.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own, unmodified
//   ball2d/SpatialGridDetector.cpp                               (AABB, getPotentialOverlaps, getPotentialOverlapsAllPairs)
//   scisim/CollisionDetection/CollisionDetectionUtilities.cpp    (computeCCDQuadraticCoeffs, ballBallCCDCollisionHappens)
//   ball2d/StaticGeometry/StaticPlane.cpp, ball2d/Portals/PlanarPortal.cpp   (portal touch tests, teleports, kinematic velocities)
//   ball2d/Constraints/{BallBall,BallStaticPlane,BallStaticDrum}Constraint.cpp, scisim/Constraints/Constraint.cpp   (isActive, constructors, normal,
//                                                                 contact point, penetrationDepth, evalgradg, computeContactBasis)
//   ball2d/SymplecticEulerMap.cpp, ball2d/VerletMap.cpp, ball2d/Forces/Ball2DGravityForce.cpp (+ Ball2DForce.cpp,
//   scisim/UnconstrainedMaps/UnconstrainedMap.cpp, FlowableSystem.cpp)       (the two unconstrained maps with the gravity force)
// compiled from /root/reference against oracle/eigen_standin (oracle/Makefile.ref).  The glue between them (swept boxes,
// pair loop) restates ball2d/Ball2DSim.cpp:553-608 and is marked as such.
#include "ball2d/SpatialGridDetector.h"
#include "scisim/CollisionDetection/CollisionDetectionUtilities.h"
#include "ball2d/Portals/PlanarPortal.h"
#include "ball2d/SymplecticEulerMap.h"
#include "ball2d/VerletMap.h"
#include "ball2d/Forces/Ball2DGravityForce.h"
#include "scisim/UnconstrainedMaps/FlowableSystem.h"
#include "ball2d/StaticGeometry/StaticDrum.h"
#include "scisim/Utilities.h"
#include "ball2d/Constraints/BallBallConstraint.h"
#include "ball2d/Constraints/BallStaticPlaneConstraint.h"
#include "ball2d/Constraints/BallStaticDrumConstraint.h"
#include "ball2d/ConstraintCache.h"

#include <memory>
#include <cstring>
#include <sstream>

#include <cstdint>

extern "C"
{

// boxes: n x [minx, miny, maxx, maxy]; writes up to cap pairs (i<j, std::set order); returns the number of pairs
uint64_t ref_ball2d_overlaps( const uint32_t n, const double* boxes, const int all_pairs, uint32_t* ij, const uint64_t cap )
{
  std::vector<AABB> aabbs;
  aabbs.reserve( n );
  for( uint32_t i = 0; i < n; ++i ) { aabbs.emplace_back( Array2s{ boxes[4 * i], boxes[4 * i + 1] }, Array2s{ boxes[4 * i + 2], boxes[4 * i + 3] } ); }
  std::set<std::pair<unsigned,unsigned>> overlaps;
  if( all_pairs ) { SpatialGridDetector::getPotentialOverlapsAllPairs( aabbs, overlaps ); }
  else { SpatialGridDetector::getPotentialOverlaps( aabbs, overlaps ); }
  uint64_t k = 0;
  for( const std::pair<unsigned,unsigned>& p : overlaps ) { if( k < cap ) { ij[2 * k] = p.first; ij[2 * k + 1] = p.second; } ++k; }
  return k;
}

// returns hit (0/1); coeffs[3] = quadratic coefficients, *toi = time of impact when hit
int ref_ball2d_ccd( const double* q0a, const double* q1a, const double ra, const double* q0b, const double* q1b, const double rb, double* coeffs, double* toi )
{
  const Vector3s c{ CollisionDetectionUtilities::computeCCDQuadraticCoeffs( Vector2s{ q0a[0], q0a[1] }, Vector2s{ q1a[0], q1a[1] }, ra, Vector2s{ q0b[0], q0b[1] }, Vector2s{ q1b[0], q1b[1] }, rb ) };
  if( coeffs != nullptr ) { coeffs[0] = c( 0 ); coeffs[1] = c( 1 ); coeffs[2] = c( 2 ); }
  const std::pair<bool,scalar> r{ CollisionDetectionUtilities::ballBallCCDCollisionHappens( c ) };
  if( toi != nullptr ) { *toi = r.first ? r.second : -1.0; }
  return r.first ? 1 : 0;
}

// Ball2DSim::computeBallBallActiveSetSpatialGrid (ball2d/Ball2DSim.cpp:553-608) without the constraint objects:
// swept boxes (restated :566-572) -> getPotentialOverlaps (reference) -> CCD per candidate (reference); counts both lists
// and writes up to cap active pairs in the order they are found (= std::set order of the candidates).
void ref_ball2d_detect( const uint32_t n, const double* q0, const double* q1, const double* r, uint64_t* n_candidates, uint64_t* n_active, uint32_t* active_ij, const uint64_t cap )
{
  std::vector<AABB> aabbs;
  aabbs.reserve( n );
  for( uint32_t i = 0; i < n; ++i )
  {
    const Array2s a{ q0[2 * i], q0[2 * i + 1] };
    const Array2s b{ q1[2 * i], q1[2 * i + 1] };
    aabbs.emplace_back( b.min( a ) - r[i], b.max( a ) + r[i] );
  }
  std::set<std::pair<unsigned,unsigned>> overlaps;
  SpatialGridDetector::getPotentialOverlaps( aabbs, overlaps );
  uint64_t na = 0;
  for( const std::pair<unsigned,unsigned>& p : overlaps )
  {
    const unsigned i = p.first, j = p.second;
    const std::pair<bool,scalar> hit{ CollisionDetectionUtilities::ballBallCCDCollisionHappens( Vector2s{ q0[2 * i], q0[2 * i + 1] }, Vector2s{ q1[2 * i], q1[2 * i + 1] }, r[i],
                                                                                              Vector2s{ q0[2 * j], q0[2 * j + 1] }, Vector2s{ q1[2 * j], q1[2 * j + 1] }, r[j] ) };
    if( hit.first )
    {
      if( active_ij != nullptr && na < cap ) { active_ij[2 * na] = i; active_ij[2 * na + 1] = j; }
      ++na;
    }
  }
  *n_candidates = overlaps.size();
  *n_active = na;
}

// ---- the part of Ball2DState::serialize (ball2d/Ball2DState.cpp:259-272) that the compiled reference classes write themselves:
// m_fixed, m_static_drums, m_static_planes, m_planar_portals, m_forces -- through Utilities::serialize (scisim/Utilities.h:43-94,
// Utilities.cpp:9-18) and the classes' own serialize methods.  Returns the number of bytes (written up to cap).
uint64_t ref_ball2d_snapshot_tail( const uint32_t nballs, const uint32_t ndrums, const double* dx, const double* dr, const uint32_t nplanes, const double* px, const double* pn,
                                   const uint32_t nportals, const double* pax, const double* pan, const double* pbx, const double* pbn, const double* pv, const double* pb, const double t,
                                   const double* g, unsigned char* out, const uint64_t cap )
{
  std::vector<bool> fixed( nballs, false );
  std::vector<StaticDrum> drums;
  for( uint32_t k = 0; k < ndrums; ++k ) { drums.emplace_back( Vector2s{ dx[2 * k], dx[2 * k + 1] }, dr[k] ); }
  std::vector<StaticPlane> planes;
  for( uint32_t k = 0; k < nplanes; ++k ) { planes.emplace_back( Vector2s{ px[2 * k], px[2 * k + 1] }, Vector2s{ pn[2 * k], pn[2 * k + 1] } ); }
  std::vector<PlanarPortal> portals;
  for( uint32_t k = 0; k < nportals; ++k )
  {
    portals.emplace_back( StaticPlane{ Vector2s{ pax[2 * k], pax[2 * k + 1] }, Vector2s{ pan[2 * k], pan[2 * k + 1] } }, StaticPlane{ Vector2s{ pbx[2 * k], pbx[2 * k + 1] }, Vector2s{ pbn[2 * k], pbn[2 * k + 1] } }, pv[k], pb[k] );
    if( pv[k] != 0.0 ) { portals.back().updateMovingPortals( t ); }
  }
  std::vector<std::unique_ptr<Ball2DForce>> forces;
  forces.emplace_back( new Ball2DGravityForce{ Vector2s{ g[0], g[1] } } );
  std::ostringstream stm;
  Utilities::serialize( fixed, stm );
  Utilities::serialize( drums, stm );
  Utilities::serialize( planes, stm );
  Utilities::serialize( portals, stm );
  Utilities::serialize( forces, stm );
  const std::string bytes = stm.str();
  for( uint64_t k = 0; k < bytes.size() && k < cap; ++k ) { out[k] = static_cast<unsigned char>( bytes[k] ); }
  return bytes.size();
}

// ---- PlanarPortal (ball2d/Portals/PlanarPortal.cpp) -----------------------------------------------------------------
void* ref_portal_create( const double* ax, const double* an, const double* bx, const double* bn, const double v, const double bounds )
{
  const StaticPlane a{ Vector2s{ ax[0], ax[1] }, Vector2s{ an[0], an[1] } };
  const StaticPlane b{ Vector2s{ bx[0], bx[1] }, Vector2s{ bn[0], bn[1] } };
  return new PlanarPortal{ a, b, v, bounds };
}
void ref_portal_destroy( void* p ) { delete static_cast<PlanarPortal*>( p ); }
void ref_portal_update( void* p, const double t ) { static_cast<PlanarPortal*>( p )->updateMovingPortals( t ); }
// same layout as orc_ball2d_portal_probe (oracle/capi.cpp)
uint32_t ref_portal_probe( const void* pv, const double* x, const double r, double* out )
{
  const PlanarPortal& p = *static_cast<const PlanarPortal*>( pv );
  const Vector2s xin{ x[0], x[1] };
  Vector2s a, b, tb, ti;
  p.teleportPointThroughPlaneA( xin, a );
  p.teleportPointThroughPlaneB( xin, b );
  p.teleportBall( xin, r, tb );
  p.teleportPointInsidePortal( xin, ti );
  const Vector2s kb{ p.getKinematicVelocityOfBall( xin, r ) };
  const Vector2s kp{ p.getKinematicVelocityOfPoint( xin ) };
  out[0] = a.x(); out[1] = a.y(); out[2] = b.x(); out[3] = b.y(); out[4] = tb.x(); out[5] = tb.y(); out[6] = ti.x(); out[7] = ti.y();
  out[8] = kb.x(); out[9] = kb.y(); out[10] = kp.x(); out[11] = kp.y();
  uint32_t flags = 0u;
  // ballTouchesPortal exits the process when both planes are touched: test that case through the planes themselves
  const bool both = p.planeA().distanceLessThanOrEqualZero( xin, r ) && p.planeB().distanceLessThanOrEqualZero( xin, r );
  if( both ) { flags |= 4u; }
  else
  {
    bool plane_idx = false;
    if( p.ballTouchesPortal( xin, r, plane_idx ) ) { flags |= 1u; if( plane_idx ) { flags |= 2u; } }
  }
  if( p.pointInsidePortal( xin ) ) { flags |= 8u; }
  return flags;
}

}

// ---- the unconstrained maps: a FlowableSystem with just what SymplecticEulerMap / VerletMap ask of Ball2DSim -- the diagonal M and Minv
// ( Minv = 1.0 / m, ball2d/Ball2DState.cpp:54-66 ) and computeForce = setZero + the gravity force ( ball2d/Ball2DSim.cpp:72-78 ) ------------
namespace
{
class ShimBall2DSystem final : public FlowableSystem
{
public:
  ShimBall2DSystem( const uint32_t n, const double* m, const double* r, const double* g )
  : m_force( Vector2s{ g[0], g[1] } )
  {
    std::vector<double> md( 2 * size_t( n ) ), mi( 2 * size_t( n ) );
    for( uint32_t b = 0; b < n; ++b ) { md[2 * b] = md[2 * b + 1] = m[b]; mi[2 * b] = mi[2 * b + 1] = 1.0 / m[b]; }
    m_M.setDiagonal( md.data(), int( md.size() ) ); m_Minv.setDiagonal( mi.data(), int( mi.size() ) );
    m_r.resize( int( n ) );
    for( uint32_t b = 0; b < n; ++b ) { m_r( int( b ) ) = r[b]; }
  }
  virtual int nqdofs() const override { return m_M.rows(); }
  virtual int nvdofs() const override { return m_M.rows(); }
  virtual unsigned numVelDoFsPerBody() const override { return 2; }
  virtual unsigned ambientSpaceDimensions() const override { return 2; }
  virtual bool isKinematicallyScripted( const int ) const override { return false; }
  virtual void computeForce( const VectorXs& q, const VectorXs& v, const scalar& t, VectorXs& F ) override
  {
    F.setZero();
    m_force.computeForce( q, v, m_M, m_r, F );
  }
  virtual void zeroOutForcesOnFixedBodies( VectorXs& ) const override {}
  virtual void linearInertialConfigurationUpdate( const VectorXs&, const VectorXs&, const scalar&, VectorXs& ) const override { std::abort(); }
  virtual const SparseMatrixsc& M() const override { return m_M; }
  virtual const SparseMatrixsc& Minv() const override { return m_Minv; }
  virtual const SparseMatrixsc& M0() const override { return m_M; }
  virtual const SparseMatrixsc& Minv0() const override { return m_Minv; }
  virtual void computeMomentum( const VectorXs&, VectorXs& ) const override { std::abort(); }
  virtual void computeAngularMomentum( const VectorXs&, VectorXs& ) const override { std::abort(); }
  virtual std::string name() const override { return "shim_ball_2d"; }
private:
  SparseMatrixsc m_M, m_Minv;
  VectorXs m_r;
  Ball2DGravityForce m_force;
};
}

extern "C"
{

// kind 0: SymplecticEulerMap::flow, 1: VerletMap::flow; q, v: 2n doubles
void ref_ball2d_flow( const int kind, const uint32_t n, const double* m, const double* r, const double* g, const double* q0, const double* v0, const unsigned iteration, const double dt,
                      double* q1, double* v1 )
{
  ShimBall2DSystem sys{ n, m, r, g };
  VectorXs vq0( int( 2 * n ) ), vv0( int( 2 * n ) ), vq1( int( 2 * n ) ), vv1( int( 2 * n ) );
  for( uint32_t k = 0; k < 2 * n; ++k ) { vq0( int( k ) ) = q0[k]; vv0( int( k ) ) = v0[k]; }
  if( kind == 0 ) { SymplecticEulerMap map; map.flow( vq0, vv0, sys, iteration, dt, vq1, vv1 ); }
  else { VerletMap map; map.flow( vq0, vv0, sys, iteration, dt, vq1, vv1 ); }
  for( uint32_t k = 0; k < 2 * n; ++k ) { q1[k] = vq1( int( k ) ); v1[k] = vv1( int( k ) ); }
}

}


// ---- the constraint classes themselves: what Ball2DSim builds per contact (ball2d/Ball2DSim.cpp:596-607, 730-762) and what the impact maps
// then ask of it (ImpactOperatorUtilities::computeN -> evalgradg, Ball2DSim::computeContactBases -> computeBasis) ---------------------------
extern "C"
{

// kind 0: BallBallConstraint{ i, j, q0, r_i, r_j, false }      geo unused
//      1: StaticPlaneConstraint{ i, r_i, StaticPlane{ x, n }, j }   geo = x[2], n[2]   (n is normalised by StaticPlane's constructor)
//      2: StaticDrumConstraint{ i, q0, r_i, X, j }               geo = X[2], R
// out[0]     the class's isActive at q1 (ball-ball: BallBallConstraint::isActive( i, j, q1, r ), what the portal path tests)
// out[1..2]  getWorldSpaceContactNormal( q0 )      out[3..4] getWorldSpaceContactPoint( q0 )      out[5] penetrationDepth( q1 )
// out[6..9]  computeBasis( q0, v ) column-major [ n | t ]
// grad_rows / grad_vals: the ( row, value ) pairs evalgradg( q0, col = 0, G ) inserts, in insertion order; returns their number
int ref_ball2d_constraint_probe( const int kind, const unsigned i, const unsigned j, const uint32_t nballs, const double* q0, const double* q1, const double* r, const double* geo,
                                 double* out, int* grad_rows, double* grad_vals )
{
  const int nb = int( nballs );
  VectorXs vq0{ 2 * nb }, vq1{ 2 * nb }, vr{ nb }, vv{ 2 * nb };
  for( uint32_t k = 0; k < 2 * nballs; ++k ) { vq0( int( k ) ) = q0[k]; vq1( int( k ) ) = q1[k]; vv( int( k ) ) = 0.0; }
  for( uint32_t k = 0; k < nballs; ++k ) { vr( int( k ) ) = r[k]; }
  std::unique_ptr<Constraint> con;
  bool active = false;
  const StaticPlane plane{ Vector2s{ geo[0], geo[1] }, ( kind == 1 ) ? Vector2s{ geo[2], geo[3] } : Vector2s{ 0.0, 1.0 } }; // outlives the constraint, which keeps a reference to it
  if( kind == 0 )
  {
    active = BallBallConstraint::isActive( i, j, vq1, vr );
    con.reset( new BallBallConstraint{ i, j, vq0, r[i], r[j], false } );
  }
  else if( kind == 1 )
  {
    active = StaticPlaneConstraint::isActive( i, vq1, vr, plane.x(), plane.n() );
    con.reset( new StaticPlaneConstraint{ i, r[i], plane, j } );
  }
  else
  {
    active = StaticDrumConstraint::isActive( i, vq1, vr, Vector2s{ geo[0], geo[1] }, geo[2] );
    con.reset( new StaticDrumConstraint{ i, vq0, r[i], Vector2s{ geo[0], geo[1] }, j } );
  }
  out[0] = active ? 1.0 : 0.0;
  VectorXs n, p;
  con->getWorldSpaceContactNormal( vq0, n );
  con->getWorldSpaceContactPoint( vq0, p );
  out[1] = n( 0 ); out[2] = n( 1 ); out[3] = p( 0 ); out[4] = p( 1 );
  out[5] = con->penetrationDepth( vq1 );
  MatrixXXsc basis;
  con->computeBasis( vq0, vv, basis );
  out[6] = basis( 0, 0 ); out[7] = basis( 1, 0 ); out[8] = basis( 0, 1 ); out[9] = basis( 1, 1 );
  std::vector<double> ones( nballs, 1.0 );
  const double g0[2] = { 0.0, 0.0 };
  ShimBall2DSystem sys{ nballs, ones.data(), r, g0 };
  SparseMatrixsc G{ 2 * int( nballs ), 1 };
  con->evalgradg( vq0, 0, G, sys );
  const int nt = int( G.trip_row.size() );
  for( int k = 0; k < nt; ++k ) { grad_rows[k] = G.trip_row[std::size_t( k )]; grad_vals[k] = G.trip_val[std::size_t( k )]; }
  return nt;
}

}

// ---- ball2d/ConstraintCache.cpp compiled unchanged: cacheConstraint for a list of constraints, then getCachedConstraint for another list.
// Constraints are built with the reference's own classes (only their indices and names matter to the cache).  type = the contact type codes of
// include/scisim_b200.h; a = first body, b = second body / static object.  r: ncomp doubles per constraint.  Returns constraintCacheEmpty() after the stores.
extern "C"
{
int ref_ball2d_cache_roundtrip_ex( const uint32_t nstore, const uint32_t* stype, const uint32_t* sa, const uint32_t* sb, const uint32_t ncomp, const double* rstore,
                                const uint32_t nquery, const uint32_t* qtype, const uint32_t* qa, const uint32_t* qb, double* rout,
                                 void* ser_out, const uint64_t ser_cap, uint64_t* ser_bytes, const void* deser_in, const uint64_t deser_bytes )
{
  uint32_t nb = 1;
  for( uint32_t k = 0; k < nstore; ++k ) { nb = std::max( nb, std::max( sa[k], stype[k] == 0 ? sb[k] : 0u ) + 1u ); }
  for( uint32_t k = 0; k < nquery; ++k ) { nb = std::max( nb, std::max( qa[k], qtype[k] == 0 ? qb[k] : 0u ) + 1u ); }
  VectorXs wq{ int( 2 * nb ) };
  for( uint32_t b = 0; b < nb; ++b ) { wq( int( 2 * b ) ) = double( b ); wq( int( 2 * b + 1 ) ) = 0.5 * double( b ); }
  const VectorXs& q = wq;
  const StaticPlane plane{ Vector2s{ 0.0, 0.0 }, Vector2s{ 0.0, 1.0 } };
  const auto make = [&]( const uint32_t type, const uint32_t a, const uint32_t b ) -> std::unique_ptr<Constraint>
  {
    if( type == 0 ) { return std::unique_ptr<Constraint>{ new BallBallConstraint{ a, b, q, 0.5, 0.5, false } }; }
    if( type == 2 ) { return std::unique_ptr<Constraint>{ new StaticPlaneConstraint{ a, 0.5, plane, b } }; }
    return std::unique_ptr<Constraint>{ new StaticDrumConstraint{ a, q, 0.5, Vector2s{ -3.0, -4.0 }, b } };
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
int ref_ball2d_cache_roundtrip( const uint32_t nstore, const uint32_t* stype, const uint32_t* sa, const uint32_t* sb, const uint32_t ncomp, const double* rstore,
                                const uint32_t nquery, const uint32_t* qtype, const uint32_t* qa, const uint32_t* qb, double* rout )
{
  return ref_ball2d_cache_roundtrip_ex( nstore, stype, sa, sb, ncomp, rstore, nquery, qtype, qa, qb, rout, nullptr, 0, nullptr, nullptr, 0 );
}
}
