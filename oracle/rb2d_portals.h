// oracle/rb2d_portals.h
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). CPU restatement of rigidbody2d's periodic / Lees-Edwards
// boundary path:
//   rigidbody2d/RigidBody2DStaticPlane.cpp:10-14,84-92    plane frame ( n as given, t = ( -n.y, n.x ) ), distanceToPoint
//   rigidbody2d/PlanarPortal.cpp:143-246                   aabbInHalfPlane / aabbTouchesPortal ( plane A first; the both-planes check is
//                                                          debug-only ), teleportPoint, getKinematicVelocityOfAABB / OfPoint
//   rigidbody2d/PlanarPortal.cpp:298-350                   teleportPointThroughPlaneA / B, updateMovingPortals ( same arithmetic as ball2d's )
//   rigidbody2d/CircleGeometry.cpp:38-43, BoxGeometry.cpp:32-42   computeAABB at one configuration
//   rigidbody2d/RigidBody2DSim.cpp:350-411                 collisionIsActive ( circles only; any box exits )
//   rigidbody2d/RigidBody2DSim.cpp:413-636                 teleportedCollisionIsActive, getTeleportedCollisionCenters, dispatchTeleportedNarrowPhaseCollision
//   rigidbody2d/RigidBody2DSim.cpp:696-714, 836-874        computeActiveSet, enforcePeriodicBoundaryConditions
//   rigidbody2d/RigidBody2DSim.cpp:876-1040                computeBodyBodyActiveSetSpatialGridWithPortals
//   rigidbody2d/TeleportedCircleCircleConstraint.cpp:10-29,155-158,177-182   normal, relative arms, contact point q0_i + r0, depth
//   rigidbody2d/KinematicKickCircleCircleConstraint.cpp:12-16
// Parity: unpinned by stored reference outputs; the portal primitives are checked bit for bit against the reference's own
// rigidbody2d/PlanarPortal.cpp compiled unchanged (oracle/_ref, tests/test_portals_cpu.py).
#ifndef ORACLE_RB2D_PORTALS_H
#define ORACLE_RB2D_PORTALS_H

#include "ball2d_portals.h"
#include "rb2d.h"

namespace orc
{

enum RB2DPortalContactType : uint32_t { CIRCLE_CIRCLE_TELEPORTED = 25, CIRCLE_CIRCLE_KICK_TELEPORTED = 26 };

// RigidBody2DStaticPlane::RigidBody2DStaticPlane( x, n ): the normal is NOT normalised here
inline Plane2D makePlaneRB2D( const V2& x, const V2& n )
{
  Plane2D p;
  p.x = x; p.n = n; p.t = V2{ -n.y, n.x };
  return p;
}

// CircleGeometry::computeAABB / BoxGeometry::computeAABB at one configuration
inline void computeAABBAt( const RB2DGeometry& g, const V2& x, const double theta, Box<2>& box )
{
  if( g.type == GEO2_CIRCLE )
  {
    box.lo[0] = x.x - g.r; box.lo[1] = x.y - g.r; box.hi[0] = x.x + g.r; box.hi[1] = x.y + g.r;
  }
  else
  {
    const M2 R = rot2( theta );
    const V2 e{ std::fabs( R.a ) * g.half.x + std::fabs( R.b ) * g.half.y, std::fabs( R.c ) * g.half.x + std::fabs( R.d ) * g.half.y };
    box.lo[0] = x.x - e.x; box.lo[1] = x.y - e.y; box.hi[0] = x.x + e.x; box.hi[1] = x.y + e.y;
  }
}

// PlanarPortal.cpp:143-165
inline bool aabbInHalfPlane( const Box<2>& b, const Plane2D& plane )
{
  if( dot( plane.n, V2{ b.lo[0], b.lo[1] } - plane.x ) <= 0 ) { return true; }
  if( dot( plane.n, V2{ b.lo[0], b.hi[1] } - plane.x ) <= 0 ) { return true; }
  if( dot( plane.n, V2{ b.hi[0], b.lo[1] } - plane.x ) <= 0 ) { return true; }
  if( dot( plane.n, V2{ b.hi[0], b.hi[1] } - plane.x ) <= 0 ) { return true; }
  return false;
}
// PlanarPortal.cpp:167-189 (release build): 0 = no, 1 = plane A, 2 = plane B
inline int aabbTouchesPortal( const Portal2D& p, const Box<2>& b )
{
  if( aabbInHalfPlane( b, p.a ) ) { return 1; }
  if( aabbInHalfPlane( b, p.b ) ) { return 2; }
  return 0;
}
// PlanarPortal.cpp:218-230
inline V2 getKinematicVelocityOfAABB( const Portal2D& p, const Box<2>& b )
{
  return aabbInHalfPlane( b, p.a ) ? ( -p.v ) * p.a.t : ( -p.v ) * p.b.t;
}

// RigidBody2DSim::enforcePeriodicBoundaryConditions (rigidbody2d/RigidBody2DSim.cpp:841-874): portal-major, in place, [x, y, theta] layout
inline void enforcePeriodicBoundaryConditionsRB2D( const std::vector<Portal2D>& portals, const uint32_t nb, double* q, double* v )
{
  for( const Portal2D& p : portals )
  {
    for( uint32_t b = 0; b < nb; ++b )
    {
      const V2 xin{ q[3 * b], q[3 * b + 1] };
      if( pointInsidePortal( p, xin ) )
      {
        const V2 xout = teleportPointInsidePortal( p, xin );
        q[3 * b] = xout.x; q[3 * b + 1] = xout.y;
        if( isLeesEdwards( p ) )
        {
          const V2 dv = getKinematicVelocityOfPoint( p, xin );
          v[3 * b] = v[3 * b] + dv.x; v[3 * b + 1] = v[3 * b + 1] + dv.y;
        }
      }
    }
  }
}

struct RB2DTeleportedInfo
{
  uint32_t p0, p1;
  bool pl0, pl1;
  V2 x0, x1;         // teleported centres at q0
  V2 delta0, delta1; // x_t0 - q0 of each body ( NaN for kinematic-kick contacts, as the reference's constructor stores )
  V2 kick;
};

struct RB2DPortalResult
{
  std::vector<RB2DContact> active;                      // regular (dispatchNarrowPhaseCollision) | teleported | planes
  std::vector<std::pair<unsigned,unsigned>> candidates; // extended index space
  std::vector<TeleportedBall2D> teleported_boxes;
  std::vector<RB2DTeleportedInfo> teleported_info;
  uint64_t n_regular = 0;                               // contacts produced by un-teleported pairs
  bool supported = true;                                // false where the reference exits
};

// RigidBody2DSim::computeActiveSet with portals (rigidbody2d/RigidBody2DSim.cpp:696-714 -> :876-1040, then the planes)
inline void computeActiveSetWithPortalsRB2D( const RB2DScene& s, const std::vector<Portal2D>& portals, const double* q0, const double* q1, RB2DPortalResult& res, const bool use_grid = true )
{
  const uint32_t nb = uint32_t( s.nbodies() );
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  res = RB2DPortalResult{};
  auto X = [&]( const double* q, unsigned b ) { return V2{ q[3 * b], q[3 * b + 1] }; };
  PairSet possible_overlaps;
  std::map<unsigned,TeleportedBall2D> teleported_aabb_body_indices;
  {
    std::vector<Box<2>> aabbs( nb );
    for( uint32_t b = 0; b < nb; ++b ) { computeAABBAt( s.geo( b ), X( q1, b ), q1[3 * b + 2], aabbs[b] ); }
    for( uint32_t p = 0; p < uint32_t( portals.size() ); ++p )
    {
      for( uint32_t b = 0; b < nb; ++b )
      {
        const int touch = aabbTouchesPortal( portals[p], aabbs[b] );
        if( touch != 0 )
        {
          const V2 xo = touch == 1 ? teleportPointThroughPlaneA( portals[p], X( q1, b ) ) : teleportPointThroughPlaneB( portals[p], X( q1, b ) );
          Box<2> bx;
          computeAABBAt( s.geo( b ), xo, q1[3 * b + 2], bx );
          aabbs.push_back( bx );
          const TeleportedBall2D tb{ b, p, touch == 2 };
          teleported_aabb_body_indices.insert( std::make_pair( unsigned( aabbs.size() - 1 ), tb ) );
          res.teleported_boxes.push_back( tb );
        }
      }
    }
    if( !aabbs.empty() )
    {
      if( use_grid ) { getPotentialOverlaps<2>( aabbs, possible_overlaps ); }
      else { getPotentialOverlapsAllPairs<2>( aabbs, possible_overlaps ); }
    }
  }
  res.candidates.assign( possible_overlaps.begin(), possible_overlaps.end() );

  // collisionIsActive (RigidBody2DSim.cpp:350-397): circles only, everything else exits
  auto circlesOnly = [&]( const unsigned a, const unsigned b ) { return s.geo( a ).type == GEO2_CIRCLE && s.geo( b ).type == GEO2_CIRCLE; };

  std::set<TeleportedCollision2D> teleported_collisions;
  for( const auto& pr : possible_overlaps )
  {
    const bool first_teleported = pr.first >= nb;
    const bool second_teleported = pr.second >= nb;
    if( !first_teleported && !second_teleported )
    {
      if( !dispatchNarrowPhaseCollision( s, pr.first, pr.second, q0, q1, res.active ) ) { res.supported = false; return; }
      continue;
    }
    uint32_t bdy0 = pr.first, bdy1 = pr.second, prtl0 = NO_PORTAL, prtl1 = NO_PORTAL;
    bool plane0 = false, plane1 = false;
    if( first_teleported ) { const TeleportedBall2D& tb = teleported_aabb_body_indices.find( pr.first )->second; bdy0 = tb.body; prtl0 = tb.portal; plane0 = tb.plane; }
    if( second_teleported ) { const TeleportedBall2D& tb = teleported_aabb_body_indices.find( pr.second )->second; bdy1 = tb.body; prtl1 = tb.portal; plane1 = tb.plane; }
    if( first_teleported && second_teleported )
    {
      if( !circlesOnly( bdy0, bdy1 ) ) { res.supported = false; return; }
      if( ballBallIsActive( X( q1, bdy0 ), X( q1, bdy1 ), s.geo( bdy0 ).r, s.geo( bdy1 ).r ) ) { continue; }
    }
    if( s.fixed[bdy0] && s.fixed[bdy1] ) { continue; }
    const TeleportedCollision2D tc{ bdy0, bdy1, prtl0, prtl1, plane0, plane1 };
    if( !circlesOnly( tc.b0, tc.b1 ) ) { res.supported = false; return; }
    V2 x0, x1;
    // getTeleportedCollisionCenters reads the [x, y, theta] layout
    x0 = X( q1, tc.b0 ); if( tc.p0 != NO_PORTAL ) { x0 = tc.pl0 == 0 ? teleportPointThroughPlaneA( portals[tc.p0], x0 ) : teleportPointThroughPlaneB( portals[tc.p0], x0 ); }
    x1 = X( q1, tc.b1 ); if( tc.p1 != NO_PORTAL ) { x1 = tc.pl1 == 0 ? teleportPointThroughPlaneA( portals[tc.p1], x1 ) : teleportPointThroughPlaneB( portals[tc.p1], x1 ); }
    if( ballBallIsActive( x0, x1, s.geo( tc.b0 ).r, s.geo( tc.b1 ).r ) ) { teleported_collisions.insert( tc ); }
  }
  res.n_regular = res.active.size();

  // dispatchTeleportedNarrowPhaseCollision (RigidBody2DSim.cpp:477-636)
  for( const TeleportedCollision2D& tc : teleported_collisions )
  {
    if( s.fixed[tc.b0] || s.fixed[tc.b1] ) { res.supported = false; return; }
    auto centre = [&]( const double* q, const uint32_t b, const uint32_t p, const bool pl )
    {
      V2 x = X( q, b );
      if( p != NO_PORTAL ) { x = pl == 0 ? teleportPointThroughPlaneA( portals[p], x ) : teleportPointThroughPlaneB( portals[p], x ); }
      return x;
    };
    RB2DTeleportedInfo info;
    info.p0 = tc.p0; info.p1 = tc.p1; info.pl0 = tc.pl0; info.pl1 = tc.pl1;
    info.x0 = centre( q0, tc.b0, tc.p0, tc.pl0 ); info.x1 = centre( q0, tc.b1, tc.p1, tc.pl1 );
    info.delta0 = info.x0 - X( q0, tc.b0 ); info.delta1 = info.x1 - X( q0, tc.b1 );
    const V2 x0_t1 = centre( q1, tc.b0, tc.p0, tc.pl0 ), x1_t1 = centre( q1, tc.b1, tc.p1, tc.pl1 );
    const bool le0 = tc.p0 != NO_PORTAL && isLeesEdwards( portals[tc.p0] );
    const bool le1 = tc.p1 != NO_PORTAL && isLeesEdwards( portals[tc.p1] );
    info.kick = V2{ 0.0, 0.0 };
    if( le0 || le1 )
    {
      if( le1 ) { Box<2> bx; computeAABBAt( s.geo( tc.b1 ), X( q1, tc.b1 ), q1[3 * tc.b1 + 2], bx ); info.kick = getKinematicVelocityOfAABB( portals[tc.p1], bx ); }
      else { Box<2> bx; computeAABBAt( s.geo( tc.b0 ), X( q1, tc.b0 ), q1[3 * tc.b0 + 2], bx ); const V2 k = getKinematicVelocityOfAABB( portals[tc.p0], bx ); info.kick = V2{ -k.x, -k.y }; }
      info.delta0 = V2{ NaN, NaN }; info.delta1 = V2{ NaN, NaN };
    }
    const double r0 = s.geo( tc.b0 ).r, r1 = s.geo( tc.b1 ).r;
    if( !ballBallIsActive( x0_t1, x1_t1, r0, r1 ) ) { continue; }
    RB2DContact c;
    c.type = ( le0 || le1 ) ? CIRCLE_CIRCLE_KICK_TELEPORTED : CIRCLE_CIRCLE_TELEPORTED;
    c.i = tc.b0; c.j = tc.b1; c.aux = 0;
    c.n = normalized( info.x0 - info.x1 );
    // getWorldSpaceContactPoint( q0 ) = q0_i + m_r0, m_r0 = ( r0 / ( r0 + r1 ) ) * ( x1 - x0 )
    c.p = X( q0, tc.b0 ) + ( r0 / ( r0 + r1 ) ) * ( info.x1 - info.x0 );
    // TeleportedCircleCircleConstraint::computePenetrationDepth( q1 ) (TeleportedCircleCircleConstraint.cpp:177-182); the kinematic-kick
    // variant stores NaN displacements and radii, and std::min( 0.0, NaN ) is 0.0
    if( le0 || le1 ) { c.depth = 0.0; }
    else { c.depth = std::min( 0.0, norm( ( X( q1, tc.b0 ) + info.delta0 ) - ( X( q1, tc.b1 ) + info.delta1 ) ) - r0 - r1 ); }
    res.active.emplace_back( c );
    res.teleported_info.push_back( info );
  }
  computeBodyPlaneActiveSet( s, q0, q1, res.active );
}

}

#endif
