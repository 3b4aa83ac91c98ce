// oracle/ball2d_portals.h
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). CPU restatement of ball2d's periodic / Lees-Edwards
// boundary path:
//   ball2d/StaticGeometry/StaticPlane.cpp:10-14,66-74            plane frame ( n normalised, t = ( -n.y, n.x ) ), signed-distance tests
//   ball2d/Portals/PlanarPortal.cpp:36-57                        TeleportedCollision (bodies ordered, compared by body pair only)
//   ball2d/Portals/PlanarPortal.cpp:107-242                      PlanarPortal (touch tests, teleports, kinematic velocities, updateMovingPortals)
//   ball2d/Ball2DSim.cpp:327-366                                 updatePeriodicBoundaryConditionsStartOfStep, enforcePeriodicBoundaryConditions
//   ball2d/Ball2DSim.cpp:368-546                                 computeBallBallActiveSetSpatialGridWithPortals
//   ball2d/Ball2DSim.cpp:610-728                                 getTeleportedBallBallCenters, teleportedBallBallCollisionHappens, generateTeleportedBallBallCollision
//   ball2d/Constraints/BallBallConstraint.cpp:16-37,219-222,270-280   isActive, normal, contact point, depth ( NaN when teleported )
//   ball2d/Constraints/KinematicKickBallBallConstraint.cpp:8-11  the kick carried by Lees-Edwards contacts
// Parity: unpinned by stored reference outputs (none exist for this path); the portal arithmetic is checked bit for bit
// against the reference's own PlanarPortal.cpp / StaticPlane.cpp compiled unchanged (oracle/_ref, tests/test_portals_cpu.py).
#ifndef ORACLE_BALL2D_PORTALS_H
#define ORACLE_BALL2D_PORTALS_H

#include "ball2d.h"

#include <climits>
#include <map>
#include <set>
#include <tuple>

namespace orc
{

// int( floor( x ) ) as the reference's x86-64 build evaluates it: cvttsd2si returns INT_MIN for NaN and for values outside
// the int range (plain portals have bounds == 0, so the reference divides by zero here and multiplies the result by 0)
inline int toIntX86( const double x )
{
  return ( x >= -2147483648.0 && x < 2147483648.0 ) ? int( x ) : INT_MIN;
}

// StaticPlane (ball2d/StaticGeometry/StaticPlane.cpp:10-14)
struct Plane2D
{
  V2 x, n, t;
};
inline Plane2D makePlane2D( const V2& x, const V2& n )
{
  Plane2D p;
  p.x = x;
  p.n = normalized( n );
  p.t = V2{ -p.n.y, p.n.x };
  return p;
}
// StaticPlane.cpp:66-74
inline bool distanceLessThanZero( const Plane2D& p, const V2& x ) { return dot( p.n, x - p.x ) < 0.0; }
inline bool distanceLessThanOrEqualZero( const Plane2D& p, const V2& x, const double r ) { return dot( p.n, x - p.x ) <= r; }

struct Portal2D
{
  Plane2D a, b;
  double v = 0.0;      // tangential velocity of the Lees-Edwards pair ( 0: plain periodic portal )
  double bounds = 0.0; // half length of the periodic tangential coordinate ( 0 for plain portals )
  double dx = 0.0;     // current tangential offset, set by updateMovingPortals
};

// PlanarPortal.cpp:231-236
inline void updateMovingPortals( Portal2D& p, const double t )
{
  const int repeat_x = toIntX86( std::floor( ( p.v * t + p.bounds ) / ( 2.0 * p.bounds ) ) );
  p.dx = p.v * t - 2.0 * repeat_x * p.bounds;
}
inline bool isLeesEdwards( const Portal2D& p ) { return p.v != 0.0; }

// PlanarPortal.cpp:186-219: teleport through the plane `from` onto the plane `to`
inline V2 teleportThrough( const Portal2D& p, const Plane2D& from, const Plane2D& to, const V2& xin )
{
  const double nA = dot( from.n, from.x - xin );
  double tA = dot( from.t, ( p.dx * from.t + from.x ) - xin );
  const int repeat_x = toIntX86( std::floor( ( tA + p.bounds ) / ( 2.0 * p.bounds ) ) );
  tA -= 2.0 * repeat_x * p.bounds;
  return ( to.x + nA * to.n ) + tA * to.t;
}
inline V2 teleportPointThroughPlaneA( const Portal2D& p, const V2& xin ) { return teleportThrough( p, p.a, p.b, xin ); }
inline V2 teleportPointThroughPlaneB( const Portal2D& p, const V2& xin ) { return teleportThrough( p, p.b, p.a, xin ); }

// PlanarPortal.cpp:107-110
inline bool pointInsidePortal( const Portal2D& p, const V2& x ) { return distanceLessThanZero( p.a, x ) || distanceLessThanZero( p.b, x ); }

// PlanarPortal.cpp:112-133.  Returns 0 = no touch, 1 = plane A, 2 = plane B, -1 = both ( the reference prints and exits )
inline int ballTouchesPortal( const Portal2D& p, const V2& x, const double r )
{
  const bool ta = distanceLessThanOrEqualZero( p.a, x, r );
  const bool tb = distanceLessThanOrEqualZero( p.b, x, r );
  if( ta && tb ) { return -1; }
  if( ta ) { return 1; }
  if( tb ) { return 2; }
  return 0;
}
// PlanarPortal.cpp:135-149
inline V2 teleportPointInsidePortal( const Portal2D& p, const V2& xin )
{
  return distanceLessThanZero( p.a, xin ) ? teleportPointThroughPlaneA( p, xin ) : teleportPointThroughPlaneB( p, xin );
}
// PlanarPortal.cpp:151-165
inline V2 teleportBall( const Portal2D& p, const V2& xin, const double r )
{
  return distanceLessThanOrEqualZero( p.a, xin, r ) ? teleportPointThroughPlaneA( p, xin ) : teleportPointThroughPlaneB( p, xin );
}
// PlanarPortal.cpp:167-179
inline V2 getKinematicVelocityOfBall( const Portal2D& p, const V2& x, const double r )
{
  return distanceLessThanOrEqualZero( p.a, x, r ) ? ( -p.v ) * p.a.t : ( -p.v ) * p.b.t;
}
// PlanarPortal.cpp:181-193
inline V2 getKinematicVelocityOfPoint( const Portal2D& p, const V2& x )
{
  return distanceLessThanZero( p.a, x ) ? ( -p.v ) * p.a.t : ( -p.v ) * p.b.t;
}

// Ball2DSim::enforcePeriodicBoundaryConditions (ball2d/Ball2DSim.cpp:336-366): portal-major, in place
inline void enforcePeriodicBoundaryConditions( const std::vector<Portal2D>& portals, const uint32_t nb, double* q, double* v )
{
  for( const Portal2D& p : portals )
  {
    for( uint32_t b = 0; b < nb; ++b )
    {
      const V2 xin{ q[2 * b], q[2 * b + 1] };
      if( pointInsidePortal( p, xin ) )
      {
        const V2 xout = teleportPointInsidePortal( p, xin );
        q[2 * b] = xout.x; q[2 * b + 1] = xout.y;
        if( isLeesEdwards( p ) )
        {
          const V2 dv = getKinematicVelocityOfPoint( p, xin );
          v[2 * b] = v[2 * b] + dv.x; v[2 * b + 1] = v[2 * b + 1] + dv.y;
        }
      }
    }
  }
}

// contact types beyond ball2d.h's: the two constraints generateTeleportedBallBallCollision creates
enum Ball2DPortalContactType : uint32_t { BALL_BALL_TELEPORTED = 3, BALL_BALL_KICK_TELEPORTED = 4 };

constexpr uint32_t NO_PORTAL = std::numeric_limits<uint32_t>::max();

// TeleportedBall (PlanarPortal.cpp:12-32)
struct TeleportedBall2D { uint32_t body, portal; bool plane; };

// TeleportedCollision (PlanarPortal.cpp:34-57)
struct TeleportedCollision2D
{
  uint32_t b0, b1, p0, p1;
  bool pl0, pl1;
  TeleportedCollision2D( uint32_t body0, uint32_t body1, uint32_t portal0, uint32_t portal1, bool plane0, bool plane1 )
  : b0( body0 ), b1( body1 ), p0( portal0 ), p1( portal1 ), pl0( plane0 ), pl1( plane1 )
  {
    if( b0 > b1 ) { std::swap( b0, b1 ); std::swap( p0, p1 ); std::swap( pl0, pl1 ); }
  }
  bool operator<( const TeleportedCollision2D& rhs ) const { return std::tie( b0, b1 ) < std::tie( rhs.b0, rhs.b1 ); }
};

// extra data of a teleported contact (what the constraint constructors receive)
struct TeleportedContactInfo
{
  uint32_t p0, p1; // portal of each body ( NO_PORTAL: not teleported )
  bool pl0, pl1;
  V2 x0, x1;       // teleported centres at q0
  V2 kick;         // KinematicKickBallBallConstraint's kick ( 0 for plain portals )
};

// Ball2DSim::getTeleportedBallBallCenters (ball2d/Ball2DSim.cpp:610-642)
inline void getTeleportedBallBallCenters( const std::vector<Portal2D>& portals, const double* q, const TeleportedCollision2D& tc, V2& x0, V2& x1 )
{
  x0 = V2{ q[2 * tc.b0], q[2 * tc.b0 + 1] };
  if( tc.p0 != NO_PORTAL ) { x0 = tc.pl0 == 0 ? teleportPointThroughPlaneA( portals[tc.p0], x0 ) : teleportPointThroughPlaneB( portals[tc.p0], x0 ); }
  x1 = V2{ q[2 * tc.b1], q[2 * tc.b1 + 1] };
  if( tc.p1 != NO_PORTAL ) { x1 = tc.pl1 == 0 ? teleportPointThroughPlaneA( portals[tc.p1], x1 ) : teleportPointThroughPlaneB( portals[tc.p1], x1 ); }
}

// BallBallConstraint::isActive (ball2d/Constraints/BallBallConstraint.cpp:16-20)
inline bool ballBallIsActive( const V2& x0, const V2& x1, const double r0, const double r1 )
{
  return squaredNorm( x0 - x1 ) <= ( r0 + r1 ) * ( r0 + r1 );
}

struct PortalActiveSetResult
{
  std::vector<Ball2DContact> active;                      // regular | teleported | drums | planes
  std::vector<std::pair<unsigned,unsigned>> candidates;   // in the extended index space ( teleported boxes follow the nb real ones )
  std::vector<TeleportedBall2D> teleported_boxes;         // box nb + k
  std::vector<TeleportedContactInfo> teleported_info;     // one per teleported contact, in active-set order
  uint64_t n_regular = 0;
  bool both_planes_touched = false;                       // the reference exits ( PlanarPortal.cpp:117-121 )
};

// Ball2DSim::computeActiveSet with portals (ball2d/Ball2DSim.cpp:151-173 -> :368-546, then drums and planes)
inline void computeActiveSetWithPortals( const Ball2DScene& s, const std::vector<Portal2D>& portals, const double* q0, const double* q1, PortalActiveSetResult& res, const bool use_grid = true )
{
  const uint32_t nb = uint32_t( s.r.size() );
  res = PortalActiveSetResult{};
  PairSet possible_overlaps;
  std::map<unsigned,TeleportedBall2D> teleported_aabb_body_indices;
  {
    std::vector<Box<2>> aabbs;
    aabbs.reserve( nb );
    for( uint32_t b = 0; b < nb; ++b )
    {
      Box<2> bx;
      for( int k = 0; k < 2; ++k ) { bx.lo[k] = q1[2 * b + k] - s.r[b]; bx.hi[k] = q1[2 * b + k] + s.r[b]; }
      aabbs.push_back( bx );
    }
    for( uint32_t p = 0; p < uint32_t( portals.size() ); ++p )
    {
      for( uint32_t b = 0; b < nb; ++b )
      {
        const V2 x{ q1[2 * b], q1[2 * b + 1] };
        const int touch = ballTouchesPortal( portals[p], x, s.r[b] );
        if( touch < 0 ) { res.both_planes_touched = true; return; }
        if( touch != 0 )
        {
          const V2 xo = teleportBall( portals[p], x, s.r[b] );
          Box<2> bx;
          bx.lo[0] = xo.x - s.r[b]; bx.lo[1] = xo.y - s.r[b]; bx.hi[0] = xo.x + s.r[b]; bx.hi[1] = xo.y + s.r[b];
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

  std::set<TeleportedCollision2D> teleported_collisions;
  for( const auto& pr : possible_overlaps )
  {
    const bool first_teleported = pr.first >= nb;
    const bool second_teleported = pr.second >= nb;
    if( !first_teleported && !second_teleported )
    {
      const uint32_t a = pr.first, b = pr.second;
      if( ballBallIsActive( V2{ q1[2 * a], q1[2 * a + 1] }, V2{ q1[2 * b], q1[2 * b + 1] }, s.r[a], s.r[b] ) )
      {
        res.active.emplace_back( makeBallBall( a, b, q0, q1, s.r[a], s.r[b] ) );
      }
    }
    else
    {
      uint32_t bdy0 = pr.first, bdy1 = pr.second, prtl0 = NO_PORTAL, prtl1 = NO_PORTAL;
      bool plane0 = false, plane1 = false;
      if( first_teleported ) { const TeleportedBall2D& tb = teleported_aabb_body_indices.find( pr.first )->second; bdy0 = tb.body; prtl0 = tb.portal; plane0 = tb.plane; }
      if( second_teleported ) { const TeleportedBall2D& tb = teleported_aabb_body_indices.find( pr.second )->second; bdy1 = tb.body; prtl1 = tb.portal; plane1 = tb.plane; }
      // both copies teleported: the collision is also found between the un-teleported bodies
      if( first_teleported && second_teleported )
      {
        if( ballBallIsActive( V2{ q1[2 * bdy0], q1[2 * bdy0 + 1] }, V2{ q1[2 * bdy1], q1[2 * bdy1 + 1] }, s.r[bdy0], s.r[bdy1] ) ) { continue; }
      }
      const TeleportedCollision2D tc{ bdy0, bdy1, prtl0, prtl1, plane0, plane1 };
      V2 x0, x1;
      getTeleportedBallBallCenters( portals, q1, tc, x0, x1 );
      if( ballBallIsActive( x0, x1, s.r[tc.b0], s.r[tc.b1] ) ) { teleported_collisions.insert( tc ); }
    }
  }
  res.n_regular = res.active.size();

  // generateTeleportedBallBallCollision (ball2d/Ball2DSim.cpp:653-728)
  for( const TeleportedCollision2D& tc : teleported_collisions )
  {
    TeleportedContactInfo info;
    info.p0 = tc.p0; info.p1 = tc.p1; info.pl0 = tc.pl0; info.pl1 = tc.pl1;
    getTeleportedBallBallCenters( portals, q0, tc, info.x0, info.x1 );
    const double ri = s.r[tc.b0], rj = s.r[tc.b1];
    const bool le0 = tc.p0 != NO_PORTAL && isLeesEdwards( portals[tc.p0] );
    const bool le1 = tc.p1 != NO_PORTAL && isLeesEdwards( portals[tc.p1] );
    Ball2DContact c;
    c.i = tc.b0; c.j = tc.b1;
    c.n = normalized( info.x0 - info.x1 );
    c.p = V2{ q0[2 * tc.b0], q0[2 * tc.b0 + 1] } - ri * c.n; // getWorldSpaceContactPoint( q0 ) uses the body's own position
    c.depth = std::numeric_limits<double>::quiet_NaN();
    info.kick = V2{ 0.0, 0.0 };
    if( !le0 && !le1 ) { c.type = BALL_BALL_TELEPORTED; }
    else
    {
      c.type = BALL_BALL_KICK_TELEPORTED;
      if( le1 ) { info.kick = getKinematicVelocityOfBall( portals[tc.p1], V2{ q1[2 * tc.b1], q1[2 * tc.b1 + 1] }, rj ); }
      else
      {
        const V2 k = getKinematicVelocityOfBall( portals[tc.p0], V2{ q1[2 * tc.b0], q1[2 * tc.b0 + 1] }, ri );
        info.kick = V2{ -k.x, -k.y };
      }
    }
    res.active.emplace_back( c );
    res.teleported_info.push_back( info );
  }

  // drums and planes exactly as without portals (ball2d/Ball2DSim.cpp:167-172)
  {
    std::vector<Ball2DContact> statics;
    computeStaticActiveSet( s, q0, q1, statics );
    res.active.insert( res.active.end(), statics.begin(), statics.end() );
  }
}

}

#endif
