// oracle/rb3d_portals.h
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). CPU restatement of rigidbody3d's periodic boundary path
//  -- ORACLE FIRST: the GPU side of this row is not built yet.
//   rigidbody3d/StaticGeometry/StaticPlane.cpp:10-15,24-27,48-66   plane: n normalised; t0 / t1 = FromTwoVectors( UnitY, n ) * UnitX / UnitZ
//   Eigen 3.3.4 Quaternion::setFromTwoVectors / _transformVector     ( restated in rotateFromUnitY below; not part of the reference tree )
//   rigidbody3d/Portals/PlanarPortal.cpp:107-185                     pointInsidePortal, aabbInHalfPlane ( 8 corners ), aabbTouchesPortal ( plane A first;
//                                                                    the both-planes check is debug-only ), teleportPointInsidePortal, teleportPoint,
//                                                                    teleportPointThroughPlaneA / B with the integer portal multiplier
//   rigidbody3d/RigidBody3DSim.cpp:642-663                           enforcePeriodicBoundaryConditions ( positions only: no Lees-Edwards in 3-D )
//   rigidbody3d/RigidBody3DSim.cpp:965-1056                          collisionIsActive ( = does the regular narrow phase produce a contact )
//   rigidbody3d/RigidBody3DSim.cpp:1072-1260                         computeActiveSetBodyBodySpatialGrid ( this IS the body-body path; zero portals = rb3d.h )
//   rigidbody3d/RigidBody3DSim.cpp:1262-1397                         teleportedCollisionHappens, getTeleportedCollisionCenters, generateTeleportedCollision
//   rigidbody3d/Constraints/TeleportedSphereSphereConstraint.cpp:14-27,318-321, KinematicObjectSphereConstraint.cpp:10-22
// Parity: unpinned by stored reference outputs; plane frames and portal primitives are checked bit for bit against the reference's
// own StaticPlane.cpp / PlanarPortal.cpp compiled unchanged (oracle/_ref; the Eigen stand-in carries the same Quaternion restatement).
#ifndef ORACLE_RB3D_PORTALS_H
#define ORACLE_RB3D_PORTALS_H

#include "ball2d_portals.h" // NO_PORTAL, TeleportedBall2D ( body, portal, plane ), TeleportedCollision2D ( ordering by body pair )
#include "rb3d.h"

namespace orc
{

enum RB3DPortalContactType : uint32_t { SPHERE_SPHERE_TELEPORTED = 19, KINEMATIC_OBJECT_SPHERE_TELEPORTED = 30 };

struct Plane3D { V3 x, n, t0, t1; };

// Quaternion::FromTwoVectors( UnitY, n ) * v  (Eigen 3.3.4: setFromTwoVectors + _transformVector); n is already unit length but Eigen
// normalises both arguments again.  The nearly-opposite branch ( n ~ -UnitY ) runs an SVD in Eigen and is not restated: NaN.
inline V3 rotateFromUnitY( const V3& n, const V3& v )
{
  const V3 v0 = normalized( V3{ 0.0, 1.0, 0.0 } );
  const V3 v1 = normalized( n );
  const double c = dot( v1, v0 );
  if( c < -1.0 + 1e-12 ) { const double nan = std::numeric_limits<double>::quiet_NaN(); return V3{ nan, nan, nan }; }
  const V3 axis = cross( v0, v1 );
  const double s = std::sqrt( ( 1.0 + c ) * 2.0 );
  const double invs = 1.0 / s;
  const V3 vec = V3{ axis.x * invs, axis.y * invs, axis.z * invs };
  const double w = s * 0.5;
  V3 uv = cross( vec, v );
  uv = uv + uv;
  return ( v + w * uv ) + cross( vec, uv );
}

inline Plane3D makePlane3D( const V3& x, const V3& n )
{
  Plane3D p;
  p.x = x;
  p.n = normalized( n );
  p.t0 = rotateFromUnitY( p.n, V3{ 1.0, 0.0, 0.0 } );
  p.t1 = rotateFromUnitY( p.n, V3{ 0.0, 0.0, 1.0 } );
  return p;
}

struct Portal3D
{
  Plane3D a, b;
  int mult[3] = { 1, 1, 1 }; // PlanarPortal::m_portal_multiplier ( Array3i )
};

inline double distanceToPoint( const Plane3D& p, const V3& x ) { return dot( p.n, x - p.x ); }

// PlanarPortal.cpp:166-185
inline V3 teleportThrough3D( const Portal3D& p, const Plane3D& from, const Plane3D& to, const V3& xin )
{
  const double nA = p.mult[0] * dot( from.n, from.x - xin );
  const double tA0 = p.mult[1] * dot( from.t0, from.x - xin );
  const double tA1 = p.mult[2] * dot( from.t1, from.x - xin );
  return ( ( to.x + nA * to.n ) + tA0 * to.t0 ) + tA1 * to.t1;
}
inline V3 teleportPointThroughPlaneA( const Portal3D& p, const V3& x ) { return teleportThrough3D( p, p.a, p.b, x ); }
inline V3 teleportPointThroughPlaneB( const Portal3D& p, const V3& x ) { return teleportThrough3D( p, p.b, p.a, x ); }
inline bool pointInsidePortal( const Portal3D& p, const V3& x ) { return distanceToPoint( p.a, x ) < 0 || distanceToPoint( p.b, x ) < 0; }
inline V3 teleportPointInsidePortal( const Portal3D& p, const V3& x ) { return distanceToPoint( p.a, x ) < 0 ? teleportPointThroughPlaneA( p, x ) : teleportPointThroughPlaneB( p, x ); }

// PlanarPortal.cpp:112-124: some corner with distanceToPoint <= 0, corners in the reference's order
inline bool aabbInHalfPlane( const Box<3>& b, const Plane3D& plane )
{
  for( int ix = 0; ix < 2; ++ix ) for( int iy = 0; iy < 2; ++iy ) for( int iz = 0; iz < 2; ++iz )
  {
    if( distanceToPoint( plane, V3{ ix ? b.hi[0] : b.lo[0], iy ? b.hi[1] : b.lo[1], iz ? b.hi[2] : b.lo[2] } ) <= 0 ) { return true; }
  }
  return false;
}
// PlanarPortal.cpp:126-147 ( release build ): 0 = no, 1 = plane A, 2 = plane B
inline int aabbTouchesPortal( const Portal3D& p, const Box<3>& b )
{
  if( aabbInHalfPlane( b, p.a ) ) { return 1; }
  if( aabbInHalfPlane( b, p.b ) ) { return 2; }
  return 0;
}

// RigidBody3DSim::enforcePeriodicBoundaryConditions: portal-major, centres of mass only
inline void enforcePeriodicBoundaryConditionsRB3D( const std::vector<Portal3D>& portals, const uint32_t nb, double* q )
{
  for( const Portal3D& p : portals )
  {
    for( uint32_t b = 0; b < nb; ++b )
    {
      const V3 cm{ q[3 * b], q[3 * b + 1], q[3 * b + 2] };
      if( pointInsidePortal( p, cm ) )
      {
        const V3 o = teleportPointInsidePortal( p, cm );
        q[3 * b] = o.x; q[3 * b + 1] = o.y; q[3 * b + 2] = o.z;
      }
    }
  }
}

struct RB3DTeleportedInfo
{
  uint32_t p0, p1;
  bool pl0, pl1;
  V3 x0, x1; // teleported centres at q0 ( constructor arguments )
};

struct RB3DPortalResult
{
  std::vector<RB3DContact> active;                      // contacts of un-teleported pairs | teleported | planes | cylinders
  std::vector<std::pair<unsigned,unsigned>> candidates; // extended index space
  std::vector<TeleportedBall2D> teleported_boxes;       // ( body, portal, plane index ) of box nb + k
  std::vector<RB3DTeleportedInfo> teleported_info;
  uint64_t n_regular = 0;
  bool supported = true;
};

inline V3 teleportedCentre( const std::vector<Portal3D>& portals, const double* q, const uint32_t b, const uint32_t p, const bool plane )
{
  V3 x = loadX( q, b );
  if( p != NO_PORTAL ) { x = plane == 0 ? teleportPointThroughPlaneA( portals[p], x ) : teleportPointThroughPlaneB( portals[p], x ); }
  return x;
}

// RigidBody3DSim::computeActiveSet with portals (rigidbody3d/RigidBody3DSim.cpp:250-262 -> :1072-1260, then planes and cylinders)
inline void computeActiveSetWithPortalsRB3D( const RB3DScene& s, const std::vector<Portal3D>& portals, const double* q0, const double* q1, RB3DPortalResult& res, const bool use_grid = true )
{
  const uint32_t nb = uint32_t( s.nbodies() );
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  res = RB3DPortalResult{};
  PairSet possible_overlaps;
  std::map<unsigned,TeleportedBall2D> teleported_aabb_body_indices;
  {
    std::vector<Box<3>> aabbs( nb );
    for( uint32_t b = 0; b < nb; ++b ) { computeAABB( s, b, loadX( q1, b ), loadR( q1, nb, b ), aabbs[b] ); }
    for( uint32_t p = 0; p < uint32_t( portals.size() ); ++p )
    {
      for( uint32_t b = 0; b < nb; ++b )
      {
        const int touch = aabbTouchesPortal( portals[p], aabbs[b] );
        if( touch != 0 )
        {
          const V3 xo = touch == 1 ? teleportPointThroughPlaneA( portals[p], loadX( q1, b ) ) : teleportPointThroughPlaneB( portals[p], loadX( q1, b ) );
          Box<3> bx;
          computeAABB( s, b, xo, loadR( q1, nb, b ), bx );
          aabbs.push_back( bx );
          const TeleportedBall2D tb{ b, p, touch == 2 };
          teleported_aabb_body_indices.insert( std::make_pair( unsigned( aabbs.size() - 1 ), tb ) );
          res.teleported_boxes.push_back( tb );
        }
      }
    }
    if( !aabbs.empty() )
    {
      if( use_grid ) { getPotentialOverlaps<3>( aabbs, possible_overlaps ); }
      else { getPotentialOverlapsAllPairs<3>( aabbs, possible_overlaps ); }
    }
  }
  res.candidates.assign( possible_overlaps.begin(), possible_overlaps.end() );

  auto spheres = [&]( const unsigned a, const unsigned b ) { return s.geo( a ).type == GEO_SPHERE && s.geo( b ).type == GEO_SPHERE; };

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
      // collisionIsActive( b0, b1, q0, q1 ): the regular narrow phase into a scratch list; exits on unsupported geometry pairs
      std::vector<RB3DContact> scratch;
      if( !dispatchNarrowPhaseCollision( s, bdy0, bdy1, q0, q1, scratch ) ) { res.supported = false; return; }
      if( !scratch.empty() ) { continue; }
    }
    if( s.fixed[bdy0] && s.fixed[bdy1] ) { continue; }
    const TeleportedCollision2D tc{ bdy0, bdy1, prtl0, prtl1, plane0, plane1 };
    // teleportedCollisionHappens( q1, tc ): spheres only, everything else exits
    if( !spheres( tc.b0, tc.b1 ) ) { res.supported = false; return; }
    const V3 x0 = teleportedCentre( portals, q1, tc.b0, tc.p0, tc.pl0 ), x1 = teleportedCentre( portals, q1, tc.b1, tc.p1, tc.pl1 );
    const double rs = s.geo( tc.b0 ).r + s.geo( tc.b1 ).r;
    if( squaredNorm( x0 - x1 ) <= rs * rs ) { teleported_collisions.insert( tc ); }
  }
  res.n_regular = res.active.size();

  // generateTeleportedCollision( q0, tc ) (RigidBody3DSim.cpp:1338-1397)
  for( const TeleportedCollision2D& tc : teleported_collisions )
  {
    RB3DTeleportedInfo info;
    info.p0 = tc.p0; info.p1 = tc.p1; info.pl0 = tc.pl0; info.pl1 = tc.pl1;
    info.x0 = teleportedCentre( portals, q0, tc.b0, tc.p0, tc.pl0 );
    info.x1 = teleportedCentre( portals, q0, tc.b1, tc.p1, tc.pl1 );
    const double r0 = s.geo( tc.b0 ).r, r1 = s.geo( tc.b1 ).r;
    RB3DContact c;
    c.aux = 0; c.depth = NaN;
    if( s.fixed[tc.b0] && !s.fixed[tc.b1] )
    {
      // KinematicObjectSphereConstraint{ idx1, r1, n, idx0, x0, 0, 0 }
      c.type = KINEMATIC_OBJECT_SPHERE_TELEPORTED; c.i = tc.b1; c.j = tc.b0;
      c.n = normalized( info.x1 - info.x0 ); c.p = info.x0;
    }
    else if( !s.fixed[tc.b0] && s.fixed[tc.b1] )
    {
      c.type = KINEMATIC_OBJECT_SPHERE_TELEPORTED; c.i = tc.b0; c.j = tc.b1;
      c.n = normalized( info.x0 - info.x1 ); c.p = info.x1;
    }
    else
    {
      // TeleportedSphereSphereConstraint{ idx0, idx1, x0, x1, r0, r1 }: point q0_i + ( r0 / ( r0 + r1 ) ) * ( x1 - x0 )
      c.type = SPHERE_SPHERE_TELEPORTED; c.i = tc.b0; c.j = tc.b1;
      c.n = normalized( info.x0 - info.x1 );
      c.p = loadX( q0, tc.b0 ) + ( r0 / ( r0 + r1 ) ) * ( info.x1 - info.x0 );
    }
    res.active.emplace_back( c );
    res.teleported_info.push_back( info );
  }
  if( !computeStaticActiveSet( s, q0, q1, res.active ) ) { res.supported = false; }
}

}

#endif
