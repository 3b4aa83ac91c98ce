// sg_portal2d.h -- planar / Lees-Edwards portal arithmetic of ball2d, shared by the kernels and the host code of
// libscisim_b200 (plain C++ when compiled without nvcc: tests/portal_math_harness.cpp runs the very same functions
// on the CPU against the oracle).
//
// Reference behaviour reproduced (file:line in the SCISim checkout):
//   ball2d/StaticGeometry/StaticPlane.cpp:10-14   n normalised, t = ( -n.y, n.x )
//   ball2d/StaticGeometry/StaticPlane.cpp:66-74   distanceLessThanZero, distanceLessThanOrEqualZero
//   ball2d/Portals/PlanarPortal.cpp:107-133       pointInsidePortal, ballTouchesPortal
//   ball2d/Portals/PlanarPortal.cpp:135-165       teleportPointInsidePortal, teleportBall
//   ball2d/Portals/PlanarPortal.cpp:167-193       getKinematicVelocityOfBall / OfPoint
//   ball2d/Portals/PlanarPortal.cpp:195-229       teleportPointThroughPlaneA / B
//   ball2d/Portals/PlanarPortal.cpp:231-236       updateMovingPortals
//   rigidbody2d/PlanarPortal.cpp:143-246,298-350  the rigid-body sim's PlanarPortal: identical teleports / offsets, AABB touch tests
//   rigidbody2d/RigidBody2DStaticPlane.cpp:10-14  its plane: normal used as given
// FP64 throughout, evaluation order as written there (Eigen evaluates these 2-vector expressions coefficient by
// coefficient, dot( a, b ) = a0*b0 + a1*b1); the library is compiled without FMA contraction.
#ifndef SG_PORTAL2D_H
#define SG_PORTAL2D_H

#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define SG_HD __host__ __device__ __forceinline__
#else
#define SG_HD inline
#endif

#define SG_MAX_PORTALS 8
#define SG_NO_PORTAL 0xffffffffu
#define SG_PLANE_B_BIT 0x80000000u // or-ed into a portal index: the body went through plane B (plane index 1)

struct SgVec2 { double x, y; };

// One PlanarPortal: two StaticPlanes (point, unit normal, tangent), the Lees-Edwards velocity, the bounds of the
// periodic tangential coordinate and the current tangential offset
struct SgPortal2D
{
  double ax[2], an[2], at[2];
  double bx[2], bn[2], bt[2];
  double v, bounds, dx;
};

struct SgPortals2D
{
  uint32_t n;
  uint32_t pad;
  SgPortal2D p[SG_MAX_PORTALS];
};

// int( floor( x ) ) as the reference's x86-64 build evaluates it (cvttsd2si): INT_MIN for NaN and out-of-range values.
// Plain portals have bounds == 0: the reference divides by zero and then multiplies the integer by 0.
SG_HD int sg_to_int_x86( const double x )
{
  return ( x >= -2147483648.0 && x < 2147483648.0 ) ? int( x ) : int( -2147483647 - 1 );
}

// n.dot( x - x_plane )
SG_HD double sg_plane_dist( const double* px, const double* pn, const SgVec2 x )
{
  return pn[0] * ( x.x - px[0] ) + pn[1] * ( x.y - px[1] );
}

// PlanarPortal::teleportPointThroughPlaneA (through_b == false) / ...B (true)
SG_HD SgVec2 sg_portal_teleport( const SgPortal2D& p, const bool through_b, const SgVec2 xin )
{
  // scalar selects rather than pointer selects: the portal usually lives in kernel-parameter (constant) space
  const double fx0 = through_b ? p.bx[0] : p.ax[0], fx1 = through_b ? p.bx[1] : p.ax[1];
  const double fn0 = through_b ? p.bn[0] : p.an[0], fn1 = through_b ? p.bn[1] : p.an[1];
  const double ft0 = through_b ? p.bt[0] : p.at[0], ft1 = through_b ? p.bt[1] : p.at[1];
  const double tx0 = through_b ? p.ax[0] : p.bx[0], tx1 = through_b ? p.ax[1] : p.bx[1];
  const double tn0 = through_b ? p.an[0] : p.bn[0], tn1 = through_b ? p.an[1] : p.bn[1];
  const double tt0 = through_b ? p.at[0] : p.bt[0], tt1 = through_b ? p.at[1] : p.bt[1];
  const double nA = fn0 * ( fx0 - xin.x ) + fn1 * ( fx1 - xin.y );
  double tA = ft0 * ( ( p.dx * ft0 + fx0 ) - xin.x ) + ft1 * ( ( p.dx * ft1 + fx1 ) - xin.y );
  const int repeat_x = sg_to_int_x86( floor( ( tA + p.bounds ) / ( 2.0 * p.bounds ) ) );
  tA = tA - ( 2.0 * double( repeat_x ) ) * p.bounds;
  SgVec2 out;
  out.x = ( tx0 + nA * tn0 ) + tA * tt0;
  out.y = ( tx1 + nA * tn1 ) + tA * tt1;
  return out;
}

// PlanarPortal::ballTouchesPortal: 0 = no, 1 = plane A, 2 = plane B, 3 = both (the reference prints and exits)
SG_HD int sg_portal_touch( const SgPortal2D& p, const SgVec2 x, const double r )
{
  const bool ta = sg_plane_dist( p.ax, p.an, x ) <= r;
  const bool tb = sg_plane_dist( p.bx, p.bn, x ) <= r;
  return ( ta ? 1 : 0 ) | ( tb ? 2 : 0 );
}

// PlanarPortal::teleportBall: through A when the ball touches A, through B otherwise
SG_HD SgVec2 sg_portal_teleport_ball( const SgPortal2D& p, const SgVec2 x, const double r )
{
  return sg_portal_teleport( p, !( sg_plane_dist( p.ax, p.an, x ) <= r ), x );
}

SG_HD bool sg_portal_point_inside( const SgPortal2D& p, const SgVec2 x )
{
  return sg_plane_dist( p.ax, p.an, x ) < 0.0 || sg_plane_dist( p.bx, p.bn, x ) < 0.0;
}

// PlanarPortal::teleportPointInsidePortal
SG_HD SgVec2 sg_portal_teleport_point_inside( const SgPortal2D& p, const SgVec2 x )
{
  return sg_portal_teleport( p, !( sg_plane_dist( p.ax, p.an, x ) < 0.0 ), x );
}

// PlanarPortal::getKinematicVelocityOfBall: -v * t of the plane the ball touches (A first)
SG_HD SgVec2 sg_portal_kinematic_velocity_of_ball( const SgPortal2D& p, const SgVec2 x, const double r )
{
  const bool a = sg_plane_dist( p.ax, p.an, x ) <= r;
  SgVec2 out;
  out.x = ( -p.v ) * ( a ? p.at[0] : p.bt[0] ); out.y = ( -p.v ) * ( a ? p.at[1] : p.bt[1] );
  return out;
}

// PlanarPortal::getKinematicVelocityOfPoint
SG_HD SgVec2 sg_portal_kinematic_velocity_of_point( const SgPortal2D& p, const SgVec2 x )
{
  const bool a = sg_plane_dist( p.ax, p.an, x ) < 0.0;
  SgVec2 out;
  out.x = ( -p.v ) * ( a ? p.at[0] : p.bt[0] ); out.y = ( -p.v ) * ( a ? p.at[1] : p.bt[1] );
  return out;
}

// PlanarPortal::updateMovingPortals( t ): dx = v t wrapped into [-bounds, bounds]
SG_HD double sg_portal_offset( const double v, const double bounds, const double t )
{
  const int repeat_x = sg_to_int_x86( floor( ( v * t + bounds ) / ( 2.0 * bounds ) ) );
  return v * t - ( 2.0 * double( repeat_x ) ) * bounds;
}

// StaticPlane's constructor: normal normalised ( n / sqrt( n.n ) when n.n > 0 ), tangent ( -n.y, n.x )
SG_HD void sg_portal_plane_frame( const double* n_in, double* n_out, double* t_out )
{
  double nx = n_in[0], ny = n_in[1];
  const double z = nx * nx + ny * ny;
  if( z > 0.0 ) { const double s = sqrt( z ); nx = nx / s; ny = ny / s; }
  n_out[0] = nx; n_out[1] = ny;
  t_out[0] = -ny; t_out[1] = nx;
}

// ---- rigidbody2d's PlanarPortal (rigidbody2d/PlanarPortal.cpp): same teleports and offsets, AABB-based touch tests -------
// RigidBody2DStaticPlane's constructor (rigidbody2d/RigidBody2DStaticPlane.cpp:10-14): the normal is used as given
SG_HD void sg_portal_plane_frame_as_given( const double* n_in, double* n_out, double* t_out )
{
  n_out[0] = n_in[0]; n_out[1] = n_in[1];
  t_out[0] = -n_in[1]; t_out[1] = n_in[0];
}

// aabbInHalfPlane (PlanarPortal.cpp:143-165): some corner has distanceToPoint <= 0; corners in the reference's order
SG_HD bool sg_aabb_in_half_plane( const double* px, const double* pn, const double* lo, const double* hi )
{
  if( sg_plane_dist( px, pn, SgVec2{ lo[0], lo[1] } ) <= 0.0 ) { return true; }
  if( sg_plane_dist( px, pn, SgVec2{ lo[0], hi[1] } ) <= 0.0 ) { return true; }
  if( sg_plane_dist( px, pn, SgVec2{ hi[0], lo[1] } ) <= 0.0 ) { return true; }
  if( sg_plane_dist( px, pn, SgVec2{ hi[0], hi[1] } ) <= 0.0 ) { return true; }
  return false;
}

// PlanarPortal::aabbTouchesPortal, release build (PlanarPortal.cpp:167-189): 0 = no, 1 = plane A, 2 = plane B; A wins
SG_HD int sg_portal_aabb_touch( const SgPortal2D& p, const double* lo, const double* hi )
{
  if( sg_aabb_in_half_plane( p.ax, p.an, lo, hi ) ) { return 1; }
  if( sg_aabb_in_half_plane( p.bx, p.bn, lo, hi ) ) { return 2; }
  return 0;
}

// PlanarPortal::getKinematicVelocityOfAABB (PlanarPortal.cpp:218-230)
SG_HD SgVec2 sg_portal_kinematic_velocity_of_aabb( const SgPortal2D& p, const double* lo, const double* hi )
{
  const bool a = sg_aabb_in_half_plane( p.ax, p.an, lo, hi );
  SgVec2 out;
  out.x = ( -p.v ) * ( a ? p.at[0] : p.bt[0] ); out.y = ( -p.v ) * ( a ? p.at[1] : p.bt[1] );
  return out;
}

// Enforces every portal on one body, portal-major like Ball2DSim::enforcePeriodicBoundaryConditions
// (ball2d/Ball2DSim.cpp:336-366): a body inside a portal is teleported, and picks up the plane's velocity when the
// portal is a Lees-Edwards one.
SG_HD void sg_portals_enforce( const SgPortals2D& ps, SgVec2& x, SgVec2& v )
{
  for( uint32_t k = 0; k < ps.n; ++k )
  {
    const SgPortal2D& p = ps.p[k];
    if( sg_portal_point_inside( p, x ) )
    {
      const SgVec2 xin = x;
      x = sg_portal_teleport_point_inside( p, xin );
      if( p.v != 0.0 )
      {
        const SgVec2 dv = sg_portal_kinematic_velocity_of_point( p, xin );
        v.x = v.x + dv.x; v.y = v.y + dv.y;
      }
    }
  }
}

// ---- teleported collisions ------------------------------------------------------------------------------------------
// TeleportedCollision (PlanarPortal.cpp:34-57): bodies ordered, portal words (index | SG_PLANE_B_BIT, or SG_NO_PORTAL) follow
struct SgTeleCollision { uint32_t b0, b1, p0, p1; };

SG_HD SgTeleCollision sg_tele_collision( uint32_t b0, uint32_t b1, uint32_t p0, uint32_t p1 )
{
  SgTeleCollision c;
  if( b0 > b1 ) { c.b0 = b1; c.b1 = b0; c.p0 = p1; c.p1 = p0; }
  else { c.b0 = b0; c.b1 = b1; c.p0 = p0; c.p1 = p1; }
  return c;
}

// Ball2DSim::getTeleportedBallBallCenters for one body (ball2d/Ball2DSim.cpp:610-642)
SG_HD SgVec2 sg_tele_center( const SgPortals2D& ps, const uint32_t portal_word, const SgVec2 x )
{
  if( portal_word == SG_NO_PORTAL ) { return x; }
  return sg_portal_teleport( ps.p[portal_word & ~SG_PLANE_B_BIT], ( portal_word & SG_PLANE_B_BIT ) != 0u, x );
}

// BallBallConstraint::isActive (ball2d/Constraints/BallBallConstraint.cpp:16-20)
SG_HD bool sg_ball_ball_active( const SgVec2 x0, const SgVec2 x1, const double r0, const double r1 )
{
  const double dx = x0.x - x1.x, dy = x0.y - x1.y;
  return dx * dx + dy * dy <= ( r0 + r1 ) * ( r0 + r1 );
}

// ---- bitonic network over (key, insertion index) pairs ------------------------------------------------------------------
// Element e of a compare-exchange step (k = size of the sorted runs being merged, j = partner distance): returns the
// partner and whether e's pair is sorted ascending.  Standard network: partner = e ^ j, ascending iff ( e & k ) == 0.
SG_HD uint32_t sg_bitonic_partner( const uint32_t e, const uint32_t j ) { return e ^ j; }
SG_HD bool sg_bitonic_ascending( const uint32_t e, const uint32_t k ) { return ( e & k ) == 0u; }
// strict order on (key, index): keys tie only between copies of one body pair, the insertion index breaks the tie
SG_HD bool sg_tele_less( const unsigned long long ka, const uint32_t ia, const unsigned long long kb, const uint32_t ib )
{
  return ka < kb || ( ka == kb && ia < ib );
}

#endif
