// oracle/rb2d.h
// Checked bit for bit against the reference's own source compiled unchanged (oracle/Makefile.ref, tests/test_oracle_vs_reference.py).
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). CPU restatement of the rigidbody2d hot path:
//   rigidbody2d/SymplecticEulerMap.cpp:15-38, VerletMap.cpp:15-55        flow (forces zeroed on kinematic bodies)
//   rigidbody2d/NearEarthGravityForce.cpp:35-48, RigidBody2DSim.cpp:104-113, RigidBody2DState.cpp:31-43 (Minv = 1/m)
//   rigidbody2d/CircleGeometry.cpp:32-37 (swept AABB), BoxGeometry.cpp:32-42 (AABB at q1, |Rot(theta1)| r)
//   rigidbody2d/RigidBody2DSim.cpp:696-714   computeActiveSet (no portals: those are in rb2d_portals.h)
//   rigidbody2d/RigidBody2DSim.cpp:1066-1100 computeBodyBodyActiveSetSpatialGrid
//   rigidbody2d/RigidBody2DSim.cpp:248-348   dispatchNarrowPhaseCollision (kinematic rules, type switch)
//   rigidbody2d/RigidBody2DSim.cpp:184-246   boxBox / boxCircle narrow phase callers
//   rigidbody2d/BoxBoxTools.cpp:50-196       2-D SAT + 1-D clip, contacts offset by half the depth
//   rigidbody2d/CircleBoxTools.cpp:8-114     Voronoi-region circle vs box
//   rigidbody2d/RigidBody2DSim.cpp:638-694   computeBodyPlaneActiveSetAllPairs
//   rigidbody2d/{CircleCircle,StaticPlaneCircle}Constraint.cpp  isActive / depth
// Rotation2D(theta).matrix() = [[c,-s],[s,c]] with libm sin/cos (device: CUDA sincos => rotated boxes compare to 1e-12).
// Parity: stored reference outputs exist for the CCD cases only; beyond them the file is pinned against the reference's sources compiled
// unchanged (oracle/_ref): both maps, the grid, BoxBoxTools, CircleBoxTools, the geometry classes' AABBs, the plane / portal classes and the constraint
// classes CircleCircle / StaticPlaneCircle / StaticPlaneBody / BodyBody (normal, contact point, depth for every contact of its active sets); and, for the
// glue, against the reference's own RigidBody2DSim (RigidBody2DSim.cpp + RigidBody2DState.cpp compiled unchanged, oracle/ref_shims/ref_rb2d_sim.cpp):
// computeActiveSet and flow as a whole, portals included, equal this file and rb2d_portals.h bit for bit (tests/test_reference_sim_cpu.py).
#ifndef ORACLE_RB2D_H
#define ORACLE_RB2D_H

#include "broadphase.h"
#include "ccd.h"

#include <limits>
#include <vector>

namespace orc
{

enum RB2DGeoType : uint32_t { GEO2_CIRCLE = 0, GEO2_BOX = 1 };
enum RB2DContactType : uint32_t { CIRCLE_CIRCLE = 20, KINEMATIC_CIRCLE = 21, BODY_BODY_2D = 22, PLANE_CIRCLE = 23, PLANE_BODY_2D = 24 };

struct M2 { double a, b, c, d; }; // [[a,b],[c,d]]
inline M2 rot2( const double theta ) { const double s = std::sin( theta ), c = std::cos( theta ); return M2{ c, -s, s, c }; }
inline V2 mul( const M2& R, const V2& v ) { return V2{ R.a * v.x + R.b * v.y, R.c * v.x + R.d * v.y }; }
inline V2 mulT( const M2& R, const V2& v ) { return V2{ R.a * v.x + R.c * v.y, R.b * v.x + R.d * v.y }; }
inline V2 col( const M2& R, const int j ) { return j == 0 ? V2{ R.a, R.c } : V2{ R.b, R.d }; }
inline double at( const V2& v, const int i ) { return i == 0 ? v.x : v.y; }

struct RB2DGeometry { uint32_t type; double r; V2 half; };

struct RB2DContact
{
  uint32_t type, i, j, aux;
  V2 n;
  V2 p;         // circle-circle / body-body: contact point; kinematic circle: kinematic body's position at q0;
                // plane-circle: x0 - r n; plane-body: body-space arm of the corner
  double depth; // penetrationDepth( q1 ) where overridden, else NaN
};

struct RB2DScene
{
  std::vector<RB2DGeometry> geometry;
  std::vector<uint32_t> geo_of_body;
  std::vector<uint8_t> fixed;
  std::vector<double> M;      // 3N diagonal: m, m, I
  V2 g{ 0.0, 0.0 };
  std::vector<V2> plane_x, plane_n;   // used as given (RigidBody2DStaticPlane does not normalise)
  std::size_t nbodies() const { return geo_of_body.size(); }
  const RB2DGeometry& geo( const std::size_t b ) const { return geometry[geo_of_body[b]]; }
};

// kind: 0 symplectic_euler, 1 verlet
inline void flow( const int kind, const RB2DScene& s, const double* q0, const double* v0, const double dt, double* q1, double* v1 )
{
  const std::size_t nb = s.nbodies();
  for( std::size_t b = 0; b < nb; ++b )
  {
    for( int k = 0; k < 3; ++k )
    {
      const std::size_t d = 3 * b + k;
      const double minv = 1.0 / s.M[d];
      // F.setZero(); F_xy += m * g; then zeroed for kinematic bodies
      double F = ( k < 2 ) ? 0.0 + s.M[3 * b] * ( k == 0 ? s.g.x : s.g.y ) : 0.0;
      if( s.fixed[b] ) { F = 0.0; }
      if( kind == 0 )
      {
        v1[d] = v0[d] + ( 0.0 + ( dt * minv ) * F );
        q1[d] = q0[d] + dt * v1[d];
      }
      else
      {
        const double sc = ( 0.5 * dt ) * minv;
        const double vh = v0[d] + ( 0.0 + sc * F );
        q1[d] = q0[d] + dt * vh;
        v1[d] = vh + sc * F;
      }
    }
  }
}

inline void computeCollisionAABB( const RB2DGeometry& g, const double* q0b, const double* q1b, Box<2>& box )
{
  if( g.type == GEO2_CIRCLE )
  {
    for( int k = 0; k < 2; ++k )
    {
      box.lo[k] = std::min( q0b[k], q1b[k] ) - g.r;
      box.hi[k] = std::max( q0b[k], q1b[k] ) + g.r;
    }
  }
  else
  {
    const M2 R = rot2( q1b[2] );
    const V2 e{ std::fabs( R.a ) * g.half.x + std::fabs( R.b ) * g.half.y, std::fabs( R.c ) * g.half.x + std::fabs( R.d ) * g.half.y };
    box.lo[0] = q1b[0] - e.x; box.lo[1] = q1b[1] - e.y;
    box.hi[0] = q1b[0] + e.x; box.hi[1] = q1b[1] + e.y;
  }
}

namespace boxbox2d
{
inline bool axisTest( const double dist, const double widths, const int crnt, double& smallest, bool& invert, int& feature )
{
  const double pen = std::fabs( dist ) - widths;
  if( pen > 0 ) { return true; }
  if( pen > smallest ) { smallest = pen; invert = dist < 0.0; feature = crnt; }
  return false;
}

// BoxBoxTools::isActive
inline void isActive( const V2& x0, const double theta0, const V2& r0, const V2& x1, const double theta1, const V2& r1, V2& n, std::vector<V2>& points )
{
  const M2 R0 = rot2( theta0 ), R1 = rot2( theta1 );
  int feature = 4;
  bool invert = false;
  {
    double min_pen = -std::numeric_limits<double>::infinity();
    // Q = ( R0^T R1 ).cwiseAbs()
    const M2 Q{ std::fabs( R0.a * R1.a + R0.c * R1.c ), std::fabs( R0.a * R1.b + R0.c * R1.d ), std::fabs( R0.b * R1.a + R0.d * R1.c ), std::fabs( R0.b * R1.b + R0.d * R1.d ) };
    const V2 p = x1 - x0;
    {
      const V2 pR0 = mulT( R0, p );
      const V2 w = mul( Q, r1 ) + r0;
      if( axisTest( pR0.x, w.x, 0, min_pen, invert, feature ) ) { return; }
      if( axisTest( pR0.y, w.y, 1, min_pen, invert, feature ) ) { return; }
    }
    {
      const V2 pR1 = mulT( R1, p );
      const V2 w = mulT( Q, r0 ) + r1;
      if( axisTest( pR1.x, w.x, 2, min_pen, invert, feature ) ) { return; }
      if( axisTest( pR1.y, w.y, 3, min_pen, invert, feature ) ) { return; }
    }
  }
  const bool first = feature <= 1;
  n = first ? col( R0, feature ) : col( R1, feature - 2 );
  if( invert ) { n = V2{ n.x * -1.0, n.y * -1.0 }; }
  const M2 Ra = first ? R0 : R1, Rb = first ? R1 : R0;
  const V2 xa = first ? x0 : x1, xb = first ? x1 : x0;
  const V2 ra = first ? r0 : r1, rb = first ? r1 : r0;
  const V2 normal2 = first ? n : V2{ -n.x, -n.y };
  const V2 n_in_b = mulT( Rb, normal2 );
  const int b_nrml = std::fabs( n_in_b.y ) > std::fabs( n_in_b.x ) ? 1 : 0;
  const int b_tngt = 1 - b_nrml;
  // xb - xa + ( sign * rb(b_nrml) ) * Rb.col(b_nrml): scalar product evaluated left to right, then the vector
  const double sc = ( at( n_in_b, b_nrml ) < 0.0 ? 1.0 : -1.0 ) * at( rb, b_nrml );
  const V2 b_face_center = ( xb - xa ) + sc * col( Rb, b_nrml );
  const int a_nrml = first ? feature : feature - 2;
  const int a_tngt = 1 - a_nrml;
  const double c_on_a = dot( b_face_center, col( Ra, a_tngt ) );
  const double costheta = dot( col( Ra, a_tngt ), col( Rb, b_tngt ) );
  double e0 = c_on_a - costheta * at( rb, b_tngt ), e1 = c_on_a + costheta * at( rb, b_tngt );
  if( e0 > e1 ) { std::swap( e0, e1 ); }
  e0 = std::min( e0, at( ra, a_tngt ) );
  e1 = std::max( e1, -at( ra, a_tngt ) );
  const double isect[2] = { std::max( -at( ra, a_tngt ), e0 ), std::min( at( ra, a_tngt ), e1 ) };
  const int num = isect[0] != isect[1] ? 2 : 1;
  for( int c = 0; c < num; ++c )
  {
    const V2 point = b_face_center + ( ( isect[c] - c_on_a ) / costheta ) * col( Rb, b_tngt );
    const double depth = at( ra, a_nrml ) - dot( normal2, point );
    if( depth >= 0.0 ) { points.emplace_back( ( xa + point ) + ( 0.5 * depth ) * normal2 ); }
  }
  n = V2{ n.x * -1.0, n.y * -1.0 };
}
}

// CircleBoxTools::isActive
inline bool circleBoxActive( const V2& x0, const double r0, const V2& x1, const double theta1, const V2& r1, V2& n, V2& p )
{
  const M2 R = rot2( theta1 );
  V2 xc = mulT( R, x0 - x1 );
  const bool invert_x = xc.x < 0.0, invert_y = xc.y < 0.0;
  if( invert_x ) { xc.x *= -1.0; }
  if( invert_y ) { xc.y *= -1.0; }
  double pen;
  if( r1.x * xc.y < r1.y * xc.x )
  {
    if( xc.y <= r1.y )
    {
      pen = xc.x - r0 - r1.x;
      if( pen > 0.0 ) { return false; }
      n = V2{ 1.0, 0.0 };
    }
    else
    {
      n = xc - r1;
      pen = squaredNorm( n );
      if( pen > r0 * r0 ) { return false; }
      pen = std::sqrt( pen );
      n = V2{ n.x / pen, n.y / pen };
      pen -= r0;
    }
  }
  else
  {
    if( xc.x <= r1.x )
    {
      pen = xc.y - r0 - r1.y;
      if( pen > 0.0 ) { return false; }
      n = V2{ 0.0, 1.0 };
    }
    else
    {
      n = xc - r1;
      pen = squaredNorm( n );
      if( pen > r0 * r0 ) { return false; }
      pen = std::sqrt( pen );
      n = V2{ n.x / pen, n.y / pen };
      pen -= r0;
    }
  }
  p = xc - ( r0 + 0.5 * pen ) * n;
  if( invert_x ) { n.x *= -1.0; p.x *= -1.0; }
  if( invert_y ) { n.y *= -1.0; p.y *= -1.0; }
  n = mul( R, n );
  p = mul( R, p ) + x1;
  return true;
}

// RigidBody2DSim::dispatchNarrowPhaseCollision for one candidate pair (rigidbody2d/RigidBody2DSim.cpp:248-348); appends to
// active_set; returns false where the reference exits (kinematic box-box, kinematic circle vs box)
inline bool dispatchNarrowPhaseCollision( const RB2DScene& s, const unsigned first, const unsigned second, const double* q0, const double* q1, std::vector<RB2DContact>& active_set )
{
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  auto X = [&]( const double* q, unsigned b ) { return V2{ q[3 * b], q[3 * b + 1] }; };
  {
    {
      unsigned i0 = first, i1 = second;
      if( s.fixed[i0] && s.fixed[i1] ) { return true; }
      if( s.fixed[i0] ) { std::swap( i0, i1 ); }
      const RB2DGeometry& g0 = s.geo( i0 );
      const RB2DGeometry& g1 = s.geo( i1 );
      if( g0.type == GEO2_CIRCLE && g1.type == GEO2_CIRCLE )
      {
        const V2 q0a = X( q0, i0 ), q1a = X( q1, i0 ), q0b = X( q0, i1 ), q1b = X( q1, i1 );
        if( ballBallCCDCollisionHappens( q0a, q1a, g0.r, q0b, q1b, g1.r ).first )
        {
          RB2DContact c;
          c.i = i0; c.j = i1; c.aux = 0;
          c.n = normalized( q0a - q0b );
          if( !s.fixed[i1] )
          {
            c.type = CIRCLE_CIRCLE;
            c.p = q0a + ( g0.r / ( g0.r + g1.r ) ) * ( q0b - q0a );
            c.depth = std::min( 0.0, norm( q1a - q1b ) - g0.r - g1.r );
          }
          else
          {
            c.type = KINEMATIC_CIRCLE;
            c.p = q0b;
            c.depth = NaN;
          }
          active_set.emplace_back( c );
        }
      }
      else if( g0.type == GEO2_BOX && g1.type == GEO2_BOX )
      {
        if( s.fixed[i0] || s.fixed[i1] ) { return false; }
        V2 n{ 0.0, 0.0 };
        std::vector<V2> points;
        boxbox2d::isActive( X( q1, i0 ), q1[3 * i0 + 2], g0.half, X( q1, i1 ), q1[3 * i1 + 2], g1.half, n, points );
        for( const V2& pt : points )
        {
          RB2DContact c;
          c.type = BODY_BODY_2D; c.i = i0; c.j = i1; c.aux = 0; c.n = n; c.p = pt; c.depth = NaN;
          active_set.emplace_back( c );
        }
      }
      else
      {
        // circle vs box (either order after the kinematic swap)
        const unsigned ic = g0.type == GEO2_CIRCLE ? i0 : i1;
        const unsigned ib = g0.type == GEO2_CIRCLE ? i1 : i0;
        if( s.fixed[ic] ) { return false; }
        V2 n, p;
        if( circleBoxActive( X( q1, ic ), s.geo( ic ).r, X( q1, ib ), q1[3 * ib + 2], s.geo( ib ).half, n, p ) )
        {
          RB2DContact c;
          c.aux = 0; c.depth = NaN;
          if( !s.fixed[ib] )
          {
            c.type = BODY_BODY_2D;
            if( ic < ib ) { c.i = ic; c.j = ib; c.p = p; c.n = n; }
            else { c.i = ib; c.j = ic; c.p = p; c.n = V2{ -n.x, -n.y }; }
          }
          else
          {
            c.type = KINEMATIC_CIRCLE; c.i = ic; c.j = ib; c.n = n; c.p = X( q0, ib );
          }
          active_set.emplace_back( c );
        }
      }
    }
  }
  return true;
}

// RigidBody2DSim::computeBodyPlaneActiveSetAllPairs (rigidbody2d/RigidBody2DSim.cpp:638-694); appends
inline void computeBodyPlaneActiveSet( const RB2DScene& s, const double* q0, const double* q1, std::vector<RB2DContact>& active_set )
{
  const std::size_t nb = s.nbodies();
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  auto X = [&]( const double* q, unsigned b ) { return V2{ q[3 * b], q[3 * b + 1] }; };
  for( uint32_t pl = 0; pl < uint32_t( s.plane_x.size() ); ++pl )
  {
    const V2 xp = s.plane_x[pl], np = s.plane_n[pl];
    for( uint32_t b = 0; b < uint32_t( nb ); ++b )
    {
      if( s.fixed[b] ) { continue; }
      const RB2DGeometry& g = s.geo( b );
      const V2 x1 = X( q1, b );
      if( g.type == GEO2_CIRCLE )
      {
        const double d = dot( np, x1 - xp );
        if( d <= g.r )
        {
          RB2DContact c;
          c.type = PLANE_CIRCLE; c.i = b; c.j = pl; c.aux = 0; c.n = np;
          c.p = X( q0, b ) - g.r * np;
          c.depth = std::min( 0.0, d - g.r );
          active_set.emplace_back( c );
        }
      }
      else
      {
        const M2 R = rot2( q1[3 * b + 2] );
        int corner = 0;
        for( int i = -1; i < 2; i += 2 )
        {
          for( int j = -1; j < 2; j += 2 )
          {
            const V2 arm{ double( i ) * g.half.x, double( j ) * g.half.y };
            const V2 tv = x1 + mul( R, arm );
            if( dot( np, tv - xp ) <= 0.0 )
            {
              RB2DContact c;
              c.type = PLANE_BODY_2D; c.i = b; c.j = pl; c.aux = uint32_t( corner ); c.n = np; c.p = arm; c.depth = NaN;
              active_set.emplace_back( c );
            }
            ++corner;
          }
        }
      }
    }
  }
}

// returns false where the reference exits (kinematic box-box, kinematic circle vs box)
inline bool computeActiveSet( const RB2DScene& s, const double* q0, const double* q1, std::vector<RB2DContact>& active_set,
                              std::vector<std::pair<unsigned,unsigned>>* candidates_out = nullptr, const bool use_grid = true )
{
  const std::size_t nb = s.nbodies();
  active_set.clear();
  if( nb > 0 )
  {
    PairSet possible_overlaps;
    {
      std::vector<Box<2>> aabbs( nb );
      for( std::size_t b = 0; b < nb; ++b ) { computeCollisionAABB( s.geo( b ), q0 + 3 * b, q1 + 3 * b, aabbs[b] ); }
      if( use_grid ) { getPotentialOverlaps<2>( aabbs, possible_overlaps ); }
      else { getPotentialOverlapsAllPairs<2>( aabbs, possible_overlaps ); }
    }
    if( candidates_out != nullptr ) { candidates_out->assign( possible_overlaps.begin(), possible_overlaps.end() ); }
    for( const auto& pr : possible_overlaps )
    {
      if( !dispatchNarrowPhaseCollision( s, pr.first, pr.second, q0, q1, active_set ) ) { return false; }
    }
  }
  computeBodyPlaneActiveSet( s, q0, q1, active_set );
  return true;
}

}

#endif
