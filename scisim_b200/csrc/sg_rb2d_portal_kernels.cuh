// sg_rb2d_portal_kernels.cuh -- per-thread kernels of the rigidbody2d portal path (driver: the end of sg_rb2d.cu).  No shared
// memory, no barriers: tests/portal_kernel_harness.cpp runs these very functions on the CPU against the oracle.  The includer
// provides Rb2dDev, ContactOut2X / put2, M2d / rot2d, SG_FIXED_BIT2 and the SG_* codes; sg_portal2d.h holds the arithmetic.
//
// Reference behaviour reproduced (file:line in the SCISim checkout):
//   rigidbody2d/RigidBody2DSim.cpp:876-1040  computeBodyBodyActiveSetSpatialGridWithPortals: boxes at q1 (computeAABB, not the swept
//                                            computeCollisionAABB), a teleported box per (portal, body whose box reaches a plane),
//                                            un-teleported pairs -> dispatchNarrowPhaseCollision (the existing k_rb2d_pairs, CCD
//                                            included), the others -> collisionIsActive / TeleportedCollision set
//   rigidbody2d/RigidBody2DSim.cpp:350-411   collisionIsActive: circles only, anything with a box exits
//   rigidbody2d/RigidBody2DSim.cpp:413-636   teleported centres, dispatchTeleportedNarrowPhaseCollision (kinematic bodies exit)
//   rigidbody2d/TeleportedCircleCircleConstraint.cpp:10-29,155-158,177-182, KinematicKickCircleCircleConstraint.cpp:12-16
//   rigidbody2d/RigidBody2DSim.cpp:841-874   enforcePeriodicBoundaryConditions
#ifndef SG_RB2D_PORTAL_KERNELS_CUH
#define SG_RB2D_PORTAL_KERNELS_CUH

#include "sg_portal2d.h"
#include "sg_pair_sort.cuh"

// CircleGeometry::computeAABB (CircleGeometry.cpp:38-43) / BoxGeometry::computeAABB (BoxGeometry.cpp:32-42) at one configuration
__device__ __forceinline__ void r2p_aabb_at( const uint32_t geo, const double2 par, const double x, const double y, const double theta, double* lo, double* hi )
{
  if( geo == SG_GEO2_CIRCLE )
  {
    lo[0] = x - par.x; lo[1] = y - par.x; hi[0] = x + par.x; hi[1] = y + par.x;
  }
  else
  {
    const M2d R = rot2d( theta );
    const double ex = fabs( R.a ) * par.x + fabs( R.b ) * par.y;
    const double ey = fabs( R.c ) * par.x + fabs( R.d ) * par.y;
    lo[0] = x - ex; lo[1] = y - ey; hi[0] = x + ex; hi[1] = y + ey;
  }
}

__global__ void __launch_bounds__( 256 ) k_r2p_boxes( const Rb2dDev dev, const double* __restrict__ q1, double* __restrict__ boxes )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= dev.n ) { return; }
  const double* c = q1 + 3 * size_t( b );
  double lo[2], hi[2];
  r2p_aabb_at( __ldg( &dev.btype[b] ) & ~SG_FIXED_BIT2, __ldg( &dev.bparam[b] ), __ldg( c ), __ldg( c + 1 ), __ldg( c + 2 ), lo, hi );
  double* o = boxes + 4 * size_t( b );
  o[0] = lo[0]; o[1] = lo[1]; o[2] = hi[0]; o[3] = hi[1];
}

// grid: ( blocks over bodies, portals ); aabbTouchesPortal on the body's box at q1
__global__ void __launch_bounds__( 256 ) k_r2p_touch( const __grid_constant__ SgPortals2D ps, const uint32_t n, const double* __restrict__ boxes, uint32_t* __restrict__ tflag )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const uint32_t p = blockIdx.y;
  const double* o = boxes + 4 * size_t( b );
  const double lo[2] = { __ldg( o ), __ldg( o + 1 ) }, hi[2] = { __ldg( o + 2 ), __ldg( o + 3 ) };
  tflag[size_t( p ) * n + b] = sg_portal_aabb_touch( ps.p[p], lo, hi ) != 0 ? 1u : 0u;
}

// teleported boxes (the body's geometry at the teleported centre, same theta) and the TeleportedBody table (RigidBody2DSim.cpp:899-922)
__global__ void __launch_bounds__( 256 ) k_r2p_tele_boxes( const __grid_constant__ SgPortals2D ps, const Rb2dDev dev, const double* __restrict__ q1, const double* __restrict__ real_boxes,
                                                          const uint32_t* __restrict__ tflag, const uint32_t* __restrict__ toff, double* __restrict__ boxes, uint32_t* __restrict__ box_body,
                                                          uint32_t* __restrict__ box_portal )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= dev.n ) { return; }
  const uint32_t p = blockIdx.y;
  const size_t idx = size_t( p ) * dev.n + b;
  if( tflag[idx] == 0u ) { return; }
  const uint32_t slot = toff[idx];
  const double* rb = real_boxes + 4 * size_t( b );
  const double blo[2] = { __ldg( rb ), __ldg( rb + 1 ) }, bhi[2] = { __ldg( rb + 2 ), __ldg( rb + 3 ) };
  const int touch = sg_portal_aabb_touch( ps.p[p], blo, bhi );
  const double* c = q1 + 3 * size_t( b );
  const SgVec2 xo = sg_portal_teleport( ps.p[p], touch == 2, SgVec2{ __ldg( c ), __ldg( c + 1 ) } ); // teleportPoint( x, plane index, x_out )
  double lo[2], hi[2];
  r2p_aabb_at( __ldg( &dev.btype[b] ) & ~SG_FIXED_BIT2, __ldg( &dev.bparam[b] ), xo.x, xo.y, __ldg( c + 2 ), lo, hi );
  double* o = boxes + ( size_t( dev.n ) + slot ) * 4;
  o[0] = lo[0]; o[1] = lo[1]; o[2] = hi[0]; o[3] = hi[1];
  box_body[slot] = b;
  box_portal[slot] = p | ( touch == 2 ? SG_PLANE_B_BIT : 0u );
}

// Per candidate of the extended box set (RigidBody2DSim.cpp:932-996).  Un-teleported pairs are compacted, in order, into
// reg_pairs for the regular narrow phase; pairs with a teleported member take the TeleportedCollision tests.  bad_flag bit 0:
// a teleported candidate involves a box (collisionIsActive exits).
template<bool EMIT>
__global__ void __launch_bounds__( 128 ) k_r2p_classify( const __grid_constant__ SgPortals2D ps, const Rb2dDev dev, const uint2* __restrict__ pairs, const unsigned long long npairs,
                                                        const double* __restrict__ q1, const uint32_t* __restrict__ box_body, const uint32_t* __restrict__ box_portal,
                                                        uint32_t* __restrict__ reg_cnt, uint32_t* __restrict__ tel_cnt, const unsigned long long* __restrict__ reg_off, const uint32_t* __restrict__ tel_off,
                                                        uint2* __restrict__ reg_pairs, unsigned long long* __restrict__ tc_key, uint32_t* __restrict__ tc_idx, uint4* __restrict__ tc_info,
                                                        uint32_t* __restrict__ bad_flag )
{
  const unsigned long long k = blockIdx.x * ( unsigned long long )( blockDim.x ) + threadIdx.x;
  if( k >= npairs ) { return; }
  const uint2 pr = pairs[k];
  const uint32_t n = dev.n;
  const bool first_teleported = pr.x >= n;
  const bool second_teleported = pr.y >= n;
  if( !first_teleported && !second_teleported )
  {
    if( !EMIT ) { reg_cnt[k] = 1u; tel_cnt[k] = 0u; }
    else { reg_pairs[reg_off[k]] = pr; }
    return;
  }
  uint32_t bdy0 = pr.x, bdy1 = pr.y, prtl0 = SG_NO_PORTAL, prtl1 = SG_NO_PORTAL;
  if( first_teleported ) { bdy0 = __ldg( &box_body[pr.x - n] ); prtl0 = __ldg( &box_portal[pr.x - n] ); }
  if( second_teleported ) { bdy1 = __ldg( &box_body[pr.y - n] ); prtl1 = __ldg( &box_portal[pr.y - n] ); }
  const SgTeleCollision c = sg_tele_collision( bdy0, bdy1, prtl0, prtl1 );
  if( EMIT )
  {
    if( tel_cnt[k] != 0u )
    {
      const uint32_t o = tel_off[k];
      tc_key[o] = ( ( unsigned long long )( c.b0 ) << 32 ) | c.b1;
      tc_idx[o] = o;
      tc_info[o] = make_uint4( c.b0, c.b1, c.p0, c.p1 );
    }
    return;
  }
  reg_cnt[k] = 0u;
  const uint32_t t0 = __ldg( &dev.btype[c.b0] ), t1 = __ldg( &dev.btype[c.b1] );
  const bool circles = ( t0 & ~SG_FIXED_BIT2 ) == SG_GEO2_CIRCLE && ( t1 & ~SG_FIXED_BIT2 ) == SG_GEO2_CIRCLE;
  const double r0 = __ldg( &dev.bparam[c.b0] ).x, r1 = __ldg( &dev.bparam[c.b1] ).x;
  const SgVec2 xa{ __ldg( q1 + 3 * size_t( c.b0 ) ), __ldg( q1 + 3 * size_t( c.b0 ) + 1 ) };
  const SgVec2 xb{ __ldg( q1 + 3 * size_t( c.b1 ) ), __ldg( q1 + 3 * size_t( c.b1 ) + 1 ) };
  uint32_t happens = 0u;
  if( first_teleported && second_teleported )
  {
    // found between the un-teleported bodies as well?  collisionIsActive( ..., q1 ) comes first and exits on any box
    if( !circles ) { atomicOr( bad_flag, 1u ); tel_cnt[k] = 0u; return; }
    if( sg_ball_ball_active( xa, xb, r0, r1 ) ) { tel_cnt[k] = 0u; return; }
  }
  if( ( t0 & SG_FIXED_BIT2 ) && ( t1 & SG_FIXED_BIT2 ) ) { tel_cnt[k] = 0u; return; }
  if( !circles ) { atomicOr( bad_flag, 1u ); tel_cnt[k] = 0u; return; }
  if( sg_ball_ball_active( sg_tele_center( ps, c.p0, xa ), sg_tele_center( ps, c.p1, xb ), r0, r1 ) ) { happens = 1u; }
  tel_cnt[k] = happens;
}

// dispatchTeleportedNarrowPhaseCollision (RigidBody2DSim.cpp:477-636) for the surviving collisions (circles by construction).
// bad_flag bit 1: a kinematically scripted body takes part (the reference exits).
__global__ void __launch_bounds__( 128 ) k_r2p_tele_contacts( const __grid_constant__ SgPortals2D ps, const Rb2dDev dev, const uint32_t nraw, const uint32_t* __restrict__ idxs,
                                                             const uint32_t* __restrict__ uflag, const uint32_t* __restrict__ uoff, const uint4* __restrict__ tc_info,
                                                             const double* __restrict__ q0, const double* __restrict__ q1, const unsigned long long base, const ContactOut2X out,
                                                             double2* __restrict__ x0t, double2* __restrict__ x1t, double2* __restrict__ delta0, double2* __restrict__ delta1,
                                                             double2* __restrict__ kick, uint32_t* __restrict__ tp0, uint32_t* __restrict__ tp1, uint32_t* __restrict__ bad_flag )
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if( e >= nraw || uflag[e] == 0u ) { return; }
  const uint32_t s = uoff[e];
  const uint4 c = tc_info[idxs[e]];
  const uint32_t b0 = c.x, b1 = c.y, p0 = c.z, p1 = c.w;
  if( ( ( __ldg( &dev.btype[b0] ) | __ldg( &dev.btype[b1] ) ) & SG_FIXED_BIT2 ) != 0u ) { atomicOr( bad_flag, 2u ); return; }
  const double r0 = __ldg( &dev.bparam[b0] ).x, r1 = __ldg( &dev.bparam[b1] ).x;
  const SgVec2 q0a{ __ldg( q0 + 3 * size_t( b0 ) ), __ldg( q0 + 3 * size_t( b0 ) + 1 ) }, q0b{ __ldg( q0 + 3 * size_t( b1 ) ), __ldg( q0 + 3 * size_t( b1 ) + 1 ) };
  const SgVec2 q1a{ __ldg( q1 + 3 * size_t( b0 ) ), __ldg( q1 + 3 * size_t( b0 ) + 1 ) }, q1b{ __ldg( q1 + 3 * size_t( b1 ) ), __ldg( q1 + 3 * size_t( b1 ) + 1 ) };
  const SgVec2 x0 = sg_tele_center( ps, p0, q0a ), x1 = sg_tele_center( ps, p1, q0b );
  SgVec2 d0{ x0.x - q0a.x, x0.y - q0a.y }, d1{ x1.x - q0b.x, x1.y - q0b.y };
  const bool le0 = p0 != SG_NO_PORTAL && ps.p[p0 & ~SG_PLANE_B_BIT].v != 0.0;
  const bool le1 = p1 != SG_NO_PORTAL && ps.p[p1 & ~SG_PLANE_B_BIT].v != 0.0;
  SgVec2 kk{ 0.0, 0.0 };
  double depth;
  if( le0 || le1 )
  {
    // kick from the box of the Lees-Edwards body at q1 (collision detection ran on q1); circles: centre -+ r
    const SgVec2 xc = le1 ? q1b : q1a;
    const double rc = le1 ? r1 : r0;
    const double lo[2] = { xc.x - rc, xc.y - rc }, hi[2] = { xc.x + rc, xc.y + rc };
    const SgVec2 kv = sg_portal_kinematic_velocity_of_aabb( ps.p[( le1 ? p1 : p0 ) & ~SG_PLANE_B_BIT], lo, hi );
    if( le1 ) { kk = kv; } else { kk.x = -kv.x; kk.y = -kv.y; }
    // KinematicKickCircleCircleConstraint stores NaN displacements and radii; std::min( 0.0, NaN ) == 0.0
    const double nan = __longlong_as_double( 0x7ff8000000000000LL );
    d0.x = nan; d0.y = nan; d1.x = nan; d1.y = nan;
    depth = 0.0;
  }
  else
  {
    // TeleportedCircleCircleConstraint::computePenetrationDepth( q1 )
    const double ex = ( q1a.x + d0.x ) - ( q1b.x + d1.x ), ey = ( q1a.y + d0.y ) - ( q1b.y + d1.y );
    depth = fmin( 0.0, sqrt( ex * ex + ey * ey ) - r0 - r1 );
  }
  double nx = x0.x - x1.x, ny = x0.y - x1.y;
  const double z = nx * nx + ny * ny;
  if( z > 0.0 ) { const double sq = sqrt( z ); nx = nx / sq; ny = ny / sq; }
  // getWorldSpaceContactPoint( q0 ) = q0_i + ( r0 / ( r0 + r1 ) ) * ( x1 - x0 )
  const double w = r0 / ( r0 + r1 );
  V2d nn; nn.x = nx; nn.y = ny;
  V2d pp; pp.x = q0a.x + w * ( x1.x - x0.x ); pp.y = q0a.y + w * ( x1.y - x0.y );
  put2( out, base + s, ( le0 || le1 ) ? SG_CIRCLE_CIRCLE_KICK_TELEPORTED : SG_CIRCLE_CIRCLE_TELEPORTED, b0, b1, 0u, nn, pp, depth );
  x0t[s] = make_double2( x0.x, x0.y ); x1t[s] = make_double2( x1.x, x1.y );
  delta0[s] = make_double2( d0.x, d0.y ); delta1[s] = make_double2( d1.x, d1.y );
  kick[s] = make_double2( kk.x, kk.y );
  tp0[s] = p0; tp1[s] = p1;
}

// RigidBody2DSim::enforcePeriodicBoundaryConditions on (q, v) in place, [x, y, theta] per body
__global__ void __launch_bounds__( 256 ) k_r2p_enforce( const __grid_constant__ SgPortals2D ps, const uint32_t n, double* __restrict__ q, double* __restrict__ v )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  SgVec2 xs{ q[3 * size_t( b )], q[3 * size_t( b ) + 1] }, vs{ v[3 * size_t( b )], v[3 * size_t( b ) + 1] };
  sg_portals_enforce( ps, xs, vs );
  q[3 * size_t( b )] = xs.x; q[3 * size_t( b ) + 1] = xs.y;
  v[3 * size_t( b )] = vs.x; v[3 * size_t( b ) + 1] = vs.y;
}

#endif
