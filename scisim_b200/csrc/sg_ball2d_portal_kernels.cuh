// sg_ball2d_portal_kernels.cuh -- the kernels of the ball2d portal path (driver and description: sg_ball2d_portals.cuh).
// All but one are plain per-thread kernels without shared memory or barriers; k_b2p_bitonic_tile uses one shared array
// pair and __syncthreads only.  tests/portal_kernel_harness.cpp therefore runs these very functions on the CPU (a launch
// = a loop over blocks; the threads of a block are a loop or, for the tile sort, host threads meeting at a barrier) against
// the oracle; this file depends only on sg_portal2d.h, ContactOut2D and the SG_* contact codes, which the includer provides.
#ifndef SG_BALL2D_PORTAL_KERNELS_CUH
#define SG_BALL2D_PORTAL_KERNELS_CUH

#include "sg_portal2d.h"
#include "sg_pair_sort.cuh"

// ---- kernels ---------------------------------------------------------------------------------------
// grid: ( blocks over balls, portals )
__global__ void __launch_bounds__( 256 ) k_b2p_touch( const __grid_constant__ SgPortals2D ps, const uint32_t n, const double2* __restrict__ q1, const double* __restrict__ r, uint32_t* __restrict__ tflag,
                                                     uint32_t* __restrict__ err )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const uint32_t p = blockIdx.y;
  const double2 x = __ldg( &q1[b] );
  const int touch = sg_portal_touch( ps.p[p], SgVec2{ x.x, x.y }, __ldg( &r[b] ) );
  if( touch == 3 ) { atomicOr( err, 1u ); }
  tflag[size_t( p ) * n + b] = ( touch == 1 || touch == 2 ) ? 1u : 0u;
}

// aabbs.emplace_back( q1 - r, q1 + r ) (ball2d/Ball2DSim.cpp:385)
__global__ void __launch_bounds__( 256 ) k_b2p_boxes( const uint32_t n, const double2* __restrict__ q1, const double* __restrict__ r, double* __restrict__ boxes )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const double2 x = __ldg( &q1[b] );
  const double rad = __ldg( &r[b] );
  double2* o = reinterpret_cast<double2*>( boxes + size_t( b ) * 4 );
  o[0] = make_double2( x.x - rad, x.y - rad );
  o[1] = make_double2( x.x + rad, x.y + rad );
}

// teleported boxes and the TeleportedBall table (ball2d/Ball2DSim.cpp:393-412)
__global__ void __launch_bounds__( 256 ) k_b2p_tele_boxes( const __grid_constant__ SgPortals2D ps, const uint32_t n, const double2* __restrict__ q1, const double* __restrict__ r,
                                                          const uint32_t* __restrict__ tflag, const uint32_t* __restrict__ toff, double* __restrict__ boxes, uint32_t* __restrict__ box_body,
                                                          uint32_t* __restrict__ box_portal )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const uint32_t p = blockIdx.y;
  const size_t idx = size_t( p ) * n + b;
  if( tflag[idx] == 0u ) { return; }
  const uint32_t slot = toff[idx];
  const double2 x = __ldg( &q1[b] );
  const double rad = __ldg( &r[b] );
  const SgVec2 xin{ x.x, x.y };
  const int touch = sg_portal_touch( ps.p[p], xin, rad );
  const SgVec2 xo = sg_portal_teleport_ball( ps.p[p], xin, rad );
  double2* o = reinterpret_cast<double2*>( boxes + ( size_t( n ) + slot ) * 4 );
  o[0] = make_double2( xo.x - rad, xo.y - rad );
  o[1] = make_double2( xo.x + rad, xo.y + rad );
  box_body[slot] = b;
  box_portal[slot] = p | ( touch == 2 ? SG_PLANE_B_BIT : 0u );
}

// Per candidate of the extended box set (ball2d/Ball2DSim.cpp:424-487).  EMIT == false: flags only.  EMIT == true: regular
// contacts at reg_off[k] (candidate order == active_set order), teleported collisions appended at tel_off[k].
template<bool EMIT>
__global__ void __launch_bounds__( 128 ) k_b2p_pairs( const __grid_constant__ SgPortals2D ps, const uint32_t n, const uint2* __restrict__ pairs, const unsigned long long npairs,
                                                     const double2* __restrict__ q0, const double2* __restrict__ q1, const double* __restrict__ r,
                                                     const uint32_t* __restrict__ box_body, const uint32_t* __restrict__ box_portal,
                                                     uint32_t* __restrict__ reg_cnt, uint32_t* __restrict__ tel_cnt, const uint32_t* __restrict__ reg_off, const uint32_t* __restrict__ tel_off,
                                                     const ContactOut2D out, unsigned long long* __restrict__ tc_key, uint32_t* __restrict__ tc_idx, uint4* __restrict__ tc_info )
{
  const unsigned long long k = blockIdx.x * ( unsigned long long )( blockDim.x ) + threadIdx.x;
  if( k >= npairs ) { return; }
  const uint2 pr = pairs[k];
  const bool first_teleported = pr.x >= n;
  const bool second_teleported = pr.y >= n;
  if( !first_teleported && !second_teleported )
  {
    const double2 x1a = __ldg( &q1[pr.x] ), x1b = __ldg( &q1[pr.y] );
    const double ra = __ldg( &r[pr.x] ), rb = __ldg( &r[pr.y] );
    if( !EMIT )
    {
      reg_cnt[k] = sg_ball_ball_active( SgVec2{ x1a.x, x1a.y }, SgVec2{ x1b.x, x1b.y }, ra, rb ) ? 1u : 0u;
      tel_cnt[k] = 0u;
    }
    else if( reg_cnt[k] != 0u )
    {
      // BallBallConstraint{ i, j, q0, ri, rj, false }: n = ( q0_i - q0_j ).normalized(); point q0_i - ri n; depth at q1
      const unsigned long long o = reg_off[k];
      if( o < out.cap )
      {
        const double2 x0a = __ldg( &q0[pr.x] ), x0b = __ldg( &q0[pr.y] );
        double nx = x0a.x - x0b.x;
        double ny = x0a.y - x0b.y;
        const double z = nx * nx + ny * ny;
        if( z > 0.0 ) { const double s = sqrt( z ); nx = nx / s; ny = ny / s; }
        const double ex = x1a.x - x1b.x;
        const double ey = x1a.y - x1b.y;
        out.type[o] = SG_BALL_BALL; out.i[o] = pr.x; out.j[o] = pr.y;
        out.n[o] = make_double2( nx, ny );
        out.p[o] = make_double2( x0a.x - ra * nx, x0a.y - ra * ny );
        out.depth[o] = fmin( 0.0, sqrt( ex * ex + ey * ey ) - ( ra + rb ) );
      }
    }
    return;
  }
  uint32_t bdy0 = pr.x, bdy1 = pr.y, prtl0 = SG_NO_PORTAL, prtl1 = SG_NO_PORTAL;
  if( first_teleported ) { bdy0 = __ldg( &box_body[pr.x - n] ); prtl0 = __ldg( &box_portal[pr.x - n] ); }
  if( second_teleported ) { bdy1 = __ldg( &box_body[pr.y - n] ); prtl1 = __ldg( &box_portal[pr.y - n] ); }
  const SgTeleCollision c = sg_tele_collision( bdy0, bdy1, prtl0, prtl1 );
  if( !EMIT )
  {
    const double2 xa = __ldg( &q1[c.b0] ), xb = __ldg( &q1[c.b1] );
    const double ra = __ldg( &r[c.b0] ), rb = __ldg( &r[c.b1] );
    bool happens = true;
    // both copies teleported and the un-teleported bodies collide as well: found there (Ball2DSim.cpp:471-480)
    if( first_teleported && second_teleported && sg_ball_ball_active( SgVec2{ xa.x, xa.y }, SgVec2{ xb.x, xb.y }, ra, rb ) ) { happens = false; }
    if( happens )
    {
      const SgVec2 ta = sg_tele_center( ps, c.p0, SgVec2{ xa.x, xa.y } );
      const SgVec2 tb = sg_tele_center( ps, c.p1, SgVec2{ xb.x, xb.y } );
      happens = sg_ball_ball_active( ta, tb, ra, rb );
    }
    reg_cnt[k] = 0u;
    tel_cnt[k] = happens ? 1u : 0u;
  }
  else if( tel_cnt[k] != 0u )
  {
    const uint32_t o = tel_off[k];
    tc_key[o] = ( ( unsigned long long )( c.b0 ) << 32 ) | c.b1;
    tc_idx[o] = o;
    tc_info[o] = make_uint4( c.b0, c.b1, c.p0, c.p1 );
  }
}

// generateTeleportedBallBallCollision (ball2d/Ball2DSim.cpp:653-728)
__global__ void __launch_bounds__( 128 ) k_b2p_tele_contacts( const __grid_constant__ SgPortals2D ps, const uint32_t nraw, const uint32_t* __restrict__ idxs, const uint32_t* __restrict__ uflag,
                                                             const uint32_t* __restrict__ uoff, const uint4* __restrict__ tc_info, const double2* __restrict__ q0, const double2* __restrict__ q1,
                                                             const double* __restrict__ r, const unsigned long long base, const ContactOut2D out, double2* __restrict__ x0t,
                                                             double2* __restrict__ x1t, double2* __restrict__ kick, uint32_t* __restrict__ tp0, uint32_t* __restrict__ tp1 )
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if( e >= nraw || uflag[e] == 0u ) { return; }
  const uint32_t s = uoff[e];
  const uint4 c = tc_info[idxs[e]];
  const uint32_t b0 = c.x, b1 = c.y, p0 = c.z, p1 = c.w;
  const double2 q0a = __ldg( &q0[b0] ), q0b = __ldg( &q0[b1] );
  const double ri = __ldg( &r[b0] ), rj = __ldg( &r[b1] );
  const SgVec2 x0 = sg_tele_center( ps, p0, SgVec2{ q0a.x, q0a.y } );
  const SgVec2 x1 = sg_tele_center( ps, p1, SgVec2{ q0b.x, q0b.y } );
  double nx = x0.x - x1.x;
  double ny = x0.y - x1.y;
  const double z = nx * nx + ny * ny;
  if( z > 0.0 ) { const double sq = sqrt( z ); nx = nx / sq; ny = ny / sq; }
  const bool le0 = p0 != SG_NO_PORTAL && ps.p[p0 & ~SG_PLANE_B_BIT].v != 0.0;
  const bool le1 = p1 != SG_NO_PORTAL && ps.p[p1 & ~SG_PLANE_B_BIT].v != 0.0;
  SgVec2 kk{ 0.0, 0.0 };
  if( le1 )
  {
    const double2 x = __ldg( &q1[b1] );
    kk = sg_portal_kinematic_velocity_of_ball( ps.p[p1 & ~SG_PLANE_B_BIT], SgVec2{ x.x, x.y }, rj );
  }
  else if( le0 )
  {
    const double2 x = __ldg( &q1[b0] );
    const SgVec2 kv = sg_portal_kinematic_velocity_of_ball( ps.p[p0 & ~SG_PLANE_B_BIT], SgVec2{ x.x, x.y }, ri );
    kk.x = -kv.x; kk.y = -kv.y;
  }
  const unsigned long long o = base + s;
  if( o < out.cap )
  {
    out.type[o] = ( le0 || le1 ) ? SG_BALL_BALL_KICK_TELEPORTED : SG_BALL_BALL_TELEPORTED;
    out.i[o] = b0; out.j[o] = b1;
    out.n[o] = make_double2( nx, ny );
    out.p[o] = make_double2( q0a.x - ri * nx, q0a.y - ri * ny ); // getWorldSpaceContactPoint( q0 ): the body's own position
    out.depth[o] = __longlong_as_double( 0x7ff8000000000000LL ); // computePenetrationDepth: NaN when teleported
  }
  x0t[s] = make_double2( x0.x, x0.y ); x1t[s] = make_double2( x1.x, x1.y ); kick[s] = make_double2( kk.x, kk.y );
  tp0[s] = p0; tp1[s] = p1;
}

// Ball2DSim::enforcePeriodicBoundaryConditions on (q, v) in place (ball2d/Ball2DSim.cpp:336-366)
__global__ void __launch_bounds__( 256 ) k_b2p_enforce( const __grid_constant__ SgPortals2D ps, const uint32_t n, double2* __restrict__ q, double2* __restrict__ v )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const double2 x = q[b], w = v[b];
  SgVec2 xs{ x.x, x.y }, vs{ w.x, w.y };
  sg_portals_enforce( ps, xs, vs );
  q[b] = make_double2( xs.x, xs.y );
  v[b] = make_double2( vs.x, vs.y );
}

#endif
