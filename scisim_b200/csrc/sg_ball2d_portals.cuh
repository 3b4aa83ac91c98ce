// sg_ball2d_portals.cuh -- ball2d active set with planar / Lees-Edwards portals (included by sg_ball2d.cu).
//
// Reference behaviour reproduced (file:line in the SCISim checkout):
//   ball2d/Ball2DSim.cpp:368-546   computeBallBallActiveSetSpatialGridWithPortals: boxes at q1 (NOT swept), one teleported box
//                                  per (portal, touching ball) appended portal-major after the real ones, getPotentialOverlaps
//                                  over all of them, BallBallConstraint::isActive at q1 for real-real pairs (no CCD on this
//                                  path), TeleportedCollision set for the rest
//   ball2d/Ball2DSim.cpp:610-728   teleported centres, teleportedBallBallCollisionHappens, generateTeleportedBallBallCollision
//   ball2d/Portals/PlanarPortal.cpp, ball2d/StaticGeometry/StaticPlane.cpp   (sg_portal2d.h)
//   ball2d/Ball2DSim.cpp:336-366   enforcePeriodicBoundaryConditions          (k_b2p_enforce)
//
// Pipeline (every list in the reference's order, nothing sorted except the few teleported collisions):
//   touch     one thread per (portal, ball): does the ball at q1 reach plane A or B?  flag by (portal-major) position
//   scan      flags -> teleported box numbers                                        (sg_scan.cuh)
//   boxes     real boxes q1 -+ r; teleported boxes at n + number, with the (ball, portal | plane) table
//   broad     the generic box pipeline of sg_broadphase.cuh over n + T boxes -> ascending candidate list
//   pairs     one thread per candidate: real-real -> isActive( q1 ); otherwise the TeleportedCollision tests.  Count, two
//             scans, then emit: regular contacts land in candidate order, teleported collisions in a (key, insertion) list
//   sort      bitonic network on (body pair, insertion number); first of each body pair survives == std::set::insert
//   contacts  teleported contacts behind the regular ones (plain or kinematic-kick), then drums and planes as always
#include "sg_boxes.cuh"
#include "sg_portal2d.h"

using PortalBoxPolicy = AabbPolicy<2, 1>;

struct PortalData
{
  SgPortals2D portals;
  BroadScratch bp;
  DevBuf boxes;                              // double[4 * ( n + T )]
  DevBuf tflag, toff, t_partials, ttotal;    // u32[P * n] flags and box numbers; T
  DevBuf box_body, box_portal;               // u32[T]: TeleportedBall table
  DevBuf err;                                // u32: a ball touches both planes of one portal
  DevBuf reg_cnt, reg_off, tel_cnt, tel_off, pr_partials, reg_total, tel_total;
  DevBuf tc_key, tc_idx, tc_info;            // u64[M], u32[M] (M = padded power of two), uint4[raw]: { b0, b1, p0, p1 }
  DevBuf uflag, uoff, u_partials, utotal;
  DevBuf x0t, x1t, kick, tp0, tp1;           // per teleported contact: constructor arguments of the constraint
  DevBuf base;                               // ScanPairCounts::Acc { candidates, body-body contacts } for the static emit
  PinBuf h;                                  // counts read back
  PinBuf h_tele;                             // sg_ball2d_teleported staging
  uint64_t n_boxes = 0, n_reg = 0, n_tel = 0;
  PortalData() { memset( &portals, 0, sizeof( portals ) ); }
  void release()
  {
    DevBuf* bufs[] = { &boxes, &tflag, &toff, &t_partials, &ttotal, &box_body, &box_portal, &err, &reg_cnt, &reg_off, &tel_cnt, &tel_off, &pr_partials, &reg_total, &tel_total,
                       &tc_key, &tc_idx, &tc_info, &uflag, &uoff, &u_partials, &utotal, &x0t, &x1t, &kick, &tp0, &tp1, &base };
    for( DevBuf* b : bufs ) { b->release(); }
    bp.release(); h.release(); h_tele.release();
  }
};

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

__global__ void __launch_bounds__( 256 ) k_b2p_sort_pad( const uint32_t first, const uint32_t m, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idxs )
{
  const uint32_t e = first + blockIdx.x * blockDim.x + threadIdx.x;
  if( e < m ) { keys[e] = ~0ull; idxs[e] = ~0u; }
}

// one compare-exchange step of the bitonic network (the lower element of each pair does the work)
__global__ void __launch_bounds__( 256 ) k_b2p_bitonic( const uint32_t m, const uint32_t j, const uint32_t k, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idxs )
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if( e >= m ) { return; }
  const uint32_t f = sg_bitonic_partner( e, j );
  if( f <= e ) { return; }
  const unsigned long long ka = keys[e], kb = keys[f];
  const uint32_t ia = idxs[e], ib = idxs[f];
  const bool a_less = sg_tele_less( ka, ia, kb, ib );
  const bool b_less = sg_tele_less( kb, ib, ka, ia );
  const bool swap = sg_bitonic_ascending( e, k ) ? b_less : a_less;
  if( swap ) { keys[e] = kb; keys[f] = ka; idxs[e] = ib; idxs[f] = ia; }
}

// std::set<TeleportedCollision>::insert keeps the first collision of each body pair: after the sort that is the first
// entry of each run of equal keys
__global__ void __launch_bounds__( 256 ) k_b2p_unique( const uint32_t nraw, const unsigned long long* __restrict__ keys, uint32_t* __restrict__ uflag )
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if( e >= nraw ) { return; }
  uflag[e] = ( e == 0u || keys[e] != keys[e - 1u] ) ? 1u : 0u;
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

// counts[g * nblocks + block] = balls of the block active against static geometry g at q1 (as the tail of k_ball2d_prep)
__global__ void __launch_bounds__( 256 ) k_b2p_static_count( const __grid_constant__ Static2D sg, const uint32_t n, const double2* __restrict__ q1, const double* __restrict__ r, uint32_t* __restrict__ counts )
{
  __shared__ uint32_t s_cnt[SG_MAX_DRUMS + SG_MAX_PLANES];
  const uint32_t ng = sg.ndrums + sg.nplanes;
  if( threadIdx.x < ng ) { s_cnt[threadIdx.x] = 0u; }
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long mask = 0ull;
  if( i < n ) { mask = static_mask( sg, __ldg( &q1[i] ), __ldg( &r[i] ) ); }
  const int lane = threadIdx.x & 31;
  for( uint32_t g = 0; g < ng; ++g )
  {
    const unsigned b = __ballot_sync( 0xffffffffu, ( mask >> g ) & 1ull );
    if( lane == 0 && b != 0u ) { atomicAdd( &s_cnt[g], __popc( b ) ); }
  }
  __syncthreads();
  if( threadIdx.x < ng ) { counts[threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x]; }
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

// ---- host driver -----------------------------------------------------------------------------------
static PortalData* ball2d_portal_data( Ball2DData* d )
{
  if( d->px == nullptr ) { d->px = new PortalData; }
  return d->px;
}

static int ball2d_portal_active_set_device( sg_ctx* ctx, Ball2DData* d, const int flow_kind, const double dt )
{
  PortalData* x = ball2d_portal_data( d );
  const uint32_t n = d->n;
  const uint32_t np_portals = x->portals.n;
  d->n_cand = d->n_bb = d->n_static = d->n_drum = d->n_plane = 0;
  x->n_boxes = x->n_reg = x->n_tel = 0;
  d->have_result = true;
  d->cand_valid = true;
  d->portal_result = true;
  if( n == 0 ) { return SG_OK; }
  if( uint64_t( n ) * np_portals >= 0x80000000ull ) { return sg_fail( ctx, SG_ERR_INVALID, "ball2d portals: balls x portals must stay below 2^31" ); }
  const unsigned nblk = sg_div_up( n, 256 );
  if( flow_kind >= 0 )
  {
    SG_LAUNCH( ctx, "ball2d_flow", double( n ) * 72.0, k_ball2d_flow<<<nblk, 256, 0, ctx->stream>>>( flow_kind, n, d->Q0(), d->v0.as<double2>(), d->m.as<double>(), d->g[0], d->g[1], dt, d->Q1(), d->v1.as<double2>() ) );
  }
  const uint32_t nflag = n * np_portals;
  SG_CUDA( ctx, x->h.ensure( 128 ) );
  SG_CUDA( ctx, x->tflag.ensure( size_t( nflag ) * 4 + 4 ) ); SG_CUDA( ctx, x->toff.ensure( size_t( nflag ) * 4 + 4 ) );
  SG_CUDA( ctx, x->t_partials.ensure( ( size_t( nflag ) / SG_SCAN_TILE + 2 ) * 4 ) );
  SG_CUDA( ctx, x->ttotal.ensure( 4 ) ); SG_CUDA( ctx, x->err.ensure( 4 ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->ttotal.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->err.ptr, 0, 4, ctx->stream ) );
  SG_LAUNCH( ctx, "b2p_touch", double( nflag ) * 28.0, k_b2p_touch<<<dim3( nblk, np_portals ), 256, 0, ctx->stream>>>( x->portals, n, d->Q1(), d->R(), x->tflag.as<uint32_t>(), x->err.as<uint32_t>() ) );
  int rc = sg_exclusive_scan<ScanU32>( ctx, "b2p_touch_scan", x->tflag.as<uint32_t>(), nullptr, nflag, nflag, x->t_partials.as<uint32_t>(), x->toff.as<uint32_t>(), x->ttotal.as<uint32_t>(), false );
  if( rc != SG_OK ) { return rc; }
  // drums and planes do not depend on the portals: count them while the teleported boxes are being numbered
  const uint32_t ng = d->sg.ndrums + d->sg.nplanes;
  const uint32_t nst = ng * nblk;
  if( ng > 0 )
  {
    rc = ball2d_static_scratch( ctx, d );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, cudaMemsetAsync( d->st_total.ptr, 0, 4, ctx->stream ) );
    SG_LAUNCH( ctx, "b2p_static_count", double( n ) * 24.0, k_b2p_static_count<<<nblk, 256, 0, ctx->stream>>>( d->sg, n, d->Q1(), d->R(), d->st_counts.as<uint32_t>() ) );
    rc = sg_exclusive_scan<ScanU32>( ctx, "ball2d_static_scan", d->st_counts.as<uint32_t>(), nullptr, nst, nst, d->st_partials.as<uint32_t>(), d->st_offsets.as<uint32_t>(), d->st_total.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
  }
  uint32_t* h32 = x->h.as<uint32_t>();
  h32[2] = 0u;
  SG_CUDA( ctx, cudaMemcpyAsync( h32, x->ttotal.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( h32 + 1, x->err.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  if( ng > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( h32 + 2, d->st_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  if( h32[1] != 0u )
  {
    return sg_fail( ctx, SG_ERR_UNSUPPORTED, "a ball touches both planes of one portal (the reference exits here: ball2d/Portals/PlanarPortal.cpp:117-121)" );
  }
  const uint32_t nt = h32[0];
  d->n_static = h32[2];
  x->n_boxes = nt;
  if( uint64_t( n ) + nt >= 0x80000000ull ) { return sg_fail( ctx, SG_ERR_INVALID, "ball2d portals: more than 2^31 - 1 boxes" ); }
  const uint32_t next = n + nt;
  SG_CUDA( ctx, x->boxes.ensure( size_t( next ) * 32 ) );
  SG_CUDA( ctx, x->box_body.ensure( size_t( nt ) * 4 + 4 ) ); SG_CUDA( ctx, x->box_portal.ensure( size_t( nt ) * 4 + 4 ) );
  SG_LAUNCH( ctx, "b2p_boxes", double( n ) * 56.0, k_b2p_boxes<<<nblk, 256, 0, ctx->stream>>>( n, d->Q1(), d->R(), x->boxes.as<double>() ) );
  if( nt > 0 )
  {
    SG_LAUNCH( ctx, "b2p_tele_boxes", double( nflag ) * 8.0 + double( nt ) * 64.0, k_b2p_tele_boxes<<<dim3( nblk, np_portals ), 256, 0, ctx->stream>>>( x->portals, n, d->Q1(), d->R(), x->tflag.as<uint32_t>(),
               x->toff.as<uint32_t>(), x->boxes.as<double>(), x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>() ) );
  }
  // broad phase over real + teleported boxes
  rc = sg_bp_prepare_scratch<PortalBoxPolicy>( ctx, x->bp, next );
  if( rc != SG_OK ) { return rc; }
  PortalBoxPolicy::In in;
  in.boxes = x->boxes.as<double>(); in.n = next;
  rc = sg_bp_bin_and_count<PortalBoxPolicy>( ctx, x->bp, in );
  if( rc != SG_OK ) { return rc; }
  unsigned long long* h64 = x->h.as<unsigned long long>() + 4;
  SG_CUDA( ctx, cudaMemcpyAsync( h64, x->bp.totals.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  const uint64_t np = h64[0];
  d->n_cand = np;
  if( np >= 0xffffffffull ) { return sg_fail( ctx, SG_ERR_INTERNAL, "ball2d portals: more than 2^32 candidate pairs" ); }
  if( np + 64 > x->bp.cand_cap ) { SG_CUDA( ctx, x->bp.cand.ensure( size_t( np + 64 ) * sizeof( uint2 ) ) ); x->bp.cand_cap = x->bp.cand.cap / sizeof( uint2 ); }
  uint32_t nraw = 0;
  if( np > 0 )
  {
    rc = sg_bp_emit_lists<PortalBoxPolicy>( ctx, x->bp, next, true, NoOut{}, 0u );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, x->reg_cnt.ensure( size_t( np ) * 4 + 4 ) ); SG_CUDA( ctx, x->reg_off.ensure( size_t( np ) * 4 + 4 ) );
    SG_CUDA( ctx, x->tel_cnt.ensure( size_t( np ) * 4 + 4 ) ); SG_CUDA( ctx, x->tel_off.ensure( size_t( np ) * 4 + 4 ) );
    SG_CUDA( ctx, x->pr_partials.ensure( ( size_t( np ) / SG_SCAN_TILE + 2 ) * 4 ) );
    SG_CUDA( ctx, x->reg_total.ensure( 4 ) ); SG_CUDA( ctx, x->tel_total.ensure( 4 ) );
    SG_CUDA( ctx, cudaMemsetAsync( x->reg_total.ptr, 0, 4, ctx->stream ) );
    SG_CUDA( ctx, cudaMemsetAsync( x->tel_total.ptr, 0, 4, ctx->stream ) );
    const ContactOut2D none{};
    SG_LAUNCH( ctx, "b2p_pairs_count", double( np ) * 96.0, k_b2p_pairs<false><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( x->portals, n, x->bp.cand.as<uint2>(), np, d->Q0(), d->Q1(), d->R(),
               x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>(), x->reg_cnt.as<uint32_t>(), x->tel_cnt.as<uint32_t>(), nullptr, nullptr, none, nullptr, nullptr, nullptr ) );
    rc = sg_exclusive_scan<ScanU32>( ctx, "b2p_regular_scan", x->reg_cnt.as<uint32_t>(), nullptr, uint32_t( np ), uint32_t( np ), x->pr_partials.as<uint32_t>(), x->reg_off.as<uint32_t>(), x->reg_total.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
    rc = sg_exclusive_scan<ScanU32>( ctx, "b2p_teleported_scan", x->tel_cnt.as<uint32_t>(), nullptr, uint32_t( np ), uint32_t( np ), x->pr_partials.as<uint32_t>(), x->tel_off.as<uint32_t>(), x->tel_total.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, cudaMemcpyAsync( h32 + 4, x->reg_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h32 + 5, x->tel_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    x->n_reg = h32[4];
    nraw = h32[5];
  }
  if( nraw > 0x40000000u ) { return sg_fail( ctx, SG_ERR_INTERNAL, "ball2d portals: more than 2^30 teleported collisions" ); }
  // every list size is bounded now: contacts = regular + (at most nraw) teleported + static
  rc = ball2d_ensure_outputs( ctx, d, 0u, x->n_reg + nraw + d->n_static + 64u );
  if( rc != SG_OK ) { return rc; }
  ContactOut2D out;
  out.type = d->c_type.as<uint32_t>(); out.i = d->c_i.as<uint32_t>(); out.j = d->c_j.as<uint32_t>();
  out.n = d->c_n.as<double2>(); out.p = d->c_p.as<double2>(); out.depth = d->c_depth.as<double>();
  out.cap = d->act_cap;
  out.gid = GidMap();
  uint32_t m = 1u;
  while( m < nraw ) { m <<= 1; }
  if( nraw > 0 )
  {
    SG_CUDA( ctx, x->tc_key.ensure( size_t( m ) * 8 ) ); SG_CUDA( ctx, x->tc_idx.ensure( size_t( m ) * 4 ) ); SG_CUDA( ctx, x->tc_info.ensure( size_t( nraw ) * 16 ) );
    SG_CUDA( ctx, x->uflag.ensure( size_t( nraw ) * 4 + 4 ) ); SG_CUDA( ctx, x->uoff.ensure( size_t( nraw ) * 4 + 4 ) );
    SG_CUDA( ctx, x->u_partials.ensure( ( size_t( nraw ) / SG_SCAN_TILE + 2 ) * 4 ) ); SG_CUDA( ctx, x->utotal.ensure( 4 ) );
    SG_CUDA( ctx, x->x0t.ensure( size_t( nraw ) * 16 ) ); SG_CUDA( ctx, x->x1t.ensure( size_t( nraw ) * 16 ) ); SG_CUDA( ctx, x->kick.ensure( size_t( nraw ) * 16 ) );
    SG_CUDA( ctx, x->tp0.ensure( size_t( nraw ) * 4 ) ); SG_CUDA( ctx, x->tp1.ensure( size_t( nraw ) * 4 ) );
  }
  if( np > 0 && ( x->n_reg > 0 || nraw > 0 ) )
  {
    SG_LAUNCH( ctx, "b2p_pairs_emit", double( np ) * 40.0 + double( x->n_reg ) * 130.0 + double( nraw ) * 28.0, k_b2p_pairs<true><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( x->portals, n, x->bp.cand.as<uint2>(), np,
               d->Q0(), d->Q1(), d->R(), x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>(), x->reg_cnt.as<uint32_t>(), x->tel_cnt.as<uint32_t>(), x->reg_off.as<uint32_t>(), x->tel_off.as<uint32_t>(), out,
               x->tc_key.as<unsigned long long>(), x->tc_idx.as<uint32_t>(), x->tc_info.as<uint4>() ) );
  }
  if( nraw > 0 )
  {
    if( m > nraw ) { SG_LAUNCH( ctx, "b2p_sort_pad", 0.0, k_b2p_sort_pad<<<sg_div_up( m - nraw, 256 ), 256, 0, ctx->stream>>>( nraw, m, x->tc_key.as<unsigned long long>(), x->tc_idx.as<uint32_t>() ) ); }
    for( uint32_t k = 2u; k <= m; k <<= 1 )
    {
      for( uint32_t j = k >> 1; j > 0u; j >>= 1 )
      {
        SG_LAUNCH( ctx, "b2p_bitonic", double( m ) * 24.0, k_b2p_bitonic<<<sg_div_up( m, 256 ), 256, 0, ctx->stream>>>( m, j, k, x->tc_key.as<unsigned long long>(), x->tc_idx.as<uint32_t>() ) );
      }
    }
    SG_CUDA( ctx, cudaMemsetAsync( x->utotal.ptr, 0, 4, ctx->stream ) );
    SG_LAUNCH( ctx, "b2p_unique", double( nraw ) * 12.0, k_b2p_unique<<<sg_div_up( nraw, 256 ), 256, 0, ctx->stream>>>( nraw, x->tc_key.as<unsigned long long>(), x->uflag.as<uint32_t>() ) );
    rc = sg_exclusive_scan<ScanU32>( ctx, "b2p_unique_scan", x->uflag.as<uint32_t>(), nullptr, nraw, nraw, x->u_partials.as<uint32_t>(), x->uoff.as<uint32_t>(), x->utotal.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
    SG_LAUNCH( ctx, "b2p_tele_contacts", double( nraw ) * 200.0, k_b2p_tele_contacts<<<sg_div_up( nraw, 128 ), 128, 0, ctx->stream>>>( x->portals, nraw, x->tc_idx.as<uint32_t>(), x->uflag.as<uint32_t>(), x->uoff.as<uint32_t>(),
               x->tc_info.as<uint4>(), d->Q0(), d->Q1(), d->R(), x->n_reg, out, x->x0t.as<double2>(), x->x1t.as<double2>(), x->kick.as<double2>(), x->tp0.as<uint32_t>(), x->tp1.as<uint32_t>() ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h32 + 6, x->utotal.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    x->n_tel = h32[6];
  }
  d->n_bb = x->n_reg + x->n_tel;
  if( ng > 0 && d->n_static > 0 )
  {
    // the static emit takes its base (number of body-body contacts) from a ScanPairCounts::Acc on the device
    SG_CUDA( ctx, x->base.ensure( sizeof( ScanPairCounts::Acc ) ) );
    unsigned long long* hb = x->h.as<unsigned long long>() + 8;
    hb[0] = d->n_cand; hb[1] = d->n_bb;
    SG_CUDA( ctx, cudaMemcpyAsync( x->base.ptr, hb, 16, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_LAUNCH( ctx, "ball2d_static_emit", double( n ) * 40.0 + double( d->n_static ) * 52.0, k_ball2d_static_emit<<<nblk, 256, 0, ctx->stream>>>( d->sg, n, d->Q0(), d->Q1(), d->R(),
               d->st_counts.as<uint32_t>(), d->st_offsets.as<uint32_t>(), x->base.as<ScanPairCounts::Acc>(), out, 0u, n ) );
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}
