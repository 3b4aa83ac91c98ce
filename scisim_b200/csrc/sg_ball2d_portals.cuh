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
//   sort      bitonic network on (body pair, insertion number), tiles of 2048 in shared memory (one launch for the usual
//             few thousand boundary collisions); first of each body pair survives == std::set::insert
//   contacts  teleported contacts behind the regular ones (plain or kinematic-kick), then drums and planes as always
#include "sg_boxes.cuh"
#include "sg_portal2d.h"
#include "sg_pair_sort_host.cuh"

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
#include "sg_ball2d_portal_kernels.cuh"

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
    rc = sg_bp_emit_lists<PortalBoxPolicy>( ctx, x->bp, in, next, true, NoOut{}, 0u );
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
    rc = sg_tele_sort_unique( ctx, nraw, m, x->tc_key.as<unsigned long long>(), x->tc_idx.as<uint32_t>(), x->uflag.as<uint32_t>(), x->uoff.as<uint32_t>(), x->u_partials.as<uint32_t>(), x->utotal.as<uint32_t>() );
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
