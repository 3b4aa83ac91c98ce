// sg_ball2d.cu -- ball2d hot path: unconstrained flow, broad phase + ball-ball CCD, drum and plane tests.
//
// Reference behaviour reproduced (file:line in the SCISim checkout):
//   ball2d/SymplecticEulerMap.cpp:21-32, ball2d/VerletMap.cpp:15-41           k_ball2d_flow
//   ball2d/Forces/Ball2DGravityForce.cpp:36-46, ball2d/Ball2DState.cpp:54-66  (gravity, Minv = 1.0/m)
//   ball2d/Ball2DSim.cpp:553-608   swept AABBs + getPotentialOverlaps + CCD    Ball2DPolicy + sg_broadphase.cuh
//   scisim/CollisionDetection/CollisionDetectionUtilities.cpp:3-121           ccd_hit()
//   ball2d/Constraints/BallBallConstraint.cpp:22-37,219-222,270-280           normal / point / depth
//   ball2d/Ball2DSim.cpp:730-762   drum-major and plane-major all-pairs loops   k_ball2d_static_*
//   ball2d/Constraints/BallStaticPlaneConstraint.cpp:10-15, BallStaticDrumConstraint.cpp:8-26
//
// HBM layout: q0,v0,q1,v1 are SCISim's own interleaved [x,y] vectors, i.e. arrays of double2 (one
// 128-bit access per ball); r and m are plain double arrays.  Sorted 64-byte records are described in
// sg_broadphase.cuh.  Contacts are SoA (type,i,j,n,p,depth) in the reference's active_set order.
#include "sg_broadphase.cuh"
#include "sg_slab.cuh"

struct Ball2DIn
{
  const double2* q0;
  const double2* q1;
  const double* r;
  uint32_t n;
  uint32_t own_first; // bodies [own_first, own_first + own_count) belong to this rank, the rest are halo ghosts
  uint32_t own_count;
  // slab mode: slots [0, ghost_counts[0]) and [own_first + own_count, ... + ghost_counts[1]) hold ghosts, the other
  // non-owned slots are unused this step.  nullptr on one GPU (every slot is a body).
  const uint32_t* ghost_counts;
  const uint32_t* gid; // slab mode: global body index of every slot; nullptr on one GPU (identity)
};

__device__ __forceinline__ bool ball2d_slot_valid( const Ball2DIn& in, const uint32_t i )
{
  if( in.ghost_counts == nullptr || i - in.own_first < in.own_count ) { return true; }
  return ( i < in.own_first ) ? ( i < __ldg( &in.ghost_counts[0] ) ) : ( i - ( in.own_first + in.own_count ) < __ldg( &in.ghost_counts[1] ) );
}

#define SG_GHOST_BIT 0x80000000u

struct alignas( 64 ) Ball2DRec
{
  double q0x, q0y, q1x, q1y;
  double r;
  uint32_t idx;
  uint32_t key;
  uint32_t c1, c2; // cell row (and layer)
  uint32_t gid;    // ORDER word: global body index (== idx on one GPU); ranks the bodies in the emitted lists and is what they carry
  uint32_t pad1;
};

struct ContactOut2D
{
  uint32_t* type;
  uint32_t* i;
  uint32_t* j;
  double2* n;
  double2* p;
  double* depth;
  unsigned long long cap;
  GidMap gid; // multi-GPU: local -> global body index (identity on one GPU)
};

// scisim/CollisionDetection/CollisionDetectionUtilities.cpp:3-121 (a = lower body index): sg_ccd.h
__device__ __forceinline__ bool ccd_hit( const Ball2DRec& a, const Ball2DRec& b )
{
  return sg_ccd_ball_ball( a.q0x, a.q0y, a.q1x, a.q1y, a.r, b.q0x, b.q0y, b.q1x, b.q1y, b.r );
}

struct Ball2DPolicy
{
  static constexpr int D = 2;
  static constexpr bool HAS_NARROW = true;
  static constexpr double IN_BYTES = 40.0;
  using In = Ball2DIn;
  using Rec = Ball2DRec;
  using Out = ContactOut2D;
  static constexpr uint32_t IDX_MASK = 0x7fffffffu;
  static constexpr uint32_t IDX_OFFSET = 40u;
  static constexpr uint32_t ORD_OFFSET = 56u;
  static constexpr bool ORD_IN_REC = true;

  // ball2d/Ball2DSim.cpp:566-571: lo = min(q1,q0) - r, hi = max(q1,q0) + r
  __device__ static void load_aabb( const In& in, const uint32_t i, double* lo, double* hi )
  {
    const double2 a = __ldg( &in.q0[i] );
    const double2 b = __ldg( &in.q1[i] );
    const double r = __ldg( &in.r[i] );
    lo[0] = fmin( b.x, a.x ) - r; lo[1] = fmin( b.y, a.y ) - r;
    hi[0] = fmax( b.x, a.x ) + r; hi[1] = fmax( b.y, a.y ) + r;
  }
  __device__ static Rec make_rec( const In& in, const uint32_t i, const uint32_t key, const uint32_t c1, const uint32_t c2 )
  {
    const double2 a = __ldg( &in.q0[i] );
    const double2 b = __ldg( &in.q1[i] );
    Rec rec;
    rec.q0x = a.x; rec.q0y = a.y; rec.q1x = b.x; rec.q1y = b.y;
    rec.r = __ldg( &in.r[i] );
    rec.idx = ( i - in.own_first < in.own_count ) ? i : ( i | SG_GHOST_BIT ); rec.key = key; rec.c1 = c1; rec.c2 = c2;
    rec.gid = ( in.gid != nullptr ) ? __ldg( &in.gid[i] ) : i; rec.pad1 = 0u;
    return rec;
  }
  __device__ static void rec_aabb( const Rec& s, double* lo, double* hi )
  {
    lo[0] = fmin( s.q1x, s.q0x ) - s.r; lo[1] = fmin( s.q1y, s.q0y ) - s.r;
    hi[0] = fmax( s.q1x, s.q0x ) + s.r; hi[1] = fmax( s.q1y, s.q0y ) + s.r;
  }
  __device__ static uint32_t rec_idx( const Rec& s ) { return s.idx & IDX_MASK; }
  __device__ static uint32_t rec_idx_raw( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_ord( const Rec& s ) { return s.gid; }
  __device__ static uint32_t rec_ord_raw( const Rec& s ) { return s.gid; }
  __device__ static bool owns( const Rec& s ) { return ( s.idx & SG_GHOST_BIT ) == 0u; }
  __device__ static bool valid( const In& in, const uint32_t i ) { return ball2d_slot_valid( in, i ); }
  __device__ static uint32_t rec_key( const Rec& s ) { return s.key; }
  __device__ static uint32_t rec_c1( const Rec& s, const GridParams& ) { return s.c1; }
  __device__ static uint32_t rec_c2( const Rec& s, const GridParams& ) { return s.c2; }
  __device__ static Rec load_pass1( const Rec* __restrict__ p ) { return sg_load_rec_global<Rec>( p ); }
  __device__ static bool narrow_test( const Rec& a, const Rec& b ) { return ccd_hit( a, b ); }
  // BallBallConstraint{ i, j, q0a, q0b, ra, rb }: n = (q0a - q0b).normalized(); point q0a - ra*n; depth at q1
  __device__ static void contact_emit( const Out& out, unsigned long long& k, const Rec& a, const Rec& b )
  {
    if( k < out.cap )
    {
      double nx = a.q0x - b.q0x;
      double ny = a.q0y - b.q0y;
      const double z = nx * nx + ny * ny;
      if( z > 0.0 ) { const double s = sqrt( z ); nx = nx / s; ny = ny / s; }
      const double ex = a.q1x - b.q1x;
      const double ey = a.q1y - b.q1y;
      out.type[k] = SG_BALL_BALL;
      out.i[k] = a.gid;
      out.j[k] = b.gid;
      out.n[k] = make_double2( nx, ny );
      out.p[k] = make_double2( a.q0x - a.r * nx, a.q0y - a.r * ny );
      out.depth[k] = fmin( 0.0, sqrt( ex * ex + ey * ey ) - ( a.r + b.r ) );
    }
    ++k;
  }
};

// ---- unconstrained flow ----------------------------------------------------------------------------
// Arithmetic order follows Eigen's evaluation of the reference expressions:
//   F = 0 + m*g;  SE: v1 = v0 + (0 + (dt*minv)*F), q1 = q0 + dt*v1
//   Verlet: vh = v0 + (0 + ((0.5*dt)*minv)*F), q1 = q0 + dt*vh, v1 = vh + ((0.5*dt)*minv)*F(q1)
__global__ void __launch_bounds__( 256 ) k_ball2d_flow( const int kind, const uint32_t n, const double2* __restrict__ q0, const double2* __restrict__ v0, const double* __restrict__ m,
                                                       const double gx, const double gy, const double dt, double2* __restrict__ q1, double2* __restrict__ v1 )
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) { return; }
  const double2 q = __ldg( &q0[i] );
  const double2 v = __ldg( &v0[i] );
  const double mass = __ldg( &m[i] );
  const double minv = 1.0 / mass;
  const double Fx = 0.0 + mass * gx;
  const double Fy = 0.0 + mass * gy;
  double2 qo, vo;
  if( kind == SG_MAP_SYMPLECTIC_EULER )
  {
    const double s = dt * minv;
    vo.x = v.x + ( 0.0 + s * Fx );
    vo.y = v.y + ( 0.0 + s * Fy );
    qo.x = q.x + dt * vo.x;
    qo.y = q.y + dt * vo.y;
  }
  else
  {
    const double s = ( 0.5 * dt ) * minv;
    const double vhx = v.x + ( 0.0 + s * Fx );
    const double vhy = v.y + ( 0.0 + s * Fy );
    qo.x = q.x + dt * vhx;
    qo.y = q.y + dt * vhy;
    vo.x = vhx + s * Fx;
    vo.y = vhy + s * Fy;
  }
  q1[i] = qo;
  v1[i] = vo;
}

// ---- static geometry (drums, then planes) ----------------------------------------------------------
struct Static2D
{
  uint32_t ndrums;
  uint32_t nplanes;
  double drum_x[SG_MAX_DRUMS], drum_y[SG_MAX_DRUMS], drum_r[SG_MAX_DRUMS];
  double plane_x[SG_MAX_PLANES], plane_y[SG_MAX_PLANES], plane_nx[SG_MAX_PLANES], plane_ny[SG_MAX_PLANES];
};

// bit g of the result: geometry g (drums first, then planes) is active for this ball at q1
__device__ __forceinline__ unsigned long long static_mask( const Static2D& sg, const double2 x1, const double r )
{
  unsigned long long mask = 0ull;
  for( uint32_t d = 0; d < sg.ndrums; ++d )
  {
    // ( X - q ).squaredNorm() >= ( R - r ) * ( R - r )
    const double dx = sg.drum_x[d] - x1.x;
    const double dy = sg.drum_y[d] - x1.y;
    const double Rr = sg.drum_r[d] - r;
    if( dx * dx + dy * dy >= Rr * Rr ) { mask |= 1ull << d; }
  }
  for( uint32_t p = 0; p < sg.nplanes; ++p )
  {
    // n.dot( q - x ) <= r
    const double dist = sg.plane_nx[p] * ( x1.x - sg.plane_x[p] ) + sg.plane_ny[p] * ( x1.y - sg.plane_y[p] );
    if( dist <= r ) { mask |= 1ull << ( sg.ndrums + p ); }
  }
  return mask;
}

// One pass over the balls that (optionally) integrates them and, from registers, also produces everything the
// detection pipeline needs before binning: the bounds of the swept AABBs (block reduce + one atomic per quantity
// per block) and counts[g * nblocks + block] = number of this block's balls active against static geometry g.
template<bool DO_FLOW>
__global__ void __launch_bounds__( 256 ) k_ball2d_prep( const __grid_constant__ Static2D sg, const int kind, const uint32_t n, const double2* __restrict__ q0, const double2* __restrict__ v0,
                                                       const double* __restrict__ m, const double* __restrict__ r, const double gx, const double gy, const double dt,
                                                       double2* __restrict__ q1, double2* __restrict__ v1, BoundsAccum* __restrict__ acc, uint32_t* __restrict__ counts,
                                                       const uint32_t own_first, const uint32_t own_count, const uint32_t* __restrict__ ghost_counts, long long* __restrict__ interval_enc,
                                                       double2* __restrict__ block_iv, const double xlim_lo, const double xlim_hi, uint32_t* __restrict__ slab_flags, const SlabCand sc )
{
  // q0, q1, r are indexed by slot ([ghosts | owned | ghosts] in slab mode); v0, m, v1 exist for owned bodies only.
  // interval_enc != nullptr (slab mode, before the exchange): the ghosts have not arrived yet -- only owned bodies are live, and
  // [min lo.x, max hi.x] of their swept boxes is reduced for the neighbours; the ghosts' share of the bounds is added
  // by the unpack kernel.
  __shared__ uint32_t s_cnt[SG_MAX_DRUMS + SG_MAX_PLANES];
  const uint32_t ng = sg.ndrums + sg.nplanes;
  if( threadIdx.x < ng ) { s_cnt[threadIdx.x] = 0u; }
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double mn[2] = { __longlong_as_double( 0x7ff0000000000000LL ), __longlong_as_double( 0x7ff0000000000000LL ) };
  double mx[2] = { __longlong_as_double( 0xfff0000000000000LL ), __longlong_as_double( 0xfff0000000000000LL ) };
  double ext = 0.0;
  unsigned long long mask = 0ull;
  bool live = i < n;
  if( live && interval_enc != nullptr ) { live = i - own_first < own_count; }
  else if( live && ghost_counts != nullptr && i - own_first >= own_count )
  {
    live = ( i < own_first ) ? ( i < __ldg( &ghost_counts[0] ) ) : ( i - ( own_first + own_count ) < __ldg( &ghost_counts[1] ) );
  }
  double ivlo = __longlong_as_double( 0x7ff0000000000000LL ), ivhi = __longlong_as_double( 0xfff0000000000000LL );
  if( live )
  {
    const double2 q = __ldg( &q0[i] );
    double2 qo;
    if( DO_FLOW )
    {
      const uint32_t io = i - own_first;
      const double2 v = __ldg( &v0[io] );
      const double mass = __ldg( &m[io] );
      const double minv = 1.0 / mass;
      const double Fx = 0.0 + mass * gx;
      const double Fy = 0.0 + mass * gy;
      double2 vo;
      if( kind == SG_MAP_SYMPLECTIC_EULER )
      {
        const double s = dt * minv;
        vo.x = v.x + ( 0.0 + s * Fx );
        vo.y = v.y + ( 0.0 + s * Fy );
        qo.x = q.x + dt * vo.x;
        qo.y = q.y + dt * vo.y;
      }
      else
      {
        const double s = ( 0.5 * dt ) * minv;
        const double vhx = v.x + ( 0.0 + s * Fx );
        const double vhy = v.y + ( 0.0 + s * Fy );
        qo.x = q.x + dt * vhx;
        qo.y = q.y + dt * vhy;
        vo.x = vhx + s * Fx;
        vo.y = vhy + s * Fy;
      }
      q1[i] = qo;
      v1[io] = vo;
    }
    else
    {
      qo = __ldg( &q1[i] );
    }
    const double rad = __ldg( &r[i] );
    double lo[2], hi[2];
    lo[0] = fmin( qo.x, q.x ) - rad; lo[1] = fmin( qo.y, q.y ) - rad;
    hi[0] = fmax( qo.x, q.x ) + rad; hi[1] = fmax( qo.y, q.y ) + rad;
    ivlo = lo[0]; ivhi = hi[0];
    // slab mode: an owned body whose swept box leaves [xlim_lo, xlim_hi] (this slab widened by half of each neighbouring slab) could reach a
    // body of a non-neighbouring rank, which no halo would carry: flag it, sg_ball2d_slab_detect then asks for a re-partition
    if( slab_flags != nullptr && ( lo[0] < xlim_lo || hi[0] > xlim_hi ) ) { slab_flags[3] = 1u; }
    sg_bp_bounds_update<2>( lo, hi, mn, mx, ext );
    if( i - own_first < own_count ) { mask = static_mask( sg, qo, rad ); } // ghosts touch no static geometry here
  }
  if( sc.band != nullptr )
  {
    // candidate lists for the halo pack (order is irrelevant: the lists are ranked by global index later)
    #pragma unroll
    for( int sd = 0; sd < 2; ++sd )
    {
      if( !sc.on[sd] ) { continue; }
      const bool c = live && ( ( sd == 0 ) ? ( ivlo <= sc.band[0] ) : ( ivhi >= sc.band[1] ) );
      const unsigned bal = __ballot_sync( 0xffffffffu, c );
      if( bal != 0u )
      {
        const int lane = threadIdx.x & 31;
        uint32_t base = 0u;
        if( lane == __ffs( bal ) - 1 ) { base = atomicAdd( &sc.count[sd], uint32_t( __popc( bal ) ) ); }
        base = __shfl_sync( 0xffffffffu, base, __ffs( bal ) - 1 );
        const uint32_t k = base + __popc( bal & ( ( 1u << lane ) - 1u ) );
        if( c && k < sc.cap ) { sc.list[sd][k] = i; }
      }
    }
  }
  sg_bp_bounds_commit<2>( mn, mx, ext, acc );
  if( interval_enc != nullptr )
  {
    // block reduce, then two atomics per block (same-address atomics from every warp would serialise in L2)
    __shared__ double s_iv[8][2];
    #pragma unroll
    for( int dd = 16; dd > 0; dd >>= 1 ) { ivlo = fmin( ivlo, __shfl_xor_sync( 0xffffffffu, ivlo, dd ) ); ivhi = fmax( ivhi, __shfl_xor_sync( 0xffffffffu, ivhi, dd ) ); }
    if( ( threadIdx.x & 31 ) == 0 ) { s_iv[threadIdx.x >> 5][0] = ivlo; s_iv[threadIdx.x >> 5][1] = ivhi; }
    __syncthreads();
    if( threadIdx.x == 0 )
    {
      for( int w = 1; w < 8; ++w ) { ivlo = fmin( ivlo, s_iv[w][0] ); ivhi = fmax( ivhi, s_iv[w][1] ); }
      if( block_iv != nullptr ) { block_iv[blockIdx.x] = make_double2( ivlo, ivhi ); } // lets the halo pack skip whole blocks
      if( ivlo <= ivhi )
      {
        atomicMin( &interval_enc[0], sg_ordered_from_double( ivlo ) );
        atomicMax( &interval_enc[1], sg_ordered_from_double( ivhi ) );
      }
    }
  }
  if( ng > 0u )
  {
    const int lane = threadIdx.x & 31;
    for( uint32_t g = 0; g < ng; ++g )
    {
      const unsigned b = __ballot_sync( 0xffffffffu, ( mask >> g ) & 1ull );
      if( lane == 0 && b != 0u ) { atomicAdd( &s_cnt[g], __popc( b ) ); }
    }
    __syncthreads();
    if( threadIdx.x < ng ) { counts[threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x]; }
  }
}

// Stable compaction: geometry-major, ball ascending, appended after the body-body contacts
__global__ void __launch_bounds__( 256 ) k_ball2d_static_emit( const __grid_constant__ Static2D sg, const uint32_t n, const double2* __restrict__ q0, const double2* __restrict__ q1, const double* __restrict__ r,
                                                              const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, const ScanPairCounts::Acc* __restrict__ pair_totals, const ContactOut2D out,
                                                              const uint32_t own_first, const uint32_t own_count )
{
  __shared__ uint32_t s_warp[8];
  const uint32_t ng = sg.ndrums + sg.nplanes;
  // most blocks touch no static geometry: find that out from the counts before reading any ball
  const int mine = ( threadIdx.x < ng ) ? int( counts[threadIdx.x * gridDim.x + blockIdx.x] != 0u ) : 0;
  if( __syncthreads_or( mine ) == 0 ) { return; }
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2 x0 = make_double2( 0.0, 0.0 ), x1 = x0;
  double rad = 0.0;
  unsigned long long mask = 0ull;
  if( i < n && i - own_first < own_count ) { x0 = __ldg( &q0[i] ); x1 = __ldg( &q1[i] ); rad = __ldg( &r[i] ); mask = static_mask( sg, x1, rad ); }
  const unsigned long long base = pair_totals->a;
  for( uint32_t g = 0; g < ng; ++g )
  {
    if( counts[g * gridDim.x + blockIdx.x] == 0u ) { continue; } // uniform across the block
    const bool act = ( mask >> g ) & 1ull;
    const unsigned b = __ballot_sync( 0xffffffffu, act );
    __syncthreads();
    if( lane == 0 ) { s_warp[warp] = __popc( b ); }
    __syncthreads();
    uint32_t before = 0u;
    for( int w = 0; w < warp; ++w ) { before += s_warp[w]; }
    if( act )
    {
      const unsigned long long k = base + offsets[g * gridDim.x + blockIdx.x] + before + __popc( b & ( ( 1u << lane ) - 1u ) );
      if( k < out.cap )
      {
        double nx, ny, depth;
        uint32_t type, j;
        if( g < sg.ndrums )
        {
          // StaticDrumConstraint: n = ( X - q0_i ).normalized(); no depth override (NaN)
          type = SG_BALL_DRUM; j = g;
          nx = sg.drum_x[g] - x0.x; ny = sg.drum_y[g] - x0.y;
          const double z = nx * nx + ny * ny;
          if( z > 0.0 ) { const double s = sqrt( z ); nx = nx / s; ny = ny / s; }
          depth = __longlong_as_double( 0x7ff8000000000000LL );
        }
        else
        {
          const uint32_t p = g - sg.ndrums;
          type = SG_BALL_PLANE; j = p;
          nx = sg.plane_nx[p]; ny = sg.plane_ny[p];
          const double dist = nx * ( x1.x - sg.plane_x[p] ) + ny * ( x1.y - sg.plane_y[p] );
          depth = fmin( 0.0, dist - rad );
        }
        out.type[k] = type;
        out.i[k] = out.gid( i );
        out.j[k] = j;
        out.n[k] = make_double2( nx, ny );
        out.p[k] = make_double2( x0.x - rad * nx, x0.y - rad * ny );
        out.depth[k] = depth;
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
struct PortalData; // sg_ball2d_portals.cuh

struct Ball2DData
{
  uint32_t n = 0;
  double g[2] = { 0.0, 0.0 };
  Static2D sg;
  DevBuf r, m, q0, v0, q1, v1;
  BroadScratch bp;
  // static geometry scratch
  DevBuf st_counts, st_offsets, st_partials, st_total;
  // contact SoA
  DevBuf c_type, c_i, c_j, c_n, c_p, c_depth;
  uint64_t act_cap = 0;
  // host staging
  PinBuf h_totals;  // {P_c, P_a, n_static}
  PinBuf h_out;     // contacts copied back
  // results of the last active-set computation
  uint64_t n_cand = 0, n_bb = 0, n_static = 0, n_drum = 0, n_plane = 0;
  bool have_result = false;
  bool cand_valid = false;
  // planar / Lees-Edwards portals (sg_ball2d_portals.cuh): allocated by sg_ball2d_set_portals
  PortalData* px = nullptr;
  struct AsmData* asmd = nullptr; // device-side assembly of N, Q, contact bases and the impulse cache (sg_ball2d_assembly.cuh)
  bool portal_result = false; // the last result came from the portal path (its candidate list lives in px->bp)
  // slab mode (one slab of a larger scene per GPU): the body arrays hold [left ghosts | owned | right ghosts] with
  // ghost_cap slots reserved on either side of the owned block (slot order == global index order); how many of
  // them are in use this step is only known on the device (ghost_counts), unused slots are skipped by every kernel.
  bool slab = false;
  uint32_t n_owned = 0, ghost_cap = 0, gid_first = 0;
  double xlim[2] = { -1.0e308, 1.0e308 }; // owned swept boxes must stay inside (sg_ball2d_slab_set_gids); outside => SG_ERR_REBALANCE
  DevBuf gid;          // u32 per slot: global body index
  DevBuf interval_enc; // 2 x long long (ordered encoding of min lo.x / max hi.x over the owned swept boxes)
  DevBuf pack_counts, pack_offsets, pack_partials, pack_total;
  // peer-memory halo exchange (NVLink): this rank's mailbox and the neighbours' mailboxes mapped into this process
  DevBuf mailbox;
  void* peer_mb[2] = { nullptr, nullptr };
  bool peer_ipc[2] = { false, false };
  bool flow_resident = false; // q0 (as given) and q1 (as computed) of the last sg_ball2d_flow are still in q0/q1 on the device
  uint32_t slab_step = 0; // tag of the current step's flags (all ranks step in lockstep)
  bool slab_prep_done = false; // this step's bounds / static counts were already produced by sg_ball2d_slab_flow
  DevBuf pack_done;            // block counter of the pack kernel's "last block raises the flag"
  DevBuf cand_state;           // SlabCandState: bands, candidate counters, pack cursors and tickets (peer-memory exchange)
  DevBuf cand_list;            // u32[2][cand_cap]: slots of the bodies that could be owed to the lower / higher neighbour this step
  uint32_t cand_cap = 0;
  DevBuf block_iv;             // double2 per 256-slot block: [min lo.x, max hi.x] of its owned swept boxes (from the flow kernel)
  size_t first_slot() const { return 0; }
  size_t owned_slot() const { return slab ? size_t( ghost_cap ) : 0; }        // first owned slot
  uint32_t own_first() const { return slab ? ghost_cap : 0u; }                 // owned range in local indices
  DevBuf ghost_counts; // u32[4]: ghosts on side 0, side 1, overflow flag, spare
  const uint32_t* GHOSTS() const { return slab ? ghost_counts.as<uint32_t>() : nullptr; }
  uint32_t own_count() const { return slab ? n_owned : n; }
  double2* Q0() const { return q0.as<double2>() + first_slot(); }
  double2* Q1() const { return q1.as<double2>() + first_slot(); }
  double* R() const { return r.as<double>() + first_slot(); }
  const uint32_t* GID() const { return slab ? gid.as<uint32_t>() + first_slot() : nullptr; }
  GidMap gid_map() const
  {
    GidMap g;
    if( slab ) { g.gid = gid.as<uint32_t>(); }
    return g;
  }
  Ball2DData() { memset( &sg, 0, sizeof( sg ) ); }
};

static void ball2d_portal_release( Ball2DData* d );
static void ball2d_asm_release( Ball2DData* d );
static bool ball2d_has_portals( const Ball2DData* d );
static int ball2d_portal_active_set_device( sg_ctx* ctx, Ball2DData* d, const int flow_kind, const double dt );
static const DevBuf& ball2d_result_candidates( const Ball2DData* d );

void sg_ball2d_release( sg_ctx* ctx )
{
  Ball2DData* d = ctx->ball2d;
  if( d == nullptr ) { return; }
  d->r.release(); d->m.release(); d->q0.release(); d->v0.release(); d->q1.release(); d->v1.release();
  d->bp.release();
  d->st_counts.release(); d->st_offsets.release(); d->st_partials.release(); d->st_total.release();
  d->c_type.release(); d->c_i.release(); d->c_j.release(); d->c_n.release(); d->c_p.release(); d->c_depth.release();
  d->h_totals.release(); d->h_out.release();
  d->gid.release(); d->ghost_counts.release(); d->interval_enc.release(); d->pack_counts.release(); d->pack_offsets.release(); d->pack_partials.release(); d->pack_total.release(); d->pack_done.release(); d->block_iv.release(); d->cand_state.release(); d->cand_list.release();
  for( int sde = 0; sde < 2; ++sde ) { if( d->peer_mb[sde] != nullptr && d->peer_ipc[sde] ) { cudaIpcCloseMemHandle( d->peer_mb[sde] ); } d->peer_mb[sde] = nullptr; }
  d->mailbox.release();
  ball2d_portal_release( d );
  ball2d_asm_release( d );
  delete d;
  ctx->ball2d = nullptr;
}

static Ball2DData* ball2d_data( sg_ctx* ctx )
{
  if( ctx->ball2d == nullptr ) { ctx->ball2d = new Ball2DData; }
  return ctx->ball2d;
}

static int ball2d_ensure_outputs( sg_ctx* ctx, Ball2DData* d, const uint64_t cand_cap, const uint64_t act_cap )
{
  if( cand_cap > d->bp.cand_cap )
  {
    SG_CUDA( ctx, d->bp.cand.ensure( size_t( cand_cap ) * sizeof( uint2 ) ) );
    d->bp.cand_cap = d->bp.cand.cap / sizeof( uint2 );
  }
  if( act_cap > d->act_cap )
  {
    SG_CUDA( ctx, d->c_type.ensure( size_t( act_cap ) * 4 ) );
    SG_CUDA( ctx, d->c_i.ensure( size_t( act_cap ) * 4 ) );
    SG_CUDA( ctx, d->c_j.ensure( size_t( act_cap ) * 4 ) );
    SG_CUDA( ctx, d->c_n.ensure( size_t( act_cap ) * 16 ) );
    SG_CUDA( ctx, d->c_p.ensure( size_t( act_cap ) * 16 ) );
    SG_CUDA( ctx, d->c_depth.ensure( size_t( act_cap ) * 8 ) );
    d->act_cap = act_cap;
  }
  return SG_OK;
}

// Runs the whole detection pipeline on the device-resident q0,q1 and leaves the counts in d->n_*.
// flow_kind >= 0: the unconstrained map is fused into the first pass (q1,v1 are produced from q0,v0 on the way).
static int ball2d_static_scratch( sg_ctx* ctx, Ball2DData* d )
{
  const uint32_t ng = d->sg.ndrums + d->sg.nplanes;
  const uint32_t nst = ng * sg_div_up( d->n, 256 );
  SG_CUDA( ctx, d->st_counts.ensure( size_t( nst ) * 4 + 4 ) );
  SG_CUDA( ctx, d->st_offsets.ensure( size_t( nst ) * 4 + 4 ) );
  SG_CUDA( ctx, d->st_partials.ensure( ( size_t( nst ) / SG_SCAN_TILE + 2 ) * 4 ) );
  SG_CUDA( ctx, d->st_total.ensure( 4 ) );
  return SG_OK;
}

static int ball2d_active_set_device( sg_ctx* ctx, Ball2DData* d, const bool want_cand, const int flow_kind = -1, const double dt = 0.0 )
{
  const uint32_t n = d->n;
  // with portals the reference takes another road altogether (Ball2DSim.cpp:159-166): boxes at q1, no CCD, teleported copies
  if( ball2d_has_portals( d ) ) { return ball2d_portal_active_set_device( ctx, d, flow_kind, dt ); }
  d->portal_result = false;
  d->n_cand = d->n_bb = d->n_static = d->n_drum = d->n_plane = 0;
  d->have_result = true;
  d->cand_valid = want_cand;
  if( n == 0 ) { return SG_OK; }
  int rc = sg_bp_prepare_scratch<Ball2DPolicy>( ctx, d->bp, n );
  if( rc != SG_OK ) { return rc; }
  SG_CUDA( ctx, d->h_totals.ensure( 64 ) );
  // first guess at the list sizes; grown (and the emit re-run) if a step overflows them
  rc = ball2d_ensure_outputs( ctx, d, want_cand ? ( d->bp.cand_cap > 0 ? d->bp.cand_cap : uint64_t( n ) * 6u + 1024u ) : 0u, d->act_cap > 0 ? d->act_cap : uint64_t( n ) * 4u + 1024u );
  if( rc != SG_OK ) { return rc; }

  const uint32_t ng = d->sg.ndrums + d->sg.nplanes;
  const unsigned nblk = sg_div_up( n, 256 );
  const uint32_t nst = ng * nblk;
  rc = ball2d_static_scratch( ctx, d );
  if( rc != SG_OK ) { return rc; }
  if( d->slab && d->slab_prep_done )
  {
    // slab mode: the flow kernel already reduced the owned bodies' bounds and static counts, the unpack kernels
    // added the ghosts' bounds
    d->slab_prep_done = false;
  }
  else if( flow_kind >= 0 && !d->slab )
  {
    SG_LAUNCH( ctx, "ball2d_flow_prep", double( n ) * ( 72.0 + 8.0 ), k_ball2d_prep<true><<<nblk, 256, 0, ctx->stream>>>( d->sg, flow_kind, n, d->Q0(), d->v0.as<double2>(), d->m.as<double>(),
               d->R(), d->g[0], d->g[1], dt, d->Q1(), d->v1.as<double2>(), d->bp.bounds_cur(), d->st_counts.as<uint32_t>(), 0u, n, nullptr, nullptr, nullptr, 0.0, 0.0, nullptr, SlabCand{ nullptr, nullptr, { nullptr, nullptr }, 0u, { false, false } } ) );
  }
  else
  {
    SG_LAUNCH( ctx, "ball2d_prep", double( n ) * 40.0, k_ball2d_prep<false><<<nblk, 256, 0, ctx->stream>>>( d->sg, 0, n, d->Q0(), nullptr, nullptr,
               d->R(), 0.0, 0.0, 0.0, d->Q1(), nullptr, d->bp.bounds_cur(), d->st_counts.as<uint32_t>(), d->own_first(), d->own_count(), d->GHOSTS(), nullptr, nullptr, 0.0, 0.0, nullptr, SlabCand{ nullptr, nullptr, { nullptr, nullptr }, 0u, { false, false } } ) );
  }
  Ball2DIn in;
  in.q0 = d->Q0(); in.q1 = d->Q1(); in.r = d->R(); in.n = n; in.own_first = d->own_first(); in.own_count = d->own_count(); in.ghost_counts = d->GHOSTS(); in.gid = d->GID();
  d->bp.ord_by_index = d->GID();
  // the (tiny, latency-bound) scan of the static-geometry counts rides along with the pair-count scan
  const bool side_scan = ng > 0 && nst <= SG_SIDE_SCAN_MAX;
  if( side_scan ) { d->bp.side.in = d->st_counts.as<uint32_t>(); d->bp.side.n = nst; d->bp.side.out = d->st_offsets.as<uint32_t>(); d->bp.side.total = d->st_total.as<uint32_t>(); }
  rc = sg_bp_bin_and_count<Ball2DPolicy>( ctx, d->bp, in, true );
  if( rc != SG_OK ) { return rc; }

  if( ng > 0 && !side_scan )
  {
    rc = sg_exclusive_scan<ScanU32>( ctx, "ball2d_static_scan", d->st_counts.as<uint32_t>(), nullptr, nst, nst, d->st_partials.as<uint32_t>(), d->st_offsets.as<uint32_t>(), d->st_total.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
  }

  for( int attempt = 0; attempt < 2; ++attempt )
  {
    ContactOut2D out;
    out.type = d->c_type.as<uint32_t>(); out.i = d->c_i.as<uint32_t>(); out.j = d->c_j.as<uint32_t>();
    out.n = d->c_n.as<double2>(); out.p = d->c_p.as<double2>(); out.depth = d->c_depth.as<double>();
    out.cap = d->act_cap;
    out.gid = d->gid_map();
    // The static-geometry contacts go behind the body-body ones (their base is the pair scan's total) and touch
    // nothing the pair emit does, so the two kernels run side by side (serially when kernels are being timed).
    cudaStream_t side = ( ng > 0 && !ctx->profile ) ? ctx->stream2 : ctx->stream;
    if( side != ctx->stream )
    {
      SG_CUDA( ctx, cudaEventRecord( ctx->ev_fork, ctx->stream ) );
      SG_CUDA( ctx, cudaStreamWaitEvent( side, ctx->ev_fork, 0 ) );
    }
    rc = sg_bp_emit_lists<Ball2DPolicy>( ctx, d->bp, in, n, want_cand, out, d->act_cap );
    if( rc != SG_OK ) { return rc; }
    if( ng > 0 )
    {
      SG_LAUNCH( ctx, "ball2d_static_emit", double( n ) * 40.0, k_ball2d_static_emit<<<nblk, 256, 0, side>>>( d->sg, n, d->Q0(), d->Q1(), d->R(),
                 d->st_counts.as<uint32_t>(), d->st_offsets.as<uint32_t>(), d->bp.totals.as<ScanPairCounts::Acc>(), out, d->own_first(), d->own_count() ) );
    }
    if( side != ctx->stream )
    {
      SG_CUDA( ctx, cudaEventRecord( ctx->ev_join, side ) );
      SG_CUDA( ctx, cudaStreamWaitEvent( ctx->stream, ctx->ev_join, 0 ) );
    }
    // counts to the host
    unsigned long long* ht = d->h_totals.as<unsigned long long>();
    ht[2] = 0ull;
    SG_CUDA( ctx, cudaMemcpyAsync( ht, d->bp.totals.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
    if( ng > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( ht + 2, d->st_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) ); }
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    sg_prof_collect( ctx );
    d->n_cand = ht[0];
    d->bp.dense = ht[0] * 2ull >= 5ull * uint64_t( n ); // >= 2.5 candidates per body: the next step's pass 1 stages the records (sg_bp_count_staged)
    d->n_bb = ht[1];
    d->n_static = ht[2] & 0xffffffffull;
    if( ctx->profile )
    {
      // list sizes are only known now: add the emitted bytes to the kernels that wrote them
      ctx->prof[sg_prof_entry( ctx, "bp_emit" )].bytes += ( want_cand ? double( d->n_cand ) * 8.0 : 0.0 ) + double( d->n_bb ) * 8.0;
      ctx->prof[sg_prof_entry( ctx, "bp_contacts" )].bytes += double( d->n_bb ) * ( 8.0 + 52.0 ) + double( n ) * 64.0; // every record is needed once; repeats hit L1/L2
      if( ng > 0 ) { ctx->prof[sg_prof_entry( ctx, "ball2d_static_emit" )].bytes += double( d->n_static ) * 52.0; }
    }
    const uint64_t need_act = d->n_bb + d->n_static;
    const bool cand_ok = !want_cand || d->n_cand <= d->bp.cand_cap;
    if( cand_ok && need_act <= d->act_cap ) { break; }
    if( attempt == 1 ) { return sg_fail( ctx, SG_ERR_INTERNAL, "ball2d: output lists still overflow after regrowth" ); }
    rc = ball2d_ensure_outputs( ctx, d, want_cand ? d->n_cand + d->n_cand / 8 + 1024 : 0u, need_act + need_act / 8 + 1024 );
    if( rc != SG_OK ) { return rc; }
  }
  return SG_OK;
}

// Copies the last result's lists into pinned memory and fills *out.  Counts per static type are derived
// on the host from the type array tail (drums precede planes).
static int ball2d_copy_out( sg_ctx* ctx, Ball2DData* d, const uint32_t flags, sg_contacts* out )
{
  memset( out, 0, sizeof( *out ) );
  out->dim = 2;
  out->n_candidates = d->n_cand;
  out->n_body_body = d->n_bb;
  const uint64_t na = d->n_bb + d->n_static;
  out->n_active = na;
  const bool want_cand = ( flags & SG_OUT_CANDIDATES ) != 0u && d->cand_valid;
  size_t bytes = 64;
  const size_t o_type = bytes; bytes += ( na * 4 + 63 ) & ~size_t( 63 );
  const size_t o_i = bytes; bytes += ( na * 4 + 63 ) & ~size_t( 63 );
  const size_t o_j = bytes; bytes += ( na * 4 + 63 ) & ~size_t( 63 );
  const size_t o_n = bytes; if( flags & SG_OUT_NORMALS ) { bytes += ( na * 16 + 63 ) & ~size_t( 63 ); }
  const size_t o_p = bytes; if( flags & SG_OUT_POINTS ) { bytes += ( na * 16 + 63 ) & ~size_t( 63 ); }
  const size_t o_d = bytes; if( flags & SG_OUT_DEPTHS ) { bytes += ( na * 8 + 63 ) & ~size_t( 63 ); }
  const size_t o_c = bytes; if( want_cand ) { bytes += ( d->n_cand * 8 + 63 ) & ~size_t( 63 ); }
  SG_CUDA( ctx, d->h_out.ensure( bytes ) );
  char* h = d->h_out.as<char>();
  if( na > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_type, d->c_type.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_i, d->c_i.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_j, d->c_j.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    if( flags & SG_OUT_NORMALS ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_n, d->c_n.ptr, na * 16, cudaMemcpyDeviceToHost, ctx->stream ) ); }
    if( flags & SG_OUT_POINTS ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_p, d->c_p.ptr, na * 16, cudaMemcpyDeviceToHost, ctx->stream ) ); }
    if( flags & SG_OUT_DEPTHS ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_d, d->c_depth.ptr, na * 8, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  }
  if( want_cand && d->n_cand > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_c, ball2d_result_candidates( d ).ptr, d->n_cand * 8, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  out->type = reinterpret_cast<const uint32_t*>( h + o_type );
  out->i = reinterpret_cast<const uint32_t*>( h + o_i );
  out->j = reinterpret_cast<const uint32_t*>( h + o_j );
  out->n = ( flags & SG_OUT_NORMALS ) ? reinterpret_cast<const double*>( h + o_n ) : nullptr;
  out->p = ( flags & SG_OUT_POINTS ) ? reinterpret_cast<const double*>( h + o_p ) : nullptr;
  out->depth = ( flags & SG_OUT_DEPTHS ) ? reinterpret_cast<const double*>( h + o_d ) : nullptr;
  out->cand_ij = want_cand ? reinterpret_cast<const uint32_t*>( h + o_c ) : nullptr;
  uint64_t nd = 0;
  for( uint64_t k = d->n_bb; k < na; ++k ) { if( out->type[k] == SG_BALL_DRUM ) { ++nd; } else { break; } }
  d->n_drum = nd; d->n_plane = d->n_static - nd;
  out->n_drum = d->n_drum; out->n_plane = d->n_plane;
  return SG_OK;
}

#include "sg_ball2d_portals.cuh"
#include "sg_ball2d_assembly.cuh"

static void ball2d_asm_release( Ball2DData* d )
{
  if( d->asmd != nullptr ) { d->asmd->release(); delete d->asmd; d->asmd = nullptr; }
}
static void ball2d_portal_release( Ball2DData* d )
{
  if( d->px != nullptr ) { d->px->release(); delete d->px; d->px = nullptr; }
}
static bool ball2d_has_portals( const Ball2DData* d ) { return d->px != nullptr && d->px->portals.n > 0u; }
static const DevBuf& ball2d_result_candidates( const Ball2DData* d ) { return ( d->portal_result && d->px != nullptr ) ? d->px->bp.cand : d->bp.cand; }

// ---- slab mode: halo selection, packing and unpacking ----------------------------------------------
// Exchange buffers hold ghost_cap + 1 records of 48 bytes; record 0 is a header whose gid field is the count, so the
// receiver learns it on the device and the host never waits for it.
struct alignas( 16 ) GhostRec { double q0x, q0y, q1x, q1y, r; uint32_t gid; uint32_t pad; };

__device__ __forceinline__ void swept_x( const double2 a, const double2 b, const double r, double& lo, double& hi )
{
  lo = fmin( b.x, a.x ) - r;
  hi = fmax( b.x, a.x ) + r;
}

// What a pack / unpack launch has to synchronise with when the exchange goes through peer-mapped mailboxes
struct SlabSync
{
  const uint32_t* wait_flag; // local flag that must reach `step` before the kernel reads its input (nullptr: none)
  uint32_t* post_flag;       // peer flag raised (after a system fence) by the last block to finish (nullptr: none)
  uint32_t* done_ctr;        // block counter for "last block"
  uint32_t* err;
  uint32_t step;
};

// Owned bodies whose swept box overlaps [iv[0], iv[1]] on x (closed, like AABB::overlaps), in body order -- for up
// to two target intervals (both neighbours) in one pass over the bodies.
// Two launches: count per block, then emit -- a block that selected anything sums the (L2-resident, few KB) counts
// of the blocks before it instead of a separate scan launch; block 0 also writes the header with the total.
struct PackArgs
{
  const double* iv[2];
  uint32_t* block_counts[2];
  uint32_t* total[2];
  GhostRec* out[2];
  SlabSync sync[2];
  bool on[2];
};

// The grid covers all slots with the flow kernel's block partition, so that kernel's per-block x-interval (block_iv)
// tells a block whether any of its owned bodies can reach a target at all: almost every block stops there.
template<bool EMIT>
__global__ void __launch_bounds__( 256 ) k_ball2d_slab_pack( const uint32_t n, const uint32_t own_first, const uint32_t own_count, const double2* __restrict__ q0, const double2* __restrict__ q1,
                                                            const double* __restrict__ r, const uint32_t* __restrict__ gid, const double2* __restrict__ block_iv, const uint32_t cap, const PackArgs args )
{
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_red[8], s_nz[8];
  if( !EMIT )
  {
    if( threadIdx.x < 2 && args.on[threadIdx.x] && args.sync[threadIdx.x].wait_flag != nullptr )
    {
      slab_wait_flag( args.sync[threadIdx.x].wait_flag, args.sync[threadIdx.x].step, args.sync[threadIdx.x].err );
    }
    __syncthreads();
  }
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool owned = i < n && i - own_first < own_count;
  // can this block contribute to any target?  (block 0 always runs the emit pass: it writes the headers)
  bool reach[2] = { false, false };
  {
    const double2 biv = ( block_iv != nullptr ) ? block_iv[blockIdx.x] : make_double2( -1.0e300, 1.0e300 );
    #pragma unroll
    for( int sd = 0; sd < 2; ++sd ) { if( args.on[sd] ) { reach[sd] = !( biv.y < args.iv[sd][0] ) && !( args.iv[sd][1] < biv.x ); } }
  }
  if( !reach[0] && !reach[1] )
  {
    if( !EMIT )
    {
      if( threadIdx.x < 2 && args.on[threadIdx.x] ) { args.block_counts[threadIdx.x][blockIdx.x] = 0u; }
      return;
    }
    if( blockIdx.x != 0u ) { return; }
  }
  double2 a = make_double2( 0.0, 0.0 ), b = a;
  double rad = 0.0, lo = 0.0, hi = 0.0;
  if( owned )
  {
    a = __ldg( &q0[i] ); b = __ldg( &q1[i] ); rad = __ldg( &r[i] );
    swept_x( a, b, rad, lo, hi );
  }
  #pragma unroll
  for( int sd = 0; sd < 2; ++sd )
  {
    if( !args.on[sd] ) { continue; }
    const SlabSync sync = args.sync[sd];
    uint32_t* block_counts = args.block_counts[sd];
    GhostRec* out = args.out[sd];
    // first read of the interval in this launch comes after the wait + barrier above: L1 cannot hold a stale line
    const double ilo = args.iv[sd][0], ihi = args.iv[sd][1];
    const bool sel = owned && !( hi < ilo ) && !( ihi < lo );
    const unsigned bal = __ballot_sync( 0xffffffffu, sel );
    if( lane == 0 ) { s_warp[warp] = __popc( bal ); }
    __syncthreads();
    uint32_t before = 0u, tot = 0u;
    for( int w = 0; w < 8; ++w ) { const uint32_t c = s_warp[w]; if( w < warp ) { before += c; } tot += c; }
    if( !EMIT )
    {
      if( threadIdx.x == 0 ) { block_counts[blockIdx.x] = tot; }
      __syncthreads();
      continue;
    }
    uint32_t participants = 1u; // block 0 only: itself + the other blocks that selected something
    if( tot != 0u || blockIdx.x == 0u )
    {
      // prefix of the preceding blocks' counts (block 0: the grand total, for the header)
      const uint32_t upto = ( blockIdx.x == 0u ) ? gridDim.x : blockIdx.x;
      uint32_t acc = 0u, nz = 0u;
      for( uint32_t bb = threadIdx.x; bb < upto; bb += blockDim.x ) { const uint32_t c = block_counts[bb]; acc += c; nz += ( c != 0u && bb != 0u ) ? 1u : 0u; }
      #pragma unroll
      for( int dd = 16; dd > 0; dd >>= 1 ) { acc += __shfl_xor_sync( 0xffffffffu, acc, dd ); nz += __shfl_xor_sync( 0xffffffffu, nz, dd ); }
      if( lane == 0 ) { s_red[warp] = acc; s_nz[warp] = nz; }
      __syncthreads();
      uint32_t sum = 0u;
      for( int w = 0; w < 8; ++w ) { sum += s_red[w]; participants += s_nz[w]; }
      if( blockIdx.x == 0u )
      {
        if( threadIdx.x == 0 )
        {
          GhostRec h;
          h.q0x = 0.0; h.q0y = 0.0; h.q1x = 0.0; h.q1y = 0.0; h.r = 0.0; h.gid = sum; h.pad = 0u;
          out[0] = h;
          if( args.total[sd] != nullptr ) { *args.total[sd] = sum; }
        }
        sum = 0u; // block 0 starts the list
      }
      if( sel )
      {
        const uint32_t k = sum + before + __popc( bal & ( ( 1u << lane ) - 1u ) );
        if( k < cap )
        {
          GhostRec g;
          g.q0x = a.x; g.q0y = a.y; g.q1x = b.x; g.q1y = b.y; g.r = rad; g.gid = gid[i]; g.pad = 0u;
          out[1u + k] = g;
        }
      }
      if( sync.post_flag != nullptr )
      {
        // The last of the blocks that wrote anything raises the neighbour's flag.  Each of them adds 1 to the counter,
        // block 0 (which knows how many there are) adds 1 - participants: exactly one add lands on zero, the last.
        // (One system-scope fence per block, by thread 0 after the barrier: the barrier orders the block's peer
        // writes before it and the fence is cumulative; the flag itself is a release store.)
        __syncthreads();
        if( threadIdx.x == 0 )
        {
          __threadfence_system();
          const int delta = ( blockIdx.x == 0u ) ? 1 - int( participants ) : 1;
          const int prev = atomicAdd( reinterpret_cast<int*>( sync.done_ctr ), delta );
          if( prev + delta == 0 ) { st_release_sys( sync.post_flag, sync.step ); }
        }
      }
    }
    __syncthreads();
  }
}

// what the shared halo pack (sg_slab.cuh) needs to know about balls
struct Ball2DSlabTraits
{
  using Rec = GhostRec;
  struct Src { const double2* q0; const double2* q1; const double* r; const uint32_t* gid; };
  __device__ static bool select( const Src& s, const uint32_t i, const double ilo, const double ihi, Rec& g )
  {
    const double2 a = __ldg( &s.q0[i] ), b = __ldg( &s.q1[i] );
    const double rad = __ldg( &s.r[i] );
    double lo, hi;
    swept_x( a, b, rad, lo, hi );
    g.q0x = a.x; g.q0y = a.y; g.q1x = b.x; g.q1y = b.y; g.r = rad; g.gid = s.gid[i]; g.pad = 0u;
    return !( hi < ilo ) && !( ihi < lo );
  }
};

static SlabCand ball2d_slab_cand( const Ball2DData* d )
{
  SlabCand sc;
  sc.band = nullptr; sc.count = nullptr; sc.list[0] = sc.list[1] = nullptr; sc.cap = 0u; sc.on[0] = sc.on[1] = false;
  if( d->mailbox.ptr == nullptr || d->cand_state.ptr == nullptr ) { return sc; }
  SlabCandState* st = d->cand_state.as<SlabCandState>();
  sc.band = st->band; sc.count = st->count;
  sc.list[0] = d->cand_list.as<uint32_t>(); sc.list[1] = d->cand_list.as<uint32_t>() + d->cand_cap;
  sc.cap = d->cand_cap;
  sc.on[0] = d->peer_mb[0] != nullptr; sc.on[1] = d->peer_mb[1] != nullptr;
  return sc;
}

// Count only (sg_ball2d_slab_pack with no send buffer): total of the per-block counts
__global__ void __launch_bounds__( 256 ) k_ball2d_slab_pack_total( const uint32_t nblk, const uint32_t* __restrict__ block_counts, uint32_t* __restrict__ total )
{
  __shared__ uint32_t s_red[8];
  uint32_t acc = 0u;
  for( uint32_t bb = threadIdx.x; bb < nblk; bb += blockDim.x ) { acc += block_counts[bb]; }
  #pragma unroll
  for( int dd = 16; dd > 0; dd >>= 1 ) { acc += __shfl_xor_sync( 0xffffffffu, acc, dd ); }
  if( ( threadIdx.x & 31 ) == 0 ) { s_red[threadIdx.x >> 5] = acc; }
  __syncthreads();
  if( threadIdx.x == 0 ) { uint32_t sum = 0u; for( int w = 0; w < 8; ++w ) { sum += s_red[w]; } *total = sum; }
}

// Ghost records into the slots either side of the owned block (both sides in one launch: the first half of the grid
// serves side 0, the second half side 1); their swept boxes join this step's bounds reduction (the owned bodies'
// share was reduced by the flow kernel).
struct UnpackArgs
{
  const GhostRec* in[2];
  uint32_t slot[2];   // first ghost slot of the side
  SlabSync sync[2];
  bool on[2];
};
__global__ void __launch_bounds__( 256 ) k_ball2d_slab_unpack( const uint32_t cap, const uint32_t blocks_per_side, const UnpackArgs args, double2* __restrict__ q0, double2* __restrict__ q1, double* __restrict__ r,
                                                              uint32_t* __restrict__ gid, uint32_t* __restrict__ ghost_counts, BoundsAccum* __restrict__ acc )
{
  const int side = ( blockIdx.x >= blocks_per_side ) ? 1 : 0;
  if( !args.on[side] ) { return; }
  const SlabSync sync = args.sync[side];
  const GhostRec* in = args.in[side];
  if( sync.wait_flag != nullptr )
  {
    if( threadIdx.x == 0 ) { slab_wait_flag( sync.wait_flag, sync.step, sync.err ); }
    __syncthreads();
  }
  const uint32_t blk = blockIdx.x - uint32_t( side ) * blocks_per_side;
  const uint32_t k = blk * blockDim.x + threadIdx.x;
  const uint32_t sent = *reinterpret_cast<const volatile uint32_t*>( &in[0].gid );
  const uint32_t count = sent < cap ? sent : cap;
  if( k == 0u )
  {
    ghost_counts[side] = count;
    if( sent > cap ) { ghost_counts[2] = 1u; } // more ghosts than reserved slots: reported by detect
  }
  if( blk * blockDim.x >= count ) { return; } // whole block past the list
  double mn[2] = { __longlong_as_double( 0x7ff0000000000000LL ), __longlong_as_double( 0x7ff0000000000000LL ) };
  double mx[2] = { __longlong_as_double( 0xfff0000000000000LL ), __longlong_as_double( 0xfff0000000000000LL ) };
  double ext = 0.0;
  if( k < count )
  {
    const int4* src = reinterpret_cast<const int4*>( &in[1u + k] );
    union { GhostRec g; int4 v[3]; } u;
    u.v[0] = src[0]; u.v[1] = src[1]; u.v[2] = src[2];
    const GhostRec& g = u.g;
    const uint32_t dst = args.slot[side] + k;
    q0[dst] = make_double2( g.q0x, g.q0y );
    q1[dst] = make_double2( g.q1x, g.q1y );
    r[dst] = g.r;
    gid[dst] = g.gid;
    double lo[2], hi[2];
    lo[0] = fmin( g.q1x, g.q0x ) - g.r; lo[1] = fmin( g.q1y, g.q0y ) - g.r;
    hi[0] = fmax( g.q1x, g.q0x ) + g.r; hi[1] = fmax( g.q1y, g.q0y ) + g.r;
    sg_bp_bounds_update<2>( lo, hi, mn, mx, ext );
  }
  if( acc != nullptr ) { sg_bp_bounds_commit<2>( mn, mx, ext, acc ); }
}



namespace
{
struct ByteSink
{
  unsigned char* p; uint64_t cap; uint64_t n;
  void put( const void* src, const uint64_t bytes ) { if( p != nullptr && n + bytes <= cap ) { memcpy( p + n, src, bytes ); } n += bytes; }
  template<typename T> void val( const T v ) { put( &v, sizeof( T ) ); }
  unsigned char* reserve( const uint64_t bytes ) { unsigned char* at = ( p != nullptr && n + bytes <= cap ) ? p + n : nullptr; n += bytes; return at; }
};
struct ByteSource
{
  const unsigned char* p; uint64_t cap; uint64_t n; bool ok;
  const unsigned char* take( const uint64_t bytes ) { if( !ok || n + bytes > cap ) { ok = false; return nullptr; } const unsigned char* at = p + n; n += bytes; return at; }
  template<typename T> T val() { T v{}; const unsigned char* at = take( sizeof( T ) ); if( at != nullptr ) { memcpy( &v, at, sizeof( T ) ); } return v; }
};
}

extern "C"
{

int sg_ball2d_set_bodies( sg_ctx* ctx, uint32_t n, const double* r, const double* m )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > 0 && ( r == nullptr || m == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_bodies: null array" ); }
  if( n >= 0x80000000u ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_bodies: at most 2^31 - 1 bodies" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  Ball2DData* d = ball2d_data( ctx );
  d->n = n;
  d->slab = false;
  d->have_result = false;
  d->flow_resident = false;
  if( n == 0 ) { return SG_OK; }
  SG_CUDA( ctx, d->r.ensure( size_t( n ) * 8 ) );
  SG_CUDA( ctx, d->m.ensure( size_t( n ) * 8 ) );
  SG_CUDA( ctx, d->q0.ensure( size_t( n ) * 16 ) );
  SG_CUDA( ctx, d->v0.ensure( size_t( n ) * 16 ) );
  SG_CUDA( ctx, d->q1.ensure( size_t( n ) * 16 ) );
  SG_CUDA( ctx, d->v1.ensure( size_t( n ) * 16 ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->r.ptr, r, size_t( n ) * 8, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->m.ptr, m, size_t( n ) * 8, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_ball2d_set_gravity( sg_ctx* ctx, const double* g )
{
  if( ctx == nullptr || g == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  d->g[0] = g[0]; d->g[1] = g[1];
  return SG_OK;
}

int sg_ball2d_set_planes( sg_ctx* ctx, uint32_t n, const double* x, const double* nrm )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > SG_MAX_PLANES ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_planes: at most %d planes", SG_MAX_PLANES ); }
  if( n > 0 && ( x == nullptr || nrm == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_planes: null array" ); }
  Ball2DData* d = ball2d_data( ctx );
  d->sg.nplanes = n;
  for( uint32_t p = 0; p < n; ++p )
  {
    // StaticPlane::StaticPlane: m_n( n.normalized() )  (Eigen 3.3: n / sqrt(n.n) when n.n > 0); host FP64, no contraction
    double nx = nrm[2 * p], ny = nrm[2 * p + 1];
    volatile double xx = nx * nx; volatile double yy = ny * ny;
    const double z = xx + yy;
    if( z > 0.0 ) { const double s = sqrt( z ); nx = nx / s; ny = ny / s; }
    d->sg.plane_x[p] = x[2 * p]; d->sg.plane_y[p] = x[2 * p + 1];
    d->sg.plane_nx[p] = nx; d->sg.plane_ny[p] = ny;
  }
  return SG_OK;
}

int sg_ball2d_set_drums( sg_ctx* ctx, uint32_t n, const double* x, const double* r )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > SG_MAX_DRUMS ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_drums: at most %d drums", SG_MAX_DRUMS ); }
  if( n > 0 && ( x == nullptr || r == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_drums: null array" ); }
  Ball2DData* d = ball2d_data( ctx );
  d->sg.ndrums = n;
  for( uint32_t k = 0; k < n; ++k ) { d->sg.drum_x[k] = x[2 * k]; d->sg.drum_y[k] = x[2 * k + 1]; d->sg.drum_r[k] = r[k]; }
  return SG_OK;
}

int sg_ball2d_set_portals( sg_ctx* ctx, uint32_t n, const double* plane_a_x, const double* plane_a_n, const double* plane_b_x, const double* plane_b_n, const double* v, const double* bounds )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > SG_MAX_PORTALS ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_portals: at most %d portals", SG_MAX_PORTALS ); }
  if( n > 0 && ( plane_a_x == nullptr || plane_a_n == nullptr || plane_b_x == nullptr || plane_b_n == nullptr || v == nullptr || bounds == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_portals: null array" ); }
  Ball2DData* d = ball2d_data( ctx );
  if( d->slab && n > 0 ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_set_portals: portals are not supported in slab mode" ); }
  if( n == 0 && d->px == nullptr ) { return SG_OK; }
  PortalData* x = ball2d_portal_data( d );
  SgPortals2D ps; // assembled on the side: a rejected call leaves the portals as they were
  memset( &ps, 0, sizeof( ps ) );
  ps.n = n;
  for( uint32_t p = 0; p < n; ++p )
  {
    SgPortal2D& pt = ps.p[p];
    for( int k = 0; k < 2; ++k ) { pt.ax[k] = plane_a_x[2 * p + k]; pt.bx[k] = plane_b_x[2 * p + k]; }
    // StaticPlane::StaticPlane (ball2d/StaticGeometry/StaticPlane.cpp:10-14)
    sg_portal_plane_frame( plane_a_n + 2 * p, pt.an, pt.at );
    sg_portal_plane_frame( plane_b_n + 2 * p, pt.bn, pt.bt );
    if( bounds[p] < 0.0 ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_set_portals: portal %u has negative bounds", p ); }
    pt.v = v[p]; pt.bounds = bounds[p]; pt.dx = 0.0; // PlanarPortal::PlanarPortal: m_dx( 0.0 )
  }
  x->portals = ps;
  d->have_result = false;
  return SG_OK;
}

int sg_ball2d_update_portals( sg_ctx* ctx, double t, double* dx_out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( d->px == nullptr ) { return SG_OK; }
  for( uint32_t p = 0; p < d->px->portals.n; ++p )
  {
    SgPortal2D& pt = d->px->portals.p[p];
    pt.dx = sg_portal_offset( pt.v, pt.bounds, t );
    if( dx_out != nullptr ) { dx_out[p] = pt.dx; }
  }
  return SG_OK;
}

int sg_ball2d_enforce_portals( sg_ctx* ctx, double* q, double* v )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( d->slab ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_enforce_portals: portals are not supported in slab mode" ); }
  if( !ball2d_has_portals( d ) || d->n == 0 ) { return SG_OK; }
  if( q == nullptr || v == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_enforce_portals: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  // q1 / v1 serve as the staging copies: whatever flow left there is overwritten, so residency ends here
  d->flow_resident = false;
  const size_t bytes = size_t( d->n ) * 16;
  SG_CUDA( ctx, cudaMemcpyAsync( d->q1.ptr, q, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->v1.ptr, v, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_LAUNCH( ctx, "b2p_enforce", double( d->n ) * 64.0, k_b2p_enforce<<<sg_div_up( d->n, 256 ), 256, 0, ctx->stream>>>( d->px->portals, d->n, d->q1.as<double2>(), d->v1.as<double2>() ) );
  SG_CUDA( ctx, cudaMemcpyAsync( q, d->q1.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( v, d->v1.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

int sg_ball2d_teleported( sg_ctx* ctx, sg_teleported* out )
{
  if( ctx == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  memset( out, 0, sizeof( *out ) );
  Ball2DData* d = ball2d_data( ctx );
  if( !d->have_result || !d->portal_result || d->px == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_teleported: the last active set was not computed with portals" ); }
  PortalData* x = d->px;
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  auto al = []( size_t b ) { return ( b + 63 ) & ~size_t( 63 ); };
  const size_t nb = x->n_boxes, nt = x->n_tel;
  size_t bytes = 64;
  const size_t o_bb = bytes; bytes += al( nb * 4 );
  const size_t o_bp = bytes; bytes += al( nb * 4 );
  const size_t o_p0 = bytes; bytes += al( nt * 4 );
  const size_t o_p1 = bytes; bytes += al( nt * 4 );
  const size_t o_x0 = bytes; bytes += al( nt * 16 );
  const size_t o_x1 = bytes; bytes += al( nt * 16 );
  const size_t o_k = bytes; bytes += al( nt * 16 );
  SG_CUDA( ctx, x->h_tele.ensure( bytes ) );
  char* h = x->h_tele.as<char>();
  if( nb > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_bb, x->box_body.ptr, nb * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_bp, x->box_portal.ptr, nb * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  }
  if( nt > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_p0, x->tp0.ptr, nt * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_p1, x->tp1.ptr, nt * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_x0, x->x0t.ptr, nt * 16, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_x1, x->x1t.ptr, nt * 16, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_k, x->kick.ptr, nt * 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  out->n_boxes = nb; out->n_regular = x->n_reg; out->n_teleported = nt;
  out->box_body = reinterpret_cast<const uint32_t*>( h + o_bb ); out->box_portal = reinterpret_cast<const uint32_t*>( h + o_bp );
  out->portal0 = reinterpret_cast<const uint32_t*>( h + o_p0 ); out->portal1 = reinterpret_cast<const uint32_t*>( h + o_p1 );
  out->x0 = reinterpret_cast<const double*>( h + o_x0 ); out->x1 = reinterpret_cast<const double*>( h + o_x1 ); out->kick = reinterpret_cast<const double*>( h + o_k );
  return SG_OK;
}

int sg_ball2d_flow( sg_ctx* ctx, int map_kind, const double* q0, const double* v0, double dt, double* q1, double* v1 )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( map_kind != SG_MAP_SYMPLECTIC_EULER && map_kind != SG_MAP_VERLET ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_flow: map kind %d is not a ball2d map", map_kind ); }
  Ball2DData* d = ball2d_data( ctx );
  if( d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_flow: context is in slab mode, use the sg_ball2d_slab_* calls" ); }
  if( d->n == 0 ) { return SG_OK; }
  if( q0 == nullptr || v0 == nullptr || q1 == nullptr || v1 == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_flow: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  // The map is per-body, so the call is pipelined over chunks of bodies on the context's two streams: while one chunk's
  // q1,v1 travel down, the next chunk's q0,v0 travel up (the two copy directions have their own engines and PCIe is
  // full duplex).  Small systems and timed (profile) runs take one chunk.
  const uint32_t n = d->n;
  const uint32_t nchunks = ( n >= ( 1u << 18 ) && !ctx->profile ) ? 8u : 1u;
  const uint32_t per = ( ( n + nchunks - 1u ) / nchunks + 255u ) & ~255u;
  if( nchunks > 1u )
  {
    SG_CUDA( ctx, cudaEventRecord( ctx->ev_fork, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamWaitEvent( ctx->stream2, ctx->ev_fork, 0 ) );
  }
  for( uint32_t c = 0; c < nchunks; ++c )
  {
    const uint32_t b0 = c * per;
    if( b0 >= n ) { break; }
    const uint32_t cnt = ( n - b0 < per ) ? n - b0 : per;
    cudaStream_t st = ( c & 1u ) ? ctx->stream2 : ctx->stream;
    const size_t off = size_t( b0 ) * 2, cb = size_t( cnt ) * 16;
    SG_CUDA( ctx, cudaMemcpyAsync( d->q0.as<double>() + off, q0 + off, cb, cudaMemcpyHostToDevice, st ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->v0.as<double>() + off, v0 + off, cb, cudaMemcpyHostToDevice, st ) );
    SG_LAUNCH( ctx, "ball2d_flow", double( cnt ) * 72.0, k_ball2d_flow<<<sg_div_up( cnt, 256 ), 256, 0, st>>>( map_kind, cnt, d->q0.as<double2>() + b0, d->v0.as<double2>() + b0, d->m.as<double>() + b0,
               d->g[0], d->g[1], dt, d->q1.as<double2>() + b0, d->v1.as<double2>() + b0 ) );
    SG_CUDA( ctx, cudaMemcpyAsync( q1 + off, d->q1.as<double>() + off, cb, cudaMemcpyDeviceToHost, st ) );
    SG_CUDA( ctx, cudaMemcpyAsync( v1 + off, d->v1.as<double>() + off, cb, cudaMemcpyDeviceToHost, st ) );
  }
  if( nchunks > 1u )
  {
    SG_CUDA( ctx, cudaEventRecord( ctx->ev_join, ctx->stream2 ) );
    SG_CUDA( ctx, cudaStreamWaitEvent( ctx->stream, ctx->ev_join, 0 ) );
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  d->flow_resident = true;
  return SG_OK;
}

int sg_ball2d_active_set( sg_ctx* ctx, const double* q0, const double* q1, uint32_t out_flags, sg_contacts* out )
{
  if( ctx == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_active_set: context is in slab mode, use the sg_ball2d_slab_* calls" ); }
  const size_t bytes = size_t( d->n ) * 16;
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( ( out_flags & SG_IN_RESIDENT ) != 0u )
  {
    // the caller vouches that (q0, q1) are the input and output of the last sg_ball2d_flow on this context: they are
    // still on the device, nothing is uploaded
    if( !d->flow_resident ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_active_set: SG_IN_RESIDENT without a preceding sg_ball2d_flow on this context" ); }
  }
  else if( d->n > 0 )
  {
    if( q0 == nullptr || q1 == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_active_set: null vector" ); }
    SG_CUDA( ctx, cudaMemcpyAsync( d->q0.ptr, q0, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->q1.ptr, q1, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    d->flow_resident = false;
  }
  const int rc = ball2d_active_set_device( ctx, d, ( out_flags & SG_OUT_CANDIDATES ) != 0u );
  if( rc != SG_OK ) { d->have_result = false; } // a failed call leaves nothing to fetch (partial or stale lists)
  if( rc != SG_OK ) { return rc; }
  return ball2d_copy_out( ctx, d, out_flags, out );
}

int sg_ball2d_upload( sg_ctx* ctx, const double* q, const double* v )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  d->flow_resident = false;
  const uint32_t nown = d->slab ? d->n_owned : d->n;
  if( nown == 0 ) { return SG_OK; }
  if( q == nullptr || v == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_upload: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->q0.as<double2>() + d->owned_slot(), q, size_t( nown ) * 16, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->v0.ptr, v, size_t( nown ) * 16, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_ball2d_slab_upload_q1( sg_ctx* ctx, const double* q1 )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_upload_q1: call sg_ball2d_slab_init first" ); }
  if( d->n_owned == 0 ) { return SG_OK; }
  if( q1 == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_upload_q1: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->q1.as<double2>() + d->owned_slot(), q1, size_t( d->n_owned ) * 16, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_ball2d_step( sg_ctx* ctx, int map_kind, double dt, sg_contacts* out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( map_kind != SG_MAP_SYMPLECTIC_EULER && map_kind != SG_MAP_VERLET ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_step: map kind %d is not a ball2d map", map_kind ); }
  Ball2DData* d = ball2d_data( ctx );
  d->flow_resident = false; // q1 is about to be overwritten by the resident step
  if( d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_step: context is in slab mode, use the sg_ball2d_slab_* calls" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const int rc = ball2d_active_set_device( ctx, d, true, map_kind, dt );
  if( rc != SG_OK ) { d->have_result = false; } // a failed call leaves nothing to fetch (partial or stale lists)
  if( rc != SG_OK ) { return rc; }
  if( out != nullptr )
  {
    memset( out, 0, sizeof( *out ) );
    out->dim = 2;
    out->n_candidates = d->n_cand;
    out->n_body_body = d->n_bb;
    out->n_active = d->n_bb + d->n_static;
  }
  return SG_OK;
}


int sg_ball2d_slab_init( sg_ctx* ctx, uint32_t n_owned, uint32_t gid_first, uint32_t ghost_cap, const double* r, const double* m )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n_owned > 0 && ( r == nullptr || m == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_init: null array" ); }
  if( uint64_t( n_owned ) + 2ull * ghost_cap >= 0x80000000ull ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_init: slab too large" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  Ball2DData* d = ball2d_data( ctx );
  if( ball2d_has_portals( d ) ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_slab_init: portals are not supported in slab mode" ); }
  d->slab = true; d->n_owned = n_owned; d->ghost_cap = ghost_cap; d->gid_first = gid_first;
  d->xlim[0] = -1.0e308; d->xlim[1] = 1.0e308;
  const size_t slots = size_t( n_owned ) + 2 * size_t( ghost_cap );
  d->n = uint32_t( slots );
  d->have_result = false;
  d->flow_resident = false;
  SG_CUDA( ctx, d->r.ensure( slots * 8 + 8 ) );
  SG_CUDA( ctx, d->q0.ensure( slots * 16 + 16 ) );
  SG_CUDA( ctx, d->q1.ensure( slots * 16 + 16 ) );
  SG_CUDA( ctx, d->gid.ensure( slots * 4 + 4 ) );
  SG_CUDA( ctx, d->m.ensure( size_t( n_owned ) * 8 + 8 ) );
  SG_CUDA( ctx, d->v0.ensure( size_t( n_owned ) * 16 + 16 ) );
  SG_CUDA( ctx, d->v1.ensure( size_t( n_owned ) * 16 + 16 ) );
  SG_CUDA( ctx, d->interval_enc.ensure( 16 ) );
  SG_CUDA( ctx, d->ghost_counts.ensure( 16 ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->ghost_counts.ptr, 0, 16, ctx->stream ) );
  k_slab_begin<<<1, 1, 0, ctx->stream>>>( d->interval_enc.as<long long>(), d->ghost_counts.as<uint32_t>() );
  if( n_owned > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( d->r.as<double>() + ghost_cap, r, size_t( n_owned ) * 8, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->m.ptr, m, size_t( n_owned ) * 8, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_LAUNCH( ctx, "slab_iota", 0.0, k_iota_u32<<<sg_div_up( n_owned, 256 ), 256, 0, ctx->stream>>>( n_owned, gid_first, d->gid.as<uint32_t>() + ghost_cap ) );
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_ball2d_slab_set_gids( sg_ctx* ctx, const uint32_t* gid_owned, const double* x_limits )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_set_gids: call sg_ball2d_slab_init first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( gid_owned != nullptr && d->n_owned > 0 )
  {
    // owned bodies are stored in ascending global order: their slot order is then the order of the emitted lists
    for( uint32_t k = 0; k < d->n_owned; ++k )
    {
      if( gid_owned[k] >= 0x80000000u || ( k > 0 && gid_owned[k] <= gid_owned[k - 1] ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_set_gids: global indices must be strictly ascending and below 2^31 (entry %u)", k ); }
    }
    SG_CUDA( ctx, cudaMemcpyAsync( d->gid.as<uint32_t>() + d->ghost_cap, gid_owned, size_t( d->n_owned ) * 4, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  }
  if( x_limits != nullptr )
  {
    if( !( x_limits[0] <= x_limits[1] ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_set_gids: empty x range" ); }
    d->xlim[0] = x_limits[0]; d->xlim[1] = x_limits[1];
  }
  SG_CUDA( ctx, cudaMemsetAsync( d->ghost_counts.ptr, 0, 16, ctx->stream ) ); // clears a pending re-partition request
  d->have_result = false;
  return SG_OK;
}

int sg_ball2d_slab_flow( sg_ctx* ctx, int map_kind, double dt, double* interval_dev )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( interval_dev == nullptr && d->mailbox.ptr == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_flow: interval_dev may only be null once a mailbox exists" ); }
  if( !d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_flow: call sg_ball2d_slab_init first" ); }
  if( map_kind != SG_MAP_SYMPLECTIC_EULER && map_kind != SG_MAP_VERLET && map_kind != SG_MAP_NONE ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_flow: map kind %d is not a ball2d map", map_kind ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  // One kernel over all slots: flow of the owned bodies, their share of the broad-phase bounds, the static-geometry
  // counts and [min lo.x, max hi.x] for the neighbours; then the interval is decoded / posted.
  int rc = sg_bp_prepare_scratch<Ball2DPolicy>( ctx, d->bp, d->n );
  if( rc != SG_OK ) { return rc; }
  rc = ball2d_static_scratch( ctx, d );
  if( rc != SG_OK ) { return rc; }
  const uint32_t n = d->n;
  SG_CUDA( ctx, d->block_iv.ensure( size_t( sg_div_up( n > 0 ? n : 1, 256 ) ) * 16 + 16 ) );
  const unsigned nblk_prep = sg_div_up( n > 0 ? n : 1, 256 );
  const SlabCand sc = ball2d_slab_cand( d );
  if( sc.band != nullptr ) { SG_CUDA( ctx, cudaMemsetAsync( sc.count, 0, 8, ctx->stream ) ); } // a flow that was never followed by an exchange must not leave candidates behind
  if( map_kind == SG_MAP_NONE )
  {
    // q1 of the owned bodies was uploaded (sg_ball2d_slab_upload_q1): everything but the integration
    SG_LAUNCH( ctx, "slab_prep", double( d->n_owned ) * 40.0,
               k_ball2d_prep<false><<<nblk_prep, 256, 0, ctx->stream>>>( d->sg, 0, n, d->Q0(), nullptr, nullptr, d->R(), 0.0, 0.0, 0.0, d->Q1(), nullptr, d->bp.bounds_cur(), d->st_counts.as<uint32_t>(),
                                                                       d->own_first(), d->own_count(), nullptr, d->interval_enc.as<long long>(), d->block_iv.as<double2>(), d->xlim[0], d->xlim[1], d->ghost_counts.as<uint32_t>(), sc ) );
  }
  else
  {
    SG_LAUNCH( ctx, "slab_flow_prep", double( d->n_owned ) * 80.0,
               k_ball2d_prep<true><<<nblk_prep, 256, 0, ctx->stream>>>( d->sg, map_kind, n, d->Q0(), d->v0.as<double2>(), d->m.as<double>(), d->R(), d->g[0], d->g[1], dt, d->Q1(), d->v1.as<double2>(),
                                                                      d->bp.bounds_cur(), d->st_counts.as<uint32_t>(), d->own_first(), d->own_count(), nullptr, d->interval_enc.as<long long>(),
                                                                      d->block_iv.as<double2>(), d->xlim[0], d->xlim[1], d->ghost_counts.as<uint32_t>(), sc ) );
  }
  if( d->mailbox.ptr == nullptr )
  {
    SG_LAUNCH( ctx, "slab_interval", 16.0, k_slab_interval_decode<<<1, 1, 0, ctx->stream>>>( d->interval_enc.as<long long>(), d->ghost_counts.as<uint32_t>(), interval_dev ) );
  }
  else
  {
    ++d->slab_step;
    SG_LAUNCH( ctx, "slab_interval", 16.0, k_slab_post_interval<<<1, 1, 0, ctx->stream>>>( d->interval_enc.as<long long>(), d->ghost_counts.as<uint32_t>(), interval_dev, static_cast<SlabMailboxHdr*>( d->peer_mb[0] ),
               static_cast<SlabMailboxHdr*>( d->peer_mb[1] ), d->slab_step ) );
  }
  d->slab_prep_done = true;
  return SG_OK;
}

// One or two targets (sides) per call: target t packs against interval_dev[t] into send_dev[t] (nullptr: count only)
static int ball2d_slab_pack_impl( sg_ctx* ctx, Ball2DData* d, const int ntargets, const double* const* interval_dev, void* const* send_dev, const uint32_t cap, uint32_t* const* count_dev, const SlabSync* sync )
{
  const uint32_t n = d->n; // all slots: same block partition as the flow kernel
  const unsigned nblk = sg_div_up( n > 0 ? n : 1, 256 );
  SG_CUDA( ctx, d->pack_counts.ensure( 2 * ( size_t( nblk ) * 4 + 4 ) ) );
  // block intervals are valid when this step's flow went through sg_ball2d_slab_flow
  const double2* biv = ( d->slab_prep_done && d->block_iv.ptr != nullptr ) ? d->block_iv.as<double2>() : nullptr;
  PackArgs count_args, emit_args;
  bool any_emit = false;
  for( int t = 0; t < 2; ++t )
  {
    const bool on = t < ntargets;
    count_args.on[t] = on; emit_args.on[t] = on && send_dev[t] != nullptr;
    count_args.iv[t] = emit_args.iv[t] = on ? interval_dev[t] : nullptr;
    count_args.block_counts[t] = emit_args.block_counts[t] = d->pack_counts.as<uint32_t>() + size_t( t ) * ( nblk + 1 );
    count_args.total[t] = nullptr; emit_args.total[t] = on ? count_dev[t] : nullptr;
    count_args.out[t] = nullptr; emit_args.out[t] = on ? static_cast<GhostRec*>( send_dev[t] ) : nullptr;
    SlabSync none; none.wait_flag = nullptr; none.post_flag = nullptr; none.done_ctr = nullptr; none.err = nullptr; none.step = 0u;
    count_args.sync[t] = on ? sync[t] : none; count_args.sync[t].post_flag = nullptr;
    emit_args.sync[t] = on ? sync[t] : none; emit_args.sync[t].wait_flag = nullptr;
    any_emit = any_emit || emit_args.on[t];
  }
  SG_LAUNCH( ctx, "slab_pack_count", double( nblk ) * 16.0, k_ball2d_slab_pack<false><<<nblk, 256, 0, ctx->stream>>>( n, d->own_first(), d->own_count(), d->q0.as<double2>(), d->q1.as<double2>(), d->r.as<double>(),
             d->gid.as<uint32_t>(), biv, cap, count_args ) );
  if( any_emit )
  {
    SG_LAUNCH( ctx, "slab_pack_emit", double( nblk ) * 16.0, k_ball2d_slab_pack<true><<<nblk, 256, 0, ctx->stream>>>( n, d->own_first(), d->own_count(), d->q0.as<double2>(), d->q1.as<double2>(), d->r.as<double>(),
               d->gid.as<uint32_t>(), biv, cap, emit_args ) );
  }
  for( int t = 0; t < ntargets; ++t )
  {
    if( send_dev[t] == nullptr && count_dev[t] != nullptr )
    {
      SG_LAUNCH( ctx, "slab_pack_total", double( nblk ) * 4.0, k_ball2d_slab_pack_total<<<1, 256, 0, ctx->stream>>>( nblk, count_args.block_counts[t], count_dev[t] ) );
    }
  }
  return SG_OK;
}

int sg_ball2d_slab_pack( sg_ctx* ctx, const double* interval_dev, void* send_dev, uint32_t cap, uint32_t* count_dev )
{
  if( ctx == nullptr || interval_dev == nullptr || count_dev == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_pack: call sg_ball2d_slab_init first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SlabSync none; none.wait_flag = nullptr; none.post_flag = nullptr; none.done_ctr = nullptr; none.err = nullptr; none.step = 0u;
  const double* ivs[1] = { interval_dev };
  void* sends[1] = { send_dev };
  uint32_t* counts[1] = { count_dev };
  return ball2d_slab_pack_impl( ctx, d, 1, ivs, sends, cap, counts, &none );
}

// recv_dev[s] != nullptr: unpack side s (both sides go in one launch)
static int ball2d_slab_unpack_impl( sg_ctx* ctx, Ball2DData* d, const void* const* recv_dev, const SlabSync* sync )
{
  const uint32_t cap = d->ghost_cap;
  const unsigned bps = sg_div_up( cap > 0 ? cap : 1, 256 );
  UnpackArgs args;
  for( int sde = 0; sde < 2; ++sde )
  {
    args.on[sde] = recv_dev[sde] != nullptr;
    args.in[sde] = static_cast<const GhostRec*>( recv_dev[sde] );
    // side 0: ghosts with smaller global indices fill slots [0, count); side 1: the slots right after the owned block
    args.slot[sde] = ( sde == 0 ) ? 0u : d->ghost_cap + d->n_owned;
    args.sync[sde] = sync[sde];
  }
  // the ghosts' boxes join the bounds the flow kernel started (when this step went through sg_ball2d_slab_flow)
  BoundsAccum* acc = ( d->slab_prep_done && d->bp.bounds.ptr != nullptr ) ? d->bp.bounds_cur() : nullptr;
  SG_LAUNCH( ctx, "slab_unpack", double( cap ) * 4.0, k_ball2d_slab_unpack<<<2 * bps, 256, 0, ctx->stream>>>( cap, bps, args, d->q0.as<double2>(), d->q1.as<double2>(), d->r.as<double>(), d->gid.as<uint32_t>(),
             d->ghost_counts.as<uint32_t>(), acc ) );
  return SG_OK;
}

int sg_ball2d_slab_unpack( sg_ctx* ctx, int side, const void* recv_dev )
{
  if( ctx == nullptr || recv_dev == nullptr || ( side != 0 && side != 1 ) ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_unpack: call sg_ball2d_slab_init first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SlabSync none; none.wait_flag = nullptr; none.post_flag = nullptr; none.done_ctr = nullptr; none.err = nullptr; none.step = 0u;
  const void* recv[2] = { side == 0 ? recv_dev : nullptr, side == 1 ? recv_dev : nullptr };
  const SlabSync syncs[2] = { none, none };
  return ball2d_slab_unpack_impl( ctx, d, recv, syncs );
}

int sg_ball2d_slab_mailbox( sg_ctx* ctx, void** mailbox_dev, void* ipc_handle_64 )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_mailbox: call sg_ball2d_slab_init first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  static_assert( sizeof( cudaIpcMemHandle_t ) == 64, "the ABI passes IPC handles as 64 opaque bytes" );
  if( d->mailbox.ptr == nullptr )
  {
    const size_t bytes = sizeof( SlabMailboxHdr ) + 2 * ( size_t( d->ghost_cap ) + 1 ) * sizeof( GhostRec );
    SG_CUDA( ctx, d->mailbox.ensure( bytes ) );
    SG_CUDA( ctx, cudaMemsetAsync( d->mailbox.ptr, 0, bytes, ctx->stream ) );
    SG_CUDA( ctx, d->pack_total.ensure( 16 ) );
    SG_CUDA( ctx, d->pack_done.ensure( 16 ) );
    SG_CUDA( ctx, cudaMemsetAsync( d->pack_done.ptr, 0, 16, ctx->stream ) );
    d->cand_cap = 4u * d->ghost_cap + 1024u;
    SG_CUDA( ctx, d->cand_state.ensure( sizeof( SlabCandState ) ) );
    SG_CUDA( ctx, d->cand_list.ensure( 2 * size_t( d->cand_cap ) * 4 ) );
    k_slab_cand_reset<<<1, 1, 0, ctx->stream>>>( d->cand_state.as<SlabCandState>() );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    d->slab_step = 0;
  }
  if( mailbox_dev != nullptr ) { *mailbox_dev = d->mailbox.ptr; }
  if( ipc_handle_64 != nullptr )
  {
    cudaIpcMemHandle_t h;
    SG_CUDA( ctx, cudaIpcGetMemHandle( &h, d->mailbox.ptr ) );
    memcpy( ipc_handle_64, &h, 64 );
  }
  return SG_OK;
}

int sg_ball2d_slab_connect( sg_ctx* ctx, int side, const void* ipc_handle_64, void* same_process_mailbox, int peer_device )
{
  if( ctx == nullptr || ( side != 0 && side != 1 ) || ( ipc_handle_64 == nullptr ) == ( same_process_mailbox == nullptr ) ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->slab || d->mailbox.ptr == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_connect: create this rank's mailbox first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( d->peer_mb[side] != nullptr && d->peer_ipc[side] ) { cudaIpcCloseMemHandle( d->peer_mb[side] ); }
  d->peer_mb[side] = nullptr;
  if( ipc_handle_64 != nullptr )
  {
    cudaIpcMemHandle_t h;
    memcpy( &h, ipc_handle_64, 64 );
    void* p = nullptr;
    SG_CUDA( ctx, cudaIpcOpenMemHandle( &p, h, cudaIpcMemLazyEnablePeerAccess ) );
    d->peer_mb[side] = p; d->peer_ipc[side] = true;
  }
  else
  {
    if( peer_device >= 0 && peer_device != ctx->device )
    {
      int can = 0;
      SG_CUDA( ctx, cudaDeviceCanAccessPeer( &can, ctx->device, peer_device ) );
      if( !can ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "device %d cannot map the memory of device %d", ctx->device, peer_device ); }
      const cudaError_t e = cudaDeviceEnablePeerAccess( peer_device, 0 );
      if( e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled ) { return sg_fail( ctx, SG_ERR_CUDA, "cudaDeviceEnablePeerAccess( %d ): %s", peer_device, cudaGetErrorString( e ) ); }
      cudaGetLastError();
    }
    d->peer_mb[side] = same_process_mailbox; d->peer_ipc[side] = false;
  }
  return SG_OK;
}

int sg_ball2d_slab_disconnect( sg_ctx* ctx )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  for( int sde = 0; sde < 2; ++sde )
  {
    if( d->peer_mb[sde] != nullptr && d->peer_ipc[sde] ) { cudaIpcCloseMemHandle( d->peer_mb[sde] ); }
    d->peer_mb[sde] = nullptr; d->peer_ipc[sde] = false;
  }
  d->mailbox.release();
  cudaGetLastError();
  return SG_OK;
}

// phase 1: for each connected neighbour wait for its interval, pack the owned bodies that reach it straight into
//          its mailbox, raise its halo flag;  phase 2: wait for the neighbours' halos and move them into the ghost
//          slots;  phase 0: both.  (A driver with several ranks in ONE process must run phase 1 on every rank
//          before phase 2 on any, because a wait kernel only returns once the neighbour's work has been launched.)
int sg_ball2d_slab_exchange( sg_ctx* ctx, int phase )
{
  if( ctx == nullptr || phase < 0 || phase > 2 ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->slab || d->mailbox.ptr == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_exchange: no mailbox" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SlabMailboxHdr* mine = d->mailbox.as<SlabMailboxHdr>();
  const uint32_t step = d->slab_step;
  if( phase == 0 || phase == 1 )
  {
    // both neighbours in one launch: the candidates the flow kernel listed (or, when a band did not hold, all owned bodies) against
    // the neighbours' intervals, selected records straight into their mailboxes, flags raised by the last block
    Pack2Args<GhostRec> pa;
    bool any = false;
    for( int side = 0; side < 2; ++side )
    {
      pa.on[side] = d->peer_mb[side] != nullptr;
      pa.iv[side] = &mine->iv[side][0]; pa.wait[side] = &mine->iv_flag[side]; pa.out[side] = nullptr; pa.post[side] = nullptr;
      if( !pa.on[side] ) { continue; }
      SlabMailboxHdr* peer = static_cast<SlabMailboxHdr*>( d->peer_mb[side] );
      // seen from the neighbour on `side`, this rank sits on its side 1 - side
      pa.out[side] = slab_mailbox_halo<GhostRec>( peer, 1 - side, d->ghost_cap );
      pa.post[side] = &peer->halo_flag[1 - side];
      any = true;
    }
    pa.err = &mine->err; pa.step = step;
    if( any )
    {
      if( !d->slab_prep_done ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_exchange: call sg_ball2d_slab_flow first" ); }
      const SlabCand sc = ball2d_slab_cand( d );
      const unsigned grid = unsigned( ctx->num_sms ); // one block per SM: the candidate lists are a few thousand entries
      Ball2DSlabTraits::Src src;
      src.q0 = d->q0.as<double2>(); src.q1 = d->q1.as<double2>(); src.r = d->r.as<double>(); src.gid = d->gid.as<uint32_t>();
      SG_LAUNCH( ctx, "slab_pack", double( d->ghost_cap ) * 2.0 * 48.0, k_slab_pack2<Ball2DSlabTraits><<<grid, 256, 0, ctx->stream>>>( d->own_first(), d->own_count(), src, d->ghost_cap, sc, d->cand_state.as<SlabCandState>(),
                 d->bp.params.as<GridParams>(), pa ) );
    }
  }
  if( phase == 0 || phase == 2 )
  {
    const void* recv[2] = { nullptr, nullptr };
    SlabSync syncs[2];
    bool any = false;
    for( int side = 0; side < 2; ++side )
    {
      syncs[side].wait_flag = nullptr; syncs[side].post_flag = nullptr; syncs[side].done_ctr = nullptr; syncs[side].err = &mine->err; syncs[side].step = step;
      if( d->peer_mb[side] == nullptr ) { continue; }
      recv[side] = slab_mailbox_halo<GhostRec>( mine, side, d->ghost_cap );
      syncs[side].wait_flag = &mine->halo_flag[side];
      any = true;
    }
    if( any )
    {
      const int rc = ball2d_slab_unpack_impl( ctx, d, recv, syncs );
      if( rc != SG_OK ) { return rc; }
    }
  }
  return SG_OK;
}

int sg_ball2d_slab_detect( sg_ctx* ctx, sg_contacts* out, uint32_t* ghosts_out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->slab ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_slab_detect: call sg_ball2d_slab_init first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const int rc = ball2d_active_set_device( ctx, d, true );
  if( rc != SG_OK ) { d->have_result = false; } // a failed call leaves nothing to fetch (partial or stale lists)
  if( rc != SG_OK ) { return rc; }
  uint32_t* hg = reinterpret_cast<uint32_t*>( d->h_totals.as<unsigned long long>() + 4 );
  SG_CUDA( ctx, cudaMemcpyAsync( hg, d->ghost_counts.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  if( hg[2] != 0u ) { return sg_fail( ctx, SG_ERR_INVALID, "slab halo exceeds the reserved ghost capacity %u: results are incomplete", d->ghost_cap ); }
  if( hg[3] != 0u ) { return sg_fail( ctx, SG_ERR_REBALANCE, "a body of this slab left [%g, %g]: it may reach a non-neighbouring slab, re-partition the scene", d->xlim[0], d->xlim[1] ); }
  if( d->mailbox.ptr != nullptr )
  {
    uint32_t err = 0u;
    SG_CUDA( ctx, cudaMemcpy( &err, &d->mailbox.as<SlabMailboxHdr>()->err, 4, cudaMemcpyDeviceToHost ) );
    if( err != 0u ) { return sg_fail( ctx, SG_ERR_INTERNAL, "slab exchange: a neighbour did not post its interval or halo within 10 s" ); }
  }
  if( ghosts_out != nullptr ) { ghosts_out[0] = hg[0]; ghosts_out[1] = hg[1]; }
  if( out != nullptr )
  {
    memset( out, 0, sizeof( *out ) );
    out->dim = 2;
    out->n_candidates = d->n_cand;
    out->n_body_body = d->n_bb;
    out->n_active = d->n_bb + d->n_static;
  }
  return SG_OK;
}

int sg_ball2d_slab_stats( sg_ctx* ctx, uint32_t* out4 )
{
  if( ctx == nullptr || out4 == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  out4[0] = out4[1] = out4[2] = out4[3] = 0u;
  if( !d->slab ) { return SG_OK; }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  uint32_t g[4];
  SG_CUDA( ctx, cudaMemcpy( g, d->ghost_counts.ptr, 16, cudaMemcpyDeviceToHost ) );
  out4[0] = g[0]; out4[1] = g[1];
  if( d->cand_state.ptr != nullptr && d->mailbox.ptr != nullptr )
  {
    SlabCandState st;
    SG_CUDA( ctx, cudaMemcpy( &st, d->cand_state.ptr, sizeof( st ), cudaMemcpyDeviceToHost ) );
    out4[2] = st.fallbacks;
    out4[3] = d->cand_cap;
  }
  return SG_OK;
}

int sg_ball2d_fetch_state( sg_ctx* ctx, double* q1, double* v1 )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const uint32_t nown = d->slab ? d->n_owned : d->n;
  if( q1 != nullptr && nown > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( q1, d->q1.as<double2>() + d->owned_slot(), size_t( nown ) * 16, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  if( v1 != nullptr && nown > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( v1, d->v1.ptr, size_t( nown ) * 16, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

int sg_ball2d_fetch( sg_ctx* ctx, uint32_t out_flags, double* q1, double* v1, sg_contacts* out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->have_result ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_fetch: no step has been run" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const uint32_t nown = d->slab ? d->n_owned : d->n;
  if( q1 != nullptr && nown > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( q1, d->q1.as<double2>() + d->owned_slot(), size_t( nown ) * 16, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  if( v1 != nullptr && nown > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( v1, d->v1.ptr, size_t( nown ) * 16, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  if( out != nullptr ) { return ball2d_copy_out( ctx, d, out_flags, out ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}


// ---- state I/O at the seam: Ball2DState's binary snapshot --------------------------------------------
// The byte stream Ball2DState::serialize writes (ball2d/Ball2DState.cpp:259-272) with the helpers of scisim/Utilities.h:43-94 and
// scisim/Math/MathUtilities.h:42-60, MathUtilities.cpp:142-154 -- raw little-endian values, no padding:
//   q, v, r            Eigen::Index rows (int64) + doubles
//   fixed              size_t count + one byte per ball
//   M, Minv            Index rows, cols, nnz; int inner[nnz]; int outer[cols + 1]; double values[nnz]   (diagonal: m resp. 1 / m per DoF)
//   drums              size_t count + { x (2 doubles), r }
//   planes             size_t count + { x, v, n, t } (2 doubles each; v = 0, n unit, t = ( -n.y, n.x ))
//   portals            size_t count + { plane A, plane B, v, bounds, dx }
//   forces             size_t count + { size_t length + "ball2d_gravity_force", g (2 doubles) }
// so that a device-resident state can be checkpointed in the reference's own format, and a reference snapshot resumed on the GPU.
// which = 0: ( q0, v0 ) as uploaded; 1: ( q1, v1 ) as the last flow / step left them.  buf = NULL: size query.
int sg_ball2d_state_serialize( sg_ctx* ctx, int which, void* buf, uint64_t cap, uint64_t* bytes )
{
  if( ctx == nullptr || bytes == nullptr || ( which != 0 && which != 1 ) ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( d->slab ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_state_serialize: a slab holds part of a scene; serialise through the owner of the whole state" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const uint64_t n = d->n;
  ByteSink out{ static_cast<unsigned char*>( buf ), cap, 0 };
  const long long dofs = ( long long )( 2 * n );
  std::vector<double> mass( n );
  if( n > 0 && buf != nullptr ) { SG_CUDA( ctx, cudaMemcpyAsync( mass.data(), d->m.ptr, n * 8, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  // q, v straight from the device into the stream
  for( int k = 0; k < 2; ++k )
  {
    out.val<long long>( dofs );
    unsigned char* at = out.reserve( n * 16 );
    const void* src = ( k == 0 ) ? ( which == 0 ? d->q0.ptr : d->q1.ptr ) : ( which == 0 ? d->v0.ptr : d->v1.ptr );
    if( at != nullptr && n > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( at, src, n * 16, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  }
  out.val<long long>( ( long long )( n ) );
  {
    unsigned char* at = out.reserve( n * 8 );
    if( at != nullptr && n > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( at, d->r.ptr, n * 8, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  out.val<size_t>( size_t( n ) );
  for( uint64_t i = 0; i < n; ++i ) { out.val<unsigned char>( 0 ); } // m_fixed: stored by the reference, never read on this path
  for( int inv = 0; inv < 2; ++inv )
  {
    out.val<long long>( dofs ); out.val<long long>( dofs ); out.val<long long>( dofs );
    for( long long k = 0; k < dofs; ++k ) { out.val<int>( int( k ) ); }
    for( long long k = 0; k <= dofs; ++k ) { out.val<int>( int( k ) ); }
    for( uint64_t i = 0; i < n; ++i ) { const double v = inv ? 1.0 / mass[i] : mass[i]; out.val<double>( v ); out.val<double>( v ); } // createMinv: 1.0 / m (Ball2DState.cpp:54-66)
  }
  out.val<size_t>( size_t( d->sg.ndrums ) );
  for( uint32_t k = 0; k < d->sg.ndrums; ++k ) { out.val<double>( d->sg.drum_x[k] ); out.val<double>( d->sg.drum_y[k] ); out.val<double>( d->sg.drum_r[k] ); }
  out.val<size_t>( size_t( d->sg.nplanes ) );
  for( uint32_t k = 0; k < d->sg.nplanes; ++k )
  {
    out.val<double>( d->sg.plane_x[k] ); out.val<double>( d->sg.plane_y[k] ); out.val<double>( 0.0 ); out.val<double>( 0.0 );
    out.val<double>( d->sg.plane_nx[k] ); out.val<double>( d->sg.plane_ny[k] ); out.val<double>( -d->sg.plane_ny[k] ); out.val<double>( d->sg.plane_nx[k] );
  }
  const uint32_t np = ( d->px != nullptr ) ? d->px->portals.n : 0u;
  out.val<size_t>( size_t( np ) );
  for( uint32_t k = 0; k < np; ++k )
  {
    const SgPortal2D& pt = d->px->portals.p[k];
    out.put( pt.ax, 16 ); out.val<double>( 0.0 ); out.val<double>( 0.0 ); out.put( pt.an, 16 ); out.put( pt.at, 16 );
    out.put( pt.bx, 16 ); out.val<double>( 0.0 ); out.val<double>( 0.0 ); out.put( pt.bn, 16 ); out.put( pt.bt, 16 );
    out.val<double>( pt.v ); out.val<double>( pt.bounds ); out.val<double>( pt.dx );
  }
  out.val<size_t>( size_t( 1 ) );
  const char name[] = "ball2d_gravity_force";
  out.val<size_t>( sizeof( name ) - 1 ); out.put( name, sizeof( name ) - 1 );
  out.val<double>( d->g[0] ); out.val<double>( d->g[1] );
  *bytes = out.n;
  if( buf != nullptr && out.n > cap ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_serialize: buffer of %llu bytes, %llu needed", ( unsigned long long )( cap ), ( unsigned long long )( out.n ) ); }
  return SG_OK;
}

// Ball2DState::deserialize (ball2d/Ball2DState.cpp:274-312): configures the context from a snapshot and uploads ( q, v ).
int sg_ball2d_state_deserialize( sg_ctx* ctx, const void* buf, uint64_t bytes )
{
  if( ctx == nullptr || buf == nullptr ) { return SG_ERR_INVALID; }
  ByteSource in{ static_cast<const unsigned char*>( buf ), bytes, 0, true };
  const long long nq = in.val<long long>();
  if( !in.ok || nq < 0 || ( nq & 1 ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_deserialize: bad q header" ); }
  const uint64_t n = uint64_t( nq ) / 2;
  const unsigned char* q = in.take( n * 16 );
  const long long nv = in.val<long long>();
  const unsigned char* v = in.take( n * 16 );
  const long long nr = in.val<long long>();
  const unsigned char* r = in.take( n * 8 );
  const size_t nfixed = in.val<size_t>();
  in.take( nfixed );
  if( !in.ok || nv != nq || uint64_t( nr ) != n || nfixed != n ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_deserialize: inconsistent vector sizes" ); }
  std::vector<double> mass( n );
  for( int inv = 0; inv < 2; ++inv )
  {
    const long long rows = in.val<long long>(), cols = in.val<long long>(), nnz = in.val<long long>();
    if( !in.ok || rows != nq || cols != nq || nnz != nq ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_deserialize: the mass matrix is not the 2N diagonal" ); }
    in.take( uint64_t( nnz ) * 4 ); in.take( uint64_t( cols + 1 ) * 4 );
    const unsigned char* vals = in.take( uint64_t( nnz ) * 8 );
    if( inv == 0 && vals != nullptr ) { for( uint64_t i = 0; i < n; ++i ) { memcpy( &mass[i], vals + 16 * i, 8 ); } }
  }
  if( !in.ok ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_deserialize: truncated snapshot" ); }
  std::vector<double> rr( n );
  if( n > 0 ) { memcpy( rr.data(), r, n * 8 ); }
  int rc = sg_ball2d_set_bodies( ctx, uint32_t( n ), rr.data(), mass.data() );
  if( rc != SG_OK ) { return rc; }
  Ball2DData* d = ball2d_data( ctx );
  const size_t nd = in.val<size_t>();
  if( !in.ok || nd > SG_MAX_DRUMS ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_deserialize: bad drum count" ); }
  d->sg.ndrums = uint32_t( nd );
  for( size_t k = 0; k < nd; ++k ) { d->sg.drum_x[k] = in.val<double>(); d->sg.drum_y[k] = in.val<double>(); d->sg.drum_r[k] = in.val<double>(); }
  const size_t npl = in.val<size_t>();
  if( !in.ok || npl > SG_MAX_PLANES ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_deserialize: bad plane count" ); }
  d->sg.nplanes = uint32_t( npl );
  for( size_t k = 0; k < npl; ++k )
  {
    // StaticPlane( std::istream& ): x, v, n, t read back as stored -- the normal is NOT normalised again (StaticPlane.cpp:22-32)
    d->sg.plane_x[k] = in.val<double>(); d->sg.plane_y[k] = in.val<double>(); in.val<double>(); in.val<double>();
    d->sg.plane_nx[k] = in.val<double>(); d->sg.plane_ny[k] = in.val<double>(); in.val<double>(); in.val<double>();
  }
  const size_t npo = in.val<size_t>();
  if( !in.ok || npo > SG_MAX_PORTALS ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_deserialize: bad portal count" ); }
  if( npo > 0 || d->px != nullptr )
  {
    PortalData* x = ball2d_portal_data( d );
    memset( &x->portals, 0, sizeof( x->portals ) );
    x->portals.n = uint32_t( npo );
    for( size_t k = 0; k < npo; ++k )
    {
      SgPortal2D& pt = x->portals.p[k];
      pt.ax[0] = in.val<double>(); pt.ax[1] = in.val<double>(); in.val<double>(); in.val<double>(); pt.an[0] = in.val<double>(); pt.an[1] = in.val<double>(); pt.at[0] = in.val<double>(); pt.at[1] = in.val<double>();
      pt.bx[0] = in.val<double>(); pt.bx[1] = in.val<double>(); in.val<double>(); in.val<double>(); pt.bn[0] = in.val<double>(); pt.bn[1] = in.val<double>(); pt.bt[0] = in.val<double>(); pt.bt[1] = in.val<double>();
      pt.v = in.val<double>(); pt.bounds = in.val<double>(); pt.dx = in.val<double>();
    }
  }
  const size_t nf = in.val<size_t>();
  d->g[0] = 0.0; d->g[1] = 0.0;
  for( size_t k = 0; k < nf && in.ok; ++k )
  {
    const size_t len = in.val<size_t>();
    const unsigned char* nm = in.take( len );
    if( nm == nullptr || len != 20 || memcmp( nm, "ball2d_gravity_force", 20 ) != 0 ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_state_deserialize: a force other than ball2d_gravity_force (the reference exits too: Ball2DState.cpp:302-310)" ); }
    // forces accumulate (Ball2DState::accumulateForce): several gravity forces add up
    d->g[0] += in.val<double>(); d->g[1] += in.val<double>();
  }
  if( !in.ok ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_state_deserialize: truncated snapshot" ); }
  if( n == 0 ) { return SG_OK; }
  // q, v may sit unaligned in the stream: through aligned copies
  std::vector<double> qq( 2 * n ), vv( 2 * n );
  memcpy( qq.data(), q, n * 16 ); memcpy( vv.data(), v, n * 16 );
  return sg_ball2d_upload( ctx, qq.data(), vv.data() );
}


// ---- device-side assembly for the solver's first step (sg_ball2d_assembly.cuh) ----------------------------------------
static ContactOut2D ball2d_contacts_view( const Ball2DData* d )
{
  ContactOut2D out;
  out.type = d->c_type.as<uint32_t>(); out.i = d->c_i.as<uint32_t>(); out.j = d->c_j.as<uint32_t>();
  out.n = d->c_n.as<double2>(); out.p = d->c_p.as<double2>(); out.depth = d->c_depth.as<double>();
  out.cap = d->act_cap;
  return out;
}

int sg_ball2d_assemble( sg_ctx* ctx, uint32_t flags, sg_assembly* out )
{
  if( ctx == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  memset( out, 0, sizeof( *out ) );
  Ball2DData* d = ball2d_data( ctx );
  if( !d->have_result ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_assemble: no active set has been computed" ); }
  if( d->slab || d->portal_result ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_assemble: slab / portal active sets are not supported" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( d->asmd == nullptr ) { d->asmd = new AsmData; }
  AsmData& A = *d->asmd;
  const uint64_t nc = d->n_bb + d->n_static;
  const uint32_t n = d->n;
  out->n_constraints = nc; out->n_dofs = 2ull * n;
  if( nc >= 0x7fffffffull / 4ull ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_assemble: more constraints than 32-bit sparse indices hold" ); }
  const ContactOut2D a = ball2d_contacts_view( d );
  SG_CUDA( ctx, A.ncnt.ensure( ( nc + 2 ) * 4 ) ); SG_CUDA( ctx, A.nouter.ensure( ( nc + 2 ) * 4 ) );
  SG_CUDA( ctx, A.ninner.ensure( ( 4 * nc + 4 ) * 4 ) ); SG_CUDA( ctx, A.nval.ensure( ( 4 * nc + 4 ) * 8 ) );
  SG_CUDA( ctx, A.deg.ensure( ( size_t( n ) + 2 ) * 4 ) ); SG_CUDA( ctx, A.inc_start.ensure( ( size_t( n ) + 2 ) * 4 ) ); SG_CUDA( ctx, A.inc_cursor.ensure( ( size_t( n ) + 2 ) * 4 ) );
  SG_CUDA( ctx, A.inc.ensure( ( 2 * nc + 4 ) * 4 ) );
  SG_CUDA( ctx, A.qcnt.ensure( ( nc + 2 ) * 4 ) ); SG_CUDA( ctx, A.qouter.ensure( ( nc + 2 ) * 4 ) );
  SG_CUDA( ctx, A.bases.ensure( ( 4 * nc + 4 ) * 8 ) );
  SG_CUDA( ctx, A.total.ensure( 16 ) ); SG_CUDA( ctx, A.bad.ensure( 4 ) );
  SG_CUDA( ctx, cudaMemsetAsync( A.deg.ptr, 0, ( size_t( n ) + 2 ) * 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( A.inc_cursor.ptr, 0, ( size_t( n ) + 2 ) * 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( A.bad.ptr, 0, 4, ctx->stream ) );
  uint32_t h_tot[2] = { 0u, 0u }, h_bad = 0u;
  if( nc > 0 )
  {
    const unsigned nblk = sg_div_up( nc, 256 );
    SG_LAUNCH( ctx, "asm_count", double( nc ) * 32.0, k_asm_count<<<nblk, 256, 0, ctx->stream>>>( nc, a, A.ncnt.as<uint32_t>(), A.deg.as<uint32_t>(), A.bad.as<uint32_t>() ) );
    int rc = sg_exclusive_scan<ScanU32>( ctx, "asm_scan", A.ncnt.as<uint32_t>(), nullptr, uint32_t( nc ), uint32_t( nc ), nullptr, A.nouter.as<uint32_t>(), A.total.as<uint32_t>(), true );
    if( rc != SG_OK ) { return rc; }
    rc = sg_exclusive_scan<ScanU32>( ctx, "asm_scan", A.deg.as<uint32_t>(), nullptr, n, n, nullptr, A.inc_start.as<uint32_t>(), nullptr, true );
    if( rc != SG_OK ) { return rc; }
    SG_LAUNCH( ctx, "asm_n_emit", double( nc ) * 96.0, k_asm_n_emit<<<nblk, 256, 0, ctx->stream>>>( nc, a, A.nouter.as<uint32_t>(), A.ninner.as<int32_t>(), A.nval.as<double>(), A.inc_start.as<uint32_t>(),
               A.inc_cursor.as<uint32_t>(), A.inc.as<uint32_t>(), A.bases.as<double>() ) );
    SG_LAUNCH( ctx, "asm_sort_inc", double( nc ) * 16.0, k_asm_sort_incidence<<<sg_div_up( n, 256 ), 256, 0, ctx->stream>>>( n, A.inc_start.as<uint32_t>(), A.inc.as<uint32_t>() ) );
    SG_LAUNCH( ctx, "asm_q_count", double( nc ) * 128.0, k_asm_q<false><<<nblk, 256, 0, ctx->stream>>>( nc, a, d->m.as<double>(), A.inc_start.as<uint32_t>(), A.inc.as<uint32_t>(), A.qcnt.as<uint32_t>(), nullptr, nullptr, nullptr ) );
    rc = sg_exclusive_scan<ScanU32>( ctx, "asm_scan", A.qcnt.as<uint32_t>(), nullptr, uint32_t( nc ), uint32_t( nc ), nullptr, A.qouter.as<uint32_t>(), A.total.as<uint32_t>() + 1, true );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, cudaMemcpyAsync( h_tot, A.total.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( &h_bad, A.bad.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    if( h_bad != 0u ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_assemble: the active set holds contact types other than ball-ball, drum and plane" ); }
    SG_CUDA( ctx, A.qinner.ensure( ( size_t( h_tot[1] ) + 4 ) * 4 ) ); SG_CUDA( ctx, A.qval.ensure( ( size_t( h_tot[1] ) + 4 ) * 8 ) );
    SG_LAUNCH( ctx, "asm_q_emit", double( h_tot[1] ) * 12.0 + double( nc ) * 128.0, k_asm_q<true><<<nblk, 256, 0, ctx->stream>>>( nc, a, d->m.as<double>(), A.inc_start.as<uint32_t>(), A.inc.as<uint32_t>(), nullptr,
               A.qouter.as<uint32_t>(), A.qinner.as<int32_t>(), A.qval.as<double>() ) );
  }
  out->n_nnz = h_tot[0]; out->q_nnz = h_tot[1];
  // results to pinned host memory (the device copies stay valid until the next assemble on this context)
  auto al = []( size_t b ) { return ( b + 63 ) & ~size_t( 63 ); };
  size_t bytes = 64;
  const size_t o_no = bytes; bytes += al( ( nc + 1 ) * 4 );
  const size_t o_ni = bytes; bytes += al( size_t( h_tot[0] ) * 4 );
  const size_t o_nv = bytes; bytes += al( size_t( h_tot[0] ) * 8 );
  const size_t o_qo = bytes; bytes += al( ( nc + 1 ) * 4 );
  const size_t o_qi = bytes; bytes += al( size_t( h_tot[1] ) * 4 );
  const size_t o_qv = bytes; bytes += al( size_t( h_tot[1] ) * 8 );
  const size_t o_b = bytes; bytes += al( 4 * nc * 8 );
  SG_CUDA( ctx, A.host.ensure( bytes ) );
  char* h = A.host.as<char>();
  memset( h + o_no, 0, 4 ); memset( h + o_qo, 0, 4 );
  if( nc > 0 )
  {
    if( flags & SG_ASM_N )
    {
      SG_CUDA( ctx, cudaMemcpyAsync( h + o_no, A.nouter.ptr, ( nc + 1 ) * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( h + o_ni, A.ninner.ptr, size_t( h_tot[0] ) * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( h + o_nv, A.nval.ptr, size_t( h_tot[0] ) * 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    }
    if( flags & SG_ASM_Q )
    {
      SG_CUDA( ctx, cudaMemcpyAsync( h + o_qo, A.qouter.ptr, ( nc + 1 ) * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( h + o_qi, A.qinner.ptr, size_t( h_tot[1] ) * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( h + o_qv, A.qval.ptr, size_t( h_tot[1] ) * 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    }
    if( flags & SG_ASM_BASES ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_b, A.bases.ptr, 4 * nc * 8, cudaMemcpyDeviceToHost, ctx->stream ) ); }
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    sg_prof_collect( ctx );
  }
  if( flags & SG_ASM_N ) { out->n_outer = reinterpret_cast<const int32_t*>( h + o_no ); out->n_inner = reinterpret_cast<const int32_t*>( h + o_ni ); out->n_values = reinterpret_cast<const double*>( h + o_nv ); }
  if( flags & SG_ASM_Q ) { out->q_outer = reinterpret_cast<const int32_t*>( h + o_qo ); out->q_inner = reinterpret_cast<const int32_t*>( h + o_qi ); out->q_values = reinterpret_cast<const double*>( h + o_qv ); }
  if( flags & SG_ASM_BASES ) { out->bases = reinterpret_cast<const double*>( h + o_b ); }
  return SG_OK;
}

// ConstrainedSystem::cacheConstraint for every constraint of the current active set (r: ncomp values per constraint, active-set order)
int sg_ball2d_cache_store( sg_ctx* ctx, uint32_t ncomp, const double* r )
{
  if( ctx == nullptr || ncomp == 0u ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->have_result ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_cache_store: no active set has been computed" ); }
  if( d->slab || d->portal_result ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_cache_store: slab / portal active sets are not supported" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( d->asmd == nullptr ) { d->asmd = new AsmData; }
  AsmData& A = *d->asmd;
  const uint64_t nc = d->n_bb + d->n_static;
  if( nc > 0 && r == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_cache_store: null impulse array" ); }
  // drums precede planes in the active set: how many drums, from the type column
  std::vector<uint32_t> types( d->n_static );
  if( d->n_static > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( types.data(), d->c_type.as<uint32_t>() + d->n_bb, d->n_static * 4, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  SG_CUDA( ctx, A.c_type.ensure( nc * 4 + 4 ) ); SG_CUDA( ctx, A.c_i.ensure( nc * 4 + 4 ) ); SG_CUDA( ctx, A.c_j.ensure( nc * 4 + 4 ) ); SG_CUDA( ctx, A.c_r.ensure( nc * ncomp * 8 + 8 ) );
  if( nc > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( A.c_type.ptr, d->c_type.ptr, nc * 4, cudaMemcpyDeviceToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( A.c_i.ptr, d->c_i.ptr, nc * 4, cudaMemcpyDeviceToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( A.c_j.ptr, d->c_j.ptr, nc * 4, cudaMemcpyDeviceToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( A.c_r.ptr, r, nc * ncomp * 8, cudaMemcpyHostToDevice, ctx->stream ) );
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  uint64_t nd = 0;
  for( uint32_t t : types ) { if( t > SG_BALL_PLANE ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_cache_store: contact type %u is not cacheable (ConstraintCache.cpp exits)", t ); } if( t == SG_BALL_DRUM ) { ++nd; } }
  A.c_n = nc; A.c_nbb = d->n_bb; A.c_ndrum = nd; A.c_ncomp = ncomp;
  return SG_OK;
}

int sg_ball2d_cache_clear( sg_ctx* ctx )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( d->asmd != nullptr ) { d->asmd->c_n = d->asmd->c_nbb = d->asmd->c_ndrum = 0; d->asmd->c_ncomp = 0; }
  return SG_OK;
}

// ConstrainedSystem::getCachedConstraintImpulse for every constraint of the current active set; *hits = how many were found
int sg_ball2d_cache_lookup( sg_ctx* ctx, uint32_t ncomp, double* r_out, uint64_t* hits )
{
  if( ctx == nullptr || ncomp == 0u ) { return SG_ERR_INVALID; }
  Ball2DData* d = ball2d_data( ctx );
  if( !d->have_result ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_cache_lookup: no active set has been computed" ); }
  if( d->slab || d->portal_result ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_ball2d_cache_lookup: slab / portal active sets are not supported" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( d->asmd == nullptr ) { d->asmd = new AsmData; }
  AsmData& A = *d->asmd;
  const uint64_t nc = d->n_bb + d->n_static;
  if( hits != nullptr ) { *hits = 0; }
  if( nc == 0 ) { return SG_OK; }
  if( r_out == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_cache_lookup: null output" ); }
  if( A.c_n > 0 && A.c_ncomp != ncomp ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_ball2d_cache_lookup: the cache holds %u values per constraint, %u asked for", A.c_ncomp, ncomp ); }
  SG_CUDA( ctx, A.lookup.ensure( nc * ncomp * 8 + 16 ) );
  SG_CUDA( ctx, A.total.ensure( 16 ) );
  SG_CUDA( ctx, cudaMemsetAsync( A.total.ptr, 0, 16, ctx->stream ) );
  SG_LAUNCH( ctx, "cache_lookup", double( nc ) * ( 12.0 + 8.0 * ncomp ), k_cache_lookup<<<sg_div_up( nc, 256 ), 256, 0, ctx->stream>>>( nc, ball2d_contacts_view( d ), ncomp, A.c_nbb, A.c_ndrum, A.c_n, A.c_type.as<uint32_t>(),
             A.c_i.as<uint32_t>(), A.c_j.as<uint32_t>(), A.c_r.as<double>(), A.lookup.as<double>(), A.total.as<unsigned long long>() ) );
  unsigned long long h = 0ull;
  SG_CUDA( ctx, cudaMemcpyAsync( r_out, A.lookup.ptr, nc * ncomp * 8, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( &h, A.total.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  if( hits != nullptr ) { *hits = h; }
  return SG_OK;
}

}
