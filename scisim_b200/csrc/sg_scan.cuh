// sg_scan.cuh -- exclusive prefix sums over device arrays whose length is only known on the device.
//
// One cooperative launch (chunk reduce -> grid barrier -> chunk scan), or one block for small arrays.  Deterministic,
// order-preserving, no atomics.  Used for (a) cell start offsets from per-cell counts ("prefix-sum cell
// ranges") and (b) per-body output offsets from per-body (candidate, active) counts, which is what puts
// the pair lists in the reference's ascending (i,j) order without a sort of the pairs themselves.
#ifndef SG_SCAN_CUH
#define SG_SCAN_CUH

#include "sg_common.cuh"

#include <cooperative_groups.h>

#define SG_SCAN_THREADS 256
#define SG_SCAN_ITEMS 4
#define SG_SCAN_TILE ( SG_SCAN_THREADS * SG_SCAN_ITEMS )

// ---- element policies -----------------------------------------------------------------------------
struct ScanU32
{
  using In = uint32_t;
  using Acc = uint32_t;
  using Out = uint32_t;
  __device__ static Acc zero() { return 0u; }
  __device__ static Acc conv( const In x ) { return x; }
  __device__ static Acc add( const Acc a, const Acc b ) { return a + b; }
  __device__ static Acc shfl_up( const Acc a, const int d ) { return __shfl_up_sync( 0xffffffffu, a, d ); }
  __device__ static Acc shfl( const Acc a, const int l ) { return __shfl_sync( 0xffffffffu, a, l ); }
  __device__ static Out out( const Acc a ) { return a; }
};

// (candidate count, active count) per body -> 64-bit running offsets for both lists at once
struct ScanPairCounts
{
  using In = uint2;
  struct Acc { unsigned long long c; unsigned long long a; };
  using Out = ulonglong2;
  __device__ static Acc zero() { return Acc{ 0ull, 0ull }; }
  __device__ static Acc conv( const In x ) { return Acc{ x.x, x.y & 0x7fffffffu }; } // bit 31 of y is a flag (sg_broadphase.cuh)
  __device__ static Acc add( const Acc a, const Acc b ) { return Acc{ a.c + b.c, a.a + b.a }; }
  __device__ static Acc shfl_up( const Acc a, const int d )
  {
    return Acc{ __shfl_up_sync( 0xffffffffu, a.c, d ), __shfl_up_sync( 0xffffffffu, a.a, d ) };
  }
  __device__ static Acc shfl( const Acc a, const int l )
  {
    return Acc{ __shfl_sync( 0xffffffffu, a.c, l ), __shfl_sync( 0xffffffffu, a.a, l ) };
  }
  __device__ static Out out( const Acc a ) { return make_ulonglong2( a.c, a.a ); }
};

// Exclusive scan of one value per thread across the block; returns the exclusive prefix, *block_total gets the sum.
template<typename P, int THREADS>
__device__ inline typename P::Acc sg_block_exclusive( const typename P::Acc v, typename P::Acc* warp_sums, typename P::Acc* block_total )
{
  using Acc = typename P::Acc;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  Acc incl = v;
  #pragma unroll
  for( int d = 1; d < 32; d <<= 1 )
  {
    const Acc o = P::shfl_up( incl, d );
    if( lane >= d ) { incl = P::add( o, incl ); }
  }
  if( lane == 31 ) { warp_sums[warp] = incl; }
  __syncthreads();
  if( warp == 0 )
  {
    constexpr int NW = THREADS / 32;
    Acc w = ( lane < NW ) ? warp_sums[lane] : P::zero();
    #pragma unroll
    for( int d = 1; d < 32; d <<= 1 )
    {
      const Acc o = P::shfl_up( w, d );
      if( lane >= d ) { w = P::add( o, w ); }
    }
    if( lane < NW ) { warp_sums[lane] = w; } // inclusive over warps
  }
  __syncthreads();
  Acc excl = P::shfl_up( incl, 1 );
  if( lane == 0 ) { excl = P::zero(); }
  if( warp > 0 ) { excl = P::add( warp_sums[warp - 1], excl ); }
  if( block_total != nullptr ) { *block_total = warp_sums[THREADS / 32 - 1]; }
  return excl;
}

// Whole scan in one 1024-thread block: for arrays small enough that launch latency, not bandwidth, is the cost.
// Each thread owns 8 consecutive elements per round.
template<typename P>
__global__ void __launch_bounds__( 1024 ) sg_scan_small( const typename P::In* __restrict__ in, const uint32_t n, typename P::Out* __restrict__ out, typename P::Acc* __restrict__ total_out, const bool write_end, const uint32_t* __restrict__ scatter )
{
  using Acc = typename P::Acc;
  constexpr int ITEMS = 8;
  __shared__ Acc warp_sums[32];
  __shared__ Acc carry_s;
  if( threadIdx.x == 0 ) { carry_s = P::zero(); }
  __syncthreads();
  for( uint32_t base = 0; base < n; base += 1024 * ITEMS )
  {
    const uint32_t e0 = base + threadIdx.x * ITEMS;
    Acc v[ITEMS];
    Acc s = P::zero();
    #pragma unroll
    for( int k = 0; k < ITEMS; ++k ) { v[k] = ( e0 + k < n ) ? P::conv( in[e0 + k] ) : P::zero(); s = P::add( s, v[k] ); }
    Acc total;
    const Acc excl = sg_block_exclusive<P, 1024>( s, warp_sums, &total );
    const Acc carry = carry_s;
    Acc run = P::add( carry, excl );
    #pragma unroll
    for( int k = 0; k < ITEMS; ++k )
    {
      if( e0 + k < n )
      {
        if( scatter == nullptr ) { out[e0 + k] = P::out( run ); }
        else { const uint32_t d = __ldg( &scatter[e0 + k] ); if( d != 0xffffffffu ) { out[d] = P::out( run ); } }
      }
      run = P::add( run, v[k] );
    }
    __syncthreads();
    if( threadIdx.x == 0 ) { carry_s = P::add( carry, total ); }
    __syncthreads();
  }
  if( threadIdx.x == 0 )
  {
    if( total_out != nullptr ) { *total_out = carry_s; }
    if( write_end ) { out[n] = P::out( carry_s ); }
  }
}

#define SG_SCAN_SMALL_MAX 32768u

#define SG_LB_THREADS 256
#define SG_LB_ITEMS 8
#define SG_LB_TILE ( SG_LB_THREADS * SG_LB_ITEMS )

template<typename Acc>
__device__ __forceinline__ Acc sg_ld_volatile( const Acc* p )
{
  static_assert( sizeof( Acc ) % 4 == 0, "Acc is a multiple of 4 bytes" );
  union { Acc a; uint32_t w[sizeof( Acc ) / 4]; } u;
  const volatile uint32_t* src = reinterpret_cast<const volatile uint32_t*>( p );
  #pragma unroll
  for( int k = 0; k < int( sizeof( Acc ) / 4 ); ++k ) { u.w[k] = src[k]; }
  return u.a;
}
template<typename Acc>
__device__ __forceinline__ void sg_st_volatile( Acc* p, const Acc& v )
{
  union { Acc a; uint32_t w[sizeof( Acc ) / 4]; } u;
  u.a = v;
  volatile uint32_t* dst = reinterpret_cast<volatile uint32_t*>( p );
  #pragma unroll
  for( int k = 0; k < int( sizeof( Acc ) / 4 ); ++k ) { dst[k] = u.w[k]; }
}

// A small u32 exclusive scan that rides along with a cooperative scan launch (one extra block), so that a second,
// latency-bound scan does not cost a launch of its own.
struct SideScan
{
  const uint32_t* in;
  uint32_t n;
  uint32_t* out;
  uint32_t* total;
};
#define SG_SIDE_SCAN_MAX 131072u // 64 extra blocks at most

// Extra block `sb` of the launch owns tile sb of the side array (SG_LB_TILE elements, kept in registers across the
// grid barrier); tile totals go through side_partials.
__device__ inline void sg_side_scan( const SideScan side, const uint32_t sb, uint32_t* side_partials, cooperative_groups::grid_group grid )
{
  __shared__ uint32_t side_warp[SG_LB_THREADS / 32];
  __shared__ uint32_t side_prefix;
  const uint32_t e0 = sb * SG_LB_TILE + threadIdx.x * SG_LB_ITEMS;
  uint32_t v[SG_LB_ITEMS];
  uint32_t ls = 0u;
  #pragma unroll
  for( int k = 0; k < SG_LB_ITEMS; ++k ) { v[k] = ( e0 + k < side.n ) ? side.in[e0 + k] : 0u; ls += v[k]; }
  uint32_t ttot;
  const uint32_t excl = sg_block_exclusive<ScanU32, SG_LB_THREADS>( ls, side_warp, &ttot );
  if( threadIdx.x == 0 ) { sg_st_volatile( &side_partials[sb], ttot ); }
  __threadfence();
  grid.sync();
  const uint32_t ntiles = ( side.n + SG_LB_TILE - 1u ) / SG_LB_TILE;
  if( threadIdx.x < 32 )
  {
    uint32_t pre = 0u, all = 0u;
    for( uint32_t b = threadIdx.x; b < ntiles; b += 32u ) { const uint32_t t = sg_ld_volatile( &side_partials[b] ); all += t; if( b < sb ) { pre += t; } }
    #pragma unroll
    for( int d = 16; d > 0; d >>= 1 ) { pre += __shfl_xor_sync( 0xffffffffu, pre, d ); all += __shfl_xor_sync( 0xffffffffu, all, d ); }
    if( threadIdx.x == 0 ) { side_prefix = pre; if( sb == 0u && side.total != nullptr ) { *side.total = all; } }
  }
  __syncthreads();
  uint32_t run = side_prefix + excl;
  #pragma unroll
  for( int k = 0; k < SG_LB_ITEMS; ++k )
  {
    if( e0 + k < side.n ) { side.out[e0 + k] = run; }
    run += v[k];
  }
}

// ---- two-phase scan in one cooperative launch -------------------------------------------------------
// All blocks are co-resident (cooperative launch, grid = a fixed multiple of the SM count): each block reduces its
// contiguous chunk, the grid synchronises once, every block sums the (few hundred) chunk totals that precede it and
// sweeps its chunk again (second read served by L1/L2).  Removes the serial tile-to-tile latency chain a look-back
// scan has when every tile is resident at once.
template<typename P>
__global__ void __launch_bounds__( SG_LB_THREADS ) sg_scan_coop( const typename P::In* __restrict__ in, const uint32_t* __restrict__ n_dev, const uint32_t n_host, typename P::Out* __restrict__ out,
                                                                const uint32_t* __restrict__ scatter, typename P::Acc* partials, typename P::Acc* __restrict__ total_out, const bool write_end,
                                                                const SideScan side )
{
  using Acc = typename P::Acc;
  namespace cg = cooperative_groups;
  __shared__ Acc warp_sums[SG_LB_THREADS / 32];
  __shared__ Acc s_prefix;
  const uint32_t n = ( n_dev != nullptr ) ? *n_dev : n_host;
  const uint32_t nb = gridDim.x - ( side.n + SG_LB_TILE - 1u ) / SG_LB_TILE;
  if( blockIdx.x >= nb )
  {
    // extra blocks: a small independent u32 scan that shares this launch and its grid barrier
    sg_side_scan( side, blockIdx.x - nb, reinterpret_cast<uint32_t*>( partials + nb ), cg::this_grid() );
    return;
  }
  // chunk per block: whole tiles
  const uint64_t tiles_total = ( uint64_t( n ) + SG_LB_TILE - 1 ) / SG_LB_TILE;
  const uint64_t tiles_per_block = ( tiles_total + nb - 1 ) / nb;
  const uint64_t c0 = uint64_t( blockIdx.x ) * tiles_per_block * SG_LB_TILE;
  const uint64_t c1 = ( c0 + tiles_per_block * SG_LB_TILE < n ) ? c0 + tiles_per_block * SG_LB_TILE : n;
  // phase 1: chunk total
  Acc s = P::zero();
  for( uint64_t e = c0 + threadIdx.x; e < c1; e += SG_LB_THREADS ) { s = P::add( s, P::conv( in[e] ) ); }
  Acc total;
  sg_block_exclusive<P, SG_LB_THREADS>( s, warp_sums, &total );
  if( threadIdx.x == 0 ) { sg_st_volatile( &partials[blockIdx.x], total ); }
  __threadfence();
  cg::this_grid().sync();
  // phase 2: prefix of the preceding chunk totals (+ grand total for the last block)
  {
    Acc pre = P::zero();
    for( uint32_t b = threadIdx.x; b < blockIdx.x; b += SG_LB_THREADS ) { pre = P::add( pre, sg_ld_volatile( &partials[b] ) ); }
    __syncthreads();
    Acc tot2;
    sg_block_exclusive<P, SG_LB_THREADS>( pre, warp_sums, &tot2 );
    if( threadIdx.x == 0 ) { s_prefix = tot2; }
    __syncthreads();
  }
  Acc carry = s_prefix;
  for( uint64_t t0 = c0; t0 < c1; t0 += SG_LB_TILE )
  {
    const uint64_t e0 = t0 + uint64_t( threadIdx.x ) * SG_LB_ITEMS;
    Acc v[SG_LB_ITEMS];
    Acc ls = P::zero();
    #pragma unroll
    for( int k = 0; k < SG_LB_ITEMS; ++k ) { v[k] = ( e0 + k < c1 ) ? P::conv( in[e0 + k] ) : P::zero(); ls = P::add( ls, v[k] ); }
    __syncthreads();
    Acc ttot;
    const Acc excl = sg_block_exclusive<P, SG_LB_THREADS>( ls, warp_sums, &ttot );
    Acc run = P::add( carry, excl );
    #pragma unroll
    for( int k = 0; k < SG_LB_ITEMS; ++k )
    {
      if( e0 + k < c1 )
      {
        if( scatter == nullptr ) { out[e0 + k] = P::out( run ); }
        else { const uint32_t d = __ldg( &scatter[e0 + k] ); if( d != 0xffffffffu ) { out[d] = P::out( run ); } }
      }
      run = P::add( run, v[k] );
    }
    carry = P::add( carry, ttot );
  }
  if( blockIdx.x == nb - 1u && threadIdx.x == 0 )
  {
    // the last block's carry is the grand total (empty chunks carry their prefix through unchanged)
    if( total_out != nullptr ) { *total_out = carry; }
    if( write_end ) { out[n] = P::out( carry ); }
  }
}

// Host driver.  cap = upper bound on the element count.
template<typename P>
static int sg_exclusive_scan( sg_ctx* ctx, const char* name, const typename P::In* in, const uint32_t* n_dev, const uint32_t n_host, const uint32_t cap,
                              typename P::Acc* partials, typename P::Out* out, typename P::Acc* total_out, const bool write_end, const uint32_t* scatter = nullptr,
                              const SideScan* side_job = nullptr )
{
  ( void ) partials;
  if( cap == 0 ) { return SG_OK; }
  SideScan side;
  side.in = nullptr; side.n = 0u; side.out = nullptr; side.total = nullptr;
  if( side_job != nullptr ) { side = *side_job; }
  if( side.n == 0u && n_dev == nullptr && n_host <= SG_SCAN_SMALL_MAX )
  {
    SG_LAUNCH( ctx, name, double( n_host ) * double( sizeof( typename P::In ) + sizeof( typename P::Out ) ), sg_scan_small<P><<<1, 1024, 0, ctx->stream>>>( in, n_host, out, total_out, write_end, scatter ) );
    return SG_OK;
  }
  static_assert( sizeof( typename P::Acc ) <= 16, "scan accumulators are at most 16 bytes" );
  // every block must be resident at once: ask the runtime how many fit (cached per scan type), use at most 4 per SM
  static int per_sm = 0;
  if( per_sm == 0 )
  {
    int occ = 0;
    SG_CUDA( ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor( &occ, sg_scan_coop<P>, SG_LB_THREADS, 0 ) );
    if( occ < 1 ) { return sg_fail( ctx, SG_ERR_INTERNAL, "sg_exclusive_scan: cooperative scan kernel does not fit on an SM" ); }
    per_sm = occ < 4 ? occ : 4;
  }
  const unsigned side_blocks = ( side.n + SG_LB_TILE - 1u ) / SG_LB_TILE;
  const unsigned nblocks = unsigned( ctx->num_sms ) * unsigned( per_sm ) - side_blocks;
  if( size_t( nblocks + side_blocks ) * 16 + 64 > ctx->scan_vals.cap ) { SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) ); SG_CUDA( ctx, ctx->scan_vals.ensure( size_t( ctx->num_sms ) * 4 * 32 + 256 ) ); }
  typename P::Acc* chunk_totals = reinterpret_cast<typename P::Acc*>( ctx->scan_vals.ptr );
  const double nelem = double( n_dev != nullptr ? cap : n_host );
  bool we = write_end;
  uint32_t nh = n_host;
  void* args[] = { ( void* ) &in, ( void* ) &n_dev, ( void* ) &nh, ( void* ) &out, ( void* ) &scatter, ( void* ) &chunk_totals, ( void* ) &total_out, ( void* ) &we, ( void* ) &side };
  SG_LAUNCH( ctx, name, nelem * double( sizeof( typename P::In ) + sizeof( typename P::Out ) ) + double( side.n ) * 8.0,
             SG_CUDA( ctx, cudaLaunchCooperativeKernel( ( const void* ) sg_scan_coop<P>, dim3( nblocks + side_blocks ), dim3( SG_LB_THREADS ), args, 0, ctx->stream ) ) );
  return SG_OK;
}

#endif
