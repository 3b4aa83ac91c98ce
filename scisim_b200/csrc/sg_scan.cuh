// sg_scan.cuh -- exclusive prefix sums over device arrays whose length is only known on the device.
//
// Three launches (tile reduce -> one-block scan of tile sums -> tile down-sweep).  Deterministic,
// order-preserving, no atomics.  Used for (a) cell start offsets from per-cell counts ("prefix-sum cell
// ranges") and (b) per-body output offsets from per-body (candidate, active) counts, which is what puts
// the pair lists in the reference's ascending (i,j) order without a sort of the pairs themselves.
#ifndef SG_SCAN_CUH
#define SG_SCAN_CUH

#include "sg_common.cuh"

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

// n_dev (nullable): element count on the device; n_host is used when n_dev == nullptr
template<typename P>
__global__ void __launch_bounds__( SG_SCAN_THREADS ) sg_scan_reduce( const typename P::In* __restrict__ in, const uint32_t* __restrict__ n_dev, const uint32_t n_host, typename P::Acc* __restrict__ partials )
{
  using Acc = typename P::Acc;
  __shared__ Acc warp_sums[SG_SCAN_THREADS / 32];
  const uint32_t n = ( n_dev != nullptr ) ? *n_dev : n_host;
  const uint64_t base = uint64_t( blockIdx.x ) * SG_SCAN_TILE;
  if( base >= n ) { return; }
  Acc s = P::zero();
  #pragma unroll
  for( int k = 0; k < SG_SCAN_ITEMS; ++k )
  {
    const uint64_t e = base + uint64_t( k ) * SG_SCAN_THREADS + threadIdx.x;
    if( e < n ) { s = P::add( s, P::conv( in[e] ) ); }
  }
  Acc total;
  sg_block_exclusive<P, SG_SCAN_THREADS>( s, warp_sums, &total );
  if( threadIdx.x == 0 ) { partials[blockIdx.x] = total; }
}

// One block: exclusive scan of the tile sums in place; total -> *total_out (and out_end[n] if given)
template<typename P>
__global__ void __launch_bounds__( 1024 ) sg_scan_partials( typename P::Acc* __restrict__ partials, const uint32_t* __restrict__ n_dev, const uint32_t n_host,
                                                            typename P::Acc* __restrict__ total_out, typename P::Out* __restrict__ out_end )
{
  using Acc = typename P::Acc;
  __shared__ Acc warp_sums[32];
  __shared__ Acc carry_s;
  const uint32_t n = ( n_dev != nullptr ) ? *n_dev : n_host;
  const uint32_t ntiles = uint32_t( ( uint64_t( n ) + SG_SCAN_TILE - 1 ) / SG_SCAN_TILE );
  if( threadIdx.x == 0 ) { carry_s = P::zero(); }
  __syncthreads();
  for( uint32_t base = 0; base < ntiles; base += 1024 )
  {
    const uint32_t e = base + threadIdx.x;
    const Acc v = ( e < ntiles ) ? partials[e] : P::zero();
    Acc total;
    const Acc excl = sg_block_exclusive<P, 1024>( v, warp_sums, &total );
    const Acc carry = carry_s;
    if( e < ntiles ) { partials[e] = P::add( carry, excl ); }
    __syncthreads();
    if( threadIdx.x == 0 ) { carry_s = P::add( carry, total ); }
    __syncthreads();
  }
  if( threadIdx.x == 0 )
  {
    if( total_out != nullptr ) { *total_out = carry_s; }
    if( out_end != nullptr ) { out_end[n] = P::out( carry_s ); }
  }
}

template<typename P>
__global__ void __launch_bounds__( SG_SCAN_THREADS ) sg_scan_down( const typename P::In* __restrict__ in, const uint32_t* __restrict__ n_dev, const uint32_t n_host,
                                                                   const typename P::Acc* __restrict__ partials, typename P::Out* __restrict__ out )
{
  using Acc = typename P::Acc;
  __shared__ Acc warp_sums[SG_SCAN_THREADS / 32];
  const uint32_t n = ( n_dev != nullptr ) ? *n_dev : n_host;
  const uint64_t base = uint64_t( blockIdx.x ) * SG_SCAN_TILE;
  if( base >= n ) { return; }
  // blocked arrangement: thread t owns SG_SCAN_ITEMS consecutive elements
  Acc v[SG_SCAN_ITEMS];
  Acc s = P::zero();
  const uint64_t e0 = base + uint64_t( threadIdx.x ) * SG_SCAN_ITEMS;
  #pragma unroll
  for( int k = 0; k < SG_SCAN_ITEMS; ++k )
  {
    v[k] = ( e0 + k < n ) ? P::conv( in[e0 + k] ) : P::zero();
    s = P::add( s, v[k] );
  }
  Acc run = P::add( partials[blockIdx.x], sg_block_exclusive<P, SG_SCAN_THREADS>( s, warp_sums, nullptr ) );
  #pragma unroll
  for( int k = 0; k < SG_SCAN_ITEMS; ++k )
  {
    if( e0 + k < n ) { out[e0 + k] = P::out( run ); }
    run = P::add( run, v[k] );
  }
}

// Whole scan in one 1024-thread block: for arrays small enough that launch latency, not bandwidth, is the cost.
template<typename P>
__global__ void __launch_bounds__( 1024 ) sg_scan_small( const typename P::In* __restrict__ in, const uint32_t n, typename P::Out* __restrict__ out, typename P::Acc* __restrict__ total_out, const bool write_end )
{
  using Acc = typename P::Acc;
  __shared__ Acc warp_sums[32];
  __shared__ Acc carry_s;
  if( threadIdx.x == 0 ) { carry_s = P::zero(); }
  __syncthreads();
  for( uint32_t base = 0; base < n; base += 1024 )
  {
    const uint32_t e = base + threadIdx.x;
    const Acc v = ( e < n ) ? P::conv( in[e] ) : P::zero();
    Acc total;
    const Acc excl = sg_block_exclusive<P, 1024>( v, warp_sums, &total );
    const Acc carry = carry_s;
    if( e < n ) { out[e] = P::out( P::add( carry, excl ) ); }
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

// Host driver.  cap = upper bound on the element count (sizes the grid); partials must hold
// ceil(cap / SG_SCAN_TILE) Acc entries.
template<typename P>
static int sg_exclusive_scan( sg_ctx* ctx, const char* name, const typename P::In* in, const uint32_t* n_dev, const uint32_t n_host, const uint32_t cap,
                              typename P::Acc* partials, typename P::Out* out, typename P::Acc* total_out, const bool write_end )
{
  if( cap == 0 ) { return SG_OK; }
  if( n_dev == nullptr && n_host <= SG_SCAN_SMALL_MAX )
  {
    SG_LAUNCH( ctx, name, double( n_host ) * double( sizeof( typename P::In ) + sizeof( typename P::Out ) ), sg_scan_small<P><<<1, 1024, 0, ctx->stream>>>( in, n_host, out, total_out, write_end ) );
    return SG_OK;
  }
  const unsigned ntiles = sg_div_up( cap, SG_SCAN_TILE );
  const double nelem = double( n_dev != nullptr ? cap : n_host );
  const double bytes_reduce = nelem * double( sizeof( typename P::In ) );
  const double bytes_down = nelem * double( sizeof( typename P::In ) + sizeof( typename P::Out ) );
  SG_LAUNCH( ctx, name, bytes_reduce, sg_scan_reduce<P><<<ntiles, SG_SCAN_THREADS, 0, ctx->stream>>>( in, n_dev, n_host, partials ) );
  SG_LAUNCH( ctx, name, 0.0, sg_scan_partials<P><<<1, 1024, 0, ctx->stream>>>( partials, n_dev, n_host, total_out, write_end ? out : nullptr ) );
  SG_LAUNCH( ctx, name, bytes_down, sg_scan_down<P><<<ntiles, SG_SCAN_THREADS, 0, ctx->stream>>>( in, n_dev, n_host, partials, out ) );
  return SG_OK;
}

#endif
