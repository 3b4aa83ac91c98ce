// sg_pair_sort.cuh -- sort + de-duplication of the teleported-collision list of the portal paths (ball2d and rigidbody2d):
// entries are ( body pair as one 64-bit key, insertion number ); after the bitonic network the first entry of each run of
// equal keys is what std::set<TeleportedCollision>::insert would have kept (ball2d/Portals/PlanarPortal.cpp:53-57).
// Lists are short (bodies at the domain boundary): tiles of SG_B2P_SORT_TILE entries are sorted in shared memory by one CTA,
// strides that cross tiles take one small launch each.  Runs unchanged on the CPU in tests/portal_kernel_harness.cpp.
#ifndef SG_PAIR_SORT_CUH
#define SG_PAIR_SORT_CUH

#include "sg_portal2d.h"

#define SG_B2P_SORT_TILE 2048   // elements sorted per CTA in shared memory (24 KB)
#define SG_B2P_SORT_THREADS 1024

static __global__ void __launch_bounds__( 256 ) k_b2p_sort_pad( const uint32_t first, const uint32_t m, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idxs )
{
  const uint32_t e = first + blockIdx.x * blockDim.x + threadIdx.x;
  if( e < m ) { keys[e] = ~0ull; idxs[e] = ~0u; }
}

// one compare-exchange step of the bitonic network in global memory (the lower element of each pair does the work); used
// for the strides that cross the tiles of k_b2p_bitonic_tile
static __global__ void __launch_bounds__( 256 ) k_b2p_bitonic( const uint32_t m, const uint32_t j, const uint32_t k, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idxs )
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

// Every compare-exchange step whose partner stays inside a tile of TILE consecutive elements, in shared memory (the list is
// a few thousand boundary collisions: one launch instead of log^2 of them).  FULL: all of k = 2 ... TILE, i.e. each tile
// sorted on its own, ascending or descending as its global position demands.  !FULL: the steps j = TILE/2 ... 1 of one
// outer k > TILE, after k_b2p_bitonic did the strides that cross tiles.  m is a power of two; the last tile is padded with
// maximal keys when m < TILE.
template<int TILE, int THREADS, bool FULL>
__global__ void __launch_bounds__( THREADS ) k_b2p_bitonic_tile( const uint32_t m, const uint32_t k_outer, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idxs )
{
  __shared__ unsigned long long s_key[TILE];
  __shared__ uint32_t s_idx[TILE];
  const uint32_t base = blockIdx.x * uint32_t( TILE );
  for( uint32_t t = threadIdx.x; t < uint32_t( TILE ); t += uint32_t( THREADS ) )
  {
    const uint32_t e = base + t;
    s_key[t] = e < m ? keys[e] : ~0ull;
    s_idx[t] = e < m ? idxs[e] : ~0u;
  }
  __syncthreads();
  const uint32_t k_first = FULL ? 2u : k_outer;
  const uint32_t k_last = FULL ? uint32_t( TILE ) : k_outer;
  for( uint32_t k = k_first; k <= k_last; k <<= 1 )
  {
    const uint32_t j_first = ( k >> 1 ) < uint32_t( TILE / 2 ) ? ( k >> 1 ) : uint32_t( TILE / 2 );
    for( uint32_t j = j_first; j > 0u; j >>= 1 )
    {
      for( uint32_t t = threadIdx.x; t < uint32_t( TILE ); t += uint32_t( THREADS ) )
      {
        const uint32_t f = sg_bitonic_partner( t, j );
        if( f > t )
        {
          const unsigned long long ka = s_key[t], kb = s_key[f];
          const uint32_t ia = s_idx[t], ib = s_idx[f];
          const bool swap = sg_bitonic_ascending( base + t, k ) ? sg_tele_less( kb, ib, ka, ia ) : sg_tele_less( ka, ia, kb, ib );
          if( swap ) { s_key[t] = kb; s_key[f] = ka; s_idx[t] = ib; s_idx[f] = ia; }
        }
      }
      __syncthreads();
    }
  }
  for( uint32_t t = threadIdx.x; t < uint32_t( TILE ); t += uint32_t( THREADS ) )
  {
    const uint32_t e = base + t;
    if( e < m ) { keys[e] = s_key[t]; idxs[e] = s_idx[t]; }
  }
}

// std::set<TeleportedCollision>::insert keeps the first collision of each body pair: after the sort that is the first
// entry of each run of equal keys
static __global__ void __launch_bounds__( 256 ) k_b2p_unique( const uint32_t nraw, const unsigned long long* __restrict__ keys, uint32_t* __restrict__ uflag )
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if( e >= nraw ) { return; }
  uflag[e] = ( e == 0u || keys[e] != keys[e - 1u] ) ? 1u : 0u;
}

#endif
