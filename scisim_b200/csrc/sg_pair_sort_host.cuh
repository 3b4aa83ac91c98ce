// sg_pair_sort_host.cuh -- host side of sg_pair_sort.cuh: the launch sequence that sorts a teleported-collision list and flags the
// first entry of every body pair, shared by the three portal drivers (ball2d, rigidbody2d, rigidbody3d).
#ifndef SG_PAIR_SORT_HOST_CUH
#define SG_PAIR_SORT_HOST_CUH

#include "sg_scan.cuh"
#include "sg_pair_sort.cuh"

// keys / idxs hold nraw entries padded to m (a power of two >= nraw); on return they are sorted by (key, insertion number),
// uflag[e] = 1 for the first entry of each key, uoff = its exclusive scan and *utotal (device) the number of distinct keys.
// Everything is asynchronous on the context's stream.
static int sg_tele_sort_unique( sg_ctx* ctx, const uint32_t nraw, const uint32_t m, unsigned long long* keys, uint32_t* idxs, uint32_t* uflag, uint32_t* uoff, uint32_t* u_partials, uint32_t* utotal )
{
  if( m > nraw ) { SG_LAUNCH( ctx, "b2p_sort_pad", 0.0, k_b2p_sort_pad<<<sg_div_up( m - nraw, 256 ), 256, 0, ctx->stream>>>( nraw, m, keys, idxs ) ); }
  // tiles of SG_B2P_SORT_TILE elements are sorted in shared memory; only strides that cross tiles take a launch each
  const unsigned ntiles = sg_div_up( m, SG_B2P_SORT_TILE );
  SG_LAUNCH( ctx, "b2p_bitonic_tile", double( m ) * 24.0, k_b2p_bitonic_tile<SG_B2P_SORT_TILE, SG_B2P_SORT_THREADS, true><<<ntiles, SG_B2P_SORT_THREADS, 0, ctx->stream>>>( m, 0u, keys, idxs ) );
  for( uint32_t k = 2u * SG_B2P_SORT_TILE; k <= m; k <<= 1 )
  {
    for( uint32_t j = k >> 1; j >= uint32_t( SG_B2P_SORT_TILE ); j >>= 1 )
    {
      SG_LAUNCH( ctx, "b2p_bitonic", double( m ) * 24.0, k_b2p_bitonic<<<sg_div_up( m, 256 ), 256, 0, ctx->stream>>>( m, j, k, keys, idxs ) );
    }
    SG_LAUNCH( ctx, "b2p_bitonic_tile", double( m ) * 24.0, k_b2p_bitonic_tile<SG_B2P_SORT_TILE, SG_B2P_SORT_THREADS, false><<<ntiles, SG_B2P_SORT_THREADS, 0, ctx->stream>>>( m, k, keys, idxs ) );
  }
  SG_CUDA( ctx, cudaMemsetAsync( utotal, 0, 4, ctx->stream ) );
  SG_LAUNCH( ctx, "b2p_unique", double( nraw ) * 12.0, k_b2p_unique<<<sg_div_up( nraw, 256 ), 256, 0, ctx->stream>>>( nraw, keys, uflag ) );
  return sg_exclusive_scan<ScanU32>( ctx, "b2p_unique_scan", uflag, nullptr, nraw, nraw, u_partials, uoff, utotal, false );
}

#endif
