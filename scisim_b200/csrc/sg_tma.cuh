// sg_tma.cuh -- the few PTX wrappers this library needs for TMA-style bulk copies (sm_90+/sm_100a):
// 1-D cp.async.bulk global -> shared with mbarrier transaction-count completion (SASS: UBLKCP + SYNCS).
#ifndef SG_TMA_CUH
#define SG_TMA_CUH

#include <cstdint>

__device__ __forceinline__ uint32_t sg_smem_u32( const void* p ) { return static_cast<uint32_t>( __cvta_generic_to_shared( p ) ); }

__device__ __forceinline__ void sg_mbar_init( uint64_t* bar, const uint32_t arrivals )
{
  asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( sg_smem_u32( bar ) ), "r"( arrivals ) : "memory" );
  // make the initialised barrier visible to the async proxy before any bulk copy names it
  asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
}

__device__ __forceinline__ void sg_mbar_arrive_expect_tx( uint64_t* bar, const uint32_t bytes )
{
  asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( sg_smem_u32( bar ) ), "r"( bytes ) : "memory" );
}

// bytes: multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void sg_bulk_g2s( void* smem_dst, const void* gmem_src, const uint32_t bytes, uint64_t* bar )
{
  asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"( sg_smem_u32( smem_dst ) ), "l"( gmem_src ), "r"( bytes ), "r"( sg_smem_u32( bar ) ) : "memory" );
}

__device__ __forceinline__ void sg_mbar_wait( uint64_t* bar, const uint32_t parity )
{
  uint32_t done = 0;
  const uint32_t addr = sg_smem_u32( bar );
  do
  {
    asm volatile( "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"( done ) : "r"( addr ), "r"( parity ) : "memory" );
  } while( done == 0 );
}

// same wait for a thread whose spinning would only steal issue slots from the warps doing the work (a producer
// waiting for its consumers): back off between polls
__device__ __forceinline__ void sg_mbar_wait_backoff( uint64_t* bar, const uint32_t parity )
{
  uint32_t done = 0;
  const uint32_t addr = sg_smem_u32( bar );
  for( ;; )
  {
    // suspend-time hint: the hardware may park the thread for up to ~20 us per try instead of returning at once
    asm volatile( "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"( done ) : "r"( addr ), "r"( parity ), "r"( 20000u ) : "memory" );
    if( done != 0 ) { break; }
    __nanosleep( 1000 );
  }
}

// non-blocking arrive (count 1) -- consumers releasing a stage back to the producer
__device__ __forceinline__ void sg_mbar_arrive( uint64_t* bar )
{
  asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( sg_smem_u32( bar ) ) : "memory" );
}

// 2-D tiled tensor copy global -> shared (SASS: UTMALDG.2D); c0 = innermost coordinate
__device__ __forceinline__ void sg_tma_load_2d( void* smem_dst, const void* tensor_map, const int c0, const int c1, uint64_t* bar )
{
  asm volatile( "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"( sg_smem_u32( smem_dst ) ), "l"( tensor_map ), "r"( c0 ), "r"( c1 ), "r"( sg_smem_u32( bar ) ) : "memory" );
}

#endif
