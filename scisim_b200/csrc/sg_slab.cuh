// sg_slab.cuh -- what the multi-GPU slab modes of ball2d and rigidbody3d share: the mailbox a rank's two
// neighbours write over NVLink, the step-tagged flags, the interval a rank publishes, the candidate bands, and the single-pass
// halo pack.  No reference counterpart (SCISim is single-process); DESIGN.md section 5.
#ifndef SG_SLAB_CUH
#define SG_SLAB_CUH

#include "sg_common.cuh"

// Slab mode, peer-memory exchange: while the flow kernel has every owned body's swept box in registers it also lists the bodies
// that COULD be owed to a neighbour this step -- those reaching the band next to that neighbour, band = the neighbour's interval
// of the previous step widened by a margin -- so that the halo pack, once the neighbour's actual interval has arrived, only looks
// at that short list (and checks that the actual interval lies inside the band; if not, it scans all bodies: correctness never
// depends on the guess).  state: [0] = left band's upper edge (a body is a candidate for the lower neighbour if lo.x <= it),
// [1] = right band's lower edge (candidate for the higher neighbour if hi.x >= it).
struct SlabCand
{
  const double* band;      // 2 doubles (device); nullptr: no candidate lists
  uint32_t* count;         // 2 counters
  uint32_t* list[2];       // slot indices
  uint32_t cap;
  bool on[2];
};


static __global__ void k_slab_begin( long long* enc, uint32_t* ghost_counts )
{
  enc[0] = sg_ordered_from_double( __longlong_as_double( 0x7ff0000000000000LL ) );
  enc[1] = sg_ordered_from_double( __longlong_as_double( 0xfff0000000000000LL ) );
  ghost_counts[0] = 0u; ghost_counts[1] = 0u;
}
// reads the reduced interval, re-arms the accumulator for the next step and clears this step's ghost counts
__device__ inline void slab_take_interval( long long* enc, uint32_t* ghost_counts, double& lo, double& hi )
{
  lo = sg_double_from_ordered( enc[0] ); hi = sg_double_from_ordered( enc[1] );
  enc[0] = sg_ordered_from_double( __longlong_as_double( 0x7ff0000000000000LL ) );
  enc[1] = sg_ordered_from_double( __longlong_as_double( 0xfff0000000000000LL ) );
  ghost_counts[0] = 0u; ghost_counts[1] = 0u;
}
static __global__ void k_slab_interval_decode( long long* enc, uint32_t* ghost_counts, double* out )
{
  double lo, hi;
  slab_take_interval( enc, ghost_counts, lo, hi );
  out[0] = lo; out[1] = hi;
}


// ---- peer-memory halo exchange -----------------------------------------------------------------------
// Each rank owns a mailbox in its own HBM that its two neighbours write over NVLink (mapped with CUDA IPC, or
// directly when the neighbour lives in the same process): the neighbour's swept interval, then -- packed by the
// neighbour's own pack kernel straight through the peer mapping -- its halo records, each followed by a
// system-scope fence and a step-tagged flag.  The consumer side is a one-thread kernel spinning on the flag in
// LOCAL memory, so the whole exchange is stream-ordered device work: no collective, no host round trip.
struct alignas( 128 ) SlabMailboxHdr
{
  double iv[2][2];        // [side]: interval of the neighbour on that side (0 = lower ranks, 1 = higher)
  uint32_t iv_flag[2];    // step tag of iv[side]
  uint32_t halo_flag[2];  // step tag of the halo records from that side
  uint32_t err;           // a wait timed out
};
// halo records of the neighbour on `side`: a header record (its gid field = the count) followed by up to cap records
template<typename Rec>
__host__ __device__ inline Rec* slab_mailbox_halo( void* mb, const int side, const uint32_t cap )
{
  return reinterpret_cast<Rec*>( static_cast<unsigned char*>( mb ) + sizeof( SlabMailboxHdr ) ) + size_t( side ) * ( size_t( cap ) + 1 );
}
__device__ __forceinline__ void st_release_sys( uint32_t* p, const uint32_t v ) { asm volatile( "st.release.sys.global.u32 [%0], %1;" ::"l"( p ), "r"( v ) : "memory" ); }
__device__ __forceinline__ uint32_t ld_acquire_sys( const uint32_t* p ) { uint32_t v; asm volatile( "ld.acquire.sys.global.u32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" ); return v; }


// one thread: returns when *flag has reached `step` (bounded: a dead neighbour must not hang the GPU)
__device__ inline void slab_wait_flag( const uint32_t* flag, const uint32_t step, uint32_t* err )
{
  unsigned long long t0, t1;
  asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t0 ) );
  while( int( ld_acquire_sys( flag ) - step ) < 0 )
  {
    __nanosleep( 200 );
    asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t1 ) );
    if( t1 - t0 > 10000000000ull ) { *err = 1u; break; } // 10 s
  }
}


// ---- single-pass halo pack over the candidate lists (peer-memory exchange) ----------------------------------------------
struct SlabCandState
{
  double band[2];        // see SlabCand
  uint32_t count[2];     // candidates listed by this step's flow kernel
  uint32_t cursor[2];    // records written into the neighbour's mailbox so far
  uint32_t ticket;       // blocks done (the last one publishes)
  uint32_t fallbacks;    // steps in which a band did not hold and all bodies were scanned (diagnostics)
};

static __global__ void k_slab_cand_reset( SlabCandState* st )
{
  // bands that make every body a candidate: the first step after (re)initialisation overflows the lists and scans all bodies
  st->band[0] = __longlong_as_double( 0x7ff0000000000000LL ); st->band[1] = __longlong_as_double( 0xfff0000000000000LL );
  st->count[0] = st->count[1] = 0u; st->cursor[0] = st->cursor[1] = 0u; st->ticket = 0u; st->fallbacks = 0u;
}

template<typename Rec>
struct Pack2Args
{
  const double* iv[2];      // the neighbour's interval of this step (local mailbox)
  const uint32_t* wait[2];  // its step-tagged flag
  Rec* out[2];              // the neighbour's mailbox (peer memory): header + records
  uint32_t* post[2];        // the neighbour's halo flag
  bool on[2];
  uint32_t* err;
  uint32_t step;
};

// Small persistent grid.  Every block: wait for the neighbours' intervals; per side, if the interval lies inside the band the
// candidates were collected with (and the list did not overflow) test the candidates, else all owned bodies; selected bodies go
// straight into the neighbour's mailbox at a slot taken from an atomic cursor.  The last block to finish writes the headers
// (counts), raises the neighbours' flags and sets the bands for the next step: this step's interval edge widened by `margin`.
// Traits: Rec (the halo record, with a uint32_t gid field that doubles as the count in the header record), Src (where the bodies live),
//         select( src, slot, ilo, ihi, rec ) -> the body's box reaches [ilo, ihi] on x (closed, like AABB::overlaps); fills rec
template<typename Traits>
__global__ void __launch_bounds__( 256 ) k_slab_pack2( const uint32_t own_first, const uint32_t own_count, const typename Traits::Src src, const uint32_t cap, const SlabCand sc, SlabCandState* st,
                                                      const GridParams* __restrict__ last_grid, const Pack2Args<typename Traits::Rec> args )
{
  if( threadIdx.x < 2 && args.on[threadIdx.x] ) { slab_wait_flag( args.wait[threadIdx.x], args.step, args.err ); }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  #pragma unroll
  for( int sd = 0; sd < 2; ++sd )
  {
    if( !args.on[sd] ) { continue; }
    const double ilo = args.iv[sd][0], ihi = args.iv[sd][1];
    const uint32_t listed = st->count[sd];
    // the band holds if every body the neighbour could need was listed: its interval does not reach past the band edge
    const bool band_ok = listed <= sc.cap && ( ( sd == 0 ) ? ( ihi <= st->band[0] ) : ( ilo >= st->band[1] ) );
    const uint32_t total = band_ok ? listed : own_count;
    for( uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x ) // whole warps stay together for the ballot
    {
      const uint32_t e = base + threadIdx.x;
      bool sel = false;
      typename Traits::Rec g;
      if( e < total )
      {
        const uint32_t i = band_ok ? sc.list[sd][e] : own_first + e;
        sel = Traits::select( src, i, ilo, ihi, g );
      }
      const unsigned bal = __ballot_sync( 0xffffffffu, sel );
      if( bal != 0u )
      {
        uint32_t k0 = 0u;
        if( lane == __ffs( bal ) - 1 ) { k0 = atomicAdd( &st->cursor[sd], uint32_t( __popc( bal ) ) ); }
        k0 = __shfl_sync( 0xffffffffu, k0, __ffs( bal ) - 1 );
        const uint32_t k = k0 + __popc( bal & ( ( 1u << lane ) - 1u ) );
        if( sel && k < cap ) { args.out[sd][1u + k] = g; }
      }
    }
  }
  // ---- the last block publishes
  __shared__ uint32_t s_last;
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    __threadfence_system(); // this block's peer writes before its ticket
    s_last = ( atomicAdd( &st->ticket, 1u ) == gridDim.x - 1u ) ? 1u : 0u;
  }
  __syncthreads();
  if( s_last == 0u || threadIdx.x != 0 ) { return; }
  __threadfence();
  const double margin = ( last_grid != nullptr && last_grid->h > 0.0 && last_grid->h < 1.0e300 ) ? last_grid->h : 0.0; // the last step's cell width (>= every swept extent); anything >= 0 is safe
  for( int sd = 0; sd < 2; ++sd )
  {
    if( !args.on[sd] ) { continue; }
    const uint32_t listed = st->count[sd];
    const double ilo = args.iv[sd][0], ihi = args.iv[sd][1];
    const bool band_ok = listed <= sc.cap && ( ( sd == 0 ) ? ( ihi <= st->band[0] ) : ( ilo >= st->band[1] ) );
    if( !band_ok ) { st->fallbacks += 1u; }
    typename Traits::Rec h;
    memset( &h, 0, sizeof( h ) );
    h.gid = *reinterpret_cast<volatile uint32_t*>( &st->cursor[sd] );
    args.out[sd][0] = h;
    // next step's band: the neighbour's edge of this step, widened (an empty neighbour posts [+inf, -inf]: nobody is a candidate)
    if( sd == 0 ) { st->band[0] = ( ilo <= ihi ) ? ihi + margin : __longlong_as_double( 0xfff0000000000000LL ); }
    else { st->band[1] = ( ilo <= ihi ) ? ilo - margin : __longlong_as_double( 0x7ff0000000000000LL ); }
    st->cursor[sd] = 0u; st->count[sd] = 0u;
  }
  st->ticket = 0u;
  __threadfence_system();
  for( int sd = 0; sd < 2; ++sd ) { if( args.on[sd] ) { st_release_sys( args.post[sd], args.step ); } }
}


// decodes this rank's interval, keeps a local copy and posts it to the neighbours (lower = side 0, higher = side 1)
static __global__ void k_slab_post_interval( long long* enc, uint32_t* ghost_counts, double* local_out, SlabMailboxHdr* lower, SlabMailboxHdr* higher, const uint32_t step )
{
  double lo, hi;
  slab_take_interval( enc, ghost_counts, lo, hi );
  if( local_out != nullptr ) { local_out[0] = lo; local_out[1] = hi; }
  if( lower != nullptr ) { lower->iv[1][0] = lo; lower->iv[1][1] = hi; }     // seen from the lower rank I am its side-1 neighbour
  if( higher != nullptr ) { higher->iv[0][0] = lo; higher->iv[0][1] = hi; }
  __threadfence_system();
  if( lower != nullptr ) { st_release_sys( &lower->iv_flag[1], step ); }
  if( higher != nullptr ) { st_release_sys( &higher->iv_flag[0], step ); }
}

static __global__ void __launch_bounds__( 256 ) k_iota_u32( const uint32_t n, const uint32_t first, uint32_t* __restrict__ out )
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if( k < n ) { out[k] = first + k; }
}


// ---- host side: the mailbox / neighbour mappings / candidate state of one slab (used by the rigidbody3d slab mode; ball2d keeps the
// same members inside its own data block) ----------------------------------------------------------------------------------------
struct SlabComm
{
  bool on = false;              // slab mode
  uint32_t n_owned = 0, ghost_cap = 0;
  double xlim[2] = { -1.0e308, 1.0e308 };
  DevBuf gid;                   // u32 per slot: global body index
  DevBuf interval_enc;          // 2 x long long: ordered encoding of [min lo.x, max hi.x] over the owned boxes
  DevBuf ghost_counts;          // u32[4]: ghosts on side 0 / 1, halo overflow, x-limit violation
  DevBuf mailbox;
  void* peer_mb[2] = { nullptr, nullptr };
  bool peer_ipc[2] = { false, false };
  uint32_t step = 0;
  bool scan_done = false;       // this step's interval / candidates were produced
  DevBuf cand_state, cand_list;
  uint32_t cand_cap = 0;
  uint32_t rec_bytes = 0;
  void release()
  {
    for( int sd = 0; sd < 2; ++sd ) { if( peer_mb[sd] != nullptr && peer_ipc[sd] ) { cudaIpcCloseMemHandle( peer_mb[sd] ); } peer_mb[sd] = nullptr; peer_ipc[sd] = false; }
    gid.release(); interval_enc.release(); ghost_counts.release(); mailbox.release(); cand_state.release(); cand_list.release();
  }
  SlabCand cand() const
  {
    SlabCand sc;
    sc.band = nullptr; sc.count = nullptr; sc.list[0] = sc.list[1] = nullptr; sc.cap = 0u; sc.on[0] = sc.on[1] = false;
    if( mailbox.ptr == nullptr || cand_state.ptr == nullptr ) { return sc; }
    SlabCandState* st = cand_state.as<SlabCandState>();
    sc.band = st->band; sc.count = st->count;
    sc.list[0] = cand_list.as<uint32_t>(); sc.list[1] = cand_list.as<uint32_t>() + cand_cap;
    sc.cap = cand_cap;
    sc.on[0] = peer_mb[0] != nullptr; sc.on[1] = peer_mb[1] != nullptr;
    return sc;
  }
};

static inline int slab_comm_mailbox( sg_ctx* ctx, SlabComm& c, void** mailbox_dev, void* ipc_handle_64 )
{
  static_assert( sizeof( cudaIpcMemHandle_t ) == 64, "the ABI passes IPC handles as 64 opaque bytes" );
  if( c.mailbox.ptr == nullptr )
  {
    const size_t bytes = sizeof( SlabMailboxHdr ) + 2 * ( size_t( c.ghost_cap ) + 1 ) * c.rec_bytes;
    SG_CUDA( ctx, c.mailbox.ensure( bytes ) );
    SG_CUDA( ctx, cudaMemsetAsync( c.mailbox.ptr, 0, bytes, ctx->stream ) );
    c.cand_cap = 4u * c.ghost_cap + 1024u;
    SG_CUDA( ctx, c.cand_state.ensure( sizeof( SlabCandState ) ) );
    SG_CUDA( ctx, c.cand_list.ensure( 2 * size_t( c.cand_cap ) * 4 ) );
    k_slab_cand_reset<<<1, 1, 0, ctx->stream>>>( c.cand_state.as<SlabCandState>() );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    c.step = 0;
  }
  if( mailbox_dev != nullptr ) { *mailbox_dev = c.mailbox.ptr; }
  if( ipc_handle_64 != nullptr )
  {
    cudaIpcMemHandle_t h;
    SG_CUDA( ctx, cudaIpcGetMemHandle( &h, c.mailbox.ptr ) );
    memcpy( ipc_handle_64, &h, 64 );
  }
  return SG_OK;
}

static inline int slab_comm_connect( sg_ctx* ctx, SlabComm& c, const int side, const void* ipc_handle_64, void* same_process_mailbox, const int peer_device )
{
  if( c.peer_mb[side] != nullptr && c.peer_ipc[side] ) { cudaIpcCloseMemHandle( c.peer_mb[side] ); }
  c.peer_mb[side] = nullptr;
  if( ipc_handle_64 != nullptr )
  {
    cudaIpcMemHandle_t h;
    memcpy( &h, ipc_handle_64, 64 );
    void* p = nullptr;
    SG_CUDA( ctx, cudaIpcOpenMemHandle( &p, h, cudaIpcMemLazyEnablePeerAccess ) );
    c.peer_mb[side] = p; c.peer_ipc[side] = true;
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
    c.peer_mb[side] = same_process_mailbox; c.peer_ipc[side] = false;
  }
  return SG_OK;
}

static inline int slab_comm_disconnect( sg_ctx* ctx, SlabComm& c )
{
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  for( int sd = 0; sd < 2; ++sd )
  {
    if( c.peer_mb[sd] != nullptr && c.peer_ipc[sd] ) { cudaIpcCloseMemHandle( c.peer_mb[sd] ); }
    c.peer_mb[sd] = nullptr; c.peer_ipc[sd] = false;
  }
  c.mailbox.release();
  cudaGetLastError();
  return SG_OK;
}

#endif
