// sg_broadphase.cuh -- uniform-grid broad phase (+ fused narrow phase) shared by every body type.
//
// Replaces SpatialGridDetector::getPotentialOverlaps (ball2d/SpatialGridDetector.cpp:106-133,
// rigidbody2d/SpatialGrid.cpp:114-141, rigidbody3d/SpatialGridDetector.cpp:110-137) and the per-pair
// loop that follows it in each sim (e.g. ball2d/Ball2DSim.cpp:580-607).
//
// Contract kept from the reference: the candidate set is EXACTLY
//   { (i<j) : for every axis  !(hi_i < lo_j) && !(hi_j < lo_i) }   in ascending (i,j) order,
// nothing else about the grid is observable.  Pipeline (one launch each):
//   bounds   reduce min/max of box lower corners and the largest box extent (fused into the caller's first kernel
//            where there is one: k_ball2d_prep)
//   hist     lay out the grid (h >= largest extent, cell count capped); cell key per body + rank inside the cell
//            (one atomicAdd per body; the histogram is re-zeroed by the scatter of the step before)
//   scan     cell counts -> cell start offsets               (one cooperative launch, sg_scan.cuh)
//   scatter  write a 64-byte record per body at cell_start[key] + rank  => bodies sorted by cell
//   pass 1   (count) per body: walk the 3^D neighbourhood as 3^(D-1) contiguous row segments, AABB test (+ narrow
//            test) against partners with a larger index; leaves counts BY BODY INDEX and, by sorted position, 64-bit
//            candidate / active masks over the visit sequence plus the walk plan.  2-D: float-box prefilter, warp-balanced
//            exact tests (sg_bp_count_l1); 3-D: records through L1/L2, cell ranges staged (sg_bp_count)
//   scan     counts -> output offsets in body-index order, scattered to sorted-position order (cooperative launch)
//   pass 2   (emit) no shared memory, no barriers: set mask bits -> partner positions via the plan -> partner indices
//            -> register sorting network -> candidate pairs at the body's offset + a (p,q) work item per active pair
//            => both lists come out in ascending (i,j) order with no sort of pairs
//   pass 3   (contacts, policies with a fused narrow phase) one thread per active pair, coalesced SoA contact stores
#ifndef SG_BROADPHASE_CUH
#define SG_BROADPHASE_CUH

#include "sg_common.cuh"
#include "sg_scan.cuh"
#include "sg_tma.cuh"
#include "sg_ccd.h"

#include <cstdlib>
#include <cuda.h> // CUtensorMap (type only; the encoder is fetched at run time, see sg_bp_encode_recs_map)

#define SG_BP_THREADS 256
#define SG_BP_LOCAL_CAP 12
#define SG_BP_BATCH 24

// Per-pipeline device scratch (owned by the caller's data block)
struct BroadScratch
{
  DevBuf bounds;        // BoundsAccum
  DevBuf params;        // GridParams
  DevBuf cell_count;    // u32[max_cells + 1]
  DevBuf cell_start;    // u32[max_cells + 1]
  DevBuf cell_partials; // u32[tiles]
  DevBuf key;           // u32[n]
  DevBuf rank;          // u32[n]
  DevBuf recs;          // Rec[n]
  DevBuf sidx;          // u32[n]  ORDER word of each sorted record: the number that ranks bodies in the emitted lists and is written into them
                        //         (the body index on one GPU, the global body index in slab mode)
  DevBuf pos_of;        // u32[n]  sorted position of body i (0xffffffff: unused slot); pass 2 runs in index order and finds its body's masks through it
  DevBuf counts;        // uint2[n]  by body index
  DevBuf masks;         // uint4[n][1 + NPLAN]  by sorted position: candidate / active masks, then where each window of the body's walk starts and how long it is
  DevBuf offsets;       // ulonglong2[n]  by body index
  DevBuf pair_partials; // ScanPairCounts::Acc[tiles]
  DevBuf totals;        // ScanPairCounts::Acc
  DevBuf cand;          // uint2[cand_cap]
  DevBuf work;          // uint2[work_cap]  (own body index, partner's sorted position) of every active pair, in output order
  uint64_t work_cap = 0;
  uint64_t cand_cap = 0;
  uint32_t max_cells = 0;
  SideScan side = { nullptr, 0u, nullptr, nullptr }; // optional small scan carried by the pair-count scan launch (set per step by the caller)
  DevBuf boxf;          // float4[n]  by sorted position (2-D pipelines): the body's box rounded outward to floats -- what pass 1's walk tests
  CUtensorMap tm_recs;  // tensor map over the sorted records (dense-scene pass 1), valid for ( tm_ptr, tm_rows )
  const void* tm_ptr = nullptr;
  uint32_t tm_rows = 0;
  void* tm_encode = nullptr;  // cuTensorMapEncodeTiled, fetched through the runtime on first use
  int staged_attr_dev = -1;   // device on which the staged kernel's shared-memory opt-in was made
  bool dense = false;         // the previous step on this scratch found >= 2.5 candidates per body (set by the caller): pass 1 stages the records
  const uint32_t* ord_by_index = nullptr; // order word of body i (multi-GPU: the global-index table; nullptr: i itself), set per step by the caller
  bool hist_clean = false; // cell_count is all zero (true after every scatter; false after (re)allocation or an aborted step)
  const void* hist_ptr = nullptr;
  uint32_t hist_slots = 0;
  int bounds_phase = 0; // which of the two BoundsAccum the current step reduces into
  BoundsAccum* bounds_cur() const { return bounds.as<BoundsAccum>() + bounds_phase; }
  void release()
  {
    bounds.release(); params.release(); cell_count.release(); cell_start.release(); cell_partials.release(); key.release(); rank.release();
    recs.release(); sidx.release(); boxf.release(); pos_of.release(); counts.release(); masks.release(); offsets.release(); pair_partials.release(); totals.release(); cand.release(); work.release();
  }
};

// ---- record I/O: 64-byte records moved as four 128-bit accesses ------------------------------------
template<typename Rec>
__device__ inline Rec sg_load_rec_global( const Rec* __restrict__ p )
{
  static_assert( sizeof( Rec ) == 64, "records are 64 bytes" );
  union { Rec r; int4 v[4]; } u;
  const int4* src = reinterpret_cast<const int4*>( p );
  u.v[0] = __ldg( src + 0 ); u.v[1] = __ldg( src + 1 ); u.v[2] = __ldg( src + 2 ); u.v[3] = __ldg( src + 3 );
  return u.r;
}
template<typename Rec>
__device__ inline Rec sg_load_rec_shared( const Rec* p )
{
  union { Rec r; int4 v[4]; } u;
  const int4* src = reinterpret_cast<const int4*>( p );
  u.v[0] = src[0]; u.v[1] = src[1]; u.v[2] = src[2]; u.v[3] = src[3];
  return u.r;
}
// Staged windows keep the 64-byte records as four 16-byte chunks with chunk c of slot s stored at chunk
// position c ^ ((s >> 1) & 3) -- the 64B-swizzle pattern -- so that a warp whose lanes read the same chunk of
// consecutive slots touches all 32 banks instead of 8.
template<typename Rec>
__device__ __forceinline__ Rec sg_load_rec_swizzled( const unsigned char* win, const uint32_t slot )
{
  union { Rec r; int4 v[4]; } u;
  const int4* src = reinterpret_cast<const int4*>( win ) + size_t( slot ) * 4;
  const uint32_t sw = ( slot >> 1 ) & 3u;
  u.v[0] = src[0u ^ sw]; u.v[1] = src[1u ^ sw]; u.v[2] = src[2u ^ sw]; u.v[3] = src[3u ^ sw];
  return u.r;
}
template<typename Rec>
__device__ inline void sg_store_rec( Rec* p, const Rec& r )
{
  union { Rec r; int4 v[4]; } u;
  u.r = r;
  int4* dst = reinterpret_cast<int4*>( p );
  dst[0] = u.v[0]; dst[1] = u.v[1]; dst[2] = u.v[2]; dst[3] = u.v[3];
}

// ---- bounds ----------------------------------------------------------------------------------------
__device__ inline void sg_bp_bounds_reset( BoundsAccum* acc )
{
  for( int k = 0; k < 3; ++k )
  {
    acc->min_lo[k] = sg_ordered_from_double( __longlong_as_double( 0x7ff0000000000000LL ) );  // +inf
    acc->max_lo[k] = sg_ordered_from_double( __longlong_as_double( 0xfff0000000000000LL ) );  // -inf
  }
  acc->max_ext = sg_ordered_from_double( 0.0 );
}

static __global__ void sg_bp_bounds_init( BoundsAccum* acc )
{
  if( threadIdx.x == 0 && blockIdx.x == 0 ) { sg_bp_bounds_reset( acc ); sg_bp_bounds_reset( acc + 1 ); }
}

// Monotone float <-> u32 map (the same trick as the ordered doubles) for the single-instruction warp reductions
__device__ __forceinline__ uint32_t sg_ordered_from_float( const float f )
{
  const uint32_t u = __float_as_uint( f );
  return ( u & 0x80000000u ) ? ~u : ( u | 0x80000000u );
}
__device__ __forceinline__ float sg_float_from_ordered( const uint32_t o )
{
  return __uint_as_float( ( o & 0x80000000u ) ? ( o & 0x7fffffffu ) : ~o );
}
// min / max over the warp of a double, rounded OUTWARD to float first so that REDUX (one instruction) can do the
// work of five 64-bit shuffle + compare rounds.  The grid only needs origin <= every lower corner, the far corner
// >= every lower corner and h >= every extent (the candidate set does not depend on the grid), so outward rounding
// by one float ulp is free.  Coordinates beyond the float range would round to +-inf and are not supported.
__device__ __forceinline__ double sg_warp_min_outward( const double x )
{
  return double( sg_float_from_ordered( __reduce_min_sync( 0xffffffffu, sg_ordered_from_float( __double2float_rd( x ) ) ) ) );
}
__device__ __forceinline__ double sg_warp_max_outward( const double x )
{
  return double( sg_float_from_ordered( __reduce_max_sync( 0xffffffffu, sg_ordered_from_float( __double2float_ru( x ) ) ) ) );
}

// Block-level reduction of per-thread partial bounds, then one atomic per quantity per block.
template<int D>
__device__ inline void sg_bp_bounds_commit( double* mn, double* mx, double ext, BoundsAccum* __restrict__ acc )
{
  #pragma unroll
  for( int k = 0; k < D; ++k )
  {
    mn[k] = sg_warp_min_outward( mn[k] );
    mx[k] = sg_warp_max_outward( mx[k] );
  }
  ext = sg_warp_max_outward( ext );
  __shared__ double s_red[SG_BP_THREADS / 32][2 * D + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if( lane == 0 )
  {
    #pragma unroll
    for( int k = 0; k < D; ++k ) { s_red[warp][k] = mn[k]; s_red[warp][D + k] = mx[k]; }
    s_red[warp][2 * D] = ext;
  }
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    for( int w = 1; w < SG_BP_THREADS / 32; ++w )
    {
      for( int k = 0; k < D; ++k ) { mn[k] = fmin( mn[k], s_red[w][k] ); mx[k] = fmax( mx[k], s_red[w][D + k] ); }
      ext = fmax( ext, s_red[w][2 * D] );
    }
    for( int k = 0; k < D; ++k )
    {
      atomicMin( &acc->min_lo[k], sg_ordered_from_double( mn[k] ) );
      atomicMax( &acc->max_lo[k], sg_ordered_from_double( mx[k] ) );
    }
    atomicMax( &acc->max_ext, sg_ordered_from_double( ext ) );
  }
}

template<int D>
__device__ __forceinline__ void sg_bp_bounds_update( const double* lo, const double* hi, double* mn, double* mx, double& ext )
{
  #pragma unroll
  for( int k = 0; k < D; ++k )
  {
    mn[k] = fmin( mn[k], lo[k] );
    mx[k] = fmax( mx[k], lo[k] );
    ext = fmax( ext, hi[k] - lo[k] );
  }
}

template<typename P>
__global__ void __launch_bounds__( SG_BP_THREADS ) sg_bp_bounds( const typename P::In in, BoundsAccum* __restrict__ acc )
{
  constexpr int D = P::D;
  double mn[D], mx[D], ext = 0.0;
  #pragma unroll
  for( int k = 0; k < D; ++k ) { mn[k] = __longlong_as_double( 0x7ff0000000000000LL ); mx[k] = __longlong_as_double( 0xfff0000000000000LL ); }
  for( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < in.n; i += gridDim.x * blockDim.x )
  {
    if( !P::valid( in, i ) ) { continue; }
    double lo[D], hi[D];
    P::load_aabb( in, i, lo, hi );
    sg_bp_bounds_update<D>( lo, hi, mn, mx, ext );
  }
  sg_bp_bounds_commit<D>( mn, mx, ext, acc );
}

// ---- grid layout -----------------------------------------------------------------------------------
template<int D>
__device__ inline void sg_cell_of( const GridParams& g, const double* lo, uint32_t* c )
{
  #pragma unroll
  for( int k = 0; k < D; ++k )
  {
    // lo >= origin exactly (origin is the minimum over the same values); unsigned cast == floor
    uint32_t v = uint32_t( ( lo[k] - g.origin[k] ) / g.h );
    c[k] = ( v < g.dims[k] ) ? v : ( g.dims[k] - 1u );
  }
}

template<int D>
__device__ inline GridParams sg_layout_grid( const BoundsAccum& acc, const uint32_t max_cells )
{
  GridParams g;
  double span[3] = { 0.0, 0.0, 0.0 };
  for( int k = 0; k < 3; ++k ) { g.origin[k] = 0.0; g.dims[k] = 1u; }
  for( int k = 0; k < D; ++k )
  {
    g.origin[k] = sg_double_from_ordered( acc.min_lo[k] );
    span[k] = sg_double_from_ordered( acc.max_lo[k] ) - g.origin[k];
  }
  const double max_ext = sg_double_from_ordered( acc.max_ext );
  // strictly larger than every extent, by a margin that dwarfs the rounding of (lo - origin) / h
  double h = max_ext * ( 1.0 + 9.5367431640625e-07 );
  if( !( h > 0.0 ) ) { h = 1.0; }
  for( int iter = 0; iter < 64; ++iter )
  {
    double cells = 1.0;
    bool ok = true;
    for( int k = 0; k < D; ++k )
    {
      const double dk = floor( span[k] / h ) + 1.0;
      if( !( dk < 4.0e9 ) ) { ok = false; }
      cells *= dk;
    }
    if( ok && cells <= double( max_cells ) ) { break; }
    // too many cells for the scratch arrays: coarsen (always legal, only costs extra box tests)
    const double ratio = ok ? cells / double( max_cells ) : 1.0e6;
    h *= 1.0009765625 * ( ( D == 2 ) ? sqrt( ratio ) : cbrt( ratio ) );
  }
  g.h = h;
  uint32_t ncells = 1u;
  for( int k = 0; k < D; ++k )
  {
    g.dims[k] = uint32_t( span[k] / h ) + 1u;
    ncells *= g.dims[k];
  }
  g.ncells = ncells;
  return g;
}

template<int D>
__device__ inline uint32_t sg_key_of( const GridParams& g, const uint32_t* c )
{
  uint32_t key = c[0] + g.dims[0] * c[1];
  if( D == 3 ) { key += g.dims[0] * g.dims[1] * c[2]; }
  return key;
}

// the outward-rounded float box of a double box: { lo.x, lo.y, hi.x, hi.y }
__device__ __forceinline__ float4 sg_box_outward( const double* lo, const double* hi )
{
  return make_float4( __double2float_rd( lo[0] ), __double2float_rd( lo[1] ), __double2float_ru( hi[0] ), __double2float_ru( hi[1] ) );
}

// ---- histogram / scatter ("single-digit radix sort" keyed by cell) ---------------------------------
template<typename P>
__global__ void __launch_bounds__( SG_BP_THREADS ) sg_bp_hist( const typename P::In in, const BoundsAccum* __restrict__ acc, BoundsAccum* __restrict__ acc_next, const uint32_t max_cells,
                                                              GridParams* __restrict__ params, uint32_t* __restrict__ cell_count,
                                                              uint32_t* __restrict__ key_out, uint32_t* __restrict__ rank_out, uint2* __restrict__ counts )
{
  constexpr int D = P::D;
  // Every block derives the same grid layout from the reduced bounds (h >= largest extent, cell count capped);
  // block 0 publishes it for the later kernels and re-arms the accumulator the next step will reduce into.
  // The histogram itself was left zeroed by the previous step's scatter.
  __shared__ GridParams g_s;
  if( threadIdx.x == 0 ) { g_s = sg_layout_grid<D>( *acc, max_cells ); if( blockIdx.x == 0 ) { *params = g_s; } }
  if( blockIdx.x == 0 && threadIdx.x == 32 ) { sg_bp_bounds_reset( acc_next ); }
  __syncthreads();
  const GridParams g = g_s;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= in.n ) { return; }
  if( !P::valid( in, i ) )
  {
    // an unused slot (multi-GPU ghost capacity): never enters the grid, owns no pairs
    key_out[i] = 0xffffffffu;
    counts[i] = make_uint2( 0u, 0u );
    return;
  }
  double lo[D], hi[D];
  P::load_aabb( in, i, lo, hi );
  uint32_t c[D];
  sg_cell_of<D>( g, lo, c );
  const uint32_t key = sg_key_of<D>( g, c );
  key_out[i] = key;
  rank_out[i] = atomicAdd( &cell_count[key], 1u );
}

template<typename P>
__global__ void __launch_bounds__( SG_BP_THREADS ) sg_bp_scatter( const typename P::In in, const GridParams* __restrict__ params, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ key_in,
                                                                 const uint32_t* __restrict__ rank_in, typename P::Rec* __restrict__ recs, uint32_t* __restrict__ pos_of,
                                                                 uint32_t* __restrict__ cell_count, const uint32_t cell_slots )
{
  // A record goes to a random place, so what counts is how many memory requests it takes: a thread storing its own record issues four
  // 16-byte writes to four different instructions' worth of scattered lines (each 32-byte sector is written in two halves).  Instead the
  // warp passes its 32 records through shared memory and four lanes store one record together: every store instruction writes eight
  // whole 64-byte records, a quarter of the requests and only full sectors.
  __shared__ int4 s_rec[SG_BP_THREADS][4];
  __shared__ uint32_t s_pos[SG_BP_THREADS];
  const uint32_t g_dims0 = params->dims[0], g_dims1 = params->dims[1];
  // the histogram has been consumed by the scan: leave it zeroed for the next step (grid-stride, coalesced)
  for( uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < cell_slots; c += gridDim.x * blockDim.x ) { cell_count[c] = 0u; }
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t pos = 0xffffffffu;
  if( i < in.n )
  {
    const uint32_t key = key_in[i];
    if( key != 0xffffffffu )
    {
      pos = __ldg( &cell_start[key] ) + rank_in[i];
      // cell coordinates travel with the record (no integer division in the pair kernels)
      const uint32_t yz = key / g_dims0;
      union { typename P::Rec r; int4 v[4]; } u;
      u.r = P::make_rec( in, i, key, ( P::D == 3 ) ? yz % g_dims1 : yz, ( P::D == 3 ) ? yz / g_dims1 : 0u );
      const uint32_t sw = ( threadIdx.x >> 1 ) & 3u; // chunk c of slot t at c ^ ((t >> 1) & 3): conflict-free 128-bit accesses both ways
      #pragma unroll
      for( uint32_t c = 0u; c < 4u; ++c ) { s_rec[threadIdx.x][c ^ sw] = u.v[c]; }
    }
    pos_of[i] = pos;
  }
  s_pos[threadIdx.x] = pos;
  __syncwarp();
  const uint32_t lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
  #pragma unroll
  for( uint32_t round = 0u; round < 4u; ++round )
  {
    const uint32_t t = wbase + round * 8u + ( lane >> 2 ); // the record this lane helps to store
    const uint32_t c = lane & 3u;                          // ... and which 16 bytes of it
    const uint32_t tp = s_pos[t];
    if( tp != 0xffffffffu ) { reinterpret_cast<int4*>( recs + tp )[c] = s_rec[t][c ^ ( ( t >> 1 ) & 3u )]; }
  }
}

// The dense by-position side arrays of the sorted records, written in one coalesced pass (the scatter's writes are random:
// it moves the 64-byte records only): the ORDER word of every record and, for the 2-D pipelines, its box rounded outward to floats.
template<typename P>
__global__ void __launch_bounds__( SG_BP_THREADS ) sg_bp_side_arrays( const uint32_t n_slots, const GridParams* __restrict__ params, const uint32_t* __restrict__ cell_start, const typename P::Rec* __restrict__ recs,
                                                                      uint32_t* __restrict__ sidx, float4* __restrict__ boxf, const uint32_t* __restrict__ ord_by_index )
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = min( n_slots, __ldg( &cell_start[params->ncells] ) ); // bodies actually binned
  if( p >= n ) { return; }
  const typename P::Rec r = sg_load_rec_global<typename P::Rec>( &recs[p] );
  sidx[p] = ( ord_by_index != nullptr ) ? __ldg( &ord_by_index[P::rec_idx( r )] ) : P::rec_ord_raw( r );
  if( P::D == 2 && boxf != nullptr )
  {
    double lo[P::D], hi[P::D];
    P::rec_aabb( r, lo, hi );
    boxf[p] = sg_box_outward( lo, hi );
  }
}

// ---- neighbourhood staging + walk ------------------------------------------------------------------
// Bodies are sorted by row-major cell key, so for a block owning sorted positions [b0,b1) with first/last
// keys kf/kl, everything its threads can visit in the row offset (dy,dz) lies in ONE contiguous range of
// cells [kf + off - 1, kl + off + 1], off = dy*dimx + dz*dimx*dimy, hence in one contiguous range of records.
// There are 3^(D-1) such windows.  Per block and window two things are staged in shared memory with coalesced
// loads: the records (pass 1: all 64 bytes, bank-swizzled; pass 2: only the 4-byte index words) and the slice of
// cell_start covering the window's cells -- so the per-thread walk touches no global memory at all.  A window with
// more than WCAP records / CSCAP cells is staged up to the cap and the tail is read through L1/L2.
// (The copy is done by the block's threads rather than by 1-D cp.async.bulk -- helpers in sg_tma.cuh -- because the
// bulk copy cannot apply the swizzle that removes the 4-way bank conflicts of 64-byte records.)
template<int D> struct BpCfg;
template<> struct BpCfg<2> { static constexpr int NW = 3; static constexpr int T = 256; static constexpr int WCAP = 272; static constexpr int CSCAP = 320; static constexpr int FAST = 8; };
template<> struct BpCfg<3> { static constexpr int NW = 9; static constexpr int T = 256; static constexpr int WCAP = 16; static constexpr int CSCAP = 320; static constexpr int FAST = 16; };

template<int D>
struct BpStage
{
  uint32_t start[BpCfg<D>::NW];   // first sorted position of the window
  uint32_t len[BpCfg<D>::NW];     // records staged
  uint32_t cs_klo[BpCfg<D>::NW];  // first cell key whose cell_start entry is staged
  uint32_t cs_len[BpCfg<D>::NW];  // cell_start entries staged
  uint32_t full;                  // every window and every cell_start slice of the tile is staged completely
};

template<int D> __host__ __device__ constexpr size_t sg_bp_cs_bytes() { return size_t( BpCfg<D>::NW ) * BpCfg<D>::CSCAP * 4; }
template<int D> __host__ __device__ constexpr size_t sg_bp_count_smem() { return ( size_t( BpCfg<D>::NW ) * BpCfg<D>::WCAP * 64 + sg_bp_cs_bytes<D>() + sizeof( BpStage<D> ) + 127 ) & ~size_t( 127 ); }

// All threads of the block call this; returns once the staged windows are readable.
// FULL: s_data receives swizzled 64-byte records; otherwise the u32 index words from sidx.
template<typename P, bool FULL>
__device__ inline void sg_bp_stage( const GridParams& g, const uint32_t n, const uint32_t* __restrict__ cell_start, const typename P::Rec* __restrict__ recs,
                                    const uint32_t* __restrict__ sidx, unsigned char* s_data, uint32_t* s_cs, BpStage<P::D>* st )
{
  constexpr int D = P::D;
  using Cfg = BpCfg<D>;
  if( threadIdx.x < uint32_t( Cfg::NW ) )
  {
    const uint32_t w = threadIdx.x;
    const uint32_t b0 = blockIdx.x * Cfg::T;
    const uint32_t b1 = ( n - b0 < uint32_t( Cfg::T ) ) ? n : b0 + Cfg::T;
    const long long kf = __ldg( &recs[b0].key );
    const long long kl = __ldg( &recs[b1 - 1u].key );
    uint32_t start = 0u, len = 0u, cs_klo = 0u, cs_len = 0u;
    const int dy = int( w % 3u ) - 1;
    const int dz = ( D == 3 ) ? int( w / 3u ) - 1 : 0;
    const long long off = ( long long )( dy ) * g.dims[0] + ( long long )( dz ) * g.dims[0] * g.dims[1];
    long long klo = kf + off - 1, khi = kl + off + 1;
    if( khi >= 0 && klo <= ( long long )( g.ncells ) - 1 )
    {
      klo = ( klo < 0 ) ? 0 : klo;
      khi = ( khi > ( long long )( g.ncells ) - 1 ) ? ( long long )( g.ncells ) - 1 : khi;
      start = __ldg( &cell_start[klo] );
      const uint32_t end = __ldg( &cell_start[khi + 1] );
      len = ( end - start < uint32_t( Cfg::WCAP ) ) ? end - start : uint32_t( Cfg::WCAP );
      cs_klo = uint32_t( klo );
      const long long ncs = khi - klo + 2; // entries klo .. khi+1
      cs_len = ( ncs < ( long long )( Cfg::CSCAP ) ) ? uint32_t( ncs ) : uint32_t( Cfg::CSCAP );
    }
    st->start[w] = start; st->len[w] = len; st->cs_klo[w] = cs_klo; st->cs_len[w] = cs_len;
    if( w == 0u ) { st->full = 0u; }
  }
  __syncthreads();
  #pragma unroll
  for( int w = 0; w < Cfg::NW; ++w )
  {
    const uint32_t len = st->len[w];
    if( FULL )
    {
      const uint32_t nchunks = len * 4u;
      const int4* src = reinterpret_cast<const int4*>( recs + st->start[w] );
      int4* dst = reinterpret_cast<int4*>( s_data + size_t( w ) * Cfg::WCAP * 64 );
      for( uint32_t c = threadIdx.x; c < nchunks; c += Cfg::T )
      {
        const uint32_t slot = c >> 2;
        dst[( slot << 2 ) | ( ( c & 3u ) ^ ( ( slot >> 1 ) & 3u ) )] = __ldg( src + c );
      }
    }
    else
    {
      const uint32_t* src = sidx + st->start[w];
      uint32_t* dst = reinterpret_cast<uint32_t*>( s_data ) + w * Cfg::WCAP;
      for( uint32_t c = threadIdx.x; c < len; c += Cfg::T ) { dst[c] = __ldg( src + c ); }
    }
    const uint32_t cs_len = st->cs_len[w];
    const uint32_t* cs_src = cell_start + st->cs_klo[w];
    for( uint32_t c = threadIdx.x; c < cs_len; c += Cfg::T ) { s_cs[w * Cfg::CSCAP + c] = __ldg( cs_src + c ); }
  }
  __syncthreads();
}

// STAGED: the caller knows (BpStage::full) that the lookup cannot miss the staged slice -- no fallback code at all
template<int D, int CSCAP = BpCfg<D>::CSCAP, bool STAGED = false>
__device__ __forceinline__ uint32_t sg_bp_cs( const BpStage<D>* st, const uint32_t* s_cs, const uint32_t* __restrict__ cell_start, const int w, const uint32_t key )
{
  const uint32_t rel = key - st->cs_klo[w];
  if( STAGED ) { return s_cs[w * CSCAP + rel]; }
  return ( rel < st->cs_len[w] ) ? s_cs[w * CSCAP + rel] : __ldg( &cell_start[key] );
}

// [qb[w], qe[w]) = sorted positions of the bodies whose cell is within one cell of (cx,c1,c2) in row window w
template<typename P, int CSCAP = BpCfg<P::D>::CSCAP, bool STAGED = false>
__device__ __forceinline__ void sg_bp_ranges( const GridParams& g, const uint32_t* __restrict__ cell_start, const uint32_t* s_cs, const BpStage<P::D>* st,
                                              const uint32_t key, const uint32_t c1, const uint32_t c2, uint32_t* qb, uint32_t* qe )
{
  constexpr int D = P::D;
  using Cfg = BpCfg<D>;
  const uint32_t cx = key - g.dims[0] * ( c1 + g.dims[1] * c2 );
  const uint32_t x0 = ( cx > 0u ) ? cx - 1u : 0u;
  const uint32_t x1 = ( cx + 1u < g.dims[0] ) ? cx + 1u : cx;
  #pragma unroll
  for( int w = 0; w < Cfg::NW; ++w )
  {
    const int dy = w % 3 - 1;
    const int dz = ( D == 3 ) ? w / 3 - 1 : 0;
    const long long y = ( long long )( c1 ) + dy;
    const long long z = ( long long )( c2 ) + dz;
    const bool ok = y >= 0 && y < ( long long )( g.dims[1] ) && z >= 0 && z < ( long long )( g.dims[2] );
    qb[w] = 0u; qe[w] = 0u;
    if( ok )
    {
      const uint32_t row = g.dims[0] * ( uint32_t( y ) + g.dims[1] * uint32_t( z ) );
      qb[w] = sg_bp_cs<D, CSCAP, STAGED>( st, s_cs, cell_start, w, row + x0 );
      qe[w] = sg_bp_cs<D, CSCAP, STAGED>( st, s_cs, cell_start, w, row + x1 + 1u );
    }
  }
}

// Calls f( w, q ) for every sorted position q != p of the ranges, in a fixed order (window-major, position
// ascending): this is THE visit sequence -- bit k of the pass-1 masks is the k-th call.
template<typename P, typename F>
__device__ __forceinline__ void sg_bp_walk_ranges( const uint32_t* qb, const uint32_t* qe, const uint32_t p, F&& f )
{
  #pragma unroll
  for( int w = 0; w < BpCfg<P::D>::NW; ++w )
  {
    for( uint32_t q = qb[w]; q < qe[w]; ++q )
    {
      if( q != p ) { f( w, q ); }
    }
  }
}

template<typename P, typename F>
__device__ __forceinline__ void sg_bp_walk_pos( const GridParams& g, const uint32_t* __restrict__ cell_start, const uint32_t* s_cs, const BpStage<P::D>* st,
                                                const uint32_t p, const uint32_t key, const uint32_t c1, const uint32_t c2, F&& f )
{
  uint32_t qb[BpCfg<P::D>::NW], qe[BpCfg<P::D>::NW];
  sg_bp_ranges<P>( g, cell_start, s_cs, st, key, c1, c2, qb, qe );
  sg_bp_walk_ranges<P>( qb, qe, p, f );
}

// The walk plan pass 1 leaves for pass 2: per body the window starts and lengths (clipped to 255; a body whose
// walk is longer than 63 visits has incomplete masks anyway), NPLAN uint4 per body.  Masks and plan of sorted position p sit
// side by side -- STRIDE = 1 + NPLAN uint4 words: [masks | plan words] -- so that pass 2, which visits the bodies in INDEX
// order, finds everything it needs about a body in one 32-byte sector (2-D) / two (3-D).  (Measured the other way round --
// pass 1 scattering them by body index so that pass 2 reads them coalesced: on 16 M randomly numbered balls pass 2 gained 0.17 ms
// and pass 1 lost 0.47 ms to the 32-byte random writes.)  `masks` and `plan` below are the same buffer, offset by one word.
template<int D> struct BpPlan;
template<> struct BpPlan<2> { static constexpr int NPLAN = 1; static constexpr int STRIDE = 2; };
template<> struct BpPlan<3> { static constexpr int NPLAN = 3; static constexpr int STRIDE = 4; };

template<int D>
__device__ __forceinline__ void sg_bp_plan_store( uint4* __restrict__ plan, const uint32_t n, const uint32_t p, const uint32_t* qb, const uint32_t* qe )
{
  constexpr int NW = BpCfg<D>::NW;
  uint32_t words[BpPlan<D>::NPLAN * 4];
  #pragma unroll
  for( int k = 0; k < BpPlan<D>::NPLAN * 4; ++k ) { words[k] = 0u; }
  #pragma unroll
  for( int w = 0; w < NW; ++w )
  {
    words[w] = qb[w];
    const uint32_t len = ( qe[w] - qb[w] < 255u ) ? qe[w] - qb[w] : 255u;
    words[NW + w / 4] |= len << ( 8 * ( w % 4 ) );
  }
  #pragma unroll
  for( int c = 0; c < BpPlan<D>::NPLAN; ++c ) { plan[size_t( p ) * BpPlan<D>::STRIDE + c] = make_uint4( words[4 * c], words[4 * c + 1], words[4 * c + 2], words[4 * c + 3] ); }
}

template<int D>
__device__ __forceinline__ void sg_bp_plan_load( const uint4* __restrict__ plan, const uint32_t n, const uint32_t p, uint32_t* qb, uint32_t* len )
{
  constexpr int NW = BpCfg<D>::NW;
  uint32_t words[BpPlan<D>::NPLAN * 4];
  #pragma unroll
  for( int c = 0; c < BpPlan<D>::NPLAN; ++c )
  {
    const uint4 v = __ldg( &plan[size_t( p ) * BpPlan<D>::STRIDE + c] );
    words[4 * c] = v.x; words[4 * c + 1] = v.y; words[4 * c + 2] = v.z; words[4 * c + 3] = v.w;
  }
  #pragma unroll
  for( int w = 0; w < NW; ++w ) { qb[w] = words[w]; len[w] = ( words[NW + w / 4] >> ( 8 * ( w % 4 ) ) ) & 255u; }
}

// record (or just its body index) at sorted position q, known to lie in window w's key range
template<typename P, bool STAGED = false>
__device__ __forceinline__ typename P::Rec sg_bp_fetch( const typename P::Rec* __restrict__ recs, const unsigned char* s_recs, const BpStage<P::D>* st, const int w, const uint32_t q )
{
  using Rec = typename P::Rec;
  const uint32_t slot = q - st->start[w];
  if( STAGED || slot < st->len[w] ) { return sg_load_rec_swizzled<Rec>( s_recs + size_t( w ) * BpCfg<P::D>::WCAP * 64, slot ); }
  return P::load_pass1( &recs[q] ); // what pass 1 needs of a partner (a policy may keep that in the record's first half)
}
// ORDER word of the body at sorted position q (window w): what decides which body of a pair owns it and how a body's partners are
// ranked.  Policies keep it inside the record (P::ORD_OFFSET) so that the staged walk never leaves shared memory.
template<typename P, bool STAGED = false>
__device__ __forceinline__ uint32_t sg_bp_fetch_ord( const typename P::Rec* __restrict__ recs, const unsigned char* s_recs, const BpStage<P::D>* st, const int w, const uint32_t q,
                                                     const uint32_t* __restrict__ sidx = nullptr )
{
  constexpr uint32_t CH = P::ORD_OFFSET / 16u, IN = P::ORD_OFFSET % 16u;
  const uint32_t slot = q - st->start[w];
  if( P::ORD_IN_REC && ( STAGED || slot < st->len[w] ) )
  {
    const unsigned char* rec = s_recs + ( size_t( w ) * BpCfg<P::D>::WCAP + slot ) * 64;
    return *reinterpret_cast<const uint32_t*>( rec + ( ( CH ^ ( ( slot >> 1 ) & 3u ) ) << 4 ) + IN ) & P::IDX_MASK;
  }
  // un-staged: the dense order array (4 B/body, eight bodies per sector) rather than one sector of every record visited
  if( sidx != nullptr ) { return __ldg( &sidx[q] ) & P::IDX_MASK; }
  return __ldg( reinterpret_cast<const uint32_t*>( reinterpret_cast<const unsigned char*>( &recs[q] ) + P::ORD_OFFSET ) ) & P::IDX_MASK;
}

// masks cover the first 63 visits of a body's walk; bit 63 of the active mask flags a longer walk (masks incomplete)
#define SG_BP_MASK_BITS 63u
#define SG_BP_MASKS_INVALID 0x8000000000000000ull

// One body's pass-1 walk: candidates with a larger index, how many are active, the two visit masks, the walk plan.
template<typename P, int CSCAP, bool STAGED>
__device__ __forceinline__ void sg_bp_count_body( const GridParams& g, const uint32_t* __restrict__ cell_start, const typename P::Rec* __restrict__ recs, const unsigned char* s_recs, const uint32_t* s_cs,
                                                  const BpStage<P::D>* st, const uint32_t n_slots, const uint32_t p, const typename P::Rec& me, const uint32_t my_idx, const uint32_t my_ord,
                                                  uint2* __restrict__ counts, uint4* __restrict__ masks, uint4* __restrict__ plan, const uint32_t* __restrict__ sidx = nullptr )
{
  constexpr int D = P::D;
  using Cfg = BpCfg<D>;
  using Rec = typename P::Rec;
  double lo[D], hi[D];
  P::rec_aabb( me, lo, hi );
  uint32_t nc = 0u, na = 0u, k = 0u;
  unsigned long long cmask = 0ull, amask = 0ull;
  uint32_t qb[Cfg::NW], qe[Cfg::NW];
  sg_bp_ranges<P, CSCAP, STAGED>( g, cell_start, s_cs, st, P::rec_key( me ), P::rec_c1( me, g ), P::rec_c2( me, g ), qb, qe );
  sg_bp_plan_store<D>( plan, n_slots, p, qb, qe );
  sg_bp_walk_ranges<P>( qb, qe, p, [&]( const int w, const uint32_t q )
  {
    const unsigned long long bit = ( k < SG_BP_MASK_BITS ) ? ( 1ull << k ) : 0ull;
    ++k;
    if( sg_bp_fetch_ord<P, STAGED>( recs, s_recs, st, w, q, sidx ) <= my_ord ) { return; } // owned by the partner: skip before touching the record
    const Rec o = sg_bp_fetch<P, STAGED>( recs, s_recs, st, w, q );
    double olo[D], ohi[D];
    P::rec_aabb( o, olo, ohi );
    bool ov = true;
    #pragma unroll
    for( int a = 0; a < D; ++a ) { ov = ov && !( hi[a] < olo[a] ) && !( ohi[a] < lo[a] ); }
    if( !ov ) { return; }
    ++nc; cmask |= bit;
    if( P::HAS_NARROW ) { if( P::narrow_test( me, o ) ) { ++na; amask |= bit; } }
  } );
  if( k > SG_BP_MASK_BITS ) { amask |= SG_BP_MASKS_INVALID; }
  counts[my_idx] = make_uint2( nc, na );
  masks[size_t( p ) * BpPlan<P::D>::STRIDE] = make_uint4( uint32_t( cmask ), uint32_t( cmask >> 32 ), uint32_t( amask ), uint32_t( amask >> 32 ) );
}

// Pass 1.  counts[body index] = { #candidates with a larger index, #of those that are active }
//          masks[sorted position] = { candidate mask (64 bit), active mask (64 bit) } over the visit sequence
template<typename P>
__global__ void __launch_bounds__( BpCfg<P::D>::T, 4 ) sg_bp_count( const uint32_t n_slots, const GridParams* __restrict__ params, const uint32_t* __restrict__ cell_start,
                                                               const typename P::Rec* __restrict__ recs, const uint32_t* __restrict__ sidx, uint2* __restrict__ counts, uint4* __restrict__ masks,
                                                               uint4* __restrict__ plan )
{
  constexpr int D = P::D;
  using Cfg = BpCfg<D>;
  using Rec = typename P::Rec;
  extern __shared__ __align__( 1024 ) unsigned char s_raw[];
  unsigned char* s_recs = s_raw;
  uint32_t* s_cs = reinterpret_cast<uint32_t*>( s_raw + size_t( Cfg::NW ) * Cfg::WCAP * 64 );
  BpStage<D>* st = reinterpret_cast<BpStage<D>*>( s_raw + size_t( Cfg::NW ) * Cfg::WCAP * 64 + sg_bp_cs_bytes<D>() );
  const GridParams g = *params;
  const uint32_t n = min( n_slots, __ldg( &cell_start[g.ncells] ) ); // bodies actually binned (unused slots are not)
  const uint32_t p = blockIdx.x * Cfg::T + threadIdx.x;
  if( blockIdx.x * Cfg::T >= n ) { return; } // (positions past the binned bodies belong to nobody: pass 2 finds a body's masks through pos_of)
  Rec me;
  if( p < n ) { me = sg_load_rec_global<Rec>( &recs[p] ); } // coalesced; in flight while the windows are staged
  sg_bp_stage<P, true>( g, n, cell_start, recs, nullptr, s_recs, s_cs, st );
  if( p >= n ) { return; }
  const uint32_t my_idx = P::rec_idx( me );
  if( !P::owns( me ) )
  {
    // a ghost body (multi-GPU halo): present only as a partner, its pairs are kept by the rank that owns it
    counts[my_idx] = make_uint2( 0u, 0u );
    masks[size_t( p ) * BpPlan<P::D>::STRIDE] = make_uint4( 0u, 0u, 0u, 0u );
    return;
  }
  const uint32_t my_ord = P::ORD_IN_REC ? P::rec_ord( me ) : ( __ldg( &sidx[p] ) & P::IDX_MASK );
  sg_bp_count_body<P, Cfg::CSCAP, false>( g, cell_start, recs, s_recs, s_cs, st, n_slots, p, me, my_idx, my_ord, counts, masks, plan, sidx );
}

// ---- pass 1, box-prefiltered (D = 2): helpers ------------------------------------------------------------
// What a body's walk over its 3 row windows needs per partner is a box test that almost always fails (config 3: 19 visits,
// 2.7 boxes touched, 1.3 pairs owned).  So the walk does not touch the 64-byte records at all: a coalesced pass after the scatter
// leaves, by sorted position, the body's box rounded OUTWARD to four floats (16 bytes) and its ORDER word; the walk is one 128-bit
// and one 32-bit load and five compares per partner, branch-free, recording the survivors as one bit each.  Only the survivors
// (a conservative superset of the overlapping boxes: outward rounding can add, never drop) get the exact FP64 tests of the
// reference -- AABB::overlaps on the swept boxes, then the narrow phase -- on the full records.
// A body whose walk does not fit the survivor bitmaps (a window of more than 64 positions, or more than 63 visits in all):
// the plain walk with exact tests, everything through L1/L2.  Leaves exact counts and the (incomplete) masks pass 2 expects.
template<typename P>
__device__ __noinline__ void sg_bpx_body_slow( const GridParams* __restrict__ params, const uint32_t* __restrict__ cell_start, const typename P::Rec* __restrict__ recs, const uint32_t* __restrict__ sidx,
                                               const uint32_t n_slots, const uint32_t p, uint2* __restrict__ counts, uint4* __restrict__ masks, uint4* __restrict__ plan )
{
  using Rec = typename P::Rec;
  const GridParams g = *params;
  BpStage<2> none;
  #pragma unroll
  for( int w = 0; w < 3; ++w ) { none.start[w] = 0u; none.len[w] = 0u; none.cs_klo[w] = 0u; none.cs_len[w] = 0u; }
  none.full = 0u;
  const Rec me = sg_load_rec_global<Rec>( &recs[p] );
  sg_bp_count_body<P, BpCfg<2>::CSCAP, false>( g, cell_start, recs, nullptr, nullptr, &none, n_slots, p, me, P::rec_idx( me ), P::rec_ord( me ), counts, masks, plan, sidx );
}

// The exact tests of one owned survivor pair -- AABB::overlaps on the swept boxes, then the narrow phase: bit 0 candidate, bit 1 active
template<typename P>
__device__ __forceinline__ uint32_t sg_bpx_exact( const typename P::Rec& a, const typename P::Rec& b )
{
  double lo[2], hi[2], olo[2], ohi[2];
  P::rec_aabb( a, lo, hi );
  P::rec_aabb( b, olo, ohi );
  const bool ov = !( hi[0] < olo[0] ) && !( ohi[0] < lo[0] ) && !( hi[1] < olo[1] ) && !( ohi[1] < lo[1] );
  if( !ov ) { return 0u; }
  uint32_t res = 1u;
  if( P::HAS_NARROW ) { if( P::narrow_test( a, b ) ) { res |= 2u; } }
  return res;
}

#define SG_BPX_QLANE 6u                       // survivors a lane hands to the warp's queue (the rest it tests itself)
#define SG_BPX_QCAP ( 32u * SG_BPX_QLANE )    // queue entries per warp

// ---- pass 1, box-prefiltered, every warp on its own (D = 2) -----------------------------------------------
// One thread per sorted body, no staging: a body's three row windows are contiguous runs of the dense 16-byte float-box array, and
// the 32 bodies of a warp walk overlapping runs, so the boxes come through L1 at the wavefront cost shared memory would have (a
// 128-bit access of 32 lanes is four wavefronts either way).  Its predecessor staged the windows of a 256-body tile with bulk copies
// (cp.async.bulk, producer warp + 8 consumer warps, two 26 KB stages); measured on the B200 this one is faster on every scene
// (config 2: 80 vs 93 us, config 3 at 2 M: 181 vs 207 us, at 16 M: 1.37 vs 1.65 ms) -- the staged kernel lost a fifth of its time
// to fast warps waiting for the slowest warp of the CTA to release a stage, and both end up bound by the L1 / shared-memory data
// pipe (63 % of its peak) and by load-to-use latency at 32 warps per SM, not by how the boxes get on chip.
// What makes the walk shorter than the staged one:
//  * far-side pruning: a body whose box ends before the next cell column (row) begins cannot touch anything binned there -- bodies
//    are binned by their LOWER corners -- so that column (the whole upper window) is dropped from its ranges.  With radii spread
//    over 1:4 most boxes are much smaller than a cell: 9 cells become 6.5 on average (config 3);
//  * the own box is read from the float-box array like the partners' (one 128-bit load instead of rebuilding it from the record);
//  * the three windows advance together, CH partners of each per round, all loads of a round requested before the first compare;
//  * the exact FP64 box test is skipped for a survivor whose float boxes overlap by more than their rounding (below): what is left
//    per pair is the narrow phase (ball-ball: division-free, sg_ccd.h).
// The survivors of all 32 lanes are compacted into the warp's queue (shared memory, the only use of it) and tested one pair per
// lane, so the FP64 work runs with all lanes busy instead of inside 32 divergent per-lane loops; every lane then collects the
// verdicts of its own pairs into its masks and counts.

// Float boxes are the FP64 boxes rounded OUTWARD: lo_f <= lo < next( lo_f ), prev( hi_f ) < hi <= hi_f.  So hi_A >= lo_B is certain once
// prev( hi_A_f ) >= next( lo_B_f ), which a gap of more than the two spacings guarantees: spacing <= 2^-23 |x| for normal floats, the
// absolute term covers subnormals.  True for all four sides => AABB::overlaps holds for the FP64 boxes, no need to build them.
__device__ __forceinline__ bool sg_box_gap_certain( const float hi, const float lo )
{
  return ( hi - lo ) > 4.0e-7f * ( fabsf( hi ) + fabsf( lo ) ) + 1.0e-37f;
}
__device__ __forceinline__ bool sg_boxes_certainly_overlap( const float4 a, const float4 b )
{
  return sg_box_gap_certain( a.z, b.x ) && sg_box_gap_certain( b.z, a.x ) && sg_box_gap_certain( a.w, b.y ) && sg_box_gap_certain( b.w, a.y );
}

template<typename P, int CH, int MINB>
__global__ void __launch_bounds__( SG_BP_THREADS, MINB ) sg_bp_count_l1( const uint32_t n_slots, const GridParams* __restrict__ params, const uint32_t* __restrict__ cell_start,
                                                                         const typename P::Rec* __restrict__ recs, const float4* __restrict__ boxf, const uint32_t* __restrict__ sidx,
                                                                         uint2* __restrict__ counts, uint4* __restrict__ masks, uint4* __restrict__ plan )
{
  using Rec = typename P::Rec;
  static_assert( P::D == 2, "the box-prefiltered pass 1 is laid out for the 2-D pipelines" );
  __shared__ uint2 s_wq[SG_BP_THREADS / 32][SG_BPX_QCAP];
  uint2* wq = s_wq[threadIdx.x >> 5];
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t p = blockIdx.x * SG_BP_THREADS + threadIdx.x;
  const GridParams g = *params;
  const uint32_t n = min( n_slots, __ldg( &cell_start[g.ncells] ) ); // bodies actually binned
  if( blockIdx.x * SG_BP_THREADS >= n ) { return; } // (positions past the binned bodies belong to nobody: pass 2 finds a body's masks through pos_of)
  bool part = p < n;
  uint32_t my_idx = 0u, my_ord = 0u, key = 0u, c1 = 0u;
  if( part )
  {
    // coalesced; only its integer words are kept.  (The stream of whole records also leaves the tile's records in L2 for the exact
    // tests below: reading the four words from a 16-byte side array instead made this kernel 8 % SLOWER on 16 M bodies.)
    const Rec me = sg_load_rec_global<Rec>( &recs[p] );
    my_idx = P::rec_idx( me ); my_ord = P::rec_ord( me ); key = P::rec_key( me ); c1 = P::rec_c1( me, g );
    if( !P::owns( me ) )
    {
      // a ghost body (multi-GPU halo): present only as a partner, its pairs are kept by the rank that owns it
      counts[my_idx] = make_uint2( 0u, 0u );
      masks[size_t( p ) * BpPlan<2>::STRIDE] = make_uint4( 0u, 0u, 0u, 0u );
      part = false;
    }
  }
  uint32_t qb[3] = { 0u, 0u, 0u }, qe[3] = { 0u, 0u, 0u };
  unsigned long long pm[3] = { 0ull, 0ull, 0ull };
  if( part )
  {
    const float4 mb = __ldg( &boxf[p] ); // { lo.x, lo.y, hi.x, hi.y } rounded outward
    // ---- ranges, pruned on the far side
    const uint32_t cx = key - g.dims[0] * c1;
    const uint32_t x0 = ( cx > 0u ) ? cx - 1u : 0u;
    uint32_t x1 = ( cx + 1u < g.dims[0] ) ? cx + 1u : cx;
    // A body B binned in column >= cx + 1 has fl( fl( lo_B - origin ) / h ) >= cx + 1, hence lo_B >= origin + (cx + 1) h (1 - 2^-51).
    // If this body's upper bound (rounded up to float) lies below that, with a margin far above the rounding of the bound computed
    // here, no box of that column can reach it.
    {
      const double edge = double( cx + 1u ) * g.h;
      if( double( mb.z ) < ( g.origin[0] + edge ) - 1.0e-9 * ( fabs( g.origin[0] ) + edge ) ) { x1 = cx; }
    }
    bool up = c1 + 1u < g.dims[1];
    {
      const double edge = double( c1 + 1u ) * g.h;
      if( double( mb.w ) < ( g.origin[1] + edge ) - 1.0e-9 * ( fabs( g.origin[1] ) + edge ) ) { up = false; }
    }
    #pragma unroll
    for( int w = 0; w < 3; ++w )
    {
      const bool ok = ( w == 0 ) ? ( c1 > 0u ) : ( ( w == 2 ) ? up : true );
      if( ok )
      {
        const uint32_t row = g.dims[0] * ( c1 + uint32_t( w ) - 1u );
        qb[w] = __ldg( &cell_start[row + x0] );
        qe[w] = __ldg( &cell_start[row + x1 + 1u] );
      }
    }
    uint32_t total = 0u;
    bool fits = true;
    #pragma unroll
    for( int w = 0; w < 3; ++w )
    {
      const uint32_t L = qe[w] - qb[w];
      fits = fits && L <= 64u;
      total += L - ( ( p - qb[w] < L ) ? 1u : 0u );
    }
    if( !fits || total > SG_BP_MASK_BITS )
    {
      // does not fit the bitmaps: the plain walk with exact tests (its own, unpruned, ranges and plan)
      sg_bpx_body_slow<P>( params, cell_start, recs, sidx, n_slots, p, counts, masks, plan );
      part = false;
    }
    else
    {
      sg_bp_plan_store<2>( plan, n_slots, p, qb, qe );
      // ---- walk: bit j of pm[w] = partner qb[w] + j survives.  The three windows advance together, CH partners of each per
      // round, boxes AND order words of the whole round requested before the first compare: one memory latency per round
      // instead of two per window and round.  (Both arrays have slack behind their last entry: a round reads up to the longest
      // window's length past a shorter window's end; those lanes are masked out.)
      const uint32_t L0 = qe[0] - qb[0], L1 = qe[1] - qb[1], L2 = qe[2] - qb[2];
      const uint32_t Lmax = max( L0, max( L1, L2 ) );
      unsigned long long m0 = 0ull, m1 = 0ull, m2 = 0ull;
      for( uint32_t j = 0u; j < Lmax; j += uint32_t( CH ) )
      {
        float4 b[3][CH];
        uint32_t o[3][CH];
        #pragma unroll
        for( int w = 0; w < 3; ++w )
        {
          #pragma unroll
          for( int u = 0; u < CH; ++u ) { b[w][u] = __ldg( boxf + qb[w] + j + u ); o[w][u] = __ldg( sidx + qb[w] + j + u ); }
        }
        uint32_t t[3] = { 0u, 0u, 0u };
        #pragma unroll
        for( int w = 0; w < 3; ++w )
        {
          const uint32_t L = ( w == 0 ) ? L0 : ( ( w == 1 ) ? L1 : L2 );
          #pragma unroll
          for( int u = 0; u < CH; ++u )
          {
            const bool pass = !( mb.z < b[w][u].x ) && !( b[w][u].z < mb.x ) && !( mb.w < b[w][u].y ) && !( b[w][u].w < mb.y ) && j + u < L
                              && ( o[w][u] & P::IDX_MASK ) > my_ord; // never the body itself: equal order words
            if( pass ) { t[w] |= 1u << u; }
          }
        }
        m0 |= static_cast<unsigned long long>( t[0] ) << j;
        m1 |= static_cast<unsigned long long>( t[1] ) << j;
        m2 |= static_cast<unsigned long long>( t[2] ) << j;
      }
      pm[0] = m0; pm[1] = m1; pm[2] = m2;
    }
  }
  // ---- compact the survivors of the warp into its queue: ( partner position, visit number | owner lane << 8 )
  const uint32_t cnt = uint32_t( __popcll( pm[0] ) + __popcll( pm[1] ) + __popcll( pm[2] ) ); // 0 for lanes that do not take part
  if( __ballot_sync( 0xffffffffu, cnt != 0u ) == 0u )
  {
    if( part ) { counts[my_idx] = make_uint2( 0u, 0u ); masks[size_t( p ) * BpPlan<2>::STRIDE] = make_uint4( 0u, 0u, 0u, 0u ); }
    return;
  }
  const uint32_t cq = ( cnt < SG_BPX_QLANE ) ? cnt : SG_BPX_QLANE;
  uint32_t off = cq;
  #pragma unroll
  for( int d = 1; d < 32; d <<= 1 ) { const uint32_t v = __shfl_up_sync( 0xffffffffu, off, d ); if( lane >= uint32_t( d ) ) { off += v; } }
  const uint32_t total_q = __shfl_sync( 0xffffffffu, off, 31 );
  off -= cq;
  uint32_t nc = 0u, na = 0u;
  unsigned long long cmask = 0ull, amask = 0ull;
  if( cnt != 0u )
  {
    uint32_t taken = 0u, base = 0u;
    #pragma unroll
    for( int w = 0; w < 3; ++w )
    {
      const uint32_t L = qe[w] - qb[w];
      const bool mine = p - qb[w] < L;
      unsigned long long m = pm[w];
      while( m != 0ull )
      {
        const uint32_t j = uint32_t( __ffsll( static_cast<long long>( m ) ) ) - 1u;
        m &= m - 1ull;
        const uint32_t q = qb[w] + j;
        const uint32_t k = base + j - ( ( mine && q > p ) ? 1u : 0u ); // number of this visit in the body's visit sequence
        if( taken < SG_BPX_QLANE ) { wq[off + taken] = make_uint2( q, k | ( lane << 8 ) ); }
        else
        {
          // more survivors than this lane's share of the queue (dense scenes): tested here
          const uint32_t res = sg_bpx_exact<P>( sg_load_rec_global<Rec>( &recs[p] ), sg_load_rec_global<Rec>( &recs[q] ) );
          if( res & 1u ) { ++nc; cmask |= 1ull << k; }
          if( res & 2u ) { ++na; amask |= 1ull << k; }
        }
        ++taken;
      }
      base += L - ( mine ? 1u : 0u );
    }
  }
  __syncwarp();
  // ---- the tests, one pair per lane
  const uint32_t warp_first = p - lane; // sorted position of lane 0's body
  for( uint32_t e = lane; e < total_q; e += 32u )
  {
    const uint2 ent = wq[e];
    const uint32_t pa = warp_first + ( ( ent.y >> 8 ) & 31u );
    uint32_t res;
    if( sg_boxes_certainly_overlap( __ldg( &boxf[pa] ), __ldg( &boxf[ent.x] ) ) )
    {
      // the usual case: the pair is a candidate for certain, what is left is the narrow phase
      res = 1u;
      if( P::HAS_NARROW ) { if( P::narrow_test( sg_load_rec_global<Rec>( &recs[pa] ), sg_load_rec_global<Rec>( &recs[ent.x] ) ) ) { res |= 2u; } }
    }
    else { res = sg_bpx_exact<P>( sg_load_rec_global<Rec>( &recs[pa] ), sg_load_rec_global<Rec>( &recs[ent.x] ) ); } // boxes within rounding of touching: the FP64 boxes decide
    wq[e].y = ent.y | ( res << 30 );
  }
  __syncwarp();
  // ---- verdicts back to the owners
  for( uint32_t i = 0u; i < cq; ++i )
  {
    const uint32_t y = wq[off + i].y;
    const unsigned long long bit = 1ull << ( y & 63u );
    if( y & 0x40000000u ) { ++nc; cmask |= bit; }
    if( y & 0x80000000u ) { ++na; amask |= bit; }
  }
  if( part )
  {
    counts[my_idx] = make_uint2( nc, na );
    masks[size_t( p ) * BpPlan<2>::STRIDE] = make_uint4( uint32_t( cmask ), uint32_t( cmask >> 32 ), uint32_t( amask ), uint32_t( amask >> 32 ) );
  }
}

// ---- pass 1 for DENSE scenes, TMA-fed (D = 2) ---------------------------------------------------------
// When most box tests pass (a lattice pile: every ball owns four of its eight neighbours) the float-box prefilter of sg_bp_count_l1 prunes
// nothing and its exact phase gathers two 64-byte records per pair through L1; here the records themselves are staged, so the walk and the
// FP64 tests run out of shared memory (config 2: 49 us against 80 us; on the polydisperse gas it is the other way round, 500 against 181 us).
// The launch picks by the previous step's candidates per body (BroadScratch::dense).
// Same work as sg_bp_count, restructured so no thread ever waits for the staging: persistent CTAs (2 per SM),
// each with 8 consumer warps and 1 producer warp.  The producer runs one tile ahead: it reads the tile's first/last
// cell keys and the cell_start entries that bound its three row windows, then asks the TMA unit for the windows --
// the 64-byte records as 2-D tensor copies with the hardware 64B swizzle (the same chunk ^ ((row >> 1) & 3)
// pattern the software staging used), the cell_start slices as 1-D bulk copies -- into the other half of a
// double-buffered shared-memory stage, completion counted in bytes on an mbarrier.  Consumers wait on that
// barrier, walk entirely out of shared memory, and hand the stage back through a second mbarrier.
#define SG_BP_TMA_ROWS 136   // rows per tensor copy (<= 256); two copies cover WCAP = 272
#define SG_BP_TMA_CSCAP 288
template<int D> __host__ __device__ constexpr size_t sg_bp_tma_stage_bytes()
{
  return ( size_t( BpCfg<D>::NW ) * BpCfg<D>::WCAP * 64 + size_t( BpCfg<D>::NW ) * SG_BP_TMA_CSCAP * 4 + sizeof( BpStage<D> ) + 1023 ) & ~size_t( 1023 );
}
#define SG_BP_TMA_STAGES 2      // shared-memory stages per CTA (measured: 1 stage x 4 CTAs/SM is 1.5x slower)
#define SG_BP_TMA_CTAS_PER_SM 2 // resident CTAs per SM (stages x CTAs x 55 KB must fit the SM's 227 KB)
template<int D> __host__ __device__ constexpr size_t sg_bp_tma_smem() { return SG_BP_TMA_STAGES * sg_bp_tma_stage_bytes<D>() + 64; }

template<typename P>
__global__ void __launch_bounds__( BpCfg<P::D>::T + 32, SG_BP_TMA_CTAS_PER_SM ) sg_bp_count_staged( const __grid_constant__ CUtensorMap tm_recs, const uint32_t n_slots, const GridParams* __restrict__ params,
                                                                            const uint32_t* __restrict__ cell_start, const typename P::Rec* __restrict__ recs, uint2* __restrict__ counts,
                                                                            uint4* __restrict__ masks, uint4* __restrict__ plan )
{
  constexpr int D = P::D;
  using Cfg = BpCfg<D>;
  using Rec = typename P::Rec;
  static_assert( D == 2, "the TMA-fed pass 1 is laid out for the 2-D pipelines" );
  static_assert( Cfg::WCAP == 2 * SG_BP_TMA_ROWS, "two tensor copies per window" );
  extern __shared__ __align__( 1024 ) unsigned char s_raw[];
  constexpr size_t STAGE = sg_bp_tma_stage_bytes<D>();
  constexpr size_t CS_OFF = size_t( Cfg::NW ) * Cfg::WCAP * 64;
  constexpr size_t ST_OFF = CS_OFF + size_t( Cfg::NW ) * SG_BP_TMA_CSCAP * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>( s_raw + SG_BP_TMA_STAGES * STAGE ); // full[0], full[1], empty[0], empty[1] (stage s uses full[s], empty[s])
  const GridParams g = *params;
  const uint32_t n = min( n_slots, __ldg( &cell_start[g.ncells] ) ); // bodies actually binned
  const uint32_t ntiles = ( n_slots + Cfg::T - 1u ) / Cfg::T;
  const bool producer = threadIdx.x >= uint32_t( Cfg::T );
  if( threadIdx.x == 0 )
  {
    sg_mbar_init( &bars[0], 1u ); sg_mbar_init( &bars[1], 1u );
    sg_mbar_init( &bars[2], Cfg::T / 32u ); sg_mbar_init( &bars[3], Cfg::T / 32u );
  }
  __syncthreads();

  if( producer )
  {
    if( threadIdx.x != uint32_t( Cfg::T ) ) { return; } // one elected lane drives the copies
    uint32_t it = 0u;
    for( uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it )
    {
      const uint32_t b0 = t * Cfg::T;
      if( b0 >= n ) { break; } // tiles past the binned bodies have nothing to stage (consumers only clear masks)
      const uint32_t sgi = it % SG_BP_TMA_STAGES, use = it / SG_BP_TMA_STAGES;
      unsigned char* stage = s_raw + sgi * STAGE;
      BpStage<D>* st = reinterpret_cast<BpStage<D>*>( stage + ST_OFF );
      // the tile's plan (two dependent rounds of global loads) does not need the stage: fetch it first, wait after
      const uint32_t b1 = ( n - b0 < uint32_t( Cfg::T ) ) ? n : b0 + Cfg::T;
      const long long kf = __ldg( &recs[b0].key );
      const long long kl = __ldg( &recs[b1 - 1u].key );
      long long klo[Cfg::NW], khi[Cfg::NW];
      uint32_t start[Cfg::NW], end[Cfg::NW];
      #pragma unroll
      for( int w = 0; w < Cfg::NW; ++w )
      {
        const long long off = ( long long )( w % 3 - 1 ) * g.dims[0];
        klo[w] = kf + off - 1; khi[w] = kl + off + 1;
        const bool ok = khi[w] >= 0 && klo[w] <= ( long long )( g.ncells ) - 1;
        klo[w] = ( klo[w] < 0 ) ? 0 : klo[w];
        khi[w] = ( khi[w] > ( long long )( g.ncells ) - 1 ) ? ( long long )( g.ncells ) - 1 : khi[w];
        if( !ok ) { klo[w] = 0; khi[w] = -1; }
        start[w] = ok ? __ldg( &cell_start[klo[w]] ) : 0u;
        end[w] = ok ? __ldg( &cell_start[khi[w] + 1] ) : 0u;
      }
      uint32_t bytes = 0u;
      uint32_t ncopies[Cfg::NW], cs_first[Cfg::NW], cs_n[Cfg::NW];
      uint32_t all_staged = 1u;
      if( use > 0u ) { sg_mbar_wait_backoff( &bars[2 + sgi], ( use - 1u ) & 1u ); } // consumers are done with this stage
      #pragma unroll
      for( int w = 0; w < Cfg::NW; ++w )
      {
        const uint32_t len = ( end[w] - start[w] < uint32_t( Cfg::WCAP ) ) ? end[w] - start[w] : uint32_t( Cfg::WCAP );
        ncopies[w] = ( len + SG_BP_TMA_ROWS - 1u ) / SG_BP_TMA_ROWS;
        // cell_start slice: entries klo .. khi+1, widened to whole 16-byte groups for the bulk copy
        const long long ncs = khi[w] - klo[w] + 2;
        cs_first[w] = uint32_t( klo[w] ) & ~3u;
        uint32_t want = ( khi[w] < klo[w] ) ? 0u : uint32_t( klo[w] - cs_first[w] + ncs );
        want = ( want + 3u ) & ~3u;
        cs_n[w] = ( want < uint32_t( SG_BP_TMA_CSCAP ) ) ? want : uint32_t( SG_BP_TMA_CSCAP );
        st->start[w] = start[w]; st->len[w] = len; st->cs_klo[w] = cs_first[w]; st->cs_len[w] = cs_n[w];
        if( end[w] - start[w] > uint32_t( Cfg::WCAP ) || want > uint32_t( SG_BP_TMA_CSCAP ) ) { all_staged = 0u; }
        bytes += ncopies[w] * uint32_t( SG_BP_TMA_ROWS * 64 ) + cs_n[w] * 4u;
      }
      st->full = all_staged;
      asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" ); // the stage's earlier generic reads vs the async writes to come
      sg_mbar_arrive_expect_tx( &bars[sgi], bytes );
      #pragma unroll
      for( int w = 0; w < Cfg::NW; ++w )
      {
        for( uint32_t c = 0u; c < ncopies[w]; ++c )
        {
          sg_tma_load_2d( stage + ( size_t( w ) * Cfg::WCAP + size_t( c ) * SG_BP_TMA_ROWS ) * 64, &tm_recs, 0, int( start[w] + c * SG_BP_TMA_ROWS ), &bars[sgi] );
        }
        if( cs_n[w] != 0u ) { sg_bulk_g2s( stage + CS_OFF + size_t( w ) * SG_BP_TMA_CSCAP * 4, cell_start + cs_first[w], cs_n[w] * 4u, &bars[sgi] ); }
      }
    }
    return;
  }

  // ---- consumers ----
  uint32_t it = 0u;
  for( uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it )
  {
    const uint32_t p = t * Cfg::T + threadIdx.x;
    if( t * Cfg::T >= n )
    {
      continue; // positions past the binned bodies belong to nobody
    }
    const uint32_t sgi = it % SG_BP_TMA_STAGES, use = it / SG_BP_TMA_STAGES;
    const unsigned char* stage = s_raw + sgi * STAGE;
    const unsigned char* s_recs = stage;
    const uint32_t* s_cs = reinterpret_cast<const uint32_t*>( stage + CS_OFF );
    const BpStage<D>* st = reinterpret_cast<const BpStage<D>*>( stage + ST_OFF );
    sg_mbar_wait( &bars[sgi], use & 1u );
    if( p < n )
    {
      const Rec me = sg_bp_fetch<P>( recs, s_recs, st, 1, p ); // own row is window 1 (dy = 0)
      const uint32_t my_idx = P::rec_idx( me );
      if( !P::owns( me ) )
      {
        counts[my_idx] = make_uint2( 0u, 0u );
        masks[size_t( p ) * BpPlan<D>::STRIDE] = make_uint4( 0u, 0u, 0u, 0u );
      }
      else
      {
        // tiles whose windows and cell_start slices were staged completely (the rule, not the exception) run a
        // walk with no fallback code in it at all
        if( st->full != 0u ) { sg_bp_count_body<P, SG_BP_TMA_CSCAP, true>( g, cell_start, recs, s_recs, s_cs, st, n_slots, p, me, my_idx, P::rec_ord( me ), counts, masks, plan ); }
        else { sg_bp_count_body<P, SG_BP_TMA_CSCAP, false>( g, cell_start, recs, s_recs, s_cs, st, n_slots, p, me, my_idx, P::rec_ord( me ), counts, masks, plan ); }
      }
    }
    // this warp is done with the stage
    __syncwarp();
    if( ( threadIdx.x & 31u ) == 0u ) { sg_mbar_arrive( &bars[2 + sgi] ); }
  }
}

// Tensor map over the sorted records: n rows of 16 x u32, box = 16 x SG_BP_TMA_ROWS, 64-byte swizzle.
// cuTensorMapEncodeTiled is fetched through the runtime (no link-time dependency on libcuda).
static inline int sg_bp_encode_recs_map( sg_ctx* ctx, void** encode_fn, const void* recs, const uint32_t n, CUtensorMap* out )
{
  typedef CUresult ( *EncodeFn )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
  if( *encode_fn == nullptr )
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres ) != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr )
    {
      cudaGetLastError();
      return sg_fail( ctx, SG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver" );
    }
    *encode_fn = fn;
  }
  const EncodeFn encode = reinterpret_cast<EncodeFn>( *encode_fn );
  const cuuint64_t gdim[2] = { 16, n };
  const cuuint64_t gstride[1] = { 64 };
  const cuuint32_t box[2] = { 16, SG_BP_TMA_ROWS };
  const cuuint32_t estr[2] = { 1, 1 };
  const CUresult r = encode( out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>( recs ), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
  if( r != CUDA_SUCCESS ) { return sg_fail( ctx, SG_ERR_CUDA, "cuTensorMapEncodeTiled( sorted records ) -> %d", int( r ) ); }
  return SG_OK;
}

// Dispatch: the 2-D pipelines take the box-prefiltered kernel (the record-staged one for dense scenes), the 3-D ones the block-staged kernel.
template<int D> struct SgBpCountLaunch;
template<> struct SgBpCountLaunch<3>
{
  template<typename P> static int run( sg_ctx* ctx, BroadScratch& s, const uint32_t n )
  {
    constexpr size_t smem = sg_bp_count_smem<3>();
    SG_CUDA( ctx, cudaFuncSetAttribute( sg_bp_count<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, int( smem ) ) );
    SG_LAUNCH( ctx, "bp_count", double( n ) * ( 64.0 + 8.0 + 16.0 + 16.0 * BpPlan<3>::NPLAN ), sg_bp_count<P><<<sg_div_up( n, BpCfg<3>::T ), BpCfg<3>::T, smem, ctx->stream>>>( n, s.params.as<GridParams>(), s.cell_start.as<uint32_t>(),
               s.recs.as<typename P::Rec>(), s.sidx.as<uint32_t>(), s.counts.as<uint2>(), s.masks.as<uint4>(), s.masks.as<uint4>() + 1 ) );
    return SG_OK;
  }
};
template<> struct SgBpCountLaunch<2>
{
  template<typename P> static int run( sg_ctx* ctx, BroadScratch& s, const uint32_t n )
  {
    if( s.dense )
    {
      if( s.recs.ptr != s.tm_ptr || n != s.tm_rows )
      {
        const int rc = sg_bp_encode_recs_map( ctx, &s.tm_encode, s.recs.ptr, n, &s.tm_recs );
        if( rc != SG_OK ) { return rc; }
        s.tm_ptr = s.recs.ptr; s.tm_rows = n;
      }
      constexpr size_t smem = sg_bp_tma_smem<2>();
      if( s.staged_attr_dev != ctx->device ) { SG_CUDA( ctx, cudaFuncSetAttribute( sg_bp_count_staged<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, int( smem ) ) ); s.staged_attr_dev = ctx->device; }
      const unsigned ntiles = sg_div_up( n, BpCfg<2>::T );
      const unsigned grid = ntiles < unsigned( ctx->num_sms ) * SG_BP_TMA_CTAS_PER_SM ? ntiles : unsigned( ctx->num_sms ) * SG_BP_TMA_CTAS_PER_SM;
      SG_LAUNCH( ctx, "bp_count", double( n ) * ( 64.0 + 8.0 + 16.0 + 16.0 * BpPlan<2>::NPLAN ), sg_bp_count_staged<P><<<grid, BpCfg<2>::T + 32, smem, ctx->stream>>>( s.tm_recs, n, s.params.as<GridParams>(), s.cell_start.as<uint32_t>(),
                 s.recs.as<typename P::Rec>(), s.counts.as<uint2>(), s.masks.as<uint4>(), s.masks.as<uint4>() + 1 ) );
      return SG_OK;
    }
    // 4 resident CTAs per SM (64 registers): measured against 5 (48 registers, 164 B of spills) and 6 (40, 284 B) on the B200 -- 185 / 231 / 285 us on 2 M
    // balls, 1.41 / 1.73 / 2.15 ms on 16 M (profiles/minb_ab_r2.jsonl) -- and against 3 (85 registers: 203 us).  The kernel is bound by issue slots and
    // load-to-use latency; spilled values cost more than the extra warps hide.
    SG_LAUNCH( ctx, "bp_count", double( n ) * ( 64.0 + 16.0 + 4.0 + 8.0 + 16.0 + 16.0 * BpPlan<2>::NPLAN ), sg_bp_count_l1<P, 2, 4><<<sg_div_up( n, SG_BP_THREADS ), SG_BP_THREADS, 0, ctx->stream>>>( n, s.params.as<GridParams>(), s.cell_start.as<uint32_t>(),
               s.recs.as<typename P::Rec>(), s.boxf.as<float4>(), s.sidx.as<uint32_t>(), s.counts.as<uint2>(), s.masks.as<uint4>(), s.masks.as<uint4>() + 1 ) );
    return SG_OK;
  }
};

// Slow paths of pass 2: redo the tests (a body with more than 32 neighbours or more candidates than the sorting
// network holds); everything comes through L1/L2.  Kept out of line so the common path stays lean in registers.
template<typename P>
__device__ __noinline__ void sg_bp_emit_slow( const uint32_t p, uint32_t nc, const uint2* __restrict__ counts, const ulonglong2 off, const uint32_t my_ord, const GridParams* __restrict__ params, const uint32_t* __restrict__ cell_start,
                                              const typename P::Rec* __restrict__ recs, const uint32_t* __restrict__ sidx, uint2* __restrict__ cand, const uint64_t cand_cap,
                                              uint2* __restrict__ work, const uint64_t work_cap )
{
  constexpr int D = P::D;
  using Cfg = BpCfg<D>;
  using Rec = typename P::Rec;
  const Rec me = sg_load_rec_global<Rec>( &recs[p] );
  if( nc == 0xffffffffu ) { nc = counts[P::rec_idx( me )].x; } // masks incomplete: the count comes from pass 1's by-index array
  if( nc == 0u ) { return; }
  unsigned long long ka = off.y;
  const GridParams g = *params;
  const uint32_t key = P::rec_key( me ), c1 = P::rec_c1( me, g ), c2 = P::rec_c2( me, g );
  BpStage<D> none; // nothing staged: sg_bp_cs falls through to cell_start
  #pragma unroll
  for( int w = 0; w < Cfg::NW; ++w ) { none.start[w] = 0u; none.len[w] = 0u; none.cs_klo[w] = 0u; none.cs_len[w] = 0u; }
  const BpStage<D>* st = &none;
  const uint32_t* s_cs = nullptr;
  auto ord_at = [&]( const int, const uint32_t q ) -> uint32_t { return __ldg( &sidx[q] ) & P::IDX_MASK; };
  double lo[D], hi[D];
  P::rec_aabb( me, lo, hi );
  auto overlaps = [&]( const Rec& o ) -> bool
  {
    double olo[D], ohi[D];
    P::rec_aabb( o, olo, ohi );
    bool ov = true;
    #pragma unroll
    for( int a = 0; a < D; ++a ) { ov = ov && !( hi[a] < olo[a] ) && !( ohi[a] < lo[a] ); }
    return ov;
  };
  auto emit_one = [&]( const unsigned long long kc, const Rec& o, const uint32_t o_ord, const uint32_t q )
  {
    if( cand != nullptr && kc < cand_cap ) { cand[kc] = make_uint2( my_ord, o_ord ); }
    if( P::HAS_NARROW ) { if( P::narrow_test( me, o ) ) { if( ka < work_cap ) { work[ka] = make_uint2( P::rec_idx( me ), q ); } ++ka; } }
  };
  if( nc <= SG_BP_LOCAL_CAP )
  {
    unsigned long long list[SG_BP_LOCAL_CAP];
    uint32_t nl = 0u;
    sg_bp_walk_pos<P>( g, cell_start, s_cs, st, p, key, c1, c2, [&]( const int w, const uint32_t q )
    {
      const uint32_t oi = ord_at( w, q );
      if( oi <= my_ord ) { return; }
      const Rec o = sg_load_rec_global<Rec>( &recs[q] );
      if( !overlaps( o ) ) { return; }
      const unsigned long long v = ( static_cast<unsigned long long>( oi ) << 32 ) | q;
      uint32_t j = nl++;
      while( j > 0u && list[j - 1u] > v ) { list[j] = list[j - 1u]; --j; }
      list[j] = v;
    } );
    for( uint32_t j = 0u; j < nl; ++j )
    {
      const uint32_t q = uint32_t( list[j] & 0xffffffffull );
      const Rec o = sg_load_rec_global<Rec>( &recs[q] );
      emit_one( off.x + j, o, uint32_t( list[j] >> 32 ), q );
    }
  }
  else
  {
    // Crowded body: partners in ascending order, SG_BP_BATCH at a time -- each walk keeps the SG_BP_BATCH smallest order words above the
    // last one emitted (bounded insertion into a sorted local list), so a body with nc candidates among m neighbours costs
    // ceil( nc / SG_BP_BATCH ) walks of m visits instead of nc walks (a mesh among thousands of small bodies: 2.6 ms -> see DESIGN 4.3)
    uint32_t last = my_ord;
    uint32_t done = 0u;
    while( done < nc )
    {
      unsigned long long best[SG_BP_BATCH]; // ( order word << 32 | position ), ascending; nb entries
      uint32_t nb = 0u;
      sg_bp_walk_pos<P>( g, cell_start, s_cs, st, p, key, c1, c2, [&]( const int w, const uint32_t q )
      {
        const uint32_t oi = ord_at( w, q );
        if( oi <= last ) { return; }
        const unsigned long long v = ( static_cast<unsigned long long>( oi ) << 32 ) | q;
        if( nb == uint32_t( SG_BP_BATCH ) && v >= best[SG_BP_BATCH - 1] ) { return; } // not among the smallest so far: skipped before touching the record
        const Rec o = sg_load_rec_global<Rec>( &recs[q] );
        if( !overlaps( o ) ) { return; }
        uint32_t j = ( nb < uint32_t( SG_BP_BATCH ) ) ? nb++ : uint32_t( SG_BP_BATCH - 1 );
        while( j > 0u && best[j - 1u] > v ) { best[j] = best[j - 1u]; --j; }
        best[j] = v;
      } );
      if( nb == 0u ) { break; } // (cannot happen: pass 1 counted nc overlapping partners)
      for( uint32_t j = 0u; j < nb && done < nc; ++j, ++done )
      {
        const uint32_t q = uint32_t( best[j] & 0xffffffffull );
        const Rec o = sg_load_rec_global<Rec>( &recs[q] );
        emit_one( off.x + done, o, uint32_t( best[j] >> 32 ), q );
      }
      last = uint32_t( best[nb - 1u] >> 32 );
    }
  }
}

// compare-exchange on packed (partner index << 32 | ...) keys
__device__ __forceinline__ void sg_cex( unsigned long long& a, unsigned long long& b )
{
  const unsigned long long lo = ( a < b ) ? a : b;
  const unsigned long long hi = ( a < b ) ? b : a;
  a = lo; b = hi;
}

template<int N> __device__ __forceinline__ void sg_sort_keys( unsigned long long* v, const uint32_t count );
template<> __device__ __forceinline__ void sg_sort_keys<8>( unsigned long long* v, const uint32_t )
{
  sg_cex( v[0], v[1] ); sg_cex( v[2], v[3] ); sg_cex( v[4], v[5] ); sg_cex( v[6], v[7] );
  sg_cex( v[0], v[2] ); sg_cex( v[1], v[3] ); sg_cex( v[4], v[6] ); sg_cex( v[5], v[7] );
  sg_cex( v[1], v[2] ); sg_cex( v[5], v[6] );
  sg_cex( v[0], v[4] ); sg_cex( v[1], v[5] ); sg_cex( v[2], v[6] ); sg_cex( v[3], v[7] );
  sg_cex( v[2], v[4] ); sg_cex( v[3], v[5] );
  sg_cex( v[1], v[2] ); sg_cex( v[3], v[4] ); sg_cex( v[5], v[6] );
}
template<> __device__ __forceinline__ void sg_sort_keys<16>( unsigned long long* v, const uint32_t count )
{
  if( count <= 8u ) { sg_sort_keys<8>( v, count ); return; } // keys fill v[0..count) in visit order, the rest is ~0
  sg_cex( v[0], v[1] ); sg_cex( v[2], v[3] ); sg_cex( v[0], v[2] ); sg_cex( v[1], v[3] ); sg_cex( v[1], v[2] ); sg_cex( v[4], v[5] ); sg_cex( v[6], v[7] ); sg_cex( v[4], v[6] );
  sg_cex( v[5], v[7] ); sg_cex( v[5], v[6] ); sg_cex( v[0], v[4] ); sg_cex( v[2], v[6] ); sg_cex( v[2], v[4] ); sg_cex( v[1], v[5] ); sg_cex( v[3], v[7] ); sg_cex( v[3], v[5] );
  sg_cex( v[1], v[2] ); sg_cex( v[3], v[4] ); sg_cex( v[5], v[6] ); sg_cex( v[8], v[9] ); sg_cex( v[10], v[11] ); sg_cex( v[8], v[10] ); sg_cex( v[9], v[11] ); sg_cex( v[9], v[10] );
  sg_cex( v[12], v[13] ); sg_cex( v[14], v[15] ); sg_cex( v[12], v[14] ); sg_cex( v[13], v[15] ); sg_cex( v[13], v[14] ); sg_cex( v[8], v[12] ); sg_cex( v[10], v[14] ); sg_cex( v[10], v[12] );
  sg_cex( v[9], v[13] ); sg_cex( v[11], v[15] ); sg_cex( v[11], v[13] ); sg_cex( v[9], v[10] ); sg_cex( v[11], v[12] ); sg_cex( v[13], v[14] ); sg_cex( v[0], v[8] ); sg_cex( v[4], v[12] );
  sg_cex( v[4], v[8] ); sg_cex( v[2], v[10] ); sg_cex( v[6], v[14] ); sg_cex( v[6], v[10] ); sg_cex( v[2], v[4] ); sg_cex( v[6], v[8] ); sg_cex( v[10], v[12] ); sg_cex( v[1], v[9] );
  sg_cex( v[5], v[13] ); sg_cex( v[5], v[9] ); sg_cex( v[3], v[11] ); sg_cex( v[7], v[15] ); sg_cex( v[7], v[11] ); sg_cex( v[3], v[5] ); sg_cex( v[7], v[9] ); sg_cex( v[11], v[13] );
  sg_cex( v[1], v[2] ); sg_cex( v[3], v[4] ); sg_cex( v[5], v[6] ); sg_cex( v[7], v[8] ); sg_cex( v[9], v[10] ); sg_cex( v[11], v[12] ); sg_cex( v[13], v[14] );
}

// Pass 2.  Each body writes its candidates (ascending partner order word) at its offset; active ones also leave a work item.
// One thread per body IN INDEX ORDER -- the order of the lists: consecutive threads own consecutive output ranges, so the
// stores are coalesced whatever the numbering of the scene has to do with space; what is random for a randomly numbered scene
// is one 32-byte read per body (masks + walk plan of its sorted position) and the gathers of its few partners' order words.
// Nothing is shared between threads: pass 1 left, per sorted position, the candidate/active masks over the visit
// sequence and the walk plan (window starts and lengths), so a thread decodes its set bits straight to sorted
// positions, gathers the partners' order words, orders its candidates (up to 8 in 2-D, 16 in 3-D) with a register sorting
// network and stores them.  No staging, no barriers, no shared memory.  The kernel is a chain of dependent gathers, i.e. latency-bound, and
// the 2-D instance (8-key network) fits 48 or even 40 registers without spilling.  Measured (config 3): when the gathers go to DRAM (16 M
// bodies) 48 warps per SM win -- 1007 / 875 / 816 us at 4 / 5 / 6 CTAs per SM; when they hit L2 (2 M) 40 warps do -- 125 / 122 / 135 us.
// The launch picks by size; the 3-D instance (16 keys) keeps 64 registers.
template<typename P, int MINB>
__global__ void __launch_bounds__( SG_BP_THREADS, MINB ) sg_bp_emit( const uint32_t n_slots, const GridParams* __restrict__ params, const uint32_t* __restrict__ cell_start,
                                                              const typename P::Rec* __restrict__ recs, const uint32_t* __restrict__ sidx, const uint32_t* __restrict__ pos_of, const uint32_t* __restrict__ ord_by_index,
                                                              const uint4* __restrict__ masks, const uint4* __restrict__ plan, const uint2* __restrict__ counts,
                                                              const ulonglong2* __restrict__ offsets, uint2* __restrict__ cand, const uint64_t cand_cap, uint2* __restrict__ work, const uint64_t work_cap )
{
  constexpr int D = P::D;
  using Cfg = BpCfg<D>;
  using Rec = typename P::Rec;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n_slots ) { return; }
  const uint32_t p = __ldg( &pos_of[i] );
  const ulonglong2 off = offsets[i];
  const uint32_t my_ord = ( ord_by_index != nullptr ) ? __ldg( &ord_by_index[i] ) : i; // the order word is what the lists carry (== the body index on one GPU)
  if( p == 0xffffffffu ) { return; } // an unused slot (multi-GPU ghost capacity)
  const uint4 m = __ldg( &masks[size_t( p ) * BpPlan<D>::STRIDE] ); // zero for ghosts
  uint32_t qb[Cfg::NW], len[Cfg::NW];
  sg_bp_plan_load<D>( plan, n_slots, p, qb, len );
  const unsigned long long cmask = m.x | ( static_cast<unsigned long long>( m.y ) << 32 );
  const unsigned long long amask = m.z | ( static_cast<unsigned long long>( m.w ) << 32 );
  const bool complete = ( amask & SG_BP_MASKS_INVALID ) == 0ull;
  if( complete && cmask == 0ull ) { return; }
  const uint32_t nc = uint32_t( __popcll( cmask ) );
  unsigned long long ka = off.y;

  if( complete && nc <= uint32_t( Cfg::FAST ) )
  {
    unsigned long long v[Cfg::FAST];
    if( Cfg::NW > 3 )
    {
      // Many windows (3-D: 9): walk the windows once and peel each window's bits off the mask -- O(windows + bits)
      // instead of O(windows x bits).  The decoded (position, active) words go through a small local array (dynamic
      // index); the index gathers are then issued from an unrolled loop so that they overlap.
      uint32_t dec[Cfg::FAST];
      uint32_t cnt = 0u, base = 0u;
      #pragma unroll
      for( int w = 0; w < Cfg::NW; ++w )
      {
        const bool mine = p - qb[w] < len[w]; // p inside this window (unsigned compare)
        const uint32_t L = len[w] - ( mine ? 1u : 0u );
        unsigned long long sub = ( base < 64u ) ? ( cmask >> base ) : 0ull;
        if( L < 64u ) { sub &= ( 1ull << L ) - 1ull; }
        while( sub != 0ull )
        {
          const uint32_t kk = uint32_t( __ffsll( static_cast<long long>( sub ) ) ) - 1u;
          sub &= sub - 1ull;
          uint32_t q = qb[w] + kk;
          if( mine && q >= p ) { ++q; }
          if( cnt < uint32_t( Cfg::FAST ) ) { dec[cnt] = q | ( uint32_t( ( amask >> ( base + kk ) ) & 1ull ) << 31 ); }
          ++cnt;
        }
        base += L;
      }
      #pragma unroll
      for( int b = 0; b < Cfg::FAST; ++b )
      {
        v[b] = ~0ull;
        if( uint32_t( b ) < nc )
        {
          const uint32_t w32 = dec[b];
          const uint32_t oi = __ldg( &sidx[w32 & 0x7fffffffu] ) & P::IDX_MASK;
          v[b] = ( static_cast<unsigned long long>( oi ) << 32 ) | w32;
        }
      }
    }
    else
    {
    // visits before window w (the body itself is skipped inside its own window)
    unsigned long long cm = cmask;
    #pragma unroll
    for( int b = 0; b < Cfg::FAST; ++b )
    {
      v[b] = ~0ull;
      if( cm != 0ull )
      {
        const uint32_t kk = uint32_t( __ffsll( static_cast<long long>( cm ) ) ) - 1u;
        cm &= cm - 1ull;
        uint32_t rem = kk, q = 0u;
        bool found = false;
        #pragma unroll
        for( int w = 0; w < Cfg::NW; ++w )
        {
          const bool mine = p - qb[w] < len[w]; // p inside this window (unsigned compare)
          const uint32_t L = len[w] - ( mine ? 1u : 0u );
          if( !found )
          {
            if( rem < L ) { q = qb[w] + rem; if( mine && q >= p ) { ++q; } found = true; }
            else { rem -= L; }
          }
        }
        const uint32_t oi = __ldg( &sidx[q] ) & P::IDX_MASK;
        v[b] = ( static_cast<unsigned long long>( oi ) << 32 ) | ( ( ( amask >> kk ) & 1ull ) << 31 ) | q;
      }
    }
    }
    // Batcher odd-even merge sort networks (empty slots hold ~0 and sink to the end): 8 keys / 19 exchanges,
    // 16 keys / 63 exchanges (3-D pipelines, where a lattice body owns 13 of its 26 neighbours)
    if( nc > 1u ) { sg_sort_keys<Cfg::FAST>( v, nc ); }
    if( cand != nullptr )
    {
      #pragma unroll
      for( int j = 0; j < Cfg::FAST; ++j )
      {
        const unsigned long long kc = off.x + j;
        if( uint32_t( j ) < nc && kc < cand_cap ) { cand[kc] = make_uint2( my_ord, uint32_t( v[j] >> 32 ) ); }
      }
    }
    if( P::HAS_NARROW && amask != 0ull )
    {
      // active pairs: only (own position, partner position) is recorded here, in output order; the contact
      // geometry is computed by sg_bp_contacts, one thread per contact, with fully coalesced stores
      #pragma unroll
      for( int j = 0; j < Cfg::FAST; ++j )
      {
        if( uint32_t( j ) < nc && ( ( v[j] >> 31 ) & 1ull ) )
        {
          if( ka < work_cap ) { work[ka] = make_uint2( i, uint32_t( v[j] & 0x7fffffffull ) ); }
          ++ka;
        }
      }
    }
    return;
  }

  sg_bp_emit_slow<P>( p, complete ? nc : 0xffffffffu, counts, off, my_ord, params, cell_start, recs, sidx, cand, cand_cap, work, work_cap );
}

// Pass 3 (policies with a fused narrow phase).  One thread per active pair, grid-stride over the work list pass 2 left in output
// order: ( body index of the pair's first body, sorted position of its partner ).  The list is in index order, so the first body's
// data is read from the caller's by-index arrays -- coalesced, consecutive contacts share it -- and only the partner's record is a
// gather (one 64-byte line instead of two: on 16 M randomly numbered balls, where every gather is a DRAM access, that is most of
// this kernel's traffic); the contact is written at its own index => every store of the SoA contact arrays is fully coalesced.
template<typename P>
__global__ void __launch_bounds__( SG_BP_THREADS, 5 ) sg_bp_contacts( const typename P::In in, const ScanPairCounts::Acc* __restrict__ totals, const uint2* __restrict__ work, const uint64_t work_cap,
                                                                  const typename P::Rec* __restrict__ recs, const typename P::Out out )
{
  using Rec = typename P::Rec;
  unsigned long long na = totals->a;
  if( na > work_cap ) { na = work_cap; }
  for( unsigned long long c = blockIdx.x * uint64_t( blockDim.x ) + threadIdx.x; c < na; c += uint64_t( gridDim.x ) * blockDim.x )
  {
    const uint2 w = __ldg( &work[c] );
    const Rec a = P::make_rec( in, w.x, 0u, 0u, 0u ); // what the scatter put into the body's record (cell words aside)
    const Rec b = sg_load_rec_global<Rec>( &recs[w.y] );
    unsigned long long k = c;
    P::contact_emit( out, k, a, b );
  }
}

// ---- host driver -----------------------------------------------------------------------------------
template<typename P>
static int sg_bp_prepare_scratch( sg_ctx* ctx, BroadScratch& s, const uint32_t n )
{
  // cell count capped at ~2 cells per body (power of two not needed: keys are row-major)
  uint64_t mc = uint64_t( n ) * 2u + 1024u;
  if( mc > 0x7fffffffull ) { mc = 0x7fffffffull; }
  s.max_cells = uint32_t( mc );
  if( s.bounds.ptr == nullptr )
  {
    SG_CUDA( ctx, s.bounds.ensure( 2 * sizeof( BoundsAccum ) ) );
    SG_LAUNCH( ctx, "bp_bounds_init", 0.0, sg_bp_bounds_init<<<1, 32, 0, ctx->stream>>>( s.bounds.as<BoundsAccum>() ) );
    s.bounds_phase = 0;
  }
  SG_CUDA( ctx, s.params.ensure( sizeof( GridParams ) ) );
  SG_CUDA( ctx, s.cell_count.ensure( ( size_t( s.max_cells ) + 2 ) * 4 ) );
  if( s.cell_count.ptr != s.hist_ptr || s.max_cells + 2u > s.hist_slots ) { s.hist_clean = false; s.hist_ptr = s.cell_count.ptr; s.hist_slots = s.max_cells + 2u; }
  SG_CUDA( ctx, s.cell_start.ensure( ( size_t( s.max_cells ) + 8 ) * 4 ) ); // + slack: bulk copies read whole 16-byte groups
  SG_CUDA( ctx, s.cell_partials.ensure( ( size_t( s.max_cells ) / SG_SCAN_TILE + 2 ) * 4 ) );
  SG_CUDA( ctx, s.key.ensure( size_t( n ) * 4 ) );
  SG_CUDA( ctx, s.rank.ensure( size_t( n ) * 4 ) );
  SG_CUDA( ctx, s.recs.ensure( size_t( n ) * 64 ) );
  SG_CUDA( ctx, s.sidx.ensure( size_t( n ) * 4 + 512 ) ); // + slack: bulk copies read whole 16-byte groups, pass 1 up to a window's length past the end
  if( P::D == 2 ) { SG_CUDA( ctx, s.boxf.ensure( size_t( n ) * 16 + 2048 ) ); } // + slack: pass 1 reads up to a window's length (<= 64 entries) past the end
  SG_CUDA( ctx, s.pos_of.ensure( size_t( n ) * 4 ) );
  SG_CUDA( ctx, s.counts.ensure( size_t( n ) * sizeof( uint2 ) ) );
  SG_CUDA( ctx, s.masks.ensure( size_t( n ) * sizeof( uint4 ) * BpPlan<P::D>::STRIDE ) ); // masks and walk plan interleaved by position
  SG_CUDA( ctx, s.offsets.ensure( size_t( n ) * sizeof( ulonglong2 ) ) );
  SG_CUDA( ctx, s.pair_partials.ensure( ( size_t( n ) / SG_SCAN_TILE + 2 ) * sizeof( ScanPairCounts::Acc ) ) );
  SG_CUDA( ctx, s.totals.ensure( sizeof( ScanPairCounts::Acc ) ) );
  return SG_OK;
}

// Sort bodies by cell and count.  After this returns (asynchronously) s.totals holds {P_c, P_a}.
template<typename P>
static int sg_bp_bin_and_count( sg_ctx* ctx, BroadScratch& s, const typename P::In& in, const bool bounds_done = false )
{
  constexpr int D = P::D;
  const uint32_t n = in.n;
  const unsigned nblk = sg_div_up( n, SG_BP_THREADS );
  const unsigned nred = nblk < unsigned( ctx->num_sms * 8 ) ? nblk : unsigned( ctx->num_sms * 8 );
  const double nb = double( n );
  if( !bounds_done ) { SG_LAUNCH( ctx, "bp_bounds", nb * P::IN_BYTES, sg_bp_bounds<P><<<nred, SG_BP_THREADS, 0, ctx->stream>>>( in, s.bounds_cur() ) ); }
  BoundsAccum* acc_cur = s.bounds_cur();
  s.bounds_phase ^= 1;
  if( !s.hist_clean )
  {
    SG_CUDA( ctx, cudaMemsetAsync( s.cell_count.ptr, 0, ( size_t( s.max_cells ) + 2 ) * 4, ctx->stream ) );
  }
  s.hist_clean = false;
  SG_LAUNCH( ctx, "bp_hist", nb * ( P::IN_BYTES + 8.0 ), sg_bp_hist<P><<<nblk, SG_BP_THREADS, 0, ctx->stream>>>( in, acc_cur, s.bounds_cur(), s.max_cells, s.params.as<GridParams>(), s.cell_count.as<uint32_t>(), s.key.as<uint32_t>(),
             s.rank.as<uint32_t>(), s.counts.as<uint2>() ) );
  const uint32_t* ncells_dev = &s.params.as<GridParams>()->ncells;
  int rc = sg_exclusive_scan<ScanU32>( ctx, "bp_cell_scan", s.cell_count.as<uint32_t>(), ncells_dev, 0u, s.max_cells, s.cell_partials.as<uint32_t>(), s.cell_start.as<uint32_t>(), nullptr, true );
  if( rc != SG_OK ) { return rc; }
  SG_LAUNCH( ctx, "bp_scatter", nb * ( P::IN_BYTES + 8.0 + 4.0 + 64.0 + 4.0 ) + double( s.max_cells ) * 4.0, sg_bp_scatter<P><<<nblk, SG_BP_THREADS, 0, ctx->stream>>>( in, s.params.as<GridParams>(), s.cell_start.as<uint32_t>(), s.key.as<uint32_t>(), s.rank.as<uint32_t>(), s.recs.as<typename P::Rec>(), s.pos_of.as<uint32_t>(), s.cell_count.as<uint32_t>(), s.max_cells + 2u ) );
  SG_LAUNCH( ctx, "bp_side", nb * ( 64.0 + 4.0 + ( D == 2 ? 16.0 : 0.0 ) ), sg_bp_side_arrays<P><<<nblk, SG_BP_THREADS, 0, ctx->stream>>>( n, s.params.as<GridParams>(), s.cell_start.as<uint32_t>(), s.recs.as<typename P::Rec>(),
             s.sidx.as<uint32_t>(), D == 2 ? s.boxf.as<float4>() : nullptr, s.ord_by_index ) );
  s.hist_clean = true;
  rc = SgBpCountLaunch<D>::template run<P>( ctx, s, n );
  if( rc != SG_OK ) { return rc; }
  rc = sg_exclusive_scan<ScanPairCounts>( ctx, "bp_pair_scan", s.counts.as<uint2>(), nullptr, n, n, s.pair_partials.as<ScanPairCounts::Acc>(), s.offsets.as<ulonglong2>(), s.totals.as<ScanPairCounts::Acc>(), false, nullptr, s.side.n != 0u ? &s.side : nullptr );
  s.side.n = 0u;
  return rc;
}

template<bool> struct SgBpContactsLaunch;
template<> struct SgBpContactsLaunch<false>
{
  template<typename P> static int run( sg_ctx*, BroadScratch&, const typename P::In&, const typename P::Out&, const uint64_t ) { return SG_OK; }
};
template<> struct SgBpContactsLaunch<true>
{
  template<typename P> static int run( sg_ctx* ctx, BroadScratch& s, const typename P::In& in, const typename P::Out& out, const uint64_t act_cap )
  {
    // (6 CTAs per SM -- 40 registers, 40 bytes spilled -- measured slower at every size: 833 vs 788 us on 16 M balls)
    SG_LAUNCH( ctx, "bp_contacts", 0.0, sg_bp_contacts<P><<<unsigned( ctx->num_sms ) * 5u, SG_BP_THREADS, 0, ctx->stream>>>( in, s.totals.as<ScanPairCounts::Acc>(), s.work.as<uint2>(), act_cap < s.work_cap ? act_cap : s.work_cap,
               s.recs.as<typename P::Rec>(), out ) );
    return SG_OK;
  }
};

// act_cap = capacity of the caller's contact arrays (0 for policies without a narrow phase)
template<typename P>
static int sg_bp_emit_lists( sg_ctx* ctx, BroadScratch& s, const typename P::In& in, const uint32_t n, const bool want_cand, const typename P::Out& out, const uint64_t act_cap )
{
  if( P::HAS_NARROW && act_cap > s.work_cap )
  {
    SG_CUDA( ctx, s.work.ensure( act_cap * sizeof( uint2 ) ) );
    s.work_cap = act_cap;
  }
  const double emit_bytes = double( n ) * ( 4.0 + 16.0 + 16.0 + 16.0 * BpPlan<P::D>::NPLAN );
  #define SG_EMIT_ARGS n, s.params.as<GridParams>(), s.cell_start.as<uint32_t>(), s.recs.as<typename P::Rec>(), s.sidx.as<uint32_t>(), s.pos_of.as<uint32_t>(), s.ord_by_index, s.masks.as<uint4>(), s.masks.as<uint4>() + 1, \
                       s.counts.as<uint2>(), s.offsets.as<ulonglong2>(), want_cand ? s.cand.as<uint2>() : nullptr, s.cand_cap, s.work.as<uint2>(), act_cap < s.work_cap ? act_cap : s.work_cap
  if( P::D == 3 ) { SG_LAUNCH( ctx, "bp_emit", emit_bytes, sg_bp_emit<P, 4><<<sg_div_up( n, SG_BP_THREADS ), SG_BP_THREADS, 0, ctx->stream>>>( SG_EMIT_ARGS ) ); }
  else if( size_t( n ) * 64 > size_t( 192 ) << 20 ) { SG_LAUNCH( ctx, "bp_emit", emit_bytes, sg_bp_emit<P, ( P::D == 2 ) ? 6 : 4><<<sg_div_up( n, SG_BP_THREADS ), SG_BP_THREADS, 0, ctx->stream>>>( SG_EMIT_ARGS ) ); } // records well beyond L2
  else { SG_LAUNCH( ctx, "bp_emit", emit_bytes, sg_bp_emit<P, ( P::D == 2 ) ? 5 : 4><<<sg_div_up( n, SG_BP_THREADS ), SG_BP_THREADS, 0, ctx->stream>>>( SG_EMIT_ARGS ) ); }
  #undef SG_EMIT_ARGS
  return SgBpContactsLaunch<P::HAS_NARROW>::template run<P>( ctx, s, in, out, act_cap );
}

#endif
