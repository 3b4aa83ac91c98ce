// sg_aabb.cu -- broad phase on caller-supplied boxes: the drop-in for
//   SpatialGridDetector::getPotentialOverlaps( const std::vector<AABB>&, std::set<std::pair<unsigned,unsigned>>& )
//   (ball2d/SpatialGridDetector.h:39, rigidbody2d/SpatialGrid.h:35, rigidbody3d/SpatialGridDetector.h:37).
// Output: every (i<j) whose boxes overlap (closed intervals), ascending -- the std::set iteration order.
#include "sg_boxes.cuh"

struct AabbData
{
  DevBuf boxes;
  BroadScratch bp;
  PinBuf h_totals;
  PinBuf h_pairs;
};

void sg_aabb_release( sg_ctx* ctx )
{
  AabbData* d = ctx->aabb;
  if( d == nullptr ) { return; }
  d->boxes.release(); d->bp.release(); d->h_totals.release(); d->h_pairs.release();
  delete d;
  ctx->aabb = nullptr;
}

template<int DIM>
static int candidate_pairs_impl( sg_ctx* ctx, AabbData* d, const uint32_t n, const double* aabbs, sg_pairs* out )
{
  using P = AabbPolicy<DIM>;
  const size_t in_bytes = size_t( n ) * 2 * DIM * 8;
  SG_CUDA( ctx, d->boxes.ensure( in_bytes ) );
  SG_CUDA( ctx, d->h_totals.ensure( 64 ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->boxes.ptr, aabbs, in_bytes, cudaMemcpyHostToDevice, ctx->stream ) );
  int rc = sg_bp_prepare_scratch<P>( ctx, d->bp, n );
  if( rc != SG_OK ) { return rc; }
  typename P::In in;
  in.boxes = d->boxes.template as<double>(); in.n = n;
  rc = sg_bp_bin_and_count<P>( ctx, d->bp, in );
  if( rc != SG_OK ) { return rc; }
  // the list size is needed before the emit can be given a buffer: one small synchronous read-back
  unsigned long long* ht = d->h_totals.template as<unsigned long long>();
  SG_CUDA( ctx, cudaMemcpyAsync( ht, d->bp.totals.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  const uint64_t np = ht[0];
  if( np > d->bp.cand_cap )
  {
    SG_CUDA( ctx, d->bp.cand.ensure( size_t( np ) * sizeof( uint2 ) ) );
    d->bp.cand_cap = d->bp.cand.cap / sizeof( uint2 );
  }
  SG_CUDA( ctx, d->h_pairs.ensure( size_t( np ) * 8 + 64 ) );
  if( np > 0 )
  {
    rc = sg_bp_emit_lists<P>( ctx, d->bp, in, n, true, NoOut{}, 0u );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, cudaMemcpyAsync( d->h_pairs.ptr, d->bp.cand.ptr, size_t( np ) * 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  }
  sg_prof_collect( ctx );
  out->n = np;
  out->ij = d->h_pairs.template as<uint32_t>();
  return SG_OK;
}

extern "C" int sg_candidate_pairs( sg_ctx* ctx, int dim, uint32_t n, const double* aabbs, sg_pairs* out )
{
  if( ctx == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  if( dim != 2 && dim != 3 ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_candidate_pairs: dim must be 2 or 3, got %d", dim ); }
  out->n = 0; out->ij = nullptr;
  if( n == 0 ) { return SG_OK; }
  if( aabbs == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_candidate_pairs: null box array" ); }
  if( n >= 0x80000000u ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_candidate_pairs: at most 2^31 - 1 boxes" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( ctx->aabb == nullptr ) { ctx->aabb = new AabbData; }
  if( dim == 2 ) { return candidate_pairs_impl<2>( ctx, ctx->aabb, n, aabbs, out ); }
  return candidate_pairs_impl<3>( ctx, ctx->aabb, n, aabbs, out );
}
