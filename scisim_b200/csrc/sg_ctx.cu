// sg_ctx.cu -- context lifetime, error text, pinned host memory, per-kernel event timing.
#include "sg_common.cuh"

#include <cstdarg>

static std::string g_create_error;

void sg_ball2d_release( sg_ctx* ctx );
void sg_aabb_release( sg_ctx* ctx );
void sg_rb3d_release( sg_ctx* ctx );
void sg_rb2d_release( sg_ctx* ctx );

int sg_fail( sg_ctx* ctx, int code, const char* fmt, ... )
{
  char buf[1024];
  va_list ap;
  va_start( ap, fmt );
  vsnprintf( buf, sizeof( buf ), fmt, ap );
  va_end( ap );
  if( ctx != nullptr ) { ctx->err = buf; } else { g_create_error = buf; }
  // leave the CUDA error state clean so the next call reports its own failure
  cudaGetLastError();
  return code;
}

int sg_prof_entry( sg_ctx* ctx, const char* name )
{
  for( size_t k = 0; k < ctx->prof.size(); ++k ) { if( strcmp( ctx->prof[k].name, name ) == 0 ) { return int( k ); } }
  ProfEntry e; e.name = name;
  ctx->prof.push_back( e );
  return int( ctx->prof.size() ) - 1;
}

static cudaEvent_t sg_get_event( sg_ctx* ctx )
{
  if( !ctx->event_pool.empty() ) { cudaEvent_t e = ctx->event_pool.back(); ctx->event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate( &e );
  return e;
}

void sg_prof_begin( sg_ctx* ctx, const char* name, double bytes )
{
  ProfPending p;
  p.entry = sg_prof_entry( ctx, name );
  ctx->prof[p.entry].bytes += bytes;
  ctx->prof[p.entry].launches += 1;
  p.e0 = sg_get_event( ctx );
  p.e1 = sg_get_event( ctx );
  cudaEventRecord( p.e0, ctx->stream );
  ctx->pending.push_back( p );
}

void sg_prof_end( sg_ctx* ctx )
{
  cudaEventRecord( ctx->pending.back().e1, ctx->stream );
}

// call after the stream has been synchronised
void sg_prof_collect( sg_ctx* ctx )
{
  for( ProfPending& p : ctx->pending )
  {
    float ms = 0.0f;
    if( cudaEventElapsedTime( &ms, p.e0, p.e1 ) == cudaSuccess ) { ctx->prof[p.entry].ms += double( ms ); }
    ctx->event_pool.push_back( p.e0 );
    ctx->event_pool.push_back( p.e1 );
  }
  ctx->pending.clear();
}

extern "C"
{

int sg_create( sg_ctx** out, int device )
{
  if( out == nullptr ) { return sg_fail( nullptr, SG_ERR_INVALID, "sg_create: null output pointer" ); }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount( &count );
  if( e != cudaSuccess || count == 0 )
  {
    return sg_fail( nullptr, SG_ERR_CUDA, "sg_create: no usable CUDA device (%s); this library has no CPU fallback", cudaGetErrorString( e ) );
  }
  if( device < 0 || device >= count ) { return sg_fail( nullptr, SG_ERR_INVALID, "sg_create: device %d out of range (0..%d)", device, count - 1 ); }
  e = cudaSetDevice( device );
  if( e != cudaSuccess ) { return sg_fail( nullptr, SG_ERR_CUDA, "sg_create: cudaSetDevice(%d): %s", device, cudaGetErrorString( e ) ); }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties( &prop, device );
  if( e != cudaSuccess ) { return sg_fail( nullptr, SG_ERR_CUDA, "sg_create: cudaGetDeviceProperties: %s", cudaGetErrorString( e ) ); }
  if( prop.major != 10 )
  {
    return sg_fail( nullptr, SG_ERR_CUDA, "sg_create: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor );
  }
  sg_ctx* ctx = new sg_ctx;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags( &ctx->stream, cudaStreamNonBlocking );
  if( e != cudaSuccess ) { delete ctx; return sg_fail( nullptr, SG_ERR_CUDA, "sg_create: cudaStreamCreate: %s", cudaGetErrorString( e ) ); }
  e = cudaStreamCreateWithFlags( &ctx->stream2, cudaStreamNonBlocking );
  if( e == cudaSuccess ) { e = cudaEventCreateWithFlags( &ctx->ev_fork, cudaEventDisableTiming ); }
  if( e == cudaSuccess ) { e = cudaEventCreateWithFlags( &ctx->ev_join, cudaEventDisableTiming ); }
  if( e != cudaSuccess ) { cudaStreamDestroy( ctx->stream ); delete ctx; return sg_fail( nullptr, SG_ERR_CUDA, "sg_create: side stream: %s", cudaGetErrorString( e ) ); }
  *out = ctx;
  return SG_OK;
}

void sg_destroy( sg_ctx* ctx )
{
  if( ctx == nullptr ) { return; }
  cudaSetDevice( ctx->device );
  cudaStreamSynchronize( ctx->stream );
  sg_prof_collect( ctx );
  sg_ball2d_release( ctx );
  sg_aabb_release( ctx );
  sg_rb3d_release( ctx );
  sg_rb2d_release( ctx );
  for( cudaEvent_t e : ctx->event_pool ) { cudaEventDestroy( e ); }
  if( ctx->timer0 != nullptr ) { cudaEventDestroy( ctx->timer0 ); cudaEventDestroy( ctx->timer1 ); }
  ctx->l2_flush.release();
  ctx->scan_vals.release();
  cudaStreamDestroy( ctx->stream );
  if( ctx->stream2 != nullptr ) { cudaStreamDestroy( ctx->stream2 ); }
  if( ctx->ev_fork != nullptr ) { cudaEventDestroy( ctx->ev_fork ); }
  if( ctx->ev_join != nullptr ) { cudaEventDestroy( ctx->ev_join ); }
  delete ctx;
}

const char* sg_last_error( const sg_ctx* ctx )
{
  return ( ctx != nullptr ) ? ctx->err.c_str() : g_create_error.c_str();
}

int sg_synchronize( sg_ctx* ctx )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

int sg_host_alloc( sg_ctx* ctx, uint64_t bytes, void** ptr )
{
  if( ctx == nullptr || ptr == nullptr ) { return SG_ERR_INVALID; }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SG_CUDA( ctx, cudaMallocHost( ptr, bytes > 0 ? bytes : 1 ) );
  return SG_OK;
}

int sg_host_free( sg_ctx* ctx, void* ptr )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  SG_CUDA( ctx, cudaFreeHost( ptr ) );
  return SG_OK;
}

void* sg_stream( sg_ctx* ctx ) { return ( ctx != nullptr ) ? static_cast<void*>( ctx->stream ) : nullptr; }

int sg_profile_enable( sg_ctx* ctx, int on )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  const int rc = sg_synchronize( ctx );
  ctx->profile = ( on != 0 );
  return rc;
}

int sg_profile_reset( sg_ctx* ctx )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  const int rc = sg_synchronize( ctx );
  for( ProfEntry& e : ctx->prof ) { e.launches = 0; e.ms = 0.0; e.bytes = 0.0; }
  return rc;
}

int sg_profile_count( sg_ctx* ctx ) { return ( ctx != nullptr ) ? int( ctx->prof.size() ) : 0; }

int sg_profile_get( sg_ctx* ctx, int k, const char** name, uint64_t* launches, double* ms, double* bytes )
{
  if( ctx == nullptr || k < 0 || k >= int( ctx->prof.size() ) ) { return SG_ERR_INVALID; }
  if( name != nullptr ) { *name = ctx->prof[k].name; }
  if( launches != nullptr ) { *launches = ctx->prof[k].launches; }
  if( ms != nullptr ) { *ms = ctx->prof[k].ms; }
  if( bytes != nullptr ) { *bytes = ctx->prof[k].bytes; }
  return SG_OK;
}

uint64_t sg_launch_count( const sg_ctx* ctx ) { return ( ctx != nullptr ) ? ctx->launch_count : 0; }

int sg_timer_begin( sg_ctx* ctx )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( ctx->timer0 == nullptr ) { SG_CUDA( ctx, cudaEventCreate( &ctx->timer0 ) ); SG_CUDA( ctx, cudaEventCreate( &ctx->timer1 ) ); }
  SG_CUDA( ctx, cudaEventRecord( ctx->timer0, ctx->stream ) );
  return SG_OK;
}

int sg_timer_end( sg_ctx* ctx, double* ms )
{
  if( ctx == nullptr || ms == nullptr || ctx->timer0 == nullptr ) { return SG_ERR_INVALID; }
  SG_CUDA( ctx, cudaEventRecord( ctx->timer1, ctx->stream ) );
  SG_CUDA( ctx, cudaEventSynchronize( ctx->timer1 ) );
  float f = 0.0f;
  SG_CUDA( ctx, cudaEventElapsedTime( &f, ctx->timer0, ctx->timer1 ) );
  *ms = double( f );
  return SG_OK;
}

int sg_flush_l2( sg_ctx* ctx )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const size_t bytes = size_t( 384 ) << 20; // 3x the 126 MB L2
  SG_CUDA( ctx, ctx->l2_flush.ensure( bytes ) );
  SG_CUDA( ctx, cudaMemsetAsync( ctx->l2_flush.ptr, int( ctx->launch_count & 0xff ), bytes, ctx->stream ) );
  return SG_OK;
}

}
