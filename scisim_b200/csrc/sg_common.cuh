// sg_common.cuh -- shared host/device plumbing for libscisim_b200 (sm_100a only).
//
// All arithmetic that must match the reference bit for bit is FP64 and this library is compiled with
// -fmad=false (the reference builds in ISO C++ mode => no FP contraction, CMakeLists.txt:53-56), IEEE
// division and square root (nvcc defaults for double), no fast-math.
#ifndef SG_COMMON_CUH
#define SG_COMMON_CUH

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scisim_b200.h"

#define SG_MAX_PLANES 32
#define SG_MAX_DRUMS 32

// ---------------------------------------------------------------------------------------------------
// Host side: context, growable device/pinned buffers, launch bookkeeping
// ---------------------------------------------------------------------------------------------------

struct DevBuf
{
  void* ptr = nullptr;
  size_t cap = 0;
  // grows (never shrinks); contents are NOT preserved across a growth
  cudaError_t ensure( size_t bytes )
  {
    if( bytes <= cap ) { return cudaSuccess; }
    if( ptr != nullptr ) { cudaFree( ptr ); ptr = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4;
    want = ( want + 255 ) & ~size_t( 255 );
    cudaError_t e = cudaMalloc( &ptr, want );
    if( e != cudaSuccess ) { want = ( bytes + 255 ) & ~size_t( 255 ); e = cudaMalloc( &ptr, want ); }
    if( e == cudaSuccess ) { cap = want; }
    return e;
  }
  void release() { if( ptr != nullptr ) { cudaFree( ptr ); } ptr = nullptr; cap = 0; }
  template<typename T> T* as() const { return static_cast<T*>( ptr ); }
};

struct PinBuf
{
  void* ptr = nullptr;
  size_t cap = 0;
  cudaError_t ensure( size_t bytes )
  {
    if( bytes <= cap ) { return cudaSuccess; }
    if( ptr != nullptr ) { cudaFreeHost( ptr ); ptr = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4;
    want = ( want + 4095 ) & ~size_t( 4095 );
    const cudaError_t e = cudaMallocHost( &ptr, want );
    if( e == cudaSuccess ) { cap = want; }
    return e;
  }
  void release() { if( ptr != nullptr ) { cudaFreeHost( ptr ); } ptr = nullptr; cap = 0; }
  template<typename T> T* as() const { return static_cast<T*>( ptr ); }
};

struct ProfEntry
{
  const char* name;
  uint64_t launches = 0;
  double ms = 0.0;
  double bytes = 0.0;
};

struct ProfPending
{
  int entry;
  cudaEvent_t e0;
  cudaEvent_t e1;
};

struct Ball2DData;
struct AabbData;
struct Rb3dData;
struct Rb2dData;

struct sg_ctx
{
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;              // side stream: independent kernels of one step run beside the main chain
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  uint64_t launch_count = 0;

  bool profile = false;
  std::vector<ProfEntry> prof;
  std::vector<ProfPending> pending;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t timer0 = nullptr;
  cudaEvent_t timer1 = nullptr;
  DevBuf l2_flush;
  DevBuf scan_vals; // per-block chunk totals of the cooperative scan (sg_scan.cuh)

  Ball2DData* ball2d = nullptr;
  AabbData* aabb = nullptr;
  Rb3dData* rb3d = nullptr;
  Rb2dData* rb2d = nullptr;
};

int sg_fail( sg_ctx* ctx, int code, const char* fmt, ... );
int sg_prof_entry( sg_ctx* ctx, const char* name );
void sg_prof_begin( sg_ctx* ctx, const char* name, double bytes );
void sg_prof_end( sg_ctx* ctx );
void sg_prof_collect( sg_ctx* ctx );

#define SG_CUDA( ctx, call )                                                                                   \
  do                                                                                                           \
  {                                                                                                            \
    const cudaError_t sg_e_ = ( call );                                                                        \
    if( sg_e_ != cudaSuccess )                                                                                 \
    {                                                                                                          \
      return sg_fail( ( ctx ), SG_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString( sg_e_ ) ); \
    }                                                                                                          \
  } while( 0 )

// Launch one kernel on the context's stream, counted and (optionally) timed.  `bytes` = algorithmic bytes
// this launch moves, used by bench.py for the live roofline figure.
#define SG_LAUNCH( ctx, name, bytes, ... )                                                                     \
  do                                                                                                           \
  {                                                                                                            \
    if( ( ctx )->profile ) { sg_prof_begin( ( ctx ), ( name ), double( bytes ) ); }                            \
    __VA_ARGS__;                                                                                               \
    ++( ctx )->launch_count;                                                                                   \
    if( ( ctx )->profile ) { sg_prof_end( ( ctx ) ); }                                                         \
    const cudaError_t sg_le_ = cudaPeekAtLastError();                                                          \
    if( sg_le_ != cudaSuccess )                                                                                \
    {                                                                                                          \
      return sg_fail( ( ctx ), SG_ERR_CUDA, "%s:%d: launch %s -> %s", __FILE__, __LINE__, ( name ), cudaGetErrorString( sg_le_ ) ); \
    }                                                                                                          \
  } while( 0 )

static inline unsigned sg_div_up( uint64_t a, uint64_t b ) { return unsigned( ( a + b - 1 ) / b ); }

// ---------------------------------------------------------------------------------------------------
// Device side helpers
// ---------------------------------------------------------------------------------------------------

// Monotone double <-> int64 map so that atomicMin/atomicMax on long long order doubles correctly
__host__ __device__ static inline long long sg_ordered_from_double( double x )
{
#ifdef __CUDA_ARCH__
  long long b = __double_as_longlong( x );
#else
  long long b; memcpy( &b, &x, 8 );
#endif
  return ( b >= 0 ) ? b : ( b ^ 0x7fffffffffffffffLL );
}
__host__ __device__ static inline double sg_double_from_ordered( long long b )
{
  b = ( b >= 0 ) ? b : ( b ^ 0x7fffffffffffffffLL );
#ifdef __CUDA_ARCH__
  return __longlong_as_double( b );
#else
  double x; memcpy( &x, &b, 8 ); return x;
#endif
}

// Uniform grid over AABB lower corners.  The candidate set is grid-independent: any
// conservative binning gives the reference's set.  With h >= every box extent, two overlapping boxes have
// lower corners less than h apart per axis, so their cells differ by at most 1 per axis.
struct GridParams
{
  double origin[3];
  double h;
  uint32_t dims[3];
  uint32_t ncells;
};

// Local body slot -> global body index (multi-GPU slabs: a table over all slots, owned and ghost; identity on one GPU)
struct GidMap
{
  const uint32_t* gid = nullptr;
#ifdef __CUDACC__
  __device__ __forceinline__ uint32_t operator()( const uint32_t i ) const { return ( gid == nullptr ) ? i : __ldg( &gid[i] ); }
#endif
};

// Reduced over all bodies before the grid is laid out (ordered-int encoding, see above)
struct BoundsAccum
{
  long long min_lo[3];
  long long max_lo[3];
  long long max_ext;
};

#endif
