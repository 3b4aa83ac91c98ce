// sg_boxes.cuh -- broad-phase policy for caller-built boxes (no narrow phase): n boxes [lo(DIM), hi(DIM)] in, every
// overlapping (i<j) out.  TAG only makes the instantiations of one translation unit distinct from another's (each user
// owns its kernels): 0 = sg_aabb.cu (sg_candidate_pairs), 1 = sg_ball2d.cu (real + teleported boxes of the portal path).
#ifndef SG_BOXES_CUH
#define SG_BOXES_CUH

#include "sg_broadphase.cuh"

template<int DIM>
struct AabbIn
{
  const double* boxes; // n * 2*DIM: lo(DIM), hi(DIM)
  uint32_t n;
};

template<int DIM> struct AabbRec;
template<> struct alignas( 64 ) AabbRec<2> { double lo[2]; double hi[2]; uint32_t idx; uint32_t key; uint32_t c1, c2; double pad[2]; };
template<> struct alignas( 64 ) AabbRec<3> { double lo[3]; double hi[3]; uint32_t idx; uint32_t key; uint32_t c1, c2; };

struct NoOut {};

template<int DIM, int TAG = 0>
struct AabbPolicy
{
  static constexpr int D = DIM;
  static constexpr bool HAS_NARROW = false;
  static constexpr double IN_BYTES = 16.0 * DIM;
  using In = AabbIn<DIM>;
  using Rec = AabbRec<DIM>;
  using Out = NoOut;
  static constexpr uint32_t IDX_MASK = 0xffffffffu;
  static constexpr uint32_t IDX_OFFSET = 16u * DIM;
  static constexpr uint32_t ORD_OFFSET = IDX_OFFSET; // bodies are ranked by their index
  static constexpr bool ORD_IN_REC = true;
  __device__ static void load_aabb( const In& in, const uint32_t i, double* lo, double* hi )
  {
    const double* b = in.boxes + size_t( i ) * 2 * DIM;
    #pragma unroll
    for( int k = 0; k < DIM; ++k ) { lo[k] = __ldg( b + k ); hi[k] = __ldg( b + DIM + k ); }
  }
  __device__ static Rec make_rec( const In& in, const uint32_t i, const uint32_t key, const uint32_t c1, const uint32_t c2 )
  {
    Rec r;
    load_aabb( in, i, r.lo, r.hi );
    r.idx = i; r.key = key; r.c1 = c1; r.c2 = c2;
    return r;
  }
  __device__ static void rec_aabb( const Rec& s, double* lo, double* hi )
  {
    #pragma unroll
    for( int k = 0; k < DIM; ++k ) { lo[k] = s.lo[k]; hi[k] = s.hi[k]; }
  }
  __device__ static uint32_t rec_idx( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_idx_raw( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_ord( const Rec& s ) { return rec_idx( s ); }
  __device__ static uint32_t rec_ord_raw( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_key( const Rec& s ) { return s.key; }
  __device__ static bool owns( const Rec& ) { return true; }
  __device__ static bool valid( const In&, const uint32_t ) { return true; }
  __device__ static uint32_t rec_c1( const Rec& s, const GridParams& ) { return s.c1; }
  __device__ static uint32_t rec_c2( const Rec& s, const GridParams& ) { return s.c2; }
  __device__ static Rec load_pass1( const Rec* __restrict__ p ) { return sg_load_rec_global<Rec>( p ); }
  __device__ static bool narrow_test( const Rec&, const Rec& ) { return false; }
  __device__ static void contact_emit( const Out&, unsigned long long&, const Rec&, const Rec& ) {}
};

#endif
