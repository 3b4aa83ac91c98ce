// tests/portal_kernel_harness.cpp -- TEST INFRASTRUCTURE.  Runs the product's portal KERNELS
// (scisim_b200/csrc/sg_ball2d_portal_kernels.cuh, exactly the source nvcc compiles) on the CPU: the CUDA keywords are
// defined away, blockIdx / threadIdx are plain (thread-local) variables and a "launch" is a loop over blocks and threads --
// legal for the kernels without barriers; the one kernel with __syncthreads (the shared-memory tile sort) gets one host
// thread per CUDA thread and a pthread barrier.  The launch sequence mirrors ball2d_portal_active_set_device
// (sg_ball2d_portals.cuh); the two pieces that are not portal code -- the prefix sums and the box broad phase, both
// covered by the GPU parity tests of the other paths -- are a sequential scan and an all-pairs sweep here.
// Built by tests/test_portals_cpu.py with g++ -O2 -std=c++17 -ffp-contract=off.  Nothing here is shipped.
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/scisim_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__( ... )
#define __grid_constant__
#define __shared__ static
struct EmuDim { unsigned x, y, z; };
static thread_local EmuDim blockIdx, blockDim, threadIdx, gridDim;
static pthread_barrier_t g_block_barrier;
static inline void __syncthreads() { pthread_barrier_wait( &g_block_barrier ); }
struct double2 { double x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline double2 make_double2( double x, double y ) { return double2{ x, y }; }
static inline uint4 make_uint4( unsigned x, unsigned y, unsigned z, unsigned w ) { return uint4{ x, y, z, w }; }
template<typename T> static inline T __ldg( const T* p ) { return *p; }
static inline unsigned atomicOr( unsigned* p, unsigned v ) { const unsigned o = *p; *p |= v; return o; }
static inline double __longlong_as_double( long long b ) { double x; std::memcpy( &x, &b, 8 ); return x; }

// as in sg_ball2d.cu, minus the multi-GPU index map
struct ContactOut2D
{
  uint32_t* type; uint32_t* i; uint32_t* j;
  double2* n; double2* p; double* depth;
  unsigned long long cap;
};

#include "../scisim_b200/csrc/sg_ball2d_portal_kernels.cuh"

template<typename F>
static void launch( unsigned gx, unsigned gy, unsigned block, F f )
{
  gridDim = EmuDim{ gx, gy, 1 }; blockDim = EmuDim{ block, 1, 1 };
  for( unsigned by = 0; by < gy; ++by ) for( unsigned bx = 0; bx < gx; ++bx ) for( unsigned t = 0; t < block; ++t )
  {
    blockIdx = EmuDim{ bx, by, 0 }; threadIdx = EmuDim{ t, 0, 0 };
    f();
  }
}
// kernels with __syncthreads: blocks one after another, the threads of a block as host threads
template<typename F>
static void launch_with_barriers( unsigned gx, unsigned block, F f )
{
  for( unsigned bx = 0; bx < gx; ++bx )
  {
    pthread_barrier_init( &g_block_barrier, nullptr, block );
    std::vector<std::thread> threads;
    for( unsigned t = 0; t < block; ++t )
    {
      threads.emplace_back( [=]() { gridDim = EmuDim{ gx, 1, 1 }; blockDim = EmuDim{ block, 1, 1 }; blockIdx = EmuDim{ bx, 0, 0 }; threadIdx = EmuDim{ t, 0, 0 }; f(); } );
    }
    for( std::thread& th : threads ) { th.join(); }
    pthread_barrier_destroy( &g_block_barrier );
  }
}
static unsigned div_up( unsigned long long a, unsigned long long b ) { return unsigned( ( a + b - 1 ) / b ); }
static uint32_t exclusive_scan( const std::vector<uint32_t>& in, std::vector<uint32_t>& out )
{
  uint32_t run = 0; out.resize( in.size() + 1 );
  for( size_t k = 0; k < in.size(); ++k ) { out[k] = run; run += in[k]; }
  return run;
}

// the sort's launch sequence as in ball2d_portal_active_set_device; TILE / THREADS are the library's (2048, 1024) or a
// small pair that makes multi-tile lists cheap to emulate
template<int TILE, int THREADS>
static void tile_sort( const uint32_t m, unsigned long long* keys, uint32_t* idxs )
{
  const unsigned ntiles = div_up( m, TILE );
  launch_with_barriers( ntiles, THREADS, [=]() { k_b2p_bitonic_tile<TILE, THREADS, true>( m, 0u, keys, idxs ); } );
  for( uint32_t k = 2u * TILE; k <= m; k <<= 1 )
  {
    for( uint32_t j = k >> 1; j >= uint32_t( TILE ); j >>= 1 ) { launch( div_up( m, 256 ), 1, 256, [&]() { k_b2p_bitonic( m, j, k, keys, idxs ); } ); }
    launch_with_barriers( ntiles, THREADS, [=]() { k_b2p_bitonic_tile<TILE, THREADS, false>( m, k, keys, idxs ); } );
  }
}
static int g_sort_mode = 0; // 0: tiles of 64 / 16 threads, 1: the library's 2048 / 1024
static void sort_teleported( const uint32_t m, unsigned long long* keys, uint32_t* idxs )
{
  if( g_sort_mode == 1 ) { tile_sort<2048, 1024>( m, keys, idxs ); } else { tile_sort<64, 16>( m, keys, idxs ); }
}

struct Result
{
  std::vector<uint2> cand;
  std::vector<uint32_t> box_body, box_portal, type, ci, cj, tp0, tp1;
  std::vector<double2> cn, cp, x0t, x1t, kick;
  std::vector<double> depth;
  uint32_t n_reg = 0, n_tel = 0;
  int err = 0;
};
static Result g_res;
static SgPortals2D g_ps;

extern "C"
{

void pk_set_portals( uint32_t n, const double* ax, const double* an, const double* bx, const double* bn, const double* v, const double* bounds, const double* dx )
{
  std::memset( &g_ps, 0, sizeof( g_ps ) );
  g_ps.n = n;
  for( uint32_t p = 0; p < n; ++p )
  {
    SgPortal2D& pt = g_ps.p[p];
    for( int k = 0; k < 2; ++k ) { pt.ax[k] = ax[2 * p + k]; pt.bx[k] = bx[2 * p + k]; }
    sg_portal_plane_frame( an + 2 * p, pt.an, pt.at );
    sg_portal_plane_frame( bn + 2 * p, pt.bn, pt.bt );
    pt.v = v[p]; pt.bounds = bounds[p]; pt.dx = dx[p];
  }
}

// body-body part of the portal active set; returns 0, or 1 where a ball touches both planes of a portal
int pk_active_set( uint32_t n, const double* q0v, const double* q1v, const double* r )
{
  Result& R = g_res;
  R = Result{};
  const double2* q0 = reinterpret_cast<const double2*>( q0v );
  const double2* q1 = reinterpret_cast<const double2*>( q1v );
  const uint32_t P = g_ps.n;
  const unsigned nblk = div_up( n, 256 );
  std::vector<uint32_t> tflag( size_t( n ) * P + 1, 0xdeadbeefu ), toff;
  unsigned err = 0;
  launch( nblk, P, 256, [&]() { k_b2p_touch( g_ps, n, q1, r, tflag.data(), &err ); } );
  if( err != 0 ) { R.err = 1; return 1; }
  tflag.resize( size_t( n ) * P );
  const uint32_t nt = exclusive_scan( tflag, toff );
  const uint32_t next = n + nt;
  std::vector<double> boxes( size_t( next ) * 4, std::nan( "" ) );
  R.box_body.assign( nt, 0xdeadbeefu ); R.box_portal.assign( nt, 0xdeadbeefu );
  launch( nblk, 1, 256, [&]() { k_b2p_boxes( n, q1, r, boxes.data() ); } );
  if( nt > 0 ) { launch( nblk, P, 256, [&]() { k_b2p_tele_boxes( g_ps, n, q1, r, tflag.data(), toff.data(), boxes.data(), R.box_body.data(), R.box_portal.data() ); } ); }
  // stand-in for the box pipeline: every (i<j) whose closed boxes overlap, ascending
  for( uint32_t i = 0; i < next; ++i ) for( uint32_t j = i + 1; j < next; ++j )
  {
    const double* a = &boxes[4 * size_t( i )]; const double* b = &boxes[4 * size_t( j )];
    if( !( a[2] < b[0] ) && !( b[2] < a[0] ) && !( a[3] < b[1] ) && !( b[3] < a[1] ) ) { R.cand.push_back( uint2{ i, j } ); }
  }
  const unsigned long long np = R.cand.size();
  std::vector<uint32_t> reg_cnt( np, 0xdeadbeefu ), tel_cnt( np, 0xdeadbeefu ), reg_off, tel_off;
  const ContactOut2D none{};
  if( np > 0 )
  {
    launch( div_up( np, 128 ), 1, 128, [&]() { k_b2p_pairs<false>( g_ps, n, R.cand.data(), np, q0, q1, r, R.box_body.data(), R.box_portal.data(), reg_cnt.data(), tel_cnt.data(), nullptr, nullptr, none, nullptr, nullptr, nullptr ); } );
  }
  R.n_reg = exclusive_scan( reg_cnt, reg_off );
  const uint32_t nraw = exclusive_scan( tel_cnt, tel_off );
  const size_t cap = size_t( R.n_reg ) + nraw + 64;
  R.type.assign( cap, 0xdeadbeefu ); R.ci.assign( cap, 0xdeadbeefu ); R.cj.assign( cap, 0xdeadbeefu );
  R.cn.assign( cap, double2{ 0, 0 } ); R.cp.assign( cap, double2{ 0, 0 } ); R.depth.assign( cap, 0.0 );
  ContactOut2D out;
  out.type = R.type.data(); out.i = R.ci.data(); out.j = R.cj.data(); out.n = R.cn.data(); out.p = R.cp.data(); out.depth = R.depth.data(); out.cap = cap;
  uint32_t m = 1u;
  while( m < nraw ) { m <<= 1; }
  std::vector<unsigned long long> tc_key( m, 0x1234ull );
  std::vector<uint32_t> tc_idx( m, 0xdeadbeefu ), uflag( nraw, 0xdeadbeefu ), uoff;
  std::vector<uint4> tc_info( nraw + 1 );
  if( np > 0 && ( R.n_reg > 0 || nraw > 0 ) )
  {
    launch( div_up( np, 128 ), 1, 128, [&]() { k_b2p_pairs<true>( g_ps, n, R.cand.data(), np, q0, q1, r, R.box_body.data(), R.box_portal.data(), reg_cnt.data(), tel_cnt.data(), reg_off.data(), tel_off.data(), out,
                                                                  tc_key.data(), tc_idx.data(), tc_info.data() ); } );
  }
  if( nraw > 0 )
  {
    if( m > nraw ) { launch( div_up( m - nraw, 256 ), 1, 256, [&]() { k_b2p_sort_pad( nraw, m, tc_key.data(), tc_idx.data() ); } ); }
    sort_teleported( m, tc_key.data(), tc_idx.data() );
    launch( div_up( nraw, 256 ), 1, 256, [&]() { k_b2p_unique( nraw, tc_key.data(), uflag.data() ); } );
    R.n_tel = exclusive_scan( uflag, uoff );
    R.x0t.assign( nraw, double2{ 0, 0 } ); R.x1t.assign( nraw, double2{ 0, 0 } ); R.kick.assign( nraw, double2{ 0, 0 } ); R.tp0.assign( nraw, 0 ); R.tp1.assign( nraw, 0 );
    launch( div_up( nraw, 128 ), 1, 128, [&]() { k_b2p_tele_contacts( g_ps, nraw, tc_idx.data(), uflag.data(), uoff.data(), tc_info.data(), q0, q1, r, R.n_reg, out, R.x0t.data(), R.x1t.data(), R.kick.data(),
                                                                       R.tp0.data(), R.tp1.data() ); } );
  }
  return 0;
}

uint64_t pk_num_candidates() { return g_res.cand.size(); }
uint32_t pk_num_boxes() { return uint32_t( g_res.box_body.size() ); }
uint32_t pk_num_regular() { return g_res.n_reg; }
uint32_t pk_num_teleported() { return g_res.n_tel; }
void pk_copy( uint32_t* cand, uint32_t* box_body, uint32_t* box_portal, uint32_t* type, uint32_t* i, uint32_t* j, double* n, double* p, double* depth, uint32_t* tp0, uint32_t* tp1, double* x0, double* x1,
              double* kick )
{
  const Result& R = g_res;
  std::memcpy( cand, R.cand.data(), R.cand.size() * 8 );
  std::memcpy( box_body, R.box_body.data(), R.box_body.size() * 4 ); std::memcpy( box_portal, R.box_portal.data(), R.box_portal.size() * 4 );
  const size_t na = size_t( R.n_reg ) + R.n_tel;
  std::memcpy( type, R.type.data(), na * 4 ); std::memcpy( i, R.ci.data(), na * 4 ); std::memcpy( j, R.cj.data(), na * 4 );
  std::memcpy( n, R.cn.data(), na * 16 ); std::memcpy( p, R.cp.data(), na * 16 ); std::memcpy( depth, R.depth.data(), na * 8 );
  std::memcpy( tp0, R.tp0.data(), size_t( R.n_tel ) * 4 ); std::memcpy( tp1, R.tp1.data(), size_t( R.n_tel ) * 4 );
  std::memcpy( x0, R.x0t.data(), size_t( R.n_tel ) * 16 ); std::memcpy( x1, R.x1t.data(), size_t( R.n_tel ) * 16 ); std::memcpy( kick, R.kick.data(), size_t( R.n_tel ) * 16 );
}

void pk_set_sort_mode( int mode ) { g_sort_mode = mode; }
void pk_sort( uint32_t m, unsigned long long* keys, uint32_t* idxs ) { sort_teleported( m, keys, idxs ); }

void pk_enforce( uint32_t n, double* q, double* v )
{
  launch( div_up( n, 256 ), 1, 256, [&]() { k_b2p_enforce( g_ps, n, reinterpret_cast<double2*>( q ), reinterpret_cast<double2*>( v ) ); } );
}

// ---- rigidbody2d portal kernels (scisim_b200/csrc/sg_rb2d_portal_kernels.cuh) ---------------------------------------------
}

// what sg_rb2d.cu provides to that header
#define SG_FIXED_BIT2 0x80000000u
#define SG_GEO2_CIRCLE 0u
#define SG_GEO2_BOX 1u
struct V2d { double x, y; };
struct M2d { double a, b, c, d; };
static inline M2d rot2d( const double theta ) { const double s = std::sin( theta ), c = std::cos( theta ); M2d R; R.a = c; R.b = -s; R.c = s; R.d = c; return R; }
struct Rb2dDev { uint32_t n; const uint32_t* btype; const double2* bparam; };
struct ContactOut2X
{
  uint32_t* type; uint32_t* i; uint32_t* j; uint32_t* aux;
  double2* n; double2* p; double* depth;
  unsigned long long cap;
};
static inline void put2( const ContactOut2X& out, const unsigned long long k, const uint32_t type, const uint32_t i, const uint32_t j, const uint32_t aux, const V2d n, const V2d p, const double depth )
{
  if( k >= out.cap ) { return; }
  out.type[k] = type; out.i[k] = i; out.j[k] = j; out.aux[k] = aux;
  out.n[k] = make_double2( n.x, n.y ); out.p[k] = make_double2( p.x, p.y ); out.depth[k] = depth;
}

#include "../scisim_b200/csrc/sg_rb2d_portal_kernels.cuh"

struct ResultRB2D
{
  std::vector<uint2> cand, reg_pairs;
  std::vector<uint32_t> box_body, box_portal, type, ci, cj, aux, tp0, tp1;
  std::vector<double2> cn, cp, x0t, x1t, d0, d1, kick;
  std::vector<double> depth;
  uint32_t n_tel = 0;
  unsigned bad = 0;
};
static ResultRB2D g_r2;

extern "C"
{

void pk_rb2d_set_portals( uint32_t n, const double* ax, const double* an, const double* bx, const double* bn, const double* v, const double* bounds, const double* dx )
{
  std::memset( &g_ps, 0, sizeof( g_ps ) );
  g_ps.n = n;
  for( uint32_t p = 0; p < n; ++p )
  {
    SgPortal2D& pt = g_ps.p[p];
    for( int k = 0; k < 2; ++k ) { pt.ax[k] = ax[2 * p + k]; pt.bx[k] = bx[2 * p + k]; }
    sg_portal_plane_frame_as_given( an + 2 * p, pt.an, pt.at );
    sg_portal_plane_frame_as_given( bn + 2 * p, pt.bn, pt.bt );
    pt.v = v[p]; pt.bounds = bounds[p]; pt.dx = dx[p];
  }
}

// the portal-specific part of rb2d_portal_active_set_device: candidates, un-teleported pair list, teleported contacts (written
// from index 0: base = 0).  btype: geometry type | SG_FIXED_BIT2; bparam: ( r, - ) or box half widths.  Returns the bad-flag word.
unsigned pk_rb2d_active_set( uint32_t n, const uint32_t* btype, const double* bparam, const double* q0, const double* q1 )
{
  ResultRB2D& R = g_r2;
  R = ResultRB2D{};
  Rb2dDev dev; dev.n = n; dev.btype = btype; dev.bparam = reinterpret_cast<const double2*>( bparam );
  const uint32_t P = g_ps.n;
  const unsigned nblk = div_up( n, 256 );
  std::vector<double> rboxes( size_t( n ) * 4, std::nan( "" ) );
  launch( nblk, 1, 256, [&]() { k_r2p_boxes( dev, q1, rboxes.data() ); } );
  std::vector<uint32_t> tflag( size_t( n ) * P, 0xdeadbeefu ), toff;
  launch( nblk, P, 256, [&]() { k_r2p_touch( g_ps, n, rboxes.data(), tflag.data() ); } );
  const uint32_t nt = exclusive_scan( tflag, toff );
  const uint32_t next = n + nt;
  std::vector<double> boxes( size_t( next ) * 4, std::nan( "" ) );
  std::copy( rboxes.begin(), rboxes.end(), boxes.begin() );
  R.box_body.assign( nt, 0xdeadbeefu ); R.box_portal.assign( nt, 0xdeadbeefu );
  if( nt > 0 ) { launch( nblk, P, 256, [&]() { k_r2p_tele_boxes( g_ps, dev, q1, rboxes.data(), tflag.data(), toff.data(), boxes.data(), R.box_body.data(), R.box_portal.data() ); } ); }
  for( uint32_t i = 0; i < next; ++i ) for( uint32_t j = i + 1; j < next; ++j )
  {
    const double* a = &boxes[4 * size_t( i )]; const double* b = &boxes[4 * size_t( j )];
    if( !( a[2] < b[0] ) && !( b[2] < a[0] ) && !( a[3] < b[1] ) && !( b[3] < a[1] ) ) { R.cand.push_back( uint2{ i, j } ); }
  }
  const unsigned long long np = R.cand.size();
  std::vector<uint32_t> reg_cnt( np, 0xdeadbeefu ), tel_cnt( np, 0xdeadbeefu ), reg_off32, tel_off;
  if( np > 0 )
  {
    launch( div_up( np, 128 ), 1, 128, [&]() { k_r2p_classify<false>( g_ps, dev, R.cand.data(), np, q1, R.box_body.data(), R.box_portal.data(), reg_cnt.data(), tel_cnt.data(), nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                      nullptr, &R.bad ); } );
  }
  if( R.bad != 0u ) { return R.bad; }
  const uint32_t nreg = exclusive_scan( reg_cnt, reg_off32 );
  const uint32_t nraw = exclusive_scan( tel_cnt, tel_off );
  std::vector<unsigned long long> reg_off( reg_off32.begin(), reg_off32.end() );
  R.reg_pairs.assign( nreg, uint2{ 0xdeadbeefu, 0xdeadbeefu } );
  uint32_t m = 1u;
  while( m < nraw ) { m <<= 1; }
  std::vector<unsigned long long> tc_key( m, 0x1234ull );
  std::vector<uint32_t> tc_idx( m, 0xdeadbeefu ), uflag( nraw, 0xdeadbeefu ), uoff;
  std::vector<uint4> tc_info( nraw + 1 );
  if( np > 0 )
  {
    launch( div_up( np, 128 ), 1, 128, [&]() { k_r2p_classify<true>( g_ps, dev, R.cand.data(), np, q1, R.box_body.data(), R.box_portal.data(), reg_cnt.data(), tel_cnt.data(), reg_off.data(), tel_off.data(),
                                                                     R.reg_pairs.data(), tc_key.data(), tc_idx.data(), tc_info.data(), &R.bad ); } );
  }
  if( nraw > 0 )
  {
    if( m > nraw ) { launch( div_up( m - nraw, 256 ), 1, 256, [&]() { k_b2p_sort_pad( nraw, m, tc_key.data(), tc_idx.data() ); } ); }
    sort_teleported( m, tc_key.data(), tc_idx.data() );
    launch( div_up( nraw, 256 ), 1, 256, [&]() { k_b2p_unique( nraw, tc_key.data(), uflag.data() ); } );
    R.n_tel = exclusive_scan( uflag, uoff );
    const size_t cap = size_t( R.n_tel ) + 8;
    R.type.assign( cap, 0xdeadbeefu ); R.ci.assign( cap, 0xdeadbeefu ); R.cj.assign( cap, 0xdeadbeefu ); R.aux.assign( cap, 0xdeadbeefu );
    R.cn.assign( cap, double2{ 0, 0 } ); R.cp.assign( cap, double2{ 0, 0 } ); R.depth.assign( cap, -1.0 );
    for( std::vector<double2>* v : { &R.x0t, &R.x1t, &R.d0, &R.d1, &R.kick } ) { v->assign( nraw, double2{ 0, 0 } ); }
    R.tp0.assign( nraw, 0 ); R.tp1.assign( nraw, 0 );
    ContactOut2X out;
    out.type = R.type.data(); out.i = R.ci.data(); out.j = R.cj.data(); out.aux = R.aux.data(); out.n = R.cn.data(); out.p = R.cp.data(); out.depth = R.depth.data(); out.cap = cap;
    launch( div_up( nraw, 128 ), 1, 128, [&]() { k_r2p_tele_contacts( g_ps, dev, nraw, tc_idx.data(), uflag.data(), uoff.data(), tc_info.data(), q0, q1, 0ull, out, R.x0t.data(), R.x1t.data(), R.d0.data(), R.d1.data(),
                                                                       R.kick.data(), R.tp0.data(), R.tp1.data(), &R.bad ); } );
  }
  return R.bad;
}
uint64_t pk_rb2d_num_candidates() { return g_r2.cand.size(); }
uint64_t pk_rb2d_num_regular_pairs() { return g_r2.reg_pairs.size(); }
uint32_t pk_rb2d_num_boxes() { return uint32_t( g_r2.box_body.size() ); }
uint32_t pk_rb2d_num_teleported() { return g_r2.n_tel; }
void pk_rb2d_copy( uint32_t* cand, uint32_t* reg_pairs, uint32_t* box_body, uint32_t* box_portal, uint32_t* type, uint32_t* i, uint32_t* j, double* n, double* p, double* depth, uint32_t* tp0, uint32_t* tp1,
                   double* x0, double* x1, double* d0, double* d1, double* kick )
{
  const ResultRB2D& R = g_r2;
  std::memcpy( cand, R.cand.data(), R.cand.size() * 8 ); std::memcpy( reg_pairs, R.reg_pairs.data(), R.reg_pairs.size() * 8 );
  std::memcpy( box_body, R.box_body.data(), R.box_body.size() * 4 ); std::memcpy( box_portal, R.box_portal.data(), R.box_portal.size() * 4 );
  const size_t nt = R.n_tel;
  if( nt == 0 ) { return; }
  std::memcpy( type, R.type.data(), nt * 4 ); std::memcpy( i, R.ci.data(), nt * 4 ); std::memcpy( j, R.cj.data(), nt * 4 );
  std::memcpy( n, R.cn.data(), nt * 16 ); std::memcpy( p, R.cp.data(), nt * 16 ); std::memcpy( depth, R.depth.data(), nt * 8 );
  std::memcpy( tp0, R.tp0.data(), nt * 4 ); std::memcpy( tp1, R.tp1.data(), nt * 4 );
  std::memcpy( x0, R.x0t.data(), nt * 16 ); std::memcpy( x1, R.x1t.data(), nt * 16 ); std::memcpy( d0, R.d0.data(), nt * 16 ); std::memcpy( d1, R.d1.data(), nt * 16 );
  std::memcpy( kick, R.kick.data(), nt * 16 );
}
void pk_rb2d_enforce( uint32_t n, double* q, double* v )
{
  launch( div_up( n, 256 ), 1, 256, [&]() { k_r2p_enforce( g_ps, n, q, v ); } );
}
// aabbTouchesPortal / teleportPoint / getKinematicVelocityOfAABB through the product header, layout of orc_rb2d_portal_probe
uint32_t pk_rb2d_probe( uint32_t p, const double* box, const double* x, double* out )
{
  const int touch = sg_portal_aabb_touch( g_ps.p[p], box, box + 2 );
  const SgVec2 xo = sg_portal_teleport( g_ps.p[p], touch == 2, SgVec2{ x[0], x[1] } );
  const SgVec2 k = sg_portal_kinematic_velocity_of_aabb( g_ps.p[p], box, box + 2 );
  out[0] = xo.x; out[1] = xo.y; out[2] = k.x; out[3] = k.y;
  return uint32_t( touch );
}

// ---- rigidbody3d portal kernels (scisim_b200/csrc/sg_rb3d_portal_kernels.cuh), sphere scenes --------------------------------
}

// what sg_rb3d.cu provides to that header
#define SG_FIXED_BIT 0x80000000u
struct Rb3dDev { uint32_t n; const uint32_t* btype; const double* bparam; };
struct ContactOut3D
{
  uint32_t* type; uint32_t* i; uint32_t* j; uint32_t* aux;
  double* n; double* p; double* depth;
  unsigned long long cap;
};

#include "../scisim_b200/csrc/sg_rb3d_portal_kernels.cuh"

struct ResultRB3D
{
  std::vector<uint2> cand, reg_pairs;
  std::vector<uint32_t> box_body, box_portal, type, ci, cj, aux, tp0, tp1;
  std::vector<double> cn, cp, depth, x0t, x1t;
  uint32_t n_tel = 0;
};
static ResultRB3D g_r3;
static SgPortals3D g_ps3;

extern "C"
{

int pk_rb3d_set_portals( uint32_t n, const double* ax, const double* an, const double* bx, const double* bn, const int* mult )
{
  std::memset( &g_ps3, 0, sizeof( g_ps3 ) );
  g_ps3.n = n;
  int ok = 1;
  for( uint32_t p = 0; p < n; ++p )
  {
    SgPortal3D& pt = g_ps3.p[p];
    for( int k = 0; k < 3; ++k ) { pt.ax[k] = ax[3 * p + k]; pt.bx[k] = bx[3 * p + k]; pt.mult[k] = mult[3 * p + k]; }
    ok = ok && sg_portal3_plane_frame( an + 3 * p, pt.an, pt.at0, pt.at1 ) && sg_portal3_plane_frame( bn + 3 * p, pt.bn, pt.bt0, pt.bt1 );
  }
  return ok;
}

// the portal-specific part of rb3d_portal_active_set_device for an all-sphere scene: candidates, un-teleported pair list, teleported
// contacts (written from index 0).  btype: SG_FIXED_BIT for kinematic spheres; bparam: 4 doubles per body, radius first
void pk_rb3d_active_set( uint32_t n, const uint32_t* btype, const double* bparam, const double* q0, const double* q1 )
{
  ResultRB3D& R = g_r3;
  R = ResultRB3D{};
  Rb3dDev dev; dev.n = n; dev.btype = btype; dev.bparam = bparam;
  const uint32_t P = g_ps3.n;
  const unsigned nblk = div_up( n, 256 );
  // real boxes: RigidBodySphere::computeAABB ( k_rb3d_aabb in the library )
  std::vector<double> rboxes( size_t( n ) * 6 );
  for( uint32_t b = 0; b < n; ++b ) { for( int k = 0; k < 3; ++k ) { rboxes[6 * size_t( b ) + k] = q1[3 * size_t( b ) + k] - bparam[4 * size_t( b )]; rboxes[6 * size_t( b ) + 3 + k] = q1[3 * size_t( b ) + k] + bparam[4 * size_t( b )]; } }
  std::vector<uint32_t> tflag( size_t( n ) * P, 0xdeadbeefu ), toff;
  launch( nblk, P, 256, [&]() { k_r3p_touch( g_ps3, n, rboxes.data(), tflag.data() ); } );
  const uint32_t nt = exclusive_scan( tflag, toff );
  const uint32_t next = n + nt;
  std::vector<double> boxes( size_t( next ) * 6, std::nan( "" ) );
  std::copy( rboxes.begin(), rboxes.end(), boxes.begin() );
  R.box_body.assign( nt, 0xdeadbeefu ); R.box_portal.assign( nt, 0xdeadbeefu );
  if( nt > 0 ) { launch( nblk, P, 256, [&]() { k_r3p_tele_boxes( g_ps3, dev, q1, rboxes.data(), tflag.data(), toff.data(), boxes.data(), R.box_body.data(), R.box_portal.data() ); } ); }
  for( uint32_t i = 0; i < next; ++i ) for( uint32_t j = i + 1; j < next; ++j )
  {
    const double* a = &boxes[6 * size_t( i )]; const double* b = &boxes[6 * size_t( j )];
    bool ov = true;
    for( int k = 0; k < 3; ++k ) { if( a[3 + k] < b[k] || b[3 + k] < a[k] ) { ov = false; } }
    if( ov ) { R.cand.push_back( uint2{ i, j } ); }
  }
  const unsigned long long np = R.cand.size();
  std::vector<uint32_t> reg_cnt( np, 0xdeadbeefu ), tel_cnt( np, 0xdeadbeefu ), reg_off32, tel_off;
  if( np > 0 )
  {
    launch( div_up( np, 128 ), 1, 128, [&]() { k_r3p_classify<false>( g_ps3, dev, R.cand.data(), np, q1, R.box_body.data(), R.box_portal.data(), reg_cnt.data(), tel_cnt.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr ); } );
  }
  const uint32_t nreg = exclusive_scan( reg_cnt, reg_off32 );
  const uint32_t nraw = exclusive_scan( tel_cnt, tel_off );
  std::vector<unsigned long long> reg_off( reg_off32.begin(), reg_off32.end() );
  R.reg_pairs.assign( nreg, uint2{ 0xdeadbeefu, 0xdeadbeefu } );
  uint32_t m = 1u;
  while( m < nraw ) { m <<= 1; }
  std::vector<unsigned long long> tc_key( m, 0x1234ull );
  std::vector<uint32_t> tc_idx( m, 0xdeadbeefu ), uflag( nraw, 0xdeadbeefu ), uoff;
  std::vector<uint4> tc_info( nraw + 1 );
  if( np > 0 )
  {
    launch( div_up( np, 128 ), 1, 128, [&]() { k_r3p_classify<true>( g_ps3, dev, R.cand.data(), np, q1, R.box_body.data(), R.box_portal.data(), reg_cnt.data(), tel_cnt.data(), reg_off.data(), tel_off.data(),
                                                                     R.reg_pairs.data(), tc_key.data(), tc_idx.data(), tc_info.data() ); } );
  }
  if( nraw > 0 )
  {
    if( m > nraw ) { launch( div_up( m - nraw, 256 ), 1, 256, [&]() { k_b2p_sort_pad( nraw, m, tc_key.data(), tc_idx.data() ); } ); }
    sort_teleported( m, tc_key.data(), tc_idx.data() );
    launch( div_up( nraw, 256 ), 1, 256, [&]() { k_b2p_unique( nraw, tc_key.data(), uflag.data() ); } );
    R.n_tel = exclusive_scan( uflag, uoff );
    const size_t cap = size_t( R.n_tel ) + 8;
    R.type.assign( cap, 0xdeadbeefu ); R.ci.assign( cap, 0xdeadbeefu ); R.cj.assign( cap, 0xdeadbeefu ); R.aux.assign( cap, 0xdeadbeefu );
    R.cn.assign( 3 * cap, 0.0 ); R.cp.assign( 3 * cap, 0.0 ); R.depth.assign( cap, -1.0 );
    R.x0t.assign( 3 * size_t( nraw ), 0.0 ); R.x1t.assign( 3 * size_t( nraw ), 0.0 ); R.tp0.assign( nraw, 0 ); R.tp1.assign( nraw, 0 );
    ContactOut3D out;
    out.type = R.type.data(); out.i = R.ci.data(); out.j = R.cj.data(); out.aux = R.aux.data(); out.n = R.cn.data(); out.p = R.cp.data(); out.depth = R.depth.data(); out.cap = cap;
    launch( div_up( nraw, 128 ), 1, 128, [&]() { k_r3p_tele_contacts( g_ps3, dev, nraw, tc_idx.data(), uflag.data(), uoff.data(), tc_info.data(), q0, 0ull, out, R.x0t.data(), R.x1t.data(), R.tp0.data(), R.tp1.data() ); } );
  }
}
uint64_t pk_rb3d_num_candidates() { return g_r3.cand.size(); }
uint64_t pk_rb3d_num_regular_pairs() { return g_r3.reg_pairs.size(); }
uint32_t pk_rb3d_num_boxes() { return uint32_t( g_r3.box_body.size() ); }
uint32_t pk_rb3d_num_teleported() { return g_r3.n_tel; }
void pk_rb3d_copy( uint32_t* cand, uint32_t* reg_pairs, uint32_t* box_body, uint32_t* box_portal, uint32_t* type, uint32_t* i, uint32_t* j, double* n, double* p, double* depth, uint32_t* tp0, uint32_t* tp1,
                   double* x0, double* x1 )
{
  const ResultRB3D& R = g_r3;
  std::memcpy( cand, R.cand.data(), R.cand.size() * 8 ); std::memcpy( reg_pairs, R.reg_pairs.data(), R.reg_pairs.size() * 8 );
  std::memcpy( box_body, R.box_body.data(), R.box_body.size() * 4 ); std::memcpy( box_portal, R.box_portal.data(), R.box_portal.size() * 4 );
  const size_t nt = R.n_tel;
  if( nt == 0 ) { return; }
  std::memcpy( type, R.type.data(), nt * 4 ); std::memcpy( i, R.ci.data(), nt * 4 ); std::memcpy( j, R.cj.data(), nt * 4 );
  std::memcpy( n, R.cn.data(), nt * 24 ); std::memcpy( p, R.cp.data(), nt * 24 ); std::memcpy( depth, R.depth.data(), nt * 8 );
  std::memcpy( tp0, R.tp0.data(), nt * 4 ); std::memcpy( tp1, R.tp1.data(), nt * 4 );
  std::memcpy( x0, R.x0t.data(), nt * 24 ); std::memcpy( x1, R.x1t.data(), nt * 24 );
}
void pk_rb3d_enforce( uint32_t n, double* q )
{
  launch( div_up( n, 256 ), 1, 256, [&]() { k_r3p_enforce( g_ps3, n, q ); } );
}

}
