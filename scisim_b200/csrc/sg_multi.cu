// sg_multi.cu -- one process, N GPUs: the ball2d hot path of a whole scene, partitioned into x-slabs,
// behind the same calls a single context offers (set_bodies / flow / active_set / upload / step / fetch).
//
// This is what a SCISim process -- single-threaded, one address space -- links against to use
// more than one GPU; SURVEY.md 8b sketched it as sg_create( ctx, n_gpus, devices ).  It is written entirely on top of
// the public per-context C ABI (sg_ball2d_slab_*, sg_slab_*), one worker thread per GPU:
//   partition   equal-count x-quantiles of the bodies' positions (arbitrary numbering), owned bodies of a slab kept in
//               ascending global index, re-done when a body leaves its slab's neighbourhood (SG_ERR_REBALANCE), when a halo
//               outgrows its slots, or every `rebalance_every` uploads
//   step        every slab: flow -> interval to the neighbours' mailboxes -> halo records written straight into the
//               neighbours' memory over NVLink by the pack kernel -> detection; no collective, no host round trip
//   fetch       per-slab lists (global indices, each ascending) merged into the reference's order:
//               ball-ball ascending (i,j) (the std::set order of ball2d/Ball2DSim.cpp:580), then drums, then planes
#include "../../include/scisim_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace
{

struct PhaseBarrier
{
  std::mutex mu;
  std::condition_variable cv;
  int n = 1, waiting = 0;
  uint64_t gen = 0;
  void wait()
  {
    std::unique_lock<std::mutex> lk( mu );
    const uint64_t g = gen;
    if( ++waiting == n ) { waiting = 0; ++gen; cv.notify_all(); return; }
    cv.wait( lk, [&] { return gen != g; } );
  }
};

struct Slab
{
  sg_ctx* ctx = nullptr;
  int device = 0;
  std::vector<uint32_t> gid;          // owned bodies, ascending global index
  double* q = nullptr;                // pinned staging, 2 * cap doubles each
  double* v = nullptr;
  double* q1 = nullptr;
  double* v1 = nullptr;
  size_t stage_cap = 0;
  sg_contacts res;                    // last fetch (pointers into the context's pinned memory)
  uint64_t n_cand = 0, n_active = 0, n_bb = 0;
  uint32_t ghosts[2] = { 0u, 0u };
  std::vector<uint64_t> dest_cand, dest_bb, dest_static;
  int rc = SG_OK;
};

} // namespace

struct sg_multi
{
  int W = 0;
  std::vector<Slab> slab;
  bool shared_device = false; // two slabs on one GPU (testing): phases are separated by stream syncs
  std::string err;
  // scene
  uint32_t n = 0;
  std::vector<double> r, m;
  double g[2] = { 0.0, 0.0 };
  std::vector<double> plane_x, plane_n, drum_x, drum_r;
  // partition
  bool partitioned = false;
  std::vector<uint32_t> rank_of;
  std::vector<double> cuts;
  uint32_t ghost_cap = 0;
  uint32_t rebalance_every = 0; // uploads between forced re-partitions (0: only when needed)
  uint32_t uploads_since_partition = 0;
  uint64_t n_partitions = 0;
  bool state_staged = false;    // the slabs' staging buffers hold q0, v0 of the bodies they own
  bool flow_pending = false;    // sg_multi_ball2d_flow ran: the slabs hold q0, q1 and this step's interval has been posted
  int last_map = SG_MAP_SYMPLECTIC_EULER;
  double last_dt = 0.0;
  bool have_result = false;
  // merged output
  std::vector<uint32_t> o_type, o_i, o_j, o_cand;
  std::vector<double> o_n, o_p, o_depth;
  // workers
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  uint64_t job_gen = 0;
  int pending = 0;
  bool quit = false;
  std::function<int( int )> job;
  PhaseBarrier phase;
};

static std::string g_multi_create_error;

static int mfail( sg_multi* m, int code, const char* fmt, ... )
{
  char buf[1024];
  va_list ap;
  va_start( ap, fmt );
  vsnprintf( buf, sizeof( buf ), fmt, ap );
  va_end( ap );
  if( m != nullptr ) { m->err = buf; } else { g_multi_create_error = buf; }
  return code;
}

static void worker_main( sg_multi* m, const int k )
{
  cudaSetDevice( m->slab[k].device );
  uint64_t seen = 0;
  for( ;; )
  {
    std::function<int( int )> job;
    {
      std::unique_lock<std::mutex> lk( m->mu );
      m->cv_go.wait( lk, [&] { return m->quit || m->job_gen != seen; } );
      if( m->quit ) { return; }
      seen = m->job_gen;
      job = m->job;
    }
    const int rc = job( k );
    {
      std::lock_guard<std::mutex> lk( m->mu );
      m->slab[k].rc = rc;
      if( --m->pending == 0 ) { m->cv_done.notify_all(); }
    }
  }
}

// runs job( k ) on every slab's worker thread; returns the first non-zero code (REBALANCE wins over other slabs' success)
static int run_all( sg_multi* m, std::function<int( int )> job )
{
  {
    std::lock_guard<std::mutex> lk( m->mu );
    m->job = std::move( job );
    m->pending = m->W;
    ++m->job_gen;
  }
  m->cv_go.notify_all();
  {
    std::unique_lock<std::mutex> lk( m->mu );
    m->cv_done.wait( lk, [&] { return m->pending == 0; } );
  }
  int rc = SG_OK;
  for( int k = 0; k < m->W; ++k )
  {
    const int rk = m->slab[k].rc;
    if( rk == SG_OK ) { continue; }
    if( rc == SG_OK || ( rk != SG_ERR_REBALANCE && rc == SG_ERR_REBALANCE ) )
    {
      rc = rk;
      const char* e = sg_last_error( m->slab[k].ctx );
      m->err = std::string( "slab " ) + std::to_string( k ) + ": " + ( e != nullptr ? e : "" );
    }
  }
  return rc;
}

static int ensure_stage( sg_multi* m, Slab& s, const size_t n_owned )
{
  if( n_owned <= s.stage_cap ) { return SG_OK; }
  for( double** p : { &s.q, &s.v, &s.q1, &s.v1 } ) { if( *p != nullptr ) { sg_host_free( s.ctx, *p ); *p = nullptr; } }
  const size_t cap = n_owned + n_owned / 8 + 64;
  for( double** p : { &s.q, &s.v, &s.q1, &s.v1 } )
  {
    void* h = nullptr;
    const int rc = sg_host_alloc( s.ctx, uint64_t( cap ) * 16, &h );
    if( rc != SG_OK ) { return mfail( m, rc, "pinned staging: %s", sg_last_error( s.ctx ) ); }
    *p = static_cast<double*>( h );
  }
  s.stage_cap = cap;
  return SG_OK;
}

// (re)partitions the scene by the x-quantiles of q and rebuilds every slab: bodies, global indices, limits, static
// geometry, mailboxes.  q: global positions (2n).
static int partition_scene( sg_multi* m, const double* q )
{
  const uint32_t n = m->n, W = uint32_t( m->W );
  m->rank_of.resize( n );
  m->cuts.assign( W + 1u, 0.0 );
  int rc = sg_slab_partition( n, q, 2u, W, m->rank_of.data(), m->cuts.data() );
  if( rc != SG_OK ) { return mfail( m, rc, "sg_slab_partition failed" ); }
  for( auto& s : m->slab ) { s.gid.clear(); }
  for( uint32_t i = 0; i < n; ++i ) { m->slab[m->rank_of[i]].gid.push_back( i ); } // ascending by construction
  uint32_t largest = 0u;
  for( auto& s : m->slab ) { largest = std::max<uint32_t>( largest, uint32_t( s.gid.size() ) ); }
  if( m->ghost_cap == 0u ) { m->ghost_cap = std::max<uint32_t>( 4096u, largest / 32u ); }
  rc = run_all( m, [m]( const int k ) -> int
  {
    Slab& s = m->slab[k];
    const uint32_t no = uint32_t( s.gid.size() );
    int r2 = ensure_stage( m, s, no );
    if( r2 != SG_OK ) { return r2; }
    // r, m of the owned bodies, gathered through the q staging buffers
    double* rr = s.q; double* mm = s.v;
    for( uint32_t e = 0; e < no; ++e ) { rr[e] = m->r[s.gid[e]]; mm[e] = m->m[s.gid[e]]; }
    sg_ball2d_slab_disconnect( s.ctx );
    if( ( r2 = sg_ball2d_slab_init( s.ctx, no, 0u, m->ghost_cap, rr, mm ) ) != SG_OK ) { return r2; }
    double lim[2];
    sg_slab_limits( uint32_t( m->W ), m->cuts.data(), uint32_t( k ), lim );
    if( ( r2 = sg_ball2d_slab_set_gids( s.ctx, s.gid.data(), lim ) ) != SG_OK ) { return r2; }
    if( ( r2 = sg_ball2d_set_gravity( s.ctx, m->g ) ) != SG_OK ) { return r2; }
    if( ( r2 = sg_ball2d_set_planes( s.ctx, uint32_t( m->plane_x.size() / 2 ), m->plane_x.data(), m->plane_n.data() ) ) != SG_OK ) { return r2; }
    if( ( r2 = sg_ball2d_set_drums( s.ctx, uint32_t( m->drum_r.size() ), m->drum_x.data(), m->drum_r.data() ) ) != SG_OK ) { return r2; }
    return sg_ball2d_slab_mailbox( s.ctx, nullptr, nullptr );
  } );
  if( rc != SG_OK ) { return rc; }
  // neighbours' mailboxes (same process: plain peer access)
  for( int k = 0; k < m->W; ++k )
  {
    for( int side = 0; side < 2; ++side )
    {
      const int peer = ( side == 0 ) ? k - 1 : k + 1;
      if( peer < 0 || peer >= m->W ) { continue; }
      void* mb = nullptr;
      if( ( rc = sg_ball2d_slab_mailbox( m->slab[peer].ctx, &mb, nullptr ) ) != SG_OK ) { return mfail( m, rc, "mailbox of slab %d: %s", peer, sg_last_error( m->slab[peer].ctx ) ); }
      const int pd = ( m->slab[peer].device == m->slab[k].device ) ? -1 : m->slab[peer].device;
      if( ( rc = sg_ball2d_slab_connect( m->slab[k].ctx, side, nullptr, mb, pd ) ) != SG_OK ) { return mfail( m, rc, "slab %d cannot map the mailbox of slab %d: %s", k, peer, sg_last_error( m->slab[k].ctx ) ); }
    }
  }
  m->partitioned = true;
  m->uploads_since_partition = 0;
  ++m->n_partitions;
  return SG_OK;
}

// gathers (q, v) of every slab's bodies into its staging buffers and uploads them
static int stage_and_upload( sg_multi* m, const double* q, const double* v )
{
  const int rc = run_all( m, [m, q, v]( const int k ) -> int
  {
    Slab& s = m->slab[k];
    const size_t no = s.gid.size();
    for( size_t e = 0; e < no; ++e )
    {
      const size_t i = s.gid[e];
      s.q[2 * e] = q[2 * i]; s.q[2 * e + 1] = q[2 * i + 1];
      if( v != nullptr ) { s.v[2 * e] = v[2 * i]; s.v[2 * e + 1] = v[2 * i + 1]; } else { s.v[2 * e] = 0.0; s.v[2 * e + 1] = 0.0; }
    }
    return sg_ball2d_upload( s.ctx, s.q, s.v );
  } );
  if( rc == SG_OK ) { m->state_staged = true; }
  return rc;
}

static int upload_impl( sg_multi* m, const double* q, const double* v, const bool force_partition )
{
  if( m->n == 0u ) { return SG_OK; }
  const bool due = m->rebalance_every != 0u && m->uploads_since_partition >= m->rebalance_every;
  if( !m->partitioned || force_partition || due )
  {
    const int rc = partition_scene( m, q );
    if( rc != SG_OK ) { return rc; }
  }
  ++m->uploads_since_partition;
  m->flow_pending = false;
  m->have_result = false;
  return stage_and_upload( m, q, v );
}

// global (q, v) back from the slabs' staging buffers (the state of the last upload)
static void unstage( sg_multi* m, std::vector<double>& q, std::vector<double>& v )
{
  q.assign( size_t( m->n ) * 2, 0.0 ); v.assign( size_t( m->n ) * 2, 0.0 );
  for( auto& s : m->slab )
  {
    for( size_t e = 0; e < s.gid.size(); ++e )
    {
      const size_t i = s.gid[e];
      q[2 * i] = s.q[2 * e]; q[2 * i + 1] = s.q[2 * e + 1];
      v[2 * i] = s.v[2 * e]; v[2 * i + 1] = s.v[2 * e + 1];
    }
  }
}

// one slab's share of a step: flow (or the prep of an uploaded q1), halo exchange, detection
static int slab_step( sg_multi* m, const int k, const int map_kind, const double dt, const bool do_flow, const bool do_detect )
{
  Slab& s = m->slab[k];
  int rc = SG_OK;
  if( do_flow )
  {
    rc = sg_ball2d_slab_flow( s.ctx, map_kind, dt, nullptr );
    if( m->shared_device ) { if( rc == SG_OK ) { rc = sg_synchronize( s.ctx ); } m->phase.wait(); }
    if( rc != SG_OK ) { if( m->shared_device && do_detect ) { m->phase.wait(); } return rc; }
  }
  if( !do_detect ) { return SG_OK; }
  if( m->shared_device )
  {
    // two slabs on one GPU: a kernel spinning on a neighbour's flag must not keep that neighbour's kernels off the SMs
    rc = sg_ball2d_slab_exchange( s.ctx, 1 );
    if( rc == SG_OK ) { rc = sg_synchronize( s.ctx ); }
    m->phase.wait();
    if( rc != SG_OK ) { return rc; }
    if( ( rc = sg_ball2d_slab_exchange( s.ctx, 2 ) ) != SG_OK ) { return rc; }
  }
  else if( ( rc = sg_ball2d_slab_exchange( s.ctx, 0 ) ) != SG_OK ) { return rc; }
  sg_contacts c;
  rc = sg_ball2d_slab_detect( s.ctx, &c, s.ghosts );
  if( rc != SG_OK ) { return rc; }
  s.n_cand = c.n_candidates; s.n_active = c.n_active; s.n_bb = c.n_body_body;
  return SG_OK;
}

// Runs the step on every slab; a slab asking for a re-partition (or out of ghost slots) gets one: from the staged
// state, and the step is repeated.
static int step_impl( sg_multi* m, const int map_kind, const double dt, const bool flow_already_done )
{
  bool do_flow = !flow_already_done;
  for( int attempt = 0; attempt < 3; ++attempt )
  {
    const int rc = run_all( m, [m, map_kind, dt, do_flow]( const int k ) -> int { return slab_step( m, k, map_kind, dt, do_flow, true ); } );
    if( rc == SG_OK ) { m->have_result = true; m->flow_pending = false; return SG_OK; }
    const bool halo_overflow = rc == SG_ERR_INVALID && m->err.find( "ghost capacity" ) != std::string::npos;
    if( rc != SG_ERR_REBALANCE && !halo_overflow ) { return rc; }
    if( !m->state_staged ) { return mfail( m, SG_ERR_INTERNAL, "re-partition needed but no state is staged" ); }
    if( attempt == 2 || ( rc == SG_ERR_REBALANCE && attempt >= 1 ) ) { return mfail( m, rc == SG_ERR_REBALANCE ? SG_ERR_UNSUPPORTED : rc, "the slabs of this scene are too thin for %d GPUs: bodies reach beyond the neighbouring slab even after a fresh partition", m->W ); }
    if( halo_overflow ) { m->ghost_cap *= 2u; }
    std::vector<double> q, v;
    unstage( m, q, v );
    // an uploaded q1 (active_set without a flow) lives in the q1 staging buffers
    std::vector<double> q1;
    if( map_kind == SG_MAP_NONE )
    {
      q1.assign( size_t( m->n ) * 2, 0.0 );
      for( auto& s : m->slab ) { for( size_t e = 0; e < s.gid.size(); ++e ) { q1[2 * size_t( s.gid[e] )] = s.q1[2 * e]; q1[2 * size_t( s.gid[e] ) + 1] = s.q1[2 * e + 1]; } }
    }
    int r2 = upload_impl( m, q.data(), v.data(), true );
    if( r2 != SG_OK ) { return r2; }
    if( map_kind == SG_MAP_NONE )
    {
      r2 = run_all( m, [m, &q1]( const int k ) -> int
      {
        Slab& s = m->slab[k];
        for( size_t e = 0; e < s.gid.size(); ++e ) { s.q1[2 * e] = q1[2 * size_t( s.gid[e] )]; s.q1[2 * e + 1] = q1[2 * size_t( s.gid[e] ) + 1]; }
        return sg_ball2d_slab_upload_q1( s.ctx, s.q1 );
      } );
      if( r2 != SG_OK ) { return r2; }
    }
    do_flow = true;
  }
  return SG_ERR_INTERNAL;
}

template<typename T>
static void ensure_size( std::vector<T>& v, const size_t n ) { if( v.size() < n ) { v.resize( n + n / 8 + 16 ); } }

extern "C"
{

int sg_create_multi( sg_multi** out, int n_gpus, const int* devices )
{
  if( out == nullptr || n_gpus <= 0 || n_gpus > 64 ) { return mfail( nullptr, SG_ERR_INVALID, "sg_create_multi: bad arguments" ); }
  sg_multi* m = new sg_multi;
  m->W = n_gpus;
  m->slab.resize( size_t( n_gpus ) );
  for( int k = 0; k < n_gpus; ++k )
  {
    m->slab[k].device = ( devices != nullptr ) ? devices[k] : k;
    memset( &m->slab[k].res, 0, sizeof( sg_contacts ) );
    for( int j = 0; j < k; ++j ) { if( m->slab[j].device == m->slab[k].device ) { m->shared_device = true; } }
    const int rc = sg_create( &m->slab[k].ctx, m->slab[k].device );
    if( rc != SG_OK )
    {
      mfail( nullptr, rc, "sg_create_multi: device %d: %s", m->slab[k].device, sg_last_error( nullptr ) );
      for( int j = 0; j < k; ++j ) { sg_destroy( m->slab[j].ctx ); }
      delete m;
      return rc;
    }
  }
  m->phase.n = n_gpus;
  for( int k = 0; k < n_gpus; ++k ) { m->threads.emplace_back( worker_main, m, k ); }
  *out = m;
  return SG_OK;
}

void sg_destroy_multi( sg_multi* m )
{
  if( m == nullptr ) { return; }
  {
    std::lock_guard<std::mutex> lk( m->mu );
    m->quit = true;
  }
  m->cv_go.notify_all();
  for( auto& t : m->threads ) { t.join(); }
  for( auto& s : m->slab )
  {
    if( s.ctx == nullptr ) { continue; }
    sg_ball2d_slab_disconnect( s.ctx );
  }
  for( auto& s : m->slab )
  {
    if( s.ctx == nullptr ) { continue; }
    for( double* p : { s.q, s.v, s.q1, s.v1 } ) { if( p != nullptr ) { sg_host_free( s.ctx, p ); } }
    sg_destroy( s.ctx );
  }
  delete m;
}

const char* sg_multi_last_error( const sg_multi* m ) { return ( m != nullptr ) ? m->err.c_str() : g_multi_create_error.c_str(); }
int sg_multi_n_gpus( const sg_multi* m ) { return ( m != nullptr ) ? m->W : 0; }
sg_ctx* sg_multi_context( sg_multi* m, int k ) { return ( m != nullptr && k >= 0 && k < m->W ) ? m->slab[k].ctx : nullptr; }

int sg_multi_set_rebalance( sg_multi* m, uint32_t every_n_uploads, uint32_t ghost_cap )
{
  if( m == nullptr ) { return SG_ERR_INVALID; }
  m->rebalance_every = every_n_uploads;
  if( ghost_cap != 0u && ghost_cap != m->ghost_cap ) { m->ghost_cap = ghost_cap; m->partitioned = false; }
  return SG_OK;
}

int sg_multi_ball2d_set_bodies( sg_multi* m, uint32_t n, const double* r, const double* mass )
{
  if( m == nullptr ) { return SG_ERR_INVALID; }
  if( n > 0 && ( r == nullptr || mass == nullptr ) ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_set_bodies: null array" ); }
  if( n >= 0x80000000u ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_set_bodies: at most 2^31 - 1 bodies" ); }
  m->n = n;
  m->r.assign( r, r + n ); m->m.assign( mass, mass + n );
  m->partitioned = false; m->state_staged = false; m->flow_pending = false; m->have_result = false;
  m->ghost_cap = 0u;
  return SG_OK;
}

int sg_multi_ball2d_set_gravity( sg_multi* m, const double* g )
{
  if( m == nullptr || g == nullptr ) { return SG_ERR_INVALID; }
  m->g[0] = g[0]; m->g[1] = g[1];
  if( m->partitioned ) { for( auto& s : m->slab ) { sg_ball2d_set_gravity( s.ctx, m->g ); } }
  return SG_OK;
}

int sg_multi_ball2d_set_planes( sg_multi* m, uint32_t n, const double* x, const double* nrm )
{
  if( m == nullptr || ( n > 0 && ( x == nullptr || nrm == nullptr ) ) ) { return SG_ERR_INVALID; }
  m->plane_x.assign( x, x + 2 * size_t( n ) ); m->plane_n.assign( nrm, nrm + 2 * size_t( n ) );
  if( m->partitioned )
  {
    for( auto& s : m->slab ) { const int rc = sg_ball2d_set_planes( s.ctx, n, x, nrm ); if( rc != SG_OK ) { return mfail( m, rc, "%s", sg_last_error( s.ctx ) ); } }
  }
  return SG_OK;
}

int sg_multi_ball2d_set_drums( sg_multi* m, uint32_t n, const double* x, const double* r )
{
  if( m == nullptr || ( n > 0 && ( x == nullptr || r == nullptr ) ) ) { return SG_ERR_INVALID; }
  m->drum_x.assign( x, x + 2 * size_t( n ) ); m->drum_r.assign( r, r + n );
  if( m->partitioned )
  {
    for( auto& s : m->slab ) { const int rc = sg_ball2d_set_drums( s.ctx, n, x, r ); if( rc != SG_OK ) { return mfail( m, rc, "%s", sg_last_error( s.ctx ) ); } }
  }
  return SG_OK;
}

int sg_multi_ball2d_upload( sg_multi* m, const double* q, const double* v )
{
  if( m == nullptr ) { return SG_ERR_INVALID; }
  if( m->n > 0 && ( q == nullptr || v == nullptr ) ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_upload: null vector" ); }
  return upload_impl( m, q, v, false );
}

int sg_multi_ball2d_step( sg_multi* m, int map_kind, double dt, sg_contacts* out )
{
  if( m == nullptr ) { return SG_ERR_INVALID; }
  if( map_kind != SG_MAP_SYMPLECTIC_EULER && map_kind != SG_MAP_VERLET ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_step: map kind %d is not a ball2d map", map_kind ); }
  if( out != nullptr ) { memset( out, 0, sizeof( *out ) ); out->dim = 2; }
  if( m->n == 0u ) { m->have_result = true; return SG_OK; }
  if( !m->state_staged ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_step: no state uploaded" ); }
  m->last_map = map_kind; m->last_dt = dt;
  const int rc = step_impl( m, map_kind, dt, false );
  if( rc != SG_OK ) { return rc; }
  if( out != nullptr ) { for( auto& s : m->slab ) { out->n_candidates += s.n_cand; out->n_active += s.n_active; out->n_body_body += s.n_bb; } }
  return SG_OK;
}

// UnconstrainedMap::flow over host vectors (scisim/UnconstrainedMaps/UnconstrainedMap.h:33): q0, v0 up (partitioned), q1, v1 down
int sg_multi_ball2d_flow( sg_multi* m, int map_kind, const double* q0, const double* v0, double dt, double* q1, double* v1 )
{
  if( m == nullptr ) { return SG_ERR_INVALID; }
  if( map_kind != SG_MAP_SYMPLECTIC_EULER && map_kind != SG_MAP_VERLET ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_flow: map kind %d is not a ball2d map", map_kind ); }
  if( m->n == 0u ) { return SG_OK; }
  if( q0 == nullptr || v0 == nullptr || q1 == nullptr || v1 == nullptr ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_flow: null vector" ); }
  int rc = upload_impl( m, q0, v0, false );
  if( rc != SG_OK ) { return rc; }
  m->last_map = map_kind; m->last_dt = dt;
  rc = run_all( m, [m, map_kind, dt, q1, v1]( const int k ) -> int
  {
    Slab& s = m->slab[k];
    int r2 = slab_step( m, k, map_kind, dt, true, false );
    if( r2 != SG_OK ) { return r2; }
    if( ( r2 = sg_ball2d_fetch_state( s.ctx, s.q1, s.v1 ) ) != SG_OK ) { return r2; }
    for( size_t e = 0; e < s.gid.size(); ++e )
    {
      const size_t i = s.gid[e];
      q1[2 * i] = s.q1[2 * e]; q1[2 * i + 1] = s.q1[2 * e + 1];
      v1[2 * i] = s.v1[2 * e]; v1[2 * i + 1] = s.v1[2 * e + 1];
    }
    return SG_OK;
  } );
  if( rc == SG_OK ) { m->flow_pending = true; }
  return rc;
}

static int multi_merge( sg_multi* m, uint32_t flags, sg_contacts* out );

// ConstrainedSystem::computeActiveSet (scisim/Constraints/ConstrainedSystem.h:20) over host vectors.  With SG_IN_RESIDENT
// (q0, q1) are the vectors of the sg_multi_ball2d_flow just done (ImpactMap.cpp:54-58) and nothing is uploaded.
int sg_multi_ball2d_active_set( sg_multi* m, const double* q0, const double* q1, uint32_t out_flags, sg_contacts* out )
{
  if( m == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  memset( out, 0, sizeof( *out ) ); out->dim = 2;
  if( m->n == 0u ) { return SG_OK; }
  int rc;
  if( ( out_flags & SG_IN_RESIDENT ) != 0u )
  {
    if( !m->flow_pending ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_active_set: SG_IN_RESIDENT without a preceding sg_multi_ball2d_flow" ); }
    rc = step_impl( m, m->last_map, m->last_dt, true );
  }
  else
  {
    if( q0 == nullptr || q1 == nullptr ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_active_set: null vector" ); }
    rc = upload_impl( m, q0, nullptr, false );
    if( rc != SG_OK ) { return rc; }
    rc = run_all( m, [m, q1]( const int k ) -> int
    {
      Slab& s = m->slab[k];
      for( size_t e = 0; e < s.gid.size(); ++e ) { s.q1[2 * e] = q1[2 * size_t( s.gid[e] )]; s.q1[2 * e + 1] = q1[2 * size_t( s.gid[e] ) + 1]; }
      return sg_ball2d_slab_upload_q1( s.ctx, s.q1 );
    } );
    if( rc != SG_OK ) { return rc; }
    rc = step_impl( m, SG_MAP_NONE, 0.0, false );
  }
  if( rc != SG_OK ) { return rc; }
  return multi_merge( m, out_flags, out );
}

int sg_multi_ball2d_fetch( sg_multi* m, uint32_t out_flags, double* q1, double* v1, sg_contacts* out )
{
  if( m == nullptr ) { return SG_ERR_INVALID; }
  if( !m->have_result ) { return mfail( m, SG_ERR_INVALID, "sg_multi_ball2d_fetch: no step has been run" ); }
  if( m->n > 0u && ( q1 != nullptr || v1 != nullptr ) )
  {
    const int rc = run_all( m, [m, q1, v1]( const int k ) -> int
    {
      Slab& s = m->slab[k];
      const int r2 = sg_ball2d_fetch_state( s.ctx, q1 != nullptr ? s.q1 : nullptr, v1 != nullptr ? s.v1 : nullptr );
      if( r2 != SG_OK ) { return r2; }
      for( size_t e = 0; e < s.gid.size(); ++e )
      {
        const size_t i = s.gid[e];
        if( q1 != nullptr ) { q1[2 * i] = s.q1[2 * e]; q1[2 * i + 1] = s.q1[2 * e + 1]; }
        if( v1 != nullptr ) { v1[2 * i] = s.v1[2 * e]; v1[2 * i + 1] = s.v1[2 * e + 1]; }
      }
      return SG_OK;
    } );
    if( rc != SG_OK ) { return rc; }
  }
  if( out == nullptr ) { return SG_OK; }
  memset( out, 0, sizeof( *out ) ); out->dim = 2;
  if( m->n == 0u ) { return SG_OK; }
  return multi_merge( m, out_flags, out );
}

/* how the last step was partitioned: cuts (world + 1 values), bodies per slab, ghosts per slab side (2 per slab), number of partitions made so far */
int sg_multi_partition_info( sg_multi* m, double* cuts, uint32_t* n_owned, uint32_t* ghosts, uint64_t* n_partitions )
{
  if( m == nullptr ) { return SG_ERR_INVALID; }
  if( cuts != nullptr && m->cuts.size() == size_t( m->W ) + 1u ) { memcpy( cuts, m->cuts.data(), m->cuts.size() * 8 ); }
  for( int k = 0; k < m->W; ++k )
  {
    if( n_owned != nullptr ) { n_owned[k] = uint32_t( m->slab[k].gid.size() ); }
    if( ghosts != nullptr ) { ghosts[2 * k] = m->slab[k].ghosts[0]; ghosts[2 * k + 1] = m->slab[k].ghosts[1]; }
  }
  if( n_partitions != nullptr ) { *n_partitions = m->n_partitions; }
  return SG_OK;
}

} // extern "C"

// Every slab copies its lists to pinned memory (in parallel), then the lists are merged into the reference's order.
static int multi_merge( sg_multi* m, const uint32_t flags, sg_contacts* out )
{
  const uint32_t W = uint32_t( m->W );
  const bool want_cand = ( flags & SG_OUT_CANDIDATES ) != 0u;
  int rc = run_all( m, [m, flags]( const int k ) -> int { return sg_ball2d_fetch( m->slab[k].ctx, flags & SG_OUT_ALL, nullptr, nullptr, &m->slab[k].res ); } );
  if( rc != SG_OK ) { return rc; }
  std::vector<const uint32_t*> first( W ), stype( W ), si( W ), sj( W );
  std::vector<uint64_t> len( W ), slen( W );
  std::vector<uint64_t*> dest( W );
  uint64_t n_bb = 0, n_static = 0, n_cand = 0, n_drum = 0, n_plane = 0;
  for( uint32_t k = 0; k < W; ++k )
  {
    Slab& s = m->slab[k];
    n_bb += s.res.n_body_body; n_static += s.res.n_active - s.res.n_body_body; n_cand += s.res.n_candidates; n_drum += s.res.n_drum; n_plane += s.res.n_plane;
    s.dest_bb.resize( s.res.n_body_body ); s.dest_static.resize( s.res.n_active - s.res.n_body_body ); s.dest_cand.resize( want_cand ? s.res.n_candidates : 0 );
  }
  // body-body contacts
  for( uint32_t k = 0; k < W; ++k ) { first[k] = m->slab[k].res.i; len[k] = m->slab[k].res.n_body_body; dest[k] = m->slab[k].dest_bb.data(); }
  if( ( rc = sg_slab_merge_dest( m->n, W, first.data(), 1u, len.data(), dest.data() ) ) != SG_OK ) { return mfail( m, rc, "merge: a slab's contact list is not ascending" ); }
  if( want_cand )
  {
    for( uint32_t k = 0; k < W; ++k ) { first[k] = m->slab[k].res.cand_ij; len[k] = m->slab[k].res.n_candidates; dest[k] = m->slab[k].dest_cand.data(); }
    if( ( rc = sg_slab_merge_dest( m->n, W, first.data(), 2u, len.data(), dest.data() ) ) != SG_OK ) { return mfail( m, rc, "merge: a slab's candidate list is not ascending" ); }
  }
  for( uint32_t k = 0; k < W; ++k )
  {
    const sg_contacts& c = m->slab[k].res;
    stype[k] = c.type + c.n_body_body; si[k] = c.i + c.n_body_body; sj[k] = c.j + c.n_body_body; slen[k] = c.n_active - c.n_body_body; dest[k] = m->slab[k].dest_static.data();
  }
  if( ( rc = sg_slab_merge_static_dest( W, stype.data(), si.data(), sj.data(), slen.data(), dest.data() ) ) != SG_OK ) { return mfail( m, rc, "merge of the static contacts failed" ); }
  const uint64_t na = n_bb + n_static;
  ensure_size( m->o_type, na ); ensure_size( m->o_i, na ); ensure_size( m->o_j, na );
  if( flags & SG_OUT_NORMALS ) { ensure_size( m->o_n, 2 * na ); }
  if( flags & SG_OUT_POINTS ) { ensure_size( m->o_p, 2 * na ); }
  if( flags & SG_OUT_DEPTHS ) { ensure_size( m->o_depth, na ); }
  if( want_cand ) { ensure_size( m->o_cand, 2 * n_cand ); }
  rc = run_all( m, [m, flags, want_cand, n_bb]( const int k ) -> int
  {
    const Slab& s = m->slab[k];
    const sg_contacts& c = s.res;
    const uint64_t nbb = c.n_body_body;
    for( uint64_t e = 0; e < c.n_active; ++e )
    {
      const uint64_t d = ( e < nbb ) ? s.dest_bb[e] : n_bb + s.dest_static[e - nbb];
      m->o_type[d] = c.type[e]; m->o_i[d] = c.i[e]; m->o_j[d] = c.j[e];
      if( flags & SG_OUT_NORMALS ) { m->o_n[2 * d] = c.n[2 * e]; m->o_n[2 * d + 1] = c.n[2 * e + 1]; }
      if( flags & SG_OUT_POINTS ) { m->o_p[2 * d] = c.p[2 * e]; m->o_p[2 * d + 1] = c.p[2 * e + 1]; }
      if( flags & SG_OUT_DEPTHS ) { m->o_depth[d] = c.depth[e]; }
    }
    if( want_cand )
    {
      for( uint64_t e = 0; e < c.n_candidates; ++e ) { const uint64_t d = s.dest_cand[e]; m->o_cand[2 * d] = c.cand_ij[2 * e]; m->o_cand[2 * d + 1] = c.cand_ij[2 * e + 1]; }
    }
    return SG_OK;
  } );
  if( rc != SG_OK ) { return rc; }
  out->dim = 2;
  out->n_candidates = n_cand; out->n_active = na; out->n_body_body = n_bb; out->n_drum = n_drum; out->n_plane = n_plane;
  out->type = m->o_type.data(); out->i = m->o_i.data(); out->j = m->o_j.data();
  out->n = ( flags & SG_OUT_NORMALS ) ? m->o_n.data() : nullptr;
  out->p = ( flags & SG_OUT_POINTS ) ? m->o_p.data() : nullptr;
  out->depth = ( flags & SG_OUT_DEPTHS ) ? m->o_depth.data() : nullptr;
  out->cand_ij = want_cand ? m->o_cand.data() : nullptr;
  return SG_OK;
}
