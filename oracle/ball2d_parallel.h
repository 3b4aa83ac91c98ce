// oracle/ball2d_parallel.h
//
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/oracle_math.h header).  NOT REFERENCE BEHAVIOUR: SCISim's hot path runs on one
// thread even in a USE_OPENMP build, and that is what cpu_baseline / --impl reference time.  This file is the
// optional second CPU figure SURVEY.md 8(d) asks for, so that the GPU speed-up is not flattered by a single-threaded
// opponent: the same ball2d step (flow, swept AABBs, broad phase, ball-ball CCD, drum / plane tests) written the way one
// would write it for a multi-core CPU -- body-parallel over all host threads (std::thread: the image's default g++ has no
// libgomp), a flat uniform grid built by a counting sort instead of std::map / std::set -- with results identical to the
// restatement's (same candidate list in the same order, same active list; tests/test_oracle_parallel.py).  The per-pair
// arithmetic is the restatement's own (ccd.h, ball2d.h).
#ifndef ORACLE_BALL2D_PARALLEL_H
#define ORACLE_BALL2D_PARALLEL_H

#include "ball2d.h"

#include <algorithm>
#include <cstdlib>
#include <thread>

namespace orc
{

struct ParallelStepResult
{
  std::vector<std::pair<unsigned,unsigned>> candidates; // ascending (i,j)
  std::vector<std::pair<unsigned,unsigned>> active;     // ball-ball contacts, ascending (i,j)
  uint64_t n_candidates = 0, n_active = 0;              // list sizes (also set when the lists are not kept)
  uint64_t n_static = 0;                                // drum + plane contacts
  int threads = 1;
};

inline int parallelThreads()
{
  if( const char* e = std::getenv( "ORC_THREADS" ) ) { const int t = std::atoi( e ); if( t > 0 ) { return t; } }
  const unsigned hc = std::thread::hardware_concurrency();
  return hc > 0 ? int( hc ) : 1;
}

// runs fn( t, first, last ) on nthreads contiguous chunks of [0, n)
template<typename F>
inline void parallelChunks( const int nthreads, const int64_t n, F fn )
{
  std::vector<std::thread> pool;
  for( int t = 1; t < nthreads; ++t ) { pool.emplace_back( [=]() { fn( t, n * t / nthreads, n * ( t + 1 ) / nthreads ); } ); }
  fn( 0, int64_t( 0 ), n / nthreads );
  for( std::thread& th : pool ) { th.join(); }
}

// q1, v1: outputs of the map (kind 0 symplectic Euler, 1 Verlet).  keep_lists = false only counts.
inline void parallelStep( const int kind, const Ball2DScene& s, const double* q0, const double* v0, const double dt, double* q1, double* v1, ParallelStepResult& res, const bool keep_lists = true )
{
  const int64_t nb = int64_t( s.r.size() );
  res = ParallelStepResult{};
  const int T = res.threads = parallelThreads();
  if( nb == 0 ) { return; }
  const double inf = std::numeric_limits<double>::infinity();
  std::vector<Box<2>> boxes( nb );
  std::vector<double> tlo0( T, inf ), tlo1( T, inf ), text( T, 0.0 );
  // ---- flow (as ball2d.h: flow) + swept boxes + bounds, body-parallel ----
  parallelChunks( T, nb, [&]( const int t, const int64_t first, const int64_t last )
  {
    double lo0 = inf, lo1 = inf, ext = 0.0;
    for( int64_t b = first; b < last; ++b )
    {
      const double minv = 1.0 / s.m[b];
      for( int k = 0; k < 2; ++k )
      {
        const int64_t d = 2 * b + k;
        const double F = 0.0 + s.m[b] * s.g[k];
        if( kind == 0 )
        {
          v1[d] = v0[d] + ( 0.0 + ( dt * minv ) * F );
          q1[d] = q0[d] + dt * v1[d];
        }
        else
        {
          const double sc = ( 0.5 * dt ) * minv;
          const double vh = v0[d] + ( 0.0 + sc * F );
          q1[d] = q0[d] + dt * vh;
          v1[d] = vh + sc * ( 0.0 + s.m[b] * s.g[k] );
        }
        boxes[b].lo[k] = std::min( q1[d], q0[d] ) - s.r[b];
        boxes[b].hi[k] = std::max( q1[d], q0[d] ) + s.r[b];
      }
      lo0 = std::min( lo0, boxes[b].lo[0] ); lo1 = std::min( lo1, boxes[b].lo[1] );
      ext = std::max( ext, std::max( boxes[b].hi[0] - boxes[b].lo[0], boxes[b].hi[1] - boxes[b].lo[1] ) );
    }
    tlo0[t] = lo0; tlo1[t] = lo1; text[t] = ext;
  } );
  const double lo0 = *std::min_element( tlo0.begin(), tlo0.end() ), lo1 = *std::min_element( tlo1.begin(), tlo1.end() ), ext = *std::max_element( text.begin(), text.end() );
  // ---- uniform grid over lower corners, cell width >= the largest box: overlapping boxes sit at most one cell apart ----
  const double h = ext > 0.0 ? ext * ( 1.0 + 1.0e-6 ) : 1.0;
  std::vector<int64_t> cx( nb ), cy( nb ), tmx( T, 0 ), tmy( T, 0 );
  parallelChunks( T, nb, [&]( const int t, const int64_t first, const int64_t last )
  {
    int64_t mx = 0, my = 0;
    for( int64_t b = first; b < last; ++b )
    {
      cx[b] = int64_t( std::floor( ( boxes[b].lo[0] - lo0 ) / h ) ); cy[b] = int64_t( std::floor( ( boxes[b].lo[1] - lo1 ) / h ) );
      mx = std::max( mx, cx[b] ); my = std::max( my, cy[b] );
    }
    tmx[t] = mx; tmy[t] = my;
  } );
  const int64_t mx = *std::max_element( tmx.begin(), tmx.end() ), my = *std::max_element( tmy.begin(), tmy.end() );
  // cap the table at ~4 cells per body by coarsening (any conservative binning gives the same pair set)
  int64_t shift = 0;
  while( ( ( mx >> shift ) + 1 ) * ( ( my >> shift ) + 1 ) > 4 * nb + 1024 ) { ++shift; }
  const int64_t dx = ( mx >> shift ) + 1, dy = ( my >> shift ) + 1;
  // counting sort by cell (serial: two streaming passes; bodies stay in ascending index order inside a cell)
  std::vector<uint32_t> cell_start( size_t( dx * dy ) + 1, 0u ), order( nb ), cell_of( nb );
  for( int64_t b = 0; b < nb; ++b ) { cell_of[b] = uint32_t( ( cx[b] >> shift ) + dx * ( cy[b] >> shift ) ); ++cell_start[cell_of[b] + 1]; }
  for( size_t c = 0; c < size_t( dx * dy ); ++c ) { cell_start[c + 1] += cell_start[c]; }
  {
    std::vector<uint32_t> fill( cell_start.begin(), cell_start.end() - 1 );
    for( int64_t b = 0; b < nb; ++b ) { order[fill[cell_of[b]]++] = uint32_t( b ); }
  }
  // ---- pairs: body i collects its partners j > i from the 3 x 3 cells around it; each thread owns a contiguous range of i,
  //      so the per-thread lists concatenate to the ascending (i,j) order of the reference's std::set ----
  std::vector<std::vector<std::pair<unsigned,unsigned>>> cand( T ), act( T );
  std::vector<uint64_t> ncand( T, 0 ), nact( T, 0 ), nstat( T, 0 );
  parallelChunks( T, nb, [&]( const int t, const int64_t first, const int64_t last )
  {
    std::vector<unsigned> partners;
    uint64_t nc = 0, na = 0, nst = 0;
    for( int64_t i = first; i < last; ++i )
    {
      partners.clear();
      const int64_t ci = cx[i] >> shift, cj = cy[i] >> shift;
      for( int64_t yy = std::max<int64_t>( cj - 1, 0 ); yy <= std::min( cj + 1, dy - 1 ); ++yy )
      {
        for( int64_t xx = std::max<int64_t>( ci - 1, 0 ); xx <= std::min( ci + 1, dx - 1 ); ++xx )
        {
          const size_t c = size_t( xx + dx * yy );
          for( uint32_t k = cell_start[c]; k < cell_start[c + 1]; ++k )
          {
            const unsigned j = order[k];
            if( int64_t( j ) <= i ) { continue; }
            // AABB::overlaps (ball2d/SpatialGridDetector.cpp): closed intervals
            if( boxes[i].hi[0] < boxes[j].lo[0] || boxes[j].hi[0] < boxes[i].lo[0] || boxes[i].hi[1] < boxes[j].lo[1] || boxes[j].hi[1] < boxes[i].lo[1] ) { continue; }
            partners.push_back( j );
          }
        }
      }
      std::sort( partners.begin(), partners.end() );
      const V2 q0a{ q0[2 * i], q0[2 * i + 1] }, q1a{ q1[2 * i], q1[2 * i + 1] };
      for( const unsigned j : partners )
      {
        ++nc;
        if( keep_lists ) { cand[t].emplace_back( unsigned( i ), j ); }
        const V2 q0b{ q0[2 * j], q0[2 * j + 1] }, q1b{ q1[2 * j], q1[2 * j + 1] };
        if( ballBallCCDCollisionHappens( q0a, q1a, s.r[i], q0b, q1b, s.r[j] ).first )
        {
          ++na;
          if( keep_lists ) { act[t].emplace_back( unsigned( i ), j ); }
        }
      }
      // drums and planes (counts; ball2d.h: computeStaticActiveSet has the contact data)
      for( size_t d = 0; d < s.drum_x.size(); ++d ) { const double Rr = s.drum_r[d] - s.r[i]; if( squaredNorm( s.drum_x[d] - q1a ) >= Rr * Rr ) { ++nst; } }
      for( size_t p = 0; p < s.plane_x.size(); ++p ) { if( dot( s.plane_n[p], q1a - s.plane_x[p] ) <= s.r[i] ) { ++nst; } }
    }
    ncand[t] = nc; nact[t] = na; nstat[t] = nst;
  } );
  for( int t = 0; t < T; ++t ) { res.n_candidates += ncand[t]; res.n_active += nact[t]; res.n_static += nstat[t]; }
  if( keep_lists )
  {
    res.candidates.reserve( res.n_candidates ); res.active.reserve( res.n_active );
    for( int t = 0; t < T; ++t ) { res.candidates.insert( res.candidates.end(), cand[t].begin(), cand[t].end() ); res.active.insert( res.active.end(), act[t].begin(), act[t].end() ); }
  }
}

}

#endif
