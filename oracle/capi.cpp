// oracle/capi.cpp
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). Plain-C entry points over the CPU
// restatement so that tests/ (ctypes) and bench.py's cpu_baseline leg can call it. Never linked into
// or loaded by the product library.
#include "ball2d.h"
#include "broadphase.h"
#include "ccd.h"

#include <chrono>
#include <cstring>

using namespace orc;

extern "C"
{

// ---- CCD (scisim/CollisionDetection/CollisionDetectionUtilities.cpp) -------------------------------
void orc_ccd_coeffs( const double* q0a, const double* q1a, double ra, const double* q0b, const double* q1b, double rb, double* c_out )
{
  const CCDCoeffs c = computeCCDQuadraticCoeffs( V2{ q0a[0], q0a[1] }, V2{ q1a[0], q1a[1] }, ra, V2{ q0b[0], q0b[1] }, V2{ q1b[0], q1b[1] }, rb );
  c_out[0] = c.c0; c_out[1] = c.c1; c_out[2] = c.c2;
}

int orc_ccd_happens( const double* c, double* t_out )
{
  const std::pair<bool,double> r = ballBallCCDCollisionHappens( CCDCoeffs{ c[0], c[1], c[2] } );
  *t_out = r.second;
  return r.first ? 1 : 0;
}

// ---- Broad phase (ball2d/SpatialGridDetector.cpp, rigidbody3d/SpatialGridDetector.cpp) -------------
// aabbs: n * 2*dim doubles, [lo(dim), hi(dim)] per box. method: 0 = literal spatial grid, 1 = all pairs.
// Returns a handle holding the ascending (i,j) list.
extern "C++"
{
struct PairResult { std::vector<std::pair<unsigned,unsigned>> pairs; double seconds; };

template<int D>
static PairResult* overlapsImpl( const uint32_t n, const double* aabbs, const int method )
{
  std::vector<Box<D>> boxes( n );
  for( uint32_t b = 0; b < n; ++b )
  {
    for( int k = 0; k < D; ++k ) { boxes[b].lo[k] = aabbs[2 * D * b + k]; boxes[b].hi[k] = aabbs[2 * D * b + D + k]; }
  }
  PairResult* res = new PairResult;
  const auto t0 = std::chrono::steady_clock::now();
  PairSet overlaps;
  if( n > 0 )
  {
    if( method == 0 ) { getPotentialOverlaps<D>( boxes, overlaps ); }
    else { getPotentialOverlapsAllPairs<D>( boxes, overlaps ); }
  }
  const auto t1 = std::chrono::steady_clock::now();
  res->seconds = std::chrono::duration<double>( t1 - t0 ).count();
  res->pairs.assign( overlaps.begin(), overlaps.end() );
  return res;
}
}

void* orc_aabb_overlaps( int dim, uint32_t n, const double* aabbs, int method )
{
  if( dim == 2 ) { return overlapsImpl<2>( n, aabbs, method ); }
  if( dim == 3 ) { return overlapsImpl<3>( n, aabbs, method ); }
  return nullptr;
}
uint64_t orc_pairs_count( const void* h ) { return static_cast<const PairResult*>( h )->pairs.size(); }
double orc_pairs_seconds( const void* h ) { return static_cast<const PairResult*>( h )->seconds; }
void orc_pairs_copy( const void* h, uint32_t* ij_out )
{
  const PairResult* r = static_cast<const PairResult*>( h );
  for( std::size_t k = 0; k < r->pairs.size(); ++k ) { ij_out[2 * k] = r->pairs[k].first; ij_out[2 * k + 1] = r->pairs[k].second; }
}
void orc_pairs_free( void* h ) { delete static_cast<PairResult*>( h ); }

// ---- ball2d ----------------------------------------------------------------------------------------
struct Ball2DHandle
{
  Ball2DScene scene;
  std::vector<Ball2DContact> active;
  std::vector<std::pair<unsigned,unsigned>> candidates;
  double seconds_flow = 0.0;
  double seconds_active = 0.0;
};

// plane_n is normalised here exactly as StaticPlane's constructor does (ball2d/StaticGeometry/StaticPlane.cpp:10-14)
void* orc_ball2d_create( uint32_t n, const double* r, const double* m, const double* g,
                         uint32_t nplanes, const double* plane_x, const double* plane_n,
                         uint32_t ndrums, const double* drum_x, const double* drum_r )
{
  Ball2DHandle* h = new Ball2DHandle;
  h->scene.r.assign( r, r + n );
  h->scene.m.assign( m, m + n );
  h->scene.g[0] = g[0]; h->scene.g[1] = g[1];
  for( uint32_t p = 0; p < nplanes; ++p )
  {
    h->scene.plane_x.push_back( V2{ plane_x[2 * p], plane_x[2 * p + 1] } );
    h->scene.plane_n.push_back( makePlaneNormal( V2{ plane_n[2 * p], plane_n[2 * p + 1] } ) );
  }
  for( uint32_t d = 0; d < ndrums; ++d )
  {
    h->scene.drum_x.push_back( V2{ drum_x[2 * d], drum_x[2 * d + 1] } );
    h->scene.drum_r.push_back( drum_r[d] );
  }
  return h;
}
void orc_ball2d_destroy( void* h ) { delete static_cast<Ball2DHandle*>( h ); }

void orc_ball2d_flow( void* hv, int kind, const double* q0, const double* v0, double dt, double* q1, double* v1 )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  const auto t0 = std::chrono::steady_clock::now();
  flow( kind, h->scene, q0, v0, dt, q1, v1 );
  h->seconds_flow = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
}

// method: 0 = literal spatial grid (reference data structures), 1 = all pairs
void orc_ball2d_active_set( void* hv, const double* q0, const double* q1, int method )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  const auto t0 = std::chrono::steady_clock::now();
  computeActiveSet( h->scene, q0, q1, h->active, &h->candidates, method == 0 );
  h->seconds_active = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
}
uint64_t orc_ball2d_num_candidates( const void* h ) { return static_cast<const Ball2DHandle*>( h )->candidates.size(); }
uint64_t orc_ball2d_num_active( const void* h ) { return static_cast<const Ball2DHandle*>( h )->active.size(); }
double orc_ball2d_seconds_flow( const void* h ) { return static_cast<const Ball2DHandle*>( h )->seconds_flow; }
double orc_ball2d_seconds_active( const void* h ) { return static_cast<const Ball2DHandle*>( h )->seconds_active; }
void orc_ball2d_copy_candidates( const void* hv, uint32_t* ij_out )
{
  const Ball2DHandle* h = static_cast<const Ball2DHandle*>( hv );
  for( std::size_t k = 0; k < h->candidates.size(); ++k ) { ij_out[2 * k] = h->candidates[k].first; ij_out[2 * k + 1] = h->candidates[k].second; }
}
void orc_ball2d_copy_active( const void* hv, uint32_t* type, uint32_t* i, uint32_t* j, double* n, double* p, double* depth )
{
  const Ball2DHandle* h = static_cast<const Ball2DHandle*>( hv );
  for( std::size_t k = 0; k < h->active.size(); ++k )
  {
    const Ball2DContact& c = h->active[k];
    type[k] = c.type; i[k] = c.i; j[k] = c.j;
    n[2 * k] = c.n.x; n[2 * k + 1] = c.n.y;
    p[2 * k] = c.p.x; p[2 * k + 1] = c.p.y;
    depth[k] = c.depth;
  }
}

}
