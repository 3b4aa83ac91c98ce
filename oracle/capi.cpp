// oracle/capi.cpp
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). Plain-C entry points over the CPU
// restatement so that tests/ (ctypes) and bench.py's cpu_baseline leg can call it. Never linked into
// or loaded by the product library.
#include "ball2d.h"
#include "ball2d_portals.h"
#include "ball2d_parallel.h"
#include "assembly2d.h"
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
  std::vector<Portal2D> portals;
  PortalActiveSetResult pres;
  std::vector<std::pair<unsigned,unsigned>> par_active;
  Assembly2D assembly;
  ConstraintCache2D cache;
  unsigned cache_ncomp = 0;
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

// ---- N, Q, contact bases and the constraint cache for the last active set (oracle/assembly2d.h) ----
// sizes: out[0] = constraints, out[1] = nnz( N ), out[2] = nnz( Q ); returns 0 when the active set holds an unsupported contact type
int orc_ball2d_assemble( void* hv, uint64_t* out )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  assemble2d( h->scene, h->active, h->assembly );
  out[0] = h->active.size(); out[1] = h->assembly.n_inner.size(); out[2] = h->assembly.q_inner.size();
  return h->assembly.supported ? 1 : 0;
}
void orc_ball2d_copy_assembly( const void* hv, int32_t* n_outer, int32_t* n_inner, double* n_val, int32_t* q_outer, int32_t* q_inner, double* q_val, double* bases )
{
  const Assembly2D& a = static_cast<const Ball2DHandle*>( hv )->assembly;
  std::copy( a.n_outer.begin(), a.n_outer.end(), n_outer ); std::copy( a.n_inner.begin(), a.n_inner.end(), n_inner ); std::copy( a.n_val.begin(), a.n_val.end(), n_val );
  std::copy( a.q_outer.begin(), a.q_outer.end(), q_outer ); std::copy( a.q_inner.begin(), a.q_inner.end(), q_inner ); std::copy( a.q_val.begin(), a.q_val.end(), q_val );
  std::copy( a.bases.begin(), a.bases.end(), bases );
}
void orc_ball2d_cache_clear( void* hv ) { static_cast<Ball2DHandle*>( hv )->cache.clear(); }
// cacheConstraint for every constraint of the last active set
void orc_ball2d_cache_store( void* hv, uint32_t ncomp, const double* r )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  h->cache.clear(); // the maps clear the cache before re-filling it (ImpactMap.cpp:98)
  h->cache_ncomp = ncomp;
  for( std::size_t c = 0; c < h->active.size(); ++c ) { h->cache.cache( h->active[c], r + c * ncomp, ncomp ); }
}
// getCachedConstraintImpulse for every constraint of the last active set; returns the number found
uint64_t orc_ball2d_cache_lookup( void* hv, uint32_t ncomp, double* r_out )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  uint64_t hits = 0;
  for( std::size_t c = 0; c < h->active.size(); ++c ) { if( h->cache.get( h->active[c], r_out + c * ncomp, ncomp ) ) { ++hits; } }
  return hits;
}

// ---- multi-core ball2d step (oracle/ball2d_parallel.h): NOT reference behaviour, the optional second CPU figure ----
// returns seconds; counts in out[0..2] = candidates, ball-ball contacts, static contacts; out[3] = threads used.
// keep_lists != 0 also leaves the two pair lists for orc_ball2d_parallel_copy.
double orc_ball2d_parallel_step( void* hv, int kind, const double* q0, const double* v0, double dt, double* q1, double* v1, int keep_lists, uint64_t* out )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  static ParallelStepResult res;
  const auto t0 = std::chrono::steady_clock::now();
  parallelStep( kind, h->scene, q0, v0, dt, q1, v1, res, keep_lists != 0 );
  const double secs = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  out[0] = res.n_candidates; out[1] = res.n_active; out[2] = res.n_static; out[3] = uint64_t( res.threads );
  h->candidates = res.candidates;
  h->par_active = res.active;
  return secs;
}
void orc_ball2d_parallel_copy( const void* hv, uint32_t* cand_ij, uint32_t* active_ij )
{
  const Ball2DHandle* h = static_cast<const Ball2DHandle*>( hv );
  for( std::size_t k = 0; k < h->candidates.size(); ++k ) { cand_ij[2 * k] = h->candidates[k].first; cand_ij[2 * k + 1] = h->candidates[k].second; }
  for( std::size_t k = 0; k < h->par_active.size(); ++k ) { active_ij[2 * k] = h->par_active[k].first; active_ij[2 * k + 1] = h->par_active[k].second; }
}

// ---- ball2d portals (oracle/ball2d_portals.h) ----------------------------------------------------------
// plane normals are normalised as StaticPlane's constructor does; dx starts at 0 (PlanarPortal.cpp:61-69)
void orc_ball2d_set_portals( void* hv, uint32_t n, const double* ax, const double* an, const double* bx, const double* bn, const double* v, const double* bounds )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  h->portals.clear();
  for( uint32_t p = 0; p < n; ++p )
  {
    Portal2D pt;
    pt.a = makePlane2D( V2{ ax[2 * p], ax[2 * p + 1] }, V2{ an[2 * p], an[2 * p + 1] } );
    pt.b = makePlane2D( V2{ bx[2 * p], bx[2 * p + 1] }, V2{ bn[2 * p], bn[2 * p + 1] } );
    pt.v = v[p]; pt.bounds = bounds[p]; pt.dx = 0.0;
    h->portals.push_back( pt );
  }
}
// Ball2DSim::updatePeriodicBoundaryConditionsStartOfStep with t = next_iteration * dt (ball2d/Ball2DSim.cpp:327-334)
void orc_ball2d_update_portals( void* hv, double t, double* dx_out )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  for( std::size_t p = 0; p < h->portals.size(); ++p ) { updateMovingPortals( h->portals[p], t ); if( dx_out != nullptr ) { dx_out[p] = h->portals[p].dx; } }
}
void orc_ball2d_enforce_portals( void* hv, double* q, double* v )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  enforcePeriodicBoundaryConditions( h->portals, uint32_t( h->scene.r.size() ), q, v );
}
// one portal's primitives at one point, for the bit-for-bit check against the reference's own PlanarPortal.cpp:
// out[0..1] through plane A, [2..3] through plane B, [4..5] teleportBall, [6..7] teleportPointInsidePortal,
// [8..9] getKinematicVelocityOfBall, [10..11] getKinematicVelocityOfPoint; flags: bit 0 ballTouchesPortal, bit 1 its plane index,
// bit 2 touches both, bit 3 pointInsidePortal
uint32_t orc_ball2d_portal_probe( const void* hv, uint32_t p, const double* x, double r, double* out )
{
  const Ball2DHandle* h = static_cast<const Ball2DHandle*>( hv );
  const Portal2D& pt = h->portals[p];
  const V2 xin{ x[0], x[1] };
  const V2 a = teleportPointThroughPlaneA( pt, xin ), b = teleportPointThroughPlaneB( pt, xin );
  const V2 tb = teleportBall( pt, xin, r ), ti = teleportPointInsidePortal( pt, xin );
  const V2 kb = getKinematicVelocityOfBall( pt, xin, r ), kp = getKinematicVelocityOfPoint( pt, xin );
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y; out[4] = tb.x; out[5] = tb.y; out[6] = ti.x; out[7] = ti.y;
  out[8] = kb.x; out[9] = kb.y; out[10] = kp.x; out[11] = kp.y;
  const int touch = ballTouchesPortal( pt, xin, r );
  uint32_t flags = 0u;
  if( touch != 0 ) { flags |= 1u; }
  if( touch == 2 ) { flags |= 2u; }
  if( touch < 0 ) { flags |= 4u; }
  if( pointInsidePortal( pt, xin ) ) { flags |= 8u; }
  return flags;
}
// returns 0, or -1 where the reference exits (a ball touching both planes of one portal)
int orc_ball2d_active_set_portals( void* hv, const double* q0, const double* q1, int method )
{
  Ball2DHandle* h = static_cast<Ball2DHandle*>( hv );
  const auto t0 = std::chrono::steady_clock::now();
  computeActiveSetWithPortals( h->scene, h->portals, q0, q1, h->pres, method == 0 );
  h->seconds_active = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  h->active = h->pres.active;
  h->candidates = h->pres.candidates;
  return h->pres.both_planes_touched ? -1 : 0;
}
uint64_t orc_ball2d_portals_num_regular( const void* h ) { return static_cast<const Ball2DHandle*>( h )->pres.n_regular; }
uint64_t orc_ball2d_portals_num_boxes( const void* h ) { return static_cast<const Ball2DHandle*>( h )->pres.teleported_boxes.size(); }
uint64_t orc_ball2d_portals_num_teleported( const void* h ) { return static_cast<const Ball2DHandle*>( h )->pres.teleported_info.size(); }
// box_portal: portal index | plane index << 31
void orc_ball2d_portals_copy_boxes( const void* hv, uint32_t* box_body, uint32_t* box_portal )
{
  const Ball2DHandle* h = static_cast<const Ball2DHandle*>( hv );
  for( std::size_t k = 0; k < h->pres.teleported_boxes.size(); ++k )
  {
    const TeleportedBall2D& tb = h->pres.teleported_boxes[k];
    box_body[k] = tb.body; box_portal[k] = tb.portal | ( tb.plane ? 0x80000000u : 0u );
  }
}
// portal0/1: portal index | plane index << 31, 0xffffffff when that body was not teleported; x0, x1, kick: 2 doubles each
void orc_ball2d_portals_copy_teleported( const void* hv, uint32_t* portal0, uint32_t* portal1, double* x0, double* x1, double* kick )
{
  const Ball2DHandle* h = static_cast<const Ball2DHandle*>( hv );
  for( std::size_t k = 0; k < h->pres.teleported_info.size(); ++k )
  {
    const TeleportedContactInfo& t = h->pres.teleported_info[k];
    portal0[k] = t.p0 == NO_PORTAL ? NO_PORTAL : ( t.p0 | ( t.pl0 ? 0x80000000u : 0u ) );
    portal1[k] = t.p1 == NO_PORTAL ? NO_PORTAL : ( t.p1 | ( t.pl1 ? 0x80000000u : 0u ) );
    x0[2 * k] = t.x0.x; x0[2 * k + 1] = t.x0.y; x1[2 * k] = t.x1.x; x1[2 * k + 1] = t.x1.y; kick[2 * k] = t.kick.x; kick[2 * k + 1] = t.kick.y;
  }
}

}

// ---- rigidbody3d -----------------------------------------------------------------------------------
#include "rb3d.h"
#include "rb3d_portals.h"

struct RB3DHandle
{
  RB3DScene scene;
  std::vector<RB3DContact> active;
  std::vector<std::pair<unsigned,unsigned>> candidates;
  std::vector<Portal3D> portals;
  RB3DPortalResult pres;
  bool supported = true;
  double seconds_flow = 0.0;
  double seconds_active = 0.0;
};

extern "C"
{

// geo_type/geo_r/geo_half/geo_mesh describe the geometry list (as m_geometry in RigidBody3DState); geo_of_body indexes it.
// plane normals are normalised here exactly as rigidbody3d/StaticGeometry/StaticPlane.cpp:10-15.
void* orc_rb3d_create( uint32_t n, const uint32_t* geo_of_body, const uint8_t* fixed, const double* m, const double* I0, const double* g,
                       uint32_t ngeo, const uint32_t* geo_type, const double* geo_r, const double* geo_half, const uint32_t* geo_mesh,
                       uint32_t nplanes, const double* plane_x, const double* plane_n )
{
  RB3DHandle* h = new RB3DHandle;
  RB3DScene& s = h->scene;
  s.geo_of_body.assign( geo_of_body, geo_of_body + n );
  s.fixed.assign( fixed, fixed + n );
  s.m.assign( m, m + n );
  for( uint32_t b = 0; b < n; ++b ) { s.I0.push_back( V3{ I0[3 * b], I0[3 * b + 1], I0[3 * b + 2] } ); }
  s.g = V3{ g[0], g[1], g[2] };
  for( uint32_t k = 0; k < ngeo; ++k )
  {
    RB3DGeometry geo;
    geo.type = geo_type[k]; geo.r = geo_r[k]; geo.half = V3{ geo_half[3 * k], geo_half[3 * k + 1], geo_half[3 * k + 2] }; geo.mesh = geo_mesh[k];
    s.geometry.push_back( geo );
  }
  for( uint32_t p = 0; p < nplanes; ++p )
  {
    s.plane_x.push_back( V3{ plane_x[3 * p], plane_x[3 * p + 1], plane_x[3 * p + 2] } );
    s.plane_n.push_back( normalized( V3{ plane_n[3 * p], plane_n[3 * p + 1], plane_n[3 * p + 2] } ) );
  }
  return h;
}
// axes are normalised here exactly as rigidbody3d/StaticGeometry/StaticCylinder.cpp:8-19
void orc_rb3d_set_cylinders( void* hv, uint32_t n, const double* x, const double* axis, const double* r )
{
  RB3DScene& s = static_cast<RB3DHandle*>( hv )->scene;
  s.cyl_x.clear(); s.cyl_axis.clear(); s.cyl_r.clear();
  for( uint32_t c = 0; c < n; ++c )
  {
    s.cyl_x.push_back( V3{ x[3 * c], x[3 * c + 1], x[3 * c + 2] } );
    s.cyl_axis.push_back( normalized( V3{ axis[3 * c], axis[3 * c + 1], axis[3 * c + 2] } ) );
    s.cyl_r.push_back( r[c] );
  }
}
void orc_rb3d_destroy( void* h ) { delete static_cast<RB3DHandle*>( h ); }

uint32_t orc_rb3d_add_mesh( void* hv, uint32_t nverts, const double* verts, uint32_t nsamples, const double* samples, uint32_t nhull, const double* hull,
                            const double* cell_delta, const uint32_t* dims, const double* origin, const double* sdf )
{
  RB3DHandle* h = static_cast<RB3DHandle*>( hv );
  RB3DMesh mesh;
  for( uint32_t k = 0; k < nverts; ++k ) { mesh.verts.push_back( V3{ verts[3 * k], verts[3 * k + 1], verts[3 * k + 2] } ); }
  for( uint32_t k = 0; k < nsamples; ++k ) { mesh.samples.push_back( V3{ samples[3 * k], samples[3 * k + 1], samples[3 * k + 2] } ); }
  for( uint32_t k = 0; k < nhull; ++k ) { mesh.hull.push_back( V3{ hull[3 * k], hull[3 * k + 1], hull[3 * k + 2] } ); }
  mesh.cell_delta = V3{ cell_delta[0], cell_delta[1], cell_delta[2] };
  mesh.dims[0] = dims[0]; mesh.dims[1] = dims[1]; mesh.dims[2] = dims[2];
  mesh.origin = V3{ origin[0], origin[1], origin[2] };
  // RigidBodyTriangleMesh.cpp:102
  mesh.grid_end = V3{ origin[0] + double( dims[0] - 1 ) * cell_delta[0], origin[1] + double( dims[1] - 1 ) * cell_delta[1], origin[2] + double( dims[2] - 1 ) * cell_delta[2] };
  mesh.sdf.assign( sdf, sdf + std::size_t( dims[0] ) * dims[1] * dims[2] );
  h->scene.meshes.push_back( mesh );
  return uint32_t( h->scene.meshes.size() - 1 );
}

void orc_rb3d_flow( void* hv, int kind, const double* q0, const double* v0, double dt, double* q1, double* v1 )
{
  RB3DHandle* h = static_cast<RB3DHandle*>( hv );
  const auto t0 = std::chrono::steady_clock::now();
  flow( kind, h->scene, q0, v0, dt, q1, v1 );
  h->seconds_flow = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
}

// the flow of every step after the first: M as left by updateMandMinv ( see flow() )
void orc_rb3d_flow_m_updated( void* hv, int kind, const double* q0, const double* v0, double dt, double* q1, double* v1 )
{
  RB3DHandle* h = static_cast<RB3DHandle*>( hv );
  flow( kind, h->scene, q0, v0, dt, q1, v1, true );
}

void orc_rb3d_update_m_minv( void* hv, const double* q, double* I_blocks, double* Iinv_blocks )
{
  updateMandMinv( static_cast<RB3DHandle*>( hv )->scene, q, I_blocks, Iinv_blocks );
}

// returns 1, or 0 when the reference would exit on an unsupported geometry pair
int orc_rb3d_active_set( void* hv, const double* q0, const double* q1, int method )
{
  RB3DHandle* h = static_cast<RB3DHandle*>( hv );
  const auto t0 = std::chrono::steady_clock::now();
  h->supported = computeActiveSet( h->scene, q0, q1, h->active, &h->candidates, method == 0 );
  h->seconds_active = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  return h->supported ? 1 : 0;
}
uint64_t orc_rb3d_num_candidates( const void* h ) { return static_cast<const RB3DHandle*>( h )->candidates.size(); }
uint64_t orc_rb3d_num_active( const void* h ) { return static_cast<const RB3DHandle*>( h )->active.size(); }
double orc_rb3d_seconds_flow( const void* h ) { return static_cast<const RB3DHandle*>( h )->seconds_flow; }
double orc_rb3d_seconds_active( const void* h ) { return static_cast<const RB3DHandle*>( h )->seconds_active; }
void orc_rb3d_copy_candidates( const void* hv, uint32_t* ij_out )
{
  const RB3DHandle* h = static_cast<const RB3DHandle*>( hv );
  for( std::size_t k = 0; k < h->candidates.size(); ++k ) { ij_out[2 * k] = h->candidates[k].first; ij_out[2 * k + 1] = h->candidates[k].second; }
}
void orc_rb3d_copy_active( const void* hv, uint32_t* type, uint32_t* i, uint32_t* j, uint32_t* aux, double* n, double* p, double* depth )
{
  const RB3DHandle* h = static_cast<const RB3DHandle*>( hv );
  for( std::size_t k = 0; k < h->active.size(); ++k )
  {
    const RB3DContact& c = h->active[k];
    type[k] = c.type; i[k] = c.i; j[k] = c.j; aux[k] = c.aux;
    n[3 * k] = c.n.x; n[3 * k + 1] = c.n.y; n[3 * k + 2] = c.n.z;
    p[3 * k] = c.p.x; p[3 * k + 1] = c.p.y; p[3 * k + 2] = c.p.z;
    depth[k] = c.depth;
  }
}

// single-mesh routines by themselves (checked against the reference's own RigidBodyTriangleMesh.cpp in tests/test_oracle_vs_reference.py)
int orc_rb3d_mesh_detect( const void* hv, uint32_t mesh, const double* x, double* n_out )
{
  const RB3DHandle* h = static_cast<const RB3DHandle*>( hv );
  V3 n{ 0.0, 0.0, 0.0 };
  const bool hit = h->scene.meshes[mesh].detectCollision( V3{ x[0], x[1], x[2] }, n );
  n_out[0] = n.x; n_out[1] = n.y; n_out[2] = n.z;
  return hit ? 1 : 0;
}
// RigidBodyTriangleMesh::computeAABB for body b of the scene placed at ( cm, R ); out = min(3), max(3)
void orc_rb3d_body_aabb( const void* hv, uint32_t b, const double* cm, const double* R, double* out )
{
  const RB3DHandle* h = static_cast<const RB3DHandle*>( hv );
  M3 Rm;
  for( int k = 0; k < 9; ++k ) { Rm.m[k] = R[k]; }
  Box<3> bx;
  computeAABB( h->scene, b, V3{ cm[0], cm[1], cm[2] }, Rm, bx );
  for( int k = 0; k < 3; ++k ) { out[k] = bx.lo[k]; out[3 + k] = bx.hi[k]; }
}

// ---- rigidbody3d portals (oracle/rb3d_portals.h) ---------------------------------------------------------
// StaticPlane( x, n ) frame: out = n (3), t0 (3), t1 (3)
void orc_rb3d_plane_frame( const double* x, const double* n, double* out )
{
  const Plane3D p = makePlane3D( V3{ x[0], x[1], x[2] }, V3{ n[0], n[1], n[2] } );
  out[0] = p.n.x; out[1] = p.n.y; out[2] = p.n.z; out[3] = p.t0.x; out[4] = p.t0.y; out[5] = p.t0.z; out[6] = p.t1.x; out[7] = p.t1.y; out[8] = p.t1.z;
}
void orc_rb3d_set_portals( void* hv, uint32_t n, const double* ax, const double* an, const double* bx, const double* bn, const int* mult )
{
  RB3DHandle* h = static_cast<RB3DHandle*>( hv );
  h->portals.clear();
  for( uint32_t p = 0; p < n; ++p )
  {
    Portal3D pt;
    pt.a = makePlane3D( V3{ ax[3 * p], ax[3 * p + 1], ax[3 * p + 2] }, V3{ an[3 * p], an[3 * p + 1], an[3 * p + 2] } );
    pt.b = makePlane3D( V3{ bx[3 * p], bx[3 * p + 1], bx[3 * p + 2] }, V3{ bn[3 * p], bn[3 * p + 1], bn[3 * p + 2] } );
    for( int k = 0; k < 3; ++k ) { pt.mult[k] = mult[3 * p + k]; }
    h->portals.push_back( pt );
  }
}
// layout of ref_rb3d_portal_probe (oracle/ref_shims/ref_rb3d.cpp)
uint32_t orc_rb3d_portal_probe( const void* hv, uint32_t p, const double* box, const double* x, double* out )
{
  const RB3DHandle* h = static_cast<const RB3DHandle*>( hv );
  const Portal3D& pt = h->portals[p];
  const V3 xin{ x[0], x[1], x[2] };
  const V3 a = teleportPointThroughPlaneA( pt, xin ), b = teleportPointThroughPlaneB( pt, xin ), ti = teleportPointInsidePortal( pt, xin );
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = b.x; out[4] = b.y; out[5] = b.z; out[6] = ti.x; out[7] = ti.y; out[8] = ti.z;
  Box<3> bb;
  for( int k = 0; k < 3; ++k ) { bb.lo[k] = box[k]; bb.hi[k] = box[3 + k]; }
  return uint32_t( aabbTouchesPortal( pt, bb ) ) | ( pointInsidePortal( pt, xin ) ? 4u : 0u );
}
void orc_rb3d_enforce_portals( void* hv, double* q )
{
  RB3DHandle* h = static_cast<RB3DHandle*>( hv );
  enforcePeriodicBoundaryConditionsRB3D( h->portals, uint32_t( h->scene.nbodies() ), q );
}
int orc_rb3d_active_set_portals( void* hv, const double* q0, const double* q1, int method )
{
  RB3DHandle* h = static_cast<RB3DHandle*>( hv );
  computeActiveSetWithPortalsRB3D( h->scene, h->portals, q0, q1, h->pres, method == 0 );
  h->active = h->pres.active;
  h->candidates = h->pres.candidates;
  h->supported = h->pres.supported;
  return h->pres.supported ? 1 : 0;
}
uint64_t orc_rb3d_portals_num_regular( const void* h ) { return static_cast<const RB3DHandle*>( h )->pres.n_regular; }
uint64_t orc_rb3d_portals_num_boxes( const void* h ) { return static_cast<const RB3DHandle*>( h )->pres.teleported_boxes.size(); }
uint64_t orc_rb3d_portals_num_teleported( const void* h ) { return static_cast<const RB3DHandle*>( h )->pres.teleported_info.size(); }
void orc_rb3d_portals_copy_boxes( const void* hv, uint32_t* box_body, uint32_t* box_portal )
{
  const RB3DHandle* h = static_cast<const RB3DHandle*>( hv );
  for( std::size_t k = 0; k < h->pres.teleported_boxes.size(); ++k )
  {
    const TeleportedBall2D& tb = h->pres.teleported_boxes[k];
    box_body[k] = tb.body; box_portal[k] = tb.portal | ( tb.plane ? 0x80000000u : 0u );
  }
}
void orc_rb3d_portals_copy_teleported( const void* hv, uint32_t* portal0, uint32_t* portal1, double* x0, double* x1 )
{
  const RB3DHandle* h = static_cast<const RB3DHandle*>( hv );
  for( std::size_t k = 0; k < h->pres.teleported_info.size(); ++k )
  {
    const RB3DTeleportedInfo& t = h->pres.teleported_info[k];
    portal0[k] = t.p0 == NO_PORTAL ? NO_PORTAL : ( t.p0 | ( t.pl0 ? 0x80000000u : 0u ) );
    portal1[k] = t.p1 == NO_PORTAL ? NO_PORTAL : ( t.p1 | ( t.pl1 ? 0x80000000u : 0u ) );
    x0[3 * k] = t.x0.x; x0[3 * k + 1] = t.x0.y; x0[3 * k + 2] = t.x0.z; x1[3 * k] = t.x1.x; x1[3 * k + 1] = t.x1.y; x1[3 * k + 2] = t.x1.z;
  }
}

}

// ---- rigidbody2d -----------------------------------------------------------------------------------
#include "rb2d.h"
#include "rb2d_portals.h"

struct RB2DHandle
{
  RB2DScene scene;
  std::vector<RB2DContact> active;
  std::vector<std::pair<unsigned,unsigned>> candidates;
  double seconds_flow = 0.0, seconds_active = 0.0;
  std::vector<Portal2D> portals;
  RB2DPortalResult pres;
};

extern "C"
{

void* orc_rb2d_create( uint32_t n, const uint32_t* geo_of_body, const uint8_t* fixed, const double* M, const double* g,
                       uint32_t ngeo, const uint32_t* geo_type, const double* geo_r, const double* geo_half,
                       uint32_t nplanes, const double* plane_x, const double* plane_n )
{
  RB2DHandle* h = new RB2DHandle;
  RB2DScene& s = h->scene;
  s.geo_of_body.assign( geo_of_body, geo_of_body + n );
  s.fixed.assign( fixed, fixed + n );
  s.M.assign( M, M + 3 * std::size_t( n ) );
  s.g = V2{ g[0], g[1] };
  for( uint32_t k = 0; k < ngeo; ++k ) { s.geometry.push_back( RB2DGeometry{ geo_type[k], geo_r[k], V2{ geo_half[2 * k], geo_half[2 * k + 1] } } ); }
  for( uint32_t p = 0; p < nplanes; ++p )
  {
    s.plane_x.push_back( V2{ plane_x[2 * p], plane_x[2 * p + 1] } );
    s.plane_n.push_back( V2{ plane_n[2 * p], plane_n[2 * p + 1] } );
  }
  return h;
}
void orc_rb2d_destroy( void* h ) { delete static_cast<RB2DHandle*>( h ); }
void orc_rb2d_flow( void* hv, int kind, const double* q0, const double* v0, double dt, double* q1, double* v1 )
{
  RB2DHandle* h = static_cast<RB2DHandle*>( hv );
  const auto t0 = std::chrono::steady_clock::now();
  flow( kind, h->scene, q0, v0, dt, q1, v1 );
  h->seconds_flow = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
}
int orc_rb2d_active_set( void* hv, const double* q0, const double* q1, int method )
{
  RB2DHandle* h = static_cast<RB2DHandle*>( hv );
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = computeActiveSet( h->scene, q0, q1, h->active, &h->candidates, method == 0 );
  h->seconds_active = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  return ok ? 1 : 0;
}
uint64_t orc_rb2d_num_candidates( const void* h ) { return static_cast<const RB2DHandle*>( h )->candidates.size(); }
uint64_t orc_rb2d_num_active( const void* h ) { return static_cast<const RB2DHandle*>( h )->active.size(); }
void orc_rb2d_copy_candidates( const void* hv, uint32_t* ij_out )
{
  const RB2DHandle* h = static_cast<const RB2DHandle*>( hv );
  for( std::size_t k = 0; k < h->candidates.size(); ++k ) { ij_out[2 * k] = h->candidates[k].first; ij_out[2 * k + 1] = h->candidates[k].second; }
}
void orc_rb2d_copy_active( const void* hv, uint32_t* type, uint32_t* i, uint32_t* j, uint32_t* aux, double* n, double* p, double* depth )
{
  const RB2DHandle* h = static_cast<const RB2DHandle*>( hv );
  for( std::size_t k = 0; k < h->active.size(); ++k )
  {
    const RB2DContact& c = h->active[k];
    type[k] = c.type; i[k] = c.i; j[k] = c.j; aux[k] = c.aux;
    n[2 * k] = c.n.x; n[2 * k + 1] = c.n.y; p[2 * k] = c.p.x; p[2 * k + 1] = c.p.y; depth[k] = c.depth;
  }
}

// ---- rigidbody2d portals (oracle/rb2d_portals.h) ---------------------------------------------------------
// plane normals are used as given (RigidBody2DStaticPlane does not normalise)
void orc_rb2d_set_portals( void* hv, uint32_t n, const double* ax, const double* an, const double* bx, const double* bn, const double* v, const double* bounds )
{
  RB2DHandle* h = static_cast<RB2DHandle*>( hv );
  h->portals.clear();
  for( uint32_t p = 0; p < n; ++p )
  {
    Portal2D pt;
    pt.a = makePlaneRB2D( V2{ ax[2 * p], ax[2 * p + 1] }, V2{ an[2 * p], an[2 * p + 1] } );
    pt.b = makePlaneRB2D( V2{ bx[2 * p], bx[2 * p + 1] }, V2{ bn[2 * p], bn[2 * p + 1] } );
    pt.v = v[p]; pt.bounds = bounds[p]; pt.dx = 0.0;
    h->portals.push_back( pt );
  }
}
void orc_rb2d_update_portals( void* hv, double t, double* dx_out )
{
  RB2DHandle* h = static_cast<RB2DHandle*>( hv );
  for( std::size_t p = 0; p < h->portals.size(); ++p ) { updateMovingPortals( h->portals[p], t ); if( dx_out != nullptr ) { dx_out[p] = h->portals[p].dx; } }
}
void orc_rb2d_enforce_portals( void* hv, double* q, double* v )
{
  RB2DHandle* h = static_cast<RB2DHandle*>( hv );
  enforcePeriodicBoundaryConditionsRB2D( h->portals, uint32_t( h->scene.nbodies() ), q, v );
}
// box = [minx, miny, maxx, maxy]; out[0..1] teleportPoint through the touched plane ( A when nothing is touched ), [2..3]
// getKinematicVelocityOfAABB; returns aabbTouchesPortal: 0 = no, 1 = plane A, 2 = plane B
uint32_t orc_rb2d_portal_probe( const void* hv, uint32_t p, const double* box, const double* x, double* out )
{
  const RB2DHandle* h = static_cast<const RB2DHandle*>( hv );
  const Portal2D& pt = h->portals[p];
  Box<2> b;
  b.lo[0] = box[0]; b.lo[1] = box[1]; b.hi[0] = box[2]; b.hi[1] = box[3];
  const int touch = aabbTouchesPortal( pt, b );
  const V2 xin{ x[0], x[1] };
  const V2 xo = touch == 2 ? teleportPointThroughPlaneB( pt, xin ) : teleportPointThroughPlaneA( pt, xin );
  const V2 k = getKinematicVelocityOfAABB( pt, b );
  out[0] = xo.x; out[1] = xo.y; out[2] = k.x; out[3] = k.y;
  return uint32_t( touch );
}
// returns 1, or 0 where the reference exits (boxes or kinematic bodies in a teleported collision, unsupported regular pairs)
int orc_rb2d_active_set_portals( void* hv, const double* q0, const double* q1, int method )
{
  RB2DHandle* h = static_cast<RB2DHandle*>( hv );
  computeActiveSetWithPortalsRB2D( h->scene, h->portals, q0, q1, h->pres, method == 0 );
  h->active = h->pres.active;
  h->candidates = h->pres.candidates;
  return h->pres.supported ? 1 : 0;
}
uint64_t orc_rb2d_portals_num_regular( const void* h ) { return static_cast<const RB2DHandle*>( h )->pres.n_regular; }
uint64_t orc_rb2d_portals_num_boxes( const void* h ) { return static_cast<const RB2DHandle*>( h )->pres.teleported_boxes.size(); }
uint64_t orc_rb2d_portals_num_teleported( const void* h ) { return static_cast<const RB2DHandle*>( h )->pres.teleported_info.size(); }
void orc_rb2d_portals_copy_boxes( const void* hv, uint32_t* box_body, uint32_t* box_portal )
{
  const RB2DHandle* h = static_cast<const RB2DHandle*>( hv );
  for( std::size_t k = 0; k < h->pres.teleported_boxes.size(); ++k )
  {
    const TeleportedBall2D& tb = h->pres.teleported_boxes[k];
    box_body[k] = tb.body; box_portal[k] = tb.portal | ( tb.plane ? 0x80000000u : 0u );
  }
}
void orc_rb2d_portals_copy_teleported( const void* hv, uint32_t* portal0, uint32_t* portal1, double* x0, double* x1, double* delta0, double* delta1, double* kick )
{
  const RB2DHandle* h = static_cast<const RB2DHandle*>( hv );
  for( std::size_t k = 0; k < h->pres.teleported_info.size(); ++k )
  {
    const RB2DTeleportedInfo& t = h->pres.teleported_info[k];
    portal0[k] = t.p0 == NO_PORTAL ? NO_PORTAL : ( t.p0 | ( t.pl0 ? 0x80000000u : 0u ) );
    portal1[k] = t.p1 == NO_PORTAL ? NO_PORTAL : ( t.p1 | ( t.pl1 ? 0x80000000u : 0u ) );
    x0[2 * k] = t.x0.x; x0[2 * k + 1] = t.x0.y; x1[2 * k] = t.x1.x; x1[2 * k + 1] = t.x1.y;
    delta0[2 * k] = t.delta0.x; delta0[2 * k + 1] = t.delta0.y; delta1[2 * k] = t.delta1.x; delta1[2 * k + 1] = t.delta1.y;
    kick[2 * k] = t.kick.x; kick[2 * k + 1] = t.kick.y;
  }
}

// geometry AABBs by themselves (checked against the reference's own Geometry sources in tests/test_oracle_vs_reference.py)
// type: RigidBodyGeometryType (0 box, 1 sphere); R row-major; out = min(3), max(3)
void orc_rb3d_aabb( int type, double r, const double* half, const double* cm, const double* R, double* out )
{
  RB3DScene s;
  RB3DGeometry g;
  g.type = uint32_t( type ); g.r = r; g.half = V3{ half[0], half[1], half[2] }; g.mesh = 0;
  s.geometry.push_back( g );
  s.geo_of_body.push_back( 0 );
  M3 Rm;
  for( int k = 0; k < 9; ++k ) { Rm.m[k] = R[k]; }
  Box<3> b;
  computeAABB( s, 0, V3{ cm[0], cm[1], cm[2] }, Rm, b );
  for( int k = 0; k < 3; ++k ) { out[k] = b.lo[k]; out[3 + k] = b.hi[k]; }
}
// type 0 circle, 1 box; swept != 0: computeCollisionAABB, else computeAABB at q1; out = min(2), max(2)
void orc_rb2d_aabb( int type, double r, const double* half, const double* q0b, const double* q1b, int swept, double* out )
{
  const RB2DGeometry g{ uint32_t( type ), r, V2{ half[0], half[1] } };
  Box<2> b;
  if( swept ) { computeCollisionAABB( g, q0b, q1b, b ); } else { computeAABBAt( g, V2{ q1b[0], q1b[1] }, q1b[2], b ); }
  out[0] = b.lo[0]; out[1] = b.lo[1]; out[2] = b.hi[0]; out[3] = b.hi[1];
}

// ---- leaf routines by themselves (checked against the reference's own sources in tests/test_oracle_vs_reference.py) ----
int orc_box_box_3d( const double* cm0, const double* R0, const double* side0, const double* cm1, const double* R1, const double* side1, double* n, double* points )
{
  using namespace orc;
  M3 A, B;
  for( int i = 0; i < 3; ++i ) { for( int j = 0; j < 3; ++j ) { A.m[3 * i + j] = R0[3 * i + j]; B.m[3 * i + j] = R1[3 * i + j]; } }
  V3 nn{ 0.0, 0.0, 0.0 };
  std::vector<V3> pts;
  boxbox::isActive( V3{ cm0[0], cm0[1], cm0[2] }, A, V3{ side0[0], side0[1], side0[2] }, V3{ cm1[0], cm1[1], cm1[2] }, B, V3{ side1[0], side1[1], side1[2] }, nn, pts );
  n[0] = nn.x; n[1] = nn.y; n[2] = nn.z;
  int k = 0;
  for( const V3& p : pts ) { if( k < 8 ) { points[3 * k] = p.x; points[3 * k + 1] = p.y; points[3 * k + 2] = p.z; } ++k; }
  return k;
}
int orc_box_box_2d( const double* x0, const double theta0, const double* r0, const double* x1, const double theta1, const double* r1, double* n, double* points )
{
  using namespace orc;
  V2 nn{ 0.0, 0.0 };
  std::vector<V2> pts;
  boxbox2d::isActive( V2{ x0[0], x0[1] }, theta0, V2{ r0[0], r0[1] }, V2{ x1[0], x1[1] }, theta1, V2{ r1[0], r1[1] }, nn, pts );
  n[0] = nn.x; n[1] = nn.y;
  int k = 0;
  for( const V2& p : pts ) { if( k < 2 ) { points[2 * k] = p.x; points[2 * k + 1] = p.y; } ++k; }
  return k;
}
int orc_circle_box_2d( const double* x0, const double r0, const double* x1, const double theta1, const double* r1, double* n, double* p )
{
  using namespace orc;
  V2 nn{ 0.0, 0.0 }, pp{ 0.0, 0.0 };
  const bool hit = circleBoxActive( V2{ x0[0], x0[1] }, r0, V2{ x1[0], x1[1] }, theta1, V2{ r1[0], r1[1] }, nn, pp );
  n[0] = nn.x; n[1] = nn.y; p[0] = pp.x; p[1] = pp.y;
  return hit ? 1 : 0;
}

}
