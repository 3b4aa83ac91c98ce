// tests/portal_math_harness.cpp -- TEST INFRASTRUCTURE.  Compiles the product's own portal arithmetic
// (scisim_b200/csrc/sg_portal2d.h, the header the CUDA kernels include) as plain C++ so that the CPU test suite can run
// it against the oracle and the reference's PlanarPortal.cpp bit for bit.  Built by tests/test_portals_cpu.py with
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC
// (no FMA contraction, like the library's -fmad=false).  Nothing here is shipped.
#include "../scisim_b200/csrc/sg_portal2d.h"

#include <cstring>

static SgPortals2D g_portals;

extern "C"
{

void pm_set_portals( uint32_t n, const double* ax, const double* an, const double* bx, const double* bn, const double* v, const double* bounds )
{
  std::memset( &g_portals, 0, sizeof( g_portals ) );
  g_portals.n = n;
  for( uint32_t p = 0; p < n; ++p )
  {
    SgPortal2D& pt = g_portals.p[p];
    for( int k = 0; k < 2; ++k ) { pt.ax[k] = ax[2 * p + k]; pt.bx[k] = bx[2 * p + k]; }
    sg_portal_plane_frame( an + 2 * p, pt.an, pt.at );
    sg_portal_plane_frame( bn + 2 * p, pt.bn, pt.bt );
    pt.v = v[p]; pt.bounds = bounds[p]; pt.dx = 0.0;
  }
}

void pm_update_portals( double t, double* dx_out )
{
  for( uint32_t p = 0; p < g_portals.n; ++p )
  {
    g_portals.p[p].dx = sg_portal_offset( g_portals.p[p].v, g_portals.p[p].bounds, t );
    if( dx_out != nullptr ) { dx_out[p] = g_portals.p[p].dx; }
  }
}

// same layout as orc_ball2d_portal_probe (oracle/capi.cpp)
uint32_t pm_probe( uint32_t p, const double* x, double r, double* out )
{
  const SgPortal2D& pt = g_portals.p[p];
  const SgVec2 xin{ x[0], x[1] };
  const SgVec2 a = sg_portal_teleport( pt, false, xin ), b = sg_portal_teleport( pt, true, xin );
  const SgVec2 tb = sg_portal_teleport_ball( pt, xin, r ), ti = sg_portal_teleport_point_inside( pt, xin );
  const SgVec2 kb = sg_portal_kinematic_velocity_of_ball( pt, xin, r ), kp = sg_portal_kinematic_velocity_of_point( pt, xin );
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y; out[4] = tb.x; out[5] = tb.y; out[6] = ti.x; out[7] = ti.y;
  out[8] = kb.x; out[9] = kb.y; out[10] = kp.x; out[11] = kp.y;
  const int touch = sg_portal_touch( pt, xin, r );
  uint32_t flags = 0u;
  if( touch == 1 || touch == 2 ) { flags |= 1u; }
  if( touch == 2 ) { flags |= 2u; }
  if( touch == 3 ) { flags |= 4u; }
  if( sg_portal_point_inside( pt, xin ) ) { flags |= 8u; }
  return flags;
}

void pm_enforce( uint32_t n, double* q, double* v )
{
  for( uint32_t b = 0; b < n; ++b )
  {
    SgVec2 x{ q[2 * b], q[2 * b + 1] }, w{ v[2 * b], v[2 * b + 1] };
    sg_portals_enforce( g_portals, x, w );
    q[2 * b] = x.x; q[2 * b + 1] = x.y; v[2 * b] = w.x; v[2 * b + 1] = w.y;
  }
}

// the launch sequence of the teleported-collision sort (sg_ball2d_portals.cuh), one "launch" = one pass over e
void pm_bitonic_sort( uint32_t m, unsigned long long* keys, uint32_t* idxs )
{
  for( uint32_t k = 2u; k <= m && k != 0u; k <<= 1 )
  {
    for( uint32_t j = k >> 1; j > 0u; j >>= 1 )
    {
      for( uint32_t e = 0; e < m; ++e )
      {
        const uint32_t f = sg_bitonic_partner( e, j );
        if( f <= e ) { continue; }
        const bool a_less = sg_tele_less( keys[e], idxs[e], keys[f], idxs[f] );
        const bool b_less = sg_tele_less( keys[f], idxs[f], keys[e], idxs[e] );
        const bool swap = sg_bitonic_ascending( e, k ) ? b_less : a_less;
        if( swap ) { const unsigned long long tk = keys[e]; keys[e] = keys[f]; keys[f] = tk; const uint32_t ti = idxs[e]; idxs[e] = idxs[f]; idxs[f] = ti; }
      }
    }
  }
}

// TeleportedCollision ordering + the collision test at q ( teleportedBallBallCollisionHappens )
int pm_tele_happens( uint32_t b0, uint32_t b1, uint32_t p0, uint32_t p1, const double* q, const double* r, uint32_t* ordered )
{
  const SgTeleCollision c = sg_tele_collision( b0, b1, p0, p1 );
  ordered[0] = c.b0; ordered[1] = c.b1; ordered[2] = c.p0; ordered[3] = c.p1;
  const SgVec2 xa = sg_tele_center( g_portals, c.p0, SgVec2{ q[2 * c.b0], q[2 * c.b0 + 1] } );
  const SgVec2 xb = sg_tele_center( g_portals, c.p1, SgVec2{ q[2 * c.b1], q[2 * c.b1 + 1] } );
  return sg_ball_ball_active( xa, xb, r[c.b0], r[c.b1] ) ? 1 : 0;
}

}
