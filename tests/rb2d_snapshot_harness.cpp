// tests/rb2d_snapshot_harness.cpp -- TEST INFRASTRUCTURE: the product's scisim_b200/csrc/sg_rb2d_snapshot.h (the header sg_rb2d.cu includes), compiled for
// the host so that the CPU suite can compare its bytes with the reference's own RigidBody2DState::serialize (tests/test_rb2d_snapshot_cpu.py).
#include "../scisim_b200/csrc/sg_rb2d_snapshot.h"

extern "C"
{

// all arrays as in sg_snapshot::Rb2dState; returns the number of bytes (written when they fit cap); 0 = not serialisable
uint64_t snap2d_serialize( const uint32_t n, const double* q, const double* v, const double* M, const uint8_t* fixed, const uint32_t* geo_of_body,
                           const uint32_t ngeo, const uint32_t* geo_type, const double* geo_r, const double* geo_half, const double* g,
                           const uint32_t npl, const double* px, const double* pn, const double* pt,
                           const uint32_t npo, const double* pax, const double* pan, const double* pat, const double* pbx, const double* pbn, const double* pbt,
                           const double* pv, const double* pbounds, const double* pdx, void* buf, const uint64_t cap )
{
  sg_snapshot::Rb2dState s;
  s.n = n;
  s.q.assign( q, q + 3 * size_t( n ) ); s.v.assign( v, v + 3 * size_t( n ) ); s.M.assign( M, M + 3 * size_t( n ) );
  s.fixed.assign( fixed, fixed + n ); s.geo_of_body.assign( geo_of_body, geo_of_body + n );
  s.geo_type.assign( geo_type, geo_type + ngeo ); s.geo_r.assign( geo_r, geo_r + ngeo ); s.geo_half.assign( geo_half, geo_half + 2 * size_t( ngeo ) );
  s.g[0] = g[0]; s.g[1] = g[1];
  s.plane_x.assign( px, px + 2 * size_t( npl ) ); s.plane_n.assign( pn, pn + 2 * size_t( npl ) ); s.plane_t.assign( pt, pt + 2 * size_t( npl ) );
  s.portal_ax.assign( pax, pax + 2 * size_t( npo ) ); s.portal_an.assign( pan, pan + 2 * size_t( npo ) ); s.portal_at.assign( pat, pat + 2 * size_t( npo ) );
  s.portal_bx.assign( pbx, pbx + 2 * size_t( npo ) ); s.portal_bn.assign( pbn, pbn + 2 * size_t( npo ) ); s.portal_bt.assign( pbt, pbt + 2 * size_t( npo ) );
  s.portal_v.assign( pv, pv + npo ); s.portal_bounds.assign( pbounds, pbounds + npo ); s.portal_dx.assign( pdx, pdx + npo );
  sg_snapshot::Sink out{ static_cast<unsigned char*>( buf ), cap, 0 };
  if( !sg_snapshot::serialize( s, out ) ) { return 0; }
  return out.n;
}

// parse a snapshot and write it again: returns the parser's code ( 0 ok, 1 malformed, 2 unsupported ); *bytes_out = length of the re-serialised stream
int snap2d_roundtrip( const void* in_buf, const uint64_t in_bytes, void* out_buf, const uint64_t cap, uint64_t* bytes_out, uint32_t* n_out, double* g_out )
{
  sg_snapshot::Source in{ static_cast<const unsigned char*>( in_buf ), in_bytes, 0, true };
  sg_snapshot::Rb2dState s;
  const char* why = "";
  const int rc = sg_snapshot::parse( in, s, &why );
  if( rc != 0 ) { return rc; }
  sg_snapshot::Sink out{ static_cast<unsigned char*>( out_buf ), cap, 0 };
  sg_snapshot::serialize( s, out );
  *bytes_out = out.n; *n_out = s.n;
  g_out[0] = s.g[0]; g_out[1] = s.g[1];
  return ( in.n == in_bytes ) ? 0 : 1;
}

}
