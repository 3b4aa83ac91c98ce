// tests/rb3d_snapshot_harness.cpp -- TEST INFRASTRUCTURE: the product's scisim_b200/csrc/sg_rb3d_snapshot.h (the header sg_rb3d.cu includes), compiled for
// the host so that the CPU suite can compare its bytes with the reference's own RigidBody3DState::serialize (tests/test_rb3d_snapshot_cpu.py).
#include "../scisim_b200/csrc/sg_rb3d_snapshot.h"

extern "C"
{

// all arrays as in sg_snapshot::Rb3dState; returns the number of bytes (written when they fit cap); 0 = not serialisable
uint64_t snap_serialize( const uint32_t n, const double* q, const double* v, const double* m, const double* I0, const double* I, const double* Iinv, const uint8_t* fixed,
                         const uint32_t* geo_of_body, const uint32_t ngeo, const uint32_t* geo_type, const double* geo_r, const double* geo_half, const double* g,
                         const uint32_t npl, const double* px, const double* pn, const uint32_t ncyl, const double* cx, const double* cax, const double* cr,
                         const uint32_t npo, const double* pax, const double* pan, const double* pbx, const double* pbn, const int32_t* mult, void* buf, const uint64_t cap,
                         const unsigned char* const* geo_record = nullptr, const uint64_t* geo_record_bytes = nullptr )
{
  sg_snapshot::Rb3dState s;
  s.n = n;
  s.q.assign( q, q + 12 * size_t( n ) ); s.v.assign( v, v + 6 * size_t( n ) ); s.m.assign( m, m + n ); s.I0.assign( I0, I0 + 3 * size_t( n ) );
  s.I.assign( I, I + 9 * size_t( n ) ); s.Iinv.assign( Iinv, Iinv + 9 * size_t( n ) );
  s.fixed.assign( fixed, fixed + n ); s.geo_of_body.assign( geo_of_body, geo_of_body + n );
  s.geo_type.assign( geo_type, geo_type + ngeo ); s.geo_r.assign( geo_r, geo_r + ngeo ); s.geo_half.assign( geo_half, geo_half + 3 * size_t( ngeo ) );
  for( int k = 0; k < 3; ++k ) { s.g[k] = g[k]; }
  s.plane_x.assign( px, px + 3 * size_t( npl ) ); s.plane_n.assign( pn, pn + 3 * size_t( npl ) );
  s.cyl_x.assign( cx, cx + 3 * size_t( ncyl ) ); s.cyl_axis.assign( cax, cax + 3 * size_t( ncyl ) ); s.cyl_r.assign( cr, cr + ncyl );
  s.portal_ax.assign( pax, pax + 3 * size_t( npo ) ); s.portal_an.assign( pan, pan + 3 * size_t( npo ) ); s.portal_bx.assign( pbx, pbx + 3 * size_t( npo ) );
  s.portal_bn.assign( pbn, pbn + 3 * size_t( npo ) ); s.portal_mult.assign( mult, mult + 3 * size_t( npo ) );
  if( geo_record != nullptr )
  {
    // triangle meshes: the record of each geometry ( type byte included ), empty for boxes and spheres
    s.geo_blob.resize( ngeo );
    for( uint32_t k = 0; k < ngeo; ++k ) { if( geo_record_bytes[k] > 0 ) { s.geo_blob[k].assign( geo_record[k], geo_record[k] + geo_record_bytes[k] ); } }
  }
  sg_snapshot::Sink out{ static_cast<unsigned char*>( buf ), cap, 0 };
  if( !sg_snapshot::serialize( s, out ) ) { return 0; }
  return out.n;
}

// the triangle-mesh records of a snapshot: for geometry k, *offset / *bytes of its record in the stream ( 0 bytes for boxes and spheres ) and the sizes of the arrays
// parse() extracted for sg_rb3d_add_mesh ( vertices, samples, hull vertices, grid cells ); returns the parser's code
int snap_mesh_records( const void* in_buf, const uint64_t in_bytes, const uint32_t ngeo_cap, uint32_t* ngeo, uint64_t* bytes, uint32_t* sizes /* 4 per geometry */, double* sdf_sum )
{
  sg_snapshot::Source in{ static_cast<const unsigned char*>( in_buf ), in_bytes, 0, true };
  sg_snapshot::Rb3dState s;
  const char* why = "";
  const int rc = sg_snapshot::parse( in, s, &why );
  if( rc != 0 ) { return rc; }
  *ngeo = uint32_t( s.geo_type.size() );
  for( uint32_t k = 0; k < *ngeo && k < ngeo_cap; ++k )
  {
    bytes[k] = s.geo_blob[k].size();
    sizes[4 * k] = uint32_t( s.mesh[k].verts.size() / 3 ); sizes[4 * k + 1] = uint32_t( s.mesh[k].samples.size() / 3 ); sizes[4 * k + 2] = uint32_t( s.mesh[k].hull.size() / 3 );
    sizes[4 * k + 3] = uint32_t( s.mesh[k].sdf.size() );
    sdf_sum[k] = 0.0;
    for( const double d : s.mesh[k].sdf ) { sdf_sum[k] += d; }
  }
  return 0;
}

// parse a snapshot and write it again: returns the parser's code ( 0 ok, 1 malformed, 2 unsupported ); *bytes_out = length of the re-serialised stream
int snap_roundtrip( const void* in_buf, const uint64_t in_bytes, void* out_buf, const uint64_t cap, uint64_t* bytes_out, uint32_t* n_out, double* g_out )
{
  sg_snapshot::Source in{ static_cast<const unsigned char*>( in_buf ), in_bytes, 0, true };
  sg_snapshot::Rb3dState s;
  const char* why = "";
  const int rc = sg_snapshot::parse( in, s, &why );
  if( rc != 0 ) { return rc; }
  sg_snapshot::Sink out{ static_cast<unsigned char*>( out_buf ), cap, 0 };
  sg_snapshot::serialize( s, out );
  *bytes_out = out.n; *n_out = s.n;
  for( int k = 0; k < 3; ++k ) { g_out[k] = s.g[k]; }
  return ( in.n == in_bytes ) ? 0 : 1;
}

}
