// RigidBody3DState's binary snapshot (rigidbody3d/RigidBody3DState.cpp:586-668), written and parsed on the host from plain arrays.  Plain C++ (no CUDA):
// sg_rb3d.cu gathers the arrays from the device-resident state and calls these; the CPU suite compiles the same header and compares its bytes with what
// the reference's own RigidBody3DState::serialize writes (tests/test_rb3d_snapshot_cpu.py).
//
// Layout, in the order RigidBody3DState::serialize writes it (scisim/Utilities.h:43-94, scisim/Math/MathUtilities.h:42-60, MathUtilities.cpp:142-154):
//   nbodies                unsigned
//   q, v                   Eigen::Index rows + doubles ( q: 3N positions | 9N row-major rotations; v: 3N linear | 3N angular )
//   M0, Minv0              sparse 6N x 6N diagonal: m, m, m per body, then the body-frame inertias ( Minv0: 1.0 / each )
//   M, Minv                sparse 6N x 6N, 12N non-zeros: the 3N linear entries, then per body the 3 x 3 world-space block, column by column
//                          ( sparse = rows, cols, nnz as Eigen::Index; nnz inner indices, cols + 1 outer indices as int; nnz doubles )
//   fixed                  size_t count + one byte per body
//   geometry               size_t count + per geometry: RigidBodyGeometryType ( uint8: BOX 0, SPHERE 1, TRIANGLE_MESH 3 ) then half widths (3) / radius /
//                          the mesh's whole record (RigidBodyTriangleMesh::serialize, rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp:215-232): file name
//                          ( size_t length + chars ), vertices and faces ( Eigen::Index columns + 3 per column; doubles / unsigned ), volume, I / rho (3),
//                          centre of mass (3), R (9), surface samples and convex-hull vertices ( as the vertices ), cell delta (3), grid dimensions
//                          ( 3 unsigned ), grid origin (3), signed distances ( Eigen::Index rows + doubles, x fastest ), grid end (3).  The library keeps a
//                          mesh's record as the caller handed it over (sg_rb3d_set_mesh_snapshot) and writes it back verbatim; reading one yields the
//                          arrays sg_rb3d_add_mesh takes
//   geometry indices       size_t count + unsigned per body
//   forces                 size_t count + { size_t length + "near_earth_gravity", g (3 doubles) }
//   static planes          size_t count + { x, n, v, omega (3 doubles each) }
//   static cylinders       size_t count + { x, axis (3 each), theta, v, omega (3 each), r }
//   planar portals         size_t count + { plane A, plane B as above, multiplier (3 int) }
//   boundary behaviour     SimBoundaryBehavior ( int: NONE 0 ), boundary min, max (3 doubles each)
#ifndef SG_RB3D_SNAPSHOT_H
#define SG_RB3D_SNAPSHOT_H

#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace sg_snapshot
{

struct Sink
{
  unsigned char* p; uint64_t cap; uint64_t n;
  void put( const void* src, const uint64_t bytes ) { if( p != nullptr && n + bytes <= cap ) { memcpy( p + n, src, bytes ); } n += bytes; }
  template<typename T> void val( const T v ) { put( &v, sizeof( T ) ); }
};

struct Source
{
  const unsigned char* p; uint64_t cap; uint64_t n; bool ok;
  const unsigned char* take( const uint64_t bytes ) { if( !ok || n + bytes > cap ) { ok = false; return nullptr; } const unsigned char* at = p + n; n += bytes; return at; }
  template<typename T> T val() { T v{}; const unsigned char* at = take( sizeof( T ) ); if( at != nullptr ) { memcpy( &v, at, sizeof( T ) ); } return v; }
  void doubles( double* out, const uint64_t count ) { const unsigned char* at = take( count * 8 ); if( at != nullptr && count > 0 ) { memcpy( out, at, count * 8 ); } }
};

// everything a snapshot holds, as plain host arrays
struct Rb3dState
{
  uint32_t n = 0;
  std::vector<double> q, v;            // 12 n, 6 n
  std::vector<double> m, I0;           // n, 3 n
  std::vector<double> I, Iinv;         // 9 n each: the world-space blocks as M / Minv store them (column-major values of each 3 x 3 block)
  std::vector<uint8_t> fixed;          // n
  std::vector<uint32_t> geo_of_body;   // n
  std::vector<uint32_t> geo_type;      // 0 box, 1 sphere, 3 triangle mesh
  std::vector<double> geo_r, geo_half; // ngeo, 3 ngeo
  // triangle meshes: geo_blob[k] = the record of geometry k as the reference writes it, type byte included (empty for boxes and spheres; may be left
  // unsized when there is no mesh); parse() also fills mesh[k] with what sg_rb3d_add_mesh takes
  struct Mesh { std::vector<double> verts, samples, hull, sdf; double delta[3] = { 0.0, 0.0, 0.0 }, origin[3] = { 0.0, 0.0, 0.0 }; uint32_t dims[3] = { 0u, 0u, 0u }; };
  std::vector<std::vector<unsigned char>> geo_blob;
  std::vector<Mesh> mesh;
  double g[3] = { 0.0, 0.0, 0.0 };
  std::vector<double> plane_x, plane_n;            // 3 each
  std::vector<double> cyl_x, cyl_axis, cyl_r;      // 3, 3, 1 each
  std::vector<double> portal_ax, portal_an, portal_bx, portal_bn; // 3 each
  std::vector<int32_t> portal_mult;                // 3 each
};

inline void put_diagonal( Sink& out, const uint32_t n, const std::vector<double>& m, const std::vector<double>& I0, const bool inverse )
{
  const long long dofs = 6ll * n;
  out.val<long long>( dofs ); out.val<long long>( dofs ); out.val<long long>( dofs );
  for( long long k = 0; k < dofs; ++k ) { out.val<int>( int( k ) ); }
  for( long long k = 0; k <= dofs; ++k ) { out.val<int>( int( k ) ); }
  for( uint32_t b = 0; b < n; ++b ) { for( int k = 0; k < 3; ++k ) { out.val<double>( inverse ? 1.0 / m[b] : m[b] ); } }          // formBodySpace(Inverse)MassMatrix (RigidBody3DState.cpp:70-138)
  for( uint32_t b = 0; b < n; ++b ) { for( int k = 0; k < 3; ++k ) { out.val<double>( inverse ? 1.0 / I0[3 * b + k] : I0[3 * b + k] ); } }
}

inline void put_world( Sink& out, const uint32_t n, const std::vector<double>& m, const std::vector<double>& blocks, const bool inverse )
{
  const long long dofs = 6ll * n, nnz = 12ll * n;
  out.val<long long>( dofs ); out.val<long long>( dofs ); out.val<long long>( nnz );
  for( uint32_t c = 0; c < 3 * n; ++c ) { out.val<int>( int( c ) ); }
  for( uint32_t b = 0; b < n; ++b ) { for( int c = 0; c < 3; ++c ) { for( int r = 0; r < 3; ++r ) { out.val<int>( int( 3 * n + 3 * b + r ) ); } } }
  for( uint32_t c = 0; c < 3 * n; ++c ) { out.val<int>( int( c ) ); }
  for( uint32_t c = 0; c <= 3 * n; ++c ) { out.val<int>( int( 3 * n + 3 * c ) ); }
  for( uint32_t b = 0; b < n; ++b ) { for( int k = 0; k < 3; ++k ) { out.val<double>( inverse ? 1.0 / m[b] : m[b] ); } }
  out.put( blocks.data(), uint64_t( 9 ) * n * 8 );
}

inline void put_plane( Sink& out, const double* x, const double* nrm )
{
  out.put( x, 24 ); out.put( nrm, 24 );
  for( int k = 0; k < 6; ++k ) { out.val<double>( 0.0 ); } // m_v, m_omega: planes of this path do not move
}

// returns false when the state cannot be written in the reference's format (a geometry that is neither box nor sphere nor a mesh with its record)
inline bool serialize( const Rb3dState& s, Sink& out )
{
  const uint32_t n = s.n;
  out.val<unsigned>( n );
  out.val<long long>( 12ll * n ); out.put( s.q.data(), uint64_t( 12 ) * n * 8 );
  out.val<long long>( 6ll * n ); out.put( s.v.data(), uint64_t( 6 ) * n * 8 );
  put_diagonal( out, n, s.m, s.I0, false );
  put_diagonal( out, n, s.m, s.I0, true );
  put_world( out, n, s.m, s.I, false );
  put_world( out, n, s.m, s.Iinv, true );
  out.val<size_t>( size_t( n ) );
  for( uint32_t b = 0; b < n; ++b ) { out.val<unsigned char>( s.fixed[b] ? 1 : 0 ); }
  out.val<size_t>( s.geo_type.size() );
  for( size_t k = 0; k < s.geo_type.size(); ++k )
  {
    if( s.geo_type[k] == 0u ) { out.val<unsigned char>( 0 ); out.put( &s.geo_half[3 * k], 24 ); }
    else if( s.geo_type[k] == 1u ) { out.val<unsigned char>( 1 ); out.val<double>( s.geo_r[k] ); }
    else if( s.geo_type[k] == 3u && k < s.geo_blob.size() && !s.geo_blob[k].empty() ) { out.put( s.geo_blob[k].data(), s.geo_blob[k].size() ); }
    else { return false; }
  }
  out.val<size_t>( size_t( n ) );
  for( uint32_t b = 0; b < n; ++b ) { out.val<unsigned>( s.geo_of_body[b] ); }
  out.val<size_t>( size_t( 1 ) );
  const char name[] = "near_earth_gravity";
  out.val<size_t>( sizeof( name ) - 1 ); out.put( name, sizeof( name ) - 1 );
  out.put( s.g, 24 );
  const size_t npl = s.plane_x.size() / 3;
  out.val<size_t>( npl );
  for( size_t k = 0; k < npl; ++k ) { put_plane( out, &s.plane_x[3 * k], &s.plane_n[3 * k] ); }
  const size_t ncy = s.cyl_r.size();
  out.val<size_t>( ncy );
  for( size_t k = 0; k < ncy; ++k )
  {
    out.put( &s.cyl_x[3 * k], 24 ); out.put( &s.cyl_axis[3 * k], 24 );
    out.val<double>( 0.0 );                                     // m_theta
    for( int c = 0; c < 6; ++c ) { out.val<double>( 0.0 ); }    // m_v, m_omega
    out.val<double>( s.cyl_r[k] );
  }
  const size_t npo = s.portal_mult.size() / 3;
  out.val<size_t>( npo );
  for( size_t k = 0; k < npo; ++k )
  {
    put_plane( out, &s.portal_ax[3 * k], &s.portal_an[3 * k] );
    put_plane( out, &s.portal_bx[3 * k], &s.portal_bn[3 * k] );
    out.put( &s.portal_mult[3 * k], 12 );
  }
  out.val<int>( 0 ); // SimBoundaryBehavior::NONE
  for( int k = 0; k < 3; ++k ) { out.val<double>( std::numeric_limits<double>::min() ); }   // RigidBody3DState.cpp:38-39
  for( int k = 0; k < 3; ++k ) { out.val<double>( std::numeric_limits<double>::max() ); }
  return true;
}

// 3 x ncols doubles behind an Eigen::Index column count (a Matrix3Xsc)
inline bool take_matrix3x( Source& in, std::vector<double>& out )
{
  const long long ncols = in.val<long long>();
  if( !in.ok || ncols < 0 || uint64_t( ncols ) > ( in.cap - in.n ) / 24 ) { in.ok = false; return false; }
  out.resize( size_t( 3 ) * size_t( ncols ) );
  in.doubles( out.data(), uint64_t( 3 ) * uint64_t( ncols ) );
  return in.ok;
}

// a triangle mesh's record after its type byte (RigidBodyTriangleMesh( std::istream& ), RigidBodyTriangleMesh.cpp:105-129)
inline bool take_mesh( Source& in, Rb3dState::Mesh& m )
{
  const size_t len = in.val<size_t>();
  if( !in.ok || len > in.cap - in.n ) { in.ok = false; return false; }
  in.take( len );                                                          // m_input_file_name
  if( !take_matrix3x( in, m.verts ) ) { return false; }
  const long long nfaces = in.val<long long>();
  if( !in.ok || nfaces < 0 || uint64_t( nfaces ) > ( in.cap - in.n ) / 12 ) { in.ok = false; return false; }
  in.take( uint64_t( nfaces ) * 12 );                                      // m_faces
  in.take( 8 + 24 + 24 + 72 );                                             // m_volume, m_I_on_rho, m_center_of_mass, m_R
  if( !take_matrix3x( in, m.samples ) || !take_matrix3x( in, m.hull ) ) { return false; }
  in.doubles( m.delta, 3 );
  for( int k = 0; k < 3; ++k ) { m.dims[k] = in.val<unsigned>(); }
  in.doubles( m.origin, 3 );
  const long long nsd = in.val<long long>();
  if( !in.ok || nsd < 0 || uint64_t( nsd ) > ( in.cap - in.n ) / 8 || uint64_t( nsd ) != uint64_t( m.dims[0] ) * m.dims[1] * m.dims[2] ) { in.ok = false; return false; }
  m.sdf.resize( size_t( nsd ) );
  in.doubles( m.sdf.data(), uint64_t( nsd ) );
  in.take( 24 );                                                           // m_grid_end: origin + ( dims - 1 ) * delta, recomputed by sg_rb3d_add_mesh
  return in.ok;
}

inline bool take_plane( Source& in, double* x, double* nrm )
{
  in.doubles( x, 3 ); in.doubles( nrm, 3 );
  double rest[6];
  in.doubles( rest, 6 );
  return in.ok;
}

// 0 ok, 1 malformed / truncated, 2 holds something this path does not support (why says what)
inline int parse( Source& in, Rb3dState& s, const char** why )
{
  *why = "";
  const unsigned n = in.val<unsigned>();
  s.n = n;
  if( !in.ok ) { *why = "truncated"; return 1; }
  if( in.val<long long>() != 12ll * n ) { *why = "q does not hold 12 doubles per body"; return 1; }
  s.q.resize( size_t( 12 ) * n ); in.doubles( s.q.data(), uint64_t( 12 ) * n );
  if( in.val<long long>() != 6ll * n ) { *why = "v does not hold 6 doubles per body"; return 1; }
  s.v.resize( size_t( 6 ) * n ); in.doubles( s.v.data(), uint64_t( 6 ) * n );
  s.m.resize( n ); s.I0.resize( size_t( 3 ) * n ); s.I.resize( size_t( 9 ) * n ); s.Iinv.resize( size_t( 9 ) * n );
  for( int mat = 0; mat < 4; ++mat )
  {
    const long long rows = in.val<long long>(), cols = in.val<long long>(), nnz = in.val<long long>();
    const long long want = ( mat < 2 ) ? 6ll * n : 12ll * n;
    if( !in.ok || rows != 6ll * n || cols != 6ll * n || nnz != want ) { *why = "a mass matrix of another shape"; return 1; }
    in.take( uint64_t( nnz ) * 4 ); in.take( uint64_t( cols + 1 ) * 4 );
    std::vector<double> vals( static_cast<size_t>( nnz ) );
    in.doubles( vals.data(), uint64_t( nnz ) );
    if( !in.ok ) { *why = "truncated"; return 1; }
    if( mat == 0 ) { for( unsigned b = 0; b < n; ++b ) { s.m[b] = vals[3 * size_t( b )]; for( int k = 0; k < 3; ++k ) { s.I0[3 * size_t( b ) + k] = vals[3 * size_t( n ) + 3 * size_t( b ) + k]; } } }
    if( mat == 2 ) { memcpy( s.I.data(), vals.data() + 3 * size_t( n ), size_t( 9 ) * n * 8 ); }
    if( mat == 3 ) { memcpy( s.Iinv.data(), vals.data() + 3 * size_t( n ), size_t( 9 ) * n * 8 ); }
  }
  if( in.val<size_t>() != size_t( n ) ) { *why = "fixed flags of another length"; return 1; }
  s.fixed.resize( n );
  for( unsigned b = 0; b < n; ++b ) { s.fixed[b] = in.val<unsigned char>(); }
  const size_t ngeo = in.val<size_t>();
  if( !in.ok || ngeo > ( 1u << 24 ) || ngeo > ( in.cap - in.n ) / 9 ) { *why = "bad geometry count"; return 1; } // a geometry takes at least a type byte and a double
  s.geo_type.assign( ngeo, 0u ); s.geo_r.assign( ngeo, 0.0 ); s.geo_half.assign( 3 * ngeo, 0.0 );
  s.geo_blob.assign( ngeo, std::vector<unsigned char>() ); s.mesh.assign( ngeo, Rb3dState::Mesh() );
  for( size_t k = 0; k < ngeo; ++k )
  {
    const uint64_t at = in.n;
    const unsigned char t = in.val<unsigned char>();
    if( !in.ok ) { *why = "truncated"; return 1; }
    if( t == 0 ) { s.geo_type[k] = 0u; in.doubles( &s.geo_half[3 * k], 3 ); }
    else if( t == 1 ) { s.geo_type[k] = 1u; s.geo_r[k] = in.val<double>(); }
    else if( t == 3 )
    {
      s.geo_type[k] = 3u;
      if( !take_mesh( in, s.mesh[k] ) ) { *why = "a malformed or truncated triangle-mesh record"; return 1; }
      s.geo_blob[k].assign( in.p + at, in.p + in.n );
    }
    else { *why = "a geometry other than box, sphere or triangle mesh (staples are not on this path)"; return 2; }
  }
  if( in.val<size_t>() != size_t( n ) ) { *why = "geometry indices of another length"; return 1; }
  s.geo_of_body.resize( n );
  for( unsigned b = 0; b < n; ++b ) { s.geo_of_body[b] = in.val<unsigned>(); if( in.ok && s.geo_of_body[b] >= ngeo ) { *why = "geometry index out of range"; return 1; } }
  const size_t nf = in.val<size_t>();
  s.g[0] = s.g[1] = s.g[2] = 0.0;
  for( size_t k = 0; k < nf && in.ok; ++k )
  {
    const size_t len = in.val<size_t>();
    const unsigned char* nm = in.take( len );
    if( nm == nullptr || len != 18 || memcmp( nm, "near_earth_gravity", 18 ) != 0 ) { *why = "a force other than near_earth_gravity (the reference exits too: RigidBody3DState.cpp:642-646)"; return 2; }
    double g[3];
    in.doubles( g, 3 );
    for( int c = 0; c < 3; ++c ) { s.g[c] += g[c]; } // forces accumulate
  }
  const size_t npl = in.val<size_t>();
  if( !in.ok || npl > 4096 ) { *why = "bad plane count"; return 1; }
  s.plane_x.resize( 3 * npl ); s.plane_n.resize( 3 * npl );
  for( size_t k = 0; k < npl; ++k ) { take_plane( in, &s.plane_x[3 * k], &s.plane_n[3 * k] ); }
  const size_t ncy = in.val<size_t>();
  if( !in.ok || ncy > 4096 ) { *why = "bad cylinder count"; return 1; }
  s.cyl_x.resize( 3 * ncy ); s.cyl_axis.resize( 3 * ncy ); s.cyl_r.resize( ncy );
  for( size_t k = 0; k < ncy; ++k )
  {
    in.doubles( &s.cyl_x[3 * k], 3 ); in.doubles( &s.cyl_axis[3 * k], 3 );
    double rest[7];
    in.doubles( rest, 7 );
    s.cyl_r[k] = in.val<double>();
  }
  const size_t npo = in.val<size_t>();
  if( !in.ok || npo > 4096 ) { *why = "bad portal count"; return 1; }
  s.portal_ax.resize( 3 * npo ); s.portal_an.resize( 3 * npo ); s.portal_bx.resize( 3 * npo ); s.portal_bn.resize( 3 * npo ); s.portal_mult.resize( 3 * npo );
  for( size_t k = 0; k < npo; ++k )
  {
    take_plane( in, &s.portal_ax[3 * k], &s.portal_an[3 * k] );
    take_plane( in, &s.portal_bx[3 * k], &s.portal_bn[3 * k] );
    for( int c = 0; c < 3; ++c ) { s.portal_mult[3 * k + c] = in.val<int>(); }
  }
  const int behaviour = in.val<int>();
  double lim[6];
  in.doubles( lim, 6 );
  if( !in.ok ) { *why = "truncated"; return 1; }
  if( behaviour != 0 ) { *why = "a simulation boundary with an exit treatment (outside this path)"; return 2; }
  return 0;
}

}

#endif
