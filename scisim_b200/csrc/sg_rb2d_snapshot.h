// RigidBody2DState's binary snapshot (rigidbody2d/RigidBody2DState.cpp:485-556), written and parsed on the host from plain arrays.  Plain C++ (no CUDA):
// sg_rb2d.cu gathers the arrays from the device-resident state and calls these; the CPU suite compiles the same header and compares its bytes with what
// the reference's own RigidBody2DState::serialize writes (tests/test_rb2d_snapshot_cpu.py).
//
// Layout, in the order RigidBody2DState::serialize writes it (scisim/Utilities.h:43-94, Utilities.cpp:9-18, scisim/Math/MathUtilities.h:42-60,
// MathUtilities.cpp:142-154):
//   q, v                   Eigen::Index rows + 3N doubles ( x, y, theta per body )
//   M, Minv                sparse 3N x 3N diagonal ( rows, cols, nnz as Eigen::Index; nnz inner indices, cols + 1 outer indices as int; nnz doubles ):
//                          m, m, I per body; Minv holds 1.0 / each (generateMinv, RigidBody2DState.cpp:30-43)
//   fixed                  size_t count + one byte per body
//   geometry indices       Eigen::Index rows + unsigned per body
//   geometry               size_t count + per geometry: RigidBody2DGeometryType ( int: CIRCLE 0, BOX 1 ) then the radius / the two half widths
//                          (CircleGeometry.cpp:52-56, BoxGeometry.cpp:50-54)
//   forces                 size_t count + { RigidBody2DForceType ( int: NEAR_EARTH_GRAVITY 0 ), g (2 doubles) } (NearEarthGravityForce.cpp:55-59)
//   static planes          size_t count + { x, n, t, v (2 doubles each), omega } (RigidBody2DStaticPlane.cpp:93-101)
//   planar portals         size_t count + { plane A, plane B as above, v, bounds, dx } (rigidbody2d/PlanarPortal.cpp:330-338)
#ifndef SG_RB2D_SNAPSHOT_H
#define SG_RB2D_SNAPSHOT_H

#include "sg_rb3d_snapshot.h" // Sink, Source

namespace sg_snapshot
{

// everything a snapshot holds, as plain host arrays
struct Rb2dState
{
  uint32_t n = 0;
  std::vector<double> q, v, M;                      // 3 n each ( M: m, m, I )
  std::vector<uint8_t> fixed;                       // n
  std::vector<uint32_t> geo_of_body;                // n
  std::vector<uint32_t> geo_type;                   // 0 circle, 1 box
  std::vector<double> geo_r, geo_half;              // ngeo, 2 ngeo
  double g[2] = { 0.0, 0.0 };
  std::vector<double> plane_x, plane_n, plane_t;    // 2 each
  std::vector<double> portal_ax, portal_an, portal_at, portal_bx, portal_bn, portal_bt; // 2 each
  std::vector<double> portal_v, portal_bounds, portal_dx;                               // 1 each
};

inline void put_diagonal2d( Sink& out, const uint32_t n, const std::vector<double>& M, const bool inverse )
{
  const long long dofs = 3ll * n;
  out.val<long long>( dofs ); out.val<long long>( dofs ); out.val<long long>( dofs );
  for( long long k = 0; k < dofs; ++k ) { out.val<int>( int( k ) ); }
  for( long long k = 0; k <= dofs; ++k ) { out.val<int>( int( k ) ); }
  for( long long k = 0; k < dofs; ++k ) { out.val<double>( inverse ? 1.0 / M[size_t( k )] : M[size_t( k )] ); }
}

inline void put_plane2d( Sink& out, const double* x, const double* nrm, const double* tng )
{
  out.put( x, 16 ); out.put( nrm, 16 ); out.put( tng, 16 );
  for( int k = 0; k < 3; ++k ) { out.val<double>( 0.0 ); } // m_v, m_omega: planes of this path do not move
}

// returns false when the state cannot be written in the reference's format (a geometry that is neither circle nor box)
inline bool serialize( const Rb2dState& s, Sink& out )
{
  const uint32_t n = s.n;
  out.val<long long>( 3ll * n ); out.put( s.q.data(), uint64_t( 3 ) * n * 8 );
  out.val<long long>( 3ll * n ); out.put( s.v.data(), uint64_t( 3 ) * n * 8 );
  put_diagonal2d( out, n, s.M, false );
  put_diagonal2d( out, n, s.M, true );
  out.val<size_t>( size_t( n ) );
  for( uint32_t b = 0; b < n; ++b ) { out.val<unsigned char>( s.fixed[b] ? 1 : 0 ); }
  out.val<long long>( ( long long )( n ) );
  for( uint32_t b = 0; b < n; ++b ) { out.val<unsigned>( s.geo_of_body[b] ); }
  out.val<size_t>( s.geo_type.size() );
  for( size_t k = 0; k < s.geo_type.size(); ++k )
  {
    if( s.geo_type[k] == 0u ) { out.val<int>( 0 ); out.val<double>( s.geo_r[k] ); }
    else if( s.geo_type[k] == 1u ) { out.val<int>( 1 ); out.put( &s.geo_half[2 * k], 16 ); }
    else { return false; }
  }
  out.val<size_t>( size_t( 1 ) );
  out.val<int>( 0 ); // RigidBody2DForceType::NEAR_EARTH_GRAVITY
  out.put( s.g, 16 );
  const size_t npl = s.plane_x.size() / 2;
  out.val<size_t>( npl );
  for( size_t k = 0; k < npl; ++k ) { put_plane2d( out, &s.plane_x[2 * k], &s.plane_n[2 * k], &s.plane_t[2 * k] ); }
  const size_t npo = s.portal_v.size();
  out.val<size_t>( npo );
  for( size_t k = 0; k < npo; ++k )
  {
    put_plane2d( out, &s.portal_ax[2 * k], &s.portal_an[2 * k], &s.portal_at[2 * k] );
    put_plane2d( out, &s.portal_bx[2 * k], &s.portal_bn[2 * k], &s.portal_bt[2 * k] );
    out.val<double>( s.portal_v[k] ); out.val<double>( s.portal_bounds[k] ); out.val<double>( s.portal_dx[k] );
  }
  return true;
}

// x, n, t as stored; false where the plane moves ( m_v, m_omega: outside this path )
inline bool take_plane2d( Source& in, double* x, double* nrm, double* tng )
{
  in.doubles( x, 2 ); in.doubles( nrm, 2 ); in.doubles( tng, 2 );
  double rest[3] = { 0.0, 0.0, 0.0 };
  in.doubles( rest, 3 );
  return rest[0] == 0.0 && rest[1] == 0.0 && rest[2] == 0.0;
}

// 0 ok, 1 malformed / truncated, 2 holds something this path does not support (why says what)
inline int parse( Source& in, Rb2dState& s, const char** why )
{
  *why = "";
  const long long nq = in.val<long long>();
  if( !in.ok ) { *why = "truncated"; return 1; }
  if( nq < 0 || nq % 3 != 0 || nq / 3 >= 0x80000000ll ) { *why = "q does not hold 3 doubles per body"; return 1; }
  const uint32_t n = uint32_t( nq / 3 );
  s.n = n;
  if( uint64_t( nq ) * 8 > in.cap - in.n ) { *why = "truncated"; return 1; }
  s.q.resize( size_t( nq ) ); in.doubles( s.q.data(), uint64_t( nq ) );
  if( in.val<long long>() != nq ) { *why = "v of another length than q"; return 1; }
  if( uint64_t( nq ) * 8 > in.cap - in.n ) { *why = "truncated"; return 1; }
  s.v.resize( size_t( nq ) ); in.doubles( s.v.data(), uint64_t( nq ) );
  s.M.resize( size_t( nq ) );
  for( int mat = 0; mat < 2; ++mat )
  {
    const long long rows = in.val<long long>(), cols = in.val<long long>(), nnz = in.val<long long>();
    if( !in.ok || rows != nq || cols != nq || nnz != nq ) { *why = "a mass matrix that is not the 3N diagonal"; return 1; }
    in.take( uint64_t( nnz ) * 4 ); in.take( uint64_t( cols + 1 ) * 4 );
    if( !in.ok ) { *why = "truncated"; return 1; }
    if( mat == 0 ) { in.doubles( s.M.data(), uint64_t( nq ) ); }   // the flow divides by M's entries; Minv is written again as 1.0 / each
    else { in.take( uint64_t( nnz ) * 8 ); }
    if( !in.ok ) { *why = "truncated"; return 1; }
  }
  if( in.val<size_t>() != size_t( n ) ) { *why = "fixed flags of another length"; return 1; }
  s.fixed.resize( n );
  for( uint32_t b = 0; b < n; ++b ) { s.fixed[b] = in.val<unsigned char>(); }
  if( in.val<long long>() != ( long long )( n ) ) { *why = "geometry indices of another length"; return 1; }
  s.geo_of_body.resize( n );
  for( uint32_t b = 0; b < n; ++b ) { s.geo_of_body[b] = in.val<unsigned>(); }
  const size_t ngeo = in.val<size_t>();
  if( !in.ok || ngeo > ( 1u << 24 ) || ngeo > ( in.cap - in.n ) / 12 ) { *why = "bad geometry count"; return 1; } // a geometry takes at least a type int and a double
  s.geo_type.assign( ngeo, 0u ); s.geo_r.assign( ngeo, 0.0 ); s.geo_half.assign( 2 * ngeo, 0.0 );
  for( size_t k = 0; k < ngeo; ++k )
  {
    const int t = in.val<int>();
    if( !in.ok ) { *why = "truncated"; return 1; }
    if( t == 0 ) { s.geo_type[k] = 0u; s.geo_r[k] = in.val<double>(); }
    else if( t == 1 ) { s.geo_type[k] = 1u; in.doubles( &s.geo_half[2 * k], 2 ); }
    else { *why = "a geometry that is neither circle nor box"; return 1; }
  }
  for( uint32_t b = 0; b < n; ++b ) { if( s.geo_of_body[b] >= ngeo ) { *why = "geometry index out of range"; return 1; } }
  const size_t nf = in.val<size_t>();
  if( !in.ok || nf > 4096 ) { *why = "bad force count"; return 1; }
  s.g[0] = s.g[1] = 0.0;
  for( size_t k = 0; k < nf; ++k )
  {
    const int t = in.val<int>();
    if( !in.ok ) { *why = "truncated"; return 1; }
    if( t != 0 ) { *why = "a force other than near-earth gravity"; return 2; }
    double g[2] = { 0.0, 0.0 };
    in.doubles( g, 2 );
    for( int c = 0; c < 2; ++c ) { s.g[c] += g[c]; } // forces accumulate
  }
  const size_t npl = in.val<size_t>();
  if( !in.ok || npl > 4096 ) { *why = "bad plane count"; return 1; }
  s.plane_x.resize( 2 * npl ); s.plane_n.resize( 2 * npl ); s.plane_t.resize( 2 * npl );
  bool still = true;
  for( size_t k = 0; k < npl; ++k ) { still = take_plane2d( in, &s.plane_x[2 * k], &s.plane_n[2 * k], &s.plane_t[2 * k] ) && still; }
  const size_t npo = in.val<size_t>();
  if( !in.ok || npo > 4096 ) { *why = "bad portal count"; return 1; }
  s.portal_ax.resize( 2 * npo ); s.portal_an.resize( 2 * npo ); s.portal_at.resize( 2 * npo );
  s.portal_bx.resize( 2 * npo ); s.portal_bn.resize( 2 * npo ); s.portal_bt.resize( 2 * npo );
  s.portal_v.resize( npo ); s.portal_bounds.resize( npo ); s.portal_dx.resize( npo );
  for( size_t k = 0; k < npo; ++k )
  {
    still = take_plane2d( in, &s.portal_ax[2 * k], &s.portal_an[2 * k], &s.portal_at[2 * k] ) && still;
    still = take_plane2d( in, &s.portal_bx[2 * k], &s.portal_bn[2 * k], &s.portal_bt[2 * k] ) && still;
    s.portal_v[k] = in.val<double>(); s.portal_bounds[k] = in.val<double>(); s.portal_dx[k] = in.val<double>();
  }
  if( !in.ok ) { *why = "truncated"; return 1; }
  if( !still ) { *why = "a moving static plane (outside this path)"; return 2; }
  return 0;
}

}

#endif
