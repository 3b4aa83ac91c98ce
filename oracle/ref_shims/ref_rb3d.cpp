// oracle/ref_shims/ref_rb3d.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own, unmodified
//   rigidbody3d/SpatialGridDetector.cpp          (3-D AABB grid)
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp  (ODE-derived box-box: BoxBoxUtilities::isActive)
//   rigidbody3d/Geometry/RigidBodySphere.cpp, RigidBodyBox.cpp (+ RigidBodyGeometry.cpp)   (computeAABB)
//   rigidbody3d/StaticGeometry/StaticPlane.cpp, rigidbody3d/Portals/PlanarPortal.cpp         (plane frames, portal touch tests and teleports)
//   rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp, rigidbody3d/Constraints/MeshMeshUtilities.cpp (+ scisim/StringUtilities.cpp)
//                                                  (mesh AABB, detectCollision on the signed distance grid, mesh-mesh and mesh-half-plane sets)
// compiled from /root/reference against oracle/eigen_standin (oracle/Makefile.ref).
#include "rigidbody3d/SpatialGridDetector.h"
#include "rigidbody3d/Constraints/BoxBoxUtilities.h"
#include "rigidbody3d/Geometry/RigidBodySphere.h"
#include "rigidbody3d/Geometry/RigidBodyBox.h"
#include "rigidbody3d/Portals/PlanarPortal.h"
#include "rigidbody3d/Geometry/RigidBodyTriangleMesh.h"
#include "rigidbody3d/Constraints/MeshMeshUtilities.h"
#include "scisim/Math/MathUtilities.h"
#include "scisim/StringUtilities.h"
#include "scisim/Utilities.h"

#include <sstream>

#include <cstdint>

extern "C"
{

// boxes: n x [minx, miny, minz, maxx, maxy, maxz]
uint64_t ref_rb3d_overlaps( const uint32_t n, const double* boxes, const int all_pairs, uint32_t* ij, const uint64_t cap )
{
  std::vector<AABB> aabbs;
  std::vector<AABB>& v = aabbs;
  v.resize( n ); // the 3-D AABB has no (min, max) constructor: RigidBody3DSim fills min()/max() (RigidBody3DSim.cpp:1057-1069)
  for( uint32_t i = 0; i < n; ++i )
  {
    v[i].min() = Array3s{ boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2] };
    v[i].max() = Array3s{ boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5] };
  }
  std::set<std::pair<unsigned,unsigned>> overlaps;
  if( all_pairs ) { SpatialGridDetector::getPotentialOverlapsAllPairs( aabbs, overlaps ); }
  else { SpatialGridDetector::getPotentialOverlaps( aabbs, overlaps ); }
  uint64_t k = 0;
  for( const std::pair<unsigned,unsigned>& p : overlaps ) { if( k < cap ) { ij[2 * k] = p.first; ij[2 * k + 1] = p.second; } ++k; }
  return k;
}

// cm: 3 doubles, R: 9 doubles row-major, side: 3 doubles (full widths).  Returns the number of contact points (<= 8
// written to points, 3 doubles each); n = contact normal.
int ref_rb3d_box_box( const double* cm0, const double* R0, const double* side0, const double* cm1, const double* R1, const double* side1, double* n, double* points )
{
  Matrix33sr A, B;
  for( int i = 0; i < 3; ++i ) { for( int j = 0; j < 3; ++j ) { A( i, j ) = R0[3 * i + j]; B( i, j ) = R1[3 * i + j]; } }
  Vector3s nn;
  nn.setZero();
  std::vector<Vector3s> pts;
  BoxBoxUtilities::isActive( Vector3s{ cm0[0], cm0[1], cm0[2] }, A, Vector3s{ side0[0], side0[1], side0[2] }, Vector3s{ cm1[0], cm1[1], cm1[2] }, B, Vector3s{ side1[0], side1[1], side1[2] }, nn, pts );
  n[0] = nn.x(); n[1] = nn.y(); n[2] = nn.z();
  int k = 0;
  for( const Vector3s& p : pts ) { if( k < 8 ) { points[3 * k] = p.x(); points[3 * k + 1] = p.y(); points[3 * k + 2] = p.z(); } ++k; }
  return k;
}

// RigidBodySphere::computeAABB (type 1) / RigidBodyBox::computeAABB (type 0); R row-major; out = min(3), max(3)
void ref_rb3d_aabb( const int type, const double r, const double* half, const double* cm, const double* R, double* out )
{
  Matrix33sr Rm;
  for( int i = 0; i < 3; ++i ) { for( int j = 0; j < 3; ++j ) { Rm( i, j ) = R[3 * i + j]; } }
  Array3s mn, mx;
  if( type == 1 ) { const RigidBodySphere g{ r }; g.computeAABB( Vector3s{ cm[0], cm[1], cm[2] }, Rm, mn, mx ); }
  else { const RigidBodyBox g{ Vector3s{ half[0], half[1], half[2] } }; g.computeAABB( Vector3s{ cm[0], cm[1], cm[2] }, Rm, mn, mx ); }
  for( int k = 0; k < 3; ++k ) { out[k] = mn( k ); out[3 + k] = mx( k ); }
}

// StaticPlane( x, n ): out = n (3), t0 (3), t1 (3)
void ref_rb3d_plane_frame( const double* x, const double* n, double* out )
{
  const StaticPlane p{ Vector3s{ x[0], x[1], x[2] }, Vector3s{ n[0], n[1], n[2] } };
  const Vector3s nn{ p.n() }, t0{ p.t0() }, t1{ p.t1() };
  for( int k = 0; k < 3; ++k ) { out[k] = nn( k ); out[3 + k] = t0( k ); out[6 + k] = t1( k ); }
}
void* ref_rb3d_portal_create( const double* ax, const double* an, const double* bx, const double* bn, const int* mult )
{
  const StaticPlane a{ Vector3s{ ax[0], ax[1], ax[2] }, Vector3s{ an[0], an[1], an[2] } };
  const StaticPlane b{ Vector3s{ bx[0], bx[1], bx[2] }, Vector3s{ bn[0], bn[1], bn[2] } };
  return new PlanarPortal{ a, b, Array3i{ mult[0], mult[1], mult[2] } };
}
void ref_rb3d_portal_destroy( void* p ) { delete static_cast<PlanarPortal*>( p ); }
// box = min(3), max(3).  out = through A (3), through B (3), teleportPointInsidePortal (3).  Returns aabbTouchesPortal ( 0 no, 1 plane A,
// 2 plane B ) | 4 * pointInsidePortal
uint32_t ref_rb3d_portal_probe( const void* pv, const double* box, const double* x, double* out )
{
  const PlanarPortal& p = *static_cast<const PlanarPortal*>( pv );
  const Vector3s xin{ x[0], x[1], x[2] };
  Vector3s a, b, ti;
  p.teleportPointThroughPlaneA( xin, a );
  p.teleportPointThroughPlaneB( xin, b );
  p.teleportPointInsidePortal( xin, ti );
  for( int k = 0; k < 3; ++k ) { out[k] = a( k ); out[3 + k] = b( k ); out[6 + k] = ti( k ); }
  bool plane_idx = false;
  const bool touches = p.aabbTouchesPortal( Array3s{ box[0], box[1], box[2] }, Array3s{ box[3], box[4], box[5] }, plane_idx );
  return ( touches ? ( plane_idx ? 2u : 1u ) : 0u ) | ( p.pointInsidePortal( xin ) ? 4u : 0u );
}

// ---- RigidBodyTriangleMesh: built through its own stream constructor (RigidBodyTriangleMesh.cpp:109-124) from a stream this shim
// writes in the order that constructor reads (the file format on both ends is the override's, see ref_shims/override) --------------------
static Matrix3Xsc toMatrix( const uint32_t n, const double* xyz )
{
  Matrix3Xsc m;
  m.resize( 3, int( n ) );
  for( uint32_t k = 0; k < n; ++k ) { for( int c = 0; c < 3; ++c ) { m( c, int( k ) ) = xyz[3 * k + c]; } }
  return m;
}
void* ref_rb3d_mesh_create( const uint32_t nverts, const double* verts, const uint32_t nsamples, const double* samples, const uint32_t nhull, const double* hull,
                            const double* cell_delta, const uint32_t* dims, const double* origin, const double* sdf )
{
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  StringUtilities::serialize( std::string{ "shim" }, stm );
  MathUtilities::serialize( toMatrix( nverts, verts ), stm );
  Matrix3Xuc faces;
  faces.resize( 3, 0 );
  MathUtilities::serialize( faces, stm );
  Utilities::serialize( scalar( 1.0 ), stm );                          // volume
  MathUtilities::serialize( Vector3s{ 1.0, 1.0, 1.0 }, stm );          // I on rho
  MathUtilities::serialize( Vector3s{ 0.0, 0.0, 0.0 }, stm );          // centre of mass
  MathUtilities::serialize( Matrix3s{ Matrix3s::Identity() }, stm );   // R
  MathUtilities::serialize( toMatrix( nsamples, samples ), stm );
  MathUtilities::serialize( toMatrix( nhull, hull ), stm );
  const Vector3s delta{ cell_delta[0], cell_delta[1], cell_delta[2] };
  MathUtilities::serialize( delta, stm );
  Vector3u gd;
  gd( 0 ) = dims[0]; gd( 1 ) = dims[1]; gd( 2 ) = dims[2];
  MathUtilities::serialize( gd, stm );
  const Vector3s org{ origin[0], origin[1], origin[2] };
  MathUtilities::serialize( org, stm );
  VectorXs sd;
  sd.resize( int( dims[0] * dims[1] * dims[2] ) );
  for( int k = 0; k < sd.size(); ++k ) { sd( k ) = sdf[k]; }
  MathUtilities::serialize( sd, stm );
  // RigidBodyTriangleMesh.cpp:102: m_grid_end = m_grid_origin + ( ( m_grid_dimensions.array() - 1 ).cast<scalar>() * m_cell_delta.array() ).matrix()
  const Vector3s grid_end{ origin[0] + scalar( dims[0] - 1 ) * cell_delta[0], origin[1] + scalar( dims[1] - 1 ) * cell_delta[1], origin[2] + scalar( dims[2] - 1 ) * cell_delta[2] };
  MathUtilities::serialize( grid_end, stm );
  stm.seekg( 0 );
  return new RigidBodyTriangleMesh{ stm };
}
void ref_rb3d_mesh_destroy( void* m ) { delete static_cast<RigidBodyTriangleMesh*>( m ); }
static Matrix33sr toR( const double* R )
{
  Matrix33sr Rm;
  for( int i = 0; i < 3; ++i ) { for( int j = 0; j < 3; ++j ) { Rm( i, j ) = R[3 * i + j]; } }
  return Rm;
}
int ref_rb3d_mesh_detect( const void* m, const double* x, double* n_out )
{
  Vector3s n{ 0.0, 0.0, 0.0 };
  const bool hit = static_cast<const RigidBodyTriangleMesh*>( m )->detectCollision( Vector3s{ x[0], x[1], x[2] }, n );
  n_out[0] = n( 0 ); n_out[1] = n( 1 ); n_out[2] = n( 2 );
  return hit ? 1 : 0;
}
void ref_rb3d_mesh_aabb( const void* m, const double* cm, const double* R, double* out )
{
  Array3s mn, mx;
  static_cast<const RigidBodyTriangleMesh*>( m )->computeAABB( Vector3s{ cm[0], cm[1], cm[2] }, toR( R ), mn, mx );
  for( int k = 0; k < 3; ++k ) { out[k] = mn( k ); out[3 + k] = mx( k ); }
}
// MeshMeshUtilities::computeActiveSet: returns the number of contacts, writes up to cap points and normals
uint64_t ref_rb3d_mesh_mesh( const void* m0, const double* cm0, const double* R0, const void* m1, const double* cm1, const double* R1, double* p_out, double* n_out, const uint64_t cap )
{
  std::vector<Vector3s> p, n;
  MeshMeshUtilities::computeActiveSet( Vector3s{ cm0[0], cm0[1], cm0[2] }, toR( R0 ), *static_cast<const RigidBodyTriangleMesh*>( m0 ),
                                       Vector3s{ cm1[0], cm1[1], cm1[2] }, toR( R1 ), *static_cast<const RigidBodyTriangleMesh*>( m1 ), p, n );
  for( uint64_t k = 0; k < p.size() && k < cap; ++k ) { for( int c = 0; c < 3; ++c ) { p_out[3 * k + c] = p[k]( c ); n_out[3 * k + c] = n[k]( c ); } }
  return p.size();
}
// MeshMeshUtilities::computeMeshHalfPlaneActiveSet: convex-hull vertices with n.( R v + cm - x0 ) <= 0
uint64_t ref_rb3d_mesh_halfplane( const void* m, const double* cm, const double* R, const double* x0, const double* n, uint32_t* verts_out, const uint64_t cap )
{
  std::vector<unsigned> verts;
  MeshMeshUtilities::computeMeshHalfPlaneActiveSet( Vector3s{ cm[0], cm[1], cm[2] }, toR( R ), *static_cast<const RigidBodyTriangleMesh*>( m ), Vector3s{ x0[0], x0[1], x0[2] }, Vector3s{ n[0], n[1], n[2] }, verts );
  for( uint64_t k = 0; k < verts.size() && k < cap; ++k ) { verts_out[k] = verts[k]; }
  return verts.size();
}

}
