// oracle/ref_shims/ref_rb3d.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own, unmodified
//   rigidbody3d/SpatialGridDetector.cpp          (3-D AABB grid)
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp  (ODE-derived box-box: BoxBoxUtilities::isActive)
//   rigidbody3d/Geometry/RigidBodySphere.cpp, RigidBodyBox.cpp (+ RigidBodyGeometry.cpp)   (computeAABB)
//   rigidbody3d/StaticGeometry/StaticPlane.cpp, rigidbody3d/Portals/PlanarPortal.cpp         (plane frames, portal touch tests and teleports)
//   rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp, rigidbody3d/Constraints/MeshMeshUtilities.cpp (+ scisim/StringUtilities.cpp)
//                                                  (mesh AABB, detectCollision on the signed distance grid, mesh-mesh and mesh-half-plane sets)
//   rigidbody3d/UnconstrainedMaps/SplitHamMap.cpp, DMVMap.cpp, rigidbody3d/Forces/NearEarthGravityForce.cpp (+ Force.cpp, scisim/UnconstrainedMaps/*.cpp)
//                                                  (the two kick-drift-kick maps, driven through a shim FlowableSystem: see ShimRB3DSystem)
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
#include "rigidbody3d/UnconstrainedMaps/SplitHamMap.h"
#include "rigidbody3d/UnconstrainedMaps/DMVMap.h"
#include "rigidbody3d/UnconstrainedMaps/ExponentialEulerMap.h"
#include "rigidbody3d/Forces/NearEarthGravityForce.h"
#include "scisim/UnconstrainedMaps/FlowableSystem.h"
#include "rigidbody3d/Constraints/SphereSphereConstraint.h"
#include "rigidbody3d/Constraints/StaticPlaneSphereConstraint.h"
#include "rigidbody3d/Constraints/StaticPlaneBoxConstraint.h"
#include "rigidbody3d/Constraints/StaticCylinderSphereConstraint.h"
#include "rigidbody3d/Constraints/StaticCylinderBodyConstraint.h"
#include "rigidbody3d/Constraints/KinematicObjectSphereConstraint.h"
#include "rigidbody3d/StaticGeometry/StaticCylinder.h"
#include "rigidbody3d/RigidBody3DState.h"
#include "rigidbody3d/ConstraintCache.h"
#include <memory>
#include <cstring>

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
// RigidBodyTriangleMesh::serialize (rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp:215-232): the mesh's own record of a state snapshot -- what a caller hands to
// sg_rb3d_set_mesh_snapshot.  Returns its length; the bytes are written when they fit cap.
uint64_t ref_rb3d_mesh_serialize( const void* m, void* buf, const uint64_t cap )
{
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  static_cast<const RigidBodyTriangleMesh*>( m )->serialize( stm );
  const std::string bytes = stm.str();
  if( bytes.size() <= cap ) { std::memcpy( buf, bytes.data(), bytes.size() ); }
  return bytes.size();
}
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

// ---- the rigidbody3d maps: a FlowableSystem with what SplitHamMap / DMVMap ask of RigidBody3DSim -- M0, Minv0 (diagonal: m, m, m per body, then
// the body-frame inertias), M, Minv (the same 3N linear entries, then one 3x3 world-space block per body), the kinematic flags and
// computeForce = setZero + NearEarthGravityForce ( RigidBody3DSim.cpp:169-181 ).  The blocks are filled as the reference fills them:
//   first flow of a simulation   formWorldSpaceMassMatrix / ...InverseMassMatrix ( RigidBody3DState.cpp:140-240 ):
//                                I = R * I0.asDiagonal() * R^T, stored with insert( row = base + col_idx, col = base + row_idx ) = I( row_idx, col_idx )
//   later flows ( m_updated )    updateMandMinv ( RigidBody3DState.cpp:428-462 ): the same product assigned through a column-major map
// The products themselves are the reference's expressions, evaluated by the stand-in. -----------------------------------------------------------
namespace
{
class ShimRB3DSystem final : public FlowableSystem
{
public:
  ShimRB3DSystem( const uint32_t n, const double* q, const double* m, const double* I0, const uint8_t* fixed, const double* g, const bool m_updated )
  : m_n( n ), m_force( Vector3s{ g[0], g[1], g[2] } ), m_fixed( fixed, fixed + n )
  {
    std::vector<double> d0( 6 * size_t( n ) ), di0( 6 * size_t( n ) );
    for( uint32_t b = 0; b < n; ++b )
    {
      for( int k = 0; k < 3; ++k ) { d0[3 * b + k] = m[b]; di0[3 * b + k] = 1.0 / m[b]; d0[3 * n + 3 * b + k] = I0[3 * b + k]; di0[3 * n + 3 * b + k] = 1.0 / I0[3 * b + k]; }
    }
    m_M0.setDiagonal( d0.data(), int( d0.size() ) ); m_Minv0.setDiagonal( di0.data(), int( di0.size() ) );
    std::vector<int> outer( 6 * size_t( n ) + 1 ), inner( 12 * size_t( n ) );
    std::vector<double> vm( 12 * size_t( n ) ), vi( 12 * size_t( n ) );
    for( uint32_t c = 0; c < 3 * n; ++c ) { outer[c] = int( c ); inner[c] = int( c ); vm[c] = m[c / 3]; vi[c] = 1.0 / m[c / 3]; }
    for( uint32_t b = 0; b < n; ++b )
    {
      const Eigen::Map<const Matrix33sr> Rmat{ q + 3 * size_t( n ) + 9 * size_t( b ) };
      const Vector3s I0b{ I0[3 * b], I0[3 * b + 1], I0[3 * b + 2] };
      const Vector3s Iinv0b{ 1.0 / I0[3 * b], 1.0 / I0[3 * b + 1], 1.0 / I0[3 * b + 2] };
      const Matrix33sr I = Rmat * I0b.asDiagonal() * Rmat.transpose();
      const Matrix33sr Iinv = Rmat * Iinv0b.asDiagonal() * Rmat.transpose();
      for( int c = 0; c < 3; ++c )
      {
        const size_t col = 3 * size_t( n ) + 3 * size_t( b ) + c;
        outer[col] = int( 3 * n + 9 * b + 3 * c );
        for( int r = 0; r < 3; ++r )
        {
          inner[3 * n + 9 * b + 3 * c + r] = int( 3 * n + 3 * b + r );
          // constructor: entry ( row r, column c ) <- I( c, r );  updateMandMinv: column-major map, entry ( r, c ) <- I( r, c )
          vm[3 * n + 9 * b + 3 * c + r] = m_updated ? I( r, c ) : I( c, r );
          vi[3 * n + 9 * b + 3 * c + r] = m_updated ? Iinv( r, c ) : Iinv( c, r );
        }
      }
    }
    outer[6 * size_t( n )] = int( 12 * n );
    m_M.setCompressed( int( 6 * n ), outer, inner, vm ); m_Minv.setCompressed( int( 6 * n ), outer, inner, vi );
  }
  virtual int nqdofs() const override { return int( 12 * m_n ); }
  virtual int nvdofs() const override { return int( 6 * m_n ); }
  virtual unsigned numVelDoFsPerBody() const override { return 6; }
  virtual unsigned ambientSpaceDimensions() const override { return 3; }
  virtual bool isKinematicallyScripted( const int i ) const override { return m_fixed[size_t( i )] != 0; }
  virtual void computeForce( const VectorXs& q, const VectorXs& v, const scalar&, VectorXs& F ) override
  {
    F.setZero();
    m_force.computeForce( q, v, m_M, F );
  }
  virtual void zeroOutForcesOnFixedBodies( VectorXs& ) const override {}
  virtual void linearInertialConfigurationUpdate( const VectorXs&, const VectorXs&, const scalar&, VectorXs& ) const override {}
  virtual const SparseMatrixsc& M() const override { return m_M; }
  virtual const SparseMatrixsc& Minv() const override { return m_Minv; }
  virtual const SparseMatrixsc& M0() const override { return m_M0; }
  virtual const SparseMatrixsc& Minv0() const override { return m_Minv0; }
  virtual void computeMomentum( const VectorXs&, VectorXs& ) const override {}
  virtual void computeAngularMomentum( const VectorXs&, VectorXs& ) const override {}
  virtual std::string name() const override { return "shim_rigid_body_3d"; }
  const double* mValues() const { return m_M.valuePtr(); }
  const double* minvValues() const { return m_Minv.valuePtr(); }
private:
  uint32_t m_n;
  NearEarthGravityForce m_force;
  std::vector<uint8_t> m_fixed;
  SparseMatrixsc m_M0, m_Minv0, m_M, m_Minv;
};
}

extern "C"
{
// kind 2: SplitHamMap::flow, 3: DMVMap::flow, 4: ExponentialEulerMap::flow (the reference's own code; for kind 4 the stand-in's JacobiSVD is NOT
// Eigen's algorithm, see eigen_standin/Eigen/Geometry: the projected rotation is pinned to rounding only, everything else bit for bit); q: 12 n, v: 6 n
void ref_rb3d_flow( const int kind, const uint32_t n, const double* q0, const double* v0, const double* m, const double* I0, const uint8_t* fixed, const double* g, const double dt,
                    const int m_updated, double* q1, double* v1 )
{
  ShimRB3DSystem sys{ n, q0, m, I0, fixed, g, m_updated != 0 };
  VectorXs q0v( int( 12 * n ) ), v0v( int( 6 * n ) ), q1v( int( 12 * n ) ), v1v( int( 6 * n ) );
  for( uint32_t k = 0; k < 12 * n; ++k ) { q0v( int( k ) ) = q0[k]; }
  for( uint32_t k = 0; k < 6 * n; ++k ) { v0v( int( k ) ) = v0[k]; }
  if( kind == 2 ) { SplitHamMap map; map.flow( q0v, v0v, sys, 1, dt, q1v, v1v ); }
  else if( kind == 4 ) { ExponentialEulerMap map; map.flow( q0v, v0v, sys, 1, dt, q1v, v1v ); }
  else { DMVMap map; map.flow( q0v, v0v, sys, 1, dt, q1v, v1v ); }
  for( uint32_t k = 0; k < 12 * n; ++k ) { q1[k] = q1v( int( k ) ); }
  for( uint32_t k = 0; k < 6 * n; ++k ) { v1[k] = v1v( int( k ) ); }
}
// the 3x3 blocks of M and Minv as the shim builds them (column-major values, 9 per body) -- to compare with the oracle's updateMandMinv
void ref_rb3d_mass_blocks( const uint32_t n, const double* q, const double* m, const double* I0, const int m_updated, double* I_blocks, double* Iinv_blocks )
{
  std::vector<uint8_t> fixed( n, 0 );
  const double g[3] = { 0.0, 0.0, 0.0 };
  ShimRB3DSystem sys{ n, q, m, I0, fixed.data(), g, m_updated != 0 };
  for( size_t k = 0; k < 9 * size_t( n ); ++k ) { I_blocks[k] = sys.mValues()[3 * size_t( n ) + k]; Iinv_blocks[k] = sys.minvValues()[3 * size_t( n ) + k]; }
}
}


// ---- the constraint classes themselves (rigidbody3d/Constraints/{SphereSphere,StaticPlaneSphere,StaticPlaneBox}Constraint.cpp
// + scisim/Constraints/Constraint.cpp, compiled unchanged): isActive at q1, the constraint built as the sim builds it, its normal, world-space
// contact point at q0 and penetrationDepth at q1 ---------------------------------------------------------------------------------------------
extern "C"
{

// q0, q1: 12 n doubles ( 3n positions | 9n row-major rotations ).
// kind 0  sphere-sphere ( i, j ), geo = r_i, r_j.  n and p are formed as RigidBody3DSim::sphereSphereNarrowPhaseCollision forms them (RigidBody3DSim.cpp:803-808,
//         two lines restated here, marked) and handed to the class's constructor
// kind 1  plane-sphere: geo = x[3], n[3], r          kind 2  plane-box: geo = x[3], n[3], half[3], aux = corner number
// kind 3  cylinder-sphere: geo = x[3], axis[3] (as given to StaticCylinder's constructor, which normalises it), R, r_sphere; j = cylinder number
// kind 4  cylinder-body (mesh hull vertex): geo = x[3], axis[3], R, collision point at q0 [3] (RigidBody3DSim.cpp:1541, formed by the caller); no isActive
//         of its own (MeshMeshUtilities::computeMeshCylinderActiveSet decides): out[0] = 1
// kind 5  free sphere i against kinematic sphere j: geo = r_i, r_j.  n as RigidBody3DSim.cpp:803 forms it (restated here, marked); KinematicSphereSphereConstraint
// out[0] isActive at q1 (plane-box: the corner is among the active corners), out[1..3] normal, out[4..6] contact point at q0, out[7] penetrationDepth( q1 )
void ref_rb3d_constraint_probe( const int kind, const unsigned i, const unsigned j, const unsigned aux, const uint32_t nbodies, const double* q0, const double* q1, const double* geo, double* out )
{
  const int nq = 12 * int( nbodies );
  VectorXs wq0{ nq }, wq1{ nq };
  for( int k = 0; k < nq; ++k ) { wq0( k ) = q0[k]; wq1( k ) = q1[k]; }
  const VectorXs& vq0 = wq0; const VectorXs& vq1 = wq1; // const: segment<3>() reads
  const Vector3s gx{ geo[0], geo[1], geo[2] }, gn{ geo[3], geo[4], geo[5] };
  const StaticPlane plane{ gx, ( kind == 1 || kind == 2 ) ? gn : Vector3s{ 0.0, 1.0, 0.0 } };   // outlive the constraints, which keep references
  const StaticCylinder cyl{ gx, ( kind == 3 || kind == 4 ) ? gn : Vector3s{ 0.0, 1.0, 0.0 }, ( kind == 3 || kind == 4 ) ? geo[6] : 1.0 };
  std::unique_ptr<Constraint> con;
  bool active = false;
  if( kind == 0 )
  {
    const scalar r0 = geo[0], r1 = geo[1];
    active = SphereSphereConstraint::isActive( vq1.segment<3>( 3 * i ), vq1.segment<3>( 3 * j ), r0, r1 );
    // restated glue (RigidBody3DSim.cpp:803, 808):
    const Vector3s n{ ( vq0.segment<3>( 3 * i ) - vq0.segment<3>( 3 * j ) ).normalized() };
    const Vector3s p{ vq0.segment<3>( 3 * i ) + ( r0 / ( r0 + r1 ) ) * ( vq0.segment<3>( 3 * j ) - vq0.segment<3>( 3 * i ) ) };
    con.reset( new SphereSphereConstraint{ i, j, n, p, r0, r1 } );
  }
  else if( kind == 1 )
  {
    active = StaticPlaneSphereConstraint::isActive( plane.x(), plane.n(), vq1.segment<3>( 3 * i ), geo[6] );
    con.reset( new StaticPlaneSphereConstraint{ i, geo[6], plane, j } );
  }
  else if( kind == 3 )
  {
    active = StaticCylinderSphereConstraint::isActive( cyl.x(), cyl.axis(), cyl.r(), vq1.segment<3>( 3 * i ), geo[7] );
    con.reset( new StaticCylinderSphereConstraint{ i, geo[7], cyl, j } );
  }
  else if( kind == 4 )
  {
    active = true;
    con.reset( new StaticCylinderBodyConstraint{ i, Vector3s{ geo[7], geo[8], geo[9] }, cyl, j, vq0 } );
  }
  else if( kind == 5 )
  {
    const scalar r0 = geo[0], r1 = geo[1];
    active = SphereSphereConstraint::isActive( vq1.segment<3>( 3 * i ), vq1.segment<3>( 3 * j ), r0, r1 );
    // restated glue (RigidBody3DSim.cpp:803):
    const Vector3s n{ ( vq0.segment<3>( 3 * i ) - vq0.segment<3>( 3 * j ) ).normalized() };
    con.reset( new KinematicSphereSphereConstraint{ i, r0, n, j, vq0.segment<3>( 3 * j ), Vector3s::Zero(), Vector3s::Zero(), r1 } );
  }
  else
  {
    const Vector3s half{ geo[6], geo[7], geo[8] };
    const Matrix33sr R{ Eigen::Map<const Matrix33sr>( vq1.segment<9>( 3 * int( nbodies ) + 9 * i ).data() ) };
    std::vector<short> corners;
    StaticPlaneBoxConstraint::isActive( plane.x(), plane.n(), vq1.segment<3>( 3 * i ), R, half, corners );
    for( const short c : corners ) { if( unsigned( c ) == aux ) { active = true; } }
    con.reset( new StaticPlaneBoxConstraint{ i, short( aux ), plane.n(), half, vq0, j } );
  }
  out[0] = active ? 1.0 : 0.0;
  if( kind == 4 )
  {
    // the class implements neither getWorldSpaceContactNormal nor ...Point (the base class exits): its normal is column 0 of its contact basis
    VectorXs vzero{ 6 * int( nbodies ) };
    vzero.setZero();
    MatrixXXsc basis;
    con->computeBasis( vq0, vzero, basis ); // the public wrapper of computeContactBasis (scisim/Constraints/Constraint.cpp:10-16)
    for( int k = 0; k < 3; ++k ) { out[1 + k] = basis( k, 0 ); out[4 + k] = geo[7 + k]; }
    out[7] = con->penetrationDepth( vq1 );
    return;
  }
  VectorXs n, p;
  con->getWorldSpaceContactNormal( vq0, n );
  con->getWorldSpaceContactPoint( vq0, p );
  for( int k = 0; k < 3; ++k ) { out[1 + k] = n( k ); out[4 + k] = p( k ); }
  out[7] = con->penetrationDepth( vq1 );
}

}


// ---- RigidBody3DState::updateMandMinv (rigidbody3d/RigidBody3DState.cpp:428-462) as EXPRESSIONS: the two assignments of its loop body, typed with the
// same Eigen types and Maps, evaluated by the stand-in.  Kept from before the file itself compiled (ref_rb3d_state_mass_matrices below now drives the real
// RigidBody3DState); a per-body probe that needs no state object.
extern "C"
{
void ref_rb3d_update_inertia_expr( const double* R_rowmajor9, const double* I0_3, const double* Iinv0_3, double* I_colmajor9, double* Iinv_colmajor9 )
{
  const Eigen::Map<const Matrix33sr> R{ R_rowmajor9 };
  {
    Eigen::Map<Matrix33sc> I{ I_colmajor9 };
    const Eigen::Map<const Vector3s> I0{ I0_3 };
    I = R * I0.asDiagonal() * R.transpose();      // RigidBody3DState.cpp:446
  }
  {
    Eigen::Map<Matrix33sc> Iinv{ Iinv_colmajor9 };
    const Eigen::Map<const Vector3s> Iinv0{ Iinv0_3 };
    Iinv = R * Iinv0.asDiagonal() * R.transpose(); // RigidBody3DState.cpp:455
  }
}
}


// ---- RigidBody3DState itself (rigidbody3d/RigidBody3DState.cpp, compiled unchanged): the mass matrices its setState builds
// ( formWorldSpaceMassMatrix / ...InverseMassMatrix, :140-240 ) and what updateMandMinv ( :428-462 ) leaves in them after the orientations changed.
// q_ctor, q_upd: 12 n doubles; all bodies are unit spheres as far as the geometry list goes (the matrices do not look at it).
// Outputs: the 9 n values behind the 3 n linear entries of M and Minv, as stored (column-compressed), after construction and after the update.
extern "C"
{
void ref_rb3d_state_mass_matrices( const uint32_t n, const double* q_ctor, const double* q_upd, const double* m, const double* I0,
                                   double* M_ctor, double* Minv_ctor, double* M_upd, double* Minv_upd )
{
  std::vector<Vector3s> X( n ), V( n, Vector3s::Zero() ), omega( n, Vector3s::Zero() ), I0v( n );
  std::vector<scalar> M( n );
  std::vector<VectorXs> R( n );
  std::vector<bool> fixed( n, false );
  std::vector<unsigned> geo_idx( n, 0u );
  std::vector<std::unique_ptr<RigidBodyGeometry>> geometry;
  geometry.emplace_back( new RigidBodySphere{ 1.0 } );
  for( uint32_t b = 0; b < n; ++b )
  {
    X[b] = Vector3s{ q_ctor[3 * b], q_ctor[3 * b + 1], q_ctor[3 * b + 2] };
    M[b] = m[b];
    I0v[b] = Vector3s{ I0[3 * b], I0[3 * b + 1], I0[3 * b + 2] };
    R[b].resize( 9 );
    for( int k = 0; k < 9; ++k ) { R[b]( k ) = q_ctor[3 * size_t( n ) + 9 * size_t( b ) + k]; }
  }
  RigidBody3DState state;
  state.setState( X, V, M, R, omega, I0v, fixed, geo_idx, geometry );
  const size_t off = 3 * size_t( n );
  for( size_t k = 0; k < 9 * size_t( n ); ++k ) { M_ctor[k] = state.M().valuePtr()[off + k]; Minv_ctor[k] = state.Minv().valuePtr()[off + k]; }
  for( size_t k = 0; k < 12 * size_t( n ); ++k ) { state.q()( int( k ) ) = q_upd[k]; }
  state.updateMandMinv();
  for( size_t k = 0; k < 9 * size_t( n ); ++k ) { M_upd[k] = state.M().valuePtr()[off + k]; Minv_upd[k] = state.Minv().valuePtr()[off + k]; }
}
}

// ---- rigidbody3d/ConstraintCache.cpp compiled unchanged: cacheConstraint for a list of constraints, then getCachedConstraint for another list.
// Constraints are built with the reference's own classes (only their indices and names matter to the cache).  type = the contact type codes of
// include/scisim_b200.h; a = first body, b = second body / static object.  r: ncomp doubles per constraint.  Returns constraintCacheEmpty() after the stores.
// (sphere constraints only: 10 sphere-sphere, 14 plane-sphere, 17 cylinder-sphere, 11 kinematic sphere-sphere -- anything else exits, as in the reference)
extern "C"
{
int ref_rb3d_cache_roundtrip_ex( const uint32_t nstore, const uint32_t* stype, const uint32_t* sa, const uint32_t* sb, const uint32_t ncomp, const double* rstore,
                              const uint32_t nquery, const uint32_t* qtype, const uint32_t* qa, const uint32_t* qb, double* rout,
                                 void* ser_out, const uint64_t ser_cap, uint64_t* ser_bytes, const void* deser_in, const uint64_t deser_bytes )
{
  const StaticPlane plane{ Vector3s{ 0.0, 0.0, 0.0 }, Vector3s{ 0.0, 1.0, 0.0 } };
  const StaticCylinder cyl{ Vector3s{ 0.0, 0.0, 0.0 }, Vector3s{ 0.0, 1.0, 0.0 }, 10.0 };
  const Vector3s n{ 1.0, 0.0, 0.0 }, p{ 0.0, 0.0, 0.0 };
  const auto make = [&]( const uint32_t type, const uint32_t a, const uint32_t b ) -> std::unique_ptr<Constraint>
  {
    if( type == 10 ) { return std::unique_ptr<Constraint>{ new SphereSphereConstraint{ a, b, n, p, 0.5, 0.5 } }; }
    if( type == 14 ) { return std::unique_ptr<Constraint>{ new StaticPlaneSphereConstraint{ a, 0.5, plane, b } }; }
    if( type == 17 ) { return std::unique_ptr<Constraint>{ new StaticCylinderSphereConstraint{ a, 0.5, cyl, b } }; }
    return std::unique_ptr<Constraint>{ new KinematicSphereSphereConstraint{ a, 0.5, n, b, p, Vector3s::Zero(), Vector3s::Zero(), 0.5 } };
  };
  ConstraintCache cache;
  VectorXs r{ int( ncomp ) };
  for( uint32_t k = 0; k < nstore; ++k )
  {
    for( uint32_t c = 0; c < ncomp; ++c ) { r( int( c ) ) = rstore[size_t( k ) * ncomp + c]; }
    cache.cacheConstraint( *make( stype[k], sa[k], sb[k] ), r );
  }
  const int empty = cache.empty() ? 1 : 0;
  // optionally: ConstraintCache::serialize of what was stored, handed to the caller, and / or the queries answered by a second cache that
  // ConstraintCache::deserialize filled from the caller's bytes
  if( ser_bytes != nullptr )
  {
    std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
    cache.serialize( stm );
    const std::string bytes = stm.str();
    *ser_bytes = bytes.size();
    if( ser_out != nullptr && bytes.size() <= ser_cap ) { std::memcpy( ser_out, bytes.data(), bytes.size() ); }
  }
  ConstraintCache restored;
  if( deser_in != nullptr )
  {
    std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
    stm.write( static_cast<const char*>( deser_in ), std::streamsize( deser_bytes ) );
    restored.deserialize( stm );
  }
  ConstraintCache& qcache = ( deser_in != nullptr ) ? restored : cache;
  for( uint32_t k = 0; k < nquery; ++k )
  {
    for( uint32_t c = 0; c < ncomp; ++c ) { r( int( c ) ) = -7.0; }
    qcache.getCachedConstraint( *make( qtype[k], qa[k], qb[k] ), r );
    for( uint32_t c = 0; c < ncomp; ++c ) { rout[size_t( k ) * ncomp + c] = r( int( c ) ); }
  }
  return empty;
}
int ref_rb3d_cache_roundtrip( const uint32_t nstore, const uint32_t* stype, const uint32_t* sa, const uint32_t* sb, const uint32_t ncomp, const double* rstore,
                              const uint32_t nquery, const uint32_t* qtype, const uint32_t* qa, const uint32_t* qb, double* rout )
{
  return ref_rb3d_cache_roundtrip_ex( nstore, stype, sa, sb, ncomp, rstore, nquery, qtype, qa, qb, rout, nullptr, 0, nullptr, nullptr, 0 );
}
}
