// example_rigidbody.cpp -- drives the rigidbody3d and rigidbody2d host shims for one step each, the way SCISim's maps
// drive those sims (umap.flow( q0, v0 ) then computeActiveSet( q0, q1 )).  Prints one summary line per sim;
// tests/test_host_shim.py compares them with the Python path on the same scenes.
#include "gpu_backend.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>

struct TinySystem final : public FlowableSystem
{
  int nq, nv, dim;
  TinySystem( const int q, const int v, const int d ) : nq( q ), nv( v ), dim( d ) {}
  int nqdofs() const override { return nq; }
  int nvdofs() const override { return nv; }
  unsigned numVelDoFsPerBody() const override { return dim == 3 ? 6 : 3; }
  unsigned ambientSpaceDimensions() const override { return unsigned( dim ); }
  bool isKinematicallyScripted( const int ) const override { return false; }
  // the sims' computeForce with their one NearEarthGravityForce: F = 0 + M g on the translational dofs (unit masses, g = -9.81 y)
  void computeForce( const VectorXs&, const VectorXs&, const scalar&, VectorXs& F ) override
  {
    F.setZero();
    const int n = dim == 3 ? nv / 6 : nv / 3;
    for( int b = 0; b < n; ++b ) { F( 3 * b + 1 ) = 0.0 + 1.0 * -9.81; }
  }
  std::string name() const override { return dim == 3 ? "rigid_body_3d" : "rigid_body_2d"; }
};

// nx*ny*nz spheres (r = 0.5) on a 0.99 lattice followed by as many unit boxes on a 0.95 lattice 100 units away, all with
// R = identity; gravity -y; one floor plane.  DMV map.
static void run_rb3d( const int nx, const int ny, const int nz )
{
  const int ns = nx * ny * nz, n = 2 * ns;
  VectorXs q0( 12 * n ), v0( 6 * n ), q1, v1, m( n ), I0( 3 * n );
  for( int b = 0; b < n; ++b )
  {
    const int k = b % ns, ix = k % nx, iy = ( k / nx ) % ny, iz = k / ( nx * ny );
    const bool box = b >= ns;
    const double s = box ? 0.95 : 0.99, off = box ? 100.0 : 0.0;
    q0( 3 * b ) = off + s * ix + 0.001 * std::sin( 12.9898 * b );
    q0( 3 * b + 1 ) = s * iy + 0.001 * std::cos( 78.233 * b );
    q0( 3 * b + 2 ) = s * iz + 0.001 * std::sin( 37.719 * b );
    for( int e = 0; e < 9; ++e ) { q0( 3 * n + 9 * b + e ) = ( e % 4 == 0 ) ? 1.0 : 0.0; }
    for( int e = 0; e < 3; ++e ) { v0( 3 * b + e ) = 0.0; v0( 3 * n + 3 * b + e ) = 0.0; I0( 3 * b + e ) = 0.1; }
    m( b ) = 1.0;
  }
  GpuRigidBody3DBackend backend( 0 );
  backend.setGeometry( { SG_GEO_SPHERE, SG_GEO_BOX }, { 0.5, 0.0 }, { 0.0, 0.0, 0.0, 0.5, 0.5, 0.5 }, { 0u, 0u } );
  std::vector<uint32_t> geo( n );
  for( int b = 0; b < n; ++b ) { geo[b] = b >= ns ? 1u : 0u; }
  backend.setBodies( geo, std::vector<uint8_t>( n, 0 ), m, I0 );
  backend.setGravity( 0.0, -9.81, 0.0 );
  backend.setPlanes( { 0.0, -0.5, 0.0 }, { 0.0, 1.0, 0.0 } );
  TinySystem fsys( 12 * n, 6 * n, 3 );
  GpuDMVMap umap( backend );
  umap.flow( q0, v0, fsys, 1, 1.0e-3, q1, v1 );
  std::vector<GpuContact3D> contacts;
  uint64_t ncand = 0;
  backend.computeActiveSet( q0, q1, contacts, &ncand );
  unsigned long nss = 0, nbb = 0, nps = 0, npb = 0;
  double psum = 0.0;
  for( const GpuContact3D& c : contacts )
  {
    if( c.type == SG_SPHERE_SPHERE ) { ++nss; }
    if( c.type == SG_BODY_BODY ) { ++nbb; }
    if( c.type == SG_PLANE_SPHERE ) { ++nps; }
    if( c.type == SG_PLANE_BOX ) { ++npb; }
    psum += c.p[0] + c.p[1] + c.p[2];
  }
  std::printf( "rb3d n=%d candidates=%llu sphere_sphere=%lu body_body=%lu plane_sphere=%lu plane_box=%lu psum=%.17g v1y=%.17g q1y0=%.17g\n", n, ( unsigned long long ) ncand, nss, nbb, nps, npb, psum,
               v1( 1 ), q1( 1 ) );
}

// nx*ny circles (r = 0.5) on a 0.99 lattice followed by as many boxes (half-widths 0.5 x 0.4, rotated by 0.1*b) 100 units
// away; gravity -y; one floor plane.  Symplectic Euler.
static void run_rb2d( const int nx, const int ny )
{
  const int ns = nx * ny, n = 2 * ns;
  VectorXs q0( 3 * n ), v0( 3 * n ), q1, v1, M( 3 * n );
  for( int b = 0; b < n; ++b )
  {
    const int k = b % ns, ix = k % nx, iy = k / nx;
    const bool box = b >= ns;
    q0( 3 * b ) = ( box ? 100.0 : 0.0 ) + 0.99 * ix + 0.001 * std::sin( 12.9898 * b );
    q0( 3 * b + 1 ) = 0.99 * iy + 0.001 * std::cos( 78.233 * b );
    q0( 3 * b + 2 ) = box ? 0.1 * k : 0.0;
    v0( 3 * b ) = 0.0; v0( 3 * b + 1 ) = 0.0; v0( 3 * b + 2 ) = 0.0;
    M( 3 * b ) = 1.0; M( 3 * b + 1 ) = 1.0; M( 3 * b + 2 ) = 0.2;
  }
  GpuRigidBody2DBackend backend( 0 );
  backend.setGeometry( { SG_GEO2D_CIRCLE, SG_GEO2D_BOX }, { 0.5, 0.0 }, { 0.0, 0.0, 0.5, 0.4 } );
  std::vector<uint32_t> geo( n );
  for( int b = 0; b < n; ++b ) { geo[b] = b >= ns ? 1u : 0u; }
  backend.setBodies( geo, std::vector<uint8_t>( n, 0 ), M );
  backend.setGravity( 0.0, -9.81 );
  backend.setPlanes( { 0.0, -0.5 }, { 0.0, 1.0 } );
  TinySystem fsys( 3 * n, 3 * n, 2 );
  GpuRB2DSymplecticEulerMap umap( backend );
  umap.flow( q0, v0, fsys, 1, 1.0e-3, q1, v1 );
  std::vector<GpuContact2D> contacts;
  uint64_t ncand = 0;
  backend.computeActiveSet( q0, q1, contacts, &ncand );
  unsigned long ncc = 0, nbb = 0, npc = 0, npb = 0;
  double psum = 0.0;
  for( const GpuContact2D& c : contacts )
  {
    if( c.type == SG_CIRCLE_CIRCLE ) { ++ncc; }
    if( c.type == SG_BODY_BODY_2D ) { ++nbb; }
    if( c.type == SG_PLANE_CIRCLE ) { ++npc; }
    if( c.type == SG_PLANE_BODY_2D ) { ++npb; }
    psum += c.p[0] + c.p[1];
  }
  std::printf( "rb2d n=%d candidates=%llu circle_circle=%lu body_body=%lu plane_circle=%lu plane_body=%lu psum=%.17g v1y=%.17g q1y0=%.17g\n", n, ( unsigned long long ) ncand, ncc, nbb, npc, npb, psum,
               v1( 1 ), q1( 1 ) );
}

int main( int argc, char** argv )
{
  const int nx = argc > 1 ? std::atoi( argv[1] ) : 8;
  const int ny = argc > 2 ? std::atoi( argv[2] ) : 6;
  const int nz = argc > 3 ? std::atoi( argv[3] ) : 5;
  run_rb3d( nx, ny, nz );
  run_rb2d( nx * 3, ny * 3 );
  return 0;
}
