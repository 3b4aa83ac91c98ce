// example_ball2d_multi.cpp -- the step of example_ball2d.cpp on SEVERAL GPUs of one process: GpuBall2DMultiBackend (sg_multi) cuts the
// scene into x-slabs behind the same calls.  Bodies are numbered by a fixed pseudo-random permutation of the lattice, so slabs are
// not index ranges.  Prints one summary line with order-sensitive checksums; tests/test_host_shim.py compares it with the
// single-GPU Python path on the same scene.   usage: example_ball2d_multi nx ny n_slabs [n_devices]
#include "gpu_backend.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>

struct TinySystem final : public FlowableSystem
{
  int n;
  explicit TinySystem( const int nballs ) : n( nballs ) {}
  int nqdofs() const override { return 2 * n; }
  int nvdofs() const override { return 2 * n; }
  unsigned numVelDoFsPerBody() const override { return 2; }
  unsigned ambientSpaceDimensions() const override { return 2; }
  bool isKinematicallyScripted( const int ) const override { return false; }
  void computeForce( const VectorXs&, const VectorXs&, const scalar&, VectorXs& F ) override
  {
    F.setZero();
    for( int b = 0; b < n; ++b ) { F( 2 * b + 1 ) = 0.0 + 1.0 * -9.81; }
  }
  std::string name() const override { return "ball_2d"; }
};

int main( int argc, char** argv )
{
  const int nx = argc > 1 ? std::atoi( argv[1] ) : 40;
  const int ny = argc > 2 ? std::atoi( argv[2] ) : 30;
  const int slabs = argc > 3 ? std::atoi( argv[3] ) : 2;
  const int ndev = argc > 4 ? std::atoi( argv[4] ) : 1;
  const int n = nx * ny;
  // body b sits at lattice site perm[b]: perm = multiplication by a unit modulo n (a bijection), so numbering and space are unrelated
  auto gcd = []( long a, long b ) { while( b != 0 ) { const long t = a % b; a = b; b = t; } return a; };
  long mult = 7919;
  while( gcd( mult, long( n ) ) != 1 ) { ++mult; }
  VectorXs q0( 2 * n ), v0( 2 * n ), q1( 2 * n ), v1( 2 * n ), r( n ), m( n );
  for( int b = 0; b < n; ++b )
  {
    const int site = int( ( long( b ) * mult ) % n ), i = site % nx, j = site / nx;
    q0( 2 * b ) = 0.99 * i + 0.001 * std::sin( 12.9898 * site );
    q0( 2 * b + 1 ) = 0.99 * j + 0.001 * std::cos( 78.233 * site );
    v0( 2 * b ) = 0.0; v0( 2 * b + 1 ) = 0.0;
    r( b ) = 0.5; m( b ) = 1.0;
  }
  std::vector<int> devices( slabs );
  for( int k = 0; k < slabs; ++k ) { devices[k] = k % ndev; }
  GpuBall2DMultiBackend backend( devices );
  backend.setBodies( r, m );
  backend.setGravity( 0.0, -9.81 );
  backend.setPlanes( { 0.0, -0.5, -0.5, 0.0 }, { 0.0, 1.0, 1.0, 0.0 } );
  TinySystem fsys( n );
  GpuMultiSymplecticEulerMap umap( backend );
  umap.flow( q0, v0, fsys, 1, 1.0e-3, q1, v1 );
  std::vector<GpuContact2D> contacts;
  uint64_t ncand = 0;
  backend.computeActiveSet( q0, q1, contacts, &ncand, true );
  unsigned long nbb = 0, npl = 0;
  unsigned long long order_sum = 0;
  double dsum = 0.0;
  for( std::size_t k = 0; k < contacts.size(); ++k )
  {
    const GpuContact2D& c = contacts[k];
    if( c.type == SG_BALL_BALL ) { ++nbb; }
    if( c.type == SG_BALL_PLANE ) { ++npl; }
    order_sum = ( order_sum * 1000003ull + c.i * 31ull + c.j + c.type ) % 1000000007ull; // depends on the ORDER of the list
    dsum += c.depth;
  }
  std::printf( "n=%d gpus=%u partitions=%llu candidates=%llu ball_ball=%lu plane=%lu order_sum=%llu depth_sum=%.17g v1y=%.17g q1y0=%.17g\n", n, backend.numGpus(),
               ( unsigned long long ) backend.numPartitions(), ( unsigned long long ) ncand, nbb, npl, order_sum, dsum, v1( 1 ), q1( 1 ) );
  return 0;
}
