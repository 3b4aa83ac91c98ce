// example_ball2d.cpp -- drives the host shim the way SCISim's maps drive a sim for one step
// (scisim/ConstrainedMaps/ImpactMaps/ImpactMap.cpp:54-58): umap.flow( q0, v0 ) then computeActiveSet( q0, q1 ).
// Prints one summary line; tests/test_host_shim.py compares it with the Python path on the same scene.
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
  // Ball2DSim::computeForce with the scene's one Ball2DGravityForce (unit masses, g = -9.81 y); with penalty != 0 a second force is
  // mixed in, which the GPU maps must refuse (GravityOnlyGuard)
  double penalty = 0.0;
  void computeForce( const VectorXs& q, const VectorXs&, const scalar&, VectorXs& F ) override
  {
    F.setZero();
    for( int b = 0; b < n; ++b ) { F( 2 * b + 1 ) = 0.0 + 1.0 * -9.81; F( 2 * b ) += penalty * q( 2 * b ); }
  }
  std::string name() const override { return "ball_2d"; }
};

int main( int argc, char** argv )
{
  const int nx = argc > 1 ? std::atoi( argv[1] ) : 40;
  const int ny = argc > 2 ? std::atoi( argv[2] ) : 30;
  const int n = nx * ny;
  VectorXs q0( 2 * n ), v0( 2 * n ), q1( 2 * n ), v1( 2 * n ), r( n ), m( n );
  for( int j = 0; j < ny; ++j )
  {
    for( int i = 0; i < nx; ++i )
    {
      const int b = j * nx + i;
      q0( 2 * b ) = 0.99 * i + 0.001 * std::sin( 12.9898 * b );
      q0( 2 * b + 1 ) = 0.99 * j + 0.001 * std::cos( 78.233 * b );
      v0( 2 * b ) = 0.0; v0( 2 * b + 1 ) = 0.0;
      r( b ) = 0.5; m( b ) = 1.0;
    }
  }
  GpuBall2DBackend backend( 0 );
  backend.setBodies( r, m );
  backend.setGravity( 0.0, -9.81 );
  backend.setPlanes( { 0.0, -0.5, -0.5, 0.0 }, { 0.0, 1.0, 1.0, 0.0 } );
  TinySystem fsys( n );
  if( argc > 3 ) { fsys.penalty = std::atof( argv[3] ); } // a scene with a second force: the map must exit instead of integrating gravity alone
  GpuSymplecticEulerMap umap( backend );
  umap.flow( q0, v0, fsys, 1, 1.0e-3, q1, v1 );
  std::vector<GpuContact2D> contacts;
  uint64_t ncand = 0;
  backend.computeActiveSet( q0, q1, contacts, &ncand );
  unsigned long nbb = 0, npl = 0;
  double nsum = 0.0;
  PairImpulseCache cache;
  VectorXs imp( 1 );
  for( const GpuContact2D& c : contacts )
  {
    if( c.type == SG_BALL_BALL ) { ++nbb; imp( 0 ) = double( c.i ) + 0.5; cache.cacheConstraint( 0, c.i, c.j, imp ); }
    if( c.type == SG_BALL_PLANE ) { ++npl; imp( 0 ) = -1.0; cache.cacheConstraint( 1, c.j, c.i, imp ); }
    nsum += c.n[0] * c.n[0] + c.n[1] * c.n[1];
  }
  // warm-start lookups: every cached pair is found, an absent one reads zero
  unsigned long hits = 0;
  for( const GpuContact2D& c : contacts )
  {
    if( c.type != SG_BALL_BALL ) { continue; }
    cache.getCachedConstraint( 0, c.i, c.j, imp );
    if( imp( 0 ) == double( c.i ) + 0.5 ) { ++hits; }
  }
  cache.getCachedConstraint( 0, 4000000000u, 7u, imp );
  // the same on the device: Q = N^T Minv N of this active set, and the impulse cache as a join over the whole set
  sg_assembly as;
  backend.assemble( SG_ASM_N | SG_ASM_Q | SG_ASM_BASES, as );
  double qtrace = 0.0;
  for( uint64_t c = 0; c < as.n_constraints; ++c ) { for( int32_t e = as.q_outer[c]; e < as.q_outer[c + 1]; ++e ) { if( uint64_t( as.q_inner[e] ) == c ) { qtrace += as.q_values[e]; } } }
  VectorXs lam( static_cast<long>( contacts.size() ) ), warm( static_cast<long>( contacts.size() ) );
  for( std::size_t k = 0; k < contacts.size(); ++k ) { lam( static_cast<long>( k ) ) = 1.0 + double( k ); }
  backend.cacheStore( 1, lam );
  const uint64_t dev_hits = backend.cacheLookup( 1, warm );
  bool warm_ok = dev_hits == contacts.size();
  for( std::size_t k = 0; k < contacts.size(); ++k ) { warm_ok = warm_ok && warm( static_cast<long>( k ) ) == lam( static_cast<long>( k ) ); }
  std::printf( "n=%d candidates=%llu ball_ball=%lu plane=%lu cache_hits=%lu miss_value=%g mean_n2=%.17g v1y=%.17g q1y0=%.17g q_nnz=%llu q_trace=%.17g dev_cache_ok=%d\n", n, ( unsigned long long ) ncand, nbb, npl, hits, imp( 0 ),
               contacts.empty() ? 0.0 : nsum / double( contacts.size() ), v1( 1 ), q1( 1 ), ( unsigned long long ) as.q_nnz, qtrace, warm_ok ? 1 : 0 );
  return 0;
}
