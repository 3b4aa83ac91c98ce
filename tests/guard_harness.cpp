// tests/guard_harness.cpp -- TEST INFRASTRUCTURE: the host shim's GravityOnlyGuard (scisim_b200/host/gpu_backend.h) re-configured from a state snapshot, as the
// three deserializeState wrappers do after the library has accepted the stream, then verified against a system whose force is 0 + m g with the masses and the
// gravity handed in separately.  No GPU involved (the guard is host logic).  Usage: guard_harness <layout 0|1|2> <snapshot file> <file: n masses, then g>
// Prints "guard ok n=<bodies>"; the guard itself prints and exits( 1 ) where the forces are not the snapshot's gravity.
#include "../scisim_b200/host/gpu_backend.h"

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <sstream>
#include <string>
#include <vector>

struct ForceOnly final : public FlowableSystem
{
  int layout;
  std::vector<double> m, g;
  int n() const { return int( m.size() ); }
  int nqdofs() const override { return layout == 0 ? 2 * n() : layout == 1 ? 3 * n() : 12 * n(); }
  int nvdofs() const override { return layout == 0 ? 2 * n() : layout == 1 ? 3 * n() : 6 * n(); }
  unsigned numVelDoFsPerBody() const override { return layout == 0 ? 2 : layout == 1 ? 3 : 6; }
  unsigned ambientSpaceDimensions() const override { return layout == 2 ? 3 : 2; }
  bool isKinematicallyScripted( const int ) const override { return false; }
  void computeForce( const VectorXs&, const VectorXs&, const scalar&, VectorXs& F ) override
  {
    F.setZero();
    const int per_body = layout == 0 ? 2 : 3, dim = layout == 2 ? 3 : 2;
    for( int b = 0; b < n(); ++b ) { for( int k = 0; k < dim; ++k ) { F( per_body * b + k ) = 0.0 + m[std::size_t( b )] * g[std::size_t( k )]; } }
  }
  std::string name() const override { return "force_only"; }
};

static std::vector<char> slurp( const char* path )
{
  std::ifstream f( path, std::ios::binary );
  return std::vector<char>( ( std::istreambuf_iterator<char>( f ) ), std::istreambuf_iterator<char>() );
}

// guard_harness split <layout> <file: a <Sim>::serialize stream = state, then the constraint cache> <prefix bytes to skip first>: the stream handling of the
// deserializeState wrappers (sghReadRest, the guard's parser, sghRewindBehindState); prints how many bytes are left behind the state and their sum
static int split( const int layout, const char* path, const int prefix )
{
  const std::vector<char> all = slurp( path );
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  stm.write( all.data(), std::streamsize( all.size() ) );
  stm.seekg( prefix );
  std::istream::pos_type start;
  const std::vector<char> rest = sghReadRest( stm, start );
  GravityOnlyGuard guard;
  const GravityOnlyGuard::Layout lay = layout == 0 ? GravityOnlyGuard::BALL2D : layout == 1 ? GravityOnlyGuard::RIGIDBODY2D : GravityOnlyGuard::RIGIDBODY3D;
  std::size_t consumed = 0;
  guard.configureFromSnapshot( lay, rest.data(), rest.size(), "guard_harness", &consumed );
  sghRewindBehindState( stm, start, consumed );
  const std::vector<char> behind( ( std::istreambuf_iterator<char>( stm ) ), std::istreambuf_iterator<char>() );
  unsigned long sum = 0;
  for( const char c : behind ) { sum += static_cast<unsigned char>( c ); }
  std::printf( "behind=%zu sum=%lu consumed=%zu\n", behind.size(), sum, consumed );
  return 0;
}

int main( int argc, char** argv )
{
  if( argc >= 5 && std::string( argv[1] ) == "split" ) { return split( std::atoi( argv[2] ), argv[3], std::atoi( argv[4] ) ); }
  if( argc < 4 ) { return 2; }
  const int layout = std::atoi( argv[1] );
  const std::vector<char> blob = slurp( argv[2] ), mg = slurp( argv[3] );
  const int dim = layout == 2 ? 3 : 2;
  ForceOnly fsys;
  fsys.layout = layout;
  const std::size_t nvals = mg.size() / 8;
  std::vector<double> vals( nvals );
  std::memcpy( vals.data(), mg.data(), nvals * 8 );
  fsys.m.assign( vals.begin(), vals.end() - dim );
  fsys.g.assign( vals.end() - dim, vals.end() );
  GravityOnlyGuard guard;
  const GravityOnlyGuard::Layout lay = layout == 0 ? GravityOnlyGuard::BALL2D : layout == 1 ? GravityOnlyGuard::RIGIDBODY2D : GravityOnlyGuard::RIGIDBODY3D;
  const unsigned n = guard.configureFromSnapshot( lay, blob.data(), blob.size(), "guard_harness" );
  if( int( n ) != fsys.n() ) { std::printf( "guard body count %u, expected %d\n", n, fsys.n() ); return 3; }
  VectorXs q( fsys.nqdofs() ), v( fsys.nvdofs() );
  q.setZero(); v.setZero();
  guard.verify( fsys, q, v, 0.0, lay, "guard_harness" );
  std::printf( "guard ok n=%u\n", n );
  return 0;
}
