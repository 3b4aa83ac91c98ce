// tests/scisim_plugin_example.cpp -- TEST INFRASTRUCTURE: the drop-in at the reference's own seam, end to end.
//
// The host shim (scisim_b200/host/gpu_backend.cpp) is compiled here with SCISIM_B200_WITH_SCISIM against the reference's OWN headers
// (scisim/UnconstrainedMaps/UnconstrainedMap.h, FlowableSystem.h, Constraints/ConstrainedSystem.h), so GpuSymplecticEulerMap / GpuVerletMap ARE
// UnconstrainedMaps of the reference, and linked with the reference's own Ball2DSim (oracle/_ref/libref_ball2d.so: Ball2DSim.cpp and everything
// it calls, compiled unchanged).  Two identical reference simulations are then stepped by the reference's own
//   Ball2DSim::flow( call_back, iteration, dt, umap )
// one with the reference's map, one with the GPU map plugged in -- the virtual call lands in the shim, which reads the system's force through
// the reference's FlowableSystem interface (GravityOnlyGuard) and integrates on the device -- and after every step the reference's own
// Ball2DSim::computeActiveSet is compared with GpuBall2DBackend::computeActiveSet, constraint by constraint.
//   usage: scisim_plugin_example [nballs] [steps] [verlet]        prints "plugin ok ..." and returns 0, or says what differed
#include "ball2d/Ball2DSim.h"
#include "ball2d/Ball2DState.h"
#include "ball2d/PythonScripting.h"
#include "ball2d/SymplecticEulerMap.h"
#include "ball2d/VerletMap.h"
#include "ball2d/Forces/Ball2DGravityForce.h"
#include "ball2d/StaticGeometry/StaticPlane.h"
#include "ball2d/Constraints/BallBallConstraint.h"
#include "ball2d/Constraints/BallStaticPlaneConstraint.h"
#include "scisim/Constraints/Constraint.h"
#include "scisim/Math/Rational.h"

#include "../scisim_b200/host/gpu_backend.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

static double uniform( uint64_t& s ) { s = s * 6364136223846793005ull + 1442695040888963407ull; return double( s >> 11 ) / 9007199254740992.0; }

static void fill( Ball2DSim& sim, const unsigned n, const double side, const std::vector<double>& px, const std::vector<double>& pn, const double gx, const double gy )
{
  Ball2DState& s = sim.state();
  s.q().resize( int( 2 * n ) ); s.v().resize( int( 2 * n ) ); s.r().resize( int( n ) );
  VectorXs mass{ int( 2 * n ) };
  uint64_t seed = 12345;
  for( unsigned b = 0; b < n; ++b )
  {
    s.r()( int( b ) ) = 0.05 + 0.35 * uniform( seed );
    const double m = 0.5 + uniform( seed );
    for( int k = 0; k < 2; ++k ) { s.q()( int( 2 * b + k ) ) = side * uniform( seed ); s.v()( int( 2 * b + k ) ) = 40.0 * ( uniform( seed ) - 0.5 ); mass( int( 2 * b + k ) ) = m; }
    s.fixed().push_back( false );
  }
  s.setMass( mass );
  for( size_t k = 0; k < px.size() / 2; ++k ) { s.staticPlanes().emplace_back( Vector2s{ px[2 * k], px[2 * k + 1] }, Vector2s{ pn[2 * k], pn[2 * k + 1] } ); }
  s.forces().emplace_back( new Ball2DGravityForce{ Vector2s{ gx, gy } } );
}

static bool same_bits( const VectorXs& a, const VectorXs& b ) { return a.size() == b.size() && ( a.size() == 0 || std::memcmp( a.data(), b.data(), size_t( a.size() ) * sizeof( double ) ) == 0 ); }

int main( int argc, char** argv )
{
  const unsigned n = ( argc > 1 ) ? unsigned( std::atoi( argv[1] ) ) : 3000u;
  const unsigned steps = ( argc > 2 ) ? unsigned( std::atoi( argv[2] ) ) : 5u;
  const bool verlet = argc > 3;
  const double side = 0.6 * std::sqrt( double( n ) ), gx = 0.3, gy = -9.81;
  const std::vector<double> px{ 0.0, 0.0, 0.0, 0.0, side, 0.0 }, pn{ 0.0, 1.0, 2.0, 0.0, -0.5, 0.0 };
  Ball2DSim sim_cpu, sim_gpu;
  fill( sim_cpu, n, side, px, pn, gx, gy );
  fill( sim_gpu, n, side, px, pn, gx, gy );

  GpuBall2DBackend backend{ 0 };
  {
    VectorXs m{ int( n ) };
    for( unsigned b = 0; b < n; ++b ) { m( int( b ) ) = sim_gpu.state().M().valuePtr()[2 * b]; }
    backend.setBodies( sim_gpu.state().r(), m );
    backend.setGravity( gx, gy );
    backend.setPlanes( px, pn );
  }
  GpuSymplecticEulerMap gpu_se{ backend };
  GpuVerletMap gpu_verlet{ backend };
  SymplecticEulerMap cpu_se;
  VerletMap cpu_verlet;
  UnconstrainedMap& gpu_map = verlet ? static_cast<UnconstrainedMap&>( gpu_verlet ) : static_cast<UnconstrainedMap&>( gpu_se );
  UnconstrainedMap& cpu_map = verlet ? static_cast<UnconstrainedMap&>( cpu_verlet ) : static_cast<UnconstrainedMap&>( cpu_se );
  if( gpu_map.name() != cpu_map.name() ) { std::printf( "map names differ: %s vs %s\n", gpu_map.name().c_str(), cpu_map.name().c_str() ); return 1; }

  PythonScripting call_back;
  const Rational<std::intmax_t> dt{ 1, 100 };
  uint64_t total_contacts = 0;
  for( unsigned it = 1; it <= steps; ++it )
  {
    const VectorXs q0{ sim_cpu.state().q() };
    sim_cpu.flow( call_back, it, dt, cpu_map );   // the reference's own step with its own map
    sim_gpu.flow( call_back, it, dt, gpu_map );   // the reference's own step, the GPU map plugged in
    if( !same_bits( sim_cpu.state().q(), sim_gpu.state().q() ) || !same_bits( sim_cpu.state().v(), sim_gpu.state().v() ) ) { std::printf( "step %u: q1 / v1 differ\n", it ); return 1; }
    const VectorXs q1{ sim_cpu.state().q() };
    const VectorXs v1{ sim_cpu.state().v() };
    std::vector<std::unique_ptr<Constraint>> active_set;
    sim_cpu.computeActiveSet( q0, q1, v1, active_set );
    std::vector<GpuContact2D> contacts;
    backend.computeActiveSet( q0, q1, contacts, nullptr, true );
    if( contacts.size() != active_set.size() ) { std::printf( "step %u: %zu contacts from the GPU, %zu constraints from the reference\n", it, contacts.size(), active_set.size() ); return 1; }
    for( size_t k = 0; k < contacts.size(); ++k )
    {
      const Constraint& con = *active_set[k];
      const std::string name{ con.name() };
      unsigned type = 99u, i = 0u, j = 0u;
      if( name == "ball_ball" ) { const BallBallConstraint& bb = static_cast<const BallBallConstraint&>( con ); type = SG_BALL_BALL; i = bb.idx0(); j = bb.idx1(); }
      else if( name == "static_plane_constraint" ) { const StaticPlaneConstraint& pc = static_cast<const StaticPlaneConstraint&>( con ); type = SG_BALL_PLANE; i = pc.ballIdx(); j = pc.planeIdx(); }
      VectorXs nrm, pt;
      con.getWorldSpaceContactNormal( q0, nrm );
      con.getWorldSpaceContactPoint( q0, pt );
      const GpuContact2D& g = contacts[k];
      if( g.type != type || g.i != i || g.j != j || std::memcmp( g.n, nrm.data(), 16 ) != 0 || std::memcmp( g.p, pt.data(), 16 ) != 0 )
      {
        std::printf( "step %u, constraint %zu: ( %u, %u, %u ) from the GPU, ( %u, %u, %u ) %s from the reference, or normal / point differ\n", it, k, g.type, g.i, g.j, type, i, j, name.c_str() );
        return 1;
      }
    }
    total_contacts += contacts.size();
  }
  std::printf( "plugin ok: %u balls, %u steps of Ball2DSim::flow with %s from the GPU shim, %llu constraints equal the reference's own\n", n, steps, gpu_map.name().c_str(), ( unsigned long long )( total_contacts ) );
  return 0;
}
