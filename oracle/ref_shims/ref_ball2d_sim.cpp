// oracle/ref_shims/ref_ball2d_sim.cpp -- TEST INFRASTRUCTURE.
// Drives the reference's OWN Ball2DSim (ball2d/Ball2DSim.cpp + Ball2DState.cpp + everything they use, compiled unchanged by oracle/Makefile.ref against
// the Eigen stand-in): Ball2DSim::computeActiveSet as a whole -- broad phase, CCD, the portal branch with its teleported collisions, drums, planes, in
// the order the reference emits them -- and Ball2DSim::flow with an unconstrained map (portal updates and periodic boundaries included).  This is what
// pins the GLUE of the oracle (oracle/ball2d.h, oracle/ball2d_portals.h), which every other shim entry leaves restated.
// The one thing stubbed is ImpactMap::flow, which a Ball2DSim::flow overload that is never called here refers to (it needs the LCP solver stack).
#include "ball2d/Ball2DSim.h"
#include "ball2d/Ball2DState.h"
#include "ball2d/PythonScripting.h"
#include "ball2d/SymplecticEulerMap.h"
#include "ball2d/VerletMap.h"
#include "ball2d/Forces/Ball2DGravityForce.h"
#include "ball2d/StaticGeometry/StaticDrum.h"
#include "ball2d/StaticGeometry/StaticPlane.h"
#include "ball2d/Portals/PlanarPortal.h"
#include "ball2d/Constraints/BallBallConstraint.h"
#include "ball2d/Constraints/BallStaticPlaneConstraint.h"
#include "ball2d/Constraints/BallStaticDrumConstraint.h"
#include "scisim/ConstrainedMaps/ImpactMaps/ImpactMap.h"
#include "scisim/ConstrainedMaps/ImpactMaps/ImpactOperatorUtilities.h"
#include "scisim/Math/Rational.h"

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

// stub: see the header comment
void ImpactMap::flow( ScriptingCallback&, FlowableSystem&, ConstrainedSystem&, UnconstrainedMap&, ImpactOperator&, const unsigned, const scalar&, const scalar&, const VectorXs&, const VectorXs&, VectorXs&, VectorXs& )
{
  std::cerr << "oracle/ref_shims: ImpactMap::flow is not part of the compiled reference subset" << std::endl;
  std::abort();
}

extern "C"
{

// q, v: 2 n;  m, r: n;  planes: x[2], n[2] each;  drums: X[2], R;  portals: plane A ( x, n ), plane B ( x, n ), velocity, bounds -- the arguments of the
// reference's constructors (ball2d/StaticGeometry/StaticPlane.cpp:10-14, StaticDrum.cpp, Portals/PlanarPortal.cpp)
void* ref_ball2d_sim_create( const uint32_t n, const double* q, const double* v, const double* m, const double* r, const uint8_t* fixed, const double* g,
                             const uint32_t nplanes, const double* px, const double* pn, const uint32_t ndrums, const double* dx, const double* dr,
                             const uint32_t nportals, const double* pax, const double* pan, const double* pbx, const double* pbn, const double* pv, const double* pbounds )
{
  Ball2DSim* sim = new Ball2DSim;
  Ball2DState& s = sim->state();
  s.q().resize( int( 2 * n ) ); s.v().resize( int( 2 * n ) ); s.r().resize( int( n ) );
  VectorXs mass{ int( 2 * n ) };
  for( uint32_t b = 0; b < n; ++b )
  {
    for( int k = 0; k < 2; ++k ) { s.q()( int( 2 * b + k ) ) = q[2 * b + k]; s.v()( int( 2 * b + k ) ) = v[2 * b + k]; mass( int( 2 * b + k ) ) = m[b]; }
    s.r()( int( b ) ) = r[b];
    s.fixed().push_back( fixed != nullptr && fixed[b] != 0 );
  }
  s.setMass( mass );
  for( uint32_t k = 0; k < nplanes; ++k ) { s.staticPlanes().emplace_back( Vector2s{ px[2 * k], px[2 * k + 1] }, Vector2s{ pn[2 * k], pn[2 * k + 1] } ); }
  for( uint32_t k = 0; k < ndrums; ++k ) { s.staticDrums().emplace_back( Vector2s{ dx[2 * k], dx[2 * k + 1] }, dr[k] ); }
  for( uint32_t k = 0; k < nportals; ++k )
  {
    const StaticPlane a{ Vector2s{ pax[2 * k], pax[2 * k + 1] }, Vector2s{ pan[2 * k], pan[2 * k + 1] } };
    const StaticPlane b{ Vector2s{ pbx[2 * k], pbx[2 * k + 1] }, Vector2s{ pbn[2 * k], pbn[2 * k + 1] } };
    s.planarPortals().emplace_back( a, b, pv[k], pbounds[k] );
  }
  s.forces().emplace_back( new Ball2DGravityForce{ Vector2s{ g[0], g[1] } } );
  return sim;
}

void ref_ball2d_sim_destroy( void* h ) { delete static_cast<Ball2DSim*>( h ); }

// Ball2DSim::computeActiveSet( q0, q1, v ) (ball2d/Ball2DSim.cpp:151-173).  Per constraint, in the reference's order: the contact type code of
// include/scisim_b200.h (from name()), the indices, the world-space normal and contact point at q0 and penetrationDepth( q1 ).  Returns the count.
uint64_t ref_ball2d_sim_active_set( void* h, const double* q0, const double* q1, const double* v, const uint64_t cap, uint32_t* type, uint32_t* ci, uint32_t* cj,
                                    double* cn, double* cp, double* depth )
{
  Ball2DSim& sim = *static_cast<Ball2DSim*>( h );
  const int nq = int( sim.state().q().size() );
  VectorXs wq0{ nq }, wq1{ nq }, wv{ nq };
  for( int k = 0; k < nq; ++k ) { wq0( k ) = q0[k]; wq1( k ) = q1[k]; wv( k ) = v != nullptr ? v[k] : 0.0; }
  const VectorXs& vq0 = wq0; const VectorXs& vq1 = wq1;
  std::vector<std::unique_ptr<Constraint>> active_set;
  sim.computeActiveSet( vq0, vq1, wv, active_set );
  uint64_t k = 0;
  for( const std::unique_ptr<Constraint>& con : active_set )
  {
    if( k < cap )
    {
      const std::string name{ con->name() };
      if( name == "ball_ball" || name == "teleported_ball_ball" || name == "kinematic_kick_ball_ball" || name == "teleported_kinematic_kick_ball_ball" )
      {
        const BallBallConstraint& bb{ static_cast<const BallBallConstraint&>( *con ) };
        type[k] = ( name == "ball_ball" ) ? 0u : ( name == "teleported_ball_ball" ? 3u : ( name == "teleported_kinematic_kick_ball_ball" ? 4u : 99u ) );
        ci[k] = bb.idx0(); cj[k] = bb.idx1();
      }
      else if( name == "static_drum_constraint" )
      {
        const StaticDrumConstraint& d{ static_cast<const StaticDrumConstraint&>( *con ) };
        type[k] = 1u; ci[k] = d.ballIdx(); cj[k] = d.drumIdx();
      }
      else if( name == "static_plane_constraint" )
      {
        const StaticPlaneConstraint& p{ static_cast<const StaticPlaneConstraint&>( *con ) };
        type[k] = 2u; ci[k] = p.ballIdx(); cj[k] = p.planeIdx();
      }
      else { type[k] = 98u; ci[k] = cj[k] = 0u; }
      VectorXs n, p;
      con->getWorldSpaceContactNormal( vq0, n );
      con->getWorldSpaceContactPoint( vq0, p );
      cn[2 * k] = n( 0 ); cn[2 * k + 1] = n( 1 ); cp[2 * k] = p( 0 ); cp[2 * k + 1] = p( 1 );
      depth[k] = con->penetrationDepth( vq1 );
    }
    ++k;
  }
  return k;
}

// Ball2DSim::flow( call_back, iteration, dt, umap ) (ball2d/Ball2DSim.cpp:259-281): portals advanced to the step's time, the unconstrained map, the
// periodic boundary conditions; the simulation's state is advanced and returned.  kind 0: SymplecticEulerMap, 1: VerletMap; dt = dt_num / dt_den.
void ref_ball2d_sim_flow( void* h, const int kind, const unsigned iteration, const long long dt_num, const long long dt_den, double* q_out, double* v_out )
{
  Ball2DSim& sim = *static_cast<Ball2DSim*>( h );
  PythonScripting call_back;
  const Rational<std::intmax_t> dt{ std::intmax_t( dt_num ), std::intmax_t( dt_den ) };
  if( kind == 0 ) { SymplecticEulerMap umap; sim.flow( call_back, iteration, dt, umap ); }
  else { VerletMap umap; sim.flow( call_back, iteration, dt, umap ); }
  const int nq = int( sim.state().q().size() );
  for( int k = 0; k < nq; ++k ) { q_out[k] = sim.state().q()( k ); v_out[k] = sim.state().v()( k ); }
}

// positions and velocities of the simulation's state written by the caller (to continue from a state of its own)
void ref_ball2d_sim_set_state( void* h, const double* q, const double* v )
{
  Ball2DSim& sim = *static_cast<Ball2DSim*>( h );
  const int nq = int( sim.state().q().size() );
  for( int k = 0; k < nq; ++k ) { sim.state().q()( k ) = q[k]; sim.state().v()( k ) = v[k]; }
}

// One benchmark step of the reference's own code, timed (bench.py --impl reference / cpu_baseline): the state is set to ( q0, v0 ) (untimed),
// Ball2DSim::flow( call_back, iteration, dt, umap ) is timed, then Ball2DSim::computeActiveSet( q0, q1, v1 ) is timed -- constraint objects allocated
// as the reference allocates them; freeing them afterwards is not timed.  Returns the seconds of the two calls; *n_active = active_set.size().
double ref_ball2d_sim_step_timed( void* h, const double* q0, const double* v0, const int kind, const unsigned iteration, const long long dt_num, const long long dt_den,
                                  uint64_t* n_active, double* seconds_flow )
{
  Ball2DSim& sim = *static_cast<Ball2DSim*>( h );
  const int nq = int( sim.state().q().size() );
  VectorXs wq0{ nq };
  for( int k = 0; k < nq; ++k ) { wq0( k ) = q0[k]; sim.state().q()( k ) = q0[k]; sim.state().v()( k ) = v0[k]; }
  PythonScripting call_back;
  const Rational<std::intmax_t> dt{ std::intmax_t( dt_num ), std::intmax_t( dt_den ) };
  SymplecticEulerMap se;
  VerletMap verlet;
  UnconstrainedMap& umap = ( kind == 0 ) ? static_cast<UnconstrainedMap&>( se ) : static_cast<UnconstrainedMap&>( verlet );
  const auto t0 = std::chrono::steady_clock::now();
  sim.flow( call_back, iteration, dt, umap );
  const auto t1 = std::chrono::steady_clock::now();
  const VectorXs q1{ sim.state().q() };
  const VectorXs v1{ sim.state().v() };
  std::vector<std::unique_ptr<Constraint>> active_set;
  const auto t2 = std::chrono::steady_clock::now();
  sim.computeActiveSet( wq0, q1, v1, active_set );
  const auto t3 = std::chrono::steady_clock::now();
  *n_active = active_set.size();
  const double tf = std::chrono::duration<double>( t1 - t0 ).count();
  if( seconds_flow != nullptr ) { *seconds_flow = tf; }
  return tf + std::chrono::duration<double>( t3 - t2 ).count();
}

// Ball2DState::serialize (ball2d/Ball2DState.cpp:259-272) of the simulation's current state: the reference's binary snapshot.  Returns its length;
// the bytes are written when they fit cap.
uint64_t ref_ball2d_sim_serialize_state( void* h, void* buf, const uint64_t cap )
{
  const Ball2DSim& sim = *static_cast<const Ball2DSim*>( h );
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  sim.state().serialize( stm );
  const std::string bytes = stm.str();
  if( bytes.size() <= cap ) { std::memcpy( buf, bytes.data(), bytes.size() ); }
  return bytes.size();
}

// Ball2DState::deserialize (ball2d/Ball2DState.cpp:274-312) of a snapshot (the product's sg_ball2d_state_serialize output, say) into a fresh Ball2DSim
void* ref_ball2d_sim_from_snapshot( const void* buf, const uint64_t bytes )
{
  std::stringstream stm( std::ios::in | std::ios::out | std::ios::binary );
  stm.write( static_cast<const char*>( buf ), std::streamsize( bytes ) );
  Ball2DSim* sim = new Ball2DSim;
  sim->state().deserialize( stm );
  return sim;
}

// ImpactOperatorUtilities::computeN (scisim/ConstrainedMaps/ImpactMaps/ImpactOperatorUtilities.cpp:10-48, compiled unchanged) on the active set
// Ball2DSim::computeActiveSet( q0, q1, v ) returns, exactly as ImpactMap::flow calls it (ImpactMap.cpp:106-107: N sized Minv.cols() x ncollisions,
// gradients evaluated at q0), and Ball2DSim::computeContactBases (Ball2DSim.cpp:188-201) at ( q0, v ).  Outputs: N column-compressed ( outer: ncols + 1,
// inner / values: nnz ) and the 2 x 2 ncols contact bases, column-major.  Returns the number of constraints; *nnz = N.nonZeros().
uint64_t ref_ball2d_sim_compute_N( void* h, const double* q0, const double* q1, const double* v, const uint64_t cap_cols, const uint64_t cap_nnz, uint64_t* nnz,
                                   int* outer, int* inner, double* values, double* bases )
{
  Ball2DSim& sim = *static_cast<Ball2DSim*>( h );
  const int nq = int( sim.state().q().size() );
  VectorXs wq0{ nq }, wq1{ nq }, wv{ nq };
  for( int k = 0; k < nq; ++k ) { wq0( k ) = q0[k]; wq1( k ) = q1[k]; wv( k ) = v[k]; }
  const VectorXs& vq0 = wq0; const VectorXs& vq1 = wq1; const VectorXs& vv = wv;
  std::vector<std::unique_ptr<Constraint>> active_set;
  sim.computeActiveSet( vq0, vq1, vv, active_set );
  const unsigned ncollisions = unsigned( active_set.size() );
  const FlowableSystem& fsys = (const FlowableSystem&) sim; // Ball2DSim derives from FlowableSystem privately; ImpactMap receives it as a FlowableSystem&
  SparseMatrixsc N{ fsys.Minv().cols(), SparseMatrixsc::Index( ncollisions ) };
  ImpactOperatorUtilities::computeN( fsys, active_set, vq0, N );
  MatrixXXsc contact_bases;
  sim.computeContactBases( vq0, vv, active_set, contact_bases );
  *nnz = uint64_t( N.nonZeros() );
  if( ncollisions <= cap_cols && uint64_t( N.nonZeros() ) <= cap_nnz && ncollisions > 0 )
  {
    for( unsigned c = 0; c <= ncollisions; ++c ) { outer[c] = N.outerIndexPtr()[c]; }
    for( int e = 0; e < N.nonZeros(); ++e ) { inner[e] = N.innerIndexPtr()[e]; values[e] = N.valuePtr()[e]; }
    for( unsigned c = 0; c < ncollisions; ++c ) { for( int k = 0; k < 4; ++k ) { bases[4 * c + k] = contact_bases( k % 2, int( 2 * c ) + k / 2 ); } }
  }
  return ncollisions;
}

}
