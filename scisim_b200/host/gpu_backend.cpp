// gpu_backend.cpp -- see gpu_backend.h
#include "gpu_backend.h"
#include "../csrc/sg_rb2d_snapshot.h" // plain C++ writers / parsers of the reference's snapshots (also includes sg_rb3d_snapshot.h): after a
                                       // deserializeState the force guard is told the masses and the gravity the device now holds

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <iterator>
#include <limits>
#include <numeric>
#include <sstream>

void GravityOnlyGuard::verify( FlowableSystem& fsys, const VectorXs& q0, const VectorXs& v0, const scalar& t, const Layout layout, const char* who )
{
  if( m_checked ) { return; }
  VectorXs F( v0.size() );
  F.setZero();
  fsys.computeForce( q0, v0, t, F );
  const long n = static_cast<long>( m_mass.size() );
  const int dim = ( layout == RIGIDBODY3D ) ? 3 : 2;
  const long per_body = ( layout == BALL2D ) ? 2 : 3;
  bool ok = F.size() == ( layout == RIGIDBODY3D ? 6 * n : per_body * n );
  for( long i = 0; ok && i < n; ++i )
  {
    for( int k = 0; k < dim; ++k ) { ok = ok && F( per_body * i + k ) == 0.0 + m_mass[std::size_t( i )] * m_g[k]; } // 0 + m g: the accumulation into a zeroed F
    if( layout == RIGIDBODY2D ) { ok = ok && F( 3 * i + 2 ) == 0.0; }
    if( layout == RIGIDBODY3D ) { for( int k = 0; k < 3; ++k ) { ok = ok && F( 3 * n + 3 * i + k ) == 0.0; } }
  }
  if( !ok )
  {
    std::cerr << who << ": the system's forces are not the single near-earth gravity configured on the GPU back end; keep the CPU map for this scene. Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
  m_checked = true;
}

// After a deserializeState the device integrates with the masses and the gravity of the snapshot: the guard is told both, so that the first GPU map flow
// after a restore is checked against them.  The library has accepted the stream when this runs; a stream it would not accept ends the process here.
unsigned GravityOnlyGuard::configureFromSnapshot( const Layout layout, const char* buf, const std::size_t bytes, const char* who, std::size_t* consumed )
{
  unsigned n = 0;
  const char* why = tryConfigureFromSnapshot( layout, buf, bytes, n, consumed );
  if( why[0] != '\0' )
  {
    std::cerr << who << ": " << why << ". Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
  return n;
}

// "" on success, else what is wrong with the stream (nothing is configured then)
const char* GravityOnlyGuard::tryConfigureFromSnapshot( const Layout layout, const char* buf, const std::size_t bytes, unsigned& n, std::size_t* consumed )
{
  sg_snapshot::Source in{ reinterpret_cast<const unsigned char*>( buf ), bytes, 0, true };
  const char* why = "";
  n = 0;
  if( layout == BALL2D )
  {
    // Ball2DState::serialize (ball2d/Ball2DState.cpp:259-272): q, v, r, fixed, M, Minv, drums, planes, portals, forces; M's values hold each mass twice
    const uint64_t nq = uint64_t( in.val<long long>() ), nb = nq / 2;
    in.take( nq * 8 );                                   // q
    in.take( 8 + nq * 8 );                               // v
    in.take( 8 + nb * 8 );                               // r
    in.take( 8 + nb );                                   // fixed
    in.take( 24 + nq * 4 + ( nq + 1 ) * 4 );             // M: header, inner, outer
    std::vector<double> mass( in.ok && nq * 8 <= bytes ? nq : 0 );
    in.doubles( mass.data(), nq );
    in.take( 24 + nq * 4 + ( nq + 1 ) * 4 + nq * 8 );    // Minv
    const uint64_t ndrums = in.val<size_t>(); in.take( ndrums * 24 );
    const uint64_t nplanes = in.val<size_t>(); in.take( nplanes * 64 );
    const uint64_t nportals = in.val<size_t>(); in.take( nportals * ( 2 * 64 + 24 ) );
    const uint64_t nforces = in.val<size_t>();
    double g[2] = { 0.0, 0.0 };
    for( uint64_t k = 0; k < nforces && in.ok; ++k )
    {
      const uint64_t len = in.val<size_t>();
      in.take( len );
      double gk[2] = { 0.0, 0.0 };
      in.doubles( gk, 2 );
      g[0] += gk[0]; g[1] += gk[1];                      // forces accumulate
    }
    if( !in.ok ) { why = "the snapshot ends early"; }
    else { n = static_cast<unsigned>( nb ); setMasses( mass.data(), n, 2 ); setGravity( g[0], g[1], 0.0 ); }
  }
  else if( layout == RIGIDBODY2D )
  {
    sg_snapshot::Rb2dState st;
    if( sg_snapshot::parse( in, st, &why ) == 0 ) { n = st.n; setMasses( st.M.data(), n, 3 ); setGravity( st.g[0], st.g[1], 0.0 ); } // M = diag( m, m, I ) per body
  }
  else
  {
    sg_snapshot::Rb3dState st;
    if( sg_snapshot::parse( in, st, &why ) == 0 ) { n = st.n; setMasses( st.m.data(), n, 1 ); setGravity( st.g[0], st.g[1], st.g[2] ); }
  }
  if( why[0] == '\0' && consumed != nullptr ) { *consumed = static_cast<std::size_t>( in.n ); }
  return why;
}

// <Sim>::deserialize reads the state and then the constraint cache from one stream (Ball2DSim.cpp:817-822, RigidBody2DSim.cpp:1147-1152, RigidBody3DSim.cpp:1600-1606).
// A state snapshot is self-delimiting only to its parser, so the wrappers below read the rest of the stream, hand it to the library (which ignores what follows
// the state) and then put the stream back to the first byte after the state, ready for PairImpulseCache::deserialize.
std::vector<char> sghReadRest( std::istream& input_stream, std::istream::pos_type& start )
{
  start = input_stream.tellg();
  return std::vector<char>( ( std::istreambuf_iterator<char>( input_stream ) ), std::istreambuf_iterator<char>() );
}

void sghRewindBehindState( std::istream& input_stream, const std::istream::pos_type start, const std::size_t consumed )
{
  if( start == std::istream::pos_type( -1 ) ) { return; } // not seekable: the caller gave a stream that ends with the state
  input_stream.clear();
  input_stream.seekg( start + std::streamoff( consumed ) );
}

GpuBall2DBackend::GpuBall2DBackend( const int device )
: m_ctx( nullptr )
, m_nbodies( 0 )
{
  const int rc = sg_create( &m_ctx, device );
  if( rc != SG_OK )
  {
    std::cerr << "GpuBall2DBackend: " << sg_last_error( nullptr ) << " Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
}

GpuBall2DBackend::~GpuBall2DBackend() { sg_destroy( m_ctx ); }

void GpuBall2DBackend::check( const int rc, const char* what ) const
{
  if( rc != SG_OK )
  {
    std::cerr << what << ": " << sg_last_error( m_ctx ) << " Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
}

void GpuBall2DBackend::setBodies( const VectorXs& r, const VectorXs& m )
{
  m_nbodies = static_cast<unsigned>( r.size() );
  check( sg_ball2d_set_bodies( m_ctx, m_nbodies, r.data(), m.data() ), "sg_ball2d_set_bodies" );
  m_guard.setMasses( m.data(), m_nbodies, 1 );
}

void GpuBall2DBackend::setGravity( const double gx, const double gy )
{
  const double g[2] = { gx, gy };
  check( sg_ball2d_set_gravity( m_ctx, g ), "sg_ball2d_set_gravity" );
  m_guard.setGravity( gx, gy, 0.0 );
}

void GpuBall2DBackend::setPlanes( const std::vector<double>& x, const std::vector<double>& n )
{
  check( sg_ball2d_set_planes( m_ctx, static_cast<uint32_t>( x.size() / 2 ), x.data(), n.data() ), "sg_ball2d_set_planes" );
}

void GpuBall2DBackend::setDrums( const std::vector<double>& x, const std::vector<double>& r )
{
  check( sg_ball2d_set_drums( m_ctx, static_cast<uint32_t>( r.size() ), x.data(), r.data() ), "sg_ball2d_set_drums" );
}

void GpuBall2DBackend::setPortals( const std::vector<double>& plane_a_x, const std::vector<double>& plane_a_n, const std::vector<double>& plane_b_x, const std::vector<double>& plane_b_n,
                                   const std::vector<double>& velocity, const std::vector<double>& bounds )
{
  check( sg_ball2d_set_portals( m_ctx, static_cast<uint32_t>( velocity.size() ), plane_a_x.data(), plane_a_n.data(), plane_b_x.data(), plane_b_n.data(), velocity.data(), bounds.data() ), "sg_ball2d_set_portals" );
}

void GpuBall2DBackend::updatePeriodicBoundaryConditionsStartOfStep( const unsigned next_iteration, const scalar& dt )
{
  // const scalar t{ next_iteration * dt };  (ball2d/Ball2DSim.cpp:329)
  const scalar t{ next_iteration * dt };
  check( sg_ball2d_update_portals( m_ctx, t, nullptr ), "sg_ball2d_update_portals" );
}

void GpuBall2DBackend::enforcePeriodicBoundaryConditions( VectorXs& q, VectorXs& v )
{
  check( sg_ball2d_enforce_portals( m_ctx, q.data(), v.data() ), "sg_ball2d_enforce_portals" );
}

void GpuBall2DBackend::teleportedContacts( std::vector<GpuTeleportedContact2D>& teleported, uint64_t* num_regular )
{
  sg_teleported t;
  check( sg_ball2d_teleported( m_ctx, &t ), "sg_ball2d_teleported" );
  teleported.resize( t.n_teleported );
  for( uint64_t k = 0; k < t.n_teleported; ++k )
  {
    GpuTeleportedContact2D& o = teleported[k];
    o.portal0 = t.portal0[k]; o.portal1 = t.portal1[k];
    for( int c = 0; c < 2; ++c )
    {
      o.x0[c] = t.x0[2 * k + c]; o.x1[c] = t.x1[2 * k + c]; o.kick[c] = t.kick[2 * k + c];
      o.delta0[c] = t.delta0 != nullptr ? t.delta0[2 * k + c] : std::numeric_limits<double>::quiet_NaN();
      o.delta1[c] = t.delta1 != nullptr ? t.delta1[2 * k + c] : std::numeric_limits<double>::quiet_NaN();
    }
  }
  if( num_regular != nullptr ) { *num_regular = t.n_regular; }
}

void GpuBall2DBackend::flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  // Outputs are pre-sized by the caller in the reference (ball2d/Ball2DSim.cpp:311-312); be lenient here
  if( q1.size() != q0.size() ) { q1.resize( q0.size() ); }
  if( v1.size() != v0.size() ) { v1.resize( v0.size() ); }
  check( sg_ball2d_flow( m_ctx, map_kind, q0.data(), v0.data(), dt, q1.data(), v1.data() ), "sg_ball2d_flow" );
}

void GpuBall2DBackend::computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact2D>& contacts, uint64_t* num_candidates, const bool from_last_flow )
{
  sg_contacts c;
  check( sg_ball2d_active_set( m_ctx, q0.data(), q1.data(), SG_OUT_NORMALS | SG_OUT_POINTS | SG_OUT_DEPTHS | ( from_last_flow ? SG_IN_RESIDENT : 0u ), &c ), "sg_ball2d_active_set" );
  contacts.resize( c.n_active );
  for( uint64_t k = 0; k < c.n_active; ++k )
  {
    GpuContact2D& o = contacts[k];
    o.type = c.type[k]; o.i = c.i[k]; o.j = c.j[k];
    o.n[0] = c.n[2 * k]; o.n[1] = c.n[2 * k + 1];
    o.p[0] = c.p[2 * k]; o.p[1] = c.p[2 * k + 1];
    o.depth = c.depth[k];
  }
  if( num_candidates != nullptr ) { *num_candidates = c.n_candidates; }
}

void GpuBall2DBackend::assemble( const uint32_t flags, sg_assembly& out )
{
  check( sg_ball2d_assemble( m_ctx, flags, &out ), "sg_ball2d_assemble" );
}

void GpuBall2DBackend::cacheStore( const unsigned ncomp, const VectorXs& r )
{
  check( sg_ball2d_cache_store( m_ctx, ncomp, r.data() ), "sg_ball2d_cache_store" );
}

uint64_t GpuBall2DBackend::cacheLookup( const unsigned ncomp, VectorXs& r )
{
  uint64_t hits = 0;
  check( sg_ball2d_cache_lookup( m_ctx, ncomp, r.data(), &hits ), "sg_ball2d_cache_lookup" );
  return hits;
}

void GpuBall2DBackend::cacheClear()
{
  check( sg_ball2d_cache_clear( m_ctx ), "sg_ball2d_cache_clear" );
}

void GpuBall2DBackend::serializeState( std::ostream& output_stream, const bool from_last_flow ) const
{
  uint64_t bytes = 0;
  check( sg_ball2d_state_serialize( m_ctx, from_last_flow ? 1 : 0, nullptr, 0, &bytes ), "sg_ball2d_state_serialize" );
  std::vector<char> buf( bytes );
  check( sg_ball2d_state_serialize( m_ctx, from_last_flow ? 1 : 0, buf.data(), bytes, &bytes ), "sg_ball2d_state_serialize" );
  output_stream.write( buf.data(), static_cast<std::streamsize>( bytes ) );
}

void GpuBall2DBackend::deserializeState( std::istream& input_stream )
{
  std::istream::pos_type start;
  const std::vector<char> buf = sghReadRest( input_stream, start );
  check( sg_ball2d_state_deserialize( m_ctx, buf.data(), buf.size() ), "sg_ball2d_state_deserialize" );
  std::size_t consumed = 0;
  m_nbodies = m_guard.configureFromSnapshot( GravityOnlyGuard::BALL2D, buf.data(), buf.size(), "GpuBall2DBackend::deserializeState", &consumed );
  sghRewindBehindState( input_stream, start, consumed );
}

void GpuBall2DBackend::getPotentialOverlaps( const std::vector<double>& aabbs, std::vector<std::pair<unsigned,unsigned>>& overlaps )
{
  sg_pairs p;
  check( sg_candidate_pairs( m_ctx, 2, static_cast<uint32_t>( aabbs.size() / 4 ), aabbs.data(), &p ), "sg_candidate_pairs" );
  // the reference appends to a std::set; the list arrives in that set's iteration order
  overlaps.reserve( overlaps.size() + p.n );
  for( uint64_t k = 0; k < p.n; ++k ) { overlaps.emplace_back( p.ij[2 * k], p.ij[2 * k + 1] ); }
}

void GpuSymplecticEulerMap::flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  m_backend.forceGuard().verify( fsys, q0, v0, iteration * dt, GravityOnlyGuard::BALL2D, "GpuSymplecticEulerMap" );
  m_backend.flow( SG_MAP_SYMPLECTIC_EULER, q0, v0, dt, q1, v1 );
}

void GpuVerletMap::flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  m_backend.forceGuard().verify( fsys, q0, v0, iteration * dt, GravityOnlyGuard::BALL2D, "GpuVerletMap" );
  m_backend.flow( SG_MAP_VERLET, q0, v0, dt, q1, v1 );
}

// ---- several GPUs of this process -------------------------------------------------------------------
GpuBall2DMultiBackend::GpuBall2DMultiBackend( const std::vector<int>& devices )
: m_multi( nullptr )
{
  const int rc = sg_create_multi( &m_multi, static_cast<int>( devices.size() ), devices.data() );
  if( rc != SG_OK )
  {
    std::cerr << "GpuBall2DMultiBackend: " << sg_multi_last_error( nullptr ) << " Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
}

GpuBall2DMultiBackend::~GpuBall2DMultiBackend() { sg_destroy_multi( m_multi ); }

void GpuBall2DMultiBackend::check( const int rc, const char* what ) const
{
  if( rc != SG_OK )
  {
    std::cerr << what << ": " << sg_multi_last_error( m_multi ) << " Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
}

void GpuBall2DMultiBackend::setBodies( const VectorXs& r, const VectorXs& m )
{
  check( sg_multi_ball2d_set_bodies( m_multi, static_cast<uint32_t>( r.size() ), r.data(), m.data() ), "sg_multi_ball2d_set_bodies" );
  m_guard.setMasses( m.data(), static_cast<unsigned>( r.size() ), 1 );
}
void GpuBall2DMultiBackend::setGravity( const double gx, const double gy )
{
  const double g[2] = { gx, gy };
  check( sg_multi_ball2d_set_gravity( m_multi, g ), "sg_multi_ball2d_set_gravity" );
  m_guard.setGravity( gx, gy, 0.0 );
}
void GpuBall2DMultiBackend::setPlanes( const std::vector<double>& x, const std::vector<double>& n ) { check( sg_multi_ball2d_set_planes( m_multi, static_cast<uint32_t>( x.size() / 2 ), x.data(), n.data() ), "sg_multi_ball2d_set_planes" ); }
void GpuBall2DMultiBackend::setDrums( const std::vector<double>& x, const std::vector<double>& r ) { check( sg_multi_ball2d_set_drums( m_multi, static_cast<uint32_t>( r.size() ), x.data(), r.data() ), "sg_multi_ball2d_set_drums" ); }

void GpuBall2DMultiBackend::flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  if( q1.size() != q0.size() ) { q1.resize( q0.size() ); }
  if( v1.size() != v0.size() ) { v1.resize( v0.size() ); }
  check( sg_multi_ball2d_flow( m_multi, map_kind, q0.data(), v0.data(), dt, q1.data(), v1.data() ), "sg_multi_ball2d_flow" );
}

void GpuBall2DMultiBackend::computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact2D>& contacts, uint64_t* num_candidates, const bool from_last_flow )
{
  sg_contacts c;
  check( sg_multi_ball2d_active_set( m_multi, q0.data(), q1.data(), SG_OUT_NORMALS | SG_OUT_POINTS | SG_OUT_DEPTHS | ( from_last_flow ? SG_IN_RESIDENT : 0u ), &c ), "sg_multi_ball2d_active_set" );
  contacts.resize( c.n_active );
  for( uint64_t k = 0; k < c.n_active; ++k )
  {
    GpuContact2D& o = contacts[k];
    o.type = c.type[k]; o.i = c.i[k]; o.j = c.j[k];
    o.n[0] = c.n[2 * k]; o.n[1] = c.n[2 * k + 1];
    o.p[0] = c.p[2 * k]; o.p[1] = c.p[2 * k + 1];
    o.depth = c.depth[k];
  }
  if( num_candidates != nullptr ) { *num_candidates = c.n_candidates; }
}

unsigned GpuBall2DMultiBackend::numGpus() const { return static_cast<unsigned>( sg_multi_n_gpus( m_multi ) ); }

uint64_t GpuBall2DMultiBackend::numPartitions()
{
  uint64_t np = 0;
  check( sg_multi_partition_info( m_multi, nullptr, nullptr, nullptr, &np ), "sg_multi_partition_info" );
  return np;
}

void GpuMultiSymplecticEulerMap::flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  m_backend.forceGuard().verify( fsys, q0, v0, iteration * dt, GravityOnlyGuard::BALL2D, "GpuMultiSymplecticEulerMap" );
  m_backend.flow( SG_MAP_SYMPLECTIC_EULER, q0, v0, dt, q1, v1 );
}

void GpuMultiVerletMap::flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  m_backend.forceGuard().verify( fsys, q0, v0, iteration * dt, GravityOnlyGuard::BALL2D, "GpuMultiVerletMap" );
  m_backend.flow( SG_MAP_VERLET, q0, v0, dt, q1, v1 );
}

void PairImpulseCache::Table::sortIfNeeded()
{
  if( sorted ) { return; }
  std::vector<std::size_t> order( keys.size() );
  std::iota( order.begin(), order.end(), std::size_t( 0 ) );
  // stable: of several entries with one key the first stored stays first, and lookup returns the first (std::map::insert keeps the first)
  std::stable_sort( order.begin(), order.end(), [this]( const std::size_t a, const std::size_t b ) { return keys[a] < keys[b]; } );
  std::vector<uint64_t> k2( keys.size() );
  std::vector<double> v2( values.size() );
  for( std::size_t n = 0; n < order.size(); ++n )
  {
    k2[n] = keys[order[n]];
    std::memcpy( &v2[n * width], &values[order[n] * width], width * sizeof( double ) );
  }
  keys.swap( k2 ); values.swap( v2 );
  sorted = true;
}

void PairImpulseCache::clear()
{
  for( Table& t : m_tables ) { t.keys.clear(); t.values.clear(); t.sorted = true; }
}

bool PairImpulseCache::empty() const
{
  return m_tables[0].keys.empty() && m_tables[1].keys.empty() && m_tables[2].keys.empty() && m_tables[3].keys.empty();
}

void PairImpulseCache::cacheConstraint( const int kind, const unsigned a, const unsigned b, const VectorXs& r )
{
  if( kind < 0 || kind > 3 )
  {
    std::cerr << "constraint kind " << kind << " not supported in PairImpulseCache::cacheConstraint. Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
  Table& t = m_tables[kind];
  const uint64_t key = ( uint64_t( a ) << 32 ) | b;
  if( t.keys.empty() ) { t.width = static_cast<unsigned>( r.size() ); }
  if( !t.keys.empty() && key <= t.keys.back() ) { t.sorted = false; }
  t.keys.push_back( key );
  t.values.insert( t.values.end(), r.data(), r.data() + r.size() );
}

void PairImpulseCache::getCachedConstraint( const int kind, const unsigned a, const unsigned b, VectorXs& r ) const
{
  if( kind < 0 || kind > 3 )
  {
    std::cerr << "constraint kind " << kind << " not supported in PairImpulseCache::getCachedConstraint. Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
  Table& t = m_tables[kind];
  t.sortIfNeeded();
  const uint64_t key = ( uint64_t( a ) << 32 ) | b;
  const auto it = std::lower_bound( t.keys.begin(), t.keys.end(), key );
  if( it != t.keys.end() && *it == key )
  {
    const std::size_t n = static_cast<std::size_t>( it - t.keys.begin() );
    for( long d = 0; d < r.size(); ++d ) { r( d ) = t.values[n * t.width + static_cast<std::size_t>( d )]; }
    return;
  }
  // If the constraint was not found set to a default force of 0 (ball2d/ConstraintCache.cpp:122)
  r.setZero();
}

// the order in which each sim's ConstraintCache::serialize writes its maps, as kinds of this class (see the header); -1 ends the list
static const int g_cache_order[3][5] = { { 0, 1, 2, -1, -1 },     // ball2d: ball-ball, plane-ball, drum-ball
                                         { 0, 2, 3, 1, -1 },      // rigidbody2d: circle-circle, body-body, kinematic object - circle, plane-circle
                                         { 0, 1, 2, 3, -1 } };    // rigidbody3d: sphere-sphere, plane-sphere, cylinder-sphere, kinematic sphere-sphere

void PairImpulseCache::serialize( const Sim sim, std::ostream& output_stream ) const
{
  for( const int* kind = g_cache_order[sim]; *kind >= 0; ++kind )
  {
    Table& t = m_tables[*kind];
    t.sortIfNeeded();
    // a key stored twice occupies two slots here and one in the reference's std::map (its first impulse): count and write the first of each run
    std::size_t unique = 0;
    for( std::size_t n = 0; n < t.keys.size(); ++n ) { if( n == 0 || t.keys[n] != t.keys[n - 1] ) { ++unique; } }
    output_stream.write( reinterpret_cast<const char*>( &unique ), sizeof( std::size_t ) );
    for( std::size_t n = 0; n < t.keys.size(); ++n )
    {
      if( n > 0 && t.keys[n] == t.keys[n - 1] ) { continue; }
      const unsigned first = static_cast<unsigned>( t.keys[n] >> 32 ), second = static_cast<unsigned>( t.keys[n] & 0xffffffffu );
      const long long rows = static_cast<long long>( t.width );
      output_stream.write( reinterpret_cast<const char*>( &first ), sizeof( unsigned ) );
      output_stream.write( reinterpret_cast<const char*>( &second ), sizeof( unsigned ) );
      output_stream.write( reinterpret_cast<const char*>( &rows ), sizeof( long long ) );
      output_stream.write( reinterpret_cast<const char*>( &t.values[n * t.width] ), std::streamsize( t.width * sizeof( double ) ) );
    }
  }
}

bool PairImpulseCache::deserialize( const Sim sim, std::istream& input_stream )
{
  clear();
  for( const int* kind = g_cache_order[sim]; *kind >= 0; ++kind )
  {
    Table& t = m_tables[*kind];
    std::size_t count = 0;
    input_stream.read( reinterpret_cast<char*>( &count ), sizeof( std::size_t ) );
    if( !input_stream.good() ) { clear(); return false; }
    for( std::size_t n = 0; n < count; ++n )
    {
      unsigned first = 0, second = 0;
      long long rows = -1;
      input_stream.read( reinterpret_cast<char*>( &first ), sizeof( unsigned ) );
      input_stream.read( reinterpret_cast<char*>( &second ), sizeof( unsigned ) );
      input_stream.read( reinterpret_cast<char*>( &rows ), sizeof( long long ) );
      // one width per map: the impulses of one constraint class have one length
      if( !input_stream.good() || rows < 0 || rows > 64 || ( n > 0 && static_cast<unsigned>( rows ) != t.width ) ) { clear(); return false; }
      if( n == 0 ) { t.width = static_cast<unsigned>( rows ); }
      const uint64_t key = ( uint64_t( first ) << 32 ) | second;
      if( !t.keys.empty() && key <= t.keys.back() ) { t.sorted = false; }
      t.keys.push_back( key );
      t.values.resize( t.values.size() + t.width );
      input_stream.read( reinterpret_cast<char*>( t.values.data() + ( t.values.size() - t.width ) ), std::streamsize( t.width * sizeof( double ) ) );
      if( !input_stream.good() ) { clear(); return false; }
    }
  }
  return true;
}

extern "C"
{
void* sgh_cache_create() { return new PairImpulseCache; }
void sgh_cache_destroy( void* cache ) { delete static_cast<PairImpulseCache*>( cache ); }
void sgh_cache_clear( void* cache ) { static_cast<PairImpulseCache*>( cache )->clear(); }
int sgh_cache_empty( const void* cache ) { return static_cast<const PairImpulseCache*>( cache )->empty() ? 1 : 0; }
void sgh_cache_store( void* cache, int kind, unsigned a, unsigned b, const double* r, unsigned ncomp )
{
  VectorXs v( static_cast<long>( ncomp ) );
  for( unsigned c = 0; c < ncomp; ++c ) { v( c ) = r[c]; }
  static_cast<PairImpulseCache*>( cache )->cacheConstraint( kind, a, b, v );
}
void sgh_cache_lookup( const void* cache, int kind, unsigned a, unsigned b, double* r, unsigned ncomp )
{
  VectorXs v( static_cast<long>( ncomp ) );
  static_cast<const PairImpulseCache*>( cache )->getCachedConstraint( kind, a, b, v );
  for( unsigned c = 0; c < ncomp; ++c ) { r[c] = v( c ); }
}
uint64_t sgh_cache_serialize( const void* cache, int sim, void* buf, uint64_t cap )
{
  std::ostringstream stm( std::ios::out | std::ios::binary );
  static_cast<const PairImpulseCache*>( cache )->serialize( static_cast<PairImpulseCache::Sim>( sim ), stm );
  const std::string bytes = stm.str();
  if( buf != nullptr && bytes.size() <= cap ) { std::memcpy( buf, bytes.data(), bytes.size() ); }
  return bytes.size();
}
uint64_t sgh_state_snapshot_length( int sim, const void* buf, uint64_t bytes )
{
  GravityOnlyGuard guard;
  unsigned n = 0;
  std::size_t consumed = 0;
  const GravityOnlyGuard::Layout layout = sim == 0 ? GravityOnlyGuard::BALL2D : sim == 1 ? GravityOnlyGuard::RIGIDBODY2D : GravityOnlyGuard::RIGIDBODY3D;
  const char* why = guard.tryConfigureFromSnapshot( layout, static_cast<const char*>( buf ), static_cast<std::size_t>( bytes ), n, &consumed );
  return why[0] == '\0' ? uint64_t( consumed ) : 0u;
}
int sgh_cache_deserialize( void* cache, int sim, const void* buf, uint64_t bytes )
{
  std::istringstream stm( std::string( static_cast<const char*>( buf ), static_cast<std::size_t>( bytes ) ), std::ios::in | std::ios::binary );
  return static_cast<PairImpulseCache*>( cache )->deserialize( static_cast<PairImpulseCache::Sim>( sim ), stm ) ? 1 : 0;
}
}

// ---- rigidbody3d ---------------------------------------------------------------------------------------------------
GpuRigidBody3DBackend::GpuRigidBody3DBackend( const int device )
: m_ctx( nullptr )
, m_nbodies( 0 )
{
  const int rc = sg_create( &m_ctx, device );
  if( rc != SG_OK )
  {
    std::cerr << "GpuRigidBody3DBackend: " << sg_last_error( nullptr ) << " Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
}

GpuRigidBody3DBackend::~GpuRigidBody3DBackend() { sg_destroy( m_ctx ); }

void GpuRigidBody3DBackend::check( const int rc, const char* what ) const
{
  if( rc != SG_OK )
  {
    // includes SG_ERR_UNSUPPORTED: the pairings for which RigidBody3DSim::dispatchNarrowPhaseCollision prints and
    // exits (RigidBody3DSim.cpp:905-961)
    std::cerr << what << ": " << sg_last_error( m_ctx ) << " Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
}

void GpuRigidBody3DBackend::setGeometry( const std::vector<uint32_t>& type, const std::vector<double>& r, const std::vector<double>& half_widths, const std::vector<uint32_t>& mesh )
{
  check( sg_rb3d_set_geometry( m_ctx, static_cast<uint32_t>( type.size() ), type.data(), r.data(), half_widths.data(), mesh.data() ), "sg_rb3d_set_geometry" );
}

uint32_t GpuRigidBody3DBackend::addMesh( const std::vector<double>& verts, const std::vector<double>& samples, const std::vector<double>& hull, const double cell_delta[3], const uint32_t dims[3],
                                         const double origin[3], const std::vector<double>& sdf )
{
  uint32_t index = 0;
  check( sg_rb3d_add_mesh( m_ctx, static_cast<uint32_t>( verts.size() / 3 ), verts.data(), static_cast<uint32_t>( samples.size() / 3 ), samples.data(), static_cast<uint32_t>( hull.size() / 3 ), hull.data(),
                           cell_delta, dims, origin, sdf.data(), &index ), "sg_rb3d_add_mesh" );
  return index;
}

void GpuRigidBody3DBackend::setMeshSnapshot( const uint32_t mesh_index, const std::string& record )
{
  check( sg_rb3d_set_mesh_snapshot( m_ctx, mesh_index, record.data(), record.size() ), "sg_rb3d_set_mesh_snapshot" );
}

void GpuRigidBody3DBackend::setBodies( const std::vector<uint32_t>& geo_of_body, const std::vector<uint8_t>& fixed, const VectorXs& m, const VectorXs& I0 )
{
  m_nbodies = static_cast<unsigned>( geo_of_body.size() );
  check( sg_rb3d_set_bodies( m_ctx, m_nbodies, geo_of_body.data(), fixed.data(), m.data(), I0.data() ), "sg_rb3d_set_bodies" );
  m_m_updated = false; // a re-initialised state starts from the constructor's (transposed) inertia block again
  m_guard.setMasses( m.data(), m_nbodies, 1 );
}

void GpuRigidBody3DBackend::setGravity( const double gx, const double gy, const double gz )
{
  const double g[3] = { gx, gy, gz };
  check( sg_rb3d_set_gravity( m_ctx, g ), "sg_rb3d_set_gravity" );
  m_guard.setGravity( gx, gy, gz );
}

void GpuRigidBody3DBackend::setPlanes( const std::vector<double>& x, const std::vector<double>& n )
{
  check( sg_rb3d_set_planes( m_ctx, static_cast<uint32_t>( x.size() / 3 ), x.data(), n.data() ), "sg_rb3d_set_planes" );
}

void GpuRigidBody3DBackend::setCylinders( const std::vector<double>& x, const std::vector<double>& axis, const std::vector<double>& r )
{
  check( sg_rb3d_set_cylinders( m_ctx, static_cast<uint32_t>( r.size() ), x.data(), axis.data(), r.data() ), "sg_rb3d_set_cylinders" );
}

void GpuRigidBody3DBackend::setPortals( const std::vector<double>& plane_a_x, const std::vector<double>& plane_a_n, const std::vector<double>& plane_b_x, const std::vector<double>& plane_b_n,
                                        const std::vector<int32_t>& multiplier )
{
  check( sg_rb3d_set_portals( m_ctx, static_cast<uint32_t>( multiplier.size() / 3 ), plane_a_x.data(), plane_a_n.data(), plane_b_x.data(), plane_b_n.data(), multiplier.data() ), "sg_rb3d_set_portals" );
}

void GpuRigidBody3DBackend::enforcePeriodicBoundaryConditions( VectorXs& q )
{
  check( sg_rb3d_enforce_portals( m_ctx, q.data() ), "sg_rb3d_enforce_portals" );
}

void GpuRigidBody3DBackend::teleportedContacts( std::vector<GpuTeleportedContact3D>& teleported, uint64_t* num_regular )
{
  sg_teleported t;
  check( sg_rb3d_teleported( m_ctx, &t ), "sg_rb3d_teleported" );
  teleported.resize( t.n_teleported );
  for( uint64_t k = 0; k < t.n_teleported; ++k )
  {
    GpuTeleportedContact3D& o = teleported[k];
    o.portal0 = t.portal0[k]; o.portal1 = t.portal1[k];
    for( int c = 0; c < 3; ++c ) { o.x0[c] = t.x0[3 * k + c]; o.x1[c] = t.x1[3 * k + c]; }
  }
  if( num_regular != nullptr ) { *num_regular = t.n_regular; }
}

void GpuRigidBody3DBackend::flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  if( q1.size() != q0.size() ) { q1.resize( q0.size() ); }
  if( v1.size() != v0.size() ) { v1.resize( v0.size() ); }
  check( sg_rb3d_flow( m_ctx, map_kind | ( m_m_updated ? SG_MAP_M_UPDATED : 0 ), q0.data(), v0.data(), dt, q1.data(), v1.data() ), "sg_rb3d_flow" );
}

void GpuRigidBody3DBackend::updateMandMinv( const VectorXs& q, double* m_values, double* minv_values, const bool from_last_flow )
{
  check( sg_rb3d_update_m_and_minv( m_ctx, from_last_flow ? nullptr : q.data(), m_values, minv_values ), "sg_rb3d_update_m_and_minv" );
  m_m_updated = true;
}

void GpuRigidBody3DBackend::serializeState( std::ostream& output_stream, const bool from_last_flow ) const
{
  uint64_t bytes = 0;
  const int updated = ( from_last_flow || m_m_updated ) ? 1 : 0; // RigidBody3DSim::flow runs updateMandMinv after every map
  check( sg_rb3d_state_serialize( m_ctx, from_last_flow ? 1 : 0, updated, nullptr, 0, &bytes ), "sg_rb3d_state_serialize" );
  std::vector<char> buf( bytes );
  check( sg_rb3d_state_serialize( m_ctx, from_last_flow ? 1 : 0, updated, buf.data(), bytes, &bytes ), "sg_rb3d_state_serialize" );
  output_stream.write( buf.data(), static_cast<std::streamsize>( bytes ) );
}

void GpuRigidBody3DBackend::deserializeState( std::istream& input_stream, const bool from_running_simulation )
{
  std::istream::pos_type start;
  const std::vector<char> buf = sghReadRest( input_stream, start );
  check( sg_rb3d_state_deserialize( m_ctx, buf.data(), buf.size() ), "sg_rb3d_state_deserialize" );
  std::size_t consumed = 0;
  m_nbodies = m_guard.configureFromSnapshot( GravityOnlyGuard::RIGIDBODY3D, buf.data(), buf.size(), "GpuRigidBody3DBackend::deserializeState", &consumed );
  sghRewindBehindState( input_stream, start, consumed );
  m_m_updated = from_running_simulation;
}

void GpuRigidBody3DBackend::computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact3D>& contacts, uint64_t* num_candidates, const bool from_last_flow )
{
  sg_contacts c;
  check( sg_rb3d_active_set( m_ctx, q0.data(), q1.data(), SG_OUT_NORMALS | SG_OUT_POINTS | SG_OUT_DEPTHS | ( from_last_flow ? SG_IN_RESIDENT : 0u ), &c ), "sg_rb3d_active_set" );
  contacts.resize( c.n_active );
  for( uint64_t k = 0; k < c.n_active; ++k )
  {
    GpuContact3D& o = contacts[k];
    o.type = c.type[k]; o.i = c.i[k]; o.j = c.j[k]; o.aux = ( c.aux != nullptr ) ? c.aux[k] : 0u;
    for( int a = 0; a < 3; ++a ) { o.n[a] = c.n[3 * k + a]; o.p[a] = c.p[3 * k + a]; }
    o.depth = c.depth[k];
  }
  if( num_candidates != nullptr ) { *num_candidates = c.n_candidates; }
}

void GpuRigidBody3DBackend::getPotentialOverlaps( const std::vector<double>& aabbs, std::vector<std::pair<unsigned,unsigned>>& overlaps )
{
  sg_pairs p;
  check( sg_candidate_pairs( m_ctx, 3, static_cast<uint32_t>( aabbs.size() / 6 ), aabbs.data(), &p ), "sg_candidate_pairs" );
  overlaps.reserve( overlaps.size() + p.n );
  for( uint64_t k = 0; k < p.n; ++k ) { overlaps.emplace_back( p.ij[2 * k], p.ij[2 * k + 1] ); }
}

void GpuSplitHamMap::flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  m_backend.forceGuard().verify( fsys, q0, v0, iteration * dt, GravityOnlyGuard::RIGIDBODY3D, "GpuSplitHamMap" );
  m_backend.flow( SG_MAP_SPLIT_HAM, q0, v0, dt, q1, v1 );
}

void GpuDMVMap::flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  m_backend.forceGuard().verify( fsys, q0, v0, iteration * dt, GravityOnlyGuard::RIGIDBODY3D, "GpuDMVMap" );
  m_backend.flow( SG_MAP_DMV, q0, v0, dt, q1, v1 );
}

// ---- rigidbody2d ---------------------------------------------------------------------------------------------------
GpuRigidBody2DBackend::GpuRigidBody2DBackend( const int device )
: m_ctx( nullptr )
, m_nbodies( 0 )
{
  const int rc = sg_create( &m_ctx, device );
  if( rc != SG_OK )
  {
    std::cerr << "GpuRigidBody2DBackend: " << sg_last_error( nullptr ) << " Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
}

GpuRigidBody2DBackend::~GpuRigidBody2DBackend() { sg_destroy( m_ctx ); }

void GpuRigidBody2DBackend::check( const int rc, const char* what ) const
{
  if( rc != SG_OK )
  {
    std::cerr << what << ": " << sg_last_error( m_ctx ) << " Exiting." << std::endl;
    std::exit( EXIT_FAILURE );
  }
}

void GpuRigidBody2DBackend::setGeometry( const std::vector<uint32_t>& type, const std::vector<double>& r, const std::vector<double>& half_widths )
{
  check( sg_rb2d_set_geometry( m_ctx, static_cast<uint32_t>( type.size() ), type.data(), r.data(), half_widths.data() ), "sg_rb2d_set_geometry" );
}

void GpuRigidBody2DBackend::setBodies( const std::vector<uint32_t>& geo_of_body, const std::vector<uint8_t>& fixed, const VectorXs& M )
{
  m_nbodies = static_cast<unsigned>( geo_of_body.size() );
  check( sg_rb2d_set_bodies( m_ctx, m_nbodies, geo_of_body.data(), fixed.data(), M.data() ), "sg_rb2d_set_bodies" );
  m_guard.setMasses( M.data(), m_nbodies, 3 ); // M = diag( m, m, I ) per body
}

void GpuRigidBody2DBackend::setGravity( const double gx, const double gy )
{
  const double g[2] = { gx, gy };
  check( sg_rb2d_set_gravity( m_ctx, g ), "sg_rb2d_set_gravity" );
  m_guard.setGravity( gx, gy, 0.0 );
}

void GpuRigidBody2DBackend::setPlanes( const std::vector<double>& x, const std::vector<double>& n )
{
  check( sg_rb2d_set_planes( m_ctx, static_cast<uint32_t>( x.size() / 2 ), x.data(), n.data() ), "sg_rb2d_set_planes" );
}

void GpuRigidBody2DBackend::setPortals( const std::vector<double>& plane_a_x, const std::vector<double>& plane_a_n, const std::vector<double>& plane_b_x, const std::vector<double>& plane_b_n,
                                        const std::vector<double>& velocity, const std::vector<double>& bounds )
{
  check( sg_rb2d_set_portals( m_ctx, static_cast<uint32_t>( velocity.size() ), plane_a_x.data(), plane_a_n.data(), plane_b_x.data(), plane_b_n.data(), velocity.data(), bounds.data() ), "sg_rb2d_set_portals" );
}

void GpuRigidBody2DBackend::updatePeriodicBoundaryConditionsStartOfStep( const unsigned next_iteration, const scalar& dt )
{
  const scalar t{ next_iteration * dt }; // rigidbody2d/RigidBody2DSim.cpp:834
  check( sg_rb2d_update_portals( m_ctx, t, nullptr ), "sg_rb2d_update_portals" );
}

void GpuRigidBody2DBackend::enforcePeriodicBoundaryConditions( VectorXs& q, VectorXs& v )
{
  check( sg_rb2d_enforce_portals( m_ctx, q.data(), v.data() ), "sg_rb2d_enforce_portals" );
}

void GpuRigidBody2DBackend::teleportedContacts( std::vector<GpuTeleportedContact2D>& teleported, uint64_t* num_regular )
{
  sg_teleported t;
  check( sg_rb2d_teleported( m_ctx, &t ), "sg_rb2d_teleported" );
  teleported.resize( t.n_teleported );
  for( uint64_t k = 0; k < t.n_teleported; ++k )
  {
    GpuTeleportedContact2D& o = teleported[k];
    o.portal0 = t.portal0[k]; o.portal1 = t.portal1[k];
    for( int c = 0; c < 2; ++c )
    {
      o.x0[c] = t.x0[2 * k + c]; o.x1[c] = t.x1[2 * k + c]; o.kick[c] = t.kick[2 * k + c];
      o.delta0[c] = t.delta0[2 * k + c]; o.delta1[c] = t.delta1[2 * k + c];
    }
  }
  if( num_regular != nullptr ) { *num_regular = t.n_regular; }
}

void GpuRigidBody2DBackend::flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  if( q1.size() != q0.size() ) { q1.resize( q0.size() ); }
  if( v1.size() != v0.size() ) { v1.resize( v0.size() ); }
  check( sg_rb2d_flow( m_ctx, map_kind, q0.data(), v0.data(), dt, q1.data(), v1.data() ), "sg_rb2d_flow" );
}

void GpuRigidBody2DBackend::computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact2D>& contacts, uint64_t* num_candidates, const bool from_last_flow )
{
  sg_contacts c;
  check( sg_rb2d_active_set( m_ctx, q0.data(), q1.data(), SG_OUT_NORMALS | SG_OUT_POINTS | SG_OUT_DEPTHS | ( from_last_flow ? SG_IN_RESIDENT : 0u ), &c ), "sg_rb2d_active_set" );
  contacts.resize( c.n_active );
  for( uint64_t k = 0; k < c.n_active; ++k )
  {
    GpuContact2D& o = contacts[k];
    o.type = c.type[k]; o.i = c.i[k]; o.j = c.j[k];
    o.n[0] = c.n[2 * k]; o.n[1] = c.n[2 * k + 1];
    o.p[0] = c.p[2 * k]; o.p[1] = c.p[2 * k + 1];
    o.depth = c.depth[k];
  }
  if( num_candidates != nullptr ) { *num_candidates = c.n_candidates; }
}

void GpuRigidBody2DBackend::serializeState( std::ostream& output_stream, const bool from_last_flow ) const
{
  uint64_t bytes = 0;
  check( sg_rb2d_state_serialize( m_ctx, from_last_flow ? 1 : 0, nullptr, 0, &bytes ), "sg_rb2d_state_serialize" );
  std::vector<char> buf( bytes );
  check( sg_rb2d_state_serialize( m_ctx, from_last_flow ? 1 : 0, buf.data(), bytes, &bytes ), "sg_rb2d_state_serialize" );
  output_stream.write( buf.data(), static_cast<std::streamsize>( bytes ) );
}

void GpuRigidBody2DBackend::deserializeState( std::istream& input_stream )
{
  std::istream::pos_type start;
  const std::vector<char> buf = sghReadRest( input_stream, start );
  check( sg_rb2d_state_deserialize( m_ctx, buf.data(), buf.size() ), "sg_rb2d_state_deserialize" );
  std::size_t consumed = 0;
  m_nbodies = m_guard.configureFromSnapshot( GravityOnlyGuard::RIGIDBODY2D, buf.data(), buf.size(), "GpuRigidBody2DBackend::deserializeState", &consumed );
  sghRewindBehindState( input_stream, start, consumed );
}

void GpuRB2DSymplecticEulerMap::flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  m_backend.forceGuard().verify( fsys, q0, v0, iteration * dt, GravityOnlyGuard::RIGIDBODY2D, "GpuRB2DSymplecticEulerMap" );
  m_backend.flow( SG_MAP_SYMPLECTIC_EULER, q0, v0, dt, q1, v1 );
}

void GpuRB2DVerletMap::flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 )
{
  m_backend.forceGuard().verify( fsys, q0, v0, iteration * dt, GravityOnlyGuard::RIGIDBODY2D, "GpuRB2DVerletMap" );
  m_backend.flow( SG_MAP_VERLET, q0, v0, dt, q1, v1 );
}
