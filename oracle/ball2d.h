// oracle/ball2d.h
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). CPU restatement of the ball2d hot path:
//   ball2d/SymplecticEulerMap.cpp:21-32            flowSymplecticEuler
//   ball2d/VerletMap.cpp:15-41                     flowVerlet
//   ball2d/Forces/Ball2DGravityForce.cpp:36-46     gravity ( F.setZero(); F_i += m_i * g, Ball2DSim.cpp:72-78 )
//   ball2d/Ball2DState.cpp:54-66                   Minv = 1.0 / m
//   ball2d/Ball2DSim.cpp:151-173                   computeActiveSet (no portals: those are in ball2d_portals.h)
//   ball2d/Ball2DSim.cpp:553-608                   computeBallBallActiveSetSpatialGrid
//   ball2d/Ball2DSim.cpp:730-745                   computeBallDrumActiveSetAllPairs
//   ball2d/Ball2DSim.cpp:747-762                   computeBallPlaneActiveSetAllPairs
//   ball2d/Constraints/BallBallConstraint.cpp:22-37,219-222,270-280   normal, contact point, depth
//   ball2d/Constraints/BallStaticPlaneConstraint.cpp:10-15,171-174,220-223
//   ball2d/Constraints/BallStaticDrumConstraint.cpp:8-26,168-171      (depth: base-class NaN)
//   ball2d/StaticGeometry/StaticPlane.cpp:10-14    plane normal normalised at construction
// Parity: the only STORED reference outputs for this path are the CCD cases and the AABB fixtures (tests/test_oracle_golden.py)
// and config 1's two bundled scenes (tests/golden/ball2d_assets.npz).  Beyond them this file is pinned against the reference's own sources compiled
// unchanged into oracle/_ref (oracle/Makefile.ref, tests/test_oracle_vs_reference.py): both maps + the gravity force (q1, v1), the spatial grid, the CCD,
// and the three constraint classes -- isActive, normal, contact point, penetration depth, contact basis, evalgradg -- for every contact of its active
// sets, bit for bit.  The GLUE (loop order, which q each test reads, the portal branch) is pinned by running the reference's own simulation class:
// Ball2DSim.cpp + Ball2DState.cpp compiled unchanged (oracle/ref_shims/ref_ball2d_sim.cpp), Ball2DSim::computeActiveSet and Ball2DSim::flow on random
// scenes, the bundled scenes and portal / Lees-Edwards scenes over many steps == this file and ball2d_portals.h, element by element, bit for bit
// (tests/test_reference_sim_cpu.py).
#ifndef ORACLE_BALL2D_H
#define ORACLE_BALL2D_H

#include "broadphase.h"
#include "ccd.h"

#include <cstdint>
#include <limits>
#include <vector>

namespace orc
{

enum Ball2DContactType : uint32_t { BALL_BALL = 0, BALL_DRUM = 1, BALL_PLANE = 2 };

// One entry per Constraint the reference would emplace into active_set, in the reference's order.
struct Ball2DContact
{
  uint32_t type;  // Ball2DContactType
  uint32_t i;     // ball index (first ball for ball-ball)
  uint32_t j;     // second ball / drum index / plane index
  V2 n;           // constraint normal (ball-ball and drum: from q0; plane: plane normal)
  V2 p;           // world-space contact point getWorldSpaceContactPoint( q0 ): x0_i - r_i * n
  double depth;   // penetrationDepth( q1 ); NaN where the reference leaves the base-class default
};

struct Ball2DScene
{
  std::vector<double> r;          // N
  std::vector<double> m;          // N (per-ball mass; both DoFs of a ball share it)
  double g[2] = { 0.0, 0.0 };
  std::vector<V2> plane_x;        // S
  std::vector<V2> plane_n;        // S, ALREADY normalised (use makePlaneNormal)
  std::vector<V2> drum_x;
  std::vector<double> drum_r;
};

inline V2 makePlaneNormal( const V2& n ) { return normalized( n ); }

// kind: 0 = symplectic_euler, 1 = verlet
inline void flow( const int kind, const Ball2DScene& s, const double* q0, const double* v0, const double dt, double* q1, double* v1 )
{
  const std::size_t nb = s.r.size();
  for( std::size_t b = 0; b < nb; ++b )
  {
    const double minv = 1.0 / s.m[b];
    for( int k = 0; k < 2; ++k )
    {
      const std::size_t d = 2 * b + k;
      // F.setZero(); F += M * g
      const double F = 0.0 + s.m[b] * s.g[k];
      if( kind == 0 )
      {
        // v1 = v0 + dt * Minv * F   (sparse * dense accumulates into a zeroed temporary)
        const double tmp = 0.0 + ( dt * minv ) * F;
        v1[d] = v0[d] + tmp;
        q1[d] = q0[d] + dt * v1[d];
      }
      else
      {
        const double tmp0 = 0.0 + ( ( 0.5 * dt ) * minv ) * F;
        double vh = v0[d] + tmp0;
        q1[d] = q0[d] + dt * vh;
        // F( q1 ): gravity does not depend on q
        const double F1 = 0.0 + s.m[b] * s.g[k];
        // v1 += 0.5 * dt * Minv * F accumulates the product straight into v1
        vh = vh + ( ( 0.5 * dt ) * minv ) * F1;
        v1[d] = vh;
      }
    }
  }
}

inline void buildAABBs( const Ball2DScene& s, const double* q0, const double* q1, std::vector<Box<2>>& aabbs )
{
  const std::size_t nb = s.r.size();
  aabbs.resize( nb );
  for( std::size_t b = 0; b < nb; ++b )
  {
    for( int k = 0; k < 2; ++k )
    {
      const double mn = std::min( q1[2 * b + k], q0[2 * b + k] );
      const double mx = std::max( q1[2 * b + k], q0[2 * b + k] );
      aabbs[b].lo[k] = mn - s.r[b];
      aabbs[b].hi[k] = mx + s.r[b];
    }
  }
}

inline Ball2DContact makeBallBall( const uint32_t i, const uint32_t j, const double* q0, const double* q1, const double ri, const double rj )
{
  Ball2DContact c;
  c.type = BALL_BALL; c.i = i; c.j = j;
  const V2 q0a{ q0[2 * i], q0[2 * i + 1] };
  const V2 q0b{ q0[2 * j], q0[2 * j + 1] };
  c.n = normalized( q0a - q0b );
  const V2 q1a{ q1[2 * i], q1[2 * i + 1] };
  const V2 q1b{ q1[2 * j], q1[2 * j + 1] };
  c.p = q0a - ri * c.n;
  c.depth = std::min( 0.0, norm( q1a - q1b ) - ( ri + rj ) );
  return c;
}

// Drums then planes, appended to active_set (ball2d/Ball2DSim.cpp:730-762)
inline void computeStaticActiveSet( const Ball2DScene& s, const double* q0, const double* q1, std::vector<Ball2DContact>& active_set )
{
  const uint32_t nb = uint32_t( s.r.size() );
  // Drums: drum-major, ball ascending
  for( uint32_t d = 0; d < uint32_t( s.drum_x.size() ); ++d )
  {
    for( uint32_t b = 0; b < nb; ++b )
    {
      const V2 x1{ q1[2 * b], q1[2 * b + 1] };
      const double Rr = s.drum_r[d] - s.r[b];
      if( squaredNorm( s.drum_x[d] - x1 ) >= Rr * Rr )
      {
        Ball2DContact c;
        c.type = BALL_DRUM; c.i = b; c.j = d;
        const V2 x0{ q0[2 * b], q0[2 * b + 1] };
        c.n = normalized( s.drum_x[d] - x0 );
        c.p = x0 - s.r[b] * c.n;
        c.depth = std::numeric_limits<double>::quiet_NaN();
        active_set.emplace_back( c );
      }
    }
  }
  // Planes: plane-major, ball ascending
  for( uint32_t p = 0; p < uint32_t( s.plane_x.size() ); ++p )
  {
    for( uint32_t b = 0; b < nb; ++b )
    {
      const V2 x1{ q1[2 * b], q1[2 * b + 1] };
      const double d = dot( s.plane_n[p], x1 - s.plane_x[p] );
      if( d <= s.r[b] )
      {
        Ball2DContact c;
        c.type = BALL_PLANE; c.i = b; c.j = p;
        c.n = s.plane_n[p];
        const V2 x0{ q0[2 * b], q0[2 * b + 1] };
        c.p = x0 - s.r[b] * c.n;
        c.depth = std::min( 0.0, d - s.r[b] );
        active_set.emplace_back( c );
      }
    }
  }
}

// candidates_out (optional): the broad-phase pair set, ascending (i,j)
// use_grid: true = literal std::map/std::set spatial grid, false = all pairs (same set, F3)
inline void computeActiveSet( const Ball2DScene& s, const double* q0, const double* q1, std::vector<Ball2DContact>& active_set,
                              std::vector<std::pair<unsigned,unsigned>>* candidates_out = nullptr, const bool use_grid = true )
{
  const uint32_t nb = uint32_t( s.r.size() );
  active_set.clear();
  // Ball-ball
  if( nb > 0 )
  {
    PairSet possible_overlaps;
    {
      std::vector<Box<2>> aabbs;
      buildAABBs( s, q0, q1, aabbs );
      if( use_grid ) { getPotentialOverlaps<2>( aabbs, possible_overlaps ); }
      else { getPotentialOverlapsAllPairs<2>( aabbs, possible_overlaps ); }
    }
    if( candidates_out != nullptr ) { candidates_out->assign( possible_overlaps.begin(), possible_overlaps.end() ); }
    for( const auto& pr : possible_overlaps )
    {
      const uint32_t a = pr.first;
      const uint32_t b = pr.second;
      const V2 q0a{ q0[2 * a], q0[2 * a + 1] };
      const V2 q1a{ q1[2 * a], q1[2 * a + 1] };
      const V2 q0b{ q0[2 * b], q0[2 * b + 1] };
      const V2 q1b{ q1[2 * b], q1[2 * b + 1] };
      const std::pair<bool,double> ccd = ballBallCCDCollisionHappens( q0a, q1a, s.r[a], q0b, q1b, s.r[b] );
      if( ccd.first ) { active_set.emplace_back( makeBallBall( a, b, q0, q1, s.r[a], s.r[b] ) ); }
    }
  }
  computeStaticActiveSet( s, q0, q1, active_set );
}

}

#endif
