// oracle/assembly2d.h -- TEST INFRASTRUCTURE.  CPU restatement of what the impact maps build first from a ball2d active set:
//   computeN            scisim/ConstrainedMaps/ImpactMaps/ImpactOperatorUtilities.cpp:10-48 with the constraints' evalgradg
//                       ( ball2d/Constraints/BallBallConstraint.cpp:88-100, BallStaticPlaneConstraint.cpp:65-73, BallStaticDrumConstraint.cpp:59-66 ),
//                       followed by N.prune( value != 0 )
//   Q = N^T Minv N      scisim/ConstrainedMaps/ImpactMaps/ImpactMap.cpp:106-110, evaluated the way Eigen 3.3.4 evaluates ( N^T * Minv ) * N for
//                       compressed column-major operands ( SparseCore/ConservativeSparseSparseProduct.h ): the inner product first --
//                       ( N^T Minv )( c, k ) = N( k, c ) * minv_k --, then per column d of N, over its coefficients in ascending row k, over the
//                       coefficients ( c, x ) of column k of the left factor: values[c] = x * N( k, d ) the first time c is met, += afterwards;
//                       the column's rows are sorted at the end
//   computeContactBases ball2d/Ball2DSim.cpp:188-201 with BallBallConstraint.cpp:244-252 ( and the plane / drum twins ): [ n | ( -n.y, n.x ) ]
//   ConstraintCache     ball2d/ConstraintCache.cpp:20-122: three std::map keyed ( i, j ), ( plane, ball ), ( drum, ball ); a miss gives zero
// PARITY UNPINNED by stored values (the reference has no test for any of this); the sparse-product order is Eigen's as read from its source.
#ifndef ORACLE_ASSEMBLY2D_H
#define ORACLE_ASSEMBLY2D_H

#include "ball2d.h"

#include <algorithm>
#include <map>
#include <vector>

namespace orc
{

struct Assembly2D
{
  std::vector<int> n_outer, n_inner, q_outer, q_inner;
  std::vector<double> n_val, q_val, bases;
  bool supported = true;
};

inline void assemble2d( const Ball2DScene& s, const std::vector<Ball2DContact>& act, Assembly2D& out )
{
  const std::size_t nc = act.size(), ndof = 2 * s.r.size();
  out = Assembly2D();
  out.n_outer.assign( 1, 0 );
  // ---- N, column by column, zeros pruned
  for( std::size_t c = 0; c < nc; ++c )
  {
    const Ball2DContact& k = act[c];
    if( k.type > 2u ) { out.supported = false; }
    auto put = [&]( const unsigned row, const double v ) { if( v != 0.0 ) { out.n_inner.push_back( int( row ) ); out.n_val.push_back( v ); } };
    put( 2u * k.i, k.n.x ); put( 2u * k.i + 1u, k.n.y );
    if( k.type == 0u ) { put( 2u * k.j, -k.n.x ); put( 2u * k.j + 1u, -k.n.y ); }
    out.n_outer.push_back( int( out.n_inner.size() ) );
    out.bases.push_back( k.n.x ); out.bases.push_back( k.n.y ); out.bases.push_back( -k.n.y ); out.bases.push_back( k.n.x );
  }
  // ---- left factor N^T * Minv, stored by column k: ( c, N( k, c ) * minv_k ), c ascending
  std::vector<std::vector<std::pair<int,double>>> left( ndof );
  for( std::size_t c = 0; c < nc; ++c )
  {
    for( int e = out.n_outer[c]; e < out.n_outer[c + 1]; ++e )
    {
      const int k = out.n_inner[e];
      const double minv = 1.0 / s.m[std::size_t( k ) / 2];
      left[std::size_t( k )].emplace_back( int( c ), out.n_val[e] * minv );
    }
  }
  // ---- Q column by column
  out.q_outer.assign( 1, 0 );
  std::vector<char> mask( nc, 0 );
  std::vector<double> values( nc, 0.0 );
  std::vector<int> indices;
  for( std::size_t d = 0; d < nc; ++d )
  {
    indices.clear();
    for( int e = out.n_outer[d]; e < out.n_outer[d + 1]; ++e )
    {
      const double y = out.n_val[e];
      for( const std::pair<int,double>& lc : left[std::size_t( out.n_inner[e] )] )
      {
        const double x = lc.second;
        if( !mask[std::size_t( lc.first )] ) { mask[std::size_t( lc.first )] = 1; values[std::size_t( lc.first )] = x * y; indices.push_back( lc.first ); }
        else { values[std::size_t( lc.first )] += x * y; }
      }
    }
    std::sort( indices.begin(), indices.end() );
    for( const int c : indices ) { out.q_inner.push_back( c ); out.q_val.push_back( values[std::size_t( c )] ); mask[std::size_t( c )] = 0; }
    out.q_outer.push_back( int( out.q_inner.size() ) );
  }
}

// ball2d/ConstraintCache.cpp
struct ConstraintCache2D
{
  std::map<std::pair<unsigned,unsigned>, std::vector<double>> ball_ball, plane_ball, drum_ball;
  void clear() { ball_ball.clear(); plane_ball.clear(); drum_ball.clear(); }
  void cache( const Ball2DContact& k, const double* r, const unsigned ncomp )
  {
    std::vector<double> v( r, r + ncomp );
    if( k.type == 0u ) { ball_ball.insert( std::make_pair( std::make_pair( k.i, k.j ), v ) ); }
    else if( k.type == 2u ) { plane_ball.insert( std::make_pair( std::make_pair( k.j, k.i ), v ) ); }
    else { drum_ball.insert( std::make_pair( std::make_pair( k.j, k.i ), v ) ); }
  }
  bool get( const Ball2DContact& k, double* r, const unsigned ncomp ) const
  {
    const auto& m = ( k.type == 0u ) ? ball_ball : ( k.type == 2u ? plane_ball : drum_ball );
    const auto it = m.find( k.type == 0u ? std::make_pair( k.i, k.j ) : std::make_pair( k.j, k.i ) );
    if( it != m.end() ) { for( unsigned q = 0; q < ncomp; ++q ) { r[q] = it->second[q]; } return true; }
    for( unsigned q = 0; q < ncomp; ++q ) { r[q] = 0.0; }
    return false;
  }
};

}

#endif
