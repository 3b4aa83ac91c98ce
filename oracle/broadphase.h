// oracle/broadphase.h
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). CPU restatement of the reference's
// spatial-grid broad phase, keeping its data structures (std::map of voxels, std::set of pairs) so that
// timing it is representative of the reference's cost:
//   ball2d/SpatialGridDetector.cpp:15-26   AABB::overlaps            -> Box<D>::overlaps
//   ball2d/SpatialGridDetector.cpp:29-40   computeCellIndex/keyForIndex
//   ball2d/SpatialGridDetector.cpp:42-74   rasterizeAABBs
//   ball2d/SpatialGridDetector.cpp:76-104  initializeSpatialGrid
//   ball2d/SpatialGridDetector.cpp:106-133 getPotentialOverlaps
//   ball2d/SpatialGridDetector.cpp:135-147 getPotentialOverlapsAllPairs
// rigidbody2d/SpatialGrid.cpp and rigidbody3d/SpatialGridDetector.cpp are the same code with D = 2 / 3
// (3-D key: x + dimx*y + dimx*dimy*z, rigidbody3d/SpatialGridDetector.cpp:36-41), hence one template.
// Parity pin: ball2dtests/collision_detection_tests.cpp (grid set == all-pairs set on 3 fixtures).
#ifndef ORACLE_BROADPHASE_H
#define ORACLE_BROADPHASE_H

#include <cmath>
#include <limits>
#include <map>
#include <set>
#include <utility>
#include <vector>

namespace orc
{

template<int D>
struct Box
{
  double lo[D];
  double hi[D];
  // Separating axis with strict '<': touching boxes overlap
  bool overlaps( const Box& o ) const
  {
    for( int k = 0; k < D; ++k ) { if( hi[k] < o.lo[k] ) { return false; } }
    for( int k = 0; k < D; ++k ) { if( o.hi[k] < lo[k] ) { return false; } }
    return true;
  }
};

using PairSet = std::set<std::pair<unsigned,unsigned>>;

template<int D>
void getPotentialOverlaps( const std::vector<Box<D>>& aabbs, PairSet& overlaps )
{
  // initializeSpatialGrid
  double min_coord[D];
  double max_coord[D];
  for( int k = 0; k < D; ++k ) { min_coord[k] = std::numeric_limits<double>::infinity(); max_coord[k] = -std::numeric_limits<double>::infinity(); }
  for( const Box<D>& b : aabbs )
  {
    for( int k = 0; k < D; ++k )
    {
      min_coord[k] = std::min( min_coord[k], b.lo[k] );
      max_coord[k] = std::max( max_coord[k], b.hi[k] );
    }
  }
  for( int k = 0; k < D; ++k ) { min_coord[k] -= 2.0e-6; max_coord[k] += 2.0e-6; }
  double h;
  {
    double delta[D];
    for( int k = 0; k < D; ++k ) { delta[k] = 0.0; }
    for( const Box<D>& b : aabbs ) { for( int k = 0; k < D; ++k ) { delta[k] += b.hi[k] - b.lo[k]; } }
    double mx = delta[0];
    for( int k = 1; k < D; ++k ) { mx = std::max( mx, delta[k] ); }
    h = mx / double( aabbs.size() );
  }
  unsigned dimensions[D];
  for( int k = 0; k < D; ++k ) { dimensions[k] = unsigned( std::ceil( ( max_coord[k] - min_coord[k] ) / h ) ); }

  // rasterizeAABBs
  std::map<unsigned,std::vector<unsigned>> voxels;
  for( std::size_t aabb_idx = 0; aabb_idx < aabbs.size(); ++aabb_idx )
  {
    unsigned index_lower[3] = { 0, 0, 0 };
    unsigned index_upper[3] = { 0, 0, 0 };
    for( int k = 0; k < D; ++k )
    {
      index_lower[k] = unsigned( ( ( aabbs[aabb_idx].lo[k] - 1.0e-6 ) - min_coord[k] ) / h );
      index_upper[k] = unsigned( ( ( aabbs[aabb_idx].hi[k] + 1.0e-6 ) - min_coord[k] ) / h );
    }
    for( unsigned x_idx = index_lower[0]; x_idx <= index_upper[0]; ++x_idx )
    {
      for( unsigned y_idx = index_lower[1]; y_idx <= index_upper[1]; ++y_idx )
      {
        for( unsigned z_idx = index_lower[2]; z_idx <= index_upper[2]; ++z_idx )
        {
          // 32-bit unsigned arithmetic; may wrap exactly as in the reference
          unsigned key = x_idx + dimensions[0] * y_idx;
          if( D == 3 ) { key += dimensions[0] * dimensions[1] * z_idx; }
          auto voxel_iterator = voxels.find( key );
          if( voxel_iterator == voxels.end() )
          {
            voxel_iterator = voxels.insert( std::make_pair( key, std::vector<unsigned>{} ) ).first;
          }
          voxel_iterator->second.emplace_back( unsigned( aabb_idx ) );
        }
      }
    }
  }

  // Pair loop over each voxel
  for( auto itr = voxels.cbegin(); itr != voxels.cend(); ++itr )
  {
    const std::vector<unsigned>& cell = itr->second;
    for( std::size_t idx0 = 0; idx0 + 1 < cell.size(); ++idx0 )
    {
      for( std::size_t idx1 = idx0 + 1; idx1 < cell.size(); ++idx1 )
      {
        if( aabbs[cell[idx0]].overlaps( aabbs[cell[idx1]] ) )
        {
          overlaps.insert( std::make_pair( cell[idx0], cell[idx1] ) );
        }
      }
    }
  }
}

template<int D>
void getPotentialOverlapsAllPairs( const std::vector<Box<D>>& aabbs, PairSet& overlaps )
{
  for( std::size_t idx0 = 0; idx0 < aabbs.size(); ++idx0 )
  {
    for( std::size_t idx1 = idx0 + 1; idx1 < aabbs.size(); ++idx1 )
    {
      if( aabbs[idx0].overlaps( aabbs[idx1] ) )
      {
        overlaps.insert( std::make_pair( unsigned( idx0 ), unsigned( idx1 ) ) );
      }
    }
  }
}

// Independent truth for large N where O(N^2) is too slow: sweep-and-prune along x, same predicate.
// Not in the reference; used only to cross-check the literal grid algorithm and the GPU at ~1e6 bodies.
template<int D>
void getPotentialOverlapsSweep( const std::vector<Box<D>>& aabbs, std::vector<std::pair<unsigned,unsigned>>& overlaps );

}

#endif
