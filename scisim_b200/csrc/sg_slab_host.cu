// sg_slab_host.cu -- host-side helpers of the multi-GPU slab decomposition (hard part H7): the
// equal-count x-quantile partition of an arbitrarily numbered scene, the x-range a slab's bodies must stay inside,
// and the merge of per-slab lists into the reference's order.  No device work, no context: plain functions over
// host arrays, used by sg_multi (one process, N GPUs) and by scisim_b200/slab.py (one process per GPU).
//
// The reference has no distributed mode.  The order contract the merge restores is the iteration order of the
// std::set the reference's broad phase fills -- ascending (i,j), ball2d/Ball2DSim.cpp:580 -- followed by the static
// contacts geometry-major, body ascending (Ball2DSim.cpp:735-761).
#include "../../include/scisim_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace
{

struct ByX
{
  const double* x;
  uint32_t stride;
  bool operator()( const uint32_t a, const uint32_t b ) const
  {
    const double xa = x[size_t( a ) * stride], xb = x[size_t( b ) * stride];
    return ( xa < xb ) || ( xa == xb && a < b );
  }
};

// splits order[lo, hi) (which will hold ranks [r0, r1)) at the quantile positions, recursively: O(n log W)
void quantile_split( uint32_t* order, const uint64_t n, const uint32_t world, const uint64_t lo, const uint64_t hi, const uint32_t r0, const uint32_t r1, const ByX& cmp )
{
  if( r1 - r0 <= 1u || hi <= lo ) { return; }
  const uint32_t rm = r0 + ( r1 - r0 ) / 2u;
  const uint64_t cut = n * uint64_t( rm ) / world; // first sorted position of rank rm
  if( cut > lo && cut < hi ) { std::nth_element( order + lo, order + cut, order + hi, cmp ); }
  quantile_split( order, n, world, lo, cut < lo ? lo : ( cut > hi ? hi : cut ), r0, rm, cmp );
  quantile_split( order, n, world, cut < lo ? lo : ( cut > hi ? hi : cut ), hi, rm, r1, cmp );
}

} // namespace

extern "C"
{

int sg_slab_partition( uint32_t n, const double* x, uint32_t x_stride, uint32_t world, uint32_t* rank_of, double* cuts )
{
  if( world == 0u || x_stride == 0u || ( n > 0u && ( x == nullptr || rank_of == nullptr ) ) ) { return SG_ERR_INVALID; }
  std::vector<uint32_t> order( n );
  std::iota( order.begin(), order.end(), 0u );
  const ByX cmp{ x, x_stride };
  quantile_split( order.data(), n, world, 0, n, 0u, world, cmp );
  // rank k owns sorted positions [k n / W, (k+1) n / W)
  std::vector<double> first_x( world, std::numeric_limits<double>::quiet_NaN() ), last_x( world, std::numeric_limits<double>::quiet_NaN() );
  double xmin = std::numeric_limits<double>::infinity(), xmax = -std::numeric_limits<double>::infinity();
  for( uint32_t k = 0; k < world; ++k )
  {
    const uint64_t b = uint64_t( n ) * k / world, e = uint64_t( n ) * ( k + 1u ) / world;
    double mn = std::numeric_limits<double>::infinity(), mx = -std::numeric_limits<double>::infinity();
    for( uint64_t p = b; p < e; ++p )
    {
      const uint32_t i = order[p];
      rank_of[i] = k;
      const double xi = x[size_t( i ) * x_stride];
      mn = std::min( mn, xi ); mx = std::max( mx, xi );
    }
    if( e > b ) { first_x[k] = mn; last_x[k] = mx; xmin = std::min( xmin, mn ); xmax = std::max( xmax, mx ); }
  }
  if( cuts != nullptr )
  {
    // cuts[k] separates rank k-1 from rank k (midway between their nearest bodies); the outer two are the scene's extent
    cuts[0] = ( n > 0u ) ? xmin : 0.0;
    cuts[world] = ( n > 0u ) ? xmax : 0.0;
    double prev_last = cuts[0];
    for( uint32_t k = 0; k < world; ++k )
    {
      if( k > 0u )
      {
        // an empty rank takes a zero-width slab at the last cut
        const double f = std::isnan( first_x[k] ) ? prev_last : first_x[k];
        cuts[k] = 0.5 * ( prev_last + f );
      }
      if( !std::isnan( last_x[k] ) ) { prev_last = last_x[k]; }
    }
  }
  return SG_OK;
}

int sg_slab_limits( uint32_t world, const double* cuts, uint32_t rank, double* limits )
{
  if( world == 0u || cuts == nullptr || limits == nullptr || rank >= world ) { return SG_ERR_INVALID; }
  // Bodies of rank k stay strictly inside ( cuts[k] - w(k-1)/2, cuts[k+1] + w(k+1)/2 ), w(j) = width of slab j: then the
  // boxes of ranks k and k+2 are separated by the midline of slab k+1, so only neighbouring slabs can touch.
  const double inf = std::numeric_limits<double>::infinity();
  limits[0] = ( rank == 0u ) ? -inf : cuts[rank] - 0.5 * ( cuts[rank] - cuts[rank - 1u] );
  limits[1] = ( rank + 1u == world ) ? inf : cuts[rank + 1u] + 0.5 * ( cuts[rank + 2u] - cuts[rank + 1u] );
  return SG_OK;
}

// Destination of every entry of n_parts lists in their merge.  Part k has len[k] entries; entry e has first index
// first[k][e * stride], lists are ascending in it and every body's entries sit in ONE part (the rank that owns it), so
// the merged list is, body by body, that part's run -- in its order.  dest[k][e] = position of the entry in the merge.
int sg_slab_merge_dest( uint32_t n_bodies, uint32_t n_parts, const uint32_t* const* first, uint32_t stride, const uint64_t* len, uint64_t* const* dest )
{
  if( n_parts == 0u ) { return SG_OK; }
  if( first == nullptr || len == nullptr || dest == nullptr || stride == 0u ) { return SG_ERR_INVALID; }
  std::vector<uint64_t> start( size_t( n_bodies ) + 1u, 0ull );
  std::vector<int> rc( n_parts, SG_OK );
  // run lengths per body (parts touch disjoint bodies: no synchronisation needed)
  auto count_part = [&]( const uint32_t k )
  {
    const uint32_t* f = first[k];
    uint32_t prev = 0u;
    for( uint64_t e = 0; e < len[k]; ++e )
    {
      const uint32_t i = f[e * stride];
      if( i >= n_bodies || ( e > 0 && i < prev ) ) { rc[k] = SG_ERR_INVALID; return; }
      ++start[i];
      prev = i;
    }
  };
  {
    std::vector<std::thread> th;
    for( uint32_t k = 1; k < n_parts; ++k ) { th.emplace_back( count_part, k ); }
    count_part( 0u );
    for( auto& t : th ) { t.join(); }
  }
  for( uint32_t k = 0; k < n_parts; ++k ) { if( rc[k] != SG_OK ) { return rc[k]; } }
  uint64_t acc = 0;
  for( size_t i = 0; i <= n_bodies; ++i ) { const uint64_t c = start[i]; start[i] = acc; acc += c; }
  auto place_part = [&]( const uint32_t k )
  {
    const uint32_t* f = first[k];
    uint64_t* d = dest[k];
    uint64_t run_begin = 0;
    for( uint64_t e = 0; e < len[k]; ++e )
    {
      const uint32_t i = f[e * stride];
      if( e > 0 && i != f[( e - 1 ) * stride] ) { run_begin = e; }
      d[e] = start[i] + ( e - run_begin );
    }
  };
  {
    std::vector<std::thread> th;
    for( uint32_t k = 1; k < n_parts; ++k ) { th.emplace_back( place_part, k ); }
    place_part( 0u );
    for( auto& t : th ) { t.join(); }
  }
  return SG_OK;
}

// Static contacts (drums, planes, ...): every part lists them type-major (ascending type code), geometry-major, body
// ascending; so does the merge.  key = ( type, j, i ).  These lists are short (bodies near walls).
int sg_slab_merge_static_dest( uint32_t n_parts, const uint32_t* const* type, const uint32_t* const* i, const uint32_t* const* j, const uint64_t* len, uint64_t* const* dest )
{
  if( n_parts == 0u ) { return SG_OK; }
  if( type == nullptr || i == nullptr || j == nullptr || len == nullptr || dest == nullptr ) { return SG_ERR_INVALID; }
  struct Ent { uint32_t t, j, i, part; uint64_t e; };
  std::vector<Ent> all;
  uint64_t total = 0;
  for( uint32_t k = 0; k < n_parts; ++k ) { total += len[k]; }
  all.reserve( total );
  for( uint32_t k = 0; k < n_parts; ++k ) { for( uint64_t e = 0; e < len[k]; ++e ) { all.push_back( Ent{ type[k][e], j[k][e], i[k][e], k, e } ); } }
  std::stable_sort( all.begin(), all.end(), []( const Ent& a, const Ent& b )
  {
    if( a.t != b.t ) { return a.t < b.t; }
    if( a.j != b.j ) { return a.j < b.j; }
    return a.i < b.i;
  } );
  for( uint64_t p = 0; p < all.size(); ++p ) { dest[all[p].part][all[p].e] = p; }
  return SG_OK;
}

}
