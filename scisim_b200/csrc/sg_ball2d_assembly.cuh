// sg_ball2d_assembly.cuh -- SURVEY.md 8(f2): what the impact maps build first from an active set, assembled on the device from the contact
// arrays the narrow phase left there, so that the solver's inputs need not be re-derived on the host from a downloaded constraint list:
//   N       ImpactOperatorUtilities::computeN (scisim/ConstrainedMaps/ImpactMaps/ImpactOperatorUtilities.cpp:10-48): column c = the
//           constraint's evalgradg -- ball-ball: ( 2i, n.x ) ( 2i+1, n.y ) ( 2j, -n.x ) ( 2j+1, -n.y ) (BallBallConstraint.cpp:88-100), plane / drum:
//           ( 2i, n.x ) ( 2i+1, n.y ) (BallStaticPlaneConstraint.cpp:65-73, BallStaticDrumConstraint.cpp:59-66) -- then N.prune( value != 0 )
//   Q       N^T * Minv * N (ImpactMap.cpp:106-110), compressed column-major with sorted rows; every entry accumulated over the shared
//           degrees of freedom in ascending order, ( N( k, c ) * minv_k ) * N( k, d ), the first term assigned and the rest added, as Eigen's
//           conservative sparse product does; an entry exists only where both columns have a (non-pruned) coefficient on a shared row
//   bases   computeContactBases (ball2d/Ball2DSim.cpp:188-201): per contact the 2x2 [ n | t ], t = ( -n.y, n.x ) (BallBallConstraint.cpp:244-252)
//   cache   ConstraintCache (ball2d/ConstraintCache.cpp:20-122) as a sorted-key join: the keys of an active set ARE sorted (ball-ball by ( i, j ),
//           drums by ( drum, ball ), planes by ( plane, ball )), so lookup is a binary search per contact, a miss gives 0
// Included at the end of sg_ball2d.cu (needs Ball2DData).  Contact types 0..2 only; teleported contacts return SG_ERR_UNSUPPORTED.
#ifndef SG_BALL2D_ASSEMBLY_CUH
#define SG_BALL2D_ASSEMBLY_CUH

struct AsmData
{
  DevBuf ncnt, nouter, ninner, nval;      // N
  DevBuf deg, inc_start, inc_cursor, inc; // body -> incident contacts (sorted ascending inside a body)
  DevBuf qcnt, qouter, qinner, qval;      // Q
  DevBuf bases;
  DevBuf partials, total;
  DevBuf bad;
  PinBuf host;
  // cache
  DevBuf c_type, c_i, c_j, c_r, lookup;
  uint64_t c_n = 0, c_nbb = 0, c_ndrum = 0;
  uint32_t c_ncomp = 0;
  void release()
  {
    DevBuf* b[] = { &ncnt, &nouter, &ninner, &nval, &deg, &inc_start, &inc_cursor, &inc, &qcnt, &qouter, &qinner, &qval, &bases, &partials, &total, &bad, &c_type, &c_i, &c_j, &c_r, &lookup };
    for( DevBuf* x : b ) { x->release(); }
    host.release();
  }
};

// the two (or one) bodies of contact c and its coefficient on row 2b + axis: s * n[axis]
struct AsmContact { uint32_t type, i, j; double nx, ny; };
__device__ __forceinline__ AsmContact asm_load( const ContactOut2D& a, const unsigned long long c )
{
  AsmContact k;
  k.type = a.type[c]; k.i = a.i[c]; k.j = a.j[c];
  const double2 n = a.n[c];
  k.nx = n.x; k.ny = n.y;
  return k;
}
__device__ __forceinline__ bool asm_has_body( const AsmContact& k, const uint32_t b ) { return k.i == b || ( k.type == SG_BALL_BALL && k.j == b ); }
__device__ __forceinline__ double asm_coeff( const AsmContact& k, const uint32_t b, const int axis )
{
  const double v = axis == 0 ? k.nx : k.ny;
  return ( k.type == SG_BALL_BALL && k.j == b && k.i != b ) ? -v : v;
}

__global__ void __launch_bounds__( 256 ) k_asm_count( const unsigned long long nc, const ContactOut2D a, uint32_t* __restrict__ ncnt, uint32_t* __restrict__ deg, uint32_t* __restrict__ bad )
{
  const unsigned long long c = blockIdx.x * uint64_t( blockDim.x ) + threadIdx.x;
  if( c >= nc ) { return; }
  const AsmContact k = asm_load( a, c );
  if( k.type > SG_BALL_PLANE ) { *bad = 1u; ncnt[c] = 0u; return; }
  const uint32_t nz = ( k.nx != 0.0 ? 1u : 0u ) + ( k.ny != 0.0 ? 1u : 0u );
  ncnt[c] = ( k.type == SG_BALL_BALL ) ? 2u * nz : nz;
  atomicAdd( &deg[k.i], 1u );
  if( k.type == SG_BALL_BALL ) { atomicAdd( &deg[k.j], 1u ); }
}

__global__ void __launch_bounds__( 256 ) k_asm_n_emit( const unsigned long long nc, const ContactOut2D a, const uint32_t* __restrict__ nouter, int32_t* __restrict__ ninner, double* __restrict__ nval,
                                                      const uint32_t* __restrict__ inc_start, uint32_t* __restrict__ inc_cursor, uint32_t* __restrict__ inc, double* __restrict__ bases )
{
  const unsigned long long c = blockIdx.x * uint64_t( blockDim.x ) + threadIdx.x;
  if( c >= nc ) { return; }
  const AsmContact k = asm_load( a, c );
  if( k.type > SG_BALL_PLANE ) { return; }
  uint32_t o = nouter[c];
  if( k.nx != 0.0 ) { ninner[o] = int32_t( 2u * k.i ); nval[o] = k.nx; ++o; }
  if( k.ny != 0.0 ) { ninner[o] = int32_t( 2u * k.i + 1u ); nval[o] = k.ny; ++o; }
  if( k.type == SG_BALL_BALL )
  {
    if( k.nx != 0.0 ) { ninner[o] = int32_t( 2u * k.j ); nval[o] = -k.nx; ++o; }
    if( k.ny != 0.0 ) { ninner[o] = int32_t( 2u * k.j + 1u ); nval[o] = -k.ny; ++o; }
  }
  inc[inc_start[k.i] + atomicAdd( &inc_cursor[k.i], 1u )] = uint32_t( c );
  if( k.type == SG_BALL_BALL ) { inc[inc_start[k.j] + atomicAdd( &inc_cursor[k.j], 1u )] = uint32_t( c ); }
  if( bases != nullptr ) { bases[4 * c] = k.nx; bases[4 * c + 1] = k.ny; bases[4 * c + 2] = -k.ny; bases[4 * c + 3] = k.nx; }
}

// a body's incident contacts in ascending order (they arrive in whatever order the atomics gave)
__global__ void __launch_bounds__( 256 ) k_asm_sort_incidence( const uint32_t nbodies, const uint32_t* __restrict__ inc_start, uint32_t* __restrict__ inc )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= nbodies ) { return; }
  const uint32_t s = inc_start[b], e = inc_start[b + 1];
  for( uint32_t x = s + 1u; x < e; ++x )
  {
    const uint32_t v = inc[x];
    uint32_t y = x;
    while( y > s && inc[y - 1u] > v ) { inc[y] = inc[y - 1u]; --y; }
    inc[y] = v;
  }
}

// column d of Q: the contacts that share a body with d, ascending; EMIT = false counts the structural entries
template<bool EMIT>
__global__ void __launch_bounds__( 256 ) k_asm_q( const unsigned long long nc, const ContactOut2D a, const double* __restrict__ mass, const uint32_t* __restrict__ inc_start, const uint32_t* __restrict__ inc,
                                                 uint32_t* __restrict__ qcnt, const uint32_t* __restrict__ qouter, int32_t* __restrict__ qinner, double* __restrict__ qval )
{
  const unsigned long long d = blockIdx.x * uint64_t( blockDim.x ) + threadIdx.x;
  if( d >= nc ) { return; }
  const AsmContact kd = asm_load( a, d );
  if( kd.type > SG_BALL_PLANE ) { if( !EMIT ) { qcnt[d] = 0u; } return; }
  const bool two = kd.type == SG_BALL_BALL;
  // bodies of d in ascending order: the rows 2b, 2b+1 of the lower body come first in every accumulation
  const uint32_t b0 = two ? min( kd.i, kd.j ) : kd.i, b1 = two ? max( kd.i, kd.j ) : 0xffffffffu;
  uint32_t pa = inc_start[b0], ea = inc_start[b0 + 1u];
  uint32_t pb = two ? inc_start[b1] : 0u, eb = two ? inc_start[b1 + 1u] : 0u;
  const double minv0 = 1.0 / mass[b0], minv1 = two ? 1.0 / mass[b1] : 0.0;
  uint32_t cnt = 0u, o = EMIT ? qouter[d] : 0u;
  while( pa < ea || pb < eb )
  {
    const uint32_t ca = ( pa < ea ) ? inc[pa] : 0xffffffffu, cb = ( pb < eb ) ? inc[pb] : 0xffffffffu;
    const uint32_t c = ca < cb ? ca : cb;
    if( ca == c ) { ++pa; }
    if( cb == c ) { ++pb; }
    const AsmContact kc = ( c == uint32_t( d ) ) ? kd : asm_load( a, c );
    // terms over the shared rows in ascending row order
    double acc = 0.0;
    bool any = false;
    #pragma unroll
    for( int s = 0; s < 2; ++s )
    {
      const uint32_t b = s == 0 ? b0 : b1;
      if( s == 1 && !two ) { break; }
      if( !asm_has_body( kc, b ) ) { continue; }
      const double minv = s == 0 ? minv0 : minv1;
      #pragma unroll
      for( int ax = 0; ax < 2; ++ax )
      {
        const double nc_ = asm_coeff( kc, b, ax ), nd_ = asm_coeff( kd, b, ax );
        if( nc_ == 0.0 || nd_ == 0.0 ) { continue; } // a pruned coefficient: no such term
        const double t = ( nc_ * minv ) * nd_;
        acc = any ? acc + t : t;
        any = true;
      }
    }
    if( !any ) { continue; }
    if( EMIT ) { qinner[o] = int32_t( c ); qval[o] = acc; ++o; }
    ++cnt;
  }
  if( !EMIT ) { qcnt[d] = cnt; }
}

// ---- ConstraintCache as a sorted-key join --------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long asm_key( const uint32_t type, const uint32_t i, const uint32_t j )
{
  // ball-ball: ( i, j ); drum / plane: ( geometry, ball ) -- the keys of ball2d/ConstraintCache.cpp
  return type == SG_BALL_BALL ? ( ( unsigned long long )( i ) << 32 ) | j : ( ( unsigned long long )( j ) << 32 ) | i;
}
__global__ void __launch_bounds__( 256 ) k_cache_lookup( const unsigned long long nc, const ContactOut2D a, const uint32_t ncomp, const unsigned long long c_nbb, const unsigned long long c_ndrum, const unsigned long long c_n,
                                                        const uint32_t* __restrict__ c_type, const uint32_t* __restrict__ c_i, const uint32_t* __restrict__ c_j, const double* __restrict__ c_r, double* __restrict__ r_out,
                                                        unsigned long long* __restrict__ hits )
{
  const unsigned long long c = blockIdx.x * uint64_t( blockDim.x ) + threadIdx.x;
  if( c >= nc ) { return; }
  const uint32_t type = a.type[c];
  const unsigned long long key = asm_key( type, a.i[c], a.j[c] );
  // the cached list is in active-set order: [ ball-ball | drums | planes ], each segment ascending in its key
  unsigned long long lo = type == SG_BALL_BALL ? 0ull : ( type == SG_BALL_DRUM ? c_nbb : c_nbb + c_ndrum );
  unsigned long long hi = type == SG_BALL_BALL ? c_nbb : ( type == SG_BALL_DRUM ? c_nbb + c_ndrum : c_n );
  while( lo < hi )
  {
    const unsigned long long mid = ( lo + hi ) >> 1;
    if( asm_key( c_type[mid], c_i[mid], c_j[mid] ) < key ) { lo = mid + 1ull; } else { hi = mid; }
  }
  const unsigned long long end = type == SG_BALL_BALL ? c_nbb : ( type == SG_BALL_DRUM ? c_nbb + c_ndrum : c_n );
  const bool hit = lo < end && c_type[lo] == type && asm_key( c_type[lo], c_i[lo], c_j[lo] ) == key;
  for( uint32_t k = 0; k < ncomp; ++k ) { r_out[c * ncomp + k] = hit ? c_r[lo * ncomp + k] : 0.0; } // a miss: r.setZero() (ConstraintCache.cpp:122)
  if( hit ) { atomicAdd( hits, 1ull ); }
}

#endif
