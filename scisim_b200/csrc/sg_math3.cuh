// sg_math3.cuh -- 3-vector / 3x3 helpers with the evaluation order the reference's Eigen expressions have
// (SURVEY.md H2): every 3-term reduction is (a0*b0 + a1*b1) + a2*b2; normalisation divides by sqrt(v.v) and
// is skipped when v.v == 0.  Built with -fmad=false, so no product-sum is contracted.
#ifndef SG_MATH3_CUH
#define SG_MATH3_CUH

struct V3d { double x, y, z; };
struct M3d { double m[9]; }; // row-major, like SCISim's Matrix33sr

__device__ __forceinline__ V3d v3( const double x, const double y, const double z ) { V3d r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3d operator-( const V3d a, const V3d b ) { return v3( a.x - b.x, a.y - b.y, a.z - b.z ); }
__device__ __forceinline__ V3d operator+( const V3d a, const V3d b ) { return v3( a.x + b.x, a.y + b.y, a.z + b.z ); }
__device__ __forceinline__ V3d operator*( const double s, const V3d a ) { return v3( s * a.x, s * a.y, s * a.z ); }
__device__ __forceinline__ V3d operator-( const V3d a ) { return v3( -a.x, -a.y, -a.z ); }
__device__ __forceinline__ double dot3( const V3d a, const V3d b ) { return ( a.x * b.x + a.y * b.y ) + a.z * b.z; }
__device__ __forceinline__ double at3( const V3d v, const int i ) { return i == 0 ? v.x : ( i == 1 ? v.y : v.z ); }
__device__ __forceinline__ V3d normalized3( const V3d a )
{
  const double z = dot3( a, a );
  if( z > 0.0 ) { const double s = sqrt( z ); return v3( a.x / s, a.y / s, a.z / s ); }
  return a;
}
__device__ __forceinline__ V3d col3( const M3d& R, const int j ) { return v3( R.m[j], R.m[3 + j], R.m[6 + j] ); }
// A x
__device__ __forceinline__ V3d mul3( const M3d& A, const V3d v )
{
  return v3( ( A.m[0] * v.x + A.m[1] * v.y ) + A.m[2] * v.z, ( A.m[3] * v.x + A.m[4] * v.y ) + A.m[5] * v.z, ( A.m[6] * v.x + A.m[7] * v.y ) + A.m[8] * v.z );
}
// A^T x
__device__ __forceinline__ V3d mulT3( const M3d& A, const V3d v )
{
  return v3( ( A.m[0] * v.x + A.m[3] * v.y ) + A.m[6] * v.z, ( A.m[1] * v.x + A.m[4] * v.y ) + A.m[7] * v.z, ( A.m[2] * v.x + A.m[5] * v.y ) + A.m[8] * v.z );
}
// A B
__device__ __forceinline__ M3d mul33( const M3d& A, const M3d& B )
{
  M3d C;
  #pragma unroll
  for( int r = 0; r < 3; ++r )
  {
    #pragma unroll
    for( int c = 0; c < 3; ++c ) { C.m[3 * r + c] = ( A.m[3 * r] * B.m[c] + A.m[3 * r + 1] * B.m[3 + c] ) + A.m[3 * r + 2] * B.m[6 + c]; }
  }
  return C;
}
// A^T B
__device__ __forceinline__ M3d mulTN33( const M3d& A, const M3d& B )
{
  M3d C;
  #pragma unroll
  for( int r = 0; r < 3; ++r )
  {
    #pragma unroll
    for( int c = 0; c < 3; ++c ) { C.m[3 * r + c] = ( A.m[r] * B.m[c] + A.m[3 + r] * B.m[3 + c] ) + A.m[6 + r] * B.m[6 + c]; }
  }
  return C;
}
// R diag(d) R^T evaluated as (R * diag(d)) * R^T
__device__ __forceinline__ M3d world_inertia3( const M3d& R, const V3d d )
{
  M3d A;
  #pragma unroll
  for( int r = 0; r < 3; ++r ) { A.m[3 * r] = R.m[3 * r] * d.x; A.m[3 * r + 1] = R.m[3 * r + 1] * d.y; A.m[3 * r + 2] = R.m[3 * r + 2] * d.z; }
  M3d B;
  #pragma unroll
  for( int r = 0; r < 3; ++r )
  {
    #pragma unroll
    for( int c = 0; c < 3; ++c ) { B.m[3 * r + c] = ( A.m[3 * r] * R.m[3 * c] + A.m[3 * r + 1] * R.m[3 * c + 1] ) + A.m[3 * r + 2] * R.m[3 * c + 2]; }
  }
  return B;
}

#endif
