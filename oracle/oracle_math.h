// oracle/oracle_math.h
//
// TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the shipped product path; it may be
// imported / linked / executed only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs, and there only as the checker or the timed CPU baseline.
//
// Fixed-size vector helpers that restate the arithmetic Eigen 3.3.4 (the reference's un-vendored
// dependency, scripts_include/get_eigen.sh:3-7) performs for the expressions on the hot path.
// Conventions:
//   * all FP64, round-to-nearest, NO FMA contraction (build with -ffp-contract=off),
//   * 2-term dot:  a0*b0 + a1*b1
//   * 3-term dot:  (a0*b0 + a1*b1) + a2*b2     (Release / SSE2-vectorised redux order)
//   * normalized(v) = v / sqrt(v.v) component-wise division, only when v.v > 0
//   * diagonal sparse * dense: res_k = 0.0 + (s*d_k)*rhs_k with s the left-assoc. scalar prefix
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H

#include <cmath>
#include <cstdint>

namespace orc
{

struct V2 { double x, y; };
struct V3 { double x, y, z; };
// Row-major 3x3
struct M3 { double m[9]; };

inline V2 operator-( const V2& a, const V2& b ) { return V2{ a.x - b.x, a.y - b.y }; }
inline V2 operator+( const V2& a, const V2& b ) { return V2{ a.x + b.x, a.y + b.y }; }
inline V2 operator*( double s, const V2& a ) { return V2{ s * a.x, s * a.y }; }
inline double dot( const V2& a, const V2& b ) { return a.x * b.x + a.y * b.y; }
inline double squaredNorm( const V2& a ) { return dot( a, a ); }
inline double norm( const V2& a ) { return std::sqrt( squaredNorm( a ) ); }
inline V2 normalized( const V2& a )
{
  const double z = squaredNorm( a );
  if( z > 0.0 ) { const double s = std::sqrt( z ); return V2{ a.x / s, a.y / s }; }
  return a;
}

inline V3 operator-( const V3& a, const V3& b ) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator+( const V3& a, const V3& b ) { return V3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator*( double s, const V3& a ) { return V3{ s * a.x, s * a.y, s * a.z }; }
inline V3 operator-( const V3& a ) { return V3{ -a.x, -a.y, -a.z }; }
inline double dot( const V3& a, const V3& b ) { return ( a.x * b.x + a.y * b.y ) + a.z * b.z; }
inline double squaredNorm( const V3& a ) { return dot( a, a ); }
inline double norm( const V3& a ) { return std::sqrt( squaredNorm( a ) ); }
inline V3 normalized( const V3& a )
{
  const double z = squaredNorm( a );
  if( z > 0.0 ) { const double s = std::sqrt( z ); return V3{ a.x / s, a.y / s, a.z / s }; }
  return a;
}
inline V3 cross( const V3& a, const V3& b )
{
  return V3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}

// y = A x, each row reduced as (a0*x0 + a1*x1) + a2*x2
inline V3 mul( const M3& A, const V3& v )
{
  return V3{ ( A.m[0] * v.x + A.m[1] * v.y ) + A.m[2] * v.z,
             ( A.m[3] * v.x + A.m[4] * v.y ) + A.m[5] * v.z,
             ( A.m[6] * v.x + A.m[7] * v.y ) + A.m[8] * v.z };
}
// y = A^T x
inline V3 mulT( const M3& A, const V3& v )
{
  return V3{ ( A.m[0] * v.x + A.m[3] * v.y ) + A.m[6] * v.z,
             ( A.m[1] * v.x + A.m[4] * v.y ) + A.m[7] * v.z,
             ( A.m[2] * v.x + A.m[5] * v.y ) + A.m[8] * v.z };
}
// C = A B
inline M3 mul( const M3& A, const M3& B )
{
  M3 C;
  for( int r = 0; r < 3; ++r )
    for( int c = 0; c < 3; ++c )
      C.m[3 * r + c] = ( A.m[3 * r + 0] * B.m[0 + c] + A.m[3 * r + 1] * B.m[3 + c] ) + A.m[3 * r + 2] * B.m[6 + c];
  return C;
}
// C = A^T B
inline M3 mulTN( const M3& A, const M3& B )
{
  M3 C;
  for( int r = 0; r < 3; ++r )
    for( int c = 0; c < 3; ++c )
      C.m[3 * r + c] = ( A.m[0 + r] * B.m[0 + c] + A.m[3 + r] * B.m[3 + c] ) + A.m[6 + r] * B.m[6 + c];
  return C;
}

}

#endif
