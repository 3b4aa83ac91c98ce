// sg_ccd.h -- ball-ball continuous collision test, shared by the kernels (ball2d pass 1) and, compiled as plain C++, by the CPU
// suite (tests/ccd_harness.cpp runs this very function against the reference's compiled CollisionDetectionUtilities.cpp).
//
// Reference: scisim/CollisionDetection/CollisionDetectionUtilities.cpp:3-121 -- coefficients of |d0 + t d1|^2 = (ra + rb)^2 and the
// roots test of the quadratic.  FP64, no FMA contraction, evaluation order as written there.
//
// The reference's verdict depends on its two quotients only through their signs and through root0 > 1, and both can be read off the
// operands EXACTLY, without dividing (IEEE round-to-nearest):
//   root1 = 2 c0 / (-c1 - s)    (c1 > 0: denominator < 0)   is negative  <=>  c0 > 0  (a quotient this far from underflow keeps its sign)
//   root1 = (-c1 + s) / (2 c2)  (c1 <= 0)                    is never negative (numerator >= 0, c2 > 0)
//   root0 = (-c1 - s) / (2 c2)  (c1 >= 0)                    is never > 1 (numerator <= 0)
//   root0 = 2 c0 / (-c1 + s)    (c1 < 0: denominator b > 0)  fl(a / b) > 1  <=>  a > b:  a > b gives a - b >= ulp(b) > b 2^-53, so a / b lies
//                                                            above the midpoint of 1 and its successor; a <= b gives a / b <= 1
// and the square root is needed for the last case only.  Operands of a magnitude where a term could overflow or be NaN, or where a
// quotient could underflow to -0 (|c1| >= 1e100, |4 c2 c0| >= 1e200, 0 < c0 < 1e-200 -- never seen in a simulation), take the
// reference's expressions verbatim.
#ifndef SG_CCD_H
#define SG_CCD_H

#include <cmath>

#ifdef __CUDACC__
#define SG_CCD_HD __host__ __device__ __forceinline__
#define SG_CCD_COLD static __host__ __device__ __noinline__
#else
#define SG_CCD_HD inline
#define SG_CCD_COLD inline
#endif

// the reference's roots test, expression for expression (c2 != 0)
SG_CCD_COLD bool sg_ccd_roots_verbatim( const double c0, const double c1, const double c2 )
{
  const double c1c1 = c1 * c1;
  const double fc2c0 = 4.0 * c2 * c0;
  if( c1c1 < fc2c0 ) { return false; }
  const double s = sqrt( c1c1 - fc2c0 );
  const double root1 = ( c1 > 0.0 ) ? ( 2.0 * c0 ) / ( -c1 - s ) : ( -c1 + s ) / ( 2.0 * c2 );
  if( root1 < 0.0 ) { return false; }
  const double root0 = ( c1 >= 0.0 ) ? ( -c1 - s ) / ( 2.0 * c2 ) : ( 2.0 * c0 ) / ( -c1 + s );
  if( root0 > 1.0 ) { return false; }
  return true;
}

// same verdict, no division (c2 != 0)
SG_CCD_HD bool sg_ccd_roots( const double c0, const double c1, const double c2 )
{
  const double c1c1 = c1 * c1;
  const double fc2c0 = 4.0 * c2 * c0;
  if( !( fabs( c1 ) < 1.0e100 && fabs( fc2c0 ) < 1.0e200 ) || ( c0 > 0.0 && c0 < 1.0e-200 ) ) { return sg_ccd_roots_verbatim( c0, c1, c2 ); }
  if( c1c1 < fc2c0 ) { return false; }
  if( c1 > 0.0 ) { return !( c0 > 0.0 ); } // receding: root1 < 0 <=> c0 > 0; root0 <= 0
  if( c1 == 0.0 ) { return true; }         // root1 = s / (2 c2) >= 0, root0 = -s / (2 c2) <= 0
  const double s = sqrt( c1c1 - fc2c0 );
  return !( 2.0 * c0 > ( -c1 + s ) );      // approaching: root1 >= 0; root0 > 1 <=> 2 c0 > -c1 + s
}

// ball a = (q0a -> q1a, ra), ball b likewise; a is the lower body index (the reference's argument order)
SG_CCD_HD bool sg_ccd_ball_ball( const double q0ax, const double q0ay, const double q1ax, const double q1ay, const double ra,
                                 const double q0bx, const double q0by, const double q1bx, const double q1by, const double rb )
{
  const double d0x = q0ax - q0bx;
  const double d0y = q0ay - q0by;
  const double d1x = ( q1ax - q1bx ) - d0x;
  const double d1y = ( q1ay - q1by ) - d0y;
  const double rs = ra + rb;
  const double c0 = ( d0x * d0x + d0y * d0y ) - rs * rs;
  const double c1 = 2.0 * ( d0x * d1x + d0y * d1y );
  const double c2 = d1x * d1x + d1y * d1y;
  if( c2 != 0.0 ) { return sg_ccd_roots( c0, c1, c2 ); }
  return c0 <= 0.0;
}

#endif
