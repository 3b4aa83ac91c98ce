// oracle/ccd.h
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). Restatement of the reference's ball-ball
// continuous-time collision test:
//   scisim/CollisionDetection/CollisionDetectionUtilities.cpp:3-22    computeCCDQuadraticCoeffs
//   scisim/CollisionDetection/CollisionDetectionUtilities.cpp:24-58   first/secondRootOfQuadratic
//   scisim/CollisionDetection/CollisionDetectionUtilities.cpp:60-121  ballBallCCDCollisionHappens
// Parity pin: the 11 known-answer cases of scisimtests/narrowphase_tests.cpp (tests/golden/ccd_cases.json).
#ifndef ORACLE_CCD_H
#define ORACLE_CCD_H

#include "oracle_math.h"

#include <algorithm>
#include <utility>

namespace orc
{

struct CCDCoeffs { double c0, c1, c2; };

inline CCDCoeffs computeCCDQuadraticCoeffs( const V2& q0a, const V2& q1a, const double ra, const V2& q0b, const V2& q1b, const double rb )
{
  const V2 q0delta = q0a - q0b;
  // Eigen evaluates q1a - q1b - q0delta coefficient-wise, left to right
  const V2 q1q0delta{ ( q1a.x - q1b.x ) - q0delta.x, ( q1a.y - q1b.y ) - q0delta.y };
  CCDCoeffs c;
  // std::pow( x, 2 ) is expanded to x*x by GCC at -O1 and above
  c.c0 = dot( q0delta, q0delta ) - ( ra + rb ) * ( ra + rb );
  c.c1 = 2.0 * dot( q0delta, q1q0delta );
  c.c2 = dot( q1q0delta, q1q0delta );
  return c;
}

inline std::pair<bool,double> ballBallCCDCollisionHappens( const CCDCoeffs& c )
{
  if( c.c2 != 0.0 )
  {
    const double c1c1 = c.c1 * c.c1;
    const double fc2c0 = 4 * c.c2 * c.c0;
    if( c1c1 < fc2c0 ) { return std::make_pair( false, 0.0 ); }
    const double dscr_sqrt = std::sqrt( c1c1 - fc2c0 );
    // secondRootOfQuadratic( a = c2, b = c1, c = c0 )
    const double root1 = ( c.c1 > 0.0 ) ? ( 2.0 * c.c0 ) / ( -c.c1 - dscr_sqrt ) : ( -c.c1 + dscr_sqrt ) / ( 2.0 * c.c2 );
    if( root1 < 0.0 ) { return std::make_pair( false, 0.0 ); }
    // firstRootOfQuadratic
    const double root0 = ( c.c1 >= 0.0 ) ? ( -c.c1 - dscr_sqrt ) / ( 2.0 * c.c2 ) : ( 2.0 * c.c0 ) / ( -c.c1 + dscr_sqrt );
    if( root0 > 1.0 ) { return std::make_pair( false, 0.0 ); }
    return std::make_pair( true, std::max( 0.0, root0 ) );
  }
  else
  {
    if( c.c0 <= 0.0 ) { return std::make_pair( true, 0.0 ); }
    return std::make_pair( false, 0.0 );
  }
}

inline std::pair<bool,double> ballBallCCDCollisionHappens( const V2& q0a, const V2& q1a, const double ra, const V2& q0b, const V2& q1b, const double rb )
{
  return ballBallCCDCollisionHappens( computeCCDQuadraticCoeffs( q0a, q1a, ra, q0b, q1b, rb ) );
}

}

#endif
