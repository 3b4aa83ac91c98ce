// oracle/ref_shims/override/scisim/Math/MathUtilities.h -- TEST INFRASTRUCTURE.
// Found ahead of the reference's own scisim/Math/MathUtilities.h (oracle/Makefile.ref puts this directory first on the
// include path) when ball2d/StaticGeometry/StaticPlane.cpp is compiled unchanged for oracle/_ref: the real header needs
// Eigen/LU and Eigen::DenseBase, which the Eigen stand-in does not provide, and StaticPlane.cpp only uses its stream
// (de)serialisers.  The one piece of arithmetic here is the reference's inline 2-D cross product (scisim/Math/MathUtilities.h:15-19, one expression),
// which the rigidbody2d constraint classes call; it is restated, and marked, below.
#ifndef SCISIM_B200_MATH_UTILITIES_OVERRIDE
#define SCISIM_B200_MATH_UTILITIES_OVERRIDE

#include "scisim/Math/MathDefines.h"
#include "scisim/Utilities.h" // as the real header does (scisim/Math/MathUtilities.h:10); plain C++, no Eigen

#include <istream>
#include <ostream>

namespace MathUtilities
{
  // RESTATED from scisim/Math/MathUtilities.h:16-19 (inline in the header this file shadows)
  inline scalar cross( const Vector2s& a, const Vector2s& b ) { return a.x() * b.y() - a.y() * b.x(); }

  // fixed sizes: the raw coefficients; the dynamic shapes of the stand-in ( R x N, N x 1 ): rows, cols, then the coefficients.  Both ends
  // of every stream that goes through here are written by oracle/ref_shims (the reference's own file format is not involved).
  template<typename T> void readShape( T&, std::istream&, decltype( &T::s )* = nullptr ) {}
  template<typename T> void readShape( T& v, std::istream& stm, decltype( &T::v )* = nullptr )
  {
    long long rc[2];
    stm.read( reinterpret_cast<char*>( rc ), sizeof( rc ) );
    v.resize( int( rc[0] ), int( rc[1] ) );
  }
  template<typename T> void writeShape( const T&, std::ostream&, decltype( &T::s )* = nullptr ) {}
  template<typename T> void writeShape( const T& v, std::ostream& stm, decltype( &T::v )* = nullptr )
  {
    const long long rc[2] = { v.rows(), v.cols() };
    stm.write( reinterpret_cast<const char*>( rc ), sizeof( rc ) );
  }
  template<typename T> T deserialize( std::istream& stm ) { T v; readShape( v, stm ); stm.read( reinterpret_cast<char*>( v.data() ), v.size() * sizeof( *v.data() ) ); return v; }
  template<typename T> void serialize( const T& v, std::ostream& stm ) { writeShape( v, stm ); stm.write( reinterpret_cast<const char*>( v.data() ), v.size() * sizeof( *v.data() ) ); }
}

#endif
