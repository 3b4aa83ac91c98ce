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
  // sparse matrices (rigidbody3d/RigidBody3DState.cpp serialises its four mass matrices): rows, cols, nnz, then outer / inner indices and values of the
  // stand-in's compressed storage -- again a private format of the shims, written and read only here
  inline void serialize( const SparseMatrixsc& A, std::ostream& stm )
  {
    const long long hdr[3] = { A.rows(), A.cols(), A.nonZeros() };
    stm.write( reinterpret_cast<const char*>( hdr ), sizeof( hdr ) );
    stm.write( reinterpret_cast<const char*>( A.outerIndexPtr() ), ( A.cols() + 1 ) * sizeof( int ) );
    stm.write( reinterpret_cast<const char*>( A.innerIndexPtr() ), A.nonZeros() * sizeof( int ) );
    stm.write( reinterpret_cast<const char*>( A.valuePtr() ), A.nonZeros() * sizeof( scalar ) );
  }
  inline void deserialize( SparseMatrixsc& A, std::istream& stm )
  {
    long long hdr[3];
    stm.read( reinterpret_cast<char*>( hdr ), sizeof( hdr ) );
    std::vector<int> outer( static_cast<size_t>( hdr[1] ) + 1 ), inner( static_cast<size_t>( hdr[2] ) );
    std::vector<scalar> values( static_cast<size_t>( hdr[2] ) );
    stm.read( reinterpret_cast<char*>( outer.data() ), outer.size() * sizeof( int ) );
    stm.read( reinterpret_cast<char*>( inner.data() ), inner.size() * sizeof( int ) );
    stm.read( reinterpret_cast<char*>( values.data() ), values.size() * sizeof( scalar ) );
    A.setCompressed( int( hdr[0] ), outer, inner, values );
  }
  template<typename T> T deserialize( std::istream& stm ) { T v; readShape( v, stm ); stm.read( reinterpret_cast<char*>( v.data() ), v.size() * sizeof( *v.data() ) ); return v; }
  template<typename T> void serialize( const T& v, std::ostream& stm ) { writeShape( v, stm ); stm.write( reinterpret_cast<const char*>( v.data() ), v.size() * sizeof( *v.data() ) ); }
}

#endif
