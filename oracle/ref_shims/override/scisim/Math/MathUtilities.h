// oracle/ref_shims/override/scisim/Math/MathUtilities.h -- TEST INFRASTRUCTURE.
// Found ahead of the reference's own scisim/Math/MathUtilities.h (oracle/Makefile.ref puts this directory first on the
// include path) when ball2d/StaticGeometry/StaticPlane.cpp is compiled unchanged for oracle/_ref: the real header needs
// Eigen/LU and Eigen::DenseBase, which the Eigen stand-in does not provide, and the compiled files only use its stream
// (de)serialisers, which are restated here in the reference's own byte layout (so that Ball2DState::serialize, compiled unchanged, writes the
// reference's snapshot format).  The one piece of arithmetic here is the reference's inline 2-D cross product (scisim/Math/MathUtilities.h:15-19, one expression),
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

  // Dense types, in the reference's own layout (scisim/Math/MathUtilities.h:42-126): a dimension that is dynamic is written first as an Eigen::Index
  // (8 bytes), rows before columns, then the raw coefficients.  Fixed sizes: the coefficients only; N x 1: the row count; R x N: the column count;
  // N x M (for which the reference has no deserialiser): both.
  namespace detail
  {
    typedef std::ptrdiff_t Index; // Eigen::Index
    template<typename T, int R, int C, int Opt, int MR, int MC> void writeShape( const Eigen::Matrix<T, R, C, Opt, MR, MC>&, std::ostream& ) {}
    template<typename T, int R, int C, int Opt, int MR, int MC> void writeShape( const Eigen::Array<T, R, C, Opt, MR, MC>&, std::ostream& ) {}
    template<typename T, int Opt, int MR, int MC> void writeShape( const Eigen::Matrix<T, Eigen::Dynamic, 1, Opt, MR, MC>& v, std::ostream& stm ) { const Index n = v.rows(); stm.write( reinterpret_cast<const char*>( &n ), sizeof( n ) ); }
    template<typename T, int R, int Opt, int MR, int MC> void writeShape( const Eigen::Matrix<T, R, Eigen::Dynamic, Opt, MR, MC>& v, std::ostream& stm ) { const Index n = v.cols(); stm.write( reinterpret_cast<const char*>( &n ), sizeof( n ) ); }
    template<typename T, int Opt, int MR, int MC> void writeShape( const Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic, Opt, MR, MC>& v, std::ostream& stm ) { const Index rc[2] = { v.rows(), v.cols() }; stm.write( reinterpret_cast<const char*>( rc ), sizeof( rc ) ); }
    template<typename T, int R, int C, int Opt, int MR, int MC> void readShape( Eigen::Matrix<T, R, C, Opt, MR, MC>&, std::istream& ) {}
    template<typename T, int R, int C, int Opt, int MR, int MC> void readShape( Eigen::Array<T, R, C, Opt, MR, MC>&, std::istream& ) {}
    template<typename T, int Opt, int MR, int MC> void readShape( Eigen::Matrix<T, Eigen::Dynamic, 1, Opt, MR, MC>& v, std::istream& stm ) { Index n; stm.read( reinterpret_cast<char*>( &n ), sizeof( n ) ); v.resize( int( n ) ); }
    template<typename T, int R, int Opt, int MR, int MC> void readShape( Eigen::Matrix<T, R, Eigen::Dynamic, Opt, MR, MC>& v, std::istream& stm ) { Index n; stm.read( reinterpret_cast<char*>( &n ), sizeof( n ) ); v.resize( R, int( n ) ); }
    template<typename T, int Opt, int MR, int MC> void readShape( Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic, Opt, MR, MC>& v, std::istream& stm ) { Index rc[2]; stm.read( reinterpret_cast<char*>( rc ), sizeof( rc ) ); v.resize( int( rc[0] ), int( rc[1] ) ); }
  }
  // Sparse matrices, RESTATED from scisim/Math/MathUtilities.cpp:142-176 (that file needs Eigen internals the stand-in does not have): rows, cols,
  // non-zeros as Eigen::Index (8 bytes each), the inner indices, the cols + 1 outer indices (32-bit StorageIndex), the values
  inline void serialize( const SparseMatrixsc& A, std::ostream& stm )
  {
    const detail::Index hdr[3] = { A.rows(), A.cols(), A.nonZeros() };
    stm.write( reinterpret_cast<const char*>( hdr ), sizeof( hdr ) );
    stm.write( reinterpret_cast<const char*>( A.innerIndexPtr() ), A.nonZeros() * sizeof( int ) );
    stm.write( reinterpret_cast<const char*>( A.outerIndexPtr() ), ( A.cols() + 1 ) * sizeof( int ) );
    stm.write( reinterpret_cast<const char*>( A.valuePtr() ), A.nonZeros() * sizeof( scalar ) );
  }
  inline void deserialize( SparseMatrixsc& A, std::istream& stm )
  {
    detail::Index hdr[3];
    stm.read( reinterpret_cast<char*>( hdr ), sizeof( hdr ) );
    std::vector<int> outer( static_cast<size_t>( hdr[1] ) + 1 ), inner( static_cast<size_t>( hdr[2] ) );
    std::vector<scalar> values( static_cast<size_t>( hdr[2] ) );
    stm.read( reinterpret_cast<char*>( inner.data() ), inner.size() * sizeof( int ) );
    stm.read( reinterpret_cast<char*>( outer.data() ), outer.size() * sizeof( int ) );
    stm.read( reinterpret_cast<char*>( values.data() ), values.size() * sizeof( scalar ) );
    A.setCompressed( int( hdr[0] ), outer, inner, values );
  }
  template<typename T> T deserialize( std::istream& stm ) { T v; detail::readShape( v, stm ); stm.read( reinterpret_cast<char*>( v.data() ), v.size() * sizeof( *v.data() ) ); return v; }
  template<typename T> void serialize( const T& v, std::ostream& stm ) { detail::writeShape( v, stm ); stm.write( reinterpret_cast<const char*>( v.data() ), v.size() * sizeof( *v.data() ) ); }
}

#endif
