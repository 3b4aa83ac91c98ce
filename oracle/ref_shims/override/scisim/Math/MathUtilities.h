// oracle/ref_shims/override/scisim/Math/MathUtilities.h -- TEST INFRASTRUCTURE.
// Found ahead of the reference's own scisim/Math/MathUtilities.h (oracle/Makefile.ref puts this directory first on the
// include path) when ball2d/StaticGeometry/StaticPlane.cpp is compiled unchanged for oracle/_ref: the real header needs
// Eigen/LU and Eigen::DenseBase, which the Eigen stand-in does not provide, and StaticPlane.cpp only uses its stream
// (de)serialisers.  No arithmetic lives here.
#ifndef SCISIM_B200_MATH_UTILITIES_OVERRIDE
#define SCISIM_B200_MATH_UTILITIES_OVERRIDE

#include "scisim/Math/MathDefines.h"
#include "scisim/Utilities.h" // as the real header does (scisim/Math/MathUtilities.h:10); plain C++, no Eigen

#include <istream>
#include <ostream>

namespace MathUtilities
{
  template<typename T> T deserialize( std::istream& stm ) { T v; stm.read( reinterpret_cast<char*>( v.data() ), v.size() * sizeof( scalar ) ); return v; }
  template<typename T> void serialize( const T& v, std::ostream& stm ) { stm.write( reinterpret_cast<const char*>( v.data() ), v.size() * sizeof( scalar ) ); }
}

#endif
