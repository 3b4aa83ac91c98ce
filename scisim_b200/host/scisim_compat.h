// scisim_compat.h -- the slice of SCISim's core interfaces the GPU shim is written against.
//
// Inside SCISim define SCISIM_B200_WITH_SCISIM and the real headers are used (scisim/Math/MathDefines.h,
// scisim/UnconstrainedMaps/{UnconstrainedMap,FlowableSystem}.h, scisim/Constraints/ConstrainedSystem.h).  Eigen is not
// available in the build container of this repository, so for the stand-alone build of the shim the same names are
// declared here with exactly the members the shim touches: VectorXs is contiguous doubles with data()/size()/resize(),
// and the abstract interfaces repeat the reference's signatures (file:line given at each).
#ifndef SCISIM_B200_COMPAT_H
#define SCISIM_B200_COMPAT_H

#ifdef SCISIM_B200_WITH_SCISIM
#include "scisim/Constraints/ConstrainedSystem.h"
#include "scisim/Math/MathDefines.h"
#include "scisim/UnconstrainedMaps/FlowableSystem.h"
#include "scisim/UnconstrainedMaps/UnconstrainedMap.h"
#else

#include <memory>
#include <ostream>
#include <string>
#include <vector>

using scalar = double; // scisim/Math/MathDefines.h:16

// stand-in for Eigen::Matrix<double,-1,1> (scisim/Math/MathDefines.h): contiguous storage only
class VectorXs
{
public:
  VectorXs() = default;
  explicit VectorXs( const long n ) : m_v( static_cast<std::size_t>( n ), 0.0 ) {}
  long size() const { return static_cast<long>( m_v.size() ); }
  void resize( const long n ) { m_v.resize( static_cast<std::size_t>( n ) ); }
  double* data() { return m_v.data(); }
  const double* data() const { return m_v.data(); }
  double& operator()( const long i ) { return m_v[static_cast<std::size_t>( i )]; }
  const double& operator()( const long i ) const { return m_v[static_cast<std::size_t>( i )]; }
  void setZero() { for( double& x : m_v ) { x = 0.0; } }
private:
  std::vector<double> m_v;
};

// scisim/UnconstrainedMaps/FlowableSystem.h:6-58 (only what the GPU maps need: sizes, kinematic flags and computeForce -- masses
// and gravity are pushed to the device once through the back end; computeForce is called once per state by
// GravityOnlyGuard to make sure the system's force IS that gravity)
class FlowableSystem
{
public:
  virtual ~FlowableSystem() = default;
  virtual int nqdofs() const = 0;
  virtual int nvdofs() const = 0;
  virtual unsigned numVelDoFsPerBody() const = 0;
  virtual unsigned ambientSpaceDimensions() const = 0;
  virtual bool isKinematicallyScripted( const int i ) const = 0;
  virtual void computeForce( const VectorXs& q, const VectorXs& v, const scalar& t, VectorXs& F ) = 0; // FlowableSystem.h:24
  virtual std::string name() const = 0;
};

// scisim/UnconstrainedMaps/UnconstrainedMap.h:13-43
class UnconstrainedMap
{
public:
  UnconstrainedMap( const UnconstrainedMap& ) = delete;
  UnconstrainedMap& operator=( const UnconstrainedMap& ) = delete;
  virtual ~UnconstrainedMap() = default;
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) = 0;
  virtual std::string name() const = 0;
  virtual void serialize( std::ostream& output_stream ) const = 0;
protected:
  UnconstrainedMap() = default;
};

// scisim/Constraints/Constraint.h:18-111 (identity only; the solver-side methods stay in SCISim)
class Constraint
{
public:
  virtual ~Constraint() = default;
  virtual std::string name() const = 0;
};

// scisim/Constraints/ConstrainedSystem.h:15-38
class ConstrainedSystem
{
public:
  virtual ~ConstrainedSystem() = default;
  virtual void computeActiveSet( const VectorXs& q0, const VectorXs& qp, const VectorXs& v, std::vector<std::unique_ptr<Constraint>>& active_set ) = 0;
  virtual void clearConstraintCache() = 0;
  virtual void cacheConstraint( const Constraint& constraint, const VectorXs& r ) = 0;
  virtual void getCachedConstraintImpulse( const Constraint& constraint, VectorXs& r ) const = 0;
  virtual bool constraintCacheEmpty() const = 0;
};

#endif
#endif
