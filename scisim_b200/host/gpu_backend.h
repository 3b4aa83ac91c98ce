// gpu_backend.h -- C++14 host shim over the C ABI (include/scisim_b200.h): the classes a SCISim maintainer links in so
// that Ball2DSim / RigidBody3DSim run their unconstrained flow and active-set computation on the GPU.
//
//   GpuBall2DBackend        owns the sg_ctx; static scene data in, contact records out (reference order)
//   GpuSymplecticEulerMap,
//   GpuVerletMap            UnconstrainedMap implementations registered beside ball2d's own maps
//                           (ball2d/Ball2DUtilities.cpp:37, ball2dutils/Ball2DSceneParser.cpp:624-631)
//   PairImpulseCache        flat sorted replacement for ball2d/ConstraintCache.{h,cpp} (std::map keyed by pair)
//
// Error convention of the reference: print to std::cerr and std::exit( EXIT_FAILURE ).
#ifndef SCISIM_B200_GPU_BACKEND_H
#define SCISIM_B200_GPU_BACKEND_H

#include "scisim_compat.h"

#include "../../include/scisim_b200.h"

#include <cstdint>
#include <utility>
#include <vector>

// One entry per Constraint the reference would emplace_back, in the reference's order
struct GpuContact2D
{
  uint32_t type; // SG_BALL_BALL, SG_BALL_DRUM, SG_BALL_PLANE
  uint32_t i;    // ball (first ball for ball-ball)
  uint32_t j;    // second ball / drum / plane
  double n[2];
  double p[2];
  double depth;
};

class GpuBall2DBackend final
{
public:
  explicit GpuBall2DBackend( const int device = 0 );
  ~GpuBall2DBackend();
  GpuBall2DBackend( const GpuBall2DBackend& ) = delete;
  GpuBall2DBackend& operator=( const GpuBall2DBackend& ) = delete;

  // Ball2DState contents (ball2d/Ball2DState.h): radii, per-ball masses, gravity, planes (x, n), drums (X, R)
  void setBodies( const VectorXs& r, const VectorXs& m );
  void setGravity( const double gx, const double gy );
  void setPlanes( const std::vector<double>& x, const std::vector<double>& n );
  void setDrums( const std::vector<double>& x, const std::vector<double>& r );

  // UnconstrainedMap::flow for the two ball2d maps
  void flow( const int map_kind, const VectorXs& q0, const VectorXs& v0, const scalar& dt, VectorXs& q1, VectorXs& v1 );
  // Ball2DSim::computeActiveSet (ball2d/Ball2DSim.cpp:151-173): contacts in active_set order
  // from_last_flow: (q0, q1) are the vectors the last flow() call read and wrote, unchanged since -- what
  // ImpactMap::flow hands to computeActiveSet (ImpactMap.cpp:54-58); they are then not uploaded a second time.
  void computeActiveSet( const VectorXs& q0, const VectorXs& q1, std::vector<GpuContact2D>& contacts, uint64_t* num_candidates = nullptr, const bool from_last_flow = false );
  // SpatialGridDetector::getPotentialOverlaps (ball2d/SpatialGridDetector.h:39) on caller-built boxes [minx,miny,maxx,maxy]
  void getPotentialOverlaps( const std::vector<double>& aabbs, std::vector<std::pair<unsigned,unsigned>>& overlaps );

  sg_ctx* context() { return m_ctx; }

private:
  void check( const int rc, const char* what ) const;
  sg_ctx* m_ctx;
  unsigned m_nbodies;
};

class GpuSymplecticEulerMap final : public UnconstrainedMap
{
public:
  explicit GpuSymplecticEulerMap( GpuBall2DBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "symplectic_euler"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuBall2DBackend& m_backend;
};

class GpuVerletMap final : public UnconstrainedMap
{
public:
  explicit GpuVerletMap( GpuBall2DBackend& backend ) : m_backend( backend ) {}
  virtual void flow( const VectorXs& q0, const VectorXs& v0, FlowableSystem& fsys, const unsigned iteration, const scalar& dt, VectorXs& q1, VectorXs& v1 ) override;
  virtual std::string name() const override { return "verlet"; }
  virtual void serialize( std::ostream& ) const override {}
private:
  GpuBall2DBackend& m_backend;
};

// ball2d/ConstraintCache.{h,cpp}: three std::map<std::pair<unsigned,unsigned>,VectorXs> (ball-ball, plane-ball,
// drum-ball).  The active set arrives sorted, so the cache is three flat arrays kept in key order: insertion appends
// (re-sorting only if a caller inserts out of order), lookup is a binary search, a miss zeroes r (ConstraintCache.cpp:122).
class PairImpulseCache final
{
public:
  void clear();
  bool empty() const;
  // kind: 0 ball-ball (i,j), 1 plane-ball (plane, ball), 2 drum-ball (drum, ball) -- the three maps of the reference
  void cacheConstraint( const int kind, const unsigned a, const unsigned b, const VectorXs& r );
  void getCachedConstraint( const int kind, const unsigned a, const unsigned b, VectorXs& r ) const;
private:
  struct Table
  {
    std::vector<uint64_t> keys;
    std::vector<double> values;
    unsigned width = 0;
    bool sorted = true;
    void sortIfNeeded();
  };
  mutable Table m_tables[3];
};

#endif
