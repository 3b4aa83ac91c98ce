// oracle/rb3d.h
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). CPU restatement of the rigidbody3d hot path:
//   rigidbody3d/UnconstrainedMaps/SplitHamMap.cpp:17-182     flowSplitHam
//   rigidbody3d/UnconstrainedMaps/DMVMap.cpp:15-60,65-100,103-207   flowDMV (solveDMV, DMV)
//   rigidbody3d/Forces/NearEarthGravityForce.cpp:39-53, RigidBody3DSim.cpp:169-181   gravity (F.setZero(); F_lin += m g)
//   rigidbody3d/RigidBody3DState.cpp:70-240                  M0/Minv0 diagonals, world-space inertia R I0 R^T
//   rigidbody3d/RigidBody3DSim.cpp:250-262                   computeActiveSet (no portals: those are in rb3d_portals.h)
//   rigidbody3d/RigidBody3DSim.cpp:1057-1141                 generateAABBs(q1) + getPotentialOverlaps + dispatch
//   rigidbody3d/RigidBody3DSim.cpp:879-962                   dispatchNarrowPhaseCollision (kinematic rules, type switch)
//   rigidbody3d/RigidBody3DSim.cpp:793-819                   sphereSphereNarrowPhaseCollision
//   rigidbody3d/RigidBody3DSim.cpp:665-694                   boxBoxNarrowPhaseCollision   (-> boxbox.h)
//   rigidbody3d/RigidBody3DSim.cpp:844-876                   meshMeshNarrowPhaseCollision (-> MeshMeshUtilities.cpp:10-65)
//   rigidbody3d/RigidBody3DSim.cpp:1414-1502                 computeBodyPlaneActiveSetAllPairs
//   rigidbody3d/Geometry/RigidBodySphere.cpp:45-49, RigidBodyBox.cpp:45-51, RigidBodyTriangleMesh.cpp:187-200   AABBs
//   rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp:269-335   SDF lookup v(i,j,k), detectCollision
//   rigidbody3d/Constraints/{SphereSphere,StaticPlaneSphere,StaticPlaneBox,BodyBody,StaticPlaneBody}Constraint.cpp
//
// Eigen 3.3.4 semantics assumed where the reference delegates arithmetic to it (see oracle_math.h): 3-term
// reductions (a0+a1)+a2; Quaternion(Matrix3) by Shoemake's method; Quaternion::normalize with the squared
// norm reduced as (x^2+z^2)+(y^2+w^2) (packet reduction, see solveDMV); AngleAxis::toRotationMatrix and Quaternion::toRotationMatrix as in
// Eigen/src/Geometry.  Parity: no reference test stores expected values for any of this; the file is pinned against
// the reference's own sources compiled unchanged into oracle/_ref (oracle/Makefile.ref, tests/test_oracle_vs_reference.py): SplitHamMap / DMVMap /
// ExponentialEulerMap with the gravity force, the spatial grid, BoxBoxUtilities, the geometry classes' AABBs, the triangle-mesh class with its SDF narrow
// phase, the plane, cylinder and portal classes, RigidBody3DState (the mass matrices of setState and of updateMandMinv), and the constraint classes
// SphereSphere / KinematicSphereSphere / StaticPlaneSphere / StaticPlaneBox / StaticCylinderSphere / StaticCylinderBody (isActive, normal, contact point,
// depth for every such contact of its active sets), bit for bit -- except ExponentialEulerMap's projected orientation, where the reference calls
// Eigen::JacobiSVD and the stand-in supplies its own SVD: agreement to rounding (4e-15), not bit for bit.  The glue of RigidBody3DSim.cpp is pinned by
// running the reference's own RigidBody3DSim (compiled unchanged with 12 more of its sources, oracle/ref_shims/ref_rb3d_sim.cpp): computeActiveSet as a
// whole (every narrow phase, kinematic rules, planes, cylinders, portals) and flow over steps equal this file and rb3d_portals.h
// (tests/test_reference_sim_cpu.py).  sin/cos come from libm here (as in the reference) and from CUDA on the device, so rotating bodies are compared to 1e-12,
// everything else bit for bit.
#ifndef ORACLE_RB3D_H
#define ORACLE_RB3D_H

#include "boxbox.h"
#include "broadphase.h"
#include "oracle_math.h"

#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

namespace orc
{

enum RB3DGeoType : uint32_t { GEO_BOX = 0, GEO_SPHERE = 1, GEO_MESH = 3 };

enum RB3DContactType : uint32_t
{
  SPHERE_SPHERE = 10, KINEMATIC_SPHERE_SPHERE = 11, BODY_BODY = 12, KINEMATIC_BODY_BODY = 13,
  PLANE_SPHERE = 14, PLANE_BOX = 15, PLANE_BODY = 16,
  CYLINDER_SPHERE = 17, CYLINDER_BODY = 18
};

struct RB3DMesh
{
  std::vector<V3> verts;      // all mesh vertices (AABB)
  std::vector<V3> samples;    // surface samples (mesh-mesh)
  std::vector<V3> hull;       // convex hull vertices (mesh-plane)
  V3 cell_delta;
  uint32_t dims[3];
  V3 origin;
  V3 grid_end;                // origin + (dims-1)*delta as the reference stores it
  std::vector<double> sdf;    // x fastest: sdf[(k*ny + j)*nx + i]
  double v( const unsigned i, const unsigned j, const unsigned k ) const { return sdf[( std::size_t( k ) * dims[1] + j ) * dims[0] + i]; }

  // RigidBodyTriangleMesh::detectCollision
  bool detectCollision( const V3& x, V3& n ) const
  {
    if( x.x < origin.x || x.y < origin.y || x.z < origin.z ) { return false; }
    if( x.x > grid_end.x || x.y > grid_end.y || x.z > grid_end.z ) { return false; }
    const unsigned ix = unsigned( std::floor( ( x.x - origin.x ) / cell_delta.x ) );
    const unsigned iy = unsigned( std::floor( ( x.y - origin.y ) / cell_delta.y ) );
    const unsigned iz = unsigned( std::floor( ( x.z - origin.z ) / cell_delta.z ) );
    // The reference only asserts indices + 1 < dims; a sample exactly on grid_end would read out of bounds there.
    if( ix + 1 >= dims[0] || iy + 1 >= dims[1] || iz + 1 >= dims[2] ) { return false; }
    const V3 bc{ ( x.x - ( origin.x + double( ix ) * cell_delta.x ) ) / cell_delta.x,
                 ( x.y - ( origin.y + double( iy ) * cell_delta.y ) ) / cell_delta.y,
                 ( x.z - ( origin.z + double( iz ) * cell_delta.z ) ) / cell_delta.z };
    const V3 bci{ 1.0 - bc.x, 1.0 - bc.y, 1.0 - bc.z };
    const double v000 = v( ix, iy, iz ),     v100 = v( ix + 1, iy, iz );
    const double v010 = v( ix, iy + 1, iz ), v110 = v( ix + 1, iy + 1, iz );
    const double v001 = v( ix, iy, iz + 1 ),     v101 = v( ix + 1, iy, iz + 1 );
    const double v011 = v( ix, iy + 1, iz + 1 ), v111 = v( ix + 1, iy + 1, iz + 1 );
    const double dist = bci.z * ( bci.y * ( bci.x * v000 + bc.x * v100 ) + bc.y * ( bci.x * v010 + bc.x * v110 ) ) +
                         bc.z * ( bci.y * ( bci.x * v001 + bc.x * v101 ) + bc.y * ( bci.x * v011 + bc.x * v111 ) );
    if( dist > 0.0 ) { return false; }
    n.x = bci.z * ( bci.y * ( v100 - v000 ) + bc.y * ( v110 - v010 ) ) + bc.z * ( bci.y * ( v101 - v001 ) + bc.y * ( v111 - v011 ) );
    n.y = bci.z * ( bci.x * ( v010 - v000 ) + bc.x * ( v110 - v100 ) ) + bc.z * ( bci.x * ( v011 - v001 ) + bc.x * ( v111 - v101 ) );
    n.z = bci.y * ( bci.x * ( v001 - v000 ) + bc.x * ( v101 - v100 ) ) + bc.y * ( bci.x * ( v011 - v010 ) + bc.x * ( v111 - v110 ) );
    n.x /= cell_delta.x; n.y /= cell_delta.y; n.z /= cell_delta.z;
    n = normalized( n );
    return true;
  }
};

struct RB3DGeometry
{
  uint32_t type;
  double r;       // sphere
  V3 half;        // box half widths
  uint32_t mesh;  // index into RB3DScene::meshes
};

struct RB3DContact
{
  uint32_t type;
  uint32_t i;      // first (non-kinematic) body
  uint32_t j;      // second body / plane index
  uint32_t aux;    // plane-box: corner number; plane-body: hull vertex; else 0
  V3 n;
  V3 p;            // sphere-sphere: contact point at q0; body-body: p; plane-box/body: world point at q0; plane-sphere: x0 - r n
  double depth;    // penetrationDepth( q1 ) where the reference overrides it, else NaN
};

struct RB3DScene
{
  std::vector<RB3DGeometry> geometry;
  std::vector<RB3DMesh> meshes;
  std::vector<uint32_t> geo_of_body;  // N
  std::vector<uint8_t> fixed;         // N
  std::vector<double> m;              // N total mass
  std::vector<V3> I0;                 // N body-frame inertia
  V3 g{ 0.0, 0.0, 0.0 };
  std::vector<V3> plane_x;
  std::vector<V3> plane_n;            // already normalised
  // static cylinders (rigidbody3d/StaticGeometry/StaticCylinder.cpp:8-19): point on the axis, unit axis, radius
  std::vector<V3> cyl_x;
  std::vector<V3> cyl_axis;           // already normalised
  std::vector<double> cyl_r;
  std::size_t nbodies() const { return geo_of_body.size(); }
  const RB3DGeometry& geo( const std::size_t b ) const { return geometry[geo_of_body[b]]; }
};

inline M3 loadR( const double* q, const std::size_t nb, const std::size_t b )
{
  M3 R;
  for( int k = 0; k < 9; ++k ) { R.m[k] = q[3 * nb + 9 * b + k]; }
  return R;
}
inline V3 loadX( const double* q, const std::size_t b ) { return V3{ q[3 * b], q[3 * b + 1], q[3 * b + 2] }; }

// R * diag(d) * R^T as Eigen evaluates it: A = R * d.asDiagonal() (column scaling), then A * R^T
inline M3 worldInertia( const M3& R, const V3& d )
{
  M3 A;
  for( int r = 0; r < 3; ++r ) { A.m[3 * r + 0] = R.m[3 * r + 0] * d.x; A.m[3 * r + 1] = R.m[3 * r + 1] * d.y; A.m[3 * r + 2] = R.m[3 * r + 2] * d.z; }
  M3 B;
  for( int r = 0; r < 3; ++r )
    for( int c = 0; c < 3; ++c )
      B.m[3 * r + c] = ( A.m[3 * r + 0] * R.m[3 * c + 0] + A.m[3 * r + 1] * R.m[3 * c + 1] ) + A.m[3 * r + 2] * R.m[3 * c + 2];
  return B;
}

// AngleAxis( angle, axis ).toRotationMatrix() (Eigen/src/Geometry/AngleAxis.h)
inline M3 angleAxisMatrix( const double angle, const V3& axis )
{
  M3 res;
  const V3 sin_axis = std::sin( angle ) * axis;
  const double c = std::cos( angle );
  const V3 cos1_axis = ( 1.0 - c ) * axis;
  double tmp;
  tmp = cos1_axis.x * axis.y;
  res.m[1] = tmp - sin_axis.z; res.m[3] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z;
  res.m[2] = tmp + sin_axis.y; res.m[6] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z;
  res.m[5] = tmp - sin_axis.x; res.m[7] = tmp + sin_axis.x;
  res.m[0] = cos1_axis.x * axis.x + c;
  res.m[4] = cos1_axis.y * axis.y + c;
  res.m[8] = cos1_axis.z * axis.z + c;
  return res;
}

struct Quat { double w, x, y, z; };

// Eigen::Quaternion( Matrix3 ) -- Shoemake
inline Quat quatFromMatrix( const M3& m )
{
  Quat q;
  auto M = [&]( int r, int c ) { return m.m[3 * r + c]; };
  double t = M( 0, 0 ) + ( M( 1, 1 ) + M( 2, 2 ) ); // diagonal().sum(): not vectorisable => unrolled tree a0 + (a1 + a2)
  if( t > 0.0 )
  {
    t = std::sqrt( t + 1.0 );
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = ( M( 2, 1 ) - M( 1, 2 ) ) * t;
    q.y = ( M( 0, 2 ) - M( 2, 0 ) ) * t;
    q.z = ( M( 1, 0 ) - M( 0, 1 ) ) * t;
  }
  else
  {
    int i = 0;
    if( M( 1, 1 ) > M( 0, 0 ) ) { i = 1; }
    if( M( 2, 2 ) > M( i, i ) ) { i = 2; }
    const int j = ( i + 1 ) % 3;
    const int k = ( j + 1 ) % 3;
    t = std::sqrt( M( i, i ) - M( j, j ) - M( k, k ) + 1.0 );
    double c[3];
    c[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = ( M( k, j ) - M( j, k ) ) * t;
    c[j] = ( M( j, i ) + M( i, j ) ) * t;
    c[k] = ( M( k, i ) + M( i, k ) ) * t;
    q.x = c[0]; q.y = c[1]; q.z = c[2];
  }
  return q;
}

inline M3 quatToMatrix( const Quat& q )
{
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r.m[0] = 1.0 - ( tyy + tzz ); r.m[1] = txy - twz; r.m[2] = txz + twy;
  r.m[3] = txy + twz; r.m[4] = 1.0 - ( txx + tzz ); r.m[5] = tyz - twx;
  r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1.0 - ( txx + tyy );
  return r;
}

// DMVMap.cpp:15-60
inline void solveDMV( const V3& am, const double h, const V3& I0, Quat& q )
{
  const double eps = std::fabs( h * 1e-15 );
  const double ha = h / 2.0;
  const double fac1 = ( I0.y - I0.z ) / I0.x;
  const double fac2 = ( I0.z - I0.x ) / I0.y;
  const double fac3 = ( I0.x - I0.y ) / I0.z;
  const double am1i = am.x * ha / I0.x;
  const double am2i = am.y * ha / I0.y;
  const double am3i = am.z * ha / I0.z;
  double cm1 = am1i + fac1 * am2i * am3i;
  double cm2 = am2i + fac2 * cm1 * am3i;
  double cm3 = am3i + fac3 * cm1 * cm2;
  for( unsigned itr = 0; itr < 50; ++itr )
  {
    const double cm1b = cm1, cm2b = cm2, cm3b = cm3;
    const double calpha = cm1 * cm1 + 1.0 + cm2 * cm2 + cm3 * cm3;
    cm1 = calpha * am1i + fac1 * cm2 * cm3;
    cm2 = calpha * am2i + fac2 * cm1 * cm3;
    cm3 = calpha * am3i + fac3 * cm1 * cm2;
    const double err = std::fabs( cm1b - cm1 ) + std::fabs( cm2b - cm2 ) + std::fabs( cm3b - cm3 );
    if( err <= eps ) { break; }
  }
  const double q0 = q.w, q1 = q.x, q2 = q.y, q3 = q.z;
  q.w = q0 - cm1 * q1 - cm2 * q2 - cm3 * q3;
  q.x = q1 + cm1 * q0 + cm3 * q2 - cm2 * q3;
  q.y = q2 + cm2 * q0 + cm1 * q3 - cm3 * q1;
  q.z = q3 + cm3 * q0 + cm2 * q1 - cm1 * q2;
  // Quaternion::normalize(): coeffs() /= norm(), norm = sqrt( coeffs.cwiseAbs2().sum() ); coeffs are stored ( x, y, z, w ).  Eigen 3.3.4
  // sums a fixed 4-vector of doubles by packets (Core/Redux.h: redux_impl<LinearVectorizedTraversal, CompleteUnrolling>): with SSE2 two
  // packets ( x2, y2 ) + ( z2, w2 ) are added lane-wise and then across (redux_vec_unroller + predux), with AVX the one packet is folded
  // high half onto low half (predux<Packet4d>) -- either way ( x2 + z2 ) + ( y2 + w2 )
  const double nrm = std::sqrt( ( q.x * q.x + q.z * q.z ) + ( q.y * q.y + q.w * q.w ) );
  q.x /= nrm; q.y /= nrm; q.z /= nrm; q.w /= nrm;
}

// kind: 2 = split_ham, 3 = dmv.  q: [3N x | 9N R row-major], v: [3N lin | 3N ang]
// m_updated: which of the reference's two world-space mass matrices multiplies v0.  RigidBody3DState's constructor stores the
// inertia block transposed, M(r,c) = I(c,r) ( formWorldSpaceMassMatrix, RigidBody3DState.cpp:165-182 ), and that is what the first
// flow of a simulation reads; RigidBody3DSim::flow then calls updateMandMinv ( RigidBody3DSim.cpp:442,517,592 ), which assigns
// I = R I0 R^T through a column-major map over the same values ( RigidBody3DState.cpp:444-446 ), so every later flow reads
// M(r,c) = I(r,c).  I is symmetric only up to rounding: ( R(r,k) d_k ) R(c,k) vs ( R(c,k) d_k ) R(r,k).
inline void flowExponentialEuler( const RB3DScene& s, const double* q0, const double* v0, const double dt, double* q1, double* v1 );
inline void flow( const int kind, const RB3DScene& s, const double* q0, const double* v0, const double dt, double* q1, double* v1, const bool m_updated = false )
{
  if( kind == 4 ) { flowExponentialEuler( s, q0, v0, dt, q1, v1 ); return; }
  const std::size_t nb = s.nbodies();
  for( std::size_t k = 0; k < 12 * nb; ++k ) { q1[k] = q0[k]; }
  for( std::size_t b = 0; b < nb; ++b )
  {
    const double m = s.m[b];
    const M3 R0 = loadR( q0, nb, b );
    const V3 vl{ v0[3 * b], v0[3 * b + 1], v0[3 * b + 2] };
    const V3 w0{ v0[3 * nb + 3 * b], v0[3 * nb + 3 * b + 1], v0[3 * nb + 3 * b + 2] };
    // v1 = M * v0 (sparse column-major accumulate into zero). World inertia block stored transposed: M(r,c) = I(c,r)
    V3 p{ 0.0 + m * vl.x, 0.0 + m * vl.y, 0.0 + m * vl.z };
    const M3 Iw = worldInertia( R0, s.I0[b] );
    V3 L{ ( ( 0.0 + Iw.m[0] * w0.x ) + Iw.m[3] * w0.y ) + Iw.m[6] * w0.z,
          ( ( 0.0 + Iw.m[1] * w0.x ) + Iw.m[4] * w0.y ) + Iw.m[7] * w0.z,
          ( ( 0.0 + Iw.m[2] * w0.x ) + Iw.m[5] * w0.y ) + Iw.m[8] * w0.z };
    if( m_updated )
    {
      L = V3{ ( ( 0.0 + Iw.m[0] * w0.x ) + Iw.m[1] * w0.y ) + Iw.m[2] * w0.z,
              ( ( 0.0 + Iw.m[3] * w0.x ) + Iw.m[4] * w0.y ) + Iw.m[5] * w0.z,
              ( ( 0.0 + Iw.m[6] * w0.x ) + Iw.m[7] * w0.y ) + Iw.m[8] * w0.z };
    }
    if( s.fixed[b] )
    {
      v1[3 * b] = p.x; v1[3 * b + 1] = p.y; v1[3 * b + 2] = p.z;
      v1[3 * nb + 3 * b] = L.x; v1[3 * nb + 3 * b + 1] = L.y; v1[3 * nb + 3 * b + 2] = L.z;
      continue;
    }
    // F.setZero(); F_lin += m * g; torque stays 0
    const V3 F{ 0.0 + m * s.g.x, 0.0 + m * s.g.y, 0.0 + m * s.g.z };
    const double hdt = 0.5 * dt;
    p = V3{ p.x + hdt * F.x, p.y + hdt * F.y, p.z + hdt * F.z };
    L = V3{ L.x + hdt * 0.0, L.y + hdt * 0.0, L.z + hdt * 0.0 };
    // q_update = dt * v0 + 0.5 * dt * dt * Minv * F
    const double sc = ( ( 0.5 * dt ) * dt ) * ( 1.0 / m );
    const V3 qu{ dt * vl.x + ( 0.0 + sc * F.x ), dt * vl.y + ( 0.0 + sc * F.y ), dt * vl.z + ( 0.0 + sc * F.z ) };
    q1[3 * b] = q0[3 * b] + qu.x; q1[3 * b + 1] = q0[3 * b + 1] + qu.y; q1[3 * b + 2] = q0[3 * b + 2] + qu.z;

    const V3 I = s.I0[b];
    M3 R1;
    if( kind == 2 )
    {
      V3 pB = mulT( R0, L );
      // five sub-rotations about -z, -y, -x, -y, -z; R1 = R * AngleAxis( -angle, axis ); pB = AngleAxis( angle, axis ) * pB
      const V3 axes[5] = { V3{ -0.0, -0.0, -1.0 }, V3{ -0.0, -1.0, -0.0 }, V3{ -1.0, -0.0, -0.0 }, V3{ -0.0, -1.0, -0.0 }, V3{ -0.0, -0.0, -1.0 } };
      R1 = R0;
      for( int st = 0; st < 5; ++st )
      {
        double angle;
        if( st == 0 || st == 4 ) { angle = 0.5 * dt * pB.z / I.z; }
        else if( st == 1 || st == 3 ) { angle = 0.5 * dt * pB.y / I.y; }
        else { angle = dt * pB.x / I.x; }
        R1 = mul( R1, angleAxisMatrix( -angle, axes[st] ) );
        pB = mul( angleAxisMatrix( angle, axes[st] ), pB );
      }
    }
    else
    {
      const V3 p_i = mulT( R0, L );
      Quat Q = quatFromMatrix( R0 );
      solveDMV( p_i, dt, I, Q );
      R1 = quatToMatrix( Q );
    }
    for( int k = 0; k < 9; ++k ) { q1[3 * nb + 9 * b + k] = R1.m[k]; }
    // second kick (gravity does not depend on q)
    p = V3{ p.x + hdt * F.x, p.y + hdt * F.y, p.z + hdt * F.z };
    L = V3{ L.x + hdt * 0.0, L.y + hdt * 0.0, L.z + hdt * 0.0 };
    v1[3 * b] = p.x / m; v1[3 * b + 1] = p.y / m; v1[3 * b + 2] = p.z / m;
    // omega = R * Iinv0.asDiagonal() * R^T * L
    const V3 Iinv{ 1.0 / I.x, 1.0 / I.y, 1.0 / I.z };
    const V3 w1 = mul( worldInertia( R1, Iinv ), L );
    v1[3 * nb + 3 * b] = w1.x; v1[3 * nb + 3 * b + 1] = w1.y; v1[3 * nb + 3 * b + 2] = w1.z;
  }
}

// ---- ExponentialEulerMap (rigidbody3d/UnconstrainedMaps/ExponentialEulerMap.cpp:13-91) ---------------------------------
// projectOrientation (:13-32): R <- U V^T of R's singular value decomposition, i.e. the orthogonal factor of its polar
// decomposition.  The reference gets U, V from Eigen::JacobiSVD; that factor is unique for a non-singular R, so any accurate
// SVD gives it to rounding.  PARITY UNPINNED BIT FOR BIT (Eigen's sweep order is not restated): the GPU is held to 1e-12.
// Here: cyclic one-sided Jacobi (Hestenes) on the columns, a different algorithm from the kernel's Newton iteration on purpose.
inline M3 polarOrthogonalFactor( const M3& A )
{
  // columns of W are rotated until mutually orthogonal: A V = W, W = U S  =>  U V^T = W S^-1 V^T
  double W[3][3], V[3][3];
  for( int r = 0; r < 3; ++r ) { for( int c = 0; c < 3; ++c ) { W[r][c] = A.m[3 * r + c]; V[r][c] = ( r == c ) ? 1.0 : 0.0; } }
  for( int sweep = 0; sweep < 60; ++sweep )
  {
    double off = 0.0;
    for( int p = 0; p < 2; ++p )
    {
      for( int q = p + 1; q < 3; ++q )
      {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for( int r = 0; r < 3; ++r ) { alpha += W[r][p] * W[r][p]; beta += W[r][q] * W[r][q]; gamma += W[r][p] * W[r][q]; }
        if( gamma == 0.0 ) { continue; }
        off = std::max( off, std::fabs( gamma ) / std::sqrt( alpha * beta ) );
        const double zeta = ( beta - alpha ) / ( 2.0 * gamma );
        const double t = ( ( zeta >= 0.0 ) ? 1.0 : -1.0 ) / ( std::fabs( zeta ) + std::sqrt( 1.0 + zeta * zeta ) );
        const double c = 1.0 / std::sqrt( 1.0 + t * t ), sn = c * t;
        for( int r = 0; r < 3; ++r )
        {
          const double wp = W[r][p], wq = W[r][q];
          W[r][p] = c * wp - sn * wq; W[r][q] = sn * wp + c * wq;
          const double vp = V[r][p], vq = V[r][q];
          V[r][p] = c * vp - sn * vq; V[r][q] = sn * vp + c * vq;
        }
      }
    }
    if( off < 1.0e-17 ) { break; }
  }
  M3 out;
  for( int r = 0; r < 3; ++r ) { for( int c = 0; c < 3; ++c ) { out.m[3 * r + c] = 0.0; } }
  for( int k = 0; k < 3; ++k )
  {
    double sk = 0.0;
    for( int r = 0; r < 3; ++r ) { sk += W[r][k] * W[r][k]; }
    sk = std::sqrt( sk );
    for( int r = 0; r < 3; ++r ) { for( int c = 0; c < 3; ++c ) { out.m[3 * r + c] += ( W[r][k] / sk ) * V[c][k]; } }
  }
  return out;
}

// flow (:34-91).  x1 = x0 + dt v0; every column of R advanced by dt omega x column, then projected; A = Minv * F with F the
// gravity force (angular part zero: the 3x3 inverse-inertia block times zeros accumulates to +0.0); v1 = v0 + dt A.
// Kinematically scripted bodies are integrated like the rest (the reference only asserts there are none).
inline void flowExponentialEuler( const RB3DScene& s, const double* q0, const double* v0, const double dt, double* q1, double* v1 )
{
  const std::size_t nb = s.nbodies();
  for( std::size_t b = 0; b < nb; ++b )
  {
    const double m = s.m[b];
    const M3 R0 = loadR( q0, nb, b );
    const V3 w{ v0[3 * nb + 3 * b], v0[3 * nb + 3 * b + 1], v0[3 * nb + 3 * b + 2] };
    for( int k = 0; k < 3; ++k ) { q1[3 * b + k] = q0[3 * b + k] + dt * v0[3 * b + k]; }
    M3 R1;
    for( int j = 0; j < 3; ++j )
    {
      const V3 c{ R0.m[j], R0.m[3 + j], R0.m[6 + j] };
      // omega.cross( c )  (Eigen cross3: a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0)
      const V3 x{ w.y * c.z - w.z * c.y, w.z * c.x - w.x * c.z, w.x * c.y - w.y * c.x };
      R1.m[j] = c.x + dt * x.x; R1.m[3 + j] = c.y + dt * x.y; R1.m[6 + j] = c.z + dt * x.z;
    }
    const M3 P = polarOrthogonalFactor( R1 );
    for( int k = 0; k < 9; ++k ) { q1[3 * nb + 9 * b + k] = P.m[k]; }
    const double minv = 1.0 / m;
    const V3 F{ 0.0 + m * s.g.x, 0.0 + m * s.g.y, 0.0 + m * s.g.z };
    const V3 A{ 0.0 + minv * F.x, 0.0 + minv * F.y, 0.0 + minv * F.z };
    v1[3 * b] = v0[3 * b] + dt * A.x; v1[3 * b + 1] = v0[3 * b + 1] + dt * A.y; v1[3 * b + 2] = v0[3 * b + 2] + dt * A.z;
    const double Aang = ( ( 0.0 + 0.0 ) + 0.0 ) + 0.0;
    for( int k = 0; k < 3; ++k ) { v1[3 * nb + 3 * b + k] = v0[3 * nb + 3 * b + k] + dt * Aang; }
  }
}

inline void computeAABB( const RB3DScene& s, const std::size_t b, const V3& cm, const M3& R, Box<3>& box )
{
  const RB3DGeometry& g = s.geo( b );
  if( g.type == GEO_SPHERE )
  {
    box.lo[0] = cm.x - g.r; box.lo[1] = cm.y - g.r; box.lo[2] = cm.z - g.r;
    box.hi[0] = cm.x + g.r; box.hi[1] = cm.y + g.r; box.hi[2] = cm.z + g.r;
  }
  else if( g.type == GEO_BOX )
  {
    M3 A;
    for( int k = 0; k < 9; ++k ) { A.m[k] = std::fabs( R.m[k] ); }
    const V3 e = mul( A, g.half );
    box.lo[0] = cm.x - e.x; box.lo[1] = cm.y - e.y; box.lo[2] = cm.z - e.z;
    box.hi[0] = cm.x + e.x; box.hi[1] = cm.y + e.y; box.hi[2] = cm.z + e.z;
  }
  else
  {
    for( int k = 0; k < 3; ++k ) { box.lo[k] = std::numeric_limits<double>::infinity(); box.hi[k] = -std::numeric_limits<double>::infinity(); }
    for( const V3& vert : s.meshes[g.mesh].verts )
    {
      const V3 t = mul( R, vert ) + cm;
      box.lo[0] = std::min( box.lo[0], t.x ); box.lo[1] = std::min( box.lo[1], t.y ); box.lo[2] = std::min( box.lo[2], t.z );
      box.hi[0] = std::max( box.hi[0], t.x ); box.hi[1] = std::max( box.hi[1], t.y ); box.hi[2] = std::max( box.hi[2], t.z );
    }
  }
}

inline V3 boxCorner( const V3& half, const int i )
{
  return V3{ half.x * double( 2 * ( i % 2 ) - 1 ), half.y * double( 2 * ( ( i >> 1 ) % 2 ) - 1 ), half.z * double( 2 * ( ( i >> 2 ) % 2 ) - 1 ) };
}

// One un-teleported candidate pair: the kinematic-kinematic skip (RigidBody3DSim.cpp:1134-1137) and
// RigidBody3DSim::dispatchNarrowPhaseCollision (:879-962); appends to active_set; returns false if the reference would
// exit( EXIT_FAILURE ) on an unsupported pair of geometry types
inline bool dispatchNarrowPhaseCollision( const RB3DScene& s, const unsigned first, const unsigned second, const double* q0, const double* q1, std::vector<RB3DContact>& active_set )
{
  const std::size_t nb = s.nbodies();
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  {
    {
      if( s.fixed[first] && s.fixed[second] ) { return true; }
      unsigned b0 = first, b1 = second;
      if( s.fixed[b0] ) { std::swap( b0, b1 ); }
      const RB3DGeometry& g0 = s.geo( b0 );
      const RB3DGeometry& g1 = s.geo( b1 );
      const bool kin = s.fixed[b1] != 0;
      if( g0.type == GEO_SPHERE && g1.type == GEO_SPHERE )
      {
        const V3 d1 = loadX( q1, b0 ) - loadX( q1, b1 );
        if( squaredNorm( d1 ) <= ( g0.r + g1.r ) * ( g0.r + g1.r ) )
        {
          RB3DContact c;
          c.i = b0; c.j = b1; c.aux = 0;
          const V3 x0a = loadX( q0, b0 ), x0b = loadX( q0, b1 );
          c.n = normalized( x0a - x0b );
          if( !kin )
          {
            c.type = SPHERE_SPHERE;
            c.p = x0a + ( g0.r / ( g0.r + g1.r ) ) * ( x0b - x0a );
          }
          else
          {
            c.type = KINEMATIC_SPHERE_SPHERE;
            c.p = x0b; // the kinematic sphere's centre handed to KinematicSphereSphereConstraint
          }
          c.depth = std::min( 0.0, norm( d1 ) - g0.r - g1.r );
          active_set.emplace_back( c );
        }
      }
      else if( g0.type == GEO_BOX && g1.type == GEO_BOX )
      {
        V3 n;
        std::vector<V3> points;
        boxbox::isActive( loadX( q1, b0 ), loadR( q1, nb, b0 ), g0.half, loadX( q1, b1 ), loadR( q1, nb, b1 ), g1.half, n, points );
        for( const V3& pt : points )
        {
          RB3DContact c;
          c.type = kin ? KINEMATIC_BODY_BODY : BODY_BODY; c.i = b0; c.j = b1; c.aux = 0; c.n = n; c.p = pt; c.depth = NaN;
          active_set.emplace_back( c );
        }
      }
      else if( g0.type == GEO_MESH && g1.type == GEO_MESH )
      {
        const RB3DMesh& mesh0 = s.meshes[g0.mesh];
        const RB3DMesh& mesh1 = s.meshes[g1.mesh];
        const V3 cm0 = loadX( q1, b0 ), cm1 = loadX( q1, b1 );
        const M3 R0 = loadR( q1, nb, b0 ), R1 = loadR( q1, nb, b1 );
        auto push = [&]( const V3& p, const V3& n )
        {
          RB3DContact c;
          c.type = kin ? KINEMATIC_BODY_BODY : BODY_BODY; c.i = b0; c.j = b1; c.aux = 0; c.n = n; c.p = p; c.depth = NaN;
          active_set.emplace_back( c );
        };
        {
          const M3 R01 = mulTN( R1, R0 );
          const V3 x01 = mulT( R1, cm0 - cm1 );
          for( const V3& smp : mesh0.samples )
          {
            const V3 x = mul( R01, smp ) + x01;
            V3 normal;
            if( mesh1.detectCollision( x, normal ) ) { push( mul( R1, x ) + cm1, mul( R1, normal ) ); }
          }
        }
        {
          const M3 R10 = mulTN( R0, R1 );
          const V3 x10 = mulT( R0, cm1 - cm0 );
          for( const V3& smp : mesh1.samples )
          {
            const V3 x = mul( R10, smp ) + x10;
            V3 normal;
            if( mesh0.detectCollision( x, normal ) ) { push( mul( R0, x ) + cm0, -mul( R0, normal ) ); }
          }
        }
      }
      else
      {
        return false;
      }
    }
  }
  return true;
}

// RigidBody3DSim::computeBodyPlaneActiveSetAllPairs / computeBodyCylinderActiveSetAllPairs (RigidBody3DSim.cpp:1414-1557); appends;
// returns false where the reference exits (cylinder vs box)
inline bool computeStaticActiveSet( const RB3DScene& s, const double* q0, const double* q1, std::vector<RB3DContact>& active_set )
{
  const std::size_t nb = s.nbodies();
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  // planes: plane-major, body ascending, corner / hull vertex ascending; kinematic bodies skipped
  for( uint32_t pl = 0; pl < uint32_t( s.plane_x.size() ); ++pl )
  {
    const V3 xp = s.plane_x[pl], np = s.plane_n[pl];
    for( uint32_t b = 0; b < uint32_t( nb ); ++b )
    {
      if( s.fixed[b] ) { continue; }
      const RB3DGeometry& g = s.geo( b );
      const V3 x1 = loadX( q1, b ), x0 = loadX( q0, b );
      if( g.type == GEO_BOX )
      {
        const M3 R1 = loadR( q1, nb, b ), R0 = loadR( q0, nb, b );
        for( int corner = 0; corner < 8; ++corner )
        {
          const V3 wp = mul( R1, boxCorner( g.half, corner ) ) + x1;
          if( dot( np, wp - xp ) <= 0.0 )
          {
            RB3DContact c;
            c.type = PLANE_BOX; c.i = b; c.j = pl; c.aux = uint32_t( corner ); c.n = np;
            c.p = mul( R0, boxCorner( g.half, corner ) ) + x0;
            c.depth = NaN;
            active_set.emplace_back( c );
          }
        }
      }
      else if( g.type == GEO_SPHERE )
      {
        const double d = dot( np, x1 - xp );
        if( d <= g.r )
        {
          RB3DContact c;
          c.type = PLANE_SPHERE; c.i = b; c.j = pl; c.aux = 0; c.n = np;
          c.p = x0 - g.r * np;
          c.depth = std::min( 0.0, d - g.r );
          active_set.emplace_back( c );
        }
      }
      else
      {
        const RB3DMesh& mesh = s.meshes[g.mesh];
        const M3 R1 = loadR( q1, nb, b ), R0 = loadR( q0, nb, b );
        for( uint32_t vi = 0; vi < uint32_t( mesh.hull.size() ); ++vi )
        {
          const V3 v = mul( R1, mesh.hull[vi] ) + x1;
          if( dot( np, v - xp ) <= 0.0 )
          {
            RB3DContact c;
            c.type = PLANE_BODY; c.i = b; c.j = pl; c.aux = vi; c.n = np;
            c.p = x0 + mul( R0, mesh.hull[vi] );
            c.depth = NaN;
            active_set.emplace_back( c );
          }
        }
      }
    }
  }
  // cylinders: cylinder-major, body ascending, hull vertex ascending; kinematic bodies skipped; boxes are not supported
  // (RigidBody3DSim::computeBodyCylinderActiveSetAllPairs, RigidBody3DSim.cpp:1504-1557)
  for( uint32_t cy = 0; cy < uint32_t( s.cyl_x.size() ); ++cy )
  {
    const V3 xc = s.cyl_x[cy], ax = s.cyl_axis[cy];
    const double rc = s.cyl_r[cy];
    for( uint32_t b = 0; b < uint32_t( nb ); ++b )
    {
      if( s.fixed[b] ) { continue; }
      const RB3DGeometry& g = s.geo( b );
      const V3 x1 = loadX( q1, b ), x0 = loadX( q0, b );
      // StaticCylinder{Sphere,Body}Constraint::computeN at q0 (StaticCylinderSphereConstraint.cpp:237-244,
      // StaticCylinderBodyConstraint.cpp:84-91): minus the part of (x - xc) perpendicular to the axis, normalised
      const V3 e0 = ( x0 - xc ) - dot( ax, x0 - xc ) * ax;
      const V3 m0 = V3{ -e0.x, -e0.y, -e0.z };
      if( g.type == GEO_SPHERE )
      {
        // StaticCylinderSphereConstraint::isActive (StaticCylinderSphereConstraint.cpp:10-19) at q1
        const V3 d = ( x1 - xc ) - dot( ax, x1 - xc ) * ax;
        if( dot( d, d ) >= ( rc - g.r ) * ( rc - g.r ) )
        {
          RB3DContact c;
          c.type = CYLINDER_SPHERE; c.i = b; c.j = cy; c.aux = 0;
          c.n = normalized( m0 );
          c.p = x0 - g.r * c.n;         // getWorldSpaceContactPoint (:324-327)
          c.depth = std::min( 0.0, rc - norm( d ) - g.r ); // computePenetrationDepth at q1 (:311-317) -- found by running the reference's own RigidBody3DSim
          active_set.emplace_back( c );
        }
      }
      else if( g.type == GEO_MESH )
      {
        // MeshMeshUtilities::computeMeshCylinderActiveSet (MeshMeshUtilities.cpp:88-109) at q1
        const RB3DMesh& mesh = s.meshes[g.mesh];
        const M3 R1 = loadR( q1, nb, b ), R0 = loadR( q0, nb, b );
        for( uint32_t vi = 0; vi < uint32_t( mesh.hull.size() ); ++vi )
        {
          const V3 v = mul( R1, mesh.hull[vi] ) + x1;
          const V3 d = ( v - xc ) - dot( ax, v - xc ) * ax;
          if( dot( d, d ) >= rc * rc )
          {
            RB3DContact c;
            c.type = CYLINDER_BODY; c.i = b; c.j = cy; c.aux = vi;
            const double nrm = std::sqrt( dot( m0, m0 ) );   // n / n.norm()
            c.n = V3{ m0.x / nrm, m0.y / nrm, m0.z / nrm };
            c.p = x0 + mul( R0, mesh.hull[vi] );
            c.depth = NaN;
            active_set.emplace_back( c );
          }
        }
      }
      else
      {
        return false; // "Collision between static cylinders and box not supported. Exiting."
      }
    }
  }
  return true;
}

// returns false if the reference would exit( EXIT_FAILURE ) on an unsupported pair of geometry types
inline bool computeActiveSet( const RB3DScene& s, const double* q0, const double* q1, std::vector<RB3DContact>& active_set,
                              std::vector<std::pair<unsigned,unsigned>>* candidates_out = nullptr, const bool use_grid = true )
{
  const std::size_t nb = s.nbodies();
  active_set.clear();
  if( nb > 0 )
  {
    PairSet possible_overlaps;
    {
      std::vector<Box<3>> aabbs( nb );
      for( std::size_t b = 0; b < nb; ++b ) { computeAABB( s, b, loadX( q1, b ), loadR( q1, nb, b ), aabbs[b] ); }
      if( use_grid ) { getPotentialOverlaps<3>( aabbs, possible_overlaps ); }
      else { getPotentialOverlapsAllPairs<3>( aabbs, possible_overlaps ); }
    }
    if( candidates_out != nullptr ) { candidates_out->assign( possible_overlaps.begin(), possible_overlaps.end() ); }
    for( const auto& pr : possible_overlaps )
    {
      if( !dispatchNarrowPhaseCollision( s, pr.first, pr.second, q0, q1, active_set ) ) { return false; }
    }
  }
  return computeStaticActiveSet( s, q0, q1, active_set );
}

}

// RigidBody3DState::updateMandMinv ( rigidbody3d/RigidBody3DState.cpp:428-462 ): per body the 3 x 3 blocks I = R I0 R^T and
// Iinv = R Iinv0 R^T as they land in the value arrays of M and Minv ( column-major maps: entry ( r, c ) at 3 c + r ), Iinv0 = 1 / I0
// ( formBodySpaceInverseMassMatrix, RigidBody3DState.cpp:116-133 ).  q: [3N x | 9N R row-major]; outputs 9 doubles per body.
inline void updateMandMinv( const RB3DScene& s, const double* q, double* I_blocks, double* Iinv_blocks )
{
  const std::size_t nb = s.nbodies();
  for( std::size_t b = 0; b < nb; ++b )
  {
    const M3 R = loadR( q, nb, b );
    const V3 I0 = s.I0[b];
    const M3 I = worldInertia( R, I0 );
    const M3 Ii = worldInertia( R, V3{ 1.0 / I0.x, 1.0 / I0.y, 1.0 / I0.z } );
    for( int r = 0; r < 3; ++r )
      for( int c = 0; c < 3; ++c ) { I_blocks[9 * b + 3 * c + r] = I.m[3 * r + c]; Iinv_blocks[9 * b + 3 * c + r] = Ii.m[3 * r + c]; }
  }
}

#endif
