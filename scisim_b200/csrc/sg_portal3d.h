// sg_portal3d.h -- planar portal arithmetic of rigidbody3d, shared by the kernels and the host code of libscisim_b200
// (plain C++ when compiled without nvcc: tests/portal_kernel_harness.cpp runs the same functions on the CPU).
//
// Reference behaviour reproduced (file:line in the SCISim checkout):
//   rigidbody3d/StaticGeometry/StaticPlane.cpp:10-15,24-27,48-66  n normalised, distanceToPoint, t0 / t1 = FromTwoVectors( UnitY, n ) * UnitX / UnitZ
//   Eigen 3.3.4 Quaternion::setFromTwoVectors / _transformVector    (sg_rotate_from_unit_y; the reference's un-vendored dependency)
//   rigidbody3d/Portals/PlanarPortal.cpp:107-185                    pointInsidePortal, aabbInHalfPlane, aabbTouchesPortal (release build: plane A
//                                                                   first), teleportPointInsidePortal, teleportPointThroughPlaneA / B with the
//                                                                   integer portal multiplier
// FP64 throughout; every 3-term reduction is ( a0*b0 + a1*b1 ) + a2*b2; the library is compiled without FMA contraction.
#ifndef SG_PORTAL3D_H
#define SG_PORTAL3D_H

#include "sg_portal2d.h" // SG_HD, SG_MAX_PORTALS, SG_NO_PORTAL, SG_PLANE_B_BIT, SgTeleCollision, the pair sort helpers

struct SgVec3 { double x, y, z; };

// One PlanarPortal: two StaticPlanes (point, unit normal, two tangents) and the integer multipliers of the three coordinates
struct SgPortal3D
{
  double ax[3], an[3], at0[3], at1[3];
  double bx[3], bn[3], bt0[3], bt1[3];
  int mult[3];
  int pad;
};

struct SgPortals3D
{
  uint32_t n;
  uint32_t pad;
  SgPortal3D p[SG_MAX_PORTALS];
};

SG_HD double sg_dot3( const double* a, const double bx, const double by, const double bz ) { return ( a[0] * bx + a[1] * by ) + a[2] * bz; }

// StaticPlane::distanceToPoint: n.dot( x - x_plane )
SG_HD double sg_plane3_dist( const double* px, const double* pn, const SgVec3 x ) { return sg_dot3( pn, x.x - px[0], x.y - px[1], x.z - px[2] ); }

// Quaternion::FromTwoVectors( UnitY, n ) * v as Eigen 3.3.4 evaluates it (both arguments are normalised again; the nearly-opposite
// branch needs an SVD and is not provided: ok = false)
SG_HD bool sg_rotate_from_unit_y( const double* n, const double* v, double* out )
{
  // v0 = UnitY.normalized() = ( 0, 1, 0 ) exactly ( 1 / sqrt( 1 ) ); v1 = n.normalized()
  double v1[3] = { n[0], n[1], n[2] };
  const double z = ( n[0] * n[0] + n[1] * n[1] ) + n[2] * n[2];
  if( z > 0.0 ) { const double s = sqrt( z ); v1[0] = n[0] / s; v1[1] = n[1] / s; v1[2] = n[2] / s; }
  const double c = ( v1[0] * 0.0 + v1[1] * 1.0 ) + v1[2] * 0.0;
  if( c < -1.0 + 1e-12 ) { return false; }
  // axis = v0.cross( v1 )
  const double axis[3] = { 1.0 * v1[2] - 0.0 * v1[1], 0.0 * v1[0] - 0.0 * v1[2], 0.0 * v1[1] - 1.0 * v1[0] };
  const double s = sqrt( ( 1.0 + c ) * 2.0 );
  const double invs = 1.0 / s;
  const double q[3] = { axis[0] * invs, axis[1] * invs, axis[2] * invs };
  const double w = s * 0.5;
  // uv = vec.cross( v ); uv += uv; v + w * uv + vec.cross( uv )
  double uv[3] = { q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0] };
  uv[0] = uv[0] + uv[0]; uv[1] = uv[1] + uv[1]; uv[2] = uv[2] + uv[2];
  const double cr[3] = { q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0] };
  out[0] = ( v[0] + w * uv[0] ) + cr[0]; out[1] = ( v[1] + w * uv[1] ) + cr[1]; out[2] = ( v[2] + w * uv[2] ) + cr[2];
  return true;
}

// StaticPlane( x, n ): n_out = n.normalized(), t0 = R * UnitX, t1 = R * UnitZ
SG_HD bool sg_portal3_plane_frame( const double* n_in, double* n_out, double* t0, double* t1 )
{
  const double z = ( n_in[0] * n_in[0] + n_in[1] * n_in[1] ) + n_in[2] * n_in[2];
  n_out[0] = n_in[0]; n_out[1] = n_in[1]; n_out[2] = n_in[2];
  if( z > 0.0 ) { const double s = sqrt( z ); n_out[0] = n_in[0] / s; n_out[1] = n_in[1] / s; n_out[2] = n_in[2] / s; }
  const double ux[3] = { 1.0, 0.0, 0.0 }, uz[3] = { 0.0, 0.0, 1.0 };
  return sg_rotate_from_unit_y( n_out, ux, t0 ) && sg_rotate_from_unit_y( n_out, uz, t1 );
}

// PlanarPortal::teleportPointThroughPlaneA (through_b == false) / ...B (true)
SG_HD SgVec3 sg_portal3_teleport( const SgPortal3D& p, const bool through_b, const SgVec3 xin )
{
  double fx[3], fn[3], ft0[3], ft1[3], tx[3], tn[3], tt0[3], tt1[3];
  for( int k = 0; k < 3; ++k )
  {
    fx[k] = through_b ? p.bx[k] : p.ax[k]; fn[k] = through_b ? p.bn[k] : p.an[k]; ft0[k] = through_b ? p.bt0[k] : p.at0[k]; ft1[k] = through_b ? p.bt1[k] : p.at1[k];
    tx[k] = through_b ? p.ax[k] : p.bx[k]; tn[k] = through_b ? p.an[k] : p.bn[k]; tt0[k] = through_b ? p.at0[k] : p.bt0[k]; tt1[k] = through_b ? p.at1[k] : p.bt1[k];
  }
  const double dx = fx[0] - xin.x, dy = fx[1] - xin.y, dz = fx[2] - xin.z;
  const double nA = double( p.mult[0] ) * sg_dot3( fn, dx, dy, dz );
  const double tA0 = double( p.mult[1] ) * sg_dot3( ft0, dx, dy, dz );
  const double tA1 = double( p.mult[2] ) * sg_dot3( ft1, dx, dy, dz );
  SgVec3 out;
  out.x = ( ( tx[0] + nA * tn[0] ) + tA0 * tt0[0] ) + tA1 * tt1[0];
  out.y = ( ( tx[1] + nA * tn[1] ) + tA0 * tt0[1] ) + tA1 * tt1[1];
  out.z = ( ( tx[2] + nA * tn[2] ) + tA0 * tt0[2] ) + tA1 * tt1[2];
  return out;
}

// aabbInHalfPlane (PlanarPortal.cpp:112-124): some corner has distanceToPoint <= 0
SG_HD bool sg_aabb3_in_half_plane( const double* px, const double* pn, const double* lo, const double* hi )
{
  for( int ix = 0; ix < 2; ++ix ) { for( int iy = 0; iy < 2; ++iy ) { for( int iz = 0; iz < 2; ++iz )
  {
    if( sg_plane3_dist( px, pn, SgVec3{ ix ? hi[0] : lo[0], iy ? hi[1] : lo[1], iz ? hi[2] : lo[2] } ) <= 0.0 ) { return true; }
  } } }
  return false;
}

// PlanarPortal::aabbTouchesPortal, release build: 0 = no, 1 = plane A, 2 = plane B
SG_HD int sg_portal3_aabb_touch( const SgPortal3D& p, const double* lo, const double* hi )
{
  if( sg_aabb3_in_half_plane( p.ax, p.an, lo, hi ) ) { return 1; }
  if( sg_aabb3_in_half_plane( p.bx, p.bn, lo, hi ) ) { return 2; }
  return 0;
}

SG_HD bool sg_portal3_point_inside( const SgPortal3D& p, const SgVec3 x ) { return sg_plane3_dist( p.ax, p.an, x ) < 0.0 || sg_plane3_dist( p.bx, p.bn, x ) < 0.0; }

// PlanarPortal::teleportPointInsidePortal
SG_HD SgVec3 sg_portal3_teleport_point_inside( const SgPortal3D& p, const SgVec3 x ) { return sg_portal3_teleport( p, !( sg_plane3_dist( p.ax, p.an, x ) < 0.0 ), x ); }

// RigidBody3DSim::enforcePeriodicBoundaryConditions for one centre of mass, portal-major (RigidBody3DSim.cpp:642-663)
SG_HD SgVec3 sg_portals3_enforce( const SgPortals3D& ps, SgVec3 x )
{
  for( uint32_t k = 0; k < ps.n; ++k ) { if( sg_portal3_point_inside( ps.p[k], x ) ) { x = sg_portal3_teleport_point_inside( ps.p[k], x ); } }
  return x;
}

// getTeleportedCollisionCenters for one body (RigidBody3DSim.cpp:1294-1336)
SG_HD SgVec3 sg_tele3_center( const SgPortals3D& ps, const uint32_t portal_word, const SgVec3 x )
{
  if( portal_word == SG_NO_PORTAL ) { return x; }
  return sg_portal3_teleport( ps.p[portal_word & ~SG_PLANE_B_BIT], ( portal_word & SG_PLANE_B_BIT ) != 0u, x );
}

// SphereSphereConstraint::isActive( x0, x1, r0, r1 )
SG_HD bool sg_sphere_sphere_active( const SgVec3 x0, const SgVec3 x1, const double r0, const double r1 )
{
  const double dx = x0.x - x1.x, dy = x0.y - x1.y, dz = x0.z - x1.z;
  return ( dx * dx + dy * dy ) + dz * dz <= ( r0 + r1 ) * ( r0 + r1 );
}

#endif
