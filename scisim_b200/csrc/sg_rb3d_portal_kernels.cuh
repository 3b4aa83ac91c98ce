// sg_rb3d_portal_kernels.cuh -- per-thread kernels of the rigidbody3d portal path, sphere scenes (driver: rb3d_portal_active_set_device
// in sg_rb3d.cu).  No shared memory, no barriers: tests/portal_kernel_harness.cpp runs these very functions on the CPU against the
// oracle.  The includer provides Rb3dDev ( n, btype, bparam: 4 doubles per body, radius first ), ContactOut3D, SG_FIXED_BIT and the SG_*
// codes; sg_portal3d.h holds the arithmetic.
//
// Reference behaviour reproduced (file:line in the SCISim checkout):
//   rigidbody3d/RigidBody3DSim.cpp:1072-1260  computeActiveSetBodyBodySpatialGrid: boxes at q1, a teleported box per (portal, body whose box
//                                             reaches a plane), un-teleported pairs -> dispatchNarrowPhaseCollision (the existing k_rb3d_pairs),
//                                             the others -> collisionIsActive / TeleportedCollision set
//   rigidbody3d/RigidBody3DSim.cpp:1262-1397  teleportedCollisionHappens, getTeleportedCollisionCenters, generateTeleportedCollision
//   rigidbody3d/Constraints/TeleportedSphereSphereConstraint.cpp:14-27,318-321, KinematicObjectSphereConstraint.cpp:10-22
//   rigidbody3d/RigidBody3DSim.cpp:642-663    enforcePeriodicBoundaryConditions
// Teleported collisions exist for spheres only in the reference (anything else exits); this path is restricted to all-sphere scenes.
#ifndef SG_RB3D_PORTAL_KERNELS_CUH
#define SG_RB3D_PORTAL_KERNELS_CUH

#include "sg_portal3d.h"
#include "sg_pair_sort.cuh"

// grid: ( blocks over bodies, portals ); aabbTouchesPortal on the body's box at q1 ( boxes: min(3), max(3) )
__global__ void __launch_bounds__( 256 ) k_r3p_touch( const __grid_constant__ SgPortals3D ps, const uint32_t n, const double* __restrict__ boxes, uint32_t* __restrict__ tflag )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const uint32_t p = blockIdx.y;
  const double* o = boxes + 6 * size_t( b );
  const double lo[3] = { __ldg( o ), __ldg( o + 1 ), __ldg( o + 2 ) }, hi[3] = { __ldg( o + 3 ), __ldg( o + 4 ), __ldg( o + 5 ) };
  tflag[size_t( p ) * n + b] = sg_portal3_aabb_touch( ps.p[p], lo, hi ) != 0 ? 1u : 0u;
}

// teleported boxes ( RigidBodySphere::computeAABB at the teleported centre ) and the TeleportedBody table (RigidBody3DSim.cpp:1090-1115)
__global__ void __launch_bounds__( 256 ) k_r3p_tele_boxes( const __grid_constant__ SgPortals3D ps, const Rb3dDev dev, const double* __restrict__ q1, const double* __restrict__ real_boxes,
                                                          const uint32_t* __restrict__ tflag, const uint32_t* __restrict__ toff, double* __restrict__ boxes, uint32_t* __restrict__ box_body,
                                                          uint32_t* __restrict__ box_portal )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= dev.n ) { return; }
  const uint32_t p = blockIdx.y;
  const size_t idx = size_t( p ) * dev.n + b;
  if( tflag[idx] == 0u ) { return; }
  const uint32_t slot = toff[idx];
  const double* rb = real_boxes + 6 * size_t( b );
  const double blo[3] = { __ldg( rb ), __ldg( rb + 1 ), __ldg( rb + 2 ) }, bhi[3] = { __ldg( rb + 3 ), __ldg( rb + 4 ), __ldg( rb + 5 ) };
  const int touch = sg_portal3_aabb_touch( ps.p[p], blo, bhi );
  const double* c = q1 + 3 * size_t( b );
  const SgVec3 xo = sg_portal3_teleport( ps.p[p], touch == 2, SgVec3{ __ldg( c ), __ldg( c + 1 ), __ldg( c + 2 ) } ); // teleportPoint( x, plane index, x_out )
  const double r = __ldg( &dev.bparam[4 * size_t( b )] );
  double* o = boxes + ( size_t( dev.n ) + slot ) * 6;
  o[0] = xo.x - r; o[1] = xo.y - r; o[2] = xo.z - r; o[3] = xo.x + r; o[4] = xo.y + r; o[5] = xo.z + r;
  box_body[slot] = b;
  box_portal[slot] = p | ( touch == 2 ? SG_PLANE_B_BIT : 0u );
}

// Per candidate of the extended box set (RigidBody3DSim.cpp:1126-1199).  Un-teleported pairs are compacted, in order, into reg_pairs
// for the regular narrow phase; pairs with a teleported member take the TeleportedCollision tests.
template<bool EMIT>
__global__ void __launch_bounds__( 128 ) k_r3p_classify( const __grid_constant__ SgPortals3D ps, const Rb3dDev dev, const uint2* __restrict__ pairs, const unsigned long long npairs,
                                                        const double* __restrict__ q1, const uint32_t* __restrict__ box_body, const uint32_t* __restrict__ box_portal,
                                                        uint32_t* __restrict__ reg_cnt, uint32_t* __restrict__ tel_cnt, const unsigned long long* __restrict__ reg_off, const uint32_t* __restrict__ tel_off,
                                                        uint2* __restrict__ reg_pairs, unsigned long long* __restrict__ tc_key, uint32_t* __restrict__ tc_idx, uint4* __restrict__ tc_info )
{
  const unsigned long long k = blockIdx.x * ( unsigned long long )( blockDim.x ) + threadIdx.x;
  if( k >= npairs ) { return; }
  const uint2 pr = pairs[k];
  const uint32_t n = dev.n;
  const bool first_teleported = pr.x >= n;
  const bool second_teleported = pr.y >= n;
  if( !first_teleported && !second_teleported )
  {
    if( !EMIT ) { reg_cnt[k] = 1u; tel_cnt[k] = 0u; }
    else { reg_pairs[reg_off[k]] = pr; }
    return;
  }
  uint32_t bdy0 = pr.x, bdy1 = pr.y, prtl0 = SG_NO_PORTAL, prtl1 = SG_NO_PORTAL;
  if( first_teleported ) { bdy0 = __ldg( &box_body[pr.x - n] ); prtl0 = __ldg( &box_portal[pr.x - n] ); }
  if( second_teleported ) { bdy1 = __ldg( &box_body[pr.y - n] ); prtl1 = __ldg( &box_portal[pr.y - n] ); }
  const SgTeleCollision c = sg_tele_collision( bdy0, bdy1, prtl0, prtl1 );
  if( EMIT )
  {
    if( tel_cnt[k] != 0u )
    {
      const uint32_t o = tel_off[k];
      tc_key[o] = ( ( unsigned long long )( c.b0 ) << 32 ) | c.b1;
      tc_idx[o] = o;
      tc_info[o] = make_uint4( c.b0, c.b1, c.p0, c.p1 );
    }
    return;
  }
  reg_cnt[k] = 0u;
  const bool f0 = ( __ldg( &dev.btype[c.b0] ) & SG_FIXED_BIT ) != 0u, f1 = ( __ldg( &dev.btype[c.b1] ) & SG_FIXED_BIT ) != 0u;
  const double r0 = __ldg( &dev.bparam[4 * size_t( c.b0 )] ), r1 = __ldg( &dev.bparam[4 * size_t( c.b1 )] );
  const SgVec3 xa{ __ldg( q1 + 3 * size_t( c.b0 ) ), __ldg( q1 + 3 * size_t( c.b0 ) + 1 ), __ldg( q1 + 3 * size_t( c.b0 ) + 2 ) };
  const SgVec3 xb{ __ldg( q1 + 3 * size_t( c.b1 ) ), __ldg( q1 + 3 * size_t( c.b1 ) + 1 ), __ldg( q1 + 3 * size_t( c.b1 ) + 2 ) };
  // both copies teleported: collisionIsActive( b0, b1, q0, q1 ) -- kinematic-kinematic never is, spheres by SphereSphereConstraint::isActive at q1
  if( first_teleported && second_teleported && !( f0 && f1 ) && sg_sphere_sphere_active( xa, xb, r0, r1 ) ) { tel_cnt[k] = 0u; return; }
  if( f0 && f1 ) { tel_cnt[k] = 0u; return; }
  tel_cnt[k] = sg_sphere_sphere_active( sg_tele3_center( ps, c.p0, xa ), sg_tele3_center( ps, c.p1, xb ), r0, r1 ) ? 1u : 0u;
}

// generateTeleportedCollision (RigidBody3DSim.cpp:1338-1397): centres teleported at q0; a kinematically scripted sphere makes it a
// KinematicObjectSphereConstraint ( free sphere first, p = the kinematic sphere's teleported centre ), otherwise a
// TeleportedSphereSphereConstraint ( p = q0_i + r_i / ( r_i + r_j ) * ( x1 - x0 ) ); no penetration depth override ( NaN )
__global__ void __launch_bounds__( 128 ) k_r3p_tele_contacts( const __grid_constant__ SgPortals3D ps, const Rb3dDev dev, const uint32_t nraw, const uint32_t* __restrict__ idxs,
                                                             const uint32_t* __restrict__ uflag, const uint32_t* __restrict__ uoff, const uint4* __restrict__ tc_info,
                                                             const double* __restrict__ q0, const unsigned long long base, const ContactOut3D out,
                                                             double* __restrict__ x0t, double* __restrict__ x1t, uint32_t* __restrict__ tp0, uint32_t* __restrict__ tp1 )
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if( e >= nraw || uflag[e] == 0u ) { return; }
  const uint32_t s = uoff[e];
  const uint4 c = tc_info[idxs[e]];
  const uint32_t b0 = c.x, b1 = c.y, p0 = c.z, p1 = c.w;
  const bool f0 = ( __ldg( &dev.btype[b0] ) & SG_FIXED_BIT ) != 0u, f1 = ( __ldg( &dev.btype[b1] ) & SG_FIXED_BIT ) != 0u;
  const double r0 = __ldg( &dev.bparam[4 * size_t( b0 )] ), r1 = __ldg( &dev.bparam[4 * size_t( b1 )] );
  const SgVec3 q0a{ __ldg( q0 + 3 * size_t( b0 ) ), __ldg( q0 + 3 * size_t( b0 ) + 1 ), __ldg( q0 + 3 * size_t( b0 ) + 2 ) };
  const SgVec3 q0b{ __ldg( q0 + 3 * size_t( b1 ) ), __ldg( q0 + 3 * size_t( b1 ) + 1 ), __ldg( q0 + 3 * size_t( b1 ) + 2 ) };
  const SgVec3 x0 = sg_tele3_center( ps, p0, q0a ), x1 = sg_tele3_center( ps, p1, q0b );
  uint32_t type, ci, cj;
  double nx, ny, nz, px, py, pz;
  if( f0 && !f1 ) { type = SG_KINEMATIC_OBJECT_SPHERE_TELEPORTED; ci = b1; cj = b0; nx = x1.x - x0.x; ny = x1.y - x0.y; nz = x1.z - x0.z; px = x0.x; py = x0.y; pz = x0.z; }
  else if( !f0 && f1 ) { type = SG_KINEMATIC_OBJECT_SPHERE_TELEPORTED; ci = b0; cj = b1; nx = x0.x - x1.x; ny = x0.y - x1.y; nz = x0.z - x1.z; px = x1.x; py = x1.y; pz = x1.z; }
  else
  {
    type = SG_SPHERE_SPHERE_TELEPORTED; ci = b0; cj = b1; nx = x0.x - x1.x; ny = x0.y - x1.y; nz = x0.z - x1.z;
    const double w = r0 / ( r0 + r1 );
    px = q0a.x + w * ( x1.x - x0.x ); py = q0a.y + w * ( x1.y - x0.y ); pz = q0a.z + w * ( x1.z - x0.z );
  }
  const double z = ( nx * nx + ny * ny ) + nz * nz;
  if( z > 0.0 ) { const double sq = sqrt( z ); nx = nx / sq; ny = ny / sq; nz = nz / sq; }
  const unsigned long long o = base + s;
  if( o < out.cap )
  {
    out.type[o] = type; out.i[o] = ci; out.j[o] = cj; out.aux[o] = 0u;
    out.n[3 * o] = nx; out.n[3 * o + 1] = ny; out.n[3 * o + 2] = nz;
    out.p[3 * o] = px; out.p[3 * o + 1] = py; out.p[3 * o + 2] = pz;
    out.depth[o] = __longlong_as_double( 0x7ff8000000000000LL );
  }
  x0t[3 * size_t( s )] = x0.x; x0t[3 * size_t( s ) + 1] = x0.y; x0t[3 * size_t( s ) + 2] = x0.z;
  x1t[3 * size_t( s )] = x1.x; x1t[3 * size_t( s ) + 1] = x1.y; x1t[3 * size_t( s ) + 2] = x1.z;
  tp0[s] = p0; tp1[s] = p1;
}

// RigidBody3DSim::enforcePeriodicBoundaryConditions: centres of mass ( the first 3n entries of q ) in place
__global__ void __launch_bounds__( 256 ) k_r3p_enforce( const __grid_constant__ SgPortals3D ps, const uint32_t n, double* __restrict__ q )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const SgVec3 x = sg_portals3_enforce( ps, SgVec3{ q[3 * size_t( b )], q[3 * size_t( b ) + 1], q[3 * size_t( b ) + 2] } );
  q[3 * size_t( b )] = x.x; q[3 * size_t( b ) + 1] = x.y; q[3 * size_t( b ) + 2] = x.z;
}

#endif
