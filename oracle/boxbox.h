// oracle/boxbox.h
// Checked bit for bit against the reference's own source compiled unchanged (oracle/Makefile.ref, tests/test_oracle_vs_reference.py).
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_math.h header). CPU restatement of the reference's ODE-derived
// 3-D box-box test:
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp:11-14     dot (explicit a0*b0 + a1*b1 + a2*b2)
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp:18-37     lineClosestApproach
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp:47-108    intersectRectQuad
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp:110-127   axisAlignedSeperatingTest
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp:129-150   edgeEdgeSeparatingTest (1.05 fudge factor)
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp:167-615   boxBox
//   rigidbody3d/Constraints/BoxBoxUtilities.cpp:617-627   isActive (normal flipped to point from body 1 to body 0)
// PARITY UNPINNED: the reference stores no expected outputs for this routine; Eigen 3-term products are
// taken as (a0*b0 + a1*b1) + a2*b2 throughout (oracle_math.h).
#ifndef ORACLE_BOXBOX_H
#define ORACLE_BOXBOX_H

#include "oracle_math.h"

#include <cstring>
#include <limits>
#include <vector>

namespace orc
{
namespace boxbox
{

inline double dot3( const double* a, const double* b ) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

inline V3 col( const M3& R, const int j ) { return V3{ R.m[j], R.m[3 + j], R.m[6 + j] }; }
inline double& at( V3& v, const int i ) { return i == 0 ? v.x : ( i == 1 ? v.y : v.z ); }
inline double at( const V3& v, const int i ) { return i == 0 ? v.x : ( i == 1 ? v.y : v.z ); }

inline void lineClosestApproach( const V3& pa, const V3& ua, const V3& pb, const V3& ub, double& alpha, double& beta )
{
  const V3 p = pb - pa;
  const double uaub = dot( ua, ub );
  const double q1 = dot( ua, p );
  const double q2 = -dot( ub, p );
  double d = 1.0 - uaub * uaub;
  if( d <= 0.0001 )
  {
    alpha = 0.0;
    beta = 0.0;
  }
  else
  {
    d = 1.0 / d;
    alpha = ( q1 + uaub * q2 ) * d;
    beta = ( uaub * q1 + q2 ) * d;
  }
}

inline int intersectRectQuad( double h[2], double p[8], double ret[16] )
{
  int nq = 4;
  int nr = 0;
  double buffer[16];
  double* q = p;
  double* r = ret;
  for( int dir = 0; dir <= 1; ++dir )
  {
    for( int sign = -1; sign <= 1; sign += 2 )
    {
      double* pq = q;
      double* pr = r;
      nr = 0;
      for( int i = nq; i > 0; --i )
      {
        if( sign * pq[dir] < h[dir] )
        {
          pr[0] = pq[0];
          pr[1] = pq[1];
          pr += 2;
          nr++;
          if( nr & 8 ) { q = r; goto done; }
        }
        double* nextq = ( i > 1 ) ? pq + 2 : q;
        if( ( sign * pq[dir] < h[dir] ) ^ ( sign * nextq[dir] < h[dir] ) )
        {
          pr[1 - dir] = pq[1 - dir] + ( nextq[1 - dir] - pq[1 - dir] ) / ( nextq[dir] - pq[dir] ) * ( sign * h[dir] - pq[dir] );
          pr[dir] = sign * h[dir];
          pr += 2;
          nr++;
          if( nr & 8 ) { q = r; goto done; }
        }
        pq += 2;
      }
      q = r;
      r = ( q == ret ) ? buffer : ret;
      nq = nr;
    }
  }
done:
  if( q != ret ) { std::memcpy( ret, q, nr * 2 * sizeof( double ) ); }
  return nr;
}

inline bool axisAlignedSeperatingTest( const double projected_center_dist, const double projected_aabb_widths, const int crnt_code, double& smallest_pen_depth, bool& invert_normal, int& code )
{
  const double pen_depth = std::fabs( projected_center_dist ) - projected_aabb_widths;
  if( pen_depth > 0 ) { return true; }
  if( pen_depth > smallest_pen_depth )
  {
    smallest_pen_depth = pen_depth;
    invert_normal = projected_center_dist < 0.0;
    code = crnt_code;
  }
  return false;
}

inline bool edgeEdgeSeparatingTest( const double projected_center_dist, const double projected_aabb_widths, const double n1, const double n2, const double n3, const int crnt_code,
                                    double& smallest_pen_depth, V3& normalC, bool& invert_normal, int& code )
{
  double s2 = std::fabs( projected_center_dist ) - projected_aabb_widths;
  if( s2 > 0 ) { return true; }
  const double l = std::sqrt( n1 * n1 + n2 * n2 + n3 * n3 );
  if( l > 0 )
  {
    s2 /= l;
    const double fudge_factor = 1.05;
    if( s2 * fudge_factor > smallest_pen_depth )
    {
      smallest_pen_depth = s2;
      normalC = V3{ n1 / l, n2 / l, n3 / l };
      invert_normal = projected_center_dist < 0;
      code = crnt_code;
    }
  }
  return false;
}

inline void boxBox( const V3& p1, const M3& R1, const V3& side1, const V3& p2, const M3& R2, const V3& side2, V3& normal, double& depth, int& code,
                    std::vector<V3>& contact_points, std::vector<double>& depths )
{
  const V3 p = p2 - p1;
  const V3 pp = mulT( R1, p );
  const M3 R = mulTN( R1, R2 );
  M3 Q;
  for( int k = 0; k < 9; ++k ) { Q.m[k] = std::fabs( R.m[k] ); }
  auto Rr = [&]( int r, int c ) { return R.m[3 * r + c]; };
  auto Qq = [&]( int r, int c ) { return Q.m[3 * r + c]; };
  const double s1[3] = { side1.x, side1.y, side1.z };
  const double s2[3] = { side2.x, side2.y, side2.z };

  depth = -std::numeric_limits<double>::infinity();
  bool invert_normal = false;
  code = 0;

  {
    const V3 QbA = mul( Q, side2 ) + side1;
    if( axisAlignedSeperatingTest( pp.x, QbA.x, 1, depth, invert_normal, code ) ) { return; }
    if( axisAlignedSeperatingTest( pp.y, QbA.y, 2, depth, invert_normal, code ) ) { return; }
    if( axisAlignedSeperatingTest( pp.z, QbA.z, 3, depth, invert_normal, code ) ) { return; }
  }
  {
    const V3 p_on_R2 = mulT( R2, p );
    const V3 QTaB = mulT( Q, side1 ) + side2;
    if( axisAlignedSeperatingTest( p_on_R2.x, QTaB.x, 4, depth, invert_normal, code ) ) { return; }
    if( axisAlignedSeperatingTest( p_on_R2.y, QTaB.y, 5, depth, invert_normal, code ) ) { return; }
    if( axisAlignedSeperatingTest( p_on_R2.z, QTaB.z, 6, depth, invert_normal, code ) ) { return; }
  }

  V3 normalC{ 0.0, 0.0, 0.0 };
  // u1 x v1..v3
  if( edgeEdgeSeparatingTest( pp.z * Rr( 1, 0 ) - pp.y * Rr( 2, 0 ), s1[1] * Qq( 2, 0 ) + s1[2] * Qq( 1, 0 ) + s2[1] * Qq( 0, 2 ) + s2[2] * Qq( 0, 1 ), 0, -Rr( 2, 0 ), Rr( 1, 0 ), 7, depth, normalC, invert_normal, code ) ) { return; }
  if( edgeEdgeSeparatingTest( pp.z * Rr( 1, 1 ) - pp.y * Rr( 2, 1 ), s1[1] * Qq( 2, 1 ) + s1[2] * Qq( 1, 1 ) + s2[0] * Qq( 0, 2 ) + s2[2] * Qq( 0, 0 ), 0, -Rr( 2, 1 ), Rr( 1, 1 ), 8, depth, normalC, invert_normal, code ) ) { return; }
  if( edgeEdgeSeparatingTest( pp.z * Rr( 1, 2 ) - pp.y * Rr( 2, 2 ), s1[1] * Qq( 2, 2 ) + s1[2] * Qq( 1, 2 ) + s2[0] * Qq( 0, 1 ) + s2[1] * Qq( 0, 0 ), 0, -Rr( 2, 2 ), Rr( 1, 2 ), 9, depth, normalC, invert_normal, code ) ) { return; }
  // u2 x v1..v3
  if( edgeEdgeSeparatingTest( pp.x * Rr( 2, 0 ) - pp.z * Rr( 0, 0 ), s1[0] * Qq( 2, 0 ) + s1[2] * Qq( 0, 0 ) + s2[1] * Qq( 1, 2 ) + s2[2] * Qq( 1, 1 ), Rr( 2, 0 ), 0, -Rr( 0, 0 ), 10, depth, normalC, invert_normal, code ) ) { return; }
  if( edgeEdgeSeparatingTest( pp.x * Rr( 2, 1 ) - pp.z * Rr( 0, 1 ), s1[0] * Qq( 2, 1 ) + s1[2] * Qq( 0, 1 ) + s2[0] * Qq( 1, 2 ) + s2[2] * Qq( 1, 0 ), Rr( 2, 1 ), 0, -Rr( 0, 1 ), 11, depth, normalC, invert_normal, code ) ) { return; }
  if( edgeEdgeSeparatingTest( pp.x * Rr( 2, 2 ) - pp.z * Rr( 0, 2 ), s1[0] * Qq( 2, 2 ) + s1[2] * Qq( 0, 2 ) + s2[0] * Qq( 1, 1 ) + s2[1] * Qq( 1, 0 ), Rr( 2, 2 ), 0, -Rr( 0, 2 ), 12, depth, normalC, invert_normal, code ) ) { return; }
  // u3 x v1..v3
  if( edgeEdgeSeparatingTest( pp.y * Rr( 0, 0 ) - pp.x * Rr( 1, 0 ), s1[0] * Qq( 1, 0 ) + s1[1] * Qq( 0, 0 ) + s2[1] * Qq( 2, 2 ) + s2[2] * Qq( 2, 1 ), -Rr( 1, 0 ), Rr( 0, 0 ), 0, 13, depth, normalC, invert_normal, code ) ) { return; }
  if( edgeEdgeSeparatingTest( pp.y * Rr( 0, 1 ) - pp.x * Rr( 1, 1 ), s1[0] * Qq( 1, 1 ) + s1[1] * Qq( 0, 1 ) + s2[0] * Qq( 2, 2 ) + s2[2] * Qq( 2, 0 ), -Rr( 1, 1 ), Rr( 0, 1 ), 0, 14, depth, normalC, invert_normal, code ) ) { return; }
  if( edgeEdgeSeparatingTest( pp.y * Rr( 0, 2 ) - pp.x * Rr( 1, 2 ), s1[0] * Qq( 1, 2 ) + s1[1] * Qq( 0, 2 ) + s2[0] * Qq( 2, 1 ) + s2[1] * Qq( 2, 0 ), -Rr( 1, 2 ), Rr( 0, 2 ), 0, 15, depth, normalC, invert_normal, code ) ) { return; }

  if( code <= 6 ) { normal = code <= 3 ? col( R1, code - 1 ) : col( R2, code - 4 ); }
  else { normal = mul( R1, normalC ); }
  if( invert_normal ) { normal = V3{ normal.x * -1.0, normal.y * -1.0, normal.z * -1.0 }; }
  depth *= -1.0;

  if( code > 6 )
  {
    V3 pa = p1;
    for( int j = 0; j < 3; ++j )
    {
      const double sign = dot( normal, col( R1, j ) ) > 0 ? 1.0 : -1.0;
      for( int i = 0; i < 3; ++i ) { at( pa, i ) += sign * s1[j] * R1.m[3 * i + j]; }
    }
    V3 pb = p2;
    for( int j = 0; j < 3; ++j )
    {
      const double sign = dot( normal, col( R2, j ) ) > 0 ? -1.0 : 1.0;
      for( int i = 0; i < 3; ++i ) { at( pb, i ) += sign * s2[j] * R2.m[3 * i + j]; }
    }
    const V3 ua = col( R1, ( code - 7 ) / 3 );
    const V3 ub = col( R2, ( code - 7 ) % 3 );
    double alpha, beta;
    lineClosestApproach( pa, ua, pb, ub, alpha, beta );
    pa = pa + alpha * ua;
    pb = pb + beta * ub;
    contact_points.emplace_back( 0.5 * ( pa + pb ) );
    depths.emplace_back( depth );
    return;
  }

  M3 Ra, Rb;
  V3 pa, pb, Sa, Sb;
  if( code <= 3 ) { Ra = R1; Rb = R2; pa = p1; pb = p2; Sa = side1; Sb = side2; }
  else { Ra = R2; Rb = R1; pa = p2; pb = p1; Sa = side2; Sb = side1; }

  const V3 normal2 = code <= 3 ? normal : -normal;
  const V3 nr = mulT( Rb, normal2 );
  int lanr, a1, a2;
  {
    const double anr[3] = { std::fabs( nr.x ), std::fabs( nr.y ), std::fabs( nr.z ) };
    if( anr[1] > anr[0] )
    {
      if( anr[1] > anr[2] ) { a1 = 0; lanr = 1; a2 = 2; }
      else { a1 = 0; a2 = 1; lanr = 2; }
    }
    else
    {
      if( anr[0] > anr[2] ) { lanr = 0; a1 = 1; a2 = 2; }
      else { a1 = 0; a2 = 1; lanr = 2; }
    }
  }
  V3 center;
  {
    const V3 d = pb - pa;
    const V3 sc = at( Sb, lanr ) * col( Rb, lanr );
    center = at( nr, lanr ) < 0 ? d + sc : d - sc;
  }
  const int codeN = code <= 3 ? code - 1 : code - 4;
  int code1, code2;
  if( codeN == 0 ) { code1 = 1; code2 = 2; }
  else if( codeN == 1 ) { code1 = 0; code2 = 2; }
  else { code1 = 0; code2 = 1; }

  const double c1 = dot( center, col( Ra, code1 ) );
  const double c2 = dot( center, col( Ra, code2 ) );
  double m11 = dot( col( Ra, code1 ), col( Rb, a1 ) );
  double m12 = dot( col( Ra, code1 ), col( Rb, a2 ) );
  double m21 = dot( col( Ra, code2 ), col( Rb, a1 ) );
  double m22 = dot( col( Ra, code2 ), col( Rb, a2 ) );

  double quad[8];
  {
    const double k1 = m11 * at( Sb, a1 );
    const double k2 = m21 * at( Sb, a1 );
    const double k3 = m12 * at( Sb, a2 );
    const double k4 = m22 * at( Sb, a2 );
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  double rect[2] = { at( Sa, code1 ), at( Sa, code2 ) };
  double ret[16];
  const int n = intersectRectQuad( rect, quad, ret );

  double point[3 * 8];
  double dep[8];
  const double det1 = 1.0 / ( m11 * m22 - m12 * m21 );
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  const double n2a[3] = { normal2.x, normal2.y, normal2.z };
  for( int j = 0; j < n; ++j )
  {
    const double k1 = m22 * ( ret[j * 2] - c1 ) - m12 * ( ret[j * 2 + 1] - c2 );
    const double k2 = -m21 * ( ret[j * 2] - c1 ) + m11 * ( ret[j * 2 + 1] - c2 );
    for( int i = 0; i < 3; ++i ) { point[3 * cnum + i] = at( center, i ) + k1 * Rb.m[3 * i + a1] + k2 * Rb.m[3 * i + a2]; }
    dep[cnum] = at( Sa, codeN ) - dot3( n2a, point + 3 * cnum );
    if( dep[cnum] >= 0 )
    {
      ret[2 * cnum] = ret[2 * j];
      ret[2 * cnum + 1] = ret[2 * j + 1];
      cnum++;
    }
  }
  for( int j = 0; j < cnum; ++j )
  {
    contact_points.emplace_back( V3{ point[3 * j] + pa.x, point[3 * j + 1] + pa.y, point[3 * j + 2] + pa.z } );
    depths.emplace_back( dep[j] );
  }
}

inline void isActive( const V3& cm0, const M3& R0, const V3& side0, const V3& cm1, const M3& R1, const V3& side1, V3& n, std::vector<V3>& points )
{
  double max_depth;
  std::vector<double> depths;
  int return_code;
  n = V3{ 0.0, 0.0, 0.0 };
  boxBox( cm0, R0, side0, cm1, R1, side1, n, max_depth, return_code, points, depths );
  n = V3{ n.x * -1.0, n.y * -1.0, n.z * -1.0 };
}

}
}

#endif
