// sg_boxbox.cuh -- 3-D box-box narrow phase, one thread per candidate pair.
//
// Device restatement, branch for branch, of rigidbody3d/Constraints/BoxBoxUtilities.cpp (ODE-derived):
//   :18-37   closest approach of two edges          sgbb_line_closest
//   :47-108  rectangle / quadrilateral clipping     sgbb_clip_rect_quad (the two ping-pong buffers are kept)
//   :110-150 face and edge separating-axis tests, edge axes biased by the 1.05 fudge factor
//   :167-615 boxBox: 15 axes at q1, then either one edge-edge contact or up to 8 clipped face contacts
//   :617-627 isActive: the normal is flipped to point from the second body to the first
// Tie-breaking (">" vs ">=") and evaluation order are the reference's; everything stays in registers / local memory.
#ifndef SG_BOXBOX_CUH
#define SG_BOXBOX_CUH

#include "sg_math3.cuh"

__device__ inline void sgbb_line_closest( const V3d pa, const V3d ua, const V3d pb, const V3d ub, double& alpha, double& beta )
{
  const V3d p = pb - pa;
  const double uaub = dot3( ua, ub );
  const double q1 = dot3( ua, p );
  const double q2 = -dot3( ub, p );
  double d = 1.0 - uaub * uaub;
  if( d <= 0.0001 ) { alpha = 0.0; beta = 0.0; }
  else
  {
    d = 1.0 / d;
    alpha = ( q1 + uaub * q2 ) * d;
    beta = ( uaub * q1 + q2 ) * d;
  }
}

// Clips the quadrilateral p[0..7] against the rectangle |x| < h[0], |y| < h[1]; result (<= 8 points) in ret.
__device__ inline int sgbb_clip_rect_quad( const double* h, double* p, double* ret )
{
  int nq = 4, nr = 0;
  double buffer[16];
  double* q = p;
  double* r = ret;
  bool full = false;
  for( int dir = 0; dir <= 1 && !full; ++dir )
  {
    for( int sign = -1; sign <= 1 && !full; sign += 2 )
    {
      double* pq = q;
      double* pr = r;
      nr = 0;
      for( int i = nq; i > 0; --i )
      {
        const bool in_cur = double( sign ) * pq[dir] < h[dir];
        if( in_cur )
        {
          pr[0] = pq[0]; pr[1] = pq[1];
          pr += 2; nr++;
          if( nr & 8 ) { q = r; full = true; break; }
        }
        double* nextq = ( i > 1 ) ? pq + 2 : q;
        const bool in_next = double( sign ) * nextq[dir] < h[dir];
        if( in_cur ^ in_next )
        {
          pr[1 - dir] = pq[1 - dir] + ( nextq[1 - dir] - pq[1 - dir] ) / ( nextq[dir] - pq[dir] ) * ( double( sign ) * h[dir] - pq[dir] );
          pr[dir] = double( sign ) * h[dir];
          pr += 2; nr++;
          if( nr & 8 ) { q = r; full = true; break; }
        }
        pq += 2;
      }
      if( full ) { break; }
      q = r;
      r = ( q == ret ) ? buffer : ret;
      nq = nr;
    }
  }
  if( q != ret ) { for( int k = 0; k < 2 * nr; ++k ) { ret[k] = q[k]; } }
  return nr;
}

struct SgbbState
{
  double depth;
  bool invert;
  int code;
  V3d normalC;
};

__device__ __forceinline__ bool sgbb_face_axis( const double dist, const double widths, const int crnt, SgbbState& s )
{
  const double pen = fabs( dist ) - widths;
  if( pen > 0 ) { return true; }
  if( pen > s.depth ) { s.depth = pen; s.invert = dist < 0.0; s.code = crnt; }
  return false;
}

__device__ __forceinline__ bool sgbb_edge_axis( const double dist, const double widths, const double n1, const double n2, const double n3, const int crnt, SgbbState& s )
{
  double s2 = fabs( dist ) - widths;
  if( s2 > 0 ) { return true; }
  const double l = sqrt( n1 * n1 + n2 * n2 + n3 * n3 );
  if( l > 0 )
  {
    s2 /= l;
    if( s2 * 1.05 > s.depth )
    {
      s.depth = s2;
      s.normalC = v3( n1 / l, n2 / l, n3 / l );
      s.invert = dist < 0;
      s.code = crnt;
    }
  }
  return false;
}

// Returns the number of contacts (0..8); points[3*k..] are world-space contact points, n the constraint normal
// (already flipped as BoxBoxUtilities::isActive does).
__device__ inline int sg_box_box( const V3d p1, const M3d& R1, const V3d side1, const V3d p2, const M3d& R2, const V3d side2, V3d& n_out, double* points )
{
  const V3d p = p2 - p1;
  const V3d pp = mulT3( R1, p );
  const M3d R = mulTN33( R1, R2 );
  M3d Q;
  #pragma unroll
  for( int k = 0; k < 9; ++k ) { Q.m[k] = fabs( R.m[k] ); }
  #define SGBB_R( r, c ) R.m[3 * ( r ) + ( c )]
  #define SGBB_Q( r, c ) Q.m[3 * ( r ) + ( c )]
  const double s1[3] = { side1.x, side1.y, side1.z };
  const double s2[3] = { side2.x, side2.y, side2.z };
  SgbbState st;
  st.depth = -__longlong_as_double( 0x7ff0000000000000LL );
  st.invert = false;
  st.code = 0;
  st.normalC = v3( 0.0, 0.0, 0.0 );
  {
    const V3d QbA = mul3( Q, side2 ) + side1;
    if( sgbb_face_axis( pp.x, QbA.x, 1, st ) ) { return 0; }
    if( sgbb_face_axis( pp.y, QbA.y, 2, st ) ) { return 0; }
    if( sgbb_face_axis( pp.z, QbA.z, 3, st ) ) { return 0; }
  }
  {
    const V3d pR2 = mulT3( R2, p );
    const V3d QTaB = mulT3( Q, side1 ) + side2;
    if( sgbb_face_axis( pR2.x, QTaB.x, 4, st ) ) { return 0; }
    if( sgbb_face_axis( pR2.y, QTaB.y, 5, st ) ) { return 0; }
    if( sgbb_face_axis( pR2.z, QTaB.z, 6, st ) ) { return 0; }
  }
  if( sgbb_edge_axis( pp.z * SGBB_R( 1, 0 ) - pp.y * SGBB_R( 2, 0 ), s1[1] * SGBB_Q( 2, 0 ) + s1[2] * SGBB_Q( 1, 0 ) + s2[1] * SGBB_Q( 0, 2 ) + s2[2] * SGBB_Q( 0, 1 ), 0, -SGBB_R( 2, 0 ), SGBB_R( 1, 0 ), 7, st ) ) { return 0; }
  if( sgbb_edge_axis( pp.z * SGBB_R( 1, 1 ) - pp.y * SGBB_R( 2, 1 ), s1[1] * SGBB_Q( 2, 1 ) + s1[2] * SGBB_Q( 1, 1 ) + s2[0] * SGBB_Q( 0, 2 ) + s2[2] * SGBB_Q( 0, 0 ), 0, -SGBB_R( 2, 1 ), SGBB_R( 1, 1 ), 8, st ) ) { return 0; }
  if( sgbb_edge_axis( pp.z * SGBB_R( 1, 2 ) - pp.y * SGBB_R( 2, 2 ), s1[1] * SGBB_Q( 2, 2 ) + s1[2] * SGBB_Q( 1, 2 ) + s2[0] * SGBB_Q( 0, 1 ) + s2[1] * SGBB_Q( 0, 0 ), 0, -SGBB_R( 2, 2 ), SGBB_R( 1, 2 ), 9, st ) ) { return 0; }
  if( sgbb_edge_axis( pp.x * SGBB_R( 2, 0 ) - pp.z * SGBB_R( 0, 0 ), s1[0] * SGBB_Q( 2, 0 ) + s1[2] * SGBB_Q( 0, 0 ) + s2[1] * SGBB_Q( 1, 2 ) + s2[2] * SGBB_Q( 1, 1 ), SGBB_R( 2, 0 ), 0, -SGBB_R( 0, 0 ), 10, st ) ) { return 0; }
  if( sgbb_edge_axis( pp.x * SGBB_R( 2, 1 ) - pp.z * SGBB_R( 0, 1 ), s1[0] * SGBB_Q( 2, 1 ) + s1[2] * SGBB_Q( 0, 1 ) + s2[0] * SGBB_Q( 1, 2 ) + s2[2] * SGBB_Q( 1, 0 ), SGBB_R( 2, 1 ), 0, -SGBB_R( 0, 1 ), 11, st ) ) { return 0; }
  if( sgbb_edge_axis( pp.x * SGBB_R( 2, 2 ) - pp.z * SGBB_R( 0, 2 ), s1[0] * SGBB_Q( 2, 2 ) + s1[2] * SGBB_Q( 0, 2 ) + s2[0] * SGBB_Q( 1, 1 ) + s2[1] * SGBB_Q( 1, 0 ), SGBB_R( 2, 2 ), 0, -SGBB_R( 0, 2 ), 12, st ) ) { return 0; }
  if( sgbb_edge_axis( pp.y * SGBB_R( 0, 0 ) - pp.x * SGBB_R( 1, 0 ), s1[0] * SGBB_Q( 1, 0 ) + s1[1] * SGBB_Q( 0, 0 ) + s2[1] * SGBB_Q( 2, 2 ) + s2[2] * SGBB_Q( 2, 1 ), -SGBB_R( 1, 0 ), SGBB_R( 0, 0 ), 0, 13, st ) ) { return 0; }
  if( sgbb_edge_axis( pp.y * SGBB_R( 0, 1 ) - pp.x * SGBB_R( 1, 1 ), s1[0] * SGBB_Q( 1, 1 ) + s1[1] * SGBB_Q( 0, 1 ) + s2[0] * SGBB_Q( 2, 2 ) + s2[2] * SGBB_Q( 2, 0 ), -SGBB_R( 1, 1 ), SGBB_R( 0, 1 ), 0, 14, st ) ) { return 0; }
  if( sgbb_edge_axis( pp.y * SGBB_R( 0, 2 ) - pp.x * SGBB_R( 1, 2 ), s1[0] * SGBB_Q( 1, 2 ) + s1[1] * SGBB_Q( 0, 2 ) + s2[0] * SGBB_Q( 2, 1 ) + s2[1] * SGBB_Q( 2, 0 ), -SGBB_R( 1, 2 ), SGBB_R( 0, 2 ), 0, 15, st ) ) { return 0; }
  #undef SGBB_R
  #undef SGBB_Q

  const int code = st.code;
  V3d normal;
  if( code <= 6 ) { normal = ( code <= 3 ) ? col3( R1, code - 1 ) : col3( R2, code - 4 ); }
  else { normal = mul3( R1, st.normalC ); }
  if( st.invert ) { normal = v3( normal.x * -1.0, normal.y * -1.0, normal.z * -1.0 ); }
  // BoxBoxUtilities::isActive: n *= -1
  n_out = v3( normal.x * -1.0, normal.y * -1.0, normal.z * -1.0 );

  if( code > 6 )
  {
    double pa[3] = { p1.x, p1.y, p1.z };
    for( int j = 0; j < 3; ++j )
    {
      const double sign = dot3( normal, col3( R1, j ) ) > 0 ? 1.0 : -1.0;
      for( int i = 0; i < 3; ++i ) { pa[i] += sign * s1[j] * R1.m[3 * i + j]; }
    }
    double pb[3] = { p2.x, p2.y, p2.z };
    for( int j = 0; j < 3; ++j )
    {
      const double sign = dot3( normal, col3( R2, j ) ) > 0 ? -1.0 : 1.0;
      for( int i = 0; i < 3; ++i ) { pb[i] += sign * s2[j] * R2.m[3 * i + j]; }
    }
    const V3d ua = col3( R1, ( code - 7 ) / 3 );
    const V3d ub = col3( R2, ( code - 7 ) % 3 );
    double alpha, beta;
    sgbb_line_closest( v3( pa[0], pa[1], pa[2] ), ua, v3( pb[0], pb[1], pb[2] ), ub, alpha, beta );
    const V3d qa = v3( pa[0], pa[1], pa[2] ) + alpha * ua;
    const V3d qb = v3( pb[0], pb[1], pb[2] ) + beta * ub;
    const V3d c = 0.5 * ( qa + qb );
    points[0] = c.x; points[1] = c.y; points[2] = c.z;
    return 1;
  }

  const bool first = code <= 3;
  const M3d& Ra = first ? R1 : R2;
  const M3d& Rb = first ? R2 : R1;
  const V3d pa = first ? p1 : p2;
  const V3d pb = first ? p2 : p1;
  const double* Sa = first ? s1 : s2;
  const double* Sb = first ? s2 : s1;
  const V3d normal2 = first ? normal : -normal;
  const V3d nr = mulT3( Rb, normal2 );
  int lanr, a1, a2;
  {
    const double anr0 = fabs( nr.x ), anr1 = fabs( nr.y ), anr2 = fabs( nr.z );
    if( anr1 > anr0 )
    {
      if( anr1 > anr2 ) { a1 = 0; lanr = 1; a2 = 2; }
      else { a1 = 0; a2 = 1; lanr = 2; }
    }
    else
    {
      if( anr0 > anr2 ) { lanr = 0; a1 = 1; a2 = 2; }
      else { a1 = 0; a2 = 1; lanr = 2; }
    }
  }
  V3d center;
  {
    const V3d d = pb - pa;
    const V3d sc = Sb[lanr] * col3( Rb, lanr );
    center = ( at3( nr, lanr ) < 0 ) ? d + sc : d - sc;
  }
  const int codeN = first ? code - 1 : code - 4;
  int code1, code2;
  if( codeN == 0 ) { code1 = 1; code2 = 2; }
  else if( codeN == 1 ) { code1 = 0; code2 = 2; }
  else { code1 = 0; code2 = 1; }
  const double c1 = dot3( center, col3( Ra, code1 ) );
  const double c2 = dot3( center, col3( Ra, code2 ) );
  double m11 = dot3( col3( Ra, code1 ), col3( Rb, a1 ) );
  double m12 = dot3( col3( Ra, code1 ), col3( Rb, a2 ) );
  double m21 = dot3( col3( Ra, code2 ), col3( Rb, a1 ) );
  double m22 = dot3( col3( Ra, code2 ), col3( Rb, a2 ) );
  double quad[8];
  {
    const double k1 = m11 * Sb[a1];
    const double k2 = m21 * Sb[a1];
    const double k3 = m12 * Sb[a2];
    const double k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  const double rect[2] = { Sa[code1], Sa[code2] };
  double ret[16];
  const int n = sgbb_clip_rect_quad( rect, quad, ret );
  const double det1 = 1.0 / ( m11 * m22 - m12 * m21 );
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  const double cen[3] = { center.x, center.y, center.z };
  for( int j = 0; j < n; ++j )
  {
    const double k1 = m22 * ( ret[j * 2] - c1 ) - m12 * ( ret[j * 2 + 1] - c2 );
    const double k2 = -m21 * ( ret[j * 2] - c1 ) + m11 * ( ret[j * 2 + 1] - c2 );
    double pt[3];
    for( int i = 0; i < 3; ++i ) { pt[i] = cen[i] + k1 * Rb.m[3 * i + a1] + k2 * Rb.m[3 * i + a2]; }
    const double dep = Sa[codeN] - ( normal2.x * pt[0] + normal2.y * pt[1] + normal2.z * pt[2] );
    if( dep >= 0 )
    {
      // world-space point = point + pa (BoxBoxUtilities.cpp:600-606)
      points[3 * cnum] = pt[0] + pa.x; points[3 * cnum + 1] = pt[1] + pa.y; points[3 * cnum + 2] = pt[2] + pa.z;
      cnum++;
    }
  }
  return cnum;
}

#endif
