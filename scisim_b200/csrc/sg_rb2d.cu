// sg_rb2d.cu -- rigidbody2d hot path: SymplecticEuler / Verlet flow on [x,y,theta] bodies, 2-D broad phase,
// circle-circle CCD / box-box / circle-box narrow phase, body-plane tests.
//
// Reference behaviour reproduced (file:line in the SCISim checkout):
//   rigidbody2d/SymplecticEulerMap.cpp:15-38, VerletMap.cpp:15-55, NearEarthGravityForce.cpp:35-48   k_rb2d_flow
//   rigidbody2d/CircleGeometry.cpp:32-37 (swept AABB), BoxGeometry.cpp:32-42 (|Rot(theta1)| r at q1)  k_rb2d_aabb
//   rigidbody2d/SpatialGrid.cpp:114-141, RigidBody2DSim.cpp:1066-1100                                 sg_broadphase.cuh
//   rigidbody2d/RigidBody2DSim.cpp:248-348 dispatch + :184-246 box-box / circle-box callers           k_rb2d_pairs
//   rigidbody2d/BoxBoxTools.cpp:50-196, CircleBoxTools.cpp:8-114                                      rb2d_box_box / rb2d_circle_box
//   scisim/CollisionDetection/CollisionDetectionUtilities.cpp:3-121                                   rb2d_ccd
//   rigidbody2d/RigidBody2DSim.cpp:638-694                                                            k_rb2d_plane_*
// Pairs the reference exits on (box-box with a kinematic body, kinematic circle vs box) raise SG_ERR_UNSUPPORTED.
#include "sg_broadphase.cuh"

#define SG_FIXED_BIT2 0x80000000u
#define SG_GEO2_CIRCLE 0u
#define SG_GEO2_BOX 1u

struct V2d { double x, y; };
struct M2d { double a, b, c, d; }; // [[a,b],[c,d]]
__device__ __forceinline__ V2d v2( const double x, const double y ) { V2d r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ V2d operator-( const V2d a, const V2d b ) { return v2( a.x - b.x, a.y - b.y ); }
__device__ __forceinline__ V2d operator+( const V2d a, const V2d b ) { return v2( a.x + b.x, a.y + b.y ); }
__device__ __forceinline__ V2d operator*( const double s, const V2d a ) { return v2( s * a.x, s * a.y ); }
__device__ __forceinline__ double dot2( const V2d a, const V2d b ) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ double at2( const V2d v, const int i ) { return i == 0 ? v.x : v.y; }
__device__ __forceinline__ M2d rot2d( const double theta ) { double s, c; sincos( theta, &s, &c ); M2d R; R.a = c; R.b = -s; R.c = s; R.d = c; return R; }
__device__ __forceinline__ V2d mul2( const M2d& R, const V2d v ) { return v2( R.a * v.x + R.b * v.y, R.c * v.x + R.d * v.y ); }
__device__ __forceinline__ V2d mulT2( const M2d& R, const V2d v ) { return v2( R.a * v.x + R.c * v.y, R.b * v.x + R.d * v.y ); }
__device__ __forceinline__ V2d col2( const M2d& R, const int j ) { return j == 0 ? v2( R.a, R.c ) : v2( R.b, R.d ); }
__device__ __forceinline__ V2d normalized2( const V2d a )
{
  const double z = dot2( a, a );
  if( z > 0.0 ) { const double s = sqrt( z ); return v2( a.x / s, a.y / s ); }
  return a;
}

struct ContactOut2X
{
  uint32_t* type; uint32_t* i; uint32_t* j; uint32_t* aux;
  double2* n; double2* p; double* depth;
  unsigned long long cap;
};
__device__ __forceinline__ void put2( const ContactOut2X& out, const unsigned long long k, const uint32_t type, const uint32_t i, const uint32_t j, const uint32_t aux, const V2d n, const V2d p, const double depth )
{
  if( k >= out.cap ) { return; }
  out.type[k] = type; out.i[k] = i; out.j[k] = j; out.aux[k] = aux;
  out.n[k] = make_double2( n.x, n.y ); out.p[k] = make_double2( p.x, p.y ); out.depth[k] = depth;
}
__device__ __forceinline__ double nan2() { return __longlong_as_double( 0x7ff8000000000000LL ); }

// ---- broad phase policy on prebuilt boxes -----------------------------------------------------------
struct Box2DIn { const double* boxes; uint32_t n; };
struct alignas( 64 ) Box2DRec { double lo[2]; double hi[2]; uint32_t idx; uint32_t key; uint32_t c1, c2; double pad[2]; };
struct NoOut2D {};
struct Box2DPolicy
{
  static constexpr int D = 2;
  static constexpr bool HAS_NARROW = false;
  static constexpr double IN_BYTES = 32.0;
  static constexpr uint32_t IDX_OFFSET = 32u;
  static constexpr uint32_t ORD_OFFSET = IDX_OFFSET; // bodies are ranked by their index
  static constexpr bool ORD_IN_REC = true;
  static constexpr uint32_t IDX_MASK = 0xffffffffu;
  using In = Box2DIn;
  using Rec = Box2DRec;
  using Out = NoOut2D;
  __device__ static void load_aabb( const In& in, const uint32_t i, double* lo, double* hi )
  {
    const double* b = in.boxes + size_t( i ) * 4;
    lo[0] = __ldg( b ); lo[1] = __ldg( b + 1 ); hi[0] = __ldg( b + 2 ); hi[1] = __ldg( b + 3 );
  }
  __device__ static Rec make_rec( const In& in, const uint32_t i, const uint32_t key, const uint32_t c1, const uint32_t c2 )
  {
    Rec r;
    load_aabb( in, i, r.lo, r.hi );
    r.idx = i; r.key = key; r.c1 = c1; r.c2 = c2; r.pad[0] = 0.0; r.pad[1] = 0.0;
    return r;
  }
  __device__ static void rec_aabb( const Rec& s, double* lo, double* hi ) { lo[0] = s.lo[0]; lo[1] = s.lo[1]; hi[0] = s.hi[0]; hi[1] = s.hi[1]; }
  __device__ static uint32_t rec_idx( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_idx_raw( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_ord( const Rec& s ) { return rec_idx( s ); }
  __device__ static uint32_t rec_ord_raw( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_key( const Rec& s ) { return s.key; }
  __device__ static bool owns( const Rec& ) { return true; }
  __device__ static bool valid( const In&, const uint32_t ) { return true; }
  __device__ static uint32_t rec_c1( const Rec& s, const GridParams& ) { return s.c1; }
  __device__ static uint32_t rec_c2( const Rec& s, const GridParams& ) { return s.c2; }
  __device__ static Rec load_pass1( const Rec* __restrict__ p ) { return sg_load_rec_global<Rec>( p ); }
  __device__ static bool narrow_test( const Rec&, const Rec& ) { return false; }
  __device__ static void contact_emit( const Out&, unsigned long long&, const Rec&, const Rec& ) {}
};

struct Rb2dDev
{
  uint32_t n;
  const uint32_t* btype;  // geometry type | SG_FIXED_BIT2
  const double2* bparam;  // circle: (r, -), box: half widths
};

struct Planes2D
{
  uint32_t n;
  double x[SG_MAX_PLANES][2];
  double nrm[SG_MAX_PLANES][2];
};

// ---- flow ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 ) k_rb2d_flow( const int kind, const uint32_t n, const double* __restrict__ q0, const double* __restrict__ v0, const double* __restrict__ M, const uint32_t* __restrict__ btype,
                                                     const double gx, const double gy, const double dt, double* __restrict__ q1, double* __restrict__ v1 )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const bool fixed = ( __ldg( &btype[b] ) & SG_FIXED_BIT2 ) != 0u;
  const double m = __ldg( &M[3 * size_t( b )] );
  #pragma unroll
  for( int k = 0; k < 3; ++k )
  {
    const size_t d = 3 * size_t( b ) + k;
    const double minv = 1.0 / __ldg( &M[d] );
    double F = ( k < 2 ) ? 0.0 + m * ( k == 0 ? gx : gy ) : 0.0;
    if( fixed ) { F = 0.0; }
    const double v = __ldg( &v0[d] ), q = __ldg( &q0[d] );
    if( kind == SG_MAP_SYMPLECTIC_EULER )
    {
      const double vo = v + ( 0.0 + ( dt * minv ) * F );
      v1[d] = vo;
      q1[d] = q + dt * vo;
    }
    else
    {
      const double sc = ( 0.5 * dt ) * minv;
      const double vh = v + ( 0.0 + sc * F );
      q1[d] = q + dt * vh;
      v1[d] = vh + sc * F;
    }
  }
}

__global__ void __launch_bounds__( 256 ) k_rb2d_aabb( const Rb2dDev dev, const double* __restrict__ q0, const double* __restrict__ q1, double* __restrict__ boxes )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= dev.n ) { return; }
  const uint32_t type = __ldg( &dev.btype[b] ) & ~SG_FIXED_BIT2;
  const double2 par = __ldg( &dev.bparam[b] );
  const double* a = q0 + 3 * size_t( b );
  const double* c = q1 + 3 * size_t( b );
  double* o = boxes + 4 * size_t( b );
  if( type == SG_GEO2_CIRCLE )
  {
    o[0] = fmin( __ldg( a ), __ldg( c ) ) - par.x; o[1] = fmin( __ldg( a + 1 ), __ldg( c + 1 ) ) - par.x;
    o[2] = fmax( __ldg( a ), __ldg( c ) ) + par.x; o[3] = fmax( __ldg( a + 1 ), __ldg( c + 1 ) ) + par.x;
  }
  else
  {
    const M2d R = rot2d( __ldg( c + 2 ) );
    const double ex = fabs( R.a ) * par.x + fabs( R.b ) * par.y;
    const double ey = fabs( R.c ) * par.x + fabs( R.d ) * par.y;
    o[0] = __ldg( c ) - ex; o[1] = __ldg( c + 1 ) - ey; o[2] = __ldg( c ) + ex; o[3] = __ldg( c + 1 ) + ey;
  }
}

// ---- narrow phase ----------------------------------------------------------------------------------
__device__ __forceinline__ bool rb2d_ccd( const V2d q0a, const V2d q1a, const double ra, const V2d q0b, const V2d q1b, const double rb )
{
  return sg_ccd_ball_ball( q0a.x, q0a.y, q1a.x, q1a.y, ra, q0b.x, q0b.y, q1b.x, q1b.y, rb ); // sg_ccd.h
}

__device__ __forceinline__ bool rb2d_axis( const double dist, const double widths, const int crnt, double& smallest, bool& invert, int& feature )
{
  const double pen = fabs( dist ) - widths;
  if( pen > 0 ) { return true; }
  if( pen > smallest ) { smallest = pen; invert = dist < 0.0; feature = crnt; }
  return false;
}

// BoxBoxTools::isActive: returns the number of contacts (0..2)
__device__ inline int rb2d_box_box( const V2d x0, const double theta0, const V2d r0, const V2d x1, const double theta1, const V2d r1, V2d& n, V2d* pts )
{
  const M2d R0 = rot2d( theta0 ), R1 = rot2d( theta1 );
  int feature = 4;
  bool invert = false;
  {
    double min_pen = -__longlong_as_double( 0x7ff0000000000000LL );
    M2d Q;
    Q.a = fabs( R0.a * R1.a + R0.c * R1.c ); Q.b = fabs( R0.a * R1.b + R0.c * R1.d );
    Q.c = fabs( R0.b * R1.a + R0.d * R1.c ); Q.d = fabs( R0.b * R1.b + R0.d * R1.d );
    const V2d p = x1 - x0;
    {
      const V2d pR0 = mulT2( R0, p );
      const V2d w = mul2( Q, r1 ) + r0;
      if( rb2d_axis( pR0.x, w.x, 0, min_pen, invert, feature ) ) { return 0; }
      if( rb2d_axis( pR0.y, w.y, 1, min_pen, invert, feature ) ) { return 0; }
    }
    {
      const V2d pR1 = mulT2( R1, p );
      const V2d w = mulT2( Q, r0 ) + r1;
      if( rb2d_axis( pR1.x, w.x, 2, min_pen, invert, feature ) ) { return 0; }
      if( rb2d_axis( pR1.y, w.y, 3, min_pen, invert, feature ) ) { return 0; }
    }
  }
  const bool first = feature <= 1;
  V2d nn = first ? col2( R0, feature ) : col2( R1, feature - 2 );
  if( invert ) { nn = v2( nn.x * -1.0, nn.y * -1.0 ); }
  const M2d Ra = first ? R0 : R1, Rb = first ? R1 : R0;
  const V2d xa = first ? x0 : x1, xb = first ? x1 : x0;
  const V2d ra = first ? r0 : r1, rb = first ? r1 : r0;
  const V2d normal2 = first ? nn : v2( -nn.x, -nn.y );
  const V2d n_in_b = mulT2( Rb, normal2 );
  const int b_nrml = fabs( n_in_b.y ) > fabs( n_in_b.x ) ? 1 : 0;
  const int b_tngt = 1 - b_nrml;
  const double sc = ( at2( n_in_b, b_nrml ) < 0.0 ? 1.0 : -1.0 ) * at2( rb, b_nrml );
  const V2d bfc = ( xb - xa ) + sc * col2( Rb, b_nrml );
  const int a_nrml = first ? feature : feature - 2;
  const int a_tngt = 1 - a_nrml;
  const double c_on_a = dot2( bfc, col2( Ra, a_tngt ) );
  const double costheta = dot2( col2( Ra, a_tngt ), col2( Rb, b_tngt ) );
  double e0 = c_on_a - costheta * at2( rb, b_tngt ), e1 = c_on_a + costheta * at2( rb, b_tngt );
  if( e0 > e1 ) { const double t = e0; e0 = e1; e1 = t; }
  e0 = fmin( e0, at2( ra, a_tngt ) );
  e1 = fmax( e1, -at2( ra, a_tngt ) );
  const double isect[2] = { fmax( -at2( ra, a_tngt ), e0 ), fmin( at2( ra, a_tngt ), e1 ) };
  const int num = isect[0] != isect[1] ? 2 : 1;
  int cnt = 0;
  for( int c = 0; c < num; ++c )
  {
    const V2d point = bfc + ( ( isect[c] - c_on_a ) / costheta ) * col2( Rb, b_tngt );
    const double depth = at2( ra, a_nrml ) - dot2( normal2, point );
    if( depth >= 0.0 ) { pts[cnt++] = ( xa + point ) + ( 0.5 * depth ) * normal2; }
  }
  n = v2( nn.x * -1.0, nn.y * -1.0 );
  return cnt;
}

// CircleBoxTools::isActive
__device__ inline bool rb2d_circle_box( const V2d x0, const double r0, const V2d x1, const double theta1, const V2d r1, V2d& n, V2d& p )
{
  const M2d R = rot2d( theta1 );
  V2d xc = mulT2( R, x0 - x1 );
  const bool invert_x = xc.x < 0.0, invert_y = xc.y < 0.0;
  if( invert_x ) { xc.x *= -1.0; }
  if( invert_y ) { xc.y *= -1.0; }
  double pen;
  const bool right_or_corner = r1.x * xc.y < r1.y * xc.x;
  const bool flat = right_or_corner ? ( xc.y <= r1.y ) : ( xc.x <= r1.x );
  if( flat )
  {
    pen = right_or_corner ? ( xc.x - r0 - r1.x ) : ( xc.y - r0 - r1.y );
    if( pen > 0.0 ) { return false; }
    n = right_or_corner ? v2( 1.0, 0.0 ) : v2( 0.0, 1.0 );
  }
  else
  {
    n = xc - r1;
    pen = dot2( n, n );
    if( pen > r0 * r0 ) { return false; }
    pen = sqrt( pen );
    n = v2( n.x / pen, n.y / pen );
    pen -= r0;
  }
  p = xc - ( r0 + 0.5 * pen ) * n;
  if( invert_x ) { n.x *= -1.0; p.x *= -1.0; }
  if( invert_y ) { n.y *= -1.0; p.y *= -1.0; }
  n = mul2( R, n );
  p = mul2( R, p ) + x1;
  return true;
}

__device__ __forceinline__ V2d ldx( const double* __restrict__ q, const uint32_t b ) { return v2( __ldg( q + 3 * size_t( b ) ), __ldg( q + 3 * size_t( b ) + 1 ) ); }

template<bool EMIT>
__global__ void __launch_bounds__( 128 ) k_rb2d_pairs( const Rb2dDev dev, const uint2* __restrict__ pairs, const unsigned long long* __restrict__ npairs_dev, const double* __restrict__ q0, const double* __restrict__ q1,
                                                      uint32_t* __restrict__ counts, const unsigned long long* __restrict__ offsets, const ContactOut2X out, uint32_t* __restrict__ bad_flag )
{
  const unsigned long long k = blockIdx.x * ( unsigned long long )( blockDim.x ) + threadIdx.x;
  if( k >= *npairs_dev ) { return; }
  const uint2 pr = pairs[k];
  uint32_t i0 = pr.x, i1 = pr.y;
  uint32_t t0 = __ldg( &dev.btype[i0] ), t1 = __ldg( &dev.btype[i1] );
  uint32_t cnt = 0u;
  const unsigned long long o = EMIT ? offsets[k] : 0ull;
  if( !( ( t0 & SG_FIXED_BIT2 ) && ( t1 & SG_FIXED_BIT2 ) ) )
  {
    if( t0 & SG_FIXED_BIT2 ) { const uint32_t ti = i0; i0 = i1; i1 = ti; const uint32_t tt = t0; t0 = t1; t1 = tt; }
    const bool f1 = ( t1 & SG_FIXED_BIT2 ) != 0u;
    const uint32_t g0 = t0 & ~SG_FIXED_BIT2, g1 = t1 & ~SG_FIXED_BIT2;
    const double2 p0 = __ldg( &dev.bparam[i0] ), p1 = __ldg( &dev.bparam[i1] );
    if( g0 == SG_GEO2_CIRCLE && g1 == SG_GEO2_CIRCLE )
    {
      const V2d q0a = ldx( q0, i0 ), q1a = ldx( q1, i0 ), q0b = ldx( q0, i1 ), q1b = ldx( q1, i1 );
      if( rb2d_ccd( q0a, q1a, p0.x, q0b, q1b, p1.x ) )
      {
        cnt = 1u;
        if( EMIT )
        {
          const V2d n = normalized2( q0a - q0b );
          if( !f1 )
          {
            const V2d d1 = q1a - q1b;
            put2( out, o, SG_CIRCLE_CIRCLE, i0, i1, 0u, n, q0a + ( p0.x / ( p0.x + p1.x ) ) * ( q0b - q0a ), fmin( 0.0, sqrt( dot2( d1, d1 ) ) - p0.x - p1.x ) );
          }
          else { put2( out, o, SG_KINEMATIC_CIRCLE, i0, i1, 0u, n, q0b, nan2() ); }
        }
      }
    }
    else if( g0 == SG_GEO2_BOX && g1 == SG_GEO2_BOX )
    {
      if( ( t0 | t1 ) & SG_FIXED_BIT2 ) { if( !EMIT ) { atomicOr( bad_flag, 1u ); } }
      else
      {
        V2d n = v2( 0.0, 0.0 ), pts[2];
        const int nc = rb2d_box_box( ldx( q1, i0 ), __ldg( q1 + 3 * size_t( i0 ) + 2 ), v2( p0.x, p0.y ), ldx( q1, i1 ), __ldg( q1 + 3 * size_t( i1 ) + 2 ), v2( p1.x, p1.y ), n, pts );
        cnt = uint32_t( nc );
        if( EMIT ) { for( int c = 0; c < nc; ++c ) { put2( out, o + c, SG_BODY_BODY_2D, i0, i1, 0u, n, pts[c], nan2() ); } }
      }
    }
    else
    {
      const bool c_first = g0 == SG_GEO2_CIRCLE;
      const uint32_t ic = c_first ? i0 : i1, ib = c_first ? i1 : i0;
      const uint32_t tc = c_first ? t0 : t1, tb = c_first ? t1 : t0;
      const double2 pc = c_first ? p0 : p1, pb = c_first ? p1 : p0;
      if( tc & SG_FIXED_BIT2 ) { if( !EMIT ) { atomicOr( bad_flag, 1u ); } }
      else
      {
        V2d n, p;
        if( rb2d_circle_box( ldx( q1, ic ), pc.x, ldx( q1, ib ), __ldg( q1 + 3 * size_t( ib ) + 2 ), v2( pb.x, pb.y ), n, p ) )
        {
          cnt = 1u;
          if( EMIT )
          {
            if( !( tb & SG_FIXED_BIT2 ) )
            {
              if( ic < ib ) { put2( out, o, SG_BODY_BODY_2D, ic, ib, 0u, n, p, nan2() ); }
              else { put2( out, o, SG_BODY_BODY_2D, ib, ic, 0u, v2( -n.x, -n.y ), p, nan2() ); }
            }
            else { put2( out, o, SG_KINEMATIC_CIRCLE, ic, ib, 0u, n, ldx( q0, ib ), nan2() ); }
          }
        }
      }
    }
  }
  if( !EMIT ) { counts[k] = cnt; }
}

// ---- planes: plane-major, body ascending, corner order (-1,-1), (-1,1), (1,-1), (1,1) ---------------
__device__ inline uint32_t rb2d_plane_contacts( const Rb2dDev& dev, const Planes2D& planes, const uint32_t pl, const uint32_t b, const double* __restrict__ q0, const double* __restrict__ q1,
                                                const bool emit, const unsigned long long base, const ContactOut2X& out )
{
  const uint32_t t = __ldg( &dev.btype[b] );
  if( t & SG_FIXED_BIT2 ) { return 0u; }
  const V2d xp = v2( planes.x[pl][0], planes.x[pl][1] ), np = v2( planes.nrm[pl][0], planes.nrm[pl][1] );
  const V2d x1 = ldx( q1, b );
  const double2 par = __ldg( &dev.bparam[b] );
  uint32_t cnt = 0u;
  if( t == SG_GEO2_CIRCLE )
  {
    const double d = dot2( np, x1 - xp );
    if( d <= par.x )
    {
      if( emit ) { put2( out, base, SG_PLANE_CIRCLE, b, pl, 0u, np, ldx( q0, b ) - par.x * np, fmin( 0.0, d - par.x ) ); }
      cnt = 1u;
    }
  }
  else
  {
    const M2d R = rot2d( __ldg( q1 + 3 * size_t( b ) + 2 ) );
    uint32_t corner = 0u;
    for( int i = -1; i < 2; i += 2 )
    {
      for( int j = -1; j < 2; j += 2 )
      {
        const V2d arm = v2( double( i ) * par.x, double( j ) * par.y );
        const V2d tv = x1 + mul2( R, arm );
        if( dot2( np, tv - xp ) <= 0.0 )
        {
          if( emit ) { put2( out, base + cnt, SG_PLANE_BODY_2D, b, pl, corner, np, arm, nan2() ); }
          ++cnt;
        }
        ++corner;
      }
    }
  }
  return cnt;
}

__global__ void __launch_bounds__( 256 ) k_rb2d_plane_count( const Rb2dDev dev, const __grid_constant__ Planes2D planes, const double* __restrict__ q0, const double* __restrict__ q1, uint32_t* __restrict__ counts )
{
  __shared__ uint32_t s_cnt[SG_MAX_PLANES];
  if( threadIdx.x < planes.n ) { s_cnt[threadIdx.x] = 0u; }
  __syncthreads();
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const ContactOut2X none = {};
  for( uint32_t pl = 0; pl < planes.n; ++pl )
  {
    uint32_t c = ( b < dev.n ) ? rb2d_plane_contacts( dev, planes, pl, b, q0, q1, false, 0ull, none ) : 0u;
    #pragma unroll
    for( int d = 16; d > 0; d >>= 1 ) { c += __shfl_xor_sync( 0xffffffffu, c, d ); }
    if( ( threadIdx.x & 31 ) == 0 && c != 0u ) { atomicAdd( &s_cnt[pl], c ); }
  }
  __syncthreads();
  if( threadIdx.x < planes.n ) { counts[threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x]; }
}

__global__ void __launch_bounds__( 256 ) k_rb2d_plane_emit( const Rb2dDev dev, const __grid_constant__ Planes2D planes, const double* __restrict__ q0, const double* __restrict__ q1,
                                                           const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, const unsigned long long* __restrict__ base_dev, const ContactOut2X out )
{
  __shared__ uint32_t s_warp[8];
  const int mine = ( threadIdx.x < planes.n ) ? int( counts[threadIdx.x * gridDim.x + blockIdx.x] != 0u ) : 0;
  if( __syncthreads_or( mine ) == 0 ) { return; }
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long base = *base_dev;
  const ContactOut2X none = {};
  for( uint32_t pl = 0; pl < planes.n; ++pl )
  {
    if( counts[pl * gridDim.x + blockIdx.x] == 0u ) { continue; }
    const uint32_t c = ( b < dev.n ) ? rb2d_plane_contacts( dev, planes, pl, b, q0, q1, false, 0ull, none ) : 0u;
    uint32_t incl = c;
    #pragma unroll
    for( int d = 1; d < 32; d <<= 1 ) { const uint32_t o = __shfl_up_sync( 0xffffffffu, incl, d ); if( lane >= d ) { incl += o; } }
    __syncthreads();
    if( lane == 31 ) { s_warp[warp] = incl; }
    __syncthreads();
    uint32_t before = incl - c;
    for( int w = 0; w < warp; ++w ) { before += s_warp[w]; }
    if( c != 0u ) { rb2d_plane_contacts( dev, planes, pl, b, q0, q1, true, base + offsets[pl * gridDim.x + blockIdx.x] + before, out ); }
  }
}

struct ScanU32To64b
{
  using In = uint32_t;
  using Acc = unsigned long long;
  using Out = unsigned long long;
  __device__ static Acc zero() { return 0ull; }
  __device__ static Acc conv( const In x ) { return x; }
  __device__ static Acc add( const Acc a, const Acc b ) { return a + b; }
  __device__ static Acc shfl_up( const Acc a, const int d ) { return __shfl_up_sync( 0xffffffffu, a, d ); }
  __device__ static Acc shfl( const Acc a, const int l ) { return __shfl_sync( 0xffffffffu, a, l ); }
  __device__ static Out out( const Acc a ) { return a; }
};

// ---- portals (rigidbody2d/PlanarPortal.h): kernels, per-context data ------------------------------------
#include "sg_rb2d_portal_kernels.cuh"
#include "sg_pair_sort_host.cuh"
#include "sg_rb2d_snapshot.h"

struct Rb2dPortalData
{
  SgPortals2D portals;
  DevBuf rboxes;                                   // double[4 n]: boxes at q1 (the touch tests read them before n + T is known)
  DevBuf tflag, toff, t_partials, ttotal;          // u32[P n] flags / teleported box numbers; T
  DevBuf box_body, box_portal;                     // u32[T]: TeleportedBody table
  DevBuf reg_cnt, tel_cnt, tel_off, pr_partials, tel_total; // u32 per candidate
  DevBuf reg_off, reg_total;                       // u64 per candidate; number of un-teleported pairs
  DevBuf reg_pairs;                                // uint2[]: the un-teleported candidates, in order
  DevBuf tc_key, tc_idx, tc_info, uflag, uoff, u_partials, utotal;
  DevBuf x0t, x1t, delta0, delta1, kick, tp0, tp1; // per teleported contact: constructor arguments
  DevBuf base;                                     // u64: contacts in front of the plane contacts
  DevBuf bad;                                      // u32: bit 0 box in a teleported candidate, bit 1 kinematic body in a teleported collision
  PinBuf h, h_tele;
  uint64_t n_boxes = 0, n_reg = 0, n_tel = 0;
  bool result = false;                             // the last active set came from the portal path
  Rb2dPortalData() { memset( &portals, 0, sizeof( portals ) ); }
  void release()
  {
    DevBuf* bufs[] = { &rboxes, &tflag, &toff, &t_partials, &ttotal, &box_body, &box_portal, &reg_cnt, &tel_cnt, &tel_off, &pr_partials, &tel_total, &reg_off, &reg_total, &reg_pairs,
                       &tc_key, &tc_idx, &tc_info, &uflag, &uoff, &u_partials, &utotal, &x0t, &x1t, &delta0, &delta1, &kick, &tp0, &tp1, &base, &bad };
    for( DevBuf* b : bufs ) { b->release(); }
    h.release(); h_tele.release();
  }
};

// ---- host ------------------------------------------------------------------------------------------
struct Rb2dData
{
  uint32_t n = 0;
  bool flow_resident = false; // q0 (as given) and q1 (as computed) of the last sg_rb2d_flow are still on the device
  double g[2] = { 0.0, 0.0 };
  Planes2D planes;
  std::vector<uint32_t> geo_type;
  std::vector<double> geo_r, geo_half;
  std::vector<uint8_t> h_fixed;          // as given to sg_rb2d_set_bodies (the snapshot writes them back)
  std::vector<uint32_t> h_geo_of_body;
  bool q1_valid = false;                 // ( q1, v1 ) on the device are the output of a flow / step on this context
  DevBuf btype, bparam, M, q0, v0, q1, v1, boxes;
  BroadScratch bp;
  DevBuf pair_counts, pair_offsets, pair_partials, narrow_total, bad_flag;
  DevBuf st_counts, st_offsets, st_partials, st_total;
  DevBuf c_type, c_i, c_j, c_aux, c_n, c_p, c_depth;
  uint64_t act_cap = 0;
  PinBuf h_totals, h_out;
  uint64_t n_cand = 0, n_bb = 0, n_static = 0;
  bool have_result = false;
  Rb2dPortalData* px = nullptr; // allocated by sg_rb2d_set_portals
  Rb2dData() { memset( &planes, 0, sizeof( planes ) ); }
};

void sg_rb2d_release( sg_ctx* ctx )
{
  Rb2dData* d = ctx->rb2d;
  if( d == nullptr ) { return; }
  DevBuf* bufs[] = { &d->btype, &d->bparam, &d->M, &d->q0, &d->v0, &d->q1, &d->v1, &d->boxes, &d->pair_counts, &d->pair_offsets, &d->pair_partials, &d->narrow_total, &d->bad_flag,
                     &d->st_counts, &d->st_offsets, &d->st_partials, &d->st_total, &d->c_type, &d->c_i, &d->c_j, &d->c_aux, &d->c_n, &d->c_p, &d->c_depth };
  for( DevBuf* b : bufs ) { b->release(); }
  d->bp.release(); d->h_totals.release(); d->h_out.release();
  if( d->px != nullptr ) { d->px->release(); delete d->px; d->px = nullptr; }
  delete d;
  ctx->rb2d = nullptr;
}

static Rb2dData* rb2d_data( sg_ctx* ctx ) { if( ctx->rb2d == nullptr ) { ctx->rb2d = new Rb2dData; } return ctx->rb2d; }

static ContactOut2X rb2d_out( const Rb2dData* d )
{
  ContactOut2X out;
  out.type = d->c_type.as<uint32_t>(); out.i = d->c_i.as<uint32_t>(); out.j = d->c_j.as<uint32_t>(); out.aux = d->c_aux.as<uint32_t>();
  out.n = d->c_n.as<double2>(); out.p = d->c_p.as<double2>(); out.depth = d->c_depth.as<double>(); out.cap = d->act_cap;
  return out;
}

static int rb2d_ensure_contacts( sg_ctx* ctx, Rb2dData* d, const uint64_t cap )
{
  if( cap <= d->act_cap ) { return SG_OK; }
  SG_CUDA( ctx, d->c_type.ensure( size_t( cap ) * 4 ) ); SG_CUDA( ctx, d->c_i.ensure( size_t( cap ) * 4 ) ); SG_CUDA( ctx, d->c_j.ensure( size_t( cap ) * 4 ) );
  SG_CUDA( ctx, d->c_aux.ensure( size_t( cap ) * 4 ) ); SG_CUDA( ctx, d->c_n.ensure( size_t( cap ) * 16 ) ); SG_CUDA( ctx, d->c_p.ensure( size_t( cap ) * 16 ) );
  SG_CUDA( ctx, d->c_depth.ensure( size_t( cap ) * 8 ) );
  d->act_cap = cap;
  return SG_OK;
}

static int rb2d_active_set_device( sg_ctx* ctx, Rb2dData* d )
{
  const uint32_t n = d->n;
  d->n_cand = d->n_bb = d->n_static = 0;
  d->have_result = true;
  if( d->px != nullptr ) { d->px->result = false; }
  if( n == 0 ) { return SG_OK; }
  Rb2dDev dev; dev.n = n; dev.btype = d->btype.as<uint32_t>(); dev.bparam = d->bparam.as<double2>();
  SG_CUDA( ctx, d->h_totals.ensure( 64 ) );
  SG_CUDA( ctx, d->st_total.ensure( 4 ) ); SG_CUDA( ctx, d->narrow_total.ensure( 8 ) ); SG_CUDA( ctx, d->bad_flag.ensure( 4 ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->st_total.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->narrow_total.ptr, 0, 8, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->bad_flag.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, d->boxes.ensure( size_t( n ) * 32 ) );
  SG_LAUNCH( ctx, "rb2d_aabb", double( n ) * ( 48.0 + 32.0 ), k_rb2d_aabb<<<sg_div_up( n, 256 ), 256, 0, ctx->stream>>>( dev, d->q0.as<double>(), d->q1.as<double>(), d->boxes.as<double>() ) );
  int rc = sg_bp_prepare_scratch<Box2DPolicy>( ctx, d->bp, n );
  if( rc != SG_OK ) { return rc; }
  Box2DIn in; in.boxes = d->boxes.as<double>(); in.n = n;
  rc = sg_bp_bin_and_count<Box2DPolicy>( ctx, d->bp, in );
  if( rc != SG_OK ) { return rc; }
  unsigned long long* ht = d->h_totals.as<unsigned long long>();
  SG_CUDA( ctx, cudaMemcpyAsync( ht, d->bp.totals.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  const uint64_t np = ht[0];
  d->n_cand = np;
  if( np >= 0xffffffffull ) { return sg_fail( ctx, SG_ERR_INTERNAL, "rb2d: more than 2^32 candidate pairs" ); }
  if( np + 64 > d->bp.cand_cap ) { SG_CUDA( ctx, d->bp.cand.ensure( size_t( np + 64 ) * sizeof( uint2 ) ) ); d->bp.cand_cap = d->bp.cand.cap / sizeof( uint2 ); }
  if( np > 0 )
  {
    rc = sg_bp_emit_lists<Box2DPolicy>( ctx, d->bp, in, n, true, NoOut2D{}, 0u );
    if( rc != SG_OK ) { return rc; }
  }
  SG_CUDA( ctx, d->pair_counts.ensure( size_t( np ) * 4 + 4 ) );
  SG_CUDA( ctx, d->pair_offsets.ensure( size_t( np ) * 8 + 8 ) );
  SG_CUDA( ctx, d->pair_partials.ensure( ( size_t( np ) / SG_SCAN_TILE + 2 ) * 8 ) );
  const unsigned long long* npairs_dev = &d->bp.totals.as<ScanPairCounts::Acc>()->c;
  if( np > 0 )
  {
    SG_LAUNCH( ctx, "rb2d_pairs_count", double( np ) * 120.0, k_rb2d_pairs<false><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( dev, d->bp.cand.as<uint2>(), npairs_dev, d->q0.as<double>(), d->q1.as<double>(),
               d->pair_counts.as<uint32_t>(), nullptr, rb2d_out( d ), d->bad_flag.as<uint32_t>() ) );
    rc = sg_exclusive_scan<ScanU32To64b>( ctx, "rb2d_pair_scan", d->pair_counts.as<uint32_t>(), nullptr, uint32_t( np ), uint32_t( np ), d->pair_partials.as<unsigned long long>(), d->pair_offsets.as<unsigned long long>(),
                                          d->narrow_total.as<unsigned long long>(), false );
    if( rc != SG_OK ) { return rc; }
  }
  const uint32_t npl = d->planes.n;
  const unsigned nblk = sg_div_up( n, 256 );
  const uint32_t nst = npl * nblk;
  if( npl > 0 )
  {
    SG_CUDA( ctx, d->st_counts.ensure( size_t( nst ) * 4 + 4 ) ); SG_CUDA( ctx, d->st_offsets.ensure( size_t( nst ) * 4 + 4 ) );
    SG_CUDA( ctx, d->st_partials.ensure( ( size_t( nst ) / SG_SCAN_TILE + 2 ) * 4 ) );
    SG_LAUNCH( ctx, "rb2d_plane_count", double( n ) * 40.0, k_rb2d_plane_count<<<nblk, 256, 0, ctx->stream>>>( dev, d->planes, d->q0.as<double>(), d->q1.as<double>(), d->st_counts.as<uint32_t>() ) );
    rc = sg_exclusive_scan<ScanU32>( ctx, "rb2d_plane_scan", d->st_counts.as<uint32_t>(), nullptr, nst, nst, d->st_partials.as<uint32_t>(), d->st_offsets.as<uint32_t>(), d->st_total.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
  }
  uint32_t* hbad = reinterpret_cast<uint32_t*>( ht + 4 );
  ht[2] = 0ull;
  SG_CUDA( ctx, cudaMemcpyAsync( ht + 1, d->narrow_total.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( ht + 2, d->st_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( hbad, d->bad_flag.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  if( *hbad != 0u )
  {
    return sg_fail( ctx, SG_ERR_UNSUPPORTED, "kinematic box-box / kinematic-circle-vs-box collisions are not supported (the reference exits here: rigidbody2d/RigidBody2DSim.cpp:186-190, 210-214)" );
  }
  d->n_bb = ht[1]; d->n_static = ht[2] & 0xffffffffull;
  rc = rb2d_ensure_contacts( ctx, d, d->n_bb + d->n_static + 64 );
  if( rc != SG_OK ) { return rc; }
  if( np > 0 && d->n_bb > 0 )
  {
    SG_LAUNCH( ctx, "rb2d_pairs_emit", double( np ) * 120.0 + double( d->n_bb ) * 60.0, k_rb2d_pairs<true><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( dev, d->bp.cand.as<uint2>(), npairs_dev, d->q0.as<double>(), d->q1.as<double>(),
               nullptr, d->pair_offsets.as<unsigned long long>(), rb2d_out( d ), d->bad_flag.as<uint32_t>() ) );
  }
  if( npl > 0 && d->n_static > 0 )
  {
    SG_LAUNCH( ctx, "rb2d_plane_emit", double( n ) * 4.0, k_rb2d_plane_emit<<<nblk, 256, 0, ctx->stream>>>( dev, d->planes, d->q0.as<double>(), d->q1.as<double>(), d->st_counts.as<uint32_t>(), d->st_offsets.as<uint32_t>(),
               d->narrow_total.as<unsigned long long>(), rb2d_out( d ) ) );
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

// RigidBody2DSim::computeActiveSet with portals (rigidbody2d/RigidBody2DSim.cpp:696-714 -> :876-1040): boxes at q1, teleported copies,
// un-teleported candidates through the regular narrow phase (k_rb2d_pairs), TeleportedCollision set for the rest, then the planes.
// Same sequence as the ball2d portal path (sg_ball2d_portals.cuh); list sizes are read back where buffers have to be sized.
static int rb2d_portal_active_set_device( sg_ctx* ctx, Rb2dData* d )
{
  Rb2dPortalData* x = d->px;
  const uint32_t n = d->n;
  const uint32_t np_portals = x->portals.n;
  d->n_cand = d->n_bb = d->n_static = 0;
  x->n_boxes = x->n_reg = x->n_tel = 0;
  d->have_result = true;
  x->result = true;
  if( n == 0 ) { return SG_OK; }
  if( uint64_t( n ) * np_portals >= 0x80000000ull ) { return sg_fail( ctx, SG_ERR_INVALID, "rigidbody2d portals: bodies x portals must stay below 2^31" ); }
  Rb2dDev dev; dev.n = n; dev.btype = d->btype.as<uint32_t>(); dev.bparam = d->bparam.as<double2>();
  const unsigned nblk = sg_div_up( n, 256 );
  const uint32_t nflag = n * np_portals;
  SG_CUDA( ctx, x->h.ensure( 128 ) );
  SG_CUDA( ctx, d->st_total.ensure( 4 ) ); SG_CUDA( ctx, d->narrow_total.ensure( 8 ) ); SG_CUDA( ctx, d->bad_flag.ensure( 4 ) ); SG_CUDA( ctx, x->bad.ensure( 4 ) );
  SG_CUDA( ctx, x->ttotal.ensure( 4 ) ); SG_CUDA( ctx, x->reg_total.ensure( 8 ) ); SG_CUDA( ctx, x->tel_total.ensure( 4 ) ); SG_CUDA( ctx, x->utotal.ensure( 4 ) ); SG_CUDA( ctx, x->base.ensure( 8 ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->st_total.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->narrow_total.ptr, 0, 8, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->bad_flag.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->bad.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->ttotal.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->reg_total.ptr, 0, 8, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->tel_total.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->utotal.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, x->rboxes.ensure( size_t( n ) * 32 ) );
  SG_CUDA( ctx, x->tflag.ensure( size_t( nflag ) * 4 + 4 ) ); SG_CUDA( ctx, x->toff.ensure( size_t( nflag ) * 4 + 4 ) );
  SG_CUDA( ctx, x->t_partials.ensure( ( size_t( nflag ) / SG_SCAN_TILE + 2 ) * 4 ) );
  SG_LAUNCH( ctx, "r2p_boxes", double( n ) * ( 44.0 + 32.0 ), k_r2p_boxes<<<nblk, 256, 0, ctx->stream>>>( dev, d->q1.as<double>(), x->rboxes.as<double>() ) );
  SG_LAUNCH( ctx, "r2p_touch", double( nflag ) * 36.0, k_r2p_touch<<<dim3( nblk, np_portals ), 256, 0, ctx->stream>>>( x->portals, n, x->rboxes.as<double>(), x->tflag.as<uint32_t>() ) );
  int rc = sg_exclusive_scan<ScanU32>( ctx, "r2p_touch_scan", x->tflag.as<uint32_t>(), nullptr, nflag, nflag, x->t_partials.as<uint32_t>(), x->toff.as<uint32_t>(), x->ttotal.as<uint32_t>(), false );
  if( rc != SG_OK ) { return rc; }
  // the planes do not depend on the portals
  const uint32_t npl = d->planes.n;
  const uint32_t nst = npl * nblk;
  if( npl > 0 )
  {
    SG_CUDA( ctx, d->st_counts.ensure( size_t( nst ) * 4 + 4 ) ); SG_CUDA( ctx, d->st_offsets.ensure( size_t( nst ) * 4 + 4 ) );
    SG_CUDA( ctx, d->st_partials.ensure( ( size_t( nst ) / SG_SCAN_TILE + 2 ) * 4 ) );
    SG_LAUNCH( ctx, "rb2d_plane_count", double( n ) * 40.0, k_rb2d_plane_count<<<nblk, 256, 0, ctx->stream>>>( dev, d->planes, d->q0.as<double>(), d->q1.as<double>(), d->st_counts.as<uint32_t>() ) );
    rc = sg_exclusive_scan<ScanU32>( ctx, "rb2d_plane_scan", d->st_counts.as<uint32_t>(), nullptr, nst, nst, d->st_partials.as<uint32_t>(), d->st_offsets.as<uint32_t>(), d->st_total.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
  }
  uint32_t* h32 = x->h.as<uint32_t>();
  unsigned long long* h64 = x->h.as<unsigned long long>() + 8; // bytes 64...
  h32[1] = 0u;
  SG_CUDA( ctx, cudaMemcpyAsync( h32, x->ttotal.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  if( npl > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( h32 + 1, d->st_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  const uint32_t nt = h32[0];
  d->n_static = h32[1];
  x->n_boxes = nt;
  if( uint64_t( n ) + nt >= 0x80000000ull ) { return sg_fail( ctx, SG_ERR_INVALID, "rigidbody2d portals: more than 2^31 - 1 boxes" ); }
  const uint32_t next = n + nt;
  SG_CUDA( ctx, d->boxes.ensure( size_t( next ) * 32 ) );
  SG_CUDA( ctx, x->box_body.ensure( size_t( nt ) * 4 + 4 ) ); SG_CUDA( ctx, x->box_portal.ensure( size_t( nt ) * 4 + 4 ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->boxes.ptr, x->rboxes.ptr, size_t( n ) * 32, cudaMemcpyDeviceToDevice, ctx->stream ) );
  if( nt > 0 )
  {
    SG_LAUNCH( ctx, "r2p_tele_boxes", double( nflag ) * 8.0 + double( nt ) * 100.0, k_r2p_tele_boxes<<<dim3( nblk, np_portals ), 256, 0, ctx->stream>>>( x->portals, dev, d->q1.as<double>(), x->rboxes.as<double>(),
               x->tflag.as<uint32_t>(), x->toff.as<uint32_t>(), d->boxes.as<double>(), x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>() ) );
  }
  rc = sg_bp_prepare_scratch<Box2DPolicy>( ctx, d->bp, next );
  if( rc != SG_OK ) { return rc; }
  Box2DIn in; in.boxes = d->boxes.as<double>(); in.n = next;
  rc = sg_bp_bin_and_count<Box2DPolicy>( ctx, d->bp, in );
  if( rc != SG_OK ) { return rc; }
  SG_CUDA( ctx, cudaMemcpyAsync( h64, d->bp.totals.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  const uint64_t np = h64[0];
  d->n_cand = np;
  if( np >= 0xffffffffull ) { return sg_fail( ctx, SG_ERR_INTERNAL, "rigidbody2d portals: more than 2^32 candidate pairs" ); }
  if( np + 64 > d->bp.cand_cap ) { SG_CUDA( ctx, d->bp.cand.ensure( size_t( np + 64 ) * sizeof( uint2 ) ) ); d->bp.cand_cap = d->bp.cand.cap / sizeof( uint2 ); }
  uint64_t nreg_pairs = 0;
  uint32_t nraw = 0;
  if( np > 0 )
  {
    rc = sg_bp_emit_lists<Box2DPolicy>( ctx, d->bp, in, next, true, NoOut2D{}, 0u );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, x->reg_cnt.ensure( size_t( np ) * 4 + 4 ) ); SG_CUDA( ctx, x->reg_off.ensure( size_t( np ) * 8 + 8 ) );
    SG_CUDA( ctx, x->tel_cnt.ensure( size_t( np ) * 4 + 4 ) ); SG_CUDA( ctx, x->tel_off.ensure( size_t( np ) * 4 + 4 ) );
    SG_CUDA( ctx, x->pr_partials.ensure( ( size_t( np ) / SG_SCAN_TILE + 2 ) * 8 ) );
    SG_LAUNCH( ctx, "r2p_classify_count", double( np ) * 60.0, k_r2p_classify<false><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( x->portals, dev, d->bp.cand.as<uint2>(), np, d->q1.as<double>(),
               x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>(), x->reg_cnt.as<uint32_t>(), x->tel_cnt.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, x->bad.as<uint32_t>() ) );
    rc = sg_exclusive_scan<ScanU32To64b>( ctx, "r2p_regular_scan", x->reg_cnt.as<uint32_t>(), nullptr, uint32_t( np ), uint32_t( np ), x->pr_partials.as<unsigned long long>(), x->reg_off.as<unsigned long long>(),
                                          x->reg_total.as<unsigned long long>(), false );
    if( rc != SG_OK ) { return rc; }
    rc = sg_exclusive_scan<ScanU32>( ctx, "r2p_teleported_scan", x->tel_cnt.as<uint32_t>(), nullptr, uint32_t( np ), uint32_t( np ), x->pr_partials.as<uint32_t>(), x->tel_off.as<uint32_t>(), x->tel_total.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, cudaMemcpyAsync( h64 + 2, x->reg_total.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h32 + 2, x->tel_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h32 + 3, x->bad.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    if( h32[3] != 0u )
    {
      return sg_fail( ctx, SG_ERR_UNSUPPORTED, "a box takes part in a teleported collision (the reference exits here: rigidbody2d/RigidBody2DSim.cpp:374-385)" );
    }
    nreg_pairs = h64[2];
    nraw = h32[2];
  }
  if( nraw > 0x40000000u ) { return sg_fail( ctx, SG_ERR_INTERNAL, "rigidbody2d portals: more than 2^30 teleported collisions" ); }
  uint32_t m = 1u;
  while( m < nraw ) { m <<= 1; }
  SG_CUDA( ctx, x->reg_pairs.ensure( size_t( nreg_pairs ) * 8 + 8 ) );
  if( nraw > 0 )
  {
    SG_CUDA( ctx, x->tc_key.ensure( size_t( m ) * 8 ) ); SG_CUDA( ctx, x->tc_idx.ensure( size_t( m ) * 4 ) ); SG_CUDA( ctx, x->tc_info.ensure( size_t( nraw ) * 16 ) );
    SG_CUDA( ctx, x->uflag.ensure( size_t( nraw ) * 4 + 4 ) ); SG_CUDA( ctx, x->uoff.ensure( size_t( nraw ) * 4 + 4 ) );
    SG_CUDA( ctx, x->u_partials.ensure( ( size_t( nraw ) / SG_SCAN_TILE + 2 ) * 4 ) );
    for( DevBuf* b : { &x->x0t, &x->x1t, &x->delta0, &x->delta1, &x->kick } ) { SG_CUDA( ctx, b->ensure( size_t( nraw ) * 16 ) ); }
    SG_CUDA( ctx, x->tp0.ensure( size_t( nraw ) * 4 ) ); SG_CUDA( ctx, x->tp1.ensure( size_t( nraw ) * 4 ) );
  }
  if( np > 0 )
  {
    SG_LAUNCH( ctx, "r2p_classify_emit", double( np ) * 28.0, k_r2p_classify<true><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( x->portals, dev, d->bp.cand.as<uint2>(), np, d->q1.as<double>(),
               x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>(), x->reg_cnt.as<uint32_t>(), x->tel_cnt.as<uint32_t>(), x->reg_off.as<unsigned long long>(), x->tel_off.as<uint32_t>(),
               x->reg_pairs.as<uint2>(), x->tc_key.as<unsigned long long>(), x->tc_idx.as<uint32_t>(), x->tc_info.as<uint4>(), x->bad.as<uint32_t>() ) );
  }
  // regular narrow phase over the un-teleported candidates: count -> scan (the emit follows once the contact arrays are sized)
  SG_CUDA( ctx, d->pair_counts.ensure( size_t( nreg_pairs ) * 4 + 4 ) );
  SG_CUDA( ctx, d->pair_offsets.ensure( size_t( nreg_pairs ) * 8 + 8 ) );
  SG_CUDA( ctx, d->pair_partials.ensure( ( size_t( nreg_pairs ) / SG_SCAN_TILE + 2 ) * 8 ) );
  const unsigned long long* nreg_dev = x->reg_total.as<unsigned long long>();
  if( nreg_pairs > 0 )
  {
    SG_LAUNCH( ctx, "rb2d_pairs_count", double( nreg_pairs ) * 120.0, k_rb2d_pairs<false><<<sg_div_up( nreg_pairs, 128 ), 128, 0, ctx->stream>>>( dev, x->reg_pairs.as<uint2>(), nreg_dev, d->q0.as<double>(), d->q1.as<double>(),
               d->pair_counts.as<uint32_t>(), nullptr, rb2d_out( d ), d->bad_flag.as<uint32_t>() ) );
    rc = sg_exclusive_scan<ScanU32To64b>( ctx, "rb2d_pair_scan", d->pair_counts.as<uint32_t>(), nullptr, uint32_t( nreg_pairs ), uint32_t( nreg_pairs ), d->pair_partials.as<unsigned long long>(),
                                          d->pair_offsets.as<unsigned long long>(), d->narrow_total.as<unsigned long long>(), false );
    if( rc != SG_OK ) { return rc; }
  }
  if( nraw > 0 )
  {
    rc = sg_tele_sort_unique( ctx, nraw, m, x->tc_key.as<unsigned long long>(), x->tc_idx.as<uint32_t>(), x->uflag.as<uint32_t>(), x->uoff.as<uint32_t>(), x->u_partials.as<uint32_t>(), x->utotal.as<uint32_t>() );
    if( rc != SG_OK ) { return rc; }
  }
  h64[3] = 0ull; h32[4] = 0u; h32[5] = 0u;
  SG_CUDA( ctx, cudaMemcpyAsync( h64 + 3, d->narrow_total.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( h32 + 4, d->bad_flag.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  if( nraw > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( h32 + 5, x->utotal.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  if( h32[4] != 0u )
  {
    return sg_fail( ctx, SG_ERR_UNSUPPORTED, "kinematic box-box / kinematic-circle-vs-box collisions are not supported (the reference exits here: rigidbody2d/RigidBody2DSim.cpp:186-190, 210-214)" );
  }
  x->n_reg = h64[3];
  x->n_tel = h32[5];
  d->n_bb = x->n_reg + x->n_tel;
  rc = rb2d_ensure_contacts( ctx, d, d->n_bb + d->n_static + 64 );
  if( rc != SG_OK ) { return rc; }
  if( nreg_pairs > 0 && x->n_reg > 0 )
  {
    SG_LAUNCH( ctx, "rb2d_pairs_emit", double( nreg_pairs ) * 120.0 + double( x->n_reg ) * 60.0, k_rb2d_pairs<true><<<sg_div_up( nreg_pairs, 128 ), 128, 0, ctx->stream>>>( dev, x->reg_pairs.as<uint2>(), nreg_dev, d->q0.as<double>(),
               d->q1.as<double>(), nullptr, d->pair_offsets.as<unsigned long long>(), rb2d_out( d ), d->bad_flag.as<uint32_t>() ) );
  }
  if( x->n_tel > 0 )
  {
    SG_LAUNCH( ctx, "r2p_tele_contacts", double( nraw ) * 250.0, k_r2p_tele_contacts<<<sg_div_up( nraw, 128 ), 128, 0, ctx->stream>>>( x->portals, dev, nraw, x->tc_idx.as<uint32_t>(), x->uflag.as<uint32_t>(), x->uoff.as<uint32_t>(),
               x->tc_info.as<uint4>(), d->q0.as<double>(), d->q1.as<double>(), x->n_reg, rb2d_out( d ), x->x0t.as<double2>(), x->x1t.as<double2>(), x->delta0.as<double2>(), x->delta1.as<double2>(),
               x->kick.as<double2>(), x->tp0.as<uint32_t>(), x->tp1.as<uint32_t>(), x->bad.as<uint32_t>() ) );
  }
  if( npl > 0 && d->n_static > 0 )
  {
    h64[4] = d->n_bb;
    SG_CUDA( ctx, cudaMemcpyAsync( x->base.ptr, h64 + 4, 8, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_LAUNCH( ctx, "rb2d_plane_emit", double( n ) * 4.0, k_rb2d_plane_emit<<<nblk, 256, 0, ctx->stream>>>( dev, d->planes, d->q0.as<double>(), d->q1.as<double>(), d->st_counts.as<uint32_t>(), d->st_offsets.as<uint32_t>(),
               x->base.as<unsigned long long>(), rb2d_out( d ) ) );
  }
  SG_CUDA( ctx, cudaMemcpyAsync( h32 + 3, x->bad.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  if( h32[3] != 0u )
  {
    return sg_fail( ctx, SG_ERR_UNSUPPORTED, "a kinematically scripted body takes part in a teleported collision (the reference exits here: rigidbody2d/RigidBody2DSim.cpp:481-485)" );
  }
  return SG_OK;
}

static int rb2d_copy_out( sg_ctx* ctx, Rb2dData* d, const uint32_t flags, sg_contacts* out )
{
  memset( out, 0, sizeof( *out ) );
  out->dim = 2;
  out->n_candidates = d->n_cand; out->n_body_body = d->n_bb; out->n_plane = d->n_static;
  const uint64_t na = d->n_bb + d->n_static;
  out->n_active = na;
  const bool want_cand = ( flags & SG_OUT_CANDIDATES ) != 0u;
  auto al = []( size_t b ) { return ( b + 63 ) & ~size_t( 63 ); };
  size_t bytes = 64;
  const size_t o_type = bytes; bytes += al( na * 4 );
  const size_t o_i = bytes; bytes += al( na * 4 );
  const size_t o_j = bytes; bytes += al( na * 4 );
  const size_t o_aux = bytes; bytes += al( na * 4 );
  const size_t o_n = bytes; bytes += al( na * 16 );
  const size_t o_p = bytes; bytes += al( na * 16 );
  const size_t o_d = bytes; bytes += al( na * 8 );
  const size_t o_c = bytes; if( want_cand ) { bytes += al( d->n_cand * 8 ); }
  SG_CUDA( ctx, d->h_out.ensure( bytes ) );
  char* h = d->h_out.as<char>();
  if( na > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_type, d->c_type.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_i, d->c_i.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_j, d->c_j.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_aux, d->c_aux.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_n, d->c_n.ptr, na * 16, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_p, d->c_p.ptr, na * 16, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_d, d->c_depth.ptr, na * 8, cudaMemcpyDeviceToHost, ctx->stream ) );
  }
  if( want_cand && d->n_cand > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_c, d->bp.cand.ptr, d->n_cand * 8, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  out->type = reinterpret_cast<const uint32_t*>( h + o_type ); out->i = reinterpret_cast<const uint32_t*>( h + o_i );
  out->j = reinterpret_cast<const uint32_t*>( h + o_j ); out->aux = reinterpret_cast<const uint32_t*>( h + o_aux );
  out->n = reinterpret_cast<const double*>( h + o_n ); out->p = reinterpret_cast<const double*>( h + o_p ); out->depth = reinterpret_cast<const double*>( h + o_d );
  out->cand_ij = want_cand ? reinterpret_cast<const uint32_t*>( h + o_c ) : nullptr;
  return SG_OK;
}

extern "C"
{

int sg_rb2d_set_geometry( sg_ctx* ctx, uint32_t ngeo, const uint32_t* type, const double* r, const double* half )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( ngeo > 0 && ( type == nullptr || r == nullptr || half == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_geometry: null array" ); }
  Rb2dData* d = rb2d_data( ctx );
  d->geo_type.assign( type, type + ngeo ); d->geo_r.assign( r, r + ngeo ); d->geo_half.assign( half, half + 2 * size_t( ngeo ) );
  return SG_OK;
}

int sg_rb2d_set_bodies( sg_ctx* ctx, uint32_t n, const uint32_t* geo_of_body, const uint8_t* fixed, const double* M )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > 0 && ( geo_of_body == nullptr || fixed == nullptr || M == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_bodies: null array" ); }
  if( n >= 0x80000000u ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_bodies: at most 2^31 - 1 bodies" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  Rb2dData* d = rb2d_data( ctx );
  std::vector<uint32_t> btype( n );
  std::vector<double> bparam( 2 * size_t( n ), 0.0 );
  for( uint32_t b = 0; b < n; ++b )
  {
    const uint32_t gi = geo_of_body[b];
    if( gi >= d->geo_type.size() ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_bodies: body %u refers to geometry %u of %zu", b, gi, d->geo_type.size() ); }
    const uint32_t t = d->geo_type[gi];
    if( t != SG_GEO2_CIRCLE && t != SG_GEO2_BOX ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb2d_set_bodies: geometry type %u is not supported", t ); }
    btype[b] = t | ( fixed[b] ? SG_FIXED_BIT2 : 0u );
    if( t == SG_GEO2_CIRCLE ) { bparam[2 * size_t( b )] = d->geo_r[gi]; }
    else { bparam[2 * size_t( b )] = d->geo_half[2 * size_t( gi )]; bparam[2 * size_t( b ) + 1] = d->geo_half[2 * size_t( gi ) + 1]; }
  }
  d->n = n;
  d->have_result = false;
  d->flow_resident = false;
  d->q1_valid = false;
  d->h_fixed.assign( fixed, fixed + n ); d->h_geo_of_body.assign( geo_of_body, geo_of_body + n );
  if( n == 0 ) { return SG_OK; }
  SG_CUDA( ctx, d->btype.ensure( size_t( n ) * 4 ) ); SG_CUDA( ctx, d->bparam.ensure( size_t( n ) * 16 ) ); SG_CUDA( ctx, d->M.ensure( size_t( n ) * 24 ) );
  SG_CUDA( ctx, d->q0.ensure( size_t( n ) * 24 ) ); SG_CUDA( ctx, d->q1.ensure( size_t( n ) * 24 ) ); SG_CUDA( ctx, d->v0.ensure( size_t( n ) * 24 ) ); SG_CUDA( ctx, d->v1.ensure( size_t( n ) * 24 ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->btype.ptr, btype.data(), size_t( n ) * 4, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->bparam.ptr, bparam.data(), size_t( n ) * 16, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->M.ptr, M, size_t( n ) * 24, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_rb2d_set_gravity( sg_ctx* ctx, const double* g )
{
  if( ctx == nullptr || g == nullptr ) { return SG_ERR_INVALID; }
  Rb2dData* d = rb2d_data( ctx );
  d->g[0] = g[0]; d->g[1] = g[1];
  return SG_OK;
}

int sg_rb2d_set_planes( sg_ctx* ctx, uint32_t n, const double* x, const double* nrm )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > SG_MAX_PLANES ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_planes: at most %d planes", SG_MAX_PLANES ); }
  if( n > 0 && ( x == nullptr || nrm == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_planes: null array" ); }
  Rb2dData* d = rb2d_data( ctx );
  d->planes.n = n;
  for( uint32_t p = 0; p < n; ++p ) { for( int k = 0; k < 2; ++k ) { d->planes.x[p][k] = x[2 * p + k]; d->planes.nrm[p][k] = nrm[2 * p + k]; } }
  return SG_OK;
}

int sg_rb2d_set_portals( sg_ctx* ctx, uint32_t n, const double* plane_a_x, const double* plane_a_n, const double* plane_b_x, const double* plane_b_n, const double* v, const double* bounds )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > SG_MAX_PORTALS ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_portals: at most %d portals", SG_MAX_PORTALS ); }
  if( n > 0 && ( plane_a_x == nullptr || plane_a_n == nullptr || plane_b_x == nullptr || plane_b_n == nullptr || v == nullptr || bounds == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_portals: null array" ); }
  Rb2dData* d = rb2d_data( ctx );
  if( n == 0 && d->px == nullptr ) { return SG_OK; }
  if( d->px == nullptr ) { d->px = new Rb2dPortalData; }
  Rb2dPortalData* x = d->px;
  SgPortals2D ps; // assembled on the side: a rejected call leaves the portals as they were
  memset( &ps, 0, sizeof( ps ) );
  ps.n = n;
  for( uint32_t p = 0; p < n; ++p )
  {
    SgPortal2D& pt = ps.p[p];
    for( int k = 0; k < 2; ++k ) { pt.ax[k] = plane_a_x[2 * p + k]; pt.bx[k] = plane_b_x[2 * p + k]; }
    // RigidBody2DStaticPlane::RigidBody2DStaticPlane( x, n ): n as given (rigidbody2d/RigidBody2DStaticPlane.cpp:10-14)
    sg_portal_plane_frame_as_given( plane_a_n + 2 * p, pt.an, pt.at );
    sg_portal_plane_frame_as_given( plane_b_n + 2 * p, pt.bn, pt.bt );
    if( bounds[p] < 0.0 ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_set_portals: portal %u has negative bounds", p ); }
    pt.v = v[p]; pt.bounds = bounds[p]; pt.dx = 0.0;
  }
  x->portals = ps;
  d->have_result = false;
  return SG_OK;
}

int sg_rb2d_update_portals( sg_ctx* ctx, double t, double* dx_out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb2dData* d = rb2d_data( ctx );
  if( d->px == nullptr ) { return SG_OK; }
  for( uint32_t p = 0; p < d->px->portals.n; ++p )
  {
    SgPortal2D& pt = d->px->portals.p[p];
    pt.dx = sg_portal_offset( pt.v, pt.bounds, t );
    if( dx_out != nullptr ) { dx_out[p] = pt.dx; }
  }
  return SG_OK;
}

int sg_rb2d_enforce_portals( sg_ctx* ctx, double* q, double* v )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb2dData* d = rb2d_data( ctx );
  if( d->px == nullptr || d->px->portals.n == 0u || d->n == 0 ) { return SG_OK; }
  if( q == nullptr || v == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_enforce_portals: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  d->flow_resident = false; // q1 / v1 serve as the staging copies
  const size_t bytes = size_t( d->n ) * 24;
  SG_CUDA( ctx, cudaMemcpyAsync( d->q1.ptr, q, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->v1.ptr, v, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_LAUNCH( ctx, "r2p_enforce", double( d->n ) * 64.0, k_r2p_enforce<<<sg_div_up( d->n, 256 ), 256, 0, ctx->stream>>>( d->px->portals, d->n, d->q1.as<double>(), d->v1.as<double>() ) );
  SG_CUDA( ctx, cudaMemcpyAsync( q, d->q1.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( v, d->v1.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

int sg_rb2d_teleported( sg_ctx* ctx, sg_teleported* out )
{
  if( ctx == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  memset( out, 0, sizeof( *out ) );
  Rb2dData* d = rb2d_data( ctx );
  if( !d->have_result || d->px == nullptr || !d->px->result ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_teleported: the last active set was not computed with portals" ); }
  Rb2dPortalData* x = d->px;
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  auto al = []( size_t b ) { return ( b + 63 ) & ~size_t( 63 ); };
  const size_t nb = x->n_boxes, nt = x->n_tel;
  size_t bytes = 64;
  const size_t o_bb = bytes; bytes += al( nb * 4 );
  const size_t o_bp = bytes; bytes += al( nb * 4 );
  const size_t o_p0 = bytes; bytes += al( nt * 4 );
  const size_t o_p1 = bytes; bytes += al( nt * 4 );
  size_t o_d[5];
  for( int k = 0; k < 5; ++k ) { o_d[k] = bytes; bytes += al( nt * 16 ); }
  SG_CUDA( ctx, x->h_tele.ensure( bytes ) );
  char* h = x->h_tele.as<char>();
  if( nb > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_bb, x->box_body.ptr, nb * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_bp, x->box_portal.ptr, nb * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  }
  if( nt > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_p0, x->tp0.ptr, nt * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_p1, x->tp1.ptr, nt * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    const DevBuf* src[5] = { &x->x0t, &x->x1t, &x->kick, &x->delta0, &x->delta1 };
    for( int k = 0; k < 5; ++k ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_d[k], src[k]->ptr, nt * 16, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  out->n_boxes = nb; out->n_regular = x->n_reg; out->n_teleported = nt;
  out->box_body = reinterpret_cast<const uint32_t*>( h + o_bb ); out->box_portal = reinterpret_cast<const uint32_t*>( h + o_bp );
  out->portal0 = reinterpret_cast<const uint32_t*>( h + o_p0 ); out->portal1 = reinterpret_cast<const uint32_t*>( h + o_p1 );
  out->x0 = reinterpret_cast<const double*>( h + o_d[0] ); out->x1 = reinterpret_cast<const double*>( h + o_d[1] ); out->kick = reinterpret_cast<const double*>( h + o_d[2] );
  out->delta0 = reinterpret_cast<const double*>( h + o_d[3] ); out->delta1 = reinterpret_cast<const double*>( h + o_d[4] );
  return SG_OK;
}

int sg_rb2d_flow( sg_ctx* ctx, int map_kind, const double* q0, const double* v0, double dt, double* q1, double* v1 )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( map_kind != SG_MAP_SYMPLECTIC_EULER && map_kind != SG_MAP_VERLET ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_flow: map kind %d is not a rigidbody2d map", map_kind ); }
  Rb2dData* d = rb2d_data( ctx );
  if( d->n == 0 ) { return SG_OK; }
  if( q0 == nullptr || v0 == nullptr || q1 == nullptr || v1 == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_flow: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const size_t bytes = size_t( d->n ) * 24;
  SG_CUDA( ctx, cudaMemcpyAsync( d->q0.ptr, q0, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->v0.ptr, v0, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_LAUNCH( ctx, "rb2d_flow", double( d->n ) * 120.0, k_rb2d_flow<<<sg_div_up( d->n, 256 ), 256, 0, ctx->stream>>>( map_kind, d->n, d->q0.as<double>(), d->v0.as<double>(), d->M.as<double>(), d->btype.as<uint32_t>(),
             d->g[0], d->g[1], dt, d->q1.as<double>(), d->v1.as<double>() ) );
  SG_CUDA( ctx, cudaMemcpyAsync( q1, d->q1.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( v1, d->v1.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  d->flow_resident = true;
  d->q1_valid = true;
  return SG_OK;
}

int sg_rb2d_active_set( sg_ctx* ctx, const double* q0, const double* q1, uint32_t out_flags, sg_contacts* out )
{
  if( ctx == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  Rb2dData* d = rb2d_data( ctx );
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( ( out_flags & SG_IN_RESIDENT ) != 0u )
  {
    // (q0, q1) are the input and output of the last sg_rb2d_flow on this context: still on the device, nothing is uploaded
    if( !d->flow_resident ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_active_set: SG_IN_RESIDENT without a preceding sg_rb2d_flow on this context" ); }
  }
  else if( d->n > 0 )
  {
    if( q0 == nullptr || q1 == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_active_set: null vector" ); }
    SG_CUDA( ctx, cudaMemcpyAsync( d->q0.ptr, q0, size_t( d->n ) * 24, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->q1.ptr, q1, size_t( d->n ) * 24, cudaMemcpyHostToDevice, ctx->stream ) );
    d->flow_resident = false;
    d->q1_valid = false; // q1 is the caller's, v1 whatever an earlier flow left
  }
  const int rc = ( d->px != nullptr && d->px->portals.n > 0u ) ? rb2d_portal_active_set_device( ctx, d ) : rb2d_active_set_device( ctx, d );
  if( rc != SG_OK ) { d->have_result = false; } // a failed call leaves nothing to fetch (partial or stale lists)
  if( rc != SG_OK ) { return rc; }
  return rb2d_copy_out( ctx, d, out_flags, out );
}

int sg_rb2d_upload( sg_ctx* ctx, const double* q, const double* v )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb2dData* d = rb2d_data( ctx );
  d->flow_resident = false;
  d->q1_valid = false;
  if( d->n == 0 ) { return SG_OK; }
  if( q == nullptr || v == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_upload: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->q0.ptr, q, size_t( d->n ) * 24, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->v0.ptr, v, size_t( d->n ) * 24, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_rb2d_step( sg_ctx* ctx, int map_kind, double dt, sg_contacts* out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( map_kind != SG_MAP_SYMPLECTIC_EULER && map_kind != SG_MAP_VERLET ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_step: map kind %d is not a rigidbody2d map", map_kind ); }
  Rb2dData* d = rb2d_data( ctx );
  d->flow_resident = false; // q1 is about to be overwritten by the resident step
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( d->n > 0 )
  {
    SG_LAUNCH( ctx, "rb2d_flow", double( d->n ) * 120.0, k_rb2d_flow<<<sg_div_up( d->n, 256 ), 256, 0, ctx->stream>>>( map_kind, d->n, d->q0.as<double>(), d->v0.as<double>(), d->M.as<double>(), d->btype.as<uint32_t>(),
               d->g[0], d->g[1], dt, d->q1.as<double>(), d->v1.as<double>() ) );
  }
  d->q1_valid = true;
  const int rc = ( d->px != nullptr && d->px->portals.n > 0u ) ? rb2d_portal_active_set_device( ctx, d ) : rb2d_active_set_device( ctx, d );
  if( rc != SG_OK ) { d->have_result = false; } // a failed call leaves nothing to fetch (partial or stale lists)
  if( rc != SG_OK ) { return rc; }
  if( out != nullptr )
  {
    memset( out, 0, sizeof( *out ) );
    out->dim = 2;
    out->n_candidates = d->n_cand;
    out->n_body_body = d->n_bb;
    out->n_plane = d->n_static;
    out->n_active = d->n_bb + d->n_static;
  }
  return SG_OK;
}

int sg_rb2d_fetch( sg_ctx* ctx, uint32_t out_flags, double* q1, double* v1, sg_contacts* out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb2dData* d = rb2d_data( ctx );
  if( !d->have_result ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_fetch: no step has been run" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( q1 != nullptr && d->n > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( q1, d->q1.ptr, size_t( d->n ) * 24, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  if( v1 != nullptr && d->n > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( v1, d->v1.ptr, size_t( d->n ) * 24, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  if( out != nullptr ) { return rb2d_copy_out( ctx, d, out_flags, out ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

// ---- state I/O at the seam: RigidBody2DState's binary snapshot (rigidbody2d/RigidBody2DState.cpp:485-556), sg_rb2d_snapshot.h ----
int sg_rb2d_state_serialize( sg_ctx* ctx, int which, void* buf, uint64_t cap, uint64_t* bytes )
{
  if( ctx == nullptr || bytes == nullptr || ( which != 0 && which != 1 ) ) { return SG_ERR_INVALID; }
  Rb2dData* d = rb2d_data( ctx );
  if( which == 1 && !d->q1_valid && d->n > 0 ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_state_serialize: which = 1 without a preceding sg_rb2d_flow / sg_rb2d_step on this context" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const uint32_t n = d->n;
  sg_snapshot::Rb2dState s;
  s.n = n;
  s.q.resize( size_t( 3 ) * n ); s.v.resize( size_t( 3 ) * n ); s.M.resize( size_t( 3 ) * n );
  if( n > 0 )
  {
    const size_t nb = size_t( n ) * 24;
    SG_CUDA( ctx, cudaMemcpyAsync( s.q.data(), ( which == 0 ) ? d->q0.ptr : d->q1.ptr, nb, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( s.v.data(), ( which == 0 ) ? d->v0.ptr : d->v1.ptr, nb, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( s.M.data(), d->M.ptr, nb, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  }
  s.fixed = d->h_fixed; s.geo_of_body = d->h_geo_of_body;
  s.geo_type.resize( d->geo_type.size() );
  for( size_t k = 0; k < d->geo_type.size(); ++k ) { s.geo_type[k] = ( d->geo_type[k] == SG_GEO2_CIRCLE ) ? 0u : ( d->geo_type[k] == SG_GEO2_BOX ) ? 1u : 2u; }
  s.geo_r = d->geo_r; s.geo_half = d->geo_half;
  s.g[0] = d->g[0]; s.g[1] = d->g[1];
  for( uint32_t p = 0; p < d->planes.n; ++p )
  {
    // RigidBody2DStaticPlane( x, n ): m_t = ( -n.y, n.x ) (rigidbody2d/RigidBody2DStaticPlane.cpp:10-14)
    s.plane_x.push_back( d->planes.x[p][0] ); s.plane_x.push_back( d->planes.x[p][1] );
    s.plane_n.push_back( d->planes.nrm[p][0] ); s.plane_n.push_back( d->planes.nrm[p][1] );
    s.plane_t.push_back( -d->planes.nrm[p][1] ); s.plane_t.push_back( d->planes.nrm[p][0] );
  }
  if( d->px != nullptr )
  {
    for( uint32_t p = 0; p < d->px->portals.n; ++p )
    {
      const SgPortal2D& pt = d->px->portals.p[p];
      for( int k = 0; k < 2; ++k )
      {
        s.portal_ax.push_back( pt.ax[k] ); s.portal_an.push_back( pt.an[k] ); s.portal_at.push_back( pt.at[k] );
        s.portal_bx.push_back( pt.bx[k] ); s.portal_bn.push_back( pt.bn[k] ); s.portal_bt.push_back( pt.bt[k] );
      }
      s.portal_v.push_back( pt.v ); s.portal_bounds.push_back( pt.bounds ); s.portal_dx.push_back( pt.dx );
    }
  }
  sg_snapshot::Sink out{ static_cast<unsigned char*>( buf ), cap, 0 };
  if( !sg_snapshot::serialize( s, out ) ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb2d_state_serialize: a geometry that is neither circle nor box" ); }
  *bytes = out.n;
  if( buf != nullptr && out.n > cap ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_state_serialize: buffer of %llu bytes, %llu needed", ( unsigned long long )( cap ), ( unsigned long long )( out.n ) ); }
  return SG_OK;
}

// RigidBody2DState::deserialize (rigidbody2d/RigidBody2DState.cpp:542-556): configures the context from a snapshot and uploads ( q, v )
int sg_rb2d_state_deserialize( sg_ctx* ctx, const void* buf, uint64_t bytes )
{
  if( ctx == nullptr || buf == nullptr ) { return SG_ERR_INVALID; }
  sg_snapshot::Source in{ static_cast<const unsigned char*>( buf ), bytes, 0, true };
  sg_snapshot::Rb2dState s;
  const char* why = "";
  const int prc = sg_snapshot::parse( in, s, &why );
  if( prc == 1 ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_state_deserialize: %s", why ); }
  if( prc == 2 ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb2d_state_deserialize: %s", why ); }
  const uint32_t npl = uint32_t( s.plane_x.size() / 2 ), npo = uint32_t( s.portal_v.size() );
  if( npl > SG_MAX_PLANES || npo > SG_MAX_PORTALS ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb2d_state_deserialize: more planes or portals than this library holds" ); }
  // the kernels take a plane's tangent as ( -n.y, n.x ), which is what the reference's constructor stores: a snapshot with another frame is not this path's
  for( uint32_t p = 0; p < npl; ++p )
  {
    if( s.plane_t[2 * p] != -s.plane_n[2 * p + 1] || s.plane_t[2 * p + 1] != s.plane_n[2 * p] ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb2d_state_deserialize: plane %u has a tangent other than ( -n.y, n.x )", p ); }
  }
  for( uint32_t p = 0; p < npo; ++p ) { if( !( s.portal_bounds[p] >= 0.0 ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb2d_state_deserialize: portal %u has negative bounds", p ); } }
  const uint32_t ngeo = uint32_t( s.geo_type.size() );
  std::vector<uint32_t> type( ngeo );
  for( uint32_t k = 0; k < ngeo; ++k ) { type[k] = ( s.geo_type[k] == 0u ) ? uint32_t( SG_GEO2_CIRCLE ) : uint32_t( SG_GEO2_BOX ); }
  int rc = sg_rb2d_set_geometry( ctx, ngeo, type.data(), s.geo_r.data(), s.geo_half.data() );
  if( rc != SG_OK ) { return rc; }
  rc = sg_rb2d_set_bodies( ctx, s.n, s.geo_of_body.data(), s.fixed.data(), s.M.data() );
  if( rc != SG_OK ) { return rc; }
  Rb2dData* d = rb2d_data( ctx );
  d->g[0] = s.g[0]; d->g[1] = s.g[1];
  rc = sg_rb2d_set_planes( ctx, npl, s.plane_x.data(), s.plane_n.data() ); // RigidBody2DStaticPlane( std::istream& ) reads x and n back as stored
  if( rc != SG_OK ) { return rc; }
  if( npo > 0 || d->px != nullptr )
  {
    rc = sg_rb2d_set_portals( ctx, npo, s.portal_ax.data(), s.portal_an.data(), s.portal_bx.data(), s.portal_bn.data(), s.portal_v.data(), s.portal_bounds.data() );
    if( rc != SG_OK ) { return rc; }
    // the frames and the Lees-Edwards offset stay as stored (PlanarPortal( std::istream& ), rigidbody2d/PlanarPortal.cpp:101-109)
    for( uint32_t p = 0; p < npo; ++p )
    {
      SgPortal2D& pt = d->px->portals.p[p];
      for( int k = 0; k < 2; ++k ) { pt.at[k] = s.portal_at[2 * p + k]; pt.bt[k] = s.portal_bt[2 * p + k]; }
      pt.dx = s.portal_dx[p];
    }
  }
  if( s.n == 0 ) { return SG_OK; }
  return sg_rb2d_upload( ctx, s.q.data(), s.v.data() );
}

}
