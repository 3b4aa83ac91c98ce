// sg_rb3d.cu -- rigidbody3d hot path: SplitHam / DMV unconstrained flow, 3-D broad phase, sphere-sphere / box-box /
// mesh-vs-SDF narrow phase, body-plane tests.
//
// Reference behaviour reproduced (file:line in the SCISim checkout):
//   rigidbody3d/UnconstrainedMaps/SplitHamMap.cpp:17-182, DMVMap.cpp:15-207            k_rb3d_flow
//   rigidbody3d/Forces/NearEarthGravityForce.cpp:39-53, RigidBody3DState.cpp:70-240     gravity, M0, R I0 R^T
//   rigidbody3d/RigidBody3DSim.cpp:1057-1069 + Geometry/*::computeAABB                  k_rb3d_aabb (AABBs at q1 only)
//   rigidbody3d/RigidBody3DSim.cpp:1072-1141                                            sg_broadphase.cuh pipeline
//   rigidbody3d/RigidBody3DSim.cpp:879-962 (dispatch, kinematic rules), :793-819 (sphere-sphere), :665-694 (box-box),
//   :844-876 + Constraints/MeshMeshUtilities.cpp:10-65 + Geometry/RigidBodyTriangleMesh.cpp:276-335 (mesh-mesh)
//   rigidbody3d/RigidBody3DSim.cpp:1414-1502 (body-plane), Constraints/StaticPlane{Sphere,Box,Body}Constraint.cpp
//
// Two pipelines share the broad phase:
//   * all bodies are spheres: records carry (x1, x0, r) and the sphere test is fused into the pair kernels
//     (Sphere3DPolicy), exactly like the ball2d path;
//   * otherwise: boxes are built by k_rb3d_aabb, the broad phase emits the sorted candidate list, and the narrow
//     phase runs over that list (one thread per sphere/box pair, one CTA per mesh pair), count -> scan -> emit so the
//     contacts of a pair stay in the order the reference's routines produce them.
#include "sg_boxbox.cuh"
#include "sg_broadphase.cuh"
#include "sg_slab.cuh"
#include "sg_rb3d_snapshot.h"

#include <cstdlib>
#include <cuda.h> // CUtensorMap (types only; the encoder is fetched at run time through cudaGetDriverEntryPoint)

#define SG_FIXED_BIT 0x80000000u
#define SG_GHOST_BIT3 0x40000000u // slab mode: the record is a halo copy of a body another rank owns

// ---- contact output (SoA, reference order) ---------------------------------------------------------
struct ContactOut3D
{
  uint32_t* type;
  uint32_t* i;
  uint32_t* j;
  uint32_t* aux;
  double* n;     // 3 per contact
  double* p;     // 3 per contact
  double* depth;
  unsigned long long cap;
  GidMap gid; // multi-GPU slabs: local body slot -> global body index (identity on one GPU)
};

__device__ __forceinline__ void put_contact( const ContactOut3D& out, const unsigned long long k, const uint32_t type, const uint32_t i, const uint32_t j, const uint32_t aux,
                                             const V3d n, const V3d p, const double depth )
{
  if( k >= out.cap ) { return; }
  // j is a body for the body-body types, a plane / cylinder number otherwise
  const bool j_is_body = type <= SG_KINEMATIC_BODY_BODY || type == SG_SPHERE_SPHERE_TELEPORTED || type == SG_KINEMATIC_OBJECT_SPHERE_TELEPORTED;
  out.type[k] = type; out.i[k] = out.gid( i ); out.j[k] = j_is_body ? out.gid( j ) : j; out.aux[k] = aux;
  out.n[3 * k] = n.x; out.n[3 * k + 1] = n.y; out.n[3 * k + 2] = n.z;
  out.p[3 * k] = p.x; out.p[3 * k + 1] = p.y; out.p[3 * k + 2] = p.z;
  out.depth[k] = depth;
}

__device__ __forceinline__ double sg_nan() { return __longlong_as_double( 0x7ff8000000000000LL ); }

__device__ __forceinline__ V3d load_v3( const double* __restrict__ a, const size_t b ) { return v3( __ldg( a + 3 * b ), __ldg( a + 3 * b + 1 ), __ldg( a + 3 * b + 2 ) ); }
__device__ __forceinline__ M3d load_m3( const double* __restrict__ a, const size_t b )
{
  M3d R;
  #pragma unroll
  for( int k = 0; k < 9; ++k ) { R.m[k] = __ldg( a + 9 * b + k ); }
  return R;
}

// ---- sphere-sphere, shared by both pipelines (RigidBody3DSim.cpp:793-819, dispatch :879-893) --------
// a has the lower body index.  Returns false when the pair is skipped (both kinematic) or not touching at q1.
__device__ __forceinline__ bool sphere_pair_active( const V3d x1a, const double ra, const bool fa, const V3d x1b, const double rb, const bool fb )
{
  if( fa && fb ) { return false; }
  const V3d d = x1a - x1b;
  return dot3( d, d ) <= ( ra + rb ) * ( ra + rb );
}
__device__ __forceinline__ void sphere_pair_emit( const ContactOut3D& out, const unsigned long long k, uint32_t ia, V3d x0a, V3d x1a, double ra, bool fa,
                                                  uint32_t ib, V3d x0b, V3d x1b, double rb, bool fb )
{
  // the kinematic body is always listed second
  if( fa )
  {
    const uint32_t ti = ia; ia = ib; ib = ti;
    V3d tv = x0a; x0a = x0b; x0b = tv;
    tv = x1a; x1a = x1b; x1b = tv;
    const double tr = ra; ra = rb; rb = tr;
    fb = true;
  }
  const V3d n = normalized3( x0a - x0b );
  const V3d d1 = x1a - x1b;
  const double depth = fmin( 0.0, sqrt( dot3( d1, d1 ) ) - ra - rb );
  if( !fb )
  {
    const V3d p = x0a + ( ra / ( ra + rb ) ) * ( x0b - x0a );
    put_contact( out, k, SG_SPHERE_SPHERE, ia, ib, 0u, n, p, depth );
  }
  else
  {
    put_contact( out, k, SG_KINEMATIC_SPHERE_SPHERE, ia, ib, 0u, n, x0b, depth );
  }
}

// ---- fused all-spheres pipeline --------------------------------------------------------------------
struct Sphere3DIn
{
  const double* q0;       // 12N
  const double* q1;
  const double* r;        // per body radius
  const uint32_t* flags;  // per body: SG_FIXED_BIT or 0
  uint32_t n;
  // slab mode (multi-GPU): slots [0, n_owned) are this rank's bodies, [n_owned, n_owned + ghost_cap) / the next ghost_cap slots hold the
  // halo copies from the lower / higher neighbour, ghost_counts[side] of them in use this step.  ghost_counts == nullptr: every slot is a body
  uint32_t n_owned;
  uint32_t ghost_cap;
  const uint32_t* ghost_counts;
};

// Everything pass 1 needs of a PARTNER -- the box at q1 (x1 -/+ r), the sphere test at q1, "kinematically scripted" -- sits in the
// first 32 bytes (one sector, two 128-bit loads instead of four; a lattice sphere owns 13 of its 26 neighbours, and those record
// reads through L1 are what bounds the 3-D pass 1): the radius carries the scripted flag in its sign.
struct alignas( 64 ) Sphere3DRec
{
  double x1[3];
  double r;      // < 0 (sign bit set): kinematically scripted; the radius is fabs( r )
  double x0[3];
  uint32_t idx;  // bit 31: kinematically scripted
  uint32_t key;
};
__device__ __forceinline__ bool sphere_rec_fixed( const double r ) { return __double_as_longlong( r ) < 0; }

struct Sphere3DPolicy
{
  static constexpr int D = 3;
  static constexpr bool HAS_NARROW = true;
  static constexpr double IN_BYTES = 32.0;
  static constexpr uint32_t IDX_OFFSET = 56u;
  static constexpr uint32_t ORD_OFFSET = IDX_OFFSET;
  static constexpr bool ORD_IN_REC = false; // the record is full: a body's order word (its global index in slab mode) lives in the dense sidx array only
  static constexpr uint32_t IDX_MASK = 0x3fffffffu;
  using In = Sphere3DIn;
  using Rec = Sphere3DRec;
  using Out = ContactOut3D;
  // RigidBodySphere::computeAABB at q1: cm -/+ r
  __device__ static void load_aabb( const In& in, const uint32_t i, double* lo, double* hi )
  {
    const double r = __ldg( &in.r[i] );
    #pragma unroll
    for( int k = 0; k < 3; ++k ) { const double c = __ldg( &in.q1[3 * size_t( i ) + k] ); lo[k] = c - r; hi[k] = c + r; }
  }
  __device__ static Rec make_rec( const In& in, const uint32_t i, const uint32_t key, const uint32_t, const uint32_t )
  {
    Rec rec;
    #pragma unroll
    for( int k = 0; k < 3; ++k ) { rec.x1[k] = __ldg( &in.q1[3 * size_t( i ) + k] ); rec.x0[k] = __ldg( &in.q0[3 * size_t( i ) + k] ); }
    const uint32_t fl = __ldg( &in.flags[i] );
    const double rad = __ldg( &in.r[i] );
    rec.r = ( fl & SG_FIXED_BIT ) ? -rad : rad;
    rec.idx = i | fl | ( ( in.ghost_counts != nullptr && i >= in.n_owned ) ? SG_GHOST_BIT3 : 0u );
    rec.key = key;
    return rec;
  }
  __device__ static void rec_aabb( const Rec& s, double* lo, double* hi )
  {
    const double rad = fabs( s.r );
    #pragma unroll
    for( int k = 0; k < 3; ++k ) { lo[k] = s.x1[k] - rad; hi[k] = s.x1[k] + rad; }
  }
  __device__ static uint32_t rec_idx( const Rec& s ) { return s.idx & IDX_MASK; }
  __device__ static uint32_t rec_idx_raw( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_ord( const Rec& s ) { return rec_idx( s ); }
  __device__ static uint32_t rec_ord_raw( const Rec& s ) { return s.idx; }
  __device__ static uint32_t rec_key( const Rec& s ) { return s.key; }
  __device__ static bool owns( const Rec& s ) { return ( s.idx & SG_GHOST_BIT3 ) == 0u; }
  __device__ static bool valid( const In& in, const uint32_t i )
  {
    if( in.ghost_counts == nullptr || i < in.n_owned ) { return true; }
    const uint32_t k = i - in.n_owned;
    return ( k < in.ghost_cap ) ? ( k < __ldg( &in.ghost_counts[0] ) ) : ( k - in.ghost_cap < __ldg( &in.ghost_counts[1] ) );
  }
  __device__ static uint32_t rec_c1( const Rec& s, const GridParams& g ) { return ( s.key / g.dims[0] ) % g.dims[1]; }
  __device__ static uint32_t rec_c2( const Rec& s, const GridParams& g ) { return s.key / ( g.dims[0] * g.dims[1] ); }
  // a partner as pass 1 needs it: the first 32 bytes of its record (box at q1, sphere test, scripted flag)
  __device__ static Rec load_pass1( const Rec* __restrict__ p )
  {
    union { Rec r; int4 v[4]; } u;
    const int4* src = reinterpret_cast<const int4*>( p );
    u.v[0] = __ldg( src ); u.v[1] = __ldg( src + 1 ); u.v[2] = make_int4( 0, 0, 0, 0 ); u.v[3] = make_int4( 0, 0, 0, 0 );
    return u.r;
  }
  __device__ static bool narrow_test( const Rec& a, const Rec& b )
  {
    return sphere_pair_active( v3( a.x1[0], a.x1[1], a.x1[2] ), fabs( a.r ), sphere_rec_fixed( a.r ), v3( b.x1[0], b.x1[1], b.x1[2] ), fabs( b.r ), sphere_rec_fixed( b.r ) ); // first 32 bytes only
  }
  __device__ static void contact_emit( const Out& out, unsigned long long& k, const Rec& a, const Rec& b )
  {
    sphere_pair_emit( out, k, a.idx & IDX_MASK, v3( a.x0[0], a.x0[1], a.x0[2] ), v3( a.x1[0], a.x1[1], a.x1[2] ), fabs( a.r ), ( a.idx & SG_FIXED_BIT ) != 0u,
                      b.idx & IDX_MASK, v3( b.x0[0], b.x0[1], b.x0[2] ), v3( b.x1[0], b.x1[1], b.x1[2] ), fabs( b.r ), ( b.idx & SG_FIXED_BIT ) != 0u );
    ++k;
  }
};

// ---- generic pipeline: boxes from k_rb3d_aabb --------------------------------------------------------
struct Box3DIn
{
  const double* boxes; // n * 6: lo(3), hi(3)
  uint32_t n;
};
struct alignas( 64 ) Box3DRec { double lo[3]; double hi[3]; uint32_t idx; uint32_t key; uint32_t c1, c2; };
struct NoOut3D {};
struct Box3DPolicy
{
  static constexpr int D = 3;
  static constexpr bool HAS_NARROW = false;
  static constexpr double IN_BYTES = 48.0;
  static constexpr uint32_t IDX_OFFSET = 48u;
  static constexpr uint32_t ORD_OFFSET = IDX_OFFSET; // bodies are ranked by their index
  static constexpr bool ORD_IN_REC = true;
  static constexpr uint32_t IDX_MASK = 0xffffffffu;
  using In = Box3DIn;
  using Rec = Box3DRec;
  using Out = NoOut3D;
  __device__ static void load_aabb( const In& in, const uint32_t i, double* lo, double* hi )
  {
    const double* b = in.boxes + size_t( i ) * 6;
    #pragma unroll
    for( int k = 0; k < 3; ++k ) { lo[k] = __ldg( b + k ); hi[k] = __ldg( b + 3 + k ); }
  }
  __device__ static Rec make_rec( const In& in, const uint32_t i, const uint32_t key, const uint32_t c1, const uint32_t c2 )
  {
    Rec r;
    load_aabb( in, i, r.lo, r.hi );
    r.idx = i; r.key = key; r.c1 = c1; r.c2 = c2;
    return r;
  }
  __device__ static void rec_aabb( const Rec& s, double* lo, double* hi )
  {
    #pragma unroll
    for( int k = 0; k < 3; ++k ) { lo[k] = s.lo[k]; hi[k] = s.hi[k]; }
  }
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

// ---- device-side scene description -----------------------------------------------------------------
struct alignas( 128 ) MeshDev // the descriptor is fetched from global memory by the TMA unit: keep every array element's copy 128-byte aligned
{
  CUtensorMap tmap;      // 3-D tiled map over the (row-padded) distance grid, box = SDF_BX x SDF_BY x SDF_BZ doubles
  const double* verts;   uint32_t nverts;
  const double* samples; uint32_t nsamples;
  const double* hull;    uint32_t nhull;
  const double* sdf;     // x fastest, rows padded to `pitch` doubles (even, so the TMA row stride is a multiple of 16 B)
  double delta[3];
  double origin[3];
  double grid_end[3];
  uint32_t dims[3];
  uint32_t pitch;
  uint32_t has_tmap;
};

// brick tile moved by one TMA tensor copy, and how many of them a CTA can hold (128 KB of shared memory)
#define SDF_BX 16
#define SDF_BY 8
#define SDF_BZ 4
#define SDF_TILE_ELEMS ( SDF_BX * SDF_BY * SDF_BZ )
#define SDF_TILES_MAX 48
#define SDF_TILES_DEFAULT 12

struct Rb3dDev
{
  uint32_t n;
  uint32_t n_live;         // bodies the body-plane loops visit: n, or this rank's own bodies in slab mode (the ghosts that follow them are not its to test)
  const uint32_t* btype;   // per body: geometry type | SG_FIXED_BIT
  const double* bparam;    // per body 4 doubles: sphere (r,-,-,-), box (hx,hy,hz,-)
  const uint32_t* bmesh;   // per body: mesh index (meshes only)
  const MeshDev* meshes;
  unsigned long long* mesh_stats; // [0] sample sweeps served from a TMA-staged brick, [1] sweeps that read the grid directly
};

#define SG_MAX_CYLINDERS 8
// static geometry: planes, then cylinders (the order RigidBody3DSim::computeActiveSet visits them, RigidBody3DSim.cpp:259-261)
struct Planes3D
{
  uint32_t n;
  uint32_t ncyl;
  double x[SG_MAX_PLANES][3];
  double nrm[SG_MAX_PLANES][3];
  double cx[SG_MAX_CYLINDERS][3];  // point on the axis
  double cax[SG_MAX_CYLINDERS][3]; // unit axis
  double cr[SG_MAX_CYLINDERS];
};

// ---- unconstrained flow ----------------------------------------------------------------------------
// AngleAxis( angle, axis ).toRotationMatrix() (Eigen/src/Geometry/AngleAxis.h)
__device__ inline M3d angle_axis_matrix( const double angle, const V3d axis )
{
  M3d res;
  double sn, cs;
  sincos( angle, &sn, &cs );
  const V3d sin_axis = sn * axis;
  const V3d cos1_axis = ( 1.0 - cs ) * axis;
  double tmp;
  tmp = cos1_axis.x * axis.y;
  res.m[1] = tmp - sin_axis.z; res.m[3] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z;
  res.m[2] = tmp + sin_axis.y; res.m[6] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z;
  res.m[5] = tmp - sin_axis.x; res.m[7] = tmp + sin_axis.x;
  res.m[0] = cos1_axis.x * axis.x + cs;
  res.m[4] = cos1_axis.y * axis.y + cs;
  res.m[8] = cos1_axis.z * axis.z + cs;
  return res;
}

// DMVMap.cpp:15-60 + 65-100 on one body: R1 from R0 and the body-frame angular momentum
__device__ inline M3d dmv_rotation( const M3d& R0, const V3d am, const double h, const V3d I0 )
{
  // Eigen::Quaternion( Matrix3 ) -- Shoemake
  double qw, qc[3];
  #define SGM( r, c ) R0.m[3 * ( r ) + ( c )]
  double t = SGM( 0, 0 ) + ( SGM( 1, 1 ) + SGM( 2, 2 ) );
  if( t > 0.0 )
  {
    t = sqrt( t + 1.0 );
    qw = 0.5 * t;
    t = 0.5 / t;
    qc[0] = ( SGM( 2, 1 ) - SGM( 1, 2 ) ) * t;
    qc[1] = ( SGM( 0, 2 ) - SGM( 2, 0 ) ) * t;
    qc[2] = ( SGM( 1, 0 ) - SGM( 0, 1 ) ) * t;
  }
  else
  {
    int i = 0;
    if( SGM( 1, 1 ) > SGM( 0, 0 ) ) { i = 1; }
    if( SGM( 2, 2 ) > SGM( i, i ) ) { i = 2; }
    const int j = ( i + 1 ) % 3;
    const int k = ( j + 1 ) % 3;
    t = sqrt( SGM( i, i ) - SGM( j, j ) - SGM( k, k ) + 1.0 );
    qc[i] = 0.5 * t;
    t = 0.5 / t;
    qw = ( SGM( k, j ) - SGM( j, k ) ) * t;
    qc[j] = ( SGM( j, i ) + SGM( i, j ) ) * t;
    qc[k] = ( SGM( k, i ) + SGM( i, k ) ) * t;
  }
  #undef SGM
  const double eps = fabs( h * 1e-15 );
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
    const double err = fabs( cm1b - cm1 ) + fabs( cm2b - cm2 ) + fabs( cm3b - cm3 );
    if( err <= eps ) { break; }
  }
  const double q0 = qw, q1 = qc[0], q2 = qc[1], q3 = qc[2];
  double w = q0 - cm1 * q1 - cm2 * q2 - cm3 * q3;
  double x = q1 + cm1 * q0 + cm3 * q2 - cm2 * q3;
  double y = q2 + cm2 * q0 + cm1 * q3 - cm3 * q1;
  double z = q3 + cm3 * q0 + cm2 * q1 - cm1 * q2;
  const double nrm = sqrt( ( x * x + z * z ) + ( y * y + w * w ) ); // Eigen 3.3.4 sums the squared ( x, y, z, w ) coefficients by packets (Core/Redux.h: redux_vec_unroller + predux): ( x2 + z2 ) + ( y2 + w2 )
  x /= nrm; y /= nrm; z /= nrm; w /= nrm;
  // Quaternion::toRotationMatrix
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M3d r;
  r.m[0] = 1.0 - ( tyy + tzz ); r.m[1] = txy - twz; r.m[2] = txz + twy;
  r.m[3] = txy + twz; r.m[4] = 1.0 - ( txx + tzz ); r.m[5] = tyz - twx;
  r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1.0 - ( txx + tyy );
  return r;
}

// ExponentialEulerMap::flow on one body (ExponentialEulerMap.cpp:34-91): explicit Euler on x and on the columns of R, then
// projectOrientation (:13-32) = the orthogonal factor of R's polar decomposition (the reference: U V^T from Eigen::JacobiSVD).
// Here by Newton's iteration X <- ( X + X^-T ) / 2, which converges quadratically to that same factor; the input is within
// O( dt^2 ) of a rotation, so a handful of iterations reach rounding level (tests hold it to 1e-12 of the oracle's Jacobi SVD).
__device__ inline void rb3d_flow_exponential_euler( const uint32_t b, const size_t nb, const V3d x0, const M3d& R0, const V3d vl, const V3d w, const double m, const double gx, const double gy, const double gz,
                                                    const double dt, double* __restrict__ q1, double* __restrict__ v1 )
{
  double* x1o = q1 + 3 * size_t( b );
  double* R1o = q1 + 3 * nb + 9 * size_t( b );
  double* vlo = v1 + 3 * size_t( b );
  double* vao = v1 + 3 * nb + 3 * size_t( b );
  x1o[0] = x0.x + dt * vl.x; x1o[1] = x0.y + dt * vl.y; x1o[2] = x0.z + dt * vl.z;
  double X[9];
  #pragma unroll
  for( int j = 0; j < 3; ++j )
  {
    const double cx = R0.m[j], cy = R0.m[3 + j], cz = R0.m[6 + j];
    X[j] = cx + dt * ( w.y * cz - w.z * cy );
    X[3 + j] = cy + dt * ( w.z * cx - w.x * cz );
    X[6 + j] = cz + dt * ( w.x * cy - w.y * cx );
  }
  for( int it = 0; it < 40; ++it )
  {
    // inverse transpose = cofactor matrix / determinant
    double C[9];
    C[0] = X[4] * X[8] - X[5] * X[7]; C[1] = X[5] * X[6] - X[3] * X[8]; C[2] = X[3] * X[7] - X[4] * X[6];
    C[3] = X[2] * X[7] - X[1] * X[8]; C[4] = X[0] * X[8] - X[2] * X[6]; C[5] = X[1] * X[6] - X[0] * X[7];
    C[6] = X[1] * X[5] - X[2] * X[4]; C[7] = X[2] * X[3] - X[0] * X[5]; C[8] = X[0] * X[4] - X[1] * X[3];
    const double det = X[0] * C[0] + X[1] * C[1] + X[2] * C[2];
    double change = 0.0, size = 0.0;
    #pragma unroll
    for( int k = 0; k < 9; ++k )
    {
      const double y = 0.5 * ( X[k] + C[k] / det );
      change += ( y - X[k] ) * ( y - X[k] );
      size += y * y;
      X[k] = y;
    }
    if( !( change > 1.0e-32 * size ) ) { break; }
  }
  #pragma unroll
  for( int k = 0; k < 9; ++k ) { R1o[k] = X[k]; }
  // A = Minv * F: linear 0 + ( 1 / m ) * ( 0 + m g ); angular: the inverse-inertia block times a zero torque
  const double minv = 1.0 / m;
  const double Ax = 0.0 + minv * ( 0.0 + m * gx ), Ay = 0.0 + minv * ( 0.0 + m * gy ), Az = 0.0 + minv * ( 0.0 + m * gz );
  vlo[0] = vl.x + dt * Ax; vlo[1] = vl.y + dt * Ay; vlo[2] = vl.z + dt * Az;
  vao[0] = w.x + dt * 0.0; vao[1] = w.y + dt * 0.0; vao[2] = w.z + dt * 0.0;
}

// One axis rotation of SplitHamMap (SplitHamMap.cpp:121-176): R1 <- R1 * AngleAxis( -angle, -e_K ), pB <- AngleAxis( angle, -e_K ) * pB.
// The reference forms each rotation with AngleAxis::toRotationMatrix on the axis -e_K = ( -0, -0, -1 ) etc. and multiplies full 3x3
// matrices.  Worked through with those axis components, toRotationMatrix( a, -e_z ) is exactly
//     [  cs  sn  0 ]                                              [ cs   0  -sn ]                 [ e    0    0 ]
//     [ -sn  cs  0 ]   with e = fl( fl( 1 - cs ) + cs ),   -e_y:  [  0   e   0  ]         -e_x:   [ 0   cs   sn ]
//     [  0   0   e ]                                              [ sn   0   cs ]                 [ 0  -sn   cs ]
// (every other entry a signed zero), and in a product ( a0 b0 + a1 b1 ) + a2 b2 a term with a zero factor adds nothing: what is left
// per entry is the same one or two roundings the full product makes.  So the structured form below gives the reference's values bit
// for bit (signs of exact zeros aside) with a third of the arithmetic -- and one sincos per rotation instead of two, sin( -a ) = -sin( a ).
// This kernel is bound by FP64 issue (61 % issue-active at 24 % occupancy, DRAM 36 %; profiles/ncu_r2_c4.md), not by its loads.
template<int K>
__device__ __forceinline__ void splitham_axis_step( M3d& R, V3d& p, const double angle )
{
  double sn, cs;
  sincos( angle, &sn, &cs );
  const double e = ( 1.0 - cs ) + cs;
  #pragma unroll
  for( int r = 0; r < 3; ++r )
  {
    const double a0 = R.m[3 * r], a1 = R.m[3 * r + 1], a2 = R.m[3 * r + 2];
    if( K == 2 ) { R.m[3 * r] = a0 * cs + a1 * sn; R.m[3 * r + 1] = a0 * ( -sn ) + a1 * cs; R.m[3 * r + 2] = a2 * e; }       // B = [ cs -sn 0 | sn cs 0 | 0 0 e ]
    else if( K == 1 ) { R.m[3 * r] = a0 * cs + a2 * ( -sn ); R.m[3 * r + 1] = a1 * e; R.m[3 * r + 2] = a0 * sn + a2 * cs; }  // B = [ cs 0 sn | 0 e 0 | -sn 0 cs ]
    else { R.m[3 * r] = a0 * e; R.m[3 * r + 1] = a1 * cs + a2 * sn; R.m[3 * r + 2] = a1 * ( -sn ) + a2 * cs; }               // B = [ e 0 0 | 0 cs -sn | 0 sn cs ]
  }
  const double px = p.x, py = p.y, pz = p.z;
  if( K == 2 ) { p = v3( cs * px + sn * py, ( -sn ) * px + cs * py, e * pz ); }
  else if( K == 1 ) { p = v3( cs * px + ( -sn ) * pz, e * py, sn * px + cs * pz ); }
  else { p = v3( e * px, cs * py + sn * pz, ( -sn ) * py + cs * pz ); }
}

__global__ void __launch_bounds__( 128 ) k_rb3d_flow( const int kind_flags, const uint32_t n, const uint32_t nrun, const double* __restrict__ q0, const double* __restrict__ v0, const double* __restrict__ mass,
                                                     const double* __restrict__ I0, const uint32_t* __restrict__ btype, const double gx, const double gy, const double gz, const double dt,
                                                     double* __restrict__ q1, double* __restrict__ v1 )
{
  // n = slots (the stride of the [3n | 9n] / [3n | 3n] layouts), nrun = bodies to integrate (the leading ones)
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= nrun ) { return; }
  const int kind = kind_flags & 0xff;
  const size_t nb = n;
  const double m = __ldg( &mass[b] );
  const V3d x0 = load_v3( q0, b );
  const M3d R0 = load_m3( q0 + 3 * nb, b );
  const V3d vl = load_v3( v0, b );
  const V3d w0 = load_v3( v0 + 3 * nb, b );
  const V3d I = load_v3( I0, b );
  if( kind == SG_MAP_EXPONENTIAL_EULER )
  {
    rb3d_flow_exponential_euler( b, nb, x0, R0, vl, w0, m, gx, gy, gz, dt, q1, v1 );
    return;
  }
  // v1 = M * v0: sparse column-major accumulate into zero.  The world inertia block is stored transposed by the state's
  // constructor and as computed once updateMandMinv has run (SG_MAP_M_UPDATED, include/scisim_b200.h)
  V3d p = v3( 0.0 + m * vl.x, 0.0 + m * vl.y, 0.0 + m * vl.z );
  const M3d Iw = world_inertia3( R0, I );
  V3d L;
  if( ( kind_flags & SG_MAP_M_UPDATED ) == 0 )
  {
    L = v3( ( ( 0.0 + Iw.m[0] * w0.x ) + Iw.m[3] * w0.y ) + Iw.m[6] * w0.z,
            ( ( 0.0 + Iw.m[1] * w0.x ) + Iw.m[4] * w0.y ) + Iw.m[7] * w0.z,
            ( ( 0.0 + Iw.m[2] * w0.x ) + Iw.m[5] * w0.y ) + Iw.m[8] * w0.z );
  }
  else
  {
    L = v3( ( ( 0.0 + Iw.m[0] * w0.x ) + Iw.m[1] * w0.y ) + Iw.m[2] * w0.z,
            ( ( 0.0 + Iw.m[3] * w0.x ) + Iw.m[4] * w0.y ) + Iw.m[5] * w0.z,
            ( ( 0.0 + Iw.m[6] * w0.x ) + Iw.m[7] * w0.y ) + Iw.m[8] * w0.z );
  }
  double* x1o = q1 + 3 * size_t( b );
  double* R1o = q1 + 3 * nb + 9 * size_t( b );
  double* vlo = v1 + 3 * size_t( b );
  double* vao = v1 + 3 * nb + 3 * size_t( b );
  if( __ldg( &btype[b] ) & SG_FIXED_BIT )
  {
    // kinematically scripted: q1 = q0, v1 keeps M * v0 (SplitHamMap.cpp:43-53)
    x1o[0] = x0.x; x1o[1] = x0.y; x1o[2] = x0.z;
    #pragma unroll
    for( int k = 0; k < 9; ++k ) { R1o[k] = R0.m[k]; }
    vlo[0] = p.x; vlo[1] = p.y; vlo[2] = p.z;
    vao[0] = L.x; vao[1] = L.y; vao[2] = L.z;
    return;
  }
  const V3d F = v3( 0.0 + m * gx, 0.0 + m * gy, 0.0 + m * gz );
  const double hdt = 0.5 * dt;
  p = v3( p.x + hdt * F.x, p.y + hdt * F.y, p.z + hdt * F.z );
  L = v3( L.x + hdt * 0.0, L.y + hdt * 0.0, L.z + hdt * 0.0 );
  const double sc = ( ( 0.5 * dt ) * dt ) * ( 1.0 / m );
  x1o[0] = x0.x + ( dt * vl.x + ( 0.0 + sc * F.x ) );
  x1o[1] = x0.y + ( dt * vl.y + ( 0.0 + sc * F.y ) );
  x1o[2] = x0.z + ( dt * vl.z + ( 0.0 + sc * F.z ) );
  M3d R1;
  if( kind == SG_MAP_SPLIT_HAM )
  {
    V3d pB = mulT3( R0, L );
    R1 = R0;
    // the five axis rotations of the splitting (SplitHamMap.cpp:121-176): z, y, x, y, z
    splitham_axis_step<2>( R1, pB, 0.5 * dt * pB.z / I.z );
    splitham_axis_step<1>( R1, pB, 0.5 * dt * pB.y / I.y );
    splitham_axis_step<0>( R1, pB, dt * pB.x / I.x );
    splitham_axis_step<1>( R1, pB, 0.5 * dt * pB.y / I.y );
    splitham_axis_step<2>( R1, pB, 0.5 * dt * pB.z / I.z );
  }
  else
  {
    R1 = dmv_rotation( R0, mulT3( R0, L ), dt, I );
  }
  #pragma unroll
  for( int k = 0; k < 9; ++k ) { R1o[k] = R1.m[k]; }
  p = v3( p.x + hdt * F.x, p.y + hdt * F.y, p.z + hdt * F.z );
  L = v3( L.x + hdt * 0.0, L.y + hdt * 0.0, L.z + hdt * 0.0 );
  vlo[0] = p.x / m; vlo[1] = p.y / m; vlo[2] = p.z / m;
  const V3d w1 = mul3( world_inertia3( R1, v3( 1.0 / I.x, 1.0 / I.y, 1.0 / I.z ) ), L );
  vao[0] = w1.x; vao[1] = w1.y; vao[2] = w1.z;
}

// RigidBody3DState::updateMandMinv (rigidbody3d/RigidBody3DState.cpp:428-462): I = R I0 R^T, Iinv = R ( 1 / I0 ) R^T per body,
// written column-major ( entry ( r, c ) at 3 c + r ) as the reference's maps over M's and Minv's value arrays store them
__global__ void __launch_bounds__( 128 ) k_rb3d_update_minertia( const uint32_t n, const double* __restrict__ q, const double* __restrict__ I0, double* __restrict__ I_blocks,
                                                                double* __restrict__ Iinv_blocks )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= n ) { return; }
  const M3d R = load_m3( q + 3 * size_t( n ), b );
  const V3d I = load_v3( I0, b );
  const M3d Iw = world_inertia3( R, I );
  const M3d Ii = world_inertia3( R, v3( 1.0 / I.x, 1.0 / I.y, 1.0 / I.z ) );
  double* o = I_blocks + 9 * size_t( b );
  double* oi = Iinv_blocks + 9 * size_t( b );
  #pragma unroll
  for( int r = 0; r < 3; ++r )
  {
    #pragma unroll
    for( int c = 0; c < 3; ++c ) { o[3 * c + r] = Iw.m[3 * r + c]; oi[3 * c + r] = Ii.m[3 * r + c]; }
  }
}

// ---- AABBs at q1 (generic pipeline) ----------------------------------------------------------------
__global__ void __launch_bounds__( 128 ) k_rb3d_aabb( const Rb3dDev dev, const double* __restrict__ q1, double* __restrict__ boxes )
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if( b >= dev.n ) { return; }
  const size_t nb = dev.n;
  const uint32_t type = __ldg( &dev.btype[b] ) & ~SG_FIXED_BIT;
  const V3d cm = load_v3( q1, b );
  double lo[3], hi[3];
  if( type == SG_GEO_SPHERE )
  {
    const double r = __ldg( &dev.bparam[4 * size_t( b )] );
    lo[0] = cm.x - r; lo[1] = cm.y - r; lo[2] = cm.z - r;
    hi[0] = cm.x + r; hi[1] = cm.y + r; hi[2] = cm.z + r;
  }
  else if( type == SG_GEO_BOX )
  {
    // extents = |R| * half
    const M3d R = load_m3( q1 + 3 * nb, b );
    M3d A;
    #pragma unroll
    for( int k = 0; k < 9; ++k ) { A.m[k] = fabs( R.m[k] ); }
    const V3d e = mul3( A, v3( __ldg( &dev.bparam[4 * size_t( b )] ), __ldg( &dev.bparam[4 * size_t( b ) + 1] ), __ldg( &dev.bparam[4 * size_t( b ) + 2] ) ) );
    lo[0] = cm.x - e.x; lo[1] = cm.y - e.y; lo[2] = cm.z - e.z;
    hi[0] = cm.x + e.x; hi[1] = cm.y + e.y; hi[2] = cm.z + e.z;
  }
  else
  {
    // min / max over R * vert + cm for ALL mesh vertices (RigidBodyTriangleMesh.cpp:187-200)
    const M3d R = load_m3( q1 + 3 * nb, b );
    const MeshDev& mesh = dev.meshes[__ldg( &dev.bmesh[b] )];
    #pragma unroll
    for( int k = 0; k < 3; ++k ) { lo[k] = __longlong_as_double( 0x7ff0000000000000LL ); hi[k] = __longlong_as_double( 0xfff0000000000000LL ); }
    for( uint32_t vi = 0; vi < mesh.nverts; ++vi )
    {
      const V3d t = mul3( R, load_v3( mesh.verts, vi ) ) + cm;
      lo[0] = fmin( lo[0], t.x ); lo[1] = fmin( lo[1], t.y ); lo[2] = fmin( lo[2], t.z );
      hi[0] = fmax( hi[0], t.x ); hi[1] = fmax( hi[1], t.y ); hi[2] = fmax( hi[2], t.z );
    }
  }
  double* o = boxes + 6 * size_t( b );
  o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = hi[0]; o[4] = hi[1]; o[5] = hi[2];
}

// ---- narrow phase over the candidate list (generic pipeline) ----------------------------------------
// Where the 8 corner values of a cell come from: straight from HBM/L2, or from a brick of the grid that the CTA
// staged in shared memory with TMA tensor copies (cp.async.bulk.tensor.3d -> UTMALDG).
struct SdfGlobal
{
  const double* sdf; size_t pitch, ny;
  __device__ __forceinline__ double at( const unsigned i, const unsigned j, const unsigned k ) const { return __ldg( sdf + ( size_t( k ) * ny + j ) * pitch + i ); }
};
struct SdfBrick
{
  const double* s; uint32_t x0, y0, z0, ntx, nty;
  __device__ __forceinline__ double at( const unsigned i, const unsigned j, const unsigned k ) const
  {
    const uint32_t lx = i - x0, ly = j - y0, lz = k - z0;
    const uint32_t tile = ( ( lz / SDF_BZ ) * nty + ( ly / SDF_BY ) ) * ntx + ( lx / SDF_BX );
    return s[tile * SDF_TILE_ELEMS + ( ( lz % SDF_BZ ) * SDF_BY + ( ly % SDF_BY ) ) * SDF_BX + ( lx % SDF_BX )];
  }
};

// cell of a point in the mesh frame; false when RigidBodyTriangleMesh::detectCollision rejects it on bounds
__device__ __forceinline__ bool sdf_cell( const MeshDev& mesh, const V3d x, unsigned& ix, unsigned& iy, unsigned& iz )
{
  if( x.x < mesh.origin[0] || x.y < mesh.origin[1] || x.z < mesh.origin[2] ) { return false; }
  if( x.x > mesh.grid_end[0] || x.y > mesh.grid_end[1] || x.z > mesh.grid_end[2] ) { return false; }
  ix = unsigned( floor( ( x.x - mesh.origin[0] ) / mesh.delta[0] ) );
  iy = unsigned( floor( ( x.y - mesh.origin[1] ) / mesh.delta[1] ) );
  iz = unsigned( floor( ( x.z - mesh.origin[2] ) / mesh.delta[2] ) );
  // the reference only asserts this; a sample exactly on grid_end is treated as a miss (same guard as the oracle)
  return !( ix + 1u >= mesh.dims[0] || iy + 1u >= mesh.dims[1] || iz + 1u >= mesh.dims[2] );
}

// RigidBodyTriangleMesh::detectCollision (RigidBodyTriangleMesh.cpp:276-335); n in the mesh frame
template<typename Acc>
__device__ inline bool sdf_detect( const MeshDev& mesh, const Acc& acc, const V3d x, V3d& n )
{
  unsigned ix, iy, iz;
  if( !sdf_cell( mesh, x, ix, iy, iz ) ) { return false; }
  const double bcx = ( x.x - ( mesh.origin[0] + double( ix ) * mesh.delta[0] ) ) / mesh.delta[0];
  const double bcy = ( x.y - ( mesh.origin[1] + double( iy ) * mesh.delta[1] ) ) / mesh.delta[1];
  const double bcz = ( x.z - ( mesh.origin[2] + double( iz ) * mesh.delta[2] ) ) / mesh.delta[2];
  const double bix = 1.0 - bcx, biy = 1.0 - bcy, biz = 1.0 - bcz;
  const double v000 = acc.at( ix, iy, iz ),         v100 = acc.at( ix + 1, iy, iz );
  const double v010 = acc.at( ix, iy + 1, iz ),     v110 = acc.at( ix + 1, iy + 1, iz );
  const double v001 = acc.at( ix, iy, iz + 1 ),     v101 = acc.at( ix + 1, iy, iz + 1 );
  const double v011 = acc.at( ix, iy + 1, iz + 1 ), v111 = acc.at( ix + 1, iy + 1, iz + 1 );
  const double dist = biz * ( biy * ( bix * v000 + bcx * v100 ) + bcy * ( bix * v010 + bcx * v110 ) ) +
                      bcz * ( biy * ( bix * v001 + bcx * v101 ) + bcy * ( bix * v011 + bcx * v111 ) );
  if( dist > 0.0 ) { return false; }
  V3d g;
  g.x = biz * ( biy * ( v100 - v000 ) + bcy * ( v110 - v010 ) ) + bcz * ( biy * ( v101 - v001 ) + bcy * ( v111 - v011 ) );
  g.y = biz * ( bix * ( v010 - v000 ) + bcx * ( v110 - v100 ) ) + bcz * ( bix * ( v011 - v001 ) + bcx * ( v111 - v101 ) );
  g.z = biy * ( bix * ( v001 - v000 ) + bcx * ( v101 - v100 ) ) + bcy * ( bix * ( v011 - v010 ) + bcx * ( v111 - v110 ) );
  g.x /= mesh.delta[0]; g.y /= mesh.delta[1]; g.z /= mesh.delta[2];
  n = normalized3( g );
  return true;
}

__device__ __forceinline__ void sg_tma_load_3d( void* smem_dst, const CUtensorMap* map, const int c0, const int c1, const int c2, uint64_t* bar )
{
  asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"( sg_smem_u32( smem_dst ) ), "l"( map ), "r"( c0 ), "r"( c1 ), "r"( c2 ), "r"( sg_smem_u32( bar ) ) : "memory" );
}

#define SG_PAIR_SKIP 0u
#define SG_PAIR_SPHERE 1u
#define SG_PAIR_BOX 2u
#define SG_PAIR_MESH 3u
#define SG_PAIR_BAD 4u

// kinematic rules of dispatchNarrowPhaseCollision: both fixed -> skip; the fixed body goes second
__device__ __forceinline__ uint32_t pair_kind( const Rb3dDev& dev, const uint2 pr, uint32_t& b0, uint32_t& b1, bool& kin )
{
  const uint32_t t0 = __ldg( &dev.btype[pr.x] ), t1 = __ldg( &dev.btype[pr.y] );
  b0 = pr.x; b1 = pr.y;
  const bool f0 = ( t0 & SG_FIXED_BIT ) != 0u, f1 = ( t1 & SG_FIXED_BIT ) != 0u;
  if( f0 && f1 ) { return SG_PAIR_SKIP; }
  if( f0 ) { b0 = pr.y; b1 = pr.x; }
  kin = f0 || f1;
  const uint32_t g0 = t0 & ~SG_FIXED_BIT, g1 = t1 & ~SG_FIXED_BIT;
  if( g0 != g1 ) { return SG_PAIR_BAD; }
  if( g0 == SG_GEO_SPHERE ) { return SG_PAIR_SPHERE; }
  if( g0 == SG_GEO_BOX ) { return SG_PAIR_BOX; }
  if( g0 == SG_GEO_MESH ) { return SG_PAIR_MESH; }
  return SG_PAIR_BAD;
}

__device__ __forceinline__ V3d body_half( const Rb3dDev& dev, const uint32_t b )
{
  return v3( __ldg( &dev.bparam[4 * size_t( b )] ), __ldg( &dev.bparam[4 * size_t( b ) + 1] ), __ldg( &dev.bparam[4 * size_t( b ) + 2] ) );
}

// One thread per candidate pair (sphere-sphere, box-box).  EMIT = false: counts[pair] = number of contacts;
// EMIT = true: contacts written at offsets[pair].  Mesh pairs are left to k_rb3d_mesh_pairs (count 0 here).
template<bool EMIT>
__global__ void __launch_bounds__( 128 ) k_rb3d_pairs( const Rb3dDev dev, const uint2* __restrict__ pairs, const unsigned long long* __restrict__ npairs_dev,
                                                      const double* __restrict__ q0, const double* __restrict__ q1, uint32_t* __restrict__ counts,
                                                      const unsigned long long* __restrict__ offsets, const ContactOut3D out, uint32_t* __restrict__ bad_flag )
{
  const unsigned long long k = blockIdx.x * ( unsigned long long )( blockDim.x ) + threadIdx.x;
  if( k >= *npairs_dev ) { return; }
  const uint2 pr = pairs[k];
  uint32_t b0, b1;
  bool kin = false;
  const uint32_t kind = pair_kind( dev, pr, b0, b1, kin );
  const size_t nb = dev.n;
  uint32_t cnt = 0u;
  if( kind == SG_PAIR_BAD ) { if( !EMIT ) { atomicOr( bad_flag, 1u ); } }
  else if( kind == SG_PAIR_SPHERE )
  {
    const double ra = __ldg( &dev.bparam[4 * size_t( pr.x )] ), rb = __ldg( &dev.bparam[4 * size_t( pr.y )] );
    const bool fa = ( __ldg( &dev.btype[pr.x] ) & SG_FIXED_BIT ) != 0u, fb = ( __ldg( &dev.btype[pr.y] ) & SG_FIXED_BIT ) != 0u;
    const V3d x1a = load_v3( q1, pr.x ), x1b = load_v3( q1, pr.y );
    if( sphere_pair_active( x1a, ra, fa, x1b, rb, fb ) )
    {
      cnt = 1u;
      if( EMIT ) { sphere_pair_emit( out, offsets[k], pr.x, load_v3( q0, pr.x ), x1a, ra, fa, pr.y, load_v3( q0, pr.y ), x1b, rb, fb ); }
    }
  }
  else if( kind == SG_PAIR_BOX )
  {
    V3d n;
    double pts[24];
    const int nc = sg_box_box( load_v3( q1, b0 ), load_m3( q1 + 3 * nb, b0 ), body_half( dev, b0 ), load_v3( q1, b1 ), load_m3( q1 + 3 * nb, b1 ), body_half( dev, b1 ), n, pts );
    cnt = uint32_t( nc );
    if( EMIT )
    {
      const unsigned long long o = offsets[k];
      for( int c = 0; c < nc; ++c ) { put_contact( out, o + c, kin ? SG_KINEMATIC_BODY_BODY : SG_BODY_BODY, b0, b1, 0u, n, v3( pts[3 * c], pts[3 * c + 1], pts[3 * c + 2] ), sg_nan() ); }
    }
  }
  if( !EMIT ) { counts[k] = cnt; }
}

// One direction of MeshMeshUtilities::computeActiveSet for one pair: src's samples against dst's distance field.
// Counting needs no order (per-thread tallies, summed by the caller); emitting writes contacts in sample order:
// ballot inside the warp, per-warp totals through a double-buffered shared array (one barrier per 256 samples),
// the running offset carried in a register by every thread.  Returns the number of hits this thread saw (count
// pass) / the new running offset (emit pass).
template<bool EMIT, typename Acc>
__device__ inline uint32_t mesh_dir_samples( const MeshDev& src, const MeshDev& dst, const Acc& acc, const M3d& Rsd, const V3d xsd, const M3d& Rd, const V3d cd, const int dir, const bool kin,
                                             const uint32_t b0, const uint32_t b1, const unsigned long long base, const ContactOut3D& out, uint32_t ( *s_warp )[8], uint32_t run )
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t buf = 0u;
  for( uint32_t s0 = 0; s0 < src.nsamples; s0 += blockDim.x )
  {
    const uint32_t si = s0 + threadIdx.x;
    bool hit = false;
    V3d x = v3( 0.0, 0.0, 0.0 ), normal = x;
    if( si < src.nsamples )
    {
      x = mul3( Rsd, load_v3( src.samples, si ) ) + xsd;
      hit = sdf_detect( dst, acc, x, normal );
    }
    if( !EMIT ) { run += hit ? 1u : 0u; continue; }
    const unsigned bal = __ballot_sync( 0xffffffffu, hit );
    if( lane == 0 ) { s_warp[buf][warp] = __popc( bal ); }
    __syncthreads();
    uint32_t before = 0u, total = 0u;
    #pragma unroll
    for( int w = 0; w < 8; ++w ) { const uint32_t c = s_warp[buf][w]; if( w < warp ) { before += c; } total += c; }
    if( hit )
    {
      const V3d pw = mul3( Rd, x ) + cd;
      V3d nw = mul3( Rd, normal );
      if( dir == 1 ) { nw = -nw; }
      put_contact( out, base + run + before + __popc( bal & ( ( 1u << lane ) - 1u ) ), kin ? SG_KINEMATIC_BODY_BODY : SG_BODY_BODY, b0, b1, 0u, nw, pw, sg_nan() );
    }
    run += total;
    buf ^= 1u;
  }
  return run;
}

// One CTA per candidate pair; only mesh-mesh pairs do work (MeshMeshUtilities.cpp:10-65): mesh0's samples against
// mesh1's distance field, then mesh1's samples against mesh0's.  Per direction the CTA first finds the brick of
// grid cells its in-range samples fall into and, when it fits, pulls it into shared memory with TMA tensor copies
// (one SDF_BX x SDF_BY x SDF_BZ tile each, completion on an mbarrier); the trilinear lookups then read shared memory.
template<bool EMIT>
__global__ void __launch_bounds__( 256 ) k_rb3d_mesh_pairs( const Rb3dDev dev, const uint2* __restrict__ pairs, const unsigned long long* __restrict__ npairs_dev, const double* __restrict__ q1,
                                                           uint32_t* __restrict__ counts, const unsigned long long* __restrict__ offsets, const ContactOut3D out, const uint32_t tile_cap )
{
  extern __shared__ __align__( 128 ) unsigned char s_dyn[];
  double* s_brick = reinterpret_cast<double*>( s_dyn );
  __shared__ uint32_t s_warp[2][8];
  __shared__ uint32_t s_count;
  __shared__ int s_ext[6];
  __shared__ __align__( 8 ) unsigned long long s_bar;
  uint64_t* bar = reinterpret_cast<uint64_t*>( &s_bar );
  if( threadIdx.x == 0 ) { sg_mbar_init( bar, 1u ); }
  __syncthreads();
  uint32_t phase = 0u;
  for( unsigned long long k = blockIdx.x; k < *npairs_dev; k += gridDim.x )
  {
    const uint2 pr = pairs[k];
    uint32_t b0, b1;
    bool kin = false;
    if( pair_kind( dev, pr, b0, b1, kin ) != SG_PAIR_MESH ) { continue; } // uniform across the CTA
    const size_t nb = dev.n;
    const MeshDev& mesh0 = dev.meshes[__ldg( &dev.bmesh[b0] )];
    const MeshDev& mesh1 = dev.meshes[__ldg( &dev.bmesh[b1] )];
    const V3d cm0 = load_v3( q1, b0 ), cm1 = load_v3( q1, b1 );
    const M3d R0 = load_m3( q1 + 3 * nb, b0 ), R1 = load_m3( q1 + 3 * nb, b1 );
    const unsigned long long base = EMIT ? offsets[k] : 0ull;
    uint32_t run = 0u; // count pass: this thread's hits; emit pass: contacts written so far for the pair (uniform)
    for( int dir = 0; dir < 2; ++dir )
    {
      const MeshDev& src = dir == 0 ? mesh0 : mesh1;
      const MeshDev& dst = dir == 0 ? mesh1 : mesh0;
      const M3d& Rs = dir == 0 ? R0 : R1;
      const M3d& Rd = dir == 0 ? R1 : R0;
      const V3d cs = dir == 0 ? cm0 : cm1;
      const V3d cd = dir == 0 ? cm1 : cm0;
      const M3d Rsd = mulTN33( Rd, Rs );
      const V3d xsd = mulT3( Rd, cs - cd );
      // brick of cells touched by the in-range samples
      if( threadIdx.x < 6 ) { s_ext[threadIdx.x] = ( threadIdx.x < 3 ) ? 0x7fffffff : -1; }
      __syncthreads();
      {
        // Only a covering brick is needed here, so the cell is estimated with a reciprocal multiply (no fp64 divide)
        // and widened by one cell each way; the exact reference arithmetic runs in the sweep below.
        int mn[3] = { 0x7fffffff, 0x7fffffff, 0x7fffffff }, mx[3] = { -1, -1, -1 };
        const double inv0 = 1.0 / dst.delta[0], inv1 = 1.0 / dst.delta[1], inv2 = 1.0 / dst.delta[2];
        for( uint32_t si = threadIdx.x; si < src.nsamples; si += blockDim.x )
        {
          const V3d x = mul3( Rsd, load_v3( src.samples, si ) ) + xsd;
          if( x.x < dst.origin[0] || x.y < dst.origin[1] || x.z < dst.origin[2] || x.x > dst.grid_end[0] || x.y > dst.grid_end[1] || x.z > dst.grid_end[2] ) { continue; }
          const int ex = int( ( x.x - dst.origin[0] ) * inv0 ), ey = int( ( x.y - dst.origin[1] ) * inv1 ), ez = int( ( x.z - dst.origin[2] ) * inv2 );
          mn[0] = min( mn[0], ex - 1 ); mn[1] = min( mn[1], ey - 1 ); mn[2] = min( mn[2], ez - 1 );
          mx[0] = max( mx[0], ex + 1 ); mx[1] = max( mx[1], ey + 1 ); mx[2] = max( mx[2], ez + 1 );
        }
        #pragma unroll
        for( int a = 0; a < 3; ++a )
        {
          #pragma unroll
          for( int d = 16; d > 0; d >>= 1 ) { mn[a] = min( mn[a], __shfl_xor_sync( 0xffffffffu, mn[a], d ) ); mx[a] = max( mx[a], __shfl_xor_sync( 0xffffffffu, mx[a], d ) ); }
          if( ( threadIdx.x & 31 ) == 0 ) { atomicMin( &s_ext[a], mn[a] ); atomicMax( &s_ext[3 + a], mx[a] ); }
        }
      }
      __syncthreads();
      // measured on B200: a tensor copy of 8-byte elements faults (illegal instruction) unless the innermost start
      // coordinate is 16-byte aligned, so bricks start on an even x
      const int x0 = max( s_ext[0], 0 ) & ~1, y0 = max( s_ext[1], 0 ), z0 = max( s_ext[2], 0 );
      const bool any = s_ext[3] >= 0;
      // corners go one cell beyond the largest cell index; cells past dims-2 are rejected by the sweep
      const int x1 = min( s_ext[3], int( dst.dims[0] ) - 2 ) + 2, y1 = min( s_ext[4], int( dst.dims[1] ) - 2 ) + 2, z1 = min( s_ext[5], int( dst.dims[2] ) - 2 ) + 2;
      const uint32_t ntx = any ? uint32_t( x1 - x0 + SDF_BX - 1 ) / SDF_BX : 0u;
      const uint32_t nty = any ? uint32_t( y1 - y0 + SDF_BY - 1 ) / SDF_BY : 0u;
      const uint32_t ntz = any ? uint32_t( z1 - z0 + SDF_BZ - 1 ) / SDF_BZ : 0u;
      const uint32_t ntiles = ntx * nty * ntz;
      const bool staged = any && dst.has_tmap != 0u && ntiles <= tile_cap;
      __syncthreads(); // s_ext is consumed; the next direction may reset it
      if( !any ) { continue; } // no sample of src lies inside dst's grid: nothing can collide (uniform)
      if( !EMIT && threadIdx.x == 0 ) { atomicAdd( &dev.mesh_stats[staged ? 0 : 1], 1ull ); }
      if( staged )
      {
        if( threadIdx.x == 0 )
        {
          asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" ); // earlier generic reads of the brick vs the async writes to come
          // the descriptor was written to global memory by a host copy: acquire it for the tensormap proxy
          asm volatile( "fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"( &dst.tmap ) : "memory" );
          sg_mbar_arrive_expect_tx( bar, ntiles * SDF_TILE_ELEMS * 8u );
          for( uint32_t t = 0; t < ntiles; ++t )
          {
            const uint32_t tx = t % ntx, ty = ( t / ntx ) % nty, tz = t / ( ntx * nty );
            sg_tma_load_3d( s_brick + size_t( t ) * SDF_TILE_ELEMS, &dst.tmap, x0 + int( tx ) * SDF_BX, y0 + int( ty ) * SDF_BY, z0 + int( tz ) * SDF_BZ, bar );
          }
        }
        sg_mbar_wait( bar, phase );
        phase ^= 1u;
        SdfBrick acc; acc.s = s_brick; acc.x0 = uint32_t( x0 ); acc.y0 = uint32_t( y0 ); acc.z0 = uint32_t( z0 ); acc.ntx = ntx; acc.nty = nty;
        run = mesh_dir_samples<EMIT>( src, dst, acc, Rsd, xsd, Rd, cd, dir, kin, b0, b1, base, out, s_warp, run );
      }
      else
      {
        SdfGlobal acc; acc.sdf = dst.sdf; acc.pitch = dst.pitch; acc.ny = dst.dims[1];
        run = mesh_dir_samples<EMIT>( src, dst, acc, Rsd, xsd, Rd, cd, dir, kin, b0, b1, base, out, s_warp, run );
      }
      __syncthreads();
    }
    if( !EMIT )
    {
      if( threadIdx.x == 0 ) { s_count = 0u; }
      __syncthreads();
      #pragma unroll
      for( int d = 16; d > 0; d >>= 1 ) { run += __shfl_xor_sync( 0xffffffffu, run, d ); }
      if( ( threadIdx.x & 31 ) == 0 && run != 0u ) { atomicAdd( &s_count, run ); }
      __syncthreads();
      if( threadIdx.x == 0 ) { counts[k] = s_count; }
    }
  }
}

// ---- body-plane (RigidBody3DSim.cpp:1414-1502): plane-major, body ascending, corner / hull vertex ascending ----
// Number of contacts of body b against plane pl; when emit_base != ~0 they are written starting there.
__device__ inline uint32_t plane_contacts( const Rb3dDev& dev, const Planes3D& planes, const uint32_t pl, const uint32_t b, const double* __restrict__ q0, const double* __restrict__ q1,
                                           const bool emit, const unsigned long long emit_base, const ContactOut3D& out )
{
  const uint32_t t = __ldg( &dev.btype[b] );
  if( t & SG_FIXED_BIT ) { return 0u; }
  const size_t nb = dev.n;
  if( pl >= planes.n )
  {
    // static cylinder pl - planes.n (RigidBody3DSim.cpp:1504-1557): spheres and mesh hull vertices OUTSIDE the cylinder
    // are in contact; boxes are rejected on the host like the reference does
    const uint32_t cy = pl - planes.n;
    const V3d xc = v3( planes.cx[cy][0], planes.cx[cy][1], planes.cx[cy][2] );
    const V3d ax = v3( planes.cax[cy][0], planes.cax[cy][1], planes.cax[cy][2] );
    const double rc = planes.cr[cy];
    const V3d x1 = load_v3( q1, b );
    uint32_t cnt = 0u;
    if( t == SG_GEO_SPHERE )
    {
      const double r = __ldg( &dev.bparam[4 * size_t( b )] );
      const V3d d = ( x1 - xc ) - dot3( ax, x1 - xc ) * ax;           // StaticCylinderSphereConstraint.cpp:10-19
      if( dot3( d, d ) >= ( rc - r ) * ( rc - r ) )
      {
        if( emit )
        {
          const V3d x0 = load_v3( q0, b );
          const V3d e0 = ( x0 - xc ) - dot3( ax, x0 - xc ) * ax;
          const V3d n = normalized3( -e0 );                          // computeN( q0 ) (:237-244)
          // computePenetrationDepth( q1 ) = min( 0, R - |d| - r ) (StaticCylinderSphereConstraint.cpp:311-317)
          put_contact( out, emit_base, SG_CYLINDER_SPHERE, b, cy, 0u, n, x0 - r * n, fmin( 0.0, rc - sqrt( dot3( d, d ) ) - r ) );
        }
        cnt = 1u;
      }
    }
    else if( t == SG_GEO_MESH )
    {
      const M3d R1 = load_m3( q1 + 3 * nb, b );
      const MeshDev& mesh = dev.meshes[__ldg( &dev.bmesh[b] )];
      for( uint32_t vi = 0; vi < mesh.nhull; ++vi )
      {
        const V3d hv = load_v3( mesh.hull, vi );
        const V3d v = mul3( R1, hv ) + x1;
        const V3d d = ( v - xc ) - dot3( ax, v - xc ) * ax;          // MeshMeshUtilities.cpp:88-109
        if( dot3( d, d ) >= rc * rc )
        {
          if( emit )
          {
            const V3d x0 = load_v3( q0, b );
            const V3d e0 = ( x0 - xc ) - dot3( ax, x0 - xc ) * ax;
            const V3d m0 = -e0;
            const double nrm = sqrt( dot3( m0, m0 ) );               // n / n.norm() (StaticCylinderBodyConstraint.cpp:84-91)
            put_contact( out, emit_base + cnt, SG_CYLINDER_BODY, b, cy, vi, v3( m0.x / nrm, m0.y / nrm, m0.z / nrm ), x0 + mul3( load_m3( q0 + 3 * nb, b ), hv ), sg_nan() );
          }
          ++cnt;
        }
      }
    }
    return cnt;
  }
  const V3d xp = v3( planes.x[pl][0], planes.x[pl][1], planes.x[pl][2] );
  const V3d np = v3( planes.nrm[pl][0], planes.nrm[pl][1], planes.nrm[pl][2] );
  const V3d x1 = load_v3( q1, b );
  uint32_t cnt = 0u;
  if( t == SG_GEO_SPHERE )
  {
    const double r = __ldg( &dev.bparam[4 * size_t( b )] );
    const double d = dot3( np, x1 - xp );
    if( d <= r )
    {
      if( emit ) { put_contact( out, emit_base, SG_PLANE_SPHERE, b, pl, 0u, np, load_v3( q0, b ) - r * np, fmin( 0.0, d - r ) ); }
      cnt = 1u;
    }
  }
  else if( t == SG_GEO_BOX )
  {
    const M3d R1 = load_m3( q1 + 3 * nb, b );
    const V3d half = body_half( dev, b );
    for( int corner = 0; corner < 8; ++corner )
    {
      const V3d cb = v3( half.x * double( 2 * ( corner % 2 ) - 1 ), half.y * double( 2 * ( ( corner >> 1 ) % 2 ) - 1 ), half.z * double( 2 * ( ( corner >> 2 ) % 2 ) - 1 ) );
      const V3d wp = mul3( R1, cb ) + x1;
      if( dot3( np, wp - xp ) <= 0.0 )
      {
        if( emit ) { put_contact( out, emit_base + cnt, SG_PLANE_BOX, b, pl, uint32_t( corner ), np, mul3( load_m3( q0 + 3 * nb, b ), cb ) + load_v3( q0, b ), sg_nan() ); }
        ++cnt;
      }
    }
  }
  else
  {
    const M3d R1 = load_m3( q1 + 3 * nb, b );
    const MeshDev& mesh = dev.meshes[__ldg( &dev.bmesh[b] )];
    for( uint32_t vi = 0; vi < mesh.nhull; ++vi )
    {
      const V3d hv = load_v3( mesh.hull, vi );
      const V3d v = mul3( R1, hv ) + x1;
      if( dot3( np, v - xp ) <= 0.0 )
      {
        if( emit ) { put_contact( out, emit_base + cnt, SG_PLANE_BODY, b, pl, vi, np, load_v3( q0, b ) + mul3( load_m3( q0 + 3 * nb, b ), hv ), sg_nan() ); }
        ++cnt;
      }
    }
  }
  return cnt;
}

// counts[pl * nblocks + block] = contacts of this block's bodies against plane pl
__global__ void __launch_bounds__( 256 ) k_rb3d_plane_count( const Rb3dDev dev, const __grid_constant__ Planes3D planes, const double* __restrict__ q0, const double* __restrict__ q1, uint32_t* __restrict__ counts )
{
  __shared__ uint32_t s_cnt[SG_MAX_PLANES + SG_MAX_CYLINDERS];
  const uint32_t ng = planes.n + planes.ncyl;
  if( threadIdx.x < ng ) { s_cnt[threadIdx.x] = 0u; }
  __syncthreads();
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const ContactOut3D none = {};
  for( uint32_t pl = 0; pl < ng; ++pl )
  {
    uint32_t c = ( b < dev.n_live ) ? plane_contacts( dev, planes, pl, b, q0, q1, false, 0ull, none ) : 0u;
    #pragma unroll
    for( int d = 16; d > 0; d >>= 1 ) { c += __shfl_xor_sync( 0xffffffffu, c, d ); }
    if( ( threadIdx.x & 31 ) == 0 && c != 0u ) { atomicAdd( &s_cnt[pl], c ); }
  }
  __syncthreads();
  if( threadIdx.x < ng ) { counts[threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x]; }
}

__global__ void __launch_bounds__( 256 ) k_rb3d_plane_emit( const Rb3dDev dev, const __grid_constant__ Planes3D planes, const double* __restrict__ q0, const double* __restrict__ q1,
                                                           const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, const unsigned long long* __restrict__ base_dev, const ContactOut3D out )
{
  __shared__ uint32_t s_warp[8];
  const uint32_t ng = planes.n + planes.ncyl;
  const int mine = ( threadIdx.x < ng ) ? int( counts[threadIdx.x * gridDim.x + blockIdx.x] != 0u ) : 0;
  if( __syncthreads_or( mine ) == 0 ) { return; }
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long base = *base_dev;
  const ContactOut3D none = {};
  for( uint32_t pl = 0; pl < ng; ++pl )
  {
    if( counts[pl * gridDim.x + blockIdx.x] == 0u ) { continue; }
    const uint32_t c = ( b < dev.n_live ) ? plane_contacts( dev, planes, pl, b, q0, q1, false, 0ull, none ) : 0u;
    // exclusive prefix of c over the block
    uint32_t incl = c;
    #pragma unroll
    for( int d = 1; d < 32; d <<= 1 ) { const uint32_t o = __shfl_up_sync( 0xffffffffu, incl, d ); if( lane >= d ) { incl += o; } }
    __syncthreads();
    if( lane == 31 ) { s_warp[warp] = incl; }
    __syncthreads();
    uint32_t before = incl - c;
    for( int w = 0; w < warp; ++w ) { before += s_warp[w]; }
    if( c != 0u ) { plane_contacts( dev, planes, pl, b, q0, q1, true, base + offsets[pl * gridDim.x + blockIdx.x] + before, out ); }
  }
}

// ---- slab mode (multi-GPU, all-sphere scenes): what sg_slab.cuh needs to know about spheres -------------------------------------
// A sphere's broad-phase box is taken at q1 only (RigidBodySphere::computeAABB through RigidBody3DSim::generateAABBs, RigidBody3DSim.cpp:1057-1069),
// so its reach on x is [ x1.x - r, x1.x + r ]; the record carries x0 as well because the contact is built from the start-of-step positions.
struct alignas( 16 ) SphereGhostRec { double x0[3]; double x1[3]; double r; uint32_t gid; uint32_t pad; };
static_assert( sizeof( SphereGhostRec ) == 64, "sphere halo records are 64 bytes" );

struct Sphere3DSlabTraits
{
  using Rec = SphereGhostRec;
  struct Src { const double* q0; const double* q1; const double* r; const uint32_t* gid; };
  __device__ static bool select( const Src& s, const uint32_t i, const double ilo, const double ihi, Rec& g )
  {
    #pragma unroll
    for( int k = 0; k < 3; ++k ) { g.x0[k] = __ldg( &s.q0[3 * size_t( i ) + k] ); g.x1[k] = __ldg( &s.q1[3 * size_t( i ) + k] ); }
    g.r = __ldg( &s.r[i] ); g.gid = s.gid[i]; g.pad = 0u;
    const double lo = g.x1[0] - g.r, hi = g.x1[0] + g.r;
    return !( hi < ilo ) && !( ihi < lo );
  }
};

// After the flow: [min lo.x, max hi.x] over this rank's spheres for the neighbours, the x-limit guard (a sphere outside could touch a body two
// slabs away: SG_ERR_REBALANCE), and the candidate lists for the halo pack (see SlabCand in sg_slab.cuh).
__global__ void __launch_bounds__( 256 ) k_rb3d_slab_scan( const uint32_t n_owned, const double* __restrict__ q1, const double* __restrict__ r, long long* __restrict__ interval_enc,
                                                          const double xlim_lo, const double xlim_hi, uint32_t* __restrict__ slab_flags, const SlabCand sc )
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n_owned;
  double lo = __longlong_as_double( 0x7ff0000000000000LL ), hi = __longlong_as_double( 0xfff0000000000000LL );
  if( live )
  {
    const double x = __ldg( &q1[3 * size_t( i )] ), rad = __ldg( &r[i] );
    lo = x - rad; hi = x + rad;
    if( lo < xlim_lo || hi > xlim_hi ) { slab_flags[3] = 1u; }
  }
  const int lane = threadIdx.x & 31;
  if( sc.band != nullptr )
  {
    #pragma unroll
    for( int sd = 0; sd < 2; ++sd )
    {
      if( !sc.on[sd] ) { continue; }
      const bool c = live && ( ( sd == 0 ) ? ( lo <= sc.band[0] ) : ( hi >= sc.band[1] ) );
      const unsigned bal = __ballot_sync( 0xffffffffu, c );
      if( bal != 0u )
      {
        uint32_t base = 0u;
        if( lane == __ffs( bal ) - 1 ) { base = atomicAdd( &sc.count[sd], uint32_t( __popc( bal ) ) ); }
        base = __shfl_sync( 0xffffffffu, base, __ffs( bal ) - 1 );
        const uint32_t k = base + __popc( bal & ( ( 1u << lane ) - 1u ) );
        if( c && k < sc.cap ) { sc.list[sd][k] = i; }
      }
    }
  }
  __shared__ double s_iv[8][2];
  #pragma unroll
  for( int dd = 16; dd > 0; dd >>= 1 ) { lo = fmin( lo, __shfl_xor_sync( 0xffffffffu, lo, dd ) ); hi = fmax( hi, __shfl_xor_sync( 0xffffffffu, hi, dd ) ); }
  if( lane == 0 ) { s_iv[threadIdx.x >> 5][0] = lo; s_iv[threadIdx.x >> 5][1] = hi; }
  __syncthreads();
  if( threadIdx.x == 0 )
  {
    for( int w = 1; w < 8; ++w ) { lo = fmin( lo, s_iv[w][0] ); hi = fmax( hi, s_iv[w][1] ); }
    if( lo <= hi )
    {
      atomicMin( &interval_enc[0], sg_ordered_from_double( lo ) );
      atomicMax( &interval_enc[1], sg_ordered_from_double( hi ) );
    }
  }
}

// Halo records from the neighbours' mailboxes into the ghost slots behind the owned bodies (first half of the grid: side 0, second: side 1)
struct SphereUnpackArgs
{
  const SphereGhostRec* in[2];
  const uint32_t* wait[2];
  bool on[2];
  uint32_t* err;
  uint32_t step;
};
__global__ void __launch_bounds__( 256 ) k_rb3d_slab_unpack( const uint32_t cap, const uint32_t blocks_per_side, const SphereUnpackArgs args, const uint32_t n_owned, double* __restrict__ q0, double* __restrict__ q1,
                                                            double* __restrict__ radius, uint32_t* __restrict__ gid, uint32_t* __restrict__ ghost_counts )
{
  const int side = ( blockIdx.x >= blocks_per_side ) ? 1 : 0;
  if( !args.on[side] ) { return; }
  if( threadIdx.x == 0 ) { slab_wait_flag( args.wait[side], args.step, args.err ); }
  __syncthreads();
  const SphereGhostRec* in = args.in[side];
  const uint32_t blk = blockIdx.x - uint32_t( side ) * blocks_per_side;
  const uint32_t k = blk * blockDim.x + threadIdx.x;
  const uint32_t sent = *reinterpret_cast<const volatile uint32_t*>( &in[0].gid );
  const uint32_t count = sent < cap ? sent : cap;
  if( k == 0u )
  {
    ghost_counts[side] = count;
    if( sent > cap ) { ghost_counts[2] = 1u; } // more ghosts than reserved slots: reported by detect
  }
  if( k >= count ) { return; }
  const int4* src = reinterpret_cast<const int4*>( &in[1u + k] );
  union { SphereGhostRec g; int4 v[4]; } u;
  u.v[0] = src[0]; u.v[1] = src[1]; u.v[2] = src[2]; u.v[3] = src[3];
  const size_t slot = size_t( n_owned ) + size_t( side ) * cap + k;
  #pragma unroll
  for( int c = 0; c < 3; ++c ) { q0[3 * slot + c] = u.g.x0[c]; q1[3 * slot + c] = u.g.x1[c]; }
  radius[slot] = u.g.r;
  gid[slot] = u.g.gid;
}

// ---- body-plane / body-cylinder, all-sphere scenes: each sphere is loaded ONCE and tested against every plane and cylinder from registers
// (the generic kernels above go through plane_contacts once per plane and body).  Same tests, same order: planes plane-major, then
// cylinders, body ascending (RigidBody3DSim.cpp:1414-1557, StaticPlaneSphereConstraint.cpp:13-17, StaticCylinderSphereConstraint.cpp:10-19).
__device__ __forceinline__ unsigned long long sphere_static_mask( const Planes3D& planes, const V3d x1, const double r )
{
  unsigned long long mask = 0ull;
  for( uint32_t pl = 0; pl < planes.n; ++pl )
  {
    const V3d xp = v3( planes.x[pl][0], planes.x[pl][1], planes.x[pl][2] );
    const V3d np = v3( planes.nrm[pl][0], planes.nrm[pl][1], planes.nrm[pl][2] );
    if( dot3( np, x1 - xp ) <= r ) { mask |= 1ull << pl; }
  }
  for( uint32_t cy = 0; cy < planes.ncyl; ++cy )
  {
    const V3d xc = v3( planes.cx[cy][0], planes.cx[cy][1], planes.cx[cy][2] );
    const V3d ax = v3( planes.cax[cy][0], planes.cax[cy][1], planes.cax[cy][2] );
    const double rc = planes.cr[cy];
    const V3d d = ( x1 - xc ) - dot3( ax, x1 - xc ) * ax;
    if( dot3( d, d ) >= ( rc - r ) * ( rc - r ) ) { mask |= 1ull << ( planes.n + cy ); }
  }
  return mask;
}

template<bool EMIT>
__global__ void __launch_bounds__( 256 ) k_rb3d_static_spheres( const uint32_t n_live, const __grid_constant__ Planes3D planes, const double* __restrict__ q0, const double* __restrict__ q1, const double* __restrict__ radius,
                                                               const uint32_t* __restrict__ flags, uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                               const unsigned long long* __restrict__ base_dev, const ContactOut3D out )
{
  __shared__ uint32_t s_w[SG_MAX_PLANES + SG_MAX_CYLINDERS][8];
  const uint32_t ng = planes.n + planes.ncyl;
  if( EMIT )
  {
    // most blocks touch nothing: the counts say so before any sphere is read
    const int mine = ( threadIdx.x < ng ) ? int( counts[threadIdx.x * gridDim.x + blockIdx.x] != 0u ) : 0;
    if( __syncthreads_or( mine ) == 0 ) { return; }
  }
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long mask = 0ull;
  V3d x1 = v3( 0.0, 0.0, 0.0 );
  double r = 0.0;
  if( b < n_live && ( __ldg( &flags[b] ) & SG_FIXED_BIT ) == 0u )
  {
    x1 = load_v3( q1, b );
    r = __ldg( &radius[b] );
    mask = sphere_static_mask( planes, x1, r );
  }
  for( uint32_t g = 0; g < ng; ++g )
  {
    const unsigned bal = __ballot_sync( 0xffffffffu, ( mask >> g ) & 1ull );
    if( lane == 0 ) { s_w[g][warp] = __popc( bal ); }
  }
  __syncthreads();
  if( !EMIT )
  {
    if( threadIdx.x < ng )
    {
      uint32_t c = 0u;
      #pragma unroll
      for( int w = 0; w < 8; ++w ) { c += s_w[threadIdx.x][w]; }
      counts[threadIdx.x * gridDim.x + blockIdx.x] = c;
    }
    return;
  }
  const unsigned long long base = *base_dev;
  const V3d x0 = ( mask != 0ull ) ? load_v3( q0, b ) : v3( 0.0, 0.0, 0.0 );
  for( uint32_t g = 0; g < ng; ++g )
  {
    if( counts[g * gridDim.x + blockIdx.x] == 0u ) { continue; } // uniform across the block
    const bool hit = ( ( mask >> g ) & 1ull ) != 0ull;
    const unsigned bal = __ballot_sync( 0xffffffffu, hit ); // every lane of the warp is here
    if( !hit ) { continue; }
    uint32_t before = __popc( bal & ( ( 1u << lane ) - 1u ) );
    for( int w = 0; w < warp; ++w ) { before += s_w[g][w]; }
    const unsigned long long k = base + offsets[g * gridDim.x + blockIdx.x] + before;
    if( g < planes.n )
    {
      const V3d xp = v3( planes.x[g][0], planes.x[g][1], planes.x[g][2] );
      const V3d np = v3( planes.nrm[g][0], planes.nrm[g][1], planes.nrm[g][2] );
      const double d = dot3( np, x1 - xp );
      put_contact( out, k, SG_PLANE_SPHERE, b, g, 0u, np, x0 - r * np, fmin( 0.0, d - r ) );
    }
    else
    {
      const uint32_t cy = g - planes.n;
      const V3d xc = v3( planes.cx[cy][0], planes.cx[cy][1], planes.cx[cy][2] );
      const V3d ax = v3( planes.cax[cy][0], planes.cax[cy][1], planes.cax[cy][2] );
      const V3d e0 = ( x0 - xc ) - dot3( ax, x0 - xc ) * ax;
      const V3d n = normalized3( -e0 );
      const V3d d1 = ( x1 - xc ) - dot3( ax, x1 - xc ) * ax;
      // computePenetrationDepth( q1 ) = min( 0, R - |d| - r ) (StaticCylinderSphereConstraint.cpp:311-317)
      put_contact( out, k, SG_CYLINDER_SPHERE, b, cy, 0u, n, x0 - r * n, fmin( 0.0, planes.cr[cy] - sqrt( dot3( d1, d1 ) ) - r ) );
    }
  }
}

// small helper kernels
__global__ void k_rb3d_sum_totals( const ScanPairCounts::Acc* __restrict__ bp_totals, const unsigned long long* __restrict__ narrow_total, const int fused, unsigned long long* __restrict__ out3 )
{
  // out3 = { P_c, n_body_body, unused }
  out3[0] = bp_totals->c;
  out3[1] = fused ? bp_totals->a : *narrow_total;
}

struct ScanU32To64
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

// ---- portals (rigidbody3d/Portals/PlanarPortal.h), sphere scenes: kernels, per-context data ----------------
#include "sg_rb3d_portal_kernels.cuh"
#include "sg_pair_sort_host.cuh"

struct Rb3dPortalData
{
  SgPortals3D portals;
  DevBuf rboxes;                                   // double[6 n]: boxes at q1 (the touch tests read them before n + T is known)
  DevBuf tflag, toff, t_partials, ttotal;          // u32[P n] flags / teleported box numbers; T
  DevBuf box_body, box_portal;                     // u32[T]: TeleportedBody table
  DevBuf reg_cnt, tel_cnt, tel_off, pr_partials, tel_total;
  DevBuf reg_off, reg_total;                       // u64 per candidate; number of un-teleported pairs
  DevBuf reg_pairs;                                // uint2[]: the un-teleported candidates, in order
  DevBuf tc_key, tc_idx, tc_info, uflag, uoff, u_partials, utotal;
  DevBuf x0t, x1t, tp0, tp1;                       // per teleported contact: constructor arguments
  PinBuf h, h_tele;
  uint64_t n_boxes = 0, n_reg = 0, n_tel = 0;
  bool result = false;                             // the last active set came from the portal path
  Rb3dPortalData() { memset( &portals, 0, sizeof( portals ) ); }
  void release()
  {
    DevBuf* bufs[] = { &rboxes, &tflag, &toff, &t_partials, &ttotal, &box_body, &box_portal, &reg_cnt, &tel_cnt, &tel_off, &pr_partials, &tel_total, &reg_off, &reg_total, &reg_pairs,
                       &tc_key, &tc_idx, &tc_info, &uflag, &uoff, &u_partials, &utotal, &x0t, &x1t, &tp0, &tp1 };
    for( DevBuf* b : bufs ) { b->release(); }
    h.release(); h_tele.release();
  }
};

// ---- host side -------------------------------------------------------------------------------------
struct MeshHost
{
  DevBuf verts, samples, hull, sdf;
  MeshDev dev;
};

struct Rb3dData
{
  uint32_t n = 0;
  bool all_spheres = false;
  bool has_free_box = false; // a box that is not kinematically scripted (static cylinders reject those)
  bool flow_resident = false; // q0 (as given) and q1 (as computed) of the last sg_rb3d_flow are still on the device
  bool q1_valid = false;      // q1 on the device is the output of the last flow kernel (sg_rb3d_flow or sg_rb3d_step)
  DevBuf minertia;            // double[18 n]: I blocks, then Iinv blocks (sg_rb3d_update_m_and_minv)
  double g[3] = { 0.0, 0.0, 0.0 };
  Planes3D planes;
  // geometry list (host copy) and per-body expansion
  std::vector<uint32_t> geo_type, geo_mesh;
  std::vector<double> geo_r, geo_half;
  std::vector<uint32_t> h_geo_of_body; // host copies of sg_rb3d_set_bodies' tables (the state snapshot writes them back)
  std::vector<uint8_t> h_fixed;
  std::vector<MeshHost*> meshes;
  std::vector<std::vector<unsigned char>> mesh_record; // per mesh: its RigidBodyTriangleMesh::serialize record (sg_rb3d_set_mesh_snapshot), written back verbatim by the state snapshot
  DevBuf d_meshes; // MeshDev[]
  DevBuf mesh_stats;
  DevBuf btype, bparam, bmesh, radius, flags, mass, I0;
  DevBuf q0, v0, q1, v1, boxes;
  BroadScratch bp;
  // narrow phase over the candidate list
  DevBuf pair_counts, pair_offsets, pair_partials, narrow_total, bad_flag, npairs_dev;
  // planes
  DevBuf st_counts, st_offsets, st_partials, st_total, totals3;
  // contacts
  DevBuf c_type, c_i, c_j, c_aux, c_n, c_p, c_depth;
  uint64_t act_cap = 0;
  PinBuf h_totals, h_out;
  uint64_t n_cand = 0, n_bb = 0, n_static = 0;
  bool have_result = false, cand_valid = false;
  Rb3dPortalData* px = nullptr; // allocated by sg_rb3d_set_portals
  SlabComm slab;                // multi-GPU slab mode (all-sphere scenes): slots [0, n_owned) are this rank's bodies, 2 x ghost_cap halo slots follow
  Rb3dData() { memset( &planes, 0, sizeof( planes ) ); }
};

void sg_rb3d_release( sg_ctx* ctx )
{
  Rb3dData* d = ctx->rb3d;
  if( d == nullptr ) { return; }
  for( MeshHost* m : d->meshes ) { m->verts.release(); m->samples.release(); m->hull.release(); m->sdf.release(); delete m; }
  DevBuf* bufs[] = { &d->d_meshes, &d->mesh_stats, &d->btype, &d->bparam, &d->bmesh, &d->radius, &d->flags, &d->mass, &d->I0, &d->q0, &d->v0, &d->q1, &d->v1, &d->boxes,
                     &d->pair_counts, &d->pair_offsets, &d->pair_partials, &d->narrow_total, &d->bad_flag, &d->npairs_dev,
                     &d->st_counts, &d->st_offsets, &d->st_partials, &d->st_total, &d->totals3,
                     &d->c_type, &d->c_i, &d->c_j, &d->c_aux, &d->c_n, &d->c_p, &d->c_depth, &d->minertia };
  for( DevBuf* b : bufs ) { b->release(); }
  d->bp.release();
  d->h_totals.release(); d->h_out.release();
  d->slab.release();
  if( d->px != nullptr ) { d->px->release(); delete d->px; d->px = nullptr; }
  delete d;
  ctx->rb3d = nullptr;
}

static Rb3dData* rb3d_data( sg_ctx* ctx )
{
  if( ctx->rb3d == nullptr ) { ctx->rb3d = new Rb3dData; }
  return ctx->rb3d;
}

static Rb3dDev rb3d_dev( const Rb3dData* d )
{
  Rb3dDev dev;
  dev.n = d->n;
  dev.n_live = d->slab.on ? d->slab.n_owned : d->n;
  dev.btype = d->btype.as<uint32_t>();
  dev.bparam = d->bparam.as<double>();
  dev.bmesh = d->bmesh.as<uint32_t>();
  dev.meshes = d->d_meshes.as<MeshDev>();
  dev.mesh_stats = d->mesh_stats.as<unsigned long long>();
  return dev;
}

static ContactOut3D rb3d_out( const Rb3dData* d )
{
  ContactOut3D out;
  out.type = d->c_type.as<uint32_t>(); out.i = d->c_i.as<uint32_t>(); out.j = d->c_j.as<uint32_t>(); out.aux = d->c_aux.as<uint32_t>();
  out.n = d->c_n.as<double>(); out.p = d->c_p.as<double>(); out.depth = d->c_depth.as<double>();
  out.cap = d->act_cap;
  if( d->slab.on ) { out.gid.gid = d->slab.gid.as<uint32_t>(); }
  return out;
}

static int rb3d_ensure_outputs( sg_ctx* ctx, Rb3dData* d, const uint64_t cand_cap, const uint64_t act_cap )
{
  if( cand_cap > d->bp.cand_cap )
  {
    SG_CUDA( ctx, d->bp.cand.ensure( size_t( cand_cap ) * sizeof( uint2 ) ) );
    d->bp.cand_cap = d->bp.cand.cap / sizeof( uint2 );
  }
  if( act_cap > d->act_cap )
  {
    SG_CUDA( ctx, d->c_type.ensure( size_t( act_cap ) * 4 ) );
    SG_CUDA( ctx, d->c_i.ensure( size_t( act_cap ) * 4 ) );
    SG_CUDA( ctx, d->c_j.ensure( size_t( act_cap ) * 4 ) );
    SG_CUDA( ctx, d->c_aux.ensure( size_t( act_cap ) * 4 ) );
    SG_CUDA( ctx, d->c_n.ensure( size_t( act_cap ) * 24 ) );
    SG_CUDA( ctx, d->c_p.ensure( size_t( act_cap ) * 24 ) );
    SG_CUDA( ctx, d->c_depth.ensure( size_t( act_cap ) * 8 ) );
    d->act_cap = act_cap;
  }
  return SG_OK;
}

static int rb3d_flow_device( sg_ctx* ctx, Rb3dData* d, const int map_kind, const double dt )
{
  const uint32_t n = d->n;
  if( n == 0 ) { return SG_OK; }
  d->q1_valid = true;
  const uint32_t nrun = d->slab.on ? d->slab.n_owned : n;
  SG_LAUNCH( ctx, "rb3d_flow", double( nrun ) * 336.0, k_rb3d_flow<<<sg_div_up( nrun > 0 ? nrun : 1, 128 ), 128, 0, ctx->stream>>>( map_kind, n, nrun, d->q0.as<double>(), d->v0.as<double>(), d->mass.as<double>(), d->I0.as<double>(),
             d->btype.as<uint32_t>(), d->g[0], d->g[1], d->g[2], dt, d->q1.as<double>(), d->v1.as<double>() ) );
  return SG_OK;
}

// body-plane contacts appended after the body-body ones; leaves the static total in st_total
static int rb3d_planes_device( sg_ctx* ctx, Rb3dData* d, const bool emit )
{
  const uint32_t n = d->n, np = d->planes.n + d->planes.ncyl;
  if( np == 0 || n == 0 ) { return SG_OK; }
  if( d->planes.ncyl > 0u && d->has_free_box )
  {
    return sg_fail( ctx, SG_ERR_UNSUPPORTED, "Collision between static cylinders and box not supported." ); // RigidBody3DSim.cpp:1550-1553
  }
  const unsigned nblk = sg_div_up( n, 256 );
  const uint32_t nst = np * nblk;
  const Rb3dDev dev = rb3d_dev( d );
  if( !emit )
  {
    SG_CUDA( ctx, d->st_counts.ensure( size_t( nst ) * 4 + 4 ) );
    SG_CUDA( ctx, d->st_offsets.ensure( size_t( nst ) * 4 + 4 ) );
    SG_CUDA( ctx, d->st_partials.ensure( ( size_t( nst ) / SG_SCAN_TILE + 2 ) * 4 ) );
    if( d->all_spheres )
    {
      SG_LAUNCH( ctx, "rb3d_plane_count", double( dev.n_live ) * 36.0, k_rb3d_static_spheres<false><<<nblk, 256, 0, ctx->stream>>>( dev.n_live, d->planes, d->q0.as<double>(), d->q1.as<double>(), d->radius.as<double>(),
                 d->flags.as<uint32_t>(), d->st_counts.as<uint32_t>(), nullptr, nullptr, rb3d_out( d ) ) );
    }
    else
    {
      SG_LAUNCH( ctx, "rb3d_plane_count", double( n ) * 32.0, k_rb3d_plane_count<<<nblk, 256, 0, ctx->stream>>>( dev, d->planes, d->q0.as<double>(), d->q1.as<double>(), d->st_counts.as<uint32_t>() ) );
    }
    return sg_exclusive_scan<ScanU32>( ctx, "rb3d_plane_scan", d->st_counts.as<uint32_t>(), nullptr, nst, nst, d->st_partials.as<uint32_t>(), d->st_offsets.as<uint32_t>(), d->st_total.as<uint32_t>(), false );
  }
  if( d->all_spheres )
  {
    SG_LAUNCH( ctx, "rb3d_plane_emit", double( n ) * 4.0, k_rb3d_static_spheres<true><<<nblk, 256, 0, ctx->stream>>>( dev.n_live, d->planes, d->q0.as<double>(), d->q1.as<double>(), d->radius.as<double>(),
               d->flags.as<uint32_t>(), d->st_counts.as<uint32_t>(), d->st_offsets.as<uint32_t>(), d->totals3.as<unsigned long long>() + 1, rb3d_out( d ) ) );
    return SG_OK;
  }
  SG_LAUNCH( ctx, "rb3d_plane_emit", double( n ) * 4.0, k_rb3d_plane_emit<<<nblk, 256, 0, ctx->stream>>>( dev, d->planes, d->q0.as<double>(), d->q1.as<double>(), d->st_counts.as<uint32_t>(), d->st_offsets.as<uint32_t>(),
             d->totals3.as<unsigned long long>() + 1, rb3d_out( d ) ) );
  return SG_OK;
}

static int rb3d_portal_active_set_device( sg_ctx* ctx, Rb3dData* d );

static int rb3d_active_set_device( sg_ctx* ctx, Rb3dData* d, const bool want_cand_in )
{
  const uint32_t n = d->n;
  // with portals the body-body path grows teleported boxes and collisions (RigidBody3DSim.cpp:1088-1115, 1139-1199)
  if( d->px != nullptr && d->px->portals.n > 0u ) { return rb3d_portal_active_set_device( ctx, d ); }
  if( d->px != nullptr ) { d->px->result = false; }
  d->n_cand = d->n_bb = d->n_static = 0;
  d->have_result = true;
  if( n == 0 ) { d->cand_valid = want_cand_in; return SG_OK; }
  const bool fused = d->all_spheres;
  const bool want_cand = want_cand_in || !fused; // the generic narrow phase consumes the candidate list
  d->cand_valid = want_cand;
  SG_CUDA( ctx, d->h_totals.ensure( 64 ) );
  SG_CUDA( ctx, d->totals3.ensure( 32 ) );
  SG_CUDA( ctx, d->st_total.ensure( 4 ) );
  SG_CUDA( ctx, d->narrow_total.ensure( 8 ) );
  SG_CUDA( ctx, d->bad_flag.ensure( 4 ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->st_total.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->narrow_total.ptr, 0, 8, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->bad_flag.ptr, 0, 4, ctx->stream ) );
  const Rb3dDev dev = rb3d_dev( d );
  int rc;
  unsigned long long* ht = d->h_totals.as<unsigned long long>();

  if( fused )
  {
    rc = sg_bp_prepare_scratch<Sphere3DPolicy>( ctx, d->bp, n );
    if( rc != SG_OK ) { return rc; }
    rc = rb3d_ensure_outputs( ctx, d, want_cand ? ( d->bp.cand_cap > 0 ? d->bp.cand_cap : uint64_t( n ) * 16u + 1024u ) : 0u, d->act_cap > 0 ? d->act_cap : uint64_t( n ) * 5u + 1024u );
    if( rc != SG_OK ) { return rc; }
    Sphere3DIn in;
    in.q0 = d->q0.as<double>(); in.q1 = d->q1.as<double>(); in.r = d->radius.as<double>(); in.flags = d->flags.as<uint32_t>(); in.n = n;
    in.n_owned = d->slab.on ? d->slab.n_owned : n; in.ghost_cap = d->slab.ghost_cap; in.ghost_counts = d->slab.on ? d->slab.ghost_counts.as<uint32_t>() : nullptr;
    d->bp.ord_by_index = d->slab.on ? d->slab.gid.as<uint32_t>() : nullptr;
    rc = sg_bp_bin_and_count<Sphere3DPolicy>( ctx, d->bp, in );
    if( rc != SG_OK ) { return rc; }
    rc = rb3d_planes_device( ctx, d, false );
    if( rc != SG_OK ) { return rc; }
    SG_LAUNCH( ctx, "rb3d_totals", 0.0, k_rb3d_sum_totals<<<1, 1, 0, ctx->stream>>>( d->bp.totals.as<ScanPairCounts::Acc>(), d->narrow_total.as<unsigned long long>(), 1, d->totals3.as<unsigned long long>() ) );
    for( int attempt = 0; attempt < 2; ++attempt )
    {
      rc = sg_bp_emit_lists<Sphere3DPolicy>( ctx, d->bp, in, n, want_cand, rb3d_out( d ), d->act_cap );
      if( rc != SG_OK ) { return rc; }
      rc = rb3d_planes_device( ctx, d, true );
      if( rc != SG_OK ) { return rc; }
      ht[2] = 0ull;
      SG_CUDA( ctx, cudaMemcpyAsync( ht, d->totals3.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( ht + 2, d->st_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
      SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
      sg_prof_collect( ctx );
      d->n_cand = ht[0]; d->n_bb = ht[1]; d->n_static = ht[2] & 0xffffffffull;
      if( ctx->profile )
      {
        // list sizes are only known now: add the emitted bytes to the kernels that wrote them
        ctx->prof[sg_prof_entry( ctx, "bp_emit" )].bytes += ( want_cand ? double( d->n_cand ) * 8.0 : 0.0 ) + double( d->n_bb ) * 8.0;
        ctx->prof[sg_prof_entry( ctx, "bp_contacts" )].bytes += double( d->n_bb ) * ( 8.0 + 76.0 ) + double( n ) * 64.0; // work item in, SoA contact out (type, i, j, aux, n, p, depth); every record needed once
      }
      const uint64_t need = d->n_bb + d->n_static;
      if( ( !want_cand || d->n_cand <= d->bp.cand_cap ) && need <= d->act_cap ) { break; }
      if( attempt == 1 ) { return sg_fail( ctx, SG_ERR_INTERNAL, "rb3d: output lists still overflow after regrowth" ); }
      rc = rb3d_ensure_outputs( ctx, d, want_cand ? d->n_cand + d->n_cand / 8 + 1024 : 0u, need + need / 8 + 1024 );
      if( rc != SG_OK ) { return rc; }
    }
    return SG_OK;
  }

  // ---- generic pipeline ----
  SG_CUDA( ctx, d->boxes.ensure( size_t( n ) * 48 ) );
  SG_LAUNCH( ctx, "rb3d_aabb", double( n ) * ( 96.0 + 48.0 ), k_rb3d_aabb<<<sg_div_up( n, 128 ), 128, 0, ctx->stream>>>( dev, d->q1.as<double>(), d->boxes.as<double>() ) );
  rc = sg_bp_prepare_scratch<Box3DPolicy>( ctx, d->bp, n );
  if( rc != SG_OK ) { return rc; }
  Box3DIn in;
  in.boxes = d->boxes.as<double>(); in.n = n;
  rc = sg_bp_bin_and_count<Box3DPolicy>( ctx, d->bp, in );
  if( rc != SG_OK ) { return rc; }
  // the candidate list must exist before the narrow phase can be sized: one small read-back
  SG_CUDA( ctx, cudaMemcpyAsync( ht, d->bp.totals.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  const uint64_t np = ht[0];
  uint32_t mesh_tiles = SDF_TILES_DEFAULT;
  d->n_cand = np;
  rc = rb3d_ensure_outputs( ctx, d, np + 64, d->act_cap );
  if( rc != SG_OK ) { return rc; }
  if( np > 0 )
  {
    rc = sg_bp_emit_lists<Box3DPolicy>( ctx, d->bp, in, n, true, NoOut3D{}, 0u );
    if( rc != SG_OK ) { return rc; }
  }
  if( np >= 0xffffffffull ) { return sg_fail( ctx, SG_ERR_INTERNAL, "rb3d: more than 2^32 candidate pairs in the generic pipeline" ); }
  SG_CUDA( ctx, d->pair_counts.ensure( size_t( np ) * 4 + 4 ) );
  SG_CUDA( ctx, d->pair_offsets.ensure( size_t( np ) * 8 + 8 ) );
  SG_CUDA( ctx, d->pair_partials.ensure( ( size_t( np ) / SG_SCAN_TILE + 2 ) * 8 ) );
  const unsigned long long* npairs_dev = &d->bp.totals.as<ScanPairCounts::Acc>()->c;
  if( np > 0 )
  {
    SG_CUDA( ctx, cudaMemsetAsync( d->pair_counts.ptr, 0, size_t( np ) * 4, ctx->stream ) );
    SG_LAUNCH( ctx, "rb3d_pairs_count", double( np ) * 250.0, k_rb3d_pairs<false><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( dev, d->bp.cand.as<uint2>(), npairs_dev, d->q0.as<double>(), d->q1.as<double>(),
               d->pair_counts.as<uint32_t>(), nullptr, rb3d_out( d ), d->bad_flag.as<uint32_t>() ) );
    if( !d->meshes.empty() )
    {
      const unsigned grid = unsigned( np < uint64_t( ctx->num_sms ) * 8u ? np : uint64_t( ctx->num_sms ) * 8u );
      SG_CUDA( ctx, cudaFuncSetAttribute( k_rb3d_mesh_pairs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_TILES_MAX * SDF_TILE_ELEMS * 8 ) );
      SG_CUDA( ctx, cudaFuncSetAttribute( k_rb3d_mesh_pairs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_TILES_MAX * SDF_TILE_ELEMS * 8 ) );
      const char* tiles_env = getenv( "SG_RB3D_TILES" ); // tuning knob: brick capacity per CTA in 4 KB tiles
      mesh_tiles = tiles_env != nullptr ? uint32_t( atoi( tiles_env ) ) : SDF_TILES_DEFAULT;
      mesh_tiles = mesh_tiles < 1u ? 1u : ( mesh_tiles > SDF_TILES_MAX ? SDF_TILES_MAX : mesh_tiles );
      SG_LAUNCH( ctx, "rb3d_mesh_count", 0.0, k_rb3d_mesh_pairs<false><<<grid, 256, mesh_tiles * SDF_TILE_ELEMS * 8, ctx->stream>>>( dev, d->bp.cand.as<uint2>(), npairs_dev, d->q1.as<double>(), d->pair_counts.as<uint32_t>(), nullptr, rb3d_out( d ), mesh_tiles ) );
    }
    rc = sg_exclusive_scan<ScanU32To64>( ctx, "rb3d_pair_scan", d->pair_counts.as<uint32_t>(), nullptr, uint32_t( np ), uint32_t( np ), d->pair_partials.as<unsigned long long>(), d->pair_offsets.as<unsigned long long>(),
                                         d->narrow_total.as<unsigned long long>(), false );
    if( rc != SG_OK ) { return rc; }
  }
  rc = rb3d_planes_device( ctx, d, false );
  if( rc != SG_OK ) { return rc; }
  SG_LAUNCH( ctx, "rb3d_totals", 0.0, k_rb3d_sum_totals<<<1, 1, 0, ctx->stream>>>( d->bp.totals.as<ScanPairCounts::Acc>(), d->narrow_total.as<unsigned long long>(), 0, d->totals3.as<unsigned long long>() ) );
  // sizes to the host, then the emits
  uint32_t* hbad = reinterpret_cast<uint32_t*>( ht + 4 );
  ht[2] = 0ull;
  SG_CUDA( ctx, cudaMemcpyAsync( ht, d->totals3.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( ht + 2, d->st_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( hbad, d->bad_flag.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  if( *hbad != 0u )
  {
    return sg_fail( ctx, SG_ERR_UNSUPPORTED, "collision between two different geometry types is not supported (the reference exits here: rigidbody3d/RigidBody3DSim.cpp:905-961)" );
  }
  d->n_bb = ht[1]; d->n_static = ht[2] & 0xffffffffull;
  const uint64_t need = d->n_bb + d->n_static;
  rc = rb3d_ensure_outputs( ctx, d, 0u, need + 64 );
  if( rc != SG_OK ) { return rc; }
  if( np > 0 && d->n_bb > 0 )
  {
    SG_LAUNCH( ctx, "rb3d_pairs_emit", double( np ) * 250.0 + double( d->n_bb ) * 72.0, k_rb3d_pairs<true><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( dev, d->bp.cand.as<uint2>(), npairs_dev, d->q0.as<double>(), d->q1.as<double>(),
               nullptr, d->pair_offsets.as<unsigned long long>(), rb3d_out( d ), d->bad_flag.as<uint32_t>() ) );
    if( !d->meshes.empty() )
    {
      const unsigned grid = unsigned( np < uint64_t( ctx->num_sms ) * 8u ? np : uint64_t( ctx->num_sms ) * 8u );
      SG_LAUNCH( ctx, "rb3d_mesh_emit", 0.0, k_rb3d_mesh_pairs<true><<<grid, 256, mesh_tiles * SDF_TILE_ELEMS * 8, ctx->stream>>>( dev, d->bp.cand.as<uint2>(), npairs_dev, d->q1.as<double>(), nullptr, d->pair_offsets.as<unsigned long long>(), rb3d_out( d ), mesh_tiles ) );
    }
  }
  rc = rb3d_planes_device( ctx, d, true );
  if( rc != SG_OK ) { return rc; }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

// RigidBody3DSim::computeActiveSet with portals, all-sphere scenes (RigidBody3DSim.cpp:250-262 -> :1072-1260): boxes at q1, teleported
// copies, un-teleported candidates through the regular narrow phase (k_rb3d_pairs), TeleportedCollision set for the rest, then planes
// and cylinders.  Same sequence as the rigidbody2d portal driver (sg_rb2d.cu).
static int rb3d_portal_active_set_device( sg_ctx* ctx, Rb3dData* d )
{
  Rb3dPortalData* x = d->px;
  const uint32_t n = d->n;
  const uint32_t np_portals = x->portals.n;
  d->n_cand = d->n_bb = d->n_static = 0;
  x->n_boxes = x->n_reg = x->n_tel = 0;
  d->have_result = true;
  d->cand_valid = true;
  x->result = true;
  if( n == 0 ) { return SG_OK; }
  if( !d->all_spheres )
  {
    return sg_fail( ctx, SG_ERR_UNSUPPORTED, "rigidbody3d portals are implemented for all-sphere scenes (the reference supports teleported collisions for spheres only: rigidbody3d/RigidBody3DSim.cpp:1262-1292)" );
  }
  if( uint64_t( n ) * np_portals >= 0x80000000ull ) { return sg_fail( ctx, SG_ERR_INVALID, "rigidbody3d portals: bodies x portals must stay below 2^31" ); }
  const Rb3dDev dev = rb3d_dev( d );
  const unsigned nblk = sg_div_up( n, 256 );
  const uint32_t nflag = n * np_portals;
  SG_CUDA( ctx, x->h.ensure( 128 ) );
  SG_CUDA( ctx, d->totals3.ensure( 32 ) ); SG_CUDA( ctx, d->st_total.ensure( 4 ) ); SG_CUDA( ctx, d->narrow_total.ensure( 8 ) ); SG_CUDA( ctx, d->bad_flag.ensure( 4 ) );
  SG_CUDA( ctx, x->ttotal.ensure( 4 ) ); SG_CUDA( ctx, x->reg_total.ensure( 8 ) ); SG_CUDA( ctx, x->tel_total.ensure( 4 ) ); SG_CUDA( ctx, x->utotal.ensure( 4 ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->st_total.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->narrow_total.ptr, 0, 8, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->bad_flag.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->ttotal.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->reg_total.ptr, 0, 8, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->tel_total.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( x->utotal.ptr, 0, 4, ctx->stream ) );
  SG_CUDA( ctx, x->rboxes.ensure( size_t( n ) * 48 ) );
  SG_CUDA( ctx, x->tflag.ensure( size_t( nflag ) * 4 + 4 ) ); SG_CUDA( ctx, x->toff.ensure( size_t( nflag ) * 4 + 4 ) );
  SG_CUDA( ctx, x->t_partials.ensure( ( size_t( nflag ) / SG_SCAN_TILE + 2 ) * 4 ) );
  SG_LAUNCH( ctx, "rb3d_aabb", double( n ) * ( 96.0 + 48.0 ), k_rb3d_aabb<<<sg_div_up( n, 128 ), 128, 0, ctx->stream>>>( dev, d->q1.as<double>(), x->rboxes.as<double>() ) );
  SG_LAUNCH( ctx, "r3p_touch", double( nflag ) * 52.0, k_r3p_touch<<<dim3( nblk, np_portals ), 256, 0, ctx->stream>>>( x->portals, n, x->rboxes.as<double>(), x->tflag.as<uint32_t>() ) );
  int rc = sg_exclusive_scan<ScanU32>( ctx, "r3p_touch_scan", x->tflag.as<uint32_t>(), nullptr, nflag, nflag, x->t_partials.as<uint32_t>(), x->toff.as<uint32_t>(), x->ttotal.as<uint32_t>(), false );
  if( rc != SG_OK ) { return rc; }
  // planes and cylinders do not depend on the portals: count them now (their emit follows the body-body contacts)
  rc = rb3d_planes_device( ctx, d, false );
  if( rc != SG_OK ) { return rc; }
  uint32_t* h32 = x->h.as<uint32_t>();
  unsigned long long* h64 = x->h.as<unsigned long long>() + 8; // bytes 64...
  h32[1] = 0u;
  SG_CUDA( ctx, cudaMemcpyAsync( h32, x->ttotal.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( h32 + 1, d->st_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  const uint32_t nt = h32[0];
  d->n_static = h32[1];
  x->n_boxes = nt;
  if( uint64_t( n ) + nt >= 0x80000000ull ) { return sg_fail( ctx, SG_ERR_INVALID, "rigidbody3d portals: more than 2^31 - 1 boxes" ); }
  const uint32_t next = n + nt;
  SG_CUDA( ctx, d->boxes.ensure( size_t( next ) * 48 ) );
  SG_CUDA( ctx, x->box_body.ensure( size_t( nt ) * 4 + 4 ) ); SG_CUDA( ctx, x->box_portal.ensure( size_t( nt ) * 4 + 4 ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->boxes.ptr, x->rboxes.ptr, size_t( n ) * 48, cudaMemcpyDeviceToDevice, ctx->stream ) );
  if( nt > 0 )
  {
    SG_LAUNCH( ctx, "r3p_tele_boxes", double( nflag ) * 8.0 + double( nt ) * 140.0, k_r3p_tele_boxes<<<dim3( nblk, np_portals ), 256, 0, ctx->stream>>>( x->portals, dev, d->q1.as<double>(), x->rboxes.as<double>(),
               x->tflag.as<uint32_t>(), x->toff.as<uint32_t>(), d->boxes.as<double>(), x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>() ) );
  }
  rc = sg_bp_prepare_scratch<Box3DPolicy>( ctx, d->bp, next );
  if( rc != SG_OK ) { return rc; }
  Box3DIn in;
  in.boxes = d->boxes.as<double>(); in.n = next;
  rc = sg_bp_bin_and_count<Box3DPolicy>( ctx, d->bp, in );
  if( rc != SG_OK ) { return rc; }
  SG_CUDA( ctx, cudaMemcpyAsync( h64, d->bp.totals.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  const uint64_t np = h64[0];
  d->n_cand = np;
  if( np >= 0xffffffffull ) { return sg_fail( ctx, SG_ERR_INTERNAL, "rigidbody3d portals: more than 2^32 candidate pairs" ); }
  rc = rb3d_ensure_outputs( ctx, d, np + 64, d->act_cap );
  if( rc != SG_OK ) { return rc; }
  uint64_t nreg_pairs = 0;
  uint32_t nraw = 0;
  if( np > 0 )
  {
    rc = sg_bp_emit_lists<Box3DPolicy>( ctx, d->bp, in, next, true, NoOut3D{}, 0u );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, x->reg_cnt.ensure( size_t( np ) * 4 + 4 ) ); SG_CUDA( ctx, x->reg_off.ensure( size_t( np ) * 8 + 8 ) );
    SG_CUDA( ctx, x->tel_cnt.ensure( size_t( np ) * 4 + 4 ) ); SG_CUDA( ctx, x->tel_off.ensure( size_t( np ) * 4 + 4 ) );
    SG_CUDA( ctx, x->pr_partials.ensure( ( size_t( np ) / SG_SCAN_TILE + 2 ) * 8 ) );
    SG_LAUNCH( ctx, "r3p_classify_count", double( np ) * 80.0, k_r3p_classify<false><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( x->portals, dev, d->bp.cand.as<uint2>(), np, d->q1.as<double>(),
               x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>(), x->reg_cnt.as<uint32_t>(), x->tel_cnt.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr ) );
    rc = sg_exclusive_scan<ScanU32To64>( ctx, "r3p_regular_scan", x->reg_cnt.as<uint32_t>(), nullptr, uint32_t( np ), uint32_t( np ), x->pr_partials.as<unsigned long long>(), x->reg_off.as<unsigned long long>(),
                                         x->reg_total.as<unsigned long long>(), false );
    if( rc != SG_OK ) { return rc; }
    rc = sg_exclusive_scan<ScanU32>( ctx, "r3p_teleported_scan", x->tel_cnt.as<uint32_t>(), nullptr, uint32_t( np ), uint32_t( np ), x->pr_partials.as<uint32_t>(), x->tel_off.as<uint32_t>(), x->tel_total.as<uint32_t>(), false );
    if( rc != SG_OK ) { return rc; }
    SG_CUDA( ctx, cudaMemcpyAsync( h64 + 2, x->reg_total.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h32 + 2, x->tel_total.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    nreg_pairs = h64[2];
    nraw = h32[2];
  }
  if( nraw > 0x40000000u ) { return sg_fail( ctx, SG_ERR_INTERNAL, "rigidbody3d portals: more than 2^30 teleported collisions" ); }
  uint32_t m = 1u;
  while( m < nraw ) { m <<= 1; }
  SG_CUDA( ctx, x->reg_pairs.ensure( size_t( nreg_pairs ) * 8 + 8 ) );
  if( nraw > 0 )
  {
    SG_CUDA( ctx, x->tc_key.ensure( size_t( m ) * 8 ) ); SG_CUDA( ctx, x->tc_idx.ensure( size_t( m ) * 4 ) ); SG_CUDA( ctx, x->tc_info.ensure( size_t( nraw ) * 16 ) );
    SG_CUDA( ctx, x->uflag.ensure( size_t( nraw ) * 4 + 4 ) ); SG_CUDA( ctx, x->uoff.ensure( size_t( nraw ) * 4 + 4 ) );
    SG_CUDA( ctx, x->u_partials.ensure( ( size_t( nraw ) / SG_SCAN_TILE + 2 ) * 4 ) );
    SG_CUDA( ctx, x->x0t.ensure( size_t( nraw ) * 24 ) ); SG_CUDA( ctx, x->x1t.ensure( size_t( nraw ) * 24 ) );
    SG_CUDA( ctx, x->tp0.ensure( size_t( nraw ) * 4 ) ); SG_CUDA( ctx, x->tp1.ensure( size_t( nraw ) * 4 ) );
  }
  if( np > 0 )
  {
    SG_LAUNCH( ctx, "r3p_classify_emit", double( np ) * 28.0, k_r3p_classify<true><<<sg_div_up( np, 128 ), 128, 0, ctx->stream>>>( x->portals, dev, d->bp.cand.as<uint2>(), np, d->q1.as<double>(),
               x->box_body.as<uint32_t>(), x->box_portal.as<uint32_t>(), x->reg_cnt.as<uint32_t>(), x->tel_cnt.as<uint32_t>(), x->reg_off.as<unsigned long long>(), x->tel_off.as<uint32_t>(),
               x->reg_pairs.as<uint2>(), x->tc_key.as<unsigned long long>(), x->tc_idx.as<uint32_t>(), x->tc_info.as<uint4>() ) );
  }
  // regular narrow phase over the un-teleported candidates: count -> scan (the emit follows once the contact arrays are sized)
  SG_CUDA( ctx, d->pair_counts.ensure( size_t( nreg_pairs ) * 4 + 4 ) );
  SG_CUDA( ctx, d->pair_offsets.ensure( size_t( nreg_pairs ) * 8 + 8 ) );
  SG_CUDA( ctx, d->pair_partials.ensure( ( size_t( nreg_pairs ) / SG_SCAN_TILE + 2 ) * 8 ) );
  const unsigned long long* nreg_dev = x->reg_total.as<unsigned long long>();
  if( nreg_pairs > 0 )
  {
    SG_CUDA( ctx, cudaMemsetAsync( d->pair_counts.ptr, 0, size_t( nreg_pairs ) * 4, ctx->stream ) );
    SG_LAUNCH( ctx, "rb3d_pairs_count", double( nreg_pairs ) * 250.0, k_rb3d_pairs<false><<<sg_div_up( nreg_pairs, 128 ), 128, 0, ctx->stream>>>( dev, x->reg_pairs.as<uint2>(), nreg_dev, d->q0.as<double>(), d->q1.as<double>(),
               d->pair_counts.as<uint32_t>(), nullptr, rb3d_out( d ), d->bad_flag.as<uint32_t>() ) );
    rc = sg_exclusive_scan<ScanU32To64>( ctx, "rb3d_pair_scan", d->pair_counts.as<uint32_t>(), nullptr, uint32_t( nreg_pairs ), uint32_t( nreg_pairs ), d->pair_partials.as<unsigned long long>(),
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
    return sg_fail( ctx, SG_ERR_UNSUPPORTED, "collision between two different geometry types is not supported (the reference exits here: rigidbody3d/RigidBody3DSim.cpp:905-961)" );
  }
  x->n_reg = h64[3];
  x->n_tel = h32[5];
  d->n_bb = x->n_reg + x->n_tel;
  rc = rb3d_ensure_outputs( ctx, d, 0u, d->n_bb + d->n_static + 64 );
  if( rc != SG_OK ) { return rc; }
  if( nreg_pairs > 0 && x->n_reg > 0 )
  {
    SG_LAUNCH( ctx, "rb3d_pairs_emit", double( nreg_pairs ) * 250.0 + double( x->n_reg ) * 72.0, k_rb3d_pairs<true><<<sg_div_up( nreg_pairs, 128 ), 128, 0, ctx->stream>>>( dev, x->reg_pairs.as<uint2>(), nreg_dev, d->q0.as<double>(),
               d->q1.as<double>(), nullptr, d->pair_offsets.as<unsigned long long>(), rb3d_out( d ), d->bad_flag.as<uint32_t>() ) );
  }
  if( x->n_tel > 0 )
  {
    SG_LAUNCH( ctx, "r3p_tele_contacts", double( nraw ) * 250.0, k_r3p_tele_contacts<<<sg_div_up( nraw, 128 ), 128, 0, ctx->stream>>>( x->portals, dev, nraw, x->tc_idx.as<uint32_t>(), x->uflag.as<uint32_t>(), x->uoff.as<uint32_t>(),
               x->tc_info.as<uint4>(), d->q0.as<double>(), x->n_reg, rb3d_out( d ), x->x0t.as<double>(), x->x1t.as<double>(), x->tp0.as<uint32_t>(), x->tp1.as<uint32_t>() ) );
  }
  // the plane / cylinder emit reads its base (all body-body contacts) from totals3[1]
  h64[4] = d->n_cand; h64[5] = d->n_bb;
  SG_CUDA( ctx, cudaMemcpyAsync( d->totals3.ptr, h64 + 4, 16, cudaMemcpyHostToDevice, ctx->stream ) );
  if( d->n_static > 0 )
  {
    rc = rb3d_planes_device( ctx, d, true );
    if( rc != SG_OK ) { return rc; }
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

static int rb3d_copy_out( sg_ctx* ctx, Rb3dData* d, const uint32_t flags, sg_contacts* out )
{
  memset( out, 0, sizeof( *out ) );
  out->dim = 3;
  out->n_candidates = d->n_cand;
  out->n_body_body = d->n_bb;
  out->n_plane = d->n_static;
  const uint64_t na = d->n_bb + d->n_static;
  out->n_active = na;
  const bool want_cand = ( flags & SG_OUT_CANDIDATES ) != 0u && d->cand_valid;
  auto al = []( size_t b ) { return ( b + 63 ) & ~size_t( 63 ); };
  size_t bytes = 64;
  const size_t o_type = bytes; bytes += al( na * 4 );
  const size_t o_i = bytes; bytes += al( na * 4 );
  const size_t o_j = bytes; bytes += al( na * 4 );
  const size_t o_aux = bytes; bytes += al( na * 4 );
  const size_t o_n = bytes; if( flags & SG_OUT_NORMALS ) { bytes += al( na * 24 ); }
  const size_t o_p = bytes; if( flags & SG_OUT_POINTS ) { bytes += al( na * 24 ); }
  const size_t o_d = bytes; if( flags & SG_OUT_DEPTHS ) { bytes += al( na * 8 ); }
  const size_t o_c = bytes; if( want_cand ) { bytes += al( d->n_cand * 8 ); }
  SG_CUDA( ctx, d->h_out.ensure( bytes ) );
  char* h = d->h_out.as<char>();
  if( na > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_type, d->c_type.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_i, d->c_i.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_j, d->c_j.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_aux, d->c_aux.ptr, na * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    if( flags & SG_OUT_NORMALS ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_n, d->c_n.ptr, na * 24, cudaMemcpyDeviceToHost, ctx->stream ) ); }
    if( flags & SG_OUT_POINTS ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_p, d->c_p.ptr, na * 24, cudaMemcpyDeviceToHost, ctx->stream ) ); }
    if( flags & SG_OUT_DEPTHS ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_d, d->c_depth.ptr, na * 8, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  }
  if( want_cand && d->n_cand > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( h + o_c, d->bp.cand.ptr, d->n_cand * 8, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  out->type = reinterpret_cast<const uint32_t*>( h + o_type );
  out->i = reinterpret_cast<const uint32_t*>( h + o_i );
  out->j = reinterpret_cast<const uint32_t*>( h + o_j );
  out->aux = reinterpret_cast<const uint32_t*>( h + o_aux );
  out->n = ( flags & SG_OUT_NORMALS ) ? reinterpret_cast<const double*>( h + o_n ) : nullptr;
  out->p = ( flags & SG_OUT_POINTS ) ? reinterpret_cast<const double*>( h + o_p ) : nullptr;
  out->depth = ( flags & SG_OUT_DEPTHS ) ? reinterpret_cast<const double*>( h + o_d ) : nullptr;
  out->cand_ij = want_cand ? reinterpret_cast<const uint32_t*>( h + o_c ) : nullptr;
  return SG_OK;
}

// expands the geometry list into per-body arrays (called when bodies or geometry change)
static int rb3d_expand_bodies( sg_ctx* ctx, Rb3dData* d, const uint32_t n, const uint32_t* geo_of_body, const uint8_t* fixed )
{
  std::vector<uint32_t> btype( n ), bmesh( n ), flags( n );
  std::vector<double> bparam( size_t( n ) * 4, 0.0 ), radius( n, 0.0 );
  bool all_spheres = n > 0;
  bool has_free_box = false;
  for( uint32_t b = 0; b < n; ++b )
  {
    const uint32_t gi = geo_of_body[b];
    if( gi >= d->geo_type.size() ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_bodies: body %u refers to geometry %u of %zu", b, gi, d->geo_type.size() ); }
    const uint32_t t = d->geo_type[gi];
    if( t != SG_GEO_BOX && t != SG_GEO_SPHERE && t != SG_GEO_MESH ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_set_bodies: geometry type %u is not supported", t ); }
    if( t == SG_GEO_MESH && d->geo_mesh[gi] >= d->meshes.size() ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_bodies: geometry %u refers to mesh %u of %zu", gi, d->geo_mesh[gi], d->meshes.size() ); }
    flags[b] = fixed[b] ? SG_FIXED_BIT : 0u;
    btype[b] = t | flags[b];
    bmesh[b] = d->geo_mesh[gi];
    if( t == SG_GEO_SPHERE ) { bparam[4 * size_t( b )] = d->geo_r[gi]; radius[b] = d->geo_r[gi]; }
    else if( t == SG_GEO_BOX ) { for( int k = 0; k < 3; ++k ) { bparam[4 * size_t( b ) + k] = d->geo_half[3 * size_t( gi ) + k]; } }
    all_spheres = all_spheres && t == SG_GEO_SPHERE;
    has_free_box = has_free_box || ( t == SG_GEO_BOX && !fixed[b] );
  }
  d->all_spheres = all_spheres;
  d->has_free_box = has_free_box;
  SG_CUDA( ctx, d->btype.ensure( size_t( n ) * 4 + 4 ) ); SG_CUDA( ctx, d->bmesh.ensure( size_t( n ) * 4 + 4 ) ); SG_CUDA( ctx, d->flags.ensure( size_t( n ) * 4 + 4 ) );
  SG_CUDA( ctx, d->bparam.ensure( size_t( n ) * 32 + 32 ) ); SG_CUDA( ctx, d->radius.ensure( size_t( n ) * 8 + 8 ) );
  if( n > 0 )
  {
    SG_CUDA( ctx, cudaMemcpyAsync( d->btype.ptr, btype.data(), size_t( n ) * 4, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->bmesh.ptr, bmesh.data(), size_t( n ) * 4, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->flags.ptr, flags.data(), size_t( n ) * 4, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->bparam.ptr, bparam.data(), size_t( n ) * 32, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->radius.ptr, radius.data(), size_t( n ) * 8, cudaMemcpyHostToDevice, ctx->stream ) );
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

extern "C"
{

int sg_rb3d_set_geometry( sg_ctx* ctx, uint32_t ngeo, const uint32_t* type, const double* r, const double* half, const uint32_t* mesh )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( ngeo > 0 && ( type == nullptr || r == nullptr || half == nullptr || mesh == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_geometry: null array" ); }
  Rb3dData* d = rb3d_data( ctx );
  d->geo_type.assign( type, type + ngeo );
  d->geo_r.assign( r, r + ngeo );
  d->geo_half.assign( half, half + 3 * size_t( ngeo ) );
  d->geo_mesh.assign( mesh, mesh + ngeo );
  return SG_OK;
}

int sg_rb3d_add_mesh( sg_ctx* ctx, uint32_t nverts, const double* verts, uint32_t nsamples, const double* samples, uint32_t nhull, const double* hull,
                      const double* cell_delta, const uint32_t* dims, const double* origin, const double* sdf, uint32_t* mesh_index )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( cell_delta == nullptr || dims == nullptr || origin == nullptr || sdf == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_add_mesh: null array" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  Rb3dData* d = rb3d_data( ctx );
  MeshHost* m = new MeshHost;
  d->meshes.push_back( m );
  const uint32_t pitch = ( dims[0] + 1u ) & ~1u; // even row length: TMA global strides must be multiples of 16 bytes
  const size_t ncell = size_t( pitch ) * dims[1] * dims[2];
  SG_CUDA( ctx, m->verts.ensure( size_t( nverts ) * 24 + 8 ) );
  SG_CUDA( ctx, m->samples.ensure( size_t( nsamples ) * 24 + 8 ) );
  SG_CUDA( ctx, m->hull.ensure( size_t( nhull ) * 24 + 8 ) );
  SG_CUDA( ctx, m->sdf.ensure( ncell * 8 + 8 ) );
  if( nverts ) { SG_CUDA( ctx, cudaMemcpyAsync( m->verts.ptr, verts, size_t( nverts ) * 24, cudaMemcpyHostToDevice, ctx->stream ) ); }
  if( nsamples ) { SG_CUDA( ctx, cudaMemcpyAsync( m->samples.ptr, samples, size_t( nsamples ) * 24, cudaMemcpyHostToDevice, ctx->stream ) ); }
  if( nhull ) { SG_CUDA( ctx, cudaMemcpyAsync( m->hull.ptr, hull, size_t( nhull ) * 24, cudaMemcpyHostToDevice, ctx->stream ) ); }
  SG_CUDA( ctx, cudaMemcpy2DAsync( m->sdf.ptr, size_t( pitch ) * 8, sdf, size_t( dims[0] ) * 8, size_t( dims[0] ) * 8, size_t( dims[1] ) * dims[2], cudaMemcpyHostToDevice, ctx->stream ) );
  m->dev.verts = m->verts.as<double>(); m->dev.nverts = nverts;
  m->dev.samples = m->samples.as<double>(); m->dev.nsamples = nsamples;
  m->dev.hull = m->hull.as<double>(); m->dev.nhull = nhull;
  m->dev.sdf = m->sdf.as<double>();
  for( int k = 0; k < 3; ++k )
  {
    m->dev.delta[k] = cell_delta[k]; m->dev.origin[k] = origin[k]; m->dev.dims[k] = dims[k];
    // RigidBodyTriangleMesh.cpp:102: grid_end = origin + (dims - 1) * delta  (host FP64; product then sum, no contraction)
    volatile double prod = double( dims[k] - 1u ) * cell_delta[k];
    m->dev.grid_end[k] = origin[k] + prod;
  }
  m->dev.pitch = pitch;
  m->dev.has_tmap = 0u;
  memset( &m->dev.tmap, 0, sizeof( m->dev.tmap ) );
  {
    // cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
    typedef CUresult ( *EncodeFn )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // SG_RB3D_NO_TMA=1 keeps every lookup on the direct path (A/B timing only; results are identical either way)
    const char* no_tma = getenv( "SG_RB3D_NO_TMA" );
    if( ( no_tma == nullptr || no_tma[0] == '0' ) && cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres ) == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr )
    {
      const cuuint64_t gdim[3] = { pitch, dims[1], dims[2] };
      const cuuint64_t gstride[2] = { cuuint64_t( pitch ) * 8, cuuint64_t( pitch ) * dims[1] * 8 };
      const cuuint32_t box[3] = { SDF_BX, SDF_BY, SDF_BZ };
      const cuuint32_t estr[3] = { 1, 1, 1 };
      const CUresult r = reinterpret_cast<EncodeFn>( fn )( &m->dev.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, m->sdf.ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
      m->dev.has_tmap = ( r == CUDA_SUCCESS ) ? 1u : 0u;
    }
    cudaGetLastError();
  }
  if( d->mesh_stats.ptr == nullptr )
  {
    SG_CUDA( ctx, d->mesh_stats.ensure( 16 ) );
    SG_CUDA( ctx, cudaMemsetAsync( d->mesh_stats.ptr, 0, 16, ctx->stream ) );
  }
  std::vector<MeshDev> all;
  for( MeshHost* mh : d->meshes ) { all.push_back( mh->dev ); }
  SG_CUDA( ctx, d->d_meshes.ensure( all.size() * sizeof( MeshDev ) ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->d_meshes.ptr, all.data(), all.size() * sizeof( MeshDev ), cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  if( mesh_index != nullptr ) { *mesh_index = uint32_t( d->meshes.size() - 1 ); }
  return SG_OK;
}

int sg_rb3d_set_bodies( sg_ctx* ctx, uint32_t n, const uint32_t* geo_of_body, const uint8_t* fixed, const double* m, const double* I0 )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > 0 && ( geo_of_body == nullptr || fixed == nullptr || m == nullptr || I0 == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_bodies: null array" ); }
  if( n >= 0x80000000u ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_bodies: at most 2^31 - 1 bodies" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  Rb3dData* d = rb3d_data( ctx );
  d->n = n;
  d->slab.on = false; // a plain body table: back to the single-GPU calls (sg_rb3d_slab_init re-enters slab mode)
  d->have_result = false;
  d->flow_resident = false;
  d->q1_valid = false;
  const int rc = rb3d_expand_bodies( ctx, d, n, geo_of_body, fixed );
  if( rc != SG_OK ) { d->n = 0; return rc; }
  d->h_geo_of_body.assign( geo_of_body, geo_of_body + n );
  d->h_fixed.assign( fixed, fixed + n );
  if( n == 0 ) { return SG_OK; }
  SG_CUDA( ctx, d->mass.ensure( size_t( n ) * 8 ) );
  SG_CUDA( ctx, d->I0.ensure( size_t( n ) * 24 ) );
  SG_CUDA( ctx, d->q0.ensure( size_t( n ) * 96 ) ); SG_CUDA( ctx, d->q1.ensure( size_t( n ) * 96 ) );
  SG_CUDA( ctx, d->v0.ensure( size_t( n ) * 48 ) ); SG_CUDA( ctx, d->v1.ensure( size_t( n ) * 48 ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->mass.ptr, m, size_t( n ) * 8, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->I0.ptr, I0, size_t( n ) * 24, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_rb3d_set_gravity( sg_ctx* ctx, const double* g )
{
  if( ctx == nullptr || g == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  d->g[0] = g[0]; d->g[1] = g[1]; d->g[2] = g[2];
  return SG_OK;
}

int sg_rb3d_set_planes( sg_ctx* ctx, uint32_t n, const double* x, const double* nrm )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > SG_MAX_PLANES ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_planes: at most %d planes", SG_MAX_PLANES ); }
  Rb3dData* d = rb3d_data( ctx );
  d->planes.n = n;
  for( uint32_t p = 0; p < n; ++p )
  {
    // StaticPlane::StaticPlane: m_n( n.normalized() ), 3-term squared norm (a0*a0 + a1*a1) + a2*a2
    double v[3] = { nrm[3 * p], nrm[3 * p + 1], nrm[3 * p + 2] };
    volatile double xx = v[0] * v[0]; volatile double yy = v[1] * v[1]; volatile double zz = v[2] * v[2];
    volatile double s01 = xx + yy;
    const double z = s01 + zz;
    if( z > 0.0 ) { const double s = sqrt( z ); v[0] = v[0] / s; v[1] = v[1] / s; v[2] = v[2] / s; }
    for( int k = 0; k < 3; ++k ) { d->planes.x[p][k] = x[3 * p + k]; d->planes.nrm[p][k] = v[k]; }
  }
  return SG_OK;
}

int sg_rb3d_set_cylinders( sg_ctx* ctx, uint32_t n, const double* x, const double* axis, const double* r )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > SG_MAX_CYLINDERS ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_cylinders: at most %d cylinders", SG_MAX_CYLINDERS ); }
  if( n > 0 && ( x == nullptr || axis == nullptr || r == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_cylinders: null array" ); }
  Rb3dData* d = rb3d_data( ctx );
  d->planes.ncyl = n;
  for( uint32_t c = 0; c < n; ++c )
  {
    // StaticCylinder::StaticCylinder: m_n( axis.normalized() ), 3-term squared norm (a0*a0 + a1*a1) + a2*a2
    double v[3] = { axis[3 * c], axis[3 * c + 1], axis[3 * c + 2] };
    volatile double xx = v[0] * v[0]; volatile double yy = v[1] * v[1]; volatile double zz = v[2] * v[2];
    volatile double s01 = xx + yy;
    const double z = s01 + zz;
    if( z > 0.0 ) { const double s = sqrt( z ); v[0] = v[0] / s; v[1] = v[1] / s; v[2] = v[2] / s; }
    for( int k = 0; k < 3; ++k ) { d->planes.cx[c][k] = x[3 * c + k]; d->planes.cax[c][k] = v[k]; }
    d->planes.cr[c] = r[c];
  }
  return SG_OK;
}

int sg_rb3d_set_portals( sg_ctx* ctx, uint32_t n, const double* plane_a_x, const double* plane_a_n, const double* plane_b_x, const double* plane_b_n, const int32_t* multiplier )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n > SG_MAX_PORTALS ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_portals: at most %d portals", SG_MAX_PORTALS ); }
  if( n > 0 && ( plane_a_x == nullptr || plane_a_n == nullptr || plane_b_x == nullptr || plane_b_n == nullptr || multiplier == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_portals: null array" ); }
  Rb3dData* d = rb3d_data( ctx );
  if( n == 0 && d->px == nullptr ) { return SG_OK; }
  if( d->px == nullptr ) { d->px = new Rb3dPortalData; }
  SgPortals3D ps;
  memset( &ps, 0, sizeof( ps ) );
  ps.n = n;
  for( uint32_t p = 0; p < n; ++p )
  {
    SgPortal3D& pt = ps.p[p];
    for( int k = 0; k < 3; ++k ) { pt.ax[k] = plane_a_x[3 * p + k]; pt.bx[k] = plane_b_x[3 * p + k]; pt.mult[k] = multiplier[3 * p + k]; }
    // StaticPlane::StaticPlane + t0() / t1() (rigidbody3d/StaticGeometry/StaticPlane.cpp:10-15,58-66)
    if( !sg_portal3_plane_frame( plane_a_n + 3 * p, pt.an, pt.at0, pt.at1 ) || !sg_portal3_plane_frame( plane_b_n + 3 * p, pt.bn, pt.bt0, pt.bt1 ) )
    {
      return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_set_portals: portal %u has a plane normal opposite to the y axis (Eigen's FromTwoVectors takes its SVD branch there; not reproduced)", p );
    }
  }
  d->px->portals = ps;
  d->have_result = false;
  return SG_OK;
}

int sg_rb3d_enforce_portals( sg_ctx* ctx, double* q )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  if( d->px == nullptr || d->px->portals.n == 0u || d->n == 0 ) { return SG_OK; }
  if( q == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_enforce_portals: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  d->flow_resident = false; // q1 serves as the staging copy
  d->q1_valid = false;
  const size_t bytes = size_t( d->n ) * 24; // only the centres of mass move
  SG_CUDA( ctx, cudaMemcpyAsync( d->q1.ptr, q, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_LAUNCH( ctx, "r3p_enforce", double( d->n ) * 48.0, k_r3p_enforce<<<sg_div_up( d->n, 256 ), 256, 0, ctx->stream>>>( d->px->portals, d->n, d->q1.as<double>() ) );
  SG_CUDA( ctx, cudaMemcpyAsync( q, d->q1.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

int sg_rb3d_teleported( sg_ctx* ctx, sg_teleported* out )
{
  if( ctx == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  memset( out, 0, sizeof( *out ) );
  Rb3dData* d = rb3d_data( ctx );
  if( !d->have_result || d->px == nullptr || !d->px->result ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_teleported: the last active set was not computed with portals" ); }
  Rb3dPortalData* x = d->px;
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  auto al = []( size_t b ) { return ( b + 63 ) & ~size_t( 63 ); };
  const size_t nb = x->n_boxes, nt = x->n_tel;
  size_t bytes = 64;
  const size_t o_bb = bytes; bytes += al( nb * 4 );
  const size_t o_bp = bytes; bytes += al( nb * 4 );
  const size_t o_p0 = bytes; bytes += al( nt * 4 );
  const size_t o_p1 = bytes; bytes += al( nt * 4 );
  const size_t o_x0 = bytes; bytes += al( nt * 24 );
  const size_t o_x1 = bytes; bytes += al( nt * 24 );
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
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_x0, x->x0t.ptr, nt * 24, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h + o_x1, x->x1t.ptr, nt * 24, cudaMemcpyDeviceToHost, ctx->stream ) );
  }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  out->n_boxes = nb; out->n_regular = x->n_reg; out->n_teleported = nt;
  out->box_body = reinterpret_cast<const uint32_t*>( h + o_bb ); out->box_portal = reinterpret_cast<const uint32_t*>( h + o_bp );
  out->portal0 = reinterpret_cast<const uint32_t*>( h + o_p0 ); out->portal1 = reinterpret_cast<const uint32_t*>( h + o_p1 );
  out->x0 = reinterpret_cast<const double*>( h + o_x0 ); out->x1 = reinterpret_cast<const double*>( h + o_x1 );
  return SG_OK;
}

int sg_rb3d_update_m_and_minv( sg_ctx* ctx, const double* q, double* m_blocks, double* minv_blocks )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  if( d->n == 0 ) { return SG_OK; }
  if( m_blocks == nullptr || minv_blocks == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_update_m_and_minv: null output" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( q != nullptr )
  {
    // q1 serves as the staging copy: whatever a flow left there is replaced
    d->flow_resident = false;
    d->q1_valid = false;
    SG_CUDA( ctx, cudaMemcpyAsync( d->q1.ptr, q, size_t( d->n ) * 96, cudaMemcpyHostToDevice, ctx->stream ) );
  }
  else if( !d->q1_valid ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_update_m_and_minv: q == NULL without a preceding sg_rb3d_flow / sg_rb3d_step on this context" ); }
  const size_t bytes = size_t( d->n ) * 72;
  SG_CUDA( ctx, d->minertia.ensure( 2 * bytes ) );
  double* blocks = d->minertia.as<double>();
  SG_LAUNCH( ctx, "rb3d_update_minertia", double( d->n ) * ( 96.0 + 144.0 ), k_rb3d_update_minertia<<<sg_div_up( d->n, 128 ), 128, 0, ctx->stream>>>( d->n, d->q1.as<double>(), d->I0.as<double>(), blocks,
             blocks + 9 * size_t( d->n ) ) );
  SG_CUDA( ctx, cudaMemcpyAsync( m_blocks, blocks, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( minv_blocks, blocks + 9 * size_t( d->n ), bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  return SG_OK;
}

int sg_rb3d_flow( sg_ctx* ctx, int map_kind, const double* q0, const double* v0, double dt, double* q1, double* v1 )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( ( map_kind & ~SG_MAP_M_UPDATED ) != SG_MAP_SPLIT_HAM && ( map_kind & ~SG_MAP_M_UPDATED ) != SG_MAP_DMV && ( map_kind & ~SG_MAP_M_UPDATED ) != SG_MAP_EXPONENTIAL_EULER ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_flow: map kind %d is not a rigidbody3d map", map_kind ); }
  Rb3dData* d = rb3d_data( ctx );
  if( d->n == 0 ) { return SG_OK; }
  if( q0 == nullptr || v0 == nullptr || q1 == nullptr || v1 == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_flow: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->q0.ptr, q0, size_t( d->n ) * 96, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->v0.ptr, v0, size_t( d->n ) * 48, cudaMemcpyHostToDevice, ctx->stream ) );
  const int rc = rb3d_flow_device( ctx, d, map_kind, dt );
  if( rc != SG_OK ) { return rc; }
  SG_CUDA( ctx, cudaMemcpyAsync( q1, d->q1.ptr, size_t( d->n ) * 96, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( v1, d->v1.ptr, size_t( d->n ) * 48, cudaMemcpyDeviceToHost, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  sg_prof_collect( ctx );
  d->flow_resident = true;
  return SG_OK;
}

int sg_rb3d_active_set( sg_ctx* ctx, const double* q0, const double* q1, uint32_t out_flags, sg_contacts* out )
{
  if( ctx == nullptr || out == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( ( out_flags & SG_IN_RESIDENT ) != 0u )
  {
    // (q0, q1) are the input and output of the last sg_rb3d_flow on this context: still on the device, nothing is uploaded
    if( !d->flow_resident ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_active_set: SG_IN_RESIDENT without a preceding sg_rb3d_flow on this context" ); }
  }
  else if( d->n > 0 )
  {
    if( q0 == nullptr || q1 == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_active_set: null vector" ); }
    SG_CUDA( ctx, cudaMemcpyAsync( d->q0.ptr, q0, size_t( d->n ) * 96, cudaMemcpyHostToDevice, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( d->q1.ptr, q1, size_t( d->n ) * 96, cudaMemcpyHostToDevice, ctx->stream ) );
    d->flow_resident = false;
    d->q1_valid = false;
  }
  const int rc = rb3d_active_set_device( ctx, d, ( out_flags & SG_OUT_CANDIDATES ) != 0u );
  if( rc != SG_OK ) { d->have_result = false; } // a failed call leaves nothing to fetch (partial or stale lists)
  if( rc != SG_OK ) { return rc; }
  return rb3d_copy_out( ctx, d, out_flags, out );
}

int sg_rb3d_upload( sg_ctx* ctx, const double* q, const double* v )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  d->flow_resident = false;
  if( d->n == 0 ) { return SG_OK; }
  if( q == nullptr || v == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_upload: null vector" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( d->slab.on )
  {
    // the caller's vectors hold the owned bodies only; on the device the [3n | 9n] / [3n | 3n] blocks are spaced by the slot count
    const size_t no = d->slab.n_owned, ns = d->n;
    if( no > 0 )
    {
      SG_CUDA( ctx, cudaMemcpyAsync( d->q0.ptr, q, no * 24, cudaMemcpyHostToDevice, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( d->q0.as<double>() + 3 * ns, q + 3 * no, no * 72, cudaMemcpyHostToDevice, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( d->v0.ptr, v, no * 24, cudaMemcpyHostToDevice, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( d->v0.as<double>() + 3 * ns, v + 3 * no, no * 24, cudaMemcpyHostToDevice, ctx->stream ) );
    }
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    return SG_OK;
  }
  SG_CUDA( ctx, cudaMemcpyAsync( d->q0.ptr, q, size_t( d->n ) * 96, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaMemcpyAsync( d->v0.ptr, v, size_t( d->n ) * 48, cudaMemcpyHostToDevice, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_rb3d_step( sg_ctx* ctx, int map_kind, double dt, sg_contacts* out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( ( map_kind & ~SG_MAP_M_UPDATED ) != SG_MAP_SPLIT_HAM && ( map_kind & ~SG_MAP_M_UPDATED ) != SG_MAP_DMV && ( map_kind & ~SG_MAP_M_UPDATED ) != SG_MAP_EXPONENTIAL_EULER ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_step: map kind %d is not a rigidbody3d map", map_kind ); }
  Rb3dData* d = rb3d_data( ctx );
  if( d->slab.on ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_step: context is in slab mode, use the sg_rb3d_slab_* calls" ); }
  d->flow_resident = false; // q1 is about to be overwritten by the resident step
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  int rc = rb3d_flow_device( ctx, d, map_kind, dt );
  if( rc != SG_OK ) { return rc; }
  rc = rb3d_active_set_device( ctx, d, true );
  if( rc != SG_OK ) { d->have_result = false; } // a failed call leaves nothing to fetch (partial or stale lists)
  if( rc != SG_OK ) { return rc; }
  if( out != nullptr )
  {
    memset( out, 0, sizeof( *out ) );
    out->dim = 3;
    out->n_candidates = d->n_cand;
    out->n_body_body = d->n_bb;
    out->n_plane = d->n_static;
    out->n_active = d->n_bb + d->n_static;
  }
  return SG_OK;
}

int sg_rb3d_mesh_stats( sg_ctx* ctx, uint64_t* staged_sweeps, uint64_t* direct_sweeps )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  unsigned long long h[2] = { 0ull, 0ull };
  if( d->mesh_stats.ptr != nullptr )
  {
    SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
    SG_CUDA( ctx, cudaMemcpyAsync( h, d->mesh_stats.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  }
  if( staged_sweeps != nullptr ) { *staged_sweeps = h[0]; }
  if( direct_sweeps != nullptr ) { *direct_sweeps = h[1]; }
  return SG_OK;
}

int sg_rb3d_fetch( sg_ctx* ctx, uint32_t out_flags, double* q1, double* v1, sg_contacts* out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  if( !d->have_result ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_fetch: no step has been run" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( d->slab.on )
  {
    const size_t no = d->slab.n_owned, ns = d->n;
    if( q1 != nullptr && no > 0 )
    {
      SG_CUDA( ctx, cudaMemcpyAsync( q1, d->q1.ptr, no * 24, cudaMemcpyDeviceToHost, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( q1 + 3 * no, d->q1.as<double>() + 3 * ns, no * 72, cudaMemcpyDeviceToHost, ctx->stream ) );
    }
    if( v1 != nullptr && no > 0 )
    {
      SG_CUDA( ctx, cudaMemcpyAsync( v1, d->v1.ptr, no * 24, cudaMemcpyDeviceToHost, ctx->stream ) );
      SG_CUDA( ctx, cudaMemcpyAsync( v1 + 3 * no, d->v1.as<double>() + 3 * ns, no * 24, cudaMemcpyDeviceToHost, ctx->stream ) );
    }
  }
  else
  {
    if( q1 != nullptr && d->n > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( q1, d->q1.ptr, size_t( d->n ) * 96, cudaMemcpyDeviceToHost, ctx->stream ) ); }
    if( v1 != nullptr && d->n > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( v1, d->v1.ptr, size_t( d->n ) * 48, cudaMemcpyDeviceToHost, ctx->stream ) ); }
  }
  if( out != nullptr ) { return rb3d_copy_out( ctx, d, out_flags, out ); }
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}


// ---- slab mode (multi-GPU), all-sphere scenes: the calls of the ball2d slab mode for rigidbody3d --------------------------------
// One spatial slab of a larger scene per context: n_owned spheres in ascending global index in slots [0, n_owned), then ghost_cap halo
// slots for the lower neighbour and ghost_cap for the higher one.  q, v of sg_rb3d_upload / sg_rb3d_fetch address the owned bodies
// ( [3 n_owned | 9 n_owned], [3 n_owned | 3 n_owned] ).  Peer-memory exchange only (sg_slab.cuh); DESIGN.md section 5.
int sg_rb3d_slab_init( sg_ctx* ctx, uint32_t n_owned, uint32_t ghost_cap, const uint32_t* geo_of_body, const double* m, const double* I0, const uint32_t* gid_owned, const double* x_limits )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  if( n_owned > 0 && ( geo_of_body == nullptr || m == nullptr || I0 == nullptr || gid_owned == nullptr ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_init: null array" ); }
  if( uint64_t( n_owned ) + 2ull * ghost_cap >= 0x40000000ull ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_init: slab too large" ); }
  Rb3dData* d = rb3d_data( ctx );
  if( d->px != nullptr && d->px->portals.n > 0u ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_slab_init: portals are not supported in slab mode" ); }
  if( d->geo_type.empty() ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_init: call sg_rb3d_set_geometry first" ); }
  for( uint32_t k = 0; k < n_owned; ++k )
  {
    if( gid_owned[k] >= 0x40000000u || ( k > 0 && gid_owned[k] <= gid_owned[k - 1] ) ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_init: global indices must be strictly ascending and below 2^30 (entry %u)", k ); }
  }
  const size_t slots = size_t( n_owned ) + 2 * size_t( ghost_cap );
  // every slot gets a body: the ghosts are spheres of the first geometry until a halo record overwrites their radius
  uint32_t sphere_geo = 0u;
  while( sphere_geo < d->geo_type.size() && d->geo_type[sphere_geo] != SG_GEO_SPHERE ) { ++sphere_geo; }
  if( sphere_geo == d->geo_type.size() ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_slab_init: slab mode is for all-sphere scenes" ); }
  std::vector<uint32_t> geo( slots, sphere_geo );
  std::vector<uint8_t> fixed( slots, 0 );
  std::vector<double> mm( slots, 1.0 ), ii( 3 * slots, 1.0 );
  for( uint32_t k = 0; k < n_owned; ++k ) { geo[k] = geo_of_body[k]; mm[k] = m[k]; ii[3 * size_t( k )] = I0[3 * size_t( k )]; ii[3 * size_t( k ) + 1] = I0[3 * size_t( k ) + 1]; ii[3 * size_t( k ) + 2] = I0[3 * size_t( k ) + 2]; }
  d->slab.on = false;
  int rc = sg_rb3d_set_bodies( ctx, uint32_t( slots ), geo.data(), fixed.data(), mm.data(), ii.data() );
  if( rc != SG_OK ) { return rc; }
  if( !d->all_spheres ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_slab_init: slab mode is for all-sphere scenes" ); }
  SlabComm& c = d->slab;
  c.on = true; c.n_owned = n_owned; c.ghost_cap = ghost_cap; c.rec_bytes = sizeof( SphereGhostRec ); c.scan_done = false;
  c.xlim[0] = ( x_limits != nullptr ) ? x_limits[0] : -1.0e308; c.xlim[1] = ( x_limits != nullptr ) ? x_limits[1] : 1.0e308;
  SG_CUDA( ctx, c.gid.ensure( slots * 4 + 4 ) );
  SG_CUDA( ctx, c.interval_enc.ensure( 16 ) );
  SG_CUDA( ctx, c.ghost_counts.ensure( 16 ) );
  SG_CUDA( ctx, cudaMemsetAsync( c.gid.ptr, 0, slots * 4, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( c.ghost_counts.ptr, 0, 16, ctx->stream ) );
  if( n_owned > 0 ) { SG_CUDA( ctx, cudaMemcpyAsync( c.gid.ptr, gid_owned, size_t( n_owned ) * 4, cudaMemcpyHostToDevice, ctx->stream ) ); }
  k_slab_begin<<<1, 1, 0, ctx->stream>>>( c.interval_enc.as<long long>(), c.ghost_counts.as<uint32_t>() );
  // the orientation slots of the ghosts are never read; zero the state once so that nothing uninitialised travels
  SG_CUDA( ctx, cudaMemsetAsync( d->q0.ptr, 0, slots * 96, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->q1.ptr, 0, slots * 96, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->v0.ptr, 0, slots * 48, ctx->stream ) );
  SG_CUDA( ctx, cudaMemsetAsync( d->v1.ptr, 0, slots * 48, ctx->stream ) );
  SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
  return SG_OK;
}

int sg_rb3d_slab_flow( sg_ctx* ctx, int map_kind, double dt )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  if( !d->slab.on ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_flow: call sg_rb3d_slab_init first" ); }
  if( d->slab.mailbox.ptr == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_flow: create the mailbox first (sg_rb3d_slab_mailbox)" ); }
  const int k = map_kind & ~SG_MAP_M_UPDATED;
  if( k != SG_MAP_SPLIT_HAM && k != SG_MAP_DMV && k != SG_MAP_EXPONENTIAL_EULER ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_flow: map kind %d is not a rigidbody3d map", map_kind ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SlabComm& c = d->slab;
  const int rc = rb3d_flow_device( ctx, d, map_kind, dt );
  if( rc != SG_OK ) { return rc; }
  const SlabCand sc = c.cand();
  if( sc.band != nullptr ) { SG_CUDA( ctx, cudaMemsetAsync( sc.count, 0, 8, ctx->stream ) ); }
  SG_LAUNCH( ctx, "slab_scan", double( c.n_owned ) * 16.0, k_rb3d_slab_scan<<<sg_div_up( c.n_owned > 0 ? c.n_owned : 1, 256 ), 256, 0, ctx->stream>>>( c.n_owned, d->q1.as<double>(), d->radius.as<double>(),
             c.interval_enc.as<long long>(), c.xlim[0], c.xlim[1], c.ghost_counts.as<uint32_t>(), sc ) );
  ++c.step;
  SG_LAUNCH( ctx, "slab_interval", 16.0, k_slab_post_interval<<<1, 1, 0, ctx->stream>>>( c.interval_enc.as<long long>(), c.ghost_counts.as<uint32_t>(), nullptr, static_cast<SlabMailboxHdr*>( c.peer_mb[0] ),
             static_cast<SlabMailboxHdr*>( c.peer_mb[1] ), c.step ) );
  c.scan_done = true;
  return SG_OK;
}

int sg_rb3d_slab_mailbox( sg_ctx* ctx, void** mailbox_dev, void* ipc_handle_64 )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  if( !d->slab.on ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_mailbox: call sg_rb3d_slab_init first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  return slab_comm_mailbox( ctx, d->slab, mailbox_dev, ipc_handle_64 );
}

int sg_rb3d_slab_connect( sg_ctx* ctx, int side, const void* ipc_handle_64, void* same_process_mailbox, int peer_device )
{
  if( ctx == nullptr || ( side != 0 && side != 1 ) || ( ipc_handle_64 == nullptr ) == ( same_process_mailbox == nullptr ) ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  if( !d->slab.on || d->slab.mailbox.ptr == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_connect: create this rank's mailbox first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  return slab_comm_connect( ctx, d->slab, side, ipc_handle_64, same_process_mailbox, peer_device );
}

int sg_rb3d_slab_disconnect( sg_ctx* ctx )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  return slab_comm_disconnect( ctx, d->slab );
}

// phase 1: wait for the neighbours' intervals, pack the spheres that reach them straight into their mailboxes; phase 2: wait for the
// neighbours' halos and move them into the ghost slots; phase 0: both
int sg_rb3d_slab_exchange( sg_ctx* ctx, int phase )
{
  if( ctx == nullptr || phase < 0 || phase > 2 ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  SlabComm& c = d->slab;
  if( !c.on || c.mailbox.ptr == nullptr ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_exchange: no mailbox" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  SlabMailboxHdr* mine = c.mailbox.as<SlabMailboxHdr>();
  const bool any = c.peer_mb[0] != nullptr || c.peer_mb[1] != nullptr;
  if( ( phase == 0 || phase == 1 ) && any )
  {
    if( !c.scan_done ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_exchange: call sg_rb3d_slab_flow first" ); }
    Pack2Args<SphereGhostRec> pa;
    for( int side = 0; side < 2; ++side )
    {
      pa.on[side] = c.peer_mb[side] != nullptr;
      pa.iv[side] = &mine->iv[side][0]; pa.wait[side] = &mine->iv_flag[side]; pa.out[side] = nullptr; pa.post[side] = nullptr;
      if( !pa.on[side] ) { continue; }
      SlabMailboxHdr* peer = static_cast<SlabMailboxHdr*>( c.peer_mb[side] );
      pa.out[side] = slab_mailbox_halo<SphereGhostRec>( peer, 1 - side, c.ghost_cap );
      pa.post[side] = &peer->halo_flag[1 - side];
    }
    pa.err = &mine->err; pa.step = c.step;
    Sphere3DSlabTraits::Src src;
    src.q0 = d->q0.as<double>(); src.q1 = d->q1.as<double>(); src.r = d->radius.as<double>(); src.gid = c.gid.as<uint32_t>();
    SG_LAUNCH( ctx, "slab_pack", double( c.ghost_cap ) * 2.0 * 64.0, k_slab_pack2<Sphere3DSlabTraits><<<unsigned( ctx->num_sms ), 256, 0, ctx->stream>>>( 0u, c.n_owned, src, c.ghost_cap, c.cand(), c.cand_state.as<SlabCandState>(),
               d->bp.params.as<GridParams>(), pa ) );
  }
  if( ( phase == 0 || phase == 2 ) && any )
  {
    SphereUnpackArgs ua;
    for( int side = 0; side < 2; ++side )
    {
      ua.on[side] = c.peer_mb[side] != nullptr;
      ua.in[side] = slab_mailbox_halo<SphereGhostRec>( mine, side, c.ghost_cap );
      ua.wait[side] = &mine->halo_flag[side];
    }
    ua.err = &mine->err; ua.step = c.step;
    const unsigned bps = sg_div_up( c.ghost_cap > 0 ? c.ghost_cap : 1, 256 );
    SG_LAUNCH( ctx, "slab_unpack", double( c.ghost_cap ) * 4.0, k_rb3d_slab_unpack<<<2 * bps, 256, 0, ctx->stream>>>( c.ghost_cap, bps, ua, c.n_owned, d->q0.as<double>(), d->q1.as<double>(), d->radius.as<double>(),
               c.gid.as<uint32_t>(), c.ghost_counts.as<uint32_t>() ) );
  }
  return SG_OK;
}

int sg_rb3d_slab_detect( sg_ctx* ctx, sg_contacts* out, uint32_t* ghosts_out )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  SlabComm& c = d->slab;
  if( !c.on ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_slab_detect: call sg_rb3d_slab_init first" ); }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  if( d->bp.params.ptr == nullptr ) { SG_CUDA( ctx, d->bp.params.ensure( sizeof( GridParams ) ) ); SG_CUDA( ctx, cudaMemsetAsync( d->bp.params.ptr, 0, sizeof( GridParams ), ctx->stream ) ); }
  const int rc = rb3d_active_set_device( ctx, d, true );
  if( rc != SG_OK ) { d->have_result = false; } // a failed call leaves nothing to fetch (partial or stale lists)
  if( rc != SG_OK ) { return rc; }
  c.scan_done = false;
  uint32_t hg[4];
  SG_CUDA( ctx, cudaMemcpy( hg, c.ghost_counts.ptr, 16, cudaMemcpyDeviceToHost ) );
  if( hg[2] != 0u ) { return sg_fail( ctx, SG_ERR_INVALID, "slab halo exceeds the reserved ghost capacity %u: results are incomplete", c.ghost_cap ); }
  if( hg[3] != 0u ) { return sg_fail( ctx, SG_ERR_REBALANCE, "a body of this slab left [%g, %g]: it may reach a non-neighbouring slab, re-partition the scene", c.xlim[0], c.xlim[1] ); }
  if( c.mailbox.ptr != nullptr )
  {
    uint32_t err = 0u;
    SG_CUDA( ctx, cudaMemcpy( &err, &c.mailbox.as<SlabMailboxHdr>()->err, 4, cudaMemcpyDeviceToHost ) );
    if( err != 0u ) { return sg_fail( ctx, SG_ERR_INTERNAL, "slab exchange: a neighbour did not post its interval or halo within 10 s" ); }
  }
  if( ghosts_out != nullptr ) { ghosts_out[0] = hg[0]; ghosts_out[1] = hg[1]; }
  if( out != nullptr )
  {
    memset( out, 0, sizeof( *out ) );
    out->dim = 3;
    out->n_candidates = d->n_cand;
    out->n_body_body = d->n_bb;
    out->n_plane = d->n_static;
    out->n_active = d->n_bb + d->n_static;
  }
  return SG_OK;
}

// A triangle mesh's own record for the state snapshot (RigidBodyTriangleMesh::serialize, rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp:215-232; type byte
// included).  It holds what this library has no use for (file name, faces, volume, moments) next to what sg_rb3d_add_mesh was given; it is checked to be a
// well-formed record whose arrays have the sizes of mesh `mesh_index`, kept on the host, and written back verbatim.
int sg_rb3d_set_mesh_snapshot( sg_ctx* ctx, uint32_t mesh_index, const void* record, uint64_t bytes )
{
  if( ctx == nullptr ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  if( mesh_index >= d->meshes.size() ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_mesh_snapshot: mesh %u of %zu", mesh_index, d->meshes.size() ); }
  if( record == nullptr || bytes == 0 ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_mesh_snapshot: empty record" ); }
  sg_snapshot::Source in{ static_cast<const unsigned char*>( record ), bytes, 0, true };
  sg_snapshot::Rb3dState::Mesh mm;
  if( in.val<unsigned char>() != 3 || !sg_snapshot::take_mesh( in, mm ) || in.n != bytes ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_mesh_snapshot: not a RigidBodyTriangleMesh record of %llu bytes", ( unsigned long long )( bytes ) ); }
  const MeshDev& dev = d->meshes[mesh_index]->dev;
  if( mm.verts.size() / 3 != dev.nverts || mm.samples.size() / 3 != dev.nsamples || mm.hull.size() / 3 != dev.nhull || mm.dims[0] != dev.dims[0] || mm.dims[1] != dev.dims[1] || mm.dims[2] != dev.dims[2] )
  {
    return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_set_mesh_snapshot: the record's vertex, sample, hull or grid sizes are not those of mesh %u", mesh_index );
  }
  if( d->mesh_record.size() < d->meshes.size() ) { d->mesh_record.resize( d->meshes.size() ); }
  d->mesh_record[mesh_index].assign( static_cast<const unsigned char*>( record ), static_cast<const unsigned char*>( record ) + bytes );
  return SG_OK;
}

// ---- state I/O at the seam: RigidBody3DState's binary snapshot (rigidbody3d/RigidBody3DState.cpp:586-668), sg_rb3d_snapshot.h ----
int sg_rb3d_state_serialize( sg_ctx* ctx, int which, int m_updated, void* buf, uint64_t cap, uint64_t* bytes )
{
  if( ctx == nullptr || bytes == nullptr || ( which != 0 && which != 1 ) ) { return SG_ERR_INVALID; }
  Rb3dData* d = rb3d_data( ctx );
  if( d->slab.on ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_state_serialize: a slab holds part of a scene; serialise through the owner of the whole state" ); }
  if( which == 1 && !d->q1_valid && d->n > 0 ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_state_serialize: which = 1 without a preceding sg_rb3d_flow / sg_rb3d_step on this context" ); }
  for( size_t k = 0; k < d->geo_type.size(); ++k )
  {
    const uint32_t t = d->geo_type[k];
    if( t == SG_GEO_BOX || t == SG_GEO_SPHERE ) { continue; }
    // a triangle mesh's snapshot holds its whole input file (names, faces, volume ...: RigidBodyTriangleMesh.cpp:215-232): written back from the record the caller attached
    const bool have = t == SG_GEO_MESH && d->geo_mesh[k] < d->mesh_record.size() && !d->mesh_record[d->geo_mesh[k]].empty();
    if( !have ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_state_serialize: geometry %zu is a triangle mesh without its record (sg_rb3d_set_mesh_snapshot), or a geometry this path does not hold", k ); }
  }
  SG_CUDA( ctx, cudaSetDevice( ctx->device ) );
  const uint32_t n = d->n;
  sg_snapshot::Rb3dState s;
  s.n = n;
  s.q.resize( size_t( 12 ) * n ); s.v.resize( size_t( 6 ) * n ); s.m.resize( n ); s.I0.resize( size_t( 3 ) * n ); s.I.resize( size_t( 9 ) * n ); s.Iinv.resize( size_t( 9 ) * n );
  if( n > 0 )
  {
    const double* qd = ( which == 0 ) ? d->q0.as<double>() : d->q1.as<double>();
    const double* vd = ( which == 0 ) ? d->v0.as<double>() : d->v1.as<double>();
    const size_t bb = size_t( n ) * 72;
    SG_CUDA( ctx, d->minertia.ensure( 2 * bb ) );
    double* blocks = d->minertia.as<double>();
    SG_LAUNCH( ctx, "rb3d_update_minertia", double( n ) * ( 96.0 + 144.0 ), k_rb3d_update_minertia<<<sg_div_up( n, 128 ), 128, 0, ctx->stream>>>( n, qd, d->I0.as<double>(), blocks, blocks + 9 * size_t( n ) ) );
    SG_CUDA( ctx, cudaMemcpyAsync( s.q.data(), qd, size_t( n ) * 96, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( s.v.data(), vd, size_t( n ) * 48, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( s.m.data(), d->mass.ptr, size_t( n ) * 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( s.I0.data(), d->I0.ptr, size_t( n ) * 24, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( s.I.data(), blocks, bb, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaMemcpyAsync( s.Iinv.data(), blocks + 9 * size_t( n ), bb, cudaMemcpyDeviceToHost, ctx->stream ) );
    SG_CUDA( ctx, cudaStreamSynchronize( ctx->stream ) );
    sg_prof_collect( ctx );
    if( !m_updated )
    {
      // as RigidBody3DState's constructor stores the blocks: transposed (formWorldSpaceMassMatrix, RigidBody3DState.cpp:165-182)
      for( std::vector<double>* blk : { &s.I, &s.Iinv } )
      {
        for( uint32_t b = 0; b < n; ++b )
        {
          double* a = blk->data() + 9 * size_t( b );
          double t;
          t = a[1]; a[1] = a[3]; a[3] = t; t = a[2]; a[2] = a[6]; a[6] = t; t = a[5]; a[5] = a[7]; a[7] = t;
        }
      }
    }
  }
  s.fixed = d->h_fixed; s.geo_of_body = d->h_geo_of_body;
  s.geo_type.resize( d->geo_type.size() );
  s.geo_blob.assign( d->geo_type.size(), std::vector<unsigned char>() );
  for( size_t k = 0; k < d->geo_type.size(); ++k )
  {
    s.geo_type[k] = ( d->geo_type[k] == SG_GEO_BOX ) ? 0u : ( d->geo_type[k] == SG_GEO_SPHERE ) ? 1u : 3u;
    if( s.geo_type[k] == 3u ) { s.geo_blob[k] = d->mesh_record[d->geo_mesh[k]]; }
  }
  s.geo_r = d->geo_r; s.geo_half = d->geo_half;
  for( int k = 0; k < 3; ++k ) { s.g[k] = d->g[k]; }
  for( uint32_t p = 0; p < d->planes.n; ++p ) { for( int k = 0; k < 3; ++k ) { s.plane_x.push_back( d->planes.x[p][k] ); s.plane_n.push_back( d->planes.nrm[p][k] ); } }
  for( uint32_t c = 0; c < d->planes.ncyl; ++c ) { for( int k = 0; k < 3; ++k ) { s.cyl_x.push_back( d->planes.cx[c][k] ); s.cyl_axis.push_back( d->planes.cax[c][k] ); } s.cyl_r.push_back( d->planes.cr[c] ); }
  if( d->px != nullptr )
  {
    for( uint32_t p = 0; p < d->px->portals.n; ++p )
    {
      const SgPortal3D& pt = d->px->portals.p[p];
      for( int k = 0; k < 3; ++k ) { s.portal_ax.push_back( pt.ax[k] ); s.portal_an.push_back( pt.an[k] ); s.portal_bx.push_back( pt.bx[k] ); s.portal_bn.push_back( pt.bn[k] ); s.portal_mult.push_back( pt.mult[k] ); }
    }
  }
  sg_snapshot::Sink out{ static_cast<unsigned char*>( buf ), cap, 0 };
  if( !sg_snapshot::serialize( s, out ) ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_state_serialize: a geometry that is neither box, sphere nor a mesh with its record" ); }
  *bytes = out.n;
  if( buf != nullptr && out.n > cap ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_state_serialize: buffer of %llu bytes, %llu needed", ( unsigned long long )( cap ), ( unsigned long long )( out.n ) ); }
  return SG_OK;
}

// RigidBody3DState::deserialize (rigidbody3d/RigidBody3DState.cpp:650-668): configures the context from a snapshot and uploads ( q, v )
int sg_rb3d_state_deserialize( sg_ctx* ctx, const void* buf, uint64_t bytes )
{
  if( ctx == nullptr || buf == nullptr ) { return SG_ERR_INVALID; }
  sg_snapshot::Source in{ static_cast<const unsigned char*>( buf ), bytes, 0, true };
  sg_snapshot::Rb3dState s;
  const char* why = "";
  const int prc = sg_snapshot::parse( in, s, &why );
  if( prc == 1 ) { return sg_fail( ctx, SG_ERR_INVALID, "sg_rb3d_state_deserialize: %s", why ); }
  if( prc == 2 ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_state_deserialize: %s", why ); }
  if( s.plane_x.size() / 3 > SG_MAX_PLANES || s.cyl_r.size() > SG_MAX_CYLINDERS || s.portal_mult.size() / 3 > SG_MAX_PORTALS ) { return sg_fail( ctx, SG_ERR_UNSUPPORTED, "sg_rb3d_state_deserialize: more planes, cylinders or portals than this library holds" ); }
  const uint32_t ngeo = uint32_t( s.geo_type.size() );
  std::vector<uint32_t> type( ngeo ), mesh( ngeo, 0u );
  int rc = SG_OK;
  for( uint32_t k = 0; k < ngeo; ++k )
  {
    type[k] = ( s.geo_type[k] == 0u ) ? uint32_t( SG_GEO_BOX ) : ( s.geo_type[k] == 1u ) ? uint32_t( SG_GEO_SPHERE ) : uint32_t( SG_GEO_MESH );
    if( s.geo_type[k] != 3u ) { continue; }
    // RigidBodyTriangleMesh( std::istream& ) (RigidBodyTriangleMesh.cpp:105-129): the mesh as stored; its record is kept for the next snapshot
    const sg_snapshot::Rb3dState::Mesh& mm = s.mesh[k];
    rc = sg_rb3d_add_mesh( ctx, uint32_t( mm.verts.size() / 3 ), mm.verts.data(), uint32_t( mm.samples.size() / 3 ), mm.samples.data(), uint32_t( mm.hull.size() / 3 ), mm.hull.data(),
                           mm.delta, mm.dims, mm.origin, mm.sdf.data(), &mesh[k] );
    if( rc != SG_OK ) { return rc; }
    rc = sg_rb3d_set_mesh_snapshot( ctx, mesh[k], s.geo_blob[k].data(), s.geo_blob[k].size() );
    if( rc != SG_OK ) { return rc; }
  }
  rc = sg_rb3d_set_geometry( ctx, ngeo, type.data(), s.geo_r.data(), s.geo_half.data(), mesh.data() );
  if( rc != SG_OK ) { return rc; }
  rc = sg_rb3d_set_bodies( ctx, s.n, s.geo_of_body.data(), s.fixed.data(), s.m.data(), s.I0.data() );
  if( rc != SG_OK ) { return rc; }
  Rb3dData* d = rb3d_data( ctx );
  for( int k = 0; k < 3; ++k ) { d->g[k] = s.g[k]; }
  // StaticPlane( std::istream& ) / StaticCylinder( std::istream& ) read x and n back as stored: the normals are NOT normalised again
  d->planes.n = uint32_t( s.plane_x.size() / 3 );
  for( uint32_t p = 0; p < d->planes.n; ++p ) { for( int k = 0; k < 3; ++k ) { d->planes.x[p][k] = s.plane_x[3 * p + k]; d->planes.nrm[p][k] = s.plane_n[3 * p + k]; } }
  d->planes.ncyl = uint32_t( s.cyl_r.size() );
  for( uint32_t c = 0; c < d->planes.ncyl; ++c ) { for( int k = 0; k < 3; ++k ) { d->planes.cx[c][k] = s.cyl_x[3 * c + k]; d->planes.cax[c][k] = s.cyl_axis[3 * c + k]; } d->planes.cr[c] = s.cyl_r[c]; }
  const uint32_t npo = uint32_t( s.portal_mult.size() / 3 );
  if( npo > 0 || d->px != nullptr )
  {
    rc = sg_rb3d_set_portals( ctx, npo, s.portal_ax.data(), s.portal_an.data(), s.portal_bx.data(), s.portal_bn.data(), s.portal_mult.data() );
    if( rc != SG_OK ) { return rc; }
    // the tangents come from the stored normals as the reference computes them on demand; the normals themselves stay as stored
    for( uint32_t p = 0; p < npo; ++p ) { for( int k = 0; k < 3; ++k ) { d->px->portals.p[p].an[k] = s.portal_an[3 * p + k]; d->px->portals.p[p].bn[k] = s.portal_bn[3 * p + k]; } }
  }
  if( s.n == 0 ) { return SG_OK; }
  return sg_rb3d_upload( ctx, s.q.data(), s.v.data() );
}

}
