// tests/ccd_harness.cpp -- TEST INFRASTRUCTURE.  Compiles the product's ball-ball CCD (scisim_b200/csrc/sg_ccd.h, the header the
// CUDA kernels include) as plain C++ so that the CPU suite can pin its division-free verdict against the reference's compiled
// CollisionDetectionUtilities.cpp and against the reference's expressions on a large sweep.  Built by tests/test_ccd_cpu.py with
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC   (no FMA contraction, like the library's -fmad=false).  Nothing here is shipped.
#include "../scisim_b200/csrc/sg_ccd.h"

#include <cstdint>

extern "C"
{

int ccdh_ball_ball( const double* q0a, const double* q1a, double ra, const double* q0b, const double* q1b, double rb )
{
  return sg_ccd_ball_ball( q0a[0], q0a[1], q1a[0], q1a[1], ra, q0b[0], q0b[1], q1b[0], q1b[1], rb ) ? 1 : 0;
}

int ccdh_roots( double c0, double c1, double c2 ) { return sg_ccd_roots( c0, c1, c2 ) ? 1 : 0; }
int ccdh_roots_verbatim( double c0, double c1, double c2 ) { return sg_ccd_roots_verbatim( c0, c1, c2 ) ? 1 : 0; }

// n triples (c0, c1, c2), c2 != 0: number of triples on which the two verdicts differ; *first = index of the first one
uint64_t ccdh_sweep( uint64_t n, const double* c, uint64_t* first, uint64_t* hits )
{
  uint64_t bad = 0, h = 0;
  for( uint64_t k = 0; k < n; ++k )
  {
    const double c0 = c[3 * k], c1 = c[3 * k + 1], c2 = c[3 * k + 2];
    if( c2 == 0.0 ) { continue; }
    const bool a = sg_ccd_roots( c0, c1, c2 ), b = sg_ccd_roots_verbatim( c0, c1, c2 );
    if( a != b ) { if( bad == 0 && first != nullptr ) { *first = k; } ++bad; }
    if( b ) { ++h; }
  }
  if( hits != nullptr ) { *hits = h; }
  return bad;
}

}
