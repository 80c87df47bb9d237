/* pfft_oracle.c -- CPU restatement ("port") of the reference's algorithm.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load this.
 * It follows, level by level, what /root/reference computes for one packed 1-D transform:
 *   planner predicates  src/portfft/common/workitem.hpp:135-185, src/portfft/common/subgroup.hpp:226-253,
 *                       src/portfft/committed_descriptor_impl.hpp:210-313, src/portfft/utils.hpp:94-132
 *   WORKITEM            src/portfft/common/workitem.hpp:64-127,200-219
 *   SUBGROUP            src/portfft/common/subgroup.hpp:170-216,271-291
 *   WORKGROUP           src/portfft/common/workgroup.hpp:85-346
 *   GLOBAL              src/portfft/dispatcher/global_dispatcher.hpp:107-256,312-412, src/portfft/common/global.hpp
 * in the descriptor's Scalar precision (fp32 or fp64), batch-parallel with OpenMP.
 *
 * Pinning: tests/test_oracle.py checks it against the numpy oracle on the reference's test grid sizes and against
 * the reference's own wi_dft / planner predicates compiled from /root/reference (oracle/_ref, tests/test_ref_shim.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* factorize, workitem.hpp:135-144 */
long long ref_factorize(long long n) {
  long long res = 1;
  for (long long i = 2; i * i <= n; i++)
    if (n % i == 0) res = i;
  return res;
}

/* wi_temps, workitem.hpp:154-169 (MaxRecursionLevel = int_log2(56) - 1 = 4) */
static long long wi_temps_rec(long long n, int level) {
  long long f0 = ref_factorize(n), f1 = n / f0;
  if (f0 < 2 || f1 < 2) return n;
  long long a = 2, b = 2;
  if (level < 4) {
    a = wi_temps_rec(f0, level + 1);
    b = wi_temps_rec(f1, level + 1);
  }
  return (a > b ? a : b) + n;
}
long long ref_wi_temps(long long n) { return wi_temps_rec(n, 0); }

/* fits_in_wi, workitem.hpp:179-185 with PORTFFT_REGISTERS_PER_WI = 128 (CMakeLists.txt:53) */
int ref_fits_in_wi(long long n, int is_double) {
  return (n + ref_wi_temps(n)) * 2 * (is_double ? 8 : 4) <= 128 * 4;
}

/* factorize_sg, subgroup.hpp:226-238 */
long long ref_factorize_sg(long long n, int sg) {
  for (long long i = sg; i > 1; i--)
    if (n % i == 0) return i;
  return 1;
}

/* fits_in_sg, subgroup.hpp:248-253 */
int ref_fits_in_sg(long long n, int sg, int is_double) { return ref_fits_in_wi(n / ref_factorize_sg(n, sg), is_double); }

/* prepare_implementation, committed_descriptor_impl.hpp:210-313: returns level 0..3 (-1: unsupported) and, for
 * GLOBAL, the factor chain of factorize_input (utils.hpp:94-132) with WI / SG factors. */
int ref_select_level(long long n, int is_double, long long local_mem_bytes, long long* factors, int* nfac) {
  *nfac = 0;
  if (ref_fits_in_wi(n, is_double)) return 0;
  if (ref_fits_in_sg(n, 32, is_double)) return 1;
  long long fn = ref_factorize(n);
  if (fn == 1) return -1;
  long long fm = n / fn;
  long long wi_n = fn / ref_factorize_sg(fn, 32), wi_m = fm / ref_factorize_sg(fm, 32);
  /* num_scalars_in_local_mem, workgroup_dispatcher.hpp:364-380, PACKED */
  long long scalar = is_double ? 8 : 4;
  long long row_bytes = 2 * scalar * fm;
  long long blp = (row_bytes % 128 == 0) ? row_bytes / 128 : 1;
  long long scalars = 2 * n + (2 * n) / (32 * blp) + 2 * (fn + fm);
  if (ref_fits_in_wi(wi_n, is_double) && ref_fits_in_wi(wi_m, is_double) && scalars * scalar <= local_mem_bytes) return 2;
  long long done = 1;
  while (n / done != 1) {
    long long f = n / done;
    if (!(ref_fits_in_wi(f, is_double) || ref_fits_in_sg(f, 32, is_double))) {
      if (ref_factorize(f) == 1) return -1;
      do {
        f = ref_factorize(f);
        if (f == 1) return -1;
      } while (!(ref_fits_in_wi(f, is_double) || ref_fits_in_sg(f, 32, is_double)));
    }
    if (*nfac >= 64) return -1;
    factors[(*nfac)++] = f;
    done *= f;
  }
  return 3;
}

#define T float
#define SUF _f32
#include "pfft_oracle_impl.h"
#undef T
#undef SUF

#define T double
#define SUF _f64
#include "pfft_oracle_impl.h"
#undef T
#undef SUF
