/* Type-generic body of the C oracle; included twice by pfft_oracle.c with T / SUF defined.
 * Complex data is interleaved (re, im) arrays of T, as in the reference's private / local memory. */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* (cospi(-2k/n), sinpi(-2k/n)): src/portfft/common/twiddle_calc.hpp:37-41 (evaluated in double, rounded once).
 * Tables are built once per length and cached, like the reference's commit-time twiddle buffers
 * (committed_descriptor_impl.hpp:737-752) and its compile-time table (common/twiddle.hpp). */
#define MAX_TABLES 256
static struct {
  long long n;
  T* w;
} FN(tables)[MAX_TABLES];
static int FN(num_tables) = 0;

static const T* FN(get_table)(long long n) {
  for (int i = 0; i < FN(num_tables); i++)
    if (FN(tables)[i].n == n) return FN(tables)[i].w;
  const T* res = NULL;
#pragma omp critical(pfft_oracle_tables)
  {
    for (int i = 0; i < FN(num_tables); i++)
      if (FN(tables)[i].n == n) res = FN(tables)[i].w;
    if (!res && FN(num_tables) < MAX_TABLES) {
      T* w = (T*)malloc(sizeof(T) * 2 * (size_t)n);
      for (long long k = 0; k < n; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        w[2 * k] = (T)cos(a);
        w[2 * k + 1] = (T)sin(a);
      }
      FN(tables)[FN(num_tables)].n = n;
      FN(tables)[FN(num_tables)].w = w;
#pragma omp flush
      FN(num_tables)++;
      res = w;
    }
  }
  return res;
}

/* multiply_complex, src/portfft/common/helpers.hpp:167-172 */
static void FN(cmul)(T ar, T ai, T br, T bi, T* cr, T* ci) {
  *cr = ar * br - ai * bi;
  *ci = ar * bi + ai * br;
}

static void FN(wi_dft)(const T* in, T* out, int n, int stride_in, int stride_out, T* scratch);

/* naive_dft, src/portfft/common/workitem.hpp:64-89 */
static void FN(naive_dft)(const T* in, T* out, int n, int stride_in, int stride_out, T* scratch) {
  const T* W = FN(get_table)(n);
  for (int o = 0; o < n; o++) {
    scratch[2 * o] = 0;
    scratch[2 * o + 1] = 0;
    for (int i = 0; i < n; i++) {
      T tr, ti;
      const T wr = W[2 * (i * o % n)], wi = W[2 * (i * o % n) + 1];
      FN(cmul)(in[2 * i * stride_in], in[2 * i * stride_in + 1], wr, wi, &tr, &ti);
      scratch[2 * o] += tr;
      scratch[2 * o + 1] += ti;
    }
  }
  for (int o = 0; o < n; o++) {
    out[2 * o * stride_out] = scratch[2 * o];
    out[2 * o * stride_out + 1] = scratch[2 * o + 1];
  }
}

/* cooley_tukey_dft, src/portfft/common/workitem.hpp:105-127 */
static void FN(ct_dft)(const T* in, T* out, int fn, int fm, int stride_in, int stride_out, T* scratch) {
  const T* W = FN(get_table)((long long)fn * fm);
  for (int i = 0; i < fm; i++) {
    FN(wi_dft)(in + 2 * i * stride_in, scratch + 2 * i * fn, fn, fm * stride_in, 1, scratch + 2 * fn * fm);
    for (int j = 0; j < fn; j++) {
      const T wr = W[2 * (i * j)], wi = W[2 * (i * j) + 1];
      FN(cmul)(scratch[2 * i * fn + 2 * j], scratch[2 * i * fn + 2 * j + 1], wr, wi, &scratch[2 * i * fn + 2 * j],
               &scratch[2 * i * fn + 2 * j + 1]);
    }
  }
  for (int i = 0; i < fn; i++)
    FN(wi_dft)(scratch + 2 * i, out + 2 * i * stride_out, fm, fn, fn * stride_out, scratch + 2 * fn * fm);
}

/* wi_dft, src/portfft/common/workitem.hpp:200-219 (recursion depth unbounded here: scratch is sized by the caller) */
static void FN(wi_dft)(const T* in, T* out, int n, int stride_in, int stride_out, T* scratch) {
  int f0 = ref_factorize(n);
  if (n == 2) {
    T a = in[0] + in[2 * stride_in], b = in[1] + in[2 * stride_in + 1];
    T c = in[0] - in[2 * stride_in];
    out[2 * stride_out + 1] = in[1] - in[2 * stride_in + 1];
    out[0] = a;
    out[1] = b;
    out[2 * stride_out] = c;
  } else if (f0 >= 2 && n / f0 >= 2) {
    FN(ct_dft)(in, out, n / f0, f0, stride_in, stride_out, scratch);
  } else if (n == 1) {
    T a = in[0], b = in[1];
    out[0] = a;
    out[1] = b;
  } else {
    FN(naive_dft)(in, out, n, stride_in, stride_out, scratch);
  }
}

/* sg_dft, src/portfft/common/subgroup.hpp:271-291: N = f_sg * f_wi, "lane" l holds x[l*f_wi + k] in slot k.
 * Per slot: cross-lane DFT of size f_sg (cross_sg_dft :205-216, the same recursive Cooley-Tukey as wi_dft but across
 * lanes), then * W_N^{l k} (:284-287), then a per-lane wi_dft of size f_wi (:289).  Output is transposed: lane l
 * slot j holds X[j*f_sg + l].  Data is addressed with `stride` complex elements between consecutive x. */
static void FN(sg_dft)(T* x, int n, int stride, T* work) {
  int f_sg = ref_factorize_sg(n, 32), f_wi = n / f_sg;
  T* lanes = work;              /* [f_sg][f_wi] complex */
  T* col = lanes + 2 * n;       /* f_sg complex */
  T* scratch = col + 2 * f_sg;  /* recursion scratch */
  const T* W = FN(get_table)(n);
  for (int k = 0; k < f_wi; k++) {
    for (int l = 0; l < f_sg; l++) {
      col[2 * l] = x[2 * (size_t)(l * f_wi + k) * stride];
      col[2 * l + 1] = x[2 * (size_t)(l * f_wi + k) * stride + 1];
    }
    FN(wi_dft)(col, col, f_sg, 1, 1, scratch);
    for (int l = 0; l < f_sg; l++) {
      const T wr = W[2 * (l * k)], wi = W[2 * (l * k) + 1];
      FN(cmul)(col[2 * l], col[2 * l + 1], wr, wi, &lanes[2 * (l * f_wi + k)], &lanes[2 * (l * f_wi + k) + 1]);
    }
  }
  for (int l = 0; l < f_sg; l++) {
    FN(wi_dft)(lanes + 2 * l * f_wi, lanes + 2 * l * f_wi, f_wi, 1, 1, scratch);
    for (int j = 0; j < f_wi; j++) {
      x[2 * (size_t)(j * f_sg + l) * stride] = lanes[2 * (l * f_wi + j)];
      x[2 * (size_t)(j * f_sg + l) * stride + 1] = lanes[2 * (l * f_wi + j) + 1];
    }
  }
}

/* one transform of a size that fits a work-item or a sub-group, in place, strided */
static void FN(small_dft)(T* x, int n, int stride, T* work, int is_double) {
  if (ref_fits_in_wi(n, is_double)) {
    FN(wi_dft)(x, x, n, stride, stride, work);
  } else {
    FN(sg_dft)(x, n, stride, work);
  }
}

/* wg_dft, src/portfft/common/workgroup.hpp:319-346 with dimension_dft :85-286: view x as n x m (row major):
 * n-point DFTs down the columns, * W_N^{i j} (:186-197), scale (:200-207), m-point DFTs along the rows, output
 * transposed X[j*n + i] (workgroup_dispatcher.hpp:247-259). */
static void FN(wg_dft)(const T* in, T* out, int N, T scale, T* work, int is_double) {
  int n = ref_factorize(N), m = N / n;
  T* y = work;
  T* w2 = y + 2 * (size_t)N;
  const T* W = FN(get_table)(N);
  memcpy(y, in, sizeof(T) * 2 * (size_t)N);
  for (int j = 0; j < m; j++) FN(small_dft)(y + 2 * j, n, m, w2, is_double);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < m; j++) {
      const T wr = W[2 * (i * j)], wi = W[2 * (i * j) + 1];
      FN(cmul)(y[2 * (i * m + j)], y[2 * (i * m + j) + 1], wr, wi, &y[2 * (i * m + j)], &y[2 * (i * m + j) + 1]);
      y[2 * (i * m + j)] *= scale;
      y[2 * (i * m + j) + 1] *= scale;
    }
    FN(small_dft)(y + 2 * (size_t)i * m, m, 1, w2, is_double);
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) {
      out[2 * (size_t)(j * n + i)] = y[2 * (i * m + j)];
      out[2 * (size_t)(j * n + i) + 1] = y[2 * (i * m + j) + 1];
    }
}

/* GLOBAL level, src/portfft/dispatcher/global_dispatcher.hpp:312-412 + common/global.hpp:135-170: one factor at a
 * time; factor f over stride-M columns with the inter-factor twiddle W_N^{c k} applied on store (twiddles computed in
 * double then cast, global_dispatcher.hpp:153-161), then the remaining length-M rows, then the transposition. */
static void FN(global_dft)(T* x, long long N, const long long* factors, int nfac, T* work, int is_double) {
  if (nfac == 1) {
    FN(small_dft)(x, (int)N, 1, work, is_double);
    return;
  }
  long long f = factors[0], M = N / f;
  const T* W = FN(get_table)(N);
  for (long long c = 0; c < M; c++) {
    FN(small_dft)(x + 2 * c, (int)f, (int)M, work, is_double);
    for (long long k = 0; k < f; k++) {
      const T wr = W[2 * (c * k)], wi = W[2 * (c * k) + 1];
      FN(cmul)(x[2 * (k * M + c)], x[2 * (k * M + c) + 1], wr, wi, &x[2 * (k * M + c)], &x[2 * (k * M + c) + 1]);
    }
  }
  for (long long k = 0; k < f; k++) FN(global_dft)(x + 2 * k * M, M, factors + 1, nfac - 1, work, is_double);
  /* transpose [f][M] -> [M][f] */
  T* t = work;
  for (long long k = 0; k < f; k++)
    for (long long c = 0; c < M; c++) {
      t[2 * (c * f + k)] = x[2 * (k * M + c)];
      t[2 * (c * f + k) + 1] = x[2 * (k * M + c) + 1];
    }
  memcpy(x, t, sizeof(T) * 2 * (size_t)N);
}

/* Batched packed interleaved C2C, the reference's compute_forward / compute_backward restated on the CPU.
 * backward = conj -> forward -> conj (committed_descriptor_impl.hpp:469-472); scale on the final output (:473). */
int FN(pfft_oracle_fft)(const T* in, T* out, long long N, long long batch, int direction, double scale_d,
                         int num_threads) {
  const int is_double = sizeof(T) == 8;
  long long factors[64];
  int nfac = 0;
  const int level = ref_select_level(N, is_double, 49152, factors, &nfac);
  if (level < 0) return -1;
  const T scale = (T)scale_d;
  int failed = 0;
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel
  {
    T* work = (T*)malloc(sizeof(T) * (size_t)(8 * N + 256));
    T* buf = (T*)malloc(sizeof(T) * 2 * (size_t)N);
    if (!work || !buf) failed = 1;
#pragma omp for schedule(static)
    for (long long b = 0; b < batch; b++) {
      if (failed) continue;
      const T* xi = in + 2 * b * N;
      T* xo = out + 2 * b * N;
      for (long long i = 0; i < N; i++) {
        buf[2 * i] = xi[2 * i];
        buf[2 * i + 1] = direction ? -xi[2 * i + 1] : xi[2 * i + 1];
      }
      T s = scale;
      if (level == 0) {
        FN(wi_dft)(buf, buf, (int)N, 1, 1, work);
      } else if (level == 1) {
        FN(sg_dft)(buf, (int)N, 1, work);
      } else if (level == 2) {
        FN(wg_dft)(buf, buf, (int)N, scale, work, is_double);
        s = 1; /* applied inside, before the row DFTs (workgroup.hpp:200-207) */
      } else {
        FN(global_dft)(buf, N, factors, nfac, work, is_double);
      }
      for (long long i = 0; i < N; i++) {
        xo[2 * i] = buf[2 * i] * s;
        xo[2 * i + 1] = (direction ? -buf[2 * i + 1] : buf[2 * i + 1]) * s;
      }
    }
    free(work);
    free(buf);
  }
  return failed ? -2 : 0;
}

#undef FN
#undef CAT
#undef CAT_
