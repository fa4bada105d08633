/*
 * geot_oracle.c -- CPU restatement of GeoT's segment-reduction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker / the timed CPU baseline.  The product path (geot_b200/) never imports,
 * links or falls back to this code.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
 *   - the reference's own sequential golden loops compiled from the sources where they lie
 *     (/root/reference/csrc/util/check.cuh:77-111 via oracle/ref_seq_shim.cu -> oracle/_ref/),
 *   - the formulas of the reference's Python tests (test/test_index_scatter.py:16-22,
 *     test/test_gather_weight_scatter.py:4-11, test/test_mh_spmm.py:4-10) as committed golden
 *     vectors under tests/golden/ (generator: tests/golden/make_golden.py),
 *   - the reference CPU kernel's segment-pointer pass (csrc/cpu/index_scatter_cpu.cpp:36-75),
 *     observed through the unmodified reference library built into oracle/_ref/.
 * mean/max/min/prod, fp64 and bf16 are not exercised by any reference test; for those the
 * pin is torch.scatter_reduce(include_self=False), the semantics the reference CPU kernel
 * states (csrc/cpu/index_scatter_cpu.cpp:95-119, ATen ReduceUtils.h init/update/write).
 *
 * All functions follow the reference's sequential definition
 *     dst[dst_index[e], :] (op)= weight[e] * src[src_index[e], :]        e = 0 .. E-1, in order
 * (check.cuh:77-85 segment_coo_sequencial, :101-111 gws_sequencial; mh_spmm_kernel.cuh:41,66 for
 * the per-head weight), with rows that receive no edge left at 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { RED_SUM = 0, RED_MEAN = 1, RED_MAX = 2, RED_MIN = 3, RED_PROD = 4 };

/* ---- segment pointers --------------------------------------------------------------------- */

/* Unique keys + start offsets of a sorted index.  Follows the reference CPU kernel's passes 1-2
 * (csrc/cpu/index_scatter_cpu.cpp:36-75): row_index[m] = m-th distinct key, row_offset[m] = first
 * position of that key, row_offset[M] = E.  Returns M (number of non-empty rows).  Buffers must
 * hold E and E+1 entries (worst case). */
int64_t geot_oracle_segment_ptr(const int64_t *index, int64_t E, int64_t *row_index,
                                int64_t *row_offset) {
  if (E <= 0) return 0;
  int64_t m = 0;
  row_index[0] = index[0];
  row_offset[0] = 0;
  m = 1;
  for (int64_t i = 1; i < E; ++i) {
    if (index[i] != index[i - 1]) {
      row_index[m] = index[i];
      row_offset[m] = i;
      ++m;
    }
  }
  row_offset[m] = E;
  return m;
}

/* CSR row pointer of a sorted COO row index: histogram then exclusive cumsum, as
 * geot::coo_to_csr does (geot/match_replace/format_transform.py:5-18, test/test_csr_gws.py:6-12).
 * rowptr has S+1 entries. */
void geot_oracle_rowptr(const int64_t *index, int64_t E, int64_t S, int64_t *rowptr) {
  memset(rowptr, 0, (size_t)(S + 1) * sizeof(int64_t));
  for (int64_t i = 0; i < E; ++i) rowptr[index[i] + 1] += 1;
  for (int64_t r = 0; r < S; ++r) rowptr[r + 1] += rowptr[r];
}

/* ---- the reduction ------------------------------------------------------------------------ */

#define NAN_MAX(a, b) (((a) != (a)) ? (a) : (((b) != (b)) ? (b) : ((a) > (b) ? (a) : (b))))
#define NAN_MIN(a, b) (((a) != (a)) ? (a) : (((b) != (b)) ? (b) : ((a) < (b) ? (a) : (b))))

/* One generic body, instantiated for (storage type T, accumulator type A).
 *   src        [N_src, H*F]     row-major
 *   src_index  [E] or NULL  (NULL: src row = e, i.e. index_scatter)
 *   dst_index  [E]          (any order; the reference's unsorted kernel has the same math)
 *   weight     NULL, or element (e,h) at weight[e*ws_e + h*ws_h]
 *                 gather_weight_scatter: H=1, ws_e=1, ws_h=0
 *                 mh_spmm   [E,H]:       ws_e=H, ws_h=1      (mh_spmm_kernel.cuh:66)
 *                 mh_spmm   [H,E]:       ws_e=1, ws_h=E      (mh_spmm_kernel.cuh:168)
 *   dst        [S, H*F], fully overwritten; rows with no edge are 0
 * The product weight*src is rounded to A before accumulation (two roundings, no fma: build with
 * -ffp-contract=off), as torch's mul -> index_add_ does in the reference's tests.
 */
#define DEFINE_REDUCE(NAME, T, A)                                                               \
  void NAME(const T *src, const int64_t *src_index, const int64_t *dst_index, const T *weight,  \
            T *dst, int64_t E, int64_t S, int64_t F, int64_t H, int64_t ws_e, int64_t ws_h,     \
            int reduce) {                                                                       \
    const int64_t W = H * F;                                                                    \
    A *acc = (A *)calloc((size_t)(S > 0 ? S : 1) * (size_t)W, sizeof(A));                       \
    int64_t *cnt = (int64_t *)calloc((size_t)(S > 0 ? S : 1), sizeof(int64_t));                 \
    for (int64_t e = 0; e < E; ++e) {                                                           \
      const int64_t r = dst_index[e];                                                           \
      const int64_t s = src_index ? src_index[e] : e;                                           \
      const T *x = src + s * W;                                                                 \
      A *a = acc + r * W;                                                                       \
      const int first = (cnt[r] == 0);                                                          \
      cnt[r] += 1;                                                                              \
      for (int64_t h = 0; h < H; ++h) {                                                         \
        const A w = weight ? (A)weight[e * ws_e + h * ws_h] : (A)1;                             \
        for (int64_t j = 0; j < F; ++j) {                                                       \
          const int64_t c = h * F + j;                                                          \
          const A v = weight ? (A)((A)x[c] * w) : (A)x[c];                                      \
          if (first) {                                                                          \
            a[c] = v;                                                                           \
          } else if (reduce == RED_SUM || reduce == RED_MEAN) {                                 \
            a[c] = a[c] + v;                                                                    \
          } else if (reduce == RED_MAX) {                                                       \
            a[c] = NAN_MAX(a[c], v);                                                            \
          } else if (reduce == RED_MIN) {                                                       \
            a[c] = NAN_MIN(a[c], v);                                                            \
          } else {                                                                              \
            a[c] = a[c] * v;                                                                    \
          }                                                                                     \
        }                                                                                       \
      }                                                                                         \
    }                                                                                           \
    for (int64_t r = 0; r < S; ++r) {                                                           \
      for (int64_t c = 0; c < W; ++c) {                                                         \
        A v = acc[r * W + c];                                                                   \
        if (reduce == RED_MEAN && cnt[r] > 0) v = v / (A)cnt[r];                                \
        dst[r * W + c] = (T)v;                                                                  \
      }                                                                                         \
    }                                                                                           \
    free(acc);                                                                                  \
    free(cnt);                                                                                  \
  }

/* f32 storage, f32 accumulate: the reference's own arithmetic (sequential order). */
DEFINE_REDUCE(geot_oracle_reduce_f32, float, float)
/* f32 storage, f64 accumulate: the tight-tolerance oracle for sum/mean. */
DEFINE_REDUCE(geot_oracle_reduce_f32_acc64, float, double)
/* f64 storage and accumulate. */
DEFINE_REDUCE(geot_oracle_reduce_f64, double, double)

/* ---- multi-threaded variant, for the timed CPU baseline only -------------------------------- */

/* Sorted dst_index only.  Rows are independent, so threads split the row range; inside a row the
 * edge order is the sequential one, hence results equal geot_oracle_reduce_f32 bit for bit.
 * Same structure as the reference CPU kernel's pass 3 (index_scatter_cpu.cpp:89-121) with the
 * src row taken from the edge (not from index[n], the reference's defect -- SURVEY 8a A5). */
int geot_oracle_reduce_f32_mt(const float *src, const int64_t *src_index, const int64_t *dst_index,
                              const float *weight, float *dst, int64_t E, int64_t S, int64_t F,
                              int64_t H, int64_t ws_e, int64_t ws_h, int reduce) {
  const int64_t W = H * F;
  int64_t *rowptr = (int64_t *)malloc((size_t)(S + 1) * sizeof(int64_t));
  geot_oracle_rowptr(dst_index, E, S, rowptr);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int64_t r = 0; r < S; ++r) {
    float *a = dst + r * W;
    const int64_t b = rowptr[r], t = rowptr[r + 1];
    if (b == t) {
      memset(a, 0, (size_t)W * sizeof(float));
      continue;
    }
    for (int64_t e = b; e < t; ++e) {
      const int64_t s = src_index ? src_index[e] : e;
      const float *x = src + s * W;
      for (int64_t h = 0; h < H; ++h) {
        const float w = weight ? weight[e * ws_e + h * ws_h] : 1.0f;
        for (int64_t j = 0; j < F; ++j) {
          const int64_t c = h * F + j;
          const float v = weight ? x[c] * w : x[c];
          if (e == b) a[c] = v;
          else if (reduce == RED_SUM || reduce == RED_MEAN) a[c] += v;
          else if (reduce == RED_MAX) a[c] = NAN_MAX(a[c], v);
          else if (reduce == RED_MIN) a[c] = NAN_MIN(a[c], v);
          else a[c] *= v;
        }
      }
    }
    if (reduce == RED_MEAN) {
      const float n = (float)(t - b);
      for (int64_t c = 0; c < W; ++c) a[c] /= n;
    }
  }
  free(rowptr);
  return nthreads;
}

/* Threads of the timed CPU baseline.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently time a
 * single-threaded baseline: the bench sets the count explicitly and reports what the runtime really uses. */
int geot_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* ---- SDDMM on a COO edge list ---------------------------------------------------------------- */

/* out[e] = < mat1[row[e], :], mat2[col[e], :] >.  Follows the reference kernel's definition
 * (csrc/cuda/sddmm_coo_kernel.cuh:44-71, the scalar tail path: offset1 = S_cooRowInd[eid]*D_kcols,
 * offset2 = S_cooColInd[eid]*D_kcols, multi += D1[offset1+c]*D2[offset2+c]) with row = dst_index and
 * col = src_index as the launcher binds them (csrc/cuda/gather_weight_scatter_cuda.cu:46-50).
 * Accumulates in double: the tight-tolerance oracle. */
void geot_oracle_sddmm_f32(const float *mat1, const int64_t *row, const float *mat2, const int64_t *col,
                           float *out, int64_t E, int64_t F) {
  for (int64_t e = 0; e < E; ++e) {
    const float *x = mat1 + row[e] * F, *y = mat2 + col[e] * F;
    double s = 0.0;
    for (int64_t c = 0; c < F; ++c) s += (double)x[c] * (double)y[c];
    out[e] = (float)s;
  }
}
void geot_oracle_sddmm_f64(const double *mat1, const int64_t *row, const double *mat2, const int64_t *col,
                           double *out, int64_t E, int64_t F) {
  for (int64_t e = 0; e < E; ++e) {
    const double *x = mat1 + row[e] * F, *y = mat2 + col[e] * F;
    double s = 0.0;
    for (int64_t c = 0; c < F; ++c) s += x[c] * y[c];
    out[e] = s;
  }
}

/* ---- CSR SpMM (csr_gws) ------------------------------------------------------------------------ */

/* out[r, :] = sum_{e in [rowptr[r], rowptr[r+1])} val[e] * src[colind[e], :] for r < nrow, and one extra
 * zero row: the reference allocates indptr.size(0) = nrow + 1 output rows (csrc/csr_gws.cpp:29-31) and its
 * kernel writes rows < nrow only (csr_gws_kernel.cuh:13-187).  Sequential edge order; f64 accumulate. */
void geot_oracle_csr_gws_f32(const int64_t *rowptr, const int64_t *colind, const float *val, const float *src,
                             float *out, int64_t nrow, int64_t F) {
  for (int64_t r = 0; r <= nrow; ++r) {
    for (int64_t c = 0; c < F; ++c) {
      double s = 0.0;
      if (r < nrow)
        for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) s += (double)((float)(val[e] * src[colind[e] * F + c]));
      out[r * F + c] = (float)s;
    }
  }
}

/* row index of every nonzero of a CSR matrix (the inverse of geot_oracle_rowptr) */
void geot_oracle_csr_to_coo(const int64_t *rowptr, int64_t nrow, int64_t *row) {
  for (int64_t r = 0; r < nrow; ++r)
    for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) row[e] = r;
}
