/*
 * oracle.c - CPU restatement of the reference's algorithms on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and
 * only as the *checker*.  Nothing under gnn_b200/ or custom_sparse_ops.py may
 * import, link or call it: the product path is the CUDA library and fails
 * loudly when that is missing.
 *
 * Parity status: PINNED.  (1) The sampler hand-off / remap functions are checked
 * bit-for-bit against arrays captured from the unmodified reference Python
 * (tests/golden/make_golden.py, run in the build container against
 * /root/reference).  (2) oracle_build_adj and the *_seqfma SpMM variants are
 * checked bit-for-bit against the reference's own CUDA extension
 * (create_coo_tensor, spmm_naive) compiled unmodified into oracle/_ref/ and run
 * on a B200 (tests/golden/make_golden_gpu.py -> tests/golden/ref_gpu_*.npz).
 * The reference repository itself ships no tests or golden vectors
 * (SURVEY.md section 4).
 *
 * All citations are file:line inside /root/reference.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the only fused
 * multiply-adds are the explicit fmaf() calls below).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- *
 * create_coo_tensor: spmm_cpp/cuda_spmm.cu:787-803 (_create_coo_tensor_kernel)
 * and :806-827 (to_coo_tensor).
 *   indices[0][i] = row, indices[1][i] = colidx[i]  (int16 -> int64 widening, :798-799)
 *   value[i] = 1. / (fullrowptr[row+1]-fullrowptr[row]) * normfact[colidx[i]]   (:800)
 * `1.` is a double literal, so the quotient and the product are evaluated in
 * double and rounded once to float on the store.
 * The result is already row-major sorted with unique (row,col) pairs, so the
 * trailing .coalesce() (:825) leaves indices and values unchanged.
 * ------------------------------------------------------------------------- */
void oracle_build_adj(const int32_t *fullrowptr, const int32_t *rowptr, const int16_t *colidx,
                      const float *normfact, int64_t nrows,
                      int64_t *out_rows, int64_t *out_cols, float *out_vals) {
  for (int64_t r = 0; r < nrows; ++r) {
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      out_rows[i] = r;
      out_cols[i] = (int64_t)colidx[i];
      out_vals[i] = (float)(1. / (double)(fullrowptr[r + 1] - fullrowptr[r]) * (double)normfact[colidx[i]]);
    }
  }
}

/* Same formula with exact int32 column ids (K > 32767 does not wrap). */
void oracle_build_adj_i32(const int32_t *fullrowptr, const int32_t *rowptr, const int32_t *colidx,
                          const float *normfact, int64_t nrows,
                          int64_t *out_rows, int64_t *out_cols, float *out_vals) {
  for (int64_t r = 0; r < nrows; ++r) {
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      out_rows[i] = r;
      out_cols[i] = (int64_t)colidx[i];
      out_vals[i] = (float)(1. / (double)(fullrowptr[r + 1] - fullrowptr[r]) * (double)normfact[colidx[i]]);
    }
  }
}

/* ------------------------------------------------------------------------- *
 * Forward SpMM  Y = A.X   (custom_sparse_ops.py:16-28 -> spmm.cpp:23-27 ->
 * cuda_spmm.cu:619-704).  The mathematical result is the plain product with
 * zero rows for empty rows (output zero-initialised, cuda_spmm.cu:626).
 *
 * _f64acc : double accumulation, one rounding to float - the arbiter for the
 *           1e-5 relative tolerance.
 * _seqfma : one float accumulator per output element, nonzeros visited in
 *           storage order, one fused multiply-add each - the summation order of
 *           the reference's deterministic kernel spmm_naive
 *           (cuda_spmm.cu:88-100 and :124-126 with nvcc's default -fmad=true).
 * ------------------------------------------------------------------------- */
void oracle_csr_spmm_f64acc(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                            int64_t M, int64_t D, const float *X, int64_t ldx, float *Y, int64_t ldy) {
  double *acc = (double *)malloc(sizeof(double) * (size_t)(D > 0 ? D : 1));
  for (int64_t r = 0; r < M; ++r) {
    for (int64_t j = 0; j < D; ++j) acc[j] = 0.0;
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const double v = (double)vals[i];
      const float *x = X + (int64_t)colidx[i] * ldx;
      for (int64_t j = 0; j < D; ++j) acc[j] += v * (double)x[j];
    }
    for (int64_t j = 0; j < D; ++j) Y[r * ldy + j] = (float)acc[j];
  }
  free(acc);
}

void oracle_csr_spmm_seqfma(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                            int64_t M, int64_t D, const float *X, int64_t ldx, float *Y, int64_t ldy) {
  for (int64_t r = 0; r < M; ++r) {
    float *y = Y + r * ldy;
    for (int64_t j = 0; j < D; ++j) y[j] = 0.0f;
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const float v = vals[i];
      const float *x = X + (int64_t)colidx[i] * ldx;
      for (int64_t j = 0; j < D; ++j) y[j] = fmaf(v, x[j], y[j]);
    }
  }
}

/* The reference's load-balanced kernel (cuda_spmm.cu:163-253): rows are cut into
 * chunks of 64 nonzeros (NNZ_PER_CHUNK, :9), each chunk is a sequential FMA chain,
 * and chunk partials are atomically added onto a zeroed output in arbitrary order.
 * This restatement adds them in chunk order (one of the orders the GPU can take);
 * rows of <= 64 nonzeros are bit-identical to the GPU, longer rows are not
 * reproducible on the GPU itself. */
void oracle_csr_spmm_chunk64(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                             int64_t M, int64_t D, const float *X, int64_t ldx, float *Y, int64_t ldy) {
  float *part = (float *)malloc(sizeof(float) * (size_t)(D > 0 ? D : 1));
  for (int64_t r = 0; r < M; ++r) {
    float *y = Y + r * ldy;
    for (int64_t j = 0; j < D; ++j) y[j] = 0.0f;
    for (int32_t b = rowptr[r]; b < rowptr[r + 1]; b += 64) {
      int32_t e = b + 64 < rowptr[r + 1] ? b + 64 : rowptr[r + 1];
      for (int64_t j = 0; j < D; ++j) part[j] = 0.0f;
      for (int32_t i = b; i < e; ++i) {
        const float v = vals[i];
        const float *x = X + (int64_t)colidx[i] * ldx;
        for (int64_t j = 0; j < D; ++j) part[j] = fmaf(v, x[j], part[j]);
      }
      for (int64_t j = 0; j < D; ++j) y[j] += part[j];
    }
  }
  free(part);
}

/* ------------------------------------------------------------------------- *
 * Backward SpMM  dX = A^T.G   (custom_sparse_ops.py:30-37).  The reference
 * materialises mat1.transpose(0,1).coalesce() - entries sorted by (col,row) -
 * and runs the forward kernel on it, so for output row c the terms arrive in
 * ascending r.  Walking A's CSR row by row and scattering into dX visits every
 * output element's terms in exactly that order.
 * ------------------------------------------------------------------------- */
void oracle_csr_spmm_t_f64acc(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                              int64_t M, int64_t K, int64_t D, const float *G, int64_t ldg,
                              float *dX, int64_t lddx) {
  double *acc = (double *)calloc((size_t)(K * D > 0 ? K * D : 1), sizeof(double));
  for (int64_t r = 0; r < M; ++r) {
    const float *g = G + r * ldg;
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const double v = (double)vals[i];
      double *a = acc + (int64_t)colidx[i] * D;
      for (int64_t j = 0; j < D; ++j) a[j] += v * (double)g[j];
    }
  }
  for (int64_t c = 0; c < K; ++c)
    for (int64_t j = 0; j < D; ++j) dX[c * lddx + j] = (float)acc[c * D + j];
  free(acc);
}

void oracle_csr_spmm_t_seqfma(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                              int64_t M, int64_t K, int64_t D, const float *G, int64_t ldg,
                              float *dX, int64_t lddx) {
  for (int64_t c = 0; c < K; ++c)
    for (int64_t j = 0; j < D; ++j) dX[c * lddx + j] = 0.0f;
  for (int64_t r = 0; r < M; ++r) {
    const float *g = G + r * ldg;
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const float v = vals[i];
      float *a = dX + (int64_t)colidx[i] * lddx;
      for (int64_t j = 0; j < D; ++j) a[j] = fmaf(v, g[j], a[j]);
    }
  }
}

/* CSR of A^T (== CSC of A), entries of each output row in ascending source row:
 * what mat1.transpose(0,1).coalesce() (custom_sparse_ops.py:34) holds.  perm[i] is
 * the CSR position the i-th transposed entry came from. */
void oracle_csr_transpose(const int32_t *rowptr, const int32_t *colidx, int64_t M, int64_t K,
                          int32_t *t_rowptr, int32_t *t_colidx, int32_t *perm) {
  for (int64_t c = 0; c <= K; ++c) t_rowptr[c] = 0;
  const int32_t nnz = rowptr[M];
  for (int32_t i = 0; i < nnz; ++i) t_rowptr[colidx[i] + 1]++;
  for (int64_t c = 0; c < K; ++c) t_rowptr[c + 1] += t_rowptr[c];
  int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));
  for (int64_t c = 0; c < K; ++c) cur[c] = t_rowptr[c];
  for (int64_t r = 0; r < M; ++r)
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const int32_t p = cur[colidx[i]]++;
      t_colidx[p] = (int32_t)r;
      perm[p] = i;
    }
  free(cur);
}

/* sorted COO rows -> CSR row pointer (cuda_spmm.cu:255-265, _calc_rowptr). */
void oracle_coo_rows_to_rowptr(const int64_t *rows, int64_t nnz, int64_t M, int32_t *rowptr) {
  for (int64_t r = 0; r <= M; ++r) rowptr[r] = 0;
  for (int64_t i = 0; i < nnz; ++i) rowptr[rows[i] + 1]++;
  for (int64_t r = 0; r < M; ++r) rowptr[r + 1] += rowptr[r];
}

/* ------------------------------------------------------------------------- *
 * sampled_nodes remap, sampler.py:143:
 *   np.where(np.in1d(after_nodes, previous_nodes))[0]
 * after_nodes is sorted unique (np.unique, sampler.py:131); previous_nodes is any
 * list of ids.  Returns the count written to out.
 * ------------------------------------------------------------------------- */
static int cmp_i64(const void *a, const void *b) {
  const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}

int64_t oracle_sampled_nodes(const int64_t *after_nodes, int64_t n_after,
                             const int64_t *previous_nodes, int64_t n_prev, int64_t *out) {
  int64_t *sorted = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_prev > 0 ? n_prev : 1));
  memcpy(sorted, previous_nodes, sizeof(int64_t) * (size_t)n_prev);
  qsort(sorted, (size_t)n_prev, sizeof(int64_t), cmp_i64);
  int64_t n = 0;
  for (int64_t i = 0; i < n_after; ++i)
    if (bsearch(&after_nodes[i], sorted, (size_t)n_prev, sizeof(int64_t), cmp_i64)) out[n++] = i;
  free(sorted);
  return n;
}

/* ------------------------------------------------------------------------- *
 * placement remap, sampler.py:150-158.  For input node j:
 *   dev = device_id_of_nodes[input_nodes[j]]                (:152)
 *   dev == -1        -> host row input_nodes[j]             (:153-154)
 *   dev == devices[i]-> row idx_of_nodes_on_device[node] of GPU i's buffer (:156-158)
 * src_dev[j] is the *index into devices[]* (-1 for host), slot[j] the row.
 * A device id that is neither -1 nor in devices[] is left as -2 (the reference
 * would silently leave such a row uninitialised, main.py:129-134).
 * ------------------------------------------------------------------------- */
void oracle_placement_remap(const int64_t *input_nodes, int64_t n0,
                            const int64_t *device_id_of_nodes, const int64_t *idx_of_nodes_on_device,
                            const int64_t *devices, int64_t world,
                            int32_t *src_dev, int64_t *slot) {
  for (int64_t j = 0; j < n0; ++j) {
    const int64_t node = input_nodes[j];
    const int64_t dev = device_id_of_nodes[node];
    if (dev == -1) { src_dev[j] = -1; slot[j] = node; continue; }
    src_dev[j] = -2; slot[j] = -1;
    for (int64_t i = 0; i < world; ++i)
      if (devices[i] == dev) { src_dev[j] = (int32_t)i; slot[j] = idx_of_nodes_on_device[node]; }
  }
}

/* ------------------------------------------------------------------------- *
 * feature gather, main.py:129-134: row j of the gathered buffer is the feature
 * row of input node j, taken from GPU i's buffer (:132) or the host table (:134).
 * bases[i] for i in [0,world) are the per-GPU buffers, bases[world] the host table;
 * all have leading dimension ld_src (floats).  Bit-exact copy.
 * ------------------------------------------------------------------------- */
void oracle_gather_rows(const float *const *bases, int64_t world, const int32_t *src_dev,
                        const int64_t *slot, int64_t n0, int64_t F, int64_t ld_src,
                        float *out, int64_t ld_out) {
  for (int64_t j = 0; j < n0; ++j) {
    const int32_t s = src_dev[j];
    if (s < -1) continue;
    const float *base = bases[s < 0 ? world : s];
    memcpy(out + j * ld_out, base + slot[j] * ld_src, sizeof(float) * (size_t)F);
  }
}

/* fused gather + SpMM (adjs[0]): Y = A . gather(...)  in double accumulation. */
void oracle_gather_spmm_f64acc(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                               int64_t M, int64_t D, const float *const *bases, int64_t world,
                               const int32_t *src_dev, const int64_t *slot, int64_t ld_src,
                               float *Y, int64_t ldy) {
  double *acc = (double *)malloc(sizeof(double) * (size_t)(D > 0 ? D : 1));
  for (int64_t r = 0; r < M; ++r) {
    for (int64_t j = 0; j < D; ++j) acc[j] = 0.0;
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const int32_t c = colidx[i];
      const int32_t s = src_dev[c];
      const float *x = bases[s < 0 ? world : s] + slot[c] * ld_src;
      const double v = (double)vals[i];
      for (int64_t j = 0; j < D; ++j) acc[j] += v * (double)x[j];
    }
    for (int64_t j = 0; j < D; ++j) Y[r * ldy + j] = (float)acc[j];
  }
  free(acc);
}
