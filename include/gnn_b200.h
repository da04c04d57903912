/*
 * gnn_b200.h - C ABI of the B200-native LADIES-layer SpMM / feature-gather path.
 *
 * Drop-in boundary for HPC-Research-Lab/GNN's `spmm_cpp` extension and the
 * gather block of its training loop.  Every entry point takes plain device (or
 * mapped host / peer) pointers, sizes, leading dimensions in ELEMENTS and a
 * cudaStream_t passed as void*; there are no torch types here.  All functions
 * are asynchronous on `stream`, never synchronise the device, never allocate,
 * and are re-entrant (the reference is entered concurrently from trainer and
 * sampler threads).  The current device must be the one
 * that owns the output buffers.
 *
 * Return value: 0 on success; a positive cudaError_t value if the CUDA runtime
 * reported an error; a negative GNN_E_* code for argument errors.  The reference
 * calls exit(-1) on CUDA errors (cuda_spmm.cu:16-24); this library returns.
 *
 * Reference interface replaced by each function is cited as file:line relative
 * to the reference repository.
 */
#ifndef GNN_B200_H_
#define GNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNN_B200_ABI_VERSION 4

#define GNN_E_BADARG   (-1)   /* null pointer, negative size, unsupported width */
#define GNN_E_WORKSPACE (-2)  /* workspace missing or too small */
#define GNN_E_RANGE    (-3)   /* size exceeds the 32-bit index limits of the path */
#define GNN_E_DRIVER   (-4)   /* a CUDA driver entry point is missing or rejected a descriptor (TMA tensor map) */

typedef void *gnn_stream_t;   /* cudaStream_t */

/* ABI version of the loaded library (GNN_B200_ABI_VERSION at build time). */
int gnn_abi_version(void);

/* Human-readable text for a return code of this library. */
const char *gnn_error_string(int code);

/* Number of kernels this library has launched in the calling process so far
 * (all threads; used by bench.py for its `gpu_launches` claim). */
int64_t gnn_launch_count(void);

/* Tell the SpMM planner how many CTA slots kernels on OTHER streams hold while it runs (process-wide; default 0;
 * returns the previous value, GNN_E_BADARG for a negative count).  The SpMM sizes its grid to fill the GPU in exactly
 * one wave; a long-running CTA of a co-running kernel - the host-row gather of the next minibatch, which the
 * reference issues next to the step from its sampler threads (sampler.py:135-139, main.py:129-134) - would push one
 * SpMM CTA into a second wave, so the plan leaves `ctas` slots out.  gnn_host_gather_ctas() is the grid of the
 * host-row gather (gnn_gather_rows_src_f32 with only_src == -1), i.e. the value a prefetching caller passes. */
int gnn_set_corunner_ctas(int ctas);
int gnn_host_gather_ctas(void);

/* Host threads that wait in a stream / event synchronise of the CURRENT device sleep (on == 1, cudaDeviceScheduleBlockingSync),
 * spin but yield their core between polls (on == 2, cudaDeviceScheduleYield) or spin (0, the CUDA default).  For processes that run more waiting threads than they have host cores: the reference's
 * sampler pool (main.py:77, --pool_num threads per GPU) next to the trainer thread. */
int gnn_set_blocking_sync(int on);

/* ---------------------------------------------------------------------------
 * gnn_build_adj - sampled CSR + LADIES weights -> COO (API) + CSR (kernels).
 *
 * Replaces create_coo_tensor: spmm_cpp/spmm.cpp:44-50, cuda_spmm.cu:787-827
 * (called from sampler.py:139).  For every stored entry i of row r:
 *     out_indices[0*nnz + i] = r
 *     out_indices[1*nnz + i] = colidx[i]                 (widened to int64)
 *     out_vals[i] = (float)( (1.0 / (double)(fullrowptr[r+1]-fullrowptr[r]))
 *                            * (double)normfact[colidx[i]] )      (cuda_spmm.cu:800)
 *     out_colidx32[i] = colidx[i]                        (int32 copy for the SpMM kernels)
 *     out_rowidx32[i] = r                                (row id per entry: lets the short-row SpMM kernels skip the
 *                                                         row search, see gnn_csr_spmm_f32_ex)
 * colidx is int16 when colidx_bytes == 2 (what sampler.py:136 uploads; read as
 * signed 16-bit exactly like cuda_spmm.cu:792) or int32 when colidx_bytes == 4.
 * out_indices, out_colidx32 and out_rowidx32 may each be NULL to skip that output.
 * Rows are already sorted and (row,col) pairs unique, so the result is coalesced
 * without the reference's trailing .coalesce() (cuda_spmm.cu:825).
 * ------------------------------------------------------------------------- */
int gnn_build_adj(const int32_t *fullrowptr, const int32_t *rowptr, const void *colidx, int colidx_bytes,
                  const float *normfact, int64_t M, int64_t K, int64_t nnz,
                  int64_t *out_indices, float *out_vals, int32_t *out_colidx32, int32_t *out_rowidx32,
                  gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_coo_to_csr - coalesced COO (int64 [2,nnz], row-major sorted) -> CSR int32.
 *
 * Replaces the per-call COO->CSR rebuild of spmm_cuda_v2: the int64->int32 casts
 * (cuda_spmm.cu:620-621) and _calc_rowptr (cuda_spmm.cu:255-265).  Used only for
 * sparse tensors that were not produced by gnn_build_adj.
 * ------------------------------------------------------------------------- */
int gnn_coo_to_csr(const int64_t *indices, int64_t M, int64_t nnz,
                   int32_t *out_rowptr, int32_t *out_colidx32, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_csr_spmm_f32 - Y[M,D] = A[M,K] . X[K,D], fp32, A in CSR.
 *
 * Replaces spmm_load_balance / spmm_naive: spmm.cpp:23-27,38-42 ->
 * spmm_cuda_v2 / spmm_cuda_v1, cuda_spmm.cu:619-704 / :134-160 (forward of
 * custom_sparse_ops.py:16-28).  Every row of Y is written (empty rows get zeros,
 * matching the zero-initialised output of cuda_spmm.cu:626).  Rows longer than
 * the internal chunk size are summed by a fixed-order two-level reduction, so
 * results are bit-reproducible run to run (the reference's atomicAdd order,
 * cuda_spmm.cu:205-209, is not).
 *
 * workspace: gnn_csr_spmm_workspace_bytes(M, nnz, D) bytes of device memory,
 * contents undefined on entry, private to this call until it completes.
 * ------------------------------------------------------------------------- */
size_t gnn_csr_spmm_workspace_bytes(int64_t M, int64_t nnz, int64_t D);

int gnn_csr_spmm_f32(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                     int64_t M, int64_t K, int64_t nnz, int64_t D,
                     const float *X, int64_t ldx, float *Y, int64_t ldy,
                     void *workspace, size_t workspace_bytes, gnn_stream_t stream);

/* Same product with the two workspace regions passed separately, for callers that keep ONE long-lived workspace per
 * stream (spmm_ext.cpp does): `counters` (gnn_csr_spmm_counter_bytes) holds one arrival counter per (row, column slab)
 * for rows whose nonzeros span several warp chunks, `partials` (gnn_csr_spmm_partial_bytes) their partial sums.
 * The counters wrap back to zero inside the kernel, so a counter region that is all-zero on entry is all-zero again
 * when the call has completed: pass GNN_SPMM_COUNTERS_ZEROED to skip the memset (one launch per SpMM instead of the
 * reference's >= 8, cuda_spmm.cu:619-704).  Without the flag the contents of both regions are undefined on entry.
 * The regions must not be shared by calls that may run concurrently (different streams).
 * `rowidx` (optional, may be NULL): int32 row id of every stored entry, as gnn_build_adj / gnn_csr_transpose emit
 * them.  Short-row blocks (top LADIES layer, sparse graphs) are latency-bound: with row ids a warp needs one round trip
 * - (col, val, row) of its chunk - before its X loads go out, instead of a search over the row pointer. */
#define GNN_SPMM_COUNTERS_ZEROED 1u
size_t gnn_csr_spmm_counter_bytes(int64_t M, int64_t nnz, int64_t D);
size_t gnn_csr_spmm_partial_bytes(int64_t M, int64_t nnz, int64_t D);

int gnn_csr_spmm_f32_ex(const int32_t *rowptr, const int32_t *rowidx, const int32_t *colidx, const float *vals,
                        int64_t M, int64_t K, int64_t nnz, int64_t D,
                        const float *X, int64_t ldx, float *Y, int64_t ldy,
                        int32_t *counters, void *partials, size_t partial_bytes, unsigned flags, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_csr_spmm_t_f32 - transpose-free backward: dX[K,D] = A^T . G from A's own CSR.
 *
 * Replaces custom_sparse_ops.py:30-37 (transpose(0,1).coalesce() + the forward pipeline) without any transposed
 * index: dX is zero-filled, then every stored entry (r, c, v) adds v * G[r, :] into dX[c, :] with vector reductions
 * (red.global.add.v4.f32 when G, dX and both leading dimensions are 16-byte aligned, scalar reductions otherwise).
 * The additions of one output row arrive in no fixed order: results are reproducible to rounding only, like the
 * reference's atomicAdd kernel (cuda_spmm.cu:205-209).  Faster than gnn_csr_transpose + gnn_csr_spmm_f32 for sparse
 * layers (the index build is four dependent launches), slower on dense LADIES blocks (profiles/ has the table).
 * rowidx: optional row id per entry (see gnn_csr_spmm_f32_ex).  No workspace.
 * ------------------------------------------------------------------------- */
int gnn_csr_spmm_t_f32(const int32_t *rowptr, const int32_t *rowidx, const int32_t *colidx, const float *vals,
                       int64_t M, int64_t K, int64_t nnz, int64_t D,
                       const float *G, int64_t ldg, float *dX, int64_t lddx, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_gather_spmm_f32 - fused input-feature gather + SpMM for the deepest layer:
 *     Y[M,D] = A[M,K] . Xg,   Xg[j,:] = *(xrows[j])   (row j never materialised)
 *
 * Replaces main.py:129-134 followed by the first spmm of models.py:18 / :60.
 * xrows[j] points at the feature row of input node j wherever the placement put
 * it (local HBM, a peer GPU's shard mapped over NVLink, or mapped pinned host
 * memory); build it with gnn_placement_remap.  Every A nonzero reads its X row
 * through the table, so use this only when rows are local or nnz/K is small;
 * otherwise stage with gnn_gather_rows_f32 first.
 * ------------------------------------------------------------------------- */
int gnn_gather_spmm_f32(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                        int64_t M, int64_t K, int64_t nnz, int64_t D,
                        const float *const *xrows, float *Y, int64_t ldy,
                        void *workspace, size_t workspace_bytes, gnn_stream_t stream);
int gnn_gather_spmm_f32_ex(const int32_t *rowptr, const int32_t *rowidx, const int32_t *colidx, const float *vals,
                           int64_t M, int64_t K, int64_t nnz, int64_t D,
                           const float *const *xrows, float *Y, int64_t ldy,
                           int32_t *counters, void *partials, size_t partial_bytes, unsigned flags, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_csr_transpose - CSR of A^T (entries of each output row in ascending source
 * row) built on the device, once per adjacency.
 *
 * Replaces mat1.transpose(0,1).coalesce() of every backward call,
 * custom_sparse_ops.py:34 (an index swap plus a full sort of nnz keys).
 * The backward product dX = A^T.G is then gnn_csr_spmm_f32 on the result.
 * Deterministic: no atomics decide the order.
 *
 * Precondition: (row, column) pairs are unique, i.e. the input is coalesced (what spmm.cpp:12-13 demands of every
 * sparse operand); a duplicate pair would set the same bitmap bit twice and corrupt the positions after it.
 *
 * workspace: gnn_csr_transpose_workspace_bytes(M, K, nnz) bytes: two K-int arrays plus a column-major row bitmap
 * of K * ceil(rows/32) * 8 bytes.  The bitmap never exceeds the budget (default 512 MiB): larger matrices
 * (products / papers-scale samp_num, ~130 K x 130 K and up) are transposed in blocks of rows, block after block,
 * with a per-column cursor - same result, more launches.  gnn_set_transpose_budget changes the budget
 * (process-wide; 0 restores the default; returns the previous value); query the workspace size after setting it.
 * ------------------------------------------------------------------------- */
int64_t gnn_set_transpose_budget(int64_t bytes);
size_t gnn_csr_transpose_workspace_bytes(int64_t M, int64_t K, int64_t nnz);

int gnn_csr_transpose(const int32_t *rowptr, const int32_t *colidx, const float *vals,
                      int64_t M, int64_t K, int64_t nnz,
                      int32_t *t_rowptr, int32_t *t_colidx, float *t_vals, int32_t *t_rowidx /* optional: row id per entry of A^T */,
                      void *workspace, size_t workspace_bytes, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_placement_remap - per-minibatch placement remap on the device.
 *
 * Replaces sampler.py:150-158 (device_id_of_nodes[previous_nodes], one boolean
 * mask and one slot list per source device, host mask + ids).  For input node j:
 *     dev = device_id_of_nodes[input_nodes[j]]
 *     dev == -1          -> src_dev[j] = -1, slot[j] = input_nodes[j]      (host table row)
 *     dev == devices[i]  -> src_dev[j] =  i, slot[j] = idx_of_nodes_on_device[input_nodes[j]]
 *     otherwise          -> src_dev[j] = -2, slot[j] = -1
 * and, when xrows != NULL,
 *     xrows[j] = bases[i] + slot[j]*ld_src     (GPU shards; leading dimension ld_src)
 *     xrows[j] = bases[world] + slot[j]*ld_host (host table; its own leading dimension, so an unpadded table can be
 *                                                registered as it is instead of being copied into a padded one)
 * or NULL for src_dev -2 and for a source whose base pointer is NULL (gather kernels skip NULL rows).  The tables are the per-rank views produced by
 * create_buffer (preprocess.py:311-407), resident on the device as int64.
 * counts (optional, int64 [world+2], zeroed by this call) receives the number of
 * rows per source: counts[i] for GPU i, counts[world] host, counts[world+1] invalid.
 * ------------------------------------------------------------------------- */
int gnn_placement_remap(const int64_t *input_nodes, int64_t n0,
                        const int64_t *device_id_of_nodes, const int64_t *idx_of_nodes_on_device,
                        const int64_t *devices, int64_t world,
                        const float *const *bases, int64_t ld_src, int64_t ld_host,
                        int32_t *src_dev, int64_t *slot, const float **xrows, int64_t *counts,
                        gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_gather_rows_f32 - out[j, 0:F] = xrows[j][0:F] for j in [0,n0), bit-exact.
 *
 * Replaces the gather block main.py:129-134 (and its copies :185-190, :228-233):
 * per source device a gather kernel on the *remote* GPU, a peer copy and a
 * boolean-mask index_put, plus a CPU gather + synchronous H2D for host rows.
 * Here one kernel on the consuming GPU pulls every row once through its pointer
 * (local HBM, peer over NVLink, mapped pinned host).  Rows with a NULL pointer
 * are left untouched.
 * ------------------------------------------------------------------------- */
int gnn_gather_rows_f32(const float *const *xrows, int64_t n0, int64_t F,
                        float *out, int64_t ld_out, gnn_stream_t stream);

/* Same gather restricted to the rows j with src_dev[j] == only_src (-1 = host rows, i = GPU i's rows,
 * GNN_SRC_DEVICES = every row held by some GPU), so transfers from different sources can be put on
 * different streams.  Host rows are PCIe-bound and use a deliberately small grid. */
#define GNN_SRC_DEVICES (-100)
/* GNN_SRC_NOT(i): every valid row NOT held by source i (e.g. everything that is not in the local shard);
 * the macro is its own inverse: GNN_SRC_NOT(GNN_SRC_NOT(i)) == i. */
#define GNN_SRC_NOT(i) (-200 - (i))
/* GNN_SRC_PEERS(i): every row held by a GPU other than i (one launch pulls all peer shards over NVLink). */
#define GNN_SRC_PEERS(i) (-100000 - (i))
int gnn_gather_rows_src_f32(const float *const *xrows, const int32_t *src_dev, int32_t only_src,
                            int64_t n0, int64_t F, float *out, int64_t ld_out, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_index_rows_f32 - out[i, 0:F] = X[idx[i], 0:F]  (the x[sampled_nodes] gather of
 * models.py:19, with the int64 row-index remap of sampler.py:143 resident on device).
 * ------------------------------------------------------------------------- */
int gnn_index_rows_f32(const float *X, int64_t ldx, const int64_t *idx, int64_t n, int64_t F,
                       float *out, int64_t ld_out, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * LADIES layer construction on the device - the array work of the sampler.
 * The weighted draw without replacement stays on the host (numpy's legacy np.random.choice algorithm on the same
 * MT19937 stream and the exact same probabilities: gnn_legacy_choice_f64 / gnn_ladies_layer_host below), so sampled
 * node sets remain bit-identical to the reference; these entry points replace the scipy/numpy array passes around it.  The graph structure (indptr int64 [N+1], indices int32 [nnz])
 * is resident on the device.
 *
 * gnn_row_slice_count  : U = lap_matrix[nodes, :]  (sampler.py:113-114): out_fullrowptr[M+1] = row pointer of
 *                        the selected rows (exclusive scan of their full-graph degrees); scratch_lens int32 [M].
 * gnn_row_slice_fill   : column ids of every selected row into out_cols[fullrowptr[M]]; when col_counts != NULL
 *                        also col_counts[c] += 1 per entry = sp.linalg.norm(U, ord=0, axis=0) (sampler.py:117);
 *                        col_counts must be zero on entry.
 * gnn_member_set       : membership tables of after_nodes (ASCENDING and distinct: np.unique output, sampler.py:131) over
 *                        the node ids: bits[v >> 5] bit (v & 31) set for every member v, rank0[w] = position inside
 *                        after_nodes of the first member of word w (written only for words that hold a member);
 *                        set == 0 clears the words again.  bits: uint32 [ceil(N / 32)], all zero between uses;
 *                        rank0: int32 [ceil(N / 32)].  The local column id of node v is
 *                        rank0[v >> 5] + popcount(bits[v >> 5] & ((1 << (v & 31)) - 1)).
 * gnn_column_slice_count / _fill : adj = U[:, after_nodes] (sampler.py:133-136).  The rows of U lie one after the other
 *                        in ucols[total] and kept entries keep their order, so the slice is an order-preserving stream
 *                        compaction over the whole array, done in chunks of entries (balanced whatever the row lengths):
 *                        _count fills chunk_prefix (scratch, int32 [2 * gnn_column_slice_chunks(total) + 2]; its first
 *                        chunks + 1 entries = kept entries before each chunk, the last of them = nnz) and
 *                        out_rowptr[M+1] (kept entries before every row; out_rowptr[M] = nnz); _fill writes the kept
 *                        entries renumbered to positions inside after_nodes, ascending within a row, as int16
 *                        (reference hand-off) or int32, from the same ucols / membership tables / chunk_prefix.
 * gnn_support_compact  : the support of the column counts (ids v with counts[v] != 0, sampler.py:117 / :124) as
 *                        (nz_out[j] = v ascending, cnt_out[j] = counts[v]) and *n_support_out = their number; the three
 *                        outputs may be pinned host memory (device-accessible pointers): only the support crosses
 *                        PCIe.  chunk_scratch: int32 [2 * gnn_column_slice_chunks(num_nodes) + 2] on the device.
 * ------------------------------------------------------------------------- */
int gnn_row_slice_count(const int64_t *indptr, const int64_t *nodes, int64_t M, int32_t *scratch_lens,
                        int32_t *out_fullrowptr, gnn_stream_t stream);
int gnn_row_slice_fill(const int64_t *indptr, const int32_t *indices, const int64_t *nodes, int64_t M,
                       const int32_t *fullrowptr, int32_t *out_cols, int32_t *col_counts, gnn_stream_t stream);
int gnn_member_set(uint32_t *bits, int32_t *rank0, const int64_t *after_nodes, int64_t K, int set, gnn_stream_t stream);
int64_t gnn_column_slice_chunks(int64_t total);
int gnn_support_compact(const int32_t *counts, int64_t num_nodes, int32_t *chunk_scratch, int64_t *nz_out, int32_t *cnt_out,
                        int64_t *n_support_out, gnn_stream_t stream);
int gnn_column_slice_count(const int32_t *ucols, int64_t total, const int32_t *fullrowptr, int64_t M, const uint32_t *bits,
                           int32_t *chunk_prefix, int32_t *out_rowptr, gnn_stream_t stream);
int gnn_column_slice_fill(const int32_t *ucols, int64_t total, const uint32_t *bits, const int32_t *rank0, const int32_t *chunk_prefix,
                          void *out_colidx, int colidx_bytes, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Fused layer epilogue - the elementwise tail of every reference layer,
 * models.py:21-25 (GraphSageConvolution) and :61-64 (GraphConvolution):
 *     out = elu(x);  mean = out.mean(1);  var = out.var(1, unbiased=False) + 1e-9
 *     y = (out - mean) * scale * rsqrt(var) + offset
 * gnn_elu_rownorm_fwd_f32 also returns mean[M] and rstd[M] for the backward.
 * gnn_elu_rownorm_bwd_f32: dx[M,C], dscale[C], doffset[C] from dy; column sums are reduced in a fixed order
 * (per-CTA partials in the workspace, gnn_elu_rownorm_workspace_bytes(C) bytes), so results are reproducible.
 * C <= 2048.
 * ------------------------------------------------------------------------- */
size_t gnn_elu_rownorm_workspace_bytes(int64_t C);
int gnn_elu_rownorm_fwd_f32(const float *x, int64_t ldx, int64_t M, int64_t C, const float *scale, const float *offset,
                            float *y, int64_t ldy, float *mean, float *rstd, gnn_stream_t stream);
int gnn_elu_rownorm_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx, int64_t M, int64_t C,
                            const float *scale, const float *mean, const float *rstd, float *dx, int64_t lddx,
                            float *dscale, float *doffset, void *workspace, size_t workspace_bytes, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Dense linears of a layer on the tcgen05 tensor cores, fp32 in / fp32 out, 3xTF32 inside.
 *
 * Replace, for models.py:18-19 `cat[linearB(x[sampled_nodes]), linearW(feat)]` and models.py:60 `linear(feat)`, the
 * index kernel + two cuBLAS fp32 GEMMs + concat of the forward and the GEMMs of the autograd backward.  Every operand
 * element a is split into hi = rn_tf32(a), lo = rn_tf32(a - hi) and a.b is evaluated as lo.hi + hi.lo + hi.hi in one
 * fp32 accumulator: the dropped terms are < 2^-21 |a||b| per product, results agree with an fp64 product to ~1e-6.
 * Inf/NaN inputs produce NaN (Inf - Inf in the split) where an fp32 GEMM would keep Inf.
 *
 * gnn_linear_split_elems(rows, cols): floats of a split weight buffer = 2 * rows * ceil32(cols).
 * gnn_linear_split_weights_f32: W[N,K] (leading dimension ldw) -> w_nk = [2][N][ceil32(K)] (hi plane, lo plane, zero
 *     padded; operand of the forward) and, unless NULL, w_kn = [2][K][ceil32(N)] (W^T, operand of dX).  Both must be
 *     16-byte aligned.  Once per optimizer step.
 * gnn_linear_tf32x3_f32: C[m, 0:N] = A[a_rows ? a_rows[m] : m, 0:K] . W^T + bias   for m < M
 *     A: [*, lda] floats, any alignment (16-byte aligned rows take 128-bit loads); a_rows: int64[M] or NULL - the
 *     x[sampled_nodes] gather of models.py:19 costs nothing extra; w_split: w_nk of W[N,K]; bias: N floats or NULL;
 *     C: leading dimension ldc >= N, so a column slice of the concatenated layer output can be written in place.
 *     dX = dY . W is the same call with A = dY, K = N_out, N = K_in, w_split = w_kn, bias = NULL.
 *     K > 1024 is cut into equal chunks (one launch each, added in fp32): the tensor core truncates on every
 *     accumulation step, and 128 steps per chain keep that bias below 2.5e-6.
 * gnn_linear_tf32x3_f32_ex: the same product with
 *     flags & GNN_LINEAR_ACCUMULATE   C += result instead of C = result;
 *     c_rows (int64[M] or NULL)       result row m is ADDED to C[c_rows[m], :] with L2 reductions (requires the flag;
 *                                     duplicates are summed).  This is the backward of x[sampled_nodes]: the dX of
 *                                     linearB lands on top of the dX the SpMM backward wrote, no zero fill, no index_add.
 * gnn_linear_wgrad_tf32x3_f32: dW[n, k] = sum_m dY[m, n] * X[x_rows ? x_rows[m] : m, k]   (n < N, k < K)
 *     split over m across CTAs; partial sums are added in ascending order (fixed => bit-reproducible).
 *     dbias (N floats or NULL): the bias gradient sum_m dY[m, n], plain fp32 sums in a fixed order, from the dY values
 *     the operand loader holds anyway (no extra pass over dY).
 *     workspace: gnn_linear_wgrad_workspace_bytes(M, N, K) bytes.
 * ------------------------------------------------------------------------- */
size_t gnn_linear_split_elems(int64_t rows, int64_t cols);
int gnn_linear_split_weights_f32(const float *W, int64_t ldw, int64_t N, int64_t K, float *w_nk, float *w_kn,
                                 gnn_stream_t stream);
/* the two weight matrices of one GraphSAGE layer (linearB, linearW: models.py:11-12) split in ONE launch */
int gnn_linear_split_weights2_f32(const float *W0, int64_t ldw0, int64_t N0, int64_t K0, float *w_nk0, float *w_kn0, const float *W1,
                                  int64_t ldw1, int64_t N1, int64_t K1, float *w_nk1, float *w_kn1, gnn_stream_t stream);
int gnn_linear_tf32x3_f32(const float *A, int64_t lda, const int64_t *a_rows, int64_t M, int64_t K, const float *w_split,
                          int64_t N, const float *bias, float *C, int64_t ldc, gnn_stream_t stream);
#define GNN_LINEAR_ACCUMULATE 1u
int gnn_linear_tf32x3_f32_ex(const float *A, int64_t lda, const int64_t *a_rows, int64_t M, int64_t K, const float *w_split,
                             int64_t N, const float *bias, float *C, int64_t ldc, const int64_t *c_rows, unsigned flags,
                             gnn_stream_t stream);
size_t gnn_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K);
int gnn_linear_wgrad_tf32x3_f32(const float *dY, int64_t lddy, const float *X, int64_t ldx, const int64_t *x_rows, int64_t M,
                                int64_t N, int64_t K, float *dW, int64_t lddw, float *dbias, void *workspace,
                                size_t workspace_bytes, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_probe_row_gather_f32 - measurement aid (bench.py): the row gather of an SpMM and nothing else.
 *
 * Every warp walks a stretch of `colidx` (the block's own column stream) and loads the addressed rows of X
 * (float4 per lane, nv = 1, 2 or 4 vectors per lane, several rows in flight), adding them up; no values, no row
 * bookkeeping, no stores.  `warps_per_sm` resident warps per SM share the stream; *bytes_gathered (host pointer,
 * written before return) = bytes the launch requests from L2.  Timing it on the benchmark's own blocks gives the
 * L2->SM gather speed of light that bounds any row-wise fp32 SpMM on a dense LADIES block.
 * X: [K, ldx] floats, 16-byte aligned rows, D >= 128*nv; sink: one float (never written in practice).
 * ------------------------------------------------------------------------- */
int gnn_probe_row_gather_f32(const float *X, int64_t ldx, int64_t D, const int32_t *colidx, int64_t nnz, int nv,
                             int warps_per_sm, float *sink, int64_t *bytes_gathered, gnn_stream_t stream);

/* ---------------------------------------------------------------------------
 * gnn_legacy_choice_f64 - HOST helper of the device LADIES sampler (no CUDA calls, allocates host scratch).
 *
 * Replaces `np.random.choice(num_nodes, s_num, p=p, replace=False)` of sampler.py:128 (numpy legacy RandomState,
 * seeded by sampler.py:96) bit for bit: the same MT19937 stream, the same sequential cumsum / normalise / right-sided
 * search / first-occurrence filter per round as numpy's mtrand.pyx, so the sampled node set is unchanged.
 *   mt_state : uint32[625] = RandomState.get_state()[1] (624 key words) followed by get_state()[2] (position);
 *              advanced in place, so the draws of consecutive layers continue one stream like the reference's.
 *   p        : n probabilities (any entries may be zero; they can never be drawn), found: `size` indices out.
 * Returns GNN_E_BADARG when fewer than `size` entries of p are non-zero (numpy raises ValueError there).
 * ------------------------------------------------------------------------- */
int gnn_legacy_choice_f64(uint32_t *mt_state, const double *p, int64_t n, int64_t size, int64_t *found);

/* gnn_ladies_layer_host - the whole host part of one LADIES layer, sampler.py:117-143, in one call outside the GIL:
 *   pi = counts (scaled by scale_factor and truncated for the nodes in skew_nodes, :119-121);  p = pi / sum(pi)  (:124)
 *   s_num = min(#p > 0, samp_num) (:126);  draw = choice(p, s_num) (:128, gnn_legacy_choice_f64)
 *   after_nodes = unique(nz[draw] ++ previous_nodes) (:131);  normfact = 1 / float32(clip(s_num * p[after_nodes], 1e-10, 1)) (:137)
 *   sampled = positions of the distinct previous_nodes inside after_nodes (:143)
 * nz: the n_nz node ids with a non-zero count, ascending; counts: their int32 column counts (the device sampler brings
 * both back from the GPU); skew_nodes: ascending distinct ids or NULL.  after_nodes / normfact need room for
 * min(n_nz, samp_num) + n_prev entries, sampled for n_prev.  Returns the number of after_nodes (>= 0) or a negative code.
 * Integer arithmetic and the reference's own IEEE expressions only: identical to the numpy code bit for bit. */
int64_t gnn_ladies_layer_host(uint32_t *mt_state, const int64_t *nz, const int32_t *counts, int64_t n_nz,
                              const int64_t *skew_nodes, int64_t n_skew, double scale_factor, const int64_t *previous_nodes,
                              int64_t n_prev, int64_t samp_num, int64_t *after_nodes, float *normfact, int64_t *sampled,
                              int64_t *n_sampled);

/* gnn_ladies_layer_host_ex - gnn_ladies_layer_host with, optionally, the whole column-count array (counts_dense[v] for every
 * v < num_nodes, or NULL) beside the compacted support: p[after_nodes] is then read from it directly instead of through
 * an id -> support-position map.  Same outputs. */
int64_t gnn_ladies_layer_host_ex(uint32_t *mt_state, const int64_t *nz, const int32_t *counts, int64_t n_nz,
                                 const int32_t *counts_dense, int64_t num_nodes, const int64_t *skew_nodes, int64_t n_skew,
                                 double scale_factor, const int64_t *previous_nodes, int64_t n_prev, int64_t samp_num,
                                 int64_t *after_nodes, float *normfact, int64_t *sampled, int64_t *n_sampled);

/* gnn_ladies_layer_host_dense - the same call fed with the WHOLE column-count array of the layer (counts_dense[v] = how many
 * rows of U = lap_matrix[previous_nodes, :] hold column v, sampler.py:117 before any compaction; num_nodes entries, the
 * device sampler copies it into pinned memory in one transfer).  The support (ids with a non-zero count) is compacted
 * on the host in one pass; *n_support (optional) receives its size, which bounds s_num.  after_nodes / normfact need room
 * for min(num_nodes, samp_num) + n_prev entries.  Outputs are those of gnn_ladies_layer_host on the compacted arrays. */
int64_t gnn_ladies_layer_host_dense(uint32_t *mt_state, const int32_t *counts_dense, int64_t num_nodes, const int64_t *skew_nodes,
                                    int64_t n_skew, double scale_factor, const int64_t *previous_nodes, int64_t n_prev,
                                    int64_t samp_num, int64_t *after_nodes, float *normfact, int64_t *sampled, int64_t *n_sampled,
                                    int64_t *n_support);

/* ---------------------------------------------------------------------------
 * Feature-shard memory that peers can map (one process per GPU).
 *
 * gnn_shard_alloc   : cudaMalloc'd buffer + its 64-byte IPC handle.
 * gnn_shard_open    : map a peer process's shard into this process (NVLink P2P).
 * gnn_shard_close   : unmap.   gnn_shard_free : free the local buffer.
 * gnn_host_register : pin + map an existing host buffer; returns its device alias
 *                     (zero-copy reads over PCIe for uncached rows, main.py:134).
 * ------------------------------------------------------------------------- */
int gnn_shard_alloc(size_t bytes, void **dev_ptr, unsigned char ipc_handle[64]);
int gnn_shard_open(const unsigned char ipc_handle[64], void **dev_ptr);
int gnn_shard_close(void *dev_ptr);
int gnn_shard_free(void *dev_ptr);
int gnn_host_register(void *host_ptr, size_t bytes, void **dev_alias);
int gnn_host_unregister(void *host_ptr);

#ifdef __cplusplus
}
#endif
#endif /* GNN_B200_H_ */
