// spmm_ext.cpp - thin torch/pybind layer over the C ABI (include/gnn_b200.h).
//
// Keeps the reference extension's Python-visible signatures
// (reference spmm_cpp/spmm.cpp:52-56):
//     spmm_load_balance(Tensor sparseMat, Tensor denseMat) -> Tensor
//     spmm_naive(Tensor sparseMat, Tensor denseMat) -> Tensor
//     create_coo_tensor(Tensor fullrowptr, Tensor rowptr, Tensor colidx, Tensor normfact, int nrows, int ncols) -> Tensor
// and the same precondition checks (spmm.cpp:10-21: CUDA + coalesced sparse
// operand, CUDA + contiguous dense operands; colidx deliberately unchecked at
// spmm.cpp:47 - here it must be int16 or int32).  Differences by design: kernels
// run on PyTorch's current stream (the reference uses the legacy default stream),
// nothing synchronises the device, the GIL is released, and CUDA failures raise
// RuntimeError instead of exit(-1) (cuda_spmm.cu:16-24).
//
// Everything below only moves pointers: torch is plumbing (memory, streams).
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <cuda_runtime_api.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "gnn_b200.h"

namespace {

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_COAL(x) TORCH_CHECK((x).is_coalesced(), #x " must be coalesced")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CHECK_DENSE(x) \
  CHECK_CUDA(x);       \
  CHECK_CONTIGUOUS(x)
#define CHECK_SPARSE(x) \
  CHECK_CUDA(x);        \
  CHECK_COAL(x)

inline void check_rc(int rc, const char *what) {
  TORCH_CHECK(rc == 0, what, " failed: ", gnn_error_string(rc), " (code ", rc, ")");
}

inline gnn_stream_t cur_stream() { return (gnn_stream_t)c10::cuda::getCurrentCUDAStream().stream(); }

inline torch::Tensor workspace(size_t bytes, const torch::Device &dev) {
  return torch::empty({(int64_t)bytes}, torch::TensorOptions().dtype(torch::kUInt8).device(dev));
}

// ---- one long-lived SpMM workspace per (device, stream) ---------------------------------
// The SpMM kernels need arrival counters (zero on entry) and a partial-sum buffer.  The counters wrap back to zero inside
// the kernel, so a region that is zeroed once stays usable forever by calls on the same stream (which the GPU runs one
// after another): no memset and no allocator round trip per spmm() call - one kernel launch.  Streams under CUDA-graph
// capture get a fresh workspace instead (memory handed out during capture belongs to the graph's private pool).
struct StreamWorkspace {
  torch::Tensor counters, partials;
};
std::mutex g_ws_mutex;
std::map<std::pair<int, void *>, StreamWorkspace> g_ws;

struct SpmmWs {
  torch::Tensor counters, partials;
  unsigned flags;
};

SpmmWs spmm_workspace(int64_t M, int64_t nnz, int64_t D, const torch::Device &dev) {
  const size_t cb = gnn_csr_spmm_counter_bytes(M, nnz, D), pb = gnn_csr_spmm_partial_bytes(M, nnz, D);
  auto stream = c10::cuda::getCurrentCUDAStream();
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream.stream(), &cap) != cudaSuccess) { (void)cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  auto bytes = torch::TensorOptions().dtype(torch::kUInt8).device(dev);
  if (cap != cudaStreamCaptureStatusNone)
    return {torch::empty({(int64_t)cb}, bytes), torch::empty({(int64_t)pb}, bytes), 0u};
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  auto &ws = g_ws[{(int)dev.index(), (void *)stream.stream()}];
  if (!ws.counters.defined() || (size_t)ws.counters.numel() < cb)
    ws.counters = torch::zeros({(int64_t)std::max<size_t>(2 * cb, (size_t)1 << 20)}, bytes);      // zeroed on this stream, once
  if (!ws.partials.defined() || (size_t)ws.partials.numel() < pb)
    ws.partials = torch::empty({(int64_t)std::max<size_t>(pb + pb / 2, (size_t)8 << 20)}, bytes);
  return {ws.counters, ws.partials, GNN_SPMM_COUNTERS_ZEROED};
}

// ---- CSR fast path -------------------------------------------------------
inline const int32_t *rowidx_ptr(const c10::optional<torch::Tensor> &rowidx, int64_t nnz) {
  if (!rowidx.has_value() || !rowidx.value().defined()) return nullptr;
  const auto &t = rowidx.value();
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kInt && t.numel() == nnz,
              "rowidx must be a contiguous int32 CUDA tensor with one entry per nonzero");
  return t.data_ptr<int32_t>();
}

torch::Tensor csr_spmm(const torch::Tensor &rowptr, const torch::Tensor &colidx, const torch::Tensor &vals, int64_t M,
                       int64_t K, const torch::Tensor &dense, const c10::optional<torch::Tensor> &rowidx, bool padded_rows) {
  CHECK_DENSE(rowptr); CHECK_DENSE(colidx); CHECK_DENSE(vals); CHECK_CUDA(dense);
  // rows may be padded (stride(0) >= D) as long as each row is contiguous: the gathered buffer is 16-byte-row aligned
  TORCH_CHECK(dense.dim() == 2 && (dense.stride(1) == 1 || dense.size(1) <= 1) && dense.stride(0) >= dense.size(1),
              "denseMat must be contiguous");
  TORCH_CHECK(rowptr.scalar_type() == torch::kInt && colidx.scalar_type() == torch::kInt, "CSR indices must be int32");
  TORCH_CHECK(vals.scalar_type() == torch::kFloat && dense.scalar_type() == torch::kFloat, "values/dense must be float32");
  TORCH_CHECK(dense.dim() == 2 && dense.size(0) == K, "dense operand must be [", K, ", D], got ", dense.sizes());
  TORCH_CHECK(rowptr.numel() == M + 1, "rowptr must have M+1 entries");
  TORCH_CHECK(vals.device() == dense.device() && rowptr.device() == dense.device() && colidx.device() == dense.device(),
              "all operands must be on the same device");
  c10::cuda::CUDAGuard g(dense.device());
  const int64_t nnz = vals.numel(), D = dense.size(1);
  // padded_rows: rows of the result start on 128-byte lines (leading dimension ceil32(D)); the consumer (the tensor-core
  // linear that follows in a layer) then reads them with 128-bit loads even when D is 602
  const int64_t ldy = padded_rows ? (D + 31) / 32 * 32 : D;
  auto out = torch::empty({M, ldy}, dense.options());
  if (ldy != D) out = out.narrow(1, 0, D);
  auto ws = spmm_workspace(M, nnz, D, dense.device());
  check_rc(gnn_csr_spmm_f32_ex(rowptr.data_ptr<int32_t>(), rowidx_ptr(rowidx, nnz), colidx.data_ptr<int32_t>(),
                               vals.data_ptr<float>(), M, K, nnz, D, dense.data_ptr<float>(), dense.size(0) > 1 ? dense.stride(0) : D, out.data_ptr<float>(), ldy,
                               reinterpret_cast<int32_t *>(ws.counters.data_ptr()), ws.partials.data_ptr(),
                               (size_t)ws.partials.numel(), ws.flags, cur_stream()),
           "gnn_csr_spmm_f32");
  return out;
}

torch::Tensor gather_spmm(const torch::Tensor &rowptr, const torch::Tensor &colidx, const torch::Tensor &vals, int64_t M,
                          int64_t K, int64_t D, const torch::Tensor &xrows, const c10::optional<torch::Tensor> &rowidx) {
  CHECK_DENSE(rowptr); CHECK_DENSE(colidx); CHECK_DENSE(vals); CHECK_DENSE(xrows);
  TORCH_CHECK(xrows.scalar_type() == torch::kLong && xrows.numel() == K, "xrows must be an int64 pointer table of K entries");
  c10::cuda::CUDAGuard g(vals.device());
  const int64_t nnz = vals.numel();
  auto out = torch::empty({M, D}, vals.options());
  auto ws = spmm_workspace(M, nnz, D, vals.device());
  check_rc(gnn_gather_spmm_f32_ex(rowptr.data_ptr<int32_t>(), rowidx_ptr(rowidx, nnz), colidx.data_ptr<int32_t>(),
                                  vals.data_ptr<float>(), M, K, nnz, D,
                                  reinterpret_cast<const float *const *>(xrows.data_ptr<int64_t>()), out.data_ptr<float>(), D,
                                  reinterpret_cast<int32_t *>(ws.counters.data_ptr()), ws.partials.data_ptr(),
                                  (size_t)ws.partials.numel(), ws.flags, cur_stream()),
           "gnn_gather_spmm_f32");
  return out;
}

// dX = A^T . G from A's own CSR (no transposed index; vector reductions into a zero-filled output)
torch::Tensor csr_spmm_t(const torch::Tensor &rowptr, const torch::Tensor &colidx, const torch::Tensor &vals, int64_t M,
                         int64_t K, const torch::Tensor &grad, const c10::optional<torch::Tensor> &rowidx) {
  CHECK_DENSE(rowptr); CHECK_DENSE(colidx); CHECK_DENSE(vals); CHECK_CUDA(grad);
  TORCH_CHECK(grad.dim() == 2 && (grad.stride(1) == 1 || grad.size(1) <= 1) && grad.stride(0) >= grad.size(1),
              "grad_output must be contiguous");
  TORCH_CHECK(rowptr.scalar_type() == torch::kInt && colidx.scalar_type() == torch::kInt, "CSR indices must be int32");
  TORCH_CHECK(vals.scalar_type() == torch::kFloat && grad.scalar_type() == torch::kFloat, "values/grad must be float32");
  TORCH_CHECK(grad.size(0) == M, "grad_output must be [", M, ", D], got ", grad.sizes());
  c10::cuda::CUDAGuard g(grad.device());
  const int64_t nnz = vals.numel(), D = grad.size(1);
  auto out = torch::empty({K, D}, grad.options());
  check_rc(gnn_csr_spmm_t_f32(rowptr.data_ptr<int32_t>(), rowidx_ptr(rowidx, nnz), colidx.data_ptr<int32_t>(),
                              vals.data_ptr<float>(), M, K, nnz, D, grad.data_ptr<float>(), grad.size(0) > 1 ? grad.stride(0) : D, out.data_ptr<float>(), D, cur_stream()),
           "gnn_csr_spmm_t_f32");
  return out;
}

// L2->SM row-gather speed of light on a block's own column stream (measurement aid of bench.py): returns bytes gathered
int64_t probe_row_gather(const torch::Tensor &X, const torch::Tensor &colidx, int64_t nv, int64_t warps_per_sm) {
  CHECK_CUDA(X); CHECK_DENSE(colidx);
  TORCH_CHECK(X.dim() == 2 && X.stride(1) == 1 && X.scalar_type() == torch::kFloat && colidx.scalar_type() == torch::kInt,
              "X must be row-major float32, colidx int32");
  c10::cuda::CUDAGuard g(X.device());
  auto sink = torch::zeros({1}, X.options());
  int64_t bytes = 0;
  check_rc(gnn_probe_row_gather_f32(X.data_ptr<float>(), X.stride(0), X.size(1), colidx.data_ptr<int32_t>(), colidx.numel(),
                                    (int)nv, (int)warps_per_sm, sink.data_ptr<float>(), &bytes, cur_stream()),
           "gnn_probe_row_gather_f32");
  return bytes;
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, c10::optional<torch::Tensor>> csr_transpose(
    const torch::Tensor &rowptr, const torch::Tensor &colidx, const torch::Tensor &vals, int64_t M, int64_t K) {
  CHECK_DENSE(rowptr); CHECK_DENSE(colidx); CHECK_DENSE(vals);
  c10::cuda::CUDAGuard g(vals.device());
  const int64_t nnz = vals.numel();
  auto iopt = rowptr.options();
  auto t_rowptr = torch::empty({K + 1}, iopt);
  auto t_colidx = torch::empty({nnz}, iopt);
  auto t_vals = torch::empty({nnz}, vals.options());
  // row ids of A^T's entries only where a short-row kernel will read them (one more scattered store per entry otherwise)
  const bool want_rows = nnz < 96 * std::max<int64_t>(K, 1);
  auto t_rowidx = want_rows ? torch::empty({nnz}, iopt) : torch::Tensor();
  const size_t wsb = gnn_csr_transpose_workspace_bytes(M, K, nnz);      // bounded by the bitmap budget (row-blocked beyond it)
  auto ws = workspace(wsb, vals.device());
  check_rc(gnn_csr_transpose(rowptr.data_ptr<int32_t>(), colidx.data_ptr<int32_t>(), vals.data_ptr<float>(), M, K, nnz,
                             t_rowptr.data_ptr<int32_t>(), t_colidx.data_ptr<int32_t>(), t_vals.data_ptr<float>(),
                             want_rows ? t_rowidx.data_ptr<int32_t>() : nullptr, ws.data_ptr(), wsb, cur_stream()),
           "gnn_csr_transpose");
  return {t_rowptr, t_colidx, t_vals, want_rows ? c10::optional<torch::Tensor>(t_rowidx) : c10::nullopt};
}

std::tuple<torch::Tensor, torch::Tensor> coo_to_csr(const torch::Tensor &sparseMat) {
  CHECK_SPARSE(sparseMat);
  TORCH_CHECK(sparseMat.dim() == 2, "sparse operand must be 2-D");
  c10::cuda::CUDAGuard g(sparseMat.device());
  auto indices = sparseMat._indices().contiguous();
  const int64_t M = sparseMat.size(0), nnz = indices.size(1);
  auto iopt = indices.options().dtype(torch::kInt);
  auto rowptr = torch::empty({M + 1}, iopt);
  auto col32 = torch::empty({nnz}, iopt);
  check_rc(gnn_coo_to_csr(indices.data_ptr<int64_t>(), M, nnz, rowptr.data_ptr<int32_t>(), col32.data_ptr<int32_t>(), cur_stream()),
           "gnn_coo_to_csr");
  return {rowptr, col32};
}

// ---- reference-named entry points ---------------------------------------
torch::Tensor spmm_load_balance(const torch::Tensor &sparseMat, const torch::Tensor &denseMat) {
  CHECK_SPARSE(sparseMat);
  CHECK_DENSE(denseMat);
  auto csr = coo_to_csr(sparseMat);
  auto vals = sparseMat._values().contiguous();
  return csr_spmm(std::get<0>(csr), std::get<1>(csr), vals, sparseMat.size(0), sparseMat.size(1), denseMat, c10::nullopt, false);
}

torch::Tensor spmm_naive(const torch::Tensor &sparseMat, const torch::Tensor &denseMat) {
  // the reference's v1 differs from v2 only in scheduling; one deterministic kernel serves both names
  return spmm_load_balance(sparseMat, denseMat);
}

std::tuple<torch::Tensor, torch::Tensor, c10::optional<torch::Tensor>> build_adj(const torch::Tensor &fullrowptr,
                                                                                 const torch::Tensor &rowptr,
                                                                                 const torch::Tensor &colidx,
                                                                                 const torch::Tensor &normfact, int64_t nrows,
                                                                                 int64_t ncols) {
  CHECK_DENSE(fullrowptr);
  CHECK_DENSE(rowptr);
  CHECK_DENSE(colidx);
  CHECK_DENSE(normfact);
  TORCH_CHECK(fullrowptr.scalar_type() == torch::kInt && rowptr.scalar_type() == torch::kInt, "row pointers must be int32");
  TORCH_CHECK(colidx.scalar_type() == torch::kShort || colidx.scalar_type() == torch::kInt, "colidx must be int16 or int32");
  TORCH_CHECK(normfact.scalar_type() == torch::kFloat, "normfact must be float32");
  TORCH_CHECK(rowptr.numel() == nrows + 1 && fullrowptr.numel() == nrows + 1, "row pointers must have nrows+1 entries");
  TORCH_CHECK(normfact.numel() >= ncols, "normfact must have ncols entries");
  TORCH_CHECK(colidx.scalar_type() != torch::kShort || ncols <= 32768, "int16 colidx cannot address ", ncols, " columns");
  c10::cuda::CUDAGuard g(colidx.device());
  const int64_t nnz = colidx.numel();
  auto indices = torch::empty({2, nnz}, colidx.options().dtype(torch::kLong));
  auto values = torch::empty({nnz}, normfact.options());
  auto col32 = torch::empty({nnz}, colidx.options().dtype(torch::kInt));
  // per-entry row ids for the short-row kernels (flat SpMM, scatter backward); dense LADIES blocks never read them
  const bool want_rows = nnz < 96 * std::max<int64_t>(nrows, 1);
  auto row32 = want_rows ? torch::empty({nnz}, colidx.options().dtype(torch::kInt)) : torch::Tensor();
  check_rc(gnn_build_adj(fullrowptr.data_ptr<int32_t>(), rowptr.data_ptr<int32_t>(), colidx.data_ptr(),
                         colidx.scalar_type() == torch::kShort ? 2 : 4, normfact.data_ptr<float>(), nrows, ncols, nnz,
                         indices.data_ptr<int64_t>(), values.data_ptr<float>(), col32.data_ptr<int32_t>(),
                         want_rows ? row32.data_ptr<int32_t>() : nullptr, cur_stream()),
           "gnn_build_adj");
  // rows ascending, columns ascending and unique inside a row: already coalesced (no sort, cuda_spmm.cu:825)
  auto coo = at::_sparse_coo_tensor_unsafe(indices, values, {nrows, ncols}, values.options().layout(torch::kSparse),
                                           /*is_coalesced=*/true);
  return {coo, col32, want_rows ? c10::optional<torch::Tensor>(row32) : c10::nullopt};
}

torch::Tensor create_coo_tensor(const torch::Tensor &fullrowptr, const torch::Tensor &rowptr, const torch::Tensor &colidx,
                                const torch::Tensor &normfact, int64_t nrows, int64_t ncols) {
  return std::get<0>(build_adj(fullrowptr, rowptr, colidx, normfact, nrows, ncols));
}

// ---- gather path ----------------------------------------------------------
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> placement_remap(
    const torch::Tensor &input_nodes, const torch::Tensor &device_id_of_nodes, const torch::Tensor &idx_of_nodes_on_device,
    const torch::Tensor &devices, const torch::Tensor &bases, int64_t ld_src, int64_t ld_host) {
  CHECK_DENSE(input_nodes); CHECK_DENSE(device_id_of_nodes); CHECK_DENSE(idx_of_nodes_on_device); CHECK_DENSE(devices);
  CHECK_DENSE(bases);
  TORCH_CHECK(input_nodes.scalar_type() == torch::kLong && device_id_of_nodes.scalar_type() == torch::kLong &&
                  idx_of_nodes_on_device.scalar_type() == torch::kLong && devices.scalar_type() == torch::kLong &&
                  bases.scalar_type() == torch::kLong,
              "placement tables must be int64");
  const int64_t world = devices.numel(), n0 = input_nodes.numel();
  TORCH_CHECK(bases.numel() == world + 1, "bases must hold world+1 pointers (last = host table)");
  c10::cuda::CUDAGuard g(input_nodes.device());
  auto src_dev = torch::empty({n0}, input_nodes.options().dtype(torch::kInt));
  auto slot = torch::empty({n0}, input_nodes.options());
  auto xrows = torch::empty({n0}, input_nodes.options());
  auto counts = torch::empty({world + 2}, input_nodes.options());
  check_rc(gnn_placement_remap(input_nodes.data_ptr<int64_t>(), n0, device_id_of_nodes.data_ptr<int64_t>(),
                               idx_of_nodes_on_device.data_ptr<int64_t>(), devices.data_ptr<int64_t>(), world,
                               reinterpret_cast<const float *const *>(bases.data_ptr<int64_t>()), ld_src, ld_host,
                               src_dev.data_ptr<int32_t>(), slot.data_ptr<int64_t>(),
                               reinterpret_cast<const float **>(xrows.data_ptr<int64_t>()), counts.data_ptr<int64_t>(),
                               cur_stream()),
           "gnn_placement_remap");
  return {src_dev, slot, xrows, counts};
}

torch::Tensor gather_rows(const torch::Tensor &xrows, int64_t F, int64_t ld_out) {
  CHECK_DENSE(xrows);
  TORCH_CHECK(xrows.scalar_type() == torch::kLong, "xrows must be an int64 pointer table");
  TORCH_CHECK(ld_out >= F, "ld_out must be >= F");
  c10::cuda::CUDAGuard g(xrows.device());
  const int64_t n0 = xrows.numel();
  auto buf = torch::empty({n0, ld_out}, xrows.options().dtype(torch::kFloat));
  check_rc(gnn_gather_rows_f32(reinterpret_cast<const float *const *>(xrows.data_ptr<int64_t>()), n0, F, buf.data_ptr<float>(),
                               ld_out, cur_stream()),
           "gnn_gather_rows_f32");
  return ld_out == F ? buf : buf.narrow(1, 0, F);
}

void gather_rows_src(const torch::Tensor &xrows, const torch::Tensor &src_dev, int64_t only_src, int64_t F, torch::Tensor out) {
  CHECK_DENSE(xrows); CHECK_DENSE(src_dev); CHECK_CUDA(out);
  TORCH_CHECK(out.dim() == 2 && out.stride(1) == 1 && out.size(1) >= F && out.size(0) == xrows.numel(), "bad output buffer");
  c10::cuda::CUDAGuard g(xrows.device());
  check_rc(gnn_gather_rows_src_f32(reinterpret_cast<const float *const *>(xrows.data_ptr<int64_t>()), src_dev.data_ptr<int32_t>(),
                                   (int32_t)only_src, xrows.numel(), F, out.data_ptr<float>(), out.stride(0), cur_stream()),
           "gnn_gather_rows_src_f32");
}

torch::Tensor index_rows(const torch::Tensor &X, const torch::Tensor &idx) {
  CHECK_CUDA(X); CHECK_DENSE(idx);
  TORCH_CHECK(X.dim() == 2 && X.stride(1) == 1 && X.scalar_type() == torch::kFloat, "X must be a row-major float32 matrix");
  TORCH_CHECK(idx.scalar_type() == torch::kLong, "idx must be int64");
  c10::cuda::CUDAGuard g(X.device());
  auto out = torch::empty({idx.numel(), X.size(1)}, X.options());
  check_rc(gnn_index_rows_f32(X.data_ptr<float>(), X.stride(0), idx.data_ptr<int64_t>(), idx.numel(), X.size(1),
                              out.data_ptr<float>(), X.size(1), cur_stream()),
           "gnn_index_rows_f32");
  return out;
}

// ---- LADIES layer construction on the device (array work of sampler.py:113-137) ----------
torch::Tensor row_slice_count(const torch::Tensor &indptr, const torch::Tensor &nodes) {
  CHECK_DENSE(indptr); CHECK_DENSE(nodes);
  TORCH_CHECK(indptr.scalar_type() == torch::kLong && nodes.scalar_type() == torch::kLong, "indptr / nodes must be int64");
  c10::cuda::CUDAGuard g(nodes.device());
  const int64_t M = nodes.numel();
  auto iopt = nodes.options().dtype(torch::kInt);
  auto lens = torch::empty({M}, iopt);
  auto fullrowptr = torch::empty({M + 1}, iopt);
  check_rc(gnn_row_slice_count(indptr.data_ptr<int64_t>(), nodes.data_ptr<int64_t>(), M, lens.data_ptr<int32_t>(),
                               fullrowptr.data_ptr<int32_t>(), cur_stream()),
           "gnn_row_slice_count");
  return fullrowptr;
}

torch::Tensor row_slice_fill(const torch::Tensor &indptr, const torch::Tensor &indices, const torch::Tensor &nodes,
                             const torch::Tensor &fullrowptr, int64_t total, c10::optional<torch::Tensor> col_counts) {
  CHECK_DENSE(indptr); CHECK_DENSE(indices); CHECK_DENSE(nodes); CHECK_DENSE(fullrowptr);
  TORCH_CHECK(indices.scalar_type() == torch::kInt && fullrowptr.scalar_type() == torch::kInt, "indices / fullrowptr must be int32");
  if (col_counts.has_value()) {
    CHECK_DENSE(col_counts.value());
    TORCH_CHECK(col_counts.value().scalar_type() == torch::kInt, "col_counts must be int32");
  }
  c10::cuda::CUDAGuard g(nodes.device());
  auto ucols = torch::empty({total}, fullrowptr.options());
  check_rc(gnn_row_slice_fill(indptr.data_ptr<int64_t>(), indices.data_ptr<int32_t>(), nodes.data_ptr<int64_t>(), nodes.numel(),
                              fullrowptr.data_ptr<int32_t>(), ucols.data_ptr<int32_t>(),
                              col_counts.has_value() ? col_counts.value().data_ptr<int32_t>() : nullptr, cur_stream()),
           "gnn_row_slice_fill");
  return ucols;
}

void member_set(torch::Tensor bits, torch::Tensor rank0, const torch::Tensor &after_nodes, bool set) {
  CHECK_DENSE(bits); CHECK_DENSE(rank0); CHECK_DENSE(after_nodes);
  TORCH_CHECK(bits.scalar_type() == torch::kInt && rank0.scalar_type() == torch::kInt && bits.numel() == rank0.numel() &&
              after_nodes.scalar_type() == torch::kLong, "bits / rank0 int32 [ceil(N/32)], after_nodes int64");
  c10::cuda::CUDAGuard g(bits.device());
  check_rc(gnn_member_set(reinterpret_cast<uint32_t *>(bits.data_ptr<int32_t>()), rank0.data_ptr<int32_t>(),
                          after_nodes.data_ptr<int64_t>(), after_nodes.numel(), set ? 1 : 0, cur_stream()),
           "gnn_member_set");
}

// -> (rowptr int32 [M+1], chunk_prefix scratch for column_slice_fill)
std::tuple<torch::Tensor, torch::Tensor> column_slice_count(const torch::Tensor &ucols, const torch::Tensor &fullrowptr,
                                                            const torch::Tensor &bits) {
  CHECK_DENSE(ucols); CHECK_DENSE(fullrowptr); CHECK_DENSE(bits);
  TORCH_CHECK(ucols.scalar_type() == torch::kInt && fullrowptr.scalar_type() == torch::kInt && bits.scalar_type() == torch::kInt,
              "ucols / fullrowptr / bits must be int32");
  c10::cuda::CUDAGuard g(ucols.device());
  const int64_t M = fullrowptr.numel() - 1, total = ucols.numel();
  auto chunk_prefix = torch::empty({2 * gnn_column_slice_chunks(total) + 2}, fullrowptr.options());
  auto rowptr = torch::empty({M + 1}, fullrowptr.options());
  check_rc(gnn_column_slice_count(ucols.data_ptr<int32_t>(), total, fullrowptr.data_ptr<int32_t>(), M,
                                  reinterpret_cast<const uint32_t *>(bits.data_ptr<int32_t>()), chunk_prefix.data_ptr<int32_t>(),
                                  rowptr.data_ptr<int32_t>(), cur_stream()),
           "gnn_column_slice_count");
  return {rowptr, chunk_prefix};
}

torch::Tensor column_slice_fill(const torch::Tensor &ucols, const torch::Tensor &bits, const torch::Tensor &rank0,
                                const torch::Tensor &chunk_prefix, int64_t nnz, bool int16_ids) {
  CHECK_DENSE(ucols); CHECK_DENSE(bits); CHECK_DENSE(rank0); CHECK_DENSE(chunk_prefix);
  TORCH_CHECK(bits.scalar_type() == torch::kInt && rank0.scalar_type() == torch::kInt && bits.numel() == rank0.numel(),
              "bits / rank0 must be int32 tables of the same size");
  TORCH_CHECK(chunk_prefix.numel() == 2 * gnn_column_slice_chunks(ucols.numel()) + 2, "chunk_prefix does not belong to ucols");
  c10::cuda::CUDAGuard g(ucols.device());
  auto colidx = torch::empty({nnz}, ucols.options().dtype(int16_ids ? torch::kShort : torch::kInt));
  check_rc(gnn_column_slice_fill(ucols.data_ptr<int32_t>(), ucols.numel(), reinterpret_cast<const uint32_t *>(bits.data_ptr<int32_t>()),
                                 rank0.data_ptr<int32_t>(), chunk_prefix.data_ptr<int32_t>(), colidx.data_ptr(), int16_ids ? 2 : 4,
                                 cur_stream()),
           "gnn_column_slice_fill");
  return colidx;
}

// One whole LADIES layer (reference sampler.py:113-143) in ONE call with the GIL released: the device passes above, the
// column counts into pinned memory, the host part (gnn_ladies_layer_host_dense: probabilities, legacy weighted draw,
// union, normfact, remap), the uploads through pinned blocks of torch's caching host allocator, the column slice.
// The sampler threads of a rank share one interpreter with the training thread; as separate Python-level calls the
// glue of a layer held the GIL for ~0.4 ms, which - not the host cores - bounded live-sampler training.
// Two stream synchronisations (counts, kept count).  Returns (fullrowptr, rowptr, colidx, normfact [device],
// after_nodes int64, sampled positions int64 [host]).
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> ladies_layer_device(
    const torch::Tensor &indptr, const torch::Tensor &indices, const torch::Tensor &indptr_host, torch::Tensor bits,
    torch::Tensor rank0, torch::Tensor counts, c10::optional<torch::Tensor> counts_host_opt, torch::Tensor mt_state, const torch::Tensor &previous_nodes,
    c10::optional<torch::Tensor> skew_nodes, double scale_factor, int64_t samp_num, bool int16_ids, int64_t device_compact_min_nodes) {
  CHECK_DENSE(indptr); CHECK_DENSE(indices); CHECK_DENSE(bits); CHECK_DENSE(rank0); CHECK_DENSE(counts);
  TORCH_CHECK(indptr.scalar_type() == torch::kLong && indices.scalar_type() == torch::kInt, "indptr int64, indices int32");
  TORCH_CHECK(bits.scalar_type() == torch::kInt && rank0.scalar_type() == torch::kInt && counts.scalar_type() == torch::kInt,
              "bits / rank0 / counts must be int32");
  TORCH_CHECK(bits.numel() == (counts.numel() + 31) / 32 && rank0.numel() == bits.numel(), "bits / rank0: ceil(N / 32) words");
  TORCH_CHECK(!indptr_host.is_cuda() && indptr_host.scalar_type() == torch::kLong && indptr_host.is_contiguous() &&
              indptr_host.numel() == indptr.numel(), "indptr_host: the host copy of indptr (int64)");
  const bool have_dense = counts_host_opt.has_value() && counts_host_opt.value().defined();
  const bool device_compact = counts.numel() >= device_compact_min_nodes;
  TORCH_CHECK(have_dense || device_compact, "without a pinned mirror of the counts the support must be compacted on the device");
  torch::Tensor counts_host;
  if (have_dense) {
    counts_host = counts_host_opt.value();
    TORCH_CHECK(!counts_host.is_cuda() && counts_host.is_pinned() && counts_host.scalar_type() == torch::kInt &&
                counts_host.is_contiguous() && counts_host.numel() == counts.numel(), "counts_host: pinned int32 mirror of counts");
  }
  TORCH_CHECK(!mt_state.is_cuda() && mt_state.is_contiguous() && mt_state.numel() * mt_state.element_size() == 625 * 4,
              "mt_state: 625 32-bit words on the host");
  TORCH_CHECK(!previous_nodes.is_cuda() && previous_nodes.scalar_type() == torch::kLong && previous_nodes.is_contiguous(),
              "previous_nodes: int64 on the host");
  const int64_t *skew = nullptr;
  int64_t n_skew = 0;
  if (skew_nodes.has_value() && scale_factor > 1.0) {
    TORCH_CHECK(!skew_nodes.value().is_cuda() && skew_nodes.value().scalar_type() == torch::kLong && skew_nodes.value().is_contiguous(),
                "skew_nodes: int64 on the host");
    skew = skew_nodes.value().data_ptr<int64_t>();
    n_skew = skew_nodes.value().numel();
  }
  c10::cuda::CUDAGuard g(indptr.device());
  const auto dev = indptr.device();
  const int64_t N = counts.numel(), M = previous_nodes.numel();
  auto pinned = [](torch::ScalarType t) { return torch::TensorOptions().dtype(t).pinned_memory(true); };
  // U = lap_matrix[previous_nodes, :]  (:113-114) and its column counts (:117)
  const int64_t *prev = previous_nodes.data_ptr<int64_t>();
  const int64_t *ip = indptr_host.data_ptr<int64_t>();
  auto prev_pin = torch::empty({M}, pinned(torch::kLong));
  int64_t total = 0;
  for (int64_t i = 0; i < M; ++i) {
    TORCH_CHECK(prev[i] >= 0 && prev[i] < N, "previous_nodes out of range");
    total += ip[prev[i] + 1] - ip[prev[i]];
    prev_pin.data_ptr<int64_t>()[i] = prev[i];
  }
  auto prev_dev = prev_pin.to(dev, /*non_blocking=*/true);
  auto fullrowptr = row_slice_count(indptr, prev_dev);
  counts.zero_();
  auto ucols = row_slice_fill(indptr, indices, prev_dev, fullrowptr, total, counts);
  // the column counts come to the host: small graphs as the whole array (the native call compacts it in one pass); larger
  // ones compacted by the device, (id, count) pairs written straight into pinned memory - only the support crosses PCIe
  // and the host scans nothing - plus the whole array while it is small enough to be worth it (p[after_nodes] lookups)
  torch::Tensor nz_pin, cnt_pin, nsup_pin, chunk_scratch;
  if (device_compact) {
    const int64_t sup_cap = std::max<int64_t>(std::min<int64_t>(N, total), 1);
    nz_pin = torch::empty({sup_cap}, pinned(torch::kLong));
    cnt_pin = torch::empty({sup_cap}, pinned(torch::kInt));
    nsup_pin = torch::empty({1}, pinned(torch::kLong));
    chunk_scratch = torch::empty({2 * gnn_column_slice_chunks(N) + 2}, counts.options());
    void *d_nz = nullptr, *d_cnt = nullptr, *d_ns = nullptr;            // device-side addresses of the pinned blocks
    TORCH_CHECK(cudaHostGetDevicePointer(&d_nz, nz_pin.data_ptr(), 0) == cudaSuccess &&
                cudaHostGetDevicePointer(&d_cnt, cnt_pin.data_ptr(), 0) == cudaSuccess &&
                cudaHostGetDevicePointer(&d_ns, nsup_pin.data_ptr(), 0) == cudaSuccess, "pinned memory is not device-accessible");
    check_rc(gnn_support_compact(counts.data_ptr<int32_t>(), N, chunk_scratch.data_ptr<int32_t>(), (int64_t *)d_nz, (int32_t *)d_cnt,
                                 (int64_t *)d_ns, cur_stream()),
             "gnn_support_compact");
  }
  if (have_dense) counts_host.copy_(counts, /*non_blocking=*/true);
  TORCH_CHECK(cudaStreamSynchronize((cudaStream_t)cur_stream()) == cudaSuccess, "stream synchronise failed");
  // :117-143 on the host
  const int64_t cap = std::min<int64_t>(N, samp_num) + M;
  auto after_pin = torch::empty({cap}, pinned(torch::kLong));
  auto nf_pin = torch::empty({cap}, pinned(torch::kFloat));
  auto sampled = torch::empty({M}, torch::TensorOptions().dtype(torch::kLong));
  int64_t n_sampled = 0, n_support = 0, n_after = 0;
  if (device_compact) {
    n_support = nsup_pin.data_ptr<int64_t>()[0];
    TORCH_CHECK(n_support > 0 && n_support <= nz_pin.numel(), "device support compaction returned ", n_support, " entries");
    n_after = gnn_ladies_layer_host_ex(reinterpret_cast<uint32_t *>(mt_state.data_ptr()), nz_pin.data_ptr<int64_t>(),
                                       cnt_pin.data_ptr<int32_t>(), n_support, have_dense ? counts_host.data_ptr<int32_t>() : nullptr, N,
                                       skew, n_skew, scale_factor, prev, M, samp_num, after_pin.data_ptr<int64_t>(),
                                       nf_pin.data_ptr<float>(), sampled.data_ptr<int64_t>(), &n_sampled);
  } else {
    n_after = gnn_ladies_layer_host_dense(reinterpret_cast<uint32_t *>(mt_state.data_ptr()), counts_host.data_ptr<int32_t>(), N,
                                          skew, n_skew, scale_factor, prev, M, samp_num, after_pin.data_ptr<int64_t>(),
                                          nf_pin.data_ptr<float>(), sampled.data_ptr<int64_t>(), &n_sampled, &n_support);
  }
  if (n_after < 0) check_rc((int)n_after, "gnn_ladies_layer_host");
  auto after_host = after_pin.narrow(0, 0, n_after);
  auto after_dev = after_host.to(dev, /*non_blocking=*/true);
  auto nf_dev = nf_pin.narrow(0, 0, n_after).to(dev, /*non_blocking=*/true);
  // adj = U[:, after_nodes]  (:133-136)
  member_set(bits, rank0, after_dev, true);
  torch::Tensor rowptr, colidx;
  try {
    torch::Tensor chunk_prefix;
    std::tie(rowptr, chunk_prefix) = column_slice_count(ucols, fullrowptr, bits);
    const int64_t nnz = rowptr[M].item<int32_t>();
    colidx = column_slice_fill(ucols, bits, rank0, chunk_prefix, nnz, int16_ids && n_after <= 32768);
  } catch (...) {
    member_set(bits, rank0, after_dev, false);      // the bitmap must be all zero for the next minibatch, whatever happened
    throw;
  }
  member_set(bits, rank0, after_dev, false);
  return {fullrowptr, rowptr, colidx, nf_dev, after_host, sampled.narrow(0, 0, n_sampled)};
}

// ---- fused layer epilogue (models.py:21-25 / :61-64) ---------------------------------------
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> elu_rownorm_fwd(const torch::Tensor &x, const torch::Tensor &scale,
                                                                         const torch::Tensor &offset) {
  CHECK_CUDA(x); CHECK_DENSE(scale); CHECK_DENSE(offset);
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.scalar_type() == torch::kFloat, "x must be a row-major float32 matrix");
  TORCH_CHECK(scale.numel() == x.size(1) && offset.numel() == x.size(1), "scale/offset must have one entry per column");
  c10::cuda::CUDAGuard g(x.device());
  const int64_t M = x.size(0), C = x.size(1);
  auto y = torch::empty({M, C}, x.options());
  auto mean = torch::empty({M}, x.options());
  auto rstd = torch::empty({M}, x.options());
  check_rc(gnn_elu_rownorm_fwd_f32(x.data_ptr<float>(), M > 1 ? x.stride(0) : C, M, C, scale.data_ptr<float>(),
                                   offset.data_ptr<float>(), y.data_ptr<float>(), C, mean.data_ptr<float>(), rstd.data_ptr<float>(),
                                   cur_stream()),
           "gnn_elu_rownorm_fwd_f32");
  return {y, mean, rstd};
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> elu_rownorm_bwd(const torch::Tensor &dy, const torch::Tensor &x,
                                                                         const torch::Tensor &scale, const torch::Tensor &mean,
                                                                         const torch::Tensor &rstd) {
  CHECK_CUDA(dy); CHECK_CUDA(x); CHECK_DENSE(scale); CHECK_DENSE(mean); CHECK_DENSE(rstd);
  TORCH_CHECK(dy.dim() == 2 && dy.stride(1) == 1 && x.stride(1) == 1 && dy.sizes() == x.sizes(), "dy and x must be row-major and equal-shaped");
  c10::cuda::CUDAGuard g(x.device());
  const int64_t M = x.size(0), C = x.size(1);
  auto dx = torch::empty({M, C}, x.options());
  auto dscale = torch::empty({C}, x.options());
  auto doffset = torch::empty({C}, x.options());
  const size_t wsb = gnn_elu_rownorm_workspace_bytes(C);
  auto ws = workspace(wsb, x.device());
  check_rc(gnn_elu_rownorm_bwd_f32(dy.data_ptr<float>(), M > 1 ? dy.stride(0) : C, x.data_ptr<float>(), M > 1 ? x.stride(0) : C, M, C,
                                   scale.data_ptr<float>(), mean.data_ptr<float>(), rstd.data_ptr<float>(), dx.data_ptr<float>(), C,
                                   dscale.data_ptr<float>(), doffset.data_ptr<float>(), ws.data_ptr(), wsb, cur_stream()),
           "gnn_elu_rownorm_bwd_f32");
  return {dx, dscale, doffset};
}

// ---- dense linears on tcgen05 (3xTF32) -----------------------------------
inline const int64_t *rows_ptr(const c10::optional<torch::Tensor> &rows, int64_t M, const torch::Device &dev) {
  if (!rows.has_value() || !rows.value().defined()) return nullptr;
  const auto &t = rows.value();
  TORCH_CHECK(t.is_cuda() && t.device() == dev && t.is_contiguous() && t.scalar_type() == torch::kLong && t.numel() == M,
              "row index must be a contiguous int64 CUDA tensor with one entry per output row");
  return t.data_ptr<int64_t>();
}
inline void check_rowmajor(const torch::Tensor &t, const char *name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat && t.dim() == 2 && (t.stride(1) == 1 || t.size(1) <= 1) &&
                  (t.size(0) <= 1 || t.stride(0) >= t.size(1)),
              name, " must be a row-major float32 CUDA matrix (rows may be padded)");
}
inline int64_t ld_of(const torch::Tensor &t) { return t.size(0) > 1 ? t.stride(0) : t.size(1); }

// W[N,K] -> (w_nk [2,N,ceil32(K)], w_kn [2,K,ceil32(N)]): TF32 hi/lo planes for the forward and for dX
std::tuple<torch::Tensor, torch::Tensor> linear_split_weights(const torch::Tensor &W, bool with_transposed) {
  check_rowmajor(W, "W");
  c10::cuda::CUDAGuard g(W.device());
  const int64_t N = W.size(0), K = W.size(1);
  auto w_nk = torch::empty({2, N, (K + 31) / 32 * 32}, W.options());
  auto w_kn = with_transposed ? torch::empty({2, K, (N + 31) / 32 * 32}, W.options()) : torch::empty({0}, W.options());
  check_rc(gnn_linear_split_weights_f32(W.data_ptr<float>(), ld_of(W), N, K, w_nk.data_ptr<float>(),
                                        with_transposed ? w_kn.data_ptr<float>() : nullptr, cur_stream()),
           "gnn_linear_split_weights_f32");
  return {w_nk, w_kn};
}

// both weight matrices of a GraphSAGE layer in one launch -> (w_nk0, w_kn0, w_nk1, w_kn1)
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> linear_split_weights2(const torch::Tensor &W0, const torch::Tensor &W1,
                                                                                             bool with_transposed) {
  check_rowmajor(W0, "W0"); check_rowmajor(W1, "W1");
  TORCH_CHECK(W0.device() == W1.device(), "both weights must be on the same device");
  c10::cuda::CUDAGuard g(W0.device());
  auto mk = [&](const torch::Tensor &W) {
    const int64_t N = W.size(0), K = W.size(1);
    return std::make_pair(torch::empty({2, N, (K + 31) / 32 * 32}, W.options()),
                          with_transposed ? torch::empty({2, K, (N + 31) / 32 * 32}, W.options()) : torch::empty({0}, W.options()));
  };
  auto a = mk(W0), b = mk(W1);
  check_rc(gnn_linear_split_weights2_f32(W0.data_ptr<float>(), ld_of(W0), W0.size(0), W0.size(1), a.first.data_ptr<float>(),
                                         with_transposed ? a.second.data_ptr<float>() : nullptr, W1.data_ptr<float>(), ld_of(W1),
                                         W1.size(0), W1.size(1), b.first.data_ptr<float>(),
                                         with_transposed ? b.second.data_ptr<float>() : nullptr, cur_stream()),
           "gnn_linear_split_weights2_f32");
  return {a.first, a.second, b.first, b.second};
}

// out[:, 0:N] = A[rows] . W^T + bias   (out may be a column slice of a wider row-major buffer)
void linear_tf32x3(const torch::Tensor &A, const c10::optional<torch::Tensor> &rows, const torch::Tensor &w_split, int64_t K,
                   const c10::optional<torch::Tensor> &bias, torch::Tensor out, const c10::optional<torch::Tensor> &out_rows,
                   bool accumulate) {
  check_rowmajor(A, "A"); check_rowmajor(out, "out");
  TORCH_CHECK(w_split.is_cuda() && w_split.is_contiguous() && w_split.scalar_type() == torch::kFloat && w_split.dim() == 3 &&
                  w_split.size(0) == 2 && w_split.size(2) == (K + 31) / 32 * 32,
              "w_split must be the [2, N, ceil32(K)] result of linear_split_weights");
  TORCH_CHECK(A.size(1) == K, "A must have K = ", K, " columns, got ", A.sizes());
  const bool has_rows = rows.has_value() && rows.value().defined();
  const bool scatter = out_rows.has_value() && out_rows.value().defined();
  const int64_t M = scatter ? out_rows.value().numel() : out.size(0), N = w_split.size(1);
  TORCH_CHECK(out.size(1) == N && out.device() == A.device() && w_split.device() == A.device(), "out must be [M, ", N, "] on A's device");
  TORCH_CHECK(has_rows || A.size(0) == M, "A must have one row per output row");
  TORCH_CHECK(!scatter || accumulate, "scattered output rows are always accumulated: pass accumulate=True");
  const float *b = nullptr;
  if (bias.has_value() && bias.value().defined()) {
    const auto &t = bias.value();
    TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kFloat && t.numel() == N, "bias must be N floats");
    b = t.data_ptr<float>();
  }
  c10::cuda::CUDAGuard g(A.device());
  check_rc(gnn_linear_tf32x3_f32_ex(A.data_ptr<float>(), ld_of(A), rows_ptr(rows, M, A.device()), M, K, w_split.data_ptr<float>(), N, b,
                                    out.data_ptr<float>(), ld_of(out), rows_ptr(out_rows, M, A.device()),
                                    accumulate ? GNN_LINEAR_ACCUMULATE : 0u, cur_stream()),
           "gnn_linear_tf32x3_f32");
}

// dW[N,K] = dY^T . X[rows];  with_bias: also db[N] = column sums of dY (from the same pass)
std::tuple<torch::Tensor, torch::Tensor> linear_wgrad_tf32x3(const torch::Tensor &dY, const torch::Tensor &X,
                                                               const c10::optional<torch::Tensor> &rows, bool with_bias) {
  check_rowmajor(dY, "dY"); check_rowmajor(X, "X");
  TORCH_CHECK(dY.device() == X.device(), "dY and X must be on the same device");
  const int64_t M = dY.size(0), N = dY.size(1), K = X.size(1);
  const bool has_rows = rows.has_value() && rows.value().defined();
  TORCH_CHECK(has_rows || X.size(0) == M, "X must have one row per row of dY");
  c10::cuda::CUDAGuard g(X.device());
  auto dW = torch::empty({N, K}, X.options());
  auto db = with_bias ? torch::empty({N}, X.options()) : torch::empty({0}, X.options());
  const size_t wsb = gnn_linear_wgrad_workspace_bytes(M, N, K);
  auto ws = workspace(wsb, X.device());
  check_rc(gnn_linear_wgrad_tf32x3_f32(dY.data_ptr<float>(), ld_of(dY), X.data_ptr<float>(), ld_of(X), rows_ptr(rows, M, X.device()), M, N, K,
                                       dW.data_ptr<float>(), K, with_bias ? db.data_ptr<float>() : nullptr, ws.data_ptr(), wsb, cur_stream()),
           "gnn_linear_wgrad_tf32x3_f32");
  return {dW, db};
}

// ---- peer-mappable feature shards ---------------------------------------
std::tuple<torch::Tensor, py::bytes> shard_alloc(int64_t rows, int64_t ld, int64_t device_index) {
  c10::cuda::CUDAGuard g(c10::Device(c10::kCUDA, (c10::DeviceIndex)device_index));
  void *p = nullptr;
  unsigned char h[64];
  check_rc(gnn_shard_alloc((size_t)rows * ld * sizeof(float), &p, h), "gnn_shard_alloc");
  auto t = torch::from_blob(p, {rows, ld}, [](void *q) { gnn_shard_free(q); },
                            torch::TensorOptions().dtype(torch::kFloat).device(torch::kCUDA, device_index));
  py::gil_scoped_acquire acq;
  return {t, py::bytes(reinterpret_cast<const char *>(h), 64)};
}

torch::Tensor shard_open(const std::string &handle, int64_t rows, int64_t ld, int64_t device_index) {
  TORCH_CHECK(handle.size() == 64, "IPC handle must be 64 bytes");
  c10::cuda::CUDAGuard g(c10::Device(c10::kCUDA, (c10::DeviceIndex)device_index));
  void *p = nullptr;
  check_rc(gnn_shard_open(reinterpret_cast<const unsigned char *>(handle.data()), &p), "gnn_shard_open");
  return torch::from_blob(p, {rows, ld}, [](void *q) { gnn_shard_close(q); },
                          torch::TensorOptions().dtype(torch::kFloat).device(torch::kCUDA, device_index));
}

int64_t host_register(const torch::Tensor &host) {
  TORCH_CHECK(!host.is_cuda() && host.is_contiguous(), "host table must be a contiguous CPU tensor");
  void *alias = nullptr;
  check_rc(gnn_host_register(host.data_ptr(), (size_t)host.numel() * host.element_size(), &alias), "gnn_host_register");
  return (int64_t)reinterpret_cast<uintptr_t>(alias);
}

void host_unregister(const torch::Tensor &host) { check_rc(gnn_host_unregister(host.data_ptr()), "gnn_host_unregister"); }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  using rel = py::call_guard<py::gil_scoped_release>;
  // reference names (spmm.cpp:52-56)
  m.def("spmm_naive", &spmm_naive, "Sparse-Dense Matrix Multiplication", rel());
  m.def("spmm_load_balance", &spmm_load_balance, "Sparse-Dense Matrix Multiplication", rel());
  m.def("create_coo_tensor", &create_coo_tensor, "Create a PyTorch sparse tensor", rel());
  // B200 path
  m.def("build_adj", &build_adj, "create_coo_tensor that also returns the int32 column ids (CSR for the kernels)", rel());
  m.def("coo_to_csr", &coo_to_csr, "coalesced COO -> (rowptr int32, colidx int32)", rel());
  m.def("csr_spmm", &csr_spmm, "Y = A.X with A in CSR (optional per-entry row ids)", py::arg("rowptr"), py::arg("colidx"),
        py::arg("vals"), py::arg("M"), py::arg("K"), py::arg("dense"), py::arg("rowidx") = py::none(), py::arg("padded_rows") = false, rel());
  m.def("csr_transpose", &csr_transpose, "CSR of A^T, deterministic", rel());
  m.def("csr_spmm_t", &csr_spmm_t, "dX = A^T.G from A's own CSR (transpose-free, vector reductions)", py::arg("rowptr"),
        py::arg("colidx"), py::arg("vals"), py::arg("M"), py::arg("K"), py::arg("grad"), py::arg("rowidx") = py::none(), rel());
  m.def("probe_row_gather", &probe_row_gather, "row gather only (L2->SM roof probe); returns bytes gathered", rel());
  m.def("set_transpose_budget", [](int64_t bytes) {
    const int64_t prev = gnn_set_transpose_budget(bytes);
    TORCH_CHECK(prev >= 0, "set_transpose_budget: ", gnn_error_string((int)prev));
    return prev;
  });
  m.def("gather_spmm", &gather_spmm, "Y = A.gather(xrows) without materialising the gathered rows", py::arg("rowptr"),
        py::arg("colidx"), py::arg("vals"), py::arg("M"), py::arg("K"), py::arg("D"), py::arg("xrows"),
        py::arg("rowidx") = py::none(), rel());
  m.def("placement_remap", &placement_remap, "device placement remap -> (src_dev, slot, xrows, counts)", rel());
  m.def("gather_rows", &gather_rows, "out[j] = *xrows[j]", rel());
  m.def("gather_rows_src", &gather_rows_src, "gather only the rows of one source", rel());
  m.def("index_rows", &index_rows, "out[i] = X[idx[i]]", rel());
  m.def("row_slice_count", &row_slice_count, "fullrowptr of lap_matrix[nodes, :]", rel());
  m.def("row_slice_fill", &row_slice_fill, "column ids of lap_matrix[nodes, :] (+ column counts)", rel());
  m.def("member_set", &member_set, "membership bitmap + word ranks of after_nodes (set) / cleared again (unset)", rel());
  m.def("column_slice_count", &column_slice_count, "rowptr of U[:, after_nodes]", rel());
  m.def("column_slice_fill", &column_slice_fill, "local column ids of U[:, after_nodes]", rel());
  m.def("ladies_layer_device", &ladies_layer_device, "one LADIES layer: device passes + host draw + uploads, GIL released", rel());
  m.def("elu_rownorm_fwd", &elu_rownorm_fwd, "y, mean, rstd = rownorm(elu(x)) * scale + offset", rel());
  m.def("elu_rownorm_bwd", &elu_rownorm_bwd, "dx, dscale, doffset", rel());
  m.def("linear_split_weights", &linear_split_weights, "W[N,K] -> TF32 hi/lo planes (w_nk, w_kn)", rel());
  m.def("linear_split_weights2", &linear_split_weights2, "two weight matrices -> TF32 planes in one launch", rel());
  m.def("linear_tf32x3", &linear_tf32x3, "out = A[rows] . W^T + bias on tcgen05 (3xTF32)", py::arg("A"), py::arg("rows"),
        py::arg("w_split"), py::arg("K"), py::arg("bias"), py::arg("out"), py::arg("out_rows") = py::none(), py::arg("accumulate") = false, rel());
  m.def("linear_wgrad_tf32x3", &linear_wgrad_tf32x3, "dW = dY^T . X[rows] on tcgen05 (3xTF32)", py::arg("dY"), py::arg("X"),
        py::arg("rows"), py::arg("with_bias") = false, rel());
  m.def("shard_alloc", &shard_alloc, "peer-mappable feature shard + IPC handle", rel());
  m.def("shard_open", &shard_open, "map a peer's shard", rel());
  m.def("host_register", &host_register, "pin+map a host table, returns the device alias", rel());
  m.def("host_unregister", &host_unregister, rel());
  m.def("launch_count", []() { return gnn_launch_count(); });
  m.def("set_corunner_ctas", [](int ctas) {
    const int prev = gnn_set_corunner_ctas(ctas);
    TORCH_CHECK(prev >= 0, "set_corunner_ctas: ", gnn_error_string(prev));
    return prev;
  });
  m.def("host_gather_ctas", []() { return gnn_host_gather_ctas(); });
  m.def("set_blocking_sync", [](int mode) { check_rc(gnn_set_blocking_sync(mode), "gnn_set_blocking_sync"); });
  m.def("abi_version", []() { return gnn_abi_version(); });
}
