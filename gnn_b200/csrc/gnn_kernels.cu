// gnn_kernels.cu - hand-written sm_100a kernels + the C ABI of include/gnn_b200.h.
//
// Path: LADIES-layer aggregation SpMM (forward A.X, backward A^T.G), adjacency
// construction, placement remap and input-feature gather of HPC-Research-Lab/GNN
// (reference spmm_cpp/cuda_spmm.cu, spmm_cpp/spmm.cpp, custom_sparse_ops.py,
// sampler.py:133-158, main.py:129-134).  Built from scratch for B200: no torch
// types, no host synchronisation, no allocation, every launch on the caller's
// stream.  The path is sparse, fp32 and bandwidth/LSU-bound: SIMT kernels with
// 128-bit coalesced row loads, warp-shuffle broadcast of the (col,val) stream,
// nnz-balanced chunking with a fixed-order fix-up instead of atomics.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gnn_b200.h"

namespace {

std::atomic<int64_t> g_launches{0};

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr unsigned kFull = 0xffffffffu;

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

#define GNN_LAUNCH_CHECK()                         \
  do {                                             \
    g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e_ = cudaGetLastError();           \
    if (e_ != cudaSuccess) return (int)e_;         \
  } while (0)

#define GNN_CUDA(call)                             \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) return (int)e_;         \
  } while (0)

// ---------------------------------------------------------------------------
// vector helpers
// ---------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void vzero(float (&a)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) a[i] = 0.f;
}

// read-only (non-coherent) vector load of VEC floats; p is VEC*4-byte aligned
template <int VEC>
__device__ __forceinline__ void ldg_vec(const float *p, float (&x)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
  } else if constexpr (VEC == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    x[0] = t.x; x[1] = t.y;
  } else {
    x[0] = __ldg(p);
  }
}

#include "chunk_rows.cuh"

// ---------------------------------------------------------------------------
// Row-split CSR SpMM
//
// Work item = (chunk of C consecutive nonzeros, column slab).  One warp per item.
// The warp walks the rows its chunk touches; a row that lies inside one chunk is
// stored directly, a row that spans several chunks is written as a partial to the
// workspace and the warp that arrives last adds the partials in ascending chunk
// order (fixed order => bit-reproducible), then stores the row.  Rows without
// nonzeros are zero-filled by the chunk that finishes the preceding non-empty row
// (chunk 0 also covers leading empty rows), so every output row is written once.
//
// Lane layout, LPR == 32: lane l owns VEC floats at column slab0 + (n*32 + l)*VEC
// for n < NV.  LPR < 32 (narrow D): the warp holds 32/LPR row-groups; group g takes
// nonzeros k = g (mod 32/LPR) and the groups are combined by xor-shuffles.
// ---------------------------------------------------------------------------
template <bool GATHER>
struct XSrc {
  const float *X;
  int ldx;                       // elements; < 2^31 (checked on the host)
  const float *const *xrows;
  // start of X row c (one IMAD.WIDE for the dense case)
  __device__ __forceinline__ const float *row(int c) const {
    if constexpr (GATHER) return reinterpret_cast<const float *>(__ldg(reinterpret_cast<const unsigned long long *>(xrows) + c));
    else return X + (int64_t)c * ldx;
  }
};

// nonzeros kept in flight per warp step (U) and CTAs per SM asked of ptxas (MINB), by vectors per lane
// (tuned on B200 over the Reddit-shaped blocks with wave fitting active, profiles/tune_r1.txt)
[[maybe_unused]] constexpr int default_u(int nv) { return nv >= 5 ? 1 : (nv == 4 ? 2 : (nv >= 2 ? 4 : 8)); }
[[maybe_unused]] constexpr int default_minb(int nv) { return (nv == 3 || nv == 4) ? 3 : (nv >= 6 ? 2 : 4); }

// Dload = floats readable per X row (D, or D rounded up to 4 when rows are padded to 16 bytes)
template <int VEC, int NV, int LPR, bool GATHER, int U>
__device__ __forceinline__ void accumulate_segment(const int *__restrict__ colidx, const float *__restrict__ vals,
                                                   const XSrc<GATHER> &xs, int s, int e, int lane, int col0, int Dload,
                                                   float (&acc)[NV][VEC]) {
  constexpr int G = 32 / LPR;                       // row groups in the warp
  const int g = lane / LPR;
  bool colok[NV];
#pragma unroll
  for (int n = 0; n < NV; ++n) colok[n] = col0 + n * LPR * VEC + VEC <= Dload;

  for (int base = s; base < e; base += 32) {
    const int i = base + lane;
    int cl = 0;
    float vl = 0.f;
    if (i < e) { cl = __ldg(colidx + i); vl = __ldg(vals + i); }
    const int n_here = min(32, e - base);
    // each group visits k = t*G + g
    const int steps = (n_here + G - 1) / G;
    int t = 0;
    for (; t + U <= steps; t += U) {
      int c[U]; float v[U]; bool ok[U];
      float x[U][NV][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = (t + u) * G + g;
        c[u] = __shfl_sync(kFull, cl, k & 31);
        v[u] = __shfl_sync(kFull, vl, k & 31);
        ok[u] = k < n_here;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float *xr = xs.row(c[u]) + col0;
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          if (ok[u] && colok[n]) ldg_vec<VEC>(xr + n * LPR * VEC, x[u][n]);
          else vzero<VEC>(x[u][n]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float vv = ok[u] ? v[u] : 0.f;
#pragma unroll
        for (int n = 0; n < NV; ++n)
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[n][q] = fmaf(vv, x[u][n][q], acc[n][q]);
      }
    }
    for (; t < steps; ++t) {
      const int k = t * G + g;
      const int c = __shfl_sync(kFull, cl, k & 31);
      const float v = __shfl_sync(kFull, vl, k & 31);
      if (k < n_here) {
        const float *xr = xs.row(c) + col0;
        float x[NV][VEC];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          if (colok[n]) ldg_vec<VEC>(xr + n * LPR * VEC, x[n]);
          else vzero<VEC>(x[n]);
        }
#pragma unroll
        for (int n = 0; n < NV; ++n)
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[n][q] = fmaf(v, x[n][q], acc[n][q]);
      }
    }
  }
  if constexpr (G > 1) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
      for (int n = 0; n < NV; ++n)
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[n][q] += __shfl_xor_sync(kFull, acc[n][q], off);
  }
}

// store VEC floats per (lane, n) to a row; scalar when the row is not VEC-aligned
template <int VEC, int NV, int LPR>
__device__ __forceinline__ void store_row(float *row, bool vec_ok, int lane, int col0, int D, const float (&acc)[NV][VEC]) {
  if (lane >= LPR) return;
#pragma unroll
  for (int n = 0; n < NV; ++n) {
    const int col = col0 + n * LPR * VEC;
    if (col >= D) continue;
    if (VEC == 4 && vec_ok && col + 4 <= D) {
      *reinterpret_cast<float4 *>(row + col) = make_float4(acc[n][0], acc[n][1 % VEC], acc[n][2 % VEC], acc[n][3 % VEC]);
    } else if (VEC == 2 && vec_ok && col + 2 <= D) {
      *reinterpret_cast<float2 *>(row + col) = make_float2(acc[n][0], acc[n][1 % VEC]);
    } else {
#pragma unroll
      for (int q = 0; q < VEC; ++q)
        if (col + q < D) row[col + q] = acc[n][q];
    }
  }
}

template <int VEC, int NV, int LPR>
__device__ __forceinline__ void zero_row(float *row, bool vec_ok, int lane, int col0, int D) {
  float z[NV][VEC];
#pragma unroll
  for (int n = 0; n < NV; ++n) vzero<VEC>(z[n]);
  store_row<VEC, NV, LPR>(row, vec_ok, lane, col0, D, z);
}

struct SpmmParams {
  const int *rowptr;
  const int *rowidx;   // row id per stored entry (optional, flat kernel only)
  const int *colidx;
  const float *vals;
  int M;
  int nnz;
  int D;
  int C;          // nonzeros per chunk
  int nchunks;
  int nslabs;
  float *Y;
  int64_t ldy;
  float *partials;   // [2*nchunks][Dp]
  int *counters;     // [M*nslabs], zero on entry
  int Dp;
  int Dload;         // floats readable per X row (>= D)
  int cshift;        // log2(C) (flat kernel: C is a power of two)
};

#include "spmm_flat.cuh"
#include "spmm_scatter_t.cuh"

template <int VEC, int NV, int LPR, bool GATHER, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
spmm_rowsplit_kernel(const SpmmParams p, const XSrc<GATHER> xs) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t item = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  if (item >= (int64_t)p.nchunks * p.nslabs) return;
  const int slab = (int)(item / p.nchunks);
  const int chunk = (int)(item % p.nchunks);
  constexpr int W = NV * LPR * VEC;                 // slab width in floats
  const int sl = lane % LPR;
  const int col0 = slab * W + sl * VEC;
  const bool y_vec_ok = ((reinterpret_cast<uintptr_t>(p.Y) | (uintptr_t)(p.ldy * 4)) & (VEC * 4 - 1)) == 0;

  int s = chunk * p.C;
  const int e = min(s + p.C, p.nnz);

  // first row with rowptr[r+1] > s: warp-wide 32-ary search (3 dependent loads for M <= 32 K instead of 15)
  const int lo = warp_first_row(p.rowptr, p.M, s, lane);
  int r = lo;
  if (chunk == 0) {  // leading empty rows
    for (int z = 0; z < r; ++z) zero_row<VEC, NV, LPR>(p.Y + (int64_t)z * p.ldy, y_vec_ok, lane, col0, p.D);
  }

  while (s < e) {
    const int row_start = __ldg(p.rowptr + r);
    const int row_end = __ldg(p.rowptr + r + 1);
    const int seg_end = min(row_end, e);
    float acc[NV][VEC];
#pragma unroll
    for (int n = 0; n < NV; ++n) vzero<VEC>(acc[n]);
    accumulate_segment<VEC, NV, LPR, GATHER, U>(p.colidx, p.vals, xs, s, seg_end, lane, col0, p.Dload, acc);

    const int c_first = row_start / p.C;
    const int c_last = (row_end - 1) / p.C;
    float *yrow = p.Y + (int64_t)r * p.ldy;
    if (c_first == c_last) {
      store_row<VEC, NV, LPR>(yrow, y_vec_ok, lane, col0, p.D, acc);
    } else {
      // partial of (row r, chunk): slot 2*chunk + (row starts strictly inside this chunk); the last chunk to arrive adds
      // the partials in ascending chunk order (spmm_flat.cuh: finish_spanning_row)
      float *slot = p.partials + ((int64_t)2 * chunk + (row_start > chunk * p.C ? 1 : 0)) * p.Dp;
      store_row<VEC, NV, LPR>(slot, true, lane, col0, p.D, acc);
      finish_spanning_row<VEC, NV, LPR>(p, r, slab, lane, col0, row_start, c_first, c_last, p.C, yrow, y_vec_ok);
    }
    s = seg_end;
    if (seg_end == row_end) {
      // this chunk finished row r: it also owns the empty rows that follow
      ++r;
      while (r < p.M && __ldg(p.rowptr + r + 1) == row_end) {
        zero_row<VEC, NV, LPR>(p.Y + (int64_t)r * p.ldy, y_vec_ok, lane, col0, p.D);
        ++r;
      }
    }
  }
}

// chunk size: a function of (nnz, D) only so that the workspace query and the launch agree.
// Large blocks use 1024 nonzeros per warp item; small ones shrink the chunk so that a few thousand
// warp items exist (the top LADIES layer is 30 K nonzeros: latency, not bandwidth, decides it).
inline int spmm_chunk(int64_t nnz, int64_t D) {
#ifdef GNN_TUNE
  if (getenv("GNN_TUNE_C")) return atoi(getenv("GNN_TUNE_C"));
#endif
  const int64_t want = nnz * cdiv(D, 128) / 4096;
  int c = 64;
  while (c < 1024 && c * 2 <= want) c *= 2;
  return c;
}

// Which kernel: the flat (nonzero-split) kernel for short-row or small blocks - top LADIES layer, the transpose of
// any sparse layer, papers100M-shaped blocks - where the per-row dependent loads of the row-split kernel, not
// bandwidth, are the run time; the row-split kernel (longer register-resident row segments, fewer shared-memory
// trips per nonzero) for the dense LADIES blocks.  A function of (M, nnz, D) only: the workspace query must agree.
constexpr int64_t kFlatMaxNnz = 4 << 20;
constexpr int64_t kFlatMeanRow = 24;           // measured crossover (profiles/r2_kernel_ab.md): below ~24 entries per row
constexpr int64_t kFlatTargetItems = 8192;     // (scatter backward) warp items wanted before chunks grow

inline bool flat_wanted(int64_t M, int64_t nnz, int64_t D) {
  (void)D;
#ifdef GNN_TUNE
  if (getenv("GNN_TUNE_FLAT")) return atoi(getenv("GNN_TUNE_FLAT")) != 0;
#endif
  return nnz <= kFlatMaxNnz && nnz < kFlatMeanRow * std::max<int64_t>(M, 1);
}

// flat chunk (entries per warp item): 64; 128 for the widest rows; 32 when the whole problem is a fraction of a wave
// anyway (profiles/r2_kernel_ab.md)
inline bool flat_tiny(int64_t nnz, int64_t D) { return cdiv(nnz, 64) * cdiv(D, 128) < 2048; }
inline int flat_chunk(int64_t nnz, int64_t D) {
#ifdef GNN_TUNE
  if (getenv("GNN_TUNE_FC")) return atoi(getenv("GNN_TUNE_FC"));
#endif
  if (flat_tiny(nnz, D)) return 32;
  return D >= 1024 ? 128 : 64;
}

struct SpmmPlan { int kind, vec, nv, lpr, nslabs, C, nchunks, Dp, u; };   // kind: 0 row-split, 1 flat

constexpr int64_t kTargetItems = 3072;   // warp items (~2/3 of a wave) wanted before wider slabs are preferred

inline SpmmPlan make_plan(int64_t M, int64_t nnz, int64_t D, int vec) {
  SpmmPlan pl;
  pl.vec = vec;
  pl.Dp = (int)(cdiv(D, 4) * 4);
  pl.u = 0;
  const int64_t nvec = cdiv(D, vec);               // vectors per row
  if (flat_wanted(M, nnz, D)) {
    pl.kind = 1;
    pl.lpr = 32;
    pl.C = flat_chunk(nnz, D);
    pl.nchunks = (int)cdiv(nnz, pl.C);
    const int64_t n = cdiv(nvec, 32);              // vector columns per lane
    // vectors per lane: wide rows walk the index stream fewer times (D >= 512: 4, D >= 256: 2); tiny problems keep the
    // most warps (1)
    pl.nv = flat_tiny(nnz, D) ? 1 : (D >= 512 && n >= 4 ? 4 : (D >= 256 && n >= 2 ? 2 : 1));
    pl.u = pl.nv == 1 ? 16 : (pl.nv == 2 ? 8 : 4);
#ifdef GNN_TUNE
    if (getenv("GNN_TUNE_FNV")) pl.nv = atoi(getenv("GNN_TUNE_FNV"));
    if (getenv("GNN_TUNE_FU")) pl.u = atoi(getenv("GNN_TUNE_FU"));
#endif
    pl.nslabs = (int)cdiv(n, pl.nv);
    return pl;
  }
  pl.kind = 0;
  pl.C = spmm_chunk(nnz, D);
  pl.nchunks = (int)cdiv(nnz, pl.C);
  if (vec == 4 && nvec <= 16) {                    // narrow rows: several nonzeros per warp step
    pl.lpr = nvec <= 4 ? 4 : (nvec <= 8 ? 8 : 16);
    pl.nv = 1;
    pl.nslabs = 1;
    return pl;
  }
  pl.lpr = 32;
  const int64_t n = cdiv(nvec, 32);                // vector columns per lane
  // vectors per lane: least padded width, then fewest slabs - among the choices that still give
  // enough warp items; the narrowest slab when nothing does
  static const int cand[] = {5, 4, 3, 2, 1};
  double best = 1e30;
  pl.nv = 1;
  for (int nv : cand) {
    const int64_t slabs = cdiv(n, nv);
    if (nv > 1 && (int64_t)pl.nchunks * slabs < kTargetItems) continue;
    const double cost = (double)(nv * slabs) + 0.5 * (double)slabs;
    if (cost < best - 1e-9) { best = cost; pl.nv = nv; }
  }
#ifdef GNN_TUNE
  if (getenv("GNN_TUNE_NV")) pl.nv = atoi(getenv("GNN_TUNE_NV"));
#endif
  pl.nslabs = (int)cdiv(n, pl.nv);
  return pl;
}

// counters: one int per (row, slab); the narrowest layout (scalar loads, one vector per lane) has ceil(D/32) slabs
inline size_t spmm_counter_bytes(int64_t M, int64_t nnz, int64_t D) {
  (void)nnz;
  return ((size_t)M * (size_t)cdiv(D, 32) * sizeof(int) + 255) / 256 * 256;
}

// Wave fitting: every warp item costs the same (C nonzeros of one slab), so the kernel runs in waves of
// `slots` = SMs x resident warps.  A grid that is 1 % over a wave boundary takes a whole extra wave (+24 % measured
// at 1.02 waves), so the chunk is shrunk - never below half the base size, which the workspace bound assumes - until
// the items fill an integral number of waves as exactly as possible.
constexpr int kHostGatherCtas = 16;

// CTA slots that kernels on OTHER streams hold while an SpMM runs (gnn_set_corunner_ctas).  A one-wave grid that
// finds even one slot taken by a long-running foreign CTA spills into a second wave: measured 1.46 -> 2.0 ms per
// Reddit-shaped minibatch beside a single-CTA host-row gather.  The wave fit therefore leaves these slots out.
std::atomic<int> g_corunner_ctas{0};

inline int wave_fit_chunk(int64_t nnz, int nslabs, int base_c, int64_t slots) {
  if (slots <= 0) return base_c;
  const int64_t items0 = cdiv(nnz, base_c) * nslabs;
  if (2 * items0 < slots) return base_c;            // small problems are latency-bound: keep the tuned chunk
  const int64_t waves = std::max<int64_t>(1, cdiv(items0, slots));
  const int64_t chunks_per_slab = std::max<int64_t>(1, waves * slots / nslabs);
  int64_t c = cdiv(nnz, chunks_per_slab);
  c = std::max<int64_t>(c, std::max(32, base_c / 2));
  c = std::min<int64_t>(c, base_c);
  return (int)c;
}

inline int device_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

template <int VEC, int NV, int LPR, bool GATHER, int U = default_u(NV), int MINB = default_minb(NV)>
int launch_spmm_t(const SpmmParams &p0, const XSrc<GATHER> &xs, cudaStream_t st) {
  static std::atomic<int> ctas_per_sm{0};            // of this instantiation (registers decide it)
  int occ = ctas_per_sm.load(std::memory_order_relaxed);
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmm_rowsplit_kernel<VEC, NV, LPR, GATHER, U, MINB>, kThreads, 0) !=
            cudaSuccess || occ <= 0) {
      (void)cudaGetLastError();
      occ = MINB;
    }
    ctas_per_sm.store(occ, std::memory_order_relaxed);
  }
  SpmmParams p = p0;
#ifdef GNN_TUNE
  if (!getenv("GNN_TUNE_C"))
#endif
  {
    const int64_t ctas = std::max<int64_t>(1, (int64_t)device_sm_count() * occ - g_corunner_ctas.load(std::memory_order_relaxed));
    p.C = wave_fit_chunk(p.nnz, p.nslabs, p0.C, ctas * kWarpsPerCta);
    p.nchunks = (int)cdiv(p.nnz, p.C);
  }
  const int64_t items = (int64_t)p.nchunks * p.nslabs;
  const unsigned grid = (unsigned)cdiv(items, kWarpsPerCta);
  spmm_rowsplit_kernel<VEC, NV, LPR, GATHER, U, MINB><<<grid, kThreads, 0, st>>>(p, xs);
  GNN_LAUNCH_CHECK();
  return 0;
}

#ifdef GNN_TUNE
// experiment build only (tools/tune_spmm.py): pick (NV, U, MINB) from the environment
inline int env_int(const char *name, int dflt) { const char *v = getenv(name); return v ? atoi(v) : dflt; }
template <int NV, int U>
int launch_tune_minb(int minb, const SpmmParams &p, const XSrc<false> &xs, cudaStream_t st) {
  switch (minb) {
    case 2: return launch_spmm_t<4, NV, 32, false, U, 2>(p, xs, st);
    case 3: return launch_spmm_t<4, NV, 32, false, U, 3>(p, xs, st);
    case 4: return launch_spmm_t<4, NV, 32, false, U, 4>(p, xs, st);
    default: return launch_spmm_t<4, NV, 32, false, U, 6>(p, xs, st);
  }
}
template <int NV>
int launch_tune_u(int u, int minb, const SpmmParams &p, const XSrc<false> &xs, cudaStream_t st) {
  switch (u) {
    case 1: return launch_tune_minb<NV, 1>(minb, p, xs, st);
    case 2: return launch_tune_minb<NV, 2>(minb, p, xs, st);
    case 4: return launch_tune_minb<NV, 4>(minb, p, xs, st);
    default: return launch_tune_minb<NV, 8>(minb, p, xs, st);
  }
}
inline int launch_tune(const SpmmPlan &pl, const SpmmParams &p, const XSrc<false> &xs, cudaStream_t st) {
  const int u = env_int("GNN_TUNE_U", 1), minb = env_int("GNN_TUNE_MINB", 2);
  switch (pl.nv) {
    case 1: return launch_tune_u<1>(u, minb, p, xs, st);
    case 2: return launch_tune_u<2>(u, minb, p, xs, st);
    case 3: return launch_tune_u<3>(u, minb, p, xs, st);
    case 4: return launch_tune_u<4>(u, minb, p, xs, st);
    case 5: return launch_tune_u<5>(u, minb, p, xs, st);
    case 6: return launch_tune_u<6>(u, minb, p, xs, st);
    default: return launch_tune_u<8>(u, minb, p, xs, st);
  }
}
#endif

template <int VEC, int NV, bool GATHER, int U>
int launch_flat_t(const SpmmParams &p, const XSrc<GATHER> &xs, cudaStream_t st) {
  const int64_t items = (int64_t)p.nchunks * p.nslabs;
  const unsigned grid = (unsigned)cdiv(items, kFlatWarps);
  if (p.rowidx) spmm_flat_kernel<VEC, NV, GATHER, U, true><<<grid, kFlatWarps * 32, 0, st>>>(p, xs);
  else spmm_flat_kernel<VEC, NV, GATHER, U, false><<<grid, kFlatWarps * 32, 0, st>>>(p, xs);
  GNN_LAUNCH_CHECK();
  return 0;
}

template <int VEC, bool GATHER>
int launch_flat(const SpmmPlan &pl, const SpmmParams &p, const XSrc<GATHER> &xs, cudaStream_t st) {
  switch (pl.nv) {
    case 1: return pl.u >= 16 ? launch_flat_t<VEC, 1, GATHER, 16>(p, xs, st) : launch_flat_t<VEC, 1, GATHER, 8>(p, xs, st);
    case 2: return launch_flat_t<VEC, 2, GATHER, 8>(p, xs, st);
    case 4: return launch_flat_t<VEC, 4, GATHER, 4>(p, xs, st);
    default: return GNN_E_BADARG;
  }
}

template <int VEC, bool GATHER>
int launch_spmm_nv(const SpmmPlan &pl, const SpmmParams &p, const XSrc<GATHER> &xs, cudaStream_t st) {
  if (pl.kind == 1) return launch_flat<VEC, GATHER>(pl, p, xs, st);
#ifdef GNN_TUNE
  if constexpr (VEC == 4 && !GATHER) {
    if (pl.lpr == 32 && getenv("GNN_TUNE_U")) return launch_tune(pl, p, xs, st);
  }
#endif
  if (pl.lpr != 32) {
    if constexpr (VEC == 4) {
      switch (pl.lpr) {
        case 4: return launch_spmm_t<4, 1, 4, GATHER>(p, xs, st);
        case 8: return launch_spmm_t<4, 1, 8, GATHER>(p, xs, st);
        default: return launch_spmm_t<4, 1, 16, GATHER>(p, xs, st);
      }
    } else {
      return GNN_E_BADARG;
    }
  }
  switch (pl.nv) {
    case 1: return launch_spmm_t<VEC, 1, 32, GATHER>(p, xs, st);
    case 2: return launch_spmm_t<VEC, 2, 32, GATHER>(p, xs, st);
    case 3: return launch_spmm_t<VEC, 3, 32, GATHER>(p, xs, st);
    case 4: return launch_spmm_t<VEC, 4, 32, GATHER>(p, xs, st);
    default: return launch_spmm_t<VEC, 5, 32, GATHER>(p, xs, st);
  }
}

__global__ void zero_rows_kernel(float *Y, int64_t ldy, int64_t M, int64_t D) {
  const int64_t n = M * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    Y[(i / D) * ldy + (i % D)] = 0.f;
}

template <bool GATHER>
int spmm_entry(const int32_t *rowptr, const int32_t *rowidx, const int32_t *colidx, const float *vals, int64_t M, int64_t K, int64_t nnz,
               int64_t D, const float *X, int64_t ldx, const float *const *xrows, float *Y, int64_t ldy,
               int *counters, void *partials, size_t partial_bytes, unsigned flags, cudaStream_t st) {
  if (M < 0 || K < 0 || nnz < 0 || D < 0) return GNN_E_BADARG;
  if (flags & ~(unsigned)GNN_SPMM_COUNTERS_ZEROED) return GNN_E_BADARG;
  if (M == 0 || D == 0) return 0;
  if (!Y || ldy < D) return GNN_E_BADARG;
  if (M >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 2048 || K >= (1ll << 31) || D >= (1ll << 24) || ldx >= (1ll << 31)) return GNN_E_RANGE;
  if (nnz == 0) {
    zero_rows_kernel<<<(unsigned)std::min<int64_t>(cdiv(M * D, 256), 148 * 16), 256, 0, st>>>(Y, ldy, M, D);
    GNN_LAUNCH_CHECK();
    return 0;
  }
  if (!rowptr || !colidx || !vals) return GNN_E_BADARG;
  if (GATHER ? (xrows == nullptr) : (X == nullptr || ldx < D)) return GNN_E_BADARG;
  int vec = 1;
  int64_t Dload = D;
  const int64_t D4 = cdiv(D, 4) * 4;
  if (GATHER) {
    vec = 4;   // contract: every row pointer is 16-byte aligned and readable up to ceil(D/4)*4 floats
    Dload = D4;
  } else {
    const uintptr_t a = reinterpret_cast<uintptr_t>(X) | (uintptr_t)(ldx * 4);
    // rows padded to 16 bytes (ldx >= ceil4(D), e.g. the gathered input buffer) may be read past D
    if ((a & 15) == 0 && (D % 4 == 0 || ldx >= D4)) { vec = 4; Dload = D4; }
    else if ((a & 7) == 0 && (D % 2 == 0 || ldx >= D + 1)) { vec = 2; Dload = cdiv(D, 2) * 2; }
  }
  const SpmmPlan pl = make_plan(M, nnz, D, vec);
  if (!counters || !partials || partial_bytes < gnn_csr_spmm_partial_bytes(M, nnz, D)) return GNN_E_WORKSPACE;

  SpmmParams p;
  p.rowptr = rowptr; p.rowidx = rowidx; p.colidx = colidx; p.vals = vals;
  p.M = (int)M; p.nnz = (int)nnz; p.D = (int)D;
  p.C = pl.C; p.nchunks = pl.nchunks; p.nslabs = pl.nslabs; p.Dp = pl.Dp;
  p.Y = Y; p.ldy = ldy; p.Dload = (int)Dload;
  p.cshift = 0;
  while ((1 << p.cshift) < pl.C) ++p.cshift;
  p.counters = counters;
  p.partials = reinterpret_cast<float *>(partials);
  // the arrival counters wrap back to zero inside the kernel (atomicInc), so a caller that keeps one workspace per
  // stream zeroes it once and passes GNN_SPMM_COUNTERS_ZEROED afterwards: one launch per SpMM instead of two
  if (!(flags & GNN_SPMM_COUNTERS_ZEROED)) GNN_CUDA(cudaMemsetAsync(p.counters, 0, (size_t)M * pl.nslabs * sizeof(int), st));
  XSrc<GATHER> xs{X, (int)ldx, xrows};
  switch (vec) {
    case 4: return launch_spmm_nv<4, GATHER>(pl, p, xs, st);
    case 2: if constexpr (!GATHER) return launch_spmm_nv<2, false>(pl, p, xs, st); else return GNN_E_BADARG;
    default: if constexpr (!GATHER) return launch_spmm_nv<1, false>(pl, p, xs, st); else return GNN_E_BADARG;
  }
}

// ---------------------------------------------------------------------------
// build_adj
// ---------------------------------------------------------------------------
constexpr int kAdjChunk = 256;

// One warp per 256 consecutive stored entries (not per row: a hub row of a LADIES layer holds thousands of entries and
// a warp-per-row grid waited 65 us for it on a 3.9 M-entry layer).  Lane-strided, so every store instruction of the
// warp covers one contiguous 128/256-byte run; the row of an entry comes from the chunk's row-pointer window in shared
// memory (chunk_rows.cuh) instead of a binary search over global memory.
template <typename ColT>
__global__ void __launch_bounds__(256)
build_adj_kernel(const int *__restrict__ fullrowptr, const int *__restrict__ rowptr, const ColT *__restrict__ colidx,
                 const float *__restrict__ normfact, int M, int64_t nnz, int64_t *__restrict__ out_idx,
                 float *__restrict__ out_vals, int *__restrict__ out_col32, int *__restrict__ out_row32) {
  constexpr int J = kAdjChunk / 32;
  __shared__ int win_s[8][32 * J + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const int64_t s64 = item * kAdjChunk;
  if (s64 >= nnz) return;
  const int s = (int)s64, e = (int)min((int64_t)s + kAdjChunk, nnz);
  int c[J];
  float nf[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int i = s + lane + 32 * j;
    c[j] = i < e ? (int)colidx[i] : 0;
  }
  ChunkRows<J> cr;
  cr.load(rowptr, M, s, e, lane, win_s[warp]);
#pragma unroll
  for (int j = 0; j < J; ++j) nf[j] = __ldg(normfact + c[j]);
  int r_prev = -1;
  double inv_deg = 0.;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int i = s + lane + 32 * j;
    if (i >= e) break;
    const int r = cr.row_of(rowptr, M, i);
    if (r != r_prev) {
      // cuda_spmm.cu:800: `1. / deg * normfact` - double quotient, double product, one rounding to float
      inv_deg = 1. / (double)(__ldg(fullrowptr + r + 1) - __ldg(fullrowptr + r));
      r_prev = r;
    }
    out_vals[i] = (float)(inv_deg * (double)nf[j]);
    if (out_col32) out_col32[i] = c[j];
    if (out_row32) out_row32[i] = r;
    if (out_idx) { out_idx[i] = r; out_idx[nnz + i] = c[j]; }
  }
}

// ---------------------------------------------------------------------------
// COO -> CSR (foreign sparse tensors)
// ---------------------------------------------------------------------------
__global__ void coo_to_csr_kernel(const int64_t *__restrict__ idx, int64_t M, int64_t nnz, int *__restrict__ rowptr,
                                  int *__restrict__ col32) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nnz) return;
  // entry i closes rows (prev, cur]: rowptr[r] = i for prev < r <= cur
  const int64_t prev = (i == 0) ? -1 : idx[i - 1];
  const int64_t cur = (i == nnz) ? M : idx[i];
  for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int)i;
  if (i < nnz) col32[i] = (int)idx[nnz + i];
}

// ---------------------------------------------------------------------------
// CSR transpose through a column-major bitmap (deterministic: bit OR is order-free)
//
// cell[c][w] = { bits of rows row0+32w..row0+32w+31 that have a nonzero in column c,
//                number of nonzeros of column c in rows [row0, row0+32w) }
// so the entry (r, c) lands at base[c] + cell.prefix + popc(cell.bits below r):
// the transposed rows come out in ascending source row without any sort.
//
// The bitmap covers a BLOCK of rows [row0, row1) at a time.  One block (the normal case: every LADIES layer of the
// benchmark shapes) = memset -> set -> prefix+scan -> fill, four stream operations.  When K*ceil(M/32)*8 bytes would
// exceed the bitmap budget (products/papers-scale samp_num, ~130 K x 130 K and up), the rows are cut into blocks that
// fit: column totals come from a histogram pass + scan, a per-column cursor carries the fill position from block to
// block.  Same output, any size, workspace bounded by the budget.
// ---------------------------------------------------------------------------
// Both passes over the nonzeros (set bits / fill) are nnz-balanced: a warp takes 256 consecutive nonzeros, finds the
// rows its chunk touches through the shared-memory row-pointer window (chunk_rows.cuh), keeps 8 independent loads per
// lane in flight.
constexpr int kTrChunk = 256;
std::atomic<int64_t> g_transpose_budget{(int64_t)512 << 20};

template <bool FILL>
__global__ void __launch_bounds__(256)
transpose_pass_kernel(uint2 *__restrict__ cells, int M, int nnz, int words_per_col, int row0, int row1,
                      const int *__restrict__ rowptr, const int *__restrict__ colidx, const float *__restrict__ vals,
                      const int *__restrict__ base, int *__restrict__ t_colidx, float *__restrict__ t_vals,
                      int *__restrict__ t_rowidx) {
  constexpr int J = kTrChunk / 32;
  __shared__ int win_s[8][32 * J + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const int64_t s64 = item * kTrChunk;
  if (s64 >= nnz) return;
  const int s = (int)s64, e = min(s + kTrChunk, nnz);
  int c[J], r[J];
  float v[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int i = s + lane + 32 * j;
    c[j] = i < e ? __ldg(colidx + i) : -1;
    if (FILL) v[j] = i < e ? __ldg(vals + i) : 0.f;
  }
  ChunkRows<J> cr;
  cr.load(rowptr, M, s, e, lane, win_s[warp]);
  if (cr.r_lo >= row1) return;                     // the whole chunk lies after this row block (warp-uniform)
#pragma unroll
  for (int j = 0; j < J; ++j) {
    r[j] = c[j] >= 0 ? cr.row_of(rowptr, M, s + lane + 32 * j) : 0;
    if (r[j] < row0 || r[j] >= row1) c[j] = -1;    // entry of another row block
    r[j] -= row0;
  }
  if (!FILL) {
#pragma unroll
    for (int j = 0; j < J; ++j)
      if (c[j] >= 0) atomicOr(&cells[(int64_t)c[j] * words_per_col + (r[j] >> 5)].x, 1u << (r[j] & 31));
  } else {
    uint2 cell[J];
    int b[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (c[j] >= 0) {
        cell[j] = __ldg(cells + (int64_t)c[j] * words_per_col + (r[j] >> 5));
        b[j] = __ldcg(base + c[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (c[j] >= 0) {
        const int pos = b[j] + (int)cell[j].y + __popc(cell[j].x & ((1u << (r[j] & 31)) - 1u));
        t_colidx[pos] = r[j] + row0;
        t_vals[pos] = v[j];
        if (t_rowidx) t_rowidx[pos] = c[j];
      }
    }
  }
}

// exclusive scan of in[0..n) into out[0..n], out[n] = total, by ONE CTA of WARPS warps.  Tiles of 1024 * WARPS elements:
// warp w owns 1024 consecutive ones as 32 coalesced rows of 32 (all loads issued together), shuffle scans per row
// with a running carry, then one 32-entry scan over the warp totals.
template <int WARPS>
__device__ __forceinline__ void block_exclusive_scan(const int *__restrict__ in, int n, int *__restrict__ out, int *sm33) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int carry = 0;
  for (int base = 0; base < n; base += 1024 * WARPS) {
    const int seg = base + warp * 1024 + lane;
    int v[32];
#pragma unroll
    for (int it = 0; it < 32; ++it) v[it] = seg + it * 32 < n ? __ldcg(in + seg + it * 32) : 0;
    int run = 0;
#pragma unroll
    for (int it = 0; it < 32; ++it) {
      int incl = v[it];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += y;
      }
      v[it] = run + incl - v[it];                       // exclusive inside the warp's segment
      run += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) sm33[warp] = run;
    __syncthreads();
    if (warp == 0) {
      int w = lane < WARPS ? sm33[lane] : 0;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(kFull, w, off);
        if (lane >= off) w += y;
      }
      sm33[lane] = w;                                   // inclusive over warps
    }
    __syncthreads();
    const int offset = carry + (warp ? sm33[warp - 1] : 0);
    carry += sm33[WARPS - 1];
#pragma unroll
    for (int it = 0; it < 32; ++it)
      if (seg + it * 32 < n) out[seg + it * 32] = v[it] + offset;
    __syncthreads();
  }
  if (tid == 0) out[n] = carry;
}

// Per column: running popcount of its bitmap words (cell.y) and the column's count in this row block.
//   cursor != NULL (multi-block mode): first cursor[c] += counts[c] (the previous block's count), then counts[c] = this block's.
//   t_rowptr != NULL (single-block mode): the CTA that finishes last turns the counts into the row pointer of A^T
//   (saves a launch; `done` is a zeroed word in the workspace).
__global__ void __launch_bounds__(256)
bitmap_prefix_kernel(uint2 *__restrict__ cells, int K, int words_per_col, int *__restrict__ counts, int *__restrict__ cursor,
                     int advance, int *__restrict__ t_rowptr, unsigned *__restrict__ done) {
  __shared__ int sm33[33];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < K; c += gridDim.x * wpb) {
    uint2 *col = cells + (int64_t)c * words_per_col;
    if (cursor && advance && lane == 0) cursor[c] += counts[c];
    int running = 0;
    for (int w0 = 0; w0 < words_per_col; w0 += 32) {
      const int w = w0 + lane;
      const int cnt = w < words_per_col ? __popc(col[w].x) : 0;
      int incl = cnt;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += y;
      }
      if (w < words_per_col) col[w].y = (unsigned)(running + incl - cnt);
      running += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) counts[c] = running;
  }
  if (t_rowptr) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sm33[32] = (atomicInc(done, gridDim.x - 1) == gridDim.x - 1);
    __syncthreads();
    if (!sm33[32]) return;
    __threadfence();
    block_exclusive_scan<8>(counts, K, t_rowptr, sm33);
  }
}

// multi-block mode only: column totals of the whole matrix (integer atomics: order-free) and cursor = row pointer
__global__ void __launch_bounds__(256)
column_histogram_kernel(const int *__restrict__ colidx, int nnz, int *__restrict__ counts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(counts + __ldg(colidx + i), 1);
}

__global__ void copy_ints_kernel(const int *__restrict__ in, int n, int *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// workspace of gnn_csr_transpose: [counts K ints][cursor K ints][done 256 B][cells K x wpc uint2]
struct TrLayout { int wpc, nblocks; size_t off_counts, off_cursor, off_done, off_cells, total; };
inline TrLayout tr_layout(int64_t M, int64_t K) {
  TrLayout L;
  const int64_t words = cdiv(M, 32);
  const int64_t budget = g_transpose_budget.load(std::memory_order_relaxed);
  const int64_t fit = std::max<int64_t>(1, budget / (K * (int64_t)sizeof(uint2)));
  L.nblocks = (int)cdiv(words, std::min(words, fit));
  L.wpc = (int)cdiv(words, L.nblocks);               // equal blocks
  const size_t kal = ((size_t)K * 4 + 255) / 256 * 256;
  L.off_counts = 0; L.off_cursor = kal; L.off_done = 2 * kal; L.off_cells = 2 * kal + 256;
  L.total = L.off_cells + (size_t)K * L.wpc * sizeof(uint2) + 256;
  return L;
}

// exclusive scan of counts[0..n) into out[0..n], out[n] = total; single CTA of 1024 threads, tiles of 4096
// elements (4 consecutive per thread, coalesced), two barriers per tile.
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(const int *__restrict__ counts, int n, int *__restrict__ out) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {
    const int i0 = base + 4 * (int)threadIdx.x;
    int v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = i0 + q < n ? counts[i0 + q] : 0;
    const int sum = v[0] + v[1] + v[2] + v[3];
    int x = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(kFull, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    const int carry = carry_s;
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(kFull, w, off);
        if (lane >= off) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    int run = carry + x - sum + (warp ? warp_sums[warp - 1] : 0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (i0 + q < n) out[i0 + q] = run;
      run += v[q];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry_s;
}

// ---------------------------------------------------------------------------
// LADIES layer construction on the device (the array work of reference sampler.py:113-137; the weighted draw
// itself stays numpy on the host so the sampled node set is bit-identical to the reference)
// ---------------------------------------------------------------------------
__global__ void row_lengths_kernel(const int64_t *__restrict__ indptr, const int64_t *__restrict__ nodes, int M,
                                   int *__restrict__ lens) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) {
    const int64_t n = nodes[i];
    lens[i] = (int)(indptr[n + 1] - indptr[n]);
  }
}

// U = lap_matrix[nodes, :] (structure): copy every row's column ids; optionally count columns (ord=0 column norm).
// Entry-parallel: a CTA takes tiles of kSliceTile consecutive OUTPUT entries, whatever rows they belong to (LADIES draws
// the high-degree nodes: rows of 20 to 19,000 entries, and a warp per row left the launch waiting for its longest rows).
// The row of an entry is searched in fullrowptr between the rows of the tile's first and last entry (no step at all
// inside a long row).
constexpr int kSliceTile = 2048;     // entries per CTA tile (256 threads x 8)

// last row r in [lo, hi] with fullrowptr[r] <= e  (rows may be empty; e < fullrowptr[M])
__device__ __forceinline__ int row_of_entry(const int *__restrict__ fullrowptr, int lo, int hi, int e) {
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(fullrowptr + mid) <= e) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
row_slice_kernel(const int64_t *__restrict__ indptr, const int *__restrict__ indices, const int64_t *__restrict__ nodes, int M,
                 const int *__restrict__ fullrowptr, int *__restrict__ ucols, int *__restrict__ counts) {
  __shared__ int s_row[2];
  const int total = __ldg(fullrowptr + M);
  const int tiles = (total + kSliceTile - 1) / kSliceTile;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int t0 = tile * kSliceTile, t1 = min(t0 + kSliceTile, total);
    if (threadIdx.x < 2) s_row[threadIdx.x] = row_of_entry(fullrowptr, 0, M - 1, threadIdx.x == 0 ? t0 : t1 - 1);
    __syncthreads();
    const int r_lo = s_row[0], r_hi = s_row[1];
    int64_t src[kSliceTile / 256];
#pragma unroll
    for (int j = 0; j < kSliceTile / 256; ++j) {
      const int e = t0 + j * 256 + (int)threadIdx.x;
      src[j] = -1;
      if (e < t1) {
        const int r = row_of_entry(fullrowptr, r_lo, r_hi, e);
        src[j] = indptr[nodes[r]] + (e - __ldg(fullrowptr + r));
      }
    }
#pragma unroll
    for (int j = 0; j < kSliceTile / 256; ++j) {
      if (src[j] >= 0) {
        const int c = __ldg(indices + src[j]);
        ucols[t0 + j * 256 + (int)threadIdx.x] = c;
        if (counts) atomicAdd(counts + c, 1);
      }
    }
    __syncthreads();
  }
}

// Membership of the sampled columns: one bit per node id plus, for every 32-bit word that holds a bit, the position of
// its first member inside after_nodes (ascending, distinct: np.unique output).  The local column id of a kept entry is
// rank0[word] + popcount(bits below it) - 4 bytes of tables per 32 node ids (58 KB on the Reddit shape: L1-resident)
// instead of a 4-byte lookup word per node id gathered from L2 (one 32-byte sector per entry).
__global__ void member_set_kernel(unsigned *__restrict__ bits, int *__restrict__ rank0, const int64_t *__restrict__ after_nodes,
                                  int K, int set) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K) return;
  const int64_t v = after_nodes[j];
  if (!set) { bits[v >> 5] = 0u; return; }                                  // (all writers of a word store the same zero)
  atomicOr(bits + (v >> 5), 1u << (v & 31));
  if (j == 0 || (after_nodes[j - 1] >> 5) != (v >> 5)) rank0[v >> 5] = j;
}

__device__ __forceinline__ bool is_member(const unsigned *__restrict__ bits, int c) {
  return (__ldg(bits + (c >> 5)) >> (c & 31)) & 1u;
}

// adj = U[:, after_nodes] (structure): keep the entries whose column was sampled, renumbered to positions in after_nodes.
// The rows of U lie one after the other in ucols and the kept entries keep their order, so the column slice is an
// order-preserving stream compaction of the whole array; rowptr[r] is the number of kept entries before fullrowptr[r].
// Entry-parallel in chunks of kSliceChunk entries per warp (balanced whatever the row lengths):
//   column_chunk_count   kept entries per chunk              -> exclusive_scan_kernel -> chunk_prefix (its last entry: nnz)
//   column_rowptr_kernel rowptr[r] = chunk_prefix[chunk of fullrowptr[r]] + kept entries of that chunk before it
//   column_fill_kernel   colidx[chunk_prefix[c] + rank inside the chunk] = local column id (ascending entries = ascending
//                        positions: the same order as the reference's row-wise slice)
constexpr int kSliceChunk = 1024;    // entries per warp chunk (32 steps of 32)

__global__ void __launch_bounds__(256)
column_chunk_count_kernel(const int *__restrict__ ucols, int total, const unsigned *__restrict__ bits, int *__restrict__ chunk_cnt) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int chunks = (total + kSliceChunk - 1) / kSliceChunk;
  for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < chunks; c += gridDim.x * wpb) {
    const int base = c * kSliceChunk;
    int kept = 0;
#pragma unroll 4
    for (int g = 0; g < kSliceChunk / 256; ++g) {
      int col[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = base + g * 256 + q * 32 + lane;
        col[q] = i < total ? __ldg(ucols + i) : -1;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) kept += (col[q] >= 0 && is_member(bits, col[q])) ? 1 : 0;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) kept += __shfl_xor_sync(kFull, kept, off);
    if (lane == 0) chunk_cnt[c] = kept;
  }
}

// one warp per row r in [0, M]; chunk_prefix holds chunks + 1 entries
__global__ void __launch_bounds__(256)
column_rowptr_kernel(const int *__restrict__ ucols, int total, const int *__restrict__ fullrowptr, int M,
                     const unsigned *__restrict__ bits, const int *__restrict__ chunk_prefix, int *__restrict__ rowptr) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int chunks = (total + kSliceChunk - 1) / kSliceChunk;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r <= M; r += gridDim.x * wpb) {
    const int e0 = __ldg(fullrowptr + r);
    if (e0 >= total) {
      if (lane == 0) rowptr[r] = __ldg(chunk_prefix + chunks);
      continue;
    }
    const int c = e0 / kSliceChunk;
    int kept = 0;
    for (int i = c * kSliceChunk + lane; i < e0; i += 32) kept += is_member(bits, __ldg(ucols + i)) ? 1 : 0;
#pragma unroll
    for (int off = 16; off; off >>= 1) kept += __shfl_xor_sync(kFull, kept, off);
    if (lane == 0) rowptr[r] = __ldg(chunk_prefix + c) + kept;
  }
}

template <typename ColT>
__global__ void __launch_bounds__(256)
column_fill_kernel(const int *__restrict__ ucols, int total, const unsigned *__restrict__ bits, const int *__restrict__ rank0,
                   const int *__restrict__ chunk_prefix, ColT *__restrict__ colidx) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int chunks = (total + kSliceChunk - 1) / kSliceChunk;
  for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < chunks; c += gridDim.x * wpb) {
    const int base = c * kSliceChunk;
    int run = __ldg(chunk_prefix + c);
    for (int g = 0; g < kSliceChunk / 256; ++g) {
      if (base + g * 256 >= total) break;
      int local[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = base + g * 256 + q * 32 + lane;
        local[q] = i < total ? __ldg(ucols + i) : -1;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = local[q];
        if (c >= 0) {
          const unsigned w = __ldg(bits + (c >> 5)), b = 1u << (c & 31);
          local[q] = (w & b) ? __ldg(rank0 + (c >> 5)) + __popc(w & (b - 1u)) : -1;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned m = __ballot_sync(kFull, local[q] >= 0);
        if (local[q] >= 0) colidx[run + __popc(m & lt)] = (ColT)local[q];
        run += __popc(m);
      }
    }
  }
}

// Support of the layer's column counts (the nodes that carry probability, sampler.py:117/124) compacted on the device:
// (node id, count) pairs in ascending id order, written straight into pinned host memory (only the support crosses
// PCIe, and the host does not scan num_nodes counters).  Same chunk-count / scan / fill scheme as the column slice.
__global__ void __launch_bounds__(256)
support_chunk_count_kernel(const int *__restrict__ counts, int64_t n, int *__restrict__ chunk_cnt) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int64_t chunks = (n + kSliceChunk - 1) / kSliceChunk;
  for (int64_t c = blockIdx.x * wpb + (threadIdx.x >> 5); c < chunks; c += (int64_t)gridDim.x * wpb) {
    const int64_t base = c * kSliceChunk;
    int kept = 0;
#pragma unroll 8
    for (int q = 0; q < kSliceChunk / 32; ++q) {
      const int64_t i = base + q * 32 + lane;
      kept += (i < n && __ldg(counts + i) != 0) ? 1 : 0;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) kept += __shfl_xor_sync(kFull, kept, off);
    if (lane == 0) chunk_cnt[c] = kept;
  }
}

__global__ void __launch_bounds__(256)
support_fill_kernel(const int *__restrict__ counts, int64_t n, const int *__restrict__ chunk_prefix, int64_t *__restrict__ nz_out,
                    int *__restrict__ cnt_out, int64_t *__restrict__ n_support_out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t chunks = (n + kSliceChunk - 1) / kSliceChunk;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_support_out = __ldg(chunk_prefix + chunks);
  for (int64_t c = blockIdx.x * wpb + (threadIdx.x >> 5); c < chunks; c += (int64_t)gridDim.x * wpb) {
    const int64_t base = c * kSliceChunk;
    int run = __ldg(chunk_prefix + c);
    if (__ldg(chunk_prefix + c + 1) == run) continue;                        // nothing to write (sparse supports: most chunks)
    for (int g = 0; g < kSliceChunk / 256; ++g) {
      int v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int64_t i = base + g * 256 + q * 32 + lane;
        v[q] = i < n ? __ldg(counts + i) : 0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned m = __ballot_sync(kFull, v[q] != 0);
        if (v[q] != 0) {
          const int o = run + __popc(m & lt);
          nz_out[o] = base + g * 256 + q * 32 + lane;
          cnt_out[o] = v[q];
        }
        run += __popc(m);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Fused layer epilogue: y = rownorm(elu(x)) * scale + offset of reference
// models.py:21-25 / :61-64  (out = F.elu(feat); mean, var(unbiased=False)+1e-9; (out-mean)*scale*rsqrt(var)+offset).
// One warp per row; the row stays in registers between the ELU, the two reductions and the affine output.
// ---------------------------------------------------------------------------
constexpr int kEpiMaxPerLane = 64;   // columns per lane kept in registers => C <= 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

template <int PER>   // PER = ceil(C / 32) rounded up to a multiple of 4
__global__ void __launch_bounds__(256)
elu_rownorm_fwd_kernel(const float *__restrict__ x, int64_t ldx, int M, int C, const float *__restrict__ scale,
                       const float *__restrict__ offset, float *__restrict__ y, int64_t ldy, float *__restrict__ mean_out,
                       float *__restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < M; r += gridDim.x * wpb) {
    const float *xr = x + (int64_t)r * ldx;
    float o[PER];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = j * 32 + lane;
      float v = 0.f;
      if (c < C) {
        v = xr[c];
        v = v > 0.f ? v : expf(v) - 1.f;
        s += v;
      }
      o[j] = v;
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = j * 32 + lane;
      if (c < C) { const float d = o[j] - mean; q += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-9f);
    float *yr = y + (int64_t)r * ldy;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = j * 32 + lane;
      if (c < C) yr[c] = (o[j] - mean) * __ldg(scale + c) * rstd + __ldg(offset + c);
    }
    if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
  }
}

// dx for every row; per-CTA partial column sums of dscale / doffset (reduced in fixed order by the second kernel)
template <int PER>
__global__ void __launch_bounds__(256)
elu_rownorm_bwd_kernel(const float *__restrict__ dy, int64_t lddy, const float *__restrict__ x, int64_t ldx, int M, int C,
                       const float *__restrict__ scale, const float *__restrict__ mean_in, const float *__restrict__ rstd_in,
                       float *__restrict__ dx, int64_t lddx, float *__restrict__ part_scale, float *__restrict__ part_offset) {
  extern __shared__ float sm[];                     // [2][wpb][C] partial sums of this CTA's warps
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  float ds[PER], db[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) { ds[j] = 0.f; db[j] = 0.f; }
  for (int r = blockIdx.x * wpb + warp; r < M; r += gridDim.x * wpb) {
    const float *xr = x + (int64_t)r * ldx;
    const float *gr = dy + (int64_t)r * lddy;
    const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
    float xh[PER], g[PER], ep[PER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = j * 32 + lane;
      xh[j] = 0.f; g[j] = 0.f; ep[j] = 0.f;
      if (c < C) {
        const float v = xr[c];
        const float e = v > 0.f ? v : expf(v) - 1.f;
        ep[j] = v > 0.f ? 1.f : e + 1.f;            // d elu / dx
        xh[j] = (e - mean) * rstd;
        const float gy = gr[c];
        ds[j] += gy * xh[j];
        db[j] += gy;
        g[j] = gy * __ldg(scale + c);               // d loss / d xhat
        s1 += g[j];
        s2 += g[j] * xh[j];
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    float *dr = dx + (int64_t)r * lddx;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = j * 32 + lane;
      if (c < C) dr[c] = rstd * (g[j] - s1 - xh[j] * s2) * ep[j];
    }
  }
  // CTA-level reduction of the column partials in fixed warp order
  float *ss = sm, *sb = sm + (size_t)wpb * C;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int c = j * 32 + lane;
    if (c < C) { ss[(size_t)warp * C + c] = ds[j]; sb[(size_t)warp * C + c] = db[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < wpb; ++w) { a += ss[(size_t)w * C + c]; b += sb[(size_t)w * C + c]; }
    part_scale[(size_t)blockIdx.x * C + c] = a;
    part_offset[(size_t)blockIdx.x * C + c] = b;
  }
}

__global__ void column_partials_reduce_kernel(const float *__restrict__ part_scale, const float *__restrict__ part_offset,
                                              int nparts, int C, float *__restrict__ dscale, float *__restrict__ doffset) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int p = 0; p < nparts; ++p) { a += part_scale[(size_t)p * C + c]; b += part_offset[(size_t)p * C + c]; }
  dscale[c] = a;
  doffset[c] = b;
}

constexpr int kEpiCtas = 296;        // 2 per SM: row loop is grid-strided, partial buffers stay small

#include "epilogue_vec.cuh"

inline bool aligned16(const void *p, int64_t ld) { return ((reinterpret_cast<uintptr_t>(p) | (uintptr_t)(ld * 4)) & 15) == 0; }

template <int PER4>
int elu_rownorm_fwd_vec_launch(const float *x, int64_t ldx, int64_t M, int64_t C, const float *scale, const float *offset,
                               float *y, int64_t ldy, float *mean, float *rstd, cudaStream_t st) {
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(M, 8), 148 * 8);
  elu_rownorm_fwd_vec_kernel<PER4><<<grid, 256, 0, st>>>(x, ldx, (int)M, (int)C, scale, offset, y, ldy, mean, rstd);
  GNN_LAUNCH_CHECK();
  return 0;
}

template <int PER4>
int elu_rownorm_bwd_vec_launch(const float *dy, int64_t lddy, const float *x, int64_t ldx, int64_t M, int64_t C,
                               const float *scale, const float *mean, const float *rstd, float *dx, int64_t lddx,
                               float *dscale, float *doffset, float *ws, cudaStream_t st) {
  const int wpb = 8;
  const size_t smem = (size_t)2 * wpb * C * sizeof(float);               // <= 64 KB for C <= 1024 (two CTAs per SM)
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(M, wpb), kEpiCtas);
  if (smem > 48 * 1024)
    GNN_CUDA(cudaFuncSetAttribute(elu_rownorm_bwd_vec_kernel<PER4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float *ps = ws, *po = ws + (size_t)kEpiCtas * C;
  elu_rownorm_bwd_vec_kernel<PER4><<<grid, wpb * 32, smem, st>>>(dy, lddy, x, ldx, (int)M, (int)C, scale, mean, rstd, dx, lddx, ps, po);
  GNN_LAUNCH_CHECK();
  column_partials_reduce_kernel<<<(unsigned)cdiv(C, 256), 256, 0, st>>>(ps, po, (int)grid, (int)C, dscale, doffset);
  GNN_LAUNCH_CHECK();
  return 0;
}

template <int PER>
int elu_rownorm_fwd_launch(const float *x, int64_t ldx, int64_t M, int64_t C, const float *scale, const float *offset,
                                  float *y, int64_t ldy, float *mean, float *rstd, cudaStream_t st) {
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(M, 8), 148 * 8);
  elu_rownorm_fwd_kernel<PER><<<grid, 256, 0, st>>>(x, ldx, (int)M, (int)C, scale, offset, y, ldy, mean, rstd);
  GNN_LAUNCH_CHECK();
  return 0;
}

template <int PER>
int elu_rownorm_bwd_launch(const float *dy, int64_t lddy, const float *x, int64_t ldx, int64_t M, int64_t C,
                                  const float *scale, const float *mean, const float *rstd, float *dx, int64_t lddx,
                                  float *dscale, float *doffset, float *ws, cudaStream_t st) {
  // warps per CTA limited by the 2 x wpb x C floats of shared memory (<= 96 KB)
  int wpb = 8;
  while (wpb > 1 && (size_t)2 * wpb * C * sizeof(float) > 96 * 1024) wpb >>= 1;
  const size_t smem = (size_t)2 * wpb * C * sizeof(float);
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(M, wpb), kEpiCtas);
  if (smem > 48 * 1024) GNN_CUDA(cudaFuncSetAttribute(elu_rownorm_bwd_kernel<PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float *ps = ws, *po = ws + (size_t)kEpiCtas * C;
  elu_rownorm_bwd_kernel<PER><<<grid, wpb * 32, smem, st>>>(dy, lddy, x, ldx, (int)M, (int)C, scale, mean, rstd, dx, lddx, ps, po);
  GNN_LAUNCH_CHECK();
  column_partials_reduce_kernel<<<(unsigned)cdiv(C, 256), 256, 0, st>>>(ps, po, (int)grid, (int)C, dscale, doffset);
  GNN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// placement remap + gathers
// ---------------------------------------------------------------------------
__global__ void placement_remap_kernel(const int64_t *__restrict__ input_nodes, int64_t n0,
                                       const int64_t *__restrict__ dev_of, const int64_t *__restrict__ idx_of,
                                       const int64_t *__restrict__ devices, int world, const float *const *__restrict__ bases,
                                       int64_t ld_src, int64_t ld_host, int *__restrict__ src_dev, int64_t *__restrict__ slot,
                                       const float **__restrict__ xrows, unsigned long long *__restrict__ counts) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n0) return;
  const int64_t node = input_nodes[j];
  const int64_t dev = dev_of[node];
  int s = -2;
  int64_t sl = -1;
  if (dev == -1) {
    s = -1; sl = node;
  } else {
    for (int i = 0; i < world; ++i)
      if (devices[i] == dev) { s = i; sl = idx_of[node]; }
  }
  src_dev[j] = s;
  slot[j] = sl;
  if (xrows) {
    // a source whose base pointer is NULL (no mapped host table, a shard that was not opened) yields a NULL row
    // pointer, which the gather kernels skip - never an address computed from NULL
    const float *base = (s == -2 || !bases) ? nullptr : bases[s < 0 ? world : s];
    xrows[j] = base ? base + sl * (s < 0 ? ld_host : ld_src) : nullptr;
  }
  if (counts) atomicAdd(counts + (s == -2 ? world + 1 : (s < 0 ? world : s)), 1ull);
}

// one warp per row; each lane keeps up to 4 independent 128-bit loads in flight
template <bool FILTER, bool INDEX>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float *const *__restrict__ xrows, const int *__restrict__ src_dev, int only_src,
                   const float *__restrict__ X, int64_t ldx, const int64_t *__restrict__ idx, int64_t n0, int F,
                   float *__restrict__ out, int64_t ld_out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t j = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); j < n0; j += (int64_t)gridDim.x * wpb) {
    if (FILTER) {
      const int sd = src_dev[j];
      const bool take = only_src == GNN_SRC_DEVICES ? sd >= 0
                        : (only_src <= GNN_SRC_PEERS(0) ? (sd >= 0 && sd != -(only_src + 100000))
                        : (only_src <= GNN_SRC_NOT(0) ? (sd != GNN_SRC_NOT(only_src) && sd != -2) : sd == only_src));
      if (!take) continue;
    }
    const float *src = INDEX ? X + idx[j] * ldx : xrows[j];
    if (!src) continue;
    float *dst = out + j * ld_out;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
      const int nv = F >> 2;
      const float4 *s4 = reinterpret_cast<const float4 *>(src);
      float4 *d4 = reinterpret_cast<float4 *>(dst);
      int v = lane;
      for (; v + 96 < nv; v += 128) {
        const float4 a = s4[v], b = s4[v + 32], c = s4[v + 64], d = s4[v + 96];
        d4[v] = a; d4[v + 32] = b; d4[v + 64] = c; d4[v + 96] = d;
      }
      for (; v < nv; v += 32) d4[v] = s4[v];
      for (int t = (nv << 2) + lane; t < F; t += 32) dst[t] = src[t];
    } else if ((((uintptr_t)src | (uintptr_t)dst) & 7) == 0) {
      const int nv = F >> 1;
      const float2 *s2 = reinterpret_cast<const float2 *>(src);
      float2 *d2 = reinterpret_cast<float2 *>(dst);
      int v = lane;
      for (; v + 96 < nv; v += 128) {
        const float2 a = s2[v], b = s2[v + 32], c = s2[v + 64], d = s2[v + 96];
        d2[v] = a; d2[v + 32] = b; d2[v + 64] = c; d2[v + 96] = d;
      }
      for (; v < nv; v += 32) d2[v] = s2[v];
      for (int t = (nv << 1) + lane; t < F; t += 32) dst[t] = src[t];
    } else {
      for (int t = lane; t < F; t += 32) dst[t] = src[t];
    }
  }
}

inline unsigned warp_grid(int64_t n_warps, int wpb) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n_warps, wpb), 148 * 32));
}

#ifdef GNN_TUNE
#include "experiments_device.cuh"   // gather-roof microbenchmark, hub-cache prototype (tools/ only)
#endif
#include "linear_tc.cuh"            // tcgen05 3xTF32 dense linears of a layer

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int gnn_abi_version(void) { return GNN_B200_ABI_VERSION; }

int gnn_set_corunner_ctas(int ctas) {
  if (ctas < 0) return GNN_E_BADARG;
  return g_corunner_ctas.exchange(ctas, std::memory_order_relaxed);
}

int gnn_host_gather_ctas(void) { return kHostGatherCtas; }

const char *gnn_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case GNN_E_BADARG: return "gnn_b200: bad argument (null pointer, negative size, unsupported width)";
    case GNN_E_WORKSPACE: return "gnn_b200: workspace missing or too small";
    case GNN_E_RANGE: return "gnn_b200: size exceeds the limits of this path";
    case GNN_E_DRIVER: return "gnn_b200: cuTensorMapEncodeTiled unavailable or rejected the weight descriptor";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "gnn_b200: unknown error";
  }
}

int64_t gnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int gnn_build_adj(const int32_t *fullrowptr, const int32_t *rowptr, const void *colidx, int colidx_bytes,
                  const float *normfact, int64_t M, int64_t K, int64_t nnz, int64_t *out_indices, float *out_vals,
                  int32_t *out_colidx32, int32_t *out_rowidx32, gnn_stream_t stream) {
  (void)K;
  if (M < 0 || nnz < 0 || (colidx_bytes != 2 && colidx_bytes != 4)) return GNN_E_BADARG;
  if (M >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 1) return GNN_E_RANGE;
  if (M == 0 || nnz == 0) return 0;
  if (!fullrowptr || !rowptr || !colidx || !normfact || !out_vals) return GNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)cdiv(cdiv(nnz, kAdjChunk), 8);
  if (colidx_bytes == 2)
    build_adj_kernel<int16_t><<<grid, 256, 0, st>>>(fullrowptr, rowptr, (const int16_t *)colidx, normfact, (int)M, nnz,
                                                    out_indices, out_vals, out_colidx32, out_rowidx32);
  else
    build_adj_kernel<int32_t><<<grid, 256, 0, st>>>(fullrowptr, rowptr, (const int32_t *)colidx, normfact, (int)M, nnz,
                                                    out_indices, out_vals, out_colidx32, out_rowidx32);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_coo_to_csr(const int64_t *indices, int64_t M, int64_t nnz, int32_t *out_rowptr, int32_t *out_colidx32,
                   gnn_stream_t stream) {
  if (M < 0 || nnz < 0 || !out_rowptr) return GNN_E_BADARG;
  if (M >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 1) return GNN_E_RANGE;
  if (nnz > 0 && (!indices || !out_colidx32)) return GNN_E_BADARG;
  coo_to_csr_kernel<<<(unsigned)cdiv(nnz + 1, 256), 256, 0, (cudaStream_t)stream>>>(indices, M, nnz, out_rowptr, out_colidx32);
  GNN_LAUNCH_CHECK();
  return 0;
}

size_t gnn_csr_spmm_counter_bytes(int64_t M, int64_t nnz, int64_t D) {
  if (M <= 0 || nnz <= 0 || D <= 0) return 256;
  return spmm_counter_bytes(M, nnz, D);
}

size_t gnn_csr_spmm_partial_bytes(int64_t M, int64_t nnz, int64_t D) {
  if (M <= 0 || nnz <= 0 || D <= 0) return 256;
  const int C = flat_wanted(M, nnz, D) ? flat_chunk(nnz, D)
                                       : std::max(32, spmm_chunk(nnz, D) / 2);     // wave fitting may halve the base chunk
  const size_t nchunks = (size_t)cdiv(nnz, C);
  const size_t Dp = (size_t)cdiv(D, 4) * 4;
  return 2 * nchunks * Dp * sizeof(float) + 256;
}

size_t gnn_csr_spmm_workspace_bytes(int64_t M, int64_t nnz, int64_t D) {
  return gnn_csr_spmm_counter_bytes(M, nnz, D) + gnn_csr_spmm_partial_bytes(M, nnz, D);
}

// workspace of the plain entry points = [counters | partials]; size and null checks happen in spmm_entry, after the
// argument and range checks
#define GNN_SPLIT_WS()                                                                              \
  const size_t cb_ = gnn_csr_spmm_counter_bytes(M, nnz, D);                                          \
  int *counters_ = workspace_bytes >= cb_ ? reinterpret_cast<int *>(workspace) : nullptr;           \
  void *partials_ = (workspace && workspace_bytes >= cb_) ? reinterpret_cast<char *>(workspace) + cb_ : nullptr; \
  const size_t pb_ = workspace_bytes > cb_ ? workspace_bytes - cb_ : 0

int gnn_csr_spmm_f32(const int32_t *rowptr, const int32_t *colidx, const float *vals, int64_t M, int64_t K, int64_t nnz,
                     int64_t D, const float *X, int64_t ldx, float *Y, int64_t ldy, void *workspace,
                     size_t workspace_bytes, gnn_stream_t stream) {
  GNN_SPLIT_WS();
  return spmm_entry<false>(rowptr, nullptr, colidx, vals, M, K, nnz, D, X, ldx, nullptr, Y, ldy, counters_, partials_, pb_, 0u,
                           (cudaStream_t)stream);
}

int gnn_gather_spmm_f32(const int32_t *rowptr, const int32_t *colidx, const float *vals, int64_t M, int64_t K,
                        int64_t nnz, int64_t D, const float *const *xrows, float *Y, int64_t ldy, void *workspace,
                        size_t workspace_bytes, gnn_stream_t stream) {
  GNN_SPLIT_WS();
  return spmm_entry<true>(rowptr, nullptr, colidx, vals, M, K, nnz, D, nullptr, 0, xrows, Y, ldy, counters_, partials_, pb_, 0u,
                          (cudaStream_t)stream);
}

int gnn_csr_spmm_f32_ex(const int32_t *rowptr, const int32_t *rowidx, const int32_t *colidx, const float *vals, int64_t M,
                        int64_t K, int64_t nnz, int64_t D, const float *X, int64_t ldx, float *Y, int64_t ldy,
                        int32_t *counters, void *partials, size_t partial_bytes, unsigned flags, gnn_stream_t stream) {
  return spmm_entry<false>(rowptr, rowidx, colidx, vals, M, K, nnz, D, X, ldx, nullptr, Y, ldy, counters, partials, partial_bytes,
                           flags, (cudaStream_t)stream);
}

int gnn_gather_spmm_f32_ex(const int32_t *rowptr, const int32_t *rowidx, const int32_t *colidx, const float *vals, int64_t M,
                           int64_t K, int64_t nnz, int64_t D, const float *const *xrows, float *Y, int64_t ldy,
                           int32_t *counters, void *partials, size_t partial_bytes, unsigned flags, gnn_stream_t stream) {
  return spmm_entry<true>(rowptr, rowidx, colidx, vals, M, K, nnz, D, nullptr, 0, xrows, Y, ldy, counters, partials, partial_bytes,
                          flags, (cudaStream_t)stream);
}

int gnn_csr_spmm_t_f32(const int32_t *rowptr, const int32_t *rowidx, const int32_t *colidx, const float *vals, int64_t M,
                       int64_t K, int64_t nnz, int64_t D, const float *G, int64_t ldg, float *dX, int64_t lddx,
                       gnn_stream_t stream) {
  if (M < 0 || K < 0 || nnz < 0 || D < 0) return GNN_E_BADARG;
  if (K == 0 || D == 0) return 0;
  if (!dX || lddx < D) return GNN_E_BADARG;
  if (M >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 2048 || K >= (1ll << 31) || D >= (1ll << 24)) return GNN_E_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (lddx == D) {
    GNN_CUDA(cudaMemsetAsync(dX, 0, (size_t)K * D * sizeof(float), st));
  } else {
    zero_rows_kernel<<<(unsigned)std::min<int64_t>(cdiv(K * D, 256), 148 * 16), 256, 0, st>>>(dX, lddx, K, D);
    GNN_LAUNCH_CHECK();
  }
  if (nnz == 0 || M == 0) return 0;
  if (!rowptr || !colidx || !vals || !G || ldg < D) return GNN_E_BADARG;
  ScatterParams p;
  p.rowptr = rowptr; p.rowidx = rowidx; p.colidx = colidx; p.vals = vals;
  p.M = (int)M; p.nnz = (int)nnz; p.D = (int)D;
  p.G = G; p.ldg = ldg; p.dX = dX; p.lddx = lddx;
  const bool vec4 = ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(dX) | (uintptr_t)(ldg * 4) | (uintptr_t)(lddx * 4)) & 15) == 0;
  const int64_t nvec = cdiv(D, vec4 ? 4 : 1);
  const int64_t n = cdiv(nvec, 32);
  // chunk: as for the flat SpMM - enough warp items for ~2 waves, at most 128 entries each
  int C = kFlatMaxC;
  while (C > 32 && cdiv(nnz, C) * n < kFlatTargetItems) C >>= 1;
  int nv = 1;
  if (n >= 2 && cdiv(nnz, C) * cdiv(n, 2) >= 2 * kFlatTargetItems) nv = 2;
#ifdef GNN_TUNE
  if (getenv("GNN_TUNE_SC")) C = atoi(getenv("GNN_TUNE_SC"));
  if (getenv("GNN_TUNE_SNV")) nv = atoi(getenv("GNN_TUNE_SNV"));
#endif
  p.C = C;
  p.nchunks = (int)cdiv(nnz, C);
  p.nslabs = (int)cdiv(n, nv);
  const unsigned grid = (unsigned)cdiv((int64_t)p.nchunks * p.nslabs, kFlatWarps);
#define GNN_SCATTER(V4_, NV_, U_)                                                                             \
  do {                                                                                                        \
    if (rowidx) spmm_scatter_t_kernel<V4_, NV_, U_, true><<<grid, kFlatWarps * 32, 0, st>>>(p);                 \
    else spmm_scatter_t_kernel<V4_, NV_, U_, false><<<grid, kFlatWarps * 32, 0, st>>>(p);                       \
  } while (0)
  if (vec4) {
    if (nv == 2) GNN_SCATTER(true, 2, 4); else GNN_SCATTER(true, 1, 8);
  } else {
    if (nv == 2) GNN_SCATTER(false, 2, 8); else GNN_SCATTER(false, 1, 8);
  }
#undef GNN_SCATTER
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_probe_row_gather_f32(const float *X, int64_t ldx, int64_t D, const int32_t *colidx, int64_t nnz, int nv,
                             int warps_per_sm, float *sink, int64_t *bytes_gathered, gnn_stream_t stream) {
  if (!X || !colidx || !sink || nnz <= 0 || nnz >= (1ll << 31) || warps_per_sm <= 0) return GNN_E_BADARG;
  if ((nv != 1 && nv != 2 && nv != 4) || D < 128 * nv || ldx < D) return GNN_E_BADARG;
  if (((reinterpret_cast<uintptr_t>(X) | (uintptr_t)(ldx * 4)) & 15) != 0) return GNN_E_BADARG;
  const int nslabs = (int)(D / (128 * nv));
  const int64_t warps = (int64_t)device_sm_count() * warps_per_sm / nslabs * nslabs;
  if (warps <= 0) return GNN_E_BADARG;
  const int per_warp = (int)std::max<int64_t>(32, nnz * nslabs / warps / 32 * 32);
  const unsigned grid = (unsigned)cdiv(warps, 8);
  cudaStream_t st = (cudaStream_t)stream;
  // the grid is rounded up to whole CTAs: count what the launch really moves
  const int64_t launched = (int64_t)grid * 8;
  if (bytes_gathered) *bytes_gathered = launched * per_warp * nv * 512;
  switch (nv) {
    case 1: row_gather_probe_kernel<1, 8><<<grid, 256, 0, st>>>(X, ldx, colidx, (int)nnz, per_warp, 128, nslabs, sink); break;
    case 2: row_gather_probe_kernel<2, 4><<<grid, 256, 0, st>>>(X, ldx, colidx, (int)nnz, per_warp, 256, nslabs, sink); break;
    default: row_gather_probe_kernel<4, 2><<<grid, 256, 0, st>>>(X, ldx, colidx, (int)nnz, per_warp, 512, nslabs, sink); break;
  }
  GNN_LAUNCH_CHECK();
  return 0;
}

int64_t gnn_set_transpose_budget(int64_t bytes) {
  if (bytes < 0) return GNN_E_BADARG;
  return g_transpose_budget.exchange(bytes == 0 ? ((int64_t)512 << 20) : bytes, std::memory_order_relaxed);
}

size_t gnn_csr_transpose_workspace_bytes(int64_t M, int64_t K, int64_t nnz) {
  (void)nnz;
  if (M <= 0 || K <= 0) return 256;
  return tr_layout(M, K).total;
}

int gnn_csr_transpose(const int32_t *rowptr, const int32_t *colidx, const float *vals, int64_t M, int64_t K, int64_t nnz,
                      int32_t *t_rowptr, int32_t *t_colidx, float *t_vals, int32_t *t_rowidx, void *workspace,
                      size_t workspace_bytes, gnn_stream_t stream) {
  if (M < 0 || K < 0 || nnz < 0 || !t_rowptr) return GNN_E_BADARG;
  if (M >= (1ll << 31) - 1 || K >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 1) return GNN_E_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 0) { GNN_CUDA(cudaMemsetAsync(t_rowptr, 0, sizeof(int), st)); return 0; }
  if (nnz == 0 || M == 0) { GNN_CUDA(cudaMemsetAsync(t_rowptr, 0, (size_t)(K + 1) * sizeof(int), st)); return 0; }
  if (!rowptr || !colidx || !vals || !t_colidx || !t_vals) return GNN_E_BADARG;
  const TrLayout L = tr_layout(M, K);
  if (!workspace || workspace_bytes < L.total) return GNN_E_WORKSPACE;
  char *ws = reinterpret_cast<char *>(workspace);
  int *counts = reinterpret_cast<int *>(ws + L.off_counts);
  int *cursor = reinterpret_cast<int *>(ws + L.off_cursor);
  unsigned *done = reinterpret_cast<unsigned *>(ws + L.off_done);
  uint2 *cells = reinterpret_cast<uint2 *>(ws + L.off_cells);
  const size_t cell_bytes = (size_t)K * L.wpc * sizeof(uint2);
  const unsigned pass_grid = (unsigned)cdiv(cdiv(nnz, kTrChunk), 8);
  const unsigned prefix_grid = warp_grid(K, 8);          // one warp per column, 256-thread CTAs
  if (L.nblocks == 1) {
    // `done` sits right in front of the cells: one memset zeroes both
    GNN_CUDA(cudaMemsetAsync(done, 0, (L.off_cells - L.off_done) + cell_bytes, st));
    transpose_pass_kernel<false><<<pass_grid, 256, 0, st>>>(cells, (int)M, (int)nnz, L.wpc, 0, (int)M, rowptr, colidx, vals,
                                                           nullptr, t_colidx, t_vals, nullptr);
    GNN_LAUNCH_CHECK();
    bitmap_prefix_kernel<<<prefix_grid, 256, 0, st>>>(cells, (int)K, L.wpc, counts, nullptr, 0, t_rowptr, done);
    GNN_LAUNCH_CHECK();
    transpose_pass_kernel<true><<<pass_grid, 256, 0, st>>>(cells, (int)M, (int)nnz, L.wpc, 0, (int)M, rowptr, colidx, vals,
                                                          t_rowptr, t_colidx, t_vals, t_rowidx);
    GNN_LAUNCH_CHECK();
    return 0;
  }
  // row-blocked: column totals first, then block after block with a per-column cursor
  GNN_CUDA(cudaMemsetAsync(counts, 0, (size_t)K * sizeof(int), st));
  column_histogram_kernel<<<(unsigned)std::min<int64_t>(cdiv(nnz, 256), 148 * 16), 256, 0, st>>>(colidx, (int)nnz, counts);
  GNN_LAUNCH_CHECK();
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, (int)K, t_rowptr);
  GNN_LAUNCH_CHECK();
  copy_ints_kernel<<<(unsigned)cdiv(K, 256), 256, 0, st>>>(t_rowptr, (int)K, cursor);
  GNN_LAUNCH_CHECK();
  for (int blk = 0; blk < L.nblocks; ++blk) {
    const int row0 = blk * L.wpc * 32;
    const int row1 = (int)std::min<int64_t>(M, (int64_t)row0 + (int64_t)L.wpc * 32);
    GNN_CUDA(cudaMemsetAsync(cells, 0, cell_bytes, st));
    transpose_pass_kernel<false><<<pass_grid, 256, 0, st>>>(cells, (int)M, (int)nnz, L.wpc, row0, row1, rowptr, colidx, vals,
                                                           nullptr, t_colidx, t_vals, nullptr);
    GNN_LAUNCH_CHECK();
    bitmap_prefix_kernel<<<prefix_grid, 256, 0, st>>>(cells, (int)K, L.wpc, counts, cursor, blk > 0 ? 1 : 0, nullptr, done);
    GNN_LAUNCH_CHECK();
    transpose_pass_kernel<true><<<pass_grid, 256, 0, st>>>(cells, (int)M, (int)nnz, L.wpc, row0, row1, rowptr, colidx, vals,
                                                          cursor, t_colidx, t_vals, t_rowidx);
    GNN_LAUNCH_CHECK();
  }
  return 0;
}

int gnn_placement_remap(const int64_t *input_nodes, int64_t n0, const int64_t *device_id_of_nodes,
                        const int64_t *idx_of_nodes_on_device, const int64_t *devices, int64_t world,
                        const float *const *bases, int64_t ld_src, int64_t ld_host, int32_t *src_dev, int64_t *slot,
                        const float **xrows, int64_t *counts, gnn_stream_t stream) {
  if (n0 < 0 || world < 0 || world > 1024) return GNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (counts) GNN_CUDA(cudaMemsetAsync(counts, 0, (size_t)(world + 2) * sizeof(int64_t), st));
  if (n0 == 0) return 0;
  if (!input_nodes || !device_id_of_nodes || !idx_of_nodes_on_device || (world > 0 && !devices) || !src_dev || !slot)
    return GNN_E_BADARG;
  if (xrows && !bases) return GNN_E_BADARG;
  placement_remap_kernel<<<(unsigned)cdiv(n0, 256), 256, 0, st>>>(input_nodes, n0, device_id_of_nodes,
                                                                 idx_of_nodes_on_device, devices, (int)world, bases, ld_src,
                                                                 ld_host, src_dev, slot, xrows, (unsigned long long *)counts);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_gather_rows_f32(const float *const *xrows, int64_t n0, int64_t F, float *out, int64_t ld_out, gnn_stream_t stream) {
  if (n0 < 0 || F < 0 || F >= (1ll << 31)) return GNN_E_BADARG;
  if (n0 == 0 || F == 0) return 0;
  if (!xrows || !out || ld_out < F) return GNN_E_BADARG;
  gather_rows_kernel<false, false><<<warp_grid(n0, 8), 256, 0, (cudaStream_t)stream>>>(xrows, nullptr, 0, nullptr, 0, nullptr,
                                                                                      n0, (int)F, out, ld_out);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_gather_rows_src_f32(const float *const *xrows, const int32_t *src_dev, int32_t only_src, int64_t n0, int64_t F,
                            float *out, int64_t ld_out, gnn_stream_t stream) {
  if (n0 < 0 || F < 0 || F >= (1ll << 31)) return GNN_E_BADARG;
  if (n0 == 0 || F == 0) return 0;
  if (!xrows || !src_dev || !out || ld_out < F) return GNN_E_BADARG;
  // host rows (only_src == -1) are PCIe-bound and normally run NEXT TO the current step's SpMMs (prefetch of the
  // next minibatch).  The grid is capped for the co-runners' sake, not the link's: every kernel launch fetches its
  // commands over the same PCIe link, behind whatever zero-copy reads are queued - 48 CTAs (47.6 GB/s) add 16 us to
  // each launch on every other stream, 24 CTAs (45.4 GB/s) 5 us, 16 CTAs (38.8 GB/s) 2.5 us, <= 8 nothing
  // (measured on B200: 100 tiny kernels beside a looping host gather).
  unsigned grid = warp_grid(n0, 8);
  if (only_src == -1) grid = std::min(grid, (unsigned)kHostGatherCtas);
  gather_rows_kernel<true, false><<<grid, 256, 0, (cudaStream_t)stream>>>(xrows, src_dev, only_src, nullptr, 0, nullptr, n0,
                                                                         (int)F, out, ld_out);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_index_rows_f32(const float *X, int64_t ldx, const int64_t *idx, int64_t n, int64_t F, float *out, int64_t ld_out,
                       gnn_stream_t stream) {
  if (n < 0 || F < 0 || F >= (1ll << 31)) return GNN_E_BADARG;
  if (n == 0 || F == 0) return 0;
  if (!X || !idx || !out || ld_out < F || ldx < F) return GNN_E_BADARG;
  gather_rows_kernel<false, true><<<warp_grid(n, 8), 256, 0, (cudaStream_t)stream>>>(nullptr, nullptr, 0, X, ldx, idx, n,
                                                                                    (int)F, out, ld_out);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_row_slice_count(const int64_t *indptr, const int64_t *nodes, int64_t M, int32_t *scratch_lens,
                        int32_t *out_fullrowptr, gnn_stream_t stream) {
  if (M < 0) return GNN_E_BADARG;
  if (M >= (1ll << 31) - 1) return GNN_E_RANGE;
  if (!out_fullrowptr || (M > 0 && (!indptr || !nodes || !scratch_lens))) return GNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) { GNN_CUDA(cudaMemsetAsync(out_fullrowptr, 0, sizeof(int), st)); return 0; }
  row_lengths_kernel<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(indptr, nodes, (int)M, scratch_lens);
  GNN_LAUNCH_CHECK();
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(scratch_lens, (int)M, out_fullrowptr);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_row_slice_fill(const int64_t *indptr, const int32_t *indices, const int64_t *nodes, int64_t M,
                       const int32_t *fullrowptr, int32_t *out_cols, int32_t *col_counts, gnn_stream_t stream) {
  if (M < 0) return GNN_E_BADARG;
  if (M == 0) return 0;
  if (!indptr || !indices || !nodes || !fullrowptr || !out_cols) return GNN_E_BADARG;
  // the entry count (fullrowptr[M]) lives on the device: a fixed grid walks the tiles
  row_slice_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(indptr, indices, nodes, (int)M, fullrowptr, out_cols, col_counts);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_member_set(uint32_t *bits, int32_t *rank0, const int64_t *after_nodes, int64_t K, int set, gnn_stream_t stream) {
  if (K < 0 || K >= (1ll << 31)) return GNN_E_BADARG;
  if (K == 0) return 0;
  if (!bits || !rank0 || !after_nodes) return GNN_E_BADARG;
  member_set_kernel<<<(unsigned)cdiv(K, 256), 256, 0, (cudaStream_t)stream>>>(bits, rank0, after_nodes, (int)K, set);
  GNN_LAUNCH_CHECK();
  return 0;
}

int64_t gnn_column_slice_chunks(int64_t total) { return total > 0 ? cdiv(total, kSliceChunk) : 0; }

int gnn_column_slice_count(const int32_t *ucols, int64_t total, const int32_t *fullrowptr, int64_t M, const uint32_t *bits,
                           int32_t *chunk_prefix, int32_t *out_rowptr, gnn_stream_t stream) {
  if (M < 0 || total < 0) return GNN_E_BADARG;
  if (total >= (1ll << 31) - kSliceChunk || M >= (1ll << 31) - 1) return GNN_E_RANGE;
  if (!out_rowptr || !chunk_prefix || (M > 0 && !fullrowptr) || (total > 0 && (!ucols || !bits))) return GNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chunks = gnn_column_slice_chunks(total);
  if (chunks == 0) {
    GNN_CUDA(cudaMemsetAsync(chunk_prefix, 0, sizeof(int), st));
    GNN_CUDA(cudaMemsetAsync(out_rowptr, 0, (size_t)(M + 1) * sizeof(int), st));
    return 0;
  }
  int32_t *chunk_cnt = chunk_prefix + chunks + 1;                                           // second half of the scratch
  column_chunk_count_kernel<<<warp_grid(chunks, 8), 256, 0, st>>>(ucols, (int)total, bits, chunk_cnt);
  GNN_LAUNCH_CHECK();
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(chunk_cnt, (int)chunks, chunk_prefix);
  GNN_LAUNCH_CHECK();
  column_rowptr_kernel<<<warp_grid(M + 1, 8), 256, 0, st>>>(ucols, (int)total, fullrowptr, (int)M, bits, chunk_prefix, out_rowptr);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_column_slice_fill(const int32_t *ucols, int64_t total, const uint32_t *bits, const int32_t *rank0, const int32_t *chunk_prefix,
                          void *out_colidx, int colidx_bytes, gnn_stream_t stream) {
  if (total < 0 || (colidx_bytes != 2 && colidx_bytes != 4)) return GNN_E_BADARG;
  if (total >= (1ll << 31) - kSliceChunk) return GNN_E_RANGE;
  if (total == 0) return 0;
  if (!ucols || !bits || !rank0 || !chunk_prefix || !out_colidx) return GNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = warp_grid(gnn_column_slice_chunks(total), 8);
  if (colidx_bytes == 2)
    column_fill_kernel<int16_t><<<grid, 256, 0, st>>>(ucols, (int)total, bits, rank0, chunk_prefix, (int16_t *)out_colidx);
  else
    column_fill_kernel<int><<<grid, 256, 0, st>>>(ucols, (int)total, bits, rank0, chunk_prefix, (int *)out_colidx);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_support_compact(const int32_t *counts, int64_t num_nodes, int32_t *chunk_scratch, int64_t *nz_out, int32_t *cnt_out,
                        int64_t *n_support_out, gnn_stream_t stream) {
  if (num_nodes <= 0 || !counts || !chunk_scratch || !nz_out || !cnt_out || !n_support_out) return GNN_E_BADARG;
  const int64_t chunks = cdiv(num_nodes, kSliceChunk);
  if (chunks >= (1ll << 31) - 2) return GNN_E_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t *chunk_cnt = chunk_scratch + chunks + 1;
  support_chunk_count_kernel<<<warp_grid(chunks, 8), 256, 0, st>>>(counts, num_nodes, chunk_cnt);
  GNN_LAUNCH_CHECK();
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(chunk_cnt, (int)chunks, chunk_scratch);
  GNN_LAUNCH_CHECK();
  support_fill_kernel<<<warp_grid(chunks, 8), 256, 0, st>>>(counts, num_nodes, chunk_scratch, nz_out, cnt_out, n_support_out);
  GNN_LAUNCH_CHECK();
  return 0;
}

size_t gnn_elu_rownorm_workspace_bytes(int64_t C) { return (size_t)2 * kEpiCtas * (size_t)(C > 0 ? C : 1) * sizeof(float); }

int gnn_elu_rownorm_fwd_f32(const float *x, int64_t ldx, int64_t M, int64_t C, const float *scale, const float *offset,
                            float *y, int64_t ldy, float *mean, float *rstd, gnn_stream_t stream) {
  if (M < 0 || C < 0) return GNN_E_BADARG;
  if (M == 0 || C == 0) return 0;
  if (C > 32 * kEpiMaxPerLane || M >= (1ll << 31)) return GNN_E_RANGE;
  if (!x || !scale || !offset || !y || !mean || !rstd || ldx < C || ldy < C) return GNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int per = (int)cdiv(C, 32);
  if (C % 4 == 0 && C <= 1024 && aligned16(x, ldx) && aligned16(y, ldy) && aligned16(scale, 4) && aligned16(offset, 4)) {
    const int per4 = (int)cdiv(C, 128);
    if (per4 <= 1) return elu_rownorm_fwd_vec_launch<1>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
    if (per4 <= 2) return elu_rownorm_fwd_vec_launch<2>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
    if (per4 <= 4) return elu_rownorm_fwd_vec_launch<4>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
    return elu_rownorm_fwd_vec_launch<8>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
  }
  if (per <= 4) return elu_rownorm_fwd_launch<4>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
  if (per <= 8) return elu_rownorm_fwd_launch<8>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
  if (per <= 16) return elu_rownorm_fwd_launch<16>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
  if (per <= 32) return elu_rownorm_fwd_launch<32>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
  return elu_rownorm_fwd_launch<64>(x, ldx, M, C, scale, offset, y, ldy, mean, rstd, st);
}

int gnn_elu_rownorm_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx, int64_t M, int64_t C,
                            const float *scale, const float *mean, const float *rstd, float *dx, int64_t lddx,
                            float *dscale, float *doffset, void *workspace, size_t workspace_bytes, gnn_stream_t stream) {
  if (M < 0 || C < 0) return GNN_E_BADARG;
  if (C > 32 * kEpiMaxPerLane || M >= (1ll << 31)) return GNN_E_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 0) return 0;
  if (!dscale || !doffset) return GNN_E_BADARG;
  if (M == 0) {
    GNN_CUDA(cudaMemsetAsync(dscale, 0, (size_t)C * sizeof(float), st));
    GNN_CUDA(cudaMemsetAsync(doffset, 0, (size_t)C * sizeof(float), st));
    return 0;
  }
  if (!dy || !x || !scale || !mean || !rstd || !dx || lddy < C || ldx < C || lddx < C) return GNN_E_BADARG;
  if (!workspace || workspace_bytes < gnn_elu_rownorm_workspace_bytes(C)) return GNN_E_WORKSPACE;
  float *ws = reinterpret_cast<float *>(workspace);
  const int per = (int)cdiv(C, 32);
  if (C % 4 == 0 && C <= 1024 && aligned16(x, ldx) && aligned16(dy, lddy) && aligned16(dx, lddx) && aligned16(scale, 4)) {
    const int per4 = (int)cdiv(C, 128);
    if (per4 <= 1) return elu_rownorm_bwd_vec_launch<1>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
    if (per4 <= 2) return elu_rownorm_bwd_vec_launch<2>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
    if (per4 <= 4) return elu_rownorm_bwd_vec_launch<4>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
    return elu_rownorm_bwd_vec_launch<8>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
  }
  if (per <= 4) return elu_rownorm_bwd_launch<4>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
  if (per <= 8) return elu_rownorm_bwd_launch<8>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
  if (per <= 16) return elu_rownorm_bwd_launch<16>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
  if (per <= 32) return elu_rownorm_bwd_launch<32>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
  return elu_rownorm_bwd_launch<64>(dy, lddy, x, ldx, M, C, scale, mean, rstd, dx, lddx, dscale, doffset, ws, st);
}

#ifdef GNN_TUNE
#include "experiments_host.cuh"
#endif

int gnn_shard_alloc(size_t bytes, void **dev_ptr, unsigned char ipc_handle[64]) {
  if (!dev_ptr || bytes == 0) return GNN_E_BADARG;
  GNN_CUDA(cudaMalloc(dev_ptr, bytes));
  if (ipc_handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
    if (e != cudaSuccess) { cudaFree(*dev_ptr); *dev_ptr = nullptr; return (int)e; }
    for (int i = 0; i < 64; ++i) ipc_handle[i] = reinterpret_cast<unsigned char *>(&h)[i];
  }
  return 0;
}

int gnn_shard_open(const unsigned char ipc_handle[64], void **dev_ptr) {
  if (!ipc_handle || !dev_ptr) return GNN_E_BADARG;
  cudaIpcMemHandle_t h;
  for (int i = 0; i < 64; ++i) reinterpret_cast<unsigned char *>(&h)[i] = ipc_handle[i];
  GNN_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int gnn_shard_close(void *dev_ptr) {
  if (!dev_ptr) return GNN_E_BADARG;
  GNN_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

int gnn_shard_free(void *dev_ptr) {
  if (!dev_ptr) return 0;
  GNN_CUDA(cudaFree(dev_ptr));
  return 0;
}

int gnn_set_blocking_sync(int on) {
  // cudaDeviceScheduleBlockingSync: threads waiting in a synchronise call sleep instead of spinning.  The sampler threads of
  // a rank wait on three small device-to-host reads per layer; with fewer host cores than threads their spinning takes
  // the cores the other threads need (8 ranks x 5 threads on a 32-core host).
  GNN_CUDA(cudaSetDeviceFlags(on == 1 ? cudaDeviceScheduleBlockingSync : on == 2 ? cudaDeviceScheduleYield : cudaDeviceScheduleAuto));
  return 0;
}

int gnn_host_register(void *host_ptr, size_t bytes, void **dev_alias) {
  if (!host_ptr || !dev_alias || bytes == 0) return GNN_E_BADARG;
  GNN_CUDA(cudaHostRegister(host_ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
  cudaError_t e = cudaHostGetDevicePointer(dev_alias, host_ptr, 0);
  if (e != cudaSuccess) { cudaHostUnregister(host_ptr); return (int)e; }
  return 0;
}

int gnn_host_unregister(void *host_ptr) {
  if (!host_ptr) return GNN_E_BADARG;
  GNN_CUDA(cudaHostUnregister(host_ptr));
  return 0;
}

// ---- dense linears of a layer on tcgen05, 3xTF32 (linear_tc.cuh) -------------------------------------------------
size_t gnn_linear_split_elems(int64_t rows, int64_t cols) {
  if (rows < 0 || cols < 0) return 0;
  return (size_t)2 * (size_t)rows * (size_t)((cols + 31) / 32 * 32);
}

int gnn_linear_split_weights_f32(const float *W, int64_t ldw, int64_t N, int64_t K, float *w_nk, float *w_kn,
                                 gnn_stream_t stream) {
  if (!W || !w_nk || N <= 0 || K <= 0 || ldw < K) return GNN_E_BADARG;
  if (N > INT32_MAX / 2 || K > INT32_MAX / 2) return GNN_E_RANGE;
  const int Kp = tc::round_up((int)K, 32), Np = tc::round_up((int)N, 32);
  const int64_t total = N * Kp + (w_kn ? K * Np : 0);
  const int grid = (int)std::min<int64_t>(cdiv(total, 256), 148 * 8);
  tc::split_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, ldw, (int)N, (int)K, Kp, Np, w_nk, w_kn);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_linear_split_weights2_f32(const float *W0, int64_t ldw0, int64_t N0, int64_t K0, float *w_nk0, float *w_kn0, const float *W1,
                                  int64_t ldw1, int64_t N1, int64_t K1, float *w_nk1, float *w_kn1, gnn_stream_t stream) {
  if (!W0 || !w_nk0 || N0 <= 0 || K0 <= 0 || ldw0 < K0 || !W1 || !w_nk1 || N1 <= 0 || K1 <= 0 || ldw1 < K1) return GNN_E_BADARG;
  if (N0 > INT32_MAX / 2 || K0 > INT32_MAX / 2 || N1 > INT32_MAX / 2 || K1 > INT32_MAX / 2) return GNN_E_RANGE;
  tc::SplitJob j0{W0, ldw0, (int)N0, (int)K0, tc::round_up((int)K0, 32), tc::round_up((int)N0, 32), w_nk0, w_kn0};
  tc::SplitJob j1{W1, ldw1, (int)N1, (int)K1, tc::round_up((int)K1, 32), tc::round_up((int)N1, 32), w_nk1, w_kn1};
  const int64_t t0 = N0 * j0.Kp + (w_kn0 ? K0 * j0.Np : 0), t1 = N1 * j1.Kp + (w_kn1 ? K1 * j1.Np : 0);
  const dim3 grid((unsigned)std::min<int64_t>(cdiv(std::max(t0, t1), 256), 148 * 4), 2);
  tc::split_weights2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(j0, j1);
  GNN_LAUNCH_CHECK();
  return 0;
}

int gnn_linear_tf32x3_f32_ex(const float *A, int64_t lda, const int64_t *a_rows, int64_t M, int64_t K, const float *w_split,
                             int64_t N, const float *bias, float *C, int64_t ldc, const int64_t *c_rows, unsigned flags,
                             gnn_stream_t stream) {
  if (M < 0 || N <= 0 || K <= 0 || lda < 0 || ldc < N) return GNN_E_BADARG;
  if (c_rows && !(flags & GNN_LINEAR_ACCUMULATE)) return GNN_E_BADARG;     // scattered rows are always added
  if (M == 0) return 0;
  if (!A || !w_split || !C) return GNN_E_BADARG;
  if (M > INT32_MAX / 2 || N > INT32_MAX / 2 || K > INT32_MAX / 2) return GNN_E_RANGE;
  if (((uintptr_t)w_split & 15u) != 0) return GNN_E_BADARG;
  tc::EncodeTiledFn encode = tc::encode_tiled_fn();
  if (!encode) return GNN_E_DRIVER;
  // a runtime call first: it makes the device's primary context current on this thread (autograd worker threads reach
  // this function without one, and the driver-API encode below then fails with CUDA_ERROR_INVALID_CONTEXT)
  auto kern = tc::linear_tc_kernel<tc::MODE_NT>;
  GNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kTcSmemBytes));
  const int Kp = tc::round_up((int)K, 32), BN = tc::tile_width_rows((int)N, M);
  CUtensorMap maps[2];
  for (int pl = 0; pl < 2; ++pl) {
    const cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
    const cuuint32_t box[2] = {(cuuint32_t)tc::kBK, (cuuint32_t)BN};
    const cuuint32_t estr[2] = {1, 1};
    void *g = (void *)(w_split + (size_t)pl * (size_t)N * (size_t)Kp);
    const CUtensorMapSwizzle swz = tc::kBK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;   // = the row bytes
    CUresult r = encode(&maps[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
      (void)cudaFree(nullptr);
      r = encode(&maps[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
      if (getenv("GNN_TC_DEBUG"))
        fprintf(stderr, "cuTensorMapEncodeTiled -> %d (plane %d, ptr %p, dims %llu x %llu, stride %llu, box %u x %u)\n", (int)r, pl, g,
                (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides[0], box[0], box[1]);
      return GNN_E_DRIVER;
    }
  }
  tc::TcParams p{};
  p.A = A; p.lda = lda; p.a_rows = a_rows;
  p.C = C; p.ldc = ldc; p.c_rows = c_rows;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.BN = BN;
  p.idesc = tc::make_idesc(BN, false);
  p.desc_lbo = 16; p.desc_sbo = 8 * tc::kBK * 4; p.desc_kstep = 32; p.desc_layout = tc::kBK == 32 ? 2u : 4u;   // SWIZZLE_128B / _64B
  p.a_vec = tc::aligned16(A, lda); p.c_vec = tc::aligned16(C, ldc);
  // The tensor core truncates on every accumulation (linear_tc.cuh): one accumulator chain covers at most K = 1024
  // (128 steps, bias < 2.5e-6); a longer K is cut into equal chunks whose results are added in fp32.
  constexpr int kChain = 1024 / tc::kBK;
  const int nkb = Kp / tc::kBK, nchunks = (nkb + kChain - 1) / kChain, kbpc = (nkb + nchunks - 1) / nchunks;
  const dim3 grid((unsigned)cdiv(M, tc::kBM), (unsigned)cdiv(N, BN));
  for (int ch = 0; ch < nchunks; ++ch) {
    p.kb0 = ch * kbpc; p.kb_per_split = kbpc;
    p.bias = ch == 0 ? bias : nullptr;
    p.accumulate = c_rows ? 2 : (((flags & GNN_LINEAR_ACCUMULATE) || ch > 0) ? 1 : 0);
    kern<<<grid, tc::tc_threads(tc::MODE_NT), tc::kTcSmemBytes, (cudaStream_t)stream>>>(p, maps[0], maps[1]);
    GNN_LAUNCH_CHECK();
  }
  return 0;
}

int gnn_linear_tf32x3_f32(const float *A, int64_t lda, const int64_t *a_rows, int64_t M, int64_t K, const float *w_split,
                          int64_t N, const float *bias, float *C, int64_t ldc, gnn_stream_t stream) {
  return gnn_linear_tf32x3_f32_ex(A, lda, a_rows, M, K, w_split, N, bias, C, ldc, nullptr, 0u, stream);
}

static void wgrad_plan(int64_t M, int64_t N, int64_t K, int &BN, int &splits, int &kbps, int64_t &ldp) {
  BN = tc::tile_width((int)K);
  const int64_t tiles = cdiv(N, tc::kBM) * cdiv(K, BN), total_kb = std::max<int64_t>(cdiv(M, tc::kBK), 1);
  const int64_t sms = tc::sm_count();
  // One accumulator chain covers at most 1024 rows: the tensor core truncates on every accumulation
  // (linear_tc.cuh) and 128 steps keep that bias below 2.5e-6; the partial sums of the splits are added in fp32.
  // Among the split counts that respect it, take the one with the smallest modelled time: waves of CTAs x (k-blocks
  // per CTA + ~3 k-block times of prologue/epilogue) + the fixed-order sum over the splits.
  constexpr int kChain = 1024 / tc::kBK, kOver = 96 / tc::kBK;       // k-blocks per 1024 rows; prologue + epilogue in k-block times
  const int64_t s_min = cdiv(total_kb, kChain), s_max = std::min<int64_t>(total_kb, std::max<int64_t>(3 * s_min, sms / tiles));
  int64_t best = s_min;
  double best_cost = 1e30;
  for (int64_t s = s_min; s <= s_max; ++s) {
    const int64_t per = cdiv(total_kb, s), eff = cdiv(total_kb, per);
    const double cost = (double)cdiv(tiles * eff, sms) * (double)(per + kOver) + 0.15 * (32.0 / tc::kBK) * (double)eff;
    if (cost < best_cost) { best_cost = cost; best = eff; }
  }
  kbps = (int)cdiv(total_kb, best);
  splits = (int)cdiv(total_kb, kbps);
  ldp = (K + 3) / 4 * 4;
}

size_t gnn_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M < 0 || N <= 0 || K <= 0) return 0;
  int BN, splits, kbps; int64_t ldp;
  wgrad_plan(M, N, K, BN, splits, kbps, ldp);
  return (size_t)splits * ((size_t)N * (size_t)ldp + (size_t)((N + 3) / 4 * 4)) * 4 + 16;      // dW partials + dbias partials
}

int gnn_linear_wgrad_tf32x3_f32(const float *dY, int64_t lddy, const float *X, int64_t ldx, const int64_t *x_rows, int64_t M,
                                int64_t N, int64_t K, float *dW, int64_t lddw, float *dbias, void *workspace,
                                size_t workspace_bytes, gnn_stream_t stream) {
  if (M < 0 || N <= 0 || K <= 0 || lddy < N || lddw < K || !dW) return GNN_E_BADARG;
  if (M > INT32_MAX / 2 || N > INT32_MAX / 2 || K > INT32_MAX / 2) return GNN_E_RANGE;
  if (M == 0) {
    GNN_CUDA(cudaMemset2DAsync(dW, (size_t)lddw * 4, 0, (size_t)K * 4, (size_t)N, (cudaStream_t)stream));
    if (dbias) GNN_CUDA(cudaMemsetAsync(dbias, 0, (size_t)N * 4, (cudaStream_t)stream));
    return 0;
  }
  if (!dY || !X) return GNN_E_BADARG;
  int BN, splits, kbps; int64_t ldp;
  wgrad_plan(M, N, K, BN, splits, kbps, ldp);
  if (!workspace || workspace_bytes < gnn_linear_wgrad_workspace_bytes(M, N, K)) return GNN_E_WORKSPACE;
  float *ws = reinterpret_cast<float *>(((uintptr_t)workspace + 15u) & ~(uintptr_t)15u);
  tc::TcParams p{};
  p.A = dY; p.lda = lddy; p.B = X; p.ldb = ldx; p.b_rows = x_rows;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.BN = BN; p.kb_per_split = kbps;
  p.idesc = tc::make_idesc(BN, true);
  p.desc_lbo = tc::kPanelBytes; p.desc_sbo = 512; p.desc_kstep = 1024; p.desc_layout = 1u;   // SWIZZLE_128B_BASE32B
  p.a_vec = tc::aligned16(dY, lddy); p.b_vec = tc::aligned16(X, ldx);
  const bool direct = splits == 1;
  p.C = direct ? dW : ws; p.ldc = direct ? lddw : ldp; p.c_split_stride = N * ldp;
  const int64_t Np4 = (N + 3) / 4 * 4;
  float *db_ws = ws + (size_t)splits * (size_t)N * (size_t)ldp;
  p.dbias = dbias ? (direct ? dbias : db_ws) : nullptr; p.dbias_split_stride = Np4;
  p.c_vec = tc::aligned16(p.C, p.ldc);
  auto kern = tc::linear_tc_kernel<tc::MODE_TN>;
  GNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kTcSmemBytes));
  const dim3 grid((unsigned)cdiv(N, tc::kBM), (unsigned)cdiv(K, BN), (unsigned)splits);
  CUtensorMap dummy{};
  kern<<<grid, tc::tc_threads(tc::MODE_TN), tc::kTcSmemBytes, (cudaStream_t)stream>>>(p, dummy, dummy);
  GNN_LAUNCH_CHECK();
  if (!direct) {
    const int rgrid = (int)std::min<int64_t>(cdiv(N * K, 256), 148 * 8);
    tc::reduce_splits_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>(ws, N * ldp, splits, (int)N, (int)K, ldp, dW, lddw, db_ws, Np4,
                                                                      dbias);
    GNN_LAUNCH_CHECK();
  }
  return 0;
}

// ---- host-side helper of the device LADIES sampler: numpy's legacy weighted draw without replacement -------------------
// Reference sampler.py:128 `np.random.choice(num_nodes, s_num, p=p, replace=False)` after `np.random.seed(seed)` (:96).
// Bit-identical restatement of numpy/random/mtrand.pyx (RandomState.choice, replace=False, p given) on MT19937:
// repeat { x = random_sample(size - n_uniq); p[found] = 0; cdf = cumsum(p) (sequential); cdf /= cdf[-1];
//          new = searchsorted(cdf, x, 'right'); keep the first occurrence of every value, in draw order }.
// 20 ms of numpy per Reddit-shaped minibatch under the GIL become ~4 ms of C without it (the sampler threads of a rank
// share one interpreter).  Pure host code: no CUDA calls; allocates its scratch on the host heap.
static inline uint32_t mt_next(uint32_t *st) {
  constexpr int N = 624, M = 397;
  constexpr uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
  uint32_t *key = st;
  uint32_t pos = st[624];
  if (pos >= (uint32_t)N) {
    int i;
    uint32_t y;
    for (i = 0; i < N - M; ++i) { y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & A); }
    for (; i < N - 1; ++i) { y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A); }
    y = (key[N - 1] & UP) | (key[0] & LO);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
    pos = 0;
  }
  uint32_t y = key[pos++];
  st[624] = pos;
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// number of leading elements of the sorted array a[0..len) that are <= t
static inline int64_t upper_count(const double *a, int64_t len, double t) {
  const double *b = a;
  while (len > 0) {
    const int64_t half = len >> 1;
    const bool le = b[half] <= t;
    b = le ? b + half + 1 : b;
    len = le ? len - half - 1 : half;
  }
  return b - a;
}

// The draw itself.  `pw` is the caller's WORKING copy of the probabilities: the entries drawn are zeroed in it (what numpy
// does to its own copy of p), nothing else is written.
static int legacy_choice_core(uint32_t *mt_state, double *pw, int64_t n, int64_t size, int64_t *found) {
  // scratch lives per thread and only grows: fresh multi-megabyte vectors per call cost more in page faults than the draw
  static thread_local std::vector<double> raw, x, coarse;
  static thread_local std::vector<int32_t> stamp, ends, start;
  static thread_local std::vector<int64_t> cand;
  static thread_local int32_t epoch = 0;          // stamp[l] == epoch: position l was already drawn in the current round
  const int64_t nc = (n + 63) / 64;
  if ((int64_t)raw.size() < n) raw.resize((size_t)n);
  if ((int64_t)x.size() < size + 1) x.resize((size_t)size + 1);
  if ((int64_t)cand.size() < size + 1) cand.resize((size_t)size + 1);
  if ((int64_t)coarse.size() < nc) coarse.resize((size_t)nc);
  if ((int64_t)stamp.size() < n + 1) stamp.resize((size_t)n + 1, 0);
  if (epoch > INT32_MAX - (1 << 20)) {            // rounds are numbered across calls so that the marks need no clearing
    std::fill(stamp.begin(), stamp.end(), 0);
    epoch = 0;
  }
  constexpr int kBuckets = 1 << 16;
  if ((int64_t)ends.size() < kBuckets + 1) ends.resize(kBuckets + 1);
  if ((int64_t)start.size() < size + 1) start.resize((size_t)size + 1);
  // thread_local vectors: their addresses stay out of the hot loops
  double *const rawp = raw.data(), *const xp = x.data(), *const coarsep = coarse.data();
  int32_t *const stampp = stamp.data(), *const endsp = ends.data(), *const startp = start.data();
  int64_t *const candp = cand.data();
  int64_t n_uniq = 0, zeroed = 0;
  int round = 0;
  double mass = 1.0;                              // estimate of cdf[-1] of the coming round (exactness does not matter)
  while (n_uniq < size) {
    const int64_t k = size - n_uniq;
    for (int64_t i = 0; i < k; ++i) {
      const uint32_t a = mt_next(mt_state) >> 5, b = mt_next(mt_state) >> 6;
      xp[i] = ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
    // p[found[0:n_uniq]] = 0 (earlier ones already are); np.cumsum adds sequentially, so the running sums before the
    // first entry zeroed in this round are the ones of the previous round, bit for bit - restart from there
    int64_t lo = round == 0 ? 0 : n;
    for (; zeroed < n_uniq; ++zeroed) {
      const int64_t f = found[zeroed];
      mass -= pw[(size_t)f];
      pw[(size_t)f] = 0.0;
      lo = std::min(lo, f);
    }
    // searchsorted(cdf / cdf[-1], x, side='right') = number of entries whose quotient raw[i] / last (the IEEE division
    // numpy applies to the whole array; rounding keeps it monotone) is <= x.  Located without dividing n numbers:
    // a division-free approximate position for x * last, then the exact quotient test on the neighbours decides.
    // Many draws (the first rounds): a 65,536-bucket table of positions; bucket = floor(running sum * scale), the same
    // monotone map for the sums and for x * last.  Few draws: a two-level binary search (coarse = every 64th sum,
    // L1-resident).
    const bool use_table = k >= 1024 && n >= 4096 && mass > 0.0;
    const double scale = use_table ? (double)kBuckets / mass : 0.0;
    double run = lo > 0 ? rawp[lo - 1] : 0.0;
    for (int64_t i = lo; i < n; ++i) { run += pw[i]; rawp[i] = run; }     // bound by the latency of the dependent additions
    if (use_table) {
      // every kTableStride-th running sum enters the table: e[b] = 1 + the last SAMPLED position whose sum falls into
      // bucket b, a lower bound of the true one - the walk below is a few entries longer, the table costs an eighth
      // (bucketing every sum inside the cumsum loop more than doubled the loop's time: measured)
      constexpr int kTableStride = 8;
      std::fill(endsp, endsp + kBuckets + 1, 0);
      int32_t *e = endsp + 1;
      for (int64_t i = kTableStride - 1; i < n; i += kTableStride) {
        const double v = rawp[i] * scale;
        e[(v >= 0.0 && v < (double)kBuckets) ? (int)v : kBuckets - 1] = (int32_t)i + 1;
      }
      int32_t m = 0;                              // ends[b] := sampled positions in buckets < b (running maximum; ends[0] = 0)
      for (int b = 1; b <= kBuckets; ++b) { m = std::max(m, endsp[b]); endsp[b] = m; }
    } else {
      for (int64_t c = 0; c < nc; ++c) coarsep[c] = rawp[std::min<int64_t>(c * 64 + 63, n - 1)];
    }
    const double last = rawp[n - 1];
    if (!(last > 0.0)) return GNN_E_BADARG;                                   // fewer non-zero entries than `size` (numpy raises)
    mass = last;
    ++round;
    ++epoch;
    int64_t got = 0;
    if (use_table) {
      // table look-ups of all draws first (independent loads), with the lines of `raw` and `stamp` they lead to requested
      // early: the walk below then finds them in cache instead of paying one miss per draw in sequence
      for (int64_t i = 0; i < k; ++i) {
        const double v = xp[i] * last * scale;
        const int32_t l0 = endsp[((v >= 0.0 && v < (double)kBuckets) ? (int)v : kBuckets - 1)];
        startp[i] = l0;
        __builtin_prefetch(rawp + l0);
        __builtin_prefetch(stampp + l0);
      }
    }
    for (int64_t i = 0; i < k; ++i) {
      const double xi = xp[i], t = xi * last;
      int64_t l;
      if (use_table) {
        l = startp[i];                                                  // every sum before l is <= t
        int steps = 0;
        while (l < n && rawp[l] <= t) {
          ++l;
          if (++steps == 48) { l += upper_count(rawp + l, n - l, t); break; }        // a crowded bucket
        }
      } else {
        const int64_t cb = upper_count(coarsep, nc, t);                 // blocks whose LAST element is <= t
        const int64_t b0 = cb * 64, bl = std::min<int64_t>(64, n - b0);
        l = bl > 0 ? b0 + upper_count(rawp + b0, bl, t) : n;
      }
      while (l < n && rawp[l] / last <= xi) ++l;
      while (l > 0 && rawp[l - 1] / last > xi) --l;
      if (stampp[l] != epoch) {            // np.unique(return_index=True) + sort: first occurrence, draw order
        stampp[l] = epoch;
        candp[got++] = l;
      }
    }
    for (int64_t i = 0; i < got; ++i) found[n_uniq + i] = candp[i];
    n_uniq += got;
  }
  return 0;
}

int gnn_legacy_choice_f64(uint32_t *mt_state, const double *p, int64_t n, int64_t size, int64_t *found) {
  if (!mt_state || !p || !found || n <= 0 || size < 0 || size > n) return GNN_E_BADARG;
  if (n > INT32_MAX) return GNN_E_RANGE;
  static thread_local std::vector<double> pw;
  pw.assign(p, p + n);
  return legacy_choice_core(mt_state, pw.data(), n, size, found);
}

// The whole host part of one LADIES layer (reference sampler.py:117-143) in one GIL-free call: probabilities from the
// integer column counts, the weighted draw above, after_nodes = unique(drawn + previous), the normalisation factors and
// the sampled_nodes remap.  Every step is integer arithmetic or the reference's own IEEE expressions, so the outputs
// equal the numpy code's bit for bit (tests/test_sampler_golden.py compares them).
// dense_counts (optional): the count of EVERY node id below num_nodes, so that p[after_nodes] needs no id -> support position map
static int64_t ladies_layer_host_impl(uint32_t *mt_state, const int64_t *nz, const int32_t *counts, int64_t n_nz,
                                      const int64_t *skew_nodes, int64_t n_skew, double scale_factor, const int64_t *previous_nodes,
                                      int64_t n_prev, int64_t samp_num, int64_t *after_nodes, float *normfact, int64_t *sampled,
                                      int64_t *n_sampled, const int32_t *dense_counts, int64_t num_nodes) {
  if (!mt_state || !nz || !counts || n_nz <= 0 || !previous_nodes || n_prev < 0 || samp_num < 0 || !after_nodes || !normfact ||
      !sampled || !n_sampled)
    return GNN_E_BADARG;
  if (n_nz > INT32_MAX) return GNN_E_RANGE;
  static thread_local std::vector<int64_t> pi, found, all, prev;
  static thread_local std::vector<int32_t> pi32;
  static thread_local std::vector<double> p;
  static thread_local std::vector<uint64_t> bits, skewbits;
  static thread_local std::vector<int32_t> rank0, pos_of;
  // pi = column counts (sampler.py:117); locality sampling scales the counts of the nodes cached on this GPU and the
  // reference stores the scaled values back into an int64 array (:119-121): truncation.  Without scaling the int32 counts
  // are read where they lie.
  const bool scaled = scale_factor > 1.0 && skew_nodes && n_skew > 0;
  if (nz[0] < 0) return GNN_E_BADARG;
  // scaled counts: membership in the locality set is a bit test (the set as a bitmap over the support's id range, built
  // per call) and the values stay int32 while the largest count * scale fits, so that this loop and the division below
  // vectorise; sets over ids far sparser than the support, or huge factors,
  // keep the merge of two ascending lists into int64 values.
  const int64_t nz_max = nz[n_nz - 1];
  const bool skew_by_bits = scaled && nz_max + 1 <= 64 * (n_nz + n_skew) + (1 << 20);
  int32_t max_count = 0;
  if (scaled && skew_by_bits) for (int64_t i = 0; i < n_nz; ++i) max_count = std::max(max_count, counts[i]);
  const bool narrow = scaled && skew_by_bits && (double)max_count * scale_factor < 2.0e9;
  auto scaled_count = [scale_factor](int64_t v) -> int64_t { return (int64_t)((double)v * scale_factor); };
  auto in_skew = [&](int64_t node) -> bool { return node <= nz_max && ((skewbits[(size_t)(node >> 6)] >> (node & 63)) & 1ull); };
  int64_t total = 0, n_pos = 0;
  if (scaled && skew_by_bits) {
    skewbits.assign((size_t)((nz_max >> 6) + 1), 0);
    for (int64_t j = 0; j < n_skew; ++j) {
      const int64_t v = skew_nodes[j];
      if (v >= 0 && v <= nz_max) skewbits[(size_t)(v >> 6)] |= 1ull << (v & 63);
    }
    const uint64_t *sb = skewbits.data();
    if (narrow) {
      pi32.resize((size_t)n_nz);
      int32_t *q = pi32.data();
      for (int64_t i = 0; i < n_nz; ++i) {
        const int64_t node = nz[i];
        const int32_t c = counts[i], sc = (int32_t)(int64_t)((double)c * scale_factor);
        const int32_t v = ((sb[node >> 6] >> (node & 63)) & 1ull) ? sc : c;
        q[i] = v;
        total += v;
        n_pos += v > 0;
      }
    } else {
      pi.resize((size_t)n_nz);
      for (int64_t i = 0; i < n_nz; ++i) {
        const int64_t node = nz[i];
        const int64_t v = ((sb[node >> 6] >> (node & 63)) & 1ull) ? scaled_count(counts[i]) : (int64_t)counts[i];
        pi[(size_t)i] = v;
        total += v;
        n_pos += v > 0;
      }
    }
  } else if (scaled) {
    pi.resize((size_t)n_nz);
    int64_t j = 0;
    for (int64_t i = 0; i < n_nz; ++i) {                                     // both ascending: merge
      int64_t v = counts[i];
      while (j < n_skew && skew_nodes[j] < nz[i]) ++j;
      if (j < n_skew && skew_nodes[j] == nz[i]) v = scaled_count(v);
      pi[(size_t)i] = v;
      total += v;
      n_pos += v > 0;
    }
  } else {
    for (int64_t i = 0; i < n_nz; ++i) { total += counts[i]; n_pos += counts[i] > 0; }
  }
  if (total <= 0) return GNN_E_BADARG;
  const double dtotal = (double)total;
  auto pi_at = [&](int64_t i) -> double {
    return !scaled ? (double)counts[i] : narrow ? (double)pi32[(size_t)i] : (double)pi[(size_t)i];
  };
  p.resize((size_t)n_nz);
  {
    double *pp = p.data();                                                                         // p = pi / np.sum(pi)  (:124)
    if (!scaled)     for (int64_t i = 0; i < n_nz; ++i) pp[i] = (double)counts[i] / dtotal;
    else if (narrow) { const int32_t *q = pi32.data(); for (int64_t i = 0; i < n_nz; ++i) pp[i] = (double)q[i] / dtotal; }
    else             { const int64_t *q = pi.data();   for (int64_t i = 0; i < n_nz; ++i) pp[i] = (double)q[i] / dtotal; }
  }
  const int64_t s_num = std::min(n_pos, samp_num);                                                 // :126
  found.resize((size_t)std::max<int64_t>(s_num, 1));
  // the draw zeroes the entries it takes in p; p[after_nodes] below is re-derived by the same division.  (Evaluating the
  // division inside the draw's first cumsum pass was measured: scalar divisions on the addition chain cost more than this
  // vectorised loop.)
  const int rc = legacy_choice_core(mt_state, p.data(), n_nz, s_num, found.data());                 // :128
  if (rc != 0) return rc;
  int64_t max_id = nz[n_nz - 1];
  bool prev_sorted = true;
  for (int64_t i = 0; i < n_prev; ++i) {
    if (previous_nodes[i] < 0) return GNN_E_BADARG;
    max_id = std::max(max_id, previous_nodes[i]);
    prev_sorted = prev_sorted && (i == 0 || previous_nodes[i - 1] < previous_nodes[i]);            // strictly: distinct too
  }
  const int64_t range = max_id + 1;
  int64_t n_after = 0, ns = 0;
  if (range <= 16 * (n_nz + n_prev) + 65536 && range < INT32_MAX) {
    // ids are dense enough for tables over the id range (the Reddit/products shapes: the support is most of the graph):
    // after_nodes = np.unique(drawn ++ previous_nodes) (:131) is a bitmap scan, p[after_nodes] a table lookup and
    // sampled_nodes (:143) a popcount rank - no sorting, no merging
    const int64_t words = (range + 63) >> 6;
    bits.assign((size_t)words, 0);
    for (int64_t i = 0; i < s_num; ++i) { const int64_t v = nz[found[(size_t)i]]; bits[(size_t)(v >> 6)] |= 1ull << (v & 63); }
    for (int64_t i = 0; i < n_prev; ++i) { const int64_t v = previous_nodes[i]; bits[(size_t)(v >> 6)] |= 1ull << (v & 63); }
    // position of a node inside the support: arithmetic when the support is one contiguous id range, else a table
    // (written for every support entry, validated on read: stale contents are harmless)
    const bool direct = dense_counts && (!scaled || skew_by_bits) && range <= num_nodes;
    const bool contiguous = nz[n_nz - 1] - nz[0] == n_nz - 1;
    if (!direct && !contiguous) {
      if ((int64_t)pos_of.size() < range) pos_of.resize((size_t)range);
      for (int64_t j = 0; j < n_nz; ++j) pos_of[(size_t)nz[j]] = (int32_t)j;
    }
    rank0.resize((size_t)words + 1);
    for (int64_t w = 0; w < words; ++w) {
      rank0[(size_t)w] = (int32_t)n_after;
      uint64_t m = bits[(size_t)w];
      while (m) {
        const int64_t node = (w << 6) + __builtin_ctzll(m);
        m &= m - 1;
        double pa;                                                             // p[node]; zero off the support
        if (direct) {
          const int64_t c = dense_counts[node];                                // p of a node straight from the whole count array
          pa = (double)((scaled && in_skew(node)) ? scaled_count(c) : c) / dtotal;
        } else {
          const int64_t j = contiguous ? node - nz[0] : (int64_t)pos_of[(size_t)node];
          pa = (j >= 0 && j < n_nz && nz[j] == node) ? pi_at(j) / dtotal : 0.0;
        }
        // normfact = 1 / np.clip(s_num * p[after_nodes], 1e-10, 1).astype(np.float32)   (:137)
        double v = (double)s_num * pa;
        v = v < 1e-10 ? 1e-10 : (v > 1.0 ? 1.0 : v);
        after_nodes[n_after] = node;
        normfact[n_after] = 1.0f / (float)v;
        ++n_after;
      }
    }
    // sampled_nodes = np.where(np.in1d(after_nodes, previous_nodes))[0]   (:143): ascending positions of the distinct previous nodes
    for (int64_t i = 0; i < n_prev; ++i) {
      const int64_t v = previous_nodes[i];
      sampled[i] = rank0[(size_t)(v >> 6)] + __builtin_popcountll(bits[(size_t)(v >> 6)] & ((1ull << (v & 63)) - 1));
    }
    ns = n_prev;
    if (!prev_sorted) {                                                       // the batch itself: any order, maybe repeats
      std::sort(sampled, sampled + n_prev);
      ns = std::unique(sampled, sampled + n_prev) - sampled;
    }
  } else {
    // sparse ids (papers100M-shaped: a support of ~1e5-1e6 nodes out of 1e8): sort the drawn nodes, merge, drop repeats
    prev.assign(previous_nodes, previous_nodes + n_prev);
    if (!prev_sorted) std::sort(prev.begin(), prev.end());
    const int64_t n_up = std::unique(prev.begin(), prev.end()) - prev.begin();
    for (int64_t i = 0; i < s_num; ++i) found[(size_t)i] = nz[found[(size_t)i]];
    std::sort(found.begin(), found.begin() + s_num);
    all.resize((size_t)(s_num + n_up));
    std::merge(found.begin(), found.begin() + s_num, prev.begin(), prev.begin() + n_up, all.begin());
    n_after = std::unique(all.begin(), all.end()) - all.begin();
    int64_t j = 0;
    for (int64_t i = 0; i < n_after; ++i) {
      const int64_t node = all[(size_t)i];
      while (j < n_nz && nz[j] < node) ++j;
      const double pa = (j < n_nz && nz[j] == node) ? pi_at(j) / dtotal : 0.0;
      double v = (double)s_num * pa;                                           // :137
      v = v < 1e-10 ? 1e-10 : (v > 1.0 ? 1.0 : v);
      after_nodes[i] = node;
      normfact[i] = 1.0f / (float)v;
    }
    int64_t k = 0;                                                             // :143
    for (int64_t i = 0; i < n_up; ++i) {
      while (k < n_after && all[(size_t)k] < prev[(size_t)i]) ++k;
      if (k < n_after && all[(size_t)k] == prev[(size_t)i]) sampled[ns++] = k;
    }
  }
  *n_sampled = ns;
  return n_after;
}

int64_t gnn_ladies_layer_host(uint32_t *mt_state, const int64_t *nz, const int32_t *counts, int64_t n_nz,
                              const int64_t *skew_nodes, int64_t n_skew, double scale_factor, const int64_t *previous_nodes,
                              int64_t n_prev, int64_t samp_num, int64_t *after_nodes, float *normfact, int64_t *sampled,
                              int64_t *n_sampled) {
  return ladies_layer_host_impl(mt_state, nz, counts, n_nz, skew_nodes, n_skew, scale_factor, previous_nodes, n_prev, samp_num,
                                after_nodes, normfact, sampled, n_sampled, nullptr, 0);
}

// Same with, optionally, the whole count array beside the compacted support (p[after_nodes] then needs no id -> position map)
int64_t gnn_ladies_layer_host_ex(uint32_t *mt_state, const int64_t *nz, const int32_t *counts, int64_t n_nz,
                                 const int32_t *counts_dense, int64_t num_nodes, const int64_t *skew_nodes, int64_t n_skew,
                                 double scale_factor, const int64_t *previous_nodes, int64_t n_prev, int64_t samp_num,
                                 int64_t *after_nodes, float *normfact, int64_t *sampled, int64_t *n_sampled) {
  if (counts_dense && num_nodes <= 0) return GNN_E_BADARG;
  return ladies_layer_host_impl(mt_state, nz, counts, n_nz, skew_nodes, n_skew, scale_factor, previous_nodes, n_prev, samp_num,
                                after_nodes, normfact, sampled, n_sampled, counts_dense, counts_dense ? num_nodes : 0);
}

// Same, fed with the device sampler's whole count array (one D2H copy of num_nodes int32 into pinned memory instead of
// a device-side compaction with its two extra stream synchronisations): the support is compacted here, in one pass.
int64_t gnn_ladies_layer_host_dense(uint32_t *mt_state, const int32_t *counts_dense, int64_t num_nodes, const int64_t *skew_nodes,
                                    int64_t n_skew, double scale_factor, const int64_t *previous_nodes, int64_t n_prev,
                                    int64_t samp_num, int64_t *after_nodes, float *normfact, int64_t *sampled, int64_t *n_sampled,
                                    int64_t *n_support) {
  if (!counts_dense || num_nodes <= 0) return GNN_E_BADARG;
  if (num_nodes > INT32_MAX) return GNN_E_RANGE;
  static thread_local std::vector<int64_t> nzv, iota;
  static thread_local std::vector<int32_t> cntv;
  int64_t support = 0;
  for (int64_t i = 0; i < num_nodes; ++i) support += counts_dense[i] != 0;
  if (support == num_nodes) {
    // every node carries probability (the lower layers of a Reddit-shaped minibatch): ids are 0..N-1, counts lie in place
    const int64_t have = (int64_t)iota.size();
    if (have < num_nodes) {
      iota.resize((size_t)num_nodes);
      for (int64_t i = have; i < num_nodes; ++i) iota[(size_t)i] = i;
    }
    if (n_support) *n_support = num_nodes;
    return ladies_layer_host_impl(mt_state, iota.data(), counts_dense, num_nodes, skew_nodes, n_skew, scale_factor, previous_nodes,
                                  n_prev, samp_num, after_nodes, normfact, sampled, n_sampled, counts_dense, num_nodes);
  }
  if ((int64_t)nzv.size() < num_nodes + 1) { nzv.resize((size_t)num_nodes + 1); cntv.resize((size_t)num_nodes + 1); }
  int64_t m = 0;
  int64_t *const nzp = nzv.data();
  int32_t *const cnp = cntv.data();
  for (int64_t i = 0; i < num_nodes; ++i) {                                   // branch-free: write always, advance on non-zero
    const int32_t c = counts_dense[i];
    nzp[m] = i;
    cnp[m] = c;
    m += c != 0;
  }
  if (n_support) *n_support = m;
  if (m == 0) return GNN_E_BADARG;
  return ladies_layer_host_impl(mt_state, nzv.data(), cntv.data(), m, skew_nodes, n_skew, scale_factor, previous_nodes, n_prev,
                                samp_num, after_nodes, normfact, sampled, n_sampled, counts_dense, num_nodes);
}

}  // extern "C"
