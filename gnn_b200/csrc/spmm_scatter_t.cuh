// spmm_scatter_t.cuh - transpose-free backward: dX[K,D] += A^T . G computed from A's own CSR.
//
// Reference custom_sparse_ops.py:30-37 sorts a transposed copy of A on every backward call.  The default path of this
// library builds the CSR of A^T once per adjacency (gnn_csr_transpose) and runs the gather kernel on it, which is
// deterministic.  This file is the alternative north_star names: no transposed index at all - a warp walks C consecutive
// entries of A, keeps G[r, slab] of the current row in registers and adds v * G[r, slab] into dX[c, slab] with
// red.global.add.v4.f32 (one 16-byte reduction per lane).  The additions of one output row arrive in arbitrary order,
// so results are reproducible only to rounding (like the reference's own atomicAdd kernel, cuda_spmm.cu:205-209); the
// caller zero-fills dX first.  It wins where the transposed index costs more than the product (sparse layers: the
// A^T build is 4 dependent launches), and loses on the dense LADIES blocks, where 4*nnz*D bytes of L2 reductions are
// far slower than the same bytes of L2 reads.  profiles/ has the A/B table; custom_sparse_ops picks per shape.
//
// Included in the middle of gnn_kernels.cu (inside its anonymous namespace).
#pragma once

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct ScatterParams {
  const int *rowptr;
  const int *rowidx;     // row id per stored entry, or NULL (then found through rowptr)
  const int *colidx;
  const float *vals;
  int M, nnz, D, C, nchunks, nslabs;
  const float *G;
  int64_t ldg;
  float *dX;
  int64_t lddx;
};

// One warp per (chunk of C consecutive entries, column slab).  Entry k adds v_k * G[row_k, slab] into dX[col_k, slab].
// G rows are loaded per ENTRY, U at a time (consecutive entries of one row hit L1), so there is no per-row dependent
// load and no row bookkeeping at all; the reductions are fire-and-forget.
template <bool VEC4, int NV, int U, bool ROWIDS>
__global__ void __launch_bounds__(kFlatWarps * 32, 8)
spmm_scatter_t_kernel(const ScatterParams p) {
  __shared__ int col_s[kFlatWarps][kFlatMaxC];
  __shared__ float val_s[kFlatWarps][kFlatMaxC];
  __shared__ int row_s[kFlatWarps][kFlatMaxC];
  __shared__ int win_s[ROWIDS ? 1 : kFlatWarps][ROWIDS ? 1 : kFlatMaxC + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t item = (int64_t)blockIdx.x * kFlatWarps + warp;
  if (item >= (int64_t)p.nchunks * p.nslabs) return;
  const int slab = (int)(item / p.nchunks);
  const int chunk = (int)(item % p.nchunks);
  constexpr int VEC = VEC4 ? 4 : 1;
  constexpr int W = NV * 32 * VEC;
  const int col0 = slab * W + lane * VEC;
  const int s = chunk * p.C;
  const int e = min(s + p.C, p.nnz);
  const int n = e - s;
  int *cs = col_s[warp];
  float *vs = val_s[warp];
  int *rs = row_s[warp];
#pragma unroll
  for (int j = 0; j < kFlatMaxC / 32; ++j) {
    const int k = 32 * j + lane;
    if (k < n) {
      cs[k] = __ldg(p.colidx + s + k);
      vs[k] = __ldg(p.vals + s + k);
      if (ROWIDS) rs[k] = __ldg(p.rowidx + s + k);
    }
  }
  if constexpr (!ROWIDS) {
    ChunkRows<kFlatMaxC / 32> cr;
    cr.load(p.rowptr, p.M, s, e, lane, win_s[warp]);
#pragma unroll
    for (int j = 0; j < kFlatMaxC / 32; ++j) {
      const int k = 32 * j + lane;
      if (k < n) rs[k] = cr.row_of(p.rowptr, p.M, s + k);
    }
  }
  __syncwarp();

  bool full[NV];                                              // the lane's vector lies completely inside [0, D)
#pragma unroll
  for (int q = 0; q < NV; ++q) full[q] = col0 + q * 32 * VEC + VEC <= p.D;

  for (int t = 0; t < n; t += U) {
    float g[U][NV][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float *gr = p.G + (int64_t)rs[min(t + u, n - 1)] * p.ldg + col0;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        if (t + u < n && full[q]) {
          ldg_vec<VEC>(gr + q * 32 * VEC, g[u][q]);
        } else {
#pragma unroll
          for (int w = 0; w < VEC; ++w)
            g[u][q][w] = (t + u < n && col0 + q * 32 * VEC + w < p.D) ? __ldg(gr + q * 32 * VEC + w) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (t + u < n) {
        const float v = vs[t + u];
        float *dr = p.dX + (int64_t)cs[t + u] * p.lddx + col0;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          if (VEC4 && full[q]) {
            red_add_v4(dr + q * 32 * VEC, v * g[u][q][0], v * g[u][q][1 % VEC], v * g[u][q][2 % VEC], v * g[u][q][3 % VEC]);
          } else {
#pragma unroll
            for (int w = 0; w < VEC; ++w)
              if (col0 + q * 32 * VEC + w < p.D) atomicAdd(dr + q * 32 * VEC + w, v * g[u][q][w]);
          }
        }
      }
    }
  }
}

// Speed-of-light probe of the row gather an SpMM performs (bench.py measures the L2->SM gather roof with it in the
// same run, on the same blocks): every warp walks a stretch of the block's own column-index stream and loads the
// addressed X rows - float4 per lane, NV vectors per lane, U rows in flight - and only adds them up.  No values, no
// per-row bookkeeping, no stores.
template <int NV, int U>
__global__ void __launch_bounds__(256)
row_gather_probe_kernel(const float *__restrict__ X, int64_t ldx, const int *__restrict__ colidx, int nnz, int per_warp,
                        int slab_floats, int nslabs, float *__restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int slab = (int)(w % nslabs);
  const int s = (int)(((w / nslabs) * per_warp) % nnz);
  const float *Xs = X + (int64_t)slab * slab_floats;
  float acc[NV][4];
#pragma unroll
  for (int q = 0; q < NV; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
  for (int base = 0; base < per_warp; base += 32) {
    const int cl = __ldg(colidx + (s + base + lane) % nnz);
    for (int t = 0; t < 32; t += U) {
      float4 x[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(kFull, cl, t + u);
        const float4 *xr = reinterpret_cast<const float4 *>(Xs + (int64_t)c * ldx) + lane;
#pragma unroll
        for (int q = 0; q < NV; ++q) x[u][q] = __ldg(xr + q * 32);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < NV; ++q) { acc[q][0] += x[u][q].x; acc[q][1] += x[u][q].y; acc[q][2] += x[u][q].z; acc[q][3] += x[u][q].w; }
    }
  }
  float t = 0.f;
#pragma unroll
  for (int q = 0; q < NV; ++q) t += acc[q][0] + acc[q][1] + acc[q][2] + acc[q][3];
  if (t == 12345.678f) sink[0] = t;
}
