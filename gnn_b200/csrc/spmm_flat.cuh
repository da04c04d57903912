// spmm_flat.cuh - nonzero-split ("flat") CSR SpMM for short-row / small blocks.
//
// The row-split kernel walks the rows of its chunk one after another: rowptr[r], rowptr[r+1] -> (col,val) of the
// segment -> X rows -> store, each arrow a dependent trip to L2/HBM.  With 4-60 entries per row (top LADIES layer,
// A^T of any sparse layer, papers100M-shaped blocks) that chain, not bandwidth, is the run time: 21 us for a
// 33 K-entry block whose HBM roof is 6 us.  This kernel removes the per-row trips:
//
//   * a warp owns C = 32..128 consecutive entries of one column slab; their (col, val) are loaded up front in one
//     coalesced pass, the row-pointer window of the chunk in a second one (chunk_rows.cuh), the row id of every entry is
//     found in shared memory, and all three are staged in shared memory;
//   * the walk over the entries then issues U X-row loads back to back REGARDLESS of row boundaries and only afterwards
//     folds them into the running row accumulator; a row change is a warp-uniform branch that stores the finished row;
//   * rows that span chunks use the same partial-slot + last-arriver protocol as the row-split kernel (fixed summation
//     order, bit-reproducible); empty rows are zero-filled by the chunk that finishes the preceding non-empty row.
//
// Included in the middle of gnn_kernels.cu (inside its anonymous namespace).
#pragma once

constexpr int kFlatWarps = 4;        // 128-thread CTAs: small grids spread over more SMs
constexpr int kFlatMaxC = 128;

// Finish a row that spans several chunks: this warp's partial is already stored in its slot.  The last warp to
// arrive adds the partials in ascending chunk order and stores the row.  atomicInc wraps to 0 on the last arrival, so
// a counter array that is zero on entry is zero again when the kernel has finished (see GNN_SPMM_COUNTERS_ZEROED).
//
// A hub row of a sparse block spans dozens of chunks (1,468 entries = 46 chunks of 32): adding its partials one
// dependent L2 load at a time was a 25 us tail on a 30 us kernel.  The loads are independent - only the ADDITIONS have
// a fixed order - so B partials are fetched together (128-bit ld.cg: partials bypass L1, which is why the last arriver
// needs no acquire fence of its own) and then added in order.
template <int VEC, int NV, int LPR>
__device__ __noinline__ void finish_spanning_row(const SpmmParams &p, int r, int slab, int lane, int col0, int row_start,
                                                 int c_first, int c_last, int C, float *yrow, bool y_vec_ok) {
  __threadfence();                                   // release: this warp's partial is visible before its arrival
  __syncwarp();
  int last = 0;
  if (lane == 0)
    last = (atomicInc(reinterpret_cast<unsigned *>(p.counters) + (int64_t)r * p.nslabs + slab, (unsigned)(c_last - c_first)) ==
            (unsigned)(c_last - c_first));
  last = __shfl_sync(kFull, last, 0);
  if (!last) return;
  constexpr int B = NV * VEC >= 12 ? 2 : (NV * VEC >= 8 ? 4 : 8);      // partials in flight per lane (register budget of the caller)
  float acc[NV][VEC];
#pragma unroll
  for (int n = 0; n < NV; ++n) vzero<VEC>(acc[n]);
  if (lane < LPR) {
    for (int ch0 = c_first; ch0 <= c_last; ch0 += B) {
      float t[B][NV][VEC];
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const int ch = min(ch0 + j, c_last);
        const float *ps = p.partials + ((int64_t)2 * ch + (row_start > ch * C ? 1 : 0)) * p.Dp;
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          const int col = col0 + n * LPR * VEC;
          if (VEC == 4 && col + 4 <= p.Dp) {         // slots are Dp (multiple of 4) wide and 16-byte aligned
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(ps + col));
            t[j][n][0] = v.x; t[j][n][1 % VEC] = v.y; t[j][n][2 % VEC] = v.z; t[j][n][3 % VEC] = v.w;
          } else {
#pragma unroll
            for (int q = 0; q < VEC; ++q) t[j][n][q] = col + q < p.D ? __ldcg(ps + col + q) : 0.f;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < B; ++j) {
        if (ch0 + j <= c_last) {
#pragma unroll
          for (int n = 0; n < NV; ++n)
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[n][q] += t[j][n][q];
        }
      }
    }
  }
  store_row<VEC, NV, LPR>(yrow, y_vec_ok, lane, col0, p.D, acc);
}

// CTAs per SM asked of ptxas: the U x NV x VEC floats in flight are the register budget
constexpr int flat_minb(int nv, int vec, int u) { return nv * vec * u >= 64 ? 4 : 8; }

// ROWIDS: the caller supplies the row id of every stored entry (gnn_build_adj / gnn_csr_transpose emit them for free).
// Then the chunk needs ONE round trip before its X loads can be issued - (col, val, row) in three coalesced loads - instead
// of the row search (3 dependent probes) plus the row-pointer window; the row pointer is touched only for the rare row
// that spans chunks.  Without row ids (foreign CSR) the search + window path of chunk_rows.cuh is used.
template <int VEC, int NV, bool GATHER, int U, bool ROWIDS>
__global__ void __launch_bounds__(kFlatWarps * 32, flat_minb(NV, VEC, U))
spmm_flat_kernel(const SpmmParams p, const XSrc<GATHER> xs) {
  __shared__ int col_s[kFlatWarps][kFlatMaxC];
  __shared__ float val_s[kFlatWarps][kFlatMaxC];
  __shared__ int row_s[kFlatWarps][kFlatMaxC];
  __shared__ int win_s[ROWIDS ? 1 : kFlatWarps][ROWIDS ? 1 : kFlatMaxC + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t item = (int64_t)blockIdx.x * kFlatWarps + warp;
  if (item >= (int64_t)p.nchunks * p.nslabs) return;          // whole warps leave; no CTA-wide barrier below
  const int slab = (int)(item / p.nchunks);
  const int chunk = (int)(item % p.nchunks);
  constexpr int W = NV * 32 * VEC;
  const int col0 = slab * W + lane * VEC;
  const bool y_vec_ok = ((reinterpret_cast<uintptr_t>(p.Y) | (uintptr_t)(p.ldy * 4)) & (VEC * 4 - 1)) == 0;
  const int C = p.C;                                          // 32, 64 or 128
  const int s = chunk * C;
  const int e = min(s + C, p.nnz);
  const int n = e - s;

  // (col, val[, row]) of the whole chunk in coalesced loads, staged in shared memory
  int *cs = col_s[warp];
  float *vs = val_s[warp];
  int *rs = row_s[warp];
  int r_before = -1, r_after = -1;                            // rows of the entries just outside the chunk (ROWIDS)
#pragma unroll
  for (int j = 0; j < kFlatMaxC / 32; ++j) {
    const int k = 32 * j + lane;
    if (k < n) {
      cs[k] = __ldg(p.colidx + s + k);
      vs[k] = __ldg(p.vals + s + k);
      if (ROWIDS) rs[k] = __ldg(p.rowidx + s + k);
    }
  }
  ChunkRows<kFlatMaxC / 32> cr;
  if constexpr (ROWIDS) {
    int edge = -1;
    if (lane == 0 && s > 0) edge = __ldg(p.rowidx + s - 1);
    if (lane == 1 && e < p.nnz) edge = __ldg(p.rowidx + e);
    r_before = __shfl_sync(kFull, edge, 0);
    r_after = __shfl_sync(kFull, edge, 1);
    cr.r_lo = 0;
    cr.win = nullptr;
  } else {
    cr.load(p.rowptr, p.M, s, e, lane, win_s[warp]);
#pragma unroll
    for (int j = 0; j < kFlatMaxC / 32; ++j) {
      const int k = 32 * j + lane;
      if (k < n) rs[k] = cr.row_of(p.rowptr, p.M, s + k);
    }
  }
  __syncwarp();

  bool colok[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) colok[q] = col0 + q * 32 * VEC + VEC <= p.Dload;

  const int r_first = rs[0];
  int cur = r_first;
  if (chunk == 0) {                                           // leading empty rows
    for (int z = 0; z < cur; ++z) zero_row<VEC, NV, 32>(p.Y + (int64_t)z * p.ldy, y_vec_ok, lane, col0, p.D);
  }
  float acc[NV][VEC];
#pragma unroll
  for (int q = 0; q < NV; ++q) vzero<VEC>(acc[q]);

  // Store the finished accumulator of row r; `at_end`: r is the row of the chunk's last entry.  A row that spans chunks
  // (at most two per chunk: the one open at its head, the one open at its tail) is stored into its partial slot and
  // finished AFTER the walk - the fence / arrival counter / fix-up call then sits outside the loop that keeps U X rows
  // live in registers.  Returns true when row r ends inside this chunk.
  int pend[2];
  int npend = 0;
  auto flush = [&](int r, bool at_end) -> bool {
    bool open_head, open_tail;
    if constexpr (ROWIDS) {
      open_head = r == r_first && r_before == r;              // the row started in an earlier chunk
      open_tail = at_end && r_after == r;                     // ... continues in a later one
    } else {
      open_head = cr.start_of(p.rowptr, r) < s;
      open_tail = cr.end_of(p.rowptr, r) > e;
    }
    if (!open_head && !open_tail) {
      store_row<VEC, NV, 32>(p.Y + (int64_t)r * p.ldy, y_vec_ok, lane, col0, p.D, acc);
    } else {
      // slot 2*chunk + (row starts strictly inside this chunk): only a row other than the chunk's first one can
      float *slot = p.partials + ((int64_t)2 * chunk + (r != r_first ? 1 : 0)) * p.Dp;
      store_row<VEC, NV, 32>(slot, true, lane, col0, p.D, acc);
      pend[npend++] = r;
    }
    return !open_tail;
  };

  for (int t = 0; t < n; t += U) {
    float x[U][NV][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float *xr = xs.row(cs[min(t + u, n - 1)]) + col0;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        if (t + u < n && colok[q]) ldg_vec<VEC>(xr + q * 32 * VEC, x[u][q]);
        else vzero<VEC>(x[u][q]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (t + u < n) {                                        // warp-uniform
        const int r = rs[t + u];
        const float v = vs[t + u];
        if (r != cur) {                                       // warp-uniform: row `cur` ended inside this chunk
          flush(cur, false);
          for (int z = cur + 1; z < r; ++z) zero_row<VEC, NV, 32>(p.Y + (int64_t)z * p.ldy, y_vec_ok, lane, col0, p.D);
          cur = r;
#pragma unroll
          for (int q = 0; q < NV; ++q) vzero<VEC>(acc[q]);
        }
#pragma unroll
        for (int q = 0; q < NV; ++q)
#pragma unroll
          for (int w = 0; w < VEC; ++w) acc[q][w] = fmaf(v, x[u][q][w], acc[q][w]);
      }
    }
  }
  const bool finished = flush(cur, true);
  for (int i = 0; i < npend; ++i) {
    const int r = pend[i];
    const int row_start = __ldg(p.rowptr + r), row_end = __ldg(p.rowptr + r + 1);
    finish_spanning_row<VEC, NV, 32>(p, r, slab, lane, col0, row_start, row_start >> p.cshift, (row_end - 1) >> p.cshift, C,
                                     p.Y + (int64_t)r * p.ldy, y_vec_ok);
  }
  if (finished) {                                             // this chunk finished row cur: the empty rows after it are ours
    if constexpr (ROWIDS) {
      const int stop = e < p.nnz ? r_after : p.M;
      for (int z = cur + 1; z < stop; ++z) zero_row<VEC, NV, 32>(p.Y + (int64_t)z * p.ldy, y_vec_ok, lane, col0, p.D);
    } else {
      const int row_end = cr.end_of(p.rowptr, cur);
      int z = cur + 1;
      while (z < p.M && __ldg(p.rowptr + z + 1) == row_end) {
        zero_row<VEC, NV, 32>(p.Y + (int64_t)z * p.ldy, y_vec_ok, lane, col0, p.D);
        ++z;
      }
    }
  }
}
