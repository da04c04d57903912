// chunk_rows.cuh - which rows does a chunk of consecutive CSR entries touch?
//
// Every nnz-balanced pass of this library (flat SpMM, scatter backward, both transpose passes, build_adj) gives a
// warp a chunk [s, e) of consecutive stored entries and needs (a) the row of the first entry and (b) the row of
// every entry.  A scalar binary search over the row pointer costs log2(M) DEPENDENT loads per warp - 14 round trips
// to a cold L2 for a 16 K-row layer, which is most of the run time of the small launches (top LADIES layer, A^T
// index of sparse layers).  Here: one 32-ary warp search (3 rounds for M <= 32 K, and the first round probes the same
// 32 addresses in every warp of the grid), then ONE coalesced load of the row-pointer window that follows into shared
// memory, then per-entry searches inside that window (shared-memory latency, no global loads).
//
// Included in the middle of gnn_kernels.cu (inside its anonymous namespace).
#pragma once

// first row r with rowptr[r+1] > s.  rowptr is non-decreasing and rowptr[M] == nnz > s, so the predicate is monotone
// and true at hi.  All 32 lanes must call.
__device__ __forceinline__ int warp_first_row(const int *__restrict__ rowptr, int M, int s, int lane) {
  int lo = 0, hi = M - 1;
  while (hi > lo) {
    const int step = (hi - lo + 32) >> 5;
    const int probe = min(lo + (lane + 1) * step - 1, hi);
    const unsigned m = __ballot_sync(kFull, __ldg(rowptr + probe + 1) > s);
    const int first = __ffs(m) - 1;
    const int nhi = min(lo + (first + 1) * step - 1, hi);
    lo += first * step;
    hi = nhi;
  }
  return lo;
}

// largest r in [lo, hi] with rowptr[r] <= i  (skips empty rows that share the same start)
__device__ __forceinline__ int row_of_nnz(const int *__restrict__ rowptr, int lo, int hi, int i) {
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(rowptr + mid) <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Row-pointer window of a chunk [s, e).  win[k] = rowptr[min(r_lo + k, M)] for k in [0, 32*J]; `win` is this warp's
// own shared-memory array of 32*J + 1 ints.
template <int J>
struct ChunkRows {
  int r_lo;          // row of entry s
  const int *win;

  int span;          // window entries win[1..span] are the only ones that can be <= an entry of the chunk

  // all 32 lanes must call; ends with __syncwarp().  e = end of the chunk (exclusive).
  __device__ __forceinline__ void load(const int *__restrict__ rowptr, int M, int s, int e, int lane, int *win_smem) {
    r_lo = warp_first_row(rowptr, M, s, lane);
    int first = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int v = __ldg(rowptr + min(r_lo + 32 * j + lane, M));
      win_smem[32 * j + lane] = v;
      if (j == 0) first = v;
    }
    if (lane == 0) win_smem[32 * J] = __ldg(rowptr + min(r_lo + 32 * J, M));
    // dense rows: the chunk touches 1-3 rows, so the per-entry search only needs the first few window entries.
    // lane l holds win[l]; the first l >= 1 with win[l] >= e bounds every search of this chunk.
    const unsigned beyond = __ballot_sync(kFull, lane >= 1 && first >= e);
    span = beyond ? __ffs(beyond) - 2 : 32 * J;        // win[1..span] < e
    __syncwarp();
    win = win_smem;
  }

  // row of entry i (s <= i < e): r_lo + #{k in [1, span] : win[k] <= i}; beyond the window (a chunk that crosses more
  // than 32J row starts, i.e. runs of empty rows) falls back to the global search
  __device__ __forceinline__ int row_of(const int *__restrict__ rowptr, int M, int i) const {
    int lo = 0, hi = span;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (win[mid + 1] <= i) lo = mid + 1; else hi = mid;
    }
    if (lo == 32 * J) return row_of_nnz(rowptr, r_lo + 32 * J, M - 1, i);
    return r_lo + lo;
  }

  // rowptr[r] / rowptr[r+1] for a row at or after r_lo: from the window when it is inside
  __device__ __forceinline__ int start_of(const int *__restrict__ rowptr, int r) const {
    const int k = r - r_lo;
    return k <= 32 * J ? win[k] : __ldg(rowptr + r);
  }
  __device__ __forceinline__ int end_of(const int *__restrict__ rowptr, int r) const {
    const int k = r - r_lo + 1;
    return k <= 32 * J ? win[k] : __ldg(rowptr + r + 1);
  }
};
