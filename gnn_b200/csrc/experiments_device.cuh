// Experiment-only code (compiled with -DGNN_TUNE by tools/build_tune.sh; never part of libgnn_b200.so).
// Included in the middle of gnn_kernels.cu, so it sees its helpers (kFull, cdiv, GNN_LAUNCH_CHECK, ...).
#ifdef GNN_TUNE
// experiment build only: speed-of-light of the L2->SM gather the SpMM performs (random rows of X, float4 per lane,
// NV vectors per lane, U rows in flight, nothing but the loads and one FADD per float)
template <int NV, int U>
__global__ void __launch_bounds__(256)
gather_roof_kernel(const float *__restrict__ X, int ldx, int K, const int *__restrict__ colidx, int nnz, int per_warp,
                   float *__restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  int s = (int)((w * per_warp) % nnz);
  float acc[NV][4];
#pragma unroll
  for (int n = 0; n < NV; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
  for (int base = 0; base < per_warp; base += 32) {
    const int cl = __ldg(colidx + (s + base + lane) % nnz);
    for (int t = 0; t < 32; t += U) {
      float4 x[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(kFull, cl, t + u);
        const float4 *xr = reinterpret_cast<const float4 *>(X + (int64_t)c * ldx) + lane;
#pragma unroll
        for (int n = 0; n < NV; ++n) x[u][n] = __ldg(xr + n * 32);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int n = 0; n < NV; ++n) { acc[n][0] += x[u][n].x; acc[n][1] += x[u][n].y; acc[n][2] += x[u][n].z; acc[n][3] += x[u][n].w; }
    }
  }
  float t = 0.f;
#pragma unroll
  for (int n = 0; n < NV; ++n) t += acc[n][0] + acc[n][1] + acc[n][2] + acc[n][3];
  if (t == 12345.678f) sink[0] = t;
}
#endif

#ifdef GNN_TUNE
// experiment build only: hub-cached SpMM prototype.  Column slab of 32 floats (128 B per X row); the slab rows of the
// H most frequent columns live in shared memory, every other nonzero goes to L2.  hubslot[c] = slot or -1.
// One CTA = one (slab, nonzero range) item; 8 lanes per X row, 4 nonzeros per warp step, UNR steps in flight.
template <int UNR>
__global__ void __launch_bounds__(1024, 1)
spmm_hub_proto_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx, const float *__restrict__ vals, int M,
                      int nnz, int D, const float *__restrict__ X, int ldx, float *__restrict__ Y, int ldy,
                      const short *__restrict__ hubslot, const int *__restrict__ hubcols, int H, int K, int ranges) {
  extern __shared__ float sm[];
  float *cache = sm;                                        // [H][32]
  short *slot_s = reinterpret_cast<short *>(sm + (size_t)H * 32);   // [K]
  const int slab = blockIdx.x / ranges, range = blockIdx.x % ranges;
  const int col_base = slab * 32;
  for (int i = threadIdx.x; i < H * 8; i += blockDim.x) {
    const int h = i >> 3, q = i & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_base + q * 4 < D) v = __ldg(reinterpret_cast<const float4 *>(X + (int64_t)hubcols[h] * ldx + col_base) + q);
    reinterpret_cast<float4 *>(cache)[i] = v;
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) slot_s[i] = hubslot[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 3, q = lane & 7;                    // 4 groups of 8 lanes
  const bool colok = col_base + q * 4 < D;
  // rows of this range: contiguous block of rows split evenly by nonzeros would need a search; the prototype splits ROWS
  const int rows_per = (M + ranges - 1) / ranges;
  const int r_lo = range * rows_per, r_hi = min(M, r_lo + rows_per);
  for (int r = r_lo + warp; r < r_hi; r += nwarps) {
    const int b = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = b; base < e; base += 32) {
      const int i = base + lane;
      int cl = 0; float vl = 0.f;
      if (i < e) { cl = __ldg(colidx + i); vl = __ldg(vals + i); }
      const int n_here = min(32, e - base);
      for (int t = 0; t < n_here; t += 4 * UNR) {
        float4 x[UNR]; float v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int k = t + u * 4 + g;
          const int c = __shfl_sync(kFull, cl, k & 31);
          v[u] = __shfl_sync(kFull, vl, k & 31);
          x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < n_here && colok) {
            const int sl = slot_s[c];
            if (sl >= 0) x[u] = reinterpret_cast<const float4 *>(cache)[sl * 8 + q];
            else x[u] = __ldg(reinterpret_cast<const float4 *>(X + (int64_t)c * ldx + col_base) + q);
          } else {
            v[u] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          acc.x = fmaf(v[u], x[u].x, acc.x); acc.y = fmaf(v[u], x[u].y, acc.y);
          acc.z = fmaf(v[u], x[u].z, acc.z); acc.w = fmaf(v[u], x[u].w, acc.w);
        }
      }
    }
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(kFull, acc.x, off); acc.y += __shfl_xor_sync(kFull, acc.y, off);
      acc.z += __shfl_xor_sync(kFull, acc.z, off); acc.w += __shfl_xor_sync(kFull, acc.w, off);
    }
    if (g == 0 && colok) {
      float *yr = Y + (int64_t)r * ldy + col_base + q * 4;
      if (col_base + q * 4 + 4 <= D) { yr[0] = acc.x; yr[1] = acc.y; yr[2] = acc.z; yr[3] = acc.w; }
      else { const float a4[4] = {acc.x, acc.y, acc.z, acc.w}; for (int z = 0; col_base + q * 4 + z < D; ++z) yr[z] = a4[z]; }
    }
  }
}
#endif

