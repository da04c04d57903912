// linear_tc.cuh - the dense half of a GraphSAGE / GCN layer on the 5th-generation tensor cores.
//
// Reference: models.py:18-19  `cat[linearB(x[sampled_nodes]), linearW(spmm(adj, x))]`, models.py:60 `linear(feat)`, and
// their autograd backward (dX = dY.W, dW = dY^T.X).  The reference runs them as fp32 cuBLAS SIMT GEMMs plus an index
// kernel and a concat; after the SpMM work of rounds 1-2 they were the largest part of a training step.
//
// Arithmetic: fp32 in, fp32 out, "3xTF32" inside: every operand element a is split into hi = rn_tf32(a) and
// lo = rn_tf32(a - hi); a.b ~ hi.hi + hi.lo + lo.hi, three tcgen05.mma.kind::tf32 into one fp32 TMEM accumulator.  The
// dropped terms are < 2^-21 |a||b| per product (fp32 rounding of one FMA: 2^-24), far inside the 1e-5 bar.
//
// Two kernels from one template:
//   NT  C[M,N] = A[rows[m], :K] . W[N,K]^T + bias      forward and dX.  A: activations, split while they are staged (one
//       pass over HBM, the row gather of x[sampled_nodes] is free: a producer thread owns a row pointer).  W: pre-split
//       once per step into [2][N][Kp] (gnn_linear_split_weights_f32) and loaded by TMA, 128-byte swizzle.
//   TN  dW[N,K] = sum_m dY[m,n] X[rows[m],k]           both operands are activations whose REDUCTION index is the slow
//       one in memory: natural [32 rows x 128 B] panels are exactly the MN-major SWIZZLE_128B operand layout, so no
//       transposes; split over m across CTAs (partials in a workspace, summed in fixed order => reproducible).
// Per CTA: 8 producer warps (global -> registers -> hi/lo -> swizzled shared), 1 TMA warp (NT), 1 MMA warp (one lane
// issues), 2 stages of 96 KB, accumulator 128 x BN fp32 in TMEM; the producer warps turn into the epilogue
// (tcgen05.ld -> shared transpose -> coalesced 128-bit stores, bias fused).
#pragma once

namespace tc {

constexpr int kBM = 128;                       // UMMA M (TMEM lanes)
#ifndef GNN_TC_BK
#define GNN_TC_BK 32
#endif
// fp32 per k-block: 32 (128-byte rows, SWIZZLE_128B, 2 stages of 96 KB) or 16 (64-byte rows, SWIZZLE_64B, 4 stages of
// 48 KB).  Measured A/B on one B200 (profiles/r2_linear_tc.md): the 4-stage variant is SLOWER (NT 60 vs 56 us, dW 160 vs
// 118 us) - the kernel is bound by shared-memory bandwidth (operand reads of three products + hi/lo writes: ~240 KB per
// 32-wide k-block against 128 B/clk), not by the refill latency more stages would hide, and halving the stage doubles the
// per-stage barrier / fence cost.
constexpr int kBK = GNN_TC_BK;
static_assert(kBK == 16 || kBK == 32, "k-block is 16 or 32 floats");
constexpr int kStages = kBK == 16 ? 4 : 2;
// producer warps: 8 for NT (4 chunks per thread and k-block; the kernel is shared-memory-bound), 16 for TN, whose producers move
// three times the bytes through registers and are issue-bound (6 chunks per thread and k-block instead of 12)
__host__ __device__ constexpr int producer_warps(int mode) { return mode == 1 ? 16 : 8; }
__host__ __device__ constexpr int tc_threads(int mode) { return (producer_warps(mode) + 2) * 32; }
constexpr int kMaxBN = 256;
constexpr uint32_t kPanelBytes = kBK * 128;    // MN-major panel: [kBK rows][128 B]
constexpr uint32_t kATile = kBM * kBK * 4;     // K-major: 128 rows x kBK floats; MN-major: 4 panels
constexpr uint32_t kBTile = kMaxBN * kBK * 4;
constexpr uint32_t kStageBytes = 2 * kATile + 2 * kBTile;
constexpr uint32_t kBarBytes = 256;
constexpr uint32_t kTcSmemBytes = 1024 + kStages * kStageBytes + kBarBytes;
constexpr uint32_t kTmemCols = 512;             // two accumulators of 128 x 256 fp32: hi.hi and the cross terms
constexpr int kEpiStride = 36;                 // floats per scratch row: 16-byte aligned, conflict-free both ways

enum { MODE_NT = 0, MODE_TN = 1 };

struct TcParams {
  const float *A; int64_t lda; const int64_t *a_rows;   // NT: activations [M,K] (rows optional); TN: dY [M,N]
  const float *B; int64_t ldb; const int64_t *b_rows;   // TN only: X [M,K] (rows optional)
  const float *bias;                                    // NT only, may be NULL
  float *dbias; int64_t dbias_split_stride;             // TN only, may be NULL: column sums of dY (bias gradient), per split
  float *C; int64_t ldc; int64_t c_split_stride;        // TN: partial of split z at C + z * c_split_stride
  const int64_t *c_rows;                                // NT: output row m goes to C[c_rows[m]] (atomic adds), or NULL
  int accumulate;                                       // NT: 0 store, 1 C += (plain read-modify-write), 2 red.add
  int kb0;                                              // NT: first k-block of this launch (K chunks of one product)
  int M, N, K;                                          // NT: C is [M,N], reduce over K.  TN: C is [N,K], reduce over M
  int BN;                                               // tile width, multiple of 16, <= 256
  int kb_per_split;                                     // TN
  uint32_t idesc;
  uint32_t desc_lbo, desc_sbo, desc_kstep, desc_layout; // shared-memory descriptor strides (bytes) and layout type of this mode
  int a_vec, b_vec, c_vec;                              // 16-byte access allowed on that operand
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// A wait that cannot hang the GPU: a protocol bug traps (the launch fails with an error) instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// shared-memory matrix descriptor with the sm_100 version bit (cute::UMMA::SmemDescriptor).  layout: 2 = SWIZZLE_128B
// (16-byte chunks, K-major operands), 1 = SWIZZLE_128B_BASE32B (32-byte chunks: the only layout MN-major tf32 accepts)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// round to nearest TF32 (ties away from zero, what cvt.rna.tf32.f32 does): add half an ulp of the 10-bit mantissa to the
// magnitude bits and clear the 13 low bits.  Two integer instructions instead of the four the cvt expands to (it
// special-cases Inf/NaN; here Inf stays Inf and NaN stays NaN as well, only the NaN payload may differ).
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// four consecutive floats of a row starting at column `col`; columns >= limit (and a null row) read as zero
__device__ __forceinline__ float4 load4(const float *row, int col, int limit, int vec_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row == nullptr || col >= limit) return v;
  if (vec_ok && col + 4 <= limit) return __ldg(reinterpret_cast<const float4 *>(row + col));
  v.x = __ldg(row + col);
  if (col + 1 < limit) v.y = __ldg(row + col + 1);
  if (col + 2 < limit) v.z = __ldg(row + col + 2);
  if (col + 3 < limit) v.w = __ldg(row + col + 3);
  return v;
}
// hi / lo planes of one 16-byte chunk at the same swizzled offset of their tiles
__device__ __forceinline__ void split_store(uint32_t hi_addr, uint32_t lo_addr, const float4 v) {
  const float hx = tf32_rn(v.x), hy = tf32_rn(v.y), hz = tf32_rn(v.z), hw = tf32_rn(v.w);
  sts128(hi_addr, hx, hy, hz, hw);
  sts128(lo_addr, tf32_rn(v.x - hx), tf32_rn(v.y - hy), tf32_rn(v.z - hz), tf32_rn(v.w - hw));
}

template <int MODE>
__global__ void __launch_bounds__(tc_threads(MODE), 1)
linear_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kStages * kStageBytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kStages, bar_accum = bars + 16 * kStages;
  const uint32_t tmem_slot = bar_accum + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kProducerWarps = producer_warps(MODE);

  // tile coordinates
  //   NT: rows of C = blockIdx.x * 128 (m), columns = blockIdx.y * BN (n), reduce over all of K
  //   TN: rows of C = blockIdx.x * 128 (n), columns = blockIdx.y * BN (k), reduce over k-blocks of split blockIdx.z
  const int row0 = blockIdx.x * kBM, col0 = blockIdx.y * p.BN;
  int kb_begin = 0, kb_end = 0;
  if (MODE == MODE_NT) {
    kb_begin = p.kb0;
    kb_end = min((p.K + kBK - 1) / kBK, p.kb0 + p.kb_per_split);
  } else {
    const int total = (p.M + kBK - 1) / kBK;
    kb_begin = blockIdx.z * p.kb_per_split;
    kb_end = min(total, kb_begin + p.kb_per_split);
  }
  const int nkb = kb_end - kb_begin;

  if (tid == 0) {
    const uint32_t producers = kProducerWarps * 32 + (MODE == MODE_NT ? 1 : 0);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, producers);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp < kProducerWarps) {
    // ===== producers: global -> registers -> (hi, lo) -> swizzled shared =====================================
    // NT (K-major): a row of the tile is kBK floats = CPR 16-byte chunks; 256 threads cover RPP rows per pass.
    //   SWIZZLE_128B (kBK 32): chunk ^= row & 7.   SWIZZLE_64B (kBK 16): chunk ^= (row >> 1) & 3  (byte bits [4,6) ^= [7,9)).
    // TN (MN-major): a panel is kBK reduction rows of 128 bytes; SWIZZLE_128B_BASE32B: 32-byte chunk ^= row & 3
    //   (byte bits [5,7) ^= [7,9)).  With kBK 16 a panel has 128 chunk slots: the two thread halves take different panels.
    constexpr int CPR = (MODE == MODE_NT) ? kBK / 4 : 8;
    constexpr int RPP = (kProducerWarps * 32) / CPR;
    constexpr int NPASS = kBM / RPP;                       // NT: row passes per thread
    constexpr int TNG = (kProducerWarps * 32) / (kBK * 8); // TN: thread groups (1 or 2)
    const int c = tid % CPR;
    const int r = (MODE == MODE_NT) ? tid / CPR : (tid / 8) % kBK;
    const int grp = (MODE == MODE_NT) ? 0 : tid / (kBK * 8);
    const uint32_t swz = (MODE == MODE_NT)
                             ? (uint32_t)(r * (kBK * 4) + ((c ^ (kBK == 32 ? (r & 7) : ((r >> 1) & 3))) << 4))
                             : (uint32_t)(r * 128 + ((c ^ ((r & 3) << 1)) << 4));
    if (MODE == MODE_NT) {
      const float *rp[NPASS];
#pragma unroll
      for (int i = 0; i < NPASS; ++i) {
        const int m = row0 + r + RPP * i;
        rp[i] = (m < p.M) ? p.A + (p.a_rows ? p.a_rows[m] : (int64_t)m) * p.lda : nullptr;
      }
      // Two register sets: the loads of k-block it+1 are issued BEFORE the stores of k-block it, so a whole k-block
      // period (not just the wait for a free stage) covers their latency.
      float4 v0[NPASS], v1[NPASS];
      // whole k-block inside K and 16-byte aligned rows (warp-uniform): one predicated 128-bit load per row; otherwise
      // (last k-block of K = 602, unaligned rows) the bounds-checked path
      auto load = [&](int kb, float4 (&v)[NPASS]) {
        const int col = kb * kBK + c * 4;
        if (p.a_vec && kb * kBK + kBK <= p.K) {
#pragma unroll
          for (int i = 0; i < NPASS; ++i)
            v[i] = rp[i] ? __ldg(reinterpret_cast<const float4 *>(rp[i] + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int i = 0; i < NPASS; ++i) v[i] = load4(rp[i], col, p.K, p.a_vec);
        }
      };
      auto put = [&](int it, const float4 (&v)[NPASS]) {
        const int s = it % kStages;
        mbar_wait(bar_empty + 8 * s, ((it / kStages) & 1) ^ 1);
        const uint32_t a_hi = base + s * kStageBytes, a_lo = a_hi + kATile;
#pragma unroll
        for (int i = 0; i < NPASS; ++i) split_store(a_hi + swz + i * (RPP * kBK * 4), a_lo + swz + i * (RPP * kBK * 4), v[i]);
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * s);
      };
      if (nkb > 0) load(kb_begin, v0);
      for (int it = 0; it < nkb; it += 2) {
        if (it + 1 < nkb) load(kb_begin + it + 1, v1);
        put(it, v0);
        if (it + 1 < nkb) {
          if (it + 2 < nkb) load(kb_begin + it + 2, v0);
          put(it + 1, v1);
        }
      }
    } else {
      const int npanels = (p.BN + 31) >> 5;
      constexpr int NA = 4 / TNG, NB = 8 / TNG;            // A / B panels per thread
      // two register sets, as in the NT producer: loads one k-block ahead of the stores; the gather index of the row
      // after that is fetched alongside, so a gathered row costs one exposed round trip, not two
      float4 va0[NA], vb0[NB], va1[NA], vb1[NB];
      float4 dsum[NA];
#pragma unroll
      for (int q = 0; q < NA; ++q) dsum[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      // which of this thread's 16-byte chunks are fully inside the operand (128-bit load), partly inside (bounds-checked
      // path) or outside: loop-invariant, only the reduction row changes per k-block
      uint32_t a_full = 0, a_part = 0, b_full = 0, b_part = 0;
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const int col = 32 * (grp * NA + q) + 4 * c, lim = p.N - row0;
        if (p.a_vec && col + 4 <= lim) a_full |= 1u << q; else if (col < lim) a_part |= 1u << q;
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int col = 32 * (grp * NB + q) + 4 * c, lim = p.K - col0;
        if (grp * NB + q < npanels) { if (p.b_vec && col + 4 <= lim) b_full |= 1u << q; else if (col < lim) b_part |= 1u << q; }
      }
      const int a_col0 = 32 * grp * NA + 4 * c, b_col0 = 32 * grp * NB + 4 * c;
      auto row_of = [&](int kb) -> int64_t {                // X row of this thread's reduction row in k-block kb (-1: past the end)
        const int m = kb * kBK + r;
        if (kb >= kb_end || m >= p.M) return -1;
        return p.b_rows ? __ldg(p.b_rows + m) : (int64_t)m;
      };
      int64_t xrow_next = nkb > 0 ? row_of(kb_begin) : -1;
      auto load = [&](int kb, float4 (&va)[NA], float4 (&vb)[NB]) {
        const int m = kb * kBK + r;
        const int64_t xrow = xrow_next;
        xrow_next = row_of(kb + 1);
        const bool ok = xrow >= 0;
        const float *ar = p.A + (int64_t)(ok ? m : 0) * p.lda + row0 + a_col0;
        const float *br = p.B + (ok ? xrow : 0) * p.ldb + col0 + b_col0;
        const uint32_t af = ok ? a_full : 0u, ap = ok ? a_part : 0u, bf = ok ? b_full : 0u, bp = ok ? b_part : 0u;
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          va[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (af >> q & 1u) va[q] = __ldg(reinterpret_cast<const float4 *>(ar + 32 * q));
          else if (ap >> q & 1u) va[q] = load4(ar - a_col0, a_col0 + 32 * q, p.N - row0, 0);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          vb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bf >> q & 1u) vb[q] = __ldg(reinterpret_cast<const float4 *>(br + 32 * q));
          else if (bp >> q & 1u) vb[q] = load4(br - b_col0, b_col0 + 32 * q, p.K - col0, 0);
        }
      };
      auto put = [&](int it, const float4 (&va)[NA], const float4 (&vb)[NB]) {
        const int s = it % kStages;
        mbar_wait(bar_empty + 8 * s, ((it / kStages) & 1) ^ 1);
        const uint32_t a_hi = base + s * kStageBytes, a_lo = a_hi + kATile, b_hi = a_lo + kATile, b_lo = b_hi + kBTile;
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          split_store(a_hi + (grp * NA + q) * kPanelBytes + swz, a_lo + (grp * NA + q) * kPanelBytes + swz, va[q]);
          dsum[q].x += va[q].x; dsum[q].y += va[q].y; dsum[q].z += va[q].z; dsum[q].w += va[q].w;    // bias gradient: column sums of dY
        }
#pragma unroll
        for (int q = 0; q < NB; ++q)
          if (grp * NB + q < npanels)
            split_store(b_hi + (grp * NB + q) * kPanelBytes + swz, b_lo + (grp * NB + q) * kPanelBytes + swz, vb[q]);
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * s);
      };
      if (nkb > 0) load(kb_begin, va0, vb0);
      for (int it = 0; it < nkb; it += 2) {
        if (it + 1 < nkb) load(kb_begin + it + 1, va1, vb1);
        put(it, va0, vb0);
        if (it + 1 < nkb) {
          if (it + 2 < nkb) load(kb_begin + it + 2, va0, vb0);
          put(it + 1, va1, vb1);
        }
      }
      // Bias gradient (column sums of this CTA's dY rows): the k-tile-0 CTAs fold their threads' running sums over the 32
      // reduction rows in a fixed order, through the second operand stage (free once the accumulator is complete).
      if (p.dbias != nullptr && blockIdx.y == 0) {
        if (nkb > 0) mbar_wait(bar_accum, 0);
        float *dbs = reinterpret_cast<float *>(tc_smem_raw + (base - smem_u32(tc_smem_raw)) + kStageBytes);     // [32 rows][128 columns]
#pragma unroll
        for (int q = 0; q < NA; ++q) *reinterpret_cast<float4 *>(dbs + r * kBM + 32 * (grp * NA + q) + 4 * c) = dsum[q];
        asm volatile("bar.sync 1, %0;" ::"r"(kProducerWarps * 32) : "memory");
        if (tid < kBM) {
          float acc = 0.f;
#pragma unroll 8
          for (int j = 0; j < kBK; ++j) acc += dbs[j * kBM + tid];
          if (row0 + tid < p.N) p.dbias[(int64_t)blockIdx.z * p.dbias_split_stride + row0 + tid] = acc;
        }
      }
    }
  } else if (warp == kProducerWarps) {
    // ===== TMA warp (NT): the pre-split weight planes =========================================================
    if (MODE == MODE_NT && lane == 0) {
      const uint32_t bytes = 2u * (uint32_t)p.BN * (uint32_t)(kBK * 4);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % kStages;
        mbar_wait(bar_empty + 8 * s, ((it / kStages) & 1) ^ 1);
        const uint32_t b_hi = base + s * kStageBytes + 2 * kATile, b_lo = b_hi + kBTile;
        mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
        tma_load_2d(b_hi, &map_hi, bar_full + 8 * s, (kb_begin + it) * kBK, col0);
        tma_load_2d(b_lo, &map_lo, bar_full + 8 * s, (kb_begin + it) * kBK, col0);
      }
    }
    __syncwarp();
  } else {
    // ===== MMA warp: one lane issues ==========================================================================
    if (lane == 0) {
      uint32_t acc = 0;
      for (int it = 0; it < nkb; ++it) {
        const int s = it % kStages;
        mbar_wait(bar_full + 8 * s, (it / kStages) & 1);
        tc_fence_after();
        const uint32_t a_hi = base + s * kStageBytes, a_lo = a_hi + kATile, b_hi = a_lo + kATile, b_lo = b_hi + kBTile;
#pragma unroll
        for (int kk = 0; kk < kBK / 8; ++kk) {
          // NT (K-major, SWIZZLE_128B / _64B): 8-row groups 8 * row bytes apart (SBO), 8 tf32 = 32 B further along the
          // swizzled row per MMA.  TN (MN-major, SWIZZLE_128B_BASE32B): 32-element MN blocks one panel apart (LBO),
          // groups of 4 reduction rows 512 B apart (SBO), 8 reduction rows = 1024 B per MMA.
          const uint32_t ko = kk * p.desc_kstep;
          const uint32_t lt = p.desc_layout;
          const uint64_t dah = smem_desc(a_hi + ko, p.desc_lbo, p.desc_sbo, lt), dal = smem_desc(a_lo + ko, p.desc_lbo, p.desc_sbo, lt);
          const uint64_t dbh = smem_desc(b_hi + ko, p.desc_lbo, p.desc_sbo, lt), dbl = smem_desc(b_lo + ko, p.desc_lbo, p.desc_sbo, lt);
          // The tensor core truncates (round toward zero) on every accumulation: a bias of ~2^-25 |acc| per step.
          // The two cross terms are 2^-11 of the main term; in their own accumulator their steps cost nothing and
          // the main accumulator sees one step per 8 k instead of three.
          umma_tf32(tmem_base + kMaxBN, dal, dbh, p.idesc, acc);
          umma_tf32(tmem_base + kMaxBN, dah, dbl, p.idesc, 1);
          umma_tf32(tmem_base, dah, dbh, p.idesc, acc);
          acc = 1;
        }
        umma_commit(bar_empty + 8 * s);
      }
      umma_commit(bar_accum);
    }
    __syncwarp();
  }

  // ===== epilogue: TMEM -> registers -> shared (transpose) -> global ============================================
  if (warp < kProducerWarps) {
    float *C = p.C + (MODE == MODE_TN ? (int64_t)blockIdx.z * p.c_split_stride : 0);
    const int rows_total = (MODE == MODE_NT) ? p.M : p.N;
    const int cols_total = (MODE == MODE_NT) ? p.N : p.K;
    const int q = warp & 3, part = warp >> 2;              // TMEM lane quarter of this warp; column chunks are dealt round-robin
    const int nchunks = (p.BN + 31) >> 5;
    if (nkb > 0) {
      mbar_wait(bar_accum, 0);
      tc_fence_after();
    }
    const uint32_t scratch = base + warp * (32 * kEpiStride * 4);
    for (int ch = part; ch < nchunks; ch += kProducerWarps / 4) {
      uint32_t v[32];
      if (nkb > 0) {
        uint32_t w[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kMaxBN + ch * 32), w);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(scratch + (lane * kEpiStride + 4 * j) * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      __syncwarp();
      const int cl = ch * 32 + (lane & 7) * 4;          // column inside the tile
      const int col = col0 + cl;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == MODE_NT && p.bias != nullptr) {
        if (col < cols_total) b4.x = __ldg(p.bias + col);
        if (col + 1 < cols_total) b4.y = __ldg(p.bias + col + 1);
        if (col + 2 < cols_total) b4.z = __ldg(p.bias + col + 2);
        if (col + 3 < cols_total) b4.w = __ldg(p.bias + col + 3);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = (lane >> 3) + 4 * i;
        const int row = row0 + q * 32 + rl;
        float4 o = lds128(scratch + (rl * kEpiStride + (lane & 7) * 4) * 4);
        o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
        if (row < rows_total && cl < p.BN && col < cols_total) {
          const int64_t crow = (MODE == MODE_NT && p.c_rows != nullptr) ? p.c_rows[row] : (int64_t)row;
          float *dst = C + crow * p.ldc + col;
          const bool full = cl + 4 <= p.BN && col + 4 <= cols_total;
          if (MODE == MODE_NT && p.accumulate == 2) {          // rows of several launches / duplicates may meet: L2 reductions
            if (p.c_vec && full) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
            } else {
              atomicAdd(dst, o.x);
              if (cl + 1 < p.BN && col + 1 < cols_total) atomicAdd(dst + 1, o.y);
              if (cl + 2 < p.BN && col + 2 < cols_total) atomicAdd(dst + 2, o.z);
              if (cl + 3 < p.BN && col + 3 < cols_total) atomicAdd(dst + 3, o.w);
            }
          } else if (p.c_vec && full) {
            if (MODE == MODE_NT && p.accumulate == 1) {
              const float4 old = *reinterpret_cast<const float4 *>(dst);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4 *>(dst) = o;
          } else {
            const bool acc = MODE == MODE_NT && p.accumulate == 1;
            dst[0] = acc ? dst[0] + o.x : o.x;
            if (cl + 1 < p.BN && col + 1 < cols_total) dst[1] = acc ? dst[1] + o.y : o.y;
            if (cl + 2 < p.BN && col + 2 < cols_total) dst[2] = acc ? dst[2] + o.z : o.z;
            if (cl + 3 < p.BN && col + 3 < cols_total) dst[3] = acc ? dst[3] + o.w : o.w;
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- weights: W[N,K] -> hi/lo planes, zero padded, in both orientations -----------------------------------------
__global__ void split_weights_kernel(const float *__restrict__ W, int64_t ldw, int N, int K, int Kp, int Np,
                                     float *__restrict__ w_nk, float *__restrict__ w_kn) {
  const int64_t n_nk = (int64_t)N * Kp, n_kn = w_kn ? (int64_t)K * Np : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nk + n_kn; i += (int64_t)gridDim.x * blockDim.x) {
    float x;
    float *hi, *lo;
    if (i < n_nk) {
      const int n = (int)(i / Kp), k = (int)(i % Kp);
      x = k < K ? W[(int64_t)n * ldw + k] : 0.f;
      hi = w_nk + i; lo = w_nk + n_nk + i;
    } else {
      const int64_t j = i - n_nk;
      const int k = (int)(j / Np), n = (int)(j % Np);
      x = n < N ? W[(int64_t)n * ldw + k] : 0.f;
      hi = w_kn + j; lo = w_kn + n_kn + j;
    }
    const float h = tf32_rn(x);
    *hi = h;
    *lo = tf32_rn(x - h);
  }
}

// the two weight matrices of a GraphSAGE layer (linearB, linearW) in one launch: blockIdx.y picks the matrix
struct SplitJob { const float *W; int64_t ldw; int N, K, Kp, Np; float *w_nk, *w_kn; };
__global__ void split_weights2_kernel(const SplitJob j0, const SplitJob j1) {
  const SplitJob &j = blockIdx.y == 0 ? j0 : j1;
  const int64_t n_nk = (int64_t)j.N * j.Kp, n_kn = j.w_kn ? (int64_t)j.K * j.Np : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nk + n_kn; i += (int64_t)gridDim.x * blockDim.x) {
    float x;
    float *hi, *lo;
    if (i < n_nk) {
      const int n = (int)(i / j.Kp), k = (int)(i % j.Kp);
      x = k < j.K ? j.W[(int64_t)n * j.ldw + k] : 0.f;
      hi = j.w_nk + i; lo = j.w_nk + n_nk + i;
    } else {
      const int64_t t = i - n_nk;
      const int k = (int)(t / j.Np), n = (int)(t % j.Np);
      x = n < j.N ? j.W[(int64_t)n * j.ldw + k] : 0.f;
      hi = j.w_kn + t; lo = j.w_kn + n_kn + t;
    }
    const float h = tf32_rn(x);
    *hi = h;
    *lo = tf32_rn(x - h);
  }
}

// dW[n,k] = sum over splits in ascending order (fixed => reproducible)
__global__ void reduce_splits_kernel(const float *__restrict__ ws, int64_t split_stride, int splits, int N, int K, int64_t ldp,
                                     float *__restrict__ out, int64_t ldo, const float *__restrict__ db_ws, int64_t db_stride,
                                     float *__restrict__ db_out) {
  const int64_t total = (int64_t)N * K;
  if (db_out != nullptr) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
      float acc = db_ws[n];
      for (int s = 1; s < splits; ++s) acc += db_ws[(int64_t)s * db_stride + n];
      db_out[n] = acc;
    }
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / K), k = (int)(i % K);
    const float *src = ws + (int64_t)n * ldp + k;
    float acc = src[0];
    for (int s = 1; s < splits; ++s) acc += src[(int64_t)s * split_stride];
    out[(int64_t)n * ldo + k] = acc;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    return (EncodeTiledFn)f;
  }();
  return fn;
}

inline int sm_count() {
  static int n = []() {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

inline int tile_width(int cols) {
  const int nt = (cols + kMaxBN - 1) / kMaxBN;
  return std::min(kMaxBN, round_up((cols + nt - 1) / nt, 16));
}

// NT: with few row tiles (the 512-row top layer) 256-wide tiles leave most SMs idle and the k loop of a CTA is a pure
// latency chain; narrower tiles (down to 64) give ~half the SMs something to do
inline int tile_width_rows(int cols, int64_t rows) {
  const int64_t row_tiles = (rows + kBM - 1) / kBM;
  int nt = (cols + kMaxBN - 1) / kMaxBN;
  const int64_t want = (sm_count() / 2 + row_tiles - 1) / row_tiles;
  if (want > nt) nt = (int)std::min<int64_t>(want, std::max(1, cols / 64));
  return std::min(kMaxBN, round_up((cols + nt - 1) / nt, 16));
}

inline uint32_t make_idesc(int bn, bool mn_major) {
  uint32_t d = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
  if (mn_major) d |= (1u << 15) | (1u << 16);
  return d;
}

inline bool aligned16(const void *p, int64_t ld) { return (((uintptr_t)p | (uintptr_t)(ld * 4)) & 15u) == 0; }

}  // namespace tc
