// epilogue_vec.cuh - 128-bit versions of the fused layer epilogue (reference models.py:21-25 / :61-64) for the aligned
// case (C % 4 == 0, 16-byte aligned rows: every activation matrix torch hands us at the model's widths).
//
// The first versions (kept below in gnn_kernels.cu for unaligned shapes) used scalar loads and, in the backward, held
// xhat / g / d-elu AND the per-lane column partials in registers: 221 registers, 8 warps per SM, 6x off the HBM roof
// (148 us for a 16 K x 1024 layer whose three streams take ~35 us).  Here: float4 loads and stores; the backward keeps
// only elu(x) and dy of its row in registers (everything else is recomputed from them) and accumulates the column sums
// of dscale / doffset in per-warp shared-memory slices - private to a warp, so still no atomics and a fixed reduction
// order (reproducible bits).
//
// Included in the middle of gnn_kernels.cu (inside its anonymous namespace).
#pragma once

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expf(v) - 1.f; }

template <int PER4>   // float4 per lane = ceil(C / 128)
__global__ void __launch_bounds__(256)
elu_rownorm_fwd_vec_kernel(const float *__restrict__ x, int64_t ldx, int M, int C, const float *__restrict__ scale,
                           const float *__restrict__ offset, float *__restrict__ y, int64_t ldy, float *__restrict__ mean_out,
                           float *__restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nv = C >> 2;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < M; r += gridDim.x * wpb) {
    const float4 *xr = reinterpret_cast<const float4 *>(x + (int64_t)r * ldx);
    float4 o[PER4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER4; ++j) {
      const int c = j * 32 + lane;
      o[j] = c < nv ? __ldcs(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < PER4; ++j) {
      if (j * 32 + lane < nv) {
        o[j].x = elu1(o[j].x); o[j].y = elu1(o[j].y); o[j].z = elu1(o[j].z); o[j].w = elu1(o[j].w);
        s += (o[j].x + o[j].y) + (o[j].z + o[j].w);
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < PER4; ++j) {
      if (j * 32 + lane < nv) {
        const float a = o[j].x - mean, b = o[j].y - mean, c2 = o[j].z - mean, d = o[j].w - mean;
        q += (a * a + b * b) + (c2 * c2 + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-9f);
    float4 *yr = reinterpret_cast<float4 *>(y + (int64_t)r * ldy);
#pragma unroll
    for (int j = 0; j < PER4; ++j) {
      const int c = j * 32 + lane;
      if (c < nv) {
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + c), of = __ldg(reinterpret_cast<const float4 *>(offset) + c);
        float4 t;
        t.x = (o[j].x - mean) * sc.x * rstd + of.x; t.y = (o[j].y - mean) * sc.y * rstd + of.y;
        t.z = (o[j].z - mean) * sc.z * rstd + of.z; t.w = (o[j].w - mean) * sc.w * rstd + of.w;
        yr[c] = t;
      }
    }
    if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
  }
}

// dx for every row; per-CTA partial column sums of dscale / doffset (reduced in fixed order by column_partials_reduce_kernel)
template <int PER4>
__global__ void __launch_bounds__(256, 2)
elu_rownorm_bwd_vec_kernel(const float *__restrict__ dy, int64_t lddy, const float *__restrict__ x, int64_t ldx, int M, int C,
                           const float *__restrict__ scale, const float *__restrict__ mean_in, const float *__restrict__ rstd_in,
                           float *__restrict__ dx, int64_t lddx, float *__restrict__ part_scale, float *__restrict__ part_offset) {
  extern __shared__ float4 sm4[];                  // [2][wpb][nv]: this warp's running column sums (dscale, doffset)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int nv = C >> 2;
  float4 *ss = sm4 + (size_t)warp * nv, *sb = sm4 + (size_t)(wpb + warp) * nv;
  for (int c = lane; c < nv; c += 32) { ss[c] = make_float4(0.f, 0.f, 0.f, 0.f); sb[c] = make_float4(0.f, 0.f, 0.f, 0.f); }
  __syncwarp();
  for (int r = blockIdx.x * wpb + warp; r < M; r += gridDim.x * wpb) {
    const float4 *xr = reinterpret_cast<const float4 *>(x + (int64_t)r * ldx);
    const float4 *gr = reinterpret_cast<const float4 *>(dy + (int64_t)r * lddy);
    const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
    float4 e[PER4], g[PER4];
#pragma unroll
    for (int j = 0; j < PER4; ++j) {
      const int c = j * 32 + lane;
      e[j] = c < nv ? __ldcs(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[j] = c < nv ? __ldcs(gr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < PER4; ++j) {
      const int c = j * 32 + lane;
      if (c < nv) {
        e[j].x = elu1(e[j].x); e[j].y = elu1(e[j].y); e[j].z = elu1(e[j].z); e[j].w = elu1(e[j].w);
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + c);
        const float hx = (e[j].x - mean) * rstd, hy = (e[j].y - mean) * rstd, hz = (e[j].z - mean) * rstd, hw = (e[j].w - mean) * rstd;
        float4 a = ss[c], b = sb[c];
        a.x += g[j].x * hx; a.y += g[j].y * hy; a.z += g[j].z * hz; a.w += g[j].w * hw;      // d loss / d scale
        b.x += g[j].x; b.y += g[j].y; b.z += g[j].z; b.w += g[j].w;                          // d loss / d offset
        ss[c] = a; sb[c] = b;
        const float gx = g[j].x * sc.x, gy = g[j].y * sc.y, gz = g[j].z * sc.z, gw = g[j].w * sc.w;   // d loss / d xhat
        s1 += (gx + gy) + (gz + gw);
        s2 += (gx * hx + gy * hy) + (gz * hz + gw * hw);
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    float4 *dr = reinterpret_cast<float4 *>(dx + (int64_t)r * lddx);
#pragma unroll
    for (int j = 0; j < PER4; ++j) {
      const int c = j * 32 + lane;
      if (c < nv) {
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + c);
        float4 t;
        // elu(v) > 0  <=>  v > 0: d elu / dv = 1 there, elu(v) + 1 = exp(v) elsewhere
        t.x = rstd * (g[j].x * sc.x - s1 - (e[j].x - mean) * rstd * s2) * (e[j].x > 0.f ? 1.f : e[j].x + 1.f);
        t.y = rstd * (g[j].y * sc.y - s1 - (e[j].y - mean) * rstd * s2) * (e[j].y > 0.f ? 1.f : e[j].y + 1.f);
        t.z = rstd * (g[j].z * sc.z - s1 - (e[j].z - mean) * rstd * s2) * (e[j].z > 0.f ? 1.f : e[j].z + 1.f);
        t.w = rstd * (g[j].w * sc.w - s1 - (e[j].w - mean) * rstd * s2) * (e[j].w > 0.f ? 1.f : e[j].w + 1.f);
        dr[c] = t;
      }
    }
  }
  __syncthreads();
  // CTA-level reduction of the warps' slices in fixed warp order
  const float *fs = reinterpret_cast<const float *>(sm4), *fb = fs + (size_t)wpb * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < wpb; ++w) { a += fs[(size_t)w * C + c]; b += fb[(size_t)w * C + c]; }
    part_scale[(size_t)blockIdx.x * C + c] = a;
    part_offset[(size_t)blockIdx.x * C + c] = b;
  }
}
