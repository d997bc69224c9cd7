// Fused residual-add + per-sample DropPath scale + LayerNorm, forward and backward (the block epilogues on
// either side of the attention call).  Reference sites: BlockEdgewise.forward experiments/cifar100_edgewise_gates.py:371-374
// (`x = x + dp1(attn(ln1(x)))`, `x = x + dp2(mlp(ln2(x)))`), DropPath mop/models/components.py:14-27, nn.LayerNorm (eps 1e-5).
//
//   forward :  x_new = x + scale[b] * r      (r optional; b = row / rows_per_sample)        -> x_new (fp32)
//              y     = (x_new - mean) * rstd * gamma + beta                                   -> y (fp32 | bf16), mean, rstd
//   backward:  dx = LN'(dy) + dx_new   (dx_new optional: gradient that reaches x_new through the residual stream)
//              dr = scale[b] * dx      (type of r),   dgamma / dbeta partials per CTA (summed by the caller: deterministic)
//
// HBM-bound, one warp per row (PL = ceil(D / 32) values per lane in registers, two-pass variance), rows strided over a
// persistent grid.  Bytes per row (D = 224, bf16 branch, bf16 y): forward 4D + 2D + 4D + 2D = 12 D, backward 2D + 4D + 4D + 4D + 2D = 16 D.
#pragma once
#include "common.cuh"

namespace mop {
namespace ln {

constexpr int kWarps = 4;

template <typename T> __device__ __forceinline__ float ldf(const void* p, size_t i) { return to_f32<T>(reinterpret_cast<const T*>(p)[i]); }
template <typename T> __device__ __forceinline__ void stf(void* p, size_t i, float v) { reinterpret_cast<T*>(p)[i] = from_f32<T>(v); }

// TR: type of the branch r / dr, TY: type of y / dy
template <int PL, typename TR, typename TY>
static __global__ void __launch_bounds__(kWarps * 32) fwd_kernel(MopLnParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D;
  float g[PL], b[PL];
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int d = lane + 32 * i;
    g[i] = d < D ? p.gamma[d] : 0.f;
    b[i] = d < D ? p.beta[d] : 0.f;
  }
  const float inv_d = 1.f / (float)D;
  for (int row = blockIdx.x * kWarps + warp; row < p.rows; row += gridDim.x * kWarps) {
    const size_t base = (size_t)row * D;
    const float sc = (p.r && p.scale) ? p.scale[row / p.rows_per_sample] : 1.f;
    float x[PL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int d = lane + 32 * i;
      float v = 0.f;
      if (d < D) {
        v = reinterpret_cast<const float*>(p.x)[base + d];
        if (p.r) v = fmaf(sc, ldf<TR>(p.r, base + d), v);
      }
      x[i] = v;
      s += v;
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int d = lane + 32 * i;
      const float c = d < D ? x[i] - mean : 0.f;
      q = fmaf(c, c, q);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + p.eps);
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int d = lane + 32 * i;
      if (d < D) {
        if (p.r) reinterpret_cast<float*>(p.x_new)[base + d] = x[i];
        stf<TY>(p.y, base + d, fmaf((x[i] - mean) * rstd, g[i], b[i]));
      }
    }
    if (lane == 0) { p.mean[row] = mean; p.rstd[row] = rstd; }
  }
}

template <int PL, typename TR, typename TY>
static __global__ void __launch_bounds__(kWarps * 32) bwd_kernel(MopLnParams p) {
  __shared__ float red[2][kWarps][PL * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D;
  float g[PL], dg[PL], db[PL];
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int d = lane + 32 * i;
    g[i] = d < D ? p.gamma[d] : 0.f;
    dg[i] = 0.f;
    db[i] = 0.f;
  }
  const float inv_d = 1.f / (float)D;
  const float* xs = reinterpret_cast<const float*>(p.r ? p.x_new : p.x);   // the tensor that was normalised
  for (int row = blockIdx.x * kWarps + warp; row < p.rows; row += gridDim.x * kWarps) {
    const size_t base = (size_t)row * D;
    const float mean = p.mean[row], rstd = p.rstd[row];
    float xh[PL], gy[PL];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int d = lane + 32 * i;
      float xv = 0.f, dyv = 0.f;
      if (d < D) { xv = (xs[base + d] - mean) * rstd; dyv = ldf<TY>(p.dy, base + d); }
      xh[i] = xv;
      dg[i] = fmaf(dyv, xv, dg[i]);
      db[i] += dyv;
      gy[i] = dyv * g[i];
      c1 += gy[i];
      c2 = fmaf(gy[i], xv, c2);
    }
    c1 = warp_sum(c1) * inv_d;
    c2 = warp_sum(c2) * inv_d;
    const float sc = (p.dr && p.scale) ? p.scale[row / p.rows_per_sample] : 1.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int d = lane + 32 * i;
      if (d < D) {
        float dx = (gy[i] - c1 - xh[i] * c2) * rstd;
        if (p.dx_new) dx += reinterpret_cast<const float*>(p.dx_new)[base + d];
        reinterpret_cast<float*>(p.dx)[base + d] = dx;
        if (p.dr) stf<TR>(p.dr, base + d, sc * dx);
      }
    }
  }
  // per-CTA partials of dgamma / dbeta (fixed order: deterministic)
#pragma unroll
  for (int i = 0; i < PL; ++i) { red[0][warp][lane + 32 * i] = dg[i]; red[1][warp][lane + 32 * i] = db[i]; }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += kWarps * 32) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { a += red[0][w][d]; c += red[1][w][d]; }
    p.dgamma_part[(size_t)blockIdx.x * D + d] = a;
    p.dbeta_part[(size_t)blockIdx.x * D + d] = c;
  }
}

inline int grid_size(int rows, int sms) {
  const int want = (rows + kWarps - 1) / kWarps, cap = sms * 8;
  return want < cap ? want : cap;
}

}  // namespace ln
}  // namespace mop
