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

// ---- vectorised variants (D % 8 == 0): a lane owns 8 consecutive features per 256-feature step (one 16-byte load of a bf16
// tensor, two of an fp32 tensor), NV = ceil(D / 256) steps.  The scalar kernels above issue D / 32 two- or four-byte accesses per
// lane and tensor; at D = 224 that is 7 instructions where one does (lanes 28..31 idle).
template <typename T> __device__ __forceinline__ void ld8(const void* p, size_t i, float* f);
template <> __device__ __forceinline__ void ld8<float>(const void* p, size_t i, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i), b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const void* p, size_t i, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w[j] << 16); f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
}
template <typename T> __device__ __forceinline__ void st8(void* p, size_t i, const float* f);
template <> __device__ __forceinline__ void st8<float>(void* p, size_t i, const float* f) {
  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(void* p, size_t i, const float* f) {
  uint4 u;
  const __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]), c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  u.x = *reinterpret_cast<const uint32_t*>(&a); u.y = *reinterpret_cast<const uint32_t*>(&b); u.z = *reinterpret_cast<const uint32_t*>(&c); u.w = *reinterpret_cast<const uint32_t*>(&d);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = u;
}

template <int NV, typename TR, typename TY>
static __global__ void __launch_bounds__(kWarps * 32) fwd_kernel_v(MopLnParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D;
  float g[NV][8], b[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int d0 = 8 * lane + 256 * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) { g[i][e] = d0 < D ? p.gamma[d0 + e] : 0.f; b[i][e] = d0 < D ? p.beta[d0 + e] : 0.f; }
  }
  const float inv_d = 1.f / (float)D;
  for (int row = blockIdx.x * kWarps + warp; row < p.rows; row += gridDim.x * kWarps) {
    const size_t base = (size_t)row * D;
    const float sc = (p.r && p.scale) ? p.scale[row / p.rows_per_sample] : 1.f;
    float x[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d0 = 8 * lane + 256 * i;
#pragma unroll
      for (int e = 0; e < 8; ++e) x[i][e] = 0.f;
      if (d0 < D) {
        ld8<float>(p.x, base + d0, x[i]);
        if (p.r) {
          float rr[8];
          ld8<TR>(p.r, base + d0, rr);
#pragma unroll
          for (int e = 0; e < 8; ++e) x[i][e] = fmaf(sc, rr[e], x[i][e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += x[i][e];
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d0 = 8 * lane + 256 * i;
      if (d0 < D) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float c = x[i][e] - mean; q = fmaf(c, c, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + p.eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d0 = 8 * lane + 256 * i;
      if (d0 < D) {
        if (p.r) st8<float>(p.x_new, base + d0, x[i]);
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = fmaf((x[i][e] - mean) * rstd, g[i][e], b[i][e]);
        st8<TY>(p.y, base + d0, y);
      }
    }
    if (lane == 0) { p.mean[row] = mean; p.rstd[row] = rstd; }
  }
}

// KW warps per CTA: the fewer CTAs, the fewer partial rows the caller has to sum afterwards (that sum cost as much as this kernel
// with 1184 CTAs of 4 warps at the bench shape); 16 warps for one 256-feature step, fewer where the staging array would not fit
template <int NV> struct BwdWarps { static constexpr int value = NV == 1 ? 16 : NV == 2 ? 8 : 4; };
template <int NV, typename TR, typename TY>
static __global__ void __launch_bounds__(BwdWarps<NV>::value * 32) bwd_kernel_v(MopLnParams p) {
  constexpr int kWarps = BwdWarps<NV>::value;   // (shadows the namespace constant inside this kernel)
  __shared__ float red[2][kWarps][NV * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D;
  float g[NV][8], dg[NV][8], db[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int d0 = 8 * lane + 256 * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) { g[i][e] = d0 < D ? p.gamma[d0 + e] : 0.f; dg[i][e] = 0.f; db[i][e] = 0.f; }
  }
  const float inv_d = 1.f / (float)D;
  const void* xs = p.r ? p.x_new : p.x;   // the tensor that was normalised
  for (int row = blockIdx.x * kWarps + warp; row < p.rows; row += gridDim.x * kWarps) {
    const size_t base = (size_t)row * D;
    const float mean = p.mean[row], rstd = p.rstd[row];
    float xh[NV][8], gy[NV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d0 = 8 * lane + 256 * i;
      float xv[8], dyv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { xv[e] = 0.f; dyv[e] = 0.f; }
      if (d0 < D) {
        ld8<float>(xs, base + d0, xv);
        ld8<TY>(p.dy, base + d0, dyv);
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[e] = (xv[e] - mean) * rstd;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        xh[i][e] = xv[e];
        dg[i][e] = fmaf(dyv[e], xv[e], dg[i][e]);
        db[i][e] += dyv[e];
        gy[i][e] = dyv[e] * g[i][e];
        c1 += gy[i][e];
        c2 = fmaf(gy[i][e], xv[e], c2);
      }
    }
    c1 = warp_sum(c1) * inv_d;
    c2 = warp_sum(c2) * inv_d;
    const float sc = (p.dr && p.scale) ? p.scale[row / p.rows_per_sample] : 1.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d0 = 8 * lane + 256 * i;
      if (d0 < D) {
        float dx[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dx[e] = (gy[i][e] - c1 - xh[i][e] * c2) * rstd;
        if (p.dx_new) {
          float dn[8];
          ld8<float>(p.dx_new, base + d0, dn);
#pragma unroll
          for (int e = 0; e < 8; ++e) dx[e] += dn[e];
        }
        st8<float>(p.dx, base + d0, dx);
        if (p.dr) {
#pragma unroll
          for (int e = 0; e < 8; ++e) dx[e] *= sc;
          st8<TR>(p.dr, base + d0, dx);
        }
      }
    }
  }
  // per-CTA partials of dgamma / dbeta (fixed order: deterministic)
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) { red[0][warp][8 * lane + 256 * i + e] = dg[i][e]; red[1][warp][8 * lane + 256 * i + e] = db[i][e]; }
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
// CTAs (= partial rows) of the vectorised backward for feature count D: the same number of resident warps per SM as above
inline int bwd_grid_size_v(int rows, int D, int sms) {
  const int nv = (D + 255) / 256, kw = nv == 1 ? 16 : nv == 2 ? 8 : 4;
  const int want = (rows + kw - 1) / kw, cap = sms * 32 / kw;
  return want < cap ? want : cap;
}

}  // namespace ln
}  // namespace mop
