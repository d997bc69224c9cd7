// ViT-MoP post-encoder token gate, fused.  Reference mop/models/vit_mop.py:95-114 with
// mop/models/components.py:255-303:
//   views = ViewsLinear(tok)            [B, V, Gh, Gw]   (D -> V per token, no bias)
//   kmaps = Kernels3(views)             conv3x3(V -> 16, zero pad 1, no bias) -> SiLU -> conv1x1(16 -> K, no bias)
//   G     = FuseExcInh([views; kmaps])  conv1x1(V+K -> hid, no bias) -> SiLU -> conv1x1(hid -> 2, bias)
//   gate  = 1 + a_pos sigmoid(G_0) - a_neg sigmoid(G_1)          (a = softplus(alpha), applied by the caller)
//   out   = tok * gate (per token)
// One CTA of 256 threads per image (T = Gh Gw <= 256 tokens): the views, the 16-channel hidden map and every other
// intermediate live in shared memory / registers; HBM traffic is tok in, out out (+ views, gate: B T (V+1) floats saved for the
// backward).  The backward recomputes the tiny network per pixel, keeps the per-pixel vectors in shared memory, and forms
// every parameter gradient as a per-CTA partial (one thread per parameter, accumulated over the images the CTA owns).
#pragma once
#include "common.cuh"
#include "../../include/mop_b200.h"

namespace mop {
namespace tokgate {

constexpr int kMaxT = 256;      // tokens per image
constexpr int kMaxV = 8;        // views
constexpr int kMaxK = 8;        // kernel maps
constexpr int kHidK = 16;       // hidden channels of Kernels3 (fixed in the reference)
constexpr int kMaxHid = 16;     // hidden channels of FuseExcInh (max(8, V + K))
constexpr int kThreads = 256;

struct Dims {
  int V, K, hid, C;             // C = V + K
  int o_k3, o_k1, o_f1, o_f2, o_b2, nnet;   // offsets into the packed network parameters
};
__host__ __device__ inline Dims dims(const MopTokenGateParams& p) {
  Dims d;
  d.V = p.V; d.K = p.K; d.hid = p.hid; d.C = p.V + p.K;
  d.o_k3 = 0;
  d.o_k1 = d.o_k3 + kHidK * d.V * 9;
  d.o_f1 = d.o_k1 + d.K * kHidK;
  d.o_f2 = d.o_f1 + d.hid * d.C;
  d.o_b2 = d.o_f2 + 2 * d.hid;
  d.nnet = d.o_b2 + 2;
  return d;
}
// shared memory (floats): wv [V][D] | net [nnet] | per-pixel rows of T floats each
inline size_t smem_floats_fwd(const MopTokenGateParams& p) { const Dims d = dims(p); return (size_t)p.V * p.D + d.nnet + (size_t)(d.V + 1) * p.T; }
inline size_t smem_floats_bwd(const MopTokenGateParams& p) {
  const Dims d = dims(p);
  // sv[V], gate, dg, dz[2], shf[hid], dhf[hid], km[K], dkm[K], shk[16], dhk[16], dvw[V]
  return (size_t)p.V * p.D + d.nnet + (size_t)(2 * d.V + 3 + 2 + 2 * d.hid + 2 * d.K + 2 * kHidK) * p.T;
}

__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_grad(float x) { const float s = 1.f / (1.f + __expf(-x)); return s * (1.f + x * (1.f - s)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

template <typename T> __device__ __forceinline__ void load8(const T* p, float* f);
template <> __device__ __forceinline__ void load8<float>(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* f);
template <> __device__ __forceinline__ void store8<float>(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]), c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b); u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}

// weights -> shared memory (once per CTA)
__device__ inline void stage_weights(const MopTokenGateParams& p, const Dims& d, float* wv_s, float* net) {
  for (int i = threadIdx.x; i < p.V * p.D; i += kThreads) wv_s[i] = p.views_w[i];
  for (int i = threadIdx.x; i < d.nnet; i += kThreads) {
    float v;
    if (i < d.o_k1) v = p.k3_w[i];
    else if (i < d.o_f1) v = p.k1_w[i - d.o_k1];
    else if (i < d.o_f2) v = p.f1_w[i - d.o_f1];
    else if (i < d.o_b2) v = p.f2_w[i - d.o_f2];
    else v = p.f2_b[i - d.o_b2];
    net[i] = v;
  }
}

// views of the tokens of one image: warp per token, lanes over the features (8 per lane and step)
template <typename T>
__device__ inline void project_tokens(const MopTokenGateParams& p, const T* x, const float* wv_s, float* sv, float* views_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, V = p.V, D = p.D;
  for (int t = warp; t < p.T; t += kThreads / 32) {
    float acc[kMaxV];
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) acc[v] = 0.f;
    for (int d0 = lane * 8; d0 < D; d0 += 256) {
      float f[8];
      load8<T>(x + (size_t)t * D + d0, f);
#pragma unroll
      for (int v = 0; v < kMaxV; ++v)
        if (v < V) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[v] = fmaf(f[e], wv_s[v * D + d0 + e], acc[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < kMaxV; ++v)
      if (v < V) {
        const float s = warp_sum(acc[v]);
        if (lane == 0) { sv[v * p.T + t] = s; if (views_out) views_out[(size_t)t * V + v] = s; }
      }
  }
}

// the network at pixel t: hidden pre-activations hk[16] (Kernels3), kernel maps km[K], fuse pre-activations hf[hid], z[2]
__device__ __forceinline__ void net_forward(const MopTokenGateParams& p, const Dims& d, const float* net, const float* sv, int t,
                                            float* hk, float* km, float* hf, float* z) {
  const int T = p.T, Gw = p.Gw, Gh = p.Gh, i = t / Gw, j = t % Gw;
#pragma unroll
  for (int o = 0; o < kHidK; ++o) hk[o] = 0.f;
  for (int v = 0; v < d.V; ++v)
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int ii = i + u - 1;
      if (ii < 0 || ii >= Gh) continue;
#pragma unroll
      for (int w = 0; w < 3; ++w) {
        const int jj = j + w - 1;
        if (jj < 0 || jj >= Gw) continue;
        const float xv = sv[v * T + ii * Gw + jj];
#pragma unroll
        for (int o = 0; o < kHidK; ++o) hk[o] = fmaf(net[d.o_k3 + (o * d.V + v) * 9 + u * 3 + w], xv, hk[o]);
      }
    }
#pragma unroll
  for (int k = 0; k < kMaxK; ++k) {
    float a = 0.f;
    if (k < d.K) {
#pragma unroll
      for (int o = 0; o < kHidK; ++o) a = fmaf(net[d.o_k1 + k * kHidK + o], silu(hk[o]), a);
    }
    km[k] = a;
  }
#pragma unroll
  for (int h = 0; h < kMaxHid; ++h) {
    float a = 0.f;
    if (h < d.hid) {
      for (int c = 0; c < d.V; ++c) a = fmaf(net[d.o_f1 + h * d.C + c], sv[c * T + t], a);
#pragma unroll
      for (int k = 0; k < kMaxK; ++k)
        if (k < d.K) a = fmaf(net[d.o_f1 + h * d.C + d.V + k], km[k], a);
    }
    hf[h] = a;
  }
  z[0] = net[d.o_b2];
  z[1] = net[d.o_b2 + 1];
#pragma unroll
  for (int h = 0; h < kMaxHid; ++h)
    if (h < d.hid) {
      const float s = silu(hf[h]);
      z[0] = fmaf(net[d.o_f2 + h], s, z[0]);
      z[1] = fmaf(net[d.o_f2 + d.hid + h], s, z[1]);
    }
}

// grid: min(B, 2 * SMs) CTAs of 256 threads, dynamic shared memory smem_floats_fwd * 4
template <typename T>
static __global__ void __launch_bounds__(kThreads) fwd_kernel(MopTokenGateParams p) {
  extern __shared__ __align__(16) float smf[];
  const Dims d = dims(p);
  float* wv_s = smf;
  float* net = wv_s + p.V * p.D;
  float* sv = net + d.nnet;            // [V][T]
  float* gate_s = sv + d.V * p.T;      // [T]
  stage_weights(p, d, wv_s, net);
  __syncthreads();
  const float a_pos = p.a_pos[0], a_neg = p.a_neg[0];
  const int Tn = p.T, D = p.D;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const T* x = reinterpret_cast<const T*>(p.x) + (size_t)b * Tn * D;
    T* out = reinterpret_cast<T*>(p.out) + (size_t)b * Tn * D;
    project_tokens<T>(p, x, wv_s, sv, p.views + (size_t)b * Tn * d.V);
    __syncthreads();
    if (threadIdx.x < Tn) {
      float hk[kHidK], km[kMaxK], hf[kMaxHid], z[2];
      net_forward(p, d, net, sv, threadIdx.x, hk, km, hf, z);
      const float g = 1.f + a_pos * sigmoidf_(z[0]) - a_neg * sigmoidf_(z[1]);
      gate_s[threadIdx.x] = g;
      p.gate[(size_t)b * Tn + threadIdx.x] = g;
    }
    __syncthreads();
    const int nc = D / 8;
#pragma unroll 4
    for (int item = threadIdx.x; item < Tn * nc; item += kThreads) {
      const int t = item / nc, c = item % nc;
      float f[8];
      load8<T>(x + (size_t)t * D + 8 * c, f);
      const float g = gate_s[t];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= g;
      store8<T>(out + (size_t)t * D + 8 * c, f);
    }
    __syncthreads();
  }
}

// number of token groups the dx / dwv phase of the backward splits the tokens of an image into (one partial row of dwv each)
// (D / 8 <= 256 chunks: every thread of a group owns one 8-feature chunk)
__host__ __device__ inline int wv_groups(int D) { return kThreads / (D / 8); }

// grid: nparts CTAs; partial rows: dnet_part [nparts][nnet + 2] (network parameters, then d a_pos, d a_neg),
// dwv_part [nparts * wv_groups(D)][V][D]
template <typename T>
static __global__ void __launch_bounds__(kThreads) bwd_kernel(MopTokenGateParams p) {
  extern __shared__ __align__(16) float smf[];
  const Dims d = dims(p);
  const int Tn = p.T, D = p.D, V = d.V, K = d.K, hid = d.hid, tid = threadIdx.x;
  float* wv_s = smf;
  float* net = wv_s + V * D;
  float* sv = net + d.nnet;            // [V][T]   views (maps 0..V-1)
  float* gate_s = sv + V * Tn;         // [T]
  float* dg = gate_s + Tn;             // [T]      d gate
  float* dz = dg + Tn;                 // [2][T]
  float* shf = dz + 2 * Tn;            // [hid][T] silu(hf)
  float* dhf = shf + hid * Tn;         // [hid][T]
  float* kmS = dhf + hid * Tn;         // [K][T]   kernel maps (maps V..V+K-1)
  float* dkm = kmS + K * Tn;           // [K][T]
  float* shk = dkm + K * Tn;           // [16][T]  silu(hk)
  float* dhk = shk + kHidK * Tn;       // [16][T]
  float* dvw = dhk + kHidK * Tn;       // [V][T]   d views
  stage_weights(p, d, wv_s, net);
  __syncthreads();
  const float a_pos = p.a_pos[0], a_neg = p.a_neg[0];
  // this thread's network parameters (fixed for the whole launch) and its dwv accumulators
  constexpr int kPer = (kHidK * kMaxV * 9 + kMaxK * kHidK + kMaxHid * (kMaxV + kMaxK) + 2 * kMaxHid + 2 + kThreads - 1) / kThreads;
  float pacc[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) pacc[j] = 0.f;
  float dap = 0.f, dan = 0.f;
  const int nc = D / 8, groups = wv_groups(D);
  const int wc = tid % nc, wg = tid / nc;
  const bool w_on = wg < groups;
  float wacc[kMaxV][8];
#pragma unroll
  for (int v = 0; v < kMaxV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) wacc[v][e] = 0.f;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const T* x = reinterpret_cast<const T*>(p.x) + (size_t)b * Tn * D;
    const T* dout = reinterpret_cast<const T*>(p.dout) + (size_t)b * Tn * D;
    T* dx = reinterpret_cast<T*>(p.dx) + (size_t)b * Tn * D;
    for (int i = tid; i < Tn * V; i += kThreads) sv[(i % V) * Tn + i / V] = p.views[(size_t)b * Tn * V + i];
    for (int i = tid; i < Tn; i += kThreads) gate_s[i] = p.gate[(size_t)b * Tn + i];
    // d gate[t] = dout[t] . x[t]: warp per token
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int t = warp; t < Tn; t += kThreads / 32) {
        float a = 0.f;
        for (int d0 = lane * 8; d0 < D; d0 += 256) {
          float f[8], g[8];
          load8<T>(x + (size_t)t * D + d0, f);
          load8<T>(dout + (size_t)t * D + d0, g);
#pragma unroll
          for (int e = 0; e < 8; ++e) a = fmaf(f[e], g[e], a);
        }
        a = warp_sum(a);
        if (lane == 0) dg[t] = a;
      }
    }
    __syncthreads();
    // per pixel: forward recompute, backward down to the hidden map of Kernels3 and the direct part of d views
    if (tid < Tn) {
      const int t = tid;
      float hk[kHidK], km[kMaxK], hf[kMaxHid], z[2];
      net_forward(p, d, net, sv, t, hk, km, hf, z);
      const float gp = sigmoidf_(z[0]), gn = sigmoidf_(z[1]), g = dg[t];
      dap = fmaf(g, gp, dap);
      dan = fmaf(-g, gn, dan);
      const float dz0 = g * a_pos * gp * (1.f - gp), dz1 = -g * a_neg * gn * (1.f - gn);
      dz[t] = dz0;
      dz[Tn + t] = dz1;
      float dmaps[kMaxV + kMaxK];
#pragma unroll
      for (int c = 0; c < kMaxV + kMaxK; ++c) dmaps[c] = 0.f;
#pragma unroll
      for (int h = 0; h < kMaxHid; ++h)
        if (h < hid) {
          const float dsh = net[d.o_f2 + h] * dz0 + net[d.o_f2 + hid + h] * dz1;
          const float dh = dsh * silu_grad(hf[h]);
          shf[h * Tn + t] = silu(hf[h]);
          dhf[h * Tn + t] = dh;
#pragma unroll
          for (int c = 0; c < kMaxV + kMaxK; ++c)
            if (c < d.C) dmaps[c] = fmaf(net[d.o_f1 + h * d.C + c], dh, dmaps[c]);
        }
#pragma unroll
      for (int v = 0; v < kMaxV; ++v)
        if (v < V) dvw[v * Tn + t] = dmaps[v];
      float dsk[kHidK];
#pragma unroll
      for (int o = 0; o < kHidK; ++o) dsk[o] = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k)
        if (k < K) {
          // dmaps is indexed at run time here (V varies): select instead of indexing a register array
          float dk = 0.f;
#pragma unroll
          for (int c = 0; c < kMaxV + kMaxK; ++c)
            if (c == V + k) dk = dmaps[c];
          kmS[k * Tn + t] = km[k];
          dkm[k * Tn + t] = dk;
#pragma unroll
          for (int o = 0; o < kHidK; ++o) dsk[o] = fmaf(net[d.o_k1 + k * kHidK + o], dk, dsk[o]);
        }
#pragma unroll
      for (int o = 0; o < kHidK; ++o) {
        shk[o * Tn + t] = silu(hk[o]);
        dhk[o * Tn + t] = dsk[o] * silu_grad(hk[o]);
      }
    }
    __syncthreads();
    // d views through the transposed 3x3 convolution
    if (tid < Tn) {
      const int t = tid, Gw = p.Gw, Gh = p.Gh, i = t / Gw, j = t % Gw;
      for (int v = 0; v < V; ++v) {
        float a = dvw[v * Tn + t];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int ii = i - (u - 1);
          if (ii < 0 || ii >= Gh) continue;
#pragma unroll
          for (int w = 0; w < 3; ++w) {
            const int jj = j - (w - 1);
            if (jj < 0 || jj >= Gw) continue;
            const int q = ii * Gw + jj;
#pragma unroll
            for (int o = 0; o < kHidK; ++o) a = fmaf(net[d.o_k3 + (o * V + v) * 9 + u * 3 + w], dhk[o * Tn + q], a);
          }
        }
        dvw[v * Tn + t] = a;   // only this thread reads / writes its own pixel of dvw in this phase
      }
    }
    // network parameter gradients: one thread per parameter, sum over the pixels of this image
#pragma unroll
    for (int jx = 0; jx < kPer; ++jx) {
      const int idx = tid + jx * kThreads;
      if (idx >= d.nnet) break;
      float a = 0.f;
      if (idx < d.o_k1) {   // k3_w[o][v][u][w]: sum_p dhk[o][p] views[v][p + (u-1, w-1)]
        const int o = idx / (V * 9), rem = idx % (V * 9), v = rem / 9, u = (rem % 9) / 3, w = rem % 3;
        const int Gw = p.Gw, Gh = p.Gh;
        for (int i = 0; i < Gh; ++i) {
          const int ii = i + u - 1;
          if (ii < 0 || ii >= Gh) continue;
          for (int j = 0; j < Gw; ++j) {
            const int jj = j + w - 1;
            if (jj < 0 || jj >= Gw) continue;
            a = fmaf(dhk[o * Tn + i * Gw + j], sv[v * Tn + ii * Gw + jj], a);
          }
        }
      } else if (idx < d.o_f1) {   // k1_w[k][o]
        const int k = (idx - d.o_k1) / kHidK, o = (idx - d.o_k1) % kHidK;
        for (int q = 0; q < Tn; ++q) a = fmaf(dkm[k * Tn + q], shk[o * Tn + q], a);
      } else if (idx < d.o_f2) {   // f1_w[h][c]
        const int h = (idx - d.o_f1) / d.C, c = (idx - d.o_f1) % d.C;
        const float* m = c < V ? sv + c * Tn : kmS + (c - V) * Tn;
        for (int q = 0; q < Tn; ++q) a = fmaf(dhf[h * Tn + q], m[q], a);
      } else if (idx < d.o_b2) {   // f2_w[s][h]
        const int s = (idx - d.o_f2) / hid, h = (idx - d.o_f2) % hid;
        for (int q = 0; q < Tn; ++q) a = fmaf(dz[s * Tn + q], shf[h * Tn + q], a);
      } else {                      // f2_b[s]
        const int s = idx - d.o_b2;
        for (int q = 0; q < Tn; ++q) a += dz[s * Tn + q];
      }
      pacc[jx] += a;
    }
    __syncthreads();
    // dx = dout * gate + d views . Wv ;  d Wv += d views^T x   (thread = (8-feature chunk, token group))
    if (w_on) {
      const int c = wc;
      for (int tb = wg; tb < Tn; tb += 4 * groups) {   // four tokens per step: all loads in flight before the first use
        float f[4][8], g[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = tb + u * groups;
          if (t < Tn) {
            load8<T>(x + (size_t)t * D + 8 * c, f[u]);
            load8<T>(dout + (size_t)t * D + 8 * c, g[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = tb + u * groups;
          if (t < Tn) {
            float o[8];
            const float gt = gate_s[t];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = g[u][e] * gt;
#pragma unroll
            for (int v = 0; v < kMaxV; ++v)
              if (v < V) {
                const float dv = dvw[v * Tn + t];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  o[e] = fmaf(dv, wv_s[v * D + 8 * c + e], o[e]);
                  wacc[v][e] = fmaf(dv, f[u][e], wacc[v][e]);
                }
              }
            store8<T>(dx + (size_t)t * D + 8 * c, o);
          }
        }
      }
    }
    __syncthreads();
  }
  // partial rows
  float* dn = p.dnet_part + (size_t)blockIdx.x * (d.nnet + 2);
#pragma unroll
  for (int jx = 0; jx < kPer; ++jx) {
    const int idx = tid + jx * kThreads;
    if (idx < d.nnet) dn[idx] = pacc[jx];
  }
  __shared__ float red[2][kThreads / 32];
  dap = warp_sum(dap);
  dan = warp_sum(dan);
  if ((tid & 31) == 0) { red[0][tid >> 5] = dap; red[1][tid >> 5] = dan; }
  __syncthreads();
  if (tid < 2) {
    float a = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) a += red[tid][w];
    dn[d.nnet + tid] = a;
  }
  if (w_on) {
    float* dw = p.dwv_part + (size_t)(blockIdx.x * groups + wg) * V * D;
#pragma unroll
    for (int v = 0; v < kMaxV; ++v)
      if (v < V) {
#pragma unroll
        for (int e = 0; e < 8; ++e) dw[(size_t)v * D + 8 * wc + e] = wacc[v][e];
      }
  }
}

}  // namespace tokgate

// GPT-MoP 1-D token gate (reference mop/models/gpt_mop.py:19-67, 102-123): views = ViewsLinear1D(x) (D -> V per token), kernel
// maps = Conv1d(V -> K, k = 3, padding 1), [g_pos, g_neg] = Conv1d 1x1([views; kernel maps]), gate = 1 + a_pos g_pos - a_neg g_neg,
// x * gate.  No nonlinearity anywhere: the caller folds the three weight tensors and alpha into ONE 3-tap filter of the views,
//   gate[t] = 1 + sum_{tau in -1..1} sum_v We[tau + 1][v] views[t + tau][v]     (zero outside the sequence),
// (tiny, differentiable PyTorch) and the kernels run per chunk of 64 tokens (+ one halo token each side, recomputed).
namespace tokgate1d {

using tokgate::kMaxV;
using tokgate::kThreads;
using tokgate::load8;
using tokgate::store8;
constexpr int kChunk = 64;
constexpr int kHalo = kChunk + 2;

inline size_t smem_floats(const MopTokenGate1dParams& p) { return (size_t)p.V * p.D + 3 * kMaxV + (size_t)(2 * kMaxV + 2) * kHalo; }
__host__ __device__ inline int wv_groups(int D) { return kThreads / (D / 8); }

// views of tokens t0 - 1 .. t0 + 64 of sequence b (zero outside [0, T)): warp per token
template <typename T>
__device__ inline void project_chunk(const MopTokenGate1dParams& p, const T* xseq, int t0, const float* wv_s, float* sv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, V = p.V, D = p.D;
  for (int i = warp; i < kHalo; i += kThreads / 32) {
    const int t = t0 - 1 + i;
    float acc[kMaxV];
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) acc[v] = 0.f;
    if (t >= 0 && t < p.T) {
      for (int d0 = lane * 8; d0 < D; d0 += 256) {
        float f[8];
        load8<T>(xseq + (size_t)t * D + d0, f);
#pragma unroll
        for (int v = 0; v < kMaxV; ++v)
          if (v < V) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[v] = fmaf(f[e], wv_s[v * D + d0 + e], acc[v]);
          }
      }
    }
#pragma unroll
    for (int v = 0; v < kMaxV; ++v)
      if (v < V) {
        const float s = warp_sum(acc[v]);
        if (lane == 0) sv[v * kHalo + i] = s;
      }
  }
}

// grid: min(chunks, 2 SMs) CTAs of 256 threads
template <typename T>
static __global__ void __launch_bounds__(kThreads) fwd_kernel(MopTokenGate1dParams p) {
  extern __shared__ __align__(16) float smf[];
  const int V = p.V, D = p.D, Tn = p.T, tid = threadIdx.x;
  float* wv_s = smf;
  float* we = wv_s + V * D;            // [3][kMaxV]
  float* sv = we + 3 * kMaxV;          // [V][66]
  float* gate_s = sv + kMaxV * kHalo;  // [64]
  for (int i = tid; i < V * D; i += kThreads) wv_s[i] = p.views_w[i];
  if (tid < 3 * kMaxV) we[tid] = (tid % kMaxV) < V ? p.w_eff[(tid / kMaxV) * V + tid % kMaxV] : 0.f;
  __syncthreads();
  const int cpt = (Tn + kChunk - 1) / kChunk, nchunks = p.B * cpt, nc = D / 8;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int b = c / cpt, t0 = (c % cpt) * kChunk;
    const T* xs = reinterpret_cast<const T*>(p.x) + (size_t)b * Tn * D;
    T* os = reinterpret_cast<T*>(p.out) + (size_t)b * Tn * D;
    project_chunk<T>(p, xs, t0, wv_s, sv);
    __syncthreads();
    if (tid < kChunk && t0 + tid < Tn) {
      float g = 1.f;
#pragma unroll
      for (int v = 0; v < kMaxV; ++v)
        if (v < V) {
          g = fmaf(we[v], sv[v * kHalo + tid], fmaf(we[kMaxV + v], sv[v * kHalo + tid + 1], fmaf(we[2 * kMaxV + v], sv[v * kHalo + tid + 2], g)));
          p.views[((size_t)b * Tn + t0 + tid) * V + v] = sv[v * kHalo + tid + 1];
        }
      gate_s[tid] = g;
      p.gate[(size_t)b * Tn + t0 + tid] = g;
    }
    __syncthreads();
    const int nt = min(kChunk, Tn - t0);
#pragma unroll 4
    for (int item = tid; item < nt * nc; item += kThreads) {
      const int t = item / nc, ch = item % nc;
      float f[8];
      load8<T>(xs + (size_t)(t0 + t) * D + 8 * ch, f);
      const float g = gate_s[t];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= g;
      store8<T>(os + (size_t)(t0 + t) * D + 8 * ch, f);
    }
    __syncthreads();
  }
}

// grid: nparts CTAs; partial rows: dweff_part [nparts][3 V], dwv_part [nparts * wv_groups(D)][V][D]
template <typename T>
static __global__ void __launch_bounds__(kThreads) bwd_kernel(MopTokenGate1dParams p) {
  extern __shared__ __align__(16) float smf[];
  const int V = p.V, D = p.D, Tn = p.T, tid = threadIdx.x;
  float* wv_s = smf;
  float* we = wv_s + V * D;
  float* sv = we + 3 * kMaxV;          // [V][66] saved views (zero outside the sequence)
  float* dvw = sv + kMaxV * kHalo;     // [V][66] d views (first 64 used)
  float* dg = dvw + kMaxV * kHalo;     // [66]    d gate of tokens t0 - 1 .. t0 + 64
  float* gate_s = dg + kHalo;          // [66]    (first 64 used)
  for (int i = tid; i < V * D; i += kThreads) wv_s[i] = p.views_w[i];
  if (tid < 3 * kMaxV) we[tid] = (tid % kMaxV) < V ? p.w_eff[(tid / kMaxV) * V + tid % kMaxV] : 0.f;
  __syncthreads();
  const int cpt = (Tn + kChunk - 1) / kChunk, nchunks = p.B * cpt, nc = D / 8, groups = wv_groups(D);
  const int wc = tid % nc, wg = tid / nc;
  const bool w_on = wg < groups;
  float wacc[kMaxV][8];
#pragma unroll
  for (int v = 0; v < kMaxV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) wacc[v][e] = 0.f;
  float eacc = 0.f;   // thread tau * kMaxV + v < 3 kMaxV: d We[tau][v]
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int b = c / cpt, t0 = (c % cpt) * kChunk;
    const T* xs = reinterpret_cast<const T*>(p.x) + (size_t)b * Tn * D;
    const T* ds = reinterpret_cast<const T*>(p.dout) + (size_t)b * Tn * D;
    T* dxs = reinterpret_cast<T*>(p.dx) + (size_t)b * Tn * D;
    for (int i = tid; i < kHalo * V; i += kThreads) {
      const int hi = i / V, v = i % V, t = t0 - 1 + hi;
      sv[v * kHalo + hi] = (t >= 0 && t < Tn) ? p.views[((size_t)b * Tn + t) * V + v] : 0.f;
    }
    if (tid < kChunk) gate_s[tid] = t0 + tid < Tn ? p.gate[(size_t)b * Tn + t0 + tid] : 0.f;
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int i = warp; i < kHalo; i += kThreads / 32) {
        const int t = t0 - 1 + i;
        float a = 0.f;
        if (t >= 0 && t < Tn) {
          for (int d0 = lane * 8; d0 < D; d0 += 256) {
            float f[8], g[8];
            load8<T>(xs + (size_t)t * D + d0, f);
            load8<T>(ds + (size_t)t * D + d0, g);
#pragma unroll
            for (int e = 0; e < 8; ++e) a = fmaf(f[e], g[e], a);
          }
        }
        a = warp_sum(a);
        if (lane == 0) dg[i] = a;
      }
    }
    __syncthreads();
    // d views[t][v] = sum_tau We[tau][v] d gate[t - tau]   (halo index of token t0 + i is i + 1)
    for (int i = tid; i < kChunk * V; i += kThreads) {
      const int t = i / V, v = i % V;
      dvw[v * kHalo + t] = we[v] * dg[t + 2] + we[kMaxV + v] * dg[t + 1] + we[2 * kMaxV + v] * dg[t];
    }
    // d We[tau][v] += sum over the chunk's tokens of d gate[t] views[t + tau][v]
    if (tid < 3 * kMaxV && (tid % kMaxV) < V) {
      const int tau = tid / kMaxV, v = tid % kMaxV, nt = min(kChunk, Tn - t0);
      float a = 0.f;
      for (int t = 0; t < nt; ++t) a = fmaf(dg[t + 1], sv[v * kHalo + t + tau], a);
      eacc += a;
    }
    __syncthreads();
    if (w_on) {
      const int nt = min(kChunk, Tn - t0);
      // four tokens per step: their eight 16/32-byte loads are issued before the first use (one token at a time left every
      // load's latency exposed: this phase is the whole cost of the kernel)
      for (int tb = wg; tb < nt; tb += 4 * groups) {
        float f[4][8], g[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = tb + u * groups;
          if (t < nt) {
            load8<T>(xs + (size_t)(t0 + t) * D + 8 * wc, f[u]);
            load8<T>(ds + (size_t)(t0 + t) * D + 8 * wc, g[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = tb + u * groups;
          if (t < nt) {
            float o[8];
            const float gt = gate_s[t];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = g[u][e] * gt;
#pragma unroll
            for (int v = 0; v < kMaxV; ++v)
              if (v < V) {
                const float dv = dvw[v * kHalo + t];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  o[e] = fmaf(dv, wv_s[v * D + 8 * wc + e], o[e]);
                  wacc[v][e] = fmaf(dv, f[u][e], wacc[v][e]);
                }
              }
            store8<T>(dxs + (size_t)(t0 + t) * D + 8 * wc, o);
          }
        }
      }
    }
    __syncthreads();
  }
  if (tid < 3 * kMaxV && (tid % kMaxV) < V) p.dweff_part[(size_t)blockIdx.x * 3 * V + (tid / kMaxV) * V + tid % kMaxV] = eacc;
  if (w_on) {
    float* dw = p.dwv_part + (size_t)(blockIdx.x * groups + wg) * V * D;
#pragma unroll
    for (int v = 0; v < kMaxV; ++v)
      if (v < V) {
#pragma unroll
        for (int e = 0; e < 8; ++e) dw[(size_t)v * D + 8 * wc + e] = wacc[v][e];
      }
  }
}

}  // namespace tokgate1d
}  // namespace mop
