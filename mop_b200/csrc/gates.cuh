// Whisper-MoP "2D mixture of products" gate, fused.  Reference mop/models/whisper_mop.py:47-124:
//   V = conv1x1(mel2d) (V views), K = conv_kxk(V, zero padding k/2) (K pattern maps), [g_pos, g_neg] = conv1x1([V; K]),
//   gate_t = 1 + a_pos mean_f(g_pos) - a_neg mean_f(g_neg).
// Every stage is linear and bias-free, so the whole module is ONE k x k filter He of the mel map followed by the mean over
// the mel bins (the caller folds the weights: He = a_pos H_pos - a_neg H_neg, tiny, differentiable in PyTorch).  With
//   R_w[b, t'] = sum over the bins f' that column w of the filter sees at row t' (a row sum minus edge bins: zero padding)
// the gate is a 1-D convolution over time:
//   gate[b, t] = 1 + (1 / F) sum_{u, w} He[u, w] R_w[b, t + u - k/2].
// The [B, V + K, T, F] intermediates of the reference (12 layers x 8 maps of 1500 x 80 per sample) are never formed.
// HBM-bound: the forward reads mel once (B T F floats) and writes R (B T k) + gate (B T); the backward reads R and dgate.
#pragma once
#include "common.cuh"

namespace mop {
namespace gates {

constexpr int kMaxK = 9;

// grid: B * ceil(T / 8) CTAs of 256 threads: one warp per time step
static __global__ void __launch_bounds__(256) mop2d_rowsums_kernel(const float* mel, float* R, int B, int T, int F, int ks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= (long long)B * T) return;
  const float* m = mel + row * F;
  const int p = ks / 2;
  float tot = 0.f, head[kMaxK / 2 + 1], tail[kMaxK / 2 + 1];   // sums of the first / last j bins, j <= p
#pragma unroll
  for (int j = 0; j <= kMaxK / 2; ++j) { head[j] = 0.f; tail[j] = 0.f; }
  for (int f = lane; f < F; f += 32) {
    const float v = m[f];
    tot += v;
#pragma unroll
    for (int j = 1; j <= kMaxK / 2; ++j) {
      if (j <= p && f < j) head[j] += v;
      if (j <= p && f >= F - j) tail[j] += v;
    }
  }
  tot = warp_sum(tot);
#pragma unroll
  for (int j = 1; j <= kMaxK / 2; ++j) { head[j] = warp_sum(head[j]); tail[j] = warp_sum(tail[j]); }
  // column w of the filter reads bins f + w - p, f in [0, F): w < p misses the last p - w bins, w > p the first w - p
  if (lane < ks) {
    const int w = lane;
    float r = tot;
    for (int j = 1; j <= p; ++j) {
      if (p - w == j) r -= tail[j];
      if (w - p == j) r -= head[j];
    }
    R[row * ks + w] = r;
  }
}

// gate[b,t] = 1 + (1/F) sum_{u,w} He[u,w] R[b, t+u-p, w]
static __global__ void __launch_bounds__(256) mop2d_gate_kernel(const float* R, const float* He, float* gate, int B, int T, int F, int ks) {
  __shared__ float h[kMaxK * kMaxK];
  if (threadIdx.x < ks * ks) h[threadIdx.x] = He[threadIdx.x];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * T) return;
  const int t = (int)(idx % T);
  const long long b0 = idx - t;
  const int p = ks / 2;
  float acc = 0.f;
  for (int u = 0; u < ks; ++u) {
    const int tt = t + u - p;
    if (tt < 0 || tt >= T) continue;
    const float* r = R + (b0 + tt) * ks;
    for (int w = 0; w < ks; ++w) acc = fmaf(h[u * ks + w], r[w], acc);
  }
  gate[idx] = 1.f + acc / (float)F;
}

// dHe_part[cta][u,w] = (1/F) sum over this CTA's (b,t) of dgate[b,t] R[b, t+u-p, w]   (summed over CTAs by the caller)
static __global__ void __launch_bounds__(256) mop2d_bwd_kernel(const float* R, const float* dgate, float* dHe_part, int B, int T, int F, int ks) {
  __shared__ float red[8][kMaxK * kMaxK];
  const int p = ks / 2, nk = ks * ks, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[kMaxK * kMaxK];
#pragma unroll
  for (int i = 0; i < kMaxK * kMaxK; ++i) acc[i] = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)B * T; idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx % T);
    const long long b0 = idx - t;
    const float dg = dgate[idx];
#pragma unroll
    for (int u = 0; u < kMaxK; ++u) {
      if (u >= ks) break;
      const int tt = t + u - p;
      if (tt < 0 || tt >= T) continue;
      const float* r = R + (b0 + tt) * ks;
#pragma unroll
      for (int w = 0; w < kMaxK; ++w)
        if (w < ks) acc[u * kMaxK + w] = fmaf(dg, r[w], acc[u * kMaxK + w]);
    }
  }
#pragma unroll
  for (int u = 0; u < kMaxK; ++u)
#pragma unroll
    for (int w = 0; w < kMaxK; ++w) {
      if (u < ks && w < ks) {
        const float s = warp_sum(acc[u * kMaxK + w]);
        if (lane == 0) red[warp][u * ks + w] = s;
      }
    }
  __syncthreads();
  if (threadIdx.x < nk) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    dHe_part[(size_t)blockIdx.x * nk + threadIdx.x] = s / (float)F;
  }
}

}  // namespace gates
}  // namespace mop
