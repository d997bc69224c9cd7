// Block-cooperative fp32 building blocks for the "fp32 mode" kernels.
//
// Every routine is called by ALL threads of a 256-thread CTA with identical
// arguments and ends with __syncthreads().  Matrices live behind generic
// pointers (global scratch that stays L1/L2 resident, or shared memory) and
// are addressed with two strides so that transposes cost nothing.
//
// These are CUDA-core FFMA routines on purpose: tcgen05 has no fp32-operand
// MMA, and the fp32 mode exists to meet the 1e-5 parity bar.
#pragma once
#include "common.cuh"

namespace mop {
namespace simt {

constexpr int kThreads = 256;
constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

struct GemmSmem {
  float a[BK][BM + PAD];
  float b[BK][BN + PAD];
};

// C[M x Nc] (ldc) = (acc ? C : 0) + alpha * nscale[n] * sum_k A(m,k) * kscale[k] * B(k,n)
//   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn];  kscale / nscale may be null.
static __device__ __noinline__ void gemm(float* C, int ldc, const float* A, int sam, int sak, const float* B, int sbk,
                                  int sbn, int M, int Nc, int K, const float* kscale, const float* nscale,
                                  float alpha, bool acc, GemmSmem& sm) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  for (int m0 = 0; m0 < M; m0 += BM) {
    for (int n0 = 0; n0 < Nc; n0 += BN) {
      float c[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
      for (int k0 = 0; k0 < K; k0 += BK) {
        // ---- stage A tile [BK][BM]
#pragma unroll
        for (int p = 0; p < (BM * BK) / kThreads; ++p) {
          int idx = tid + p * kThreads;
          int m, k;
          if (sak == 1) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
          int gm = m0 + m, gk = k0 + k;
          float v = 0.f;
          if (gm < M && gk < K) {
            v = A[(size_t)gm * sam + (size_t)gk * sak];
            if (kscale) v *= kscale[gk];
          }
          sm.a[k][m] = v;
        }
        // ---- stage B tile [BK][BN]
#pragma unroll
        for (int p = 0; p < (BN * BK) / kThreads; ++p) {
          int idx = tid + p * kThreads;
          int n, k;
          if (sbk == 1) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
          int gn = n0 + n, gk = k0 + k;
          float v = 0.f;
          if (gn < Nc && gk < K) v = B[(size_t)gk * sbk + (size_t)gn * sbn];
          sm.b[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          float4 av = *reinterpret_cast<const float4*>(&sm.a[k][ty * 4]);
          float4 bv = *reinterpret_cast<const float4*>(&sm.b[k][tx * 4]);
          float a4[4] = {av.x, av.y, av.z, av.w};
          float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) c[i][j] = fmaf(a4[i], b4[j], c[i][j]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int gn = n0 + tx * 4 + j;
          if (gn >= Nc) continue;
          float v = alpha * c[i][j];
          if (nscale) v *= nscale[gn];
          float* dst = C + (size_t)gm * ldc + gn;
          *dst = acc ? (*dst + v) : v;
        }
      }
    }
  }
  __syncthreads();
}

// dst[i,:] = softmax(src[i,:]) for an N x N row-major map (dst may alias src).
__device__ __forceinline__ void softmax_rows(float* dst, const float* src, int rows, int cols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kThreads / 32;
  for (int i = warp; i < rows; i += nw) {
    const float* s = src + (size_t)i * cols;
    float* d = dst + (size_t)i * cols;
    float mx = -INFINITY;
    for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, s[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < cols; j += 32) {
      float e = expf(s[j] - mx);
      d[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    float inv = 1.0f / sum;
    for (int j = lane; j < cols; j += 32) d[j] *= inv;
  }
  __syncthreads();
}

// in place: X[i,:] = P[i,:] * (X[i,:] - sum_j X[i,j] P[i,j])   (softmax backward)
__device__ __forceinline__ void softmax_bwd_rows(float* X, const float* P, int rows, int cols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kThreads / 32;
  for (int i = warp; i < rows; i += nw) {
    float* x = X + (size_t)i * cols;
    const float* p = P + (size_t)i * cols;
    float dot = 0.f;
    for (int j = lane; j < cols; j += 32) dot = fmaf(x[j], p[j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < cols; j += 32) x[j] = p[j] * (x[j] - dot);
  }
  __syncthreads();
}

// row means and column means of f(map): rmean[i] = mean_j f(M[i,j]), cmean[j] = mean_i f(M[i,j])
// LOG: f(x) = log(x + eps), else identity.
template <bool LOG>
__device__ __forceinline__ void row_col_means(const float* M, int N, float eps, float* rmean, float* cmean) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kThreads / 32;
  const float invN = 1.0f / (float)N;
  for (int i = warp; i < N; i += nw) {
    const float* r = M + (size_t)i * N;
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s += LOG ? logf(r[j] + eps) : r[j];
    s = warp_sum(s);
    if (lane == 0) rmean[i] = s * invN;
  }
  for (int j = threadIdx.x; j < N; j += kThreads) {
    float s = 0.f;
    for (int i = 0; i < N; ++i) {
      float v = M[(size_t)i * N + j];
      s += LOG ? logf(v + eps) : v;
    }
    cmean[j] = s * invN;
  }
  __syncthreads();
}

// block-wide sum; every thread gets the result.  `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < kThreads / 32) ? red[lane] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

}  // namespace simt
}  // namespace mop
