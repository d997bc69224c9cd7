// Plain / causal / biased / cross attention, fp32 mode (flash-style tiles, CUDA-core math).
//
// Replaces softmax(q k^T * scale [masks] [+ bias]) v of
//   components.py:61-64, attention_variants.py:42-46, whisper_mop.py:163-175,212-219.
// Forward keeps running row max / sum (online softmax) and saves the row
// log-sum-exp; backward recomputes the probabilities tile by tile:
//   P = exp(S - lse), dV = P^T dO, dP = dO V^T, dS = P (.) (dP - rowsum(dO (.) O)),
//   dQ = scale dS K, dK = scale dS^T Q.
// Two backward kernels (one owning K/V tiles, one owning Q tiles) keep every
// output written by exactly one CTA: no atomics, deterministic.
#pragma once
#include "simt_blas.cuh"

namespace mop {
namespace sdpa {

constexpr int TQ = 64, TK = 64;
constexpr int kMaxDk = 128;

struct Tiles {  // dynamic shared memory carve-up (floats)
  float *q, *k, *v, *s, *o, *aux, *m, *l, *dlt;
  simt::GemmSmem* gs;
};

__host__ __device__ inline size_t smem_bytes(int dk) {
  return sizeof(simt::GemmSmem) + sizeof(float) * ((size_t)TQ * dk * 5 + (size_t)TQ * TK + 3 * TQ);
}

__device__ inline Tiles carve(unsigned char* raw, int dk) {
  Tiles t;
  t.gs = reinterpret_cast<simt::GemmSmem*>(raw);
  float* f = reinterpret_cast<float*>(raw + sizeof(simt::GemmSmem));
  t.q = f; f += TQ * dk;
  t.k = f; f += TK * dk;
  t.v = f; f += TK * dk;
  t.o = f; f += TQ * dk;
  t.aux = f; f += TQ * dk;
  t.s = f; f += TQ * TK;
  t.m = f; f += TQ;
  t.l = f; f += TQ;
  t.dlt = f;
  return t;
}

template <typename T>
__device__ inline void load_tile(float* dst, const T* base, int64_t sn, int row0, int rows_total, int dk) {
  for (int idx = threadIdx.x; idx < TQ * dk; idx += simt::kThreads) {
    int r = idx / dk, d = idx % dk;
    int gr = row0 + r;
    dst[idx] = gr < rows_total ? to_f32<T>(base[(int64_t)gr * sn + d]) : 0.f;
  }
}

// masked / biased score for element (gi, gj); returns -inf when dead
__device__ __forceinline__ float apply_masks(const MopSdpaParams& p, int b, int h, int gi, int gj, float s) {
  if (p.zero_mask) {
    float mk = p.zero_mask[(int64_t)b * p.zm_sb + (int64_t)h * p.zm_sh + (int64_t)gi * p.zm_sq + (int64_t)gj * p.zm_sk];
    if (mk == 0.f) s = -INFINITY;
  }
  if (p.causal && gj > gi) s = -INFINITY;
  if (p.bias) s += p.bias[(int64_t)b * p.bias_sb + (int64_t)h * p.bias_sh + (int64_t)gi * p.bias_sq + (int64_t)gj * p.bias_sk];
  return s;
}

template <typename T>
static __global__ void __launch_bounds__(simt::kThreads) fwd_kernel(MopSdpaParams p) {
  extern __shared__ __align__(16) unsigned char raw[];
  Tiles t = carve(raw, p.dk);
  const int dk = p.dk;
  const int nqb = (p.Nq + TQ - 1) / TQ;
  const int qb = blockIdx.x % nqb, bh = blockIdx.x / nqb, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * TQ;
  const int rows = min(TQ, p.Nq - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = simt::kThreads / 32;
  const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
  const T* qp = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_sb + (int64_t)h * p.q_sh;
  const T* kp = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.k_sb + (int64_t)h * p.k_sh;
  const T* vp = reinterpret_cast<const T*>(p.v) + (int64_t)b * p.v_sb + (int64_t)h * p.v_sh;
  load_tile<T>(t.q, qp, p.q_sn, q0, p.Nq, dk);
  for (int idx = threadIdx.x; idx < TQ * dk; idx += simt::kThreads) t.o[idx] = 0.f;
  for (int r = threadIdx.x; r < TQ; r += simt::kThreads) { t.m[r] = -INFINITY; t.l[r] = 0.f; }
  __syncthreads();
  int k_end = p.Nk;
  if (p.causal) k_end = min(p.Nk, q0 + rows);  // columns j <= i only
  for (int k0 = 0; k0 < k_end; k0 += TK) {
    const int cols = min(TK, p.Nk - k0);
    load_tile<T>(t.k, kp, p.k_sn, k0, p.Nk, dk);
    load_tile<T>(t.v, vp, p.v_sn, k0, p.Nk, dk);
    __syncthreads();
    simt::gemm(t.s, TK, t.q, dk, 1, t.k, 1, dk, rows, cols, dk, nullptr, nullptr, p.scale, false, *t.gs);
    for (int r = warp; r < rows; r += nw) {
      float* srow = t.s + r * TK;
      float mx = -INFINITY;
      for (int j = lane; j < cols; j += 32) {
        float s = apply_masks(p, b, h, q0 + r, k0 + j, srow[j]);
        srow[j] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      float m_old = t.m[r];
      float m_new = fmaxf(m_old, mx);
      float corr = (m_new == -INFINITY) ? 1.f : expf(m_old - m_new);
      float sum = 0.f;
      const uint32_t rkey = dropout_row_key(drop, (uint32_t)bh, (uint32_t)(q0 + r));
      for (int j = lane; j < cols; j += 32) {
        float e = (m_new == -INFINITY) ? 0.f : expf(srow[j] - m_new);
        sum += e;   // the softmax denominator is that of the un-dropped row; 1/(1-p) rides on the kept entries
        srow[j] = drop.on ? e * dropout_factor(drop, rkey, (uint32_t)(k0 + j)) : e;
      }
      sum = warp_sum(sum);
      for (int d = lane; d < dk; d += 32) t.o[r * dk + d] *= corr;
      if (lane == 0) { t.m[r] = m_new; t.l[r] = t.l[r] * corr + sum; }
    }
    __syncthreads();
    simt::gemm(t.o, dk, t.s, TK, 1, t.v, dk, 1, rows, dk, cols, nullptr, nullptr, 1.f, true, *t.gs);
  }
  T* y = reinterpret_cast<T*>(p.y);
  for (int idx = threadIdx.x; idx < rows * dk; idx += simt::kThreads) {
    int r = idx / dk, d = idx % dk;
    // a fully masked row gives 0/0 = NaN, like softmax over all -inf in the reference
    y[(((int64_t)b * p.Nq + q0 + r) * p.H + h) * dk + d] = from_f32<T>(t.o[idx] / t.l[r]);
  }
  if (p.lse)
    for (int r = threadIdx.x; r < rows; r += simt::kThreads)
      p.lse[((int64_t)b * p.H + h) * p.Nq + q0 + r] = t.m[r] + logf(t.l[r]);
}

// Recompute P (into t.s) and dS (into t.s, in place) for the tile (q0.., k0..).
// Needs t.q, t.k, t.v, t.aux = dO tile, t.l = lse rows, t.dlt = delta rows.
// After the call: t.s = dS = P (.) (M (.) dP - delta); `p_keep` (TQ*TK floats) holds M (.) P, the factor of dV (M = dropout factor).
__device__ inline void recompute_tile(const MopSdpaParams& p, const Tiles& t, int b, int h, int q0, int rows, int k0,
                                      int cols, float* p_keep) {
  const int dk = p.dk;
  simt::gemm(t.s, TK, t.q, dk, 1, t.k, 1, dk, rows, cols, dk, nullptr, nullptr, p.scale, false, *t.gs);
  for (int idx = threadIdx.x; idx < rows * TK; idx += simt::kThreads) {
    int r = idx / TK, j = idx % TK;
    float pr = 0.f;
    if (j < cols) {
      float s = apply_masks(p, b, h, q0 + r, k0 + j, t.s[idx]);
      pr = (s == -INFINITY) ? 0.f : expf(s - t.l[r]);
    }
    p_keep[idx] = pr;
  }
  __syncthreads();
  // dP = dO V^T
  simt::gemm(t.s, TK, t.aux, dk, 1, t.v, 1, dk, rows, cols, dk, nullptr, nullptr, 1.f, false, *t.gs);
  const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
  for (int idx = threadIdx.x; idx < rows * TK; idx += simt::kThreads) {
    int r = idx / TK, j = idx % TK;
    const float mk = drop.on ? dropout_factor(drop, dropout_row_key(drop, (uint32_t)(b * p.H + h), (uint32_t)(q0 + r)), (uint32_t)(k0 + j)) : 1.f;
    t.s[idx] = (j < cols) ? p_keep[idx] * (mk * t.s[idx] - t.dlt[r]) : 0.f;
    p_keep[idx] *= mk;
  }
  __syncthreads();
}

template <typename T>
__device__ inline void load_row_stats(const MopSdpaParams& p, const Tiles& t, int b, int h, int q0, int rows) {
  // t.aux holds the dO tile already; delta_r = sum_d dO[r,d] * O[r,d]
  const int dk = p.dk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = simt::kThreads / 32;
  const T* y = reinterpret_cast<const T*>(p.y);
  for (int r = warp; r < TQ; r += nw) {
    float s = 0.f;
    if (r < rows)
      for (int d = lane; d < dk; d += 32)
        s = fmaf(t.aux[r * dk + d], to_f32<T>(y[(((int64_t)b * p.Nq + q0 + r) * p.H + h) * dk + d]), s);
    s = warp_sum(s);
    if (lane == 0) {
      t.dlt[r] = s;
      t.l[r] = r < rows ? p.lse[((int64_t)b * p.H + h) * p.Nq + q0 + r] : 0.f;
    }
  }
  __syncthreads();
}

// grid: (b,h,k-block).  Owns dK, dV of its keys.
template <typename T>
static __global__ void __launch_bounds__(simt::kThreads) bwd_dkdv_kernel(MopSdpaParams p, float* pbuf_base) {
  extern __shared__ __align__(16) unsigned char raw[];
  Tiles t = carve(raw, p.dk);
  const int dk = p.dk;
  const int nkb = (p.Nk + TK - 1) / TK;
  const int kbk = blockIdx.x % nkb, bh = blockIdx.x / nkb, b = bh / p.H, h = bh % p.H;
  const int k0 = kbk * TK, cols = min(TK, p.Nk - k0);
  float* pkeep = pbuf_base + (size_t)blockIdx.x * (TQ * TK + 2 * TK * kMaxDk);
  float* dkacc = pkeep + TQ * TK;
  float* dvacc = dkacc + TK * kMaxDk;
  const T* qp = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_sb + (int64_t)h * p.q_sh;
  const T* kp = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.k_sb + (int64_t)h * p.k_sh;
  const T* vp = reinterpret_cast<const T*>(p.v) + (int64_t)b * p.v_sb + (int64_t)h * p.v_sh;
  const T* dyp = reinterpret_cast<const T*>(p.dy) + ((int64_t)b * p.Nq * p.H + h) * dk;
  load_tile<T>(t.k, kp, p.k_sn, k0, p.Nk, dk);
  load_tile<T>(t.v, vp, p.v_sn, k0, p.Nk, dk);
  for (int idx = threadIdx.x; idx < TK * dk; idx += simt::kThreads) { dkacc[idx] = 0.f; dvacc[idx] = 0.f; }
  __syncthreads();
  int q_begin = 0;
  if (p.causal) q_begin = (k0 / TQ) * TQ;  // rows i >= j only
  for (int q0 = q_begin; q0 < p.Nq; q0 += TQ) {
    const int rows = min(TQ, p.Nq - q0);
    load_tile<T>(t.q, qp, p.q_sn, q0, p.Nq, dk);
    load_tile<T>(t.aux, dyp, (int64_t)p.H * dk, q0, p.Nq, dk);
    __syncthreads();
    load_row_stats<T>(p, t, b, h, q0, rows);
    recompute_tile(p, t, b, h, q0, rows, k0, cols, pkeep);
    // dV += P^T dO ; dK += scale dS^T Q
    simt::gemm(dvacc, dk, pkeep, 1, TK, t.aux, dk, 1, cols, dk, rows, nullptr, nullptr, 1.f, true, *t.gs);
    simt::gemm(dkacc, dk, t.s, 1, TK, t.q, dk, 1, cols, dk, rows, nullptr, nullptr, p.scale, true, *t.gs);
  }
  T* dK = reinterpret_cast<T*>(p.dk_);
  T* dV = reinterpret_cast<T*>(p.dv);
  for (int idx = threadIdx.x; idx < cols * dk; idx += simt::kThreads) {
    int r = idx / dk, d = idx % dk;
    int64_t o = (((int64_t)b * p.Nk + k0 + r) * p.H + h) * dk + d;
    dK[o] = from_f32<T>(dkacc[idx]);
    dV[o] = from_f32<T>(dvacc[idx]);
  }
}

// grid: (b,h,q-block).  Owns dQ of its queries.
template <typename T>
static __global__ void __launch_bounds__(simt::kThreads) bwd_dq_kernel(MopSdpaParams p, float* pbuf_base) {
  extern __shared__ __align__(16) unsigned char raw[];
  Tiles t = carve(raw, p.dk);
  const int dk = p.dk;
  const int nqb = (p.Nq + TQ - 1) / TQ;
  const int qb = blockIdx.x % nqb, bh = blockIdx.x / nqb, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * TQ, rows = min(TQ, p.Nq - q0);
  float* pkeep = pbuf_base + (size_t)blockIdx.x * (TQ * TK + 2 * TK * kMaxDk);
  const T* qp = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_sb + (int64_t)h * p.q_sh;
  const T* kp = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.k_sb + (int64_t)h * p.k_sh;
  const T* vp = reinterpret_cast<const T*>(p.v) + (int64_t)b * p.v_sb + (int64_t)h * p.v_sh;
  const T* dyp = reinterpret_cast<const T*>(p.dy) + ((int64_t)b * p.Nq * p.H + h) * dk;
  load_tile<T>(t.q, qp, p.q_sn, q0, p.Nq, dk);
  load_tile<T>(t.aux, dyp, (int64_t)p.H * dk, q0, p.Nq, dk);
  for (int idx = threadIdx.x; idx < TQ * dk; idx += simt::kThreads) t.o[idx] = 0.f;
  __syncthreads();
  load_row_stats<T>(p, t, b, h, q0, rows);
  int k_end = p.Nk;
  if (p.causal) k_end = min(p.Nk, q0 + rows);
  for (int k0 = 0; k0 < k_end; k0 += TK) {
    const int cols = min(TK, p.Nk - k0);
    load_tile<T>(t.k, kp, p.k_sn, k0, p.Nk, dk);
    load_tile<T>(t.v, vp, p.v_sn, k0, p.Nk, dk);
    __syncthreads();
    recompute_tile(p, t, b, h, q0, rows, k0, cols, pkeep);
    simt::gemm(t.o, dk, t.s, TK, 1, t.k, dk, 1, rows, dk, cols, nullptr, nullptr, p.scale, true, *t.gs);
  }
  T* dQ = reinterpret_cast<T*>(p.dq);
  for (int idx = threadIdx.x; idx < rows * dk; idx += simt::kThreads) {
    int r = idx / dk, d = idx % dk;
    dQ[(((int64_t)b * p.Nq + q0 + r) * p.H + h) * dk + d] = from_f32<T>(t.o[idx]);
  }
}

inline size_t bwd_workspace_floats(const MopSdpaParams* p) {
  size_t nkb = (p->Nk + TK - 1) / TK, nqb = (p->Nq + TQ - 1) / TQ;
  size_t ctas = (size_t)p->B * p->H * (nkb > nqb ? nkb : nqb);
  return ctas * (TQ * TK + 2 * TK * kMaxDk);
}

}  // namespace sdpa
}  // namespace mop
