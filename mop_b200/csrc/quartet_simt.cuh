// GPT "Quartet" causal attention, fp32 mode (flash-style tiles, CUDA-core math).
//
// Replaces quartet_attn_patch.py:88-121:  two score maps, each z-scored per row over the FULL
// unmasked row with the unbiased std (:95-98), mixed as (1-m) n1 + m gamma n1 n2 (:103-106), causal
// fill, optional additive mask, softmax, PV.  `use_quartet = 0` keeps the single z-scored map (:108-110).
//
// Centred form (oracle/quartet.py::quartet_core_manual): with
// kc_j = k_j - mean_j k_j the score minus its row mean is c_ij = s q_i . kc_j exactly, so
//   n_ij = c_ij / (sigma_i + eps),  sigma_i^2 = sum_j c_ij^2 / (T-1)        (all j, masked or not)
// Forward: a statistics sweep over all key tiles, then a causal sweep with online softmax.
// Backward, given dn (nonzero for j <= i only):
//   g_i   = sum_j dn_ij c_ij / ((sigma_i+eps)^2 (T-1) sigma_i)
//   dq_i  = s sum_{j<=i} W_ij kc_j - s^2 g_i (Kc^T Kc) q_i,      W = dn / (sigma+eps)
//   dkc_j = s sum_{i>=j} W_ij q_i  - s^2 (sum_i g_i q_i q_i^T) kc_j,   dk = dkc - mean_j dkc
// i.e. causal tiles plus two dk x dk Gram corrections; the dense non-causal part of dS is never formed.
// Kernels (all deterministic, every output written by exactly one CTA):
//   prep -> fwd                                   (forward)
//   prep -> bwd_dq -> gmat -> bwd_dkdv -> finish  (backward)
#pragma once
#include "simt_blas.cuh"

namespace mop {
namespace quartet {

constexpr int TQ = 64, TK = 64;
constexpr int kMaxDk = 64;

struct Ws {  // workspace layout in floats
  size_t kc, qf, gram, gvec, mmat, dkc, spart, acc, total;
  int nm, nqb;
};

__host__ __device__ inline Ws layout(const MopQuartetParams* p, int backward) {
  Ws w;
  const size_t BH = (size_t)p->B * p->H, T = p->T, dk = p->dk;
  w.nm = p->use_quartet ? 2 : 1;
  w.nqb = (p->T + TQ - 1) / TQ;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };
  w.kc = take(w.nm * BH * T * dk);
  w.qf = take(w.nm * BH * T * dk);
  w.gram = take(w.nm * BH * dk * dk);
  w.gvec = w.mmat = w.dkc = w.spart = w.acc = 0;
  if (backward) {
    w.gvec = take(w.nm * BH * T);
    w.mmat = take(w.nm * BH * dk * dk);
    w.dkc = take(w.nm * BH * T * dk);
    w.spart = take(BH * w.nqb * 2);
    w.acc = take(BH * w.nqb * 3 * TQ * dk);
  }
  w.total = o;
  return w;
}

// element (b, t, h, d) of a [B,T,H,dk] contiguous tensor
__device__ __forceinline__ size_t at(const MopQuartetParams& p, int b, int t, int h) { return (((size_t)b * p.T + t) * p.H + h) * p.dk; }

// grid: B*H*nm.  kc = k - mean(k), qf = q (fp32 copies), gram = kc^T kc.
template <typename T>
static __global__ void __launch_bounds__(simt::kThreads) prep_kernel(MopQuartetParams p, Ws w, float* ws) {
  __shared__ simt::GemmSmem gs;
  __shared__ float kbar[kMaxDk];
  const int nm = w.nm, bh = blockIdx.x / nm, map = blockIdx.x % nm, b = bh / p.H, h = bh % p.H, dk = p.dk, Tn = p.T;
  const T* k = reinterpret_cast<const T*>(map ? p.k2 : p.k);
  const T* q = reinterpret_cast<const T*>(map ? p.q2 : p.q);
  float* kc = ws + w.kc + ((size_t)map * p.B * p.H + bh) * Tn * dk;
  float* qf = ws + w.qf + ((size_t)map * p.B * p.H + bh) * Tn * dk;
  float* gram = ws + w.gram + ((size_t)map * p.B * p.H + bh) * dk * dk;
  for (int d = threadIdx.x; d < dk; d += simt::kThreads) {
    float s = 0.f;
    for (int t = 0; t < Tn; ++t) s += to_f32<T>(k[at(p, b, t, h) + d]);
    kbar[d] = s / (float)Tn;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < Tn * dk; idx += simt::kThreads) {
    int t = idx / dk, d = idx % dk;
    kc[idx] = to_f32<T>(k[at(p, b, t, h) + d]) - kbar[d];
    qf[idx] = to_f32<T>(q[at(p, b, t, h) + d]);
  }
  __syncthreads();
  simt::gemm(gram, dk, kc, 1, dk, kc, dk, 1, dk, dk, Tn, nullptr, nullptr, 1.f, false, gs);
}

struct Tiles {
  float *q, *q2, *dy, *k1, *k2, *v, *s1, *s2, *pp, *sig1, *sig2, *lse, *dlt, *g1, *g2, *m, *l;
  simt::GemmSmem* gs;
};
__host__ __device__ inline size_t smem_bytes(int dk) {
  return sizeof(simt::GemmSmem) + sizeof(float) * ((size_t)TQ * dk * 6 + (size_t)TQ * TK * 3 + 8 * TQ);
}
__device__ inline Tiles carve(unsigned char* raw, int dk) {
  Tiles t;
  t.gs = reinterpret_cast<simt::GemmSmem*>(raw);
  float* f = reinterpret_cast<float*>(raw + sizeof(simt::GemmSmem));
  t.q = f; f += TQ * dk; t.q2 = f; f += TQ * dk; t.dy = f; f += TQ * dk;
  t.k1 = f; f += TK * dk; t.k2 = f; f += TK * dk; t.v = f; f += TK * dk;
  t.s1 = f; f += TQ * TK; t.s2 = f; f += TQ * TK; t.pp = f; f += TQ * TK;
  t.sig1 = f; f += TQ; t.sig2 = f; f += TQ; t.lse = f; f += TQ; t.dlt = f; f += TQ;
  t.g1 = f; f += TQ; t.g2 = f; f += TQ; t.m = f; f += TQ; t.l = f;
  return t;
}

__device__ inline void load_f32_tile(float* dst, const float* src, int row0, int rows_total, int dk) {
  for (int idx = threadIdx.x; idx < TQ * dk; idx += simt::kThreads) {
    int r = idx / dk, gr = row0 + r;
    dst[idx] = gr < rows_total ? src[(size_t)gr * dk + idx % dk] : 0.f;
  }
}
template <typename T>
__device__ inline void load_bthd_tile(float* dst, const T* src, const MopQuartetParams& p, int b, int h, int row0) {
  const int dk = p.dk;
  for (int idx = threadIdx.x; idx < TQ * dk; idx += simt::kThreads) {
    int r = idx / dk, gr = row0 + r;
    dst[idx] = gr < p.T ? to_f32<T>(src[at(p, b, gr, h) + idx % dk]) : 0.f;
  }
}

struct Mix {
  float m, gam, eps;
  bool quart;
};
__device__ __forceinline__ Mix load_mix(const MopQuartetParams& p) {
  Mix x;
  x.quart = p.use_quartet != 0;
  x.m = x.quart ? 1.f / (1.f + expf(-p.mixture[0])) : 0.f;
  x.gam = x.quart ? p.quartet_scale[0] : 0.f;
  x.eps = p.eps;
  return x;
}
// score for element (gi, gj) from the centred scores; -inf above the diagonal; additive mask after the fill
__device__ __forceinline__ float mix_score(const MopQuartetParams& p, const Mix& x, int b, int h, int gi, int gj, float n1, float n2) {
  float sc = x.quart ? (1.f - x.m) * n1 + x.m * x.gam * n1 * n2 : n1;
  if (gj > gi) sc = -INFINITY;
  if (p.add_mask) sc += p.add_mask[(int64_t)b * p.am_sb + (int64_t)h * p.am_sh + (int64_t)gi * p.am_sq + (int64_t)gj * p.am_sk];
  return sc;
}

// grid: (b,h,q-block)
template <typename T>
static __global__ void __launch_bounds__(simt::kThreads) fwd_kernel(MopQuartetParams p, Ws w, float* ws) {
  extern __shared__ __align__(16) unsigned char raw[];
  Tiles t = carve(raw, p.dk);
  const int dk = p.dk, Tn = p.T, nqb = w.nqb;
  const int qb = blockIdx.x % nqb, bh = blockIdx.x / nqb, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * TQ, rows = min(TQ, Tn - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = simt::kThreads / 32;
  const Mix mx = load_mix(p);
  const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
  const size_t BH = (size_t)p.B * p.H;
  const float* kc1 = ws + w.kc + (size_t)bh * Tn * dk;
  const float* kc2 = ws + w.kc + (BH + bh) * Tn * dk;
  load_bthd_tile<T>(t.q, reinterpret_cast<const T*>(p.q), p, b, h, q0);
  if (mx.quart) load_bthd_tile<T>(t.q2, reinterpret_cast<const T*>(p.q2), p, b, h, q0);
  float* o = t.dy;  // output accumulator reuses the dy tile
  for (int idx = threadIdx.x; idx < TQ * dk; idx += simt::kThreads) o[idx] = 0.f;
  for (int r = threadIdx.x; r < TQ; r += simt::kThreads) { t.sig1[r] = 0.f; t.sig2[r] = 0.f; t.m[r] = -INFINITY; t.l[r] = 0.f; }
  __syncthreads();
  // ---- sweep 1: sigma over the full row (masked positions included, as the reference does)
  for (int k0 = 0; k0 < Tn; k0 += TK) {
    const int cols = min(TK, Tn - k0);
    load_f32_tile(t.k1, kc1, k0, Tn, dk);
    if (mx.quart) load_f32_tile(t.k2, kc2, k0, Tn, dk);
    __syncthreads();
    simt::gemm(t.s1, TK, t.q, dk, 1, t.k1, 1, dk, rows, cols, dk, nullptr, nullptr, p.scale, false, *t.gs);
    if (mx.quart) simt::gemm(t.s2, TK, t.q2, dk, 1, t.k2, 1, dk, rows, cols, dk, nullptr, nullptr, p.scale, false, *t.gs);
    for (int r = warp; r < rows; r += nw) {
      float a = 0.f, c = 0.f;
      for (int j = lane; j < cols; j += 32) {
        a = fmaf(t.s1[r * TK + j], t.s1[r * TK + j], a);
        if (mx.quart) c = fmaf(t.s2[r * TK + j], t.s2[r * TK + j], c);
      }
      a = warp_sum(a); c = warp_sum(c);
      if (lane == 0) { t.sig1[r] += a; t.sig2[r] += c; }
    }
    __syncthreads();
  }
  for (int r = threadIdx.x; r < TQ; r += simt::kThreads) {
    t.sig1[r] = sqrtf(t.sig1[r] / (float)(Tn - 1));
    t.sig2[r] = sqrtf(t.sig2[r] / (float)(Tn - 1));
  }
  __syncthreads();
  // ---- sweep 2: causal tiles, online softmax, PV
  const int k_end = min(Tn, q0 + rows);
  for (int k0 = 0; k0 < k_end; k0 += TK) {
    const int cols = min(TK, Tn - k0);
    load_f32_tile(t.k1, kc1, k0, Tn, dk);
    if (mx.quart) load_f32_tile(t.k2, kc2, k0, Tn, dk);
    load_bthd_tile<T>(t.v, reinterpret_cast<const T*>(p.v), p, b, h, k0);
    __syncthreads();
    simt::gemm(t.s1, TK, t.q, dk, 1, t.k1, 1, dk, rows, cols, dk, nullptr, nullptr, p.scale, false, *t.gs);
    if (mx.quart) simt::gemm(t.s2, TK, t.q2, dk, 1, t.k2, 1, dk, rows, cols, dk, nullptr, nullptr, p.scale, false, *t.gs);
    for (int r = warp; r < rows; r += nw) {
      float* srow = t.s1 + r * TK;
      const float i1 = 1.f / (t.sig1[r] + mx.eps), i2 = mx.quart ? 1.f / (t.sig2[r] + mx.eps) : 0.f;
      float vmax = -INFINITY;
      for (int j = lane; j < cols; j += 32) {
        float sc = mix_score(p, mx, b, h, q0 + r, k0 + j, srow[j] * i1, mx.quart ? t.s2[r * TK + j] * i2 : 0.f);
        srow[j] = sc;
        vmax = fmaxf(vmax, sc);
      }
      vmax = warp_max(vmax);
      const float m_old = t.m[r], m_new = fmaxf(m_old, vmax);
      const float corr = (m_new == -INFINITY) ? 1.f : expf(m_old - m_new);
      float sum = 0.f;
      const uint32_t rkey = dropout_row_key(drop, (uint32_t)bh, (uint32_t)(q0 + r));
      for (int j = lane; j < cols; j += 32) {
        float e = (m_new == -INFINITY) ? 0.f : expf(srow[j] - m_new);
        sum += e;   // denominator of the un-dropped row; the kept entries carry 1/(1-p)
        srow[j] = drop.on ? e * dropout_factor(drop, rkey, (uint32_t)(k0 + j)) : e;
      }
      sum = warp_sum(sum);
      for (int d = lane; d < dk; d += 32) o[r * dk + d] *= corr;
      if (lane == 0) { t.m[r] = m_new; t.l[r] = t.l[r] * corr + sum; }
    }
    __syncthreads();
    simt::gemm(o, dk, t.s1, TK, 1, t.v, dk, 1, rows, dk, cols, nullptr, nullptr, 1.f, true, *t.gs);
  }
  T* y = reinterpret_cast<T*>(p.y);
  for (int idx = threadIdx.x; idx < rows * dk; idx += simt::kThreads) {
    int r = idx / dk, d = idx % dk;
    y[at(p, b, q0 + r, h) + d] = from_f32<T>(o[idx] / t.l[r]);
    if (p.y_f32) p.y_f32[at(p, b, q0 + r, h) + d] = o[idx] / t.l[r];
  }
  if (p.stats)
    for (int r = threadIdx.x; r < rows; r += simt::kThreads) {
      float* st = p.stats + (((size_t)b * p.H + h) * Tn + q0 + r) * 3;
      st[0] = t.sig1[r]; st[1] = t.sig2[r]; st[2] = t.m[r] + logf(t.l[r]);
    }
}

// Recompute one causal tile for the backward: on return
//   t.pp = M (.) P (probabilities times the dropout factor M: the operand of dV), t.s1 = W1 = dn1/(sigma1+eps), t.s2 = W2, row accumulators g1/g2 += sum_j dn c,
//   and the two scalar partials (d mixture-logit / d quartet_scale) are added to sc[0], sc[1] (per thread).
__device__ inline void recompute_tile(const MopQuartetParams& p, const Tiles& t, const Mix& mx, int b, int h, int q0, int rows,
                                      int k0, int cols, bool accumulate_rows, float* sc) {
  const int dk = p.dk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = simt::kThreads / 32;
  simt::gemm(t.s1, TK, t.q, dk, 1, t.k1, 1, dk, rows, cols, dk, nullptr, nullptr, p.scale, false, *t.gs);
  if (mx.quart) simt::gemm(t.s2, TK, t.q2, dk, 1, t.k2, 1, dk, rows, cols, dk, nullptr, nullptr, p.scale, false, *t.gs);
  simt::gemm(t.pp, TK, t.dy, dk, 1, t.v, 1, dk, rows, cols, dk, nullptr, nullptr, 1.f, false, *t.gs);  // dP = dO V^T
  const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
  for (int r = warp; r < rows; r += nw) {
    const float i1 = 1.f / (t.sig1[r] + mx.eps), i2 = mx.quart ? 1.f / (t.sig2[r] + mx.eps) : 0.f;
    float a1 = 0.f, a2 = 0.f;
    const uint32_t rkey = dropout_row_key(drop, (uint32_t)(b * p.H + h), (uint32_t)(q0 + r));
    for (int j = lane; j < cols; j += 32) {
      const float c1 = t.s1[r * TK + j], c2 = mx.quart ? t.s2[r * TK + j] : 0.f;
      const float n1 = c1 * i1, n2 = c2 * i2;
      const float s = mix_score(p, mx, b, h, q0 + r, k0 + j, n1, n2);
      const float pr = (s == -INFINITY) ? 0.f : expf(s - t.lse[r]);
      const float mk = drop.on ? dropout_factor(drop, rkey, (uint32_t)(k0 + j)) : 1.f;
      const float D = pr * (mk * t.pp[r * TK + j] - t.dlt[r]);
      float dn1 = D, dn2 = 0.f;
      if (mx.quart) {
        dn1 = D * ((1.f - mx.m) + mx.m * mx.gam * n2);
        dn2 = D * (mx.m * mx.gam * n1);
        sc[0] += D * (-n1 + mx.gam * n1 * n2);
        sc[1] += D * (mx.m * n1 * n2);
      }
      a1 = fmaf(dn1, c1, a1);
      a2 = fmaf(dn2, c2, a2);
      t.pp[r * TK + j] = pr * mk;
      t.s1[r * TK + j] = dn1 * i1;
      if (mx.quart) t.s2[r * TK + j] = dn2 * i2;
    }
    if (accumulate_rows) {
      a1 = warp_sum(a1); a2 = warp_sum(a2);
      if (lane == 0) { t.g1[r] += a1; t.g2[r] += a2; }
    }
  }
  __syncthreads();
}

template <typename T>
__device__ inline void load_q_side(const MopQuartetParams& p, const Tiles& t, const Mix& mx, int b, int h, int q0, int rows) {
  const int dk = p.dk, Tn = p.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = simt::kThreads / 32;
  load_bthd_tile<T>(t.q, reinterpret_cast<const T*>(p.q), p, b, h, q0);
  if (mx.quart) load_bthd_tile<T>(t.q2, reinterpret_cast<const T*>(p.q2), p, b, h, q0);
  load_bthd_tile<T>(t.dy, reinterpret_cast<const T*>(p.dy), p, b, h, q0);
  __syncthreads();
  const T* y = reinterpret_cast<const T*>(p.y);
  for (int r = warp; r < TQ; r += nw) {
    float s = 0.f;
    if (r < rows)
      for (int d = lane; d < dk; d += 32) s = fmaf(t.dy[r * dk + d], p.y_f32 ? p.y_f32[at(p, b, q0 + r, h) + d] : to_f32<T>(y[at(p, b, q0 + r, h) + d]), s);
    s = warp_sum(s);
    if (lane == 0) {
      const float* st = p.stats + (((size_t)b * p.H + h) * Tn + min(q0 + r, Tn - 1)) * 3;
      t.dlt[r] = s; t.sig1[r] = st[0]; t.sig2[r] = st[1]; t.lse[r] = st[2];
    }
  }
  __syncthreads();
}

// grid: (b,h,q-block).  Owns dq, dq2, the row coefficients g and the scalar partials of its rows.
template <typename T>
static __global__ void __launch_bounds__(simt::kThreads) bwd_dq_kernel(MopQuartetParams p, Ws w, float* ws) {
  extern __shared__ __align__(16) unsigned char raw[];
  __shared__ float red[32];
  Tiles t = carve(raw, p.dk);
  const int dk = p.dk, Tn = p.T, nqb = w.nqb;
  const int qb = blockIdx.x % nqb, bh = blockIdx.x / nqb, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * TQ, rows = min(TQ, Tn - q0);
  const Mix mx = load_mix(p);
  const size_t BH = (size_t)p.B * p.H;
  const float* kc1 = ws + w.kc + (size_t)bh * Tn * dk;
  const float* kc2 = ws + w.kc + (BH + bh) * Tn * dk;
  float* acc1 = ws + w.acc + (size_t)blockIdx.x * 3 * TQ * dk;
  float* acc2 = acc1 + TQ * dk;
  float* tmp = acc2 + TQ * dk;
  load_q_side<T>(p, t, mx, b, h, q0, rows);
  for (int idx = threadIdx.x; idx < TQ * dk; idx += simt::kThreads) { acc1[idx] = 0.f; acc2[idx] = 0.f; }
  for (int r = threadIdx.x; r < TQ; r += simt::kThreads) { t.g1[r] = 0.f; t.g2[r] = 0.f; }
  __syncthreads();
  float sc[2] = {0.f, 0.f};
  const int k_end = min(Tn, q0 + rows);
  for (int k0 = 0; k0 < k_end; k0 += TK) {
    const int cols = min(TK, Tn - k0);
    load_f32_tile(t.k1, kc1, k0, Tn, dk);
    if (mx.quart) load_f32_tile(t.k2, kc2, k0, Tn, dk);
    load_bthd_tile<T>(t.v, reinterpret_cast<const T*>(p.v), p, b, h, k0);
    __syncthreads();
    recompute_tile(p, t, mx, b, h, q0, rows, k0, cols, true, sc);
    simt::gemm(acc1, dk, t.s1, TK, 1, t.k1, dk, 1, rows, dk, cols, nullptr, nullptr, p.scale, true, *t.gs);
    if (mx.quart) simt::gemm(acc2, dk, t.s2, TK, 1, t.k2, dk, 1, rows, dk, cols, nullptr, nullptr, p.scale, true, *t.gs);
  }
  // g_i, Gram correction, outputs
  for (int r = threadIdx.x; r < TQ; r += simt::kThreads) {
    const float s1 = t.sig1[r], s2 = t.sig2[r];
    t.g1[r] = (r < rows) ? t.g1[r] / ((s1 + mx.eps) * (s1 + mx.eps) * (float)(Tn - 1) * s1) : 0.f;
    t.g2[r] = (r < rows && mx.quart) ? t.g2[r] / ((s2 + mx.eps) * (s2 + mx.eps) * (float)(Tn - 1) * s2) : 0.f;
  }
  __syncthreads();
  for (int map = 0; map < w.nm; ++map) {
    const float* gram = ws + w.gram + ((size_t)map * BH + bh) * dk * dk;
    const float* qt = map ? t.q2 : t.q;
    const float* gv = map ? t.g2 : t.g1;
    float* acc = map ? acc2 : acc1;
    simt::gemm(tmp, dk, qt, dk, 1, gram, dk, 1, rows, dk, dk, nullptr, nullptr, 1.f, false, *t.gs);
    T* out = reinterpret_cast<T*>(map ? p.dq2 : p.dq);
    float* gout = ws + w.gvec + ((size_t)map * BH + bh) * Tn;
    for (int idx = threadIdx.x; idx < rows * dk; idx += simt::kThreads) {
      int r = idx / dk, d = idx % dk;
      out[at(p, b, q0 + r, h) + d] = from_f32<T>(acc[idx] - p.scale * p.scale * gv[r] * tmp[idx]);
    }
    for (int r = threadIdx.x; r < rows; r += simt::kThreads) gout[q0 + r] = gv[r];
    __syncthreads();
  }
  const float s0 = simt::block_sum(sc[0], red), s1 = simt::block_sum(sc[1], red);
  if (threadIdx.x == 0) {
    float* sp = ws + w.spart + (size_t)blockIdx.x * 2;
    sp[0] = mx.m * (1.f - mx.m) * s0;
    sp[1] = s1;
  }
}

// grid: B*H*nm.  mmat = sum_i g_i q_i q_i^T
static __global__ void __launch_bounds__(simt::kThreads) gmat_kernel(MopQuartetParams p, Ws w, float* ws) {
  __shared__ simt::GemmSmem gs;
  const int nm = w.nm, bh = blockIdx.x / nm, map = blockIdx.x % nm, dk = p.dk, Tn = p.T;
  const size_t BH = (size_t)p.B * p.H;
  const float* qf = ws + w.qf + ((size_t)map * BH + bh) * Tn * dk;
  const float* gv = ws + w.gvec + ((size_t)map * BH + bh) * Tn;
  float* mm = ws + w.mmat + ((size_t)map * BH + bh) * dk * dk;
  simt::gemm(mm, dk, qf, 1, dk, qf, dk, 1, dk, dk, Tn, gv, nullptr, 1.f, false, gs);
}

// grid: (b,h,k-block).  Owns dv and the centred-key gradients dkc of its keys.
template <typename T>
static __global__ void __launch_bounds__(simt::kThreads) bwd_dkdv_kernel(MopQuartetParams p, Ws w, float* ws) {
  extern __shared__ __align__(16) unsigned char raw[];
  Tiles t = carve(raw, p.dk);
  const int dk = p.dk, Tn = p.T, nkb = w.nqb;
  const int kb = blockIdx.x % nkb, bh = blockIdx.x / nkb, b = bh / p.H, h = bh % p.H;
  const int k0 = kb * TK, cols = min(TK, Tn - k0);
  const Mix mx = load_mix(p);
  const size_t BH = (size_t)p.B * p.H;
  const float* kc1 = ws + w.kc + (size_t)bh * Tn * dk;
  const float* kc2 = ws + w.kc + (BH + bh) * Tn * dk;
  float* a1 = ws + w.acc + (size_t)blockIdx.x * 3 * TQ * dk;
  float* a2 = a1 + TK * dk;
  float* av = a2 + TK * dk;
  load_f32_tile(t.k1, kc1, k0, Tn, dk);
  if (mx.quart) load_f32_tile(t.k2, kc2, k0, Tn, dk);
  load_bthd_tile<T>(t.v, reinterpret_cast<const T*>(p.v), p, b, h, k0);
  for (int idx = threadIdx.x; idx < TK * dk; idx += simt::kThreads) { a1[idx] = 0.f; a2[idx] = 0.f; av[idx] = 0.f; }
  __syncthreads();
  float sc[2] = {0.f, 0.f};
  for (int q0 = k0; q0 < Tn; q0 += TQ) {   // rows i >= j only (TQ == TK)
    const int rows = min(TQ, Tn - q0);
    load_q_side<T>(p, t, mx, b, h, q0, rows);
    recompute_tile(p, t, mx, b, h, q0, rows, k0, cols, false, sc);
    simt::gemm(av, dk, t.pp, 1, TK, t.dy, dk, 1, cols, dk, rows, nullptr, nullptr, 1.f, true, *t.gs);              // dV += P^T dO
    simt::gemm(a1, dk, t.s1, 1, TK, t.q, dk, 1, cols, dk, rows, nullptr, nullptr, p.scale, true, *t.gs);           // s W1^T q
    if (mx.quart) simt::gemm(a2, dk, t.s2, 1, TK, t.q2, dk, 1, cols, dk, rows, nullptr, nullptr, p.scale, true, *t.gs);
  }
  T* dv = reinterpret_cast<T*>(p.dv);
  for (int idx = threadIdx.x; idx < cols * dk; idx += simt::kThreads) {
    int r = idx / dk, d = idx % dk;
    dv[at(p, b, k0 + r, h) + d] = from_f32<T>(av[idx]);
  }
  // dkc -= s^2 kc M ; stored fp32, centred by finish_kernel
  for (int map = 0; map < w.nm; ++map) {
    const float* mm = ws + w.mmat + ((size_t)map * BH + bh) * dk * dk;
    float* acc = map ? a2 : a1;
    simt::gemm(acc, dk, map ? t.k2 : t.k1, dk, 1, mm, dk, 1, cols, dk, dk, nullptr, nullptr, -p.scale * p.scale, true, *t.gs);
    float* out = ws + w.dkc + ((size_t)map * BH + bh) * Tn * dk;
    for (int idx = threadIdx.x; idx < cols * dk; idx += simt::kThreads) out[(size_t)k0 * dk + idx] = acc[idx];
    __syncthreads();
  }
}

// grid: B*H*nm.  dk = dkc - mean_j dkc ; map 0 of each (b,h) also reduces the scalar partials.
template <typename T>
static __global__ void __launch_bounds__(simt::kThreads) finish_kernel(MopQuartetParams p, Ws w, float* ws) {
  __shared__ float mean[kMaxDk];
  const int nm = w.nm, bh = blockIdx.x / nm, map = blockIdx.x % nm, b = bh / p.H, h = bh % p.H, dk = p.dk, Tn = p.T;
  const size_t BH = (size_t)p.B * p.H;
  const float* dkc = ws + w.dkc + ((size_t)map * BH + bh) * Tn * dk;
  for (int d = threadIdx.x; d < dk; d += simt::kThreads) {
    float s = 0.f;
    for (int t = 0; t < Tn; ++t) s += dkc[(size_t)t * dk + d];
    mean[d] = s / (float)Tn;
  }
  __syncthreads();
  T* out = reinterpret_cast<T*>(map ? p.dk2 : p.dk_);
  for (int idx = threadIdx.x; idx < Tn * dk; idx += simt::kThreads) {
    int t = idx / dk, d = idx % dk;
    out[at(p, b, t, h) + d] = from_f32<T>(dkc[idx] - mean[d]);
  }
  if (map == 0 && threadIdx.x < 2 && p.dscalar_part) {
    float s = 0.f;
    for (int c = 0; c < w.nqb; ++c) s += ws[w.spart + ((size_t)bh * w.nqb + c) * 2 + threadIdx.x];
    p.dscalar_part[(size_t)bh * 2 + threadIdx.x] = s;
  }
}

}  // namespace quartet
}  // namespace mop
