// Plain / causal / biased / cross attention on tcgen05 / TMEM (bf16 operands, fp32 accumulation, fp32 softmax).
//
// Replaces softmax(q k^T * scale [masks] [+ bias]) v of components.py:61-64, attention_variants.py:42-46,
// whisper_mop.py:163-175,212-219 for bf16 activations with dk % 8 == 0, dk <= 64 (models A/E, ViT-B/16, Whisper,
// GPT-2 heads); everything else stays on the fp32-mode kernels of sdpa_simt.cuh.
//
// Flash-style 64x64 tiles built from the primitives of tc_common.cuh:
//   * one CTA (128 threads) owns a 64-row query tile; K/V tiles stream through shared memory;
//   * S = Q K^T and the PV / gradient products are M=64 tcgen05.mma into TMEM (<= 256 columns per CTA,
//     ~33-50 KB of shared memory) so that several CTAs are resident per SM and hide each other's latencies;
//   * 16x256b TMEM loads give every thread a 2-row x 16-column fragment: online-softmax statistics are quad
//     shuffles; masks (causal, mask==0, key padding) and the additive bias are applied on the fragment;
//   * backward = one kernel that owns a key tile (dK, dV accumulate in TMEM over the query tiles) and one that
//     owns a query tile (dQ accumulates in TMEM over the key tiles): every output is written by exactly one CTA.
#pragma once
#include "tc_common.cuh"

namespace mop {
namespace sdpatc {

using namespace tc;

constexpr int kTile = 64 * 64 * 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void publish() {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// rows [row0, row0+64) of a strided [.., n, h, :] bf16 tensor -> chunk-major tile (zero fill outside n < n_total, d < dk)
__device__ __forceinline__ void load_rows(unsigned char* tile, const __nv_bfloat16* base, int64_t sn, int row0, int n_total, int dk) {
  for (int idx = threadIdx.x; idx < 64 * 8; idx += 128) {
    const int r = idx & 63, ch = idx >> 6, n = row0 + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < n_total && ch * 8 < dk) v = *reinterpret_cast<const uint4*>(base + (int64_t)n * sn + ch * 8);
    *reinterpret_cast<uint4*>(tile + ch * 1024 + r * 16) = v;
  }
}

// score of element (gi, gj) after masks; s is the raw dot product.  Returns the value in NATURAL units (scaled).
__device__ __forceinline__ float masked_score(const MopSdpaParams& p, int b, int h, int gi, int gj, float s) {
  s *= p.scale;
  if (gj >= p.Nk) return -INFINITY;
  if (p.zero_mask && gi < p.Nq) {
    float mk = p.zero_mask[(int64_t)b * p.zm_sb + (int64_t)h * p.zm_sh + (int64_t)gi * p.zm_sq + (int64_t)gj * p.zm_sk];
    if (mk == 0.f) s = -INFINITY;
  }
  if (p.causal && gj > gi) s = -INFINITY;
  if (p.bias && gi < p.Nq) s += p.bias[(int64_t)b * p.bias_sb + (int64_t)h * p.bias_sh + (int64_t)gi * p.bias_sq + (int64_t)gj * p.bias_sk];
  return s;
}

struct __align__(1024) SmemFwd {
  unsigned char Q[kTile], K[kTile], V[kTile], P[kTile];
  uint64_t bar;
  uint32_t tmem_slot;
};

// grid: B*H*ceil(Nq/64)
__global__ void __launch_bounds__(128, 4) fwd_kernel(MopSdpaParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemFwd& sm = *reinterpret_cast<SmemFwd*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, dk = p.dk;
  const int nqb = (p.Nq + 63) >> 6;
  const int qb = blockIdx.x % nqb, bh = blockIdx.x / nqb, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * 64;
  const int ksteps = (dk + 15) >> 4;
  const Frag f;
  if (warp == 0) tmem_alloc<128>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
  const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(p.q) + (int64_t)b * p.q_sb + (int64_t)h * p.q_sh;
  const __nv_bfloat16* kp = reinterpret_cast<const __nv_bfloat16*>(p.k) + (int64_t)b * p.k_sb + (int64_t)h * p.k_sh;
  const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(p.v) + (int64_t)b * p.v_sb + (int64_t)h * p.v_sh;
  load_rows(sm.Q, qp, p.q_sn, q0, p.Nq, dk);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_slot, tlane = tbase + ((uint32_t)(32 * warp) << 16);
  uint32_t phase = 0;
  const uint32_t id_s = idesc_bf16(64, 64, 0, 0), id_pv = idesc_bf16(64, 64, 0, 1);
  float o[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  int k_end = p.Nk;
  if (p.causal) k_end = min(p.Nk, q0 + 64);
  for (int k0 = 0; k0 < k_end; k0 += 64) {
    load_rows(sm.K, kp, p.k_sn, k0, p.Nk, dk);
    load_rows(sm.V, vp, p.v_sn, k0, p.Nk, dk);
    publish();
    if (tid == 0) {
      for (int k = 0; k < ksteps; ++k)
        mma_ss(tbase, desc_kmajor(smem_u32(sm.Q), 64, 16 * k), desc_kmajor(smem_u32(sm.K), 64, 16 * k), id_s, k > 0 ? 1u : 0u);
      mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
    float s[32];
    tmem_ld_16x256b_x8(tlane, s);
    tmem_ld_wait();
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gi = q0 + ((e & 2) ? f.row_hi : f.row_lo), gj = k0 + f.col(n) + (e & 1);
        s[4 * n + e] = masked_score(p, b, h, gi, gj, s[4 * n + e]);
      }
      mx_lo = fmaxf(mx_lo, fmaxf(s[4 * n], s[4 * n + 1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[4 * n + 2], s[4 * n + 3]));
    }
    mx_lo = quad_max(mx_lo); mx_hi = quad_max(mx_hi);
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float base_lo = (mn_lo == -INFINITY) ? 0.f : mn_lo, base_hi = (mn_hi == -INFINITY) ? 0.f : mn_hi;
    const float c_lo = ex2((m_lo - base_lo) * kLog2e), c_hi = ex2((m_hi - base_hi) * kLog2e);
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[4 * n] = ex2((s[4 * n] - base_lo) * kLog2e);
      s[4 * n + 1] = ex2((s[4 * n + 1] - base_lo) * kLog2e);
      s[4 * n + 2] = ex2((s[4 * n + 2] - base_hi) * kLog2e);
      s[4 * n + 3] = ex2((s[4 * n + 3] - base_hi) * kLog2e);
      sum_lo += s[4 * n] + s[4 * n + 1];
      sum_hi += s[4 * n + 2] + s[4 * n + 3];
    }
    l_lo = l_lo * c_lo + quad_sum(sum_lo);
    l_hi = l_hi * c_hi + quad_sum(sum_hi);
    m_lo = mn_lo; m_hi = mn_hi;
    frag_store_bf16(sm.P, f, s);
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[4 * n] *= c_lo; o[4 * n + 1] *= c_lo; o[4 * n + 2] *= c_hi; o[4 * n + 3] *= c_hi; }
    publish();
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_ss(tbase + 64, desc_kmajor(smem_u32(sm.P), 64, 16 * k), desc_mnmajor(smem_u32(sm.V), 64, 16 * k), id_pv, k > 0 ? 1u : 0u);
      mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
    tmem_ld_16x256b_x8(tlane + 64, s);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] += s[i];
  }
  // epilogue
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y);
  const float i_lo = 1.f / l_lo, i_hi = 1.f / l_hi;   // fully masked row: 0/0 = NaN like the reference softmax
  const int r_lo = q0 + f.row_lo, r_hi = q0 + f.row_hi;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = f.col(n);
    if (c < dk) {
      if (r_lo < p.Nq) *reinterpret_cast<uint32_t*>(y + (((int64_t)b * p.Nq + r_lo) * p.H + h) * dk + c) = pack_bf16(o[4 * n] * i_lo, o[4 * n + 1] * i_lo);
      if (r_hi < p.Nq) *reinterpret_cast<uint32_t*>(y + (((int64_t)b * p.Nq + r_hi) * p.H + h) * dk + c) = pack_bf16(o[4 * n + 2] * i_hi, o[4 * n + 3] * i_hi);
    }
  }
  if (p.lse && (f.lane & 3) == 0) {
    if (r_lo < p.Nq) p.lse[((int64_t)b * p.H + h) * p.Nq + r_lo] = m_lo + kLn2 * lg2(l_lo);
    if (r_hi < p.Nq) p.lse[((int64_t)b * p.H + h) * p.Nq + r_hi] = m_hi + kLn2 * lg2(l_hi);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tbase);
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// delta[b,h,i] = sum_d dy[b,i,h,d] * y[b,i,h,d]; one warp per row
__global__ void __launch_bounds__(256) delta_kernel(MopSdpaParams p, float* delta) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int total = p.B * p.Nq * p.H;
  if (row >= total) return;
  const int h = row % p.H, i = (row / p.H) % p.Nq, b = row / (p.H * p.Nq);
  const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(p.y) + (int64_t)row * p.dk;
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(p.dy) + (int64_t)row * p.dk;
  float s = 0.f;
  for (int d = lane; d < p.dk; d += 32) s = fmaf(__bfloat162float(y[d]), __bfloat162float(dy[d]), s);
  s = warp_sum(s);
  if (lane == 0) delta[((int64_t)b * p.H + h) * p.Nq + i] = s;
}

struct __align__(1024) SmemBwd {
  unsigned char Q[kTile], K[kTile], V[kTile], dO[kTile], P[kTile], dS[kTile];
  float lse[64], dlt[64];
  uint64_t bar;
  uint32_t tmem_slot;
};

// Recompute P and dS (both as bf16 operand tiles) for query tile q0 x key tile k0.
// TMEM: S at column 0, dP at column 64 (both already computed by the caller's MMA batch).
__device__ __forceinline__ void p_and_ds(const MopSdpaParams& p, SmemBwd& sm, const Frag& f, uint32_t tlane, int b, int h, int q0, int k0) {
  float s[32], dp[32];
  tmem_ld_16x256b_x8(tlane, s);
  tmem_ld_16x256b_x8(tlane + 64, dp);
  tmem_ld_wait();
  const float ls_lo = sm.lse[f.row_lo], ls_hi = sm.lse[f.row_hi], dl_lo = sm.dlt[f.row_lo], dl_hi = sm.dlt[f.row_hi];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool hi = (e & 2) != 0;
      const int gi = q0 + (hi ? f.row_hi : f.row_lo), gj = k0 + f.col(n) + (e & 1);
      const float sc = masked_score(p, b, h, gi, gj, s[4 * n + e]);
      const float pr = (sc == -INFINITY || gi >= p.Nq) ? 0.f : ex2((sc - (hi ? ls_hi : ls_lo)) * kLog2e);
      s[4 * n + e] = pr;
      dp[4 * n + e] = pr * (dp[4 * n + e] - (hi ? dl_hi : dl_lo)) * p.scale;
    }
  }
  frag_store_bf16(sm.P, f, s);
  frag_store_bf16(sm.dS, f, dp);
}

__device__ __forceinline__ void load_row_stats(const MopSdpaParams& p, SmemBwd& sm, const float* delta, int b, int h, int q0) {
  if (threadIdx.x < 64) {
    const int i = q0 + threadIdx.x;
    const int64_t o = ((int64_t)b * p.H + h) * p.Nq + min(i, p.Nq - 1);
    sm.lse[threadIdx.x] = p.lse[o];
    sm.dlt[threadIdx.x] = delta[o];
  }
}

// grid: B*H*ceil(Nk/64).  Owns dK, dV of its key tile (TMEM columns 128.. and 192..).
__global__ void __launch_bounds__(128, 2) bwd_dkdv_kernel(MopSdpaParams p, const float* delta) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemBwd& sm = *reinterpret_cast<SmemBwd*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, dk = p.dk;
  const int nkb = (p.Nk + 63) >> 6;
  const int kb = blockIdx.x % nkb, bh = blockIdx.x / nkb, b = bh / p.H, h = bh % p.H;
  const int k0 = kb * 64, ksteps = (dk + 15) >> 4;
  const Frag f;
  if (warp == 0) tmem_alloc<256>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
  const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(p.q) + (int64_t)b * p.q_sb + (int64_t)h * p.q_sh;
  const __nv_bfloat16* kp = reinterpret_cast<const __nv_bfloat16*>(p.k) + (int64_t)b * p.k_sb + (int64_t)h * p.k_sh;
  const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(p.v) + (int64_t)b * p.v_sb + (int64_t)h * p.v_sh;
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(p.dy) + ((int64_t)b * p.Nq * p.H + h) * dk;
  load_rows(sm.K, kp, p.k_sn, k0, p.Nk, dk);
  load_rows(sm.V, vp, p.v_sn, k0, p.Nk, dk);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_slot, tlane = tbase + ((uint32_t)(32 * warp) << 16);
  uint32_t phase = 0;
  const uint32_t id_kk = idesc_bf16(64, 64, 0, 0), id_mm = idesc_bf16(64, 64, 1, 1);
  bool first = true;
  const int q_begin = p.causal ? k0 : 0;   // rows i >= j only
  for (int q0 = q_begin; q0 < p.Nq; q0 += 64) {
    load_rows(sm.Q, qp, p.q_sn, q0, p.Nq, dk);
    load_rows(sm.dO, dyp, (int64_t)p.H * dk, q0, p.Nq, dk);
    load_row_stats(p, sm, delta, b, h, q0);
    publish();
    if (tid == 0) {
      for (int k = 0; k < ksteps; ++k) {   // S = Q K^T ; dP = dO V^T   (rows = queries, cols = keys)
        mma_ss(tbase, desc_kmajor(smem_u32(sm.Q), 64, 16 * k), desc_kmajor(smem_u32(sm.K), 64, 16 * k), id_kk, k > 0 ? 1u : 0u);
        mma_ss(tbase + 64, desc_kmajor(smem_u32(sm.dO), 64, 16 * k), desc_kmajor(smem_u32(sm.V), 64, 16 * k), id_kk, k > 0 ? 1u : 0u);
      }
      mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
    p_and_ds(p, sm, f, tlane, b, h, q0, k0);
    publish();
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {        // dV += P^T dO ; dK += dS^T Q   (K index = queries)
        const uint32_t acc = (first && k == 0) ? 0u : 1u;
        mma_ss(tbase + 128, desc_mnmajor(smem_u32(sm.P), 64, 16 * k), desc_mnmajor(smem_u32(sm.dO), 64, 16 * k), id_mm, acc);
        mma_ss(tbase + 192, desc_mnmajor(smem_u32(sm.dS), 64, 16 * k), desc_mnmajor(smem_u32(sm.Q), 64, 16 * k), id_mm, acc);
      }
      mma_commit(&sm.bar);
    }
    first = false;
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();   // operand tiles are rewritten by the next iteration
  }
  float dv[32], dkk[32];
  if (first) {
#pragma unroll
    for (int i = 0; i < 32; ++i) { dv[i] = 0.f; dkk[i] = 0.f; }
  } else {
    tmem_ld_16x256b_x8(tlane + 128, dv);
    tmem_ld_16x256b_x8(tlane + 192, dkk);
    tmem_ld_wait();
  }
  __nv_bfloat16* dK = reinterpret_cast<__nv_bfloat16*>(p.dk_);
  __nv_bfloat16* dV = reinterpret_cast<__nv_bfloat16*>(p.dv);
  const int r_lo = k0 + f.row_lo, r_hi = k0 + f.row_hi;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = f.col(n);
    if (c < dk) {
      if (r_lo < p.Nk) {
        const int64_t o = (((int64_t)b * p.Nk + r_lo) * p.H + h) * dk + c;
        *reinterpret_cast<uint32_t*>(dK + o) = pack_bf16(dkk[4 * n], dkk[4 * n + 1]);
        *reinterpret_cast<uint32_t*>(dV + o) = pack_bf16(dv[4 * n], dv[4 * n + 1]);
      }
      if (r_hi < p.Nk) {
        const int64_t o = (((int64_t)b * p.Nk + r_hi) * p.H + h) * dk + c;
        *reinterpret_cast<uint32_t*>(dK + o) = pack_bf16(dkk[4 * n + 2], dkk[4 * n + 3]);
        *reinterpret_cast<uint32_t*>(dV + o) = pack_bf16(dv[4 * n + 2], dv[4 * n + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tbase);
}

// grid: B*H*ceil(Nq/64).  Owns dQ of its query tile (TMEM column 128..).
__global__ void __launch_bounds__(128, 2) bwd_dq_kernel(MopSdpaParams p, const float* delta) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemBwd& sm = *reinterpret_cast<SmemBwd*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, dk = p.dk;
  const int nqb = (p.Nq + 63) >> 6;
  const int qb = blockIdx.x % nqb, bh = blockIdx.x / nqb, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * 64, ksteps = (dk + 15) >> 4;
  const Frag f;
  if (warp == 0) tmem_alloc<256>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
  const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(p.q) + (int64_t)b * p.q_sb + (int64_t)h * p.q_sh;
  const __nv_bfloat16* kp = reinterpret_cast<const __nv_bfloat16*>(p.k) + (int64_t)b * p.k_sb + (int64_t)h * p.k_sh;
  const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(p.v) + (int64_t)b * p.v_sb + (int64_t)h * p.v_sh;
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(p.dy) + ((int64_t)b * p.Nq * p.H + h) * dk;
  load_rows(sm.Q, qp, p.q_sn, q0, p.Nq, dk);
  load_rows(sm.dO, dyp, (int64_t)p.H * dk, q0, p.Nq, dk);
  load_row_stats(p, sm, delta, b, h, q0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_slot, tlane = tbase + ((uint32_t)(32 * warp) << 16);
  uint32_t phase = 0;
  const uint32_t id_kk = idesc_bf16(64, 64, 0, 0), id_km = idesc_bf16(64, 64, 0, 1);
  bool first = true;
  int k_end = p.Nk;
  if (p.causal) k_end = min(p.Nk, q0 + 64);
  for (int k0 = 0; k0 < k_end; k0 += 64) {
    load_rows(sm.K, kp, p.k_sn, k0, p.Nk, dk);
    load_rows(sm.V, vp, p.v_sn, k0, p.Nk, dk);
    publish();
    if (tid == 0) {
      for (int k = 0; k < ksteps; ++k) {
        mma_ss(tbase, desc_kmajor(smem_u32(sm.Q), 64, 16 * k), desc_kmajor(smem_u32(sm.K), 64, 16 * k), id_kk, k > 0 ? 1u : 0u);
        mma_ss(tbase + 64, desc_kmajor(smem_u32(sm.dO), 64, 16 * k), desc_kmajor(smem_u32(sm.V), 64, 16 * k), id_kk, k > 0 ? 1u : 0u);
      }
      mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
    p_and_ds(p, sm, f, tlane, b, h, q0, k0);
    publish();
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)           // dQ += dS K   (K index = keys)
        mma_ss(tbase + 128, desc_kmajor(smem_u32(sm.dS), 64, 16 * k), desc_mnmajor(smem_u32(sm.K), 64, 16 * k), id_km, (first && k == 0) ? 0u : 1u);
      mma_commit(&sm.bar);
    }
    first = false;
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
  }
  float dq[32];
  if (first) {
#pragma unroll
    for (int i = 0; i < 32; ++i) dq[i] = 0.f;
  } else {
    tmem_ld_16x256b_x8(tlane + 128, dq);
    tmem_ld_wait();
  }
  __nv_bfloat16* dQ = reinterpret_cast<__nv_bfloat16*>(p.dq);
  const int r_lo = q0 + f.row_lo, r_hi = q0 + f.row_hi;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = f.col(n);
    if (c < dk) {
      if (r_lo < p.Nq) *reinterpret_cast<uint32_t*>(dQ + (((int64_t)b * p.Nq + r_lo) * p.H + h) * dk + c) = pack_bf16(dq[4 * n], dq[4 * n + 1]);
      if (r_hi < p.Nq) *reinterpret_cast<uint32_t*>(dQ + (((int64_t)b * p.Nq + r_hi) * p.H + h) * dk + c) = pack_bf16(dq[4 * n + 2], dq[4 * n + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tbase);
}

inline bool supported(const MopSdpaParams* p) {
  auto al8 = [](int64_t v) { return v % 8 == 0; };
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return p->dtype == MOP_BF16 && p->dk % 8 == 0 && p->dk <= 64 && al8(p->q_sb) && al8(p->q_sn) && al8(p->q_sh) && al8(p->k_sb) &&
         al8(p->k_sn) && al8(p->k_sh) && al8(p->v_sb) && al8(p->v_sn) && al8(p->v_sh) && al16(p->q) && al16(p->k) && al16(p->v);
}

}  // namespace sdpatc
}  // namespace mop
