// Plain / causal / biased / cross attention on tcgen05 / TMEM, second generation (bf16 operands, fp32 softmax).
//
// Replaces softmax(q k^T * scale [masks] [+ bias]) v of components.py:61-64, attention_variants.py:42-46,
// whisper_mop.py:163-175,212-219 for bf16 activations with dk % 8 == 0, dk <= 64 (models A/E, ViT-B/16, Whisper,
// GPT-2 heads).  Same structure as the Quartet kernels (quartet_tc.cuh): M=128 accumulators read one thread per
// row (32x32b), 64-wide streamed tiles double buffered with cp.async, every output owned by one CTA.
//   fwd       128 queries per CTA, online softmax in registers, P V in TMEM; saves lse
//   bwd_dq    128 queries per CTA (two warpgroups split the tile columns): S, dP by MMA, dS (bf16) -> dQ += dS K;
//             also leaves delta = dO . y for bwd_dkdv
//   bwd_dkdv  128 keys per CTA, transposed tiles (thread per key row, per-query lse / delta broadcast from shared
//             memory): dV += P^T dO, dK += dS^T Q
// Masks: keys >= Nk, zero-mask (mask == 0 -> -inf), causal fill, then the additive bias - in that order, as the
// fp32-mode kernels (sdpa_simt.cuh) do.  The kernels are compiled with and without
// the optional mask / bias tensors (EXTRA).
#pragma once
#include "quartet_tc.cuh"

namespace mop {
namespace sdpa2 {

using namespace tc;
using qtc::ex2;
using qtc::kLn2;
using qtc::kLog2e;
using qtc::kT128;
using qtc::kT64;
using qtc::lg2;
using qtc::load_act_tile;
using qtc::load_act_tile_async;
using qtc::pack8;
using qtc::publish;
using qtc::unpack8;

// optional masks of element (gi, gj); s is the scaled score (already -inf for padded keys)
__device__ __forceinline__ float apply_extra(const MopSdpaParams& p, int b, int h, int gi, int gj, float s) {
  if (p.zero_mask) {
    const float mk = p.zero_mask[(int64_t)b * p.zm_sb + (int64_t)h * p.zm_sh + (int64_t)gi * p.zm_sq + (int64_t)gj * p.zm_sk];
    if (mk == 0.f) s = -INFINITY;
  }
  if (p.causal && gj > gi) s = -INFINITY;
  if (p.bias) s += p.bias[(int64_t)b * p.bias_sb + (int64_t)h * p.bias_sh + (int64_t)gi * p.bias_sq + (int64_t)gj * p.bias_sk];
  return s;
}

struct __align__(128) SmemF {
  unsigned char Q[kT128], P[kT128];
  unsigned char K[2][kT64], V[2][kT64];
  uint64_t bar;      // MMA completion
  uint64_t ld[2];    // TMA completion of key / value buffer 0 / 1
  uint64_t ldq;      // TMA completion of the query tile
  uint32_t tmem_slot;
};

// grid: B*H*ceil(Nq/128), 128 threads; TMEM 128 columns (S | O): up to three CTAs per SM
// Q / K / V tiles arrive by TMA (tc_common.cuh: tma_load_tile): one instruction per tile, issued by the MMA thread one tile ahead.
template <bool EXTRA>
__global__ void __launch_bounds__(128, 3) fwd_kernel(MopSdpaParams p, const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                     const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, dk = p.dk, Nq = p.Nq, Nk = p.Nk;
  const int nqb = (Nq + 127) >> 7, BH = p.B * p.H;
  const int qb = nqb - 1 - (int)(blockIdx.x / (unsigned)BH), bh = blockIdx.x % BH, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * 128, gi = q0 + tid;
  const bool row_ok = gi < Nq;
  const int dks = (dk + 15) >> 4;
  if (warp == 0) tmem_alloc<128>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 1); mbar_init(&sm.ld[0], 1); mbar_init(&sm.ld[1], 1); mbar_init(&sm.ldq, 1); fence_mbar_init(); }
  const int k_end = p.causal ? min(Nk, q0 + 128) : Nk;
  const int ntiles = (k_end + 63) >> 6;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  auto fetch = [&](int buf, int k0) {   // thread 0 only
    mbar_expect_tx(&sm.ld[buf], 2 * kT64);
    tma_load_tile(sm.K[buf], &tmK, k0, h, b, &sm.ld[buf]);
    tma_load_tile(sm.V[buf], &tmV, k0, h, b, &sm.ld[buf]);
  };
  if (tid == 0) {
    mbar_expect_tx(&sm.ldq, kT128);
    tma_load_tile(sm.Q, &tmQ, q0, h, b, &sm.ldq);
    if (ntiles > 0) fetch(0, 0);
  }
  const uint32_t tb = sm.tmem_slot, tl = tb + ((uint32_t)(32 * warp) << 16);
  uint32_t phase = 0;
  float m_run = -INFINITY, l_run = 0.f;
  for (int it = 0; it < ntiles; ++it) {
    const int k0 = it * 64, buf = it & 1;
    if (it > 0) { mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after(); }   // P V of tile it-1: its buffers and P are free
    if (tid == 0) {
      if (it + 1 < ntiles) fetch(buf ^ 1, k0 + 64);
      if (it == 0) mbar_wait(&sm.ldq, 0);
      mbar_wait(&sm.ld[buf], (uint32_t)(it >> 1) & 1u);
      const uint32_t id = idesc_bf16(128, 64, 0, 0);
      for (int ks = 0; ks < dks; ++ks) mma_ss(tb, desc_kmajor(smem_u32(sm.Q), 128, 16 * ks), desc_kmajor(smem_u32(sm.K[buf]), 64, 16 * ks), id, ks > 0 ? 1u : 0u);
      mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
    float sc[64];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v1[16];
      tmem_ld_32x32b_x16(tl + 16 * c, v1);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) sc[16 * c + e] = v1[e] * p.scale;
    }
    if (k0 + 64 > Nk) {   // padded keys (uniform branch)
#pragma unroll
      for (int e = 0; e < 64; ++e)
        if (k0 + e >= Nk) sc[e] = -INFINITY;
    }
    if constexpr (EXTRA) {
      if (row_ok) {
#pragma unroll 8
        for (int e = 0; e < 64; ++e)
          if (k0 + e < Nk) sc[e] = apply_extra(p, b, h, gi, k0 + e, sc[e]);
      }
    } else {
      if (p.causal && k0 + 63 > q0) {   // diagonal tiles only
#pragma unroll
        for (int e = 0; e < 64; ++e)
          if (k0 + e > gi) sc[e] = -INFINITY;
      }
    }
    float tmax = -INFINITY;
#pragma unroll
    for (int e = 0; e < 64; ++e) tmax = fmaxf(tmax, sc[e]);
    const float m_new = fmaxf(m_run, tmax);
    const float mb = (m_new == -INFINITY) ? 0.f : m_new * kLog2e;
    const float corr = (m_run == -INFINITY) ? 0.f : ex2(fmaf(m_run, kLog2e, -mb));
    const bool need = it > 0 && m_new > m_run;
    if (__any_sync(0xffffffffu, need)) {
      const float scl = need ? corr : 1.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float o[16];
        tmem_ld_32x32b_x16(tl + 64 + 16 * c, o);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] *= scl;
        tmem_st_32x32b_x16(tl + 64 + 16 * c, o);
      }
      tmem_st_wait();
    }
    l_run *= corr;
    m_run = m_new;
    float ps = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float pv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { pv[e] = ex2(fmaf(sc[8 * c + e], kLog2e, -mb)); ps += pv[e]; }
      *reinterpret_cast<uint4*>(sm.P + c * (128 * 16) + tid * 16) = pack8(pv);
    }
    l_run += ps;
    publish();
    if (tid == 0) {
      const uint32_t id = idesc_bf16(128, 64, 0, 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_ss(tb + 64, desc_kmajor(smem_u32(sm.P), 128, 16 * ks), desc_mnmajor(smem_u32(sm.V[buf]), 64, 16 * ks), id, (it > 0 || ks > 0) ? 1u : 0u);
      mma_commit(&sm.bar);
    }
  }
  if (ntiles > 0) { mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after(); }
  const float il = 1.f / l_run;   // fully masked row: 0/0 = NaN like the reference softmax
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + (((int64_t)b * Nq + (row_ok ? gi : 0)) * p.H + h) * dk;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float o[16];
    if (ntiles > 0) {
      tmem_ld_32x32b_x16(tl + 64 + 16 * c, o);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) o[e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) o[e] *= il;
    if (row_ok) {
      if (16 * c < dk) *reinterpret_cast<uint4*>(y + 16 * c) = pack8(o);
      if (16 * c + 8 < dk) *reinterpret_cast<uint4*>(y + 16 * c + 8) = pack8(o + 8);
    }
  }
  if (p.lse && row_ok) p.lse[((int64_t)b * p.H + h) * Nq + gi] = m_run + kLn2 * lg2(l_run);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tb);
}

// probability and dS of one element given the raw dot product and dP
__device__ __forceinline__ void elem(const MopSdpaParams& p, float raw, float dp, float lse, float dlt, bool masked, float& pr, float& ds) {
  pr = masked ? 0.f : ex2(fmaf(raw, p.scale * kLog2e, -lse * kLog2e));
  ds = pr * (dp - dlt) * p.scale;
}
template <bool EXTRA>
__device__ __forceinline__ void elem_x(const MopSdpaParams& p, int b, int h, int gi, int gj, float raw, float dp, float lse, float dlt,
                                       bool masked, float& pr, float& ds) {
  if constexpr (EXTRA) {
    float s = masked ? -INFINITY : raw * p.scale;
    if (!masked) s = apply_extra(p, b, h, gi, gj, s);
    pr = (s == -INFINITY) ? 0.f : ex2((s - lse) * kLog2e);
    ds = pr * (dp - dlt) * p.scale;
  } else {
    elem(p, raw, dp, lse, dlt, masked || (p.causal && gj > gi), pr, ds);
  }
}

struct __align__(128) SmemQ {
  unsigned char Q[kT128], dO[kT128], W[kT128];
  unsigned char K[2][kT64], V[2][kT64];
  uint64_t bar;      // MMA completion
  uint64_t ld[2];    // TMA completion of key / value buffer 0 / 1
  uint64_t ldq;      // TMA completion of the query-side tiles
  uint32_t tmem_slot;
};

// grid: B*H*ceil(Nq/128), 256 threads; TMEM 256 columns (S | dP | dQ): two CTAs per SM
template <bool EXTRA>
__global__ void __launch_bounds__(256, 2) bwd_dq_kernel(MopSdpaParams p, float* delta, const __grid_constant__ CUtensorMap tmQ,
                                                        const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmK,
                                                        const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemQ& sm = *reinterpret_cast<SmemQ*>(smem_raw);
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp4 = (tid >> 5) & 3, dk = p.dk, Nq = p.Nq, Nk = p.Nk;
  const int nqb = (Nq + 127) >> 7, BH = p.B * p.H;
  const int qb = nqb - 1 - (int)(blockIdx.x / (unsigned)BH), bh = blockIdx.x % BH, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * 128, gi = q0 + t;
  const bool row_ok = gi < Nq;
  const int dks = (dk + 15) >> 4, c0 = 32 * wg;
  if (tid < 32) tmem_alloc<256>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 1); mbar_init(&sm.ld[0], 1); mbar_init(&sm.ld[1], 1); mbar_init(&sm.ldq, 1); fence_mbar_init(); }
  const size_t ystride = (size_t)p.H * dk;
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(p.dy) + ((int64_t)b * Nq * p.H + h) * dk;
  const __nv_bfloat16* yp = reinterpret_cast<const __nv_bfloat16*>(p.y) + ((int64_t)b * Nq * p.H + h) * dk;
  const int k_end = p.causal ? min(Nk, q0 + 128) : Nk;
  const int ntiles = (k_end + 63) >> 6;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  auto fetch = [&](int buf, int k0) {   // thread 0 only
    mbar_expect_tx(&sm.ld[buf], 2 * kT64);
    tma_load_tile(sm.K[buf], &tmK, k0, h, b, &sm.ld[buf]);
    tma_load_tile(sm.V[buf], &tmV, k0, h, b, &sm.ld[buf]);
  };
  if (tid == 0) {
    mbar_expect_tx(&sm.ldq, 2 * kT128);
    tma_load_tile(sm.Q, &tmQ, q0, h, b, &sm.ldq);
    tma_load_tile(sm.dO, &tmdO, q0, h, b, &sm.ldq);
    if (ntiles > 0) fetch(0, 0);
  }
  const float lse = p.lse[((int64_t)b * p.H + h) * Nq + (row_ok ? gi : Nq - 1)];
  float dlt = 0.f;
  if (row_ok) {
    for (int d0 = 0; d0 < dk; d0 += 8) {
      float a[8], c[8];
      unpack8(*reinterpret_cast<const uint4*>(yp + (size_t)gi * ystride + d0), a);
      unpack8(*reinterpret_cast<const uint4*>(dyp + (size_t)gi * ystride + d0), c);
#pragma unroll
      for (int e = 0; e < 8; ++e) dlt = fmaf(a[e], c[e], dlt);
    }
    if (wg == 0) delta[((int64_t)b * p.H + h) * Nq + gi] = dlt;
  }
  const uint32_t tb = sm.tmem_slot, tl = tb + ((uint32_t)(32 * warp4) << 16);
  uint32_t phase = 0;
  for (int it = 0; it < ntiles; ++it) {
    const int k0 = it * 64, buf = it & 1;
    if (it > 0) { mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after(); }
    if (tid == 0) {
      if (it + 1 < ntiles) fetch(buf ^ 1, k0 + 64);
      if (it == 0) mbar_wait(&sm.ldq, 0);
      mbar_wait(&sm.ld[buf], (uint32_t)(it >> 1) & 1u);
      const uint32_t id = idesc_bf16(128, 64, 0, 0);
      for (int ks = 0; ks < dks; ++ks) {
        mma_ss(tb, desc_kmajor(smem_u32(sm.Q), 128, 16 * ks), desc_kmajor(smem_u32(sm.K[buf]), 64, 16 * ks), id, ks > 0 ? 1u : 0u);
        mma_ss(tb + 64, desc_kmajor(smem_u32(sm.dO), 128, 16 * ks), desc_kmajor(smem_u32(sm.V[buf]), 64, 16 * ks), id, ks > 0 ? 1u : 0u);
      }
      mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col = c0 + 16 * c;
      float v1[16], dp[16], ws_[16];
      tmem_ld_32x32b_x16(tl + col, v1);
      tmem_ld_32x32b_x16(tl + 64 + col, dp);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int gj = k0 + col + e;
        float pr;
        elem_x<EXTRA>(p, b, h, gi, gj, v1[e], dp[e], lse, dlt, gj >= Nk || !row_ok, pr, ws_[e]);
      }
      const int ch = col >> 3;
      *reinterpret_cast<uint4*>(sm.W + ch * (128 * 16) + t * 16) = pack8(ws_);
      *reinterpret_cast<uint4*>(sm.W + (ch + 1) * (128 * 16) + t * 16) = pack8(ws_ + 8);
    }
    publish();
    if (tid == 0) {
      const uint32_t id = idesc_bf16(128, 64, 0, 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_ss(tb + 128, desc_kmajor(smem_u32(sm.W), 128, 16 * ks), desc_mnmajor(smem_u32(sm.K[buf]), 64, 16 * ks), id, (it > 0 || ks > 0) ? 1u : 0u);
      mma_commit(&sm.bar);
    }
  }
  if (ntiles > 0) { mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after(); }
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.dq) + (((int64_t)b * Nq + (row_ok ? gi : 0)) * p.H + h) * dk;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int col = c0 + 16 * c;
    float acc[16];
    if (ntiles > 0) {
      tmem_ld_32x32b_x16(tl + 128 + col, acc);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    }
    if (row_ok) {
      if (col < dk) *reinterpret_cast<uint4*>(out + col) = pack8(acc);
      if (col + 8 < dk) *reinterpret_cast<uint4*>(out + col + 8) = pack8(acc + 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<256>(tb);
}

struct __align__(128) SmemK {
  unsigned char K[kT128], V[kT128], PT[kT128], WT[kT128];
  unsigned char Q[2][kT64], dO[2][kT64];
  float vec[2][2][64];   // per query of the tile: lse, delta
  uint64_t bar;      // MMA completion
  uint64_t ld[2];    // TMA completion of query-side buffer 0 / 1
  uint64_t ldk;      // TMA completion of the key / value tiles
  uint32_t tmem_slot;
};

// grid: B*H*ceil(Nk/128), 256 threads (thread per key row; two warpgroups split the 64 query columns); TMEM 256 columns
// (S^T | dP^T | dV | dK): two CTAs per SM
template <bool EXTRA>
__global__ void __launch_bounds__(256, 2) bwd_dkdv_kernel(MopSdpaParams p, const float* delta, const __grid_constant__ CUtensorMap tmQ,
                                                          const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmK,
                                                          const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemK& sm = *reinterpret_cast<SmemK*>(smem_raw);
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp4 = (tid >> 5) & 3, dk = p.dk, Nq = p.Nq, Nk = p.Nk;
  const int nkb = (Nk + 127) >> 7;
  const int kb = blockIdx.x % nkb, bh = blockIdx.x / nkb, b = bh / p.H, h = bh % p.H;
  const int k0 = kb * 128, gj = k0 + t;
  const bool key_ok = gj < Nk;
  const int dks = (dk + 15) >> 4, c0 = 32 * wg;
  if (tid < 32) tmem_alloc<256>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 1); mbar_init(&sm.ld[0], 1); mbar_init(&sm.ld[1], 1); mbar_init(&sm.ldk, 1); fence_mbar_init(); }
  const float* lsep = p.lse + ((int64_t)b * p.H + h) * Nq;
  const float* dltp = delta + ((int64_t)b * p.H + h) * Nq;
  const int qstart = p.causal ? min(k0, Nq) & ~63 : 0;   // queries i >= j only when causal
  const int ntiles = (Nq - qstart + 63) >> 6;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // query-side tiles by TMA (thread 0); the per-query statistics by 4-byte cp.async (threads 0..63)
  auto fetch = [&](int buf, int q0) {
    if (tid == 0) {
      mbar_expect_tx(&sm.ld[buf], 2 * kT64);
      tma_load_tile(sm.Q[buf], &tmQ, q0, h, b, &sm.ld[buf]);
      tma_load_tile(sm.dO[buf], &tmdO, q0, h, b, &sm.ld[buf]);
    }
    if (tid < 64) {
      const int i = min(q0 + tid, Nq - 1);
      cp_async4(&sm.vec[buf][0][tid], lsep + i);
      cp_async4(&sm.vec[buf][1][tid], dltp + i);
    }
    cp_async_commit();
  };
  if (tid == 0) {
    mbar_expect_tx(&sm.ldk, 2 * kT128);
    tma_load_tile(sm.K, &tmK, k0, h, b, &sm.ldk);
    tma_load_tile(sm.V, &tmV, k0, h, b, &sm.ldk);
  }
  if (ntiles > 0) fetch(0, qstart);
  const uint32_t tb = sm.tmem_slot, tl = tb + ((uint32_t)(32 * warp4) << 16);
  uint32_t phase = 0;
  for (int it = 0; it < ntiles; ++it) {
    const int q0 = qstart + it * 64, buf = it & 1;
    if (it > 0) { mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after(); }
    if (it + 1 < ntiles) { fetch(buf ^ 1, q0 + 64); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();   // the per-query vectors of this tile are visible to every thread
    if (tid == 0) {
      if (it == 0) mbar_wait(&sm.ldk, 0);
      mbar_wait(&sm.ld[buf], (uint32_t)(it >> 1) & 1u);
      const uint32_t id = idesc_bf16(128, 64, 0, 0);
      for (int ks = 0; ks < dks; ++ks) {   // transposed tiles: rows = keys, columns = queries
        mma_ss(tb, desc_kmajor(smem_u32(sm.K), 128, 16 * ks), desc_kmajor(smem_u32(sm.Q[buf]), 64, 16 * ks), id, ks > 0 ? 1u : 0u);
        mma_ss(tb + 64, desc_kmajor(smem_u32(sm.V), 128, 16 * ks), desc_kmajor(smem_u32(sm.dO[buf]), 64, 16 * ks), id, ks > 0 ? 1u : 0u);
      }
      mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int colb = c0 + 16 * c;
      float v1[16], dp[16], pt[16], wt[16];
      tmem_ld_32x32b_x16(tl + colb, v1);
      tmem_ld_32x32b_x16(tl + 64 + colb, dp);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int col = colb + e, gi = q0 + col;
        elem_x<EXTRA>(p, b, h, gi, gj, v1[e], dp[e], sm.vec[buf][0][col], sm.vec[buf][1][col], gi >= Nq || !key_ok, pt[e], wt[e]);
      }
      const int ch = colb >> 3;
      *reinterpret_cast<uint4*>(sm.PT + ch * (128 * 16) + t * 16) = pack8(pt);
      *reinterpret_cast<uint4*>(sm.PT + (ch + 1) * (128 * 16) + t * 16) = pack8(pt + 8);
      *reinterpret_cast<uint4*>(sm.WT + ch * (128 * 16) + t * 16) = pack8(wt);
      *reinterpret_cast<uint4*>(sm.WT + (ch + 1) * (128 * 16) + t * 16) = pack8(wt + 8);
    }
    publish();
    if (tid == 0) {
      const uint32_t id = idesc_bf16(128, 64, 0, 1);
      const uint32_t acc0 = it > 0 ? 1u : 0u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {   // K index = queries of this tile
        mma_ss(tb + 128, desc_kmajor(smem_u32(sm.PT), 128, 16 * ks), desc_mnmajor(smem_u32(sm.dO[buf]), 64, 16 * ks), id, (acc0 || ks > 0) ? 1u : 0u);
        mma_ss(tb + 192, desc_kmajor(smem_u32(sm.WT), 128, 16 * ks), desc_mnmajor(smem_u32(sm.Q[buf]), 64, 16 * ks), id, (acc0 || ks > 0) ? 1u : 0u);
      }
      mma_commit(&sm.bar);
    }
  }
  if (ntiles > 0) { mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after(); }
  __nv_bfloat16* dv = reinterpret_cast<__nv_bfloat16*>(p.dv) + (((int64_t)b * Nk + (key_ok ? gj : 0)) * p.H + h) * dk;
  __nv_bfloat16* dkp = reinterpret_cast<__nv_bfloat16*>(p.dk_) + (((int64_t)b * Nk + (key_ok ? gj : 0)) * p.H + h) * dk;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int col = c0 + 16 * c;
    float av[16], ak[16];
    if (ntiles > 0) {
      tmem_ld_32x32b_x16(tl + 128 + col, av);
      tmem_ld_32x32b_x16(tl + 192 + col, ak);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) { av[e] = 0.f; ak[e] = 0.f; }
    }
    if (key_ok) {
      if (col < dk) { *reinterpret_cast<uint4*>(dv + col) = pack8(av); *reinterpret_cast<uint4*>(dkp + col) = pack8(ak); }
      if (col + 8 < dk) { *reinterpret_cast<uint4*>(dv + col + 8) = pack8(av + 8); *reinterpret_cast<uint4*>(dkp + col + 8) = pack8(ak + 8); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<256>(tb);
}

inline bool supported(const MopSdpaParams* p) {
  auto al8 = [](int64_t v) { return v % 8 == 0; };
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return p->dtype == MOP_BF16 && p->dk % 8 == 0 && p->dk <= 64 && al8(p->q_sb) && al8(p->q_sn) && al8(p->q_sh) && al8(p->k_sb) &&
         al8(p->k_sn) && al8(p->k_sh) && al8(p->v_sb) && al8(p->v_sn) && al8(p->v_sh) && al16(p->q) && al16(p->k) && al16(p->v);
}

}  // namespace sdpa2
}  // namespace mop
