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

constexpr int kStages = 4;   // key / value tile ring of the forward kernel
struct __align__(128) SmemF {
  unsigned char Q[kT128], P[2][kT128];
  unsigned char K[kStages][kT64], V[kStages][kT64];
  uint64_t bar_s[2];      // completion of the S MMA into score buffer 0 / 1            (tensor pipe -> softmax warps)
  uint64_t bar_pv[2];     // completion of P V of even / odd tiles                      (tensor pipe -> softmax warps, MMA warp)
  uint64_t p_ready[2];    // P(t) written into P[t & 1]: 128 arrivals                   (softmax warps -> P V warp)
  uint64_t s_free[2];     // score buffer t & 1 has been read: 128 arrivals             (softmax warps -> S warp)
  // every barrier is waited on phase by phase by each of its waiters, and its next completion depends on that waiter's
  // progress - a waiter can never fall two phases behind (a parity wait cannot tell phase k from phase k + 2)
  uint64_t ld[kStages];   // TMA completion of ring stage s
  uint64_t fr[kStages];   // ring stage s free again: P V of its tile complete               (tensor pipe -> TMA warp)
  uint64_t ldq;           // TMA completion of the query tile
  uint32_t tmem_slot;
};

// grid: B*H*ceil(Nq/128), 192 threads, two CTAs per SM; TMEM 256 columns (S buffer 0 | S buffer 1 | O).
// Warp-specialised, no CTA-wide barrier inside the loop:
//   warps 0-3  softmax, one query row per thread: S(t) from TMEM -> exponentials -> P(t) (bf16, two buffers) -> arrive p_ready
//   warp 4     one lane issues the TMA loads (Q once; K / V tiles into a four-stage ring, refilled when P V of the old tile is
//              done) and S(t) = Q K_t^T as soon as K_t has landed and score buffer t&1 has been read (two tiles ahead)
//   warp 5     one lane issues O += P(t) V_t when p_ready(t) completes
// The single-lane roles are split over two warps because a lone issuing lane needs ~10 cycles per instruction: one lane doing
// everything (~220 instructions per tile) was the bottleneck of the first pipelined version.  The running maximum is only
// raised when a tile exceeds it by more than 2^8 (exponentials stay <= 256; fp32 sums and bf16 products keep their relative
// precision), which removes nearly all rescales of O.
template <bool EXTRA, bool DROP = false>
static __global__ void __launch_bounds__(192, 2) fwd_kernel(MopSdpaParams p, const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                     const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();   // the 128-byte-swizzled TMA tiles need a 1024-byte aligned base
  const int tid = threadIdx.x, warp = tid >> 5, dk = p.dk, Nq = p.Nq, Nk = p.Nk;
  const int nqb = (Nq + 127) >> 7, BH = p.B * p.H;
  const int qb = nqb - 1 - (int)(blockIdx.x / (unsigned)BH), bh = blockIdx.x % BH, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * 128;
  if (warp == 0) tmem_alloc<256>(&sm.tmem_slot);
  if (tid == 0) {
    mbar_init(&sm.bar_s[0], 1); mbar_init(&sm.bar_s[1], 1); mbar_init(&sm.bar_pv[0], 1); mbar_init(&sm.bar_pv[1], 1); mbar_init(&sm.ldq, 1);
    mbar_init(&sm.p_ready[0], 128); mbar_init(&sm.p_ready[1], 128); mbar_init(&sm.s_free[0], 128); mbar_init(&sm.s_free[1], 128);
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.ld[s], 1); mbar_init(&sm.fr[s], 1); }
    fence_mbar_init();
  }
  const int k_end = p.causal ? min(Nk, q0 + 128) : Nk;
  const int ntiles = (k_end + 63) >> 6;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sm.tmem_slot;
  if (warp >= 4) {
    if (tid == 128) {
      // ---- TMA loads + S(t) = Q K_t^T ----------------------------------------------------------------------------
      auto fetch = [&](int t) {
        const int s = t & (kStages - 1);
        mbar_expect_tx(&sm.ld[s], 2 * kT64);
        tma_load_tile_sw(sm.K[s], &tmK, 64 * t, h, b, &sm.ld[s]);
        tma_load_tile_sw(sm.V[s], &tmV, 64 * t, h, b, &sm.ld[s]);
      };
      const uint32_t id_s = idesc_bf16(128, 64, 0, 0);
      const uint64_t dq = desc_k_sw(smem_u32(sm.Q), 0), dk0 = desc_k_sw(smem_u32(sm.K[0]), 0);
      if (ntiles > 0) {
        mbar_expect_tx(&sm.ldq, kT128);
        tma_load_tile_sw(sm.Q, &tmQ, q0, h, b, &sm.ldq);
        for (int t = 0; t < min(ntiles, kStages); ++t) fetch(t);
        mbar_wait(&sm.ldq, 0);
      }
      TS_DECL
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & (kStages - 1);
        TS(t, 0);
        if (t >= 2) { mbar_wait(&sm.s_free[t & 1], (uint32_t)((t - 2) >> 1) & 1u); tc_fence_after(); }   // S(t-2) has been read
        TS(t, 1);
        mbar_wait(&sm.ld[s], (uint32_t)(t >> 2) & 1u);
        TS(t, 2);
        const uint64_t dkk = dk0 + (uint64_t)(s * (kT64 >> 4));
        const uint32_t dst = tb + 64 * (t & 1);
        // one K step = 16 columns = +32 bytes inside the 128-byte swizzle row (>> 4 in the descriptor's address field);
        // the tiles are zero filled past dk, so the full 64-wide contraction is always right
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_ss(dst, dq + (uint64_t)(ks * 2), dkk + (uint64_t)(ks * 2), id_s, ks > 0 ? 1u : 0u);
        mma_commit(&sm.bar_s[t & 1]);
        TS(t, 3);
        if (t >= 2 && t + 2 < ntiles) {   // refill the ring stage of tile t-2 once P V(t-2) (issued right after p_ready(t-2)) is done
          mbar_wait(&sm.fr[(t - 2) & (kStages - 1)], (uint32_t)((t - 2) >> 2) & 1u);
          fetch(t + 2);
        }
        TS(t, 4);
      }
      TS_DUMP("S-lane", ntiles, 5);
    } else if (tid == 160) {
      // ---- O += P(t) V_t ---------------------------------------------------------------------------------------
      const uint32_t id_pv = idesc_bf16(128, 64, 0, 1);
      const uint64_t dp0 = desc_kmajor(smem_u32(sm.P[0]), 128, 0), dv0 = desc_mn_sw(smem_u32(sm.V[0]), 0);
      TS_DECL
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & (kStages - 1);
        TS(t, 0);
        mbar_wait(&sm.ld[s], (uint32_t)(t >> 2) & 1u);   // V_t (long since landed: S(t) already used K_t of the same stage)
        TS(t, 1);
        mbar_wait(&sm.p_ready[t & 1], (uint32_t)(t >> 1) & 1u);   // P(t) written into P[t & 1]
        tc_fence_after();
        TS(t, 2);
        const uint64_t dp = dp0 + (uint64_t)((t & 1) * (kT128 >> 4)), dv = dv0 + (uint64_t)(s * (kT64 >> 4));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_ss(tb + 128, dp + (uint64_t)(ks * 256), dv + (uint64_t)(ks * 128), id_pv, (t > 0 || ks > 0) ? 1u : 0u);
        mma_commit(&sm.bar_pv[t & 1]);
        mma_commit(&sm.fr[s]);
        TS(t, 3);
      }
      TS_DUMP("PV-lane", ntiles, 4);
    }
  } else {
    // ---- softmax warps: one query row per thread ------------------------------------------------------------
    const int gi = q0 + tid;
    const bool row_ok = gi < Nq;
    const uint32_t tl = tb + ((uint32_t)(32 * warp) << 16);
    const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
    const uint32_t rkey = dropout_row_key(drop, (uint32_t)bh, (uint32_t)gi);
    const float coef = EXTRA ? kLog2e : p.scale * kLog2e;   // exponent = score * coef - m_ref (base 2)
    float m_ref = -INFINITY, l_run = 0.f;
    TS_DECL
    for (int it = 0; it < ntiles; ++it) {
      const int k0 = it * 64;
      TS(it, 0);
      mbar_wait(&sm.bar_s[it & 1], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      TS(it, 1);
      float sc[64];
      tmem_ld_32x32b_x32(tl + 64 * (it & 1), sc);
      tmem_ld_32x32b_x32(tl + 64 * (it & 1) + 32, sc + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&sm.s_free[it & 1]);
      TS(it, 2);
      if (k0 + 64 > Nk) {   // padded keys (uniform branch)
#pragma unroll
        for (int e = 0; e < 64; ++e)
          if (k0 + e >= Nk) sc[e] = -INFINITY;
      }
      if constexpr (EXTRA) {
#pragma unroll
        for (int e = 0; e < 64; ++e) sc[e] *= p.scale;
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
      float t0 = sc[0], t1 = sc[1], t2 = sc[2], t3 = sc[3];
#pragma unroll
      for (int e = 4; e < 64; e += 4) { t0 = fmaxf(t0, sc[e]); t1 = fmaxf(t1, sc[e + 1]); t2 = fmaxf(t2, sc[e + 2]); t3 = fmaxf(t3, sc[e + 3]); }
      const float tm2 = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)) * coef;
      const bool need = tm2 > m_ref + 8.f;                       // -inf + 8 = -inf: the first finite tile always raises it
      const float m_new = need ? tm2 : m_ref;
      const float corr = need ? ((m_ref == -INFINITY) ? 0.f : ex2(m_ref - m_new)) : 1.f;
      const float mb = (m_new == -INFINITY) ? 0.f : m_new;
      TS(it, 3);
      if (it > 1) mbar_wait(&sm.bar_pv[it & 1], (uint32_t)((it - 2) >> 1) & 1u);
      TS(it, 4);   // P V(it-2) (issued a tile ago): P[it & 1] is free
      if (it > 0 && __any_sync(0xffffffffu, need)) {   // rare: O must be final up to tile it-1 before it is rescaled
        mbar_wait(&sm.bar_pv[(it - 1) & 1], (uint32_t)((it - 1) >> 1) & 1u);
        tc_fence_after();
        {
          float o[64];
          tmem_ld_32x32b_x32(tl + 128, o);
          tmem_ld_32x32b_x32(tl + 160, o + 32);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 64; ++e) o[e] *= corr;
          tmem_st_32x32b_x32(tl + 128, o);
          tmem_st_32x32b_x32(tl + 160, o + 32);
          tmem_st_wait();
        }
      }
      l_run *= corr;
      m_ref = m_new;
      float2 ps0 = make_float2(0.f, 0.f), ps1 = ps0;   // packed fp32 pairs: one FFMA2 / FADD2 per two elements
      const float2 coef2 = make_float2(coef, coef), nmb2 = make_float2(-mb, -mb);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float2 a = fma2(make_float2(sc[8 * c + e], sc[8 * c + e + 1]), coef2, nmb2);
          pv[e] = ex2(a.x); pv[e + 1] = ex2(a.y);
        }
        ps0 = add2(ps0, add2(make_float2(pv[0], pv[1]), make_float2(pv[2], pv[3])));
        ps1 = add2(ps1, add2(make_float2(pv[4], pv[5]), make_float2(pv[6], pv[7])));
        if constexpr (DROP) {   // the row sum above is that of the un-dropped probabilities; 1/(1-p) is folded into the final 1/l
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (!dropout_keep(rkey, (uint32_t)(k0 + 8 * c + e), drop.thresh)) pv[e] = 0.f;
        }
        *reinterpret_cast<uint4*>(sm.P[it & 1] + c * (128 * 16) + tid * 16) = pack8(pv);
      }
      l_run += (ps0.x + ps0.y) + (ps1.x + ps1.y);
      fence_async_smem();   // P (generic proxy) -> tensor pipe (async proxy)
      tc_fence_before();    // this thread's TMEM reads / writes precede the MMAs issued after the arrive is observed
      mbar_arrive(&sm.p_ready[it & 1]);
      TS(it, 5);
    }
    if (tid == 0) TS_DUMP("softmax", ntiles, 6);
    if (ntiles > 1) mbar_wait(&sm.bar_pv[ntiles & 1], (uint32_t)((ntiles - 2) >> 1) & 1u);
    if (ntiles > 0) { mbar_wait(&sm.bar_pv[(ntiles - 1) & 1], (uint32_t)((ntiles - 1) >> 1) & 1u); tc_fence_after(); }
    const float il = (DROP ? drop.inv_keep : 1.f) / l_run;   // fully masked row: 0/0 = NaN like the reference softmax
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + (((int64_t)b * Nq + (row_ok ? gi : 0)) * p.H + h) * dk;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float o[16];
      if (ntiles > 0) {
        tmem_ld_32x32b_x16(tl + 128 + 16 * c, o);
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
    if (p.lse && row_ok) p.lse[((int64_t)b * p.H + h) * Nq + gi] = kLn2 * (m_ref + lg2(l_run));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tb);
}

// probability and dS / scale of one element given the raw dot product and dP (the factor `scale` of dS is applied once, to the
// dQ / dK accumulators, instead of per element)
__device__ __forceinline__ void elem(const MopSdpaParams& p, float raw, float dp, float lse, float dlt, bool masked, float& pr, float& ds) {
  pr = masked ? 0.f : ex2(fmaf(raw, p.scale * kLog2e, -lse * kLog2e));
  ds = pr * (dp - dlt);
}
// the same for an unmasked pair of elements on packed fp32 math: 3 FMA-pipe instructions + 2 MUFU for two elements
__device__ __forceinline__ void elem2(float2 raw, float2 dp, float2 coef, float2 nlse, float2 ndlt, bool m0, bool m1, float2& pr, float2& ds) {
  const float2 a = fma2(raw, coef, nlse);
  pr = make_float2(m0 ? 0.f : ex2(a.x), m1 ? 0.f : ex2(a.y));   // m0 / m1: causal zeroing (diagonal tiles)
  ds = mul2(pr, add2(dp, ndlt));
}
template <bool EXTRA>
__device__ __forceinline__ void elem_x(const MopSdpaParams& p, int b, int h, int gi, int gj, float raw, float dp, float lse, float dlt,
                                       bool masked, float& pr, float& ds) {
  if constexpr (EXTRA) {
    float s = masked ? -INFINITY : raw * p.scale;
    if (!masked) s = apply_extra(p, b, h, gi, gj, s);
    pr = (s == -INFINITY) ? 0.f : ex2((s - lse) * kLog2e);
    ds = pr * (dp - dlt);
  } else {
    elem(p, raw, dp, lse, dlt, masked || (p.causal && gj > gi), pr, ds);
  }
}

struct __align__(128) SmemQ {
  unsigned char Q[kT128], dO[kT128], W[kT128];
  unsigned char K[2][kT64], V[2][kT64];
  uint64_t bar;      // completion of S and dP (two issuing threads)
  uint64_t bar2;     // completion of dQ += dS K
  uint64_t ld[2];    // TMA completion of key / value buffer 0 / 1
  uint64_t ldq;      // TMA completion of the query-side tiles
  uint32_t tmem_slot;
  float dlt[128];    // dO . y of the tile's query rows
};

// grid: B*H*ceil(Nq/128), 256 threads; TMEM 256 columns (S | dP | dQ): two CTAs per SM
template <bool EXTRA, bool DROP = false>
static __global__ void __launch_bounds__(256, 2) bwd_dq_kernel(MopSdpaParams p, float* delta, const __grid_constant__ CUtensorMap tmQ,
                                                        const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmK,
                                                        const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemQ& sm = *reinterpret_cast<SmemQ*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();   // the 128-byte-swizzled TMA tiles need a 1024-byte aligned base
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp4 = (tid >> 5) & 3, dk = p.dk, Nq = p.Nq, Nk = p.Nk;
  const int nqb = (Nq + 127) >> 7, BH = p.B * p.H;
  const int qb = nqb - 1 - (int)(blockIdx.x / (unsigned)BH), bh = blockIdx.x % BH, b = bh / p.H, h = bh % p.H;
  const int q0 = qb * 128, gi = q0 + t;
  const bool row_ok = gi < Nq;
  const int dks = (dk + 15) >> 4, c0 = 32 * wg;
  if (tid < 32) tmem_alloc<256>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 2); mbar_init(&sm.bar2, 1); mbar_init(&sm.ld[0], 1); mbar_init(&sm.ld[1], 1); mbar_init(&sm.ldq, 1); fence_mbar_init(); }
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
    tma_load_tile_sw(sm.K[buf], &tmK, k0, h, b, &sm.ld[buf]);
    tma_load_tile_sw(sm.V[buf], &tmV, k0, h, b, &sm.ld[buf]);
  };
  if (tid == 0) {
    mbar_expect_tx(&sm.ldq, 2 * kT128);
    tma_load_tile_sw(sm.Q, &tmQ, q0, h, b, &sm.ldq);
    tma_load_tile_sw(sm.dO, &tmdO, q0, h, b, &sm.ldq);
    if (ntiles > 0) fetch(0, 0);
  }
  const float lse = p.lse[((int64_t)b * p.H + h) * Nq + (row_ok ? gi : Nq - 1)];
  const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
  const uint32_t rkey = dropout_row_key(drop, (uint32_t)bh, (uint32_t)gi);
  // delta = dO . y per query row: eight lanes share a row (16 B of y and of dO per lane; four rows per warp instruction, all
  // 16 rows of a warp in flight) - one thread per 128-byte row costs 32 sectors per load and leaves eight dependent round trips
  {
    const int lane = tid & 31, wrp = tid >> 5, sub = lane >> 3, l8 = lane & 7;
    uint4 ya[4], da[4];
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      const int r = q0 + 16 * wrp + 4 * ps + sub;
      const bool ok = r < Nq && 8 * l8 < dk;
      const size_t o = (size_t)(ok ? r : 0) * ystride + 8 * l8;
      ya[ps] = ok ? __ldg(reinterpret_cast<const uint4*>(yp + o)) : make_uint4(0, 0, 0, 0);
      da[ps] = ok ? __ldg(reinterpret_cast<const uint4*>(dyp + o)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      float a[8], c[8];
      unpack8(ya[ps], a);
      unpack8(da[ps], c);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(a[e], c[e], d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      if (l8 == 0) sm.dlt[16 * wrp + 4 * ps + sub] = d;
    }
  }
  __syncthreads();
  const float dlt = row_ok ? sm.dlt[t] : 0.f;
  if (row_ok && wg == 0) delta[((int64_t)b * p.H + h) * Nq + gi] = dlt;
  const uint32_t tb = sm.tmem_slot, tl = tb + ((uint32_t)(32 * warp4) << 16);
  const float2 coef2 = make_float2(p.scale * kLog2e, p.scale * kLog2e), nlse2 = make_float2(-lse * kLog2e, -lse * kLog2e), ndlt2 = make_float2(-dlt, -dlt);
  uint32_t phase = 0, phase2 = 0;
  // The MMAs of one tile are issued by three threads of three different warps (S, dP, dQ): a lone issuing lane needs ~10 cycles
  // per instruction, so one thread issuing all twelve MMAs and the TMA loads kept the other seven warps waiting.
  const uint32_t id_in = idesc_bf16(128, 64, 0, 0), id_out = idesc_bf16(128, 64, 0, 1);
  const uint64_t d_q = desc_k_sw(smem_u32(sm.Q), 0), d_do = desc_k_sw(smem_u32(sm.dO), 0), d_w = desc_kmajor(smem_u32(sm.W), 128, 0);
  for (int it = 0; it < ntiles; ++it) {
    const int k0 = it * 64, buf = it & 1;
    if (it > 0) { mbar_wait(&sm.bar2, phase2); phase2 ^= 1; tc_fence_after(); }   // dQ(it-1): W and the other K / V buffer are free
    if (tid == 0) {
      if (it + 1 < ntiles) fetch(buf ^ 1, k0 + 64);
      if (it == 0) mbar_wait(&sm.ldq, 0);
      mbar_wait(&sm.ld[buf], (uint32_t)(it >> 1) & 1u);
      const uint64_t d_k = desc_k_sw(smem_u32(sm.K[buf]), 0);
      for (int ks = 0; ks < dks; ++ks) mma_ss(tb, d_q + (uint64_t)(2 * ks), d_k + (uint64_t)(2 * ks), id_in, ks > 0 ? 1u : 0u);
      mma_commit(&sm.bar);
    } else if (tid == 32) {
      if (it == 0) mbar_wait(&sm.ldq, 0);
      mbar_wait(&sm.ld[buf], (uint32_t)(it >> 1) & 1u);
      const uint64_t d_v = desc_k_sw(smem_u32(sm.V[buf]), 0);
      for (int ks = 0; ks < dks; ++ks) mma_ss(tb + 64, d_do + (uint64_t)(2 * ks), d_v + (uint64_t)(2 * ks), id_in, ks > 0 ? 1u : 0u);
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
      if constexpr (DROP) {   // dS = P (.) (M (.) dP - delta), M = dropout factor of the forward
#pragma unroll
        for (int e = 0; e < 16; ++e) dp[e] *= dropout_factor(drop, rkey, (uint32_t)(k0 + col + e));
      }
      if (!EXTRA) {
        // zero-filled key rows (>= Nk) contribute nothing to dS K, rows >= Nq are never written: only the causal diagonal masks
        const bool diag = p.causal && k0 + 63 > q0;
        const int lim = gi - k0 - col;   // element e is above the diagonal when e > lim
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float2 pr, ds;
          elem2(make_float2(v1[e], v1[e + 1]), make_float2(dp[e], dp[e + 1]), coef2, nlse2, ndlt2, diag && e > lim, diag && e + 1 > lim, pr, ds);
          ws_[e] = ds.x; ws_[e + 1] = ds.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int gj = k0 + col + e;
          float pr;
          elem_x<EXTRA>(p, b, h, gi, gj, v1[e], dp[e], lse, dlt, gj >= Nk || !row_ok, pr, ws_[e]);
        }
      }
      const int ch = col >> 3;
      *reinterpret_cast<uint4*>(sm.W + ch * (128 * 16) + t * 16) = pack8(ws_);
      *reinterpret_cast<uint4*>(sm.W + (ch + 1) * (128 * 16) + t * 16) = pack8(ws_ + 8);
    }
    publish();
    if (tid == 64) {
      const uint64_t d_kt = desc_mn_sw(smem_u32(sm.K[buf]), 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tb + 128, d_w + (uint64_t)(256 * ks), d_kt + (uint64_t)(128 * ks), id_out, (it > 0 || ks > 0) ? 1u : 0u);
      mma_commit(&sm.bar2);
    }
  }
  if (ntiles > 0) { mbar_wait(&sm.bar2, phase2); phase2 ^= 1; tc_fence_after(); }
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
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] *= p.scale;
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
  uint32_t rk[2][64];    // per query of the tile: dropout row key
  uint64_t bar;      // completion of S^T and dP^T (two issuing threads)
  uint64_t bar2;     // completion of dV += P^T dO and dK += dS^T Q (two issuing threads)
  uint64_t ld[2];    // TMA completion of query-side buffer 0 / 1
  uint64_t ldk;      // TMA completion of the key / value tiles
  uint32_t tmem_slot;
};

// grid: B*H*ceil(Nk/128), 256 threads (thread per key row; two warpgroups split the 64 query columns); TMEM 256 columns
// (S^T | dP^T | dV | dK): two CTAs per SM
template <bool EXTRA, bool DROP = false>
static __global__ void __launch_bounds__(256, 2) bwd_dkdv_kernel(MopSdpaParams p, const float* delta, const __grid_constant__ CUtensorMap tmQ,
                                                          const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmK,
                                                          const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemK& sm = *reinterpret_cast<SmemK*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();   // the 128-byte-swizzled TMA tiles need a 1024-byte aligned base
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp4 = (tid >> 5) & 3, dk = p.dk, Nq = p.Nq, Nk = p.Nk;
  const int nkb = (Nk + 127) >> 7;
  const int kb = blockIdx.x % nkb, bh = blockIdx.x / nkb, b = bh / p.H, h = bh % p.H;
  const int k0 = kb * 128, gj = k0 + t;
  const bool key_ok = gj < Nk;
  const int dks = (dk + 15) >> 4, c0 = 32 * wg;
  if (tid < 32) tmem_alloc<256>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 2); mbar_init(&sm.bar2, 2); mbar_init(&sm.ld[0], 1); mbar_init(&sm.ld[1], 1); mbar_init(&sm.ldk, 1); fence_mbar_init(); }
  const float* lsep = p.lse + ((int64_t)b * p.H + h) * Nq;
  const float* dltp = delta + ((int64_t)b * p.H + h) * Nq;
  const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
  const int qstart = p.causal ? min(k0, Nq) & ~63 : 0;   // queries i >= j only when causal
  const int ntiles = (Nq - qstart + 63) >> 6;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // query-side tiles by TMA (thread 0); the per-query statistics by 4-byte cp.async (threads 0..63)
  auto fetch = [&](int buf, int q0) {
    if (tid == 0) {
      mbar_expect_tx(&sm.ld[buf], 2 * kT64);
      tma_load_tile_sw(sm.Q[buf], &tmQ, q0, h, b, &sm.ld[buf]);
      tma_load_tile_sw(sm.dO[buf], &tmdO, q0, h, b, &sm.ld[buf]);
    }
    if (tid < 64) {
      const int i = min(q0 + tid, Nq - 1);
      cp_async4(&sm.vec[buf][0][tid], lsep + i);
      cp_async4(&sm.vec[buf][1][tid], dltp + i);
      if constexpr (DROP) sm.rk[buf][tid] = dropout_row_key(drop, (uint32_t)bh, (uint32_t)i);
    }
    cp_async_commit();
  };
  if (tid == 0) {
    mbar_expect_tx(&sm.ldk, 2 * kT128);
    tma_load_tile_sw(sm.K, &tmK, k0, h, b, &sm.ldk);
    tma_load_tile_sw(sm.V, &tmV, k0, h, b, &sm.ldk);
  }
  if (ntiles > 0) fetch(0, qstart);
  const uint32_t tb = sm.tmem_slot, tl = tb + ((uint32_t)(32 * warp4) << 16);
  uint32_t phase = 0, phase2 = 0;
  // four issuing threads in four warps (S^T, dP^T | dV, dK): see bwd_dq_kernel
  const uint32_t id_in = idesc_bf16(128, 64, 0, 0), id_out = idesc_bf16(128, 64, 0, 1);
  const uint64_t d_k = desc_k_sw(smem_u32(sm.K), 0), d_v = desc_k_sw(smem_u32(sm.V), 0);
  const uint64_t d_pt = desc_kmajor(smem_u32(sm.PT), 128, 0), d_wt = desc_kmajor(smem_u32(sm.WT), 128, 0);
  for (int it = 0; it < ntiles; ++it) {
    const int q0 = qstart + it * 64, buf = it & 1;
    if (it > 0) { mbar_wait(&sm.bar2, phase2); phase2 ^= 1; tc_fence_after(); }
    if (it + 1 < ntiles) { fetch(buf ^ 1, q0 + 64); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();   // the per-query vectors of this tile are visible to every thread
    if (tid == 0 || tid == 32) {   // transposed tiles: rows = keys, columns = queries
      if (it == 0) mbar_wait(&sm.ldk, 0);
      mbar_wait(&sm.ld[buf], (uint32_t)(it >> 1) & 1u);
      const uint64_t da = tid ? d_v : d_k, db = desc_k_sw(smem_u32(tid ? sm.dO[buf] : sm.Q[buf]), 0);
      for (int ks = 0; ks < dks; ++ks) mma_ss(tb + (tid ? 64 : 0), da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), id_in, ks > 0 ? 1u : 0u);
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
      uint32_t km = 0xFFFFu;   // keep bits of this thread's 16 (query) columns
      if constexpr (DROP) {
        km = 0u;
#pragma unroll
        for (int e = 0; e < 16; ++e) km |= (dropout_keep(sm.rk[buf][colb + e], (uint32_t)gj, drop.thresh) ? 1u : 0u) << e;
#pragma unroll
        for (int e = 0; e < 16; ++e) dp[e] = ((km >> e) & 1u) ? dp[e] * drop.inv_keep : 0.f;
      }
      if (!EXTRA) {
        // zero-filled query rows (>= Nq) contribute nothing to P^T dO / dS^T Q, keys >= Nk are never written: only the diagonal masks
        const bool diag = p.causal && q0 < k0 + 127;
        const int lim = gj - q0 - colb;   // query column x is above the diagonal (masked) when x < lim
        const float2 coef2 = make_float2(p.scale * kLog2e, p.scale * kLog2e), nl2e = make_float2(-kLog2e, -kLog2e), neg1 = make_float2(-1.f, -1.f);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(&sm.vec[buf][0][colb + e]), d4 = *reinterpret_cast<const float4*>(&sm.vec[buf][1][colb + e]);
          float2 pr, ds;
          elem2(make_float2(v1[e], v1[e + 1]), make_float2(dp[e], dp[e + 1]), coef2, mul2(make_float2(l4.x, l4.y), nl2e), mul2(make_float2(d4.x, d4.y), neg1),
                diag && e < lim, diag && e + 1 < lim, pr, ds);
          pt[e] = pr.x; pt[e + 1] = pr.y; wt[e] = ds.x; wt[e + 1] = ds.y;
          elem2(make_float2(v1[e + 2], v1[e + 3]), make_float2(dp[e + 2], dp[e + 3]), coef2, mul2(make_float2(l4.z, l4.w), nl2e), mul2(make_float2(d4.z, d4.w), neg1),
                diag && e + 2 < lim, diag && e + 3 < lim, pr, ds);
          pt[e + 2] = pr.x; pt[e + 3] = pr.y; wt[e + 2] = ds.x; wt[e + 3] = ds.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = colb + e, gi = q0 + col;
          elem_x<EXTRA>(p, b, h, gi, gj, v1[e], dp[e], sm.vec[buf][0][col], sm.vec[buf][1][col], gi >= Nq || !key_ok, pt[e], wt[e]);
        }
      }
      if constexpr (DROP) {   // dV = (M (.) P)^T dO
#pragma unroll
        for (int e = 0; e < 16; ++e) pt[e] = ((km >> e) & 1u) ? pt[e] * drop.inv_keep : 0.f;
      }
      const int ch = colb >> 3;
      *reinterpret_cast<uint4*>(sm.PT + ch * (128 * 16) + t * 16) = pack8(pt);
      *reinterpret_cast<uint4*>(sm.PT + (ch + 1) * (128 * 16) + t * 16) = pack8(pt + 8);
      *reinterpret_cast<uint4*>(sm.WT + ch * (128 * 16) + t * 16) = pack8(wt);
      *reinterpret_cast<uint4*>(sm.WT + (ch + 1) * (128 * 16) + t * 16) = pack8(wt + 8);
    }
    publish();
    if (tid == 64 || tid == 96) {   // K index = queries of this tile
      const bool second = tid == 96;
      const uint64_t da = second ? d_wt : d_pt, db = desc_mn_sw(smem_u32(second ? sm.Q[buf] : sm.dO[buf]), 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tb + (second ? 192 : 128), da + (uint64_t)(256 * ks), db + (uint64_t)(128 * ks), id_out, (it > 0 || ks > 0) ? 1u : 0u);
      mma_commit(&sm.bar2);
    }
  }
  if (ntiles > 0) { mbar_wait(&sm.bar2, phase2); phase2 ^= 1; tc_fence_after(); }
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
#pragma unroll
    for (int e = 0; e < 16; ++e) ak[e] *= p.scale;
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
