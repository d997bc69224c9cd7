// Fused tcgen05/TMEM Edgewise forward for the config-1/2 hot shape (N = 64 tokens, dk <= 64, V <= 5 shared-projection views,
// low-rank gates r <= 4), second generation: CTAs of 256 threads (eight warps = 4 TMEM sub-partitions x 2 column halves),
// TWO CTAs per SM (101 KB of shared memory, 256 TMEM columns each), persistent over (batch, head) problems.
//
// Against ewtc::edgewise_kernel<false> (128 threads x 2 CTAs, 4.3 % of the tensor peak):
//   * sixteen warps per SM instead of eight: every element pass is split between the two column halves; a row softmax needs
//     ONE exchange between the two warps that share its rows (each half keeps exp(s - m_half); the halves swap (m, l, row sum)
//     through shared memory behind a 64-thread named barrier and rescale) - no CTA-wide barrier inside the softmax passes;
//   * raw Q, K, V tiles arrive by TMA (128-byte swizzle); K is an MMA operand as it lands;
//   * element loops are rolled over views / 8-column octets (x1 / x4 TMEM accesses): a few hundred instructions per phase;
//   * the small per-problem vectors the backward needs (row statistics, feature means, gate factors) are written to `aux`.
//
// Math: SURVEY.md appendix A (reference attention_variants.py:500-562, :319-331); specification oracle/edgewise_manual.py.
#pragma once
#include "edgewise_n64_bwd.cuh"

namespace mop {
namespace ew64 {

constexpr int kFwdThreads = 256;

// TMEM tiles of the forward (8 tiles of 64 columns x 16 lanes, two banks; ewtc::ttile<false>)
constexpr int fS = 0;   // S_k, 0..4
constexpr int fF = 5;   // chain product F (fp32)
constexpr int fR = 6;   // chain product R
constexpr int fY = 7;   // output accumulator

struct __align__(1024) SmemF {
  unsigned char Kr[kTile], Qr[kTile], Vr[kTile];   // raw tiles as TMA lands them; Q and K are reused by the chain once dead
  unsigned char A[kMaxV][kTile];                   // Q (.) c_k -> A_k; after the chain: A[0] = mixed map, A[1] = V_1, A[2] = w V_V
  unsigned char X[2][kTile];                       // chain ping-pong partners of Qr (forward chain) / Kr (reverse chain)
  float rho[kMaxC][64], kap[kMaxC][64];
  float ab[2][kMaxQ][64];                          // gate factors a, b; before that: column-sum partials red[7][4][64]
  float cvec[kMaxV][64], vs1[64], vsL[64];
  float exch[2][2][64][4];                         // [parity][column half][row]: max, sum, row sum (softmax exchange)
  float hw[2][kMaxQ * kMaxC + kMaxQ];
  uint64_t bar_mma, bar_in;
  uint32_t tmem_slot;
};
static_assert(sizeof(SmemF) + 1024 <= 115712, "forward shared memory: two CTAs per SM need <= 113 KB each");
static_assert(sizeof(float) * 7 * 4 * 64 <= sizeof(float) * 2 * kMaxQ * 64, "column-sum partials alias the gate-factor arrays");

__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}

static __global__ void __launch_bounds__(kFwdThreads, 2)
edgewise_fwd3_kernel(MopEdgewiseParams p, const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int sp = wid & 3, cb = wid >> 2;                          // TMEM sub-partition, column half
  const int row_lo = 16 * sp + (lane >> 2), row_hi = row_lo + 8, cq = 2 * (lane & 3);
  const int c0 = 32 * cb;                                         // my columns: c0 + 8n + cq + {0,1}, n = 0..3
  const int V = p.V, r = p.gate_rank, C = 2 * V + 2, dk = p.dk, H = p.H;
  const int ksteps = (dk + 15) >> 4;

  if (wid == 0) tmem_alloc<256>(&sm.tmem_slot);
  constexpr int kIssuers = 8;
  if (tid == 0) { mbar_init(&sm.bar_mma, kIssuers); mbar_init(&sm.bar_in, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const bool leader = lane == 0;
  auto mine = [&](int idx) { return (idx & (kIssuers - 1)) == wid; };
  const uint32_t tbase = sm.tmem_slot;
  const uint32_t tlane = tbase + ((uint32_t)(32 * sp) << 16);
  auto tcol = [&](int t) -> uint32_t { return tlane + ewtc::ttile<false>(t) + (uint32_t)c0; };
  uint32_t ph_mma = 0, ph_in = 0;
  const float w = 1.f / (1.f + __expf(-p.chain_value_logit[0]));
  const float bn = p.beta_not / (float)max(1, V - 1);
  const float sscale = rsqrtf((float)dk);
  const float eps = p.eps;
  auto sa = [&](const void* ptr) -> uint32_t { return smem_u32(ptr); };
  auto gemm = [&](int dt, uint32_t a_tile, int a_kind, uint32_t b_tile, int b_kind, bool acc, int ks) {
    const uint32_t id = idesc_bf16(64, 64, (a_kind & 1) ? 1u : 0u, (b_kind & 1) ? 1u : 0u);
    for (int k = 0; k < ks; ++k)
      mma_ss(tbase + ewtc::ttile<false>(dt), op_desc(a_kind, a_tile, k), op_desc(b_kind, b_tile, k), id, (acc || k > 0) ? 1u : 0u);
  };
  auto wait_mma = [&]() { mbar_wait(&sm.bar_mma, ph_mma); ph_mma ^= 1; tc_fence_after(); };
  // my 16 values (x4 fragment: octets n = 0..3) of a map <-> bf16 chunk-major tile
  const uint32_t o_lo = (uint32_t)(row_lo * 16 + 2 * cq), o_hi = o_lo + 128;
  auto put16 = [&](unsigned char* tile, const float* v) {
    unsigned char* t = tile + cb * 4096;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      *reinterpret_cast<uint32_t*>(t + n * 1024 + o_lo) = pack_bf16(v[4 * n + 0], v[4 * n + 1]);
      *reinterpret_cast<uint32_t*>(t + n * 1024 + o_hi) = pack_bf16(v[4 * n + 2], v[4 * n + 3]);
    }
  };
  // the two warps that share my rows (one per column half) meet here
  auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + sp) : "memory"); };
  // per-column sums of my fragment over the 16 rows of this warp -> red[sp][column]
  float (*red)[4][64] = reinterpret_cast<float (*)[4][64]>(&sm.ab[0][0][0]);
  auto colsum16 = [&](float (*dst)[64], const float* v) {
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float s = v[4 * n + e] + v[4 * n + 2 + e];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (lane < 4) dst[sp][c0 + 8 * n + cq + e] = s;
      }
  };
  // row softmax of my half-row fragments with ONE exchange: v becomes probabilities; returns (max * log2e, 1 / sum) per row,
  // `extra` (row sums of the raw values) is summed over the two halves on the way
  int xpar = 0;
  auto softmax16 = [&](float* v, float& m_lo, float& m_hi, float& il_lo, float& il_hi, float& ex_lo, float& ex_hi) {
    float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      mlo = fmaxf(mlo, fmaxf(v[4 * n], v[4 * n + 1]));
      mhi = fmaxf(mhi, fmaxf(v[4 * n + 2], v[4 * n + 3]));
    }
    mlo = quad_max(mlo) * kLog2e;
    mhi = quad_max(mhi) * kLog2e;
    float slo = 0.f, shi = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      v[4 * n + 0] = fast_exp2(fmaf(v[4 * n + 0], kLog2e, -mlo)); v[4 * n + 1] = fast_exp2(fmaf(v[4 * n + 1], kLog2e, -mlo));
      v[4 * n + 2] = fast_exp2(fmaf(v[4 * n + 2], kLog2e, -mhi)); v[4 * n + 3] = fast_exp2(fmaf(v[4 * n + 3], kLog2e, -mhi));
      slo += v[4 * n + 0] + v[4 * n + 1];
      shi += v[4 * n + 2] + v[4 * n + 3];
    }
    slo = quad_sum(slo);
    shi = quad_sum(shi);
    float (*mybuf)[4] = sm.exch[xpar][cb];
    float (*other)[4] = sm.exch[xpar][cb ^ 1];
    xpar ^= 1;
    if ((lane & 3) == 0) {
      *reinterpret_cast<float4*>(mybuf[row_lo]) = make_float4(mlo, slo, ex_lo, 0.f);
      *reinterpret_cast<float4*>(mybuf[row_hi]) = make_float4(mhi, shi, ex_hi, 0.f);
    }
    pair_sync();
    const float4 olo = *reinterpret_cast<const float4*>(other[row_lo]), ohi = *reinterpret_cast<const float4*>(other[row_hi]);
    m_lo = fmaxf(mlo, olo.x);
    m_hi = fmaxf(mhi, ohi.x);
    const float flo = fast_exp2(mlo - m_lo), fhi = fast_exp2(mhi - m_hi);
    il_lo = 1.f / fmaf(slo, flo, olo.y * fast_exp2(olo.x - m_lo));
    il_hi = 1.f / fmaf(shi, fhi, ohi.y * fast_exp2(ohi.x - m_hi));
    ex_lo += olo.z;
    ex_hi += ohi.z;
    const float klo = flo * il_lo, khi = fhi * il_hi;
#pragma unroll
    for (int n = 0; n < 4; ++n) { v[4 * n + 0] *= klo; v[4 * n + 1] *= klo; v[4 * n + 2] *= khi; v[4 * n + 3] *= khi; }
  };

  const int G = p.B * H;
  auto load_inputs = [&](int g) {
    mbar_expect_tx(&sm.bar_in, 3 * kTile);
    tma_load_tile_sw(sm.Qr, &tmQ, 0, g % H, g / H, &sm.bar_in);
    tma_load_tile_sw(sm.Kr, &tmK, 0, g % H, g / H, &sm.bar_in);
    tma_load_tile_sw(sm.Vr, &tmV, 0, g % H, g / H, &sm.bar_in);
  };
  if (tid == 0 && (int)blockIdx.x < G) load_inputs(blockIdx.x);
  {
    const int nW = 4 * r * C;
    for (int idx = tid; idx < 2 * (nW + 4 * r); idx += kFwdThreads) {
      const int half = idx / (nW + 4 * r), rem = idx % (nW + 4 * r);
      sm.hw[half][rem] = rem < nW ? (half ? p.col_w : p.row_w)[rem] : (half ? p.col_b : p.row_b)[rem - nW];
    }
  }
  const int hw_bias = 4 * r * C;

  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int pb = g / H, ph = g % H;
    float* aux = p.aux ? p.aux + (size_t)g * kAuxFloats : nullptr;
    // ---- phase 0: scale vectors, operand tiles --------------------------------------------------------------------------
    for (int idx = tid; idx < V * 64; idx += kFwdThreads) {
      const int i = idx >> 6, d = idx & 63;
      float c = 0.f;
      if (d < dk) c = sscale * p.q_scale[((size_t)i * H + ph) * dk + d] * p.k_scale[((size_t)i * H + ph) * dk + d];
      sm.cvec[i][d] = c;
    }
    if (tid < 64) {
      const int d = tid;
      float a = 0.f, b = 0.f;
      if (d < dk) { a = p.v_scale[((size_t)0 * H + ph) * dk + d]; b = p.v_scale[((size_t)(V - 1) * H + ph) * dk + d]; }
      sm.vs1[d] = a;
      sm.vsL[d] = w * b;
    }
    mbar_wait(&sm.bar_in, ph_in);
    ph_in ^= 1;
    __syncthreads();
    for (int idx = tid; idx < 64 * 8; idx += kFwdThreads) {
      const int rr = idx & 63, ch = idx >> 6;
      const uint4 q = *reinterpret_cast<const uint4*>(sm.Qr + sw128_off(rr, 8 * ch));
      for (int i = 0; i < V; ++i) *reinterpret_cast<uint4*>(sm.A[i] + ch * 1024 + rr * 16) = scale_chunk(q, &sm.cvec[i][ch * 8]);
    }
    publish();
    // ---- S_k = (Q c_k) K^T ----------------------------------------------------------------------------------------------
    if (leader) {
      for (int i = 0; i < V; ++i)
        if (mine(i)) gemm(fS + i, sa(sm.A[i]), CM_K, sa(sm.Kr), SW_K, false, ksteps);
      mma_commit(&sm.bar_mma);
    }
    wait_mma();
    // ---- per-view softmax (one exchange per view), row / column means of S_k -------------------------------------------------
    for (int k = 0; k < V; ++k) {
      float v[16];
      tmem_ld_x4(tcol(fS + k), v);
      tmem_ld_wait();
      float rs_lo = 0.f, rs_hi = 0.f;
#pragma unroll
      for (int n = 0; n < 4; ++n) { rs_lo += v[4 * n] + v[4 * n + 1]; rs_hi += v[4 * n + 2] + v[4 * n + 3]; }
      rs_lo = quad_sum(rs_lo);
      rs_hi = quad_sum(rs_hi);
      colsum16(red[k], v);
      float mlo, mhi, ilo, ihi;
      softmax16(v, mlo, mhi, ilo, ihi, rs_lo, rs_hi);
      put16(sm.A[k], v);
      if (cb == 0 && (lane & 3) == 0) {
        sm.rho[k][row_lo] = rs_lo * (1.f / 64.f);
        sm.rho[k][row_hi] = rs_hi * (1.f / 64.f);
        if (aux) {
          *reinterpret_cast<float2*>(aux + kAuxStats + (k * 64 + row_lo) * 2) = make_float2(mlo, ilo);
          *reinterpret_cast<float2*>(aux + kAuxStats + (k * 64 + row_hi) * 2) = make_float2(mhi, ihi);
        }
      }
    }
    // ---- chain products F = A_0..A_{V-1} (Qr <-> X[0]), R = A_{V-1}..A_0 (Kr <-> X[1]) ---------------------------------------
    uint32_t sF;
    {
      uint32_t xf = sa(sm.A[0]), xr = sa(sm.A[V - 1]);
      for (int s = 1; s < V; ++s) {
        publish();
        if (leader) {
          if (mine(0)) gemm(fF, xf, CM_K, sa(sm.A[s]), CM_MN, false, 4);
          if (mine(4)) gemm(fR, xr, CM_K, sa(sm.A[V - 1 - s]), CM_MN, false, 4);
          mma_commit(&sm.bar_mma);
        }
        wait_mma();
        unsigned char* pf = (s & 1) ? sm.Qr : sm.X[0];
        unsigned char* pr = (s & 1) ? sm.Kr : sm.X[1];
        float v[16];
        tmem_ld_x4(tcol(fF), v);
        tmem_ld_wait();
        put16(pf, v);
        xf = sa(pf);
        if (s < V - 1) {
          tmem_ld_x4(tcol(fR), v);
          tmem_ld_wait();
          put16(pr, v);
          xr = sa(pr);
        }
      }
      sF = xf;
    }
    // ---- log-chain features: row / column means of log(F + eps), log(R + eps) -------------------------------------------------
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      float v[16];
      tmem_ld_x4(tcol(which ? fR : fF), v);
      tmem_ld_wait();
      float slo = 0.f, shi = 0.f;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[4 * n + e] = kLn2 * fast_log2(v[4 * n + e] + eps);
        slo += v[4 * n] + v[4 * n + 1];
        shi += v[4 * n + 2] + v[4 * n + 3];
      }
      slo = quad_sum(slo);
      shi = quad_sum(shi);
      if ((lane & 3) == 0) { sm.exch[which][cb][row_lo][0] = slo; sm.exch[which][cb][row_hi][0] = shi; }
      colsum16(red[kMaxV + which], v);
    }
    __syncthreads();
    for (int idx = tid; idx < (V + 2) * 64; idx += kFwdThreads) {
      const int m = idx >> 6, j = idx & 63;
      const int slot = m < V ? m : kMaxV + (m - V), ch = m < V ? m : 2 * V + (m - V);
      sm.kap[ch][j] = ((red[slot][0][j] + red[slot][1][j]) + (red[slot][2][j] + red[slot][3][j])) * (1.f / 64.f);
      if (m >= V) sm.rho[ch][j] = (sm.exch[m - V][0][j][0] + sm.exch[m - V][1][j][0]) * (1.f / 64.f);
    }
    __syncthreads();
    // ---- low-rank gate factors: thread = (a | b, half of the 16 slots, token): the weight rows are warp-uniform (broadcast
    //      16-byte loads), the twelve feature means of the token sit in registers ---------------------------------------------
    {
      const int which = tid >> 7, half = (tid >> 6) & 1, tok = tid & 63;
      const float* W = sm.hw[which];
      const float* bias = W + hw_bias;
      float (*own)[64] = which ? sm.kap : sm.rho;
      float (*swp)[64] = which ? sm.rho : sm.kap;
      float ft[kMaxC];   // feature c as this projection sees it: S_c, S_c^T (roles swapped), log F, log R
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) {
        float v = 0.f;
        if (c < V) v = own[c][tok];
        else if (c < 2 * V) v = swp[c - V][tok];
        else if (c < C) v = own[c][tok];
        ft[c] = v;
      }
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int qq = 8 * half + j, t = qq >> 2, k = qq & 3, q = t * r + k;
        float a = 0.f;
        if (k < r) {
          a = bias[q];
          const float* wr = W + q * C;   // C = 2V + 2 is even: rows are 8-byte aligned, warp-uniform (broadcast) 8-byte loads
#pragma unroll
          for (int c = 0; c < kMaxC; c += 2)
            if (c < C) {
              const float2 w2 = *reinterpret_cast<const float2*>(wr + c);
              a = fmaf(w2.x, ft[c], fmaf(w2.y, ft[c + 1], a));
            }
        }
        acc[j] = a;
      }
      if (aux) {   // feature means for the backward
        for (int idx = tid; idx < 2 * C * 64; idx += kFwdThreads) {
          const int hf = idx / (C * 64), rem = idx % (C * 64);
          aux[(hf ? kAuxKap : kAuxRho) + rem] = (hf ? &sm.kap[0][0] : &sm.rho[0][0])[rem];
        }
      }
      __syncthreads();   // every thread has read the column-sum partials that a, b alias
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sm.ab[which][8 * half + j][tok] = acc[j];
        if (aux) aux[(which ? kAuxB : kAuxA) + (8 * half + j) * 64 + tok] = acc[j];
      }
    }
    __syncthreads();
    // ---- mix the score maps (one octet of 8 columns at a time), re-normalise ------------------------------------------------
    float amix[16];
    {
      float alo[kMaxQ], ahi[kMaxQ];
#pragma unroll
      for (int q = 0; q < kMaxQ; ++q) { alo[q] = sm.ab[0][q][row_lo]; ahi[q] = sm.ab[0][q][row_hi]; }
#pragma unroll
      for (int n = 0; n < 4; ++n) {   // (unrolled: amix[] stays in registers; the body is ~250 instructions)
        float s[kMaxV][4], fv[4];
#pragma unroll
        for (int i = 0; i < kMaxV; ++i)
          if (i < V) tmem_ld_x1(tcol(fS + i) + 8 * n, s[i]);
        tmem_ld_x1(tcol(fF) + 8 * n, fv);
        const int c = c0 + 8 * n + cq;
        float z[4][4];   // [gate][element]
#pragma unroll
        for (int t = 0; t < 4; ++t) { z[t][0] = z[t][1] = z[t][2] = z[t][3] = 0.f; }
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
          const float2 bb = *reinterpret_cast<const float2*>(&sm.ab[1][q][c]);
          z[q >> 2][0] = fmaf(alo[q], bb.x, z[q >> 2][0]);
          z[q >> 2][1] = fmaf(alo[q], bb.y, z[q >> 2][1]);
          z[q >> 2][2] = fmaf(ahi[q], bb.x, z[q >> 2][2]);
          z[q >> 2][3] = fmaf(ahi[q], bb.y, z[q >> 2][3]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float s0 = s[0][e];
          float sum = s0, mx = s0;
#pragma unroll
          for (int i = 1; i < kMaxV; ++i)
            if (i < V) { sum += s[i][e]; mx = fmaxf(mx, s[i][e]); }
          const float nm = -mx * kLog2e;
          float se = 0.f;
#pragma unroll
          for (int i = 0; i < kMaxV; ++i)
            if (i < V) se += fast_exp2(fmaf(s[i][e], kLog2e, nm));
          const float lse = fmaf(kLn2, fast_log2(se), mx);
          const float U = sum - s0, O = lse - s0, lf = kLn2 * fast_log2(fv[e] + eps);
          const float g0 = fast_sigmoid(z[0][e]), g1 = fast_sigmoid(z[1][e]), g2 = fast_sigmoid(z[2][e]), g3 = fast_sigmoid(z[3][e]);
          amix[4 * n + e] = fmaf(g3, lf, fmaf(g1, O, fmaf(fmaf(-bn, g2, g0), U, s0)));
        }
      }
    }
    {
      float mlo, mhi, ilo, ihi, d0 = 0.f, d1 = 0.f;
      softmax16(amix, mlo, mhi, ilo, ihi, d0, d1);
      if (aux && cb == 0 && (lane & 3) == 0) {
        *reinterpret_cast<float2*>(aux + kAuxStats + (V * 64 + row_lo) * 2) = make_float2(mlo, ilo);
        *reinterpret_cast<float2*>(aux + kAuxStats + (V * 64 + row_hi) * 2) = make_float2(mhi, ihi);
      }
    }
    put16(sm.A[0], amix);   // every A_k is dead: A[0] <- mixed map, A[1] <- V_1, A[2] <- w V_V
    for (int idx = tid; idx < 64 * 8; idx += kFwdThreads) {
      const int rr = idx & 63, ch = idx >> 6;
      const uint4 v = *reinterpret_cast<const uint4*>(sm.Vr + sw128_off(rr, 8 * ch));
      *reinterpret_cast<uint4*>(sm.A[1] + ch * 1024 + rr * 16) = scale_chunk(v, &sm.vs1[ch * 8]);
      *reinterpret_cast<uint4*>(sm.A[2] + ch * 1024 + rr * 16) = scale_chunk(v, &sm.vsL[ch * 8]);
    }
    // ---- y = A V_1 + w F V_V ------------------------------------------------------------------------------------------------
    publish();
    if (leader) {
      if (mine(0)) {   // both accumulate into the same tile: one issuer, in order
        gemm(fY, sa(sm.A[0]), CM_K, sa(sm.A[1]), CM_MN, false, 4);
        gemm(fY, sF, CM_K, sa(sm.A[2]), CM_MN, true, 4);
      }
      mma_commit(&sm.bar_mma);
    }
    wait_mma();
    {
      float yv[16];
      tmem_ld_x4(tcol(fY), yv);
      tmem_ld_wait();
      __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y);
      const size_t r_lo = (((size_t)pb * 64 + row_lo) * H + ph) * dk, r_hi = (((size_t)pb * 64 + row_hi) * H + ph) * dk;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int c = c0 + 8 * n + cq;
        if (c < dk) {
          *reinterpret_cast<uint32_t*>(y + r_lo + c) = pack_bf16(yv[4 * n], yv[4 * n + 1]);
          *reinterpret_cast<uint32_t*>(y + r_hi + c) = pack_bf16(yv[4 * n + 2], yv[4 * n + 3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // tiles, vectors and TMEM are reused by the next problem
    if (tid == 0 && g + (int)gridDim.x < G) load_inputs(g + gridDim.x);
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<256>(tbase);
}

}  // namespace ew64
}  // namespace mop
