// Fused tcgen05/TMEM Edgewise backward, two warpgroups (256 threads) cooperating on ONE (batch, head) problem.
//
// Same math, shared-memory tiles and TMEM tiles as ewtc::edgewise_kernel<true> (edgewise_tc.cuh), but the
// element-wise work of every phase is split between the two warpgroups so that each SM sub-partition has two
// warps to switch between (the single-warpgroup kernel issues on 17 % of its cycles).  Both warpgroups can read
// every TMEM tile (a warp's lane window is warp_id % 4), so no data is duplicated:
//   * per-view softmax, feature terms, dS_k conversion:   views i with i % 2 == warpgroup
//   * chain products, log features, chain seeds, sweeps:   warpgroup 0 = forward chain F, warpgroup 1 = reverse chain R
//   * mix / gate-gradient pass, value-gradient epilogue:    column halves (row statistics exchanged through smem)
//   * final projections:                                     warpgroup 0 = dQ (+ scale partials), warpgroup 1 = dK
// The reverse-chain contributions to dS_k are written to their own TMEM tiles (one write each, no read-modify-write
// race with the forward-chain side) and summed when dS_k is converted to its bf16 operand tile.
#pragma once
#include "edgewise_tc.cuh"

namespace mop {
namespace ewtc {

// reverse-chain dS_k contributions (fp32): TMEM tiles that are dead once the chain seeds are built
__host__ __device__ constexpr int tileSR(int k) { return k < 3 ? 5 + k : 9 + k; }          // 5,6,7,12,13
__host__ __device__ constexpr int tileT2(int k) { return k < 4 ? 8 + k : 14; }             // T_k : 8,9,10,11,14
__host__ __device__ constexpr int tileU2(int k) { return k == 0 ? 15 : k - 1; }            // U_k : 15,0,1,2,3

struct SmemBwd2Extra {
  float pmax[2][64], psum[2][64], prs[2][64];
  float da2[2][kMaxQ][64];
};
struct __align__(1024) SmemBwd2 {
  unsigned char T[Slots<true>::N][kTile];
  SmemVec v;
  SmemBwdVec bv;
  SmemBwd2Extra x;
};

static __global__ void __launch_bounds__(256, 1) edgewise_bwd2_kernel(MopEdgewiseParams p) {
  using SB = Slots<true>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemBwd2& sm = *reinterpret_cast<SmemBwd2*>(smem_raw);
  SmemVec& sv_ = sm.v;
  SmemBwdVec& bv_ = sm.bv;
  SmemBwd2Extra& xv = sm.x;
  const int tid = threadIdx.x, wg = tid >> 7, warp4 = (tid >> 5) & 3;
  const int V = p.V, r = p.gate_rank, C = 2 * V + 2, dk = p.dk, H = p.H;
  const int ksteps = (dk + 15) >> 4;
  const Frag f;

  if (tid < 32) tmem_alloc<512>(&sv_.tmem_slot);
  // MMAs are issued by lane 0 of all eight warps, dealt round-robin (see edgewise_tc.cuh); every leader commits
  constexpr int kIssuers = 8;
  const int wid = tid >> 5;
  const bool leader = (tid & 31) == 0;
  auto mine = [&](int idx) { return (idx & (kIssuers - 1)) == wid; };
  if (tid == 0) { mbar_init(&sv_.bar, kIssuers); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sv_.tmem_slot;
  const uint32_t tlane = tbase + ((uint32_t)(32 * warp4) << 16);
  uint32_t phase = 0;
  const float w = 1.f / (1.f + __expf(-p.chain_value_logit[0]));
  const float bn = p.beta_not / (float)max(1, V - 1);
  const float sscale = rsqrtf((float)dk);

  auto tile = [&](int slot) -> unsigned char* { return sm.T[slot]; };
  auto taddr = [&](int slot) -> uint32_t { return smem_u32(sm.T[slot]); };
  auto gemm = [&](int dt, uint32_t dcol, uint32_t a_tile, bool a_mn, uint32_t b_tile, bool b_mn, bool acc, int ks, uint32_t n) {
    const uint32_t id = idesc_bf16(64, n, a_mn ? 1u : 0u, b_mn ? 1u : 0u);
    for (int k = 0; k < ks; ++k) {
      uint64_t ad = a_mn ? desc_mnmajor(a_tile, 64, 16 * k) : desc_kmajor(a_tile, 64, 16 * k);
      uint64_t bd = b_mn ? desc_mnmajor(b_tile, 64, 16 * k) : desc_kmajor(b_tile, 64, 16 * k);
      mma_ss(tbase + ttile<true>(dt) + dcol, ad, bd, id, (acc || k > 0) ? 1u : 0u);
    }
  };
  auto wait_mma = [&]() { mbar_wait(&sv_.bar, phase); phase ^= 1; tc_fence_after(); };
  auto ld_tile = [&](int t, float* v) { tmem_ld_16x256b_x8(tlane + ttile<true>(t), v); tmem_ld_wait(); };
  auto st_tile = [&](int t, const float* v) { tmem_st_16x256b_x8(tlane + ttile<true>(t), v); tmem_st_wait(); };

  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p.qkv);
  __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(p.dqkv);
  const size_t hd = (size_t)H * dk;
  const int G = p.B * H;
  {   // head weights: read through L2 once per CTA instead of once per use
    const int nW = 4 * r * C;
    for (int idx = tid; idx < 2 * (nW + 4 * r); idx += 256) {
      const int half = idx / (nW + 4 * r), rem = idx % (nW + 4 * r);
      sv_.hw[half][rem] = rem < nW ? (half ? p.col_w : p.row_w)[rem] : (half ? p.col_b : p.row_b)[rem - nW];
    }
    __syncthreads();
  }
  const float* hw_row = sv_.hw[0];
  const float* hw_col = sv_.hw[1];
  const int hw_bias = 4 * r * C;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int pb = g / H, ph = g % H;
    const size_t in_lo = (((size_t)pb * 64 + f.row_lo) * 3) * hd + (size_t)ph * dk;   // q row; +hd: k; +2hd: v
    const size_t in_hi = (((size_t)pb * 64 + f.row_hi) * 3) * hd + (size_t)ph * dk;
    // ---- stage 0 ----------------------------------------------------------------------------------------
    for (int idx = tid; idx < V * 64; idx += 256) {
      int i = idx >> 6, d = idx & 63;
      float c = 0.f;
      if (d < dk) c = sscale * p.q_scale[((size_t)i * H + ph) * dk + d] * p.k_scale[((size_t)i * H + ph) * dk + d];
      sv_.cvec[i][d] = c;
    }
    if (tid < 64) {
      const int d = tid;
      float a = 0.f, b = 0.f;
      if (d < dk) { a = p.v_scale[((size_t)0 * H + ph) * dk + d]; b = p.v_scale[((size_t)(V - 1) * H + ph) * dk + d]; }
      sv_.vs1[d] = a;
      sv_.vsL[d] = w * b;
    }
    __syncthreads();
    for (int idx = tid; idx < 64 * 8; idx += 256) {
      const int rr = idx & 63, ch = idx >> 6;
      const uint32_t off = ch * 1024 + rr * 16;
      uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, dyv = q;
      if (ch * 8 < dk) {
        const __nv_bfloat16* base = qkv + (((size_t)pb * 64 + rr) * 3) * hd + (size_t)ph * dk + ch * 8;
        q = *reinterpret_cast<const uint4*>(base);
        k = *reinterpret_cast<const uint4*>(base + hd);
        v = *reinterpret_cast<const uint4*>(base + 2 * hd);
        dyv = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + (((size_t)pb * 64 + rr) * H + ph) * dk + ch * 8);
      }
      *reinterpret_cast<uint4*>(tile(SB::K) + off) = k;
      *reinterpret_cast<uint4*>(tile(SB::V1) + off) = scale_chunk(v, &sv_.vs1[ch * 8]);
      *reinterpret_cast<uint4*>(tile(SB::VL) + off) = scale_chunk(v, &sv_.vsL[ch * 8]);
      for (int i = 0; i < V; ++i) *reinterpret_cast<uint4*>(tile(SB::QC + i) + off) = scale_chunk(q, &sv_.cvec[i][ch * 8]);
      *reinterpret_cast<uint4*>(tile(SB::DY) + off) = dyv;
    }
    publish();
    // ---- stage 1: S_i, dA ---------------------------------------------------------------------------------
    if (leader) {
      for (int i = 0; i < V; ++i)
        if (mine(i)) gemm(kTS + i, 0, taddr(SB::QC + i), false, taddr(SB::K), false, false, ksteps, 64);
      if (mine(V)) gemm(kTY, 0, taddr(SB::DY), false, taddr(SB::V1), false, false, ksteps, 64);
      mma_commit(&sv_.bar);
    }
    wait_mma();
    for (int i = wg; i < V; i += 2) {   // per-view softmax: views split between the warpgroups
      float v[32];
      ld_tile(kTS + i, v);
      float slo = 0.f, shi = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) { slo += v[4 * n] + v[4 * n + 1]; shi += v[4 * n + 2] + v[4 * n + 3]; }
      slo = quad_sum(slo); shi = quad_sum(shi);
      if ((f.lane & 3) == 0) { sv_.rho[i][f.row_lo] = slo * (1.f / 64.f); sv_.rho[i][f.row_hi] = shi * (1.f / 64.f); }
      colsum_to(sv_.red[i], f, v);
      frag_softmax(v);
      frag_store_bf16(tile(SB::A + i), f, v);
    }
    // ---- chain products: warpgroup 0 keeps the forward prefixes, warpgroup 1 the reverse suffixes ---------------
    uint32_t sF;
    {
      uint32_t xf = taddr(SB::A + 0), xr = taddr(SB::A + V - 1);
      for (int s = 1; s < V; ++s) {
        publish();
        if (leader) {
          if (mine(0)) gemm(kTF, 0, xf, false, taddr(SB::A + s), true, false, 4, 64);
          if (mine(4)) gemm(kTR, 0, xr, false, taddr(SB::A + V - 1 - s), true, false, 4, 64);   // a warp of the other warpgroup
          mma_commit(&sv_.bar);
        }
        wait_mma();
        float v[32];
        if (wg == 0) {
          ld_tile(kTF, v);
          frag_store_bf16(tile(SB::P(s)), f, v);
        } else if (s < V - 1) {
          ld_tile(kTR, v);
          frag_store_bf16(tile(SB::R(s)), f, v);
        }
        xf = taddr(SB::P(s));
        if (s < V - 1) xr = taddr(SB::R(s));
      }
      sF = xf;
    }
    {
      float v[32];
      ld_tile(wg ? kTR : kTF, v);
      float slo = 0.f, shi = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[4 * n + e] = fast_log(v[4 * n + e] + p.eps);
        slo += v[4 * n] + v[4 * n + 1];
        shi += v[4 * n + 2] + v[4 * n + 3];
      }
      slo = quad_sum(slo); shi = quad_sum(shi);
      if ((f.lane & 3) == 0) {
        sv_.rho[2 * V + wg][f.row_lo] = slo * (1.f / 64.f);
        sv_.rho[2 * V + wg][f.row_hi] = shi * (1.f / 64.f);
      }
      colsum_to(sv_.red[kMaxV + wg], f, v);
    }
    __syncthreads();
    for (int idx = tid; idx < (V + 2) * 64; idx += 256) {
      int m = idx >> 6, j = idx & 63;
      int slot = m < V ? m : kMaxV + (m - V);
      float s = sv_.red[slot][0][j] + sv_.red[slot][1][j] + sv_.red[slot][2][j] + sv_.red[slot][3][j];
      sv_.kap[m < V ? m : 2 * V + (m - V)][j] = s * (1.f / 64.f);
    }
    __syncthreads();
    // ---- gate factors: thread = (a|b, token, half of the 16 slots) ------------------------------------------------
    {
      const int which = tid >> 7, tok = (tid & 127) >> 1, half = tid & 1;
      const float* W = which ? hw_col : hw_row;
      const float* bias = W + hw_bias;
      float (*own)[64] = which ? sv_.kap : sv_.rho;
      float (*swp)[64] = which ? sv_.rho : sv_.kap;
      for (int qq = 8 * half; qq < 8 * half + 8; ++qq) {
        const int t = qq >> 2, k = qq & 3, q = t * r + k;
        float acc = 0.f;
        if (k < r) {
          acc = bias[q];
          for (int c = 0; c < V; ++c) {
            acc = fmaf(W[q * C + c], own[c][tok], acc);
            acc = fmaf(W[q * C + V + c], swp[c][tok], acc);
          }
          acc = fmaf(W[q * C + 2 * V], own[2 * V][tok], acc);
          acc = fmaf(W[q * C + 2 * V + 1], own[2 * V + 1][tok], acc);
        }
        (which ? sv_.b : sv_.a)[qq][tok] = acc;
        if (which == 0) *reinterpret_cast<__nv_bfloat16*>(bv_.a_bf + tile_off(64, tok, qq)) = __float2bfloat16_rn(acc);
      }
    }
    __syncthreads();
    // ---- mix + gate-gradient pass: each warpgroup owns 32 columns (blocks 2wg, 2wg+1) of every row ----------------
    float alo[kMaxQ], ahi[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) { alo[q] = sv_.a[q][f.row_lo]; ahi[q] = sv_.a[q][f.row_hi]; }
    float am[16];   // this thread's 16 elements: [2 blocks][8]
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const int blk = 2 * wg + bb;
      float sv[kMaxV][8], fv[8];
#pragma unroll
      for (int i = 0; i < kMaxV; ++i)
        if (i < V) tmem_ld_16x256b_x2(tlane + ttile<true>(kTS + i) + 16 * blk, sv[i]);
      tmem_ld_16x256b_x2(tlane + ttile<true>(kTF) + 16 * blk, fv);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = 16 * blk + 8 * (e >> 2) + f.cq + (e & 1);
        const bool hi = (e & 2) != 0;
        float s0 = sv[0][e], sum = s0, mx = s0;
#pragma unroll
        for (int i = 1; i < kMaxV; ++i)
          if (i < V) { sum += sv[i][e]; mx = fmaxf(mx, sv[i][e]); }
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxV; ++i)
          if (i < V) se += fast_exp2((sv[i][e] - mx) * kLog2e);
        const float lse = mx + fast_log(se);
        const float U = sum - s0, O = lse - s0, lf = fast_log(fv[e] + p.eps);
        float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) z[q >> 2] = fmaf(hi ? ahi[q] : alo[q], sv_.b[q][col], z[q >> 2]);
        am[8 * bb + e] = s0 + fast_sigmoid(z[0]) * U + fast_sigmoid(z[1]) * O - fast_sigmoid(z[2]) * bn * U + fast_sigmoid(z[3]) * lf;
      }
    }
    {
      // row softmax across the two column halves
      float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; ++i) { if (i & 2) mhi = fmaxf(mhi, am[i]); else mlo = fmaxf(mlo, am[i]); }
      mlo = quad_max(mlo); mhi = quad_max(mhi);
      if ((f.lane & 3) == 0) { xv.pmax[wg][f.row_lo] = mlo; xv.pmax[wg][f.row_hi] = mhi; }
      __syncthreads();
      mlo = fmaxf(xv.pmax[0][f.row_lo], xv.pmax[1][f.row_lo]);
      mhi = fmaxf(xv.pmax[0][f.row_hi], xv.pmax[1][f.row_hi]);
      float elo = 0.f, ehi = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        am[i] = fast_exp2((am[i] - ((i & 2) ? mhi : mlo)) * kLog2e);
        if (i & 2) ehi += am[i]; else elo += am[i];
      }
      elo = quad_sum(elo); ehi = quad_sum(ehi);
      if ((f.lane & 3) == 0) { xv.psum[wg][f.row_lo] = elo; xv.psum[wg][f.row_hi] = ehi; }
      __syncthreads();
      const float ilo = 1.f / (xv.psum[0][f.row_lo] + xv.psum[1][f.row_lo]), ihi = 1.f / (xv.psum[0][f.row_hi] + xv.psum[1][f.row_hi]);
#pragma unroll
      for (int i = 0; i < 16; ++i) am[i] *= (i & 2) ? ihi : ilo;
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          const int c = 16 * (2 * wg + bb) + 8 * nn + f.cq;
          *reinterpret_cast<uint32_t*>(tile(SB::AMIX) + tile_off(64, f.row_lo, c)) = pack_bf16(am[8 * bb + 4 * nn], am[8 * bb + 4 * nn + 1]);
          *reinterpret_cast<uint32_t*>(tile(SB::AMIX) + tile_off(64, f.row_hi, c)) = pack_bf16(am[8 * bb + 4 * nn + 2], am[8 * bb + 4 * nn + 3]);
        }
    }
    // D = A (.) (dA - rowsum(dA (.) A))
    float D[16];
    {
      tmem_ld_16x256b_x2(tlane + ttile<true>(kTY) + 16 * (2 * wg), D);
      tmem_ld_16x256b_x2(tlane + ttile<true>(kTY) + 16 * (2 * wg + 1), D + 8);
      tmem_ld_wait();
      float rlo = 0.f, rhi = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { if (i & 2) rhi = fmaf(D[i], am[i], rhi); else rlo = fmaf(D[i], am[i], rlo); }
      rlo = quad_sum(rlo); rhi = quad_sum(rhi);
      if ((f.lane & 3) == 0) { xv.prs[wg][f.row_lo] = rlo; xv.prs[wg][f.row_hi] = rhi; }
      __syncthreads();
      rlo = xv.prs[0][f.row_lo] + xv.prs[1][f.row_lo];
      rhi = xv.prs[0][f.row_hi] + xv.prs[1][f.row_hi];
#pragma unroll
      for (int i = 0; i < 16; ++i) D[i] = am[i] * (D[i] - ((i & 2) ? rhi : rlo));
    }
    {
      float da_lo[kMaxQ], da_hi[kMaxQ];
#pragma unroll
      for (int q = 0; q < kMaxQ; ++q) { da_lo[q] = 0.f; da_hi[q] = 0.f; }
#pragma unroll 1
      for (int bb = 0; bb < 2; ++bb) {
        const int blk = 2 * wg + bb;
        float sv[kMaxV][8], fv[8];
#pragma unroll
        for (int i = 0; i < kMaxV; ++i)
          if (i < V) tmem_ld_16x256b_x2(tlane + ttile<true>(kTS + i) + 16 * blk, sv[i]);
        tmem_ld_16x256b_x2(tlane + ttile<true>(kTF) + 16 * blk, fv);
        tmem_ld_wait();
        float hf[8], dgv[4][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = 16 * blk + 8 * (e >> 2) + f.cq + (e & 1);
          const bool hi = (e & 2) != 0;
          const float d = bb ? D[8 + e] : D[e];
          float s0 = sv[0][e], sum = s0, mx = s0;
#pragma unroll
          for (int i = 1; i < kMaxV; ++i)
            if (i < V) { sum += sv[i][e]; mx = fmaxf(mx, sv[i][e]); }
          float ex[kMaxV], se = 0.f;
#pragma unroll
          for (int i = 0; i < kMaxV; ++i)
            if (i < V) { ex[i] = fast_exp2((sv[i][e] - mx) * kLog2e); se += ex[i]; }
          const float inv_se = fast_rcp(se);
          const float lse = mx + fast_log(se);
          const float U = sum - s0, O = lse - s0;
          const float fe = fv[e] + p.eps, lf = fast_log(fe);
          float z[4] = {0.f, 0.f, 0.f, 0.f}, bq[kMaxQ];
#pragma unroll
          for (int q = 0; q < kMaxQ; ++q) { bq[q] = sv_.b[q][col]; z[q >> 2] = fmaf(hi ? ahi[q] : alo[q], bq[q], z[q >> 2]); }
          const float g0 = fast_sigmoid(z[0]), g1 = fast_sigmoid(z[1]), g2 = fast_sigmoid(z[2]), g3 = fast_sigmoid(z[3]);
          dgv[0][e] = d * U * g0 * (1.f - g0);
          dgv[1][e] = d * O * g1 * (1.f - g1);
          dgv[2][e] = -bn * d * U * g2 * (1.f - g2);
          dgv[3][e] = d * lf * g3 * (1.f - g3);
#pragma unroll
          for (int q = 0; q < kMaxQ; ++q) {
            if (hi) da_hi[q] = fmaf(dgv[q >> 2][e], bq[q], da_hi[q]);
            else da_lo[q] = fmaf(dgv[q >> 2][e], bq[q], da_lo[q]);
          }
          hf[e] = d * g3 * fast_rcp(fe);
          const float e1 = d * g1, e0 = d * (g0 - g2 * bn);
#pragma unroll
          for (int i = 0; i < kMaxV; ++i)
            if (i < V) {
              const float pi = ex[i] * inv_se;
              sv[i][e] = (i == 0) ? (d - e1 + e1 * pi) : (e0 + e1 * pi);
            }
        }
#pragma unroll
        for (int i = 0; i < kMaxV; ++i)
          if (i < V) tmem_st_16x256b_x2(tlane + ttile<true>(kTS + i) + 16 * blk, sv[i]);
        tmem_st_16x256b_x2(tlane + ttile<true>(kTY) + 16 * blk, hf);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int nn = 0; nn < 2; ++nn) {
            const int c = 16 * blk + 8 * nn + f.cq;
            *reinterpret_cast<uint32_t*>(tile(SB::X + t) + tile_off(64, f.row_lo, c)) = pack_bf16(dgv[t][4 * nn], dgv[t][4 * nn + 1]);
            *reinterpret_cast<uint32_t*>(tile(SB::X + t) + tile_off(64, f.row_hi, c)) = pack_bf16(dgv[t][4 * nn + 2], dgv[t][4 * nn + 3]);
          }
        }
      }
      tmem_st_wait();
#pragma unroll
      for (int q = 0; q < kMaxQ; ++q) {
        const float lo = quad_sum(da_lo[q]), hi = quad_sum(da_hi[q]);
        if ((f.lane & 3) == 0) { xv.da2[wg][q][f.row_lo] = lo; xv.da2[wg][q][f.row_hi] = hi; }
      }
    }
    // ---- early GEMMs ---------------------------------------------------------------------------------------------
    publish();
    if (leader) {
      if (mine(0)) gemm(kTY, 0, taddr(SB::DY), false, taddr(SB::VL), false, true, ksteps, 64);
      if (mine(1)) gemm(kTdV1, 0, taddr(SB::AMIX), true, taddr(SB::DY), true, false, 4, 64);
      if (mine(2)) gemm(kTdVL, 0, sF, true, taddr(SB::DY), true, false, 4, 64);
      for (int t = 0; t < 4; ++t)
        if (mine(3 + t)) gemm(kTdb, 16 * t, taddr(SB::X + t), true, smem_u32(bv_.a_bf), true, false, 4, 16);
      mma_commit(&sv_.bar);
    }
    for (int idx = tid; idx < kMaxQ * 64; idx += 256) bv_.da[idx >> 6][idx & 63] = xv.da2[0][idx >> 6][idx & 63] + xv.da2[1][idx >> 6][idx & 63];
    wait_mma();
    {
      // value gradients, v_scale column sums and db: each warpgroup handles its 32 columns
      float d1[16], dl[16], vb[16];
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const uint32_t c0 = 16 * (2 * wg + bb);
        tmem_ld_16x256b_x2(tlane + ttile<true>(kTdV1) + c0, d1 + 8 * bb);
        tmem_ld_16x256b_x2(tlane + ttile<true>(kTdVL) + c0, dl + 8 * bb);
        tmem_ld_16x256b_x2(tlane + ttile<true>(kTdb) + c0, vb + 8 * bb);
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // j = 8-column group within this warpgroup's half
        const int c = 32 * wg + 8 * j + f.cq;
        float2 vlo = make_float2(0.f, 0.f), vhi = vlo;
        const float* a = d1 + 4 * j;
        const float* b = dl + 4 * j;
        if (c < dk) {
          vlo = unpack_bf16(*reinterpret_cast<const uint32_t*>(qkv + in_lo + 2 * hd + c));
          vhi = unpack_bf16(*reinterpret_cast<const uint32_t*>(qkv + in_hi + 2 * hd + c));
          const float a0 = sv_.vs1[c], a1 = sv_.vs1[c + 1], b0 = sv_.vsL[c], b1 = sv_.vsL[c + 1];
          *reinterpret_cast<uint32_t*>(dqkv + in_lo + 2 * hd + c) = pack_bf16(a[0] * a0 + b[0] * b0, a[1] * a1 + b[1] * b1);
          *reinterpret_cast<uint32_t*>(dqkv + in_hi + 2 * hd + c) = pack_bf16(a[2] * a0 + b[2] * b0, a[3] * a1 + b[3] * b1);
        }
        // column sums over the 64 rows: this warp's 16 rows -> red[.][warp][col]
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float s1 = a[e] * (e ? vlo.y : vlo.x) + a[2 + e] * (e ? vhi.y : vhi.x);
          float sl = b[e] * (e ? vlo.y : vlo.x) + b[2 + e] * (e ? vhi.y : vhi.x);
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); sl += __shfl_xor_sync(0xffffffffu, sl, o); }
          if (f.lane < 4) { sv_.red[0][warp4][c + e] = s1; sv_.red[1][warp4][c + e] = sl; }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = c + (e & 1), t = col >> 4, within = col & 15;
          if ((within >> 2) == t) sv_.b[within][(e & 2) ? f.row_hi : f.row_lo] = vb[4 * j + e];   // db[q][token]
        }
      }
    }
    __syncthreads();
    if (tid < 32) {
      float s = 0.f;
      for (int d = tid; d < dk; d += 32) s = fmaf(sv_.vsL[d], sv_.red[1][0][d] + sv_.red[1][1][d] + sv_.red[1][2][d] + sv_.red[1][3][d], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tid == 0) p.dlogit_part[g] = (1.f - w) * s;
    }
    {
      float* ds = p.dscale_part + (size_t)g * 3 * V * dk + (size_t)2 * V * dk;
      for (int idx = tid; idx < V * dk; idx += 256) {
        const int k = idx / dk, d = idx % dk;
        float val = 0.f;
        if (k == 0) val = sv_.red[0][0][d] + sv_.red[0][1][d] + sv_.red[0][2][d] + sv_.red[0][3][d];
        if (k == V - 1) val += w * (sv_.red[1][0][d] + sv_.red[1][1][d] + sv_.red[1][2][d] + sv_.red[1][3][d]);
        ds[idx] = val;
      }
    }
    for (int idx = tid; idx < C * 64; idx += 256) {
      const int c = idx >> 6, tok = idx & 63;
      float sr = 0.f, sc = 0.f;
      for (int qq = 0; qq < kMaxQ; ++qq) {
        const int t = qq >> 2, k = qq & 3;
        if (k < r) {
          const int q = t * r + k;
          sr = fmaf(hw_row[q * C + c], bv_.da[qq][tok], sr);
          sc = fmaf(hw_col[q * C + c], sv_.b[qq][tok], sc);
        }
      }
      bv_.drho[c][tok] = sr * (1.f / 64.f);
      bv_.dkap[c][tok] = sc * (1.f / 64.f);
    }
    {
      const int nW = 4 * r * C, nP = nW + 4 * r;
      float* dh = p.dhead_part + (size_t)g * 2 * nP;
      for (int idx = tid; idx < 2 * nP; idx += 256) {
        const int half = idx / nP, rem = idx % nP;
        float (*dv)[64] = half ? sv_.b : bv_.da;
        float s = 0.f;
        if (rem < nW) {
          const int q = rem / C, c = rem % C, qq = 4 * (q / r) + (q % r);
          const float* ft;
          if (c < V) ft = half ? sv_.kap[c] : sv_.rho[c];
          else if (c < 2 * V) ft = half ? sv_.rho[c - V] : sv_.kap[c - V];
          else ft = half ? sv_.kap[c] : sv_.rho[c];
          // four independent partial sums over float4 loads: the serial 64-step fma chain was shared-memory latency bound
          float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4* dp = reinterpret_cast<const float4*>(dv[qq]);
          const float4* fp = reinterpret_cast<const float4*>(ft);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 d4 = dp[i], f4 = fp[i];
            a4.x = fmaf(d4.x, f4.x, a4.x); a4.y = fmaf(d4.y, f4.y, a4.y); a4.z = fmaf(d4.z, f4.z, a4.z); a4.w = fmaf(d4.w, f4.w, a4.w);
          }
          s = (a4.x + a4.y) + (a4.z + a4.w);
        } else {
          const int q = rem - nW, qq = 4 * (q / r) + (q % r);
          float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4* dp = reinterpret_cast<const float4*>(dv[qq]);
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float4 d4 = dp[i]; a4.x += d4.x; a4.y += d4.y; a4.z += d4.z; a4.w += d4.w; }
          s = (a4.x + a4.y) + (a4.z + a4.w);
        }
        dh[idx] = s;
      }
    }
    __syncthreads();
    // ---- chain seeds (wg0: X_F, wg1: X_R) and feature terms of dS_k (views split) -------------------------------------
    {
      float x[32], den[32];
      if (wg == 0) {
        ld_tile(kTY, x);
        ld_tile(kTF, den);
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = f.col(n) + (e & 1), row = (e & 2) ? f.row_hi : f.row_lo;
            x[4 * n + e] += (bv_.drho[2 * V][row] + bv_.dkap[2 * V][col]) * fast_rcp(den[4 * n + e] + p.eps);
          }
        frag_store_bf16(tile(SB::X + 0), f, x);
      } else {
        ld_tile(kTR, den);
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = f.col(n) + (e & 1), row = (e & 2) ? f.row_hi : f.row_lo;
            x[4 * n + e] = (bv_.drho[2 * V + 1][row] + bv_.dkap[2 * V + 1][col]) * fast_rcp(den[4 * n + e] + p.eps);
          }
        frag_store_bf16(tile(SB::X + 2), f, x);
      }
      for (int k = wg; k < V; k += 2) {
        ld_tile(kTS + k, x);
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = f.col(n) + (e & 1), row = (e & 2) ? f.row_hi : f.row_lo;
            x[4 * n + e] += bv_.drho[k][row] + bv_.dkap[k][col] + bv_.drho[V + k][col] + bv_.dkap[V + k][row];
          }
        st_tile(kTS + k, x);
      }
    }
    // ---- chain sweep: wg0 = forward chain (accumulates into tile k), wg1 = reverse chain (writes tile SR(k)) ----------
    {
      int xf = SB::X + 0, xr = SB::X + 2;
      for (int s = 0; s <= V - 2; ++s) {
        const int kF = V - 1 - s, kR = s;
        publish();
        if (leader) {
          const uint32_t pPrev = (kF - 1 == 0) ? taddr(SB::A + 0) : taddr(SB::P(kF - 1));
          const uint32_t rNext = (kR + 1 == V - 1) ? taddr(SB::A + V - 1) : taddr(SB::R(V - 1 - (kR + 1)));
          if (mine(0)) gemm(kTAF, 0, pPrev, true, taddr(xf), true, false, 4, 64);
          if (mine(1)) gemm(kTXF, 0, taddr(xf), false, taddr(SB::A + kF), false, false, 4, 64);
          if (mine(4)) gemm(kTAR, 0, rNext, true, taddr(xr), true, false, 4, 64);
          if (mine(5)) gemm(kTXR, 0, taddr(xr), false, taddr(SB::A + kR), false, false, 4, 64);
          mma_commit(&sv_.bar);
        }
        wait_mma();
        const int xf_n = (xf == SB::X) ? SB::X + 1 : SB::X, xr_n = (xr == SB::X + 2) ? SB::X + 3 : SB::X + 2;
        float x[32], pk[32];
        if (wg == 0) {
          float acc[32];
          ld_tile(kTAF, x);
          frag_load_bf16(tile(SB::A + kF), f, pk);
          frag_softmax_bwd(x, pk);
          ld_tile(kTS + kF, acc);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] += x[i];
          st_tile(kTS + kF, acc);
          if (s < V - 2) { ld_tile(kTXF, x); frag_store_bf16(tile(xf_n), f, x); }
        } else {
          ld_tile(kTAR, x);
          frag_load_bf16(tile(SB::A + kR), f, pk);
          frag_softmax_bwd(x, pk);
          st_tile(tileSR(kR), x);
          if (s < V - 2) { ld_tile(kTXR, x); frag_store_bf16(tile(xr_n), f, x); }
        }
        if (s < V - 2) { xf = xf_n; xr = xr_n; }
      }
      float x[32], pk[32];
      if (wg == 0) {   // dA_0 += X_F
        float acc[32];
        ld_tile(kTXF, x);
        frag_load_bf16(tile(SB::A + 0), f, pk);
        frag_softmax_bwd(x, pk);
        ld_tile(kTS + 0, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += x[i];
        st_tile(kTS + 0, acc);
      } else {         // dA_{V-1} = X_R
        ld_tile(kTXR, x);
        frag_load_bf16(tile(SB::A + V - 1), f, pk);
        frag_softmax_bwd(x, pk);
        st_tile(tileSR(V - 1), x);
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- dS_k = forward part + reverse part -> bf16 operand tiles (views split); reload unscaled Q ---------------------
    auto ds_slot = [&](int k) { return k < 4 ? 1 + k : SB::X + 0; };
    for (int k = wg; k < V; k += 2) {
      float x[32], y2[32];
      ld_tile(kTS + k, x);
      ld_tile(tileSR(k), y2);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] += y2[i];
      frag_store_bf16(tile(ds_slot(k)), f, x);
    }
    for (int idx = tid; idx < 64 * 8; idx += 256) {
      const int rr = idx & 63, ch = idx >> 6;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (ch * 8 < dk) q = *reinterpret_cast<const uint4*>(qkv + (((size_t)pb * 64 + rr) * 3) * hd + (size_t)ph * dk + ch * 8);
      *reinterpret_cast<uint4*>(tile(SB::X + 1) + ch * 1024 + rr * 16) = q;
    }
    publish();
    if (leader) {
      for (int k = 0; k < V; ++k) {
        if (mine(2 * k)) gemm(tileT2(k), 0, taddr(ds_slot(k)), false, taddr(SB::K), true, false, 4, 64);
        if (mine(2 * k + 1)) gemm(tileU2(k), 0, taddr(ds_slot(k)), true, taddr(SB::X + 1), true, false, 4, 64);
      }
      mma_commit(&sv_.bar);
    }
    wait_mma();
    {
      float accv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) accv[i] = 0.f;
      if (wg == 0) {   // dQ = sum_k T_k (.) c_k ; Z_k = colsum(T_k (.) Q)
        float qf[32];
        frag_load_bf16(tile(SB::X + 1), f, qf);
        for (int k = 0; k < V; ++k) {
          float t[32], z[32];
          ld_tile(tileT2(k), t);
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float c0 = sv_.cvec[k][f.col(n)], c1 = sv_.cvec[k][f.col(n) + 1];
            accv[4 * n] = fmaf(t[4 * n], c0, accv[4 * n]); accv[4 * n + 1] = fmaf(t[4 * n + 1], c1, accv[4 * n + 1]);
            accv[4 * n + 2] = fmaf(t[4 * n + 2], c0, accv[4 * n + 2]); accv[4 * n + 3] = fmaf(t[4 * n + 3], c1, accv[4 * n + 3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) z[4 * n + e] = t[4 * n + e] * qf[4 * n + e];
          }
          colsum_to(sv_.red[k], f, z);
        }
      } else {         // dK = sum_k U_k (.) c_k
        for (int k = 0; k < V; ++k) {
          float t[32];
          ld_tile(tileU2(k), t);
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float c0 = sv_.cvec[k][f.col(n)], c1 = sv_.cvec[k][f.col(n) + 1];
            accv[4 * n] = fmaf(t[4 * n], c0, accv[4 * n]); accv[4 * n + 1] = fmaf(t[4 * n + 1], c1, accv[4 * n + 1]);
            accv[4 * n + 2] = fmaf(t[4 * n + 2], c0, accv[4 * n + 2]); accv[4 * n + 3] = fmaf(t[4 * n + 3], c1, accv[4 * n + 3]);
          }
        }
      }
      const size_t sel = wg ? hd : 0;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int c = f.col(n);
        if (c < dk) {
          *reinterpret_cast<uint32_t*>(dqkv + in_lo + sel + c) = pack_bf16(accv[4 * n], accv[4 * n + 1]);
          *reinterpret_cast<uint32_t*>(dqkv + in_hi + sel + c) = pack_bf16(accv[4 * n + 2], accv[4 * n + 3]);
        }
      }
    }
    __syncthreads();
    {
      float* ds = p.dscale_part + (size_t)g * 3 * V * dk;
      for (int idx = tid; idx < V * dk; idx += 256) {
        const int k = idx / dk, d = idx % dk;
        const float z = sscale * (sv_.red[k][0][d] + sv_.red[k][1][d] + sv_.red[k][2][d] + sv_.red[k][3][d]);
        const size_t pi = ((size_t)k * H + ph) * dk + d;
        ds[idx] = p.k_scale[pi] * z;
        ds[(size_t)V * dk + idx] = p.q_scale[pi] * z;
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tbase);
}

}  // namespace ewtc
}  // namespace mop
