// Fused tcgen05/TMEM Edgewise backward for the config-1/2 hot shape (N = 64 tokens, dk <= 64, V <= 5 shared-projection views,
// low-rank gates r <= 4), third generation: ONE CTA of 512 threads (sixteen warps = 4 TMEM sub-partitions x 4 column blocks)
// per SM, one (batch, head) problem at a time.
//
// What changed against ewtc::edgewise_bwd2_kernel (256 threads, 2.9 % of the tensor peak, issue slots 27 % busy, 1.9 stall
// cycles per issue waiting for instructions):
//   * every small per-problem vector the forward already has - softmax row statistics of the V score maps and of the mixed map,
//     row/column feature means, the low-rank gate factors a, b - arrives through `aux` (17 KB per problem, one bulk copy)
//     instead of being recomputed: no row-max / row-sum exchanges, no log-feature passes, no gate-factor phase;
//   * the raw Q, K, V, dY tiles arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) and are MMA operands as they land; the
//     loads of V, dY and aux for the NEXT problem are issued as soon as this problem has consumed them;
//   * the gate pre-activations z_t = sum_k a_tk[i] b_tk[j] are four K=16 MMAs on bf16 hi+lo splits (error 2^-17) instead of
//     16 FMAs + 16 shared loads per map element in two passes;
//   * the mixed map is evaluated ONCE per element: pass 1 leaves the ten coefficient maps that multiply D = A (dA - delta) in
//     TMEM (C_k over S_k, W_t over z_t), pass 2 is ten multiplies; the row term delta needs one exchange;
//   * chain backward: dA_k accumulates its forward-chain and reverse-chain contribution in ONE TMEM tile by MMA accumulation,
//     the softmax backward of view k runs once, fused with the direct and the feature terms of dS_k and the bf16 conversion;
//   * element loops are rolled over 8-column octets / views / gates (x1, x2 TMEM accesses), so that a warp executes ~3 k
//     instructions per problem out of ~50 KB of code and four warps per scheduler share every instruction-cache line;
//   * d chain_value_logit = (1-w) sum F (dY (w V_V)^T) is formed from the fp32 chain product in TMEM, not from its bf16 tile.
//
// Math: SURVEY.md appendix A / D.1 (reference attention_variants.py:500-562, :319-331); executable specification
// oracle/edgewise_manual.py.  Tile layouts, descriptors and the M=64 two-bank TMEM map are those of edgewise_tc.cuh.
#pragma once
#include "edgewise_tc.cuh"

namespace mop {
namespace ew64 {

using namespace tc;
using ewtc::fast_exp2;
using ewtc::fast_log2;
using ewtc::fast_rcp;
using ewtc::fast_sigmoid;
using ewtc::kAuxA;
using ewtc::kAuxB;
using ewtc::kAuxFloats;
using ewtc::kAuxKap;
using ewtc::kAuxRho;
using ewtc::kAuxStats;
using ewtc::kLn2;
using ewtc::kLog2e;
using ewtc::kMaxC;
using ewtc::kMaxQ;
using ewtc::kMaxV;
using ewtc::kTile;
using ewtc::publish;
using ewtc::scale_chunk;
using ewtc::tmem_ld_16x256b_x2;
using ewtc::tmem_st_16x256b_x2;

constexpr int kThreads = 512;

// ---- TMEM tiles (16 tiles of 64 columns x 16 lanes per sub-partition, two banks; ewtc::ttile<true>) -------------------------
constexpr int tS = 0;      // S_k -> C_k -> C_k D (direct part of dS_k)                 0..4      final: T_k = dS_k K
constexpr int tF = 5;      // chain product F (fp32)
constexpr int tR = 6;      // chain product R
constexpr int tT1 = 7;     // dA = dY V_1^T -> T1 = A dA                                 then dV1 = A^T dY
constexpr int tZ = 8;      // z_t -> W_t (gate t)                                        8..11   then dVL (8), db (9)
constexpr int tT2 = 12;    // A (fp32)
constexpr int tW4 = 13;    // g_chain / (F + eps) -> Hf = D g_chain / (F + eps)          sweep: X_F' accumulator
constexpr int tG = 14;     // G = dY (w V_V)^T                                           sweep: X_R' accumulator
constexpr int tdV1 = 7, tdVL = 8, tdb = 9;
constexpr int tdA = 8;     // dA_k accumulators of the chain sweep                       8..12   final: U_k = dS_k^T Q
constexpr int tXF = 13, tXR = 14;

// operand kinds of the 64-row tiles
enum : int { CM_K = 0, CM_MN = 1, SW_K = 2, SW_MN = 3 };   // chunk-major | 128-byte swizzled (TMA), used K-major | MN-major

struct __align__(1024) Smem {
  unsigned char Kr[kTile], Qr[kTile], Vr[kTile], DYr[kTile];   // raw tiles as TMA lands them (128-byte swizzle)
  unsigned char A[kMaxV][kTile];       // Q (.) c_k  ->  A_k  ->  dS_k                (chunk-major, as every tile below)
  unsigned char P[kMaxV - 1][kTile];   // P(s) = A_0..A_s, s = 1..V-1 at P[s-1] (P(V-1) = F); P[0], P[1] first hold V_1, w V_V
  unsigned char R[kMaxV - 2][kTile];   // R(s) = A_{V-1}..A_{V-1-s}, s = 1..V-2 at R[s-1]
  unsigned char AMIX[kTile];
  unsigned char X[4][kTile];           // dG_t (gate pre-activation gradients)  ->  chain-sweep X ping-pong (F: 0/1, R: 2/3)
  unsigned char ahl[4][64 * 16 * 2];   // per gate: [token][a_hi(4) a_lo(4) a_hi(4) 0(4)] bf16
  unsigned char bhl[4][64 * 16 * 2];   // per gate: [token][b_hi(4) b_hi(4) b_lo(4) 0(4)]
  float aux[kAuxFloats];               // stats | rho | kap | a | b  (one bulk copy per problem)
  float da[kMaxQ][64], db[kMaxQ][64];
  float rt[kMaxV][64], ct[kMaxV][64];  // feature-mean gradients as they enter dS_k: row term drho_k + dkap_{V+k}, column term dkap_k + drho_{V+k}
  float sd[4][64];                     // chain-seed terms: drho_{2V}, dkap_{2V}, drho_{2V+1}, dkap_{2V+1}
  float cvec[kMaxV][64], vs1[64], vsL[64];
  float part[4][64];                   // row-term exchange between the four column blocks
  float red[kMaxV][4][64];             // column-sum partials per sub-partition / row dots of the final pass (two buffers)
  float hw[2][kMaxQ * kMaxC + kMaxQ];  // gate-head weights + biases (row / column projection), staged once per CTA
  float wsum[16];
  float acc_head[2 * (kMaxQ * kMaxC + kMaxQ)];   // gate-head parameter gradients summed over this CTA's problems (one owner thread per entry)
  uint64_t bar_mma, bar_in;
  uint32_t tmem_slot;
};
static_assert(sizeof(Smem) + 1024 <= 232448, "backward shared memory over the 227 KB limit");

__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x1(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}

__device__ __forceinline__ uint64_t op_desc(int kind, uint32_t tile, int kstep) {
  switch (kind) {
    case CM_K: return desc_kmajor(tile, 64, 16 * kstep);
    case CM_MN: return desc_mnmajor(tile, 64, 16 * kstep);
    case SW_K: return desc_k_sw(tile, 16 * kstep);
    default: return desc_mn_sw(tile, 16 * kstep);
  }
}

static __global__ void __launch_bounds__(kThreads, 1)
edgewise_bwd3_kernel(MopEdgewiseParams p, const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDY) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int sp = wid & 3, cb = wid >> 2;                         // TMEM sub-partition (16 rows of every tile), column block
  const int row_lo = 16 * sp + (lane >> 2), row_hi = row_lo + 8, cq = 2 * (lane & 3);
  const int c0 = 16 * cb;                                        // this thread's columns: c0 + 8n + cq + {0,1}, n = 0, 1
  const int V = p.V, r = p.gate_rank, C = 2 * V + 2, dk = p.dk, H = p.H;
  const int ksteps = (dk + 15) >> 4;

  if (wid == 0) tmem_alloc<512>(&sm.tmem_slot);
  constexpr int kIssuers = 16;   // lane 0 of every warp issues (and commits) its share of each MMA batch
  if (tid == 0) { mbar_init(&sm.bar_mma, kIssuers); mbar_init(&sm.bar_in, 2); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const bool leader = lane == 0;
  auto mine = [&](int idx) { return (idx & (kIssuers - 1)) == wid; };
  const uint32_t tbase = sm.tmem_slot;
  const uint32_t tlane = tbase + ((uint32_t)(32 * sp) << 16);
  auto tcol = [&](int t) -> uint32_t { return tlane + ewtc::ttile<true>(t) + (uint32_t)c0; };   // my 16 columns of tile t
  uint32_t ph_mma = 0, ph_in = 0;
  const float w = 1.f / (1.f + __expf(-p.chain_value_logit[0]));
  const float bn = p.beta_not / (float)max(1, V - 1);
  const float sscale = rsqrtf((float)dk);
  const float eps = p.eps;

  auto sa = [&](const void* ptr) -> uint32_t { return smem_u32(ptr); };
  // D[tile dt, columns dcol ..] (+)= op(A) op(B) over `ks` K-steps of 16; 64-row operand tiles
  auto gemm = [&](int dt, uint32_t dcol, uint32_t a_tile, int a_kind, uint32_t b_tile, int b_kind, bool acc, int ks, uint32_t n) {
    const uint32_t id = idesc_bf16(64, n, (a_kind & 1) ? 1u : 0u, (b_kind & 1) ? 1u : 0u);
    for (int k = 0; k < ks; ++k)
      mma_ss(tbase + ewtc::ttile<true>(dt) + dcol, op_desc(a_kind, a_tile, k), op_desc(b_kind, b_tile, k), id, (acc || k > 0) ? 1u : 0u);
  };
  auto wait_mma = [&]() { mbar_wait(&sm.bar_mma, ph_mma); ph_mma ^= 1; tc_fence_after(); };
  // my 8 values of a 64x64 map as bf16 into / from a chunk-major tile: element (row, 16 cblk + 8n + cq) sits at
  // (2 cblk + n) * 1024 + row * 16 + 2 cq
  const uint32_t o_lo = (uint32_t)(row_lo * 16 + 2 * cq), o_hi = o_lo + 128;
  auto put_bf16 = [&](unsigned char* tile, int cblk, const float* v) {
    unsigned char* t = tile + cblk * 2048;
    *reinterpret_cast<uint32_t*>(t + o_lo) = pack_bf16(v[0], v[1]);
    *reinterpret_cast<uint32_t*>(t + o_hi) = pack_bf16(v[2], v[3]);
    *reinterpret_cast<uint32_t*>(t + 1024 + o_lo) = pack_bf16(v[4], v[5]);
    *reinterpret_cast<uint32_t*>(t + 1024 + o_hi) = pack_bf16(v[6], v[7]);
  };
  auto get_bf16 = [&](const unsigned char* tile, int cblk, float* v) {
    const unsigned char* t = tile + cblk * 2048;
    const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(t + o_lo)), b = unpack_bf16(*reinterpret_cast<const uint32_t*>(t + o_hi));
    const float2 c = unpack_bf16(*reinterpret_cast<const uint32_t*>(t + 1024 + o_lo)), d = unpack_bf16(*reinterpret_cast<const uint32_t*>(t + 1024 + o_hi));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  };

  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p.qkv);
  __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(p.dqkv);
  const size_t hd = (size_t)H * dk;
  const int G = p.B * H;
  // input stage: {V, dY, aux} and {Q, K} are two transactions on one barrier (count 2) so that the first group can be
  // re-issued for the next problem as soon as this one has consumed it
  auto load_vdy = [&](int g) {
    mbar_expect_tx(&sm.bar_in, 2 * kTile + kAuxFloats * 4);
    tma_load_tile_sw(sm.Vr, &tmV, 0, g % H, g / H, &sm.bar_in);
    tma_load_tile_sw(sm.DYr, &tmDY, 0, g % H, g / H, &sm.bar_in);
    bulk_g2s(sm.aux, p.aux + (size_t)g * kAuxFloats, kAuxFloats * 4, &sm.bar_in);
  };
  auto load_qk = [&](int g) {
    mbar_expect_tx(&sm.bar_in, 2 * kTile);
    tma_load_tile_sw(sm.Qr, &tmQ, 0, g % H, g / H, &sm.bar_in);
    tma_load_tile_sw(sm.Kr, &tmK, 0, g % H, g / H, &sm.bar_in);
  };
  if (tid == 0 && (int)blockIdx.x < G) { load_vdy(blockIdx.x); load_qk(blockIdx.x); }
  {   // head weights: read through L2 once per CTA
    const int nW = 4 * r * C;
    for (int idx = tid; idx < 2 * (nW + 4 * r); idx += kThreads) {
      const int half = idx / (nW + 4 * r), rem = idx % (nW + 4 * r);
      sm.hw[half][rem] = rem < nW ? (half ? p.col_w : p.row_w)[rem] : (half ? p.col_b : p.row_b)[rem - nW];
    }
  }
  const float* hw_row = sm.hw[0];
  const float* hw_col = sm.hw[1];
  float (*rho)[64] = reinterpret_cast<float (*)[64]>(sm.aux + kAuxRho);
  float (*kap)[64] = reinterpret_cast<float (*)[64]>(sm.aux + kAuxKap);
  float (*bfac)[64] = reinterpret_cast<float (*)[64]>(sm.aux + kAuxB);
  const float* stats = sm.aux + kAuxStats;

  // Parameter-gradient partials are summed over the problems of this CTA (fixed order: deterministic) and written once at the
  // end, one row per CTA: 7x fewer bytes and a 7x shorter host-side reduction than one row per (b,h) problem.  Only the scale
  // gradients depend on the head h of the problem, so they are kept per head: thread tid < V*dk owns element tid of
  // (d q_scale, d k_scale, d v_scale); rows are indexed [cta][head].
  for (int idx = tid; idx < 2 * (kMaxQ * kMaxC + kMaxQ); idx += kThreads) sm.acc_head[idx] = 0.f;
  float acc_sq = 0.f, acc_sk = 0.f, acc_sv = 0.f;   // the grid is a multiple of H: every problem of this CTA has the same head
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int pb = g / H, ph = g % H;
    const size_t in_lo = (((size_t)pb * 64 + row_lo) * 3) * hd + (size_t)ph * dk;   // q row; +hd: k; +2hd: v
    const size_t in_hi = (((size_t)pb * 64 + row_hi) * 3) * hd + (size_t)ph * dk;
    // =========================================================================================================================
    // phase 0: per-head scale vectors, operand tiles
    // =========================================================================================================================
    for (int idx = tid; idx < V * 64; idx += kThreads) {
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
    {
      const int rr = tid & 63, ch = tid >> 6;   // one 16-byte chunk (8 columns) of one row per thread
      const uint32_t src = sw128_off(rr, 8 * ch), dst = ch * 1024 + rr * 16;
      const uint4 q = *reinterpret_cast<const uint4*>(sm.Qr + src);
      const uint4 v = *reinterpret_cast<const uint4*>(sm.Vr + src);
      for (int i = 0; i < V; ++i) *reinterpret_cast<uint4*>(sm.A[i] + dst) = scale_chunk(q, &sm.cvec[i][ch * 8]);
      *reinterpret_cast<uint4*>(sm.P[0] + dst) = scale_chunk(v, &sm.vs1[ch * 8]);
      *reinterpret_cast<uint4*>(sm.P[1] + dst) = scale_chunk(v, &sm.vsL[ch * 8]);
    }
    {
      // gate-factor operand tiles: a, b split into bf16 hi + lo so that a_hi b_hi + a_lo b_hi + a_hi b_lo is exact to 2^-17
      const int tok = tid & 63, t = (tid >> 6) & 3, which = tid >> 8;
      const float* src = sm.aux + (which ? kAuxB : kAuxA);
      uint32_t hi2[2], lo2[2];
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const float x0 = src[(4 * t + 2 * k2) * 64 + tok], x1 = src[(4 * t + 2 * k2 + 1) * 64 + tok];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
        hi2[k2] = pack_bf16(__bfloat162float(h0), __bfloat162float(h1));
        lo2[k2] = pack_bf16(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
      }
      unsigned char* tile = which ? sm.bhl[t] : sm.ahl[t];
      const uint4 first = which ? make_uint4(hi2[0], hi2[1], hi2[0], hi2[1]) : make_uint4(hi2[0], hi2[1], lo2[0], lo2[1]);
      const uint4 second = which ? make_uint4(lo2[0], lo2[1], 0u, 0u) : make_uint4(hi2[0], hi2[1], 0u, 0u);
      *reinterpret_cast<uint4*>(tile + tok * 16) = first;           // slots 0..7
      *reinterpret_cast<uint4*>(tile + 1024 + tok * 16) = second;   // slots 8..15
    }
    publish();
    // =========================================================================================================================
    // batch 1: S_k = (Q c_k) K^T, dA = dY V_1^T, G = dY (w V_V)^T, z_t = a_t b_t^T
    // =========================================================================================================================
    if (leader) {
      for (int i = 0; i < V; ++i)
        if (mine(i)) gemm(tS + i, 0, sa(sm.A[i]), CM_K, sa(sm.Kr), SW_K, false, ksteps, 64);
      if (mine(5)) gemm(tT1, 0, sa(sm.DYr), SW_K, sa(sm.P[0]), CM_K, false, ksteps, 64);
      if (mine(6)) gemm(tG, 0, sa(sm.DYr), SW_K, sa(sm.P[1]), CM_K, false, ksteps, 64);
      for (int t = 0; t < 4; ++t)
        if (mine(7 + t)) gemm(tZ + t, 0, sa(sm.ahl[t]), CM_K, sa(sm.bhl[t]), CM_K, false, 1, 64);
      mma_commit(&sm.bar_mma);
    }
    wait_mma();
    // per-view softmax maps from the saved row statistics: A_k = exp2(S_k log2e - m_k) / l_k   (overwrites Q (.) c_k)
    for (int k = 0; k < V; ++k) {
      float v[8];
      tmem_ld_16x256b_x2(tcol(tS + k), v);
      const float2 slo = *reinterpret_cast<const float2*>(stats + (k * 64 + row_lo) * 2);
      const float2 shi = *reinterpret_cast<const float2*>(stats + (k * 64 + row_hi) * 2);
      tmem_ld_wait();
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        v[4 * n + 0] = fast_exp2(fmaf(v[4 * n + 0], kLog2e, -slo.x)) * slo.y;
        v[4 * n + 1] = fast_exp2(fmaf(v[4 * n + 1], kLog2e, -slo.x)) * slo.y;
        v[4 * n + 2] = fast_exp2(fmaf(v[4 * n + 2], kLog2e, -shi.x)) * shi.y;
        v[4 * n + 3] = fast_exp2(fmaf(v[4 * n + 3], kLog2e, -shi.x)) * shi.y;
      }
      put_bf16(sm.A[k], cb, v);
    }
    // chain products F = A_0..A_{V-1}, R = A_{V-1}..A_0, prefixes / suffixes kept as bf16 tiles for the sweep
    uint32_t sF;
    {
      uint32_t xf = sa(sm.A[0]), xr = sa(sm.A[V - 1]);
      for (int s = 1; s < V; ++s) {
        publish();
        if (leader) {
          if (mine(0)) gemm(tF, 0, xf, CM_K, sa(sm.A[s]), CM_MN, false, 4, 64);
          if (mine(1)) gemm(tR, 0, xr, CM_K, sa(sm.A[V - 1 - s]), CM_MN, false, 4, 64);
          mma_commit(&sm.bar_mma);
        }
        wait_mma();
        float v[8];
        tmem_ld_16x256b_x2(tcol(tF), v);
        tmem_ld_wait();
        put_bf16(sm.P[s - 1], cb, v);
        xf = sa(sm.P[s - 1]);
        if (s < V - 1) {
          tmem_ld_16x256b_x2(tcol(tR), v);
          tmem_ld_wait();
          put_bf16(sm.R[s - 1], cb, v);
          xr = sa(sm.R[s - 1]);
        }
      }
      sF = xf;
    }
    // =========================================================================================================================
    // pass 1 (one evaluation of the mix per element): A, T1 = A dA, and the coefficient maps that multiply D in pass 2
    //   W_t = d z_t / d(.) factors of the four gates, W4 = g_chain / (F + eps), C_k = d Smix / d S_k (direct part)
    // =========================================================================================================================
    {
      const float2 mlo = *reinterpret_cast<const float2*>(stats + (V * 64 + row_lo) * 2);
      const float2 mhi = *reinterpret_cast<const float2*>(stats + (V * 64 + row_hi) * 2);
      float dl_lo = 0.f, dl_hi = 0.f;
#pragma unroll 1
      for (int n = 0; n < 2; ++n) {
        const uint32_t off = 8 * n;
        float s[kMaxV][4], fv[4], z[4][4], da4[4], a4[4];
#pragma unroll
        for (int i = 0; i < kMaxV; ++i)
          if (i < V) tmem_ld_x1(tcol(tS + i) + off, s[i]);
        tmem_ld_x1(tcol(tF) + off, fv);
#pragma unroll
        for (int t = 0; t < 4; ++t) tmem_ld_x1(tcol(tZ + t) + off, z[t]);
        tmem_ld_x1(tcol(tT1) + off, da4);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool hi = (e & 2) != 0;
          const float s0 = s[0][e];
          float sum = s0, mx = s0;
#pragma unroll
          for (int i = 1; i < kMaxV; ++i)
            if (i < V) { sum += s[i][e]; mx = fmaxf(mx, s[i][e]); }
          const float nm = -mx * kLog2e;
          float ex[kMaxV], se = 0.f;
#pragma unroll
          for (int i = 0; i < kMaxV; ++i)
            if (i < V) { ex[i] = fast_exp2(fmaf(s[i][e], kLog2e, nm)); se += ex[i]; }
          const float inv_se = fast_rcp(se);
          const float lse = fmaf(kLn2, fast_log2(se), mx);
          const float U = sum - s0, O = lse - s0;
          const float fe = fv[e] + eps, lf = kLn2 * fast_log2(fe), rfe = fast_rcp(fe);
          const float g0 = fast_sigmoid(z[0][e]), g1 = fast_sigmoid(z[1][e]), g2 = fast_sigmoid(z[2][e]), g3 = fast_sigmoid(z[3][e]);
          const float gu = fmaf(-bn, g2, g0);
          const float am = fmaf(g3, lf, fmaf(g1, O, fmaf(gu, U, s0)));
          const float A = fast_exp2(fmaf(am, kLog2e, hi ? -mhi.x : -mlo.x)) * (hi ? mhi.y : mlo.y);
          a4[e] = A;
          const float t1 = A * da4[e];
          da4[e] = t1;
          if (hi) dl_hi += t1; else dl_lo += t1;
          z[0][e] = U * g0 * (1.f - g0);
          z[1][e] = O * g1 * (1.f - g1);
          z[2][e] = -bn * U * g2 * (1.f - g2);
          z[3][e] = lf * g3 * (1.f - g3);
          fv[e] = g3 * rfe;
          const float e1 = g1 * inv_se;
#pragma unroll
          for (int i = 0; i < kMaxV; ++i)
            if (i < V) s[i][e] = (i == 0) ? fmaf(e1, ex[0], 1.f - g1) : fmaf(e1, ex[i], gu);
        }
#pragma unroll
        for (int i = 0; i < kMaxV; ++i)
          if (i < V) tmem_st_x1(tcol(tS + i) + off, s[i]);
#pragma unroll
        for (int t = 0; t < 4; ++t) tmem_st_x1(tcol(tZ + t) + off, z[t]);
        tmem_st_x1(tcol(tW4) + off, fv);
        tmem_st_x1(tcol(tT1) + off, da4);
        tmem_st_x1(tcol(tT2) + off, a4);
        *reinterpret_cast<uint32_t*>(sm.AMIX + (2 * cb + n) * 1024 + o_lo) = pack_bf16(a4[0], a4[1]);
        *reinterpret_cast<uint32_t*>(sm.AMIX + (2 * cb + n) * 1024 + o_hi) = pack_bf16(a4[2], a4[3]);
      }
      dl_lo = quad_sum(dl_lo);
      dl_hi = quad_sum(dl_hi);
      if ((lane & 3) == 0) { sm.part[cb][row_lo] = dl_lo; sm.part[cb][row_hi] = dl_hi; }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // =========================================================================================================================
    // pass 2: D = T1 - delta T2; direct part of dS_k = C_k D (in place), Hf = W4 D (in place), dG_t = W_t D (bf16 tiles),
    //         da[q][i] = sum_j dG_t[i,j] b_q[j] in fp32 (rows of D sum to zero: this sum cancels heavily)
    // =========================================================================================================================
    {
      const float dlo = (sm.part[0][row_lo] + sm.part[1][row_lo]) + (sm.part[2][row_lo] + sm.part[3][row_lo]);
      const float dhi = (sm.part[0][row_hi] + sm.part[1][row_hi]) + (sm.part[2][row_hi] + sm.part[3][row_hi]);
      {
        float d[8], t2[8];
        tmem_ld_16x256b_x2(tcol(tT1), d);
        tmem_ld_16x256b_x2(tcol(tT2), t2);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = fmaf((e & 2) ? -dhi : -dlo, t2[e], d[e]);
        for (int k = 0; k <= V; ++k) {   // k = V: the chain-gate term
          float c[8];
          const uint32_t ta = tcol(k < V ? tS + k : tW4);
          tmem_ld_16x256b_x2(ta, c);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) c[e] *= d[e];
          tmem_st_16x256b_x2(ta, c);
        }
      }
      {
        // warpgroup `cb` owns gate t = cb over ALL 64 columns (16 at a time): complete row sums, one writer per dG tile
        const int t = cb;
        float acc[4][4];   // [rank][row_lo even/odd column, row_hi even/odd column]
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          float gq[8], t2[8], wv[8];
          const uint32_t co = tlane + 16 * ch;
          tmem_ld_16x256b_x2(co + ewtc::ttile<true>(tT1), gq);
          tmem_ld_16x256b_x2(co + ewtc::ttile<true>(tT2), t2);
          tmem_ld_16x256b_x2(co + ewtc::ttile<true>(tZ + t), wv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) gq[e] = fmaf((e & 2) ? -dhi : -dlo, t2[e], gq[e]) * wv[e];
          put_bf16(sm.X[t], ch, gq);
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const int c = 16 * ch + 8 * n + cq;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 bb = *reinterpret_cast<const float2*>(&bfac[4 * t + k][c]);
              acc[k][0] = fmaf(gq[4 * n + 0], bb.x, acc[k][0]);
              acc[k][1] = fmaf(gq[4 * n + 1], bb.y, acc[k][1]);
              acc[k][2] = fmaf(gq[4 * n + 2], bb.x, acc[k][2]);
              acc[k][3] = fmaf(gq[4 * n + 3], bb.y, acc[k][3]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = quad_sum(acc[k][0] + acc[k][1]), hi = quad_sum(acc[k][2] + acc[k][3]);
          if ((lane & 3) == 0) { sm.da[4 * t + k][row_lo] = lo; sm.da[4 * t + k][row_hi] = hi; }
        }
      }
      tmem_st_wait();
    }
    // =========================================================================================================================
    // batch 2: dV1 = A^T dY, dVL = F^T dY, db_t = dG_t^T [a_hi a_lo ..]
    // =========================================================================================================================
    publish();
    if (leader) {
      if (mine(0)) gemm(tdV1, 0, sa(sm.AMIX), CM_MN, sa(sm.DYr), SW_MN, false, 4, 64);
      if (mine(1)) gemm(tdVL, 0, sF, CM_MN, sa(sm.DYr), SW_MN, false, 4, 64);
      for (int t = 0; t < 4; ++t)
        if (mine(2 + t)) gemm(tdb, 16 * t, sa(sm.X[t]), CM_MN, sa(sm.ahl[t]), CM_MN, false, 4, 16);
      mma_commit(&sm.bar_mma);
    }
    wait_mma();
    {
      // value gradients, v_scale column sums (my 16 columns), db (column block t = cb holds the 16 operand slots of gate t)
      float d1[8], dl[8], vb[8];
      tmem_ld_16x256b_x2(tcol(tdV1), d1);
      tmem_ld_16x256b_x2(tcol(tdVL), dl);
      tmem_ld_16x256b_x2(tcol(tdb), vb);
      tmem_ld_wait();
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int c = c0 + 8 * n + cq;
        float2 vlo = make_float2(0.f, 0.f), vhi = vlo;
        const float* a = d1 + 4 * n;
        const float* b = dl + 4 * n;
        if (c < dk) {
          vlo = unpack_bf16(*reinterpret_cast<const uint32_t*>(sm.Vr + sw128_off(row_lo, c)));
          vhi = unpack_bf16(*reinterpret_cast<const uint32_t*>(sm.Vr + sw128_off(row_hi, c)));
          const float a0 = sm.vs1[c], a1 = sm.vs1[c + 1], b0 = sm.vsL[c], b1 = sm.vsL[c + 1];
          *reinterpret_cast<uint32_t*>(dqkv + in_lo + 2 * hd + c) = pack_bf16(a[0] * a0 + b[0] * b0, a[1] * a1 + b[1] * b1);
          *reinterpret_cast<uint32_t*>(dqkv + in_hi + 2 * hd + c) = pack_bf16(a[2] * a0 + b[2] * b0, a[3] * a1 + b[3] * b1);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float s1 = a[e] * (e ? vlo.y : vlo.x) + a[2 + e] * (e ? vhi.y : vhi.x);
          float sl = b[e] * (e ? vlo.y : vlo.x) + b[2 + e] * (e ? vhi.y : vhi.x);
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); sl += __shfl_xor_sync(0xffffffffu, sl, o); }
          if (lane < 4) { sm.red[0][sp][c + e] = s1; sm.red[1][sp][c + e] = sl; }
        }
      }
      // db_q[j] = D[j][k] + D[j][4+k] (hi + lo operand slots): slots 0..3 sit in lanes with lane%4 < 2, slots 4..7 two lanes on
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float other = __shfl_xor_sync(0xffffffffu, vb[e], 2);
        if ((lane & 3) < 2) sm.db[4 * cb + cq + (e & 1)][(e & 2) ? row_hi : row_lo] = vb[e] + other;
      }
    }
    __syncthreads();
    if (tid < V * dk) {
      const int k = tid / dk, d = tid % dk;
      float val = 0.f;
      if (k == 0) val = (sm.red[0][0][d] + sm.red[0][1][d]) + (sm.red[0][2][d] + sm.red[0][3][d]);
      if (k == V - 1) val += w * ((sm.red[1][0][d] + sm.red[1][1][d]) + (sm.red[1][2][d] + sm.red[1][3][d]));
      acc_sv += val;
    }
    // feature-mean gradients, already combined the way dS_k and the chain seeds use them.  thread = (token, output group):
    // outputs o < V: row term of view o; V <= o < 2V: column term of view o - V; 2V .. 2V+3: seed terms
    {
      const int tok = tid & 63, grp = tid >> 6;
      float dav[kMaxQ], dbv[kMaxQ];
#pragma unroll
      for (int qq = 0; qq < kMaxQ; ++qq) { dav[qq] = sm.da[qq][tok]; dbv[qq] = sm.db[qq][tok]; }
      for (int o = grp; o < 2 * V + 4; o += 8) {
        // out = (1/64) sum_q ( hw_row[q][cr] da[q] [+ hw_col[q][cc] db[q]] )
        int cr = -1, cc = -1;
        if (o < V) { cr = o; cc = V + o; }
        else if (o < 2 * V) { cc = o - V; cr = o; }
        else if (o == 2 * V) cr = 2 * V;
        else if (o == 2 * V + 1) cc = 2 * V;
        else if (o == 2 * V + 2) cr = 2 * V + 1;
        else cc = 2 * V + 1;
        float acc = 0.f;
#pragma unroll
        for (int qq = 0; qq < kMaxQ; ++qq) {
          const int t = qq >> 2, k = qq & 3;
          if (k < r) {
            const int q = t * r + k;
            if (cr >= 0) acc = fmaf(hw_row[q * C + cr], dav[qq], acc);
            if (cc >= 0) acc = fmaf(hw_col[q * C + cc], dbv[qq], acc);
          }
        }
        acc *= (1.f / 64.f);
        if (o < V) sm.rt[o][tok] = acc;
        else if (o < 2 * V) sm.ct[o - V][tok] = acc;
        else sm.sd[o - 2 * V][tok] = acc;
      }
    }
    __syncthreads();
    // gate-head parameter partials <da_q, feature_c>, <db_q, feature_c> and the bias sums: thread = (projection, slot q, channel
    // half, 8-token group); runs while the first sweep step's MMAs are in flight
    auto head_partials = [&]() {
      const int tg = tid & 7, chalf = (tid >> 3) & 1, qq = (tid >> 4) & 15, half = tid >> 8;
      const int t = qq >> 2, k = qq & 3;
      const bool live = k < r;   // unused slots hold zeros; no early exit: the shuffles below need every lane
      const int q = t * r + k, nW = 4 * r * C, nP = nW + 4 * r;
      float* dh = sm.acc_head + (size_t)half * nP;
      const float* dv = (half ? sm.db[qq] : sm.da[qq]) + 8 * tg;
      const float4 d0 = *reinterpret_cast<const float4*>(dv), d1 = *reinterpret_cast<const float4*>(dv + 4);
      auto tg_sum = [&](float v) {
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        return v;
      };
      const int cbeg = chalf ? (C + 1) / 2 : 0, cend = chalf ? C : (C + 1) / 2;
      for (int c = cbeg; c < cend; ++c) {
        const float* ft;   // feature of channel c as seen by the row (half 0) / column (half 1) projection
        if (c < V) ft = half ? kap[c] : rho[c];
        else if (c < 2 * V) ft = half ? rho[c - V] : kap[c - V];
        else ft = half ? kap[c] : rho[c];
        const float4 f0 = *reinterpret_cast<const float4*>(ft + 8 * tg), f1 = *reinterpret_cast<const float4*>(ft + 8 * tg + 4);
        float a = d0.x * f0.x;
        a = fmaf(d0.y, f0.y, a); a = fmaf(d0.z, f0.z, a); a = fmaf(d0.w, f0.w, a);
        a = fmaf(d1.x, f1.x, a); a = fmaf(d1.y, f1.y, a); a = fmaf(d1.z, f1.z, a); a = fmaf(d1.w, f1.w, a);
        a = tg_sum(a);
        if (tg == 0 && live) dh[q * C + c] += a;
      }
      const float bsum = tg_sum(((d0.x + d0.y) + (d0.z + d0.w)) + ((d1.x + d1.y) + (d1.z + d1.w)));
      if (chalf && tg == 0 && live) dh[nW + q] += bsum;
    };
    // =========================================================================================================================
    // chain seeds X_F = Hf + G + dfeat_{2V} / (F + eps), X_R = dfeat_{2V+1} / (R + eps);  d logit = (1 - w) sum F (.) G
    // =========================================================================================================================
    {
      float hf[8], gg[8], Fv[8], Rv[8];
      tmem_ld_16x256b_x2(tcol(tW4), hf);
      tmem_ld_16x256b_x2(tcol(tG), gg);
      tmem_ld_16x256b_x2(tcol(tF), Fv);
      tmem_ld_16x256b_x2(tcol(tR), Rv);
      tmem_ld_wait();
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = c0 + 8 * (e >> 2) + cq + (e & 1), row = (e & 2) ? row_hi : row_lo;
        dot = fmaf(Fv[e], gg[e], dot);
        hf[e] = hf[e] + gg[e] + (sm.sd[0][row] + sm.sd[1][col]) * fast_rcp(Fv[e] + eps);
        Rv[e] = (sm.sd[2][row] + sm.sd[3][col]) * fast_rcp(Rv[e] + eps);
      }
      put_bf16(sm.X[0], cb, hf);
      put_bf16(sm.X[2], cb, Rv);
      dot = warp_sum(dot);
      if (lane == 0) sm.wsum[wid] = dot;
    }
    // =========================================================================================================================
    // chain sweep.  F = A_0..A_{V-1}: dA_k += P(k-1)^T X, X <- X A_k^T for k = V-1..1, dA_0 += X.
    //               R = A_{V-1}..A_0: dA_k += R(V-2-k)^T X, X <- X A_k^T for k = 0..V-2, dA_{V-1} += X.
    //   Both contributions of a view accumulate in ONE fp32 TMEM tile (first writer overwrites).
    // =========================================================================================================================
    {
      int xf = 0, xr = 2;
      uint32_t touched = 0;
      for (int s = 0; s <= V - 2; ++s) {
        const int kF = V - 1 - s, kR = s;
        publish();
        if (s == 0 && tid == 0) {
          float tot = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) tot += sm.wsum[i];
          p.dlogit_part[g] = (1.f - w) * tot;
        }
        if (leader) {
          const uint32_t pPrev = (kF - 1 == 0) ? sa(sm.A[0]) : sa(sm.P[kF - 2]);
          const uint32_t rNext = (kR + 1 == V - 1) ? sa(sm.A[V - 1]) : sa(sm.R[V - 3 - kR]);
          const bool accF = (touched >> kF) & 1u, accR = ((touched >> kR) & 1u) || kR == kF;
          if (mine(0)) {
            gemm(tdA + kF, 0, pPrev, CM_MN, sa(sm.X[xf]), CM_MN, accF, 4, 64);
            if (kR == kF) gemm(tdA + kR, 0, rNext, CM_MN, sa(sm.X[xr]), CM_MN, true, 4, 64);   // same tile: same issuer, in order
          }
          if (mine(1) && kR != kF) gemm(tdA + kR, 0, rNext, CM_MN, sa(sm.X[xr]), CM_MN, accR, 4, 64);
          if (mine(2)) gemm(tXF, 0, sa(sm.X[xf]), CM_K, sa(sm.A[kF]), CM_K, false, 4, 64);
          if (mine(3)) gemm(tXR, 0, sa(sm.X[xr]), CM_K, sa(sm.A[kR]), CM_K, false, 4, 64);
          mma_commit(&sm.bar_mma);
        }
        touched |= (1u << kF) | (1u << kR);
        if (s == 0) head_partials();
        wait_mma();
        if (s < V - 2) {
          xf ^= 1;
          xr ^= 1;
          float v[8];
          tmem_ld_16x256b_x2(tcol(tXF), v);
          tmem_ld_wait();
          put_bf16(sm.X[xf], cb, v);
          tmem_ld_16x256b_x2(tcol(tXR), v);
          tmem_ld_wait();
          put_bf16(sm.X[xr], cb, v);
        }
      }
    }
    // =========================================================================================================================
    // final pass per view: dS_k = C_k D + feature terms + A_k (.) (dA_k - rowsum(dA_k (.) A_k))  ->  bf16 tile (over A_k)
    // =========================================================================================================================
    auto load_dA = [&](int k, float* x, float* pk) {
      tmem_ld_16x256b_x2(tcol(tdA + k), x);
      get_bf16(sm.A[k], cb, pk);
      tmem_ld_wait();
      if (k == 0 || k == V - 1) {   // last links of the two chains (fp32 accumulators of the last sweep step)
        float y2[8];
        if (k == 0) {
          tmem_ld_16x256b_x2(tcol(tXF), y2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] += y2[e];
        }
        if (k == V - 1) {
          tmem_ld_16x256b_x2(tcol(tXR), y2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] += y2[e];
        }
      }
    };
    for (int k = 0; k < V; ++k) {   // row dots of every view first: ONE exchange between the four column blocks
      float x[8], pk[8];
      load_dA(k, x, pk);
      float lo = 0.f, hi = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { if (e & 2) hi = fmaf(x[e], pk[e], hi); else lo = fmaf(x[e], pk[e], lo); }
      lo = quad_sum(lo);
      hi = quad_sum(hi);
      if ((lane & 3) == 0) { sm.red[k][cb][row_lo] = lo; sm.red[k][cb][row_hi] = hi; }
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + sp) : "memory");   // the four warps (one per column block) that share these rows
    for (int k = 0; k < V; ++k) {
      float x[8], pk[8], cd[8];
      tmem_ld_16x256b_x2(tcol(tS + k), cd);
      load_dA(k, x, pk);
      const float dlo = (sm.red[k][0][row_lo] + sm.red[k][1][row_lo]) + (sm.red[k][2][row_lo] + sm.red[k][3][row_lo]);
      const float dhi = (sm.red[k][0][row_hi] + sm.red[k][1][row_hi]) + (sm.red[k][2][row_hi] + sm.red[k][3][row_hi]);
      const float rlo = sm.rt[k][row_lo], rhi = sm.rt[k][row_hi];
      const float2 c01 = *reinterpret_cast<const float2*>(&sm.ct[k][c0 + cq]), c23 = *reinterpret_cast<const float2*>(&sm.ct[k][c0 + 8 + cq]);
      const float ctv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float ft = ((e & 2) ? rhi : rlo) + ctv[2 * (e >> 2) + (e & 1)];
        x[e] = fmaf(pk[e], x[e] - ((e & 2) ? dhi : dlo), cd[e] + ft);
      }
      put_bf16(sm.A[k], cb, x);
    }
    // =========================================================================================================================
    // batch 3: T_k = dS_k K, U_k = dS_k^T Q; dQ = sum_k T_k (.) c_k, dK = sum_k U_k (.) c_k, scale partials
    // =========================================================================================================================
    publish();
    // V, dY and the aux vectors of this problem are consumed (every thread is past the head partials): fetch the next problem's
    if (tid == 0 && g + (int)gridDim.x < G) load_vdy(g + gridDim.x);
    if (leader) {
      for (int k = 0; k < V; ++k) {
        if (mine(2 * k)) gemm(tS + k, 0, sa(sm.A[k]), CM_K, sa(sm.Kr), SW_MN, false, 4, 64);
        if (mine(2 * k + 1)) gemm(tdA + k, 0, sa(sm.A[k]), CM_MN, sa(sm.Qr), SW_MN, false, 4, 64);
      }
      mma_commit(&sm.bar_mma);
    }
    wait_mma();
    {
      float dq[8], dkk[8], qf[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { dq[e] = 0.f; dkk[e] = 0.f; }
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int c = c0 + 8 * n + cq;
        const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(sm.Qr + sw128_off(row_lo, c)));
        const float2 b = unpack_bf16(*reinterpret_cast<const uint32_t*>(sm.Qr + sw128_off(row_hi, c)));
        qf[4 * n + 0] = a.x; qf[4 * n + 1] = a.y; qf[4 * n + 2] = b.x; qf[4 * n + 3] = b.y;
      }
      for (int k = 0; k < V; ++k) {
        float tq[8], tk[8];
        tmem_ld_16x256b_x2(tcol(tS + k), tq);
        tmem_ld_16x256b_x2(tcol(tdA + k), tk);
        tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int c = c0 + 8 * n + cq;
          const float2 cv = *reinterpret_cast<const float2*>(&sm.cvec[k][c]);
          dq[4 * n + 0] = fmaf(tq[4 * n + 0], cv.x, dq[4 * n + 0]); dq[4 * n + 1] = fmaf(tq[4 * n + 1], cv.y, dq[4 * n + 1]);
          dq[4 * n + 2] = fmaf(tq[4 * n + 2], cv.x, dq[4 * n + 2]); dq[4 * n + 3] = fmaf(tq[4 * n + 3], cv.y, dq[4 * n + 3]);
          dkk[4 * n + 0] = fmaf(tk[4 * n + 0], cv.x, dkk[4 * n + 0]); dkk[4 * n + 1] = fmaf(tk[4 * n + 1], cv.y, dkk[4 * n + 1]);
          dkk[4 * n + 2] = fmaf(tk[4 * n + 2], cv.x, dkk[4 * n + 2]); dkk[4 * n + 3] = fmaf(tk[4 * n + 3], cv.y, dkk[4 * n + 3]);
#pragma unroll
          for (int e = 0; e < 2; ++e) {   // Z_k[d] = sum_i T_k[i,d] Q[i,d]
            float z = tq[4 * n + e] * qf[4 * n + e] + tq[4 * n + 2 + e] * qf[4 * n + 2 + e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
            if (lane < 4) sm.red[k][sp][c + e] = z;
          }
        }
      }
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int c = c0 + 8 * n + cq;
        if (c < dk) {
          *reinterpret_cast<uint32_t*>(dqkv + in_lo + c) = pack_bf16(dq[4 * n + 0], dq[4 * n + 1]);
          *reinterpret_cast<uint32_t*>(dqkv + in_hi + c) = pack_bf16(dq[4 * n + 2], dq[4 * n + 3]);
          *reinterpret_cast<uint32_t*>(dqkv + in_lo + hd + c) = pack_bf16(dkk[4 * n + 0], dkk[4 * n + 1]);
          *reinterpret_cast<uint32_t*>(dqkv + in_hi + hd + c) = pack_bf16(dkk[4 * n + 2], dkk[4 * n + 3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // tiles, vectors and TMEM are reused by the next problem; Q and K tiles are free
    if (tid == 0 && g + (int)gridDim.x < G) load_qk(g + gridDim.x);
    if (tid < V * dk) {
      const int k = tid / dk, d = tid % dk;
      const float z = sscale * ((sm.red[k][0][d] + sm.red[k][1][d]) + (sm.red[k][2][d] + sm.red[k][3][d]));
      const size_t pi = ((size_t)k * H + ph) * dk + d;
      acc_sq = fmaf(p.k_scale[pi], z, acc_sq);
      acc_sk = fmaf(p.q_scale[pi], z, acc_sk);
    }
    __syncthreads();   // red[] is written again early in the next problem
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  {   // one row of parameter-gradient partials per CTA
    const int nP = 4 * r * C + 4 * r;
    float* dh = p.dhead_part + (size_t)blockIdx.x * 2 * nP;
    for (int idx = tid; idx < 2 * nP; idx += kThreads) dh[idx] = sm.acc_head[idx];
    if (tid < V * dk) {
      float* ds = p.dscale_part + (size_t)blockIdx.x * 3 * V * dk;
      ds[tid] = acc_sq;
      ds[V * dk + tid] = acc_sk;
      ds[2 * V * dk + tid] = acc_sv;
    }
  }
  if (wid == 0) tmem_dealloc<512>(tbase);
}

inline bool supported_bwd(const MopEdgewiseParams* p) { return ewtc::supported(p) && p->aux != nullptr; }
// persistent grid of the backward: a multiple of H so that every CTA sees ONE head (its scale-gradient partials are per head)
inline int bwd_grid(const MopEdgewiseParams* p, int sms) {
  const int G = p->B * p->H, cap = G < sms ? G : sms;
  const int g = cap / p->H * p->H;
  return g > 0 ? g : p->H;
}

}  // namespace ew64
}  // namespace mop
