// Edgewise (Mixture-of-Products) attention forward on tcgen05 / TMEM for token counts up to 200
// (ViT-B/16: N = 196, dk = 64, V = 5): bf16 operands, fp32 accumulation and fp32 statistics.
//
// One persistent CTA of 256 threads (two warpgroups) per SM owns one (batch, head) problem at a time:
//   * every N x N map is cut into two M=128 row blocks; warpgroup w owns row block w and reads its fp32
//     accumulator (TMEM columns [256w, 256w + 208)) with the 32x32b shape, i.e. ONE THREAD PER ROW: row
//     softmax statistics are thread local, column statistics are a 16-step shuffle butterfly + shared atomics;
//   * bf16 MMA operands live in shared memory in the chunk-major layout of tc_common.cuh:
//       A  (86.5 KB)  the current per-view softmax A_k  (B operand of the chain products, K = its rows)
//       X  (83.2 KB)  the running chain product (A operand; row block w is rewritten in place by warpgroup w)
//       Q  (25.6 KB)  unscaled queries (A operand),  K (26.6 KB)  keys scaled by q_scale*k_scale/sqrt(dk) per view
//     which is all of the 227 KB;
//   * pass R (views V-1..0):  S_k = Q Ks_k^T -> softmax -> A_k ;  Y <- Y A_k        -> R = A_{V-1}..A_0
//     pass F (views 1..V-1):  A_k comes back by bulk copy      ;  X <- X A_k        -> F = A_0..A_{V-1} (stays in X)
//     Between the passes the bf16 A_k wait in a per-CTA scratch slot of the caller's workspace (V x 86.5 KB per
//     CTA, 64 MB for 148 CTAs: L2 resident; moved by cp.async.bulk in both directions).  Only the row / column
//     means of S_k, log(F+eps), log(R+eps) leave the passes (gate features);
//   * final stage, flash style over 32-column panels: the V score panels are recomputed into TMEM, mixed with
//     the rank-r gates in registers (AND sum, OR log-sum-exp, NOT, chain log F from X), online softmax,
//     P V_1 accumulated in TMEM; y = (P V_1)/l + F (w V_V).
//
// Math: SURVEY.md appendix A (reference attention_variants.py:500-562, :319-331); executable specification
// oracle/edgewise_manual.py.  The token-count-64 specialisation is edgewise_tc.cuh.
#pragma once
#include "edgewise_tc.cuh"

namespace mop {
namespace ewl {

using namespace tc;
using ewtc::fast_exp2;
using ewtc::fast_log2;
using ewtc::fast_rcp;
using ewtc::kLn2;
using ewtc::kLog2e;
using ewtc::kMaxQ;
using ewtc::kMaxV;
using ewtc::scale_chunk;

constexpr int kNmax = 208;                    // 13 MMA k-steps (padded token count)
constexpr int kMaxTokens = 200;               // rows held by the X / Q tiles
constexpr int kRA = 208;                      // rows of the A / K / V tiles (K-dimension operands: every row finite)
constexpr int kRX = 200;                      // rows of the X / Q tiles (M-dimension operands: over-read is harmless)
constexpr int kMapChunks = kNmax / 8;         // 26 column chunks of 8
constexpr int kBufX = kRX * 16 * kMapChunks;  // 83200
constexpr int kBufA = kRA * 16 * kMapChunks;  // 86528
constexpr int kQt = kRX * 16 * 8;             // 25600
constexpr int kKt = kRA * 16 * 8;             // 26624
constexpr int kPanel = 32;                    // final-stage panel width (columns)
constexpr int kKsP = kPanel * 16 * 8;         // one scaled key panel: 4096

struct __align__(128) Smem {
  unsigned char X[kBufX];
  unsigned char A[kBufA];
  unsigned char Q[kQt];
  unsigned char K[kKt];
  float colsum[kMaxV + 2][kNmax];  // column sums of S_k (k < V), log F (V), log R (V+1)
  float cvec[kMaxV][64];           // q_scale*k_scale/sqrt(dk) per view
  float vs1[64], vsL[64];          // v_scale[0], sigmoid(chain_value_logit) * v_scale[V-1]
  uint64_t bar[2];                 // MMA completion, one per warpgroup
  uint32_t tmem_slot;
};
// final-stage aliases inside A (dead once the chain passes are done)
constexpr int kOffBfac = 0;                              // float [208][16]: column gate factors
constexpr int kOffVt = kNmax * 16 * 4;                   // 13312: value tile (V_1, later w V_V)
constexpr int kOffKsP = kOffVt + kKt;                    // 39936: [2 warpgroups][kMaxV][kKsP]
static_assert(kOffKsP + 2 * kMaxV * kKsP <= kBufA, "final-stage aliases overflow the A buffer");
// P panels ([2][128 rows x 32 cols] bf16) alias the K tile
constexpr int kPt = 128 * 16 * (kPanel / 8);             // 8192
static_assert(2 * kPt <= kKt, "P panels overflow the K tile");

// sigmoid from ex2 + rcp (relative error ~1e-7).  tanh.approx (2^-11) is not enough here: the column-factor gradients of the
// gate head are residuals of row sums of D g(1-g) that cancel to ~1 % of their terms, and forward and backward must use the
// same gate values for the rows of D to sum to zero.
__device__ __forceinline__ float fast_sigmoid(float x) { return fast_rcp(1.f + fast_exp2(-kLog2e * x)); }

__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

// 16 fp32 -> 2 x (8 bf16), optionally scaled
// (callers zero v[] for padded rows themselves: 0 * NaN would not be zero)
__device__ __forceinline__ void pack16(const float* v, float sc, uint4& lo, uint4& hi) {
  lo.x = pack_bf16(v[0] * sc, v[1] * sc); lo.y = pack_bf16(v[2] * sc, v[3] * sc);
  lo.z = pack_bf16(v[4] * sc, v[5] * sc); lo.w = pack_bf16(v[6] * sc, v[7] * sc);
  hi.x = pack_bf16(v[8] * sc, v[9] * sc); hi.y = pack_bf16(v[10] * sc, v[11] * sc);
  hi.z = pack_bf16(v[12] * sc, v[13] * sc); hi.w = pack_bf16(v[14] * sc, v[15] * sc);
}


static __global__ void __launch_bounds__(256, 1) edgewise_fwd_kernel(MopEdgewiseParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp4 = (tid >> 5) & 3, lane = tid & 31;
  const int N = p.N, V = p.V, r = p.gate_rank, C = 2 * V + 2, dk = p.dk, H = p.H;
  const int KS = (N + 15) >> 4, NN = KS * 16;   // MMA k-steps over tokens; padded token count
  const int cfull = N >> 4;                     // column chunks of 16 without padding
  const int dks = (dk + 15) >> 4;
  const int row = 128 * wg + t;
  const bool row_ok = row < N;
  const bool blk_on = 128 * wg < N;                   // this warpgroup owns rows
  const bool warp_on = 128 * wg + 32 * warp4 < N;     // this warp owns at least one valid row
  const float invN = 1.f / (float)N;
  const uint32_t map_bytes = (uint32_t)(2 * KS) * (kRA * 16);   // the chunks of A that hold data
  unsigned char* spill = reinterpret_cast<unsigned char*>(p.workspace) + (size_t)blockIdx.x * kMaxV * kBufA;

  if (tid < 32) tmem_alloc<512>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_slot;
  const uint32_t tD = tbase + 256u * (uint32_t)wg;                      // accumulator of this warpgroup (MMA address)
  const uint32_t tl = tD + ((uint32_t)(32 * warp4) << 16);              // the same, this warp's lane window
  uint32_t phase = 0;
  const float w = 1.f / (1.f + __expf(-p.chain_value_logit[0]));
  const float bn = p.beta_not / (float)max(1, V - 1);
  const float sscale = rsqrtf((float)dk);
  const uint32_t sX = smem_u32(sm.X), sA = smem_u32(sm.A), sQ = smem_u32(sm.Q), sK = smem_u32(sm.K);

  auto mma_wait = [&]() { mbar_wait(&sm.bar[wg], phase); phase ^= 1; tc_fence_after(); };
  auto publish_cta = [&]() { fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after(); };
  auto publish_wg = [&]() { fence_async_smem(); tc_fence_before(); wg_sync(wg); tc_fence_after(); };
  // D = X[row block] A  (A: the map in the A buffer, K index = its rows)
  auto chain_mma = [&]() {
    const uint32_t id = idesc_bf16(128, NN, 0, 1);
    for (int ks = 0; ks < KS; ++ks)
      mma_ss(tD, desc_kmajor(sX + 128 * wg * 16, kRX, 16 * ks), desc_mnmajor(sA, kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
    mma_commit(&sm.bar[wg]);
  };
  // accumulator -> bf16 row of X (STORE) and / or row + column sums of log(D + eps) (LOGS)
  auto chain_epilogue = [&](bool store, bool logs, int slot, float& rowmean) {
    float ls = 0.f;
    for (int c = 0; c < KS; ++c) {
      float v[16];
      tmem_ld_32x32b_x16(tl + 16 * c, v);
      tmem_ld_wait();
      if (store && row < kRX) {
        uint4 lo, hi;
        pack16(v, 1.f, lo, hi);
        *reinterpret_cast<uint4*>(sm.X + (2 * c) * (kRX * 16) + row * 16) = lo;
        *reinterpret_cast<uint4*>(sm.X + (2 * c + 1) * (kRX * 16) + row * 16) = hi;
      }
      if (logs) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const bool ok = row_ok && 16 * c + e < N;
          v[e] = ok ? kLn2 * fast_log2(v[e] + p.eps) : 0.f;
          ls += v[e];
        }
        int col;
        const float cs = warp_colsum16(v, lane, &col);
        if ((lane & 1) == 0) atomicAdd(&sm.colsum[slot][16 * c + col], cs);
      }
    }
    if (logs) rowmean = ls * invN;
  };

  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p.qkv);
  const size_t hd = (size_t)H * dk;
  const int G = p.B * H;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int pb = g / H, ph = g % H;
    auto in_row = [&](int n) { return qkv + (((size_t)pb * N + n) * 3) * hd + (size_t)ph * dk; };   // q; +hd: k; +2hd: v
    // =================================================================================================
    // stage 0: per-view scale vectors, unscaled Q tile, this thread's key row (registers), zeroed column sums
    // =================================================================================================
    for (int idx = tid; idx < V * 64; idx += 256) {
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
    for (int idx = tid; idx < (kMaxV + 2) * kNmax; idx += 256) (&sm.colsum[0][0])[idx] = 0.f;
    uint4 kraw[8];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      uint4 q = make_uint4(0, 0, 0, 0);
      kraw[ch] = q;
      if (tid < N && ch * 8 < dk) {
        q = *reinterpret_cast<const uint4*>(in_row(tid) + ch * 8);
        kraw[ch] = *reinterpret_cast<const uint4*>(in_row(tid) + hd + ch * 8);
      }
      if (tid < kRX) *reinterpret_cast<uint4*>(sm.Q + ch * (kRX * 16) + tid * 16) = q;
    }
    __syncthreads();
    float rho[kMaxV], rhoF = 0.f, rhoR = 0.f;   // row means of S_k, log F, log R of this thread's row
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) rho[i] = 0.f;
    // =================================================================================================
    // pass R (views V-1 .. 0): A_k = softmax(S_k) (spilled to the L2-resident scratch), Y <- Y A_k
    // =================================================================================================
    for (int idx = 0; idx < V; ++idx) {
      const int k = V - 1 - idx;
      const bool first = idx == 0, last = idx == V - 1;
      // scaled keys of view k (every MMA issued so far has completed: each thread waited for its warpgroup's
      // MMAs before the last CTA barrier)
      if (tid < kRA) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<uint4*>(sm.K + ch * (kRA * 16) + tid * 16) = scale_chunk(kraw[ch], &sm.cvec[k][ch * 8]);
      }
      publish_cta();
      if (blk_on) {
        if (t == 0) {
          const uint32_t id = idesc_bf16(128, NN, 0, 0);
          for (int ks = 0; ks < dks; ++ks)
            mma_ss(tD, desc_kmajor(sQ + 128 * wg * 16, kRX, 16 * ks), desc_kmajor(sK, kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
          mma_commit(&sm.bar[wg]);
        }
        mma_wait();
      }
      // ---- row softmax of S_k, thread per row; row / column sums of S_k for the gate features
      if (warp_on) {
        float mx = -INFINITY, rs = 0.f;
        for (int c = 0; c < KS; ++c) {
          float v[16];
          tmem_ld_32x32b_x16(tl + 16 * c, v);
          tmem_ld_wait();
          if (c < cfull) {
#pragma unroll
            for (int e = 0; e < 16; ++e) { mx = fmaxf(mx, v[e]); rs += v[e]; }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if (16 * c + e < N) { mx = fmaxf(mx, v[e]); rs += v[e]; } else v[e] = 0.f;
            }
          }
          if (!row_ok) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0.f;
          }
          int col;
          const float cs = warp_colsum16(v, lane, &col);
          if ((lane & 1) == 0) atomicAdd(&sm.colsum[k][16 * c + col], cs);
        }
#pragma unroll
        for (int i = 0; i < kMaxV; ++i)
          if (i == k) rho[i] = rs * invN;
        float l = 0.f;
        const float mb = mx * kLog2e;
        for (int c = 0; c < KS; ++c) {
          float v[16];
          tmem_ld_32x32b_x16(tl + 16 * c, v);
          tmem_ld_wait();
          if (c < cfull) {
#pragma unroll
            for (int e = 0; e < 16; ++e) { v[e] = fast_exp2(fmaf(v[e], kLog2e, -mb)); l += v[e]; }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) { v[e] = (16 * c + e < N) ? fast_exp2(fmaf(v[e], kLog2e, -mb)) : 0.f; l += v[e]; }
          }
          tmem_st_32x32b_x16(tl + 16 * c, v);
        }
        tmem_st_wait();
        const float inv_l = 1.f / l;
        for (int c = 0; c < KS; ++c) {
          float v[16];
          tmem_ld_32x32b_x16(tl + 16 * c, v);
          tmem_ld_wait();
          uint4 lo, hi;
          pack16(v, inv_l, lo, hi);
          if (!row_ok) lo = hi = make_uint4(0, 0, 0, 0);   // padded rows are zero rows of A_k
          if (row < kRA) {
            *reinterpret_cast<uint4*>(sm.A + (2 * c) * (kRA * 16) + row * 16) = lo;
            *reinterpret_cast<uint4*>(sm.A + (2 * c + 1) * (kRA * 16) + row * 16) = hi;
            if (k >= 1) {   // pass F needs A_k again: its tile image goes to the scratch slot (L2)
              *reinterpret_cast<uint4*>(spill + (size_t)k * kBufA + (2 * c) * (kRA * 16) + row * 16) = lo;
              *reinterpret_cast<uint4*>(spill + (size_t)k * kBufA + (2 * c + 1) * (kRA * 16) + row * 16) = hi;
            }
          }
          if (first && row < kRX) {
            *reinterpret_cast<uint4*>(sm.X + (2 * c) * (kRX * 16) + row * 16) = lo;
            *reinterpret_cast<uint4*>(sm.X + (2 * c + 1) * (kRX * 16) + row * 16) = hi;
          }
        }
      } else if (row < kRA) {
        // rows of A_k that no active warp writes: they are K-dimension rows of the chain MMAs, keep them zero
        for (int c = 0; c < 2 * KS; ++c) *reinterpret_cast<uint4*>(sm.A + c * (kRA * 16) + row * 16) = make_uint4(0, 0, 0, 0);
      }
      publish_cta();
      if (first) continue;
      if (blk_on) {
        if (t == 0) chain_mma();
        mma_wait();
      }
      if (warp_on) chain_epilogue(!last, last, V + 1, rhoR);   // R itself is never needed again
      if (last && row < kRX) {
        // X <- A_0: start of the forward chain (this warpgroup's MMA, the only reader of these X rows, is complete)
        for (int c = 0; c < 2 * KS; ++c)
          *reinterpret_cast<uint4*>(sm.X + c * (kRX * 16) + row * 16) =
              row < kRA ? *reinterpret_cast<const uint4*>(sm.A + c * (kRA * 16) + row * 16) : make_uint4(0, 0, 0, 0);
      }
    }
    // =================================================================================================
    // pass F (views 1 .. V-1): A_k comes back from the scratch, X <- X A_k; F stays in X
    // =================================================================================================
    publish_cta();   // X = A_0 visible to the MMAs; nobody reads the A buffer any more
    // (a single cp.async.bulk of the 86 KB image was measured at ~4 B/cycle; 256 threads x cp.async 16 B are ~10x faster)
    cp_async_block(sm.A, spill + (size_t)1 * kBufA, map_bytes);
    cp_async_commit();
    for (int k = 1; k < V; ++k) {
      const bool last = k == V - 1;
      cp_async_wait<0>();
      publish_cta();   // A_k landed (and, for k > 1, the rewritten X rows are visible)
      if (blk_on) {
        if (t == 0) chain_mma();
        mma_wait();
      }
      if (!last) {
        tc_fence_before();
        __syncthreads();   // both warpgroups' MMAs have read A_k: the next map may land while X is rewritten
        cp_async_block(sm.A, spill + (size_t)(k + 1) * kBufA, map_bytes);
        cp_async_commit();
      }
      if (warp_on) chain_epilogue(true, last, V, rhoF);
    }
    // =================================================================================================
    // final stage
    // =================================================================================================
    publish_cta();   // F complete in X; chain MMAs of both warpgroups are done: A and K buffers are free, colsum is final
    float* bfac = reinterpret_cast<float*>(sm.A + kOffBfac);
    unsigned char* Vt = sm.A + kOffVt;
    unsigned char* KsP = sm.A + kOffKsP + wg * (kMaxV * kKsP);
    unsigned char* Pt = sm.K + wg * kPt;
    auto load_values = [&](const float* vscale) {
      if (tid < kRA) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4 vv = make_uint4(0, 0, 0, 0);
          if (tid < N && ch * 8 < dk) vv = scale_chunk(*reinterpret_cast<const uint4*>(in_row(tid) + 2 * hd + ch * 8), &vscale[ch * 8]);
          *reinterpret_cast<uint4*>(Vt + ch * (kRA * 16) + tid * 16) = vv;
        }
      }
    };
    load_values(sm.vs1);
    // gate factors: a (row factors of this thread's row, registers) and b (column factors of column `row`, shared)
    float afac[kMaxQ];
    {
      float fr[2 * kMaxV + 2], fc[2 * kMaxV + 2];
#pragma unroll
      for (int c = 0; c < kMaxV; ++c) {
        const float kap = (c < V && row < kNmax) ? sm.colsum[c][row] * invN : 0.f;
        fr[c] = rho[c]; fr[kMaxV + c] = kap;     // row projection sees S_c (row mean) and S_c^T (row mean = column mean of S_c)
        fc[c] = kap;    fc[kMaxV + c] = rho[c];  // column projection: roles swapped
      }
      const float kapF = row < kNmax ? sm.colsum[V][row] * invN : 0.f, kapR = row < kNmax ? sm.colsum[V + 1][row] * invN : 0.f;
      fr[2 * kMaxV] = rhoF; fr[2 * kMaxV + 1] = rhoR;
      fc[2 * kMaxV] = kapF; fc[2 * kMaxV + 1] = kapR;
#pragma unroll
      for (int qq = 0; qq < kMaxQ; ++qq) {
        const int tg = qq >> 2, kk = qq & 3, q = tg * r + kk;
        float a = 0.f, b = 0.f;
        if (kk < r && row_ok) {
          a = __ldg(p.row_b + q);
          b = __ldg(p.col_b + q);
#pragma unroll
          for (int c = 0; c < kMaxV; ++c)
            if (c < V) {
              a = fmaf(__ldg(p.row_w + q * C + c), fr[c], a);
              a = fmaf(__ldg(p.row_w + q * C + V + c), fr[kMaxV + c], a);
              b = fmaf(__ldg(p.col_w + q * C + c), fc[c], b);
              b = fmaf(__ldg(p.col_w + q * C + V + c), fc[kMaxV + c], b);
            }
          a = fmaf(__ldg(p.row_w + q * C + 2 * V), fr[2 * kMaxV], a);
          a = fmaf(__ldg(p.row_w + q * C + 2 * V + 1), fr[2 * kMaxV + 1], a);
          b = fmaf(__ldg(p.col_w + q * C + 2 * V), fc[2 * kMaxV], b);
          b = fmaf(__ldg(p.col_w + q * C + 2 * V + 1), fc[2 * kMaxV + 1], b);
        }
        afac[qq] = a;
        if (row < kNmax) bfac[row * 16 + qq] = b;
      }
    }
    publish_cta();
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t tS = tD, tO = tD + 160;            // MMA addresses: score panels (V x 32 columns), P V_1 accumulator
    const uint32_t tlS = tl, tlO = tl + 160;          // this warp's lane window
    if (blk_on) {
      const int npanels = (N + kPanel - 1) / kPanel;
      // raw key chunks of the next panel, prefetched while the current one is processed
      uint4 raw[2];
      auto fetch_panel = [&](int j0) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int item = t + 128 * it, jj = item & 31, ch = item >> 5, j = j0 + jj;
          raw[it] = make_uint4(0, 0, 0, 0);
          if (j < N && ch * 8 < dk) raw[it] = *reinterpret_cast<const uint4*>(in_row(j) + hd + ch * 8);
        }
      };
      fetch_panel(0);
      for (int pn = 0; pn < npanels; ++pn) {
        const int j0 = pn * kPanel;
        // scaled key panels of the V views (the score MMAs of the previous panel have completed)
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int item = t + 128 * it, jj = item & 31, ch = item >> 5;
          for (int k = 0; k < V; ++k)
            *reinterpret_cast<uint4*>(KsP + k * kKsP + ch * (kPanel * 16) + jj * 16) = scale_chunk(raw[it], &sm.cvec[k][ch * 8]);
        }
        publish_wg();
        if (t == 0) {
          const uint32_t id = idesc_bf16(128, kPanel, 0, 0);
          for (int k = 0; k < V; ++k)
            for (int ks = 0; ks < dks; ++ks)
              mma_ss(tS + 32 * k, desc_kmajor(sQ + 128 * wg * 16, kRX, 16 * ks), desc_kmajor(smem_u32(KsP) + k * kKsP, kPanel, 16 * ks), id, ks > 0 ? 1u : 0u);
          mma_commit(&sm.bar[wg]);
        }
        if (pn + 1 < npanels) fetch_panel(j0 + kPanel);
        mma_wait();   // also covers the P V_1 MMA of the previous panel
        if (warp_on) {
          float pmax = -INFINITY;
#pragma unroll 1
          for (int sub = 0; sub < kPanel / 8; ++sub) {
            float sv[kMaxV][8];
#pragma unroll
            for (int i = 0; i < kMaxV; ++i)
              if (i < V) tmem_ld_32x32b_x8(tlS + 32 * i + 8 * sub, sv[i]);
            tmem_ld_wait();
            const int jc = j0 + 8 * sub;
            uint4 fraw = make_uint4(0, 0, 0, 0);
            if (row < kRX && jc < NN) fraw = *reinterpret_cast<const uint4*>(sm.X + (jc >> 3) * (kRX * 16) + row * 16);
            const uint32_t fw[4] = {fraw.x, fraw.y, fraw.z, fraw.w};
            float val[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = jc + e;
              const float2 f2 = unpack_bf16(fw[e >> 1]);
              const float fval = (e & 1) ? f2.y : f2.x;
              float s0 = sv[0][e], sum = s0, mxv = s0;
#pragma unroll
              for (int i = 1; i < kMaxV; ++i)
                if (i < V) { sum += sv[i][e]; mxv = fmaxf(mxv, sv[i][e]); }
              float se = 0.f;
#pragma unroll
              for (int i = 0; i < kMaxV; ++i)
                if (i < V) se += fast_exp2((sv[i][e] - mxv) * kLog2e);
              const float lse = mxv + kLn2 * fast_log2(se);
              const float U = sum - s0, O = lse - s0, lf = kLn2 * fast_log2(fval + p.eps);
              float z[4];
              const int jb = j < kNmax ? j : kNmax - 1;
              const float4* bp = reinterpret_cast<const float4*>(bfac + jb * 16);
#pragma unroll
              for (int tg = 0; tg < 4; ++tg) {
                const float4 b4 = bp[tg];
                z[tg] = fmaf(afac[4 * tg], b4.x, fmaf(afac[4 * tg + 1], b4.y, fmaf(afac[4 * tg + 2], b4.z, afac[4 * tg + 3] * b4.w)));
              }
              const float x = s0 + fast_sigmoid(z[0]) * U + fast_sigmoid(z[1]) * O - fast_sigmoid(z[2]) * bn * U + fast_sigmoid(z[3]) * lf;
              val[e] = (j < N && row_ok) ? x : -INFINITY;
              pmax = fmaxf(pmax, val[e]);
            }
            tmem_st_32x32b_x8(tlS + 8 * sub, val);   // park the mixed scores in the (consumed) columns of S_0
          }
          tmem_st_wait();
          // online softmax in base 2 with an INTEGER reference exponent per row: rescaling by exact powers of two
          // commutes with the bf16 rounding of P, and the row sum is taken over the ROUNDED probabilities, so that
          // y_base = (sum_j P_ij V_1j) / l_i is exactly the softmax-weighted mean the backward differentiates
          // (its D = A (dA - dY . y_base) then has zero row sums, which the gate-head gradients rely on).
          const float m_new = fmaxf(m_run, ceilf(pmax * kLog2e));
          const float mb = (m_new == -INFINITY) ? 0.f : m_new;
          float sc = 0.f;
          if (m_run != -INFINITY) {
            const int diff = max((int)(m_run - m_new), -126);
            sc = __int_as_float((127 + diff) << 23);
          }
          const bool need = pn > 0 && m_new > m_run;
          if (__any_sync(0xffffffffu, need)) {
            const float scl = need ? sc : 1.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float o[16];
              tmem_ld_32x32b_x16(tlO + 16 * c, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] *= scl;
              tmem_st_32x32b_x16(tlO + 16 * c, o);
            }
            tmem_st_wait();
          }
          l_run *= sc;
          m_run = m_new;
          float smix[kPanel];
          tmem_ld_32x32b_x32(tlS, smix);
          tmem_ld_wait();
          float ps = 0.f;
#pragma unroll
          for (int c = 0; c < kPanel / 8; ++c) {
            uint32_t u[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              u[e2] = pack_bf16(fast_exp2(fmaf(smix[8 * c + 2 * e2], kLog2e, -mb)), fast_exp2(fmaf(smix[8 * c + 2 * e2 + 1], kLog2e, -mb)));
              ps += __uint_as_float(u[e2] << 16) + __uint_as_float(u[e2] & 0xffff0000u);   // the rounded values (exp2(-inf) = 0)
            }
            *reinterpret_cast<uint4*>(Pt + c * (128 * 16) + t * 16) = make_uint4(u[0], u[1], u[2], u[3]);
          }
          l_run += ps;
        } else {
#pragma unroll
          for (int c = 0; c < kPanel / 8; ++c) *reinterpret_cast<uint4*>(Pt + c * (128 * 16) + t * 16) = make_uint4(0, 0, 0, 0);
        }
        publish_wg();
        if (t == 0) {
          const uint32_t id = idesc_bf16(128, 64, 0, 1);
          const int nks = min(kPanel, NN - j0) >> 4;
          for (int ks = 0; ks < nks; ++ks)
            mma_ss(tO, desc_kmajor(smem_u32(Pt), 128, 16 * ks), desc_mnmajor(smem_u32(Vt), kRA, j0 + 16 * ks), id, (pn > 0 || ks > 0) ? 1u : 0u);
          if (pn == npanels - 1) mma_commit(&sm.bar[wg]);   // otherwise covered by the next panel's commit
        }
      }
      mma_wait();
    }
    // ---- y = (P V_1) / l + F (w V_V)
    tc_fence_before();
    __syncthreads();          // both warpgroups are done with V_1
    load_values(sm.vsL);
    publish_cta();
    if (blk_on) {
      if (t == 0) {
        const uint32_t id = idesc_bf16(128, 64, 0, 1);
        for (int ks = 0; ks < KS; ++ks)
          mma_ss(tS, desc_kmajor(sX + 128 * wg * 16, kRX, 16 * ks), desc_mnmajor(smem_u32(Vt), kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
        mma_commit(&sm.bar[wg]);
      }
      mma_wait();
      if (warp_on) {
        const float il = 1.f / l_run;
        if (p.row_stats && row_ok) *reinterpret_cast<float2*>(p.row_stats + (((size_t)pb * H + ph) * N + row) * 2) = make_float2(m_run, l_run);
        __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + (((size_t)pb * N + row) * H + ph) * dk;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float oa[16], of[16];
          tmem_ld_32x32b_x16(tlO + 16 * c, oa);
          tmem_ld_32x32b_x16(tlS + 16 * c, of);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const int d0 = 16 * c + 8 * h8;
              if (d0 < dk) {
                if (p.y_base) {
                  float* yb = p.y_base + (((size_t)pb * N + row) * H + ph) * dk + d0;
                  *reinterpret_cast<float4*>(yb) = make_float4(oa[8 * h8] * il, oa[8 * h8 + 1] * il, oa[8 * h8 + 2] * il, oa[8 * h8 + 3] * il);
                  *reinterpret_cast<float4*>(yb + 4) = make_float4(oa[8 * h8 + 4] * il, oa[8 * h8 + 5] * il, oa[8 * h8 + 6] * il, oa[8 * h8 + 7] * il);
                }
                uint4 u;
                u.x = pack_bf16(fmaf(oa[8 * h8 + 0], il, of[8 * h8 + 0]), fmaf(oa[8 * h8 + 1], il, of[8 * h8 + 1]));
                u.y = pack_bf16(fmaf(oa[8 * h8 + 2], il, of[8 * h8 + 2]), fmaf(oa[8 * h8 + 3], il, of[8 * h8 + 3]));
                u.z = pack_bf16(fmaf(oa[8 * h8 + 4], il, of[8 * h8 + 4]), fmaf(oa[8 * h8 + 5], il, of[8 * h8 + 5]));
                u.w = pack_bf16(fmaf(oa[8 * h8 + 6], il, of[8 * h8 + 6]), fmaf(oa[8 * h8 + 7], il, of[8 * h8 + 7]));
                *reinterpret_cast<uint4*>(y + d0) = u;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // tiles, vectors and TMEM are reused by the next problem
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tbase);
}

inline bool supported(const MopEdgewiseParams* p) {
  return p->dtype == MOP_BF16 && p->N >= 1 && p->N <= kMaxTokens && p->dk <= 64 && p->dk % 8 == 0 && p->V >= 2 && p->V <= kMaxV &&
         p->Vp == 1 && p->gate_mode == MOP_GATE_LOWRANK && p->gate_rank >= 1 && p->gate_rank <= 4 && p->q_scale != nullptr && p->lens_n == 0;
}

// one persistent CTA per SM (sizing without a device: B200)
inline int grid_size(const MopEdgewiseParams* p) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int G = p->B * p->H;
  return G < sms ? G : sms;
}

}  // namespace ewl
}  // namespace mop
