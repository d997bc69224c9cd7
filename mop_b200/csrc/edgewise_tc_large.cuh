// Edgewise (Mixture-of-Products) attention forward on tcgen05 / TMEM for token counts up to 200
// (ViT-B/16: N = 196, dk = 64, V = 5): bf16 operands, fp32 accumulation and fp32 statistics.
//
// Shared pieces of the forward (edgewise_tc_large_fwd2.cuh, 512 threads) and the backward (edgewise_tc_large_bwd.cuh, 256 threads).
// One persistent CTA per SM owns one (batch, head) problem at a time:
//   * every N x N map is cut into two M=128 row blocks; row block w has its fp32 accumulator in TMEM columns
//     [256w, 256w + 208), read with the 32x32b shape, i.e. a thread sees ONE ROW (the backward runs one thread per row,
//     the forward two, each taking half of the columns): row softmax statistics are thread local (or one exchange
//     away), column statistics are a 16-step shuffle butterfly + shared atomics;
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
// what the forward hands to the backward per (b, h) problem (MopEdgewiseParams::aux, floats, [field][208 tokens] so that one
// thread per token reads / writes coalesced): per-view softmax statistics, the 14 feature means, the gate factors a and b
constexpr int kAuxLStats = 0;                             // [V][2][208]: max * log2(e), 1 / sum of the row of S_k
constexpr int kAuxLFeat = kAuxLStats + kMaxV * 2 * kNmax; // [14][208]: rho_0..4, kap_0..4, rhoF, rhoR, kapF, kapR
constexpr int kAuxLA = kAuxLFeat + (2 * kMaxV + 4) * kNmax;   // [16][208] row gate factors
constexpr int kAuxLB = kAuxLA + kMaxQ * kNmax;            // [16][208] column gate factors
constexpr int kAuxLFloats = kAuxLB + kMaxQ * kNmax;       // 11648 floats = 46.6 KB
constexpr int kKsP = kPanel * 16 * 8;         // one scaled key panel: 4096

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

__device__ __forceinline__ void unpack8(uint4 u, float* f) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}
// 16 fp32 -> 2 x (8 bf16), optionally scaled
// (callers zero v[] for padded rows themselves: 0 * NaN would not be zero)
__device__ __forceinline__ void pack16(const float* v, float sc, uint4& lo, uint4& hi) {
  lo.x = pack_bf16(v[0] * sc, v[1] * sc); lo.y = pack_bf16(v[2] * sc, v[3] * sc);
  lo.z = pack_bf16(v[4] * sc, v[5] * sc); lo.w = pack_bf16(v[6] * sc, v[7] * sc);
  hi.x = pack_bf16(v[8] * sc, v[9] * sc); hi.y = pack_bf16(v[10] * sc, v[11] * sc);
  hi.z = pack_bf16(v[12] * sc, v[13] * sc); hi.w = pack_bf16(v[14] * sc, v[15] * sc);
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
