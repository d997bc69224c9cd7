// Edgewise (Mixture-of-Products) attention backward on tcgen05 / TMEM for token counts up to 200 (ViT-B/16: N = 196).
//
// Execution model of edgewise_tc_large.cuh with one thread per row: one persistent CTA of 256 threads per SM, one
// (batch, head) problem at a time, M=128 row blocks, one thread per row of every fp32 accumulator.  The backward
// touches ~25 N x N maps per problem, far more than fits on chip, so every map that is needed again later is kept
// as a bf16 tile image in a per-CTA scratch region of the caller's workspace (27 slots x 86.5 KB; written by the
// owning threads, read back by the whole CTA with cp.async straight into the operand buffer, overlapped with the
// epilogues wherever the target buffer is free early).  Nothing in
// the scratch is shared between CTAs and nothing survives the launch.
//
// Phases per problem (executable specification: oracle/edgewise_manual.py):
//   1  forward recompute: pass R and pass F of the forward, keeping A_k, the suffix products A_{V-1}..A_k, the
//      prefix products A_0..A_k, F and R.  The softmax statistics, the feature means and the gate factors a, b come from
//      the forward (MopEdgewiseParams::aux): A_k is one exp2 pass, no row / column statistics are formed again;
//   2  mixed-map backward, flash style over 32-column panels: score panels + dA = dY V_1^T panel by MMA, the mix is
//      recomputed in registers, A = exp2(mix - lse) with the row statistics saved by the forward,
//      D = A (dA - delta), delta = dY . (A V_1) with A V_1 in fp32 from the forward; keeps A, the four gate pre-activation gradients, the direct
//      part of dS_k and D g_chain / (F + eps); the row-factor gradients da are fp32 register sums;
//   3  column-factor gradients db = sum_t dG_t^T a (MMA), gate-head parameter partials, feature-mean gradients;
//   4  value gradients dV = (A^T dY) vs_1 + (F^T dY) w vs_V, v_scale partials, chain_value_logit partial;
//   5  chain seeds and sweeps.  The F chain (k = V-1..1):  X' = X A_k^T,  dA_k = P_{k-1}^T X, recorded as bf16 rows in the
//      scratch; the R chain (k = 0..V-2, suffix products) forms its own dA_k, adds the recorded one, and finishes the view:
//      dS_k = A_k (dA_k - rowsum(dA_k A_k)) + direct part + rank-1 feature terms;
//   6  each dS_k goes straight through T = dS_k K, U = dS_k^T Q into fp32 dQ / dK rows (one pair of contractions per view).
#pragma once
#include "edgewise_tc_large.cuh"

#ifdef MOP_PHASE_TIMING
// stamps are kept in local memory and printed after the first problem (a printf per stamp costs ~60k cycles)
#define MOP_TS(name) do { if (threadIdx.x == 0 && blockIdx.x == 0 && g == 0 && ts_n < 32) { ts_v[ts_n] = clock64(); ts_name[ts_n++] = #name; } } while (0)
#else
#define MOP_TS(name) do { } while (0)
#endif

namespace mop {
namespace ewl {

// scratch slots (kBufA bytes each)
constexpr int kSlotA = 0;       // A_k, k = 0..4
constexpr int kSlotSfx = 5;     // suffix product A_{V-1}..A_k for k = 1..3 at kSlotSfx + k - 1
constexpr int kSlotR = 8;       // R = A_{V-1}..A_0
constexpr int kSlotPfx = 9;     // prefix product A_0..A_k for k = 1..3 at kSlotPfx + k - 1
constexpr int kSlotF = 12;      // F = A_0..A_{V-1}
constexpr int kSlotAmix = 13;   // A = softmax(mix)
constexpr int kSlotDG = 14;     // gate pre-activation gradients, 4 maps
constexpr int kSlotDS = 18;     // direct part of dS_k, 5 maps
constexpr int kSlotHf = 23;     // D g_chain / (F + eps)
constexpr int kSlotXN = 24;     // next running product of a sweep (X tile image, kBufX bytes)
constexpr int kSlotAcc = 25;    // fp32 dQ rows, [16][208] float4 (slot 25) and dK rows (slot 26)
constexpr int kSlotDAF = 27;    // dA_k of the F chain (bf16), 5 maps: added to the R chain's dA_k before the softmax backward
constexpr int kBwdSlots = 32;

struct __align__(128) SmemBwd {
  unsigned char X[kBufX];
  unsigned char A[kBufA];
  unsigned char Q[kQt];
  unsigned char K[kKt];
  float colsum[kMaxV + 2][kNmax];  // phase 1: column sums of S_k, log F, log R.  Phase 3 on: rows 0..4 = column part
                                   // of the rank-1 feature gradient of S_k, row 5 / 6 = dkappa of log F / log R
  float cvec[kMaxV][64];
  float vs1[64], vsL[64];
  float zsum[kMaxV][64];          // sum_n (dS_k K)[n,d] Q[n,d]
  float vsum[2][64];              // sum_j dV_1[j,d] V[j,d], sum_j (F^T dY)[j,d] V[j,d]
  float amean[kMaxQ];             // mean over tokens of the row gate factors a[q][.]
  uint64_t bar[2];
  uint32_t tmem_slot;
};
static_assert(sizeof(SmemBwd) <= 232448, "backward shared memory over the 227 KB limit");

// phase-3 staging inside X: dY tile | a (bf16, one masked copy per gate) | da, db rows
constexpr int kOffDy = 0;                          // kQt bytes (R = 200)
constexpr int kAbf = kRA * 16 * 2;                 // 6656: [208 x 16] bf16
constexpr int kOffAbf = kQt;                       // 4 x kAbf
constexpr int kOffDab = kOffAbf + 4 * kAbf;        // float da[16][208], db[16][208]
constexpr int kOffHw = kOffDab + 2 * 16 * kNmax * 4;      // gate-head weights (row | column projection), phase 3
static_assert(kOffHw + 2 * kMaxQ * (2 * kMaxV + 2) * 4 <= kBufX, "phase-3 staging overflows the X buffer");

// column sums over the 32 rows of a warp for 16 columns, added to dst[col] (shared memory)
__device__ __forceinline__ void colsum16_to(float* dst, const float* v, int lane) {
  int col;
  const float cs = warp_colsum16(v, lane, &col);
  if ((lane & 1) == 0) atomicAdd(dst + col, cs);
}

__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

static __global__ void __launch_bounds__(256, 1) edgewise_bwd_kernel(MopEdgewiseParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemBwd& sm = *reinterpret_cast<SmemBwd*>(smem_raw);
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp4 = (tid >> 5) & 3, lane = tid & 31;
  const int N = p.N, V = p.V, r = p.gate_rank, C = 2 * V + 2, dk = p.dk, H = p.H;
  const int KS = (N + 15) >> 4, NN = KS * 16;
  const int cfull = N >> 4;
  const int dks = (dk + 15) >> 4;
  const int row = 128 * wg + t;
  const bool row_ok = row < N;
  const bool blk_on = 128 * wg < N;
  const bool warp_on = 128 * wg + 32 * warp4 < N;
  const float invN = 1.f / (float)N;
  const uint32_t map_bytes = (uint32_t)(2 * KS) * (kRA * 16);
  const uint32_t xmap_bytes = (uint32_t)(2 * KS) * (kRX * 16);
  unsigned char* scratch = reinterpret_cast<unsigned char*>(p.workspace) + (size_t)blockIdx.x * kBwdSlots * kBufA;
  auto slot = [&](int s) -> unsigned char* { return scratch + (size_t)s * kBufA; };

  if (tid < 32) tmem_alloc<512>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_slot;
  const uint32_t tD = tbase + 256u * (uint32_t)wg;
  const uint32_t tl = tD + ((uint32_t)(32 * warp4) << 16);
  uint32_t phase = 0;
  const float w = 1.f / (1.f + __expf(-p.chain_value_logit[0]));
  const float bn = p.beta_not / (float)max(1, V - 1);
  const float sscale = rsqrtf((float)dk);
  const uint32_t sX = smem_u32(sm.X), sA = smem_u32(sm.A), sQ = smem_u32(sm.Q), sK = smem_u32(sm.K);

  auto mma_wait = [&]() { mbar_wait(&sm.bar[wg], phase); phase ^= 1; tc_fence_after(); };
  auto commit = [&]() { mma_commit(&sm.bar[wg]); };
  // generic-proxy shared-memory writes -> visible to the async proxy (MMA operand reads); CTA barrier.  Scratch maps in
  // global memory are written and read back (cp.async, ld.global.cg) through the generic proxy: the barrier orders them.
  auto publish_cta = [&]() { fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after(); };
  auto publish_wg = [&]() { fence_async_smem(); tc_fence_before(); wg_sync(wg); tc_fence_after(); };
  auto sync_cta = [&]() { tc_fence_before(); __syncthreads(); tc_fence_after(); };
  // load of a scratch map into the A buffer (or of an X image into the X buffer) by the whole CTA; every thread waits
  // (measured: one cp.async.bulk of 86 KB took ~25k cycles - 4 B/cycle; 256 threads x cp.async 16 B move it ~10x faster)
  auto load_start = [&](void* dst, const void* src, uint32_t bytes) { cp_async_block(dst, src, bytes); cp_async_commit(); };
  auto load_wait = [&]() { cp_async_wait<0>(); fence_async_smem(); __syncthreads(); };
  // D[rows of this block, NN] = Xtile[rows] * (A buffer, K index = its rows)          (forward chain step)
  auto mma_x_a = [&]() {
    const uint32_t id = idesc_bf16(128, NN, 0, 1);
    for (int ks = 0; ks < KS; ++ks)
      mma_ss(tD, desc_kmajor(sX + 128 * wg * 16, kRX, 16 * ks), desc_mnmajor(sA, kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
  };
  // D[rows, m] = sum_j Xtile[rows, j] * Abuf[m, j]                                      (X A_k^T)
  auto mma_x_at = [&]() {
    const uint32_t id = idesc_bf16(128, NN, 0, 0);
    for (int ks = 0; ks < KS; ++ks)
      mma_ss(tD, desc_kmajor(sX + 128 * wg * 16, kRX, 16 * ks), desc_kmajor(sA, kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
  };
  // D[m in this block, j] = sum_i Abuf[i, m] * Xtile[i, j]                              (P^T X)
  auto mma_at_x = [&]() {
    const uint32_t id = idesc_bf16(128, NN, 1, 1);
    for (int ks = 0; ks < KS; ++ks)
      mma_ss(tD, desc_mnmajor(sA + (16 * wg) * (kRA * 16), kRA, 16 * ks), desc_mnmajor(sX, kRX, 16 * ks), id, ks > 0 ? 1u : 0u);
  };
  // D[m in this block, 0..ncols) at column dcol = sum_i Abuf[i, m] * tile[i, :]         (map^T times a [tokens x n] tile)
  auto mma_at_tile = [&](uint32_t dcol, uint32_t tile, uint32_t R, uint32_t ncols, bool acc) {
    const uint32_t id = idesc_bf16(128, ncols, 1, 1);
    for (int ks = 0; ks < KS; ++ks)
      mma_ss(tD + dcol, desc_mnmajor(sA + (16 * wg) * (kRA * 16), kRA, 16 * ks), desc_mnmajor(tile, R, 16 * ks), id, (acc || ks > 0) ? 1u : 0u);
  };
  // D[rows of this block, 0..64) at column dcol = Abuf[rows, :] * tile (tile rows = tokens)
  auto mma_a_tile = [&](uint32_t dcol, uint32_t tile, uint32_t R) {
    const uint32_t id = idesc_bf16(128, 64, 0, 1);
    for (int ks = 0; ks < KS; ++ks)
      mma_ss(tD + dcol, desc_kmajor(sA + 128 * wg * 16, kRA, 16 * ks), desc_mnmajor(tile, R, 16 * ks), id, ks > 0 ? 1u : 0u);
  };

  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p.qkv);
  __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(p.dqkv);
  const size_t hd = (size_t)H * dk;
  const int G = p.B * H;
#ifdef MOP_PHASE_TIMING
  long long ts_v[32];
  const char* ts_name[32];
  int ts_n = 0;
#endif
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int pb = g / H, ph = g % H;
    auto in_row = [&](int n) { return qkv + (((size_t)pb * N + n) * 3) * hd + (size_t)ph * dk; };
    auto out_row = [&](int n) { return dqkv + (((size_t)pb * N + n) * 3) * hd + (size_t)ph * dk; };
    auto tok_row = [&](const void* base, int n) { return reinterpret_cast<const __nv_bfloat16*>(base) + (((size_t)pb * N + n) * H + ph) * dk; };
    // row image of a scratch map (R = 208 layout): chunk c of this thread's row
    auto map_chunk = [&](int s, int c) -> unsigned char* { return slot(s) + (size_t)c * (kRA * 16) + row * 16; };
    // what the forward left for this problem: per-view softmax statistics, feature means, gate factors ([field][208 tokens])
    const float* aux = p.aux + (size_t)g * kAuxLFloats;
    // =================================================================================================
    // stage 0
    // =================================================================================================
    for (int idx = tid; idx < V * 64; idx += 256) {
      const int i = idx >> 6, d = idx & 63;
      float c = 0.f;
      if (d < dk) c = sscale * p.q_scale[((size_t)i * H + ph) * dk + d] * p.k_scale[((size_t)i * H + ph) * dk + d];
      sm.cvec[i][d] = c;
      sm.zsum[i][d] = 0.f;
    }
    if (tid < 64) {
      const int d = tid;
      float a = 0.f, b = 0.f;
      if (d < dk) { a = p.v_scale[((size_t)0 * H + ph) * dk + d]; b = p.v_scale[((size_t)(V - 1) * H + ph) * dk + d]; }
      sm.vs1[d] = a;
      sm.vsL[d] = w * b;
      sm.vsum[0][d] = 0.f;
      sm.vsum[1][d] = 0.f;
      if (d < kMaxQ) sm.amean[d] = 0.f;
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
    // accumulator -> bf16 row of X and / or of a scratch map
    auto chain_epilogue = [&](bool store_x, int gslot) {
      for (int c = 0; c < KS; ++c) {
        float v[16];
        tmem_ld_32x32b_x16(tl + 16 * c, v);
        tmem_ld_wait();
        uint4 lo, hi;
        pack16(v, 1.f, lo, hi);
        if (!row_ok) lo = hi = make_uint4(0, 0, 0, 0);
        if (store_x && row < kRX) {
          *reinterpret_cast<uint4*>(sm.X + (2 * c) * (kRX * 16) + row * 16) = lo;
          *reinterpret_cast<uint4*>(sm.X + (2 * c + 1) * (kRX * 16) + row * 16) = hi;
        }
        if (gslot >= 0 && row < kRA) {
          *reinterpret_cast<uint4*>(map_chunk(gslot, 2 * c)) = lo;
          *reinterpret_cast<uint4*>(map_chunk(gslot, 2 * c + 1)) = hi;
        }
      }
    };
    MOP_TS(T0);
    // =================================================================================================
    // phase 1a: pass R (views V-1 .. 0)
    // =================================================================================================
    for (int idx = 0; idx < V; ++idx) {
      const int k = V - 1 - idx;
      const bool first = idx == 0, last = idx == V - 1;
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
          commit();
        }
        mma_wait();
      }
      if (warp_on) {
        // A_k = exp2(S_k log2e - mb) / l with the row statistics of the forward: one pass, bit-identical to the forward's A_k
        const float mb = row_ok ? aux[kAuxLStats + (2 * k) * kNmax + row] : 0.f;
        const float inv_l = row_ok ? aux[kAuxLStats + (2 * k + 1) * kNmax + row] : 0.f;
        for (int c = 0; c < KS; ++c) {
          float v[16];
          tmem_ld_32x32b_x16(tl + 16 * c, v);
          tmem_ld_wait();
          if (c < cfull) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fast_exp2(fmaf(v[e], kLog2e, -mb));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = (16 * c + e < N) ? fast_exp2(fmaf(v[e], kLog2e, -mb)) : 0.f;
          }
          uint4 lo, hi;
          pack16(v, inv_l, lo, hi);
          if (!row_ok) lo = hi = make_uint4(0, 0, 0, 0);
          if (row < kRA) {
            *reinterpret_cast<uint4*>(sm.A + (2 * c) * (kRA * 16) + row * 16) = lo;
            *reinterpret_cast<uint4*>(sm.A + (2 * c + 1) * (kRA * 16) + row * 16) = hi;
            *reinterpret_cast<uint4*>(map_chunk(kSlotA + k, 2 * c)) = lo;       // A_k is needed again by pass F and the sweeps
            *reinterpret_cast<uint4*>(map_chunk(kSlotA + k, 2 * c + 1)) = hi;
          }
          if (first && row < kRX) {
            *reinterpret_cast<uint4*>(sm.X + (2 * c) * (kRX * 16) + row * 16) = lo;
            *reinterpret_cast<uint4*>(sm.X + (2 * c + 1) * (kRX * 16) + row * 16) = hi;
          }
        }
      } else if (row < kRA) {
        for (int c = 0; c < 2 * KS; ++c) *reinterpret_cast<uint4*>(sm.A + c * (kRA * 16) + row * 16) = make_uint4(0, 0, 0, 0);
      }
      publish_cta();
      if (first) continue;
      if (blk_on) {
        if (t == 0) { mma_x_a(); commit(); }
        mma_wait();
      }
      if (warp_on) chain_epilogue(!last, last ? kSlotR : kSlotSfx + k - 1);
      if (last && row < kRX) {
        for (int c = 0; c < 2 * KS; ++c)
          *reinterpret_cast<uint4*>(sm.X + c * (kRX * 16) + row * 16) =
              row < kRA ? *reinterpret_cast<const uint4*>(sm.A + c * (kRA * 16) + row * 16) : make_uint4(0, 0, 0, 0);
      }
    }
    MOP_TS(T1);
    // =================================================================================================
    // phase 1b: pass F (views 1 .. V-1)
    // =================================================================================================
    publish_cta();
    load_start(sm.A, slot(kSlotA + 1), map_bytes);
    for (int k = 1; k < V; ++k) {
      const bool last = k == V - 1;
      load_wait();
      if (blk_on) {
        if (t == 0) { mma_x_a(); commit(); }
        mma_wait();
      }
      if (!last) {
        sync_cta();
        load_start(sm.A, slot(kSlotA + k + 1), map_bytes);
      }
      if (warp_on) chain_epilogue(true, last ? kSlotF : kSlotPfx + k - 1);
      if (!last) publish_cta();
    }
    publish_cta();   // F in X and in its slot; colsum final; A and K buffers free
    MOP_TS(T2);
    // =================================================================================================
    // gate factors; delta = dY . (y - F w V_V)
    // =================================================================================================
    float* bfac = reinterpret_cast<float*>(sm.A + kOffBfac);
    unsigned char* Vt = sm.A + kOffVt;
    unsigned char* KsP = sm.A + kOffKsP + wg * (kMaxV * kKsP);
    auto load_values = [&](unsigned char* dst, const float* vscale) {
      if (tid < kRA) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4 vv = make_uint4(0, 0, 0, 0);
          if (tid < N && ch * 8 < dk) vv = scale_chunk(*reinterpret_cast<const uint4*>(in_row(tid) + 2 * hd + ch * 8), &vscale[ch * 8]);
          *reinterpret_cast<uint4*>(dst + ch * (kRA * 16) + tid * 16) = vv;
        }
      }
    };
    // gate factors a (registers) / b (shared) and the feature means of this token, as the forward computed them
    float afac[kMaxQ];
    float fr[2 * kMaxV + 2], fc[2 * kMaxV + 2];
    {
#pragma unroll
      for (int c = 0; c < kMaxV; ++c) {
        const float rh = (c < V && row_ok) ? aux[kAuxLFeat + c * kNmax + row] : 0.f;
        const float kp = (c < V && row_ok) ? aux[kAuxLFeat + (kMaxV + c) * kNmax + row] : 0.f;
        fr[c] = rh; fr[kMaxV + c] = kp;
        fc[c] = kp; fc[kMaxV + c] = rh;
      }
      fr[2 * kMaxV] = row_ok ? aux[kAuxLFeat + (2 * kMaxV) * kNmax + row] : 0.f;
      fr[2 * kMaxV + 1] = row_ok ? aux[kAuxLFeat + (2 * kMaxV + 1) * kNmax + row] : 0.f;
      fc[2 * kMaxV] = row_ok ? aux[kAuxLFeat + (2 * kMaxV + 2) * kNmax + row] : 0.f;
      fc[2 * kMaxV + 1] = row_ok ? aux[kAuxLFeat + (2 * kMaxV + 3) * kNmax + row] : 0.f;
      float bq[kMaxQ];
#pragma unroll
      for (int qq = 0; qq < kMaxQ; ++qq) {   // every load is issued before the first shuffle / atomic below
        const bool on = (qq & 3) < r && row_ok;
        afac[qq] = on ? aux[kAuxLA + qq * kNmax + row] : 0.f;
        bq[qq] = on ? aux[kAuxLB + qq * kNmax + row] : 0.f;
      }
#pragma unroll
      for (int qq = 0; qq < kMaxQ; ++qq) {
        if (row < kNmax) bfac[row * 16 + qq] = bq[qq];
        const float asum = warp_sum(afac[qq]);   // a = 0 for padded rows
        if (lane == 0 && asum != 0.f) atomicAdd(&sm.amean[qq], asum);
      }
    }
    MOP_TS(T2a);
    // exact (fp32) column sums of the four gate pre-activation gradients, accumulated in phase 2: [4][208] in the K region
    float* cs_s = reinterpret_cast<float*>(sm.K);
    for (int idx = tid; idx < 4 * kNmax; idx += 256) cs_s[idx] = 0.f;
    // delta = rowsum(dA . A) = dY . (A V_1), with A V_1 in fp32 from the forward: eight lanes per row (32 B of y_base and 16 B
    // of dY per lane; four rows per warp instruction, seven passes per warp, all loads of a pass group in flight) - one thread
    // per 256-byte row costs 32 sectors per load instruction and 24 dependent round trips
    float* delta_s = cs_s + 4 * kNmax;   // [208]
    {
      const int wrp = tid >> 5, sub = lane >> 3, l8 = lane & 7;
      for (int ps0 = 0; ps0 < 7; ps0 += 4) {
        float4 ya[4], yb4[4];
        uint4 da4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ps = ps0 + i, rr = 28 * wrp + 4 * ps + sub;   // 8 warps x 28 rows >= 208
          const bool ok = ps < 7 && rr < N && 8 * l8 < dk;
          const size_t o = (((size_t)pb * N + (ok ? rr : 0)) * H + ph) * dk + 8 * l8;
          ya[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.y_base + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
          yb4[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.y_base + o + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          da4[i] = ok ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + o)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ps = ps0 + i, rr = 28 * wrp + 4 * ps + sub;
          float dv[8];
          unpack8(da4[i], dv);
          float d = fmaf(dv[0], ya[i].x, fmaf(dv[1], ya[i].y, fmaf(dv[2], ya[i].z, dv[3] * ya[i].w)));
          d = fmaf(dv[4], yb4[i].x, fmaf(dv[5], yb4[i].y, fmaf(dv[6], yb4[i].z, fmaf(dv[7], yb4[i].w, d))));
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if (l8 == 0 && ps < 7 && rr < kNmax) delta_s[rr] = d;
        }
      }
    }
    publish_cta();
    float delta = 0.f, lse2 = 0.f, inv_lmix = 0.f;
    if (row_ok) {
      delta = delta_s[row];
      const float2 st2 = *reinterpret_cast<const float2*>(p.row_stats + (((size_t)pb * H + ph) * N + row) * 2);
      lse2 = st2.x;          // integer reference exponent (base 2) of this row of the mixed map
      inv_lmix = 1.f / st2.y;   // 1 / sum of the bf16-rounded exp2(mix - reference)
    }
    MOP_TS(T2b);
    sync_cta();   // X (F) and the value tile are free
    // dY tile (A operand of dA = dY V_1^T, B operand of the value-gradient MMAs), V_1 tile
    unsigned char* dYt = sm.X + kOffDy;
    if (tid < kRX) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tid < N && ch * 8 < dk) v = *reinterpret_cast<const uint4*>(tok_row(p.dy, tid) + ch * 8);
        *reinterpret_cast<uint4*>(dYt + ch * (kRX * 16) + tid * 16) = v;
      }
    }
    load_values(Vt, sm.vs1);
    publish_cta();
    MOP_TS(T3);
    // =================================================================================================
    // phase 2: mixed-map backward over 32-column panels
    // =================================================================================================
    float da[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) da[q] = 0.f;
    const uint32_t tS = tD, tdA = tD + 160;
    const uint32_t tlS = tl, tldA = tl + 160;
    if (blk_on) {
      const int npanels = (N + kPanel - 1) / kPanel;
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
      uint4 fnext = row < kRA ? ldcg16(map_chunk(kSlotF, 0)) : make_uint4(0, 0, 0, 0);
      for (int pn = 0; pn < npanels; ++pn) {
        const int j0 = pn * kPanel;
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
          for (int ks = 0; ks < dks; ++ks)   // dA panel = dY V_1[panel]^T
            mma_ss(tdA, desc_kmajor(smem_u32(dYt) + 128 * wg * 16, kRX, 16 * ks), desc_kmajor(smem_u32(Vt) + j0 * 16, kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
          commit();
        }
        if (pn + 1 < npanels) fetch_panel(j0 + kPanel);
        mma_wait();
        if (warp_on) {
#pragma unroll 1
          for (int sub = 0; sub < kPanel / 8; ++sub) {
            const int jc = j0 + 8 * sub;
            if (jc >= NN) break;
            float sv[kMaxV][8], dav[8];
#pragma unroll
            for (int i = 0; i < kMaxV; ++i)
              if (i < V) tmem_ld_32x32b_x8(tlS + 32 * i + 8 * sub, sv[i]);
            tmem_ld_32x32b_x8(tldA + 8 * sub, dav);
            tmem_ld_wait();
            if (!row_ok) {   // padded rows hold whatever the over-read operand rows produced: make them exact zeros
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                dav[e] = 0.f;
#pragma unroll
                for (int i = 0; i < kMaxV; ++i) sv[i][e] = 0.f;
              }
            }
            float fv[8];
            unpack8(fnext, fv);
            if (jc + 8 < NN && row < kRA) fnext = ldcg16(map_chunk(kSlotF, (jc + 8) >> 3));   // next chunk of this row of F
            float o_am[8], o_hf[8], o_dg[4][8], o_ds[kMaxV][8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = jc + e;
              const bool ok = j < N && row_ok;
              float s0 = sv[0][e], sum = s0, mxv = s0;
#pragma unroll
              for (int i = 1; i < kMaxV; ++i)
                if (i < V) { sum += sv[i][e]; mxv = fmaxf(mxv, sv[i][e]); }
              float ex[kMaxV], se = 0.f;
#pragma unroll
              for (int i = 0; i < kMaxV; ++i)
                if (i < V) { ex[i] = fast_exp2((sv[i][e] - mxv) * kLog2e); se += ex[i]; }
              const float inv_se = fast_rcp(se);
              const float lse = mxv + kLn2 * fast_log2(se);
              const float U = sum - s0, O = lse - s0;
              const float fe = fv[e] + p.eps, lf = kLn2 * fast_log2(fe);
              float z[4], bq[kMaxQ];
              const float4* bp = reinterpret_cast<const float4*>(bfac + (j < kNmax ? j : kNmax - 1) * 16);
#pragma unroll
              for (int tg = 0; tg < 4; ++tg) {
                const float4 b4 = bp[tg];
                bq[4 * tg] = b4.x; bq[4 * tg + 1] = b4.y; bq[4 * tg + 2] = b4.z; bq[4 * tg + 3] = b4.w;
                z[tg] = fmaf(afac[4 * tg], b4.x, fmaf(afac[4 * tg + 1], b4.y, fmaf(afac[4 * tg + 2], b4.z, afac[4 * tg + 3] * b4.w)));
              }
              const float g0 = fast_sigmoid(z[0]), g1 = fast_sigmoid(z[1]), g2 = fast_sigmoid(z[2]), g3 = fast_sigmoid(z[3]);
              const float mixv = s0 + g0 * U + g1 * O - g2 * bn * U + g3 * lf;
              const float am = ok ? __bfloat162float(__float2bfloat16_rn(fast_exp2(fmaf(mixv, kLog2e, -lse2)))) * inv_lmix : 0.f;
              const float d = ok ? am * (dav[e] - delta) : 0.f;
              o_am[e] = am;
              const float dg0 = d * U * g0 * (1.f - g0), dg1 = d * O * g1 * (1.f - g1), dg2 = -bn * d * U * g2 * (1.f - g2), dg3 = d * lf * g3 * (1.f - g3);
              o_dg[0][e] = dg0; o_dg[1][e] = dg1; o_dg[2][e] = dg2; o_dg[3][e] = dg3;
#pragma unroll
              for (int q = 0; q < kMaxQ; ++q) {
                const float dgq = (q >> 2) == 0 ? dg0 : (q >> 2) == 1 ? dg1 : (q >> 2) == 2 ? dg2 : dg3;
                da[q] = fmaf(dgq, bq[q], da[q]);
              }
              o_hf[e] = d * g3 * fast_rcp(fe);
              const float e1 = d * g1, e0 = d * (g0 - g2 * bn);
#pragma unroll
              for (int i = 0; i < kMaxV; ++i)
                if (i < V) {
                  const float pi = ex[i] * inv_se;
                  o_ds[i][e] = (i == 0) ? (d - e1 + e1 * pi) : (e0 + e1 * pi);
                }
            }
            {
              // column sums over this warp's rows: gates (0,1) and (2,3) share one 16-column butterfly each
              float pr[16];
#pragma unroll
              for (int e = 0; e < 8; ++e) { pr[e] = o_dg[0][e]; pr[8 + e] = o_dg[1][e]; }
              int col;
              float csum = warp_colsum16(pr, lane, &col);
              if ((lane & 1) == 0) atomicAdd(&cs_s[(col >> 3) * kNmax + jc + (col & 7)], csum);
#pragma unroll
              for (int e = 0; e < 8; ++e) { pr[e] = o_dg[2][e]; pr[8 + e] = o_dg[3][e]; }
              csum = warp_colsum16(pr, lane, &col);
              if ((lane & 1) == 0) atomicAdd(&cs_s[(2 + (col >> 3)) * kNmax + jc + (col & 7)], csum);
            }
            if (row < kRA) {
              const int ch = jc >> 3;
              *reinterpret_cast<uint4*>(map_chunk(kSlotAmix, ch)) = pack8(o_am);
              *reinterpret_cast<uint4*>(map_chunk(kSlotHf, ch)) = pack8(o_hf);
#pragma unroll
              for (int tg = 0; tg < 4; ++tg) *reinterpret_cast<uint4*>(map_chunk(kSlotDG + tg, ch)) = pack8(o_dg[tg]);
#pragma unroll
              for (int i = 0; i < kMaxV; ++i)
                if (i < V) *reinterpret_cast<uint4*>(map_chunk(kSlotDS + i, ch)) = pack8(o_ds[i]);
            }
          }
        }
      }
    }
    sync_cta();   // phase-2 MMAs complete: A buffer (value tile, key panels, b factors) and the dY tile region are reusable
    MOP_TS(T4);
    // =================================================================================================
    // phase 3: db = sum_t dG_t^T a (MMA), head-parameter partials, feature-mean gradients
    // =================================================================================================
    float* da_s = reinterpret_cast<float*>(sm.X + kOffDab);          // [16][208]
    float* db_s = da_s + 16 * kNmax;                                  // [16][208]
    float* fr_s = reinterpret_cast<float*>(sm.K);                     // [12][208] row-projection features
    float* fc_s = fr_s + (2 * kMaxV + 2) * kNmax;                     // [12][208] column-projection features
    static_assert(2 * (2 * kMaxV + 2) * kNmax * 4 <= kKt, "feature staging overflows the K tile");
    float csr[4];   // column sums of dG_t at column `row`
#pragma unroll
    for (int tg = 0; tg < 4; ++tg) csr[tg] = row < kNmax ? cs_s[tg * kNmax + row] : 0.f;
    __syncthreads();   // cs read before the K region is reused for the feature staging
    if (row < kNmax) {
#pragma unroll
      for (int q = 0; q < kMaxQ; ++q) {
        da_s[q * kNmax + row] = row_ok ? da[q] : 0.f;
        // bf16 image of a, one copy per gate with the other gates' columns zeroed: [208 rows][16 cols], R = 208
        // centred: db = mean(a) * colsum(dG_t) (exact, fp32) + dG_t^T (a - mean(a)) (MMA); the column sums of dG_t cancel
        // heavily once summed over columns, which bf16-rounded operands would not reproduce
        const __nv_bfloat16 av = __float2bfloat16_rn(row_ok ? afac[q] - sm.amean[q] * invN : 0.f);
#pragma unroll
        for (int tg = 0; tg < 4; ++tg)
          *reinterpret_cast<__nv_bfloat16*>(sm.X + kOffAbf + tg * kAbf + tile_off(kRA, row, q)) = (q >> 2) == tg ? av : __float2bfloat16_rn(0.f);
      }
#pragma unroll
      for (int c = 0; c < 2 * kMaxV + 2; ++c) { fr_s[c * kNmax + row] = row_ok ? fr[c] : 0.f; fc_s[c * kNmax + row] = row_ok ? fc[c] : 0.f; }
    }
    // the projection weights for the feature-mean gradients below (the L1 next to 227 KB of shared memory is a few KB: read
    // with __ldg inside the loops they came from L2 one dependent load at a time)
    float* hw_s = reinterpret_cast<float*>(sm.X + kOffHw);
    {
      const int nW = 4 * r * C;
      for (int idx = tid; idx < 2 * nW; idx += 256) hw_s[idx] = idx < nW ? p.row_w[idx] : p.col_w[idx - nW];
    }
    publish_cta();
    for (int tg = 0; tg < 4; ++tg) {
      load_start(sm.A, slot(kSlotDG + tg), map_bytes);
      load_wait();
      if (blk_on) {
        if (t == 0) { mma_at_tile(0, sX + kOffAbf + tg * kAbf, kRA, 16, tg > 0); commit(); }
        mma_wait();
      }
      sync_cta();   // both warpgroups' MMAs have read the map
    }
    load_start(sm.A, slot(kSlotAmix), map_bytes);   // for phase 4: lands during the vector work below
    MOP_TS(T4a);
    float db[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) db[q] = 0.f;
    if (warp_on) {
      tmem_ld_32x32b_x16(tl, db);
      tmem_ld_wait();
    }
    if (row < kNmax) {
#pragma unroll
      for (int q = 0; q < kMaxQ; ++q) { db[q] = row_ok ? fmaf(sm.amean[q] * invN, csr[q >> 2], db[q]) : 0.f; db_s[q * kNmax + row] = db[q]; }
    }
    // feature-mean gradients of this thread's token (as a row i and as a column j)
    float drho[2 * kMaxV + 2], dkap[2 * kMaxV + 2];
#pragma unroll
    for (int c = 0; c < 2 * kMaxV + 2; ++c) { drho[c] = 0.f; dkap[c] = 0.f; }
    if (row_ok) {
#pragma unroll
      for (int qq = 0; qq < kMaxQ; ++qq) {
        const int tg = qq >> 2, kk = qq & 3, q = tg * r + kk;
        if (kk < r) {
#pragma unroll
          for (int c = 0; c < kMaxV; ++c)
            if (c < V) {
              drho[c] = fmaf(hw_s[q * C + c], da[qq], drho[c]);
              drho[kMaxV + c] = fmaf(hw_s[q * C + V + c], da[qq], drho[kMaxV + c]);
              dkap[c] = fmaf(hw_s[4 * r * C + q * C + c], db[qq], dkap[c]);
              dkap[kMaxV + c] = fmaf(hw_s[4 * r * C + q * C + V + c], db[qq], dkap[kMaxV + c]);
            }
          drho[2 * kMaxV] = fmaf(hw_s[q * C + 2 * V], da[qq], drho[2 * kMaxV]);
          drho[2 * kMaxV + 1] = fmaf(hw_s[q * C + 2 * V + 1], da[qq], drho[2 * kMaxV + 1]);
          dkap[2 * kMaxV] = fmaf(hw_s[4 * r * C + q * C + 2 * V], db[qq], dkap[2 * kMaxV]);
          dkap[2 * kMaxV + 1] = fmaf(hw_s[4 * r * C + q * C + 2 * V + 1], db[qq], dkap[2 * kMaxV + 1]);
        }
      }
#pragma unroll
      for (int c = 0; c < 2 * kMaxV + 2; ++c) { drho[c] *= invN; dkap[c] *= invN; }
    }
    float* cprime_s = &sm.colsum[0][0];
    float* dkapF_s = &sm.colsum[kMaxV][0];
    float* dkapR_s = &sm.colsum[kMaxV + 1][0];
    MOP_TS(T4b);
    // dS_k[i,j] += rterm_k[i] + cprime_k[j]:  channel k is S_k, channel V+k is S_k^T (row/column roles swapped)
    float rterm[kMaxV];
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
      rterm[k] = drho[k] + dkap[kMaxV + k];
      if (row < kNmax) cprime_s[k * kNmax + row] = dkap[k] + drho[kMaxV + k];
    }
    const float drhoF = drho[2 * kMaxV], drhoR = drho[2 * kMaxV + 1];
    if (row < kNmax) { dkapF_s[row] = dkap[2 * kMaxV]; dkapR_s[row] = dkap[2 * kMaxV + 1]; }
    __syncthreads();
    {
      // gate-head parameter partials: row_w [4r][C], row_b [4r], col_w, col_b
      const int nW = 4 * r * C, nP = nW + 4 * r;
      float* dh = p.dhead_part + (size_t)g * 2 * nP;
      for (int idx = tid; idx < 2 * nP; idx += 256) {
        const int half = idx / nP, rem = idx % nP;
        const float* dv = half ? db_s : da_s;
        const float* ft = half ? fc_s : fr_s;
        float s = 0.f;
        if (rem < nW) {
          const int q = rem / C, c = rem % C, qq = 4 * (q / r) + (q % r);
          const int cc = c < V ? c : c < 2 * V ? kMaxV + (c - V) : 2 * kMaxV + (c - 2 * V);
          for (int i = 0; i < N; ++i) s = fmaf(dv[qq * kNmax + i], ft[cc * kNmax + i], s);
        } else {
          const int q = rem - nW, qq = 4 * (q / r) + (q % r);
          for (int i = 0; i < N; ++i) s += dv[qq * kNmax + i];
        }
        dh[idx] = s;
      }
    }
    sync_cta();
    MOP_TS(T5);
    // =================================================================================================
    // phase 4: dV = (A^T dY) vs_1 + (F^T dY) w vs_V ; v_scale partials ; chain_value_logit partial
    // =================================================================================================
    load_wait();   // A = softmax(mix), started at the end of phase 3
    if (blk_on) {
      if (t == 0) { mma_at_tile(0, smem_u32(dYt), kRX, 64, false); commit(); }
      mma_wait();
    }
    sync_cta();
    load_start(sm.A, slot(kSlotF), map_bytes);
    load_wait();
    if (blk_on) {
      if (t == 0) { mma_at_tile(64, smem_u32(dYt), kRX, 64, false); commit(); }
      mma_wait();
      if (warp_on) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float d1[16], dl[16], p1[16], pl[16];
          tmem_ld_32x32b_x16(tl + 16 * c, d1);
          tmem_ld_32x32b_x16(tl + 64 + 16 * c, dl);
          tmem_ld_wait();
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const int d0 = 16 * c + 8 * h8;
            float vv[8], o[8];
            unpack8((row_ok && d0 < dk) ? *reinterpret_cast<const uint4*>(in_row(row) + 2 * hd + d0) : make_uint4(0, 0, 0, 0), vv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float a1 = d1[8 * h8 + e], al = dl[8 * h8 + e];
              o[e] = a1 * sm.vs1[d0 + e] + al * sm.vsL[d0 + e];
              p1[8 * h8 + e] = row_ok ? a1 * vv[e] : 0.f;
              pl[8 * h8 + e] = row_ok ? al * vv[e] : 0.f;
            }
            if (row_ok && d0 < dk) *reinterpret_cast<uint4*>(out_row(row) + 2 * hd + d0) = pack8(o);
          }
          colsum16_to(&sm.vsum[0][16 * c], p1, lane);
          colsum16_to(&sm.vsum[1][16 * c], pl, lane);
        }
      }
    }
    sync_cta();
    if (tid < 32) {
      float s = 0.f;
      for (int d = tid; d < dk; d += 32) s = fmaf(sm.vsL[d], sm.vsum[1][d], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tid == 0) p.dlogit_part[g] = (1.f - w) * s;
    }
    if (p.dscale_part) {
      float* ds = p.dscale_part + (size_t)g * 3 * V * dk + (size_t)2 * V * dk;
      for (int idx = tid; idx < V * dk; idx += 256) {
        const int k = idx / dk, d = idx % dk;
        float val = 0.f;
        if (k == 0) val = sm.vsum[0][d];
        if (k == V - 1) val += w * sm.vsum[1][d];
        ds[idx] = val;
      }
    }
    MOP_TS(T6);
    // =================================================================================================
    // phase 5 / 6: chain seeds, sweeps and contributions
    // =================================================================================================
    // value tile w V_V (B operand of the seed MMA) in the K region; later the unscaled key tile lives there
    load_values(sm.K, sm.vsL);
    publish_cta();
    if (blk_on) {
      if (t == 0) {   // dY (w V_V)^T -> [rows, NN]
        const uint32_t id = idesc_bf16(128, NN, 0, 0);
        for (int ks = 0; ks < dks; ++ks)
          mma_ss(tD, desc_kmajor(smem_u32(dYt) + 128 * wg * 16, kRX, 16 * ks), desc_kmajor(sK, kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
        commit();
      }
      mma_wait();
    }
    sync_cta();   // dY tile and value tile are dead: X may be overwritten, K region gets the unscaled keys
    if (tid < kRA) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 kk = make_uint4(0, 0, 0, 0);
        if (tid < N && ch * 8 < dk) kk = *reinterpret_cast<const uint4*>(in_row(tid) + hd + ch * 8);
        *reinterpret_cast<uint4*>(sm.K + ch * (kRA * 16) + tid * 16) = kk;
      }
    }
    // seed of the F sweep: X = D g_chain/(F+eps) + dY (w V_V)^T + (drho_F[i] + dkap_F[j]) / (F + eps)
    if (warp_on) {
      uint4 pf[2][2];   // [F | Hf][half] of the next 16 columns
      auto fetch_fh = [&](int c) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          pf[0][hh] = (row < kRA && c < KS) ? ldcg16(map_chunk(kSlotF, 2 * c + hh)) : make_uint4(0, 0, 0, 0);
          pf[1][hh] = (row < kRA && c < KS) ? ldcg16(map_chunk(kSlotHf, 2 * c + hh)) : make_uint4(0, 0, 0, 0);
        }
      };
      fetch_fh(0);
      for (int c = 0; c < KS; ++c) {
        float v[16], f8[8], h8v[8];
        tmem_ld_32x32b_x16(tl + 16 * c, v);
        tmem_ld_wait();
        const uint4 cf0 = pf[0][0], cf1 = pf[0][1], ch0 = pf[1][0], ch1 = pf[1][1];
        fetch_fh(c + 1);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          unpack8(hh ? cf1 : cf0, f8);
          unpack8(hh ? ch1 : ch0, h8v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = 16 * c + 8 * hh + e;
            const float x = v[8 * hh + e] + h8v[e] + (drhoF + dkapF_s[j < kNmax ? j : 0]) * fast_rcp(f8[e] + p.eps);
            v[8 * hh + e] = (row_ok && j < N) ? x : 0.f;
          }
        }
        if (row < kRX) {
          uint4 lo, hi;
          pack16(v, 1.f, lo, hi);
          *reinterpret_cast<uint4*>(sm.X + (2 * c) * (kRX * 16) + row * 16) = lo;
          *reinterpret_cast<uint4*>(sm.X + (2 * c + 1) * (kRX * 16) + row * 16) = hi;
        }
      }
    }
    bool acc_started = false;   // dQ / dK rows in the scratch hold valid partial sums
    // fp32 dQ / dK rows, stored [16 groups of 4 features][208 rows] so that the one-thread-per-row accesses coalesce
    float4* dq_acc = reinterpret_cast<float4*>(slot(kSlotAcc)) + row;
    float4* dk_acc = reinterpret_cast<float4*>(slot(kSlotAcc + 1)) + row;
    // C (bf16, all rows, in the A buffer) is one additive part of dS_k:  T = C K -> dQ, scale sums ; U = C^T Q -> dK
    // next_slot >= 0: once both warpgroups' MMAs have read C, that scratch map starts to load into the A buffer, so that its
    // latency hides behind the epilogue (the caller waits for it with load_wait before the next use of the A buffer).
    auto contribute = [&](int k, int next_slot) {
      publish_cta();   // C rows written by their owners
      if (k == 1) MOP_TS(c_pub);
      if (blk_on) {
        if (t == 0) {
          mma_a_tile(0, sK, kRA);
          mma_at_tile(64, sQ, kRX, 64, false);
          commit();
        }
        mma_wait();
      }
      if (k == 1) MOP_TS(c_mma);
      if (next_slot >= 0) {
        sync_cta();
        load_start(sm.A, slot(next_slot), map_bytes);
      }
      if (k == 1) MOP_TS(c_sync);
      if (blk_on) {
        if (warp_on) {
          // T part: dQ rows (fp32 partial sums live in the scratch; all loads are issued before the first use)
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            float4* acc = part ? dk_acc : dq_acc;
            float4 av[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (acc_started && row_ok) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (4 * i < dk) av[i] = __ldcg(acc + i * kNmax);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float tv[16];
              tmem_ld_32x32b_x16(tl + 64 * part + 16 * c, tv);
              tmem_ld_wait();
              const float* cv = &sm.cvec[k][16 * c];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                av[4 * c + i].x = fmaf(tv[4 * i + 0], cv[4 * i + 0], av[4 * c + i].x);
                av[4 * c + i].y = fmaf(tv[4 * i + 1], cv[4 * i + 1], av[4 * c + i].y);
                av[4 * c + i].z = fmaf(tv[4 * i + 2], cv[4 * i + 2], av[4 * c + i].z);
                av[4 * c + i].w = fmaf(tv[4 * i + 3], cv[4 * i + 3], av[4 * c + i].w);
              }
              if (part == 0) {
                float zq[16], qv[8];
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                  unpack8(row < kRX ? *reinterpret_cast<const uint4*>(sm.Q + (2 * c + h8) * (kRX * 16) + row * 16) : make_uint4(0, 0, 0, 0), qv);
#pragma unroll
                  for (int e = 0; e < 8; ++e) zq[8 * h8 + e] = row_ok ? tv[8 * h8 + e] * qv[e] : 0.f;
                }
                colsum16_to(&sm.zsum[k][16 * c], zq, lane);
              }
            }
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (4 * i < dk) acc[i * kNmax] = av[i];
            }
            if (k == 1) MOP_TS(c_part);
          }
        }
      }
      acc_started = true;
      sync_cta();   // the A buffer may be refilled
    };
    // dS_k rows, complete: x = dA_k of the R chain (accumulator, or the X tile for the final link) + dA_k of the F chain
    // (scratch, bf16);  C[m,:] = Ak[m,:] * (x[m,:] - sum_j x[m,j] Ak[m,j]) + direct part of dS_k (scratch, bf16) + rank-1
    // feature terms, for this thread's row m.  A_k sits in the A buffer and is overwritten in place by C (bf16).  The summed x
    // is parked in the accumulator columns between the two passes.  Rows >= N stay zero.
    auto softmax_bwd_row = [&](bool from_tmem, int k) {
      if (!warp_on) return;   // (warp-uniform: the TMEM accesses below are warp-collective)
      const bool has_row = row < kRA;
      float rt = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxV; ++i)
        if (i == k) rt = rterm[i];
      float dot = 0.f;
      // the scratch rows come from L2 (~1k cycles away with two warps per scheduler): four 16-column chunks are in flight
      uint4 cur[8], nxt[8];
      auto fetch4 = [&](uint4* dst, int s, int c0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = (has_row && c0 + (i >> 1) < KS) ? ldcg16(map_chunk(s, 2 * c0 + i)) : make_uint4(0, 0, 0, 0);
      };
      fetch4(cur, kSlotDAF + k, 0);
      for (int c0 = 0; c0 < KS; c0 += 4) {
        fetch4(nxt, kSlotDAF + k, c0 + 4);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = c0 + ci;
          if (c < KS) {
            float v[16], a8[8], f8[8];
            if (from_tmem) { tmem_ld_32x32b_x16(tl + 16 * c, v); tmem_ld_wait(); }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (!from_tmem) unpack8(row < kRX ? *reinterpret_cast<const uint4*>(sm.X + (2 * c + hh) * (kRX * 16) + row * 16) : make_uint4(0, 0, 0, 0), v + 8 * hh);
              unpack8(cur[2 * ci + hh], f8);
              unpack8(has_row ? *reinterpret_cast<const uint4*>(sm.A + (2 * c + hh) * (kRA * 16) + row * 16) : make_uint4(0, 0, 0, 0), a8);
#pragma unroll
              for (int e = 0; e < 8; ++e) { v[8 * hh + e] += f8[e]; dot = fmaf(v[8 * hh + e], a8[e], dot); }
            }
            tmem_st_32x32b_x16(tl + 16 * c, v);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
      }
      tmem_st_wait();
      fetch4(cur, kSlotDS + k, 0);
      for (int c0 = 0; c0 < KS; c0 += 4) {
        fetch4(nxt, kSlotDS + k, c0 + 4);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = c0 + ci;
          if (c < KS) {
            float v[16], a8[8], d8[8];
            tmem_ld_32x32b_x16(tl + 16 * c, v);
            tmem_ld_wait();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              unsigned char* ap = sm.A + (2 * c + hh) * (kRA * 16) + row * 16;
              unpack8(has_row ? *reinterpret_cast<const uint4*>(ap) : make_uint4(0, 0, 0, 0), a8);
              unpack8(cur[2 * ci + hh], d8);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int j = 16 * c + 8 * hh + e;
                a8[e] = (row_ok && j < N) ? fmaf(a8[e], v[8 * hh + e] - dot, d8[e] + rt + cprime_s[k * kNmax + j]) : 0.f;
              }
              if (has_row) *reinterpret_cast<uint4*>(ap) = pack8(a8);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
      }
    };
    // One sweep.  On entry X holds the seed (written by its row owners) and the first view's A_k is in (or on its way into)
    // the A buffer.  Step s visits view k with the partial product `pslot` multiplying X from the left (transposed); the
    // final link is dA += X.  Loads are issued as early as their target buffer is free so that they overlap the epilogues:
    // P during the X' stash, X' and the next A_k during the dA epilogue / the softmax backward + contribution.
    auto sweep = [&](bool fchain) {
      for (int s = 0; s < V - 1; ++s) {
        const int k = fchain ? V - 1 - s : s;
        const int knext = fchain ? k - 1 : k + 1;   // view of the next step (or of the final link)
        int pslot;
        if (fchain) pslot = (k - 1 == 0) ? kSlotA + 0 : kSlotPfx + (k - 1) - 1;                 // P_{k-1}
        else pslot = (k + 1 == V - 1) ? kSlotA + V - 1 : kSlotSfx + (k + 1) - 1;               // A_{V-1}..A_{k+1}
        if (s == 1) MOP_TS(s_begin);
        cp_async_wait<0>();
        publish_cta();   // A_k and X (seed or reloaded image) landed / visible
        if (s == 1) MOP_TS(s_landed);
        if (blk_on) {
          if (t == 0) { mma_x_at(); commit(); }     // X' = X A_k^T
          mma_wait();
        }
        sync_cta();      // both warpgroups' MMAs have read A_k
        if (s == 1) MOP_TS(s_mma1);
        load_start(sm.A, slot(pslot), map_bytes);
        if (warp_on) {
          for (int c = 0; c < KS; ++c) {
            float v[16];
            tmem_ld_32x32b_x16(tl + 16 * c, v);
            tmem_ld_wait();
            if (row < kRX) {
              uint4 lo, hi;
              pack16(v, 1.f, lo, hi);
              if (!row_ok) lo = hi = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(slot(kSlotXN) + (size_t)(2 * c) * (kRX * 16) + row * 16) = lo;
              *reinterpret_cast<uint4*>(slot(kSlotXN) + (size_t)(2 * c + 1) * (kRX * 16) + row * 16) = hi;
            }
          }
        }
        if (s == 1) MOP_TS(s_stash);
        cp_async_wait<0>();
        publish_cta();   // P landed; the X' image is complete in the scratch
        if (s == 1) MOP_TS(s_pland);
        if (blk_on) {
          if (t == 0) { mma_at_x(); commit(); }     // dA_k part = P^T X
          mma_wait();
        }
        sync_cta();      // A buffer (P) and X are free
        if (s == 1) MOP_TS(s_mma2);
        if (fchain) {
          // the F chain only records its dA_k (bf16 rows in the scratch): the R chain adds it to its own dA_k, so that every
          // view goes through ONE softmax backward and ONE pair of dQ / dK contractions.  The next view's A and X' load meanwhile.
          load_start(sm.A, slot(kSlotA + knext), map_bytes);
          load_start(sm.X, slot(kSlotXN), xmap_bytes);
          if (warp_on) {
            for (int c = 0; c < KS; ++c) {
              float v[16];
              tmem_ld_32x32b_x16(tl + 16 * c, v);
              tmem_ld_wait();
              if (row < kRA) {
                uint4 lo, hi;
                pack16(v, 1.f, lo, hi);
                if (!row_ok) lo = hi = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(map_chunk(kSlotDAF + k, 2 * c)) = lo;
                *reinterpret_cast<uint4*>(map_chunk(kSlotDAF + k, 2 * c + 1)) = hi;
              }
            }
          }
        } else {
          // A_k comes back (turned into dS_k in place), X' replaces X
          load_start(sm.A, slot(kSlotA + k), map_bytes);
          load_start(sm.X, slot(kSlotXN), xmap_bytes);
          cp_async_wait<1>();   // A_k only
          fence_async_smem();
          __syncthreads();
          if (s == 1) MOP_TS(s_aland);
          softmax_bwd_row(true, k);
          if (s == 1) MOP_TS(s_smbwd);
          contribute(k, kSlotA + knext);
          if (s == 1) MOP_TS(s_contrib);
        }
      }
      // final link: dA += X
      cp_async_wait<0>();
      fence_async_smem();
      __syncthreads();   // the last X' landed (and the A map whose load the last step started)
      if (fchain) {
        // dA_0 of the F chain is the running product itself: this thread's row of X goes to the scratch.  The A buffer
        // already holds A_0, the first view of the R sweep.
        if (row < kRA) {
          for (int c = 0; c < 2 * KS; ++c)
            *reinterpret_cast<uint4*>(map_chunk(kSlotDAF + 0, c)) =
                row_ok ? *reinterpret_cast<const uint4*>(sm.X + c * (kRX * 16) + row * 16) : make_uint4(0, 0, 0, 0);
        }
      } else {
        softmax_bwd_row(false, V - 1);
        contribute(V - 1, -1);
      }
    };
    MOP_TS(T7);
    sync_cta();
    load_start(sm.A, slot(kSlotA + V - 1), map_bytes);   // first view of the F sweep
    sweep(true);
    // seed of the R sweep: X = (drho_R[i] + dkap_R[j]) / (R + eps)
    if (row < kRX) {
      uint4 rnext[2];
      rnext[0] = ldcg16(map_chunk(kSlotR, 0));
      rnext[1] = ldcg16(map_chunk(kSlotR, 1));
      for (int c = 0; c < 2 * KS; ++c) {
        float r8[8];
        unpack8(rnext[c & 1], r8);
        if (c + 2 < 2 * KS) rnext[c & 1] = ldcg16(map_chunk(kSlotR, c + 2));
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = 8 * c + e;
          r8[e] = (row_ok && j < N) ? (drhoR + dkapR_s[j < kNmax ? j : 0]) * fast_rcp(r8[e] + p.eps) : 0.f;
        }
        *reinterpret_cast<uint4*>(sm.X + c * (kRX * 16) + row * 16) = pack8(r8);
      }
    }
    MOP_TS(T8pre);
    sweep(false);
    MOP_TS(T9);
#ifdef MOP_PHASE_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0 && g == 0)
      for (int i = 0; i < ts_n; ++i) printf("ts %s %lld\n", ts_name[i], ts_v[i]);
#endif
    // =================================================================================================
    // outputs: dQ, dK rows; q/k scale partials
    // =================================================================================================
    if (row_ok) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c * 8 < dk) {
          const float4 a0 = __ldcg(dq_acc + (2 * c) * kNmax), a1 = __ldcg(dq_acc + (2 * c + 1) * kNmax);
          const float4 b0 = __ldcg(dk_acc + (2 * c) * kNmax), b1 = __ldcg(dk_acc + (2 * c + 1) * kNmax);
          const float qa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, ka[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          *reinterpret_cast<uint4*>(out_row(row) + 8 * c) = pack8(qa);
          *reinterpret_cast<uint4*>(out_row(row) + hd + 8 * c) = pack8(ka);
        }
      }
    }
    if (p.dscale_part) {
      float* ds = p.dscale_part + (size_t)g * 3 * V * dk;
      for (int idx = tid; idx < V * dk; idx += 256) {
        const int k = idx / dk, d = idx % dk;
        const float z = sscale * sm.zsum[k][d];
        const size_t pi = ((size_t)k * H + ph) * dk + d;
        ds[idx] = p.k_scale[pi] * z;
        ds[(size_t)V * dk + idx] = p.q_scale[pi] * z;
      }
    }
    sync_cta();
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tbase);
}

inline bool supported_bwd(const MopEdgewiseParams* p) { return supported(p) && p->row_stats != nullptr && p->y_base != nullptr && p->aux != nullptr; }

}  // namespace ewl
}  // namespace mop
