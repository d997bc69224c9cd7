// Edgewise (Mixture-of-Products) attention forward on tcgen05 / TMEM for token counts up to 200 (ViT-B/16: N = 196,
// dk = 64, V = 5), 512 threads: TWO THREADS PER ROW.
//
// Algorithm, buffers and passes: see the top of edgewise_tc_large.cuh (pass R, pass F, flash-style final stage).  The first version
// of this kernel ran one thread per row (256 threads); what changed is who does the element work.  Every fp32 accumulator row (TMEM lane) is shared by the two
// threads tid and tid ^ 256 - warps w and w + 8 see the same 32 TMEM lanes - and each takes half of the row's columns
// (chunks of 16 in the passes, 16 of the 32 panel columns in the final stage).  Row statistics (max, sums, log sums) are
// combined through a 2 KB exchange buffer (all the shared memory that is left) and 64-thread named barriers.  With one thread per row the SM
// ran 7 busy warps (2 per scheduler) and every TMEM load, MUFU result and shared-memory round trip was exposed; 14 busy warps
// hide most of it.
//
// Math: SURVEY.md appendix A (reference attention_variants.py:500-562, :319-331); executable specification
// oracle/edgewise_manual.py.
#pragma once
#include "edgewise_tc_large.cuh"

#ifdef MOP_PHASE_TIMING
#define MOP_TSF(name) do { if (threadIdx.x == 0 && blockIdx.x == 0 && g == 0 && ts_n < 40) { ts_v[ts_n] = clock64(); ts_name[ts_n++] = #name; } } while (0)
#else
#define MOP_TSF(name) do { } while (0)
#endif

namespace mop {
namespace ewl {

struct __align__(128) Smem2 {
  unsigned char X[kBufX];
  unsigned char A[kBufA];
  unsigned char Q[kQt];
  unsigned char K[kKt];
  float colsum[2][kNmax];          // column sums of log(F + eps) (0) and log(R + eps) (1)
  float cvec[kMaxV][64];
  float vs1[64], vsL[64];
  float qsum[64], ksraw[64];       // column sums of the query / raw key rows (row and column means of S_k are rank-1: see row_col_means)
  float xch[512];                  // row-statistic exchange between the two threads of a row
  float hw[2][kMaxQ * (2 * kMaxV + 2) + kMaxQ];   // gate-head weights + biases (row / column projection), staged once per CTA
  uint64_t bar[2];
  uint32_t tmem_slot;
};
static_assert(sizeof(Smem2) + 128 <= 232448, "forward shared memory over the 227 KB limit");

static __global__ void __launch_bounds__(512, 1) edgewise_fwd2_kernel(MopEdgewiseParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem2& sm = *reinterpret_cast<Smem2*>(smem_raw);
  const int tid = threadIdx.x, half = tid >> 8, wg = (tid >> 7) & 1, t = tid & 127, warp4 = (tid >> 5) & 3, lane = tid & 31;
  const int N = p.N, V = p.V, r = p.gate_rank, C = 2 * V + 2, dk = p.dk, H = p.H;
  const int KS = (N + 15) >> 4, NN = KS * 16;
  const int cfull = N >> 4;
  const int dks = (dk + 15) >> 4;
  const int row = 128 * wg + t;
  const bool row_ok = row < N;
  const bool blk_on = 128 * wg < N;
  const bool warp_on = 128 * wg + 32 * warp4 < N;
  const bool issuer = t == 0 && half == 0;              // the thread of this row block that issues its MMAs
  const int KH = (KS + 1) >> 1;
  const int cb = half ? KH : 0, ce = half ? KS : KH;    // this thread's 16-column chunks of a row
  const float invN = 1.f / (float)N;
  const uint32_t map_bytes = (uint32_t)(2 * KS) * (kRA * 16);
  unsigned char* spill = reinterpret_cast<unsigned char*>(p.workspace) + (size_t)blockIdx.x * kMaxV * kBufA;

  if (tid < 32) tmem_alloc<512>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); fence_mbar_init(); }
  {
    // with 227 KB of shared memory the L1 is a few KB: weights read with __ldg per problem came from L2 every time
    const int nW = 4 * r * C, nP = nW + 4 * r;
    for (int idx = tid; idx < 2 * nP; idx += 512) {
      const int hf = idx / nP, rem = idx % nP;
      sm.hw[hf][rem] = rem < nW ? (hf ? p.col_w : p.row_w)[rem] : (hf ? p.col_b : p.row_b)[rem - nW];
    }
  }
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
  auto publish_cta = [&]() { fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after(); };
  // the 256 threads of one row block
  auto publish_wg = [&]() { fence_async_smem(); tc_fence_before(); asm volatile("bar.sync %0, 256;" ::"r"(wg + 1) : "memory"); tc_fence_after(); };
  // the two warps that share 32 rows: the other thread's value of a row statistic (second barrier: the slot may be rewritten)
  const int pair_bar = 3 + ((tid >> 5) & 7);
  auto exchange = [&](float v) {
    sm.xch[tid] = v;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    const float o = sm.xch[tid ^ 256];
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    return o;
  };
  auto chain_mma = [&]() {
    const uint32_t id = idesc_bf16(128, NN, 0, 1);
    for (int ks = 0; ks < KS; ++ks)
      mma_ss(tD, desc_kmajor(sX + 128 * wg * 16, kRX, 16 * ks), desc_mnmajor(sA, kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
    mma_commit(&sm.bar[wg]);
  };
  // accumulator -> bf16 row of X (STORE) and / or row + column sums of log(D + eps) (LOGS): this thread's chunks
  auto chain_epilogue = [&](bool store, bool logs, int slot, float& rowmean) {
    float ls = 0.f;
    for (int c = cb; c < ce; ++c) {
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
    if (logs) rowmean = (ls + exchange(ls)) * invN;
  };

  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p.qkv);
  const size_t hd = (size_t)H * dk;
  const int G = p.B * H;
#ifdef MOP_PHASE_TIMING
  long long ts_v[40];
  const char* ts_name[40];
  int ts_n = 0;
#endif
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int pb = g / H, ph = g % H;
    auto in_row = [&](int n) { return qkv + (((size_t)pb * N + n) * 3) * hd + (size_t)ph * dk; };   // q; +hd: k; +2hd: v
    float* aux = p.aux ? p.aux + (size_t)g * kAuxLFloats : nullptr;   // statistics / means / factors for the backward
    MOP_TSF(F0);
    // a [rows x 64] operand tile from the token rows of q (which = 0), k (1) or v (2), optionally scaled per feature:
    // item = (row, 8-feature chunk), consecutive threads take the chunks of one row (128 contiguous bytes)
    // (R <= 208: at most four items per thread, all loads issued before the first use)
    auto load_tile = [&](unsigned char* dst, int R, int which, const float* scale) {
      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int item = tid + 512 * i, n = item >> 3, ch = item & 7;
        v[i] = make_uint4(0, 0, 0, 0);
        if (n < N && ch * 8 < dk) v[i] = __ldg(reinterpret_cast<const uint4*>(in_row(n) + (size_t)which * hd + ch * 8));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int item = tid + 512 * i, n = item >> 3, ch = item & 7;
        if (item < R * 8) *reinterpret_cast<uint4*>(dst + ch * (R * 16) + n * 16) = scale ? scale_chunk(v[i], scale + ch * 8) : v[i];
      }
    };
    // =================================================================================================
    // stage 0: per-view scale vectors, unscaled Q tile, zeroed column sums
    // =================================================================================================
    for (int idx = tid; idx < V * 64; idx += 512) {
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
    for (int idx = tid; idx < 2 * kNmax; idx += 512) (&sm.colsum[0][0])[idx] = 0.f;
    if (tid < 128) (tid < 64 ? sm.qsum : sm.ksraw)[tid & 63] = 0.f;
    load_tile(sm.Q, kRX, 0, nullptr);
    __syncthreads();
    // column sums of the query and of the raw key rows: thread = (8-feature chunk, row group), the four lanes of a warp that
    // share a chunk are combined by shuffles, then shared atomics
    {
      const int ch = tid & 7;
      float aq[8], ak[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { aq[e] = 0.f; ak[e] = 0.f; }
      for (int n = tid >> 3; n < N; n += 64) {
        if (ch * 8 < dk) {
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(in_row(n) + ch * 8)), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) aq[e] += f[e];
          unpack8(__ldg(reinterpret_cast<const uint4*>(in_row(n) + hd + ch * 8)), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) ak[e] += f[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        aq[e] += __shfl_xor_sync(0xffffffffu, aq[e], 8);
        aq[e] += __shfl_xor_sync(0xffffffffu, aq[e], 16);
        ak[e] += __shfl_xor_sync(0xffffffffu, ak[e], 8);
        ak[e] += __shfl_xor_sync(0xffffffffu, ak[e], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { atomicAdd(&sm.qsum[8 * ch + e], aq[e]); atomicAdd(&sm.ksraw[8 * ch + e], ak[e]); }
      }
    }
    __syncthreads();
    // row means rho_k[i] and column means kap_k[j] of S_k for this thread's token (as row i and as column j), both threads of a row
    float rho[kMaxV], kap[kMaxV], rhoF = 0.f, rhoR = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) { rho[i] = 0.f; kap[i] = 0.f; }
    // =================================================================================================
    // pass R (views V-1 .. 0): A_k = softmax(S_k) (spilled to the L2-resident scratch), Y <- Y A_k
    // =================================================================================================
    for (int idx = 0; idx < V; ++idx) {
      const int k = V - 1 - idx;
      const bool first = idx == 0, last = idx == V - 1;
      // scaled keys of view k (every MMA issued so far has completed: each thread waited for its row block's MMAs before
      // the last CTA barrier)
      if (idx == 1) MOP_TSF(v_begin);
      load_tile(sm.K, kRA, 1, sm.cvec[k]);
      publish_cta();
      if (idx == 1) MOP_TSF(v_ktile);
      if (blk_on) {
        if (issuer) {
          const uint32_t id = idesc_bf16(128, NN, 0, 0);
          for (int ks = 0; ks < dks; ++ks)
            mma_ss(tD, desc_kmajor(sQ + 128 * wg * 16, kRX, 16 * ks), desc_kmajor(sK, kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
          mma_commit(&sm.bar[wg]);
        }
        mma_wait();
      }
      if (idx == 1) MOP_TSF(v_smma);
      // ---- row softmax of S_k, two threads per row; row / column sums of S_k for the gate features
      if (warp_on) {
        float mx = -INFINITY;
        for (int c = cb; c < ce; ++c) {
          float v[16];
          tmem_ld_32x32b_x16(tl + 16 * c, v);
          tmem_ld_wait();
          if (c < cfull) {
#pragma unroll
            for (int e = 0; e < 16; ++e) mx = fmaxf(mx, v[e]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (16 * c + e < N) mx = fmaxf(mx, v[e]);
          }
        }
        mx = fmaxf(mx, exchange(mx));
        // S_k = Q Ks_k^T is a product, so its row and column sums are rank-1: row i sums to q_i . (sum_j ks_j), column j to
        // (sum_i q_i) . ks_j - two 64-term dot products instead of sums over the map.  First thread of the row: the row mean,
        // with sum_j ks_j = c_k (.) (sum of the raw keys); second thread: the column mean from the scaled key row in the K tile.
        {
          float part = 0.f;
          const unsigned char* src = half ? sm.K : sm.Q;
          const int R = half ? kRA : kRX;
          if (row < R) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              float f[8];
              unpack8(*reinterpret_cast<const uint4*>(src + ch * (R * 16) + row * 16), f);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float wv = half ? sm.qsum[8 * ch + e] : sm.cvec[k][8 * ch + e] * sm.ksraw[8 * ch + e];
                part = fmaf(f[e], wv, part);
              }
            }
          }
          part = row_ok ? part * invN : 0.f;
          const float other = exchange(part);
#pragma unroll
          for (int i = 0; i < kMaxV; ++i)
            if (i == k) { rho[i] = half ? other : part; kap[i] = half ? part : other; }
        }
        float l = 0.f;
        const float mb = mx * kLog2e;
        for (int c = cb; c < ce; ++c) {
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
        l += exchange(l);
        const float inv_l = 1.f / l;
        if (aux && half == 0 && row_ok) {
          aux[kAuxLStats + (2 * k) * kNmax + row] = mb;
          aux[kAuxLStats + (2 * k + 1) * kNmax + row] = inv_l;
        }
        for (int c = cb; c < ce; ++c) {
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
        for (int c = 2 * cb; c < 2 * ce; ++c) *reinterpret_cast<uint4*>(sm.A + c * (kRA * 16) + row * 16) = make_uint4(0, 0, 0, 0);
      }
      publish_cta();
      if (idx == 1) MOP_TSF(v_softmax);
      if (first) continue;
      if (blk_on) {
        if (issuer) chain_mma();
        mma_wait();
      }
      if (idx == 1) MOP_TSF(v_cmma);
      if (warp_on) chain_epilogue(!last, last, 1, rhoR);   // R itself is never needed again
      if (idx == 1) MOP_TSF(v_cepi);
      if (last && row < kRX) {
        // X <- A_0: start of the forward chain (this row block's MMA, the only reader of these X rows, is complete)
        for (int c = 2 * cb; c < 2 * ce; ++c)
          *reinterpret_cast<uint4*>(sm.X + c * (kRX * 16) + row * 16) =
              row < kRA ? *reinterpret_cast<const uint4*>(sm.A + c * (kRA * 16) + row * 16) : make_uint4(0, 0, 0, 0);
      }
    }
    // =================================================================================================
    // pass F (views 1 .. V-1): A_k comes back from the scratch, X <- X A_k; F stays in X
    // =================================================================================================
    MOP_TSF(F1);
    publish_cta();   // X = A_0 visible to the MMAs; nobody reads the A buffer any more
    cp_async_block(sm.A, spill + (size_t)1 * kBufA, map_bytes);
    cp_async_commit();
    for (int k = 1; k < V; ++k) {
      const bool last = k == V - 1;
      cp_async_wait<0>();
      publish_cta();   // A_k landed (and, for k > 1, the rewritten X rows are visible)
      if (blk_on) {
        if (issuer) chain_mma();
        mma_wait();
      }
      if (!last) {
        tc_fence_before();
        __syncthreads();   // both row blocks' MMAs have read A_k: the next map may land while X is rewritten
        cp_async_block(sm.A, spill + (size_t)(k + 1) * kBufA, map_bytes);
        cp_async_commit();
      }
      if (warp_on) chain_epilogue(true, last, 0, rhoF);
    }
    // =================================================================================================
    // final stage
    // =================================================================================================
    MOP_TSF(F2);
    publish_cta();   // F complete in X; chain MMAs of both row blocks are done: A and K buffers are free, colsum is final
    float* bfac = reinterpret_cast<float*>(sm.A + kOffBfac);
    unsigned char* Vt = sm.A + kOffVt;
    unsigned char* KsP = sm.A + kOffKsP + wg * (kMaxV * kKsP);
    unsigned char* Pt = sm.K + wg * kPt;
    load_tile(Vt, kRA, 2, sm.vs1);
    // gate factors: a (row factors of this thread's row, registers, both threads of the row) and b (column factors of column
    // `row`, shared: slots 0..7 from the first thread of the row, 8..15 from the second)
    float afac[kMaxQ];
    {
      // the C = 2V + 2 features of this token in the order of the projection weights: as a row (S_c row means, S_c^T row
      // means = column means of S_c, log F, log R row means) and as a column (roles swapped)
      constexpr int kMaxC = 2 * kMaxV + 2;
      float fr[kMaxC], fc[kMaxC];
      const float kapF = row < kNmax ? sm.colsum[0][row] * invN : 0.f, kapR = row < kNmax ? sm.colsum[1][row] * invN : 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxV; ++i) {
          if (c == i && i < V) { a = rho[i]; b = kap[i]; }
          if (c == V + i && i < V) { a = kap[i]; b = rho[i]; }
        }
        if (c == 2 * V) { a = rhoF; b = kapF; }
        if (c == 2 * V + 1) { a = rhoR; b = kapR; }
        fr[c] = a; fc[c] = b;
      }
      if (aux && half == 0 && row_ok) {
#pragma unroll
        for (int i = 0; i < kMaxV; ++i) { aux[kAuxLFeat + i * kNmax + row] = rho[i]; aux[kAuxLFeat + (kMaxV + i) * kNmax + row] = kap[i]; }
        aux[kAuxLFeat + (2 * kMaxV) * kNmax + row] = rhoF;
        aux[kAuxLFeat + (2 * kMaxV + 1) * kNmax + row] = rhoR;
        aux[kAuxLFeat + (2 * kMaxV + 2) * kNmax + row] = kapF;
        aux[kAuxLFeat + (2 * kMaxV + 3) * kNmax + row] = kapR;
      }
#pragma unroll
      for (int qq = 0; qq < kMaxQ; ++qq) {
        const int tg = qq >> 2, kk = qq & 3, q = tg * r + kk;
        float a = 0.f, b = 0.f;
        const bool mine = (qq >> 3) == half;   // this thread writes b of slot qq
        if (kk < r && row_ok) {
          // C is even: the weight rows (shared memory, warp-uniform addresses) are read as float2
          const float2* wr = reinterpret_cast<const float2*>(sm.hw[0] + q * C);
          const float2* wc = reinterpret_cast<const float2*>(sm.hw[1] + q * C);
          a = sm.hw[0][4 * r * C + q];
          if (mine) b = sm.hw[1][4 * r * C + q];
#pragma unroll
          for (int c2 = 0; c2 < kMaxC / 2; ++c2)
            if (2 * c2 < C) {
              const float2 w2 = wr[c2];
              a = fmaf(w2.x, fr[2 * c2], fmaf(w2.y, fr[2 * c2 + 1], a));
              if (mine) {
                const float2 v2 = wc[c2];
                b = fmaf(v2.x, fc[2 * c2], fmaf(v2.y, fc[2 * c2 + 1], b));
              }
            }
        }
        afac[qq] = a;
        if (mine && row < kNmax) bfac[row * 16 + qq] = b;
        if (aux && mine && row_ok) { aux[kAuxLA + qq * kNmax + row] = a; aux[kAuxLB + qq * kNmax + row] = b; }
      }
    }
    publish_cta();
    MOP_TSF(F3);
    float m_run = -INFINITY, l_run = 0.f;             // l_run: this thread's columns only (combined at the end)
    const uint32_t tS = tD, tO = tD + 160;            // MMA addresses: score panels (V x 32 columns), P V_1 accumulator
    const uint32_t tlS = tl, tlO = tl + 160;          // this warp's lane window
    if (blk_on) {
      const int npanels = (N + kPanel - 1) / kPanel;
      // raw key chunk of the next panel (one (key, 8-feature chunk) item per thread of the row block), prefetched
      const int pitem = t + 128 * half, pjj = pitem & 31, pch = pitem >> 5;
      uint4 raw;
      auto fetch_panel = [&](int j0) {
        const int j = j0 + pjj;
        raw = make_uint4(0, 0, 0, 0);
        if (j < N && pch * 8 < dk) raw = __ldg(reinterpret_cast<const uint4*>(in_row(j) + hd + pch * 8));
      };
      fetch_panel(0);
      for (int pn = 0; pn < npanels; ++pn) {
        const int j0 = pn * kPanel;
        // scaled key panels of the V views (the score MMAs of the previous panel have completed)
        for (int k = 0; k < V; ++k)
          *reinterpret_cast<uint4*>(KsP + k * kKsP + pch * (kPanel * 16) + pjj * 16) = scale_chunk(raw, &sm.cvec[k][pch * 8]);
        publish_wg();
        if (issuer) {
          const uint32_t id = idesc_bf16(128, kPanel, 0, 0);
          for (int k = 0; k < V; ++k)
            for (int ks = 0; ks < dks; ++ks)
              mma_ss(tS + 32 * k, desc_kmajor(sQ + 128 * wg * 16, kRX, 16 * ks), desc_kmajor(smem_u32(KsP) + k * kKsP, kPanel, 16 * ks), id, ks > 0 ? 1u : 0u);
          mma_commit(&sm.bar[wg]);
        }
        if (pn + 1 < npanels) fetch_panel(j0 + kPanel);
        if (pn == 1) MOP_TSF(p_issue);
        mma_wait();   // also covers the P V_1 MMA of the previous panel
        if (pn == 1) MOP_TSF(p_mma);
        if (warp_on) {
          float pmax = -INFINITY;
#pragma unroll 1
          for (int s2 = 0; s2 < 2; ++s2) {
            const int sub = 2 * half + s2;   // this thread's 8-column blocks of the panel
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
          if (pn == 1) MOP_TSF(p_mix);
          pmax = fmaxf(pmax, exchange(pmax));
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
            for (int c2 = 0; c2 < 2; ++c2) {   // this thread's half of the 64 accumulator columns
              const int c = 2 * half + c2;
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
          float smix[16];
          tmem_ld_32x32b_x16(tlS + 16 * half, smix);
          tmem_ld_wait();
          float ps = 0.f;
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            uint32_t u[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              u[e2] = pack_bf16(fast_exp2(fmaf(smix[8 * c2 + 2 * e2], kLog2e, -mb)), fast_exp2(fmaf(smix[8 * c2 + 2 * e2 + 1], kLog2e, -mb)));
              ps += __uint_as_float(u[e2] << 16) + __uint_as_float(u[e2] & 0xffff0000u);   // the rounded values (exp2(-inf) = 0)
            }
            *reinterpret_cast<uint4*>(Pt + (2 * half + c2) * (128 * 16) + t * 16) = make_uint4(u[0], u[1], u[2], u[3]);
          }
          l_run += ps;
        } else {
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) *reinterpret_cast<uint4*>(Pt + (2 * half + c2) * (128 * 16) + t * 16) = make_uint4(0, 0, 0, 0);
        }
        if (pn == 1) MOP_TSF(p_soft);
        publish_wg();
        if (issuer) {
          const uint32_t id = idesc_bf16(128, 64, 0, 1);
          const int nks = min(kPanel, NN - j0) >> 4;
          for (int ks = 0; ks < nks; ++ks)
            mma_ss(tO, desc_kmajor(smem_u32(Pt), 128, 16 * ks), desc_mnmajor(smem_u32(Vt), kRA, j0 + 16 * ks), id, (pn > 0 || ks > 0) ? 1u : 0u);
          if (pn == npanels - 1) mma_commit(&sm.bar[wg]);   // otherwise covered by the next panel's commit
        }
      }
      mma_wait();
    }
    MOP_TSF(F4);
    if (warp_on) l_run += exchange(l_run);   // row sum of the rounded probabilities over both threads' columns
    // ---- y = (P V_1) / l + F (w V_V)
    tc_fence_before();
    __syncthreads();          // both row blocks are done with V_1
    load_tile(Vt, kRA, 2, sm.vsL);
    publish_cta();
    if (blk_on) {
      if (issuer) {
        const uint32_t id = idesc_bf16(128, 64, 0, 1);
        for (int ks = 0; ks < KS; ++ks)
          mma_ss(tS, desc_kmajor(sX + 128 * wg * 16, kRX, 16 * ks), desc_mnmajor(smem_u32(Vt), kRA, 16 * ks), id, ks > 0 ? 1u : 0u);
        mma_commit(&sm.bar[wg]);
      }
      mma_wait();
      if (warp_on) {
        const float il = 1.f / l_run;
        if (p.row_stats && row_ok && half == 0) *reinterpret_cast<float2*>(p.row_stats + (((size_t)pb * H + ph) * N + row) * 2) = make_float2(m_run, l_run);
        __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + (((size_t)pb * N + row) * H + ph) * dk;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int c = 2 * half + c2;
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
    MOP_TSF(F5);
#ifdef MOP_PHASE_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0 && g == 0)
      for (int i = 0; i < ts_n; ++i) printf("tf %s %lld\n", ts_name[i], ts_v[i]);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tbase);
}

}  // namespace ewl
}  // namespace mop
