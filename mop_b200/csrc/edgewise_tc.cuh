// Edgewise (Mixture-of-Products) attention core on tcgen05 / TMEM, bf16 operands, fp32 accumulation
// and fp32 statistics.  Specialised for the config-1/2 hot shape: N = 64 tokens, dk <= 64 (dk % 8 == 0),
// V <= 5 shared-projection views, low-rank gate head with 4r <= 16.
//
// One CTA of 128 threads owns one (batch, head) problem at a time (persistent loop over problems):
//   * every N x N map is a 64x64 tile: fp32 accumulators live in TMEM (8 tiles = 512 columns),
//     bf16 MMA operands live in shared memory in the chunk-major layout of tc_common.cuh;
//   * contractions are M=64,N=64,K=16 tcgen05.mma (cta_group::1) issued by one thread and tracked
//     with tcgen05.commit -> mbarrier;
//   * accumulators are read with tcgen05.ld.16x256b so that all 128 threads hold a fragment
//     (2 rows x 16 columns each); row statistics are quad shuffles, column statistics are
//     xor-shuffles + a 4-warp shared-memory reduction;
//   * no N x N map ever reaches HBM: inputs are Q,K,V (bf16), output is y (bf16).
//
// Math: SURVEY.md appendix A / D.1 (reference attention_variants.py:500-562, :319-331); the
// executable specification is oracle/edgewise_manual.py.
#pragma once
#include "tc_common.cuh"

namespace mop {
namespace ewtc {

using namespace tc;

constexpr int kTile = 64 * 64 * 2;  // one bf16 64x64 operand tile
constexpr int kMaxV = 5;
constexpr int kMaxQ = 16;           // 4 * gate_rank
constexpr int kMaxC = 2 * kMaxV + 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// TMEM column map (fp32 64x64 tiles)
constexpr uint32_t kColS = 0;      // S_1..S_5 : 0, 64, ..., 256
constexpr uint32_t kColF = 320;    // forward chain accumulator
constexpr uint32_t kColR = 384;    // reverse chain accumulator
constexpr uint32_t kColY = 448;    // output accumulator

struct __align__(1024) SmemFwd {
  unsigned char K[kTile];
  unsigned char V1[kTile];          // V (.) v_scale[0]
  unsigned char VL[kTile];          // w * V (.) v_scale[V-1]
  unsigned char Qc[kMaxV][kTile];   // Q (.) (s q_scale_i k_scale_i); reused as XF0,XF1,XR0,XR1,Amix after stage 1
  unsigned char A[kMaxV][kTile];    // per-view probabilities
  float rho[kMaxC][64];
  float kap[kMaxC][64];
  float a[kMaxQ][64];
  float b[kMaxQ][64];
  float cvec[kMaxV][64];
  float vs1[64], vsL[64];
  float red[kMaxV + 2][4][64];      // cross-warp column sums
  uint64_t bar;
  uint32_t tmem_slot;
};

__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_log2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ float fast_log(float x) { return kLn2 * fast_log2(x); }

// x2 variant of the fragment load: 16 columns, 8 registers
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// all threads: make generic-proxy smem writes and finished TMEM reads visible before the next MMA batch
__device__ __forceinline__ void publish() {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// per-column sums of a 64x64 fragment over the 16 rows this warp owns -> red[warp][col]
__device__ __forceinline__ void colsum_to(float (*red)[64], const Frag& f, const float* v) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float s = v[4 * n + e] + v[4 * n + 2 + e];
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (f.lane < 4) red[f.warp][f.col(n) + e] = s;
    }
  }
}

struct Problem {
  int b, h;
};

// ------------------------------------------------------------------------------------------------
// stage 0: Q,K,V of one (b,h) -> shared-memory operand tiles (per-view scaled queries, scaled values)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_scales(const MopEdgewiseParams& p, SmemFwd& sm, int h, float w) {
  const int dk = p.dk, H = p.H, V = p.V;
  const float s = rsqrtf((float)dk);
  for (int idx = threadIdx.x; idx < V * 64; idx += 128) {
    int i = idx >> 6, d = idx & 63;
    float c = 0.f;
    if (d < dk) c = s * (p.q_scale ? p.q_scale[((size_t)i * H + h) * dk + d] * p.k_scale[((size_t)i * H + h) * dk + d] : 1.f);
    sm.cvec[i][d] = c;
  }
  for (int d = threadIdx.x; d < 64; d += 128) {
    float a = 0.f, b = 0.f;
    if (d < dk) {
      a = p.v_scale ? p.v_scale[((size_t)0 * H + h) * dk + d] : 1.f;
      b = p.v_scale ? p.v_scale[((size_t)(V - 1) * H + h) * dk + d] : 1.f;
    }
    sm.vs1[d] = a;
    sm.vsL[d] = w * b;
  }
}

__device__ __forceinline__ uint4 scale_chunk(uint4 raw, const float* sc) {
  uint32_t in[4] = {raw.x, raw.y, raw.z, raw.w}, out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = unpack_bf16(in[j]);
    out[j] = pack_bf16(f.x * sc[2 * j], f.y * sc[2 * j + 1]);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ void load_qkv_tiles(const MopEdgewiseParams& p, SmemFwd& sm, const Problem& pr) {
  const int dk = p.dk, H = p.H, V = p.V, dk8 = dk >> 3;
  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p.qkv);
  const size_t hd = (size_t)H * dk;
  for (int idx = threadIdx.x; idx < 64 * 8; idx += 128) {
    const int r = idx & 63, ch = idx >> 6;
    const uint32_t off = ch * 1024 + r * 16;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q;
    if (ch < dk8) {
      const __nv_bfloat16* base = qkv + (((size_t)pr.b * 64 + r) * 3) * hd + (size_t)pr.h * dk + ch * 8;
      q = *reinterpret_cast<const uint4*>(base);
      k = *reinterpret_cast<const uint4*>(base + hd);
      v = *reinterpret_cast<const uint4*>(base + 2 * hd);
    }
    *reinterpret_cast<uint4*>(sm.K + off) = k;
    *reinterpret_cast<uint4*>(sm.V1 + off) = scale_chunk(v, &sm.vs1[ch * 8]);
    *reinterpret_cast<uint4*>(sm.VL + off) = scale_chunk(v, &sm.vsL[ch * 8]);
    for (int i = 0; i < V; ++i) *reinterpret_cast<uint4*>(sm.Qc[i] + off) = scale_chunk(q, &sm.cvec[i][ch * 8]);
  }
}

// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) fwd_kernel(MopEdgewiseParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemFwd& sm = *reinterpret_cast<SmemFwd*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int V = p.V, r = p.gate_rank, C = 2 * V + 2;
  const int ksteps = (p.dk + 15) >> 4;
  const Frag f;

  if (warp == 0) tmem_alloc<512>(&sm.tmem_slot);
  if (tid == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_slot;
  const uint32_t tlane = tbase + ((uint32_t)(32 * warp) << 16);  // this warp's TMEM lanes
  uint32_t phase = 0;
  const float w = 1.f / (1.f + __expf(-p.chain_value_logit[0]));
  const float bn = p.beta_not / (float)max(1, V - 1);
  const uint32_t sK = smem_u32(sm.K), sV1 = smem_u32(sm.V1), sVL = smem_u32(sm.VL);
  unsigned char* XF[2] = {sm.Qc[0], sm.Qc[1]};
  unsigned char* XR[2] = {sm.Qc[2], sm.Qc[3]};
  unsigned char* Amix = sm.Qc[4];

  auto gemm = [&](uint32_t dcol, uint32_t a_tile, bool a_mn, uint32_t b_tile, bool b_mn, bool acc, int ks) {
    const uint32_t id = idesc_bf16(64, 64, a_mn ? 1u : 0u, b_mn ? 1u : 0u);
    for (int k = 0; k < ks; ++k) {
      uint64_t ad = a_mn ? desc_mnmajor(a_tile, 64, 16 * k) : desc_kmajor(a_tile, 64, 16 * k);
      uint64_t bd = b_mn ? desc_mnmajor(b_tile, 64, 16 * k) : desc_kmajor(b_tile, 64, 16 * k);
      mma_ss(tbase + dcol, ad, bd, id, (acc || k > 0) ? 1u : 0u);
    }
  };
  auto wait_mma = [&]() { mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after(); };

  const int G = p.B * p.H;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const Problem pr{g / p.H, g % p.H};
    // ---- stage 0: operands ------------------------------------------------------------------
    load_scales(p, sm, pr.h, w);
    __syncthreads();
    load_qkv_tiles(p, sm, pr);
    publish();
    // ---- stage 1: S_i = Qc_i K^T ------------------------------------------------------------
    if (tid == 0) {
      for (int i = 0; i < V; ++i) gemm(kColS + 64 * i, smem_u32(sm.Qc[i]), false, sK, false, false, ksteps);
      mma_commit(&sm.bar);
    }
    wait_mma();
    // ---- per-view row softmax, row/column means of S_i ------------------------------------------
    for (int i = 0; i < V; ++i) {
      float v[32];
      tmem_ld_16x256b_x8(tlane + kColS + 64 * i, v);
      tmem_ld_wait();
      float mlo = -INFINITY, mhi = -INFINITY, slo = 0.f, shi = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        mlo = fmaxf(mlo, fmaxf(v[4 * n], v[4 * n + 1]));
        mhi = fmaxf(mhi, fmaxf(v[4 * n + 2], v[4 * n + 3]));
        slo += v[4 * n] + v[4 * n + 1];
        shi += v[4 * n + 2] + v[4 * n + 3];
      }
      mlo = quad_max(mlo); mhi = quad_max(mhi);
      slo = quad_sum(slo); shi = quad_sum(shi);
      if ((f.lane & 3) == 0) { sm.rho[i][f.row_lo] = slo * (1.f / 64.f); sm.rho[i][f.row_hi] = shi * (1.f / 64.f); }
      colsum_to(sm.red[i], f, v);
      float elo = 0.f, ehi = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        v[4 * n] = fast_exp2((v[4 * n] - mlo) * kLog2e);
        v[4 * n + 1] = fast_exp2((v[4 * n + 1] - mlo) * kLog2e);
        v[4 * n + 2] = fast_exp2((v[4 * n + 2] - mhi) * kLog2e);
        v[4 * n + 3] = fast_exp2((v[4 * n + 3] - mhi) * kLog2e);
        elo += v[4 * n] + v[4 * n + 1];
        ehi += v[4 * n + 2] + v[4 * n + 3];
      }
      const float ilo = 1.f / quad_sum(elo), ihi = 1.f / quad_sum(ehi);
#pragma unroll
      for (int n = 0; n < 8; ++n) { v[4 * n] *= ilo; v[4 * n + 1] *= ilo; v[4 * n + 2] *= ihi; v[4 * n + 3] *= ihi; }
      frag_store_bf16(sm.A[i], f, v);
    }
    // ---- stage 1b: chain products F = A_1..A_V, R = A_V..A_1 ---------------------------------------
    {
      uint32_t xf = smem_u32(sm.A[0]), xr = smem_u32(sm.A[V - 1]);
      for (int s = 1; s < V; ++s) {
        publish();
        if (tid == 0) {
          gemm(kColF, xf, false, smem_u32(sm.A[s]), true, false, 4);
          gemm(kColR, xr, false, smem_u32(sm.A[V - 1 - s]), true, false, 4);
          mma_commit(&sm.bar);
        }
        wait_mma();
        float v[32];
        tmem_ld_16x256b_x8(tlane + kColF, v);
        tmem_ld_wait();
        frag_store_bf16(XF[s & 1], f, v);
        xf = smem_u32(XF[s & 1]);
        if (s < V - 1) {
          tmem_ld_16x256b_x8(tlane + kColR, v);
          tmem_ld_wait();
          frag_store_bf16(XR[s & 1], f, v);
          xr = smem_u32(XR[s & 1]);
        }
      }
      // log-chain features: row / column means of log(F+eps), log(R+eps)
      for (int which = 0; which < 2; ++which) {
        float v[32];
        tmem_ld_16x256b_x8(tlane + (which ? kColR : kColF), v);
        tmem_ld_wait();
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
          sm.rho[2 * V + which][f.row_lo] = slo * (1.f / 64.f);
          sm.rho[2 * V + which][f.row_hi] = shi * (1.f / 64.f);
        }
        colsum_to(sm.red[kMaxV + which], f, v);
      }
      const uint32_t sF = xf;  // bf16 copy of F (operand of the value-transport GEMM)
      __syncthreads();
      // finish column means; channels V+c are the transposes: rho_{V+c} = kap_c, kap_{V+c} = rho_c
      for (int idx = tid; idx < (V + 2) * 64; idx += 128) {
        int m = idx >> 6, j = idx & 63;
        int slot = m < V ? m : kMaxV + (m - V);
        float s = sm.red[slot][0][j] + sm.red[slot][1][j] + sm.red[slot][2][j] + sm.red[slot][3][j];
        sm.kap[m < V ? m : 2 * V + (m - V)][j] = s * (1.f / 64.f);
      }
      __syncthreads();
      // ---- stage 2: low-rank gate factors a[q][i], b[q][j] --------------------------------------------
      {
        const int which = tid >> 6, tok = tid & 63;  // 0: a (row factors), 1: b (column factors)
        const float* W = which ? p.col_w : p.row_w;
        const float* bias = which ? p.col_b : p.row_b;
        float (*own)[64] = which ? sm.kap : sm.rho;    // feature of channel c
        float (*swp)[64] = which ? sm.rho : sm.kap;    // feature of channel V+c
        for (int qq = 0; qq < kMaxQ; ++qq) {  // slot qq = t*4 + k  <->  reference row q = t*r + k; unused slots are 0
          const int t = qq >> 2, k = qq & 3, q = t * r + k;
          float acc = 0.f;
          if (k < r) {
            acc = __ldg(bias + q);
            for (int c = 0; c < V; ++c) {
              acc = fmaf(__ldg(W + q * C + c), own[c][tok], acc);
              acc = fmaf(__ldg(W + q * C + V + c), swp[c][tok], acc);
            }
            acc = fmaf(__ldg(W + q * C + 2 * V), own[2 * V][tok], acc);
            acc = fmaf(__ldg(W + q * C + 2 * V + 1), own[2 * V + 1][tok], acc);
          }
          (which ? sm.b : sm.a)[qq][tok] = acc;
        }
      }
      __syncthreads();
      // ---- stage 2c/3: mix, re-normalise ---------------------------------------------------------------
      float smix[32];
      {
        float alo[kMaxQ], ahi[kMaxQ];
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) { alo[q] = sm.a[q][f.row_lo]; ahi[q] = sm.a[q][f.row_hi]; }
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
          float sv[kMaxV][8], fv[8];
#pragma unroll
          for (int i = 0; i < kMaxV; ++i)
            if (i < V) tmem_ld_16x256b_x2(tlane + kColS + 64 * i + 16 * blk, sv[i]);
          tmem_ld_16x256b_x2(tlane + kColF + 16 * blk, fv);
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
            for (int q = 0; q < kMaxQ; ++q) z[q >> 2] = fmaf(hi ? ahi[q] : alo[q], sm.b[q][col], z[q >> 2]);
            smix[8 * blk + e] = s0 + fast_sigmoid(z[0]) * U + fast_sigmoid(z[1]) * O - fast_sigmoid(z[2]) * bn * U +
                                fast_sigmoid(z[3]) * lf;
          }
        }
      }
      {
        float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          mlo = fmaxf(mlo, fmaxf(smix[4 * n], smix[4 * n + 1]));
          mhi = fmaxf(mhi, fmaxf(smix[4 * n + 2], smix[4 * n + 3]));
        }
        mlo = quad_max(mlo); mhi = quad_max(mhi);
        float elo = 0.f, ehi = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          smix[4 * n] = fast_exp2((smix[4 * n] - mlo) * kLog2e);
          smix[4 * n + 1] = fast_exp2((smix[4 * n + 1] - mlo) * kLog2e);
          smix[4 * n + 2] = fast_exp2((smix[4 * n + 2] - mhi) * kLog2e);
          smix[4 * n + 3] = fast_exp2((smix[4 * n + 3] - mhi) * kLog2e);
          elo += smix[4 * n] + smix[4 * n + 1];
          ehi += smix[4 * n + 2] + smix[4 * n + 3];
        }
        const float ilo = 1.f / quad_sum(elo), ihi = 1.f / quad_sum(ehi);
#pragma unroll
        for (int n = 0; n < 8; ++n) { smix[4 * n] *= ilo; smix[4 * n + 1] *= ilo; smix[4 * n + 2] *= ihi; smix[4 * n + 3] *= ihi; }
        frag_store_bf16(Amix, f, smix);
      }
      // ---- y = A V_1 + w F V_V ------------------------------------------------------------------------
      publish();
      if (tid == 0) {
        gemm(kColY, smem_u32(Amix), false, sV1, true, false, 4);
        gemm(kColY, sF, false, sVL, true, true, 4);
        mma_commit(&sm.bar);
      }
      wait_mma();
      float yv[32];
      tmem_ld_16x256b_x8(tlane + kColY, yv);
      tmem_ld_wait();
      __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y);
      const size_t row_lo = (((size_t)pr.b * 64 + f.row_lo) * p.H + pr.h) * p.dk;
      const size_t row_hi = (((size_t)pr.b * 64 + f.row_hi) * p.H + pr.h) * p.dk;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int c = f.col(n);
        if (c < p.dk) {
          *reinterpret_cast<uint32_t*>(y + row_lo + c) = pack_bf16(yv[4 * n], yv[4 * n + 1]);
          *reinterpret_cast<uint32_t*>(y + row_hi + c) = pack_bf16(yv[4 * n + 2], yv[4 * n + 3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // tiles and TMEM are reused by the next problem
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

inline bool supported(const MopEdgewiseParams* p) {
  return p->dtype == MOP_BF16 && p->N == 64 && p->dk <= 64 && p->dk % 8 == 0 && p->V >= 2 && p->V <= kMaxV && p->Vp == 1 &&
         p->gate_mode == MOP_GATE_LOWRANK && p->gate_rank >= 1 && p->gate_rank <= 4;
}

}  // namespace ewtc
}  // namespace mop
