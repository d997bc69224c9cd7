// Edgewise (Mixture-of-Products) attention core on tcgen05 / TMEM: fused forward and fused backward.
// bf16 operands, fp32 accumulation and fp32 statistics.  Specialised for the config-1/2 hot shape:
// N = 64 tokens, dk <= 64 (dk % 8 == 0), V <= 5 shared-projection views, low-rank gate head, r <= 4.
//
// One CTA of 128 threads owns one (batch, head) problem at a time (persistent loop over problems):
//   * every N x N map is a 64x64 tile.  fp32 accumulators live in TMEM: an M=64 accumulator uses
//     lanes 0-15 of each 32-lane subpartition, so a second bank sits at lane offset 16 -> 16 tiles
//     in the 512 allocated columns.  bf16 MMA operands live in shared memory in the chunk-major
//     layout of tc_common.cuh (the same bytes serve as X and X^T);
//   * contractions are M=64,N=64|16,K=16 tcgen05.mma (cta_group::1) issued by one thread, batched
//     per dependency level and tracked with tcgen05.commit -> mbarrier;
//   * accumulators are read/written with tcgen05.ld/st.16x256b so that all 128 threads hold a
//     fragment (2 rows x 16 columns); row statistics are quad shuffles, column statistics are
//     xor-shuffles + a 4-warp shared-memory reduction;
//   * no N x N map ever reaches HBM.  Forward: Q,K,V in, y out.  Backward: Q,K,V,dy in, dqkv and
//     per-(b,h) parameter-gradient partials out; everything else is recomputed on chip.
//
// Math: SURVEY.md appendix A / D.1 (reference attention_variants.py:500-562, :319-331); the
// executable specification is oracle/edgewise_manual.py.
#pragma once
#include <type_traits>

#include "tc_common.cuh"

namespace mop {
namespace ewtc {

using namespace tc;

constexpr int kTile = 64 * 64 * 2;  // one bf16 64x64 operand tile
constexpr int kMaxV = 5;
constexpr int kMaxQ = 16;           // gate-factor slots: q = 4*gate + k, k < r <= 4 (unused slots hold 0)
constexpr int kMaxC = 2 * kMaxV + 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---- `aux`: small per-(b,h) vectors handed from the forward to the backward (MopEdgewiseParams::aux), in floats -----------
//   stats [(kMaxV+1)][64][2]  per row of S_k (k < V) and of the mixed map (k = V): max * log2(e), 1 / sum exp
//   rho   [kMaxC][64], kap [kMaxC][64]   row / column feature means (rows c < V and 2V, 2V+1 are used)
//   a     [kMaxQ][64], b [kMaxQ][64]     low-rank gate factors, slot q = 4 * gate + k (unused slots 0)
constexpr int kAuxStats = 0;
constexpr int kAuxRho = (kMaxV + 1) * 64 * 2;
constexpr int kAuxKap = kAuxRho + kMaxC * 64;
constexpr int kAuxA = kAuxKap + kMaxC * 64;
constexpr int kAuxB = kAuxA + kMaxQ * 64;
constexpr int kAuxFloats = kAuxB + kMaxQ * 64;   // 4352 floats = 17 KB

// ---- TMEM tile map: tile i -> lane offset 16*(i/8), column 64*(i%8) ---------------------------------
// backward: 16 tiles in 512 columns (tile i -> lane offset 16*(i/8), column 64*(i%8));
// forward: 8 tiles in 256 columns (lane offset 16*(i/4), column 64*(i%4)) so that two CTAs fit on one SM.
template <bool BWD>
__host__ __device__ constexpr uint32_t ttile(int i) {
  return BWD ? ((uint32_t)(i >= 8 ? 16 : 0) << 16) + 64u * (uint32_t)(i & 7) : ((uint32_t)(i >= 4 ? 16 : 0) << 16) + 64u * (uint32_t)(i & 3);
}
template <bool BWD> struct TmemCols { static constexpr uint32_t value = BWD ? 512 : 256; };
constexpr int kTS = 0;     // S_1..S_5 (later: dS_k accumulators)   tiles 0..4
constexpr int kTF = 5;     // forward chain product F
constexpr int kTR = 6;     // reverse chain product R
constexpr int kTY = 7;     // fwd: y accumulator;  bwd: dA, then dF
constexpr int kTdV1 = 8, kTdVL = 9, kTdb = 12;                       // bwd, "early" GEMMs
constexpr int kTAF = 8, kTXF = 9, kTAR = 10, kTXR = 11;                // bwd, chain sweep
constexpr int kTT = 8;                                                 // bwd, T_k = dS_k K : tiles 8..12
__host__ __device__ constexpr int tileU(int k) { return k < 3 ? 5 + k : 10 + k; }  // U_k = dS_k^T Q : 5,6,7,13,14

// ---- shared-memory tile slots -------------------------------------------------------------------------
template <bool BWD> struct Slots;
template <> struct Slots<false> {
  // A_i overwrites Qc_i (dead once S_i is in TMEM); the forward chain product is updated in place in the dead K
  // tile, the reverse one in tile 8 (an MMA that read X has completed before X is rewritten); Amix reuses tile 8.
  static constexpr int QC = 0, A = 0, K = 5, V1 = 6, VL = 7, AMIX = 8, N = 9;
  __device__ static int P(int) { return 5; }
  __device__ static int R(int) { return 8; }
};
template <> struct Slots<true> {
  static constexpr int K = 0, V1 = 1, VL = 2, DY = 3, AMIX = 4, QC = 5, A = 12, X = 17, N = 21;
  __device__ static int P(int s) { return 5 + (s - 1); }        // P(s) = A_0..A_s kept for the chain backward, s = 1..V-1
  __device__ static int R(int s) { return 9 + (s - 1); }        // step s produces A_{V-1}..A_{V-1-s}, kept for s = 1..V-2
};

struct SmemVec {
  float rho[kMaxC][64];
  float kap[kMaxC][64];
  float a[kMaxQ][64];               // bwd: reused for da
  float b[kMaxQ][64];               // bwd: reused for db
  float cvec[kMaxV][64];
  float vs1[64], vsL[64];
  float red[kMaxV + 2][4][64];      // cross-warp column sums
  float scal[8];
  float hw[2][kMaxQ * kMaxC + kMaxQ];   // gate-head weights + biases of the row / column projection (staged once per CTA)
  uint64_t bar;
  uint32_t tmem_slot;
};

struct SmemBwdVec {
  float drho[kMaxC][64];
  float dkap[kMaxC][64];
  float da[kMaxQ][64];              // grad of the row factors (fp32, CUDA-core accumulation)
  unsigned char a_bf[64 * 16 * 2];  // bf16 [token][q] operand tile of the column-factor gradient GEMM
};
struct Empty {};
template <bool BWD>
struct __align__(1024) Smem {
  unsigned char T[Slots<BWD>::N][kTile];
  SmemVec v;
  typename std::conditional<BWD, SmemBwdVec, Empty>::type bv;
};

__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_log2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ float fast_log(float x) { return kLn2 * fast_log2(x); }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// x2 variants of the fragment load/store: 16 columns, 8 registers
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// all threads: make generic-proxy smem writes and finished TMEM accesses visible before the next MMA batch
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

// row softmax of a fragment in place (fp32): v becomes probabilities; st4 (optional) receives max_lo, max_hi, 1/sum_lo, 1/sum_hi
__device__ __forceinline__ void frag_softmax(float* v, float* st4 = nullptr) {
  float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    mlo = fmaxf(mlo, fmaxf(v[4 * n], v[4 * n + 1]));
    mhi = fmaxf(mhi, fmaxf(v[4 * n + 2], v[4 * n + 3]));
  }
  mlo = quad_max(mlo); mhi = quad_max(mhi);
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
  if (st4) { st4[0] = mlo; st4[1] = mhi; st4[2] = ilo; st4[3] = ihi; }
}

// in place: x = P (.) (x - rowsum(x (.) P))   (softmax backward on fragments)
__device__ __forceinline__ void frag_softmax_bwd(float* x, const float* P) {
  float dlo = 0.f, dhi = 0.f;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    dlo = fmaf(x[4 * n], P[4 * n], fmaf(x[4 * n + 1], P[4 * n + 1], dlo));
    dhi = fmaf(x[4 * n + 2], P[4 * n + 2], fmaf(x[4 * n + 3], P[4 * n + 3], dhi));
  }
  dlo = quad_sum(dlo); dhi = quad_sum(dhi);
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    x[4 * n] = P[4 * n] * (x[4 * n] - dlo);
    x[4 * n + 1] = P[4 * n + 1] * (x[4 * n + 1] - dlo);
    x[4 * n + 2] = P[4 * n + 2] * (x[4 * n + 2] - dhi);
    x[4 * n + 3] = P[4 * n + 3] * (x[4 * n + 3] - dhi);
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

// sum over the whole CTA (128 threads); every thread gets the result
__device__ __forceinline__ float cta_sum(float v, float* scratch4) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch4[threadIdx.x >> 5] = v;
  __syncthreads();
  return scratch4[0] + scratch4[1] + scratch4[2] + scratch4[3];
}

template <bool BWD>
static __global__ void __launch_bounds__(128, BWD ? 1 : 2) edgewise_kernel(MopEdgewiseParams p) {
  using SL = Slots<BWD>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem<BWD>& sm = *reinterpret_cast<Smem<BWD>*>(smem_raw);
  SmemVec& sv_ = sm.v;
  auto& bv_ = sm.bv;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int V = p.V, r = p.gate_rank, C = 2 * V + 2, dk = p.dk, H = p.H;
  const int ksteps = (dk + 15) >> 4;
  const Frag f;

  if (warp == 0) tmem_alloc<TmemCols<BWD>::value>(&sv_.tmem_slot);
  // The MMAs of one step are issued by lane 0 of ALL four warps (independent destination tiles, dealt round-robin): a lone
  // issuing lane needs ~10 cycles per instruction and ~12 instructions per tcgen05.mma, and everybody else waits for it.
  // Every leader commits (a commit with nothing outstanding arrives at once), so the barrier always counts four arrivals.
  constexpr int kIssuers = 4;
  if (tid == 0) { mbar_init(&sv_.bar, kIssuers); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const bool leader = (tid & 31) == 0;
  auto mine = [&](int idx) { return (idx & (kIssuers - 1)) == warp; };
  const uint32_t tbase = sv_.tmem_slot;
  const uint32_t tlane = tbase + ((uint32_t)(32 * warp) << 16);  // this warp's TMEM subpartition
  uint32_t phase = 0;
  const float w = 1.f / (1.f + __expf(-p.chain_value_logit[0]));
  const float bn = p.beta_not / (float)max(1, V - 1);
  const float sscale = rsqrtf((float)dk);

  auto tile = [&](int slot) -> unsigned char* { return sm.T[slot]; };
  auto taddr = [&](int slot) -> uint32_t { return smem_u32(sm.T[slot]); };
  // D[tile dt (+dcol)] (+)= op(A) op(B), K = 16*ks.  a_mn/b_mn: operand tile is used MN-major.
  auto gemm = [&](int dt, uint32_t dcol, uint32_t a_tile, bool a_mn, uint32_t b_tile, bool b_mn, bool acc, int ks, uint32_t n) {
    const uint32_t id = idesc_bf16(64, n, a_mn ? 1u : 0u, b_mn ? 1u : 0u);
    for (int k = 0; k < ks; ++k) {
      uint64_t ad = a_mn ? desc_mnmajor(a_tile, 64, 16 * k) : desc_kmajor(a_tile, 64, 16 * k);
      uint64_t bd = b_mn ? desc_mnmajor(b_tile, 64, 16 * k) : desc_kmajor(b_tile, 64, 16 * k);
      mma_ss(tbase + ttile<BWD>(dt) + dcol, ad, bd, id, (acc || k > 0) ? 1u : 0u);
    }
  };
  auto wait_mma = [&]() { mbar_wait(&sv_.bar, phase); phase ^= 1; tc_fence_after(); };
  auto ld_tile = [&](int t, float* v) { tmem_ld_16x256b_x8(tlane + ttile<BWD>(t), v); tmem_ld_wait(); };
  auto st_tile = [&](int t, const float* v) { tmem_st_16x256b_x8(tlane + ttile<BWD>(t), v); tmem_st_wait(); };

  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p.qkv);
  const size_t hd = (size_t)H * dk;
  const int G = p.B * H;
  {   // head weights: read through L2 once per CTA instead of once per use
    const int nW = 4 * r * C;
    for (int idx = tid; idx < 2 * (nW + 4 * r); idx += 128) {
      const int half = idx / (nW + 4 * r), rem = idx % (nW + 4 * r);
      sv_.hw[half][rem] = rem < nW ? (half ? p.col_w : p.row_w)[rem] : (half ? p.col_b : p.row_b)[rem - nW];
    }
    __syncthreads();
  }
  const int hw_bias = 4 * r * C;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int pb = g / H, ph = g % H;
    float* aux = nullptr;
    if constexpr (!BWD) aux = p.aux ? p.aux + (size_t)g * kAuxFloats : nullptr;
    auto put_stats = [&](int k, const float* st4) {   // row statistics of map k for the backward
      if (aux && (f.lane & 3) == 0) {
        *reinterpret_cast<float2*>(aux + kAuxStats + (k * 64 + f.row_lo) * 2) = make_float2(st4[0] * kLog2e, st4[2]);
        *reinterpret_cast<float2*>(aux + kAuxStats + (k * 64 + f.row_hi) * 2) = make_float2(st4[1] * kLog2e, st4[3]);
      }
    };
    // =================================================================================================
    // stage 0: scales and operand tiles
    // =================================================================================================
    for (int idx = tid; idx < V * 64; idx += 128) {
      int i = idx >> 6, d = idx & 63;
      float c = 0.f;
      if (d < dk) c = sscale * (p.q_scale ? p.q_scale[((size_t)i * H + ph) * dk + d] * p.k_scale[((size_t)i * H + ph) * dk + d] : 1.f);
      sv_.cvec[i][d] = c;
    }
    for (int d = tid; d < 64; d += 128) {
      float a = 0.f, b = 0.f;
      if (d < dk) {
        a = p.v_scale ? p.v_scale[((size_t)0 * H + ph) * dk + d] : 1.f;
        b = p.v_scale ? p.v_scale[((size_t)(V - 1) * H + ph) * dk + d] : 1.f;
      }
      sv_.vs1[d] = a;
      sv_.vsL[d] = w * b;
    }
    __syncthreads();
    for (int idx = tid; idx < 64 * 8; idx += 128) {
      const int rr = idx & 63, ch = idx >> 6;
      const uint32_t off = ch * 1024 + rr * 16;
      uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, dyv = q;
      if (ch * 8 < dk) {
        const __nv_bfloat16* base = qkv + (((size_t)pb * 64 + rr) * 3) * hd + (size_t)ph * dk + ch * 8;
        q = *reinterpret_cast<const uint4*>(base);
        k = *reinterpret_cast<const uint4*>(base + hd);
        v = *reinterpret_cast<const uint4*>(base + 2 * hd);
        if constexpr (BWD)
          dyv = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + (((size_t)pb * 64 + rr) * H + ph) * dk + ch * 8);
      }
      *reinterpret_cast<uint4*>(tile(SL::K) + off) = k;
      *reinterpret_cast<uint4*>(tile(SL::V1) + off) = scale_chunk(v, &sv_.vs1[ch * 8]);
      *reinterpret_cast<uint4*>(tile(SL::VL) + off) = scale_chunk(v, &sv_.vsL[ch * 8]);
      for (int i = 0; i < V; ++i) *reinterpret_cast<uint4*>(tile(SL::QC + i) + off) = scale_chunk(q, &sv_.cvec[i][ch * 8]);
      if constexpr (BWD) *reinterpret_cast<uint4*>(tile(Slots<true>::DY) + off) = dyv;
    }
    publish();
    // =================================================================================================
    // stage 1: S_i = Qc_i K^T   (bwd: also dA = dY V_1^T)
    // =================================================================================================
    if (leader) {
      for (int i = 0; i < V; ++i)
        if (mine(i)) gemm(kTS + i, 0, taddr(SL::QC + i), false, taddr(SL::K), false, false, ksteps, 64);
      if constexpr (BWD)
        if (mine(V)) gemm(kTY, 0, taddr(Slots<true>::DY), false, taddr(SL::V1), false, false, ksteps, 64);
      mma_commit(&sv_.bar);
    }
    wait_mma();
    // per-view row softmax + row / column means of S_i
    for (int i = 0; i < V; ++i) {
      float v[32];
      ld_tile(kTS + i, v);
      float slo = 0.f, shi = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) { slo += v[4 * n] + v[4 * n + 1]; shi += v[4 * n + 2] + v[4 * n + 3]; }
      slo = quad_sum(slo); shi = quad_sum(shi);
      if ((f.lane & 3) == 0) { sv_.rho[i][f.row_lo] = slo * (1.f / 64.f); sv_.rho[i][f.row_hi] = shi * (1.f / 64.f); }
      colsum_to(sv_.red[i], f, v);
      float st4[4];
      frag_softmax(v, st4);
      put_stats(i, st4);
      frag_store_bf16(tile(SL::A + i), f, v);
    }
    // chain products F = A_0..A_{V-1}, R = A_{V-1}..A_0
    uint32_t sF;
    {
      uint32_t xf = taddr(SL::A + 0), xr = taddr(SL::A + V - 1);
      for (int s = 1; s < V; ++s) {
        publish();
        if (leader) {
          if (mine(0)) gemm(kTF, 0, xf, false, taddr(SL::A + s), true, false, 4, 64);
          if (mine(1)) gemm(kTR, 0, xr, false, taddr(SL::A + V - 1 - s), true, false, 4, 64);
          mma_commit(&sv_.bar);
        }
        wait_mma();
        float v[32];
        ld_tile(kTF, v);
        frag_store_bf16(tile(SL::P(s)), f, v);
        xf = taddr(SL::P(s));
        if (s < V - 1) {
          ld_tile(kTR, v);
          frag_store_bf16(tile(SL::R(s)), f, v);
          xr = taddr(SL::R(s));
        }
      }
      sF = xf;  // bf16 copy of F
    }
    // log-chain features: row / column means of log(F+eps), log(R+eps)
    for (int which = 0; which < 2; ++which) {
      float v[32];
      ld_tile(which ? kTR : kTF, v);
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
        sv_.rho[2 * V + which][f.row_lo] = slo * (1.f / 64.f);
        sv_.rho[2 * V + which][f.row_hi] = shi * (1.f / 64.f);
      }
      colsum_to(sv_.red[kMaxV + which], f, v);
    }
    __syncthreads();
    for (int idx = tid; idx < (V + 2) * 64; idx += 128) {
      int m = idx >> 6, j = idx & 63;
      int slot = m < V ? m : kMaxV + (m - V);
      float s = sv_.red[slot][0][j] + sv_.red[slot][1][j] + sv_.red[slot][2][j] + sv_.red[slot][3][j];
      sv_.kap[m < V ? m : 2 * V + (m - V)][j] = s * (1.f / 64.f);
    }
    __syncthreads();
    // =================================================================================================
    // stage 2: low-rank gate factors.  Feature channel c < V is S_c, V + c is S_c^T (row/col roles
    // swapped), 2V / 2V+1 are log F / log R.
    // =================================================================================================
    {
      const int which = tid >> 6, tok = tid & 63;  // 0: a (row factors), 1: b (column factors)
      const float* W = sv_.hw[which];
      const float* bias = W + hw_bias;
      float (*own)[64] = which ? sv_.kap : sv_.rho;
      float (*swp)[64] = which ? sv_.rho : sv_.kap;
      for (int qq = 0; qq < kMaxQ; ++qq) {  // slot qq = 4t + k  <->  reference row q = t*r + k
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
        if (aux) aux[(which ? kAuxB : kAuxA) + qq * 64 + tok] = acc;
        if constexpr (BWD)
          if (which == 0) *reinterpret_cast<__nv_bfloat16*>(bv_.a_bf + tile_off(64, tok, qq)) = __float2bfloat16_rn(acc);
      }
      if (aux)   // feature means (rows c < V and 2V, 2V+1 are the ones in use)
        for (int idx = tid; idx < 2 * C * 64; idx += 128) {
          const int hf = idx / (C * 64), rem = idx % (C * 64);
          aux[(hf ? kAuxKap : kAuxRho) + rem] = (hf ? &sv_.kap[0][0] : &sv_.rho[0][0])[rem];
        }
    }
    __syncthreads();
    // =================================================================================================
    // stage 2c/3: mix the score maps, re-normalise
    // =================================================================================================
    float amix[32];
    float alo[kMaxQ], ahi[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) { alo[q] = sv_.a[q][f.row_lo]; ahi[q] = sv_.a[q][f.row_hi]; }
#pragma unroll
    for (int blk = 0; blk < 4; ++blk) {   // (kept unrolled: amix[] must stay in registers)
      float sv[kMaxV][8], fv[8];
#pragma unroll
      for (int i = 0; i < kMaxV; ++i)
        if (i < V) tmem_ld_16x256b_x2(tlane + ttile<BWD>(kTS + i) + 16 * blk, sv[i]);
      tmem_ld_16x256b_x2(tlane + ttile<BWD>(kTF) + 16 * blk, fv);
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
        amix[8 * blk + e] = s0 + fast_sigmoid(z[0]) * U + fast_sigmoid(z[1]) * O - fast_sigmoid(z[2]) * bn * U + fast_sigmoid(z[3]) * lf;
      }
    }
    {
      float st4[4];
      frag_softmax(amix, st4);
      put_stats(V, st4);
    }
    frag_store_bf16(tile(SL::AMIX), f, amix);

    if constexpr (!BWD) {
      // ===============================================================================================
      // y = A V_1 + w F V_V
      // ===============================================================================================
      publish();
      if (leader) {
        if (mine(0)) {   // both accumulate into the same tile: one issuer, in order
          gemm(kTY, 0, taddr(SL::AMIX), false, taddr(SL::V1), true, false, 4, 64);
          gemm(kTY, 0, sF, false, taddr(SL::VL), true, true, 4, 64);
        }
        mma_commit(&sv_.bar);
      }
      wait_mma();
      float yv[32];
      ld_tile(kTY, yv);
      __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y);
      const size_t row_lo = (((size_t)pb * 64 + f.row_lo) * H + ph) * dk;
      const size_t row_hi = (((size_t)pb * 64 + f.row_hi) * H + ph) * dk;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int c = f.col(n);
        if (c < dk) {
          *reinterpret_cast<uint32_t*>(y + row_lo + c) = pack_bf16(yv[4 * n], yv[4 * n + 1]);
          *reinterpret_cast<uint32_t*>(y + row_hi + c) = pack_bf16(yv[4 * n + 2], yv[4 * n + 3]);
        }
      }
    } else {
      using SB = Slots<true>;
      // ===============================================================================================
      // B1: D = A (.) (dA - rowsum(dA (.) A)); then per element: gate pre-activation grads dG_t (bf16
      //     tiles, operands of the da/db GEMMs), the direct part of dS_k (fp32, overwrites S_k in TMEM)
      //     and Hf = D g_chain / (F + eps) (overwrites dA: initial value of the dF accumulator)
      // ===============================================================================================
      float D[32];
      ld_tile(kTY, D);
      frag_softmax_bwd(D, amix);
      // da[q][i] = sum_j dG_t[i,j] b[q][j] is accumulated here in fp32: rows of D sum to zero, so this sum
      // cancels heavily and bf16-rounded dG terms (the MMA route) would not.
      float da_lo[kMaxQ], da_hi[kMaxQ];
#pragma unroll
      for (int q = 0; q < kMaxQ; ++q) { da_lo[q] = 0.f; da_hi[q] = 0.f; }
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk) {
        float sv[kMaxV][8], fv[8];
#pragma unroll
        for (int i = 0; i < kMaxV; ++i)
          if (i < V) tmem_ld_16x256b_x2(tlane + ttile<BWD>(kTS + i) + 16 * blk, sv[i]);
        tmem_ld_16x256b_x2(tlane + ttile<BWD>(kTF) + 16 * blk, fv);
        tmem_ld_wait();
        float hf[8], dgv[4][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = 16 * blk + 8 * (e >> 2) + f.cq + (e & 1);
          const bool hi = (e & 2) != 0;
          const float d = D[8 * blk + e];
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
          // direct part of dS_k replaces S_k
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
          if (i < V) tmem_st_16x256b_x2(tlane + ttile<BWD>(kTS + i) + 16 * blk, sv[i]);
        tmem_st_16x256b_x2(tlane + ttile<BWD>(kTY) + 16 * blk, hf);
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
        if ((f.lane & 3) == 0) { bv_.da[q][f.row_lo] = lo; bv_.da[q][f.row_hi] = hi; }
      }
      // ===============================================================================================
      // B2: dF += dY (w V_V)^T ; dV1 = A^T dY ; dVL = F^T dY ; db_t = dG_t^T a
      // ===============================================================================================
      publish();
      if (leader) {
        if (mine(0)) gemm(kTY, 0, taddr(SB::DY), false, taddr(SB::VL), false, true, ksteps, 64);
        if (mine(1)) gemm(kTdV1, 0, taddr(SB::AMIX), true, taddr(SB::DY), true, false, 4, 64);
        if (mine(2)) gemm(kTdVL, 0, sF, true, taddr(SB::DY), true, false, 4, 64);
        for (int t = 0; t < 4; ++t)
          if (mine(3 + t)) gemm(kTdb, 16 * t, taddr(SB::X + t), true, smem_u32(bv_.a_bf), true, false, 4, 16);
        mma_commit(&sv_.bar);
      }
      wait_mma();
      __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(p.dqkv);
      const size_t in_lo = (((size_t)pb * 64 + f.row_lo) * 3) * hd + (size_t)ph * dk;   // q row; +hd: k; +2hd: v
      const size_t in_hi = (((size_t)pb * 64 + f.row_hi) * 3) * hd + (size_t)ph * dk;
      {
        // value gradients and the v_scale partials
        float d1[32], dl[32];
        ld_tile(kTdV1, d1);
        ld_tile(kTdVL, dl);
        float p1[32], pl[32];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const int c = f.col(n);
          float2 vlo = make_float2(0.f, 0.f), vhi = vlo;
          if (c < dk) {
            vlo = unpack_bf16(*reinterpret_cast<const uint32_t*>(qkv + in_lo + 2 * hd + c));
            vhi = unpack_bf16(*reinterpret_cast<const uint32_t*>(qkv + in_hi + 2 * hd + c));
            const float a0 = sv_.vs1[c], a1 = sv_.vs1[c + 1], b0 = sv_.vsL[c], b1 = sv_.vsL[c + 1];
            *reinterpret_cast<uint32_t*>(dqkv + in_lo + 2 * hd + c) =
                pack_bf16(d1[4 * n] * a0 + dl[4 * n] * b0, d1[4 * n + 1] * a1 + dl[4 * n + 1] * b1);
            *reinterpret_cast<uint32_t*>(dqkv + in_hi + 2 * hd + c) =
                pack_bf16(d1[4 * n + 2] * a0 + dl[4 * n + 2] * b0, d1[4 * n + 3] * a1 + dl[4 * n + 3] * b1);
          }
          p1[4 * n] = d1[4 * n] * vlo.x; p1[4 * n + 1] = d1[4 * n + 1] * vlo.y;
          p1[4 * n + 2] = d1[4 * n + 2] * vhi.x; p1[4 * n + 3] = d1[4 * n + 3] * vhi.y;
          pl[4 * n] = dl[4 * n] * vlo.x; pl[4 * n + 1] = dl[4 * n + 1] * vlo.y;
          pl[4 * n + 2] = dl[4 * n + 2] * vhi.x; pl[4 * n + 3] = dl[4 * n + 3] * vhi.y;
        }
        colsum_to(sv_.red[0], f, p1);
        colsum_to(sv_.red[1], f, pl);
      }
      {
        // column-factor gradients: useful columns of accumulator block t are slots 4t..4t+3
        float vb[32];
        ld_tile(kTdb, vb);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = f.col(n) + (e & 1), t = col >> 4, within = col & 15;
            if ((within >> 2) == t) sv_.b[within][(e & 2) ? f.row_hi : f.row_lo] = vb[4 * n + e];   // db[q][j]
          }
        }
      }
      __syncthreads();
      // chain_value_logit: w(1-w) <dY, F V_V> = (1-w) sum_d (w vs_V[d]) sum_m (F^T dY)[m,d] V[m,d]  - the same
      // column sums as the v_scale[V-1] gradient (fp32 accumulators x exact V: no operand rounding in the
      // heavily cancelling sum)
      if (tid < 32) {
        float s = 0.f;
        for (int d = tid; d < dk; d += 32) s = fmaf(sv_.vsL[d], sv_.red[1][0][d] + sv_.red[1][1][d] + sv_.red[1][2][d] + sv_.red[1][3][d], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) p.dlogit_part[g] = (1.f - w) * s;
      }
      // v_scale partials (views 0 and V-1), feature-mean grads, gate-head parameter partials
      if (p.dscale_part) {
        float* ds = p.dscale_part + (size_t)g * 3 * V * dk + (size_t)2 * V * dk;
        for (int idx = tid; idx < V * dk; idx += 128) {
          const int k = idx / dk, d = idx % dk;
          float val = 0.f;
          if (k == 0) val = sv_.red[0][0][d] + sv_.red[0][1][d] + sv_.red[0][2][d] + sv_.red[0][3][d];
          if (k == V - 1) val += w * (sv_.red[1][0][d] + sv_.red[1][1][d] + sv_.red[1][2][d] + sv_.red[1][3][d]);
          ds[idx] = val;
        }
      }
      for (int idx = tid; idx < C * 64; idx += 128) {
        const int c = idx >> 6, tok = idx & 63;
        float sr = 0.f, sc = 0.f;
        for (int qq = 0; qq < kMaxQ; ++qq) {
          const int t = qq >> 2, k = qq & 3;
          if (k < r) {
            const int q = t * r + k;
            sr = fmaf(sv_.hw[0][q * C + c], bv_.da[qq][tok], sr);
            sc = fmaf(sv_.hw[1][q * C + c], sv_.b[qq][tok], sc);
          }
        }
        bv_.drho[c][tok] = sr * (1.f / 64.f);
        bv_.dkap[c][tok] = sc * (1.f / 64.f);
      }
      {
        const int nW = 4 * r * C, nP = nW + 4 * r;
        float* dh = p.dhead_part + (size_t)g * 2 * nP;
        for (int idx = tid; idx < 2 * nP; idx += 128) {
          const int half = idx / nP, rem = idx % nP;
          float (*dv)[64] = half ? sv_.b : bv_.da;
          float s = 0.f;
          if (rem < nW) {
            const int q = rem / C, c = rem % C, qq = 4 * (q / r) + (q % r);
            // feature of channel c as seen by the row (half 0) / column (half 1) projection
            const float* ft;
            if (c < V) ft = half ? sv_.kap[c] : sv_.rho[c];
            else if (c < 2 * V) ft = half ? sv_.rho[c - V] : sv_.kap[c - V];
            else ft = half ? sv_.kap[c] : sv_.rho[c];
            for (int i = 0; i < 64; ++i) s = fmaf(dv[qq][i], ft[i], s);
          } else {
            const int q = rem - nW, qq = 4 * (q / r) + (q % r);
            for (int i = 0; i < 64; ++i) s += dv[qq][i];
          }
          dh[idx] = s;
        }
      }
      __syncthreads();
      // ===============================================================================================
      // B3: chain seeds X_F = dF + dfeat_{2V}/(F+eps), X_R = dfeat_{2V+1}/(R+eps); feature terms of dS_k
      // ===============================================================================================
      {
        float x[32], den[32];
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
        ld_tile(kTR, den);
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = f.col(n) + (e & 1), row = (e & 2) ? f.row_hi : f.row_lo;
            x[4 * n + e] = (bv_.drho[2 * V + 1][row] + bv_.dkap[2 * V + 1][col]) * fast_rcp(den[4 * n + e] + p.eps);
          }
        frag_store_bf16(tile(SB::X + 2), f, x);
        for (int k = 0; k < V; ++k) {
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
      // ===============================================================================================
      // B4: chain sweep.  F = A_0..A_{V-1}: dA_k += P_{k-1}^T X, X <- X A_k^T for k = V-1..1, dA_0 += X.
      //                   R = A_{V-1}..A_0: dA_k += R_{k+1}^T X, X <- X A_k^T for k = 0..V-2, dA_{V-1} += X.
      //     Both sweeps run in lock step; each contribution goes through the softmax backward of A_k
      //     (linear in dA_k) and is accumulated into the fp32 dS_k tile.
      // ===============================================================================================
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
            if (mine(2)) gemm(kTAR, 0, rNext, true, taddr(xr), true, false, 4, 64);
            if (mine(3)) gemm(kTXR, 0, taddr(xr), false, taddr(SB::A + kR), false, false, 4, 64);
            mma_commit(&sv_.bar);
          }
          wait_mma();
          float x[32], pk[32], acc[32];
          ld_tile(kTAF, x);
          frag_load_bf16(tile(SB::A + kF), f, pk);
          frag_softmax_bwd(x, pk);
          ld_tile(kTS + kF, acc);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] += x[i];
          st_tile(kTS + kF, acc);
          ld_tile(kTAR, x);
          frag_load_bf16(tile(SB::A + kR), f, pk);
          frag_softmax_bwd(x, pk);
          ld_tile(kTS + kR, acc);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] += x[i];
          st_tile(kTS + kR, acc);
          if (s < V - 2) {
            xf = (xf == SB::X) ? SB::X + 1 : SB::X;            // ping-pong (X+0, X+1) and (X+2, X+3)
            xr = (xr == SB::X + 2) ? SB::X + 3 : SB::X + 2;
            ld_tile(kTXF, x);
            frag_store_bf16(tile(xf), f, x);
            ld_tile(kTXR, x);
            frag_store_bf16(tile(xr), f, x);
          }
        }
        // last links: dA_0 += X_F (fp32 accumulator of the last step), dA_{V-1} += X_R
        float x[32], pk[32], acc[32];
        ld_tile(kTXF, x);
        frag_load_bf16(tile(SB::A + 0), f, pk);
        frag_softmax_bwd(x, pk);
        ld_tile(kTS + 0, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += x[i];
        st_tile(kTS + 0, acc);
        ld_tile(kTXR, x);
        frag_load_bf16(tile(SB::A + V - 1), f, pk);
        frag_softmax_bwd(x, pk);
        ld_tile(kTS + V - 1, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += x[i];
        st_tile(kTS + V - 1, acc);
      }
      // ===============================================================================================
      // B5: dS_k -> bf16 operand tiles; T_k = dS_k K, U_k = dS_k^T Q; dQ, dK, scale partials
      // ===============================================================================================
      // dS_k tiles: slots 1..4 (V1, VL, dY, Amix are dead) and X+0; unscaled Q is reloaded into X+1
      auto ds_slot = [&](int k) { return k < 4 ? 1 + k : SB::X + 0; };
      for (int k = 0; k < V; ++k) {
        float x[32];
        ld_tile(kTS + k, x);
        frag_store_bf16(tile(ds_slot(k)), f, x);
      }
      for (int idx = tid; idx < 64 * 8; idx += 128) {
        const int rr = idx & 63, ch = idx >> 6;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (ch * 8 < dk) q = *reinterpret_cast<const uint4*>(qkv + (((size_t)pb * 64 + rr) * 3) * hd + (size_t)ph * dk + ch * 8);
        *reinterpret_cast<uint4*>(tile(SB::X + 1) + ch * 1024 + rr * 16) = q;
      }
      publish();
      if (leader) {
        for (int k = 0; k < V; ++k) {
          if (mine(2 * k)) gemm(kTT + k, 0, taddr(ds_slot(k)), false, taddr(SB::K), true, false, 4, 64);
          if (mine(2 * k + 1)) gemm(tileU(k), 0, taddr(ds_slot(k)), true, taddr(SB::X + 1), true, false, 4, 64);
        }
        mma_commit(&sv_.bar);
      }
      wait_mma();
      {
        float dq[32], dkk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { dq[i] = 0.f; dkk[i] = 0.f; }
        float qf[32];
        frag_load_bf16(tile(SB::X + 1), f, qf);
        for (int k = 0; k < V; ++k) {
          float t[32], z[32];
          ld_tile(kTT + k, t);
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float c0 = sv_.cvec[k][f.col(n)], c1 = sv_.cvec[k][f.col(n) + 1];
            dq[4 * n] = fmaf(t[4 * n], c0, dq[4 * n]); dq[4 * n + 1] = fmaf(t[4 * n + 1], c1, dq[4 * n + 1]);
            dq[4 * n + 2] = fmaf(t[4 * n + 2], c0, dq[4 * n + 2]); dq[4 * n + 3] = fmaf(t[4 * n + 3], c1, dq[4 * n + 3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) z[4 * n + e] = t[4 * n + e] * qf[4 * n + e];
          }
          colsum_to(sv_.red[k], f, z);
          ld_tile(tileU(k), t);
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float c0 = sv_.cvec[k][f.col(n)], c1 = sv_.cvec[k][f.col(n) + 1];
            dkk[4 * n] = fmaf(t[4 * n], c0, dkk[4 * n]); dkk[4 * n + 1] = fmaf(t[4 * n + 1], c1, dkk[4 * n + 1]);
            dkk[4 * n + 2] = fmaf(t[4 * n + 2], c0, dkk[4 * n + 2]); dkk[4 * n + 3] = fmaf(t[4 * n + 3], c1, dkk[4 * n + 3]);
          }
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const int c = f.col(n);
          if (c < dk) {
            *reinterpret_cast<uint32_t*>(dqkv + in_lo + c) = pack_bf16(dq[4 * n], dq[4 * n + 1]);
            *reinterpret_cast<uint32_t*>(dqkv + in_hi + c) = pack_bf16(dq[4 * n + 2], dq[4 * n + 3]);
            *reinterpret_cast<uint32_t*>(dqkv + in_lo + hd + c) = pack_bf16(dkk[4 * n], dkk[4 * n + 1]);
            *reinterpret_cast<uint32_t*>(dqkv + in_hi + hd + c) = pack_bf16(dkk[4 * n + 2], dkk[4 * n + 3]);
          }
        }
      }
      __syncthreads();
      if (p.dscale_part) {
        float* ds = p.dscale_part + (size_t)g * 3 * V * dk;
        for (int idx = tid; idx < V * dk; idx += 128) {
          const int k = idx / dk, d = idx % dk;
          const float z = sscale * (sv_.red[k][0][d] + sv_.red[k][1][d] + sv_.red[k][2][d] + sv_.red[k][3][d]);
          const size_t pi = ((size_t)k * H + ph) * dk + d;
          ds[idx] = p.k_scale[pi] * z;
          ds[(size_t)V * dk + idx] = p.q_scale[pi] * z;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // tiles, vectors and TMEM are reused by the next problem
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TmemCols<BWD>::value>(tbase);
}

inline bool supported(const MopEdgewiseParams* p) {
  return p->dtype == MOP_BF16 && p->N == 64 && p->dk <= 64 && p->dk % 8 == 0 && p->V >= 2 && p->V <= kMaxV && p->Vp == 1 &&
         p->gate_mode == MOP_GATE_LOWRANK && p->gate_rank >= 1 && p->gate_rank <= 4 && p->q_scale != nullptr;
}

}  // namespace ewtc
}  // namespace mop
