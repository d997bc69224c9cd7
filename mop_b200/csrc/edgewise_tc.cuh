// Edgewise (Mixture-of-Products) attention core on tcgen05 / TMEM: building blocks shared by the N = 64 kernels
// (edgewise_n64_fwd.cuh, edgewise_n64_bwd.cuh) and the N <= 200 kernels (edgewise_tc_large*.cuh): tile / TMEM maps, fragment
// helpers, the layout of the forward -> backward `aux` vectors.  The first-generation N = 64 kernels that lived here
// (128 / 256 threads, everything recomputed in the backward) were replaced in round 2.
// bf16 operands, fp32 accumulation and fp32 statistics.  Hot shape of configs 1/2:
// N = 64 tokens, dk <= 64 (dk % 8 == 0), V <= 5 shared-projection views, low-rank gate head, r <= 4.
//
// One CTA owns one (batch, head) problem at a time (persistent loop over problems):
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

inline bool supported(const MopEdgewiseParams* p) {
  return p->dtype == MOP_BF16 && p->N == 64 && p->dk <= 64 && p->dk % 8 == 0 && p->V >= 2 && p->V <= kMaxV && p->Vp == 1 &&
         p->gate_mode == MOP_GATE_LOWRANK && p->gate_rank >= 1 && p->gate_rank <= 4 && p->q_scale != nullptr && p->lens_n == 0;
}

}  // namespace ewtc
}  // namespace mop
