// GPT "Quartet" causal attention on tcgen05 / TMEM (bf16 operands, fp32 accumulation and statistics).
//
// Replaces quartet_attn_patch.py:88-121 for bf16 activations with dk <= 64 (dk % 8 == 0): two score maps, each
// z-scored per row over the FULL row with the unbiased std, mixed as (1-m) n1 + m gamma n1 n2, causal fill,
// optional additive mask, softmax, PV (`use_quartet = 0`: the single z-scored map).  Same algebra as the fp32-mode
// kernels (quartet_simt.cuh): with kc_j = k_j - mean_j k_j the centred score is
// c_ij = s q_i . kc_j, and sigma_i^2 = s^2 q_i^T (Kc^T Kc) q_i / (T-1), so the forward needs only the CAUSAL half of
// both score maps plus a dk x dk Gram matrix per (batch, head, map); the backward adds two Gram-type corrections.
//
// Kernels (M=128 accumulators, one thread per accumulator row = tcgen05.ld 32x32b):
//   prep      per (b,h,map): kc (bf16, the operand the MMAs see), G = Kc^T Kc from that rounded kc (as bf16 hi + lo tiles)
//   fwd       per (b,h,128 queries): sigma_i^2 = s^2 q_i.(G q_i)/(T-1) with G q by MMA; then stream 64-key tiles (cp.async, double
//             buffered) up to the diagonal: S1, S2 by MMA, mix + online softmax in
//             registers (thread per query row), P V accumulated in TMEM
//   bwd_dq    per (b,h,128 queries): S1, S2, dP = dO V^T by MMA; W1, W2 (bf16) -> dQ1 += W1 Kc1, dQ2 += W2 Kc2 in TMEM;
//             row coefficients g_i, the Gram correction and the two scalar partials
//   gmat      per (b,h,map): M = sum_i g_i q_i q_i^T
//   bwd_dkdv  per (b,h,128 keys), transposed tiles (thread per key row, per-query statistics broadcast from shared
//             memory): dV += P^T dO, dKc1 += W1^T Q, dKc2 += W2^T Q2 in TMEM, then the M correction
//   finish    dk = dkc - mean_j dkc; scalar partials
// Every output is written by exactly one CTA (no atomics); all buffers are the caller's.
#pragma once
#include "quartet_simt.cuh"
#include "tc_common.cuh"

// -DMOP_FWD_TIMELINE: one CTA of the forward kernel prints clock64 stamps of its roles per tile (tools/ts_timeline.py); development only
#ifdef MOP_FWD_TIMELINE
#define TS_DECL long long ts_[24][6]; const bool ts_on = blockIdx.x == MOP_FWD_TIMELINE;
#define TS(t, k) do { if (ts_on && (t) < 24) ts_[t][k] = clock64(); } while (0)
#define TS_DUMP(name, n, K) do { if (ts_on) for (int t_ = 0; t_ < (n) && t_ < 24; ++t_) printf("%s t= %d %lld %lld %lld %lld %lld %lld\n", name, t_, ts_[t_][0], ts_[t_][1], ts_[t_][2], ts_[t_][3], K > 4 ? ts_[t_][4] : 0LL, K > 5 ? ts_[t_][5] : 0LL); } while (0)
#else
#define TS_DECL
#define TS(t, k)
#define TS_DUMP(name, n, K)
#endif
namespace mop {
namespace qtc {

using namespace tc;
using quartet::at;
using quartet::load_mix;
using quartet::Mix;

constexpr int kT128 = 128 * 16 * 8;   // [128 x 64] bf16 tile image, R = 128
constexpr int kT64 = 64 * 16 * 8;     // [64 x 64] bf16 tile image, R = 64
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct Ws {  // byte offsets into the workspace
  size_t kc, gram, ksum, gacc, gvec, mmat, macc, dkc, dksum, spart, delta, total;
  int nm, nqb;
};
__host__ __device__ inline Ws layout(const MopQuartetParams* p, int backward) {
  Ws w;
  const size_t BH = (size_t)p->B * p->H, T = p->T;
  w.nm = p->use_quartet ? 2 : 1;
  w.nqb = (p->T + 127) / 128;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
  w.kc = take(w.nm * BH * T * 64 * 2);       // bf16 [nm][BH][T][64] (columns >= dk are zero)
  w.gram = take(w.nm * BH * 2 * kT64);       // bf16 hi / lo tile images of Kc^T Kc: [nm][BH][2][8 KB]
  w.ksum = take(w.nm * BH * 64 * 4);         // fp32 column sums of k          } zeroed by one memset before the prep kernels,
  w.gacc = take(w.nm * BH * 64 * 64 * 4);    // fp32 Kc^T Kc accumulators      } accumulated with atomics by their row groups
  w.gvec = w.mmat = w.macc = w.dkc = w.dksum = w.spart = w.delta = 0;
  if (backward) {
    w.gvec = take(w.nm * BH * T * 4);
    w.mmat = take(w.nm * BH * 2 * kT64);     // bf16 hi / lo tile images of sum_i g_i q_i q_i^T
    w.macc = take(w.nm * BH * 64 * 64 * 4);  // fp32 accumulators of the same when gmat is split over row groups (memset, atomics)
    w.dkc = take(w.nm * BH * T * 64 * 4);    // fp32 [nm][BH][T][64]
    w.dksum = take(w.nm * BH * 64 * 4);      // fp32 [nm][BH][64]: column sums of dkc (zeroed by prep, accumulated by bwd_dkdv)
    w.spart = take(BH * w.nqb * 2 * 4);
    w.delta = take(BH * T * 4);              // fp32 [BH][T]: dO . y per query (bwd_dq -> bwd_dkdv)
  }
  w.total = o;
  return w;
}

__device__ __forceinline__ void publish() {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}
__device__ __forceinline__ void unpack8(uint4 u, float* f) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}

// fp32 64x64 matrix (shared) -> two bf16 chunk-major tile images hi, lo with hi + lo = G to ~2^-17 (global, 2 x 8 KB).
// The symmetric matrices Kc^T Kc and sum g q q^T enter the quadratic forms through two MMAs each (X hi + X lo).
__device__ inline void write_hilo_tiles(const float* G, unsigned char* dst) {
  for (int idx = threadIdx.x; idx < 64 * 8; idx += blockDim.x) {
    const int r = idx & 63, ch = idx >> 6;
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g = G[r * 64 + ch * 8 + e];
      hi[e] = __bfloat162float(__float2bfloat16_rn(g));
      lo[e] = g - hi[e];
    }
    *reinterpret_cast<uint4*>(dst + ch * 1024 + r * 16) = pack8(hi);
    *reinterpret_cast<uint4*>(dst + kT64 + ch * 1024 + r * 16) = pack8(lo);
  }
}
// n (<= 4) 8 KB tile images global -> shared, 256 threads: every load is issued before the first store, so the CTA pays one
// L2 round trip (a copy loop per tile paid eight: these tiles are read once per CTA, at the end of a short kernel)
__device__ __forceinline__ void copy_tiles64(unsigned char* const* dst, const unsigned char* const* src, int n) {
  uint4 v[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (i < n) v[i][j] = __ldg(reinterpret_cast<const uint4*>(src[i] + (threadIdx.x + 256 * j) * 16));
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (i < n) *reinterpret_cast<uint4*>(dst[i] + (threadIdx.x + 256 * j) * 16) = v[i][j];
}
// Z[128 x 64] at TMEM column dcol = X (Mhi + Mlo): X a [128 x 64] K-major tile, M symmetric (tile images in shared memory)
__device__ __forceinline__ void mma_x_sym(uint32_t d_tmem, uint32_t xtile, uint32_t mhi, uint32_t mlo) {
  const uint32_t id = idesc_bf16(128, 64, 0, 0);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) mma_ss(d_tmem, desc_k_sw(xtile, 16 * ks), desc_kmajor(mhi, 64, 16 * ks), id, ks > 0 ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) mma_ss(d_tmem, desc_k_sw(xtile, 16 * ks), desc_kmajor(mlo, 64, 16 * ks), id, 1u);
}

// rows [row0, row0 + R) of a [.., T, H, dk] bf16 activation (token stride `stride` elements) -> chunk-major tile, zero fill
template <int R>
__device__ __forceinline__ void load_act_tile(unsigned char* tile, const __nv_bfloat16* base, size_t stride, int row0, int T, int dk) {
  for (int idx = threadIdx.x; idx < R * 8; idx += blockDim.x) {
    const int r = idx % R, ch = idx / R, n = row0 + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < T && ch * 8 < dk) v = *reinterpret_cast<const uint4*>(base + (size_t)n * stride + ch * 8);
    *reinterpret_cast<uint4*>(tile + ch * (R * 16) + r * 16) = v;
  }
}

// ---- 64 x 64 Gram-type matrices on the tensor core ----------------------------------------------------------
// G = sum over 64-row chunks of A_chunk^T B_chunk with M = N = 64 (MN-major operands: the chunk's rows are the K
// index), accumulated in TMEM (M=64 accumulator: lanes 0-15 of each sub-partition, read as 16x256b fragments).
// The caller stages each chunk's operand tiles (chunk-major, R = 64) into one of kGramBufs shared buffers.
constexpr int kGramBufs = 2;
struct GramPipe {
  uint64_t bar[kGramBufs];
  uint32_t tmem_slot;
  uint32_t phase_bits;   // per-thread copy kept in a register by the caller
};
// after the last chunk: accumulator -> fp32 G[64][64] in shared memory (threads 0..127)
__device__ inline void gram_readout(uint32_t tb, float* G) {
  if (threadIdx.x < 128) {
    const Frag f;
    float v[32];
    tmem_ld_16x256b_x8(tb + ((uint32_t)(32 * f.warp) << 16), v);
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      G[f.row_lo * 64 + f.col(n)] = v[4 * n];
      G[f.row_lo * 64 + f.col(n) + 1] = v[4 * n + 1];
      G[f.row_hi * 64 + f.col(n)] = v[4 * n + 2];
      G[f.row_hi * 64 + f.col(n) + 1] = v[4 * n + 3];
    }
  }
}

// The key-side preparation is split over row groups of kPrepRows keys so that it fills the GPU at every (B, T) - one CTA per
// (b, h, map) left 96 CTAs with 64 serial chunks each at T = 4096, B = 4:
//   prep_sum_kernel    grid (B*H*nm) * groups: column sums of k -> atomics into w.ksum
//   prep_kernel        grid (B*H*nm) * groups: kc = bf16(k - kbar) (workspace), partial Kc^T Kc of the group's rounded rows on
//                      the tensor core -> atomics into w.gacc
//   acc_to_tiles_kernel  grid B*H*nm: w.gacc -> bf16 hi / lo tile images
constexpr int kPrepRows = 256;
static __global__ void __launch_bounds__(256) prep_sum_kernel(MopQuartetParams p, Ws w, unsigned char* ws) {
  __shared__ float part[32][64];
  const int nm = w.nm, T = p.T, ngr = (T + kPrepRows - 1) / kPrepRows, dk = p.dk, tid = threadIdx.x;
  const int grp = blockIdx.x % ngr, bm = blockIdx.x / ngr, bh = bm / nm, map = bm % nm, b = bh / p.H, h = bh % p.H;
  const size_t BH = (size_t)p.B * p.H, stride = (size_t)p.H * dk;
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(map ? p.k2 : p.k) + at(p, b, 0, h);
  const int r0 = tid >> 3, ch = tid & 7;   // rows r0 + 32 j of the group, the 8 columns of chunk ch
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  uint4 u[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {   // eight independent 16-byte loads in flight
    const int t = grp * kPrepRows + r0 + 32 * j;
    u[j] = (t < T && ch * 8 < dk) ? *reinterpret_cast<const uint4*>(k + (size_t)t * stride + ch * 8) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float f[8];
    unpack8(u[j], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] += f[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[r0][ch * 8 + e] = s[e];
  __syncthreads();
  if (tid < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += part[i][tid];
    atomicAdd(reinterpret_cast<float*>(ws + w.ksum) + ((size_t)map * BH + bh) * 64 + tid, t);
  }
}

static __global__ void __launch_bounds__(256) prep_kernel(MopQuartetParams p, Ws w, unsigned char* ws) {
  __shared__ __align__(128) unsigned char tiles[kGramBufs][kT64];
  __shared__ __align__(16) float G[64 * 64];
  __shared__ float kbar[64];
  __shared__ GramPipe gp;
  const int nm = w.nm, T = p.T, ngr = (T + kPrepRows - 1) / kPrepRows, dk = p.dk, tid = threadIdx.x;
  const int grp = blockIdx.x % ngr, bm = blockIdx.x / ngr, bh = bm / nm, map = bm % nm, b = bh / p.H, h = bh % p.H;
  const size_t BH = (size_t)p.B * p.H, stride = (size_t)p.H * dk;
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(map ? p.k2 : p.k) + at(p, b, 0, h);
  __nv_bfloat16* kc = reinterpret_cast<__nv_bfloat16*>(ws + w.kc) + ((size_t)map * BH + bh) * T * 64;
  if (tid < 32) tmem_alloc<64>(&gp.tmem_slot);
  if (tid == 0) { for (int i = 0; i < kGramBufs; ++i) mbar_init(&gp.bar[i], 1); fence_mbar_init(); }
  if (w.dksum && grp == 0 && tid < 64) (reinterpret_cast<float*>(ws + w.dksum) + ((size_t)map * BH + bh) * 64)[tid] = 0.f;   // backward only
  if (tid < 64) kbar[tid] = (reinterpret_cast<const float*>(ws + w.ksum) + ((size_t)map * BH + bh) * 64)[tid] / (float)T;
  const int r0 = tid >> 3, ch = tid & 7;   // this thread: rows r0, r0 + 32 of a chunk and the 8 columns of chunk ch
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = gp.tmem_slot;
  uint32_t phases = 0;
  const int c_lo = grp * (kPrepRows / 64), nchunks = min(kPrepRows / 64, ((T + 63) >> 6) - c_lo);   // this group's 64-row chunks
  auto load_rows = [&](int c, uint4* u) {   // this thread's two 16-byte pieces of chunk c (zero outside the tensor / the group)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int t = (c_lo + c) * 64 + r0 + 32 * it;
      u[it] = (c < nchunks && t < T && ch * 8 < dk) ? *reinterpret_cast<const uint4*>(k + (size_t)t * stride + ch * 8) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 cur[2], nxt[2];
  load_rows(0, cur);
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c % kGramBufs;
    load_rows(c + 1, nxt);   // in flight while this chunk is converted, staged and multiplied
    if (c >= kGramBufs) { mbar_wait(&gp.bar[buf], (phases >> buf) & 1u); phases ^= 1u << buf; }   // the MMAs that read this buffer are done
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = r0 + 32 * it, t = (c_lo + c) * 64 + r;
      float f[8];
      unpack8(cur[it], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = (t < T && ch * 8 + e < dk) ? f[e] - kbar[ch * 8 + e] : 0.f;
      const uint4 u = pack8(f);
      if (t < T) *reinterpret_cast<uint4*>(kc + (size_t)t * 64 + ch * 8) = u;
      *reinterpret_cast<uint4*>(tiles[buf] + ch * 1024 + r * 16) = u;
    }
    publish();
    if (tid == 0) {
      const uint32_t id = idesc_bf16(64, 64, 1, 1), tl = smem_u32(tiles[buf]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tb, desc_mnmajor(tl, 64, 16 * ks), desc_mnmajor(tl, 64, 16 * ks), id, (c > 0 || ks > 0) ? 1u : 0u);
      mma_commit(&gp.bar[buf]);
    }
    cur[0] = nxt[0]; cur[1] = nxt[1];
  }
  {   // the last commit covers every MMA issued before it
    const int buf = (nchunks - 1) % kGramBufs;
    mbar_wait(&gp.bar[buf], (phases >> buf) & 1u);
    tc_fence_after();
  }
  gram_readout(tb, G);
  tc_fence_before();
  __syncthreads();
  float* acc = reinterpret_cast<float*>(ws + w.gacc) + ((size_t)map * BH + bh) * 64 * 64;
  for (int i = tid; i < 64 * 64; i += 256) atomicAdd(acc + i, G[i]);
  if (tid < 32) tmem_dealloc<64>(tb);
}

// fp32 64x64 accumulators (byte offset acc_off) -> bf16 hi / lo tile images (byte offset tiles_off); blockIdx.x = map * BH + bh
static __global__ void __launch_bounds__(256) acc_to_tiles_kernel(unsigned char* ws, size_t acc_off, size_t tiles_off) {
  __shared__ __align__(16) float G[64 * 64];
  const float* acc = reinterpret_cast<const float*>(ws + acc_off) + (size_t)blockIdx.x * 64 * 64;
  for (int i = threadIdx.x; i < 64 * 64 / 4; i += 256) reinterpret_cast<float4*>(G)[i] = reinterpret_cast<const float4*>(acc)[i];
  __syncthreads();
  write_hilo_tiles(G, ws + tiles_off + (size_t)blockIdx.x * 2 * kT64);
}

// One CTA per (b, h, map): used when B*H*nm alone fills the GPU (fewer launches, no atomics).
// grid: B*H*nm, 256 threads.  kbar, kc = bf16(k - kbar) (written to the workspace and staged for the MMA),
// G = Kc^T Kc from that rounded kc -> bf16 hi / lo tile images.
static __global__ void __launch_bounds__(256) prep_fused_kernel(MopQuartetParams p, Ws w, unsigned char* ws) {
  __shared__ __align__(128) unsigned char tiles[kGramBufs][kT64];
  __shared__ __align__(16) float G[64 * 64];
  __shared__ float part[32][64];
  __shared__ float kbar[64];
  __shared__ GramPipe gp;
  const int nm = w.nm, bh = blockIdx.x / nm, map = blockIdx.x % nm, b = bh / p.H, h = bh % p.H, dk = p.dk, T = p.T, tid = threadIdx.x;
  const size_t BH = (size_t)p.B * p.H, stride = (size_t)p.H * dk;
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(map ? p.k2 : p.k) + at(p, b, 0, h);
  __nv_bfloat16* kc = reinterpret_cast<__nv_bfloat16*>(ws + w.kc) + ((size_t)map * BH + bh) * T * 64;
  if (tid < 32) tmem_alloc<64>(&gp.tmem_slot);
  if (tid == 0) { for (int i = 0; i < kGramBufs; ++i) mbar_init(&gp.bar[i], 1); fence_mbar_init(); }
  if (w.dksum && tid < 64) (reinterpret_cast<float*>(ws + w.dksum) + ((size_t)map * BH + bh) * 64)[tid] = 0.f;   // backward only
  const int r0 = tid >> 3, ch = tid & 7;   // this thread: rows r0, r0 + 32, ... and the 8 columns of chunk ch
  {
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    if (ch * 8 < dk)
      for (int t = r0; t < T; t += 128) {   // four independent 16-byte loads in flight per thread (the loop is latency bound)
        uint4 u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          u[j] = t + 32 * j < T ? *reinterpret_cast<const uint4*>(k + (size_t)(t + 32 * j) * stride + ch * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float f[8];
          unpack8(u[j], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) s[e] += f[e];
        }
      }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[r0][ch * 8 + e] = s[e];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < 64) {
    float s = 0.f;
    for (int i = 0; i < 32; ++i) s += part[i][tid];
    kbar[tid] = s / (float)T;
  }
  __syncthreads();
  const uint32_t tb = gp.tmem_slot;
  uint32_t phases = 0;
  const int nchunks = (T + 63) >> 6;
  auto load_rows = [&](int c, uint4* u) {   // this thread's two 16-byte pieces of chunk c (zero outside the tensor)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int t = c * 64 + r0 + 32 * it;
      u[it] = (c < nchunks && t < T && ch * 8 < dk) ? *reinterpret_cast<const uint4*>(k + (size_t)t * stride + ch * 8) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 cur[2], nxt[2];
  load_rows(0, cur);
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c % kGramBufs;
    load_rows(c + 1, nxt);   // in flight while this chunk is converted, staged and multiplied
    if (c >= kGramBufs) { mbar_wait(&gp.bar[buf], (phases >> buf) & 1u); phases ^= 1u << buf; }   // the MMAs that read this buffer are done
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = r0 + 32 * it, t = c * 64 + r;
      float f[8];
      unpack8(cur[it], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = (t < T && ch * 8 + e < dk) ? f[e] - kbar[ch * 8 + e] : 0.f;
      const uint4 u = pack8(f);
      if (t < T) *reinterpret_cast<uint4*>(kc + (size_t)t * 64 + ch * 8) = u;
      *reinterpret_cast<uint4*>(tiles[buf] + ch * 1024 + r * 16) = u;
    }
    publish();
    if (tid == 0) {
      const uint32_t id = idesc_bf16(64, 64, 1, 1), tl = smem_u32(tiles[buf]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tb, desc_mnmajor(tl, 64, 16 * ks), desc_mnmajor(tl, 64, 16 * ks), id, (c > 0 || ks > 0) ? 1u : 0u);
      mma_commit(&gp.bar[buf]);
    }
    cur[0] = nxt[0]; cur[1] = nxt[1];
  }
  {   // the last commit covers every MMA issued before it
    const int buf = (nchunks - 1) % kGramBufs;
    mbar_wait(&gp.bar[buf], (phases >> buf) & 1u);
    tc_fence_after();
  }
  gram_readout(tb, G);
  tc_fence_before();
  __syncthreads();
  write_hilo_tiles(G, ws + w.gram + ((size_t)map * BH + bh) * 2 * kT64);
  if (tid < 32) tmem_dealloc<64>(tb);
}

// mixed score of one element from the raw MMA dot products (natural units); masks applied by the caller
__device__ __forceinline__ float mix_n(const Mix& x, float n1, float n2) { return x.quart ? n1 * ((1.f - x.m) + x.m * x.gam * n2) : n1; }

template <int R>
__device__ __forceinline__ void load_act_tile_async(unsigned char* tile, const __nv_bfloat16* base, size_t stride, int row0, int T, int dk) {
  for (int idx = threadIdx.x; idx < R * 8; idx += blockDim.x) {
    const int r = idx % R, ch = idx / R, n = row0 + r;
    const bool ok = n < T && ch * 8 < dk;
    cp_async16(tile + ch * (R * 16) + r * 16, base + (size_t)(ok ? n : 0) * stride + (ok ? ch * 8 : 0), ok);
  }
}

constexpr int kKStages = 3;   // key ring of the forward kernel (the value ring has two stages)
struct __align__(128) SmemF {
  unsigned char Q[kT128], Q2[kT128], P[kT128];
  unsigned char K1[kKStages][kT64], K2[kKStages][kT64], V[2][kT64];
  uint64_t bar_s;        // S1 and S2 of a tile complete (two issuing lanes -> two arrivals); phase 0 = the sigma prologue
  uint64_t bar_pv;       // P V of a tile complete                                   (tensor pipe -> softmax warps, lane B)
  uint64_t p_ready;      // P(t) written: 128 arrivals                                (softmax warps -> lane B)
  uint64_t s_free;       // the score columns have been read: 128 arrivals            (softmax warps -> lanes A and B)
  uint64_t ldk[kKStages], ldv[2];   // TMA completion of the key / value ring stages
  uint64_t ldq;          // TMA completion of the query tiles
  uint32_t tmem_slot;
};

// grid: B*H*nqb, 192 threads, two CTAs per SM (256 TMEM columns each: S1 | S2 | O).
// tmQ / tmQ2 / tmV: the activations [B,T,H,dk]; tmKc: the centred keys in the workspace ([nm*B*H, T, 1, 64])
// Warp-specialised like sdpa2::fwd_kernel (sdpa_tc2.cuh), no CTA-wide barrier inside the loop:
//   warps 0-3       one query row per thread: scores -> registers (arrive s_free) -> mix -> lazy online softmax -> P (arrive p_ready)
//   lane A (warp 4) key TMA loads (three-stage ring, refilled as soon as the scores of the old tile have been read) and
//                   S1(t+1) = Q Kc1^T, issued the moment tile t's scores are in registers
//   lane B (warp 5) S2(t+1) = Q2 Kc2^T at the same moment, then O += P(t) V_t when p_ready(t) completes, and the value TMA loads
// One lane doing all of it needed ~2900 cycles per tile (~110 cycles per tcgen05.mma, TMA latency exposed by a two-stage ring)
// against ~2000 for the softmax warps.  Every mbarrier is waited on phase by phase by each of its waiters and its next phase
// cannot complete before that waiter has moved on (parity waits cannot tell phase k from k + 2).
#ifndef MOP_QLAZY
#define MOP_QLAZY 8.f
#endif
constexpr float kLazy = MOP_QLAZY;   // the running maximum is raised only when a tile exceeds it by 2^kLazy
template <bool HAS_MASK, bool DROP = false>
static __global__ void __launch_bounds__(192, 2) fwd_kernel(MopQuartetParams p, Ws w, unsigned char* ws, const __grid_constant__ CUtensorMap tmQ,
                                                     const __grid_constant__ CUtensorMap tmQ2, const __grid_constant__ CUtensorMap tmKc,
                                                     const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();   // the 128-byte-swizzled TMA tiles need a 1024-byte aligned base
  const int tid = threadIdx.x, warp = tid >> 5, dk = p.dk, T = p.T, nqb = w.nqb;
  // heavy (late) query blocks first: the causal work per block grows with its index
  const int qb = nqb - 1 - (int)(blockIdx.x / ((unsigned)p.B * p.H)), bh = blockIdx.x % (p.B * p.H), b = bh / p.H, h = bh % p.H;
  const int q0 = qb * 128;
  const Mix mx = load_mix(p);
  const size_t BH = (size_t)p.B * p.H;
  if (warp == 0) tmem_alloc<256>(&sm.tmem_slot);
  if (tid == 0) {
    mbar_init(&sm.bar_s, 2); mbar_init(&sm.bar_pv, 1); mbar_init(&sm.p_ready, 128); mbar_init(&sm.s_free, 128);
    for (int s = 0; s < kKStages; ++s) mbar_init(&sm.ldk[s], 1);
    mbar_init(&sm.ldv[0], 1); mbar_init(&sm.ldv[1], 1); mbar_init(&sm.ldq, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();   // barriers initialised, TMEM allocated
  tc_fence_after();
  const int k_end = min(T, q0 + 128);
  const int ntiles = (k_end + 63) >> 6;   // >= 1
  auto fetch_k = [&](int t, int s) {   // lane A
    mbar_expect_tx(&sm.ldk[s], (mx.quart ? 2 : 1) * kT64);
    tma_load_tile_sw(sm.K1[s], &tmKc, 64 * t, 0, bh, &sm.ldk[s]);
    if (mx.quart) tma_load_tile_sw(sm.K2[s], &tmKc, 64 * t, 0, (int)BH + bh, &sm.ldk[s]);
  };
  auto fetch_v = [&](int t) {          // lane B
    mbar_expect_tx(&sm.ldv[t & 1], kT64);
    tma_load_tile_sw(sm.V[t & 1], &tmV, 64 * t, h, b, &sm.ldv[t & 1]);
  };
  const uint32_t tb = sm.tmem_slot;
  if (tid == 128) {
    // query tiles by TMA; the Gram tile images (hi, lo per map; 8 KB each, contiguous in the workspace) by bulk copies into key-ring
    // stages 1 and 2, which the sigma prologue borrows - all on one transaction barrier, no thread copies anything
    mbar_expect_tx(&sm.ldq, (mx.quart ? 2 : 1) * (kT128 + 2 * kT64));
    tma_load_tile_sw(sm.Q, &tmQ, q0, h, b, &sm.ldq);
    bulk_g2s(sm.K1[1], ws + w.gram + (size_t)bh * 2 * kT64, kT64, &sm.ldq);
    bulk_g2s(sm.K2[1], ws + w.gram + (size_t)bh * 2 * kT64 + kT64, kT64, &sm.ldq);
    if (mx.quart) {
      tma_load_tile_sw(sm.Q2, &tmQ2, q0, h, b, &sm.ldq);
      bulk_g2s(sm.K1[2], ws + w.gram + (BH + bh) * 2 * kT64, kT64, &sm.ldq);
      bulk_g2s(sm.K2[2], ws + w.gram + (BH + bh) * 2 * kT64 + kT64, kT64, &sm.ldq);
    }
    fetch_k(0, 0);
    // ---- sigma prologue: G q by MMA (G = hi + lo) into the score columns
    mbar_wait(&sm.ldq, 0);
    mma_x_sym(tb, smem_u32(sm.Q), smem_u32(sm.K1[1]), smem_u32(sm.K2[1]));
    if (mx.quart) mma_x_sym(tb + 64, smem_u32(sm.Q2), smem_u32(sm.K1[2]), smem_u32(sm.K2[2]));
    mma_commit(&sm.bar_s);   // phase 0 of bar_s (two arrivals: this one and lane B's); tile t completes phase t + 1
  } else if (tid == 160) {
    fetch_v(0);
    mma_commit(&sm.bar_s);   // nothing outstanding: arrives at once
  }
  float s1v = 0.f, s2v = 0.f, a1 = 0.f, a2 = 0.f;
  if (warp < 4) {
    mbar_wait(&sm.ldq, 0);   // every softmax thread reads its query row below
    mbar_wait(&sm.bar_s, 0);
    tc_fence_after();
    const uint32_t tl = tb + ((uint32_t)(32 * warp) << 16);
    float quad1 = 0.f, quad2 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float z1[16], z2[16], x[8];
      tmem_ld_32x32b_x16(tl + 16 * c, z1);
      if (mx.quart) tmem_ld_32x32b_x16(tl + 64 + 16 * c, z2);
      tmem_ld_wait();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        unpack8(*reinterpret_cast<const uint4*>(sm.Q + sw128_off(tid, 8 * (2 * c + hh))), x);
#pragma unroll
        for (int e = 0; e < 8; ++e) quad1 = fmaf(z1[8 * hh + e], x[e], quad1);
        if (mx.quart) {
          unpack8(*reinterpret_cast<const uint4*>(sm.Q2 + sw128_off(tid, 8 * (2 * c + hh))), x);
#pragma unroll
          for (int e = 0; e < 8; ++e) quad2 = fmaf(z2[8 * hh + e], x[e], quad2);
        }
      }
    }
    s1v = p.scale * sqrtf(fmaxf(quad1, 0.f) / (float)(T - 1));
    s2v = mx.quart ? p.scale * sqrtf(fmaxf(quad2, 0.f) / (float)(T - 1)) : 0.f;
    a1 = p.scale / (s1v + mx.eps);
    a2 = mx.quart ? p.scale / (s2v + mx.eps) : 0.f;
  }
  tc_fence_before();
  __syncthreads();   // the Gram tiles and the prologue's TMEM columns are dead: key stages 1, 2 and the score columns are free
  tc_fence_after();
  if (warp >= 4) {
    const uint32_t id_s = idesc_bf16(128, 64, 0, 0);
    if (tid == 128) {
      // ---- lane A: key TMA loads + S1 ----------------------------------------------------------------------------
      const uint64_t dq1 = desc_k_sw(smem_u32(sm.Q), 0), dk1 = desc_k_sw(smem_u32(sm.K1[0]), 0);
      for (int t = 1; t < min(ntiles, kKStages); ++t) fetch_k(t, t);
      TS_DECL
      int st = 0, par = 0;   // ring stage / parity of the tile whose S1 is issued next
      for (int t = 0; t < ntiles; ++t) {   // issues S1(t); for t >= 1 right after the scores of tile t-1 have been read
        TS(t, 0);
        if (t >= 1) {
          mbar_wait(&sm.s_free, (uint32_t)(t - 1) & 1u);   // => S1 / S2(t-1) complete: their key stage is free
          tc_fence_after();
          if (t + 2 < ntiles) fetch_k(t + 2, st == 0 ? kKStages - 1 : st - 1);   // tile t+2 takes the stage of tile t-1
        }
        TS(t, 1);
        mbar_wait(&sm.ldk[st], (uint32_t)par);
        TS(t, 2);
        const uint64_t off = (uint64_t)(st * (kT64 >> 4));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_ss(tb, dq1 + (uint64_t)(2 * ks), dk1 + off + (uint64_t)(2 * ks), id_s, ks > 0 ? 1u : 0u);
        mma_commit(&sm.bar_s);
        TS(t, 3);
        if (++st == kKStages) { st = 0; par ^= 1; }
      }
      TS_DUMP("S-lane", ntiles, 4);
    } else if (tid == 160) {
      // ---- lane B: S2, O += P(t) V_t, value TMA loads ----------------------------------------------------------------
      const uint32_t id_pv = idesc_bf16(128, 64, 0, 1);
      const uint64_t dq2 = desc_k_sw(smem_u32(sm.Q2), 0), dk2 = desc_k_sw(smem_u32(sm.K2[0]), 0);
      const uint64_t dp = desc_kmajor(smem_u32(sm.P), 128, 0), dv0 = desc_mn_sw(smem_u32(sm.V[0]), 0);
      if (ntiles > 1) fetch_v(1);
      TS_DECL
      int st = 0, par = 0;
      auto issue_s2 = [&](int t) {   // S2(t) (nothing to do without the second map) + this lane's arrival on bar_s
        if (mx.quart) {
          mbar_wait(&sm.ldk[st], (uint32_t)par);
          const uint64_t off = (uint64_t)(st * (kT64 >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_ss(tb + 64, dq2 + (uint64_t)(2 * ks), dk2 + off + (uint64_t)(2 * ks), id_s, ks > 0 ? 1u : 0u);
        }
        mma_commit(&sm.bar_s);
        if (++st == kKStages) { st = 0; par ^= 1; }
      };
      issue_s2(0);
      for (int t = 0; t < ntiles; ++t) {
        TS(t, 0);
        if (t + 1 < ntiles) {
          mbar_wait(&sm.s_free, (uint32_t)t & 1u);   // the scores of tile t are in registers
          tc_fence_after();
          issue_s2(t + 1);
        }
        if (t >= 1 && t + 1 < ntiles) {   // refill the value stage of tile t-1.  Must precede this tile's commit on bar_pv: a
          mbar_wait(&sm.bar_pv, (uint32_t)(t - 1) & 1u);   // parity wait cannot tell phase t-1 from phase t+1
          fetch_v(t + 1);
        }
        TS(t, 1);
        mbar_wait(&sm.ldv[t & 1], (uint32_t)(t >> 1) & 1u);
        mbar_wait(&sm.p_ready, (uint32_t)t & 1u);
        tc_fence_after();
        TS(t, 2);
        const uint64_t dv = dv0 + (uint64_t)((t & 1) * (kT64 >> 4));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_ss(tb + 128, dp + (uint64_t)(ks * 256), dv + (uint64_t)(ks * 128), id_pv, (t > 0 || ks > 0) ? 1u : 0u);
        mma_commit(&sm.bar_pv);
        TS(t, 3);
      }
      TS_DUMP("PV-lane", ntiles, 4);
    }
  } else {
    // ---- softmax warps: one query row per thread ------------------------------------------------------------
    const int gi = q0 + tid;
    const bool row_ok = gi < T;
    const uint32_t tl = tb + ((uint32_t)(32 * warp) << 16);
    const float fa_ = mx.quart ? a1 * (1.f - mx.m) : a1, fb_ = mx.quart ? a1 * mx.m * mx.gam * a2 : 0.f;
    const float2 fA2 = make_float2(fa_, fa_), fB2 = make_float2(fb_, fb_), l2e2 = make_float2(kLog2e, kLog2e);
    float m_ref = -INFINITY, l_run = 0.f;   // m_ref in base-2 units
    const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
    const uint32_t rkey = dropout_row_key(drop, (uint32_t)bh, (uint32_t)gi);
    TS_DECL
    for (int it = 0; it < ntiles; ++it) {
      const int k0 = it * 64;
      TS(it, 0);
      mbar_wait(&sm.bar_s, (uint32_t)(it + 1) & 1u);
      tc_fence_after();
      TS(it, 1);
      float sc[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v1[16], v2[16];
        tmem_ld_32x32b_x16(tl + 16 * c, v1);
        if (mx.quart) tmem_ld_32x32b_x16(tl + 64 + 16 * c, v2);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 2) {   // mix = n1 ((1-m) + m gamma n2) = u (fA + fB v) on packed fp32 math
          const float2 tt = mx.quart ? fma2(fB2, make_float2(v2[e], v2[e + 1]), fA2) : fA2;
          const float2 r = mul2(make_float2(v1[e], v1[e + 1]), tt);
          sc[16 * c + e] = r.x; sc[16 * c + e + 1] = r.y;
        }
      }
      tc_fence_before();
      mbar_arrive(&sm.s_free);
      TS(it, 2);
      if ((k0 + 63 > q0) || (k0 + 64 > T)) {   // tiles on the diagonal / past the end only (uniform branch)
#pragma unroll
        for (int e = 0; e < 64; ++e)
          if (k0 + e > gi || k0 + e >= T) sc[e] = -INFINITY;
      }
      if constexpr (HAS_MASK) {
        if (row_ok) {
          const float* am = p.add_mask + (int64_t)b * p.am_sb + (int64_t)h * p.am_sh + (int64_t)gi * p.am_sq + (int64_t)k0 * p.am_sk;
#pragma unroll 8
          for (int e = 0; e < 64; ++e)
            if (k0 + e < T) sc[e] += am[(int64_t)e * p.am_sk];
        }
      }
      float t0 = sc[0], t1 = sc[1], t2 = sc[2], t3 = sc[3];
#pragma unroll
      for (int e = 4; e < 64; e += 4) { t0 = fmaxf(t0, sc[e]); t1 = fmaxf(t1, sc[e + 1]); t2 = fmaxf(t2, sc[e + 2]); t3 = fmaxf(t3, sc[e + 3]); }
      const float tm2 = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)) * kLog2e;
      const bool need = tm2 > m_ref + kLazy;                        // -inf + 8 = -inf: the first finite tile always raises it
      const float m_new = need ? tm2 : m_ref;
      const float corr = need ? ((m_ref == -INFINITY) ? 0.f : ex2(m_ref - m_new)) : 1.f;
      const float mb = (m_new == -INFINITY) ? 0.f : m_new;
      TS(it, 3);
      if (it > 0) { mbar_wait(&sm.bar_pv, (uint32_t)(it - 1) & 1u); tc_fence_after(); }   // P V(it-1) (issued most of a tile ago): P is free, O final
      TS(it, 4);
      if (it > 0 && __any_sync(0xffffffffu, need)) {   // rare
        float o[64];
        tmem_ld_32x32b_x32(tl + 128, o);
        tmem_ld_32x32b_x32(tl + 160, o + 32);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 64; ++e) o[e] *= corr;
        tmem_st_32x32b_x32(tl + 128, o);
        tmem_st_32x32b_x32(tl + 160, o + 32);
        tmem_st_wait();
      }
      l_run *= corr;
      m_ref = m_new;
      float2 ps0 = make_float2(0.f, 0.f), ps1 = ps0;
      const float2 nmb2 = make_float2(-mb, -mb);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float2 a = fma2(make_float2(sc[8 * c + e], sc[8 * c + e + 1]), l2e2, nmb2);
          pv[e] = ex2(a.x); pv[e + 1] = ex2(a.y);
        }
        ps0 = add2(ps0, add2(make_float2(pv[0], pv[1]), make_float2(pv[2], pv[3])));
        ps1 = add2(ps1, add2(make_float2(pv[4], pv[5]), make_float2(pv[6], pv[7])));
        if constexpr (DROP) {   // the row sum above is that of the un-dropped probabilities; 1/(1-p) is folded into the final 1/l
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (!dropout_keep(rkey, (uint32_t)(k0 + 8 * c + e), drop.thresh)) pv[e] = 0.f;
        }
        *reinterpret_cast<uint4*>(sm.P + c * (128 * 16) + tid * 16) = pack8(pv);
      }
      l_run += (ps0.x + ps0.y) + (ps1.x + ps1.y);
      fence_async_smem();   // P (generic proxy) -> tensor pipe (async proxy)
      tc_fence_before();
      mbar_arrive(&sm.p_ready);
      TS(it, 5);
    }
    if (tid == 0) TS_DUMP("softmax", ntiles, 6);
    mbar_wait(&sm.bar_pv, (uint32_t)(ntiles - 1) & 1u);
    tc_fence_after();
    const float il = (DROP ? drop.inv_keep : 1.f) / l_run;   // fully masked row: 0/0 = NaN like the reference softmax
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + at(p, b, row_ok ? gi : 0, h);
    float* y32 = p.y_f32 ? p.y_f32 + at(p, b, row_ok ? gi : 0, h) : nullptr;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float o[16];
      tmem_ld_32x32b_x16(tl + 128 + 16 * c, o);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) o[e] *= il;
      if (row_ok) {
        if (16 * c < dk) *reinterpret_cast<uint4*>(y + 16 * c) = pack8(o);
        if (16 * c + 8 < dk) *reinterpret_cast<uint4*>(y + 16 * c + 8) = pack8(o + 8);
        if (y32) {
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            if (16 * c + e < dk) *reinterpret_cast<float4*>(y32 + 16 * c + e) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
        }
      }
    }
    if (p.stats && row_ok) {
      float* st = p.stats + (((size_t)b * p.H + h) * T + gi) * 3;
      st[0] = s1v; st[1] = s2v; st[2] = kLn2 * (m_ref + lg2(l_run));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tb);
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
struct __align__(128) SmemQ {
  unsigned char Q[kT128], Q2[kT128], dO[kT128], W1[2][kT128], W2[2][kT128];   // W: one buffer per tile parity
  unsigned char K1[3][kT64], K2[3][kT64], V[3][kT64];   // three-stage key / value ring
  float gx[2][2][128];   // [warpgroup][map][row]: row-coefficient partial sums
  float red[16];
  uint64_t bar;      // MMA completion (Gram epilogue)
  uint64_t bar_in[2];   // S1, S2, dP of a tile complete (three issuing threads); one barrier per input buffer (tile parity)
  uint64_t bar_out;  // dQ1, dQ2 of a tile complete (two issuing threads)
  uint64_t ld[3];    // TMA completion of the ring stages
  uint64_t ldq;      // TMA completion of the query-side tiles
  uint32_t tmem_slot;
};

// Per-element backward of the mix given the raw dot products; returns the probability and writes dn1, dn2.
struct ElemOut { float pr, dn1, dn2, c1, c2; };
__device__ __forceinline__ ElemOut elem_bwd(const Mix& mx, float r1, float r2, float dp, float scale, float i1, float i2, float lse, float dlt,
                                            bool masked, float addm, float* sc0, float* sc1) {
  ElemOut o;
  o.c1 = r1 * scale; o.c2 = r2 * scale;
  const float n1 = o.c1 * i1, n2 = o.c2 * i2;
  const float s = mix_n(mx, n1, n2) + addm;
  o.pr = masked ? 0.f : ex2((s - lse) * kLog2e);
  const float D = o.pr * (dp - dlt);
  o.dn1 = D; o.dn2 = 0.f;
  if (mx.quart) {
    o.dn1 = D * ((1.f - mx.m) + mx.m * mx.gam * n2);
    o.dn2 = D * (mx.m * mx.gam * n1);
    *sc0 += D * (-n1 + mx.gam * n1 * n2);
    *sc1 += D * (mx.m * n1 * n2);
  }
  return o;
}

// grid: B*H*nqb, 256 threads (two warpgroups split the 64 columns of every tile), one CTA per SM
// TMEM: two input buffers {S1 | S2 | dP} at columns 0 and 192, dQ1 at 384, dQ2 at 448 (64 columns each).
// Software pipeline: S1 / S2 / dP of tile t+1 are issued (three lanes) as soon as tile t's have completed, into the other input
// buffer, so they run during tile t's element math; dQ += W K of tile t (two lanes) runs during tile t+1's element math (W has one
// buffer per tile parity); one lane refills the three-stage key / value ring once dQ(t-1) is done.
template <bool HAS_MASK, bool DROP = false>
static __global__ void __launch_bounds__(256, 1) bwd_dq_kernel(MopQuartetParams p, Ws w, unsigned char* ws, const unsigned char* pws, const __grid_constant__ CUtensorMap tmQ,
                                                        const __grid_constant__ CUtensorMap tmQ2, const __grid_constant__ CUtensorMap tmdO,
                                                        const __grid_constant__ CUtensorMap tmKc, const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemQ& sm = *reinterpret_cast<SmemQ*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();   // the 128-byte-swizzled TMA tiles need a 1024-byte aligned base
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp4 = (tid >> 5) & 3, lane = tid & 31, dk = p.dk, T = p.T, nqb = w.nqb;
  const int qb = nqb - 1 - (int)(blockIdx.x / ((unsigned)p.B * p.H)), bh = blockIdx.x % (p.B * p.H), b = bh / p.H, h = bh % p.H;
  const int q0 = qb * 128, gi = q0 + t;
  const bool row_ok = gi < T;
  const int dks = (dk + 15) >> 4;
  const int c0 = 32 * wg;   // this warpgroup's columns inside a 64-key tile
  const Mix mx = load_mix(p);
  const size_t BH = (size_t)p.B * p.H, stride = (size_t)p.H * dk;
  if (tid < 32) tmem_alloc<512>(&sm.tmem_slot);
  if (tid == 0) {
    mbar_init(&sm.bar, 1); mbar_init(&sm.bar_in[0], 3); mbar_init(&sm.bar_in[1], 3); mbar_init(&sm.bar_out, 2);
    mbar_init(&sm.ld[0], 1); mbar_init(&sm.ld[1], 1); mbar_init(&sm.ld[2], 1); mbar_init(&sm.ldq, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int k_end = min(T, q0 + 128);
  const int ntiles = (k_end + 63) >> 6;   // >= 1
  auto fetch = [&](int tile) {   // one thread: key tile `tile` -> ring stage tile % 3
    const int s = tile % 3, k0 = 64 * tile;
    mbar_expect_tx(&sm.ld[s], (mx.quart ? 3 : 2) * kT64);
    tma_load_tile_sw(sm.K1[s], &tmKc, k0, 0, bh, &sm.ld[s]);
    if (mx.quart) tma_load_tile_sw(sm.K2[s], &tmKc, k0, 0, (int)BH + bh, &sm.ld[s]);
    tma_load_tile_sw(sm.V[s], &tmV, k0, h, b, &sm.ld[s]);
  };
  if (tid == 0) {
    mbar_expect_tx(&sm.ldq, (mx.quart ? 3 : 2) * kT128);
    tma_load_tile_sw(sm.Q, &tmQ, q0, h, b, &sm.ldq);
    if (mx.quart) tma_load_tile_sw(sm.Q2, &tmQ2, q0, h, b, &sm.ldq);
    tma_load_tile_sw(sm.dO, &tmdO, q0, h, b, &sm.ldq);
    fetch(0);
    if (ntiles > 1) fetch(1);
  }
  // per-row statistics (both warpgroups need them)
  const float* st = p.stats + (((size_t)b * p.H + h) * T + (row_ok ? gi : T - 1)) * 3;
  const float s1v = st[0], s2v = st[1], lse = st[2];
  const float i1 = 1.f / (s1v + mx.eps), i2 = mx.quart ? 1.f / (s2v + mx.eps) : 0.f;
  const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
  const uint32_t rkey = dropout_row_key(drop, (uint32_t)bh, (uint32_t)gi);
  // delta = dO . y per query row: eight lanes share a row (32 B of fp32 y and 16 B of dO per lane and pass: four rows = 1 KB
  // per warp instruction, all of a warp's 16 rows in flight) - one thread per 256-byte row costs 32 sectors per load.
  {
    const int wrp = tid >> 5, sub = lane >> 3, l8 = lane & 7;
    float part[4];
    float4 ya[4], yb[4];
    uint4 da[4], yh[4];
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      const int r = q0 + 16 * wrp + 4 * ps + sub;
      const bool ok = r < T && 8 * l8 < dk;
      const size_t o = at(p, b, ok ? r : 0, h) + 8 * l8;
      da[ps] = ok ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + o)) : make_uint4(0, 0, 0, 0);
      if (p.y_f32) {
        ya[ps] = ok ? __ldg(reinterpret_cast<const float4*>(p.y_f32 + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
        yb[ps] = ok ? __ldg(reinterpret_cast<const float4*>(p.y_f32 + o + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        yh[ps] = ok ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.y) + o)) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      float a[8], c[8];
      if (p.y_f32) {
        a[0] = ya[ps].x; a[1] = ya[ps].y; a[2] = ya[ps].z; a[3] = ya[ps].w; a[4] = yb[ps].x; a[5] = yb[ps].y; a[6] = yb[ps].z; a[7] = yb[ps].w;
      } else {
        unpack8(yh[ps], a);
      }
      unpack8(da[ps], c);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(a[e], c[e], d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      part[ps] = d;
    }
    if (l8 == 0) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) sm.gx[0][0][16 * wrp + 4 * ps + sub] = part[ps];
    }
  }
  __syncthreads();
  const float dlt = row_ok ? sm.gx[0][0][t] : 0.f;
  if (row_ok && wg == 0) (reinterpret_cast<float*>(ws + w.delta) + (size_t)bh * T)[gi] = dlt;   // for bwd_dkdv
  __syncthreads();   // gx is written again by the epilogue only, but keep the read and any later write apart
  const uint32_t tb = sm.tmem_slot, tl = tb + ((uint32_t)(32 * warp4) << 16);
  uint32_t phase = 0, ph_in = 0, ph_out = 0;
  float g1 = 0.f, g2 = 0.f, sc0 = 0.f, sc1 = 0.f;
  // Without an additive mask the tiles run on packed fp32 math with the per-row constants folded in (zero-filled key rows >= T
  // contribute nothing; on diagonal tiles the probabilities above the diagonal are zeroed).  With u, v the raw dot products:  t = f / (sigma1+eps) = A1 + B1 v,  s = scale u t,
  // D = P (dP - delta),  W1 = D t,  W2 = D u B1;  the row sums the epilogue needs all follow from Sa = sum D u, Sb = sum D u v.
  const float fA = mx.quart ? 1.f - mx.m : 1.f, fB = mx.quart ? mx.m * mx.gam * p.scale * i2 : 0.f;   // f = fA + fB v
  const float2 A1 = make_float2(fA * i1, fA * i1), B1 = make_float2(fB * i1, fB * i1), cs2 = make_float2(p.scale * kLog2e, p.scale * kLog2e);
  const float2 L2 = make_float2(-lse * kLog2e, -lse * kLog2e), nd2 = make_float2(-dlt, -dlt);
  float2 Sa2 = make_float2(0.f, 0.f), Sb2 = make_float2(0.f, 0.f);
  // S1 / S2 / dP of tile `tile` into input buffer tile & 1: lane 0 of warps 0, 1, 2, one product each; all three commit (an empty
  // commit arrives at once).  A lone issuing lane needs ~110 cycles per tcgen05.mma and one CTA per SM hides none of it.
  auto issue_in = [&](int tile) {
    const int s = tile % 3;
    if (tile == 0) mbar_wait(&sm.ldq, 0);
    mbar_wait(&sm.ld[s], (uint32_t)(tile / 3) & 1u);
    const uint32_t id = idesc_bf16(128, 64, 0, 0), xc = tb + 192u * (uint32_t)(tile & 1);
    if (tid == 0) {
      for (int ks = 0; ks < dks; ++ks) mma_ss(xc, desc_k_sw(smem_u32(sm.Q), 16 * ks), desc_k_sw(smem_u32(sm.K1[s]), 16 * ks), id, ks > 0 ? 1u : 0u);
    } else if (tid == 32) {
      if (mx.quart)
        for (int ks = 0; ks < dks; ++ks) mma_ss(xc + 64, desc_k_sw(smem_u32(sm.Q2), 16 * ks), desc_k_sw(smem_u32(sm.K2[s]), 16 * ks), id, ks > 0 ? 1u : 0u);
    } else {
      for (int ks = 0; ks < dks; ++ks) mma_ss(xc + 128, desc_k_sw(smem_u32(sm.dO), 16 * ks), desc_k_sw(smem_u32(sm.V[s]), 16 * ks), id, ks > 0 ? 1u : 0u);
    }
    mma_commit(&sm.bar_in[tile & 1]);
  };
  const bool in_lane = tid == 0 || tid == 32 || tid == 64;
  if (in_lane) issue_in(0);
  for (int it = 0; it < ntiles; ++it) {
    const int k0 = it * 64, par = it & 1;
    const uint32_t tx = tl + 192u * (uint32_t)par;   // this tile's input buffer
    // Tile it+1's products are issued below, BEFORE the CTA barrier of this tile, so a barrier shared by all tiles could complete
    // two phases while a late warp (slow prologue loads) had not yet waited for the first - the parity wait then blocks for
    // ever (seen as a rare hang of whole sub-partitions).  One barrier per tile parity: the phase a warp waits for can be
    // followed by at most one more before the CTA barrier of tile it+1, which needs that warp.
    mbar_wait(&sm.bar_in[par], (uint32_t)(it >> 1) & 1u); tc_fence_after();
    if (in_lane && it + 1 < ntiles) issue_in(it + 1);   // the other input buffer was consumed before the barrier of tile it-1
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col = c0 + 16 * c;
      float v1[16], v2[16], dp[16], w1[16], w2[16];
      tmem_ld_32x32b_x16(tx + col, v1);
      if (mx.quart) tmem_ld_32x32b_x16(tx + 64 + col, v2);
      tmem_ld_32x32b_x16(tx + 128 + col, dp);
      tmem_ld_wait();
      if constexpr (DROP) {   // D = P (.) (M (.) dP - delta), M = dropout factor of the forward
#pragma unroll
        for (int e = 0; e < 16; ++e) dp[e] *= dropout_factor(drop, rkey, (uint32_t)(k0 + col + e));
      }
      if (!HAS_MASK) {
        const int lim = gi - k0 - col;   // causal: element e of this chunk is masked when e > lim (diagonal tiles only)
        const bool diag = k0 + 63 > q0;
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float2 u = make_float2(v1[e], v1[e + 1]), v = mx.quart ? make_float2(v2[e], v2[e + 1]) : make_float2(0.f, 0.f);
          const float2 tt = fma2(B1, v, A1);
          const float2 a = fma2(mul2(u, tt), cs2, L2);
          float2 pr = make_float2(ex2(a.x), ex2(a.y));
          if (diag) { pr.x = e > lim ? 0.f : pr.x; pr.y = e + 1 > lim ? 0.f : pr.y; }
          const float2 D = mul2(pr, add2(make_float2(dp[e], dp[e + 1]), nd2));
          const float2 Du = mul2(D, u), W1 = mul2(D, tt), W2 = mul2(Du, B1);
          Sa2 = add2(Sa2, Du);
          Sb2 = fma2(Du, v, Sb2);
          w1[e] = W1.x; w1[e + 1] = W1.y; w2[e] = W2.x; w2[e + 1] = W2.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int gj = k0 + col + e;
          const bool masked = gj > gi || gj >= T || !row_ok;
          float addm = 0.f;
          if constexpr (HAS_MASK)
            if (!masked) addm = p.add_mask[(int64_t)b * p.am_sb + (int64_t)h * p.am_sh + (int64_t)gi * p.am_sq + (int64_t)gj * p.am_sk];
          const ElemOut o = elem_bwd(mx, v1[e], mx.quart ? v2[e] : 0.f, dp[e], p.scale, i1, i2, lse, dlt, masked, addm, &sc0, &sc1);
          g1 = fmaf(o.dn1, o.c1, g1);
          g2 = fmaf(o.dn2, o.c2, g2);
          w1[e] = o.dn1 * i1;
          w2[e] = o.dn2 * i2;
        }
      }
      const int ch = col >> 3;
      *reinterpret_cast<uint4*>(sm.W1[par] + ch * (128 * 16) + t * 16) = pack8(w1);   // W[par] is free: dQ(it-2) completed before the barrier of tile it-1
      *reinterpret_cast<uint4*>(sm.W1[par] + (ch + 1) * (128 * 16) + t * 16) = pack8(w1 + 8);
      if (mx.quart) {
        *reinterpret_cast<uint4*>(sm.W2[par] + ch * (128 * 16) + t * 16) = pack8(w2);
        *reinterpret_cast<uint4*>(sm.W2[par] + (ch + 1) * (128 * 16) + t * 16) = pack8(w2 + 8);
      }
    }
    if (tid == 96) {   // ring refill (lane 0 of warp 3): dQ(it-1), issued a tile ago, has read the keys of stage (it-1) % 3
      if (it >= 1) { mbar_wait(&sm.bar_out, ph_out); ph_out ^= 1; }
      if (it + 2 < ntiles) fetch(it + 2);
    }
    publish();   // W visible to the tensor pipe; every thread now knows dQ(it-1) has completed
    if (tid == 128 || tid == 160) {
      const uint32_t id = idesc_bf16(128, 64, 0, 1);
      const int s = it % 3;
      if (tid == 128) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_ss(tb + 384, desc_kmajor(smem_u32(sm.W1[par]), 128, 16 * ks), desc_mn_sw(smem_u32(sm.K1[s]), 16 * ks), id, (it > 0 || ks > 0) ? 1u : 0u);
      } else if (mx.quart) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_ss(tb + 448, desc_kmajor(smem_u32(sm.W2[par]), 128, 16 * ks), desc_mn_sw(smem_u32(sm.K2[s]), 16 * ks), id, (it > 0 || ks > 0) ? 1u : 0u);
      }
      mma_commit(&sm.bar_out);
    }
  }
  // dQ(ntiles-2) completed before the last barrier, so the parity of the last phase is unambiguous for every thread
  mbar_wait(&sm.bar_out, (uint32_t)(ntiles - 1) & 1u); tc_fence_after();
  mbar_wait(&sm.ldq, 0);   // (already complete: orders the TMA-written query tiles before the generic reads below)
  if (row_ok) {   // fold the packed-math tiles into the row sums
    const float Sa = Sa2.x + Sa2.y, Sb = Sb2.x + Sb2.y, al1 = p.scale * i1, al2 = p.scale * i2;
    g1 += p.scale * (fA * Sa + fB * Sb);
    if (mx.quart) {
      g2 += mx.m * mx.gam * al1 * p.scale * Sb;
      sc0 += al1 * (mx.gam * al2 * Sb - Sa);
      sc1 += mx.m * al1 * al2 * Sb;
    }
  }
  // row coefficients: combine the two column halves
  sm.gx[wg][0][t] = g1;
  sm.gx[wg][1][t] = g2;
  // G q by MMA: the Gram tiles (hi, lo per map) go to the dead key / value / W buffers; Z1 -> columns [0,64), Z2 -> [64,128)
  {
    unsigned char* const dst[4] = {sm.K1[0], sm.K2[0], sm.V[0], sm.W1[0]};
    // (pws: the workspace that holds the key preparation - this call's, or the forward call's)
    const unsigned char* const src[4] = {pws + w.gram + (size_t)bh * 2 * kT64, pws + w.gram + (size_t)bh * 2 * kT64 + kT64,
                                         pws + w.gram + (BH + bh) * 2 * kT64, pws + w.gram + (BH + bh) * 2 * kT64 + kT64};
    copy_tiles64(dst, src, mx.quart ? 4 : 2);
  }
  publish();
  if (tid == 0) {
    mma_x_sym(tb, smem_u32(sm.Q), smem_u32(sm.K1[0]), smem_u32(sm.K2[0]));
    if (mx.quart) mma_x_sym(tb + 64, smem_u32(sm.Q2), smem_u32(sm.V[0]), smem_u32(sm.W1[0]));
    mma_commit(&sm.bar);
  }
  mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
  const float Tm1 = (float)(T - 1);
  const float g1t = sm.gx[0][0][t] + sm.gx[1][0][t], g2t = sm.gx[0][1][t] + sm.gx[1][1][t];
  const float gg1 = row_ok ? g1t / ((s1v + mx.eps) * (s1v + mx.eps) * Tm1 * s1v) : 0.f;
  const float gg2 = (row_ok && mx.quart) ? g2t / ((s2v + mx.eps) * (s2v + mx.eps) * Tm1 * s2v) : 0.f;
  for (int map = 0; map < w.nm; ++map) {
    const float gv = map ? gg2 : gg1;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(map ? p.dq2 : p.dq) + at(p, b, row_ok ? gi : 0, h);
#pragma unroll
    for (int c = 0; c < 2; ++c) {   // this warpgroup's 32 output columns
      const int col = c0 + 16 * c;
      float acc[16], wv[16];
      tmem_ld_32x32b_x16(tl + (map ? 448 : 384) + col, acc);
      tmem_ld_32x32b_x16(tl + (map ? 64 : 0) + col, wv);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = p.scale * acc[e] - p.scale * p.scale * gv * wv[e];
      if (row_ok) {
        if (col < dk) *reinterpret_cast<uint4*>(out + col) = pack8(acc);
        if (col + 8 < dk) *reinterpret_cast<uint4*>(out + col + 8) = pack8(acc + 8);
      }
    }
    if (row_ok && wg == 0) (reinterpret_cast<float*>(ws + w.gvec) + ((size_t)map * BH + bh) * T)[gi] = gv;
  }
  // scalar partials of this query block
  sc0 = warp_sum(sc0); sc1 = warp_sum(sc1);
  if (lane == 0) { sm.red[tid >> 5] = sc0; sm.red[8 + (tid >> 5)] = sc1; }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += sm.red[i]; c += sm.red[8 + i]; }
    float* sp = reinterpret_cast<float*>(ws + w.spart) + ((size_t)bh * nqb + qb) * 2;
    sp[0] = mx.m * (1.f - mx.m) * a;
    sp[1] = c;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tb);
}

// grid: B*H*nm, 256 threads.  M = sum_i g_i q_i q_i^T = (g q)^T q on the tensor core, g q split into bf16 hi + lo
// groups > 1: the queries of one (b, h, map) are split over `groups` CTAs of `cpg` 64-row chunks each; partial sums go to
// w.macc with atomics and acc_to_tiles_kernel writes the tile images (used when B*H*nm alone does not fill the GPU)
static __global__ void __launch_bounds__(256) gmat_kernel(MopQuartetParams p, Ws w, unsigned char* ws, int groups, int cpg) {
  __shared__ __align__(128) unsigned char tiles[1][3][kT64];   // [g q hi | g q lo | q] (single buffer: several CTAs share an SM)
  __shared__ GramPipe gp;
  float* G = reinterpret_cast<float*>(&tiles[0][0][0]);        // 16 KB: read out once every MMA has completed
  const int nm = w.nm, grp = blockIdx.x % groups, bm = blockIdx.x / groups, bh = bm / nm, map = bm % nm, b = bh / p.H, h = bh % p.H;
  const int dk = p.dk, T = p.T, tid = threadIdx.x;
  const size_t BH = (size_t)p.B * p.H, stride = (size_t)p.H * dk;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(map ? p.q2 : p.q) + at(p, b, 0, h);
  const float* gv = reinterpret_cast<const float*>(ws + w.gvec) + ((size_t)map * BH + bh) * T;
  if (tid < 32) tmem_alloc<64>(&gp.tmem_slot);
  if (tid == 0) { mbar_init(&gp.bar[0], 1); mbar_init(&gp.bar[1], 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = gp.tmem_slot;
  const int r0 = tid >> 3, ch = tid & 7;
  uint32_t phases = 0;
  const int c_lo = grp * cpg, nchunks = min(cpg, ((T + 63) >> 6) - c_lo);   // this CTA's 64-row chunks (>= 1)
  auto load_rows = [&](int c, uint4* u, float* gg) {   // this thread's two 16-byte pieces of chunk c and their row coefficients
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int t = (c_lo + c) * 64 + r0 + 32 * it;
      const bool in = c < nchunks && t < T;
      u[it] = (in && ch * 8 < dk) ? *reinterpret_cast<const uint4*>(q + (size_t)t * stride + ch * 8) : make_uint4(0, 0, 0, 0);
      gg[it] = in ? gv[t] : 0.f;
    }
  };
  uint4 cur[2], nxt[2];
  float gcur[2], gnxt[2];
  load_rows(0, cur, gcur);
  for (int c = 0; c < nchunks; ++c) {
    const int buf = 0;
    load_rows(c + 1, nxt, gnxt);   // in flight while this chunk is converted, staged and multiplied
    if (c >= 1) { mbar_wait(&gp.bar[buf], (phases >> buf) & 1u); phases ^= 1u << buf; }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = r0 + 32 * it;
      const uint4 raw = cur[it];
      const float g = gcur[it];
      float f[8], hi[8], lo[8];
      unpack8(raw, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float x = g * f[e];
        hi[e] = __bfloat162float(__float2bfloat16_rn(x));
        lo[e] = x - hi[e];
      }
      *reinterpret_cast<uint4*>(tiles[buf][0] + ch * 1024 + r * 16) = pack8(hi);
      *reinterpret_cast<uint4*>(tiles[buf][1] + ch * 1024 + r * 16) = pack8(lo);
      *reinterpret_cast<uint4*>(tiles[buf][2] + ch * 1024 + r * 16) = raw;
    }
    publish();
    if (tid == 0) {
      const uint32_t id = idesc_bf16(64, 64, 1, 1);
      const uint32_t th = smem_u32(tiles[buf][0]), tlo = smem_u32(tiles[buf][1]), tq = smem_u32(tiles[buf][2]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tb, desc_mnmajor(th, 64, 16 * ks), desc_mnmajor(tq, 64, 16 * ks), id, (c > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_ss(tb, desc_mnmajor(tlo, 64, 16 * ks), desc_mnmajor(tq, 64, 16 * ks), id, 1u);
      mma_commit(&gp.bar[buf]);
    }
    cur[0] = nxt[0]; cur[1] = nxt[1]; gcur[0] = gnxt[0]; gcur[1] = gnxt[1];
  }
  mbar_wait(&gp.bar[0], phases & 1u);
  tc_fence_after();
  gram_readout(tb, G);
  tc_fence_before();
  __syncthreads();
  if (groups == 1) {
    write_hilo_tiles(G, ws + w.mmat + ((size_t)map * BH + bh) * 2 * kT64);
  } else {
    float* acc = reinterpret_cast<float*>(ws + w.macc) + ((size_t)map * BH + bh) * 64 * 64;
    for (int i = tid; i < 64 * 64; i += 256) atomicAdd(acc + i, G[i]);
  }
  if (tid < 32) tmem_dealloc<64>(tb);
}

struct __align__(128) SmemK {
  unsigned char K1[kT128], K2[kT128], V[kT128], PT[2][kT128], W1T[2][kT128], W2T[2][kT128];   // P^T / W^T: one buffer per tile parity
  unsigned char Q[3][kT64], Q2[3][kT64], dO[3][kT64];   // three-stage ring of query-side tiles
  uint32_t rk[3][64];    // per query of the tile: dropout row key
  float vec[3][8][64];   // per query of the tile: sigma1 -> 1/(sigma1+eps), sigma2 -> .., lse, delta | packed-math constants A1, B1, L, -delta
  uint64_t bar;      // MMA completion (M kc epilogue)
  uint64_t bar_in;   // S1^T, S2^T, dP^T of a tile complete (three issuing threads)
  uint64_t bar_out;  // dV, dKc1, dKc2 of a tile complete (three issuing threads)
  uint64_t ld[3];    // TMA completion of the query-side ring stages
  uint64_t ldk;      // TMA completion of the key / value tiles
  uint32_t tmem_slot;
};


// grid: B*H*nqb (128 keys per CTA), 256 threads (thread per key row; two warpgroups split the 64 query columns of every
// tile), one CTA per SM.  TMEM: S1^T | S2^T | dP^T | dV | dKc1 | dKc2  (64 columns each).
// The three accumulators leave room for one input buffer only, so the input products of tile t+1 are issued at the barrier of tile
// t (ahead of the output products in the tensor pipe's queue); the output products of tile t run during tile t+1 (P^T / W^T have
// one buffer per tile parity) and a lane refills the three-stage query ring once the outputs of tile t-1 are done.
template <bool HAS_MASK, bool DROP = false>
static __global__ void __launch_bounds__(256, 1) bwd_dkdv_kernel(MopQuartetParams p, Ws w, unsigned char* ws, const __grid_constant__ CUtensorMap tmQ,
                                                          const __grid_constant__ CUtensorMap tmQ2, const __grid_constant__ CUtensorMap tmdO,
                                                          const __grid_constant__ CUtensorMap tmKc, const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemK& sm = *reinterpret_cast<SmemK*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();   // the 128-byte-swizzled TMA tiles need a 1024-byte aligned base
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp4 = (tid >> 5) & 3, dk = p.dk, T = p.T, nkb = w.nqb;
  const int kb = blockIdx.x % nkb, bh = blockIdx.x / nkb, b = bh / p.H, h = bh % p.H;   // early key blocks (most work) first
  const Dropout drop = make_dropout(p.dropout_p, p.dropout_seed, p.dropout_offset);
  const int k0 = kb * 128, gj = k0 + t;
  const bool key_ok = gj < T;
  const int dks = (dk + 15) >> 4;
  const int c0 = 32 * wg;
  const Mix mx = load_mix(p);
  const size_t BH = (size_t)p.B * p.H, stride = (size_t)p.H * dk;
  if (tid < 32) tmem_alloc<512>(&sm.tmem_slot);
  if (tid == 0) {
    mbar_init(&sm.bar, 1); mbar_init(&sm.bar_in, 3); mbar_init(&sm.bar_out, 3);
    mbar_init(&sm.ld[0], 1); mbar_init(&sm.ld[1], 1); mbar_init(&sm.ld[2], 1); mbar_init(&sm.ldk, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const float* stats = p.stats + ((size_t)b * p.H + h) * T * 3;
  const float* delta = reinterpret_cast<const float*>(ws + w.delta) + (size_t)bh * T;
  // query tile `tile` -> ring stage tile % 3: tiles by TMA (thread `tma_tid`), the per-query statistics by 4-byte cp.async
  // (threads 0..63); every thread commits exactly one cp.async group per call, `on` or not, so that the group count stays uniform
  auto fetch = [&](int tile, int tma_tid, bool on) {
    const int buf = tile % 3, q0 = k0 + 64 * tile;
    if (on && tid == tma_tid) {
      mbar_expect_tx(&sm.ld[buf], (mx.quart ? 3 : 2) * kT64);
      tma_load_tile_sw(sm.Q[buf], &tmQ, q0, h, b, &sm.ld[buf]);
      if (mx.quart) tma_load_tile_sw(sm.Q2[buf], &tmQ2, q0, h, b, &sm.ld[buf]);
      tma_load_tile_sw(sm.dO[buf], &tmdO, q0, h, b, &sm.ld[buf]);
    }
    if (on && tid < 64) {
      const int i = min(q0 + tid, T - 1);
      cp_async4(&sm.vec[buf][0][tid], stats + (size_t)i * 3);
      cp_async4(&sm.vec[buf][1][tid], stats + (size_t)i * 3 + 1);
      cp_async4(&sm.vec[buf][2][tid], stats + (size_t)i * 3 + 2);
      cp_async4(&sm.vec[buf][3][tid], delta + i);
      if constexpr (DROP) sm.rk[buf][tid] = dropout_row_key(drop, (uint32_t)bh, (uint32_t)i);
    }
    cp_async_commit();
  };
  if (tid == 0) {
    mbar_expect_tx(&sm.ldk, (mx.quart ? 3 : 2) * kT128);
    tma_load_tile_sw(sm.K1, &tmKc, k0, 0, bh, &sm.ldk);
    if (mx.quart) tma_load_tile_sw(sm.K2, &tmKc, k0, 0, (int)BH + bh, &sm.ldk);
    tma_load_tile_sw(sm.V, &tmV, k0, h, b, &sm.ldk);
  }
  const int ntiles = (T - k0 + 63) >> 6;   // queries i >= j only (k0 is a multiple of 64); >= 1
  fetch(0, 0, true);
  fetch(1, 0, ntiles > 1);
  const uint32_t tb = sm.tmem_slot, tl = tb + ((uint32_t)(32 * warp4) << 16);
  uint32_t phase = 0, ph_in = 0, ph_out = 0;
  float dum0 = 0.f, dum1 = 0.f;
  const bool in_lane = tid == 0 || tid == 32 || tid == 64;
  auto issue_in = [&](int tile) {   // transposed tiles: rows = keys, columns = queries; one issuing thread per product, all commit
    const int s = tile % 3;
    if (tile == 0) mbar_wait(&sm.ldk, 0);
    mbar_wait(&sm.ld[s], (uint32_t)(tile / 3) & 1u);
    const uint32_t id = idesc_bf16(128, 64, 0, 0);
    if (tid == 0) {
      for (int ks = 0; ks < dks; ++ks) mma_ss(tb, desc_k_sw(smem_u32(sm.K1), 16 * ks), desc_k_sw(smem_u32(sm.Q[s]), 16 * ks), id, ks > 0 ? 1u : 0u);
    } else if (tid == 32) {
      if (mx.quart)
        for (int ks = 0; ks < dks; ++ks) mma_ss(tb + 64, desc_k_sw(smem_u32(sm.K2), 16 * ks), desc_k_sw(smem_u32(sm.Q2[s]), 16 * ks), id, ks > 0 ? 1u : 0u);
    } else {
      for (int ks = 0; ks < dks; ++ks) mma_ss(tb + 128, desc_k_sw(smem_u32(sm.V), 16 * ks), desc_k_sw(smem_u32(sm.dO[s]), 16 * ks), id, ks > 0 ? 1u : 0u);
    }
    mma_commit(&sm.bar_in);
  };
  if (in_lane) issue_in(0);
  for (int it = 0; it < ntiles; ++it) {
    const int q0 = k0 + it * 64, buf = it % 3, par = it & 1;
    cp_async_wait<1>();   // groups 0 .. it+1 are committed: the statistics of tile `it` have landed
    if (tid < 64) {   // sigma -> 1 / (sigma + eps), in place (each thread converts the values it fetched itself)
      sm.vec[buf][0][tid] = 1.f / (sm.vec[buf][0][tid] + mx.eps);
      sm.vec[buf][1][tid] = mx.quart ? 1.f / (sm.vec[buf][1][tid] + mx.eps) : 0.f;
      const float i1 = sm.vec[buf][0][tid], i2 = sm.vec[buf][1][tid];
      sm.vec[buf][4][tid] = (mx.quart ? 1.f - mx.m : 1.f) * i1;                       // t = f / (sigma1+eps) = A1 + B1 v  (see bwd_dq_kernel)
      sm.vec[buf][5][tid] = mx.quart ? mx.m * mx.gam * p.scale * i2 * i1 : 0.f;
      sm.vec[buf][6][tid] = -sm.vec[buf][2][tid] * kLog2e;
      sm.vec[buf][7][tid] = -sm.vec[buf][3][tid];
    }
    __syncthreads();   // the per-query vectors of this tile are visible to every thread
    mbar_wait(&sm.bar_in, ph_in); ph_in ^= 1; tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int colb = c0 + 16 * c;
      float v1[16], v2[16], dp[16], pt[16], w1[16], w2[16];
      tmem_ld_32x32b_x16(tl + colb, v1);
      if (mx.quart) tmem_ld_32x32b_x16(tl + 64 + colb, v2);
      tmem_ld_32x32b_x16(tl + 128 + colb, dp);
      tmem_ld_wait();
      uint32_t km = 0xFFFFu;   // keep bits of this thread's 16 (query) columns
      if constexpr (DROP) {
        km = 0u;
#pragma unroll
        for (int e = 0; e < 16; ++e) km |= (dropout_keep(sm.rk[buf][colb + e], (uint32_t)gj, drop.thresh) ? 1u : 0u) << e;
#pragma unroll
        for (int e = 0; e < 16; ++e) dp[e] = ((km >> e) & 1u) ? dp[e] * drop.inv_keep : 0.f;
      }
      if (!HAS_MASK) {   // packed fp32 math (zero-filled query rows >= T contribute nothing; causal zeroing on diagonal tiles)
        const bool diag = q0 < k0 + 127;
        const int lim = gj - q0 - colb;   // query column x of this chunk is masked when x < lim
        const float2 cs2 = make_float2(p.scale * kLog2e, p.scale * kLog2e);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 A4 = *reinterpret_cast<const float4*>(&sm.vec[buf][4][colb + e]), B4 = *reinterpret_cast<const float4*>(&sm.vec[buf][5][colb + e]);
          const float4 L4 = *reinterpret_cast<const float4*>(&sm.vec[buf][6][colb + e]), N4 = *reinterpret_cast<const float4*>(&sm.vec[buf][7][colb + e]);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int x = e + 2 * hh;
            const float2 A1 = hh ? make_float2(A4.z, A4.w) : make_float2(A4.x, A4.y), B1 = hh ? make_float2(B4.z, B4.w) : make_float2(B4.x, B4.y);
            const float2 L2 = hh ? make_float2(L4.z, L4.w) : make_float2(L4.x, L4.y), nd2 = hh ? make_float2(N4.z, N4.w) : make_float2(N4.x, N4.y);
            const float2 u = make_float2(v1[x], v1[x + 1]), v = mx.quart ? make_float2(v2[x], v2[x + 1]) : make_float2(0.f, 0.f);
            const float2 tt = fma2(B1, v, A1);
            const float2 a = fma2(mul2(u, tt), cs2, L2);
            float2 pr = make_float2(ex2(a.x), ex2(a.y));
            if (diag) { pr.x = x < lim ? 0.f : pr.x; pr.y = x + 1 < lim ? 0.f : pr.y; }
            const float2 D = mul2(pr, add2(make_float2(dp[x], dp[x + 1]), nd2));
            const float2 W1 = mul2(D, tt), W2 = mul2(mul2(D, u), B1);
            pt[x] = pr.x; pt[x + 1] = pr.y; w1[x] = W1.x; w1[x + 1] = W1.y; w2[x] = W2.x; w2[x + 1] = W2.y;
          }
        }
      } else
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int col = colb + e, gi = q0 + col;
        const bool masked = gj > gi || gi >= T || !key_ok;
        float addm = 0.f;
        if constexpr (HAS_MASK)
          if (!masked) addm = p.add_mask[(int64_t)b * p.am_sb + (int64_t)h * p.am_sh + (int64_t)gi * p.am_sq + (int64_t)gj * p.am_sk];
        const float i1 = sm.vec[buf][0][col], i2 = sm.vec[buf][1][col];
        const ElemOut o = elem_bwd(mx, v1[e], mx.quart ? v2[e] : 0.f, dp[e], p.scale, i1, i2, sm.vec[buf][2][col], sm.vec[buf][3][col], masked, addm, &dum0, &dum1);
        pt[e] = o.pr;
        w1[e] = o.dn1 * i1;
        w2[e] = o.dn2 * i2;
      }
      const int ch = colb >> 3;
      if constexpr (DROP) {   // dV = (M (.) P)^T dO
#pragma unroll
        for (int e = 0; e < 16; ++e) pt[e] = ((km >> e) & 1u) ? pt[e] * drop.inv_keep : 0.f;
      }
      *reinterpret_cast<uint4*>(sm.PT[par] + ch * (128 * 16) + t * 16) = pack8(pt);   // free: the outputs of tile it-2 completed before
      *reinterpret_cast<uint4*>(sm.PT[par] + (ch + 1) * (128 * 16) + t * 16) = pack8(pt + 8);   // the barrier of tile it-1
      *reinterpret_cast<uint4*>(sm.W1T[par] + ch * (128 * 16) + t * 16) = pack8(w1);
      *reinterpret_cast<uint4*>(sm.W1T[par] + (ch + 1) * (128 * 16) + t * 16) = pack8(w1 + 8);
      if (mx.quart) {
        *reinterpret_cast<uint4*>(sm.W2T[par] + ch * (128 * 16) + t * 16) = pack8(w2);
        *reinterpret_cast<uint4*>(sm.W2T[par] + (ch + 1) * (128 * 16) + t * 16) = pack8(w2 + 8);
      }
    }
    if (tid == 96 && it >= 1) { mbar_wait(&sm.bar_out, ph_out); ph_out ^= 1; }   // outputs of tile it-1 (issued a tile ago): stage (it-1) % 3 is free
    fetch(it + 2, 96, it + 2 < ntiles);
    publish();   // P^T / W^T visible to the tensor pipe; the input columns have been read; every thread knows the outputs of tile it-1 are done
    if (in_lane && it + 1 < ntiles) issue_in(it + 1);   // queued ahead of this tile's output products
    if (tid == 128 || tid == 160 || tid == 192) {   // K index = queries of this tile; one issuing thread per product
      const uint32_t id = idesc_bf16(128, 64, 0, 1);
      const uint32_t acc0 = it > 0 ? 1u : 0u;
      if (tid == 128) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_ss(tb + 192, desc_kmajor(smem_u32(sm.PT[par]), 128, 16 * ks), desc_mn_sw(smem_u32(sm.dO[buf]), 16 * ks), id, (acc0 || ks > 0) ? 1u : 0u);
      } else if (tid == 160) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_ss(tb + 256, desc_kmajor(smem_u32(sm.W1T[par]), 128, 16 * ks), desc_mn_sw(smem_u32(sm.Q[buf]), 16 * ks), id, (acc0 || ks > 0) ? 1u : 0u);
      } else if (mx.quart) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_ss(tb + 320, desc_kmajor(smem_u32(sm.W2T[par]), 128, 16 * ks), desc_mn_sw(smem_u32(sm.Q2[buf]), 16 * ks), id, (acc0 || ks > 0) ? 1u : 0u);
      }
      mma_commit(&sm.bar_out);
    }
  }
  cp_async_wait<0>();
  // the outputs of tile ntiles-2 completed before the last barrier, so the parity of the last phase is unambiguous for every thread
  mbar_wait(&sm.bar_out, (uint32_t)(ntiles - 1) & 1u); tc_fence_after();
  // dV: this warpgroup's 32 columns
  __nv_bfloat16* dv = reinterpret_cast<__nv_bfloat16*>(p.dv) + at(p, b, key_ok ? gj : 0, h);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int col = c0 + 16 * c;
    float acc[16];
    tmem_ld_32x32b_x16(tl + 192 + col, acc);
    tmem_ld_wait();
    if (key_ok) {
      if (col < dk) *reinterpret_cast<uint4*>(dv + col) = pack8(acc);
      if (col + 8 < dk) *reinterpret_cast<uint4*>(dv + col + 8) = pack8(acc + 8);
    }
  }
  // M kc by MMA: the M tiles (hi, lo per map) go to the dead query / dO buffers; Z1 -> columns [0,64), Z2 -> [64,128)
  __syncthreads();
  {
    unsigned char* const dst[4] = {sm.Q[0], sm.Q[1], sm.dO[0], sm.dO[1]};
    const unsigned char* const src[4] = {ws + w.mmat + (size_t)bh * 2 * kT64, ws + w.mmat + (size_t)bh * 2 * kT64 + kT64,
                                         ws + w.mmat + (BH + bh) * 2 * kT64, ws + w.mmat + (BH + bh) * 2 * kT64 + kT64};
    copy_tiles64(dst, src, mx.quart ? 4 : 2);
  }
  publish();
  if (tid == 0) {
    mma_x_sym(tb, smem_u32(sm.K1), smem_u32(sm.Q[0]), smem_u32(sm.Q[1]));
    if (mx.quart) mma_x_sym(tb + 64, smem_u32(sm.K2), smem_u32(sm.dO[0]), smem_u32(sm.dO[1]));
    mma_commit(&sm.bar);
  }
  mbar_wait(&sm.bar, phase); phase ^= 1; tc_fence_after();
  for (int map = 0; map < w.nm; ++map) {
    float* out = reinterpret_cast<float*>(ws + w.dkc) + (((size_t)map * BH + bh) * T + (key_ok ? gj : 0)) * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col = c0 + 16 * c;
      float acc[16], wv[16];
      tmem_ld_32x32b_x16(tl + (map ? 320 : 256) + col, acc);
      tmem_ld_32x32b_x16(tl + (map ? 64 : 0) + col, wv);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = key_ok ? p.scale * acc[e] - p.scale * p.scale * wv[e] : 0.f;
      if (key_ok) {
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4)
          *reinterpret_cast<float4*>(out + col + 4 * e4) = make_float4(acc[4 * e4], acc[4 * e4 + 1], acc[4 * e4 + 2], acc[4 * e4 + 3]);
      }
      // column sums of this warp's 32 key rows -> global (finish_kernel subtracts the mean over all T keys)
      {
        int cix;
        const float cs = warp_colsum16(acc, tid & 31, &cix);   // 16-step butterfly: even lanes hold one column sum each
        float* sum = reinterpret_cast<float*>(ws + w.dksum) + ((size_t)map * BH + bh) * 64 + col;
        if ((tid & 1) == 0) atomicAdd(sum + cix, cs);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tb);
}

// grid: (B*H*nm) * ceil(T/64), 256 threads: dk = dkc - mean_j dkc for 64 keys, the column sums come from bwd_dkdv (w.dksum);
// the first chunk of map 0 also folds the scalar partials of the query blocks
static __global__ void __launch_bounds__(256) finish_kernel(MopQuartetParams p, Ws w, unsigned char* ws) {
  __shared__ float mean[64];
  const int nm = w.nm, T = p.T, nch = (T + 63) >> 6, dk = p.dk, tid = threadIdx.x;
  const int chunk = blockIdx.x % nch, bm = blockIdx.x / nch, bh = bm / nm, map = bm % nm, b = bh / p.H, h = bh % p.H;
  const size_t BH = (size_t)p.B * p.H;
  const float* dkc = reinterpret_cast<const float*>(ws + w.dkc) + ((size_t)map * BH + bh) * T * 64;
  if (tid < 64) mean[tid] = (reinterpret_cast<const float*>(ws + w.dksum) + ((size_t)map * BH + bh) * 64)[tid] / (float)T;
  __syncthreads();
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(map ? p.dk2 : p.dk_) + at(p, b, 0, h);
  const size_t stride = (size_t)p.H * dk;
  for (int idx = tid; idx < 64 * 8; idx += 256) {
    const int t = chunk * 64 + (idx >> 3), ch = idx & 7;
    if (t >= T || ch * 8 >= dk) continue;
    const float4 a = *reinterpret_cast<const float4*>(dkc + (size_t)t * 64 + ch * 8), c = *reinterpret_cast<const float4*>(dkc + (size_t)t * 64 + ch * 8 + 4);
    const float f[8] = {a.x - mean[ch * 8], a.y - mean[ch * 8 + 1], a.z - mean[ch * 8 + 2], a.w - mean[ch * 8 + 3],
                        c.x - mean[ch * 8 + 4], c.y - mean[ch * 8 + 5], c.z - mean[ch * 8 + 6], c.w - mean[ch * 8 + 7]};
    *reinterpret_cast<uint4*>(out + (size_t)t * stride + ch * 8) = pack8(f);
  }
  if (chunk == 0 && map == 0 && tid < 2 && p.dscalar_part) {
    float s = 0.f;
    const float* sp = reinterpret_cast<const float*>(ws + w.spart) + (size_t)bh * w.nqb * 2;
    for (int c = 0; c < w.nqb; ++c) s += sp[c * 2 + tid];
    p.dscalar_part[(size_t)bh * 2 + tid] = s;
  }
}

inline bool supported(const MopQuartetParams* p) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return p->dtype == MOP_BF16 && p->dk % 8 == 0 && p->dk <= 64 && p->T >= 2 && al16(p->q) && al16(p->k) && al16(p->v) &&
         (!p->use_quartet || (al16(p->q2) && al16(p->k2)));
}

}  // namespace qtc
}  // namespace mop
