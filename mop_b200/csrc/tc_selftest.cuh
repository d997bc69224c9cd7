// Bring-up / regression kernel for the tcgen05 primitives in tc_common.cuh.
// D = (a_mn ? A^T : A) * (b_mn ? B : B^T) for 64x64 fp32 inputs rounded to bf16, accumulators read
// back through the 16x256b fragment path; D2 = D + 1 after a tcgen05.st / tcgen05.ld round trip.
#pragma once
#include "tc_common.cuh"

namespace mop {
namespace tc {

static __global__ void __launch_bounds__(128, 1) selftest_kernel(const float* A, const float* B, float* D, float* D2, int a_mn,
                                                          int b_mn, int lane_off, int col_off) {
  __shared__ __align__(128) unsigned char tA[64 * 64 * 2];
  __shared__ __align__(128) unsigned char tB[64 * 64 * 2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int idx = tid; idx < 64 * 64; idx += 128) {
    int r = idx >> 6, c = idx & 63;
    *reinterpret_cast<__nv_bfloat16*>(tA + tile_off(64, r, c)) = __float2bfloat16_rn(A[idx]);
    *reinterpret_cast<__nv_bfloat16*>(tB + tile_off(64, r, c)) = __float2bfloat16_rn(B[idx]);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t dacc = tbase + ((uint32_t)lane_off << 16) + (uint32_t)col_off;
  if (tid == 0) {
    gemm64<4>(dacc, smem_u32(tA), a_mn != 0, smem_u32(tB), b_mn != 0, false);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  Frag f;
  float v[32];
  const uint32_t taddr = tbase + ((uint32_t)(32 * warp + lane_off) << 16) + (uint32_t)col_off;
  tmem_ld_16x256b_x8(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    D[f.row_lo * 64 + f.col(n)] = v[4 * n];
    D[f.row_lo * 64 + f.col(n) + 1] = v[4 * n + 1];
    D[f.row_hi * 64 + f.col(n)] = v[4 * n + 2];
    D[f.row_hi * 64 + f.col(n) + 1] = v[4 * n + 3];
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] += 1.0f;
  tmem_st_16x256b_x8(taddr, v);
  tmem_st_wait();
  float u[32];
  tmem_ld_16x256b_x8(taddr, u);
  tmem_ld_wait();
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    D2[f.row_lo * 64 + f.col(n)] = u[4 * n];
    D2[f.row_lo * 64 + f.col(n) + 1] = u[4 * n + 1];
    D2[f.row_hi * 64 + f.col(n)] = u[4 * n + 2];
    D2[f.row_hi * 64 + f.col(n) + 1] = u[4 * n + 3];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tbase);
}


// Bring-up of the M=128 building blocks used by the large-N kernels (edgewise_tc_large.cuh):
//   D[256 x Nn] = A[Ma x K] * op(B) in two M=128 row blocks, operands in chunk-major tiles with Ra / Rb rows,
//   accumulators at TMEM columns 0 and 256, read back with the 32x32b "thread per row" shape after a
//   tcgen05.st / tcgen05.ld round trip of block 1 (+1 then -1).
//   b_mn = 0: B is [Nn x K] row-major (tile rows = N index, K-major operand)
//   b_mn = 1: B is [Kb x Nn] row-major (tile rows = K index, MN-major operand); the GEMM uses rows
//             [b_k0, b_k0 + K) of it.
static __global__ void __launch_bounds__(256, 1) selftest128_kernel(const float* A, const float* B, float* D, int Ma, int Nn, int K,
                                                             int b_mn, int Ra, int Rb, int Kb, int b_k0) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, t = tid & 127;
  unsigned char* tA = smem;
  const int bytesA = Ra * 16 * (K / 8);
  unsigned char* tB = smem + bytesA;
  const int colsB = b_mn ? Nn : K, rowsB = b_mn ? Kb : Nn;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int idx = tid; idx < Ra * K; idx += 256) {
    int r = idx / K, c = idx % K;
    *reinterpret_cast<__nv_bfloat16*>(tA + tile_off(Ra, r, c)) = __float2bfloat16_rn(r < Ma ? A[(size_t)r * K + c] : 0.f);
  }
  for (int idx = tid; idx < Rb * colsB; idx += 256) {
    int r = idx / colsB, c = idx % colsB;
    *reinterpret_cast<__nv_bfloat16*>(tB + tile_off(Rb, r, c)) = __float2bfloat16_rn(r < rowsB ? B[(size_t)r * colsB + c] : 0.f);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t id = idesc_bf16(128, Nn, 0, b_mn ? 1u : 0u);
    for (int blk = 0; blk < 2; ++blk)
      for (int k = 0; k < K / 16; ++k) {
        uint64_t ad = desc_kmajor(smem_u32(tA) + blk * 128 * 16, Ra, 16 * k);
        uint64_t bd = b_mn ? desc_mnmajor(smem_u32(tB), Rb, b_k0 + 16 * k) : desc_kmajor(smem_u32(tB), Rb, 16 * k);
        mma_ss(tbase + 256 * blk, ad, bd, id, k > 0 ? 1u : 0u);
      }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + 256 * wg;
  for (int c = 0; c < Nn; c += 16) {
    float v[16];
    tmem_ld_32x32b_x16(tl + c, v);
    tmem_ld_wait();
    if (wg == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += 1.0f;
      tmem_st_32x32b_x16(tl + c, v);
      tmem_st_wait();
      float u[8];
      tmem_ld_32x32b_x8(tl + c, u);
      tmem_ld_32x32b_x8(tl + c + 8, v + 8);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = u[i];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] -= 1.0f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)(128 * wg + t) * Nn + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}


// Bring-up of the TMA tile load: rows [row0, row0 + R) of head `head`, batch `batch` of a [B][N][H][dk] bf16 tensor -> 128-byte
// swizzled tile in shared memory -> copied out verbatim (R * 128 bytes) for comparison on the host.
static __global__ void __launch_bounds__(128, 1) selftest_tma_kernel(const __grid_constant__ CUtensorMap tm, unsigned char* out, int R, int row0,
                                                              int head, int batch) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < R * 128 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_async_smem();
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, (uint32_t)R * 128u);
    tma_load_tile_sw(smem, &tm, row0, head, batch, &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = tid; i < R * 128 / 16; i += 128) reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(smem)[i];
}

}  // namespace tc
}  // namespace mop
