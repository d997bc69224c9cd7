// Bring-up / regression kernel for the tcgen05 primitives in tc_common.cuh.
// D = (a_mn ? A^T : A) * (b_mn ? B : B^T) for 64x64 fp32 inputs rounded to bf16, accumulators read
// back through the 16x256b fragment path; D2 = D + 1 after a tcgen05.st / tcgen05.ld round trip.
#pragma once
#include "tc_common.cuh"

namespace mop {
namespace tc {

__global__ void __launch_bounds__(128, 1) selftest_kernel(const float* A, const float* B, float* D, float* D2, int a_mn,
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

}  // namespace tc
}  // namespace mop
