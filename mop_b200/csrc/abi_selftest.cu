// libmop_b200.so - C ABI entry points (see include/mop_b200.h): tcgen05 primitive self-tests (tests/test_gpu_tc_primitives.py)
// Host side only validates, sizes scratch and enqueues kernels on the caller's stream.
#include "abi_host.h"
#include "tc_selftest.cuh"

using namespace mop;

extern "C" int mop_selftest_tma(const void* x, void* out, int B, int N, int H, int dk, int64_t sb, int64_t sn, int64_t sh, int R, int row0,
                                int head, int batch, void* stream) {
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device");
  MOP_REQUIRE(R > 0 && R <= 256 && dk % 8 == 0 && dk <= 64, MOP_EINVAL, "bad selftest shape");
  CUtensorMap tm;
  int rc = make_tile_map_sw(&tm, x, B, N, H, dk, sb, sn, sh, R);
  if (rc != MOP_OK) return rc;
  MOP_CHECK_CUDA(cudaFuncSetAttribute(tc::selftest_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, R * 128 + 1024));
  tc::selftest_tma_kernel<<<1, 128, R * 128 + 1024, (cudaStream_t)stream>>>(tm, reinterpret_cast<unsigned char*>(out), R, row0, head, batch);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_selftest_umma128(const float* A, const float* B, float* D, int Ma, int Nn, int K, int b_mn, int Ra, int Rb,
                                    int Kb, int b_k0, void* stream) {
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device");
  MOP_REQUIRE(Ma > 0 && Ma <= 256 && Nn % 16 == 0 && Nn >= 16 && Nn <= 256 && K % 16 == 0 && K > 0 && Ra % 8 == 0 && Rb % 8 == 0, MOP_EINVAL,
              "bad selftest shape");
  const size_t smem = (size_t)Ra * 16 * (K / 8) + (size_t)Rb * 16 * ((b_mn ? Nn : K) / 8) + 4096;   // slack: M-row over-read
  MOP_REQUIRE(smem <= 227 * 1024, MOP_EINVAL, "selftest tiles do not fit in shared memory");
  MOP_CHECK_CUDA(cudaFuncSetAttribute(tc::selftest128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::selftest128_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(A, B, D, Ma, Nn, K, b_mn, Ra, Rb, Kb, b_k0);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_selftest_umma(const float* A, const float* B, float* D, float* D2, int a_mn, int b_mn, int lane_off,
                                 int col_off, void* stream) {
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device");
  MOP_REQUIRE((lane_off == 0 || lane_off == 16) && col_off >= 0 && col_off <= 64, MOP_EINVAL, "bad lane/col offset");
  tc::selftest_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(A, B, D, D2, a_mn, b_mn, lane_off, col_off);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
