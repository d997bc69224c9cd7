// libmop_b200.so - C ABI entry points (see include/mop_b200.h): fused Whisper-MoP 2D gate
#include "abi_host.h"
#include "gates.cuh"

using namespace mop;

extern "C" {

int mop_mop2d_partial_rows(void) {
  int sms = sm_count();
  return (sms > 0 ? sms : 148) * 2;
}

int mop_mop2d_fwd(const float* mel, const float* He, float* R, float* gate, int B, int T, int F, int ks, void* stream) {
  MOP_REQUIRE(mel && He && R && gate && B > 0 && T > 0 && F > 0, MOP_EINVAL, "bad mop2d arguments");
  MOP_REQUIRE(ks >= 1 && ks <= gates::kMaxK && (ks & 1) && ks / 2 < F, MOP_EUNSUPPORTED, "kernel_size=%d (odd, <= %d)", ks, gates::kMaxK);
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)B * T;
  gates::mop2d_rowsums_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(mel, R, B, T, F, ks);
  gates::mop2d_gate_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(R, He, gate, B, T, F, ks);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

int mop_mop2d_bwd(const float* R, const float* dgate, float* dHe_part, int nparts, int B, int T, int F, int ks, void* stream) {
  MOP_REQUIRE(R && dgate && dHe_part && B > 0 && T > 0 && F > 0, MOP_EINVAL, "bad mop2d arguments");
  MOP_REQUIRE(ks >= 1 && ks <= gates::kMaxK && (ks & 1), MOP_EUNSUPPORTED, "kernel_size=%d (odd, <= %d)", ks, gates::kMaxK);
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  MOP_REQUIRE(nparts >= 1, MOP_EWORKSPACE, "dHe_part needs at least one partial row");
  gates::mop2d_bwd_kernel<<<nparts, 256, 0, (cudaStream_t)stream>>>(R, dgate, dHe_part, B, T, F, ks);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

}  // extern "C"
