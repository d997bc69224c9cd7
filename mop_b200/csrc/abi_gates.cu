// libmop_b200.so - C ABI entry points (see include/mop_b200.h): fused Whisper-MoP 2D gate
#include "abi_host.h"
#include "gates.cuh"
#include "token_gate.cuh"

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

static int check_token_gate(const MopTokenGateParams* p, bool bwd) {
  MOP_REQUIRE(p != nullptr, MOP_EINVAL, "null params");
  MOP_REQUIRE(p->struct_bytes == (int)sizeof(MopTokenGateParams), MOP_EINVAL, "MopTokenGateParams size mismatch: caller %d, library %zu", p->struct_bytes, sizeof(MopTokenGateParams));
  MOP_REQUIRE(p->dtype == MOP_F32 || p->dtype == MOP_BF16, MOP_EINVAL, "dtype");
  MOP_REQUIRE(p->B > 0 && p->T > 0 && p->Gh > 0 && p->Gw > 0 && p->Gh * p->Gw == p->T, MOP_EINVAL, "tokens must form a Gh x Gw grid");
  MOP_REQUIRE(p->T <= tokgate::kMaxT, MOP_EUNSUPPORTED, "T=%d tokens per image (<= %d)", p->T, tokgate::kMaxT);
  MOP_REQUIRE(p->D >= 8 && p->D % 8 == 0 && p->D <= 8 * tokgate::kThreads, MOP_EUNSUPPORTED, "D=%d (multiple of 8, <= %d)", p->D, 8 * tokgate::kThreads);
  MOP_REQUIRE(p->V >= 1 && p->V <= tokgate::kMaxV && p->K >= 1 && p->K <= tokgate::kMaxK && p->hid >= 1 && p->hid <= tokgate::kMaxHid, MOP_EUNSUPPORTED,
              "V=%d K=%d hid=%d (<= %d, %d, %d)", p->V, p->K, p->hid, tokgate::kMaxV, tokgate::kMaxK, tokgate::kMaxHid);
  MOP_REQUIRE(p->x && p->views_w && p->k3_w && p->k1_w && p->f1_w && p->f2_w && p->f2_b && p->a_pos && p->a_neg && p->views && p->gate, MOP_EINVAL, "null tensor");
  if (bwd) {
    MOP_REQUIRE(p->dout && p->dx && p->dwv_part && p->dnet_part, MOP_EINVAL, "null gradient tensor");
    MOP_REQUIRE(p->nparts >= 1 && p->nparts <= p->B, MOP_EWORKSPACE, "nparts=%d (1..B)", p->nparts);
  } else {
    MOP_REQUIRE(p->out != nullptr, MOP_EINVAL, "null out");
  }
  return MOP_OK;
}

int mop_token_gate_partial_rows(int B) {
  int sms = sm_count();
  const int cap = (sms > 0 ? sms : 148) * 2;
  return B < cap ? (B > 0 ? B : 1) : cap;
}
int mop_token_gate_wv_groups(int D) { return (D >= 8 && D / 8 <= tokgate::kThreads) ? tokgate::wv_groups(D) : 0; }
int mop_token_gate_net_params(const MopTokenGateParams* p) { return p ? tokgate::dims(*p).nnet : 0; }

int mop_token_gate_fwd(MopTokenGateParams* p, void* stream) {
  int rc = check_token_gate(p, false);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  const size_t smem = tokgate::smem_floats_fwd(*p) * sizeof(float);
  MOP_REQUIRE(smem <= 200 * 1024, MOP_EUNSUPPORTED, "token gate needs %zu bytes of shared memory", smem);
  const int grid = mop_token_gate_partial_rows(p->B);
  if (p->dtype == MOP_BF16) {
    if ((rc = allow_smem(tokgate::fwd_kernel<__nv_bfloat16>, smem))) return rc;
    tokgate::fwd_kernel<__nv_bfloat16><<<grid, tokgate::kThreads, smem, (cudaStream_t)stream>>>(*p);
  } else {
    if ((rc = allow_smem(tokgate::fwd_kernel<float>, smem))) return rc;
    tokgate::fwd_kernel<float><<<grid, tokgate::kThreads, smem, (cudaStream_t)stream>>>(*p);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

int mop_token_gate_bwd(MopTokenGateParams* p, void* stream) {
  int rc = check_token_gate(p, true);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  const size_t smem = tokgate::smem_floats_bwd(*p) * sizeof(float);
  MOP_REQUIRE(smem <= 200 * 1024, MOP_EUNSUPPORTED, "token gate backward needs %zu bytes of shared memory", smem);
  if (p->dtype == MOP_BF16) {
    if ((rc = allow_smem(tokgate::bwd_kernel<__nv_bfloat16>, smem))) return rc;
    tokgate::bwd_kernel<__nv_bfloat16><<<p->nparts, tokgate::kThreads, smem, (cudaStream_t)stream>>>(*p);
  } else {
    if ((rc = allow_smem(tokgate::bwd_kernel<float>, smem))) return rc;
    tokgate::bwd_kernel<float><<<p->nparts, tokgate::kThreads, smem, (cudaStream_t)stream>>>(*p);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

static int check_token_gate1d(const MopTokenGate1dParams* p, bool bwd) {
  MOP_REQUIRE(p != nullptr, MOP_EINVAL, "null params");
  MOP_REQUIRE(p->struct_bytes == (int)sizeof(MopTokenGate1dParams), MOP_EINVAL, "MopTokenGate1dParams size mismatch: caller %d, library %zu", p->struct_bytes, sizeof(MopTokenGate1dParams));
  MOP_REQUIRE(p->dtype == MOP_F32 || p->dtype == MOP_BF16, MOP_EINVAL, "dtype");
  MOP_REQUIRE(p->B > 0 && p->T > 0, MOP_EINVAL, "empty input");
  MOP_REQUIRE(p->D >= 8 && p->D % 8 == 0 && p->D <= 8 * tokgate::kThreads, MOP_EUNSUPPORTED, "D=%d (multiple of 8, <= %d)", p->D, 8 * tokgate::kThreads);
  MOP_REQUIRE(p->V >= 1 && p->V <= tokgate::kMaxV, MOP_EUNSUPPORTED, "V=%d (<= %d)", p->V, tokgate::kMaxV);
  MOP_REQUIRE(p->x && p->views_w && p->w_eff && p->views && p->gate, MOP_EINVAL, "null tensor");
  if (bwd) {
    MOP_REQUIRE(p->dout && p->dx && p->dwv_part && p->dweff_part, MOP_EINVAL, "null gradient tensor");
    MOP_REQUIRE(p->nparts >= 1, MOP_EWORKSPACE, "nparts=%d", p->nparts);
  } else {
    MOP_REQUIRE(p->out != nullptr, MOP_EINVAL, "null out");
  }
  return MOP_OK;
}

int mop_token_gate1d_partial_rows(int B, int T) {
  int sms = sm_count();
  const long long chunks = (long long)(B > 0 ? B : 1) * ((T + tokgate1d::kChunk - 1) / tokgate1d::kChunk), cap = (sms > 0 ? sms : 148) * 2;
  return (int)(chunks < cap ? (chunks > 0 ? chunks : 1) : cap);
}

int mop_token_gate1d_fwd(MopTokenGate1dParams* p, void* stream) {
  int rc = check_token_gate1d(p, false);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  const size_t smem = tokgate1d::smem_floats(*p) * sizeof(float);
  const int grid = mop_token_gate1d_partial_rows(p->B, p->T);
  if (p->dtype == MOP_BF16) {
    if ((rc = allow_smem(tokgate1d::fwd_kernel<__nv_bfloat16>, smem))) return rc;
    tokgate1d::fwd_kernel<__nv_bfloat16><<<grid, tokgate::kThreads, smem, (cudaStream_t)stream>>>(*p);
  } else {
    if ((rc = allow_smem(tokgate1d::fwd_kernel<float>, smem))) return rc;
    tokgate1d::fwd_kernel<float><<<grid, tokgate::kThreads, smem, (cudaStream_t)stream>>>(*p);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

int mop_token_gate1d_bwd(MopTokenGate1dParams* p, void* stream) {
  int rc = check_token_gate1d(p, true);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  const size_t smem = tokgate1d::smem_floats(*p) * sizeof(float);
  if (p->dtype == MOP_BF16) {
    if ((rc = allow_smem(tokgate1d::bwd_kernel<__nv_bfloat16>, smem))) return rc;
    tokgate1d::bwd_kernel<__nv_bfloat16><<<p->nparts, tokgate::kThreads, smem, (cudaStream_t)stream>>>(*p);
  } else {
    if ((rc = allow_smem(tokgate1d::bwd_kernel<float>, smem))) return rc;
    tokgate1d::bwd_kernel<float><<<p->nparts, tokgate::kThreads, smem, (cudaStream_t)stream>>>(*p);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

}  // extern "C"
