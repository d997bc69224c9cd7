// libmop_b200.so - C ABI entry points (see include/mop_b200.h): plain / causal / biased / cross attention
// Host side only validates, sizes scratch and enqueues kernels on the caller's stream.
#include "abi_host.h"
#include "sdpa_simt.cuh"
#include "sdpa_tc2.cuh"

namespace mop {
static int check_sdpa(const MopSdpaParams* p, bool bwd) {
  MOP_REQUIRE(p != nullptr, MOP_EINVAL, "params is NULL");
  MOP_REQUIRE(p->struct_bytes == (int32_t)sizeof(MopSdpaParams), MOP_EABI,
              "MopSdpaParams size mismatch: caller %d, library %d", p->struct_bytes, (int)sizeof(MopSdpaParams));
  MOP_REQUIRE(p->dtype == MOP_F32 || p->dtype == MOP_BF16, MOP_EINVAL, "bad dtype %d", p->dtype);
  MOP_REQUIRE(p->B > 0 && p->H > 0 && p->Nq > 0 && p->Nk > 0 && p->dk > 0, MOP_EINVAL, "bad shape");
  MOP_REQUIRE(p->dk <= sdpa::kMaxDk, MOP_EUNSUPPORTED, "head dim %d > %d", p->dk, sdpa::kMaxDk);
  MOP_REQUIRE(p->q && p->k && p->v && p->y, MOP_EINVAL, "q/k/v/y must be set");
  if (bwd) MOP_REQUIRE(p->dy && p->dq && p->dk_ && p->dv && p->lse, MOP_EINVAL, "backward buffers (dy,dq,dk,dv,lse) missing");
  return MOP_OK;
}
}  // namespace mop

using namespace mop;

extern "C" {

size_t mop_sdpa_workspace_bytes(const MopSdpaParams* p, int backward) {
  if (check_sdpa(p, false) != MOP_OK) return 0;
  if (!backward) return 0;
  const size_t simt = sdpa::bwd_workspace_floats(p) * sizeof(float), tcn = (size_t)p->B * p->H * p->Nq * sizeof(float);
  return simt > tcn ? simt : tcn;
}

int mop_sdpa_fwd(MopSdpaParams* p, void* stream) {
  int rc = check_sdpa(p, false);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  const bool tc_ok = sdpa2::supported(p);
  MOP_REQUIRE(p->impl == MOP_IMPL_AUTO || p->impl == MOP_IMPL_SIMT || (p->impl == MOP_IMPL_TCGEN05 && tc_ok), MOP_EUNSUPPORTED,
              "impl %d not available (tcgen05 path: bf16, dk%%8==0, dk<=64, 16-byte aligned rows)", p->impl);
  const int grid = p->B * p->H * ((p->Nq + sdpa::TQ - 1) / sdpa::TQ);
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_ok && p->impl != MOP_IMPL_SIMT) {
    const size_t smem_tc = sizeof(sdpa2::SmemF) + 128;
    const bool extra = p->bias != nullptr || p->zero_mask != nullptr;
    const bool drop = p->dropout_p > 0.f;   // kernels are compiled with and without the optional tensors / the dropout code
    auto kf = extra ? (drop ? sdpa2::fwd_kernel<true, true> : sdpa2::fwd_kernel<true, false>) : (drop ? sdpa2::fwd_kernel<false, true> : sdpa2::fwd_kernel<false, false>);
    if ((rc = allow_smem(kf, smem_tc))) return rc;
    CUtensorMap tmQ, tmK, tmV;
    if ((rc = make_tile_map_sw(&tmQ, p->q, p->B, p->Nq, p->H, p->dk, p->q_sb, p->q_sn, p->q_sh, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmK, p->k, p->B, p->Nk, p->H, p->dk, p->k_sb, p->k_sn, p->k_sh, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmV, p->v, p->B, p->Nk, p->H, p->dk, p->v_sb, p->v_sn, p->v_sh, 64))) return rc;
    kf<<<p->B * p->H * ((p->Nq + 127) / 128), 192, smem_tc, st>>>(*p, tmQ, tmK, tmV);
    MOP_CHECK_CUDA(cudaGetLastError());
    p->impl_used = MOP_IMPL_TCGEN05;
    return MOP_OK;
  }
  const size_t smem = sdpa::smem_bytes(p->dk);
  if (p->dtype == MOP_F32) {
    if ((rc = allow_smem(sdpa::fwd_kernel<float>, smem))) return rc;
    sdpa::fwd_kernel<float><<<grid, simt::kThreads, smem, st>>>(*p);
  } else {
    if ((rc = allow_smem(sdpa::fwd_kernel<__nv_bfloat16>, smem))) return rc;
    sdpa::fwd_kernel<__nv_bfloat16><<<grid, simt::kThreads, smem, st>>>(*p);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  p->impl_used = MOP_IMPL_SIMT;
  return MOP_OK;
}

int mop_sdpa_bwd(MopSdpaParams* p, void* stream) {
  int rc = check_sdpa(p, true);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  const bool tc_ok = sdpa2::supported(p);
  MOP_REQUIRE(p->impl == MOP_IMPL_AUTO || p->impl == MOP_IMPL_SIMT || (p->impl == MOP_IMPL_TCGEN05 && tc_ok), MOP_EUNSUPPORTED,
              "impl %d not available (tcgen05 path: bf16, dk%%8==0, dk<=64, 16-byte aligned rows)", p->impl);
  if (tc_ok && p->impl != MOP_IMPL_SIMT) {
    const size_t need_tc = (size_t)p->B * p->H * p->Nq * sizeof(float);
    MOP_REQUIRE(p->workspace && p->workspace_bytes >= need_tc, MOP_EWORKSPACE, "workspace too small: have %zu, need %zu", p->workspace_bytes, need_tc);
    float* delta = reinterpret_cast<float*>(p->workspace);
    cudaStream_t st2 = (cudaStream_t)stream;
    const size_t smem_q = sizeof(sdpa2::SmemQ) + 128, smem_k = sizeof(sdpa2::SmemK) + 128;
    const bool extra = p->bias != nullptr || p->zero_mask != nullptr;
    const bool drop = p->dropout_p > 0.f;
    auto kq = extra ? (drop ? sdpa2::bwd_dq_kernel<true, true> : sdpa2::bwd_dq_kernel<true, false>) : (drop ? sdpa2::bwd_dq_kernel<false, true> : sdpa2::bwd_dq_kernel<false, false>);
    auto kk = extra ? (drop ? sdpa2::bwd_dkdv_kernel<true, true> : sdpa2::bwd_dkdv_kernel<true, false>) : (drop ? sdpa2::bwd_dkdv_kernel<false, true> : sdpa2::bwd_dkdv_kernel<false, false>);
    if ((rc = allow_smem(kq, smem_q))) return rc;
    if ((rc = allow_smem(kk, smem_k))) return rc;
    // TMA tensor maps: 128-row boxes for the stationary tiles, 64-row boxes for the streamed ones
    const int64_t sY = (int64_t)p->H * p->dk, sYb = (int64_t)p->Nq * sY;
    CUtensorMap tmQ, tmdO, tmK, tmV, tmQs, tmdOs, tmKL, tmVL;
    if ((rc = make_tile_map_sw(&tmQ, p->q, p->B, p->Nq, p->H, p->dk, p->q_sb, p->q_sn, p->q_sh, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmdO, p->dy, p->B, p->Nq, p->H, p->dk, sYb, sY, p->dk, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmK, p->k, p->B, p->Nk, p->H, p->dk, p->k_sb, p->k_sn, p->k_sh, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmV, p->v, p->B, p->Nk, p->H, p->dk, p->v_sb, p->v_sn, p->v_sh, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmQs, p->q, p->B, p->Nq, p->H, p->dk, p->q_sb, p->q_sn, p->q_sh, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmdOs, p->dy, p->B, p->Nq, p->H, p->dk, sYb, sY, p->dk, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmKL, p->k, p->B, p->Nk, p->H, p->dk, p->k_sb, p->k_sn, p->k_sh, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmVL, p->v, p->B, p->Nk, p->H, p->dk, p->v_sb, p->v_sn, p->v_sh, 128))) return rc;
    kq<<<p->B * p->H * ((p->Nq + 127) / 128), 256, smem_q, st2>>>(*p, delta, tmQ, tmdO, tmK, tmV);
    kk<<<p->B * p->H * ((p->Nk + 127) / 128), 256, smem_k, st2>>>(*p, delta, tmQs, tmdOs, tmKL, tmVL);
    MOP_CHECK_CUDA(cudaGetLastError());
    p->impl_used = MOP_IMPL_TCGEN05;
    return MOP_OK;
  }
  const size_t need = sdpa::bwd_workspace_floats(p) * sizeof(float);
  MOP_REQUIRE(p->workspace && p->workspace_bytes >= need, MOP_EWORKSPACE, "workspace too small: have %zu, need %zu", p->workspace_bytes, need);
  const size_t smem = sdpa::smem_bytes(p->dk);
  const int gk = p->B * p->H * ((p->Nk + sdpa::TK - 1) / sdpa::TK);
  const int gq = p->B * p->H * ((p->Nq + sdpa::TQ - 1) / sdpa::TQ);
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(p->workspace);
  if (p->dtype == MOP_F32) {
    if ((rc = allow_smem(sdpa::bwd_dkdv_kernel<float>, smem))) return rc;
    if ((rc = allow_smem(sdpa::bwd_dq_kernel<float>, smem))) return rc;
    sdpa::bwd_dkdv_kernel<float><<<gk, simt::kThreads, smem, st>>>(*p, ws);
    sdpa::bwd_dq_kernel<float><<<gq, simt::kThreads, smem, st>>>(*p, ws);
  } else {
    if ((rc = allow_smem(sdpa::bwd_dkdv_kernel<__nv_bfloat16>, smem))) return rc;
    if ((rc = allow_smem(sdpa::bwd_dq_kernel<__nv_bfloat16>, smem))) return rc;
    sdpa::bwd_dkdv_kernel<__nv_bfloat16><<<gk, simt::kThreads, smem, st>>>(*p, ws);
    sdpa::bwd_dq_kernel<__nv_bfloat16><<<gq, simt::kThreads, smem, st>>>(*p, ws);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  p->impl_used = MOP_IMPL_SIMT;
  return MOP_OK;
}
}  // extern "C"
