// libmop_b200.so - C ABI entry points (see include/mop_b200.h): Quartet causal attention
// Host side only validates, sizes scratch and enqueues kernels on the caller's stream.
#include "abi_host.h"
#include "quartet_simt.cuh"
#include "quartet_tc.cuh"

using namespace mop;

namespace mop {
static int check_quartet(const MopQuartetParams* p, bool bwd) {
  MOP_REQUIRE(p != nullptr, MOP_EINVAL, "params is NULL");
  MOP_REQUIRE(p->struct_bytes == (int32_t)sizeof(MopQuartetParams), MOP_EABI,
              "MopQuartetParams size mismatch: caller %d, library %d", p->struct_bytes, (int)sizeof(MopQuartetParams));
  MOP_REQUIRE(p->dtype == MOP_F32 || p->dtype == MOP_BF16, MOP_EINVAL, "bad dtype %d", p->dtype);
  MOP_REQUIRE(p->B > 0 && p->H > 0 && p->T > 1 && p->dk > 0, MOP_EINVAL, "bad shape (T must be >= 2: unbiased std)");
  MOP_REQUIRE(p->dk <= quartet::kMaxDk, MOP_EUNSUPPORTED, "head dim %d > %d", p->dk, quartet::kMaxDk);
  MOP_REQUIRE(p->q && p->k && p->v && p->y, MOP_EINVAL, "q/k/v/y must be set");
  if (p->use_quartet) MOP_REQUIRE(p->q2 && p->k2 && p->mixture && p->quartet_scale, MOP_EINVAL, "quartet tensors missing");
  if (bwd) {
    MOP_REQUIRE(p->dy && p->dq && p->dk_ && p->dv && p->stats, MOP_EINVAL, "backward buffers missing");
    if (p->use_quartet) MOP_REQUIRE(p->dq2 && p->dk2 && p->dscalar_part, MOP_EINVAL, "quartet backward buffers missing");
  }
  return MOP_OK;
}
template <typename T>
static int quartet_run(MopQuartetParams* p, cudaStream_t st, bool bwd) {
  const quartet::Ws w = quartet::layout(p, bwd ? 1 : 0);
  float* ws = reinterpret_cast<float*>(p->workspace);
  const size_t smem = quartet::smem_bytes(p->dk);
  const int BH = p->B * p->H;
  int rc;
  quartet::prep_kernel<T><<<BH * w.nm, simt::kThreads, 0, st>>>(*p, w, ws);
  if (!bwd) {
    if ((rc = allow_smem(quartet::fwd_kernel<T>, smem))) return rc;
    quartet::fwd_kernel<T><<<BH * w.nqb, simt::kThreads, smem, st>>>(*p, w, ws);
  } else {
    if ((rc = allow_smem(quartet::bwd_dq_kernel<T>, smem))) return rc;
    if ((rc = allow_smem(quartet::bwd_dkdv_kernel<T>, smem))) return rc;
    quartet::bwd_dq_kernel<T><<<BH * w.nqb, simt::kThreads, smem, st>>>(*p, w, ws);
    quartet::gmat_kernel<<<BH * w.nm, simt::kThreads, 0, st>>>(*p, w, ws);
    quartet::bwd_dkdv_kernel<T><<<BH * w.nqb, simt::kThreads, smem, st>>>(*p, w, ws);
    quartet::finish_kernel<T><<<BH * w.nm, simt::kThreads, 0, st>>>(*p, w, ws);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  p->impl_used = MOP_IMPL_SIMT;
  return MOP_OK;
}
// tcgen05 path: prep -> fwd  |  prep -> bwd_dq -> gmat -> bwd_dkdv -> finish  (quartet_tc.cuh)
static int quartet_run_tc(MopQuartetParams* p, cudaStream_t st, bool bwd) {
  const qtc::Ws w = qtc::layout(p, bwd ? 1 : 0);
  unsigned char* ws = reinterpret_cast<unsigned char*>(p->workspace);
  const int BH = p->B * p->H;
  // the backward kernels own all 512 TMEM columns of their SM: ask for more than half of the shared memory
  const size_t smem_f = sizeof(qtc::SmemF) + 128, one_per_sm = 117 * 1024;
  const size_t smem_q = sizeof(qtc::SmemQ) + 128 > one_per_sm ? sizeof(qtc::SmemQ) + 128 : one_per_sm;
  const size_t smem_k = sizeof(qtc::SmemK) + 128 > one_per_sm ? sizeof(qtc::SmemK) + 128 : one_per_sm;
  int rc;
  // the key preparation of the forward call, if the caller kept that workspace (same layout prefix): nothing to redo
  const bool reuse = bwd && p->fwd_workspace != nullptr && p->fwd_workspace_bytes >= qtc::layout(p, 0).total &&
                     (reinterpret_cast<uintptr_t>(p->fwd_workspace) & 255) == 0;
  unsigned char* pws = reuse ? reinterpret_cast<unsigned char*>(const_cast<void*>(p->fwd_workspace)) : ws;
  if (reuse) {
    MOP_CHECK_CUDA(cudaMemsetAsync(ws + w.dksum, 0, (size_t)w.nm * BH * 64 * 4, st));   // otherwise zeroed by the prep kernels
  } else if (BH * w.nm >= 2 * sm_count()) {   // enough (b, h, map) problems to fill the GPU: one CTA each, one launch
    qtc::prep_fused_kernel<<<BH * w.nm, 256, 0, st>>>(*p, w, ws);
  } else {   // key-side preparation split over groups of kPrepRows keys (quartet_tc.cuh)
    const int groups = (p->T + qtc::kPrepRows - 1) / qtc::kPrepRows;
    MOP_CHECK_CUDA(cudaMemsetAsync(ws + w.ksum, 0, w.gacc + (size_t)w.nm * BH * 64 * 64 * 4 - w.ksum, st));   // ksum and gacc are adjacent
    qtc::prep_sum_kernel<<<BH * w.nm * groups, 256, 0, st>>>(*p, w, ws);
    qtc::prep_kernel<<<BH * w.nm * groups, 256, 0, st>>>(*p, w, ws);
    qtc::acc_to_tiles_kernel<<<BH * w.nm, 256, 0, st>>>(ws, w.gacc, w.gram);
  }
  const bool hm = p->add_mask != nullptr, drop = p->dropout_p > 0.f;   // compiled with / without the mask and the dropout code
  // TMA tensor maps: activations [B,T,H,dk] (contiguous) and the centred keys in the workspace ([nm*B*H, T, 1, 64])
  const int64_t sT = (int64_t)p->H * p->dk, sB = (int64_t)p->T * sT;
  CUtensorMap tmQ, tmQ2, tmKc, tmV;
  if ((rc = make_tile_map_sw(&tmQ, p->q, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 128))) return rc;
  if ((rc = make_tile_map_sw(&tmQ2, p->use_quartet ? p->q2 : p->q, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 128))) return rc;
  if ((rc = make_tile_map_sw(&tmKc, pws + w.kc, w.nm * BH, p->T, 1, 64, (int64_t)p->T * 64, 64, 64, 64))) return rc;
  if ((rc = make_tile_map_sw(&tmV, p->v, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 64))) return rc;
  if (!bwd) {
    auto kf = hm ? (drop ? qtc::fwd_kernel<true, true> : qtc::fwd_kernel<true, false>) : (drop ? qtc::fwd_kernel<false, true> : qtc::fwd_kernel<false, false>);
    if ((rc = allow_smem(kf, smem_f))) return rc;
    kf<<<BH * w.nqb, 192, smem_f, st>>>(*p, w, ws, tmQ, tmQ2, tmKc, tmV);
  } else {
    auto kq = hm ? (drop ? qtc::bwd_dq_kernel<true, true> : qtc::bwd_dq_kernel<true, false>) : (drop ? qtc::bwd_dq_kernel<false, true> : qtc::bwd_dq_kernel<false, false>);
    auto kk = hm ? (drop ? qtc::bwd_dkdv_kernel<true, true> : qtc::bwd_dkdv_kernel<true, false>) : (drop ? qtc::bwd_dkdv_kernel<false, true> : qtc::bwd_dkdv_kernel<false, false>);
    if ((rc = allow_smem(kq, smem_q))) return rc;
    if ((rc = allow_smem(kk, smem_k))) return rc;
    CUtensorMap tmdO, tmQs, tmQ2s, tmdOs, tmKcL, tmVL;   // dO (128-row box); 64-row boxes of q, q2, dO; 128-row boxes of kc, v
    if ((rc = make_tile_map_sw(&tmdO, p->dy, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmQs, p->q, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmQ2s, p->use_quartet ? p->q2 : p->q, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmdOs, p->dy, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmKcL, pws + w.kc, w.nm * BH, p->T, 1, 64, (int64_t)p->T * 64, 64, 64, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmVL, p->v, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 128))) return rc;
    kq<<<BH * w.nqb, 256, smem_q, st>>>(*p, w, ws, pws, tmQ, tmQ2, tmdO, tmKc, tmV);
    const int nct = (p->T + 63) / 64, cpg = 4, groups = (nct + cpg - 1) / cpg;
    if (BH * w.nm >= 2 * sm_count() || groups == 1) {
      qtc::gmat_kernel<<<BH * w.nm, 256, 0, st>>>(*p, w, ws, 1, nct);
    } else {   // split over groups of four 64-row chunks: partial sums by atomics, then the tile images
      MOP_CHECK_CUDA(cudaMemsetAsync(ws + w.macc, 0, (size_t)w.nm * BH * 64 * 64 * 4, st));
      qtc::gmat_kernel<<<BH * w.nm * groups, 256, 0, st>>>(*p, w, ws, groups, cpg);
      qtc::acc_to_tiles_kernel<<<BH * w.nm, 256, 0, st>>>(ws, w.macc, w.mmat);
    }
    kk<<<BH * w.nqb, 256, smem_k, st>>>(*p, w, ws, tmQs, tmQ2s, tmdOs, tmKcL, tmVL);
    qtc::finish_kernel<<<BH * w.nm * ((p->T + 63) / 64), 256, 0, st>>>(*p, w, ws);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  p->impl_used = MOP_IMPL_TCGEN05;
  return MOP_OK;
}
static int quartet_launch(MopQuartetParams* p, void* stream, bool bwd) {
  int rc = check_quartet(p, bwd);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  const bool tc_ok = qtc::supported(p);
  MOP_REQUIRE(p->impl == MOP_IMPL_AUTO || p->impl == MOP_IMPL_SIMT || (p->impl == MOP_IMPL_TCGEN05 && tc_ok), MOP_EUNSUPPORTED,
              "impl %d not available (tcgen05 path: bf16, dk%%8==0, dk<=64, 16-byte aligned tensors)", p->impl);
  if (tc_ok && p->impl != MOP_IMPL_SIMT) {
    const size_t need_tc = qtc::layout(p, bwd ? 1 : 0).total;
    MOP_REQUIRE(p->workspace && p->workspace_bytes >= need_tc, MOP_EWORKSPACE, "workspace too small: have %zu, need %zu", p->workspace_bytes, need_tc);
    MOP_REQUIRE((reinterpret_cast<uintptr_t>(p->workspace) & 255) == 0, MOP_EINVAL, "workspace must be 256-byte aligned");
    return quartet_run_tc(p, (cudaStream_t)stream, bwd);
  }
  const size_t need = quartet::layout(p, bwd ? 1 : 0).total * sizeof(float);
  MOP_REQUIRE(p->workspace && p->workspace_bytes >= need, MOP_EWORKSPACE, "workspace too small: have %zu, need %zu", p->workspace_bytes, need);
  return p->dtype == MOP_F32 ? quartet_run<float>(p, (cudaStream_t)stream, bwd) : quartet_run<__nv_bfloat16>(p, (cudaStream_t)stream, bwd);
}
}  // namespace mop
extern "C" {
size_t mop_quartet_workspace_bytes(const MopQuartetParams* p, int backward) {
  if (check_quartet(p, false) != MOP_OK) return 0;
  if (p->impl != MOP_IMPL_SIMT && qtc::supported(p)) return qtc::layout(p, backward).total;
  return quartet::layout(p, backward).total * sizeof(float);
}
int mop_quartet_fwd(MopQuartetParams* p, void* stream) { return quartet_launch(p, stream, false); }
int mop_quartet_bwd(MopQuartetParams* p, void* stream) { return quartet_launch(p, stream, true); }

}  // extern "C"
