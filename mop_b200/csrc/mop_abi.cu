// libmop_b200.so - C ABI entry points (see include/mop_b200.h).
// Host side only validates, sizes scratch and enqueues kernels on the caller's stream.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "edgewise_simt.cuh"
#include "edgewise_tc.cuh"
#include "edgewise_tc_bwd2.cuh"
#include "edgewise_tc_large.cuh"
#include "edgewise_tc_large_bwd.cuh"
#include "quartet_simt.cuh"
#include "quartet_tc.cuh"
#include "sdpa_tc2.cuh"
#include "sdpa_simt.cuh"
#include "tc_selftest.cuh"

namespace mop {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return MOP_ECUDA;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

// ---------------------------------------------------------------------------
// TMA tensor maps (tc_common.cuh: tma_load_tile).  The driver entry point is looked up at run time: no link-time libcuda.
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<TensorMapEncodeFn>(f);
  }();
  return fn;
}
// bf16 tensor [B][N][H][dk] with element strides (sb, sn, sh), last dimension contiguous, described to the TMA unit as
// (column, token, head, batch); box = 64 columns x box_rows tokens, 128-byte swizzle (tc_common.cuh: tma_load_tile_sw)
int make_tile_map_sw(CUtensorMap* tm, const void* base, int B, int N, int H, int dk, int64_t sb, int64_t sn, int64_t sh, int box_rows) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  MOP_REQUIRE(enc != nullptr, MOP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  auto stride = [](int64_t elems, int dim) -> cuuint64_t { return (dim > 1 && elems > 0) ? (cuuint64_t)elems * 2 : 16; };
  const cuuint64_t dims[4] = {(cuuint64_t)dk, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {stride(sn, N), stride(sh, H), stride(sb, B)};
  const cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MOP_REQUIRE(rc == CUDA_SUCCESS, MOP_ECUDA, "cuTensorMapEncodeTiled (128B swizzle) failed with code %d (N=%d H=%d dk=%d strides %lld %lld %lld)", (int)rc, N,
              H, dk, (long long)sb, (long long)sn, (long long)sh);
  return MOP_OK;
}

// ---------------------------------------------------------------------------
static int check_edgewise(const MopEdgewiseParams* p, bool bwd) {
  MOP_REQUIRE(p != nullptr, MOP_EINVAL, "params is NULL");
  MOP_REQUIRE(p->struct_bytes == (int32_t)sizeof(MopEdgewiseParams), MOP_EABI,
              "MopEdgewiseParams size mismatch: caller %d, library %d", p->struct_bytes, (int)sizeof(MopEdgewiseParams));
  MOP_REQUIRE(p->dtype == MOP_F32 || p->dtype == MOP_BF16, MOP_EINVAL, "bad dtype %d", p->dtype);
  MOP_REQUIRE(p->B > 0 && p->H > 0 && p->N > 0 && p->dk > 0, MOP_EINVAL, "bad shape B=%d H=%d N=%d dk=%d", p->B, p->H, p->N, p->dk);
  MOP_REQUIRE(p->V >= 2 && p->V <= ew::kMaxViews, MOP_EUNSUPPORTED, "n_views=%d outside [2,%d]", p->V, ew::kMaxViews);
  MOP_REQUIRE(p->Vp == 1 || p->Vp == p->V, MOP_EINVAL, "Vp must be 1 or V");
  MOP_REQUIRE(p->gate_mode == MOP_GATE_DENSE || p->gate_mode == MOP_GATE_LOWRANK, MOP_EINVAL, "bad gate_mode %d", p->gate_mode);
  MOP_REQUIRE(p->qkv && p->y && p->chain_value_logit, MOP_EINVAL, "qkv / y / chain_value_logit must be set");
  MOP_REQUIRE((p->q_scale != nullptr) == (p->k_scale != nullptr) && (p->q_scale != nullptr) == (p->v_scale != nullptr),
              MOP_EINVAL, "q/k/v_scale must be all set or all NULL");
  if (p->gate_mode == MOP_GATE_LOWRANK) {
    MOP_REQUIRE(p->gate_rank >= 1 && p->gate_rank <= ew::kMaxRank, MOP_EUNSUPPORTED, "gate_rank=%d outside [1,%d]", p->gate_rank, ew::kMaxRank);
    MOP_REQUIRE(p->row_w && p->row_b && p->col_w && p->col_b, MOP_EINVAL, "lowrank head tensors missing");
  } else {
    MOP_REQUIRE(p->hidden >= 1 && p->hidden <= ew::kMaxHidden, MOP_EUNSUPPORTED, "hidden=%d outside [1,%d]", p->hidden, ew::kMaxHidden);
    MOP_REQUIRE(p->conv1_w && p->conv1_b && p->conv2_w && p->conv2_b, MOP_EINVAL, "dense head tensors missing");
    MOP_REQUIRE(!p->use_k3 || (p->mid3_w && p->mid3_b), MOP_EINVAL, "use_k3 set but mid3 tensors missing");
  }
  if (bwd) {
    MOP_REQUIRE(p->dy && p->dqkv && p->dhead_part && p->dlogit_part, MOP_EINVAL, "backward buffers missing");
    MOP_REQUIRE((p->q_scale == nullptr) || p->dscale_part, MOP_EINVAL, "dscale_part missing");
  }
  return MOP_OK;
}

static int edgewise_grid(const MopEdgewiseParams* p) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;  // sizing only (no device): B200
  int G = p->B * p->H;
  return G < 2 * sms ? G : 2 * sms;
}

static ew::Layout edgewise_layout(const MopEdgewiseParams* p, int bwd) {
  ew::Layout L;
  const bool dense = p->gate_mode == MOP_GATE_DENSE;
  L.build(p->N, p->dk, p->V, p->Vp, dense ? 1 : p->gate_rank, dense ? p->hidden : 1, dense ? 1 : 0,
          dense && p->use_k3 ? 1 : 0, bwd);
  return L;
}

}  // namespace mop

using namespace mop;

extern "C" {

int mop_abi_version(void) { return MOP_ABI_VERSION; }
const char* mop_last_error(void) { return g_err; }
int mop_device_sm_count(void) {
  int n = sm_count();
  if (n < 0) { set_error("no CUDA device"); return MOP_ECUDA; }
  return n;
}

size_t mop_edgewise_head_param_count(const MopEdgewiseParams* p) {
  if (!p) return 0;
  const bool dense = p->gate_mode == MOP_GATE_DENSE;
  return ew::head_param_count(p->gate_mode, p->V, dense ? 1 : p->gate_rank, p->hidden, dense && p->use_k3);
}

int mop_edgewise_needs_row_stats(const MopEdgewiseParams* p) {
  if (check_edgewise(p, false) != MOP_OK) return 0;
  return (p->impl != MOP_IMPL_SIMT && !ewtc::supported(p) && ewl::supported(p)) ? 1 : 0;
}

size_t mop_edgewise_workspace_bytes(const MopEdgewiseParams* p, int backward) {
  if (check_edgewise(p, false) != MOP_OK) return 0;
  if (p->impl != MOP_IMPL_SIMT && !ewtc::supported(p) && ewl::supported(p)) {
    // per-CTA scratch slots of bf16 map images (forward: the V per-view softmax maps; backward: see edgewise_tc_large_bwd.cuh)
    if (!backward) return (size_t)ewl::grid_size(p) * ewtc::kMaxV * ewl::kBufA;
    if (p->row_stats && p->y_base) return (size_t)ewl::grid_size(p) * ewl::kBwdSlots * ewl::kBufA;
  }
  ew::Layout L = edgewise_layout(p, backward ? 1 : 0);
  return (size_t)edgewise_grid(p) * L.total * sizeof(float);
}

static int edgewise_launch(MopEdgewiseParams* p, void* stream, bool bwd) {
  int rc = check_edgewise(p, bwd);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = ewtc::supported(p);
  // token counts up to 200 (ViT-B/16: 196); the backward needs the row statistics saved by the forward
  const bool large_ok = !tc_ok && (bwd ? ewl::supported_bwd(p) : ewl::supported(p));
  MOP_REQUIRE(p->impl == MOP_IMPL_AUTO || p->impl == MOP_IMPL_SIMT || (p->impl == MOP_IMPL_TCGEN05 && (tc_ok || large_ok)), MOP_EUNSUPPORTED,
              "impl %d not available for this shape (tcgen05 path: bf16, dk%%8==0, dk<=64, V<=5, share_qkv, lowrank r<=4; "
              "N=64 forward+backward, N<=200 forward)", p->impl);
  if (large_ok && p->impl != MOP_IMPL_SIMT) {
    const size_t smem = sizeof(ewl::Smem) + 128, smem_b = sizeof(ewl::SmemBwd) + 128;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MOP_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      MOP_CHECK_CUDA(cudaFuncSetAttribute(ewl::edgewise_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MOP_CHECK_CUDA(cudaFuncSetAttribute(ewl::edgewise_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
      configured_dev = dev;
    }
    const size_t need = (size_t)ewl::grid_size(p) * (bwd ? ewl::kBwdSlots : ewtc::kMaxV) * ewl::kBufA;
    MOP_REQUIRE(p->workspace && p->workspace_bytes >= need, MOP_EWORKSPACE, "workspace too small: have %zu, need %zu", p->workspace_bytes, need);
    MOP_REQUIRE((reinterpret_cast<uintptr_t>(p->workspace) & 15) == 0, MOP_EINVAL, "workspace must be 16-byte aligned");
    if (bwd) ewl::edgewise_bwd_kernel<<<ewl::grid_size(p), 256, smem_b, st>>>(*p);
    else ewl::edgewise_fwd_kernel<<<ewl::grid_size(p), 256, smem, st>>>(*p);
    MOP_CHECK_CUDA(cudaGetLastError());
    p->impl_used = MOP_IMPL_TCGEN05;
    return MOP_OK;
  }
  if (tc_ok && p->impl != MOP_IMPL_SIMT) {
    const size_t smem_f = sizeof(ewtc::Smem<false>) + 1024, smem_b = sizeof(ewtc::Smem<true>) + 1024, smem_b2 = sizeof(ewtc::SmemBwd2) + 1024;
    static const bool one_wg = getenv("MOP_EW_BWD_1WG") != nullptr;   // A/B switch: single-warpgroup backward
    static thread_local int configured_dev = -1;
    int dev = 0;
    MOP_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      MOP_CHECK_CUDA(cudaFuncSetAttribute(ewtc::edgewise_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
      MOP_CHECK_CUDA(cudaFuncSetAttribute(ewtc::edgewise_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
      MOP_CHECK_CUDA(cudaFuncSetAttribute(ewtc::edgewise_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b2));
      configured_dev = dev;
    }
    const int G = p->B * p->H, sms = sm_count();
    const int grid = bwd ? (G < sms ? G : sms) : (G < 2 * sms ? G : 2 * sms);   // forward: two CTAs per SM
    if (bwd && !one_wg) ewtc::edgewise_bwd2_kernel<<<grid, 256, smem_b2, st>>>(*p);
    else if (bwd) ewtc::edgewise_kernel<true><<<grid, 128, smem_b, st>>>(*p);
    else ewtc::edgewise_kernel<false><<<grid, 128, smem_f, st>>>(*p);
    MOP_CHECK_CUDA(cudaGetLastError());
    p->impl_used = MOP_IMPL_TCGEN05;
    return MOP_OK;
  }
  ew::Layout L = edgewise_layout(p, bwd ? 1 : 0);
  const int grid = edgewise_grid(p);
  const size_t need = (size_t)grid * L.total * sizeof(float);
  MOP_REQUIRE(p->workspace && p->workspace_bytes >= need, MOP_EWORKSPACE, "workspace too small: have %zu, need %zu", p->workspace_bytes, need);
  float* ws = reinterpret_cast<float*>(p->workspace);
  if (p->dtype == MOP_F32) {
    if (bwd) ew::bwd_kernel<float><<<grid, simt::kThreads, 0, st>>>(*p, L, ws);
    else ew::fwd_kernel<float><<<grid, simt::kThreads, 0, st>>>(*p, L, ws);
  } else {
    if (bwd) ew::bwd_kernel<__nv_bfloat16><<<grid, simt::kThreads, 0, st>>>(*p, L, ws);
    else ew::fwd_kernel<__nv_bfloat16><<<grid, simt::kThreads, 0, st>>>(*p, L, ws);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  p->impl_used = MOP_IMPL_SIMT;
  return MOP_OK;
}

int mop_edgewise_fwd(MopEdgewiseParams* p, void* stream) { return edgewise_launch(p, stream, false); }
int mop_edgewise_bwd(MopEdgewiseParams* p, void* stream) { return edgewise_launch(p, stream, true); }

}  // extern "C"

// ---------------------------------------------------------------------------
// SDPA
// ---------------------------------------------------------------------------
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
template <typename K> static int allow_smem(K kernel, size_t bytes) {
  MOP_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return MOP_OK;
}
}  // namespace mop

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
    if ((rc = allow_smem(extra ? sdpa2::fwd_kernel<true> : sdpa2::fwd_kernel<false>, smem_tc))) return rc;
    CUtensorMap tmQ, tmK, tmV;
    if ((rc = make_tile_map_sw(&tmQ, p->q, p->B, p->Nq, p->H, p->dk, p->q_sb, p->q_sn, p->q_sh, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmK, p->k, p->B, p->Nk, p->H, p->dk, p->k_sb, p->k_sn, p->k_sh, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmV, p->v, p->B, p->Nk, p->H, p->dk, p->v_sb, p->v_sn, p->v_sh, 64))) return rc;
    (extra ? sdpa2::fwd_kernel<true> : sdpa2::fwd_kernel<false>)<<<p->B * p->H * ((p->Nq + 127) / 128), 192, smem_tc, st>>>(*p, tmQ, tmK, tmV);
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
    if ((rc = allow_smem(extra ? sdpa2::bwd_dq_kernel<true> : sdpa2::bwd_dq_kernel<false>, smem_q))) return rc;
    if ((rc = allow_smem(extra ? sdpa2::bwd_dkdv_kernel<true> : sdpa2::bwd_dkdv_kernel<false>, smem_k))) return rc;
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
    (extra ? sdpa2::bwd_dq_kernel<true> : sdpa2::bwd_dq_kernel<false>)<<<p->B * p->H * ((p->Nq + 127) / 128), 256, smem_q, st2>>>(*p, delta, tmQ, tmdO, tmK, tmV);
    (extra ? sdpa2::bwd_dkdv_kernel<true> : sdpa2::bwd_dkdv_kernel<false>)<<<p->B * p->H * ((p->Nk + 127) / 128), 256, smem_k, st2>>>(*p, delta, tmQs, tmdOs, tmKL, tmVL);
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

// ---------------------------------------------------------------------------
// Quartet
// ---------------------------------------------------------------------------
}  // extern "C" (reopened below)
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
  if (BH * w.nm >= 2 * sm_count()) {   // enough (b, h, map) problems to fill the GPU: one CTA each, one launch
    qtc::prep_fused_kernel<<<BH * w.nm, 256, 0, st>>>(*p, w, ws);
  } else {   // key-side preparation split over groups of kPrepRows keys (quartet_tc.cuh)
    const int groups = (p->T + qtc::kPrepRows - 1) / qtc::kPrepRows;
    MOP_CHECK_CUDA(cudaMemsetAsync(ws + w.ksum, 0, w.gacc + (size_t)w.nm * BH * 64 * 64 * 4 - w.ksum, st));   // ksum and gacc are adjacent
    qtc::prep_sum_kernel<<<BH * w.nm * groups, 256, 0, st>>>(*p, w, ws);
    qtc::prep_kernel<<<BH * w.nm * groups, 256, 0, st>>>(*p, w, ws);
    qtc::acc_to_tiles_kernel<<<BH * w.nm, 256, 0, st>>>(ws, w.gacc, w.gram);
  }
  const bool hm = p->add_mask != nullptr;
  // TMA tensor maps: activations [B,T,H,dk] (contiguous) and the centred keys in the workspace ([nm*B*H, T, 1, 64])
  const int64_t sT = (int64_t)p->H * p->dk, sB = (int64_t)p->T * sT;
  CUtensorMap tmQ, tmQ2, tmKc, tmV;
  if ((rc = make_tile_map_sw(&tmQ, p->q, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 128))) return rc;
  if ((rc = make_tile_map_sw(&tmQ2, p->use_quartet ? p->q2 : p->q, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 128))) return rc;
  if ((rc = make_tile_map_sw(&tmKc, ws + w.kc, w.nm * BH, p->T, 1, 64, (int64_t)p->T * 64, 64, 64, 64))) return rc;
  if ((rc = make_tile_map_sw(&tmV, p->v, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 64))) return rc;
  if (!bwd) {
    if ((rc = allow_smem(hm ? qtc::fwd_kernel<true> : qtc::fwd_kernel<false>, smem_f))) return rc;
    (hm ? qtc::fwd_kernel<true> : qtc::fwd_kernel<false>)<<<BH * w.nqb, 192, smem_f, st>>>(*p, w, ws, tmQ, tmQ2, tmKc, tmV);
  } else {
    if ((rc = allow_smem(hm ? qtc::bwd_dq_kernel<true> : qtc::bwd_dq_kernel<false>, smem_q))) return rc;
    if ((rc = allow_smem(hm ? qtc::bwd_dkdv_kernel<true> : qtc::bwd_dkdv_kernel<false>, smem_k))) return rc;
    CUtensorMap tmdO, tmQs, tmQ2s, tmdOs, tmKcL, tmVL;   // dO (128-row box); 64-row boxes of q, q2, dO; 128-row boxes of kc, v
    if ((rc = make_tile_map_sw(&tmdO, p->dy, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmQs, p->q, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmQ2s, p->use_quartet ? p->q2 : p->q, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmdOs, p->dy, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 64))) return rc;
    if ((rc = make_tile_map_sw(&tmKcL, ws + w.kc, w.nm * BH, p->T, 1, 64, (int64_t)p->T * 64, 64, 64, 128))) return rc;
    if ((rc = make_tile_map_sw(&tmVL, p->v, p->B, p->T, p->H, p->dk, sB, sT, p->dk, 128))) return rc;
    (hm ? qtc::bwd_dq_kernel<true> : qtc::bwd_dq_kernel<false>)<<<BH * w.nqb, 256, smem_q, st>>>(*p, w, ws, tmQ, tmQ2, tmdO, tmKc, tmV);
    const int nct = (p->T + 63) / 64, cpg = 4, groups = (nct + cpg - 1) / cpg;
    if (BH * w.nm >= 2 * sm_count() || groups == 1) {
      qtc::gmat_kernel<<<BH * w.nm, 256, 0, st>>>(*p, w, ws, 1, nct);
    } else {   // split over groups of four 64-row chunks: partial sums by atomics, then the tile images
      MOP_CHECK_CUDA(cudaMemsetAsync(ws + w.macc, 0, (size_t)w.nm * BH * 64 * 64 * 4, st));
      qtc::gmat_kernel<<<BH * w.nm * groups, 256, 0, st>>>(*p, w, ws, groups, cpg);
      qtc::acc_to_tiles_kernel<<<BH * w.nm, 256, 0, st>>>(ws, w.macc, w.mmat);
    }
    (hm ? qtc::bwd_dkdv_kernel<true> : qtc::bwd_dkdv_kernel<false>)<<<BH * w.nqb, 256, smem_k, st>>>(*p, w, ws, tmQs, tmQ2s, tmdOs, tmKcL, tmVL);
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

// ---------------------------------------------------------------------------
// tcgen05 primitive self-test (tests/test_gpu_tc_primitives.py)
// ---------------------------------------------------------------------------
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
