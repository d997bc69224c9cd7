// libmop_b200.so - C ABI entry points (see include/mop_b200.h): Edgewise attention
// Host side only validates, sizes scratch and enqueues kernels on the caller's stream.
#include "abi_host.h"
#include "edgewise_simt.cuh"
#include "edgewise_tc.cuh"
#include "edgewise_tc_large.cuh"
#include "edgewise_tc_large_bwd.cuh"
#include "edgewise_tc_large_fwd2.cuh"

namespace mop {
// ---------------------------------------------------------------------------
static int check_edgewise(const MopEdgewiseParams* p, bool bwd) {
  MOP_REQUIRE(p != nullptr, MOP_EINVAL, "params is NULL");
  MOP_REQUIRE(p->struct_bytes == (int32_t)sizeof(MopEdgewiseParams), MOP_EABI,
              "MopEdgewiseParams size mismatch: caller %d, library %d", p->struct_bytes, (int)sizeof(MopEdgewiseParams));
  MOP_REQUIRE(p->dtype == MOP_F32 || p->dtype == MOP_BF16, MOP_EINVAL, "bad dtype %d", p->dtype);
  MOP_REQUIRE(p->B > 0 && p->H > 0 && p->N > 0 && p->dk > 0, MOP_EINVAL, "bad shape B=%d H=%d N=%d dk=%d", p->B, p->H, p->N, p->dk);
  // V = 1 happens with a single-dilation Q/K lens bank (attention_variants.py:472-498): chain = A_1, all mix terms vanish
  MOP_REQUIRE(p->V >= 1 && p->V <= ew::kMaxViews, MOP_EUNSUPPORTED, "n_views=%d outside [1,%d]", p->V, ew::kMaxViews);
  MOP_REQUIRE(p->Vp == 1 || p->Vp == p->V, MOP_EINVAL, "Vp must be 1 or V");
  MOP_REQUIRE(p->gate_mode == MOP_GATE_DENSE || p->gate_mode == MOP_GATE_LOWRANK || p->gate_mode == MOP_GATE_CONST, MOP_EINVAL,
              "bad gate_mode %d", p->gate_mode);
  MOP_REQUIRE(p->qkv && p->y && p->chain_value_logit, MOP_EINVAL, "qkv / y / chain_value_logit must be set");
  MOP_REQUIRE((p->q_scale != nullptr) == (p->k_scale != nullptr) && (p->q_scale != nullptr) == (p->v_scale != nullptr),
              MOP_EINVAL, "q/k/v_scale must be all set or all NULL");
  if (p->gate_mode == MOP_GATE_CONST) {
    MOP_REQUIRE(p->hops >= 2 && p->hops <= ew::kMaxViews, MOP_EUNSUPPORTED, "hops=%d outside [2,%d]", p->hops, ew::kMaxViews);
  } else if (p->gate_mode == MOP_GATE_LOWRANK) {
    MOP_REQUIRE(p->gate_rank >= 1 && p->gate_rank <= ew::kMaxRank, MOP_EUNSUPPORTED, "gate_rank=%d outside [1,%d]", p->gate_rank, ew::kMaxRank);
    MOP_REQUIRE(p->row_w && p->row_b && p->col_w && p->col_b, MOP_EINVAL, "lowrank head tensors missing");
  } else {
    MOP_REQUIRE(p->hidden >= 1 && p->hidden <= ew::kMaxHidden, MOP_EUNSUPPORTED, "hidden=%d outside [1,%d]", p->hidden, ew::kMaxHidden);
    MOP_REQUIRE(p->conv1_w && p->conv1_b && p->conv2_w && p->conv2_b, MOP_EINVAL, "dense head tensors missing");
    MOP_REQUIRE(!p->use_k3 || (p->mid3_w && p->mid3_b), MOP_EINVAL, "use_k3 set but mid3 tensors missing");
  }
  MOP_REQUIRE(p->lens_n >= 0 && p->lens_n <= ew::kMaxLens, MOP_EUNSUPPORTED, "lens_n=%d outside [0,%d]", p->lens_n, ew::kMaxLens);
  if (p->lens_n > 0) {
    MOP_REQUIRE(p->lens_w != nullptr && p->gate_mode != MOP_GATE_CONST, MOP_EINVAL, "S lens bank needs lens_w and a gate head");
    for (int l = 0; l < p->lens_n; ++l) MOP_REQUIRE(p->lens_dil[l] >= 1, MOP_EINVAL, "lens dilation must be >= 1");
    MOP_REQUIRE(!bwd || p->dlens_part, MOP_EINVAL, "dlens_part missing");
  }
  if (bwd) {
    MOP_REQUIRE(p->dy && p->dqkv && p->dlogit_part && (p->dhead_part || p->gate_mode == MOP_GATE_CONST), MOP_EINVAL, "backward buffers missing");
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
  const bool dense = p->gate_mode == MOP_GATE_DENSE, cg = p->gate_mode == MOP_GATE_CONST;
  L.build(p->N, p->dk, p->V, p->Vp, (dense || cg) ? 1 : p->gate_rank, dense ? p->hidden : 1, dense ? 1 : 0,
          dense && p->use_k3 ? 1 : 0, bwd, cg ? p->hops : 0, cg ? 1 : 0, p->lens_n);
  return L;
}

}  // namespace mop

using namespace mop;

namespace mop {
// Sums of the per-CTA / per-problem gradient partials of one Edgewise backward call in ONE launch (the caller used three
// torch reductions per layer).  Fixed summation order: deterministic.
//   dscale[c][v][h][d] = sum_j dscale_part[(j H + h)][c][v][d]      (row i of the partials belongs to head i % H)
//   dhead[n]           = sum_i dhead_part[i][n]
//   dlogit             = sum_g dlogit_part[g]
static __global__ void __launch_bounds__(256) reduce_partials_kernel(const float* dscale_part, const float* dhead_part, const float* dlogit_part,
                                                                     int R, int H, int V, int dk, int nhead, int G, float* dscale, float* dhead,
                                                                     float* dlogit) {
  const int nscale = dscale_part ? 3 * V * H * dk : 0, nh = dhead_part ? nhead : 0;
  if ((int)blockIdx.x == (int)gridDim.x - 1) {   // last CTA: the scalar
    __shared__ float red[8];
    float a = 0.f;
    for (int g = threadIdx.x; g < G; g += 256) a += dlogit_part[g];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      dlogit[0] = t;
    }
    return;
  }
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx < nscale) {
    const int d = idx % dk, h = (idx / dk) % H, cv = idx / (dk * H);   // cv = c * V + v
    const float* src = dscale_part + (size_t)h * 3 * V * dk + (size_t)cv * dk + d;
    float a = 0.f;
    for (int j = 0; j < R / H; ++j) a += src[(size_t)j * H * 3 * V * dk];
    dscale[idx] = a;
  } else if (idx < nscale + nh) {
    const int n = idx - nscale;
    float a = 0.f;
    for (int i = 0; i < R; ++i) a += dhead_part[(size_t)i * nhead + n];
    dhead[n] = a;
  }
}
}  // namespace mop

extern "C" {
int mop_edgewise_reduce_partials(const float* dscale_part, const float* dhead_part, const float* dlogit_part, int R, int H, int V, int dk,
                                 int nhead, int G, float* dscale, float* dhead, float* dlogit, void* stream) {
  MOP_REQUIRE(dlogit_part && dlogit && R > 0 && H > 0 && R % H == 0 && G > 0, MOP_EINVAL, "bad partial-sum arguments (R=%d H=%d G=%d)", R, H, G);
  MOP_REQUIRE((dscale_part == nullptr) == (dscale == nullptr) && (dhead_part == nullptr) == (dhead == nullptr), MOP_EINVAL, "partial / output pairs must be set together");
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  const int n = (dscale_part ? 3 * V * H * dk : 0) + (dhead_part ? nhead : 0);
  mop::reduce_partials_kernel<<<(n + 255) / 256 + 1, 256, 0, (cudaStream_t)stream>>>(dscale_part, dhead_part, dlogit_part, R, H, V, dk, nhead, G, dscale,
                                                                               dhead, dlogit);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}


size_t mop_edgewise_head_param_count(const MopEdgewiseParams* p) {
  if (!p) return 0;
  if (p->gate_mode == MOP_GATE_CONST) return 0;
  const bool dense = p->gate_mode == MOP_GATE_DENSE;
  return ew::head_param_count(p->gate_mode, p->V, dense ? 1 : p->gate_rank, p->hidden, dense && p->use_k3, p->lens_n);
}

int mop_edgewise_needs_row_stats(const MopEdgewiseParams* p) {
  if (check_edgewise(p, false) != MOP_OK) return 0;
  return (p->impl != MOP_IMPL_SIMT && !ewtc::supported(p) && ewl::supported(p)) ? 1 : 0;
}

int mop_edgewise_partial_rows(const MopEdgewiseParams* p) {
  if (check_edgewise(p, false) != MOP_OK) return 0;
  // the N = 64 tcgen05 backward sums its partials over the problems of a CTA (rows ordered [cta], head = cta % H)
  if (p->impl != MOP_IMPL_SIMT && ewtc::supported(p)) return edgewise_n64_bwd_partial_rows(p);
  return p->B * p->H;
}

size_t mop_edgewise_aux_floats(const MopEdgewiseParams* p) {
  if (check_edgewise(p, false) != MOP_OK) return 0;
  if (p->impl == MOP_IMPL_SIMT) return 0;
  if (ewtc::supported(p)) return (size_t)p->B * p->H * ewtc::kAuxFloats;
  if (ewl::supported(p)) return (size_t)p->B * p->H * ewl::kAuxLFloats;
  return 0;
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
    const size_t smem = sizeof(ewl::Smem2) + 128, smem_b = sizeof(ewl::SmemBwd) + 128;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MOP_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      MOP_CHECK_CUDA(cudaFuncSetAttribute(ewl::edgewise_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MOP_CHECK_CUDA(cudaFuncSetAttribute(ewl::edgewise_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
      configured_dev = dev;
    }
    const size_t need = (size_t)ewl::grid_size(p) * (bwd ? ewl::kBwdSlots : ewtc::kMaxV) * ewl::kBufA;
    MOP_REQUIRE(p->workspace && p->workspace_bytes >= need, MOP_EWORKSPACE, "workspace too small: have %zu, need %zu", p->workspace_bytes, need);
    MOP_REQUIRE((reinterpret_cast<uintptr_t>(p->workspace) & 15) == 0, MOP_EINVAL, "workspace must be 16-byte aligned");
    if (bwd) ewl::edgewise_bwd_kernel<<<ewl::grid_size(p), 256, smem_b, st>>>(*p);
    else ewl::edgewise_fwd2_kernel<<<ewl::grid_size(p), 512, smem, st>>>(*p);
    MOP_CHECK_CUDA(cudaGetLastError());
    p->impl_used = MOP_IMPL_TCGEN05;
    return MOP_OK;
  }
  if (tc_ok && p->impl != MOP_IMPL_SIMT) {   // N = 64: edgewise_n64_fwd.cuh / edgewise_n64_bwd.cuh (abi_edgewise64.cu)
    if (bwd) {
      MOP_REQUIRE(edgewise_n64_bwd_supported(p), MOP_EINVAL,
                  "the N = 64 tcgen05 backward needs `aux` (mop_edgewise_aux_floats() floats written by mop_edgewise_fwd of the same call)");
      rc = edgewise_n64_bwd_launch(p, st);
    } else {
      rc = edgewise_n64_fwd_launch(p, st);
    }
    if (rc != MOP_OK) return rc;
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
