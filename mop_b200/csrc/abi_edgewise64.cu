// libmop_b200.so - launcher of the third-generation N = 64 Edgewise backward (own translation unit: compiles in parallel)
#include "abi_host.h"
#include "edgewise_n64_bwd.cuh"
#include "edgewise_n64_fwd.cuh"

namespace mop {

bool edgewise_n64_bwd_supported(const MopEdgewiseParams* p) { return ew64::supported_bwd(p); }
int edgewise_n64_bwd_partial_rows(const MopEdgewiseParams* p) {
  int sms = sm_count();
  return ew64::bwd_grid(p, sms > 0 ? sms : 148);
}

int edgewise_n64_bwd_launch(MopEdgewiseParams* p, cudaStream_t st) {
  const size_t smem = sizeof(ew64::Smem) + 1024;
  int rc;
  if ((rc = allow_smem(ew64::edgewise_bwd3_kernel, smem))) return rc;
  MOP_REQUIRE((reinterpret_cast<uintptr_t>(p->qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->dy) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->aux) & 15) == 0,
              MOP_EINVAL, "qkv, dy and aux must be 16-byte aligned");
  const int H = p->H, dk = p->dk, N = p->N;
  const int64_t hd = (int64_t)H * dk;
  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p->qkv);
  CUtensorMap tmQ, tmK, tmV, tmDY;   // [B, N, 3, H, dk] slices and dy [B, N, H, dk]: one 64 x 64 box per (b, h)
  if ((rc = make_tile_map_sw(&tmQ, qkv, p->B, N, H, dk, (int64_t)N * 3 * hd, 3 * hd, dk, 64))) return rc;
  if ((rc = make_tile_map_sw(&tmK, qkv + hd, p->B, N, H, dk, (int64_t)N * 3 * hd, 3 * hd, dk, 64))) return rc;
  if ((rc = make_tile_map_sw(&tmV, qkv + 2 * hd, p->B, N, H, dk, (int64_t)N * 3 * hd, 3 * hd, dk, 64))) return rc;
  if ((rc = make_tile_map_sw(&tmDY, p->dy, p->B, N, H, dk, (int64_t)N * hd, hd, dk, 64))) return rc;
  const int grid = ew64::bwd_grid(p, sm_count());
  ew64::edgewise_bwd3_kernel<<<grid, ew64::kThreads, smem, st>>>(*p, tmQ, tmK, tmV, tmDY);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

int edgewise_n64_fwd_launch(MopEdgewiseParams* p, cudaStream_t st) {
  const size_t smem = sizeof(ew64::SmemF) + 1024;
  int rc;
  if ((rc = allow_smem(ew64::edgewise_fwd3_kernel, smem))) return rc;
  MOP_REQUIRE((reinterpret_cast<uintptr_t>(p->qkv) & 15) == 0, MOP_EINVAL, "qkv must be 16-byte aligned");
  const int H = p->H, dk = p->dk, N = p->N;
  const int64_t hd = (int64_t)H * dk;
  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(p->qkv);
  CUtensorMap tmQ, tmK, tmV;
  if ((rc = make_tile_map_sw(&tmQ, qkv, p->B, N, H, dk, (int64_t)N * 3 * hd, 3 * hd, dk, 64))) return rc;
  if ((rc = make_tile_map_sw(&tmK, qkv + hd, p->B, N, H, dk, (int64_t)N * 3 * hd, 3 * hd, dk, 64))) return rc;
  if ((rc = make_tile_map_sw(&tmV, qkv + 2 * hd, p->B, N, H, dk, (int64_t)N * 3 * hd, 3 * hd, dk, 64))) return rc;
  const int G = p->B * p->H, sms = sm_count();
  const int grid = G < 2 * sms ? G : 2 * sms;   // two CTAs per SM
  ew64::edgewise_fwd3_kernel<<<grid, ew64::kFwdThreads, smem, st>>>(*p, tmQ, tmK, tmV);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

}  // namespace mop
