// libmop_b200.so - C ABI entry points (see include/mop_b200.h): fused residual-add + DropPath scale + LayerNorm
#include "abi_host.h"
#include "layernorm.cuh"

namespace mop {
static int check_ln(const MopLnParams* p, bool bwd) {
  MOP_REQUIRE(p != nullptr, MOP_EINVAL, "params is NULL");
  MOP_REQUIRE(p->struct_bytes == (int32_t)sizeof(MopLnParams), MOP_EABI, "MopLnParams size mismatch: caller %d, library %d", p->struct_bytes,
              (int)sizeof(MopLnParams));
  MOP_REQUIRE(p->rows > 0 && p->D > 0 && p->D <= 1024, MOP_EUNSUPPORTED, "rows=%d D=%d (D <= 1024)", p->rows, p->D);
  MOP_REQUIRE((p->r_dtype == MOP_F32 || p->r_dtype == MOP_BF16) && (p->y_dtype == MOP_F32 || p->y_dtype == MOP_BF16), MOP_EINVAL, "bad dtype");
  MOP_REQUIRE(p->x && p->gamma && p->beta && p->mean && p->rstd, MOP_EINVAL, "x / gamma / beta / mean / rstd must be set");
  MOP_REQUIRE(p->scale == nullptr || p->rows_per_sample > 0, MOP_EINVAL, "rows_per_sample must be set with scale");
  if (!bwd) {
    MOP_REQUIRE(p->y != nullptr, MOP_EINVAL, "y must be set");
    MOP_REQUIRE((p->r == nullptr) || p->x_new, MOP_EINVAL, "x_new must be set when a branch r is added");
  } else {
    MOP_REQUIRE(p->dy && p->dx && p->dgamma_part && p->dbeta_part, MOP_EINVAL, "backward buffers missing");
    MOP_REQUIRE((p->r == nullptr) || p->x_new, MOP_EINVAL, "x_new (the normalised tensor) must be passed back when a branch was added");
  }
  return MOP_OK;
}

template <int PL, typename TR, typename TY>
static void launch(const MopLnParams& p, bool bwd, int grid, cudaStream_t st) {
  if (bwd) ln::bwd_kernel<PL, TR, TY><<<grid, ln::kWarps * 32, 0, st>>>(p);
  else ln::fwd_kernel<PL, TR, TY><<<grid, ln::kWarps * 32, 0, st>>>(p);
}
template <int NV, typename TR, typename TY>
static void launch_v(const MopLnParams& p, bool bwd, int grid, cudaStream_t st) {
  if (bwd) ln::bwd_kernel_v<NV, TR, TY><<<grid, ln::BwdWarps<NV>::value * 32, 0, st>>>(p);
  else ln::fwd_kernel_v<NV, TR, TY><<<grid, ln::kWarps * 32, 0, st>>>(p);
}
template <int NV>
static void launch_vt(const MopLnParams& p, bool bwd, int grid, cudaStream_t st) {
  const bool rb = p.r_dtype == MOP_BF16, yb = p.y_dtype == MOP_BF16;
  if (rb && yb) launch_v<NV, __nv_bfloat16, __nv_bfloat16>(p, bwd, grid, st);
  else if (rb) launch_v<NV, __nv_bfloat16, float>(p, bwd, grid, st);
  else if (yb) launch_v<NV, float, __nv_bfloat16>(p, bwd, grid, st);
  else launch_v<NV, float, float>(p, bwd, grid, st);
}
template <int PL>
static void launch_t(const MopLnParams& p, bool bwd, int grid, cudaStream_t st) {
  const bool rb = p.r_dtype == MOP_BF16, yb = p.y_dtype == MOP_BF16;
  if (rb && yb) launch<PL, __nv_bfloat16, __nv_bfloat16>(p, bwd, grid, st);
  else if (rb) launch<PL, __nv_bfloat16, float>(p, bwd, grid, st);
  else if (yb) launch<PL, float, __nv_bfloat16>(p, bwd, grid, st);
  else launch<PL, float, float>(p, bwd, grid, st);
}
static int ln_launch(MopLnParams* p, void* stream, bool bwd) {
  int rc = check_ln(p, bwd);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device (libmop_b200 has no CPU fallback)");
  // forward: one warp per row over a persistent grid.  backward: exactly nparts CTAs (every partial row is written - a CTA without
  // rows writes zeros -, so the caller may sum all of them): mop_ln_partial_rows_d() is the recommended count
  int grid = ln::grid_size(p->rows, sm_count());
  if (bwd) {
    MOP_REQUIRE(p->nparts >= 1, MOP_EWORKSPACE, "dgamma_part / dbeta_part need at least one partial row (nparts=%d)", p->nparts);
    grid = p->nparts < sm_count() * 16 ? p->nparts : sm_count() * 16;
    MOP_REQUIRE(grid == p->nparts, MOP_EWORKSPACE, "nparts=%d partial rows (at most %d)", p->nparts, sm_count() * 16);
  }
  cudaStream_t st = (cudaStream_t)stream;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool aligned = al16(p->x) && al16(p->r) && al16(p->x_new) && al16(p->y) && al16(p->dy) && al16(p->dx_new) && al16(p->dx) && al16(p->dr);
  if (p->D % 8 == 0 && aligned) {   // 16-byte accesses: D % 8 == 0 keeps every row of a 16-byte aligned tensor aligned
    if (p->D <= 256) launch_vt<1>(*p, bwd, grid, st);
    else if (p->D <= 512) launch_vt<2>(*p, bwd, grid, st);
    else if (p->D <= 768) launch_vt<3>(*p, bwd, grid, st);
    else launch_vt<4>(*p, bwd, grid, st);
  } else if (p->D <= 256) launch_t<8>(*p, bwd, grid, st);
  else if (p->D <= 768) launch_t<24>(*p, bwd, grid, st);
  else launch_t<32>(*p, bwd, grid, st);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
}  // namespace mop

using namespace mop;

extern "C" {
int mop_ln_partial_rows(int rows) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  return ln::grid_size(rows, sms);
}
int mop_ln_partial_rows_d(int rows, int D) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  return D % 8 == 0 ? ln::bwd_grid_size_v(rows, D, sms) : ln::grid_size(rows, sms);
}
int mop_ln_fwd(MopLnParams* p, void* stream) { return ln_launch(p, stream, false); }
int mop_ln_bwd(MopLnParams* p, void* stream) { return ln_launch(p, stream, true); }
}
