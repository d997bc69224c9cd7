// libmop_b200.so - C ABI entry points (see include/mop_b200.h): error plumbing, device queries, TMA tensor-map encoder
// Host side only validates, sizes scratch and enqueues kernels on the caller's stream.
#include "abi_host.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

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
// The encoded descriptors are kept in a small per-thread cache keyed by every argument (a training loop asks for the same ~10
// maps every step - the caching allocator hands the same addresses back -, and each driver call costs about a microsecond of host
// time per map: eight per plain-attention backward).  A descriptor depends on nothing but these arguments.
struct TileMapKey {
  const void* base; int B, N, H, dk, box_rows; int64_t sb, sn, sh;
  bool operator==(const TileMapKey& o) const {
    return base == o.base && B == o.B && N == o.N && H == o.H && dk == o.dk && box_rows == o.box_rows && sb == o.sb && sn == o.sn && sh == o.sh;
  }
};
int make_tile_map_sw(CUtensorMap* tm, const void* base, int B, int N, int H, int dk, int64_t sb, int64_t sn, int64_t sh, int box_rows) {
  constexpr int kSlots = 64;
  thread_local TileMapKey keys[kSlots];
  thread_local CUtensorMap maps[kSlots];
  thread_local int used = 0, next = 0;
  const TileMapKey key{base, B, N, H, dk, box_rows, sb, sn, sh};
  for (int i = 0; i < used; ++i)
    if (keys[i] == key) { *tm = maps[i]; return MOP_OK; }
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
  const int slot = used < kSlots ? used++ : (next = (next + 1) % kSlots);   // round-robin replacement once full
  keys[slot] = key;
  maps[slot] = *tm;
  return MOP_OK;
}
}  // namespace mop

namespace mop {
static __global__ void dropout_mask_kernel(float* out, int BH, int Nq, int Nk, Dropout d) {
  const size_t n = (size_t)BH * Nq * Nk;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t j = (uint32_t)(idx % Nk), i = (uint32_t)((idx / Nk) % Nq), bh = (uint32_t)(idx / ((size_t)Nk * Nq));
    out[idx] = d.on ? dropout_factor(d, dropout_row_key(d, bh, i), j) : 1.f;
  }
}
}  // namespace mop

using namespace mop;

extern "C" {
int mop_dropout_mask(float* out, int BH, int Nq, int Nk, float p, uint64_t seed, uint64_t offset, void* stream) {
  MOP_REQUIRE(out && BH > 0 && Nq > 0 && Nk > 0 && p >= 0.f && p <= 1.f, MOP_EINVAL, "bad dropout-mask arguments");
  MOP_REQUIRE(sm_count() > 0, MOP_ECUDA, "no CUDA device");
  dropout_mask_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(out, BH, Nq, Nk, make_dropout(p, seed, offset));
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
int mop_abi_version(void) { return MOP_ABI_VERSION; }
const char* mop_last_error(void) { return g_err; }
int mop_device_sm_count(void) {
  int n = sm_count();
  if (n < 0) { set_error("no CUDA device"); return MOP_ECUDA; }
  return n;
}
}  // extern "C"
