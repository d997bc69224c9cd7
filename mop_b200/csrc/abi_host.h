// Host-side helpers shared by the translation units of libmop_b200.so.
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched with cudaGetDriverEntryPoint)

#include "common.cuh"

namespace mop {
// bf16 tensor [B][N][H][dk] with element strides (sb, sn, sh), last dimension contiguous, described to the TMA unit as
// (column, token, head, batch); box = 64 columns x box_rows tokens, 128-byte swizzle (tc_common.cuh: tma_load_tile_sw)
int make_tile_map_sw(CUtensorMap* tm, const void* base, int B, int N, int H, int dk, int64_t sb, int64_t sn, int64_t sh, int box_rows);

// opt-in to `bytes` of dynamic shared memory; done once per (kernel, device) and remembered per thread
struct SmemOptIn { const void* fn; int dev; size_t bytes; };
template <typename K> static int allow_smem(K kernel, size_t bytes) {
  static thread_local SmemOptIn seen[64];
  static thread_local int n_seen = 0;
  int dev = 0;
  MOP_CHECK_CUDA(cudaGetDevice(&dev));
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < n_seen; ++i)
    if (seen[i].fn == key && seen[i].dev == dev && bytes <= seen[i].bytes) return MOP_OK;
  MOP_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (n_seen < 64) seen[n_seen++] = SmemOptIn{key, dev, bytes};
  return MOP_OK;
}

// third-generation N = 64 Edgewise backward (abi_edgewise64.cu)
bool edgewise_n64_bwd_supported(const MopEdgewiseParams* p);
int edgewise_n64_bwd_partial_rows(const MopEdgewiseParams* p);
int edgewise_n64_bwd_launch(MopEdgewiseParams* p, cudaStream_t st);
int edgewise_n64_fwd_launch(MopEdgewiseParams* p, cudaStream_t st);
}  // namespace mop
