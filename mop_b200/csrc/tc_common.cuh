// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) used by the tensor-core kernels.
//
// Shared-memory operand tiles use ONE physical layout, "chunk-major":
//     byte_offset(r, c) = (c / 8) * (R * 16) + r * 16 + (c % 8) * 2          (bf16, R rows)
// i.e. the matrix is cut into 8-column chunks; inside a chunk every row is 16 contiguous bytes.
// A 8-row x 16-byte block is exactly one UMMA "core matrix" of the no-swizzle canonical layouts,
// so the same tile can be handed to tcgen05.mma either as
//   * a K-major operand   (tile rows = M/N index, tile cols = K index):  SBO = 128,    LBO = R*16
//   * an MN-major operand (tile rows = K index,  tile cols = M/N index): SBO = R*16,   LBO = 128
// which means no kernel in this library ever transposes a tile: X, X^T, A_k and A_k^T are the same bytes.
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched with cudaGetDriverEntryPoint)

#include "common.cuh"

namespace mop {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {   // release.cta: this thread's earlier writes are visible to the waiter
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Bounded wait: a lost arrive must never hang the GPU (gpurun strike) - trap after ~10 s of WALL time instead (%globaltimer;
// an SM-cycle bound of 2 s fired spuriously when the context shared the GPU with another process right after box start).
__device__ __forceinline__ bool mbar_try(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
static __device__ __noinline__ void mbar_wait_slow(uint32_t a, uint32_t parity) {
  const unsigned long long t0 = global_ns();
  const long long c0 = clock64();
  while (!mbar_try(a, parity)) {
    if (global_ns() - t0 > 10000000000ULL) {
      printf("mop_b200: mbarrier wait timed out after %llu ms / %lld SM cycles (block %d thread %d, barrier at shared offset %u, parity %u)\n",
             (global_ns() - t0) / 1000000ULL, clock64() - c0, blockIdx.x, threadIdx.x, a, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  if (mbar_try(a, parity)) return;   // already complete: no clock read, no call
  mbar_wait_slow(a, parity);
}

// ---- bulk (TMA, non-tensor) copies: one thread moves a contiguous block ---------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, completion (bytes) signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources reusable
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // writes performed

// ---- cp.async (LDGSTS): per-thread 16 / 4 byte global -> shared copies tracked by commit / wait groups ------------
// 16-byte global -> shared copy without registers (zero fill when !valid); completion tracked by cp.async groups
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// contiguous block global -> shared by the whole CTA (16-byte pieces, L2-only); caller commits / waits
__device__ __forceinline__ void cp_async_block(void* smem_dst, const void* gsrc, uint32_t bytes) {
  for (uint32_t off = threadIdx.x * 16u; off < bytes; off += blockDim.x * 16u)
    cp_async16(reinterpret_cast<unsigned char*>(smem_dst) + off, reinterpret_cast<const unsigned char*>(gsrc) + off, true);
}

// ---- TMA tensor-map tile loads --------------------------------------------------------------------------------
// A [.., tokens, .., dk] bf16 tensor is described to the TMA unit as (column, token, head, batch); a box {64, R, 1, 1} lands in
// shared memory as R rows of 128 bytes with the 16-byte chunks of row r XOR-swizzled by (r & 7) - the canonical SWIZZLE_128B
// operand layout of tcgen05.mma (descriptors: desc_k_sw / desc_mn_sw), usable K-major (rows = M/N index) and MN-major
// (rows = K index) from the same bytes.  Rows / columns outside the tensor are zero filled; the tile base must be 1024-byte
// aligned.  (Host side: mop::make_tile_map_sw.)  A first version used 16-byte inner boxes (chunk-major tiles): eight times the
// L2 requests and half of every 32-byte sector wasted - lts tag throughput 57 %, TMA latency ~4000 cycles under load.
__device__ __forceinline__ void tma_load_tile_sw(void* smem_dst, const CUtensorMap* tm, int row0, int head, int batch, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row0), "r"(head), "r"(batch), "r"(smem_u32(bar))
      : "memory");
}
// byte offset of element (r, c) inside a SWIZZLE_128B tile of 64 bf16 columns
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((((c >> 3) ^ r) & 7) << 4) + ((c & 7) << 1)); }

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2): one issue slot for two elements --------------------------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)), "l"(*reinterpret_cast<uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  uint64_t rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&rd);
}

// ---- proxies / fences ---------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp) ---------------------------------------------------------------
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle (layout_type 0), Blackwell version field = 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// chunk-major tile with R rows used as a K-major operand starting at K offset k0 (multiple of 8)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, uint32_t R, uint32_t k0) {
  return smem_desc(tile_saddr + (k0 >> 3) * R * 16, R * 16, 128);
}
// chunk-major tile with R rows (= K extent) used as an MN-major operand starting at K offset k0
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, uint32_t R, uint32_t k0) {
  return smem_desc(tile_saddr + k0 * 16, 128, R * 16);
}
// SWIZZLE_128B tiles (rows of 64 bf16 = 128 bytes, 8-row groups of 1024 bytes; layout_type 2)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);
}
// rows = M / N index, the 64 columns = K: K-major operand starting at K offset k0 (multiple of 16)
__device__ __forceinline__ uint64_t desc_k_sw(uint32_t tile_saddr, uint32_t k0) { return smem_desc_sw128(tile_saddr + k0 * 2, 16, 1024); }
// the 64 columns = M / N index, rows = K: MN-major operand starting at K offset (row) k0 (multiple of 16)
__device__ __forceinline__ uint64_t desc_mn_sw(uint32_t tile_saddr, uint32_t k0) { return smem_desc_sw128(tile_saddr + k0 * 128, 1024, 1024); }
// instruction descriptor: kind::f16, bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one K=16 step.  Issued by ONE thread.
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers, 16 lanes x 256 bit shape --------------------------------------------------
// One warp reads 16 TMEM lanes x (8*X) columns.  Thread T receives, for n in [0,X):
//   v[4n+0], v[4n+1] = lane (T/4),     columns 8n + 2(T%4) + {0,1}
//   v[4n+2], v[4n+3] = lane (T/4) + 8, same columns
// (the mma.sync accumulator fragment shape).  taddr = (lane_base << 16) | column.
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x8(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


// ---- TMEM <-> registers, 32 lanes x 32 bit shape ("thread per row") --------------------------------------
// One warp reads its 32 TMEM lanes x X consecutive columns: thread T receives lane (32*(warp%4) + T),
// columns [col, col + X).  With an M=128 accumulator (row i <-> lane i) every thread owns one row.
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
               : "memory");
}

// ---- chunk-major tile addressing (bf16) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t tile_off(uint32_t R, uint32_t r, uint32_t c) { return (c >> 3) * R * 16 + r * 16 + (c & 7) * 2; }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

// Fragment geometry of a 64x64 tile held by 128 threads (4 warps x 16 TMEM lanes):
//   row_lo = 16*warp + lane/4, row_hi = row_lo + 8, columns 8n + 2*(lane%4) + {0,1} for n in [0,8)
struct Frag {
  int warp, lane, row_lo, row_hi, cq;
  __device__ Frag() {
    warp = (threadIdx.x >> 5) & 3;
    lane = threadIdx.x & 31;
    row_lo = 16 * warp + (lane >> 2);
    row_hi = row_lo + 8;
    cq = 2 * (lane & 3);
  }
  __device__ __forceinline__ int col(int n) const { return 8 * n + cq; }
};

// write a 64x64 fp32 fragment as bf16 into a chunk-major 64-row tile (conflict-free 4-byte stores)
__device__ __forceinline__ void frag_store_bf16(unsigned char* tile, const Frag& f, const float* v) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    *reinterpret_cast<uint32_t*>(tile + tile_off(64, f.row_lo, f.col(n))) = pack_bf16(v[4 * n + 0], v[4 * n + 1]);
    *reinterpret_cast<uint32_t*>(tile + tile_off(64, f.row_hi, f.col(n))) = pack_bf16(v[4 * n + 2], v[4 * n + 3]);
  }
}
__device__ __forceinline__ void frag_load_bf16(const unsigned char* tile, const Frag& f, float* v) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(tile + tile_off(64, f.row_lo, f.col(n))));
    float2 b = unpack_bf16(*reinterpret_cast<const uint32_t*>(tile + tile_off(64, f.row_hi, f.col(n))));
    v[4 * n + 0] = a.x; v[4 * n + 1] = a.y; v[4 * n + 2] = b.x; v[4 * n + 3] = b.y;
  }
}

// sum over the 4 threads that share a row
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// ---- issue a 64x64x(16*KSTEPS) GEMM into TMEM: D = op(A) * op(B) ------------------------------------------
// a_mn / b_mn: operand tile is used MN-major (see header comment).  One thread issues.
template <int KSTEPS>
__device__ __forceinline__ void gemm64(uint32_t d_tmem, uint32_t a_tile, bool a_mn, uint32_t b_tile, bool b_mn, bool accumulate) {
  const uint32_t id = idesc_bf16(64, 64, a_mn ? 1u : 0u, b_mn ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) {
    uint64_t ad = a_mn ? desc_mnmajor(a_tile, 64, 16 * k) : desc_kmajor(a_tile, 64, 16 * k);
    uint64_t bd = b_mn ? desc_mnmajor(b_tile, 64, 16 * k) : desc_kmajor(b_tile, 64, 16 * k);
    mma_ss(d_tmem, ad, bd, id, (accumulate || k > 0) ? 1u : 0u);
  }
}

// column sums over the 32 rows of a warp: v[e] is this lane's value of column e (16 columns).  After the
// butterfly lane L (L even) holds the sum of column bitrev-ish index col(L); returns it, *col receives the index.
__device__ __forceinline__ float warp_colsum16(const float* v, int lane, int* col) {
  float a8[8], a4[4], a2[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = h16 ? v[i] : v[i + 8], keep = h16 ? v[i + 8] : v[i];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = h8 ? a8[i] : a8[i + 4], keep = h8 ? a8[i + 4] : a8[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = h4 ? a4[i] : a4[i + 2], keep = h4 ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = h2 ? a2[0] : a2[1], keep = h2 ? a2[1] : a2[0];
  float s = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  *col = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
  return s;
}

}  // namespace tc
}  // namespace mop
