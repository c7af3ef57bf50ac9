// sm_100a primitives used by the tcgen05/TMA attention kernels: mbarrier, TMA (cp.async.bulk.tensor), TMEM
// allocation, tcgen05.mma / commit / ld / st, shared-memory matrix descriptors and the instruction descriptor.
// Everything is inline PTX; encodings follow the PTX ISA "tcgen05" chapter (cross-checked against the CuTe
// headers cute/arch/mma_sm100_desc.hpp shipped in this image, used as an encoding reference only).
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the driver entry point is resolved at run time)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace hstu {
namespace umma {

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

#ifndef HSTU_WAIT_LIMIT_CLK
#define HSTU_WAIT_LIMIT_CLK 4000000000ll  // bounded wait (~2 s): a protocol bug traps instead of hanging the GPU
#endif

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifdef HSTU_DEBUG_SPIN
// debug build: report the wait site that timed out (line number) instead of trapping, then fall through
__device__ __forceinline__ void mbar_wait_line(uint64_t* bar, uint32_t parity, int line) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 20)) {
      printf("MBAR TIMEOUT line %d block (%d,%d,%d) thread %d parity %u\n", line, (int)blockIdx.x, (int)blockIdx.y,
             (int)blockIdx.z, (int)threadIdx.x, parity);
      return;
    }
  }
}
#define mbar_wait(bar, parity) mbar_wait_line(bar, parity, __LINE__)
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  // bounded in TIME (try_wait may suspend the thread for a while, so a spin COUNT bounds nothing): a protocol bug traps
  // after ~2 s instead of hanging the GPU
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > HSTU_WAIT_LIMIT_CLK) __trap();
  }
}
#endif

// ---------------------------------------------------------------------------------------------
// fences
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// TMA loads (tile mode) into shared memory, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA reduce-add of a shared-memory box into global memory (fp32 add done by the L2 / TMA unit; bulk-group completion)
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMEM allocation (one full warp executes these)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------
enum : int { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };  // UMMA layout_type field values

__host__ __device__ constexpr int swizzle_layout_type(int swizzle_bytes) {
  return swizzle_bytes == 128 ? SWZ_128B : swizzle_bytes == 64 ? SWZ_64B : swizzle_bytes == 32 ? SWZ_32B : SWZ_NONE;
}

// 64-bit shared-memory matrix descriptor.  lbo/sbo in bytes.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

// K-major operand tile: rows x K, K contiguous, stored as boxes of [rows][SW bytes] (row pitch = SW), 8-row swizzle
// atoms of 8*SW bytes.  Descriptor for the 16-element K slice starting `k_byte_off` bytes into the row of the box at
// `box_addr`: SBO = 8*SW (stride between 8-row groups), LBO unused for swizzled K-major (encoded as 16 B).
template <int SW>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t box_addr, uint32_t k_byte_off) {
  return make_smem_desc(box_addr + k_byte_off, 16, 8 * SW, swizzle_layout_type(SW));
}

// MN-major operand tile: K rows x MN, MN contiguous, stored as boxes of [K rows][SW bytes]; a box holds SW/2 elements
// of MN; consecutive boxes (next SW/2 MN elements) are `box_stride` bytes apart.  Descriptor for the 16 K-rows
// starting at row k0 (multiple of 8): SBO = 8*SW (next 8 K rows), LBO = box_stride (next MN block).
template <int SW>
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, uint32_t k0_rows, uint32_t box_stride) {
  return make_smem_desc(tile_addr + k0_rows * SW, box_stride, 8 * SW, swizzle_layout_type(SW));
}

// 32-bit instruction descriptor for kind::f16 (fp32 accumulate).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major, bool a_bf16, bool b_bf16) {
  return (1u << 4)                                  // c_format = F32
         | ((a_bf16 ? 1u : 0u) << 7)                // a_format: 0 = F16, 1 = BF16
         | ((b_bf16 ? 1u : 0u) << 10)               // b_format
         | ((a_mn_major ? 1u : 0u) << 15)           // a_major: 0 = K, 1 = MN
         | ((b_mn_major ? 1u : 0u) << 16)           // b_major
         | ((uint32_t)(N >> 3) << 17)               // n_dim
         | ((uint32_t)(M >> 4) << 24);              // m_dim
}

// ---------------------------------------------------------------------------------------------
// tcgen05.mma (single thread issues), commit, ld / st
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on `bar` once all previously issued tcgen05.mma of this thread have completed (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread t of the warp gets TMEM lane (lane_base + t), columns [col, col+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
#ifdef HSTU_EXP_HALF_LDTM
  if ((taddr & 32u) != 0) return;  // ablation experiment only: skip every other 32-column TMEM load (stale registers)
#endif
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// register-level helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// saturating form (clamps to +-65504 instead of producing inf): used for the fp16 tensor-core operands P / dS
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// In-place bf16 -> fp16 conversion (times `scale`, a power of two) of `bytes` bytes of shared memory by NT threads (t = index
// of the calling thread).  Elementwise and in place, so it is independent of the swizzled tile layout.  fp16 holds every bf16
// significand exactly; magnitudes above 65504 saturate and magnitudes below 2^-24 flush to zero.
template <int NT>
__device__ __forceinline__ void convert_bf16_to_f16_inplace(uint8_t* base, int bytes, int t, float scale) {
#ifdef HSTU_EXP_NO_CONVERT
  if (bytes > 0) return;  // ablation experiment only (wrong numerics): what does the in-place conversion cost?
#endif
  const uint32_t s0 = smem_u32(base);
  constexpr int U = 4;  // independent 16-byte chunks in flight per thread (the loop is latency-bound otherwise)
  for (int off0 = t * 16; off0 < bytes; off0 += U * NT * 16) {
    uint32_t w[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int off = off0 + u * NT * 16;
      if (off < bytes)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]) : "r"(s0 + off));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int off = off0 + u * NT * 16;
      if (off < bytes) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lo = __uint_as_float(w[u][i] << 16) * scale, hi = __uint_as_float(w[u][i] & 0xffff0000u) * scale;
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[u][i]) : "f"(hi), "f"(lo));
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s0 + off), "r"(w[u][0]), "r"(w[u][1]), "r"(w[u][2]), "r"(w[u][3]) : "memory");
      }
    }
  }
}

template <int NREG>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(NREG)); }
template <int NREG>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(NREG)); }

// 16-bit tensor-core operand pair in the format the kernel's MMAs use (kind::f16 needs the SAME format for A and B: a mixed
// fp16 x bf16 instruction descriptor raises "illegal instruction" on B200 -- umma_selftest reports the probe)
template <bool OP_BF16>
__device__ __forceinline__ uint32_t pack_operand(float lo, float hi) {
  return OP_BF16 ? pack_bf16x2(lo, hi) : pack_f16x2_sat(lo, hi);
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#ifdef HSTU_EXP_NO_STS
  if (saddr != 0xffffffffu) return;  // ablation experiment only
#endif
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_b16(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
#ifdef HSTU_EXP_NO_MUFU
  return x * 0.25f;  // ablation experiment only (wrong numerics): takes the MUFU pipe out of the picture
#else
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// Byte offset of the 16-byte chunk `chunk` (0..SW/16-1) of row `row` inside a [rows][SW bytes] swizzled box whose base
// is aligned to 8*SW bytes (the TMA / UMMA 128B/64B/32B swizzle: chunk index XOR (row % 8) masked to the atom width).
template <int SW>
__device__ __forceinline__ uint32_t swizzled_chunk_offset(uint32_t row, uint32_t chunk) {
  constexpr uint32_t kChunks = SW / 16;  // 8, 4, 2
  // Swizzle<B,4,3>: byte-address bits [4, 4+B) ^= bits [7, 7+B).  With a row pitch of SW bytes the bits [7,7+B) of the
  // address are (row * SW / 128) % 2^B, i.e. row%8 for SW=128, (row/2)%4 for SW=64, (row/4)%2 for SW=32.
  uint32_t x = (SW == 128) ? (row & 7u) : (SW == 64) ? ((row >> 1) & 3u) : ((row >> 2) & 1u);
  return row * SW + (((chunk ^ x) & (kChunks - 1)) << 4);
}

}  // namespace umma

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
// 3-D map over a [rows, heads, d] bf16/fp16 tensor with arbitrary row / head strides (elements): dims (d, heads, rows),
// box (box_cols, 1, box_rows), swizzle = box_cols * 2 bytes (32/64/128).  Returns 0 on success.
int make_tmap_rows_heads_f32(CUtensorMap* out, const void* base, long long rows, int heads, int d, int box_cols, int box_rows);
int make_tmap_rows_heads(CUtensorMap* out, const void* base, long long rows, int heads, int d, long long row_stride,
                         long long head_stride, int box_cols, int box_rows);

}  // namespace hstu
