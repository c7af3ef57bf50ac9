// On-device self test of the tcgen05 / TMA building blocks: libhstu_b200_selftest.so (include/hstu_b200_selftest.h), a TEST
// library of its own -- nothing of this file is linked into the product library.
//
// One small GEMM kernel  D[128, N] = A[128, K] * B[N, K]^T  exercises, with exactly-representable integer data,
// every operand layout the attention kernels rely on:
//   * K-major operands loaded by TMA with 128B / 64B / 32B swizzle (Q, K, dO tiles),
//   * MN-major operands loaded by TMA (V for P.V; Q, dO, K for the backward GEMMs),
//   * A written by the threads themselves into the swizzled layout (P / dS tiles),
//   * A read from TMEM (tcgen05.mma with a TMEM A operand, written by tcgen05.st),
//   * accumulator read-back with tcgen05.ld 32x32b.
// A second micro-benchmark measures SFU (MUFU) throughput of the candidate sigmoid formulations.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "internal.h"
#include "umma.cuh"

extern "C" const char* hstu_selftest_last_error(void);

namespace hstu {
using namespace umma;

struct StCfg {
  int N, K;
  int a_mode;  // 0 K-major TMA, 1 MN-major TMA, 2 K-major written by threads, 3 TMEM
  int b_mode;  // 0 K-major TMA, 1 MN-major TMA
  int sw_a, sw_b;
  int variant;  // 1: swap LBO/SBO of MN-major descriptors (diagnostic)
  int a_f16;    // 1: A holds fp16 values while B stays bf16 (mixed operand formats of kind::f16: the fp16 P / dS operands)
  int c_f16;    // 1: A and B fp16, accumulator fp16 (c_format = F16): N values packed into N / 2 TMEM columns
};

template <int SW>
__device__ uint64_t op_desc(int mode, uint32_t base, int rows_k_or_mn, int k_step, int variant) {
  // mode 0/2: K-major tile of `rows` rows; boxes along K each rows*SW bytes
  if (mode == 0 || mode == 2) {
    const uint32_t kbyte = k_step * 32;
    const uint32_t box = kbyte / SW, off = kbyte % SW;
    return desc_kmajor<SW>(base + box * rows_k_or_mn * SW, off);
  }
  // mode 1: MN-major tile: boxes along MN, each Krows*SW bytes; rows_k_or_mn = K rows of the tile
  uint64_t d = desc_mnmajor<SW>(base, k_step * 16, rows_k_or_mn * SW);
  if (variant == 1) d = make_smem_desc(base + k_step * 16 * SW, 8 * SW, rows_k_or_mn * SW, swizzle_layout_type(SW));
  return d;
}

__device__ uint64_t op_desc_rt(int sw, int mode, uint32_t base, int rows, int k_step, int variant) {
  if (sw == 128) return op_desc<128>(mode, base, rows, k_step, variant);
  if (sw == 64) return op_desc<64>(mode, base, rows, k_step, variant);
  return op_desc<32>(mode, base, rows, k_step, variant);
}

__device__ uint32_t swz_off_rt(int sw, uint32_t row, uint32_t chunk) {
  if (sw == 128) return swizzled_chunk_offset<128>(row, chunk);
  if (sw == 64) return swizzled_chunk_offset<64>(row, chunk);
  return swizzled_chunk_offset<32>(row, chunk);
}

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const __nv_bfloat16* __restrict__ Ag, float* __restrict__ Dg,
                                                            StCfg cfg, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = 128, N = cfg.N, K = cfg.K;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 128 * 2 * 2;  // A tile is at most 128 x 256 bf16... (we cap K at 128 for A) -> 64 KB reserved
  if (tid == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_d = tmem_base;          // accumulator: columns [0, N)
  const uint32_t tmem_a = tmem_base + 256;    // TMEM A operand: columns [256, 256 + K/2)

  if (tid == 0) {
    uint32_t bytes = 0;
    // A
    if (cfg.a_mode == 0) {
      const int cols = cfg.sw_a / 2, nbox = K / cols;
      for (int b = 0; b < nbox; ++b) tma_load_3d(sA + b * M * cfg.sw_a, &tmA, &bar_full, b * cols, 0, 0);
      bytes += M * K * 2;
    } else if (cfg.a_mode == 1) {
      const int cols = cfg.sw_a / 2, nbox = M / cols;
      for (int b = 0; b < nbox; ++b) tma_load_3d(sA + b * K * cfg.sw_a, &tmA, &bar_full, b * cols, 0, 0);
      bytes += M * K * 2;
    }
    // B
    if (cfg.b_mode == 0) {
      const int cols = cfg.sw_b / 2, nbox = K / cols;
      for (int b = 0; b < nbox; ++b) tma_load_3d(sB + b * N * cfg.sw_b, &tmB, &bar_full, b * cols, 0, 0);
    } else {
      const int cols = cfg.sw_b / 2, nbox = N / cols;
      for (int b = 0; b < nbox; ++b) tma_load_3d(sB + b * K * cfg.sw_b, &tmB, &bar_full, b * cols, 0, 0);
    }
    bytes += N * K * 2;
    mbar_arrive_expect_tx(&bar_full, bytes);
  }
  if (cfg.a_mode == 2) {
    // thread = row; write A[row][:] (K-major) into the swizzled boxes by hand, 16 bytes at a time
    const int chunks_per_box = cfg.sw_a / 16;
    for (int c = 0; c < K / 8; ++c) {
      uint4 v = *reinterpret_cast<const uint4*>(Ag + (size_t)tid * K + c * 8);
      const int box = c / chunks_per_box, cc = c % chunks_per_box;
      *reinterpret_cast<uint4*>(sA + box * M * cfg.sw_a + swz_off_rt(cfg.sw_a, tid, cc)) = v;
    }
    fence_proxy_async_smem();
  } else if (cfg.a_mode == 3) {
    // thread = row = TMEM lane; 16 bf16 (one UMMA K step) occupy 8 TMEM columns
    for (int c = 0; c < K / 32; ++c) {
      uint32_t r[16];
      const uint4* src = reinterpret_cast<const uint4*>(Ag + (size_t)tid * K + c * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v = src[i];
        r[4 * i + 0] = v.x;
        r[4 * i + 1] = v.y;
        r[4 * i + 2] = v.z;
        r[4 * i + 3] = v.w;
      }
      tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + c * 16, r);
    }
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&bar_full, 0);
    tc_fence_after_sync();
    uint32_t idesc = make_idesc(M, N, cfg.a_mode == 1, cfg.b_mode == 1, cfg.a_f16 == 0 && cfg.c_f16 == 0, cfg.c_f16 == 0);
    if (cfg.c_f16) idesc &= ~(3u << 4);  // c_format = F16
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t bd = op_desc_rt(cfg.sw_b, cfg.b_mode, b_base, cfg.b_mode == 1 ? K : N, k, cfg.variant);
      if (cfg.a_mode == 3) {
        mma_ts(tmem_d, tmem_a + k * 8, bd, idesc, k > 0);
      } else {
        const uint64_t ad = op_desc_rt(cfg.sw_a, cfg.a_mode, a_base, cfg.a_mode == 1 ? K : M, k, cfg.variant);
        mma_ss(tmem_d, ad, bd, idesc, k > 0);
      }
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after_sync();
  if (cfg.c_f16) {
    for (int c0 = 0; c0 < N / 2; c0 += 16) {  // fp16 accumulator: column c holds elements 2c (low half) and 2c + 1
      uint32_t r[16];
      tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __half2 h2 = *reinterpret_cast<const __half2*>(&r[i]);
        Dg[(size_t)tid * N + 2 * (c0 + i)] = __low2float(h2);
        Dg[(size_t)tid * N + 2 * (c0 + i) + 1] = __high2float(h2);
      }
    }
  } else
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) Dg[(size_t)tid * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
  if (tid == 0) *status = 1;
}

// ---- MUFU micro-benchmark ------------------------------------------------------------------------
template <int MODE>
__global__ void mufu_bench_kernel(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.5f, c = a + 1.0f, d = a + 1.5f;
  uint32_t pa = 0x3c003800u + threadIdx.x, pb = 0x38003c00u + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // tanh.approx.f32
      a = tanh_approx(a); b = tanh_approx(b); c = tanh_approx(c); d = tanh_approx(d);
    } else if (MODE == 1) {  // ex2 + rcp
      asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(b)); asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(b));
    } else if (MODE == 2) {  // tanh.approx.bf16x2
      asm("tanh.approx.bf16x2 %0, %0;" : "+r"(pa)); asm("tanh.approx.bf16x2 %0, %0;" : "+r"(pb));
    } else if (MODE == 3) {  // tanh.approx.f16x2
      asm("tanh.approx.f16x2 %0, %0;" : "+r"(pa)); asm("tanh.approx.f16x2 %0, %0;" : "+r"(pb));
    } else if (MODE == 4) {  // ex2.approx.f16x2
      asm("ex2.approx.f16x2 %0, %0;" : "+r"(pa)); asm("ex2.approx.f16x2 %0, %0;" : "+r"(pb));
    } else if (MODE == 5) {  // FMA pipe reference: 4 independent FFMA chains
      a = fmaf(a, 1.0001f, 0.5f); b = fmaf(b, 1.0001f, 0.5f); c = fmaf(c, 1.0001f, 0.5f); d = fmaf(d, 1.0001f, 0.5f);
    } else if (MODE == 6) {  // pack two fp32 into bf16x2 (the result is fed back through a cheap integer op to keep a chain)
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pa) : "f"(a), "f"(b)); a = __uint_as_float(pa | 0x3f000000u);
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pb) : "f"(c), "f"(d)); c = __uint_as_float(pb | 0x3f000000u);
    } else if (MODE == 7) {  // pack two fp32 into f16x2
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pa) : "f"(a), "f"(b)); a = __uint_as_float(pa | 0x3f000000u);
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pb) : "f"(c), "f"(d)); c = __uint_as_float(pb | 0x3f000000u);
    } else if (MODE == 8) {  // saturating form
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pa) : "f"(a), "f"(b)); a = __uint_as_float(pa | 0x3f000000u);
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pb) : "f"(c), "f"(d)); c = __uint_as_float(pb | 0x3f000000u);
    } else if (MODE == 9) {  // packed half2 FMA
      asm("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(pa) : "r"(pb)); asm("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(pb) : "r"(pa));
    } else if (MODE == 10) {  // bf16x2 -> two fp32 by shifts, times a scale, -> f16x2 (the in-place operand conversion)
      const float lo = __uint_as_float(pa << 16) * 0.5f, hi = __uint_as_float(pa & 0xffff0000u) * 0.5f;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pa) : "f"(hi), "f"(lo));
      const float lo2 = __uint_as_float(pb << 16) * 0.5f, hi2 = __uint_as_float(pb & 0xffff0000u) * 0.5f;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pb) : "f"(hi2), "f"(lo2));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(pa) + __uint_as_float(pb);
}

static void rep(char* buf, size_t cap, const char* fmt, ...);

// ---- tcgen05.mma issue-rate / latency micro-benchmark --------------------------------------------
// One thread issues `count` MMAs (M = 128, K = 16) back to back, commits, and waits; cycles are read with clock64().
// mode bit 0: A from TMEM (TS) instead of shared memory; `nacc`: number of accumulators the MMAs rotate over.
__global__ void __launch_bounds__(128) umma_rate_kernel(int N, int mode, int nacc, int count, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const int amode = mode & 3, bmode = (mode >> 2) & 3;
    const uint32_t idesc = make_idesc(128, N, amode == 2, bmode >= 2, true, true);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 32768;
    const uint64_t ad = amode == 2 ? desc_mnmajor<128>(sa, 0, 16384) : amode == 3 ? desc_kmajor<64>(sa, 0) : desc_kmajor<128>(sa, 0);
    const uint64_t bd = bmode == 0 ? desc_kmajor<128>(sb, 0) : bmode == 1 ? desc_kmajor<64>(sb, 0)
                        : bmode == 2 ? desc_mnmajor<64>(sb, 0, 8192) : desc_mnmajor<128>(sb, 0, 16384);
    // latency of one MMA: issue -> commit -> barrier observed
    long long t0 = clock64();
    mma_ss(tmem, ad, bd, idesc, 0);
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    const uint32_t acc_stride = nacc > 1 ? (uint32_t)N : 0u;
    if (amode == 1) {
      for (int i = 0; i < count; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_ts(tmem + (j & 3) * acc_stride, tmem + 448, bd + (uint64_t)((j & 3) * 2), idesc, 1);
      }
    } else {
      for (int i = 0; i < count; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          mma_ss(tmem + (j & 3) * acc_stride, ad + (uint64_t)((j & 3) * 2), bd + (uint64_t)((j & 3) * 2), idesc, 1);
      }
    }
    long long t2 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 1);
    long long t3 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t1;
    out[2] = t3 - t1;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// ---- cost of tcgen05.commit in the issue stream: groups of `per_group` MMAs followed by `ncommit` commits ----
__global__ void __launch_bounds__(128) umma_commit_kernel(int per_group, int ncommit, int groups, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw3[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw3) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 32, false, false, true, true);
    const uint64_t ad = desc_kmajor<128>(smem_u32(smem), 0);
    const uint64_t bd = desc_kmajor<128>(smem_u32(smem) + 32768, 0);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      for (int j = 0; j < per_group; ++j) mma_ss(tmem, ad, bd, idesc, 1);
      for (int c = 0; c < ncommit; ++c) mma_commit(&bars[c]);
    }
    long long t1 = clock64();
    mma_commit(&bars[3]);
    mbar_wait(&bars[3], 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// ---- TMEM read / write rate: `nwarps` warps (warp w reads the lane quarter w & 3) each issue `iters` x 4 tcgen05.ld 32x32b.x32
// (4 KB per instruction) or tcgen05.st .x16, waiting once per group of four.  Reports bytes per clock of the whole SM. ----
__global__ void __launch_bounds__(512) tmem_rate_kernel(int iters, int store, long long* out) {
  __shared__ uint32_t tmem_base_s;
  __shared__ long long t0s[16], t1s[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t r[4][32];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 32; ++i) r[k][i] = tid + i;
  __syncthreads();
  const long long t0 = clock64();
  if (!store) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tmem_ld32(base + ((it * 4 + k) * 32 & 511), r[k]);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += r[k][0] ^ r[k][31];
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = r[k][i] + it;
        tmem_st16(base + ((it * 4 + k) * 16 & 511), v);
      }
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  if (lane == 0) {
    t0s[warp] = t0;
    t1s[warp] = t1;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    long long a = t0s[0], b = t1s[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      a = a < t0s[w] ? a : t0s[w];
      b = b > t1s[w] ? b : t1s[w];
    }
    out[0] = b - a;
    out[1] = acc;
  }
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

static void tmem_rate_case(int nwarps, int store, char* report, size_t cap) {
  long long* d = nullptr;
  long long h[2] = {0, 0};
  const int iters = 2048;
  cudaMalloc(&d, sizeof(h));
  tmem_rate_kernel<<<1, nwarps * 32>>>(iters, store, d);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    rep(report, cap, "tmem-rate/ CUDA-ERROR\n");
    return;
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  const double bytes = (double)nwarps * iters * 4 * (store ? 2048.0 : 4096.0);
  rep(report, cap, "tmem-rate/%s %2d warps: %7.1f B/clk per SM (%lld clk for %.0f KB)\n", store ? "tcgen05.st x16" : "tcgen05.ld x32",
      nwarps, bytes / (double)h[0], h[0], bytes / 1024);
}

static void commit_case(int per_group, int ncommit, char* report, size_t cap) {
  long long* d = nullptr;
  long long h[2] = {0, 0};
  const int groups = 256;
  cudaMalloc(&d, sizeof(h));
  cudaFuncSetAttribute(umma_commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  umma_commit_kernel<<<1, 128, 66560>>>(per_group, ncommit, groups, d);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    rep(report, cap, "mma-commit/ CUDA-ERROR\n");
    return;
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  rep(report, cap, "mma-commit/%d MMAs (N=32) + %d commits per group: issue %7.1f clk/group, complete %7.1f clk/group\n", per_group,
      ncommit, (double)h[0] / groups, (double)h[1] / groups);
}

static void rate_case(const char* name, int N, int mode, int nacc, char* report, size_t cap) {
  long long* d = nullptr;
  long long h[3] = {0, 0, 0};
  const int count = 2048;
  cudaMalloc(&d, sizeof(h));
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  umma_rate_kernel<<<1, 128, 66560>>>(N, mode, nacc, count, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    rep(report, cap, "mma-rate/%-26s CUDA-ERROR %s\n", name, cudaGetErrorString(e));
    return;
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  rep(report, cap, "mma-rate/%-26s single MMA round trip %5lld clk; %d MMAs: issue %6.1f clk/MMA, complete %6.1f clk/MMA\n", name,
      h[0], count, (double)h[1] / count, (double)h[2] / count);
}


// ---- several issuing threads sharing the tensor pipe: `nissuers` warps each issue `groups` x (8 MMAs N=32 + 1 commit), optionally
// waiting for their own commit before the next group (a dependent hand-off).  Reports aggregate clocks per MMA. ----
__global__ void __launch_bounds__(128) umma_multi_kernel(int nissuers, int groups, int wait_each, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw4[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw4) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_base_s;
  __shared__ long long t_begin[4], t_end[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp < nissuers && lane == 0) {
    const uint32_t idesc = make_idesc(128, 32, false, false, true, true);
    const uint64_t ad = desc_kmajor<128>(smem_u32(smem) + warp * 4096, 0);
    const uint64_t bd = desc_kmajor<128>(smem_u32(smem) + 32768 + warp * 4096, 0);
    const uint32_t acc = tmem + warp * 64;
    t_begin[warp] = clock64();
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int j = 0; j < 8; ++j) mma_ss(acc, ad + (uint64_t)((j & 3) * 2), bd + (uint64_t)((j & 3) * 2), idesc, 1);
      mma_commit(&bars[warp]);
      if (wait_each) mbar_wait(&bars[warp], g & 1);
    }
    mma_commit(&bars[4 + warp]);   // final drain on a fresh barrier (the per-group barrier may be many phases ahead)
    mbar_wait(&bars[4 + warp], 0);
    t_end[warp] = clock64();
  }
  __syncthreads();
  if (tid == 0) {
    long long b = t_begin[0], e = t_end[0];
    for (int i = 1; i < nissuers; ++i) {
      b = t_begin[i] < b ? t_begin[i] : b;
      e = t_end[i] > e ? t_end[i] : e;
    }
    out[0] = e - b;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


static void multi_case(int nissuers, int wait_each, char* report, size_t cap) {
  long long* d = nullptr;
  long long h[1] = {0};
  const int groups = 256;
  cudaMalloc(&d, sizeof(h));
  cudaFuncSetAttribute(umma_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  umma_multi_kernel<<<1, 128, 66560>>>(nissuers, groups, wait_each, d);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    rep(report, cap, "mma-multi/ CUDA-ERROR\n");
    return;
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  rep(report, cap, "mma-multi/%d issuing threads x (8 MMAs N=32 + commit%s): %6.1f clk per MMA aggregate, %7.1f clk per group per thread\n",
      nissuers, wait_each ? " + wait own commit" : "", (double)h[0] / (groups * 8.0 * nissuers), (double)h[0] / groups);
}

static void rep(char* buf, size_t cap, const char* fmt, ...) {
  size_t n = strlen(buf);
  if (n + 1 >= cap) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf + n, cap - n, fmt, ap);
  va_end(ap);
}

static int run_gemm_case(const char* name, StCfg cfg, char* report, size_t cap) {
  const int M = 128, N = cfg.N, K = cfg.K;
  std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);   // logical A[m][k], B[n][k]
  std::vector<float> ref((size_t)M * N, 0.f), got((size_t)M * N, 0.f);
  uint32_t s = 12345u + (uint32_t)(N * 131 + K * 7 + cfg.a_mode * 3 + cfg.b_mode);
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 16) % 7) - 3); };
  for (auto& x : hA) x = __float2bfloat16(rnd());
  for (auto& x : hB) x = __float2bfloat16(rnd());
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += __bfloat162float(hA[(size_t)m * K + k]) * __bfloat162float(hB[(size_t)n * K + k]);
      ref[(size_t)m * N + n] = acc;
    }
  // device storage: K-major operands as [rows][K]; MN-major operands as [K][rows]
  std::vector<__nv_bfloat16> dAh = hA, dBh = hB;
  if (cfg.a_mode == 1) for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) dAh[(size_t)k * M + m] = hA[(size_t)m * K + k];
  if (cfg.a_f16 || cfg.c_f16) {  // same values, fp16 bit patterns (the tensor map / copies only move 16-bit words)
    for (auto& x : dAh) {
      const __half hv = __float2half(__bfloat162float(x));
      memcpy((void*)&x, &hv, 2);
    }
  }
  if (cfg.b_mode == 1) for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) dBh[(size_t)k * N + n] = hB[(size_t)n * K + k];
  if (cfg.c_f16) {
    for (auto& x : dBh) {
      const __half hv = __float2half(__bfloat162float(x));
      memcpy((void*)&x, &hv, 2);
    }
  }
  __nv_bfloat16 *dA = nullptr, *dB = nullptr;
  float* dD = nullptr;
  int* dS = nullptr;
  cudaMalloc(&dA, dAh.size() * 2);
  cudaMalloc(&dB, dBh.size() * 2);
  cudaMalloc(&dD, got.size() * 4);
  cudaMalloc(&dS, 4);
  cudaMemcpy(dA, dAh.data(), dAh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, dBh.data(), dBh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, got.size() * 4);
  cudaMemset(dS, 0, 4);
  CUtensorMap tA, tB;
  int rc = 0;
  if (cfg.a_mode == 1) rc |= make_tmap_rows_heads(&tA, dA, K, 1, M, M, M, cfg.sw_a / 2, K);
  else rc |= make_tmap_rows_heads(&tA, dA, M, 1, K, K, K, cfg.sw_a / 2, M);
  if (cfg.b_mode == 1) rc |= make_tmap_rows_heads(&tB, dB, K, 1, N, N, N, cfg.sw_b / 2, K);
  else rc |= make_tmap_rows_heads(&tB, dB, N, 1, K, K, K, cfg.sw_b / 2, N);
  int fails = 0;
  if (rc != 0) {
    rep(report, cap, "%-34s TENSORMAP-ERROR %s\n", name, hstu_selftest_last_error());
    fails = 1;
  } else {
    const size_t smem = 1024 + 65536 + 65536;
    cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_selftest_kernel<<<1, 128, smem>>>(tA, tB, dA, dD, cfg, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      rep(report, cap, "%-34s CUDA-ERROR %s\n", name, cudaGetErrorString(e));
      return -1000;  // context is probably poisoned
    }
    cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    double maxerr = 0;
    for (size_t i = 0; i < got.size(); ++i) {
      double er = fabs((double)got[i] - (double)ref[i]);
      if (er > maxerr) maxerr = er;
      if (er != 0.0) ++bad;
    }
    rep(report, cap, "%-34s %s  mismatches=%zu/%zu maxerr=%g  (d[0]=%g ref=%g, d[last]=%g ref=%g)\n", name,
        bad == 0 ? "PASS" : "FAIL", bad, got.size(), maxerr, got[0], ref[0], got.back(), ref.back());
    fails = bad == 0 ? 0 : 1;
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  cudaFree(dS);
  return fails;
}

template <int MODE>
static void mufu_case(const char* name, int per_iter, char* report, size_t cap) {
  float* out = nullptr;
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  mufu_bench_kernel<MODE><<<blocks, threads>>>(out, 64);
  cudaEventRecord(e0);
  mufu_bench_kernel<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * iters * per_iter;
  rep(report, cap, "mufu/%-28s %8.1f G results/s (%.3f ms)\n", name, ops / (ms * 1e-3) / 1e9, ms);
  cudaFree(out);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

static int umma_selftest(char* report, size_t cap) {
  if (report == nullptr || cap < 64) return -1;
  report[0] = 0;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    rep(report, cap, "no CUDA device\n");
    return -1;
  }
  rep(report, cap, "device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  if (prop.major != 10) {
    rep(report, cap, "not an sm_100 device: tcgen05 self test skipped\n");
    return -1;
  }
  struct Case { const char* name; StCfg cfg; };
  const Case cases[] = {
      //                                   N    K   a  b  swa  swb var
      {"KK sw128 K64 N128",               {128, 64, 0, 0, 128, 128, 0}},
      {"KK sw128 K128 N128 (2 boxes)",    {128, 128, 0, 0, 128, 128, 0}},
      {"KK sw64 K32 N128",                {128, 32, 0, 0, 64, 64, 0}},
      {"KK sw32 K16 N64",                 {64, 16, 0, 0, 32, 32, 0}},
      {"KK sw128 K128 N256",              {256, 128, 0, 0, 128, 128, 0}},
      {"B MN sw128 N128 K128",            {128, 128, 0, 1, 128, 128, 0}},
      {"B MN sw128 N64 K128",             {64, 128, 0, 1, 128, 128, 0}},
      {"B MN sw64 N32 K128",              {32, 128, 0, 1, 128, 64, 0}},
      {"B MN sw128 N256 K64",             {256, 64, 0, 1, 128, 128, 0}},
      {"B MN sw128 N128 K128 [swapped]",  {128, 128, 0, 1, 128, 128, 1}},
      {"A MN sw128, B K sw128 K128 N128", {128, 128, 1, 0, 128, 128, 0}},
      {"A MN sw128, B K sw64 K128 N32",   {32, 128, 1, 0, 128, 64, 0}},
      {"A MN sw128, B MN sw64 N32",       {32, 128, 1, 1, 128, 64, 0}},
      {"A MN sw128, B MN sw128 N64",      {64, 128, 1, 1, 128, 128, 0}},
      {"A MN sw128, B MN sw128 N128",     {128, 128, 1, 1, 128, 128, 0}},
      {"A manual sw128 K128, B MN sw64",  {32, 128, 2, 1, 128, 64, 0}},
      {"A manual sw128 K128, B MN sw128", {128, 128, 2, 1, 128, 128, 0}},
      {"A manual sw64 K32, B K sw64",     {128, 32, 2, 0, 64, 64, 0}},
      {"A TMEM K128, B MN sw64 N32",      {32, 128, 3, 1, 128, 64, 0}},
      {"A TMEM K128, B MN sw128 N128",    {128, 128, 3, 1, 128, 128, 0}},
      {"A TMEM K64, B K sw128 N128",      {128, 64, 3, 0, 128, 128, 0}},
  };
  // Probe (only with HSTU_SELFTEST_MIXED=1, run it in a process of its own): A fp16 x B bf16 in ONE kind::f16 instruction.
  // The instruction descriptor has separate a_format / b_format fields, but the hardware rejects the combination with
  // "illegal instruction" (which poisons the CUDA context), so the attention kernels convert their operands to one format.
  const Case mixed_cases[] = {
      {"A f16 TMEM K128, B bf16 MN sw64",  {32, 128, 3, 1, 128, 64, 0, 1}},
      {"A f16 K sw128, B bf16 K sw128",    {128, 128, 0, 0, 128, 128, 0, 1}},
  };
  // Probe (HSTU_SELFTEST_F16ACC=1, own process): fp16 accumulators
  const Case f16acc_cases[] = {
      // fp16 accumulators (c_format = F16): N values in N / 2 TMEM columns -- half the tcgen05.ld traffic of the score tiles
      {"F16 acc: KK sw128 K128 N128",     {128, 128, 0, 0, 128, 128, 0, 0, 1}},
      {"F16 acc: KK sw64 K32 N64",        {64, 32, 0, 0, 64, 64, 0, 0, 1}},
      {"F16 acc: A TMEM K64, B MN N32",   {32, 64, 3, 1, 128, 64, 0, 0, 1}},
  };
  if (const char* env = getenv("HSTU_SELFTEST_F16ACC"); env && env[0] == '1') {
    for (const Case& c : f16acc_cases) {
      int r = run_gemm_case(c.name, c.cfg, report, cap);
      if (r == -1000) {
        rep(report, cap, "fp16 accumulator: rejected by the hardware (CUDA error above); probe ends here\n");
        return 0;
      }
    }
    return 0;
  }
  if (const char* env = getenv("HSTU_SELFTEST_MIXED"); env && env[0] == '1') {
    for (const Case& c : mixed_cases) {
      int r = run_gemm_case(c.name, c.cfg, report, cap);
      if (r == -1000) {
        rep(report, cap, "mixed fp16 x bf16 operands: rejected by the hardware (CUDA error above); probe ends here\n");
        return 0;
      }
      rep(report, cap, "mixed fp16 x bf16 operands: %s\n", r == 0 ? "ACCEPTED and exact" : "accepted but WRONG");
    }
    return 0;
  }
  int fails = 0;
  for (const Case& c : cases) {
    int r = run_gemm_case(c.name, c.cfg, report, cap);
    if (r == -1000) {
      rep(report, cap, "aborting self test after a CUDA error\n");
      return fails + 100;
    }
    if (c.cfg.variant == 0) fails += r;  // diagnostic variants do not count
  }
  // mode = amode | bmode << 2;  amode: 0 smem K-major SW128, 1 TMEM, 2 smem MN-major SW128, 3 smem K-major SW64
  //                            bmode: 0 K-major SW128, 1 K-major SW64, 2 MN-major SW64, 3 MN-major SW128
  rate_case("A K128  B K128  N=256", 256, 0 | 0 << 2, 1, report, cap);
  rate_case("A K128  B K128  N=128", 128, 0 | 0 << 2, 1, report, cap);
  rate_case("A K128  B K128  N=64", 64, 0 | 0 << 2, 1, report, cap);
  rate_case("A K128  B K128  N=32", 32, 0 | 0 << 2, 1, report, cap);
  rate_case("A TMEM  B K128  N=128", 128, 1 | 0 << 2, 1, report, cap);
  rate_case("A TMEM  B K128  N=32", 32, 1 | 0 << 2, 1, report, cap);
  rate_case("A K64   B K64   N=128", 128, 3 | 1 << 2, 1, report, cap);
  rate_case("A TMEM  B K64   N=64", 64, 1 | 1 << 2, 1, report, cap);
  rate_case("A K128  B MN64  N=32", 32, 0 | 2 << 2, 1, report, cap);
  rate_case("A TMEM  B MN64  N=32", 32, 1 | 2 << 2, 1, report, cap);
  rate_case("A K128  B MN128 N=64", 64, 0 | 3 << 2, 1, report, cap);
  rate_case("A TMEM  B MN128 N=128", 128, 1 | 3 << 2, 1, report, cap);
  rate_case("A MN128 B MN64  N=32", 32, 2 | 2 << 2, 1, report, cap);
  rate_case("A MN128 B K128  N=128", 128, 2 | 0 << 2, 1, report, cap);
  commit_case(8, 0, report, cap);
  commit_case(8, 1, report, cap);
  commit_case(8, 3, report, cap);
  commit_case(4, 1, report, cap);
  commit_case(2, 2, report, cap);
  for (int w = 0; w < 2; ++w)
    for (int n = 1; n <= 3; ++n) multi_case(n, w, report, cap);

  for (int nw : {1, 4, 8, 12, 16}) tmem_rate_case(nw, 0, report, cap);
  for (int nw : {4, 8}) tmem_rate_case(nw, 1, report, cap);
  mufu_case<0>("tanh.approx.f32", 4, report, cap);
  mufu_case<1>("ex2+rcp f32 (sigmoid)", 2, report, cap);
  mufu_case<2>("tanh.approx.bf16x2", 4, report, cap);
  mufu_case<3>("tanh.approx.f16x2", 4, report, cap);
  mufu_case<4>("ex2.approx.f16x2", 4, report, cap);
  mufu_case<5>("ffma f32 (reference)", 4, report, cap);
  // packing conversions: which pipe / rate?  (per_iter counts PACK INSTRUCTIONS, two fp32 -> one 32-bit register each)
  mufu_case<6>("cvt.rn.bf16x2.f32 (packs)", 2, report, cap);
  mufu_case<7>("cvt.rn.f16x2.f32 (packs)", 2, report, cap);
  mufu_case<8>("cvt.rn.satfinite.f16x2.f32", 2, report, cap);
  mufu_case<9>("fma.rn.f16x2 (instr)", 2, report, cap);
  mufu_case<10>("bf16x2 -> f16x2 via fp32", 2, report, cap);
  rep(report, cap, "failed checks: %d\n", fails);
  return fails;
}

}  // namespace hstu

// ---- the test library's own C ABI and the one internal symbol tmap.cu needs ----
namespace hstu {
static thread_local char g_selftest_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_selftest_err, sizeof(g_selftest_err), fmt, ap);
  va_end(ap);
}
}  // namespace hstu

extern "C" {
const char* hstu_selftest_last_error(void) { return hstu::g_selftest_err; }
int hstu_umma_selftest(char* report, size_t report_bytes) { return hstu::umma_selftest(report, report_bytes); }
}
