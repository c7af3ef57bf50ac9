// Jagged HSTU attention backward on tcgen05 + TMEM with TMA-staged tiles.  bf16 / fp16, dqk == dv in {32, 64, 128}.
//
// One CTA per (128-row KEY tile, head, sequence); early key tiles (the heavy ones under a causal mask) are scheduled
// first.  K and V of the tile stay in shared memory; the CTA streams the query tiles that can attend to it
// (Q_i and dO_i, TMA ring of 4 / 3 / 1 stages for d = 32 / 64 / 128).  Work is pipelined in half-tiles ("units":
// 128 key rows x 64 query rows); three warps issue MMAs, each with its own wait -> issue -> commit loop:
//   X:  S^T  = K Q^T          (A = K  K-major,  B = 64 rows of Q_i  K-major)   -> TMEM slot, columns [0, 64)
//       dP^T = V dO^T         (A = V  K-major,  B = 64 rows of dO_i K-major)   -> TMEM slot, columns [64, 128)
//   The slots form a ring of 3 / 2 / 1.  The two elementwise warpgroups (warpgroup h owns half h of every query tile; one
//   key row per thread) turn S^T, dP^T into
//        P^T  = silu(alpha S)            * mask      (1/N folded into the dV epilogue)
//        dS^T = dP sig (1 + x (1 - sig)) * mask      (alpha/N folded into the dK epilogue / dQ convert)
//   from ONE tanh per score (packed fp32x2 arithmetic).  Both are ALWAYS fp16 tensor-core operands, whatever the input dtype
//   (kind::f16 takes the A and B formats independently): an 11-bit significand keeps their rounding inside the 1e-3 parity
//   budget, a bf16 P / dS does not (measured 1.7e-3).  P is bounded by |alpha s|; dS inherits the scale of dO, so it is stored
//   as dS * 2^-e with e = floor(log2 max|dO|) (dout_amax_kernel) and the power of two is undone in the epilogues.
//   P^T overwrites the front of the slot in TMEM, dS^T goes to a box in shared memory ([kv][q], q contiguous, 128B swizzle;
//   two box pairs).
//   Y:  dV  += P^T  dO        (A = P^T from the TMEM slot,  B = 64 rows of dO_i MN-major)   [kv x d]  all query tiles
//   Z:  dK  += dS^T Q         (A = dS^T box K-major,  B = 64 rows of Q_i  MN-major)         [kv x d]
//       dQ_i = dS   K         (A = both dS^T boxes of the tile read MN-major, B = K MN-major)   [128 q x d]
//   dQ_i (two TMEM accumulators, drained two tiles late) is staged as fp32 in a swizzled shared-memory box and added to an
//   fp32 accumulator in global memory with one TMA reduce-add per warpgroup (each key-tile CTA contributes to every later
//   query tile); a small convert kernel scales it by alpha/N and writes bf16 dq.
//
// Reference math: ops/triton/triton_hstu_attention.py:995-1006,1222 and SURVEY.md appendix A; unlike the Triton
// kernel dQ is accumulated in fp32, not in the input dtype (triton_attention_utils.py:47-60).
#include <string.h>

#include "common.cuh"
#include "internal.h"
#include "umma.cuh"

namespace hstu {
using namespace umma;

bool umma_fwd_supported(const hstu_attn_params& p);

struct alignas(64) BwdParams {
  CUtensorMap tmQ, tmK, tmV, tmDO, tmDQ;  // tmDQ: the fp32 dQ accumulator (TMA reduce-add destination)
  const void* seq_offsets;
  const void* num_targets;
  void* dk;
  void* dv;
  float* dq_acc;  // [L, H, D] fp32, zero-initialised
  const uint32_t* dout_amax_bits;  // fp32 bit pattern of max |dO| over the whole tensor (written by dout_amax_kernel)
  long long dk_row_stride, dk_head_stride, dv_row_stride, dv_head_stride;
  int offsets_i64, targets_i64;
  int max_seq_len, heads;
  int win, min_full, ctx;
  float alpha_half;
  float dv_scale;  // 1 / N
  float dk_scale;  // alpha / N
};

#ifdef HSTU_BWD_PSMEM
constexpr bool kBwdPsmem = true;
#else
constexpr bool kBwdPsmem = false;
#endif

template <int D>
struct BwdCfg {
  static constexpr int SW = (D * 2 >= 128) ? 128 : D * 2;
  static constexpr int BOX_COLS = SW / 2;
  static constexpr int NBOX = D / BOX_COLS;
  static constexpr int BOX_BYTES = 128 * SW;
  static constexpr int TILE_BYTES = 128 * D * 2;
  static constexpr int PT_BYTES = 128 * 128 * 2;          // one pair buffer: two [128 kv][64 q] boxes
  static constexpr int STAGES = (D <= 32) ? 4 : (D == 64 ? 3 : 1);  // Q_i / dO_i TMA ring depth (what shared memory allows)
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE_BYTES;
  static constexpr int OFF_Q = OFF_V + TILE_BYTES;
  static constexpr int OFF_DO = OFF_Q + STAGES * TILE_BYTES;
  // dS^T boxes [kv][q] (tile i -> pair buffer i & 1): read K-major as A of dK (M = kv) and MN-major as A of dQ (M = q)
  static constexpr int OFF_DST = OFF_DO + STAGES * TILE_BYTES;
  // dQ staging (source of the TMA reduce-add): two alternating boxes of 128 query rows x 32 fp32 columns (128B swizzle); the
  // drain warpgroup sends the D columns of a dQ tile in DQ_NPASS boxes
  static constexpr int DQ_BOX_COLS = 32;
  static constexpr int DQ_NPASS = D / DQ_BOX_COLS;
  static constexpr int DQS_BYTES = 128 * DQ_BOX_COLS * 4;
  static constexpr int OFF_DQS = OFF_DST + 2 * PT_BYTES;
  static constexpr int OFF_PT = OFF_DQS + 2 * DQS_BYTES;   // PSM: NPB P^T boxes [128 kv][64 q] fp16
  static constexpr int OFF_BAR = OFF_PT + ((kBwdPsmem && D == 32) ? 3 * 16384 : 0);
  static constexpr int BAR_BYTES = 512;   // sizeof(BwdBars) <= 512 is asserted next to the struct (it had outgrown the 256 bytes reserved)
  static_assert(OFF_BAR + BAR_BYTES + 1024 <= 232448, "shared memory budget");
  static constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + 1024;
  // TMEM: a ring of NSLOT score slots, each {S^T half-tile: 64 columns, dP^T half-tile: 64 columns} (a half-tile is
  // 128 key rows x 64 query rows; after the elementwise stage the front of the S^T half holds P^T as bf16), the dV and
  // dK accumulators and NDQ dQ accumulators (tile i -> buffer i % NDQ; with two, the warpgroups drain dQ two tiles late and
  // never wait for it).  d = 128 fills TMEM with one slot and one dQ accumulator: 128 + 3 * 128 columns.
  //
  // PRING mode, r02.  In r01 the slot of a unit also carried its P^T (written over the scores), so the issuer of the
  // scores had to wait for the dV GEMM of the unit three back: a dependency cycle X -> elementwise -> YV -> X of ~3700 clk per
  // three units that set the pace of the CTA (profiles/r02_bwd_timelines.txt (A)).  Now P^T goes to its own small ring of NPR
  // buffers (32 columns each) and a warpgroup loads ALL its scores (64 + 64 registers) before it starts the arithmetic: each
  // warpgroup owns one slot (unit u -> slot u & 1) and hands it back (scores_free) right after the load, so the scores of its
  // next unit are computed while it works.  The two warpgroups stay half a unit apart by themselves, which keeps the MUFU / pack
  // pipe busy through each other's load / store phases.  (A whole-tile variant with one 256-column slot and N = 128 score MMAs,
  // (C) / (D) in the same file, has 28 % less tensor-pipe work but forces both warpgroups into lockstep: 2.8 ms instead of 2.25.)
  // d = 128 keeps the r01 layout: its dK / dV accumulators leave room for one slot only and none for a P^T ring.
  // Measured on B200 (bwd ms, bf16; profiles/r02_bwd_variants.txt): d = 32 (B 16, H 8, Lmax 8192): ring of 3 slots 2.36, PRING 2.77;
  // d = 64 (B 512, H 4, Lmax 2048): ring of 2 slots 5.55, PRING 5.04.  So PRING is used where it wins: d = 64.
#ifndef HSTU_BWD_PRING_MASK
#define HSTU_BWD_PRING_MASK 64   /* bit mask over head dims: 32 | 64 */
#endif
  static constexpr bool PRING = (HSTU_BWD_PRING_MASK & D) != 0 && D <= 64;
  // PSM (d = 32): THREE elementwise warpgroups, P^T through shared memory.  ncu / timelines of the 2-warpgroup kernel: no pipe above
  // 45 %, issue slots 55 % busy, two elementwise warps per scheduler -- the stage is bound by the latency of its own dependent
  // chains, and a score slot is held from the score GEMMs until the dV GEMM has consumed the P^T written over it (cycle
  // X -> elementwise -> YV -> X of ~3800 clk per three units).  Here unit u is processed by warpgroup u % 3 out of slot u % 3;
  // P^T goes to a ring of NPB boxes in shared memory ([kv][q] fp16, the layout of the dS^T boxes; dV becomes an SS GEMM), so the
  // slot returns to the score issuer as soon as its values are in registers (scores_free), and a third warp per scheduler
  // fills the latency gaps of the other two.  640 threads; register budgets 48 / 120 / 72 (pool 640 x 96).
  // (Probed and rejected: fp16 accumulators for the score GEMMs are NOT packed in TMEM -- one value per 32-bit column -- so they save
  // neither columns nor tcgen05.ld traffic; and the TMEM read port is not the limit: 840 B / clk / SM measured, ~60 needed.)
  static constexpr bool PSM = kBwdPsmem && D == 32 && !PRING;
  static constexpr int NEW = PSM ? 3 : 2;              // elementwise warpgroups
  static constexpr int THREADS = 128 * (2 + NEW);      // issuers + elementwise + drain
  static constexpr int NPB = 3;                        // PSM: P^T boxes, unit u -> u % NPB
  static constexpr int SLOT_COLS = 128;                // {S^T half-tile | dP^T half-tile}
  static constexpr int HALF_COLS = SLOT_COLS / 2;
  static constexpr int NSLOT = PRING ? 2 : (D <= 32 ? 3 : (D == 64 ? 2 : 1));
  static constexpr int NPR = (D <= 32) ? 4 : 2;  // PRING: P^T buffers (unit u -> u % NPR)
  static constexpr int NDQ = (D <= 32) ? 2 : ((D == 64 && !PRING) ? 2 : 1);
  // s_full barrier instances: unit u -> [u % NSF].  At least two even with one slot: the warpgroups alternate units, and a
  // warpgroup must never wait for phase k+1 of a barrier before phase k has completed (the parity test would pass at once).
  static constexpr int NSF = NSLOT < 2 ? 2 : NSLOT;
  static constexpr int LAG = NDQ;                // the warpgroups drain dQ of tile i - LAG after their unit of tile i
  static constexpr int TMEM_SLOT = 0;
  static constexpr int TMEM_P = NSLOT * SLOT_COLS;     // PRING: NPR buffers of 32 columns
  static constexpr int TMEM_DV = NSLOT * SLOT_COLS + (PRING ? NPR * 32 : 0);
  static constexpr int TMEM_DK = TMEM_DV + D;
  static constexpr int TMEM_DQ = TMEM_DK + D;   // NDQ buffers of D columns
  static_assert(TMEM_DQ + NDQ * D <= 512, "TMEM budget");
  // scripts/sim_bwd_protocol.py: a 3-slot score ring needs the Q/dO ring to be at least 4 deep (the scores of tile i+2 are
  // requested before tile i releases its stage), otherwise the producer and the MMA issuer wait on each other.
  static_assert(NSLOT != 3 || STAGES >= 4, "3-slot score ring needs >= 4 Q/dO stages");
  static_assert(!PRING || STAGES >= 3, "PRING: tile i + 1 is staged while tile i is in flight and tile i - 1 drains");
};

struct BwdBars {
  uint64_t kv_full;
  uint64_t q_full[4];
  uint64_t kv_ready, q_ready[4];  // bf16 inputs: the tiles have been converted to fp16 in place (128 threads of the drain warpgroup)
  uint64_t s_full[3], unit_done[4];
  uint64_t tile_done[4];  // tile i -> [i % 4]: both issuers have finished every GEMM of query tile i (count 2)
  uint64_t slot_free[3];  // unit u -> [u % NSLOT]: dV of the unit has consumed P^T in the slot
  uint64_t dq_empty[2], fin_full;
  // PRING: scores_free[h] (128 arrivals: warpgroup h has loaded the scores of its unit), p_free[u % NPR] (YV: dV of the unit has
  // consumed the P^T buffer)
  uint64_t scores_free[3], p_free[4];
  uint32_t tmem_base;
};
static_assert(sizeof(BwdBars) <= 512, "BwdBars must fit the bytes reserved for it (BwdCfg::BAR_BYTES)");

#ifdef HSTU_TRACE
// Debug timeline: CTA (0,0,0) records clock64() stamps of its pipeline events into g_trace[role][index][slot].
__device__ long long* g_trace = nullptr;
#define HSTU_TSTAMP(role, idx, k)                                                                              \
  do {                                                                                                           \
    if (g_trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < 256)                \
      g_trace[((role) * 256 + (idx)) * 4 + (k)] = clock64();                                                     \
  } while (0)
#else
#define HSTU_TSTAMP(role, idx, k) do { } while (0)
#endif

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// 2^-e with e = floor(log2(amax)) (amax given as fp32 bits): amax * scale lies in [1, 2).  Zero / denormal / non-finite amax
// are clamped to the representable exponent range, so the result is always a finite, non-zero power of two.
__host__ __device__ __forceinline__ float ds_scale_from_amax(uint32_t amax_bits) {
  uint32_t e = (amax_bits >> 23) & 0xffu;
  e = e < 1u ? 127u : (e > 253u ? 253u : e);  // amax == 0 -> scale 1
  const uint32_t bits = (254u - e) << 23;
#ifdef __CUDA_ARCH__
  return __uint_as_float(bits);
#else
  float f;
  memcpy(&f, &bits, 4);
  return f;
#endif
}

// max |dO| over the [rows, heads, D] tensor -> atomicMax on the fp32 bit pattern (order-preserving for non-negative floats)
template <bool BF16>
__global__ void dout_amax_kernel(const uint16_t* __restrict__ dout, long long rows, int heads, int D, long long row_stride,
                                 long long head_stride, uint32_t* __restrict__ amax_bits) {
  const long long nvec = rows * heads * (D / 8);
  uint32_t m = 0;  // max of the 15-bit magnitudes (monotone in |x| for both 16-bit formats)
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nvec; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % (D / 8));
    const long long rh = idx / (D / 8);
    const int hh = (int)(rh % heads);
    const long long r = rh / heads;
    const uint4 x = *reinterpret_cast<const uint4*>(dout + r * row_stride + hh * head_stride + v * 8);
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m = max(m, w[i] & 0x7fffu);
      m = max(m, (w[i] >> 16) & 0x7fffu);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0) {
    const uint32_t fbits = BF16 ? (m << 16) : __float_as_uint(__half2float(__ushort_as_half((unsigned short)m)));
    atomicMax(amax_bits, fbits);
  }
}

__device__ __forceinline__ void bulk_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// 512 threads = 4 warpgroups with their own register budgets (setmaxnreg; 128 regs / thread at launch -- a CTA can only
// redistribute the pool it was launched with: requests beyond it spin forever in setmaxnreg.inc):
//   warps 0-3    issuers X (scores), YV (dV), YK (dK), Z (dQ): one elected lane each                          -> 64 regs
//   warps 4-11   two elementwise warpgroups (one per half of the query tile).  PRING mode: a thread loads its 64 scores and 64
//                dP values at once (128 registers), hands the score slot back and only then starts the arithmetic  -> 176 regs
//   warps 12-15  dQ drain warpgroup; its elected lane is also the TMA producer; converts bf16 tiles to fp16      -> 96 regs
// (r02 also tried FOUR elementwise warpgroups at 768 threads / 80 registers: better MUFU utilisation per unit, 76 % instead of
// 60 %, but with one score slot the four groups run in lockstep and nothing overlaps the score GEMMs: 2.6 ms instead of 2.25.)
template <int D, bool BF16>
__global__ void __launch_bounds__(BwdCfg<D>::THREADS, 1) attn_bwd_umma_kernel(const __grid_constant__ BwdParams p) {
  using Cfg = BwdCfg<D>;
  constexpr int SW = Cfg::SW;
  constexpr int NST = Cfg::STAGES;
  // bf16 inputs: every TMA-landed tile (K, V, Q_i, dO_i) is converted to fp16 in place in shared memory (dO times the power of
  // two 2^-e, e = floor(log2 max|dO|), so that dP, dS, dK, dQ and dV all carry that factor until their epilogues) and ALL MMAs of
  // the kernel run fp16 x fp16.  fp16 inputs skip the conversion; there the factor enters through the dS constants instead.
  constexpr bool CONV = BF16;
  const int b = blockIdx.z, h = blockIdx.y;
  const int n0 = (int)blockIdx.x * 128;
  const long long row0 = load_index(p.seq_offsets, p.offsets_i64, b);
  int len = (int)(load_index(p.seq_offsets, p.offsets_i64, b + 1) - row0);
  if (len > p.max_seq_len) {  // rows past max_seq_len: zero gradients (dq: the accumulator stays zero there)
    if (blockIdx.x == 0) {
      zero_rows(p.dk, 2, p.dk_row_stride, (long long)h * p.dk_head_stride, D, row0 + p.max_seq_len, row0 + len);
      zero_rows(p.dv, 2, p.dv_row_stride, (long long)h * p.dv_head_stride, D, row0 + p.max_seq_len, row0 + len);
    }
    len = p.max_seq_len;
  }
  if (n0 >= len) return;
  const int n_tgt = p.num_targets ? (int)load_index(p.num_targets, p.targets_i64, b) : -1;
  const SeqMask msk = make_seq_mask(len, n_tgt, p.win, p.min_full, p.ctx);
  const int nrows = min(128, len - n0);
  int lo, hi, ctx_hi;
  q_range_for_kv_rows(msk, n0, n0 + nrows, &lo, &hi, &ctx_hi);
  // query tiles: the contextual prefix tiles that lie strictly before the main range, then the main range
  const int main_t0 = lo / 128;
  const int n_main = (hi + 127) / 128 - main_t0;
  const int n_pre = min((ctx_hi + 127) / 128, main_t0);
  const int T = n_pre + n_main;
  auto q_tile = [&](int i) { return i < n_pre ? i : main_t0 + (i - n_pre); };

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sDO = smem + Cfg::OFF_DO;
  uint8_t* sDST = smem + Cfg::OFF_DST;
  uint8_t* sDQS = smem + Cfg::OFF_DQS;
  uint8_t* sPT = smem + Cfg::OFF_PT;   // PSM only
  BwdBars* bars = reinterpret_cast<BwdBars*>(smem + Cfg::OFF_BAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bars->kv_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->tile_done[i], 3);   // the commits of YV, YK and Z
    }
    for (int i = 0; i < 3; ++i) mbar_init(&bars->s_full[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bars->unit_done[i], 128);
    for (int i = 0; i < 2; ++i) mbar_init(&bars->dq_empty[i], 128);
    mbar_init(&bars->fin_full, 2);         // YV (dV) and YK (dK)
    mbar_init(&bars->kv_ready, 128);
    // Q_i / dO_i tiles are converted by the drain warpgroup (128 arrivals), or -- with a single stage (d = 128), where the
    // conversion is on the critical path of every tile -- by the two elementwise warpgroups together (256 arrivals)
    for (int i = 0; i < 4; ++i) mbar_init(&bars->q_ready[i], NST == 1 ? 256 : 128);
    for (int i = 0; i < 3; ++i) mbar_init(&bars->slot_free[i], 1);
    for (int i = 0; i < 3; ++i) mbar_init(&bars->scores_free[i], 128);
    for (int i = 0; i < 4; ++i) mbar_init(&bars->p_free[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;
  uint64_t* const kv_rdy = CONV ? &bars->kv_ready : &bars->kv_full;
  uint64_t* const q_rdy = CONV ? bars->q_ready : bars->q_full;
  const float ds_scale = ds_scale_from_amax(__ldg(p.dout_amax_bits));  // 2^-e

  if (warp < 4) {
    if constexpr (Cfg::PSM) reg_dealloc<48>(); else reg_dealloc<64>();
    // ---------------- MMA issuers ----------------
    // Work is pipelined in "units" u = 2 i + h: half h (64 query rows) of query tile i.  Unit u's scores live in TMEM
    // slot u % NSLOT; warpgroup h turns them into P^T (fp16, over the front of the slot) and the dS^T box (pair i & 1,
    // box h) in shared memory.  FOUR threads issue, each with its own wait -> issue -> commit loop: a tcgen05.mma blocks its
    // issuing thread until the tensor pipe accepts it and a commit costs the thread ~100-200 clk more, so the thread with the
    // most MMAs per unit sets the pace of the CTA (r01: one thread with dV + dK = 8 MMAs + 2 commits per unit was busy 100 % of
    // the time and the whole kernel ran at its rate; profiles/r01_bwd_timeline.txt).
    //   X (warp 0): S^T, dP^T.   YV (warp 1): dV.   YK (warp 2): dK.   Z (warp 3): dQ of every tile.
    // The whole warp runs the warp-uniform control flow, one fixed lane issues; descriptors are built once.
    constexpr int NSLOT = Cfg::NSLOT;
    const bool leader = lane == 0;
    const int U = 2 * T;
    if (warp == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, 64, false, false, false, false);  // S^T, dP^T half-tiles (fp16 x fp16)
      const uint64_t dk_k = desc_kmajor<SW>(smem_u32(sK), 0);                        // K as K-major A (S^T)
      const uint64_t dv_k = desc_kmajor<SW>(smem_u32(sV), 0);                        // V as K-major A (dP^T)
      const uint64_t dq_k = desc_kmajor<SW>(smem_u32(sQ), 0);                        // Q_i rows as K-major B
      const uint64_t ddo_k = desc_kmajor<SW>(smem_u32(sDO), 0);                      // dO_i rows as K-major B
      mbar_wait(kv_rdy, 0);
      tc_fence_after_sync();
      for (int u = 0; u < U; ++u) {
        const int i = u >> 1, hf = u & 1, st = i % NST, slot = u % NSLOT;
        if (leader) HSTU_TSTAMP(0, u, 0);
        if (Cfg::PRING) {
          if (i >= 1) {  // warpgroup hf has loaded the scores of its previous unit out of this slot
            mbar_wait(&bars->scores_free[hf], (i - 1) & 1);
            tc_fence_after_sync();
          }
        } else if (Cfg::PSM) {
          if (u >= NSLOT) {  // warpgroup (u - NSLOT) % 3 has loaded the scores of unit u - NSLOT out of this slot
            mbar_wait(&bars->scores_free[slot], (u / NSLOT - 1) & 1);
            tc_fence_after_sync();
          }
        } else if (u >= NSLOT) {  // the slot still holds P^T of unit u - NSLOT until its dV GEMM has completed
          mbar_wait(&bars->slot_free[slot], (u / NSLOT - 1) & 1);
          tc_fence_after_sync();
        }
        if (hf == 0) {
          mbar_wait(&q_rdy[st], (i / NST) & 1);
          tc_fence_after_sync();
        }
        // query rows [64 hf, 64 hf + 64) of the staged Q_i / dO_i tiles
        const uint64_t row_off = (uint64_t)((st * Cfg::TILE_BYTES + hf * 64 * SW) >> 4);
        const uint32_t ts = tmem + Cfg::TMEM_SLOT + slot * Cfg::SLOT_COLS;
        if (leader) {
          HSTU_TSTAMP(0, u, 1);
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t kb = ks * 32, bx = kb / SW, off = kb % SW;
            const uint64_t o = (uint64_t)((bx * Cfg::BOX_BYTES + off) >> 4);
            mma_ss(ts, dk_k + o, dq_k + row_off + o, idesc_s, ks > 0);
          }
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t kb = ks * 32, bx = kb / SW, off = kb % SW;
            const uint64_t o = (uint64_t)((bx * Cfg::BOX_BYTES + off) >> 4);
            mma_ss(ts + Cfg::HALF_COLS, dv_k + o, ddo_k + row_off + o, idesc_s, ks > 0);
          }
          mma_commit(&bars->s_full[u % Cfg::NSF]);
          HSTU_TSTAMP(0, u, 2);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---- issuer YV: dV += P^T dO (A = P^T from the unit's TMEM slot) of every unit; its commit frees the slot ----
      constexpr uint32_t idesc_kv = make_idesc(128, D, false, true, false, false);   // A = P^T from TMEM, B MN-major (fp16 x fp16)
      const uint64_t ddo_mn = desc_mnmajor<SW>(smem_u32(sDO), 0, Cfg::BOX_BYTES);    // dO_i rows as MN-major B
      for (int u = 0; u < U; ++u) {
        const int i = u >> 1, hf = u & 1, st = i % NST, pb = i & 1, slot = u % NSLOT;
        // P^T / dS^T of the unit are written.  One barrier per (half, tile parity): with a 3-slot score ring a warpgroup may
        // finish TWO units before this thread gets here; a single barrier per half would then be two phases ahead and the
        // parity wait would alias.
        if (leader) HSTU_TSTAMP(1, u, 0);
        mbar_wait(&bars->unit_done[hf * 2 + pb], (i >> 1) & 1);
        tc_fence_after_sync();
        const uint64_t rows = (uint64_t)((st * Cfg::TILE_BYTES + hf * 64 * SW) >> 4);  // MN-major B: K rows = the 64 query rows
        const uint32_t tp = Cfg::PRING ? tmem + Cfg::TMEM_P + (u % Cfg::NPR) * 32 : tmem + Cfg::TMEM_SLOT + slot * Cfg::SLOT_COLS;
        if (leader) {
          HSTU_TSTAMP(1, u, 1);
          if constexpr (Cfg::PSM) {
            // A = the unit's P^T box in shared memory (K-major, like the dS^T box of the dK GEMM)
            const uint64_t dpt_k = desc_kmajor<128>(smem_u32(sPT), 0) + (uint64_t)(((u % Cfg::NPB) * 16384) >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ss(tmem + Cfg::TMEM_DV, dpt_k + (uint64_t)((ks * 32) >> 4), ddo_mn + rows + (uint64_t)((ks * 16 * SW) >> 4), idesc_kv,
                     (u > 0) || (ks > 0));
            mma_commit(&bars->p_free[u % Cfg::NPB]);   // the P^T box may be rewritten
          } else {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)  // K = the 64 query rows of this half
            mma_ts(tmem + Cfg::TMEM_DV, tp + ks * 8, ddo_mn + rows + (uint64_t)((ks * 16 * SW) >> 4), idesc_kv, (u > 0) || (ks > 0));
          // PRING: the P^T buffer may be rewritten; otherwise: the score issuer may overwrite the slot
          mma_commit(Cfg::PRING ? &bars->p_free[u % Cfg::NPR] : &bars->slot_free[slot]);
          }
          if (hf == 1) mma_commit(&bars->tile_done[i & 3]);  // this issuer is done with dO_i
          HSTU_TSTAMP(1, u, 2);
        }
        __syncwarp();
      }
      if (leader) mma_commit(&bars->fin_full);
      __syncwarp();
    } else if (warp == 2) {
      // ---- issuer YK: dK += dS^T Q (A = the unit's dS^T box in shared memory) of every unit ----
      constexpr uint32_t idesc_kv = make_idesc(128, D, false, true, false, false);   // A = dS^T K-major, B MN-major (fp16 x fp16)
      const uint64_t dds_k = desc_kmajor<128>(smem_u32(sDST), 0);                    // dS^T box [kv][q] as K-major A
      const uint64_t dq_mn = desc_mnmajor<SW>(smem_u32(sQ), 0, Cfg::BOX_BYTES);      // Q_i rows as MN-major B
      for (int u = 0; u < U; ++u) {
        const int i = u >> 1, hf = u & 1, st = i % NST, pb = i & 1;
        if (leader) HSTU_TSTAMP(4, u, 0);
        mbar_wait(&bars->unit_done[hf * 2 + pb], (i >> 1) & 1);
        tc_fence_after_sync();
        const uint64_t box = (uint64_t)((pb * Cfg::PT_BYTES + hf * 16384) >> 4);
        const uint64_t rows = (uint64_t)((st * Cfg::TILE_BYTES + hf * 64 * SW) >> 4);
        if (leader) {
          HSTU_TSTAMP(4, u, 1);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_ss(tmem + Cfg::TMEM_DK, dds_k + box + (uint64_t)((ks * 32) >> 4), dq_mn + rows + (uint64_t)((ks * 16 * SW) >> 4),
                   idesc_kv, (u > 0) || (ks > 0));
          if (hf == 1) mma_commit(&bars->tile_done[i & 3]);  // this issuer is done with Q_i and (as A of dK) the dS^T boxes of tile i
          HSTU_TSTAMP(4, u, 2);
        }
        __syncwarp();
      }
      if (leader) mma_commit(&bars->fin_full);
      __syncwarp();
    } else {
      // ---- issuer Z: dQ_i = dS_i K (A = the dS^T box pair read MN-major, M = 128 query rows) of every query tile ----
      constexpr uint32_t idesc_dq = make_idesc(128, D, true, true, false, false);    // A = dS^T MN-major, B MN-major (fp16 x fp16)
      const uint64_t dds_mn = desc_mnmajor<128>(smem_u32(sDST), 0, 16384);           // the box pair as MN-major A
      const uint64_t dk_mn = desc_mnmajor<SW>(smem_u32(sK), 0, Cfg::BOX_BYTES);      // K as MN-major B
      mbar_wait(kv_rdy, 0);
      for (int i = 0; i < T; ++i) {
        const int pb = i & 1;
        if (leader) HSTU_TSTAMP(5, i, 0);
        mbar_wait(&bars->unit_done[0 * 2 + pb], (i >> 1) & 1);
        mbar_wait(&bars->unit_done[1 * 2 + pb], (i >> 1) & 1);
        if (leader) HSTU_TSTAMP(5, i, 1);
        if (i >= Cfg::NDQ) mbar_wait(&bars->dq_empty[i % Cfg::NDQ], ((i / Cfg::NDQ) - 1) & 1);  // dQ_{i-NDQ} has been drained from this accumulator
        tc_fence_after_sync();
        if (leader) {
          HSTU_TSTAMP(5, i, 2);
          const uint64_t pair = (uint64_t)((pb * Cfg::PT_BYTES) >> 4);  // the dS^T boxes of this query tile
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // K = 128 key rows, 16 per step
            mma_ss(tmem + Cfg::TMEM_DQ + (i % Cfg::NDQ) * D, dds_mn + pair + (uint64_t)((ks * 16 * 128) >> 4),
                   dk_mn + (uint64_t)((ks * 16 * SW) >> 4), idesc_dq, ks > 0);
          mma_commit(&bars->tile_done[i & 3]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 + 4 * Cfg::NEW) {
    if constexpr (Cfg::PSM) reg_dealloc<72>(); else reg_dealloc<96>();
    // ---------------- dQ drain warpgroup (+ TMA producer on its elected lane) ----------------
    // dQ tile of query tile i: TMEM (lane = query row) -> swizzled fp32 staging box (32 columns) in shared memory -> ONE TMA
    // reduce-add per box into dq_acc.  Two staging boxes alternate, so a reduce may still be reading one while the next is being
    // filled.  (r01 had the elementwise warpgroups drain dQ themselves: ~750 clk per tile taken from the stage that is the
    // slowest of the kernel.)  The producer duty follows the same barrier: once tile i is done, its Q / dO stage is free for
    // tile i + NST.
    const int quad = warp & 3;
    const int row = quad * 32 + lane;              // query row inside the tile == TMEM lane
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const bool elected = warp == 4 + 4 * Cfg::NEW && lane == 0;
    auto load_tile = [&](int i) {
      const int st = i % NST;
      mbar_arrive_expect_tx(&bars->q_full[st], 2 * Cfg::TILE_BYTES);
      const int qrow = (int)(row0 + (long long)q_tile(i) * 128);
#pragma unroll
      for (int bx = 0; bx < Cfg::NBOX; ++bx) {
        tma_load_3d(sQ + st * Cfg::TILE_BYTES + bx * Cfg::BOX_BYTES, &p.tmQ, &bars->q_full[st], bx * Cfg::BOX_COLS, h, qrow);
        tma_load_3d(sDO + st * Cfg::TILE_BYTES + bx * Cfg::BOX_BYTES, &p.tmDO, &bars->q_full[st], bx * Cfg::BOX_COLS, h, qrow);
      }
    };
    if (elected) {
      prefetch_tensormap(&p.tmK);
      prefetch_tensormap(&p.tmV);
      prefetch_tensormap(&p.tmQ);
      prefetch_tensormap(&p.tmDO);
      prefetch_tensormap(&p.tmDQ);
      mbar_arrive_expect_tx(&bars->kv_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
      for (int bx = 0; bx < Cfg::NBOX; ++bx) {
        tma_load_3d(sK + bx * Cfg::BOX_BYTES, &p.tmK, &bars->kv_full, bx * Cfg::BOX_COLS, h, (int)(row0 + n0));
        tma_load_3d(sV + bx * Cfg::BOX_BYTES, &p.tmV, &bars->kv_full, bx * Cfg::BOX_COLS, h, (int)(row0 + n0));
      }
      for (int i = 0; i < NST && i < T; ++i) load_tile(i);
    }
    const int ct = tid - 128 * (1 + Cfg::NEW);  // index inside this warpgroup
    auto convert_tile = [&](int i) {  // Q_i (as is) and dO_i (times 2^-e) of stage i % NST: bf16 -> fp16 in place
      const int st = i % NST;
      mbar_wait(&bars->q_full[st], (i / NST) & 1);
      convert_bf16_to_f16_inplace<128>(sQ + st * Cfg::TILE_BYTES, Cfg::TILE_BYTES, ct, 1.0f);
      convert_bf16_to_f16_inplace<128>(sDO + st * Cfg::TILE_BYTES, Cfg::TILE_BYTES, ct, ds_scale);
      fence_proxy_async_smem();
      mbar_arrive(&bars->q_ready[st]);
    };
    if (CONV) {
      mbar_wait(&bars->kv_full, 0);
      convert_bf16_to_f16_inplace<128>(sK, Cfg::TILE_BYTES, ct, 1.0f);
      convert_bf16_to_f16_inplace<128>(sV, Cfg::TILE_BYTES, ct, 1.0f);
      fence_proxy_async_smem();
      mbar_arrive(&bars->kv_ready);
      if (NST > 1)
        for (int i = 0; i < NST && i < T; ++i) convert_tile(i);
    }
    int box = 0;  // staging box counter (box & 1 = buffer)
    for (int i = 0; i < T; ++i) {
      if (elected) HSTU_TSTAMP(6, i, 0);
      mbar_wait(&bars->tile_done[i & 3], (i >> 2) & 1);
      tc_fence_after_sync();
      if (elected) HSTU_TSTAMP(6, i, 1);
      if (elected && i + NST < T) load_tile(i + NST);
      const int qpos = q_tile(i) * 128 + row;
      const bool q_ok = qpos < len;                // rows past the end of this sequence belong to the next one: add zeros
#pragma unroll
      for (int ps = 0; ps < Cfg::DQ_NPASS; ++ps, ++box) {
        const uint32_t sbox = smem_u32(sDQS + (box & 1) * Cfg::DQS_BYTES);
#ifdef HSTU_EXP_BWD_NO_DRAIN
        if (ps == Cfg::DQ_NPASS - 1) {  // ablation (wrong numerics): no dQ traffic at all
          tc_fence_before_sync();
          mbar_arrive(&bars->dq_empty[i % Cfg::NDQ]);
        }
        continue;
#endif
        if (elected) bulk_wait_group_read1();      // the reduce that used this box (two boxes ago) has finished reading it
        named_bar_sync(1, 128);
        if constexpr (Cfg::PSM) {
          // 72 registers in this mode: two loads of 16 columns instead of one of 32 (which spilled 22 registers per tile)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[16];
            tmem_ld16(tmem + Cfg::TMEM_DQ + (i % Cfg::NDQ) * D + ps * 32 + hh * 16 + lane_bits, r);
            tmem_ld_wait();
            if (hh == 1 && ps == Cfg::DQ_NPASS - 1) {
              tc_fence_before_sync();
              mbar_arrive(&bars->dq_empty[i % Cfg::NDQ]);
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              st_shared_v4(sbox + swizzled_chunk_offset<128>(row, hh * 4 + (e >> 2)), q_ok ? r[e] : 0u, q_ok ? r[e + 1] : 0u,
                           q_ok ? r[e + 2] : 0u, q_ok ? r[e + 3] : 0u);
          }
        } else {
        uint32_t r[32];
        tmem_ld32(tmem + Cfg::TMEM_DQ + (i % Cfg::NDQ) * D + ps * 32 + lane_bits, r);
        tmem_ld_wait();
        if (ps == Cfg::DQ_NPASS - 1) {
          tc_fence_before_sync();
          mbar_arrive(&bars->dq_empty[i % Cfg::NDQ]);
        }
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          st_shared_v4(sbox + swizzled_chunk_offset<128>(row, e >> 2), q_ok ? r[e] : 0u, q_ok ? r[e + 1] : 0u,
                       q_ok ? r[e + 2] : 0u, q_ok ? r[e + 3] : 0u);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (elected) {
          tma_reduce_add_3d(&p.tmDQ, sbox, ps * 32, h, (int)(row0 + q_tile(i) * 128));
          bulk_commit_group();
        }
      }
      // Convert the tile whose load was issued ONE iteration ago (it has landed by now).  Waiting here for the load issued
      // above would put a full TMA round trip into every iteration of this loop -- and this loop paces the dQ accumulators and
      // the Q / dO stages of the whole CTA (r02: 2.25 -> 2.9 ms with that wait).
      if (elected) HSTU_TSTAMP(6, i, 2);
      if (CONV && NST >= 4) {
        if (i >= 1 && i - 1 + NST < T) convert_tile(i - 1 + NST);
      } else if (CONV && NST > 1 && i + NST < T) {
        convert_tile(i + NST);  // three stages (d = 64): the lag would leave the issuer one tile short; take the wait instead
      }
      if (elected) HSTU_TSTAMP(6, i, 3);
    }
    if (elected) bulk_wait_group_read0();          // shared memory must stay valid until the last reduce has read it
  } else {
    if constexpr (Cfg::PSM) reg_alloc<120>(); else reg_alloc<176>();
    // ---------------- elementwise warpgroups ----------------
    const int wg = (warp - 4) >> 2;                // owns query columns [64*wg, 64*wg + 64) of every tile
    const int quad = warp & 3;
    const int row = quad * 32 + lane;              // key row inside the tile == TMEM lane
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const int j_pos = n0 + row;
    const float2 ah2 = make_float2(p.alpha_half, p.alpha_half);
    // dS^T is an fp16 tensor-core operand: it is stored as dS * 2^-e.  With bf16 inputs the factor is already in dO (applied
    // when dO was converted); with fp16 inputs it is folded into the constants of the sigmoid-derivative polynomial here.
    // It is removed again in the dK / dV epilogues and in dq_convert_kernel.
    // The derivative factor g = sig (1 + x (1 - sig)) is evaluated as (1 + g2) / 2 with g2 = t + h (1 - t^2), t = tanh(h), h = x / 2
    // (3 packed FMAs); the 1/2 is left out here and applied with the other factors in the epilogues: dS^T holds 2 * 2^-e * dS.
    const float2 one2v = make_float2(1.0f, 1.0f);
    const float2 esc2v = make_float2(ds_scale, ds_scale);
    const bool fast = msk.fast != 0;
    const bool j_ok = j_pos < len;
    const bool j_hist = j_ok && (!msk.has_tgt || j_pos < msk.max_id);  // fast mask: valid = (j_hist & i > j) | (i == j)
    const int cbase = wg * 64;
    const bool stamp = quad == 0 && lane == 0;

    // p = x sig(x) and 2 dS = dP (1 + g2) from one tanh; packed fp32x2 arithmetic (FMUL2 / FFMA2): two elements per issued
    // instruction, one MUFU.TANH per element
#define HSTU_BWD_ELEM2(S0, S1, DP0, DP1, P0, P1, D0, D1)                                                       \
  {                                                                                                            \
    const float2 hh = __fmul2_rn(make_float2(__uint_as_float(S0), __uint_as_float(S1)), ah2);                  \
    const float2 t = make_float2(tanh_approx(hh.x), tanh_approx(hh.y));                                        \
    const float2 pv = __ffma2_rn(hh, t, hh);                    /* p = x sig(x) = h (1 + t)            */      \
    const float2 w = __ffma2_rn(make_float2(-t.x, -t.y), t, one2v);  /* 1 - t^2                          */      \
    const float2 g2 = __ffma2_rn(hh, w, t);                     /* 2 g - 1 = t + h (1 - t^2)           */      \
    float2 dpe = make_float2(__uint_as_float(DP0), __uint_as_float(DP1));                                      \
    if (!CONV) dpe = __fmul2_rn(dpe, esc2v);                    /* fp16 inputs: dO is not pre-scaled   */      \
    const float2 dv = __ffma2_rn(dpe, g2, dpe);                 /* 2 dS = dP (1 + g2)                  */      \
    P0 = pv.x; P1 = pv.y; D0 = dv.x; D1 = dv.y;                                                                \
  }
    // NE consecutive query columns starting at column `col0` of this half: scores / dP (SREF(e), DREF(e) = element e of the run)
    // -> packed fp16 P^T (PP[e / 2]) and dS^T (DD[e / 2]).  ONE block of straight-line code per mask mode: the tile-uniform branch
    // sits outside, so the scheduler can interleave all NE / 2 independent MUFU -> FFMA2 chains (r02: with the branch and the
    // stores inside 16-column chunks every chunk exposed its own latency: 2750 clk per tile instead of 1700).
#define HSTU_BWD_RUN(NE, SREF, DREF, PP, DD, col0)                                                             \
  if (mode == 0) {                                                                                             \
    _Pragma("unroll") for (int e = 0; e < NE; e += 2) {                                                        \
      float p0, p1, d0, d1;                                                                                    \
      HSTU_BWD_ELEM2(SREF(e), SREF(e + 1), DREF(e), DREF(e + 1), p0, p1, d0, d1);                              \
      PP[e >> 1] = pack_f16x2_sat(p0, p1);                                                                     \
      DD[e >> 1] = pack_f16x2_sat(d0, d1);                                                                     \
    }                                                                                                          \
  } else if (mode == 1) {                                                                                      \
    /* valid(i, j) = ((j is history) & (i > j)) | (i == j), restricted to i < len and j < len */               \
    const int lo_c = j_hist ? jr : 0x7fffffff;       /* columns > lo_c are valid (if j is a history position) */ \
    const int dg_c = j_ok ? jr : -0x7fffffff;        /* the diagonal column */                                 \
    _Pragma("unroll") for (int e = 0; e < NE; e += 2) {                                                        \
      float p0, p1, d0, d1;                                                                                    \
      HSTU_BWD_ELEM2(SREF(e), SREF(e + 1), DREF(e), DREF(e + 1), p0, p1, d0, d1);                              \
      const int c0 = (col0) + e;                                                                               \
      const bool v0 = ((c0 > lo_c) | (c0 == dg_c)) & (c0 < len_rel);                                           \
      const bool v1 = ((c0 + 1 > lo_c) | (c0 + 1 == dg_c)) & (c0 + 1 < len_rel);                               \
      p0 = v0 ? p0 : 0.f; d0 = v0 ? d0 : 0.f;                                                                  \
      p1 = v1 ? p1 : 0.f; d1 = v1 ? d1 : 0.f;                                                                  \
      PP[e >> 1] = pack_f16x2_sat(p0, p1);                                                                     \
      DD[e >> 1] = pack_f16x2_sat(d0, d1);                                                                     \
    }                                                                                                          \
  } else {                                                                                                     \
    _Pragma("unroll") for (int e = 0; e < NE; e += 2) {                                                        \
      float p0, p1, d0, d1;                                                                                    \
      HSTU_BWD_ELEM2(SREF(e), SREF(e + 1), DREF(e), DREF(e + 1), p0, p1, d0, d1);                              \
      const int i_pos = m0 + cbase + (col0) + e;                                                               \
      const bool v0 = j_ok && i_pos < len && mask_valid(msk, i_pos, j_pos);                                    \
      const bool v1 = j_ok && i_pos + 1 < len && mask_valid(msk, i_pos + 1, j_pos);                            \
      p0 = v0 ? p0 : 0.f; d0 = v0 ? d0 : 0.f;                                                                  \
      p1 = v1 ? p1 : 0.f; d1 = v1 ? d1 : 0.f;                                                                  \
      PP[e >> 1] = pack_f16x2_sat(p0, p1);                                                                     \
      DD[e >> 1] = pack_f16x2_sat(d0, d1);                                                                     \
    }                                                                                                          \
  }

    if constexpr (Cfg::PSM) {
      // warpgroup wg takes units wg, wg + 3, ...: unit u = 2 i + hf lives in slot u % 3 == wg
      for (int u = wg; u < 2 * T; u += 3) {
        const int i = u >> 1, hf = u & 1, slot = wg;
        const int m0 = q_tile(i) * 128;
        const int cbase = hf * 64;                    // query columns [64 hf, 64 hf + 64) of the tile
        if (stamp) HSTU_TSTAMP(wg == 2 ? 7 : 2 + wg, u, 0);
        mbar_wait(&bars->s_full[slot], (u / 3) & 1);
        tc_fence_after_sync();
        if (stamp) HSTU_TSTAMP(wg == 2 ? 7 : 2 + wg, u, 1);
        const int mh0 = m0 + cbase;
        const bool full = fast && (mh0 >= n0 + 128) && (mh0 + 64 <= len) && (!msk.has_tgt || n0 + 128 <= msk.max_id);
        const int mode = full ? 0 : (fast ? 1 : 2);
        const uint32_t sDSTw = smem_u32(sDST + (i & 1) * Cfg::PT_BYTES + hf * 16384);
        const uint32_t sPTw = smem_u32(sPT + (u % Cfg::NPB) * 16384);
        const int jr = j_pos - m0 - cbase;
        const int len_rel = len - m0 - cbase;
        const uint32_t st_addr = tmem + Cfg::TMEM_SLOT + slot * Cfg::SLOT_COLS + lane_bits;
        const uint32_t dp_addr = st_addr + Cfg::HALF_COLS;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t s[32], dp[32];
          tmem_ld32(st_addr + c * 32, s);
          tmem_ld32(dp_addr + c * 32, dp);
          tmem_ld_wait();
          if (c == 1) {
            tc_fence_before_sync();
            mbar_arrive(&bars->scores_free[slot]);     // both chunks are in registers: the slot goes back to the score issuer
          } else {
            if (i >= 2) mbar_wait(&bars->tile_done[(i - 2) & 3], ((i - 2) >> 2) & 1);   // GEMMs of tile i-2 are done with the dS^T box pair
            if (u >= Cfg::NPB) mbar_wait(&bars->p_free[u % Cfg::NPB], ((u / Cfg::NPB) - 1) & 1);  // dV of unit u - NPB has read this P^T box
          }
          uint32_t pp[16], dd[16];
#define HSTU_S32(e) s[e]
#define HSTU_D32(e) dp[e]
          HSTU_BWD_RUN(32, HSTU_S32, HSTU_D32, pp, dd, c * 32);
#undef HSTU_S32
#undef HSTU_D32
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            st_shared_v4(sPTw + swizzled_chunk_offset<128>(row, c * 4 + j4), pp[4 * j4], pp[4 * j4 + 1], pp[4 * j4 + 2], pp[4 * j4 + 3]);
            st_shared_v4(sDSTw + swizzled_chunk_offset<128>(row, c * 4 + j4), dd[4 * j4], dd[4 * j4 + 1], dd[4 * j4 + 2], dd[4 * j4 + 3]);
          }
        }
        fence_proxy_async_smem();
        if (stamp) HSTU_TSTAMP(wg == 2 ? 7 : 2 + wg, u, 2);
        mbar_arrive(&bars->unit_done[hf * 2 + (i & 1)]);
      }
    } else
    for (int i = 0; i < T; ++i) {
      const int u = 2 * i + wg, slot = u % Cfg::NSLOT;
      const int m0 = q_tile(i) * 128;
      if (stamp) HSTU_TSTAMP(2 + wg, i, 0);
      if (CONV && NST == 1) {
        // single Q / dO stage: the tile has just landed and nothing else can run before it is converted, so both warpgroups
        // do it (this one: Q, the other: dO times 2^-e), 256 threads instead of the 128 of the drain warpgroup
        mbar_wait(&bars->q_full[0], i & 1);
        if (wg == 0) convert_bf16_to_f16_inplace<128>(sQ, Cfg::TILE_BYTES, tid - 128, 1.0f);
        else convert_bf16_to_f16_inplace<128>(sDO, Cfg::TILE_BYTES, tid - 256, ds_scale);
        fence_proxy_async_smem();
        mbar_arrive(&bars->q_ready[0]);
      }
      mbar_wait(&bars->s_full[u % Cfg::NSF], (u / Cfg::NSF) & 1);
      tc_fence_after_sync();
      if (stamp) HSTU_TSTAMP(2 + wg, i, 1);
      // classification of this half-tile (uniform over the warpgroup)
      const int mh0 = m0 + cbase;                   // first query row of the half
      const bool full = fast && (mh0 >= n0 + 128) && (mh0 + 64 <= len) && (!msk.has_tgt || n0 + 128 <= msk.max_id);
      const int mode = full ? 0 : (fast ? 1 : 2);
      const uint32_t sDSTw = smem_u32(sDST + (i & 1) * Cfg::PT_BYTES + wg * 16384);
      const int jr = j_pos - m0 - cbase;           // query column (relative to this warpgroup's half) equal to j
      const int len_rel = len - m0 - cbase;        // columns >= len_rel are past the sequence end
      const uint32_t st_addr = tmem + Cfg::TMEM_SLOT + slot * Cfg::SLOT_COLS + lane_bits;   // the slot holds one half-tile {S^T | dP^T}
      const uint32_t dp_addr = st_addr + Cfg::HALF_COLS;
#ifdef HSTU_BWD_PRING_NOPF
      if (Cfg::PRING) {
        // PRING without the register-hungry prefetch: the chunk loop of the ring path (64 live inputs), the slot goes back to the
        // score issuer as soon as the SECOND chunk is in registers, P^T goes to its own ring
        const uint32_t p_addr = tmem + Cfg::TMEM_P + (u % Cfg::NPR) * 32 + lane_bits;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t s[32], dp[32];
          tmem_ld32(st_addr + c * 32, s);
          tmem_ld32(dp_addr + c * 32, dp);
          tmem_ld_wait();
          if (c == 1) {
            tc_fence_before_sync();
            mbar_arrive(&bars->scores_free[wg]);
          } else {
            if (i >= 2) mbar_wait(&bars->tile_done[(i - 2) & 3], ((i - 2) >> 2) & 1);
            if (u >= Cfg::NPR) {
              mbar_wait(&bars->p_free[u % Cfg::NPR], ((u / Cfg::NPR) - 1) & 1);
              tc_fence_after_sync();
            }
          }
          uint32_t pp[16], dd[16];
#define HSTU_S32(e) s[e]
#define HSTU_D32(e) dp[e]
          HSTU_BWD_RUN(32, HSTU_S32, HSTU_D32, pp, dd, c * 32);
#undef HSTU_S32
#undef HSTU_D32
          tmem_st16(p_addr + c * 16, pp);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            st_shared_v4(sDSTw + swizzled_chunk_offset<128>(row, c * 4 + j4), dd[4 * j4], dd[4 * j4 + 1], dd[4 * j4 + 2], dd[4 * j4 + 3]);
        }
      } else
#endif
      if (Cfg::PRING) {
        // Two chunks of 32 query columns; the second is loaded while the first is processed, and as soon as it has landed the
        // slot goes back to the issuer (scores_free): the score GEMMs of this warpgroup's next unit run during the second half
        // of its arithmetic.  (Loading all 64 + 64 values first and processing them as one block left ptxas no registers to
        // interleave the last MUFU -> FFMA2 chains: 2500 clk per unit instead of 1700, profiles/r02_bwd_timelines.txt (E).)
        uint32_t s[2][32], dp[2][32];
        tmem_ld32(st_addr, s[0]);
        tmem_ld32(dp_addr, dp[0]);
        tmem_ld_wait();
        tmem_ld32(st_addr + 32, s[1]);
        tmem_ld32(dp_addr + 32, dp[1]);
        const uint32_t p_addr = tmem + Cfg::TMEM_P + (u % Cfg::NPR) * 32 + lane_bits;  // 64 fp16 = 32 columns: A of the dV GEMM
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pp[16], dd[16];
#define HSTU_SC(e) s[c][e]
#define HSTU_DC(e) dp[c][e]
          HSTU_BWD_RUN(32, HSTU_SC, HSTU_DC, pp, dd, c * 32);
#undef HSTU_SC
#undef HSTU_DC
          if (c == 0) {
            tmem_ld_wait();               // chunk 1 is in registers: the slot is free
            tc_fence_before_sync();
            mbar_arrive(&bars->scores_free[wg]);
            if (i >= 2) mbar_wait(&bars->tile_done[(i - 2) & 3], ((i - 2) >> 2) & 1);  // GEMMs of tile i-2 are done with this box pair
            if (u >= Cfg::NPR) {
              mbar_wait(&bars->p_free[u % Cfg::NPR], ((u / Cfg::NPR) - 1) & 1);         // dV of unit u - NPR has consumed the P^T buffer
              tc_fence_after_sync();
            }
          }
          tmem_st16(p_addr + c * 16, pp);
          // dS^T [kv][q] (16-byte stores): A of dK as stored, A of dQ read MN-major
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            st_shared_v4(sDSTw + swizzled_chunk_offset<128>(row, c * 4 + j4), dd[4 * j4], dd[4 * j4 + 1], dd[4 * j4 + 2], dd[4 * j4 + 3]);
        }
      } else {
#ifdef HSTU_EXP_BWD_NO_ELEM
        if (i >= 2) mbar_wait(&bars->tile_done[(i - 2) & 3], ((i - 2) >> 2) & 1);  // ablation (wrong numerics): barriers only
        if (T < 0)
#endif
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // 2 chunks of 32 query columns
          uint32_t s[32], dp[32];
          tmem_ld32(st_addr + c * 32, s);
          tmem_ld32(dp_addr + c * 32, dp);
          tmem_ld_wait();
          if (c == 0 && i >= 2) mbar_wait(&bars->tile_done[(i - 2) & 3], ((i - 2) >> 2) & 1);  // GEMMs of tile i-2 are done with this box pair
          uint32_t pp[16], dd[16];
#define HSTU_S32(e) s[e]
#define HSTU_D32(e) dp[e]
          HSTU_BWD_RUN(32, HSTU_S32, HSTU_D32, pp, dd, c * 32);
#undef HSTU_S32
#undef HSTU_D32
          // P^T chunk c (32 fp16 = 16 columns) overwrites the already-read front of the S^T half of the slot: A of the dV GEMM
          tmem_st16(st_addr + c * 16, pp);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            st_shared_v4(sDSTw + swizzled_chunk_offset<128>(row, c * 4 + j4), dd[4 * j4], dd[4 * j4 + 1], dd[4 * j4 + 2], dd[4 * j4 + 3]);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      fence_proxy_async_smem();
      if (stamp) HSTU_TSTAMP(2 + wg, i, 2);
      mbar_arrive(&bars->unit_done[wg * 2 + (i & 1)]);
    }
#undef HSTU_BWD_RUN
#undef HSTU_BWD_ELEM2
    // ---------------- epilogue: dV (warpgroup 0) / dK (warpgroup 1): TMEM -> scale -> global ----------------
    mbar_wait(&bars->fin_full, 0);
    tc_fence_after_sync();
    const bool is_dv = wg == 0;
    if (wg < 2) {
    const int ecol0 = 0;
    const uint32_t acc = tmem + (is_dv ? Cfg::TMEM_DV : Cfg::TMEM_DK) + ecol0 + lane_bits;
    // undo 2^-e (a power of two: exact): dK always carries it, dV only when dO itself was scaled
    const float scale = is_dv ? (CONV ? p.dv_scale / ds_scale : p.dv_scale) : 0.5f * p.dk_scale / ds_scale;  // dS^T holds 2 * 2^-e * dS
    uint16_t* gptr = (is_dv
        ? reinterpret_cast<uint16_t*>(p.dv) + (row0 + j_pos) * p.dv_row_stride + (long long)h * p.dv_head_stride
        : reinterpret_cast<uint16_t*>(p.dk) + (row0 + j_pos) * p.dk_row_stride + (long long)h * p.dk_head_stride) + ecol0;
#pragma unroll
    for (int c = 0; c < D / 16; ++c) {
      uint32_t o[16];
      tmem_ld16(acc + c * 16, o);
      tmem_ld_wait();
      if (j_ok) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float a = __uint_as_float(o[e]) * scale, bb = __uint_as_float(o[e + 1]) * scale;
          pk[e >> 1] = BF16 ? pack_bf16x2(a, bb) : pack_f16x2(a, bb);
        }
        uint4* dst = reinterpret_cast<uint4*>(gptr + c * 16);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// dq[r, h, :] = convert(dq_acc[r, h, :] * scale)
template <bool BF16>
__global__ void dq_convert_kernel(const float* __restrict__ acc, uint16_t* __restrict__ dq, long long rows, int heads, int D,
                                  long long row_stride, long long head_stride, float scale_in,
                                  const uint32_t* __restrict__ amax_bits) {
  const float scale = 0.5f * scale_in / ds_scale_from_amax(__ldg(amax_bits));  // dq_acc holds sums of 2 * 2^-e * dS K (no alpha)
  const long long nvec = rows * heads * (D / 8);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nvec; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % (D / 8));
    const long long rh = idx / (D / 8);
    const int hh = (int)(rh % heads);
    const long long r = rh / heads;
    const float4 a = *reinterpret_cast<const float4*>(acc + rh * D + v * 8);
    const float4 c = *reinterpret_cast<const float4*>(acc + rh * D + v * 8 + 4);
    uint4 o;
    if (BF16) {
      o = make_uint4(pack_bf16x2(a.x * scale, a.y * scale), pack_bf16x2(a.z * scale, a.w * scale),
                     pack_bf16x2(c.x * scale, c.y * scale), pack_bf16x2(c.z * scale, c.w * scale));
    } else {
      o = make_uint4(pack_f16x2(a.x * scale, a.y * scale), pack_f16x2(a.z * scale, a.w * scale),
                     pack_f16x2(c.x * scale, c.y * scale), pack_f16x2(c.z * scale, c.w * scale));
    }
    *reinterpret_cast<uint4*>(dq + r * row_stride + hh * head_stride + v * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static bool aligned_view(const void* ptr, long long row_stride, long long head_stride) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (row_stride % 8) == 0 && (head_stride % 8) == 0;
}

static bool umma_bwd_supported(const hstu_attn_params& p) {
  if (!umma_fwd_supported(p)) return false;  // dtype / dims / alignment of q, k, v (out is not used by the backward)
  // d = 256: dK and dV alone fill the 512 TMEM columns; the forward has a tcgen05 kernel, the backward runs on the generic path
  if (p.dqk != 32 && p.dqk != 64 && p.dqk != 128) return false;
  return aligned_view(p.dout, p.do_row_stride, p.do_head_stride) && aligned_view(p.dq, p.dq_row_stride, p.dq_head_stride) &&
         aligned_view(p.dk, p.dk_row_stride, p.dk_head_stride) && aligned_view(p.dv_out, p.dv_row_stride, p.dv_head_stride);
}

bool umma_supported(const hstu_attn_params& p, bool bwd) {
  if (!bwd) return umma_fwd_supported(p);
  hstu_attn_params q = p;
  if (q.out == nullptr) q.out = const_cast<void*>(q.q);  // the forward check also looks at `out`
  q.o_row_stride = 8;
  q.o_head_stride = 8;
  return umma_bwd_supported(q);
}

size_t umma_workspace_bytes(const hstu_attn_params& p, bool bwd) {
  if (!bwd) return 0;
  // fp32 dQ accumulator [L, H, D] + one 256-byte slot for the max |dO| word
  return (size_t)p.total_rows * p.heads * p.dqk * sizeof(float) + 256;
}

template <int D, bool BF16>
static int launch_bwd_umma(const hstu_attn_params& p, cudaStream_t st) {
  using Cfg = BwdCfg<D>;
  const size_t need = umma_workspace_bytes(p, true);
  if (p.workspace == nullptr || p.workspace_bytes < need) {
    set_error("hstu_attn_bwd: workspace of %zu bytes required (got %zu)", need, p.workspace_bytes);
    return HSTU_ERR_WORKSPACE;
  }
  BwdParams bp;
  memset(&bp, 0, sizeof(bp));
  if (int e = make_tmap_rows_heads(&bp.tmQ, p.q, p.total_rows, p.heads, D, p.q_row_stride, p.q_head_stride, Cfg::BOX_COLS, 128)) return e;
  if (int e = make_tmap_rows_heads(&bp.tmK, p.k, p.total_rows, p.heads, D, p.k_row_stride, p.k_head_stride, Cfg::BOX_COLS, 128)) return e;
  if (int e = make_tmap_rows_heads(&bp.tmV, p.v, p.total_rows, p.heads, D, p.v_row_stride, p.v_head_stride, Cfg::BOX_COLS, 128)) return e;
  if (int e = make_tmap_rows_heads(&bp.tmDO, p.dout, p.total_rows, p.heads, D, p.do_row_stride, p.do_head_stride, Cfg::BOX_COLS, 128)) return e;
  if (int e = make_tmap_rows_heads_f32(&bp.tmDQ, p.workspace, p.total_rows, p.heads, D, Cfg::DQ_BOX_COLS, 128)) return e;
  bp.seq_offsets = p.seq_offsets;
  bp.num_targets = p.num_targets;
  bp.dk = p.dk;
  bp.dv = p.dv_out;
  bp.dq_acc = reinterpret_cast<float*>(p.workspace);
  uint32_t* amax_bits = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(p.workspace) + (need - 256));
  bp.dout_amax_bits = amax_bits;
  bp.dk_row_stride = p.dk_row_stride;
  bp.dk_head_stride = p.dk_head_stride;
  bp.dv_row_stride = p.dv_row_stride;
  bp.dv_head_stride = p.dv_head_stride;
  bp.offsets_i64 = p.offsets_are_i64;
  bp.targets_i64 = p.num_targets_are_i64;
  bp.max_seq_len = p.max_seq_len;
  bp.heads = p.heads;
  bp.win = p.max_attn_len;
  bp.min_full = p.min_full_attn_seq_len;
  bp.ctx = p.contextual_seq_len;
  bp.alpha_half = 0.5f * p.alpha;
  bp.dv_scale = 1.0f / (float)p.max_seq_len;
  bp.dk_scale = p.alpha / (float)p.max_seq_len;
  HSTU_CUDA_OK(cudaMemsetAsync(p.workspace, 0, need, st));
  const long long nvec = p.total_rows * p.heads * (D / 8);
  long long blocks = (nvec + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  dout_amax_kernel<BF16><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(p.dout), p.total_rows, p.heads, D,
                                                      p.do_row_stride, p.do_head_stride, amax_bits);
  HSTU_CUDA_OK(cudaGetLastError());
  auto kern = attn_bwd_umma_kernel<D, BF16>;
  HSTU_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  dim3 grid((p.max_seq_len + 127) / 128, p.heads, p.batch);
#ifdef HSTU_TRACE
  long long* tbuf = nullptr;
  const size_t tbytes = sizeof(long long) * 8 * 256 * 4;
  cudaMalloc(&tbuf, tbytes);
  cudaMemset(tbuf, 0, tbytes);
  cudaMemcpyToSymbol(g_trace, &tbuf, sizeof(tbuf));
#endif
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(bp);
  HSTU_CUDA_OK(cudaGetLastError());
#ifdef HSTU_TRACE
  {
    cudaDeviceSynchronize();
    static long long host[8 * 256 * 4];
    cudaMemcpy(host, tbuf, tbytes, cudaMemcpyDeviceToHost);
    FILE* f = fopen("gpurun_out/bwd_trace.txt", "w");
    if (f) {
      for (int r = 0; r < 8; ++r)
        for (int i = 0; i < 256; ++i) {
          const long long* e = host + (r * 256 + i) * 4;
          if (e[0] || e[1] || e[2] || e[3]) fprintf(f, "%d %d %lld %lld %lld %lld\n", r, i, e[0], e[1], e[2], e[3]);
        }
      fclose(f);
    }
    cudaFree(tbuf);
    tbuf = nullptr;
    cudaMemcpyToSymbol(g_trace, &tbuf, sizeof(tbuf));
  }
#endif
  blocks = (nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  dq_convert_kernel<BF16><<<(int)blocks, 256, 0, st>>>(bp.dq_acc, reinterpret_cast<uint16_t*>(p.dq), p.total_rows, p.heads, D,
                                                       p.dq_row_stride, p.dq_head_stride, bp.dk_scale, amax_bits);
  HSTU_CUDA_OK(cudaGetLastError());
  return 0;
}

int attn_umma_bwd(const hstu_attn_params& p, cudaStream_t st) {
  const bool bf = p.dtype == HSTU_BF16;
  switch (p.dqk) {
    case 32: return bf ? launch_bwd_umma<32, true>(p, st) : launch_bwd_umma<32, false>(p, st);
    case 64: return bf ? launch_bwd_umma<64, true>(p, st) : launch_bwd_umma<64, false>(p, st);
    case 128: return bf ? launch_bwd_umma<128, true>(p, st) : launch_bwd_umma<128, false>(p, st);
  }
  set_error("tcgen05 backward: unsupported head dim %d", p.dqk);
  return HSTU_ERR_UNSUPPORTED;
}

}  // namespace hstu
