// Jagged HSTU attention forward on the 5th-gen tensor cores (tcgen05 + TMEM) with TMA-staged tiles.  bf16 / fp16,
// dqk == dv in {32, 64, 128, 256}.  Two kernels share the roles below: attn_fwd_umma_kernel (one CTA per work item; long sequences)
// and attn_fwd_umma_persist_kernel (one CTA per SM walking a static item list with all rings / phases on a global key-tile
// counter; short sequences, see its header further down).
//
// One CTA per (128-row query tile, head, sequence); heavy (late) query tiles are scheduled first.  Warp roles:
//   warp 0  lane 0 : TMA producer for Q (once) and the K tiles (3 stages; stage t % 3 is reused once s_full of tile t
//                    has completed, i.e. the score GEMM has read it)
//   warp 3  lane 0 : TMA producer for the V tiles (3 stages, reused once pv_done of tile t has completed)
//   warp 1  lane 0 : tcgen05.mma issuer 1:  S_t = Q K_t^T  (SS, both K-major) into TMEM slot t % 3, one commit per tile
//   warp 2         : TMEM allocation / release; lane 0 = issuer 2:  O += P_t V_t  (TS: A = P_t read from the front of
//                    slot t % 3, B = V_t MN-major), rotating over NACC accumulators, one commit per tile.
//                    Two issuing threads because tcgen05.commit and the barrier waits stall only their own issuer
//                    (umma_selftest: one thread 71-84 clk / MMA, two threads 41-47).
//   warps 4-7, 8-11: two "silu" warpgroups.  Warpgroup g owns key tiles t = g, g+2, ...: tcgen05.ld S (one query row per
//                    thread, next 32 columns prefetched), p = silu(alpha*s) * mask via one MUFU (tanh) + packed fp32x2
//                    FMUL2 / FFMA2, packs bf16 pairs and writes them with tcgen05.st over the S columns already read.
//   warps 12-15    : (bf16 inputs only) converter warpgroup: kind::f16 needs ONE operand format per instruction (a mixed fp16 x
//                    bf16 descriptor is an illegal instruction on B200) and the parity budget needs P in fp16 (11-bit
//                    significand; a bf16 P alone costs 1.7e-3 of relative error).  So the V tiles -- the B operand of P.V --
//                    are converted bf16 -> fp16 IN PLACE in shared memory when their TMA load lands (exact for every bf16
//                    value in the fp16 range); S = Q K^T stays bf16 x bf16.  V has slack (it is needed only after the silu
//                    stage), so the hop is off the critical path; its cost is shared-memory bandwidth (one read + one write of
//                    the tile).  fp16 inputs skip this stage.
// The 1/N factor of the reference is applied once in the epilogue (O tile: TMEM -> registers -> 128-bit global stores,
// rows past the sequence end are not written).  Rows of neighbouring sequences that a 128-row TMA box drags in are
// neutralised by the mask (P = 0 for key positions >= len), never by re-reading memory.
//
// Reference semantics: ops/pytorch/pt_hstu_attention.py:130-171; tile skipping mirrors the idea of
// ops/triton/triton_hstu_attention.py:517-543 (loop bounds from the mask) but is derived from common.cuh's ranges.
#include "common.cuh"
#include "internal.h"
#include "umma.cuh"

namespace hstu {
using namespace umma;

struct alignas(64) FwdParams {
  CUtensorMap tmQ, tmK, tmV;
  const void* seq_offsets;
  const void* num_targets;
  void* out;
  long long o_row_stride, o_head_stride;
  int offsets_i64, targets_i64;
  int max_seq_len;
  int win, min_full, ctx;
  float alpha_half;  // alpha / 2
  float inv_n;       // 1 / max_seq_len
  int heads, batch;  // persistent kernel: the work items are enumerated inside the kernel
};

#ifdef HSTU_FWD_PSMEM
constexpr bool kFwdPsmem = true;
#else
constexpr bool kFwdPsmem = false;
#endif

template <int D>
struct FwdCfg {
  static constexpr int SW = (D * 2 >= 128) ? 128 : D * 2;  // swizzle width (bytes) of the Q/K/V boxes
  static constexpr int BOX_COLS = SW / 2;
  static constexpr int NBOX = D / BOX_COLS;
  static constexpr int BOX_BYTES = 128 * SW;
  static constexpr int TILE_BYTES = 128 * D * 2;
  // K / V TMA rings and the ring of score slots in TMEM.  d <= 128: three of each (stage = slot = tile % 3).  d = 256: a tile is
  // 64 KB, so ONE K and ONE V stage next to Q, and the 256-column O accumulator leaves room for TWO 128-column score slots.  A K
  // stage is released by the score GEMM that read it (s_full), a V stage and a score slot by the P.V GEMM (pv_done).
  static constexpr int STAGES = (D <= 128) ? 3 : 1;
  static constexpr int NSLOT = (D <= 128) ? 3 : 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + TILE_BYTES;
  static constexpr int OFF_V = OFF_K + STAGES * TILE_BYTES;
  // PSM (d <= 64): P goes to shared memory (three boxes of [128 q][128 keys] fp16, K-major, 128-byte swizzle: the A operand of an SS
  // P.V GEMM) instead of back into its score slot.  The slot then returns to the score issuer as soon as the silu warpgroup has
  // LOADED it (slot_free) instead of after the P.V GEMM of the tile: the serial chain Q K^T -> silu -> P V -> Q K^T of a slot -- what
  // bounds this kernel at d = 32 (profiles/r02_ablations.txt) -- loses its last two links.
  static constexpr bool PSM = kFwdPsmem && D <= 64;
  static constexpr int P_BYTES = 128 * 128 * 2;            // two 128-byte-swizzle boxes of 64 keys
  static constexpr int OFF_P = OFF_V + STAGES * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_P + (PSM ? 3 * P_BYTES : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;  // + barriers + alignment slack
  // ring of 3 score slots of 128 columns; P (bf16 pairs, 64 columns) overwrites the front of its own S slot
  static constexpr int TMEM_S = 0;
  // O is accumulated in NACC independent TMEM accumulators (k-step ks of every P.V goes to accumulator ks % NACC; they are
  // summed in the epilogue): consecutive tcgen05.mma into ONE accumulator are dependent and expose the MMA latency when
  // the instruction itself is short (N = d = 32: 16 clk of work).
  static constexpr int NACC = (D <= 32) ? 4 : (D <= 64 ? 2 : 1);
  static constexpr int TMEM_O = NSLOT * 128;  // accumulators: columns [TMEM_O, TMEM_O + NACC * D)
  static_assert(TMEM_O + NACC * D <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct FwdBars {
  uint64_t q_full;
  uint64_t k_full[3], v_full[3];
  uint64_t v_ready[3];  // bf16 inputs: the V tile has been converted to fp16 (128 converter threads)
  uint64_t s_full[3], p_full[3], pv_done[3];
  uint64_t slot_free[3];  // PSM: the silu warpgroup has loaded the scores of the tile out of the slot (128 arrivals)
  uint64_t o_full;
  uint32_t tmem_base;
};

// THREE silu warpgroups for d <= 64 (tile i -> warpgroup i % 3 == its score slot), two otherwise: at d <= 64 the kernel is bound by
// the MUFU pipe (one tanh per score; ncu: 70 % busy with two warpgroups) and each warpgroup leaves it idle while it waits for
// its next scores, loads and stores; a third warp per scheduler fills those gaps.
#ifdef HSTU_FWD_3WG
template <int D> constexpr int kFwdSiluWgs = (D <= 64) ? 3 : 2;
#else
template <int D> constexpr int kFwdSiluWgs = 2;
#endif
template <int D, bool BF16> constexpr int kFwdThreads = 128 * (1 + kFwdSiluWgs<D> + (BF16 ? 1 : 0));

template <int D, bool BF16>
__global__ void __launch_bounds__(kFwdThreads<D, BF16>, 1) attn_fwd_umma_kernel(const __grid_constant__ FwdParams p) {
  using Cfg = FwdCfg<D>;
  constexpr int NWG = kFwdSiluWgs<D>;
  constexpr bool CONV = BF16;  // bf16 V tiles are converted to fp16 in shared memory: P.V runs fp16 x fp16
  constexpr int SW = Cfg::SW;
  constexpr int NST = Cfg::STAGES;
  constexpr int NSL = Cfg::NSLOT;
  const int b = blockIdx.z, h = blockIdx.y;
  const int m0 = (int)(gridDim.x - 1 - blockIdx.x) * 128;
  const long long row0 = load_index(p.seq_offsets, p.offsets_i64, b);
  int len = (int)(load_index(p.seq_offsets, p.offsets_i64, b + 1) - row0);
  if (len > p.max_seq_len) {  // rows past max_seq_len are ignored on the way in and zero on the way out
    if (blockIdx.x == 0) zero_rows(p.out, 2, p.o_row_stride, (long long)h * p.o_head_stride, D, row0 + p.max_seq_len, row0 + len);
    len = p.max_seq_len;
  }
  if (m0 >= len) return;
  const int n_tgt = p.num_targets ? (int)load_index(p.num_targets, p.targets_i64, b) : -1;
  const SeqMask msk = make_seq_mask(len, n_tgt, p.win, p.min_full, p.ctx);
  const int mrows = min(128, len - m0);
  int lo, hi;
  kv_range_for_q_rows(msk, m0, m0 + mrows, &lo, &hi);
  const int t0 = lo / 128;
  const int T = (hi + 127) / 128 - t0;  // >= 1 (the diagonal tile)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sP = smem + Cfg::OFF_P;   // PSM only
  FwdBars* bars = reinterpret_cast<FwdBars*>(smem + Cfg::OFF_BAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 128);
      mbar_init(&bars->pv_done[i], 1);
      mbar_init(&bars->slot_free[i], 128);
      mbar_init(&bars->v_ready[i], 128);
    }
    mbar_init(&bars->o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;
  uint64_t* const q_rdy = &bars->q_full;
  uint64_t* const k_rdy = bars->k_full;
  uint64_t* const v_rdy = CONV ? bars->v_ready : bars->v_full;
  // bf16 inputs run 512 threads (128 registers / thread at launch): per-warpgroup budgets are set at the top of each role

  if (warp >= 4 + 4 * NWG) {
    // ---------------- converter warpgroup (bf16 inputs): TMA-landed tile -> fp16 in place -> ready ----------------
    if constexpr (NWG == 3) reg_dealloc<56>(); else reg_dealloc<64>();
    const int t = tid - 128 * (1 + NWG);
    for (int i = 0; i < T; ++i) {
      const int st = i % NST;
      mbar_wait(&bars->v_full[st], (i / NST) & 1);
      convert_bf16_to_f16_inplace<128>(sV + st * Cfg::TILE_BYTES, Cfg::TILE_BYTES, t, 1.0f);
      fence_proxy_async_smem();
      mbar_arrive(&bars->v_ready[st]);
    }
  } else if (warp < 4) {
   if (CONV) { if constexpr (NWG == 3) reg_dealloc<56>(); else reg_dealloc<80>(); }
   if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: Q, then K tiles ----------------
      prefetch_tensormap(&p.tmQ);
      prefetch_tensormap(&p.tmK);
      mbar_arrive_expect_tx(&bars->q_full, Cfg::TILE_BYTES);
#pragma unroll
      for (int bx = 0; bx < Cfg::NBOX; ++bx)
        tma_load_3d(sQ + bx * Cfg::BOX_BYTES, &p.tmQ, &bars->q_full, bx * Cfg::BOX_COLS, h, (int)(row0 + m0));
      for (int i = 0; i < T; ++i) {
        const int st = i % NST;
        if (i >= NST) mbar_wait(&bars->s_full[(i - NST) % NSL], ((i - NST) / NSL) & 1);  // Q K_{i-NST}^T has consumed this stage
#ifdef HSTU_EXP_NO_KLOAD
        if (i >= NST) {  // ablation experiment only
          mbar_arrive(&bars->k_full[st]);
          continue;
        }
#endif
        mbar_arrive_expect_tx(&bars->k_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int bx = 0; bx < Cfg::NBOX; ++bx)
          tma_load_3d(sK + st * Cfg::TILE_BYTES + bx * Cfg::BOX_BYTES, &p.tmK, &bars->k_full[st], bx * Cfg::BOX_COLS, h,
                      (int)(row0 + (long long)(t0 + i) * 128));
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ---------------- TMA producer: V tiles ----------------
      prefetch_tensormap(&p.tmV);
      for (int i = 0; i < T; ++i) {
        const int st = i % NST;
        if (i >= NST) mbar_wait(&bars->pv_done[(i - NST) % NSL], ((i - NST) / NSL) & 1);  // P_{i-NST} V_{i-NST} has consumed this stage
#ifdef HSTU_EXP_NO_VLOAD
        if (i >= NST) {  // ablation experiment only: no TMA traffic for V after the ring has been filled once
          mbar_arrive(&bars->v_full[st]);
          continue;
        }
#endif
        mbar_arrive_expect_tx(&bars->v_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int bx = 0; bx < Cfg::NBOX; ++bx)
          tma_load_3d(sV + st * Cfg::TILE_BYTES + bx * Cfg::BOX_BYTES, &p.tmV, &bars->v_full[st], bx * Cfg::BOX_COLS, h,
                      (int)(row0 + (long long)(t0 + i) * 128));
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer 1: S_i = Q K_i^T ----------------
    // Two issuing threads (this warp and warp 2): every tcgen05.mma blocks its issuer ~45 clk and every tcgen05.commit
    // ~100-200 clk (measured, profiles/r01_umma_selftest.txt), so one thread issuing 10 MMAs + 4 commits per tile was the
    // critical path of the whole kernel.  The whole warp runs the warp-uniform control flow, one fixed lane issues;
    // descriptors are built once and only their address field is advanced.  One commit per tile and thread: K / V stages
    // are released through the slot barriers (stage = slot = tile % 3).
    const bool leader = lane == 0;
    constexpr uint32_t idesc_qk = make_idesc(128, 128, false, false, BF16, BF16);    // Q, K in the input dtype
    const uint64_t dq0 = desc_kmajor<SW>(smem_u32(sQ), 0);
    const uint64_t dk0 = desc_kmajor<SW>(smem_u32(sK), 0);
    mbar_wait(q_rdy, 0);
    for (int i = 0; i < T; ++i) {
      const int st = i % NST, sl = i % NSL;
      if (i >= NSL) {
        if constexpr (Cfg::PSM) mbar_wait(&bars->slot_free[sl], ((i / NSL) - 1) & 1);  // S_{i-NSL} has been loaded out of this slot
        else mbar_wait(&bars->pv_done[sl], ((i / NSL) - 1) & 1);                        // P_{i-NSL} (front of this slot) has been consumed
      }
      mbar_wait(&k_rdy[st], (i / NST) & 1);
      tc_fence_after_sync();
      const uint64_t kd = dk0 + (uint64_t)((st * Cfg::TILE_BYTES) >> 4);
      const uint32_t ts = tmem + Cfg::TMEM_S + sl * 128;
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint32_t kb = ks * 32, bx = kb / SW, off = kb % SW;
          const uint64_t o = (uint64_t)((bx * Cfg::BOX_BYTES + off) >> 4);
          mma_ss(ts, dq0 + o, kd + o, idesc_qk, ks > 0);
        }
        mma_commit(&bars->s_full[sl]);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer 2: O += P_i V_i (A = P from TMEM, B = V read MN-major) ----------------
    const bool leader = lane == 0;
    constexpr uint32_t idesc_pv = make_idesc(128, D, false, true, false, false);     // P (TMEM) and V both fp16
    const uint64_t dv0 = desc_mnmajor<SW>(smem_u32(sV), 0, Cfg::BOX_BYTES);
    for (int i = 0; i < T; ++i) {
      const int st = i % NST, sl = i % NSL;
      mbar_wait(&bars->p_full[sl], (i / NSL) & 1);   // P_i sits in TMEM (front of its score slot)
      mbar_wait(&v_rdy[st], (i / NST) & 1);
      tc_fence_after_sync();
      const uint64_t vd = dv0 + (uint64_t)((st * Cfg::TILE_BYTES) >> 4);
      const uint32_t tp = tmem + Cfg::TMEM_S + sl * 128;
      if (leader) {
        if constexpr (Cfg::PSM) {
          // A = P box sl in shared memory: [128 q][64 keys] x 2 boxes, K-major; k-step ks = 16 keys = 32 bytes inside the 128-byte row
          const uint64_t pd = desc_kmajor<128>(smem_u32(sP), 0) + (uint64_t)((sl * Cfg::P_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_ss(tmem + Cfg::TMEM_O + (ks % Cfg::NACC) * D, pd + (uint64_t)(((ks >> 2) * 16384 + (ks & 3) * 32) >> 4),
                   vd + (uint64_t)((ks * 16 * SW) >> 4), idesc_pv, (i > 0) || (ks >= Cfg::NACC));
        } else {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // 16 bf16 of K per 8 TMEM columns
          mma_ts(tmem + Cfg::TMEM_O + (ks % Cfg::NACC) * D, tp + ks * 8, vd + (uint64_t)((ks * 16 * SW) >> 4), idesc_pv,
                 (i > 0) || (ks >= Cfg::NACC));
        }
        mma_commit(&bars->pv_done[sl]);
      }
      __syncwarp();
    }
    if (leader) mma_commit(&bars->o_full);
    __syncwarp();
   }
  } else {
    // ---------------- silu warpgroups ----------------
    if (CONV) { if constexpr (NWG == 3) reg_alloc<120>(); else reg_alloc<168>(); }
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;              // query row inside the tile == TMEM lane
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const int i_pos = m0 + row;
    const float2 ah2 = make_float2(p.alpha_half, p.alpha_half);
    const bool fast = msk.fast != 0;
    // plain causal (+targets): valid(i, j) = (j < lim_i) | (j == i)   (common.cuh: mask_valid, fast path)
    const int lim_i = msk.has_tgt ? min(i_pos, msk.max_id) : i_pos;
    const int full_lim = fast ? min(m0, msk.has_tgt ? msk.max_id : 0x7fffffff) : -1;
    for (int i = wg, it = 0; i < T; i += NWG, ++it) {
      const uint32_t s_taddr = tmem + Cfg::TMEM_S + (i % NSL) * 128 + lane_bits;
      mbar_wait(&bars->s_full[i % NSL], (i / NSL) & 1);
      tc_fence_after_sync();
      const int n0 = (t0 + i) * 128;
      const int mode = (n0 + 128 <= full_lim) ? 0 : (fast ? 1 : 2);  // tile-uniform: no divergence
      const int lim_rel = lim_i - n0, diag_rel = i_pos - n0;
      uint32_t sbuf[2][32];
#ifdef HSTU_EXP_NO_ELEM
      if (T < 0)  // ablation experiment only: skip the whole elementwise stage
#endif
      tmem_ld32(s_taddr, sbuf[0]);
#pragma unroll
#ifdef HSTU_EXP_NO_ELEM
      for (int c = 0; c < (T < 0 ? 4 : 0); ++c) {
#else
      for (int c = 0; c < 4; ++c) {
#endif
        tmem_ld_wait();
        if (c < 3) tmem_ld32(s_taddr + (c + 1) * 32, sbuf[(c + 1) & 1]);  // prefetch the next 32 columns
        if constexpr (Cfg::PSM) {
          if (c == 3) {                      // the last chunk is in registers: the slot goes back to the score issuer
            tc_fence_before_sync();
            mbar_arrive(&bars->slot_free[i % NSL]);
          }
        }
        const uint32_t(&s)[32] = sbuf[c & 1];
        uint32_t pk[16];
        if (mode == 0) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 hh = __fmul2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), ah2);
            const float2 pv = __ffma2_rn(hh, make_float2(tanh_approx(hh.x), tanh_approx(hh.y)), hh);
            const float p0 = pv.x, p1 = pv.y;
            pk[e >> 1] = pack_f16x2_sat(p0, p1);
          }
        } else if (mode == 1) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const int j0 = c * 32 + e;
            const float2 hh = __fmul2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), ah2);
            const float2 pv = __ffma2_rn(hh, make_float2(tanh_approx(hh.x), tanh_approx(hh.y)), hh);
            float p0 = pv.x, p1 = pv.y;
            p0 = ((j0 < lim_rel) | (j0 == diag_rel)) ? p0 : 0.f;
            p1 = ((j0 + 1 < lim_rel) | (j0 + 1 == diag_rel)) ? p1 : 0.f;
            pk[e >> 1] = pack_f16x2_sat(p0, p1);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const int j = n0 + c * 32 + e;
            const float2 hh = __fmul2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), ah2);
            const float2 pv = __ffma2_rn(hh, make_float2(tanh_approx(hh.x), tanh_approx(hh.y)), hh);
            float p0 = pv.x, p1 = pv.y;
            p0 = (j < len && mask_valid(msk, i_pos, j)) ? p0 : 0.f;
            p1 = (j + 1 < len && mask_valid(msk, i_pos, j + 1)) ? p1 : 0.f;
            pk[e >> 1] = pack_f16x2_sat(p0, p1);
          }
        }
        if constexpr (Cfg::PSM) {
          if (c == 0 && i >= NSL) mbar_wait(&bars->pv_done[i % NSL], ((i / NSL) - 1) & 1);  // P V of tile i - 3 has read this P box
          // P chunk c = keys [32 c, 32 c + 32) of this thread's query row: four 16-byte pieces of box c / 2
          const uint32_t pbox = smem_u32(sP + (i % NSL) * Cfg::P_BYTES + (c >> 1) * 16384);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            st_shared_v4(pbox + swizzled_chunk_offset<128>(row, (c & 1) * 4 + j4), pk[4 * j4], pk[4 * j4 + 1], pk[4 * j4 + 2], pk[4 * j4 + 3]);
        } else {
        // P chunk c (32 bf16 = 16 columns) goes to columns [16 c, 16 c + 16) of the slot: a region of S that has already
        // been read (S chunk c covers columns [32 c, 32 c + 32))
        tmem_st16(s_taddr + c * 16, pk);
        }
      }
      if constexpr (Cfg::PSM) {
        fence_proxy_async_smem();
      } else {
        tmem_st_wait();
        tc_fence_before_sync();
      }
      mbar_arrive(&bars->p_full[i % NSL]);
    }
    // ---------------- epilogue: O (TMEM) -> * 1/N -> global ----------------
    mbar_wait(&bars->o_full, 0);
    tc_fence_after_sync();
    constexpr int HALF = D / 2;
    const int cbase = (wg & 1) * HALF;
    if (wg < 2) {
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + (row0 + m0 + row) * p.o_row_stride + (long long)h * p.o_head_stride + cbase;
#pragma unroll
    for (int c = 0; c < HALF / 16; ++c) {
      uint32_t o[16];
      tmem_ld16(tmem + Cfg::TMEM_O + cbase + c * 16 + lane_bits, o);
      tmem_ld_wait();
#pragma unroll
      for (int a = 1; a < Cfg::NACC; ++a) {  // sum the independent accumulators
        uint32_t o2[16];
        tmem_ld16(tmem + Cfg::TMEM_O + a * D + cbase + c * 16 + lane_bits, o2);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) + __uint_as_float(o2[e]));
      }
      if (row < mrows) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float a = __uint_as_float(o[e]) * p.inv_n, bb = __uint_as_float(o[e + 1]) * p.inv_n;
          pk[e >> 1] = BF16 ? pack_bf16x2(a, bb) : pack_f16x2(a, bb);
        }
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 16);
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


// ------------------------------------------------------------------------------------------------
// Persistent forward (d <= 64: max_seq_len <= 4096; d = 128: max_seq_len <= 512): ONE CTA per SM walks a static list of work items (query tile, head, sequence), heavy tiles
// first.  TMEM is allocated and the barriers are initialised once; the K / V rings, the score slots and the barrier phases run on
// a GLOBAL key-tile counter, so the loads and score GEMMs of item k + 1 start while item k is still in its silu / P.V / epilogue
// phase.  Q has two buffers (q_full / q_empty), O two TMEM accumulator sets (o_full / o_empty) for the same reason.  Short sequences
// (a handful of key tiles per item) otherwise pay TMEM allocation, barrier set-up, the first TMA round trip, the pipeline drain and
// the epilogue once per CTA with nothing to overlap them: ~5 us per item against ~1 us of work at Lmax = 512.
// Measured on B200 (fwd ms, bf16, one-CTA-per-item kernel -> this one): B 512, H 4, d 64: Lmax 512 0.349 -> 0.267, Lmax 2048 2.179 ->
// 2.102; B 512, H 4, d 32, Lmax 512: 0.324 -> 0.244; B 128, H 8, d 32, Lmax 256: 0.089 -> 0.062; headline (B 16, H 8, d 32, Lmax 8192):
// 1.328 -> 1.344, which is why long sequences keep the other kernel (HSTU_FWD_PERSIST=0 / 1 in the environment forces either).
// ------------------------------------------------------------------------------------------------
template <int D>
struct FwdPCfg {
  using C = FwdCfg<D>;
  // d <= 64: two Q buffers and two O accumulator sets of 64 columns; d = 128: one of each (a 128-column O next to the three score
  // slots fills the tensor memory, two 32 KB Q tiles would not fit the shared memory) -- the next item's K / V loads and score GEMMs
  // still overlap the current item, its P.V GEMMs wait for the epilogue
  static constexpr int NQ = (D <= 64) ? 2 : 1;
  static constexpr int NO = (D <= 64) ? 2 : 1;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NQ * C::TILE_BYTES;
  static constexpr int OFF_V = OFF_K + 3 * C::TILE_BYTES;
  static constexpr int OFF_BAR = OFF_V + 3 * C::TILE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static constexpr int O_COLS = (D <= 64) ? 64 : D;                 // columns of one O accumulator set
  static constexpr int NACC = O_COLS / D;                           // independent accumulators per set (summed in the epilogue)
  static constexpr int TMEM_O = 3 * 128;                            // set b: columns [TMEM_O + O_COLS b, + O_COLS)
  static_assert(D <= 128 && TMEM_O + NO * O_COLS <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct FwdPBars {
  uint64_t q_full[2], q_empty[2];
  uint64_t k_full[3], v_full[3], v_ready[3];
  uint64_t s_full[3], p_full[3], pv_done[3];
  uint64_t o_full[2], o_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(FwdPBars) <= 512, "barrier block");

struct FwdItem {
  int b, h, m0, len, len_true, T, t0, mrows;
  long long row0;
  SeqMask msk;
};

// item k of the static list: query tiles in descending order (heavy first), then (sequence, head)
__device__ __forceinline__ FwdItem fwd_item(const FwdParams& p, long long k, int ntiles) {
  FwdItem it;
  const int hb = p.heads * p.batch;
  const int qt = ntiles - 1 - (int)(k / hb);
  const int r = (int)(k % hb);
  it.h = r % p.heads;
  it.b = r / p.heads;
  it.m0 = qt * 128;
  it.row0 = load_index(p.seq_offsets, p.offsets_i64, it.b);
  it.len_true = (int)(load_index(p.seq_offsets, p.offsets_i64, it.b + 1) - it.row0);
  it.len = min(it.len_true, p.max_seq_len);
  it.T = 0;
  it.t0 = 0;
  it.mrows = 0;
  if (it.m0 < it.len) {
    const int n_tgt = p.num_targets ? (int)load_index(p.num_targets, p.targets_i64, it.b) : -1;
    it.msk = make_seq_mask(it.len, n_tgt, p.win, p.min_full, p.ctx);
    it.mrows = min(128, it.len - it.m0);
    int lo, hi;
    kv_range_for_q_rows(it.msk, it.m0, it.m0 + it.mrows, &lo, &hi);
    it.t0 = lo / 128;
    it.T = (hi + 127) / 128 - it.t0;
  }
  return it;
}

template <int D, bool BF16>
__global__ void __launch_bounds__(BF16 ? 512 : 384, 1) attn_fwd_umma_persist_kernel(const __grid_constant__ FwdParams p) {
  using Cfg = FwdCfg<D>;
  using PC = FwdPCfg<D>;
  constexpr bool CONV = BF16;
  constexpr int SW = Cfg::SW;
  const int ntiles = (p.max_seq_len + 127) / 128;
  const long long n_items = (long long)ntiles * p.heads * p.batch;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + PC::OFF_Q;
  uint8_t* sK = smem + PC::OFF_K;
  uint8_t* sV = smem + PC::OFF_V;
  FwdPBars* bars = reinterpret_cast<FwdPBars*>(smem + PC::OFF_BAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->q_empty[i], 1);
      mbar_init(&bars->o_full[i], 1);
      mbar_init(&bars->o_empty[i], 256);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_ready[i], 128);
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 128);
      mbar_init(&bars->pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;
  uint64_t* const v_rdy = CONV ? bars->v_ready : bars->v_full;

  // every role walks the same item list with the same two counters: g0 = key tiles before this item, n = non-empty items before it
  if (warp >= 12) {
    reg_dealloc<64>();
    const int t = tid - 384;
    int g0 = 0;
    for (long long k = blockIdx.x; k < n_items; k += gridDim.x) {
      const FwdItem it = fwd_item(p, k, ntiles);
      for (int i = 0; i < it.T; ++i) {
        const int g = g0 + i, st = g % 3;
        mbar_wait(&bars->v_full[st], (g / 3) & 1);
        convert_bf16_to_f16_inplace<128>(sV + st * Cfg::TILE_BYTES, Cfg::TILE_BYTES, t, 1.0f);
        fence_proxy_async_smem();
        mbar_arrive(&bars->v_ready[st]);
      }
      g0 += it.T;
    }
  } else if (warp < 4) {
    if (CONV) reg_dealloc<80>();
    if (warp == 0) {
      if (lane == 0) {
        // ---------------- TMA producer: Q of every item, K tiles ----------------
        prefetch_tensormap(&p.tmQ);
        prefetch_tensormap(&p.tmK);
        int g0 = 0, n = 0;
        for (long long k = blockIdx.x; k < n_items; k += gridDim.x) {
          const FwdItem it = fwd_item(p, k, ntiles);
          if (it.T == 0) continue;
          const int qb = n % PC::NQ;
          if (n >= PC::NQ) mbar_wait(&bars->q_empty[qb], ((n / PC::NQ) - 1) & 1);   // the score GEMMs of item n - NQ have read this buffer
          mbar_arrive_expect_tx(&bars->q_full[qb], Cfg::TILE_BYTES);
#pragma unroll
          for (int bx = 0; bx < Cfg::NBOX; ++bx)
            tma_load_3d(sQ + qb * Cfg::TILE_BYTES + bx * Cfg::BOX_BYTES, &p.tmQ, &bars->q_full[qb], bx * Cfg::BOX_COLS, it.h,
                        (int)(it.row0 + it.m0));
          for (int i = 0; i < it.T; ++i) {
            const int g = g0 + i, st = g % 3;
            if (g >= 3) mbar_wait(&bars->s_full[st], ((g / 3) - 1) & 1);    // Q K^T of key tile g - 3 has consumed this stage
            mbar_arrive_expect_tx(&bars->k_full[st], Cfg::TILE_BYTES);
#pragma unroll
            for (int bx = 0; bx < Cfg::NBOX; ++bx)
              tma_load_3d(sK + st * Cfg::TILE_BYTES + bx * Cfg::BOX_BYTES, &p.tmK, &bars->k_full[st], bx * Cfg::BOX_COLS, it.h,
                          (int)(it.row0 + (long long)(it.t0 + i) * 128));
          }
          g0 += it.T;
          ++n;
        }
      }
    } else if (warp == 3) {
      if (lane == 0) {
        // ---------------- TMA producer: V tiles ----------------
        prefetch_tensormap(&p.tmV);
        int g0 = 0;
        for (long long k = blockIdx.x; k < n_items; k += gridDim.x) {
          const FwdItem it = fwd_item(p, k, ntiles);
          for (int i = 0; i < it.T; ++i) {
            const int g = g0 + i, st = g % 3;
            if (g >= 3) mbar_wait(&bars->pv_done[st], ((g / 3) - 1) & 1);   // P V of key tile g - 3 has consumed this stage
            mbar_arrive_expect_tx(&bars->v_full[st], Cfg::TILE_BYTES);
#pragma unroll
            for (int bx = 0; bx < Cfg::NBOX; ++bx)
              tma_load_3d(sV + st * Cfg::TILE_BYTES + bx * Cfg::BOX_BYTES, &p.tmV, &bars->v_full[st], bx * Cfg::BOX_COLS, it.h,
                          (int)(it.row0 + (long long)(it.t0 + i) * 128));
          }
          g0 += it.T;
        }
      }
    } else if (warp == 1) {
      // ---------------- MMA issuer 1: S = Q K^T of every key tile of every item ----------------
      const bool leader = lane == 0;
      constexpr uint32_t idesc_qk = make_idesc(128, 128, false, false, BF16, BF16);
      const uint64_t dq0 = desc_kmajor<SW>(smem_u32(sQ), 0);
      const uint64_t dk0 = desc_kmajor<SW>(smem_u32(sK), 0);
      int g0 = 0, n = 0;
      for (long long k = blockIdx.x; k < n_items; k += gridDim.x) {
        const FwdItem it = fwd_item(p, k, ntiles);
        if (it.T == 0) continue;
        const int qb = n % PC::NQ;
        mbar_wait(&bars->q_full[qb], (n / PC::NQ) & 1);
        const uint64_t qd = dq0 + (uint64_t)((qb * Cfg::TILE_BYTES) >> 4);
        for (int i = 0; i < it.T; ++i) {
          const int g = g0 + i, st = g % 3;
          if (g >= 3) mbar_wait(&bars->pv_done[st], ((g / 3) - 1) & 1);     // P of key tile g - 3 (front of this slot) has been consumed
          mbar_wait(&bars->k_full[st], (g / 3) & 1);
          tc_fence_after_sync();
          const uint64_t kd = dk0 + (uint64_t)((st * Cfg::TILE_BYTES) >> 4);
          const uint32_t ts = tmem + st * 128;
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint32_t kb = ks * 32, bx = kb / SW, off = kb % SW;
              const uint64_t o = (uint64_t)((bx * Cfg::BOX_BYTES + off) >> 4);
              mma_ss(ts, qd + o, kd + o, idesc_qk, ks > 0);
            }
            mma_commit(&bars->s_full[st]);
            if (i == it.T - 1) mma_commit(&bars->q_empty[qb]);               // every score GEMM of this item has read Q
          }
          __syncwarp();
        }
        g0 += it.T;
        ++n;
      }
    } else {
      // ---------------- MMA issuer 2: O += P V ----------------
      const bool leader = lane == 0;
      constexpr uint32_t idesc_pv = make_idesc(128, D, false, true, false, false);
      const uint64_t dv0 = desc_mnmajor<SW>(smem_u32(sV), 0, Cfg::BOX_BYTES);
      int g0 = 0, n = 0;
      for (long long k = blockIdx.x; k < n_items; k += gridDim.x) {
        const FwdItem it = fwd_item(p, k, ntiles);
        if (it.T == 0) continue;
        const int ob = n % PC::NO;
        if (n >= PC::NO) {                                                   // the epilogue of item n - NO has read this O buffer
          mbar_wait(&bars->o_empty[ob], ((n / PC::NO) - 1) & 1);
          tc_fence_after_sync();
        }
        for (int i = 0; i < it.T; ++i) {
          const int g = g0 + i, st = g % 3;
          mbar_wait(&bars->p_full[st], (g / 3) & 1);
          mbar_wait(&v_rdy[st], (g / 3) & 1);
          tc_fence_after_sync();
          const uint64_t vd = dv0 + (uint64_t)((st * Cfg::TILE_BYTES) >> 4);
          const uint32_t tp = tmem + st * 128;
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_ts(tmem + PC::TMEM_O + ob * PC::O_COLS + (ks % PC::NACC) * D, tp + ks * 8, vd + (uint64_t)((ks * 16 * SW) >> 4), idesc_pv,
                     (i > 0) || (ks >= PC::NACC));
            mma_commit(&bars->pv_done[st]);
            if (i == it.T - 1) mma_commit(&bars->o_full[ob]);
          }
          __syncwarp();
        }
        g0 += it.T;
        ++n;
      }
    }
  } else {
    // ---------------- silu warpgroups: key tiles with (global index) % 2 == wg; then half of the O tile each ----------------
    if (CONV) reg_alloc<168>();
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const float2 ah2 = make_float2(p.alpha_half, p.alpha_half);
    int g0 = 0, n = 0;
    for (long long k = blockIdx.x; k < n_items; k += gridDim.x) {
      const FwdItem it = fwd_item(p, k, ntiles);
      if (it.m0 == 0 && it.len_true > it.len) {
        // rows of this sequence beyond max_seq_len: zeros (the lightest query tile of every (sequence, head) does it)
        const long long nz = (long long)(it.len_true - it.len) * D;
        uint16_t* ob = reinterpret_cast<uint16_t*>(p.out);
        for (long long idx = tid - 128; idx < nz; idx += 256) {
          const long long r = it.row0 + it.len + idx / D;
          ob[r * p.o_row_stride + (long long)it.h * p.o_head_stride + idx % D] = 0;
        }
      }
      if (it.T == 0) continue;
      const SeqMask msk = it.msk;
      const int m0 = it.m0, len = it.len;
      const int i_pos = m0 + row;
      const bool fast = msk.fast != 0;
      const int lim_i = msk.has_tgt ? min(i_pos, msk.max_id) : i_pos;
      const int full_lim = fast ? min(m0, msk.has_tgt ? msk.max_id : 0x7fffffff) : -1;
      for (int i = 0; i < it.T; ++i) {
        const int g = g0 + i;
        if ((g & 1) != wg) continue;
        const int st = g % 3;
        const uint32_t s_taddr = tmem + st * 128 + lane_bits;
        mbar_wait(&bars->s_full[st], (g / 3) & 1);
        tc_fence_after_sync();
        const int n0 = (it.t0 + i) * 128;
        const int mode = (n0 + 128 <= full_lim) ? 0 : (fast ? 1 : 2);
        const int lim_rel = lim_i - n0, diag_rel = i_pos - n0;
        uint32_t sbuf[2][32];
        tmem_ld32(s_taddr, sbuf[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c < 3) tmem_ld32(s_taddr + (c + 1) * 32, sbuf[(c + 1) & 1]);
          const uint32_t(&s)[32] = sbuf[c & 1];
          uint32_t pk[16];
          if (mode == 0) {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float2 hh = __fmul2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), ah2);
              const float2 pv = __ffma2_rn(hh, make_float2(tanh_approx(hh.x), tanh_approx(hh.y)), hh);
              pk[e >> 1] = pack_f16x2_sat(pv.x, pv.y);
            }
          } else if (mode == 1) {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const int j0 = c * 32 + e;
              const float2 hh = __fmul2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), ah2);
              const float2 pv = __ffma2_rn(hh, make_float2(tanh_approx(hh.x), tanh_approx(hh.y)), hh);
              float p0 = pv.x, p1 = pv.y;
              p0 = ((j0 < lim_rel) | (j0 == diag_rel)) ? p0 : 0.f;
              p1 = ((j0 + 1 < lim_rel) | (j0 + 1 == diag_rel)) ? p1 : 0.f;
              pk[e >> 1] = pack_f16x2_sat(p0, p1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const int j = n0 + c * 32 + e;
              const float2 hh = __fmul2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), ah2);
              const float2 pv = __ffma2_rn(hh, make_float2(tanh_approx(hh.x), tanh_approx(hh.y)), hh);
              float p0 = pv.x, p1 = pv.y;
              p0 = (j < len && mask_valid(msk, i_pos, j)) ? p0 : 0.f;
              p1 = (j + 1 < len && mask_valid(msk, i_pos, j + 1)) ? p1 : 0.f;
              pk[e >> 1] = pack_f16x2_sat(p0, p1);
            }
          }
          tmem_st16(s_taddr + c * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(&bars->p_full[st]);
      }
      // ---------------- epilogue of this item: O buffer n & 1 -> * 1/N -> global; then hand the buffer back ----------------
      const int ob = n % PC::NO;
      mbar_wait(&bars->o_full[ob], (n / PC::NO) & 1);
      tc_fence_after_sync();
      constexpr int HALF = D / 2;
      const int cbase = wg * HALF;
      uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + (it.row0 + m0 + row) * p.o_row_stride + (long long)it.h * p.o_head_stride + cbase;
#pragma unroll
      for (int c = 0; c < HALF / 16; ++c) {
        uint32_t o[16];
        tmem_ld16(tmem + PC::TMEM_O + ob * PC::O_COLS + cbase + c * 16 + lane_bits, o);
        tmem_ld_wait();
#pragma unroll
        for (int a = 1; a < PC::NACC; ++a) {
          uint32_t o2[16];
          tmem_ld16(tmem + PC::TMEM_O + ob * PC::O_COLS + a * D + cbase + c * 16 + lane_bits, o2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) + __uint_as_float(o2[e]));
        }
        if (row < it.mrows) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float a = __uint_as_float(o[e]) * p.inv_n, bb = __uint_as_float(o[e + 1]) * p.inv_n;
            pk[e >> 1] = BF16 ? pack_bf16x2(a, bb) : pack_f16x2(a, bb);
          }
          uint4* dst = reinterpret_cast<uint4*>(orow + c * 16);
          dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bars->o_empty[ob]);
      g0 += it.T;
      ++n;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int D, bool BF16>
static int launch_fwd_persist(const FwdParams& fp, const hstu_attn_params& p, cudaStream_t st) {
  using PC = FwdPCfg<D>;
  auto kern = attn_fwd_umma_persist_kernel<D, BF16>;
  HSTU_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PC::SMEM_BYTES));
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    HSTU_CUDA_OK(cudaGetDevice(&dev));
    HSTU_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long items = (long long)((p.max_seq_len + 127) / 128) * p.heads * p.batch;
  const int grid = (int)(items < n_sm ? items : n_sm);
  kern<<<grid, BF16 ? 512 : 384, PC::SMEM_BYTES, st>>>(fp);
  HSTU_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static bool is_sm100() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0;
    cudaDeviceProp prop;
    cached = (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.major == 10) ? 1 : 0;
  }
  return cached == 1;
}

static bool aligned_view(const void* ptr, long long row_stride, long long head_stride) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (row_stride % 8) == 0 && (head_stride % 8) == 0;
}

bool umma_fwd_supported(const hstu_attn_params& p) {
  if (p.dtype != HSTU_BF16 && p.dtype != HSTU_F16) return false;
  if (p.dqk != p.dv || (p.dqk != 32 && p.dqk != 64 && p.dqk != 128 && p.dqk != 256)) return false;
  if (p.delta_q_len != 0 || p.pos_w != nullptr || p.ts_w != nullptr) return false;
  if (p.total_rows >= (1ll << 31) - 256) return false;
  if (!aligned_view(p.q, p.q_row_stride, p.q_head_stride) || !aligned_view(p.k, p.k_row_stride, p.k_head_stride) ||
      !aligned_view(p.v, p.v_row_stride, p.v_head_stride) || !aligned_view(p.out, p.o_row_stride, p.o_head_stride))
    return false;
  return is_sm100();
}

template <int D, bool BF16>
static int launch_fwd_umma(const hstu_attn_params& p, cudaStream_t st) {
  using Cfg = FwdCfg<D>;
  FwdParams fp;
  memset(&fp, 0, sizeof(fp));
  if (int e = make_tmap_rows_heads(&fp.tmQ, p.q, p.total_rows, p.heads, D, p.q_row_stride, p.q_head_stride, Cfg::BOX_COLS, 128)) return e;
  if (int e = make_tmap_rows_heads(&fp.tmK, p.k, p.total_rows, p.heads, D, p.k_row_stride, p.k_head_stride, Cfg::BOX_COLS, 128)) return e;
  if (int e = make_tmap_rows_heads(&fp.tmV, p.v, p.total_rows, p.heads, D, p.v_row_stride, p.v_head_stride, Cfg::BOX_COLS, 128)) return e;
  fp.seq_offsets = p.seq_offsets;
  fp.num_targets = p.num_targets;
  fp.out = p.out;
  fp.o_row_stride = p.o_row_stride;
  fp.o_head_stride = p.o_head_stride;
  fp.offsets_i64 = p.offsets_are_i64;
  fp.targets_i64 = p.num_targets_are_i64;
  fp.max_seq_len = p.max_seq_len;
  fp.win = p.max_attn_len;
  fp.min_full = p.min_full_attn_seq_len;
  fp.ctx = p.contextual_seq_len;
  fp.alpha_half = 0.5f * p.alpha;
  fp.inv_n = 1.0f / (float)p.max_seq_len;
  fp.heads = p.heads;
  fp.batch = p.batch;
  if constexpr (D <= 128) {
    static const int forced = [] { const char* e = getenv("HSTU_FWD_PERSIST"); return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0); }();
    // measured thresholds (profiles/r02_ablations.txt): d <= 64 wins up to Lmax 2048 (3.5 %) and loses 1 % at 8192; d = 128, with
    // one Q buffer and one O set, wins 12 % at Lmax 512 and loses 20 % at 2048
    constexpr int kPersistMaxLen = (D <= 64) ? 4096 : 512;
    if (forced == 1 || (forced < 0 && p.max_seq_len <= kPersistMaxLen)) return launch_fwd_persist<D, BF16>(fp, p, st);
  }
  auto kern = attn_fwd_umma_kernel<D, BF16>;
  HSTU_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  dim3 grid((p.max_seq_len + 127) / 128, p.heads, p.batch);
  kern<<<grid, kFwdThreads<D, BF16>, Cfg::SMEM_BYTES, st>>>(fp);
  HSTU_CUDA_OK(cudaGetLastError());
  return 0;
}

int attn_umma_fwd(const hstu_attn_params& p, cudaStream_t st) {
  const bool bf = p.dtype == HSTU_BF16;
  switch (p.dqk) {
    case 32: return bf ? launch_fwd_umma<32, true>(p, st) : launch_fwd_umma<32, false>(p, st);
    case 64: return bf ? launch_fwd_umma<64, true>(p, st) : launch_fwd_umma<64, false>(p, st);
    case 128: return bf ? launch_fwd_umma<128, true>(p, st) : launch_fwd_umma<128, false>(p, st);
    case 256: return bf ? launch_fwd_umma<256, true>(p, st) : launch_fwd_umma<256, false>(p, st);
  }
  set_error("tcgen05 forward: unsupported head dim %d", p.dqk);
  return HSTU_ERR_UNSUPPORTED;
}

}  // namespace hstu
