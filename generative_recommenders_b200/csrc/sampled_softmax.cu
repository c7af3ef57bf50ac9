// Fused sampled-softmax loss with dot-product similarity.
//
// Reference: research/modeling/sequential/losses/sampled_softmax.py:43-89 (SampledSoftmaxLoss.jagged_forward) with
// research/modeling/sequential/autoregressive_losses.py:73-121 (LocalNegativesSampler: ids = all_item_ids[randint], embeddings =
// l2-normalised rows of the item embedding table) and research/rails/similarities/dot_product_similarity_fn.py:31-67 (bmm):
//     pos_logit_i  = q_i . norm(pos_emb_i) / T
//     neg_logit_ir = q_i . norm(E[id_ir]) / T,   replaced by -5e4 where id_ir == positive id_i
//     loss_i       = -log_softmax([pos_logit_i, neg_logit_i1 .. neg_logit_iR])[0] = lse_i - pos_logit_i
//     loss         = sum_i w_i loss_i / sum_i w_i                                   (the final weighted mean is done by the host)
//     norm(x)      = x / max(||x||_2, eps)   (only if l2_norm)
// The eager path gathers a [N, R, D] tensor of negatives and runs a bmm over it; here the R table rows of a query are read once,
// straight from the table, by the warp that owns the query: G = D / 8 (16-bit) or D / 4 (fp32) lanes hold one row as one 128-bit
// load each, so a warp handles 32 / G negatives per step; dot product and sum of squares are reduced inside the lane group.
// HBM / L2-gather bound: N R D e bytes of table rows against 2 N R D flops.
// Saved for the backward: logits [N, R + 1] (column 0 = positive) and the reciprocal norms [N, R + 1] (fp32).
// Backward: with p = softmax(logits_i), c_i = dloss w_i / sum w:   dlogit = c_i (p - [j == 0]) / T  (masked negatives: 0),
//     dq_i += dlogit_j  ehat_j,      d e_j = dlogit_j rn_j (q_i - ehat_j (ehat_j . q_i))   (rn_j = 1 / max(||e_j||, eps); when the norm
//     is clamped, or without l2_norm, d e_j = dlogit_j rn_j q_i),   table gradient accumulated with fp32 vector atomics.
#include "common.cuh"
#include "internal.h"

namespace hstu {

template <typename T> struct Vec16;  // 16 bytes of T
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };
template <> struct Vec16<__half> { static constexpr int N = 8; };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&f)[Vec16<T>::N]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int i = 0; i < Vec16<T>::N; ++i) f[i] = Cvt<T>::to_f(e[i]);
}

__device__ __forceinline__ float group_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per query row.  G lanes per table row (D = G * VEC, G a power of two <= 32).
template <typename T>
__global__ void __launch_bounds__(256) ssl_fwd_kernel(const SslArgs a) {
  constexpr int VEC = Vec16<T>::N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * 8 + warp;
  if (i >= a.N) return;
  const int G = a.D / VEC, NPW = 32 / G;           // lanes per row, rows per warp step
  const int g = lane / G, gl = lane % G;           // which row of the step, which slice of it
  const T* q = reinterpret_cast<const T*>(a.q) + i * a.D + gl * VEC;
  float qf[VEC];
  load_vec<T>(q, qf);
  const float inv_t = 1.0f / a.temperature;
  float* logits = a.logits + i * (a.R + 1);
  float* rnorm = a.rnorm + i * (a.R + 1);
  const long long pos_id = a.pos_ids[i];
  const int64_t* ids = a.neg_ids + i * a.R;
  float m = -INFINITY, ssum = 0.f;                 // running max / sum of exp over the logits this lane group has seen
  // j = 0 is the positive (its embedding comes from pos_emb), j = 1..R the sampled negatives
  for (int j0 = 0; j0 < a.R + 1; j0 += NPW) {
    const int j = j0 + g;
    const bool ok = j <= a.R;
    float ef[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) ef[k] = 0.f;
    long long id = -1;
    if (ok) {
      const T* src;
      if (j == 0) {
        src = reinterpret_cast<const T*>(a.pos_emb) + i * a.D;
      } else {
        id = ids[j - 1];
        src = reinterpret_cast<const T*>(a.table) + id * a.D;
      }
      load_vec<T>(src + gl * VEC, ef);
    }
    float dot = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      dot += qf[k] * ef[k];
      sq += ef[k] * ef[k];
    }
    dot = group_sum(dot, G);
    sq = group_sum(sq, G);
    const float rn = a.l2_norm ? 1.0f / fmaxf(sqrtf(sq), a.l2_eps) : 1.0f;
    float logit = dot * rn * inv_t;
    if (j > 0 && id == pos_id) logit = -5e4f;      // sampled_softmax.py:80-84
    if (ok) {
      if (gl == 0) {
        logits[j] = logit;
        rnorm[j] = rn;
      }
      const float nm = fmaxf(m, logit);
      ssum = ssum * __expf(m - nm) + __expf(logit - nm);
      m = nm;
    }
  }
  // merge the (max, sum) pairs of the NPW lane groups (lanes of one group hold identical values)
  for (int o = G; o < 32; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, ssum, o);
    const float nm = fmaxf(m, m2);
    ssum = (m == -INFINITY ? 0.f : ssum * __expf(m - nm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - nm));
    m = nm;
  }
  if (lane == 0) {
    const float lse = m + logf(ssum);
    a.lse[i] = lse;
    a.loss_rows[i] = lse - logits[0];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) ssl_bwd_kernel(const SslArgs a) {
  constexpr int VEC = Vec16<T>::N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * 8 + warp;
  if (i >= a.N) return;
  const int G = a.D / VEC, NPW = 32 / G;
  const int g = lane / G, gl = lane % G;
  float qf[VEC], dq[VEC];
  load_vec<T>(reinterpret_cast<const T*>(a.q) + i * a.D + gl * VEC, qf);
#pragma unroll
  for (int k = 0; k < VEC; ++k) dq[k] = 0.f;
  const float inv_t = 1.0f / a.temperature;
  const float c = a.row_coef[i] * inv_t;           // dloss * w_i / sum(w) / T
  const float lse = a.lse[i];
  const float* logits = a.logits + i * (a.R + 1);
  const float* rnorm = a.rnorm + i * (a.R + 1);
  const long long pos_id = a.pos_ids[i];
  const int64_t* ids = a.neg_ids + i * a.R;
  for (int j0 = 0; j0 < a.R + 1; j0 += NPW) {
    const int j = j0 + g;
    if (j > a.R) continue;
    long long id = -1;
    const T* src;
    if (j == 0) {
      src = reinterpret_cast<const T*>(a.pos_emb) + i * a.D;
    } else {
      id = ids[j - 1];
      src = reinterpret_cast<const T*>(a.table) + id * a.D;
    }
    const bool masked = j > 0 && id == pos_id;
    const float logit = logits[j], rn = rnorm[j];
    const float p = __expf(logit - lse);
    const float dl = masked ? 0.f : c * (p - (j == 0 ? 1.f : 0.f));
    if (dl == 0.f) continue;
    float ef[VEC];
    load_vec<T>(src + gl * VEC, ef);
    // ehat . q = logit * T; the norm was clamped iff rn == 1 / eps (then the normalisation is a constant scale)
    const float edq = logit * a.temperature;
    const bool through_norm = a.l2_norm && rn < 1.0f / a.l2_eps;
    float de[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float eh = ef[k] * rn;
      dq[k] += dl * eh;
      de[k] = dl * rn * (through_norm ? qf[k] - eh * edq : qf[k]);
    }
    if (j == 0) {
      T* dst = reinterpret_cast<T*>(a.d_pos_emb) + i * a.D + gl * VEC;
      uint4 ov;
      T* oe = reinterpret_cast<T*>(&ov);
#pragma unroll
      for (int k = 0; k < VEC; ++k) oe[k] = Cvt<T>::from_f(de[k]);
      *reinterpret_cast<uint4*>(dst) = ov;
    } else {
      float* dst = a.d_table + id * a.D + gl * VEC;
#pragma unroll
      for (int k = 0; k < VEC; k += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "f"(de[k]), "f"(de[k + 1]), "f"(de[k + 2]),
                     "f"(de[k + 3]) : "memory");
    }
  }
  // dq: sum over the NPW lane groups, then the lanes of group 0 store their slice
#pragma unroll
  for (int k = 0; k < VEC; ++k)
    for (int o = G; o < 32; o <<= 1) dq[k] += __shfl_xor_sync(0xffffffffu, dq[k], o);
  if (g == 0) {
    T* dst = reinterpret_cast<T*>(a.d_q) + i * a.D + gl * VEC;
    uint4 ov;
    T* oe = reinterpret_cast<T*>(&ov);
#pragma unroll
    for (int k = 0; k < VEC; ++k) oe[k] = Cvt<T>::from_f(dq[k]);
    *reinterpret_cast<uint4*>(dst) = ov;
  }
}

static int check_shape(const SslArgs& a, int elem_bytes) {
  const int vec = 16 / elem_bytes;
  const int G = a.D / vec;
  if (a.D % vec != 0 || G < 1 || G > 32 || (G & (G - 1)) != 0) {
    set_error("sampled softmax: embedding dim %d must be %d * 2^k with 2^k <= 32 (one 128-bit load per lane)", a.D, vec);
    return HSTU_ERR_UNSUPPORTED;
  }
  return 0;
}

int sampled_softmax_fwd(const SslArgs& a, int dtype, cudaStream_t st) {
  if (a.N == 0) return 0;
  if (int e = check_shape(a, dtype_bytes(dtype))) return e;
  const unsigned blocks = (unsigned)((a.N + 7) / 8);
  switch (dtype) {
    case HSTU_F32: ssl_fwd_kernel<float><<<blocks, 256, 0, st>>>(a); break;
    case HSTU_BF16: ssl_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(a); break;
    case HSTU_F16: ssl_fwd_kernel<__half><<<blocks, 256, 0, st>>>(a); break;
    default: set_error("sampled softmax: bad dtype %d", dtype); return HSTU_ERR_INVALID_ARGUMENT;
  }
  HSTU_CUDA_OK(cudaGetLastError());
  return 0;
}

int sampled_softmax_bwd(const SslArgs& a, int dtype, cudaStream_t st) {
  if (a.N == 0) return 0;
  if (int e = check_shape(a, dtype_bytes(dtype))) return e;
  const unsigned blocks = (unsigned)((a.N + 7) / 8);
  switch (dtype) {
    case HSTU_F32: ssl_bwd_kernel<float><<<blocks, 256, 0, st>>>(a); break;
    case HSTU_BF16: ssl_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(a); break;
    case HSTU_F16: ssl_bwd_kernel<__half><<<blocks, 256, 0, st>>>(a); break;
    default: set_error("sampled softmax: bad dtype %d", dtype); return HSTU_ERR_INVALID_ARGUMENT;
  }
  HSTU_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace hstu
