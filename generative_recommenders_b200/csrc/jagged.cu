// Jagged row concat / split (pure copies; integer-exact routing).
//
//   concat:  out_b = [ right_b[:n_prefix] | left_b | right_b[n_prefix:] ]      split = the inverse
//
// Reference: ops/jagged_tensors.py:55-207 (facade), ops/pytorch/pt_jagged_tensors.py:31-246 (eager),
// ops/triton/triton_jagged_tensors.py:31-142 (grid (max_seq_len, B), one program per row).
// Here: one warp per output row, 128-bit loads/stores along the row when the layout allows; grid =
// (ceil(max_seq_len / 8), B).  A NULL offsets pointer means the side is dense with `dense_len` rows per entry.
#include "common.cuh"

namespace hstu {

template <int VB>  // bytes per lane access: 16, 4, 2, 1
struct VecT;
template <> struct VecT<16> { using type = uint4; };
template <> struct VecT<4> { using type = uint32_t; };
template <> struct VecT<2> { using type = uint16_t; };
template <> struct VecT<1> { using type = uint8_t; };

template <int VB, bool SPLIT>
__global__ void __launch_bounds__(256) jagged_rows_kernel(const char* __restrict__ a, const char* __restrict__ b,
                                                           char* __restrict__ c, char* __restrict__ c2,
                                                           const void* __restrict__ off_l, const void* __restrict__ off_r,
                                                           int is_i64, int dense_l, int dense_r, int n_prefix,
                                                           long long row_bytes) {
  // concat: a = left, b = right, c = out.   split: a = in, c = left, c2 = right.
  using V = typename VecT<VB>::type;
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pos = blockIdx.x * 8 + warp;
  long long l0, l1, r0, r1;
  if (off_l) {
    l0 = load_index(off_l, is_i64, bidx);
    l1 = load_index(off_l, is_i64, bidx + 1);
  } else {
    l0 = (long long)bidx * dense_l;
    l1 = l0 + dense_l;
  }
  if (off_r) {
    r0 = load_index(off_r, is_i64, bidx);
    r1 = load_index(off_r, is_i64, bidx + 1);
  } else {
    r0 = (long long)bidx * dense_r;
    r1 = r0 + dense_r;
  }
  const long long len_l = l1 - l0, len_r = r1 - r0;
  if (pos >= len_l + len_r) return;
  const long long npre = n_prefix < len_r ? n_prefix : len_r;
  const long long cat_row = l0 + r0 + pos;
  bool from_left;
  long long side_row;
  if (pos < npre) {
    from_left = false;
    side_row = r0 + pos;
  } else if (pos < npre + len_l) {
    from_left = true;
    side_row = l0 + (pos - npre);
  } else {
    from_left = false;
    side_row = r0 + (pos - len_l);
  }
  const V* src;
  V* dst;
  if (!SPLIT) {
    src = reinterpret_cast<const V*>((from_left ? a : b) + side_row * row_bytes);
    dst = reinterpret_cast<V*>(c + cat_row * row_bytes);
  } else {
    src = reinterpret_cast<const V*>(a + cat_row * row_bytes);
    dst = reinterpret_cast<V*>((from_left ? c : c2) + side_row * row_bytes);
  }
  const int nvec = (int)(row_bytes / VB);
  for (int i = lane; i < nvec; i += 32) dst[i] = src[i];
}

static int pick_vb(long long row_bytes, std::initializer_list<const void*> ptrs) {
  int vb = 16;
  auto ok = [&](int v) {
    if (row_bytes % v) return false;
    for (const void* p : ptrs)
      if (p && (reinterpret_cast<uintptr_t>(p) % v)) return false;
    return true;
  };
  while (vb > 1 && !ok(vb)) vb = vb == 16 ? 4 : vb / 2;
  return vb;
}

int jagged_concat_split(bool split, const void* a, const void* b, void* c, void* c2, const void* off_l, const void* off_r,
                        int is_i64, int batch, int dense_l, int dense_r, int n_prefix, int D, int elem_bytes,
                        int max_seq_len, cudaStream_t st) {
  if (batch <= 0 || max_seq_len <= 0 || D <= 0) return 0;
  const long long row_bytes = (long long)D * elem_bytes;
  const int vb = pick_vb(row_bytes, {a, b, c, c2});
  dim3 grid((max_seq_len + 7) / 8, batch);
#define LAUNCH(VB)                                                                                                 \
  do {                                                                                                             \
    if (split)                                                                                                     \
      jagged_rows_kernel<VB, true><<<grid, 256, 0, st>>>((const char*)a, nullptr, (char*)c, (char*)c2, off_l, off_r, \
                                                         is_i64, dense_l, dense_r, n_prefix, row_bytes);           \
    else                                                                                                           \
      jagged_rows_kernel<VB, false><<<grid, 256, 0, st>>>((const char*)a, (const char*)b, (char*)c, nullptr, off_l,  \
                                                          off_r, is_i64, dense_l, dense_r, n_prefix, row_bytes);   \
  } while (0)
  switch (vb) {
    case 16: LAUNCH(16); break;
    case 4: LAUNCH(4); break;
    case 2: LAUNCH(2); break;
    default: LAUNCH(1); break;
  }
#undef LAUNCH
  HSTU_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace hstu
