// Jagged HSTU attention backward on tcgen05 -- dispatch glue (kernel: see below).
#include "common.cuh"
#include "internal.h"
#include "umma.cuh"

namespace hstu {

bool umma_fwd_supported(const hstu_attn_params& p);

bool umma_supported(const hstu_attn_params& p, bool bwd) {
  if (bwd) return false;  // tcgen05 backward not enabled yet: the generic kernels run the backward
  return umma_fwd_supported(p);
}

size_t umma_workspace_bytes(const hstu_attn_params& p, bool bwd) {
  (void)p;
  (void)bwd;
  return 0;
}

int attn_umma_bwd(const hstu_attn_params& p, cudaStream_t st) {
  (void)p;
  (void)st;
  set_error("tcgen05 backward is not available in this build");
  return HSTU_ERR_UNSUPPORTED;
}

}  // namespace hstu
