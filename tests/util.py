"""Shared helpers of the parity tests."""
import math

import torch

from oracle import hstu_oracle as O

# north-star tolerance: activations / grads within 1e-3 relative (L2) of the reference evaluated in fp32 on the same
# input values.  Outputs stored in bf16/fp16 additionally carry the unavoidable storage rounding of that dtype
# (measured on the reference tensor itself, ~1.7e-3 for bf16); the two are combined in quadrature and NOTHING else is
# granted: the tensor-core kernels keep their P / dS operands in fp16 (11-bit significand) precisely so that no operand
# rounding term is needed here.  fp32: 1e-3 flat would be far too loose for an fp32 kernel, so fp32 paths are held to 2e-5.
TOL = {torch.float32: 2e-5, torch.bfloat16: 1e-3, torch.float16: 1e-3}


def assert_rel(actual: torch.Tensor, ref32: torch.Tensor, what: str, tol: float = None) -> float:
    """rel-L2(actual, ref32) <= sqrt(tol^2 + q^2), q = storage rounding of actual.dtype measured on ref32."""
    dt = actual.dtype
    t = TOL[dt] if tol is None else tol
    q = O.storage_quantisation(ref32.float().cpu(), dt)
    err = O.rel_l2(actual.float().cpu(), ref32.float().cpu())
    lim = math.sqrt(t * t + q * q)
    assert err <= lim, f"{what}: rel-L2 error {err:.3e} > {lim:.3e} (tol {t:.1e}, storage rounding {q:.2e})"
    return err


def offsets_from(lengths, device="cpu", dtype=torch.int64):
    off = torch.zeros(len(lengths) + 1, dtype=dtype, device=device)
    off[1:] = torch.cumsum(torch.as_tensor(lengths, dtype=dtype, device=device), 0)
    return off
