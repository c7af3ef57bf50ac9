"""GPU parity of jagged_dense_bmm_broadcast_add through the C ABI against golden vectors of
the reference eager path (ops/jagged_tensors.py:210-253, ops/pytorch/pt_jagged.py:77-98; recipe of
ops/tests/jagged_tensors_test.py) and against the oracle at the reference's "large tensor" proportions."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, golden
from oracle import hstu_oracle as O
from util import offsets_from

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


def _run(g):
    from generative_recommenders_b200.common import HammerKernel
    from generative_recommenders_b200.ops.jagged_tensors import jagged_dense_bmm_broadcast_add

    j, d, b = (g[n].to(DEV).requires_grad_() for n in ("jagged", "dense", "bias"))
    out = jagged_dense_bmm_broadcast_add(max_seq_len=g["max_seq_len"], seq_offsets=g["seq_offsets"].to(DEV), jagged=j, dense=d,
                                         bias=b, kernel=HammerKernel.CUDA)
    out.backward(g["dout"].to(DEV))
    return out.detach().cpu(), j.grad.cpu(), d.grad.cpu(), b.grad.cpu()


@pytest.mark.parametrize("fname", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "jagged_bmm_*.pt"))))
def test_jagged_dense_bmm_golden(fname):
    g = golden(fname)
    out, dj, dd, db = _run(g)
    # the reference's own criterion (jagged_tensors_test.py: torch.testing.assert_close with the default tolerances of the dtype)
    for name, a, r in (("out", out, g["out"]), ("d_jagged", dj, g["d_jagged"]), ("d_dense", dd, g["d_dense"]), ("d_bias", db, g["d_bias"])):
        torch.testing.assert_close(a, r, msg=lambda m, n=name: f"{fname}:{n}: {m}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_jagged_dense_bmm_large_vs_oracle(dtype):
    gen = torch.Generator().manual_seed(12)
    B, max_len, K, N = 24, 700, 256, 192
    lengths = torch.randint(max_len + 1, (B,), generator=gen)
    off = offsets_from(lengths.tolist())
    L = int(off[-1])
    g = dict(max_seq_len=max_len, seq_offsets=off, jagged=torch.empty(L, K).uniform_(-1, 1, generator=gen).to(dtype),
             dense=torch.empty(B, K, N).uniform_(-0.1, 0.1, generator=gen).to(dtype),
             bias=torch.empty(B, N).uniform_(-1, 1, generator=gen).to(dtype), dout=(torch.randn(L, N, generator=gen) * 0.1).to(dtype))
    out, dj, dd, db = _run(g)
    j, d, b = (g[n].float().requires_grad_() for n in ("jagged", "dense", "bias"))
    ref = O.jagged_dense_bmm_broadcast_add(max_len, off, j, d, b)
    ref.backward(g["dout"].float())
    tol = 2e-5 if dtype == torch.float32 else 4e-3  # bf16: one storage rounding of the result (fp32 accumulation inside)
    for name, a, r in (("out", out, ref.detach()), ("d_jagged", dj, j.grad), ("d_dense", dd, d.grad), ("d_bias", db, b.grad)):
        assert O.rel_l2(a.float(), r) <= tol, name
