"""GPU parity of add_timestamp_positional_embeddings through the C ABI against golden vectors of the
reference eager path (ops/position.py:43-96; the recipe of ops/tests/position_test.py:96-233) and against the oracle at a
large size (the reference's own large case: B=130, D=512, max_uih_len=32768 is scaled to what the CPU oracle does in seconds)."""
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
    from generative_recommenders_b200.ops.position import add_timestamp_positional_embeddings

    pos_w = g["pos_w"].to(DEV).requires_grad_()
    ts_w = g["ts_w"].to(DEV).requires_grad_()
    x = g["x"].to(DEV).requires_grad_()
    nt = None if g["num_targets"] is None else g["num_targets"].to(DEV)
    out = add_timestamp_positional_embeddings(
        alpha=g["alpha"], max_seq_len=g["max_seq_len"], max_contextual_seq_len=g["max_contextual_seq_len"],
        position_embeddings_weight=pos_w, timestamp_embeddings_weight=ts_w, seq_offsets=g["seq_offsets"].to(DEV),
        seq_lengths=g["seq_lengths"].to(DEV), seq_embeddings=x, timestamps=g["timestamps"].to(DEV), num_targets=nt,
        interleave_targets=g["interleave_targets"], time_bucket_fn=g["time_bucket_fn"], kernel=HammerKernel.CUDA)
    out.backward(g["dout"].to(DEV))
    return out.detach().cpu(), x.grad.cpu(), pos_w.grad.cpu(), ts_w.grad.cpu()


@pytest.mark.parametrize("fname", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "position_*.pt"))))
def test_position_embeddings_golden(fname):
    g = golden(fname)
    out, dx, dpos, dts = _run(g)
    assert torch.equal(out, g["out"]), "out must be bit-identical (same roundings as the eager path)"
    assert torch.equal(dx, g["dx"])
    # fp32 scatter-adds: the summation order differs (atomics), nothing else
    torch.testing.assert_close(dpos, g["dpos_w"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dts, g["dts_w"], rtol=1e-5, atol=1e-6)


def test_position_embeddings_large_vs_oracle():
    gen = torch.Generator().manual_seed(5)
    B, D, max_uih, max_t, ctx = 64, 512, 4096, 10, 10
    nt = torch.randint(max_t + 1, (B,), generator=gen)
    lengths = torch.randint(int(0.7 * max_uih), max_uih + 1, (B,), generator=gen) + nt
    off = offsets_from(lengths.tolist())
    N = max_uih + max_t
    L = int(off[-1])
    g = dict(alpha=0.5, max_seq_len=N, max_contextual_seq_len=ctx, interleave_targets=False, time_bucket_fn="log",
             pos_w=torch.empty(N, D).uniform_(-1, 1, generator=gen), ts_w=torch.empty(600, D).uniform_(-1, 1, generator=gen),
             seq_offsets=off, seq_lengths=lengths, x=torch.empty(L, D).uniform_(-0.1, 0.1, generator=gen).to(torch.bfloat16),
             num_targets=nt, dout=(torch.randn(L, D, generator=gen) * 0.01).to(torch.bfloat16))
    ts_dense = torch.randint(86400, (B, N), generator=gen).cumsum(dim=1)
    g["timestamps"] = ts_dense[torch.arange(N) < lengths.unsqueeze(1)]
    out, dx, dpos, dts = _run(g)
    rout, rdx, rdpos, rdts = O.add_timestamp_positional_embeddings(
        g["alpha"], ctx, g["pos_w"], g["ts_w"], off, lengths, g["x"], g["timestamps"], nt, False, "log", g["dout"])
    assert torch.equal(out, rout) and torch.equal(dx, rdx)
    torch.testing.assert_close(dpos, rdpos, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dts, rdts, rtol=1e-4, atol=1e-4)  # a handful of buckets each sum ~1e5 rows
