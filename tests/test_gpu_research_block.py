"""GPU parity of the research-path block against golden vectors produced by the UNMODIFIED
reference module `SequentialTransductionUnitJagged` (research/modeling/sequential/hstu.py:226-444; fixtures
tests/golden/research_block_*.pt written by tests/golden/make_golden.py): output, input gradient and every parameter
gradient, fp32 (tolerance: rel-L2 <= 2e-5, tests/util.py; the two bias tables accumulate many tiny terms: 1e-4)."""
import pytest
import torch

from conftest import golden
from util import assert_rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _build(g):
    from generative_recommenders_b200.modules.research_hstu import (
        RelativeBucketedTimeAndPositionBasedBias,
        SequentialTransductionUnitJagged,
    )

    blk = SequentialTransductionUnitJagged(
        embedding_dim=g["D"], linear_hidden_dim=g["dv"], attention_dim=g["dqk"], dropout_ratio=0.0, attn_dropout_ratio=0.0,
        num_heads=g["H"], linear_activation="silu",
        relative_attention_bias_module=RelativeBucketedTimeAndPositionBasedBias(max_seq_len=g["n"], num_buckets=128),
        normalization="rel_bias", linear_config="uvqk", concat_ua=g["concat_ua"], epsilon=g["eps"], max_length=g["n"],
    )
    blk.load_state_dict(g["state_dict"], strict=True)  # the reference's own state dict
    return blk.to(DEV)


@pytest.mark.parametrize("name", ["plain", "concat_ua"])
def test_research_block_golden(name):
    g = golden(f"research_block_{name}.pt")
    blk = _build(g)
    x = g["x"].to(DEV).requires_grad_()
    n = g["n"]
    y, cache = blk(x, g["seq_offsets"].to(DEV), g["timestamps"].to(DEV), torch.tril(torch.ones(n, n, device=DEV)))
    y.backward(g["dy"].to(DEV))
    assert_rel(y, g["y"], f"research block {name} y")
    assert_rel(x.grad, g["dx"], f"research block {name} dx")
    for k, ref in g["grads"].items():
        got = dict(blk.named_parameters())[k].grad
        assert got is not None, k
        assert_rel(got, ref, f"research block {name} d{k}", tol=1e-4 if "_rel_attn_bias" in k else None)
    assert cache[3] is y


def test_research_block_dropout_and_eval_are_consistent():
    """training with p > 0 drops ~p of the output-stage activations (statistically), eval is deterministic."""
    g = golden("research_block_plain.pt")
    blk = _build(g)
    blk._dropout_ratio = 0.5
    x = g["x"].to(DEV)
    n = g["n"]
    args = (x, g["seq_offsets"].to(DEV), g["timestamps"].to(DEV), torch.tril(torch.ones(n, n, device=DEV)))
    blk.eval()
    y0, _ = blk(*args)
    y1, _ = blk(*args)
    assert torch.equal(y0, y1)
    blk.train()
    yt, _ = blk(*args)
    assert torch.isfinite(yt).all() and not torch.equal(yt, y0)


def test_research_block_cached_delta_forward_golden():
    """Incremental decoding (hstu.py:284-444 with delta_x_offsets / cache): a full forward that returns its cache states, then the
    last row of every sequence is replaced and only those rows are recomputed against the cache.  Outputs and all four cache
    tensors against the unmodified reference module."""
    g = golden("research_block_cache.pt")
    g = dict(g, concat_ua=False, eps=1e-6)
    blk = _build(g).eval()
    n = g["n"]
    mask = torch.tril(torch.ones(n, n, device=DEV))
    off, ts = g["seq_offsets"].to(DEV), g["timestamps"].to(DEV)
    with torch.no_grad():
        y0, cache = blk(g["x"].to(DEV), off, ts, mask, return_cache_states=True)
        assert_rel(y0, g["y0"], "cached: first full forward")
        for name, a, r in zip(("v", "padded_q", "padded_k", "outputs"), cache, g["cache0"]):
            assert a.shape == r.shape, name
            assert_rel(a, r, f"cache state {name} after the full forward")
        delta = (g["delta_rows"].to(DEV), g["delta_pos"].to(DEV))
        y1, cache1 = blk(g["x2"].to(DEV), off, ts, mask, delta_x_offsets=delta, cache=cache)
    assert_rel(y1, g["y1"], "cached: delta forward output")
    for name, a, r in zip(("v", "padded_q", "padded_k", "outputs"), cache1, g["cache1"]):
        assert_rel(a, r, f"cache state {name} after the delta forward")
    # the recomputed rows equal what a full forward on the updated input gives for those rows
    assert_rel(y1[delta[0]], g["y_full"][g["delta_rows"]], "delta rows vs full forward")
