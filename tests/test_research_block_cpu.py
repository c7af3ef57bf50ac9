"""Host-side checks of the research-path block that need no GPU: the reference's state dict loads unchanged, unsupported
options fail loudly, and there is no CPU fallback."""
import pytest
import torch

from conftest import golden


def _mk(**kw):
    from generative_recommenders_b200.modules.research_hstu import (
        RelativeBucketedTimeAndPositionBasedBias,
        SequentialTransductionUnitJagged,
    )

    args = dict(embedding_dim=32, linear_hidden_dim=16, attention_dim=16, dropout_ratio=0.0, attn_dropout_ratio=0.0, num_heads=2,
                linear_activation="silu", relative_attention_bias_module=RelativeBucketedTimeAndPositionBasedBias(24, 128),
                normalization="rel_bias", linear_config="uvqk", concat_ua=False, epsilon=1e-6, max_length=24)
    args.update(kw)
    return SequentialTransductionUnitJagged(**args)


@pytest.mark.parametrize("name", ["plain", "concat_ua"])
def test_reference_state_dict_loads_strictly(name):
    g = golden(f"research_block_{name}.pt")
    from generative_recommenders_b200.modules.research_hstu import RelativeBucketedTimeAndPositionBasedBias

    blk = _mk(embedding_dim=g["D"], linear_hidden_dim=g["dv"], attention_dim=g["dqk"], num_heads=g["H"], concat_ua=g["concat_ua"],
              relative_attention_bias_module=RelativeBucketedTimeAndPositionBasedBias(g["n"], 128), max_length=g["n"])
    missing, unexpected = blk.load_state_dict(g["state_dict"], strict=True)
    assert not missing and not unexpected
    assert set(blk.state_dict().keys()) == set(g["state_dict"].keys())
    for k, v in g["state_dict"].items():
        assert torch.equal(blk.state_dict()[k], v)


def test_unsupported_options_raise():
    with pytest.raises(ValueError):
        _mk(linear_config="uv")
    with pytest.raises(NotImplementedError):
        _mk(normalization="softmax_rel_bias")
    with pytest.raises(NotImplementedError):
        _mk(linear_activation="none")
    with pytest.raises(ValueError):
        _mk(relative_attention_bias_module=None)
    with pytest.raises(NotImplementedError):
        _mk(attn_dropout_ratio=0.1)


def test_no_cpu_fallback():
    blk = _mk()
    x = torch.randn(10, 32)
    off = torch.tensor([0, 4, 10])
    with pytest.raises(RuntimeError):
        blk(x, off, None, torch.tril(torch.ones(24, 24)))
    with pytest.raises(RuntimeError):  # a delta call needs a cache that carries padded q / k
        blk(x, off, None, torch.tril(torch.ones(24, 24)), delta_x_offsets=(off[:2], off[:2]), cache=(x, None, None, x))


def test_non_causal_mask_and_wrong_sizes_are_rejected():
    """The kernel applies the reference's causal mask and never reads `invalid_attn_mask`: anything else must not be silently
    ignored; a mask whose size disagrees with the bias module would index pos_w / timestamps out of bounds."""
    blk = _mk()
    x = torch.randn(10, 32)
    off = torch.tensor([0, 4, 10])
    with pytest.raises(NotImplementedError):
        blk(x, off, None, torch.ones(24, 24))                      # full (non-causal) mask
    with pytest.raises(RuntimeError):
        blk(x, off, None, torch.tril(torch.ones(20, 20)))          # n != the bias module's max_seq_len


def test_bias_shape_contract():
    from generative_recommenders_b200 import _lib
    from generative_recommenders_b200.ops.hstu_attention import _fill_bias

    p = _lib.AttnParams()
    p.max_seq_len, p.batch = 24, 2
    with pytest.raises(RuntimeError):
        _fill_bias(p, (torch.zeros(40), None, None), None)                              # pos_w must be 2n-1 = 47
    with pytest.raises(RuntimeError):
        _fill_bias(p, (torch.zeros(47), torch.zeros(129), torch.zeros(2, 23, dtype=torch.int64)), None)  # timestamps [B, n]
    with pytest.raises(RuntimeError):
        _fill_bias(p, (torch.zeros(47), torch.zeros(129), None), None)
