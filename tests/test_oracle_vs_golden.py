"""CPU: pin the oracle (oracle/hstu_oracle.py) against golden vectors produced by the unmodified
reference eager path (tests/golden/make_golden.py).  Tolerances: fp32 restatement vs fp32 eager
differ only by summation order -> rel-L2 <= 2e-6; integer row routing is bit-exact."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, golden
from oracle import hstu_oracle as O

F32_TOL = 2e-6


def _attn_files():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "attn_*.pt")))


@pytest.mark.parametrize("fname", _attn_files())
def test_attention_fwd_bwd(fname):
    g = golden(fname)
    kw = dict(num_targets=g["num_targets"], max_attn_len=g["max_attn_len"],
              contextual_seq_len=g["contextual_seq_len"], min_full_attn_seq_len=g["min_full_attn_seq_len"])
    out = O.hstu_mha_fwd(g["max_seq_len"], g["alpha"], g["q"], g["k"], g["v"], g["seq_offsets"], **kw)
    dq, dk, dv = O.hstu_mha_bwd(g["max_seq_len"], g["alpha"], g["dout"], g["q"], g["k"], g["v"],
                                g["seq_offsets"], **kw)
    ref = g["ref_f32"]
    for name, a, r in (("out", out, ref["out"]), ("dq", dq, ref["dq"]), ("dk", dk, ref["dk"]),
                       ("dv", dv, ref["dv"])):
        assert O.rel_l2(a, r) <= F32_TOL, (fname, name, O.rel_l2(a, r))
    # the reference's own comparison (ops/tests/hstu_attention_test.py:152-163) against eager in the
    # native dtype: assert_close at dtype-default tolerances
    nat = g["ref_native"]
    dt = nat["out"].dtype
    torch.testing.assert_close(out.to(dt), nat["out"])


@pytest.mark.parametrize("fname", ["delta_plain.pt", "delta_ctx.pt"])
def test_delta_attention(fname):
    g = golden(fname)
    out = O.delta_hstu_mha_fwd(g["max_seq_len"], g["alpha"], g["delta_q"], g["k"], g["v"], g["seq_offsets"],
                               g["num_targets"], g["max_attn_len"], g["contextual_seq_len"])
    assert O.rel_l2(out, g["out"]) <= F32_TOL


@pytest.mark.parametrize("fname", ["layer_norm_37x64.pt", "layer_norm_50x200.pt"])
def test_layer_norm(fname):
    g = golden(fname)
    y, mean, rstd = O.layer_norm_fwd(g["x"], g["w"], g["b"], g["eps"])
    assert O.rel_l2(y, g["ln"]["y"]) <= F32_TOL
    dx, dw, db = O.layer_norm_bwd(g["dy"], g["x"], g["w"], mean, rstd)
    for a, r in ((dx, g["ln"]["dx"]), (dw, g["ln"]["dw"]), (db, g["ln"]["db"])):
        assert O.rel_l2(a, r) <= 5e-6
    assert O.rel_l2(O.swish_layer_norm_fwd(g["x"], g["w"], g["b"], g["eps"]), g["swish"]["y"]) <= F32_TOL


@pytest.mark.parametrize("fname", ["compute_output_ln_concat.pt", "compute_output_gn_concat.pt",
                                   "compute_output_ln_plain.pt"])
def test_compute_output(fname):
    g = golden(fname)
    ts = [g[k].clone().requires_grad_() for k in ("attn", "u", "x", "norm_weight", "norm_bias", "output_weight")]
    out = O.hstu_compute_output_fwd(ts[0], ts[1], ts[2], ts[3], ts[4], ts[5], g["eps"], False, g["concat_ux"],
                                    g["group_norm"], g["num_heads"], g["linear_dim"])
    assert O.rel_l2(out, g["out"]) <= F32_TOL
    out.backward(g["dout"])
    for t, r in zip(ts, g["grads"]):
        assert O.rel_l2(t.grad, r) <= 1e-5


@pytest.mark.parametrize("fname", ["stu_ln.pt", "stu_gn_ctx_window.pt"])
def test_stu_stack(fname):
    g = golden(fname)
    cfg = g["cfg"]
    x = g["x"].clone().requires_grad_()
    params = {k: v.clone().requires_grad_() for k, v in g["state_dict"].items()}
    h = x
    for layer in range(cfg["layers"]):
        p = {k.split(".")[-1]: v for k, v in params.items() if k.startswith(f"_stu_layers.{layer}.")}
        # hstu_mha_fwd is not differentiable (explicit loops are, through torch ops) -> autograd works
        h = O.stu_layer_fwd(h, g["x_offsets"], g["max_seq_len"], g["num_targets"], p, cfg["num_heads"],
                            cfg["attention_dim"], cfg["hidden_dim"], cfg["max_attn_len"],
                            cfg["contextual_seq_len"], None, cfg["use_group_norm"])
    assert O.rel_l2(h, g["y"]) <= 5e-6


def test_jagged_routing_bit_exact():
    g = golden("jagged.pt")
    cat = O.concat_2D_jagged(g["vl"], g["vr"], g["max_l"], g["max_r"], g["ol"], g["orr"])
    assert torch.equal(cat, g["cat_jj"])
    cat_jd = O.concat_2D_jagged(g["vl"], g["dense_r"], g["max_l"], g["max_r"], g["ol"], None)
    assert torch.equal(cat_jd, g["cat_jd"])
    l, r = O.split_2D_jagged(g["cat_jj"], g["max_l"], g["max_r"], g["ol"], g["orr"])
    assert torch.equal(l, g["sp_l"]) and torch.equal(r, g["sp_r"])
    l, r = O.split_2D_jagged(g["cat_jd"], g["max_l"], g["max_r"], g["ol"], None)
    assert torch.equal(l, g["sp_dl"]) and torch.equal(r, g["sp_dr"])
    l2 = O.concat_2D_jagged(g["vl"], g["vr"], g["max_l"], g["max_r"], g["ol"], g["orr"], n_prefix_from_right=g["ctx"])
    assert torch.equal(l2, g["l2cat"])
    pre, l2x = O.split_2D_jagged(g["l2cat"], None, None, g["ol"], g["orr"], n_prefix_to_right=g["ctx"])
    assert torch.equal(pre, g["l2_pre"]) and torch.equal(l2x, g["l2_l2"])


def test_research_rel_bias_attention():
    g = golden("research_attn.pt")
    H, dqk, dv = g["H"], g["dqk"], g["dv"]
    out = O.hstu_rel_bias_attention_fwd(g["n"], g["q"].view(-1, H, dqk), g["k"].view(-1, H, dqk),
                                        g["v"].view(-1, H, dv), g["seq_offsets"], g["pos_w"], g["ts_w"],
                                        g["timestamps"])
    assert O.rel_l2(out.reshape(-1, H * dv), g["out"]) <= F32_TOL


@pytest.mark.parametrize("name", ["plain", "concat_ua"])
def test_research_block(name):
    """The research block restated in the oracle vs the unmodified reference module, fwd + every grad."""
    g = golden(f"research_block_{name}.pt")
    sd = g["state_dict"]
    leaves = {k: sd[k].clone().requires_grad_() for k in sd}
    x = g["x"].clone().requires_grad_()
    y = O.research_block_fwd(x, g["seq_offsets"], g["timestamps"], leaves["_uvqk"], leaves["_o.weight"], leaves["_o.bias"],
                             leaves["_rel_attn_bias._pos_w"], leaves["_rel_attn_bias._ts_w"], g["n"], g["H"], g["dqk"], g["dv"],
                             g["concat_ua"], g["eps"])
    assert O.rel_l2(y, g["y"]) < 2e-6
    y.backward(g["dy"])
    assert O.rel_l2(x.grad, g["dx"]) < 2e-6
    for k, ref in g["grads"].items():
        assert O.rel_l2(leaves[k].grad, ref) < 5e-6, k


@pytest.mark.parametrize("fname", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "position_*.pt"))))
def test_position_embeddings_oracle_matches_reference_bit_exactly(fname):
    """ops/position.py:43-96 through the eager path: integer routing (position index, time bucket) and all three roundings of the
    16-bit path are restated exactly, so out / dx are bit-identical and the fp32 scatter-adds agree to the last few ulps."""
    g = golden(fname)
    out, dx, dpos, dts = O.add_timestamp_positional_embeddings(
        g["alpha"], g["max_contextual_seq_len"], g["pos_w"], g["ts_w"], g["seq_offsets"], g["seq_lengths"], g["x"], g["timestamps"],
        g["num_targets"], g["interleave_targets"], g["time_bucket_fn"], g["dout"])
    assert torch.equal(out, g["out"]) and torch.equal(dx, g["dx"])
    torch.testing.assert_close(dpos, g["dpos_w"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(dts, g["dts_w"], rtol=1e-6, atol=1e-7)


def _ssl_jagged(g):
    N = g["supervision_ids"].shape[1]
    keep = torch.arange(N).unsqueeze(0) < g["lengths"].unsqueeze(1)
    return keep, g["output_embeddings"][keep], g["supervision_ids"][keep], g["supervision_embeddings"][keep], g["weights"][keep]


@pytest.mark.parametrize("fname", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "ssl_*.pt"))))
def test_sampled_softmax_oracle_matches_reference(fname):
    """SampledSoftmaxLoss.forward of the unmodified reference (LocalNegativesSampler + DotProductSimilarity) vs the restatement:
    loss and the gradients w.r.t. the output embeddings, the supervision embeddings and the item table."""
    g = golden(fname)
    keep, q, ids, pe, w = _ssl_jagged(g)
    q, pe, tb = q.float().requires_grad_(), pe.float().requires_grad_(), g["table"].float().requires_grad_()
    loss = O.sampled_softmax_loss(q, ids, pe, w, g["sampled_ids"], tb, g["temperature"], g["l2_norm"], g["l2_norm_eps"])
    loss.backward()
    tol = 2e-6 if g["table"].dtype == torch.float32 else 2e-2  # the bf16 fixture is the reference evaluated in bf16
    assert abs(float(loss) - float(g["loss"])) <= tol * abs(float(g["loss"]))
    dq = torch.zeros_like(g["d_out"].float())
    dq[keep] = q.grad
    dp = torch.zeros_like(g["d_sup"].float())
    dp[keep] = pe.grad
    assert O.rel_l2(dq, g["d_out"].float()) <= tol
    assert O.rel_l2(dp, g["d_sup"].float()) <= tol
    assert O.rel_l2(tb.grad, g["d_table"].float()) <= tol


def test_local_negatives_sampler_draws_the_reference_ids():
    """Same generator call as autoregressive_losses.py:106-121: under the saved RNG state the sampler reproduces the ids the
    reference drew when the fixture was made."""
    from generative_recommenders_b200.modules.sampled_softmax import LocalNegativesSampler

    g = golden("ssl_l2_f32.pt")
    _, _, ids, _, _ = _ssl_jagged(g)
    sampler = LocalNegativesSampler(g["V"], torch.nn.Embedding(g["V"], 4), list(range(g["V"])), True, 1e-6)
    state = torch.get_rng_state()
    torch.set_rng_state(g["rng_state"])
    drawn = sampler.sample_ids(ids, g["R"])
    torch.set_rng_state(state)
    assert torch.equal(drawn, g["sampled_ids"])


@pytest.mark.parametrize("fname", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "jagged_bmm_*.pt"))))
def test_jagged_dense_bmm_oracle_matches_reference(fname):
    g = golden(fname)
    j, d, b = (g[n].clone().requires_grad_() for n in ("jagged", "dense", "bias"))
    out = O.jagged_dense_bmm_broadcast_add(g["max_seq_len"], g["seq_offsets"], j, d, b)
    out.backward(g["dout"])
    torch.testing.assert_close(out, g["out"])
    for a, r in ((j.grad, g["d_jagged"]), (d.grad, g["d_dense"]), (b.grad, g["d_bias"])):
        torch.testing.assert_close(a, r)
