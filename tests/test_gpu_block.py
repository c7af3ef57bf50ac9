"""GPU parity of the rest of the HSTU block: layer norms, the fused output stage, jagged concat/split (bit-exact),
the fused preprocess+attention op and the STU layer/stack forward + backward, against the golden vectors of the
reference eager path (tests/golden/*.pt) and the CPU oracle."""
import pytest
import torch

from conftest import golden
from oracle import hstu_oracle as O
from util import assert_rel, offsets_from

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _k():
    from generative_recommenders_b200.common import HammerKernel

    return HammerKernel.CUDA


@pytest.mark.parametrize("fname", ["layer_norm_37x64.pt", "layer_norm_50x200.pt"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layer_norm_golden(fname, dtype):
    from generative_recommenders_b200.ops.layer_norm import layer_norm, swish_layer_norm

    g = golden(fname)
    for key, fn in (("ln", layer_norm), ("swish", swish_layer_norm)):
        x = g["x"].to(DEV, dtype).requires_grad_()
        w = g["w"].to(DEV).requires_grad_()
        b = g["b"].to(DEV).requires_grad_()
        y = fn(x, w, b, eps=g["eps"], kernel=_k())
        y.backward(g["dy"].to(DEV, dtype))
        if dtype == torch.float32:
            ref = g[key]
            assert_rel(y, ref["y"], f"{key} y")
            assert_rel(x.grad, ref["dx"], f"{key} dx")
            assert_rel(w.grad, ref["dw"], f"{key} dw")
            assert_rel(b.grad, ref["db"], f"{key} db")
        else:  # bf16 activations: compare with the oracle evaluated in fp32 on the bf16-valued inputs
            xq = g["x"].to(dtype).float()
            wq, bq = g["w"].to(dtype).float(), g["b"].to(dtype).float()
            yr, mean, rstd = O.layer_norm_fwd(xq, wq, bq, g["eps"])
            if key == "swish":
                yr = xq * torch.sigmoid(yr)
            assert_rel(y, yr, f"{key} y bf16")
            if key == "ln":
                dxr, dwr, dbr = O.layer_norm_bwd(g["dy"].to(dtype).float(), xq, wq, mean, rstd)
                assert_rel(x.grad, dxr, "ln dx bf16")
                assert_rel(w.grad.float(), dwr, "ln dw bf16", tol=1e-3)


@pytest.mark.parametrize("n,d", [(0, 64), (1, 32), (1000, 512), (777, 50), (64, 1024), (4099, 256)])
def test_layer_norm_shapes(n, d):
    """N in [0, ...], D in [32, 512] (+ odd and max D), like ops/tests/layer_norm_test.py:62-142."""
    from generative_recommenders_b200.ops.layer_norm import RMSNorm, layer_norm

    torch.manual_seed(n + d)
    x = torch.randn(n, d, device=DEV).requires_grad_()
    w = (torch.randn(d, device=DEV) * 0.3 + 1).requires_grad_()
    b = (torch.randn(d, device=DEV) * 0.1).requires_grad_()
    y = layer_norm(x, w, b, eps=1e-5, kernel=_k())
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, wr, br = x.detach().cpu(), w.detach().cpu(), b.detach().cpu()
    yr, mean, rstd = O.layer_norm_fwd(xr, wr, br, 1e-5)
    dxr, dwr, dbr = O.layer_norm_bwd(dy.cpu(), xr, wr, mean, rstd)
    if n > 0:
        assert_rel(y, yr, "y")
        assert_rel(x.grad, dxr, "dx")
        assert_rel(w.grad, dwr, "dw", tol=1e-4)
        assert_rel(b.grad, dbr, "db", tol=1e-4)
    else:
        assert y.shape == (0, d) and float(w.grad.abs().sum()) == 0.0
    rms = RMSNorm(d).to(DEV)
    x2 = torch.randn(max(n, 1), d, device=DEV, requires_grad=True)
    y2 = rms(x2)
    y2.sum().backward()
    assert_rel(y2, O.rms_norm_fwd(x2.detach().cpu(), rms.weight.detach().cpu(), 1e-5), "rms y")
    xr2 = x2.detach().cpu().requires_grad_()
    wr2 = rms.weight.detach().cpu().requires_grad_()
    O.rms_norm_fwd(xr2, wr2, 1e-5).sum().backward()
    assert_rel(x2.grad, xr2.grad, "rms dx", tol=1e-4)
    assert_rel(rms.weight.grad, wr2.grad, "rms dw", tol=1e-4)


@pytest.mark.parametrize("fname", ["compute_output_ln_concat.pt", "compute_output_gn_concat.pt", "compute_output_ln_plain.pt"])
@pytest.mark.parametrize("recompute", [False, True])
def test_compute_output_golden(fname, recompute):
    from generative_recommenders_b200.ops.hstu_compute import hstu_compute_output

    g = golden(fname)
    ts = [g[k].to(DEV).requires_grad_() for k in ("attn", "u", "x", "norm_weight", "norm_bias", "output_weight")]
    out = hstu_compute_output(attn=ts[0], u=ts[1], x=ts[2], norm_weight=ts[3], norm_bias=ts[4], norm_eps=g["eps"],
                              output_weight=ts[5], num_heads=g["num_heads"], linear_dim=g["linear_dim"], dropout_ratio=0.0,
                              training=True, concat_ux=g["concat_ux"], group_norm=g["group_norm"],
                              recompute_y_in_backward=recompute, kernel=_k())
    out.backward(g["dout"].to(DEV))
    assert_rel(out, g["out"], "out", tol=1e-4)  # TF32 is off by default for torch.addmm in fp32
    for t, r, nm in zip(ts, g["grads"], ("dattn", "du", "dx", "dnw", "dnb", "dwo")):
        assert_rel(t.grad, r, nm, tol=1e-4)


def test_dropout_statistics_and_backward_mask_consistency():
    """p > 0 cannot be compared bit-wise with eager (different generators); check keep-rate,
    scaling, and that backward applies the very same mask as forward."""
    from generative_recommenders_b200.ops.hstu_compute import hstu_compute_output

    torch.manual_seed(0)
    L, H, dv, D, p = 4096, 4, 32, 64, 0.25
    attn = torch.randn(L, H * dv, device=DEV)
    u = torch.randn(L, H * dv, device=DEV, requires_grad=True)
    x = torch.zeros(L, D, device=DEV)
    nw, nb = torch.ones(H * dv, device=DEV), torch.zeros(H * dv, device=DEV)
    wo = torch.zeros(3 * H * dv, D, device=DEV)
    wo[: H * dv, :] = 0.0
    # identity-like readout of the first D columns of the `u` part: out[:, j] = dropout(u)[:, j]
    wo[torch.arange(D), torch.arange(D)] = 1.0
    out = hstu_compute_output(attn=attn, u=u, x=x, norm_weight=nw, norm_bias=nb, norm_eps=1e-6, output_weight=wo,
                              num_heads=H, linear_dim=dv, dropout_ratio=p, training=True, concat_ux=True,
                              group_norm=False, recompute_y_in_backward=True, kernel=_k())
    kept = out != 0
    rate = kept.float().mean().item()
    assert abs(rate - (1 - p)) < 0.01, rate
    torch.testing.assert_close(out[kept], (u.detach()[:, :D] / (1 - p))[kept], rtol=1e-5, atol=1e-6)
    out.sum().backward()
    g = u.grad[:, :D]
    torch.testing.assert_close(g[kept], torch.full_like(g[kept], 1 / (1 - p)), rtol=1e-5, atol=1e-6)
    assert float(g[~kept].abs().max()) < 1e-6


def test_jagged_concat_split_bit_exact():
    from generative_recommenders_b200.ops import jagged_tensors as J

    g = golden("jagged.pt")
    c = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in g.items()}
    n = c["max_l"] + c["max_r"]
    for dt in (torch.float32, torch.bfloat16):
        vl, vr, dr = c["vl"].to(dt), c["vr"].to(dt), c["dense_r"].to(dt)
        assert torch.equal(J.concat_2D_jagged(n, vl, vr, c["max_l"], c["max_r"], c["ol"], c["orr"], kernel=_k()), c["cat_jj"].to(dt))
        assert torch.equal(J.concat_2D_jagged(n, vl, dr, c["max_l"], c["max_r"], c["ol"], None, kernel=_k()), c["cat_jd"].to(dt))
        l, r = J.split_2D_jagged(n, c["cat_jj"].to(dt), None, None, c["max_l"], c["max_r"], c["ol"], c["orr"], kernel=_k())
        assert torch.equal(l, c["sp_l"].to(dt)) and torch.equal(r, c["sp_r"].to(dt))
        l, r = J.split_2D_jagged(n, c["cat_jd"].to(dt), None, None, c["max_l"], c["max_r"], c["ol"], None, kernel=_k())
        assert torch.equal(l, c["sp_dl"].to(dt)) and torch.equal(r, c["sp_dr"].to(dt))
        l2 = J.hstu_concat_l2_embeddings(c["max_l"], vl, c["ol"], c["max_r"], vr, c["orr"], c["ctx"], kernel=_k())
        assert torch.equal(l2, c["l2cat"].to(dt))
        pre, l2x = J.hstu_split_l2_embeddings(n, c["l2cat"].to(dt), c["ol"], c["orr"], c["ctx"], kernel=_k())
        assert torch.equal(pre, c["l2_pre"].to(dt)) and torch.equal(l2x, c["l2_l2"].to(dt))
    # int32 offsets, odd row width (no 16-byte vector path), autograd round trip
    vl = c["vl"][:, :7].contiguous().requires_grad_()
    vr = c["vr"][:, :7].contiguous().requires_grad_()
    out = J.concat_2D_jagged(n, vl, vr, c["max_l"], c["max_r"], c["ol"].int(), c["orr"].int(), kernel=_k())
    assert torch.equal(out, c["cat_jj"][:, :7])
    out.backward(out.detach())
    assert torch.equal(vl.grad, vl.detach()) and torch.equal(vr.grad, vr.detach())


def test_jagged_large():
    """130 x 32768 x 512-wide rows: the reference's large-tensor case (ops/tests/jagged_tensors_test.py:189-321), scaled to
    keep the test fast: offsets beyond 2^31 bytes are exercised."""
    from generative_recommenders_b200.ops import jagged_tensors as J

    torch.manual_seed(1)
    B, D = 130, 512
    ll = torch.randint(0, 20000, (B,))
    lr = torch.randint(0, 12768, (B,))
    ol, orr = offsets_from(ll.tolist(), DEV), offsets_from(lr.tolist(), DEV)
    vl = torch.randn(int(ol[-1]), D, device=DEV, dtype=torch.bfloat16)
    vr = torch.randn(int(orr[-1]), D, device=DEV, dtype=torch.bfloat16)
    out = J.concat_2D_jagged(32768, vl, vr, 20000, 12768, ol, orr, kernel=_k())
    l, r = J.split_2D_jagged(32768, out, None, None, 20000, 12768, ol, orr, kernel=_k())
    assert torch.equal(l, vl) and torch.equal(r, vr)
    b = B - 1
    s = int(ol[b] + orr[b])
    assert torch.equal(out[s : s + int(ll[b])], vl[int(ol[b]) :]) and torch.equal(out[s + int(ll[b]) :], vr[int(orr[b]) :])


def _build_stack(g, dtype=torch.float32, **over):
    from generative_recommenders_b200.modules.stu import STULayer, STULayerConfig, STUStack

    cfg = g["cfg"]
    layers = [
        STULayer(STULayerConfig(embedding_dim=cfg["embedding_dim"], num_heads=cfg["num_heads"], hidden_dim=cfg["hidden_dim"],
                                attention_dim=cfg["attention_dim"], output_dropout_ratio=0.0, causal=True, target_aware=True,
                                max_attn_len=cfg["max_attn_len"] or None, use_group_norm=cfg["use_group_norm"],
                                contextual_seq_len=cfg["contextual_seq_len"],
                                recompute_normed_x=over.get("recompute", False), recompute_uvqk=over.get("recompute", False),
                                recompute_y=over.get("recompute", False)))
        for _ in range(cfg["layers"])
    ]
    stack = STUStack(layers)
    stack.load_state_dict(g["state_dict"])  # reference parameter names load unchanged
    return stack.to(DEV)


@pytest.mark.parametrize("fname", ["stu_ln.pt", "stu_gn_ctx_window.pt"])
@pytest.mark.parametrize("recompute", [False, True])
def test_stu_stack_golden(fname, recompute):
    """2-layer STUStack fwd + x.grad + every parameter gradient vs the reference eager run (modules/tests/stu_test.py:47-172)."""
    g = golden(fname)
    stack = _build_stack(g, recompute=recompute)
    x = g["x"].to(DEV).requires_grad_()
    y = stack(x=x, x_lengths=g["x_lengths"].to(DEV), x_offsets=g["x_offsets"].to(DEV), max_seq_len=g["max_seq_len"],
              num_targets=g["num_targets"].to(DEV))
    y.backward(g["dout"].to(DEV))
    assert_rel(y, g["y"], "y", tol=1e-4)
    assert_rel(x.grad, g["dx"], "dx", tol=1e-4)
    for n, p in stack.named_parameters():
        assert_rel(p.grad, g["param_grads"][n], n, tol=2e-4)


def test_stu_stack_bf16_vs_oracle():
    """bf16 activations (the training dtype of configs 2-4): forward of a 2-layer stack vs the fp32 oracle on bf16-valued
    parameters; every GEMM/activation is rounded to bf16 between ops, so the budget is a few bf16 ulps."""
    g = golden("stu_ln.pt")
    stack = _build_stack(g).to(torch.bfloat16)
    x = g["x"].to(DEV, torch.bfloat16)
    y = stack(x=x, x_lengths=g["x_lengths"].to(DEV), x_offsets=g["x_offsets"].to(DEV), max_seq_len=g["max_seq_len"],
              num_targets=g["num_targets"].to(DEV))
    cfg = g["cfg"]
    h = x.float().cpu()
    sd = {k: v.to(torch.bfloat16).float() for k, v in g["state_dict"].items()}
    for layer in range(cfg["layers"]):
        p = {k.split(".")[-1]: v for k, v in sd.items() if k.startswith(f"_stu_layers.{layer}.")}
        h = O.stu_layer_fwd(h, g["x_offsets"], g["max_seq_len"], g["num_targets"], p, cfg["num_heads"],
                            cfg["attention_dim"], cfg["hidden_dim"])
    assert_rel(y, h, "stu bf16 y", tol=1.5e-2)


def test_cached_forward_matches_full_forward():
    """Prefill + cached_forward on the last `delta` rows == full forward on those rows (modules/tests/stu_test.py:327-457)."""
    from generative_recommenders_b200.modules.stu import STULayer, STULayerConfig, STUStack

    torch.manual_seed(3)
    D, H, dqk, dv, delta = 32, 2, 16, 16, 5
    stack = STUStack([STULayer(STULayerConfig(embedding_dim=D, num_heads=H, hidden_dim=dv, attention_dim=dqk,
                                              output_dropout_ratio=0.0, target_aware=True)) for _ in range(2)]).to(DEV)
    stack.eval()
    lengths = torch.tensor([40, 23, 64], device=DEV)
    full_len = lengths + delta
    nt = torch.full((3,), delta, device=DEV)
    off = offsets_from(full_len.tolist(), DEV)
    N = 64 + delta
    x = torch.randn(int(off[-1]), D, device=DEV)
    with torch.no_grad():
        y_full = stack(x=x, x_lengths=full_len, x_offsets=off, max_seq_len=N, num_targets=nt)
        # prefill on the prefix (cache = everything but the last delta rows), then the delta step
        for layer in stack._stu_layers:
            layer.reset_kv_cache()
        stack(x=x, x_lengths=full_len, x_offsets=off, max_seq_len=N, num_targets=nt, max_kv_caching_len=64,
              kv_caching_lengths=lengths)
        rows = torch.cat([torch.arange(int(off[i + 1]) - delta, int(off[i + 1]), device=DEV) for i in range(3)])
        y_delta = stack.cached_forward(delta_x=x[rows], num_targets=nt)
    assert_rel(y_delta, y_full[rows], "cached vs full", tol=1e-4)
