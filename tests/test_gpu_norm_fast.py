"""The 16-bit / length-256 specialisations of the row-wise kernels (csrc/norm_fast.cu) against the oracle, through the C ABI.

They replace the general kernels of csrc/norm.cu for the shape every bench configuration uses (D = H dv = 256, bf16), so they
get their own parity cases: LayerNorm forward / backward, the output stage (u * LN(attn) with the three concat modes, silu(u) on
and off) forward / backward, odd row counts (the two-rows-per-warp loop has a tail), strided inputs, and the dropout masks of
the fast forward against those of the general backward (both must evaluate the same counter-based function).
"""
import pytest
import torch

from oracle import hstu_oracle as O
from util import assert_rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from generative_recommenders_b200.ops import hstu_compute as hc
    from generative_recommenders_b200.ops import layer_norm as ln
    return hc, ln


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n", [1, 7, 16, 1237, 20011])
def test_layer_norm_256_fwd_bwd(dtype, n):
    hc, ln = _ops()
    torch.manual_seed(n)
    D = 256
    x = (torch.randn(n, D, device=DEV) * 1.7 + 0.3).to(dtype)
    w = (1 + 0.2 * torch.randn(D, device=DEV)).to(dtype)
    b = (0.1 * torch.randn(D, device=DEV)).to(dtype)
    dy = torch.randn(n, D, device=DEV).to(dtype)
    y, mean, rstd = ln.cuda_layer_norm_fwd(x, w, b, 1e-6, False)
    dx, dw, db = ln.cuda_layer_norm_bwd(dy, x, w, b, mean, rstd, False)
    xr, wr, br = x.float().cpu(), w.float().cpu(), b.float().cpu()
    yr, mr, rr = O.layer_norm_fwd(xr, wr, br, 1e-6)
    dxr, dwr, dbr = O.layer_norm_bwd(dy.float().cpu(), xr, wr, mr, rr)
    assert_rel(y, yr, "y")
    assert_rel(mean, mr, "mean", tol=2e-5)
    assert_rel(rstd, rr, "rstd", tol=2e-5)
    assert_rel(dx, dxr, "dx")
    assert_rel(dw, dwr, "dw", tol=1e-4)
    assert_rel(db, dbr, "db", tol=1e-4)


def test_layer_norm_256_strided_rows_take_the_fast_path_too():
    hc, ln = _ops()
    torch.manual_seed(3)
    big = torch.randn(513, 1024, device=DEV).to(torch.bfloat16)
    x = big[:, 256:512]  # row stride 1024 elements, 16-byte aligned
    w = torch.ones(256, device=DEV, dtype=torch.bfloat16)
    b = torch.zeros(256, device=DEV, dtype=torch.bfloat16)
    y, _, _ = ln.cuda_layer_norm_fwd(x, w, b, 1e-5, False)
    assert_rel(y, O.layer_norm_fwd(x.float().cpu(), w.float().cpu(), b.float().cpu(), 1e-5)[0], "y strided")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("concat", [0, 1, 2])
@pytest.mark.parametrize("silu_u", [False, True])
def test_output_stage_256_fwd_bwd(dtype, concat, silu_u):
    hc, ln = _ops()
    torch.manual_seed(10 * concat + int(silu_u))
    n, H, dv = 3001, 8, 32
    W = H * dv
    attn = (torch.randn(n, W, device=DEV) * 0.8).to(dtype)
    u = torch.randn(n, W, device=DEV).to(dtype)
    w = (1 + 0.2 * torch.randn(W, device=DEV)).to(dtype)
    b = (0.1 * torch.randn(W, device=DEV)).to(dtype)
    out, mean, rstd = hc.cuda_norm_mul_dropout_fwd(attn, u, w, b, 1e-6, 0.0, 0, silu_u, concat, False, H, dv)
    dout = torch.randn_like(out)
    dattn, du, dw, db = hc.cuda_norm_mul_dropout_bwd(dout, attn, u, w, b, mean, rstd, 0.0, 0, silu_u, concat, False, H, dv)

    ar, ur, wr, br = (t.float().cpu().requires_grad_() for t in (attn, u, w, b))
    uf = torch.nn.functional.silu(ur) if silu_u else ur
    nrm = O.layer_norm_fwd(ar, wr, br, 1e-6)[0]
    yr = uf * nrm
    if concat == 1:
        yr = torch.cat([uf, ar, yr], dim=1)      # concat_ux (pt_hstu_linear.py:57-58)
    elif concat == 2:
        yr = torch.cat([uf, nrm, yr], dim=1)     # concat_ua of the research block (hstu.py:427-437)
    yr.backward(dout.float().cpu())
    assert_rel(out, yr.detach(), "out")
    assert_rel(dattn, ar.grad, "dattn")
    assert_rel(du, ur.grad, "du")
    assert_rel(dw, wr.grad, "dw", tol=1e-4)
    assert_rel(db, br.grad, "db", tol=1e-4)


def test_dropout_masks_of_the_fast_and_general_kernels_agree():
    """len 256 bf16 -> fast kernels; the same tensors viewed as 2 heads of 128 with group norm -> general kernels.  The
    counter-based dropout depends only on (seed, flat element index), so the zero patterns of the outputs must coincide; and the
    fast backward must zero exactly the gradient elements the fast forward dropped."""
    hc, ln = _ops()
    torch.manual_seed(5)
    n, W, p, seed = 2048, 256, 0.3, 987654321
    attn = (torch.randn(n, W, device=DEV) + 3.0).to(torch.bfloat16)   # bounded away from zero: a zero output is a dropped one
    u = (torch.rand(n, W, device=DEV) + 0.5).to(torch.bfloat16)
    w = torch.ones(W, device=DEV, dtype=torch.bfloat16)
    b = torch.full((W,), 4.0, device=DEV, dtype=torch.bfloat16)      # LN(attn) + 4 > 0
    out_f, mean, rstd = hc.cuda_norm_mul_dropout_fwd(attn, u, w, b, 1e-6, p, seed, False, 1, False, 8, 32)
    wg = torch.ones(2, device=DEV, dtype=torch.bfloat16)
    bg = torch.full((2,), 4.0, device=DEV, dtype=torch.bfloat16)
    out_g, _, _ = hc.cuda_norm_mul_dropout_fwd(attn, u, wg, bg, 1e-6, p, seed, False, 1, True, 2, 128)
    assert out_f.shape == out_g.shape == (n, 3 * W)
    assert torch.equal(out_f == 0, out_g == 0)
    keep = (out_f != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.01, keep
    dout = torch.ones_like(out_f)
    dattn, du, _, _ = hc.cuda_norm_mul_dropout_bwd(dout, attn, u, w, b, mean, rstd, p, seed, False, 1, False, 8, 32)
    # du = dropout(dout_u) + dropout(dout_y) * LN(attn): zero iff both the u part and the y part were dropped
    dropped_u = out_f[:, :W] == 0
    dropped_y = out_f[:, 2 * W:] == 0
    assert torch.equal(du == 0, dropped_u & dropped_y)
