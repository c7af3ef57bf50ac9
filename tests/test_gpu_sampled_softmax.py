"""GPU parity of the fused sampled-softmax loss through the C ABI against golden vectors of the
reference's SampledSoftmaxLoss (research/modeling/sequential/losses/sampled_softmax.py:29-193) and against the oracle at the
shape of BASELINE config 3 (Amazon-Books: D = 64, 512 negatives)."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, golden
from oracle import hstu_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


def _run(q, ids, pe, w, neg, table, T, l2, eps):
    from generative_recommenders_b200.ops.sampled_softmax import sampled_softmax_loss

    qd, ped, tbd = q.to(DEV).requires_grad_(), pe.to(DEV).requires_grad_(), table.to(DEV).requires_grad_()
    loss = sampled_softmax_loss(qd, ids.to(DEV), ped, w.to(DEV), neg.to(DEV), tbd, T, l2, eps)
    loss.backward()
    return loss.detach().cpu(), qd.grad.cpu(), ped.grad.cpu(), tbd.grad.cpu()


@pytest.mark.parametrize("fname", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "ssl_*.pt"))))
def test_sampled_softmax_golden(fname):
    g = golden(fname)
    N = g["supervision_ids"].shape[1]
    keep = torch.arange(N).unsqueeze(0) < g["lengths"].unsqueeze(1)
    q, ids, pe, w = g["output_embeddings"][keep], g["supervision_ids"][keep], g["supervision_embeddings"][keep], g["weights"][keep]
    loss, dq, dp, dt = _run(q, ids, pe, w, g["sampled_ids"], g["table"], g["temperature"], g["l2_norm"], g["l2_norm_eps"])
    f32 = g["table"].dtype == torch.float32
    if f32:
        ref_loss, rdq, rdp, rdt = g["loss"], g["d_out"][keep], g["d_sup"][keep], g["d_table"]
    else:  # the bf16 fixture is the reference evaluated IN bf16; the kernel accumulates in fp32: compare with the fp32 oracle
        qq, pp, tt = q.float().requires_grad_(), pe.float().requires_grad_(), g["table"].float().requires_grad_()
        ref_loss = O.sampled_softmax_loss(qq, ids, pp, w, g["sampled_ids"], tt, g["temperature"], g["l2_norm"], g["l2_norm_eps"])
        ref_loss.backward()
        ref_loss, rdq, rdp, rdt = ref_loss.detach(), qq.grad, pp.grad, tt.grad
        # and the reference's own bf16 result is within bf16 resolution of ours
        assert abs(float(loss) - float(g["loss"])) <= 2e-2 * abs(float(g["loss"]))
    tol = 2e-5 if f32 else 4e-3  # bf16: storage rounding of the gradients (fp32 math inside)
    assert abs(float(loss) - float(ref_loss)) <= (1e-5 if f32 else 4e-3) * abs(float(ref_loss))
    assert O.rel_l2(dq.float(), rdq.float()) <= tol
    assert O.rel_l2(dp.float(), rdp.float()) <= tol
    assert O.rel_l2(dt.float(), rdt.float()) <= tol


@pytest.mark.parametrize("D,R,dtype", [(64, 512, torch.bfloat16), (256, 128, torch.bfloat16), (32, 128, torch.float32)])
def test_sampled_softmax_config_shapes_vs_oracle(D, R, dtype):
    gen = torch.Generator().manual_seed(7 + D)
    N, V = 3000, 5000
    table = (torch.randn(V, D, generator=gen) * 0.2).to(dtype)
    ids = torch.randint(0, V, (N,), generator=gen)
    neg = torch.randint(0, V, (N, R), generator=gen)
    neg[5, 3] = ids[5]  # a sampled negative that collides with the positive: masked to -5e4
    q = (torch.randn(N, D, generator=gen) * 0.5).to(dtype)
    pe = table[ids].clone()
    w = (torch.rand(N, generator=gen) > 0.3).float()
    loss, dq, dp, dt = _run(q, ids, pe, w, neg, table, 0.05, True, 1e-6)
    qq, pp, tt = q.float().requires_grad_(), pe.float().requires_grad_(), table.float().requires_grad_()
    ref = O.sampled_softmax_loss(qq, ids, pp, w, neg, tt, 0.05, True, 1e-6)
    ref.backward()
    f32 = dtype == torch.float32
    assert abs(float(loss) - float(ref)) <= (1e-5 if f32 else 4e-3) * abs(float(ref))
    tol = 3e-5 if f32 else 4e-3
    assert O.rel_l2(dq.float(), qq.grad) <= tol
    assert O.rel_l2(dp.float(), pp.grad) <= tol
    assert O.rel_l2(dt.float(), tt.grad) <= tol


def test_sampled_softmax_module_runs_and_is_seed_reproducible():
    from generative_recommenders_b200.modules.sampled_softmax import LocalNegativesSampler, SampledSoftmaxLoss

    torch.manual_seed(0)
    V, D, B, N = 300, 64, 4, 20
    emb = torch.nn.Embedding(V, D).to(DEV)
    sampler = LocalNegativesSampler(V, emb, list(range(V)), True, 1e-6).to(DEV)
    mod = SampledSoftmaxLoss(num_to_sample=64, softmax_temperature=0.05)
    lengths = torch.tensor([20, 3, 11, 1], device=DEV)
    out = torch.randn(B, N, D, device=DEV, requires_grad=True)
    ids = torch.randint(0, V, (B, N), device=DEV)
    w = torch.ones(B, N, device=DEV)
    losses = []
    for _ in range(2):
        torch.cuda.manual_seed(123)
        loss, aux = mod(lengths=lengths, output_embeddings=out, supervision_ids=ids, supervision_embeddings=emb(ids),
                        supervision_weights=w, negatives_sampler=sampler)
        losses.append(float(loss))
    assert aux == {} and losses[0] == losses[1] and losses[0] > 0
    loss.backward()
    assert torch.isfinite(out.grad).all() and torch.isfinite(emb.weight.grad).all()
