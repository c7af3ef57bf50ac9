"""Fused sampled-softmax loss (dot-product similarity, negatives gathered straight from the item embedding table) -- host side
of `hstu_sampled_softmax_fwd/_bwd` (include/hstu_b200.h).

Reference math: research/modeling/sequential/losses/sampled_softmax.py:43-89 with LocalNegativesSampler
(autoregressive_losses.py:73-121) and DotProductSimilarity (rails/similarities/dot_product_similarity_fn.py:31-67).  The eager
path materialises the [N, R, D] tensor of negatives; here it never exists.
"""
import ctypes as C

import torch

from .. import _lib


class _SampledSoftmaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, pos_ids, pos_emb, weights, neg_ids, table, temperature, l2_norm, l2_eps):
        dev = _lib.require_cuda(q, pos_ids, pos_emb, weights, neg_ids, table)
        if not (q.dtype == pos_emb.dtype == table.dtype):
            raise RuntimeError("output embeddings, supervision embeddings and the item table must share a dtype")
        q, pos_emb, table = q.contiguous(), pos_emb.contiguous(), table.contiguous()
        pos_ids = pos_ids.to(torch.int64).contiguous()
        neg_ids = neg_ids.to(torch.int64).contiguous()
        N, D = q.shape
        R = neg_ids.shape[1] if neg_ids.dim() == 2 else 0
        if tuple(pos_emb.shape) != (N, D) or pos_ids.numel() != N or (R and neg_ids.shape[0] != N) or table.shape[1] != D:
            raise RuntimeError("sampled softmax: inconsistent shapes")
        logits = torch.empty((N, R + 1), dtype=torch.float32, device=dev)
        rnorm = torch.empty((N, R + 1), dtype=torch.float32, device=dev)
        lse = torch.empty(N, dtype=torch.float32, device=dev)
        loss_rows = torch.empty(N, dtype=torch.float32, device=dev)
        p = _lib.SslParams()
        p.abi_version, p.dtype, p.N, p.R, p.D = _lib.ABI_VERSION, _lib.dtype_code(q), N, R, D
        p.l2_norm, p.l2_eps, p.temperature = int(bool(l2_norm)), float(l2_eps), float(temperature)
        p.q, p.pos_emb, p.table = q.data_ptr(), pos_emb.data_ptr(), table.data_ptr()
        p.pos_ids, p.neg_ids = pos_ids.data_ptr(), neg_ids.data_ptr()
        p.logits, p.rnorm, p.lse, p.loss_rows = logits.data_ptr(), rnorm.data_ptr(), lse.data_ptr(), loss_rows.data_ptr()
        with torch.cuda.device(dev), _lib.timed("ssl_fwd", dev):
            _lib.check(_lib.lib().hstu_sampled_softmax_fwd(C.byref(p), _lib.stream_ptr(dev)), "hstu_sampled_softmax_fwd")
            _lib.note_launch(1)
        w = weights.float()
        wsum = w.sum()
        loss = (loss_rows * w).sum() / wsum  # sampled_softmax.py:87-89
        ctx.save_for_backward(q, pos_emb, table, pos_ids, neg_ids, logits, rnorm, lse, w, wsum)
        ctx.cfg = (float(temperature), bool(l2_norm), float(l2_eps))
        return loss.to(q.dtype)

    @staticmethod
    def backward(ctx, dloss):
        q, pos_emb, table, pos_ids, neg_ids, logits, rnorm, lse, w, wsum = ctx.saved_tensors
        temperature, l2_norm, l2_eps = ctx.cfg
        dev = q.device
        N, D = q.shape
        R = logits.shape[1] - 1
        row_coef = (dloss.float() * w / wsum).contiguous()
        d_q = torch.empty_like(q)
        d_pos = torch.zeros_like(pos_emb)
        d_table = torch.zeros(table.shape, dtype=torch.float32, device=dev)
        p = _lib.SslParams()
        p.abi_version, p.dtype, p.N, p.R, p.D = _lib.ABI_VERSION, _lib.dtype_code(q), N, R, D
        p.l2_norm, p.l2_eps, p.temperature = int(l2_norm), l2_eps, temperature
        p.q, p.pos_emb, p.table = q.data_ptr(), pos_emb.data_ptr(), table.data_ptr()
        p.pos_ids, p.neg_ids = pos_ids.data_ptr(), neg_ids.data_ptr()
        p.logits, p.rnorm, p.lse = logits.data_ptr(), rnorm.data_ptr(), lse.data_ptr()
        p.row_coef, p.d_q, p.d_pos_emb, p.d_table = row_coef.data_ptr(), d_q.data_ptr(), d_pos.data_ptr(), d_table.data_ptr()
        with torch.cuda.device(dev), _lib.timed("ssl_bwd", dev):
            _lib.check(_lib.lib().hstu_sampled_softmax_bwd(C.byref(p), _lib.stream_ptr(dev)), "hstu_sampled_softmax_bwd")
            _lib.note_launch(1)
        return d_q, None, d_pos, None, None, d_table.to(table.dtype), None, None, None


def sampled_softmax_loss(output_embeddings: torch.Tensor, supervision_ids: torch.Tensor, supervision_embeddings: torch.Tensor,
                         supervision_weights: torch.Tensor, sampled_ids: torch.Tensor, item_embedding_table: torch.Tensor,
                         softmax_temperature: float, l2_norm: bool, l2_norm_eps: float) -> torch.Tensor:
    """loss = sum_i w_i * (-log_softmax([q_i . n(pos_i), q_i . n(E[id_i1]), ...] / T)[0]) / sum_i w_i; negatives whose id equals the
    positive id are masked to -5e4; n(.) = l2 normalisation with the norm clamped at l2_norm_eps (only if l2_norm)."""
    return _SampledSoftmaxFunction.apply(output_embeddings, supervision_ids, supervision_embeddings, supervision_weights,
                                         sampled_ids, item_embedding_table, softmax_temperature, l2_norm, l2_norm_eps)
