"""Timestamp + position embedding add in front of the STU stack -- host side of `hstu_position_embeddings_fwd/_bwd`.

Same call surface as the reference facade generative_recommenders/ops/position.py:43-96
(`add_timestamp_positional_embeddings`); semantics of the eager path ops/pytorch/pt_position.py:39-134, including its
time-bucket clamp `num_time_buckets = ts_embeddings.size(1) - 1` (:98).  `kernel` must be HammerKernel.CUDA.
"""
from typing import Optional

import torch

from .. import _lib
from ..common import HammerKernel, require_cuda_kernel


def _idx(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype not in (torch.int32, torch.int64):
        raise RuntimeError(f"{what} must be int32 or int64")
    return t.contiguous()


class _AddTimestampPositionEmbeddingsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, max_seq_len, max_contextual_seq_len, pos_w, ts_w, seq_offsets, seq_lengths, seq_embeddings,
                timestamps, num_targets, interleave_targets, time_bucket_fn):
        dev = _lib.require_cuda(seq_embeddings, pos_w, ts_w, seq_offsets, seq_lengths, timestamps, num_targets)
        x = seq_embeddings.contiguous()
        L, D = x.shape
        B = seq_lengths.numel()
        pw = pos_w.detach().float().contiguous()
        tw = ts_w.detach().float().contiguous()
        if pw.shape[1] != D or tw.shape[1] != D:
            raise RuntimeError("embedding tables must have the embedding dim of seq_embeddings")
        # pt_position.py:98 takes the bucket clamp from size(1); beyond the table it would index out of range there
        num_time_buckets = min(tw.shape[1] - 1, tw.shape[0] - 1)
        off, lens = _idx(seq_offsets, "seq_offsets"), _idx(seq_lengths, "seq_lengths")
        nt = None if num_targets is None else _idx(num_targets, "num_targets")
        ts = timestamps.to(torch.int64).contiguous()
        if ts.numel() != L:
            raise RuntimeError("timestamps must have one entry per row of seq_embeddings")
        out = torch.empty_like(x)
        pos_inds = torch.empty(L, dtype=torch.int32, device=dev)
        ts_inds = torch.empty(L, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev), _lib.timed("position_fwd", dev):
            _lib.check(_lib.lib().hstu_position_embeddings_fwd(
                x.data_ptr(), out.data_ptr(), pw.data_ptr(), tw.data_ptr(), off.data_ptr(), lens.data_ptr(), _lib.ptr(nt),
                ts.data_ptr(), pos_inds.data_ptr(), ts_inds.data_ptr(), L, B, D, pw.shape[0], num_time_buckets,
                int(max_contextual_seq_len), float(alpha), int(bool(interleave_targets)), int(time_bucket_fn == "log"),
                int(off.dtype == torch.int64), int(lens.dtype == torch.int64),
                int(nt is not None and nt.dtype == torch.int64), _lib.dtype_code(x), _lib.stream_ptr(dev)),
                "hstu_position_embeddings_fwd")
            _lib.note_launch(1)
        ctx.save_for_backward(pos_inds, ts_inds)
        ctx.cfg = (float(alpha), tuple(pos_w.shape), tuple(ts_w.shape), pos_w.dtype, ts_w.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        pos_inds, ts_inds = ctx.saved_tensors
        alpha, pshape, tshape, pdt, tdt = ctx.cfg
        dev = dout.device
        dout = dout.contiguous()
        L, D = dout.shape
        dseq = torch.empty_like(dout)
        dpos = torch.zeros(pshape, dtype=torch.float32, device=dev)
        dts = torch.zeros(tshape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _lib.timed("position_bwd", dev):
            _lib.check(_lib.lib().hstu_position_embeddings_bwd(
                dout.data_ptr(), dseq.data_ptr(), dpos.data_ptr(), dts.data_ptr(), pos_inds.data_ptr(), ts_inds.data_ptr(),
                L, D, alpha, _lib.dtype_code(dout), _lib.stream_ptr(dev)), "hstu_position_embeddings_bwd")
            _lib.note_launch(1)
        return (None, None, None, dpos.to(pdt), dts.to(tdt), None, None, dseq, None, None, None, None)


def add_timestamp_positional_embeddings(
    alpha: float,
    max_seq_len: int,
    max_contextual_seq_len: int,
    position_embeddings_weight: torch.Tensor,
    timestamp_embeddings_weight: torch.Tensor,
    seq_offsets: torch.Tensor,
    seq_lengths: torch.Tensor,
    seq_embeddings: torch.Tensor,
    timestamps: torch.Tensor,
    num_targets: Optional[torch.Tensor],
    interleave_targets: bool,
    time_bucket_fn: str = "sqrt",
    kernel: HammerKernel = HammerKernel.CUDA,
) -> torch.Tensor:
    """Drop-in for generative_recommenders.ops.position.add_timestamp_positional_embeddings (position.py:43-96):
    out = seq_embeddings * alpha + (pos_emb[pos_ind] + ts_emb[time_bucket]).to(dtype), rows addressed by seq_offsets."""
    assert time_bucket_fn in ["sqrt", "log"]
    require_cuda_kernel(kernel, "add_timestamp_positional_embeddings")
    return _AddTimestampPositionEmbeddingsFunction.apply(
        alpha, max_seq_len, max_contextual_seq_len, position_embeddings_weight, timestamp_embeddings_weight, seq_offsets,
        seq_lengths, seq_embeddings, timestamps, num_targets, interleave_targets, time_bucket_fn)
