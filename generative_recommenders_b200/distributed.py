"""Data-parallel plumbing of the HSTU block: shard the user batch across ranks, all-reduce parameter gradients.

The reference trains with plain DDP (research/trainer/train.py:78,269: NCCL process group, bucketed gradient all-reduce);
every op of the hot path is per-sequence, so forward/backward need no communication.  Here:
  * `shard_sequences`  - balanced partition of the user sequences over ranks by attention cost (sum of len^2),
  * `rebase_offsets`   - seq_offsets of a shard, starting at 0,
  * `LayerBucketAllReduce` - one flat all-reduce per STU layer, issued as soon as that layer's gradients are final
                             (post-accumulate-grad hooks), on a side stream so that it overlaps the remaining backward.
Works with NCCL on GPUs and with gloo on CPU tensors (used by the world_size-2 CPU tests).
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_sequences(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time partition of sequence indices by cost len^2; deterministic; every index appears once."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda w: (loads[w], w))
        shards[r].append(i)
        loads[r] += int(lengths[i]) ** 2
    for s in shards:
        s.sort()
    return shards


def rebase_offsets(lengths: Sequence[int], indices: Sequence[int], dtype=torch.int64) -> torch.Tensor:
    off = torch.zeros(len(indices) + 1, dtype=dtype)
    if len(indices):
        off[1:] = torch.cumsum(torch.tensor([int(lengths[i]) for i in indices], dtype=dtype), 0)
    return off


def gather_rows(values: torch.Tensor, seq_offsets: torch.Tensor, indices: Sequence[int]) -> torch.Tensor:
    """Rows of the selected sequences, concatenated in the order of `indices`."""
    off = seq_offsets.tolist()
    if not len(indices):
        return values[:0]
    return torch.cat([values[off[i]:off[i + 1]] for i in indices], dim=0)


class LayerBucketAllReduce:
    """Averages gradients across ranks, one flat bucket per layer, overlapped with the backward of earlier layers.

    The gradients of a layer LIVE in one flat buffer of the parameters' dtype (`p.grad` is a view into it, as with DDP's
    `gradient_as_bucket_view`), so the collective runs in place on that buffer: no concatenation, no fp32 staging copy, no
    scatter back.  A post-accumulate-grad hook counts the parameters of a layer; when the last one has its gradient the
    layer's slice is all-reduced (AVG on NCCL, SUM + divide on gloo) on a side stream.  `wait()` reduces whatever has not been
    reduced in this iteration (layers with frozen or unused parameters: their slots hold zeros on every rank) and resets the
    bookkeeping, so a partial backward can never leave ranks out of step.  Use `zero_grad()` of this object instead of the
    optimizer's: it clears the flat buffer with one memset and keeps the views attached.
    """

    def __init__(self, layers: Sequence[torch.nn.Module], world_size: int, device: torch.device):
        self.world = world_size
        self.device = device
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self.launched = 0
        self._layers = []
        by_dtype = {}
        for layer in layers:
            params = [p for p in layer.parameters() if p.requires_grad]
            if not params:
                continue
            dt = params[0].dtype
            if any(p.dtype != dt for p in params):
                raise RuntimeError("LayerBucketAllReduce: the parameters of one layer must share a dtype")
            by_dtype.setdefault(dt, []).append(params)
        self._flats = []
        for dt, groups in by_dtype.items():
            # every layer slice starts at a multiple of 128 elements (>= 256 B): aligned collectives, vectorised optimizer
            sizes = [(sum(p.numel() for p in g) + 127) // 128 * 128 for g in groups]
            flat = torch.zeros(sum(sizes), dtype=dt, device=device)
            self._flats.append(flat)
            lo = 0
            for g, sz in zip(groups, sizes):
                state = {"params": g, "slice": flat[lo:lo + sz], "seen": 0, "done": False, "views": []}
                o = lo
                for p in g:
                    view = flat[o:o + p.numel()].view_as(p)
                    state["views"].append(view)
                    p.grad = view
                    o += p.numel()
                    p.register_post_accumulate_grad_hook(self._make_hook(state))
                self._layers.append(state)
                lo += sz
        backend = dist.get_backend() if (world_size > 1 and dist.is_initialized()) else None
        self._avg = backend == "nccl"

    def _make_hook(self, state):
        def hook(_p):
            state["seen"] += 1
            if state["seen"] == len(state["params"]) and not state["done"]:
                self._reduce(state)
        return hook

    def _reduce(self, state) -> None:
        state["done"] = True
        if self.world <= 1:
            return
        self.launched += 1
        if self.stream is not None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                self._reduce_now(state["slice"])
        else:
            self._reduce_now(state["slice"])

    def _reduce_now(self, flat: torch.Tensor) -> None:
        if self._avg:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat)
            flat.div_(self.world)

    def zero_grad(self) -> None:
        for flat in self._flats:
            flat.zero_()
        for state in self._layers:
            for p, view in zip(state["params"], state["views"]):
                if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                    p.grad = view  # re-attach (e.g. after an optimizer.zero_grad(set_to_none=True))

    def wait(self) -> None:
        """Call after backward, before the optimizer step."""
        for state in self._layers:
            if not state["done"]:
                self._reduce(state)  # frozen / unused parameters: reduce what is there so that ranks stay in step
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        for state in self._layers:
            state["seen"], state["done"] = 0, False
