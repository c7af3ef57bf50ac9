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
    """Averages gradients across ranks, one flat bucket per layer, overlapped with the backward of earlier layers."""

    def __init__(self, layers: Sequence[torch.nn.Module], world_size: int, device: torch.device):
        self.world = world_size
        self.device = device
        self.pending: List[torch.Tensor] = []
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self.launched = 0
        for layer in layers:
            params = [p for p in layer.parameters() if p.requires_grad]
            state = {"left": len(params)}
            for p in params:
                p.register_post_accumulate_grad_hook(self._make_hook(params, state))

    def _make_hook(self, params, state):
        def hook(_p):
            state["left"] -= 1
            if state["left"] == 0:
                state["left"] = len(params)
                self._reduce(params)
        return hook

    def _reduce(self, params) -> None:
        self.launched += 1
        if self.stream is not None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                self._reduce_now(params)
        else:
            self._reduce_now(params)

    def _reduce_now(self, params) -> None:
        flat = torch.cat([p.grad.reshape(-1) for p in params]).float()
        dist.all_reduce(flat)
        flat.div_(self.world)
        o = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[o:o + n].view_as(p.grad))
            o += n
        self.pending.append(flat)  # keep alive until the consumer stream has waited

    def wait(self) -> None:
        """Call after backward, before the optimizer step."""
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self.pending.clear()
