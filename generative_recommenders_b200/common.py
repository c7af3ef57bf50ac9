"""Backend selector + synthetic-length generators.  Mirrors generative_recommenders/common.py:102-216 (reference)."""
import abc
from enum import Enum, unique
from typing import Any, Optional

import torch


@unique
class HammerKernel(Enum):
    # same members as the reference enum (common.py:102-107); this package implements CUDA only
    TRITON = "TRITON"
    PYTORCH = "PYTORCH"
    CUDA = "CUDA"
    TRITON_CC = "TRITON_CC"


def require_cuda_kernel(kernel: HammerKernel, op: str) -> None:
    if kernel != HammerKernel.CUDA:
        raise NotImplementedError(
            f"{op}: generative_recommenders_b200 implements HammerKernel.CUDA only (got {kernel}); "
            "use the reference package for its PyTorch / Triton backends"
        )


class HammerModule(torch.nn.Module, abc.ABC):
    """Same contract as the reference HammerModule (common.py:110-170) with CUDA as the default backend."""

    _is_inference: bool = False
    _hammer_kernel: Optional[HammerKernel] = None

    def __init__(self, is_inference: bool = False, hammer_kernel: Optional[HammerKernel] = None) -> None:
        super().__init__()
        self._is_inference = is_inference
        self._hammer_kernel = hammer_kernel

    def hammer_kernel(self) -> HammerKernel:
        return self._hammer_kernel if self._hammer_kernel is not None else HammerKernel.CUDA

    def recursive_setattr(self, name: str, value: Any) -> None:
        for _, module in self.named_modules():
            if hasattr(module, name):
                setattr(module, name, value)

    def set_is_inference(self, is_inference: bool) -> None:
        self._is_inference = is_inference
        self.recursive_setattr("_is_inference", is_inference)

    def set_hammer_kernel(self, hammer_kernel: HammerKernel) -> None:
        self._hammer_kernel = hammer_kernel
        self.recursive_setattr("_hammer_kernel", hammer_kernel)

    @property
    def is_inference(self) -> bool:
        return self._is_inference


def generate_sparse_seq_len(size: int, max_seq_len: int, sparsity: float, device: torch.device) -> torch.Tensor:
    """Synthetic sequence lengths, same recipe as the reference bench (common.py:173-201)."""
    if sparsity == 0.0:
        return torch.zeros(size=(size,), device=device, dtype=torch.int)
    elif sparsity == 1.0:
        return torch.ones(size=(size,), device=device, dtype=torch.int) * max_seq_len
    elif sparsity >= 0.5:
        min_seq_len = int((2 * sparsity - 1.0) * max_seq_len)
        return torch.randint(low=min_seq_len, high=max_seq_len, size=(size,), device=device, dtype=torch.int)
    else:
        hi = int(2 * sparsity * max_seq_len)
        return torch.randint(low=0, high=hi, size=(size,), device=device, dtype=torch.int)


def apply_sampling(lengths: torch.Tensor, alpha: float, max_seq_len: int) -> torch.Tensor:
    """Stochastic-length sub-sampling (common.py:204-216)."""
    threshold = int(max_seq_len ** (alpha / 2))
    no_sample_prob = (max_seq_len**alpha) / torch.pow(lengths, 2)
    users_to_sample = torch.logical_and(lengths > threshold, torch.rand_like(no_sample_prob) < 1 - no_sample_prob)
    return torch.where(users_to_sample, threshold, lengths)


def switch_to_contiguous_if_needed(x: torch.Tensor) -> torch.Tensor:
    """Only the last dim has to be dense (common.py:240-247): q/k/v may be strided views of uvqk."""
    return x if x.stride(-1) == 1 else x.contiguous()
