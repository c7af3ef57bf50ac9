"""CPU: the C-ABI library loads, exports every symbol include/hstu_b200.h declares, and its host-side mask /
tile-range logic agrees with the oracle (no GPU, no compute kernels)."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import hstu_oracle as O


@pytest.fixture(scope="module")
def lib():
    from generative_recommenders_b200 import _lib
    from generative_recommenders_b200.build import build

    build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "hstu_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(hstu_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/hstu_b200.h but not exported"
    from generative_recommenders_b200 import _lib

    assert names == set(_lib.EXPORTED_SYMBOLS), names ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.hstu_abi_version() == 1


def test_struct_layout_matches_header(lib):
    from generative_recommenders_b200 import _lib

    # sizeof(hstu_attn_params): 16 int32/float + ... ; a mismatch would make every call fail the abi/arg checks
    p = _lib.AttnParams()
    p.abi_version = 1
    p.dtype = _lib.BF16
    p.batch, p.heads, p.dqk, p.dv, p.max_seq_len, p.total_rows = 2, 2, 32, 32, 64, 0
    assert lib.hstu_attn_select_impl(C.byref(p), 0) in (_lib.IMPL_GENERIC, _lib.IMPL_UMMA)
    p.max_seq_len = 0
    assert lib.hstu_attn_select_impl(C.byref(p), 0) < 0
    assert b"max_seq_len" in lib.hstu_last_error()
    p.max_seq_len = 64
    p.abi_version = 7
    assert lib.hstu_attn_select_impl(C.byref(p), 0) < 0
    assert b"ABI" in lib.hstu_last_error()


CASES = [
    # (len, n_tgt, window, min_full, ctx)
    (1, -1, 0, 0, 0), (37, -1, 0, 0, 0), (300, 5, 0, 0, 0), (200, 0, 0, 0, 0), (150, 150, 0, 0, 0),
    (260, 7, 11, 0, 0), (260, -1, 40, 0, 0), (190, 9, 13, 20, 0), (140, 4, 0, 0, 6), (333, 12, 25, 0, 3),
    (333, 12, 25, 30, 3), (90, 3, 200, 0, 5), (129, 1, 1, 0, 1), (257, 128, 3, 2, 2),
]


@pytest.mark.parametrize("case", CASES)
def test_mask_matches_oracle(lib, case):
    n, nt, win, mf, ctx = case
    ref = O.attn_valid_mask(n, None if nt < 0 else nt, win, ctx, mf)
    got = np.zeros_like(ref)
    for i in range(n):
        for j in range(n):
            got[i, j] = lib.hstu_mask_valid(n, nt, win, mf, ctx, i, j)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tile", [32, 64, 128])
def test_tile_ranges_cover_the_mask(lib, case, tile):
    n, nt, win, mf, ctx = case
    ref = O.attn_valid_mask(n, None if nt < 0 else nt, win, ctx, mf)
    lo, hi, chi = C.c_int32(), C.c_int32(), C.c_int32()
    for m0 in range(0, n, tile):
        m1 = min(n, m0 + tile)
        assert lib.hstu_kv_range_for_q_rows(n, nt, win, mf, ctx, m0, m1, C.byref(lo), C.byref(hi)) == 0
        cols = np.nonzero(ref[m0:m1].any(axis=0))[0]
        assert cols.min() >= lo.value and cols.max() < hi.value, (case, m0, lo.value, hi.value, cols.min(), cols.max())
        # not absurdly loose: the plain-causal upper bound is exact
        if ctx == 0:
            assert hi.value == m1
        assert lib.hstu_q_range_for_kv_rows(n, nt, win, mf, ctx, m0, m1, C.byref(lo), C.byref(hi), C.byref(chi)) == 0
        rows = np.nonzero(ref[:, m0:m1].any(axis=1))[0]
        for r in rows:
            assert (lo.value <= r < hi.value) or r < chi.value, (case, m0, r, lo.value, hi.value, chi.value)


def test_host_layer_rejects_cpu_and_foreign_kernels(lib):
    from generative_recommenders_b200.common import HammerKernel
    from generative_recommenders_b200.ops.hstu_attention import hstu_mha
    from generative_recommenders_b200.ops.layer_norm import layer_norm

    q = torch.zeros(4, 1, 16)
    off = torch.tensor([0, 4])
    with pytest.raises(NotImplementedError):
        hstu_mha(8, 0.25, q, q, q, off, kernel=HammerKernel.PYTORCH)
    with pytest.raises(RuntimeError, match="CUDA"):
        hstu_mha(8, 0.25, q, q, q, off, kernel=HammerKernel.CUDA)  # CPU tensors: no fallback
    with pytest.raises(RuntimeError, match="CUDA"):
        layer_norm(torch.zeros(2, 8), torch.ones(8), torch.zeros(8))
    with pytest.raises(Exception):
        hstu_mha(0, 0.25, q, q, q, off)


def test_synthetic_length_generators_match_reference_recipe():
    from generative_recommenders_b200.common import apply_sampling, generate_sparse_seq_len

    torch.manual_seed(1001)
    l = generate_sparse_seq_len(512, 8192, 0.95, torch.device("cpu"))
    assert l.dtype == torch.int32 and int(l.min()) >= int(0.9 * 8192) and int(l.max()) < 8192
    assert torch.equal(apply_sampling(l, 2.0, 8192), l)  # alpha = 2: threshold = max_seq_len, no-op
    assert int(generate_sparse_seq_len(4, 100, 0.0, torch.device("cpu")).sum()) == 0
