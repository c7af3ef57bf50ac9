"""The mbarrier protocol of the persistent tcgen05 forward kernel, model-checked on the CPU under random schedules
(scripts/sim_fwd_protocol.py mirrors the waits / arrives / commits of attn_fwd_umma_persist_kernel in csrc/attn_umma_fwd.cu)."""
import importlib.util
import os
import random

import pytest

_spec = importlib.util.spec_from_file_location(
    "sim_fwd_protocol", os.path.join(os.path.dirname(__file__), "..", "scripts", "sim_fwd_protocol.py"))
sim = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sim)


@pytest.mark.parametrize("conv,buffers", [(True, 2), (False, 2), (True, 1), (False, 1)])
def test_persistent_forward_protocol(conv, buffers):
    rnd = random.Random(3)
    lists = list(sim.item_lists(rnd, 30)) + [[1] * 9, [4] * 5, [64, 1, 64], [0, 0, 3], [2], [0]]
    for items in lists:
        for seed in range(8):
            sim.run(items, seed, conv, buffers, buffers)
