#!/usr/bin/env python3
"""Comparator recipe: copy the reference's OWN Triton jagged attention (unmodified) into the git-ignored baseline/_ref/ so that
`bench.py --workload attn --impl triton` can time it on the same B200, on the same seeded inputs, next to the kernels of this
repo (north_star: ">= the reference's Triton hstu_attention").  Nothing under baseline/_ref/ is product code, nothing of it is
committed (.gitignore), and the product package never imports it; it travels to the GPU box with the gpurun snapshot because
/root/reference does not exist there.

    python scripts/fetch_triton_baseline.py            # needs /root/reference (build container only)

Files (paths relative to /root/reference/generative_recommenders/):
    common.py                               triton_autotune, autotune_max_seq_len, switch_to_contiguous_if_needed, ...
    ops/triton/triton_hstu_attention.py     _hstu_attn_fwd / _hstu_attn_bwd kernels, triton_hstu_mha (:2066-2092)
    ops/triton/triton_attention_utils.py    acc_dq
"""
import os
import shutil
import sys

SRC = "/root/reference/generative_recommenders"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref", "generative_recommenders")
FILES = ["common.py", "ops/triton/triton_hstu_attention.py", "ops/triton/triton_attention_utils.py"]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not found: the comparator can only be fetched in the build container", file=sys.stderr)
        return 1
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
        print("copied", f)
    return 0


if __name__ == "__main__":
    sys.exit(main())
