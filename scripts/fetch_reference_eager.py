#!/usr/bin/env python3
"""CPU-baseline recipe: copy the reference's OWN PyTorch-eager HSTU block (unmodified Python) into the git-ignored oracle/_ref/
so that `bench.py` can time the real reference (cpu_baseline.kind = "reference") on the GPU box's host cores, where
/root/reference does not exist.  The three fbgemm_gpu jagged ops it calls are supplied by oracle/fbgemm_shim.py (the package is
a third-party dependency that is not vendored under /root/reference).  Nothing under oracle/_ref/ is committed or imported by the
product package.

    python scripts/fetch_reference_eager.py            # needs /root/reference (build container only)

Copied (relative to /root/reference/generative_recommenders/): common.py, modules/stu.py, ops/*.py, ops/pytorch/*.py,
ops/triton/*.py (imported by the facades at module import time; never called on the CPU path).
"""
import glob
import os
import shutil
import sys

SRC = "/root/reference/generative_recommenders"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "oracle", "_ref", "generative_recommenders")
PATTERNS = ["common.py", "modules/stu.py", "ops/*.py", "ops/pytorch/*.py", "ops/triton/*.py"]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not found: the reference can only be fetched in the build container", file=sys.stderr)
        return 1
    n = 0
    for pat in PATTERNS:
        for src in sorted(glob.glob(os.path.join(SRC, pat))):
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            n += 1
    print(f"copied {n} files to {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
