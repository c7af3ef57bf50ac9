"""Build a compile-time variant of libhstu_b200.so next to the product library, for A/B runs on the GPU box without rebuilding there:

    python scripts/build_variant.py NAME "FLAG1 FLAG2=3" norm.cu[,other.cu]
    HSTU_B200_LIB=generative_recommenders_b200/lib/variants/libhstu_b200_NAME.so python scripts/rowwise_bench.py

Only the listed translation units are recompiled (with -DFLAG...); the rest is linked from the objects of the default build."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from generative_recommenders_b200 import build as B  # noqa: E402

name, flags, srcs = sys.argv[1], sys.argv[2].split(), sys.argv[3].split(",")
B.build()
out_dir = os.path.join(B.HERE, "lib", "variants")
os.makedirs(out_dir, exist_ok=True)
objs = []
for s in B.SOURCES:
    if s in srcs:
        obj = os.path.join(B.OBJ, f"{name}__{s.replace('.cu', '.o')}")
        cmd = [B._nvcc()] + B.NVCC_FLAGS + ["-D" + f for f in flags] + ["-c", os.path.join(B.CSRC, s), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stdout + r.stderr)
        objs.append(obj)
    else:
        objs.append(os.path.join(B.OBJ, s.replace(".cu", ".o")))
lib = os.path.join(out_dir, f"libhstu_b200_{name}.so")
r = subprocess.run([B._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
if r.returncode != 0:
    raise SystemExit(r.stdout + r.stderr)
print(lib)
