"""Row-wise kernels of one STU layer at the bench shape, each timed alone through the C ABI (CUDA events, rotating buffer sets
larger than L2 so that every launch reads cold data), against the HBM bandwidth of MEASURED_PEAKS.json.

    python scripts/rowwise_bench.py [--rows 123699] [--iters 20] [--only ln_fwd,nmd_bwd]

Prints one JSON line: per kernel ms, algorithmic GB/s, fraction of the measured HBM peak.  Algorithmic bytes per row (D = H dv
elements of e bytes): ln_fwd 2 D e, ln_bwd 3 D e, nmd_fwd (concat_ux) 5 D e, nmd_bwd 7 D e, silu_fwd 2 D e, silu_bwd 3 D e.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=123699)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--only", default="")
    ap.add_argument("--profile", action="store_true", help="one launch per kernel, no warm-up (for ncu)")
    args = ap.parse_args()
    from generative_recommenders_b200 import _lib
    from generative_recommenders_b200.build import build
    from generative_recommenders_b200.ops import hstu_compute as hc
    from generative_recommenders_b200.ops import layer_norm as ln

    build()
    _lib.lib()
    dev = torch.device("cuda", 0)
    L, H, dv = args.rows, 8, 32
    D = H * dv
    dt = torch.bfloat16
    torch.manual_seed(0)
    S = args.sets
    mk = lambda w=D: [torch.randn(L, w, device=dev, dtype=dt) for _ in range(S)]
    x, dy, attn, u, dy3 = mk(), mk(), mk(), mk(), mk(3 * D)
    dx = [torch.empty(L, D, device=dev, dtype=dt) for _ in range(S)]
    w = torch.randn(D, device=dev, dtype=dt)
    b = torch.randn(D, device=dev, dtype=dt)
    stats = [ln.cuda_layer_norm_fwd(x[i], w, b, 1e-6, False)[1:] for i in range(S)]
    nstats = [hc.cuda_norm_mul_dropout_fwd(attn[i], u[i], w, b, 1e-6, 0.2, 1234, False, True, False, H, dv)[1:] for i in range(S)]
    e = 2
    cases = {
        "ln_fwd": (2 * D * e, lambda i: ln.cuda_layer_norm_fwd(x[i], w, b, 1e-6, False)),
        "ln_bwd": (3 * D * e, lambda i: ln.cuda_layer_norm_bwd(dy[i], x[i], w, b, stats[i][0], stats[i][1], False)),
        "nmd_fwd": (5 * D * e, lambda i: hc.cuda_norm_mul_dropout_fwd(attn[i], u[i], w, b, 1e-6, 0.2, 1234, False, True, False, H, dv)),
        "nmd_bwd": (7 * D * e, lambda i: hc.cuda_norm_mul_dropout_bwd(dy3[i], attn[i], u[i], w, b, nstats[i][0], nstats[i][1], 0.2,
                                                                     1234, False, True, False, H, dv)),
        "silu_fwd": (2 * D * e, lambda i: hc.cuda_silu_fwd(x[i])),
        "silu_bwd": (3 * D * e, lambda i: hc.cuda_silu_bwd(dy[i], x[i], dx[i])),
    }
    only = [s for s in args.only.split(",") if s]
    peak = 6570.6
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"rows": L, "D": D, "hbm_peak_gbs": peak, "kernels": {}}
    for name, (bytes_per_row, fn) in cases.items():
        if only and name not in only:
            continue
        if args.profile:
            fn(0)
            continue
        for i in range(3):
            fn(i % S)
        torch.cuda.synchronize()
        pairs = []
        for i in range(args.iters):  # one event pair per launch: host-side call overhead between launches is not counted
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(i % S)
            e1.record()
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in pairs) / args.iters
        gbs = bytes_per_row * L / ms / 1e6
        out["kernels"][name] = {"ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
