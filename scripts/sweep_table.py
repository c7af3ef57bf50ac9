#!/usr/bin/env python3
"""Table "ours / Triton" from the JSON lines of `bench.py --workload attn` (ours) and `--impl triton` (the reference's kernel)."""
import json
import os
import sys


def load(path):
    out = {}
    if not os.path.exists(path):
        return out
    for line in open(path):
        try:
            d = json.loads(line)
        except Exception:
            continue
        s = d.get("config", {}).get("attn_shape")
        if not s:
            continue
        out[(s["batch"], s["heads"], s["d"], s["lmax"])] = d
    return out


def main():
    ours, tri = load(sys.argv[1]), load(sys.argv[2]) if len(sys.argv) > 2 else {}
    peak_tf, peak_bw = 1467.7, 6570.6
    try:
        pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        peak_tf, peak_bw = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
    except Exception:
        pass
    print(f"peaks (MEASURED_PEAKS.json): bf16 sustained {peak_tf} TFLOP/s, HBM {peak_bw} GB/s; ALGORITHMIC work / CUDA-event time")
    print("   B  H    d  Lmax |  ours fwd ms  TF/s frac  GB/s | ours bwd ms  TF/s frac  GB/s | triton fwd ms  bwd ms | ours/triton fwd  bwd  fwd+bwd")
    for key in sorted(set(ours) | set(tri), key=lambda k: (k[2], k[3])):
        o, t = ours.get(key), tri.get(key)
        row = "%4d %2d %4d %5d |" % key
        if o and "roofline" in o:
            r = o["roofline"]
            row += " %11.3f %5.0f %4.2f %5.0f | %11.3f %5.0f %4.2f %5.0f |" % (
                r["fwd"]["ms_per_launch"], r["fwd"]["achieved"], r["fwd"]["achieved"] / peak_tf, r["fwd"]["hbm_gbs_algorithmic"],
                r["ms_per_launch"], r["achieved"], r["achieved"] / peak_tf, r["bwd_hbm_gbs_algorithmic"])
        else:
            row += " " * 27 + "n/a" + " " * 26 + "|"
        if t and "kernel_ms_per_call" in t:
            tf, tb = t["kernel_ms_per_call"]["attn_fwd"], t["kernel_ms_per_call"]["attn_bwd"]
            row += " %13.3f %7.3f |" % (tf, tb)
            if o and "roofline" in o:
                of, ob = o["roofline"]["fwd"]["ms_per_launch"], o["roofline"]["ms_per_launch"]
                row += " %15.2f %4.2f %8.2f" % (tf / of, tb / ob, (tf + tb) / (of + ob))
        elif t:
            row += " unavailable: " + str(t.get("unavailable"))[:80]
        print(row)


if __name__ == "__main__":
    main()
