#!/usr/bin/env python3
"""Extract per-launch DRAM traffic of the attention kernels from an .ncu-rep into profiles/ncu_traffic.json.

usage: ncu_traffic.py rep out.json "<workload description>" batch lmax heads d
bench.py reads the JSON and reports `roofline.traffic` only when its attention shape equals the captured one."""
import csv, io, json, subprocess, sys
rep, out, desc = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(r, name):
    i = hdr.index(name)
    v, u = float(r[i].replace(",", "")), units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
res = {"source": rep.split("/")[-1], "workload": desc, "kernels": {},
       "attn_shape": dict(zip(("batch", "lmax", "heads", "d"), map(int, sys.argv[4:8])))}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    key = "attn_bwd" if "attn_bwd" in name else "attn_fwd" if "attn_fwd" in name else None
    if key is None: continue
    rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
    res["kernels"][key] = {"kernel": name.split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                           "gpu_time_ms_under_ncu": col(r, "gpu__time_duration.sum") / 1e6 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else None}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
