"""Summarise gpurun_out/bwd_trace.txt (HSTU_TRACE build of the backward kernel, CTA (0,0,0)): merged event timeline in clocks.

roles: 0 = issuer X (scores), 1 = issuer YV (dV), 4 = issuer YK (dK), 5 = issuer Z (dQ), 6 = drain warpgroup, 2 / 3 = elementwise
warpgroup 0 / 1.  Index = half-tile unit u = 2 * tile + half for X / YV / YK, query tile for W0 / W1 / Z / DR.
usage: trace_report.py [trace] [first_index] [count]"""
import collections
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bwd_trace.txt"
i0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cnt = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rows = collections.defaultdict(dict)
for line in open(path):
    p = line.split()
    if len(p) >= 6:
        rows[int(p[0])][int(p[1])] = [int(x) for x in p[2:6]]
if not rows[0]:
    print("empty trace")
    sys.exit(0)
t0 = min(v[0] for v in rows[0].values() if v[0])
names = {0: ("X ", ["waits slot / q_ready", "issues S^T, dP^T", "commit issued"]),
         1: ("YV", ["waits P^T / dS^T ready", "issues dV", "commits issued"]),
         4: ("YK", ["waits P^T / dS^T ready", "issues dK", "commit issued"]),
         2: ("W0", ["waits s_full", "elementwise starts", "arrives ready"]),
         3: ("W1", ["waits s_full", "elementwise starts", "arrives ready"]),
         5: ("Z ", ["waits units", "units ready, waits dq_empty", "issues dQ"]),
         6: ("DR", ["waits tile_done", "tile done: loads next, drains dQ", "drained", "converted"])}
ev = []
for role, (nm, labels) in names.items():
    per_tile = role in (2, 3, 5, 6)  # the warpgroups, Z and the drain count query tiles, the issuers X / YV / YK half-tile units
    for i in (range(i0 // 2, (i0 + cnt + 1) // 2) if per_tile else range(i0, i0 + cnt)):
        a = rows[role].get(i)
        if a:
            ev += [(a[k] - t0, f"{nm} #{i} {labels[k]}") for k in range(len(labels)) if a[k]]
for t, e in sorted(set(ev)):
    print(f"{t:8d}  {e}")
