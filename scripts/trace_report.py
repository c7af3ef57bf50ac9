"""Summarise gpurun_out/bwd_trace.txt (HSTU_TRACE build of the backward kernel): per-actor phase durations in clocks."""
import collections, sys
rows = collections.defaultdict(dict)
for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bwd_trace.txt"):
    p = l.split()
    if len(p) < 6: continue
    rows[int(p[0])][int(p[1])] = [int(x) for x in p[2:6]]
t0 = min(v[0] for v in rows[0].values() if v[0])
lo, hi = 20, 32
for u in range(lo, hi):
    a = rows[0][u]; n = rows[0].get(u + 1, [0])[0]
    print('X', u, a[0] - t0, 'pre(dV wait+issue, q_full)', a[1] - a[0], 'scores+commit', a[2] - a[1], 'gap', n - a[2])
for u in range(lo, hi):
    a = rows[1][u]; n = rows[1].get(u + 1, [0])[0]
    print('Y', u, a[0] - t0, 'wait_unit', a[1] - a[0], 'dK', a[2] - a[1], 'dq_empty wait', (a[3] - a[2]) if a[3] else '-', 'dQ+commit', n - (a[3] if a[3] else a[2]))
for wg in (0, 1):
    for i in range(lo // 2, hi // 2):
        a = rows[2 + wg][i]
        print('WG', wg, i, a[0] - t0, 'wait_s', a[1] - a[0], 'elem', a[2] - a[1], 'drain', a[3] - a[2])
