"""Summarise gpurun_out/bwd_trace.txt (HSTU_TRACE build of the backward kernel, CTA (0,0,0)): merged event timeline in clocks.

roles: 0 = issuer X (scores), 1 = issuer YV (dV), 4 = issuer YK (dK), 2/3 = elementwise warpgroups 0/1.
usage: trace_report.py [trace] [first_unit] [n_units]"""
import collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bwd_trace.txt"
u0 = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nu = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rows = collections.defaultdict(dict)
for l in open(path):
    p = l.split()
    if len(p) >= 6: rows[int(p[0])][int(p[1])] = [int(x) for x in p[2:6]]
t0 = min(v[0] for v in rows[0].values() if v[0])
ev = []
for u in range(u0, u0 + nu):
    a = rows[0].get(u)
    if a: ev += [(a[0] - t0, f'X  u{u} waits slot_free / q_full'), (a[1] - t0, f'X  u{u} issues S^T, dP^T'), (a[2] - t0, f'X  u{u} commit issued')]
    a = rows[1].get(u)
    if a: ev += [(a[0] - t0, f'YV u{u} waits unit_done'), (a[1] - t0, f'YV u{u} issues dV'), (a[2] - t0, f'YV u{u} commits issued')]
    a = rows[4].get(u)
    if a: ev += [(a[0] - t0, f'YK u{u} waits unit_done'), (a[1] - t0, f'YK u{u} issues dK'), (a[2] - t0, f'YK u{u} issued')]
for wg in (0, 1):
    for i in range(u0 // 2, (u0 + nu) // 2 + 1):
        a = rows[2 + wg].get(i)
        if a:
            u = 2 * i + wg
            ev += [(a[0] - t0, f'W{wg} u{u} waits s_full'), (a[1] - t0, f'W{wg} u{u} elementwise starts'),
                   (a[2] - t0, f'W{wg} u{u} arrives unit_done')]
for t, e in sorted(set(ev)):
    print(f'{t:8d}  {e}')
