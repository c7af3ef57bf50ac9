#!/usr/bin/env python3
"""Discrete simulation of the mbarrier protocol of the PERSISTENT forward kernel (attn_fwd_umma_persist_kernel, csrc/attn_umma_fwd.cu).

One CTA walks a list of work items; item k has T_k key tiles (0 = skipped by every role).  All ring stages, score slots and barrier
phases run on a global key-tile counter g; Q buffers / O accumulator sets on the count n of non-empty items.  The model checks, under
random schedules, what scripts/sim_bwd_protocol.py checks for the backward: no deadlock, no parity wait that passes falsely, no waiter
two phases behind.  Actors: KP (TMA: Q + K tiles), VP (TMA: V tiles), CV (bf16 -> fp16 converter), QK and PV (MMA issuers; commits are
asynchronous), W0 / W1 (silu warpgroups: key tiles with g % 2 == w, then half of the epilogue each).
usage: sim_fwd_protocol.py [--seeds N]"""
import argparse
import importlib.util
import os
import random

_spec = importlib.util.spec_from_file_location("sim_bwd_protocol", os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim_bwd_protocol.py"))
_b = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_b)
Bar, Violation, _simulate = _b.Bar, _b.Violation, _b._simulate


def run(items, seed, conv=True, NQ=2, NO=2):
    """items: list of key-tile counts per work item; NQ Q buffers and NO O accumulator sets (2 / 2 for d <= 64, 1 / 1 for d = 128)."""
    rnd = random.Random(seed)
    B = {}
    for i in range(2):
        B[f"qf{i}"], B[f"qe{i}"], B[f"of{i}"] = Bar(1), Bar(1), Bar(1)
        B[f"oe{i}"] = Bar(2)             # 256 threads modelled as one arrival per warpgroup
    for i in range(3):
        B[f"kf{i}"], B[f"vf{i}"], B[f"vr{i}"], B[f"sf{i}"], B[f"pf{i}"], B[f"pd{i}"] = Bar(1), Bar(1), Bar(1), Bar(1), Bar(1), Bar(1)
    vrdy = "vr" if conv else "vf"

    def walk():
        g0 = n = 0
        for T in items:
            if T == 0:
                continue
            yield g0, n, T
            g0 += T
            n += 1

    def KP():
        for g0, n, T in walk():
            qb = n % NQ
            if n >= NQ:
                yield ("wait", f"qe{qb}", n // NQ - 1)
            yield ("async", f"qf{qb}")
            for i in range(T):
                g = g0 + i
                if g >= 3:
                    yield ("wait", f"sf{g % 3}", g // 3 - 1)
                yield ("async", f"kf{g % 3}")

    def VP():
        for g0, n, T in walk():
            for i in range(T):
                g = g0 + i
                if g >= 3:
                    yield ("wait", f"pd{g % 3}", g // 3 - 1)
                yield ("async", f"vf{g % 3}")

    def CV():
        for g0, n, T in walk():
            for i in range(T):
                g = g0 + i
                yield ("wait", f"vf{g % 3}", g // 3)
                yield ("arrive", f"vr{g % 3}")

    def QK():
        for g0, n, T in walk():
            yield ("wait", f"qf{n % NQ}", n // NQ)
            for i in range(T):
                g = g0 + i
                if g >= 3:
                    yield ("wait", f"pd{g % 3}", g // 3 - 1)
                yield ("wait", f"kf{g % 3}", g // 3)
                yield ("async", f"sf{g % 3}")
                if i == T - 1:
                    yield ("async", f"qe{n % NQ}")

    def PV():
        for g0, n, T in walk():
            if n >= NO:
                yield ("wait", f"oe{n % NO}", n // NO - 1)
            for i in range(T):
                g = g0 + i
                yield ("wait", f"pf{g % 3}", g // 3)
                yield ("wait", f"{vrdy}{g % 3}", g // 3)
                yield ("async", f"pd{g % 3}")
                if i == T - 1:
                    yield ("async", f"of{n % NO}")

    def W(w):
        for g0, n, T in walk():
            for i in range(T):
                g = g0 + i
                if (g & 1) != w:
                    continue
                yield ("wait", f"sf{g % 3}", g // 3)
                yield ("arrive", f"pf{g % 3}")
            yield ("wait", f"of{n % NO}", n // NO)
            yield ("arrive", f"oe{n % NO}")

    actors = {"KP": KP(), "VP": VP(), "QK": QK(), "PV": PV(), "W0": W(0), "W1": W(1)}
    if conv:
        actors["CV"] = CV()
    return _simulate(actors, B, rnd)


def item_lists(rnd, count):
    for _ in range(count):
        n = rnd.randint(1, 12)
        yield [rnd.choice([0, 1, 1, 2, 3, 4, 7]) for _ in range(n)]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=300)
    a = ap.parse_args()
    rnd = random.Random(1)
    bad = 0
    for items in list(item_lists(rnd, 40)) + [[1] * 9, [4] * 5, [64, 1, 64], [0, 0, 3]]:
        for seed in range(a.seeds // 20):
            for conv, nq in ((True, 2), (False, 2), (True, 1)):
                try:
                    run(items, seed, conv, nq, nq)
                except Violation as e:
                    bad += 1
                    print(items, seed, conv, e)
    print("violations:", bad)
    raise SystemExit(1 if bad else 0)
