#!/usr/bin/env python
"""List-scheduling model of the cube kernel's work queue (CPU only): how well do the work items of a bench-G shard
balance over the 888 warp pairs of one B200 (148 SMs x 6 pairs)?

Items are cut as segment_kernel cuts them (cut_spaxel: bulk items of psub particles, the last eighth of a spaxel in
items a quarter of that size, all small items behind all bulk items), cost = particles + E (the expansion of the cells,
in particle equivalents), greedy assignment in queue order.  Prints ideal / makespan per psub.

    python tools/queue_model.py [n_particles ...]
"""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rubix_b200 import synthetic  # noqa: E402

P, E = 888, 14.0


def counts_for(d, n, S=25):
    c = d["coords"][:n]
    e = synthetic.spatial_edges(S)
    xi = np.clip(np.searchsorted(e, c[:, 0], "right") - 1, 0, S - 1)
    yi = np.clip(np.searchsorted(e, c[:, 1], "right") - 1, 0, S - 1)
    ok = (c[:, 0] >= e[0]) & (c[:, 0] <= e[-1]) & (c[:, 1] >= e[0]) & (c[:, 1] <= e[-1])
    return np.bincount((xi + S * yi)[ok], minlength=S * S)


def cut(c, psub, small_shift=2, tail_shift=3):
    psmall = max(32, psub >> small_shift)
    tail = 0
    if c > psub:
        tail = min(c, (((c >> tail_shift) + psmall - 1) // psmall) * psmall)
    bulk = c - tail
    return ([min(psub, bulk - q * psub) for q in range((bulk + psub - 1) // psub)],
            [min(psmall, tail - q * psmall) for q in range((tail + psmall - 1) // psmall)])


def makespan(queue):
    h = [0.0] * P
    heapq.heapify(h)
    for it in queue:
        heapq.heappush(h, heapq.heappop(h) + it + E)
    return max(h)


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [1_000_000, 1_250_000, 2_500_000, 5_000_000, 10_000_000]
    d = synthetic.bench_g(max(sizes))
    for n in sizes:
        cnt = counts_for(d, n)
        ideal = cnt.sum() / P
        row = []
        for psub in (256, 384, 512, 644, 768, 1024, 2048):
            b, s = [], []
            for c in cnt:
                x, y = cut(int(c), psub)
                b += x
                s += y
            row.append(f"{psub}: {ideal / makespan(b + s):.3f} ({len(b) + len(s)} items)")
        print(n, " ".join(row))
