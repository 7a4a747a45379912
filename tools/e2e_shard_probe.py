#!/usr/bin/env python
"""rbx_build_cube_host on one GPU: time of a rank's host shard -> partial cube on the device against the number of
copy ranges (option host_chunks), for the shard sizes of the 2 / 4 / 8-GPU strong-scaling runs."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rubix_b200 import _lib, ops, synthetic  # noqa: E402

tpl, _ = bench.load_template()
wave = synthetic.muse_wave()
plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method="linear")
S = 25
edges = synthetic.spatial_edges(S)
full = synthetic.bench_g(5_000_000)
out = torch.empty((S, S, len(wave)), dtype=torch.float32, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for n in (1_250_000, 2_500_000, 5_000_000):
    h = {k: torch.from_numpy(np.ascontiguousarray(v[:n])).pin_memory().numpy() for k, v in full.items()}
    for chunks in (1, 2, 3, 4, 5, 6):
        _lib.set_option("host_chunks", chunks)
        ts = []
        for it in range(7):
            flush.fill_(1.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ops.build_cube_host(plan, h["coords"], h["velocity"], h["mass"], h["metallicity"], h["age"], edges, S, out=out)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print(f"n {n} ranges {chunks}: {1e3 * np.median(ts[2:]):.3f} ms", flush=True)
    _lib.set_option("host_chunks", -1)
