#!/usr/bin/env python
"""One rank's device work of the 8-GPU large-FOV step (config 4) on ONE GPU, for a launch list: 1.25e6 particles onto
150 x 150 spaxels, slab-major partial cube (8 slabs, 12-channel halo), PSF + LSF of one slab.  The reduce-scatter
itself needs the other ranks and is not part of this."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rubix_b200 import ops, synthetic  # noqa: E402

class A: pass
args = A(); args.spaxels = 150; args.gpus = 8; args.galaxies = 1; args.particles = 10_000_000; args.weak = False; args.method = "linear"
tpl, _ = bench.load_template()
wave = synthetic.muse_wave()
plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method="linear")
edges = ops.dev(bench.spatial_edges(args))
d = bench.galaxy(args, 1_250_000)
pk, lk = bench.host_kernels()
t = {k: ops.dev(v) for k, v in d.items()}
wslab, ws = ops.slab_geometry(len(wave), 8, 12)
slabs = torch.empty((8, 150 * 150, ws), dtype=torch.float32, device="cuda")
for _ in range(4):
    ops.assign_build_cube_slabs(plan, t["coords"], edges, t["velocity"], t["mass"], t["metallicity"], t["age"], 150, 8, 12, out=slabs)
    out = ops.psf_lsf_own_slab(slabs[0], 150, len(wave), 0, 8, pk, lk, 12)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
ops.assign_build_cube_slabs(plan, t["coords"], edges, t["velocity"], t["mass"], t["metallicity"], t["age"], 150, 8, 12, out=slabs)
out = ops.psf_lsf_own_slab(slabs[0], 150, len(wave), 0, 8, pk, lk, 12)
b.record(); torch.cuda.synchronize()
print("one rank's device work without the collective: %.3f ms" % a.elapsed_time(b))
