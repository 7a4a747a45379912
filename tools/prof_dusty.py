#!/usr/bin/env python
"""Launches the one-pass dusty cube kernel and the binned form once each on 2*10^5 particles (for ncu)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rubix_b200 import dust, ops, synthetic  # noqa: E402

tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
wave = synthetic.muse_wave()
plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method="linear")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
p = synthetic.bench_g(n)
pix = ops.spaxel_assign(p["coords"], synthetic.spatial_edges(25))
av = ops.dev(np.random.default_rng(42).uniform(0, 3, n).astype(np.float32))
axav = ops.dev(dust.extinction_curve("Cardelli89", wave, 3.1))
spec = ops.scale_by_mass(ops.ssp_lookup(plan, p["metallicity"], p["age"]), p["mass"])
for _ in range(3):
    ops.build_cube_dusty(plan, spec, p["velocity"], pix, 25, av, axav)
    ops.build_cube_dusty_binned(plan, p["velocity"], p["mass"], p["metallicity"], p["age"], pix, 25, av, axav)
torch.cuda.synchronize()
