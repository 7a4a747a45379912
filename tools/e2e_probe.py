#!/usr/bin/env python
"""Where the end-to-end time of rbx_pipeline_host goes: PCIe copy times alone, device step alone, and the
host call with 1..4 particle ranges (RBX_HOST_CHUNKS)."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rubix_b200 import ops, synthetic  # noqa: E402
from rubix_b200.telescope import gaussian_kernel_2d, lsf_kernel  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
wave = synthetic.muse_wave(); edges = synthetic.spatial_edges(25)
d = synthetic.bench_g(n)
plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method="linear")
pk, lk = gaussian_kernel_2d(5, 5, 0.6), lsf_kernel(0.5, 1.25)
pin = {k: torch.from_numpy(v).pin_memory() for k, v in d.items()}
dev = {k: torch.empty_like(v, device="cuda") for k, v in pin.items()}
hcube = torch.empty((25, 25, 3721)).pin_memory(); dcube = torch.empty((25, 25, 3721), device="cuda")
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
out = {}
out["h2d_ms"] = timed(lambda: [dev[k].copy_(pin[k], non_blocking=True) for k in pin])
out["d2h_cube_ms"] = timed(lambda: hcube.copy_(dcube, non_blocking=True))
def dev_step():
    pix = ops.filter_and_assign(dev["coords"], ops.dev(edges))
    c = ops.build_cube(plan, dev["velocity"], dev["mass"], dev["metallicity"], dev["age"], pix, 25)
    return ops.psf_lsf(c, pk, lk)
out["device_step_ms"] = timed(dev_step)
hn = {k: v.numpy() for k, v in pin.items()}
for c in (1, 2, 3, 4):
    os.environ["RBX_HOST_CHUNKS"] = str(c)
    out[f"host_call_chunks{c}_ms"] = timed(lambda: ops.pipeline_host(plan, hn["coords"], hn["velocity"], hn["mass"], hn["metallicity"], hn["age"], edges, 25, pk, lk, out=hcube.numpy()))
print(json.dumps(out, indent=1))
