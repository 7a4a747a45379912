import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import c_oracle
from rubix_b200 import ops, synthetic
from helpers import well_conditioned
tpl = np.load("rubix_b200/templates/bc03lr_f32.npz")
wave = synthetic.muse_wave(); edges = synthetic.spatial_edges(25)
gen = sys.argv[1] if len(sys.argv) > 1 else "bench_u"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
method = sys.argv[3] if len(sys.argv) > 3 else "linear"
d = well_conditioned(getattr(synthetic, gen)(n), np.float32(1.1) * tpl["wavelength"], wave)
plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method=method)
coords = ops.dev(d["coords"]); mass, met, age = ops.dev(d["mass"]).clone(), ops.dev(d["metallicity"]).clone(), ops.dev(d["age"]).clone()
ops.filter_particles(coords, edges, mass, met, age)
pix = ops.spaxel_assign(coords, edges)
out = ops.build_cube(plan, d["velocity"], mass, met, age, pix, 25).cpu().numpy().astype(np.float64)
ref = c_oracle.particles_to_cube(d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"], edges, 25, tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method=method, dtype=np.float64, n_threads=8)
err = np.abs(out - ref)
print("max err", err.max(), "max ref", ref.max())
sp = err.max(-1).reshape(-1)
bad = np.argsort(-sp)[:8]
cnt = np.bincount(pix.cpu().numpy(), minlength=625)
for s in bad:
    w = err.reshape(625, -1)[s].argmax()
    e = err.reshape(625, -1)[s]
    print(f"spaxel {s} count {cnt[s]} maxerr {sp[s]:.3e} at chan {w}  nbad(>1e-5*max) {(e > 1e-5 * ref.max()).sum()} first/last bad {np.nonzero(e > 1e-5*ref.max())[0][[0,-1]] if (e > 1e-5*ref.max()).any() else None}")
print("spaxels with error:", (sp > 1e-5 * ref.max()).sum())
