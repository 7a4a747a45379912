#!/usr/bin/env python
"""RubixPipeline timings, calc_ifu against calc_dusty_ifu (all stages on the device, device-resident inputs):

    python tools/bench_dusty_pipeline.py [--particles 1000000] [--gas 1000000] [--reps 5]
"""
import argparse, copy, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rubix_b200 import core, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=1_000_000)
ap.add_argument("--gas", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
CONFIG = {
    "pipeline": {"name": "calc_ifu"},
    "logger": {"log_level": "ERROR", "log_file_path": None, "format": "%(message)s"},
    "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}, "lsf": {"sigma": 0.5}},
    "cosmology": {"name": "PLANCK15"}, "galaxy": {"dist_z": 0.1},
    "ssp": {"template": {"name": "BruzualCharlot2003"}, "method": "cubic",
            "dust": {"extinction_model": "Cardelli89", "dust_grain_density": 3.5, "Rv": 3.1}},
    "data": {"args": {"particle_type": ["stars", "gas"]}}, "b200": {"fused": True},
}
p = synthetic.bench_g(args.particles)
rng = np.random.default_rng(42)
ng = args.gas
gas = dict(coords=np.stack([rng.normal(0, 1.5, ng), rng.normal(0, 1.5, ng), rng.normal(0, 0.5, ng)], 1).astype(np.float32),
           mass=(rng.uniform(0.5, 2, ng) * 1e3).astype(np.float32), metals=rng.uniform(1e-4, 1e-2, (ng, 9)).astype(np.float32))
gas["metals"][:, 0] = 0.74


def data():
    rd = core.make_rubix_data(**p)
    rd.gas.coords, rd.gas.mass, rd.gas.metals = (torch.from_numpy(gas[k]).cuda() for k in ("coords", "mass", "metals"))
    rd.gas.velocity = torch.zeros_like(rd.gas.coords)
    return rd


out = {"particles": args.particles, "gas_cells": ng}
for name in ("calc_ifu", "calc_dusty_ifu"):
    for method in ("linear", "cubic"):
        cfg = copy.deepcopy(CONFIG)
        cfg["pipeline"]["name"] = name
        cfg["ssp"]["method"] = method
        ts = []
        for r in range(args.reps + 2):
            rd = data()
            pipe = core.RubixPipeline(cfg, data=rd)
            chain = pipe.assemble()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x = rd
            for fn in chain:
                x = fn(x)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        out[f"{name}_{method}_ms"] = float(np.median(ts[2:]))
        # one more run with a synchronisation after every stage: where the time goes
        rd = data()
        chain = core.RubixPipeline(cfg, data=rd).assemble()
        per, x = {}, rd
        for fn in chain:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x = fn(x)
            torch.cuda.synchronize()
            per[fn.__name__] = round((time.perf_counter() - t0) * 1e3, 3)
        out[f"{name}_{method}_per_stage_ms"] = per
        out[f"{name}_{method}_stages"] = [fn.__name__ for fn in chain]
print(json.dumps(out))
