#!/usr/bin/env python
"""Device timings of the dust variant (calc_dusty_ifu): per-star A_V, the stage kernel, and the one-pass
resample + extinction + cube kernel against the staged kernels (CUDA events, L2 flushed between repetitions).

    python tools/bench_dusty.py [--particles 1000000] [--gas 1000000] [--staged 200000] [--reps 5]
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rubix_b200 import dust, ops, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=1_000_000)
ap.add_argument("--gas", type=int, default=1_000_000)
ap.add_argument("--staged", type=int, default=200_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--method", default="linear")
args = ap.parse_args()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def timeit(fn, reps=args.reps):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
wave = synthetic.muse_wave()
edges = synthetic.spatial_edges(25)
plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method=args.method)
n, ng = args.particles, args.gas
p = synthetic.bench_g(n)
coords, vel = ops.dev(p["coords"]), ops.dev(p["velocity"])
mass, met, age = ops.dev(p["mass"]), ops.dev(p["metallicity"]), ops.dev(p["age"])
pix = ops.spaxel_assign(coords, ops.dev(edges))
rng = np.random.default_rng(42)
gc = ops.dev(np.stack([rng.normal(0, 1.5, ng), rng.normal(0, 1.5, ng), rng.normal(0, 0.5, ng)], 1).astype(np.float32))
gpix = ops.spaxel_assign(gc, ops.dev(edges))
gmass = ops.dev((rng.uniform(0.5, 2, ng) * 1e3).astype(np.float32))
metals = rng.uniform(1e-4, 1e-2, (ng, 9)).astype(np.float32)
metals[:, 0] = 0.74
metals = ops.dev(metals)
dtg = dust.dust_to_gas_parameters("broken power law fit", "Z")
axav = ops.dev(dust.extinction_curve("Cardelli89", wave, 3.1))
out = {"particles": n, "gas_cells": ng, "method": args.method}
av = ops.dust_av(gc, gpix, gmass, metals, coords, pix, 625, dtg, dust.extinction_constant(3.5), 0.145)
out["av_median"] = float(av.median())
out["dust_av_ms"] = timeit(lambda: ops.dust_av(gc, gpix, gmass, metals, coords, pix, 625, dtg, dust.extinction_constant(3.5), 0.145))

spec = ops.scale_by_mass(ops.ssp_lookup(plan, met, age), mass)
out["ssp_lookup_scale_ms"] = timeit(lambda: ops.scale_by_mass(ops.ssp_lookup(plan, met, age), mass))
out["one_pass_dusty_ms"] = timeit(lambda: ops.build_cube_dusty(plan, spec, vel, pix, 25, av, axav))
out["one_pass_nodust_ms"] = timeit(lambda: ops.build_cube_dusty(plan, spec, vel, pix, 25))
out["one_pass_particles_per_s"] = n / out["one_pass_dusty_ms"] * 1e3
out["binned_dusty_ms"] = timeit(lambda: ops.build_cube_dusty_binned(plan, vel, mass, met, age, pix, 25, av, axav))
out["binned_particles_per_s"] = n / out["binned_dusty_ms"] * 1e3
_a = ops.build_cube_dusty_binned(plan, vel, mass, met, age, pix, 25, av, axav)
_b = ops.build_cube_dusty(plan, spec, vel, pix, 25, av, axav)
out["binned_vs_one_pass_rel_to_max"] = float((_a - _b).abs().max() / _b.abs().max())
out["av_range"] = [float(av.min()), float(av.max())]
out["fused_no_dust_build_cube_ms"] = timeit(lambda: ops.build_cube(plan, vel, mass, met, age, pix, 25))

m = min(args.staged, n)
sp_m, vel_m, pix_m, av_m = spec[:m].contiguous(), vel[:m].contiguous(), pix[:m].contiguous(), av[:m].contiguous()


def staged():
    r = ops.doppler_resample(plan, sp_m, vel_m)
    r = ops.apply_extinction(r, av_m, axav, out=r)
    return ops.segment_sum(r, pix_m, 625)


out["staged_particles"] = m
out["staged_ms"] = timeit(staged)
out["staged_particles_per_s"] = m / out["staged_ms"] * 1e3
out["one_pass_same_particles_ms"] = timeit(lambda: ops.build_cube_dusty(plan, sp_m, vel_m, pix_m, 25, av_m, axav))
a = ops.build_cube_dusty(plan, sp_m, vel_m, pix_m, 25, av_m, axav)
b = staged().reshape(25, 25, -1)
out["one_pass_vs_staged_rel_to_max"] = float((a - b).abs().max() / b.abs().max())
print(json.dumps(out))
