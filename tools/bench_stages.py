#!/usr/bin/env python
"""Per-stage device timings (CUDA events, L2 flushed between repetitions) against the HBM roofline:
spaxel assignment, the fused particle->cube call, PSF+LSF at S=25 and S=150.

    python tools/bench_stages.py [--particles 1000000] [--reps 10]
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rubix_b200 import ops, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--conv-only", action="store_true")
args = ap.parse_args()
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, ValueError, KeyError, TypeError):
    peak = 6650.0
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def timeit(fn, reps=args.reps):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


out = {}
tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
wave = synthetic.muse_wave()
pk = ops.dev(np.load(os.path.join(ROOT, "tests", "golden", "muse_wave.npy"))[:1])  # placeholder, replaced below
from rubix_b200.telescope import gaussian_kernel_2d, lsf_kernel  # noqa: E402
pk_h, lk_h = gaussian_kernel_2d(5, 5, 0.6), lsf_kernel(0.5, 1.25)
pk, lk = ops.dev(pk_h), ops.dev(lk_h)
for S in (25, 150):
    cube = torch.rand((S, S, 3721), device="cuda")
    byts = 8 * cube.numel()
    for name, fn in (("psf_lsf_taps", lambda: ops.psf_lsf(cube, pk_h, lk_h)),
                     ("psf_only_taps", lambda: ops.convolve_psf(cube, pk_h)),
                     ("lsf_only_taps", lambda: ops.convolve_lsf(cube, lk_h)),
                     ("copy", lambda: cube.clone())):
        med, mn = timeit(fn)
        out[f"{name}_S{S}"] = {"ms_median": med, "ms_min": mn, "algorithmic_bytes": byts,
                               "achieved_gbs": byts / (med * 1e-3) / 1e9,
                               "frac_of_measured_hbm": byts / (med * 1e-3) / 1e9 / peak}
    med, mn = timeit(lambda: ops.psf_lsf(cube, pk, lk))
    byts = 8 * cube.numel()
    out[f"psf_lsf_S{S}"] = {"ms_median": med, "ms_min": mn, "algorithmic_bytes": byts,
                            "achieved_gbs": byts / (med * 1e-3) / 1e9, "frac_of_measured_hbm": byts / (med * 1e-3) / 1e9 / peak}
    med, mn = timeit(lambda: ops.convolve_psf(cube, pk))
    out[f"psf_only_S{S}"] = {"ms_median": med, "achieved_gbs": byts / (med * 1e-3) / 1e9}
    med, mn = timeit(lambda: ops.convolve_lsf(cube, lk))
    out[f"lsf_only_S{S}"] = {"ms_median": med, "achieved_gbs": byts / (med * 1e-3) / 1e9}
    del cube
if args.conv_only:
    print(json.dumps(out, indent=1))
    sys.exit(0)
n = args.particles
d = synthetic.bench_g(n)
for S in (25, 150):
    edges = ops.dev(synthetic.spatial_edges(S))
    coords, vel = ops.dev(d["coords"]), ops.dev(d["velocity"])
    mass, met, age = ops.dev(d["mass"]), ops.dev(d["metallicity"]), ops.dev(d["age"])
    med, mn = timeit(lambda: ops.spaxel_assign(coords, edges))
    out[f"spaxel_assign_S{S}"] = {"ms_median": med, "achieved_gbs": 16 * n / (med * 1e-3) / 1e9}
    pix = ops.spaxel_assign(coords, edges)
    for method in ("linear", "cubic"):
        plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method=method)
        cube = torch.empty((S, S, 3721), device="cuda")
        med, mn = timeit(lambda: ops.build_cube(plan, vel, mass, met, age, pix, S, out=cube))
        out[f"build_cube_{method}_S{S}"] = {"ms_median": med, "ms_min": mn, "particles_per_s": n / (med * 1e-3)}
        del cube
print(json.dumps(out, indent=1))
