#!/usr/bin/env python
"""Time the REAL reference (AstroAI-Lab/rubix, JAX) on this machine's CPU for the bench.py workload.

NOT runnable in the build image or on the GPU box of this project: jax, interpax, equinox, h5py and astropy are
not installed there and there is no network.  It is shipped for any machine that has
them (``pip install rubix`` or a checkout on PYTHONPATH), so that the numbers of ``bench.py --impl reference``
(the C restatement of the same algorithm) can be put next to the genuine JAX-CPU pipeline:

    python baseline/run_reference_jax.py --particles 100000 [--method linear] [--rubix /path/to/rubix]

It builds the hot-path closures from rubix's own factories (rubix.core.ifu / psf / lsf / telescope), feeds them
the same synthetic bench-G particles as bench.py (rubix_b200.synthetic.bench_g, seed 42), jit-compiles the
chain once, and times the compiled call only (the reference's own log line includes setup and compilation:
rubix/core/pipeline.py:148-170).  Prints one JSON line shaped like bench.py's.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=100_000)
    ap.add_argument("--method", default="linear", choices=["linear", "cubic"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--rubix", default=None, help="path of a rubix checkout (else the installed package)")
    args = ap.parse_args()
    if args.rubix:
        sys.path.insert(0, args.rubix)
    try:
        import jax
        import jax.numpy as jnp
        from rubix.core.data import Galaxy, GasData, RubixData, StarsData, get_reshape_data
        from rubix.core.ifu import (get_calculate_datacube, get_calculate_spectra, get_doppler_shift_and_resampling,
                                    get_scale_spectrum_by_mass)
        from rubix.core.lsf import get_convolve_lsf
        from rubix.core.psf import get_convolve_psf
        from rubix.core.telescope import get_filter_particles, get_spaxel_assignment
    except ImportError as e:  # the expected outcome in this project's images
        print(json.dumps({"impl": "reference-jax", "unavailable": f"{type(e).__name__}: {e}"}))
        return
    from rubix_b200 import synthetic

    config = {
        "pipeline": {"name": "calc_ifu"},
        "logger": {"log_level": "WARNING", "log_file_path": None,
                   "format": "%(asctime)s - %(name)s - %(levelname)s - %(message)s"},
        "data": {"name": "IllustrisAPI", "args": {"particle_type": ["stars"]}, "load_galaxy_args": {}, "subset": {"use_subset": False}},
        "simulation": {"name": "IllustrisTNG", "args": {"path": "unused"}},
        "output_path": "unused",
        "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}, "lsf": {"sigma": 0.5},
                      "noise": {"signal_to_noise": 1, "noise_distribution": "normal"}},
        "cosmology": {"name": "PLANCK15"},
        "galaxy": {"dist_z": 0.1, "rotation": {"type": "face-on"}},
        "ssp": {"template": {"name": "BruzualCharlot2003"}, "method": args.method},
    }
    d = synthetic.bench_g(args.particles, seed=42)
    rd = RubixData(Galaxy(), StarsData(), GasData())
    rd.galaxy.redshift, rd.galaxy.center, rd.galaxy.halfmassrad_stars = 0.1, jnp.zeros(3), 1.5
    rd.stars.coords, rd.stars.velocity = jnp.array(d["coords"]), jnp.array(d["velocity"])
    rd.stars.mass, rd.stars.metallicity, rd.stars.age = jnp.array(d["mass"]), jnp.array(d["metallicity"]), jnp.array(d["age"])
    stages = [get(config) for get in (get_filter_particles, get_spaxel_assignment, get_calculate_spectra, get_reshape_data,
                                      get_scale_spectrum_by_mass, get_doppler_shift_and_resampling,
                                      get_calculate_datacube, get_convolve_psf, get_convolve_lsf)]

    def chain(x):
        for f in stages:
            x = f(x)
        return x

    fn = jax.jit(chain)
    out = fn(rd)
    jax.block_until_ready(out.stars.datacube)          # compile + first run, not timed
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = fn(rd)
        jax.block_until_ready(out.stars.datacube)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    print(json.dumps({"impl": "reference-jax", "metric": "particles/s binned into MUSE datacube", "value": args.particles / t,
                      "unit": "particles/s", "ms_per_step": t * 1e3, "particles": args.particles, "method": args.method,
                      "jax_backend": jax.default_backend(), "devices": [str(x) for x in jax.devices()],
                      "host_threads": os.cpu_count()}))


if __name__ == "__main__":
    main()
