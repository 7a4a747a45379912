"""Shared helpers of the GPU parity tests."""

import numpy as np


def well_conditioned(data, lamz, wave, direction=2, tol=1e-2):
    """Drop particles sitting on a knife edge of the reference's own formula.

    rubix/spectra/ifu.py:241-244 masks the SSP knots with ``tmin <= lam' <= tmax``; a knot within a
    float32 ulp (~5e-4 A) of either end flips in or out of ``total_lum`` depending on how lam_z * d
    was rounded, which changes that particle's whole spectrum by ~0.5 % (one 22 A bin of 4650 A).
    float32 and float64 evaluations of the reference disagree there (oracle f32 vs f64: 3.6e-5 of
    the cube maximum for one such particle in bench_g(20000)), so those particles cannot pin
    anything; they are removed from the parity inputs (about 1 in 10^4)."""
    d = np.exp(data["velocity"][:, direction].astype(np.float64) / 299792.458)
    lz = lamz.astype(np.float64)
    lo = np.searchsorted(lz, wave[0] / d.max()) - 2
    hi = np.searchsorted(lz, wave[-1] / d.min()) + 2
    keep = np.empty(len(d), dtype=bool)
    for a in range(0, len(d), 200_000):   # in blocks: 10^6 particles x ~250 knots of float64 would be 2 GB
        x = lz[None, max(lo, 0):hi] * d[a:a + 200_000, None]
        dist = np.minimum(np.abs(x - float(wave[0])).min(1), np.abs(x - float(wave[-1])).min(1))
        keep[a:a + 200_000] = dist > tol
    return {k: v[keep] for k, v in data.items()}


def cube_close(out, ref, tag="", rtol_max=5e-6):
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert out.shape == ref.shape
    assert np.isfinite(out).all(), f"{tag}: non-finite values in the cube"
    err = np.abs(out - ref).max()
    mx = np.abs(ref).max()
    tot = np.abs(ref).sum()
    print(f"[{tag}] max|d|={err:.3e} max|ref|={mx:.3e} rel_to_max={err / max(mx, 1e-300):.3e} "
          f"rel_to_total={err / max(tot, 1e-300):.3e}")
    # north star: rtol 1e-5 per voxel relative to the cube's total flux
    assert err <= 1e-5 * tot, f"{tag}: north-star tolerance violated"
    assert err <= rtol_max * mx, f"{tag}: max error {err:.3e} > {rtol_max} * {mx:.3e}"


def build_c_host(tmp_path):
    """Compile tests/abi/c_host_pipeline.c (plain C99, only include/rubix_b200.h) against the in-tree library."""
    import os
    import subprocess
    from rubix_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_host_pipeline")
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "abi", "c_host_pipeline.c"), "-o", exe, "-L", libdir, "-lrubix_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def write_c_host_inputs(d, tpl, wave, particles, edges, S, psf, lsf, method):
    """The raw arrays c_host_pipeline reads (float32 / int32, C order)."""
    import os
    f32 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    n = len(particles["mass"])
    dims = np.array([len(tpl["metallicity"]), len(tpl["age"]), len(tpl["wavelength"]), len(wave), n, len(edges), S,
                     psf.shape[0], psf.shape[1], len(lsf), {"linear": 0, "cubic": 1}[method]], dtype=np.int32)
    dims.tofile(os.path.join(d, "dims.i32"))
    for name, a in (("metallicity", tpl["metallicity"]), ("age", tpl["age"]), ("wavelength", tpl["wavelength"]),
                    ("flux", tpl["flux"]), ("wave", wave), ("coords", particles["coords"]),
                    ("velocity", particles["velocity"]), ("mass", particles["mass"]), ("met", particles["metallicity"]),
                    ("age_p", particles["age"]), ("edges", edges), ("psf", psf), ("lsf", lsf)):
        f32(a).tofile(os.path.join(d, name + ".f32"))
