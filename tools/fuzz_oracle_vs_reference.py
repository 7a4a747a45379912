"""Differential fuzzing of the oracle against the REFERENCE'S OWN SOURCE (build container only).

tools/refshim.py runs the reference files unchanged on numpy; this script throws seeded random inputs -- including the
degenerate ones the fixed vectors do not hold (bands that miss the spectrum, zero and negative spectra, one-channel
grids, particles on edges, empty spaxels, no particle inside the half-mass radius, kernels as large as the image) -- at
both and reports every disagreement beyond float64 rounding.  Exit status 1 if there is any.

    python tools/fuzz_oracle_vs_reference.py [cases per function, default 300] [seed, default 0]
"""

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refshim  # noqa: E402
from make_ref_golden import reference_modules  # noqa: E402
from oracle import rubix_oracle as orc  # noqa: E402


def close(a, b, tol=1e-11):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    a, b = np.nan_to_num(a, nan=0.0), np.nan_to_num(b, nan=0.0)
    scale = max(np.abs(b).max(initial=0.0), 1e-300)
    return bool(np.abs(a - b).max(initial=0.0) <= tol * scale)


def grid(rng, n, lo, hi):
    g = np.sort(rng.uniform(lo, hi, n))
    return g + np.arange(n) * 1e-9 * (hi - lo)          # strictly increasing


def fuzz(cases, seed):
    m = reference_modules()
    ifu, tel, kern, psf, lsf, align, noise = (m[k] for k in ("ifu", "tel", "kern", "psf", "lsf", "align", "noise"))
    rng = np.random.default_rng(seed)
    bad = []

    def check(name, k, got, want, tol=1e-11, exact=False):
        ok = np.array_equal(np.asarray(got), np.asarray(want)) if exact else close(got, want, tol)
        if not ok:
            bad.append((name, k))

    for k in range(cases):
        # ---- a4: resample_spectrum ------------------------------------------------------------------------------
        L, W = int(rng.integers(2, 40)), int(rng.integers(1, 60))
        lam = grid(rng, L, 1000.0, 2000.0)
        mode = k % 6
        t_lo, t_hi = [(1100, 1900), (500, 900), (2100, 2500), (800, 1500), (1500, 2600), (900, 2100)][mode]
        t = grid(rng, W, t_lo, t_hi)
        s = rng.uniform(0, 5, L)
        if k % 7 == 0:
            s[:] = 0.0
        if k % 11 == 0:
            s -= 2.5                                     # negative flux: the sums can cancel
        if k % 13 == 0:
            t[rng.integers(0, W)] = lam[rng.integers(0, L)]   # a channel exactly on a knot
        with np.errstate(all="ignore"):
            check("resample_spectrum", k, orc.resample_spectrum(s, lam, t), ifu.resample_spectrum(s, lam, t))
        check("calculate_diff", k, orc.calculate_diff(t), ifu.calculate_diff(t), exact=True)
        # ---- a0: spaxel assignment + mask -----------------------------------------------------------------------
        nb = int(rng.integers(1, 12))
        edges = grid(rng, nb + 1, -5.0, 5.0).astype(np.float32)
        c = rng.uniform(-7, 7, (50, 3)).astype(np.float32)
        c[:nb + 1, 0] = edges
        c[nb + 1:2 * nb + 2, 1] = edges[:min(nb + 1, 50 - nb - 1)]
        check("square_spaxel_assignment", k, orc.square_spaxel_assignment(c, edges),
              np.asarray(tel.square_spaxel_assignment(c, edges)), exact=True)
        check("mask_particles_outside_aperture", k, orc.mask_particles_outside_aperture(c, edges),
              np.asarray(tel.mask_particles_outside_aperture(c, edges)), exact=True)
        # ---- a3 -------------------------------------------------------------------------------------------------
        vel = rng.normal(0, 500, (5, 3))
        for direction in "xyz":
            check("velocity_doppler_shift", k, orc.velocity_doppler_shift(lam, vel, direction, dtype=np.float64),
                  ifu.velocity_doppler_shift(lam, vel, direction), 1e-15)
        # ---- a5 -------------------------------------------------------------------------------------------------
        S = int(rng.integers(1, 5))
        ids = rng.integers(0, S * S + 3, 30)
        spec = rng.uniform(0, 1, (30, 7))
        check("calculate_cube", k, orc.calculate_cube(spec, ids, S), ifu.calculate_cube(spec, ids, S), exact=True)
        # ---- a6 / a7 --------------------------------------------------------------------------------------------
        mk, nk = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        H, Wd = mk + int(rng.integers(0, 6)), nk + int(rng.integers(0, 6))      # the image is at least the kernel
        cube = rng.uniform(0, 1, (H, Wd, 4))
        kernel = rng.uniform(-1, 1, (mk, nk))
        check("apply_psf", k, orc.apply_psf(cube, kernel), psf.apply_psf(cube, kernel))
        sigma = float(rng.uniform(0.2, 3.0))
        check("gaussian_kernel_2d", k, orc.gaussian_kernel_2d(mk, nk, sigma, dtype=np.float64),
              kern.gaussian_kernel_2d(mk, nk, sigma), 1e-14)
        wres, ext = float(rng.uniform(0.5, 3.0)), int(rng.integers(1, 14))
        # the reference builds its taps with arange(-ext * wres, ext * wres + wres, wres): 2 ext + 1 of them unless
        # rounding adds one (then its own slice no longer returns the input length, and the oracle mirrors that)
        cube = rng.uniform(0, 1, (2, 3, 2 * ext + 1 + int(rng.integers(0, 30))))
        want = lsf.apply_lsf(cube, sigma, wres, ext) if len(lsf._get_kernel(sigma, wres, factor=ext)) == 2 * ext + 1 \
            else None
        if want is not None:
            check("apply_lsf", k, orc.apply_lsf(cube, sigma, wres, ext), want)
            check("lsf_kernel", k, orc.lsf_kernel(sigma, wres, ext, dtype=np.float64),
                  lsf._get_kernel(sigma, wres, factor=ext), 1e-14)
        # ---- rotate_galaxy / S2N --------------------------------------------------------------------------------
        n = int(rng.integers(3, 60))
        pos, masses = rng.normal(0, 2, (n, 3)), rng.uniform(0.1, 2, n)
        radius = [0.01, 1.0, 3.0, 100.0][k % 4]          # nobody inside ... everybody inside
        check("moment_of_inertia_tensor", k, orc.moment_of_inertia_tensor(pos, masses, radius),
              align.moment_of_inertia_tensor(pos, masses, radius))
        ang = rng.uniform(-180, 180, 3)
        check("euler_rotation_matrix", k, orc.euler_rotation_matrix(*ang), align.euler_rotation_matrix(*ang), 1e-14)
        nc = rng.uniform(0, 1, (4, 3, 5))
        if k % 3 == 0:
            nc[1, 2] = 0.0
        with np.errstate(all="ignore"):
            check("calculate_S2N", k, orc.calculate_S2N(nc, 20.0), noise.calculate_S2N(nc, 20.0))

    # ---- dust: apply_spaxel_extinction --------------------------------------------------------------------------
    from types import SimpleNamespace as NS
    for f in ("helpers", "generic_models", "dust_baseclasses", "extinction_models"):
        refshim.load(f"rubix/spectra/dust/{f}.py")
    de = refshim.load("rubix/spectra/dust/dust_extinction.py")
    for k in range(max(cases // 3, 1)):
        S = int(rng.integers(1, 6))
        ng, ns, W = int(rng.integers(0, 25)), int(rng.integers(1, 15)), 6
        if ng == 0:
            continue          # the reference indexes the first gas cell unconditionally (jnp.interp on empty arrays)
        gz = rng.normal(0, 1, ng).astype(np.float32)
        sz = rng.normal(0, 1.5, ns).astype(np.float32)
        if k % 4 == 0:
            sz[0] = gz[0]                                # a star exactly at a gas cell
        gp, sp = rng.integers(0, S, ng).astype(np.int32), rng.integers(0, S, ns).astype(np.int32)
        gmass = rng.uniform(1e4, 1e6, ng).astype(np.float32)
        metals = rng.uniform(1e-3, 0.05, (ng, 9)).astype(np.float32)
        metals[:, 0] = 0.74
        spectra = rng.uniform(0.5, 2, (ns, W))
        wave = np.linspace(4000.0, 9000.0, W).astype(np.float32)
        f64 = lambda a: np.asarray(a, dtype=np.float64)
        gc = np.concatenate([np.zeros((ng, 2)), f64(gz)[:, None]], axis=1)
        sc = np.concatenate([np.zeros((ns, 2)), f64(sz)[:, None]], axis=1)
        for model in ("Cardelli89", "Gordon23"):
            rd = NS(gas=NS(coords=gc[None], pixel_assignment=gp[None], metals=f64(metals)[None], mass=f64(gmass)[None]),
                    stars=NS(coords=sc[None], pixel_assignment=sp[None], mass=np.ones((1, ns)), spectra=spectra[None]))
            cfg = {"ssp": {"dust": {"extinction_model": model, "Rv": 3.1, "dust_grain_density": 3.5}}}
            want = np.asarray(de.apply_spaxel_extinction(cfg, rd, f64(wave), S, np.float64(0.2)))[0]
            got, _ = orc.apply_spaxel_extinction(spectra, wave, gz, gp, gmass, metals, sz, sp, S, 0.2,
                                                 {"extinction_model": model, "Rv": 3.1, "dust_grain_density": 3.5,
                                                  "dust_to_gas_model": "broken power law fit", "Xco": "Z"})
            check("apply_spaxel_extinction " + model, k, got, want, 1e-10)
    return bad


def float32_order(cases, seed):
    """The oracle's float32 mode claims the reference's OPERATION ORDER.  With numpy's float32 arithmetic on both sides
    -- the reference's source expressions evaluated as written, ``jnp.interp`` stood in by jax's own formula in the
    input dtype -- the cosmological and velocity Doppler shifts, resample_spectrum and calculate_cube must then agree
    BIT FOR BIT (what XLA fuses or contracts on top of that order cannot be known here)."""
    import jax.numpy as jnp      # the stand-in

    def interp_in_dtype(x, xp, fp, left=None, right=None, period=None):
        x, xp, fp = np.asarray(x), np.asarray(xp), np.asarray(fp)
        i = np.clip(np.searchsorted(xp, x, side="right"), 1, len(xp) - 1)
        df, dx, delta = fp[i] - fp[i - 1], xp[i] - xp[i - 1], x - xp[i - 1]
        dx0 = np.abs(dx) <= np.spacing(np.finfo(xp.dtype).eps)
        with np.errstate(all="ignore"):
            f = np.where(dx0, fp[i - 1], fp[i - 1] + (delta / np.where(dx0, 1, dx)) * df)
        f = np.where(x < xp[0], fp[0], f)
        return np.where(x > xp[-1], fp[-1], f)

    keep = jnp.interp
    jnp.interp = interp_in_dtype
    try:
        sys.modules.pop("rubix.spectra.ifu", None)
        ifu = refshim.load("rubix/spectra/ifu.py")
        tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
        wave = np.load(os.path.join(ROOT, "tests", "golden", "muse_wave.npy"))
        rng = np.random.default_rng(seed)
        f = np.float32
        bad = []
        lam_z = ifu.cosmological_doppler_shift(0.1, tpl["wavelength"])
        if lam_z.dtype != f or not np.array_equal(lam_z, orc.cosmological_doppler_shift(0.1, tpl["wavelength"])):
            bad.append(("cosmological_doppler_shift", 0))
        vel = rng.normal(0, 300, (cases, 3)).astype(f)
        sh, so = ifu.velocity_doppler_shift(lam_z, vel, "z"), orc.velocity_doppler_shift(lam_z, vel, "z")
        if sh.dtype != f or not np.array_equal(sh, so):
            bad.append(("velocity_doppler_shift", 0))
        rows = tpl["flux"].reshape(-1, 842)[rng.integers(0, 1326, cases)] * rng.uniform(0.5, 1.5, cases).astype(f)[:, None]
        res = []
        for k in range(cases):
            with np.errstate(all="ignore"):
                a, b = ifu.resample_spectrum(rows[k], sh[k], wave), orc.resample_spectrum(rows[k], so[k], wave)
            res.append(a)
            if a.dtype != f or not np.array_equal(a, b):
                bad.append(("resample_spectrum", k))
        ids = rng.integers(0, 30, cases)
        if not np.array_equal(ifu.calculate_cube(np.stack(res), ids, 5), orc.calculate_cube(np.stack(res), ids, 5)):
            bad.append(("calculate_cube", 0))
    finally:
        jnp.interp = keep
        sys.modules.pop("rubix.spectra.ifu", None)
    return bad


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bad = fuzz(cases, seed)
        bad32 = float32_order(min(cases, 200), seed)
    print(f"float32 operation order, {min(cases, 200)} particles: " +
          ("oracle float32 mode == reference source in float32, bit for bit" if not bad32 else f"differs: {bad32[:10]}"))
    bad = bad + bad32
    names = sorted({b[0] for b in bad})
    print(f"fuzz: {cases} cases per function, seed {seed}: " +
          ("oracle == reference source everywhere" if not bad else f"{len(bad)} disagreement(s) in {names}: {bad[:12]}"))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
