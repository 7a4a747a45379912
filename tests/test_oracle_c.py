"""The C restatement (oracle/rubix_oracle.c) must agree with the numpy restatement, which is the
one pinned to the reference's known-answer tests (tests/test_oracle_golden.py)."""

import numpy as np
import pytest

from oracle import c_oracle
from oracle import rubix_oracle as orc


def _args(bc03, muse_wave, s, n):
    edges = np.linspace(-4.7619, 4.7619, 26).astype(np.float32)
    return (s["coords"][:n], s["velocity"][:n], s["mass"][:n], s["metallicity"][:n], s["age"][:n],
            edges, 25, bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1)


@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_c_oracle_matches_numpy_f64(bc03, muse_wave, tng_subset, method):
    args = _args(bc03, muse_wave, tng_subset, 300)
    ref, idx = orc.particles_to_cube(*args, method=method, dtype=np.float64)
    out = c_oracle.particles_to_cube(*args, method=method, dtype=np.float64)
    assert np.abs(out - ref).max() <= 1e-11 * np.abs(ref).max()
    out3 = c_oracle.particles_to_cube(*args, method=method, dtype=np.float64, n_threads=3)
    assert np.abs(out3 - ref).max() <= 1e-11 * np.abs(ref).max()
    i2, m2 = c_oracle.spaxel_assign(args[0], args[5])
    assert np.array_equal(i2, idx)
    assert np.array_equal(m2, orc.mask_particles_outside_aperture(args[0], args[5]))


@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_c_oracle_f32_tracks_f64(bc03, muse_wave, tng_subset, method):
    args = _args(bc03, muse_wave, tng_subset, 2000)
    c64 = c_oracle.particles_to_cube(*args, method=method, dtype=np.float64, n_threads=4)
    c32 = c_oracle.particles_to_cube(*args, method=method, dtype=np.float32, n_threads=4)
    assert c64.max() > 0
    assert np.abs(c32 - c64).max() <= 5e-6 * np.abs(c64).max()


def test_c_oracle_no_filter_and_direction(bc03, muse_wave, tng_subset):
    args = _args(bc03, muse_wave, tng_subset, 200)
    for kw in (dict(apply_filter=False), dict(direction="x")):
        ref, _ = orc.particles_to_cube(*args, method="linear", dtype=np.float64, **kw)
        out = c_oracle.particles_to_cube(*args, method="linear", dtype=np.float64, **kw)
        assert np.abs(out - ref).max() <= 1e-11 * np.abs(ref).max()


def test_c_psf_lsf_match_numpy():
    rng = np.random.default_rng(5)
    cube = rng.random((7, 9, 300))
    k2 = orc.gaussian_kernel_2d(5, 5, 0.6, dtype=np.float64)
    assert np.allclose(c_oracle.apply_psf(cube, k2), orc.apply_psf(cube, k2), atol=1e-14)
    k2 = rng.random((4, 3))
    assert np.allclose(c_oracle.apply_psf(cube, k2), orc.apply_psf(cube, k2), atol=1e-13)
    k1 = orc.lsf_kernel(0.5, 1.25, dtype=np.float64)
    assert np.allclose(c_oracle.apply_lsf(cube, k1), orc.apply_lsf(cube, 0.5, 1.25), atol=1e-14)
