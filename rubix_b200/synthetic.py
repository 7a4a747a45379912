"""Synthetic star particles and SSP templates for tests and bench.py.

``bench_u`` / ``bench_g`` follow SURVEY.md section 8(d); the RNG seed 42 is the reference's own
(rubix/core/data.py:565, rubix/debug.py:10).  ``synthetic_ssp`` builds a BC03lr-shaped template
(6 x 221 x 842, same wavelength/age/metallicity axes layout) for boxes where the golden fixture is
not wanted.
"""

from __future__ import annotations

import numpy as np

MUSE = dict(fov=5.0, spatial_res=0.2, wave_range=[4700.15, 9351.4], wave_res=1.25)


def muse_wave() -> np.ndarray:
    """telescope.wave_seq for MUSE (rubix/telescope/telescopes.yaml:2-10, telescope/utils.py:53)."""
    return np.arange(MUSE["wave_range"][0], MUSE["wave_range"][1], MUSE["wave_res"], dtype=np.float32)


def spatial_edges(num_spaxels: int = 25, half_aperture: float = 4.7619) -> np.ndarray:
    """Bin edges as get_spatial_bin_edges would hand them over for dist_z = 0.1 / PLANCK15
    (aperture 9.5238 kpc): ``num_spaxels + 1`` float32 edges."""
    size = np.float32(2 * half_aperture / num_spaxels)
    return (np.float32(-half_aperture) + size * np.arange(num_spaxels + 1, dtype=np.float32)).astype(np.float32)


def bench_u(n: int, half: float = 4.7, seed: int = 42):
    rng = np.random.default_rng(seed)
    coords = rng.uniform(-half, half, (n, 3)).astype(np.float32)
    vel = rng.uniform(-100, 100, (n, 3)).astype(np.float32)
    met = rng.uniform(1e-4, 0.05, n).astype(np.float32)
    age = rng.uniform(0.0, 10.30, n).astype(np.float32)
    mass = np.ones(n, dtype=np.float32)
    return dict(coords=coords, velocity=vel, mass=mass, metallicity=met, age=age)


def bench_g(n: int, sigma_kpc: float = 1.5, seed: int = 42, scale: float = 1.0):
    """Galaxy-like: centrally concentrated, ~0.3 % outside the aperture, 2 % with Z above the grid."""
    rng = np.random.default_rng(seed)
    coords = (rng.normal(0, sigma_kpc * scale, (n, 3))).astype(np.float32)
    vel = np.zeros((n, 3), dtype=np.float32)
    vel[:, 2] = rng.normal(0, 150, n)
    vel[:, :2] = rng.normal(0, 150, (n, 2))
    logz = rng.uniform(-4, -1.3, n)
    met = (10.0 ** logz).astype(np.float32)
    hi = rng.random(n) < 0.02
    met[hi] = rng.uniform(0.051, 0.1, hi.sum()).astype(np.float32)
    age = rng.uniform(5.1, 10.30, n).astype(np.float32)
    mass = (rng.uniform(0.5, 1.5, n) * 1e5).astype(np.float32)
    return dict(coords=coords, velocity=vel, mass=mass, metallicity=met, age=age)


def synthetic_ssp(nz: int = 6, na: int = 221, seed: int = 7):
    """A smooth, positive, BC03lr-shaped template: same axes as the real file (metallicity 1e-4..0.05,
    'age' 0, 5.1..10.3 used verbatim, 842 wavelengths 91..19950 A with 20 A spacing in the optical)."""
    rng = np.random.default_rng(seed)
    met = np.array([1e-4, 4e-4, 4e-3, 8e-3, 2e-2, 5e-2], dtype=np.float32)[:nz]
    age = np.concatenate([[0.0], np.linspace(5.1, 10.30103, na - 1)]).astype(np.float32)
    wl = np.concatenate([
        np.linspace(91, 3530, 200, endpoint=False),
        np.arange(3540, 3540 + 20 * 442, 20),
        np.linspace(12400, 19950, 200),
    ]).astype(np.float32)
    assert len(wl) == 842
    x = (wl / 5000.0).astype(np.float64)
    flux = np.empty((nz, na, len(wl)), dtype=np.float64)
    lines = rng.uniform(3600, 9900, 40)
    depth = rng.uniform(0.05, 0.5, 40)
    for i in range(nz):
        for j in range(na):
            temp = 0.6 + 1.2 * (1 - j / na) + 0.2 * i / nz
            cont = x ** -3 / (np.exp(1.0 / (x * temp)) - 1.0 + 1e-9)
            absorb = np.ones_like(x)
            for l0, dp in zip(lines, depth):
                absorb -= dp * (0.5 + 0.5 * j / na) * np.exp(-0.5 * ((wl - l0) / 15.0) ** 2)
            flux[i, j] = 1e-3 * cont * np.clip(absorb, 0.05, None) * (1.5 - j / na)
    return dict(metallicity=met, age=age, wavelength=wl, flux=flux.astype(np.float32))
