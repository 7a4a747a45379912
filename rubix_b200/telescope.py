"""Telescope description and grids (host side), mirroring rubix/telescope/{base,factory,utils}.py."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import numpy as np

from .config import TELESCOPES
from .utils import read_yaml


@dataclass
class BaseTelescope:
    """Fields of rubix/telescope/base.py:10-39 (aperture_region is a numpy mask here)."""

    fov: float
    spatial_res: float
    wave_range: List[float]
    wave_res: float
    lsf_fwhm: float
    signal_to_noise: Optional[float]
    sbin: int
    aperture_region: np.ndarray
    pixel_type: str
    wave_seq: np.ndarray
    wave_edges: np.ndarray
    name: str = ""


def calculate_wave_seq(wave_range, wave_res) -> np.ndarray:
    """rubix/telescope/utils.py:53: ``jnp.arange(lo, hi, res)`` in float32.  Bit-exact with the
    ``wave`` array the reference wrote to notebooks/data/dummy_datacube.h5 for MUSE."""
    return np.arange(wave_range[0], wave_range[1], wave_res, dtype=np.float32)


def calculate_wave_edges(wave_seq, wave_res) -> np.ndarray:
    """rubix/telescope/utils.py:70-73."""
    start = wave_seq[0] - np.float32(wave_res / 2)
    end = wave_seq[-1] + np.float32(wave_res / 2)
    return np.arange(start, end, wave_res, dtype=np.float32)


def calculate_spatial_bin_edges(fov, spatial_bins, dist_z, cosmology):
    """rubix/telescope/utils.py:30-37, float32 like the jnp evaluation: the number of edges can be
    ``spatial_bins + 1`` or ``+ 2`` depending on rounding -- callers must not assume either."""
    f = np.float32
    ang = f(cosmology.angular_scale(dist_z))
    aperture = f(ang * f(fov))
    size = f(aperture / f(spatial_bins))
    edges = np.arange(f(-aperture / f(2)), f(aperture / f(2) + size), size, dtype=np.float32)
    return edges, size


def square_aperture(n: int) -> np.ndarray:
    return np.ones(n * n, dtype=np.float32)


def circular_aperture(n: int) -> np.ndarray:
    c = (n - 1) / 2.0
    y, x = np.mgrid[0:n, 0:n]
    return (((x - c) ** 2 + (y - c) ** 2) <= (n / 2.0) ** 2).astype(np.float32).ravel()


def hexagonal_aperture(n: int) -> np.ndarray:
    """rubix/telescope/apertures.py:12-40 (HEXAGONAL_APERTURE): a flat-topped hexagon inscribed in the n x n grid,
    pixel centres at 1 .. n relative to the centre n / 2 + 1 / 2; same strict / inclusive comparisons."""
    c = n / 2 + 0.5
    idx = np.arange(1, n + 1, dtype=np.float64)
    xx = idx[:, None] - c    # first array axis, as in the reference's ap_region[x - 1, y - 1]
    yy = idx[None, :] - c
    h = n * np.sqrt(3.0) / 4
    rr = 2 * (n / 4) * h - (n / 4) * np.abs(yy) - h * np.abs(xx)
    return ((rr >= 0) & (np.abs(xx) < n / 2) & (np.abs(yy) < h)).astype(np.float32).ravel()


class TelescopeFactory:
    """rubix/telescope/factory.py:18-110: ``create_telescope(name)`` from the built-in table, a YAML
    path or a ``{name: {...}}`` dict (custom telescopes, e.g. the large-FOV MUSE variant)."""

    def __init__(self, telescopes_config: Optional[Union[dict, str]] = None):
        if telescopes_config is None:
            self.telescopes_config: Dict[str, dict] = TELESCOPES
        elif isinstance(telescopes_config, str):
            self.telescopes_config = read_yaml(telescopes_config)
        else:
            self.telescopes_config = telescopes_config

    def create_telescope(self, name: str) -> BaseTelescope:
        if name not in self.telescopes_config:
            raise ValueError(f"Telescope {name} not found in config")
        c = self.telescopes_config[name]
        sbin = int(np.floor(c["fov"] / c["spatial_res"]))
        ap = c.get("aperture_type", "square")
        if ap == "square":
            region = square_aperture(sbin)
        elif ap == "circular":
            region = circular_aperture(sbin)  # aperture masks are outside the particle->cube path
        elif ap == "hexagonal":
            region = hexagonal_aperture(sbin)
        else:
            raise ValueError(f"Unknown aperture type: {ap}")
        wave_seq = calculate_wave_seq(c["wave_range"], c["wave_res"])
        return BaseTelescope(fov=c["fov"], spatial_res=c["spatial_res"], wave_range=list(c["wave_range"]),
                             wave_res=c["wave_res"], lsf_fwhm=c["lsf_fwhm"],
                             signal_to_noise=c.get("signal_to_noise"), sbin=sbin, aperture_region=region,
                             pixel_type=c.get("pixel_type", "square"), wave_seq=wave_seq,
                             wave_edges=calculate_wave_edges(wave_seq, c["wave_res"]), name=name)


def gaussian_kernel_2d(m: int, n: int, sigma: float) -> np.ndarray:
    """rubix/telescope/psf/kernels.py:26-31 in float32 (config-time constant, like the reference
    which builds it in the factory, outside the jitted function)."""
    x = np.arange(-((m - 1) / 2), ((m - 1) / 2) + 1).astype(np.float32)
    y = np.arange(-((n - 1) / 2), ((n - 1) / 2) + 1).astype(np.float32)
    X, Y = np.meshgrid(x, y, indexing="ij")
    v = np.exp(-(X**2 + Y**2) / np.float32(2 * sigma**2)).astype(np.float32)
    return (v / v.sum(dtype=np.float32)).astype(np.float32)


def get_psf_kernel(name: str, m: int, n: int, **kwargs) -> np.ndarray:
    """rubix/telescope/psf/psf.py:14-31."""
    if name == "gaussian":
        return gaussian_kernel_2d(m=m, n=n, **kwargs)
    raise ValueError(f"Unknown PSF kernel name: {name}")


def lsf_kernel(sigma: float, wave_res: float, factor: int = 12) -> np.ndarray:
    """rubix/telescope/lsf/lsf.py:12-26 in float32."""
    x = np.arange(-factor * wave_res, factor * wave_res + wave_res, wave_res).astype(np.float32)
    r = np.exp(np.float32(-0.5) * (x**2) / np.float32(sigma**2)).astype(np.float32)
    return (r / r.sum(dtype=np.float32)).astype(np.float32)
