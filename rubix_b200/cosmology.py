"""Flat wCDM distances in float32, restating rubix/cosmology/base.py (comoving distance by a
256-point trapezoid, :99-116; angular scale :290-311) so that ``get_spatial_bin_edges`` can be
evaluated on the host without jax.  Runs once per configuration; not on the hot path."""

from __future__ import annotations

import numpy as np

C_SPEED = 2.99792458e8  # m/s, rubix/cosmology/base.py:12


class RubixCosmology:
    def __init__(self, Om0: float, w0: float, wa: float, h: float):
        f = np.float32
        self.Om0, self.w0, self.wa, self.h = f(Om0), f(w0), f(wa), f(h)

    def _rho_de_z(self, z):
        f = np.float32
        a = f(1.0) / (f(1.0) + z)
        return (a ** (f(-3.0) * (f(1.0) + self.w0 + self.wa)) * np.exp(f(-3.0) * self.wa * (f(1.0) - a))).astype(f)

    def _Ez(self, z):
        f = np.float32
        zp1 = f(1.0) + z
        return np.sqrt(self.Om0 * zp1**3 + (f(1.0) - self.Om0) * self._rho_de_z(z)).astype(f)

    def comoving_distance_to_z(self, redshift):
        f = np.float32
        z = np.linspace(0, redshift, 256).astype(f)
        y = (f(1.0) / self._Ez(z)).astype(f)
        acc = f(0.0)
        for k in range(1, len(z)):  # the reference's lax.scan trapezoid, sequential float32
            acc = f(acc + (z[k] - z[k - 1]) * (y[k] + y[k - 1]) / f(2.0))
        return f(f(f(acc * f(C_SPEED)) * f(1e-5)) / self.h)

    def angular_diameter_distance_to_z(self, redshift):
        return np.float32(self.comoving_distance_to_z(redshift) / np.float32(1 + redshift))

    def luminosity_distance_to_z(self, redshift):
        return np.float32(self.comoving_distance_to_z(redshift) * np.float32(1 + redshift))

    def angular_scale(self, z):
        """kpc per arcsec at redshift z (base.py:290-311)."""
        f = np.float32
        return f(f(self.angular_diameter_distance_to_z(z) * f(np.pi / (180 * 60 * 60))) * f(1e3))


PLANCK15 = RubixCosmology(0.3075, -1.0, 0.0, 0.6774)  # rubix/cosmology/__init__.py:3


def get_cosmology(config: dict) -> RubixCosmology:
    """rubix/core/cosmology.py:11-43."""
    name = config["cosmology"]["name"].upper()
    if name == "PLANCK15":
        return PLANCK15
    if name == "CUSTOM":
        return RubixCosmology(**config["cosmology"]["args"])
    raise ValueError(f"Cosmology {config['cosmology']['name']} not supported. Try PLANCK15 or CUSTOM.")
