"""SSP template grids on the host, mirroring rubix/spectra/ssp/{grid,factory}.py for the HDF5
templates (BC03lr) without h5py."""

from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .config import SSP, TEMPLATE_PATHS
from .h5lite import H5File


@dataclass
class SSPGrid:
    """rubix/spectra/ssp/grid.py:22-40: age (A,), metallicity (Z,), wavelength (L,), flux (Z, A, L),
    all float32."""

    age: np.ndarray
    metallicity: np.ndarray
    wavelength: np.ndarray
    flux: np.ndarray
    name: str = ""

    def __post_init__(self):
        for k in ("age", "metallicity", "wavelength", "flux"):
            setattr(self, k, np.ascontiguousarray(np.asarray(getattr(self, k)), dtype=np.float32))
        if self.flux.shape != (len(self.metallicity), len(self.age), len(self.wavelength)):
            raise ValueError(f"flux shape {self.flux.shape} does not match (metallicity, age, wavelength)")

    def keys(self):
        return ["age", "metallicity", "wavelength", "flux"]

    def get_lookup_interpolation(self, method: str = "cubic", extrap: int = 0):
        """grid.py:61-124: returns ``lookup(metallicity, age) -> (n, L)`` evaluated on the GPU
        (interpax.interp2d semantics; only ``extrap=0`` is what rubix binds)."""
        if extrap != 0:
            raise NotImplementedError("rubix binds extrap=0; other values are not implemented")
        from . import ops
        plan = ops.Plan(self.metallicity, self.age, self.wavelength, self.flux,
                        np.array([1.0], dtype=np.float32), 0.0, method=method)

        def lookup(metallicity, age):
            return ops.ssp_lookup(plan, metallicity, age)

        lookup.__doc__ = "Interpolation function for SSP grid, args: f(metallicity, age)"
        return lookup


class HDF5SSPGrid(SSPGrid):
    @classmethod
    def from_file(cls, config: dict, file_location: str) -> "SSPGrid":
        """grid.py:304-335: read the four datasets, 10**x where ``in_log``, cast to float32."""
        if config.get("format", "").lower() not in ["hdf5", "fsps"]:
            raise ValueError("Configured file format is not HDF5.")
        path = os.path.join(file_location, config["file_name"])
        if not os.path.exists(path):
            raise FileNotFoundError(f"SSP template {path} not found")
        data = {}
        with H5File(path) as f:
            for field, info in config["fields"].items():
                arr = f[info["name"]].read()
                if info.get("in_log"):
                    arr = np.power(10.0, arr)
                data[field] = np.asarray(arr, dtype=np.float32)
        return cls(name=config["name"], **data)


def get_ssp_template(template: str) -> SSPGrid:
    """rubix/spectra/ssp/factory.py:13-76 for the HDF5 family.  Looks in TEMPLATE_PATHS; the float32
    ``.npz`` fixture of BC03lr under tests/golden/ is accepted in place of the ``.h5`` file."""
    cfg = SSP["templates"]
    if template not in cfg:
        raise ValueError(f"SSP template {template} not found in the supported configuration file.")
    c = cfg[template]
    if c["format"].lower() not in ("hdf5", "fsps"):
        raise ValueError("Currently only HDF5 format and fits files in the format of pyPipe3D format are "
                         "supported for SSP templates.")
    for d in TEMPLATE_PATHS:
        if os.path.exists(os.path.join(d, c["file_name"])):
            return HDF5SSPGrid.from_file(c, d)
        npz = os.path.join(d, os.path.splitext(c["file_name"])[0].lower() + "_f32.npz")
        if os.path.exists(npz):
            z = np.load(npz)
            return SSPGrid(age=z["age"], metallicity=z["metallicity"], wavelength=z["wavelength"],
                           flux=z["flux"], name=c["name"])
    raise FileNotFoundError(f"SSP template file {c['file_name']} not found in {TEMPLATE_PATHS}")
