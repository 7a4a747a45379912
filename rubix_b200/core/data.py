"""Containers crossing the boundary (rubix/core/data.py:65-355): ``RubixData{galaxy, stars, gas}``.
Fields are CUDA tensors (float32 / int32) or ``None``; numpy arrays are accepted and moved to the
device by the stage functions."""

from __future__ import annotations

from dataclasses import dataclass, field, fields
from typing import Any, Optional

import numpy as np


def _repr(obj, title):
    out = [title]
    for f in fields(obj):
        v = getattr(obj, f.name)
        out.append(f"{f.name}: None" if v is None else
                   f"{f.name}: shape = {tuple(getattr(v, 'shape', ()))}, dtype = {getattr(v, 'dtype', type(v))}")
    return "\n\t".join(out)


@dataclass
class Galaxy:
    redshift: Optional[Any] = None
    center: Optional[Any] = None
    halfmassrad_stars: Optional[Any] = None

    def __repr__(self):
        return _repr(self, "Galaxy:")


@dataclass
class StarsData:
    """Field order of rubix/core/data.py:140-149."""
    coords: Optional[Any] = None
    velocity: Optional[Any] = None
    mass: Optional[Any] = None
    metallicity: Optional[Any] = None
    age: Optional[Any] = None
    pixel_assignment: Optional[Any] = None
    spatial_bin_edges: Optional[Any] = None
    mask: Optional[Any] = None
    spectra: Optional[Any] = None
    datacube: Optional[Any] = None

    def __repr__(self):
        return _repr(self, "StarsData:")


@dataclass
class GasData:
    coords: Optional[Any] = None
    velocity: Optional[Any] = None
    mass: Optional[Any] = None
    density: Optional[Any] = None
    internal_energy: Optional[Any] = None
    metallicity: Optional[Any] = None
    metals: Optional[Any] = None
    sfr: Optional[Any] = None
    electron_abundance: Optional[Any] = None
    pixel_assignment: Optional[Any] = None
    spatial_bin_edges: Optional[Any] = None
    mask: Optional[Any] = None
    spectra: Optional[Any] = None
    datacube: Optional[Any] = None

    def __repr__(self):
        return _repr(self, "GasData:")


@dataclass
class RubixData:
    galaxy: Galaxy = field(default_factory=Galaxy)
    stars: StarsData = field(default_factory=StarsData)
    gas: GasData = field(default_factory=GasData)

    def __repr__(self):
        return f"RubixData:\n\t{self.galaxy}\n\t{self.stars}\n\t{self.gas}"


def make_rubix_data(coords, velocity, mass, metallicity, age, device: bool = True) -> RubixData:
    """Build a RubixData with star particles from host or device arrays (float32)."""
    rd = RubixData()
    vals = dict(coords=coords, velocity=velocity, mass=mass, metallicity=metallicity, age=age)
    for k, v in vals.items():
        if device:
            from .. import ops
            v = ops.dev(v)
        else:
            v = np.ascontiguousarray(np.asarray(v), dtype=np.float32)
        setattr(rd.stars, k, v)
    return rd


def device_count() -> int:
    """The reference pads/reshapes to ``jax.device_count()`` devices inside one process
    (rubix/core/data.py:461).  Here there is one process per GPU, so the in-process device axis has
    length 1; cross-GPU sharding is done by rank (rubix_b200.parallel)."""
    return 1


def reshape_array(arr, n_dev: Optional[int] = None):
    """rubix/core/data.py:447-487: (n, ...) -> (n_dev, ceil(n / n_dev), ...) zero-padded."""
    import torch
    n_dev = n_dev or device_count()
    if not isinstance(arr, torch.Tensor):
        arr = torch.from_numpy(np.ascontiguousarray(np.asarray(arr)))
    n = arr.shape[0]
    per = (n + n_dev - 1) // n_dev
    pad = per * n_dev - n
    if pad:
        arr = torch.cat([arr, torch.zeros((pad,) + tuple(arr.shape[1:]), dtype=arr.dtype, device=arr.device)], 0)
    return arr.reshape((n_dev, per) + tuple(arr.shape[1:]))


_RESHAPED = ("coords", "velocity", "mass", "metallicity", "age", "pixel_assignment", "mask",
             "density", "internal_energy", "metals", "sfr", "electron_abundance")


def get_reshape_data(config: dict):
    """rubix/core/data.py:639-671: add the leading device axis to every per-particle array."""
    import torch

    def reshape_data(rubixdata: RubixData) -> RubixData:
        for part in (rubixdata.stars, rubixdata.gas):
            if part.coords is None:
                continue
            for k in _RESHAPED:
                v = getattr(part, k, None)
                if isinstance(v, (torch.Tensor, np.ndarray)):
                    setattr(part, k, reshape_array(v))
        return rubixdata

    return reshape_data
