"""rubix/core/telescope.py mirrors: get_telescope, get_spatial_bin_edges, get_spaxel_assignment,
get_filter_particles."""

from __future__ import annotations

from typing import Callable, Union

from ..cosmology import get_cosmology
from ..logger import get_logger
from ..telescope import BaseTelescope, TelescopeFactory, calculate_spatial_bin_edges
from .data import RubixData


def get_telescope(config: Union[str, dict]) -> BaseTelescope:
    """rubix/core/telescope.py:21-47.  Extension: ``config["telescope"]["custom"]`` may hold a
    ``{name: {...}}`` table for telescopes that are not built in (the reference has a TODO for it)."""
    custom = config["telescope"].get("custom") if isinstance(config["telescope"], dict) else None
    factory = TelescopeFactory(custom) if custom else TelescopeFactory()
    return factory.create_telescope(config["telescope"]["name"])


def get_spatial_bin_edges(config: dict):
    """rubix/core/telescope.py:51-78 (host, float32)."""
    telescope = get_telescope(config)
    cosmology = get_cosmology(config)
    edges, _ = calculate_spatial_bin_edges(fov=telescope.fov, spatial_bins=telescope.sbin,
                                           dist_z=config["galaxy"]["dist_z"], cosmology=cosmology)
    return edges


def _flat(t, trailing):
    """Collapse an optional leading device axis: (n_dev, P, *trailing) or (P, *trailing) -> (n, *trailing)."""
    return t.reshape((-1,) + tuple(trailing))


def get_spaxel_assignment(config: dict) -> Callable:
    """rubix/core/telescope.py:82-124."""
    logger = get_logger(config.get("logger", None))
    telescope = get_telescope(config)
    if telescope.pixel_type not in ["square"]:
        raise ValueError(f"Pixel type {telescope.pixel_type} not supported")
    spatial_bin_edges = get_spatial_bin_edges(config)

    def spaxel_assignment(rubixdata: RubixData) -> RubixData:
        from .. import ops
        logger.info("Assigning particles to spaxels...")
        edges = ops.dev(spatial_bin_edges)
        for part in (rubixdata.stars, rubixdata.gas):
            if part.coords is not None:
                coords = ops.dev(part.coords)
                pix = ops.spaxel_assign(_flat(coords, (3,)), edges)
                part.pixel_assignment = pix.reshape(coords.shape[:-1])
                part.spatial_bin_edges = edges
        return rubixdata

    return spaxel_assignment


def get_filter_particles(config: dict) -> Callable:
    """rubix/core/telescope.py:128-199: every per-particle attribute except coords / velocity is set
    to 0 outside the aperture; ``mask`` is stored."""
    logger = get_logger(config.get("logger", None))
    spatial_bin_edges = get_spatial_bin_edges(config)

    def filter_particles(rubixdata: RubixData) -> RubixData:
        import numpy as np
        import torch
        from .. import ops
        logger.info("Filtering particles outside the aperture...")
        edges = ops.dev(spatial_bin_edges)
        for name in config["data"]["args"]["particle_type"] if "data" in config else ["stars"]:
            part = getattr(rubixdata, name)
            if part.coords is None:
                continue
            coords = ops.dev(part.coords)
            part.coords = coords
            n = coords.reshape(-1, 3).shape[0]
            hot = {}
            for k in ("mass", "metallicity", "age"):
                v = getattr(part, k, None)
                if v is not None:
                    hot[k] = ops.dev(v).clone()
            mask = ops.filter_particles(coords.reshape(-1, 3), edges,
                                        *(hot[k].reshape(-1) if k in hot else None
                                          for k in ("mass", "metallicity", "age")))
            for k, v in hot.items():
                setattr(part, k, v)
            # the remaining per-particle attributes (gas fields, pixel ids ...) follow the same rule
            for k, v in vars(part).items():
                if k in ("coords", "velocity", "mask", "spectra", "datacube", "spatial_bin_edges") or k in hot:
                    continue
                if name == "gas" and k == "metals":   # rubix/core/telescope.py:186: gas metals are not masked
                    continue
                if isinstance(v, np.ndarray) and v.dtype.kind in "fiu" and v.size and v.shape[:1] == tuple(coords.shape[:1]):
                    # prepare_input keeps the gas fields as numpy arrays: they follow the mask rule like every
                    # other per-particle attribute (rubix/core/telescope.py:176-190)
                    v = ops.dev(v, dtype=torch.float32 if v.dtype.kind == "f" else torch.int32)
                if isinstance(v, torch.Tensor) and v.numel() and v.reshape(-1).shape[0] % n == 0 and v.shape[:1] == coords.shape[:1]:
                    m = mask.reshape(coords.shape[:-1])
                    while m.ndim < v.ndim:
                        m = m.unsqueeze(-1)
                    setattr(part, k, torch.where(m, v, torch.zeros((), dtype=v.dtype, device=v.device)))
            part.mask = mask.reshape(coords.shape[:-1])
        return rubixdata

    return filter_particles
