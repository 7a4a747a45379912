"""rubix/core/psf.py mirror."""

from __future__ import annotations

from typing import Callable

from ..logger import get_logger
from ..telescope import get_psf_kernel
from .data import RubixData


class DeferredPSF:
    """In fused mode ``convolve_psf`` leaves this in ``stars.datacube``: the raw cube plus the PSF
    taps, so that ``convolve_lsf`` can run PSF and LSF in one pass.  ``materialize()`` gives the
    PSF-convolved cube the reference would hold at this point."""

    def __init__(self, cube, kernel):
        self.cube, self.kernel = cube, kernel
        self.shape, self.dtype = tuple(cube.shape), cube.dtype

    def materialize(self):
        from .. import ops
        return ops.convolve_psf(self.cube, self.kernel)

    def __array__(self, dtype=None, copy=None):
        a = self.materialize().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a


def get_convolve_psf(config: dict) -> Callable:
    """rubix/core/psf.py:15-75 (same validation order and messages)."""
    logger = get_logger(config.get("logger", None))
    if "psf" not in config["telescope"]:
        raise ValueError("PSF configuration not found in telescope configuration")
    if "name" not in config["telescope"]["psf"]:
        raise ValueError("PSF name not found in telescope configuration")
    if config["telescope"]["psf"]["name"] == "gaussian":
        if "size" not in config["telescope"]["psf"]:
            raise ValueError("PSF size not found in telescope configuration")
        if "sigma" not in config["telescope"]["psf"]:
            raise ValueError("PSF sigma not found in telescope configuration")
        m = n = config["telescope"]["psf"]["size"]
        psf_kernel = get_psf_kernel("gaussian", m, n, sigma=config["telescope"]["psf"]["sigma"])
    else:
        raise ValueError(f"Unknown PSF kernel name: {config['telescope']['psf']['name']}")
    fuse = isinstance(config.get("b200"), dict) and config["b200"].get("fused") is True

    def convolve_psf(rubixdata: RubixData) -> RubixData:
        """Convolve the input datacube with the PSF kernel."""
        from .. import ops
        logger.info("Convolving with PSF...")
        cube = ops.dev(rubixdata.stars.datacube)
        rubixdata.stars.datacube = DeferredPSF(cube, psf_kernel) if fuse else ops.convolve_psf(cube, psf_kernel)
        return rubixdata

    return convolve_psf
