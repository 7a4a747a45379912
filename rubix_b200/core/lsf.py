"""rubix/core/lsf.py mirror."""

from __future__ import annotations

from typing import Callable

from ..logger import get_logger
from ..telescope import lsf_kernel
from .data import RubixData
from .psf import DeferredPSF
from .telescope import get_telescope


def get_convolve_lsf(config: dict) -> Callable:
    """rubix/core/lsf.py:14-67 (same validation and messages)."""
    logger = get_logger(config.get("logger", None))
    if "lsf" not in config["telescope"]:
        raise ValueError("LSF configuration not found in telescope configuration")
    if "sigma" not in config["telescope"]["lsf"]:
        raise ValueError("LSF sigma size not found in telescope configuration")
    sigma = config["telescope"]["lsf"]["sigma"]
    telescope = get_telescope(config)
    kernel = lsf_kernel(sigma, telescope.wave_res, factor=12)

    def convolve_lsf(rubixdata: RubixData) -> RubixData:
        """Convolve the input datacube with the LSF."""
        from .. import ops
        logger.info("Convolving with LSF...")
        cube = rubixdata.stars.datacube
        if isinstance(cube, DeferredPSF):
            rubixdata.stars.datacube = ops.psf_lsf(cube.cube, cube.kernel, kernel, ext=12)
        else:
            rubixdata.stars.datacube = ops.convolve_lsf(ops.dev(cube), kernel, ext=12)
        return rubixdata

    return convolve_lsf
