"""rubix/core/dust.py mirror: ``get_extinction(config)`` -> ``calculate_extinction(rubixdata)``, the extra
stage of the ``calc_dusty_ifu`` pipeline (pipeline_config.yml:62-126) between
``doppler_shift_and_resampling`` and ``calculate_datacube``.

The per-star A_V is computed on the device (``ops.dust_av`` -> rbx_dust_av).  On materialised spectra the
stage multiplies them in place like the reference (``ops.apply_extinction``); on a deferred recipe
(``config["b200"]["fused"]``) it only attaches (A_V, A(lambda)/A(V)) and ``calculate_datacube`` runs
Doppler shift, resampling, extinction and the per-spaxel sum in one kernel (rbx_build_cube_dusty).
"""

from __future__ import annotations

from typing import Callable

import numpy as np

from .. import dust as _dust
from ..cosmology import get_cosmology
from ..logger import get_logger
from ..telescope import calculate_spatial_bin_edges
from .data import RubixData
from .telescope import get_telescope


def get_extinction(config: dict) -> Callable:
    """rubix/core/dust.py:15-65 (same validation order and messages) + the configuration part of
    ``apply_spaxel_extinction`` (rubix/spectra/dust/dust_extinction.py:218-232, :286-296)."""
    logger = get_logger(config.get("logger", None))
    if "dust" not in config["ssp"]:
        raise ValueError("Dust configuration not found in config file.")
    if "extinction_model" not in config["ssp"]["dust"]:
        raise ValueError("Extinction model not found in dust configuration.")

    telescope = get_telescope(config)
    n_spaxels = int(telescope.sbin ** 2)
    wavelength = np.asarray(telescope.wave_seq, dtype=np.float32)
    _, spatial_bin_size = calculate_spatial_bin_edges(fov=telescope.fov, spatial_bins=telescope.sbin,
                                                      dist_z=config["galaxy"]["dist_z"], cosmology=get_cosmology(config))
    spaxel_area = float(np.float32(spatial_bin_size) ** 2)
    package = _dust.package_defaults()   # dust_to_gas_model / Xco come from the package-level rubix_config (:288-290)
    # per-configuration constants, evaluated once (the reference re-traces them inside its jit): the A(lambda)/A(V)
    # curve on the telescope grid, the dust-to-gas fit, the A_V constant.  An unknown model raises when the stage
    # runs, like the reference (dust_extinction.py:224-228).
    dcfg = config["ssp"]["dust"]
    ext_model = dcfg["extinction_model"]
    consts = {}
    if ext_model in _dust.RV_MODELS:
        consts["axav"] = _dust.extinction_curve(ext_model, wavelength, dcfg["Rv"])
        consts["dtg"] = _dust.dust_to_gas_parameters(package["dust_to_gas_model"], package["Xco"])
        consts["ext_const"] = _dust.extinction_constant(dcfg["dust_grain_density"])

    def calculate_extinction(rubixdata: RubixData) -> RubixData:
        """Apply the dust extinction to the spaxel data."""
        from .. import ops
        from .ifu import DeferredSpectra
        logger.info("Applying dust extinction to the spaxel data...")
        if ext_model not in _dust.RV_MODELS:   # dust_extinction.py:224-228
            raise ValueError(f"Extinction model '{ext_model}' is not available. Choose from {_dust.RV_MODELS}.")
        axav, dtg, ext_const = consts["axav"], consts["dtg"], consts["ext_const"]
        st, gas = rubixdata.stars, rubixdata.gas
        if gas.coords is None or gas.pixel_assignment is None or gas.metals is None or gas.mass is None:
            raise ValueError("calculate_extinction needs gas particles with coords, pixel_assignment, metals and mass")
        import torch

        def shard0(a, trailing, dtype=torch.float32):
            """First device shard, like the reference's ``[0]`` (dust_extinction.py:240-245, :283-291)."""
            t = ops.dev(a, dtype)
            return t[0] if t.ndim == trailing + 2 else t

        gas_coords, star_coords = shard0(gas.coords, 1), shard0(st.coords, 1)
        gpix, spix = shard0(gas.pixel_assignment, 0, torch.int32), shard0(st.pixel_assignment, 0, torch.int32)
        gmass, metals = shard0(gas.mass, 0), shard0(gas.metals, 1)
        av = ops.dust_av(gas_coords, gpix, gmass, metals, star_coords, spix, n_spaxels, dtg, ext_const, spaxel_area)
        if "axav_d" not in consts or consts["axav_d"].device != av.device:
            consts["axav_d"] = ops.dev(axav)
        axav_d = consts["axav_d"]
        if isinstance(st.spectra, DeferredSpectra):
            if not st.spectra.resampled:
                raise ValueError("calculate_extinction: spectra are not on the telescope wavelength grid "
                                 "(doppler_shift_and_resampling has not run)")
            st.spectra.extinction = (av, axav_d)
        else:
            spec = ops.dev(st.spectra)
            flat = spec.reshape(-1, spec.shape[-1])
            st.spectra = ops.apply_extinction(flat, av, axav_d).reshape(spec.shape)
        return rubixdata

    return calculate_extinction
