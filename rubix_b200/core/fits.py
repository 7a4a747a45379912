"""rubix/core/fits.py mirror: ``store_fits(config, data, filepath)`` / ``load_fits(filepath)`` with the numpy-only
FITS writer of rubix_b200.fitslite (the reference uses astropy.io.fits and mpdaf, neither is in the image).
Same HDU layout, keywords and file name as rubix/core/fits.py:13-101."""

from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np

from ..fitslite import read_fits, write_fits
from ..logger import get_logger
from .telescope import get_telescope


def store_fits(config, data, filepath):
    """rubix/core/fits.py:13-101: empty primary HDU with the run's provenance, the cube transposed to
    (wavelength, x, y) in an IMAGE extension "DATA" with a linear AWAV axis."""
    logger = get_logger(config.get("logger", None))
    if "cube_type" not in config["data"]["args"] or config["data"]["args"]["cube_type"] == "stars":
        datacube, parttype = data.stars.datacube, "stars"
    elif config["data"]["args"]["cube_type"] == "gas":
        datacube, parttype = data.gas.datacube, "gas"
    else:
        raise ValueError(f"unknown cube_type {config['data']['args']['cube_type']!r}")
    if hasattr(datacube, "detach"):
        datacube = datacube.detach().cpu().numpy()
    datacube = np.asarray(datacube)
    telescope = get_telescope(config)

    galaxy_id = config["data"]["load_galaxy_args"]["id"]
    snapshot = config["data"]["args"]["snapshot"]
    hdr = {
        "PIPELINE": config["pipeline"]["name"],
        "DIST_z": config["galaxy"]["dist_z"],
        "ROTATION": config["galaxy"]["rotation"]["type"],
        "SIM": config["simulation"]["name"],
        "GALAXYID": galaxy_id,
        "SNAPSHOT": snapshot,
        "SUBSET": config["data"]["subset"]["use_subset"],
        "SSP": config["ssp"]["template"]["name"],
        "INSTR": config["telescope"]["name"],
        "PSF": config["telescope"]["psf"]["name"],
        "PSF_SIZE": config["telescope"]["psf"]["size"],
        "PSFSIGMA": config["telescope"]["psf"]["sigma"],
        "LSF": config["telescope"]["lsf"]["sigma"],
        "S_TO_N": config["telescope"]["noise"]["signal_to_noise"],
        "N_DISTR": config["telescope"]["noise"]["noise_distribution"],
        "COSMO": config["cosmology"]["name"],
    }
    object_name = f"{config['simulation']['name']} {galaxy_id}"
    hdr1 = {
        "EXTNAME": "DATA", "OBJECT": object_name, "BUNIT": "erg/(s*cm^2*A)",
        "CRPIX1": (datacube.shape[0] - 1) / 2, "CRPIX2": (datacube.shape[1] - 1) / 2,
        "CD1_1": telescope.spatial_res / 3600, "CD1_2": 0, "CD2_1": 0, "CD2_2": telescope.spatial_res / 3600,
        "CUNIT1": "deg", "CUNIT2": "deg", "CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CTYPE3": "AWAV",
        "CUNIT3": "Angstrom", "CD3_3": float(telescope.wave_res), "CRPIX3": 1, "CRVAL3": float(telescope.wave_range[0]),
        "CD1_3": 0, "CD2_3": 0, "CD3_1": 0, "CD3_2": 0,
    }
    output_filename = (f"{filepath}{config['simulation']['name']}_id{galaxy_id}_snap{snapshot}_"
                       f"{parttype}_subset{config['data']['subset']['use_subset']}.fits")
    os.makedirs(os.path.dirname(output_filename) or ".", exist_ok=True)
    write_fits(output_filename, hdr, [(np.ascontiguousarray(datacube.T), hdr1)])
    logger.info(f"Datacube saved to {output_filename}")
    return output_filename


def load_fits(filepath):
    """rubix/core/fits.py:104-115 returns an ``mpdaf.obj.Cube``; here a namespace with the same essentials:
    ``data`` (wavelength, x, y), ``primary_header``, ``data_header`` and ``wave`` (the linear AWAV axis)."""
    hdus = read_fits(filepath)
    primary, _ = hdus[0]
    hdr, data = next((h, d) for h, d in hdus[1:] if h.get("EXTNAME") == "DATA")
    wave = float(hdr["CRVAL3"]) + float(hdr["CD3_3"]) * (np.arange(data.shape[0]) + 1 - float(hdr["CRPIX3"]))
    return SimpleNamespace(data=data, primary_header=primary, data_header=hdr, wave=wave, shape=data.shape)
