"""Package-level configuration, mirroring the keys of the reference's ``rubix.config``
(rubix/config/rubix_config.yml, rubix/config/pipeline_config.yml, rubix/telescope/telescopes.yaml)
that the particle -> datacube path reads.  Values are the reference's; the layout is ours."""

from __future__ import annotations

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)

#: rubix_config.yml:1-12
CONSTANTS = {
    "SPEED_OF_LIGHT": 299792.458,  # km/s
    "LSOL_TO_ERG": 3.828e33,
    "MPC_TO_CM": 3.08568e24,
    "KPC_TO_CM": 3.08568e21,       # rubix_config.yml:6
    "MSUN_TO_GRAMS": 1.989e33,     # rubix_config.yml:7
    "MASS_OF_PROTON": 1.67262e-24, # rubix_config.yml:18
}

#: rubix_config.yml:145-150 (``ssp.dust``): package-level defaults of the dust variant
DUST = {"extinction_model": "Cardelli89", "Rv": 3.1, "dust_to_gas_model": "broken power law fit", "Xco": "Z",
        "dust_grain_density": 3.5}

#: rubix_config.yml:242-245
IFU = {"doppler": {"velocity_direction": "z"}}

#: rubix/telescope/telescopes.yaml (fov in arcsec, wavelengths in Angstrom)
def _t(fov, sres, wr, wres, fwhm, ap="square"):
    return dict(fov=fov, spatial_res=sres, wave_range=list(wr), wave_res=wres, lsf_fwhm=fwhm,
                signal_to_noise=None, aperture_type=ap, pixel_type="square")


TELESCOPES = {
    "MUSE": _t(5.0, 0.2, (4700.15, 9351.4), 1.25, 2.51),
    "NIRSpec_PRISM_CLEAR": _t(3.0, 0.1, (6000, 53000), 114.36, 471.0),
    "NIRSpec_G140M_F070LP": _t(3.0, 0.1, (7000, 12600), 6.38, 15.02),
    "NIRSpec_G140M_F100LP": _t(3.0, 0.1, (9800, 18800), 6.38, 15.02),
    "NIRSpec_G235M_F170LP": _t(3.0, 0.1, (17000, 31500), 10.69, 25.2),
    "NIRSpec_G395M_F290LP": _t(3.0, 0.1, (28800, 52000), 17.98, 42.39),
    "NIRSpec_G140H_F070LP": _t(3.0, 0.1, (7000, 12600), 2.41, 5.65),
    "NIRSpec_G140H_F100LP": _t(3.0, 0.1, (9800, 18700), 2.37, 5.65),
    "NIRSpec_G235H_F170LP": _t(3.0, 0.1, (17000, 31500), 3.98, 9.42),
    "NIRSpec_G395H_F290LP": _t(3.0, 0.1, (28800, 52000), 6.69, 15.78),
    "MaNGA_12": _t(12, 0.5, (3622, 10354), 1.04, 2.85, "hexagonal"),
    "MaNGA_32": _t(32, 0.5, (3622, 10354), 1.04, 2.85, "hexagonal"),
    "SAMI": _t(15, 0.5, (3750, 5750), 1.04, 2.65, "circular"),
    "HECTOR": _t(30, 0.1, (3720, 5910), 1.6, 1.3, "hexagonal"),
    "CALIFA": _t(74, 1.0, (3700, 4750), 2.7, 2.7, "hexagonal"),
}

#: rubix_config.yml:137-176: HDF5 templates and how their fields are read (no log transform, the
#: file's values are used verbatim in the internal units).
_FIELDS = {k: {"name": k, "in_log": False} for k in ("age", "metallicity", "wavelength", "flux")}
SSP = {
    "units": {"age": "Gyr", "metallicity": "", "wavelength": "Angstrom", "flux": "Lsun/Angstrom"},
    "templates": {
        "BruzualCharlot2003": {"name": "Bruzual & Charlot (2003)", "format": "HDF5",
                               "file_name": "BC03lr.h5", "fields": _FIELDS},
        "FSPS": {"name": "FSPS (Conroy et al. 2009)", "format": "fsps", "file_name": "fsps.h5",
                 "fields": {k: {"name": k, "in_log": k in ("age", "metallicity")} for k in _FIELDS}},
    },
}

#: where SSP template files are looked for, in order: $RUBIX_B200_TEMPLATE_PATH, then the package's templates/
#: directory (BC03lr as the float32 arrays the reference's loader produces, tools/make_golden.py).
TEMPLATE_PATHS = [p for p in (os.environ.get("RUBIX_B200_TEMPLATE_PATH"), os.path.join(_HERE, "templates")) if p]

#: pipeline_config.yml:1-60 (calc_ifu): node name -> depends_on
def _chain(names):
    out, prev = {}, None
    for n in names:
        out[n] = {"name": n, "depends_on": prev, "args": [], "kwargs": {}}
        prev = n
    return {"Transformers": out}


PIPELINES = {
    "calc_ifu": _chain(["rotate_galaxy", "filter_particles", "spaxel_assignment", "reshape_data",
                        "calculate_spectra", "scale_spectrum_by_mass", "doppler_shift_and_resampling",
                        "calculate_datacube", "convolve_psf", "convolve_lsf", "apply_noise"]),
    # pipeline_config.yml:62-126
    "calc_dusty_ifu": _chain(["rotate_galaxy", "filter_particles", "spaxel_assignment", "reshape_data",
                              "calculate_spectra", "scale_spectrum_by_mass", "doppler_shift_and_resampling",
                              "calculate_extinction", "calculate_datacube", "convolve_psf", "convolve_lsf",
                              "apply_noise"]),
}

SSP["dust"] = DUST
rubix_config = {"constants": CONSTANTS, "ifu": IFU, "ssp": SSP, "telescopes": TELESCOPES, "pipelines": PIPELINES}
