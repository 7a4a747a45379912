"""store_fits / load_fits mirror (rubix/core/fits.py:13-115) on the numpy-only FITS writer."""

import copy
import os
from types import SimpleNamespace

import numpy as np

from rubix_b200 import fitslite
from rubix_b200.core.fits import load_fits, store_fits

CONFIG = {
    "pipeline": {"name": "calc_ifu"},
    "logger": {"log_level": "WARNING", "log_file_path": None, "format": "%(message)s"},
    "simulation": {"name": "IllustrisTNG"},
    "data": {"args": {"snapshot": 99, "particle_type": ["stars"]}, "load_galaxy_args": {"id": 11},
             "subset": {"use_subset": True, "subset_size": 1000}},
    "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}, "lsf": {"sigma": 0.5},
                  "noise": {"signal_to_noise": 100, "noise_distribution": "normal"}},
    "cosmology": {"name": "PLANCK15"},
    "galaxy": {"dist_z": 0.1, "rotation": {"type": "edge-on"}},
    "ssp": {"template": {"name": "BruzualCharlot2003"}},
}


def test_store_and_load_fits_round_trip(tmp_path):
    rng = np.random.default_rng(42)
    cube = rng.uniform(0, 1, (25, 25, 3721)).astype(np.float32)
    data = SimpleNamespace(stars=SimpleNamespace(datacube=cube), gas=SimpleNamespace(datacube=None))
    name = store_fits(copy.deepcopy(CONFIG), data, str(tmp_path) + "/")
    # rubix/core/fits.py:92-95
    assert os.path.basename(name) == "IllustrisTNG_id11_snap99_stars_subsetTrue.fits"
    assert os.path.getsize(name) % 2880 == 0
    raw = open(name, "rb").read()
    assert raw[:30] == b"SIMPLE  =                    T"
    back = load_fits(name)
    assert back.data.shape == (3721, 25, 25) and np.array_equal(back.data, cube.T)   # ImageHDU(datacube.T)
    h = back.data_header
    assert h["EXTNAME"] == "DATA" and h["CTYPE3"] == "AWAV" and h["BUNIT"] == "erg/(s*cm^2*A)"
    assert h["NAXIS1"] == 25 and h["NAXIS3"] == 3721 and h["BITPIX"] == -32
    assert abs(h["CD1_1"] - 0.2 / 3600) < 1e-15 and h["CRPIX1"] == 12.0 and h["CD3_3"] == 1.25
    assert abs(back.wave[0] - 4700.15) < 1e-9 and abs(back.wave[-1] - (4700.15 + 1.25 * 3720)) < 1e-6
    p = back.primary_header
    assert p["PIPELINE"] == "calc_ifu" and p["GALAXYID"] == 11 and p["SNAPSHOT"] == 99 and p["SUBSET"] is True
    assert p["INSTR"] == "MUSE" and p["PSFSIGMA"] == 0.6 and p["N_DISTR"] == "normal" and p["DIST_Z"] == 0.1
    assert p["ROTATION"] == "edge-on" and p["NAXIS"] == 0


def test_fitslite_dtypes_and_strings(tmp_path):
    path = str(tmp_path / "a.fits")
    a = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    b = np.arange(6, dtype=np.int32).reshape(3, 2)
    fitslite.write_fits(path, {"NOTE": "it's", "FLAG": False, "X": 1.5e-7}, [(a, {"EXTNAME": "A"}), (b, {"EXTNAME": "B"})])
    hdus = fitslite.read_fits(path)
    assert len(hdus) == 3 and hdus[0][1] is None
    assert hdus[0][0]["NOTE"] == "it's" and hdus[0][0]["FLAG"] is False and hdus[0][0]["X"] == 1.5e-7
    assert np.array_equal(hdus[1][1], a) and hdus[1][0]["NAXIS1"] == 4 and hdus[1][0]["BITPIX"] == -64
    assert np.array_equal(hdus[2][1], b) and hdus[2][0]["BITPIX"] == 32
