"""CPU tests of the host-side mirror of the rubix.core interface: config validation and error
messages (the reference's tests/test_core_psf.py, test_core_lsf.py, test_core_ssp.py,
test_core_pipeline.py:80-85), grids, padding/sharding, the linear pipeline ordering and the numpy-only
HDF5 reader.  No CUDA calls."""

import copy
import os

import numpy as np
import pytest

from rubix_b200 import parallel
from rubix_b200.config import PIPELINES
from rubix_b200.core import data as cdata
from rubix_b200.core import pipeline as cpipe
from rubix_b200.core.lsf import get_convolve_lsf
from rubix_b200.core.psf import get_convolve_psf
from rubix_b200.core.ssp import get_method, get_ssp
from rubix_b200.core.telescope import get_spatial_bin_edges, get_telescope
from rubix_b200.utils import get_pipeline_config

CONFIG = {
    "pipeline": {"name": "calc_ifu"},
    "logger": {"log_level": "WARNING", "log_file_path": None,
               "format": "%(asctime)s - %(name)s - %(levelname)s - %(message)s"},
    "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6},
                  "lsf": {"sigma": 0.5}, "noise": {"signal_to_noise": 1, "noise_distribution": "normal"}},
    "cosmology": {"name": "PLANCK15"},
    "galaxy": {"dist_z": 0.1, "rotation": {"type": "edge-on"}},
    "ssp": {"template": {"name": "BruzualCharlot2003"}},
}


def _cfg(**over):
    c = copy.deepcopy(CONFIG)
    for k, v in over.items():
        c[k] = v
    return c


# ---- reference tests/test_core_psf.py -----------------------------------------------------------
@pytest.mark.parametrize("tel,msg", [
    ({"name": "MUSE"}, "PSF configuration not found in telescope configuration"),
    ({"name": "MUSE", "psf": {}}, "PSF name not found in telescope configuration"),
    ({"name": "MUSE", "psf": {"name": "gaussian", "sigma": 1.0}}, "PSF size not found in telescope configuration"),
    ({"name": "MUSE", "psf": {"name": "gaussian", "size": 5}}, "PSF sigma not found in telescope configuration"),
    ({"name": "MUSE", "psf": {"name": "airy", "size": 5, "sigma": 1.0}}, "Unknown PSF kernel name: airy"),
])
def test_psf_config_errors(tel, msg):
    with pytest.raises(ValueError, match=msg):
        get_convolve_psf(_cfg(telescope=tel))


# ---- reference tests/test_core_lsf.py -----------------------------------------------------------
@pytest.mark.parametrize("tel,msg", [
    ({"name": "MUSE"}, "LSF configuration not found in telescope configuration"),
    ({"name": "MUSE", "lsf": {}}, "LSF sigma size not found in telescope configuration"),
])
def test_lsf_config_errors(tel, msg):
    with pytest.raises(ValueError, match=msg):
        get_convolve_lsf(_cfg(telescope=tel))


# ---- reference tests/test_core_ssp.py:23-31 ------------------------------------------------------
@pytest.mark.parametrize("cfg,msg", [
    ({}, "Configuration does not contain 'ssp' field"),
    ({"ssp": {}}, "Configuration does not contain 'template' field"),
    ({"ssp": {"template": {}}}, "Configuration does not contain 'name' field"),
])
def test_ssp_config_errors(cfg, msg):
    with pytest.raises(ValueError, match=msg):
        get_ssp(cfg)


def test_ssp_template_and_default_method():
    ssp = get_ssp(CONFIG)
    assert ssp.flux.shape == (6, 221, 842) and ssp.flux.dtype == np.float32
    assert get_method(CONFIG) == "cubic"  # rubix/core/ssp.py:57-62
    assert get_method(_cfg(ssp={"template": {"name": "BruzualCharlot2003"}, "method": "linear"})) == "linear"
    with pytest.raises(ValueError, match="not found in the supported configuration"):
        get_ssp({"ssp": {"template": {"name": "nope"}}})


def test_closure_names_match_pipeline_nodes():
    """rubix/pipeline/abstract_pipeline.py:80-82: transformers are registered by ``__name__``."""
    from rubix_b200.core import (get_calculate_datacube, get_calculate_spectra, get_doppler_shift_and_resampling,
                                 get_filter_particles, get_reshape_data, get_scale_spectrum_by_mass,
                                 get_spaxel_assignment)
    fns = [get_filter_particles(CONFIG), get_spaxel_assignment(CONFIG), get_reshape_data(CONFIG),
           get_calculate_spectra(CONFIG), get_scale_spectrum_by_mass(CONFIG),
           get_doppler_shift_and_resampling(CONFIG), get_calculate_datacube(CONFIG), get_convolve_psf(CONFIG),
           get_convolve_lsf(CONFIG)]
    names = [f.__name__ for f in fns]
    assert names == ["filter_particles", "spaxel_assignment", "reshape_data", "calculate_spectra",
                     "scale_spectrum_by_mass", "doppler_shift_and_resampling", "calculate_datacube",
                     "convolve_psf", "convolve_lsf"]
    assert set(names) <= set(PIPELINES["calc_ifu"]["Transformers"])
    for f in fns:  # rubix/pipeline/transformer.py:18 deep-copies every transformer
        assert copy.deepcopy(f).__name__ == f.__name__


# ---- telescope grids ------------------------------------------------------------------------------
def test_muse_telescope_and_edges(muse_wave):
    t = get_telescope(CONFIG)
    assert t.sbin == 25 and t.wave_seq.dtype == np.float32
    assert np.array_equal(t.wave_seq, muse_wave)
    edges = get_spatial_bin_edges(CONFIG)
    assert edges.dtype == np.float32 and len(edges) in (26, 27)
    assert abs(float(edges[0]) + 4.7619) < 2e-3 and np.all(np.diff(edges) > 0)
    with pytest.raises(ValueError, match="Telescope nope not found in config"):
        get_telescope(_cfg(telescope={"name": "nope"}))


def test_custom_large_fov_telescope():
    tel = {"name": "MUSE_WIDE", "psf": CONFIG["telescope"]["psf"], "lsf": {"sigma": 0.5},
           "custom": {"MUSE_WIDE": dict(fov=30.0, spatial_res=0.2, wave_range=[4700.15, 9351.4], wave_res=1.25,
                                        lsf_fwhm=2.51, signal_to_noise=None, aperture_type="square",
                                        pixel_type="square")}}
    t = get_telescope(_cfg(telescope=tel))
    assert t.sbin == 150 and len(t.wave_seq) == 3721


# ---- pipeline ordering (reference tests/test_pipeline.py, tests/test_core_pipeline.py:80-85) -------
def test_pipeline_order_and_errors():
    order = cpipe.order_by_depends_on(PIPELINES["calc_ifu"])
    assert order[0] == "rotate_galaxy" and order[-1] == "apply_noise"
    assert order.index("calculate_spectra") < order.index("scale_spectrum_by_mass") < \
        order.index("doppler_shift_and_resampling") < order.index("calculate_datacube") < \
        order.index("convolve_psf") < order.index("convolve_lsf")
    with pytest.raises(ValueError, match="Pipeline nope not found in the configuration"):
        get_pipeline_config("nope")
    two_roots = {"Transformers": {"a": {"depends_on": None}, "b": {"depends_on": None}}}
    with pytest.raises(ValueError, match="exactly one starting point"):
        cpipe.order_by_depends_on(two_roots)
    branch = {"Transformers": {"a": {"depends_on": None}, "b": {"depends_on": "a"}, "c": {"depends_on": "a"}}}
    with pytest.raises(ValueError, match="Branching is not allowed"):
        cpipe.order_by_depends_on(branch)


# ---- padding / sharding (reference tests/test_core_data.py:226-248) -------------------------------
@pytest.mark.parametrize("n,n_dev", [(10, 2), (11, 2), (7, 3), (1, 4), (8, 8)])
def test_reshape_array_padding(n, n_dev):
    a = np.arange(1, 3 * n + 1, dtype=np.float32).reshape(n, 3)
    r = cdata.reshape_array(a, n_dev).numpy()
    per = -(-n // n_dev)
    assert r.shape == (n_dev, per, 3)
    flat = r.reshape(-1, 3)
    assert np.array_equal(flat[:n], a) and not flat[n:].any()
    r1 = cdata.reshape_array(a[:, 0], n_dev).numpy()
    assert r1.shape == (n_dev, per)


@pytest.mark.parametrize("n,world", [(10, 1), (10, 3), (7, 8), (1000003, 8), (0, 4)])
def test_shard_ranges_cover_once(n, world):
    seen = np.zeros(n, dtype=int)
    for r in range(world):
        lo, hi = parallel.shard_range(n, r, world)
        assert 0 <= lo <= hi <= n
        seen[lo:hi] += 1
    assert (seen == 1).all()
    w = [parallel.wavelength_slab(3721, r, world) for r in range(world)]
    assert w[0][0] == 0 and w[-1][1] == 3721 and all(w[i][1] == w[i + 1][0] for i in range(world - 1))


# ---- numpy-only HDF5 reader against the committed fixture of the reference's galaxy file ----------
def test_tng_fixture_fields(tng_subset):
    for k in ("coords", "velocity", "mass", "metallicity", "age"):
        assert tng_subset[k].dtype == np.float32
    assert tng_subset["coords"].shape[1] == 3


# ---- rubix/core/rotation.py ------------------------------------------------------------------------
@pytest.mark.parametrize("galaxy,msg", [
    ({"dist_z": 0.1}, "Rotation information not provided in galaxy config"),
    ({"rotation": {"beta": 0, "gamma": 0}}, "alpha not provided in rotation information"),
    ({"rotation": {"alpha": 0, "gamma": 0}}, "beta not provided in rotation information"),
    ({"rotation": {"alpha": 0, "beta": 0}}, "gamma not provided in rotation information"),
    ({"rotation": {"type": "sideways"}}, "Invalid type provided in rotation information"),
])
def test_rotation_config_errors(galaxy, msg):
    """tests/test_core_rotation.py:15-46."""
    from rubix_b200.core import get_galaxy_rotation
    with pytest.raises(ValueError, match=msg):
        get_galaxy_rotation({"galaxy": galaxy, "data": {"args": {"particle_type": ["stars"]}}})


def test_rotation_factory_name():
    from rubix_b200.core import get_galaxy_rotation
    for rot in ({"type": "face-on"}, {"type": "edge-on"}, {"alpha": 10, "beta": 20, "gamma": 30}):
        fn = get_galaxy_rotation({"galaxy": {"rotation": rot}, "data": {"args": {"particle_type": ["stars"]}}})
        assert fn.__name__ == "rotate_galaxy"  # the YAML node name (pipeline_config.yml)


# ---- rubix/core/noise.py ---------------------------------------------------------------------------
@pytest.mark.parametrize("tel,msg", [
    ({}, "Noise information not provided in telescope config"),
    ({"noise": {"noise_distribution": "normal"}}, "Signal to noise information not provided in noise config"),
    ({"noise": {"signal_to_noise": 1}}, "Noise distribution not provided in noise config"),
])
def test_noise_config_errors(tel, msg):
    from rubix_b200.core import get_apply_noise
    with pytest.raises(ValueError, match=msg):
        get_apply_noise({"telescope": tel})


def test_noise_factory_name():
    from rubix_b200.core import get_apply_noise
    fn = get_apply_noise({"telescope": {"noise": {"signal_to_noise": 1, "noise_distribution": "normal"}}})
    assert fn.__name__ == "apply_noise"


def test_hexagonal_aperture_matches_the_reference_loop():
    """rubix/telescope/apertures.py:12-40 restated as the reference writes it (a double loop with .at[].set)."""
    import numpy as np
    from rubix_b200.telescope import TelescopeFactory, hexagonal_aperture

    def ref(sbin):
        ap = np.zeros((sbin, sbin))
        xc = yc = sbin / 2 + 0.5
        for x in range(1, sbin + 1):
            for y in range(1, sbin + 1):
                xx, yy = x - xc, y - yc
                rr = (2 * (sbin / 4) * (sbin * np.sqrt(3) / 4)) - ((sbin / 4) * abs(yy)) - ((sbin * np.sqrt(3) / 4) * abs(xx))
                if rr >= 0 and abs(xx) < sbin / 2 and abs(yy) < sbin * np.sqrt(3) / 4:
                    ap[x - 1, y - 1] = 1
        return ap.flatten()

    for n in (4, 5, 24, 25, 74):
        assert np.array_equal(hexagonal_aperture(n), ref(n))
    tel = TelescopeFactory({"HEX": dict(fov=5.0, spatial_res=0.2, wave_range=[4700.15, 9351.4], wave_res=1.25,
                                         lsf_fwhm=2.51, aperture_type="hexagonal", pixel_type="square")}).create_telescope("HEX")
    assert np.array_equal(np.asarray(tel.aperture_region), ref(25))


def test_rubix_pipeline_strict_mode_validates_like_the_reference(tng_subset):
    """config["b200"]["strict"]: all twelve factories are called unconditionally, as rubix/core/pipeline.py:105-133 does,
    so a configuration without galaxy.rotation / ssp.dust / telescope.noise raises the factory's own error (by default the
    mirror skips those stages with a warning)."""
    import copy
    from rubix_b200 import core
    base = {
        "pipeline": {"name": "calc_ifu"}, "cosmology": {"name": "PLANCK15"}, "galaxy": {"dist_z": 0.1},
        "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}, "lsf": {"sigma": 0.5}},
        "ssp": {"template": {"name": "BruzualCharlot2003"}}, "data": {"args": {"particle_type": ["stars"]}},
    }
    rd = core.make_rubix_data(**tng_subset, device=False)
    assert len(core.RubixPipeline(copy.deepcopy(base), data=rd).assemble()) == 9            # relaxed: three stages skipped
    strict = copy.deepcopy(base)
    strict["b200"] = {"strict": True}
    with pytest.raises(ValueError, match="Rotation information not provided in galaxy config"):
        core.RubixPipeline(strict, data=rd).assemble()
    strict["galaxy"]["rotation"] = {"type": "face-on"}
    with pytest.raises(ValueError, match="Dust configuration not found in config file."):
        core.RubixPipeline(strict, data=rd).assemble()
    strict["ssp"]["dust"] = {"extinction_model": "Cardelli89", "Rv": 3.1, "dust_grain_density": 3.5}
    with pytest.raises(ValueError, match="Noise information not provided in telescope config"):
        core.RubixPipeline(strict, data=rd).assemble()
    strict["telescope"]["noise"] = {"signal_to_noise": 10, "noise_distribution": "normal"}
    names = [fn.__name__ for fn in core.RubixPipeline(strict, data=rd).assemble()]
    assert len(names) == 11 and names[0] == "rotate_galaxy" and names[-1] == "apply_noise"   # calc_ifu has no dust node
