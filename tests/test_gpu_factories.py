"""GPU tests of the drop-in factories (the rubix.core mirror): the staged path reproduces the
reference's observable intermediates (tests/test_core_ifu.py shapes / exact mass scaling / zero spectra
for out-of-grid Z), the fused path gives the same cube, and RubixPipeline runs the whole chain."""

import copy

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import c_oracle  # noqa: E402
from oracle import rubix_oracle as orc  # noqa: E402

CONFIG = {
    "pipeline": {"name": "calc_ifu"},
    "logger": {"log_level": "WARNING", "log_file_path": None,
               "format": "%(asctime)s - %(name)s - %(levelname)s - %(message)s"},
    "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}, "lsf": {"sigma": 0.5}},
    "cosmology": {"name": "PLANCK15"},
    "galaxy": {"dist_z": 0.1},
    "ssp": {"template": {"name": "BruzualCharlot2003"}},
}


@pytest.fixture(scope="module")
def core():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rubix_b200 import core as _core
    return _core


def _run_chain(core, cfg, data):
    rd = core.make_rubix_data(**data)
    for get in (core.get_filter_particles, core.get_spaxel_assignment, core.get_reshape_data,
                core.get_calculate_spectra, core.get_scale_spectrum_by_mass,
                core.get_doppler_shift_and_resampling, core.get_calculate_datacube):
        rd = get(cfg)(rd)
    return rd


def _oracle_cube(cfg, data, bc03, muse_wave, method):
    from rubix_b200.core.telescope import get_spatial_bin_edges
    edges = get_spatial_bin_edges(cfg)
    nb = len(edges) - 1
    # ids are x + nb*y; the reference's segment_sum keeps ids < sbin^2 (rubix/core/ifu.py:320)
    ref = c_oracle.particles_to_cube(data["coords"], data["velocity"], data["mass"], data["metallicity"],
                                     data["age"], edges, 25, bc03["metallicity"], bc03["age"], bc03["wavelength"],
                                     bc03["flux"], muse_wave, 0.1, method=method, dtype=np.float64, n_threads=8)
    return ref, nb


def test_staged_intermediates_match_reference_contract(core, bc03, muse_wave, tng_subset):
    cfg = copy.deepcopy(CONFIG)
    cfg["b200"] = {"fused": False}
    d = {k: v[:300].copy() for k, v in tng_subset.items()}
    d["metallicity"][:5] = 0.1  # above the grid -> zero spectra (tests/test_core_ifu.py:207-246)
    rd = core.make_rubix_data(**d)
    rd = core.get_filter_particles(cfg)(rd)
    rd = core.get_spaxel_assignment(cfg)(rd)
    rd = core.get_reshape_data(cfg)(rd)
    assert rd.stars.coords.shape == (1, 300, 3) and rd.stars.pixel_assignment.dtype == torch.int32
    rd = core.get_calculate_spectra(cfg)(rd)
    spec = rd.stars.spectra
    assert tuple(spec.shape) == (1, 300, 842)
    assert not torch.isnan(spec).any()
    assert float(spec[0, :5].abs().max()) == 0.0
    before = spec.clone()
    rd = core.get_scale_spectrum_by_mass(cfg)(rd)
    # tests/test_core_ifu.py:277-279: exactly spectra * mass[..., None]
    assert torch.equal(rd.stars.spectra, before * rd.stars.mass[..., None])
    rd = core.get_doppler_shift_and_resampling(cfg)(rd)
    assert tuple(rd.stars.spectra.shape) == (1, 300, 3721) and not torch.isnan(rd.stars.spectra).any()
    rd = core.get_calculate_datacube(cfg)(rd)
    assert tuple(rd.stars.datacube.shape) == (25, 25, 3721)


@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_fused_and_staged_factories_agree_with_oracle(core, bc03, muse_wave, tng_subset, method):
    from helpers import cube_close as _cube_close, well_conditioned as _well_conditioned
    d = _well_conditioned(tng_subset, np.float32(1.1) * bc03["wavelength"], muse_wave)
    cubes = {}
    for fused in (False, True):
        cfg = copy.deepcopy(CONFIG)
        cfg["ssp"]["method"] = method
        cfg["b200"] = {"fused": fused}
        rd = _run_chain(core, cfg, d)
        if fused:
            from rubix_b200.core.ifu import DeferredSpectra
            assert isinstance(rd.stars.spectra, DeferredSpectra) and rd.stars.spectra.shape == (1, len(d["mass"]), 3721)
        cubes[fused] = rd.stars.datacube.cpu().numpy()
    ref, nb = _oracle_cube(cfg, d, bc03, muse_wave, method)
    # the edge count follows float32 rounding (26 edges, nb = 25, for MUSE at z = 0.1; a configuration that rounds to
    # 27 edges would give nb = 26 and ids >= 625, which both sides drop like segment_sum does)
    _cube_close(cubes[True], ref, f"factory fused {method} (nb={nb})")
    _cube_close(cubes[True], cubes[False].astype(np.float64), f"factory fused vs staged {method}", rtol_max=1e-5)


def test_rubix_pipeline_end_to_end(core, bc03, muse_wave, tng_subset):
    from helpers import cube_close as _cube_close, well_conditioned as _well_conditioned
    d = _well_conditioned(tng_subset, np.float32(1.1) * bc03["wavelength"], muse_wave)
    outs = {}
    for fused in (False, True):
        cfg = copy.deepcopy(CONFIG)
        cfg["b200"] = {"fused": fused}
        pipe = core.RubixPipeline(cfg, data=core.make_rubix_data(**d, device=False))
        out = pipe.run()
        cube = out.stars.datacube
        assert tuple(cube.shape) == (25, 25, 3721) and not torch.isnan(cube).any()
        outs[fused] = cube.cpu().numpy()
    raw, nb = _oracle_cube(cfg, d, bc03, muse_wave, "cubic")
    ref = orc.apply_lsf(orc.apply_psf(raw, orc.gaussian_kernel_2d(5, 5, 0.6).astype(np.float64)), 0.5, 1.25)
    _cube_close(outs[True], ref, f"RubixPipeline fused (nb={nb})")
    _cube_close(outs[True], outs[False].astype(np.float64), "RubixPipeline fused vs staged", rtol_max=1e-5)


def test_rubix_pipeline_all_stages(core, bc03, muse_wave, tng_subset):
    """calc_ifu with galaxy.rotation and telescope.noise in the config: rotate_galaxy -> ... -> apply_noise all on
    the device, against the oracle run stage by stage on the same inputs."""
    from helpers import cube_close as _cube_close, well_conditioned as _well_conditioned
    from rubix_b200.core.telescope import get_spatial_bin_edges
    d = _well_conditioned(tng_subset, np.float32(1.1) * bc03["wavelength"], muse_wave)
    cfg = copy.deepcopy(CONFIG)
    cfg["galaxy"]["rotation"] = {"alpha": 20.0, "beta": -35.0, "gamma": 70.0}
    cfg["telescope"]["noise"] = {"signal_to_noise": 50.0, "noise_distribution": "normal"}
    cfg["data"] = {"args": {"particle_type": ["stars"]}}
    cfg["b200"] = {"fused": True}
    rd = core.make_rubix_data(**d, device=False)
    rd.galaxy.halfmassrad_stars = 2.5
    pipe = core.RubixPipeline(cfg, data=rd)
    names = [fn.__name__ for fn in pipe.assemble()]
    assert names[0] == "rotate_galaxy" and names[-1] == "apply_noise" and len(names) == 11
    cube = pipe.run().stars.datacube
    assert tuple(cube.shape) == (25, 25, 3721) and not torch.isnan(cube).any()
    # oracle, stage by stage
    pos, vel, _ = orc.rotate_galaxy(d["coords"], d["velocity"], d["mass"], 2.5, 20.0, -35.0, 70.0)
    pos32, vel32 = pos.astype(np.float32), vel.astype(np.float32)
    edges = get_spatial_bin_edges(cfg)
    mass, met, age, _ = orc.filter_particles(pos32, d["mass"], d["metallicity"], d["age"], edges)
    raw = c_oracle.particles_to_cube(pos32, vel32, mass.astype(np.float32), met.astype(np.float32),
                                     age.astype(np.float32), edges, 25,
                                     bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1,
                                     method="cubic", dtype=np.float64, n_threads=8)
    conv = orc.apply_lsf(orc.apply_psf(raw, orc.gaussian_kernel_2d(5, 5, 0.6).astype(np.float64)), 0.5, 1.25)
    ref = orc.apply_noise(conv, 50.0, "normal")
    # particles the float32 rotation puts within rounding of a spaxel edge may land in the neighbouring spaxel:
    # compare in the sense of the north-star bound (relative to the cube's total flux)
    out = cube.cpu().numpy().astype(np.float64)
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).sum()
    assert np.abs(out.sum() - ref.sum()) <= 1e-5 * ref.sum()
