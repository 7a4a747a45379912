"""rubix/core/ifu.py mirrors: the four factories of the particle -> cube path.

Each factory has the reference's signature ``get_*(config: dict) -> Callable`` and returns a
closure ``f(rubixdata) -> rubixdata`` whose ``__name__`` equals the pipeline node name
(rubix/pipeline/abstract_pipeline.py:80-82) and which mutates and returns its argument
(rubix/core/ifu.py:118,155,291,337).

Two execution modes, chosen per call by ``config["b200"]["fused"]`` (True / False / "auto"):

* staged  -- every stage launches its own CUDA kernel and ``stars.spectra`` holds the same
  intermediate the reference would hold: (1, P, L) after calculate_spectra and
  scale_spectrum_by_mass, (1, P, W) after doppler_shift_and_resampling.  This is what the stepwise
  notebook flow and the reference's contract tests observe.
* fused   -- the first three stages only record what has to be done (``stars.spectra`` becomes a
  :class:`DeferredSpectra` that can still be materialised on request) and ``calculate_datacube``
  launches the fused kernel, which never writes the 14.9 kB/particle intermediate.

"auto" uses the staged mode while the (P, W) intermediate stays below ``AUTO_FUSE_BYTES``.
"""

from __future__ import annotations

import os

from typing import Callable

import numpy as np

from ..config import rubix_config
from ..logger import get_logger
from .data import RubixData
from .ssp import get_method, get_ssp
from .telescope import get_telescope

AUTO_FUSE_BYTES = 256 << 20

_plan_cache = {}


def _plan_key(config: dict):
    tel = config["telescope"]
    return (config["ssp"]["template"]["name"], get_method(config), tel["name"], repr(tel.get("custom")),
            float(config["galaxy"]["dist_z"]), rubix_config["ifu"]["doppler"]["velocity_direction"])


def get_plan(config: dict):
    """Device tables for this configuration (SSP template, telescope wave_seq, redshift, method,
    Doppler direction).  Cached per process and device; closures only keep the config, so they stay
    ``deepcopy``-able like the reference's (rubix/pipeline/transformer.py:18)."""
    import torch
    from .. import ops
    key = _plan_key(config) + (torch.cuda.current_device(),)
    plan = _plan_cache.get(key)
    if plan is None:
        ssp = get_ssp(config)
        telescope = get_telescope(config)
        plan = ops.Plan(ssp.metallicity, ssp.age, ssp.wavelength, ssp.flux, telescope.wave_seq,
                        config["galaxy"]["dist_z"], method=get_method(config),
                        direction=rubix_config["ifu"]["doppler"]["velocity_direction"])
        _plan_cache[key] = plan
    return plan


class DeferredSpectra:
    """What ``stars.spectra`` holds in fused mode: the recipe instead of the (n_dev, P, L|W) array.
    ``materialize()`` runs the stage kernels and returns the array the reference would hold."""

    def __init__(self, config: dict, stars, n: int):
        self._config, self._stars, self.n = config, stars, n
        self.scaled = False
        self.resampled = False
        self.extinction = None   # (A_V per star, A(lambda)/A(V) per channel) once calculate_extinction has run
        self.dtype = np.float32

    @property
    def shape(self):
        plan = get_plan(self._config)
        return (1, self.n, plan.W if self.resampled else plan.L)

    @property
    def ndim(self):
        return 3

    def materialize(self):
        from .. import ops
        plan = get_plan(self._config)
        st = self._stars
        spec = ops.ssp_lookup(plan, st.metallicity.reshape(-1), st.age.reshape(-1))
        if self.scaled:
            spec = ops.scale_by_mass(spec, st.mass.reshape(-1))
        if self.resampled:
            spec = ops.doppler_resample(plan, spec, st.velocity.reshape(-1, 3))
        if self.extinction is not None:
            spec = ops.apply_extinction(spec, *self.extinction, out=spec)
        return spec.unsqueeze(0)

    def __array__(self, dtype=None, copy=None):
        a = self.materialize().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return f"DeferredSpectra(shape={self.shape}, scaled={self.scaled}, resampled={self.resampled})"


def _use_fused(config: dict, n: int, W: int) -> bool:
    mode = config.get("b200", {}).get("fused", "auto") if isinstance(config.get("b200", {}), dict) else "auto"
    if mode is True or mode is False:
        return mode
    return n * W * 4 > AUTO_FUSE_BYTES


def get_calculate_spectra(config: dict) -> Callable:
    """rubix/core/ifu.py:29-123: SSP lookup of every star -> ``stars.spectra`` (1, P, L)."""
    logger = get_logger(config.get("logger", None))
    get_ssp(config)  # same early validation / errors as the reference factory

    def calculate_spectra(rubixdata: RubixData) -> RubixData:
        from .. import ops
        logger.info("Calculating IFU cube...")
        st = rubixdata.stars
        st.metallicity, st.age = ops.dev(st.metallicity), ops.dev(st.age)
        logger.debug(f"Input shapes: Metallicity: {len(st.metallicity)}, Age: {len(st.age)}")
        plan = get_plan(config)
        n = st.metallicity.numel()
        if _use_fused(config, n, plan.W):
            st.spectra = DeferredSpectra(config, st, n)
        else:
            st.spectra = ops.ssp_lookup(plan, st.metallicity.reshape(-1), st.age.reshape(-1)).unsqueeze(0)
        logger.debug(f"Calculation Finished! Spectra shape: {tuple(st.spectra.shape)}")
        return rubixdata

    return calculate_spectra


def get_scale_spectrum_by_mass(config: dict) -> Callable:
    """rubix/core/ifu.py:127-159: ``spectra * mass[..., None]``."""
    logger = get_logger(config.get("logger", None))

    def scale_spectrum_by_mass(rubixdata: RubixData) -> RubixData:
        from .. import ops
        logger.info("Scaling Spectra by Mass...")
        st = rubixdata.stars
        st.mass = ops.dev(st.mass)
        if isinstance(st.spectra, DeferredSpectra):
            st.spectra.scaled = True
        else:
            spec = ops.dev(st.spectra)
            lead = st.mass.shape
            if spec.shape[:-1] != lead:  # (1, P, L) x (n_dev, P, 1) broadcast of the reference
                spec = spec.expand(lead + spec.shape[-1:]).contiguous()
            st.spectra = ops.scale_by_mass(spec, st.mass.reshape(-1)).reshape(spec.shape)
        return rubixdata

    return scale_spectrum_by_mass


def get_doppler_shift_and_resampling(config: dict) -> Callable:
    """rubix/core/ifu.py:226-295: per-particle Doppler shift of the (1+z)-shifted SSP wavelengths and
    flux-conserving resampling onto ``telescope.wave_seq`` -> ``spectra`` (n_dev, P, W)."""
    logger = get_logger(config.get("logger", None))
    _ = config["galaxy"]["dist_z"]  # KeyError like the reference when missing
    get_telescope(config)
    get_ssp(config)

    def doppler_shift_and_resampling(rubixdata: RubixData) -> RubixData:
        from .. import ops
        for particle_name in ["stars", "gas"]:
            part = getattr(rubixdata, particle_name)
            if part.spectra is None:
                continue
            logger.info("Doppler shifting and resampling spectra...")
            part.velocity = ops.dev(part.velocity)
            if isinstance(part.spectra, DeferredSpectra):
                part.spectra.resampled = True
                continue
            plan = get_plan(config)
            spec = ops.dev(part.spectra)
            out = ops.doppler_resample(plan, spec.reshape(-1, spec.shape[-1]), part.velocity.reshape(-1, 3))
            part.spectra = out.reshape(spec.shape[:-1] + (plan.W,))
        return rubixdata

    return doppler_shift_and_resampling


#: particles per pass of the dusty cube (bounds the (n, L) SSP spectra held at once: 3.4 kB per particle)
DUSTY_CHUNK = 1 << 20


def _dusty_cube(plan, st, mass, pix, num_spaxels: int, extinction, impl=None):
    """calc_dusty_ifu with deferred spectra: SSP lookup and mass scaling per chunk of particles, then
    resampling + extinction + per-spaxel sum in one kernel (rbx_build_cube_dusty).  ``impl``
    (``config["b200"]["dusty_impl"]``, else ``RBX_DUSTY_IMPL``): "binned" (default: through the knot-based cube
    kernel, no cube atomics) or "onepass"."""
    from .. import ops
    av, axav = extinction
    met, age, vel = st.metallicity.reshape(-1), st.age.reshape(-1), st.velocity.reshape(-1, 3)
    n = met.numel()
    if (impl or os.environ.get("RBX_DUSTY_IMPL", "binned")) == "binned":
        cube = ops.build_cube_dusty_binned(plan, vel, mass, met, age, pix, num_spaxels, av, axav)
        if cube is not None:
            return cube
    cube = None
    for lo in range(0, max(n, 1), DUSTY_CHUNK):
        hi = min(n, lo + DUSTY_CHUNK)
        spec = ops.ssp_lookup(plan, met[lo:hi], age[lo:hi])
        spec = ops.scale_by_mass(spec, mass[lo:hi])
        part = ops.build_cube_dusty(plan, spec, vel[lo:hi], pix[lo:hi], num_spaxels, av[lo:hi], axav)
        cube = part if cube is None else cube.add_(part)
    return cube


def get_calculate_datacube(config: dict) -> Callable:
    """rubix/core/ifu.py:299-341: per-spaxel sum of the resampled spectra -> ``stars.datacube``
    (S, S, W).  With ``torch.distributed`` initialised and ``config["b200"]["distributed"]`` true the
    per-rank partial cubes are summed with one NCCL all-reduce (the reference's
    ``jnp.sum(ifu_cubes, axis=0)`` over its device axis)."""
    logger = get_logger(config.get("logger", None))
    telescope = get_telescope(config)
    num_spaxels = int(telescope.sbin)

    def calculate_datacube(rubixdata: RubixData) -> RubixData:
        from .. import ops
        logger.info("Calculating Data Cube...")
        st = rubixdata.stars
        pix = ops.dev(st.pixel_assignment, dtype=__import__("torch").int32).reshape(-1)
        if isinstance(st.spectra, DeferredSpectra):
            d = st.spectra
            if not d.resampled:
                raise ValueError("calculate_datacube: spectra are not on the telescope wavelength grid "
                                 "(doppler_shift_and_resampling has not run)")
            plan = get_plan(config)
            mass = st.mass.reshape(-1) if d.scaled else __import__("torch").ones_like(st.metallicity.reshape(-1))
            if d.extinction is not None:
                b200 = config.get("b200") if isinstance(config.get("b200"), dict) else {}
                cube = _dusty_cube(plan, st, mass, pix, num_spaxels, d.extinction, impl=b200.get("dusty_impl"))
            else:
                cube = ops.build_cube(plan, st.velocity.reshape(-1, 3), mass, st.metallicity.reshape(-1),
                                      st.age.reshape(-1), pix, num_spaxels)
        else:
            spec = ops.dev(st.spectra)
            # b200.deterministic: the staged sum in the reference's CPU order (sorted runs, no atomics), bit-reproducible
            det = isinstance(config.get("b200"), dict) and bool(config["b200"].get("deterministic"))
            cube = ops.segment_sum(spec.reshape(-1, spec.shape[-1]), pix, num_spaxels * num_spaxels, deterministic=det)
            cube = cube.reshape(num_spaxels, num_spaxels, spec.shape[-1])
        if isinstance(config.get("b200"), dict) and config["b200"].get("distributed"):
            from ..parallel import allreduce_cube
            cube = allreduce_cube(cube)
        logger.debug(f"Datacube Shape: {tuple(cube.shape)}")
        st.datacube = cube
        return rubixdata

    return calculate_datacube
