"""jax.ffi binding of the path (the JAX-side half of rubix_b200/csrc/jax_ffi.cc).

Needs jax >= 0.4.38 and ``librubix_b200_jax.so`` (``make -C rubix_b200/csrc jax_ffi``); neither is
available in the build image, so this module is not imported by the package and has not been
executed.  It shows the binding a rubix maintainer would use: the handlers are registered once and
``build_cube`` / ``psf_lsf`` become traceable functions that can sit inside rubix's ``jax.jit``.
The plan (device tables of one configuration) is created eagerly, outside jit, through the ctypes
binding in :mod:`rubix_b200._lib`; only its address travels as a static attribute.
"""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_TARGETS = {"rbx_spaxel_assign": "RbxSpaxelAssign", "rbx_build_cube": "RbxBuildCube",
            "rbx_ssp_lookup": "RbxSspLookup", "rbx_doppler_resample": "RbxDopplerResample",
            "rbx_psf_lsf": "RbxPsfLsf", "rbx_assign_build_cube": "RbxAssignBuildCube",
            "rbx_psf_lsf_taps": "RbxPsfLsfTaps", "rbx_dust_av": "RbxDustAv",
            "rbx_apply_extinction": "RbxApplyExtinction"}
_registered = False


def register():
    global _registered
    if _registered:
        return
    import jax
    lib = ctypes.CDLL(os.path.join(_HERE, "librubix_b200_jax.so"))
    for name, symbol in _TARGETS.items():
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    _registered = True


def spaxel_assign(coords, edges):
    import jax
    import jax.numpy as jnp
    register()
    out = jax.ShapeDtypeStruct(coords.shape[:1], jnp.int32)
    return jax.ffi.ffi_call("rbx_spaxel_assign", out)(coords, edges)


def build_cube(plan_handle: int, workspace_bytes: int, velocity, mass, metallicity, age, pixel, num_spaxels: int,
               n_wave: int):
    """calculate_spectra -> scale_spectrum_by_mass -> doppler_shift_and_resampling -> calculate_datacube
    (rubix/core/ifu.py) as one custom call.  ``plan_handle``/``workspace_bytes`` come from
    ``ops.Plan(...).handle.value`` and ``rbx_build_cube_workspace_bytes``."""
    import jax
    import jax.numpy as jnp
    import numpy as np
    register()
    outs = (jax.ShapeDtypeStruct((num_spaxels, num_spaxels, n_wave), jnp.float32),
            jax.ShapeDtypeStruct((workspace_bytes,), jnp.uint8))
    cube, _ = jax.ffi.ffi_call("rbx_build_cube", outs)(velocity, mass, metallicity, age, pixel,
                                                       plan=np.int64(plan_handle),
                                                       num_spaxels=np.int32(num_spaxels))
    return cube


def psf_lsf(cube, psf_kernel, lsf_kernel, ext: int = 12):
    import jax
    import numpy as np
    register()
    return jax.ffi.ffi_call("rbx_psf_lsf", jax.ShapeDtypeStruct(cube.shape, cube.dtype))(
        cube, psf_kernel, lsf_kernel, ext=np.int32(ext))


def assign_build_cube(plan_handle: int, workspace_bytes: int, coords, edges, velocity, mass, metallicity, age,
                      num_spaxels: int, n_wave: int, apply_filter: bool = True):
    """filter_particles + spaxel_assignment + the four fused ``ifu`` stages as one custom call."""
    import jax
    import jax.numpy as jnp
    import numpy as np
    register()
    outs = (jax.ShapeDtypeStruct((num_spaxels, num_spaxels, n_wave), jnp.float32),
            jax.ShapeDtypeStruct((workspace_bytes,), jnp.uint8))
    cube, _ = jax.ffi.ffi_call("rbx_assign_build_cube", outs)(
        coords, edges, velocity, mass, metallicity, age, plan=np.int64(plan_handle),
        num_spaxels=np.int32(num_spaxels), apply_filter=np.int32(1 if apply_filter else 0))
    return cube


def psf_lsf_taps(cube, psf_kernel, lsf_kernel, ext: int = 12):
    """PSF + LSF with the (host, config-constant) taps passed as static attributes."""
    import jax
    import numpy as np
    register()
    pk = np.ascontiguousarray(psf_kernel, dtype=np.float32)
    lk = np.ascontiguousarray(lsf_kernel, dtype=np.float32).reshape(-1)
    return jax.ffi.ffi_call("rbx_psf_lsf_taps", jax.ShapeDtypeStruct(cube.shape, cube.dtype))(
        cube, psf=pk.reshape(-1), psf_size=np.int32(pk.shape[0]), lsf=lk, ext=np.int32(ext))


def dust_av(workspace_bytes: int, gas_coords, gas_pixel, gas_mass, gas_metals, star_coords, star_pixel, n_spaxels: int,
            dust_to_gas, ext_const: float, spaxel_area: float):
    """A_V per star for ``apply_spaxel_extinction`` (rubix/spectra/dust/dust_extinction.py:240-337) as one custom
    call; ``workspace_bytes`` from ``rbx_dust_av_workspace_bytes(n_gas, n_spaxels)``."""
    import jax
    import jax.numpy as jnp
    import numpy as np
    register()
    outs = (jax.ShapeDtypeStruct(star_pixel.shape, jnp.float32), jax.ShapeDtypeStruct((workspace_bytes,), jnp.uint8))
    av, _ = jax.ffi.ffi_call("rbx_dust_av", outs)(
        gas_coords, gas_pixel, gas_mass, gas_metals, star_coords, star_pixel, n_spaxels=np.int32(n_spaxels),
        dust_to_gas=np.asarray(dust_to_gas, dtype=np.float32), ext_const=np.float32(ext_const),
        spaxel_area=np.float32(spaxel_area))
    return av


def apply_extinction(spectra, av, axav):
    """``spectra * 10**(-0.4 * axav * av[:, None])`` (dust_extinction.py:341-356) as one custom call."""
    import jax
    register()
    return jax.ffi.ffi_call("rbx_apply_extinction", jax.ShapeDtypeStruct(spectra.shape, spectra.dtype))(spectra, av, axav)
