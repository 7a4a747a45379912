"""ctypes front-end of oracle/rubix_oracle.c (TEST INFRASTRUCTURE ONLY -- see that file's header).

``build()`` compiles it with the Makefile next to it; ``particles_to_cube`` etc. mirror the numpy
oracle's signatures so tests can swap one for the other.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librubix_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rubix_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def particles_to_cube(coords, velocity, mass, metallicity, age, spatial_bin_edges, num_spaxels,
                      ssp_metallicity, ssp_age, ssp_wavelength, ssp_flux, target_wavelength,
                      redshift, method="cubic", direction="z", dtype=np.float32,
                      apply_filter=True, n_threads=1):
    """Same contract as oracle.rubix_oracle.particles_to_cube (returns cube (S,S,W) only)."""
    coords, velocity = _f32(coords), _f32(velocity)
    mass, metallicity, age = _f32(mass), _f32(metallicity), _f32(age)
    edges = _f32(spatial_bin_edges)
    zg, ag, wl, fl = _f32(ssp_metallicity), _f32(ssp_age), _f32(ssp_wavelength), _f32(ssp_flux)
    t = _f32(target_wavelength)
    n = coords.shape[0]
    W = len(t)
    is64 = np.dtype(dtype) == np.float64
    cube = np.zeros((num_spaxels, num_spaxels, W), dtype=np.float64 if is64 else np.float32)
    fn = getattr(lib(), ("rbxo64_" if is64 else "rbxo32_") + "particles_to_cube")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    rc = fn(_p(coords), _p(velocity), _p(mass), _p(metallicity), _p(age), n, _p(edges), len(edges),
            int(num_spaxels), _p(zg), len(zg), _p(ag), len(ag), _p(wl), len(wl), _p(fl), _p(t), W,
            float(redshift), 1 if method == "cubic" else 0, {"x": 0, "y": 1, "z": 2}[direction],
            1 if apply_filter else 0, int(n_threads), _p(cube))
    if rc != 0:
        raise RuntimeError(f"oracle failed: {rc}")
    return cube


def spaxel_assign(coords, spatial_bin_edges):
    coords, edges = _f32(coords), _f32(spatial_bin_edges)
    n = coords.shape[0]
    idx = np.zeros(n, dtype=np.int32)
    mask = np.zeros(n, dtype=np.uint8)
    fn = lib().rbxo32_spaxel_assign
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    fn(_p(coords), n, _p(edges), len(edges), _p(idx), _p(mask))
    return idx, mask.astype(bool)


def apply_psf(cube, kernel):
    cube = np.ascontiguousarray(cube)
    is64 = cube.dtype == np.float64
    if not is64:
        cube = cube.astype(np.float32)
    k = np.ascontiguousarray(kernel, dtype=cube.dtype)
    out = np.empty_like(cube)
    fn = getattr(lib(), ("rbxo64_" if is64 else "rbxo32_") + "apply_psf")
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    fn(_p(cube), _p(out), cube.shape[0], cube.shape[1], cube.shape[2], _p(k), k.shape[0], k.shape[1])
    return out


def apply_lsf(cube, kernel, ext=12):
    cube = np.ascontiguousarray(cube)
    is64 = cube.dtype == np.float64
    if not is64:
        cube = cube.astype(np.float32)
    k = np.ascontiguousarray(kernel, dtype=cube.dtype)
    out = np.empty_like(cube)
    rows = int(np.prod(cube.shape[:-1]))
    fn = getattr(lib(), ("rbxo64_" if is64 else "rbxo32_") + "apply_lsf")
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int]
    fn(_p(cube), _p(out), rows, cube.shape[-1], _p(k), len(k), int(ext))
    return out
