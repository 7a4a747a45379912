"""Device-level operators: thin Python wrappers over the C ABI (include/rubix_b200.h).

PyTorch is used only for device memory and streams.  Every function takes/returns CUDA float32 /
int32 tensors (numpy inputs are copied to the current device) and launches on the current stream.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib

METHODS = {"linear": 0, "cubic": 1}
DIRECTIONS = {"x": 0, "y": 1, "z": 2}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("rubix_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def dev(x, dtype=torch.float32) -> torch.Tensor:
    """Contiguous CUDA tensor of ``dtype`` from a tensor / numpy array / sequence."""
    _require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    return t.to(device="cuda", dtype=dtype).contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Plan:
    """Per-configuration device tables (``rbx_plan``): SSP template, telescope wavelength grid,
    redshift, ``ssp.method`` and the Doppler velocity direction."""

    def __init__(self, metallicity, age, wavelength, flux, target_wave, redshift: float,
                 method: str = "cubic", direction: str = "z"):
        _require_cuda()
        if method not in METHODS:
            raise ValueError(f"unknown ssp.method {method!r}: expected 'linear' or 'cubic'")
        if direction not in DIRECTIONS:
            raise ValueError(
                f"{direction} is not a valid direction. Supported directions are 'x', 'y', or 'z'.")
        f32 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float32)
        self.metallicity, self.age = f32(metallicity), f32(age)
        self.wavelength, self.flux, self.target_wave = f32(wavelength), f32(flux), f32(target_wave)
        nz, na, L = len(self.metallicity), len(self.age), len(self.wavelength)
        if self.flux.shape != (nz, na, L):
            raise ValueError(f"flux shape {self.flux.shape} != ({nz}, {na}, {L})")
        self.method, self.direction, self.redshift = method, direction, float(redshift)
        self.L, self.W = L, len(self.target_wave)
        self._h = C.c_void_p()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib().rbx_plan_create(
            C.byref(self._h), vp(self.metallicity), nz, vp(self.age), na, vp(self.wavelength), L,
            vp(self.flux), vp(self.target_wave), self.W, self.redshift, METHODS[method],
            DIRECTIONS[direction], _stream()))
        self.device = torch.cuda.current_device()

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().rbx_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spaxel_assign(coords, edges, with_mask: bool = False):
    """rubix/telescope/utils.py:138-151 -> int32 pixel ids (and the inclusive aperture mask)."""
    coords, edges = dev(coords), dev(edges)
    if coords.ndim != 2 or coords.shape[1] != 3:
        raise ValueError(f"coords must have shape (n, 3), got {tuple(coords.shape)}")
    n = coords.shape[0]
    pixel = torch.empty(n, dtype=torch.int32, device="cuda")
    mask = torch.empty(n, dtype=torch.uint8, device="cuda") if with_mask else None
    _lib.check(_lib.lib().rbx_spaxel_assign(_p(coords), n, _p(edges), edges.numel(), _p(pixel), _p(mask), _stream()))
    return (pixel, mask.bool()) if with_mask else pixel


def filter_particles(coords, edges, mass=None, metallicity=None, age=None):
    """rubix/core/telescope.py:155-174: zero mass / metallicity / age (in place) outside the aperture;
    returns the boolean mask."""
    coords, edges = dev(coords), dev(edges)
    n = coords.shape[0]
    for t in (mass, metallicity, age):
        if t is not None and not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32
                                  and t.is_contiguous() and t.numel() == n):
            raise ValueError("filter_particles works in place on contiguous CUDA float32 tensors of length n")
    mask = torch.empty(n, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().rbx_filter_particles(_p(coords), n, _p(edges), edges.numel(), _p(mass),
                                               _p(metallicity), _p(age), _p(mask), _stream()))
    return mask.bool()


def filter_and_assign(coords, edges, mass=None, metallicity=None, age=None) -> torch.Tensor:
    """filter_particles + spaxel_assignment in one pass.  Without mass / metallicity / age nothing is
    modified and particles outside the aperture get pixel -1 (dropped by build_cube / segment_sum)."""
    coords, edges = dev(coords), dev(edges)
    n = coords.shape[0]
    pixel = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().rbx_filter_and_assign(_p(coords), n, _p(edges), edges.numel(), _p(mass), _p(metallicity),
                                                _p(age), _p(pixel), None, _stream()))
    return pixel


def sort_by_spaxel(pixel, num_segments: int, with_sorted: bool = True, with_offsets: bool = True):
    """Stable device radix sort of particles by spaxel id (``rbx_sort_by_spaxel``): returns
    ``(order, sorted_ids, offsets)`` -- ``order`` is numpy's stable argsort of the ids with every id outside
    ``[0, num_segments)`` counted as ``num_segments`` (the particles segment_sum drops, rubix/spectra/ifu.py:286,
    sort to the end), ``offsets[s]`` the first sorted position with id >= s (``num_segments + 1`` entries)."""
    pixel = dev(pixel, torch.int32)
    if pixel.ndim != 1:
        raise ValueError(f"pixel must have shape (n,), got {tuple(pixel.shape)}")
    n, S2 = pixel.numel(), int(num_segments)
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    srt = torch.empty(n, dtype=torch.int32, device="cuda") if with_sorted else None
    off = torch.empty(S2 + 1, dtype=torch.int32, device="cuda") if with_offsets else None
    nbytes = int(_lib.lib().rbx_sort_by_spaxel_workspace_bytes(n, S2))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().rbx_sort_by_spaxel(_p(pixel), n, S2, _p(order), _p(srt), _p(off), _p(ws), nbytes, _stream()))
    return order, srt, off


def euler_rotation_matrix(alpha: float, beta: float, gamma: float) -> np.ndarray:
    """rubix/galaxy/alignment.py:164-209: R = R_z R_y R_x (degrees), float32."""
    a, b, g = np.deg2rad([alpha, beta, gamma])
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    return np.ascontiguousarray(Rz @ Ry @ Rx, dtype=np.float32)


def rotate_galaxy(coords, velocity, mass, halfmass_radius: float, alpha: float, beta: float, gamma: float,
                  comm=None):
    """rubix/galaxy/alignment.py:233-265 on the device: inertia tensor of the particles within the
    half-mass radius, eigenvector alignment, Euler rotation.  Returns (coords, velocity, R).

    ``comm`` (an ``ops.Comm``): the arrays are this rank's contiguous shard of ONE galaxy; the inertia sums are
    all-reduced so that every rank applies the rotation of the whole galaxy, as the reference does before it
    splits the particles over devices (rubix/core/rotation.py:76-115)."""
    coords, mass = dev(coords), dev(mass).reshape(-1)
    velocity = None if velocity is None else dev(velocity)
    n = coords.shape[0]
    if coords.ndim != 2 or coords.shape[1] != 3 or mass.numel() != n or (velocity is not None and velocity.shape != coords.shape):
        raise ValueError("rotate_galaxy: coords / velocity must be (n, 3) and mass (n,)")
    out_c = torch.empty_like(coords)
    out_v = None if velocity is None else torch.empty_like(velocity)
    R = torch.empty(9, dtype=torch.float32, device="cuda")
    L = _lib.lib()
    ws = _workspace(L.rbx_rotate_galaxy_workspace_bytes())
    E = euler_rotation_matrix(alpha, beta, gamma)
    radius = float(np.float32(halfmass_radius))
    if comm is None or comm.world == 1:
        _lib.check(L.rbx_rotate_galaxy(_p(coords), _p(velocity), _p(mass), n, radius, E.ctypes.data_as(C.c_void_p),
                                       _p(out_c), _p(out_v), _p(R), _p(ws), ws.numel(), _stream()))
    else:
        mom = torch.empty(12, dtype=torch.float64, device="cuda")
        _lib.check(L.rbx_rotate_moments(_p(coords), _p(mass), n, radius, 1 if comm.rank == 0 else 0, _p(mom), _p(ws),
                                        ws.numel(), _stream()))
        comm.allreduce_f64(mom)
        _lib.check(L.rbx_rotate_apply(_p(coords), _p(velocity), n, _p(mom), E.ctypes.data_as(C.c_void_p), _p(out_c),
                                      _p(out_v), _p(R), _stream()))
    return out_c, out_v, R.reshape(3, 3)


NOISE_DISTRIBUTIONS = {"normal": 0, "uniform": 1}


def apply_noise(cube, signal_to_noise: float, distribution: str = "normal", key=(0, 0)) -> torch.Tensor:
    """rubix/core/noise.py:63-78: cube + cube * N * S2N (``key`` = the two words of jax.random.PRNGKey(0))."""
    if distribution not in NOISE_DISTRIBUTIONS:
        raise ValueError(f"Invalid noise type: {distribution}. Supported types: {list(NOISE_DISTRIBUTIONS)}")
    cube = dev(cube)
    ny, nx, W = cube.shape
    out = torch.empty_like(cube)
    L = _lib.lib()
    ws = _workspace(L.rbx_apply_noise_workspace_bytes(ny, nx))
    _lib.check(L.rbx_apply_noise(_p(cube), _p(out), ny, nx, W, float(signal_to_noise), NOISE_DISTRIBUTIONS[distribution],
                                 int(key[0]), int(key[1]), _p(ws), ws.numel(), _stream()))
    return out


def noise_samples(n: int, distribution: str = "normal", key=(0, 0)):
    """The raw sample stream of apply_noise and its 32-bit words (tests)."""
    _require_cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    bits = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().rbx_noise_samples(_p(out), _p(bits), n, NOISE_DISTRIBUTIONS[distribution], int(key[0]),
                                            int(key[1]), _stream()))
    return out, bits


def dust_av(gas_coords, gas_pixel, gas_mass, gas_metals, star_coords, star_pixel, n_spaxels: int, dust_to_gas,
            ext_const: float, spaxel_area: float, return_cells: bool = False):
    """A_V of every star from the gas cells of its spaxel (rubix/spectra/dust/dust_extinction.py:240-337).
    ``dust_to_gas`` = (a_high, alpha_high, a_low, alpha_low, x_transition), see rubix_b200.dust."""
    gas_coords, star_coords = dev(gas_coords).reshape(-1, 3), dev(star_coords).reshape(-1, 3)
    gas_pixel, star_pixel = dev(gas_pixel, torch.int32).reshape(-1), dev(star_pixel, torch.int32).reshape(-1)
    gas_mass = dev(gas_mass).reshape(-1)
    ng, ns = gas_coords.shape[0], star_coords.shape[0]
    gas_metals = dev(gas_metals).reshape(ng, -1) if ng else dev(gas_metals).reshape(0, 5)
    if gas_pixel.numel() != ng or gas_mass.numel() != ng or star_pixel.numel() != ns:
        raise ValueError("dust_av: per-particle arrays of different lengths")
    av = torch.empty(ns, dtype=torch.float32, device="cuda")
    cells = torch.empty(ng, dtype=torch.float32, device="cuda") if return_cells else None
    L = _lib.lib()
    ws = _workspace(L.rbx_dust_av_workspace_bytes(ng, int(n_spaxels)))
    dtg = np.ascontiguousarray(np.asarray(dust_to_gas, dtype=np.float32))
    if dtg.shape != (5,):
        raise ValueError("dust_to_gas must hold (a_high, alpha_high, a_low, alpha_low, x_transition)")
    _lib.check(L.rbx_dust_av(_p(gas_coords), _p(gas_pixel), _p(gas_mass), _p(gas_metals), gas_metals.shape[1], ng,
                             _p(star_coords), _p(star_pixel), ns, int(n_spaxels), dtg.ctypes.data_as(C.c_void_p),
                             float(ext_const), float(spaxel_area), _p(av), _p(cells), _p(ws), ws.numel(), _stream()))
    return (av, cells) if return_cells else av


def apply_extinction(spectra, av, axav, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """spectra (n, W) * 10^(-0.4 * axav (W,) * av (n,))  (rubix/spectra/dust/dust_baseclasses.py:126-164)."""
    spectra, av, axav = dev(spectra), dev(av).reshape(-1), dev(axav).reshape(-1)
    W = spectra.shape[-1]
    n = spectra.numel() // W if W else 0
    if av.numel() != n or axav.numel() != W:
        raise ValueError("apply_extinction: av needs one entry per spectrum, axav one per wavelength")
    out = torch.empty_like(spectra) if out is None else out
    _lib.check(_lib.lib().rbx_apply_extinction(_p(spectra), _p(av), _p(axav), n, W, _p(out), _stream()))
    return out


def build_cube_dusty(plan: Plan, spectra, velocity, pixel, num_spaxels: int, av=None, axav=None) -> torch.Tensor:
    """doppler_shift_and_resampling -> calculate_extinction -> calculate_datacube in one kernel: mass-scaled SSP
    spectra (n, L) -> cube (S, S, W) without the (n, W) intermediate.  ``av`` / ``axav`` None = no dust."""
    spectra, velocity = dev(spectra), dev(velocity).reshape(-1, 3)
    pixel = dev(pixel, torch.int32).reshape(-1)
    n = velocity.shape[0]
    spectra = spectra.reshape(n, plan.L) if n else spectra.reshape(0, plan.L)
    if pixel.numel() != n:
        raise ValueError("build_cube_dusty: per-particle arrays of different lengths")
    if (av is None) != (axav is None):
        raise ValueError("build_cube_dusty: av and axav go together")
    if av is not None:
        av, axav = dev(av).reshape(-1), dev(axav).reshape(-1)
        if av.numel() != n or axav.numel() != plan.W:
            raise ValueError("build_cube_dusty: av needs one entry per particle, axav one per channel")
    cube = torch.empty((num_spaxels, num_spaxels, plan.W), dtype=torch.float32, device="cuda")
    L = _lib.lib()
    ws = _workspace(L.rbx_build_cube_dusty_workspace_bytes(n))
    _lib.check(L.rbx_build_cube_dusty(plan.handle, _p(spectra), _p(velocity), _p(pixel), _p(av), _p(axav), n,
                                      int(num_spaxels), _p(cube), _p(ws), ws.numel(), _stream()))
    return cube


#: the binned-moment dusty cube holds moments x spaxels x bins x W floats; beyond this many bytes (or for A_V ranges
#: that need more than DUSTY_MAX_BINS bins) callers use build_cube_dusty instead
DUSTY_MOMENT_BYTES = 24 << 30
DUSTY_MAX_BINS = 4096


def build_cube_dusty_binned(plan: Plan, velocity, mass, metallicity, age, pixel, num_spaxels: int, av, axav,
                            x_max: float = 0.1) -> Optional[torch.Tensor]:
    """calculate_spectra .. calculate_extinction .. calculate_datacube through the knot-based cube kernel: stars
    binned by A_V, the extinction factor expanded to third order inside a bin, one run of the cube kernel over four
    weighted copies of every star (include/rubix_b200.h, rbx_dusty_bins / rbx_dusty_combine).  Returns None when the A_V range or the cube size rule it out (non-finite
    A_V, too many bins, too much memory) -- the caller then takes build_cube_dusty."""
    velocity, mass = dev(velocity).reshape(-1, 3), dev(mass).reshape(-1)
    metallicity, age = dev(metallicity).reshape(-1), dev(age).reshape(-1)
    pixel, av, axav = dev(pixel, torch.int32).reshape(-1), dev(av).reshape(-1), dev(axav).reshape(-1)
    n, S = mass.numel(), int(num_spaxels)
    nseg = S * S
    if n == 0:
        return torch.zeros((S, S, plan.W), dtype=torch.float32, device="cuda")
    lo, hi = (float(v) for v in torch.aminmax(av))          # one device -> host read per cube
    kmax = float(axav.abs().max())
    if not (np.isfinite(lo) and np.isfinite(hi) and np.isfinite(kmax)):
        return None
    step = x_max / max(0.4 * np.log(10.0) * kmax, 1e-30)
    n_bins = max(1, int(np.ceil((hi - lo) / step)))
    L = _lib.lib()
    M = L.rbx_dusty_moments()
    Sv = int(np.ceil(np.sqrt(nseg * n_bins * M)))
    if n_bins > DUSTY_MAX_BINS or Sv * Sv * plan.W * 4 > DUSTY_MOMENT_BYTES:
        return None
    vpix = torch.empty(M * n, dtype=torch.int32, device="cuda")
    wmass = torch.empty(M * n, dtype=torch.float32, device="cuda")
    vvel = torch.empty((M * n, 3), dtype=torch.float32, device="cuda")
    vmet, vage = torch.empty_like(wmass), torch.empty_like(wmass)
    _lib.check(L.rbx_dusty_bins(_p(av), _p(pixel), _p(mass), _p(velocity), _p(metallicity), _p(age), n, nseg, n_bins, lo,
                                step, _p(vpix), _p(wmass), _p(vvel), _p(vmet), _p(vage), _stream()))
    vcube = build_cube(plan, vvel, wmass, vmet, vage, vpix, Sv)
    cube = torch.empty((S, S, plan.W), dtype=torch.float32, device="cuda")
    _lib.check(L.rbx_dusty_combine(_p(vcube), nseg, n_bins, lo, step, _p(axav), plan.W, _p(cube), _stream()))
    return cube


def ssp_lookup(plan: Plan, metallicity, age) -> torch.Tensor:
    metallicity, age = dev(metallicity).reshape(-1), dev(age).reshape(-1)
    n = metallicity.numel()
    if age.numel() != n:
        raise ValueError("metallicity and age must have the same length")
    out = torch.empty((n, plan.L), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().rbx_ssp_lookup(plan.handle, _p(metallicity), _p(age), n, _p(out), _stream()))
    return out


def scale_by_mass(spectra, mass) -> torch.Tensor:
    spectra, mass = dev(spectra), dev(mass).reshape(-1)
    L = spectra.shape[-1]
    n = spectra.numel() // L if L else 0
    if mass.numel() != n:
        raise ValueError("mass must have one entry per spectrum")
    out = torch.empty_like(spectra)
    _lib.check(_lib.lib().rbx_scale_by_mass(_p(spectra), _p(mass), n, L, _p(out), _stream()))
    return out


def doppler_resample(plan: Plan, spectra, velocity) -> torch.Tensor:
    spectra, velocity = dev(spectra), dev(velocity)
    n = velocity.shape[0]
    if spectra.shape != (n, plan.L) or velocity.shape != (n, 3):
        raise ValueError(f"expected spectra (n, {plan.L}) and velocity (n, 3)")
    out = torch.empty((n, plan.W), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().rbx_doppler_resample(plan.handle, _p(spectra), _p(velocity), n, _p(out), _stream()))
    return out


def segment_sum(spectra, pixel, num_segments: int, deterministic: bool = False) -> torch.Tensor:
    """rubix/spectra/ifu.py:286: ``segment_sum`` of the staged spectra.  The default adds with float atomics (any
    order, like the reference on a GPU backend); ``deterministic=True`` sorts the particles by spaxel
    (``rbx_sort_by_spaxel``) and adds every spaxel's run one particle after the other in particle order
    (``rbx_segment_sum_sorted``): bit-reproducible, and the float32 sum the reference forms on the CPU backend."""
    spectra, pixel = dev(spectra), dev(pixel, torch.int32).reshape(-1)
    n, W = spectra.shape
    if pixel.numel() != n:
        raise ValueError("pixel must have one entry per spectrum")
    if n == 0:                                   # an empty galaxy: nothing to launch
        return torch.zeros((num_segments, W), dtype=torch.float32, device="cuda")
    cube = torch.empty((num_segments, W), dtype=torch.float32, device="cuda")
    if deterministic:
        order, _, off = sort_by_spaxel(pixel, num_segments, with_sorted=False)
        _lib.check(_lib.lib().rbx_segment_sum_sorted(_p(spectra), _p(order), _p(off), W, num_segments, _p(cube),
                                                     _stream()))
        return cube
    _lib.check(_lib.lib().rbx_segment_sum(_p(spectra), _p(pixel), n, W, num_segments, _p(cube), 1, _stream()))
    return cube


_ws_cache = {}


def _workspace(nbytes: int) -> torch.Tensor:
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device="cuda")
        _ws_cache[key] = ws
    return ws


def build_cube(plan: Plan, velocity, mass, metallicity, age, pixel, num_spaxels: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused calculate_spectra -> scale_spectrum_by_mass -> doppler_shift_and_resampling ->
    calculate_datacube.  Returns the (S, S, W) cube."""
    velocity, mass = dev(velocity), dev(mass).reshape(-1)
    metallicity, age = dev(metallicity).reshape(-1), dev(age).reshape(-1)
    pixel = dev(pixel, torch.int32).reshape(-1)
    n = mass.numel()
    if velocity.shape != (n, 3) or metallicity.numel() != n or age.numel() != n or pixel.numel() != n:
        raise ValueError("particle arrays disagree in length")
    S = int(num_spaxels)
    cube = out if out is not None else torch.empty((S, S, plan.W), dtype=torch.float32, device="cuda")
    L = _lib.lib()
    nbytes = L.rbx_build_cube_workspace_bytes(plan.handle, n, S)
    ws = _workspace(nbytes)
    _lib.check(L.rbx_build_cube(plan.handle, _p(velocity), _p(mass), _p(metallicity), _p(age), _p(pixel),
                                n, S, _p(cube), _p(ws), ws.numel(), _stream()))
    return cube


def assign_build_cube(plan: Plan, coords, edges, velocity, mass, metallicity, age, num_spaxels: int,
                      apply_filter: bool = True, out: Optional[torch.Tensor] = None, return_pixel: bool = False):
    """filter_particles + spaxel_assignment + the fused cube build in one call (``rbx_assign_build_cube``): the
    spaxel index is computed inside the first kernel of the build, the particle arrays are read once."""
    coords, edges, velocity = dev(coords), dev(edges), dev(velocity)
    mass, metallicity, age = dev(mass).reshape(-1), dev(metallicity).reshape(-1), dev(age).reshape(-1)
    n = mass.numel()
    if coords.shape != (n, 3) or velocity.shape != (n, 3) or metallicity.numel() != n or age.numel() != n:
        raise ValueError("particle arrays disagree in length")
    S = int(num_spaxels)
    cube = out if out is not None else torch.empty((S, S, plan.W), dtype=torch.float32, device="cuda")
    pixel = torch.empty(n, dtype=torch.int32, device="cuda") if return_pixel else None
    L = _lib.lib()
    ws = _workspace(L.rbx_build_cube_workspace_bytes(plan.handle, n, S))
    _lib.check(L.rbx_assign_build_cube(plan.handle, _p(coords), _p(edges), edges.numel(), 1 if apply_filter else 0,
                                       _p(velocity), _p(mass), _p(metallicity), _p(age), n, S, _p(pixel), _p(cube),
                                       _p(ws), ws.numel(), _stream()))
    return (cube, pixel) if return_pixel else cube


def assign_build_cube_packed(plan: Plan, x, y, edges, vlos, mass, metallicity, age, num_spaxels: int,
                             apply_filter: bool = True, out: Optional[torch.Tensor] = None):
    """``rbx_assign_build_cube_packed``: the same build from structure-of-arrays particles (x, y and the
    line-of-sight velocity as arrays of their own: the 24 bytes per particle the path reads)."""
    x, y, edges, vlos = dev(x).reshape(-1), dev(y).reshape(-1), dev(edges), dev(vlos).reshape(-1)
    mass, metallicity, age = dev(mass).reshape(-1), dev(metallicity).reshape(-1), dev(age).reshape(-1)
    n = mass.numel()
    if any(t.numel() != n for t in (x, y, vlos, metallicity, age)):
        raise ValueError("particle arrays disagree in length")
    S = int(num_spaxels)
    cube = out if out is not None else torch.empty((S, S, plan.W), dtype=torch.float32, device="cuda")
    L = _lib.lib()
    ws = _workspace(L.rbx_build_cube_workspace_bytes(plan.handle, n, S))
    _lib.check(L.rbx_assign_build_cube_packed(plan.handle, _p(x), _p(y), _p(edges), edges.numel(),
                                              1 if apply_filter else 0, _p(vlos), _p(mass), _p(metallicity), _p(age),
                                              n, S, None, _p(cube), _p(ws), ws.numel(), _stream()))
    return cube


def build_cube_status(plan: Plan, n: int, num_spaxels: int):
    """(error, impl) of the last cube build of ``n`` particles on this stream's workspace: error 0 = ok (otherwise
    the cube was filled with NaN); impl 0 = the warp kernel ran, 1 = the general group kernel."""
    L = _lib.lib()
    ws = _workspace(L.rbx_build_cube_workspace_bytes(plan.handle, int(n), int(num_spaxels)))
    err, impl = C.c_int(), C.c_int()
    _lib.check(L.rbx_build_cube_status(plan.handle, int(n), int(num_spaxels), _p(ws), C.byref(err), C.byref(impl),
                                       _stream()))
    return int(err.value), int(impl.value)


def build_cube_cell_layout(plan: Plan, n: int, num_spaxels: int) -> bool:
    """True when the warp kernel of the last cube build of ``n`` particles on this stream's workspace kept its cells in
    the transposed block layout, False for cells in channel order / the group kernel."""
    L = _lib.lib()
    ws = _workspace(L.rbx_build_cube_workspace_bytes(plan.handle, int(n), int(num_spaxels)))
    tr = C.c_int()
    _lib.check(L.rbx_build_cube_cell_layout(plan.handle, int(n), int(num_spaxels), _p(ws), C.byref(tr), _stream()))
    return bool(tr.value)


def slab_geometry(W: int, nslab: int, halo: int = 12):
    """(wslab, ws): channels per wavelength slab and its stored width including the halo on both sides."""
    a, b = C.c_int(), C.c_int()
    _lib.check(_lib.lib().rbx_slab_geometry(int(W), int(nslab), int(halo), C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


def assign_build_cube_slabs(plan: Plan, coords, edges, velocity, mass, metallicity, age, num_spaxels: int, nslab: int,
                            halo: int = 12, apply_filter: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``rbx_assign_build_cube_slabs``: the partial cube stored slab-major, (nslab, S*S, wslab + 2 halo), every slab
    with the halo channels of its neighbours, ready for one reduce-scatter over the ranks."""
    coords, edges, velocity = dev(coords), dev(edges), dev(velocity)
    mass, metallicity, age = dev(mass).reshape(-1), dev(metallicity).reshape(-1), dev(age).reshape(-1)
    n = mass.numel()
    if coords.shape != (n, 3) or velocity.shape != (n, 3) or metallicity.numel() != n or age.numel() != n:
        raise ValueError("particle arrays disagree in length")
    S = int(num_spaxels)
    _, ws_ch = slab_geometry(plan.W, nslab, halo)
    slabs = out if out is not None else torch.empty((nslab, S * S, ws_ch), dtype=torch.float32, device="cuda")
    L = _lib.lib()
    ws = _workspace(L.rbx_build_cube_workspace_bytes(plan.handle, n, S))
    _lib.check(L.rbx_assign_build_cube_slabs(plan.handle, _p(coords), _p(edges), edges.numel(), 1 if apply_filter else 0,
                                             _p(velocity), _p(mass), _p(metallicity), _p(age), n, S, int(nslab),
                                             int(halo), _p(slabs), _p(ws), ws.numel(), _stream()))
    return slabs


def psf_lsf_own_slab(slab, num_spaxels: int, W: int, rank: int, nslab: int, psf_kernel, lsf_kernel, halo: int = 12,
                     ext: int = 12) -> torch.Tensor:
    """PSF + LSF of one summed slab (S*S, wslab + 2 halo) as rbx_reduce_scatter_cube leaves it on rank ``rank``;
    returns the (S, S, n_own) interior, n_own = the slab's channels inside [0, W)."""
    S = int(num_spaxels)
    wslab, ws_ch = slab_geometry(W, nslab, halo)
    slab = dev(slab).reshape(S, S, ws_ch)
    hp, hl = _host_taps(psf_kernel), _host_taps(lsf_kernel)
    if hp is None or hl is None:
        raise ValueError("psf_lsf_own_slab needs host (numpy) kernels")
    out = torch.empty_like(slab)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = _lib.lib().rbx_psf_lsf_taps_pitched(_p(slab), ws_ch, _p(out), ws_ch, S, S, ws_ch, vp(hp), hp.shape[0],
                                             hp.shape[1], vp(hl.reshape(-1)), hl.size, ext, _stream())
    if rc == _lib.RBX_ERR_UNSUPPORTED:
        out = psf_lsf(slab, psf_kernel, lsf_kernel, ext, host_taps=False)
    else:
        _lib.check(rc)
    n_own = max(0, min(wslab, W - rank * wslab))
    return out[:, :, halo:halo + n_own]


class Comm:
    """The exchange step behind the C ABI (``rbx_comm_*``, NCCL bound at run time).  ``unique_id()`` on one rank,
    the 128 bytes handed to every rank by the host, then ``Comm(id, rank, world)`` on every rank."""

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.lib().rbx_comm_unique_id(buf))
        return buf.raw

    def __init__(self, uid: bytes, rank: int, world: int):
        _require_cuda()
        self._h = C.c_void_p()
        buf = C.create_string_buffer(bytes(uid), 128)
        _lib.check(_lib.lib().rbx_comm_init(C.byref(self._h), buf, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    @classmethod
    def from_torch_distributed(cls):
        """Bootstrap over an initialised torch.distributed group (any backend): rank 0's id is broadcast."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        obj = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        return cls(obj[0], rank, world)

    def nccl_version(self) -> int:
        v = C.c_int()
        _lib.check(_lib.lib().rbx_comm_info(self._h, None, None, C.byref(v)))
        return int(v.value)

    def reduce(self, send: torch.Tensor, recv: Optional[torch.Tensor] = None, root: int = 0):
        recv = send if recv is None else recv
        _lib.check(_lib.lib().rbx_reduce_cube(self._h, _p(send), _p(recv), send.numel(), int(root), _stream()))
        return recv

    def allreduce(self, send: torch.Tensor, recv: Optional[torch.Tensor] = None):
        recv = send if recv is None else recv
        _lib.check(_lib.lib().rbx_allreduce_cube(self._h, _p(send), _p(recv), send.numel(), _stream()))
        return recv

    def allreduce_f64(self, buf: torch.Tensor):
        _lib.check(_lib.lib().rbx_allreduce_f64(self._h, _p(buf), _p(buf), buf.numel(), _stream()))
        return buf

    def reduce_scatter(self, send: torch.Tensor, recv: torch.Tensor):
        if send.numel() != recv.numel() * self.world:
            raise ValueError("reduce_scatter: send must hold world * recv elements")
        _lib.check(_lib.lib().rbx_reduce_scatter_cube(self._h, _p(send), _p(recv), recv.numel(), _stream()))
        return recv

    def allgather(self, send: torch.Tensor, recv: torch.Tensor):
        if recv.numel() != send.numel() * self.world:
            raise ValueError("allgather: recv must hold world * send elements")
        _lib.check(_lib.lib().rbx_allgather_cube(self._h, _p(send), _p(recv), send.numel(), _stream()))
        return recv

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().rbx_comm_destroy(self._h)
            self._h = C.c_void_p()


def _host_taps(k):
    """float32 host copy of a kernel given as numpy / list (None for CUDA tensors: those stay on the device)."""
    if k is None or (isinstance(k, torch.Tensor) and k.is_cuda):
        return None
    if isinstance(k, torch.Tensor):
        k = k.numpy()
    return np.ascontiguousarray(np.asarray(k), dtype=np.float32)


def _taps_call(cube, out, pk, lk, ext) -> bool:
    """rbx_psf_lsf_taps (host-known taps: separable PSF, pruned LSF).  False when the taps need the
    general device-tap kernels."""
    ny, nx, W = cube.shape
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    M, N = pk.shape if pk is not None else (0, 0)
    K = lk.size if lk is not None else 0
    rc = _lib.lib().rbx_psf_lsf_taps(_p(cube), _p(out), ny, nx, W, vp(pk), M, N, vp(lk), K, ext, _stream())
    if rc == _lib.RBX_ERR_UNSUPPORTED:
        return False
    _lib.check(rc)
    return True


def psf_lsf_slab(cube, lo: int, hi: int, psf_kernel, lsf_kernel, ext: int = 12) -> torch.Tensor:
    """PSF + LSF of the wavelength slab [lo, hi) of a full (ny, nx, W) cube: the multi-GPU PSF / LSF stage
    shards by wavelength.  The slab is read in place with a halo of ``ext`` channels on each
    side (strided view, no copy) and the (ny, nx, hi - lo) interior is returned."""
    cube = dev(cube)
    ny, nx, W = cube.shape
    hp, hl = _host_taps(psf_kernel), _host_taps(lsf_kernel)
    if hp is None or hl is None:
        raise ValueError("psf_lsf_slab needs host (numpy) kernels")
    hlo, hhi = max(lo - ext, 0), min(hi + ext, W)
    ws = hhi - hlo
    out = torch.empty((ny, nx, ws), dtype=torch.float32, device="cuda")
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    src = C.c_void_p(cube.data_ptr() + 4 * hlo)
    rc = _lib.lib().rbx_psf_lsf_taps_pitched(src, W, _p(out), ws, ny, nx, ws, vp(hp), hp.shape[0], hp.shape[1],
                                             vp(hl.reshape(-1)), hl.size, ext, _stream())
    if rc == _lib.RBX_ERR_UNSUPPORTED:  # general taps: device-tap kernels on a contiguous copy of the slab
        out = psf_lsf(cube[:, :, hlo:hhi].contiguous(), psf_kernel, lsf_kernel, ext, host_taps=False)
    else:
        _lib.check(rc)
    return out[:, :, lo - hlo:lo - hlo + (hi - lo)]


def convolve_psf(cube, kernel, host_taps: bool = True) -> torch.Tensor:
    """a6.  A kernel given on the host (numpy) takes the host-tap kernel when it is an outer product."""
    cube = dev(cube)
    ny, nx, W = cube.shape
    out = torch.empty_like(cube)
    hk = _host_taps(kernel) if host_taps else None
    if hk is not None and hk.ndim == 2 and _taps_call(cube, out, hk, None, 0):
        return out
    kernel = dev(kernel)
    _lib.check(_lib.lib().rbx_convolve_psf(_p(cube), _p(out), ny, nx, W, _p(kernel), kernel.shape[0],
                                           kernel.shape[1], _stream()))
    return out


def convolve_lsf(cube, kernel, ext: int = 12, host_taps: bool = True) -> torch.Tensor:
    """a7.  A kernel given on the host (numpy) takes the host-tap kernel (negligible taps dropped)."""
    cube = dev(cube)
    W = cube.shape[-1]
    rows = cube.numel() // W
    out = torch.empty_like(cube)
    hk = _host_taps(kernel) if host_taps else None
    if hk is not None and cube.ndim == 3 and hk.size == 2 * ext + 1 and _taps_call(cube, out, None, hk.reshape(-1), ext):
        return out
    kernel = dev(kernel).reshape(-1)
    _lib.check(_lib.lib().rbx_convolve_lsf(_p(cube), _p(out), rows, W, _p(kernel), kernel.numel(), ext, _stream()))
    return out


def psf_lsf(cube, psf_kernel, lsf_kernel, ext: int = 12, host_taps: bool = True) -> torch.Tensor:
    """a6 + a7 in one pass.  Kernels given on the host (numpy, as the reference builds them) take
    rbx_psf_lsf_taps; CUDA tensors, or taps it does not cover, take the device-tap kernels."""
    cube = dev(cube)
    ny, nx, W = cube.shape
    out = torch.empty_like(cube)
    hp, hl = (_host_taps(psf_kernel), _host_taps(lsf_kernel)) if host_taps else (None, None)
    if hp is not None and hl is not None and hp.ndim == 2 and _taps_call(cube, out, hp, hl.reshape(-1), ext):
        return out
    pk, lk = dev(psf_kernel), dev(lsf_kernel).reshape(-1)
    rc = _lib.lib().rbx_psf_lsf(_p(cube), _p(out), ny, nx, W, _p(pk), pk.shape[0], pk.shape[1], _p(lk),
                                lk.numel(), ext, _stream())
    if rc == _lib.RBX_ERR_UNSUPPORTED:  # taps too large for the fused tile: two CUDA passes
        return convolve_lsf(convolve_psf(cube, pk), lk, ext)
    _lib.check(rc)
    return out


def pipeline_host_packed(plan: Plan, x, y, vlos, mass, metallicity, age, edges, num_spaxels: int,
                         psf_kernel=None, lsf_kernel=None, ext: int = 12, apply_filter: bool = True,
                         out: Optional[np.ndarray] = None) -> np.ndarray:
    """``rbx_pipeline_host_packed``: the host-buffer call with structure-of-arrays particles (24 B each over PCIe)."""
    f32 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    x, y, vlos, mass = f32(x), f32(y), f32(vlos), f32(mass)
    metallicity, age, edges = f32(metallicity), f32(age), f32(edges)
    n = x.shape[0]
    S = int(num_spaxels)
    cube = out if out is not None else np.empty((S, S, plan.W), dtype=np.float32)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    pk = None if psf_kernel is None else f32(psf_kernel)
    lk = None if lsf_kernel is None else f32(lsf_kernel)
    M, N = (pk.shape if pk is not None else (0, 0))
    K = len(lk) if lk is not None else 0
    _lib.check(_lib.lib().rbx_pipeline_host_packed(
        plan.handle, vp(x), vp(y), vp(vlos), vp(mass), vp(metallicity), vp(age), n, vp(edges), len(edges), S,
        1 if apply_filter else 0, vp(pk), M, N, vp(lk), K, ext, vp(cube), _stream()))
    return cube


def build_cube_host(plan: Plan, coords, velocity, mass, metallicity, age, edges, num_spaxels: int,
                    out: torch.Tensor, nslab: int = 1, halo: int = 12, apply_filter: bool = True) -> torch.Tensor:
    """``rbx_build_cube_host``: a rank's HOST shard (numpy arrays, pinned for the copy / compute overlap) -> its partial
    cube in the device tensor ``out`` ((S, S, W), or the slab-major (nslab, S*S, ws) block for ``nslab > 1``).  Stream
    ordered: the host arrays must stay alive until the stream has passed the call (the exchange and the PSF + LSF that
    follow on the same stream are fine)."""
    req = lambda a: a if (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous) else \
        np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    coords, velocity, mass = req(coords), req(velocity), req(mass)
    metallicity, age, edges = req(metallicity), req(age), req(edges)
    _require_cuda()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib().rbx_build_cube_host(
        plan.handle, vp(coords), vp(velocity), vp(mass), vp(metallicity), vp(age), coords.shape[0], vp(edges),
        len(edges), int(num_spaxels), 1 if apply_filter else 0, int(nslab), int(halo), _p(out), _stream()))
    return out


def pipeline_host(plan: Plan, coords, velocity, mass, metallicity, age, edges, num_spaxels: int,
                  psf_kernel=None, lsf_kernel=None, ext: int = 12, apply_filter: bool = True,
                  out: Optional[np.ndarray] = None) -> np.ndarray:
    """Host-buffer entry point (``rbx_pipeline_host``): numpy in, numpy cube out, copies included."""
    f32 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    coords, velocity, mass = f32(coords), f32(velocity), f32(mass)
    metallicity, age, edges = f32(metallicity), f32(age), f32(edges)
    n = coords.shape[0]
    S = int(num_spaxels)
    cube = out if out is not None else np.empty((S, S, plan.W), dtype=np.float32)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    pk = None if psf_kernel is None else f32(psf_kernel)
    lk = None if lsf_kernel is None else f32(lsf_kernel)
    M, N = (pk.shape if pk is not None else (0, 0))
    K = len(lk) if lk is not None else 0
    _lib.check(_lib.lib().rbx_pipeline_host(
        plan.handle, vp(coords), vp(velocity), vp(mass), vp(metallicity), vp(age), n, vp(edges), len(edges), S,
        1 if apply_filter else 0, vp(pk), M, N, vp(lk), K, ext, vp(cube), _stream()))
    return cube
