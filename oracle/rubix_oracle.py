"""CPU oracle for the rubix particle -> IFU datacube path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, not the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``rubix_b200/`` imports it.

It is a plain-numpy restatement of the reference algorithm (AstroAI-Lab/rubix @ dbb4487, mounted
at /root/reference while this was written; paths below are relative to it).  Every function keeps
the reference's *dataflow* (materialised ``(P, L)`` and ``(P, W)`` intermediates) and, with
``dtype=np.float32``, the reference's operation order, so that it can stand in for the JAX-CPU
pipeline, which cannot be imported in this environment (no jax / interpax / equinox / h5py).
With ``dtype=np.float64`` the same formulas are evaluated in double precision on the same float32
inputs; that is the "truth" the float32 CUDA path and the float32 oracle are both compared with.

PARITY PINS
-----------
* Checked against every known-answer test the reference holds for this path
  (tests/test_oracle_golden.py lists them with file:line).
* Checked against OUTPUTS OF THE REFERENCE'S OWN SOURCE run here: tools/refshim.py executes the reference files for a0,
  a2 - a7, rotate_galaxy, calculate_S2N, the dust modules, the cosmology and the rubix.core factories of the path
  unchanged with numpy standing in for jax.numpy (float64); the vectors are committed as
  tests/golden/ref_numpy_{stages,cube,dust}.npz with their generator tools/make_ref_golden.py, and
  tests/test_oracle_vs_reference_source.py holds this file's float64 mode (and the C form) to them at 1e-11
  (integer results exactly).  That pins the LOGIC of everything but a1 to the reference's code; it is not jax
  arithmetic (numpy's rounding order, numpy's double-precision ``interp``).  tools/fuzz_oracle_vs_reference.py adds
  random and degenerate inputs (600 cases per function agree) and checks the OPERATION ORDER of the float32 mode: with
  numpy float32 arithmetic on both sides (``jnp.interp`` by jax's formula in the input dtype) the Doppler shifts,
  ``resample_spectrum`` and ``calculate_cube`` agree with the reference's source bit for bit.
* ``interp2d`` is NOT in the reference tree: it is ``interpax.interp2d`` (PyPI ``interpax``,
  unpinned in the reference's pyproject.toml:38, no lock file).  Its published algorithm is restated
  in :func:`interp2d` below.  The reference's tests pin it only at grid nodes and out-of-grid
  (tests/test_core_ssp.py:95-181, tests/test_ssp_grid.py:653-698).  **Off-node interpolated values:
  parity unpinned against interpax** (no golden vector exists in the reference and interpax cannot be run here);
  tests/test_interp2d_scipy.py pins them to two independent scipy implementations of the published algorithm.
"""

from __future__ import annotations

import numpy as np

SPEED_OF_LIGHT = 299792.458  # rubix/config/rubix_config.yml:8  (km/s)


# --------------------------------------------------------------------------------------
# a0  spaxel assignment + aperture mask
# --------------------------------------------------------------------------------------
def square_spaxel_assignment(coords, spatial_bin_edges):
    """rubix/telescope/utils.py:138-151.

    ``jnp.digitize(x, edges)`` is ``searchsorted(edges, x, side='right')`` for increasing edges.
    Returns int32 flat indices ``x + nbins * y``.
    """
    coords = np.asarray(coords)
    edges = np.asarray(spatial_bin_edges)
    xi = np.searchsorted(edges, coords[:, 0], side="right") - 1
    yi = np.searchsorted(edges, coords[:, 1], side="right") - 1
    nb = len(edges) - 1
    xi = np.clip(xi, 0, nb - 1)
    yi = np.clip(yi, 0, nb - 1)
    return (xi + nb * yi).astype(np.int32)


def mask_particles_outside_aperture(coords, spatial_bin_edges):
    """rubix/telescope/utils.py:170-174 (inclusive on both boundaries)."""
    coords = np.asarray(coords)
    edges = np.asarray(spatial_bin_edges)
    lo, hi = edges.min(), edges.max()
    m = (coords[:, 0] >= lo) & (coords[:, 0] <= hi)
    m &= (coords[:, 1] >= lo) & (coords[:, 1] <= hi)
    return m


def filter_particles(coords, mass, metallicity, age, spatial_bin_edges):
    """rubix/core/telescope.py:155-174: masked particles get mass = metallicity = age = 0
    (coords and velocity are left alone)."""
    m = mask_particles_outside_aperture(coords, spatial_bin_edges)
    z = lambda a: np.where(m, a, np.zeros((), dtype=np.asarray(a).dtype))
    return z(mass), z(metallicity), z(age), m


def reshape_array(arr, n_dev):
    """rubix/core/data.py:461-487: pad with zeros to a multiple of n_dev, add leading device axis."""
    arr = np.asarray(arr)
    n = arr.shape[0]
    per = (n + n_dev - 1) // n_dev
    pad = per * n_dev - n
    if pad:
        arr = np.concatenate([arr, np.zeros((pad,) + arr.shape[1:], arr.dtype)], axis=0)
    return arr.reshape((n_dev, per) + arr.shape[1:])


# --------------------------------------------------------------------------------------
# a1  SSP lookup: interpax.interp2d(xq=Z, yq=age, x=Zgrid, y=agegrid, f=flux, method, extrap=0)
# --------------------------------------------------------------------------------------
# Inverse of the bicubic-patch matrix (coefficients a_ij of sum a_ij x^i y^j from
# [f, fx, fy, fxy] at the four corners, corner order (0,0),(1,0),(0,1),(1,1)); interpax A_BICUBIC.
A_BICUBIC = np.array(
    [
        [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [-3, 3, 0, 0, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [2, -2, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, -2, -1, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 2, -2, 0, 0, 1, 1, 0, 0],
        [-3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0],
        [9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1],
        [-6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1],
        [2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0],
        [-6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1],
        [4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1],
    ],
    dtype=np.float64,
)


def approx_df(x, f, axis):
    """interpax ``approx_df(x, f, method="cubic", axis)``: node derivatives for the C1 cubic
    spline: one-sided secant at both ends, plain mean of the two adjacent secants inside."""
    x = np.asarray(x)
    f = np.asarray(f)
    dx = np.diff(x)
    df = np.diff(f, axis=axis)
    with np.errstate(divide="ignore"):
        dxi = np.where(dx == 0, 0, 1 / np.where(dx == 0, 1, dx)).astype(f.dtype)
    shape = [1] * f.ndim
    shape[axis] = -1
    df = dxi.reshape(shape) * df
    first = np.take(df, [0], axis=axis)
    last = np.take(df, [-1], axis=axis)
    n = df.shape[axis]
    mid = f.dtype.type(0.5) * (
        np.take(df, np.arange(0, n - 1), axis=axis) + np.take(df, np.arange(1, n), axis=axis)
    )
    return np.concatenate([first, mid, last], axis=axis)


def interp2d(xq, yq, x, y, f, method="cubic", extrap=0, dtype=np.float32, hermite=False):
    """Restatement of ``interpax.interp2d`` as bound by rubix/spectra/ssp/grid.py:113-120
    (``x``=metallicity grid, ``y``=age grid, ``f``=flux ``(nx, ny, L)``, ``extrap=0``) and called at
    rubix/core/ifu.py:107-110 with 1-D ``xq``, ``yq`` of equal length.  Returns ``(P, L)``.

    ``hermite=True`` evaluates the same bicubic patch through the Hermite basis instead of the
    16x16 coefficient matrix (mathematically identical; this is the form the CUDA kernel uses).
    """
    dt = np.dtype(dtype).type
    xq = np.atleast_1d(np.asarray(xq)).astype(dtype)
    yq = np.atleast_1d(np.asarray(yq)).astype(dtype)
    x = np.asarray(x).astype(dtype)
    y = np.asarray(y).astype(dtype)
    f = np.asarray(f).astype(dtype)
    i = np.clip(np.searchsorted(x, xq, side="right"), 1, len(x) - 1)
    j = np.clip(np.searchsorted(y, yq, side="right"), 1, len(y) - 1)
    x0, x1 = x[i - 1], x[i]
    y0, y1 = y[j - 1], y[j]
    dx = x1 - x0
    dy = y1 - y0
    with np.errstate(divide="ignore", invalid="ignore"):
        dxi = np.where(dx == 0, dt(0), dt(1) / dx).astype(dtype)
        dyi = np.where(dy == 0, dt(0), dt(1) / dy).astype(dtype)

    if method == "linear":
        f00 = f[i - 1, j - 1]
        f01 = f[i - 1, j]
        f10 = f[i, j - 1]
        f11 = f[i, j]
        tx = np.stack([x1 - xq, xq - x0])  # (2, P)
        ty = np.stack([y1 - yq, yq - y0])
        F = np.stack([np.stack([f00, f01]), np.stack([f10, f11])])  # (2,2,P,L)
        acc = np.einsum("ijkl,ik,jk->kl", F, tx, ty).astype(dtype)
        fq = ((dxi * dyi)[:, None] * acc).astype(dtype)
    elif method == "cubic":
        fx = approx_df(x, f, 0)
        fy = approx_df(y, f, 1)
        fxy = approx_df(y, fx, 1)
        tx = (xq - x0) * dxi
        ty = (yq - y0) * dyi
        vals = []
        for name, tab in (("f", f), ("fx", fx), ("fy", fy), ("fxy", fxy)):
            for jj in (0, 1):
                for ii in (0, 1):
                    v = tab[i - 1 + ii, j - 1 + jj]  # (P, L)
                    if "x" in name:
                        v = dx[:, None] * v
                    if "y" in name:
                        v = dy[:, None] * v
                    vals.append(v.astype(dtype))
        F = np.stack(vals, axis=0)  # (16, P, L)
        if hermite:
            def hb(t):
                t2 = t * t
                t3 = t2 * t
                return (
                    (dt(2) * t3 - dt(3) * t2 + dt(1)),  # value at 0
                    (dt(-2) * t3 + dt(3) * t2),  # value at 1
                    (t3 - dt(2) * t2 + t),  # slope at 0
                    (t3 - t2),  # slope at 1
                )
            hx = hb(tx)
            hy = hb(ty)
            fq = np.zeros(F.shape[1:], dtype=dtype)
            k = 0
            for (sx, sy) in ((0, 0), (2, 0), (0, 2), (2, 2)):  # f, fx, fy, fxy
                for jj in (0, 1):
                    for ii in (0, 1):
                        w = hx[sx + ii] * hy[sy + jj]
                        fq = fq + w[:, None] * F[k]
                        k += 1
            fq = fq.astype(dtype)
        else:
            coef = np.einsum("ab,bpl->apl", A_BICUBIC.astype(dtype), F).astype(dtype)  # (16,P,L)
            coef = coef.reshape((4, 4) + coef.shape[1:], order="F")  # [i_pow_x, j_pow_y, P, L]
            ttx = np.stack([np.ones_like(tx), tx, tx * tx, tx * tx * tx])  # (4, P)
            tty = np.stack([np.ones_like(ty), ty, ty * ty, ty * ty * ty])
            fq = np.einsum("jkil,ji,ki->il", coef, ttx, tty).astype(dtype)
    else:
        raise ValueError(f"unknown method {method}")

    # extrap=0 (an int, not a bool): values outside [x0, x_last] are replaced by 0, boundaries
    # inclusive (interpax _extrap: where(xq < x[0], lo, fq); where(xq > x[-1], hi, fq)).
    lo = dt(extrap)
    fq = np.where((xq < x[0])[:, None], lo, fq)
    fq = np.where((xq > x[-1])[:, None], lo, fq)
    fq = np.where((yq < y[0])[:, None], lo, fq)
    fq = np.where((yq > y[-1])[:, None], lo, fq)
    return fq.astype(dtype)


def calculate_spectra(metallicity, age, ssp_metallicity, ssp_age, ssp_flux, method="cubic",
                      dtype=np.float32, chunk_size=250000):
    """rubix/core/ifu.py:95-118 (250k-particle chunks, concatenated).  1-D inputs (device shard 0)."""
    out = []
    n = len(metallicity)
    for s in range(0, n, chunk_size):
        e = min(s + chunk_size, n)
        out.append(interp2d(metallicity[s:e], age[s:e], ssp_metallicity, ssp_age, ssp_flux,
                            method=method, dtype=dtype))
    if not out:
        return np.zeros((0, np.asarray(ssp_flux).shape[-1]), dtype=dtype)
    return np.concatenate(out, axis=0)


# --------------------------------------------------------------------------------------
# a2  mass scaling
# --------------------------------------------------------------------------------------
def scale_spectrum_by_mass(spectra, mass):
    """rubix/core/ifu.py:152-154: ``spectra * mass[..., None]`` (a single multiply)."""
    return spectra * np.asarray(mass, dtype=spectra.dtype)[..., None]


# --------------------------------------------------------------------------------------
# a3  Doppler shift
# --------------------------------------------------------------------------------------
def cosmological_doppler_shift(z, wavelength, dtype=np.float32):
    """rubix/spectra/ifu.py:80: ``(1 + z) * wavelength`` (python float times f32 array -> f32)."""
    return (np.dtype(dtype).type(1 + z) * np.asarray(wavelength, dtype=dtype)).astype(dtype)


def doppler_factor(velocity_component, dtype=np.float32, c=SPEED_OF_LIGHT):
    """rubix/spectra/ifu.py:190: ``exp(v / c)`` per particle."""
    v = np.asarray(velocity_component, dtype=dtype)
    return np.exp(v / np.dtype(dtype).type(c)).astype(dtype)


def velocity_doppler_shift(wavelength, velocity, direction="z", dtype=np.float32):
    """rubix/spectra/ifu.py:216-220: (P, L) shifted wavelengths."""
    comp = {"x": 0, "y": 1, "z": 2}[direction]
    d = doppler_factor(np.asarray(velocity)[:, comp], dtype=dtype)
    return (np.asarray(wavelength, dtype=dtype)[None, :] * d[:, None]).astype(dtype)


# --------------------------------------------------------------------------------------
# a4  flux-conserving resample
# --------------------------------------------------------------------------------------
def calculate_diff(vec):
    """rubix/spectra/ifu.py:84-102: ``jnp.diff(vec, prepend=vec[0])`` -> [0, v1-v0, ...]."""
    vec = np.asarray(vec)
    return np.diff(vec, prepend=vec[..., :1], axis=-1)


def jnp_interp(x, xp, fp):
    """``jax.numpy.interp(x, xp, fp)`` (non-periodic, left/right = end values), 1-D ``xp``."""
    x = np.asarray(x)
    xp = np.asarray(xp)
    fp = np.asarray(fp)
    i = np.clip(np.searchsorted(xp, x, side="right"), 1, len(xp) - 1)
    df = fp[i] - fp[i - 1]
    dx = xp[i] - xp[i - 1]
    delta = x - xp[i - 1]
    eps = np.spacing(np.finfo(xp.dtype).eps)
    dx0 = np.abs(dx) <= eps
    with np.errstate(invalid="ignore", divide="ignore"):
        f = np.where(dx0, fp[i - 1], fp[i - 1] + (delta / np.where(dx0, 1, dx)) * df)
    f = np.where(x < xp[0], fp[0], f)
    f = np.where(x > xp[-1], fp[-1], f)
    return f.astype(fp.dtype)


def resample_spectrum(initial_spectrum, initial_wavelength, target_wavelength):
    """rubix/spectra/ifu.py:241-260 for one particle (dtype follows the inputs)."""
    s = np.asarray(initial_spectrum)
    lam = np.asarray(initial_wavelength)
    t = np.asarray(target_wavelength)
    dt = s.dtype.type
    in_range = (lam >= t.min()) & (lam <= t.max())
    wave_diff = calculate_diff(lam) * in_range
    total_lum = np.sum(s * wave_diff, dtype=s.dtype)
    p = jnp_interp(t, lam, s)
    new_total = np.sum(p * calculate_diff(t), dtype=s.dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        scale = total_lum / new_total
    scale = np.nan_to_num(scale, nan=0.0)  # +-inf -> +-max finite, like jnp.nan_to_num
    return (p * dt(scale)).astype(s.dtype)


def resample_spectra(spectra, wavelengths, target_wavelength):
    """vmapped :func:`resample_spectrum` (rubix/core/ifu.py:164-183)."""
    spectra = np.asarray(spectra)
    out = np.empty((spectra.shape[0], len(target_wavelength)), dtype=spectra.dtype)
    for k in range(spectra.shape[0]):
        out[k] = resample_spectrum(spectra[k], wavelengths[k], target_wavelength)
    return out


# --------------------------------------------------------------------------------------
# a5  cube
# --------------------------------------------------------------------------------------
def calculate_cube(spectra, spaxel_index, num_spaxels, acc_dtype=None):
    """rubix/spectra/ifu.py:286-287: ``segment_sum`` (indices outside [0, S^2) are dropped, as XLA
    scatter does) then C-order reshape to (S, S, W) => cube[y, x, w].  Adds run in particle order."""
    spectra = np.asarray(spectra)
    idx = np.asarray(spaxel_index)
    acc_dtype = acc_dtype or spectra.dtype
    cube = np.zeros((num_spaxels * num_spaxels, spectra.shape[-1]), dtype=acc_dtype)
    ok = (idx >= 0) & (idx < num_spaxels * num_spaxels)
    np.add.at(cube, idx[ok], spectra[ok].astype(acc_dtype))
    return cube.reshape(num_spaxels, num_spaxels, spectra.shape[-1])


# --------------------------------------------------------------------------------------
# a6  PSF
# --------------------------------------------------------------------------------------
def gaussian_kernel_2d(m, n, sigma, dtype=np.float32):
    """rubix/telescope/psf/kernels.py:26-31."""
    x = np.arange(-((m - 1) / 2), ((m - 1) / 2) + 1).astype(dtype)
    y = np.arange(-((n - 1) / 2), ((n - 1) / 2) + 1).astype(dtype)
    X, Y = np.meshgrid(x, y, indexing="ij")
    dt = np.dtype(dtype).type
    values = np.exp(-(X**2 + Y**2) / dt(2 * sigma**2)).astype(dtype)
    return (values / np.sum(values, dtype=dtype)).astype(dtype)


def convolve2d_same(plane, kernel):
    """``jax.scipy.signal.convolve2d(plane, kernel, mode="same")`` for plane >= kernel in both
    dims: zero-padded true convolution, out[i,j] = sum_mn K[m,n] plane[i-m+(M-1)//2, j-n+(N-1)//2]."""
    plane = np.asarray(plane)
    kernel = np.asarray(kernel)
    M, N = kernel.shape
    H, Wd = plane.shape
    cm, cn = (M - 1) // 2, (N - 1) // 2
    pad = np.zeros((H + M - 1, Wd + N - 1), dtype=plane.dtype)
    pad[M - 1 - cm:M - 1 - cm + H, N - 1 - cn:N - 1 - cn + Wd] = plane
    out = np.zeros_like(plane)
    for m in range(M):
        for n in range(N):
            # plane[i - m + cm] == pad[i - m + cm + (M-1-cm)] = pad[i + (M-1-m)]
            out = out + kernel[m, n] * pad[M - 1 - m:M - 1 - m + H, N - 1 - n:N - 1 - n + Wd]
    return out.astype(plane.dtype)


def apply_psf(datacube, psf_kernel):
    """rubix/telescope/psf/psf.py:56-57: convolve every wavelength slice."""
    datacube = np.asarray(datacube)
    kernel = np.asarray(psf_kernel).astype(datacube.dtype)
    M, N = kernel.shape
    H, Wd, L = datacube.shape
    cm, cn = (M - 1) // 2, (N - 1) // 2
    pad = np.zeros((H + M - 1, Wd + N - 1, L), dtype=datacube.dtype)
    pad[M - 1 - cm:M - 1 - cm + H, N - 1 - cn:N - 1 - cn + Wd] = datacube
    out = np.zeros_like(datacube)
    for m in range(M):
        for n in range(N):
            out = out + kernel[m, n] * pad[M - 1 - m:M - 1 - m + H, N - 1 - n:N - 1 - n + Wd]
    return out.astype(datacube.dtype)


# --------------------------------------------------------------------------------------
# a7  LSF
# --------------------------------------------------------------------------------------
def lsf_kernel(sigma, wave_res, factor=12, dtype=np.float32):
    """rubix/telescope/lsf/lsf.py:12-26."""
    x = np.arange(-factor * wave_res, factor * wave_res + wave_res, wave_res).astype(dtype)
    dt = np.dtype(dtype).type
    res = np.exp(dt(-0.5) * (x**2) / dt(sigma**2)).astype(dtype)
    return (res / np.sum(res, dtype=dtype)).astype(dtype)


def apply_lsf(datacube, lsf_sigma, wave_resolution, extend_factor=12):
    """rubix/telescope/lsf/lsf.py:59-65,96-105: full convolution along the last axis, then the
    slice ``[extend_factor : W + K - 1 - extend_factor]``."""
    datacube = np.asarray(datacube)
    shape = datacube.shape
    flat = datacube.reshape(-1, shape[-1])
    k = lsf_kernel(lsf_sigma, wave_resolution, extend_factor, dtype=datacube.dtype)
    K = len(k)
    Wn = shape[-1]
    full = np.zeros((flat.shape[0], Wn + K - 1), dtype=datacube.dtype)
    for m in range(K):
        full[:, m:m + Wn] += k[m] * flat
    end = Wn + K - 1 - extend_factor
    return full[:, extend_factor:end].reshape(shape[:-1] + (end - extend_factor,))


# --------------------------------------------------------------------------------------
# f4  dust extinction (calc_dusty_ifu): rubix/core/dust.py:15-65,
#     rubix/spectra/dust/dust_extinction.py:13-358, extinction_models.py, generic_models.py
# --------------------------------------------------------------------------------------
MSUN_TO_GRAMS = 1.989e33     # rubix/config/rubix_config.yml:7
KPC_TO_CM = 3.08568e21       # rubix/config/rubix_config.yml:6
DUST_EFFECTIVE_WAVELENGTH = 5448.0  # Johnson V, dust_extinction.py:103

#: Remy-Ruyer et al. 2014 table 1 as coded in dust_extinction.py:41-91: (a_high, alpha_high, a_low,
#: alpha_low, x_transition); the single power laws have one branch.
DUST_TO_GAS = {
    ("MW", "power law slope free"): (2.21, 1.62, 2.21, 1.62, -np.inf),
    ("MW", "broken power law fit"): (2.21, 1.00, 0.68, 3.08, 7.96),
    ("Z", "power law slope free"): (2.21, 2.02, 2.21, 2.02, -np.inf),
    ("Z", "broken power law fit"): (2.21, 1.00, 0.96, 3.10, 8.10),
}


def calculate_dust_to_gas_ratio(gas_metallicity, model, Xco, dtype=np.float64):
    """dust_extinction.py:13-93: ``1 / 10**(a + alpha*(8.69 - x))`` with the branch of Table 1."""
    if model == "power law slope fixed":
        raise NotImplementedError("power law slope fixed not implemented yet.")
    a_h, al_h, a_l, al_l, xt = DUST_TO_GAS[(Xco, model)]
    x = np.asarray(gas_metallicity, dtype=dtype)
    x_sol = dtype(8.69)
    with np.errstate(over="ignore", invalid="ignore"):
        hi = dtype(10.0) ** (dtype(a_h) + dtype(al_h) * (x_sol - x))
        lo = dtype(10.0) ** (dtype(a_l) + dtype(al_l) * (x_sol - x))
        return (dtype(1.0) / np.where(x > xt, hi, lo)).astype(dtype)


def dust_extinction_constant(dust_grain_density, effective_wavelength=DUST_EFFECTIVE_WAVELENGTH):
    """dust_extinction.py:150-162: A_V per unit dust surface density (Msun / kpc^2), a host scalar."""
    conv = MSUN_TO_GRAMS / KPC_TO_CM ** 2
    return 3.0 * np.pi * conv / (0.4 * np.log(10.0) * effective_wavelength * 1e-8 * dust_grain_density)


def calculate_extinction(dust_column_density, dust_grain_density, dtype=np.float64):
    return (np.asarray(dust_column_density, dtype=dtype) * dtype(dust_extinction_constant(dust_grain_density))).astype(dtype)


def cardelli89(wave, Rv=3.1):
    """extinction_models.py:104-178, evaluated on ``wave`` exactly as the reference passes it.

    NB the CCM89 polynomials are functions of the wavenumber x = 1/lambda [1/micron], but
    apply_spaxel_extinction hands ``wavelength / 1e4`` (microns) to ``evaluate`` unconverted
    (dust_extinction.py:343-345), so the MUSE band lands in the branch ``0.3 <= wave < 1.1``
    (a = 0.574 wave^1.61, b = -0.527 wave^1.61).  Restated literally: parity is with the reference."""
    w = np.asarray(wave, dtype=np.float64)
    a = np.zeros_like(w)
    b = np.zeros_like(w)
    ir = (0.3 <= w) & (w < 1.1)
    opt = (1.1 <= w) & (w < 3.3)
    nuv = (3.3 <= w) & (w <= 8.0)
    fnuv = (5.9 <= w) & (w <= 8.0)
    fuv = (8.0 < w) & (w <= 10.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        a = np.where(ir, 0.574 * w ** 1.61, a)
        b = np.where(ir, -0.527 * w ** 1.61, b)
        y = w - 1.82
        a = np.where(opt, 1 + 0.17699 * y - 0.50447 * y ** 2 - 0.02427 * y ** 3 + 0.72085 * y ** 4
                     + 0.01979 * y ** 5 - 0.77530 * y ** 6 + 0.32999 * y ** 7, a)
        b = np.where(opt, 1.41338 * y + 2.28305 * y ** 2 + 1.07233 * y ** 3 - 5.38434 * y ** 4
                     - 0.62251 * y ** 5 + 5.30260 * y ** 6 - 2.09002 * y ** 7, b)
        a = np.where(nuv, 1.752 - 0.316 * w - 0.104 / ((w - 4.67) ** 2 + 0.341), a)
        b = np.where(nuv, -3.09 + 1.825 * w + 1.206 / ((w - 4.62) ** 2 + 0.263), b)
        y = w - 5.9
        a = np.where(fnuv, a + (-0.04473 * y ** 2 - 0.009779 * y ** 3), a)
        b = np.where(fnuv, b + (0.2130 * y ** 2 + 0.1207 * y ** 3), b)
        y = w - 8.0
        a = np.where(fuv, -1.073 - 0.628 * y + 0.137 * y ** 2 - 0.070 * y ** 3, a)
        b = np.where(fuv, 13.670 + 4.257 * y - 0.420 * y ** 2 + 0.374 * y ** 3, b)
    return a + b / Rv


def _smoothstep(x, x_min, x_max):
    """helpers.py:_smoothstep with N = 1: 3x^2 - 2x^3 on the clipped, normalised argument."""
    x = np.clip((x - x_min) / (x_max - x_min), 0, 1)
    return (3.0 - 2.0 * x) * x ** 2


def _drude1d(x, amplitude, x_0, fwhm):
    return amplitude * ((fwhm / x_0) ** 2) / ((x / x_0 - x_0 / x) ** 2 + (fwhm / x_0) ** 2)


def _modified_drude(x, scale, x_o, gamma_o, asym):
    gamma = 2.0 * gamma_o / (1.0 + np.exp(asym * (x - x_o)))
    return scale * ((gamma / x_o) ** 2) / ((x / x_o - x_o / x) ** 2 + (gamma / x_o) ** 2)


def _fm90(x, C1, C2, C3, C4, xo, gamma):
    e = C1 + C2 * x
    x2 = x ** 2
    e = e + C3 * (x2 / ((x2 - xo ** 2) ** 2 + x2 * gamma ** 2))
    y = np.where(x >= 5.9, x - 5.9, 0.0)
    return np.where(x >= 5.9, e + C4 * (0.5392 * y ** 2 + 0.05644 * y ** 3), e)


def _g23_nirmir(wave, params):
    (scale, alpha, alpha2, swave, swidth, s1a, s1c, s1f, s1y, s2a, s2c, s2f, s2y) = params
    p1 = scale * wave ** (-alpha)
    p2 = scale * (swave ** (-alpha) / swave ** (-alpha2)) * wave ** (-alpha2)
    wgt = _smoothstep(wave, swave - swidth / 2, swave + swidth / 2)
    return p1 * (1.0 - wgt) + p2 * wgt + _modified_drude(wave, s1a, s1c, s1f, s1y) + _modified_drude(wave, s2a, s2c, s2f, s2y)


def gordon23(wave, Rv=3.1):
    """extinction_models.py:262-389 (Gordon et al. 2023), ``wave`` in microns."""
    w = np.asarray(wave, dtype=np.float64)
    a = np.zeros_like(w)
    b = np.zeros_like(w)
    ir = (1.0 <= w) & (w < 35.0)
    opt = (0.3 <= w) & (w < 1.1)
    uv = (0.09 <= w) & (w <= 0.3)
    optir = (w >= 0.9) & (w <= 1.1)
    uvopt = (w >= 0.3) & (w <= 0.33)
    ir_a = [0.38526, 1.68467, 0.78791, 4.30578, 4.78338, 0.06652, 9.8434, 2.21205, -0.24703,
            0.0267, 19.58294, 17.0, -0.27]
    opt_a = [-0.35848, 0.7122, 0.08746, -0.05403, 0.00674, 0.03893, 2.288, 0.243, 0.02965, 2.054, 0.179,
             0.01747, 1.587, 0.243]
    opt_b = [0.12354, -2.68335, 2.01901, -0.39299, 0.03355, 0.18453, 2.288, 0.243, 0.19728, 2.054, 0.179,
             0.1713, 1.587, 0.243]

    def poly_drude(x, p):
        c = p[:5]
        poly = c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * c[4])))
        return poly + _drude1d(x, p[5], p[6], p[7]) + _drude1d(x, p[8], p[9], p[10]) + _drude1d(x, p[11], p[12], p[13])

    def powerlaw_b(x):
        return -1.01251 * x ** 1.06099   # PowerLaw1d(amplitude=-1.01251, x_0=1, alpha=-1.06099)

    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        x = 1.0 / w
        a = np.where(ir, _g23_nirmir(w, ir_a), a)
        b = np.where(ir, powerlaw_b(w), b)
        a = np.where(opt, poly_drude(x, opt_a), a)
        b = np.where(opt, poly_drude(x, opt_b), b)
        wgt = _smoothstep(w, 0.9, 1.1)
        a = np.where(optir, (1.0 - wgt) * poly_drude(x, opt_a) + wgt * _g23_nirmir(w, ir_a), a)
        b = np.where(optir, (1.0 - wgt) * poly_drude(x, opt_b) + wgt * powerlaw_b(w), b)
        fa = _fm90(x, 0.81297, 0.2775, 1.06295, 0.11303, 4.60, 0.99)
        fb = _fm90(x, -2.97868, 1.89808, 3.10334, 0.65484, 4.60, 0.99)
        a = np.where(uv, fa, a)
        b = np.where(uv, fb, b)
        wgt = _smoothstep(w, 0.3, 0.33)
        a = np.where(uvopt, (1.0 - wgt) * fa + wgt * poly_drude(x, opt_a), a)
        b = np.where(uvopt, (1.0 - wgt) * fb + wgt * poly_drude(x, opt_b), b)
    return a + b * (1.0 / Rv - 1.0 / 3.1)


EXTINCTION_MODELS = {"Cardelli89": cardelli89, "Gordon23": gordon23}


def jnp_interp_left_extrapolate(x, xp, fp):
    """``jnp.interp(x, xp, fp, left="extrapolate")``: as :func:`jnp_interp` without the left clamp."""
    x, xp, fp = np.asarray(x), np.asarray(xp), np.asarray(fp)
    i = np.clip(np.searchsorted(xp, x, side="right"), 1, len(xp) - 1)
    df = fp[i] - fp[i - 1]
    dx = xp[i] - xp[i - 1]
    delta = x - xp[i - 1]
    eps = np.spacing(np.finfo(xp.dtype).eps)
    dx0 = np.abs(dx) <= eps
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        f = np.where(dx0, fp[i - 1], fp[i - 1] + (delta / np.where(dx0, 1, dx)) * df)
    return np.where(x > xp[-1], fp[-1], f).astype(fp.dtype)


def dust_cell_extinction(gas_mass, gas_metals, dust_grain_density, spaxel_area, model="broken power law fit",
                         Xco="Z", dtype=np.float64):
    """dust_extinction.py:283-297: A_V contribution of every gas cell (input order)."""
    m = np.asarray(gas_metals, dtype=dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        log_OH = dtype(12.0) + np.log10(m[:, 4] / (dtype(16.0) * m[:, 0]))
    dtg = calculate_dust_to_gas_ratio(log_OH, model, Xco, dtype=dtype)
    dust_mass = np.asarray(gas_mass, dtype=dtype) * dtg
    return (calculate_extinction(dust_mass, dust_grain_density, dtype=dtype) / dtype(spaxel_area)).astype(dtype)


def stars_av(gas_z, gas_pixel, gas_cell_extinction, star_z, star_pixel, n_spaxels, dtype=np.float64):
    """dust_extinction.py:240-337, the reference algorithm spaxel by spaxel: both particle sets are
    lexsorted by (pixel, z); for every spaxel the gas cells outside it are pushed to z * 1e30 with
    value 0, the cumulative extinction of the cells inside is interpolated at the stars' z
    (``left="extrapolate"``, right end value) and masked to the spaxel's stars.  Returns A_V per star
    in input order.  The 1e30 factor is applied in float32 like the reference (x64 is off)."""
    gz = np.asarray(gas_z, dtype=np.float32)
    gp = np.asarray(gas_pixel)
    sz = np.asarray(star_z, dtype=np.float32)
    sp = np.asarray(star_pixel)
    g_idx = np.lexsort((gz, gp))
    s_idx = np.lexsort((sz, sp))
    gz_s, gp_s = gz[g_idx], gp[g_idx]
    ext_s = np.asarray(gas_cell_extinction, dtype=dtype)[g_idx]
    sz_s, sp_s = sz[s_idx], sp[s_idx]
    ids = np.arange(n_spaxels)
    gb = np.concatenate([np.searchsorted(gp_s, ids, side="left"), [len(g_idx)]])
    sb = np.concatenate([np.searchsorted(sp_s, ids, side="left"), [len(s_idx)]])
    av = np.zeros(len(sz), dtype=dtype)
    ar_g = np.arange(len(g_idx))
    for s in range(n_spaxels):
        s0, s1 = sb[s], sb[s + 1]
        if s1 <= s0:
            continue   # star_mask is empty: the spaxel adds nothing
        gmask = (ar_g >= gb[s]) & (ar_g < gb[s + 1])
        cum = np.cumsum(ext_s * gmask) * gmask
        with np.errstate(over="ignore"):
            xp = (gz_s * np.where(gmask, np.float32(1), np.float32(1e30))).astype(np.float32)
        order = np.argsort(xp, kind="stable")     # jax.lax.sort_key_val is stable
        xp_o, fp_o = xp[order].astype(dtype), cum[order]
        av[s0:s1] = jnp_interp_left_extrapolate(sz_s[s0:s1].astype(dtype), xp_o, fp_o)
    out = np.zeros_like(av)
    out[s_idx] = av                                   # extinction[undo_sort]
    return out


def extinguish(wave_angstrom, av, model="Cardelli89", Rv=3.1, dtype=np.float64):
    """dust_baseclasses.py:126-164 vmapped over stars (dust_extinction.py:341-345):
    ``10 ** (-0.4 * axav(wave / 1e4) * Av)`` -> (n_star, n_wave)."""
    axav = EXTINCTION_MODELS[model](np.asarray(wave_angstrom, dtype=np.float32).astype(np.float64) / 1e4, Rv).astype(dtype)
    return np.power(dtype(10.0), dtype(-0.4) * axav[None, :] * np.asarray(av, dtype=dtype)[:, None])


def apply_spaxel_extinction(spectra, wave_angstrom, gas_z, gas_pixel, gas_mass, gas_metals, star_z, star_pixel,
                            n_spaxels, spaxel_area, dust_cfg, dtype=np.float64):
    """dust_extinction.py:169-358 end to end: (n_star, n_wave) spectra times the per-star extinction."""
    model = dust_cfg["extinction_model"]
    if model not in EXTINCTION_MODELS:
        raise ValueError(f"Extinction model '{model}' is not available. Choose from {list(EXTINCTION_MODELS)}.")
    cell = dust_cell_extinction(gas_mass, gas_metals, dust_cfg["dust_grain_density"], spaxel_area,
                                dust_cfg.get("dust_to_gas_model", "broken power law fit"), dust_cfg.get("Xco", "Z"), dtype)
    av = stars_av(gas_z, gas_pixel, cell, star_z, star_pixel, n_spaxels, dtype)
    return np.asarray(spectra, dtype=dtype) * extinguish(wave_angstrom, av, model, dust_cfg["Rv"], dtype), av


# --------------------------------------------------------------------------------------
# grids
# --------------------------------------------------------------------------------------

# ---- rotate_galaxy: the stage before the path (rubix/galaxy/alignment.py) -----------------------------
def moment_of_inertia_tensor(positions, masses, halfmass_radius, dtype=np.float64):
    """rubix/galaxy/alignment.py:67-125.  The reference selects the particles inside the half-mass
    radius with ``jnp.where(mask, size=N)[0]``, which pads the index list with 0 up to N entries: particle
    0 is therefore added (N - n_inside) more times.  Reproduced, not fixed."""
    pos32 = np.asarray(positions, dtype=np.float32)
    dist = np.sqrt(np.sum(pos32 ** 2, axis=1, dtype=np.float32))
    inside = dist <= np.float32(halfmass_radius)
    idx = np.nonzero(inside)[0]
    idx = np.concatenate([idx, np.zeros(len(pos32) - len(idx), dtype=idx.dtype)])
    p = np.asarray(positions, dtype=dtype)[idx]
    m = np.asarray(masses, dtype=dtype)[idx]
    I = np.zeros((3, 3), dtype=dtype)
    for i in range(3):
        for j in range(3):
            if i == j:
                I[i, j] = np.sum(m * np.sum(p ** 2, axis=1) - m * p[:, i] ** 2)
            else:
                I[i, j] = -np.sum(m * p[:, i] * p[:, j])
    return I


def rotation_matrix_from_inertia_tensor(I, normalise_signs=True):
    """rubix/galaxy/alignment.py:128-146: eigh, eigenvectors ordered by ascending eigenvalue.  eigh's
    eigenvector signs are backend-dependent (LAPACK / cuSOLVER) in the reference itself; with
    ``normalise_signs`` every eigenvector gets its largest-magnitude component positive (the convention of
    the CUDA path), otherwise numpy's LAPACK signs are kept."""
    w, v = np.linalg.eigh(np.asarray(I, dtype=np.float64))
    R = v[:, np.argsort(w, kind="stable")]
    if normalise_signs:
        for c in range(3):
            if R[np.argmax(np.abs(R[:, c])), c] < 0:
                R[:, c] = -R[:, c]
    return R


def euler_rotation_matrix(alpha, beta, gamma):
    """rubix/galaxy/alignment.py:164-209: Rotation.from_euler about x, y, z (degrees), R = R_z R_y R_x."""
    a, b, g = np.deg2rad([alpha, beta, gamma])
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def rotate_galaxy(positions, velocities, masses, halfmass_radius, alpha, beta, gamma, dtype=np.float64,
                  normalise_signs=True):
    """rubix/galaxy/alignment.py:233-265: (p @ R) @ E for coordinates and velocities."""
    I = moment_of_inertia_tensor(positions, masses, halfmass_radius, dtype=np.float64)
    R = rotation_matrix_from_inertia_tensor(I, normalise_signs).astype(dtype)
    E = euler_rotation_matrix(alpha, beta, gamma).astype(dtype)
    pos = (np.asarray(positions, dtype=dtype) @ R) @ E
    vel = (np.asarray(velocities, dtype=dtype) @ R) @ E
    return pos, vel, R


# ---- apply_noise: the stage after the path (rubix/telescope/noise/noise.py) ---------------------------
def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011; jax._src.prng.threefry2x32_p).  uint32 arrays."""
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    ks = [k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA))]
    R = [[13, 15, 26, 6], [17, 29, 16, 24]]
    x0 = np.asarray(x0, dtype=np.uint32).copy()
    x1 = np.asarray(x1, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for g in range(5):
            for r in R[g & 1]:
                x0 += x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 ^= x0
            x0 += ks[(g + 1) % 3]
            x1 += ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def random_bits(key, n):
    """jax's partitionable threefry (default since jax 0.5): element i uses the counter (hi32(i), lo32(i)) and
    bits = x0 ^ x1."""
    i = np.arange(n, dtype=np.uint64)
    a, b = threefry2x32(key, (i >> np.uint64(32)).astype(np.uint32), (i & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    return a ^ b


def _erfinv_f32(x):
    """XLA's float32 ErfInv (Giles 2010, single precision), evaluated in float64 here."""
    x = np.asarray(x, dtype=np.float64)
    w = -np.log1p(-x * x)
    small = w < 5.0
    ws = np.where(small, w - 2.5, np.sqrt(np.maximum(w, 5.0)) - 3.0)
    cs = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
          -0.00417768164, 0.246640727, 1.50140941]
    cl = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
          0.00943887047, 1.00167406, 2.83297682]
    ps = np.zeros_like(ws) + cs[0]
    pl = np.zeros_like(ws) + cl[0]
    for c in cs[1:]:
        ps = ps * ws + c
    for c in cl[1:]:
        pl = pl * ws + c
    return np.where(small, ps, pl) * x


def sample_noise(n, distribution="normal", key=(0, 0)):
    """rubix/telescope/noise/noise.py:8-34 with key = jax.random.PRNGKey(0) = (0, 0): jax.random.uniform takes
    the top 23 bits into [1, 2) - 1; jax.random.normal maps a uniform on (-1, 1) through sqrt(2) erfinv."""
    bits = random_bits(key, n)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    if distribution == "uniform":
        return f.astype(np.float64)
    if distribution != "normal":
        raise ValueError(f"Invalid noise type: {distribution}. Supported types: ['normal', 'uniform']")
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = np.maximum(lo, f * (np.float32(1.0) - lo) + lo)
    return np.sqrt(2.0) * _erfinv_f32(u)


def calculate_S2N(datacube, signal_to_noise):
    """rubix/telescope/noise/noise.py:37-78.  jnp.median propagates NaN: one spaxel without flux -> median 0."""
    flux = np.sum(np.asarray(datacube, dtype=np.float64), axis=-1)
    mask = flux > 0
    median = np.median(flux) if mask.all() and flux.size else 0.0
    factor = np.sqrt(median) / signal_to_noise
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(mask, factor / np.sqrt(flux), 0.0)


def apply_noise(datacube, signal_to_noise, distribution="normal"):
    """rubix/core/noise.py:63-78 + noise.py:81-115: datacube + datacube * N * S2N[:, :, None]."""
    cube = np.asarray(datacube, dtype=np.float64)
    noise = sample_noise(cube.size, distribution).reshape(cube.shape) * calculate_S2N(cube, signal_to_noise)[:, :, None]
    return cube + cube * noise


def calculate_wave_seq(wave_range, wave_res, dtype=np.float32):
    """rubix/telescope/utils.py:53 (``jnp.arange`` in f32).  Pinned bit-exactly by the ``wave``
    dataset of the reference's notebooks/data/dummy_datacube.h5 (tests/golden/muse_wave.npy)."""
    return np.arange(wave_range[0], wave_range[1], wave_res, dtype=dtype)


# --------------------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------------------
def particles_to_cube(coords, velocity, mass, metallicity, age, spatial_bin_edges, num_spaxels,
                      ssp_metallicity, ssp_age, ssp_wavelength, ssp_flux, target_wavelength,
                      redshift, method="cubic", direction="z", dtype=np.float32,
                      apply_filter=True, chunk=20000, acc_dtype=None, extinction=None):
    """filter_particles -> spaxel_assignment -> calculate_spectra -> scale_spectrum_by_mass ->
    doppler_shift_and_resampling -> [calculate_extinction] -> calculate_datacube, in the reference's
    order (rubix/config/pipeline_config.yml:1-45, :62-126), chunked over particles only to bound memory.
    All inputs are float32 arrays; ``dtype`` selects the arithmetic precision.  ``extinction`` (n, W)
    is the per-star factor of the dusty variant (:func:`extinguish`), applied to the resampled spectra."""
    coords = np.asarray(coords, dtype=np.float32)
    edges = np.asarray(spatial_bin_edges, dtype=np.float32)
    if apply_filter:
        mass, metallicity, age, _ = filter_particles(coords, mass, metallicity, age, edges)
    idx = square_spaxel_assignment(coords, edges)
    t = np.asarray(target_wavelength, dtype=dtype)
    lam_z = cosmological_doppler_shift(redshift, ssp_wavelength, dtype=dtype)
    acc_dtype = acc_dtype or dtype
    W = len(t)
    cube = np.zeros((num_spaxels, num_spaxels, W), dtype=acc_dtype)
    n = coords.shape[0]
    for s in range(0, n, chunk):
        e = min(s + chunk, n)
        spec = calculate_spectra(metallicity[s:e], age[s:e], ssp_metallicity, ssp_age, ssp_flux,
                                 method=method, dtype=dtype)
        spec = scale_spectrum_by_mass(spec, np.asarray(mass[s:e], dtype=dtype))
        lam = velocity_doppler_shift(lam_z, velocity[s:e], direction=direction, dtype=dtype)
        res = resample_spectra(spec, lam, t)
        if extinction is not None:
            res = res * np.asarray(extinction[s:e], dtype=dtype)   # dust_extinction.py:356
        cube += calculate_cube(res, idx[s:e], num_spaxels, acc_dtype=acc_dtype)
    return cube, idx


def full_pipeline(*args, psf_kernel=None, lsf_sigma=None, wave_res=None, **kw):
    """particles_to_cube -> convolve_psf -> convolve_lsf (pipeline_config.yml:40-55)."""
    cube, idx = particles_to_cube(*args, **kw)
    if psf_kernel is not None:
        cube = apply_psf(cube, np.asarray(psf_kernel).astype(cube.dtype))
    if lsf_sigma is not None:
        cube = apply_lsf(cube, lsf_sigma, wave_res)
    return cube, idx
