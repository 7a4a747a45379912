"""The algorithm of the cube kernels is an exact reformulation of the reference formula, shown on the CPU.

tests/knot_model.py states in float64 numpy what DESIGN.md section 5 describes -- kinks accumulated per channel cell,
the two normalisation sums from the knots alone, one prefix sum per spaxel -- and this file holds it to the oracle's
float64 evaluation of the reference (jnp.interp on every channel, rubix/spectra/ifu.py:241-260): they agree to
float64 rounding, so what the GPU parity tests measure is float32 rounding of the kernels, not an approximation in the
method."""

import numpy as np
import pytest

from oracle import rubix_oracle as orc

from knot_model import arange_segment_moment, particle_knots, particles_to_cube_knots


def _particles(bc03, n, seed, vmax=300.0):
    rng = np.random.default_rng(seed)
    zq = rng.uniform(1e-4, 0.05, n).astype(np.float32)
    aq = rng.uniform(5.1, 10.3, n).astype(np.float32)
    spec = orc.interp2d(zq, aq, bc03["metallicity"], bc03["age"], bc03["flux"], method="linear", dtype=np.float64)
    spec = spec * rng.uniform(0.5, 1.5, n)[:, None]
    vel = np.zeros((n, 3))
    vel[:, 2] = rng.normal(0, vmax, n)
    lam_z = orc.cosmological_doppler_shift(0.1, bc03["wavelength"], dtype=np.float64)
    return spec, orc.velocity_doppler_shift(lam_z, vel, "z", dtype=np.float64)


@pytest.mark.parametrize("grid", ["muse", "coarse", "non_arange", "beyond_ssp"])
def test_knot_form_equals_the_reference_formula(bc03, muse_wave, grid):
    if grid == "muse":
        t = muse_wave.astype(np.float64)
    elif grid == "coarse":                      # channels wider than the SSP spacing: several knots per cell
        t = np.arange(4700.15, 9351.4, 30.0)
    elif grid == "non_arange":                  # slowly growing channel width: the prefix-table path
        t = 5000.0 + np.cumsum(1.25 * (1 + 0.2 * np.linspace(0, 1, 1500)))
    else:                                       # a band that sticks out of the shifted SSP range: both end clamps
        t = np.linspace(60.0, 24000.0, 900)
    spec, knots = _particles(bc03, 60, 11)
    ref = orc.resample_spectra(spec, knots, t)
    W = len(t)
    for p in range(len(spec)):
        k, dm, total, new, s0 = particle_knots(spec[p], knots[p], t)
        # the spectrum on the channels from the cells: one prefix sum
        A, B = np.zeros(W + 1), np.zeros(W + 1)
        np.add.at(A, k, dm)
        np.add.at(B, k, dm * knots[p])
        interp = s0 + t * np.cumsum(A[:W]) - np.cumsum(B[:W])
        direct = orc.jnp_interp(t, knots[p], spec[p])
        assert np.abs(interp - direct).max() <= 1e-10 * np.abs(direct).max()
        # the two sums against the reference's per-channel / per-knot sums
        in_band = (knots[p] >= t.min()) & (knots[p] <= t.max())
        assert np.isclose(total, np.sum(spec[p] * orc.calculate_diff(knots[p]) * in_band), rtol=1e-13, atol=0)
        assert np.isclose(new, np.sum(direct * orc.calculate_diff(t)), rtol=1e-11, atol=0)
        out = interp * np.nan_to_num(total / new, nan=0.0)
        assert np.abs(out - ref[p]).max() <= 2e-10 * np.abs(ref[p]).max()


def test_arange_closed_form_of_the_segment_moment(muse_wave):
    """sum_w dt_w (t_w - x) over a knot segment = D (D / 2 + t[k-1] - x + delta / 2) on an exactly uniform grid."""
    t = 4700.0 + 1.25 * np.arange(3721)          # exactly representable: the identity holds to rounding
    rng = np.random.default_rng(5)
    for _ in range(200):
        k_lo = int(rng.integers(1, 3600))
        k_hi = k_lo + int(rng.integers(1, 40))
        x = t[k_lo] - rng.uniform(0, 1.25)
        direct = np.sum(1.25 * (t[k_lo:k_hi] - x))
        assert np.isclose(arange_segment_moment(t, k_lo, k_hi, x), direct, rtol=1e-12)


def test_cube_from_cells_equals_the_reference_cube(bc03, muse_wave):
    """Many particles per spaxel, accumulated as cells and expanded once per spaxel, against segment_sum of the
    per-particle resampled spectra (rubix/spectra/ifu.py:286-287); ids outside the cube are dropped."""
    t = muse_wave.astype(np.float64)
    spec, knots = _particles(bc03, 240, 3, vmax=600.0)
    rng = np.random.default_rng(9)
    pix = rng.integers(-1, 11, len(spec))        # 9 spaxels plus ids that are dropped on either side
    spec[7] = 0.0                                # a zero spectrum: total / new = 0 / 0 -> 0
    ref = orc.calculate_cube(orc.resample_spectra(spec, knots, t), pix, 3).reshape(9, -1)
    got = particles_to_cube_knots(spec, knots, pix, 9, t)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()


def test_channel_unit_knot_positions_are_at_least_as_accurate_as_the_reference_float32(bc03, muse_wave):
    """The kernels place a knot with ONE float32 FMA in channel units, u' = a eps + b (a = lam_z / delta,
    b = (lam_z - t0) / delta - 1/2 rounded once from double in plan.cu, eps = expm1(v / c) in the particle record;
    DESIGN.md section 5, 'channel units').  Emulated here with numpy -- the FMA as a float64 product (exact for two
    float32 factors) and sum, rounded once to float32 -- and compared, in units of channels, with the exact position of
    the same float32 inputs and with what the reference's own float32 arithmetic gives (lam_z * exp(v / c),
    rubix/spectra/ifu.py:190): the claim is an error below 3e-4 channels, no worse than the reference's."""
    f = np.float32
    lam_z = (f(1.1) * bc03["wavelength"]).astype(f)                     # rubix/spectra/ifu.py:80 in float32
    t0, delta = float(muse_wave[0]), 1.25
    band = (lam_z > 4500) & (lam_z < 9500)
    lz = lam_z[band].astype(np.float64)
    a = (lz / delta).astype(f)
    b = ((lz - t0) / delta - 0.5).astype(f)
    rng = np.random.default_rng(2)
    v = np.concatenate([rng.normal(0, 300, 4000), rng.uniform(-1000, 1000, 1000)]).astype(f)
    c = 299792.458
    eps = np.expm1(v.astype(np.float64) / c).astype(f)                  # expm1f: correctly rounded to within an ulp
    u_kernel = (a.astype(np.float64)[None, :] * eps.astype(np.float64)[:, None] + b.astype(np.float64)[None, :]).astype(f)
    u_exact = (lz[None, :] * np.exp(v.astype(np.float64) / c)[:, None] - t0) / delta - 0.5
    d32 = np.exp((v / f(c)).astype(f)).astype(f)                        # the reference: exp(v / c) in float32 ...
    x32 = (lam_z[band][None, :] * d32[:, None]).astype(f)               # ... times lam_z in float32
    u_ref32 = (x32.astype(np.float64) - t0) / delta - 0.5
    err_kernel = np.abs(u_kernel.astype(np.float64) - u_exact).max()
    err_ref32 = np.abs(u_ref32 - u_exact).max()
    print(f"knot position error in channels: kernel form {err_kernel:.2e}, reference float32 form {err_ref32:.2e}")
    assert err_kernel <= 3e-4
    assert err_kernel <= err_ref32
