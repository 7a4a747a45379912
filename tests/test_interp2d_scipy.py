"""Independent pin of the oracle's ``interp2d`` restatement (interpax is an unpinned,
un-vendored dependency of the reference, rubix/spectra/ssp/grid.py:113-120, and cannot be installed
here).  scipy is an independent implementation of the same published algorithms:

* linear:  ``scipy.interpolate.RegularGridInterpolator(method="linear")`` -- bilinear interpolation;
* cubic:   interpax's C1 cubic = per-axis cubic Hermite interpolation with node derivatives from
  ``approx_df`` (one-sided secants at the ends, plain mean of the two adjacent secants inside) and
  the cross derivative ``fxy = approx_df_y(fx)``.  A tensor product of
  ``scipy.interpolate.CubicHermiteSpline`` built from those node derivatives evaluates the same
  bicubic patch without the 16x16 ``A_BICUBIC`` matrix or the Hermite-basis weights of the oracle.

The node derivatives are re-derived here with ``numpy.gradient``-free explicit secants, so the only
thing shared with the oracle is the published definition.  This pins the restatement to an
independent implementation of the published algorithm -- still not to interpax itself.
"""

import numpy as np
import pytest
from scipy.interpolate import CubicHermiteSpline, RegularGridInterpolator

from oracle import rubix_oracle as orc


def _queries(x, y, n, seed):
    """Off-node queries plus the awkward ones: cell boundaries, grid edges, just inside the edges."""
    rng = np.random.default_rng(seed)
    xq = rng.uniform(x[0], x[-1], n)
    yq = rng.uniform(y[0], y[-1], n)
    k = min(len(x), n // 8)
    xq[:k] = x[:k]                                   # x on a node, y off-node
    yq[k:k + 40] = y[rng.integers(0, len(y), 40)]    # y on a node, x off-node
    xq[k + 40], yq[k + 40] = x[0], y[0]              # the four corners
    xq[k + 41], yq[k + 41] = x[-1], y[-1]
    xq[k + 42], yq[k + 42] = x[0], y[-1]
    xq[k + 43], yq[k + 43] = x[-1], y[0]
    xq[k + 44] = np.nextafter(x[-1], -np.inf)
    yq[k + 45] = np.nextafter(y[0], np.inf)
    return xq, yq


def _secant_slopes(x, f, axis):
    """Node derivatives as interpax.approx_df(method='cubic') defines them, written out directly."""
    f = np.moveaxis(f, axis, 0)
    s = (f[1:] - f[:-1]) / (x[1:] - x[:-1]).reshape((-1,) + (1,) * (f.ndim - 1))
    d = np.empty_like(f)
    d[0], d[-1] = s[0], s[-1]
    d[1:-1] = 0.5 * (s[:-1] + s[1:])
    return np.moveaxis(d, 0, axis)


@pytest.fixture(scope="module")
def grid(bc03):
    x = bc03["metallicity"].astype(np.float64)
    y = bc03["age"].astype(np.float64)
    f = bc03["flux"].astype(np.float64)[:, :, 380:620:6]   # 40 wavelengths across the MUSE band at z = 0.1
    return x, y, f


def test_linear_matches_regular_grid_interpolator(grid):
    x, y, f = grid
    xq, yq = _queries(x, y, 10000, 1)
    ours = orc.interp2d(xq, yq, x, y, f, method="linear", dtype=np.float64)
    ref = RegularGridInterpolator((x, y), f, method="linear")(np.stack([xq, yq], axis=1))
    scale = np.abs(ref).max()
    assert np.abs(ours - ref).max() <= 1e-12 * scale


@pytest.mark.parametrize("hermite", [False, True])
def test_cubic_matches_tensor_product_hermite_spline(grid, hermite):
    x, y, f = grid
    xq, yq = _queries(x, y, 10000, 2)
    ours = orc.interp2d(xq, yq, x, y, f, method="cubic", dtype=np.float64, hermite=hermite)
    fx = _secant_slopes(x, f, 0)
    fy = _secant_slopes(y, f, 1)
    fxy = _secant_slopes(y, fx, 1)
    # along y at every metallicity node: values and x-derivatives at yq ...
    F = CubicHermiteSpline(y, f, fy, axis=1)(yq)      # (nx, Q, L)
    G = CubicHermiteSpline(y, fx, fxy, axis=1)(yq)    # (nx, Q, L)
    # ... then along x, query by query (the spline of query q evaluated at xq[q] only)
    ref = np.empty_like(ours)
    for lo in range(0, len(xq), 500):
        sl = slice(lo, lo + 500)
        v = CubicHermiteSpline(x, F[:, sl], G[:, sl], axis=0)(xq[sl])   # (q, q, L)
        ref[sl] = v[np.arange(v.shape[0]), np.arange(v.shape[0])]
    scale = np.abs(ref).max()
    assert np.abs(ours - ref).max() <= 1e-11 * scale


def test_float32_restatement_close_to_the_pinned_float64(grid, bc03):
    """The float32 op-order mirror (what the CUDA kernels are compared with per stage) stays within float32
    rounding of the scipy-pinned float64 evaluation, both methods."""
    x, y, f = grid
    xq, yq = _queries(x, y, 3000, 3)
    xq32, yq32 = xq.astype(np.float32), yq.astype(np.float32)
    for method in ("linear", "cubic"):
        a = orc.interp2d(xq32, yq32, bc03["metallicity"], bc03["age"], bc03["flux"][:, :, 380:620:6], method=method,
                         dtype=np.float32)
        b = orc.interp2d(xq32.astype(np.float64), yq32.astype(np.float64), x, y, f, method=method, dtype=np.float64)
        assert np.abs(a - b).max() <= 4e-6 * np.abs(b).max()


def test_outside_the_grid_is_zero_and_edges_are_inclusive(grid):
    # extrap=0: strictly outside -> 0; exactly on the outer nodes -> interpolated (inclusive), as interpax
    x, y, f = grid
    for method in ("linear", "cubic"):
        out = orc.interp2d([x[0], x[-1], np.nextafter(x[-1], np.inf), x[2]],
                           [y[0], y[-1], y[5], np.nextafter(y[0], -np.inf)], x, y, f, method=method, dtype=np.float64)
        assert np.allclose(out[0], f[0, 0], rtol=1e-12) and np.allclose(out[1], f[-1, -1], rtol=1e-12)
        assert (out[2] == 0).all() and (out[3] == 0).all()
