"""A float64 numpy model of the ALGORITHM the cube kernels use -- test infrastructure.

The reference evaluates every particle on every telescope channel (``jnp.interp`` + two sums, rubix/spectra/ifu.py:
241-260).  The CUDA kernels never touch a channel per particle: the resampled spectrum p(t) is piecewise linear with its
break points at the Doppler-shifted SSP knots, so

* p(t_w) = s_0 + sum over the knots x_j < t_w of dm_j (t_w - x_j), with the slope changes ("kinks")
  dm_j = m_j - m_{j-1} and m_{-1} = m_{L-1} = 0 (``jnp.interp`` clamps to the end values): a knot only adds
  (dm_j, dm_j x_j) to the CELL k_j of the first channel at or above it, and one prefix sum over the channels per SPAXEL
  turns the summed cells into the summed spectra;
* total = sum_j s_j (x_j - x_{j-1}) [x_j in band] needs the knots only;
* new = sum_w p(t_w) dt_w is evaluated per knot segment: the channel widths of a segment telescope to
  D_j = t[k_{j+1}-1] - t[k_j-1], and sum_w dt_w (t_w - x_j) has the closed form D_j (D_j / 2 + t[k_j-1] - x_j + delta / 2)
  on an arange grid (a prefix table on any other grid).

This file states exactly that in numpy, in double precision and without the float32 devices of the kernels (channel
units, chunk re-anchoring, window tables), so that tests/test_knot_reformulation_cpu.py can show the reformulation is
EXACT -- equal to the reference formula to float64 rounding -- independently of any GPU run.
"""

import numpy as np


def particle_knots(s, x, t):
    """One particle: knots ``x`` (L,), mass-weighted spectrum ``s`` (L,), channels ``t`` (W,), all float64.
    Returns (cells k_j, kinks dm_j, total, new, s_0) with ``new`` from the per-segment closed sums."""
    L, W = len(x), len(t)
    m = np.zeros(L + 1)                       # m[j + 1] = slope on [x_j, x_{j+1}]; m[0] = m[L] = 0 (end clamps)
    m[1:L] = np.diff(s) / np.diff(x)
    dm = m[1:] - m[:-1]                       # kink at knot j
    k = np.searchsorted(t, x, side="left")    # first channel with t_w >= x_j (W when the knot is beyond the band)
    in_band = (x >= t[0]) & (x <= t[-1])
    total = np.sum(s[1:] * np.diff(x) * in_band[1:])          # diff0: the first difference is 0
    # new = sum_{w >= 1} p(t_w) dt_w, segment by segment.  P1[w] = sum_{u <= w} dt_u, P2[w] = sum_{u <= w} dt_u t_u
    dt = np.diff(t, prepend=t[0])
    P1 = np.concatenate([[0.0], np.cumsum(dt)])               # P1[w + 1] = sum up to and including channel w
    P2 = np.concatenate([[0.0], np.cumsum(dt * t)])
    lo = np.concatenate([[0], k])                             # segment -1 (clamp to s_0), 0 .. L-2, L-1 (clamp to s_{L-1})
    hi = np.concatenate([k, [W]])
    base = np.concatenate([[s[0]], s])                        # value at the segment's left knot
    slope = np.concatenate([[0.0], m[1:L], [0.0]])
    left = np.concatenate([[0.0], x])                         # the left knot itself (unused where the slope is 0)
    D = P1[hi] - P1[lo]                                       # telescoping channel widths of the segment
    T = (P2[hi] - P2[lo]) - left * D                          # sum dt_w (t_w - x_j)
    new = np.sum(base * D + slope * T)
    return k, dm, total, new, s[0]


def arange_segment_moment(t, k_lo, k_hi, x):
    """The closed form the kernels use on an arange grid for sum_{w = k_lo}^{k_hi - 1} dt_w (t_w - x):
    D (D / 2 + t[k_lo - 1] - x + delta / 2) with D = t[k_hi - 1] - t[k_lo - 1] (needs k_lo >= 1)."""
    delta = t[1] - t[0]
    D = t[k_hi - 1] - t[k_lo - 1]
    return D * (D / 2 + t[k_lo - 1] - x + delta / 2)


def particles_to_cube_knots(spectra, knots, pixel, num_segments, t):
    """The cube from cells: every particle adds scale * (dm_j, dm_j x_j) to cell k_j of its spaxel row and scale * s_0 to
    the row's constant; ONE prefix sum per spaxel expands the cells.  ``spectra`` / ``knots`` are (n, L) float64."""
    W = len(t)
    A = np.zeros((num_segments, W + 1))       # sum of kinks per cell (cell W: knots beyond the band, never expanded)
    B = np.zeros((num_segments, W + 1))       # sum of kink * knot position per cell
    C = np.zeros(num_segments)                # sum of scale * s_0
    for p in range(len(spectra)):
        seg = pixel[p]
        if seg < 0 or seg >= num_segments:
            continue
        k, dm, total, new, s0 = particle_knots(spectra[p], knots[p], t)
        with np.errstate(invalid="ignore", divide="ignore"):
            scale = np.nan_to_num(total / new, nan=0.0)
        np.add.at(A[seg], k, scale * dm)
        np.add.at(B[seg], k, scale * dm * knots[p])
        C[seg] += scale * s0
    # p(t_w) = s_0 + t_w * sum_{k <= w} A_k - sum_{k <= w} B_k
    return C[:, None] + t[None, :] * np.cumsum(A[:, :W], axis=1) - np.cumsum(B[:, :W], axis=1)
