"""Property tests (hypothesis, CPU) of the multi-GPU host logic in rubix_b200/parallel.py: the contiguous particle split
of rubix/core/data.py:471-482 and the slab-major layout of the large-FOV exchange."""

import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

from rubix_b200 import _lib, parallel


@settings(max_examples=300, deadline=None)
@given(n=st.integers(0, 5000), world=st.integers(1, 16))
def test_shard_ranges_partition_the_particles_like_reshape_array(n, world):
    """Contiguous, disjoint, in rank order, ceil(n / world) per rank except the tail -- the reference pads the tail with
    zero-mass particles instead (reshape_array), which contribute exactly 0."""
    per = -(-n // world)
    ranges = [parallel.shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for r, (lo, hi) in enumerate(ranges):
        assert 0 <= lo <= hi <= n and hi - lo <= per
        assert (lo, hi) == (min(r * per, n), min((r + 1) * per, n))
        if r:
            assert lo == ranges[r - 1][1]
    data = {"a": np.arange(n), "b": np.arange(2 * n).reshape(n, 2)}
    got = [parallel.shard_particles(data, r, world) for r in range(world)]
    assert np.array_equal(np.concatenate([g["a"] for g in got]), data["a"])
    assert np.array_equal(np.concatenate([g["b"] for g in got]), data["b"])


@settings(max_examples=200, deadline=None)
@given(W=st.integers(1, 400), nslab=st.integers(1, 9), halo=st.integers(0, 14), nseg=st.integers(1, 4))
def test_slab_pack_and_interior_are_inverse_and_halos_hold_the_neighbours(W, nslab, halo, nseg):
    rng = np.random.default_rng(W * 131 + nslab * 17 + halo)
    cube = rng.random((nseg, W))
    wslab, ws = parallel.slab_geometry(W, nslab, halo)
    assert wslab == -(-W // nslab) and ws == wslab + 2 * halo          # trailing slabs may own no channel at all
    a, b = C.c_int(), C.c_int()
    assert _lib.lib().rbx_slab_geometry(W, nslab, halo, C.byref(a), C.byref(b)) == 0 and (a.value, b.value) == (wslab, ws)
    packed = parallel.slab_pack(cube, nslab, halo)
    assert packed.shape == (nslab, nseg, ws)
    # the owned channels of all slabs, in rank order, are the cube
    own = [parallel.slab_interior(packed[r], W, r, nslab, halo) for r in range(nslab)]
    assert np.array_equal(np.concatenate(own, axis=-1), cube)
    # every stored channel is the cube's channel at that wavelength, or zero padding outside [0, W)
    for r in range(nslab):
        lo = r * wslab - halo
        idx = np.arange(lo, lo + ws)
        inside = (idx >= 0) & (idx < W)
        assert np.array_equal(packed[r][:, inside], cube[:, idx[inside]])
        assert not packed[r][:, ~inside].any()
    # packing is linear: the reduce-scatter of packed partial cubes is the packed sum
    other = rng.random((nseg, W))
    assert np.allclose(parallel.slab_pack(cube + other, nslab, halo), packed + parallel.slab_pack(other, nslab, halo),
                       rtol=0, atol=1e-15)
