"""GPU test of rbx_sort_by_spaxel (SURVEY 8b minimum set; the north star's "device radix sort of particles by
spaxel index"): integer work, so everything is compared BIT-EXACTLY with numpy's stable argsort of the same ids --
the order of equal ids is the particle order, which is the summation order of the reference's segment_sum on one
device (rubix/spectra/ifu.py:286-287)."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rubix_b200 import ops as _ops
    return _ops


def _expect(pix, nseg):
    key = np.where((pix >= 0) & (pix < nseg), pix, nseg).astype(np.int64)
    order = np.argsort(key, kind="stable").astype(np.int32)
    srt = key[order].astype(np.int32)
    off = np.searchsorted(srt, np.arange(nseg + 1), side="left").astype(np.int32)
    return order, srt, off


def _check(ops, pix, nseg):
    order, srt, off = ops.sort_by_spaxel(pix, nseg)
    torch.cuda.synchronize()
    e_order, e_srt, e_off = _expect(np.asarray(pix), nseg)
    assert np.array_equal(order.cpu().numpy(), e_order)
    assert np.array_equal(srt.cpu().numpy(), e_srt)
    assert np.array_equal(off.cpu().numpy(), e_off)


# tile = 4096 keys: below, at and just above one tile, ragged last tiles, many tiles (look-back across tiles)
@pytest.mark.parametrize("n", [1, 2, 31, 4095, 4096, 4097, 12289, 200_003])
# key widths of 1, 2, 10, 15 and 17 bits: one, two and three radix passes (the output pair alternates with the parity)
@pytest.mark.parametrize("nseg", [1, 2, 625, 22500, 70000])
def test_sort_by_spaxel_matches_stable_argsort(ops, n, nseg):
    rng = np.random.default_rng(42 + n + nseg)
    # ids with dropped ones on both sides: -1 (rbx_filter_and_assign's mark) and >= nseg
    pix = rng.integers(-1, nseg + 2, size=n).astype(np.int32)
    _check(ops, pix, nseg)


def test_sort_by_spaxel_of_assigned_particles_at_bench_size(ops):
    """10^6 bench-G particles through rbx_filter_and_assign (-1 outside the aperture) -> sort: the MUSE key layout
    (625 spaxels + the dropped bucket: 10 bits, two passes), heavily skewed towards the central spaxels."""
    from rubix_b200 import synthetic
    d = synthetic.bench_g(1_000_000)
    edges = synthetic.spatial_edges(25)
    pix = ops.filter_and_assign(d["coords"], edges)
    assert int((pix < 0).sum()) > 0
    _check(ops, pix.cpu().numpy(), 625)
    # bit-reproducible: a second run gives the same permutation
    a = ops.sort_by_spaxel(pix, 625)[0]
    b = ops.sort_by_spaxel(pix, 625)[0]
    assert torch.equal(a, b)


def test_sort_by_spaxel_edge_cases(ops):
    # empty input: offsets all zero
    order, srt, off = ops.sort_by_spaxel(np.zeros(0, np.int32), 625)
    assert order.numel() == 0 and srt.numel() == 0 and int(off.abs().sum()) == 0 and off.numel() == 626
    # every particle in one spaxel / every particle dropped / already sorted / reversed
    n = 50_000
    _check(ops, np.full(n, 312, np.int32), 625)
    _check(ops, np.full(n, -1, np.int32), 625)
    _check(ops, np.sort(np.random.default_rng(1).integers(0, 625, n)).astype(np.int32), 625)
    _check(ops, np.sort(np.random.default_rng(2).integers(0, 625, n))[::-1].astype(np.int32).copy(), 625)
    # outputs are optional
    order, srt, off = ops.sort_by_spaxel(np.arange(100, dtype=np.int32)[::-1].copy(), 128, with_sorted=False,
                                         with_offsets=False)
    assert srt is None and off is None
    assert np.array_equal(order.cpu().numpy(), np.arange(100, dtype=np.int32)[::-1])


def test_sort_feeds_segment_sum_order(ops, bc03, muse_wave):
    """The permutation groups the particles exactly as the cube build does: summing the staged spectra run by run in
    sorted order gives the cube of rbx_segment_sum (float32 adds in particle order inside a spaxel) to rounding."""
    from rubix_b200 import synthetic
    d = synthetic.bench_g(3000)
    edges = synthetic.spatial_edges(25)
    plan = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1,
                    method="linear")
    pix = ops.filter_and_assign(d["coords"], edges)
    spec = ops.doppler_resample(plan, ops.scale_by_mass(ops.ssp_lookup(plan, d["metallicity"], d["age"]), d["mass"]),
                                d["velocity"])
    cube = ops.segment_sum(spec, pix, 625).cpu().numpy().reshape(625, -1)
    order, srt, off = (t.cpu().numpy() for t in ops.sort_by_spaxel(pix, 625))
    spec = spec.cpu().numpy().astype(np.float64)
    ref = np.zeros_like(cube, dtype=np.float64)
    for s in range(625):
        ref[s] = spec[order[off[s]:off[s + 1]]].sum(axis=0)
    assert np.abs(cube - ref).max() <= 4e-6 * ref.max()


def test_sorted_segment_sum_is_the_sequential_float32_sum(ops, bc03, muse_wave):
    """rbx_segment_sum_sorted: bit-identical to the oracle's float32 calculate_cube (np.add.at: one particle after the
    other in particle order) on the same staged spectra, bit-identical across reruns, and within
    rounding of the atomic form."""
    from oracle import rubix_oracle as orc
    from rubix_b200 import synthetic
    d = synthetic.bench_g(4000)
    edges = synthetic.spatial_edges(25)
    plan = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1,
                    method="cubic")
    pix = ops.filter_and_assign(d["coords"], edges)
    spec = ops.doppler_resample(plan, ops.scale_by_mass(ops.ssp_lookup(plan, d["metallicity"], d["age"]), d["mass"]),
                                d["velocity"])
    det = ops.segment_sum(spec, pix, 625, deterministic=True)
    again = ops.segment_sum(spec, pix, 625, deterministic=True)
    atomic = ops.segment_sum(spec, pix, 625)
    assert torch.equal(det, again)
    ref = orc.calculate_cube(spec.cpu().numpy(), pix.cpu().numpy(), 25).reshape(625, -1)
    assert ref.dtype == np.float32
    assert np.array_equal(det.cpu().numpy(), ref)
    assert float((det - atomic).abs().max()) <= 4e-6 * float(ref.max())
    # an empty galaxy gives a zero cube
    z = ops.segment_sum(torch.zeros((0, plan.W), device="cuda"), torch.zeros(0, dtype=torch.int32, device="cuda"), 625,
                        deterministic=True)
    assert z.shape == (625, plan.W) and float(z.abs().max()) == 0.0


def test_factory_staged_deterministic_option(bc03, muse_wave, tng_subset):
    """config["b200"]["deterministic"]: the staged calculate_datacube sums in sorted runs -- bit-identical reruns,
    within rounding of the atomic form, and ids >= sbin^2 (26 bins per axis from the 27 edges) are dropped alike."""
    import copy
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rubix_b200 import core
    from test_gpu_factories import CONFIG, _run_chain
    d = {k: v[:2000].copy() for k, v in tng_subset.items()}
    cubes = []
    for det in (True, True, False):
        cfg = copy.deepcopy(CONFIG)
        cfg["ssp"]["method"] = "linear"
        cfg["b200"] = {"fused": False, "deterministic": det}
        cubes.append(_run_chain(core, cfg, d).stars.datacube)
    assert torch.equal(cubes[0], cubes[1])
    assert float((cubes[0] - cubes[2]).abs().max()) <= 4e-6 * float(cubes[2].max())
    assert float(cubes[0].abs().max()) > 0
