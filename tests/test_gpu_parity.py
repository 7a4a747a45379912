"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Integer work (spaxel ids, masks) must be bit-exact.  Floating-point tolerances, all relative to
the float64 evaluation of the reference formulas on the same float32 inputs:

* per-stage arrays:  max |delta| <= 4e-6 * max |ref|      (a few float32 ulps of accumulated rounding)
* cubes:             max |delta| <= 5e-6 * max |cube|      and the north-star bound
                     max |delta| <= 1e-5 * sum |cube| / n_voxel_scale (see _cube_close)
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import c_oracle  # noqa: E402
from oracle import rubix_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rubix_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def plans(ops, bc03, muse_wave):
    out = {}
    for m in ("linear", "cubic"):
        out[m] = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1,
                          method=m, direction="z")
    return out


from helpers import cube_close as _cube_close, well_conditioned as _well_conditioned  # noqa: E402


# ---- a0 -------------------------------------------------------------------------------------------
def test_spaxel_assign_kat(ops):
    # tests/test_telescope_utils.py:20-35 (z column added: our coords are (n,3))
    coords = np.array([[0.5, 1.5, 0], [2.5, 3.5, 0]], dtype=np.float32)
    out = ops.spaxel_assign(coords, np.array([0, 1, 2, 3, 4], dtype=np.float32))
    assert out.dtype == torch.int32
    assert out.cpu().tolist() == [4, 14]


def test_spaxel_assign_bit_exact_random(ops):
    from rubix_b200.synthetic import spatial_edges
    rng = np.random.default_rng(0)
    for S in (25, 150):
        edges = spatial_edges(S)
        coords = rng.normal(0, 3.0, (200000, 3)).astype(np.float32)
        # put particles exactly on edges and just beside them
        k = len(edges)
        coords[:k, 0] = edges
        coords[k:2 * k, 1] = edges
        coords[2 * k:3 * k, 0] = np.nextafter(edges, np.float32(np.inf))
        coords[3 * k:4 * k, 1] = np.nextafter(edges, np.float32(-np.inf))
        pix, mask = ops.spaxel_assign(coords, edges, with_mask=True)
        assert np.array_equal(pix.cpu().numpy(), orc.square_spaxel_assignment(coords, edges))
        assert np.array_equal(mask.cpu().numpy(), orc.mask_particles_outside_aperture(coords, edges))
    # reference quirk: len(edges) - 1 may differ from num_spaxels (S+2 edges); ids follow the edges
    edges = np.linspace(-1, 1, 28).astype(np.float32)
    pix = ops.spaxel_assign(coords, edges)
    assert np.array_equal(pix.cpu().numpy(), orc.square_spaxel_assignment(coords, edges))


def test_mask_boundaries_and_empty(ops):
    # tests/test_telescope_utils.py:38-79
    e = np.array([0, 1], dtype=np.float32)
    cases = [([[0.5, 0.5, 0], [0.2, 0.2, 0]], 2), ([[1.5, 1.5, 0], [-0.1, -0.1, 0]], 0),
             ([[0, 0, 0], [1, 1, 0], [0, 1, 0], [1, 0, 0]], 4),
             ([[0.5, 0.5, 0], [1.5, 1.5, 0], [0, 0, 0], [-0.1, -0.1, 0]], 2)]
    for coords, expected in cases:
        _, m = ops.spaxel_assign(np.array(coords, dtype=np.float32), e, with_mask=True)
        assert int(m.sum()) == expected
    pix, m = ops.spaxel_assign(np.zeros((0, 3), dtype=np.float32), e, with_mask=True)
    assert pix.numel() == 0 and m.numel() == 0


def test_filter_particles(ops):
    rng = np.random.default_rng(1)
    coords = rng.uniform(-2, 2, (5000, 3)).astype(np.float32)
    e = np.linspace(-1, 1, 11).astype(np.float32)
    mass = ops.dev(rng.random(5000).astype(np.float32) + 1)
    met = ops.dev(rng.random(5000).astype(np.float32) + 1)
    age = ops.dev(rng.random(5000).astype(np.float32) + 1)
    m0, z0, a0 = mass.cpu().numpy().copy(), met.cpu().numpy().copy(), age.cpu().numpy().copy()
    mask = ops.filter_particles(coords, e, mass, met, age).cpu().numpy()
    em, ez, ea, emask = orc.filter_particles(coords, m0, z0, a0, e)
    assert np.array_equal(mask, emask)
    assert np.array_equal(mass.cpu().numpy(), em) and np.array_equal(met.cpu().numpy(), ez)
    assert np.array_equal(age.cpu().numpy(), ea)


def test_filter_and_assign_one_pass(ops):
    from rubix_b200.synthetic import bench_g, spatial_edges
    d = bench_g(50000, sigma_kpc=3.0)
    edges = spatial_edges(25)
    ref_idx, ref_mask = c_oracle.spaxel_assign(d["coords"], edges)
    assert (~ref_mask).sum() > 100
    # marking mode: nothing modified, pixel -1 outside the aperture
    pix = ops.filter_and_assign(d["coords"], edges).cpu().numpy()
    assert np.array_equal(pix[ref_mask], ref_idx[ref_mask]) and (pix[~ref_mask] == -1).all()
    # in-place mode: same as filter_particles followed by spaxel_assign
    mass, met, age = ops.dev(d["mass"]).clone(), ops.dev(d["metallicity"]).clone(), ops.dev(d["age"]).clone()
    pix2 = ops.filter_and_assign(d["coords"], edges, mass, met, age).cpu().numpy()
    assert np.array_equal(pix2, ref_idx)
    for t, h in ((mass, d["mass"]), (met, d["metallicity"]), (age, d["age"])):
        assert np.array_equal(t.cpu().numpy(), np.where(ref_mask, h, np.float32(0)))


# ---- a1 -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_ssp_lookup_nodes_and_outside(ops, plans, bc03, method):
    # tests/test_core_ssp.py:158-173 (grid-node identity, rtol 1e-5 / atol 1e-6) and :116-126,176-181
    Z, A = np.meshgrid(bc03["metallicity"], bc03["age"], indexing="ij")
    out = ops.ssp_lookup(plans[method], Z.ravel(), A.ravel()).cpu().numpy()
    assert np.allclose(out, bc03["flux"].reshape(-1, 842), rtol=1e-5, atol=1e-6)
    zq = np.array([0.1, 1e-5, 0.02, 0.02, 0.0], dtype=np.float32)
    aq = np.array([5.0, 5.0, 11.0, -1.0, 0.0], dtype=np.float32)
    out = ops.ssp_lookup(plans[method], zq, aq).cpu().numpy()
    assert (out == 0).all()


@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_ssp_lookup_off_node(ops, plans, bc03, method):
    rng = np.random.default_rng(2)
    zq = rng.uniform(1e-4, 0.05, 3000).astype(np.float32)
    aq = rng.uniform(0.0, 10.3, 3000).astype(np.float32)
    out = ops.ssp_lookup(plans[method], zq, aq).cpu().numpy().astype(np.float64)
    ref = orc.interp2d(zq, aq, bc03["metallicity"], bc03["age"], bc03["flux"], method=method, dtype=np.float64)
    err = np.abs(out - ref).max()
    print(f"[lookup {method}] max|d|={err:.3e} max|ref|={np.abs(ref).max():.3e}")
    assert err <= 4e-6 * np.abs(ref).max()


# ---- a2 -------------------------------------------------------------------------------------------
def test_scale_by_mass_exact(ops):
    # tests/test_core_ifu.py:249-282 (array_equal)
    rng = np.random.default_rng(3)
    spec = rng.random((257, 842)).astype(np.float32)
    mass = (rng.random(257) * 1e5).astype(np.float32)
    out = ops.scale_by_mass(spec, mass).cpu().numpy()
    assert np.array_equal(out, spec * mass[:, None])


# ---- a3 + a4 --------------------------------------------------------------------------------------
def test_resample_kat(ops):
    # tests/test_spectra_ifu.py:175-228 through a 5-bin "template" plan
    lam = np.array([4000.0, 5000.0, 6000.0, 7000.0, 8000.0], dtype=np.float32)
    t = np.array([4500.0, 5500.0, 6500.0, 7500.0], dtype=np.float32)
    flux = np.zeros((2, 2, 5), dtype=np.float32)
    plan = ops.Plan([0.0, 1.0], [0.0, 1.0], lam, flux, t, 0.0, method="linear", direction="y")
    s = np.array([[1.0, 2.0, 3.0, 4.0, 5.0], [0, 0, 0, 0, 0]], dtype=np.float32)
    out = ops.doppler_resample(plan, s, np.zeros((2, 3), dtype=np.float32)).cpu().numpy()
    assert np.allclose(out[0], [1.2857143, 2.142857, 3.0, 3.857143], rtol=1e-5)
    assert (out[1] == 0).all() and not np.isnan(out).any()


def test_doppler_resample_matches_oracle(ops, plans, bc03, muse_wave):
    rng = np.random.default_rng(4)
    n = 400
    zq = rng.uniform(1e-4, 0.05, n).astype(np.float32)
    aq = rng.uniform(5.1, 10.3, n).astype(np.float32)
    vel = rng.normal(0, 200, (n, 3)).astype(np.float32)
    spec = orc.interp2d(zq, aq, bc03["metallicity"], bc03["age"], bc03["flux"], method="linear")
    out = ops.doppler_resample(plans["linear"], spec, vel).cpu().numpy().astype(np.float64)
    lam = orc.velocity_doppler_shift(orc.cosmological_doppler_shift(0.1, bc03["wavelength"], np.float64), vel,
                                     "z", dtype=np.float64)
    ref = orc.resample_spectra(spec.astype(np.float64), lam, muse_wave.astype(np.float64))
    err = np.abs(out - ref).max()
    print(f"[resample] max|d|={err:.3e} max|ref|={np.abs(ref).max():.3e}")
    assert err <= 4e-6 * np.abs(ref).max()


# ---- a5 -------------------------------------------------------------------------------------------
def test_segment_sum_kat(ops):
    # tests/test_spectra_ifu.py:231-257, plus out-of-range ids are dropped like XLA's scatter
    spectra = np.array([[100, 200, 300], [400, 500, 600], [700, 800, 900], [1, 2, 3], [9, 9, 9]], dtype=np.float32)
    idx = np.array([0, 1, 1, 3, 7], dtype=np.int32)
    cube = ops.segment_sum(spectra, idx, 4).cpu().numpy().reshape(2, 2, 3)
    assert np.array_equal(cube, [[[100, 200, 300], [1100, 1300, 1500]], [[0, 0, 0], [1, 2, 3]]])


# ---- fused a1..a5 -----------------------------------------------------------------------------------
def _run_fused(ops, plan, data, edges, S, apply_filter=True):
    coords = ops.dev(data["coords"])
    mass, met, age = ops.dev(data["mass"]).clone(), ops.dev(data["metallicity"]).clone(), ops.dev(data["age"]).clone()
    if apply_filter:
        ops.filter_particles(coords, edges, mass, met, age)
    pix = ops.spaxel_assign(coords, edges)
    cube = ops.build_cube(plan, data["velocity"], mass, met, age, pix, S)
    torch.cuda.synchronize()
    return cube.cpu().numpy()


@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_fused_cube_tng_subset(ops, plans, bc03, muse_wave, tng_subset, method):
    from rubix_b200.synthetic import spatial_edges
    edges = spatial_edges(25)
    tng_subset = _well_conditioned(tng_subset, np.float32(1.1) * bc03["wavelength"], muse_wave)
    out = _run_fused(ops, plans[method], tng_subset, edges, 25)
    ref = c_oracle.particles_to_cube(tng_subset["coords"], tng_subset["velocity"], tng_subset["mass"],
                                     tng_subset["metallicity"], tng_subset["age"], edges, 25, bc03["metallicity"],
                                     bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1, method=method,
                                     dtype=np.float64, n_threads=8)
    assert ref.max() > 0
    _cube_close(out, ref, f"fused tng {method}")


@pytest.mark.parametrize("method", ["linear", "cubic"])
@pytest.mark.parametrize("gen", ["bench_u", "bench_g"])
def test_fused_cube_synthetic(ops, plans, bc03, muse_wave, method, gen):
    from rubix_b200 import synthetic
    edges = synthetic.spatial_edges(25)
    data = _well_conditioned(getattr(synthetic, gen)(20000), np.float32(1.1) * bc03["wavelength"], muse_wave)
    out = _run_fused(ops, plans[method], data, edges, 25)
    ref = c_oracle.particles_to_cube(data["coords"], data["velocity"], data["mass"], data["metallicity"],
                                     data["age"], edges, 25, bc03["metallicity"], bc03["age"], bc03["wavelength"],
                                     bc03["flux"], muse_wave, 0.1, method=method, dtype=np.float64, n_threads=8)
    _cube_close(out, ref, f"fused {gen} {method}")


@pytest.mark.parametrize("env", ["fused_force_lut=1", "fused_force_cas=1", "fused_impl=1",
                                 "psub=64", "small_shift=1", "fused_no_skew=1", "fused_chs=9", "sort_impl=1",
                                 "fused_variant=1", "fused_variant=2", "fused_tr=0"])
@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_fused_cube_alternate_code_paths(ops, plans, bc03, muse_wave, method, env):
    """The general paths -- the group kernel (option fused_impl = 1), its lookup-table channel search for
    non-arange telescope grids and its shared cell region with CAS adds for SSP grids finer than the
    telescope's -- forced on the MUSE configuration (they are otherwise only taken by configurations the
    oracle is slow on); and the warp kernel with other work-item cuts / without the bank skew / larger chunks /
    one warp per cell array (fused_variant = 1) / six arrays with two warps each (2) instead of seven / cells in
    channel order instead of the transposed block layout (fused_tr = 0)."""
    from rubix_b200 import _lib, synthetic
    env, val = env.split("=")
    edges = synthetic.spatial_edges(25)
    data = _well_conditioned(synthetic.bench_g(20000, seed=5), np.float32(1.1) * bc03["wavelength"], muse_wave)
    _lib.set_option(env, int(val))
    try:
        out = _run_fused(ops, plans[method], data, edges, 25)
        err, impl = ops.build_cube_status(plans[method], len(data["mass"]), 25)
    finally:
        _lib.set_option(env, -1)
    assert err == 0
    assert impl == (1 if env in ("fused_force_lut", "fused_force_cas", "fused_impl") else 0)
    if env in ("fused_tr", "fused_variant") and val in ("0", "1"):
        assert not ops.build_cube_cell_layout(plans[method], len(data["mass"]), 25)
    ref = c_oracle.particles_to_cube(data["coords"], data["velocity"], data["mass"], data["metallicity"],
                                     data["age"], edges, 25, bc03["metallicity"], bc03["age"], bc03["wavelength"],
                                     bc03["flux"], muse_wave, 0.1, method=method, dtype=np.float64, n_threads=8)
    _cube_close(out, ref, f"fused {env} {method}")


def test_fused_cube_non_affine_grid(ops, bc03):
    """A telescope grid that is not an arange (slowly growing channel width) takes the lookup-table path."""
    from rubix_b200 import synthetic
    rng = np.random.default_rng(3)
    steps = (1.25 * (1 + 0.2 * np.linspace(0, 1, 1500)) + rng.uniform(0, 0.01, 1500)).astype(np.float32)
    wave = (np.float32(5000.0) + np.cumsum(steps, dtype=np.float64)).astype(np.float32)
    plan = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], wave, 0.05, method="linear")
    edges = synthetic.spatial_edges(9)
    data = _well_conditioned(synthetic.bench_g(6000, seed=9), np.float32(1.05) * bc03["wavelength"], wave)
    out = _run_fused(ops, plan, data, edges, 9)
    ref = c_oracle.particles_to_cube(data["coords"], data["velocity"], data["mass"], data["metallicity"],
                                     data["age"], edges, 9, bc03["metallicity"], bc03["age"], bc03["wavelength"],
                                     bc03["flux"], wave, 0.05, method="linear", dtype=np.float64, n_threads=8)
    _cube_close(out, ref, "fused non-affine grid")


@pytest.mark.parametrize("method", ["linear", "cubic"])
@pytest.mark.parametrize("z,w0,w1,dw", [(0.3, 5000.0, 9000.0, 2.0), (0.0, 3800.0, 7000.0, 0.8), (0.02, 4000.0, 9800.0, 5.0),
                                        (0.1, 4700.15, 9351.4, 30.0)])
def test_fused_cube_other_arange_grids(ops, bc03, method, z, w0, w1, dw):
    """Other redshifts and arange telescope grids: a different knot window, other chunk sizes and cell skews, a
    grid much finer than the SSP's (0.8 A), a coarse one (5 A) and one coarser than the SSP's 20 A spacing
    (30 A: two knots of a lane can share a channel -> the group kernel's CAS mode)."""
    from rubix_b200 import synthetic
    wave = np.arange(w0, w1, dw, dtype=np.float32)
    plan = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], wave, z, method=method)
    edges = synthetic.spatial_edges(7)
    lamz = (np.float32(1.0 + z) * bc03["wavelength"]).astype(np.float32)
    data = _well_conditioned(synthetic.bench_g(4000, seed=23), lamz, wave)
    out = _run_fused(ops, plan, data, edges, 7)
    ref = c_oracle.particles_to_cube(data["coords"], data["velocity"], data["mass"], data["metallicity"],
                                     data["age"], edges, 7, bc03["metallicity"], bc03["age"], bc03["wavelength"],
                                     bc03["flux"], wave, z, method=method, dtype=np.float64, n_threads=8)
    assert ref.max() > 0
    _cube_close(out, ref, f"fused arange grid z={z} [{w0}, {w1}) step {dw} {method}")


def test_fused_cube_large_fov_150(ops, plans, bc03, muse_wave):
    """BASELINE config 4 geometry (150 x 150 spaxels x 3721 channels) at a particle count the oracle
    finishes in seconds: many small spaxel segments, 335 MB cube."""
    from rubix_b200 import synthetic
    edges = synthetic.spatial_edges(150)
    data = _well_conditioned(synthetic.bench_g(60000, seed=21), np.float32(1.1) * bc03["wavelength"], muse_wave)
    out = _run_fused(ops, plans["linear"], data, edges, 150)
    ref = c_oracle.particles_to_cube(data["coords"], data["velocity"], data["mass"], data["metallicity"],
                                     data["age"], edges, 150, bc03["metallicity"], bc03["age"], bc03["wavelength"],
                                     bc03["flux"], muse_wave, 0.1, method="linear", dtype=np.float64, n_threads=8)
    _cube_close(out, ref, "fused 150x150")
    pk, lk = orc.gaussian_kernel_2d(5, 5, 0.6), orc.lsf_kernel(0.5, 1.25)
    conv = ops.psf_lsf(out.astype(np.float32), pk, lk).cpu().numpy()
    sl = slice(40, 110)  # oracle convolution of the central region only (the full 335 MB cube is slow in numpy)
    refc = orc.apply_lsf(orc.apply_psf(ref[sl, sl], pk.astype(np.float64)), 0.5, 1.25)[2:-2, 2:-2]
    assert np.abs(conv[sl, sl][2:-2, 2:-2] - refc).max() <= 5e-6 * np.abs(refc).max()


def test_fused_equals_stage_path(ops, plans, tng_subset):
    """The fused kernel and the materialising stage kernels are two CUDA implementations of the same
    stages; they must agree to float32 rounding."""
    from rubix_b200.synthetic import spatial_edges
    edges = spatial_edges(25)
    plan = plans["cubic"]
    d = tng_subset
    coords = ops.dev(d["coords"])
    mass, met, age = ops.dev(d["mass"]).clone(), ops.dev(d["metallicity"]).clone(), ops.dev(d["age"]).clone()
    ops.filter_particles(coords, edges, mass, met, age)
    pix = ops.spaxel_assign(coords, edges)
    spec = ops.scale_by_mass(ops.ssp_lookup(plan, met, age), mass)
    res = ops.doppler_resample(plan, spec, d["velocity"])
    staged = ops.segment_sum(res, pix, 625).cpu().numpy().reshape(25, 25, -1)
    fused = ops.build_cube(plan, d["velocity"], mass, met, age, pix, 25).cpu().numpy()
    _cube_close(fused, staged, "fused vs staged", rtol_max=1e-5)


def test_fused_edge_cases(ops, plans):
    from rubix_b200.synthetic import spatial_edges, bench_u
    plan = plans["linear"]
    # empty input -> zero cube
    z = np.zeros((0,), dtype=np.float32)
    cube = ops.build_cube(plan, np.zeros((0, 3), dtype=np.float32), z, z, z, np.zeros((0,), dtype=np.int32), 25)
    assert cube.shape == (25, 25, 3721) and float(cube.abs().max()) == 0.0
    # all particles invalid (mass 0 / Z out of grid / pixel out of range) -> exactly zero
    d = bench_u(1000)
    pix = np.full(1000, 3, dtype=np.int32)
    assert float(ops.build_cube(plan, d["velocity"], np.zeros(1000, np.float32), d["metallicity"], d["age"], pix, 25).abs().max()) == 0.0
    assert float(ops.build_cube(plan, d["velocity"], d["mass"], np.full(1000, 0.1, np.float32), d["age"], pix, 25).abs().max()) == 0.0
    assert float(ops.build_cube(plan, d["velocity"], d["mass"], d["metallicity"], d["age"], np.full(1000, 625, np.int32), 25).abs().max()) == 0.0
    # one particle
    one = ops.build_cube(plan, d["velocity"][:1], d["mass"][:1], d["metallicity"][:1], d["age"][:1], pix[:1], 25)
    rows = one.abs().amax(-1).reshape(-1)
    assert float(rows[3]) > 0 and int((rows != 0).sum()) == 1
    # a single crowded spaxel (exercises segment splitting + the two-level reduction), run twice:
    # deterministic bit-for-bit
    d = bench_u(30000)
    pix = np.full(30000, 7, dtype=np.int32)
    a = ops.build_cube(plan, d["velocity"], d["mass"], d["metallicity"], d["age"], pix, 25).cpu().numpy()
    b = ops.build_cube(plan, d["velocity"], d["mass"], d["metallicity"], d["age"], pix, 25).cpu().numpy()
    assert np.array_equal(a, b)
    assert a[0, 7].min() > 0 and np.count_nonzero(np.abs(a).max(-1)) == 1


def test_fused_linearity_and_sharding(ops, plans):
    """Size-independent properties: the cube is a sum over particles (shards add up) and scales
    linearly with mass."""
    from rubix_b200.synthetic import spatial_edges, bench_g
    plan = plans["linear"]
    edges = spatial_edges(25)
    d = bench_g(200000)
    pix = ops.spaxel_assign(d["coords"], edges)
    args = lambda sl: (d["velocity"][sl], d["mass"][sl], d["metallicity"][sl], d["age"][sl], pix[sl])
    full = ops.build_cube(plan, *args(slice(None)), 25).double()
    parts = sum(ops.build_cube(plan, *args(slice(k, None, 4)), 25).double() for k in range(4))
    assert float((full - parts).abs().max()) <= 2e-6 * float(full.abs().max())
    v, m, z, a, p = args(slice(None))
    twice = ops.build_cube(plan, v, 2 * m, z, a, p, 25).double()
    assert float((twice - 2 * full).abs().max()) <= 1e-6 * float(full.abs().max())


# ---- a6 / a7 ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("host_taps", [True, False])
def test_psf_matches_oracle(ops, host_taps):
    rng = np.random.default_rng(6)
    cube = rng.random((25, 25, 300)).astype(np.float32)
    for shape in ((5, 5), (3, 3), (4, 4), (3, 5)):
        k = rng.random(shape).astype(np.float32)
        out = ops.convolve_psf(cube, k, host_taps=host_taps).cpu().numpy()
        ref = orc.apply_psf(cube.astype(np.float64), k.astype(np.float64))
        assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    # tests/test_telescope_psf.py:22-54
    c = np.zeros((10, 10, 3), dtype=np.float32)
    c[5, 5, :] = 1
    out = ops.convolve_psf(c, np.ones((3, 3), dtype=np.float32), host_taps=host_taps).cpu().numpy()
    assert np.array_equal(out, orc.apply_psf(c, np.ones((3, 3), dtype=np.float32)))
    # separable (outer-product) kernels: the host-tap path runs them as two 1-D passes
    for P in (3, 5, 7):
        k = orc.gaussian_kernel_2d(P, P, 0.9)
        out = ops.convolve_psf(cube, k, host_taps=host_taps).cpu().numpy()
        ref = orc.apply_psf(cube.astype(np.float64), k.astype(np.float64))
        assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("host_taps", [True, False])
def test_lsf_matches_oracle(ops, host_taps):
    # tests/test_telescope_lsf.py:6-60 (delta -> normalised gaussian, atol 1e-5)
    for pos in (20, 50, 75):
        s = np.zeros((1, 1, 100), dtype=np.float32)
        s[0, 0, pos] = 1
        k = orc.lsf_kernel(2.0, 1.0)
        out = ops.convolve_lsf(s, k, host_taps=host_taps).cpu().numpy()[0, 0]
        x = np.arange(100)
        g = np.exp(-0.5 * ((x - pos) ** 2) / 4.0)
        assert np.allclose(out, g / g.sum(), atol=1e-5)
    rng = np.random.default_rng(7)
    cube = rng.random((7, 9, 500)).astype(np.float32)
    k = orc.lsf_kernel(0.5, 1.25)
    out = ops.convolve_lsf(cube, k, host_taps=host_taps).cpu().numpy()
    ref = orc.apply_lsf(cube.astype(np.float64), 0.5, 1.25)
    assert out.shape == cube.shape
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    # wide LSF (sigma 3 A: all 25 taps matter) and a medium one (13-tap window)
    for sigma in (3.0, 1.2):
        k = orc.lsf_kernel(sigma, 1.25)
        out = ops.convolve_lsf(cube, k, host_taps=host_taps).cpu().numpy()
        ref = orc.apply_lsf(cube.astype(np.float64), sigma, 1.25)
        assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("host_taps", [True, False])
@pytest.mark.parametrize("S,W,P", [(25, 3721, 5), (13, 517, 5), (31, 130, 5), (12, 700, 3), (17, 300, 7), (9, 200, 4),
                                   (47, 250, 5), (3, 40, 3)])
def test_fused_psf_lsf(ops, S, W, P, host_taps):
    """Device taps: P = 3/5/7 take the register-tiled kernel (5x5 and 4x8 spaxel tiles), P = 4 the generic
    one.  Host taps: the marching kernel (separable PSF, pruned LSF; TX = 5 below 40 columns, 10 above)."""
    rng = np.random.default_rng(8)
    cube = rng.random((S, S, W)).astype(np.float32)
    pk = orc.gaussian_kernel_2d(P, P, 0.6) if P != 4 else rng.random((4, 4)).astype(np.float32)
    lk = orc.lsf_kernel(0.5, 1.25)
    out = ops.psf_lsf(cube, pk, lk, host_taps=host_taps).cpu().numpy()
    ref = orc.apply_lsf(orc.apply_psf(cube.astype(np.float64), pk.astype(np.float64)), 0.5, 1.25)
    err = np.abs(out - ref).max()
    print(f"[psf+lsf {S}x{W} P={P} host_taps={host_taps}] max|d|={err:.3e}")
    assert err <= 2e-6 * np.abs(ref).max()
    two = ops.convolve_lsf(ops.convolve_psf(cube, pk, host_taps=host_taps), lk, host_taps=host_taps).cpu().numpy()
    assert np.abs(two - ref).max() <= 2e-6 * np.abs(ref).max()


def test_psf_lsf_taps_wide_cube(ops):
    """150 x 150 spaxels (config 4's cube shape, fewer channels): several y segments and x strips, the
    spectral tile boundary, a non-square cube; the host-tap kernel against the device-tap kernel and the
    float64 oracle.  A delta cube checks every tap lands where the reference puts it."""
    rng = np.random.default_rng(12)
    pk, lk = orc.gaussian_kernel_2d(5, 5, 0.6), orc.lsf_kernel(0.5, 1.25)
    for shape in ((150, 150, 260), (37, 150, 131)):
        cube = rng.random(shape).astype(np.float32)
        out = ops.psf_lsf(cube, pk, lk).cpu().numpy()
        dev_taps = ops.psf_lsf(cube, pk, lk, host_taps=False).cpu().numpy()
        ref = orc.apply_lsf(orc.apply_psf(cube.astype(np.float64), pk.astype(np.float64)), 0.5, 1.25)
        assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
        assert np.abs(out - dev_taps).max() <= 2e-6 * np.abs(ref).max()
    d = np.zeros((150, 150, 130), dtype=np.float32)
    d[0, 0, 0] = d[149, 149, 129] = d[74, 9, 121] = d[49, 10, 122] = 1.0
    out = ops.psf_lsf(d, pk, lk).cpu().numpy()
    ref = orc.apply_lsf(orc.apply_psf(d.astype(np.float64), pk.astype(np.float64)), 0.5, 1.25)
    assert np.abs(out - ref).max() <= 2e-7


def test_fused_psf_lsf_asymmetric_kernels(ops):
    """Non-symmetric taps: catches flipped or transposed kernels (a Gaussian cannot)."""
    rng = np.random.default_rng(11)
    cube = rng.random((10, 15, 333)).astype(np.float32)
    pk = rng.random((5, 5)).astype(np.float32)
    lk = rng.random(25).astype(np.float32)
    out = ops.psf_lsf(cube, pk, lk).cpu().numpy()   # not an outer product: falls through to the device taps
    mid = orc.apply_psf(cube.astype(np.float64), pk.astype(np.float64))
    ref = np.stack([[np.convolve(mid[y, x], lk.astype(np.float64), mode="full")[12:12 + 333] for x in range(15)]
                    for y in range(10)])
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    # asymmetric outer-product PSF + asymmetric 7-tap LSF: the marching kernel's tap orientation
    pk = np.outer(rng.random(5) + 0.1, rng.random(5) + 0.1).astype(np.float32)
    lk = np.zeros(25, dtype=np.float32)
    lk[9:16] = rng.random(7).astype(np.float32) + 0.1
    out = ops.psf_lsf(cube, pk, lk).cpu().numpy()
    mid = orc.apply_psf(cube.astype(np.float64), pk.astype(np.float64))
    ref = np.stack([[np.convolve(mid[y, x], lk.astype(np.float64), mode="full")[12:12 + 333] for x in range(15)]
                    for y in range(10)])
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.abs(ops.psf_lsf(cube, pk, lk, host_taps=False).cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()


def test_psf_lsf_wavelength_slabs(ops):
    """The multi-GPU PSF / LSF stage: every rank convolves its wavelength slab (read in place, +-12 channel
    halo) of the summed cube; the slabs put together equal the convolution of the whole cube."""
    from rubix_b200 import parallel
    rng = np.random.default_rng(13)
    pk, lk = orc.gaussian_kernel_2d(5, 5, 0.6), orc.lsf_kernel(0.5, 1.25)
    for shape, world in (((25, 25, 1001), 3), ((50, 40, 517), 8)):
        cube = rng.random(shape).astype(np.float32)
        whole = ops.psf_lsf(cube, pk, lk).cpu().numpy()
        dcube = ops.dev(cube)
        parts = []
        for r in range(world):
            lo, hi = parallel.wavelength_slab(shape[2], r, world)
            parts.append(ops.psf_lsf_slab(dcube, lo, hi, pk, lk).cpu().numpy())
        got = np.concatenate(parts, axis=2)
        assert got.shape == whole.shape
        assert np.abs(got - whole).max() <= 1e-6 * np.abs(whole).max()
        ref = orc.apply_lsf(orc.apply_psf(cube.astype(np.float64), pk.astype(np.float64)), 0.5, 1.25)
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("apply_filter", [True, False])
def test_assign_build_cube_equals_two_calls(ops, plans, apply_filter):
    """rbx_assign_build_cube computes the spaxel index inside the build's first kernel: same pixels (bit-exact
    integer work) and a bit-identical cube compared with spaxel assignment followed by rbx_build_cube."""
    from rubix_b200 import synthetic
    edges = synthetic.spatial_edges(25)
    d = synthetic.bench_g(50000, seed=17, scale=1.6)   # ~1 % of the particles outside the aperture
    pix = ops.filter_and_assign(d["coords"], edges) if apply_filter else ops.spaxel_assign(d["coords"], edges)
    two = ops.build_cube(plans["linear"], d["velocity"], d["mass"], d["metallicity"], d["age"], pix, 25)
    one, pix1 = ops.assign_build_cube(plans["linear"], d["coords"], edges, d["velocity"], d["mass"], d["metallicity"],
                                      d["age"], 25, apply_filter=apply_filter, return_pixel=True)
    assert torch.equal(pix1, pix)
    assert (pix < 0).any() == apply_filter
    assert torch.equal(one, two)


def test_pipeline_host_particle_ranges(ops, plans, bc03, muse_wave):
    """rbx_pipeline_host bins a galaxy in particle ranges (copy / compute overlap, accumulating cube build):
    1, 2 and 5 ranges give the same cube, equal to the oracle's."""
    from rubix_b200 import synthetic
    edges = synthetic.spatial_edges(25)
    d = _well_conditioned(synthetic.bench_g(30011, seed=9), np.float32(1.1) * bc03["wavelength"], muse_wave)
    pk, lk = orc.gaussian_kernel_2d(5, 5, 0.6), orc.lsf_kernel(0.5, 1.25)
    outs = {}
    from rubix_b200 import _lib
    try:
        for c in (1, 2, 5):
            _lib.set_option("host_chunks", c)
            outs[c] = ops.pipeline_host(plans["linear"], d["coords"], d["velocity"], d["mass"], d["metallicity"],
                                        d["age"], edges, 25, pk, lk).copy()
    finally:
        _lib.set_option("host_chunks", -1)
    ref = c_oracle.particles_to_cube(d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"], edges, 25,
                                     bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave,
                                     0.1, method="linear", dtype=np.float64, n_threads=8)
    ref = orc.apply_lsf(orc.apply_psf(ref, pk.astype(np.float64)), 0.5, 1.25)
    for c, out in outs.items():
        _cube_close(out, ref, f"pipeline_host {c} ranges")
    assert np.abs(outs[2] - outs[1]).max() <= 2e-6 * np.abs(ref).max()
    assert np.abs(outs[5] - outs[1]).max() <= 2e-6 * np.abs(ref).max()


def test_gaussian_kernels_on_device(ops):
    from rubix_b200 import _lib
    import ctypes as C
    k = torch.empty(25, device="cuda")
    _lib.check(_lib.lib().rbx_gaussian_psf_kernel(5, 5, 0.6, C.c_void_p(k.data_ptr()), None))
    assert np.allclose(k.cpu().numpy().reshape(5, 5), orc.gaussian_kernel_2d(5, 5, 0.6), rtol=1e-6, atol=1e-9)
    _lib.check(_lib.lib().rbx_gaussian_lsf_kernel(0.5, 1.25, 12, C.c_void_p(k.data_ptr()), None))
    assert np.allclose(k.cpu().numpy(), orc.lsf_kernel(0.5, 1.25), rtol=1e-5, atol=1e-12)


# ---- host-buffer entry point ----------------------------------------------------------------------
def test_pipeline_host_end_to_end(ops, plans, bc03, muse_wave, tng_subset):
    from rubix_b200.synthetic import spatial_edges
    edges = spatial_edges(25)
    pk = orc.gaussian_kernel_2d(5, 5, 0.6)
    lk = orc.lsf_kernel(0.5, 1.25)
    d = _well_conditioned(tng_subset, np.float32(1.1) * bc03["wavelength"], muse_wave)
    out = ops.pipeline_host(plans["cubic"], d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"],
                            edges, 25, pk, lk)
    ref = c_oracle.particles_to_cube(d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"], edges, 25,
                                     bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave,
                                     0.1, method="cubic", dtype=np.float64, n_threads=8)
    ref = orc.apply_lsf(orc.apply_psf(ref, pk.astype(np.float64)), 0.5, 1.25)
    _cube_close(out, ref, "pipeline_host cubic + psf + lsf")


# ---- rotate_galaxy ---------------------------------------------------------------------
@pytest.mark.parametrize("n,angles", [(3, (90.0, 0.0, 0.0)), (5000, (0.0, 0.0, 0.0)), (200000, (33.0, 71.0, 152.0))])
def test_rotate_galaxy_matches_oracle(ops, n, angles):
    """Inertia tensor (with the reference's index-0 padding), eigenvector alignment, Euler rotation against
    the float64 oracle.  tests/test_galaxy_alignment.py:160-190 is the n = 3 case."""
    rng = np.random.default_rng(21)
    if n == 3:
        pos, vel, m, r = np.eye(3, dtype=np.float32), np.roll(np.eye(3, dtype=np.float32), 1, axis=1), np.ones(3, np.float32), 2.0
    else:
        # a flattened, tilted disc: distinct principal axes
        pos = (rng.normal(size=(n, 3)) * np.array([4.0, 2.5, 0.6])).astype(np.float32)
        tilt = orc.euler_rotation_matrix(25.0, -40.0, 10.0).astype(np.float32)
        pos = (pos @ tilt).astype(np.float32)
        vel = rng.normal(scale=150.0, size=(n, 3)).astype(np.float32)
        m = (rng.random(n) + 0.5).astype(np.float32)
        r = 3.0
    c, v, R = ops.rotate_galaxy(pos, vel, m, r, *angles)
    c, v, R = c.cpu().numpy(), v.cpu().numpy(), R.cpu().numpy()
    pref, vref, Rref = orc.rotate_galaxy(pos, vel, m, r, *angles)
    if n == 3:  # degenerate tensor: any orthonormal basis is valid; check the rotation is one and was applied
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-6)
        E = orc.euler_rotation_matrix(*angles)
        assert np.allclose(c, (pos.astype(np.float64) @ R) @ E, atol=1e-6)
        return
    assert np.abs(R - Rref).max() <= 2e-6
    assert np.abs(c - pref).max() <= 1e-5 * np.abs(pref).max()
    assert np.abs(v - vref).max() <= 1e-5 * np.abs(vref).max()
    # bit-identical reruns (fixed reduction order)
    c2, v2, R2 = ops.rotate_galaxy(pos, vel, m, r, *angles)
    assert torch.equal(c2, torch.from_numpy(c).cuda()) and torch.equal(R2.cpu(), torch.from_numpy(R))


def test_rotate_then_bin(ops, plans, bc03, muse_wave):
    """rotate_galaxy -> filter -> spaxel assignment -> cube through the factories' operators: an edge-on
    rotation of a thin disc puts the flux into a band of spaxel rows."""
    from rubix_b200.synthetic import spatial_edges
    rng = np.random.default_rng(22)
    n = 20000
    pos = (rng.normal(size=(n, 3)) * np.array([1.5, 1.5, 0.05])).astype(np.float32)
    vel = rng.normal(scale=50.0, size=(n, 3)).astype(np.float32)
    m = np.ones(n, np.float32)
    met = np.full(n, 0.02, np.float32)
    age = np.full(n, 9.0, np.float32)
    edges = spatial_edges(25)
    cubes = {}
    for name, ang in (("face", (0.0, 0.0, 0.0)), ("edge", (90.0, 0.0, 0.0))):
        c, v, _ = ops.rotate_galaxy(pos, vel, m, 2.0, *ang)
        pix = ops.filter_and_assign(c, edges)
        cubes[name] = ops.build_cube(plans["linear"], v, m, met, age, pix, 25).sum(dim=2).cpu().numpy()
    # the principal-axis frame orders axes by ascending moment: x, y span the disc for "face-on" output
    rows_face = (cubes["face"].sum(axis=1) > 0).sum() + (cubes["face"].sum(axis=0) > 0).sum()
    rows_edge = min((cubes["edge"].sum(axis=1) > 0).sum(), (cubes["edge"].sum(axis=0) > 0).sum())
    assert rows_edge <= 3 and rows_face >= 20


# ---- apply_noise -----------------------------------------------------------------------
def test_noise_stream_matches_oracle(ops):
    """The counter-based random words are integer work: bit-exact against the oracle's threefry; the float
    mapping (uniform, sqrt(2) erfinv) to 1e-6."""
    n = 100003
    for dist in ("normal", "uniform"):
        smp, bits = ops.noise_samples(n, dist)
        assert np.array_equal(bits.cpu().numpy().view(np.uint32), orc.random_bits((0, 0), n))
        ref = orc.sample_noise(n, dist)
        # float32 evaluation of w = -log1p(-u^2) and the erfinv polynomial against the oracle's float64 one: the
        # cancellation in 1 - u^2 costs a few 1e-6 relative in the tails (|N| > 3)
        assert (np.abs(smp.cpu().numpy() - ref) <= 1e-5 * np.maximum(1.0, np.abs(ref))).all()
    smp7, bits7 = ops.noise_samples(1000, "normal", key=(7, 11))
    assert np.array_equal(bits7.cpu().numpy().view(np.uint32), orc.random_bits((7, 11), 1000))


@pytest.mark.parametrize("shape", [(25, 25, 3721), (7, 9, 130)])
@pytest.mark.parametrize("dist", ["normal", "uniform"])
def test_apply_noise_matches_oracle(ops, shape, dist):
    rng = np.random.default_rng(31)
    cube = (rng.random(shape) * 3 + 0.2).astype(np.float32)
    out = ops.apply_noise(cube, 12.5, dist).cpu().numpy()
    ref = orc.apply_noise(cube, 12.5, dist)
    assert np.abs(out - ref).max() <= 5e-6 * np.abs(ref).max()
    assert np.abs(out - cube).max() > 0
    # a spaxel without flux switches the noise off altogether (the reference's NaN-propagating median)
    cube[1, 2] = 0.0
    assert np.array_equal(ops.apply_noise(cube, 12.5, dist).cpu().numpy(), cube)
