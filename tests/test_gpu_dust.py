"""GPU parity tests of the dust variant (calc_dusty_ifu) through the C ABI: per-star A_V against the
oracle's spaxel-by-spaxel restatement of apply_spaxel_extinction, the extinction factor, the one-pass
resample + extinction + cube kernel against the staged kernels and the float64 oracle, and the dusty pipeline."""

import copy

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import rubix_oracle as orc  # noqa: E402
from rubix_b200 import dust as hdust  # noqa: E402
from test_dust_cpu import CONFIG  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rubix_b200 import ops as _ops
    return _ops


def _gas(rng, ng, S, zmean=0.0):
    z = rng.normal(zmean, 1.0, ng).astype(np.float32)
    coords = np.stack([rng.normal(0, 1, ng), rng.normal(0, 1, ng), z], 1).astype(np.float32)
    pix = rng.integers(0, S, ng).astype(np.int32)
    mass = rng.uniform(0.5, 2.0, ng).astype(np.float32) * 1e5
    metals = rng.uniform(1e-4, 1e-2, (ng, 9)).astype(np.float32)
    metals[:, 0] = rng.uniform(0.70, 0.76, ng)          # hydrogen
    metals[:, 4] = rng.uniform(5e-4, 1.2e-2, ng)        # oxygen: 12 + log(O/H) on both sides of the break
    return coords, pix, mass, metals


def _stars(rng, ns, S, spread=1.6):
    coords = np.stack([rng.normal(0, 1, ns), rng.normal(0, 1, ns), rng.normal(0, spread, ns)], 1).astype(np.float32)
    return coords, rng.integers(0, S, ns).astype(np.int32)


def _av_oracle(gc, gp, gm, gmet, sc, sp, S, model, xco, rho, area):
    cell = orc.dust_cell_extinction(gm, gmet, rho, area, model, xco)
    return orc.stars_av(gc[:, 2], gp, cell, sc[:, 2], sp, S), cell


@pytest.mark.parametrize("model,xco", [("broken power law fit", "Z"), ("power law slope free", "MW")])
def test_dust_av_matches_oracle(ops, model, xco):
    rng = np.random.default_rng(42)
    S = 36
    gc, gp, gm, gmet = _gas(rng, 3000, S)
    gp[gp == 7] = 8          # a spaxel without gas
    gp[:1] = 9               # ... and re-point a few cells so that spaxels 10 / 11 hold one and two cells
    gp[gp == 10] = 12
    gp[gp == 11] = 12
    gp[1] = 10
    gp[2:4] = 11
    sc, sp = _stars(rng, 5000, S)
    sp[:40] = 7
    sp[40:80] = 10
    sp[80:120] = 11
    sp[120:125] = -1         # outside every spaxel: A_V = 0
    area = 0.145
    dtg = hdust.dust_to_gas_parameters(model, xco)
    av, cells = ops.dust_av(gc, gp, gm, gmet, sc, sp, S, dtg, hdust.extinction_constant(3.5), area, return_cells=True)
    torch.cuda.synchronize()
    ref, cell_ref = _av_oracle(gc, gp, gm, gmet, sc, sp, S, model, xco, 3.5, area)
    cells = cells.cpu().numpy().astype(np.float64)
    # float32 exponent a + alpha (8.69 - x) carries ~1e-6 absolute error -> ~3e-6 relative in 10^x
    assert np.abs(cells - cell_ref).max() <= 1e-5 * np.abs(cell_ref).max()
    av = av.cpu().numpy().astype(np.float64)
    assert np.all(av[sp == 7] == 0) and np.all(av[sp < 0] == 0)
    assert np.abs(av - ref).max() <= 1e-5 * np.abs(ref).max(), np.abs(av - ref).max() / np.abs(ref).max()
    assert ref.max() > 0


@pytest.mark.parametrize("zmean", [-50.0, 50.0])
def test_dust_av_one_sided_gas(ops, zmean):
    """All gas in front of (behind) the galaxy: no far cell on one side, so the table of a spaxel ends with its own
    cells -- left extrapolation through the first pair, right end value."""
    rng = np.random.default_rng(7)
    S = 9
    gc, gp, gm, gmet = _gas(rng, 600, S, zmean)
    sc, sp = _stars(rng, 900, S, spread=80.0)
    dtg = hdust.dust_to_gas_parameters("broken power law fit", "Z")
    av = ops.dust_av(gc, gp, gm, gmet, sc, sp, S, dtg, hdust.extinction_constant(3.5), 0.1).cpu().numpy().astype(np.float64)
    ref, _ = _av_oracle(gc, gp, gm, gmet, sc, sp, S, "broken power law fit", "Z", 3.5, 0.1)
    scale = np.abs(ref).max()
    assert np.abs(av - ref).max() <= 2e-5 * scale, np.abs(av - ref).max() / scale


def test_dust_av_single_spaxel_and_no_gas(ops):
    rng = np.random.default_rng(3)
    gc, gp, gm, gmet = _gas(rng, 50, 1)
    sc, sp = _stars(rng, 200, 1, spread=3.0)
    dtg = hdust.dust_to_gas_parameters("broken power law fit", "Z")
    av = ops.dust_av(gc, gp, gm, gmet, sc, sp, 1, dtg, hdust.extinction_constant(3.5), 0.1).cpu().numpy().astype(np.float64)
    ref, _ = _av_oracle(gc, gp, gm, gmet, sc, sp, 1, "broken power law fit", "Z", 3.5, 0.1)
    assert np.abs(av - ref).max() <= 2e-5 * np.abs(ref).max()
    av0 = ops.dust_av(gc[:0], gp[:0], gm[:0], gmet[:0], sc, sp, 1, dtg, 1.0, 0.1)
    assert torch.all(av0 == 0)


@pytest.mark.parametrize("model", ["Cardelli89", "Gordon23"])
def test_apply_extinction_matches_oracle(ops, model, muse_wave):
    rng = np.random.default_rng(5)
    n = 64
    spec = rng.uniform(0, 10, (n, len(muse_wave))).astype(np.float32)
    av = rng.uniform(0, 3, n).astype(np.float32)
    av[:3] = 0.0
    axav = hdust.extinction_curve(model, muse_wave, 3.1)
    out = ops.apply_extinction(spec, av, axav).cpu().numpy()
    ref = spec.astype(np.float64) * orc.extinguish(muse_wave, av, model, 3.1)
    assert np.array_equal(out[:3], spec[:3])          # A_V = 0: factor exactly 1
    assert np.abs(out - ref).max() <= 4e-6 * np.abs(ref).max()
    assert np.all(out <= spec)


def _cube_inputs(tng_subset, bc03, muse_wave, n):
    from helpers import well_conditioned
    d = well_conditioned(tng_subset, np.float32(1.1) * bc03["wavelength"], muse_wave)
    return {k: v[:n].copy() for k, v in d.items()}


@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_build_cube_dusty_matches_oracle_and_staged(ops, bc03, muse_wave, tng_subset, method):
    from helpers import cube_close
    from rubix_b200.core.telescope import get_spatial_bin_edges
    d = _cube_inputs(tng_subset, bc03, muse_wave, 1500)
    n = len(d["mass"])
    edges = get_spatial_bin_edges(CONFIG)
    rng = np.random.default_rng(11)
    av = rng.uniform(0, 2.5, n).astype(np.float32)
    axav = hdust.extinction_curve("Cardelli89", muse_wave, 3.1)
    plan = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1, method=method)
    pix = ops.spaxel_assign(d["coords"], edges)
    spec = ops.scale_by_mass(ops.ssp_lookup(plan, d["metallicity"], d["age"]), d["mass"])
    # one pass
    cube = ops.build_cube_dusty(plan, spec, d["velocity"], pix, 25, av, axav)
    # staged: resample -> extinction -> segment sum
    res = ops.doppler_resample(plan, spec, d["velocity"])
    staged = ops.segment_sum(ops.apply_extinction(res, av, axav), pix, 625).reshape(25, 25, -1)
    ext = orc.extinguish(muse_wave, av, "Cardelli89", 3.1)
    ref, _ = orc.particles_to_cube(d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"], edges, 25,
                                   bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1,
                                   method=method, dtype=np.float64, apply_filter=False, extinction=ext)
    out = cube.cpu().numpy()
    cube_close(out, ref, f"dusty one-pass {method}")
    cube_close(staged.cpu().numpy(), ref, f"dusty staged {method}")
    assert np.abs(out.astype(np.float64) - ref).max() <= 1e-5 * np.abs(ref).sum()
    # without dust the kernel is doppler_shift_and_resampling + calculate_datacube
    nodust = ops.build_cube_dusty(plan, spec, d["velocity"], pix, 25).cpu().numpy()
    ref0, _ = orc.particles_to_cube(d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"], edges, 25,
                                    bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1,
                                    method=method, dtype=np.float64, apply_filter=False)
    cube_close(nodust, ref0, f"one-pass without dust {method}")
    assert np.all(out <= nodust * (1 + 1e-5) + 1e-30)


@pytest.mark.parametrize("method", ["linear", "cubic"])
@pytest.mark.parametrize("model", ["Cardelli89", "Gordon23"])
def test_build_cube_dusty_binned_matches_oracle(ops, bc03, muse_wave, tng_subset, method, model):
    """The binned-moment form (A_V bins, third-order expansion inside a bin, four runs of the knot-based cube
    kernel) against the float64 oracle and the one-pass kernel."""
    from helpers import cube_close
    from rubix_b200.core.telescope import get_spatial_bin_edges
    d = _cube_inputs(tng_subset, bc03, muse_wave, 1500)
    n = len(d["mass"])
    edges = get_spatial_bin_edges(CONFIG)
    rng = np.random.default_rng(13)
    av = rng.uniform(-0.2, 2.5, n).astype(np.float32)     # left extrapolation can make A_V slightly negative
    av[:7] = 0.0
    axav = hdust.extinction_curve(model, muse_wave, 3.1)
    plan = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1, method=method)
    pix = ops.spaxel_assign(d["coords"], edges)
    cube = ops.build_cube_dusty_binned(plan, d["velocity"], d["mass"], d["metallicity"], d["age"], pix, 25, av, axav)
    assert cube is not None
    ext = orc.extinguish(muse_wave, av, model, 3.1)
    ref, _ = orc.particles_to_cube(d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"], edges, 25,
                                   bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1,
                                   method=method, dtype=np.float64, apply_filter=False, extinction=ext)
    cube_close(cube.cpu().numpy(), ref, f"dusty binned {method} {model}")
    spec = ops.scale_by_mass(ops.ssp_lookup(plan, d["metallicity"], d["age"]), d["mass"])
    one = ops.build_cube_dusty(plan, spec, d["velocity"], pix, 25, av, axav)
    cube_close(cube.cpu().numpy(), one.cpu().numpy().astype(np.float64), "binned vs one-pass", rtol_max=1e-5)
    # a single A_V value: one bin, the expansion is exact
    av1 = np.full(n, 0.7, np.float32)
    c1 = ops.build_cube_dusty_binned(plan, d["velocity"], d["mass"], d["metallicity"], d["age"], pix, 25, av1, axav)
    c0 = ops.build_cube(plan, d["velocity"], d["mass"], d["metallicity"], d["age"], pix, 25).cpu().numpy().astype(np.float64)
    want = c0 * np.power(10.0, -0.4 * 0.7 * axav.astype(np.float64))[None, None, :]
    cube_close(c1.cpu().numpy(), want, "binned, single A_V", rtol_max=4e-6)
    # non-finite A_V: the caller is sent to the one-pass kernel
    avn = av.copy()
    avn[3] = np.nan
    assert ops.build_cube_dusty_binned(plan, d["velocity"], d["mass"], d["metallicity"], d["age"], pix, 25, avn, axav) is None


def test_build_cube_dusty_edge_cases(ops, bc03, muse_wave, tng_subset):
    d = _cube_inputs(tng_subset, bc03, muse_wave, 200)
    plan = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1, method="linear")
    spec = ops.scale_by_mass(ops.ssp_lookup(plan, d["metallicity"], d["age"]), d["mass"])
    axav = hdust.extinction_curve("Gordon23", muse_wave, 3.1)
    # empty input -> zero cube
    z = ops.build_cube_dusty(plan, spec[:0], d["velocity"][:0], np.zeros(0, np.int32), 25, np.zeros(0, np.float32), axav)
    assert tuple(z.shape) == (25, 25, 3721) and torch.all(z == 0)
    # ids outside [0, S^2) are dropped (jax.ops.segment_sum); one crowded spaxel; linear in exp(-A_V)
    pix = np.full(200, 312, np.int32)
    pix[:10] = -1
    pix[10:20] = 625
    av = np.zeros(200, np.float32)
    a = ops.build_cube_dusty(plan, spec, d["velocity"], pix, 25, av, axav).cpu().numpy()
    assert np.count_nonzero(a.reshape(625, -1).sum(1)) == 1
    b = ops.build_cube_dusty(plan, spec[20:], d["velocity"][20:], pix[20:], 25, av[20:], axav).cpu().numpy()
    assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()
    c = ops.build_cube_dusty(plan, spec, d["velocity"], pix, 25, av + 1.0, axav).cpu().numpy()
    want = a.reshape(625, -1)[312].astype(np.float64) * np.power(10.0, -0.4 * axav.astype(np.float64))
    assert np.abs(c.reshape(625, -1)[312] - want).max() <= 4e-6 * want.max()


def _dusty_data(core, tng_subset, bc03, muse_wave, n=1200, ng=4000):
    d = _cube_inputs(tng_subset, bc03, muse_wave, n)
    rng = np.random.default_rng(42)
    rd = core.make_rubix_data(**d, device=False)
    gc = np.stack([rng.normal(0, 2.0, ng), rng.normal(0, 2.0, ng), rng.normal(0, 1.0, ng)], 1).astype(np.float32)
    gc[:50] *= 10.0   # some gas outside the aperture: masked by filter_particles (mass -> 0), metals kept
    metals = rng.uniform(1e-4, 1e-2, (ng, 9)).astype(np.float32)
    metals[:, 0] = 0.74
    metals[:, 4] = rng.uniform(2e-3, 1.2e-2, ng)
    rd.gas.coords, rd.gas.velocity = gc, np.zeros_like(gc)
    rd.gas.mass = (rng.uniform(0.5, 2.0, ng) * 3e5).astype(np.float32)
    rd.gas.metals = metals
    return d, rd


def test_dusty_pipeline_fused_equals_staged_and_oracle(bc03, muse_wave, tng_subset, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from helpers import cube_close
    from rubix_b200 import core
    from rubix_b200.core.telescope import get_spatial_bin_edges
    from rubix_b200.telescope import calculate_spatial_bin_edges
    from rubix_b200.cosmology import get_cosmology
    outs, avs = {}, {}
    for fused in (False, True, "one-pass"):
        cfg = copy.deepcopy(CONFIG)
        cfg["b200"] = {"fused": bool(fused)}
        monkeypatch.setenv("RBX_DUSTY_IMPL", "onepass" if fused == "one-pass" else "binned")
        cfg["data"] = {"args": {"particle_type": ["stars", "gas"]}}
        d, rd = _dusty_data(core, tng_subset, bc03, muse_wave)
        pipe = core.RubixPipeline(cfg, data=rd)
        names = [fn.__name__ for fn in pipe.assemble()]
        assert names.index("calculate_extinction") == names.index("calculate_datacube") - 1
        out = pipe.run()
        assert tuple(out.stars.datacube.shape) == (25, 25, 3721) and not torch.isnan(out.stars.datacube).any()
        outs[fused] = out.stars.datacube.cpu().numpy()
    cube_close(outs[True], outs[False].astype(np.float64), "dusty pipeline fused (binned) vs staged", rtol_max=1e-5)
    cube_close(outs["one-pass"], outs[False].astype(np.float64), "dusty pipeline fused (one-pass) vs staged", rtol_max=1e-5)
    # oracle, stage by stage on the same inputs
    cfg = copy.deepcopy(CONFIG)
    d, rd = _dusty_data(core, tng_subset, bc03, muse_wave)
    edges = get_spatial_bin_edges(cfg)
    _, size = calculate_spatial_bin_edges(fov=5.0, spatial_bins=25, dist_z=0.1, cosmology=get_cosmology(cfg))
    mass, met, age, _ = orc.filter_particles(d["coords"], d["mass"], d["metallicity"], d["age"], edges)
    gmask = orc.mask_particles_outside_aperture(rd.gas.coords, edges)
    gmass = np.where(gmask, rd.gas.mass, 0).astype(np.float32)
    spix = orc.square_spaxel_assignment(d["coords"], edges)
    gpix = orc.square_spaxel_assignment(rd.gas.coords, edges)
    cell = orc.dust_cell_extinction(gmass, rd.gas.metals, 3.5, float(np.float32(size) ** 2), "broken power law fit", "Z")
    av = orc.stars_av(rd.gas.coords[:, 2], gpix, cell, d["coords"][:, 2], spix, 625)
    assert av.max() > 0.05   # the test galaxy is actually dusty
    ext = orc.extinguish(muse_wave, av, "Cardelli89", 3.1)
    raw, _ = orc.particles_to_cube(d["coords"], d["velocity"], mass.astype(np.float32), met.astype(np.float32),
                                   age.astype(np.float32), edges, 25, bc03["metallicity"], bc03["age"], bc03["wavelength"],
                                   bc03["flux"], muse_wave, 0.1, method="cubic", dtype=np.float64, apply_filter=False,
                                   extinction=ext)
    ref = orc.apply_lsf(orc.apply_psf(raw, orc.gaussian_kernel_2d(5, 5, 0.6).astype(np.float64)), 0.5, 1.25)
    cube_close(outs[True], ref, "dusty pipeline vs oracle", rtol_max=2e-5)
    assert np.abs(outs[True].astype(np.float64) - ref).max() <= 1e-5 * np.abs(ref).sum()


def test_dusty_cube_properties_at_full_size(ops, bc03, muse_wave):
    """BASELINE config 2 size (10^6 particles, full MUSE grid), where the oracle is too slow: properties that do not
    depend on size -- A_V = 0 reproduces the dust-free cube, extinction only removes flux, star shards add up, and
    the binned-moment form agrees with the one-pass kernel."""
    from rubix_b200 import synthetic
    n = 1_000_000
    p = synthetic.bench_g(n)
    edges = synthetic.spatial_edges(25)
    plan = ops.Plan(bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1, method="linear")
    pix = ops.spaxel_assign(p["coords"], edges)
    vel, mass, met, age = (ops.dev(p[k]) for k in ("velocity", "mass", "metallicity", "age"))
    axav = ops.dev(hdust.extinction_curve("Gordon23", muse_wave, 3.1))
    av = ops.dev(np.random.default_rng(42).uniform(0.0, 3.0, n).astype(np.float32))
    clean = ops.build_cube(plan, vel, mass, met, age, pix, 25)
    mx = float(clean.max())
    zero = ops.build_cube_dusty_binned(plan, vel, mass, met, age, pix, 25, torch.zeros_like(av), axav)
    assert float((zero - clean).abs().max()) <= 2e-6 * mx
    dusty = ops.build_cube_dusty_binned(plan, vel, mass, met, age, pix, 25, av, axav)
    assert torch.isfinite(dusty).all() and float(dusty.min()) >= 0.0
    assert bool((dusty <= clean * (1 + 1e-5) + 1e-6 * mx).all()) and float(dusty.sum()) < 0.9 * float(clean.sum())
    h = n // 2
    parts = sum(ops.build_cube_dusty_binned(plan, vel[s], mass[s], met[s], age[s], pix[s], 25, av[s], axav)
                for s in (slice(0, h), slice(h, n)))
    assert float((parts - dusty).abs().max()) <= 5e-6 * float(dusty.max())
    m = 200_000   # the one-pass kernel on a subset (it needs the (m, L) SSP spectra)
    spec = ops.scale_by_mass(ops.ssp_lookup(plan, met[:m], age[:m]), mass[:m])
    one = ops.build_cube_dusty(plan, spec, vel[:m], pix[:m], 25, av[:m], axav)
    binned = ops.build_cube_dusty_binned(plan, vel[:m], mass[:m], met[:m], age[:m], pix[:m], 25, av[:m], axav)
    assert float((one - binned).abs().max()) <= 5e-6 * float(one.max())
