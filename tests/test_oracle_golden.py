"""Pin the CPU oracle against every known-answer test the reference holds for the hot path.

Each test names the reference test it restates (paths relative to the reference tree).
These run on CPU (no GPU marker).
"""

import numpy as np
import pytest

from oracle import rubix_oracle as orc


# ---- rubix/telescope/utils.py -----------------------------------------------------------------
def test_square_spaxel_assignment_kat():
    # tests/test_telescope_utils.py:20-35
    coords = np.array([[0.5, 1.5], [2.5, 3.5]])
    edges = np.array([0, 1, 2, 3, 4])
    out = orc.square_spaxel_assignment(coords, edges)
    assert out.dtype == np.int32
    assert np.array_equal(out, [4, 14])


@pytest.mark.parametrize(
    "coords,expected",
    [
        ([[0.5, 0.5, 0], [0.2, 0.2, 0]], 2),  # test_all_particles_inside :45-51
        ([[1.5, 1.5, 0], [-0.1, -0.1, 0]], 0),  # test_all_particles_outside :54-59
        ([[0, 0, 0], [1, 1, 0], [0, 1, 0], [1, 0, 0]], 4),  # test_particles_on_boundary :62-68
        ([[0.5, 0.5, 0], [1.5, 1.5, 0], [0, 0, 0], [-0.1, -0.1, 0]], 2),  # test_mixed :71-79
    ],
)
def test_mask_particles(coords, expected):
    m = orc.mask_particles_outside_aperture(np.array(coords, dtype=np.float32), np.array([0, 1]))
    assert m.sum() == expected


def test_mask_no_particles():
    # tests/test_telescope_utils.py:38-42
    m = orc.mask_particles_outside_aperture(np.zeros((0, 3)), np.array([0, 1]))
    assert len(m) == 0


# ---- rubix/spectra/ifu.py -----------------------------------------------------------------------
def test_cosmological_doppler_shift():
    # tests/test_spectra_ifu.py:29-34
    w = np.array([5000.0, 6000.0, 7000.0], dtype=np.float32)
    assert np.allclose(orc.cosmological_doppler_shift(0.1, w), w * (1 + 0.1))


def test_calculate_diff():
    # tests/test_spectra_ifu.py:37-50
    assert np.array_equal(orc.calculate_diff(np.array([1.0, 2.0, 4.0, 7.0])), [0, 1, 2, 3])


def test_velocity_doppler_shift():
    # tests/test_spectra_ifu.py:150-172
    w = np.array([5000.0, 6000.0, 7000.0], dtype=np.float32)
    v = np.array([[300.0, 400.0, 500.0], [600.0, 700.0, 800.0]], dtype=np.float32)
    out = orc.velocity_doppler_shift(w, v, direction="y")
    exp = np.stack([w * np.exp(400.0 / 299792.458), w * np.exp(700.0 / 299792.458)])
    assert np.allclose(out, exp, rtol=1e-5)


def test_resample_spectrum_kat():
    # tests/test_spectra_ifu.py:175-199 and the numbers quoted in SURVEY.md section 8(c)
    s = np.array([1.0, 2.0, 3.0, 4.0, 5.0], dtype=np.float32)
    lam = np.array([4000.0, 5000.0, 6000.0, 7000.0, 8000.0], dtype=np.float32)
    t = np.array([4500.0, 5500.0, 6500.0, 7500.0], dtype=np.float32)
    out = orc.resample_spectrum(s, lam, t)
    # by hand: p = [1.5, 2.5, 3.5, 4.5]; total = 2*1000+3*1000+4*1000 = 9000;
    # new = 2.5*1000+3.5*1000+4.5*1000 = 10500; scale = 6/7
    assert np.allclose(out, np.array([1.5, 2.5, 3.5, 4.5]) * 6 / 7, rtol=1e-6)
    assert np.allclose(out, [1.2857143, 2.142857, 3.0, 3.857143], rtol=1e-6)
    assert not np.isnan(out).any()


def test_resample_spectrum_zero():
    # tests/test_spectra_ifu.py:202-228: zero spectrum -> exactly zero, no NaN
    lam = np.array([4000.0, 5000.0, 6000.0, 7000.0, 8000.0], dtype=np.float32)
    t = np.array([4500.0, 5500.0, 6500.0, 7500.0], dtype=np.float32)
    out = orc.resample_spectrum(np.zeros(5, dtype=np.float32), lam, t)
    assert (out == 0).all() and not np.isnan(out).any()


def test_calculate_cube_kat():
    # tests/test_spectra_ifu.py:231-257
    spectra = np.array([[100, 200, 300], [400, 500, 600], [700, 800, 900], [1, 2, 3]], dtype=np.float32)
    idx = np.array([0, 1, 1, 3], dtype=np.int32)
    cube = orc.calculate_cube(spectra, idx, 2)
    exp = np.array([[[100, 200, 300], [1100, 1300, 1500]], [[0, 0, 0], [1, 2, 3]]])
    assert np.array_equal(cube, exp)


def test_scale_by_mass_exact():
    # tests/test_core_ifu.py:249-282 (array_equal on spectra * mass)
    rng = np.random.default_rng(0)
    spec = rng.random((1, 7, 11), dtype=np.float32)
    mass = rng.random((1, 7), dtype=np.float32)
    assert np.array_equal(orc.scale_spectrum_by_mass(spec, mass), spec * mass[..., None])


# ---- interp2d (interpax) ------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_lookup_grid_nodes_identity(bc03, method):
    # tests/test_core_ssp.py:158-173 (all 6x221 nodes, rtol 1e-5 / atol 1e-6), :95-113, test_ssp_grid.py:696-698
    Z, A = np.meshgrid(bc03["metallicity"], bc03["age"], indexing="ij")
    out = orc.interp2d(Z.ravel(), A.ravel(), bc03["metallicity"], bc03["age"], bc03["flux"], method=method)
    exp = bc03["flux"].reshape(-1, bc03["flux"].shape[-1])
    assert np.allclose(out, exp, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_lookup_out_of_grid_is_zero(bc03, method):
    # tests/test_core_ssp.py:116-126,176-181; test_ssp_grid.py:679-685; test_core_ifu.py:218-243 (Z=0.1 > 0.05)
    zq = np.array([0.1, 1e-5, 0.02, 0.02], dtype=np.float32)
    aq = np.array([5.0, 5.0, 11.0, -1.0], dtype=np.float32)
    out = orc.interp2d(zq, aq, bc03["metallicity"], bc03["age"], bc03["flux"], method=method)
    assert (out == 0).all()
    # boundaries are inclusive
    out = orc.interp2d(bc03["metallicity"][[0, -1]], bc03["age"][[0, -1]], bc03["metallicity"],
                       bc03["age"], bc03["flux"], method=method)
    assert np.allclose(out[0], bc03["flux"][0, 0], rtol=1e-5, atol=1e-6)
    assert np.allclose(out[1], bc03["flux"][-1, -1], rtol=1e-5, atol=1e-6)


def test_bicubic_matrix_equals_hermite_form(bc03):
    # the 16x16 coefficient matrix and the Hermite-basis evaluation are the same patch
    rng = np.random.default_rng(1)
    zq = rng.uniform(1e-4, 0.05, 64).astype(np.float32)
    aq = rng.uniform(0.0, 10.3, 64).astype(np.float32)
    a = orc.interp2d(zq, aq, bc03["metallicity"], bc03["age"], bc03["flux"], "cubic", dtype=np.float64)
    b = orc.interp2d(zq, aq, bc03["metallicity"], bc03["age"], bc03["flux"], "cubic", dtype=np.float64,
                     hermite=True)
    assert np.allclose(a, b, rtol=1e-10, atol=1e-14)


def test_linear_is_bilinear(bc03):
    rng = np.random.default_rng(2)
    zq = rng.uniform(1e-4, 0.05, 32).astype(np.float32)
    aq = rng.uniform(5.1, 10.3, 32).astype(np.float32)
    out = orc.interp2d(zq, aq, bc03["metallicity"], bc03["age"], bc03["flux"], "linear", dtype=np.float64)
    Zg, Ag, F = (bc03[k].astype(np.float64) for k in ("metallicity", "age", "flux"))
    for k in range(len(zq)):
        i = np.searchsorted(Zg, zq[k], side="right")
        j = np.searchsorted(Ag, aq[k], side="right")
        tx = (zq[k] - Zg[i - 1]) / (Zg[i] - Zg[i - 1])
        ty = (aq[k] - Ag[j - 1]) / (Ag[j] - Ag[j - 1])
        exp = ((1 - tx) * (1 - ty) * F[i - 1, j - 1] + (1 - tx) * ty * F[i - 1, j]
               + tx * (1 - ty) * F[i, j - 1] + tx * ty * F[i, j])
        assert np.allclose(out[k], exp, rtol=1e-12, atol=1e-18)


# ---- PSF / LSF ----------------------------------------------------------------------------------
def test_gaussian_kernel_properties():
    # tests/test_telescope_psf_kernels.py:6-28, tests/test_telescope_psf.py:9-14
    k = orc.gaussian_kernel_2d(5, 5, 1.0)
    assert k.shape == (5, 5) and (k >= 0).all() and np.isclose(k.sum(), 1)
    assert k[2, 2] == k.max()
    k = orc.gaussian_kernel_2d(3, 3, 2.0)
    assert k.shape == (3, 3) and k.sum() == pytest.approx(1)


def test_apply_psf_matches_scipy_same():
    # tests/test_telescope_psf.py:22-54 (== convolve2d(mode="same")), plus asymmetric / even kernels
    from scipy.signal import convolve2d

    cube = np.zeros((10, 10, 3), dtype=np.float32)
    cube[5, 5, :] = 1
    k = np.ones((3, 3), dtype=np.float32)
    out = orc.apply_psf(cube, k)
    for i in range(3):
        assert np.allclose(out[:, :, i], convolve2d(cube[:, :, i], k, mode="same"))
    rng = np.random.default_rng(3)
    cube = rng.random((9, 11, 2)).astype(np.float64)
    for shape in ((5, 5), (3, 5), (4, 4), (2, 3)):
        k = rng.random(shape)
        out = orc.apply_psf(cube, k)
        for i in range(2):
            assert np.allclose(out[:, :, i], convolve2d(cube[:, :, i], k, mode="same"), atol=1e-12)
            assert np.allclose(orc.convolve2d_same(cube[:, :, i], k), out[:, :, i], atol=1e-12)


def test_apply_lsf_delta():
    # tests/test_telescope_lsf.py:6-60
    for pos in (20, 50, 75):
        spectra = np.zeros((1, 100), dtype=np.float32)
        spectra[0, pos] = 1
        out = orc.apply_lsf(spectra, 2.0, 1.0)
        x = np.arange(100)
        g = np.exp(-0.5 * ((x - pos) ** 2) / 2.0**2)
        g /= g.sum()
        assert out.shape == spectra.shape
        assert np.allclose(out[0], g, atol=1e-5)


def test_apply_lsf_is_same_mode_and_unit_integral():
    # tests/test_core_lsf.py:37-74 (shape preserved, kernel integral 1); lsf.py:59-65 == 'same'
    rng = np.random.default_rng(4)
    cube = rng.random((3, 4, 200)).astype(np.float64)
    out = orc.apply_lsf(cube, 0.5, 1.25)
    assert out.shape == cube.shape
    k = orc.lsf_kernel(0.5, 1.25, dtype=np.float64)
    assert len(k) == 25 and np.isclose(k.sum(), 1)
    ref = np.stack([np.convolve(r, k, mode="same") for r in cube.reshape(-1, 200)]).reshape(cube.shape)
    assert np.allclose(out, ref, atol=1e-13)


# ---- grids --------------------------------------------------------------------------------------
def test_muse_wave_grid_bit_exact(muse_wave):
    # rubix/telescope/telescopes.yaml:2-10 + telescope/utils.py:53, pinned by the cube the reference
    # shipped in notebooks/data/dummy_datacube.h5
    w = orc.calculate_wave_seq([4700.15, 9351.4], 1.25)
    assert w.dtype == np.float32 and w.shape == (3721,)
    assert np.array_equal(w, muse_wave)


def test_reshape_array_padding():
    # tests/test_core_data.py:226-248 (device_count faked to 2 / 3)
    a = np.arange(10, dtype=np.float32)
    r = orc.reshape_array(a, 3)
    assert r.shape == (3, 4) and r.ravel()[-2:].tolist() == [0, 0]
    r = orc.reshape_array(np.ones((5, 3), dtype=np.float32), 2)
    assert r.shape == (2, 3, 3) and (r[1, 2] == 0).all()


# ---- the f32 restatement tracks the f64 evaluation of the same formulas -----------------------
@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_f32_oracle_tracks_f64(bc03, muse_wave, tng_subset, method):
    n = 96
    s = {k: v[:n] for k, v in tng_subset.items()}
    edges = np.linspace(-4.7619, 4.7619, 26).astype(np.float32)
    args = (s["coords"], s["velocity"], s["mass"], s["metallicity"], s["age"], edges, 25,
            bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], muse_wave, 0.1)
    c32, i32 = orc.particles_to_cube(*args, method=method, dtype=np.float32)
    c64, i64 = orc.particles_to_cube(*args, method=method, dtype=np.float64)
    assert np.array_equal(i32, i64)
    assert c64.max() > 0
    assert np.abs(c32 - c64).max() <= 2e-5 * np.abs(c64).max()


# ---- rotate_galaxy (the stage before the path) ---------------------------------------------------
def test_rotation_known_answers():
    """tests/test_galaxy_alignment.py: inertia tensor of three unit masses on the axes = diag(2, 2, 2)
    (:67-80), eigenvectors of diag(1, 2, 3) = identity (:83-99), euler (90, 0, 0) =
    [[1,0,0],[0,0,-1],[0,1,0]] (:120-134), apply_rotation / rotate_galaxy of the unit vectors (:137-190)."""
    pos = np.eye(3)
    I = orc.moment_of_inertia_tensor(pos, np.ones(3), 2.0)
    assert np.array_equal(I, 2.0 * np.eye(3))
    assert np.allclose(orc.rotation_matrix_from_inertia_tensor(np.diag([1.0, 2.0, 3.0])), np.eye(3))
    E = orc.euler_rotation_matrix(90.0, 0.0, 0.0)
    assert np.allclose(E, [[1, 0, 0], [0, 0, -1], [0, 1, 0]], atol=1e-15)
    assert np.allclose(pos @ E, [[1, 0, 0], [0, 0, -1], [0, 1, 0]], atol=1e-15)
    vel = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
    p, v, R = orc.rotate_galaxy(pos, vel, np.ones(3), 2.0, 90.0, 0.0, 0.0)
    # a degenerate tensor (all eigenvalues equal): any orthonormal R is an eigenbasis; LAPACK returns the identity
    assert np.allclose(R, np.eye(3))
    assert np.allclose(p, [[1, 0, 0], [0, 0, -1], [0, 1, 0]], atol=1e-15)
    assert np.allclose(v, [[0, 0, -1], [0, 1, 0], [1, 0, 0]], atol=1e-15)


def test_inertia_tensor_padding_quirk():
    """jnp.where(mask, size=N) pads with index 0 (alignment.py:103-106): particle 0 is added once per
    particle outside the radius."""
    rng = np.random.default_rng(3)
    pos = rng.normal(size=(50, 3))
    m = rng.random(50) + 0.5
    r = 1.2
    inside = np.linalg.norm(pos.astype(np.float32), axis=1) <= np.float32(r)
    def tensor(p, w):
        return np.array([[np.sum(w * (np.sum(p ** 2, 1) - p[:, i] ** 2)) if i == j else -np.sum(w * p[:, i] * p[:, j])
                          for j in range(3)] for i in range(3)])
    want = tensor(pos[inside], m[inside]) + (50 - inside.sum()) * tensor(pos[:1], m[:1])
    assert np.allclose(orc.moment_of_inertia_tensor(pos, m, r), want, rtol=1e-12)
    R = orc.rotation_matrix_from_inertia_tensor(want)
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-12)
    w = np.diag(R.T @ want @ R)
    assert np.all(np.diff(w) >= 0)


# ---- apply_noise (the stage after the path) --------------------------------------------------------
def test_threefry2x32_known_answers():
    """Random123 / jax tests/random_test.py testThreefry2x32 vectors."""
    kat = [((0x0, 0x0), (0x0, 0x0), (0x6b200159, 0x99ba4efe)),
           ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
           ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]
    for key, ctr, want in kat:
        a, b = orc.threefry2x32(key, np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
        assert (int(a[0]), int(b[0])) == want


def test_noise_samples_and_s2n():
    n = orc.sample_noise(200000)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1.0) < 0.01 and np.isfinite(n).all()
    u = orc.sample_noise(200000, "uniform")
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    with pytest.raises(ValueError, match="Invalid noise type"):
        orc.sample_noise(4, "poisson")
    rng = np.random.default_rng(4)
    cube = rng.random((6, 5, 40)) + 0.1
    s2n = orc.calculate_S2N(cube, 10.0)
    flux = cube.sum(-1)
    assert np.allclose(s2n, np.sqrt(np.median(flux)) / 10.0 / np.sqrt(flux))
    cube[2, 3] = 0.0   # one spaxel without flux: jnp.median returns NaN -> nan_to_num -> 0 -> no noise at all
    assert np.array_equal(orc.calculate_S2N(cube, 10.0), np.zeros((6, 5)))
    assert np.array_equal(orc.apply_noise(cube, 10.0), cube)
