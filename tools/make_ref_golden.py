"""Golden vectors from the REFERENCE'S OWN SOURCE, executed on numpy through tools/refshim.py.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tools/make_ref_golden.py

The reference functions below are loaded unchanged from /root/reference and called with float64 arrays (seeded inputs,
committed with the outputs), so the vectors are what the reference's formulas MEAN, free of float32 rounding order.
``tests/test_oracle_vs_reference_source.py`` holds the oracle (its float64 mode, the one every GPU parity test compares
against) to them, and re-runs the reference source live whenever /root/reference is present.

Outputs: tests/golden/ref_numpy_stages.npz, ref_numpy_cube.npz, ref_numpy_dust.npz (the dusty variant: extinction curves,
dust-to-gas ratios, cell extinction and apply_spaxel_extinction of rubix/spectra/dust/).
Not covered: a1 (interpax.interp2d is not installable and is not stood in for) -- the cube fixture puts every particle
ON a node of the SSP grid, where the reference's own tests pin the lookup to the template row
(tests/test_core_ssp.py:158-173), so a2 - a7 are exercised end to end without it.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def reference_modules():
    m = {"ifu": refshim.load("rubix/spectra/ifu.py"), "tel": refshim.load("rubix/telescope/utils.py")}
    refshim.load("rubix/telescope/psf/kernels.py")
    m["kern"] = sys.modules["rubix.telescope.psf.kernels"]
    m["psf"] = refshim.load("rubix/telescope/psf/psf.py")
    m["lsf"] = refshim.load("rubix/telescope/lsf/lsf.py")
    m["align"] = refshim.load("rubix/galaxy/alignment.py")
    m["noise"] = refshim.load("rubix/telescope/noise/noise.py")
    return m


def stage_inputs():
    """Seeded inputs of the per-stage vectors (float64 unless integer work)."""
    rng = np.random.default_rng(42)
    tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
    wave = np.load(os.path.join(OUT, "muse_wave.npy")).astype(np.float64)
    # a0: float32 edges as get_spatial_bin_edges hands them over (26 and 27 of them), coords on / next to edges
    half = np.float32(4.7619)
    e26 = (-half + np.float32(2 * half / 25) * np.arange(26, dtype=np.float32)).astype(np.float32)
    e27 = (-half + np.float32(2 * half / 25) * np.arange(27, dtype=np.float32)).astype(np.float32)
    c = rng.normal(0, 2.5, (3000, 3)).astype(np.float32)
    c[:26, 0] = e26
    c[26:52, 1] = e26
    c[52:78, 0] = np.nextafter(e26, np.float32(np.inf))
    c[78:104, 0] = np.nextafter(e26, np.float32(-np.inf))
    c[104:110] = np.float32([[-9, 0, 0], [9, 0, 0], [0, -9, 0], [0, 9, 0], [4.7619, 4.7619, 0], [-4.7619, -4.7619, 0]])
    # a3 / a4: twelve template rows at z = 0.1 with velocities up to +-900 km/s, a zero spectrum, a constant one
    lam_ssp = tpl["wavelength"].astype(np.float64)
    rows = tpl["flux"].reshape(-1, lam_ssp.size)[rng.integers(0, 6 * 221, 12)].astype(np.float64)
    rows[3] = 0.0
    rows[7] = 1.0
    vel = rng.normal(0, 300, (12, 3))
    vel[0] = [0.0, 0.0, 0.0]
    vel[1, 2], vel[2, 2] = 900.0, -900.0
    return dict(edges26=e26, edges27=e27, coords=c, lam_ssp=lam_ssp, wave=wave, rows=rows, vel=vel,
                cube_small=rng.random((9, 7, 40)), psf_odd=None, lsf_cube=rng.random((3, 4, 300)),
                gal_pos=rng.normal(0, 3, (500, 3)), gal_vel=rng.normal(0, 100, (500, 3)),
                gal_mass=rng.uniform(0.5, 1.5, 500), noise_cube=rng.random((6, 5, 30)))


def run_stages(m, x):
    """Every per-stage vector, from the reference's functions."""
    ifu, tel, kern, psf, lsf, align, noise = (m[k] for k in ("ifu", "tel", "kern", "psf", "lsf", "align", "noise"))
    o = {}
    for tag in ("26", "27"):
        e = x["edges" + tag]
        o["pixel" + tag] = np.asarray(tel.square_spaxel_assignment(x["coords"], e)).astype(np.int32)
        o["mask" + tag] = np.asarray(tel.mask_particles_outside_aperture(x["coords"], e)).astype(bool)
    lam_z = ifu.cosmological_doppler_shift(0.1, x["lam_ssp"])
    o["lam_z"] = lam_z
    shifted = ifu.velocity_doppler_shift(lam_z, x["vel"], "z")
    o["shifted"] = shifted
    o["diff"] = ifu.calculate_diff(x["wave"])
    o["resampled"] = np.stack([ifu.resample_spectrum(x["rows"][k], shifted[k], x["wave"]) for k in range(12)])
    ids = np.array([0, 3, 3, 8, 1, 9, 7, 2, 0, 3, 4, 12])      # 9, 12: beyond the 3 x 3 cube -> dropped
    o["cube_ids"] = ids
    o["cube"] = ifu.calculate_cube(o["resampled"], ids, 3)
    for name, (a, b, s) in {"psf55": (5, 5, 0.6), "psf46": (4, 6, 1.3), "psf33": (3, 3, 2.0)}.items():
        k = kern.gaussian_kernel_2d(a, b, s)
        o[name] = k
        o[name + "_applied"] = psf.apply_psf(x["cube_small"], k)
    skew = np.arange(1.0, 16.0).reshape(3, 5) / 120.0        # asymmetric taps: pins the orientation of the convolution
    o["psf_skew"] = skew
    o["psf_skew_applied"] = psf.apply_psf(x["cube_small"], skew)
    o["lsf_kernel"] = lsf._get_kernel(0.5, 1.25, factor=12)
    o["lsf_kernel_wide"] = lsf._get_kernel(3.0, 1.25, factor=12)
    o["lsf_applied"] = lsf.apply_lsf(x["lsf_cube"], 0.5, 1.25)
    o["lsf_applied_wide"] = lsf.apply_lsf(x["lsf_cube"], 3.0, 1.25)
    o["inertia"] = align.moment_of_inertia_tensor(x["gal_pos"], x["gal_mass"], 4.0)
    o["euler"] = align.euler_rotation_matrix(20.0, -35.0, 70.0)
    p, v = align.rotate_galaxy(x["gal_pos"], x["gal_vel"], x["gal_mass"], 4.0, 20.0, -35.0, 70.0)
    o["gal_pos_rot"], o["gal_vel_rot"] = p, v
    # center_particles (rubix/galaxy/alignment.py:14-64): coordinates relative to the centre, velocities relative to the
    # median velocity within 10 kpc -- float32 arrays, as prepare_input holds them
    from types import SimpleNamespace as NS
    f32 = np.float32
    centre = f32([3.0, -2.0, 1.0])
    for key in ("stars", "gas"):
        pos = (x["gal_pos"] * (4.0 if key == "gas" else 1.0) + centre).astype(f32)   # the gas reaches beyond 10 kpc
        vel = (x["gal_vel"] + f32([120.0, -40.0, 15.0])).astype(f32)
        rd = NS(stars=NS(coords=pos, velocity=vel), gas=NS(coords=pos, velocity=vel), galaxy=NS(center=centre))
        rd = align.center_particles(rd, key)
        part = getattr(rd, key)
        o[f"centre_{key}_coords"], o[f"centre_{key}_velocity"] = part.coords, part.velocity
    o["centre"] = centre
    try:
        align.center_particles(NS(stars=NS(coords=x["gal_pos"].astype(f32) + f32(100.0), velocity=x["gal_vel"].astype(f32)),
                                  galaxy=NS(center=centre)), "stars")
        o["centre_error"] = np.array("")
    except Exception as e:   # noqa: BLE001
        o["centre_error"] = np.array(f"{type(e).__name__}: {e}")
    nc = x["noise_cube"].copy()
    o["s2n"] = noise.calculate_S2N(nc, 50.0)
    nc[2, 3] = 0.0                                            # a flux-less spaxel: the NaN-propagating median
    o["s2n_with_dark_spaxel"] = noise.calculate_S2N(nc, 50.0)
    return {k: np.asarray(v) for k, v in o.items()}


def cube_inputs(n=1500, S=7):
    """Particles ON nodes of the SSP grid (float32 values, as the CUDA path receives them) on a 7 x 7 spaxel grid."""
    rng = np.random.default_rng(4242)
    tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
    half = np.float32(4.7619)
    edges = (-half + np.float32(2 * half / S) * np.arange(S + 1, dtype=np.float32)).astype(np.float32)
    iz, ia = rng.integers(0, 6, n), rng.integers(0, 221, n)
    coords = rng.normal(0, 2.5, (n, 3)).astype(np.float32)       # some fall outside the aperture
    vel = rng.normal(0, 200, (n, 3)).astype(np.float32)
    mass = rng.uniform(0.5, 1.5, n).astype(np.float32)
    # keep only particles without a Doppler-shifted knot within 0.01 A of a band edge: there the in-range mask of
    # rubix/spectra/ifu.py:241-244 flips with the rounding of lam * d, float32 and float64 evaluations of the reference
    # disagree, and the vector could pin nothing for a float32 implementation (tests/helpers.py: well_conditioned)
    wave = np.load(os.path.join(OUT, "muse_wave.npy")).astype(np.float64)
    lam = 1.1 * tpl["wavelength"].astype(np.float64)[None, :] * np.exp(vel[:, 2].astype(np.float64) / 299792.458)[:, None]
    keep = np.minimum(np.abs(lam - wave[0]).min(1), np.abs(lam - wave[-1]).min(1)) > 1e-2
    iz, ia, coords, vel, mass = iz[keep], ia[keep], coords[keep], vel[keep], mass[keep]
    return dict(edges=edges, node_z=iz, node_age=ia, coords=coords, velocity=vel, mass=mass,
                metallicity=tpl["metallicity"][iz], age=tpl["age"][ia])


def run_cube(m, x, S=7):
    """filter_particles -> spaxel_assignment -> (node lookup) -> scale -> Doppler -> resample -> cube -> PSF -> LSF with
    the reference's functions, float64 arithmetic on the float32 input values."""
    ifu, tel, kern, psf, lsf = (m[k] for k in ("ifu", "tel", "kern", "psf", "lsf"))
    tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
    wave = np.load(os.path.join(OUT, "muse_wave.npy")).astype(np.float64)
    mask = np.asarray(tel.mask_particles_outside_aperture(x["coords"], x["edges"]))
    mass = np.where(mask, x["mass"].astype(np.float64), 0.0)     # rubix/core/telescope.py:155-174
    pix = np.asarray(tel.square_spaxel_assignment(x["coords"], x["edges"]))
    spectra = tpl["flux"].astype(np.float64)[x["node_z"], x["node_age"]]     # a1 at a node = the template row
    spectra = spectra * mass[:, None]                            # rubix/core/ifu.py:152-154
    lam_z = ifu.cosmological_doppler_shift(0.1, tpl["wavelength"].astype(np.float64))
    shifted = ifu.velocity_doppler_shift(lam_z, x["velocity"].astype(np.float64), "z")
    res = np.stack([ifu.resample_spectrum(spectra[k], shifted[k], wave) for k in range(len(mass))])
    cube = ifu.calculate_cube(res, pix, S)
    conv = lsf.apply_lsf(psf.apply_psf(cube, kern.gaussian_kernel_2d(5, 5, 0.6)), 0.5, 1.25)
    # the cubes are 49 x 3721 doubles (1.5 MB each): the fixture keeps every 4th channel of every spaxel, and all 3721
    # channels summed over the spaxels (thin() below), so that every voxel still enters a compared number
    return {"pixel": pix.astype(np.int32), "mask": mask, **thin("cube", cube), **thin("cube_psf_lsf", conv)}


def run_cosmology():
    """rubix/cosmology/base.py + utils.py and calculate_spatial_bin_edges (rubix/telescope/utils.py:30-37) in float32:
    the class holds float32 parameters, and with x64 off ``jnp.linspace`` of Python floats is float32 too (the stand-in
    is switched to that for this function).  Both sides are numpy here, so this pins the LOGIC (256-point trapezoid by
    scan, constants, the arange that yields 26 or 27 edges), not XLA's float32 pow / exp roundings."""
    import jax.numpy as jnp      # the stand-in installed by refshim
    refshim.load("rubix/cosmology/utils.py")
    sys.modules.pop("rubix.cosmology.base", None)
    keep = jnp.linspace
    jnp.linspace = lambda *a, **k: np.linspace(*a, **k).astype(np.float32)
    try:
        base = refshim.load("rubix/cosmology/base.py")
        sys.modules.pop("rubix.telescope.utils", None)
        tel = refshim.load("rubix/telescope/utils.py")          # re-bound to the real BaseCosmology
        c = base.BaseCosmology(0.3075, -1.0, 0.0, 0.6774)       # PLANCK15, rubix/cosmology/__init__.py:3
        zs = np.array([0.01, 0.05, 0.1, 0.3, 1.0])
        o = {"z": zs, "angular_scale": np.array([c.angular_scale(float(z)) for z in zs]),
             "comoving": np.array([c.comoving_distance_to_z(float(z)) for z in zs]),
             "luminosity": np.array([c.luminosity_distance_to_z(float(z)) for z in zs])}
        for tag, fov, nb in (("muse", 5.0, 25), ("fov30", 30.0, 150), ("z03", 5.0, 25)):
            # spatial_bins as a Python int: numpy would promote float32 / np.int64 to float64, jax (x64 off) keeps float32
            e, size = tel.calculate_spatial_bin_edges(fov, int(nb), 0.3 if tag == "z03" else 0.1, c)
            o["edges_" + tag], o["size_" + tag] = np.asarray(e), np.asarray(size)
    finally:
        jnp.linspace = keep
    return o


# malformed / valid configurations handed to the reference's factories and to the mirror's (tests compare the outcome)
BOUNDARY_CASES = {
    "get_convolve_psf": [
        {"telescope": {"name": "MUSE"}},
        {"telescope": {"name": "MUSE", "psf": {}}},
        {"telescope": {"name": "MUSE", "psf": {"name": "gaussian"}}},
        {"telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5}}},
        {"telescope": {"name": "MUSE", "psf": {"name": "moffat", "size": 5, "sigma": 0.6}}},
        {"telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}}},
    ],
    "get_convolve_lsf": [
        {"telescope": {"name": "MUSE"}},
        {"telescope": {"name": "MUSE", "lsf": {}}},
        {"telescope": {"name": "MUSE", "lsf": {"sigma": 0.5}}},
    ],
    "get_apply_noise": [
        {"telescope": {"name": "MUSE"}},
        {"telescope": {"name": "MUSE", "noise": {}}},
        {"telescope": {"name": "MUSE", "noise": {"signal_to_noise": 10}}},
        {"telescope": {"name": "MUSE", "noise": {"signal_to_noise": 10, "noise_distribution": "normal"}}},
    ],
    "get_galaxy_rotation": [
        {"galaxy": {"dist_z": 0.1}},
        {"galaxy": {"dist_z": 0.1, "rotation": {"type": "sideways"}}},
        {"galaxy": {"dist_z": 0.1, "rotation": {"alpha": 1.0}}},
        {"galaxy": {"dist_z": 0.1, "rotation": {"alpha": 1.0, "beta": 2.0}}},
        {"galaxy": {"dist_z": 0.1, "rotation": {"type": "face-on"}}},
        {"galaxy": {"dist_z": 0.1, "rotation": {"type": "edge-on"}}},
        {"galaxy": {"dist_z": 0.1, "rotation": {"alpha": 20.0, "beta": -35.0, "gamma": 70.0}}},
    ],
    "get_ssp": [
        {},
        {"ssp": {}},
        {"ssp": {"template": {}}},
    ],
    # the interpolation method the lookup is built with (rubix/core/ssp.py:57-62: "cubic" when ssp.method is absent)
    "ssp_method": [
        {"ssp": {"template": {"name": "BruzualCharlot2003"}}},
        {"ssp": {"template": {"name": "BruzualCharlot2003"}, "method": "linear"}},
        {"ssp": {"template": {"name": "BruzualCharlot2003"}, "method": "cubic"}},
    ],
    "get_extinction": [
        {"ssp": {"template": {"name": "BruzualCharlot2003"}}, "telescope": {"name": "MUSE"}, "galaxy": {"dist_z": 0.1},
         "cosmology": {"name": "PLANCK15"}},
        {"ssp": {"template": {"name": "BruzualCharlot2003"}, "dust": {}}, "telescope": {"name": "MUSE"},
         "galaxy": {"dist_z": 0.1}, "cosmology": {"name": "PLANCK15"}},
        {"ssp": {"template": {"name": "BruzualCharlot2003"},
                 "dust": {"extinction_model": "Cardelli89", "Rv": 3.1, "dust_grain_density": 3.5}},
         "telescope": {"name": "MUSE"}, "galaxy": {"dist_z": 0.1}, "cosmology": {"name": "PLANCK15"}},
    ],
    "get_cosmology": [
        {"cosmology": {"name": "WMAP9"}},
        {"cosmology": {"name": "planck15"}},
        {"cosmology": {"name": "CUSTOM", "args": {"Om0": 0.3, "w0": -1.0, "wa": 0.0, "h": 0.7}}},
    ],
}


def outcome(factory, cfg):
    """What a factory call gives: ("error", type name, message) or ("ok", the returned object's __name__ / type)."""
    try:
        f = factory(cfg)
    except Exception as e:   # noqa: BLE001
        return ["error", type(e).__name__, str(e)]
    return ["ok", f if isinstance(f, str) else getattr(f, "__name__", type(f).__name__)]


def run_boundary(x):
    """The reference's own factories (rubix/core/{psf,lsf,noise,rotation,cosmology}.py): outcomes for BOUNDARY_CASES
    and what their closures do to a stand-in RubixData (only the attributes the closures touch)."""
    import json
    from types import SimpleNamespace as NS
    refshim.load("rubix/telescope/psf/kernels.py")
    refshim.load("rubix/telescope/psf/psf.py")
    refshim.load("rubix/telescope/lsf/lsf.py")
    refshim.load("rubix/telescope/noise/noise.py")
    refshim.load("rubix/galaxy/alignment.py")
    refshim.load("rubix/cosmology/utils.py")
    sys.modules.pop("rubix.cosmology.base", None)
    base = refshim.load("rubix/cosmology/base.py")
    cosmo_pkg = sys.modules["rubix.cosmology"]
    cosmo_pkg.RubixCosmology = base.BaseCosmology                      # rubix/cosmology/__init__.py:1-3
    cosmo_pkg.PLANCK15 = base.BaseCosmology(0.3075, -1.0, 0.0, 0.6774)
    # get_telescope needs the telescope factory (yaml + equinox classes); the LSF factory reads one attribute of it
    sys.modules["rubix.core.telescope"] = type(sys)("rubix.core.telescope")
    sys.modules["rubix.core.telescope"].get_telescope = lambda config: NS(     # telescopes.yaml: MUSE
        wave_res=1.25, sbin=np.int64(25), fov=5.0, wave_seq=np.arange(4700.15, 9351.4, 1.25))
    # get_ssp_template reads the HDF5 template (h5py): stood in by an object whose lookup factory hands back the
    # interpolation method it was asked for
    for pkg in ("rubix.spectra.ssp",):
        sys.modules.setdefault(pkg, type(sys)(pkg)).__path__ = []
    fmod = type(sys)("rubix.spectra.ssp.factory")
    fmod.get_ssp_template = lambda name: NS(get_lookup_interpolation=lambda method: method)
    sys.modules["rubix.spectra.ssp.factory"] = fmod
    sys.modules.pop("rubix.core.ssp", None)
    for f in ("helpers", "generic_models", "dust_baseclasses", "extinction_models", "dust_extinction"):
        refshim.load(f"rubix/spectra/dust/{f}.py")
    sys.modules.pop("rubix.core.dust", None)
    core = {k: refshim.load(f"rubix/core/{k}.py") for k in ("psf", "lsf", "noise", "rotation", "cosmology", "ssp", "dust")}
    fac = {"get_convolve_psf": core["psf"].get_convolve_psf, "get_convolve_lsf": core["lsf"].get_convolve_lsf,
           "get_apply_noise": core["noise"].get_apply_noise, "get_galaxy_rotation": core["rotation"].get_galaxy_rotation,
           "get_cosmology": core["cosmology"].get_cosmology, "get_ssp": core["ssp"].get_ssp,
           "ssp_method": core["ssp"].get_lookup_interpolation, "get_extinction": core["dust"].get_extinction}
    table = {name: [outcome(fac[name], cfg) for cfg in cases] for name, cases in BOUNDARY_CASES.items()}
    o = {"outcomes_json": np.array(json.dumps(table))}
    cube = x["lsf_cube"]
    rd = NS(stars=NS(datacube=cube.copy()))
    o["psf_closure"] = fac["get_convolve_psf"](BOUNDARY_CASES["get_convolve_psf"][-1])(rd).stars.datacube
    o["psf_lsf_closures"] = fac["get_convolve_lsf"](BOUNDARY_CASES["get_convolve_lsf"][-1])(rd).stars.datacube
    # apply_noise: the random numbers come from jax.random (threefry, not available here), so the stand-in hands out a
    # fixed array of standard normals, stored with the result: what is pinned is the ARITHMETIC around them
    # (rubix/core/noise.py:63-78, rubix/telescope/noise/noise.py:81-115), for a cube with and without a dark spaxel
    rnd = sys.modules["jax.random"]
    normals = np.random.default_rng(123).standard_normal(x["noise_cube"].shape)
    rnd.PRNGKey = lambda seed: ("key", seed)
    rnd.normal = lambda key, shape: normals.reshape(shape)
    rnd.uniform = lambda key, shape: (normals.reshape(shape) % 1.0)
    o["noise_normals"] = normals
    for tag, dark in (("noise_closure", False), ("noise_closure_dark_spaxel", True)):
        nc = x["noise_cube"].copy()
        if dark:
            nc[2, 3] = 0.0
        with np.errstate(all="ignore"):
            o[tag] = fac["get_apply_noise"](BOUNDARY_CASES["get_apply_noise"][-1])(NS(stars=NS(datacube=nc))).stars.datacube
    cfg = dict(BOUNDARY_CASES["get_galaxy_rotation"][-1], data={"args": {"particle_type": ["stars"]}})
    rd = NS(stars=NS(coords=x["gal_pos"].copy(), velocity=x["gal_vel"].copy(), mass=x["gal_mass"].copy()),
            galaxy=NS(halfmassrad_stars=4.0))
    rd = fac["get_galaxy_rotation"](cfg)(rd)
    o["rotation_closure_coords"], o["rotation_closure_velocity"] = rd.stars.coords, rd.stars.velocity
    return {k: np.asarray(v) for k, v in o.items()}


def run_core_closures(xc, S=7, n_dev=2):
    """The closures of rubix/core/ifu.py -- scale_spectrum_by_mass, doppler_shift_and_resampling, calculate_datacube --
    run from source on a stand-in RubixData with a DEVICE AXIS of two: the particles are padded and reshaped to
    (2, P, ...) as rubix/core/data.py:447-487 does, pmap runs the per-device cube, jnp.sum(axis=0) adds them
    (rubix/core/ifu.py:324-333: the reduction the multi-GPU path does with one NCCL collective).  get_telescope and
    get_ssp are stood in by objects holding the three attributes the closures read; the lookup (a1, interpax) is
    replaced by the template rows of the node particles, as in run_cube."""
    from types import SimpleNamespace as NS
    tpl = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
    wave = np.load(os.path.join(OUT, "muse_wave.npy")).astype(np.float64)
    tel = refshim.load("rubix/telescope/utils.py")
    refshim.load("rubix/spectra/ifu.py")
    sys.modules["rubix.core.telescope"] = type(sys)("rubix.core.telescope")
    sys.modules["rubix.core.telescope"].get_telescope = lambda config: NS(wave_seq=wave, sbin=np.int64(S), wave_res=1.25)
    ssp = type(sys)("rubix.core.ssp")
    ssp.get_ssp = lambda config: NS(wavelength=tpl["wavelength"].astype(np.float64))
    ssp.get_lookup_interpolation = ssp.get_lookup_interpolation_pmap = ssp.get_lookup_interpolation_vmap = None
    sys.modules["rubix.core.ssp"] = ssp
    sys.modules.pop("rubix.core.ifu", None)
    ifu = refshim.load("rubix/core/ifu.py")
    cfg = {"galaxy": {"dist_z": 0.1}, "telescope": {"name": "MUSE"}}
    mask = np.asarray(tel.mask_particles_outside_aperture(xc["coords"], xc["edges"]))
    n = len(mask)
    per = -(-n // n_dev)

    def shard(a):                                   # rubix/core/data.py:471-482: zero padding, then (n_dev, per, ...)
        a = np.asarray(a)
        pad = np.zeros((per * n_dev - n,) + a.shape[1:], dtype=a.dtype)
        return np.concatenate([a, pad]).reshape((n_dev, per) + a.shape[1:])

    rows = tpl["flux"].astype(np.float64)[xc["node_z"], xc["node_age"]]
    rd = NS(stars=NS(spectra=shard(rows), mass=shard(np.where(mask, xc["mass"].astype(np.float64), 0.0)),
                     velocity=shard(xc["velocity"].astype(np.float64)),
                     pixel_assignment=shard(np.asarray(tel.square_spaxel_assignment(xc["coords"], xc["edges"])))),
            gas=NS(spectra=None, velocity=None))
    for get in (ifu.get_scale_spectrum_by_mass, ifu.get_doppler_shift_and_resampling, ifu.get_calculate_datacube):
        rd = get(cfg)(rd)
    return np.asarray(rd.stars.datacube), tuple(rd.stars.spectra.shape)


def run_pipeline_orders():
    """rubix/pipeline/{transformer,abstract_pipeline,linear_pipeline}.py from source on the reference's own
    rubix/config/pipeline_config.yml: for every pipeline defined there, the node order LinearTransformerPipeline assembles
    from the depends_on chain, and the order in which its composed expression calls the twelve functions
    rubix/core/pipeline.py:105-133 registers (stood in by recorders of the same names)."""
    import json
    import yaml
    for pkg in ("rubix.pipeline",):
        sys.modules.setdefault(pkg, type(sys)(pkg)).__path__ = []
    sys.modules["jax"].make_jaxpr = lambda f, **k: f
    refshim.load("rubix/pipeline/transformer.py")
    sys.modules["rubix.pipeline"].abstract_pipeline = refshim.load("rubix/pipeline/abstract_pipeline.py")
    lin = refshim.load("rubix/pipeline/linear_pipeline.py")
    cfgs = yaml.safe_load(open(os.path.join(refshim.REF, "rubix", "config", "pipeline_config.yml")))
    names = ["rotate_galaxy", "filter_particles", "spaxel_assignment", "calculate_spectra", "reshape_data",
             "scale_spectrum_by_mass", "doppler_shift_and_resampling", "calculate_extinction", "calculate_datacube",
             "convolve_psf", "convolve_lsf", "apply_noise"]

    def recorder(name):
        def fn(trace):
            return trace + [name]
        fn.__name__ = name
        return fn

    out = {}
    for pname, cfg in cfgs.items():
        pipe = lin.LinearTransformerPipeline(cfg, [recorder(n) for n in names])
        out[pname] = {"nodes": list(pipe._names), "called": pipe.expression([])}
    return {"json": np.array(json.dumps(out))}


def run_config_data():
    """The reference's configuration files as data: pipeline_config.yml, telescopes.yaml and the parts of rubix_config.yml
    the path reads (constants, ifu.doppler, ssp.dust, the BC03 template entry)."""
    import json
    import yaml
    rd = lambda *p: yaml.safe_load(open(os.path.join(refshim.REF, "rubix", *p)))
    cfg = rd("config", "rubix_config.yml")
    out = {"pipelines": rd("config", "pipeline_config.yml"), "telescopes": rd("telescope", "telescopes.yaml"),
           "constants": cfg["constants"], "ifu": cfg["ifu"], "dust": cfg["ssp"]["dust"],
           "bc03": cfg["ssp"]["templates"]["BruzualCharlot2003"]}
    return {"json": np.array(json.dumps(out, sort_keys=True))}


FITS_CONFIG = {
    "pipeline": {"name": "calc_ifu"}, "galaxy": {"dist_z": 0.1, "rotation": {"type": "edge-on"}},
    "simulation": {"name": "IllustrisTNG"}, "cosmology": {"name": "PLANCK15"},
    "data": {"args": {"snapshot": 99, "particle_type": ["stars"]}, "load_galaxy_args": {"id": 14},
             "subset": {"use_subset": True, "subset_size": 1000}},
    "ssp": {"template": {"name": "BruzualCharlot2003"}},
    "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}, "lsf": {"sigma": 0.5},
                  "noise": {"signal_to_noise": 10, "noise_distribution": "normal"}},
}


def run_store_fits():
    """store_fits (rubix/core/fits.py:13-101) from source with astropy.io.fits stood in by recorders: the two headers
    (keywords, values, order), the array handed to the IMAGE extension and the output file name."""
    import json
    import types
    from types import SimpleNamespace as NS
    rec = {}

    class Header(dict):
        pass

    class HDUList(list):
        def writeto(self, name, overwrite=False):
            rec["filename"] = name

    fits = types.ModuleType("astropy.io.fits")
    fits.Header = Header
    fits.PrimaryHDU = lambda header=None: rec.setdefault("primary", header)
    fits.ImageHDU = lambda data, header=None: rec.update(image=np.asarray(data), image_header=header)
    fits.HDUList = HDUList
    for name, mod in (("astropy", types.ModuleType("astropy")), ("astropy.io", types.ModuleType("astropy.io")),
                      ("astropy.io.fits", fits), ("matplotlib", types.ModuleType("matplotlib")),
                      ("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot")),
                      ("matplotlib.colors", types.ModuleType("matplotlib.colors")),
                      ("mpdaf", types.ModuleType("mpdaf")), ("mpdaf.obj", types.ModuleType("mpdaf.obj"))):
        sys.modules[name] = mod
    sys.modules["astropy.io"].fits = fits
    sys.modules["matplotlib.colors"].LogNorm = object
    sys.modules["mpdaf.obj"].Cube = object
    sys.modules["rubix.core.telescope"] = types.ModuleType("rubix.core.telescope")
    sys.modules["rubix.core.telescope"].get_telescope = lambda config: NS(spatial_res=0.2, wave_res=1.25,
                                                                          wave_range=[4700.15, 9351.4])
    sys.modules.pop("rubix.core.fits", None)
    mod = refshim.load("rubix/core/fits.py")
    cube = np.arange(2 * 3 * 5, dtype=np.float32).reshape(2, 3, 5)
    import tempfile
    d = tempfile.mkdtemp()
    mod.store_fits(FITS_CONFIG, NS(stars=NS(datacube=cube)), d + "/out_")
    for m in ("astropy", "astropy.io", "astropy.io.fits", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "mpdaf",
              "mpdaf.obj"):
        sys.modules.pop(m, None)
    return {"json": np.array(json.dumps({"primary": list(rec["primary"].items()),
                                         "image_header": list(rec["image_header"].items()),
                                         "filename": os.path.basename(rec["filename"])})),
            "image": rec["image"]}


def run_data_classes():
    """rubix/core/data.py from source: the attribute names of Galaxy / StarsData / GasData / RubixData (a8, the types that
    cross the boundary) and reshape_array for a device count of two and three (zero padding, 1-D and 2-D)."""
    import json
    import types
    sys.modules["jax.tree_util"].register_pytree_node_class = lambda c=None: (c if c is not None else (lambda k: k))
    g = types.ModuleType("rubix.galaxy")
    g.IllustrisAPI, g.get_input_handler, g.__path__ = object, (lambda *a, **k: None), []
    sys.modules["rubix.galaxy"] = g
    refshim.load("rubix/galaxy/alignment.py")
    u = sys.modules.get("rubix.utils") or types.ModuleType("rubix.utils")
    u.load_galaxy_data = lambda *a, **k: None
    if not hasattr(u, "read_yaml"):
        u.read_yaml = lambda path: None
    sys.modules["rubix.utils"] = u
    sys.modules.pop("rubix.core.data", None)
    data = refshim.load("rubix/core/data.py")
    names = {}
    for cls in ("Galaxy", "StarsData", "GasData", "RubixData"):
        obj = getattr(data, cls)()
        names[cls] = sorted(a for a in dir(obj) if not a.startswith("_") and not callable(getattr(obj, a)))
    o = {"names_json": np.array(json.dumps(names))}
    a1, a2 = np.arange(1.0, 8.0), np.arange(1.0, 15.0).reshape(7, 2)
    for n_dev in (2, 3):
        sys.modules["jax"].device_count = lambda n=n_dev: n
        o[f"reshape1d_{n_dev}"], o[f"reshape2d_{n_dev}"] = data.reshape_array(a1), data.reshape_array(a2)
    # prepare_input (rubix/core/data.py:491-603) itself, its HDF5 reader stood in by a function returning arrays
    f32 = np.float32
    x = stage_inputs()
    centre = f32([3.0, -2.0, 1.0])
    stars = dict(coords=(x["gal_pos"] + centre).astype(f32), velocity=(x["gal_vel"] + f32([120.0, -40.0, 15.0])).astype(f32),
                 mass=np.linspace(0.5, 1.5, 500).astype(f32), metallicity=np.full(500, 0.01, f32),
                 age=np.linspace(5.0, 10.0, 500).astype(f32))
    gas = dict(coords=(x["gal_pos"] * 4.0 + centre).astype(f32), velocity=stars["velocity"][::-1].copy(),
               mass=np.arange(500, dtype=f32), metals=np.arange(4500, dtype=f32).reshape(500, 9))
    units = {"galaxy": {"redshift": "", "center": "kpc", "halfmassrad_stars": "kpc"},
             "stars": {k: "u" for k in stars}, "gas": {k: "u" for k in gas}}
    for tag, types_, subset in (("both", ["stars", "gas"], None), ("both_subset", ["stars", "gas"], 40),
                                ("gas_only_subset", ["gas"], 25)):
        raw = {"redshift": 0.1, "subhalo_center": centre, "subhalo_halfmassrad_stars": 2.0,
               "particle_data": {k: dict(v) for k, v in (("stars", stars), ("gas", gas)) if k in types_}}
        data.load_galaxy_data = lambda path, raw=raw: (raw, units)
        cfg = {"output_path": "/nonexistent", "data": {"args": {"particle_type": types_}}}
        if subset:
            cfg["data"]["subset"] = {"use_subset": True, "subset_size": subset}
        rd = data.prepare_input(cfg)
        for part in types_:
            for k in raw["particle_data"][part]:
                o[f"prepare_{tag}_{part}_{k}"] = np.asarray(getattr(getattr(rd, part), k))
    sys.modules.pop("rubix.core.data", None)          # other runs register their own stand-in for this module
    sys.modules["rubix.core.data"] = types.ModuleType("rubix.core.data")
    sys.modules["rubix.core.data"].RubixData = sys.modules["rubix.core.data"].StarsData = object
    sys.modules["rubix.core.data"].GasData = object
    return o


def run_telescopes():
    """rubix/telescope/{apertures,base,factory}.py from source: every telescope of telescopes.yaml through
    TelescopeFactory.create_telescope -- sbin, aperture mask (square / circular / hexagonal), length and end points of
    wave_seq and wave_edges.  The wave grids are numpy float64 aranges here (float32 in jax, x64 off): lengths and end
    points are compared, the MUSE grid itself is pinned bit-exactly by tests/golden/muse_wave.npy."""
    import json
    import warnings
    import yaml
    mod = type(sys)("rubix.utils")
    mod.read_yaml = lambda path: yaml.safe_load(open(path))
    sys.modules["rubix.utils"] = mod
    sys.modules.pop("rubix.telescope.utils", None)
    refshim.load("rubix/telescope/utils.py")
    refshim.load("rubix/telescope/apertures.py")
    refshim.load("rubix/telescope/base.py")
    fac = refshim.load("rubix/telescope/factory.py")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f = fac.TelescopeFactory()
    o, meta = {}, {}
    for name in f.telescopes_config:
        t = f.create_telescope(name)
        o["aperture_" + name] = np.packbits(np.asarray(t.aperture_region) > 0)
        ws, we = np.asarray(t.wave_seq), np.asarray(t.wave_edges)
        meta[name] = {"sbin": int(t.sbin), "n_aperture": int(np.asarray(t.aperture_region).size),
                      "aperture_sum": float(np.asarray(t.aperture_region).sum()), "pixel_type": str(t.pixel_type),
                      "n_wave": int(ws.size), "wave_first": float(ws[0]), "wave_last": float(ws[-1]),
                      "n_edges": int(we.size), "edge_first": float(we[0]), "edge_last": float(we[-1]),
                      "fov": float(t.fov), "spatial_res": float(t.spatial_res), "wave_res": float(t.wave_res)}
    try:
        f.create_telescope("HST")
    except Exception as e:   # noqa: BLE001
        meta["__unknown__"] = [type(e).__name__, str(e)]
    o["meta_json"] = np.array(json.dumps(meta))
    return o


def dust_inputs():
    """Gas cells and stars on 12 spaxels: crowded spaxels, spaxels with 0 / 1 / 2 gas cells, stars in front of and
    behind all the gas of their spaxel, a spaxel with gas and no stars.  float32 values (the CUDA path's inputs)."""
    rng = np.random.default_rng(77)
    S, ng, ns, W = 12, 700, 260, 24
    gp = rng.integers(0, 8, ng)                     # spaxels 0 .. 7 crowded
    gp[:1] = 8                                      # spaxel 8: one gas cell
    gp[1:3] = 9                                     # spaxel 9: two gas cells; 10: none; 11: gas but no stars
    gp[3:9] = 11
    sp = rng.integers(0, 11, ns)
    gz = rng.normal(0, 1.0, ng).astype(np.float32)
    sz = rng.normal(0, 1.6, ns).astype(np.float32)
    sz[:4] = [-30.0, 30.0, -29.0, 31.0]             # in front of / behind everything
    sp[:4] = [2, 2, 8, 9]
    gas_coords = np.concatenate([rng.normal(0, 1, (ng, 2)).astype(np.float32), gz[:, None]], axis=1)
    star_coords = np.concatenate([rng.normal(0, 1, (ns, 2)).astype(np.float32), sz[:, None]], axis=1)
    metals = rng.uniform(0.001, 0.05, (ng, 9)).astype(np.float32)
    metals[:, 0] = rng.uniform(0.70, 0.76, ng).astype(np.float32)            # hydrogen
    metals[:, 4] = (10 ** rng.uniform(-4.2, -2.0, ng)).astype(np.float32)    # oxygen: 12 + log(O/H) from 7.6 to 9.8
    return dict(S=np.int64(S), gas_coords=gas_coords, gas_pixel=gp.astype(np.int32),
                gas_mass=rng.uniform(1e4, 1e6, ng).astype(np.float32), gas_metals=metals, star_coords=star_coords,
                star_pixel=sp.astype(np.int32), spectra=rng.uniform(0.5, 2.0, (ns, W)),
                wave=np.linspace(3600.0, 9900.0, W).astype(np.float32), spaxel_area=np.float64(0.145))


def run_dust(x):
    """apply_spaxel_extinction (rubix/spectra/dust/dust_extinction.py:169-358) and its helpers, for both extinction
    curves, on a stand-in RubixData carrying exactly the fields the function reads."""
    from types import SimpleNamespace as NS
    for f in ("helpers", "generic_models", "dust_baseclasses", "extinction_models"):
        refshim.load(f"rubix/spectra/dust/{f}.py")
    de = refshim.load("rubix/spectra/dust/dust_extinction.py")
    em = sys.modules["rubix.spectra.dust.extinction_models"]
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    o = {}
    mu = f64(x["wave"]) / 1e4
    o["cardelli89_axav"] = em.Cardelli89(Rv=3.1)(mu)
    o["cardelli89_axav_rv45"] = em.Cardelli89(Rv=4.5)(mu)
    o["gordon23_axav"] = em.Gordon23(Rv=3.1)(mu)
    o["gordon23_axav_rv25"] = em.Gordon23(Rv=2.5)(mu)
    log_oh = np.linspace(7.0, 9.5, 26)
    for model in ("power law slope free", "broken power law fit"):
        for xco in ("MW", "Z"):
            o[f"dtg_{model.split()[0]}_{xco}"] = de.calculate_dust_to_gas_ratio(log_oh, model, xco)
    o["cell_extinction"] = de.calculate_extinction(f64(x["gas_mass"]), 3.5)
    for model in ("Cardelli89", "Gordon23"):
        for spectra, tag in ((np.ones_like(x["spectra"]), "factor"), (x["spectra"], "spectra")):
            rd = NS(gas=NS(coords=f64(x["gas_coords"])[None], pixel_assignment=x["gas_pixel"][None],
                           metals=f64(x["gas_metals"])[None], mass=f64(x["gas_mass"])[None]),
                    stars=NS(coords=f64(x["star_coords"])[None], pixel_assignment=x["star_pixel"][None],
                             mass=np.ones((1, len(x["star_pixel"]))), spectra=spectra[None]))
            cfg = {"ssp": {"dust": {"extinction_model": model, "Rv": 3.1, "dust_grain_density": 3.5}}}
            o[f"{model}_{tag}"] = np.asarray(de.apply_spaxel_extinction(cfg, rd, f64(x["wave"]), int(x["S"]),
                                                                        f64(x["spaxel_area"])))[0]
    return {k: np.asarray(v) for k, v in o.items()}


def thin(name, cube):
    cube = np.asarray(cube, dtype=np.float64)
    return {name + "_every4th": cube[:, :, ::4].copy(), name + "_spectrum": cube.sum(axis=(0, 1)),
            name + "_image": cube.sum(axis=2)}


def check():
    """Re-run the reference source and compare with the committed fixtures bit for bit (exit status 1 on a mismatch)."""
    m = reference_modules()
    st = np.load(os.path.join(OUT, "ref_numpy_stages.npz"))
    cu = np.load(os.path.join(OUT, "ref_numpy_cube.npz"))
    bad = []
    x = stage_inputs()
    for k, v in run_stages(m, x).items():
        if "out_" + k in st.files and not np.array_equal(np.asarray(v), st["out_" + k],
                                                         equal_nan=np.asarray(v).dtype.kind == "f"):
            bad.append(k)
    xc = cube_inputs()
    bad += [k for k, v in xc.items() if not np.array_equal(v, cu["in_" + k])]
    bad += [k for k, v in run_cube(m, xc).items() if not np.array_equal(v, cu["out_" + k])]
    bad += [k for k, v in thin("cube_via_core_closures", run_core_closures(xc)[0]).items()
            if not k.endswith("every4th") and not np.array_equal(v, cu["out_" + k])]
    bad += [k for k, v in run_cosmology().items() if not np.array_equal(v, st["cosmo_" + k])]
    bad += [k for k, v in run_boundary(x).items() if not np.array_equal(v, st["boundary_" + k])]
    bad += [k for k, v in run_telescopes().items() if not np.array_equal(v, st["telescope_" + k])]
    bad += [k for k, v in run_pipeline_orders().items() if not np.array_equal(v, st["pipeline_" + k])]
    bad += [k for k, v in run_data_classes().items() if not np.array_equal(v, st["data_" + k])]
    bad += [k for k, v in run_config_data().items() if not np.array_equal(v, st["config_" + k])]
    bad += [k for k, v in run_store_fits().items() if not np.array_equal(v, st["fits_" + k])]
    du = np.load(os.path.join(OUT, "ref_numpy_dust.npz"))
    xd = dust_inputs()
    bad += [k for k, v in xd.items() if not np.array_equal(v, du["in_" + k])]
    bad += [k for k, v in run_dust(xd).items() if not np.array_equal(v, du["out_" + k])]
    print("reference source vs committed fixtures:", "identical" if not bad else f"MISMATCH in {bad}")
    return 1 if bad else 0


def main():
    if "--check" in sys.argv:
        sys.exit(check())
    m = reference_modules()
    x = stage_inputs()
    x.pop("psf_odd")
    o = run_stages(m, x)
    np.savez_compressed(os.path.join(OUT, "ref_numpy_stages.npz"), **{"in_" + k: v for k, v in x.items()
                                                                      if k not in ("lam_ssp", "wave")},
                        **{"out_" + k: v for k, v in o.items() if k not in ("diff", "lam_z")},
                        **{"cosmo_" + k: v for k, v in run_cosmology().items()},
                        **{"boundary_" + k: v for k, v in run_boundary(x).items()},
                        **{"telescope_" + k: v for k, v in run_telescopes().items()},
                        **{"pipeline_" + k: v for k, v in run_pipeline_orders().items()},
                        **{"data_" + k: v for k, v in run_data_classes().items()},
                        **{"config_" + k: v for k, v in run_config_data().items()},
                        **{"fits_" + k: v for k, v in run_store_fits().items()})
    xc = cube_inputs()
    oc = run_cube(m, xc)
    via, shape = run_core_closures(xc)
    oc.update({k: v for k, v in thin("cube_via_core_closures", via).items() if not k.endswith("every4th")})
    oc["core_closures_spectra_shape"] = np.array(shape)
    np.savez_compressed(os.path.join(OUT, "ref_numpy_cube.npz"), **{"in_" + k: v for k, v in xc.items()},
                        **{"out_" + k: v for k, v in oc.items()})
    xd = dust_inputs()
    np.savez_compressed(os.path.join(OUT, "ref_numpy_dust.npz"), **{"in_" + k: v for k, v in xd.items()},
                        **{"out_" + k: v for k, v in run_dust(xd).items()})
    for f in ("ref_numpy_stages.npz", "ref_numpy_cube.npz", "ref_numpy_dust.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
