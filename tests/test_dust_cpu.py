"""CPU tests of the dust variant: the oracle's restatement of rubix/spectra/dust against the
known answers the reference holds (tests/test_dust_classes.py, tests/test_dust_extinction.py), the host-side
curves of rubix_b200.dust against the oracle, and the ``get_extinction`` mirror's error behaviour."""

import copy

import numpy as np
import pytest

from oracle import rubix_oracle as orc
from rubix_b200 import dust


# ---- generic models: the reference's known answers (tests/test_dust_classes.py) -------------------------------
def test_drude1d_known_answer():   # tests/test_dust_classes.py:40-47
    x = np.array([1.0, 2.0, 3.0])
    want = np.array([1.0, 0.30769232, 0.12328766])
    assert np.allclose(orc._drude1d(x, 1.0, 1.0, 1.0), want, rtol=1e-6)
    assert np.allclose(dust._drude1d(x.astype(np.float32), 1.0, 1.0, 1.0), want, rtol=1e-6)


def test_modified_drude_known_answer():   # tests/test_dust_classes.py:60-68
    x = np.array([1.0, 2.0, 3.0])
    want = np.array([1.0, 0.30769232, 0.12328766])
    assert np.allclose(orc._modified_drude(x, 1.0, 1.0, 1.0, 0.0), want, rtol=1e-6)
    assert np.allclose(dust._modified_drude(x.astype(np.float32), 1.0, 1.0, 1.0, 0.0), want, rtol=1e-6)


def test_fm90_known_answer():   # tests/test_dust_classes.py:71-81
    x = np.array([4.0, 5.0, 6.0])
    want = np.array([4.1879544, 5.723751, 4.7574277])
    assert np.allclose(orc._fm90(x, 0.10, 0.70, 3.23, 0.41, 4.59, 0.95), want, rtol=1e-6)
    assert np.allclose(dust._fm90(x.astype(np.float32), 0.10, 0.70, 3.23, 0.41, 4.59, 0.95), want, rtol=2e-6)


def test_smoothstep_is_the_cubic_hermite_step():   # helpers.py: N = 1 -> 3x^2 - 2x^3
    x = np.linspace(-0.5, 1.5, 41)
    t = np.clip(x, 0, 1)
    assert np.allclose(orc._smoothstep(x, 0.0, 1.0), 3 * t ** 2 - 2 * t ** 3)


def test_cardelli89_range_and_anchor():   # tests/test_dust_classes.py:112-123
    wave = np.array([0.5, 1.0, 2.0, 3.0, 5.0, 8.0, 10.0])
    r = orc.cardelli89(wave, 3.1)
    assert r.shape == wave.shape and np.all(r >= 0) and np.all(r <= 10)
    # CCM89's normalisation: a(1.82) = 1, b(1.82) = 0 for every R(V)
    for rv in (2.0, 3.1, 5.5):
        assert abs(orc.cardelli89(np.array([1.82]), rv)[0] - 1.0) < 1e-12
    # the branch the MUSE band falls into as the reference calls it (microns passed unconverted)
    assert np.allclose(orc.cardelli89(np.array([0.5]), 3.1), (0.574 - 0.527 / 3.1) * 0.5 ** 1.61)


def test_gordon23_range_and_rv_independence_at_v():   # tests/test_dust_classes.py:154-166
    wave = np.array([0.1, 0.3, 0.5, 1.0, 2.0, 5.0, 10.0, 20.0, 30.0])
    r = orc.gordon23(wave, 3.1)
    assert r.shape == wave.shape and np.all(r >= 0) and np.all(r <= 10)
    # at R(V) = 3.1 the b term drops out: a + b (1/Rv - 1/3.1) = a
    assert np.allclose(orc.gordon23(wave, 3.1), orc.gordon23(wave, 3.1 + 1e-15), rtol=1e-9)
    # A(V)/A(V) ~ 1 at 0.55 micron
    assert abs(orc.gordon23(np.array([0.55]), 3.1)[0] - 1.0) < 0.03


@pytest.mark.parametrize("model", ["Cardelli89", "Gordon23"])
@pytest.mark.parametrize("rv", [2.5, 3.1, 4.5])
def test_host_curves_match_oracle(model, rv, muse_wave):
    a = dust.extinction_curve(model, muse_wave, rv)
    b = orc.EXTINCTION_MODELS[model](muse_wave.astype(np.float64) / 1e4, rv)
    assert a.dtype == np.float32 and a.shape == muse_wave.shape
    assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()
    wide = np.array([0.0912, 0.1, 0.3, 0.31, 0.33, 0.5, 0.9, 0.95, 1.0, 1.1, 2.0, 3.3, 5.0, 5.9, 6.5, 8.0, 9.0, 10.0, 20.0, 30.0]) * 1e4
    a = dust.extinction_curve(model, wide, rv)
    b = orc.EXTINCTION_MODELS[model](wide.astype(np.float32).astype(np.float64) / 1e4, rv)
    assert np.abs(a - b).max() <= 5e-6 * np.abs(b).max()


def test_extinguish_is_power_of_ten():   # tests/test_dust_classes.py:136-151
    wave = np.array([5000.0, 6000.0, 7000.0], dtype=np.float32)
    av = np.array([0.0, 3.1])
    ext = orc.extinguish(wave, av, "Cardelli89", 3.1)
    want = np.power(10.0, -0.4 * orc.cardelli89(wave.astype(np.float64) / 1e4, 3.1)[None, :] * av[:, None])
    assert np.allclose(ext, want, rtol=1e-12)
    assert np.all(ext[0] == 1.0)


# ---- dust-to-gas ratio and the A_V constant (tests/test_dust_extinction.py:156-190) ---------------------------
@pytest.mark.parametrize("xco", ["MW", "Z"])
@pytest.mark.parametrize("model", ["power law slope free", "broken power law fit"])
def test_dust_to_gas_ratio(model, xco):
    x = np.array([7.5, 8.0, 8.5])
    r = orc.calculate_dust_to_gas_ratio(x, model, xco)
    assert r.shape == (3,) and np.all(r >= 0)
    a_h, al_h, a_l, al_l, xt = dust.dust_to_gas_parameters(model, xco)
    want = np.where(x > xt, 1 / 10 ** (a_h + al_h * (8.69 - x)), 1 / 10 ** (a_l + al_l * (8.69 - x)))
    assert np.allclose(r, want, rtol=1e-6)
    assert np.all(np.diff(r) > 0)   # more metals, more dust


def test_dust_to_gas_fixed_slope_not_implemented():
    with pytest.raises(NotImplementedError):
        orc.calculate_dust_to_gas_ratio(np.array([8.0]), "power law slope fixed", "MW")
    with pytest.raises(NotImplementedError):
        dust.dust_to_gas_parameters("power law slope fixed", "Z")


def test_extinction_constant_matches_formula():
    c = dust.extinction_constant(3.5)
    want = 3 * np.pi * (1.989e33 / 3.08568e21 ** 2) / (0.4 * np.log(10) * 5448e-8 * 3.5)
    assert abs(c - want) <= 1e-12 * want
    assert abs(orc.dust_extinction_constant(3.5) - want) <= 1e-12 * want


# ---- the per-spaxel column (dust_extinction.py:240-337) --------------------------------------------------------
def test_reference_mock_case():   # tests/test_dust_extinction.py:36-75: shape (n_star, n_wave), all >= 0
    gas_coords = np.array([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]], dtype=np.float32)
    metals = np.array([[0.01, 0.02, 0.03, 0.04, 0.05], [0.06, 0.07, 0.08, 0.09, 0.1]], dtype=np.float32)
    spectra = np.array([[1.0, 2.0], [3.0, 4.0]])
    cfg = {"dust_grain_density": 3.0, "extinction_model": "Cardelli89", "Rv": 3.1,
           "dust_to_gas_model": "power law slope free", "Xco": "MW"}
    for model in ("Cardelli89", "Gordon23"):
        cfg["extinction_model"] = model
        out, av = orc.apply_spaxel_extinction(spectra, np.array([5000.0, 6000.0]), gas_coords[:, 2], np.array([0, 1]),
                                              np.array([1.0, 2.0]), metals, gas_coords[:, 2], np.array([0, 1]), 2, 1.0, cfg)
        assert out.shape == (2, 2) and np.all(out >= 0) and np.all(out <= spectra)
        assert np.all(av >= 0)
    with pytest.raises(ValueError, match="is not available"):
        cfg["extinction_model"] = "Nope"
        orc.apply_spaxel_extinction(spectra, np.array([5000.0, 6000.0]), gas_coords[:, 2], np.array([0, 1]),
                                    np.array([1.0, 2.0]), metals, gas_coords[:, 2], np.array([0, 1]), 2, 1.0, cfg)


def test_stars_av_is_the_cumulative_column_in_front_of_the_star():
    rng = np.random.default_rng(42)
    ng, ns, S = 400, 120, 6
    gz = rng.normal(0, 1, ng).astype(np.float32)
    gp = rng.integers(0, S, ng)
    sz = rng.normal(0, 1.5, ns).astype(np.float32)
    sp = rng.integers(0, S, ns)
    ext = rng.uniform(0, 1, ng)
    av = orc.stars_av(gz, gp, ext, sz, sp, S)
    for k in range(ns):
        m = gp == sp[k]
        o = np.argsort(gz[m], kind="stable")
        z, c = gz[m][o].astype(np.float64), np.cumsum(ext[m][o])
        x = float(sz[k])
        if z[0] <= x <= z[-1]:
            want = np.interp(x, z, c)
        elif x > z[-1]:
            want = c[-1]          # behind all the gas of the spaxel: the whole column
        else:
            want = c[0]           # in front of it: the table's left neighbour is a far cell with value 0
        assert abs(av[k] - want) <= 1e-9 * max(1.0, abs(want)), (k, av[k], want)
    assert np.all(av >= 0)


def test_stars_av_without_far_cells_extrapolates_left():
    # one spaxel only: nothing sits at z * 1e30, so left="extrapolate" runs through the first two cells
    gz = np.array([0.0 + 1.0, 2.0, 3.0], dtype=np.float32)
    ext = np.array([1.0, 1.0, 1.0])
    av = orc.stars_av(gz, np.zeros(3, int), ext, np.array([0.0, 1.5, 10.0], dtype=np.float32), np.zeros(3, int), 1)
    assert np.allclose(av, [0.0, 1.5, 3.0])


def test_stars_outside_every_spaxel_keep_zero():
    gz = np.array([-1.0, 1.0], dtype=np.float32)
    av = orc.stars_av(gz, np.array([0, 0]), np.array([1.0, 1.0]), np.array([0.5, 0.5], dtype=np.float32), np.array([0, -1]), 2)
    assert av[1] == 0.0 and av[0] > 0


# ---- the get_extinction mirror (rubix/core/dust.py:15-65, tests/test_dust_extinction.py:209-228) ----------------
CONFIG = {
    "pipeline": {"name": "calc_dusty_ifu"},
    "logger": {"log_level": "WARNING", "log_file_path": None,
               "format": "%(asctime)s - %(name)s - %(levelname)s - %(message)s"},
    "telescope": {"name": "MUSE", "psf": {"name": "gaussian", "size": 5, "sigma": 0.6}, "lsf": {"sigma": 0.5}},
    "cosmology": {"name": "PLANCK15"},
    "galaxy": {"dist_z": 0.1},
    "ssp": {"template": {"name": "BruzualCharlot2003"},
            "dust": {"extinction_model": "Cardelli89", "dust_to_gas_ratio": 0.01, "dust_to_metals_ratio": 0.4,
                     "dust_grain_density": 3.5, "Rv": 3.1}},
}


def test_get_extinction_errors():
    from rubix_b200.core import get_extinction
    cfg = copy.deepcopy(CONFIG)
    del cfg["ssp"]["dust"]
    with pytest.raises(ValueError, match="Dust configuration not found in config file."):
        get_extinction(cfg)
    cfg = copy.deepcopy(CONFIG)
    del cfg["ssp"]["dust"]["extinction_model"]
    with pytest.raises(ValueError, match="Extinction model not found in dust configuration."):
        get_extinction(cfg)
    fn = get_extinction(copy.deepcopy(CONFIG))
    assert fn.__name__ == "calculate_extinction"   # the YAML node name (pipeline_config.yml:102-106)


def test_dusty_pipeline_order():
    from rubix_b200.core.pipeline import order_by_depends_on
    from rubix_b200.utils import get_pipeline_config
    order = order_by_depends_on(get_pipeline_config("calc_dusty_ifu"))
    i = order.index("calculate_extinction")
    assert order[i - 1] == "doppler_shift_and_resampling" and order[i + 1] == "calculate_datacube"
    assert len(order) == 12


def test_unknown_extinction_model_message():
    with pytest.raises(ValueError, match="Extinction model 'Nope' is not available"):
        dust.extinction_curve("Nope", np.array([5000.0]), 3.1)


def test_prepare_input_loads_and_centres_gas(monkeypatch, tmp_path):
    """rubix/core/data.py:541-600: every stored gas attribute is loaded, coordinates are centred on the subhalo
    centre, and with a subset the gas indices are drawn (seed 42) from the length the STAR arrays have at that point --
    already cut to subset_size -- so the gas keeps a permutation of its first subset_size cells (the reference's own
    prepare_input run from source gives exactly this: tests/test_oracle_vs_reference_source.py)."""
    from rubix_b200.core import pipeline as pl
    rng = np.random.default_rng(0)
    ns, ng = 50, 80
    stars = dict(coords=rng.normal(10, 1, (ns, 3)).astype(np.float32), velocity=rng.normal(0, 1, (ns, 3)).astype(np.float32),
                 mass=np.ones(ns, np.float32), metallicity=np.full(ns, 0.01, np.float32), age=np.full(ns, 5.0, np.float32))
    gas = dict(coords=rng.normal(10, 1, (ng, 3)).astype(np.float32), velocity=np.zeros((ng, 3), np.float32),
               mass=np.arange(ng, dtype=np.float32), metals=rng.uniform(0, 1, (ng, 9)).astype(np.float32),
               density=np.ones(ng, np.float32))
    raw = {"particle_data": {"stars": stars, "gas": gas}, "redshift": 0.1,
           "subhalo_center": np.array([10.0, 10.0, 10.0], np.float32), "subhalo_halfmassrad_stars": 2.0}
    monkeypatch.setattr(pl, "load_rubix_galaxy", lambda path, types: raw)
    cfg = copy.deepcopy(CONFIG)
    cfg["output_path"] = str(tmp_path)
    cfg["data"] = {"args": {"particle_type": ["stars", "gas"]}, "subset": {"use_subset": False}}
    rd = pl.prepare_input(cfg)
    assert np.allclose(rd.gas.coords, gas["coords"] - 10.0) and rd.gas.metals.shape == (ng, 9)
    assert np.array_equal(rd.gas.mass, gas["mass"]) and rd.gas.density is not None
    cfg["data"]["subset"] = {"use_subset": True, "subset_size": 20}
    rd = pl.prepare_input(cfg)
    np.random.seed(42)
    idx = np.random.choice(np.arange(ns), size=20, replace=False)
    assert len(rd.stars.mass) == 20 and np.array_equal(rd.stars.coords, (stars["coords"] - 10.0)[idx])
    np.random.seed(42)
    perm = np.random.choice(np.arange(20), size=20, replace=False)
    assert np.array_equal(rd.gas.mass, gas["mass"][perm])


# ---- the reference's class interface (tests/test_dust_classes.py:112-166) ---------------------------------------
def test_extinction_model_classes():
    wave = np.array([0.5, 1.0, 2.0, 3.0, 5.0, 8.0, 10.0], dtype=np.float32)
    model = dust.Cardelli89(Rv=3.1)
    r = model.evaluate(wave)
    assert r.shape == wave.shape and np.all(r >= 0) and np.all(r <= 10)          # :112-123
    assert np.array_equal(model(wave), r)
    with pytest.raises(ValueError, match="neither Av or Ebv passed, one of them is required!"):   # :126-133
        model.extinguish(wave)
    got = model.extinguish(wave, Ebv=1.0)                                        # :136-151: Av = Rv * Ebv
    assert np.allclose(got, np.power(10.0, -0.4 * r.astype(np.float64) * 3.1), rtol=2e-6)
    assert np.allclose(model.extinguish(wave, Av=3.1), got)
    g = dust.Gordon23(Rv=3.1)
    wave = np.array([0.1, 0.3, 0.5, 1.0, 2.0, 5.0, 10.0, 20.0, 30.0], dtype=np.float32)
    r = g.evaluate(wave)
    assert r.shape == wave.shape and np.all(r >= 0) and np.all(r <= 10)          # :154-166
    assert set(dust.Rv_model_classes) == set(dust.RV_MODELS)


def test_get_extinction_errors_on_the_reference_minimal_configs():   # tests/test_dust_extinction.py:209-227
    from rubix_b200.core import get_extinction
    with pytest.raises(ValueError, match="Dust configuration not found in config file."):
        get_extinction({"ssp": {}, "galaxy": {"dist_z": 0.1}})
    with pytest.raises(ValueError, match="Extinction model not found in dust configuration."):
        get_extinction({"ssp": {"dust": {}}, "galaxy": {"dist_z": 0.1}})


def test_unknown_model_raises_when_the_stage_runs():   # tests/test_dust_extinction.py:192-206
    from rubix_b200.core import RubixData, get_extinction
    cfg = copy.deepcopy(CONFIG)
    cfg["ssp"]["dust"]["extinction_model"] = "InvalidModel"
    fn = get_extinction(cfg)   # the factory accepts it, like the reference's
    with pytest.raises(ValueError, match="Extinction model 'InvalidModel' is not available. Choose from"):
        fn(RubixData())
