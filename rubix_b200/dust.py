"""Host side of the dust-extinction variant (``calc_dusty_ifu``): the per-configuration pieces of
rubix/spectra/dust that do not depend on the particles.

* the A(lambda)/A(V) curves of rubix/spectra/dust/extinction_models.py (``Cardelli89``, ``Gordon23``),
  evaluated once per configuration on the telescope wavelength grid in float32 (the reference runs
  them through jnp with x64 off) -- like the PSF / LSF taps they are kernel *parameters*: the device
  multiplies every star's spectrum by ``10^(-0.4 * axav * Av)`` (rbx_apply_extinction /
  rbx_build_cube_dusty);
* the dust-to-gas fit parameters of Remy-Ruyer et al. 2014 as coded in
  rubix/spectra/dust/dust_extinction.py:41-91 and the A_V-per-surface-density constant of :150-162.

The per-particle work (gas cell A_V, the (pixel, z) sort, the cumulative column, the per-star
interpolation) is CUDA: rubix_b200/csrc/dust.cu.
"""

from __future__ import annotations

import numpy as np

from .config import CONSTANTS, DUST

RV_MODELS = ["Cardelli89", "Gordon23"]   # rubix/spectra/dust/extinction_models.py:10-13

F = np.float32


def dust_to_gas_parameters(model: str, Xco: str):
    """(a_high, alpha_high, a_low, alpha_low, x_transition) for ``1 / 10**(a + alpha * (8.69 - x))``;
    the *high* pair applies for ``x > x_transition`` (dust_extinction.py:41-91)."""
    if model == "power law slope fixed":
        raise NotImplementedError("power law slope fixed not implemented yet.")
    table = {
        ("MW", "power law slope free"): (2.21, 1.62, 2.21, 1.62, -np.inf),
        ("MW", "broken power law fit"): (2.21, 1.00, 0.68, 3.08, 7.96),
        ("Z", "power law slope free"): (2.21, 2.02, 2.21, 2.02, -np.inf),
        ("Z", "broken power law fit"): (2.21, 1.00, 0.96, 3.10, 8.10),
    }
    if (Xco, model) not in table:
        raise ValueError(f"unknown dust-to-gas model {model!r} / Xco {Xco!r}")
    return np.asarray(table[(Xco, model)], dtype=np.float32)


def extinction_constant(dust_grain_density: float, effective_wavelength: float = 5448.0) -> float:
    """dust_extinction.py:150-162: A_V per unit dust surface density in Msun / kpc^2."""
    conv = float(CONSTANTS["MSUN_TO_GRAMS"]) / float(CONSTANTS["KPC_TO_CM"]) ** 2
    return float(3.0 * np.pi * conv / (0.4 * np.log(10.0) * effective_wavelength * 1e-8 * dust_grain_density))


# ---------------------------------------------------------------------------------------------
# extinction curves (float32, the argument exactly as the reference passes it)
# ---------------------------------------------------------------------------------------------
def cardelli89(wave, Rv: float = 3.1) -> np.ndarray:
    """extinction_models.py:104-178.  The reference hands this function ``wavelength / 1e4`` (microns,
    dust_extinction.py:343-345) although the CCM89 polynomials are written in 1/micron; the branches
    are evaluated on that argument as it comes -- parity is with the reference, not with CCM89."""
    w = np.asarray(wave, dtype=F)
    a = np.zeros_like(w)
    b = np.zeros_like(w)
    ir = (w >= F(0.3)) & (w < F(1.1))
    opt = (w >= F(1.1)) & (w < F(3.3))
    nuv = (w >= F(3.3)) & (w <= F(8.0))
    fnuv = (w >= F(5.9)) & (w <= F(8.0))
    fuv = (w > F(8.0)) & (w <= F(10.0))
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        wp = np.power(w, F(1.61))
        a = np.where(ir, F(0.574) * wp, a)
        b = np.where(ir, F(-0.527) * wp, b)
        y = w - F(1.82)
        a = np.where(opt, F(1) + F(0.17699) * y - F(0.50447) * y ** 2 - F(0.02427) * y ** 3 + F(0.72085) * y ** 4
                     + F(0.01979) * y ** 5 - F(0.77530) * y ** 6 + F(0.32999) * y ** 7, a)
        b = np.where(opt, F(1.41338) * y + F(2.28305) * y ** 2 + F(1.07233) * y ** 3 - F(5.38434) * y ** 4
                     - F(0.62251) * y ** 5 + F(5.30260) * y ** 6 - F(2.09002) * y ** 7, b)
        a = np.where(nuv, F(1.752) - F(0.316) * w - F(0.104) / ((w - F(4.67)) ** 2 + F(0.341)), a)
        b = np.where(nuv, F(-3.09) + F(1.825) * w + F(1.206) / ((w - F(4.62)) ** 2 + F(0.263)), b)
        y = w - F(5.9)
        a = np.where(fnuv, a + (F(-0.04473) * y ** 2 - F(0.009779) * y ** 3), a)
        b = np.where(fnuv, b + (F(0.2130) * y ** 2 + F(0.1207) * y ** 3), b)
        y = w - F(8.0)
        a = np.where(fuv, F(-1.073) - F(0.628) * y + F(0.137) * y ** 2 - F(0.070) * y ** 3, a)
        b = np.where(fuv, F(13.670) + F(4.257) * y - F(0.420) * y ** 2 + F(0.374) * y ** 3, b)
    return (a + b / F(Rv)).astype(F)


def _smoothstep(x, x_min, x_max):
    """rubix/spectra/dust/helpers.py ``_smoothstep`` with N = 1."""
    x = np.clip((x - F(x_min)) / F(x_max - x_min), F(0), F(1))
    return ((F(3.0) + F(-2.0) * x) * x ** 2).astype(F)


def _drude1d(x, amplitude, x_0, fwhm):
    """generic_models.py ``Drude1d``."""
    r = F(fwhm / x_0) ** 2
    return F(amplitude) * r / ((x / F(x_0) - F(x_0) / x) ** 2 + r)


def _modified_drude(x, scale, x_o, gamma_o, asym):
    """generic_models.py ``_modified_drude``."""
    gamma = F(2.0 * gamma_o) / (F(1.0) + np.exp(F(asym) * (x - F(x_o))))
    return F(scale) * ((gamma / F(x_o)) ** 2) / ((x / F(x_o) - F(x_o) / x) ** 2 + (gamma / F(x_o)) ** 2)


def _fm90(x, C1, C2, C3, C4, xo, gamma):
    """generic_models.py ``FM90``."""
    e = F(C1) + F(C2) * x
    x2 = x ** 2
    e = e + F(C3) * (x2 / ((x2 - F(xo ** 2)) ** 2 + x2 * F(gamma ** 2)))
    far = x >= F(5.9)
    y = np.where(far, x - F(5.9), F(0))
    return np.where(far, e + F(C4) * (F(0.5392) * y ** 2 + F(0.05644) * y ** 3), e).astype(F)


_G23_IR_A = (0.38526, 1.68467, 0.78791, 4.30578, 4.78338, 0.06652, 9.8434, 2.21205, -0.24703,
             0.0267, 19.58294, 17.0, -0.27)
_G23_OPT_A = (-0.35848, 0.7122, 0.08746, -0.05403, 0.00674, 0.03893, 2.288, 0.243, 0.02965, 2.054, 0.179,
              0.01747, 1.587, 0.243)
_G23_OPT_B = (0.12354, -2.68335, 2.01901, -0.39299, 0.03355, 0.18453, 2.288, 0.243, 0.19728, 2.054, 0.179,
              0.1713, 1.587, 0.243)


def _g23_nirmir(wave, params):
    """``Gordon23.nirmir_intercept`` (extinction_models.py:391-431)."""
    scale, alpha, alpha2, swave, swidth, s1a, s1c, s1f, s1y, s2a, s2c, s2f, s2y = params
    p1 = F(scale) * np.power(wave, F(-alpha))
    ratio = F(swave ** (-alpha) / swave ** (-alpha2))
    p2 = F(scale) * ratio * np.power(wave, F(-alpha2))
    wgt = _smoothstep(wave, swave - swidth / 2, swave + swidth / 2)
    out = p1 * (F(1.0) - wgt) + p2 * wgt
    out = out + _modified_drude(wave, s1a, s1c, s1f, s1y)
    return (out + _modified_drude(wave, s2a, s2c, s2f, s2y)).astype(F)


def _g23_poly_drude(x, p):
    """``compound_polynomial_drude_model`` (extinction_models.py:318-352): Horner polynomial + three Drudes."""
    c0 = F(p[4])
    for i in (3, 2, 1, 0):
        c0 = F(p[i]) + c0 * x
    return (c0 + _drude1d(x, p[5], p[6], p[7]) + _drude1d(x, p[8], p[9], p[10]) + _drude1d(x, p[11], p[12], p[13])).astype(F)


def gordon23(wave, Rv: float = 3.1) -> np.ndarray:
    """extinction_models.py:262-389 (Gordon et al. 2023), ``wave`` in microns."""
    w = np.asarray(wave, dtype=F)
    a = np.zeros_like(w)
    b = np.zeros_like(w)
    ir = (w >= F(1.0)) & (w < F(35.0))
    opt = (w >= F(0.3)) & (w < F(1.1))
    uv = (w >= F(0.09)) & (w <= F(0.3))
    optir = (w >= F(0.9)) & (w <= F(1.1))
    uvopt = (w >= F(0.3)) & (w <= F(0.33))
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        x = F(1) / w
        pl = F(-1.01251) * np.power(w / F(1.0), F(1.06099))   # PowerLaw1d(amplitude=-1.01251, x_0=1, alpha=-1.06099)
        nir = _g23_nirmir(w, _G23_IR_A)
        pa, pb = _g23_poly_drude(x, _G23_OPT_A), _g23_poly_drude(x, _G23_OPT_B)
        a = np.where(ir, nir, a)
        b = np.where(ir, pl, b)
        a = np.where(opt, pa, a)
        b = np.where(opt, pb, b)
        wgt = _smoothstep(w, 0.9, 1.1)
        a = np.where(optir, (F(1.0) - wgt) * pa + wgt * nir, a)
        b = np.where(optir, (F(1.0) - wgt) * pb + wgt * pl, b)
        fa = _fm90(x, 0.81297, 0.2775, 1.06295, 0.11303, 4.60, 0.99)
        fb = _fm90(x, -2.97868, 1.89808, 3.10334, 0.65484, 4.60, 0.99)
        a = np.where(uv, fa, a)
        b = np.where(uv, fb, b)
        wgt = _smoothstep(w, 0.3, 0.33)
        a = np.where(uvopt, (F(1.0) - wgt) * fa + wgt * pa, a)
        b = np.where(uvopt, (F(1.0) - wgt) * fb + wgt * pb, b)
    return (a + b * F(1 / Rv - 1 / 3.1)).astype(F)


Rv_model_dict = {"Cardelli89": cardelli89, "Gordon23": gordon23}


def extinction_curve(model: str, wave_angstrom, Rv: float) -> np.ndarray:
    """A(lambda)/A(V) on the telescope grid, float32: ``ext(wavelength / 1e4)`` (dust_extinction.py:231-232, :345)."""
    if model not in RV_MODELS:
        raise ValueError(f"Extinction model '{model}' is not available. Choose from {RV_MODELS}.")
    wave = np.asarray(wave_angstrom, dtype=F) / F(1e4)
    return np.ascontiguousarray(Rv_model_dict[model](wave, float(Rv)), dtype=F)


def package_defaults() -> dict:
    """rubix_config.yml:145-150 (``ssp.dust`` of the package-level configuration)."""
    return dict(DUST)


# ---------------------------------------------------------------------------------------------
# the reference's class interface (rubix/spectra/dust/dust_baseclasses.py, extinction_models.py), host / numpy:
# ``Cardelli89(Rv=3.1)(wave)``, ``.evaluate(wave)``, ``.extinguish(wave, Av=..., Ebv=...)``
# ---------------------------------------------------------------------------------------------
class BaseExtRvModel:
    """dust_baseclasses.py:72-164: an R(V)-dependent extinction curve A(lambda)/A(V)."""

    _curve = None
    wave_range_l = wave_range_h = Rv_range_l = Rv_range_h = None

    def __init__(self, Rv: float = 3.1):
        self.Rv = float(Rv)

    def evaluate(self, wave) -> np.ndarray:
        return type(self)._curve(wave, self.Rv)

    def __call__(self, wave) -> np.ndarray:
        return self.evaluate(wave)

    def extinguish(self, wave, Av=None, Ebv=None) -> np.ndarray:
        """dust_baseclasses.py:126-164: the fractional extinction ``10 ** (-0.4 * axav * Av)``."""
        axav = self(wave)
        if (Av is None) and (Ebv is None):
            raise ValueError("neither Av or Ebv passed, one of them is required!")
        if Av is None:
            Av = self.Rv * Ebv
        return np.power(F(10.0), F(-0.4) * axav * F(Av)).astype(F)


class Cardelli89(BaseExtRvModel):
    """extinction_models.py:22-178."""
    _curve = staticmethod(cardelli89)
    wave_range_l, wave_range_h, Rv_range_l, Rv_range_h = 0.3, 10.0, 2.0, 6.0


class Gordon23(BaseExtRvModel):
    """extinction_models.py:181-431."""
    _curve = staticmethod(gordon23)
    wave_range_l, wave_range_h, Rv_range_l, Rv_range_h = 0.0912, 32.0, 2.3, 5.6


Rv_model_classes = {"Cardelli89": Cardelli89, "Gordon23": Gordon23}
