"""Run the REFERENCE'S OWN SOURCE FILES for the path on numpy -- test infrastructure, used by tools/make_ref_golden.py.

rubix is pure Python on jax / jaxtyping / beartype / equinox; none of them is installable in this image, so the package
cannot be imported.  The functions on the hot path, however, only use the numpy-compatible subset of ``jax.numpy``
(plus ``vmap``, ``jax.ops.segment_sum``, ``jax.scipy.signal.convolve / convolve2d``, ``Rotation``).  This module puts
minimal stand-ins for those modules into ``sys.modules`` and executes the reference files UNCHANGED, straight from
``/root/reference`` (nothing is copied into the repo), so that their results -- in float64, where numpy and jax agree on
what the formulas mean -- can be frozen as golden vectors for the oracle (``tests/golden/ref_numpy_*.npz``).

What this is not: jax arithmetic.  ``jnp.interp`` is numpy's (double precision), float32 rounding orders are numpy's,
``interpax`` (the SSP lookup, a1) is absent and is NOT stood in for.  The vectors pin the LOGIC of a0 and a2 - a7
(masks, index clipping, diff0, the flux-conserving scale, nan_to_num, segment ids, convolution alignment, kernel
normalisation, the inertia tensor with its index-0 padding) to the reference's own code.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import scipy.signal
import yaml
from scipy.spatial.transform import Rotation as _ScipyRotation

REF = os.environ.get("RUBIX_REFERENCE", "/root/reference")


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Set:
            @staticmethod
            def set(v):
                out = arr.copy()
                out[idx] = v
                return out.view(_Arr)

        return _Set


class _Arr(np.ndarray):
    """ndarray with jax's functional ``x.at[idx].set(v)``."""

    @property
    def at(self):
        return _At(self)


def _where(cond, x=None, y=None, size=None, fill_value=0):
    if x is None and y is None:
        idx = np.nonzero(np.asarray(cond))
        if size is None:
            return idx
        # jnp.where(mask, size=N): indices padded with fill_value (0) up to the static size
        return tuple(np.concatenate([i[:size], np.full(max(size - len(i), 0), fill_value, dtype=i.dtype)]) for i in idx)
    return np.where(cond, x, y)


def _vmap(fn, in_axes=0, out_axes=0):
    """Loop form of jax.vmap for int / None in_axes and a single array output stacked on axis 0."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[a if ax is None else np.take(a, k, axis=ax) for a, ax in zip(args, axes)]) for k in range(n)]
        return np.stack(outs, axis=out_axes)
    return mapped


def _pmap(fn):
    """jax.pmap as a loop over the leading (device) axis of every positional and keyword argument."""
    def mapped(*args, **kw):
        n = len(args[0]) if args else len(next(iter(kw.values())))
        return np.stack([fn(*[a[i] for a in args], **{k: v[i] for k, v in kw.items()}) for i in range(n)])
    return mapped


def _segment_sum(data, segment_ids, num_segments):
    data, ids = np.asarray(data), np.asarray(segment_ids)
    out = np.zeros((num_segments,) + data.shape[1:], dtype=data.dtype)
    if (ids < 0).any():                               # never produced by the reference (ids are clipped to >= 0)
        raise ValueError("negative segment ids: not part of what this stand-in covers")
    ok = ids < num_segments                           # "values outside [0, num_segments) are dropped"
    np.add.at(out, ids[ok], data[ok])
    return out


def _interp(x, xp, fp, left=None, right=None, period=None):
    """jnp.interp.  Without "extrapolate" it is numpy's; with it, the formula of jax's implementation (the documented
    behaviour: linear continuation of the first / last segment) -- numpy has no such option."""
    if period is not None:
        raise NotImplementedError("periodic interp is not used on the path")
    if left != "extrapolate" and right != "extrapolate":
        return np.interp(x, xp, fp, left=left, right=right)
    x, xp, fp = np.asarray(x), np.asarray(xp), np.asarray(fp)
    i = np.clip(np.searchsorted(xp, x, side="right"), 1, len(xp) - 1)
    df, dx, delta = fp[i] - fp[i - 1], xp[i] - xp[i - 1], x - xp[i - 1]
    dx0 = np.abs(dx) <= np.spacing(np.finfo(xp.dtype).eps)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        f = np.where(dx0, fp[i - 1], fp[i - 1] + (delta / np.where(dx0, 1, dx)) * df)
    if left != "extrapolate":
        f = np.where(x < xp[0], fp[0] if left is None else left, f)
    if right != "extrapolate":
        f = np.where(x > xp[-1], fp[-1] if right is None else right, f)
    return f


def _scan(f, init, xs):
    carry, ys = init, []
    for x in (zip(*xs) if isinstance(xs, tuple) else xs):     # a tuple of arrays is scanned element-wise
        carry, y = f(carry, x)
        ys.append(y)
    return carry, (None if all(y is None for y in ys) else np.stack(ys))


def _sort_key_val(keys, values):
    order = np.argsort(keys, kind="stable")           # lax.sort_key_val is a stable sort by key
    return np.asarray(keys)[order], np.asarray(values)[order]


class _EqxModule:
    """equinox.Module, as far as the extinction models need it: subclasses are dataclasses."""

    def __init_subclass__(cls, **kw):
        import dataclasses
        super().__init_subclass__(**kw)
        dataclasses.dataclass(cls)


def _eqx_field(converter=None, static=False, default=None, **kw):
    import dataclasses
    return dataclasses.field(default=default)


class _Rotation:
    """jax.scipy.spatial.transform.Rotation, the part alignment.py uses (scipy's has the same conventions)."""

    def __init__(self, r):
        self.r = r

    @classmethod
    def from_euler(cls, seq, angles, degrees=False):
        return cls(_ScipyRotation.from_euler(seq, angles, degrees=degrees))

    def __mul__(self, other):
        return _Rotation(self.r * other.r)

    def as_matrix(self):
        return self.r.as_matrix()


class _Subscriptable:
    def __class_getitem__(cls, item):
        return cls

    def __getitem__(self, item):
        return self


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Stand-ins for jax, jaxtyping, beartype and the few rubix modules the path's files import at module level."""
    global _installed
    if _installed:
        return
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "_rbx_shim", False):
        raise RuntimeError("a real jax is imported: run the reference directly instead of the shim")
    jnp = _module("jax.numpy")
    for k in dir(np):
        if not k.startswith("_"):
            setattr(jnp, k, getattr(np, k))
    jnp.where = _where
    jnp.zeros = lambda *a, **k: np.zeros(*a, **k).view(_Arr)
    jnp.array = lambda *a, **k: np.array(*a, **k)
    jnp.ndarray = np.ndarray
    jnp.interp = _interp
    sig = _module("jax.scipy.signal", convolve=scipy.signal.convolve, convolve2d=scipy.signal.convolve2d)
    tr = _module("jax.scipy.spatial.transform", Rotation=_Rotation)
    sp = _module("jax.scipy.spatial", transform=tr)
    jsp = _module("jax.scipy", signal=sig, spatial=sp)
    ops = _module("jax.ops", segment_sum=_segment_sum)
    rnd = _module("jax.random")
    lax = _module("jax.lax", scan=_scan, sort_key_val=_sort_key_val, exp=np.exp,
                  cond=lambda pred, t, f, operand=None: t(operand) if pred else f(operand))
    import functools
    tree = _module("jax.tree_util", Partial=functools.partial)
    _module("jax", numpy=jnp, scipy=jsp, ops=ops, random=rnd, lax=lax, tree_util=tree, vmap=_vmap, pmap=_pmap,
            jit=lambda f, **k: f, Array=np.ndarray, _rbx_shim=True)
    _module("equinox", Module=_EqxModule, AbstractVar=_Subscriptable, field=_eqx_field, filter_jit=lambda c: c)
    ident = lambda *a, **k: (a[0] if a and callable(a[0]) and not k else (lambda f: f))
    names = {k: _Subscriptable for k in ("Array", "Float", "Int", "Bool", "PyTree", "Shaped", "Num")}
    _module("jaxtyping", jaxtyped=lambda *a, **k: (lambda f: f), **names)
    _module("beartype", beartype=ident)
    cfg = yaml.safe_load(open(os.path.join(REF, "rubix", "config", "rubix_config.yml")))
    _module("rubix", config=cfg, __path__=[])
    _module("rubix.cosmology", __path__=[])
    _module("rubix.cosmology.base", BaseCosmology=object)
    import logging
    _module("rubix.core", __path__=[])
    _module("rubix.core.data", RubixData=object, StarsData=object, GasData=object)
    _module("rubix.logger", get_logger=lambda *a, **k: logging.getLogger("rubix-refshim"))
    for pkg in ("rubix.spectra", "rubix.spectra.dust", "rubix.telescope", "rubix.telescope.psf", "rubix.telescope.lsf",
                "rubix.telescope.noise", "rubix.galaxy"):
        _module(pkg, __path__=[])
    _installed = True


def load(relpath: str):
    """Execute one reference source file (e.g. ``rubix/spectra/ifu.py``) from REF under the stand-ins."""
    install()
    name = relpath[:-3].replace("/", ".")
    if name in sys.modules and getattr(sys.modules[name], "__file__", None):
        return sys.modules[name]
    path = os.path.join(REF, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = name.rsplit(".", 1)[0]
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "rubix"))
