"""Regenerate the committed fixtures under tests/golden/ from the reference's own data files.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tools/make_golden.py

Outputs
-------
rubix_b200/templates/bc03lr_f32.npz   The BC03lr SSP template exactly as the reference's loader hands it to
                              interp2d: every dataset cast to float32, no log transform, no unit
                              change (rubix/spectra/ssp/grid.py:323-331, rubix_config.yml:152-176).
tests/golden/muse_wave.npy    The ``wave`` dataset of notebooks/data/dummy_datacube.h5: a cube the
                              reference itself wrote, pinning telescope.wave_seq for MUSE bit-exactly.
tests/golden/tng50_subset.npz 4096 star particles of tests/output/rubix_galaxy.h5 (TNG50 subhalo 14),
                              drawn with the reference's own subset rule (np.random.seed(42) +
                              np.random.choice, rubix/core/data.py:565-571), cast to float32 and
                              centred like rubix/galaxy/alignment.py:36-57.  Velocities are converted
                              from the file's kpc/s to km/s so the Doppler shifts are realistic.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rubix_b200.h5lite import H5File  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
KPC_IN_KM = 3.0856775814913673e16


def main():
    global OUT
    check = "--check" in sys.argv
    committed = OUT
    if check:   # regenerate into a scratch directory and compare array by array with the committed fixtures
        import tempfile
        OUT = tempfile.mkdtemp()
        os.makedirs(os.path.join(OUT, "templates"), exist_ok=True)
    os.makedirs(OUT, exist_ok=True)

    with H5File(f"{REF}/rubix/spectra/ssp/templates/BC03lr.h5") as f:
        tpl = {k: f[k].read().astype(np.float32) for k in ("age", "metallicity", "wavelength", "flux")}
    tpl_path = (os.path.join(OUT, "templates", "bc03lr_f32.npz") if check
                else os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
    np.savez_compressed(tpl_path, **tpl)

    with H5File(f"{REF}/notebooks/data/dummy_datacube.h5") as f:
        np.save(os.path.join(OUT, "muse_wave.npy"), f["wave"].read())

    with H5File(f"{REF}/tests/output/rubix_galaxy.h5") as f:
        center = f["galaxy/center"].read().astype(np.float32)
        stars = {k: f[f"particles/stars/{k}"].read() for k in
                 ("coords", "velocity", "mass", "metallicity", "age")}
    n = len(stars["coords"])
    np.random.seed(42)
    idx = np.random.choice(np.arange(n), size=4096, replace=False)
    coords = stars["coords"].astype(np.float32)
    vel = stars["velocity"].astype(np.float32)
    mask = np.linalg.norm(coords - center, axis=1) < 10
    central_v = np.median(vel[mask], axis=0)
    sub = {
        "coords": (coords - center)[idx],
        "velocity": ((vel - central_v)[idx].astype(np.float64) * KPC_IN_KM).astype(np.float32),
        "mass": stars["mass"].astype(np.float32)[idx],
        "metallicity": stars["metallicity"].astype(np.float32)[idx],
        "age": stars["age"].astype(np.float32)[idx],
    }
    np.savez_compressed(os.path.join(OUT, "tng50_subset.npz"), **sub)
    if check:
        bad = []
        pairs = [(tpl_path, os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz")),
                 (os.path.join(OUT, "tng50_subset.npz"), os.path.join(committed, "tng50_subset.npz"))]
        for new, old in pairs:
            a, b = np.load(new), np.load(old)
            bad += [f"{os.path.basename(old)}:{k}" for k in set(a.files) | set(b.files)
                    if k not in a.files or k not in b.files or not np.array_equal(a[k], b[k])]
        if not np.array_equal(np.load(os.path.join(OUT, "muse_wave.npy")), np.load(os.path.join(committed, "muse_wave.npy"))):
            bad.append("muse_wave.npy")
        print("reference data files vs committed fixtures:", "identical" if not bad else f"MISMATCH in {bad}")
        sys.exit(1 if bad else 0)
    for fn in sorted(os.listdir(OUT)):
        if not os.path.isdir(os.path.join(OUT, fn)):
            print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
