"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rubix_b200.h declares,
and reports errors (not crashes) when no device is present.  No compute calls here."""

import ctypes as C
import os
import re

import pytest

from rubix_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "rubix_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rbx_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rubix_b200.h but not exported"
    assert declared == set(_lib.EXPORTED)


def test_version_and_error_string(lib):
    assert lib.rbx_version() >= 100
    assert isinstance(lib.rbx_last_error(), bytes)


def test_argument_validation_without_device(lib):
    # null plan / bad shapes are rejected before any CUDA call
    assert lib.rbx_ssp_lookup(None, None, None, 0, None, None) != 0
    assert b"null plan" in lib.rbx_last_error()
    assert lib.rbx_spaxel_assign(None, 5, None, 1, None, None, None) != 0
    assert lib.rbx_convolve_lsf(None, None, 1, 1, None, 1, 0, None) != 0
    assert lib.rbx_build_cube_workspace_bytes(None, 10, 25) == 0
    # rbx_sort_by_spaxel: sizes checked first, the workspace query needs no device
    assert lib.rbx_sort_by_spaxel(None, -1, 625, None, None, None, None, 0, None) != 0
    assert lib.rbx_sort_by_spaxel(None, 10, 0, None, None, None, None, 0, None) != 0
    assert lib.rbx_sort_by_spaxel(None, 10, 625, None, None, None, None, 0, None) != 0
    assert b"null pointer" in lib.rbx_last_error()
    ws = lib.rbx_sort_by_spaxel_workspace_bytes(10**6, 625)
    assert 3 * 4 * 10**6 <= ws <= 4 * 4 * 10**6        # three n-word arrays + the tile states of two passes
    h = C.c_void_p()
    rc = lib.rbx_plan_create(C.byref(h), None, 2, None, 2, None, 2, None, None, 1, 0.1, 0, 2, None)
    assert rc == -1  # RBX_ERR_INVALID_ARGUMENT


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_option_table(lib):
    """rbx_set_option / rbx_get_option: the switches the launch path reads instead of getenv()."""
    assert _lib.get_option("psub") == -1
    _lib.set_option("psub", 64)
    assert _lib.get_option("psub") == 64
    _lib.set_option("psub", -5)          # any negative value = the library's own choice
    assert _lib.get_option("psub") == -1
    with pytest.raises(_lib.RubixB200Error, match="unknown option"):
        _lib.set_option("no_such_switch", 1)
    # nothing under csrc/ reads the environment on a launch path: getenv appears once, in the option seeding
    csrc = os.path.join(ROOT, "rubix_b200", "csrc")
    hits = [(f, i) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".cc"))
            for i, line in enumerate(open(os.path.join(csrc, f)), 1) if "getenv(" in line and not line.lstrip().startswith("//")]
    assert [f for f, _ in hits] == ["options.cu"], hits


def test_slab_geometry_and_comm_argument_checks(lib):
    w, ws = C.c_int(), C.c_int()
    assert lib.rbx_slab_geometry(3721, 8, 12, C.byref(w), C.byref(ws)) == 0
    assert (w.value, ws.value) == (466, 490)             # ceil(3721 / 8), + 2 x 12 halo channels
    assert lib.rbx_slab_geometry(3721, 0, 12, C.byref(w), C.byref(ws)) != 0
    # the exchange entry points validate before touching NCCL or CUDA
    assert lib.rbx_reduce_cube(None, None, None, 0, 0, None) != 0
    assert lib.rbx_reduce_scatter_cube(None, None, None, 0, None) != 0
    assert lib.rbx_comm_init(None, None, 0, 1) != 0
    assert lib.rbx_build_cube_status(None, 0, 25, None, None, None, None) != 0


def test_plain_c_client_compiles_links_and_gets_error_codes(lib, tmp_path):
    """The boundary is a C ABI: tests/abi/abi_client.c includes the header as C99 (-Wall -Wextra -Werror -pedantic),
    links librubix_b200.so and checks versions, geometry helpers, the option table and argument validation."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_client")
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi", "abi_client.c"), "-o", exe, "-L", libdir, "-lrubix_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "abi_client: ok" in res.stdout, res.stdout + res.stderr


def test_c_host_of_the_whole_path_fails_loudly_without_a_device(lib, tmp_path, bc03, muse_wave):
    """tests/abi/c_host_pipeline.c drives rbx_plan_create + rbx_pipeline_host from C with host buffers only.  Without a
    GPU it must stop with the library's error (no CPU fallback); on a GPU box tests/test_gpu_reference_vectors.py
    compares its cube with the reference-source vector."""
    import shutil
    import subprocess
    import numpy as np
    from helpers import build_c_host, write_c_host_inputs
    from rubix_b200 import synthetic
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a device is present: the GPU test runs the C host")
    except ImportError:
        pass
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = build_c_host(tmp_path)
    write_c_host_inputs(str(tmp_path), bc03, muse_wave, synthetic.bench_g(100), synthetic.spatial_edges(5), 5,
                        np.full((3, 3), 1 / 9, np.float32), np.ones(1, np.float32), "linear")
    res = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 1 and "rbx_plan_create" in res.stderr, res.stdout + res.stderr
    assert not (tmp_path / "cube.f32").exists()
