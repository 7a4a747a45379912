"""ctypes binding of ``librubix_b200.so`` (the C ABI declared in include/rubix_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises, it never
falls back to a CPU or PyTorch implementation.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "librubix_b200.so")
CSRC = os.path.join(_HERE, "csrc")

EXPORTED = [
    "rbx_last_error", "rbx_version", "rbx_launch_count",
    "rbx_plan_create", "rbx_plan_destroy", "rbx_plan_dims",
    "rbx_spaxel_assign", "rbx_filter_particles", "rbx_filter_and_assign",
    "rbx_ssp_lookup", "rbx_scale_by_mass", "rbx_doppler_resample", "rbx_segment_sum",
    "rbx_build_cube_workspace_bytes", "rbx_build_cube", "rbx_assign_build_cube",
    "rbx_convolve_psf", "rbx_convolve_lsf", "rbx_psf_lsf", "rbx_psf_lsf_taps", "rbx_psf_lsf_taps_pitched",
    "rbx_gaussian_psf_kernel", "rbx_gaussian_lsf_kernel",
    "rbx_pipeline_host",
    "rbx_rotate_galaxy", "rbx_rotate_galaxy_workspace_bytes",
    "rbx_apply_noise", "rbx_apply_noise_workspace_bytes", "rbx_noise_samples",
    "rbx_dust_av", "rbx_dust_av_workspace_bytes", "rbx_apply_extinction",
    "rbx_build_cube_dusty", "rbx_build_cube_dusty_workspace_bytes",
    "rbx_dusty_moments", "rbx_dusty_bins", "rbx_dusty_combine",
    "rbx_profile_enable", "rbx_profile_fused",
    "rbx_set_option", "rbx_get_option",
    "rbx_assign_build_cube_packed", "rbx_build_cube_status", "rbx_slab_geometry", "rbx_assign_build_cube_slabs",
    "rbx_pipeline_host_packed",
    "rbx_comm_unique_id", "rbx_comm_init", "rbx_comm_destroy", "rbx_comm_info",
    "rbx_reduce_cube", "rbx_allreduce_cube", "rbx_reduce_scatter_cube", "rbx_allgather_cube", "rbx_allreduce_f64",
    "rbx_rotate_moments", "rbx_rotate_apply",
    "rbx_build_cube_cell_layout", "rbx_build_cube_host",
    "rbx_sort_by_spaxel", "rbx_sort_by_spaxel_workspace_bytes", "rbx_segment_sum_sorted",
]

RBX_OK = 0
RBX_ERR_UNSUPPORTED = -4

_lib = None


class RubixB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rubix_b200 error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building librubix_b200.so failed")
    return SO_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C rubix_b200/csrc`). There is no CPU fallback."
        )
    L = C.CDLL(SO_PATH)
    vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
    L.rbx_last_error.restype = C.c_char_p
    L.rbx_last_error.argtypes = []
    L.rbx_version.restype = i32
    L.rbx_launch_count.restype = i64
    sigs = {
        "rbx_plan_create": [C.POINTER(vp), vp, i32, vp, i32, vp, i32, vp, vp, i32, f64, i32, i32, vp],
        "rbx_plan_destroy": [vp],
        "rbx_plan_dims": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
        "rbx_spaxel_assign": [vp, i64, vp, i32, vp, vp, vp],
        "rbx_filter_particles": [vp, i64, vp, i32, vp, vp, vp, vp, vp],
        "rbx_filter_and_assign": [vp, i64, vp, i32, vp, vp, vp, vp, vp, vp],
        "rbx_ssp_lookup": [vp, vp, vp, i64, vp, vp],
        "rbx_scale_by_mass": [vp, vp, i64, i32, vp, vp],
        "rbx_doppler_resample": [vp, vp, vp, i64, vp, vp],
        "rbx_segment_sum": [vp, vp, i64, i32, i32, vp, i32, vp],
        "rbx_build_cube": [vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, sz, vp],
        "rbx_assign_build_cube": [vp, vp, vp, i32, i32, vp, vp, vp, vp, i64, i32, vp, vp, vp, sz, vp],
        "rbx_convolve_psf": [vp, vp, i32, i32, i32, vp, i32, i32, vp],
        "rbx_convolve_lsf": [vp, vp, i64, i32, vp, i32, i32, vp],
        "rbx_psf_lsf": [vp, vp, i32, i32, i32, vp, i32, i32, vp, i32, i32, vp],
        "rbx_psf_lsf_taps": [vp, vp, i32, i32, i32, vp, i32, i32, vp, i32, i32, vp],
        "rbx_psf_lsf_taps_pitched": [vp, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp, i32, i32, vp],
        "rbx_gaussian_psf_kernel": [i32, i32, f32, vp, vp],
        "rbx_gaussian_lsf_kernel": [f32, f32, i32, vp, vp],
        "rbx_pipeline_host": [vp, vp, vp, vp, vp, vp, i64, vp, i32, i32, i32, vp, i32, i32, vp, i32, i32, vp, vp],
    }
    sigs["rbx_rotate_galaxy"] = [vp, vp, vp, i64, f32, vp, vp, vp, vp, vp, sz, vp]
    sigs["rbx_apply_noise"] = [vp, vp, i32, i32, i32, f32, i32, C.c_uint32, C.c_uint32, vp, sz, vp]
    sigs["rbx_noise_samples"] = [vp, vp, i64, i32, C.c_uint32, C.c_uint32, vp]
    sigs["rbx_dust_av"] = [vp, vp, vp, vp, i32, i64, vp, vp, i64, i32, vp, f32, f32, vp, vp, vp, sz, vp]
    sigs["rbx_apply_extinction"] = [vp, vp, vp, i64, i32, vp, vp]
    sigs["rbx_build_cube_dusty"] = [vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, sz, vp]
    sigs["rbx_dusty_bins"] = [vp, vp, vp, vp, vp, vp, i64, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp]
    sigs["rbx_dusty_combine"] = [vp, i32, i32, f32, f32, vp, i32, vp, vp]
    sigs["rbx_set_option"] = [C.c_char_p, i64]
    sigs["rbx_get_option"] = [C.c_char_p, C.POINTER(i64)]
    sigs["rbx_assign_build_cube_packed"] = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, i64, i32, vp, vp, vp, sz, vp]
    sigs["rbx_build_cube_status"] = [vp, i64, i32, vp, C.POINTER(i32), C.POINTER(i32), vp]
    sigs["rbx_build_cube_cell_layout"] = [vp, i64, i32, vp, C.POINTER(i32), vp]
    sigs["rbx_slab_geometry"] = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    sigs["rbx_assign_build_cube_slabs"] = [vp, vp, vp, i32, i32, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp, sz, vp]
    sigs["rbx_pipeline_host_packed"] = [vp, vp, vp, vp, vp, vp, vp, i64, vp, i32, i32, i32, vp, i32, i32, vp, i32, i32,
                                        vp, vp]
    sigs["rbx_build_cube_host"] = [vp, vp, vp, vp, vp, vp, i64, vp, i32, i32, i32, i32, i32, vp, vp]
    sigs["rbx_sort_by_spaxel"] = [vp, i64, i32, vp, vp, vp, vp, sz, vp]
    sigs["rbx_segment_sum_sorted"] = [vp, vp, vp, i32, i32, vp, vp]
    sigs["rbx_comm_unique_id"] = [vp]
    sigs["rbx_comm_init"] = [C.POINTER(vp), vp, i32, i32]
    sigs["rbx_comm_destroy"] = [vp]
    sigs["rbx_comm_info"] = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    sigs["rbx_reduce_cube"] = [vp, vp, vp, i64, i32, vp]
    sigs["rbx_allreduce_cube"] = [vp, vp, vp, i64, vp]
    sigs["rbx_reduce_scatter_cube"] = [vp, vp, vp, i64, vp]
    sigs["rbx_allgather_cube"] = [vp, vp, vp, i64, vp]
    sigs["rbx_allreduce_f64"] = [vp, vp, vp, i64, vp]
    sigs["rbx_rotate_moments"] = [vp, vp, i64, f32, i32, vp, vp, sz, vp]
    sigs["rbx_rotate_apply"] = [vp, vp, i64, vp, vp, vp, vp, vp, vp]
    L.rbx_dusty_moments.argtypes = []
    L.rbx_dusty_moments.restype = i32
    L.rbx_dust_av_workspace_bytes.argtypes = [i64, i32]
    L.rbx_dust_av_workspace_bytes.restype = sz
    L.rbx_build_cube_dusty_workspace_bytes.argtypes = [i64]
    L.rbx_build_cube_dusty_workspace_bytes.restype = sz
    L.rbx_apply_noise_workspace_bytes.argtypes = [i32, i32]
    L.rbx_apply_noise_workspace_bytes.restype = sz
    L.rbx_sort_by_spaxel_workspace_bytes.argtypes = [i64, i32]
    L.rbx_sort_by_spaxel_workspace_bytes.restype = sz
    L.rbx_rotate_galaxy_workspace_bytes.argtypes = []
    L.rbx_rotate_galaxy_workspace_bytes.restype = sz
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    L.rbx_profile_enable.argtypes = [i32]
    L.rbx_profile_enable.restype = i32
    L.rbx_profile_fused.argtypes = [C.POINTER(f64), C.POINTER(i64), i32]
    L.rbx_profile_fused.restype = i32
    L.rbx_build_cube_workspace_bytes.argtypes = [vp, i64, i32]
    L.rbx_build_cube_workspace_bytes.restype = sz
    _lib = L
    return L


def check(code: int) -> None:
    if code != RBX_OK:
        raise RubixB200Error(code, lib().rbx_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(lib().rbx_launch_count())


def set_option(name: str, value: int) -> None:
    """Tuning / test switch of the library (include/rubix_b200.h: rbx_set_option); value < 0 = the library's own choice."""
    check(lib().rbx_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = C.c_int64()
    check(lib().rbx_get_option(name.encode(), C.byref(v)))
    return int(v.value)
