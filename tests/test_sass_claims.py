"""The SASS of the shipped library against what DESIGN.md says about it (CPU; `cuobjdump -sass`, about 13 s).

Static mnemonic counts per kernel (the same parsing as tools/sass_summary.py, whose output is committed as
profiles/r02_sass_v20.txt): which kernels use packed FFMA2 and 128-bit loads, where the TMA bulk copies (UBLKCP) and
mbarrier waits (SYNCS) are, that the radix pass ranks with MATCH.ANY, and -- the determinism claims -- that no kernel on
the cube path adds into the cube with RED / ATOM (the work queue's two ATOMs aside)."""

import collections
import os
import re
import shutil
import subprocess

import pytest

from rubix_b200 import _lib


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    out = subprocess.run(["cuobjdump", "-sass", _lib.SO_PATH], capture_output=True, text=True, timeout=600).stdout
    kern, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kern[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            kern[name][m.group(1)] += 1
    names = list(kern)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    table = {}
    for n, d in zip(names, dem):
        if "rbx::" in d:
            table[re.sub(r"\(.*", "", d).replace("void ", "")] = kern[n]
    assert len(table) > 60
    return table


def count(c, pattern):
    return sum(v for k, v in c.items() if re.match(pattern, k))


def kernels(table, prefix):
    got = {k: v for k, v in table.items() if k.startswith("rbx::" + prefix)}
    assert got, prefix
    return got


def test_cube_kernels(sass):
    warp = kernels(sass, "fused_cube_warp_kernel<")
    assert len(warp) >= 8                                     # linear / cubic x one warp / pairs x chunk sizes x layouts
    for name, c in warp.items():
        assert count(c, r"^FFMA2") > 0, name                  # packed f32x2 math
        assert count(c, r"^LDG\.E\.128") > 0, name            # template rows from the window tables: 16-byte loads
        assert count(c, r"^RED") == 0 and count(c, r"^ATOM") <= 2, name     # the work queue only: no cube atomics
        assert count(c, r"^UBLKCP") == 0, name                # rows go to registers, not through TMA
        assert count(c, r"^BAR") >= 1, name                   # turns of a warp pair over named barriers
    group = kernels(sass, "fused_cube_kernel<")
    assert any(count(c, r"^UBLKCP") > 0 for c in group.values())            # lookup tables staged with TMA bulk copies
    for name, c in group.items():
        assert count(c, r"^RED") == 0, name
    for name in ("reduce_partials_kernel", "segment_sum_sorted_kernel"):
        c = kernels(sass, name)["rbx::" + name]
        assert count(c, r"^RED") == 0 and count(c, r"^ATOM") == 0, name     # fixed-order sums
    assert count(kernels(sass, "segment_sum_kernel")["rbx::segment_sum_kernel"], r"^RED") >= 1   # the atomic stage form


def test_sort_and_prep_kernels(sass):
    radix = kernels(sass, "radix_pass_kernel")["rbx::radix_pass_kernel"]
    assert count(radix, r"^MATCH") >= 16 and count(radix, r"^RED") == 0    # one match-any vote per item of a thread
    prep = kernels(sass, "prep_kernel")["rbx::prep_kernel"]
    assert count(prep, r"^STG\.E\.128") >= 1                               # records written as 16-byte stores
    keys = kernels(sass, "spaxel_keys_kernel")["rbx::spaxel_keys_kernel"]
    assert count(keys, r"^ATOM") >= 1                                      # shared-memory digit histograms


def test_psf_lsf_march_kernel_uses_tma_bulk_copies(sass):
    march = kernels(sass, "psf_lsf_march_kernel<")
    assert len(march) >= 30
    for name, c in march.items():
        assert count(c, r"^UBLKCP") >= 2 and count(c, r"^SYNCS") >= 1, name   # cp.async.bulk + mbarrier waits
        assert count(c, r"^RED") == 0 and count(c, r"^ATOM") == 0, name
    assert any(count(c, r"^FFMA2") > 0 for c in march.values())              # column pairs through FFMA2


def test_register_budgets_behind_the_occupancy_claims():
    """cuobjdump -res-usage: the defaults of the linear cube kernel (two warps per cell array, 12 warps per SM) need
    <= 168 registers and no stack (65536 / (12 x 32) = 170); prep_kernel <= 64 (four 256-thread blocks per SM);
    radix_pass_kernel <= 80 (three blocks per SM, its launch bound)."""
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-res-usage", _lib.SO_PATH], capture_output=True, text=True, timeout=300).stdout
    usage, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            usage[name] = tuple(int(v) for v in m.groups())
    pick = lambda frag: {k: v for k, v in usage.items() if frag in k}
    linear_pairs = pick("fused_cube_warp_kernelILi0ELb1ELi384E")            # method linear, pairs, 384-cell arrays
    assert len(linear_pairs) == 2                                           # channel-order and transposed cell layout
    for name, (reg, stack, _) in linear_pairs.items():
        assert reg <= 168 and stack == 0, (name, reg, stack)
    (reg, _, _), = pick("rbx11prep_kernel").values()
    assert reg <= 64
    (reg, _, shared), = pick("rbx17radix_pass_kernel").values()
    assert reg <= 80 and shared <= 48 * 1024
