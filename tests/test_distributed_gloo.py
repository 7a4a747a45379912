"""World-size-2 test (gloo, CPU) of the multi-GPU host logic: contiguous particle shards
(rubix/core/data.py:471-482), per-rank partial cubes, one sum-reduce (rubix/core/ifu.py:333), and the
wavelength-slab split used for the PSF/LSF stage.  The partial cubes come from the CPU oracle here
(the checker); on the GPU box the same plumbing carries the CUDA cubes (bench.py --gpus N)."""

import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle
        from oracle import rubix_oracle as orc
        from rubix_b200 import parallel, synthetic
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        tpl = np.load(os.path.join(root, "rubix_b200", "templates", "bc03lr_f32.npz"))
        wave = synthetic.muse_wave()[:400]
        edges = synthetic.spatial_edges(5)
        data = synthetic.bench_g(3001, seed=42)  # odd count: the last shard is shorter
        mine = parallel.shard_particles(data, rank, world)
        cube = c_oracle.particles_to_cube(mine["coords"], mine["velocity"], mine["mass"], mine["metallicity"],
                                          mine["age"], edges, 5, tpl["metallicity"], tpl["age"], tpl["wavelength"],
                                          tpl["flux"], wave, 0.1, method="linear", dtype=np.float64, n_threads=1)
        partial = cube.copy()                       # t shares cube's memory and is summed in place
        t = torch.from_numpy(np.ascontiguousarray(cube))
        parallel.allreduce_cube(t)
        # PSF + LSF sharded by wavelength slab with a +-12 channel halo, then gathered
        pk = orc.gaussian_kernel_2d(5, 5, 0.6).astype(np.float64)
        lo, hi = parallel.wavelength_slab(len(wave), rank, world)
        hlo, hhi = max(lo - 12, 0), min(hi + 12, len(wave))
        slab = orc.apply_lsf(orc.apply_psf(t.numpy()[:, :, hlo:hhi], pk), 0.5, 1.25)[:, :, lo - hlo:lo - hlo + hi - lo]
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, slab))
        # the large-FOV exchange: slab-major partial cubes with halos, summed, every rank keeps its own
        # slab (gloo has no reduce-scatter: all-reduce + slice is the same sum), PSF + LSF on slab + halo
        W = len(wave)
        packed = torch.from_numpy(parallel.slab_pack(partial.reshape(25, W), world, 12))
        dist.all_reduce(packed)
        own = packed[rank].numpy().reshape(5, 5, -1)
        own = parallel.slab_interior(orc.apply_lsf(orc.apply_psf(own, pk), 0.5, 1.25), W, rank, world, 12)
        parts2 = [None] * world
        dist.all_gather_object(parts2, (rank, own))
        if rank == 0:
            full = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])], axis=2)
            full2 = np.concatenate([p[1] for p in sorted(parts2, key=lambda p: p[0])], axis=2)
            np.savez(os.path.join(out_dir, "out.npz"), cube=t.numpy(), conv=full, conv_slabs=full2)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_reduce_and_slab_convolution(tmp_path, bc03):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    from oracle import c_oracle
    from oracle import rubix_oracle as orc
    from rubix_b200 import synthetic
    wave = synthetic.muse_wave()[:400]
    edges = synthetic.spatial_edges(5)
    d = synthetic.bench_g(3001, seed=42)
    ref = c_oracle.particles_to_cube(d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"], edges, 5,
                                     bc03["metallicity"], bc03["age"], bc03["wavelength"], bc03["flux"], wave, 0.1,
                                     method="linear", dtype=np.float64, n_threads=1)
    assert ref.max() > 0
    assert np.abs(got["cube"] - ref).max() <= 1e-12 * ref.max()
    conv = orc.apply_lsf(orc.apply_psf(ref, orc.gaussian_kernel_2d(5, 5, 0.6).astype(np.float64)), 0.5, 1.25)
    # slab-wise LSF with a 12-channel halo is exact (the kernel reaches +-12 channels)
    assert np.abs(got["conv"] - conv).max() <= 1e-12 * conv.max()
    # the same through the slab-major layout (what rbx_assign_build_cube_slabs + rbx_reduce_scatter_cube carry)
    assert got["conv_slabs"].shape == conv.shape
    assert np.abs(got["conv_slabs"] - conv).max() <= 1e-12 * conv.max()


def test_slab_geometry_matches_the_library():
    import ctypes as C
    from rubix_b200 import _lib, parallel
    for W, g, h in ((3721, 8, 12), (3721, 2, 12), (400, 3, 5), (7, 8, 0)):
        a, b = C.c_int(), C.c_int()
        assert _lib.lib().rbx_slab_geometry(W, g, h, C.byref(a), C.byref(b)) == 0
        assert (a.value, b.value) == parallel.slab_geometry(W, g, h)


def _dusty_inputs():
    from rubix_b200 import synthetic
    d = synthetic.bench_g(601, seed=42)
    rng = np.random.default_rng(7)
    ng = 900
    gas = dict(coords=np.stack([rng.normal(0, 1.5, ng), rng.normal(0, 1.5, ng), rng.normal(0, 1.0, ng)], 1).astype(np.float32),
               mass=(rng.uniform(0.5, 2, ng) * 2e5).astype(np.float32),
               metals=rng.uniform(1e-4, 1e-2, (ng, 9)).astype(np.float32))
    gas["metals"][:, 0] = 0.74
    return d, gas


def _dusty_cube(orc, stars, gas, tpl, wave, edges, S):
    """calc_dusty_ifu on the oracle for one set of stars and ALL the gas cells."""
    spix = orc.square_spaxel_assignment(stars["coords"], edges)
    gpix = orc.square_spaxel_assignment(gas["coords"], edges)
    cell = orc.dust_cell_extinction(gas["mass"], gas["metals"], 3.5, 0.145, "broken power law fit", "Z")
    av = orc.stars_av(gas["coords"][:, 2], gpix, cell, stars["coords"][:, 2], spix, S * S)
    ext = orc.extinguish(wave, av, "Cardelli89", 3.1)
    cube, _ = orc.particles_to_cube(stars["coords"], stars["velocity"], stars["mass"], stars["metallicity"], stars["age"],
                                    edges, S, tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1,
                                    method="linear", dtype=np.float64, extinction=ext)
    return cube, av


def _dusty_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import rubix_oracle as orc
        from rubix_b200 import parallel, synthetic
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        tpl = np.load(os.path.join(root, "rubix_b200", "templates", "bc03lr_f32.npz"))
        wave = synthetic.muse_wave()[:300]
        edges = synthetic.spatial_edges(5)
        stars, gas = _dusty_inputs()
        mine = parallel.shard_particles(stars, rank, world)   # stars shard by rank, the gas cells are replicated
        cube, av = _dusty_cube(orc, mine, gas, tpl, wave, edges, 5)
        t = torch.from_numpy(np.ascontiguousarray(cube))
        parallel.allreduce_cube(t)
        parts = [None] * world
        dist.all_gather_object(parts, (rank, av))
        if rank == 0:
            np.savez(os.path.join(out_dir, "dusty.npz"), cube=t.numpy(),
                     av=np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_dusty_variant_shards_stars_and_replicates_gas(tmp_path, bc03):
    """The dusty variant on two ranks: a star's A_V needs all the gas cells of its spaxel, so stars are sharded and
    the gas is replicated; the partial cubes add up to the single-process cube with no further exchange."""
    world = 2
    mp.spawn(_dusty_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "dusty.npz")
    from oracle import rubix_oracle as orc
    from rubix_b200 import synthetic
    stars, gas = _dusty_inputs()
    ref, av = _dusty_cube(orc, stars, gas, bc03, synthetic.muse_wave()[:300], synthetic.spatial_edges(5), 5)
    assert av.max() > 0.1 and ref.max() > 0
    assert np.array_equal(got["av"], av)                       # per-star A_V does not depend on the star sharding
    assert np.abs(got["cube"] - ref).max() <= 1e-12 * ref.max()


def _reference_vector_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle
        from oracle import rubix_oracle as orc
        from rubix_b200 import parallel, synthetic
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        tpl = np.load(os.path.join(root, "rubix_b200", "templates", "bc03lr_f32.npz"))
        fx = np.load(os.path.join(root, "tests", "golden", "ref_numpy_cube.npz"))
        data = {k: fx["in_" + k] for k in ("coords", "velocity", "mass", "metallicity", "age")}
        wave = synthetic.muse_wave()
        mine = parallel.shard_particles(data, rank, world)
        cube = c_oracle.particles_to_cube(mine["coords"], mine["velocity"], mine["mass"], mine["metallicity"],
                                          mine["age"], fx["in_edges"], 7, tpl["metallicity"], tpl["age"],
                                          tpl["wavelength"], tpl["flux"], wave, 0.1, method="linear", dtype=np.float64,
                                          n_threads=1)
        W = len(wave)
        packed = torch.from_numpy(parallel.slab_pack(cube.reshape(49, W), world, 12))
        t = torch.from_numpy(np.ascontiguousarray(cube))
        parallel.allreduce_cube(t)
        dist.all_reduce(packed)
        pk = orc.gaussian_kernel_2d(5, 5, 0.6, dtype=np.float64)   # the vector is float64 throughout
        own = packed[rank].numpy().reshape(7, 7, -1)
        own = parallel.slab_interior(orc.apply_lsf(orc.apply_psf(own, pk), 0.5, 1.25), W, rank, world, 12)
        parts = [None] * world
        dist.all_gather_object(parts, (rank, own, len(mine["mass"])))
        if rank == 0:
            np.savez(os.path.join(out_dir, "refvec.npz"), cube=t.numpy(), counts=np.array([p[2] for p in parts]),
                     conv=np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])], axis=2))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_reproduce_the_reference_closures_over_two_devices(tmp_path):
    """The reference-source vector tests/golden/ref_numpy_cube.npz holds the cube rubix/core/ifu.py's own closures give
    with the particles reshaped to TWO devices (pmap + jnp.sum(axis=0)).  Two gloo ranks, sharded by
    parallel.shard_particles (the same contiguous ceil(n / 2) split), reduced with parallel.allreduce_cube, PSF + LSF
    per wavelength slab of the slab-major exchange: the same cube."""
    world = 2
    mp.spawn(_reference_vector_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "refvec.npz")
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_numpy_cube.npz"))
    n = len(fx["in_mass"])
    assert list(got["counts"]) == [-(-n // 2), n - -(-n // 2)] and fx["out_core_closures_spectra_shape"][1] == -(-n // 2)
    for name, c in (("cube", got["cube"]), ("cube_psf_lsf", got["conv"])):
        for suffix, v in (("_every4th", c[:, :, ::4]), ("_spectrum", c.sum(axis=(0, 1))), ("_image", c.sum(axis=2))):
            ref = fx["out_" + name + suffix]
            assert np.abs(v - ref).max() <= 1e-11 * np.abs(ref).max(), (name, suffix)
    for suffix, v in (("_spectrum", got["cube"].sum(axis=(0, 1))), ("_image", got["cube"].sum(axis=2))):
        ref = fx["out_cube_via_core_closures" + suffix]
        assert np.abs(v - ref).max() <= 1e-11 * np.abs(ref).max()
