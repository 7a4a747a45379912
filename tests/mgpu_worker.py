"""Worker of tests/test_gpu_scale.py::test_two_ranks_nccl_against_oracle (launched with torchrun, one rank per GPU).

Every rank bins its contiguous particle shard (rubix/core/data.py:471-482) with the CUDA path; the partial cubes are
summed through the C ABI's NCCL calls (rbx_comm_*), and rank 0 compares with the oracle on the UNSHARDED input:
  A. MUSE 25 x 25: rbx_reduce_cube onto rank 0, PSF + LSF there           (rubix/core/ifu.py:324-333)
  B. 60 x 60 spaxels, slab-major partial cubes: rbx_reduce_scatter_cube, PSF + LSF per wavelength slab with the
     12-channel halo, slabs gathered for the comparison                    
  C. rotate_galaxy under sharding: every rank must apply the SAME rotation (moments summed over ranks)
torch.distributed (gloo) only carries the 128-byte NCCL id and gathers results for the check.
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("gloo")
    from oracle import c_oracle
    from oracle import rubix_oracle as orc
    from helpers import cube_close
    from rubix_b200 import ops, parallel, synthetic

    tpl = dict(np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz")))
    wave = np.load(os.path.join(ROOT, "tests", "golden", "muse_wave.npy"))
    W = len(wave)
    pk, lk = orc.gaussian_kernel_2d(5, 5, 0.6), orc.lsf_kernel(0.5, 1.25)
    plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1, method="linear")
    comm = ops.Comm.from_torch_distributed()
    if rank == 0:
        print("NCCL version", comm.nccl_version(), flush=True)

    def oracle(d, edges, S):
        ref = c_oracle.particles_to_cube(d["coords"], d["velocity"], d["mass"], d["metallicity"], d["age"], edges, S,
                                         tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, 0.1,
                                         method="linear", dtype=np.float64, n_threads=8)
        return ref, orc.apply_lsf(orc.apply_psf(ref, pk.astype(np.float64)), 0.5, 1.25)

    # ---- A: reduce onto rank 0 -------------------------------------------------------------------------
    S = 25
    edges = synthetic.spatial_edges(S)
    full = synthetic.bench_g(300_001, seed=21)            # odd count: the last shard is shorter
    mine = parallel.shard_particles(full, rank, world)
    cube = ops.assign_build_cube(plan, mine["coords"], edges, mine["velocity"], mine["mass"], mine["metallicity"],
                                 mine["age"], S)
    comm.reduce(cube, root=0)
    if rank == 0:
        ref, refc = oracle(full, edges, S)
        cube_close(cube.cpu().numpy(), ref, "2 ranks: reduced cube")
        cube_close(ops.psf_lsf(cube, pk, lk).cpu().numpy(), refc, "2 ranks: reduced cube + PSF + LSF")

    # ---- B: slab-major partial cubes, reduce-scatter, PSF + LSF per slab ------------------------------------
    S = 60
    edges = synthetic.spatial_edges(S)
    full = synthetic.bench_g(200_000, seed=22)
    full["coords"] *= np.float32(1.6)
    mine = parallel.shard_particles(full, rank, world)
    halo = 12
    wslab, ws = ops.slab_geometry(W, world, halo)
    slabs = ops.assign_build_cube_slabs(plan, mine["coords"], edges, mine["velocity"], mine["mass"],
                                        mine["metallicity"], mine["age"], S, world, halo)
    own = torch.empty((S * S, ws), dtype=torch.float32, device="cuda")
    comm.reduce_scatter(slabs, own)
    out = ops.psf_lsf_own_slab(own, S, W, rank, world, pk, lk, halo).contiguous()
    pad = torch.zeros((S, S, wslab), dtype=torch.float32, device="cuda")
    pad[:, :, :out.shape[2]] = out
    parts = [torch.empty_like(pad).cpu() for _ in range(world)]
    dist.all_gather(parts, pad.cpu())
    if rank == 0:
        _, refc = oracle(full, edges, S)
        got = np.concatenate([p.numpy() for p in parts], axis=2)[:, :, :W]
        cube_close(got, refc, "2 ranks: reduce-scatter + slab PSF + LSF")

    # ---- C: rotate_galaxy of a sharded galaxy = rotate_galaxy of the whole galaxy ------------------------------
    full = synthetic.bench_g(100_000, seed=23)
    full["coords"][:, 2] *= np.float32(0.2)
    mine = parallel.shard_particles(full, rank, world)
    c_all, v_all, R_all = ops.rotate_galaxy(full["coords"], full["velocity"], full["mass"], 1.5, 30.0, 40.0, 50.0)
    c_my, v_my, R_my = ops.rotate_galaxy(mine["coords"], mine["velocity"], mine["mass"], 1.5, 30.0, 40.0, 50.0, comm=comm)
    lo, hi = parallel.shard_range(100_000, rank, world)
    assert float((R_my - R_all).abs().max()) <= 2e-5, "sharded rotate_galaxy: rotation differs from the unsharded one"
    assert float((c_my - c_all[lo:hi]).abs().max()) <= 2e-4 * float(c_all.abs().max())
    assert float((v_my - v_all[lo:hi]).abs().max()) <= 2e-4 * float(v_all.abs().max())

    ok = torch.ones(1)
    dist.all_reduce(ok)
    comm.close()
    if rank == 0:
        print("MGPU OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
