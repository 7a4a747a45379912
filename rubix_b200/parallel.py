"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200
box, gloo in CPU tests).

The path is data-parallel over particles with exactly one exchange: the per-rank partial cubes are
summed (the reference's ``jnp.sum(ifu_cubes, axis=0)`` over its device axis, rubix/core/ifu.py:333).
Particles are split into contiguous ranges of ceil(N / world) like the reference's ``reshape_array``
(rubix/core/data.py:471-482); the zero padding of the last range contributes exactly 0.
"""

from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def shard_particles(data: Dict[str, np.ndarray], rank: int, world: int) -> Dict[str, np.ndarray]:
    n = len(next(iter(data.values())))
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in data.items()}


def wavelength_slab(W: int, rank: int, world: int) -> Tuple[int, int]:
    """Channel range [lo, hi) owned by ``rank`` when the PSF / LSF stage is sharded by wavelength."""
    return shard_range(W, rank, world)


def slab_geometry(W: int, nslab: int, halo: int = 12) -> Tuple[int, int]:
    """(wslab, ws) of the slab-major partial cube (``rbx_slab_geometry``): ceil(W / nslab) channels per slab, stored
    with ``halo`` channels of each neighbour."""
    wslab = (W + nslab - 1) // nslab
    return wslab, wslab + 2 * halo


def slab_pack(cube: np.ndarray, nslab: int, halo: int = 12) -> np.ndarray:
    """Host restatement of the layout ``rbx_assign_build_cube_slabs`` writes: (nseg, W) -> (nslab, nseg, ws); slab r
    holds channels [r wslab - halo, (r + 1) wslab + halo), zeros outside [0, W).  One reduce-scatter over the first
    axis then leaves rank r with its summed slab and the halo the LSF needs."""
    nseg, W = cube.shape
    wslab, ws = slab_geometry(W, nslab, halo)
    out = np.zeros((nslab, nseg, ws), dtype=cube.dtype)
    for r in range(nslab):
        lo = r * wslab - halo
        a, b = max(lo, 0), min(lo + ws, W)
        if b > a:
            out[r, :, a - lo:b - lo] = cube[:, a:b]
    return out


def slab_interior(slab: np.ndarray, W: int, rank: int, nslab: int, halo: int = 12) -> np.ndarray:
    """The channels rank ``rank`` owns, cut out of its (.., ws) slab after the PSF / LSF pass."""
    wslab, _ = slab_geometry(W, nslab, halo)
    n_own = max(0, min(wslab, W - rank * wslab))
    return slab[..., halo:halo + n_own]


_comm = None


def get_comm():
    """The process-wide ``ops.Comm`` (NCCL through the C ABI, ``rbx_comm_*``), bootstrapped once over the initialised
    ``torch.distributed`` group (which only carries the 128-byte NCCL id).  None when the job has one rank."""
    global _comm
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    if _comm is None:
        from . import ops
        _comm = ops.Comm.from_torch_distributed()
    return _comm


def close_comm():
    global _comm
    if _comm is not None:
        _comm.close()
        _comm = None


def _on_cuda(t) -> bool:
    return bool(getattr(t, "is_cuda", False))


def allreduce_cube(cube):
    """Sum the partial cubes of all ranks in place: ``rbx_allreduce_cube`` for CUDA tensors; CPU tensors (the gloo
    tests of the host logic) go through torch.distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if _on_cuda(cube):
            get_comm().allreduce(cube)
        else:
            dist.all_reduce(cube, op=dist.ReduceOp.SUM)
    return cube


def reduce_cube(cube, dst: int = 0):
    """Sum the partial cubes onto rank ``dst`` in place (``rbx_reduce_cube``; other ranks' buffers are unchanged)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if _on_cuda(cube):
            get_comm().reduce(cube, root=dst)
        else:
            dist.reduce(cube, dst=dst, op=dist.ReduceOp.SUM)
    return cube
