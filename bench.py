#!/usr/bin/env python
"""bench.py -- particles/s binned into a MUSE datacube (with PSF + LSF) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--particles P] [--weak] [--method linear|cubic] [--spaxels S] [--galaxies G]

A "step" is one pass of the hot path over one synthetic galaxy (bench-G of SURVEY.md section 8(d)):
filter_particles -> spaxel_assignment -> fused SSP lookup / mass scaling / Doppler shift / flux-conserving
resample / cube accumulation -> [sum of the per-rank partial cubes] -> PSF -> LSF.

Default workload = BASELINE.json's target, config 3: ONE galaxy of 10^7 star particles onto the MUSE grid
(25 x 25 x 3721), STRONG scaling: rank r bins the contiguous particle range [r ceil(P/N), (r+1) ceil(P/N))
(rubix/core/data.py:447-487), the partial cubes are summed onto rank 0 with one NCCL reduce through the C ABI
(rbx_reduce_cube; rubix/core/ifu.py:324-333) and rank 0 applies PSF + LSF.  ``--spaxels 150`` is config 4: the
partial cubes are built slab-major, reduce-scattered (rbx_reduce_scatter_cube) and every rank convolves its own
wavelength slab.  ``--particles 1000000`` is config 2; ``--weak`` keeps P particles PER GPU; ``--galaxies G`` is the
survey batch (config 5: G rotated galaxies of P particles per GPU and step, replicas only).

* ``value``     whole-job particles/s with the particle arrays resident in HBM (CUDA events around each step on the
                launch stream, L2 flushed between steps, max over ranks).
* ``e2e``       the same metric through the host-buffer C-ABI call (``rbx_pipeline_host``): pinned host arrays in
                the reference's (n, 3) layout in, host cube out, H2D + D2H inside the timed region.
* ``parity``    the step's CUDA result on a bounded particle sample against the float64 oracle, in the same run
                (N > 1: the sharded, NCCL-reduced result against the oracle on the unsharded sample).
* ``roofline``  the dominant kernel (fused_cube_warp_kernel) timed by CUDA events inside the library on its launch
                stream.  It is bound by FP32 / instruction issue, not by HBM (SURVEY 8d "binding roofline"), so the
                object reports the executed FP32 rate against the MEASURED FMA peak (profiles/fp32_peak.json from
                tools/fma_peak.cu) next to the HBM and issue-slot fractions.
* ``cpu_baseline``  the C restatement of the reference (oracle/rubix_oracle.c, float32, all host threads) on a
                bounded sample of the same workload (N = 1 only).

``--impl reference`` times that CPU restatement as the reference arm (the reference itself is pure Python/JAX and
cannot be installed here: no jax / interpax wheels, no network) on rank 0, with the same ``config``.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particles/s binned into MUSE datacube (SSP lookup + Doppler resample + cube + PSF + LSF)"
PSF = dict(size=5, sigma=0.6)
LSF = dict(sigma=0.5, wave_res=1.25)
REDSHIFT = 0.1
HALO = 12   # LSF half width in channels (rubix/telescope/lsf/lsf.py:12-26: extend_factor 12)


def load_template():
    path = os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz")
    if os.path.exists(path):
        d = np.load(path)
        return {k: d[k] for k in d.files}, "BC03lr (the reference's template as float32 arrays)"
    from rubix_b200.synthetic import synthetic_ssp
    return synthetic_ssp(), "synthetic BC03lr-shaped template"


def host_kernels():
    """PSF / LSF taps exactly as the reference builds them (kernels.py:26-31, lsf.py:12-26), float32."""
    m = PSF["size"]
    x = np.arange(-((m - 1) / 2), ((m - 1) / 2) + 1).astype(np.float32)
    X, Y = np.meshgrid(x, x, indexing="ij")
    v = np.exp(-(X**2 + Y**2) / np.float32(2 * PSF["sigma"] ** 2)).astype(np.float32)
    pk = (v / v.sum(dtype=np.float32)).astype(np.float32)
    wr = LSF["wave_res"]
    xs = np.arange(-12 * wr, 12 * wr + wr, wr).astype(np.float32)
    r = np.exp(np.float32(-0.5) * xs**2 / np.float32(LSF["sigma"] ** 2)).astype(np.float32)
    lk = (r / r.sum(dtype=np.float32)).astype(np.float32)
    return pk, lk


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (NVML; falls back to nvidia-smi)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM)))
            mask = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            for name, bit in self.REASONS:
                if mask & bit:
                    self.reasons.add(name)
            return
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                              "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                             capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 6:
            self.sm.append(float(parts[0]))
            self.max_mhz = float(parts[1])
            for (name, _), val in zip(self.REASONS, parts[2:6]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                self.sample()
            except Exception:
                pass
            time.sleep(0.002 if self.nvml is not None else 0.1)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm)}


# ---- the workload, named identically by both arms -------------------------------------------------------------
def total_particles(args):
    """Star particles one step processes over all ranks."""
    if args.galaxies > 1:
        return args.gpus * args.galaxies * args.particles
    return args.gpus * args.particles if args.weak else args.particles


def scaling_of(args):
    return "weak" if (args.weak or args.galaxies > 1) else "strong"


def slab_mode_for(args):
    return args.gpus > 1 and args.galaxies == 1 and args.spaxels * args.spaxels * 3721 * 4 > (64 << 20)


def config_for(args, tpl_name):
    S, g = args.spaxels, args.gpus
    grid = f"MUSE {S}x{S} spaxels x 3721 channels, BC03 SSP, ssp.method={args.method}, gaussian PSF 5/0.6 + LSF sigma 0.5"
    if args.galaxies > 1:
        workload = (f"survey batch: {args.galaxies} synthetic galaxies per GPU x {args.particles} star particles (bench-G "
                    f"discs), each rotated to its own inclination on the device, {grid}")
        par = f"replicas x{g}: galaxies are independent, no collective"
    else:
        if args.weak:
            workload = f"{args.particles} synthetic star particles per GPU (bench-G), {grid}"
        else:
            workload = (f"one galaxy of {args.particles} synthetic star particles (bench-G) sharded over {g} GPU(s) in "
                        f"contiguous ranges of ceil(P/N), {grid}")
        if g == 1:
            par = "single GPU, no collective"
        elif slab_mode_for(args):
            par = (f"particle-sharded x{g}; slab-major partial cubes, one NCCL reduce-scatter through the C ABI "
                   f"(rbx_reduce_scatter_cube), PSF+LSF per wavelength slab (+-{HALO} channel halo) on every rank")
        else:
            par = (f"particle-sharded x{g}; one NCCL reduce of the 9.3 MB partial cubes through the C ABI "
                   "(rbx_reduce_cube), PSF+LSF on rank 0")
    if fov_scale(args) > 1.0:
        workload += (f"; custom telescope fov {5 * fov_scale(args):.0f} arcsec at MUSE's 0.2 arcsec spaxels, galaxy sigma "
                     f"{fov_scale(args):.0f} kpc")
    return {"workload": workload, "template": tpl_name, "method": args.method, "particles_total": total_particles(args),
            "spaxels": S, "galaxies_per_gpu": args.galaxies, "scaling": scaling_of(args),
            "l2": "flushed (256 MB fill) between timed steps", "parallelism": par}


def fov_scale(args):
    """The large-FOV grid (config 4: fov 30 arcsec instead of MUSE's 5) keeps the spaxel size and widens the field."""
    return args.spaxels / 25.0 if args.spaxels > 60 else 1.0


def galaxy(args, n, seed=42):
    """bench-G; on the large-FOV grid the galaxy is spread over the wider field."""
    from rubix_b200 import synthetic
    d = synthetic.bench_g(n, seed=seed)
    if fov_scale(args) > 1.0:
        d["coords"] *= np.float32(fov_scale(args) / 1.5)
    return d


def spatial_edges(args):
    from rubix_b200 import synthetic
    return synthetic.spatial_edges(args.spaxels, half_aperture=4.7619 * fov_scale(args))


def cpu_arm(data, tpl, wave, edges, S, method, pk, lk, sample, threads, dtype=np.float32):
    """One pass of the reference algorithm (C restatement) over the first ``sample`` particles."""
    from oracle import c_oracle
    sl = slice(0, sample)
    t0 = time.perf_counter()
    cube = c_oracle.particles_to_cube(data["coords"][sl], data["velocity"][sl], data["mass"][sl],
                                      data["metallicity"][sl], data["age"][sl], edges, S, tpl["metallicity"],
                                      tpl["age"], tpl["wavelength"], tpl["flux"], wave, REDSHIFT, method=method,
                                      dtype=dtype, n_threads=threads)
    cube = c_oracle.apply_psf(cube, pk.astype(dtype))
    cube = c_oracle.apply_lsf(cube, lk.astype(dtype))
    return time.perf_counter() - t0, cube


def one_core_rate(data, tpl, wave, edges, S, method, pk, lk, sample):
    """The same pass on ONE host thread over a smaller sample (1 core and all cores are both reported)."""
    m = max(1, min(sample, 40000))
    t, _ = cpu_arm(data, tpl, wave, edges, S, method, pk, lk, m, 1)
    return {"value": m / t, "unit": "particles/s", "cores": 1, "sample": f"first {m} particles, one pass, one thread"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rubix_b200 import synthetic
    tpl, tpl_name = load_template()
    wave = synthetic.muse_wave()
    S = args.spaxels
    edges = spatial_edges(args)
    pk, lk = host_kernels()
    threads = os.cpu_count() or 1
    total = total_particles(args)
    sample = min(total, args.cpu_sample)
    data = galaxy(args, sample, seed=42)   # bench-G draws are i.i.d.: a sample of the workload's particles
    for _ in range(args.warmup):
        cpu_arm(data, tpl, wave, edges, S, args.method, pk, lk, min(sample, 20000), threads)
    times = [cpu_arm(data, tpl, wave, edges, S, args.method, pk, lk, sample, threads)[0] for _ in range(args.steps)]
    t = float(np.mean(times))
    val = sample / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "particles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_for(args, tpl_name),
        "cpu_baseline": {"value": val, "unit": "particles/s", "cores": threads, "kind": "port",
                         "sample": (f"each step = one pass over {sample} of the workload's {total} particles "
                                    f"(PSF+LSF of the full cube included), oracle/rubix_oracle.c float32 on all {threads} "
                                    f"host threads (OpenMP); ms_per_step is the SAMPLE's time -- the work is linear in the "
                                    f"particle count, the whole workload extrapolates to {total / val:.1f} s per step"),
                         "implementation": "C restatement of rubix@dbb4487, not jax: the reference is pure Python/JAX "
                                           "and cannot be installed in this image",
                         "one_core": one_core_rate(data, tpl, wave, edges, S, args.method, pk, lk, sample)},
        "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def measured_peaks():
    """(hbm GB/s, source), fp32 peak record or None."""
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for key in ("hbm_gbs", "hbm_gbs_burst", "hbm_gb_s", "hbm"):
            if isinstance(pk.get(key), (int, float)) and pk[key] > 0:
                peak, src = float(pk[key]), f"measured (MEASURED_PEAKS.json {key})"
                break
    except (OSError, ValueError, AttributeError):
        pass
    fp32 = None
    try:
        fp32 = json.load(open(os.path.join(ROOT, "profiles", "fp32_peak.json")))
    except (OSError, ValueError):
        pass
    return (peak, src), fp32


def ncu_counts(method, n_local, S):
    """Per-particle instruction / flop counts of the cube kernel from the committed ncu --set full captures
    (profiles/fused_ncu.json, written by tools/ncu_summary.py).  The capture of this configuration if there is one,
    else the one with the nearest particle count for the same method and grid (flagged in the output)."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "fused_ncu.json")))
    except (OSError, ValueError):
        return None
    best = None
    for key, rec in table.items():
        parts = key.split("_")
        if len(parts) < 2 or parts[0] != method or not parts[1].isdigit() or not isinstance(rec, dict):
            continue
        s = int(parts[2]) if len(parts) > 2 and parts[2].isdigit() else 25
        if s != S:
            continue
        dist = abs(np.log(int(parts[1]) / max(n_local, 1)))
        if best is None or dist < best[0]:
            best = (dist, key, rec)
    if best is None:
        return None
    rec = dict(best[2])
    rec["ncu_config"] = best[1]
    rec["ncu_config_exact"] = bool(best[0] < 1e-9)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=10_000_000,
                    help="star particles of the galaxy (total over all GPUs; per GPU with --weak; per galaxy with --galaxies)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --particles per GPU")
    ap.add_argument("--method", default="linear", choices=["linear", "cubic"])
    ap.add_argument("--spaxels", type=int, default=25)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--parity-sample", type=int, default=200_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run oracle comparison")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stage", action="store_true", help="skip the timing of the HBM-bound PSF+LSF kernel on a 150x150 cube")
    ap.add_argument("--galaxies", type=int, default=1,
                    help="survey batch (config 5): galaxies per GPU and step, each with its own inclination "
                         "(rotate_galaxy on the device); replicas only, no collective")
    args = ap.parse_args()
    args.galaxies = max(1, args.galaxies)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from rubix_b200 import _lib, ops, parallel, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = parallel.get_comm()   # NCCL through the C ABI (rbx_comm_*); torch.distributed carries the id + barriers

    tpl, tpl_name = load_template()
    wave = synthetic.muse_wave()
    S = args.spaxels
    W = len(wave)
    edges_h = spatial_edges(args)
    pk_h, lk_h = host_kernels()
    G = args.galaxies
    plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, REDSHIFT,
                    method=args.method, direction="z")
    slab_mode = slab_mode_for(args)

    # ---- this rank's particles --------------------------------------------------------------------------------
    if G > 1:
        n = args.particles
        data = None
    elif args.weak:
        n = args.particles
        data = galaxy(args, n, seed=42 + rank)
    else:   # strong scaling: contiguous ranges of ceil(P / world) of ONE galaxy (rubix/core/data.py:447-487)
        full = galaxy(args, args.particles, seed=42)
        data = {k: np.ascontiguousarray(v) for k, v in parallel.shard_particles(full, rank, world).items()}
        n = len(data["mass"])
        del full
    total = total_particles(args)

    edges = ops.dev(edges_h)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2
    if data is not None:
        coords, vel = ops.dev(data["coords"]), ops.dev(data["velocity"])
        mass, met, age = ops.dev(data["mass"]), ops.dev(data["metallicity"]), ops.dev(data["age"])
    wslab = W
    if slab_mode:
        wslab, ws_ch = ops.slab_geometry(W, world, HALO)
        slabs = torch.empty((world, S * S, ws_ch), dtype=torch.float32, device="cuda")
        own = torch.empty((S * S, ws_ch), dtype=torch.float32, device="cuda")
        cube = None
    else:
        cube = torch.empty((S, S, W), dtype=torch.float32, device="cuda")

    gal = []
    if G > 1:   # survey batch: a flattened disc per galaxy, its own Euler angles
        rng = np.random.default_rng(1000 + rank)
        for g in range(G):
            dg = synthetic.bench_g(n, seed=4200 + 64 * rank + g)
            dg["coords"][:, 2] *= 0.2
            gal.append(dict(coords=ops.dev(dg["coords"]), velocity=ops.dev(dg["velocity"]), mass=ops.dev(dg["mass"]),
                            metallicity=ops.dev(dg["metallicity"]), age=ops.dev(dg["age"]),
                            angles=tuple(float(a) for a in rng.uniform(0.0, 180.0, 3)), host=dg))

    def survey_step():
        out = None
        for g in gal:   # independent galaxies: no collective (SURVEY 8e "replicas only")
            c, v, _ = ops.rotate_galaxy(g["coords"], g["velocity"], g["mass"], 1.5, *g["angles"])
            ops.assign_build_cube(plan, c, edges, v, g["mass"], g["metallicity"], g["age"], S, out=cube)
            out = ops.psf_lsf(cube, pk_h, lk_h)
        return out

    def device_step(c, v, m, z, a):
        """One pass over this rank's particles (device tensors) -> rank 0's cube / this rank's slab."""
        if slab_mode:
            ops.assign_build_cube_slabs(plan, c, edges, v, m, z, a, S, world, HALO, out=slabs)
            comm.reduce_scatter(slabs, own)
            return ops.psf_lsf_own_slab(own, S, W, rank, world, pk_h, lk_h, HALO)
        ops.assign_build_cube(plan, c, edges, v, m, z, a, S, out=cube)
        if world > 1:
            comm.reduce(cube, root=0)
        if rank == 0:
            return ops.psf_lsf(cube, pk_h, lk_h)  # host taps, as the reference builds them
        return cube

    def step():
        if G > 1:
            return survey_step()
        return device_step(coords, vel, mass, met, age)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    import ctypes as C
    _lib.lib().rbx_profile_enable(1)
    _lib.lib().rbx_profile_fused(None, None, 1)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.fill_(1.0)  # evict L2 between timed steps (not timed)
        a.record()
        step()
        b.record()
    barrier()
    launches = _lib.launch_count() - launches0
    sampler.stop_flag = True
    mean_ms = C.c_double()
    nl = C.c_int64()
    _lib.lib().rbx_profile_fused(C.byref(mean_ms), C.byref(nl), 1)
    _lib.lib().rbx_profile_enable(0)
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = total / (ms_per_step * 1e-3)
    err_impl = ops.build_cube_status(plan, n, S)

    # ---- e2e: host buffers through the C ABI (right after the timed loop: the GPU and the PCIe link are still in
    # their active power state; the oracle of the parity leg below keeps the GPU idle for a second or two) ---------
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        hcube = torch.empty((S, S, W), dtype=torch.float32).pin_memory() if not slab_mode else None
        if world == 1 and G == 1:
            hp = {k: pin(v) for k, v in data.items()}
            hnp = {k: v.numpy() for k, v in hp.items()}

            def e2e_step():
                out = ops.pipeline_host(plan, hnp["coords"], hnp["velocity"], hnp["mass"], hnp["metallicity"],
                                        hnp["age"], edges_h, S, pk_h, lk_h, out=hcube.numpy())
                return float(out[S // 2, S // 2, 100])  # the host reads the result
            e2e_api = "rbx_pipeline_host (C ABI, pinned host buffers in the reference's (n, 3) layout)"
            d2h = hcube.numel() * 4
            h2d = sum(v.nbytes for v in hnp.values())
        elif G > 1:
            for g in gal:
                g["pin"] = {k: pin(v) for k, v in g["host"].items()}

            def e2e_step():
                r = 0.0
                for g in gal:
                    p = {k: v.cuda(non_blocking=True) for k, v in g["pin"].items()}
                    c, v, _ = ops.rotate_galaxy(p["coords"], p["velocity"], p["mass"], 1.5, *g["angles"])
                    ops.assign_build_cube(plan, c, edges, v, p["mass"], p["metallicity"], p["age"], S, out=cube)
                    hcube.copy_(ops.psf_lsf(cube, pk_h, lk_h), non_blocking=True)
                    torch.cuda.synchronize()
                    r += float(hcube[S // 2, S // 2, 100])
                return r
            e2e_api = ("per galaxy: pinned host arrays -> device, rotate_galaxy + cube + PSF + LSF through the C ABI, "
                       "cube to the host")
            d2h = hcube.numel() * 4 * G
            h2d = sum(v.nbytes for v in gal[0]["host"].values()) * G
        else:
            # N > 1: pinned host shard -> device (async copies), device path, NCCL through the C ABI, result to the host
            hp = {k: pin(v) for k, v in data.items()}
            hnp = {k: v.numpy() for k, v in hp.items()}
            hslab = (torch.empty((S, S, max(0, min(wslab, W - rank * wslab))), dtype=torch.float32).pin_memory()
                     if slab_mode else None)

            def e2e_step():
                # rbx_build_cube_host: the shard goes to the device in ranges on a second stream while the kernels of
                # the previous range run (what rbx_pipeline_host does on one GPU); exchange + PSF + LSF on the device
                if slab_mode:
                    ops.build_cube_host(plan, hnp["coords"], hnp["velocity"], hnp["mass"], hnp["metallicity"], hnp["age"],
                                        edges_h, S, out=slabs, nslab=world, halo=HALO)
                    comm.reduce_scatter(slabs, own)
                    out = ops.psf_lsf_own_slab(own, S, W, rank, world, pk_h, lk_h, HALO)
                else:
                    ops.build_cube_host(plan, hnp["coords"], hnp["velocity"], hnp["mass"], hnp["metallicity"], hnp["age"],
                                        edges_h, S, out=cube)
                    comm.reduce(cube, root=0)
                    out = ops.psf_lsf(cube, pk_h, lk_h) if rank == 0 else cube
                if slab_mode:
                    hslab.copy_(out, non_blocking=True)
                    torch.cuda.synchronize()
                    return float(hslab[S // 2, S // 2, 0])
                if rank == 0:
                    hcube.copy_(out, non_blocking=True)
                torch.cuda.synchronize()
                return float(hcube[S // 2, S // 2, 100]) if rank == 0 else 0.0
            e2e_api = ("rbx_build_cube_host (pinned host shard copied in ranges under the previous range's kernels) + "
                       "rbx_reduce_scatter_cube through the C ABI, PSF+LSF per wavelength slab, each slab to its rank's "
                       "host" if slab_mode else
                       "rbx_build_cube_host (pinned host shard copied in ranges under the previous range's kernels) + "
                       "rbx_reduce_cube through the C ABI, PSF+LSF on rank 0, cube to rank 0's host")
            d2h = (hslab.numel() if slab_mode else hcube.numel()) * 4
            h2d = sum(v.numel() * 4 for v in hp.values())

        def time_e2e(fn):
            for _ in range(4):
                fn()
            barrier()
            e_steps = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(e_steps):
                flush.fill_(1.0)
                torch.cuda.synchronize()
                fn()
            torch.cuda.synchronize()
            # the L2 flush (a 256 MB fill, ~0.1 ms) is inside this wall-clock bracket; negligible vs the PCIe copies
            s = (time.perf_counter() - t0) / e_steps
            te = torch.tensor([s], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item())

        e2e_s = time_e2e(e2e_step)
        h2d += edges_h.nbytes + pk_h.nbytes + lk_h.nbytes
        e2e = {"value": total / e2e_s, "unit": "particles/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "api": e2e_api,
               "bytes_are": "this rank's (rank 0's) copies per step"}
        if world == 1 and G == 1:   # the same call with structure-of-arrays host buffers (24 B / particle over PCIe)
            sx, sy = pin(hnp["coords"][:, 0]).numpy(), pin(hnp["coords"][:, 1]).numpy()
            sv = pin(hnp["velocity"][:, 2]).numpy()

            def packed_step():
                out = ops.pipeline_host_packed(plan, sx, sy, sv, hnp["mass"], hnp["metallicity"], hnp["age"], edges_h, S,
                                               pk_h, lk_h, out=hcube.numpy())
                return float(out[S // 2, S // 2, 100])
            ps = time_e2e(packed_step)
            e2e["packed"] = {"value": total / ps, "ms_per_step": ps * 1e3,
                             "h2d_bytes_per_step": int(24 * n + edges_h.nbytes + pk_h.nbytes + lk_h.nbytes),
                             "api": "rbx_pipeline_host_packed (x, y, line-of-sight velocity, mass, Z, age as separate "
                                    "host arrays: the 24 B / particle the path reads)"}

    # ---- parity: the same step on a bounded sample against the float64 oracle -----------------------------------
    parity = None
    if not args.no_parity:
        P = max(world, min(args.parity_sample, total if G == 1 else n))
        if G > 1:
            g0 = gal[0]
            c, v, _ = ops.rotate_galaxy(g0["coords"][:P], g0["velocity"][:P], g0["mass"][:P], 1.5, *g0["angles"])
            ops.assign_build_cube(plan, c, edges, v, g0["mass"][:P], g0["metallicity"][:P], g0["age"][:P], S, out=cube)
            got = ops.psf_lsf(cube, pk_h, lk_h).cpu().numpy()
            pdata = dict(coords=c.cpu().numpy(), velocity=v.cpu().numpy(), mass=g0["host"]["mass"][:P],
                         metallicity=g0["host"]["metallicity"][:P], age=g0["host"]["age"][:P])
            what = f"first {P} particles of galaxy 0 (rotated on the device), cube + PSF + LSF vs the float64 oracle"
        else:
            pdata = galaxy(args, P, seed=42)            # a P-particle galaxy drawn like the workload's
            mine = parallel.shard_particles(pdata, rank, world)
            out = device_step(ops.dev(mine["coords"]), ops.dev(mine["velocity"]), ops.dev(mine["mass"]),
                              ops.dev(mine["metallicity"]), ops.dev(mine["age"]))
            if slab_mode:   # gather the slabs (padded to wslab channels)
                pad = torch.zeros((S, S, wslab), dtype=torch.float32, device="cuda")
                pad[:, :, :out.shape[2]] = out
                parts = [torch.empty_like(pad) for _ in range(world)]
                dist.all_gather(parts, pad)
                got = torch.cat(parts, dim=2)[:, :, :W].cpu().numpy() if rank == 0 else None
            else:
                got = out.cpu().numpy() if rank == 0 else None
            what = (f"a {P}-particle bench-G galaxy sharded over {world} rank(s) like the timed step, reduced cube + "
                    "PSF + LSF vs the float64 oracle on the unsharded particles")
        if rank == 0:
            _, ref = cpu_arm(pdata, tpl, wave, edges_h, S, args.method, pk_h, lk_h, P, os.cpu_count() or 1, dtype=np.float64)
            err = float(np.abs(got.astype(np.float64) - ref).max())
            mx, tot = float(np.abs(ref).max()), float(np.abs(ref).sum())
            parity = {"max_abs_err_over_max": err / mx, "over_total": err / tot, "n": int(P),
                      "ok": bool(np.isfinite(got).all() and err <= 1e-5 * tot and err <= 2e-5 * mx),
                      "tolerance": "max|d| <= 1e-5 * sum|cube| (BASELINE north star) and <= 2e-5 * max|cube| (a sparse "
                                   "cube shows single knife-edge particles, tests/test_gpu_scale.py; dense cubes sit "
                                   "near 1e-6)",
                      "what": what}
        del pdata

    # ---- the HBM-bound stage of the path on its large configuration: PSF + LSF of a 150 x 150 x 3721 cube ---------
    stage = None
    if rank == 0 and world == 1 and G == 1 and not args.no_stage:
        big = torch.rand((150, 150, W), dtype=torch.float32, device="cuda")
        for _ in range(3):
            ops.psf_lsf(big, pk_h, lk_h)
        ts = []
        for _ in range(10):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.psf_lsf(big, pk_h, lk_h); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        stage = {"kernel": "psf_lsf_march_kernel", "cube": "150x150x3721 float32 (config 4), read once + written once",
                 "ms_median": float(np.median(ts)), "ms_min": float(np.min(ts)), "algorithmic_bytes": int(8 * big.numel())}
        del big

    if rank == 0:
        (hbm_peak, hbm_src), fp32 = measured_peaks()
        nz, na, L = tpl["flux"].shape
        alg_bytes = 40 * n + 4 * nz * na * L + 4 * S * S * W  # B_A per launch
        fused_s = mean_ms.value * 1e-3
        hbm_achieved = alg_bytes / fused_s / 1e9 if fused_s > 0 else 0.0
        sm_mhz = sampler.summary()["sm_mhz"] or 1965.0
        ncu = ncu_counts(args.method, n, S)
        roof = {"bound": "fp32", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None,
                "traffic": ncu.get("dram_bytes_per_launch") if ncu else None,
                "kernel": "fused_cube_warp_kernel", "kernel_ms": mean_ms.value, "kernel_launches_timed": int(nl.value),
                "kernel_share_of_step": mean_ms.value / ms_per_step if ms_per_step > 0 else None,
                "particles_per_launch": n,
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "algorithmic_bytes": alg_bytes, "peak_source": hbm_src},
                "hbm_frac": hbm_achieved / hbm_peak, "fp32_frac": None, "issue_frac": None}
        if ncu and fused_s > 0 and G == 1:
            if fp32 and ncu.get("flop_per_particle"):
                tf = ncu["flop_per_particle"] * n / fused_s / 1e12
                roof.update({"achieved": tf, "peak": fp32["fp32_tflops"], "frac": tf / fp32["fp32_tflops"],
                             "fp32_frac": tf / fp32["fp32_tflops"],
                             "peak_source": "measured (profiles/fp32_peak.json: tools/fma_peak.cu, best of FFMA / FFMA2)",
                             "flop_per_particle": ncu["flop_per_particle"]})
            if ncu.get("warp_inst_per_particle"):
                slots = (fp32["ffma_warp_inst_per_s"] if fp32 and fp32.get("ffma_warp_inst_per_s")
                         else 148 * 4 * sm_mhz * 1e6)
                roof["issue_frac"] = ncu["warp_inst_per_particle"] * n / fused_s / slots
                roof["issue"] = {"warp_inst_per_particle": ncu["warp_inst_per_particle"], "issue_slots_per_s": slots,
                                 "source": ("measured FFMA issue rate (profiles/fp32_peak.json)" if fp32
                                            else "148 SMs x 4 schedulers x SM clock")}
            roof["counts_source"] = (f"profiles/fused_ncu.json[{ncu['ncu_config']}] (ncu --set full, tools/ncu_summary.py); "
                                     + ("this configuration" if ncu["ncu_config_exact"] else
                                        "NEAREST captured particle count -- per-particle counts vary by a few % with the size"))
        roof["note"] = ("the cube kernel is bound by FP32 / instruction issue and the shared-memory pipe, not by HBM "
                        "(40 B per particle against ~30 kflop: SURVEY 8d, DESIGN.md section 5); frac = executed FP32 "
                        "flops (SASS opcode mix from ncu x live kernel time) / measured FMA peak")
        line = {
            "metric": METRIC, "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling_of(args),
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(args, tpl_name),
            "cube_build_ms": ms_per_step,
            "particles_per_gpu": n * G,
            "cube_kernel": {"error": err_impl[0],
                            "impl": "fused_cube_warp_kernel" if err_impl[1] == 0 else "fused_cube_kernel"},
            "roofline": roof,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if stage is not None:
            gbs = stage["algorithmic_bytes"] / (stage["ms_median"] * 1e-3) / 1e9
            stage.update({"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                          "peak_source": hbm_src})
            line["roofline_psf_lsf"] = stage
        if parity is not None:
            line["parity"] = parity
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu and world == 1 and G == 1:
            threads = os.cpu_count() or 1
            sample = min(n, args.cpu_sample)
            cpu_arm(data, tpl, wave, edges_h, S, args.method, pk_h, lk_h, min(sample, 20000), threads)
            ct, _ = cpu_arm(data, tpl, wave, edges_h, S, args.method, pk_h, lk_h, sample, threads)
            line["cpu_baseline"] = {"value": sample / ct, "unit": "particles/s", "cores": threads, "kind": "port",
                                    "sample": f"first {sample} of the workload's {n} particles, one pass, "
                                              f"oracle/rubix_oracle.c float32 on {threads} host threads",
                                    "one_core": one_core_rate(data, tpl, wave, edges_h, S, args.method, pk_h, lk_h,
                                                              sample)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        parallel.close_comm()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
