#!/usr/bin/env python
"""bench.py -- particles/s binned into a MUSE datacube (with PSF + LSF) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--particles P] [--method linear|cubic] [--spaxels S]

A "step" is one pass of the hot path over one batch of synthetic star particles (bench-G of
SURVEY.md section 8(d)): filter_particles -> spaxel_assignment -> fused SSP lookup / mass scaling /
Doppler shift / flux-conserving resample / cube accumulation -> PSF -> LSF.

* ``value``  whole-job particles/s with the particle arrays already resident in HBM (CUDA events
             around each step on the launch stream, L2 flushed between steps, max over ranks).
* ``e2e``    the same metric through the host-buffer C-ABI call (``rbx_pipeline_host``): pinned host
             arrays in, host cube out, H2D + D2H inside the timed region.
* ``roofline``  the dominant kernel (fused_cube_warp_kernel), timed by CUDA events inside the library on
             its launch stream; achieved = algorithmic HBM bytes / duration against the measured
             copy bandwidth in MEASURED_PEAKS.json.  The kernel is instruction-issue bound, not HBM
             bound, so ``frac`` is small by construction; ``issue`` adds the
             achieved particles/s against the issue-slot ceiling derived from the SASS.
* ``cpu_baseline``  the C restatement of the reference (oracle/rubix_oracle.c, float32, all host
             threads) on a bounded sample of the same workload.

``--impl reference`` times that CPU restatement as the reference arm (the reference itself is
pure Python/JAX and cannot be installed here: no jax / interpax wheels, no network).
N > 1: one process per GPU (torchrun), every rank bins its own ``--particles`` particles of the
galaxy (weak scaling), the partial cubes are summed with one NCCL reduce, rank 0 applies PSF + LSF.
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


def load_template():
    path = os.path.join(ROOT, "tests", "golden", "bc03lr_f32.npz")
    if os.path.exists(path):
        d = np.load(path)
        return {k: d[k] for k in d.files}, "BC03lr (float32 fixture of the reference's template)"
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


def cpu_arm(data, tpl, wave, edges, S, method, pk, lk, sample, threads):
    """One pass of the reference algorithm (C restatement, float32) over ``sample`` particles."""
    from oracle import c_oracle
    sl = slice(0, sample)
    t0 = time.perf_counter()
    cube = c_oracle.particles_to_cube(data["coords"][sl], data["velocity"][sl], data["mass"][sl],
                                      data["metallicity"][sl], data["age"][sl], edges, S, tpl["metallicity"],
                                      tpl["age"], tpl["wavelength"], tpl["flux"], wave, REDSHIFT, method=method,
                                      dtype=np.float32, n_threads=threads)
    cube = c_oracle.apply_psf(cube, pk)
    cube = c_oracle.apply_lsf(cube, lk)
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
    edges = synthetic.spatial_edges(S)
    pk, lk = host_kernels()
    threads = os.cpu_count() or 1
    sample = min(args.particles, args.cpu_sample)
    data = synthetic.bench_g(sample, seed=42)
    for _ in range(args.warmup):
        cpu_arm(data, tpl, wave, edges, S, args.method, pk, lk, min(sample, 20000), threads)
    times = [cpu_arm(data, tpl, wave, edges, S, args.method, pk, lk, sample, threads)[0] for _ in range(args.steps)]
    t = float(np.mean(times))
    val = sample / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "particles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "template": tpl_name, "method": args.method,
                   "note": "C restatement of rubix@dbb4487 (oracle/rubix_oracle.c), not jax: the reference is "
                           "pure Python/JAX and cannot be installed in this image"},
        "cpu_baseline": {"value": val, "unit": "particles/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} bench-G particles per step, all {threads} host threads (OpenMP)",
                         "one_core": one_core_rate(data, tpl, wave, edges, S, args.method, pk, lk, sample)},
        "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(args):
    if getattr(args, "galaxies", 1) > 1:
        return (f"survey batch: {args.galaxies} synthetic galaxies per GPU x {args.particles} star particles, each "
                f"rotated to its own inclination, MUSE {args.spaxels}x{args.spaxels} spaxels x 3721 channels, BC03 SSP, "
                f"ssp.method={args.method}, gaussian PSF 5/0.6 + LSF sigma 0.5")
    return (f"{args.particles} synthetic star particles per GPU (bench-G), MUSE {args.spaxels}x{args.spaxels} "
            f"spaxels x 3721 channels, BC03 SSP, ssp.method={args.method}, gaussian PSF 5/0.6 + LSF sigma 0.5")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--method", default="linear", choices=["linear", "cubic"])
    ap.add_argument("--spaxels", type=int, default=25)
    ap.add_argument("--cpu-sample", type=int, default=400_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--galaxies", type=int, default=1,
                    help="survey batch (config 5): galaxies per GPU and step, each with its own inclination "
                         "(rotate_galaxy on the device); replicas only, no collective")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from rubix_b200 import _lib, ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    tpl, tpl_name = load_template()
    wave = synthetic.muse_wave()
    S = args.spaxels
    edges_h = synthetic.spatial_edges(S)
    pk_h, lk_h = host_kernels()
    n = args.particles
    data = synthetic.bench_g(n, seed=42 + rank)
    plan = ops.Plan(tpl["metallicity"], tpl["age"], tpl["wavelength"], tpl["flux"], wave, REDSHIFT,
                    method=args.method, direction="z")

    # device-resident inputs for `value`
    coords, vel = ops.dev(data["coords"]), ops.dev(data["velocity"])
    mass0, met0, age0 = ops.dev(data["mass"]), ops.dev(data["metallicity"]), ops.dev(data["age"])
    mass, met, age = mass0.clone(), met0.clone(), age0.clone()
    edges, pk, lk = ops.dev(edges_h), ops.dev(pk_h), ops.dev(lk_h)
    cube = torch.empty((S, S, plan.W), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    # N > 1, two exchange patterns: a MUSE-size cube (9.3 MB) is summed onto rank 0, which applies
    # PSF + LSF (14 us); a large-FOV cube (S = 150: 335 MB) is all-reduced and every rank convolves its own
    # wavelength slab in place (+-12 channel halo), the result staying sharded by wavelength.
    from rubix_b200 import parallel
    slab_mode = world > 1 and cube.numel() * 4 > (64 << 20)
    slab_lo, slab_hi = parallel.wavelength_slab(plan.W, rank, world)

    G = max(1, args.galaxies)
    if G > 1:   # survey batch: a flattened disc per galaxy, its own Euler angles
        rng = np.random.default_rng(1000 + rank)
        gal = []
        for g in range(G):
            dg = synthetic.bench_g(n, seed=4200 + 64 * rank + g)
            dg["coords"][:, 2] *= 0.2
            gal.append(dict(coords=ops.dev(dg["coords"]), velocity=ops.dev(dg["velocity"]), mass=ops.dev(dg["mass"]),
                            metallicity=ops.dev(dg["metallicity"]), age=ops.dev(dg["age"]),
                            angles=tuple(float(a) for a in rng.uniform(0.0, 180.0, 3))))

    def survey_step():
        out = None
        for g in gal:   # independent galaxies: no collective (SURVEY 8e "replicas only")
            c, v, _ = ops.rotate_galaxy(g["coords"], g["velocity"], g["mass"], 1.5, *g["angles"])
            ops.assign_build_cube(plan, c, edges, v, g["mass"], g["metallicity"], g["age"], S, out=cube)
            out = ops.psf_lsf(cube, pk_h, lk_h)
        return out

    def step():
        if G > 1:
            return survey_step()
        # filter_particles + spaxel_assignment inside the first kernel of the fused cube build
        ops.assign_build_cube(plan, coords, edges, vel, mass, met, age, S, out=cube)
        if slab_mode:
            dist.all_reduce(cube, op=dist.ReduceOp.SUM)
            return ops.psf_lsf_slab(cube, slab_lo, slab_hi, pk_h, lk_h)
        if world > 1:
            dist.reduce(cube, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            return ops.psf_lsf(cube, pk_h, lk_h)  # host taps, as the reference builds them
        return cube

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    _lib.lib().rbx_profile_enable(1)
    import ctypes as C
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
    value = world * G * n / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI (rank-local galaxy; N>1 reduces on the device) -------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hp = {k: pin(v) for k, v in data.items()}
    hcube = torch.empty((S, S, plan.W), dtype=torch.float32).pin_memory()
    hnp = {k: v.numpy() for k, v in hp.items()}

    if world == 1:
        def e2e_step():
            out = ops.pipeline_host(plan, hnp["coords"], hnp["velocity"], hnp["mass"], hnp["metallicity"],
                                    hnp["age"], edges_h, S, pk_h, lk_h, out=hcube.numpy())
            return float(out[S // 2, S // 2, 100])  # the host reads the result
        e2e_api = "rbx_pipeline_host (C ABI, pinned host buffers)"
    else:
        # N > 1: pinned host shard -> device (async copies), device path, NCCL reduce, PSF+LSF on rank 0,
        # cube back to rank 0's host
        dcoords, dvel = torch.empty_like(coords), torch.empty_like(vel)

        def e2e_step():
            dcoords.copy_(hp["coords"], non_blocking=True); dvel.copy_(hp["velocity"], non_blocking=True)
            mass.copy_(hp["mass"], non_blocking=True); met.copy_(hp["metallicity"], non_blocking=True)
            age.copy_(hp["age"], non_blocking=True)
            ops.assign_build_cube(plan, dcoords, edges, dvel, mass, met, age, S, out=cube)
            if slab_mode:
                dist.all_reduce(cube, op=dist.ReduceOp.SUM)
                hslab.copy_(ops.psf_lsf_slab(cube, slab_lo, slab_hi, pk_h, lk_h), non_blocking=True)
                torch.cuda.synchronize()
                return float(hslab[S // 2, S // 2, 0])
            dist.reduce(cube, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                hcube.copy_(ops.psf_lsf(cube, pk_h, lk_h), non_blocking=True)
            torch.cuda.synchronize()
            return float(hcube[S // 2, S // 2, 100]) if rank == 0 else 0.0
        hslab = torch.empty((S, S, slab_hi - slab_lo), dtype=torch.float32).pin_memory() if slab_mode else None
        e2e_api = ("device ops through the C ABI with pinned host shards, NCCL all-reduce, PSF+LSF per wavelength slab, "
                   "each slab to its rank's host" if slab_mode else
                   "device ops through the C ABI with pinned host shards, NCCL reduce, cube to rank 0's host")

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_steps = max(3, min(args.steps, 10))
    for _ in range(e_steps):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        e2e_step()
    torch.cuda.synchronize()
    # the L2 flush (a 256 MB fill, ~0.1 ms) is inside this wall-clock bracket; negligible vs PCIe copies
    e2e_s = (time.perf_counter() - t0) / e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * n / float(te.item())
    h2d = sum(v.nbytes for v in hnp.values()) + edges_h.nbytes + pk_h.nbytes + lk_h.nbytes
    d2h = (hslab.numel() if (world > 1 and slab_mode) else hcube.numel()) * 4

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            pk = json.load(open(peaks_path))
            for key in ("hbm_gbs", "hbm_gbs_burst", "hbm_gb_s", "hbm"):
                if isinstance(pk.get(key), (int, float)) and pk[key] > 0:
                    peak, peak_src = float(pk[key]), f"measured (MEASURED_PEAKS.json {key})"
                    break
        except (OSError, ValueError, AttributeError):
            pass
        nz, na, L = tpl["flux"].shape
        alg_bytes = 40 * n + 4 * nz * na * L + 4 * S * S * plan.W  # B_A per launch
        fused_s = mean_ms.value * 1e-3
        achieved = alg_bytes / fused_s / 1e9 if fused_s > 0 else 0.0
        # figures of the last committed ncu --set full capture of this kernel (tools/ncu_summary.py)
        ncu = {}
        ncu_path = os.path.join(ROOT, "profiles", "fused_ncu.json")
        if os.path.exists(ncu_path):
            ncu = json.load(open(ncu_path)).get(f"{args.method}_{n}", {})
        sm_mhz = sampler.summary()["sm_mhz"] or 1965.0
        issue = None
        if ncu.get("warp_inst_per_particle") and fused_s > 0:
            slots = 148 * 4 * sm_mhz * 1e6  # warp instructions the GPU can issue per second
            issue = {"warp_inst_per_particle": ncu["warp_inst_per_particle"], "issue_slots_per_s": slots,
                     "achieved_frac": ncu["warp_inst_per_particle"] * n / fused_s / slots,
                     "source": "profiles/fused_ncu.json (ncu --set full) x live kernel time"}
        line = {
            "metric": METRIC, "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "template": tpl_name, "method": args.method,
                       "particles_per_gpu": n * G, "galaxies_per_gpu": G,
                       "l2": "flushed (256 MB fill) between timed steps",
                       "parallelism": (f"particle-sharded x{world}, one NCCL all-reduce of the partial cubes, PSF+LSF "
                                       "sharded by wavelength slab" if slab_mode else
                                       f"particle-sharded x{world}, one NCCL reduce of the partial cubes")},
            "cube_build_ms": ms_per_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu.get("dram_bytes_per_launch"),
                         "kernel": "fused_cube_warp_kernel",
                         "kernel_ms": mean_ms.value, "kernel_launches_timed": int(nl.value),
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_src, "issue": issue,
                         "note": "this kernel is bound by instruction issue and the shared-memory (LSU wavefront) pipe, "
                                 "not by HBM (DESIGN.md section 5): 40 B and "
                                 f"~{ncu.get('warp_inst_per_particle', 460):.0f} warp instructions per particle; kernel "
                                 f"share of step = {mean_ms.value / ms_per_step:.2f}"},
            "e2e": {"value": e2e_val, "unit": "particles/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": float(te.item()) * 1e3,
                    "api": e2e_api},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = min(n, args.cpu_sample)
            cpu_arm(data, tpl, wave, edges_h, S, args.method, pk_h, lk_h, min(sample, 20000), threads)
            ct, _ = cpu_arm(data, tpl, wave, edges_h, S, args.method, pk_h, lk_h, sample, threads)
            line["cpu_baseline"] = {"value": sample / ct, "unit": "particles/s", "cores": threads, "kind": "port",
                                    "sample": f"first {sample} particles of the same workload, one pass, "
                                              f"oracle/rubix_oracle.c float32 on {threads} host threads",
                                    "one_core": one_core_rate(data, tpl, wave, edges_h, S, args.method, pk_h, lk_h,
                                                              sample)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
