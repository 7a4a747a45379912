"""Times rbx_sort_by_spaxel (stand-alone stable radix sort of particles by spaxel id) on the current GPU with CUDA
events, L2 flushed between runs; writes one JSON object.  Algorithmic bytes as SURVEY 8d: 12 B per particle (the id
read twice, the permutation written once).  usage: python tools/sort_probe.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rubix_b200 import _lib, ops, synthetic  # noqa: E402

peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs")
except Exception:
    pass
peak = float(peak or 6545.9)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"peak_gbs": peak, "cases": []}
for n, S in ((10**6, 25), (10**7, 25), (10**7, 150)):
    d = synthetic.bench_g(n)
    pix = ops.filter_and_assign(d["coords"], synthetic.spatial_edges(S))
    nseg = S * S
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    srt = torch.empty(n, dtype=torch.int32, device="cuda")
    off = torch.empty(nseg + 1, dtype=torch.int32, device="cuda")
    nb = int(_lib.lib().rbx_sort_by_spaxel_workspace_bytes(n, nseg))
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: t.data_ptr()
    ms = []
    for it in range(8):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(_lib.lib().rbx_sort_by_spaxel(P(pix), n, nseg, P(order), P(srt), P(off), P(ws), nb, st))
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms = sorted(ms[3:])
    med = ms[len(ms) // 2]
    key = np.where(pix.cpu().numpy() < 0, nseg, pix.cpu().numpy())
    ok = bool(np.array_equal(order.cpu().numpy(), np.argsort(key, kind="stable").astype(np.int32)))
    out["cases"].append({"n": n, "num_segments": nseg, "ms_median": med, "ms_min": ms[0], "bit_exact_vs_numpy": ok,
                         "algorithmic_bytes": 12 * n, "achieved_gbs": 12 * n / med / 1e6,
                         "hbm_frac": 12 * n / med / 1e6 / peak,
                         "launches": "memset + spaxel_keys_kernel + 2 radix_pass_kernel + segment_offsets_kernel"})
    print(out["cases"][-1], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
