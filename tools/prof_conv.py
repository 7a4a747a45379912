#!/usr/bin/env python
"""Launch the host-tap PSF+LSF kernel a few times on a 150x150x3721 cube (for ncu -k regex:march)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rubix_b200 import ops  # noqa: E402
from rubix_b200.telescope import gaussian_kernel_2d, lsf_kernel  # noqa: E402
S = int(sys.argv[1]) if len(sys.argv) > 1 else 150
cube = torch.rand((S, S, 3721), device="cuda")
pk, lk = gaussian_kernel_2d(5, 5, 0.6), lsf_kernel(0.5, 1.25)
for _ in range(4):
    out = ops.psf_lsf(cube, pk, lk)
torch.cuda.synchronize()
print(float(out[S // 2, S // 2, 100]))
