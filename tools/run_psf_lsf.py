"""Run the fused PSF+LSF call a few times (for ncu): python tools/run_psf_lsf.py [S] [W]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rubix_b200 import ops
from rubix_b200.telescope import gaussian_kernel_2d, lsf_kernel
S = int(sys.argv[1]) if len(sys.argv) > 1 else 150
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3721
cube = torch.rand((S, S, W), device="cuda")
pk, lk = ops.dev(gaussian_kernel_2d(5, 5, 0.6)), ops.dev(lsf_kernel(0.5, 1.25))
for _ in range(4):
    out = ops.psf_lsf(cube, pk, lk)
torch.cuda.synchronize()
print(float(out.sum()))
