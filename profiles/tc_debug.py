"""tcgen05 conv bring-up: TC (tf32) vs CUDA-core fp32 on random volumes."""
import sys, os
import torch

torch.set_grad_enabled(False)   # inference scripts: the fused (non-autograd) path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
cases = [(8, 8, (2, 16, 8)), (8, 8, (4, 16, 8)), (8, 8, (5, 20, 13)), (16, 16, (6, 32, 24)),
         (32, 8, (8, 32, 40)), (32, 32, (4, 16, 16)), (8, 1, (8, 32, 16)), (16, 8, (48, 128, 160)),
         (8, 8, (1, 7, 33)), (16, 8, (23, 9, 61)), (32, 1, (3, 5, 29)), (8, 8, (11, 30, 31)),
         (64, 64, (2, 16, 8)), (64, 64, (6, 16, 20)), (64, 64, (4, 32, 40)), (64, 64, (1, 64, 80))]
for cin, cout, dims in cases:
    nb = 2 if dims[0] == 11 else 1
    x = torch.randn(nb, cin, *dims, device=dev)
    wt = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.1
    scale = torch.rand(cout, device=dev) + 0.5
    shift = torch.randn(cout, device=dev) * 0.1
    wp = ops.pack_conv3d_weight(wt, ops.CONV)
    skip = torch.randn(nb, cout, *dims, device=dev)
    ref = ops.conv3d(x, wp, cin, cout, scale, shift, 0.01, skip, ops.CONV, 1, ops.FP32)
    got = ops.conv3d(x, wp, cin, cout, scale, shift, 0.01, skip, ops.CONV, 1, ops.TF32)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    print(f"cin={cin} cout={cout} dims={dims}: max|err|={err:.3e} max|ref|={ref.abs().max().item():.2f} "
          f"rel={err / ref.abs().max().item():.2e} nan={int(torch.isnan(got).sum())}", flush=True)
