"""tcgen05 stride-2 / transposed conv bring-up: TF32 tensor-core path vs fp32 CUDA cores."""
import os, sys
import torch

torch.set_grad_enabled(False)   # inference scripts: the fused (non-autograd) path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
cases = [("conv", 8, 16, (4, 16, 16)), ("conv", 8, 16, (8, 32, 40)), ("conv", 16, 32, (6, 20, 24)),
         ("conv", 8, 16, (48, 128, 160)), ("convT", 16, 8, (2, 16, 8)), ("convT", 16, 8, (5, 18, 11)),
         ("convT", 32, 16, (3, 16, 16)), ("convT", 16, 8, (24, 64, 80)), ("convT", 32, 16, (12, 32, 40)),
         ("conv", 32, 64, (4, 16, 16)), ("conv", 32, 64, (12, 32, 40)), ("convT", 64, 32, (2, 16, 8)),
         ("convT", 64, 32, (6, 16, 20)), ("convT", 64, 32, (1, 64, 80))]
for kind, cin, cout, dims in cases:
    x = torch.randn(1, cin, *dims, device=dev)
    scale = torch.rand(cout, device=dev) + 0.5
    shift = torch.randn(cout, device=dev) * 0.1
    if kind == "conv":
        wt = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.1
        k, skip = ops.CONV, None
    else:
        wt = torch.randn(cin, cout, 3, 3, 3, device=dev) * 0.1
        k = ops.CONV_TRANSPOSE
        skip = torch.randn(1, cout, *[2 * d for d in dims], device=dev)
    wp = ops.pack_conv3d_weight(wt, k)
    ref = ops.conv3d(x, wp, cin, cout, scale, shift, 0.01, skip, k, 2, ops.FP32)
    got = ops.conv3d(x, wp, cin, cout, scale, shift, 0.01, skip, k, 2, ops.TF32)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.conv3d(x, wp, cin, cout, scale, shift, 0.01, skip, k, 2, ops.TF32)
    e1.record(); torch.cuda.synchronize()
    print(f"{kind} cin={cin} cout={cout} dims={dims}: max|err|={err:.3e} max|ref|={ref.abs().max().item():.2f} "
          f"rel={err / ref.abs().max().item():.2e} nan={int(torch.isnan(got).sum())} {e0.elapsed_time(e1)/3*1e3:.1f} us", flush=True)
