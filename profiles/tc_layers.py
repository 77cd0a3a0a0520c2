"""Run every tensor-core-eligible CostRegNet layer shape of cfg2 once (hang finder / timer)."""
import os, sys, time
import torch

torch.set_grad_enabled(False)   # inference scripts: the fused (non-autograd) path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import ops
dev = "cuda:0"
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
shapes = []
for l, (D, H, W) in {2: (48, 128, 160), 1: (32, 256, 320), 0: (8, 512, 640)}.items():
    C = 8 * 2 ** l
    shapes += [(f"l{l}.conv0", C, 8, (D, H, W)), (f"l{l}.conv2", 16, 16, (D // 2, H // 2, W // 2)),
               (f"l{l}.conv4", 32, 32, (D // 4, H // 4, W // 4)),
               (f"l{l}.conv6", 64, 64, (D // 8, H // 8, W // 8)), (f"l{l}.prob", 8, 1, (D, H, W))]
for i, (name, cin, cout, dims) in enumerate(shapes):
    if only >= 0 and i != only:
        continue
    x = torch.randn(1, cin, *dims, device=dev)
    wt = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.1
    wp = ops.pack_conv3d_weight(wt, ops.CONV)
    sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
    for _ in range(2):
        y = ops.conv3d(x, wp, cin, cout, sc, sh, 0.01, None, ops.CONV, 1, ops.TF32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = ops.conv3d(x, wp, cin, cout, sc, sh, 0.01, None, ops.CONV, 1, ops.TF32)
    e1.record(); torch.cuda.synchronize()
    nbytes = 4 * (cin + cout) * dims[0] * dims[1] * dims[2]
    us = e0.elapsed_time(e1) * 1e3 / 5
    print(f"{i:2d} {name:10s} cin={cin:2d} cout={cout:2d} dims={dims}: {us:8.1f} us  {nbytes/us/1e3:7.1f} GB/s", flush=True)
