"""Per-layer timing of the Cout<=8 conv kernel (conv3d_tma_n8.cu) at the six cfg2 shapes, L2 flushed,
CUDA events; arguments: CASMVS_N8_DCHUNK caps to apply (0 = the library's own choice).  The library
reads the variable once per process, so every value runs in its own child process.

    python profiles/bench_n8.py 0 4 8
"""
import os, subprocess, sys
if len(sys.argv) > 2:
    for c in sys.argv[1:]:
        env = dict(os.environ)
        env.pop("CASMVS_N8_DCHUNK", None)
        if int(c):
            env["CASMVS_N8_DCHUNK"] = c
        subprocess.run([sys.executable, os.path.abspath(__file__), c], env=env, check=True)
    sys.exit(0)
import torch
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import ops
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [("s0.conv0", 8, 8, (8, 512, 640)), ("s0.prob", 8, 1, (8, 512, 640)),
          ("s1.conv0", 16, 8, (32, 256, 320)), ("s1.prob", 8, 1, (32, 256, 320)),
          ("s2.conv0", 32, 8, (48, 128, 160)), ("s2.prob", 8, 1, (48, 128, 160))]
chunks = [int(c) for c in sys.argv[1:]] or [0]
for name, cin, cout, dims in shapes:
    x = torch.randn(1, cin, *dims, device=dev)
    x = x.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    wt = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.1
    wp = ops.pack_conv3d_weight(wt, ops.CONV)
    sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
    row = []
    for c in chunks:
        ts = []
        for it in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = ops.conv3d(x, wp, cin, cout, sc, sh, 0.01, None, ops.CONV, 1, ops.TF32)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        row.append(f"{c or 'auto'}:{ts[len(ts)//2]:6.1f}")
    print(f"{name:9s} {cin:2d}->{cout}  " + "  ".join(row) + "  us", flush=True)
