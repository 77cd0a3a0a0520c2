"""Per-role clock64 timeline of CTA 0 of the tcgen05 conv kernel (debug aid)."""
import os, sys
import torch

torch.set_grad_enabled(False)   # inference scripts: the fused (non-autograd) path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dev = "cuda:0"
dbg = torch.zeros(4 * 64 * 4, dtype=torch.int64, device=dev)
os.environ["CASMVS_TC_DBG"] = hex(dbg.data_ptr())
from casmvsnet_pl_b200 import ops
cin, cout, dims = int(sys.argv[1]), int(sys.argv[2]), tuple(int(v) for v in sys.argv[3:6])
x = torch.randn(1, cin, *dims, device=dev)
wt = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.1
wp = ops.pack_conv3d_weight(wt, ops.CONV)
for _ in range(3):
    y = ops.conv3d(x, wp, cin, cout, None, None, 0.01, None, ops.CONV, 1, ops.TF32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y = ops.conv3d(x, wp, cin, cout, None, None, 0.01, None, ops.CONV, 1, ops.TF32)
e1.record()
torch.cuda.synchronize()
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us for cin={cin} cout={cout} dims={dims}")
t = dbg.cpu().reshape(4, 64, 4)
t0 = int(t[3, 0, 0])
print("setup cycles", int(t[3, 0, 1]) - t0)
for it in range(44):
    pr = [int(v) - t0 if v else -1 for v in t[0, it]]
    mm = [int(v) - t0 if v else -1 for v in t[1, it]]
    ep = [int(v) - t0 if v else -1 for v in t[2, it]]
    print(f"it {it:2d} prod start/gotslot/issued/prevlanded {pr}  mma wait/full/acc/issued {mm}  epi wait/full/freed {ep}")
