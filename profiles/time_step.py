"""Step time of the cfg2 cascade under CUDA-graph replay (ms, median of 5 x 100 replays): the quick
A/B tool for the experiment switches of DESIGN.md section 8.

    CASMVS_DCHUNK_FIXED=3 python profiles/time_step.py [--hot]
"""
import os
import sys

import torch

torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import ABN, synth                      # noqa: E402
from casmvsnet_pl_b200.graph import GraphedCascade             # noqa: E402
from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet      # noqa: E402

torch.manual_seed(0)
m = CascadeMVSNet(norm_act=ABN, precision=os.environ.get("CASMVS_PRECISION", "tf32"))
synth.randomize_model_(m, 0)
m = m.eval().cuda()
imgs, pm, dmin, dint = synth.make_inputs(B=1, V=3, W=640, H=512, seed=0)
g = GraphedCascade(m, imgs.cuda(), pm.cuda(), dmin, dint)
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        g()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 100)
ts.sort()
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("CASMVS_"))
print(f"{ts[2]:.4f} ms/step (min {ts[0]:.4f})  {tag or 'defaults'}")
