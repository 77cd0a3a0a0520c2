"""One steady-state cfg2 step inside cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` launch lists and captures."""
import os
import sys

import torch

torch.set_grad_enabled(False)   # inference scripts: the fused (non-autograd) path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import ABN, synth                      # noqa: E402
from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet      # noqa: E402

prec = os.environ.get("CASMVS_PRECISION", "tf32")
torch.manual_seed(0)
m = CascadeMVSNet(norm_act=ABN, precision=prec)
synth.randomize_model_(m, 0)
m = m.eval().cuda()
imgs, pm, dmin, dint = synth.make_inputs(B=1, V=3, W=640, H=512, seed=0)
imgs, pm = imgs.cuda(), pm.cuda()
for _ in range(4):
    m(imgs, pm, dmin, dint)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m(imgs, pm, dmin, dint)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
