"""Functional + timing run of the other BASELINE.json configs (one reference view each):
cfg3 640x512 V=3 gwc G=8; cfg4 1152x864 V=5; cfg5 1920x1056 V=7 D=64/32/8.
Checks finiteness, and at level 2 compares K1 against the CPU oracle (seconds at these sizes)."""
import os, sys, time, json
import torch

torch.set_grad_enabled(False)   # inference scripts: the fused (non-autograd) path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import ABN, ops, synth
from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet
from oracle import casmvs_oracle as O

dev = "cuda:0"
cfgs = [("cfg3", 640, 512, 3, 8, (8, 32, 48)), ("cfg4", 1152, 864, 5, 1, (8, 32, 48)),
        ("cfg5", 1920, 1056, 7, 1, (8, 32, 64))]
for name, W, H, V, G, nd in cfgs:
    torch.manual_seed(0)
    m = CascadeMVSNet(n_depths=list(nd), num_groups=G, norm_act=ABN, precision="tf32")
    synth.randomize_model_(m, 0)
    m = m.eval().to(dev)
    imgs, pm, dmin, dint = synth.make_inputs(B=1, V=V, W=W, H=H, seed=0)
    imgs, pm = imgs.to(dev), pm.to(dev)
    for _ in range(2):
        res = m(imgs, pm, dmin, dint)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        res = m(imgs, pm, dmin, dint)
    e1.record(); torch.cuda.synchronize()
    ok = all(torch.isfinite(v).all().item() for v in res.values())
    # tcgen05 path vs the fp32 CUDA-core / cuDNN-fp32 path of the same model at this size
    d_tf32 = res["depth_0"].clone()
    m.set_precision("fp32")
    ref32 = m(imgs, pm, dmin, dint)["depth_0"]
    m.set_precision("tf32")
    rel_l1 = ((d_tf32 - ref32).abs().mean() / ref32.abs().mean()).item()
    del ref32
    # K1 parity at level 2 against the oracle
    with torch.no_grad():
        f = m.feature(imgs.reshape(V, 3, H, W))["level_2"]
    f = f.view(1, V, *f.shape[1:])
    h, w = f.shape[-2:]
    dv = ops.uniform_hypotheses(dmin, dint * 4, nd[2], 1, h, w, dev)
    got = ops.warp_cost(f, pm[:, :, 2], dv, G, ops.NCHW).cpu()
    fc, pc, dc = f.cpu().contiguous(), pm[:, :, 2].cpu(), dv.cpu()
    want = O.variance_cost_volume(fc, pc, dc) if G == 1 else O.groupwise_cost_volume(fc, pc, dc, G)
    err = (got - want).abs().max().item()
    print(json.dumps(dict(cfg=name, W=W, H=H, V=V, G=G, n_depths=nd, finite=ok,
                          ms_per_depth_map=round(e0.elapsed_time(e1) / 5, 3),
                          depth0_rel_l1_tf32_vs_fp32=rel_l1, k1_level2_max_err=err, k1_level2_max_ref=want.abs().max().item(),
                          depth0_range=[res["depth_0"].min().item(), res["depth_0"].max().item()],
                          peak_mem_GB=round(torch.cuda.max_memory_allocated() / 2**30, 2))), flush=True)
    del m, res
    torch.cuda.empty_cache()
