"""K1 (fused warp + variance) micro-benchmark at the three BASELINE cfg2 stage
shapes: CUDA-event time per launch with the L2 flushed, algorithmic GB/s and the
fraction of the measured HBM peak.  Also checks the result against the variant
with the register window cache disabled (bit-identical is expected).

    CASMVS_K1_CACHE=1 CASMVS_K1_DCHUNK=0 python profiles/bench_k1.py [--views 3] [--gwc 0] [--profile]
"""
import argparse
import json
import os
import sys

import torch

torch.set_grad_enabled(False)   # inference scripts: the fused (non-autograd) path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import ops, synth   # noqa: E402
import bench                                 # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=3)
ap.add_argument("--gwc", type=int, default=0)
ap.add_argument("--W", type=int, default=640)
ap.add_argument("--H", type=int, default=512)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--ladder", action="store_true", help="hypotheses as first + step*d generated in the kernel (what the cascade runs)")
ap.add_argument("--profile", action="store_true", help="one launch per level inside cudaProfilerStart/Stop")
a = ap.parse_args()
dev = "cuda:0"
V, G = a.views, max(a.gwc, 1)
peak, src = bench.measured_peak_hbm()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pm = synth.projection_matrices(V, a.W, a.H).unsqueeze(0).to(dev)
tot_b = tot_ms = 0
out = []
for l, D in ((2, 48), (1, 32), (0, 8)):
    C, h, w = 8 * 2 ** l, a.H >> l, a.W >> l
    f = synth.make_level_feats(1, V, l, a.W, a.H, seed=1).to(dev)
    f = f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)      # channels-last storage
    # per-pixel hypotheses like a real stage: plane spacing interval*ratio around a smooth surface
    if l == 2:
        dv = ops.uniform_hypotheses(425.0, 2.65 * 4, D, 1, h, w, dev)
    else:
        ys, xs = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
        cur = (650 + 40 * torch.sin(3 * xs) * torch.cos(2 * ys)).reshape(1, 1, h, w).to(dev)
        dv = ops.depth_hypotheses(cur, D, 2.65 * (2 if l == 1 else 1))
    pml = pm[:, :, l].contiguous()
    if a.ladder:
        step = 2.65 * (4 if l == 2 else 2 if l == 1 else 1)
        lad = ops.Ladder(425.0 if l == 2 else dv[:, 0].contiguous(), step, D, 1, h, w, dev)
        assert torch.equal(lad.materialize(), dv)
        run = lambda: ops.warp_cost_ladder(f, pml, lad, G)            # noqa: E731
    else:
        run = lambda: ops.warp_cost(f, pml, dv, G, ops.NHWC)          # noqa: E731
    nbytes = 4 * (V * C * h * w + (C if G == 1 else G) * D * h * w + D * h * w) + 48 * (V - 1)
    if a.profile:
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        flush.zero_()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        run()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        continue
    ts = []
    for it in range(3 + a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = run()
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    tot_b += nbytes
    tot_ms += ms
    out.append(dict(level=l, C=C, D=D, ms=round(ms, 4), min_ms=round(min(ts), 4),
                    GBps=round(nbytes / ms / 1e6, 1), frac=round(nbytes / ms / 1e6 / peak, 3),
                    checksum=float(o.double().abs().sum())))
if not a.profile:
    print(json.dumps(dict(smem=os.environ.get("CASMVS_K1_SMEM", "1"),
                          variant=os.environ.get("CASMVS_K1S_VARIANT", "0"),
                          margin=os.environ.get("CASMVS_K1_MARGIN_X", "dflt"), ladder=a.ladder,
                          dchunk=os.environ.get("CASMVS_K1S_DCHUNK", "auto"), views=V, gwc=a.gwc,
                          total_ms=round(tot_ms, 4), GBps=round(tot_b / tot_ms / 1e6, 1),
                          frac_of_peak=round(tot_b / tot_ms / 1e6 / peak, 3), peak=peak, levels=out)))
