#!/usr/bin/env python
"""Benchmark of the cascade-MVS depth hot path on B200 (contract: task prompt §④).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one `CascadeMVSNet.forward` (FeatureNet + three cascade stages) over
one batch of synthetic DTU-shaped views per GPU: BASELINE.json configs[1] —
640x512, V=3, D=48/32/8, variance cost, B=1 per GPU.  Metric: depth-maps/sec.

  value     : whole-job depth-maps/s, inputs resident in HBM when timing starts
  e2e       : same through the public API with HOST (pinned) inputs; H2D copy of
              imgs+proj and D2H read of depth_0 + confidence_2 inside the timed region
  roofline  : the fused warp+variance kernel (K1): algorithmic bytes of its three
              launches / their CUDA-event time, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the CPU oracle port of the reference path on this box's host cores
  --impl reference : the reference arm = the same CPU port, all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_IMG, H_IMG, VIEWS = 640, 512, 3
N_DEPTHS = (8, 32, 48)          # level 0..2  (BASELINE writes coarse->fine 48/32/8)
RATIOS = (1, 2, 4)
METRIC = "depth-maps/sec at 640x512 V=3 D=48/32/8"
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback


def k1_algorithmic_bytes(V, G=1, W=W_IMG, H=H_IMG, n_depths=N_DEPTHS):
    """4*[V*C*h*w + Cin3d*D*h*w + D*h*w] + 48*(V-1) per stage."""
    per = []
    for l in (2, 1, 0):
        C, D, h, w = 8 * 2 ** l, n_depths[l], H >> l, W >> l
        cin3d = C if G == 1 else G
        per.append(4 * (V * C * h * w + cin3d * D * h * w + D * h * w) + 48 * (V - 1))
    return per


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_forward_factory(threads):
    """The CPU reference arm / cpu_baseline: oracle port of the reference path."""
    import torch
    from casmvsnet_pl_b200 import ABN, synth
    from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet
    from oracle import casmvs_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = CascadeMVSNet(n_depths=list(N_DEPTHS), interval_ratios=list(RATIOS), norm_act=ABN)
    synth.randomize_model_(m, 0)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    imgs, pm, dmin, dint = synth.make_inputs(B=1, V=VIEWS, W=W_IMG, H=H_IMG, seed=0)

    def fwd():
        return O.cascade_forward(sd, imgs, pm, dmin, dint, N_DEPTHS, RATIOS, 1)
    return fwd


def best_cpu_threads(cores):
    """torch/oneDNN over-subscribe badly on many-core hosts (128 threads measured 7x
    slower than 8 on the GPU box), so the CPU arm uses the thread count that is
    FASTEST for it: probed on the coarsest stage (128x160, D=48) of the workload."""
    import torch
    from casmvsnet_pl_b200 import ABN, synth
    from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet
    from oracle import casmvs_oracle as O
    torch.manual_seed(0)
    m = CascadeMVSNet(n_depths=list(N_DEPTHS), interval_ratios=list(RATIOS), norm_act=ABN)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    feats = synth.make_level_feats(1, VIEWS, 2, W_IMG, H_IMG)
    pm = synth.projection_matrices(VIEWS, W_IMG, H_IMG)[:, 2].unsqueeze(0)
    dv = O.initial_hypotheses(425.0, 2.65 * 4, 48, 1, H_IMG // 4, W_IMG // 4).contiguous()
    cands = sorted({t for t in (4, 8, 16, 32, 64, cores) if t <= cores})
    best, best_t = None, None
    for t in cands:
        torch.set_num_threads(t)
        with torch.no_grad():
            O.predict_depth(feats, pm, dv, sd, "cost_reg_2.", 1)
            t0 = time.perf_counter()
            O.predict_depth(feats, pm, dv, sd, "cost_reg_2.", 1)
            dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, best_t = dt, t
    return best_t


def time_cpu_port(steps, warmup, threads):
    fwd = cpu_port_forward_factory(threads)
    for _ in range(warmup):
        fwd()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fwd()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], sum(ts)


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import torch
    avail = len(os.sched_getaffinity(0))
    cores = best_cpu_threads(avail)
    steps = max(1, args.steps)
    med, total = time_cpu_port(steps, max(1, min(args.warmup, 2)), cores)
    value = steps / total
    sample = (f"{steps} full forwards of the cfg2 workload (640x512, V=3, D=48/32/8, B=1) on "
              f"{cores} torch threads (fastest of 4..{avail} available; more threads are slower); "
              f"oracle port of the reference PyTorch-CPU path "
              f"(/root/reference does not travel to the GPU box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "depth-maps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 640x512, V=3, D=48/32/8, variance cost, B=1",
                   "device": "cpu", "torch_threads": torch.get_num_threads()},
        "cpu_baseline": {"value": value, "unit": "depth-maps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "depth-maps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CASMVS_PRECISION", "tf32"),
                    choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from casmvsnet_pl_b200 import ABN, _lib, ops, synth
    from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a B200; no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    model = CascadeMVSNet(n_depths=list(N_DEPTHS), interval_ratios=list(RATIOS), norm_act=ABN,
                          precision=args.precision)
    synth.randomize_model_(model, 0)
    model = model.eval().to(dev)
    B = 1
    imgs_h, pm_h, dmin, dint = synth.make_inputs(B=B, V=VIEWS, W=W_IMG, H=H_IMG, seed=rank)
    imgs_h, pm_h = imgs_h.pin_memory(), pm_h.pin_memory()
    imgs_d, pm_d = imgs_h.to(dev), pm_h.to(dev)
    gather_buf = [torch.empty(B, H_IMG, W_IMG, device=dev) for _ in range(world)] if world > 1 else None

    graphed = None
    if not args.no_graph:
        from casmvsnet_pl_b200.graph import GraphedCascade
        graphed = GraphedCascade(model, imgs_d, pm_d, dmin, dint)

    def step_resident():
        res = graphed() if graphed is not None else model(imgs_d, pm_d, dmin, dint)
        if world > 1:
            dist.all_gather(gather_buf, res["depth_0"])     # the path's only collective (§8e)
        return res

    out_depth_h = torch.empty(B, H_IMG, W_IMG).pin_memory()
    out_conf_h = torch.empty(B, H_IMG // 4, W_IMG // 4).pin_memory()

    def step_e2e():
        if graphed is not None:
            res = graphed(imgs_h, pm_h)                          # H2D into the static buffers
        else:
            res = model(imgs_h.to(dev, non_blocking=True), pm_h.to(dev, non_blocking=True),
                        dmin, dint)
        if world > 1:
            dist.all_gather(gather_buf, res["depth_0"])
        out_depth_h.copy_(res["depth_0"], non_blocking=True)     # what eval.py:224-226 reads back
        out_conf_h.copy_(res["confidence_2"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return res

    pipe = None
    if graphed is not None and world == 1:
        from casmvsnet_pl_b200.graph import PipelinedCascade
        pipe = PipelinedCascade(model, imgs_d, pm_d, dmin, dint)

    def run_e2e_pipelined(steps):
        # every step: H2D of that step's inputs from pinned memory, forward, D2H of its results;
        # copies of neighbouring steps overlap the compute (two slots)
        for _ in range(steps):
            pipe.submit(imgs_h, pm_h)
        pipe.drain()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(args.warmup):
        step_resident()
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = _lib.launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = _lib.launch_count() - n0
    if graphed is not None:      # graph replays launch the captured libcasmvs kernels
        launches += graphed.kernels_per_replay * args.steps
    if pipe is not None:
        run_e2e_pipelined(3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e_pipelined(args.steps)        # drain() waits for the last D2H
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1)
    else:
        ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None

    # ---- K1 roofline: the three launches of one depth map, CUDA events, L2 flushed ----
    roofline = None
    hot = None
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
        per_stage_bytes = k1_algorithmic_bytes(VIEWS)
        stage_ms = []
        with torch.no_grad():
            feats = model.feature(imgs_d.reshape(B * VIEWS, 3, H_IMG, W_IMG))
            for i, l in enumerate((2, 1, 0)):
                f = feats[f"level_{l}"]
                f = f.view(B, VIEWS, *f.shape[1:])
                D = N_DEPTHS[l]
                h, w = f.shape[-2:]
                dv = ops.uniform_hypotheses(dmin + 3.0 * l, dint * RATIOS[l], D, B, h, w, dev)
                pml = pm_d[:, :, l].contiguous()
                ts = []
                for it in range(3 + 10):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.warp_cost(f, pml, dv, 1, ops.NHWC)
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= 3:
                        ts.append(e0.elapsed_time(e1))
                stage_ms.append(sum(ts) / len(ts))
        tot_bytes = sum(per_stage_bytes)
        tot_ms = sum(stage_ms)
        achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
        roofline = {"kernel": "warp_cost_kernel (K1, fused warp+variance), 3 launches / depth map",
                    "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                    "per_stage": [{"level": l, "algorithmic_bytes": b, "ms": m,
                                   "GBps": b / (m * 1e-3) / 1e9}
                                  for l, b, m in zip((2, 1, 0), per_stage_bytes, stage_ms)],
                    "l2": "flushed (256 MiB memset) before every timed launch"}
        prof = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.isfile(prof):
            try:
                roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_depth_map")
            except Exception:
                pass
        # hot path only (features resident): K4+K1+K2+K3 x 3 stages
        def hot_only():
            depth_l = None
            for l in (2, 1, 0):
                f = feats[f"level_{l}"]
                f = f.view(B, VIEWS, *f.shape[1:])
                D = N_DEPTHS[l]
                h, w = f.shape[-2:]
                if l == 2:
                    dv = ops.uniform_hypotheses(dmin, dint * RATIOS[l], D, B, h, w, dev)
                else:
                    dv = ops.depth_hypotheses(depth_l, D, dint * RATIOS[l], upsample=True)
                depth_l, _ = model.predict_depth(f, pm_d[:, :, l], dv, getattr(model, f"cost_reg_{l}"))
        with torch.no_grad():
            for _ in range(3):
                hot_only()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                hot_only()
            e1.record()
            torch.cuda.synchronize()
            hot_ms = e0.elapsed_time(e1) / args.steps
        hot = {"ms_per_depth_map": hot_ms, "depth_maps_per_s": 1e3 / hot_ms,
               "what": "features resident -> depth/confidence (K4,K1,K2,K3 x 3 stages), no FeatureNet"}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        avail = len(os.sched_getaffinity(0))
        cores = best_cpu_threads(avail)
        med, total = time_cpu_port(3, 1, cores)
        cpu_baseline = {"value": 1.0 / med, "unit": "depth-maps/s", "cores": cores, "kind": "port",
                        "sample": "1 warm-up + 3 timed full forwards of the same cfg2 workload "
                                  f"(median {med:.2f} s/depth-map), oracle port of the reference "
                                  f"PyTorch-CPU path on {cores} torch threads (fastest of "
                                  f"4..{avail} available)"}

    if rank == 0:
        maps = world * B * args.steps
        value = maps / (ms_total * 1e-3)
        h2d = imgs_h.numel() * 4 + pm_h.numel() * 4
        d2h = out_depth_h.numel() * 4 + out_conf_h.numel() * 4
        line = {
            "metric": METRIC, "value": value, "unit": "depth-maps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else f"f32 (conv products {args.precision}, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": "cfg2: 640x512, V=3, D=48/32/8, variance cost, B=1 per GPU "
                                   "(BASELINE.json configs[1])",
                       "step": ("CascadeMVSNet.forward = FeatureNet (cuDNN fp32 convs + fused FPN "
                                "kernel) + 3 cascade stages") if args.precision == "fp32" else
                               ("CascadeMVSNet.forward = FeatureNet (own kernels: planar tcgen05 "
                                "convs, RGB block, FPN merges) + 3 cascade stages (K4, K1, K2 x 11 "
                                "layers on tcgen05, K3)"),
                       "parallelism": f"dp{world} (independent reference views per rank, one "
                                      "all_gather of depth_0 per step)" if world > 1 else "single GPU",
                       "precision": args.precision,
                       "cuda_graph": not args.no_graph,
                       "l2": "per-step working set (>1 GB of intermediates) exceeds the 126 MB L2; "
                             "K1 roofline launches are preceded by an explicit L2 flush"},
            "e2e": {"value": maps / (ms_e2e * 1e-3), "unit": "depth-maps/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "how": ("pinned host inputs -> H2D -> CUDA-graph forward -> D2H of depth_0 + "
                            "confidence_2, every step; copies of neighbouring steps overlap compute "
                            "(2-slot pipeline)") if pipe is not None else
                           "pinned host inputs -> H2D -> forward -> D2H, serial"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "hot_path": hot,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
