"""CUDA-graph wrapper for fixed-shape inference (the eval.py loop feeds the same
shapes for every reference view).  One forward is ~100 kernel launches of a few
tens of microseconds each, so replaying a captured graph removes the host launch
overhead from the critical path.  All kernels of libcasmvs.so are enqueued on the
caller's stream without synchronisation or allocation, hence capturable.
"""
from __future__ import annotations

import torch


class GraphedCascade:
    """`g = GraphedCascade(model, imgs, proj_mats, depth_min, depth_interval)` captures
    one forward; `g(imgs, proj_mats)` copies the new inputs into the static buffers,
    replays the graph and returns the (static) result tensors."""

    def __init__(self, model, imgs, proj_mats, init_depth_min, depth_interval, warmup=3):
        assert imgs.is_cuda and proj_mats.is_cuda
        for t in (init_depth_min, depth_interval):
            if torch.is_tensor(t) and not t.is_cuda:
                raise ValueError("GraphedCascade needs device tensors (or floats) for the depth "
                                 "parameters: a host->device copy cannot be captured")
        self.model = model
        self.imgs = imgs.clone()
        self.proj = proj_mats.clone()
        self.dmin, self.dint = init_depth_min, depth_interval
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):          # cuDNN autotune, weight packing, smem attributes
                model(self.imgs, self.proj, self.dmin, self.dint)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        # the capture below must not touch the operand-image builders' events: tell the library
        # that everything the warm-up built is complete
        _lib.check(_lib.load().casmvs_settle_weight_images(), "settle_weight_images")
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = model(self.imgs, self.proj, self.dmin, self.dint)
        # libcasmvs kernels recorded in the graph (each replay launches all of them)
        self.kernels_per_replay = _lib.launch_count() - n0
        # the graph embeds raw pointers to the library's tensor-core operand images: remember
        # which packed buffers it depends on and how many images each of them has
        self._deps = [(b, b.data_ptr(), self._image_count(b)) for b in self._packed_buffers()]

    def _packed_buffers(self):
        fn = getattr(self.model, "packed_buffers", None)
        return fn() if fn is not None else []

    @staticmethod
    def _image_count(b):
        from . import _lib
        import ctypes
        return _lib.load().casmvs_weight_image_count(ctypes.c_void_p(b.data_ptr()),
                                                     b.numel() * b.element_size())

    def _check_weights(self):
        from . import _lib
        now = {id(b) for b in self._packed_buffers()}
        ok = all(id(b) in now and b.data_ptr() == ptr and self._image_count(b) == n
                 for b, ptr, n in self._deps)
        if not ok:
            raise _lib.CasMVSError(
                "packed weights were re-created after this CUDA graph was captured (load_state_dict, "
                "a second model, ...): the graph references freed operand images; build a new "
                "GraphedCascade")

    def __call__(self, imgs=None, proj_mats=None):
        self._check_weights()
        if imgs is not None:
            self.imgs.copy_(imgs, non_blocking=True)
        if proj_mats is not None:
            self.proj.copy_(proj_mats, non_blocking=True)
        self.graph.replay()
        return self.out


class PipelinedCascade:
    """Host-buffer streaming inference: the eval loop's H2D copy of view i+1 and the D2H read of
    view i-1 overlap the graph replay of view i.  Every slot has its own static inputs, captured
    graph, pinned result buffers AND compute stream: consecutive views also overlap each other
    on the GPU (one cfg2 forward leaves ~15 % of the machine idle in its launch / tail gaps;
    three views in flight: 1.067 -> 0.92 ms per depth map, bench.py `throughput_modes`).
    `concurrent=False` replays all slots on the caller's stream (round-2 behaviour before this).

        pipe = PipelinedCascade(model, imgs_example, proj_example, depth_min, depth_interval)
        for imgs_h, proj_h in views:            # pinned host tensors
            done = pipe.submit(imgs_h, proj_h)  # returns the results of the view submitted
            ...                                 # `slots` calls earlier (or None while filling);
                                                # valid until `slots` further submits
        tail = pipe.drain()
    Results are (depth_0, confidence_2) pinned host tensors, what eval.py:224-226 reads back.
    """

    def __init__(self, model, imgs, proj_mats, init_depth_min, depth_interval, slots=3,
                 concurrent=True):
        self.slots = [GraphedCascade(model, imgs, proj_mats, init_depth_min, depth_interval,
                                     warmup=3 if i == 0 else 1) for i in range(slots)]
        self.copy_stream = torch.cuda.Stream()      # H2D
        self.d2h_stream = torch.cuda.Stream()       # D2H (separate: it waits on compute)
        cur = torch.cuda.current_stream()
        # one compute stream per slot (the captured graphs share nothing mutable: static inputs,
        # outputs and graph memory are per slot, weights and operand images are read-only)
        self.compute = [torch.cuda.Stream() if concurrent else cur for _ in range(slots)]
        for cs in self.compute:
            cs.wait_stream(cur)
        self.h2d_done = [torch.cuda.Event() for _ in range(slots)]
        self.compute_done = [torch.cuda.Event() for _ in range(slots)]
        # pinned result buffers: a ring twice as long as the slot ring, so that the tensors a
        # submit() returns are not the target of any copy enqueued by that same call -- they
        # stay valid until `slots` further submits
        self.out_h = []
        g0 = self.slots[0]
        for _ in range(2 * slots):
            self.out_h.append((torch.empty(g0.out["depth_0"].shape).pin_memory(),
                               torch.empty(g0.out["confidence_2"].shape).pin_memory()))
        self.d2h_done = [torch.cuda.Event() for _ in range(2 * slots)]
        self.slot_read = [None] * slots
        self.n = 0
        self.pending = []

    def submit(self, imgs_h, proj_h, keep=None):
        """keep: optional device tensor that receives a copy of depth_0 on the compute stream
        (multi-GPU runs gather the per-rank depth maps from it once, at the end)."""
        i = self.n % len(self.slots)
        g = self.slots[i]
        ret = None
        if len(self.pending) == len(self.slots):          # slot i still holds an older view
            ret = self._collect()
        with torch.cuda.stream(self.copy_stream):
            # the slot's previous replay must have consumed its inputs before they are overwritten
            self.copy_stream.wait_event(self.compute_done[i])
            g.imgs.copy_(imgs_h, non_blocking=True)
            g.proj.copy_(proj_h, non_blocking=True)
            self.h2d_done[i].record(self.copy_stream)
        cs = self.compute[i]
        cs.wait_event(self.h2d_done[i])
        if self.slot_read[i] is not None:
            cs.wait_event(self.slot_read[i])
        g._check_weights()
        with torch.cuda.stream(cs):
            g.graph.replay()
            if keep is not None:
                keep.copy_(g.out["depth_0"])
        self.compute_done[i].record(cs)
        j = self.n % len(self.out_h)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(self.compute_done[i])
            self.out_h[j][0].copy_(g.out["depth_0"], non_blocking=True)
            self.out_h[j][1].copy_(g.out["confidence_2"], non_blocking=True)
            self.d2h_done[j].record(self.d2h_stream)
        # the slot's next replay overwrites g.out: it must not start before this D2H has read it
        self.slot_read[i] = self.d2h_done[j]
        self.pending.append(j)
        self.n += 1
        return ret

    def run_resident(self, steps, keep=None, every=0, on_chunk=None):
        """`steps` forwards over the inputs already resident in the slots' static buffers (no host
        copies): slot k % slots replays on its own stream, the caller's stream joins them at the
        end.  keep(k): optional device tensor that receives depth_0 of forward k.
        on_chunk(k0, k1, events): called after every `every` forwards (and after the last one) with
        the events that complete forwards k0..k1-1 -- a multi-GPU caller starts the gather of that
        chunk on a side stream while the next forwards run."""
        cur = torch.cuda.current_stream()
        for cs in self.compute:
            cs.wait_stream(cur)
        k0 = 0
        for k in range(steps):
            i = k % len(self.slots)
            g = self.slots[i]
            if self.slot_read[i] is not None:       # a pending D2H of this slot's previous result
                self.compute[i].wait_event(self.slot_read[i])
            g._check_weights()
            with torch.cuda.stream(self.compute[i]):
                g.graph.replay()
                if keep is not None:
                    keep(k).copy_(g.out["depth_0"])
            self.compute_done[i].record(self.compute[i])
            if on_chunk is not None and every > 0 and ((k + 1) % every == 0 or k + 1 == steps):
                on_chunk(k0, k + 1, list(self.compute_done))
                k0 = k + 1
        for cs in self.compute:
            cur.wait_stream(cs)
        return self.slots[(steps - 1) % len(self.slots)].out if steps else None

    def _collect(self):
        j = self.pending.pop(0)
        self.d2h_done[j].synchronize()
        return self.out_h[j]

    def drain(self):
        out = []
        while self.pending:
            out.append(self._collect())
        return out
