"""Public serving-style API: register batches of cloud pairs that live in HOST memory.

    reg = Registrar(net, batch=32, n_points=16384)        # net: pcd_reg_hregnet_b200.models.HRegNet on a CUDA device
    R, t = reg(src_host, dst_host)                        # [B,3,3], [B,3] pinned host tensors (final pose, level 1)

One call = H2D copy of the two clouds from pinned memory, the whole HRegNet forward (one CUDA-graph replay of the
~70 kernel launches of the path; captured once per (batch, n_points)), D2H copy of the poses.  Everything is
stream-ordered on one CUDA stream; the only host synchronisation is the final wait for the poses.

    for R, t in reg.map(batches):                         # batches: iterable of (src_host, dst_host) pinned tensors

is the throughput form: the H2D copy of batch i+1 runs on a copy stream into a staging buffer while batch i is being
registered (the graph reads fixed input buffers, so a 12.6 MB device-to-device copy in front of each replay moves the
staged clouds in), and the poses of batch i are read back while batch i+1 runs.  Every batch still pays its own H2D and
D2H copies; they are only taken off the critical path.  With `in_flight` = 2 (the default) the forward is captured
twice (two sets of input / activation buffers, shared weights) and consecutive batches replay on two streams: the
sampling kernels are latency-bound with one small cluster per cloud and the shared-MLP kernels fill the chip, so two
forwards half a step apart use SMs the other leaves idle (6.10 -> 5.55 ms per 32-pair batch on one B200,
tools/split_probe.py --pipelined).
"""
import gc

import torch


class Registrar:
    """net: HRegNet / Model_V2 / Model_V4 of this package on a CUDA device (eval mode is forced).
    batch, n_points: the fixed shape [batch, n_points, 3] of both clouds of a call.
    use_cuda_graph: capture the forward once and replay it (False: eager launches, one forward at a time).
    warmup: eager forwards before the capture.  in_flight: forwards map() keeps enqueued at once (captures of the
    forward, each replayed on its own stream; 1 = single stream).  slot: which capture of `net` this object is -- the
    lanes of map() number themselves; nets that draw from the host generator keep one set of buffers per slot."""

    def __init__(self, net, batch, n_points, use_cuda_graph=True, warmup=2, in_flight=2, slot=0):
        self.net = net.eval()
        self.device = next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("Registrar needs the net on a CUDA device: there is no CPU path (move it with .cuda())")
        self.batch, self.n_points = batch, n_points
        both = torch.zeros(2 * batch, n_points, 3, device=self.device)   # back to back: the forward stacks them as a view
        self.src, self.dst = both[:batch], both[batch:]
        self.R_host = torch.empty(batch, 3, 3).pin_memory()
        self.t_host = torch.empty(batch, 3).pin_memory()
        self.graph = None
        self.out = None
        self.use_cuda_graph = use_cuda_graph
        self._warm = warmup
        self.in_flight = max(1, int(in_flight))   # forwards map() keeps enqueued at once (each on its own stream)
        self._slot = slot                         # which capture of `net` this is (lanes of map(): 0, 1, ...)
        self._pipe = None           # lazily created state of map(): copy stream, staging buffers, result slots

    def _prologue(self):
        # host-side draws of the forward (Model_V2 / Model_V4: FineReg2's batch shuffles), ahead of a captured forward
        if self.use_cuda_graph and hasattr(self.net, "host_prologue"):
            self.net.host_prologue(self.batch, self.device, self._slot)

    def _forward(self):
        with torch.no_grad():
            return self.net(self.src, self.dst)

    def load(self, src_host, dst_host):
        self.src.copy_(src_host, non_blocking=True)
        self.dst.copy_(dst_host, non_blocking=True)

    def capture(self):
        """Warm up (folds BN, sets kernel attributes, fills the allocator) and capture the forward."""
        bind = getattr(self.net, "bind_host_draws", None) if self.use_cuda_graph else None
        if bind is not None:
            bind(self._slot)
        try:
            self._capture()
        finally:
            if bind is not None:
                bind(None)
        return self

    def _capture(self):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(self._warm):
                self._prologue()
                self.out = self._forward()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        if self.use_cuda_graph:
            # No garbage collection inside the capture: a collected CUDA graph / tensor cycle frees device memory
            # (cudaFree synchronises with the legacy stream), which invalidates a capture in progress
            # (cudaErrorStreamCaptureImplicit).  Collect what is pending first, then keep the collector off.
            gc.collect()
            was_enabled = gc.isenabled()
            gc.disable()
            try:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.out = self._forward()
            finally:
                if was_enabled:
                    gc.enable()

    def run_device(self):
        """Forward on the clouds already resident in self.src / self.dst; returns the device result dict."""
        if self.graph is not None:
            self._prologue()
            self.graph.replay()
        else:
            self.out = self._forward()
        return self.out

    def __call__(self, src_host, dst_host, sync=True):
        if self.graph is None and self.use_cuda_graph:
            self.capture()
        self.load(src_host, dst_host)
        out = self.run_device()
        self.R_host.copy_(out["rotation"][-1], non_blocking=True)
        self.t_host.copy_(out["translation"][-1], non_blocking=True)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return self.R_host, self.t_host

    # ---- pipelined form ------------------------------------------------------------------------------------------
    def _pipe_state(self):
        """Lanes of map(): lane 0 is this object; every further lane is a second capture of the same forward (own input
        buffers, own graph, own activations; the weights are shared) that replays on its own stream."""
        if self._pipe is None:
            main = torch.cuda.current_stream(self.device)
            lanes = []
            for i in range(self.in_flight if self.use_cuda_graph else 1):   # eager forwards are host-bound: one lane
                # lane 0 is this object, stored as None: a reference to self here would make a cycle, and a Registrar
                # that waits for the cycle collector may release its graphs in the middle of somebody's capture
                reg = None if i == 0 else Registrar(self.net, self.batch, self.n_points, self.use_cuda_graph,
                                                    self._warm, in_flight=1, slot=self._slot + i)
                if reg is not None:
                    reg.src.copy_(self.src)
                    reg.dst.copy_(self.dst)
                    if self.use_cuda_graph:
                        reg.capture()
                lane = dict(reg=reg, stream=torch.cuda.Stream(device=self.device) if self.in_flight > 1 and self.use_cuda_graph else None,
                            stage=(torch.empty_like(self.src), torch.empty_like(self.dst)),
                            staged=torch.cuda.Event(), stage_free=torch.cuda.Event())
                lane["stage_free"].record(main)
                lanes.append(lane)
            self._pipe = dict(
                copy=torch.cuda.Stream(device=self.device), lanes=lanes,
                slots=[(torch.empty(self.batch, 3, 3).pin_memory(), torch.empty(self.batch, 3).pin_memory(),
                        torch.cuda.Event()) for _ in range(self.in_flight + 2)])
        return self._pipe

    def map(self, batches, post=None):
        """Registers an iterable of (src_host, dst_host) pinned batches; yields (R_host, t_host) per batch, in order.
        Up to `in_flight` forwards are enqueued before the oldest result is waited for; with in_flight > 1 consecutive
        batches run on different streams, so the latency-bound sampling kernels of one batch (one small cluster per
        cloud, most SMs idle) run beside the tensor-core kernels of its neighbour.  Every batch is still copied H2D,
        registered in full and read back D2H.
        The yielded tensors are one of in_flight + 2 result slots: consume (or copy) them before asking for the batch
        after next.  `post(out)` -- optional -- is called right after each forward has been enqueued, on that forward's
        stream, with the device result dict (e.g. to enqueue a collective on the poses)."""
        if self.graph is None and self.use_cuda_graph:
            self.capture()
        P = self._pipe_state()
        lanes, slots = P["lanes"], P["slots"]
        main = torch.cuda.current_stream(self.device)
        for lane in lanes:
            if lane["stream"] is not None:
                lane["stream"].wait_stream(main)
        pending = []
        n = 0
        try:
            for src_host, dst_host in batches:
                lane = lanes[n % len(lanes)]
                run = lane["stream"] if lane["stream"] is not None else main
                reg = lane["reg"] if lane["reg"] is not None else self
                with torch.cuda.stream(P["copy"]):
                    P["copy"].wait_event(lane["stage_free"])            # the lane's previous batch has left its staging buffers
                    lane["stage"][0].copy_(src_host, non_blocking=True)
                    lane["stage"][1].copy_(dst_host, non_blocking=True)
                    lane["staged"].record(P["copy"])
                with torch.cuda.stream(run):
                    run.wait_event(lane["staged"])
                    reg.src.copy_(lane["stage"][0], non_blocking=True)
                    reg.dst.copy_(lane["stage"][1], non_blocking=True)
                    lane["stage_free"].record(run)
                    out = reg.run_device()
                    if post is not None:
                        post(out)
                    R_h, t_h, done = slots[n % len(slots)]
                    R_h.copy_(out["rotation"][-1], non_blocking=True)
                    t_h.copy_(out["translation"][-1], non_blocking=True)
                    done.record(run)
                pending.append((R_h, t_h, done))
                n += 1
                if len(pending) > len(lanes):
                    R_h, t_h, done = pending.pop(0)
                    done.synchronize()
                    yield R_h, t_h
        finally:                                   # also when the caller abandons the generator early
            for lane in lanes:
                if lane["stream"] is not None:
                    main.wait_stream(lane["stream"])
        for R_h, t_h, done in pending:
            done.synchronize()
            yield R_h, t_h
