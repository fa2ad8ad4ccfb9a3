"""CUDA-graph capture of the hot loop body (src/solver.py:188-196) and of inference.

The C ABI launches a static sequence of kernels (no host sync, no allocation, device-side Adam step counter), so one
training step — forward, PIT loss, backward, gradient all-reduce, clip, Adam — is captured once per input shape and
replayed with a single graph launch; inputs are copied into the graph's static buffers (that copy is the H2D transfer
when the caller hands over pinned host tensors)."""
import torch

from . import _lib
from .pit_criterion import _pit_forward_raw, cal_loss


class GraphedTrainStep:
    """step = GraphedTrainStep(model_or_dp, optimizer); loss = step(mixture, source, lengths)

    `optimizer` must be capturable (conv_tasnet_b200.optim.FusedAdam is).  Under data parallelism
    (ShardedDataParallel) the collectives stay outside the graphs and overlap them: graph 0 = forward + PIT loss + its
    gradient + backward stage 0, graphs 1..R+1 = the remaining backward stages; after each replay the finished slice of
    the flat gradient buffer is all-reduced asynchronously on NCCL's stream while the next graph runs; the last graph
    (clip + Adam) waits for the collectives.  Falls back to eager launches if capture is refused."""

    def __init__(self, model, optimizer, warmup=3):
        self.model, self.optimizer, self.warmup = model, optimizer, warmup
        self._shape, self._graph, self._graph2, self._static, self._loss = None, None, None, None, None
        self._stage_graphs, self._keep = None, None
        self.captured = False

    def _fwd_bwd(self, mix, src, lens):
        est = self.model(mix)
        loss, _, _, _ = cal_loss(src, est, lens)
        self.optimizer.zero_grad()
        loss.backward()
        return loss

    def _eager(self, mix, src, lens):
        loss = self._fwd_bwd(mix, src, lens)
        self.optimizer.step()
        return loss

    def _dp(self):
        return self.model if getattr(self.model, "_enabled", False) and hasattr(self.model, "all_reduce_flat") else None

    def _capture(self, mix, src, lens):
        dev = mix.device
        self._static = (torch.empty_like(mix, device=dev), torch.empty_like(src, device=dev),
                        torch.empty_like(lens, device=dev))
        for d, s in zip(self._static, (mix, src, lens)):
            d.copy_(s)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._eager(*self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        dp = self._dp()
        try:
            if dp is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._loss = self._eager(*self._static)
                self._graph, self._graph2 = g, None
            else:  # collectives stay eager, between the per-stage graphs
                self._capture_staged(dp)
            self.captured = True
        except Exception:  # capture refused: keep working, eagerly
            torch.cuda.synchronize(dev)
            self._graph, self._graph2, self._stage_graphs, self.captured = None, None, None, False
        self._shape = (tuple(mix.shape), tuple(src.shape))

    def _capture_staged(self, dp):
        """forward + loss + backward as R+2 graphs cut at the gradient-bucket boundaries (driving the C ABI directly:
        ctn_model_forward, ctn_pit_forward/backward, ctn_model_backward_stage), then the optimizer graph"""
        m = dp.module
        mix, src, lens = self._static
        lens = lens.to(torch.int64)
        L = _lib.lib()
        pool = torch.cuda.graph_pool_handle()
        if any(p.grad is None for p in m.parameters()):
            for p, v in zip(m.parameters(), m.grad_views()):
                p.grad = v
        graphs = []
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            est, ws, token = m._run_forward(mix, training=True)
            loss, _max_snr, _idx, coef, _ = _pit_forward_raw(src, est, lens, False)
            d_est = torch.empty_like(est)
            one = torch.ones(1, dtype=torch.float32, device=est.device)
            B, C, T = src.shape
            _lib.check(L.ctn_pit_backward(_lib.ptr(src), _lib.ptr(est), _lib.ptr(lens), _lib.ptr(coef), _lib.ptr(one),
                                          B, C, T, _lib.ptr(d_est), _lib.stream()))
            m._backward_stage(mix, d_est, ws, 0)
        graphs.append(g)
        for stage in range(1, m.R + 2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                m._backward_stage(mix, d_est, ws, stage)
            graphs.append(g)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, pool=pool):
            self.optimizer.step()
        self._keep = (est, ws, token, coef, d_est, one, lens)  # tensors the graphs reference
        self._loss = loss.view(())
        self._graph, self._graph2, self._stage_graphs = graphs[0], g2, graphs

    def __call__(self, mixture, source, lengths):
        if self._shape != (tuple(mixture.shape), tuple(source.shape)):
            self._capture(mixture.to(self._device_of(), non_blocking=True), source.to(self._device_of(), non_blocking=True),
                          torch.as_tensor(lengths).to(device=self._device_of(), dtype=torch.int64, non_blocking=True))
        if self._graph is None:
            dev = self._device_of()
            return self._eager(mixture.to(dev, non_blocking=True), source.to(dev, non_blocking=True),
                               torch.as_tensor(lengths).to(dev, non_blocking=True))
        for d, s in zip(self._static, (mixture, source, lengths)):
            if d.data_ptr() != (s.data_ptr() if isinstance(s, torch.Tensor) and s.is_cuda else -1):
                d.copy_(torch.as_tensor(s), non_blocking=True)
        if self._graph2 is None:
            self._graph.replay()
        else:
            dp = self._dp()
            for stage, g in enumerate(self._stage_graphs):
                g.replay()
                dp._on_stage(dp.module, stage)  # async all-reduce of the slice this stage finished
            dp._on_stage(dp.module, -1)         # the current stream waits for the collectives
            self._graph2.replay()
        return self._loss

    def _device_of(self):
        m = getattr(self.model, "module", self.model)
        return m.flat_params.device


class GraphedInference:
    """infer = GraphedInference(model); est = infer(mixture)   (est is a static buffer, overwritten by the next call)"""

    def __init__(self, model):
        self.model = model
        self._shape, self._graph, self._in, self._out = None, None, None, None

    def __call__(self, mixture):
        if self._shape != tuple(mixture.shape):
            m = getattr(self.model, "module", self.model)
            dev = m.flat_params.device
            self._in = torch.empty(mixture.shape, dtype=torch.float32, device=dev)
            self._in.copy_(mixture)
            with torch.no_grad():
                for _ in range(2):
                    self.model(self._in)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._out = self.model(self._in)
            self._graph, self._shape = g, tuple(mixture.shape)
        if self._in.data_ptr() != (mixture.data_ptr() if mixture.is_cuda else -1):
            self._in.copy_(mixture, non_blocking=True)
        self._graph.replay()
        return self._out
