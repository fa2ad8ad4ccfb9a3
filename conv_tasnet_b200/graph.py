"""CUDA-graph capture of the hot loop body (src/solver.py:188-196) and of inference.

The C ABI launches a static sequence of kernels (no host sync, no allocation, device-side Adam step counter), so one
training step — forward, PIT loss, backward, gradient all-reduce, clip, Adam — is captured once per input shape and
replayed with a single graph launch; inputs are copied into the graph's static buffers (that copy is the H2D transfer
when the caller hands over pinned host tensors).

Each captured shape owns its workspace (the model's grow-only cache may replace its buffer at any time; a graph holds raw
pointers), the warm-up before a capture leaves no trace (parameters, Adam state and BatchNorm statistics are restored),
and a change of the optimizer's hyper-parameters (src/solver.py:169-176 halves the learning rate through
`optimizer.load_state_dict`) triggers a re-capture, since the kernels take them by value."""
import collections
import warnings

import torch

from . import _lib
from .pit_criterion import _pit_forward_raw, cal_loss


class _Captured:
    """one input shape: static input buffers, the workspace the graphs point into, the graphs, the loss tensor"""

    def __init__(self):
        self.static = self.ws = self.graph = self.graph2 = self.stage_graphs = self.keep = self.loss = self.hyper = None
        self.stage_groups = None


class GraphedTrainStep:
    """step = GraphedTrainStep(model_or_dp, optimizer); loss = step(mixture, source, lengths)

    `optimizer` must be capturable (conv_tasnet_b200.optim.FusedAdam is).  Under data parallelism
    (ShardedDataParallel) the collectives stay outside the graphs and overlap them: graph 0 = forward + PIT loss + its
    gradient + backward stage 0, graphs 1..R+1 = the remaining backward stages; after each replay the finished slice of
    the flat gradient buffer is all-reduced asynchronously on NCCL's stream while the next graph runs; the last graph
    (clip + Adam) waits for the collectives.  If capture is refused the step runs eagerly and a warning says why
    (`strict=True` raises instead).  Up to `max_shapes` input shapes stay captured (least recently used evicted)."""

    def __init__(self, model, optimizer, warmup=3, max_shapes=2, strict=False, dp_graphs=None):
        self.model, self.optimizer, self.warmup = model, optimizer, warmup
        self.max_shapes, self.strict = max_shapes, strict
        # data parallel only: into how many graphs the R + 2 backward stages are grouped (each graph boundary is a point
        # where the gradients finished so far start their all-reduce; more graphs = more overlap but more graph launches
        # and NCCL kernels competing with 1-CTA-per-SM GEMMs).  None: CTN_DP_GRAPHS or 2; R + 2 = one graph per stage.
        self.dp_graphs = dp_graphs
        self._cap = collections.OrderedDict()  # shape key -> _Captured (graph is None: capture refused, run eagerly)
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

    def _module(self):
        return getattr(self.model, "module", self.model)

    def _snapshot(self):
        m, o = self._module(), self.optimizer
        snap = [m.flat_params.clone()]
        for name in ("exp_avg", "exp_avg_sq", "step_count", "grad_norm"):
            if hasattr(o, name):
                snap.append(getattr(o, name).clone())
        if getattr(m, "_bn_state", None) is not None:
            snap += [m._bn_state.clone(), m._bn_count.clone()]
        return snap

    def _restore(self, snap):
        m, o = self._module(), self.optimizer
        it = iter(snap)
        m.flat_params.copy_(next(it))
        for name in ("exp_avg", "exp_avg_sq", "step_count", "grad_norm"):
            if hasattr(o, name):
                getattr(o, name).copy_(next(it))
        if getattr(m, "_bn_state", None) is not None:
            m._bn_state.copy_(next(it))
            m._bn_count.copy_(next(it))

    def _hyper(self):
        return self.optimizer.hyper() if hasattr(self.optimizer, "hyper") else None

    def _capture(self, mix, src, lens):
        dev = mix.device
        m = self._module()
        c = _Captured()
        c.static = (torch.empty_like(mix, device=dev), torch.empty_like(src, device=dev),
                    torch.empty_like(lens, device=dev))
        for d, s in zip(c.static, (mix, src, lens)):
            d.copy_(s)
        c.ws = torch.empty(m.workspace_bytes(mix.shape[0], mix.shape[1], True), dtype=torch.uint8, device=dev)
        c.hyper = self._hyper()
        m._ws_override = c.ws
        try:
            if self.warmup > 0:  # side-effect free: the state the warm-up steps changed is put back
                snap = self._snapshot()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(self.warmup):
                        self._eager(*c.static)
                torch.cuda.current_stream(dev).wait_stream(side)
                self._restore(snap)
            torch.cuda.synchronize(dev)
            dp = self._dp()
            try:
                if dp is None:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        c.loss = self._eager(*c.static)
                    c.graph = g
                elif dp.peer_active():  # the exchange is a kernel of ours: the whole step is one graph
                    self._capture_peer(dp, c)
                else:  # collectives stay eager, between the per-stage graphs
                    self._capture_staged(dp, c)
            except Exception as e:  # capture refused: keep working, eagerly — but say so
                torch.cuda.synchronize(dev)
                c.graph = c.graph2 = c.stage_graphs = None
                if self.strict:
                    raise
                warnings.warn(f"GraphedTrainStep: CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly")
        finally:
            m._ws_override = None
        return c

    def _capture_peer(self, dp, c):
        """data parallel with the peer-memory exchange: forward + PIT + every backward stage + the one-kernel all-reduce
        (csrc/peer_reduce.cu) + clip + Adam in ONE graph — nothing stays on the host between the kernels"""
        m = dp.module
        mix, src, lens = c.static
        lens = lens.to(torch.int64)
        L = _lib.lib()
        if any(p.grad is None for p in m.parameters()):
            for p, v in zip(m.parameters(), m.grad_views()):
                p.grad = v
        # uneven shards: the weight world * M_local / M_global of this rank's loss gradient (see _capture_staged)
        if getattr(dp, "weight_by_batch", False):
            one = dp.grad_scale_for(mix.shape[0], mix.device).to(torch.float32).clone()
        else:
            one = torch.ones(1, dtype=torch.float32, device=mix.device)
        torch.cuda.synchronize(mix.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            est, ws, token = m._run_forward(mix, training=True)
            loss, _max_snr, _idx, coef, _ = _pit_forward_raw(src, est, lens, False)
            d_est = torch.empty_like(est)
            B, C, T = src.shape
            _lib.check(L.ctn_pit_backward(_lib.ptr(src), _lib.ptr(est), _lib.ptr(lens), _lib.ptr(coef), _lib.ptr(one),
                                          B, C, T, _lib.ptr(d_est), _lib.stream()))
            for stage in range(m.R + 2):
                m._backward_stage(mix, d_est, ws, stage)
            dp.peer_all_reduce()
            self.optimizer.step()
        c.keep = (est, ws, token, coef, d_est, one, lens)  # tensors the graph references
        c.loss = loss.view(())
        c.graph, c.graph2, c.stage_graphs, c.stage_groups = g, None, None, None

    def _capture_staged(self, dp, c):
        """forward + loss + backward as R+2 graphs cut at the gradient-bucket boundaries (driving the C ABI directly:
        ctn_model_forward, ctn_pit_forward/backward, ctn_model_backward_stage), then the optimizer graph"""
        m = dp.module
        mix, src, lens = c.static
        lens = lens.to(torch.int64)
        L = _lib.lib()
        pool = torch.cuda.graph_pool_handle()
        if any(p.grad is None for p in m.parameters()):
            for p, v in zip(m.parameters(), m.grad_views()):
                p.grad = v
        graphs = []
        # uneven shards: this rank's loss gradient is weighted by world * M_local / M_global (data_parallel.py); the shard
        # sizes of a captured shape are fixed, so the factor is formed once, eagerly (one 4-byte all-reduce), and baked in
        if getattr(dp, "weight_by_batch", False):
            one = dp.grad_scale_for(mix.shape[0], mix.device).to(torch.float32).clone()
        else:
            one = torch.ones(1, dtype=torch.float32, device=mix.device)
        torch.cuda.synchronize(mix.device)
        import os
        nstage = m.R + 2
        ng = self.dp_graphs if self.dp_graphs is not None else int(os.environ.get("CTN_DP_GRAPHS", "2"))
        ng = max(1, min(nstage, ng))
        # stage groups of (nearly) equal size, the first group takes the remainder: [0..a), [a..b), ...
        bounds = [round(i * nstage / ng) for i in range(ng + 1)]
        groups = [(bounds[i], bounds[i + 1]) for i in range(ng) if bounds[i + 1] > bounds[i]]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            est, ws, token = m._run_forward(mix, training=True)
            loss, _max_snr, _idx, coef, _ = _pit_forward_raw(src, est, lens, False)
            d_est = torch.empty_like(est)
            B, C, T = src.shape
            _lib.check(L.ctn_pit_backward(_lib.ptr(src), _lib.ptr(est), _lib.ptr(lens), _lib.ptr(coef), _lib.ptr(one),
                                          B, C, T, _lib.ptr(d_est), _lib.stream()))
            for stage in range(*groups[0]):
                m._backward_stage(mix, d_est, ws, stage)
        graphs.append(g)
        for lo, hi in groups[1:]:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                for stage in range(lo, hi):
                    m._backward_stage(mix, d_est, ws, stage)
            graphs.append(g)
        c.stage_groups = groups
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, pool=pool):
            self.optimizer.step()
        c.keep = (est, ws, token, coef, d_est, one, lens)  # tensors the graphs reference
        c.loss = loss.view(())
        c.graph, c.graph2, c.stage_graphs = graphs[0], g2, graphs

    def __call__(self, mixture, source, lengths):
        dev = self._device_of()
        key = (tuple(mixture.shape), tuple(source.shape))
        c = self._cap.get(key)
        if c is not None and c.hyper != self._hyper():  # lr / betas / clip norm are baked into the captured kernels
            del self._cap[key]
            c = None
        if c is None:
            while len(self._cap) >= max(1, self.max_shapes):
                self._cap.popitem(last=False)
            c = self._capture(mixture.to(dev, non_blocking=True), source.to(dev, non_blocking=True),
                              torch.as_tensor(lengths).to(device=dev, dtype=torch.int64, non_blocking=True))
            self._cap[key] = c
        else:
            self._cap.move_to_end(key)
        self.captured = c.graph is not None
        if c.graph is None:
            return self._eager(mixture.to(dev, non_blocking=True), source.to(dev, non_blocking=True),
                               torch.as_tensor(lengths).to(dev, non_blocking=True))
        for d, s in zip(c.static, (mixture, source, lengths)):
            if d.data_ptr() != (s.data_ptr() if isinstance(s, torch.Tensor) and s.is_cuda else -1):
                d.copy_(torch.as_tensor(s), non_blocking=True)
        if c.graph2 is None:
            c.graph.replay()
        else:
            dp = self._dp()
            for (lo, hi), g in zip(c.stage_groups, c.stage_graphs):
                g.replay()
                dp._on_stages(dp.module, lo, hi)  # async all-reduce of the (contiguous) slice these stages finished
            dp._on_stage(dp.module, -1)         # the current stream waits for the collectives
            c.graph2.replay()
        return c.loss

    def _device_of(self):
        return self._module().flat_params.device


class GraphedInference:
    """infer = GraphedInference(model); est = infer(mixture)   (est is a static buffer, overwritten by the next call)

    One captured graph (with its own workspace) per input shape, up to `max_shapes` (least recently used evicted)."""

    def __init__(self, model, max_shapes=4):
        self.model, self.max_shapes = model, max_shapes
        self._cap = collections.OrderedDict()  # shape -> (graph, static input, static output, workspace)

    def __call__(self, mixture):
        key = (tuple(mixture.shape), str(getattr(getattr(self.model, "module", self.model), "inference_dtype", None)))
        entry = self._cap.get(key)
        if entry is None:
            while len(self._cap) >= max(1, self.max_shapes):
                self._cap.popitem(last=False)
            m = getattr(self.model, "module", self.model)
            dev = m.flat_params.device
            x = torch.empty(mixture.shape, dtype=torch.float32, device=dev)
            x.copy_(mixture)
            ws = torch.empty(m.workspace_bytes(mixture.shape[0], mixture.shape[1], False), dtype=torch.uint8, device=dev)
            m._ws_override = ws
            try:
                with torch.no_grad():
                    for _ in range(2):
                        self.model(x)
                    torch.cuda.synchronize(dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        out = self.model(x)
            finally:
                m._ws_override = None
            entry = self._cap[key] = (g, x, out, ws)
        else:
            self._cap.move_to_end(key)
        g, x, out, _ = entry
        if x.data_ptr() != (mixture.data_ptr() if mixture.is_cuda else -1):
            x.copy_(mixture, non_blocking=True)
        g.replay()
        return out
