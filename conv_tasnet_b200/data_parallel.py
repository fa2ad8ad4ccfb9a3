"""Batch-sharded data parallelism: the B200 replacement for `torch.nn.DataParallel(model)` in src/train.py:83-85.

One process per GPU (torchrun), every rank holds the full weights and optimizer state and works on its own shard of
the batch (gLN/cLN/PIT are per-sample, so sharding is exact).  The only exchange per step is the gradient
all-reduce: the hand-written backward runs in R+2 stages (ctn_model_backward_stage), each finishing one contiguous
slice of the flat gradient buffer, and every finished slice is all-reduced (NCCL over NVLink/NVSwitch) on NCCL's stream
while the next stage computes.  No parameter broadcast per step, no hub GPU.

Uneven shards (the reference's batches vary in size, src/data.py:84-108): every rank's loss is the mean over its OWN
items, so the global-batch-mean gradient is sum_r (M_r / M_global) g_r.  The wrapper all-reduces the local batch size
(4 bytes, asynchronous, no host sync) in every training forward and scales the incoming d_est of the model's backward
by world * M_r / M_global on the device; the AVG all-reduce of the gradients then yields exactly that weighted sum.

Peer-memory exchange (the default under an NCCL process group of at most 8 ranks on one node; `peer_reduce=False` or
CTN_PEER_REDUCE=0 keeps the bucketed NCCL all-reduces): the flat gradient buffer of every rank is a CUDA-IPC
allocation mapped by all ranks of the node, and the whole all-reduce is ONE hand-written kernel per step
(csrc/peer_reduce.cu: flag barrier, reduce-scatter with peer loads in rank order, all-gather with peer stores, flag
barrier) launched on the compute stream after the last backward stage — no NCCL call and nothing for the host to do, so
`graph.GraphedTrainStep` captures the data-parallel step as a single CUDA graph exactly like the single-GPU step.

The wrapper exposes `.module`, `__call__`, `.parameters()`, `.train()/.eval()`, `.cuda()` so the reference's
solver.py (which expects a DataParallel-style object, solver.py:61,97,141,188,194) runs unchanged.
"""
import ctypes
import os
import socket

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib


class _RawCuda:
    """a device allocation that torch did not make, exposed through __cuda_array_interface__ (zero-copy torch.as_tensor)"""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}


class PeerExchange:
    """This rank's exchange buffer (256-byte flag block + `n_floats` fp32) in CUDA-IPC memory, every other rank's buffer
    mapped into this process, and the one-kernel all-reduce over them (include/ctn_b200.h: ctn_peer_*).

    All ranks of the group must sit on one node (NVLink / NVSwitch peers) and construct it collectively."""

    FLAG_BYTES = 256

    def __init__(self, n_floats, device, group=None):
        L = _lib.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not 2 <= self.world <= 8:
            raise RuntimeError(f"PeerExchange: world size {self.world} not in [2, 8]")
        self.n = int(n_floats)
        self.device = torch.device(device)
        self._base, self._peers = None, []
        nbytes = self.FLAG_BYTES + 4 * self.n
        err = None
        try:
            with torch.cuda.device(self.device):
                p = ctypes.c_void_p()
                _lib.check(L.ctn_peer_alloc(nbytes, ctypes.byref(p)))
                self._base = p.value
                h = ctypes.create_string_buffer(64)
                _lib.check(L.ctn_peer_export(self._base, h))
                mine = (bytes(h.raw), socket.gethostname(), os.getpid())
        except Exception as e:  # every rank still takes part in the exchange below, then all of them give up together
            err, mine = e, (None, socket.gethostname(), os.getpid())
        table = [None] * self.world
        dist.all_gather_object(table, mine, group=group)
        if err is None and (any(t[0] is None for t in table) or len({t[1] for t in table}) != 1):
            err = RuntimeError("PeerExchange: a rank could not allocate / export its buffer, or the ranks are on "
                               "different hosts")
        bases = [None] * self.world
        if err is None:
            try:
                with torch.cuda.device(self.device):
                    for r, (handle, _host, _pid) in enumerate(table):
                        if r == self.rank:
                            bases[r] = self._base
                        else:
                            q = ctypes.c_void_p()
                            _lib.check(L.ctn_peer_open(handle, ctypes.byref(q)))
                            bases[r] = q.value
                            self._peers.append(q.value)
            except Exception as e:
                err = e
        oks = [None] * self.world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            self.close()
            raise RuntimeError(f"PeerExchange: set-up failed on some rank ({err})")
        self._bufs = (ctypes.c_void_p * self.world)(*[b + self.FLAG_BYTES for b in bases])
        self._flags = (ctypes.c_void_p * self.world)(*bases)
        self.tensor = torch.as_tensor(_RawCuda(self._base + self.FLAG_BYTES, self.n), device=self.device)
        self._flag_view = torch.as_tensor(_RawCuda(self._base, self.FLAG_BYTES // 4), device=self.device).view(torch.int32)

    def all_reduce(self, scale=None, offset=0, count=None):
        """sum over the ranks (x scale, default 1 / world) of floats [offset, offset + count) of the exchange buffers,
        left in every rank's buffer; one kernel on the current stream"""
        count = self.n - offset if count is None else count
        scale = 1.0 / self.world if scale is None else scale
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctn_peer_all_reduce(self._bufs, self._flags, self.rank, self.world, offset, count,
                                                      scale, _lib.stream()))

    def error(self):
        """True when a wait inside the kernel gave up (a rank never arrived); synchronises"""
        return bool(self._flag_view[18].item())

    def close(self):
        L = _lib.lib()
        for q in self._peers:
            L.ctn_peer_close(q)
        self._peers = []
        if self._base is not None:
            L.ctn_peer_free(self._base)
            self._base = None



class ShardedDataParallel(nn.Module):
    def __init__(self, module, process_group=None, overlap=True, broadcast_parameters=True, weight_by_batch=True,
                 peer_reduce=None):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.overlap = overlap
        self.weight_by_batch = weight_by_batch
        self._pending = []
        self._peer = None
        self._enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        if self._enabled:
            module._grad_sync = self._on_stage
            if broadcast_parameters:
                self.broadcast_parameters()
            if peer_reduce is None:  # default: on wherever it can work (NCCL group = one GPU per rank, at most 8 of them)
                env = os.environ.get("CTN_PEER_REDUCE")
                auto = (dist.get_backend(process_group) == "nccl" and dist.get_world_size(process_group) <= 8
                        and module.flat_params.is_cuda)
                if env is not None:
                    peer_reduce = env == "1"
                elif auto:  # the set-up fails on every rank or on none (PeerExchange agrees on it): NCCL otherwise
                    try:
                        self.enable_peer_reduce()
                    except RuntimeError as e:
                        import warnings
                        warnings.warn(f"ShardedDataParallel: peer-memory gradient exchange unavailable ({e}); using NCCL")
                    peer_reduce = False
            if peer_reduce:
                self.enable_peer_reduce()

    def enable_peer_reduce(self):
        """Move the flat gradient buffer into CUDA-IPC memory shared with the other ranks of the node and exchange
        gradients with the one-kernel peer all-reduce instead of NCCL (collective: every rank calls it)."""
        m = self.module
        flat = m.flat_params
        self._peer = PeerExchange(flat.numel(), flat.device, self.process_group)
        for p in m.parameters():
            p.grad = None
        m._flat_grad = self._peer.tensor

    def peer_active(self):
        """the model's flat gradient buffer still is the shared one (a re-flatten, e.g. after .to(), drops it)"""
        return self._peer is not None and self.module.flat_grads.data_ptr() == self._peer.tensor.data_ptr()

    def peer_all_reduce(self):
        self._peer.all_reduce()

    def broadcast_parameters(self, src=0):
        """Make every replica start from rank `src`'s weights (once; DataParallel re-broadcasts every step)."""
        dist.broadcast(self.module.flat_params, src=src, group=self.process_group)
        if getattr(self.module, "_bns", None):  # BatchNorm branch: running statistics and batch counters too
            dist.broadcast(self.module._bn_state, src=src, group=self.process_group)
            dist.broadcast(self.module._bn_count, src=src, group=self.process_group)

    def grad_scale_for(self, m_local, device):
        """world * M_local / M_global as a 1-element device tensor (no host sync): the factor that turns the AVG
        all-reduce of per-rank batch-mean gradients into the gradient of the GLOBAL batch mean."""
        world = dist.get_world_size(self.process_group)
        total = torch.full((1,), float(m_local), dtype=torch.float32, device=device)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.process_group)
        return (float(m_local) * world) / total

    def forward(self, *inputs, **kwargs):
        if self._enabled and self.weight_by_batch and torch.is_grad_enabled() and inputs:
            x = inputs[0]
            if x.shape[0] < 1:
                raise ValueError("ShardedDataParallel: this rank's shard is empty; every rank needs at least one item "
                                 "(shard_batch refuses such splits on all ranks)")
            self.module._grad_scale = self.grad_scale_for(x.shape[0], x.device)
        return self.module(*inputs, **kwargs)

    # called by ConvTasNet._run_backward after each backward stage (stage >= 0), once to drain (stage == -1), and with
    # stage == -2 + an explicit buffer from the slow path (gradients computed aside): synchronous whole-buffer reduce
    def _on_stage(self, model, stage, buf=None):
        if stage != -2 and self.peer_active():  # one kernel for the whole buffer, after the last stage
            if stage == -1:
                self._peer.all_reduce()
            return
        if stage == -2:
            self._reduce(buf)
            self._drain()
        elif stage >= 0:
            off, cnt = model.grad_bucket(stage)
            self._reduce(model.flat_grads[off:off + cnt])
            if not self.overlap:
                self._drain()
        else:
            self._drain()

    def _on_stages(self, model, lo, hi):
        """all-reduce the gradient slices of backward stages [lo, hi) as ONE contiguous slice (the buckets of consecutive
        stages are adjacent in the flat buffer, in descending order) — used by graph.GraphedTrainStep"""
        spans = [model.grad_bucket(s) for s in range(lo, hi)]
        start = min(o for o, _ in spans)
        end = max(o + c for o, c in spans)
        assert sum(c for _, c in spans) == end - start, "stage buckets are not contiguous"
        self._reduce(model.flat_grads[start:end])
        if not self.overlap:
            self._drain()

    def _reduce(self, view):
        avg = _has_avg(view, self.process_group)
        work = dist.all_reduce(view, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.process_group,
                               async_op=True)
        self._pending.append((work, None if avg else view))

    def _drain(self):
        world = dist.get_world_size(self.process_group)
        for work, view in self._pending:
            work.wait()  # on CUDA: the current stream waits for NCCL's stream, no host block
            if view is not None:
                view.div_(world)
        self._pending = []

    def all_reduce_flat(self):
        """Un-overlapped variant: all-reduce the whole flat gradient buffer in R+2 bucket calls."""
        if not self._enabled:
            return
        if self.peer_active():
            self._peer.all_reduce()
            return
        for stage in range(self.module.R + 2):
            self._on_stage(self.module, stage)
        self._drain()


def _has_avg(t, group=None):
    return t.is_cuda and dist.get_backend(group) == "nccl"  # NCCL implements AVG; gloo does not


def shard_sizes(n, world):
    """Items per rank for a batch of n: as even as possible, larger shards first (like torch.chunk's scatter in
    nn.DataParallel).  Raises on every rank alike when some rank would get nothing."""
    if n < world:
        raise ValueError(f"cannot shard a batch of {n} over {world} ranks: every rank needs at least one item")
    base, extra = divmod(n, world)
    return [base + (1 if r < extra else 0) for r in range(world)]


def shard_batch(rank, world, *tensors):
    """Contiguous dim-0 shard of each tensor for this rank (the scatter DataParallel did inside forward)."""
    out = []
    for t in tensors:
        sizes = shard_sizes(t.shape[0], world)
        start = sum(sizes[:rank])
        out.append(t[start:start + sizes[rank]])
    return out
