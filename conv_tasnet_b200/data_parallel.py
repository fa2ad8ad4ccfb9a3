"""Batch-sharded data parallelism: the B200 replacement for `torch.nn.DataParallel(model)` in src/train.py:83-85.

One process per GPU (torchrun), every rank holds the full weights and optimizer state and works on its own shard of
the batch (gLN/cLN/PIT are per-sample, so sharding is exact).  The only exchange per step is the gradient
all-reduce: the hand-written backward runs in R+2 stages (ctn_model_backward_stage), each finishing one contiguous
slice of the flat gradient buffer, and every finished slice is all-reduced (NCCL over NVLink/NVSwitch, AVG) on
NCCL's stream while the next stage computes.  No parameter broadcast per step, no hub GPU.

The wrapper exposes `.module`, `__call__`, `.parameters()`, `.train()/.eval()`, `.cuda()` so the reference's
solver.py (which expects a DataParallel-style object, solver.py:61,97,141,188,194) runs unchanged.
"""
import torch
import torch.distributed as dist
import torch.nn as nn


class ShardedDataParallel(nn.Module):
    def __init__(self, module, process_group=None, overlap=True, broadcast_parameters=True):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.overlap = overlap
        self._pending = []
        self._enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        if self._enabled:
            module._grad_sync = self._on_stage
            if broadcast_parameters:
                self.broadcast_parameters()

    def broadcast_parameters(self, src=0):
        """Make every replica start from rank `src`'s weights (once; DataParallel re-broadcasts every step)."""
        dist.broadcast(self.module.flat_params, src=src, group=self.process_group)
        if getattr(self.module, "_bns", None):  # BatchNorm branch: running statistics and batch counters too
            dist.broadcast(self.module._bn_state, src=src, group=self.process_group)
            dist.broadcast(self.module._bn_count, src=src, group=self.process_group)

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    # called by ConvTasNet._run_backward after each backward stage (stage >= 0) and once to drain (stage == -1)
    def _on_stage(self, model, stage):
        if stage >= 0:
            off, cnt = model.grad_bucket(stage)
            work = dist.all_reduce(model.flat_grads[off:off + cnt], op=dist.ReduceOp.AVG if _has_avg(model.flat_grads)
                                   else dist.ReduceOp.SUM, group=self.process_group, async_op=True)
            if not _has_avg(model.flat_grads):
                self._pending.append((work, model.flat_grads[off:off + cnt]))
            else:
                self._pending.append((work, None))
            if not self.overlap:
                self._drain()
        else:
            self._drain()

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
        for stage in range(self.module.R + 2):
            self._on_stage(self.module, stage)
        self._drain()


def _has_avg(t):
    return t.is_cuda  # NCCL implements AVG; gloo (CPU tests) does not


def shard_batch(rank, world, *tensors):
    """Contiguous dim-0 shard of each tensor for this rank (the scatter DataParallel did inside forward)."""
    out = []
    for t in tensors:
        n = t.shape[0]
        per = (n + world - 1) // world
        out.append(t[rank * per:min(n, (rank + 1) * per)])
    return out
