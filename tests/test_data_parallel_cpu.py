"""world_size-2 gloo test (CPU) of the data-parallel host logic: bucketed gradient all-reduce over the staged
backward's flat slices, parameter broadcast, batch sharding.  The CUDA kernels are not involved (no GPU here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conv_tasnet_b200 import ConvTasNet
    from conv_tasnet_b200.data_parallel import ShardedDataParallel, shard_batch
    torch.manual_seed(100 + rank)  # different initial weights per rank on purpose
    model = ConvTasNet(16, 8, 8, 16, 3, 3, 2, 2)
    dp = ShardedDataParallel(model, overlap=overlap)
    assert dp.module is model and list(dp.parameters())[0] is list(model.parameters())[0]
    # broadcast made the replicas identical to rank 0
    ref = [torch.empty_like(model.flat_params) for _ in range(world)]
    dist.all_gather(ref, model.flat_params)
    assert torch.equal(ref[0], ref[1])
    # staged all-reduce averages every slice of the flat gradient exactly once
    g = model.flat_grads
    g.copy_(torch.arange(g.numel(), dtype=torch.float32) * (rank + 1))
    for stage in range(model.R + 2):
        model._grad_sync(model, stage)
    model._grad_sync(model, -1)
    want = torch.arange(g.numel(), dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    assert torch.allclose(g, want, rtol=1e-6), (g - want).abs().max()
    assert dp._pending == []
    # shard_batch = the scatter DataParallel did
    x = torch.arange(6).view(6, 1)
    (mine,) = shard_batch(rank, world, x)
    assert mine.flatten().tolist() == [3 * rank, 3 * rank + 1, 3 * rank + 2]
    # uneven shards (the reference's batches vary in size, src/data.py:84-108): every rank's gradient is the mean over
    # its own items; scaled by world * M_r / M_global and averaged, the result is the mean over the global batch
    from conv_tasnet_b200.data_parallel import shard_sizes
    assert shard_sizes(7, 2) == [4, 3] and shard_sizes(5, 3) == [2, 2, 1] and shard_sizes(6, 2) == [3, 3]
    items = torch.arange(7, dtype=torch.float32)            # per-item "gradients" 0..6: global mean 3
    (mine,) = shard_batch(rank, world, items)
    assert mine.tolist() == ([0, 1, 2, 3] if rank == 0 else [4, 5, 6])
    scale = dp.grad_scale_for(mine.numel(), torch.device("cpu"))
    assert abs(scale.item() - world * mine.numel() / 7.0) < 1e-6
    g.fill_(0)
    g += mine.mean() * scale                                 # what the scaled backward leaves in the flat buffer
    for stage in range(model.R + 2):
        model._grad_sync(model, stage)
    model._grad_sync(model, -1)
    assert torch.allclose(g, torch.full_like(g, 3.0), rtol=1e-6), g[:4]
    # a batch smaller than the world is refused on every rank alike (no rank may wait in a collective for another)
    with pytest.raises(ValueError):
        shard_batch(rank, world, torch.zeros(1, 3))
    # the slow path (gradients computed aside): one synchronous whole-buffer reduce
    aside = torch.full((5,), float(rank + 1))
    model._grad_sync(model, -2, aside)
    assert torch.allclose(aside, torch.full((5,), (1 + world) / 2.0)) and dp._pending == []
    # utterance-sharded inference plans: disjoint, complete, balanced, identical on every rank
    from conv_tasnet_b200.separate import batches_by_length, shard_utterances
    lens = [480000, 31000, 64000, 479000, 8000, 250000, 250001, 12]
    shards = [shard_utterances(lens, world, r) for r in range(world)]
    assert sorted(sum(shards, [])) == list(range(len(lens)))
    loads = [sum(lens[i] for i in sh) for sh in shards]
    assert abs(loads[0] - loads[1]) <= max(lens)
    assert all(lens[a] >= lens[b] for sh in shards for a, b in zip(sh, sh[1:]))  # longest first within a shard
    for sh in shards:
        for batch in batches_by_length(sh, lens, 3, 1000000):
            assert len(batch) <= 3 and len(batch) * lens[batch[0]] <= 1000000 or len(batch) == 1
    # grad_views alias the flat buffer in parameters() order
    views = model.grad_views()
    assert sum(v.numel() for v in views) == sum(p.numel() for p in model.parameters())
    assert views[0].data_ptr() == g.data_ptr()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_bucketed_allreduce_world2_gloo(overlap):
    mp.spawn(_worker, args=(2, _free_port(), overlap), nprocs=2, join=True)
