"""Utterance-sharded inference — the loops of src/separate.py:39-57 and src/evaluate.py:42-71 across the GPUs of one box.

The reference runs them on one device: sort by length (src/data.py:55-59,207-209), batch, `model(mixture)`,
`remove_pad`, then per utterance on the host.  Here every rank (one process per GPU) takes a shard of the utterances and
runs the same forward on it; utterances are independent (gLN / cLN statistics are per utterance), so there is NO
collective on the data path.  The only optional exchange is the final (sum, count) all-reduce of the SI-SNRi metric.

    plan = shard_utterances(lengths, world, rank)          # which utterances this rank separates, longest first
    sep = ShardedSeparator(model)                          # rank / world from torch.distributed when initialised
    for index, est in sep.separate(mixtures):              # est [C, len(mixtures[index])] on the device
        ...
    avg_sisnri = sep.evaluate(mixtures, sources)           # src/evaluate.py's "Average SISNR improvement"
"""
import torch
import torch.distributed as dist

from .evaluate import cal_SISNRi_batch
from .pit_criterion import cal_loss


def shard_utterances(lengths, world, rank):
    """Indices of the utterances rank `rank` of `world` processes, longest first.  Longest-processing-time assignment:
    utterances sorted by length (descending, ties by index), each given to the rank with the least audio so far — the
    shards' total audio differs by at most one utterance.  Deterministic, so every rank computes the same plan."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        load[r] += int(lengths[i])
        if r == rank:
            mine.append(i)
    return mine


def batches_by_length(indices, lengths, max_batch, max_batch_samples):
    """Cut a length-sorted index list into batches of at most `max_batch` utterances and `max_batch_samples` padded
    samples (batch size x longest member) — neighbours in the sorted order have similar lengths, so padding is small."""
    out, cur = [], []
    for i in indices:
        longest = int(lengths[cur[0]]) if cur else int(lengths[i])
        if cur and (len(cur) + 1 > max_batch or (len(cur) + 1) * longest > max_batch_samples):
            out.append(cur)
            cur = []
        cur.append(i)
    if cur:
        out.append(cur)
    return out


class ShardedSeparator:
    def __init__(self, model, rank=None, world=None, max_batch=8, max_batch_seconds=480.0, sample_rate=8000):
        self.model = getattr(model, "module", model)
        ddp = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank() if ddp else 0)
        self.world = world if world is not None else (dist.get_world_size() if ddp else 1)
        self.max_batch = max_batch
        self.max_batch_samples = int(max_batch_seconds * sample_rate)

    def _device(self):
        return self.model.flat_params.device

    def _padded(self, tensors, idxs, lengths):
        """pad_list (src/data.py:322-331) of the batch members, built on the device"""
        dev = self._device()
        T = max(int(lengths[i]) for i in idxs)
        first = tensors[idxs[0]]
        out = torch.zeros((len(idxs),) + tuple(first.shape[:-1]) + (T,), dtype=torch.float32, device=dev)
        for b, i in enumerate(idxs):
            out[b, ..., :int(lengths[i])] = tensors[i].to(dev, non_blocking=True)
        return out

    def plan(self, lengths):
        mine = shard_utterances(lengths, self.world, self.rank)
        return batches_by_length(mine, lengths, self.max_batch, self.max_batch_samples)

    @torch.no_grad()
    def separate(self, mixtures):
        """mixtures: list of 1-D float tensors (any device).  Yields (index, est [C, length]) for this rank's shard, each
        trimmed to its utterance's length like remove_pad (src/utils.py:50-67)."""
        lengths = [int(m.shape[-1]) for m in mixtures]
        was_training = self.model.training
        self.model.eval()
        try:
            for idxs in self.plan(lengths):
                est = self.model(self._padded(mixtures, idxs, lengths))  # [B, C, T]
                for b, i in enumerate(idxs):
                    yield i, est[b, :, :lengths[i]]
        finally:
            self.model.train(was_training)

    @torch.no_grad()
    def evaluate(self, mixtures, sources, reduce=True):
        """Average SI-SNR improvement over ALL utterances (src/evaluate.py:42-71): every rank scores its shard with the
        batched device metric (PIT-reordered estimate, ctn_sisnri), then one 2-float all-reduce forms the global mean.
        sources: list of [C, length] tensors."""
        lengths = [int(m.shape[-1]) for m in mixtures]
        dev = self._device()
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        was_training = self.model.training
        self.model.eval()
        try:
            for idxs in self.plan(lengths):
                mix = self._padded(mixtures, idxs, lengths)
                src = self._padded(sources, idxs, lengths)
                lens = torch.tensor([lengths[i] for i in idxs], dtype=torch.int64, device=dev)
                est = self.model(mix)
                _loss, _snr, _est, reordered = cal_loss(src, est, lens)
                sisnri = cal_SISNRi_batch(src, reordered, mix, lens)
                acc[0] += sisnri.double().sum()
                acc[1] += len(idxs)
        finally:
            self.model.train(was_training)
        if reduce and self.world > 1 and dist.is_initialized():
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        return (acc[0] / acc[1].clamp_min(1)).item()
