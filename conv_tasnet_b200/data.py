"""Device-side batch assembly: the work of src/data.py::_collate_fn / _collate_fn_eval + pad_list
(:159-183, :239-260, :322-331) and of the three `.cuda()` copies in the training / evaluation loops
(src/solver.py:184-187, src/evaluate.py:44-47), done with ONE pinned staging buffer, ONE async host->device copy of
the un-padded samples and one kernel (ctn_assemble_batch) that pads, transposes the sources and writes the lengths.
"""
import numpy as np
import torch

from . import _lib


def _al16(n):
    return (n + 15) & ~15


class DeviceBatchAssembler:
    """`assembler(mixtures, sources)` -> (padded_mixture [B,T], lengths [B] int64, padded_source [B,C,T]) on the GPU,
    the triple `_collate_fn` returns (src/data.py:183) followed by solver.py's `.cuda()` calls.

    mixtures: list of B 1-D arrays (T_b samples each, any float dtype); sources: list of B [T_b, C] arrays (the loader's
    layout, src/data.py:264-300) or None for the evaluation collate (then padded_source is None).
    Two pinned staging buffers alternate so that packing batch i+1 on the host overlaps the copy of batch i."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceBatchAssembler needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._pinned = [None, None]
        self._event = [None, None]
        self._dev = None
        self._turn = 0

    def _buffers(self, nbytes):
        i = self._turn
        self._turn ^= 1
        if self._pinned[i] is None or self._pinned[i].numel() < nbytes:
            self._pinned[i] = torch.empty(int(nbytes * 1.25) + 64, dtype=torch.uint8).pin_memory()
            self._event[i] = None
        if self._event[i] is not None:
            self._event[i].synchronize()  # the copy that last read this staging buffer has finished
        if self._dev is None or self._dev.numel() < nbytes:
            self._dev = torch.empty(int(nbytes * 1.25) + 64, dtype=torch.uint8, device=self.device)
        return i, self._pinned[i], self._dev

    def __call__(self, mixtures, sources=None):
        B = len(mixtures)
        if B == 0:
            raise ValueError("empty batch")
        lens = [int(np.shape(m)[0]) for m in mixtures]
        T, total = max(lens), sum(lens)
        if T == 0:
            raise ValueError("every utterance of the batch is empty")
        C = 1
        if sources is not None:
            if len(sources) != B:
                raise ValueError("mixtures and sources differ in batch size")
            C = int(np.shape(sources[0])[1])
            for s, n in zip(sources, lens):
                if np.shape(s) != (n, C):
                    raise ValueError(f"source of shape {np.shape(s)} does not match its mixture ({n} samples, C={C})")
        o_off, o_mix = 0, _al16(8 * (B + 1))
        o_src = o_mix + _al16(4 * total)
        nbytes = o_src + (_al16(4 * total * C) if sources is not None else 0)
        with torch.cuda.device(self.device):
            i, pinned, dev = self._buffers(nbytes)
            host = pinned.numpy()
            offs = host[o_off:o_off + 8 * (B + 1)].view(np.int64)
            offs[0] = 0
            np.cumsum(lens, out=offs[1:])
            mix_h = host[o_mix:o_mix + 4 * total].view(np.float32)
            src_h = host[o_src:o_src + 4 * total * C].view(np.float32).reshape(total, C) if sources is not None else None
            for b, n in enumerate(lens):
                mix_h[offs[b]:offs[b] + n] = mixtures[b]
                if src_h is not None:
                    src_h[offs[b]:offs[b] + n] = sources[b]
            dev[:nbytes].copy_(pinned[:nbytes], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._event[i] = ev
            mix = torch.empty(B, T, dtype=torch.float32, device=self.device)
            src = torch.empty(B, C, T, dtype=torch.float32, device=self.device) if sources is not None else None
            lengths = torch.empty(B, dtype=torch.int64, device=self.device)
            base = dev.data_ptr()
            _lib.check(_lib.lib().ctn_assemble_batch(base + o_mix, base + o_src if sources is not None else None,
                                                     base + o_off, B, C, T, _lib.ptr(mix), _lib.ptr(src),
                                                     _lib.ptr(lengths), _lib.stream()))
        return mix, lengths, src
