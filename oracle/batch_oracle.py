"""CPU oracle for the batch assembly next to the hot path (TEST INFRASTRUCTURE, not product; numpy).

Restates src/data.py::pad_list (:322-331) + the collate functions' tensor work (:172-183, :252-260) and
src/utils.py::remove_pad (:50-67).  Pinned by tests/golden/batch.npz, produced by running the reference's own
`pad_list` / `remove_pad` (tests/golden/make_golden_batch.py)."""
import numpy as np


def pad_list(xs, pad_value=0.0):
    """list of B arrays [T_b, ...] -> [B, max T_b, ...] filled with pad_value (src/data.py:322-331)"""
    max_len = max(x.shape[0] for x in xs)
    out = np.full((len(xs), max_len) + tuple(xs[0].shape[1:]), pad_value, dtype=np.float32)
    for i, x in enumerate(xs):
        out[i, :x.shape[0]] = x
    return out


def collate(mixtures, sources=None):
    """-> (padded_mixture [B,T] f32, lengths [B] i64, padded_source [B,C,T] f32 or None) (src/data.py:172-183)"""
    lengths = np.array([m.shape[0] for m in mixtures], dtype=np.int64)
    mix = pad_list([np.asarray(m, dtype=np.float32) for m in mixtures])
    src = None
    if sources is not None:
        src = np.ascontiguousarray(pad_list([np.asarray(s, dtype=np.float32) for s in sources]).transpose(0, 2, 1))
    return mix, lengths, src


def remove_pad(inputs, lengths):
    """[B,C,T] or [B,T] + lengths -> list of [C, T_b] / [T_b] arrays (src/utils.py:50-67)"""
    out = []
    for x, n in zip(inputs, lengths):
        n = int(n)
        out.append(x[:, :n].reshape(x.shape[0], -1).copy() if inputs.ndim == 3 else x[:n].reshape(-1).copy())
    return out
