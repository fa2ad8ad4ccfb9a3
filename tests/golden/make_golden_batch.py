"""Golden vectors for the batch assembly (pad_list / collate / remove_pad), produced by RUNNING THE REFERENCE's own
functions in the build container (src/data.py needs librosa at import time, which is absent: a stub module stands in,
the functions used here never touch it):

    python tests/golden/make_golden_batch.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CTN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.modules.setdefault("librosa", types.ModuleType("librosa"))
from src.data import pad_list  # noqa: E402
from src.utils import remove_pad  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(3)
    arrays = {}
    cases = [([403, 1, 257, 403], 2), ([3200, 3137, 17], 2), ([5], 3), ([64, 0, 9], 1)]
    for i, (lens, C) in enumerate(cases):
        mixtures = [rng.standard_normal(n).astype(np.float32) for n in lens]
        sources = [rng.standard_normal((n, C)).astype(np.float32) for n in lens]
        # _collate_fn, src/data.py:172-183
        mix_pad = pad_list([torch.from_numpy(m).float() for m in mixtures], 0)
        src_pad = pad_list([torch.from_numpy(s).float() for s in sources], 0).permute((0, 2, 1)).contiguous()
        lengths = torch.from_numpy(np.array(lens))
        arrays[f"c{i}_packed_mix"] = np.concatenate(mixtures)
        arrays[f"c{i}_packed_src"] = np.concatenate(sources, axis=0)
        arrays[f"c{i}_lengths"] = np.array(lens, dtype=np.int64)
        arrays[f"c{i}_mix_pad"] = mix_pad.numpy()
        arrays[f"c{i}_src_pad"] = src_pad.numpy()
        rp3 = remove_pad(src_pad, lengths)
        rp2 = remove_pad(mix_pad, lengths)
        arrays[f"c{i}_rp3"] = np.concatenate([r.reshape(-1) for r in rp3])
        arrays[f"c{i}_rp2"] = np.concatenate([r.reshape(-1) for r in rp2])
        arrays[f"c{i}_rp3_shapes"] = np.array([r.shape for r in rp3], dtype=np.int64)
    arrays["n_cases"] = np.int64(len(cases))
    path = os.path.join(OUT, "batch.npz")
    np.savez_compressed(path, **arrays)
    print(f"batch.npz: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
