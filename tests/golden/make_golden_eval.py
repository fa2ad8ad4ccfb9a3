"""Golden vectors for the evaluation metric (cal_SISNRi / cal_SISNR, src/evaluate.py:94-130), produced by RUNNING THE
REFERENCE's own functions in the build container (src/evaluate.py imports librosa / mir_eval / the data pipeline at
import time; they are absent here and unused by these two functions, so stub modules stand in):

    python tests/golden/make_golden_eval.py
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("CTN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "src"))  # src/evaluate.py imports its siblings as top-level modules
for name in ("librosa", "mir_eval", "mir_eval.separation", "visdom"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["mir_eval.separation"].bss_eval_sources = None
sys.modules["mir_eval"].separation = sys.modules["mir_eval.separation"]
from src.evaluate import cal_SISNR, cal_SISNRi  # noqa: E402
from src.utils import remove_pad  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    import torch
    rng = np.random.default_rng(11)
    arrays = {}
    cases = [(4, 900, [900, 640, 900, 123], 0.02), (3, 8000, [8000, 7977, 4000], 0.3), (2, 257, [257, 31], 1e-3)]
    for i, (B, T, lens, noise) in enumerate(cases):
        src = (rng.standard_normal((B, 2, T)) * 0.05).astype(np.float32)
        for b, n in enumerate(lens):
            src[b, :, n:] = 0
        mix = src.sum(1)
        est = (src + noise * 0.05 * rng.standard_normal((B, 2, T))).astype(np.float32)
        lengths = torch.tensor(lens)
        want, each = [], []
        # the reference's evaluation loop: remove_pad, then the metric per utterance (src/evaluate.py:53-63)
        for s, e, m in zip(remove_pad(torch.from_numpy(src), lengths), remove_pad(torch.from_numpy(est), lengths),
                           remove_pad(torch.from_numpy(mix), lengths)):
            want.append(cal_SISNRi(s, e, m))
            each.append([cal_SISNR(s[0], e[0]), cal_SISNR(s[1], e[1])])
        arrays[f"c{i}_src"], arrays[f"c{i}_est"], arrays[f"c{i}_mix"] = src, est, mix
        arrays[f"c{i}_lengths"] = np.array(lens, dtype=np.int64)
        arrays[f"c{i}_sisnri"] = np.array(want, dtype=np.float64)
        arrays[f"c{i}_sisnr"] = np.array(each, dtype=np.float64)
    arrays["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **arrays)
    print({k: v.shape for k, v in arrays.items()}, arrays["c0_sisnri"], arrays["c2_sisnri"])


if __name__ == "__main__":
    main()
