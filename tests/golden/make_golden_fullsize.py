"""fp64 truth and reference fp32 gradients for the remaining BASELINE configs at FULL size.

    python tests/golden/make_golden_fullsize.py cfg4     configs[3]: C=3, M=16 x 4 s training step, fp64 oracle  (~15 min, 8 cores)
    python tests/golden/make_golden_fullsize.py cfg5     configs[4]: two 60 s utterances of the 8 x 60 s batch, fp64 forward
    python tests/golden/make_golden_fullsize.py ref32    configs[1]: the REFERENCE's own fp32 gradients (imports /root/reference)

cfg4 / cfg5 use the same storage as make_golden_fp64.py (sub-sampled output, loss, per-parameter gradient norm + sampled
entries + the fp32 oracle's own error).  `ref32` runs the unmodified reference implementation (src/conv_tasnet.py +
src/pit_criterion.py) on the configs[1] batch in fp32 and stores, per parameter, max|g| over the whole tensor and the
entries at make_golden_fp64.sample_index — what the north star's "gradients must match to 1e-3 rel [against the
reference]" is measured against (tests/test_model_gpu.py writes the per-tensor table)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import conv_tasnet_oracle as O  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_fp64 import sample_index  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def chunked_step(cfg, sd, mix, src, lens, chunk):
    """O.train_step_grads over the batch in chunks of `chunk` samples (gLN and PIT are per sample, the loss is the batch
    mean: the gradient is the size-weighted mean of the chunks' gradients) — bounds the memory of the fp64 autograd"""
    M = mix.shape[0]
    loss, ests, snrs, grads = 0.0, [], [], None
    for b0 in range(0, M, chunk):
        sl = slice(b0, min(M, b0 + chunk))
        l, e, g, snr, _ = O.train_step_grads(cfg, sd, mix[sl], src[sl], lens[sl])
        w = (sl.stop - sl.start) / M
        loss = loss + l * w
        ests.append(e)
        snrs.append(snr)
        if grads is None:
            grads = {k: v * w for k, v in g.items()}
        else:
            for k, v in g.items():
                grads[k] += v * w
        print("chunk", b0, "done", flush=True)
    return loss, torch.cat(ests), grads, torch.cat(snrs)


def train_golden(cfg, M, T, seed_w, seed_x, name, stride):
    sd = O.init_state_dict(cfg, seed_w)
    mix, src, lens = O.synthetic_batch(M, T, cfg.C, cfg.L, seed_x)
    sd64 = {k: v.double() for k, v in sd.items()}
    loss, est, grads, max_snr = chunked_step(cfg, sd64, mix.double(), src.double(), lens, 4)
    loss32, est32, grads32, _ = chunked_step(cfg, sd, mix, src, lens, 4)
    arrays = dict(M=np.int64(M), T=np.int64(T), seed_w=np.int64(seed_w), seed_x=np.int64(seed_x),
                  loss=np.float64(loss.item()), max_snr=max_snr.numpy(), est_stride=np.int64(stride),
                  est_sub=est.numpy()[..., ::stride].astype(np.float32), est_abs_max=np.float64(est.abs().max().item()),
                  names=np.array(list(grads.keys())))
    gn, ref_err, samp = [], [], []
    for k, g in grads.items():
        f = g.flatten()
        idx = sample_index(f.numel())
        gn.append(f.norm().item())
        s64, s32 = f[idx], grads32[k].flatten()[idx].double()
        ref_err.append(((s32 - s64).norm() / s64.norm()).item())
        samp.append(s64.numpy().astype(np.float32))
    arrays.update(g_norm=np.array(gn), ref32_rel_l2=np.array(ref_err), g_samples=np.concatenate(samp),
                  g_sample_counts=np.array([len(s) for s in samp]))
    out = os.path.join(OUT, name)
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out) // 1024, "KiB; worst fp32-oracle rel-L2 error", max(ref_err), flush=True)


def forward_golden(cfg, lengths, seed_w, seed_x, name, stride):
    """fp64 forward of single utterances (gLN statistics are per utterance: utterance b of a batch equals the same
    utterance run alone)"""
    sd64 = {k: v.double() for k, v in O.init_state_dict(cfg, seed_w).items()}
    arrays = dict(seed_w=np.int64(seed_w), seed_x=np.int64(seed_x), est_stride=np.int64(stride),
                  lengths=np.array(lengths, dtype=np.int64))
    for i, T in enumerate(lengths):
        mix, _, _ = O.synthetic_batch(1, T, cfg.C, cfg.L, seed_x + i)
        with torch.no_grad():
            est = O.forward(cfg, sd64, mix.double(), training=False)
        arrays[f"est_sub{i}"] = est.numpy()[..., ::stride].astype(np.float32)
        arrays[f"est_abs_max{i}"] = np.float64(est.abs().max().item())
        print("utterance", i, T, "done", flush=True)
    out = os.path.join(OUT, name)
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out) // 1024, "KiB", flush=True)


def reference_fp32(cfg, M, T, seed_w, seed_x, name):
    import types
    ref = os.environ.get("CTN_REFERENCE", "/root/reference")
    sys.path.insert(0, ref)
    sys.path.insert(0, os.path.join(ref, "src"))
    for mod in ("librosa", "visdom"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    from src.conv_tasnet import ConvTasNet  # noqa: E402
    from src.pit_criterion import cal_loss  # noqa: E402
    sd = O.init_state_dict(cfg, seed_w)
    model = ConvTasNet(cfg.N, cfg.L, cfg.B, cfg.H, cfg.P, cfg.X, cfg.R, cfg.C, norm_type=cfg.norm_type,
                       causal=cfg.causal, mask_nonlinear=cfg.mask_nonlinear)
    model.load_state_dict(sd)
    model.train()
    mix, src, lens = O.synthetic_batch(M, T, cfg.C, cfg.L, seed_x)
    est = model(mix)
    loss, max_snr, est_m, _ = cal_loss(src, est, lens)
    loss.backward()
    names, gmax, samp = [], [], []
    for k, p in model.named_parameters():
        f = p.grad.flatten()
        idx = sample_index(f.numel())
        names.append(k)
        gmax.append(f.abs().max().item())
        samp.append(f[idx].numpy().astype(np.float32))
    out = os.path.join(OUT, name)
    np.savez_compressed(out, M=np.int64(M), T=np.int64(T), seed_w=np.int64(seed_w), seed_x=np.int64(seed_x),
                        loss=np.float64(loss.item()), names=np.array(names), g_abs_max=np.array(gmax),
                        g_samples=np.concatenate(samp), g_sample_counts=np.array([len(s) for s in samp]))
    print(out, os.path.getsize(out) // 1024, "KiB; reference fp32 loss", loss.item(), flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    what = sys.argv[1]
    if what == "cfg4":
        cfg = O.Config(**{**O.PAPER.as_dict(), "C": 3})
        train_golden(cfg, 16, 32000, 0, 1238, "paper_cfg4_fp64.npz", 149)
    elif what == "cfg5":
        forward_golden(O.PAPER, [480000, 479893], 0, 1239, "paper_cfg5_fp64.npz", 211)
    elif what == "ref32":
        reference_fp32(O.PAPER, 3, 32000, 0, 1234, "paper_cfg2_ref32.npz")
    else:
        raise SystemExit(__doc__)
