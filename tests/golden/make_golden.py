"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference modules `src.conv_tasnet`, `src.pit_criterion`, `src.utils`
(pure torch) and records inputs, weights and outputs as small .npz files.  The tests
then compare the oracle (CPU) and the CUDA path (GPU) with these files; nothing at test time reads
/root/reference.
"""
import io
import os
import sys
import contextlib

import numpy as np
import torch

REF = os.environ.get("CTN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from src.conv_tasnet import ConvTasNet  # noqa: E402
from src.pit_criterion import cal_loss, cal_si_snr_with_pit, reorder_source  # noqa: E402
from src.utils import overlap_and_add  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def np_(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------- overlap_and_add
def golden_ola():
    arrays = {}
    # the reference's own print-only known answer (src/utils.py:70-77), seed 123
    torch.manual_seed(123)
    sig = torch.randint(5, (2, 2, 3, 4))
    with contextlib.redirect_stdout(io.StringIO()):
        res = overlap_and_add(sig, 2)
    arrays["kat_signal"], arrays["kat_step"], arrays["kat_result"] = np_(sig), np.int64(2), np_(res)
    g = torch.Generator().manual_seed(7)
    for i, (shape, step) in enumerate([((3, 2, 57, 20), 10), ((2, 1, 9, 16), 8), ((1, 3, 11, 8), 4),
                                       ((2, 2, 6, 6), 2), ((1, 1, 5, 5), 2), ((4, 1, 20), 10)]):
        sig = torch.randn(shape, generator=g)
        arrays[f"r{i}_signal"], arrays[f"r{i}_step"] = np_(sig), np.int64(step)
        arrays[f"r{i}_result"] = np_(overlap_and_add(sig, step))
    arrays["n_random"] = np.int64(6)
    save("ola.npz", **arrays)


# ---------------------------------------------------------------- PIT SI-SNR
def golden_pit():
    arrays = {}
    # the reference's seed-123 smoke block (src/pit_criterion.py:117-133) -> loss 45.9221
    torch.manual_seed(123)
    B, C, T = 2, 3, 32000
    source = torch.randint(4, (B, C, T)).float()
    est = torch.randint(4, (B, C, T)).float()
    source[1, :, -3:] = 0
    est[1, :, -3:] = 0
    lengths = torch.LongTensor([T, T - 3])
    loss, max_snr, est_m, reord = cal_loss(source, est.clone(), lengths)
    arrays.update(kat_source=np_(source).astype(np.int8), kat_est=np_(est).astype(np.int8),
                  kat_lengths=np_(lengths), kat_loss=np_(loss), kat_max_snr=np_(max_snr),
                  kat_reorder_crc=np.float64(reord.double().sum().item()))
    g = torch.Generator().manual_seed(11)
    cases = [(3, 2, 1000, [1000, 1000, 763]), (4, 3, 777, [777, 500, 777, 1]), (1, 2, 64, [64]),
             (5, 3, 320, [320, 320, 300, 17, 320]), (2, 4, 200, [200, 150]), (2, 1, 50, [50, 40])]
    for i, (B, C, T, lens) in enumerate(cases):
        src = torch.randn(B, C, T, generator=g) * 0.05
        perm = torch.stack([torch.randperm(C, generator=g) for _ in range(B)])
        est = torch.gather(src, 1, perm.view(B, C, 1).expand(B, C, T)) + 0.3 * 0.05 * torch.randn(B, C, T, generator=g)
        if i % 2 == 1:  # also an unrelated-estimate case (small margins between permutations)
            est = torch.randn(B, C, T, generator=g) * 0.05
        lens_t = torch.LongTensor(lens)
        for b, n in enumerate(lens):
            src[b, :, n:] = 0
        est_in = est.clone().requires_grad_(True)
        est_work = est_in * 1.0  # non-leaf so the in-place mask is allowed
        loss, max_snr, est_masked, reord = cal_loss(src, est_work, lens_t)
        (grad,) = torch.autograd.grad(loss, est_in)
        with torch.no_grad():
            ms2, perms, idx = cal_si_snr_with_pit(src, est.clone(), lens_t)
        arrays.update({f"c{i}_source": np_(src), f"c{i}_est": np_(est), f"c{i}_lengths": np_(lens_t),
                       f"c{i}_loss": np_(loss), f"c{i}_max_snr": np_(max_snr), f"c{i}_est_masked": np_(est_masked),
                       f"c{i}_reorder": np_(reord), f"c{i}_perms": np_(perms), f"c{i}_idx": np_(idx),
                       f"c{i}_grad_est": np_(grad)})
    arrays["n_cases"] = np.int64(len(cases))
    save("pit.npz", **arrays)


# ---------------------------------------------------------------- small models, fwd + loss + grads
SMALL = {
    "gln": dict(N=16, L=8, B=8, H=16, P=3, X=3, R=2, C=2, norm_type="gLN", causal=False, mask_nonlinear="relu"),
    "cln_causal": dict(N=16, L=8, B=8, H=16, P=3, X=3, R=2, C=2, norm_type="cLN", causal=True, mask_nonlinear="relu"),
    "softmax_c3": dict(N=12, L=6, B=8, H=20, P=3, X=2, R=2, C=3, norm_type="gLN", causal=False, mask_nonlinear="softmax"),
    "gln_causal_p2": dict(N=16, L=10, B=12, H=16, P=2, X=3, R=1, C=2, norm_type="gLN", causal=True, mask_nonlinear="relu"),
    "cln_p5": dict(N=8, L=4, B=8, H=8, P=5, X=2, R=1, C=2, norm_type="cLN", causal=False, mask_nonlinear="relu"),
}


def synthetic(M, T, C, L, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(M, C, T, generator=g) * 0.05
    lengths = torch.full((M,), T, dtype=torch.long)
    short = T - 3 * (L // 2) - 7
    lengths[-1] = short
    src[-1, :, short:] = 0
    mix = src.sum(1).clamp_(-0.9, 0.9)
    return mix, src, lengths


def run_model(cfg, M, T, seed_w, seed_x, perturb_scalars=True):
    torch.manual_seed(seed_w)
    model = ConvTasNet(**cfg)
    if perturb_scalars:  # make PReLU slopes distinct so a kernel cannot get away with 0.25 everywhere
        g = torch.Generator().manual_seed(99)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if p.dim() == 1:
                    p.copy_(0.05 + 0.4 * torch.rand(1, generator=g))
    mix, src, lengths = synthetic(M, T, cfg["C"], cfg["L"], seed_x)
    est = model(mix)
    est_raw = est.detach().clone()
    loss, max_snr, est_masked, reord = cal_loss(src, est, lengths)
    model.zero_grad()
    loss.backward()
    return model, mix, src, lengths, est_raw, loss, max_snr, est_masked, reord


def golden_small_models():
    for name, cfg in SMALL.items():
        T = 403 if name != "cln_p5" else 131
        model, mix, src, lengths, est_raw, loss, max_snr, est_masked, reord = run_model(cfg, 3, T, 5, 17)
        arrays = {"cfg_" + k: np.array(v) for k, v in cfg.items()}
        arrays.update(mixture=np_(mix), source=np_(src), lengths=np_(lengths), est_source=np_(est_raw),
                      loss=np_(loss), max_snr=np_(max_snr), est_masked=np_(est_masked), reorder=np_(reord))
        for k, v in model.state_dict().items():
            arrays["w:" + k] = np_(v)
        for k, p in model.named_parameters():
            arrays["g:" + k] = np_(p.grad)
        save(f"model_{name}.npz", **arrays)


# ---------------------------------------------------------------- paper config: seeded init + outputs
def golden_paper():
    cfg = dict(N=256, L=20, B=256, H=512, P=3, X=8, R=4, C=2, norm_type="gLN", causal=False, mask_nonlinear="relu")
    model, mix, src, lengths, est_raw, loss, max_snr, est_masked, reord = run_model(
        cfg, 1, 32000, 0, 1234 + 1, perturb_scalars=False)
    arrays = {"cfg_" + k: np.array(v) for k, v in cfg.items()}
    names, wsum, wabs, wfirst, gnorm = [], [], [], [], []
    for k, p in model.named_parameters():
        names.append(k)
        wsum.append(p.detach().double().sum().item())
        wabs.append(p.detach().double().abs().sum().item())
        wfirst.append(p.detach().flatten()[0].item())
        gnorm.append(p.grad.double().norm().item())
    arrays.update(names=np.array(names), w_sum=np.array(wsum), w_abs=np.array(wabs),
                  w_first=np.array(wfirst, dtype=np.float32), g_norm=np.array(gnorm),
                  seed_w=np.int64(0), seed_x=np.int64(1235), M=np.int64(1), T=np.int64(32000),
                  est_stride=np.int64(37), est_sub=np_(est_raw)[..., ::37], est_abs_max=np.float64(est_raw.abs().max().item()),
                  loss=np_(loss), max_snr=np_(max_snr),
                  g_enc=np_(model.encoder.conv1d_U.weight.grad), g_dec=np_(model.decoder.basis_signals.weight.grad))
    save("paper_cfg1.npz", **arrays)


if __name__ == "__main__":
    golden_ola()
    golden_pit()
    golden_small_models()
    golden_paper()
