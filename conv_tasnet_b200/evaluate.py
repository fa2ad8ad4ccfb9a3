"""Batched, device-side SI-SNR improvement — the metric of src/evaluate.py:94-130 without the per-utterance
device->host copies and numpy loops.  Works on the padded batch the evaluation loop already has
(`padded_source`, `reorder_estimate_source` from cal_loss, `padded_mixture`, `mixture_lengths`); positions past each
utterance's length are ignored, exactly like the reference's remove_pad + per-utterance call."""
import torch

EPS = 1e-8


def _masked_sisnr(ref, out, mask, n):
    """SI-SNR over the last axis of zero-meaned signals (src/evaluate.py:114-130), fp64 reductions."""
    ref = ref.double() * mask
    out = out.double() * mask
    ref = (ref - ref.sum(-1, keepdim=True) / n) * mask
    out = (out - out.sum(-1, keepdim=True) / n) * mask
    ref_energy = (ref * ref).sum(-1, keepdim=True) + EPS
    proj = (ref * out).sum(-1, keepdim=True) * ref / ref_energy
    noise = out - proj
    ratio = (proj * proj).sum(-1) / ((noise * noise).sum(-1) + EPS)
    return 10.0 * torch.log(ratio + EPS) / torch.log(torch.tensor(10.0, dtype=torch.float64, device=ref.device))


def cal_SISNR(ref_sig, out_sig, eps=1e-8):
    """[T] (or [..., T]) tensors -> SI-SNR in dB; same formula as src/evaluate.py:114-130."""
    assert ref_sig.shape[-1] == out_sig.shape[-1]
    n = torch.tensor(float(ref_sig.shape[-1]), dtype=torch.float64, device=ref_sig.device)
    mask = torch.ones((), dtype=torch.float64, device=ref_sig.device)
    return _masked_sisnr(ref_sig, out_sig, mask, n)


def cal_SISNRi(src_ref, src_est, mix):
    """[C,T], [C,T] (reordered by the best PIT permutation), [T] -> average SI-SNR improvement over the C sources
    (src/evaluate.py:94-111 hard-codes C = 2; this is the same average for any C)."""
    est = cal_SISNR(src_ref, src_est)
    base = cal_SISNR(src_ref, mix.unsqueeze(0).expand_as(src_ref))
    return (est - base).mean()


def cal_SISNRi_batch(padded_source, reorder_estimate_source, padded_mixture, mixture_lengths):
    """[B,C,T], [B,C,T], [B,T], [B] -> [B] average SI-SNRi per utterance, computed on the inputs' device."""
    B, C, T = padded_source.shape
    dev = padded_source.device
    lengths = torch.as_tensor(mixture_lengths).to(dev)
    mask = (torch.arange(T, device=dev).view(1, 1, T) < lengths.view(B, 1, 1)).double()
    n = lengths.view(B, 1, 1).double()
    est = _masked_sisnr(padded_source, reorder_estimate_source, mask, n)
    base = _masked_sisnr(padded_source, padded_mixture.unsqueeze(1).expand(B, C, T), mask, n)
    return (est - base).mean(dim=1)
