"""Batched, device-side SI-SNR improvement — the metric of src/evaluate.py:94-130 without the per-utterance
device->host copies and numpy loops.  Works on the padded batch the evaluation loop already has
(`padded_source`, `reorder_estimate_source` from cal_loss, `padded_mixture`, `mixture_lengths`); positions past each
utterance's length are ignored, exactly like the reference's remove_pad + per-utterance call.

The arithmetic is ONE hand-written kernel (csrc/sisnri.cu, C ABI `ctn_sisnri`): a streaming pass over the three
signals accumulates eight masked moments per (utterance, speaker) in fp64 and the last block of an utterance turns
them into SI-SNR(ref, est) - SI-SNR(ref, mix).  CUDA tensors only, like the rest of the package (no CPU fallback)."""
import torch

from . import _lib


def _run(src, est, mix, lengths, want_sisnr):
    src = src.detach().to(torch.float32).contiguous()
    est = est.detach().to(torch.float32).contiguous()
    mix = mix.detach().to(torch.float32).contiguous()
    B, C, T = src.shape
    if est.shape != src.shape or mix.shape != (B, T):
        raise ValueError("expected source / estimate [B,C,T] and mixture [B,T]")
    lengths = torch.as_tensor(lengths).to(device=src.device, dtype=torch.int64).contiguous()
    L = _lib.lib()
    out = torch.empty(B, dtype=torch.float32, device=src.device)
    each = torch.empty(B, C, dtype=torch.float32, device=src.device) if want_sisnr else None
    with torch.cuda.device(src.device):
        ws = torch.empty(L.ctn_sisnri_workspace_bytes(B, C), dtype=torch.uint8, device=src.device)
        _lib.check(L.ctn_sisnri(_lib.ptr(src), _lib.ptr(est), _lib.ptr(mix), _lib.ptr(lengths), B, C, T, _lib.ptr(out),
                                _lib.ptr(each), _lib.ptr(ws), _lib.stream()))
    return out, each


def cal_SISNRi_batch(padded_source, reorder_estimate_source, padded_mixture, mixture_lengths):
    """[B,C,T], [B,C,T], [B,T], [B] -> [B] average SI-SNRi per utterance (float32, on the inputs' device)."""
    return _run(padded_source, reorder_estimate_source, padded_mixture, mixture_lengths, False)[0]


def cal_SISNRi(src_ref, src_est, mix):
    """[C,T], [C,T] (reordered by the best PIT permutation), [T] -> average SI-SNR improvement over the C sources
    (src/evaluate.py:94-111 hard-codes C = 2; this is the same average for any C).  Returns a 0-dim tensor."""
    lengths = torch.full((1,), src_ref.shape[-1], dtype=torch.int64, device=src_ref.device)
    return cal_SISNRi_batch(src_ref.unsqueeze(0), src_est.unsqueeze(0), mix.unsqueeze(0), lengths)[0]


def cal_SISNR(ref_sig, out_sig, eps=1e-8):
    """[T] tensors -> SI-SNR in dB, the formula of src/evaluate.py:114-130 (eps fixed at the reference's 1e-8)."""
    if eps != 1e-8:
        raise ValueError("the kernel implements the reference's eps = 1e-8")
    assert ref_sig.shape[-1] == out_sig.shape[-1]
    T = ref_sig.shape[-1]
    lengths = torch.full((1,), T, dtype=torch.int64, device=ref_sig.device)
    _, each = _run(ref_sig.reshape(1, 1, T), out_sig.reshape(1, 1, T), out_sig.reshape(1, T), lengths, True)
    return each[0, 0]
