"""Drop-in for src/pit_criterion.py on B200: utterance-level PIT SI-SNR as one fused moments kernel + argmax,
an elementwise backward, and a gather for reorder_source (C ABI: ctn_pit_forward / ctn_pit_backward /
ctn_reorder_source).  Same signatures, return values and quirks as the reference:
  * `estimate_source` is masked IN PLACE and returned (src/pit_criterion.py:38,24)
  * the target mean uses the un-masked sum (:42), max_snr is [B,1] (:75), reorder applies the permutation
    itself, not its inverse (:92-98), permutations are lexicographic (:67), argmax keeps the first maximum (:73)
  * fp32 only (the reference hard-codes .float(), :72)
"""
from itertools import permutations

import torch

from . import _lib

EPS = 1e-8


def _check(source, estimate_source):
    assert source.size() == estimate_source.size()  # src/pit_criterion.py:34
    if source.dim() != 3:
        raise ValueError("source / estimate_source must be [B, C, T]")
    if not (source.is_cuda and estimate_source.is_cuda):
        raise RuntimeError("conv_tasnet_b200.pit_criterion runs on CUDA tensors only (no CPU fallback)")
    if source.dtype != torch.float32 or estimate_source.dtype != torch.float32:
        raise TypeError("PIT SI-SNR is fp32 only, like the reference (src/pit_criterion.py:72)")
    if not estimate_source.is_contiguous():
        raise RuntimeError("estimate_source must be contiguous (it is masked in place)")


def _lengths(source_lengths, device):
    return torch.as_tensor(source_lengths).to(device=device, dtype=torch.int64).contiguous()


def _pit_forward_raw(source, estimate_source, lengths, want_reorder):
    B, C, T = source.shape
    dev = source.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        max_snr = torch.empty(B, 1, dtype=torch.float32, device=dev)
        idx = torch.empty(B, dtype=torch.int64, device=dev)
        coef = torch.empty(B, C, 4, dtype=torch.float32, device=dev)
        reorder = torch.empty_like(estimate_source) if want_reorder else None
        ws = torch.empty(L.ctn_pit_workspace_bytes(B, C), dtype=torch.uint8, device=dev)
        _lib.check(L.ctn_pit_forward(_lib.ptr(source), _lib.ptr(estimate_source), _lib.ptr(lengths), B, C, T,
                                     _lib.ptr(loss), _lib.ptr(max_snr), _lib.ptr(idx), _lib.ptr(reorder),
                                     _lib.ptr(coef), _lib.ptr(ws), _lib.stream()))
    return loss, max_snr, idx, coef, reorder


class _PitLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, source, estimate_source, lengths):
        source = source.contiguous()
        loss, max_snr, idx, coef, reorder = _pit_forward_raw(source, estimate_source, lengths, True)
        ctx.mark_dirty(estimate_source)
        ctx.mark_non_differentiable(max_snr, idx, reorder)
        ctx.save_for_backward(source, estimate_source, lengths, coef)
        return loss.view(()), max_snr, idx, reorder, estimate_source

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss, _g_snr, _g_idx, _g_reorder, g_est):
        source, est_masked, lengths, coef = ctx.saved_tensors
        B, C, T = source.shape
        with torch.cuda.device(source.device):
            d_est = torch.empty_like(est_masked)
            g = g_loss.to(torch.float32).contiguous().view(1)
            _lib.check(_lib.lib().ctn_pit_backward(_lib.ptr(source), _lib.ptr(est_masked), _lib.ptr(lengths),
                                                   _lib.ptr(coef), _lib.ptr(g), B, C, T, _lib.ptr(d_est),
                                                   _lib.stream()))
            if g_est is not None:  # someone also back-propagated through the returned (masked) estimate
                d_est += g_est * get_mask(source, lengths)
        return None, d_est, None


def cal_loss(source, estimate_source, source_lengths):
    """
    Args:
        source: [B, C, T], B is batch size
        estimate_source: [B, C, T]
        source_lengths: [B]
    Returns (loss, max_snr [B,1], estimate_source (masked in place), reorder_estimate_source)
    (src/pit_criterion.py:12-24)
    """
    _check(source, estimate_source)
    lengths = _lengths(source_lengths, source.device)
    loss, max_snr, _idx, reorder, est = _PitLossFn.apply(source, estimate_source, lengths)
    return loss, max_snr, est, reorder


def cal_si_snr_with_pit(source, estimate_source, source_lengths):
    """-> (max_snr [B,1], perms [C!,C], max_snr_idx [B]); masks estimate_source in place (src/pit_criterion.py:27-77).
    Not differentiable through this entry point (use cal_loss for training)."""
    _check(source, estimate_source)
    lengths = _lengths(source_lengths, source.device)
    with torch.no_grad():
        _, max_snr, idx, _, _ = _pit_forward_raw(source.contiguous(), estimate_source, lengths, False)
    C = source.size(1)
    perms = source.new_tensor(list(permutations(range(C))), dtype=torch.long)
    return max_snr, perms, idx


def reorder_source(source, perms, max_snr_idx):
    """out[b, c] = source[b, perms[max_snr_idx[b]][c]]  (src/pit_criterion.py:80-99).
    `perms` must be the lexicographic table cal_si_snr_with_pit returns."""
    if not source.is_cuda:
        raise RuntimeError("conv_tasnet_b200.pit_criterion runs on CUDA tensors only (no CPU fallback)")
    B, C = source.shape[:2]
    if perms.shape != (torch.tensor(list(permutations(range(C)))).shape):
        raise ValueError("perms must be the [C!, C] lexicographic permutation table")
    src = source.contiguous().float()
    inner = src[0, 0].numel()
    out = torch.empty_like(src)
    idx = max_snr_idx.to(device=source.device, dtype=torch.int64).contiguous()
    with torch.cuda.device(source.device):
        _lib.check(_lib.lib().ctn_reorder_source(_lib.ptr(src), _lib.ptr(idx), B, C, inner, _lib.ptr(out),
                                                 _lib.stream()))
    return out


def get_mask(source, source_lengths):
    """[B,1,T] ones with zeros from source_lengths[b] on (src/pit_criterion.py:102-114), built on the device
    without the reference's per-item host sync."""
    B, _, T = source.size()
    lengths = _lengths(source_lengths, source.device)
    t = torch.arange(T, device=source.device).view(1, 1, T)
    return (t < lengths.view(B, 1, 1)).to(source.dtype)
