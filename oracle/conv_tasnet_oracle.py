"""CPU oracle for the Conv-TasNet hot path (TEST INFRASTRUCTURE, not product).

This file is a functional restatement, in plain PyTorch CPU ops, of the algorithm of
the reference's model + loss path.  It exists only to check the CUDA path; nothing in
`conv_tasnet_b200/` imports it.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may call it.

Parity pinning: every function below is checked in `tests/test_oracle_golden.py`
against golden vectors produced by *running the reference itself*
(`/root/reference/src/{conv_tasnet,pit_criterion,utils}.py`) in the build container;
the generating script is `tests/golden/make_golden.py`, the vectors are committed
under `tests/golden/*.npz`.  The reference's own two reproducible known-answers
(`src/utils.py:70-77` overlap-add print-out, `src/pit_criterion.py:117-133` seed-123
loss 45.9221) are part of those vectors.

The arithmetic itself lives in torch (third party, not vendored by the reference, which
pins only "PyTorch 0.4.1+", README.md:17; this image has torch 2.11.0).  The oracle works
in any floating dtype (fp32 to mirror the reference, fp64 to measure the reference's own
rounding noise), on a `state_dict` with exactly the reference's key names and shapes.

Citations are `file:line` into /root/reference.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-8  # src/conv_tasnet.py:10, src/pit_criterion.py:9


@dataclass(frozen=True)
class Config:
    """Hyper-parameters, same names/meaning as ConvTasNet.__init__ (src/conv_tasnet.py:14-30)."""
    N: int = 256
    L: int = 20
    B: int = 256
    H: int = 512
    P: int = 3
    X: int = 8
    R: int = 4
    C: int = 2
    norm_type: str = "gLN"
    causal: bool = False
    mask_nonlinear: str = "relu"

    def as_dict(self):
        return asdict(self)


PAPER = Config()


# --------------------------------------------------------------------------------------
# parameter inventory (state_dict contract; src/conv_tasnet.py:150-278)
# --------------------------------------------------------------------------------------
def is_batch_norm(cfg: Config) -> bool:
    """chose_norm (src/conv_tasnet.py:298-309): every norm_type other than gLN / cLN builds nn.BatchNorm1d."""
    return cfg.norm_type not in ("gLN", "cLN")


def param_spec(cfg: Config) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) in the order `ConvTasNet(...).state_dict()` yields them."""
    spec = [("encoder.conv1d_U.weight", (cfg.N, 1, cfg.L)),
            ("separator.network.0.gamma", (1, cfg.N, 1)),
            ("separator.network.0.beta", (1, cfg.N, 1)),
            ("separator.network.1.weight", (cfg.B, cfg.N, 1))]
    # Chomp1d sits at index 1 of DepthwiseSeparableConv.net in the causal variant
    # (src/conv_tasnet.py:264-269), shifting the later indices by one.
    sh = 1 if cfg.causal else 0

    def norm_entries(prefix):
        if is_batch_norm(cfg):  # nn.BatchNorm1d(H) parameters and buffers, in state_dict order (:306-309)
            return [(prefix + "weight", (cfg.H,)), (prefix + "bias", (cfg.H,)), (prefix + "running_mean", (cfg.H,)),
                    (prefix + "running_var", (cfg.H,)), (prefix + "num_batches_tracked", ())]
        return [(prefix + "gamma", (1, cfg.H, 1)), (prefix + "beta", (1, cfg.H, 1))]

    for r in range(cfg.R):
        for x in range(cfg.X):
            p = f"separator.network.2.{r}.{x}.net."
            spec += [(p + "0.weight", (cfg.H, cfg.B, 1)),
                     (p + "1.weight", (1,))]
            spec += norm_entries(p + "2.")
            spec += [(p + "3.net.0.weight", (cfg.H, 1, cfg.P)),
                     (p + f"3.net.{1 + sh}.weight", (1,))]
            spec += norm_entries(p + f"3.net.{2 + sh}.")
            spec += [(p + f"3.net.{3 + sh}.weight", (cfg.B, cfg.H, 1))]
    spec += [("separator.network.3.weight", (cfg.C * cfg.N, cfg.B, 1)),
             ("decoder.basis_signals.weight", (cfg.L, cfg.N))]
    return spec


def init_state_dict(cfg: Config, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random weights with the reference's *distribution* (xavier-normal on every >1-D
    tensor including gamma/beta, PReLU slope 0.25; src/conv_tasnet.py:41-43).  Not the
    reference's RNG stream: seeded identical streams are the product's job (tested there)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_spec(cfg):
        leaf = name.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            sd[name] = torch.zeros((), dtype=torch.int64)
        elif leaf in ("running_mean", "bias") and len(shape) == 1:
            sd[name] = torch.zeros(shape, dtype=dtype)
        elif leaf == "running_var" or (leaf == "weight" and len(shape) == 1 and shape[0] != 1):
            sd[name] = torch.ones(shape, dtype=dtype)  # BatchNorm1d defaults (1-D: not touched by the xavier loop, :41-43)
        elif len(shape) == 1:
            sd[name] = torch.full(shape, 0.25, dtype=dtype)
        else:
            rf = 1
            for s in shape[2:]:
                rf *= s
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
            std = math.sqrt(2.0 / (fan_in + fan_out))
            sd[name] = (torch.randn(shape, generator=g, dtype=torch.float64) * std).to(dtype)
    return sd


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def n_frames(T: int, L: int) -> int:
    """K of src/conv_tasnet.py:113 for a strided conv without padding."""
    S = L // 2
    return (T - L) // S + 1


def encoder(mixture: torch.Tensor, U: torch.Tensor) -> torch.Tensor:
    """[M,T] -> [M,N,K]: relu(conv1d(x, U, stride=L//2)) (src/conv_tasnet.py:106,119-120)."""
    L = U.shape[-1]
    return F.relu(F.conv1d(mixture.unsqueeze(1), U, stride=L // 2))


def channelwise_layer_norm(y, gamma, beta):
    """cLN: statistics over channels for every (m,k); biased variance (src/conv_tasnet.py:332-334)."""
    mean = y.mean(dim=1, keepdim=True)
    var = y.var(dim=1, keepdim=True, unbiased=False)
    return gamma * (y - mean) / torch.pow(var + EPS, 0.5) + beta


def global_layer_norm(y, gamma, beta):
    """gLN: statistics over (channels, time) per sample, two-pass variance (src/conv_tasnet.py:358-360)."""
    mean = y.mean(dim=(1, 2), keepdim=True)
    var = ((y - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    return gamma * (y - mean) / torch.pow(var + EPS, 0.5) + beta


def batch_norm(y, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    """nn.BatchNorm1d(C) on [M, C, K] (src/conv_tasnet.py:306-309; torch defaults eps 1e-5, momentum 0.1, affine):
    training: statistics over (M, K) per channel, biased variance to normalise, running statistics updated in place
    with the unbiased one; evaluation: the running statistics."""
    if training:
        mean = y.mean(dim=(0, 2))
        var = y.var(dim=(0, 2), unbiased=False)
        n = y.shape[0] * y.shape[2]
        with torch.no_grad():
            running_mean.mul_(1 - momentum).add_(momentum * mean.to(running_mean.dtype))
            running_var.mul_(1 - momentum).add_(momentum * (var * (n / max(n - 1, 1))).to(running_var.dtype))
    else:
        mean, var = running_mean.to(y.dtype), running_var.to(y.dtype)
    xhat = (y - mean.view(1, -1, 1)) / torch.sqrt(var.view(1, -1, 1) + eps)
    return xhat * weight.view(1, -1, 1) + bias.view(1, -1, 1)


def _norm(cfg: Config, sd, prefix: str, y, training: bool):
    """the norm chose_norm built at `prefix` (src/conv_tasnet.py:298-309)"""
    if cfg.norm_type == "gLN":
        return global_layer_norm(y, sd[prefix + "gamma"], sd[prefix + "beta"])
    if cfg.norm_type == "cLN":
        return channelwise_layer_norm(y, sd[prefix + "gamma"], sd[prefix + "beta"])
    if training:
        sd[prefix + "num_batches_tracked"] += 1
    return batch_norm(y, sd[prefix + "weight"], sd[prefix + "bias"], sd[prefix + "running_mean"],
                      sd[prefix + "running_var"], training)


def temporal_block(cfg: Config, sd, prefix: str, x: torch.Tensor, dilation: int, training: bool = True) -> torch.Tensor:
    """x + pointwise(norm(prelu(depthwise(norm(prelu(conv1x1(x))))))) (src/conv_tasnet.py:223-243,253-278)."""
    sh = 1 if cfg.causal else 0
    pad = (cfg.P - 1) * dilation if cfg.causal else (cfg.P - 1) * dilation // 2  # :182
    y = F.conv1d(x, sd[prefix + "0.weight"])
    y = F.prelu(y, sd[prefix + "1.weight"])
    y = _norm(cfg, sd, prefix + "2.", y, training)
    y = F.conv1d(y, sd[prefix + "3.net.0.weight"], padding=pad, dilation=dilation, groups=cfg.H)
    if cfg.causal:
        y = y[:, :, :-pad].contiguous()  # Chomp1d, :295
    y = F.prelu(y, sd[prefix + f"3.net.{1 + sh}.weight"])
    y = _norm(cfg, sd, prefix + f"3.net.{2 + sh}.", y, training)
    y = F.conv1d(y, sd[prefix + f"3.net.{3 + sh}.weight"])
    return y + x  # no output ReLU (:243)


def separator(cfg: Config, sd, mixture_w: torch.Tensor, training: bool = True) -> torch.Tensor:
    """[M,N,K] -> mask [M,C,N,K] (src/conv_tasnet.py:172-214).  First norm is always cLN (:172)."""
    M, N, K = mixture_w.shape
    y = channelwise_layer_norm(mixture_w, sd["separator.network.0.gamma"], sd["separator.network.0.beta"])
    y = F.conv1d(y, sd["separator.network.1.weight"])
    for r in range(cfg.R):
        for x in range(cfg.X):
            y = temporal_block(cfg, sd, f"separator.network.2.{r}.{x}.net.", y, 2 ** x, training)  # :181
    score = F.conv1d(y, sd["separator.network.3.weight"]).view(M, cfg.C, N, K)  # channel = c*N + n (:208)
    if cfg.mask_nonlinear == "softmax":
        return F.softmax(score, dim=1)
    if cfg.mask_nonlinear == "relu":
        return F.relu(score)
    raise ValueError("Unsupported mask non-linear function")  # :213-214


def overlap_and_add(signal: torch.Tensor, frame_step: int) -> torch.Tensor:
    """[..., frames, frame_length] -> [..., (frames-1)*step + frame_length] (src/utils.py:9-47).

    Restated as a direct sum: out[..., k*step + l] += signal[..., k, l], accumulating frames in
    ascending k (the order index_add_ visits them on CPU, src/utils.py:45)."""
    *outer, frames, frame_length = signal.shape
    out = signal.new_zeros(*outer, (frames - 1) * frame_step + frame_length)
    for k in range(frames):
        out[..., k * frame_step:k * frame_step + frame_length] += signal[..., k, :]
    return out


def overlap_and_add_fast(signal: torch.Tensor, frame_step: int) -> torch.Tensor:
    """Same result as `overlap_and_add` for frame_length == 2*frame_step (two contributors per
    sample => order independent), vectorised so the CPU baseline is not dominated by a Python loop."""
    *outer, frames, frame_length = signal.shape
    if frame_length != 2 * frame_step:
        return overlap_and_add(signal, frame_step)
    out = signal.new_zeros(*outer, frames + 1, frame_step)
    out[..., :frames, :] += signal[..., :frame_step]
    out[..., 1:, :] += signal[..., frame_step:]
    return out.reshape(*outer, -1)


def decoder(mixture_w, est_mask, V, L: int):
    """mask*w -> basis -> overlap-add (src/conv_tasnet.py:140-145).  V is Linear(N,L).weight [L,N]."""
    source_w = (mixture_w.unsqueeze(1) * est_mask).transpose(2, 3)  # [M,C,K,N]
    frames = source_w @ V.t()  # [M,C,K,L]
    return overlap_and_add_fast(frames, L // 2)


def forward(cfg: Config, sd, mixture: torch.Tensor, training: bool = True) -> torch.Tensor:
    """ConvTasNet.forward (src/conv_tasnet.py:45-60): [M,T] -> [M,C,T], right-padded with zeros.
    `training` is nn.Module.training (default True, like a freshly built module); it only matters for the BatchNorm
    branch, whose running statistics inside `sd` are then updated in place."""
    w = encoder(mixture, sd["encoder.conv1d_U.weight"])
    mask = separator(cfg, sd, w, training)
    est = decoder(w, mask, sd["decoder.basis_signals.weight"], cfg.L)
    return F.pad(est, (0, mixture.shape[-1] - est.shape[-1]))


# --------------------------------------------------------------------------------------
# PIT SI-SNR (src/pit_criterion.py)
# --------------------------------------------------------------------------------------
def get_mask(source, source_lengths):
    """[B,1,T] of ones with zeros from source_lengths[b] on (src/pit_criterion.py:102-114)."""
    T = source.shape[-1]
    t = torch.arange(T, device=source.device).view(1, 1, T)
    return (t < source_lengths.view(-1, 1, 1).to(t.device)).to(source.dtype)


def permutations_of(C: int) -> torch.Tensor:
    """[C!, C] int64, lexicographic like itertools.permutations(range(C)) (src/pit_criterion.py:67)."""
    return torch.tensor(list(itertools.permutations(range(C))), dtype=torch.long)


def cal_si_snr_with_pit(source, estimate_source, source_lengths):
    """(max_snr [B,1], perms [C!,C], max_snr_idx [B]); masks `estimate_source` IN PLACE
    (src/pit_criterion.py:27-77)."""
    assert source.size() == estimate_source.size()
    B, C, T = source.shape
    mask = get_mask(source, source_lengths)
    estimate_source *= mask  # :38 in place on the caller's tensor
    n = source_lengths.view(-1, 1, 1).to(source.dtype)
    tgt = (source - source.sum(dim=2, keepdim=True) / n) * mask  # target mean uses the un-masked sum (:42)
    est = (estimate_source - estimate_source.sum(dim=2, keepdim=True) / n) * mask
    s_t = tgt.unsqueeze(1)  # [B,1,C,T]
    s_e = est.unsqueeze(2)  # [B,C,1,T]
    dot = (s_e * s_t).sum(dim=3, keepdim=True)
    energy = (s_t ** 2).sum(dim=3, keepdim=True) + EPS
    proj = dot * s_t / energy
    noise = s_e - proj
    ratio = (proj ** 2).sum(dim=3) / ((noise ** 2).sum(dim=3) + EPS)
    snr = 10 * torch.log10(ratio + EPS)  # [B, C_est, C_tgt]
    perms = permutations_of(C).to(source.device)
    # snr_set[b,p] = sum_i snr[b, i, perms[p][i]]  (one-hot einsum at :69-72)
    rows = torch.arange(C, device=source.device)
    snr_set = snr[:, rows.unsqueeze(0), perms].sum(dim=2)  # [B, C!]
    max_snr_idx = torch.argmax(snr_set, dim=1)
    max_snr = snr_set.max(dim=1, keepdim=True).values / C
    return max_snr, perms, max_snr_idx


def reorder_source(source, perms, max_snr_idx):
    """out[b,c] = source[b, perms[idx[b]][c]] — the permutation, not its inverse
    (src/pit_criterion.py:80-99)."""
    sel = perms[max_snr_idx]  # [B,C]
    return torch.gather(source, 1, sel.view(*sel.shape, *([1] * (source.dim() - 2))).expand_as(source))


def cal_loss(source, estimate_source, source_lengths):
    """(loss, max_snr, estimate_source (masked in place), reordered) (src/pit_criterion.py:12-24)."""
    max_snr, perms, idx = cal_si_snr_with_pit(source, estimate_source, source_lengths)
    loss = 0 - torch.mean(max_snr)
    return loss, max_snr, estimate_source, reorder_source(estimate_source, perms, idx)


# --------------------------------------------------------------------------------------
# whole step (what the CPU baseline times)
# --------------------------------------------------------------------------------------
def synthetic_batch(M: int, T: int, C: int, L: int, seed: int, dtype=torch.float32):
    """SURVEY §8d synthetic inputs: sources ~ N(0, 0.05^2), mixture = clamp(sum, +-0.9), the last
    item is 3*S+7 samples short with a zeroed tail."""
    g = torch.Generator().manual_seed(seed)
    src = (torch.randn(M, C, T, generator=g, dtype=torch.float32) * 0.05).to(dtype)
    lengths = torch.full((M,), T, dtype=torch.long)
    short = T - 3 * (L // 2) - 7
    if short > 0:
        lengths[-1] = short
        src[-1, :, short:] = 0
    mix = src.sum(1).clamp_(-0.9, 0.9)
    return mix, src, lengths


def is_buffer(name: str) -> bool:
    """BatchNorm1d buffers inside a state_dict (no gradient; the running statistics are updated in place)"""
    return name.rsplit(".", 1)[-1] in ("running_mean", "running_var", "num_batches_tracked")


def train_step_grads(cfg: Config, sd, mixture, source, lengths, training: bool = True):
    """forward + cal_loss + backward through autograd; returns (loss, est (masked), grads by name).
    BatchNorm buffers in `sd` are used (and, when `training`, updated) in place, like the module's."""
    params = {k: (v if is_buffer(k) else v.detach().clone().requires_grad_(True)) for k, v in sd.items()}
    est = forward(cfg, params, mixture, training)
    loss, max_snr, est_masked, reordered = cal_loss(source, est, lengths)
    names = [k for k in params if not is_buffer(k)]
    grads = torch.autograd.grad(loss, [params[k] for k in names])
    return loss.detach(), est_masked.detach(), dict(zip(names, grads)), max_snr.detach(), reordered.detach()


# --------------------------------------------------------------------------------------
# evaluation metric (src/evaluate.py:94-130), numpy like the reference
# --------------------------------------------------------------------------------------
def cal_SISNR_np(ref_sig, out_sig, eps=1e-8):
    """src/evaluate.py:114-130"""
    import numpy as np
    ref_sig = ref_sig - np.mean(ref_sig)
    out_sig = out_sig - np.mean(out_sig)
    ref_energy = np.sum(ref_sig ** 2) + eps
    proj = np.sum(ref_sig * out_sig) * ref_sig / ref_energy
    noise = out_sig - proj
    ratio = np.sum(proj ** 2) / (np.sum(noise ** 2) + eps)
    return 10 * np.log(ratio + eps) / np.log(10.0)


def cal_SISNRi_np(src_ref, src_est, mix):
    """src/evaluate.py:94-111 (two sources)"""
    s1, s2 = cal_SISNR_np(src_ref[0], src_est[0]), cal_SISNR_np(src_ref[1], src_est[1])
    b1, b2 = cal_SISNR_np(src_ref[0], mix), cal_SISNR_np(src_ref[1], mix)
    return ((s1 - b1) + (s2 - b2)) / 2
