"""torch.library registration of the hot path: the whole-model forward / backward and the PIT SI-SNR loss as dispatcher
ops (`torch.ops.ctn_b200.*`), so that dispatcher-level users (torch.compile graphs, torch.export, functorch-style
callers) see them as opaque CUDA ops with shape inference (`register_fake`) and a backward formula
(`register_autograd`).  Each op is a thin call into the C ABI of include/ctn_b200.h — the same entry points the
`ConvTasNet` module and `cal_loss` use; there is no CPU implementation (device_types="cuda" only).

    est, workspace = torch.ops.ctn_b200.model_forward(flat_params, mixture, cfg, training)
    grad = torch.ops.ctn_b200.model_backward(flat_params, mixture, d_est, workspace, cfg)       # flat gradient
    loss, max_snr, idx, reorder, coef = torch.ops.ctn_b200.pit_forward(source, est, lengths)     # masks est in place
    d_est = torch.ops.ctn_b200.pit_backward(source, est_masked, lengths, coef, grad_loss)

`cfg` is the list of ints [N, L, B, H, P, X, R, C, norm (0 gLN / 1 cLN), causal, mask (0 relu / 1 softmax)];
`flat_params` is the model's flat parameter buffer (`ConvTasNet.flat_params`, reference state_dict order) and
`workspace` the uint8 buffer of `workspace_bytes(cfg, M, T, training)` bytes the forward allocates and returns: it holds
the activation stash the backward consumes (functional ops: autograd formulas cannot be registered on mutating ones; the
backward's scratch region inside the workspace is private to the op).  Functional wrappers: `separate(...)` (differentiable w.r.t. flat_params) and `pit_loss`.
Replaces: ConvTasNet.forward src/conv_tasnet.py:45-60 (+ its autograd), cal_loss src/pit_criterion.py:12-24."""
import ctypes
from typing import List, Tuple

import torch

from . import _lib

_NS = "ctn_b200"


def _cfg_struct(cfg: List[int]):
    if len(cfg) != 11:
        raise ValueError("cfg must be [N, L, B, H, P, X, R, C, norm, causal, mask]")
    if cfg[8] not in (0, 1):
        raise ValueError("the dispatcher ops cover norm_type gLN (0) and cLN (1); BatchNorm carries running statistics "
                         "and goes through the ConvTasNet module")
    return _lib.CtnConfig(*[int(v) for v in cfg])


def workspace_bytes(cfg: List[int], M: int, T: int, training: bool) -> int:
    n = _lib.lib().ctn_workspace_bytes(ctypes.byref(_cfg_struct(cfg)), M, T, 1 if training else 0)
    if n < 0:
        _lib.check(1)
    return n


def param_floats(cfg: List[int]) -> int:
    n = _lib.lib().ctn_param_floats(ctypes.byref(_cfg_struct(cfg)))
    if n < 0:
        _lib.check(1)
    return n


# ------------------------------------------------------------------------------------------- whole-model forward
@torch.library.custom_op(f"{_NS}::model_forward", mutates_args=(), device_types="cuda")
def model_forward(flat_params: torch.Tensor, mixture: torch.Tensor, cfg: List[int],
                  training: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    c = _cfg_struct(cfg)
    mixture = mixture.contiguous()
    M, T = mixture.shape
    est = torch.empty(M, cfg[7], T, dtype=torch.float32, device=mixture.device)
    workspace = torch.empty(workspace_bytes(cfg, M, T, training), dtype=torch.uint8, device=mixture.device)
    with torch.cuda.device(mixture.device):
        _lib.check(_lib.lib().ctn_model_forward(ctypes.byref(c), _lib.ptr(flat_params), _lib.ptr(mixture), M, T,
                                                _lib.ptr(est), _lib.ptr(workspace), workspace.numel(),
                                                1 if training else 0, _lib.stream()))
    return est, workspace


@model_forward.register_fake
def _(flat_params, mixture, cfg, training):
    M, T = mixture.shape
    if isinstance(M, int) and isinstance(T, int):
        nbytes = workspace_bytes(list(cfg), M, T, bool(training))  # a pure host-side function of the shape
    else:  # symbolic shapes: opaque to the tracer
        nbytes = torch.library.get_ctx().new_dynamic_size()
    return mixture.new_empty(M, cfg[7], T, dtype=torch.float32), mixture.new_empty(nbytes, dtype=torch.uint8)


@torch.library.custom_op(f"{_NS}::model_backward", mutates_args=(), device_types="cuda")
def model_backward(flat_params: torch.Tensor, mixture: torch.Tensor, d_est: torch.Tensor, workspace: torch.Tensor,
                   cfg: List[int]) -> torch.Tensor:
    c = _cfg_struct(cfg)
    mixture, d_est = mixture.contiguous(), d_est.contiguous()
    M, T = mixture.shape
    grads = torch.empty_like(flat_params)
    with torch.cuda.device(mixture.device):
        _lib.check(_lib.lib().ctn_model_backward(ctypes.byref(c), _lib.ptr(flat_params), _lib.ptr(mixture), M, T,
                                                 _lib.ptr(d_est), _lib.ptr(grads), _lib.ptr(workspace),
                                                 workspace.numel(), 0, _lib.stream()))
    return grads


@model_backward.register_fake
def _(flat_params, mixture, d_est, workspace, cfg):
    return torch.empty_like(flat_params)


def _fwd_setup(ctx, inputs, output):
    flat_params, mixture, cfg, training = inputs
    _est, workspace = output
    ctx.training = training
    ctx.save_for_backward(flat_params, mixture, workspace)
    ctx.cfg = list(cfg)


def _fwd_backward(ctx, d_est, _d_ws):
    if not ctx.training:
        raise RuntimeError("ctn_b200::model_forward: differentiating needs training=True (the inference workspace keeps no "
                           "activation stash)")
    flat_params, mixture, workspace = ctx.saved_tensors
    return torch.ops.ctn_b200.model_backward(flat_params, mixture, d_est, workspace, ctx.cfg), None, None, None


model_forward.register_autograd(_fwd_backward, setup_context=_fwd_setup)


# ------------------------------------------------------------------------------------------- PIT SI-SNR
@torch.library.custom_op(f"{_NS}::pit_forward", mutates_args=("estimate_source",), device_types="cuda")
def pit_forward(source: torch.Tensor, estimate_source: torch.Tensor,
                lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    B, C, T = source.shape
    dev = source.device
    L = _lib.lib()
    source = source.contiguous()
    with torch.cuda.device(dev):
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        max_snr = torch.empty(B, 1, dtype=torch.float32, device=dev)
        idx = torch.empty(B, dtype=torch.int64, device=dev)
        coef = torch.empty(B, C, 4, dtype=torch.float32, device=dev)
        reorder = torch.empty_like(estimate_source)
        ws = torch.empty(L.ctn_pit_workspace_bytes(B, C), dtype=torch.uint8, device=dev)
        _lib.check(L.ctn_pit_forward(_lib.ptr(source), _lib.ptr(estimate_source), _lib.ptr(lengths), B, C, T,
                                     _lib.ptr(loss), _lib.ptr(max_snr), _lib.ptr(idx), _lib.ptr(reorder), _lib.ptr(coef),
                                     _lib.ptr(ws), _lib.stream()))
    return loss, max_snr, idx, reorder, coef


@pit_forward.register_fake
def _(source, estimate_source, lengths):
    B, C, T = source.shape
    f32 = dict(dtype=torch.float32)
    return (source.new_empty(1, **f32), source.new_empty(B, 1, **f32), source.new_empty(B, dtype=torch.int64),
            torch.empty_like(estimate_source), source.new_empty(B, C, 4, **f32))


@torch.library.custom_op(f"{_NS}::pit_backward", mutates_args=(), device_types="cuda")
def pit_backward(source: torch.Tensor, est_masked: torch.Tensor, lengths: torch.Tensor, coef: torch.Tensor,
                 grad_loss: torch.Tensor) -> torch.Tensor:
    B, C, T = source.shape
    source, est_masked = source.contiguous(), est_masked.contiguous()
    g = grad_loss.to(torch.float32).contiguous().view(1)
    with torch.cuda.device(source.device):
        d_est = torch.empty_like(est_masked)
        _lib.check(_lib.lib().ctn_pit_backward(_lib.ptr(source), _lib.ptr(est_masked), _lib.ptr(lengths), _lib.ptr(coef),
                                               _lib.ptr(g), B, C, T, _lib.ptr(d_est), _lib.stream()))
    return d_est


@pit_backward.register_fake
def _(source, est_masked, lengths, coef, grad_loss):
    return torch.empty_like(est_masked)


# ------------------------------------------------------------------------------------------- functional wrappers
class _PitLoss(torch.autograd.Function):
    """autograd glue over the two PIT ops (the forward masks its input in place, which `register_autograd` formulas
    may not depend on; an autograd.Function may)."""

    @staticmethod
    def forward(ctx, source, estimate_source, lengths):
        loss, max_snr, idx, reorder, coef = torch.ops.ctn_b200.pit_forward(source, estimate_source, lengths)
        ctx.mark_dirty(estimate_source)
        ctx.mark_non_differentiable(max_snr, idx, reorder)
        ctx.save_for_backward(source, estimate_source, lengths, coef)
        return loss.view(()), max_snr, idx, reorder, estimate_source

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss, _a, _b, _c, _d):
        source, est_masked, lengths, coef = ctx.saved_tensors
        return None, torch.ops.ctn_b200.pit_backward(source, est_masked, lengths, coef, g_loss), None


def separate(flat_params, mixture, cfg, training=None):
    """ConvTasNet.forward on the flat parameter buffer: mixture [M, T] -> est_source [M, C, T]; differentiable with
    respect to `flat_params` when training (default: torch.is_grad_enabled() and flat_params.requires_grad)."""
    if training is None:
        training = torch.is_grad_enabled() and flat_params.requires_grad
    est, _workspace = torch.ops.ctn_b200.model_forward(flat_params, mixture, list(cfg), bool(training))
    return est


def pit_loss(source, estimate_source, source_lengths):
    """cal_loss through the dispatcher ops: (loss, max_snr [B,1], estimate_source masked in place, reordered estimate)."""
    lengths = torch.as_tensor(source_lengths).to(device=source.device, dtype=torch.int64).contiguous()
    loss, max_snr, _idx, reorder, est = _PitLoss.apply(source, estimate_source, lengths)
    return loss, max_snr, est, reorder
