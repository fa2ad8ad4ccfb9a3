"""Drop-in for src/conv_tasnet.py::ConvTasNet on B200.

Same constructor, attributes, forward signature, state_dict keys/shapes, checkpoint package and seeded
initialisation as the reference (src/conv_tasnet.py:13-94); the math runs in hand-written sm_100a kernels
behind the C ABI (include/ctn_b200.h).  Parameters are views into ONE flat fp32 buffer (and gradients into
one flat gradient buffer) so the data-parallel all-reduce and the fused optimizer touch contiguous memory.
"""
import ctypes
import math
import weakref

import torch
import torch.nn as nn

from . import _lib

EPS = 1e-8


class _Weight(nn.Module):
    """Parameter holder that consumes the RNG exactly like the nn.Conv1d / nn.Linear the reference builds at this
    position (kaiming_uniform_(a=sqrt(5)), torch/nn/modules/conv.py reset_parameters), so that a seeded
    `ConvTasNet(...)` yields bit-identical weights to the reference's constructor."""

    def __init__(self, *shape):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class _PReLU(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.full((1,), 0.25))  # nn.PReLU() default (src/conv_tasnet.py:224,259)


class _Norm(nn.Module):
    """gamma/beta of ChannelwiseLayerNorm / GlobalLayerNorm (src/conv_tasnet.py:313-361)."""

    def __init__(self, channel_size):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(1, channel_size, 1))
        self.beta = nn.Parameter(torch.zeros(1, channel_size, 1))


class _Seq(nn.Sequential):
    def forward(self, *a, **k):  # parameter container only; compute happens in ConvTasNet.forward
        raise RuntimeError("sub-modules of the B200 ConvTasNet hold parameters only; call ConvTasNet.forward")


class _Slot(nn.Module):
    """Parameter-less place holder (keeps the reference's Sequential indices, e.g. Chomp1d at net.1)."""


class Encoder(nn.Module):
    def __init__(self, L, N):
        super().__init__()
        self.L, self.N = L, N
        self.conv1d_U = _Weight(N, 1, L)  # nn.Conv1d(1, N, L, stride=L//2, bias=False), src/conv_tasnet.py:106


class Decoder(nn.Module):
    def __init__(self, N, L):
        super().__init__()
        self.N, self.L = N, L
        self.basis_signals = _Weight(L, N)  # nn.Linear(N, L, bias=False), src/conv_tasnet.py:129


def _chose_norm(norm_type, channel_size):
    """Parameter holder at the position of chose_norm (src/conv_tasnet.py:298-309): gamma/beta [1,C,1] for gLN / cLN,
    and for anything else the reference's nn.BatchNorm1d itself (weight, bias, running_mean, running_var,
    num_batches_tracked: same state_dict keys, same defaults; its forward is never called)."""
    if norm_type in ("gLN", "cLN"):
        return _Norm(channel_size)
    return nn.BatchNorm1d(channel_size)


def _ds_conv(B, H, P, causal, norm_type):
    mods = [_Weight(H, 1, P)]  # depthwise, src/conv_tasnet.py:253
    if causal:
        mods.append(_Slot())  # Chomp1d
    mods += [_PReLU(), _chose_norm(norm_type, H), _Weight(B, H, 1)]
    net = _Seq(*mods)
    holder = nn.Module()
    holder.net = net
    return holder


def _temporal_block(B, H, P, causal, norm_type):
    holder = nn.Module()
    holder.net = _Seq(_Weight(H, B, 1), _PReLU(), _chose_norm(norm_type, H), _ds_conv(B, H, P, causal, norm_type))
    return holder


class TemporalConvNet(nn.Module):
    def __init__(self, N, B, H, P, X, R, C, causal, norm_type="gLN"):
        super().__init__()
        repeats = [_Seq(*[_temporal_block(B, H, P, causal, norm_type) for _ in range(X)]) for _ in range(R)]
        self.network = _Seq(_Norm(N), _Weight(B, N, 1), _Seq(*repeats), _Weight(C * N, B, 1))


class _ConvTasNetFn(torch.autograd.Function):
    """forward/backward of the whole network as two C-ABI calls (ctn_model_forward / ctn_model_backward[_stage])."""

    @staticmethod
    def forward(ctx, model, mixture, _anchor):
        est, ws, token = model._run_forward(mixture, training=True)
        ctx.model, ctx.ws, ctx.token = model, ws, token
        # data parallel with uneven shards: world * M_local / M_global, set by ShardedDataParallel.forward
        ctx.scale, model._grad_scale = model._grad_scale, None
        ctx.save_for_backward(mixture)
        return est

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_est):
        (mixture,) = ctx.saved_tensors
        if ctx.scale is not None:
            d_est = d_est * ctx.scale
        ctx.model._run_backward(mixture, d_est.contiguous(), ctx.ws)
        ctx.ws = ctx.token = None
        return None, None, None


class ConvTasNet(nn.Module):
    def __init__(self, N, L, B, H, P, X, R, C, norm_type="gLN", causal=False, mask_nonlinear='relu'):
        """Same arguments as src/conv_tasnet.py:14-30."""
        super().__init__()
        self.N, self.L, self.B, self.H, self.P, self.X, self.R, self.C = N, L, B, H, P, X, R, C
        self.norm_type = norm_type
        self.causal = causal
        self.mask_nonlinear = mask_nonlinear
        self._cfg = _lib.make_config(N, L, B, H, P, X, R, C, norm_type, causal, mask_nonlinear)
        if self._cfg.mask_nonlinear < 0:
            # the reference only objects at forward time (src/conv_tasnet.py:213-214): keep a valid geometry config and
            # let _run_forward raise the ValueError
            self._cfg.mask_nonlinear = 0
        self.encoder = Encoder(L, N)
        self.separator = TemporalConvNet(N, B, H, P, X, R, C, causal, norm_type)
        self.decoder = Decoder(N, L)
        for p in self.parameters():  # includes gamma/beta [1,C,1] (src/conv_tasnet.py:41-43)
            if p.dim() > 1:
                nn.init.xavier_normal_(p)
        # flat storage (built lazily on the device the parameters live on)
        self._flat = None
        self._flat_grad = None
        self._layout = None
        self._ws_cache = {}       # training flag -> (grow-only workspace, weakref of the autograd token using it)
        self._ws_override = None  # set by graph.GraphedTrainStep / GraphedInference: the graph owns its workspace
        # torch.bfloat16: no_grad forwards store the H-wide activations as bf16 and use single-bf16 MMAs ("bf16 forward",
        # BASELINE configs[2]; opt-in — the reference has no reduced-precision path, outputs stay within 2e-2 of fp32)
        self.inference_dtype = torch.float32
        self._grad_sync = None  # set by data_parallel.ShardedDataParallel
        self._grad_scale = None  # set per forward by ShardedDataParallel (uneven shards), consumed by the backward
        self._overwrite_next = False  # set by optim.FusedAdam: the next backward overwrites the flat gradients
        self._plist = None
        # BatchNorm branch: running statistics of every nn.BatchNorm1d in ONE flat buffer (the C ABI's norm_state)
        self._bns = [m for m in self.modules() if isinstance(m, nn.BatchNorm1d)]
        self._bn_state = None
        self._bn_count = None

    # ------------------------------------------------------------------ flat parameter storage
    def _param_layout(self):
        if self._layout is None:
            L = _lib.lib()
            n = L.ctn_param_tensors(ctypes.byref(self._cfg))
            offs, nums = (ctypes.c_int64 * n)(), (ctypes.c_int64 * n)()
            _lib.check(L.ctn_param_layout(ctypes.byref(self._cfg), offs, nums, n))
            total = L.ctn_param_floats(ctypes.byref(self._cfg))
            self._layout = (list(offs), list(nums), int(total))
        return self._layout

    def _flat_ok(self):
        if self._flat is None or self._plist is None:
            return False
        base = self._flat.data_ptr()
        offs = self._layout[0]
        for i in (0, len(self._plist) // 2, len(self._plist) - 1):
            if self._plist[i].data_ptr() != base + 4 * offs[i]:
                return False
        if self._bns:
            H, st = self.H, self._bn_state
            if st is None or self._bns[0].running_mean.data_ptr() != st.data_ptr() \
                    or self._bns[-1].running_var.data_ptr() != st.data_ptr() + 4 * (2 * len(self._bns) - 1) * H \
                    or self._bns[-1].num_batches_tracked.data_ptr() != self._bn_count.data_ptr() + 8 * (len(self._bns) - 1):
                return False
        return True

    def _flatten(self):
        """(Re)build the flat buffer on the parameters' current device and re-point every parameter at its slice."""
        plist = list(self.parameters())
        offs, nums, total = self._param_layout()
        if len(plist) != len(offs):
            raise RuntimeError("parameter inventory does not match the C layout")
        dev = plist[0].device
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o, n in zip(plist, offs, nums):
            if p.numel() != n or p.dtype != torch.float32:
                raise TypeError("ConvTasNet parameters must stay fp32 with the reference shapes "
                                "(reduced precision has no reference path, SURVEY §8c)")
            flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = flat[o:o + n].view(p.shape)
        self._flat, self._plist = flat, plist
        self._flat_grad = None
        self._ws_cache = {}
        if self._bns:  # modules() order = block order, norm 1 then norm 2: the layout ctn_model_forward_bn expects
            H = self.H
            state = torch.empty(2 * len(self._bns) * H, dtype=torch.float32, device=dev)
            count = torch.empty(len(self._bns), dtype=torch.int64, device=dev)
            for i, m in enumerate(self._bns):
                state[2 * i * H:(2 * i + 1) * H].copy_(m.running_mean)
                state[(2 * i + 1) * H:(2 * i + 2) * H].copy_(m.running_var)
                count[i] = m.num_batches_tracked
                m.running_mean = state[2 * i * H:(2 * i + 1) * H]
                m.running_var = state[(2 * i + 1) * H:(2 * i + 2) * H]
                m.num_batches_tracked = count[i]
            self._bn_state, self._bn_count = state, count

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._flat = None  # .cuda()/.to() moved every parameter separately; re-flatten on next use
        return out

    @property
    def flat_params(self):
        if not self._flat_ok():
            self._flatten()
        return self._flat

    @property
    def flat_grads(self):
        self.flat_params
        if self._flat_grad is None:
            self._flat_grad = torch.zeros_like(self._flat)
        return self._flat_grad

    def grad_views(self):
        """per-parameter views into the flat gradient buffer, in parameters() order"""
        offs, nums, _ = self._param_layout()
        fg = self.flat_grads
        return [fg[o:o + n].view(p.shape) for p, o, n in zip(self._plist, offs, nums)]

    # ------------------------------------------------------------------ forward / backward plumbing
    def workspace_bytes(self, M, T, training):
        nbytes = _lib.lib().ctn_workspace_bytes(ctypes.byref(self._cfg), M, T, 1 if training else 0)
        if nbytes < 0:
            _lib.check(1)
        return nbytes

    def _workspace(self, M, T, training):
        """-> (workspace, is_the_cached_one).

        ONE grow-only workspace per `training` flag, reused for every (M, T) that fits (the reference's loops feed
        variable-length batches: src/solver.py:183-188 cross-validation, src/evaluate.py:44, src/separate.py:44 — a
        workspace per distinct shape would pin memory until OOM).  While an autograd graph still holds the cached
        workspace's activation stash (its token is alive) a temporary one is allocated instead.  CUDA-graph wrappers
        own their workspace and pass it through `_ws_override`."""
        nbytes = self.workspace_bytes(M, T, training)
        if self._ws_override is not None:
            if self._ws_override.numel() < nbytes:
                raise RuntimeError("graph-owned workspace is too small for this shape")
            return self._ws_override, False
        key = bool(training)
        entry = self._ws_cache.get(key)
        busy = entry is not None and entry[1] is not None and entry[1]() is not None
        if entry is not None and not busy and entry[0].numel() >= nbytes:
            return entry[0], True
        if entry is not None and not busy:
            self._ws_cache.pop(key)  # too small: release it before the larger allocation
            del entry
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self._flat.device)
        if not busy:
            self._ws_cache[key] = (ws, None)
            return ws, True
        return ws, False

    def _run_forward(self, mixture, training):
        if self.mask_nonlinear not in ("relu", "softmax"):
            raise ValueError("Unsupported mask non-linear function")  # src/conv_tasnet.py:213-214
        if mixture.dim() != 2:
            raise ValueError("mixture must be [M, T]")
        if not mixture.is_cuda:
            raise RuntimeError("conv_tasnet_b200.ConvTasNet runs on CUDA only (no CPU fallback); move the model and "
                               "the mixture to a B200 with .cuda()")
        if mixture.dtype != torch.float32:
            raise TypeError("mixture must be float32 (the reference path is fp32 only, SURVEY §8c)")
        flat = self.flat_params
        if flat.device != mixture.device:
            raise RuntimeError(f"model is on {flat.device}, mixture on {mixture.device}")
        mixture = mixture.contiguous()
        M, T = mixture.shape
        with torch.cuda.device(mixture.device):
            ws, cached = self._workspace(M, T, training)
            est = torch.empty(M, self.C, T, dtype=torch.float32, device=mixture.device)
            if self._bns:  # nn.BatchNorm1d semantics: batch statistics + running update in .train(), running in .eval()
                _lib.check(_lib.lib().ctn_model_forward_bn(
                    ctypes.byref(self._cfg), _lib.ptr(flat), _lib.ptr(self._bn_state), _lib.ptr(mixture), M, T,
                    _lib.ptr(est), _lib.ptr(ws), ws.numel(), 1 if training else 0, 1 if self.training else 0,
                    _lib.stream()))
                if self.training:
                    self._bn_count += 1
            else:
                mode = 1 if training else (2 if self.inference_dtype == torch.bfloat16 else 0)
                _lib.check(_lib.lib().ctn_model_forward(ctypes.byref(self._cfg), _lib.ptr(flat), _lib.ptr(mixture), M,
                                                        T, _lib.ptr(est), _lib.ptr(ws), ws.numel(), mode, _lib.stream()))
        token = None
        if training:
            token = _Token()
            if cached:
                self._ws_cache[True] = (ws, weakref.ref(token))
        return est, ws, token

    def _run_backward(self, mixture, d_est, ws):
        M, T = mixture.shape
        L = _lib.lib()
        plist = self._plist
        grads = self.flat_grads
        n_none = sum(1 for p in plist if p.grad is None)
        views = None
        if n_none == len(plist):
            accumulate, target = 0, grads
        elif n_none == 0 and plist[0].grad.data_ptr() == grads.data_ptr() + 4 * self._layout[0][0] \
                and plist[-1].grad.data_ptr() == grads.data_ptr() + 4 * self._layout[0][-1]:
            accumulate, target = (0 if self._overwrite_next else 1), grads
        else:  # mixed ownership of .grad: compute aside and add per tensor (slow path)
            accumulate, target = 0, torch.empty_like(grads)
        self._overwrite_next = False
        args = (ctypes.byref(self._cfg), _lib.ptr(self._flat), _lib.ptr(mixture), M, T, _lib.ptr(d_est),
                _lib.ptr(target), _lib.ptr(ws), ws.numel(), accumulate)
        with torch.cuda.device(mixture.device):
            if self._grad_sync is None:
                _lib.check(L.ctn_model_backward(*args, _lib.stream()))
            elif target is not grads:
                # mixed ownership of .grad under data parallelism: the gradients computed aside are all-reduced as one
                # buffer before they are added to the user's tensors (the replicas must not diverge)
                _lib.check(L.ctn_model_backward(*args, _lib.stream()))
                self._grad_sync(self, -2, target)
            else:
                for stage in range(self.R + 2):
                    _lib.check(L.ctn_model_backward_stage(*args, stage, _lib.stream()))
                    self._grad_sync(self, stage)
                self._grad_sync(self, -1)  # drain
        if n_none == len(plist):
            views = self.grad_views()
            for p, v in zip(plist, views):
                p.grad = v
        elif target is not grads:
            offs, nums, _ = self._param_layout()
            for p, o, n in zip(plist, offs, nums):
                v = target[o:o + n].view(p.shape)
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    p.grad += v

    def _backward_stage(self, mixture, d_est, ws, stage):
        """One stage of the staged backward straight into the flat gradient buffer (stage 0 zero-initialises it).
        Used by graph.GraphedTrainStep, which captures one CUDA graph per stage so that the data-parallel all-reduce of
        a finished slice overlaps the next stage's graph."""
        M, T = mixture.shape
        grads = self.flat_grads
        with torch.cuda.device(mixture.device):
            _lib.check(_lib.lib().ctn_model_backward_stage(
                ctypes.byref(self._cfg), _lib.ptr(self._flat), _lib.ptr(mixture), M, T, _lib.ptr(d_est),
                _lib.ptr(grads), _lib.ptr(ws), ws.numel(), 0, stage, _lib.stream()))

    def half_inference(self, enabled=True):
        """Opt into the reduced-precision inference path: no_grad forwards keep the H-wide activations in bf16 and feed
        the tensor cores single bf16 operands (fp32 accumulation, fp32 residual stream / statistics / I/O).  Training
        and forwards under autograd are unaffected.  Returns self."""
        if enabled and (self._bns or self.N % 64 or self.B % 64 or self.H % 64):
            raise ValueError("half_inference needs gLN / cLN and N, B, H multiples of 64")
        self.inference_dtype = torch.bfloat16 if enabled else torch.float32
        return self

    def grad_bucket(self, stage):
        off, cnt = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.lib().ctn_grad_bucket(ctypes.byref(self._cfg), stage, ctypes.byref(off), ctypes.byref(cnt)))
        return off.value, cnt.value

    def forward(self, mixture):
        """
        Args:
            mixture: [M, T], M is batch size, T is #samples
        Returns:
            est_source: [M, C, T]
        (src/conv_tasnet.py:45-60)
        """
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            self.flat_params  # make sure the anchor parameter is a live view
            return _ConvTasNetFn.apply(self, mixture, self._plist[0])
        est, _, _ = self._run_forward(mixture, training=False)
        return est

    # ------------------------------------------------------------------ (de)serialisation, src/conv_tasnet.py:62-94
    @classmethod
    def load_model(cls, path):
        package = torch.load(path, map_location=lambda storage, loc: storage)
        return cls.load_model_from_package(package)

    @classmethod
    def load_model_from_package(cls, package):
        model = cls(package['N'], package['L'], package['B'], package['H'],
                    package['P'], package['X'], package['R'], package['C'],
                    norm_type=package['norm_type'], causal=package['causal'],
                    mask_nonlinear=package['mask_nonlinear'])
        model.load_state_dict(package['state_dict'])
        return model

    @staticmethod
    def serialize(model, optimizer, epoch, tr_loss=None, cv_loss=None):
        package = {
            'N': model.N, 'L': model.L, 'B': model.B, 'H': model.H,
            'P': model.P, 'X': model.X, 'R': model.R, 'C': model.C,
            'norm_type': model.norm_type, 'causal': model.causal,
            'mask_nonlinear': model.mask_nonlinear,
            'state_dict': model.state_dict(),
            'optim_dict': optimizer.state_dict(),
            'epoch': epoch,
        }
        if tr_loss is not None:
            package['tr_loss'] = tr_loss
            package['cv_loss'] = cv_loss
        return package


class _Token:
    """Alive while an autograd graph still needs the activation stash of a workspace."""
    __slots__ = ("__weakref__",)
