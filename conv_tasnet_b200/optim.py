"""The step tail of src/solver.py:192-196 — clip_grad_norm_(max_norm) + Adam — fused on the model's flat parameter /
gradient buffers (ctn_clip_grad_norm, ctn_adam_step): 4 kernel launches instead of ~30 multi-tensor ones over 294
tensors, and no host sync (the norm stays on the device unless asked for).

`state_dict()` / `load_state_dict()` speak torch.optim.Adam's format ({'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}},
'param_groups': [{..., 'params': [0..n-1]}]}), so the `optim_dict` of a checkpoint package (src/conv_tasnet.py:89,
src/solver.py:62,126-129) is interchangeable with the reference's torch.optim.Adam in both directions."""
import torch

from . import _lib


class FusedAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics on ConvTasNet.flat_params / flat_grads."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None):
        self.model = model
        p = model.flat_params
        if not p.is_cuda:
            raise RuntimeError("FusedAdam runs on CUDA only (no CPU fallback); move the model to the GPU first")
        n = len(model._plist)
        # hyper-parameters live where torch keeps them, so `optimizer.param_groups[0]['lr'] = ...` works
        self.param_groups = [{"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay,
                              "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                              "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                              "params": list(range(n))}]
        self.max_grad_norm = max_grad_norm
        self.exp_avg = torch.zeros_like(p)
        self.exp_avg_sq = torch.zeros_like(p)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=p.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=p.device)
        self._scratch = torch.empty(8192, dtype=torch.uint8, device=p.device)

    # scalar views of the single parameter group (read at every step, so edits to param_groups take effect)
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = v

    @property
    def betas(self):
        return tuple(self.param_groups[0]["betas"])

    @property
    def eps(self):
        return self.param_groups[0]["eps"]

    @property
    def weight_decay(self):
        return self.param_groups[0]["weight_decay"]

    def hyper(self):
        """Everything a captured CUDA graph bakes in by value (graph.GraphedTrainStep re-captures when it changes)."""
        return (float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                None if self.max_grad_norm is None else float(self.max_grad_norm))

    def zero_grad(self, set_to_none=False):
        # the next backward overwrites the flat gradient buffer (no memset pass over 35 MB, no 294 tiny kernels)
        self.model._overwrite_next = True
        if set_to_none:
            for p in self.model.parameters():
                p.grad = None

    def step(self):
        m = self.model
        p, g = m.flat_params, m.flat_grads
        if p.device != self.exp_avg.device:
            raise RuntimeError("model moved to another device after FusedAdam was built; rebuild the optimizer")
        if p.numel() != self.exp_avg.numel():
            raise RuntimeError("model was re-flattened with a different size after FusedAdam was built")
        L = _lib.lib()
        with torch.cuda.device(p.device):
            if self.max_grad_norm is not None:
                _lib.check(L.ctn_clip_grad_norm(_lib.ptr(g), g.numel(), float(self.max_grad_norm),
                                                _lib.ptr(self.grad_norm), _lib.ptr(self._scratch), _lib.stream()))
            _lib.check(L.ctn_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                       p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                       _lib.ptr(self.step_count), _lib.stream()))

    # ------------------------------------------------------------------ torch.optim.Adam-compatible checkpoints
    def state_dict(self):
        m = self.model
        offs, nums, _ = m._param_layout()
        m.flat_params  # make sure _plist is current
        state = {}
        if int(self.step_count.item()) > 0:  # torch creates the per-parameter state lazily at the first step
            step = self.step_count.to(torch.float32).reshape(())
            for i, (p, o, n) in enumerate(zip(m._plist, offs, nums)):
                # copies, not views: the flat moment buffers keep being updated in place, and torch's
                # Optimizer.load_state_dict adopts the tensors it is given without copying them
                state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        return {"state": state, "param_groups": [dict(self.param_groups[0])]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in g:
                self.param_groups[0][k] = tuple(g[k]) if k == "betas" else g[k]
        if g.get("amsgrad", False) or g.get("maximize", False):
            raise ValueError("FusedAdam implements plain Adam (amsgrad=False, maximize=False)")
        if "state" not in sd:  # round-1 flat form
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.step_count.copy_(sd["step"])
            return
        m = self.model
        offs, nums, _ = m._param_layout()
        state = sd["state"]
        if len(state) == 0:
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            self.step_count.zero_()
            return
        if len(state) != len(offs):
            raise ValueError(f"optimizer state has {len(state)} entries, the model has {len(offs)} parameters")
        steps = set()
        for i, (o, n) in enumerate(zip(offs, nums)):
            st = state[i] if i in state else state[str(i)]
            self.exp_avg[o:o + n].copy_(torch.as_tensor(st["exp_avg"]).reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(torch.as_tensor(st["exp_avg_sq"]).reshape(-1))
            steps.add(int(torch.as_tensor(st["step"]).item()))
        if len(steps) != 1:
            raise ValueError("FusedAdam keeps one step counter: every parameter must have taken the same number of steps")
        self.step_count.fill_(steps.pop())
