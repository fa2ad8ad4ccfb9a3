"""Golden vectors for the BatchNorm branch (norm_type other than gLN / cLN, src/conv_tasnet.py:306-309), produced by
RUNNING THE REFERENCE ITSELF in the build container (needs /root/reference; nothing at test time reads it):

    python tests/golden/make_golden_bn.py

Per case: the initial state_dict (BatchNorm weight / bias perturbed away from their defaults, running statistics those
of a trained-like state, so that nothing is an identity), an evaluation-mode forward + backward (running statistics), then a training-mode
forward + loss + backward (batch statistics) and the running statistics it leaves behind.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CTN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from src.conv_tasnet import ConvTasNet  # noqa: E402
from src.pit_criterion import cal_loss  # noqa: E402
from make_golden import np_, save, synthetic  # noqa: E402

CASES = {
    "bn": dict(N=16, L=8, B=8, H=16, P=3, X=3, R=2, C=2, norm_type="BN", causal=False, mask_nonlinear="relu"),
    "bn_causal_c3": dict(N=12, L=6, B=8, H=20, P=3, X=2, R=2, C=3, norm_type="BN", causal=True, mask_nonlinear="softmax"),
}


def step(model, mix, src, lengths):
    est = model(mix)
    est_raw = est.detach().clone()
    loss, max_snr, est_masked, reord = cal_loss(src, est, lengths)
    model.zero_grad()
    loss.backward()
    grads = {k: np_(p.grad) for k, p in model.named_parameters()}
    return est_raw, loss, max_snr, grads


def main():
    for name, cfg in CASES.items():
        torch.manual_seed(5)
        model = ConvTasNet(**cfg)
        g = torch.Generator().manual_seed(99)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if p.dim() == 1 and p.numel() == 1:      # PReLU slopes: distinct values
                    p.copy_(0.05 + 0.4 * torch.rand(1, generator=g))
                elif p.dim() == 1 and n.endswith("weight"):  # BatchNorm weight
                    p.copy_(0.5 + torch.rand(p.shape, generator=g))
                elif p.dim() == 1:                        # BatchNorm bias
                    p.copy_(0.2 * torch.randn(p.shape, generator=g))
        # running statistics of a "trained" state: the batch statistics of another batch (momentum 1 for one training
        # forward), then perturbed by up to +-20 % so that evaluation mode differs from training mode
        bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm1d)]
        for m in bns:
            m.momentum = 1.0
        with torch.no_grad():
            model(synthetic(3, 403, cfg["C"], cfg["L"], 18)[0])
            for m in bns:
                m.momentum = 0.1
                m.running_var.mul_(0.8 + 0.4 * torch.rand(m.running_var.shape, generator=g))
                m.running_mean.mul_(0.8 + 0.4 * torch.rand(m.running_mean.shape, generator=g))
                m.num_batches_tracked.zero_()
        mix, src, lengths = synthetic(3, 403, cfg["C"], cfg["L"], 17)
        arrays = {"cfg_" + k: np.array(v) for k, v in cfg.items()}
        arrays.update(mixture=np_(mix), source=np_(src), lengths=np_(lengths))
        for k, v in model.state_dict().items():
            arrays["w:" + k] = np_(v).copy()
        model.eval()
        est, loss, max_snr, grads = step(model, mix, src, lengths)
        arrays.update(eval_est_source=np_(est), eval_loss=np_(loss), eval_max_snr=np_(max_snr))
        arrays.update({"ge:" + k: v for k, v in grads.items()})
        for k, v in model.state_dict().items():  # evaluation must not touch the buffers
            assert np.array_equal(np_(v), arrays["w:" + k]), k
        model.train()
        est, loss, max_snr, grads = step(model, mix, src, lengths)
        arrays.update(est_source=np_(est), loss=np_(loss), max_snr=np_(max_snr))
        arrays.update({"g:" + k: v for k, v in grads.items()})
        for k, v in model.named_buffers():
            arrays["after:" + k] = np_(v)
        save(f"model_{name}.npz", **arrays)


if __name__ == "__main__":
    main()
