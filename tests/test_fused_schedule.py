"""CPU proof that the fused kernel schedule (oracle/fused_schedule.py: channels-last, norm folded into the
next 1x1 conv, hand-derived backward, moments-form PIT) equals the op-by-op oracle — forward, loss, and every
gradient — in fp64 (tight) and fp32 (the tolerance the CUDA path is held to)."""
import pytest
import torch

from conftest import golden_model, load_golden, rel_err
from oracle import conv_tasnet_oracle as O
from oracle import fused_schedule as FS

SMALL = ["gln", "cln_causal", "softmax_c3", "gln_causal_p2", "cln_p5"]


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("dtype,tol_out,tol_grad", [(torch.float64, 1e-10, 1e-8), (torch.float32, 1e-4, 1e-3)])
def test_fused_schedule_equals_oracle(name, dtype, tol_out, tol_grad):
    cfgd, sd, z = golden_model(name)
    cfg = O.Config(**cfgd)
    sd = {k: v.to(dtype) for k, v in sd.items()}
    mix = torch.from_numpy(z["mixture"]).to(dtype)
    src = torch.from_numpy(z["source"]).to(dtype)
    lens = torch.from_numpy(z["lengths"])
    loss_o, est_o, grads_o, max_snr_o, reord_o = O.train_step_grads(cfg, sd, mix, src, lens)
    pf, est, grads = FS.train_step(cfg, sd, mix, src, lens)
    assert rel_err(pf["est_masked"], est_o) < tol_out
    assert abs(pf["loss"].item() - loss_o.item()) < (1e-8 if dtype == torch.float64 else 1e-3)
    assert rel_err(pf["max_snr"], max_snr_o) < max(tol_out, 1e-5)
    assert rel_err(pf["reorder"], reord_o) < tol_out
    assert set(grads) == set(grads_o)
    for k in grads_o:
        assert grads[k].shape == grads_o[k].shape, k
        assert rel_err(grads[k], grads_o[k]) < tol_grad, k
    # and against the reference's own fp32 numbers
    assert rel_err(est, z["est_source"]) < 1e-4
    for k in grads_o:
        assert rel_err(grads[k], z["g:" + k]) < 1e-3, k


@pytest.mark.parametrize("name", ["bn", "bn_causal_c3"])
@pytest.mark.parametrize("training", [True, False])
def test_fused_schedule_batchnorm_branch(name, training):
    """BatchNorm as the kernels compute it (identity statistics + per-channel (s, t); backward from the per-channel
    sums A, B) equals nn.BatchNorm1d semantics of the op-by-op oracle in fp64: outputs, every gradient, and the
    running statistics left behind; and the reference's fp32 vectors."""
    cfgd, sd, z = golden_model(name)
    cfg = O.Config(**cfgd)
    to64 = lambda d: {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in d.items()}
    mix = torch.from_numpy(z["mixture"]).double()
    src = torch.from_numpy(z["source"]).double()
    lens = torch.from_numpy(z["lengths"])
    sd_o, sd_f = to64(sd), to64(sd)
    loss_o, est_o, grads_o, _, _ = O.train_step_grads(cfg, sd_o, mix, src, lens, training=training)
    pf, est, grads = FS.train_step(cfg, sd_f, mix, src, lens, training=training)
    assert rel_err(pf["est_masked"], est_o) < 1e-10
    assert abs(pf["loss"].item() - loss_o.item()) < 1e-8
    assert set(grads) == set(grads_o)
    for k in grads_o:
        assert grads[k].shape == grads_o[k].shape, k
        assert rel_err(grads[k], grads_o[k]) < 1e-8, k
    for k in sd_o:
        if O.is_buffer(k):
            assert rel_err(sd_f[k], sd_o[k]) < 1e-12, k
    pre = "g:" if training else "ge:"
    assert rel_err(est, z["est_source" if training else "eval_est_source"]) < 1e-4
    errs = sorted(rel_err(grads[k], z[pre + k]) for k in grads_o)
    assert errs[len(errs) // 2] < 1e-4 and errs[-1] < 5e-3


def test_pit_moments_form_matches_reference_cases():
    z = load_golden("pit.npz")
    for i in range(int(z["n_cases"])):
        src = torch.from_numpy(z[f"c{i}_source"])
        est = torch.from_numpy(z[f"c{i}_est"])
        lens = torch.from_numpy(z[f"c{i}_lengths"])
        for dt in (torch.float64, torch.float32):
            pf = FS.pit_fwd(src.to(dt), est.to(dt), lens)
            assert torch.equal(pf["idx"], torch.from_numpy(z[f"c{i}_idx"]))  # bit exact permutation choice
            assert abs(pf["loss"].item() - float(z[f"c{i}_loss"])) < 1e-3
            assert rel_err(pf["max_snr"], z[f"c{i}_max_snr"]) < 1e-4
            assert torch.equal(pf["est_masked"].float(), torch.from_numpy(z[f"c{i}_est_masked"]))
            assert torch.equal(pf["reorder"].float(), torch.from_numpy(z[f"c{i}_reorder"]))
            g = FS.pit_bwd(src.to(dt), pf["est_masked"], pf)
            assert rel_err(g, z[f"c{i}_grad_est"]) < (1e-6 if dt == torch.float64 else 1e-3)
