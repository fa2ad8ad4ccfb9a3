"""Parity of every CUDA kernel (called through the C ABI) with the kernel-level oracle oracle/fused_schedule.py,
on seeded inputs, at sizes the oracle finishes in seconds.  Tolerances are max|a-b|/max|b| (conftest.rel_err):
fp32 forward ops 1e-5 (north-star budget is 1e-4 end to end), reductions/gradients 1e-4 (budget 1e-3)."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from gpu_util import P, call, dev, gln_acc, stats_args
from oracle import conv_tasnet_oracle as O
from oracle import fused_schedule as FS

pytestmark = pytest.mark.gpu
TOL_F, TOL_G = 1e-5, 1e-4
TOL_TC = 2e-5  # 1x1 convs on tcgen05 use the bf16x3 split: per-GEMM error bound ~2^-16 (measured 5e-6)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev())


def cfg_of(**kw):
    base = dict(N=16, L=8, B=8, H=16, P=3, X=2, R=1, C=2, norm_type="gLN", causal=False, mask_nonlinear="relu")
    base.update(kw)
    return O.Config(**base)


@pytest.mark.parametrize("M,T,N,L", [(2, 403, 16, 8), (3, 32000, 256, 20), (1, 131, 8, 4), (2, 100, 12, 5),
                                     (2, 2000, 512, 32), (1, 1000, 320, 20), (2, 4000, 256, 40), (1, 900, 64, 64),
                                     (2, 700, 32, 34)])
def test_encoder_fwd_bwd(M, T, N, L):
    mix, U = rnd(M, T, seed=1, scale=0.1), rnd(N, L, seed=2, scale=0.3)
    K = O.n_frames(T, L)
    w = torch.empty(M, K, N, device=dev())
    call("ctn_encoder_fwd", P(mix), P(U), M, T, N, L, P(w))
    want = FS.encoder_fwd(mix.cpu(), U.cpu())
    assert rel_err(w.cpu(), want) < TOL_F
    dwa, dwb = rnd(M, K, N, seed=3), rnd(M, K, N, seed=4)
    dU = torch.zeros(N, L, device=dev())
    call("ctn_encoder_bwd", P(mix), P(w), P(dwa), P(dwb), M, T, N, L, P(dU))
    want = FS.encoder_bwd(mix.cpu().double(), want.double(), (dwa + dwb).cpu().double(), L)
    assert rel_err(dU.cpu(), want) < TOL_G


@pytest.mark.parametrize("F,Ch,prelu", [(403, 16, False), (9597, 512, True), (77, 20, True)])
def test_row_stats(F, Ch, prelu):
    x = rnd(1, F, Ch, seed=5) + 0.3
    alpha = torch.tensor([0.17], device=dev()) if prelu else None
    out = torch.empty(F, 2, device=dev())
    call("ctn_row_stats", P(x), P(alpha), F, Ch, P(out))
    a = FS.prelu(x.cpu().double(), 0.17) if prelu else x.cpu().double()
    mu, r = FS.row_stats(a)
    assert rel_err(out[:, 0].cpu(), mu.view(-1)) < TOL_F
    assert rel_err(out[:, 1].cpu(), r.view(-1)) < TOL_F


@pytest.mark.parametrize("M,K,Kd,Ochan", [(2, 99, 8, 16), (3, 3199, 256, 512), (3, 3199, 512, 256), (1, 130, 36, 8),
                                            (2, 257, 20, 12)])
@pytest.mark.parametrize("kn", [0, 1])
def test_conv1x1_plain_and_stats(M, K, Kd, Ochan, kn):
    A, W = rnd(M, K, Kd, seed=6), rnd(Ochan, Kd, seed=7, scale=1 / math.sqrt(Kd))
    Wdev = W.t().contiguous() if kn else W
    D = torch.empty(M, K, Ochan, device=dev())
    alpha_out = torch.tensor([0.21], device=dev())
    stat = torch.zeros(M, 2, dtype=torch.float64, device=dev())
    call("ctn_conv1x1", P(A), P(Wdev), kn, P(D), M * K, Ochan, Kd, K, None, None, None, None, None, None, P(stat),
         P(alpha_out))
    want = A.cpu().double() @ W.cpu().double().t()
    assert rel_err(D.cpu(), want) < TOL_TC
    assert rel_err(stat.cpu(), gln_acc(FS.prelu(want, 0.21))) < TOL_TC


@pytest.mark.parametrize("norm", ["gLN", "cLN"])
@pytest.mark.parametrize("M,K,Kd,Ochan", [(2, 99, 16, 8), (3, 3199, 512, 256), (5, 37, 20, 12)])
def test_conv1x1_prelu_normfold_residual(norm, M, K, Kd, Ochan):
    z, W = rnd(M, K, Kd, seed=8) + 0.2, rnd(Ochan, Kd, seed=9, scale=1 / math.sqrt(Kd))
    gamma, beta, res = rnd(Kd, seed=10), rnd(Kd, seed=11), rnd(M, K, Ochan, seed=12)
    alpha = torch.tensor([0.3], device=dev())
    a = FS.prelu(z.cpu().double(), 0.3)
    acc, rs, (mu, r) = stats_args(norm, a)
    acc, rs = (None if acc is None else acc.to(dev())), (None if rs is None else rs.to(dev()))
    Wg, c1, c2 = torch.empty_like(W), torch.empty(Ochan, device=dev()), torch.empty(Ochan, device=dev())
    call("ctn_prep_normfold", P(W), P(gamma), P(beta), Ochan, Kd, P(Wg), P(c1), P(c2))
    assert rel_err(Wg.cpu(), (W * gamma.view(1, -1)).cpu()) < 1e-6
    assert rel_err(c1.cpu(), (W.cpu().double() @ beta.cpu().double())) < TOL_F
    D = torch.empty(M, K, Ochan, device=dev())
    call("ctn_conv1x1", P(z), P(Wg), 0, P(D), M * K, Ochan, Kd, K, P(alpha), P(c1), P(c2), P(acc), P(rs), P(res), None,
         None)
    mu_d, r_d = (FS.sample_stats(a) if norm == "gLN" else FS.row_stats(a))
    want = FS.gemm_normfold(a, mu_d, r_d, W.cpu().double(), gamma.cpu().double(), beta.cpu().double(),
                            res.cpu().double())
    assert rel_err(D.cpu(), want) < TOL_TC


@pytest.mark.parametrize("norm", [None, "gLN", "cLN"])
@pytest.mark.parametrize("M,K,Ochan,I", [(2, 99, 8, 16), (3, 3199, 256, 512), (3, 3199, 512, 256), (4, 50, 36, 8)])
def test_wgrad(norm, M, K, Ochan, I):
    G, z = rnd(M, K, Ochan, seed=13), rnd(M, K, I, seed=14) + 0.1
    dW = torch.zeros(Ochan, I, device=dev())
    if norm is None:
        call("ctn_wgrad", P(G), P(z), P(dW), M * K, Ochan, I, K, None, None, None, None, None)
        act = z.cpu().double()
    else:
        gamma, beta = rnd(I, seed=15), rnd(I, seed=16)
        alpha = torch.tensor([0.3], device=dev())
        a = FS.prelu(z.cpu().double(), 0.3)
        acc, rs, _ = stats_args(norm, a)
        acc, rs = (None if acc is None else acc.to(dev())), (None if rs is None else rs.to(dev()))
        call("ctn_wgrad", P(G), P(z), P(dW), M * K, Ochan, I, K, P(alpha), P(gamma), P(beta), P(acc), P(rs))
        mu, r = FS.sample_stats(a) if norm == "gLN" else FS.row_stats(a)
        act = gamma.cpu().double().view(1, 1, -1) * (a - mu) * r + beta.cpu().double().view(1, 1, -1)
    want = torch.einsum("mko,mki->oi", G.cpu().double(), act)
    assert rel_err(dW.cpu(), want) < TOL_G


@pytest.mark.parametrize("norm", ["gLN", "cLN"])
@pytest.mark.parametrize("causal,Pk,dil", [(False, 3, 1), (False, 3, 16), (True, 3, 4), (True, 2, 2), (False, 5, 2),
                                           (False, 3, 128)])
@pytest.mark.parametrize("M,K,H", [(2, 99, 16), (3, 400, 512)])
def test_dwconv_fwd_bwd(norm, causal, Pk, dil, M, K, H):
    cfg = cfg_of(causal=causal, P=Pk, H=H, norm_type=norm)
    z1 = rnd(M, K, H, seed=17) + 0.1
    g1, b1, Wd = rnd(H, seed=18), rnd(H, seed=19), rnd(H, Pk, seed=20)
    a1, a2 = torch.tensor([0.25], device=dev()), torch.tensor([0.4], device=dev())
    a = FS.prelu(z1.cpu().double(), 0.25)
    acc, rs, _ = stats_args(norm, a)
    acc, rs = (None if acc is None else acc.to(dev())), (None if rs is None else rs.to(dev()))
    mu, r = FS.sample_stats(a) if norm == "gLN" else FS.row_stats(a)
    z2 = torch.empty_like(z1)
    stat = torch.zeros(M, 2, dtype=torch.float64, device=dev())
    call("ctn_dwconv_fwd", P(z1), P(a1), P(acc), P(rs), P(g1), P(b1), P(Wd), M, K, H, Pk, dil, int(causal), P(z2),
         P(stat), P(a2))
    want = FS.dwconv_fwd(cfg, z1.cpu().double(), 0.25, mu, r, g1.cpu().double(), b1.cpu().double(), Wd.cpu().double(),
                         dil)
    assert rel_err(z2.cpu(), want) < TOL_F
    assert rel_err(stat.cpu(), gln_acc(FS.prelu(want, 0.4))) < TOL_F
    # backward
    dz2 = rnd(M, K, H, seed=21)
    dn1 = torch.empty_like(z1)
    dWd, dg, db = torch.zeros(H, Pk, device=dev()), torch.zeros(H, device=dev()), torch.zeros(H, device=dev())
    red = torch.zeros(M, 2, dtype=torch.float64, device=dev())
    call("ctn_dwconv_bwd", P(dz2), P(z1), P(a1), P(acc), P(rs), P(g1), P(b1), P(Wd), M, K, H, Pk, dil, int(causal),
         P(dn1), P(dWd), P(dg), P(db), P(red))
    w_dn1, w_dWd = FS.dwconv_bwd(cfg, dz2.cpu().double(), z1.cpu().double(), 0.25, mu, r, g1.cpu().double(),
                                 b1.cpu().double(), Wd.cpu().double(), dil)
    assert rel_err(dn1.cpu(), w_dn1) < TOL_F
    assert rel_err(dWd.cpu(), w_dWd) < TOL_G
    yh = (a - mu) * r
    assert rel_err(dg.cpu(), (w_dn1 * yh).sum(dim=(0, 1))) < TOL_G
    assert rel_err(db.cpu(), w_dn1.sum(dim=(0, 1))) < TOL_G
    gh = w_dn1 * g1.cpu().double().view(1, 1, -1)
    want_red = torch.stack([gh.sum(dim=(1, 2)), (gh * yh).sum(dim=(1, 2))], dim=1)
    assert rel_err(red.cpu(), want_red) < TOL_G


@pytest.mark.parametrize("norm", ["gLN", "cLN"])
@pytest.mark.parametrize("prelu", [True, False])
@pytest.mark.parametrize("M,K,Ch", [(2, 99, 16), (3, 700, 512), (1, 33, 20)])
def test_norm_bwd(norm, prelu, M, K, Ch):
    cfg = cfg_of(norm_type=norm)
    z, dn, gamma = rnd(M, K, Ch, seed=22) + 0.1, rnd(M, K, Ch, seed=23), rnd(Ch, seed=24)
    alpha = torch.tensor([0.35], device=dev()) if prelu else None
    a = FS.prelu(z.cpu().double(), 0.35) if prelu else z.cpu().double()
    acc, rs, _ = stats_args(norm, a)
    acc, rs = (None if acc is None else acc.to(dev())), (None if rs is None else rs.to(dev()))
    mu, r = FS.sample_stats(a) if norm == "gLN" else FS.row_stats(a)
    dg, db = torch.zeros(Ch, device=dev()), torch.zeros(Ch, device=dev())
    red = torch.zeros(M, 2, dtype=torch.float64, device=dev())
    dalpha = torch.zeros(1, device=dev())
    call("ctn_norm_bwd_reduce", P(dn), P(z), P(alpha), P(acc), P(rs), P(gamma), M, K, Ch, P(dg), P(db), P(red))
    dz = dn.clone()
    call("ctn_norm_bwd_apply", P(dz), P(z), P(alpha), P(acc), P(rs), P(gamma), P(red), M, K, Ch,
         P(dalpha) if prelu else None)
    w_dz, w_dg, w_db, w_da = FS.norm_bwd(cfg, dn.cpu().double(), z.cpu().double(),
                                         torch.tensor(0.35, dtype=torch.float64) if prelu else None, mu, r,
                                         gamma.cpu().double())
    assert rel_err(dz.cpu(), w_dz) < TOL_G
    assert rel_err(dg.cpu(), w_dg) < TOL_G
    assert rel_err(db.cpu(), w_db) < TOL_G
    if prelu:
        assert abs(dalpha.item() - w_da.item()) < TOL_G * max(1.0, abs(w_da.item()), w_dz.abs().sum().item() * 1e-3)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("M,K,Ch", [(2, 99, 16), (3, 700, 512), (1, 33, 20)])
def test_batchnorm_stats_and_backward(training, M, K, Ch):
    """BatchNorm branch kernels (bn_stats / bn_finalize / bn_bwd_finalize / bn_bwd_apply + the shared per-channel
    reduction) against oracle/fused_schedule.py::bn_forward_stats / bn_bwd"""
    z, dn = rnd(M, K, Ch, seed=31) + 0.1, rnd(M, K, Ch, seed=32)
    weight, bias = rnd(Ch, seed=33).abs() + 0.5, rnd(Ch, seed=34)
    rm, rv = 0.05 * rnd(Ch, seed=35), rnd(Ch, seed=36).abs() + 0.3
    alpha = torch.tensor([0.35], device=dev())
    F = M * K
    sd = {"n.weight": weight.cpu().double(), "n.bias": bias.cpu().double(), "n.running_mean": rm.cpu().double().clone(),
          "n.running_var": rv.cpu().double().clone(), "n.num_batches_tracked": torch.zeros((), dtype=torch.int64)}
    a = FS.prelu(z.cpu().double(), 0.35)
    w_mean, w_rstd, w_s, w_t = FS.bn_forward_stats(sd, "n.", a, training)
    mean, rstd, s, t = (torch.empty(Ch, device=dev()) for _ in range(4))
    scratch = torch.empty(16 * Ch + 16, dtype=torch.uint8, device=dev())
    rm_d, rv_d = rm.clone(), rv.clone()
    call("ctn_batchnorm_stats", P(z), P(alpha), P(weight), P(bias), P(rm_d), P(rv_d), F, Ch, 1 if training else 0,
         P(scratch), P(mean), P(rstd), P(s), P(t))
    for got, want in ((mean, w_mean), (rstd, w_rstd), (s, w_s), (t, w_t), (rm_d, sd["n.running_mean"]),
                      (rv_d, sd["n.running_var"])):
        assert rel_err(got.cpu(), want) < TOL_F
    # backward: per-channel sums through the shared reduction kernel (no statistics = identity), then finalize + apply
    A, Bs = torch.zeros(Ch, device=dev()), torch.zeros(Ch, device=dev())
    call("ctn_norm_bwd_reduce", P(dn), P(z), P(alpha), None, None, P(s), M, K, Ch, P(Bs), P(A), None)
    dw, db, dalpha = torch.zeros(Ch, device=dev()), torch.zeros(Ch, device=dev()), torch.zeros(1, device=dev())
    dz = dn.clone()
    call("ctn_batchnorm_bwd", P(dz), P(z), P(alpha), P(A), P(Bs), P(mean), P(rstd), P(s), 1 if training else 0, F, Ch,
         P(dw), P(db), P(dalpha), P(scratch))
    w_dz, w_dw, w_db, w_da = FS.bn_bwd(dn.cpu().double(), z.cpu().double(), torch.tensor(0.35, dtype=torch.float64),
                                       w_mean, w_rstd, weight.cpu().double(), training)
    assert rel_err(dz.cpu(), w_dz) < TOL_G
    assert rel_err(dw.cpu(), w_dw) < TOL_G
    assert rel_err(db.cpu(), w_db) < TOL_G
    assert abs(dalpha.item() - w_da.item()) < TOL_G * max(1.0, abs(w_da.item()), w_dz.abs().sum().item() * 1e-3)


@pytest.mark.parametrize("softmax", [0, 1])
@pytest.mark.parametrize("M,K,C,N,L,pad", [(2, 99, 2, 16, 8, 3), (3, 3199, 2, 256, 20, 0), (2, 64, 3, 12, 6, 5),
                                            (1, 40, 2, 8, 5, 0), (2, 33, 4, 8, 4, 1), (2, 150, 4, 512, 32, 2),
                                            (1, 77, 3, 320, 20, 0), (2, 500, 2, 256, 32, 7), (2, 199, 2, 256, 40, 3),
                                            (1, 60, 3, 512, 64, 0), (2, 45, 2, 32, 34, 5)])
def test_decoder_fwd_bwd(softmax, M, K, C, N, L, pad):
    cfg = cfg_of(C=C, N=N, L=L, mask_nonlinear="softmax" if softmax else "relu")
    S = L // 2
    T = (K - 1) * S + L + pad
    score, w, V = rnd(M, K, C * N, seed=25), rnd(M, K, N, seed=26).abs(), rnd(L, N, seed=27, scale=0.2)
    est = torch.full((M, C, T), float("nan"), device=dev())
    call("ctn_decoder_fwd", P(score), P(w), P(V), M, K, C, N, L, T, softmax, P(est))
    want = FS.decoder_fwd(cfg, score.cpu().double(), w.cpu().double(), V.cpu().double(), T)
    assert rel_err(est.cpu(), want) < TOL_F
    d_est = rnd(M, C, T, seed=28)
    d_score, d_w = torch.empty_like(score), torch.empty_like(w)
    dV = torch.zeros(L, N, device=dev())
    call("ctn_decoder_bwd", P(d_est), P(score), P(w), P(V), M, K, C, N, L, T, softmax, P(d_score), P(d_w), P(dV))
    w_ds, w_dw, w_dV = FS.decoder_bwd(cfg, d_est.cpu().double(), score.cpu().double(), w.cpu().double(),
                                      V.cpu().double())
    assert rel_err(d_score.cpu(), w_ds) < TOL_F
    assert rel_err(d_w.cpu(), w_dw) < TOL_F
    assert rel_err(dV.cpu(), w_dV) < TOL_G


def test_overlap_and_add_golden_bit_exact():
    from conv_tasnet_b200.utils import overlap_and_add
    z = load_golden("ola.npz")
    out = overlap_and_add(torch.from_numpy(z["kat_signal"]).float().to(dev()), int(z["kat_step"]))
    assert torch.equal(out.cpu(), torch.from_numpy(z["kat_result"]).float())  # reference's own known answer
    for i in range(int(z["n_random"])):
        sig, step = torch.from_numpy(z[f"r{i}_signal"]).to(dev()), int(z[f"r{i}_step"])
        want = torch.from_numpy(z[f"r{i}_result"])
        got = overlap_and_add(sig, step).cpu()
        assert got.shape == want.shape
        if sig.shape[-1] <= 2 * step:
            assert torch.equal(got, want)  # <= 2 contributors: order independent => bit exact
        else:
            assert rel_err(got, want) < 1e-6
    with pytest.raises(ValueError):
        overlap_and_add(torch.zeros(1, 3, 4, device=dev()), 5)


def test_pit_golden_bit_exact_choice():
    from conv_tasnet_b200.pit_criterion import cal_loss, cal_si_snr_with_pit, reorder_source
    z = load_golden("pit.npz")
    # the reference's seed-123 known answer
    src = torch.from_numpy(z["kat_source"]).float().to(dev())
    est = torch.from_numpy(z["kat_est"]).float().to(dev())
    loss, max_snr, est_m, reord = cal_loss(src, est, torch.from_numpy(z["kat_lengths"]))
    assert abs(loss.item() - 45.9221) < 2e-4 and abs(loss.item() - float(z["kat_loss"])) < 1e-4
    assert max_snr.shape == (2, 1) and np.allclose(max_snr.cpu().numpy(), z["kat_max_snr"], atol=1e-4)
    assert abs(reord.double().sum().item() - float(z["kat_reorder_crc"])) < 1e-6
    assert est_m.data_ptr() == est.data_ptr()  # masked in place, same object returned
    for i in range(int(z["n_cases"])):
        src = torch.from_numpy(z[f"c{i}_source"]).to(dev())
        est_in = torch.from_numpy(z[f"c{i}_est"]).to(dev()).requires_grad_(True)
        lens = torch.from_numpy(z[f"c{i}_lengths"])
        est = est_in * 1.0
        loss, max_snr, est_m, reord = cal_loss(src, est, lens if i % 2 else lens.to(dev()))
        loss.backward()
        assert abs(loss.item() - float(z[f"c{i}_loss"])) < 1e-3
        assert rel_err(max_snr.cpu(), z[f"c{i}_max_snr"]) < 1e-4
        assert torch.equal(est_m.detach().cpu(), torch.from_numpy(z[f"c{i}_est_masked"]))
        assert torch.equal(reord.cpu(), torch.from_numpy(z[f"c{i}_reorder"]))  # bit exact reorder
        assert rel_err(est_in.grad.cpu(), z[f"c{i}_grad_est"]) < 1e-3
        ms, perms, idx = cal_si_snr_with_pit(src, torch.from_numpy(z[f"c{i}_est"]).to(dev()), lens)
        assert torch.equal(perms.cpu(), torch.from_numpy(z[f"c{i}_perms"]))
        assert torch.equal(idx.cpu(), torch.from_numpy(z[f"c{i}_idx"]))  # bit exact permutation choice
        r2 = reorder_source(torch.from_numpy(z[f"c{i}_est_masked"]).to(dev()), perms, idx)
        assert torch.equal(r2.cpu(), torch.from_numpy(z[f"c{i}_reorder"]))


def test_pit_full_size_properties():
    """BASELINE sizes: permuting the estimates permutes the chosen index consistently and leaves the loss unchanged."""
    from conv_tasnet_b200.pit_criterion import cal_loss
    g = torch.Generator().manual_seed(3)
    B, C, T = 16, 3, 32000
    src = (torch.randn(B, C, T, generator=g) * 0.05).to(dev())
    est = src[:, [2, 0, 1]] + 0.02 * torch.randn(B, C, T, generator=g).to(dev())
    lens = torch.full((B,), T)
    l1, s1, _, r1 = cal_loss(src, est.clone(), lens)
    l2, s2, _, r2 = cal_loss(src, est[:, [1, 2, 0]].clone(), lens)
    assert abs(l1.item() - l2.item()) < 1e-4
    ref = O.cal_loss(src.cpu(), est.cpu().clone(), lens)
    assert abs(l1.item() - ref[0].item()) < 1e-3
    assert torch.equal(r1.cpu(), ref[3])


def test_clip_and_adam_match_torch():
    n = 100003
    g = torch.Generator().manual_seed(9)
    p0, g0 = torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.1
    pt = torch.nn.Parameter(p0.clone().to(dev()))
    opt = torch.optim.Adam([pt], lr=1e-3)
    p = p0.clone().to(dev())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device=dev())
    scratch = torch.empty(8192, dtype=torch.uint8, device=dev())
    norm = torch.empty(1, device=dev())
    for it in range(3):
        gi = (g0 * (it + 1) * 30).to(dev())
        pt.grad = gi.clone()
        tn = torch.nn.utils.clip_grad_norm_([pt], 5.0)
        opt.step()
        gc = gi.clone()
        call("ctn_clip_grad_norm", P(gc), n, 5.0, P(norm), P(scratch))
        assert abs(norm.item() - tn.item()) < 1e-4 * tn.item()
        assert rel_err(gc.cpu(), pt.grad.cpu()) < 1e-6
        call("ctn_adam_step", P(p), P(gc), P(m), P(v), n, 1e-3, 0.9, 0.999, 1e-8, 0.0, P(step))
        assert rel_err(p.cpu(), pt.detach().cpu()) < 1e-6
    assert step.item() == 3


def test_sisnri_kernel_matches_reference_goldens():
    """ctn_sisnri (csrc/sisnri.cu) through conv_tasnet_b200.evaluate against tests/golden/eval.npz — the reference's own
    cal_SISNRi / cal_SISNR (src/evaluate.py:94-130) run per utterance after remove_pad — and against the fp64 oracle.
    Tolerance in dB: 1e-3 against the reference (which computes in float32 numpy), 1e-4 against the fp64 oracle."""
    from conv_tasnet_b200.evaluate import cal_SISNR, cal_SISNRi, cal_SISNRi_batch
    z = load_golden("eval.npz")
    for i in range(int(z["n_cases"])):
        src, est, mix = (torch.from_numpy(z[f"c{i}_{k}"]).to(dev()) for k in ("src", "est", "mix"))
        lens = torch.from_numpy(z[f"c{i}_lengths"])
        got = cal_SISNRi_batch(src, est, mix, lens).cpu().double()
        assert (got - torch.from_numpy(z[f"c{i}_sisnri"])).abs().max().item() < 1e-3
        for b, n in enumerate(lens.tolist()):
            want = O.cal_SISNRi_np(z[f"c{i}_src"][b, :, :n].astype(np.float64), z[f"c{i}_est"][b, :, :n].astype(np.float64),
                                   z[f"c{i}_mix"][b, :n].astype(np.float64))
            assert abs(got[b].item() - want) < 1e-4
        # the single-utterance entry points of the reference API
        n = int(lens[0])
        one = cal_SISNRi(src[0, :, :n], est[0, :, :n], mix[0, :n]).item()
        assert abs(one - z[f"c{i}_sisnri"][0]) < 1e-3
        s = cal_SISNR(src[0, 1, :n], est[0, 1, :n]).item()
        assert abs(s - z[f"c{i}_sisnr"][0, 1]) < 1e-3


def test_sisnri_kernel_three_speakers_and_empty_tail():
    """C = 3 (the reference hard-codes two speakers; the kernel averages over any C) and an utterance one sample long."""
    from conv_tasnet_b200.evaluate import cal_SISNRi_batch
    g = torch.Generator().manual_seed(21)
    B, C, T = 3, 3, 5000
    src = torch.randn(B, C, T, generator=g) * 0.05
    lens = torch.tensor([5000, 4099, 2])
    for b, n in enumerate(lens.tolist()):
        src[b, :, n:] = 0
    mix = src.sum(1)
    est = src + 0.01 * torch.randn(B, C, T, generator=g)
    got = cal_SISNRi_batch(src.to(dev()), est.to(dev()), mix.to(dev()), lens).cpu()
    for b, n in enumerate(lens.tolist()):
        s, e, m = (x[b, ..., :n].double().numpy() for x in (src, est, mix))
        want = np.mean([O.cal_SISNR_np(s[c], e[c]) - O.cal_SISNR_np(s[c], m) for c in range(C)])
        assert abs(got[b].item() - want) < 1e-3, (b, got[b].item(), want)


def test_get_mask_matches_reference():
    """get_mask (src/pit_criterion.py:102-114) built on the device: ones up to each length, zeros after, [B,1,T]."""
    from conv_tasnet_b200.pit_criterion import get_mask
    src = torch.zeros(4, 2, 37, device=dev())
    lens = torch.tensor([37, 1, 0, 20])
    got = get_mask(src, lens)
    want = O.get_mask(src.cpu(), lens)
    assert got.shape == (4, 1, 37) and got.dtype == src.dtype
    assert torch.equal(got.cpu(), want)
    assert torch.equal(get_mask(src, lens.to(dev())).cpu(), want)  # lengths already on the device


@pytest.mark.parametrize("causal,Pk,dil", [(False, 3, 1), (False, 3, 16), (True, 3, 4), (True, 2, 2), (False, 5, 2),
                                           (False, 3, 128)])
@pytest.mark.parametrize("M,K,H", [(2, 99, 16), (3, 400, 512), (1, 37, 8)])
def test_dwconv_bwd_with_fused_gln_apply_equals_the_two_pass_schedule(causal, Pk, dil, M, K, H):
    """ctn_dwconv_bwd_gln_fused (norm2 backward + PReLU' applied on load) against ctn_norm_bwd_apply followed by
    ctn_dwconv_bwd — the schedule whose kernels are each checked against the oracle above — on the same inputs."""
    z1, z2 = rnd(M, K, H, seed=31) + 0.1, rnd(M, K, H, seed=32) - 0.05
    dn2 = rnd(M, K, H, seed=33)
    g1, b1, g2, Wd = rnd(H, seed=34), rnd(H, seed=35), rnd(H, seed=36), rnd(H, Pk, seed=37)
    a1, a2 = torch.tensor([0.25], device=dev()), torch.tensor([0.4], device=dev())
    acc1 = gln_acc(FS.prelu(z1.cpu().double(), 0.25)).to(dev())
    acc2 = gln_acc(FS.prelu(z2.cpu().double(), 0.4)).to(dev())
    # the per-sample sums the reduce pass leaves for norm2
    dg2, db2 = torch.zeros(H, device=dev()), torch.zeros(H, device=dev())
    red2 = torch.zeros(M, 2, dtype=torch.float64, device=dev())
    call("ctn_norm_bwd_reduce", P(dn2), P(z2), P(a2), P(acc2), None, P(g2), M, K, H, P(dg2), P(db2), P(red2))

    def outputs():
        return (torch.empty_like(z1), torch.zeros(H, Pk, device=dev()), torch.zeros(H, device=dev()),
                torch.zeros(H, device=dev()), torch.zeros(M, 2, dtype=torch.float64, device=dev()),
                torch.zeros(1, device=dev()))

    # two passes
    dn1_a, dWd_a, dg_a, db_a, red_a, dal_a = outputs()
    dz2 = dn2.clone()
    call("ctn_norm_bwd_apply", P(dz2), P(z2), P(a2), P(acc2), None, P(g2), P(red2), M, K, H, P(dal_a))
    call("ctn_dwconv_bwd", P(dz2), P(z1), P(a1), P(acc1), None, P(g1), P(b1), P(Wd), M, K, H, Pk, dil, int(causal),
         P(dn1_a), P(dWd_a), P(dg_a), P(db_a), P(red_a))
    # fused
    dn1_b, dWd_b, dg_b, db_b, red_b, dal_b = outputs()
    call("ctn_dwconv_bwd_gln_fused", P(dn2), P(z2), P(a2), P(acc2), P(g2), P(red2), P(dal_b), P(z1), P(a1), P(acc1),
         P(g1), P(b1), P(Wd), M, K, H, Pk, dil, int(causal), P(dn1_b), P(dWd_b), P(dg_b), P(db_b), P(red_b))
    assert rel_err(dn1_b.cpu(), dn1_a.cpu()) < 1e-6
    assert rel_err(dWd_b.cpu(), dWd_a.cpu()) < 1e-5
    assert rel_err(dg_b.cpu(), dg_a.cpu()) < 1e-5 and rel_err(db_b.cpu(), db_a.cpu()) < 1e-5
    assert rel_err(red_b.cpu(), red_a.cpu()) < 1e-6
    assert abs(dal_b.item() - dal_a.item()) < 1e-4 * max(1.0, abs(dal_a.item()))


@pytest.mark.parametrize("F", [130, 9597, 51184])
@pytest.mark.parametrize("mode,Kd,Ochan", [(0, 256, 512), (0, 512, 256), (1, 256, 512), (2, 256, 256), (3, 256, 512),
                                           (4, 512, 256), (4, 256, 512), (0, 64, 48)])
def test_conv1x1_planes_every_operand_flavour(F, mode, Kd, Ochan):
    """ctn_conv1x1_planes: pre-split weight planes, every operand flavour of the tcgen05 kernels — 0 bf16x3, 1 TF32x3,
    2..4 the single-bf16 MMA of the reduced-precision inference path (fp32 / bf16-stored activations in, fp32 / bf16
    out) — at one, a few and many frame tiles per SM (the persistent kernel's segment loop), 5 launches each.
    Expected value: the same product in fp64 from the operands as the flavour rounds them."""
    A, W = rnd(F, Kd, seed=41), rnd(Ochan, Kd, seed=42, scale=1 / 16)
    if mode == 1:
        hi = (W.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)  # round to nearest tf32 (ties away)
        lo = W - hi
    else:
        hi = W.to(torch.bfloat16)
        lo = (W - hi.float()).to(torch.bfloat16)
    a_in = A.to(torch.bfloat16) if mode == 4 else A
    D = torch.empty(F, Ochan, device=dev(), dtype=torch.bfloat16 if mode == 3 else torch.float32)
    a_ref = A if mode < 2 else A.to(torch.bfloat16).float()
    w_ref = W if mode < 2 else hi.float()
    want = a_ref.double() @ w_ref.double().t()
    tol = 4e-3 if mode == 3 else TOL_TC  # mode 3 stores bf16: 2^-9 per element
    for _ in range(5):
        D.zero_()
        call("ctn_conv1x1_planes", P(a_in), P(hi), P(lo), mode, P(D), F, Ochan, Kd, 3199)
        assert rel_err(D.float().cpu(), want.cpu()) < tol
