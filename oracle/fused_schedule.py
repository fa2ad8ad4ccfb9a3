"""Kernel-level oracle (TEST INFRASTRUCTURE): the exact fused schedule the CUDA path runs, in torch.

`oracle/conv_tasnet_oracle.py` restates the reference op by op.  This file restates the *same math*
in the form the sm_100a kernels compute it — channels-last `[M,K,Ch]` activations, norm folded into
the following 1x1 conv, hand-derived backward (App. A.5/A.6) — one python function
per C-ABI entry point of `include/ctn_b200.h`.  `tests/test_fused_schedule.py` proves on CPU (fp64)
that this schedule equals the op-by-op oracle (forward, loss, every gradient); the `-m gpu` tests
then compare each CUDA kernel with the function of the same name here.

Only tests/ may import this.  Citations are file:line into /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

from .conv_tasnet_oracle import Config, EPS, is_batch_norm, permutations_of


# ------------------------------------------------------------------ helpers
def prelu(z, a):
    return torch.where(z > 0, z, a * z)


def dprelu(z, a):  # torch: grad passes where z > 0, alpha elsewhere (z == 0 takes the alpha branch)
    return torch.where(z > 0, torch.ones_like(z), a * torch.ones_like(z))


def block_names(cfg: Config, r: int, x: int):
    p = f"separator.network.2.{r}.{x}.net."
    sh = 1 if cfg.causal else 0
    gk, bk = ("weight", "bias") if is_batch_norm(cfg) else ("gamma", "beta")  # nn.BatchNorm1d names (:306-309)
    return dict(W1=p + "0.weight", a1=p + "1.weight", g1=p + "2." + gk, b1=p + "2." + bk, n1=p + "2.",
                Wd=p + "3.net.0.weight", a2=p + f"3.net.{1 + sh}.weight",
                g2=p + f"3.net.{2 + sh}." + gk, b2=p + f"3.net.{2 + sh}." + bk, n2=p + f"3.net.{2 + sh}.",
                W2=p + f"3.net.{3 + sh}.weight")


# ------------------------------------------------------------------ forward kernels
def encoder_fwd(mix, U):
    """mix [M,T], U [N,L] -> w [M,K,N] = relu(frames @ U^T)   (src/conv_tasnet.py:119-120)"""
    L = U.shape[1]
    frames = mix.unfold(1, L, L // 2)  # [M,K,L]
    return torch.relu(frames @ U.t())


def row_stats(a):
    """cLN statistics per frame over channels: (mean, rstd) [M,K,1] (src/conv_tasnet.py:332-334)"""
    mu = a.mean(dim=2, keepdim=True)
    var = ((a - mu) ** 2).mean(dim=2, keepdim=True)
    return mu, 1.0 / torch.sqrt(var + EPS)


def sample_stats(a):
    """gLN statistics per sample over (frames, channels): (mean, rstd) [M,1,1] (src/conv_tasnet.py:358-359)"""
    mu = a.mean(dim=(1, 2), keepdim=True)
    var = ((a - mu) ** 2).mean(dim=(1, 2), keepdim=True)
    return mu, 1.0 / torch.sqrt(var + EPS)


def norm_stats(cfg, a):
    return sample_stats(a) if cfg.norm_type == "gLN" else row_stats(a)


def bn_forward_stats(sd, prefix, a, training, momentum=0.1, eps=1e-5):
    """BatchNorm branch (batchnorm.cu: bn_stats + bn_finalize): per-channel statistics of a [M,K,Ch] over (M,K) in
    training (running statistics updated in place, unbiased variance) or the running statistics in evaluation
    -> (mean, rstd, s, t) with n = s*a + t, s = weight*rstd, t = bias - mean*s."""
    Ch = a.shape[2]
    if training:
        mean = a.mean(dim=(0, 1))
        var = ((a - mean.view(1, 1, Ch)) ** 2).mean(dim=(0, 1))
        n = a.shape[0] * a.shape[1]
        with torch.no_grad():
            sd[prefix + "running_mean"].mul_(1 - momentum).add_(momentum * mean.to(sd[prefix + "running_mean"].dtype))
            sd[prefix + "running_var"].mul_(1 - momentum).add_(
                momentum * (var * (n / max(n - 1, 1))).to(sd[prefix + "running_var"].dtype))
            sd[prefix + "num_batches_tracked"] += 1
    else:
        mean, var = sd[prefix + "running_mean"].to(a.dtype), sd[prefix + "running_var"].to(a.dtype)
    rstd = 1.0 / torch.sqrt(var + eps)
    s = sd[prefix + "weight"] * rstd
    return mean, rstd, s, sd[prefix + "bias"] - mean * s


def bn_bwd(dn, z, alpha, mean, rstd, weight, training):
    """Backward of n = weight*(prelu(z)-mean)*rstd + bias through the batch statistics (bn_bwd_finalize + bn_bwd_apply):
    A = sum dn, B = sum dn*p  ->  dweight = rstd (B - mean A), dbias = A,
    dp = s (dn - A/F - (p - mean) rstd dweight / F) in training, s dn in evaluation.  -> dz, dweight, dbias, dalpha"""
    Ch = z.shape[2]
    p = prelu(z, alpha)
    F_ = z.shape[0] * z.shape[1]
    A = dn.sum(dim=(0, 1))
    Bv = (dn * p).sum(dim=(0, 1))
    dweight = rstd * (Bv - mean * A)
    s = (weight * rstd).view(1, 1, Ch)
    if training:
        dp = s * (dn - (A / F_).view(1, 1, Ch) - (p - mean.view(1, 1, Ch)) * (rstd * dweight / F_).view(1, 1, Ch))
    else:
        dp = s * dn
    dz = dp * dprelu(z, alpha)
    dalpha = (dp * torch.where(z > 0, torch.zeros_like(z), z)).sum().view(1)
    return dz, dweight, A, dalpha


def gemm_normfold(a, mu, r, W, gamma, beta, res=None):
    """(gamma*(a-mu)*r+beta) @ W^T (+res) computed as r*(a @ (W*gamma)^T) + W@beta - mu*r*((W*gamma)@1)
    (App. A.4): the GEMM consumes the raw activation; norm is a rank-1 epilogue."""
    Wg = W * gamma.view(1, -1)
    c1 = W @ beta.view(-1)
    c2 = Wg.sum(dim=1)
    out = r * (a @ Wg.t()) + c1 - mu * r * c2
    return out if res is None else out + res


def dwconv_fwd(cfg, z1, a1, mu1, r1, g1, b1, Wd, d):
    """z2[m,k,h] = sum_p Wd[h,p] * n1[m, k+(p-c)*d, h], n1 = g1*(prelu(z1)-mu1)*r1+b1 inside [0,K), 0 outside
    (zero padding is applied after the norm; Chomp1d keeps the first K outputs: src/conv_tasnet.py:253-256,295)."""
    M, K, H = z1.shape
    P = Wd.shape[1]
    n1 = g1.view(1, 1, H) * (prelu(z1, a1) - mu1) * r1 + b1.view(1, 1, H)
    c = (P - 1) if cfg.causal else (P - 1) // 2
    z2 = torch.zeros_like(z1)
    for p in range(P):
        off = (p - c) * d
        lo, hi = max(0, -off), min(K, K - off)
        if hi > lo:
            z2[:, lo:hi] += Wd[:, p].view(1, 1, H) * n1[:, lo + off:hi + off]
    return z2


def decoder_fwd(cfg, score, w, V, T):
    """score [M,K,C*N] (channel c*N+n), w [M,K,N], V [L,N] -> est [M,C,T]
    mask nonlinearity, mask*w, basis, overlap-add with step L//2, right zero pad (src/conv_tasnet.py:208-214,140-145,57-59)."""
    M, K, N = w.shape
    C, L = cfg.C, V.shape[0]
    S = L // 2
    sc = score.view(M, K, C, N)
    if cfg.mask_nonlinear == "softmax":
        mask = torch.softmax(sc, dim=2)
    elif cfg.mask_nonlinear == "relu":
        mask = torch.relu(sc)
    else:
        raise ValueError("Unsupported mask non-linear function")
    frames = (mask * w.unsqueeze(2)) @ V.t()  # [M,K,C,L]
    est = w.new_zeros(M, C, T)
    for k in range(K):
        est[:, :, k * S:k * S + L] += frames[:, k]
    return est


def model_fwd(cfg: Config, sd: Dict[str, torch.Tensor], mix, keep=True, training=True):
    """Whole forward in kernel order.  Returns est [M,C,T] and the stash the backward consumes.
    BatchNorm branch: identity statistics (mu 0, r 1) with the per-channel (s, t) in place of (gamma, beta)."""
    T = mix.shape[1]
    st = {}
    bn = is_batch_norm(cfg)
    zero, one = mix.new_zeros(1, 1, 1), mix.new_ones(1, 1, 1)
    U = sd["encoder.conv1d_U.weight"][:, 0, :]
    w = encoder_fwd(mix, U)
    mu0, r0 = row_stats(w)
    x = gemm_normfold(w, mu0, r0, sd["separator.network.1.weight"][:, :, 0],
                      sd["separator.network.0.gamma"].view(-1), sd["separator.network.0.beta"].view(-1))
    st.update(w=w, mu0=mu0, r0=r0, blocks=[])
    for r in range(cfg.R):
        for xi in range(cfg.X):
            nm = block_names(cfg, r, xi)
            d = 2 ** xi
            z1 = x @ sd[nm["W1"]][:, :, 0].t()
            bn1 = bn2 = None
            if bn:
                bn1 = bn_forward_stats(sd, nm["n1"], prelu(z1, sd[nm["a1"]]), training)
                mu1, r1, ga1, be1 = zero, one, bn1[2], bn1[3]
            else:
                mu1, r1 = norm_stats(cfg, prelu(z1, sd[nm["a1"]]))
                ga1, be1 = sd[nm["g1"]].view(-1), sd[nm["b1"]].view(-1)
            z2 = dwconv_fwd(cfg, z1, sd[nm["a1"]], mu1, r1, ga1, be1, sd[nm["Wd"]][:, 0, :], d)
            a2 = prelu(z2, sd[nm["a2"]])
            if bn:
                bn2 = bn_forward_stats(sd, nm["n2"], a2, training)
                mu2, r2, ga2, be2 = zero, one, bn2[2], bn2[3]
            else:
                mu2, r2 = norm_stats(cfg, a2)
                ga2, be2 = sd[nm["g2"]].view(-1), sd[nm["b2"]].view(-1)
            out = gemm_normfold(a2, mu2, r2, sd[nm["W2"]][:, :, 0], ga2, be2, res=x)
            st["blocks"].append(dict(x=x, z1=z1, z2=z2, mu1=mu1, r1=r1, mu2=mu2, r2=r2, d=d, nm=nm, bn1=bn1, bn2=bn2,
                                     ga1=ga1, be1=be1, ga2=ga2, be2=be2, training=training))
            x = out
    score = x @ sd["separator.network.3.weight"][:, :, 0].t()
    st.update(y=x, score=score)
    est = decoder_fwd(cfg, score, w, sd["decoder.basis_signals.weight"], T)
    return est, st


# ------------------------------------------------------------------ PIT SI-SNR (moments form, App. A.6)
def pit_fwd(src, est, lengths):
    """One streaming pass of masked moments -> pairwise SI-SNR -> argmax over C! permutations.
    Returns loss, max_snr [B,1], idx [B], reordered est, masked est, and the per-(b,i) coefficients
    (c_e, c_s, c_0, j) such that d loss / d est[b,i,t] = mask * (c_e*est + c_s*src[b,j] + c_0)."""
    B, C, T = src.shape
    dt = src.dtype
    t = torch.arange(T).view(1, 1, T)
    mask = (t < lengths.view(B, 1, 1)).to(dt)
    n = lengths.view(B, 1).to(dt)
    e = est * mask  # in-place masking of the caller's tensor in the product (src/pit_criterion.py:38)
    Se = e.sum(2)                  # [B,C]
    Ss_all = src.sum(2)            # un-masked target sum (:42)
    sm = src * mask
    Ss = sm.sum(2)
    See = (e * e).sum(2)
    Sss = (sm * sm).sum(2)
    Ses = torch.einsum("bit,bjt->bij", e, sm)
    mue, mus = Se / n, Ss_all / n
    # centred, masked moments
    Eee = See - 2 * mue * Se + n * mue * mue                      # [B,C]   sum (e-mue)^2 over t<len
    Ess = Sss - 2 * mus * Ss + n * mus * mus
    dot = Ses - mue.unsqueeze(2) * Ss.unsqueeze(1) - mus.unsqueeze(1) * Se.unsqueeze(2) \
        + n.unsqueeze(2) * mue.unsqueeze(2) * mus.unsqueeze(1)       # [B,i,j]
    E = Ess.unsqueeze(1) + EPS
    a = dot / E
    Pw = a * a * Ess.unsqueeze(1)
    Q = Eee.unsqueeze(2) - 2 * a * dot + a * a * Ess.unsqueeze(1)
    rho = Pw / (Q + EPS)
    snr = 10 * torch.log10(rho + EPS)                              # [B,i,j]
    perms = permutations_of(C)
    rows = torch.arange(C)
    snr_set = snr[:, rows.unsqueeze(0), perms].sum(2)              # [B,C!]
    idx = torch.argmax(snr_set, dim=1)
    max_snr = snr_set.gather(1, idx.view(B, 1)) / C
    loss = -max_snr.mean()
    sel = perms[idx]                                                # [B,C]: est i pairs with target sel[b,i]
    reorder = torch.gather(e, 1, sel.view(B, C, 1).expand(B, C, T))
    # backward coefficients for the chosen pairs
    bi = torch.arange(B).view(B, 1)
    ii = rows.view(1, C).expand(B, C)
    a_, P_, Q_, rho_ = a[bi, ii, sel], Pw[bi, ii, sel], Q[bi, ii, sel], rho[bi, ii, sel]
    Ess_, E_ = Ess.gather(1, sel), Ess.gather(1, sel) + EPS
    mus_, Ssm_ = mus.gather(1, sel), Ss.gather(1, sel)
    qs = dot[bi, ii, sel] - a_ * Ess_                               # <q, s_bar>
    k0 = (10.0 / math.log(10.0)) / (rho_ + EPS) * (-1.0 / (B * C))
    A_ = k0 * (-P_ * 2.0 / (Q_ + EPS) ** 2)                         # coefficient of q(t)
    B_ = k0 * (2 * a_ * Ess_ / E_ / (Q_ + EPS) + P_ * 2.0 * (qs / E_) / (Q_ + EPS) ** 2)  # coefficient of s_bar(t)
    # G = A_*q + B_*s_bar = A_*e_bar + (B_ - A_*a_)*s_bar ; then through (x*mask - sum/len)*mask
    ce = A_
    cs = B_ - A_ * a_
    sum_sbar = Ssm_ - n * mus_                                      # masked sum of s_bar (e_bar sums to 0)
    c0 = -ce * mue - cs * mus_ - cs * sum_sbar / n
    return dict(loss=loss, max_snr=max_snr, idx=idx, perms=perms, reorder=reorder, est_masked=e,
                ce=ce, cs=cs, c0=c0, sel=sel, mask=mask)


def pit_bwd(src, est_masked, pf, grad_loss=1.0):
    B, C, T = src.shape
    s_sel = torch.gather(src, 1, pf["sel"].view(B, C, 1).expand(B, C, T))
    g = pf["mask"] * (pf["ce"].unsqueeze(2) * est_masked + pf["cs"].unsqueeze(2) * s_sel + pf["c0"].unsqueeze(2))
    return g * grad_loss


# ------------------------------------------------------------------ backward kernels
def decoder_bwd(cfg, d_est, score, w, V):
    """-> d_score [M,K,C*N], d_w [M,K,N], dV [L,N]"""
    M, K, N = w.shape
    C, L = cfg.C, V.shape[0]
    S = L // 2
    d_frames = d_est[:, :, :(K - 1) * S + L].unfold(2, L, S)       # [M,C,K,L]  (gather = OLA transpose)
    d_frames = d_frames.permute(0, 2, 1, 3)                          # [M,K,C,L]
    sc = score.view(M, K, C, N)
    mask = torch.softmax(sc, dim=2) if cfg.mask_nonlinear == "softmax" else torch.relu(sc)
    sw = mask * w.unsqueeze(2)
    d_sw = d_frames @ V                                              # [M,K,C,N]
    dV = torch.einsum("mkcl,mkcn->ln", d_frames, sw)
    d_w = (d_sw * mask).sum(2)
    d_mask = d_sw * w.unsqueeze(2)
    if cfg.mask_nonlinear == "softmax":
        d_sc = mask * (d_mask - (d_mask * mask).sum(2, keepdim=True))
    else:
        d_sc = d_mask * (sc > 0).to(sc.dtype)
    return d_sc.reshape(M, K, C * N), d_w, dV


def norm_bwd(cfg, dn, z, alpha, mu, r, gamma):
    """Backward of n = gamma*(prelu(z)-mu)*r+beta for gLN / cLN (App. A.5).
    -> dz, dgamma [Ch], dbeta [Ch], dalpha [1]"""
    Ch = z.shape[2]
    a = prelu(z, alpha) if alpha is not None else z
    yh = (a - mu) * r
    gh = dn * gamma.view(1, 1, Ch)
    dims = (1, 2) if mu.shape[1] == 1 else (2,)
    da = r * (gh - gh.mean(dim=dims, keepdim=True) - yh * (gh * yh).mean(dim=dims, keepdim=True))
    dgamma = (dn * yh).sum(dim=(0, 1))
    dbeta = dn.sum(dim=(0, 1))
    if alpha is None:
        return da, dgamma, dbeta, None
    dz = da * dprelu(z, alpha)
    dalpha = (da * torch.where(z > 0, torch.zeros_like(z), z)).sum().view(1)
    return dz, dgamma, dbeta, dalpha


def dwconv_bwd(cfg, dz2, z1, a1, mu1, r1, g1, b1, Wd, d):
    """-> dn1 [M,K,H], dWd [H,P]"""
    M, K, H = z1.shape
    P = Wd.shape[1]
    n1 = g1.view(1, 1, H) * (prelu(z1, a1) - mu1) * r1 + b1.view(1, 1, H)
    c = (P - 1) if cfg.causal else (P - 1) // 2
    dn1 = torch.zeros_like(z1)
    dWd = torch.zeros_like(Wd)
    for p in range(P):
        off = (p - c) * d
        lo, hi = max(0, -off), min(K, K - off)
        if hi > lo:
            dn1[:, lo + off:hi + off] += Wd[:, p].view(1, 1, H) * dz2[:, lo:hi]
            dWd[:, p] = (dz2[:, lo:hi] * n1[:, lo + off:hi + off]).sum(dim=(0, 1))
    return dn1, dWd


def encoder_bwd(mix, w, dw, L):
    """dU [N,L] = sum_f relu'(w)*dw * frame  (no input gradient is needed)"""
    frames = mix.unfold(1, L, L // 2)
    dpre = dw * (w > 0).to(w.dtype)
    return torch.einsum("mkn,mkl->nl", dpre, frames)


def model_bwd(cfg: Config, sd, mix, st, d_est):
    """Whole backward in kernel order -> grads by state_dict name."""
    G = {}
    V = sd["decoder.basis_signals.weight"]
    Wm = sd["separator.network.3.weight"][:, :, 0]
    d_score, d_w, dV = decoder_bwd(cfg, d_est, st["score"], st["w"], V)
    G["decoder.basis_signals.weight"] = dV
    G["separator.network.3.weight"] = torch.einsum("mko,mkb->ob", d_score, st["y"]).unsqueeze(2)
    g = d_score @ Wm
    for bi in reversed(range(len(st["blocks"]))):
        b = st["blocks"][bi]
        nm = b["nm"]
        W1, W2 = sd[nm["W1"]][:, :, 0], sd[nm["W2"]][:, :, 0]
        bn = b["bn1"] is not None
        g1, b1, g2, b2 = b["ga1"], b["be1"], b["ga2"], b["be2"]  # (gamma, beta), or BatchNorm's (s, t)
        a1, a2 = sd[nm["a1"]], sd[nm["a2"]]
        H = W1.shape[0]
        gshape = (H,) if bn else (1, H, 1)
        dn2 = g @ W2
        n2 = g2.view(1, 1, H) * (prelu(b["z2"], a2) - b["mu2"]) * b["r2"] + b2.view(1, 1, H)
        G[nm["W2"]] = torch.einsum("mko,mkh->oh", g, n2).unsqueeze(2)
        if bn:
            dz2, dg2, db2, da2 = bn_bwd(dn2, b["z2"], a2, b["bn2"][0], b["bn2"][1], sd[nm["g2"]], b["training"])
        else:
            dz2, dg2, db2, da2 = norm_bwd(cfg, dn2, b["z2"], a2, b["mu2"], b["r2"], g2)
        G[nm["g2"]], G[nm["b2"]], G[nm["a2"]] = dg2.view(gshape), db2.view(gshape), da2
        dn1, dWd = dwconv_bwd(cfg, dz2, b["z1"], a1, b["mu1"], b["r1"], g1, b1, sd[nm["Wd"]][:, 0, :], b["d"])
        G[nm["Wd"]] = dWd.unsqueeze(1)
        if bn:
            dz1, dg1, db1, da1 = bn_bwd(dn1, b["z1"], a1, b["bn1"][0], b["bn1"][1], sd[nm["g1"]], b["training"])
        else:
            dz1, dg1, db1, da1 = norm_bwd(cfg, dn1, b["z1"], a1, b["mu1"], b["r1"], g1)
        G[nm["g1"]], G[nm["b1"]], G[nm["a1"]] = dg1.view(gshape), db1.view(gshape), da1
        G[nm["W1"]] = torch.einsum("mkh,mkb->hb", dz1, b["x"]).unsqueeze(2)
        g = g + dz1 @ W1
    # bottleneck + first cLN
    Wb = sd["separator.network.1.weight"][:, :, 0]
    g0, b0 = sd["separator.network.0.gamma"].view(-1), sd["separator.network.0.beta"].view(-1)
    N = Wb.shape[1]
    n0 = g0.view(1, 1, N) * (st["w"] - st["mu0"]) * st["r0"] + b0.view(1, 1, N)
    G["separator.network.1.weight"] = torch.einsum("mkb,mkn->bn", g, n0).unsqueeze(2)
    dn0 = g @ Wb
    dw_sep, dg0, db0, _ = norm_bwd(cfg, dn0, st["w"], None, st["mu0"], st["r0"], g0)
    G["separator.network.0.gamma"], G["separator.network.0.beta"] = dg0.view(1, N, 1), db0.view(1, N, 1)
    dU = encoder_bwd(mix, st["w"], dw_sep + d_w, cfg.L)
    G["encoder.conv1d_U.weight"] = dU.unsqueeze(1)
    return G


def train_step(cfg, sd, mix, src, lengths, training=True):
    est, st = model_fwd(cfg, sd, mix, training=training)
    pf = pit_fwd(src, est, lengths)
    d_est = pit_bwd(src, pf["est_masked"], pf)
    return pf, est, model_bwd(cfg, sd, mix, st, d_est)
