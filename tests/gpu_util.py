"""Thin torch->C-ABI call helpers for the -m gpu tests (every call goes through libctn_b200.so)."""
import ctypes

import torch

from conv_tasnet_b200 import _lib


def dev():
    return torch.device("cuda:0")


def P(t):
    return None if t is None else _lib.ptr(t)


def call(name, *args):
    _lib.check(getattr(_lib.lib(), name)(*args, _lib.stream()))
    torch.cuda.synchronize()


def gln_acc(a):
    """a [M,K,Ch] -> [M,2] float64 (sum, sumsq): what the producing kernel accumulates for gLN"""
    a = a.double()
    return torch.stack([a.sum(dim=(1, 2)), (a * a).sum(dim=(1, 2))], dim=1).contiguous()


def rowstat(mu, r):
    return torch.cat([mu, r], dim=2).float().contiguous()


def stats_args(norm, a):
    """-> (gln_acc tensor or None, rowstat tensor or None, (mu, r) as the oracle wants them)"""
    from oracle import fused_schedule as FS
    if norm == "gLN":
        mu, r = FS.sample_stats(a.double())
        return gln_acc(a), None, (mu.float(), r.float())
    mu, r = FS.row_stats(a.double())
    return None, rowstat(mu, r), (mu.float(), r.float())
