"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the flat parameter
layout matches the reference's state_dict, seeded construction reproduces the reference's weights, and the product
refuses to run without CUDA (no fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_model, load_golden
from oracle import conv_tasnet_oracle as O


@pytest.fixture(scope="module")
def lib():
    from conv_tasnet_b200.build import build_library
    build_library()
    from conv_tasnet_b200 import _lib
    return _lib


def test_library_exports_every_header_symbol(lib):
    header = open(os.path.join(ROOT, "include", "ctn_b200.h")).read()
    declared = set(re.findall(r"\b(ctn_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 28
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ctn_b200.h but not exported"
        assert name in lib.SIGNATURES, f"{name} has no ctypes signature"
    assert L.ctn_version() == 100


@pytest.mark.parametrize("causal", [False, True])
def test_param_layout_matches_state_dict(lib, causal):
    cfg = O.Config(N=16, L=8, B=8, H=16, P=3, X=3, R=2, C=2, norm_type="cLN" if causal else "gLN", causal=causal)
    c = lib.make_config(**cfg.as_dict())
    L = lib.lib()
    n = L.ctn_param_tensors(ctypes.byref(c))
    spec = O.param_spec(cfg)
    assert n == len(spec)
    offs, nums = (ctypes.c_int64 * n)(), (ctypes.c_int64 * n)()
    assert L.ctn_param_layout(ctypes.byref(c), offs, nums, n) == 0
    assert [int(x) for x in nums] == [int(np.prod(s)) for _, s in spec]
    assert all(o % 4 == 0 for o in offs) and list(offs) == sorted(offs)
    assert all(offs[i] + nums[i] <= offs[i + 1] for i in range(n - 1))
    assert L.ctn_param_floats(ctypes.byref(c)) >= offs[n - 1] + nums[n - 1]
    # buckets of the staged backward tile the whole buffer
    total, seen = L.ctn_param_floats(ctypes.byref(c)), 0
    for stage in range(cfg.R + 2):
        off, cnt = ctypes.c_int64(), ctypes.c_int64()
        assert L.ctn_grad_bucket(ctypes.byref(c), stage, ctypes.byref(off), ctypes.byref(cnt)) == 0
        seen += cnt.value
    assert seen == total


def test_batchnorm_layout_and_seeded_init(lib):
    """BN branch: weight / bias sit where gamma / beta do; running statistics travel in norm_state; a seeded constructor
    call yields the reference's state_dict (keys, shapes, values; fixture values from tests/golden/model_bn.npz keys)."""
    import torch
    from conv_tasnet_b200 import ConvTasNet
    cfg = O.Config(N=16, L=8, B=8, H=16, P=3, X=3, R=2, C=2, norm_type="BN")
    c = lib.make_config(**cfg.as_dict())
    L = lib.lib()
    assert c.norm_type == 2
    assert L.ctn_norm_state_floats(ctypes.byref(c)) == cfg.R * cfg.X * 4 * cfg.H
    assert L.ctn_norm_state_floats(ctypes.byref(lib.make_config(**O.PAPER.as_dict()))) == 0
    spec = [(k, s) for k, s in O.param_spec(cfg) if not O.is_buffer(k)]
    n = L.ctn_param_tensors(ctypes.byref(c))
    assert n == len(spec)
    offs, nums = (ctypes.c_int64 * n)(), (ctypes.c_int64 * n)()
    assert L.ctn_param_layout(ctypes.byref(c), offs, nums, n) == 0
    assert [int(x) for x in nums] == [int(np.prod(s)) for _, s in spec]
    torch.manual_seed(0)
    m = ConvTasNet(**cfg.as_dict())
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == O.param_spec(cfg)
    sd = O.init_state_dict(cfg)
    for k, v in m.state_dict().items():  # BatchNorm1d defaults: weight 1, bias 0, running (0, 1), counter 0
        if k.split(".")[-1] in ("bias", "running_mean", "running_var", "num_batches_tracked") or \
                (k.endswith("weight") and v.dim() == 1 and v.numel() > 1):
            assert torch.equal(v, sd[k]), k


def test_geometry_and_validation(lib):
    L = lib.lib()
    c = lib.make_config(**O.PAPER.as_dict())
    assert L.ctn_num_frames(ctypes.byref(c), 32000) == 3199 and L.ctn_num_frames(ctypes.byref(c), 480000) == 47999
    inf, tr = L.ctn_workspace_bytes(ctypes.byref(c), 3, 32000, 0), L.ctn_workspace_bytes(ctypes.byref(c), 3, 32000, 1)
    assert 0 < inf < tr < 4 << 30
    assert L.ctn_param_floats(ctypes.byref(c)) >= 8710720
    bad = lib.make_config(**dict(O.PAPER.as_dict(), N=250))
    assert L.ctn_param_floats(ctypes.byref(bad)) == -1 and b"multiples of 4" in L.ctn_last_error()
    bad = lib.make_config(**dict(O.PAPER.as_dict(), P=4))
    assert L.ctn_workspace_bytes(ctypes.byref(bad), 1, 32000, 0) == -1 and b"odd P" in L.ctn_last_error()


def test_seeded_constructor_reproduces_reference_weights():
    from conv_tasnet_b200 import ConvTasNet
    z = load_golden("paper_cfg1.npz")
    torch.manual_seed(int(z["seed_w"]))
    model = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 2)
    assert sum(p.numel() for p in model.parameters()) == 8710720
    for (k, p), name, ws, wa, w0 in zip(model.named_parameters(), z["names"], z["w_sum"], z["w_abs"], z["w_first"]):
        assert k == str(name)
        assert abs(p.detach().double().sum().item() - ws) <= 1e-12 * max(1.0, abs(ws)), k
        assert abs(p.detach().double().abs().sum().item() - wa) <= 1e-12 * max(1.0, wa), k
        assert p.detach().flatten()[0].item() == w0, k


@pytest.mark.parametrize("name", ["gln", "cln_causal"])
def test_state_dict_contract_and_flat_views(name):
    from conv_tasnet_b200 import ConvTasNet
    cfgd, sd, z = golden_model(name)
    model = ConvTasNet(**cfgd)
    assert list(model.state_dict().keys()) == list(sd.keys())
    assert all(model.state_dict()[k].shape == sd[k].shape for k in sd)
    model.load_state_dict(sd)
    flat = model.flat_params
    offs, nums, total = model._param_layout()
    assert flat.numel() == total
    for p, o, n in zip(model.parameters(), offs, nums):
        assert p.data_ptr() == flat.data_ptr() + 4 * o
    model.load_state_dict(sd)  # in-place copy keeps the views
    assert model._flat_ok()
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k])
    for a in ("N", "L", "B", "H", "P", "X", "R", "C", "norm_type", "causal", "mask_nonlinear"):
        assert getattr(model, a) == cfgd[a]


def test_no_cpu_fallback():
    from conv_tasnet_b200 import ConvTasNet, cal_loss, overlap_and_add
    cfgd, sd, z = golden_model("gln")
    model = ConvTasNet(**cfgd)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 403))
    with pytest.raises(RuntimeError, match="CUDA"):
        cal_loss(torch.zeros(1, 2, 8), torch.zeros(1, 2, 8), torch.tensor([8]))
    with pytest.raises(RuntimeError, match="CUDA"):
        overlap_and_add(torch.zeros(1, 3, 4), 2)
    bn = ConvTasNet(**dict(cfgd, norm_type="BN"))  # the BatchNorm fall-through of chose_norm (src/conv_tasnet.py:306)
    assert bn._cfg.norm_type == 2 and len(bn._bns) == 2 * cfgd["R"] * cfgd["X"]


def test_remove_pad_matches_reference_semantics():
    from conv_tasnet_b200 import remove_pad
    x = torch.arange(2 * 3 * 5, dtype=torch.float32).view(2, 3, 5)
    out = remove_pad(x, torch.tensor([5, 2]))
    assert out[0].shape == (3, 5) and out[1].shape == (3, 2) and np.array_equal(out[1], x[1, :, :2].numpy())
    out = remove_pad(x[:, 0], torch.tensor([4, 1]))
    assert out[0].shape == (4,) and out[1].shape == (1,)


def test_oracle_sisnri_matches_reference_goldens():
    """oracle.cal_SISNRi_np / cal_SISNR_np (the CPU checker of the ctn_sisnri kernel) against tests/golden/eval.npz, which
    tests/golden/make_golden_eval.py produced by running the reference's own src/evaluate.py:94-130 per utterance after
    remove_pad, like its evaluation loop."""
    from conftest import load_golden
    z = load_golden("eval.npz")
    for i in range(int(z["n_cases"])):
        src, est, mix, lens = z[f"c{i}_src"], z[f"c{i}_est"], z[f"c{i}_mix"], z[f"c{i}_lengths"]
        for b, n in enumerate(lens.tolist()):
            got = O.cal_SISNRi_np(src[b, :, :n], est[b, :, :n], mix[b, :n])
            assert abs(got - z[f"c{i}_sisnri"][b]) < 1e-4  # the reference works in float32 numpy
            for c in range(2):
                assert abs(O.cal_SISNR_np(src[b, c, :n], est[b, c, :n]) - z[f"c{i}_sisnr"][b, c]) < 1e-4



def test_dispatcher_ops_are_registered_with_fake_kernels():
    """torch.library registration (north star: "thin C-ABI torch.library extension"): the ops exist, are CUDA-only, and
    their fake kernels infer shapes without running anything (what torch.compile / torch.export need)."""
    import conv_tasnet_b200  # noqa: F401  (importing the package registers the ops)
    from conv_tasnet_b200 import ops
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in ("model_forward", "model_backward", "pit_forward", "pit_backward"):
        assert hasattr(torch.ops.ctn_b200, name)
    assert "Tensor(a1!) estimate_source" in str(torch.ops.ctn_b200.pit_forward.default._schema)  # in-place mask declared
    cfg = [16, 8, 8, 16, 3, 2, 1, 2, 0, 0, 0]
    with FakeTensorMode():
        fp = torch.empty(ops.param_floats(cfg), device="cuda")
        mix = torch.empty(2, 400, device="cuda")
        est, ws = torch.ops.ctn_b200.model_forward(fp, mix, cfg, True)
        assert est.shape == (2, 2, 400) and est.device.type == "cuda"
        assert ws.numel() == ops.workspace_bytes(cfg, 2, 400, True) and ws.dtype == torch.uint8
        grads = torch.ops.ctn_b200.model_backward(fp, mix, est, ws, cfg)
        assert grads.shape == fp.shape
        src, lens = torch.empty(2, 2, 400, device="cuda"), torch.empty(2, dtype=torch.int64, device="cuda")
        loss, max_snr, idx, reorder, coef = torch.ops.ctn_b200.pit_forward(src, est, lens)
        assert loss.shape == (1,) and max_snr.shape == (2, 1) and idx.dtype == torch.int64 and reorder.shape == est.shape
        assert torch.ops.ctn_b200.pit_backward(src, est, lens, coef, loss).shape == est.shape
    with pytest.raises(Exception):  # no CPU kernel: the op refuses CPU tensors instead of falling back
        torch.ops.ctn_b200.model_forward(torch.zeros(ops.param_floats(cfg)), torch.zeros(2, 400), cfg, False)


def test_peer_all_reduce_rejects_bad_arguments_before_touching_the_gpu(lib):
    """ctn_peer_all_reduce (the one-kernel gradient exchange) validates its arguments on the host: world size, rank,
    4-float granularity, unmapped buffers — every error is reported through ctn_last_error, nothing is launched"""
    L = lib.lib()
    bufs = (ctypes.c_void_p * 8)(*[0x1000 * (i + 1) for i in range(8)])
    flags = (ctypes.c_void_p * 8)(*[0x100000 + 0x100 * i for i in range(8)])
    assert L.ctn_peer_all_reduce(bufs, flags, 0, 1, 0, 1024, 1.0, None) != 0 and b"world size" in L.ctn_last_error()
    assert L.ctn_peer_all_reduce(bufs, flags, 0, 9, 0, 1024, 1.0, None) != 0 and b"world size" in L.ctn_last_error()
    assert L.ctn_peer_all_reduce(bufs, flags, 2, 2, 0, 1024, 1.0, None) != 0 and b"rank" in L.ctn_last_error()
    assert L.ctn_peer_all_reduce(bufs, flags, 0, 2, 2, 1024, 1.0, None) != 0 and b"multiples of 4" in L.ctn_last_error()
    assert L.ctn_peer_all_reduce(bufs, flags, 0, 2, 0, 1022, 1.0, None) != 0 and b"multiples of 4" in L.ctn_last_error()
    holes = (ctypes.c_void_p * 8)(0x1000, None)
    assert L.ctn_peer_all_reduce(holes, flags, 0, 2, 0, 1024, 1.0, None) != 0 and b"not mapped" in L.ctn_last_error()
    odd = (ctypes.c_void_p * 8)(0x1000, 0x2004)
    assert L.ctn_peer_all_reduce(odd, flags, 0, 2, 0, 1024, 1.0, None) != 0 and b"16-byte aligned" in L.ctn_last_error()
    assert L.ctn_peer_export(None, ctypes.create_string_buffer(64)) != 0
    assert L.ctn_peer_alloc(0, ctypes.byref(ctypes.c_void_p())) != 0
