"""Batch assembly next to the hot path: the oracle against the reference-generated vectors (CPU), and the
device-side assembler / packed remove_pad against the oracle, bit exact (pure data movement)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import batch_oracle as BO


def cases():
    z = load_golden("batch.npz")
    for i in range(int(z["n_cases"])):
        lens = z[f"c{i}_lengths"]
        offs = np.concatenate([[0], np.cumsum(lens)])
        mixtures = [z[f"c{i}_packed_mix"][offs[b]:offs[b + 1]] for b in range(len(lens))]
        sources = [z[f"c{i}_packed_src"][offs[b]:offs[b + 1]] for b in range(len(lens))]
        yield i, z, lens, mixtures, sources


def test_oracle_collate_and_remove_pad_match_reference():
    for i, z, lens, mixtures, sources in cases():
        mix, lengths, src = BO.collate(mixtures, sources)
        assert np.array_equal(mix, z[f"c{i}_mix_pad"]) and np.array_equal(src, z[f"c{i}_src_pad"])
        assert np.array_equal(lengths, lens)
        rp3, rp2 = BO.remove_pad(src, lengths), BO.remove_pad(mix, lengths)
        assert np.array_equal(np.concatenate([r.reshape(-1) for r in rp3]), z[f"c{i}_rp3"])
        assert np.array_equal(np.concatenate([r.reshape(-1) for r in rp2]), z[f"c{i}_rp2"])
        assert np.array_equal(np.array([r.shape for r in rp3]), z[f"c{i}_rp3_shapes"])


@pytest.mark.gpu
def test_device_assembler_and_packed_remove_pad_bit_exact():
    from conv_tasnet_b200.data import DeviceBatchAssembler
    from conv_tasnet_b200.utils import remove_pad
    asm = DeviceBatchAssembler()
    for rounds in range(2):  # the staging buffers are reused (and alternate) across calls
        for i, z, lens, mixtures, sources in cases():
            mix, lengths, src = asm(mixtures, sources)
            assert mix.is_cuda and lengths.dtype == torch.int64
            assert np.array_equal(mix.cpu().numpy(), z[f"c{i}_mix_pad"])
            assert np.array_equal(src.cpu().numpy(), z[f"c{i}_src_pad"])
            assert np.array_equal(lengths.cpu().numpy(), lens)
            mix_only, lengths2, none = asm(mixtures)  # evaluation collate (src/data.py:239-260)
            assert none is None and torch.equal(mix_only, mix) and torch.equal(lengths2, lengths)
            for got, want in zip(remove_pad(src, lengths), BO.remove_pad(z[f"c{i}_src_pad"], lens)):
                assert got.shape == want.shape and np.array_equal(got, want)
            for got, want in zip(remove_pad(mix, lengths.cpu()), BO.remove_pad(z[f"c{i}_mix_pad"], lens)):
                assert got.shape == want.shape and np.array_equal(got, want)
    with pytest.raises(ValueError):
        asm([np.zeros(4, np.float32)], [np.zeros((5, 2), np.float32)])
    with pytest.raises(ValueError):
        asm([])


@pytest.mark.gpu
def test_assembled_batch_feeds_the_training_step():
    """the assembler's triple goes straight into model + cal_loss, like solver.py:181-190 after the .cuda() calls"""
    from conv_tasnet_b200 import ConvTasNet, cal_loss
    from conv_tasnet_b200.data import DeviceBatchAssembler
    rng = np.random.default_rng(0)
    lens = [403, 350, 403]
    sources = [0.05 * rng.standard_normal((n, 2)).astype(np.float32) for n in lens]
    mixtures = [s.sum(1) for s in sources]
    mix, lengths, src = DeviceBatchAssembler()(mixtures, sources)
    torch.manual_seed(0)
    model = ConvTasNet(16, 8, 8, 16, 3, 2, 2, 2).cuda()
    loss, max_snr, est, reord = cal_loss(src, model(mix), lengths)
    loss.backward()
    ref_mix, ref_len, ref_src = BO.collate(mixtures, sources)
    loss2, *_ = cal_loss(torch.from_numpy(ref_src).cuda(), model(torch.from_numpy(ref_mix).cuda()),
                         torch.from_numpy(ref_len).cuda())
    assert loss.item() == loss2.item()
