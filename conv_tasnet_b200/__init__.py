"""conv_tasnet_b200 — B200-native (sm_100a) drop-in for the hot path of OfekCohen1/Conv-TasNet.

Mirrors the reference's module layout for the path it replaces:
    conv_tasnet_b200.conv_tasnet    <->  src/conv_tasnet.py    (ConvTasNet, same ctor / forward / state_dict)
    conv_tasnet_b200.pit_criterion  <->  src/pit_criterion.py  (cal_loss, cal_si_snr_with_pit, reorder_source, get_mask)
    conv_tasnet_b200.utils          <->  src/utils.py          (overlap_and_add, remove_pad)
    conv_tasnet_b200.data_parallel  <->  nn.DataParallel use in src/train.py:83-85 (one process per GPU + NCCL)
    conv_tasnet_b200.optim          <->  the clip + Adam tail of src/solver.py:192-196
    conv_tasnet_b200.data           <->  _collate_fn / pad_list + .cuda() of src/data.py:159-183,322-331 (device-side)
    conv_tasnet_b200.ops            <->  the same forward / backward / PIT loss as torch.library dispatcher ops
                                         (torch.ops.ctn_b200.*: register_fake + register_autograd)
    conv_tasnet_b200.evaluate       <->  cal_SISNRi / cal_SISNR of src/evaluate.py:94-130 (one kernel, batched)
    conv_tasnet_b200.separate       <->  the utterance-sharded inference loop of src/separate.py / src/evaluate.py
All compute goes through libctn_b200.so (hand-written CUDA behind the C ABI of include/ctn_b200.h).
"""
from .conv_tasnet import ConvTasNet  # noqa: F401
from .pit_criterion import cal_loss, cal_si_snr_with_pit, reorder_source, get_mask  # noqa: F401
from .utils import overlap_and_add, remove_pad  # noqa: F401
from . import ops  # noqa: F401,E402  (registers torch.ops.ctn_b200.*)
