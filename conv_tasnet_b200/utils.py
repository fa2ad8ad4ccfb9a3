"""Drop-in for src/utils.py: overlap_and_add on the GPU (ctn_overlap_and_add), remove_pad as host glue."""
import torch

from . import _lib


def overlap_and_add(signal, frame_step):
    """[..., frames, frame_length] -> [..., (frames-1)*frame_step + frame_length]  (src/utils.py:9-47).
    Each output sample sums the frames that cover it in ascending frame order (what index_add_ does on CPU)."""
    if not signal.is_cuda:
        raise RuntimeError("conv_tasnet_b200.utils.overlap_and_add runs on CUDA tensors only (no CPU fallback)")
    outer = signal.shape[:-2]
    frames, frame_length = signal.shape[-2:]
    if frame_step > frame_length or frame_step < 1:
        raise ValueError("frame_step must be in [1, frame_length]")
    sig = signal.contiguous().to(torch.float32)
    n_outer = 1
    for s in outer:
        n_outer *= s
    out = torch.empty(*outer, (frames - 1) * frame_step + frame_length, dtype=torch.float32, device=signal.device)
    with torch.cuda.device(signal.device):
        done = 0
        while done < n_outer:  # grid.y limit
            n = min(65535, n_outer - done)
            _lib.check(_lib.lib().ctn_overlap_and_add(
                sig.data_ptr() + done * frames * frame_length * 4, n, frames, frame_length, frame_step,
                out.data_ptr() + done * out.shape[-1] * 4, _lib.stream()))
            done += n
    return out.to(signal.dtype)


def remove_pad(inputs, inputs_lengths):
    """
    Args:
        inputs: torch.Tensor, [B, C, T] or [B, T], B is batch size
        inputs_lengths: torch.Tensor, [B]
    Returns:
        results: a list containing B items, each item is [C, T], T varies
    (src/utils.py:50-67).  CUDA inputs: one kernel packs the valid samples of the whole batch (ctn_pack_valid) and ONE
    device->host copy brings exactly those back, instead of one slice + copy per item; CPU inputs are sliced in numpy.
    """
    if inputs.dim() not in (2, 3):
        return []
    T = inputs.shape[-1]
    lengths = [min(max(int(n), 0), T) for n in torch.as_tensor(inputs_lengths).tolist()][:inputs.shape[0]]
    C = inputs.shape[1] if inputs.dim() == 3 else 1
    if inputs.is_cuda and inputs.dtype == torch.float32 and len(lengths) > 0 and sum(lengths) > 0:
        B = len(lengths)
        offs = [0]
        for n in lengths:
            offs.append(offs[-1] + n)
        with torch.cuda.device(inputs.device):
            meta = torch.tensor(lengths + offs[:-1], dtype=torch.int64).pin_memory().to(inputs.device, non_blocking=True)
            packed = torch.empty(offs[-1] * C, dtype=torch.float32, device=inputs.device)
            _lib.check(_lib.lib().ctn_pack_valid(_lib.ptr(inputs.detach().contiguous()), meta.data_ptr(),
                                                 meta.data_ptr() + 8 * B, B, C, T, _lib.ptr(packed), _lib.stream()))
            host = torch.empty(packed.shape, dtype=torch.float32).pin_memory()
            host.copy_(packed, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        flat = host.numpy()
        out = []
        for b, n in enumerate(lengths):
            item = flat[offs[b] * C:offs[b + 1] * C].copy()
            out.append(item.reshape(C, n) if inputs.dim() == 3 else item)
        return out
    host = inputs.detach().cpu().numpy()
    if inputs.dim() == 3:
        return [host[b, :, :n].reshape(host.shape[1], -1).copy() for b, n in enumerate(lengths)]
    return [host[b, :n].reshape(-1).copy() for b, n in enumerate(lengths)]
