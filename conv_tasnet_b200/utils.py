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
    (src/utils.py:50-67; one device->host copy for the whole batch instead of one per item)
    """
    host = inputs.detach().cpu().numpy()
    lengths = [int(n) for n in torch.as_tensor(inputs_lengths).tolist()]
    if inputs.dim() == 3:
        return [host[b, :, :n].reshape(host.shape[1], -1).copy() for b, n in enumerate(lengths)]
    if inputs.dim() == 2:
        return [host[b, :n].reshape(-1).copy() for b, n in enumerate(lengths)]
    return []
