"""Index helpers of the hot path (sgmse/util/other.py:102-109)."""
import torch


def pad_spec(Y: torch.Tensor) -> torch.Tensor:
    """Zero-pad the frame axis (dim 3) on the right to the next multiple of 64."""
    T = Y.size(3)
    num_pad = (64 - T % 64) % 64
    return torch.nn.functional.pad(Y, (0, num_pad, 0, 0))


def padded_frames(T: int) -> int:
    return (T + 63) // 64 * 64
