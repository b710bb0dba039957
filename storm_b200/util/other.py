"""Index helpers of the hot path (sgmse/util/other.py:102-109)."""
import torch


def pad_spec(Y: torch.Tensor) -> torch.Tensor:
    """Zero-pad the frame axis (dim 3) on the right to the next multiple of 64."""
    T = Y.size(3)
    num_pad = (64 - T % 64) % 64
    return torch.nn.functional.pad(Y, (0, num_pad, 0, 0))


def padded_frames(T: int) -> int:
    return (T + 63) // 64 * 64


def _rows(x) -> torch.Tensor:
    import numpy as np
    t = torch.from_numpy(np.ascontiguousarray(x)) if not torch.is_tensor(x) else x
    t = t.to(torch.float32)
    return (t.reshape(1, -1) if t.dim() == 1 else t).cuda()


def si_sdr(s, s_hat):
    """Scale-invariant SDR in dB (sgmse/util/other.py:82-86), computed on the GPU (``storm_si_sdr``, fp64 accumulation).
    1-D inputs -> python float like the reference; [B, L] inputs -> float32 tensor [B] on the device."""
    from .. import _native
    out = _native.si_sdr(_rows(s), _rows(s_hat), eps=0.0)
    one = (s.dim() if torch.is_tensor(s) else getattr(s, "ndim", 1)) == 1
    return float(out[0]) if one else out


def si_sdr_torch(s, s_hat):
    """torch form with the 1e-10 guards and truncation to the common length (sgmse/util/other.py:88-94)."""
    from .. import _native
    n = min(s.shape[-1], s_hat.shape[-1])
    out = _native.si_sdr(_rows(s[..., :n]), _rows(s_hat[..., :n]), eps=1e-10)
    return out[0] if s.dim() == 1 else out
