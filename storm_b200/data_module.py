"""STFT front / back end of the path: the ``SpecsDataModule`` methods ``ScoreModel.enhance`` uses
(sgmse/data_module.py:140-223: ctor options, stft, istft, spec_fwd, spec_back, _get_window).
Dataset / DataLoader parts of the reference class are training-only and out of scope (SURVEY.md
section 2, row 9).  stft / istft run as native kernels (direct DFT of the non-power-of-two n_fft = 510)."""
from __future__ import annotations

import torch

from . import _native


def get_window(window_type: str, window_length: int) -> torch.Tensor:
    if window_type == "sqrthann":
        return torch.sqrt(torch.hann_window(window_length, periodic=True))
    if window_type == "hann":
        return torch.hann_window(window_length, periodic=True)
    raise NotImplementedError(f"Window type {window_type} not implemented!")


class SpecsDataModule:
    def __init__(self, base_dir="", format="wsj0", spatial_channels=1, batch_size=8, n_fft=510, hop_length=128,
                 num_frames=256, window="hann", num_workers=8, dummy=False, spec_factor=0.15, spec_abs_exponent=0.5,
                 gpu=True, return_time=False, **kwargs):
        self.n_fft, self.hop_length, self.num_frames = n_fft, hop_length, num_frames
        self.window_type = window
        self.window = get_window(window, n_fft)
        self.spec_factor, self.spec_abs_exponent = spec_factor, spec_abs_exponent
        self.batch_size, self.spatial_channels = batch_size, spatial_channels
        self.kwargs = kwargs

    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--n_fft", type=int, default=510, help="Number of FFT bins. 510 by default.")
        parser.add_argument("--hop_length", type=int, default=128, help="Window hop length. 128 by default.")
        parser.add_argument("--num_frames", type=int, default=256, help="Number of frames for the dataset. 256 by default.")
        parser.add_argument("--window", type=str, choices=("sqrthann", "hann"), default="hann", help="The window function to use for the STFT.")
        parser.add_argument("--spec_factor", type=float, default=0.33, help="Factor to multiply complex STFT coefficients by.")
        parser.add_argument("--spec_abs_exponent", type=float, default=0.5, help="Exponent e for the transformation abs(z)**e * exp(1j*angle(z)).")
        return parser

    def spec_config(self, identity_transform: bool = False) -> _native.SpecConfig:
        if identity_transform:
            return _native.make_spec_config(self.n_fft, self.hop_length, self.window_type, 1.0, 1.0)
        return _native.make_spec_config(self.n_fft, self.hop_length, self.window_type, self.spec_factor, self.spec_abs_exponent)

    # ---- reference-named methods --------------------------------------------------------------------------
    def stft(self, sig: torch.Tensor) -> torch.Tensor:
        """[..., L] float -> [..., n_fft/2+1, 1 + L//hop] complex64 (data_module.py:217-219)."""
        lead = sig.shape[:-1]
        spec = _native.stft(sig.reshape(-1, sig.shape[-1]), None, self.spec_config(identity_transform=True))
        return spec.reshape(*lead, *spec.shape[-2:])

    def istft(self, spec: torch.Tensor, length=None) -> torch.Tensor:
        """Inverse of ``stft`` (data_module.py:221-223); needs an engine handle only for scratch space."""
        lead = spec.shape[:-2]
        F, T = spec.shape[-2:]
        L = length if length is not None else self.hop_length * (T - 1)
        wav = _scratch_engine().istft(spec.reshape(-1, F, T), None, L, self.spec_config(identity_transform=True))
        return wav.reshape(*lead, L)

    def spec_fwd(self, spec: torch.Tensor) -> torch.Tensor:          # data_module.py:182-186
        if self.spec_abs_exponent != 1:
            spec = torch.polar(spec.abs() ** self.spec_abs_exponent, spec.angle())
        return spec * self.spec_factor

    def spec_back(self, spec: torch.Tensor) -> torch.Tensor:         # data_module.py:188-193
        spec = spec / self.spec_factor
        if self.spec_abs_exponent != 1:
            spec = torch.polar(spec.abs() ** (1 / self.spec_abs_exponent), spec.angle())
        return spec


_SCRATCH = None


def _scratch_engine() -> _native.Engine:
    """A minimal handle used only for its workspace pool by the handle-bound iSTFT entry point."""
    global _SCRATCH
    if _SCRATCH is None:
        _SCRATCH = _native.Engine(nf=8, ch_mult=(1,), num_res_blocks=1, attn_resolutions=(0,))
    return _SCRATCH
