"""ctypes binding of libstorm_b200.so (include/storm_b200.h) - the only way Python reaches the CUDA path.

There is no CPU or PyTorch fallback: if the shared library is missing and cannot be built, or no CUDA
device is present when a handle is created, this module raises.  torch is used only to own device memory
and to provide the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("STORM_B200_LIB", os.path.join(_HERE, "libstorm_b200.so"))   # override: A/B of two builds
MAX_LEVELS = 8

E_INVALID, E_CUDA, E_STATE, E_NOMEM = -1, -2, -3, -4
PREDICTORS = {"reverse_diffusion": 0, "euler_maruyama": 1, "none": 2}
CORRECTORS = {"ald": 0, "langevin": 1, "none": 2}


class NetConfig(C.Structure):
    _fields_ = [("nf", C.c_int32), ("num_levels", C.c_int32), ("ch_mult", C.c_int32 * MAX_LEVELS),
                ("num_res_blocks", C.c_int32), ("num_attn_resolutions", C.c_int32),
                ("attn_resolutions", C.c_int32 * MAX_LEVELS), ("image_size", C.c_int32),
                ("input_channels", C.c_int32), ("discriminative", C.c_int32), ("engine", C.c_int32)]


class SamplerConfig(C.Structure):
    _fields_ = [("predictor", C.c_int32), ("corrector", C.c_int32), ("N", C.c_int32), ("corrector_steps", C.c_int32),
                ("snr", C.c_float), ("t_eps", C.c_float), ("theta", C.c_float), ("sigma_min", C.c_float),
                ("sigma_max", C.c_float), ("denoise", C.c_int32), ("probability_flow", C.c_int32)]


class SpecConfig(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("hop_length", C.c_int32), ("sqrt_window", C.c_int32),
                ("spec_factor", C.c_float), ("spec_abs_exponent", C.c_float)]


_lib = None

_PROTOS = {
    "storm_last_error": (C.c_char_p, []),
    "storm_version": (C.c_char_p, []),
    "storm_create": (C.c_int, [C.POINTER(NetConfig), C.POINTER(C.c_void_p)]),
    "storm_destroy": (C.c_int, [C.c_void_p]),
    "storm_param_floats": (C.c_int64, [C.c_void_p]),
    "storm_param_entries": (C.c_int32, [C.c_void_p]),
    "storm_param_info": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int64)]),
    "storm_load_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "storm_weights_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "storm_repack_weights": (C.c_int, [C.c_void_p, C.c_void_p]),
    "storm_reserve": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "storm_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "storm_net_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "storm_set_debug_taps": (C.c_int, [C.c_void_p, C.c_int32]),
    "storm_get_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "storm_pc_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.POINTER(SamplerConfig), C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "storm_pc_num_draws": (C.c_int32, [C.POINTER(SamplerConfig)]),
    "storm_philox_normal": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p]),
    "storm_ode_drift": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.POINTER(SamplerConfig), C.c_void_p]),
    "storm_num_frames": (C.c_int32, [C.c_int32, C.POINTER(SpecConfig)]),
    "storm_padded_frames": (C.c_int32, [C.c_int32]),
    "storm_wave_norm": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "storm_stft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(SpecConfig), C.c_void_p]),
    "storm_istft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                              C.POINTER(SpecConfig), C.c_void_p]),
    "storm_enhance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(SamplerConfig),
                                C.POINTER(SpecConfig), C.c_uint64, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "storm_wave_norm_ragged": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "storm_stft_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.POINTER(SpecConfig), C.c_void_p]),
    "storm_istft_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.POINTER(SpecConfig), C.c_void_p]),
    "storm_enhance_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(SamplerConfig),
                                       C.POINTER(SpecConfig), C.c_uint64, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "storm_si_sdr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "storm_launch_count": (C.c_int64, [C.c_void_p]),
    "storm_reset_launch_count": (C.c_int, [C.c_void_p]),
    "storm_debug_conv": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "storm_debug_conv_timed": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float,
                                         C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]),
    "storm_debug_conv_ex": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.POINTER(C.c_float), C.c_void_p]),
    "storm_debug_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p]),
    "storm_debug_attention": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p]),
    "storm_set_graph": (C.c_int, [C.c_void_p, C.c_int32]),
    "storm_graph_replays": (C.c_int64, [C.c_void_p]),
    "storm_set_profile": (C.c_int, [C.c_void_p, C.c_int32]),
    "storm_get_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


def exported_symbols() -> List[str]:
    return sorted(_PROTOS)


def lib():
    """Load (building if necessary and possible) the native library.  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        try:
            from . import build_native
            build_native.build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"storm_b200: native library {_LIB_PATH} is missing and could not be built ({e}); "
                               "there is no CPU fallback - run `python -c 'import __graft_entry__ as g; g.build()'`") from e
    l = C.CDLL(_LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def _check(rc: int):
    if rc == 0:
        return
    msg = lib().storm_last_error().decode()
    if rc == E_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def _stream(device=None) -> int:
    """torch's current stream on ``device`` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


def _on(t):
    """Context that makes the device of tensor ``t`` (or the given device / index) current: every native call launches on
    the device that owns its operands, whatever the process-wide current device is (``model.cuda(1)`` in a process whose
    current device is 0 must not launch device-0 kernels on device-1 pointers)."""
    dev = t.device if torch.is_tensor(t) else t
    return torch.cuda.device(dev)


def _engine_dev(fn):
    """Engine method runs with the engine's device current (the C side also guards, include/storm_b200.h)."""
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


def _tensor_dev(fn):
    """Handle-free entry point runs on the device of its first tensor argument."""
    @functools.wraps(fn)
    def wrapped(t, *a, **k):
        if torch.is_tensor(t) and t.is_cuda:
            with torch.cuda.device(t.device):
                return fn(t, *a, **k)
        return fn(t, *a, **k)
    return wrapped


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _cf(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"storm_b200: {name} must be a CUDA tensor (the path has no CPU implementation)")
    return t.contiguous()


def make_sampler_config(predictor="reverse_diffusion", corrector="ald", N=30, corrector_steps=1, snr=0.5, t_eps=0.03,
                        theta=1.5, sigma_min=0.05, sigma_max=0.5, denoise=True, probability_flow=False) -> SamplerConfig:
    if predictor not in PREDICTORS:
        raise ValueError(f"Predictor with name '{predictor}' unknown.")
    if corrector not in CORRECTORS:
        raise ValueError(f"Corrector with name '{corrector}' unknown.")
    return SamplerConfig(PREDICTORS[predictor], CORRECTORS[corrector], int(N), int(corrector_steps), float(snr), float(t_eps),
                         float(theta), float(sigma_min), float(sigma_max), int(bool(denoise)), int(bool(probability_flow)))


def make_spec_config(n_fft=510, hop_length=128, window="hann", spec_factor=0.15, spec_abs_exponent=0.5) -> SpecConfig:
    if window not in ("hann", "sqrthann"):
        raise NotImplementedError(f"Window type {window} not implemented!")
    return SpecConfig(int(n_fft), int(hop_length), int(window == "sqrthann"), float(spec_factor), float(spec_abs_exponent))


class Engine:
    """One native handle (one per process / GPU)."""

    def __init__(self, nf=128, ch_mult=(1, 2, 2, 2), num_res_blocks=1, attn_resolutions=(0,), image_size=256,
                 input_channels=4, discriminative=False, engine: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("storm_b200: no CUDA device - the enhancement path runs on B200 only (no CPU fallback)")
        cfg = NetConfig()
        cfg.nf = nf
        cfg.num_levels = len(ch_mult)
        for i, v in enumerate(ch_mult):
            cfg.ch_mult[i] = int(v)
        cfg.num_res_blocks = num_res_blocks
        cfg.num_attn_resolutions = len(attn_resolutions)
        for i, v in enumerate(attn_resolutions):
            cfg.attn_resolutions[i] = int(v)
        cfg.image_size = image_size
        cfg.input_channels = input_channels
        cfg.discriminative = int(discriminative)
        cfg.engine = int(os.environ.get("STORM_B200_ENGINE", engine))
        self.cfg = cfg
        self.in_cplx = 1 if discriminative else input_channels // 2
        self.discriminative = bool(discriminative)
        h = C.c_void_p()
        self.device = torch.device("cuda", torch.cuda.current_device())     # the handle lives on the device current at creation
        _check(lib().storm_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def _own(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if t.is_cuda and t.device != self.device:
            raise RuntimeError(f"storm_b200: {name} is on {t.device} but this engine was created on {self.device}")
        return t

    def close(self):
        if getattr(self, "_h", None):
            lib().storm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- parameters ---------------------------------------------------------------------------------------
    def param_table(self) -> List[Tuple[str, Tuple[int, ...], int]]:
        n = lib().storm_param_entries(self._h)
        out = []
        name = C.create_string_buffer(256)
        shape = (C.c_int64 * 4)()
        ndim = C.c_int32()
        off = C.c_int64()
        for i in range(n):
            _check(lib().storm_param_info(self._h, i, name, 256, shape, C.byref(ndim), C.byref(off)))
            out.append((name.value.decode(), tuple(shape[k] for k in range(ndim.value)), off.value))
        return out

    def param_floats(self) -> int:
        return lib().storm_param_floats(self._h)

    @_engine_dev
    def load_blob(self, blob: torch.Tensor):
        blob = blob.contiguous()
        if blob.dtype != torch.float32:
            raise ValueError("weight blob must be float32")
        _check(lib().storm_load_weights(self._h, blob.data_ptr(), blob.numel(), int(blob.is_cuda), _stream()))
        torch.cuda.current_stream().synchronize()   # blob may be a temporary

    @_engine_dev
    def weights_view(self) -> torch.Tensor:
        """Zero-copy view of the raw device blob (for torch.distributed.broadcast)."""
        p = C.c_void_p()
        n = C.c_int64()
        _check(lib().storm_weights_device(self._h, C.byref(p), C.byref(n)))
        return _DevView(p.value, n.value, self.device).tensor()

    @_engine_dev
    def repack(self):
        _check(lib().storm_repack_weights(self._h, _stream()))

    # ---- network ------------------------------------------------------------------------------------------
    @_engine_dev
    def net_forward(self, x: torch.Tensor, t: Optional[torch.Tensor]) -> torch.Tensor:
        x = self._own(_cf(x, "x"), "x")
        if x.dtype != torch.complex64 or x.dim() != 4 or x.shape[1] != self.in_cplx:
            raise ValueError(f"expected complex64 [B, {self.in_cplx}, F, T], got {x.dtype} {tuple(x.shape)}")
        B, _, F, T = x.shape
        if not self.discriminative:
            if t is None:
                raise ValueError("time_cond is required for a score (non-discriminative) backbone")
            t = _cf(t.to(torch.float32), "time_cond")
            if t.numel() != B:
                raise ValueError("time_cond must have one entry per batch element")
        out = torch.empty((B, 1, F, T), dtype=torch.complex64, device=x.device)
        _check(lib().storm_net_forward(self._h, x.data_ptr(), _ptr(t), out.data_ptr(), B, F, T, _stream()))
        return out

    def set_taps(self, on: bool):
        _check(lib().storm_set_debug_taps(self._h, int(on)))

    @_engine_dev
    def get_tap(self, name: str) -> torch.Tensor:
        shape = (C.c_int64 * 4)()
        _check(lib().storm_get_tap(self._h, name.encode(), None, 0, shape, _stream()))
        out = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
        _check(lib().storm_get_tap(self._h, name.encode(), out.data_ptr(), out.numel(), shape, _stream()))
        return out

    @_engine_dev
    def reserve(self, B, F, T):
        _check(lib().storm_reserve(self._h, B, F, T))

    def workspace_bytes(self) -> int:
        return lib().storm_workspace_bytes(self._h)

    def launch_count(self) -> int:
        return lib().storm_launch_count(self._h)

    def reset_launch_count(self):
        lib().storm_reset_launch_count(self._h)

    def set_graph(self, on: bool):
        """CUDA-graph replay of whole trajectories (default on)."""
        _check(lib().storm_set_graph(self._h, int(on)))

    def graph_replays(self) -> int:
        return lib().storm_graph_replays(self._h)

    def set_profile(self, on: bool):
        _check(lib().storm_set_profile(self._h, int(on)))

    @_engine_dev
    def get_profile(self) -> Dict[str, Dict[str, float]]:
        n = (C.c_int64 * 2)()
        ms = (C.c_double * 2)()
        fl = (C.c_double * 2)()
        _check(lib().storm_get_profile(self._h, n, ms, fl))
        return {k: dict(launches=int(n[i]), ms=float(ms[i]), flops=float(fl[i])) for i, k in enumerate(("tcgen05", "simt"))}

    # ---- sampler ------------------------------------------------------------------------------------------
    @_engine_dev
    def pc_sample(self, Y: torch.Tensor, sc: SamplerConfig, seed: int = 0, utt_ids: Optional[Sequence[int]] = None,
                  noise: Optional[torch.Tensor] = None, cond: Optional[Sequence[torch.Tensor]] = None) -> Tuple[torch.Tensor, int]:
        """Y: SDE mean / prior centre.  cond: conditioning spectrograms stacked after x in the network input
        (default [Y]; StoRM passes [Y_noisy, Y_denoised] with Y = Y_denoised)."""
        Y = self._own(_cf(Y, "Y"), "Y")
        if Y.dtype != torch.complex64 or Y.dim() != 4 or Y.shape[1] != 1:
            raise ValueError(f"expected complex64 [B, 1, F, T], got {Y.dtype} {tuple(Y.shape)}")
        B, _, F, T = Y.shape
        X = torch.empty_like(Y)
        ids = None
        if utt_ids is not None:
            if len(utt_ids) != B:
                raise ValueError("utt_ids must have one entry per batch element")
            ids = (C.c_uint64 * B)(*[int(u) for u in utt_ids])
        if noise is not None:
            noise = _cf(noise, "noise")
            nd = lib().storm_pc_num_draws(C.byref(sc))
            if noise.dtype != torch.complex64 or noise.numel() != nd * Y.numel():
                raise ValueError(f"injected noise must be complex64 with {nd} draws of shape {tuple(Y.shape)}")
        c0 = c1 = None
        if cond is not None:
            cond = [_cf(c, "cond") for c in cond]
            if len(cond) not in (1, 2) or any(c.shape != Y.shape or c.dtype != torch.complex64 for c in cond):
                raise ValueError("cond must be one or two complex64 tensors shaped like Y")
            c0 = cond[0]
            c1 = cond[1] if len(cond) == 2 else None
        nfe = C.c_int32()
        _check(lib().storm_pc_sample(self._h, Y.data_ptr(), _ptr(c0), _ptr(c1), X.data_ptr(), B, F, T, C.byref(sc), C.c_uint64(seed),
                                     C.cast(ids, C.c_void_p) if ids is not None else None, _ptr(noise), C.byref(nfe), _stream()))
        return X, int(nfe.value)

    @_engine_dev
    def ode_drift(self, X: torch.Tensor, Y: torch.Tensor, t: float, sc: SamplerConfig) -> torch.Tensor:
        X, Y = _cf(X, "X"), _cf(Y, "Y")
        B, _, F, T = Y.shape
        out = torch.empty_like(X)
        _check(lib().storm_ode_drift(self._h, X.data_ptr(), Y.data_ptr(), float(t), out.data_ptr(), B, F, T, C.byref(sc), _stream()))
        return out

    # ---- STFT ---------------------------------------------------------------------------------------------
    @_engine_dev
    def istft(self, spec: torch.Tensor, norm: Optional[torch.Tensor], L: int, sp: SpecConfig,
              lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """lengths: optional int32 CUDA tensor [B] (ragged batch of one frame bucket, row stride L)."""
        spec = _cf(spec, "spec")
        B, F, T = spec.shape
        if F != sp.n_fft // 2 + 1:
            raise ValueError("spectrogram does not have n_fft/2+1 bins")
        wav = torch.empty((B, L), dtype=torch.float32, device=spec.device)
        _check(lib().storm_istft_ragged(self._h, spec.data_ptr(), _ptr(norm), B, T, wav.data_ptr(), L, _ptr(_lens(lengths, B, L, sp, wav.device)),
                                        C.byref(sp), _stream()))
        return wav

    @_engine_dev
    def enhance(self, wav: torch.Tensor, sc: SamplerConfig, sp: SpecConfig, seed: int = 0,
                utt_ids: Optional[Sequence[int]] = None, lengths: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, int]:
        wav = self._own(_cf(wav.to(torch.float32), "wav"), "wav")
        B, L = wav.shape
        lengths = _lens(lengths, B, L, sp, wav.device)
        out = torch.empty_like(wav)
        ids = None
        if utt_ids is not None:
            ids = (C.c_uint64 * B)(*[int(u) for u in utt_ids])
        nfe = C.c_int32()
        _check(lib().storm_enhance_ragged(self._h, wav.data_ptr(), out.data_ptr(), B, L, _ptr(lengths), C.byref(sc), C.byref(sp),
                                          C.c_uint64(seed), C.cast(ids, C.c_void_p) if ids is not None else None, C.byref(nfe),
                                          _stream()))
        return out, int(nfe.value)


class _DevView:
    """Expose a raw device range as a torch tensor through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int, device="cuda"):
        self.device = device
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

    def tensor(self) -> torch.Tensor:
        return torch.as_tensor(self, device=self.device)


# ---- handle-free entry points ------------------------------------------------------------------------------
def check_lengths(lengths, B: int, L: int, sp: Optional[SpecConfig] = None) -> List[int]:
    """Host-side validation of the per-utterance lengths of a ragged batch (the library cannot see device values):
    B entries, every L_b in (0, L], longer than the STFT reflect padding, and - when a SpecConfig is given - in the same
    padded-frame bucket as the row stride L (otherwise the batch would not reproduce per-file results)."""
    lh = [int(v) for v in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
    if len(lh) != B:
        raise ValueError(f"lengths has {len(lh)} entries for a batch of {B}")
    for v in lh:
        if v < 1 or v > L:
            raise ValueError(f"utterance length {v} outside (0, {L}]")
        if sp is not None:
            if v <= sp.n_fft // 2:
                raise ValueError(f"utterance of {v} samples is shorter than the STFT reflect padding ({sp.n_fft // 2})")
            if padded_frames(num_frames(v, sp)) != padded_frames(num_frames(L, sp)):
                raise ValueError(f"lengths {v} and {L} fall into different padded-frame buckets; batch them separately")
    return lh


def _lens(lengths, B: int, L: int, sp: Optional[SpecConfig] = None, device="cuda") -> Optional[torch.Tensor]:
    """Validated lengths as the int32 CUDA tensor the ragged entry points take (None = every row is L samples)."""
    if lengths is None:
        return None
    return torch.tensor(check_lengths(lengths, B, L, sp), dtype=torch.int32, device=device)


@_tensor_dev
def wave_norm(wav: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    wav = _cf(wav, "wav")
    B, L = wav.shape
    norm = torch.empty((B,), dtype=torch.float32, device=wav.device)
    _check(lib().storm_wave_norm_ragged(wav.data_ptr(), B, L, _ptr(_lens(lengths, B, L, None, wav.device)), norm.data_ptr(), _stream()))
    return norm


@_tensor_dev
def si_sdr(s: torch.Tensor, s_hat: torch.Tensor, lengths: Optional[torch.Tensor] = None, eps: float = 0.0) -> torch.Tensor:
    """SI-SDR in dB per row of [B, L] CUDA tensors (sgmse/util/other.py:82-94); eps=1e-10 is the si_sdr_torch form."""
    s, s_hat = _cf(s.to(torch.float32), "s"), _cf(s_hat.to(torch.float32), "s_hat")
    if s.shape != s_hat.shape or s.dim() != 2:
        raise ValueError("si_sdr expects two [B, L] tensors of the same shape")
    B, L = s.shape
    out = torch.empty((B,), dtype=torch.float32, device=s.device)
    _check(lib().storm_si_sdr(s.data_ptr(), s_hat.data_ptr(), B, L, _ptr(_lens(lengths, B, L, None, s.device)), float(eps), out.data_ptr(), _stream()))
    return out


def num_frames(L: int, sp: SpecConfig) -> int:
    return lib().storm_num_frames(L, C.byref(sp))


def padded_frames(frames: int) -> int:
    return lib().storm_padded_frames(frames)


@_tensor_dev
def stft(wav: torch.Tensor, norm: Optional[torch.Tensor], sp: SpecConfig, T_pad: Optional[int] = None,
         lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    wav = _cf(wav.to(torch.float32), "wav")
    B, L = wav.shape
    frames = num_frames(L, sp)
    T_pad = frames if T_pad is None else T_pad
    spec = torch.empty((B, sp.n_fft // 2 + 1, T_pad), dtype=torch.complex64, device=wav.device)
    _check(lib().storm_stft_ragged(wav.data_ptr(), _ptr(norm), B, L, _ptr(_lens(lengths, B, L, sp, wav.device)), spec.data_ptr(), T_pad,
                                   C.byref(sp), _stream()))
    return spec


def debug_conv(engine: int, x0, w, x1=None, ss=None, silu=False, x2=None, w2=None, bias=None, res=None, res_up=None, out_scale=1.0,
               timed=False, max_prec=1):
    """Test hook around storm_debug_conv_ex: NHWC fp32 CUDA tensors, w OIHW [Cout, Cin, k, k]; res_up = residual at half
    resolution (added FIR-upsampled); max_prec=0 forces the bf16 x 3 operand split.
    timed=True -> (out, ms, counters)."""
    x0 = _cf(x0, "x0")
    B, H, W, C0 = x0.shape
    Cout = w.shape[0]
    taps = w.shape[2] * w.shape[3]
    out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x0.device)
    cont = lambda t: None if t is None else _cf(t, "arg")   # noqa: E731
    x1, ss, x2, w2, bias, res, res_up, w = cont(x1), cont(ss), cont(x2), cont(w2), cont(bias), cont(res), cont(res_up), cont(w)
    dbg = torch.zeros(16, dtype=torch.int64, device=x0.device) if timed else None
    ms = C.c_float()
    with torch.cuda.device(x0.device):
        _check(lib().storm_debug_conv_ex(engine, x0.data_ptr(), _ptr(x1), C0, 0 if x1 is None else x1.shape[-1], _ptr(ss), int(silu),
                                         taps, w.data_ptr(), _ptr(x2), 0 if x2 is None else x2.shape[-1], _ptr(w2), Cout, _ptr(bias),
                                         _ptr(res), _ptr(res_up), float(out_scale), out.data_ptr(), B, H, W, int(max_prec), _ptr(dbg),
                                         C.byref(ms) if timed else None, _stream()))
    if timed:
        return out, float(ms.value), dbg.cpu().tolist()
    return out


@_tensor_dev
def debug_attention(engine: int, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Test hook: q, k, v [B, N, C] fp32 CUDA; the tcgen05 kernel is handed v channel-major as the engine does."""
    q, k, v = _cf(q, "q"), _cf(k, "k"), _cf(v, "v")
    B, N, Cc = q.shape
    o = torch.empty_like(q)
    vv = v.transpose(1, 2).contiguous() if engine == 0 else v
    _check(lib().storm_debug_attention(engine, q.data_ptr(), k.data_ptr(), vv.data_ptr(), o.data_ptr(), B, N, Cc, _stream()))
    return o


def debug_resample(x: torch.Tensor, ss=None, mode: int = 2, want_h: bool = True, want_x: bool = True):
    """Test hook: FIR resampling pass ([B, H, W, C] fp32 NHWC CUDA; mode 1 = up, 2 = down) -> (FIR(SiLU(x * scale + shift)), FIR(x))."""
    x = _cf(x, "x")
    B, H, W, Cc = x.shape
    OH, OW = (2 * H, 2 * W) if mode == 1 else (H // 2, W // 2)
    ss_t = _cf(ss, "ss") if ss is not None else None
    h = torch.empty((B, OH, OW, Cc), device=x.device) if want_h else None
    xo = torch.empty((B, OH, OW, Cc), device=x.device) if want_x else None
    with torch.cuda.device(x.device):
        _check(lib().storm_debug_resample(x.data_ptr(), ss_t.data_ptr() if ss_t is not None else None, h.data_ptr() if want_h else None,
                                          xo.data_ptr() if want_x else None, B, H, W, Cc, mode, _stream()))
    return h, xo


def philox_normal(seed: int, utt_id: int, draw: int, n: int, device=None) -> torch.Tensor:
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((n,), dtype=torch.complex64, device=device)
    with torch.cuda.device(device):
        _check(lib().storm_philox_normal(C.c_uint64(seed), C.c_uint64(utt_id), C.c_uint32(draw), n, out.data_ptr(), _stream()))
    return out
