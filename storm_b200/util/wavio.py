"""wav I/O around the batch path (the reference uses torchaudio.load / torchaudio.save per file, enhancement.py:66-72).

``load_wav`` / ``save_wav`` go through torchaudio when it has a working backend and through scipy.io.wavfile otherwise
(PCM16/PCM32/float wav), so the CLI also runs in images without an audio backend.  ``Prefetcher`` and ``AsyncWriter``
overlap file reads and writes with GPU compute: file k+1.. are read while bucket k is being enhanced, results are written
by a small thread pool (file I/O and CUDA synchronisation both release the GIL).
"""
from __future__ import annotations

import wave
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Callable, Iterable, Iterator, List, Sequence, Tuple, TypeVar

import numpy as np
import torch

T = TypeVar("T")


def _scipy_load(path: str) -> Tuple[torch.Tensor, int]:
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    if x.ndim == 1:
        x = x[:, None]
    return torch.from_numpy(np.ascontiguousarray(x.T)), int(sr)          # [channels, L] like torchaudio.load


def load_wav(path: str) -> Tuple[torch.Tensor, int]:
    """-> (float32 [channels, L], sample rate), the return convention of torchaudio.load."""
    try:
        import torchaudio
        y, sr = torchaudio.load(path)
        return y.to(torch.float32), int(sr)
    except Exception:  # noqa: BLE001  (no backend / unsupported container): plain wav reader
        return _scipy_load(path)


def save_wav(path: str, x: torch.Tensor, sr: int) -> None:
    """x: float32 [channels, L] (or [L]); written as 32-bit float wav like torchaudio.save of a float tensor."""
    x = x.detach().to("cpu", torch.float32)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    try:
        import torchaudio
        torchaudio.save(path, x, sr)
        return
    except Exception:  # noqa: BLE001
        from scipy.io import wavfile
        wavfile.write(path, sr, np.ascontiguousarray(x.numpy().T))


def wav_length(path: str) -> int:
    """Number of sample frames without decoding the file when the header can be parsed (PCM wav), else by loading."""
    try:
        with wave.open(path, "rb") as w:
            return int(w.getnframes())
    except Exception:  # noqa: BLE001  (float wav, extensible header, ...)
        return int(load_wav(path)[0].shape[1])


class Prefetcher:
    """Iterate ``fn(item)`` over ``items`` with up to ``depth`` results being produced ahead by worker threads."""

    def __init__(self, items: Sequence[T], fn: Callable[[T], object], depth: int = 2, workers: int = 4):
        self._items, self._fn, self._depth = list(items), fn, max(1, depth)
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))

    def __iter__(self) -> Iterator[object]:
        pending: List[Future] = []
        it = iter(self._items)
        try:
            for _ in range(self._depth):
                nxt = next(it, None)
                if nxt is None:
                    break
                pending.append(self._pool.submit(self._fn, nxt))
            while pending:
                res = pending.pop(0).result()
                nxt = next(it, None)
                if nxt is not None:
                    pending.append(self._pool.submit(self._fn, nxt))
                yield res
        finally:
            self._pool.shutdown(wait=True)


class AsyncWriter:
    """Write files from a thread pool; ``close()`` waits for all of them and re-raises the first failure."""

    def __init__(self, workers: int = 4):
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self._futs: List[Future] = []

    def save(self, path: str, x: torch.Tensor, sr: int) -> None:
        self._futs.append(self._pool.submit(save_wav, path, x, sr))

    def close(self) -> None:
        self._pool.shutdown(wait=True)
        for f in self._futs:
            f.result()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def load_many(paths: Iterable[str], workers: int = 4) -> List[Tuple[torch.Tensor, int]]:
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        return list(pool.map(load_wav, list(paths)))
