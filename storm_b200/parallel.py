"""Multi-GPU plumbing of the path: one process per GPU, utterances sharded round-robin, ONE broadcast of the
flat weight blob at load, no per-step collective.  The reference has no inference-time
parallelism (enhancement.py:61,66-72 is a single-process loop); its training DDP (train.py:150) is out of scope.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Utterance b goes to rank b mod world."""
    return list(range(rank, n, world))


def padded_frame_count(L: int, hop_length: int = 128) -> int:
    """Frames of a centred STFT, 1 + L // hop (data_module.py:217-219), padded to a multiple of 64 (util/other.py:102-109)."""
    frames = 1 + L // hop_length
    return (frames + 63) // 64 * 64


def bucket_utterances(lengths: Sequence[int], hop_length: int = 128, max_batch: int = 8) -> List[List[int]]:
    """Group utterance indices into batches that reproduce per-file results: a batch only holds utterances with the same
    padded frame count (one [B, F, T_pad] tensor; every other per-utterance quantity - reflect padding, frame count,
    istft length, normalisation - is handled per row by the ragged C-ABI entry points).  Buckets are ordered by frame
    count, indices inside a bucket keep their input order, each batch has at most ``max_batch`` utterances."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    by_T = {}
    for i, L in enumerate(lengths):
        if L < 1:
            raise ValueError(f"utterance {i} is empty")
        by_T.setdefault(padded_frame_count(int(L), hop_length), []).append(i)
    batches: List[List[int]] = []
    for T in sorted(by_T):
        idx = by_T[T]
        batches.extend(idx[k:k + max_batch] for k in range(0, len(idx), max_batch))
    return batches


def flat_params(module: torch.nn.Module) -> torch.Tensor:
    return torch.cat([p.detach().reshape(-1).to(torch.float32) for p in module.parameters()])


def load_flat_params(module: torch.nn.Module, flat: torch.Tensor) -> None:
    off = 0
    with torch.no_grad():
        for p in module.parameters():
            n = p.numel()
            p.copy_(flat[off:off + n].reshape(p.shape).to(p.device))
            off += n
    if hasattr(module, "mark_weights_changed"):
        module.mark_weights_changed()


def broadcast_weights(dnn: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank's backbone weights equal to rank ``src``'s with a single collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    p0 = next(dnn.parameters())
    if p0.is_cuda:
        eng = dnn.engine()                      # uploads this rank's current weights
        dist.broadcast(eng.weights_view(), src=src, group=group)     # NCCL over NVLink, in place in the engine's arena
        torch.cuda.synchronize()
        eng.repack()
        # keep the Python-side parameters consistent with what the engine now holds
        blob = eng.weights_view()
        with torch.no_grad():
            sd = dnn.state_dict()
            for name, shape, off in eng.param_table():
                sd[name].copy_(blob[off:off + sd[name].numel()].reshape(sd[name].shape))
        dnn._dirty = False
    else:
        flat = flat_params(dnn)
        dist.broadcast(flat, src=src, group=group)
        load_flat_params(dnn, flat)


def gather_utterances(local: torch.Tensor, ids: Sequence[int], total: int, dst: int = 0, group=None):
    """Collect per-rank results [B_local, L] on rank ``dst`` in utterance order (a single final gather)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    payload = (list(ids), local.cpu())
    out = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(payload, out, dst=dst, group=group)
    if out is None:
        return None
    L = local.shape[1]
    res = torch.zeros(total, L)
    for ids_r, t in out:
        for i, u in enumerate(ids_r):
            res[u] = t[i]
    return res


def max_over_ranks(value: float, device: str = "cpu", group=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
