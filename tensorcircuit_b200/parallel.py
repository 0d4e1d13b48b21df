"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests).  Two things shard on this path:

* the vmap batch -- independent states, block-partitioned over ranks, no data-path collective,
  one all-gather of the per-element results at the end (this file);
* a state larger than one GPU -- the top log2(G) index bits are the rank id
  (``tensorcircuit_b200.dist``)."""

from __future__ import annotations

from typing import Any, Tuple

import numpy as np


def world() -> Tuple[int, int]:
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # pragma: no cover
        pass
    return 0, 1


def shard_bounds(total: int, rank: int, size: int) -> Tuple[int, int]:
    """Contiguous block partition; the first ``total % size`` ranks get one extra element."""
    base, rem = divmod(total, size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(total: int) -> Tuple[int, int]:
    rank, size = world()
    return shard_bounds(total, rank, size)


def _gather_array(a: np.ndarray, total: int) -> np.ndarray:
    import torch
    import torch.distributed as dist

    rank, size = world()
    per = max(shard_bounds(total, r, size)[1] - shard_bounds(total, r, size)[0] for r in range(size))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    was_complex = np.iscomplexobj(a)
    a = np.ascontiguousarray(a)
    if was_complex:
        a = np.stack([a.real, a.imag], axis=-1)
    pad = np.zeros((per,) + a.shape[1:], dtype=a.dtype)
    pad[: a.shape[0]] = a
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(size)]
    dist.all_gather(outs, t)
    parts = []
    for r in range(size):
        lo, hi = shard_bounds(total, r, size)
        parts.append(outs[r][: hi - lo].cpu().numpy())
    out = np.concatenate(parts, axis=0)
    if was_complex:
        out = out[..., 0] + 1j * out[..., 1]
    return out


def gather_batch(out: Any, total: int) -> Any:
    rank, size = world()
    if size == 1:
        return out
    if isinstance(out, (tuple, list)):
        return type(out)(gather_batch(o, total) for o in out)
    if isinstance(out, dict):
        return {k: gather_batch(v, total) for k, v in out.items()}
    return _gather_array(np.asarray(out), total)
