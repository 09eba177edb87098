"""Multi-GPU plumbing (torch.distributed): one process per GPU, requests sharded by contiguous index range.

SURVEY.md 8(e): the path shards with no data-path collective (inputs are independent, engine.go:302-310);
collectives are only (1) the broadcast of the flattened table blob from rank 0 at (re)load time and (2) an
all-gather of the packed decision bitmaps so that every rank holds the whole result.  Backend "nccl" on GPUs
(NVLink / NVSwitch), "gloo" in the CPU tests of this logic.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous index range [lo, hi) of `rank`: GPU g takes [g*N/G, (g+1)*N/G)."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def broadcast_blob(blob: bytes | None, device) -> bytes:
    """Rank 0 passes the table blob, the others None; returns the blob on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return blob
    rank = dist.get_rank()
    ln = torch.tensor([len(blob) if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(ln, 0)
    t = torch.empty(int(ln.item()), dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return t.cpu().numpy().tobytes()


def all_gather_bitmaps(local: torch.Tensor, out: torch.Tensor | None = None, async_op: bool = False):
    """local: uint8[n_local * kbytes] on this rank (equal sizes on all ranks). Returns uint8[world * n_local * kbytes],
    rank-major = request-index order for contiguous shards.  With async_op=True returns (out, work): the collective
    runs on NCCL's own stream and overlaps the next batch's kernel; call work.wait() before reading `out`."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return (local, None) if async_op else local
    world = dist.get_world_size()
    if out is None:
        out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    work = dist.all_gather_into_tensor(out, local, async_op=async_op)
    return (out, work) if async_op else out
