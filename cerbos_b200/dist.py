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


class PeerGather:
    """Fused all-gather target: one buffer of world * slice_bytes per rank, allocated by the library, exported over
    CUDA IPC and mapped by every other rank (NVLink / NVSwitch peer memory).  The check kernels of rank r store every
    result byte into slice r of ALL buffers (cgpu_check_device_gather), then release a step number into every
    rank's flag array; `wait(step)` makes a stream wait until all slices of that step have landed locally.
    No NCCL call is involved after construction (the handles travel once through all_gather_object)."""

    def __init__(self, ctx, slice_bytes: int, n_buffers: int = 1):
        self.ctx = ctx
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.slice_bytes = int(slice_bytes)
        self.n_buffers = n_buffers
        total = self.world * self.slice_bytes
        self._own = [ctx.peer_alloc(total) for _ in range(n_buffers)]
        # uint32[n_lanes][8]: cell [l][r] = the latest step of lane l that rank r has fully stored here (a lane = one
        # CUDA stream of the issuing rank: steps of one lane complete in order, so each array only counts up)
        self.n_lanes = 4
        self._own_flags = ctx.peer_alloc(4 * 8 * self.n_lanes)
        mine = {"bufs": [h for _, h in self._own], "flags": self._own_flags[1]}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        self._opened = []
        self.bufs = []      # [buffer][rank] -> device pointer valid here
        for j in range(n_buffers):
            row = []
            for r in range(self.world):
                if r == self.rank:
                    row.append(self._own[j][0])
                else:
                    ptr = ctx.peer_open(everyone[r]["bufs"][j])
                    self._opened.append(ptr)
                    row.append(ptr)
            self.bufs.append(row)
        self.flags = []     # [rank] -> that rank's flag array
        for r in range(self.world):
            if r == self.rank:
                self.flags.append(self._own_flags[0])
            else:
                ptr = ctx.peer_open(everyone[r]["flags"])
                self._opened.append(ptr)
                self.flags.append(ptr)
        dist.barrier()

    def lane_flags(self, lane: int):
        """flag arrays (one pointer per rank) of `lane`"""
        return [f + 32 * lane for f in self.flags]

    def local_flags(self, lane: int) -> int:
        return self._own_flags[0] + 32 * lane

    def wait(self, step: int, stream: int = 0, lane: int = 0):
        """Steps of one lane are numbered 1, 2, 3 ... (a rank completes them in order)."""
        self.ctx.gather_wait(self.local_flags(lane), self.world, step, stream)

    def read(self, j: int):
        return self.ctx.peer_read(self._own[j][0], self.world * self.slice_bytes)

    def close(self):
        dist.barrier()
        for p in self._opened:
            self.ctx.peer_close(p)
        dist.barrier()
        for p, _ in self._own:
            self.ctx.peer_free(p)
        self.ctx.peer_free(self._own_flags[0])
