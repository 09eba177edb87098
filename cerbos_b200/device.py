"""Device-resident batches: request columns as torch CUDA tensors + the packed decision bitmap.

torch is plumbing here (device memory, streams, torch.distributed); the evaluation is
cgpu_check_device (cerbos_b200/csrc/cerbos_b200.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from .encode import Batch


def _to_u8(arr: np.ndarray) -> torch.Tensor:
    a = np.ascontiguousarray(arr)
    return torch.from_numpy(a.view(np.uint8).reshape(-1).copy())


class DeviceBatch:
    """Columns of an encoded Batch uploaded to a CUDA device."""

    def __init__(self, batch: Batch, device="cuda:0"):
        self.n = batch.n
        self.max_actions = batch.max_actions
        self.kbytes = (max(batch.max_actions, 1) + 7) // 8
        self.sizes = [int(np.ascontiguousarray(c).nbytes) for c in batch.columns]
        self.tensors = [_to_u8(c).to(device) for c in batch.columns]
        self.ptrs = [t.data_ptr() for t in self.tensors]
        self.bitmap = torch.zeros(self.n * self.kbytes + 8, dtype=torch.uint8, device=device)

    def nbytes(self):
        return sum(self.sizes)

    def run(self, table, now_ns=0, flags=0, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        table.check_device(self.ptrs, self.sizes, self.n, self.max_actions, self.bitmap.data_ptr(), now_ns, flags, s)

    def prepare(self, table, now_ns=0, flags=0):
        """-> f(stream_handle): launches the check with pre-built arguments (for timing loops / graphs)."""
        return table.prepared_device_call(self.ptrs, self.sizes, self.n, self.max_actions, self.bitmap.data_ptr(),
                                          now_ns, flags)

    def effects(self) -> np.ndarray:
        """uint8[n, K] 1 = ALLOW, 2 = DENY (synchronises)."""
        bm = self.bitmap[: self.n * self.kbytes].cpu().numpy().reshape(self.n, self.kbytes)
        bits = np.unpackbits(bm, axis=1, bitorder="little")[:, : max(self.max_actions, 1)]
        return np.where(bits == 1, 1, 2).astype(np.uint8)
