"""cgpu_check_narrow over pinned host buffers for several pipeline chunk sizes (CERBOS_B200_CHECK_CHUNK, read per call):
how large must a chunk be for its per-column copies to use the PCIe link well?  One workload batch, results verified once."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W  # noqa: E402
from cerbos_b200 import capi, narrow as NW  # noqa: E402
from oracle import cref  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 23
w = W.WORKLOADS[name]()
_, ft, enc = W.build(w)
b = W.columns_parallel(w, n, 0, enc)
ctx = capi.Context(0)
t = ctx.load_table(ft.blob)
t.wait_ready()
nb = NW.narrow_batch(b, len(enc.slots))


def pin(a):
    a = np.ascontiguousarray(a)
    p = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory()
    p.numpy()[: a.nbytes] = a.view(np.uint8).reshape(-1)
    return p.data_ptr(), p


bb, nr, keep = t.prepare_narrow(nb, 0, 0, pin=pin)
out = torch.empty(n * b.max_actions, dtype=torch.uint8).pin_memory()
t.check_narrow_into(bb, nr, out.data_ptr())
cnt = min(n, 1 << 20)
cols = list(b.columns)
cols[0] = np.ascontiguousarray(b.columns[0][:cnt]); cols[1] = np.ascontiguousarray(b.columns[1][:cnt])
cols[2] = np.ascontiguousarray(b.columns[2][:, :cnt]); cols[3] = np.ascontiguousarray(b.columns[3][:, :cnt])
want = cref.check(ft.blob, cols, cnt, b.max_actions, 0, 0, n_threads=os.cpu_count() or 1)
assert (out.numpy()[: cnt * b.max_actions].reshape(cnt, b.max_actions) == want).all()
print(f"{name} n={n} wire {nb.wire_bytes() / n:.1f} B/request")
for lg in (17, 18, 19, 20, 21, 22):
    os.environ["CERBOS_B200_CHECK_CHUNK"] = str(1 << lg)
    for _ in range(2):
        t.check_narrow_into(bb, nr, out.data_ptr())
    t0 = time.perf_counter()
    reps = 6
    for _ in range(reps):
        t.check_narrow_into(bb, nr, out.data_ptr())
    dt = (time.perf_counter() - t0) / reps
    ok = (out.numpy()[: cnt * b.max_actions].reshape(cnt, b.max_actions) == want).all()
    print(f"chunk 2^{lg}: {dt * 1e3:8.3f} ms  {n * b.max_actions / dt:.3e} decisions/s  {nb.wire_bytes() / dt / 1e9:.1f} GB/s  ok={ok}")
