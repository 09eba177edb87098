"""Small invocations of the kernels added in round 2's second half, for `compute-sanitizer --tool memcheck`:
cb_spec_uc_global (leaf programs, index-form rows), uc_merge_rows, cb_spec_strpred with single-string leaf programs,
widen_kernel with the second narrow form (u16 ids in two windows, u8 numbers, packed header fields) and widen_heap16_kernel.
Every result is compared with the oracle, so a run is also a parity check."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W  # noqa: E402
from cerbos_b200 import capi, narrow as NW  # noqa: E402
from cerbos_b200.device import DeviceBatch  # noqa: E402
from oracle import cref  # noqa: E402

ctx = capi.Context(0)
for name, n in (("C5", 3000), ("C3", 5000), ("C2", 5000)):
    w = W.WORKLOADS[name]()
    _, ft, enc = W.build(w)
    b = w.columns(w.fields(n), enc)
    want = cref.check(ft.blob, b.columns, b.n, b.max_actions)
    t = ctx.load_table(ft.blob)
    ok, note = t.wait_ready()
    assert ok, note
    db = DeviceBatch(b, "cuda:0")
    db.run(t)
    ctx.sync()
    assert (db.effects() == want).all(), name
    for v2 in (True, False):
        nb = NW.narrow_batch(b, len(enc.slots), v2=v2)
        assert (t.check_narrow(nb) == want).all(), (name, v2)
    assert (t.check(b.columns, b.n, b.max_actions) == want).all(), name
    print(name, "ok", ctx.last_kernel_config())
    t.release()
ctx.close()
print("SANITIZE_RUN_OK")
