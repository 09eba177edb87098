"""Narrow wire form (cerbos_b200/narrow.py + cgpu_check_narrow): exactness of the narrowing on the CPU (a Python model of
the widening kernel must give back the canonical columns bit for bit), the full path on the GPU against the oracle."""
import os

import numpy as np
import pytest

from cerbos_b200 import narrow as NW
import workloads as W
from cerbos_b200.encode import Encoder
from cerbos_b200.table import layout as L
from cerbos_b200.table.flatten import flatten
from helpers import engine_decisions, store_rule_table

BOX = lambda tag: np.uint64((L.V64_BOX_BASE | tag) << 48)   # noqa: E731


def widen_slot(cl, col):
    """what widen_kernel does, in numpy"""
    if cl == NW.SLOT_U64:
        return col.astype(np.uint64)
    if cl == NW.SLOT_U8:
        c = col.astype(np.uint64)
        return np.where(c <= 1, BOX(L.V64_BOOL) | c, np.where(c == 2, BOX(L.V64_NULL), np.where(c == 3, BOX(L.V64_ABSENT), BOX(L.V64_ERROR))))
    w = col.astype(np.uint64)
    spec = np.where(w == 0xFFFFFFFF, BOX(L.V64_ABSENT), np.where(w == 0xFFFFFFFE, BOX(L.V64_ERROR), BOX(L.V64_NULL)))
    if cl == NW.SLOT_U32_ID:
        return np.where(w < 0xFFFFFFF0, BOX(L.V64_STRING) | w, np.where(w >= 0xFFFFFFFD, spec, BOX(L.V64_BOOL) | (w == 0xFFFFFFFB).astype(np.uint64)))
    if cl == NW.SLOT_U32_HEAP:
        ref = np.where((w & 0x80000000) != 0, BOX(L.V64_MAP), BOX(L.V64_LIST)) | np.uint64(L.V64_HEAP_BATCH_BIT) | (w & np.uint64(0x7FFFFFFF))
        return np.where(w >= 0xFFFFFFFD, spec, ref)
    f = col.view(np.float32)
    is_spec = ((col & 0x7FC00000) == 0x7FC00000) & ((col & 0x3FFFFF) != 0)
    code = col & 3
    sp = np.where(code == 1, BOX(L.V64_ABSENT), np.where(code == 2, BOX(L.V64_ERROR), BOX(L.V64_NULL)))
    with np.errstate(invalid="ignore"):
        d = f.astype(np.float64).view(np.uint64)
    return np.where(is_spec, sp, np.where(col == 0x7FC00000, np.uint64(L.V64_CANON_NAN), d))


def roundtrip(batch, n_slots):
    nb = NW.narrow_batch(batch, n_slots)
    assert nb is not None
    slots = np.asarray(batch.columns[3])
    for v in range(n_slots):
        assert (widen_slot(int(nb.slot_class[v]), nb.slot_cols[v]) == slots[v]).all(), v
    hdr0 = np.asarray(batch.columns[0]).reshape(-1, 4)
    k = nb.hdr16[:, 0].astype(np.uint32)
    kc = np.where(k == 0xFFFF, L.KIND_NONE, np.where(k & 0x8000, (k & 0x7FFF) | L.KIND_CLASS_CSR_BIT, k)).astype(np.uint32)
    assert (kc == hdr0[:, 1]).all() and (nb.principal_id == hdr0[:, 0]).all()
    for j, col in ((1, 2), (2, 3)):
        s = nb.hdr16[:, j].astype(np.uint32)
        assert (np.where(s == 0xFFFF, L.SCOPE_NONE, np.where(s & 0x8000, (s & 0x7FFF) | L.SCOPE_INEXACT_BIT, s)).astype(np.uint32) == hdr0[:, col]).all()
    heap = np.asarray(batch.columns[4])
    if nb.heap_u32:
        w = nb.tables[0].astype(np.uint64)
        assert (np.where(w & 0x80000000, BOX(L.V64_STRING) | (w & np.uint64(0x7FFFFFFF)), w) == heap).all()
    return nb


@pytest.mark.parametrize("name,n", [("C2", 5000), ("C3", 4000), ("C5", 300)])
def test_workload_columns_narrow_exactly(name, n):
    w = W.WORKLOADS[name]()
    _, ft, enc = W.build(w)
    b = w.columns(w.fields(n), enc)
    nb = roundtrip(b, len(enc.slots))
    if name == "C2":
        assert nb.request_bytes() / n == 33        # 72 B / request of columns -> 33
    if name == "C3":
        assert nb.heap_u32 and nb.request_bytes() / n == 62 and nb.tables[0].nbytes * 2 == np.asarray(b.columns[4]).nbytes   # 132 -> 62, heap halved


def test_goldens_narrow_exactly():
    ft = flatten(store_rule_table(), globals_={"environment": "test"})
    for lenient in (False, True):
        inputs = [inp for _, len_, inp, _ in engine_decisions() if len_ == lenient]
        enc = Encoder(ft.manifest, lenient_scope_search=lenient)
        roundtrip(enc.encode(inputs), len(enc.slots))


def test_slot_classes():
    f = lambda xs: np.array(xs, dtype=np.float64).view(np.uint64)   # noqa: E731
    assert NW.narrow_slot(f([1.0, 2.5, -0.0, 1e30]))[0] == NW.SLOT_U64          # 1e30 is not a float32
    assert NW.narrow_slot(f([1.0, 2.5, -0.0, 3e38]))[0] == NW.SLOT_U64
    cl, col = NW.narrow_slot(np.concatenate([f([1.0, 2.5, -0.0, 65536.0]), np.array([BOX(L.V64_ABSENT), BOX(L.V64_NULL), L.V64_CANON_NAN], dtype=np.uint64)]))
    assert cl == NW.SLOT_F32
    cl, col = NW.narrow_slot(np.array([BOX(L.V64_BOOL) | np.uint64(1), BOX(L.V64_BOOL), BOX(L.V64_ERROR)], dtype=np.uint64))
    assert cl == NW.SLOT_U8 and list(col) == [1, 0, 4]
    cl, col = NW.narrow_slot(np.array([BOX(L.V64_STRING) | np.uint64(7), BOX(L.V64_BOOL) | np.uint64(1), BOX(L.V64_NULL)], dtype=np.uint64))
    assert cl == NW.SLOT_U32_ID
    mixed = np.array([BOX(L.V64_STRING) | np.uint64(7), f([1.0])[0]], dtype=np.uint64)
    assert NW.narrow_slot(mixed)[0] == NW.SLOT_U64


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("C2", (1 << 19) + 13), ("C3", (1 << 18) + 5), ("C5", 2000)])
def test_check_narrow_on_gpu(name, n):
    from cerbos_b200 import capi
    from oracle import cref
    w = W.WORKLOADS[name]()
    _, ft, enc = W.build(w)
    b = w.columns(w.fields(n), enc)
    want = cref.check(ft.blob, b.columns, b.n, b.max_actions, n_threads=os.cpu_count() or 1)
    nb = NW.narrow_batch(b, len(enc.slots))
    c = capi.Context(0)
    t = c.load_table(ft.blob)
    t.wait_ready()
    assert (t.check_narrow(nb) == want).all()
    assert (t.check(b.columns, b.n, b.max_actions) == want).all()
    t.release()
    c.close()


@pytest.mark.gpu
def test_goldens_narrow_on_gpu():
    from cerbos_b200 import capi
    from oracle import cref
    ft = flatten(store_rule_table(), globals_={"environment": "test"})
    inputs = [inp for _, len_, inp, _ in engine_decisions() if not len_]
    enc = Encoder(ft.manifest)
    b = enc.encode(inputs * 30)
    want = cref.check(ft.blob, b.columns, b.n, b.max_actions, 1_704_067_200_000_000_000)
    nb = NW.narrow_batch(b, len(enc.slots))
    c = capi.Context(0)
    t = c.load_table(ft.blob)
    got = t.check_narrow(nb, 1_704_067_200_000_000_000)
    valid = want != 0
    assert (got[valid] == want[valid]).all()
    t.release()
    c.close()
