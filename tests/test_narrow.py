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


def widen_slot(cl, col, base=0, base2=0):
    """what widen_kernel does, in numpy"""
    if cl == NW.SLOT_U64:
        return col.astype(np.uint64)
    if cl == NW.SLOT_U16_ID:
        w = col.astype(np.uint64)
        spec = np.where(w == 0xFFFF, BOX(L.V64_ABSENT), np.where(w == 0xFFFE, BOX(L.V64_ERROR), BOX(L.V64_NULL)))
        sid = np.where(w < 0x8000, w + np.uint64(base), w - np.uint64(0x8000) + np.uint64(base2))
        return np.where(w < 0xFFF0, BOX(L.V64_STRING) | sid, np.where(w >= 0xFFFD, spec, BOX(L.V64_BOOL) | (w == 0xFFFB).astype(np.uint64)))
    if cl == NW.SLOT_U8_NUM:
        spec = np.where(col == 0xFF, BOX(L.V64_ABSENT), np.where(col == 0xFE, BOX(L.V64_ERROR), BOX(L.V64_NULL)))
        return np.where(col >= 0xFD, spec, col.astype(np.float64).view(np.uint64))
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


def roundtrip(batch, n_slots, v2=True):
    nb = NW.narrow_batch(batch, n_slots, v2=v2)
    assert nb is not None
    n = batch.n
    slots = np.asarray(batch.columns[3])
    for v in range(n_slots):
        assert (widen_slot(int(nb.slot_class[v]), nb.slot_cols[v], int(nb.slot_base[v]), int(nb.slot_base2[v])) == slots[v]).all(), v
        assert nb.slot_cols[v].nbytes == n * NW.ELEM_BYTES[int(nb.slot_class[v])]
    hdr0 = np.asarray(batch.columns[0]).reshape(-1, 4)
    hdr1 = np.asarray(batch.columns[1])
    # the four 16-bit header fields: constant ones come from hdr_const, the others from the packed columns, in order
    f16, q = [], 0
    for f in range(4):
        if (nb.hdr_const_mask >> f) & 1:
            f16.append(np.full(n, nb.hdr_const[f], dtype=np.uint32))
        else:
            f16.append(nb.hdr16.reshape(n, -1)[:, q].astype(np.uint32))
            q += 1
    assert q == (nb.hdr16.size // n if n else 0)
    k = f16[0]
    kc = np.where(k == 0xFFFF, L.KIND_NONE, np.where(k & 0x8000, (k & 0x7FFF) | L.KIND_CLASS_CSR_BIT, k)).astype(np.uint32)
    pid = nb.principal_id.astype(np.uint32) + np.uint32(nb.principal_base) if nb.principal_base is not None else nb.principal_id
    assert (kc == hdr0[:, 1]).all() and (pid == hdr0[:, 0]).all()
    for j, col in ((1, 2), (2, 3)):
        s = f16[j]
        assert (np.where(s == 0xFFFF, L.SCOPE_NONE, np.where(s & 0x8000, (s & 0x7FFF) | L.SCOPE_INEXACT_BIT, s)).astype(np.uint32) == hdr0[:, col]).all()
    assert (f16[3] == hdr1["aset"]).all()
    ver = np.asarray(nb.versions) if nb.versions is not None else np.tile(np.array(nb.versions_value, dtype=np.uint8), (n, 1))
    for j, name in ((0, "rv"), (1, "pv")):
        assert (np.where(ver[:, j] == 0xFF, L.NONE16, ver[:, j].astype(np.uint32)).astype(np.uint16) == hdr1[name]).all()
    heap = np.asarray(batch.columns[4])
    if nb.heap_bits == 16:
        w = nb.tables[0].astype(np.uint64)
        sid = (w & np.uint64(0x3FFF)) + np.where(w & 0x4000, np.uint64(nb.heap_base2), np.uint64(nb.heap_base))
        assert (np.where(w & 0x8000, BOX(L.V64_STRING) | sid, w) == heap).all()
    elif nb.heap_u32:
        w = nb.tables[0].astype(np.uint64)
        assert (np.where(w & 0x80000000, BOX(L.V64_STRING) | (w & np.uint64(0x7FFFFFFF)), w) == heap).all()
    return nb


@pytest.mark.parametrize("name,n", [("C2", 5000), ("C3", 4000), ("C5", 300)])
def test_workload_columns_narrow_exactly(name, n):
    w = W.WORKLOADS[name]()
    _, ft, enc = W.build(w)
    b = w.columns(w.fields(n), enc)
    nb1 = roundtrip(b, len(enc.slots), v2=False)
    nb = roundtrip(b, len(enc.slots))
    print(name, "request bytes", nb1.request_bytes() / n, "->", nb.request_bytes() / n, "wire", nb1.wire_bytes() / n, "->", nb.wire_bytes() / n)
    if name == "C2":
        assert nb1.request_bytes() / n == 33 and nb.request_bytes() / n <= 19      # 72 B / request of columns -> 33 -> 19
    if name == "C3":
        assert nb1.heap_u32 and nb1.request_bytes() / n == 62 and nb1.tables[0].nbytes * 2 == np.asarray(b.columns[4]).nbytes   # 132 -> 62, heap halved
        assert nb.heap_bits == 16 and nb.request_bytes() / n <= 34 and nb.tables[0].nbytes * 4 == np.asarray(b.columns[4]).nbytes


def test_goldens_narrow_exactly():
    ft = flatten(store_rule_table(), globals_={"environment": "test"})
    for lenient in (False, True):
        inputs = [inp for _, len_, inp, _ in engine_decisions() if len_ == lenient]
        enc = Encoder(ft.manifest, lenient_scope_search=lenient)
        roundtrip(enc.encode(inputs), len(enc.slots))
        roundtrip(enc.encode(inputs), len(enc.slots), v2=False)


@pytest.mark.parametrize("seed", range(8))
def test_random_batches_narrow_exactly(seed):
    """Random policy sets x random requests (tests/fuzzgen.py: missing attributes, nulls, mixed-type columns, nested lists and
    maps, unknown roles, odd scopes): both narrow forms must widen back to the canonical columns bit for bit -- or decline
    the batch (None) -- whatever classes the columns fall into."""
    import random
    from cerbos_b200.policy.compile import build_rule_table
    from fuzzgen import rand_policies, rand_request
    r = random.Random(4200 + seed)
    ft = flatten(build_rule_table(rand_policies(r)))
    enc = Encoder(ft.manifest)
    b = enc.encode([rand_request(r) for _ in range(400)])
    for v2 in (True, False):
        if NW.narrow_batch(b, len(enc.slots), v2=v2) is not None:
            roundtrip(b, len(enc.slots), v2=v2)


def test_slot_classes():
    f = lambda xs: np.array(xs, dtype=np.float64).view(np.uint64)   # noqa: E731
    assert NW.narrow_slot(f([1.0, 7.0, 0.0, 239.0]))[0] == NW.SLOT_U8_NUM and NW.narrow_slot(f([1.0, 240.0]))[0] == NW.SLOT_F32
    assert NW.narrow_slot(f([1.0, -0.0]))[0] == NW.SLOT_F32 and NW.narrow_slot(f([1.0, 0.5]))[0] == NW.SLOT_F32     # -0.0 / fractions are not small integers
    cl, col, base = NW.narrow_slot(np.array([BOX(L.V64_STRING) | np.uint64(70000), BOX(L.V64_STRING) | np.uint64(70010), BOX(L.V64_BOOL) | np.uint64(1), BOX(L.V64_ABSENT)], dtype=np.uint64))
    assert cl == NW.SLOT_U16_ID and base == (70000, 70000) and list(col) == [0, 10, 0xFFFB, 0xFFFF]
    cl, col, base = NW.narrow_slot(np.array([BOX(L.V64_STRING) | np.uint64(5), BOX(L.V64_STRING) | np.uint64(70000), BOX(L.V64_STRING) | np.uint64(70002)], dtype=np.uint64))
    assert cl == NW.SLOT_U16_ID and base == (5, 70000) and list(col) == [0, 0x8000, 0x8002]                     # two windows: table strings, batch strings
    far = np.array([BOX(L.V64_STRING) | np.uint64(5), BOX(L.V64_STRING) | np.uint64(70000), BOX(L.V64_STRING) | np.uint64(200000)], dtype=np.uint64)
    assert NW.narrow_slot(far)[0] == NW.SLOT_U32_ID                                                           # three clusters: no
    assert NW.narrow_slot(f([1.0, 2.5, -0.0, 1e30]))[0] == NW.SLOT_U64          # 1e30 is not a float32
    assert NW.narrow_slot(f([1.0, 2.5, -0.0, 3e38]))[0] == NW.SLOT_U64
    cl, col, _ = NW.narrow_slot(np.concatenate([f([1.0, 2.5, -0.0, 65536.0]), np.array([BOX(L.V64_ABSENT), BOX(L.V64_NULL), L.V64_CANON_NAN], dtype=np.uint64)]))
    assert cl == NW.SLOT_F32
    cl, col, _ = NW.narrow_slot(np.array([BOX(L.V64_BOOL) | np.uint64(1), BOX(L.V64_BOOL), BOX(L.V64_ERROR)], dtype=np.uint64))
    assert cl == NW.SLOT_U8 and list(col) == [1, 0, 4]
    cl, col, _ = NW.narrow_slot(np.array([BOX(L.V64_STRING) | np.uint64(7), BOX(L.V64_BOOL) | np.uint64(1), BOX(L.V64_NULL)], dtype=np.uint64), v2=False)
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
    assert (t.check_narrow(NW.narrow_batch(b, len(enc.slots), v2=False)) == want).all()    # the first form of the wire format
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
