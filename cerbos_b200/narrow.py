"""Narrow wire form of an encoded batch (include/cerbos_b200.h: cgpu_check_narrow).

`cgpu_check` is bound by the PCIe link, so the per-request columns are re-expressed in their narrowest EXACT form before they
cross it: 16-bit dictionary ids in the header, 8-bit versions and roles, and per attribute slot whichever of u32 string id /
u32 heap reference / float32 / u8 holds every value of the column exactly (else the 8-byte value as is); a heap of string
lists travels as 32-bit words.  A widening kernel rebuilds the canonical columns in HBM.  Nothing is approximated: a column
that does not fit a class keeps its wide form, a batch whose ids do not fit 16 bits is not narrowed at all (-> None).
"""
from __future__ import annotations

import numpy as np

from .table import layout as L

SLOT_U64, SLOT_U32_ID, SLOT_U32_HEAP, SLOT_F32, SLOT_U8 = 0, 1, 2, 3, 4
_BOX = np.uint64(0xFFF0) << np.uint64(48)
_PAY = np.uint64((1 << 48) - 1)


def _tags(col: np.ndarray) -> np.ndarray:
    top = (col >> np.uint64(48)).astype(np.uint32)
    return np.where((top & 0xFFF0) == 0xFFF0, top & 0xF, 0)


class NarrowBatch:
    def __init__(self, n, max_actions, role_cols, principal_id, hdr16, versions, roles, slot_class, slot_cols, tables, heap_u32):
        self.n, self.max_actions, self.role_cols = n, max_actions, role_cols
        self.principal_id, self.hdr16, self.versions, self.roles = principal_id, hdr16, versions, roles
        self.slot_class, self.slot_cols, self.tables, self.heap_u32 = slot_class, slot_cols, tables, heap_u32

    def request_bytes(self) -> int:
        """bytes of the per-request columns (what scales with the batch)"""
        return int(self.principal_id.nbytes + self.hdr16.nbytes + self.versions.nbytes + self.roles.nbytes + sum(c.nbytes for c in self.slot_cols))

    def wire_bytes(self) -> int:
        return int(self.principal_id.nbytes + self.hdr16.nbytes + self.versions.nbytes + self.roles.nbytes +
                   sum(c.nbytes for c in self.slot_cols) + sum(np.asarray(t).nbytes for t in self.tables))


def narrow_slot(col: np.ndarray):
    """-> (class, narrow column) of one u64 slot column"""
    tag = _tags(col)
    pay = col & _PAY
    is_abs, is_err, is_null = tag == L.V64_ABSENT, tag == L.V64_ERROR, tag == L.V64_NULL
    special = is_abs | is_err | is_null
    is_bool, is_str = tag == L.V64_BOOL, tag == L.V64_STRING
    if (is_bool | special).all():
        out = np.where(is_bool, pay.astype(np.uint8), np.where(is_null, 2, np.where(is_abs, 3, 4))).astype(np.uint8)
        return SLOT_U8, out
    if (is_str | is_bool | special).all() and (not is_str.any() or int(pay[is_str].max()) < 0xFFFFFFF0):
        out = pay.astype(np.uint32)
        out[is_abs], out[is_err], out[is_null] = 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFD
        out[is_bool] = np.where(pay[is_bool] != 0, 0xFFFFFFFB, 0xFFFFFFFC).astype(np.uint32)
        return SLOT_U32_ID, out
    is_heap = ((tag == L.V64_LIST) | (tag == L.V64_MAP)) & ((col & np.uint64(L.V64_HEAP_BATCH_BIT)) != 0)
    if (is_heap | special).all():
        off = pay & np.uint64(L.V64_HEAP_BATCH_BIT - 1)
        if not is_heap.any() or int(off[is_heap].max()) < 0x7FFFFFF0:
            out = off.astype(np.uint32) | np.where(tag == L.V64_MAP, 0x80000000, 0).astype(np.uint32)
            out[is_abs], out[is_err], out[is_null] = 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFD
            return SLOT_U32_HEAP, out
    is_num = tag == 0
    if (is_num | special).all():
        d = col.view(np.float64)
        with np.errstate(over="ignore", invalid="ignore"):
            f = d.astype(np.float32)
            back = f.astype(np.float64)
        nan = is_num & (col == np.uint64(L.V64_CANON_NAN))
        exact = (back.view(np.uint64) == col) | ~is_num | nan
        if exact.all():
            out = f.view(np.uint32).copy()
            out[nan] = 0x7FC00000
            out[is_abs], out[is_err], out[is_null] = 0x7FC00001, 0x7FC00002, 0x7FC00003
            return SLOT_F32, out
    return SLOT_U64, col


def narrow_batch(batch, n_slots: int):
    """batch: cerbos_b200.encode.Batch.  -> NarrowBatch, or None when an id does not fit its 16 / 8-bit field."""
    hdr0 = np.asarray(batch.columns[0]).reshape(-1, 4)
    hdr1 = np.asarray(batch.columns[1])
    roles = np.asarray(batch.columns[2])
    slots = np.asarray(batch.columns[3])
    n = batch.n
    kc, rs, ps = hdr0[:, 1], hdr0[:, 2], hdr0[:, 3]
    kc_none, kc_csr = kc == L.KIND_NONE, ((kc & L.KIND_CLASS_CSR_BIT) != 0) & (kc != L.KIND_NONE)
    kid = kc & np.uint32(~L.KIND_CLASS_CSR_BIT & 0xFFFFFFFF)
    if (kid[~kc_none] >= 0x7FFF).any():
        return None
    k16 = np.where(kc_none, 0xFFFF, kid | np.where(kc_csr, 0x8000, 0)).astype(np.uint16)

    def scope16(s):
        none = s == L.SCOPE_NONE
        inexact = ((s & L.SCOPE_INEXACT_BIT) != 0) & ~none
        sid = s & np.uint32(~L.SCOPE_INEXACT_BIT & 0xFFFFFFFF)
        if (sid[~none] >= 0x7FFF).any():
            return None
        return np.where(none, 0xFFFF, sid | np.where(inexact, 0x8000, 0)).astype(np.uint16)

    rs16, ps16 = scope16(rs), scope16(ps)
    aset = hdr1["aset"]
    rv, pv = hdr1["rv"], hdr1["pv"]
    if rs16 is None or ps16 is None or (aset > 0xFFFF).any():
        return None
    if ((rv != L.NONE16) & (rv >= 0xFF)).any() or ((pv != L.NONE16) & (pv >= 0xFF)).any():
        return None
    rmask = (roles != L.ROLE_PAD) & (roles != L.ROLE_UNKNOWN)
    if (roles[rmask] >= 0xFE).any():
        return None
    hdr16 = np.ascontiguousarray(np.stack([k16, rs16, ps16, aset.astype(np.uint16)], axis=1))
    versions = np.ascontiguousarray(np.stack([np.where(rv == L.NONE16, 0xFF, rv).astype(np.uint8), np.where(pv == L.NONE16, 0xFF, pv).astype(np.uint8)], axis=1))
    roles8 = np.ascontiguousarray(np.where(roles == L.ROLE_PAD, 0xFF, np.where(roles == L.ROLE_UNKNOWN, 0xFE, roles)).astype(np.uint8))
    classes, cols = [], []
    for v in range(n_slots):
        c, col = narrow_slot(np.ascontiguousarray(slots[v]))
        classes.append(c)
        cols.append(np.ascontiguousarray(col))
    tables = [np.ascontiguousarray(np.asarray(c)) for c in batch.columns[4:]]
    heap = tables[0]
    htag = _tags(heap)
    hpay = heap & _PAY
    small = (heap < np.uint64(1 << 31))
    hstr = (htag == L.V64_STRING) & (hpay < np.uint64(1 << 31))
    heap_u32 = bool((small | hstr).all())
    if heap_u32:
        tables[0] = np.where(hstr, hpay | np.uint64(1 << 31), heap).astype(np.uint32)
    return NarrowBatch(n, batch.max_actions, roles.shape[0], np.ascontiguousarray(hdr0[:, 0]), hdr16, versions, roles8,
                       np.array(classes or [0], dtype=np.uint8), cols, tables, heap_u32)
