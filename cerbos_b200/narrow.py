"""Narrow wire form of an encoded batch (include/cerbos_b200.h: cgpu_check_narrow).

`cgpu_check` is bound by the PCIe link, so the per-request columns are re-expressed in their narrowest EXACT form before they
cross it: 16-bit dictionary ids in the header (a field that has one value in the whole batch does not travel at all), 8-bit
versions and roles, the principal id as a 16-bit offset when the batch's principals sit within 65 536 ids, and per attribute
slot whichever of u8 small number / u8 bool / u16 string id relative to the column's smallest / u32 string id / u32 heap
reference / float32 holds every value of the column exactly (else the 8-byte value as is); a heap of string lists travels as
16- or 32-bit words.  A widening kernel rebuilds the canonical columns in HBM.  Nothing is approximated: a column
that does not fit a class keeps its wide form, a batch whose ids do not fit 16 bits is not narrowed at all (-> None).
"""
from __future__ import annotations

import numpy as np

from .table import layout as L

SLOT_U64, SLOT_U32_ID, SLOT_U32_HEAP, SLOT_F32, SLOT_U8, SLOT_U16_ID, SLOT_U8_NUM = 0, 1, 2, 3, 4, 5, 6
ELEM_BYTES = {SLOT_U64: 8, SLOT_U32_ID: 4, SLOT_U32_HEAP: 4, SLOT_F32: 4, SLOT_U8: 1, SLOT_U16_ID: 2, SLOT_U8_NUM: 1}
_BOX = np.uint64(0xFFF0) << np.uint64(48)
_PAY = np.uint64((1 << 48) - 1)


def _tags(col: np.ndarray) -> np.ndarray:
    top = (col >> np.uint64(48)).astype(np.uint32)
    return np.where((top & 0xFFF0) == 0xFFF0, top & 0xF, 0)


class NarrowBatch:
    def __init__(self, n, max_actions, role_cols, principal_id, hdr16, versions, roles, slot_class, slot_cols, tables, heap_u32,
                 slot_base=None, principal_base=None, hdr_const_mask=0, hdr_const=(0, 0, 0, 0), versions_value=None, heap_bits=0, heap_base=0,
                 slot_base2=None, heap_base2=0):
        self.n, self.max_actions, self.role_cols = n, max_actions, role_cols
        # principal_id: u32 ids, or (principal_base is not None) u16 offsets from principal_base
        # hdr16: u16[n][fields that vary]; versions: u8[n][2], or None when versions_value holds the batch-wide pair
        self.principal_id, self.hdr16, self.versions, self.roles = principal_id, hdr16, versions, roles
        self.slot_class, self.slot_cols, self.tables, self.heap_u32 = slot_class, slot_cols, tables, heap_u32
        self.slot_base = slot_base if slot_base is not None else np.zeros(max(len(slot_cols), 1), dtype=np.uint32)
        self.principal_base, self.hdr_const_mask, self.hdr_const = principal_base, hdr_const_mask, tuple(int(x) for x in hdr_const)
        self.versions_value, self.heap_bits, self.heap_base, self.heap_base2 = versions_value, heap_bits, heap_base, heap_base2
        self.slot_base2 = slot_base2 if slot_base2 is not None else np.zeros_like(self.slot_base)

    def request_bytes(self) -> int:
        """bytes of the per-request columns (what scales with the batch)"""
        return int(self.principal_id.nbytes + self.hdr16.nbytes + (self.versions.nbytes if self.versions is not None else 0) + self.roles.nbytes +
                   sum(c.nbytes for c in self.slot_cols))

    def wire_bytes(self) -> int:
        return self.request_bytes() + int(sum(np.asarray(t).nbytes for t in self.tables))


def two_windows(ids: np.ndarray, width: int):
    """ids (non-empty) -> (base, base2) such that every id lies in [base, base + width) or [base2, base2 + width), or None"""
    lo = int(ids.min())
    rest = ids[ids >= np.uint64(lo + width)]
    lo2 = int(rest.min()) if rest.size else lo
    return (lo, lo2) if not rest.size or int(rest.max()) - lo2 < width else None


def narrow_slot(col: np.ndarray, v2: bool = True):
    """-> (class, narrow column, (base, base2)) of one u64 slot column (the two windows a SLOT_U16_ID column's ids are offsets into)"""
    return _narrow_slot(col, v2)


def _narrow_slot(col: np.ndarray, v2: bool):
    tag = _tags(col)
    pay = col & _PAY
    is_abs, is_err, is_null = tag == L.V64_ABSENT, tag == L.V64_ERROR, tag == L.V64_NULL
    special = is_abs | is_err | is_null
    is_bool, is_str = tag == L.V64_BOOL, tag == L.V64_STRING
    if (is_bool | special).all():
        out = np.where(is_bool, pay.astype(np.uint8), np.where(is_null, 2, np.where(is_abs, 3, 4))).astype(np.uint8)
        return SLOT_U8, out, (0, 0)
    if v2 and (is_str | is_bool | special).all() and is_str.any():
        win = two_windows(pay[is_str], 0x7FF0)
        if win is not None:
            lo, lo2 = win
            in2 = is_str & (pay >= np.uint64(lo + 0x7FF0))
            out = np.where(in2, pay - np.uint64(lo2) + np.uint64(0x8000), np.where(is_str, pay - np.uint64(lo), 0)).astype(np.uint16)
            out[is_abs], out[is_err], out[is_null] = 0xFFFF, 0xFFFE, 0xFFFD
            out[is_bool] = np.where(pay[is_bool] != 0, 0xFFFB, 0xFFFC).astype(np.uint16)
            return SLOT_U16_ID, out, win
    if (is_str | is_bool | special).all() and (not is_str.any() or int(pay[is_str].max()) < 0xFFFFFFF0):
        out = pay.astype(np.uint32)
        out[is_abs], out[is_err], out[is_null] = 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFD
        out[is_bool] = np.where(pay[is_bool] != 0, 0xFFFFFFFB, 0xFFFFFFFC).astype(np.uint32)
        return SLOT_U32_ID, out, (0, 0)
    is_heap = ((tag == L.V64_LIST) | (tag == L.V64_MAP)) & ((col & np.uint64(L.V64_HEAP_BATCH_BIT)) != 0)
    if (is_heap | special).all():
        off = pay & np.uint64(L.V64_HEAP_BATCH_BIT - 1)
        if not is_heap.any() or int(off[is_heap].max()) < 0x7FFFFFF0:
            out = off.astype(np.uint32) | np.where(tag == L.V64_MAP, 0x80000000, 0).astype(np.uint32)
            out[is_abs], out[is_err], out[is_null] = 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFD
            return SLOT_U32_HEAP, out, (0, 0)
    is_num = tag == 0
    if v2 and (is_num | special).all():
        d = col.view(np.float64)
        with np.errstate(invalid="ignore"):
            small = np.where(is_num & (d >= 0) & (d <= 239), d, 0).astype(np.uint8)
        if ((small.astype(np.float64).view(np.uint64) == col) | ~is_num).all():     # bit for bit: no -0.0, no fraction, no NaN
            out = small.copy()
            out[is_abs], out[is_err], out[is_null] = 0xFF, 0xFE, 0xFD
            return SLOT_U8_NUM, out, (0, 0)
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
            return SLOT_F32, out, (0, 0)
    return SLOT_U64, col, (0, 0)


def narrow_batch(batch, n_slots: int, v2: bool = True):
    """batch: cerbos_b200.encode.Batch.  -> NarrowBatch, or None when an id does not fit its 16 / 8-bit field.
    v2 = False keeps to the first form of the wire format (no constant elision, no 16-bit string ids / principal ids / heap)."""
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
    classes, cols, bases = [], [], []
    for v in range(n_slots):
        c, col, base = narrow_slot(np.ascontiguousarray(slots[v]), v2)
        classes.append(c)
        cols.append(np.ascontiguousarray(col))
        bases.append(base)
    tables = [np.ascontiguousarray(np.asarray(c)) for c in batch.columns[4:]]
    heap = tables[0]
    htag = _tags(heap)
    hpay = heap & _PAY
    small = (heap < np.uint64(1 << 31))
    hstr = (htag == L.V64_STRING) & (hpay < np.uint64(1 << 31))
    heap_u32 = bool((small | hstr).all())
    heap_bits = heap_base = heap_base2 = 0
    all_str = htag == L.V64_STRING
    if v2 and heap.size and (all_str | (heap < np.uint64(1 << 15))).all():
        win = two_windows(hpay[all_str], 1 << 14) if all_str.any() else (0, 0)
        if win is not None:
            heap_bits, (heap_base, heap_base2) = 16, win
            in2 = all_str & (hpay >= np.uint64(heap_base + (1 << 14)))
            tables[0] = np.where(in2, (hpay - np.uint64(heap_base2)) | np.uint64(0xC000),
                                 np.where(all_str, (hpay - np.uint64(heap_base)) | np.uint64(0x8000), heap)).astype(np.uint16)
    if not heap_bits and heap_u32:
        tables[0] = np.where(hstr, hpay | np.uint64(1 << 31), heap).astype(np.uint32)
    pid = np.ascontiguousarray(hdr0[:, 0])
    principal_base = None
    hdr_const_mask, hdr_const, versions_value = 0, [0, 0, 0, 0], None
    if v2 and n:
        lo = int(pid.min())
        if int(pid.max()) - lo <= 0xFFFF:
            principal_base, pid = lo, (pid - np.uint32(lo)).astype(np.uint16)
        keep = []
        for f in range(4):
            if (hdr16[:, f] == hdr16[0, f]).all():
                hdr_const_mask |= 1 << f
                hdr_const[f] = int(hdr16[0, f])
            else:
                keep.append(f)
        hdr16 = np.ascontiguousarray(hdr16[:, keep]) if keep else np.zeros((0,), dtype=np.uint16)
        if (versions == versions[0]).all():
            versions_value, versions = (int(versions[0, 0]), int(versions[0, 1])), None
    return NarrowBatch(n, batch.max_actions, roles.shape[0], pid, hdr16, versions, roles8,
                       np.array(classes or [0], dtype=np.uint8), cols, tables, heap_u32 and not heap_bits,
                       slot_base=np.array([b[0] for b in bases] or [0], dtype=np.uint32), slot_base2=np.array([b[1] for b in bases] or [0], dtype=np.uint32),
                       heap_base2=heap_base2, principal_base=principal_base, hdr_const_mask=hdr_const_mask,
                       hdr_const=hdr_const, versions_value=versions_value, heap_bits=heap_bits, heap_base=heap_base)
