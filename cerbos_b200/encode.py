"""CheckInput batch -> SoA request columns (host side, per Check call).

This is the host half of the replacement for the per-input work the reference does
with strings and maps inside ``RuleTable.check`` (internal/ruletable/ruletable.go:785-884):
default version/scope (:789-799, evaluator.go:99-113), ``namer.SanitizedResource`` (:851),
resolving the request scope against the per-kind scope sets (GetAllScopes :611-645),
and -- because actions / resource kinds / roles are open vocabularies matched by globs
(internal/util/globs_common.go, glob_map.go:138-186) -- turning every distinct string
into dictionary ids / pattern classes *once per batch*.  The device sees integers only.

Column order (``cgpu_batch.columns``; see include/cerbos_b200.h):
   0 hdr0        u32[N][4]  principal_id(string id), kind_class, resource_scope, principal_scope
   1 hdr1        {u16 resource_version, u16 principal_version, u32 action_set_id}[N]
   2 roles       u32[role_cols][N]
   3 slots       u64[n_slots][N]        NaN-boxed attribute values (layout.V64_*)
   4 heap        u64[]                  lists / maps referenced from slots
   5 bstr_off    u32[n_batch_strings+1]
   6 bstr_bytes  u8[]
   7 class_off   u32[n_classes+1]       kind class -> resource pattern ids
   8 class_pats  u32[]
   9 aset_k      u32[n_asets]           number of actions in each action set
  10 aset_spread u64[n_pass][n_asets][n_apats]   action-pattern -> (action x role-column) bit spread
  11 row_am      u64[n_pass][n_asets][n_rows]    the same spread OR-ed over the patterns of every table row
"""
from __future__ import annotations

import json
import math
import struct

import numpy as np

from .policy import namer
from .policy.globs import key_matches
from .table import layout as L

N_COLUMNS = 12


def _f64_bits(d: float) -> int:
    if math.isnan(d):
        return L.V64_CANON_NAN
    return struct.unpack("<Q", struct.pack("<d", float(d)))[0]


def _box(tag: int, payload: int = 0) -> int:
    return ((L.V64_BOX_BASE | tag) << 48) | (payload & ((1 << 48) - 1))


V_ABSENT = _box(L.V64_ABSENT)
V_ERROR = _box(L.V64_ERROR)
V_NULL = _box(L.V64_NULL)


class Batch:
    """Encoded batch: list of numpy columns + scalars, ready for cgpu_check."""

    def __init__(self, n, max_actions, role_cols, columns, action_lists, n_pass, kc):
        self.n = n
        self.max_actions = max_actions
        self.role_cols = role_cols
        self.columns = columns
        self.action_lists = action_lists  # per request: list of action strings (for decoding results)
        self.n_pass = n_pass
        self.kc = kc

    def nbytes(self):
        return sum(int(c.nbytes) for c in self.columns)


def passes_for(max_actions: int, role_cols: int):
    kc = max(1, min(max(max_actions, 1), 64 // max(role_cols, 1)))
    n_pass = (max(max_actions, 1) + kc - 1) // kc
    return kc, n_pass


class Encoder:
    """Encodes against one flattened table (its MANIFEST)."""

    def __init__(self, manifest: dict, default_version="default", default_scope="", lenient_scope_search=False):
        self.m = manifest
        self.default_version = default_version
        self.default_scope = default_scope
        self.lenient = lenient_scope_search
        self.version_ids = {v: i for i, v in enumerate(manifest["versions"])}
        self.scope_ids = {s: i for i, s in enumerate(manifest["scopes"])}
        self.role_ids = {r: i for i, r in enumerate(manifest["roles"])}
        self.table_strings = {s: i for i, s in enumerate(manifest["strings"])}
        self.n_table_strings = len(manifest["strings"])
        self.apats = manifest["apats"]
        self.respats = manifest["respats"]
        self.slots = [tuple(p) for p in manifest["slots"]]
        self.row_pat_start = np.array(manifest.get("row_pat_start", []), dtype=np.int64)
        self.row_apats = np.array(manifest.get("row_apats", []), dtype=np.int64)
        self._apat_cache: dict[str, tuple] = {}
        self._class_cache: dict[str, tuple] = {}

    # ---------------------------------------------------------------- dictionaries
    def action_patterns(self, action: str) -> tuple:
        r = self._apat_cache.get(action)
        if r is None:
            r = tuple(i for i, p in enumerate(self.apats) if key_matches(p, action))
            self._apat_cache[action] = r
        return r

    def kind_patterns(self, kind: str) -> tuple:
        r = self._class_cache.get(kind)
        if r is None:
            sk = namer.sanitize(kind)
            r = tuple(i for i, p in enumerate(self.respats) if key_matches(p, sk))
            if len(r) > L.MAX_CLASS_PATS:
                raise ValueError(f"resource kind {kind!r} matches more than {L.MAX_CLASS_PATS} resource patterns")
            self._class_cache[kind] = r
        return r

    def resolve_scope(self, scope: str) -> int:
        sid = self.scope_ids.get(scope)
        if sid is not None:
            return sid
        if self.lenient:
            for anc in namer.scope_parents(scope):
                sid = self.scope_ids.get(anc)
                if sid is not None:
                    return sid | L.SCOPE_INEXACT_BIT
        return L.SCOPE_NONE

    # ---------------------------------------------------------------- encode
    def encode(self, inputs: list) -> Batch:
        n = len(inputs)
        n_slots = len(self.slots)
        max_roles = max([len((i.get("principal") or {}).get("roles") or []) for i in inputs] + [1])
        if max_roles > L.MAX_ROLE_COLS:
            raise ValueError(f"more than {L.MAX_ROLE_COLS} roles on one principal is not supported")
        role_cols = max_roles
        max_actions = max([len(i.get("actions") or []) for i in inputs] + [1])
        kc, n_pass = passes_for(max_actions, role_cols)

        hdr0 = np.zeros((n, 4), dtype=np.uint32)
        hdr1 = np.zeros(n, dtype=np.dtype([("rv", "<u2"), ("pv", "<u2"), ("aset", "<u4")]))
        roles = np.full((role_cols, n), L.ROLE_PAD, dtype=np.uint32)
        slots = np.zeros((max(n_slots, 1), n), dtype=np.uint64)
        heap: list[int] = []
        bstr: dict[str, int] = {}
        bstr_list: list[bytes] = []
        classes: dict[tuple, int] = {}
        class_list: list[tuple] = []
        asets: dict[tuple, int] = {}
        aset_list: list[tuple] = []
        nts = self.n_table_strings

        def sid(s: str) -> int:
            i = self.table_strings.get(s)
            if i is not None:
                return i
            i = bstr.get(s)
            if i is None:
                i = len(bstr_list)
                bstr[s] = i
                bstr_list.append(s.encode("utf-8"))
            return nts + i

        def v64(v) -> int:
            if v is None:
                return V_NULL
            if isinstance(v, bool):
                return _box(L.V64_BOOL, int(v))
            if isinstance(v, (int, float)):
                return _f64_bits(float(v))
            if isinstance(v, str):
                return _box(L.V64_STRING, sid(v))
            if isinstance(v, (list, tuple)):
                elems = [v64(x) for x in v]
                off = len(heap)
                heap.append(len(elems))
                heap.extend(elems)
                return _box(L.V64_LIST, off | L.V64_HEAP_BATCH_BIT)
            if isinstance(v, dict):
                keys = [_box(L.V64_STRING, sid(str(k))) for k in v.keys()]
                vals = [v64(x) for x in v.values()]
                off = len(heap)
                heap.append(len(keys))
                heap.extend(keys)
                heap.extend(vals)
                return _box(L.V64_MAP, off | L.V64_HEAP_BATCH_BIT)
            raise TypeError(f"not a JSON value: {type(v)}")

        def walk(root: dict, segs) -> int:
            cur = root
            for j, s in enumerate(segs):
                if not isinstance(cur, dict):
                    return V_ERROR
                if s not in cur:
                    return V_ABSENT if j == len(segs) - 1 else V_ERROR
                cur = cur[s]
            return v64(cur)

        for i, inp in enumerate(inputs):
            p = inp.get("principal") or {}
            r = inp.get("resource") or {}
            aux = inp.get("auxData", inp.get("aux_data")) or {}
            p_scope = namer.scope_value(p.get("scope") or self.default_scope)
            r_scope = namer.scope_value(r.get("scope") or self.default_scope)
            p_ver = p.get("policyVersion", p.get("policy_version")) or self.default_version
            r_ver = r.get("policyVersion", r.get("policy_version")) or self.default_version
            kp = self.kind_patterns(r.get("kind", ""))
            cid = self.kind_class(kp, classes, class_list)
            acts = tuple(inp.get("actions") or [])
            aid = asets.get(acts)
            if aid is None:
                aid = len(aset_list)
                asets[acts] = aid
                aset_list.append(acts)
            hdr0[i] = (sid(p.get("id", "")), cid, self.resolve_scope(r_scope), self.resolve_scope(p_scope))
            hdr1[i] = (self.version_ids.get(r_ver, L.NONE16), self.version_ids.get(p_ver, L.NONE16), aid)
            for j, role in enumerate(p.get("roles") or []):
                roles[j, i] = self.role_ids.get(role, L.ROLE_UNKNOWN)
            for s, path in enumerate(self.slots):
                if path[0] == "aux_data":
                    val = walk(aux.get("jwt") or {}, path[2:]) if len(path) > 2 else v64(aux.get("jwt") or {})
                else:
                    msg = p if path[0] == "principal" else r
                    fld = path[1]
                    if fld == "attr":
                        attr = msg.get("attr") or {}
                        val = walk(attr, path[2:]) if len(path) > 2 else v64(attr)
                    elif fld == "roles":
                        val = v64(list(msg.get("roles") or []))
                    elif fld == "scope":
                        val = v64(namer.scope_value(msg.get("scope") or ""))
                    elif fld == "policy_version":
                        val = v64(msg.get("policyVersion", msg.get("policy_version")) or "")
                    else:  # id, kind
                        val = v64(msg.get(fld, "") or "")
                slots[s, i] = val

        # kind classes (CSR)
        class_off = np.zeros(len(class_list) + 1, dtype=np.uint32)
        cp = []
        for c, pats in enumerate(class_list):
            class_off[c] = len(cp)
            cp.extend(pats)
        class_off[len(class_list)] = len(cp)
        class_pats = np.array(cp or [0], dtype=np.uint32)

        aset_k, aset_spread, row_am = self.build_action_sets(aset_list, role_cols, max_actions)

        off = np.zeros(len(bstr_list) + 1, dtype=np.uint32)
        pos = 0
        for j, b in enumerate(bstr_list):
            off[j] = pos
            pos += len(b)
        off[len(bstr_list)] = pos
        bbytes = np.frombuffer(b"".join(bstr_list) + b"\0" * 16, dtype=np.uint8)
        cols = [hdr0, hdr1, roles, slots, np.array(heap or [0], dtype=np.uint64), off, bbytes,
                class_off, class_pats, aset_k, aset_spread, row_am]
        return Batch(n, max_actions, role_cols, cols, [list(a.get("actions") or []) for a in inputs], n_pass, kc)

    def build_action_sets(self, aset_list, role_cols: int, max_actions: int):
        """aset_k[u32 n_asets], aset_spread[u64 n_pass][n_asets][n_apats]: bit (kk*role_cols + i), for
        every role column i, of action kk (within its pass) whose string matches the pattern."""
        kc, n_pass = passes_for(max_actions, role_cols)
        n_ap = max(len(self.apats), 1)
        n_as = max(len(aset_list), 1)
        aset_k = np.zeros(n_as, dtype=np.uint32)
        spread = np.zeros((n_pass, n_as, n_ap), dtype=np.uint64)
        for a, acts in enumerate(aset_list):
            aset_k[a] = len(acts)
            for k, act in enumerate(acts):
                ps, kk = divmod(k, kc)
                for ap in self.action_patterns(act):
                    spread[ps, a, ap] |= np.uint64(1 << (kk * role_cols))
        return aset_k, spread, self.row_action_masks(spread)

    def row_action_masks(self, spread: np.ndarray) -> np.ndarray:
        """row_am[pass][aset][row] = OR of spread[pass][aset][p] over the action patterns p of table row."""
        n_rows = len(self.row_pat_start)
        n_pass, n_as, _ = spread.shape
        if n_rows == 0:
            return np.zeros((n_pass, n_as, 1), dtype=np.uint64)
        out = np.zeros((n_pass, n_as, n_rows), dtype=np.uint64)
        for ps in range(n_pass):
            for a in range(n_as):
                out[ps, a] = np.bitwise_or.reduceat(spread[ps, a][self.row_apats], self.row_pat_start)
        return out

    @staticmethod
    def kind_class(kp: tuple, classes: dict, class_list: list) -> int:
        """hdr0.kind_class: the pattern id itself when the kind matches exactly one resource pattern."""
        if len(kp) == 0:
            return L.KIND_NONE
        if len(kp) == 1:
            return kp[0]
        cid = classes.get(kp)
        if cid is None:
            cid = len(class_list)
            classes[kp] = cid
            class_list.append(kp)
        return cid | L.KIND_CLASS_CSR_BIT


def manifest_from_blob(blob: bytes) -> dict:
    magic, version, n_sec, _flags, _total, _res = struct.unpack_from("<IIIIQQ", blob, 0)
    if magic != L.MAGIC or version != L.VERSION:
        raise ValueError("not a cerbos_b200 table blob (bad magic / version)")
    for i in range(n_sec):
        sid_, _eb, off, nb = struct.unpack_from("<IIQQ", blob, 32 + 24 * i)
        if sid_ == L.SECTIONS["MANIFEST"]:
            return json.loads(bytes(blob[off:off + nb]).decode("utf-8"))
    raise ValueError("table blob has no MANIFEST section")
