"""Synthetic CheckResources workloads C1..C3 (BASELINE.json `configs`).

Every workload provides
  * ``policies()``            policy documents (dicts, same shape as the reference's YAML),
  * ``fields(n)``             per-request logical fields as numpy arrays, drawn from SplitMix64
                              streams (seed 0xCE4B05 + cfg#, stream index = request index),
  * ``inputs(fields, idx)``   CheckInput dicts for a subset (what the reference's Go harness or the
                              Python oracle would be handed), and
  * ``columns(fields, enc)``  the same requests encoded directly (vectorised) into the SoA batch
                              columns of cerbos_b200/encode.py -- used for the 2^20 .. 2^24 batches
                              where a per-request Python encoder would take minutes.
tests/test_workloads.py checks that both routes give identical decisions.
"""
from __future__ import annotations

import numpy as np

from .encode import Batch, Encoder, passes_for
from .table import layout as L

SEED_BASE = 0xCE4B05
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_DRAWS = 32  # draws reserved per request


_START = 0  # first request index of the window being generated (set by fields(n, start))


def splitmix(seed: int, n: int, draw: int) -> np.ndarray:
    """draw-th SplitMix64 output of the stream of every request _START.._START+n-1."""
    with np.errstate(over="ignore"):
        idx = (np.arange(n, dtype=np.uint64) + np.uint64(_START)) * np.uint64(_DRAWS) + np.uint64(draw + 1)
        z = np.uint64(seed) + idx * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _uniform(seed, n, draw, k):
    return (splitmix(seed, n, draw) % np.uint64(k)).astype(np.int64)


def _prob(seed, n, draw, p):
    return (splitmix(seed, n, draw) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53)) < p


def _box(tag, payload):
    return (np.uint64((L.V64_BOX_BASE | tag) << 48)) | payload.astype(np.uint64)


def _str_ids(enc: Encoder, strings):
    """Assigns ids the way the encoder would: table strings keep their table id, the rest are
    appended to the batch string table. Returns (ids array, batch string list)."""
    nts = enc.n_table_strings
    ids, extra, seen = [], [], {}
    for s in strings:
        i = enc.table_strings.get(s)
        if i is None:
            i = seen.get(s)
            if i is None:
                i = nts + len(extra)
                seen[s] = i
                extra.append(s.encode("utf-8"))
        ids.append(i)
    return np.array(ids, dtype=np.uint64), extra


def _kind_classes(enc: Encoder, kinds):
    """-> (hdr kind_class value per kind, CSR class list) using the encoder's direct / CSR encoding."""
    classes, class_list, vals = {}, [], []
    for k in kinds:
        vals.append(enc.kind_class(enc.kind_patterns(k), classes, class_list))
    return np.array(vals, dtype=np.uint32), class_list


def _finish_batch(enc: Encoder, n, hdr0, hdr1, roles, slots, heap, bstr_list, class_list, aset_list, max_actions):
    role_cols = roles.shape[0]
    class_off = np.zeros(len(class_list) + 1, dtype=np.uint32)
    cp = []
    for c, pats in enumerate(class_list):
        class_off[c] = len(cp)
        cp.extend(pats)
    class_off[len(class_list)] = len(cp)
    aset_k, aset_spread, row_am = enc.build_action_sets(aset_list, role_cols, max_actions)
    off = np.zeros(len(bstr_list) + 1, dtype=np.uint32)
    pos = 0
    for j, b in enumerate(bstr_list):
        off[j] = pos
        pos += len(b)
    off[len(bstr_list)] = pos
    bbytes = np.frombuffer(b"".join(bstr_list) + b"\0" * 16, dtype=np.uint8)
    kc, n_pass = passes_for(max_actions, role_cols)
    cols = [hdr0, hdr1, roles, slots, heap, off, bbytes, class_off, np.array(cp or [0], dtype=np.uint32),
            aset_k, aset_spread, row_am]
    return Batch(n, max_actions, role_cols, cols, None, n_pass, kc)


# =========================================================================================== C1
class C1:
    """hack/loadtest-shaped: 1 resource policy, 3 actions, role-only rules, 32x32 = 1024 pairs."""
    name = "C1"
    cfg = 1
    actions = ["view", "edit", "delete"]
    role_names = ["user", "editor", "admin", "guest"]
    default_n = 1024
    role_cols = 2
    n_slots = 0
    heap_bytes = 0

    def policies(self):
        return [{"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {
            "resource": "document", "version": "default", "rules": [
                {"actions": ["view"], "effect": "EFFECT_ALLOW", "roles": ["user", "admin"]},
                {"actions": ["edit"], "effect": "EFFECT_ALLOW", "roles": ["editor", "admin"]},
                {"actions": ["delete"], "effect": "EFFECT_ALLOW", "roles": ["admin"]},
            ]}}]

    def fields(self, n=None, start=0):
        n = n or self.default_n
        seed = SEED_BASE + self.cfg
        pr = (np.arange(n) + start) // 32 % 32      # principal index, materialised flat (one row per pair)
        rs = (np.arange(n) + start) % 32
        # roles are a property of the principal: derive them from the principal's own stream
        r0 = _uniform(seed, 32, 0, 4)[pr]
        two = _prob(seed, 32, 1, 0.5)[pr]
        r1 = (r0 + 1 + _uniform(seed, 32, 2, 3)[pr]) % 4
        return {"n": n, "pr": pr, "rs": rs, "r0": r0, "r1": np.where(two, r1, -1)}

    def inputs(self, f, idx):
        out = []
        for i in idx:
            roles = [self.role_names[f["r0"][i]]] + ([self.role_names[f["r1"][i]]] if f["r1"][i] >= 0 else [])
            out.append({"requestId": str(i), "actions": list(self.actions),
                        "principal": {"id": f"user{f['pr'][i]}", "roles": roles},
                        "resource": {"kind": "document", "id": f"doc{f['rs'][i]}"}})
        return out

    def columns(self, f, enc: Encoder) -> Batch:
        n = f["n"]
        pid_ids, extra = _str_ids(enc, [f"user{i}" for i in range(32)])
        hdr0 = np.zeros((n, 4), dtype=np.uint32)
        hdr0[:, 0] = pid_ids[f["pr"]]
        kvals, class_list = _kind_classes(enc, ["document"])
        hdr0[:, 1] = kvals[0]
        hdr0[:, 2] = enc.resolve_scope("")
        hdr0[:, 3] = enc.resolve_scope("")
        hdr1 = np.zeros(n, dtype=np.dtype([("rv", "<u2"), ("pv", "<u2"), ("aset", "<u4")]))
        hdr1["rv"] = enc.version_ids.get("default", L.NONE16)
        hdr1["pv"] = hdr1["rv"]
        rmap = np.array([enc.role_ids.get(r, L.ROLE_UNKNOWN) for r in self.role_names] + [L.ROLE_PAD], dtype=np.uint32)
        roles = np.stack([rmap[f["r0"]], rmap[f["r1"]]]).astype(np.uint32)
        slots = np.zeros((1, n), dtype=np.uint64)
        return _finish_batch(enc, n, hdr0, hdr1, roles, slots, np.zeros(1, dtype=np.uint64), extra,
                             class_list, [tuple(self.actions)], 3)

    # 24 + 4R + 8A + S + ceil(K/8)
    def bytes_per_request(self):
        return 24 + 4 * 2 + 0 + 0 + 1


# =========================================================================================== C2
class C2:
    """10 resource policies x 8 actions, 2 derived roles with CEL on request.resource.attr; batch 2^20.
    This is the configuration BASELINE.json's metric is quoted on for 1xB200."""
    name = "C2"
    cfg = 2
    actions = [f"a{i}" for i in range(8)]
    role_names = ["user", "manager", "admin"]
    statuses = ["OPEN", "PENDING", "CLOSED", "ARCHIVED"]
    default_n = 1 << 20
    n_principals = 65536
    n_kinds = 10
    role_cols = 2

    def policies(self):
        docs = [{"apiVersion": "api.cerbos.dev/v1", "derivedRoles": {"name": "c2_roles", "definitions": [
            {"name": "owner", "parentRoles": ["user"],
             "condition": {"match": {"expr": "R.attr.owner == P.id"}}},
            {"name": "dept_manager", "parentRoles": ["manager"],
             "condition": {"match": {"expr": 'R.attr.dept == P.attr.dept && R.attr.status in ["OPEN","PENDING"]'}}},
        ]}}]
        for k in range(self.n_kinds):
            docs.append({"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {
                "resource": f"kind_{k}", "version": "default", "importDerivedRoles": ["c2_roles"], "rules": [
                    {"actions": ["a0", "a1", "a2"], "effect": "EFFECT_ALLOW", "roles": ["user"]},
                    {"actions": ["a3", "a4"], "effect": "EFFECT_ALLOW", "derivedRoles": ["owner"]},
                    {"actions": ["a5", "a6"], "effect": "EFFECT_ALLOW", "derivedRoles": ["dept_manager"]},
                    {"actions": ["a7"], "effect": "EFFECT_ALLOW", "roles": ["admin"]},
                    {"actions": ["*"], "effect": "EFFECT_DENY", "roles": ["*"],
                     "condition": {"match": {"expr": "R.attr.locked == true"}}},
                ]}})
        return docs

    def fields(self, n=None, start=0):
        """Requests [start, start+n) of the workload's stream."""
        global _START
        n = n or self.default_n
        seed = SEED_BASE + self.cfg
        _START = start
        try:
            return self._fields(n, seed)
        finally:
            _START = 0

    def _fields(self, n, seed):
        pid = _uniform(seed, n, 1, self.n_principals)
        own_other = _uniform(seed, n, 5, self.n_principals)
        r0 = _uniform(seed, n, 2, 3)
        r1 = (r0 + 1 + _uniform(seed, n, 4, 2)) % 3
        return {
            "n": n,
            "kind": _uniform(seed, n, 0, self.n_kinds),
            "pid": pid,
            "r0": r0,
            "r1": np.where(_prob(seed, n, 3, 0.5), r1, -1),
            "owner": np.where(_prob(seed, n, 6, 0.25), pid, own_other),
            "pdept": _uniform(seed, n, 7, 16),
            "rdept": _uniform(seed, n, 8, 16),
            "status": _uniform(seed, n, 9, 4),
            "locked": _prob(seed, n, 10, 0.05),
        }

    def inputs(self, f, idx):
        out = []
        for i in idx:
            roles = [self.role_names[f["r0"][i]]] + ([self.role_names[f["r1"][i]]] if f["r1"][i] >= 0 else [])
            out.append({"requestId": str(i), "actions": list(self.actions),
                        "principal": {"id": f"p{f['pid'][i]}", "roles": roles, "attr": {"dept": f"d{f['pdept'][i]}"}},
                        "resource": {"kind": f"kind_{f['kind'][i]}", "id": f"r{i}", "attr": {
                            "owner": f"p{f['owner'][i]}", "dept": f"d{f['rdept'][i]}",
                            "status": self.statuses[f["status"][i]], "locked": bool(f["locked"][i])}}})
        return out

    def columns(self, f, enc: Encoder) -> Batch:
        n = f["n"]
        pid_ids, extra = _str_ids(enc, [f"p{i}" for i in range(self.n_principals)])
        nts = enc.n_table_strings
        dept_ids, e2 = _str_ids(enc, [f"d{i}" for i in range(16)])
        # continue numbering of batch strings after `extra`
        dept_ids = np.where(dept_ids >= nts, dept_ids + np.uint64(len(extra)), dept_ids)
        extra = extra + e2
        st_ids, e3 = _str_ids(enc, self.statuses)
        st_ids = np.where(st_ids >= nts, st_ids + np.uint64(len(extra)), st_ids)
        extra = extra + e3
        kvals, class_list = _kind_classes(enc, [f"kind_{k}" for k in range(self.n_kinds)])
        hdr0 = np.zeros((n, 4), dtype=np.uint32)
        hdr0[:, 0] = pid_ids[f["pid"]]
        hdr0[:, 1] = kvals[f["kind"]]
        hdr0[:, 2] = enc.resolve_scope("")
        hdr0[:, 3] = enc.resolve_scope("")
        hdr1 = np.zeros(n, dtype=np.dtype([("rv", "<u2"), ("pv", "<u2"), ("aset", "<u4")]))
        hdr1["rv"] = enc.version_ids.get("default", L.NONE16)
        hdr1["pv"] = hdr1["rv"]
        rmap = np.array([enc.role_ids.get(r, L.ROLE_UNKNOWN) for r in self.role_names] + [L.ROLE_PAD], dtype=np.uint32)
        roles = np.stack([rmap[f["r0"]], rmap[f["r1"]]]).astype(np.uint32)
        slots = np.zeros((max(len(enc.slots), 1), n), dtype=np.uint64)
        vals = {
            ("principal", "attr", "dept"): _box(L.V64_STRING, dept_ids[f["pdept"]]),
            ("resource", "attr", "owner"): _box(L.V64_STRING, pid_ids[f["owner"]]),
            ("resource", "attr", "dept"): _box(L.V64_STRING, dept_ids[f["rdept"]]),
            ("resource", "attr", "status"): _box(L.V64_STRING, st_ids[f["status"]]),
            ("resource", "attr", "locked"): _box(L.V64_BOOL, f["locked"].astype(np.uint64)),
        }
        for s, path in enumerate(enc.slots):
            slots[s] = vals[path]
        return _finish_batch(enc, n, hdr0, hdr1, roles, slots, np.zeros(1, dtype=np.uint64), extra,
                             class_list, [tuple(self.actions)], 8)

    def bytes_per_request(self):
        return 24 + 4 * 2 + 8 * 5 + 0 + 1   # = 73


WORKLOADS = {"C1": C1, "C2": C2}


def build(workload, globals_=None):
    """-> (rule table, FlatTable, Encoder)"""
    from .policy.compile import build_rule_table
    from .table.flatten import flatten
    rt = build_rule_table(workload.policies())
    ft = flatten(rt, globals_=globals_)
    return rt, ft, Encoder(ft.manifest)
