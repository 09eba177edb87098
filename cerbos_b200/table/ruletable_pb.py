"""Serialized ``runtimev1.RuleTable`` -> rule-table model (rule-table bundle ingestion).

The reference ships pre-compiled policy sets as rule-table bundles: a ``runtimev1.RuleTable`` protobuf message
(api/private/cerbos/runtime/v1/runtime.proto ``RuleTable``), optionally ChaCha20-Poly1305 encrypted, which
``OpenRuleTableBundle`` unmarshals and hands to ``ruletable.NewRuleTable`` (internal/storage/hub/ruletable_bundle.go:36-87);
``Manager`` applies the same rows on policy events (internal/ruletable/manager.go:126-181).  This module reads that wire
format directly -- a hand-written protobuf reader, no generated code (there is no protoc in this image) -- into the model
the flattener consumes (cerbos_b200/policy/model.py), so a PDP's own compiler output becomes the GPU table without going
through YAML again:

    blob = flatten(decode_rule_table(open("bundle_unencrypted.crrt", "rb").read())).blob

Conditions travel as ``Expr{original, checked}``: the CEL source text is re-parsed by cerbos_b200.cel.parser (the checked
AST carries nothing the lowering needs beyond what the text says).  Rule outputs, schemas and source attributes are
skipped: they do not influence an effect.
"""
from __future__ import annotations

import struct

from ..cel import parser as celparser
from ..policy.model import (Cond, DerivedRole, Expr, KIND_PRINCIPAL, KIND_RESOURCE, Params, Row, RuleTable, Variable)


class WireError(ValueError):
    pass


def _varint(buf, i):
    shift = v = 0
    while True:
        if i >= len(buf):
            raise WireError("truncated varint")
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, i
        shift += 7
        if shift > 63:
            raise WireError("varint too long")


def fields(buf):
    """Yields (field number, wire type, value) of one message: varint -> int, 64-bit / 32-bit -> bytes, length-delimited -> memoryview."""
    buf = memoryview(buf)
    i, n = 0, len(buf)
    while i < n:
        key, i = _varint(buf, i)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v, i = bytes(buf[i:i + 8]), i + 8
        elif wt == 2:
            ln, i = _varint(buf, i)
            if i + ln > n:
                raise WireError("truncated length-delimited field")
            v, i = buf[i:i + ln], i + ln
        elif wt == 5:
            v, i = bytes(buf[i:i + 4]), i + 4
        else:
            raise WireError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _s(v) -> str:
    return bytes(v).decode("utf-8")


def _map_entries(buf):
    """A map field is repeated {key = 1, value = 2} entries: -> (key raw, value raw) of one entry (None when absent)."""
    k = v = None
    for fno, _, val in fields(buf):
        if fno == 1:
            k = val
        elif fno == 2:
            v = val
    return k, v


def decode_value(buf):
    """google.protobuf.Value -> JSON-ish Python (numbers are floats, as structpb has them)."""
    out = None
    for fno, wt, v in fields(buf):
        if fno == 1:
            out = None
        elif fno == 2:
            out = struct.unpack("<d", v)[0]
        elif fno == 3:
            out = _s(v)
        elif fno == 4:
            out = bool(v)
        elif fno == 5:
            out = {}
            for f2, _, ent in fields(v):
                if f2 == 1:
                    k, val = _map_entries(ent)
                    out[_s(k) if k is not None else ""] = decode_value(val) if val is not None else None
        elif fno == 6:
            out = [decode_value(e) for f2, _, e in fields(v) if f2 == 1]
    return out


def decode_expr(buf) -> Expr:
    original = ""
    for fno, _, v in fields(buf):
        if fno == 1:
            original = _s(v)
    return Expr(original=original, ast=celparser.parse(original))


def decode_condition(buf) -> Cond:
    for fno, _, v in fields(buf):
        if fno in (1, 2, 3):
            kids = [decode_condition(e) for f2, _, e in fields(v) if f2 == 1]
            return Cond({1: "all", 2: "any", 3: "none"}[fno], children=kids)
        if fno == 4:
            return Cond("expr", expr=decode_expr(v))
    raise WireError("empty Condition")


def _variables(entries) -> list:
    out = []
    for v in entries:
        name, expr = "", None
        for fno, _, val in fields(v):
            if fno == 1:
                name = _s(val)
            elif fno == 2:
                expr = decode_expr(val)
        out.append(Variable(name=name, expr=expr))
    return out


def decode_params(buf, key: str) -> Params:
    ordered, constants = [], {}
    for fno, _, v in fields(buf):
        if fno == 1:
            ordered.append(v)
        elif fno == 2:
            k, val = _map_entries(v)
            constants[_s(k)] = decode_value(val) if val is not None else None
    return Params(key=key, variables=_variables(ordered), constants=constants)


def _string_set(buf) -> list:
    """map<string, google.protobuf.Empty> (one entry) -> its key"""
    k, _ = _map_entries(buf)
    return _s(k) if k is not None else ""


_KINDS = {3: KIND_PRINCIPAL, 4: KIND_RESOURCE}


def decode_rule_row(buf) -> Row:
    r = Row()
    params_raw = dr_params_raw = None
    kind = 0
    for fno, _, v in fields(buf):
        if fno == 1:
            r.origin_fqn = _s(v)
        elif fno == 2:
            r.resource = _s(v)
        elif fno == 3:
            r.role = _s(v)
        elif fno == 4:
            r.action = _s(v)
        elif fno == 15:
            r.allow_actions = [_string_set(e) for f2, _, e in fields(v) if f2 == 1]
        elif fno == 5:
            r.condition = decode_condition(v)
        elif fno == 6:
            r.dr_condition = decode_condition(v)
        elif fno == 7:
            r.effect = int(v)
        elif fno == 8:
            r.scope = _s(v)
        elif fno == 9:
            r.scope_permissions = int(v)
        elif fno == 10:
            r.version = _s(v)
        elif fno == 11:
            r.origin_derived_role = _s(v)
        elif fno == 12:      # Output emit_output { When when = 1 { Expr rule_activated = 1; Expr condition_not_met = 2 } }
            for f1, _, w in fields(v):
                if f1 == 1:
                    for f2, _, e in fields(w):
                        if f2 == 1:
                            r.emit_activated = decode_expr(e)
                        elif f2 == 2:
                            r.emit_not_met = decode_expr(e)
        elif fno == 13:
            r.name = _s(v)
        elif fno == 14:
            r.principal = _s(v)
        elif fno == 16:
            params_raw = v
        elif fno == 17:
            dr_params_raw = v
        elif fno == 18:
            r.evaluation_key = _s(v)
        elif fno == 19:
            kind = int(v)
        elif fno == 20:
            r.from_role_policy = bool(v)
    r.policy_kind = _KINDS.get(kind, KIND_RESOURCE)
    # the per-request variable caches are keyed by the policy (ruletable.go:869-884): the origin FQN identifies it
    r.params = decode_params(params_raw, r.origin_fqn) if params_raw is not None else None
    r.dr_params = decode_params(dr_params_raw, f"{r.origin_derived_role}@{r.origin_fqn}") if dr_params_raw is not None else None
    if r.allow_actions is not None and r.action is None:
        r.action = None
    return r


def decode_derived_role(buf) -> DerivedRole:
    name, parents, cond, ordered, constants, origin = "", [], None, [], {}, ""
    for fno, _, v in fields(buf):
        if fno == 1:
            name = _s(v)
        elif fno == 2:
            parents.append(_string_set(v))
        elif fno == 4:
            cond = decode_condition(v)
        elif fno == 5:
            ordered.append(v)
        elif fno == 6:
            k, val = _map_entries(v)
            constants[_s(k)] = decode_value(val) if val is not None else None
        elif fno == 7:
            origin = _s(v)
    return DerivedRole(name=name, parent_roles=sorted(parents), condition=cond,
                       params=Params(key=f"{origin}#{name}", variables=_variables(ordered), constants=constants), origin_fqn=origin)


def decode_rule_table(buf, skip=None) -> RuleTable:
    """Serialized runtimev1.RuleTable -> RuleTable model.  skip(row) -> True drops a row (e.g. policies the device does not cover)."""
    rt = RuleTable()
    meta_fqn: dict[int, str] = {}
    pdr_raw: list = []
    for fno, _, v in fields(buf):
        if fno == 1:
            row = decode_rule_row(v)
            if skip is None or not skip(row):
                rt.rows.append(row)
        elif fno == 3:           # map<uint64, RuleTableMetadata>: module id -> fqn
            k, val = _map_entries(v)
            if val is not None:
                for f2, _, x in fields(val):
                    if f2 == 1:
                        meta_fqn[int(k or 0)] = _s(x)
        elif fno == 4:           # map<string scope, RoleParentRoles>
            k, val = _map_entries(v)
            scope = _s(k) if k is not None else ""
            rmap = rt.scope_parent_roles.setdefault(scope, {})
            if val is not None:
                for f2, _, ent in fields(val):
                    if f2 == 1:
                        role, pr = _map_entries(ent)
                        rmap[_s(role)] = [_s(x) for f3, _, x in fields(pr) if f3 == 1] if pr is not None else []
        elif fno == 5:           # map<uint64 module id, PolicyDerivedRoles>
            pdr_raw.append(_map_entries(v))
    for k, val in pdr_raw:
        fqn = meta_fqn.get(int(k or 0))
        if fqn is None or val is None:
            continue
        drs = rt.policy_derived_roles.setdefault(fqn, {})
        for f2, _, ent in fields(val):
            if f2 == 1:
                nm, dr = _map_entries(ent)
                if dr is not None:
                    d = decode_derived_role(dr)
                    drs[_s(nm)] = d
    return rt
