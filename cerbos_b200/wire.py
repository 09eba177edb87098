"""Protobuf wire encoding of enginev1.CheckInput (api/public/cerbos/engine/v1/engine.proto `CheckInput`, `Principal`,
`Resource`, `AuxData`; attributes as google.protobuf.Value) from protojson-shaped dicts -- what a Go host would obtain
with proto.Marshal.  Used by the tests and the bench to feed the native encoder (cgpu_encode); hand-written, no
generated code.  Map entries are written in dict order so that both encoders see the attributes in the same order."""
from __future__ import annotations

import struct


def _varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(fno: int, payload: bytes) -> bytes:
    return _varint((fno << 3) | 2) + _varint(len(payload)) + payload


def _str(fno: int, s: str) -> bytes:
    return _ld(fno, s.encode("utf-8")) if s else b""


def value(v) -> bytes:
    """google.protobuf.Value"""
    if v is None:
        return _varint((1 << 3) | 0) + _varint(0)
    if isinstance(v, bool):
        return _varint((4 << 3) | 0) + _varint(1 if v else 0)
    if isinstance(v, (int, float)):
        return _varint((2 << 3) | 1) + struct.pack("<d", float(v))
    if isinstance(v, str):
        return _ld(3, v.encode("utf-8"))
    if isinstance(v, dict):
        return _ld(5, b"".join(_ld(1, _ld(1, str(k).encode("utf-8")) + _ld(2, value(x))) for k, x in v.items()))
    if isinstance(v, (list, tuple)):
        return _ld(6, b"".join(_ld(1, value(x)) for x in v))
    raise TypeError(type(v))


def _attr(fno: int, attr: dict) -> bytes:
    return b"".join(_ld(fno, _ld(1, str(k).encode("utf-8")) + _ld(2, value(v))) for k, v in (attr or {}).items())


def check_input(inp: dict) -> bytes:
    p, r = inp.get("principal") or {}, inp.get("resource") or {}
    aux = inp.get("auxData", inp.get("aux_data"))
    res = (_str(1, r.get("kind", "")) + _str(2, r.get("policyVersion", r.get("policy_version")) or "") + _str(3, r.get("id", "")) +
           _attr(4, r.get("attr")) + _str(5, r.get("scope") or ""))
    prin = (_str(1, p.get("id", "")) + _str(2, p.get("policyVersion", p.get("policy_version")) or "") +
            b"".join(_ld(3, x.encode("utf-8")) for x in (p.get("roles") or [])) + _attr(4, p.get("attr")) + _str(5, p.get("scope") or ""))
    out = _str(1, inp.get("requestId", "")) + _ld(2, res) + _ld(3, prin) + b"".join(_ld(4, a.encode("utf-8")) for a in (inp.get("actions") or []))
    if aux is not None:
        out += _ld(5, _attr(1, (aux or {}).get("jwt")))
    return out
