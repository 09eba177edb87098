"""Constant folding helpers for the table compiler: RFC 3339 timestamps and Go-style
duration strings (what cel-go's timestamp()/duration() accept) -> int64 nanoseconds.

Device timestamps are int64 nanoseconds since the Unix epoch (1678-2262); constants
outside that range make the table build fail (Unsupported) instead of diverging.
"""
from __future__ import annotations

import re

INT64_MIN = -(1 << 63)
INT64_MAX = (1 << 63) - 1

_RFC3339 = re.compile(
    r"\A([0-9]{4})-([0-9]{2})-([0-9]{2})T([0-9]{2}):([0-9]{2}):([0-9]{2})(?:[.,]([0-9]{1,9})[0-9]*)?(Z|[+-][0-9]{2}:[0-9]{2})\Z")


def days_from_civil(y: int, m: int, d: int) -> int:
    y -= m <= 2
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 719468


def parse_timestamp_ns(s: str) -> int:
    m = _RFC3339.match(s)
    if not m:
        raise ValueError(f"invalid timestamp {s!r}")
    y, mo, d, h, mi, sec = (int(m.group(i)) for i in range(1, 7))
    frac = m.group(7) or ""
    ns = int((frac + "000000000")[:9]) if frac else 0
    leap = y % 4 == 0 and (y % 100 != 0 or y % 400 == 0)
    dim = [31, 29 if leap else 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    if not (1 <= mo <= 12 and 1 <= d <= dim[mo - 1] and h < 24 and mi < 60 and sec < 60):
        raise ValueError(f"invalid timestamp {s!r}")
    tz = m.group(8)
    off = 0
    if tz != "Z":       # Go's time.Parse(time.RFC3339): 'T' / 'Z' literally; the offset's range test is `>` (24:60 passes)
        oh, om = int(tz[1:3]), int(tz[4:6])
        if oh > 24 or om > 60:
            raise ValueError(f"invalid timestamp {s!r}")
        off = (oh * 3600 + om * 60) * (1 if tz[0] == "+" else -1)
    secs = days_from_civil(y, mo, d) * 86400 + h * 3600 + mi * 60 + sec - off
    if secs < -62135596800 or secs > 253402300799:     # cel-go: the instant within 0001-01-01 .. 9999-12-31
        raise ValueError(f"timestamp {s!r} out of range")
    total = secs * 1_000_000_000 + ns
    if total < INT64_MIN or total > INT64_MAX:
        from .bytecode import Unsupported
        raise Unsupported(f"timestamp constant {s!r} outside the device range (1678..2262)")
    return total


_DUR_UNITS = {"ns": 1, "us": 1_000, "µs": 1_000, "μs": 1_000, "ms": 1_000_000,
              "s": 1_000_000_000, "m": 60_000_000_000, "h": 3_600_000_000_000}
_DUR_PART = re.compile(r"([0-9]*)(?:\.([0-9]*))?(ns|us|µs|μs|ms|s|m|h)")   # (ASCII digits only: \d would take any Unicode digit)


def parse_duration_ns(s: str) -> int:
    """Go time.ParseDuration."""
    orig = s
    if not s:
        raise ValueError("invalid duration")
    neg = False
    if s[0] in "+-":
        neg = s[0] == "-"
        s = s[1:]
    if s == "0":
        return 0
    if not s:
        raise ValueError(f"invalid duration {orig!r}")
    total = 0
    pos = 0
    while pos < len(s):
        m = _DUR_PART.match(s, pos)
        if not m:
            raise ValueError(f"invalid duration {orig!r}")
        whole, frac, unit = m.group(1), m.group(2), m.group(3)
        if whole == "" and not frac:
            raise ValueError(f"invalid duration {orig!r}")
        mult = _DUR_UNITS[unit]
        v = int(whole or "0") * mult
        if frac:
            v += int(frac) * mult // (10 ** len(frac))
        total += v
        pos = m.end()
    if neg:
        total = -total
    if total < INT64_MIN or total > INT64_MAX:
        raise ValueError(f"invalid duration {orig!r}")
    return total


def parse_ip(s: str):
    """Go net.ParseIP -> (family, 128-bit int in IPv6 form for v6 / 32-bit int for v4) or None.
    IPv4 dotted quads reject leading zeros (Go >= 1.17); IPv6 accepts '::' compression and an
    embedded dotted quad in the last 32 bits; zones ('%eth0') are rejected."""
    if "." in s and ":" not in s:
        parts = s.split(".")
        if len(parts) != 4:
            return None
        v = 0
        for p in parts:
            if not p.isdigit() or not p.isascii() or len(p) > 3 or (len(p) > 1 and p[0] == "0") or int(p) > 255:
                return None
            v = (v << 8) | int(p)
        return (4, v)
    if ":" not in s or "%" in s:
        return None
    tail4 = None
    if "." in s:
        head, _, last = s.rpartition(":")
        t = parse_ip(last)
        if t is None or t[0] != 4:
            return None
        tail4 = t[1]
        s = head + ":0:0"  # placeholder groups replaced below
    if s.count("::") > 1:
        return None
    if "::" in s:
        left, right = s.split("::")
        lg = left.split(":") if left else []
        rg = right.split(":") if right else []
        if len(lg) + len(rg) > 7:
            return None
        groups = lg + ["0"] * (8 - len(lg) - len(rg)) + rg
    else:
        groups = s.split(":")
        if len(groups) != 8:
            return None
    v = 0
    for g in groups:
        if not (1 <= len(g) <= 4) or any(c not in "0123456789abcdefABCDEF" for c in g):
            return None
        v = (v << 16) | int(g, 16)
    if tail4 is not None:
        v = (v & ~0xFFFFFFFF) | tail4
    return (6, v)


def parse_cidr(s: str):
    """Go net.ParseCIDR -> (family, prefix_bits, hi64, lo64) of the *network*, or None.
    For IPv4 the address lives in lo64's low 32 bits."""
    addr, sep, bits = s.partition("/")
    if not sep or not bits.isdigit() or not bits.isascii() or (len(bits) > 1 and bits[0] == "0") or len(bits) > 3:
        return None
    ip = parse_ip(addr)
    if ip is None:
        return None
    fam, v = ip
    width = 32 if fam == 4 else 128
    n = int(bits)
    if n > width:
        return None
    mask = ((1 << width) - 1) ^ ((1 << (width - n)) - 1)
    v &= mask
    return (fam, n, (v >> 64) & 0xFFFFFFFFFFFFFFFF, v & 0xFFFFFFFFFFFFFFFF)
