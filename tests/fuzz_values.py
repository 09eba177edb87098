"""Seeded random CEL expressions over the functions that BUILD values at run time (cel-go ext.Strings / ext.Lists,
Cerbos except / intersect, comprehensions that collect, concatenation, RE2 matches) and random requests whose attributes
change type from request to request -- for the differential test of oracle #1 against the kernel core
(tests/test_fuzz_values.py).  Typed generators: S string, I int, L list of strings, N list of ints, B bool."""
import random

STR_LITS = ["a", "b", "ab", "abc", "a,b,,c", "héllo wörld", " pad ", "", "A.B.c", "x1", "日本語", "a-b_c", "\u00a0x\u2003", "\u0085a\u3000", "\ufeffb\u200b",
            "\x1cq\x1f", "e\u0301a", "\U0001F600b", "aaa", "abab", "ÀB"]
PATTERNS = ["^a", "b$", "^[a-c]+$", "l+o", "^$", "a|x", "\\\\d", "^h.llo", "[[:alpha:]]+", "(ab)+", "."]
ATTR_S = ["P.attr.s", "R.attr.t", "P.attr.u", "R.attr.csv"]
ATTR_I = ["P.attr.n", "R.attr.k"]
ATTR_L = ["P.attr.l", "R.attr.m", "P.attr.tags"]
ATTR_N = ["R.attr.nums", "P.attr.ids"]


def _lit(s):
    """CEL source of a string literal (the lists above hold CEL source text: a backslash there is already doubled)"""
    return '"' + s + '"'


def S(r, d=0):
    k = r.randrange(14 if d < 3 else 3)
    if k == 0:
        return _lit(r.choice(STR_LITS))
    if k in (1, 2):
        return r.choice(ATTR_S)
    if k == 3:
        return f"{S(r, d + 1)}.{r.choice(['lowerAscii', 'upperAscii', 'trim', 'reverse'])}()"
    if k == 4:
        return f"{S(r, d + 1)}.charAt({I(r, d + 1)})"
    if k == 5:
        return f"{S(r, d + 1)}.substring({I(r, d + 1)}" + (f", {I(r, d + 1)})" if r.random() < 0.6 else ")")
    if k == 6:
        return f"{S(r, d + 1)}.replace({S(r, d + 2)}, {S(r, d + 2)}" + (f", {I(r, d + 1)})" if r.random() < 0.4 else ")")
    if k == 7:
        return f"{L(r, d + 1)}.join(" + (f"{_lit(r.choice(['', ',', '-', '::']))})" if r.random() < 0.7 else ")")
    if k == 8:
        return f"({S(r, d + 1)} + {S(r, d + 1)})"
    if k == 9:
        return f"{L(r, d + 1)}[{I(r, d + 1)}]"
    if k == 10:
        return f"({B(r, d + 1)} ? {S(r, d + 1)} : {S(r, d + 1)})"
    if k == 11:
        return f"base64.encode(bytes({S(r, d + 1)}))"
    if k == 12:
        return f"string({r.choice([I, I, S, S, B, B, I, D])(r, d + 1)})" if r.random() < 0.8 else f"string(bytes({S(r, d + 1)}))"
    return r.choice(ATTR_S)


def I(r, d=0):
    k = r.randrange(10 if d < 3 else 2)
    if k == 0:
        return str(r.choice([0, 1, 2, 3, 5, -1, 7, 100]))
    if k == 1:
        return r.choice(ATTR_I)
    if k == 2:
        return f"size({r.choice([S, L, N])(r, d + 1)})"
    if k == 3:
        return f"{S(r, d + 1)}.{r.choice(['indexOf', 'lastIndexOf'])}({S(r, d + 2)}" + (f", {I(r, d + 1)})" if r.random() < 0.3 else ")")
    if k == 4:
        return f"({I(r, d + 1)} {r.choice(['+', '-', '*', '%', '/'])} {I(r, d + 1)})"
    if k == 5:
        return f"{N(r, d + 1)}[{I(r, d + 1)}]"
    if k == 6:
        return f"({B(r, d + 1)} ? {I(r, d + 1)} : {I(r, d + 1)})"
    if k == 7:
        return f"-({I(r, d + 1)})"
    if k == 8:
        return f"int({r.choice([S(r, d + 1), I(r, d + 1)])})"
    return r.choice(ATTR_I)


def L(r, d=0):
    k = r.randrange(15 if d < 3 else 3)
    if k == 0:
        return "[" + ", ".join(_lit(r.choice(STR_LITS)) for _ in range(r.randrange(0, 4))) + "]"
    if k in (1, 2):
        return r.choice(ATTR_L)
    if k == 3:
        return f"{S(r, d + 1)}.split({_lit(r.choice([',', '', '.', 'l', ' ', 'ab']))}" + (f", {I(r, d + 1)})" if r.random() < 0.3 else ")")
    if k == 4:
        return f"({L(r, d + 1)} + {L(r, d + 1)})"
    if k == 5:
        return f"{L(r, d + 1)}.filter(x, {PX(r, d + 1)})"
    if k == 6:
        return f"{L(r, d + 1)}.map(x, {SX(r, d + 1)})"
    if k == 7:
        return f"{L(r, d + 1)}.map(x, {PX(r, d + 1)}, {SX(r, d + 1)})"
    if k == 8:
        return f"{r.choice(['except', 'intersect'])}({L(r, d + 1)}, {L(r, d + 1)})"
    if k == 9:
        return f"{L(r, d + 1)}.{r.choice(['sort', 'distinct', 'reverse', 'flatten'])}()"
    if k == 10:
        return f"{L(r, d + 1)}.slice({I(r, d + 1)}, {I(r, d + 1)})"
    if k == 11:
        return f"[{S(r, d + 1)}, {S(r, d + 1)}]"
    if k == 12:
        return f"{L(r, d + 1)}.sortBy(x, {r.choice(['x', 'size(x)', 'x.lowerAscii()'])})"
    if k == 13:
        return f"{L(r, d + 1)}.transformList(i, x, {r.choice(['x', 'x + x', 'x.upperAscii()'])})"
    return r.choice(ATTR_L)


def N(r, d=0):
    k = r.randrange(9 if d < 3 else 2)
    if k == 0:
        return "[" + ", ".join(str(r.choice([0, 1, 2, 3, -4, 10])) for _ in range(r.randrange(0, 4))) + "]"
    if k == 1:
        return r.choice(ATTR_N)
    if k == 2:
        return f"lists.range({I(r, d + 1)})"
    if k == 3:
        return f"{N(r, d + 1)}.{r.choice(['sort', 'distinct', 'reverse'])}()"
    if k == 4:
        return f"{N(r, d + 1)}.map(y, {r.choice(['y * 2', 'y + 1', 'y % 3', '-y'])})"
    if k == 5:
        return f"{N(r, d + 1)}.filter(y, y {r.choice(['>', '<', '==', '!='])} {I(r, d + 1)})"
    if k == 6:
        return f"{L(r, d + 1)}.map(x, size(x))"
    if k == 7:
        return f"({N(r, d + 1)} + {N(r, d + 1)})"
    return r.choice(ATTR_N)


ATTR_D = ["P.attr.d", "R.attr.e"]


def D(r, d=0):
    k = r.randrange(9 if d < 3 else 2)
    if k == 0:
        return r.choice(["0.0", "1.5", "-2.5", "0.5", "1e300", "3.0", "(0.0 / 0.0)", "(1.0 / 0.0)", "-0.5"])
    if k == 1:
        return r.choice(ATTR_D)
    if k == 2:
        return f"math.{r.choice(['ceil', 'floor', 'round', 'trunc', 'abs', 'sign', 'sqrt'])}({D(r, d + 1)})"
    if k == 3:
        return f"({D(r, d + 1)} {r.choice(['+', '-', '*', '/'])} {D(r, d + 1)})"
    if k == 4:
        return f"double({I(r, d + 1)})"
    if k == 5:
        return f"math.{r.choice(['greatest', 'least'])}({D(r, d + 1)}, {D(r, d + 1)})"
    if k == 6:
        return f"math.sqrt({I(r, d + 1)})"
    if k == 7:
        return f"math.{r.choice(['greatest', 'least'])}({r.choice(ATTR_N)})"
    return r.choice(ATTR_D)


def M(r, d=0):
    """predicates over ext.Math"""
    k = r.randrange(9)
    if k == 0:
        return f"math.{r.choice(['isNaN', 'isInf', 'isFinite'])}({D(r, d + 1)})"
    if k == 1:
        return f"{D(r, d + 1)} {r.choice(['==', '<', '>=', '!='])} {D(r, d + 1)}"
    if k == 2:
        return f"math.{r.choice(['greatest', 'least'])}({I(r, d + 1)}, {D(r, d + 1)}, {I(r, d + 1)}) {r.choice(['==', '<', '>'])} {r.choice([I, D])(r, d + 1)}"
    if k == 3:
        return f"math.{r.choice(['bitAnd', 'bitOr', 'bitXor'])}({I(r, d + 1)}, {I(r, d + 1)}) {r.choice(['==', '<', '>'])} {I(r, d + 1)}"
    if k == 4:
        return f"math.{r.choice(['bitShiftLeft', 'bitShiftRight'])}({I(r, d + 1)}, {I(r, d + 1)}) {r.choice(['==', '<', '>'])} {I(r, d + 1)}"
    if k == 5:
        return f"math.{r.choice(['abs', 'sign', 'bitNot'])}({I(r, d + 1)}) {r.choice(['==', '<', '>'])} {I(r, d + 1)}"
    if k == 6:
        return f"math.{r.choice(['greatest', 'least'])}({N(r, d + 1)}) {r.choice(['==', '<', '>'])} {I(r, d + 1)}"
    if k == 7:
        x = D(r, d + 1)
        return f"math.floor({x}) <= {x} && {x} <= math.ceil({x})"
    x = I(r, d + 1)
    return f"math.bitXor(math.bitNot({x}), {x}) == -1 && math.abs({x}) >= 0 && math.greatest({x}, 0 - {x}) == math.abs({x})"


def SX(r, d):
    """a string built from the comprehension variable x"""
    return r.choice(["x", "x.upperAscii()", 'x + "!"', "x.charAt(0)", "x.trim()", 'x.replace("a", "_")', "x.substring(1)", "x.reverse()"])


def PX(r, d):
    """a predicate over the comprehension variable x"""
    return r.choice(['x != ""', "size(x) > 1", 'x.startsWith("a")', 'x.contains("b")', f"x.matches({_lit(r.choice(PATTERNS))})",
                     f"x in {L(r, d + 1)}", f"x == {S(r, d + 1)}", 'x < "b"'])


def B(r, d=0):
    k = r.randrange(18 if d < 3 else 6)
    if k == 0:
        return f"{S(r, d + 1)} {r.choice(['==', '!=', '<', '>='])} {S(r, d + 1)}"
    if k == 1:
        return f"{I(r, d + 1)} {r.choice(['==', '!=', '<', '<=', '>', '>='])} {I(r, d + 1)}"
    if k == 2:
        return f"{L(r, d + 1)} {r.choice(['==', '!='])} {L(r, d + 1)}"
    if k == 3:
        return f"{S(r, d + 1)} in {L(r, d + 1)}"
    if k == 4:
        return f"{S(r, d + 1)}.matches({_lit(r.choice(PATTERNS))})"
    if k == 5:
        return f"{S(r, d + 1)}.{r.choice(['startsWith', 'endsWith', 'contains'])}({S(r, d + 1)})"
    if k == 6:
        return f"{N(r, d + 1)} == {N(r, d + 1)}"
    if k == 7:
        return f"{I(r, d + 1)} in {N(r, d + 1)}"
    if k == 8:
        return f"{L(r, d + 1)}.{r.choice(['exists', 'all', 'exists_one'])}(x, {PX(r, d + 1)})"
    if k == 9:
        return f"{r.choice(['hasIntersection', 'isSubset'])}({L(r, d + 1)}, {L(r, d + 1)})"
    if k == 10:
        return f"({B(r, d + 1)} {r.choice(['&&', '||'])} {B(r, d + 1)})"
    if k == 11:
        return f"!({B(r, d + 1)})"
    if k == 12:
        return f"size({L(r, d + 1)}) {r.choice(['==', '>', '<='])} {r.randrange(4)}"
    if k == 13:
        return f"{L(r, d + 1)}.transformMap(i, x, {r.choice(['x', 'i'])}) == {{}}"
    if k == 14:
        return f"base64.decode({S(r, d + 1)}) == bytes({S(r, d + 1)})"
    if k == 16:
        # cel.bind (ext.Bindings): the bound value used twice, once inside a comprehension
        kind = r.choice("SLI")
        if kind == "S":
            return f"cel.bind(bv, {S(r, d + 1)}, bv + bv == {S(r, d + 1)} || {L(r, d + 1)}.exists(x, x == bv) || size(bv) > 2)"
        if kind == "L":
            return f"cel.bind(bv, {L(r, d + 1)}, size(bv) > 1 && bv[0] in bv && bv.all(x, x in bv))"
        return f"cel.bind(bv, {I(r, d + 1)}, cel.bind(bw, bv + 1, bw > bv && bv * 2 - bv == bv))"
    if k == 15:
        # the same compound value on both sides: true unless its evaluation fails, so an error on one side only shows
        x = r.choice([S, I, L, N])(r, d)
        return f"{x} == {x}"
    return f"{S(r, d + 1)} == {S(r, d + 1)}"


def identity(r):
    """properties that hold for every value of the right type: both sides must find them true on the same requests, and
    fail on the same requests"""
    k = r.randrange(12)
    s, l, n, i = S(r, 1), L(r, 1), N(r, 1), I(r, 1)
    return [f"{s}.reverse().reverse() == {s}", f"{l}.reverse().reverse() == {l}", f"{l}.sort().sort() == {l}.sort()",
            f'{s}.split(",").join(",") == {s}', f"{s}.substring(0, {i}) + {s}.substring({i}) == {s}",
            f"{s}.lowerAscii().upperAscii() == {s}.upperAscii()", f"size({l}.distinct()) <= size({l})",
            f"size({s}.split(\"\")) == size({s})", f"{l}.slice(0, size({l})) == {l}", f"{n}.sort().reverse() == {n}.sort().reverse().distinct() || size({n}) >= 0",
            f"{s}.indexOf({s}.charAt({i})) <= {i}", f"({l} + {l}).distinct() == {l}.distinct()"][k]


def rand_attr(r, kind):
    """a value of the nominal kind nine times in ten, anything else otherwise (the type errors are part of the test)"""
    if r.random() < 0.1:
        kind = r.choice("silnbdmz")
    if kind == "s":
        return r.choice(STR_LITS + ["ABC", "a,b", "aGVsbG8=", "l", "ll", "hello", "  x  ", "a.b.c"])
    if kind == "i":
        return r.choice([0, 1, 2, 3, 5, -1, 7, 100, -3])
    if kind == "l":
        return [r.choice(STR_LITS + ["ab", "b", "hello"]) for _ in range(r.randrange(0, 5))]
    if kind == "n":
        return [r.choice([0, 1, 2, 3, -4, 10, 2.0, 2.5]) for _ in range(r.randrange(0, 5))]
    if kind == "b":
        return r.random() < 0.5
    if kind == "d":
        return r.choice([0.5, 2.0, -1.5, 1e10, 2.5, -0.5, 0.49999999999999994, 4503599627370497.5, 7.0])
    if kind == "m":
        return {"a": "b", "n": 1}
    return None


def rand_request(r):
    spec_p = [("s", "s"), ("u", "s"), ("n", "i"), ("l", "l"), ("tags", "l"), ("ids", "n"), ("d", "d")]
    spec_r = [("t", "s"), ("csv", "s"), ("k", "i"), ("m", "l"), ("nums", "n"), ("e", "d")]

    def attrs(spec):
        return {name: rand_attr(r, kind) for name, kind in spec if r.random() < 0.9}
    return {"requestId": "v", "principal": {"id": "p", "roles": ["user"], "attr": attrs(spec_p)},
            "resource": {"kind": "doc", "id": "d", "attr": attrs(spec_r)}}


# ---- timestamps and durations: texts that travel in request attributes (well formed, nearly well formed, malformed) ----------
ZONES = ["UTC", "+05:30", "-08:00", "America/New_York", "Europe/London", "Asia/Kolkata", "Australia/Lord_Howe", "+00:00", "-00:30", "Mars/Olympus"]
TS_FIELDS = ["getFullYear", "getMonth", "getDayOfYear", "getDayOfMonth", "getDate", "getDayOfWeek", "getHours", "getMinutes", "getSeconds", "getMilliseconds"]


def rand_ts_text(r):
    y = r.choice([1970, 1999, 2000, 2021, 2024, 2038, 2100, 2200, 1900, 1800, 1700, 1677, 1678, 2262, 2263, 1, 9999, 0])
    mo, dd = r.randrange(1, 13), r.randrange(1, 29)
    if r.random() < 0.15:
        mo, dd = r.choice([(2, 29), (2, 30), (4, 31), (12, 31), (13, 1), (0, 10), (1, 0), (6, 30)])
    h, mi, s = r.randrange(0, 24), r.randrange(0, 60), r.randrange(0, 60)
    if r.random() < 0.08:
        h, mi, s = r.choice([(24, 0, 0), (23, 60, 0), (23, 59, 60), (0, 0, 0)])
    frac = r.choice(["", "", ".5", ".021", ".123456789", ".000000001", ".1234567891", ".", ",5", ".999999999"])
    zone = r.choice(["Z", "Z", "+00:00", "-05:00", "+05:30", "+14:00", "-23:59", "+24:00", "z", "", "+0530", "+05", " UTC", "-00:00", "+23:60", "-24:60", "+25:00", "+00:61"])
    sep = r.choice(["T"] * 8 + ["t", " "])
    txt = f"{y:04d}-{mo:02d}-{dd:02d}{sep}{h:02d}:{mi:02d}:{s:02d}{frac}{zone}"
    k = r.random()
    if k < 0.05:
        txt = txt.replace("-", "/", 1)
    elif k < 0.08:
        txt = txt[: r.randrange(0, len(txt))]
    elif k < 0.1:
        txt = " " + txt
    elif k < 0.12:
        txt = f"{y}-{mo}-{dd}T{h}:{mi}:{s}Z"
    elif k < 0.14:
        txt = txt + "\n"
    elif k < 0.16:
        txt = txt.replace("0", "\u0660", 1)      # an ARABIC-INDIC DIGIT ZERO is a digit to Python's \\d, not to Go
    return txt


def rand_dur_text(r):
    if r.random() < 0.25:
        return r.choice(["\u0661s", "1s\n", "", "5", "1d", "+3s", ".5s", "1e3s", "-", "1h ", " 1h", "1H", "h", "1.s", "1..5s", "0", "+0", "-0", "1us", "1µs", "1μs",
                         "9223372036s", "9223372037s", "2562047h47m16.854775807s", "2562047h47m16.854775808s", "-2562047h47m16.854775808s",
                         "0.000000001s", "0.0000000001s", "1.5h30m", "1ns1ns", "3ms2s", "100000000000000000000h"])
    parts = []
    for _ in range(r.randrange(1, 4)):
        n = r.choice([str(r.randrange(0, 100)), f"{r.randrange(0, 100)}.{r.randrange(0, 1000)}", str(r.randrange(0, 100000))])
        parts.append(n + r.choice(["h", "m", "s", "ms", "us", "ns"]))
    return r.choice(["", "", "-", "+"]) + "".join(parts)


def TS(r, d=0):
    k = r.randrange(7 if d < 2 else 2)
    if k == 0:
        return f'timestamp("{rand_ts_text(r)}")'
    if k in (1, 2):
        return f"timestamp({r.choice(['R.attr.ts', 'P.attr.ts2'])})"
    if k == 3:
        return f"({TS(r, d + 1)} {r.choice(['+', '-'])} {DU(r, d + 1)})"
    if k == 4:
        return "now()"
    if k == 5:
        return f"timestamp({r.choice(['R.attr.secs', '0', '1700000000', '-1', '253402300800', 'int(R.attr.secs)'])})"
    return f"timestamp({r.choice(['R.attr.ts', 'P.attr.ts2'])})"


def DU(r, d=0):
    k = r.randrange(7 if d < 2 else 2)
    if k == 0:
        return f'duration("{rand_dur_text(r)}")'
    if k == 1:
        return f"duration({r.choice(['R.attr.dur', 'P.attr.dur2'])})"
    if k == 2:
        return f"({TS(r, d + 1)} - {TS(r, d + 1)})"
    if k == 3:
        return f"({DU(r, d + 1)} {r.choice(['+', '-'])} {DU(r, d + 1)})"
    if k == 4:
        return f"timeSince({TS(r, d + 1)})"
    if k == 5:
        return f"duration({r.choice(['R.attr.secs', '5', 'int(R.attr.secs)'])})"
    return f"duration({r.choice(['R.attr.dur', 'P.attr.dur2'])})"


def TB(r):
    k = r.randrange(8)
    if k == 0:
        return f"{TS(r)} {r.choice(['<', '<=', '>', '>=', '==', '!='])} {TS(r)}"
    if k == 1:
        return f"{DU(r)} {r.choice(['<', '<=', '>', '>=', '==', '!='])} {DU(r)}"
    if k == 2:
        tz = r.choice(ZONES)
        return f"{TS(r)}.{r.choice(TS_FIELDS)}({'' if r.random() < 0.4 else chr(34) + tz + chr(34)}) {r.choice(['==', '<', '>='])} {r.randrange(0, 32)}"
    if k == 3:
        return f"{DU(r)}.{r.choice(['getHours', 'getMinutes', 'getSeconds', 'getMilliseconds'])}() {r.choice(['==', '<', '>='])} {r.choice([0, 1, 59, 60, 3600, -1, 90])}"
    if k == 4:
        x = TS(r, 1)
        return f"({x} + duration(\"1h\")) - {x} == duration(\"60m\")"
    if k == 5:
        x = TS(r, 1)
        return f"{x}.getDayOfWeek() >= 0 && {x}.getDayOfYear() >= {x}.getDayOfMonth() && {x}.getDate() == {x}.getDayOfMonth() + 1"
    if k == 6:
        x = DU(r, 1)
        return f"{x}.getMinutes() == {x}.getSeconds() / 60 && {x}.getHours() == {x}.getMinutes() / 60"
    return f"int({TS(r)}) {r.choice(['==', '<', '>='])} {r.choice(['0', '1700000000', 'int(R.attr.secs)'])}"


def rand_time_request(r):
    def ts_or_other():
        k = r.random()
        return rand_ts_text(r) if k < 0.9 else r.choice([5, None, True, ["x"], "yesterday"])

    def dur_or_other():
        k = r.random()
        return rand_dur_text(r) if k < 0.9 else r.choice([5, None, "soon", 1.5])
    pa, ra = {}, {}
    if r.random() < 0.95:
        ra["ts"] = ts_or_other()
    if r.random() < 0.95:
        pa["ts2"] = ts_or_other()
    if r.random() < 0.95:
        ra["dur"] = dur_or_other()
    if r.random() < 0.95:
        pa["dur2"] = dur_or_other()
    if r.random() < 0.9:
        ra["secs"] = r.choice([0, 1, 1700000000, -1, 86400, 4102444800, 1.5, 253402300799, 253402300800, -62135596800, -62135596801, 9.3e9, "5"])
    return {"requestId": "t", "principal": {"id": "p", "roles": ["user"], "attr": pa}, "resource": {"kind": "doc", "id": "d", "attr": ra}}


# ---- core semantics: heterogeneous equality, cross-type ordering, conversions and their range errors, arithmetic overflow,
# ---- maps, has(), index errors, error absorption of && / || -- over attributes whose type changes from request to request
CORE_ATTRS = ["P.attr.a", "R.attr.b", "P.attr.c", "R.attr.m.k", 'R.attr.m["k2"]', "P.attr.lst[0]", "P.attr.lst[2]", "R.attr.m", "P.attr.lst", "R.attr.nested.x.y"]
CORE_LITS = ["1", "1u", "1.0", '"1"', "true", "null", "[1, 2.0]", '{"a": 1}', "0", "-1", "2", "9223372036854775807", "-9223372036854775807",
             "18446744073709551615u", "1e19", "-1e19", "0.5", '""', '"a"', "[]", "{}", "9007199254740993", "2.0", "3u", '["a", "b"]', "false", 'b"a"']


def CV(r, d=0):
    k = r.randrange(12 if d < 3 else 3)
    if k == 0:
        return r.choice(CORE_LITS)
    if k in (1, 2):
        return r.choice(CORE_ATTRS)
    if k == 3:
        return f"{r.choice(['int', 'uint', 'int', 'uint', 'dyn', 'bytes', 'int', 'uint', 'dyn', 'double', 'string'])}({CV(r, d + 1)})"
    if k == 4:
        return f"({CV(r, d + 1)} {r.choice(['+', '-', '*', '/', '%'])} {CV(r, d + 1)})"
    if k == 5:
        return f"-({CV(r, d + 1)})"
    if k == 6:
        return f"size({CV(r, d + 1)})"
    if k == 7:
        return f"({CB(r, d + 1)} ? {CV(r, d + 1)} : {CV(r, d + 1)})"
    if k == 8:
        return f"{r.choice(['R.attr.m', 'P.attr.lst', 'R.attr.nested.x'])}[{CV(r, d + 1)}]"
    if k == 9:
        return f"[{CV(r, d + 1)}, {CV(r, d + 1)}]"
    if k == 10:
        return f"{{{r.choice(['1', '\"a\"', 'true', '2u', 'P.attr.c'])}: {CV(r, d + 1)}}}"
    return r.choice(CORE_ATTRS)


def CB(r, d=0):
    k = r.randrange(14 if d < 3 else 5)
    if k in (0, 1):
        return f"{CV(r, d + 1)} {r.choice(['==', '!='])} {CV(r, d + 1)}"
    if k == 2:
        return f"{CV(r, d + 1)} {r.choice(['<', '<=', '>', '>='])} {CV(r, d + 1)}"
    if k == 3:
        return f"{CV(r, d + 1)} in {CV(r, d + 1)}"
    if k == 4:
        return f"has({r.choice(['R.attr.m.k', 'R.attr.m.zz', 'P.attr.a', 'R.attr.q', 'R.attr.nested.x.y', 'R.attr.nested.x.q', 'P.attr.lst', 'R.attr.b.c'])})"
    if k == 5:
        return f"({CB(r, d + 1)} {r.choice(['&&', '||'])} {CB(r, d + 1)})"
    if k == 6:
        return f"!({CB(r, d + 1)})"
    if k == 7:
        return f"R.attr.m.{r.choice(['exists', 'all', 'exists_one'])}(k, {r.choice(['k == \"k\"', 'R.attr.m[k] == 1', 'k != P.attr.c', 'size(k) > 1'])})"
    if k == 8:
        return f"R.attr.m.{r.choice(['exists', 'all'])}(k, v, {r.choice(['v == 1', 'k == \"k2\" && v != null', 'v in P.attr.lst', 'v > 0'])})"
    if k == 9:
        return f"P.attr.lst.{r.choice(['exists', 'all', 'exists_one'])}(x, {r.choice(['x == 1', 'x > 0', 'x in R.attr.m', 'x == P.attr.a', 'x != null'])})"
    if k == 10:
        x = CV(r, d + 1)
        return f"{x} == {x}"
    if k == 11:
        return f"({CB(r, d + 1)} ? {CB(r, d + 1)} : {CB(r, d + 1)})"
    if k == 12:
        tn = r.choice(["string", "int", "uint", "double", "bool", "list", "map", "null_type", "bytes", "type"])
        return r.choice([f"type({CV(r, d + 1)}) {r.choice(['==', '!='])} {tn}", f"type({CV(r, d + 1)}) == type({CV(r, d + 1)})", f"bool({CV(r, d + 1)})"])
    return f"{CV(r, d + 1)} == {CV(r, d + 1)}"


def rand_core_value(r, depth=0):
    k = r.randrange(12 if depth < 2 else 8)
    if k == 0:
        return r.choice([0, 1, 2, -1, 3, 100, 2.0])
    if k == 1:
        return r.choice([0.5, -0.5, 1e19, -1e19, 9007199254740993, 9223372036854775807, 9223372036854775808, -9223372036854775808, 1.5, 4294967296, -0.0, 1e-7])
    if k in (2, 3):
        return r.choice(["1", "a", "", "k", "k2", "true", "1.5", "abc", "-1", "9223372036854775808", "0x10", " 1", "1e3", "日本", "\u0661\u0662", "12\n", "+7", "-0",
                         "1_000", "t", "F", "TRUE", "False", "18446744073709551615", "18446744073709551616", "-9223372036854775808", "007"])
    if k == 4:
        return r.random() < 0.5
    if k == 5:
        return None
    if k == 6:
        return 1
    if k == 7:
        return "a"
    if k in (8, 9):
        return [rand_core_value(r, depth + 1) for _ in range(r.randrange(0, 4))]
    return {kk: rand_core_value(r, depth + 1) for kk in r.sample(["k", "k2", "a", "1", "x", "y"], r.randrange(0, 4))}


def rand_core_request(r):
    pa = {n: rand_core_value(r) for n in ("a", "c") if r.random() < 0.9}
    ra = {n: rand_core_value(r) for n in ("b",) if r.random() < 0.9}
    if r.random() < 0.9:
        pa["lst"] = [rand_core_value(r, 1) for _ in range(r.randrange(0, 4))] if r.random() < 0.9 else rand_core_value(r)
    if r.random() < 0.9:
        ra["m"] = {kk: rand_core_value(r, 1) for kk in r.sample(["k", "k2", "a", "zz"], r.randrange(0, 4))} if r.random() < 0.9 else rand_core_value(r)
    if r.random() < 0.8:
        ra["nested"] = {"x": {"y": rand_core_value(r, 1)}} if r.random() < 0.8 else rand_core_value(r)
    return {"requestId": "c", "principal": {"id": "p", "roles": ["user"], "attr": pa}, "resource": {"kind": "doc", "id": "d", "attr": ra}}


# ---- inIPAddrRange (cerbos_lib.go:472-510 over Go's net.ParseIP / ParseCIDR / IPNet.Contains) ------------------------------------
CIDRS = ["10.0.0.0/8", "10.1.2.0/24", "192.168.0.0/16", "0.0.0.0/0", "10.1.2.3/32", "::/0", "2001:db8::/32", "::ffff:10.0.0.0/104", "::ffff:0:0/96",
         "fe80::/10", "2001:db8:0:1::/64", "::1/128", "10.0.0.0/33", "10.0.0/8", "2001:db8::/129", "10.0.0.0", "10.0.0.0/08", "::ffff:10.1.2.0/120",
         "10.0.0.0/ 8", "1.2.3.4/0", "::ffff:10.0.0.0/95"]


def rand_ip_text(r):
    k = r.random()
    if k < 0.35:
        parts = [r.choice([0, 1, 2, 3, 10, 168, 192, 255, 256, 127]) for _ in range(4)]
        if r.random() < 0.5:
            parts[0:2] = r.choice([[10, 1], [10, 0], [192, 168]])
        txt = ".".join(str(p) for p in parts)
        m = r.random()
        if m < 0.06:
            txt = txt.replace(".", ".0", 1)
        elif m < 0.1:
            txt = txt.rsplit(".", 1)[0]
        elif m < 0.13:
            txt += ".1"
        elif m < 0.16:
            txt = " " + txt
        elif m < 0.19:
            txt = txt.replace("1", "١", 1)
        return txt
    if k < 0.55:
        v4 = ".".join(str(r.choice([10, 1, 2, 3, 0, 255, 192, 168])) for _ in range(4))
        return r.choice(["::ffff:", "::FFFF:", "0:0:0:0:0:ffff:", "::", "::ffff:0:", "64:ff9b::"]) + v4
    if k < 0.9:
        hx = [r.choice(["2001", "db8", "0", "1", "fe80", "ffff", "ABCD", "0001", "00001", "g", ""]) for _ in range(r.choice([8, 8, 8, 7, 9, 4]))]
        txt = ":".join(hx)
        m = r.random()
        if m < 0.4:
            i = r.randrange(0, 6)
            txt = ":".join(hx[:i]) + "::" + ":".join(hx[i + 2:])
        if m > 0.9:
            txt += "%eth0"
        return txt
    return r.choice(["", "::", "::1", "localhost", "1.2.3", "1.2.3.4.5", ":::", "1::2::3", "::ffff:1.2.3", "2001:db8::1", "fe80::1%1", "[::1]", "0x10.1.2.3",
                     "010.1.2.3", "1.2.3.4/8", "::ffff:10.1.2.3", "2001:DB8:0:1:0:0:0:1"])


def IPB(r):
    ip = r.choice(["R.attr.ip", "P.attr.ip2", "R.attr.ip", f'"{rand_ip_text(r)}"'])
    return f'{ip}.inIPAddrRange("{r.choice(CIDRS)}")' if r.random() < 0.5 else f'inIPAddrRange({ip}, "{r.choice(CIDRS)}")'


def rand_ip_request(r):
    def v():
        return rand_ip_text(r) if r.random() < 0.93 else r.choice([5, None, ["10.0.0.1"], True])
    return {"requestId": "i", "principal": {"id": "p", "roles": ["user"], "attr": {"ip2": v()}}, "resource": {"kind": "doc", "id": "d", "attr": {"ip": v()}}}
