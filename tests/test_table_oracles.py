"""Oracle #2 (C interpreter of the flattened table) and the host build of the kernel core against
oracle #1 (structural Python restatement), on the reference goldens and the synthetic workloads."""
import numpy as np
import pytest

import workloads as W
from cerbos_b200.encode import Encoder
from cerbos_b200.policy.compile import build_rule_table
from cerbos_b200.table import layout as L
from cerbos_b200.table.bytecode import Unsupported
from cerbos_b200.table.flatten import flatten
from helpers import engine_decisions, load_golden, store_rule_table
from hostsim import driver as hostsim
from oracle import cref
from oracle.celeval import parse_timestamp
from oracle.check import CheckOracle

G = {"environment": "test"}
NOW = parse_timestamp("2024-01-01T00:00:00Z")


@pytest.fixture(scope="module")
def store_flat():
    return flatten(store_rule_table(), globals_=G)


def test_layout_header_in_sync():
    import os
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "cerbos_b200_format.h")
    assert open(hdr).read() == L.c_header(), "run `python -m cerbos_b200.table.layout`"


def test_c_oracle_and_kernel_core_on_engine_goldens(store_flat):
    ft = store_flat
    orc = CheckOracle(store_rule_table(), globals_=G)
    n = 0
    for cid, lenient, inp, want in engine_decisions():
        enc = Encoder(ft.manifest, lenient_scope_search=lenient)
        b = enc.encode([inp])
        fl = L.BATCH_FLAG_LENIENT if lenient else 0
        c_out = cref.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, fl)
        k_out = hostsim.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, fl)
        py = orc.check(inp, NOW, lenient=lenient)
        for k, a in enumerate(inp["actions"]):
            assert c_out[0, k] == py["actions"][a]["effect"], (cid, a)
            assert k_out[0, k] == py["actions"][a]["effect"], (cid, a)
            n += 1
    assert n == 166


def test_batched_goldens_mixed_shapes(store_flat):
    """All strict golden inputs in ONE batch: mixed action counts, role counts, scopes, principals."""
    ft = store_flat
    inputs = [inp for _, lenient, inp, _ in engine_decisions() if not lenient]
    b = Encoder(ft.manifest).encode(inputs)
    c_out = cref.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns)
    k_out = hostsim.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns)
    valid = c_out != 0
    assert (c_out[valid] == k_out[valid]).all()
    one = np.concatenate([cref.check(ft.blob, Encoder(ft.manifest).encode([i]).columns, 1, len(i["actions"]), NOW.ns)[0]
                          for i in inputs])
    assert (c_out[valid] == one).all()


def _leafs(c):
    if "expr" in c:
        yield c["expr"]
    for k in ("all", "any", "none"):
        if k in c:
            for x in c[k]["of"]:
                yield from _leafs(x)


def _cel_cases():
    for tc in load_golden("cel_eval.json"):
        for e in _leafs(tc["condition"]):
            yield tc["file"], e, tc["request"]
    for tc in load_golden("cerbos_lib_test.json"):
        yield "cerbos_lib_test", tc["expr"], {"principal": {"id": "x", "roles": ["r"]}, "resource": {"kind": "k", "id": "1"}}


def test_bytecode_vs_cel_oracle_on_golden_expressions():
    """Every golden CEL leaf that lowers to bytecode must evaluate like oracle #1 (the rest must be
    rejected at table build -- never silently diverge)."""
    now = parse_timestamp("2021-04-22T10:05:20.021-05:00")
    lowered = run_time_values = 0
    for f, e, req in _cel_cases():
        inp = {"principal": dict(req.get("principal") or {}), "resource": dict(req.get("resource") or {}), "actions": ["a"]}
        if "auxData" in req:
            inp["auxData"] = req["auxData"]
        inp["resource"]["kind"] = "leave_request"
        inp["principal"].setdefault("roles", ["r"])
        pol = {"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {"resource": "leave_request", "version": "default",
               "rules": [{"actions": ["a"], "effect": "EFFECT_ALLOW", "roles": ["*"], "condition": {"match": {"expr": e}}}]}}
        try:
            rt = build_rule_table([pol])
            ft = flatten(rt)
        except Unsupported:
            continue
        except Exception:
            continue
        lowered += 1
        b = Encoder(ft.manifest).encode([inp])
        want = CheckOracle(rt).check(inp, now)["actions"]["a"]["effect"]
        assert hostsim.check(ft.blob, b.columns, 1, 1, now.ns)[0, 0] == want, (f, e)
        try:
            c_out = cref.check(ft.blob, b.columns, 1, 1, now.ns)[0, 0]
        except RuntimeError as x:
            # oracle #2 does not port the functions that build lists / strings at run time (it flags them): those
            # expressions are pinned by oracle #1 (which the reference goldens pin) against the kernel core only
            assert "-2" in str(x), (f, e, x)
            run_time_values += 1
            continue
        assert c_out == want, (f, e)
    assert lowered == 208 and run_time_values <= 70, (lowered, run_time_values)   # every golden leaf lowers, the 18 SPIFFE ones included


RUN_TIME_VALUE_CASES = [
    'P.attr["1-2-3"].transformMapEntry(indexVar, valueVar, {valueVar: indexVar}) == {1: 0, 2: 1, 3: 2}', '{2.0: 5}[2] == 5',
    'P.attr.s.trim() == "héllo wörld"', 'P.attr.s.charAt(1) == "é"', 'P.attr.s.indexOf("ö") == 7', 'P.attr.s.lastIndexOf("l") == 10',
    'P.attr.s.substring(1, 5) == "éllo"', 'P.attr.s.upperAscii() == "HéLLO WöRLD  "', 'P.attr.s.replace("l", "L", 2) == "héLLo wörld  "',
    'P.attr.s.replace("", "-", 3) == "-h-é-llo wörld  "', 'R.attr.csv.split(",") == ["a","b","","c"]', 'R.attr.csv.split(",", 2) == ["a","b,,c"]',
    'R.attr.csv.split("") == ["a",",","b",",",",","c"]', 'P.attr.e.split(",") == [""]', 'R.attr.csv.split(",").join("-") == "a-b--c"',
    'P.attr.teams.map(t, t.upperAscii()).sort() == ["COMMERCIAL","COMMUNICATIONS","DESIGN","PRODUCT"]',
    'P.attr.teams.filter(t, t.startsWith("co")).reverse() == ["commercial","communications"]', 'P.attr.teams.except(["design"]).size() == 3',
    'P.attr.teams.map(t, t.size() > 7, t + "!") == ["communications!", "commercial!"]', '(P.attr.teams + ["x"]).slice(3, 5) == ["commercial", "x"]',
    'lists.range(4).map(i, i * i) == [0, 1, 4, 9]', '[[1,2],[3],[4,[5]]].flatten() == [1,2,3,4,[5]]', '[3,1,2,3].distinct().sort() == [1,2,3]',
    'hierarchy(R.attr.csv.split(",").filter(x, x != "")).ancestorOf(hierarchy("a.b.c.d"))', 'hierarchy("a:b:c", ":")[2] == "c"',
    'hierarchy(P.attr.department)[0] == "marketing"', 'hierarchy(P.attr.department)[1] == "x"', 'P.attr.s.reverse() == "  dlröw olléh"',
    'P.attr.teams.transformMap(i, t, i % 2 == 0, t.charAt(0)) == {0: "d", 2: "p"}', '"abc".lastIndexOf("") == 3', '"abc".indexOf("c", 3) == -1',
    'P.attr.teams.exists(t, t.lowerAscii() == "DESIGN".lowerAscii())', 'intersect(P.attr.teams, ["product","x","design","q","z"]) == ["design","product"]',
    'P.attr.s.charAt(99) == "x"', '[1, "a"].sort() == []', 'P.attr.teams.map(t, t.nope)  == []', '[P.id, R.attr.owner] == ["john", "john"]',
    '{P.id: R.attr.csv}.john == "a,b,,c"', '{P.id: 1, R.attr.owner: 2} == {}', 'P.attr.teams.transformList(i, t, t.size() + i).sort() == [6, 9, 12, 15]',
    '"a,b".split(",", 0) == []', 'P.attr.teams.join() == "designcommunicationsproductcommercial"', '[1, 2].join(",") == "1,2"',
    'hierarchy(["a", "b"]) == hierarchy("a.b")', 'hierarchy(["a.b", "c"]) == hierarchy("a.b.c")', 'hierarchy(["a", "b"]).siblingOf(hierarchy("a:c", ":"))',
    # RE2 search through a DFA table built at table load (cel/regex_dfa.py)
    'P.attr.department.matches("^[mM].*g$")', 'size(P.attr.teams.filter(t, t.matches("^comm"))) == 2', 'P.attr.s.matches("^h.llo w.rld\\\\s+$")',
    'R.attr.csv.matches("^(\\\\w?,)+\\\\w$")', 'P.attr.e.matches("^$")', 'P.attr.teams.all(t, t.matches("^[a-z]+$"))', 'P.attr.department.matches("[")',
    'R.attr.owner.matches("^jo(hn|e)$") && !R.attr.owner.matches("x")', 'P.attr.teams.matches("a")',
    # bytes / base64
    'base64.decode("aGVsbG8=") == bytes("hello")', 'base64.encode(bytes("hello")) == "aGVsbG8="', 'base64.decode(R.attr.b64) == b"hello"',
    'base64.encode(bytes(P.attr.department)) == "bWFya2V0aW5n"', 'base64.decode("a") == b""', 'base64.decode("aGVsbG8h") == b"hello!"',
    'size(bytes("héllo")) == 6', 'base64.decode("aGVsbG9=") == b"x"',
    # IANA zones (transition table built from the host's tz database at table load)
    'timestamp(R.attr.lastAccessed).getMonth("NZ") == 3', 'timestamp(R.attr.lastAccessed).getHours("NZ") == 3',
    'timestamp(R.attr.lastAccessed).getHours("America/New_York") == 11', 'timestamp(R.attr.lastAccessed).getDayOfWeek("Asia/Kolkata") == 2',
    'timestamp(R.attr.lastAccessed).getHours("Mars/Olympus") == 1',
    'timestamp("2021-12-25T00:30:00Z").getDate("Europe/London") == 25 && timestamp("2021-06-25T23:30:00Z").getDate("Europe/London") == 26',
    # sortBy
    'P.attr.people.sortBy(e, e.score).map(e, e.name) == ["bar", "foo", "baz"]', 'P.attr.teams.sortBy(t, t.size()) == ["design", "product", "commercial", "communications"]',
    'P.attr.teams.sortBy(t, t) == ["commercial", "communications", "design", "product"]', 'P.attr.people.sortBy(e, e.nope) == []',
]
RUN_TIME_VALUE_REQUEST = {"principal": {"attr": {"1-2-3": [1, 2, 3], "department": "marketing", "teams": ["design", "communications", "product", "commercial"],
                                                 "s": "héllo wörld  ", "e": "",
                                                 "people": [{"name": "foo", "score": 0}, {"name": "bar", "score": -10}, {"name": "baz", "score": 1000}]},
                                        "id": "john", "roles": ["employee"]},
                          "resource": {"attr": {"owner": "john", "csv": "a,b,,c", "lastAccessed": "2021-04-20T10:00:20.021-05:00", "b64": "aGVsbG8"},
                                       "id": "test", "kind": "leave_request"}, "actions": ["a"]}


def run_time_value_table(e):
    pol = {"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {"resource": "leave_request", "version": "default",
           "rules": [{"actions": ["a"], "effect": "EFFECT_ALLOW", "roles": ["*"], "condition": {"match": {"expr": e}}}]}}
    rt = build_rule_table([pol])
    return rt, flatten(rt)


def test_run_time_values_vs_cel_oracle():
    """List / string producing functions, collecting comprehensions, dynamic literals, concatenation, hierarchy(list) and
    hierarchy[i] (device scratch arena; cerbos_lib.go:287, 433, cel-go ext.Strings / ext.Lists): the kernel core must
    evaluate every case like oracle #1, including the error cases (index out of range, mixed-type sort, missing field)."""
    now = parse_timestamp("2021-04-22T10:05:20.021-05:00")
    for e in RUN_TIME_VALUE_CASES:
        rt, ft = run_time_value_table(e)
        b = Encoder(ft.manifest).encode([RUN_TIME_VALUE_REQUEST])
        want = CheckOracle(rt).check(RUN_TIME_VALUE_REQUEST, now)["actions"]["a"]["effect"]
        assert hostsim.check(ft.blob, b.columns, 1, 1, now.ns)[0, 0] == want, e


# (expression, holds): `holds` None = the evaluation fails, so the expression AND its negation are both denied
MATH_AND_ENCODER_CASES = [
    # cel-go ext.Math (ext/math.go) beyond the documentation rows: mixed numeric types, ties, NaN, overflow, shifts
    ("math.greatest(1, 2.5, 2u) == 2.5", True), ("math.least([3, 1.0, 2]) == 1.0", True), ("math.greatest(R.attr.nums) == 7", True),
    ("math.least(R.attr.nums) == -2.5", True), ("math.greatest(2, 2.0) == 2 && math.least(2u, 2) == 2u", True),
    ("math.greatest([]) == 0", None), ("math.greatest(0.0 / 0.0, 1.0) == 1.0", None), ("math.least(R.attr.mixed) == 1", None),
    ("math.greatest(P.attr.n) == 12 && math.least(P.attr.x) == 2.5", True), ("math.greatest(P.attr.department) == 1", None),
    ("math.round(0.49999999999999994) == 0.0 && math.round(2.5) == 3.0 && math.round(-2.5) == -3.0", True),
    ("math.round(4503599627370497.0) == 4503599627370497.0 && math.trunc(-0.9) == 0.0", True), ("math.ceil(1) == 1.0", None),
    ("math.isNaN(math.sqrt(-1.0)) && math.sqrt(81) == 9.0 && math.sqrt(2u) > 1.41 && math.sqrt(P.attr.x) > 1.58", True),
    ("math.isInf(1e308 * 10.0) && !math.isFinite(-1.0 / 0.0) && math.isFinite(P.attr.x)", True), ("math.isNaN(1)", None),
    ("math.abs(-9223372036854775807 - 1) > 0", None), ("math.abs(int(P.attr.n) - 20) == 8 && math.abs(5u) == 5u && math.abs(-0.5) == 0.5", True),
    ("math.sign(-0.0) == 0.0 && math.sign(7u) == 1u && math.sign(int(P.attr.n) - 20) == -1 && math.isNaN(math.sign(0.0 / 0.0))", True),
    ("math.bitShiftRight(-8, 1) == 9223372036854775804 && math.bitShiftLeft(1, 63) < 0 && math.bitShiftLeft(int(P.attr.n), 64) == 0", True),
    ("math.bitShiftLeft(1, -1) == 0", None), ("math.bitShiftLeft(1, 1u) == 2", None), ("math.bitAnd(1, 1u) == 1", None),
    ("math.bitNot(int(P.attr.n)) == -13 && math.bitXor(int(P.attr.n), 5) == 9 && math.bitOr(12u, 3u) == 15u", True), ("math.bitNot(1.0) == 0", None), ("math.bitNot(P.attr.n) == -13", None),
    # string(x) (cel-go ConvertToType): a JSON number is a double -- an integral one prints without a fraction
    ('string(12) == "12" && string(-5) == "-5" && string(7u) == "7" && string(true) == "true" && string("x") == "x"', True),
    ('string(P.attr.n) == "12" && string(int(P.attr.n) * -1000000) == "-12000000" && "id-" + string(P.attr.n) == "id-12"', True),
    ('string(0.0 / 0.0) == "NaN" && string(1.0 / 0.0) == "+Inf" && string(-1.0 / 0.0) == "-Inf" && string(P.attr.n * -0.0) == "-0"', True),
    ('string(b"abc") == "abc" && string(bytes("héllo")) == "héllo" && string(base64.decode("aGVsbG8=")) == "hello"', True),
    ('string(base64.decode("/w==")) == "x"', None), ('string(base64.decode("7aCA")) == "x"', None), ('string(R.attr.nums) == "x"', None),
    # bool(x) = strconv.ParseBool on strings; type(x) and the type names; cel.bind (ext.Bindings)
    ('bool("t") && bool("TRUE") && bool("1") && !bool("F") && !bool("false") && !bool("0") && bool(true)', True), ('bool("tRUE")', None),
    ('bool("yes")', None), ('bool("")', None), ('bool(1)', None), ('bool(P.attr.department)', None), ('bool(R.attr.yes) && !bool(R.attr.no)', True),
    ('type(P.attr.n) == double && type(P.attr.department) == string && type(R.attr.nums) == list && type(P.attr) == map', True),
    ('type(1) == int && type(1u) == uint && type(b"a") == bytes && type(null) == null_type && type(type(1)) == type && type(true) == bool', True),
    ('type(P.attr.n) != int && string != bytes && type(P.attr.nope) == string', None), ('type(P.attr.n) == type(P.attr.x) && type(P.attr.n) != type(1)', True),
    ('cel.bind(d, P.attr.department, d.startsWith("mark") && d.size() == 9 && cel.bind(e, d + d, e.size() == 18))', True),
    ('cel.bind(z, P.attr.nope, true)', True), ('cel.bind(z, P.attr.nope, z == 1)', None), ('cel.bind(xs, R.attr.nums, xs.all(v, v in xs) && xs[0] == 3)', True),
    # base64: Go's decoder skips CR / LF, drops non-zero trailing bits, takes unpadded text, refuses misplaced padding
    ('size(base64.decode("x1")) == 1 && base64.decode("aGVsbG9=") == b"hello"', True), ('base64.decode("aGVs\\nbG8=\\r\\n") == b"hello"', True),
    ('size(base64.decode("aaGVsbG8=")) >= 0', None), ('size(base64.decode("abcaGVsbG8=")) >= 0', None), ('size(base64.decode("a")) >= 0', None),
    ('size(base64.decode("ab=c")) >= 0', None), ('size(base64.decode("ab==")) == 1 && size(base64.decode("abc")) == 2', True),
    ('size(base64.decode(R.attr.b64 + "=")) == 5 && size(base64.decode(R.attr.b64 + "==")) >= 0', None),
]
MATH_REQUEST = {"principal": {"id": "john", "roles": ["employee"], "attr": {"n": 12, "x": 2.5, "department": "marketing"}},
                "resource": {"kind": "leave_request", "id": "r", "attr": {"nums": [3, -2.5, 7, 0], "mixed": [1, "a"], "b64": "aGVsbG8", "yes": "True", "no": "FALSE"}}, "actions": ["a"]}


def test_math_and_encoder_functions_vs_cel_oracle():
    """ext.Math and the base64 edge cases: the expected answer is written out, oracle #1 must give it and the kernel core must
    give oracle #1's -- on the expression and on its negation, so that a failed evaluation (both denied) differs from `false`."""
    now = parse_timestamp("2021-04-22T10:05:20.021-05:00")
    programs = 0
    for e, holds in MATH_AND_ENCODER_CASES:
        for neg in (False, True):
            expr = f"!({e})" if neg else e
            rt, ft = run_time_value_table(expr)
            b = Encoder(ft.manifest).encode([MATH_REQUEST])
            want = CheckOracle(rt).check(MATH_REQUEST, now)["actions"]["a"]["effect"]
            assert want == (1 if holds is not None and holds != neg else 2), (expr, want)
            assert hostsim.check(ft.blob, b.columns, 1, 1, now.ns)[0, 0] == want, expr
            src, _ = hostsim.generate_uc(ft.blob)           # the generated leaf programs take the same functions
            programs += "CB_HD bool uc_atom_" in src        # (a bind whose body has a flat form needs none)
    assert programs >= 2 * len(MATH_AND_ENCODER_CASES) - 8


def test_string_of_a_fractional_double_is_flagged():
    """string(2.5) needs shortest-digit printing (strconv 'f', -1): the device flags the request (the call fails loudly)
    instead of approximating; oracle #1 prints it."""
    rt, ft = run_time_value_table('string(P.attr.x) == "2.5"')
    b = Encoder(ft.manifest).encode([MATH_REQUEST])
    assert CheckOracle(rt).check(MATH_REQUEST)["actions"]["a"]["effect"] == 1
    with pytest.raises(RuntimeError, match="-2"):
        hostsim.check(ft.blob, b.columns, 1, 1)


def test_c5_vectorised_columns_answer_like_the_generic_encoder():
    """workloads.C5.columns (numpy, what the bench generates 2^23 requests per GPU with) against the generic encoder over the
    same requests as protojson dicts: same headers and roles, same heap volume, and -- the heap and the batch dictionary
    are laid out in another order -- the same 8 decisions per request from oracle #2; chunked generation (columns_parallel,
    heap references rebased incl. the lists nested in the grants maps) gives the same answers again."""
    w = W.C5()
    _, ft, enc = W.build(w)
    n = 5000
    f = w.fields(n, start=777)
    b = w.columns(f, enc)
    b2 = enc.encode(w.inputs(f, range(n)))
    assert (np.asarray(b.columns[0])[:, 1:] == np.asarray(b2.columns[0])[:, 1:]).all() and (np.asarray(b.columns[2]) == np.asarray(b2.columns[2])).all()
    assert len(b.columns[4]) == len(b2.columns[4])
    want = cref.check(ft.blob, b2.columns, b2.n, b2.max_actions, n_threads=4)
    assert (cref.check(ft.blob, b.columns, b.n, b.max_actions, n_threads=4) == want).all()
    assert 0.3 < (want == 1).mean() < 0.5
    b3 = W.columns_parallel(w, n, 777, enc, chunk=1024)
    assert (cref.check(ft.blob, b3.columns, b3.n, b3.max_actions, n_threads=4) == want).all()
    assert (hostsim.check(ft.blob, b3.columns, b3.n, b3.max_actions, mode=1)[:512] == want[:512]).all()


SPIFFE_IDS = ["spiffe://cerbos.dev/ns/privileged/sa/curl", "spiffe://cerbos.dev", "spiffe://example.com/a", "spiffe://Cerbos.dev/x", "spiffe://cerbos.dev/",
              "spiffe://cerbos.dev//a", "spiffe://cerbos.dev/./a", "spiffe://cerbos.dev/a/..", "spiffe://cerbos.dev/a b", "http://cerbos.dev/x", "", "spiffe:///x",
              "spiffe://cerbos.dev/A_b-c.d/e", "spiffe://a_b-c.1/x", "cerbos.dev"]
SPIFFE_CASES = [
    'spiffeID(P.id).path() == "/ns/privileged/sa/curl"', 'spiffeID(P.id).path() == ""', 'spiffeID(P.id).trustDomain() == spiffeTrustDomain("cerbos.dev")',
    'spiffeID(P.id).trustDomain().name() == "cerbos.dev"', 'spiffeID(P.id).trustDomain().id() == "spiffe://cerbos.dev"',
    'spiffeID(P.id).isMemberOf(spiffeTrustDomain(R.attr.td))', 'spiffeID(P.id).isMemberOf(spiffeTrustDomain("spiffe://cerbos.dev/some/path"))',
    'spiffeID(P.id) == P.id', 'P.id == spiffeID(P.id)', 'spiffeID(P.id) != "spiffe://cerbos.dev"', 'spiffeID(P.id) == spiffeID(R.attr.other)',
    'spiffeTrustDomain(P.id) == "cerbos.dev"', 'spiffeTrustDomain(P.id) == "spiffe://cerbos.dev/any"', 'spiffeTrustDomain(P.id) == "not a domain!"',
    'spiffeTrustDomain(P.id) != spiffeTrustDomain(R.attr.other)', 'spiffeTrustDomain(R.attr.td).id() == "spiffe://cerbos.dev"',
    'spiffeMatchAny().matchesID(P.id)', '!spiffeMatchAny().matchesID(P.id)', 'spiffeMatchAny().matchesID(spiffeID(P.id))',
    'spiffeMatchExact(R.attr.other).matchesID(P.id)', 'spiffeMatchExact(spiffeID(R.attr.other)).matchesID(spiffeID(P.id))',
    'spiffeMatchOneOf([R.attr.other, "spiffe://cerbos.dev/ns/privileged/sa/curl"]).matchesID(P.id)',
    'spiffeMatchOneOf([spiffeID(R.attr.other), spiffeID("spiffe://cerbos.dev")]).matchesID(spiffeID(P.id))',
    'spiffeMatchOneOf(R.attr.ids).matchesID(P.id)', 'spiffeMatchOneOf(["not an id"]).matchesID(P.id)',
    'spiffeMatchTrustDomain(R.attr.td).matchesID(P.id)', 'spiffeMatchTrustDomain(spiffeTrustDomain("example.com")).matchesID(P.id)',
    '!spiffeMatchTrustDomain("spiffe://example.com").matchesID(P.id)', 'spiffeID(P.id).isMemberOf("cerbos.dev")', 'spiffeID(R.attr.n) == "x"',
]


def test_spiffe_functions_vs_cel_oracle():
    """conditions/types/spiffe.go on the kernel core against oracle #1 (which the reference's TestCerbosLib rows, cel_eval goldens
    and documentation rows pin): valid and malformed ids (go-spiffe's FromString / ValidatePath rules: character sets, empty
    / dot segments, trailing slash, missing trust domain, wrong scheme), equality in both operand orders, every matcher."""
    now = parse_timestamp("2021-04-22T10:05:20.021-05:00")
    n = allowed = 0
    for e in SPIFFE_CASES:
        rt, ft = run_time_value_table(e)
        orc = CheckOracle(rt)
        enc = Encoder(ft.manifest)
        for pid in SPIFFE_IDS:
            for other in (SPIFFE_IDS[0], SPIFFE_IDS[2], "bad id"):
                inp = {"principal": {"id": pid, "roles": ["api"]}, "actions": ["a"],
                       "resource": {"kind": "leave_request", "id": "r", "attr": {"td": "cerbos.dev", "other": other, "ids": [other, SPIFFE_IDS[1]], "n": 5}}}
                b = enc.encode([inp])
                want = orc.check(inp, now)["actions"]["a"]["effect"]
                assert hostsim.check(ft.blob, b.columns, 1, 1, now.ns)[0, 0] == want, (e, pid, other)
                n += 1
                allowed += want == 1
    assert n == len(SPIFFE_CASES) * len(SPIFFE_IDS) * 3 and 150 < allowed < n - 150, (n, allowed)


@pytest.mark.parametrize("cls,n", [(W.C1, 1024), (W.C2, 1 << 14), (W.C3, 1 << 12)])
def test_workloads_three_way(cls, n):
    w = cls()
    rt, ft, enc = W.build(w)
    f = w.fields(n)
    b = w.columns(f, enc)
    c_out = cref.check(ft.blob, b.columns, b.n, b.max_actions, n_threads=4)
    for mode in (0, 1, 2, 3):   # fast body / general 64-bit / general 32-bit / fast body on staged column tiles
        k_out = hostsim.check(ft.blob, b.columns, b.n, b.max_actions, mode=mode)
        assert (c_out == k_out).all(), mode
    idx = list(range(0, n, max(1, n // 256)))
    inputs = w.inputs(f, idx)
    b2 = enc.encode(inputs)
    assert (cref.check(ft.blob, b2.columns, b2.n, b2.max_actions) == c_out[idx]).all(), "vectorised columns != encoder"
    orc = CheckOracle(rt)
    for j, inp in enumerate(inputs[:256]):
        g = orc.check(inp)
        for k, a in enumerate(w.actions):
            assert g["actions"][a]["effect"] == c_out[idx[j], k]


def test_many_actions_and_roles_passes():
    """K x role_cols > 64 forces several passes of the bit-parallel walk."""
    docs = W.C2().policies()
    rt = build_rule_table(docs)
    ft = flatten(rt)
    enc = Encoder(ft.manifest)
    acts = [f"a{i % 8}" if i % 3 else f"zz{i}" for i in range(40)]
    acts = list(dict.fromkeys(acts)) + ["a0", "a7"]
    inp = {"actions": acts, "principal": {"id": "p1", "roles": ["user", "manager", "admin", "x1", "x2"], "attr": {"dept": "d1"}},
           "resource": {"kind": "kind_3", "id": "r", "attr": {"owner": "p1", "dept": "d1", "status": "OPEN", "locked": False}}}
    b = enc.encode([inp])
    assert b.n_pass > 1
    c_out = cref.check(ft.blob, b.columns, 1, b.max_actions)
    k_out = hostsim.check(ft.blob, b.columns, 1, b.max_actions)
    py = CheckOracle(rt).check(inp)
    for k, a in enumerate(acts):
        assert c_out[0, k] == py["actions"][a]["effect"] == k_out[0, k], a


def test_hierarchy_functions_on_attributes():
    """hierarchy(...) relations over request attributes (fused ops, no hierarchy value at run time): oracle #1 vs oracle #2
    vs the kernel core, incl. custom delimiters, empty / trailing segments and non-string operands (error => no match)."""
    import random
    exprs = [
        'hierarchy(P.attr.scope).ancestorOf(hierarchy(R.attr.scope))',
        'hierarchy(R.attr.scope).descendentOf(hierarchy(P.attr.scope))',
        'hierarchy(P.attr.scope).immediateParentOf(hierarchy(R.attr.scope))',
        'hierarchy(R.attr.scope).immediateChildOf(hierarchy(P.attr.scope))',
        'hierarchy(P.attr.scope).siblingOf(hierarchy(R.attr.scope))',
        'hierarchy(P.attr.scope).overlaps(hierarchy(R.attr.scope))',
        'hierarchy(P.attr.scope) == hierarchy(R.attr.scope)',
        'hierarchy(P.attr.scope) != hierarchy(R.attr.scope)',
        'hierarchy(P.attr.path, "::").ancestorOf(hierarchy(R.attr.path, "::"))',
        'hierarchy(P.attr.path, "::").overlaps(hierarchy(R.attr.scope))',
        'hierarchy(P.attr.scope).size() >= 3',
        'size(hierarchy(R.attr.path, "::")) == 2',
        'hierarchy(P.attr.scope).commonAncestors(hierarchy(R.attr.scope)) == hierarchy("a.b")',
        'hierarchy(P.attr.scope).commonAncestors(hierarchy(R.attr.scope)).size() > 0',
        'hierarchy("a.b.c").ancestorOf(hierarchy(R.attr.scope)) || hierarchy(R.attr.scope).siblingOf(hierarchy("a.b.x"))',
    ]
    rules = [{"actions": [f"a{i}"], "effect": "EFFECT_ALLOW", "roles": ["*"], "condition": {"match": {"expr": e}}} for i, e in enumerate(exprs)]
    pol = {"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {"resource": "doc", "version": "default", "rules": rules}}
    rt = build_rule_table([pol])
    ft = flatten(rt)
    r = random.Random(5)
    segs = ["a", "b", "c", "x", "", "ab"]

    def rand_scope(delim):
        return delim.join(r.choice(segs) for _ in range(r.randrange(0, 5)))

    inputs = []
    for _ in range(400):
        p = {"scope": rand_scope("."), "path": rand_scope("::")}
        q = {"scope": rand_scope("."), "path": rand_scope("::")}
        if r.random() < 0.3:
            q["scope"] = p["scope"] + "." + r.choice(segs)       # a child
        if r.random() < 0.1:
            q["scope"] = p["scope"]
        if r.random() < 0.05:
            p["scope"] = 7                                        # not a string: error => condition false
        if r.random() < 0.05:
            del q["scope"]
        inputs.append({"requestId": "h", "actions": [f"a{i}" for i in range(len(exprs))],
                       "principal": {"id": "u", "roles": ["user"], "attr": p}, "resource": {"kind": "doc", "id": "d", "attr": q}})
    orc = CheckOracle(rt)
    b = Encoder(ft.manifest).encode(inputs)
    c_out = cref.check(ft.blob, b.columns, b.n, b.max_actions)
    allow = 0
    for j, inp in enumerate(inputs):
        py = orc.check(inp)
        for k, a in enumerate(inp["actions"]):
            assert c_out[j, k] == py["actions"][a]["effect"], (exprs[k], inp["principal"]["attr"], inp["resource"]["attr"])
            allow += c_out[j, k] == 1
    assert allow > 200
    for mode in (0, 1):
        assert (hostsim.check(ft.blob, b.columns, b.n, b.max_actions, mode=mode) == c_out).all(), mode


def test_workload_c5_adversarial():
    """BASELINE.json configs[4]: 1000 policies, deep CEL (nested trees, ternaries, dynamic map keys, 2-variable
    comprehensions over maps, JWT claims, `in ... .split()`, timestamps vs now(), chained variables), Zipf kinds."""
    w = W.C5()
    rt, ft, enc = W.build(w)
    f = w.fields(1536)
    inputs = w.inputs(f, range(f["n"]))
    b = enc.encode(inputs)
    c_out = cref.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, n_threads=4)
    assert 0.2 < (c_out == 1).mean() < 0.7
    for mode in (0, 1, 2):
        assert (hostsim.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, mode=mode) == c_out).all(), mode
    orc = CheckOracle(rt)
    for j in range(0, f["n"], 3):
        g = orc.check(inputs[j], NOW)
        for k, a in enumerate(w.actions):
            assert g["actions"][a]["effect"] == c_out[j, k], (j, a)


def test_timestamp_and_duration_accessors():
    """getFullYear ... getMilliseconds (UTC) on timestamps from request attributes over 1875-2160, durations incl. negative
    ones: Python's datetime (expected values travel as attributes), oracle #1, oracle #2 and the kernel core must agree."""
    import datetime
    import random
    fields = ["getFullYear", "getMonth", "getDayOfYear", "getDayOfMonth", "getDate", "getDayOfWeek", "getHours", "getMinutes", "getSeconds",
              "getMilliseconds"]
    exprs = [f"timestamp(R.attr.ts).{f}() == int(R.attr.want_{f})" for f in fields] + [
        "(now() - timestamp(R.attr.ts)).getHours() > 24",
        'duration("90m").getMinutes() == 90 && duration("-1500ms").getMilliseconds() == -1500 && duration("-1500ms").getSeconds() == -1',
        "now().getDayOfWeek() == 1 && now().getHours() == 0",
        # fixed-offset zones: -05:30 (east-negative) and +09:00; "UTC"; an invalid zone text is an error => no match
        'timestamp(R.attr.ts).getHours("-05:30") == int(R.attr.h_m530) && timestamp(R.attr.ts).getDate("-05:30") == int(R.attr.d_m530)',
        'timestamp(R.attr.ts).getDayOfWeek("+09:00") == int(R.attr.dow_p9) && timestamp(R.attr.ts).getMinutes("UTC") == int(R.attr.want_getMinutes)',
        '!(timestamp(R.attr.ts).getHours("x:y") >= 0)']
    rules = [{"actions": [f"a{i}"], "effect": "EFFECT_ALLOW", "roles": ["*"], "condition": {"match": {"expr": e}}} for i, e in enumerate(exprs)]
    rt = build_rule_table([{"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {"resource": "doc", "version": "default", "rules": rules}}])
    ft = flatten(rt)
    r = random.Random(3)
    inputs, perturbed = [], []
    for _ in range(300):
        t = datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc) + datetime.timedelta(
            seconds=r.randrange(-3_000_000_000, 6_000_000_000), milliseconds=r.randrange(0, 1000))
        iso = t.strftime("%Y-%m-%dT%H:%M:%S.") + f"{t.microsecond // 1000:03d}Z"
        want = {"getFullYear": t.year, "getMonth": t.month - 1, "getDayOfYear": t.timetuple().tm_yday - 1, "getDayOfMonth": t.day - 1,
                "getDate": t.day, "getDayOfWeek": (t.weekday() + 1) % 7, "getHours": t.hour, "getMinutes": t.minute, "getSeconds": t.second,
                "getMilliseconds": t.microsecond // 1000}
        wrong = r.choice(fields) if r.random() < 0.2 else None
        if wrong:
            want[wrong] += 1
        perturbed.append(wrong)
        t1, t2 = t - datetime.timedelta(minutes=330), t + datetime.timedelta(hours=9)
        attr = {"ts": iso, **{f"want_{k}": v for k, v in want.items()}, "h_m530": t1.hour, "d_m530": t1.day, "dow_p9": (t2.weekday() + 1) % 7}
        inputs.append({"requestId": "t", "actions": [f"a{i}" for i in range(len(exprs))], "principal": {"id": "u", "roles": ["r"]},
                       "resource": {"kind": "doc", "id": "d", "attr": attr}})
    orc = CheckOracle(rt)
    b = Encoder(ft.manifest).encode(inputs)
    c_out = cref.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns)
    for j, inp in enumerate(inputs):
        py = orc.check(inp, NOW)
        for k, a in enumerate(inp["actions"]):
            assert c_out[j, k] == py["actions"][a]["effect"], (exprs[k], inp["resource"]["attr"]["ts"])
        for k, f in enumerate(fields):      # datetime's answer: ALLOW unless this field's expectation was perturbed
            assert c_out[j, k] == (2 if perturbed[j] == f else 1), (f, inp["resource"]["attr"]["ts"])
    assert (c_out[:, len(fields) + 1] == 1).all() and (c_out[:, len(fields) + 2] == 1).all()     # 2024-01-01T00:00:00Z is a Monday
    assert (c_out[:, len(fields) + 3] == 1).all() and (c_out[:, len(fields) + 5] == 2).all()     # zone offsets; invalid zone => error
    assert ((c_out[:, len(fields) + 4] == 1) == np.array([p != "getMinutes" for p in perturbed])).all()
    for mode in (0, 1):
        assert (hostsim.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, mode=mode) == c_out).all(), mode


# timestamp(<text>): Go's time.Parse(time.RFC3339, text), then cel-go's range check on the instant.  "ok" = a timestamp,
# "err" = a failed evaluation, "flag" = a valid CEL timestamp the device's int64 nanoseconds cannot hold (the call fails loudly)
TIMESTAMP_TEXTS = [
    ("2021-04-22T10:05:20Z", "ok"), ("2021-04-22T10:05:20.021-05:00", "ok"), ("2021-04-22T10:05:20,5Z", "ok"), ("2021-04-22T10:05:20.1234567891Z", "ok"),
    ("2021-04-22T10:05:20+24:00", "ok"), ("2021-04-22T10:05:20+23:60", "ok"), ("2021-04-22T10:05:20-24:60", "ok"), ("2024-02-29T23:59:59Z", "ok"),
    ("2021-04-22t10:05:20Z", "err"), ("2021-04-22T10:05:20z", "err"), ("2021-04-22T10:05:20+25:00", "err"), ("2021-04-22T10:05:20+00:61", "err"),
    ("2021-04-22T10:05:20Z\n", "err"), (" 2021-04-22T10:05:20Z", "err"), ("2021-04-22T24:00:00Z", "err"), ("2021-04-22T23:59:60Z", "err"),
    ("2021-02-29T00:00:00Z", "err"), ("2021-04-22 10:05:20Z", "err"), ("2021-04-22T10:05:20.Z", "err"), ("2021-04-22T10:05:20", "err"),
    ("2021-04-22T10:05:20+0530", "err"), ("2021-4-22T10:05:20Z", "err"), ("2021-04-22T10:05:2\u0660Z", "err"), ("2021-13-01T00:00:00Z", "err"),
    # the range is on the INSTANT: year 0000 is fine to parse, and fine as a timestamp once the offset moves it into year 1
    ("0000-06-01T00:00:00Z", "err"), ("0001-01-01T00:00:00+00:01", "err"), ("9999-12-31T23:59:59-00:01", "err"),
    ("0000-12-31T05:35:43-23:59", "flag"), ("0001-01-01T00:00:00Z", "flag"), ("9999-12-31T23:59:59Z", "flag"), ("1677-09-21T00:12:43Z", "flag"),
    ("1677-09-21T00:12:44Z", "ok"), ("2262-04-11T23:47:16Z", "ok"), ("2262-04-11T23:47:17Z", "flag"),
]


def test_timestamp_texts_like_go_time_parse():
    """The expected outcome of each text is written out from Go's parser and cel-go's range check; oracle #1 must give it,
    oracle #2 and the kernel core must follow -- for texts in request attributes and for the same texts as literals."""
    rt, ft = run_time_value_table("timestamp(R.attr.ts) <= timestamp(R.attr.ts)")
    enc = Encoder(ft.manifest)
    for text, kind in TIMESTAMP_TEXTS:
        inp = {"principal": {"id": "p", "roles": ["r"]}, "resource": {"kind": "leave_request", "id": "r", "attr": {"ts": text}}, "actions": ["a"]}
        want = CheckOracle(rt).check(inp)["actions"]["a"]["effect"]
        assert want == (2 if kind == "err" else 1), (text, want)
        b = enc.encode([inp])
        for fn in (hostsim.check, cref.check):
            if kind == "flag":
                with pytest.raises(RuntimeError, match="-2"):
                    fn(ft.blob, b.columns, 1, 1)
            else:
                assert fn(ft.blob, b.columns, 1, 1)[0, 0] == want, (text, fn.__module__)
        if not text.isascii() or not text.isprintable():
            continue
        # the same text as a literal: parsed by the table builder (table/consts.py)
        lit = f'timestamp("{text}") <= timestamp("{text}")'
        if kind == "flag":
            with pytest.raises(Unsupported):
                run_time_value_table(lit)
            continue
        rt2, ft2 = run_time_value_table(lit)
        b2 = Encoder(ft2.manifest).encode([inp])
        assert CheckOracle(rt2).check(inp)["actions"]["a"]["effect"] == want, text
        assert hostsim.check(ft2.blob, b2.columns, 1, 1)[0, 0] == want and cref.check(ft2.blob, b2.columns, 1, 1)[0, 0] == want, text


def test_duration_of_request_strings():
    """duration(<attribute string>) parsed on the device (Go time.ParseDuration): random and boundary texts (sign, fractions up
    to 22 digits, every unit incl. both micro signs, int64 limits, malformed input => CEL error => no match)."""
    import random
    from oracle.celeval import CelError, parse_duration
    exprs = ['duration(R.attr.d) == duration(R.attr.e)', 'duration(R.attr.d) > duration("1h")', 'duration(R.attr.d).getSeconds() == int(R.attr.secs)',
             'timestamp(R.attr.ts) + duration(R.attr.cooldown) > now()', 'duration(R.attr.d) < duration("0s")', 'duration(R.attr.cooldown).getMinutes() == 62']
    rules = [{"actions": [f"a{i}"], "effect": "EFFECT_ALLOW", "roles": ["*"], "condition": {"match": {"expr": e}}} for i, e in enumerate(exprs)]
    rt = build_rule_table([{"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {"resource": "doc", "version": "default", "rules": rules}}])
    ft = flatten(rt)
    r = random.Random(11)
    units = ["ns", "us", "\u00b5s", "\u03bcs", "ms", "s", "m", "h"]
    edge = ["", "0", "-0", "+", "1", "h", "1.5", "..s", "1..5s", "9223372036854775807ns", "9223372036854775808ns", "-9223372036854775808ns",
            "-9223372036854775809ns", "2562047h47m16.854775807s", "2562047h47m16.854775808s", "1e3s", " 1s", "1s ", "1H", ".5s", "5.s",
            "0.0000000000001h", "0.9999999999999999999999h"]

    def rd():
        if r.random() < 0.08:
            return r.choice(edge)
        s = r.choice(["", "", "-", "+"])
        for _ in range(r.randrange(1, 4)):
            w = str(r.randrange(0, 10 ** r.randrange(1, 7))) if r.random() < 0.9 else ""
            f = ("." + "".join(r.choice("0123456789") for _ in range(r.randrange(0, 14)))) if r.random() < 0.4 else ""
            if not w and len(f) < 2:
                w = "3"
            s += w + f + r.choice(units)
        return s

    inputs = []
    for _ in range(800):
        d = rd()
        e = d if r.random() < 0.3 else rd()
        try:
            ns = parse_duration(d).ns
            secs = ns // 10 ** 9 if ns >= 0 else -((-ns) // 10 ** 9)
        except CelError:
            secs = 0
        inputs.append({"requestId": "x", "actions": [f"a{i}" for i in range(len(exprs))], "principal": {"id": "u", "roles": ["r"]},
                       "resource": {"kind": "doc", "id": "d", "attr": {"d": d, "e": e, "secs": secs, "ts": "2023-12-31T23:00:00Z", "cooldown": "1h2m30s"}}})
    orc = CheckOracle(rt)
    b = Encoder(ft.manifest).encode(inputs)
    c_out = cref.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns)
    for j, inp in enumerate(inputs):
        py = orc.check(inp, NOW)
        for k, a in enumerate(inp["actions"]):
            assert c_out[j, k] == py["actions"][a]["effect"], (exprs[k], inp["resource"]["attr"]["d"], inp["resource"]["attr"]["e"])
    assert (c_out[:, 3] == 1).all() and (c_out[:, 5] == 1).all() and 100 < (c_out[:, 0] == 1).sum() < 500
    for mode in (0, 1):
        assert (hostsim.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, mode=mode) == c_out).all(), mode


def test_check_resources_api_goldens():
    """The reference's API-level CheckResources goldens (cr_case_00 ... 08): oracle #1, oracle #2 and the kernel core."""
    from helpers import check_resources_api_cases
    rt = store_rule_table()
    ft = flatten(rt, globals_={"environment": "test"})
    orc = CheckOracle(rt, globals_={"environment": "test"})
    enc = Encoder(ft.manifest)
    names = {"EFFECT_ALLOW": 1, "EFFECT_DENY": 2}
    n = 0
    for f, ci, want in check_resources_api_cases():
        py = orc.check(ci)
        b = enc.encode([ci])
        c_out = cref.check(ft.blob, b.columns, 1, b.max_actions)
        k_out = hostsim.check(ft.blob, b.columns, 1, b.max_actions)
        for k, a in enumerate(ci["actions"]):
            assert py["actions"][a]["effect"] == names[want[a]], (f, a, "oracle #1")
            assert c_out[0, k] == names[want[a]], (f, a, "oracle #2")
            assert k_out[0, k] == names[want[a]], (f, a, "kernel core")
            n += 1
    assert n == 48


def test_verify_suite_goldens():
    """145 engine answers from the reference's policy-test goldens (verify/cases): oracle #1, oracle #2 and the kernel core,
    under every engine configuration the suites use (globals, default policy version / scope, lenient scope search, now)."""
    import json
    from helpers import verify_suite_cases
    names = {"EFFECT_ALLOW": 1, "EFFECT_DENY": 2}
    rt = store_rule_table()
    n = 0
    for (gl, dver, dscope, lenient), cases in verify_suite_cases():
        ft = flatten(rt, globals_=json.loads(gl))
        orc = CheckOracle(rt, globals_=json.loads(gl), default_version=dver, default_scope=dscope, lenient_scope_search=lenient)
        enc = Encoder(ft.manifest, default_version=dver, default_scope=dscope, lenient_scope_search=lenient)
        fl = L.BATCH_FLAG_LENIENT if lenient else 0
        for c in cases:
            now = parse_timestamp(c["now"]) if c["now"] else NOW
            py = orc.check(c["input"], now)
            b = enc.encode([c["input"]])
            c_out = cref.check(ft.blob, b.columns, 1, b.max_actions, now.ns, fl)
            k_out = hostsim.check(ft.blob, b.columns, 1, b.max_actions, now.ns, fl)
            for k, a in enumerate(c["input"]["actions"]):
                w = names[c["want"][a]]
                assert py["actions"][a]["effect"] == w, (c["file"], c["test"], a, "oracle #1")
                assert c_out[0, k] == w, (c["file"], c["test"], a, "oracle #2")
                assert k_out[0, k] == w, (c["file"], c["test"], a, "kernel core")
                n += 1
    assert n == 145


def test_decision_metadata_on_engine_goldens(store_flat):
    """ActionEffect.Policy / Scope and EffectiveDerivedRoles (ruletable.go:753-782, 913-922, 936-979, 1082-1148): the
    metadata body of the kernel core (cb::eval_request_meta, the optional plane behind cgpu_check_meta) against oracle #1
    on all 166 reference decisions -- and its effects against the bit-parallel body."""
    from cerbos_b200 import meta as M
    ft = store_flat
    orc = CheckOracle(store_rule_table(), globals_=G)
    n = n_edr = 0
    for cid, lenient, inp, want in engine_decisions():
        enc = Encoder(ft.manifest, lenient_scope_search=lenient)
        b = enc.encode([inp])
        fl = L.BATCH_FLAG_LENIENT if lenient else 0
        eff, am, rm = hostsim.check_meta(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, fl)
        k_out = hostsim.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, fl)
        py = orc.check(inp, NOW, lenient=lenient)
        p, r = inp.get("principal") or {}, inp.get("resource") or {}
        for k, a in enumerate(inp["actions"]):
            pol, sc = M.decode_action(int(am[0, k]), rm[0], ft.manifest, p.get("id", ""), r.get("kind", ""),
                                      p.get("policyVersion") or "default", r.get("policyVersion") or "default")
            w = py["actions"][a]
            assert (int(eff[0, k]), pol, sc) == (w["effect"], w["policy"], w["scope"]), (cid, a)
            assert eff[0, k] == k_out[0, k], (cid, a)
            # ... and the reference's own recorded answer
            names = {"EFFECT_ALLOW": 1, "EFFECT_DENY": 2}
            wa = want["actions"][a]
            assert (names[wa["effect"]], wa.get("policy", ""), wa.get("scope", "")) == (int(eff[0, k]), pol, sc), (cid, a)
            n += 1
        edr = M.decode_edr(int(rm[0]["effective_derived_roles"]), ft.manifest)
        assert edr == py["effectiveDerivedRoles"] == sorted(want.get("effectiveDerivedRoles", want.get("effective_derived_roles")) or []), cid
        n_edr += bool(edr)
    assert n == 166 and n_edr >= 20
