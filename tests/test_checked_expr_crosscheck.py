"""Independent pin of the CEL front end: the reference's own compiled rule-table bundle (tests/golden/
ruletable_bundle_unencrypted.crrt = internal/test/testdata/bundle/v2_ruletable/bundle_unencrypted.crrt) carries every
condition twice -- `Expr.original` (the source text our parser reads) and `Expr.checked` (the CheckedExpr cel-go's parser and
type checker produced, google/api/expr/v1alpha1/checked.proto).  This test decodes the checked tree with a hand-written
protobuf reader and compares it, node for node, with what cerbos_b200.cel.parser makes of the text: same operators, same
function names, same receivers, same literals, same select chains.  Two normalisations, both semantics-free: cel-go balances
chains of && / || (compared as flattened operand lists), and it expands macros into fold comprehensions (all / exists /
exists_one / map / filter are folded back by their documented shapes before comparing)."""
import os
import struct

from cerbos_b200.cel.ast import Call, Const, Ident, ListLit, Macro, MapLit, Select, UInt
from cerbos_b200.cel.parser import parse
from cerbos_b200.table.ruletable_pb import fields

BUNDLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ruletable_bundle_unencrypted.crrt")


def _s(v):
    return bytes(v).decode("utf-8")


def _sint(v):          # int64 varint
    return v - (1 << 64) if v >> 63 else v


# ---- google.api.expr.v1alpha1.Expr -> a plain tuple tree
def dec_const(buf):
    for fno, wt, v in fields(buf):
        if fno == 1:
            return ("const", None)
        if fno == 2:
            return ("const", bool(v))
        if fno == 3:
            return ("const", _sint(v))
        if fno == 4:
            return ("const", ("u", v))
        if fno == 5:
            return ("const", struct.unpack("<d", v)[0])
        if fno == 6:
            return ("const", _s(v))
        if fno == 7:
            return ("const", bytes(v))
    return ("const", 0)     # proto3: an absent int64 field is 0 -- but a Constant always sets its oneof


def dec_expr(buf):
    for fno, wt, v in fields(buf):
        if fno == 3:
            c = dec_const(v)
            return c if any(True for _ in fields(v)) else ("const", None)
        if fno == 4:
            return ("ident", next((_s(x) for f, _, x in fields(v) if f == 1), ""))
        if fno == 5:
            op, fld, test = None, "", False
            for f, _, x in fields(v):
                if f == 1:
                    op = dec_expr(x)
                elif f == 2:
                    fld = _s(x)
                elif f == 3:
                    test = bool(x)
            return ("select", op, fld, test)
        if fno == 6:
            tgt, fn, args = None, "", []
            for f, _, x in fields(v):
                if f == 1:
                    tgt = dec_expr(x)
                elif f == 2:
                    fn = _s(x)
                elif f == 3:
                    args.append(dec_expr(x))
            return ("call", fn, tgt, args)
        if fno == 7:
            return ("list", [dec_expr(x) for f, _, x in fields(v) if f == 1])
        if fno == 8:
            ents = []
            for f, _, x in fields(v):
                if f == 2:
                    k = val = None
                    for f2, _, y in fields(x):
                        if f2 == 3:
                            k = dec_expr(y)
                        elif f2 == 4:
                            val = dec_expr(y)
                    ents.append((k, val))
            return ("map", ents)
        if fno == 9:
            d = {}
            for f, _, x in fields(v):
                d[f] = _s(x) if f in (1, 3, 8) else dec_expr(x)
            return ("fold", d.get(1, ""), d.get(2), d.get(3, ""), d.get(4), d.get(5), d.get(6), d.get(7), d.get(8, ""))
    return ("const", None)


def checked_root(buf):
    """CheckedExpr {reference_map = 2, type_map = 3, expr = 4, source_info = 5, expr_version = 6} -> its expr, or None if `buf`
    is not one (cel-go always fills source_info)"""
    try:
        fs = list(fields(buf))
    except Exception:
        return None
    nums = {f for f, _, _ in fs}
    if not ({4, 5} <= nums <= {2, 3, 4, 5, 6}):
        return None
    return dec_expr(next(v for f, _, v in fs if f == 4))


# ---- normal form shared by both sides
def flat_logic(fn, args):
    out = []
    for a in args:
        if a[0] == "call" and a[1] == fn and a[2] is None:
            out += flat_logic(fn, a[3])
        else:
            out.append(a)
    return out


def fold_macro(t):
    """fold comprehension -> ("macro", name, vars, range, args) by the shapes cel-go's macro expander emits (parser/macro.go)"""
    _, iv, rng, accu, init, cond, step, result, iv2 = t
    acc = ("ident", accu)
    if init == ("const", True) and step[0] == "call" and step[1] == "_&&_" and step[3][0] == acc:
        return ("macro", "all", [iv] + ([iv2] if iv2 else []), rng, [step[3][1]])
    if init == ("const", False) and step[0] == "call" and step[1] == "_||_" and step[3][0] == acc:
        return ("macro", "exists", [iv] + ([iv2] if iv2 else []), rng, [step[3][1]])
    if init == ("const", 0) and step[0] == "call" and step[1] == "_?_:_":
        return ("macro", "exists_one", [iv] + ([iv2] if iv2 else []), rng, [step[3][0]])
    if init == ("list", []) and step[0] == "call":
        if step[1] == "_+_" and step[3][0] == acc and step[3][1][0] == "list":                       # map(x, t)
            return ("macro", "map", [iv], rng, [step[3][1][1][0]])
        if step[1] == "_?_:_" and step[3][1][0] == "call" and step[3][1][1] == "_+_":                # filter(x, p) / map(x, p, t)
            elem = step[3][1][3][1][1][0]
            if elem == ("ident", iv):
                return ("macro", "filter", [iv], rng, [step[3][0]])
            return ("macro", "map", [iv], rng, [step[3][0], elem])
    return t


def norm(t):
    k = t[0]
    if k == "call":
        fn, tgt, args = t[1], t[2], [norm(a) for a in t[3]]
        if fn in ("_&&_", "_||_") and tgt is None:
            return ("logic", fn, flat_logic_n(fn, args))
        return ("call", fn, norm(tgt) if tgt is not None else None, args)
    if k == "select":
        return ("select", norm(t[1]), t[2], t[3])
    if k == "list":
        return ("list", [norm(x) for x in t[1]])
    if k == "map":
        return ("map", [(norm(a), norm(b)) for a, b in t[1]])
    if k == "macro":
        return ("macro", t[1], t[2], norm(t[3]), [norm(a) for a in t[4]])
    if k == "fold":
        m = fold_macro(t)
        if m[0] == "macro":
            return ("macro", m[1], m[2], norm(m[3]), [norm(a) for a in m[4]])
        return t
    return t


def flat_logic_n(fn, args):
    out = []
    for a in args:
        if a[0] == "logic" and a[1] == fn:
            out += a[2]
        else:
            out.append(a)
    return out


def ours(n):
    """cerbos_b200.cel.ast node -> the same tuple tree"""
    if isinstance(n, Const):
        v = n.value
        if isinstance(v, UInt):
            return ("const", ("u", int(v)))
        return ("const", v)
    if isinstance(n, Ident):
        return ("ident", n.name)
    if isinstance(n, Select):
        return ("select", ours(n.operand), n.field, bool(n.test_only))
    if isinstance(n, Call):
        return ("call", n.fn, ours(n.target) if n.target is not None else None, [ours(a) for a in n.args])
    if isinstance(n, ListLit):
        return ("list", [ours(e) for e in n.elems])
    if isinstance(n, MapLit):
        return ("map", [(ours(k), ours(v)) for k, v in n.entries])
    if isinstance(n, Macro):
        name = {"all2": "all", "exists2": "exists", "exists_one2": "exists_one"}.get(n.name, n.name)   # two-variable forms: the variable list says so
        return ("macro", name, list(n.vars), ours(n.target), [ours(a) for a in n.args])
    raise AssertionError(type(n))


def walk_exprs(buf, out, depth=0):
    """every runtimev1.Expr {original = 1, checked = 2} nested anywhere in the message"""
    if depth > 24:
        return
    try:
        fs = list(fields(buf))
    except Exception:
        return
    nums = {f for f, _, _ in fs}
    if nums and nums <= {1, 2} and all(wt == 2 for _, wt, _ in fs):
        try:
            orig = next((_s(v) for f, _, v in fs if f == 1), None)
            chk = next((v for f, _, v in fs if f == 2), None)
            if orig is not None and chk is not None and orig.strip() and checked_root(chk) is not None:
                out.append((orig, bytes(chk)))       # (whether OUR parser takes the text is the test's business)
                return
        except Exception:
            pass
    for _, wt, v in fs:
        if wt == 2 and len(v) > 4:
            walk_exprs(v, out, depth + 1)


def test_parser_matches_the_reference_bundles_checked_expressions():
    found = []
    walk_exprs(open(BUNDLE, "rb").read(), found)
    exprs = {}
    for orig, chk in found:
        exprs.setdefault(orig, chk)
    assert len(exprs) >= 90, len(exprs)
    same = macros = 0
    for orig, chk in sorted(exprs.items()):
        want = norm(checked_root(chk))
        got = norm(ours(parse(orig)))
        assert got == want, (orig, got, want)
        same += 1
        macros += "macro" in repr(want)
    assert same == len(exprs) and macros >= 1, (same, macros)
