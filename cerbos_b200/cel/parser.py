"""CEL text -> AST (recursive descent).

Stands in for the cel-go v0.27.0 parser + macro expander that the reference
invokes through ``conditions.StdEnv.Compile`` (internal/conditions/cel.go:62-75,
internal/compile/conditions.go:63-89).  Grammar follows the CEL language
definition (cel-spec langdef.md "Syntax"); macros enabled are the standard ones
plus the extensions the reference turns on: TwoVarComprehensions, Bindings,
Lists (sortBy), Math (greatest/least are ordinary variadic calls here).

Namespaced helper functions (``math.abs``, ``base64.encode``, ``lists.range``,
``strings.quote``, ``cel.bind``) are resolved at parse time into global calls
named ``ns.fn`` because the reference environment declares no variables called
``math``/``base64``/``lists``/``strings``/``cel``.
"""
from __future__ import annotations

from .ast import Call, Const, Ident, ListLit, Macro, MapLit, Node, Select, UInt


class CelSyntaxError(ValueError):
    pass


_RESERVED = {
    "as", "break", "const", "continue", "else", "for", "function", "if", "import",
    "let", "loop", "package", "namespace", "return", "var", "void", "while",
}
_NAMESPACES = {"math", "base64", "lists", "strings", "cel"}

_PUNCT3 = ()
_PUNCT2 = ("&&", "||", "<=", ">=", "==", "!=")
_PUNCT1 = "()[]{}.,?:+-*/%!<>"

INT64_MIN = -(1 << 63)
INT64_MAX = (1 << 63) - 1
UINT64_MAX = (1 << 64) - 1


class _Tok:
    __slots__ = ("kind", "val", "pos")

    def __init__(self, kind, val, pos):
        self.kind = kind  # 'id','int','uint','float','str','bytes','p','eof'
        self.val = val
        self.pos = pos

    def __repr__(self):
        return f"<{self.kind} {self.val!r}@{self.pos}>"


def _is_id_start(c):
    return c == "_" or ("a" <= c <= "z") or ("A" <= c <= "Z")


def _is_id_char(c):
    return _is_id_start(c) or ("0" <= c <= "9")


def _unescape(body: str, is_bytes: bool, pos: int):
    """Process CEL escape sequences. Returns str (or bytes when is_bytes)."""
    out_s: list[str] = []
    out_b = bytearray()
    i = 0
    n = len(body)

    def emit_cp(cp):
        if is_bytes:
            out_b.extend(chr(cp).encode("utf-8"))
        else:
            out_s.append(chr(cp))

    while i < n:
        c = body[i]
        if c != "\\":
            if is_bytes:
                out_b.extend(c.encode("utf-8"))
            else:
                out_s.append(c)
            i += 1
            continue
        i += 1
        if i >= n:
            raise CelSyntaxError(f"dangling escape at {pos}")
        e = body[i]
        i += 1
        simple = {"a": 7, "b": 8, "f": 12, "n": 10, "r": 13, "t": 9, "v": 11,
                  "\\": 92, "?": 63, '"': 34, "'": 39, "`": 96}
        if e in simple:
            emit_cp(simple[e])
        elif e in "xX":
            h = body[i:i + 2]
            if len(h) != 2:
                raise CelSyntaxError(f"bad \\x escape at {pos}")
            v = int(h, 16)
            i += 2
            if is_bytes:
                out_b.append(v)
            else:
                out_s.append(chr(v))
        elif e == "u":
            h = body[i:i + 4]
            if len(h) != 4 or is_bytes:
                raise CelSyntaxError(f"bad \\u escape at {pos}")
            out_s.append(chr(int(h, 16)))
            i += 4
        elif e == "U":
            h = body[i:i + 8]
            if len(h) != 8 or is_bytes:
                raise CelSyntaxError(f"bad \\U escape at {pos}")
            out_s.append(chr(int(h, 16)))
            i += 8
        elif e in "0123":
            h = body[i - 1:i + 2]
            if len(h) != 3:
                raise CelSyntaxError(f"bad octal escape at {pos}")
            v = int(h, 8)
            i += 2
            if is_bytes:
                out_b.append(v)
            else:
                out_s.append(chr(v))
        else:
            raise CelSyntaxError(f"unknown escape \\{e} at {pos}")
    return bytes(out_b) if is_bytes else "".join(out_s)


def _lex(src: str):
    toks = []
    i = 0
    n = len(src)
    while i < n:
        c = src[i]
        if c in " \t\r\n\f":
            i += 1
            continue
        if c == "/" and src[i:i + 2] == "//":
            while i < n and src[i] != "\n":
                i += 1
            continue
        start = i
        # string / bytes literals (with optional r/b prefixes)
        j = i
        raw = False
        is_bytes = False
        while j < n and src[j] in "rRbB" and j - i < 2:
            if src[j] in "rR":
                if raw:
                    break
                raw = True
            else:
                if is_bytes:
                    break
                is_bytes = True
            j += 1
        if j < n and src[j] in "\"'" and (j == i or all(ch in "rRbB" for ch in src[i:j])):
            q = src[j]
            if src[j:j + 3] == q * 3:
                end = src.find(q * 3, j + 3)
                # escaped quotes inside triple-quoted non-raw strings
                if not raw:
                    k = j + 3
                    while True:
                        end = src.find(q * 3, k)
                        if end < 0:
                            break
                        # count preceding backslashes
                        b = 0
                        m = end - 1
                        while m >= j + 3 and src[m] == "\\":
                            b += 1
                            m -= 1
                        if b % 2 == 0:
                            break
                        k = end + 1
                if end < 0:
                    raise CelSyntaxError(f"unterminated string at {start}")
                body = src[j + 3:end]
                i = end + 3
            else:
                k = j + 1
                while k < n and src[k] != q:
                    if src[k] == "\n":
                        raise CelSyntaxError(f"newline in string at {start}")
                    if src[k] == "\\" and not raw:
                        k += 1
                    k += 1
                if k >= n:
                    raise CelSyntaxError(f"unterminated string at {start}")
                body = src[j + 1:k]
                i = k + 1
            if raw:
                val = body.encode("utf-8") if is_bytes else body
            else:
                val = _unescape(body, is_bytes, start)
            toks.append(_Tok("bytes" if is_bytes else "str", val, start))
            continue
        if _is_id_start(c):
            j = i + 1
            while j < n and _is_id_char(src[j]):
                j += 1
            toks.append(_Tok("id", src[i:j], start))
            i = j
            continue
        if c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            j = i
            if src[j:j + 2] in ("0x", "0X"):
                j += 2
                while j < n and src[j] in "0123456789abcdefABCDEF":
                    j += 1
                text = src[i:j]
                if j < n and src[j] in "uU":
                    toks.append(_Tok("uint", int(text, 16), start))
                    j += 1
                else:
                    toks.append(_Tok("int", int(text, 16), start))
                i = j
                continue
            while j < n and src[j].isdigit():
                j += 1
            is_float = False
            if j < n and src[j] == "." and j + 1 < n and src[j + 1].isdigit():
                is_float = True
                j += 1
                while j < n and src[j].isdigit():
                    j += 1
            if j < n and src[j] in "eE":
                k = j + 1
                if k < n and src[k] in "+-":
                    k += 1
                if k < n and src[k].isdigit():
                    is_float = True
                    j = k
                    while j < n and src[j].isdigit():
                        j += 1
            text = src[i:j]
            if is_float:
                toks.append(_Tok("float", float(text), start))
            elif j < n and src[j] in "uU":
                toks.append(_Tok("uint", int(text), start))
                j += 1
            else:
                toks.append(_Tok("int", int(text), start))
            i = j
            continue
        two = src[i:i + 2]
        if two in _PUNCT2:
            toks.append(_Tok("p", two, start))
            i += 2
            continue
        if c in _PUNCT1:
            toks.append(_Tok("p", c, start))
            i += 1
            continue
        raise CelSyntaxError(f"unexpected character {c!r} at {i}")
    toks.append(_Tok("eof", None, n))
    return toks


_REL_OPS = {"<": "_<_", "<=": "_<=_", ">": "_>_", ">=": "_>=_", "==": "_==_", "!=": "_!=_"}
_ADD_OPS = {"+": "_+_", "-": "_-_"}
_MUL_OPS = {"*": "_*_", "/": "_/_", "%": "_%_"}

_ONE_VAR_MACROS = {"all": 2, "exists": 2, "exists_one": 2, "filter": 2}
_TWO_VAR_MACROS = {"all": 3, "exists": 3, "existsOne": 3, "exists_one": 3}


class _Parser:
    def __init__(self, src: str):
        self.src = src
        self.toks = _lex(src)
        self.i = 0
        self.depth = 0

    # -- token helpers
    @property
    def tok(self):
        return self.toks[self.i]

    def peek(self, k=1):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def is_p(self, v):
        t = self.tok
        return t.kind == "p" and t.val == v

    def accept(self, v):
        if self.is_p(v):
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise CelSyntaxError(f"expected {v!r} at {self.tok.pos} in {self.src!r}, got {self.tok.val!r}")

    # -- grammar
    def parse(self) -> Node:
        e = self.expr()
        if self.tok.kind != "eof":
            raise CelSyntaxError(f"unexpected token {self.tok.val!r} at {self.tok.pos} in {self.src!r}")
        return e

    def expr(self) -> Node:
        self.depth += 1
        if self.depth > 200:
            raise CelSyntaxError("expression nesting too deep")
        c = self.cond_or()
        if self.accept("?"):
            a = self.cond_or()
            self.expect(":")
            b = self.expr()
            c = Call("_?_:_", None, [c, a, b])
        self.depth -= 1
        return c

    def cond_or(self) -> Node:
        e = self.cond_and()
        while self.accept("||"):
            r = self.cond_and()
            e = Call("_||_", None, [e, r])
        return e

    def cond_and(self) -> Node:
        e = self.relation()
        while self.accept("&&"):
            r = self.relation()
            e = Call("_&&_", None, [e, r])
        return e

    def relation(self) -> Node:
        e = self.addition()
        while True:
            t = self.tok
            if t.kind == "p" and t.val in _REL_OPS:
                self.i += 1
                r = self.addition()
                e = Call(_REL_OPS[t.val], None, [e, r])
            elif t.kind == "id" and t.val == "in":
                self.i += 1
                r = self.addition()
                e = Call("@in", None, [e, r])
            else:
                return e

    def addition(self) -> Node:
        e = self.multiplication()
        while True:
            t = self.tok
            if t.kind == "p" and t.val in _ADD_OPS:
                self.i += 1
                r = self.multiplication()
                e = Call(_ADD_OPS[t.val], None, [e, r])
            else:
                return e

    def multiplication(self) -> Node:
        e = self.unary()
        while True:
            t = self.tok
            if t.kind == "p" and t.val in _MUL_OPS:
                self.i += 1
                r = self.unary()
                e = Call(_MUL_OPS[t.val], None, [e, r])
            else:
                return e

    def unary(self) -> Node:
        if self.is_p("!"):
            n = 0
            while self.accept("!"):
                n += 1
            e = self.member()
            if n % 2 == 1:
                e = Call("!_", None, [e])
            return e
        if self.is_p("-"):
            # a single '-' directly before a numeric literal is part of the literal
            if self.peek().kind in ("int", "float"):
                self.i += 1
                t = self.tok
                self.i += 1
                if t.kind == "int":
                    v = -t.val
                    if v < INT64_MIN:
                        raise CelSyntaxError("invalid int literal")
                    e = Const(v)
                else:
                    e = Const(-t.val)
                return self.member_tail(e)
            n = 0
            while self.accept("-"):
                n += 1
            e = self.member()
            if n % 2 == 1:
                e = Call("-_", None, [e])
            return e
        return self.member()

    def member(self) -> Node:
        return self.member_tail(self.primary())

    def member_tail(self, e: Node) -> Node:
        while True:
            if self.accept("."):
                t = self.tok
                if t.kind != "id":
                    raise CelSyntaxError(f"expected identifier after '.' at {t.pos} in {self.src!r}")
                self.i += 1
                name = t.val
                if self.accept("("):
                    args = self.expr_list(")")
                    self.expect(")")
                    e = self.make_member_call(e, name, args)
                else:
                    e = Select(e, name)
            elif self.accept("["):
                idx = self.expr()
                self.expect("]")
                e = Call("_[_]", None, [e, idx])
            else:
                return e

    def expr_list(self, closer) -> list:
        out = []
        if self.is_p(closer):
            return out
        while True:
            out.append(self.expr())
            if not self.accept(","):
                return out
            if self.is_p(closer):  # trailing comma (lists / maps)
                return out

    def primary(self) -> Node:
        t = self.tok
        if t.kind == "p":
            if t.val == "(":
                self.i += 1
                e = self.expr()
                self.expect(")")
                return e
            if t.val == "[":
                self.i += 1
                elems = self.expr_list("]")
                self.expect("]")
                return ListLit(elems)
            if t.val == "{":
                self.i += 1
                entries = []
                if not self.is_p("}"):
                    while True:
                        k = self.expr()
                        self.expect(":")
                        v = self.expr()
                        entries.append((k, v))
                        if not self.accept(","):
                            break
                        if self.is_p("}"):
                            break
                self.expect("}")
                return MapLit(entries)
            if t.val == ".":
                # leading dot: root-namespace identifier
                self.i += 1
                t = self.tok
                if t.kind != "id":
                    raise CelSyntaxError(f"expected identifier at {t.pos}")
            else:
                raise CelSyntaxError(f"unexpected {t.val!r} at {t.pos} in {self.src!r}")
        if t.kind == "int":
            self.i += 1
            if t.val > INT64_MAX:
                raise CelSyntaxError("invalid int literal")
            return Const(t.val)
        if t.kind == "uint":
            self.i += 1
            if t.val > UINT64_MAX:
                raise CelSyntaxError("invalid uint literal")
            return Const(UInt(t.val))
        if t.kind in ("float", "str", "bytes"):
            self.i += 1
            return Const(t.val)
        if t.kind == "id":
            self.i += 1
            name = t.val
            if name == "true":
                return Const(True)
            if name == "false":
                return Const(False)
            if name == "null":
                return Const(None)
            if name == "in" or name in _RESERVED:
                raise CelSyntaxError(f"reserved identifier {name!r} at {t.pos}")
            if self.accept("("):
                args = self.expr_list(")")
                self.expect(")")
                return self.make_global_call(name, args)
            return Ident(name)
        raise CelSyntaxError(f"unexpected end of expression in {self.src!r}")

    # -- macros
    @staticmethod
    def _ident_name(n: Node, what: str) -> str:
        if not isinstance(n, Ident):
            raise CelSyntaxError(f"{what}: argument must be a simple name")
        return n.name

    def make_global_call(self, name: str, args: list) -> Node:
        if name == "has" and len(args) == 1:
            a = args[0]
            if not isinstance(a, Select) or a.test_only:
                raise CelSyntaxError("invalid argument to has() macro")
            return Select(a.operand, a.field, test_only=True)
        return Call(name, None, args)

    def make_member_call(self, target: Node, name: str, args: list) -> Node:
        n = len(args)
        # namespaced functions
        if isinstance(target, Ident) and target.name in _NAMESPACES:
            ns = target.name
            if ns == "cel" and name == "bind" and n == 3:
                var = self._ident_name(args[0], "cel.bind")
                return Macro("bind", args[1], [var], [args[2]])
            return Call(f"{ns}.{name}", None, args)
        if name in ("all", "exists", "exists_one", "filter") and n == 2:
            v = self._ident_name(args[0], name)
            return Macro(name, target, [v], [args[1]])
        if name == "existsOne" and n == 2:
            v = self._ident_name(args[0], name)
            return Macro("exists_one", target, [v], [args[1]])
        if name == "map" and n in (2, 3):
            v = self._ident_name(args[0], name)
            return Macro("map", target, [v], args[1:])
        if name in ("all", "exists", "existsOne", "exists_one") and n == 3:
            v1 = self._ident_name(args[0], name)
            v2 = self._ident_name(args[1], name)
            nm = {"all": "all2", "exists": "exists2"}.get(name, "exists_one2")
            return Macro(nm, target, [v1, v2], [args[2]])
        if name in ("transformList", "transformMap", "transformMapEntry") and n in (3, 4):
            v1 = self._ident_name(args[0], name)
            v2 = self._ident_name(args[1], name)
            return Macro(name, target, [v1, v2], args[2:])
        if name == "sortBy" and n == 2:
            v = self._ident_name(args[0], name)
            return Macro("sortBy", target, [v], [args[1]])
        return Call(name, target, args)


def parse(src: str) -> Node:
    """Parse CEL source text into an AST. Raises CelSyntaxError."""
    return _Parser(src).parse()
