"""RE2-style regular expressions -> byte-level DFA tables for the device (`matches`, cel-go strings overload).

cel-go's `matches` is an unanchored RE2 search (`regexp.MatchString`).  The device cannot run a backtracking or NFA
engine per request cheaply, so a constant pattern is compiled here, at table build, to a DFA over BYTES of the UTF-8
text -- `.` and negated classes expand to the UTF-8 encodings of a code point, so no decoding happens on the device --
bracketed by two virtual symbols, begin-of-text and end-of-text, which is what `^` / `\\A` and `$` / `\\z` consume:

    text  = BOT b1 b2 ... bn EOT          search = ANY* pattern ANY*       (ANY: all 258 symbols)

Supported: literals, `.`, classes with ranges / negation / `\\d \\w \\s` / POSIX `[:alpha:]`, escapes, groups (capturing,
non-capturing, named), alternation, `* + ? {m} {m,} {m,n}` (lazy forms too: laziness cannot change a yes/no answer),
`^ $ \\A \\z`, the flags `(?i)` `(?s)` at the start of the pattern or of a group.  Rejected with RegexUnsupported (the
table build then fails loudly, never a silent divergence): `\\b \\B`, Unicode classes `\\p{..}`, non-ASCII members of a
class or under `(?i)`, multi-line mode, more than 1024 DFA states.
"""
from __future__ import annotations

BOT, EOT, NSYM = 256, 257, 258
MAX_STATES = 1024
MAX_REPEAT = 64


class RegexUnsupported(Exception):
    pass


class RegexError(Exception):
    """Invalid pattern (a CEL error at evaluation time)."""


# ---- NFA fragments: states are ints; transitions: list of (symbol set as int bitmask, target); eps: list of targets
class _NFA:
    def __init__(self):
        self.trans = []   # state -> [(mask, to)]
        self.eps = []     # state -> [to]
        self.cond = []    # state -> [(BOT | EOT, to)]: zero-width, taken only at the begin / the end of the text

    def new(self):
        self.trans.append([])
        self.eps.append([])
        self.cond.append([])
        return len(self.trans) - 1


_ASCII_ALL = (1 << 128) - 1
_CLASSES = {
    "d": sum(1 << c for c in range(48, 58)),
    "w": sum(1 << c for c in list(range(48, 58)) + list(range(65, 91)) + list(range(97, 123)) + [95]),
    "s": sum(1 << c for c in (9, 10, 12, 13, 32)),
}
_POSIX = {
    "alpha": sum(1 << c for c in list(range(65, 91)) + list(range(97, 123))), "digit": _CLASSES["d"],
    "alnum": sum(1 << c for c in list(range(48, 58)) + list(range(65, 91)) + list(range(97, 123))),
    "upper": sum(1 << c for c in range(65, 91)), "lower": sum(1 << c for c in range(97, 123)),
    "space": sum(1 << c for c in (9, 10, 11, 12, 13, 32)), "punct": sum(1 << c for c in range(33, 127) if not chr(c).isalnum()),
    "xdigit": sum(1 << c for c in list(range(48, 58)) + list(range(65, 71)) + list(range(97, 103))), "word": _CLASSES["w"],
    "blank": (1 << 9) | (1 << 32), "cntrl": sum(1 << c for c in list(range(0, 32)) + [127]),
    "print": sum(1 << c for c in range(32, 127)), "graph": sum(1 << c for c in range(33, 127)),
}


def _fold(mask: int) -> int:
    for c in range(65, 91):
        if (mask >> c) & 1 or (mask >> (c + 32)) & 1:
            mask |= (1 << c) | (1 << (c + 32))
    return mask


class _Parser:
    def __init__(self, pat: str):
        self.p = pat
        self.i = 0
        self.nfa = _NFA()
        self.icase = False
        self.dotall = False

    # ---- helpers building fragments (start, end)
    def _sym(self, mask):
        a, b = self.nfa.new(), self.nfa.new()
        self.nfa.trans[a].append((mask, b))
        return a, b

    def _eps(self):
        a = self.nfa.new()
        return a, a

    def _assert(self, which):
        """`^` / `\\A` (which = BOT) and `$` / `\\z` (EOT) consume nothing: `^^a$$`, `$^` on the empty text, `(^|x)^a` all hold"""
        a, b = self.nfa.new(), self.nfa.new()
        self.nfa.cond[a].append((which, b))
        return a, b

    def _seq(self, f, g):
        self.nfa.eps[f[1]].append(g[0])
        return f[0], g[1]

    def _alt(self, frags):
        a, b = self.nfa.new(), self.nfa.new()
        for f in frags:
            self.nfa.eps[a].append(f[0])
            self.nfa.eps[f[1]].append(b)
        return a, b

    def _star(self, f):
        a, b = self.nfa.new(), self.nfa.new()
        self.nfa.eps[a] += [f[0], b]
        self.nfa.eps[f[1]] += [f[0], b]
        return a, b

    def _opt(self, f):
        a, b = self.nfa.new(), self.nfa.new()
        self.nfa.eps[a] += [f[0], b]
        self.nfa.eps[f[1]].append(b)
        return a, b

    def _bytes(self, bs):
        f = self._eps()
        for by in bs:
            f = self._seq(f, self._sym(1 << by))
        return f

    def _codepoint_set(self, ascii_mask: int, non_ascii: bool):
        """one code point: an ASCII byte of `ascii_mask`, or (non_ascii) any multi-byte UTF-8 sequence"""
        alts = []
        if ascii_mask:
            alts.append(self._sym(ascii_mask))
        if non_ascii:
            cont = sum(1 << c for c in range(0x80, 0xC0))
            for lead_lo, lead_hi, n in ((0xC2, 0xDF, 1), (0xE0, 0xEF, 2), (0xF0, 0xF4, 3)):
                f = self._sym(sum(1 << c for c in range(lead_lo, lead_hi + 1)))
                for _ in range(n):
                    f = self._seq(f, self._sym(cont))
                alts.append(f)
        if not alts:
            a, b = self.nfa.new(), self.nfa.new()   # matches nothing
            return a, b
        return alts[0] if len(alts) == 1 else self._alt(alts)

    # ---- grammar
    def parse(self):
        self._flags_prefix()
        f = self._alternation()
        if self.i != len(self.p):
            raise RegexError("unexpected )")
        return f

    def _flags_prefix(self):
        while self.p.startswith("(?", self.i):
            j = self.i + 2
            k = j
            while k < len(self.p) and self.p[k] in "imsU-":
                k += 1
            if k < len(self.p) and self.p[k] == ")" and k > j:
                self._apply_flags(self.p[j:k])
                self.i = k + 1
            else:
                break

    def _apply_flags(self, fl):
        on = True
        for ch in fl:
            if ch == "-":
                on = False
            elif ch == "i":
                self.icase = on
            elif ch == "s":
                self.dotall = on
            elif ch == "m":
                if on:
                    raise RegexUnsupported("multi-line mode (?m)")
            elif ch == "U":
                pass

    def _alternation(self):
        frags = [self._concat()]
        while self.i < len(self.p) and self.p[self.i] == "|":
            self.i += 1
            frags.append(self._concat())
        return frags[0] if len(frags) == 1 else self._alt(frags)

    def _concat(self):
        f = self._eps()
        while self.i < len(self.p) and self.p[self.i] not in "|)":
            f = self._seq(f, self._repeat())
        return f

    def _repeat(self):
        start = self.i
        atom_src = None
        f = self._atom()
        atom_src = (start, self.i)
        while self.i < len(self.p) and self.p[self.i] in "*+?{":
            ch = self.p[self.i]
            if ch == "{":
                j = self.p.find("}", self.i)
                body = self.p[self.i + 1:j] if j > 0 else ""
                parts = body.split(",")
                if j < 0 or not parts[0].isdigit() or len(parts) > 2 or (len(parts) == 2 and parts[1] and not parts[1].isdigit()):
                    break   # a literal '{'
                lo = int(parts[0])
                hi = lo if len(parts) == 1 else (int(parts[1]) if parts[1] else None)
                if lo > MAX_REPEAT or (hi is not None and (hi > MAX_REPEAT or hi < lo)):
                    raise RegexUnsupported("repetition count too large")
                self.i = j + 1
                f = self._counted(atom_src, lo, hi)
            else:
                self.i += 1
                if ch == "*":
                    f = self._star(f)
                elif ch == "+":
                    f = self._seq(f, self._star(self._reparse(atom_src)))
                else:
                    f = self._opt(f)
            if self.i < len(self.p) and self.p[self.i] == "?":   # lazy: same language
                self.i += 1
            atom_src = None if atom_src is None else (start, self.i)
            if self.i < len(self.p) and (self.p[self.i] in "*+?" or self._is_count(self.i)):
                # Go's regexp (Perl flags) does not stack repetition operators: a**, a*+, a{2}{3}, a??? are errors
                raise RegexError("invalid nested repetition operator")
        return f

    def _is_count(self, i):
        if not self.p.startswith("{", i):
            return False
        j = self.p.find("}", i)
        parts = (self.p[i + 1:j] if j > 0 else "").split(",")
        return j > 0 and parts[0].isdigit() and len(parts) <= 2 and (len(parts) == 1 or not parts[1] or parts[1].isdigit())

    def _reparse(self, src):
        """a fresh copy of the fragment whose source is p[src[0]:src[1]] (fragments are not shared)"""
        sub = _Parser(self.p[src[0]:src[1]])
        sub.nfa, sub.icase, sub.dotall = self.nfa, self.icase, self.dotall
        f = sub._alternation()
        return f

    def _counted(self, src, lo, hi):
        f = self._eps()
        for _ in range(lo):
            f = self._seq(f, self._reparse(src))
        if hi is None:
            f = self._seq(f, self._star(self._reparse(src)))
        else:
            for _ in range(hi - lo):
                f = self._seq(f, self._opt(self._reparse(src)))
        return f

    def _atom(self):
        ch = self.p[self.i]
        if ch == "(":
            self.i += 1
            saved = (self.icase, self.dotall)
            if self.p.startswith("?", self.i):
                if self.p.startswith("?:", self.i):
                    self.i += 2
                elif self.p.startswith("?P<", self.i) or self.p.startswith("?<", self.i):
                    j = self.p.find(">", self.i)
                    if j < 0:
                        raise RegexError("bad named group")
                    self.i = j + 1
                else:
                    j = self.i + 1
                    while j < len(self.p) and self.p[j] in "imsU-":
                        j += 1
                    if j < len(self.p) and self.p[j] == ":":
                        self._apply_flags(self.p[self.i + 1:j])
                        self.i = j + 1
                    elif j < len(self.p) and self.p[j] == ")":
                        self._apply_flags(self.p[self.i + 1:j])   # flags for the rest of the enclosing group
                        self.i = j + 1
                        return self._eps()
                    else:
                        raise RegexUnsupported("group syntax (?" + self.p[self.i + 1:self.i + 3])
            f = self._alternation()
            if self.i >= len(self.p) or self.p[self.i] != ")":
                raise RegexError("missing )")
            self.i += 1
            self.icase, self.dotall = saved
            return f
        if ch == "[":
            return self._cls()
        if ch == ".":
            self.i += 1
            mask = _ASCII_ALL if self.dotall else _ASCII_ALL & ~(1 << 10)
            return self._codepoint_set(mask, True)
        if ch == "^":
            self.i += 1
            return self._assert(BOT)
        if ch == "$":
            self.i += 1
            return self._assert(EOT)
        if ch == "\\":
            return self._escape_atom()
        if ch in "*+?":
            raise RegexError("missing argument to repetition operator")
        self.i += 1
        return self._literal(ch)

    def _literal(self, ch):
        bs = ch.encode("utf-8")
        if self.icase:
            if len(bs) > 1:
                if ch.lower() != ch.upper():
                    raise RegexUnsupported("case-insensitive match of a non-ASCII letter")
            elif ch.isalpha():
                return self._sym(_fold(1 << bs[0]))
        return self._bytes(bs)

    def _escape_char(self):
        """after a backslash at self.i: returns ('cls', mask, negated) | ('chr', str)"""
        self.i += 1
        if self.i >= len(self.p):
            raise RegexError("trailing backslash")
        ch = self.p[self.i]
        self.i += 1
        if ch in "dws":
            return ("cls", _CLASSES[ch], False)
        if ch in "DWS":
            return ("cls", _CLASSES[ch.lower()], True)
        simple = {"t": "\t", "n": "\n", "r": "\r", "f": "\f", "v": "\v", "a": "\a"}
        if ch in simple:
            return ("chr", simple[ch])
        if ch == "x":
            if self.p.startswith("{", self.i):
                j = self.p.find("}", self.i)
                if j < 0:
                    raise RegexError("bad \\x{")
                cp = int(self.p[self.i + 1:j], 16)
                self.i = j + 1
            else:
                cp = int(self.p[self.i:self.i + 2], 16)
                self.i += 2
            return ("chr", chr(cp))
        if ch in "pP":
            raise RegexUnsupported("Unicode class \\p")
        if ch in "bB":
            raise RegexUnsupported("word boundary \\b")
        if ch == "A":
            return ("sym", BOT)
        if ch == "z":
            return ("sym", EOT)
        if ch == "Q":
            j = self.p.find("\\E", self.i)
            lit = self.p[self.i:] if j < 0 else self.p[self.i:j]
            self.i = len(self.p) if j < 0 else j + 2
            return ("lit", lit)
        if ch.isalnum():
            raise RegexError(f"invalid escape \\{ch}")
        return ("chr", ch)

    def _escape_atom(self):
        r = self._escape_char()
        if r[0] == "cls":
            return self._codepoint_set(_ASCII_ALL & ~r[1], True) if r[2] else self._sym(r[1])
        if r[0] == "sym":
            return self._assert(r[1])
        if r[0] == "lit":
            f = self._eps()
            for ch in r[1]:
                f = self._seq(f, self._literal(ch))
            return f
        return self._literal(r[1])

    def _cls(self):
        self.i += 1
        neg = False
        if self.p.startswith("^", self.i):
            neg = True
            self.i += 1
        mask = 0
        first = True
        while True:
            if self.i >= len(self.p):
                raise RegexError("missing ]")
            ch = self.p[self.i]
            if ch == "]" and not first:
                self.i += 1
                break
            first = False
            if ch == "[" and self.p.startswith("[:", self.i):
                j = self.p.find(":]", self.i)
                name = self.p[self.i + 2:j] if j > 0 else ""
                nneg = name.startswith("^")
                name = name.lstrip("^")
                if name not in _POSIX:
                    raise RegexError("bad POSIX class")
                m = _POSIX[name]
                mask |= (_ASCII_ALL & ~m) if nneg else m
                if nneg:
                    raise RegexUnsupported("negated POSIX class inside a class")
                self.i = j + 2
                continue
            if ch == "\\":
                r = self._escape_char()
                if r[0] == "cls":
                    if r[2]:
                        raise RegexUnsupported("negated shorthand inside a class")
                    mask |= r[1]
                    continue
                if r[0] != "chr":
                    raise RegexUnsupported("escape inside a class")
                lo = r[1]
            else:
                lo = ch
                self.i += 1
            hi = lo
            if self.p.startswith("-", self.i) and not self.p.startswith("-]", self.i):
                self.i += 1
                if self.p[self.i] == "\\":
                    r = self._escape_char()
                    if r[0] != "chr":
                        raise RegexError("bad range")
                    hi = r[1]
                else:
                    hi = self.p[self.i]
                    self.i += 1
            if ord(lo) > 127 or ord(hi) > 127:
                raise RegexUnsupported("non-ASCII member of a character class")
            if ord(hi) < ord(lo):
                raise RegexError("bad range")
            for c in range(ord(lo), ord(hi) + 1):
                mask |= 1 << c
        if self.icase:
            mask = _fold(mask)
        if neg:
            return self._codepoint_set(_ASCII_ALL & ~mask, True)
        return self._sym(mask)


def _unquote(p: str) -> str:
    """\\Q..\\E -> the same text with its metacharacters escaped, so that a repetition operator after \\E binds to the last
    character only (as in Go's regexp), not to the whole quoted run"""
    out, i = [], 0
    while i < len(p):
        if p[i] == "\\" and i + 1 < len(p):
            if p[i + 1] == "Q":
                j = p.find("\\E", i + 2)
                lit = p[i + 2:] if j < 0 else p[i + 2:j]
                out.append("".join(c if (c.isalnum() or ord(c) > 127 or c in " _") else "\\" + c for c in lit))
                i = len(p) if j < 0 else j + 2
            else:
                out.append(p[i:i + 2])
                i += 2
        else:
            out.append(p[i])
            i += 1
    return "".join(out)


def compile_dfa(pattern: str):
    """-> dict(n_states, n_classes, start, classmap (258 ints), accept (list of bool), trans (n_states x n_classes ints))"""
    ps = _Parser(_unquote(pattern))
    try:
        frag = ps.parse()
    except (IndexError, ValueError) as e:
        raise RegexError(str(e)) from e
    nfa = ps.nfa
    # search = ANY* pattern ANY*
    any_mask = (1 << NSYM) - 1
    s0, acc = nfa.new(), nfa.new()
    nfa.trans[s0].append((any_mask, s0))
    nfa.eps[s0].append(frag[0])
    nfa.eps[frag[1]].append(acc)
    nfa.trans[acc].append((any_mask, acc))
    # symbol classes: symbols with the same membership in every transition set; BOT and EOT each on their own (the
    # construction treats them specially)
    sig = [0] * NSYM
    masks = sorted({m for st in nfa.trans for m, _ in st})
    for bit, m in enumerate(masks):
        for sym in range(NSYM):
            if (m >> sym) & 1:
                sig[sym] |= 1 << bit
    sig[BOT] |= 1 << len(masks)
    sig[EOT] |= 1 << (len(masks) + 1)
    class_of, reps = {}, []
    classmap = []
    for sym in range(NSYM):
        c = class_of.get(sig[sym])
        if c is None:
            c = len(reps)
            class_of[sig[sym]] = c
            reps.append(sym)
        classmap.append(c)

    def closure(states, at_begin, at_end):
        stack, seen = list(states), set(states)
        while stack:
            s = stack.pop()
            nxt = list(nfa.eps[s])
            for which, t in nfa.cond[s]:
                if (which == BOT and at_begin) or (which == EOT and at_end):
                    nxt.append(t)
            for t in nxt:
                if t not in seen:
                    seen.add(t)
                    stack.append(t)
        return frozenset(seen)

    def move(states, sym):
        nxt = set()
        for s in states:
            for m, t in nfa.trans[s]:
                if (m >> sym) & 1:
                    nxt.add(t)
        return nxt

    # a DFA state = (NFA states, "the position is the begin of the text"): `^` edges are followed right after BOT, `$`
    # edges when EOT arrives -- before and after it is consumed, so that what follows a `$` (another `$`, `\z`, a `^` on
    # the empty text) is still reached
    start = (closure({s0}, False, False), False)
    ids = {start: 0}
    order = [start]
    trans = []
    k = 0
    while k < len(order):
        cur, cur_begin = order[k]
        row = []
        for rep in reps:
            if rep == BOT:
                d = (closure(move(cur, rep), True, False), True)
            elif rep == EOT:
                pre = closure(cur, cur_begin, True)
                d = (closure(move(pre, rep), cur_begin, True), cur_begin)
            else:
                d = (closure(move(cur, rep), False, False), False)
            if d not in ids:
                if len(ids) >= MAX_STATES:
                    raise RegexUnsupported("pattern needs more than %d DFA states" % MAX_STATES)
                ids[d] = len(order)
                order.append(d)
            row.append(ids[d])
        trans.append(row)
        k += 1
    accept = [acc in st for st, _ in order]
    return {"n_states": len(order), "n_classes": len(reps), "start": 0, "classmap": classmap, "accept": accept, "trans": trans}


def dfa_words(d) -> list:
    """u64 words of the device table: [n_states | n_classes << 16 | start << 32], class map (258 bytes, padded to 33
    words), accept bitmap, transitions (u16, four per word, row-major)."""
    ns, nc = d["n_states"], d["n_classes"]
    words = [ns | (nc << 16) | (d["start"] << 32)]
    cm = bytes(d["classmap"]) + b"\0" * (264 - NSYM)
    words += [int.from_bytes(cm[i:i + 8], "little") for i in range(0, 264, 8)]
    for w in range((ns + 63) // 64):
        v = 0
        for b in range(64):
            if w * 64 + b < ns and d["accept"][w * 64 + b]:
                v |= 1 << b
        words.append(v)
    flat = [t for row in d["trans"] for t in row]
    flat += [0] * (-len(flat) % 4)
    for i in range(0, len(flat), 4):
        words.append(flat[i] | (flat[i + 1] << 16) | (flat[i + 2] << 32) | (flat[i + 3] << 48))
    return words


def dfa_match(d, text: str) -> bool:
    """Reference walk of the table (what the device does), for tests."""
    s = d["start"]
    for sym in [BOT] + list(text.encode("utf-8")) + [EOT]:
        s = d["trans"][s][d["classmap"][sym]]
    return d["accept"][s]
