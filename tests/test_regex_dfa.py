"""cel/regex_dfa.py: RE2-style patterns -> byte-level DFA tables (the device's `matches`), against Python's `re` on
patterns both engines read the same way (Python's `$` also matches before a trailing newline: no test text ends in one)."""
import re

import pytest

from cerbos_b200.cel.regex_dfa import RegexError, RegexUnsupported, compile_dfa, dfa_match, dfa_words

PATTERNS = [r"^[mM].*g$", r"^comm", r"foo|bar", r"a+b*c?", r"^(ab|cd){2,3}$", r"\d{3}-\d{4}", r"^\w+@\w+\.(com|org)$", r"[^a-c]x", r"^.$", r"^..$",
            r"é+", r"(?i)hello", r"^$", r"a.c", r"x{2,}", r"\.", r"^a|b$", r"(?i)[a-c]z", r"^(?:a|b)*c", r"\s+end$",
            r"[a-z0-9._%+-]+@[a-z0-9.-]+\.[a-z]{2,}", r"^/api/v[0-9]+/users/[^/]+$", r"a\x41b", r"a*?b", r"^.{3}$", r"[\d-]+", r"[]a]", r"^\D+$",
            r"^(a|ab)(c|bcd)(d*)$", r"(?s)a.c", r"^\S+\s\S+$", r"(ab)+$", r"^[^@]+@[^@]+$"]
TEXTS = ["marketing", "Mg", "communications", "foo", "xbar", "ac", "aabbc", "abab", "abcdab", "555-1234", "a@b.com", "a@b.net", "dx", "ax", "é", "éé",
         "ééé", "HeLLo", "", "abc", "a\nc", "xx", "x", "a.b", "b", "a", "cz", "Az", "aabc", "the   end", "john.doe@example.com", "/api/v2/users/42",
         "/api/v2/users/4/2", "aAb", "axb", "ab", "日本語", "12-34", "]", "abc def", "abcd", "a@b@c"]


def test_dfa_agrees_with_python_re():
    n = 0
    for p in PATTERNS:
        d = compile_dfa(p)
        assert d["n_states"] <= 1024 and len(dfa_words(d)) > 34
        for t in TEXTS:
            assert dfa_match(d, t) == (re.search(p, t) is not None), (p, t)
            n += 1
    assert n == len(PATTERNS) * len(TEXTS)


def test_posix_classes_and_quoting():
    assert dfa_match(compile_dfa(r"^[[:alpha:]]+$"), "abcXYZ") and not dfa_match(compile_dfa(r"^[[:alpha:]]+$"), "ab1")
    assert dfa_match(compile_dfa(r"\Qa.b\E"), "xa.by") and not dfa_match(compile_dfa(r"\Qa.b\E"), "axb")
    assert dfa_match(compile_dfa(r"^a\z"), "a") and not dfa_match(compile_dfa(r"\Aa$"), "ba")


@pytest.mark.parametrize("p", [r"\bword\b", r"\p{Greek}+", r"(?i)é", r"[é]", r"(?m)^a$", r"(a|b|c|d|e|f){40}x{40}y{40}z{40}(a|b|c|d|e|f){40}"])
def test_rejected_loudly(p):
    with pytest.raises(RegexUnsupported):
        compile_dfa(p)


@pytest.mark.parametrize("p", ["[", "(", "a)", "*a", r"\8", "a{2,1}"])
def test_invalid_patterns(p):
    with pytest.raises((RegexError, RegexUnsupported)):
        compile_dfa(p)


# ---- random patterns: the DFA compiler against oracle #1's reading of RE2 (oracle/celeval.py: _re2_to_python + `re`) ------
_ATOMS = ["a", "b", "c", "x", "1", " ", "-", "_", ".", "\\.", "\\d", "\\w", "\\s", "\\D", "\\W", "\\S", "[a-c]", "[^a-c]", "[[:alpha:]]", "[[:digit:]x]",
          "[\\d-]", "[^\\s]", "é", "日", "\\n", "\\t", "[]a]", "[a\\]]", "\\x41", "\\Qa.b\\E", "[[:space:]]", "[[:punct:]]", "[[:upper:][:digit:]]",
          "\\v", "\\f", "[\\w.]", "\\-", "\\{", "a{,2}"]
_TEXT = ["a", "b", "c", "x", "1", " ", "-", "_", ".", "A", "B", "\n", "\t", "\v", "\f", "é", "É", "日", "]", "{", ",", "2", "!", "\r"]


def _rand_pattern(r, d=0):
    k = r.randrange(10 if d < 3 else 3)
    if k < 3:
        return r.choice(_ATOMS)
    if k == 3:
        return _rand_pattern(r, d + 1) + _rand_pattern(r, d + 1)
    if k == 4:
        return _rand_pattern(r, d + 1) + r.choice(["*", "+", "?", "{2}", "{1,3}", "{2,}", "*?", "+?", "??"])
    if k == 5:
        return "(" + _rand_pattern(r, d + 1) + "|" + _rand_pattern(r, d + 1) + ")"
    if k == 6:
        return "(?:" + _rand_pattern(r, d + 1) + ")"
    if k == 7:
        return r.choice(["^", "\\A", ""]) + _rand_pattern(r, d + 1) + r.choice(["$", "\\z", ""])
    if k == 8:
        fl = r.choice(["(?i)", "(?s)", "(?i:", "(?s:", "(?is)"])
        return fl + _rand_pattern(r, d + 1) + (")" if fl.endswith(":") else "")
    return _rand_pattern(r, d + 1) + _rand_pattern(r, d + 1) + _rand_pattern(r, d + 1)


@pytest.mark.parametrize("seed", range(4))
def test_random_patterns_against_the_oracle(seed):
    """Same verdict on validity (stacked repetition operators, bad groups ... are errors in Go's regexp) and the same
    answer on every text -- texts with newlines, \\v, non-ASCII letters, a trailing newline (where `$` must not match)."""
    import random
    from oracle.celeval import CelError, _regex
    r = random.Random(4200 + seed)
    compared = invalid = 0
    for _ in range(500):
        p = _rand_pattern(r)
        try:
            d, perr = compile_dfa(p), False
        except RegexUnsupported:
            continue
        except RegexError:
            d, perr = None, True
        try:
            o, oerr = _regex(p), False
        except CelError as e:
            if "not supported by this oracle" in str(e):
                continue
            o, oerr = None, True
        assert perr == oerr, p
        if perr:
            invalid += 1
            continue
        for _ in range(20):
            t = "".join(r.choice(_TEXT) for _ in range(r.randrange(0, 7)))
            assert dfa_match(d, t) == (o.search(t) is not None), (p, t)
            compared += 1
    assert compared > 8000 and invalid > 3


def test_assertions_are_zero_width():
    for p, t, want in [(r"^^a$$", "a", True), (r"\A\Aa\z$", "a", True), (r"$^", "", True), (r"$^", "a", False), (r"(^|x)^a", "a", True),
                       (r"a$\z", "ba", True), (r"^$?a", "a", True), (r"a^b", "ab", False), (r"a$b", "ab", False)]:
        assert dfa_match(compile_dfa(p), t) == want, (p, t)


def test_quoted_run_takes_repetition_on_its_last_character():
    d = compile_dfa(r"^\Qa.b\E*$")
    assert dfa_match(d, "a.bbb") and dfa_match(d, "a.") and not dfa_match(d, "a.ba.b")


@pytest.mark.parametrize("p", ["a**", "a*+", "a{2}{3}", "a+?*", "a???"])
def test_stacked_repetition_is_invalid(p):
    with pytest.raises(RegexError):
        compile_dfa(p)
