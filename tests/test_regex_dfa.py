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
