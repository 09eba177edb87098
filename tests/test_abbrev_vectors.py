"""internal/conditions/cel_test.go:18-60 (TestExpandAbbrev): R / P / V / G (and C) are aliases of request.resource,
request.principal, variables, globals (constants).  A condition written with an alias must compile to exactly the table its
long form compiles to -- same attribute slots, same bytecode, same flat terms."""
import pytest

from cerbos_b200.policy.compile import build_rule_table
from cerbos_b200.table.flatten import flatten

PAIRS = [   # (abbreviated, expanded) -- the rows of TestExpandAbbrev, each inside a comparison so that it is a condition
    ('R.attr.department == "x"', 'request.resource.attr.department == "x"'),
    ('P.attr.department == "x"', 'request.principal.attr.department == "x"'),
    ('R.id == P.id', 'request.resource.id == request.principal.id'),
    ('V.is_admin', 'variables.is_admin'),
    ('G.environment == "test"', 'globals.environment == "test"'),
    ('C.limit > P.attr.n', 'constants.limit > request.principal.attr.n'),
]


def _blob(expr):
    pol = {"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {
        "resource": "leave_request", "version": "default",
        "variables": {"local": {"is_admin": '"admin" in request.principal.roles'}}, "constants": {"local": {"limit": 3}},
        "rules": [{"actions": ["a"], "effect": "EFFECT_ALLOW", "roles": ["*"], "condition": {"match": {"expr": expr}}}]}}
    ft = flatten(build_rule_table([pol]), globals_={"environment": "test"})
    return {k: bytes(v) for k, v in ft.sections.items() if k not in ("MANIFEST",)}, ft.manifest["slots"]


@pytest.mark.parametrize("short,full", PAIRS, ids=[p[0] for p in PAIRS])
def test_alias_and_long_form_compile_to_the_same_table(short, full):
    a, slots_a = _blob(short)
    b, slots_b = _blob(full)
    assert slots_a == slots_b
    assert a == b
