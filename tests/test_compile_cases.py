"""The reference compiler's own test cases (internal/test/testdata/compile, run by internal/compile/compile_test.go;
extracted by tests/golden/make_golden.py: compile_cases) against cerbos_b200/policy/compile.py, the front end whose rows
feed the table: every set of policy files the reference compiles must compile here, every set it rejects must be rejected
(imports that do not exist, unknown / ambiguous derived roles, undefined / cyclical / redefined variables and constants,
invalid identifiers, index lookups into V / C / G, rules without roles, scoped policies without their ancestors, script
conditions, YAML comments inside expressions).  The error kinds are compared by acceptance only -- messages and source
positions are presentation; schema errors are out of scope (schemas do not reach the decision path)."""
import pytest

from cerbos_b200.policy.compile import PolicyCompileError, build_rule_table
from cerbos_b200.table.flatten import flatten
from helpers import load_golden

CASES = load_golden("compile_cases.json")
SCHEMA_ONLY = {"invalid_schemas", "missing_schemas"}       # their only errors are "invalid schema": schema validation is not built


@pytest.mark.parametrize("tc", CASES, ids=lambda tc: tc["name"])
def test_reference_compile_case(tc):
    docs = [d for ds in tc["files"].values() for d in ds]
    kinds = {e["error"] for e in tc["wantErrors"]}
    if tc["name"] in SCHEMA_ONLY:
        assert kinds == {"invalid schema"}
        build_rule_table(docs)                   # (compiles: nothing else is wrong with these policies)
        return
    if not kinds:
        rt = build_rule_table(docs)
        assert rt.rows
        flatten(rt)                              # ... and the table builds
    else:
        with pytest.raises(PolicyCompileError):
            build_rule_table(docs)


def test_validate_identifier_known_answers():
    """internal/conditions/identifiers_test.go:13-47"""
    from cerbos_b200.policy.compile import validate_identifier
    for ok in ["_", "_0", "_x", "foo", "foo_bar", "foo42bar", "fooBar", "no", "x_", "x0", "yes"]:
        validate_identifier("variable", ok)
    for bad in ["", "0", "123", "false", "foo?", "in", "null", "true"]:
        with pytest.raises(PolicyCompileError):
            validate_identifier("constant", bad)
