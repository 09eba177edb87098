"""internal/ruletable/ruletable_test.go:34-392 (TestRuleTableManager) as a known-answer scenario: a policy store evolves file by
file and after every event one CheckInput must get the recorded effect.  The reference applies storage events incrementally
(manager.go:126-181); here every state is a full build-then-swap of the table (what Engine.reload does, manager.go:88-124
semantics): a state whose policies do not compile keeps the previous table -- "maintain valid state".  Each state is checked
through oracle #1, oracle #2 and the kernel core."""
import pytest

from cerbos_b200.encode import Encoder
from cerbos_b200.policy.compile import build_rule_table
from cerbos_b200.table.flatten import flatten
from hostsim import driver as hostsim
from oracle import cref
from oracle.check import CheckOracle

ALLOW, DENY = 1, 2


def _rp(resource, rules, **extra):
    return {"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": dict({"resource": resource, "version": "default", "rules": rules}, **extra)}


class Store:
    """files -> policy documents; build() = the table of the current files, or None when they do not compile"""

    def __init__(self):
        self.files = {}
        self.current = None        # (rule table, flattened table) that answers requests

    def put(self, name, doc):
        self.files[name] = doc
        self._rebuild()

    def delete(self, name):
        del self.files[name]
        self._rebuild()

    def _rebuild(self):
        try:
            rt = build_rule_table(list(self.files.values()))
            self.current = (rt, flatten(rt))
            self.last_build_ok = True
        except Exception:          # noqa: BLE001 -- an invalid store state: the previous table stays (manager.go:108-112)
            self.last_build_ok = False

    def effect(self, inp):
        rt, ft = self.current
        want = CheckOracle(rt).check(inp)["actions"][inp["actions"][0]]["effect"]
        b = Encoder(ft.manifest).encode([inp])
        assert cref.check(ft.blob, b.columns, 1, 1)[0, 0] == want
        for mode in (0, 1):
            assert hostsim.check(ft.blob, b.columns, 1, 1, mode=mode)[0, 0] == want
        return want


def test_rule_table_manager_scenario():
    s = Store()
    rock = {"requestId": "1", "resource": {"kind": "rock", "id": "1"}, "principal": {"id": "sam", "roles": ["user"]}, "actions": ["throw"]}
    # a simple, valid policy: ALLOW
    s.put("resource_policies/rock.yaml", _rp("rock", [{"actions": ["throw"], "roles": ["user"], "effect": "EFFECT_ALLOW"}]))
    assert s.effect(rock) == ALLOW
    # maintain_valid_state_on_missing_derived_role: the update references derived roles that do not exist -> still ALLOW
    s.put("resource_policies/rock.yaml", _rp("rock", [{"actions": ["throw"], "derivedRoles": ["special_user"], "effect": "EFFECT_ALLOW"}],
                                             importDerivedRoles=["special_roles"]))
    assert not s.last_build_ok               # (the import cannot be resolved: compile error, compile.go import checks)
    assert s.effect(rock) == ALLOW
    # adding_missing_derived_role_re_enables_updates: the derived role's condition is false -> DENY
    dr = lambda expr: {"apiVersion": "api.cerbos.dev/v1", "derivedRoles": {"name": "special_roles", "definitions": [   # noqa: E731
        {"name": "special_user", "parentRoles": ["user"], "condition": {"match": {"expr": expr}}}]}}
    s.put("derived_roles/special_roles.yaml", dr("true == false"))
    assert s.last_build_ok and s.effect(rock) == DENY
    # updating_derived_role_affects_rule_table
    s.put("derived_roles/special_roles.yaml", dr("true == true"))
    assert s.effect(rock) == ALLOW
    # adding_and_referencing_export_const_affects_rule_table
    consts = lambda v: {"apiVersion": "api.cerbos.dev/v1", "exportConstants": {"name": "special_constants", "definitions": {"FlakeyTrue": v}}}   # noqa: E731
    s.put("export_constants/special_constants.yaml", consts(True))
    s.put("resource_policies/rock.yaml", _rp("rock", [{"actions": ["throw"], "roles": ["user"], "effect": "EFFECT_ALLOW",
                                                       "condition": {"match": {"expr": "C.FlakeyTrue == true"}}}],
                                             constants={"import": ["special_constants"]}))
    assert s.effect(rock) == ALLOW
    s.put("export_constants/special_constants.yaml", consts(False))
    assert s.effect(rock) == DENY
    # deleting_role_policy_does_not_restore_parent_roles
    doc = {"requestId": "role-parent-delete", "resource": {"kind": "document", "id": "1"}, "principal": {"id": "sam", "roles": ["employee"]}, "actions": ["view"]}
    s.put("resource_policies/document.yaml", _rp("document", [{"actions": ["view"], "roles": ["user"], "effect": "EFFECT_ALLOW"}]))
    s.put("role_policies/employee.yaml", {"apiVersion": "api.cerbos.dev/v1", "rolePolicy": {"role": "employee", "version": "default", "parentRoles": ["user"]}})
    assert s.effect(doc) == ALLOW
    s.delete("role_policies/employee.yaml")
    s.put("resource_policies/unrelated_parent_role_control.yaml", _rp("unrelated_parent_role_control", [{"actions": ["noop"], "roles": ["admin"], "effect": "EFFECT_ALLOW"}]))
    unrelated = {"requestId": "unrelated-update-control", "resource": {"kind": "unrelated_parent_role_control", "id": "1"},
                 "principal": {"id": "sam", "roles": ["admin"]}, "actions": ["noop"]}
    assert s.effect(unrelated) == ALLOW
    assert s.effect(doc) == DENY
