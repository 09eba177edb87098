"""Rule-table data model (host side).

Python mirror of the parts of ``runtimev1.RuleTable`` the decision path reads
(api/private/cerbos/runtime/v1/runtime.proto ``RuleTable``, ``RuleTable.RuleRow``,
``Condition``, ``Expr``) plus ``index.Row`` (internal/ruletable/index/index.go:70-89).
This is what the flattener (cerbos_b200/table/flatten.py) consumes; in the Go
integration the same information comes from the reference's own proto structs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

from ..cel.ast import Node

EFFECT_ALLOW = 1   # api/public/cerbos/effect/v1/effect.proto
EFFECT_DENY = 2
EFFECT_NO_MATCH = 3

KIND_RESOURCE = "RESOURCE"
KIND_PRINCIPAL = "PRINCIPAL"

SP_UNSPECIFIED = 0
SP_OVERRIDE_PARENT = 1
SP_REQUIRE_PARENTAL_CONSENT = 2

_SP_NAMES = {
    None: SP_UNSPECIFIED, "": SP_UNSPECIFIED, "SCOPE_PERMISSIONS_UNSPECIFIED": SP_UNSPECIFIED,
    "SCOPE_PERMISSIONS_OVERRIDE_PARENT": SP_OVERRIDE_PARENT,
    "SCOPE_PERMISSIONS_REQUIRE_PARENTAL_CONSENT_FOR_ALLOWS": SP_REQUIRE_PARENTAL_CONSENT,
    0: 0, 1: 1, 2: 2,
}
_EFFECT_NAMES = {"EFFECT_ALLOW": EFFECT_ALLOW, "EFFECT_DENY": EFFECT_DENY, 1: 1, 2: 2}


def parse_scope_permissions(v) -> int:
    return _SP_NAMES[v]


def parse_effect(v) -> int:
    return _EFFECT_NAMES[v]


@dataclass
class Expr:
    original: str
    ast: Node


@dataclass
class Cond:
    """Condition tree: op in {'expr','all','any','none'}."""
    op: str
    expr: Optional[Expr] = None
    children: list = field(default_factory=list)


@dataclass
class Variable:
    name: str
    expr: Expr


@dataclass
class Params:
    """index.rowParams: ordered variable programs + constants, cached per request by `key`."""
    key: str
    variables: list  # [Variable] in dependency order
    constants: dict  # name -> JSON value (google.protobuf.Value semantics)


@dataclass
class DerivedRole:
    name: str
    parent_roles: list
    condition: Optional[Cond]
    params: Params
    origin_fqn: str


@dataclass
class Row:
    origin_fqn: str = ""
    resource: str = ""
    role: str = ""
    action: Optional[str] = None            # ActionSet.action
    allow_actions: Optional[list] = None    # ActionSet.allow_actions (role policies)
    condition: Optional[Cond] = None
    dr_condition: Optional[Cond] = None
    effect: int = 0
    scope: str = ""
    scope_permissions: int = SP_UNSPECIFIED
    version: str = ""
    origin_derived_role: str = ""
    name: str = ""
    principal: str = ""
    params: Optional[Params] = None
    dr_params: Optional[Params] = None
    evaluation_key: str = ""
    policy_kind: str = KIND_RESOURCE
    from_role_policy: bool = False
    no_match_for_scope_permissions: bool = False
    # rule outputs (RuleRow.emit_output.when, runtime.proto:260-267): expressions evaluated to a VALUE when the row is visited
    # and its condition is / is not satisfied (ruletable.go:1065-1106)
    emit_activated: Optional["Expr"] = None
    emit_not_met: Optional["Expr"] = None


@dataclass
class RuleTable:
    rows: list = field(default_factory=list)
    # scope -> role -> [direct parent roles]   (RuleTable.scope_parent_roles)
    scope_parent_roles: dict = field(default_factory=dict)
    # resource policy FQN -> {derived role name -> DerivedRole}  (RuleTable.policy_derived_roles)
    policy_derived_roles: dict = field(default_factory=dict)
