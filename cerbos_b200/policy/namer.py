"""Naming / scope helpers (host side).

Restates the small string rules of internal/namer/namer.go that the decision path
depends on: sanitize (:213-218 with the patterns at :17-20), FQNs (:104-160),
ScopeParents (:77-87), ScopeValue (:276-278).
"""
from __future__ import annotations

import re

# namer.go:17-20.  [[:alpha:]] / [[:word:]] / \w are ASCII classes in Go regexp.
_OLD_NAME = re.compile(r"^[A-Za-z][0-9A-Za-z_@.\-/]*(:[A-Za-z][0-9A-Za-z_@.\-/]*)*$")
_INVALID = re.compile(r"[^0-9A-Za-z_.]+")

DEFAULT_VERSION = "default"


def sanitize(v: str) -> str:
    if _OLD_NAME.match(v):
        return _INVALID.sub("_", v)
    return v


def with_scope(fqn: str, scope: str) -> str:
    return fqn if scope == "" else f"{fqn}/{scope}"


def resource_policy_fqn(resource: str, version: str, scope: str) -> str:
    return with_scope(f"cerbos.resource.{sanitize(resource)}.v{sanitize(version)}", scope)


def principal_policy_fqn(principal: str, version: str, scope: str) -> str:
    return with_scope(f"cerbos.principal.{sanitize(principal)}.v{sanitize(version)}", scope)


def role_policy_fqn(role: str, version: str, scope: str) -> str:
    if version == "":
        version = DEFAULT_VERSION
    return with_scope(f"cerbos.role.{sanitize(role)}.v{sanitize(version)}", scope)


def derived_roles_fqn(name: str) -> str:
    return f"cerbos.derived_roles.{sanitize(name)}"


def policy_key_from_fqn(fqn: str) -> str:
    return fqn[len("cerbos."):] if fqn.startswith("cerbos.") else fqn


def scope_value(scope: str) -> str:
    return scope[1:] if scope.startswith(".") else scope


def scope_parents(scope: str):
    """Every dotted prefix of `scope` from longest to "" (namer.go:77-87)."""
    out = []
    for i in range(len(scope) - 1, -1, -1):
        if scope[i] == "." or i == 0:
            out.append(scope[:i])
    return out
