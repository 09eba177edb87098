"""CheckInput -> CEL activation values -- TEST INFRASTRUCTURE (oracle).

Restates checkInputToRequest (internal/ruletable/ruletable.go:1227-1245) and the
activation built by buildEvalVars (:1303-1317) over the message shapes of
api/public/cerbos/engine/v1/engine.proto (Request, Request.Principal,
Request.Resource, AuxData, Runtime), with the camelCase field aliases of
internal/conditions/types/jsonfield.go:23-29.
"""
from __future__ import annotations

from .celeval import AbsentMsg, CelMap, Msg, from_json

_P_ALIASES = {"policyVersion": "policy_version"}
_REQ_ALIASES = {"auxData": "aux_data"}
_RT_ALIASES = {"effectiveDerivedRoles": "effective_derived_roles"}


def _attr(d):
    return CelMap((str(k), from_json(v)) for k, v in (d or {}).items())


def scope_value(scope: str) -> str:
    """namer.ScopeValue (internal/namer/namer.go:276-278)."""
    return scope[1:] if scope.startswith(".") else scope


def build_request(inp: dict) -> Msg:
    """inp is a CheckInput in protojson form (camelCase or snake_case keys)."""
    p = inp.get("principal") or {}
    r = inp.get("resource") or {}
    aux = inp.get("auxData", inp.get("aux_data"))
    principal = Msg("cerbos.engine.v1.Request.Principal", {
        "id": p.get("id", ""),
        "roles": list(p.get("roles") or []),
        "attr": _attr(p.get("attr")),
        "policy_version": p.get("policyVersion", p.get("policy_version", "")) or "",
        "scope": scope_value(p.get("scope", "") or ""),
    }, _P_ALIASES)
    resource = Msg("cerbos.engine.v1.Request.Resource", {
        "kind": r.get("kind", ""),
        "id": r.get("id", ""),
        "attr": _attr(r.get("attr")),
        "policy_version": r.get("policyVersion", r.get("policy_version", "")) or "",
        "scope": scope_value(r.get("scope", "") or ""),
    }, _P_ALIASES)
    if aux is None:
        aux_msg = AbsentMsg("<absent>", {"jwt": CelMap()})
    else:
        aux_msg = Msg("cerbos.engine.v1.AuxData", {"jwt": _attr(aux.get("jwt"))})
    return Msg("cerbos.engine.v1.Request", {"principal": principal, "resource": resource, "aux_data": aux_msg},
               _REQ_ALIASES)


def build_runtime(effective_derived_roles) -> Msg:
    return Msg("cerbos.engine.v1.Runtime", {"effective_derived_roles": sorted(effective_derived_roles or [])},
               _RT_ALIASES)


def build_activation(request: Msg, constants=None, variables=None, globals_=None, runtime=None) -> dict:
    c = constants if constants is not None else CelMap()
    v = variables if variables is not None else CelMap()
    g = globals_ if globals_ is not None else CelMap()
    return {
        "request": request,
        "R": request.fields["resource"],
        "P": request.fields["principal"],
        "runtime": runtime if runtime is not None else (lambda: build_runtime(None)),
        "constants": c, "C": c,
        "variables": v, "V": v,
        "globals": g, "G": g,
    }
