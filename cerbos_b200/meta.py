"""Decision metadata: decoding of the device's metadata plane (cgpu_check_meta) into the strings of
``CheckOutput.ActionEffect.policy / scope`` and ``CheckOutput.effectiveDerivedRoles``
(internal/ruletable/ruletable.go:753-782, 913-922, 936-979, 1082-1148; names per internal/namer/namer.go:104-160).

The device reports, per (request, action), the id of the deciding scope, where the policy name comes from and -- for a
role policy -- the role; per request the first scope of each chain and a bit set of derived role names.  Only ids travel:
the names are assembled here from the table's MANIFEST dictionaries and the request's own strings.
"""
from __future__ import annotations

import numpy as np

from .policy import namer
from .table import layout as L

REQUEST_META_DTYPE = np.dtype([("principal_first_scope", "<u2"), ("resource_first_scope", "<u2"), ("flags", "<u4"),
                               ("effective_derived_roles", "<u8")])
NO_POLICY_MATCH = "NO_MATCH"
NO_MATCH_SCOPE_PERMISSIONS = "NO_MATCH_FOR_SCOPE_PERMISSIONS"


def decode_action(word: int, rm, manifest: dict, principal_id: str, kind: str, p_ver: str, r_ver: str):
    """-> (policy, scope) strings of one decision."""
    scope_id, src, role = word & 0xFFFF, (word >> 16) & 0xFF, (word >> 24) & 0xFF
    scopes = manifest["scopes"]
    scope = "" if scope_id == 0xFFFF else scopes[scope_id]
    if src == L.META_SRC["PRINCIPAL_POLICY"]:
        first = scopes[int(rm["principal_first_scope"])]
        policy = namer.policy_key_from_fqn(namer.principal_policy_fqn(principal_id, p_ver, first))
    elif src == L.META_SRC["RESOURCE_POLICY"]:
        first = scopes[int(rm["resource_first_scope"])]
        policy = namer.policy_key_from_fqn(namer.resource_policy_fqn(kind, r_ver, first))
    elif src == L.META_SRC["ROLE_POLICY"]:
        policy = namer.policy_key_from_fqn(namer.role_policy_fqn(manifest["roles"][role], r_ver, scope))
    elif src == L.META_SRC["NO_MATCH_FOR_SCOPE_PERMISSIONS"]:
        policy = NO_MATCH_SCOPE_PERMISSIONS
    else:
        policy = NO_POLICY_MATCH
    return policy, scope


def decode_edr(mask: int, manifest: dict):
    names = manifest.get("derived_roles") or []
    return [nm for i, nm in enumerate(names) if (mask >> i) & 1]
