"""Structural CPU restatement of the CheckResources decision path -- TEST INFRASTRUCTURE (oracle #1).

Follows the reference function by function over string-keyed rule rows:

  RuleTable.check                 internal/ruletable/ruletable.go:785-1155
  GetAllScopes / ScopeParents     ruletable.go:611-645 ; internal/namer/namer.go:77-87
  indexRules (scope maps, perms)  ruletable.go:563-601
  Impl.GetRows + role-policy DENY synthesis   internal/ruletable/index/index.go:564-801
  Row.Matches                     index.go:91-116
  AddParentRoles / closure        index.go:805-881
  ScopedPrincipalExists / ScopedResourceExists   index.go:1089-1172
  SatisfiesCondition / variables  ruletable.go:1284-1499
  defaults                        internal/evaluator/evaluator.go:99-113

CEL evaluation is oracle/celeval.py.  Pinned against the reference's engine goldens
(tests/golden/engine_cases.json) by tests/test_oracle_goldens.py: effect, policy and
scope of every decision plus effectiveDerivedRoles.
"""
from __future__ import annotations

from cerbos_b200.policy import namer
from cerbos_b200.policy.globs import key_matches
from cerbos_b200.policy.model import (Cond, EFFECT_ALLOW, EFFECT_DENY, EFFECT_NO_MATCH, KIND_PRINCIPAL, KIND_RESOURCE,
                                      Row, RuleTable, SP_OVERRIDE_PARENT, SP_REQUIRE_PARENTAL_CONSENT,
                                      SP_UNSPECIFIED)

from .activation import build_activation, build_request, build_runtime
from .celeval import CelError, CelMap, Evaluator, Timestamp, UInt, from_json

NO_POLICY_MATCH = "NO_MATCH"
NO_MATCH_SCOPE_PERMISSIONS = "NO_MATCH_FOR_SCOPE_PERMISSIONS"


def native_to_cel(v):
    """Go-native config values (evaluator.Conf.Globals) -> CEL: ints stay ints."""
    if v is None or isinstance(v, (bool, str, float, bytes)):
        return v
    if isinstance(v, int):
        return v
    if isinstance(v, (list, tuple)):
        return [native_to_cel(x) for x in v]
    if isinstance(v, dict):
        return CelMap((str(k), native_to_cel(x)) for k, x in v.items())
    raise TypeError(type(v))


def to_json_value(v):
    """cel-go ConvertToNative(*structpb.Value): numbers are doubles, bytes base64, map keys their string form, timestamps
    RFC 3339, durations seconds with an "s"; types without a JSON form fail."""
    import base64
    from .celeval import Duration, Timestamp, UInt, conv_string
    if v is None or isinstance(v, (bool, str)):
        return v
    if isinstance(v, (int, UInt, float)):
        return float(v)
    if isinstance(v, bytes):
        return base64.b64encode(v).decode("ascii")
    if isinstance(v, list):
        return [to_json_value(x) for x in v]
    if isinstance(v, CelMap):
        return {conv_string(k): to_json_value(x) for k, x in v.items()}
    if isinstance(v, (Timestamp, Duration)):
        return conv_string(v)
    raise CelError(f"no protobuf value for {type(v).__name__}")


class CheckOracle:
    def __init__(self, rt: RuleTable, globals_=None, default_version="default", default_scope="",
                 lenient_scope_search=False):
        self.rt = rt
        self.rows = [r for r in rt.rows]
        self.globals = native_to_cel(globals_ or {})
        self.default_version = default_version
        self.default_scope = default_scope
        self.lenient = lenient_scope_search
        # indexRules (ruletable.go:563-601)
        self.principal_scopes, self.resource_scopes, self.scope_perms = set(), set(), {}
        for r in self.rows:
            if r.scope_permissions != SP_UNSPECIFIED:
                self.scope_perms[r.scope] = r.scope_permissions
            (self.principal_scopes if r.policy_kind == KIND_PRINCIPAL else self.resource_scopes).add(r.scope)
        self.has_role_policy_rules = any(r.allow_actions for r in self.rows)
        # compileParentRoleAncestors (index.go:842-881)
        self.parent_closure = {}
        for scope, roles in rt.scope_parent_roles.items():
            comp = {}
            for role in roles:
                acc, visited = [], set()
                self._collect(scope, role, acc, visited)
                comp[role] = acc
            self.parent_closure[scope] = comp

    def _collect(self, scope, role, acc, visited):
        if role in visited:
            return
        visited.add(role)
        for pr in self.rt.scope_parent_roles.get(scope, {}).get(role, []):
            if pr not in acc:
                acc.append(pr)
            self._collect(scope, pr, acc, visited)

    # ------------------------------------------------------------- index queries
    def add_parent_roles(self, scope, roles):
        out = list(roles)
        comp = self.parent_closure.get(scope)
        if not comp:
            return out
        for role in roles:
            out.extend(comp.get(role, []))
        return out

    def get_all_scopes(self, kind, scope, name, version):
        scope_map = self.principal_scopes if kind == KIND_PRINCIPAL else self.resource_scopes
        fqn_fn = namer.principal_policy_fqn if kind == KIND_PRINCIPAL else namer.resource_policy_fqn
        scopes, first_key, first_fqn = [], "", ""
        if scope in scope_map:
            first_fqn = fqn_fn(name, version, scope)
            first_key = namer.policy_key_from_fqn(first_fqn)
            scopes.append(scope)
        elif not self.lenient:
            return [], "", ""
        for s in namer.scope_parents(scope):
            if s in scope_map:
                scopes.append(s)
                if first_key == "":
                    first_fqn = fqn_fn(name, version, s)
                    first_key = namer.policy_key_from_fqn(first_fqn)
        return scopes, first_key, first_fqn

    def scoped_principal_exists(self, version, scopes):
        return any(r.policy_kind == KIND_PRINCIPAL and r.version == version and r.scope in scopes for r in self.rows)

    def scoped_resource_exists(self, version, resource, scopes):
        return any(r.policy_kind == KIND_RESOURCE and r.version == version and r.scope in scopes
                   and key_matches(r.resource, resource) for r in self.rows)

    def get_rows(self, version, resource, scopes, roles, actions):
        res, seen = [], set()

        def add(r):
            if id(r) not in seen:
                seen.add(id(r))
                res.append(r)

        if not any(r.version == version for r in self.rows):
            return res
        for scope in scopes:
            scope_version = [r for r in self.rows if r.scope == scope and r.version == version]
            if not scope_version:
                continue
            scope_resource = [r for r in scope_version if key_matches(r.resource, resource)]
            if not scope_resource:
                continue
            for role in roles:
                role_match = lambda r: key_matches(r.role, role)  # noqa: E731
                if not any(role_match(r) for r in self.rows):
                    continue
                role_resource = [r for r in scope_resource if role_match(r)]
                if self.has_role_policy_rules:
                    role_scope = [r for r in scope_version if role_match(r)]
                    if any(r.allow_actions for r in role_scope):
                        ars = [r for r in role_resource if r.allow_actions]
                        for action in actions:
                            matched = [ar for ar in ars for a in ar.allow_actions if key_matches(a, action)]
                            if not matched:
                                add(Row(action=action, origin_fqn=namer.role_policy_fqn(role, version, scope),
                                        resource=resource, role=role, effect=EFFECT_DENY, scope=scope,
                                        version=version, policy_kind=KIND_RESOURCE, from_role_policy=True,
                                        no_match_for_scope_permissions=True))
                            else:
                                for ar in matched:
                                    if ar.condition is not None:
                                        add(Row(action=action, origin_fqn=ar.origin_fqn, resource=resource,
                                                condition=Cond("none", children=[ar.condition]), role=ar.role,
                                                effect=EFFECT_DENY, scope=scope,
                                                scope_permissions=SP_REQUIRE_PARENTAL_CONSENT, version=version,
                                                evaluation_key=ar.evaluation_key, policy_kind=KIND_RESOURCE,
                                                from_role_policy=True))
                for action in actions:
                    for r in role_resource:
                        if r.action is not None and not r.allow_actions and key_matches(r.action, action):
                            add(r)
        return res

    @staticmethod
    def row_matches(r: Row, pt, scope, action, principal_id, roles):
        if r.policy_kind != pt:
            return False
        if pt == KIND_PRINCIPAL and r.principal != principal_id:
            return False
        if scope != r.scope:
            return False
        if r.role != "*" and r.role not in roles:
            return False
        a = r.action or ""
        return a == action or key_matches(a, action)

    # ------------------------------------------------------------- CEL glue
    def _eval_vars(self, ctx, params):
        """evaluatePrograms (ruletable.go:1319-1344): eager, in order, erroring variable left unset."""
        constants = CelMap((k, from_json(v)) for k, v in (params.constants or {}).items())
        variables = CelMap()
        for var in params.variables:
            try:
                variables.put(var.name, self._eval(ctx, var.expr, constants, variables))
            except CelError:
                continue
        return constants, variables

    def _eval(self, ctx, expr, constants, variables):
        act = build_activation(ctx["request"], constants, variables, self.globals,
                               runtime=lambda: build_runtime(ctx["edr"]))
        return Evaluator(act, ctx["now"]).eval(expr.ast)

    def _output_value(self, ctx, expr, constants, variables):
        """evaluateProtobufValueCELExpr (ruletable.go:1443-1465): the expression's value as a protobuf Value (here: its JSON
        shape).  A CEL evaluation error gives no value at all (evaluateCELExpr returns nil for an error value, :1477-1482)."""
        try:
            v = self._eval(ctx, expr, constants or CelMap(), variables or CelMap())
        except CelError:
            return None
        try:
            return to_json_value(v)
        except CelError:
            return "<failed to convert evaluation to protobuf value>"

    def satisfies(self, ctx, cond, constants, variables):
        """SatisfiesCondition (ruletable.go:1346-1441): error / non-bool -> false.  (Its `error` return is for failures that are
        not CEL evaluation errors: evaluateCELExpr turns a CEL error VALUE into nil (:1477-1482), which evaluateBoolCELExpr reads
        as false (:1431-1433).  Pinned by golden engine/case_20: any(<missing attribute>, true) is satisfied.)"""
        if cond is None:
            return True
        if cond.op == "expr":
            try:
                return self._eval(ctx, cond.expr, constants, variables) is True
            except CelError:
                return False
        if cond.op == "all":
            return all(self.satisfies(ctx, c, constants, variables) for c in cond.children)
        if cond.op == "any":
            return any(self.satisfies(ctx, c, constants, variables) for c in cond.children)
        if cond.op == "none":
            return not any(self.satisfies(ctx, c, constants, variables) for c in cond.children)
        raise ValueError(cond.op)

    # ------------------------------------------------------------- check
    def check(self, inp: dict, now: Timestamp | None = None, lenient=None) -> dict:
        """Returns {"actions": {action: {"effect","policy","scope"}}, "effectiveDerivedRoles": [...]}."""
        lenient_saved = self.lenient
        if lenient is not None:
            self.lenient = lenient
        try:
            return self._check(inp, now)
        finally:
            self.lenient = lenient_saved

    def _check(self, inp, now):
        p, r = inp.get("principal") or {}, inp.get("resource") or {}
        actions = list(inp.get("actions") or [])
        p_scope = namer.scope_value(p.get("scope") or self.default_scope)
        r_scope = namer.scope_value(r.get("scope") or self.default_scope)
        p_ver = p.get("policyVersion") or self.default_version
        r_ver = r.get("policyVersion") or self.default_version
        p_id, p_roles, kind = p.get("id", ""), list(p.get("roles") or []), r.get("kind", "")

        effects = {a: {"effect": EFFECT_DENY, "policy": NO_POLICY_MATCH, "scope": ""} for a in actions}
        out = {"actions": effects, "effectiveDerivedRoles": [], "outputs": []}

        p_scopes, p_key, _ = self.get_all_scopes(KIND_PRINCIPAL, p_scope, p_id, p_ver)
        r_scopes, r_key, _ = self.get_all_scopes(KIND_RESOURCE, r_scope, kind, r_ver)
        if not p_scopes and not r_scopes:
            return out
        if not actions:
            return out

        ctx = {"request": build_request(inp), "now": now, "edr": []}
        resource = namer.sanitize(kind)
        p_exists = self.scoped_principal_exists(p_ver, p_scopes)
        r_exists = self.scoped_resource_exists(r_ver, resource, r_scopes)
        if not p_exists and not r_exists:
            return out

        all_roles = self.add_parent_roles(r_scope, p_roles)
        including_parents = set(all_roles)
        combined = list(dict.fromkeys(p_scopes + r_scopes))
        cand = self.get_rows(r_ver, resource, combined, all_roles, actions)

        var_cache, cond_cache, processed_dr = {}, {}, set()
        all_edr = set()
        for action in actions:
            info = {"effect": EFFECT_NO_MATCH, "policy": "", "scope": ""}
            for pt in (KIND_PRINCIPAL, KIND_RESOURCE):
                main_key, scopes = (p_key, p_scopes) if pt == KIND_PRINCIPAL else (r_key, r_scopes)
                info["effect"] = EFFECT_NO_MATCH
                for i, role in enumerate(p_roles):
                    if i > 0 and pt == KIND_PRINCIPAL:
                        break
                    role_effects = set()
                    rinfo = {"effect": EFFECT_NO_MATCH, "policy": NO_POLICY_MATCH, "scope": ""}
                    if (pt == KIND_RESOURCE and r_exists) or (pt == KIND_PRINCIPAL and p_exists):
                        rinfo["policy"] = main_key
                    parent_roles = self.add_parent_roles(r_scope, [role])
                    broke = False
                    for scope in scopes:
                        if pt == KIND_RESOURCE and scope not in processed_dr:
                            edr = set()
                            drs = self.rt.policy_derived_roles.get(namer.resource_policy_fqn(kind, r_ver, scope))
                            for name, dr in (drs or {}).items():
                                if not (set(dr.parent_roles) & including_parents) and "*" not in dr.parent_roles:
                                    continue
                                ck = dr.origin_fqn + "#" + name
                                if ck not in var_cache:
                                    var_cache[ck] = self._eval_vars(ctx, dr.params)
                                c, v = var_cache[ck]
                                if self.satisfies(ctx, dr.condition, c, v):
                                    edr.add(name)
                                    all_edr.add(name)
                            ctx["edr"] = sorted(edr)
                            processed_dr.add(scope)
                        if rinfo["effect"] != EFFECT_NO_MATCH:
                            break
                        for row in cand:
                            if not self.row_matches(row, pt, scope, action, p_id, parent_roles):
                                continue
                            constants = variables = None
                            if row.params is not None:
                                if row.params.key not in var_cache:
                                    var_cache[row.params.key] = self._eval_vars(ctx, row.params)
                                constants, variables = var_cache[row.params.key]
                            if row.evaluation_key in cond_cache and row.evaluation_key != "":
                                sat = cond_cache[row.evaluation_key]
                            else:
                                sat = None
                                if row.dr_condition is not None:
                                    dc = dv = None
                                    if row.dr_params is not None:
                                        k = "DR:" + row.dr_params.key + "@" + row.origin_fqn
                                        if k not in var_cache:
                                            var_cache[k] = self._eval_vars(ctx, row.dr_params)
                                        dc, dv = var_cache[k]
                                    if not self.satisfies(ctx, row.dr_condition, dc, dv):
                                        cond_cache[row.evaluation_key] = False
                                        sat = False
                                if sat is None:
                                    sat = self.satisfies(ctx, row.condition, constants, variables)
                                    cond_cache[row.evaluation_key] = sat
                            # rule outputs (ruletable.go:1065-1080, 1096-1106): every visit of a row emits its entry -- the
                            # same rule visited for two roles emits twice
                            emit = row.emit_activated if sat else row.emit_not_met
                            if emit is not None:
                                out["outputs"].append({"action": action, "src": f"{namer.policy_key_from_fqn(row.origin_fqn)}#{row.name}",
                                                       "val": self._output_value(ctx, emit, constants, variables)})
                            if sat:
                                role_effects.add(row.effect)
                                if row.effect == EFFECT_DENY:
                                    rinfo["effect"] = EFFECT_DENY
                                    rinfo["scope"] = scope
                                    if row.from_role_policy:
                                        rinfo["policy"] = namer.policy_key_from_fqn(row.origin_fqn)
                                    broke = True
                                    break
                                elif row.no_match_for_scope_permissions:
                                    rinfo["policy"] = NO_MATCH_SCOPE_PERMISSIONS
                                    rinfo["scope"] = scope
                        if broke:
                            break
                        if EFFECT_ALLOW in role_effects:
                            sp = self.scope_perms.get(scope, SP_UNSPECIFIED)
                            if sp == SP_REQUIRE_PARENTAL_CONSENT:
                                role_effects.discard(EFFECT_ALLOW)
                            elif sp == SP_OVERRIDE_PARENT:
                                rinfo["effect"] = EFFECT_ALLOW
                                rinfo["scope"] = scope
                                break
                    if info["effect"] == EFFECT_NO_MATCH:
                        info = dict(rinfo)
                    if rinfo["effect"] == EFFECT_ALLOW:
                        info = dict(rinfo)
                        break
                    elif (rinfo["effect"] == EFFECT_DENY and info["policy"] == NO_MATCH_SCOPE_PERMISSIONS
                          and rinfo["policy"] != NO_MATCH_SCOPE_PERMISSIONS):
                        info = dict(rinfo)
                if info["effect"] in (EFFECT_ALLOW, EFFECT_DENY):
                    break
            if info["effect"] == EFFECT_NO_MATCH:
                info["effect"] = EFFECT_DENY
            effects[action] = info
        out["effectiveDerivedRoles"] = sorted(all_edr)
        return out
