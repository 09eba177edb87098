"""Policy documents -> rule-table rows (host side, once per policy change).

Stands in for two Go stages that stay on the host in the reference:

* ``compile.Compile`` (internal/compile/compile.go:54-235, 237-357, 379-549;
  variables.go:29-313; constants.go) -- import derived roles / export_variables /
  export_constants, keep only the *used* variables and constants (transitively,
  dependency-ordered), parse condition CEL;
* ``ruletable.AddPolicy`` (internal/ruletable/ruletable.go:105-416) -- cartesian
  (action x role) row generation, derived roles expanded to one row per parent
  role with ``DerivedRoleCondition``, REQUIRE_PARENTAL_CONSENT conditional ALLOWs
  rewritten to ``DENY none(cond)``, no-op rows for empty policies, role-policy
  allow-action rows and parent-role maps.

Input: policy documents as plain dicts (YAML/JSON already parsed, protojson field
names as in api/public/cerbos/policy/v1/policy.proto:35-318).
"""
from __future__ import annotations

from ..cel import parser as celparser
from ..cel.ast import Call, Ident, Select, walk
from . import namer
from .model import (Cond, DerivedRole, EFFECT_ALLOW, EFFECT_DENY, Expr, KIND_PRINCIPAL, KIND_RESOURCE, Params, Row,
                    RuleTable, SP_OVERRIDE_PARENT, SP_REQUIRE_PARENTAL_CONSENT, SP_UNSPECIFIED, Variable,
                    parse_effect, parse_scope_permissions)

ANY_ROLE = "*"


class PolicyCompileError(ValueError):
    pass


# ------------------------------------------------------------------ conditions

_DEFS_IDENTS = ("V", "variables", "C", "constants", "G", "globals")
_CEL_KEYWORDS = ("false", "in", "null", "true")


def validate_identifier(kind: str, name) -> None:
    """conditions.ValidateIdentifier (internal/conditions/identifiers.go:24-34): names of constants and variables."""
    import re
    name = str(name)
    if name in _CEL_KEYWORDS:
        raise PolicyCompileError(f"invalid {kind} name: \"{name}\" is a reserved keyword and can't be used as an identifier")
    if not re.match(r"^[_a-zA-Z][_a-zA-Z0-9]*$", name):
        raise PolicyCompileError(f"invalid {kind} name: \"{name}\" is not a valid identifier")


def compile_expr(src: str) -> Expr:
    try:
        e = Expr(original=src, ast=celparser.parse(src))
    except celparser.CelSyntaxError as ex:
        raise PolicyCompileError(f"invalid expression `{src.strip()}`: {ex}") from ex
    # variables / constants / globals are record-like (cerbos.Variables): field selection only -- V["x"] has no overload
    # and the reference's type checker rejects it (testdata/compile/variables_index_lookup.yaml)
    for n in walk(e.ast):
        if isinstance(n, Call) and n.fn == "_[_]" and n.target is None and len(n.args) == 2 and isinstance(n.args[0], Ident) and n.args[0].name in _DEFS_IDENTS:
            raise PolicyCompileError(f"invalid expression `{src.strip()}`: found no matching overload for '_[_]' applied to '(cerbos.Variables, ...)'")
    return e


def _compile_output(out, defs):
    """policy.proto Output -> (when.rule_activated, when.condition_not_met) (compile.go:416-432, 520-536: the deprecated
    `expr` is the rule_activated expression unless `when.ruleActivated` is given too)"""
    out = out or {}
    when = out.get("when") or {}
    activated = when.get("ruleActivated") or out.get("expr") or None
    not_met = when.get("conditionNotMet") or None
    if out.get("expr"):
        defs.use_expr(compile_expr(out["expr"]))
    res = []
    for src in (activated, not_met):
        e = compile_expr(src) if src else None
        if e is not None:
            defs.use_expr(e)
        res.append(e)
    return tuple(res)


def compile_match(m: dict) -> Cond:
    """policy.proto Match: {expr} | {all|any|none: {of: [...]}} (policy.proto:295-318)."""
    if m is None:
        return None
    if "expr" in m:
        return Cond("expr", expr=compile_expr(m["expr"]))
    for op in ("all", "any", "none"):
        if op in m:
            of = (m[op] or {}).get("of") or []
            return Cond(op, children=[compile_match(x) for x in of])
    raise PolicyCompileError(f"invalid match block: {m!r}")


def compile_condition(c: dict):
    if c is None:
        return None
    if "script" in c:
        raise PolicyCompileError("scripts in conditions are no longer supported")  # compile/conditions.go:33-35
    if "match" not in c:
        raise PolicyCompileError(f"invalid condition: {c!r}")
    return compile_match(c["match"])


def cond_exprs(c: Cond):
    if c is None:
        return
    if c.op == "expr":
        yield c.expr
    else:
        for ch in c.children:
            yield from cond_exprs(ch)


def expr_references(e: Expr):
    """(constants, variables) referenced as C.x / constants.x / V.x / variables.x select nodes
    (compile/variables.go:216-245)."""
    consts, vars_ = set(), set()
    for n in walk(e.ast):
        if isinstance(n, Select) and isinstance(n.operand, Ident):
            if n.operand.name in ("C", "constants"):
                consts.add(n.field)
            elif n.operand.name in ("V", "variables"):
                vars_.add(n.field)
    return consts, vars_


# ------------------------------------------------------------------ variable / constant scopes

class _Defs:
    """Variable + constant definitions visible to one policy module, with usage tracking."""

    def __init__(self, where: str):
        self.where = where
        self.var_defs: dict[str, Expr] = {}
        self.const_defs: dict[str, object] = {}
        self.used_vars: set[str] = set()
        self.used_consts: set[str] = set()

    def add_vars(self, defs: dict, source: str):
        for name, src in (defs or {}).items():
            validate_identifier("variable", name)
            if name in self.var_defs:
                raise PolicyCompileError(f"{self.where}: variable '{name}' has multiple definitions ({source})")
            self.var_defs[name] = compile_expr(src)

    def add_consts(self, defs: dict, source: str):
        for name, val in (defs or {}).items():
            validate_identifier("constant", name)
            if name in self.const_defs:
                raise PolicyCompileError(f"{self.where}: constant '{name}' has multiple definitions ({source})")
            self.const_defs[name] = val

    def reset_usage(self):
        self.used_vars = set()
        self.used_consts = set()

    def use_expr(self, e: Expr):
        consts, vars_ = expr_references(e)
        for c in consts:
            if c not in self.const_defs:
                raise PolicyCompileError(f"{self.where}: undefined constant '{c}'")
            self.used_consts.add(c)
        for v in vars_:
            self._use_var(v)

    def _use_var(self, name: str, stack=()):
        if name not in self.var_defs:
            raise PolicyCompileError(f"{self.where}: undefined variable '{name}'")
        if name in stack:
            raise PolicyCompileError(f"{self.where}: variables {list(stack) + [name]} form a cycle")
        if name in self.used_vars:
            return
        self.used_vars.add(name)
        consts, vars_ = expr_references(self.var_defs[name])
        for c in consts:
            if c not in self.const_defs:
                raise PolicyCompileError(f"{self.where}: undefined constant '{c}'")
            self.used_consts.add(c)
        for v in vars_:
            self._use_var(v, stack + (name,))

    def use_cond(self, c: Cond):
        for e in cond_exprs(c):
            self.use_expr(e)

    def ordered_used_vars(self) -> list:
        """Dependency order, name-sorted ties (compile/variables.go:251-313)."""
        out, done = [], set()

        def visit(name, stack=()):
            if name in done:
                return
            if name in stack:
                raise PolicyCompileError(f"{self.where}: variable cycle at '{name}'")
            _, deps = expr_references(self.var_defs[name])
            for d in sorted(deps):
                visit(d, stack + (name,))
            done.add(name)
            out.append(Variable(name, self.var_defs[name]))

        for name in sorted(self.used_vars):
            visit(name)
        return out

    def params(self, key: str) -> Params:
        return Params(key=key, variables=self.ordered_used_vars(),
                      constants={k: self.const_defs[k] for k in sorted(self.used_consts)})


# ------------------------------------------------------------------ policy set

def _kind_of(doc: dict) -> str:
    for k in ("resourcePolicy", "principalPolicy", "rolePolicy", "derivedRoles", "exportVariables", "exportConstants"):
        if k in doc:
            return k
    raise PolicyCompileError(f"unknown policy type: {list(doc)}")


class PolicySet:
    """All (enabled) policy documents of a store, indexed for import resolution."""

    def __init__(self, docs):
        self.resource, self.principal, self.role = [], [], []
        self.derived_roles, self.export_vars, self.export_consts = {}, {}, {}
        self._rp_keys, self._pp_keys = set(), set()
        for doc in docs:
            if doc.get("disabled"):
                continue
            kind = _kind_of(doc)
            body = doc[kind]
            if kind == "resourcePolicy":
                self.resource.append(doc)
                self._rp_keys.add((namer.sanitize(body["resource"]), body.get("version", ""),
                                   namer.scope_value(body.get("scope", "") or "")))
            elif kind == "principalPolicy":
                self.principal.append(doc)
                self._pp_keys.add((namer.sanitize(body["principal"]), body.get("version", ""),
                                   namer.scope_value(body.get("scope", "") or "")))
            elif kind == "rolePolicy":
                self.role.append(doc)
            elif kind == "derivedRoles":
                self.derived_roles[body["name"]] = doc
            elif kind == "exportVariables":
                self.export_vars[body["name"]] = doc
            else:
                self.export_consts[body["name"]] = doc

    # -- definitions scopes
    def _defs_for(self, doc: dict, body: dict, where: str) -> _Defs:
        d = _Defs(where)
        consts = body.get("constants") or {}
        for imp in consts.get("import") or []:
            ec = self.export_consts.get(imp)
            if ec is None:
                raise PolicyCompileError(f"{where}: constants import '{imp}' cannot be found")
            d.add_consts(ec["exportConstants"].get("definitions"), f"import '{imp}'")
        d.add_consts(consts.get("local"), "policy local constants")
        vars_ = body.get("variables") or {}
        for imp in vars_.get("import") or []:
            ev = self.export_vars.get(imp)
            if ev is None:
                raise PolicyCompileError(f"{where}: variables import '{imp}' cannot be found")
            d.add_vars(ev["exportVariables"].get("definitions"), f"import '{imp}'")
        d.add_vars(vars_.get("local"), "policy local variables")
        d.add_vars(doc.get("variables"), "deprecated top-level policy variables")
        return d

    def _check_ancestors(self, keys, name, version, scope, where):
        for anc in namer.scope_parents(scope):
            if (name, version, anc) not in keys:
                raise PolicyCompileError(f"{where}: missing ancestor policy at scope '{anc}'")

    # -- derived roles
    def _compile_derived_roles(self, set_name: str, where: str) -> dict:
        doc = self.derived_roles.get(set_name)
        if doc is None:
            raise PolicyCompileError(f"{where}: derived roles import '{set_name}' cannot be found")
        body = doc["derivedRoles"]
        defs = self._defs_for(doc, body, f"derived_roles.{set_name}")
        origin = namer.derived_roles_fqn(body["name"])
        out = {}
        for d in body.get("definitions") or []:
            parents = []
            for pr in d.get("parentRoles") or []:
                if pr == ANY_ROLE:
                    parents = [ANY_ROLE]
                    break
                if pr not in parents:
                    parents.append(pr)
            defs.reset_usage()
            cond = compile_condition(d.get("condition"))
            defs.use_cond(cond)
            out[d["name"]] = DerivedRole(name=d["name"], parent_roles=parents, condition=cond,
                                         params=defs.params(namer.derived_roles_fqn(d["name"])), origin_fqn=origin)
        return out

    # -- rows
    def build_rule_table(self) -> RuleTable:
        rt = RuleTable()
        for doc in self.resource:
            rt.rows.extend(self._resource_rows(rt, doc))
        for doc in self.principal:
            rt.rows.extend(self._principal_rows(doc))
        for doc in self.role:
            rt.rows.extend(self._role_rows(rt, doc))
        return rt

    def _resource_rows(self, rt: RuleTable, doc: dict) -> list:
        rp = doc["resourcePolicy"]
        resource = namer.sanitize(rp["resource"])
        version = rp.get("version", "")
        scope = namer.scope_value(rp.get("scope", "") or "")
        fqn = namer.resource_policy_fqn(rp["resource"], version, scope)
        where = namer.policy_key_from_fqn(fqn)
        self._check_ancestors(self._rp_keys, resource, version, scope, where)

        # imported derived roles, restricted to the ones some rule references (compile.go:237-317)
        imports: dict[str, list] = {}
        for imp in rp.get("importDerivedRoles") or []:
            for name, dr in self._compile_derived_roles(imp, where).items():
                imports.setdefault(name, []).append(dr)
        referenced: dict[str, DerivedRole] = {}
        for rule in rp.get("rules") or []:
            for r in rule.get("derivedRoles") or []:
                if r not in imports:
                    raise PolicyCompileError(f"{where}: derived role '{r}' is not defined in any imports")
                if len(imports[r]) > 1:
                    raise PolicyCompileError(f"{where}: derived role '{r}' is defined in more than one import")
                referenced[r] = imports[r][0]
        if referenced:
            rt.policy_derived_roles[fqn] = referenced

        defs = self._defs_for(doc, rp, where)
        sp_raw = parse_scope_permissions(rp.get("scopePermissions"))
        sp = sp_raw if sp_raw != SP_UNSPECIFIED else SP_OVERRIDE_PARENT

        rules = []
        for i, rule in enumerate(rp.get("rules") or []):
            if not rule.get("roles") and not rule.get("derivedRoles"):
                raise PolicyCompileError(f"{where}: rule #{i + 1} does not specify any roles or derived roles")
            name = rule.get("name") or f"rule-{i + 1:03d}"
            cond = compile_condition(rule.get("condition"))
            defs.use_cond(cond)
            emit = _compile_output(rule.get("output"), defs)
            roles = []
            for r in rule.get("roles") or []:
                if r == ANY_ROLE:
                    roles = [ANY_ROLE]
                    break
                if r not in roles:
                    roles.append(r)
            actions = list(dict.fromkeys(rule.get("actions") or []))
            drs = list(dict.fromkeys(rule.get("derivedRoles") or []))
            rules.append((name, cond, parse_effect(rule["effect"]), roles, actions, drs, emit))
        params = defs.params(fqn)

        rows = []
        if not rules:
            rows.append(Row(origin_fqn=fqn, resource=resource, scope=scope, scope_permissions=sp, version=version,
                            policy_kind=KIND_RESOURCE, params=Params(fqn, [], {}), dr_params=Params("", [], {})))
        for name, cond, effect, roles, actions, drs, emit in rules:
            rule_fqn = f"{where}#{name}"
            eval_key = f"{fqn}#{rule_fqn}"
            for a in actions:
                for r in roles:
                    rows.append(self._consent_rewrite(Row(
                        origin_fqn=fqn, resource=resource, role=r, action=a, condition=cond, effect=effect,
                        scope=scope, scope_permissions=sp, version=version, name=name, params=params,
                        evaluation_key=eval_key, policy_kind=KIND_RESOURCE, emit_activated=emit[0], emit_not_met=emit[1]), sp_raw))
                for dr in drs:
                    rdr = referenced.get(dr)
                    if rdr is None:
                        continue
                    dr_key = f"{namer.derived_roles_fqn(dr)}#{rule_fqn}"
                    for pr in rdr.parent_roles:
                        rows.append(self._consent_rewrite(Row(
                            origin_fqn=fqn, resource=resource, role=pr, action=a, condition=cond,
                            dr_condition=rdr.condition, effect=effect, scope=scope, scope_permissions=sp,
                            version=version, origin_derived_role=dr, name=name, params=params,
                            dr_params=rdr.params, evaluation_key=dr_key, policy_kind=KIND_RESOURCE,
                            emit_activated=emit[0], emit_not_met=emit[1]), sp_raw))
        return rows

    @staticmethod
    def _consent_rewrite(row: Row, sp_raw: int) -> Row:
        # ruletable.go:198-209, 303-314, 352-363
        if sp_raw == SP_REQUIRE_PARENTAL_CONSENT and row.effect == EFFECT_ALLOW and row.condition is not None:
            row.condition = Cond("none", children=[row.condition])
            row.effect = EFFECT_DENY
        return row

    def _principal_rows(self, doc: dict) -> list:
        pp = doc["principalPolicy"]
        principal = pp["principal"]
        version = pp.get("version", "")
        scope = namer.scope_value(pp.get("scope", "") or "")
        fqn = namer.principal_policy_fqn(principal, version, scope)
        where = namer.policy_key_from_fqn(fqn)
        self._check_ancestors(self._pp_keys, namer.sanitize(principal), version, scope, where)
        defs = self._defs_for(doc, pp, where)
        sp_raw = parse_scope_permissions(pp.get("scopePermissions"))
        sp = sp_raw if sp_raw != SP_UNSPECIFIED else SP_OVERRIDE_PARENT

        resource_rules: dict[str, list] = {}
        for rule in pp.get("rules") or []:
            ars = []
            for i, act in enumerate(rule.get("actions") or []):
                name = act.get("name") or f"{rule['resource']}_rule-{i + 1:03d}"
                cond = compile_condition(act.get("condition"))
                defs.use_cond(cond)
                ars.append((act["action"], name, parse_effect(act["effect"]), cond, _compile_output(act.get("output"), defs)))
            resource_rules[rule["resource"]] = ars  # map keyed by resource: later block overwrites (compile.go:505-540)
        params = defs.params(fqn)

        rows = []
        if not resource_rules:
            rows.append(Row(origin_fqn=fqn, scope=scope, scope_permissions=sp, version=version, principal=principal,
                            policy_kind=KIND_PRINCIPAL, params=Params(fqn, [], {}), dr_params=Params("", [], {})))
        for resource, ars in resource_rules.items():
            for action, name, effect, cond, emit in ars:
                rule_fqn = f"{where}#{name}"
                rows.append(self._consent_rewrite(Row(
                    origin_fqn=fqn, resource=namer.sanitize(resource), role=ANY_ROLE, action=action, condition=cond,
                    effect=effect, scope=scope, scope_permissions=sp, version=version, name=name,
                    principal=principal, params=params, evaluation_key=f"{fqn}#{rule_fqn}",
                    policy_kind=KIND_PRINCIPAL, emit_activated=emit[0], emit_not_met=emit[1]), sp_raw))
        return rows

    def _role_rows(self, rt: RuleTable, doc: dict) -> list:
        rp = doc["rolePolicy"]
        role = rp["role"]
        version = rp.get("version") or namer.DEFAULT_VERSION  # compile.go:84-87
        scope = namer.scope_value(rp.get("scope", "") or "")
        fqn = namer.role_policy_fqn(role, version, scope)
        rows = []
        per_resource: dict[str, int] = {}
        for rule in rp.get("rules") or []:
            resource = rule["resource"]
            idx = per_resource.get(resource, 0)
            per_resource[resource] = idx + 1
            cond = compile_condition(rule.get("condition"))
            rows.append(Row(origin_fqn=fqn, role=role, resource=resource,
                            allow_actions=list(dict.fromkeys(rule.get("allowActions") or [])), condition=cond,
                            scope=scope, version=version,
                            evaluation_key=f"{namer.policy_key_from_fqn(fqn)}#{role}_rule-{idx:03d}",   # ruletable.go:398 (index within the resource's rule list)
                            policy_kind=KIND_RESOURCE, from_role_policy=True))
        rt.scope_parent_roles.setdefault(scope, {})[role] = list(rp.get("parentRoles") or [])
        return rows


def build_rule_table(docs) -> RuleTable:
    """docs: iterable of policy documents (dicts)."""
    return PolicySet(list(docs)).build_rule_table()
