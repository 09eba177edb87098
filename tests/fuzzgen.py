"""Seeded random policy sets + requests for differential tests (oracle #1 vs oracle #2 vs kernel core vs GPU)."""
import random

KINDS = ["doc", "album:object", "leave_request", "report"]
ROLES = ["user", "manager", "admin", "auditor", "guest"]
ACTIONS = ["view", "view:public", "edit", "delete", "approve", "share:*", "*"]
REQ_ACTIONS = ["view", "view:public", "view:private", "edit", "delete", "approve", "share:team", "share:all", "zzz"]
SCOPES = ["", "acme", "acme.hr", "acme.hr.uk", "beta"]
STRS = ["a", "b", "eng", "ops", "x1", "gold", "silver", ""]


def rand_expr(r: random.Random, depth=0):
    """A CEL expression from the device-supported subset (so that tables flatten)."""
    attrs_s = ["P.attr.dept", "R.attr.dept", "R.attr.owner", "P.attr.team", "R.attr.tier", "request.resource.attr.status"]
    attrs_n = ["P.attr.level", "R.attr.min_level", "R.attr.size"]
    attrs_l = ["P.attr.groups", "R.attr.allowed", "P.attr.tags"]
    attrs_b = ["R.attr.public", "P.attr.vip"]
    k = r.random()
    if depth < 2 and k < 0.25:
        op = r.choice(["&&", "||"])
        return f"({rand_expr(r, depth + 1)} {op} {rand_expr(r, depth + 1)})"
    if depth < 2 and k < 0.30:
        return f"!({rand_expr(r, depth + 1)})"
    if depth < 2 and k < 0.35:
        return f"({rand_expr(r, depth + 1)} ? {rand_expr(r, depth + 1)} : {rand_expr(r, depth + 1)})"
    c = r.randrange(16)
    if c == 0:
        return f"{r.choice(attrs_s)} == {r.choice(attrs_s + ['P.id', 'R.kind', 'R.id'])}"
    if c == 1:
        return f'{r.choice(attrs_s)} {r.choice(["==", "!="])} "{r.choice(STRS)}"'
    if c == 2:
        return f"{r.choice(attrs_n)} {r.choice(['<', '<=', '>', '>=', '==', '!='])} {r.choice(attrs_n + ['3', '5.5', '0'])}"
    if c == 3:
        return f'"{r.choice(STRS)}" in {r.choice(attrs_l)}'
    if c == 4:
        return f'{r.choice(attrs_s)} in ["{r.choice(STRS)}", "{r.choice(STRS)}"]'
    if c == 5:
        return f"{r.choice(attrs_b)} == {r.choice(['true', 'false'])}"
    if c == 6:
        return f"hasIntersection({r.choice(attrs_l)}, {r.choice(attrs_l)})"
    if c == 7:
        return f"{r.choice(attrs_l)}.exists(g, g == {r.choice(attrs_s)})"
    if c == 8:
        return f"size({r.choice(attrs_l + attrs_s)}) {r.choice(['>', '==', '<='])} {r.randrange(4)}"
    if c == 9:
        return f'{r.choice(attrs_s)}.{r.choice(["startsWith", "endsWith", "contains"])}("{r.choice(STRS)}")'
    if c == 10:
        return f"has({r.choice(attrs_s + attrs_l)})"
    if c == 11:
        return f"isSubset({r.choice(attrs_l)}, {r.choice(attrs_l)})"
    if c == 12:
        return f'{r.choice(attrs_l)}.all(g, g != "{r.choice(STRS)}")'
    if c == 13:
        return f"{r.choice(attrs_l)}[{r.randrange(3)}] == {r.choice(attrs_s)}"
    if c == 14:
        return f"{r.choice(attrs_n)} + {r.choice(attrs_n)} > {r.randrange(10)}.0"
    return f"{r.choice(attrs_s)} in {r.choice(attrs_l)}"


def rand_cond(r: random.Random, depth=0):
    k = r.random()
    if depth < 2 and k < 0.2:
        op = r.choice(["all", "any", "none"])
        return {op: {"of": [rand_cond(r, depth + 1) for _ in range(r.randrange(1, 4))]}}
    return {"expr": rand_expr(r)}


def rand_policies(r: random.Random):
    docs = []
    dr_defs = []
    for i in range(r.randrange(0, 4)):
        d = {"name": f"dr{i}", "parentRoles": r.sample(ROLES, r.randrange(1, 3))}
        if r.random() < 0.8:
            d["condition"] = {"match": rand_cond(r)}
        dr_defs.append(d)
    if dr_defs:
        docs.append({"apiVersion": "api.cerbos.dev/v1", "derivedRoles": {"name": "drs", "definitions": dr_defs}})
    for kind in r.sample(KINDS, r.randrange(1, len(KINDS) + 1)):
        # a scoped policy needs all its ancestors
        scopes = {""}
        for s in r.sample(SCOPES, r.randrange(0, 3)):
            parts = s.split(".") if s else []
            for j in range(len(parts) + 1):
                scopes.add(".".join(parts[:j]))
        for s in sorted(scopes):
            rules = []
            for _ in range(r.randrange(1, 7)):
                rule = {"actions": r.sample(ACTIONS, r.randrange(1, 4)), "effect": r.choice(["EFFECT_ALLOW", "EFFECT_ALLOW", "EFFECT_DENY"])}
                if dr_defs and r.random() < 0.3:
                    rule["derivedRoles"] = [r.choice(dr_defs)["name"]]
                if "derivedRoles" not in rule or r.random() < 0.3:
                    rule["roles"] = r.sample(ROLES + ["*"], r.randrange(1, 3))
                if r.random() < 0.45:
                    rule["condition"] = {"match": rand_cond(r)}
                rules.append(rule)
            rp = {"resource": kind, "version": r.choice(["default", "default", "v2"]), "rules": rules}
            if r.random() < 0.4:
                # variables and constants (local and imported; a variable may use earlier ones, constants, or fail): the
                # table builder inlines them, oracle #1 evaluates them eagerly per request (ruletable.go:1319-1344)
                rp["constants"] = {"local": {"c0": r.choice(STRS), "c1": r.choice([1, 3, 5.5]), "c2": [r.choice(STRS), r.choice(STRS)]}}
                rp["variables"] = {"local": {
                    "flag": rand_expr(r, 1), "lvl": "P.attr.level", "who": r.choice(["P.id", "R.attr.owner", "P.attr.nope"]),
                    "both": f"V.flag || ({rand_expr(r, 1)})", "big": "variables.lvl > C.c1", "inc": "C.c0 in constants.c2 || V.who == R.attr.owner"}}
                if r.random() < 0.5:
                    rp["variables"]["import"] = ["shared_vars"]
                    rp["constants"]["import"] = ["shared_consts"]
                extra = ["V.flag", "V.both", "V.big", "V.inc", "V.who == P.id", "V.lvl >= 3", "R.attr.dept == C.c0", "V.who in C.c2", "!V.flag"]
                if "import" in rp["variables"]:
                    extra += ["V.sv_owner", "V.sv_lvl > C.sk1", "C.sk0 == P.attr.dept"]
                for rule in rules:
                    if r.random() < 0.6:
                        x = r.choice(extra)
                        if "condition" in rule and r.random() < 0.5:
                            rule["condition"] = {"match": {r.choice(["all", "any"]): {"of": [rule["condition"]["match"], {"expr": x}]}}}
                        else:
                            rule["condition"] = {"match": {"expr": x}}
            if dr_defs and r.random() < 0.3 and scopes == {""}:
                # runtime.effectiveDerivedRoles in a condition (ruletable.go:936-979) -- only where the kind has one scope: along
                # a scope chain the reference's value depends on the order in which actions and roles were walked (the set is
                # replaced when a scope is FIRST processed and kept afterwards), which the device does not reproduce
                rules.append({"actions": [r.choice(ACTIONS)], "effect": r.choice(["EFFECT_ALLOW", "EFFECT_DENY"]), "roles": ["*"],
                              "condition": {"match": {"expr": f'"{r.choice(dr_defs)["name"]}" in runtime.effectiveDerivedRoles'}}})
            if dr_defs:
                rp["importDerivedRoles"] = ["drs"]
            if s:
                rp["scope"] = s
            if r.random() < 0.3:
                rp["scopePermissions"] = "SCOPE_PERMISSIONS_REQUIRE_PARENTAL_CONSENT_FOR_ALLOWS"
            docs.append({"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": rp})
    # make every (kind, version) ancestor-complete: drop scoped policies whose ancestors miss for that version
    have = {(d["resourcePolicy"]["resource"], d["resourcePolicy"]["version"], d["resourcePolicy"].get("scope", "")) for d in docs if "resourcePolicy" in d}
    keep = []
    for d in docs:
        if "resourcePolicy" in d:
            rp = d["resourcePolicy"]
            s = rp.get("scope", "")
            parts = s.split(".") if s else []
            if not all((rp["resource"], rp["version"], ".".join(parts[:j])) in have for j in range(len(parts))):
                continue
        keep.append(d)
    docs = keep
    if any("import" in (d.get("resourcePolicy", {}).get("variables") or {}) for d in docs):
        docs.append({"apiVersion": "api.cerbos.dev/v1", "exportVariables": {"name": "shared_vars", "definitions": {
            "sv_owner": "R.attr.owner == P.id", "sv_lvl": "P.attr.level"}}})
        docs.append({"apiVersion": "api.cerbos.dev/v1", "exportConstants": {"name": "shared_consts", "definitions": {"sk0": r.choice(STRS), "sk1": 2}}})
    # de-duplicate (kind, version, scope)
    seen, out = set(), []
    for d in docs:
        if "resourcePolicy" in d:
            k = (d["resourcePolicy"]["resource"], d["resourcePolicy"]["version"], d["resourcePolicy"].get("scope", ""))
            if k in seen:
                continue
            seen.add(k)
        out.append(d)
    docs = out
    if r.random() < 0.4:   # role policies with parent roles
        for i in range(r.randrange(1, 3)):
            docs.append({"apiVersion": "api.cerbos.dev/v1", "rolePolicy": {
                "role": f"custom{i}", "scope": r.choice(SCOPES), "parentRoles": r.sample(ROLES + [f"custom{(i + 1) % 2}"], r.randrange(0, 3)),
                "rules": [{"resource": r.choice(KINDS + ["*"]), "allowActions": r.sample(ACTIONS, r.randrange(1, 3)),
                           **({"condition": {"match": rand_cond(r)}} if r.random() < 0.4 else {})} for _ in range(r.randrange(1, 3))]}})
    if r.random() < 0.4:   # principal policies
        for pid in r.sample(["alice", "bob"], r.randrange(1, 3)):
            docs.append({"apiVersion": "api.cerbos.dev/v1", "principalPolicy": {
                "principal": pid, "version": "default",
                "rules": [{"resource": r.choice(KINDS + ["*"]), "actions": [
                    {"action": r.choice(ACTIONS), "effect": r.choice(["EFFECT_ALLOW", "EFFECT_DENY"]),
                     **({"condition": {"match": rand_cond(r)}} if r.random() < 0.5 else {})} for _ in range(r.randrange(1, 3))]}
                    for _ in range(r.randrange(1, 3))]}})
    return docs


def rand_value(r: random.Random, kind):
    if kind == "s":
        return r.choice(STRS + ["alice", "bob", "doc"])
    if kind == "n":
        return r.choice([0, 1, 3, 5, 5.5, 9, -2])
    if kind == "l":
        return [r.choice(STRS) for _ in range(r.randrange(0, 5))]
    return r.random() < 0.5


def rand_request(r: random.Random):
    def attrs(spec):
        out = {}
        for name, kind in spec:
            if r.random() < 0.8:
                out[name] = rand_value(r, kind if r.random() < 0.9 else r.choice("snlb"))
        return out
    p = {"id": r.choice(["alice", "bob", "carol", "a"]), "roles": r.sample(ROLES + ["custom0", "custom1", "ghost"], r.randrange(1, 5)),
         "attr": attrs([("dept", "s"), ("team", "s"), ("level", "n"), ("groups", "l"), ("tags", "l"), ("vip", "b")])}
    res = {"kind": r.choice(KINDS + ["unknown"]), "id": r.choice(["d1", "a", "x1"]),
           "attr": attrs([("dept", "s"), ("owner", "s"), ("tier", "s"), ("status", "s"), ("min_level", "n"), ("size", "n"), ("allowed", "l"), ("public", "b")])}
    if r.random() < 0.6:
        res["scope"] = r.choice(SCOPES + ["acme.hr.uk.london", "nope"])
    if r.random() < 0.2:
        p["scope"] = r.choice(SCOPES)
    if r.random() < 0.2:
        res["policyVersion"] = r.choice(["default", "v2", "v9"])
    if r.random() < 0.1:
        p["policyVersion"] = r.choice(["default", "v2"])
    return {"requestId": "f", "actions": r.sample(REQ_ACTIONS, r.randrange(1, 6)), "principal": p, "resource": res}
