"""Shared helpers for the test-suite (golden loading, oracle construction)."""
import functools
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EFFECT_NAMES = {1: "EFFECT_ALLOW", 2: "EFFECT_DENY", 3: "EFFECT_NO_MATCH"}


def load_golden(name):
    with open(os.path.join(GOLDEN, name), encoding="utf-8") as f:
        return json.load(f)


@functools.lru_cache(maxsize=None)
def store_rule_table():
    from cerbos_b200.policy.compile import build_rule_table
    docs = [e["policy"] for e in load_golden("store_policies.json")]
    return build_rule_table(docs)


def engine_decisions():
    """Yields (case_id, lenient, input, action, want_effect_dict) for every golden decision."""
    for case in load_golden("engine_cases.json"):
        outs = {(o.get("requestId"), o.get("resourceId")): o for o in case["wantOutputs"]}
        for inp in case["inputs"]:
            want = outs[(inp.get("requestId"), inp["resource"]["id"])]
            yield f"{case['suite']}/{case['file']}", case["lenient"], inp, want


def check_resources_api_cases():
    """The reference's API-shaped CheckResources goldens (internal/test/testdata/server/checks/check_resources/cr_case_0*.yaml,
    extracted by tests/golden/make_golden.py): one CheckInput per resource entry, the way the service builds them
    (internal/svc/cerbos_svc.go:249-263), with the decoded JWT claims as auxData.  Yields (file, CheckInput, {action: effect name}).
    Skipped: request-validation error cases (rejected before the engine) and cr_case_02 (schema enforcement, out of scope)."""
    for c in load_golden("check_resources_cases.json"):
        if c.get("wantError") or c["file"] == "cr_case_02.yaml":
            continue
        inp = c["input"]
        for i, entry in enumerate(inp["resources"]):
            ci = {"requestId": inp.get("requestId", ""), "actions": entry["actions"], "principal": inp["principal"], "resource": entry["resource"]}
            if c.get("jwtClaims"):
                ci["auxData"] = {"jwt": c["jwtClaims"]}
            want = {a: (w if isinstance(w, str) else w.get("effect")) for a, w in c["wantResponse"]["results"][i]["actions"].items()}
            yield c["file"], ci, want


def check_resources_api_outputs():
    """The rule outputs the same goldens record (results[i].outputs): yields (file, CheckInput, [OutputEntry, ...])."""
    for c in load_golden("check_resources_cases.json"):
        if c.get("wantError") or c["file"] == "cr_case_02.yaml":
            continue
        inp = c["input"]
        for i, entry in enumerate(inp["resources"]):
            ci = {"requestId": inp.get("requestId", ""), "actions": entry["actions"], "principal": inp["principal"], "resource": entry["resource"]}
            if c.get("jwtClaims"):
                ci["auxData"] = {"jwt": c["jwtClaims"]}
            yield c["file"], ci, c["wantResponse"]["results"][i].get("outputs") or []


def verify_suite_cases():
    """Engine answers recorded by the reference's policy-test goldens (internal/test/testdata/verify/cases/*.golden, extracted by
    tests/golden/make_golden.py::verify_cases): grouped by engine configuration.
    Yields ((globals json, default version, default scope, lenient), [case, ...])."""
    import json
    groups = {}
    for c in load_golden("verify_cases.json"):
        key = (json.dumps(c["globals"] or {}, sort_keys=True), str(c["defaultPolicyVersion"] or "default"), c["defaultScope"] or "", bool(c["lenient"]))
        groups.setdefault(key, []).append(c)
    yield from groups.items()
