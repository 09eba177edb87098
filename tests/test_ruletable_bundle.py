"""Rule-table bundle ingestion (cerbos_b200/table/ruletable_pb.py): the reference compiler's own serialized
runtimev1.RuleTable for the `store` policies (tests/golden/ruletable_bundle_unencrypted.crrt, copied by
tests/golden/make_golden.py from internal/test/testdata/bundle/v2_ruletable) is decoded, flattened and must answer the
reference's engine goldens -- effect, policy, scope, effectiveDerivedRoles -- exactly; and its rows must be the rows
cerbos_b200/policy/compile.py generates from the policy documents (pins SURVEY 8(a) row a9 on the reference's output)."""
import os

import pytest

from cerbos_b200 import meta as M
from cerbos_b200.encode import Encoder
from cerbos_b200.table import layout as L
from cerbos_b200.table.flatten import flatten
from cerbos_b200.table.ruletable_pb import WireError, decode_rule_table, fields
from conftest import GOLDEN
from helpers import engine_decisions, store_rule_table
from hostsim import driver as hostsim
from oracle.celeval import parse_timestamp

NOW = parse_timestamp("2024-01-01T00:00:00Z")
G = {"environment": "test"}


@pytest.fixture(scope="module")
def bundle():
    with open(os.path.join(GOLDEN, "ruletable_bundle_unencrypted.crrt"), "rb") as f:
        return f.read()


def _row_key(r):
    return (r.origin_fqn, r.resource, r.role, r.action, tuple(sorted(r.allow_actions or [])), r.effect, r.scope, r.scope_permissions, r.version,
            r.origin_derived_role, r.principal, r.policy_kind, r.from_role_policy, r.evaluation_key, r.name,
            r.condition is not None, r.dr_condition is not None)


def test_bundle_rows_equal_the_rows_of_our_policy_compiler(bundle):
    rt = decode_rule_table(bundle)
    ours = store_rule_table()
    assert len(rt.rows) == len(ours.rows) == 126
    assert sorted(map(_row_key, rt.rows), key=repr) == sorted(map(_row_key, ours.rows), key=repr)
    assert {s: {r: sorted(p) for r, p in m.items() if p} for s, m in rt.scope_parent_roles.items() if any(m.values())} == \
           {s: {r: sorted(p) for r, p in m.items() if p} for s, m in ours.scope_parent_roles.items() if any(m.values())}
    assert {f: sorted(d) for f, d in rt.policy_derived_roles.items() if d} == {f: sorted(d) for f, d in ours.policy_derived_roles.items() if d}


def test_engine_goldens_through_the_bundle_built_table(bundle):
    ft = flatten(decode_rule_table(bundle), globals_=G)
    names = {"EFFECT_ALLOW": 1, "EFFECT_DENY": 2}
    n = 0
    for cid, lenient, inp, want in engine_decisions():
        b = Encoder(ft.manifest, lenient_scope_search=lenient).encode([inp])
        fl = L.BATCH_FLAG_LENIENT if lenient else 0
        eff, am, rm = hostsim.check_meta(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, fl)
        k_out = hostsim.check(ft.blob, b.columns, b.n, b.max_actions, NOW.ns, fl)
        p, r = inp.get("principal") or {}, inp.get("resource") or {}
        for k, a in enumerate(inp["actions"]):
            pol, sc = M.decode_action(int(am[0, k]), rm[0], ft.manifest, p.get("id", ""), r.get("kind", ""),
                                      p.get("policyVersion") or "default", r.get("policyVersion") or "default")
            wa = want["actions"][a]
            assert (names[wa["effect"]], wa.get("policy", ""), wa.get("scope", "")) == (int(eff[0, k]), pol, sc), (cid, a)
            assert k_out[0, k] == eff[0, k], (cid, a)
            n += 1
        wedr = sorted(want.get("effectiveDerivedRoles", want.get("effective_derived_roles")) or [])
        assert M.decode_edr(int(rm[0]["effective_derived_roles"]), ft.manifest) == wedr, cid
    assert n == 166


def test_wire_reader_rejects_truncated_input(bundle):
    with pytest.raises(WireError):
        decode_rule_table(bundle[: len(bundle) // 2 + 1])
    assert list(fields(b"")) == []


@pytest.mark.gpu
def test_engine_from_bundle_on_gpu(bundle):
    """Engine.from_rule_table_bundle: the same goldens on the device."""
    from cerbos_b200.engine import Engine
    eng = Engine.from_rule_table_bundle(bundle, globals_=G)
    n = 0
    for cid, lenient, inp, want in engine_decisions():
        if lenient:
            continue
        got = eng.check([inp], now_ns=NOW.ns, include_meta=True)[0]
        for a, wv in want["actions"].items():
            g = got["actions"][a]
            assert (g["effect"], g["policy"], g["scope"]) == (wv["effect"], wv.get("policy", ""), wv.get("scope", "")), (cid, a)
            n += 1
    assert n > 100
    eng.close()
