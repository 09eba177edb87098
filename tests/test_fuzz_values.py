"""Differential fuzzing of the functions that build values at run time (device scratch arena: cel-go ext.Strings / ext.Lists /
ext.Encoders, Cerbos except / intersect, collecting comprehensions, concatenation, RE2 matches over a DFA): random typed
expressions (tests/fuzz_values.py) x random requests whose attributes change type, oracle #1 against the kernel core compiled
for the host.  Every expression is checked together with its negation, so "false" and "error" (both DENY) are told apart;
a third of the expressions are identities (x.reverse().reverse() == x ...) that are true whenever nothing fails.
oracle #2 does not port these functions (it flags them), so this is the two-way check that covers them."""
import random

import pytest

import fuzz_values as FV
from cerbos_b200.encode import Encoder
from cerbos_b200.policy.compile import build_rule_table
from cerbos_b200.table.flatten import flatten
from hostsim import driver as hostsim
from oracle.check import CheckOracle


def _table(es):
    rules = [{"actions": [f"a{i}"], "effect": "EFFECT_ALLOW", "roles": ["*"], "condition": {"match": {"expr": e}}} for i, e in enumerate(es)]
    pol = {"apiVersion": "api.cerbos.dev/v1", "resourcePolicy": {"resource": "doc", "version": "default", "rules": rules}}
    rt = build_rule_table([pol])
    return rt, flatten(rt)


def run_seed(seed, n_expr=10, n_req=30):
    """-> (mismatches, requests the device flags as outside its exact range, comparisons, allows)"""
    r = random.Random(50000 + seed)
    es = []
    while len(es) < n_expr:
        k = r.random()
        e = FV.identity(r) if k < 0.3 else FV.M(r) if k < 0.45 else FV.B(r)
        try:
            _table([e])           # (a construct the table build rejects -- string(x), a pattern the DFA compiler refuses -- is drawn again)
            es.append(e)
        except Exception:  # noqa: BLE001
            pass
    es = es + [f"!({e})" for e in es]
    rt, ft = _table(es)
    orc = CheckOracle(rt)
    enc = Encoder(ft.manifest)
    mism, flagged, total, allows = [], 0, 0, 0
    for _ in range(n_req):
        inp = dict(FV.rand_request(r), actions=[f"a{i}" for i in range(len(es))])
        want = orc.check(inp)["actions"]
        b = enc.encode([inp])
        try:
            got = hostsim.check(ft.blob, b.columns, 1, b.max_actions)
        except RuntimeError as x:
            if "-2" not in str(x):
                raise
            flagged += 1          # a run-time value outside the device's exact range: the call fails loudly, nothing to compare
            continue
        for i, e in enumerate(es):
            total += 1
            allows += want[f"a{i}"]["effect"] == 1
            if got[0, i] != want[f"a{i}"]["effect"]:
                mism.append((e, inp, int(got[0, i])))
    return mism, flagged, total, allows


@pytest.mark.parametrize("seed", range(24))
def test_run_time_values_random(seed):
    mism, flagged, total, allows = run_seed(seed)
    assert not mism, mism[:3]
    assert total >= 200 and flagged <= 20 and allows >= 10     # (an expression or its negation is ALLOWed whenever the evaluation succeeds)


def run_time_seed(seed, n_expr=8, n_req=40, gen=None, req=None):
    """timestamps / durations (or another family: gen / req): oracle #1, oracle #2 (it ports all of this) and the kernel core"""
    from oracle import cref
    from oracle.celeval import parse_timestamp
    now = parse_timestamp("2024-03-10T06:59:59.5Z")      # half a second before the US spring-forward hour
    r = random.Random(81000 + seed)
    gen, req = gen or FV.TB, req or FV.rand_time_request
    es = []
    probe = dict(req(random.Random(seed)), actions=["a0"])
    while len(es) < n_expr:
        e = gen(r)
        try:
            _, ft1 = _table([e])
            b1 = Encoder(ft1.manifest).encode([probe])
            hostsim.check(ft1.blob, b1.columns, 1, 1, now.ns)      # (an expression that flags whatever the request holds --
            es.append(e)                                            # string(0.5) -- would void the whole seed: drawn again)
        except Exception:  # noqa: BLE001
            pass
    es = es + [f"!({e})" for e in es]
    rt, ft = _table(es)
    orc = CheckOracle(rt)
    enc = Encoder(ft.manifest)
    mism, flagged, total, allows = [], 0, 0, 0
    for _ in range(n_req):
        inp = dict(req(r), actions=[f"a{i}" for i in range(len(es))])
        want = orc.check(inp, now)["actions"]
        b = enc.encode([inp])
        outs = []
        for fn in (hostsim.check, cref.check):
            try:
                outs.append(fn(ft.blob, b.columns, 1, b.max_actions, now.ns))
            except RuntimeError as x:
                if "-2" not in str(x):
                    raise
                outs.append(None)
        if outs[0] is None:
            # a value outside the device's exact range (a timestamp beyond 1678..2262 in nanoseconds ...): the call fails
            # loudly -- nothing to compare.  (oracle #2 flags the functions it does not port: then it has no opinion.)
            flagged += 1
            continue
        for i, e in enumerate(es):
            total += 1
            w = want[f"a{i}"]["effect"]
            allows += w == 1
            if outs[0][0, i] != w or (outs[1] is not None and outs[1][0, i] != w):
                mism.append((e, inp["principal"]["attr"], inp["resource"]["attr"], int(outs[0][0, i]), int(outs[1][0, i]) if outs[1] is not None else None, w))
    return mism, flagged, total, allows


@pytest.mark.parametrize("seed", range(16))
def test_time_values_random(seed):
    mism, flagged, total, allows = run_time_seed(seed)
    assert not mism, mism[:3]
    assert total >= 100 and allows >= 5


@pytest.mark.parametrize("seed", range(16))
def test_core_semantics_random(seed):
    """heterogeneous equality, cross-type ordering, conversions and their range errors, overflow, maps, has(), index errors,
    && / || error absorption (tests/fuzz_values.py: CV / CB) -- three ways"""
    mism, flagged, total, allows = run_time_seed(5000 + seed, n_expr=10, n_req=40, gen=FV.CB, req=FV.rand_core_request)
    assert not mism, mism[:3]
    assert total >= 100 and allows >= 5


@pytest.mark.parametrize("seed", range(8))
def test_ip_ranges_random(seed):
    """inIPAddrRange over well-formed and malformed IPv4 / IPv6 / IPv4-mapped texts and CIDRs (valid, out of range, malformed) -- three ways"""
    mism, flagged, total, allows = run_time_seed(9000 + seed, n_expr=10, n_req=60, gen=FV.IPB, req=FV.rand_ip_request)
    assert not mism, mism[:3]
    assert total >= 1000 and flagged == 0 and allows >= 20
