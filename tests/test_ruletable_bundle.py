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


def test_encrypted_bundle_decrypts_to_the_same_rule_table(bundle):
    """bundle.crrts + encryption_key.txt of the reference's fixture (crypto.DecryptChaCha20Poly1305Stream, ruletable_bundle.go:
    56-70): every 64 KiB chunk authenticates under the STREAM nonces, and the message decodes to the rule table of
    bundle_unencrypted.crrt (Go serialises maps in random order, so the bytes differ; rows, conditions, parent roles agree)."""
    from cerbos_b200.table import bundle_crypto as BC
    enc = open(os.path.join(GOLDEN, "ruletable_bundle_encrypted.crrts"), "rb").read()
    key = open(os.path.join(GOLDEN, "ruletable_bundle_encryption_key.txt")).read()
    assert len(enc) == len(bundle) + 3 * BC.TAG            # three chunks, one tag each, nothing else
    plain = BC.decrypt_stream(key, enc)
    assert len(plain) == len(bundle)
    a, b = decode_rule_table(plain), decode_rule_table(bundle)
    assert sorted(map(_row_key, a.rows), key=repr) == sorted(map(_row_key, b.rows), key=repr)
    conds = lambda rt: sorted(repr((r.condition, r.dr_condition)) for r in rt.rows)   # noqa: E731
    assert conds(a) == conds(b)
    spr = lambda rt: {s: {r: sorted(p) for r, p in m.items() if p} for s, m in rt.scope_parent_roles.items() if any(m.values())}   # noqa: E731
    pdr = lambda rt: {f: sorted(d) for f, d in rt.policy_derived_roles.items() if d}   # noqa: E731
    assert spr(a) == spr(b) and pdr(a) == pdr(b)
    # the written-out cipher gives the library's bytes; a wrong key, a flipped bit, a dropped chunk are refused
    k = BC.parse_key(key)
    n = 3
    pure = b"".join(BC.open_chacha20poly1305(k, i.to_bytes(11, "big") + (b"\x01" if i == n - 1 else b"\x00"),
                                             enc[i * (BC.CHUNK + BC.TAG):(i + 1) * (BC.CHUNK + BC.TAG)]) for i in range(n))
    assert pure == plain
    for bad in (lambda: BC.decrypt_stream("00" * 32, enc), lambda: BC.decrypt_stream(key, enc[:100] + bytes([enc[100] ^ 1]) + enc[101:]),
                lambda: BC.decrypt_stream(key, enc[: BC.CHUNK + BC.TAG]), lambda: BC.decrypt_stream(key[:10], enc)):
        with pytest.raises(BC.BundleCryptoError):
            bad()


def test_chacha20poly1305_known_answer():
    """RFC 8439 section 2.8.2: the AEAD test vector, against the implementation written out in bundle_crypto.py."""
    from cerbos_b200.table import bundle_crypto as BC
    key = bytes(range(0x80, 0xA0))
    nonce = bytes.fromhex("070000004041424344454647")
    aad = bytes.fromhex("50515253c0c1c2c3c4c5c6c7")
    pt = b"Ladies and Gentlemen of the class of '99: If I could offer you only one tip for the future, sunscreen would be it."
    ct = BC._chacha20_xor(key, nonce, 1, pt)
    assert ct[:16].hex() == "d31a8d34648e60db7b86afbc53ef7ec2" and ct[-2:].hex() == "6116"
    otk = BC._chacha20_block(list(__import__("struct").unpack("<8I", key)), 0, list(__import__("struct").unpack("<3I", nonce)))[:32]
    tag = BC._poly1305(otk, aad + BC._pad16(aad) + ct + BC._pad16(ct) + __import__("struct").pack("<QQ", len(aad), len(ct)))
    assert tag.hex() == "1ae10b594f09e26a7e902ecbd0600691"
    assert BC.open_chacha20poly1305(key, nonce, ct + tag, aad) == pt


def test_wire_reader_rejects_truncated_input(bundle):
    with pytest.raises(WireError):
        decode_rule_table(bundle[: len(bundle) // 2 + 1])
    assert list(fields(b"")) == []


@pytest.mark.gpu
def test_engine_from_bundle_on_gpu(bundle):
    """Engine.from_rule_table_bundle: the same goldens on the device."""
    from cerbos_b200.engine import Engine
    # (the encrypted form of the same bundle, as a PDP receives it from Cerbos Hub)
    enc = open(os.path.join(GOLDEN, "ruletable_bundle_encrypted.crrts"), "rb").read()
    key = open(os.path.join(GOLDEN, "ruletable_bundle_encryption_key.txt")).read()
    eng = Engine.from_rule_table_bundle(enc, key=key, globals_=G)
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
