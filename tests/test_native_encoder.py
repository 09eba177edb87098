"""Native batch encoder (cerbos_b200/csrc/cb_encode.h: serialized enginev1.CheckInput -> columns) against the Python
encoder (cerbos_b200/encode.py) on the same inputs: all twelve columns byte for byte -- the reference's engine goldens
(principal / role policies, scopes, globs, JWT aux data, lenient scope search) and the synthetic workloads."""
import numpy as np
import pytest

from cerbos_b200 import wire
import workloads as W
from cerbos_b200.encode import Encoder
from cerbos_b200.table.flatten import flatten
from helpers import engine_decisions, store_rule_table
from hostsim import driver as hostsim


def _same(py_batch, cols, dims):
    assert dims == [py_batch.max_actions, py_batch.role_cols, py_batch.kc, py_batch.n_pass]
    for i, (a, b) in enumerate(zip(py_batch.columns, cols)):
        want = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        assert want.size == b.size and (want == b).all(), f"column {i} differs"


@pytest.mark.parametrize("lenient", [False, True])
def test_goldens_byte_for_byte(lenient):
    ft = flatten(store_rule_table(), globals_={"environment": "test"})
    inputs = [inp for _, len_, inp, _ in engine_decisions() if len_ == lenient]
    assert len(inputs) >= 3
    py = Encoder(ft.manifest, lenient_scope_search=lenient).encode(inputs)
    cols, dims = hostsim.native_encode(ft.blob, [wire.check_input(i) for i in inputs], lenient=lenient)
    _same(py, cols, dims)
    for inp in inputs[:40]:     # and one by one (different batch-level dictionaries every time)
        _same(Encoder(ft.manifest, lenient_scope_search=lenient).encode([inp]),
              *hostsim.native_encode(ft.blob, [wire.check_input(inp)], lenient=lenient))


@pytest.mark.parametrize("name,n", [("C1", 200), ("C2", 3000), ("C3", 2000), ("C5", 300)])
def test_workloads_byte_for_byte(name, n):
    w = W.WORKLOADS[name]()
    _, ft, enc = W.build(w)
    f = w.fields(n)
    inputs = w.inputs(f, range(n))
    msgs = [wire.check_input(i) for i in inputs]
    py = enc.encode(inputs)
    _same(py, *hostsim.native_encode(ft.blob, msgs))
    for threads in (2, 5):       # shards encoded concurrently and merged in order: still the same bytes
        _same(py, *hostsim.native_encode(ft.blob, msgs, threads=threads))


def test_defaults_and_odd_shapes():
    ft = flatten(store_rule_table(), globals_={"environment": "test"})
    inputs = [
        {"actions": ["view:public", "x" * 3], "principal": {"id": "a", "roles": ["employee"], "attr": {"m": {"k": [1, 2.5, None, True, {"z": "s"}], "e": {}}, "l": []}},
         "resource": {"kind": "leave_request", "id": "1", "scope": ".acme.hr", "attr": {"owner": "a", "n": -0.0}}},
        {"actions": [], "principal": {"id": "", "roles": ["r1", "r2", "r3"], "policyVersion": "20210210"}, "resource": {"kind": "nope:kind", "id": "2", "policyVersion": "zzz"}},
        {"actions": ["a"], "principal": {"id": "b", "roles": ["employee"], "scope": "acme.hr.uk.nowhere"}, "resource": {"kind": "leave_request", "id": "3"},
         "auxData": {"jwt": {"aud": ["x", "y"], "iss": "me", "nested": {"deep": {"er": 1}}}}},
    ]
    for dv, ds, len_ in (("default", "", False), ("20210210", "acme", True)):
        py = Encoder(ft.manifest, default_version=dv, default_scope=ds, lenient_scope_search=len_).encode(inputs)
        _same(py, *hostsim.native_encode(ft.blob, [wire.check_input(i) for i in inputs], default_version=dv, default_scope=ds, lenient=len_))


def test_malformed_message_is_rejected():
    ft = flatten(store_rule_table(), globals_={"environment": "test"})
    with pytest.raises(RuntimeError):
        hostsim.native_encode(ft.blob, [b"\x12\xff\xff\xff"])


def test_library_entry_points_without_a_device():
    """cgpu_encoder_create / cgpu_encode / cgpu_encoded_batch of the product library (no GPU needed: marshalling only)."""
    from cerbos_b200 import capi
    w = W.C3()
    _, ft, enc = W.build(w)
    inputs = w.inputs(w.fields(500), range(500))
    ne = capi.NativeEncoder(ft.blob)
    eb = ne.encode([wire.check_input(i) for i in inputs])
    py = enc.encode(inputs)
    b = eb.batch(123)
    assert (b.n_requests, b.max_actions, b.now_unix_nanos, b.n_columns) == (500, 8, 123, 12)
    for i, (a, c) in enumerate(zip(py.columns, eb.columns())):
        assert (np.ascontiguousarray(a).view(np.uint8).reshape(-1) == c).all(), i
    eb.free()
    with pytest.raises(capi.CgpuError):
        ne.encode([b"\x12\xff\xff\xff"])
    ne.close()


@pytest.mark.gpu
def test_encode_then_check_on_gpu():
    """Serialized CheckInputs -> cgpu_encode -> cgpu_check (the path a Go PDP takes) against the oracle."""
    import os
    from cerbos_b200 import capi
    from oracle import cref
    for name, n in (("C2", 50000), ("C3", 20000)):
        w = W.WORKLOADS[name]()
        _, ft, enc = W.build(w)
        inputs = w.inputs(w.fields(n), range(n))
        py = enc.encode(inputs)
        want = cref.check(ft.blob, py.columns, py.n, py.max_actions, n_threads=os.cpu_count() or 1)
        c = capi.Context(0)
        t = c.load_table(ft.blob)
        ne = capi.NativeEncoder(ft.blob)
        eb = ne.encode([wire.check_input(i) for i in inputs])
        assert (t.check_encoded(eb) == want).all(), name
        eb.free(); ne.close(); t.release(); c.close()


# ---- native narrowing (cb_narrow.h: cgpu_narrow_build) against its specification, cerbos_b200/narrow.py ------------------
def _arr(ptr, nbytes, dtype):
    import ctypes
    return np.frombuffer(ctypes.string_at(ptr, nbytes), dtype=dtype).copy() if ptr and nbytes else np.zeros(0, dtype=dtype)


def _assert_narrowed_equals_spec(eb, want_batch, n_slots, form):
    """every array and every parameter cgpu_narrow_build produces == narrow.narrow_batch of the same columns"""
    import ctypes
    from cerbos_b200 import narrow as NW
    want = NW.narrow_batch(want_batch, n_slots, v2=(form == 2))
    got = eb.narrow(form)
    if want is None:
        assert got is None
        return None
    assert got is not None
    n = want.n
    b, nr = got.view(0)
    if want.principal_base is not None:
        assert nr.principal_base == want.principal_base and (_arr(nr.principal_id16, n * 2, np.uint16) == want.principal_id).all()
    else:
        assert not nr.principal_id16 and (_arr(nr.principal_id, n * 4, np.uint32) == want.principal_id).all()
    assert nr.hdr_const_mask == want.hdr_const_mask and tuple(nr.hdr_const) == tuple(want.hdr_const)
    assert (_arr(nr.hdr16, want.hdr16.nbytes, np.uint16) == want.hdr16.reshape(-1)).all()
    if want.versions is None:
        assert nr.versions_const == 1 and tuple(nr.versions_value) == tuple(want.versions_value)
    else:
        assert nr.versions_const == 0 and (_arr(nr.versions, n * 2, np.uint8) == want.versions.reshape(-1)).all()
    assert nr.role_cols == want.role_cols and (_arr(nr.roles, want.roles.nbytes, np.uint8) == want.roles.reshape(-1)).all()
    cls = _arr(nr.slot_class, max(n_slots, 1), np.uint8)
    assert (cls[:n_slots] == want.slot_class[:n_slots]).all()
    assert (_arr(nr.slot_base, 4 * max(n_slots, 1), np.uint32)[:n_slots] == want.slot_base[:n_slots]).all()
    assert (_arr(nr.slot_base2, 4 * max(n_slots, 1), np.uint32)[:n_slots] == want.slot_base2[:n_slots]).all()
    for v in range(n_slots):
        assert ctypes.string_at(nr.slot_cols[v], n * NW.ELEM_BYTES[int(cls[v])]) == np.ascontiguousarray(want.slot_cols[v]).tobytes(), v
    assert (nr.heap_bits, nr.heap_u32, nr.heap_base, nr.heap_base2) == (want.heap_bits, 1 if want.heap_u32 else 0, want.heap_base, want.heap_base2)
    for i in range(4, 12):
        tb = np.ascontiguousarray(np.asarray(want.tables[i - 4])).tobytes()
        assert b.column_bytes[i] == len(tb) and ctypes.string_at(b.columns[i], len(tb)) == tb, i
    assert all(b.column_bytes[i] == 0 for i in range(4))
    return got


@pytest.mark.parametrize("block", ["0", "1"])
@pytest.mark.parametrize("name,n", [("C1", 300), ("C2", 3000), ("C3", 2500), ("C5", 1500)])
def test_native_narrowing_equals_narrow_py_on_workloads(name, n, block, monkeypatch):
    """block = "1": the single-block layout the library uses for page-locked memory on a GPU host, here in plain memory"""
    from cerbos_b200 import capi, wire
    monkeypatch.setenv("CERBOS_B200_NARROW_BLOCK", block)
    w = W.WORKLOADS[name]()
    _, ft, enc = W.build(w)
    inputs = w.inputs(w.fields(n), range(n))
    ne = capi.NativeEncoder(ft.blob)
    eb = ne.encode([wire.check_input(i) for i in inputs])
    want_b = enc.encode(inputs)
    for form in (2, 1):
        got = _assert_narrowed_equals_spec(eb, want_b, len(enc.slots), form)
        got.free()
    eb.free()
    ne.close()


@pytest.mark.parametrize("seed", range(6))
def test_native_narrowing_equals_narrow_py_on_random_batches(seed):
    """fuzzed policy sets x requests: mixed-type columns (-> the 8-byte class), nulls, nested values, odd scopes"""
    import random
    from cerbos_b200 import capi, wire
    from cerbos_b200.policy.compile import build_rule_table
    from fuzzgen import rand_policies, rand_request
    r = random.Random(8800 + seed)
    ft = flatten(build_rule_table(rand_policies(r)))
    enc = Encoder(ft.manifest)
    inputs = [rand_request(r) for _ in range(300)]
    ne = capi.NativeEncoder(ft.blob)
    eb = ne.encode([wire.check_input(i) for i in inputs])
    want_b = enc.encode(inputs)
    for form in (2, 1):
        got = _assert_narrowed_equals_spec(eb, want_b, len(enc.slots), form)
        if got is not None:
            got.free()
    eb.free()
    ne.close()


@pytest.mark.parametrize("family", ["val", "core", "time", "ip"])
def test_value_fuzz_requests_byte_for_byte(family):
    """The request generators of tests/fuzz_values.py (attributes that change type from request to request, Unicode and astral
    characters, numbers beyond 2^53, nulls, nested lists and maps, malformed timestamps / durations / addresses): native
    encoder == Python encoder, every column, single- and multi-threaded."""
    import random
    import fuzz_values as FV
    from test_fuzz_values import _table
    gen, req = {"val": (FV.B, FV.rand_request), "core": (FV.CB, FV.rand_core_request), "time": (FV.TB, FV.rand_time_request),
                "ip": (FV.IPB, FV.rand_ip_request)}[family]
    for seed in range(3):
        r = random.Random(123000 + seed)
        es = []
        while len(es) < 8:
            e = gen(r)
            try:
                _table([e])
                es.append(e)
            except Exception:  # noqa: BLE001 -- a construct the table build refuses: drawn again
                pass
        _, ft = _table(es)
        inputs = [dict(req(r), actions=[f"a{i}" for i in range(len(es))]) for _ in range(50)]
        py = Encoder(ft.manifest).encode(inputs)
        msgs = [wire.check_input(i) for i in inputs]
        _same(py, *hostsim.native_encode(ft.blob, msgs))
        _same(py, *hostsim.native_encode(ft.blob, msgs, threads=3))
