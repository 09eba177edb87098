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
