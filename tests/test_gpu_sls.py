"""GPU tier: SLS wire format of LOG events (next row) -- the C-ABI kernels and the host-layer
SLSEventGroupSerializer against the oracle restatement (which tests/test_oracle_sls.py pins on the reference's
unit-test cases and on the protobuf runtime)."""
import json
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402  (checker only)
from tests.test_oracle_sls import _fixture_group, _random_events  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    import loongcollector_b200 as lc
    e = lc.Engine(0)
    yield e
    e.close()


def test_logs_bytes_match_oracle_on_random_events(eng):
    rng = random.Random(2024)
    for _ in range(40):
        events = _random_events(rng, 60)
        for enable_ns in (True, False):
            want, _ = orc.sls_serialize_logs(events, enable_ns)
            assert eng.sls_serialize_logs(events, enable_ns) == want


def test_logs_bytes_large_parsed_batch(eng):
    """20000 events with the 10 nginx fields each (the shape the regex stage produces)."""
    from loongcollector_b200 import synth
    buf, off, ln = synth.nginx_lines(20000, seed=5, line_bytes=None)
    o = orc.Regex(synth.NGINX_PATTERN)
    st, co, cl = orc.regex_parse_batch(o, buf, off, ln, 10)
    keys = [k.encode() for k in synth.NGINX_KEYS]
    events = []
    for i in range(off.size):
        contents = []
        if st[i] == 0:
            contents = [(keys[g], bytes(buf[co[i, g]:co[i, g] + cl[i, g]])) for g in range(10)]
        events.append((1700000000 + i, i % 1000 if i % 3 else None, contents))
    want, offs = orc.sls_serialize_logs(events, True)
    got = eng.sls_serialize_logs(events, True)
    assert got == want and len(offs) == int((st == 0).sum())


def test_capacity_error_reports_needed_size(eng):
    import ctypes as C

    import loongcollector_b200 as lc
    from loongcollector_b200 import capi
    base = np.frombuffer(b"keyvalue", np.uint8)
    need = C.c_uint64(0)
    out = np.zeros(4, np.uint8)
    rc = capi.lib().lc_sls_serialize_logs(eng._h, capi._p(base), 8, 1, capi._p(np.array([1234567890], np.uint32)), None,
                                          capi._p(np.array([0, 1], np.uint64)), capi._p(np.array([0], np.uint32)),
                                          capi._p(np.array([3], np.uint32)), capi._p(np.array([3], np.uint32)),
                                          capi._p(np.array([5], np.uint32)), capi._p(out), 4, C.byref(need))
    want, _ = orc.sls_serialize_logs([(1234567890, None, [(b"key", b"value")])], False)
    assert rc == lc.capi.LC_ERR_CAPACITY and need.value == len(want)


def test_host_serializer_matches_oracle():
    from loongcollector_b200 import capi
    rng = random.Random(11)
    # the reference's five unit-test cases (SLSSerializerUnittest.cpp:82-147)
    for args, ns in (((False,), False), ((True,), True), ((False,), True), ((False, True, True), False),
                     ((False, True, False), False)):
        g = _fixture_group(*args)
        want, werr = orc.sls_serialize_group(g, ns)
        got, gerr = capi.host_sls_serialize(g.to_json(True), ns)
        assert got == want and (gerr is None) == (werr is None), (args, ns, gerr, werr)
    for _ in range(30):
        evs = []
        for _ in range(rng.randint(1, 40)):
            contents = {"k%d" % j: "".join(rng.choice('ab "xyz\n') for _ in range(rng.choice([0, 1, 5, 130])))
                        for j in range(rng.choice([0, 1, 3, 9]))}
            ev = {"type": 1, "timestamp": rng.choice([5, 1234567890]), "contents": contents}
            if rng.random() < 0.5:
                ev["timestampNanosecond"] = rng.randint(0, 999999999)
            evs.append(ev)
        root = {"events": evs, "tags": {"__topic__": "t", "__source__": "1.2.3.4", "__pack_id__": "ABCD-1",
                                        "host.name": "h" * rng.choice([1, 200])}}
        for ns in (False, True):
            g = orc.Group.from_json(json.loads(json.dumps(root)))
            want, werr = orc.sls_serialize_group(g, ns)
            got, gerr = capi.host_sls_serialize(root, ns)
            assert got == want and (gerr is None) == (werr is None), (root, ns, gerr, werr)
