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


@pytest.mark.parametrize("fail_key", [None, b"rawLog"])
def test_parsed_tables_to_wire_bytes_on_device(eng, fail_key):
    """regex parse (device tables) -> lc_sls_serialize_parsed_dev: the wire bytes of the parsed events, written from the
    capture tables + constant keys without any host-built span list; == the oracle's serialiser over the events the
    oracle's ProcessorParseRegexNative leaves (10 fields per matching line; failed lines erased or kept as rawLog)."""
    import torch

    import loongcollector_b200 as lc
    from loongcollector_b200 import synth
    n = 30000
    buf, off, ln = synth.nginx_lines(n, seed=9, line_bytes=None, bad_fraction=0.05)
    rx = lc.Regex(synth.NGINX_PATTERN)
    keys = [k.encode() for k in synth.NGINX_KEYS]
    d_buf = torch.from_numpy(buf).cuda()
    d_off = torch.from_numpy(off.view(np.int32)).cuda()
    d_len = torch.from_numpy(ln.view(np.int32)).cuda()
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_co = torch.empty(n * 10, dtype=torch.int32, device="cuda")
    d_cl = torch.empty(n * 10, dtype=torch.int32, device="cuda")
    times = (1700000000 + np.arange(n)).astype(np.uint32)
    nss = np.where(np.arange(n) % 3 == 0, 0xFFFFFFFF, np.arange(n) % 1000).astype(np.uint32)
    d_t = torch.from_numpy(times.view(np.int32)).cuda()
    d_ns = torch.from_numpy(nss.view(np.int32)).cuda()
    cap = int(buf.size) * 2 + 64 * n
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    eng.regex_parse_dev(rx, d_buf.data_ptr(), buf.size, d_off.data_ptr(), d_len.data_ptr(), n, 10, d_st.data_ptr(),
                        d_co.data_ptr(), d_cl.data_ptr())
    got_len = eng.sls_serialize_parsed_dev(d_buf.data_ptr(), buf.size, d_off.data_ptr(), d_len.data_ptr(),
                                           d_st.data_ptr(), d_co.data_ptr(), d_cl.data_ptr(), 10, n, keys, fail_key,
                                           d_t.data_ptr(), d_ns.data_ptr(), d_out.data_ptr(), cap)
    got = bytes(d_out[:got_len].cpu().numpy())
    st, co, cl = orc.regex_parse_batch(orc.Regex(synth.NGINX_PATTERN), buf, off, ln, 10)
    assert (st != 0).sum() > 100
    events = []
    for i in range(n):
        if st[i] == 0:
            contents = [(keys[g], bytes(buf[co[i, g]:co[i, g] + cl[i, g]])) for g in range(10)]
        elif fail_key:
            contents = [(fail_key, bytes(buf[off[i]:off[i] + ln[i]]))]
        else:
            contents = []
        events.append((int(times[i]), None if nss[i] == 0xFFFFFFFF else int(nss[i]), contents))
    want, _ = orc.sls_serialize_logs(events, True)
    assert got == want
    # capacity error reports the needed size
    from loongcollector_b200 import capi
    with pytest.raises(lc.LcError) as ei:
        eng.sls_serialize_parsed_dev(d_buf.data_ptr(), buf.size, d_off.data_ptr(), d_len.data_ptr(), d_st.data_ptr(),
                                     d_co.data_ptr(), d_cl.data_ptr(), 10, n, keys, fail_key, d_t.data_ptr(),
                                     d_ns.data_ptr(), d_out.data_ptr(), 100)
    assert ei.value.code == capi.LC_ERR_CAPACITY
