"""CPU tier: the oracle's restatement of the SLS wire format (hand-rolled protobuf writer of the reference) is pinned
(1) on the reference's own unit-test cases (core/unittest/serializer/SLSSerializerUnittest.cpp:82-147: the five
LOG cases, whose expectations are the parsed fields), and (2) byte for byte against the protobuf runtime's canonical
encoding of the same messages, built from a descriptor that restates core/protobuf/sls/sls_logs.proto."""
import random

import pytest

from oracle import oracle as orc

descriptor_pb2 = pytest.importorskip("google.protobuf.descriptor_pb2")


def _messages():
    from google.protobuf import descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="sls_logs_restated.proto", package="sls_logs_restated", syntax="proto2")
    F = descriptor_pb2.FieldDescriptorProto

    def field(m, name, num, ftype, label, type_name=None):
        f = m.field.add(name=name, number=num, type=ftype, label=label)
        if type_name:
            f.type_name = type_name

    log = fd.message_type.add(name="Log")
    content = log.nested_type.add(name="Content")
    field(content, "Key", 1, F.TYPE_BYTES, F.LABEL_REQUIRED)
    field(content, "Value", 2, F.TYPE_BYTES, F.LABEL_REQUIRED)
    field(log, "Time", 1, F.TYPE_UINT32, F.LABEL_REQUIRED)
    field(log, "Contents", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".sls_logs_restated.Log.Content")
    field(log, "values", 3, F.TYPE_BYTES, F.LABEL_REPEATED)
    field(log, "Time_ns", 4, F.TYPE_FIXED32, F.LABEL_OPTIONAL)
    tag = fd.message_type.add(name="LogTag")
    field(tag, "Key", 1, F.TYPE_BYTES, F.LABEL_REQUIRED)
    field(tag, "Value", 2, F.TYPE_BYTES, F.LABEL_REQUIRED)
    grp = fd.message_type.add(name="LogGroup")
    field(grp, "Logs", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".sls_logs_restated.Log")
    field(grp, "Category", 2, F.TYPE_BYTES, F.LABEL_OPTIONAL)
    field(grp, "Topic", 3, F.TYPE_BYTES, F.LABEL_OPTIONAL)
    field(grp, "Source", 4, F.TYPE_BYTES, F.LABEL_OPTIONAL)
    field(grp, "MachineUUID", 5, F.TYPE_BYTES, F.LABEL_OPTIONAL)
    field(grp, "LogTags", 6, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".sls_logs_restated.LogTag")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = message_factory.GetMessageClass
    return get(pool.FindMessageTypeByName("sls_logs_restated.LogGroup"))


def _fixture_group(enable_nanosecond, with_empty=False, with_non_empty=True):
    """CreateBatchedLogEvents (SLSSerializerUnittest.cpp:776-808)"""
    g = orc.Group()
    g.tags = {"__topic__": "topic", "__source__": "source", "__machine_uuid__": "machine_uuid",
              "__pack_id__": "pack_id"}
    for non_empty in ([True] if with_non_empty else []) + ([False] if with_empty else []):
        e = orc.Event()
        if non_empty:
            e.set(b"key", b"value")
        e.timestamp = 1234567890
        e.ns = 1 if enable_nanosecond else None
        g.events.append(e)
    return g


def test_reference_unit_test_cases():
    LogGroup = _messages()

    def parse(b):
        m = LogGroup()
        m.ParseFromString(b)
        return m

    res, err = orc.sls_serialize_group(_fixture_group(False), enable_ns=False)  # :85-102
    m = parse(res)
    assert len(m.Logs) == 1 and len(m.Logs[0].Contents) == 1
    assert (m.Logs[0].Contents[0].Key, m.Logs[0].Contents[0].Value) == (b"key", b"value")
    assert m.Logs[0].Time == 1234567890 and not m.Logs[0].HasField("Time_ns")
    assert [(t.Key, t.Value) for t in m.LogTags] == [(b"__pack_id__", b"pack_id")]
    assert (m.MachineUUID, m.Source, m.Topic) == (b"machine_uuid", b"source", b"topic")
    m = parse(orc.sls_serialize_group(_fixture_group(True), enable_ns=True)[0])  # :103-113
    assert m.Logs[0].Time == 1234567890 and m.Logs[0].Time_ns == 1
    m = parse(orc.sls_serialize_group(_fixture_group(False), enable_ns=True)[0])  # :114-124
    assert m.Logs[0].Time == 1234567890 and not m.Logs[0].HasField("Time_ns")
    m = parse(orc.sls_serialize_group(_fixture_group(False, True, True), enable_ns=False)[0])  # :125-142
    assert len(m.Logs) == 1 and len(m.Logs[0].Contents) == 1 and m.Logs[0].Time == 1234567890
    res, err = orc.sls_serialize_group(_fixture_group(False, True, False), enable_ns=False)  # :143-147
    assert res is None and err


def test_logs_are_byte_identical_to_protobuf_runtime():
    LogGroup = _messages()
    rng = random.Random(99)
    for _ in range(300):
        events = []
        m = LogGroup()
        for _ in range(rng.randint(1, 12)):
            t = rng.choice([0, 5, (1 << 28) - 1, 1 << 28, 1234567890, 0xFFFFFFFF])
            ns = rng.choice([None, 0, 1, 999999999])
            contents = [(bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 7, 127, 128, 300]))),
                         bytes(rng.randrange(256) for _ in range(rng.choice([0, 3, 127, 128, 200, 17000]))))
                        for _ in range(rng.randint(0, 6))]
            events.append((t, ns, contents))
            if contents:
                lg = m.Logs.add()
                lg.Time = max(t, orc.SLS_MIN_LOG_TIME)
                for k, v in contents:
                    c = lg.Contents.add()
                    c.Key, c.Value = k, v
                if ns is not None:
                    lg.Time_ns = ns
        got, offs = orc.sls_serialize_logs(events, enable_ns=True)
        assert got == m.SerializeToString()
        assert len(offs) == len(m.Logs) and all(got[o] == 0x0A for o in offs)


def _random_events(rng, n_max=40):
    events = []
    for _ in range(rng.randint(0, n_max)):
        t = rng.choice([0, 5, (1 << 28) - 1, 1 << 28, 1234567890, 0xFFFFFFFF])
        ns = rng.choice([None, 0, 1, 999999999])
        contents = [(bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 7, 127, 128, 300]))),
                     bytes(rng.randrange(256) for _ in range(rng.choice([0, 3, 127, 128, 200, 17000]))))
                    for _ in range(rng.choice([0, 0, 1, 2, 6, 11]))]
        events.append((t, ns, contents))
    return events


def test_kernel_size_and_emit_functions_on_the_host():
    """csrc/lc_exec.cuh: lc_sls_log_size / lc_sls_emit_log (the arithmetic of the sls kernels) compiled for the host."""
    from tests.emul import emul
    rng = random.Random(7)
    for _ in range(200):
        events = _random_events(rng)
        for enable_ns in (True, False):
            want, _ = orc.sls_serialize_logs(events, enable_ns)
            assert emul.sls_serialize_logs(events, enable_ns) == want
