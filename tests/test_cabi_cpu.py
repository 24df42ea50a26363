"""CPU tier: the C-ABI library loads, exports every symbol include/*.h declares, compiles regexes on the
host, fails loudly without a GPU, and the host layer's GPU-free paths reproduce the reference fixtures."""
import ctypes
import os
import re

import pytest

import loongcollector_b200 as lc
from tests.golden_util import load_cases, input_with_metadata, strip_event_meta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("lc_b200.h", "lc_b200_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"\b(lc_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = lc.lib()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "missing export: " + s


def test_version_and_regex_compile_on_host():
    assert "sm_100a" in lc.version()
    r = lc.Regex(r"(\w+)\t(\w+).*")
    assert r.ngroups == 2 and r.info["mode"] == 0
    with pytest.raises(lc.LcError) as ei:
        lc.Regex(r"(a)\1")
    assert ei.value.code == 4


@pytest.mark.skipif(lc.device_count() > 0, reason="needs a box WITHOUT a GPU")
def test_no_cpu_fallback():
    with pytest.raises(lc.LcError) as ei:
        lc.Engine(0)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    # the host layer fails just as loudly when its processors need the engine
    p = lc.HostProcessor("processor_split_string_native", {})
    with pytest.raises(lc.LcError):
        p.process({"events": [{"type": 1, "timestamp": 1, "contents": {"content": "a\nb"}}]})


def test_host_layer_whole_line_mode_needs_no_gpu():
    """Regex "(.*)" is the reference's whole-line shortcut (ProcessorParseRegexNative.cpp:68,147-148): no regex
    runs, so the reference fixture replays without a device."""
    import json
    case = [c for c in load_cases("regex") if c["fn"] == "TestProcessWholeLine"][0]
    step = case["pipeline"][0]
    p = lc.HostProcessor(step["type"], step["config"])
    out = p.process(input_with_metadata(case), True)
    assert json.dumps(strip_event_meta(out), sort_keys=True) == json.dumps(case["expected"], sort_keys=True)


def test_host_layer_init_errors():
    with pytest.raises(lc.LcError):
        lc.HostProcessor("processor_parse_regex_native", {"SourceKey": "content", "Keys": ["a"]})  # no Regex
    with pytest.raises(lc.LcError):
        lc.HostProcessor("processor_parse_regex_native", {"SourceKey": "c", "Regex": "(a", "Keys": ["a"]})
    with pytest.raises(lc.LcError):
        lc.HostProcessor("processor_parse_delimiter_native", {"SourceKey": "c", "Separator": "12345", "Keys": ["a"]})
    with pytest.raises(lc.LcError):
        lc.HostProcessor("no_such_processor", {})


def _roundtrip(so, cfg, group):
    import json
    L = lc.lib()
    L.lc_host_dynamic_plugin_roundtrip.restype = ctypes.c_void_p
    L.lc_host_dynamic_plugin_roundtrip.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_void_p),
                                                   ctypes.POINTER(ctypes.c_void_p)]
    L.lc_host_string_free.argtypes = [ctypes.c_void_p]
    ver, name, err = ctypes.c_int(-1), ctypes.c_void_p(), ctypes.c_void_p()
    out = L.lc_host_dynamic_plugin_roundtrip(so.encode(), json.dumps(cfg).encode(),
                                             None if group is None else json.dumps(group).encode(), 1,
                                             ctypes.byref(ver), ctypes.byref(name), ctypes.byref(err))
    nm = ctypes.string_at(name.value).decode() if name.value else None
    er = ctypes.string_at(err.value).decode() if err.value else None
    res = json.loads(ctypes.string_at(out).decode()) if out else None
    for p in (out, name.value, err.value):
        if p:
            L.lc_host_string_free(p)
    return ver.value, nm, res, er


def test_dynamic_plugins_export_processor_interface_and_init_like_the_agent_loads_them():
    """lib<name>.so per processor: dlopen + dlsym("processor_interface") + version == 100, then init / finalize the way
    DynamicCProcessorProxy drives them (CProcessor.h:23-45, PluginRegistry.cpp:255-275).  Init needs no GPU."""
    from loongcollector_b200 import _build
    cfgs = {
        "processor_parse_regex_b200": {"SourceKey": "content", "Regex": r"(\w+)\t(\w+).*", "Keys": ["a", "b"]},
        "processor_parse_delimiter_b200": {"SourceKey": "content", "Separator": ",", "Keys": ["a", "b"]},
        "processor_split_string_b200": {},
        "processor_split_multiline_log_string_b200": {"Multiline": {"StartPattern": r"\d+-.*"}},
    }
    for name, ptype in _build.PLUGINS:
        so = _build.plugin_path(name)
        assert os.path.exists(so), so
        raw = ctypes.CDLL(so)
        assert hasattr(raw, "processor_interface")
        ver, nm, res, err = _roundtrip(so, cfgs[name], None)
        assert (ver, nm, res, err) == (100, name, None, None) or (ver, nm, err) == (100, name, None), (name, err)
    # a config the reference's Init rejects: init returns non-zero and leaves plugin_state NULL
    ver, nm, res, err = _roundtrip(_build.plugin_path("processor_parse_regex_b200"),
                                   {"SourceKey": "content", "Regex": "(a", "Keys": ["a"]}, None)
    assert ver == 100 and res is None and err == "init returned non-zero"
    # whole-line mode needs no device: the full init -> process -> finalize round trip runs here
    grp = {"events": [{"type": 1, "timestamp": 3, "timestampNanosecond": 0, "contents": {"content": "l1\nl2"}}]}
    ver, nm, res, err = _roundtrip(_build.plugin_path("processor_parse_regex_b200"),
                                   {"SourceKey": "content", "Regex": "(.*)", "Keys": ["msg"]}, grp)
    assert err is None and res["events"][0]["contents"] == {"msg": "l1\nl2"}
