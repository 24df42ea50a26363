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
