"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see lc_oracle.c header).

ctypes binding of oracle/liblc_oracle.so plus an event-group level restatement of the four
reference processors' ``Process(PipelineEventGroup&)`` so that the reference's own unit-test
fixtures (tests/golden/*.json) can be replayed.  Nothing under loongcollector_b200/ imports this.

Reference lines restated here (paths relative to the reference checkout):
  * LogEvent content semantics ........ core/models/LogEvent.cpp:50-106
  * group/event JSON ................. core/models/PipelineEventGroup.cpp:405-483, LogEvent.cpp:170-208,
                                         RawEvent.cpp:52-73
  * split ............................. core/plugin/processor/inner/ProcessorSplitLogStringNative.cpp:74-161
  * multiline ......................... core/plugin/processor/inner/ProcessorSplitMultilineLogStringNative.cpp:95-340
  * regex parse ....................... core/plugin/processor/ProcessorParseRegexNative.cpp:29-253
  * delimiter ......................... core/plugin/processor/ProcessorParseDelimiterNative.cpp:30-409
  * common options .................... core/plugin/processor/CommonParserOptions.cpp:28-117
  * multiline options ................. core/file_server/MultilineOptions.cpp:22-222
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liblc_oracle.so")
    src = os.path.join(_HERE, "lc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    L.orc_regex_compile.restype = C.c_void_p
    L.orc_regex_compile.argtypes = [C.c_char_p, C.c_uint64, C.c_int]
    L.orc_regex_free.argtypes = [C.c_void_p]
    L.orc_regex_ngroups.restype = C.c_uint32
    L.orc_regex_ngroups.argtypes = [C.c_void_p]
    L.orc_matcher_create.restype = C.c_void_p
    L.orc_matcher_create.argtypes = [C.c_void_p]
    L.orc_matcher_free.argtypes = [C.c_void_p]
    L.orc_regex_prefix_match.restype = C.c_int
    L.orc_regex_prefix_match.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.orc_regex_full_match.restype = C.c_int
    L.orc_regex_full_match.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.orc_regex_parse_batch.restype = None
    L.orc_regex_parse_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_regex_match_batch.restype = None
    L.orc_regex_match_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_split_lines.restype = C.c_uint64
    L.orc_split_lines.argtypes = [C.c_void_p, C.c_uint64, C.c_uint8, C.c_void_p, C.c_void_p, C.c_uint64]
    L.orc_multiline_split.restype = C.c_uint64
    L.orc_multiline_split.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_delim_trim.restype = C.c_int
    L.orc_delim_trim.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.orc_delim_fsm.restype = C.c_int64
    L.orc_delim_fsm.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int64]
    L.orc_delim_unquote.restype = C.c_uint32
    L.orc_delim_unquote.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint8, C.c_void_p]
    L.orc_delim_split.restype = C.c_int64
    L.orc_delim_split.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int64]
    L.orc_delim_parse_batch.restype = None
    L.orc_delim_parse_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                        C.c_uint8, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_remove_last_incomplete_log.restype = C.c_int32
    L.orc_remove_last_incomplete_log.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.POINTER(C.c_int32)]
    _LIB = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        assert buf.dtype == np.uint8
        return np.ascontiguousarray(buf)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


# ----------------------------------------------------------------------------- flat API
def _boost_to_pcre2(pattern: bytes) -> bytes:
    """Boost-only escapes PCRE2 reads differently: \\< and \\> (word start / end, perl_matcher::match_word_start /
    match_word_end) are literal '<' / '>' for PCRE2; outside character classes they become \\b(?=\\w) and
    \\b(?<=\\w), which is also how PCRE2 itself expands [[:<:]] and [[:>:]]."""
    out = bytearray()
    i, in_set = 0, False
    while i < len(pattern):
        c = pattern[i:i + 1]
        if c == b"\\" and i + 1 < len(pattern):
            e = pattern[i + 1:i + 2]
            if not in_set and e == b"Q":  # \\Q ... \\E is literal text
                j = pattern.find(b"\\E", i + 2)
                j = len(pattern) if j < 0 else j + 2
                out += pattern[i:j]
                i = j
                continue
            if not in_set and e == b"<":
                out += b"\\b(?=\\w)"
            elif not in_set and e == b">":
                out += b"\\b(?<=\\w)"
            else:
                out += c + e
            i += 2
            continue
        if c == b"[" and not in_set:
            in_set = True
            out += c
            i += 1
            if pattern[i:i + 1] == b"^":
                out += b"^"
                i += 1
            if pattern[i:i + 1] == b"]":  # a leading ']' is a literal
                out += b"]"
                i += 1
            continue
        if c == b"]" and in_set:
            in_set = False
        out += c
        i += 1
    return bytes(out)


class Regex:
    """boost::regex(pattern) stand-in (PCRE2, DOTALL|MULTILINE, bytes). One matcher scratch per object."""

    def __init__(self, pattern, jit: bool = False):
        if isinstance(pattern, str):
            pattern = pattern.encode("utf-8")
        self.pattern = pattern
        pattern = _boost_to_pcre2(pattern)
        L = lib()
        self._re = L.orc_regex_compile(pattern, len(pattern), 1 if jit else 0)
        if not self._re:
            raise ValueError("oracle: invalid regex %r" % (pattern,))
        self.ngroups = int(L.orc_regex_ngroups(self._re))
        self._m = L.orc_matcher_create(self._re)

    def new_matcher(self):
        """An extra matcher scratch over the same compiled code (one per CPU thread)."""
        return lib().orc_matcher_create(self._re)

    def prefix_match(self, data: bytes) -> bool:
        a = _as_u8(data)
        return bool(lib().orc_regex_prefix_match(self._m, _ptr(a) if a.size else None, a.size))

    def full_match(self, data: bytes):
        """-> None, or list of (off,len) for groups 1..ngroups (unset = (len(data), 0))."""
        a = _as_u8(data)
        co = np.zeros(max(self.ngroups, 1), np.uint32)
        cl = np.zeros(max(self.ngroups, 1), np.uint32)
        ok = lib().orc_regex_full_match(self._m, _ptr(a) if a.size else None, a.size, _ptr(co), _ptr(cl))
        if not ok:
            return None
        return [(int(co[g]), int(cl[g])) for g in range(self.ngroups)]

    def __del__(self):
        try:
            L = lib()
            if getattr(self, "_m", None):
                L.orc_matcher_free(self._m)
            if getattr(self, "_re", None):
                L.orc_regex_free(self._re)
        except Exception:
            pass


def remove_last_incomplete_log(buf, start: Optional["Regex"], end: Optional["Regex"], allow_rollback: bool = True):
    """LogFileReader::RemoveLastIncompleteLog (raw text): -> (bytes to keep, rollbackLineFeedCount)."""
    a = _as_u8(buf)
    padded = np.concatenate([a, np.zeros(1, np.uint8)])  # the reader's buffer is NUL terminated one past the end
    rb = C.c_int32(0)
    keep = lib().orc_remove_last_incomplete_log(_ptr(padded), a.size, start._m if start else None,
                                                end._m if end else None, 1 if allow_rollback else 0, C.byref(rb))
    return int(keep), int(rb.value)


def multiline_regs(cfg: dict):
    """(start, end) Regex objects a LogFileReader holds for this Multiline config: MultilineOptions::Init
    (core/file_server/MultilineOptions.cpp:125-160,205-222) -- Continue alone is ignored, patterns are kept with their
    trailing '$' / '.*' removed, a pattern reduced to nothing is dropped."""
    def reg(p):
        if not p:
            return None
        if p.endswith("$"):
            p = p[:-1]
        while p.endswith(".*"):
            p = p[:-2]
        return Regex(p) if p else None
    start, end = reg(cfg.get("StartPattern")), reg(cfg.get("EndPattern"))
    return start, end


def split_lines(buf, split_char: int = 10):
    a = _as_u8(buf)
    L = lib()
    n = int(L.orc_split_lines(_ptr(a), a.size, split_char, None, None, 0)) if a.size else 0
    off = np.zeros(n, np.uint32)
    ln = np.zeros(n, np.uint32)
    if n:
        L.orc_split_lines(_ptr(a), a.size, split_char, _ptr(off), _ptr(ln), n)
    return off, ln


def multiline_split(buf, start: Optional[Regex], cont: Optional[Regex], end: Optional[Regex], discard: bool):
    """-> (off, len, flags[bit0 isLast, bit1 matched], counters[matched_events, input_lines, unmatch_lines])"""
    a = _as_u8(buf)
    L = lib()
    ctr = np.zeros(3, np.uint64)
    s = start._m if start else None
    c = cont._m if cont else None
    e = end._m if end else None
    if a.size == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint8), ctr
    n = int(L.orc_multiline_split(_ptr(a), a.size, s, c, e, int(discard), None, None, None, 0, None))
    off = np.zeros(n, np.uint32)
    ln = np.zeros(n, np.uint32)
    fl = np.zeros(n, np.uint8)
    L.orc_multiline_split(_ptr(a), a.size, s, c, e, int(discard), _ptr(off), _ptr(ln), _ptr(fl), n, _ptr(ctr))
    return off, ln, fl, ctr


def regex_parse_batch(rx: Regex, base, ev_off, ev_len, nkeys: int, matcher=None):
    a = _as_u8(base)
    ev_off = np.ascontiguousarray(ev_off, np.uint32)
    ev_len = np.ascontiguousarray(ev_len, np.uint32)
    n = ev_off.size
    G = rx.ngroups
    status = np.zeros(n, np.uint8)
    co = np.zeros((n, max(G, 1)), np.uint32)
    cl = np.zeros((n, max(G, 1)), np.uint32)
    if G == 0:
        # degenerate: no groups; keep the row stride of 1 but pass G=0 semantics by looping
        m = matcher or rx._m
        for i in range(n):
            ok = lib().orc_regex_full_match(m, C.c_void_p(a.ctypes.data + int(ev_off[i])), int(ev_len[i]), None, None)
            status[i] = 0 if (ok and 1 > nkeys) else (2 if ok else 1)
        return status, co[:, :0], cl[:, :0]
    lib().orc_regex_parse_batch(matcher or rx._m, _ptr(a), _ptr(ev_off), _ptr(ev_len), n, nkeys, _ptr(status),
                                _ptr(co), _ptr(cl))
    return status, co, cl


def regex_match_batch(rx: Regex, base, ev_off, ev_len):
    a = _as_u8(base)
    ev_off = np.ascontiguousarray(ev_off, np.uint32)
    ev_len = np.ascontiguousarray(ev_len, np.uint32)
    out = np.zeros(ev_off.size, np.uint8)
    lib().orc_regex_match_batch(rx._m, _ptr(a), _ptr(ev_off), _ptr(ev_len), ev_off.size, _ptr(out))
    return out.astype(bool)


def delim_parse_batch(base, ev_off, ev_len, sep: bytes, quote: int, nkeys: int, extend: bool, allow_short: bool,
                      max_fields: int):
    a = _as_u8(base)
    ev_off = np.ascontiguousarray(ev_off, np.uint32)
    ev_len = np.ascontiguousarray(ev_len, np.uint32)
    n = ev_off.size
    mode_quote = 1 if (len(sep) == 1 and quote != sep[0]) else 0
    status = np.zeros(n, np.uint8)
    nf = np.zeros(n, np.uint32)
    fo = np.zeros((n, max_fields), np.uint32)
    fl = np.zeros((n, max_fields), np.uint32)
    fd = np.zeros((n, max_fields), np.uint32)
    sp = np.frombuffer(sep, np.uint8)
    lib().orc_delim_parse_batch(_ptr(a), _ptr(ev_off), _ptr(ev_len), n, _ptr(sp), len(sep), quote, mode_quote, nkeys,
                                int(extend), int(allow_short), max_fields, _ptr(status), _ptr(nf), _ptr(fo), _ptr(fl),
                                _ptr(fd))
    return status, nf, fo, fl, fd


# ----------------------------------------------------------------------------- event model
LOG, METRIC, SPAN, RAW = 1, 2, 3, 4  # PipelineEvent::Type (core/models/PipelineEvent.h)

# metadata keys that FromJson understands (PipelineEventGroup.cpp:396-407); everything else -> UNKNOWN
_KNOWN_META = {"log.file.path_resolved", "source.id", "has.part.log", "log.file.offset"}
META_LOG_FILE_OFFSET_KEY = "log.file.offset"  # EventGroupMetaKey::LOG_FILE_OFFSET_KEY (PipelineEventGroup.cpp:348,375)


class Event:
    __slots__ = ("type", "contents", "timestamp", "ns", "pos", "raw", "extra")

    def __init__(self, type_=LOG):
        self.type = type_
        self.contents = []  # list of [key(bytes), val(bytes), live]
        self.timestamp = 0
        self.ns = None
        self.pos = (0, 0)
        self.raw = b""
        self.extra = None  # opaque JSON for metric/span events (passed through)

    # --- LogEvent.cpp:50-106
    def _find(self, key):
        for it in reversed(self.contents):
            if it[0] == key and it[2]:
                return it
        return None

    def has(self, key):
        return self._find(key) is not None

    def get(self, key):
        it = self._find(key)
        return it[1] if it else b""

    def set(self, key, val):
        it = self._find(key)
        if it:
            it[0], it[1] = key, val
        else:
            self.contents.append([key, val, True])

    def delete(self, key):
        it = self._find(key)
        if it:
            it[2] = False

    def live(self):
        return [(k, v) for k, v, l in self.contents if l]

    def size(self):
        return sum(1 for c in self.contents if c[2])


class Group:
    def __init__(self):
        self.events = []
        self.metadata = {}
        self.tags = {}

    @staticmethod
    def from_json(s):
        root = json.loads(s) if isinstance(s, (str, bytes)) else s
        g = Group()
        if root is None:
            return g
        for k in sorted(root.get("metadata", {})):
            g.metadata[k if k in _KNOWN_META or k == META_LOG_FILE_OFFSET_KEY else "unknown"] = root["metadata"][k]
        for k in sorted(root.get("tags", {})):
            g.tags[k] = root["tags"][k]
        for ev in root.get("events", []):
            t = int(ev["type"])
            e = Event(t if t in (LOG, METRIC, SPAN) else RAW)
            e.timestamp = int(ev.get("timestamp", 0))
            if "timestampNanosecond" in ev:
                e.ns = int(ev["timestampNanosecond"])
            if e.type == LOG:
                if "fileOffset" in ev and "rawSize" in ev:
                    e.pos = (int(ev["fileOffset"]), int(ev["rawSize"]))
                for k in sorted(ev.get("contents", {}), key=lambda x: x.encode("utf-8")):
                    e.set(k.encode("utf-8"), ev["contents"][k].encode("utf-8"))
            elif e.type == RAW:
                e.raw = ev.get("content", "").encode("utf-8")
            else:
                e.extra = ev
            g.events.append(e)
        return g

    def to_json(self, enable_event_meta=False):
        """Same shape as PipelineEventGroup::ToJson; an empty group serialises to None ("null")."""
        root = {}
        if self.metadata:
            root["metadata"] = dict(self.metadata)
        if self.tags:
            root["tags"] = dict(self.tags)
        if self.events:
            evs = []
            for e in self.events:
                if e.type in (METRIC, SPAN):
                    evs.append(e.extra)
                    continue
                d = {"type": e.type, "timestamp": e.timestamp}
                if e.ns is not None:
                    d["timestampNanosecond"] = e.ns
                if e.type == RAW:
                    d["content"] = e.raw.decode("utf-8", "surrogateescape")
                else:
                    if enable_event_meta:
                        d["fileOffset"], d["rawSize"] = e.pos
                    live = e.live()
                    if live:
                        d["contents"] = {k.decode("utf-8", "surrogateescape"): v.decode("utf-8", "surrogateescape")
                                         for k, v in live}
                evs.append(d)
            root["events"] = evs
        return root or None


# ----------------------------------------------------------------------------- processors
def _b(x):
    return x.encode("utf-8") if isinstance(x, str) else x


def _last_segment(cfg: dict, dotted: str, default=None):
    """ParamExtractor resolves dotted names to their last segment (core/common/ParamExtractor.cpp:23-29)."""
    key = dotted.split(".")[-1]
    return cfg.get(key, default)


class CommonParserOptions:
    """CommonParserOptions.cpp:28-117"""

    legacy_raw_key = b"__raw_log__"

    def __init__(self, cfg):
        self.keep_fail = bool(cfg.get("KeepingSourceWhenParseFail", False)) if isinstance(
            cfg.get("KeepingSourceWhenParseFail", False), bool) else False
        self.keep_ok = bool(cfg.get("KeepingSourceWhenParseSucceed", False)) if isinstance(
            cfg.get("KeepingSourceWhenParseSucceed", False), bool) else False
        r = cfg.get("RenamedSourceKey", "")
        self.renamed = _b(r) if isinstance(r, str) and r else _b(cfg["SourceKey"])
        self.coping_raw = bool(cfg.get("CopingRawLog", False)) if isinstance(cfg.get("CopingRawLog", False),
                                                                              bool) else False

    def should_add_source(self, ok):
        return (ok and self.keep_ok) or ((not ok) and self.keep_fail)

    def should_add_legacy_raw(self, ok):
        return (not ok) and self.keep_fail and self.coping_raw

    def should_erase(self, ok, ev: Event, metadata):
        if (not ok) and (not self.keep_fail):
            live = ev.live()
            if not live:
                return True
            offkey = metadata.get(META_LOG_FILE_OFFSET_KEY)
            if len(live) == 1 and offkey is not None and live[0][0] == _b(offkey):
                return True
            if len(live) == 2 and ev.has(b"_time_") and ev.has(b"_source_"):
                return True
        return False


def _add_log(ev: Event, key, val, overwritten=True):
    if (not overwritten) and ev.has(key):
        return
    ev.set(key, val)


class ProcessorSplitLogStringNative:
    name = "processor_split_string_native"

    def __init__(self, cfg):
        self.source_key = _b(cfg.get("SourceKey", "content"))
        sc = cfg.get("SplitChar", 10)
        self.split_char = (int(sc) & 0xFF) if isinstance(sc, int) and not isinstance(sc, bool) else 10
        self.raw = bool(cfg.get("EnableRawContent", False))

    def process(self, g: Group):
        if not g.events:
            return
        new = []
        for e in g.events:
            if e.type != LOG or e.size() != 1 or not e.has(self.source_key):
                new.append(e)
                continue
            val = e.get(self.source_key)
            off, ln = split_lines(val, self.split_char)
            for o, l in zip(off.tolist(), ln.tolist()):
                content = val[o:o + l]
                if self.raw:
                    t = Event(RAW)
                    t.raw = content
                    t.timestamp, t.ns = e.timestamp, e.ns
                else:
                    t = Event(LOG)
                    t.set(self.source_key, content)
                    t.timestamp, t.ns = e.timestamp, e.ns
                    offset = e.pos[0] + o
                    length = (e.pos[1] - o) if (o + l == len(val)) else l + 1
                    t.pos = (offset, length)
                    if META_LOG_FILE_OFFSET_KEY in g.metadata:
                        t.set(_b(g.metadata[META_LOG_FILE_OFFSET_KEY]), str(offset).encode())
                new.append(t)
        g.events = new


class MultilineOptions:
    """MultilineOptions.cpp:22-222 (custom mode only; JSON mode selects the plain splitter upstream)."""

    def __init__(self, cfg):
        self.ok = True
        self.start = self.cont = self.end = ""
        has = {}
        for name in ("StartPattern", "ContinuePattern", "EndPattern"):
            pat = _last_segment(cfg, "Multiline." + name, "")
            if not isinstance(pat, str):
                pat = ""
            valid, compiled = self._parse_regex(pat)
            if not valid:
                pat, compiled = "", False
            has[name] = compiled
            setattr(self, {"StartPattern": "start", "ContinuePattern": "cont", "EndPattern": "end"}[name], pat)
        # NB: the combination rules (:125-160) reset only the *RegPtr members; the processor keys off the
        # pattern STRINGS (ProcessorSplitMultilineLogStringNative.cpp:68-80), so all three strings stay.
        self.is_multiline = has["StartPattern"] or has["EndPattern"]
        t = _last_segment(cfg, "Multiline.UnmatchedContentTreatment", "single_line")
        self.discard = (t == "discard")
        # The *RegPtr members (used by ProcessorMergeMultilineLogNative): compiled from the TRIMMED pattern
        # (trailing '$' and '.*' removed, ParseRegex :196-215), null when the trimmed pattern is empty, and the
        # continue pattern is dropped when it is the only one or when all three are given (:125-160).
        self.start_reg = self._trim(self.start) if has["StartPattern"] else None
        self.cont_reg = self._trim(self.cont) if has["ContinuePattern"] else None
        self.end_reg = self._trim(self.end) if has["EndPattern"] else None
        if self.start_reg is None and self.end_reg is None and self.cont_reg is not None:
            self.cont_reg = None
        elif self.start_reg is not None and self.cont_reg is not None and self.end_reg is not None:
            self.cont_reg = None
        self.ignore_warning = bool(cfg.get("IgnoringUnmatchWarning", False))

    @staticmethod
    def _trim(pattern):
        p = pattern
        if p.endswith("$"):
            p = p[:-1]
        while p.endswith(".*"):
            p = p[:-2]
        return p

    @staticmethod
    def _parse_regex(pattern):
        p = pattern
        if p.endswith("$"):
            p = p[:-1]
        while p.endswith(".*"):
            p = p[:-2]
        if not p:
            return True, False
        try:
            Regex(p)
        except ValueError:
            return False, False
        return True, True


class ProcessorSplitMultilineLogStringNative:
    name = "processor_split_multiline_log_string_native"

    def __init__(self, cfg):
        self.source_key = _b(cfg.get("SourceKey", "content"))
        self.opts = MultilineOptions(cfg)
        self.raw = bool(cfg.get("EnableRawContent", False))
        self.start = Regex(self.opts.start) if self.opts.start else None
        self.cont = Regex(self.opts.cont) if self.opts.cont else None
        self.end = Regex(self.opts.end) if self.opts.end else None
        self.counters = {"matched_events": 0, "matched_lines": 0, "unmatched_lines": 0}

    def process(self, g: Group):
        if not g.events:
            return
        new = []
        in_lines = un_lines = 0
        for e in g.events:
            if e.type != LOG or e.size() != 1 or not e.has(self.source_key):
                new.append(e)
                continue
            val = e.get(self.source_key)
            off, ln, fl, ctr = multiline_split(val, self.start, self.cont, self.end, self.opts.discard)
            self.counters["matched_events"] += int(ctr[0])
            in_lines += int(ctr[1])
            un_lines += int(ctr[2])
            for o, l, f in zip(off.tolist(), ln.tolist(), fl.tolist()):
                content = val[o:o + l]
                if self.raw:
                    t = Event(RAW)
                    t.raw = content
                    t.timestamp, t.ns = e.timestamp, e.ns
                else:
                    t = Event(LOG)
                    t.set(self.source_key, content)
                    t.timestamp, t.ns = e.timestamp, e.ns
                    offset = e.pos[0] + o
                    length = (e.pos[1] - o) if (f & 1) else l + 1
                    t.pos = (offset, length)
                    if META_LOG_FILE_OFFSET_KEY in g.metadata:
                        t.set(_b(g.metadata[META_LOG_FILE_OFFSET_KEY]), str(offset).encode())
                new.append(t)
        self.counters["matched_lines"] += in_lines - un_lines
        self.counters["unmatched_lines"] += un_lines
        g.events = new


def _keys_param(cfg):
    keys = cfg.get("Keys")
    if not isinstance(keys, list):
        raise ValueError("mandatory list param Keys")
    return [_b(k) for k in keys]


class ProcessorParseRegexNative:
    name = "processor_parse_regex_native"

    def __init__(self, cfg):
        self.source_key = _b(cfg["SourceKey"])
        self.regex_str = cfg["Regex"]
        self.rx = Regex(self.regex_str)
        self.whole_line = self.regex_str == "(.*)"
        self.keys = _keys_param(cfg)
        if len(self.keys) == 1 and b"," in self.keys[0]:
            self.keys = [k for k in self.keys[0].split(b",")]
        self.source_overwritten = self.source_key in self.keys
        self.common = CommonParserOptions(cfg)
        self.counters = {"discarded": 0, "out_failed": 0, "out_key_not_found": 0, "out_successful": 0}

    def process(self, g: Group):
        if not g.events:
            return
        out = []
        for e in g.events:
            if self._event(e, g.metadata):
                out.append(e)
        g.events = out

    def _event(self, e: Event, metadata):
        if e.type != LOG:
            self.counters["out_failed"] += 1
            return True
        if not e.has(self.source_key):
            self.counters["out_key_not_found"] += 1
            return True
        raw = e.get(self.source_key)
        if self.whole_line:
            _add_log(e, self.keys[0] if self.keys else b"content", raw)
            ok = True
        else:
            caps = self.rx.full_match(raw)
            if caps is None:
                self.counters["out_failed"] += 1
                ok = False
            elif self.rx.ngroups + 1 <= len(self.keys):
                ok = False
            else:
                ok = True
                for i, k in enumerate(self.keys):
                    o, l = caps[i]
                    _add_log(e, k, raw[o:o + l])
        if (not ok) or (not self.source_overwritten):
            e.delete(self.source_key)
        if self.common.should_add_source(ok):
            _add_log(e, self.common.renamed, raw, False)
        if self.common.should_add_legacy_raw(ok):
            _add_log(e, self.common.legacy_raw_key, raw, False)
        if self.common.should_erase(ok, e, metadata):
            self.counters["discarded"] += 1
            return False
        self.counters["out_successful"] += 1
        return True


class ProcessorParseDelimiterNative:
    name = "processor_parse_delimiter_native"

    def __init__(self, cfg):
        self.source_key = _b(cfg["SourceKey"])
        sep = cfg["Separator"]
        if len(_b(sep)) > 4:
            raise ValueError("Separator has more than 4 chars")
        if sep == "\\t":
            sep = "\t"
        self.sep = _b(sep)
        q = cfg.get("Quote", "")
        self.quote = ord('"')
        if len(self.sep) == 1:
            if isinstance(q, str) and len(_b(q)) > 1:
                raise ValueError("Quote is not a single char")
            if isinstance(q, str) and q:
                self.quote = _b(q)[0]
        # multi-char separator: a configured Quote is ignored with a warning (:97-107)
        self.keys = _keys_param(cfg)
        self.source_overwritten = self.source_key in self.keys
        a = cfg.get("AllowingShortenedFields", True)
        self.allow_short = a if isinstance(a, bool) else True
        t = cfg.get("OverflowedFieldsTreatment", "extend")
        self.overflow = t if t in ("keep", "discard") else "extend"
        self.partial = self.overflow == "discard"
        self.common = CommonParserOptions(cfg)
        self.counters = {"discarded": 0, "out_failed": 0, "out_key_not_found": 0, "out_successful": 0}

    def process(self, g: Group):
        if not g.events:
            return
        g.events = [e for e in g.events if self._event(e, g.metadata)]

    def _event(self, e: Event, metadata):
        L = lib()
        if e.type != LOG:
            self.counters["out_failed"] += 1
            return True
        if not e.has(self.source_key):
            self.counters["out_key_not_found"] += 1
            return True
        buf = e.get(self.source_key)
        a = _as_u8(buf)
        b = C.c_int32(0)
        en = C.c_int32(0)
        if a.size == 0 or not L.orc_delim_trim(_ptr(a), a.size, C.byref(b), C.byref(en)):
            self.counters["out_failed"] += 1
            return True
        beg, end = b.value, en.value
        extend = self.overflow == "extend"
        use_quote = len(self.sep) == 1 and self.quote != self.sep[0]
        ok = False
        cols = []
        if self.keys:
            cap = a.size + 2
            fo = np.zeros(cap, np.uint32)
            fl = np.zeros(cap, np.uint32)
            fd = np.zeros(cap, np.uint32)
            if use_quote:
                k = int(L.orc_delim_fsm(_ptr(a), beg, end, self.sep[0], self.quote, _ptr(fo), _ptr(fl), _ptr(fd),
                                        cap))
                ok = k >= 0
                if ok:
                    for j in range(k):
                        if fd[j]:
                            dst = np.zeros(int(fl[j]) + 1, np.uint8)
                            w = L.orc_delim_unquote(_ptr(a), int(fo[j]), int(fl[j]), self.quote, _ptr(dst))
                            assert w == int(fl[j]) - int(fd[j])
                            cols.append(bytes(dst[:w]))
                        else:
                            cols.append(buf[int(fo[j]):int(fo[j]) + int(fl[j])])
                    if (not extend) and len(cols) > len(self.keys):
                        extra = b"".join(self.sep[:1] + c for c in cols[len(self.keys):])
                        cols = cols[:len(self.keys)] + [extra]
            else:
                sp = np.frombuffer(self.sep, np.uint8)
                k = int(L.orc_delim_split(_ptr(a), beg, end, _ptr(sp), len(self.sep), len(self.keys), int(extend),
                                          _ptr(fo), _ptr(fl), cap))
                ok = k > 0
                cols = [buf[int(fo[j]):int(fo[j]) + int(fl[j])] for j in range(max(k, 0))]
            if ok:
                if len(cols) <= 0 or ((not self.allow_short) and len(cols) < len(self.keys)):
                    ok = False
        if ok:
            for idx, c in enumerate(cols):
                if idx < len(self.keys):
                    if self.partial and self.keys[idx] == b"_":
                        continue
                    _add_log(e, self.keys[idx], c)
                else:
                    if self.partial:
                        continue
                    _add_log(e, b"__column%d__" % idx, c)
            self.counters["out_successful"] += 1
        else:
            self.counters["out_failed"] += 1
        if (not ok) or (not self.source_overwritten):
            e.delete(self.source_key)
        if self.common.should_add_source(ok):
            _add_log(e, self.common.renamed, buf, False)
        if self.common.should_add_legacy_raw(ok):
            _add_log(e, self.common.legacy_raw_key, buf, False)
        if self.common.should_erase(ok, e, metadata):
            self.counters["discarded"] += 1
            return False
        return True


def _none_utf8(data: bytes, modify: bool):
    """ProcessorFilterNative::noneUtf8 (ProcessorFilterNative.cpp:297-378): returns (has_invalid, sanitised) where
    every byte that starts an invalid sequence is replaced by a blank (only that byte; scanning resumes after it)."""
    b = bytearray(data)
    i, n, bad = 0, len(b), False

    def cont(k):
        return k < n and (b[k] & 0xC0) == 0x80

    while i < n:
        c = b[i]
        step = 1
        inv = False
        if c & 0x80 == 0:
            pass
        elif c & 0xE0 == 0xC0:
            if i + 1 >= n or not cont(i + 1):
                inv = True
            else:
                u = ((c & 0x1F) << 6) | (b[i + 1] & 0x3F)
                inv = not (0x80 <= u <= 0x7FF)
                step = 2
        elif c & 0xF0 == 0xE0:
            if i + 2 >= n or not cont(i + 1) or not cont(i + 2):
                inv = True
            else:
                u = (((c & 0x0F) << 12) | ((b[i + 1] & 0x3F) << 6) | (b[i + 2] & 0x3F)) & 0xFFFF
                inv = not (u >= 0x800)
                step = 3
        elif c & 0xF8 == 0xF0:
            if i + 3 >= n or not cont(i + 1) or not cont(i + 2) or not cont(i + 3):
                inv = True
            else:
                u = ((c & 0x07) << 18) | ((b[i + 1] & 0x3F) << 12) | ((b[i + 2] & 0x3F) << 6) | (b[i + 3] & 0x3F)
                inv = not (0x10000 <= u <= 0x10FFFF)
                step = 4
        else:
            inv = True
        if inv:
            if not modify:
                return True, data
            b[i] = 0x20
            bad = True
            i += 1
        else:
            i += step
    return bad, bytes(b)


class ProcessorFilterNative:
    """core/plugin/processor/ProcessorFilterNative.cpp:30-275,380-488 -- the first "next" row of SURVEY.md 8(f)."""
    name = "processor_filter_regex_native"

    def __init__(self, cfg):
        self.mode = "bypass"
        self.exp = None
        self.rule = None
        ce = cfg.get("ConditionExp")
        if ce is not None:
            if not isinstance(ce, dict):
                raise ValueError("object param ConditionExp is not of type object")
            self.exp = self._parse(ce)
            if self.exp is None:
                raise ValueError("object param ConditionExp is not valid")
            self.mode = "expression"
        if self.mode == "bypass":
            keys = cfg.get("FilterKey") or []
            regs = cfg.get("FilterRegex") or []
            if len(keys) != len(regs):
                raise ValueError("param FilterKey and FilterRegex does not have the same size")
            if keys:
                self.rule = [(_b(k), Regex(r)) for k, r in zip(keys, regs)]
                self.mode = "rule"
        if self.mode == "bypass":
            inc = cfg.get("Include") or {}
            if inc:
                self.rule = [(_b(k), Regex(inc[k])) for k in sorted(inc)]
                self.mode = "rule"
        d = cfg.get("DiscardingNonUTF8", False)
        self.discard_non_utf8 = d if isinstance(d, bool) else False

    def _parse(self, v):
        if not isinstance(v, dict):
            return None
        if isinstance(v.get("operator"), str) and isinstance(v.get("operands"), list):
            op = v["operator"].lower()
            ops = v["operands"]
            if op == "not" and len(ops) == 1:
                c = self._parse(ops[0])
                return ("not", c) if c else None
            if op in ("and", "or") and len(ops) == 2:
                l, r = self._parse(ops[0]), self._parse(ops[1])
                return (op, l, r) if l and r else None
            return None
        if (isinstance(v.get("key"), str) and isinstance(v.get("exp"), str)) or not isinstance(v.get("type"), str):
            t = v.get("type", "")
            if not isinstance(t, str) or t.lower() != "regex":
                return None
            return ("regex", _b(v.get("key", "")), Regex(v.get("exp", "")))
        return None

    def _eval(self, node, e: Event):
        if node[0] == "regex":
            if not e.has(node[1]):
                return False
            return node[2].full_match(e.get(node[1])) is not None
        if node[0] == "not":
            return not self._eval(node[1], e)
        if node[0] == "and":
            return self._eval(node[1], e) and self._eval(node[2], e)
        return self._eval(node[1], e) or self._eval(node[2], e)

    def process(self, g: Group):
        if not g.events:
            return
        g.events = [e for e in g.events if self._event(e)]

    def _event(self, e: Event):
        if e.type != LOG:
            return True
        res = True
        if self.mode == "expression":
            res = e.size() > 0 and self._eval(self.exp, e)
        elif self.mode == "rule":
            res = e.size() > 0 and all(e.has(k) and rx.full_match(e.get(k)) is not None for k, rx in self.rule)
        if res and self.discard_non_utf8:
            renamed = []
            for c in e.contents:
                if not c[2]:
                    continue
                bad, fixed = _none_utf8(c[1], True)
                if bad:
                    c[1] = fixed
                badk, fixedk = _none_utf8(c[0], True)
                if badk:
                    renamed.append((fixedk, c[1]))
                    c[2] = False
            for k, v in renamed:
                e.set(k, v)
        return res


class ProcessorMergeMultilineLogNative:
    """SURVEY.md 8(f) rank 2.  core/plugin/processor/inner/ProcessorMergeMultilineLogNative.cpp:33-420:
    merges already-split LogEvents of a group back into records, by the docker partial-log flag ("P" content,
    :116-159) or by the start / continue / end patterns (anchored prefix probes, :161-318)."""
    name = "processor_merge_multiline_log_native"
    PART_LOG_FLAG = b"P"

    def __init__(self, cfg):
        self.source_key = _b(cfg.get("SourceKey", "content"))
        mt = cfg.get("MergeType")
        if mt == "flag":
            self.by_flag = True
            self.opts = None
        elif mt == "regex":
            self.by_flag = False
            self.opts = MultilineOptions(cfg)
            self.start = Regex(self.opts.start_reg) if self.opts.start_reg is not None else None
            self.cont = Regex(self.opts.cont_reg) if self.opts.cont_reg is not None else None
            self.end = Regex(self.opts.end_reg) if self.opts.end_reg is not None else None
        else:
            raise ValueError("MergeType")  # Init returns false (:52-74)
        self.counters = {"merged_events": 0, "unmatched_events": 0}

    def process(self, g: Group):
        if not g.events:
            return
        if not self.by_flag:
            self._by_regex(g)
        elif "has.part.log" in g.metadata:
            self._by_flag(g)
            del g.metadata["has.part.log"]

    # MergeEvents (:320-346): the target keeps its other contents; values are joined in place
    def _merge(self, evs, line_break):
        if not evs:
            return
        self.counters["merged_events"] += len(evs)
        if len(evs) > 1:
            sep = b"\n" if line_break else b""
            evs[0].set(self.source_key, sep.join(e.get(self.source_key) for e in evs))
        evs.clear()

    # HandleUnmatchLogs (:348-385)
    def _unmatched(self, src, out, begin, end):
        self.counters["unmatched_events"] += end - begin + 1
        if self.opts.discard:
            return
        out.extend(src[begin:end + 1])

    def _by_flag(self, g: Group):
        src, out, evs = g.events, [], []
        partial, begin = False, 0
        for cur, e in enumerate(src):
            if e.type != LOG:
                if not evs:
                    begin = cur
                out.extend(src[begin:])
                g.events = out
                return
            if e.size() == 0:
                continue
            evs.append(e)
            if partial:
                if not e.has(self.PART_LOG_FLAG):
                    self._merge(evs, False)
                    out.append(src[begin])
                    begin = cur + 1
                    partial = False
            elif e.has(self.PART_LOG_FLAG):
                e.delete(self.PART_LOG_FLAG)
                partial = True
            else:
                self._merge(evs, False)
                out.append(src[begin])
                begin = cur + 1
        if partial:
            self._merge(evs, False)
            out.append(src[begin])
        g.events = out

    def _by_regex(self, g: Group):
        S, C, E = self.start, self.cont, self.end
        src, out, evs = g.events, [], []
        begin = 0
        partial = S is None and C is None and E is not None
        for cur, e in enumerate(src):
            if e.type != LOG or (e.size() != 0 and not e.has(self.source_key)):
                if not evs:
                    begin = cur
                out.extend(src[begin:])
                g.events = out
                return
            if e.size() == 0:
                continue
            val = e.get(self.source_key)
            if not partial:
                first = S if S is not None else C
                if first.prefix_match(val):
                    evs.append(e)
                    begin = cur
                    partial = True
                elif E is not None and S is None and C is not None and E.prefix_match(val):
                    begin = cur
                    self.counters["merged_events"] += 1
                    out.append(src[begin])
                else:
                    self._unmatched(src, out, cur, cur)
                continue
            if C is not None and C.prefix_match(val):
                evs.append(e)
                continue
            if E is not None:
                evs.append(e)
                if C is not None:
                    if E.prefix_match(val):
                        self._merge(evs, True)
                        out.append(src[begin])
                    else:
                        self._unmatched(src, out, begin, cur)
                        evs.clear()
                    partial = False
                elif E.prefix_match(val):
                    self._merge(evs, True)
                    out.append(src[begin])
                    if S is not None:
                        partial = False
                    else:
                        begin = cur + 1
            elif C is None:
                if not S.prefix_match(val):
                    evs.append(e)
                else:
                    self._merge(evs, True)
                    out.append(src[begin])
                    begin = cur
                    evs.append(e)
            else:
                self._merge(evs, True)
                out.append(src[begin])
                if not S.prefix_match(val):
                    self._unmatched(src, out, cur, cur)
                    partial = False
                else:
                    begin = cur
                    evs.append(e)
        if partial and begin < len(src):
            if E is None:
                self._merge(evs, True)
                out.append(src[begin])
            else:
                self._unmatched(src, out, begin, len(src) - 1)
        g.events = out


# ----------------------------------------------------------------------------- SLS wire format (next row, rank 4)
# Restates core/protobuf/sls/LogGroupSerializer.cpp:33-143,232-262 (hand-rolled protobuf writer) and the LOG-event
# path of SLSEventGroupSerializer::Serialize (core/collection_pipeline/serializer/SLSSerializer.cpp:162-269,377-395).
SLS_MIN_LOG_TIME = 1 << 28  # AddLogTime clamps so that the varint always takes 5 bytes (:108-117)
SLS_MAX_LOG_GROUP_SIZE = 10 * 1024 * 1024  # max_send_log_group_size (FlusherSLS.cpp:61)
_SLS_TOPIC, _SLS_SOURCE, _SLS_UUID = b"__topic__", b"__source__", b"__machine_uuid__"


def _sls_varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _sls_string(b: bytes) -> bytes:
    return _sls_varint(len(b)) + b


def _sls_pair(tag: int, key: bytes, val: bytes) -> bytes:
    """Contents (field 2 of Log) / LogTags (field 6 of LogGroup): nested {Key = 1, Value = 2}"""
    inner = b"\x0a" + _sls_string(key) + b"\x12" + _sls_string(val)
    return bytes([tag]) + _sls_varint(len(inner)) + inner


def sls_serialize_logs(events, enable_ns: bool):
    """The `Logs` fields (field 1) of the LogGroup for a list of (time, ns or None, [(key, value), ...]); events
    without contents are skipped (LogEvent::Empty).  Returns (bytes, offsets of each emitted record)."""
    out = bytearray()
    offs = []
    for t, ns, contents in events:
        if not contents:
            continue
        body = b"\x08" + _sls_varint(max(int(t) & 0xFFFFFFFF, SLS_MIN_LOG_TIME))
        for k, v in contents:
            body += _sls_pair(0x12, k, v)
        if enable_ns and ns is not None:
            body += b"\x25" + int(ns).to_bytes(4, "little")
        offs.append(len(out))
        out += b"\x0a" + _sls_varint(len(body)) + body
    return bytes(out), offs


def sls_serialize_group(g: "Group", enable_ns: bool = False):
    """SLSEventGroupSerializer::Serialize for a group of LOG events: (bytes, None) or (None, error message)."""
    if not g.events:
        return None, "empty event group"
    evs = [(e.timestamp, e.ns, e.live()) for e in g.events]
    logs, _ = sls_serialize_logs(evs, enable_ns)
    if not logs:
        return None, "all empty logs"
    tail = bytearray()
    for k in sorted(g.tags, key=lambda x: _b(x)):  # SizedMap wraps a std::map: key order
        kb, vb = _b(k), _b(g.tags[k])
        if kb == _SLS_TOPIC:
            tail += b"\x1a" + _sls_string(vb)
        elif kb == _SLS_SOURCE:
            tail += b"\x22" + _sls_string(vb)
        elif kb == _SLS_UUID:
            tail += b"\x2a" + _sls_string(vb)
        else:
            tail += _sls_pair(0x32, kb, vb)
    if len(logs) + len(tail) > SLS_MAX_LOG_GROUP_SIZE:
        return None, "log group exceeds size limit"
    return logs + bytes(tail), None


PROCESSORS = {
    p.name: p
    for p in (ProcessorSplitLogStringNative, ProcessorSplitMultilineLogStringNative, ProcessorParseRegexNative,
              ProcessorParseDelimiterNative, ProcessorFilterNative, ProcessorMergeMultilineLogNative)
}
