"""The oracle's delimiter FSM restatement checked against the REFERENCE's own translation unit
(core/parser/DelimiterModeFsmParser.cpp compiled in place by oracle/build_ref.sh into oracle/_ref/)."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libref_delim.so")

pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref not built (needs /root/reference)")


def ref_fsm(lib, line: bytes, sep: int, quote: int):
    out = np.zeros(len(line) * 5 + 64, np.uint8)
    used = C.c_uint64(0)
    k = lib.ref_delim_fsm(line, 0, len(line), C.c_char(bytes([sep])), C.c_char(bytes([quote])),
                          out.ctypes.data_as(C.c_void_p), out.size, C.byref(used))
    if k < 0:
        return None
    cols, at = [], 0
    raw = out.tobytes()
    for _ in range(k):
        l = int.from_bytes(raw[at:at + 4], "little")
        cols.append(raw[at + 4:at + 4 + l])
        at += 4 + l
    return cols


def oracle_fsm(line: bytes, sep: int, quote: int):
    L = orc.lib()
    a = np.frombuffer(line, np.uint8) if line else np.zeros(0, np.uint8)
    cap = len(line) + 2
    fo, fl, fd = (np.zeros(cap, np.uint32) for _ in range(3))
    k = int(L.orc_delim_fsm(a.ctypes.data_as(C.c_void_p) if a.size else None, 0, len(line), sep, quote,
                            fo.ctypes.data_as(C.c_void_p), fl.ctypes.data_as(C.c_void_p),
                            fd.ctypes.data_as(C.c_void_p), cap))
    if k < 0:
        return None
    cols = []
    for j in range(k):
        if fd[j]:
            dst = np.zeros(int(fl[j]) + 1, np.uint8)
            w = L.orc_delim_unquote(a.ctypes.data_as(C.c_void_p), int(fo[j]), int(fl[j]), quote,
                                    dst.ctypes.data_as(C.c_void_p))
            cols.append(bytes(dst[:w]))
        else:
            cols.append(line[int(fo[j]):int(fo[j]) + int(fl[j])])
    return cols


def test_fsm_restatement_equals_reference_translation_unit():
    lib = C.CDLL(SO)
    lib.ref_delim_fsm.restype = C.c_int64
    lib.ref_delim_fsm.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_char, C.c_char, C.c_void_p, C.c_uint64,
                                  C.POINTER(C.c_uint64)]
    rng = random.Random(20260922)
    n = fails = 0
    for sep, quote in ((44, 39), (44, 34), (124, 39), (9, 34)):
        alpha = "ab1 " + chr(sep) * 3 + chr(quote) * 3 + "x,"
        for _ in range(20000):
            line = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 24))).encode()
            r, o = ref_fsm(lib, line, sep, quote), oracle_fsm(line, sep, quote)
            assert r == o, (line, sep, quote, r, o)
            n += 1
            fails += r is None
    assert n == 80000 and 1000 < fails < n - 1000  # both outcomes exercised
    # the reference's own TestProcessDoubleQuote inputs (ProcessorParseDelimiterNativeUnittest.cpp:1841-1904)
    for url in ("''PutData?Category=YunOsAccountOpLog", "PutData?Category=YunOs''AccountOpLog",
                "PutData?Category=YunOsAccountOpLog''", "''PutData?Category=YunOsAccountOpLog'",
                "'PutData?Category=Yun'Os'AccountOpLog'", "'PutData?Category=YunOs''AccountOpLog'",
                "'PutData?Category=YunOsAccountOpLog''", "'''PutData?Category=YunOs''AccountOpLog'''"):
        line = ("2013-10-31 21:03:49,POST," + url + ",0.024").encode()
        assert ref_fsm(lib, line, 44, 39) == oracle_fsm(line, 44, 39), url
