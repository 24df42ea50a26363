"""ctypes wrapper of the TEST-ONLY emulation library (tests/emul/lc_emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liblc_emul.so")
        srcs = [os.path.join(_HERE, "lc_emul.cpp")] + [
            os.path.join(_HERE, "..", "..", "loongcollector_b200", "csrc", f)
            for f in ("regex_compiler.cpp", "regex_compiler.h", "lc_exec.cuh", "lc_tables.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call([os.path.join(_HERE, "build.sh")])
        L = C.CDLL(so)
        L.emul_compile.restype = C.c_void_p
        L.emul_compile.argtypes = [C.c_char_p, C.c_uint64]
        L.emul_free.argtypes = [C.c_void_p]
        for f in ("emul_valid", "emul_supported"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.emul_error.restype = C.c_char_p
        L.emul_error.argtypes = [C.c_void_p]
        L.emul_ngroups.restype = C.c_uint32
        L.emul_ngroups.argtypes = [C.c_void_p]
        L.emul_info.argtypes = [C.c_void_p, C.c_void_p]
        L.emul_prefix_match.restype = C.c_int
        L.emul_prefix_match.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
        L.emul_full_match.restype = C.c_int
        L.emul_full_match.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.emul_full_match_fast2.restype = C.c_int
        L.emul_full_match_fast2.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.emul_full_match_tdfa.restype = C.c_int
        L.emul_full_match_tdfa.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.emul_tdfa_info.argtypes = [C.c_void_p, C.c_void_p]
        L.emul_delim_fsm.restype = C.c_int64
        L.emul_delim_fsm.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64]
        L.emul_sls_serialize_logs.restype = C.c_uint64
        L.emul_sls_serialize_logs.argtypes = [C.c_void_p] + [C.c_uint64] + [C.c_void_p] * 8 + [C.c_uint64]
        L.emul_fast2_bytes.restype = C.c_uint32
        L.emul_fast2_bytes.argtypes = [C.c_void_p]
        L.emul_fast_bytes.restype = C.c_uint32
        L.emul_fast_bytes.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class EmulRegex:
    def __init__(self, pattern):
        if isinstance(pattern, str):
            pattern = pattern.encode("utf-8")
        L = lib()
        self._h = L.emul_compile(pattern, len(pattern))
        self.valid = bool(L.emul_valid(self._h))
        self.supported = bool(L.emul_supported(self._h))
        self.error = L.emul_error(self._h).decode()
        self.ngroups = int(L.emul_ngroups(self._h))
        info = np.zeros(8, np.uint32)
        L.emul_info(self._h, info.ctypes.data_as(C.c_void_p))
        self.mode, self.nclasses, self.nw, self.npc, self.nrev, self.npre, self.blob_bytes, self.ninsts = \
            [int(x) for x in info]

    def prefix_match(self, data: bytes) -> bool:
        assert self.supported, self.error
        return bool(lib().emul_prefix_match(self._h, data, len(data)))

    def full_match(self, data: bytes):
        assert self.supported, self.error
        co = np.zeros(max(self.ngroups, 1), np.uint32)
        cl = np.zeros(max(self.ngroups, 1), np.uint32)
        ok = lib().emul_full_match(self._h, data, len(data), co.ctypes.data_as(C.c_void_p),
                                   cl.ctypes.data_as(C.c_void_p))
        if not ok:
            return None
        return [(int(co[g]), int(cl[g])) for g in range(self.ngroups)]

    def full_match_fast2(self, data: bytes, mis: int = 0):
        """stride-2 tables; returns 'n/a' when the pattern has no fast2 layout"""
        co = np.zeros(max(self.ngroups, 1), np.uint32)
        cl = np.zeros(max(self.ngroups, 1), np.uint32)
        rc = lib().emul_full_match_fast2(self._h, data, len(data), mis, co.ctypes.data_as(C.c_void_p),
                                         cl.ctypes.data_as(C.c_void_p))
        if rc < 0:
            return "n/a"
        if rc == 0:
            return None
        return [(int(co[g]), int(cl[g])) for g in range(self.ngroups)]

    def full_match_tdfa(self, data: bytes, mis: int = 0):
        """single-pass tagged DFA; returns 'n/a' when the pattern has no tdfa layout"""
        co = np.zeros(max(self.ngroups, 1), np.uint32)
        cl = np.zeros(max(self.ngroups, 1), np.uint32)
        rc = lib().emul_full_match_tdfa(self._h, data, len(data), mis, co.ctypes.data_as(C.c_void_p),
                                        cl.ctypes.data_as(C.c_void_p))
        if rc < 0:
            return "n/a"
        if rc == 0:
            return None
        return [(int(co[g]), int(cl[g])) for g in range(self.ngroups)]

    @property
    def tdfa_info(self):
        info = np.zeros(6, np.uint32)
        lib().emul_tdfa_info(self._h, info.ctypes.data_as(C.c_void_p))
        return dict(zip(("states", "classes", "regs", "max_threads", "has_slow", "bytes"), (int(x) for x in info)))

    @property
    def fast2_bytes(self):
        return int(lib().emul_fast2_bytes(self._h))

    @property
    def fast_bytes(self):
        return int(lib().emul_fast_bytes(self._h))

    def __del__(self):
        try:
            lib().emul_free(self._h)
        except Exception:
            pass


def delim_fast(buf: np.ndarray, line_off: int, begin: int, end: int, sep: int, quote: int, cap: int = 64):
    """Bit-parallel delimiter path of the kernels (lc_delim_fast); None when the record is handed to the state machine."""
    L = lib()
    L.emul_delim_fast.restype = C.c_int64
    L.emul_delim_fast.argtypes = L.emul_delim_fsm.argtypes
    fo = np.zeros(cap, np.uint32)
    fl = np.zeros(cap, np.uint32)
    fd = np.zeros(cap, np.uint32)
    n = L.emul_delim_fast(buf.ctypes.data + line_off, begin, end, sep, quote, fo.ctypes.data_as(C.c_void_p),
                          fl.ctypes.data_as(C.c_void_p), fd.ctypes.data_as(C.c_void_p), cap)
    if n == -2:
        return None
    k = min(n, cap)
    return int(n), fo[:k].copy(), fl[:k].copy(), fd[:k].copy()


def delim_fsm(buf: np.ndarray, line_off: int, begin: int, end: int, sep: int, quote: int, cap: int = 64):
    """Run-skipping delimiter FSM of the kernels on the line starting at buf[line_off] (buf must keep 16 spare
    bytes on both sides of the line).  Returns None on an FSM error, else the list of (off, len, dq) columns."""
    fo = np.zeros(cap, np.uint32)
    fl = np.zeros(cap, np.uint32)
    fd = np.zeros(cap, np.uint32)
    n = lib().emul_delim_fsm(buf.ctypes.data + line_off, begin, end, sep, quote, fo.ctypes.data_as(C.c_void_p),
                             fl.ctypes.data_as(C.c_void_p), fd.ctypes.data_as(C.c_void_p), cap)
    if n < 0:
        return None
    k = min(int(n), cap)
    return int(n), list(zip(fo[:k].tolist(), fl[:k].tolist(), fd[:k].tolist()))


def sls_serialize_logs(events, enable_ns=True):
    """Host build of the kernels' SLS size / emit functions; same packing as capi.Engine.sls_serialize_logs."""
    arena = bytearray()
    koff, klen, voff, vlen, begin, times, nss = [], [], [], [], [0], [], []
    for t, ns, contents in events:
        for k, v in contents:
            koff.append(len(arena))
            klen.append(len(k))
            arena += k
            voff.append(len(arena))
            vlen.append(len(v))
            arena += v
        begin.append(len(koff))
        times.append(int(t) & 0xFFFFFFFF)
        nss.append(0xFFFFFFFF if (ns is None or not enable_ns) else int(ns))
    base = np.frombuffer(bytes(arena), np.uint8) if arena else np.zeros(1, np.uint8)
    a32 = lambda x: np.array(x if x else [0], np.uint32)  # noqa: E731
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    arrs = [a32(times), a32(nss), np.array(begin, np.uint64), a32(koff), a32(klen), a32(voff), a32(vlen)]
    cap = len(arena) + 32 * len(koff) + 16 * len(times) + 64
    out = np.zeros(cap, np.uint8)
    total = lib().emul_sls_serialize_logs(p(base), len(times), *[p(a) for a in arrs], p(out), cap)
    assert total <= cap
    return bytes(out[:total])
