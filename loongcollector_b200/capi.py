"""ctypes binding of include/lc_b200.h.  Host arrays are numpy; device pointers are plain ints
(e.g. ``torch.Tensor.data_ptr()``) -- torch is only plumbing for HBM allocations and streams."""
import ctypes as C
import os

import numpy as np

from . import _build

LC_OK, LC_ERR_INVALID_ARG, LC_ERR_CUDA, LC_ERR_REGEX_INVALID, LC_ERR_REGEX_UNSUPPORTED, LC_ERR_CAPACITY, \
    LC_ERR_TOO_LARGE = range(7)
LC_ML_IS_LAST, LC_ML_MATCHED = 1, 2

_LIB = None


class LcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lc_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads libloongcollector_b200.so (building it in-tree if sources are newer). Raises if unavailable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = _build.SO
    if _build.needs_build():
        if os.path.exists("/usr/local/cuda/bin/nvcc") or os.environ.get("NVCC"):
            _build.build()
        elif not os.path.exists(so):
            raise ImportError("libloongcollector_b200.so is missing and nvcc is not available; "
                              "run `python -m loongcollector_b200._build` (there is no CPU fallback)")
    L = C.CDLL(so)
    vp, u64, u32, u8, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint8, C.c_int
    L.lc_version.restype = C.c_char_p
    L.lc_last_error.restype = C.c_char_p
    L.lc_device_count.restype = i32
    L.lc_engine_create.argtypes = [i32, C.POINTER(vp)]
    L.lc_engine_destroy.argtypes = [vp]
    L.lc_engine_sync.argtypes = [vp]
    L.lc_engine_stream.restype = vp
    L.lc_engine_stream.argtypes = [vp]
    L.lc_engine_launch_count.restype = u64
    L.lc_engine_launch_count.argtypes = [vp]
    L.lc_host_alloc.restype = vp
    L.lc_host_alloc.argtypes = [C.c_size_t]
    L.lc_host_free.argtypes = [vp]
    L.lc_regex_compile.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    L.lc_regex_free.argtypes = [vp]
    L.lc_regex_error.restype = C.c_char_p
    L.lc_regex_error.argtypes = [vp]
    L.lc_regex_ngroups.restype = u32
    L.lc_regex_ngroups.argtypes = [vp]
    L.lc_regex_info.argtypes = [vp, vp]
    split_args = [vp, vp, u64, u8, vp, vp, u64, C.POINTER(u64)]
    L.lc_split_lines.argtypes = split_args
    L.lc_split_lines_dev.argtypes = split_args
    parse_args = [vp, vp, vp, u64, vp, vp, u64, u32, vp, vp, vp]
    L.lc_regex_parse.argtypes = parse_args
    L.lc_regex_parse_dev.argtypes = parse_args
    L.lc_engine_set_stream.argtypes = [vp, vp]
    L.lc_regex_parse_strided_dev.argtypes = [vp, vp, vp, u64, vp, vp, u32, u64, u32, vp, vp, vp]
    multi_args = [vp, vp, u32, vp, vp, u64, vp, vp, u64, vp, vp, vp, u32, vp, vp]
    L.lc_regex_parse_multi.argtypes = multi_args
    L.lc_regex_parse_multi_dev.argtypes = multi_args
    rl_args = [vp, vp, u64, vp, vp, i32, C.POINTER(u64), C.POINTER(C.c_int32)]
    L.lc_remove_last_incomplete_log.argtypes = rl_args
    L.lc_remove_last_incomplete_log_dev.argtypes = rl_args
    L.lc_sls_serialize_parsed_dev.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, u32, u64, vp, vp, u32, C.c_char_p, u32, vp,
                                              vp, vp, u64, C.POINTER(u64)]
    L.lc_regex_prefix_match.argtypes = [vp, vp, vp, u64, vp, vp, u64, vp]
    L.lc_regex_match.argtypes = [vp, vp, vp, u64, vp, vp, u64, vp]
    L.lc_regex_match_dev.argtypes = [vp, vp, vp, u64, vp, vp, u64, vp]
    ml_args = [vp, vp, u64, vp, vp, vp, i32, vp, vp, vp, u64, C.POINTER(u64), vp]
    L.lc_multiline_split.argtypes = ml_args
    L.lc_multiline_split_dev.argtypes = ml_args
    dl_args = [vp, vp, u64, vp, vp, u64, vp, u32, u8, u32, i32, i32, u32, vp, vp, vp, vp, vp]
    L.lc_delim_parse.argtypes = dl_args
    L.lc_delim_parse_dev.argtypes = dl_args
    L.lc_delim_parse_tap_dev.argtypes = dl_args + [u32, vp, vp]
    L.lc_delim_regex_chain.argtypes = dl_args + [u32, vp, u32, vp, vp, vp]
    L.lc_sls_serialize_logs.argtypes = [vp, vp, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, u64, C.POINTER(u64)]
    _LIB = L
    return L


def version():
    return lib().lc_version().decode()


def device_count():
    return int(lib().lc_device_count())


def _check(rc):
    if rc != LC_OK:
        raise LcError(rc, lib().lc_last_error().decode())


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def _u8(buf):
    if isinstance(buf, np.ndarray):
        assert buf.dtype == np.uint8 and buf.flags.c_contiguous
        return buf
    return np.frombuffer(bytes(buf), dtype=np.uint8)


class Regex:
    """Compiled pattern (host object; boost::regex(pattern) replacement)."""

    def __init__(self, pattern):
        if isinstance(pattern, str):
            pattern = pattern.encode("utf-8")
        self.pattern = pattern
        h = C.c_void_p()
        L = lib()
        rc = L.lc_regex_compile(pattern, len(pattern), C.byref(h))
        self._h = h
        if rc != LC_OK:
            msg = L.lc_last_error().decode()
            L.lc_regex_free(h)
            self._h = None
            raise LcError(rc, msg)
        self.ngroups = int(L.lc_regex_ngroups(h))
        info = np.zeros(8, np.uint32)
        L.lc_regex_info(h, _p(info))
        self.info = dict(zip(("mode", "classes", "walkers", "ctx", "rev_states", "prefix_states", "table_bytes",
                              "insts"), (int(x) for x in info)))

    def __del__(self):
        try:
            if self._h:
                lib().lc_regex_free(self._h)
        except Exception:
            pass


def _rh(r):
    return r._h if r is not None else None


class Engine:
    """One engine per (GPU, host thread)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().lc_engine_create(device, C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if self._h:
            lib().lc_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(lib().lc_engine_sync(self._h))

    @property
    def stream(self):
        return int(lib().lc_engine_stream(self._h) or 0)

    @property
    def launches(self):
        return int(lib().lc_engine_launch_count(self._h))

    # ---- host-buffer API ---------------------------------------------------------------------
    def split_lines(self, buf, split_char=10, cap=None):
        a = _u8(buf)
        cap = int(cap if cap is not None else max(1, a.size))
        off = np.empty(cap, np.uint32)
        ln = np.empty(cap, np.uint32)
        n = C.c_uint64(0)
        _check(lib().lc_split_lines(self._h, _p(a), a.size, split_char, _p(off), _p(ln), cap, C.byref(n)))
        return off[:n.value], ln[:n.value]

    def regex_parse(self, rx, base, ev_off, ev_len, nkeys):
        a = _u8(base)
        ev_off = np.ascontiguousarray(ev_off, np.uint32)
        ev_len = np.ascontiguousarray(ev_len, np.uint32)
        n, G = ev_off.size, rx.ngroups
        status = np.empty(n, np.uint8)
        co = np.empty((n, G), np.uint32)
        cl = np.empty((n, G), np.uint32)
        _check(lib().lc_regex_parse(self._h, rx._h, _p(a), a.size, _p(ev_off), _p(ev_len), n, nkeys, _p(status),
                                    _p(co), _p(cl)))
        return status, co, cl

    @staticmethod
    def _multi_handles(rxs, nkeys):
        arr = (C.c_void_p * len(rxs))(*[r._h.value if isinstance(r._h, C.c_void_p) else r._h for r in rxs])
        nk = np.ascontiguousarray(nkeys, np.uint32)
        assert nk.size == len(rxs)
        return arr, nk

    def regex_parse_multi(self, rxs, nkeys, base, ev_off, ev_len, sel=None, row_pitch=None):
        """First-match-wins over several patterns in one grid; returns (which, status, cap_off, cap_len)."""
        a = _u8(base)
        ev_off = np.ascontiguousarray(ev_off, np.uint32)
        ev_len = np.ascontiguousarray(ev_len, np.uint32)
        n = ev_off.size
        G = int(row_pitch if row_pitch is not None else max(r.ngroups for r in rxs))
        arr, nk = self._multi_handles(rxs, nkeys)
        which = np.empty(n, np.uint8)
        status = np.empty(n, np.uint8)
        co = np.empty((n, G), np.uint32)
        cl = np.empty((n, G), np.uint32)
        if sel is not None:
            sel = np.ascontiguousarray(sel, np.uint8)
        _check(lib().lc_regex_parse_multi(self._h, arr, len(rxs), _p(nk), _p(a), a.size, _p(ev_off), _p(ev_len), n,
                                          _p(sel), _p(which), _p(status), G, _p(co), _p(cl)))
        return which, status, co, cl

    def regex_parse_multi_dev(self, rxs, nkeys, d_base, base_len, d_ev_off, d_ev_len, n, d_sel, d_which, d_status,
                              row_pitch, d_cap_off, d_cap_len):
        arr, nk = self._multi_handles(rxs, nkeys)
        _check(lib().lc_regex_parse_multi_dev(self._h, arr, len(rxs), _p(nk), _p(d_base), base_len, _p(d_ev_off),
                                              _p(d_ev_len), n, _p(d_sel), _p(d_which), _p(d_status), row_pitch,
                                              _p(d_cap_off), _p(d_cap_len)))

    def regex_parse_strided_dev(self, rx, d_base, base_len, d_ev_off, d_ev_len, ev_stride, n, nkeys, d_status,
                                d_cap_off, d_cap_len):
        _check(lib().lc_regex_parse_strided_dev(self._h, rx._h, _p(d_base), base_len, _p(d_ev_off), _p(d_ev_len),
                                                ev_stride, n, nkeys, _p(d_status), _p(d_cap_off), _p(d_cap_len)))

    def set_stream(self, stream):
        _check(lib().lc_engine_set_stream(self._h, _p(stream) if stream else None))

    def remove_last_incomplete_log(self, buf, start, end, allow_rollback=True):
        """LogFileReader::RemoveLastIncompleteLog (raw text) -> (bytes to keep, rollbackLineFeedCount)."""
        a = _u8(buf)
        keep, rb = C.c_uint64(0), C.c_int32(0)
        _check(lib().lc_remove_last_incomplete_log(self._h, _p(a), a.size, _rh(start), _rh(end),
                                                   int(bool(allow_rollback)), C.byref(keep), C.byref(rb)))
        return int(keep.value), int(rb.value)

    def regex_prefix_match(self, rx, base, ev_off, ev_len):
        a = _u8(base)
        ev_off = np.ascontiguousarray(ev_off, np.uint32)
        ev_len = np.ascontiguousarray(ev_len, np.uint32)
        out = np.empty(ev_off.size, np.uint8)
        _check(lib().lc_regex_prefix_match(self._h, rx._h, _p(a), a.size, _p(ev_off), _p(ev_len), ev_off.size,
                                           _p(out)))
        return out.astype(bool)

    def regex_match(self, rx, base, ev_off, ev_len):
        a = _u8(base)
        ev_off = np.ascontiguousarray(ev_off, np.uint32)
        ev_len = np.ascontiguousarray(ev_len, np.uint32)
        out = np.empty(ev_off.size, np.uint8)
        _check(lib().lc_regex_match(self._h, rx._h, _p(a), a.size, _p(ev_off), _p(ev_len), ev_off.size, _p(out)))
        return out.astype(bool)

    def multiline_split(self, buf, start, cont, end, discard, cap=None):
        a = _u8(buf)
        cap = int(cap if cap is not None else max(1, a.size))
        off = np.empty(cap, np.uint32)
        ln = np.empty(cap, np.uint32)
        fl = np.empty(cap, np.uint8)
        ctr = np.zeros(3, np.uint64)
        n = C.c_uint64(0)
        _check(lib().lc_multiline_split(self._h, _p(a), a.size, _rh(start), _rh(cont), _rh(end), int(bool(discard)),
                                        _p(off), _p(ln), _p(fl), cap, C.byref(n), _p(ctr)))
        return off[:n.value], ln[:n.value], fl[:n.value], ctr

    def delim_parse(self, base, ev_off, ev_len, sep: bytes, quote: int, nkeys, extend, allow_short, max_fields):
        a = _u8(base)
        ev_off = np.ascontiguousarray(ev_off, np.uint32)
        ev_len = np.ascontiguousarray(ev_len, np.uint32)
        n = ev_off.size
        status = np.empty(n, np.uint8)
        nf = np.empty(n, np.uint32)
        fo = np.empty((n, max_fields), np.uint32)
        fl = np.empty((n, max_fields), np.uint32)
        fd = np.empty((n, max_fields), np.uint32)
        sp = np.frombuffer(sep, np.uint8)
        _check(lib().lc_delim_parse(self._h, _p(a), a.size, _p(ev_off), _p(ev_len), n, _p(sp), len(sep), quote, nkeys,
                                    int(bool(extend)), int(bool(allow_short)), max_fields, _p(status), _p(nf), _p(fo),
                                    _p(fl), _p(fd)))
        return status, nf, fo, fl, fd

    # ---- device-pointer API (ints = device addresses) ---------------------------------------------
    def sls_serialize_logs(self, events, enable_ns=True):
        """events: list of (time, ns or None, [(key bytes, value bytes), ...]) -> the LogGroup's `Logs` fields (bytes).
        Packs the keys and values into one arena the way LogEvent contents alias the SourceBuffer."""
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
        n = len(times)
        base = np.frombuffer(bytes(arena), np.uint8) if arena else np.zeros(1, np.uint8)
        a32 = lambda x: np.array(x if x else [0], np.uint32)  # noqa: E731
        need = C.c_uint64(0)
        cap = len(arena) + 32 * len(koff) + 16 * n + 64
        out = np.zeros(max(cap, 1), np.uint8)
        _check(lib().lc_sls_serialize_logs(self._h, _p(base), len(arena), n, _p(a32(times)), _p(a32(nss)),
                                           _p(np.array(begin, np.uint64)), _p(a32(koff)), _p(a32(klen)),
                                           _p(a32(voff)), _p(a32(vlen)), _p(out), cap, C.byref(need)))
        return bytes(out[:need.value])

    def sls_serialize_parsed_dev(self, d_base, base_len, d_ev_off, d_ev_len, d_status, d_cap_off, d_cap_len, row_pitch,
                                 n, keys, fail_key, d_ev_time, d_ev_time_ns, d_out, out_cap):
        """keys: list of bytes; fail_key: bytes or None.  Returns the number of wire bytes written to d_out."""
        arr = (C.c_char_p * max(len(keys), 1))(*keys)
        kl = np.array([len(k) for k in keys] or [0], np.uint32)
        need = C.c_uint64(0)
        _check(lib().lc_sls_serialize_parsed_dev(self._h, _p(d_base), base_len, _p(d_ev_off), _p(d_ev_len),
                                                 _p(d_status), _p(d_cap_off), _p(d_cap_len), row_pitch, n, arr, _p(kl),
                                                 len(keys), fail_key, len(fail_key) if fail_key else 0, _p(d_ev_time),
                                                 _p(d_ev_time_ns), _p(d_out), out_cap, C.byref(need)))
        return int(need.value)

    def split_lines_dev(self, d_buf, length, split_char, d_off, d_len, cap):
        n = C.c_uint64(0)
        _check(lib().lc_split_lines_dev(self._h, _p(d_buf), length, split_char, _p(d_off), _p(d_len), cap,
                                        C.byref(n)))
        return n.value

    def regex_parse_dev(self, rx, d_base, base_len, d_ev_off, d_ev_len, n, nkeys, d_status, d_cap_off, d_cap_len):
        _check(lib().lc_regex_parse_dev(self._h, rx._h, _p(d_base), base_len, _p(d_ev_off), _p(d_ev_len), n, nkeys,
                                        _p(d_status), _p(d_cap_off), _p(d_cap_len)))

    def multiline_split_dev(self, d_buf, length, start, cont, end, discard, d_off, d_len, d_flags, cap):
        n = C.c_uint64(0)
        ctr = np.zeros(3, np.uint64)
        _check(lib().lc_multiline_split_dev(self._h, _p(d_buf), length, _rh(start), _rh(cont), _rh(end),
                                            int(bool(discard)), _p(d_off), _p(d_len), _p(d_flags), cap, C.byref(n),
                                            _p(ctr)))
        return n.value, ctr

    def delim_regex_chain(self, base, ev_off, ev_len, sep: bytes, quote, nkeys, extend, allow_short, max_fields, column,
                          rx, regex_nkeys=None):
        """Delimiter stage + regex stage on one column with host buffers (lc_delim_regex_chain); returns
        (status, nfields, f_off, f_len, f_dq, re_status, cap_off, cap_len)."""
        a = _u8(base)
        ev_off = np.ascontiguousarray(ev_off, np.uint32)
        ev_len = np.ascontiguousarray(ev_len, np.uint32)
        n, MF, G = ev_off.size, int(max_fields), rx.ngroups
        st, nf = np.empty(n, np.uint8), np.empty(n, np.uint32)
        fo, fl, fd = (np.empty((n, MF), np.uint32) for _ in range(3))
        rs = np.empty(n, np.uint8)
        co, cl = np.empty((n, G), np.uint32), np.empty((n, G), np.uint32)
        sp = np.frombuffer(sep, np.uint8)
        _check(lib().lc_delim_regex_chain(self._h, _p(a), a.size, _p(ev_off), _p(ev_len), n, _p(sp), len(sep), quote,
                                          nkeys, int(bool(extend)), int(bool(allow_short)), MF, _p(st), _p(nf), _p(fo),
                                          _p(fl), _p(fd), column, rx._h, G if regex_nkeys is None else regex_nkeys,
                                          _p(rs), _p(co), _p(cl)))
        return st, nf, fo, fl, fd, rs, co, cl

    def delim_parse_dev(self, d_base, base_len, d_ev_off, d_ev_len, n, sep: bytes, quote, nkeys, extend, allow_short,
                        max_fields, d_status, d_nf, d_fo, d_fl, d_fd, tap_col=None, d_tap_off=None, d_tap_len=None):
        sp = np.frombuffer(sep, np.uint8)
        _check(lib().lc_delim_parse_tap_dev(self._h, _p(d_base), base_len, _p(d_ev_off), _p(d_ev_len), n, _p(sp),
                                            len(sep), quote, nkeys, int(bool(extend)), int(bool(allow_short)),
                                            max_fields, _p(d_status), _p(d_nf), _p(d_fo), _p(d_fl), _p(d_fd),
                                            0xFFFFFFFF if tap_col is None else tap_col, _p(d_tap_off), _p(d_tap_len)))


class HostProcessor:
    """A B200-backed Processor of the C++ host layer (include/lc_b200_host.h), driven with JSON event groups
    the way the reference's unit tests drive the original classes."""

    def __init__(self, ptype: str, config: dict):
        import json
        L = lib()
        L.lc_host_processor_create.restype = C.c_void_p
        L.lc_host_processor_create.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.lc_host_processor_destroy.argtypes = [C.c_void_p]
        L.lc_host_processor_process.restype = C.c_void_p
        L.lc_host_processor_process.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.lc_host_processor_counters.restype = C.c_void_p
        L.lc_host_processor_counters.argtypes = [C.c_void_p]
        L.lc_host_string_free.argtypes = [C.c_void_p]
        err = C.c_void_p()
        self._h = L.lc_host_processor_create(ptype.encode(), json.dumps(config).encode("utf-8"), C.byref(err))
        if not self._h:
            msg = C.string_at(err.value).decode() if err.value else "unknown error"
            if err.value:
                L.lc_host_string_free(err)
            raise LcError(LC_ERR_INVALID_ARG, msg)
        self.type = ptype

    def process(self, group, enable_event_meta=True):
        """group: dict in the reference's event-group JSON shape (or None). Returns the processed group."""
        import json
        L = lib()
        err = C.c_void_p()
        out = L.lc_host_processor_process(self._h, json.dumps(group).encode("utf-8"), int(enable_event_meta),
                                          C.byref(err))
        if not out:
            msg = C.string_at(err.value).decode() if err.value else "unknown error"
            if err.value:
                L.lc_host_string_free(err)
            raise LcError(LC_ERR_CUDA, msg)
        s = C.string_at(out).decode("utf-8")
        L.lc_host_string_free(out)
        return json.loads(s)

    def counters(self):
        import json
        L = lib()
        out = L.lc_host_processor_counters(self._h)
        s = C.string_at(out).decode()
        L.lc_host_string_free(out)
        return json.loads(s)

    def __del__(self):
        try:
            if self._h:
                lib().lc_host_processor_destroy(self._h)
        except Exception:
            pass


def host_sls_serialize(group, enable_ns=False):
    """SLSEventGroupSerializer::Serialize of the C++ host layer on a JSON event group: (bytes, None) or (None, error)."""
    import json
    L = lib()
    L.lc_host_sls_serialize.restype = C.c_void_p
    L.lc_host_sls_serialize.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_void_p)]
    L.lc_host_string_free.argtypes = [C.c_void_p]
    err = C.c_void_p()
    n = C.c_ulonglong(0)
    out = L.lc_host_sls_serialize(json.dumps(group).encode("utf-8"), int(bool(enable_ns)), C.byref(n), C.byref(err))
    if not out:
        msg = C.string_at(err.value).decode() if err.value else "unknown error"
        if err.value:
            L.lc_host_string_free(err)
        return None, msg
    data = C.string_at(out, n.value)
    L.lc_host_string_free(out)
    return data, None
