"""GPU tier: BASELINE.json configs C4 (delimiter -> regex chain) and C5 (multi-pattern, Zipf lengths) at BASELINE
shape, every output row compared with the CPU oracle; the one-grid multi-pattern entry point; the strided event table;
events beyond the 16-bit capture registers; caller-provided streams.  Everything goes through the C-ABI."""
import os
import random

import numpy as np
import pytest

from oracle import oracle as orc  # checker only

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _lc():
    import loongcollector_b200 as lc
    return lc


@pytest.fixture(scope="module")
def eng():
    lc = _lc()
    e = lc.Engine(0)
    yield e
    e.close()


def _events(lines):
    base = b"".join(lines)
    ln = np.array([len(x) for x in lines], np.uint32)
    off = np.zeros(len(lines), np.uint32)
    if len(lines) > 1:
        off[1:] = np.cumsum(ln[:-1])
    return np.frombuffer(base, np.uint8) if base else np.zeros(0, np.uint8), off, ln


def merge_first_match(per_pattern, nkeys, gmax, sel=None):
    """Oracle results of every pattern -> the expected output of lc_regex_parse_multi (first match in array order wins;
    regex_match true == status != NOMATCH; rows zero unless the winner's status is OK)."""
    n = per_pattern[0][0].size
    which = np.full(n, 0xFF, np.uint8)
    status = np.ones(n, np.uint8)
    co = np.zeros((n, gmax), np.uint32)
    cl = np.zeros((n, gmax), np.uint32)
    for p, (st, o, l) in enumerate(per_pattern):
        take = (which == 0xFF) & (st != 1)
        if sel is not None:
            take &= (sel == 0xFF) | (sel == p)
        which[take] = p
        status[take] = st[take]
        g = o.shape[1]
        okrow = take & (st == 0)
        co[okrow, :g] = o[okrow]
        cl[okrow, :g] = l[okrow]
    return which, status, co, cl


def _expand_regex(pst, pco, pcl, poff, idx, off):
    """Oracle result on the pool lines -> exact expected tables of the sampled batch."""
    st = pst[idx]
    rel = (pco.astype(np.int64) - poff[:, None].astype(np.int64)) * (pst == 0)[:, None]
    co = ((rel[idx] + off[:, None].astype(np.int64)) * (st == 0)[:, None]).astype(np.uint32)
    return st, co, pcl[idx]


# ------------------------------------------------------------------------------------------- C5 / multi-pattern
def test_multi_small_mixed_noise_selectors_and_resume(monkeypatch):
    """Two and three patterns over noise + nginx + apache lines: first-match-wins, per-line selectors, rows zero-filled
    beyond the winner's groups; LC_B200_MULTI_SPLIT=1 forces one launch per pattern (the resume path)."""
    lc = _lc()
    from loongcollector_b200 import synth
    rng = random.Random(5)
    buf, off, ln, kind = synth.zipf_mixed_lines(3000, seed=77, pool=300)
    lines = [bytes(buf[o:o + l]) for o, l in zip(off, ln)]
    alpha = "ab c1-2\t\"[]x.=:/ yyd"
    lines += ["".join(rng.choice(alpha) for _ in range(rng.randint(0, 60))).encode() for _ in range(1500)]
    lines += [b"", b"k=v", b"GET /a 200", b"10.0.0.1 tail"]
    rng.shuffle(lines)
    base, off, ln = _events(lines)
    pats = [synth.NGINX_PATTERN, synth.APACHE_PATTERN, r"(\w+)=(.*)"]
    nkeys = [10, 11, 3]  # the third pattern has 2 groups but 3 keys: KEYS_MISMATCH for its matches
    per = [orc.regex_parse_batch(orc.Regex(p), base, off, ln, k) for p, k in zip(pats, nkeys)]
    per = [(st, co[:, :orc.Regex(p).ngroups], cl[:, :orc.Regex(p).ngroups]) for (st, co, cl), p in zip(per, pats)]
    sel = np.array([rng.choice([0xFF, 0xFF, 0, 1, 2]) for _ in lines], np.uint8)
    for split in ("0", "1"):
        monkeypatch.setenv("LC_B200_MULTI_SPLIT", split)
        e = lc.Engine(0)
        try:
            rxs = [lc.Regex(p) for p in pats]
            for npat in (2, 3):
                for s in (None, np.where(sel < npat, sel, 0xFF).astype(np.uint8)):
                    for pitch in (None, 16):
                        gmax = pitch or max(r.ngroups for r in rxs[:npat])
                        got = e.regex_parse_multi(rxs[:npat], nkeys[:npat], base, off, ln, sel=s, row_pitch=pitch)
                        want = merge_first_match(per[:npat], nkeys[:npat], gmax, s)
                        for g, w, name in zip(got, want, ("which", "status", "cap_off", "cap_len")):
                            assert np.array_equal(g, w), (split, npat, s is not None, pitch, name,
                                                          np.nonzero(np.atleast_1d(g != w).reshape(len(lines), -1)
                                                                     .any(axis=1))[0][:5])
        finally:
            e.close()


def test_multi_single_pattern_equals_regex_parse(eng):
    lc = _lc()
    from loongcollector_b200 import synth
    buf, off, ln = synth.nginx_lines(20000, seed=3, line_bytes=None)
    rx = lc.Regex(synth.NGINX_PATTERN)
    st, co, cl = eng.regex_parse(rx, buf, off, ln, 10)
    wh, mst, mco, mcl = eng.regex_parse_multi([rx], [10], buf, off, ln)
    assert np.array_equal(mst, st) and np.array_equal(mco, co) and np.array_equal(mcl, cl)
    assert np.array_equal(wh == 0, st == 0)


def test_events_beyond_16bit_registers_single_and_multi(eng):
    """Events of 65535 bytes or more are redone by the 32-bit-register kernel behind the staged one (no host round
    trip); neighbours in the same warp batch are unaffected."""
    lc = _lc()
    rng = random.Random(11)
    lines = []
    for L in (65534, 65535, 65536, 70001, 200000):
        lines.append(b"k=" + b"v" * (L - 2))
        lines.append(b"GET /" + b"a" * (L - 9) + b" 200")
        lines.append(b"x" * L)
    lines += [b"k=v", b"GET /abc 200", b"", b"nomatch here"] * 20
    rng.shuffle(lines)
    base, off, ln = _events(lines)
    pats = [r"(\w+) /(\w+) (\d+)", r"(\w)=(.*)", r"(x*)(.*)"]
    for p in pats:
        rx = lc.Regex(p)
        st, co, cl = eng.regex_parse(rx, base, off, ln, rx.ngroups)
        est, eco, ecl = orc.regex_parse_batch(orc.Regex(p), base, off, ln, rx.ngroups)
        assert np.array_equal(st, est) and np.array_equal(co, eco) and np.array_equal(cl, ecl), p
        got = eng.regex_match(rx, base, off, ln)
        assert np.array_equal(got, est != 1), p
    rxs = [lc.Regex(p) for p in pats]
    nk = [r.ngroups for r in rxs]
    per = [orc.regex_parse_batch(orc.Regex(p), base, off, ln, k) for p, k in zip(pats, nk)]
    got = eng.regex_parse_multi(rxs, nk, base, off, ln)
    want = merge_first_match(per, nk, max(nk))
    for g, w, name in zip(got, want, ("which", "status", "cap_off", "cap_len")):
        assert np.array_equal(g, w), name


def test_full_size_c5_multi_pattern_one_grid(eng):
    """C5 at >= 1 Mi lines: nginx + apache lines, Zipf(1.1) lengths clipped to [120, 8191] B (mean ~3.5 KB), BOTH
    patterns offered to every line in one grid (no generator ground truth), length-ordered visiting.  The lines are
    samples of an 8 Ki-line pool, so the oracle's per-pattern results on the pool, merged first-match-wins, expand to
    the exact expected tables -- every which / status / offset / length is compared."""
    import torch
    lc = _lc()
    from loongcollector_b200 import synth
    n = 1 << 20
    pool, kinds = synth.zipf_mixed_pool(n)
    buf, off, ln, kind = synth.zipf_mixed_lines(n)
    assert buf.size < (1 << 32) - 16
    idx = synth.pool_index(len(pool), n, synth.DEFAULT_SEED + 1)
    pbuf, poff, plen = _events([p[:-1] for p in pool])
    pats = [synth.NGINX_PATTERN, synth.APACHE_PATTERN]
    nkeys = [10, 11]
    per = []
    for p, k in zip(pats, nkeys):
        st, co, cl = orc.regex_parse_batch(orc.Regex(p), pbuf, poff, plen, k)
        per.append(_expand_regex(st, co, cl, poff, idx, off))
    # both verdicts of both patterns occur: nginx lines match only pattern 0, apache lines only pattern 1
    assert np.all(per[0][0][~kind] == 0) and np.all(per[0][0][kind] == 1)
    assert np.all(per[1][0][kind] == 0) and np.all(per[1][0][~kind] == 1)
    want = merge_first_match(per, nkeys, 11)
    rxs = [lc.Regex(p) for p in pats]
    d_buf = torch.from_numpy(buf).cuda()
    d_off = torch.from_numpy(off.view(np.int32)).cuda()
    d_len = torch.from_numpy(ln.view(np.int32)).cuda()
    d_wh = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_co = torch.empty(n * 11, dtype=torch.int32, device="cuda")
    d_cl = torch.empty(n * 11, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    eng.regex_parse_multi_dev(rxs, nkeys, d_buf.data_ptr(), buf.size, d_off.data_ptr(), d_len.data_ptr(), n, None,
                              d_wh.data_ptr(), d_st.data_ptr(), 11, d_co.data_ptr(), d_cl.data_ptr())
    eng.sync()
    assert np.array_equal(d_wh.cpu().numpy(), want[0])
    assert np.array_equal(d_st.cpu().numpy(), want[1])
    assert np.array_equal(d_co.cpu().numpy().view(np.uint32).reshape(n, 11), want[2])
    assert np.array_equal(d_cl.cpu().numpy().view(np.uint32).reshape(n, 11), want[3])
    # the single-pattern entry point on the matching halves (length-ordered path), both patterns
    for p, (pat, k) in enumerate(zip(pats, nkeys)):
        m = (kind == bool(p))
        so, sl = np.ascontiguousarray(off[m]), np.ascontiguousarray(ln[m])
        mcount = so.size
        G = rxs[p].ngroups
        t_off = torch.from_numpy(so.view(np.int32)).cuda()
        t_len = torch.from_numpy(sl.view(np.int32)).cuda()
        t_st = torch.empty(mcount, dtype=torch.uint8, device="cuda")
        t_co = torch.empty(mcount * G, dtype=torch.int32, device="cuda")
        t_cl = torch.empty(mcount * G, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        eng.regex_parse_dev(rxs[p], d_buf.data_ptr(), buf.size, t_off.data_ptr(), t_len.data_ptr(), mcount, k,
                            t_st.data_ptr(), t_co.data_ptr(), t_cl.data_ptr())
        eng.sync()
        assert np.array_equal(t_st.cpu().numpy(), per[p][0][m])
        assert np.array_equal(t_co.cpu().numpy().view(np.uint32).reshape(mcount, G), per[p][1][m])
        assert np.array_equal(t_cl.cpu().numpy().view(np.uint32).reshape(mcount, G), per[p][2][m])


# ------------------------------------------------------------------------------------------- C4 chain
def test_full_size_c4_delimiter_regex_chain(eng):
    """C4 at >= 1 Mi CSV lines: ProcessorParseDelimiterNative then ProcessorParseRegexNative on column 3, the regex
    reading the delimiter's field table IN PLACE (strided event table).  Lines are samples of a 16 Ki-line pool: the
    oracle's delimiter rows and the regex captures of column 3 on the pool expand to the exact expected tables; every
    status / nfields / f_off / f_len / f_dq row and every regex status / capture is compared."""
    import torch
    lc = _lc()
    from loongcollector_b200 import synth
    n = 1 << 20
    MF = 11
    pool = synth.csv_pool(n)
    buf, off, ln = synth.csv_lines(n)
    idx = synth.pool_index(len(pool), n, synth.DEFAULT_SEED + 1)
    pbuf, poff, plen = _events([p[:-1] for p in pool])
    pst, pnf, pfo, pfl, pfd = orc.delim_parse_batch(pbuf, poff, plen, b",", ord('"'), 10, True, True, MF)
    # rows of failed / blank lines and unused columns are zero: offsets are relative only where a field was stored
    stored = (np.arange(MF)[None, :] < np.minimum(pnf, MF)[:, None]) & np.isin(pst, (0, 3))[:, None]
    rel = (pfo.astype(np.int64) - poff[:, None].astype(np.int64)) * stored
    e_st, e_nf = pst[idx], pnf[idx]
    e_fo = ((rel[idx] + off[:, None].astype(np.int64)) * stored[idx]).astype(np.uint32)
    e_fl, e_fd = pfl[idx], pfd[idx]
    rx = lc.Regex(synth.CSV_URL_PATTERN)
    G = rx.ngroups
    rst, rco, rcl = orc.regex_parse_batch(orc.Regex(synth.CSV_URL_PATTERN), pbuf, np.ascontiguousarray(pfo[:, 3]),
                                          np.ascontiguousarray(pfl[:, 3]), G)
    # captures are relative to the line start (column 3 sits at the same place in every copy of a pool line)
    x_st, x_co, x_cl = _expand_regex(rst, rco, rcl, poff, idx, off)
    assert (pfd != 0).any() and (rst == 0).any() and (rst == 1).any()

    d_buf = torch.from_numpy(buf).cuda()
    d_off = torch.from_numpy(off.view(np.int32)).cuda()
    d_len = torch.from_numpy(ln.view(np.int32)).cuda()
    st4 = torch.empty(n, dtype=torch.uint8, device="cuda")
    nf4 = torch.empty(n, dtype=torch.int32, device="cuda")
    fo4 = torch.empty(n * MF, dtype=torch.int32, device="cuda")
    fl4 = torch.empty(n * MF, dtype=torch.int32, device="cuda")
    fd4 = torch.empty(n * MF, dtype=torch.int32, device="cuda")
    rs = torch.empty(n, dtype=torch.uint8, device="cuda")
    rco_d = torch.empty(n * G, dtype=torch.int32, device="cuda")
    rcl_d = torch.empty(n * G, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    eng.delim_parse_dev(d_buf.data_ptr(), buf.size, d_off.data_ptr(), d_len.data_ptr(), n, b",", ord('"'), 10, True,
                        True, MF, st4.data_ptr(), nf4.data_ptr(), fo4.data_ptr(), fl4.data_ptr(), fd4.data_ptr())
    eng.regex_parse_strided_dev(rx, d_buf.data_ptr(), buf.size, fo4.data_ptr() + 3 * 4, fl4.data_ptr() + 3 * 4, MF, n, G,
                                rs.data_ptr(), rco_d.data_ptr(), rcl_d.data_ptr())
    eng.sync()
    # the same chain through the delimiter's dense column tap (what bench.py --config c4 times)
    to4 = torch.empty(n, dtype=torch.int32, device="cuda")
    tl4 = torch.empty(n, dtype=torch.int32, device="cuda")
    st5 = torch.empty(n, dtype=torch.uint8, device="cuda")
    fo5 = torch.empty(n * MF, dtype=torch.int32, device="cuda")
    rs2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    rco2 = torch.empty(n * G, dtype=torch.int32, device="cuda")
    rcl2 = torch.empty(n * G, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    eng.delim_parse_dev(d_buf.data_ptr(), buf.size, d_off.data_ptr(), d_len.data_ptr(), n, b",", ord('"'), 10, True,
                        True, MF, st5.data_ptr(), nf4.data_ptr(), fo5.data_ptr(), fl4.data_ptr(), fd4.data_ptr(), 3,
                        to4.data_ptr(), tl4.data_ptr())
    eng.regex_parse_dev(rx, d_buf.data_ptr(), buf.size, to4.data_ptr(), tl4.data_ptr(), n, G, rs2.data_ptr(),
                        rco2.data_ptr(), rcl2.data_ptr())
    eng.sync()
    assert torch.equal(to4, fo5.view(n, MF)[:, 3]) and torch.equal(tl4, fl4.view(n, MF)[:, 3])
    assert torch.equal(st5, st4) and torch.equal(fo5, fo4)
    assert torch.equal(rs2, rs) and torch.equal(rco2, rco_d) and torch.equal(rcl2, rcl_d)
    assert np.array_equal(st4.cpu().numpy(), e_st)
    assert np.array_equal(nf4.cpu().numpy().view(np.uint32), e_nf)
    assert np.array_equal(fo4.cpu().numpy().view(np.uint32).reshape(n, MF), e_fo)
    assert np.array_equal(fl4.cpu().numpy().view(np.uint32).reshape(n, MF), e_fl)
    assert np.array_equal(fd4.cpu().numpy().view(np.uint32).reshape(n, MF), e_fd)
    assert np.array_equal(rs.cpu().numpy(), x_st)
    assert np.array_equal(rco_d.cpu().numpy().view(np.uint32).reshape(n, G), x_co)
    assert np.array_equal(rcl_d.cpu().numpy().view(np.uint32).reshape(n, G), x_cl)


def test_strided_event_table_small_and_fallback_kernels(monkeypatch):
    """lc_regex_parse_strided_dev == lc_regex_parse_dev on the gathered table, for the single-pass kernel and for the
    two-pass fall-backs (which densify the table first)."""
    import torch
    lc = _lc()
    from loongcollector_b200 import synth
    buf, off, ln = synth.nginx_lines(5000, seed=8, line_bytes=None)
    stride = 7
    wide_off = np.zeros(5000 * stride, np.uint32)
    wide_len = np.zeros(5000 * stride, np.uint32)
    wide_off[2::stride] = off
    wide_len[2::stride] = ln
    est, eco, ecl = orc.regex_parse_batch(orc.Regex(synth.NGINX_PATTERN), buf, off, ln, 10)
    for variant in ("", "fast2", "generic"):
        if variant:
            monkeypatch.setenv("LC_B200_REGEX_KERNEL", variant)
        e = lc.Engine(0)
        try:
            rx = lc.Regex(synth.NGINX_PATTERN)
            d_buf = torch.from_numpy(buf).cuda()
            d_o = torch.from_numpy(wide_off.view(np.int32)).cuda()
            d_l = torch.from_numpy(wide_len.view(np.int32)).cuda()
            st = torch.empty(5000, dtype=torch.uint8, device="cuda")
            co = torch.empty(5000 * 10, dtype=torch.int32, device="cuda")
            cl = torch.empty(5000 * 10, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            e.regex_parse_strided_dev(rx, d_buf.data_ptr(), buf.size, d_o.data_ptr() + 8, d_l.data_ptr() + 8, stride,
                                      5000, 10, st.data_ptr(), co.data_ptr(), cl.data_ptr())
            e.sync()
            assert np.array_equal(st.cpu().numpy(), est), variant
            assert np.array_equal(co.cpu().numpy().view(np.uint32).reshape(5000, 10), eco), variant
            assert np.array_equal(cl.cpu().numpy().view(np.uint32).reshape(5000, 10), ecl), variant
        finally:
            e.close()


def test_engine_on_caller_stream_and_back_to_back_async_calls(eng):
    """*_dev regex calls never wait for the device: several calls queued back to back on a caller-provided stream
    (lc_engine_set_stream) produce the same tables as synchronous calls."""
    import torch
    lc = _lc()
    from loongcollector_b200 import synth
    rx = lc.Regex(synth.NGINX_PATTERN)
    s = torch.cuda.Stream()
    outs = []
    batches = [synth.nginx_lines(30000 + 1000 * k, seed=100 + k, line_bytes=256) for k in range(4)]
    eng.set_stream(s.cuda_stream)
    try:
        with torch.cuda.stream(s):
            for buf, off, ln in batches:
                n = off.size
                d_buf = torch.from_numpy(buf).cuda()
                d_o = torch.from_numpy(off.view(np.int32)).cuda()
                d_l = torch.from_numpy(ln.view(np.int32)).cuda()
                st = torch.empty(n, dtype=torch.uint8, device="cuda")
                co = torch.empty(n * 10, dtype=torch.int32, device="cuda")
                cl = torch.empty(n * 10, dtype=torch.int32, device="cuda")
                eng.regex_parse_dev(rx, d_buf.data_ptr(), buf.size, d_o.data_ptr(), d_l.data_ptr(), n, 10,
                                    st.data_ptr(), co.data_ptr(), cl.data_ptr())
                outs.append((d_buf, d_o, d_l, st, co, cl))
        s.synchronize()
    finally:
        eng.set_stream(None)
    for (buf, off, ln), (_, _, _, st, co, cl) in zip(batches, outs):
        est, eco, ecl = orc.regex_parse_batch(orc.Regex(synth.NGINX_PATTERN), buf, off, ln, 10)
        assert np.array_equal(st.cpu().numpy(), est)
        assert np.array_equal(co.cpu().numpy().view(np.uint32).reshape(-1, 10), eco)
        assert np.array_equal(cl.cpu().numpy().view(np.uint32).reshape(-1, 10), ecl)


def test_host_entry_points_reject_events_outside_the_arena(eng):
    lc = _lc()
    base = np.frombuffer(b"k=v,a,b\n" * 8, np.uint8)
    off = np.array([0, 8, 60], np.uint32)
    ln = np.array([7, 7, 9], np.uint32)  # the last event ends at 69 > 64
    rx = lc.Regex(r"(\w)=(.*)")
    for call in (lambda: eng.regex_parse(rx, base, off, ln, 2), lambda: eng.regex_match(rx, base, off, ln),
                 lambda: eng.regex_prefix_match(rx, base, off, ln),
                 lambda: eng.delim_parse(base, off, ln, b",", ord('"'), 3, True, True, 4),
                 lambda: eng.regex_parse_multi([rx], [2], base, off, ln)):
        with pytest.raises(lc.LcError) as ei:
            call()
        assert ei.value.code == lc.capi.LC_ERR_INVALID_ARG


def test_plugin_batched_and_per_group_process_match_the_oracle(eng):
    """ProcessorInstance::Process of the B200-backed ProcessorParseRegexNative over <= 512 KB groups with pinned
    arenas -- one group per call and the batched vector override -- leaves exactly the events / contents the flat
    oracle predicts, and the CPU reference arm (oracle/ref_plugin.cpp) agrees on the same checksum."""
    import bench
    from loongcollector_b200 import synth

    class A:
        lines = 200000
    import torch
    cfg = bench.C2(A, 0, 1, eng, torch.device("cuda", 0))
    cfg.setup_host()
    est, eco, ecl = orc.regex_parse_batch(orc.Regex(synth.NGINX_PATTERN), cfg.buf, cfg.off, cfg.ln, 10)
    want = bench.expected_plugin_stats(cfg.buf, est, eco, ecl, synth.NGINX_KEYS)
    for mode in (0, 1):
        secs, stats = cfg.e2e_plugin(2, mode=mode)
        assert stats["groups"] == 98 and stats["in_events"] == A.lines
        for k, v in want.items():
            assert stats[k] == v, (mode, k, stats[k], v)
        # ProcessorInstance counters (a8): both repetitions are counted
        assert stats["ctr_in_events"] == 2 * A.lines and stats["ctr_out_events"] == 2 * want["out_events"]
        assert stats["ctr_in_bytes"] > stats["ctr_in_events"] * 255 and stats["ctr_process_ns"] > 0
    secs, cstats = bench.cpu_plugin_regex(cfg.buf, cfg.off, cfg.ln, 4, 1)
    for k, v in want.items():
        assert cstats[k] == v, ("cpu arm", k)
    assert cstats["ctr_in_bytes"] * 2 == stats["ctr_in_bytes"] and cstats["ctr_out_bytes"] * 2 == stats["ctr_out_bytes"]


@pytest.mark.gpu
def test_delimiter_regex_chain_host_call_matches_the_two_stages(eng):
    """lc_delim_regex_chain (one upload, chunked through three streams) = lc_delim_parse followed by lc_regex_parse on the
    column, and both equal the oracle; large enough for several chunks, with quoted / malformed lines and blank lines."""
    from loongcollector_b200 import synth
    import loongcollector_b200 as lc
    n = 1 << 20
    buf, off, ln = synth.csv_lines(n, seed=77)
    buf = buf.copy()
    # a few malformed and blank lines
    for i in range(0, n, 9973):
        buf[off[i]] = ord('"')
    for i in range(5, n, 19997):
        buf[off[i]:off[i] + ln[i]] = ord(" ")
    rx = lc.Regex(synth.CSV_URL_PATTERN)
    MF = 11
    st, nf, fo, fl, fd, rs, co, cl = eng.delim_regex_chain(buf, off, ln, b",", ord('"'), 10, True, True, MF, 3, rx)
    st2, nf2, fo2, fl2, fd2 = eng.delim_parse(buf, off, ln, b",", ord('"'), 10, True, True, MF)
    assert np.array_equal(st, st2) and np.array_equal(nf, nf2)
    assert np.array_equal(fo, fo2) and np.array_equal(fl, fl2) and np.array_equal(fd, fd2)
    rs2, co2, cl2 = eng.regex_parse(rx, buf, fo2[:, 3].copy(), fl2[:, 3].copy(), rx.ngroups)
    assert np.array_equal(rs, rs2) and np.array_equal(co, co2) and np.array_equal(cl, cl2)
    ns = 1 << 15
    est, enf, efo, efl, efd = orc.delim_parse_batch(buf, off[:ns], ln[:ns], b",", ord('"'), 10, True, True, MF)
    assert np.array_equal(st[:ns], est) and np.array_equal(fo[:ns], efo) and np.array_equal(fl[:ns], efl)
    erst, eco, ecl = orc.regex_parse_batch(orc.Regex(synth.CSV_URL_PATTERN), buf, efo[:, 3].copy(), efl[:, 3].copy(), 2)
    assert np.array_equal(rs[:ns], erst) and np.array_equal(co[:ns], eco) and np.array_equal(cl[:ns], ecl)
    assert (st != 0).any() and (rs == 0).any()
