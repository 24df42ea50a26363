#!/usr/bin/env python3
"""Device-resident throughput of the other BASELINE.json configs on ONE GPU (C1 split, C3 multiline,
C4 delimiter -> regex), each next to the CPU oracle on a bounded sample and with the algorithmic-bytes
roofline of SURVEY.md section 8(d).  bench.py stays the contract line (C2); this prints one JSON line per config.

  python tools/bench_configs.py [--scale 1.0]      (scale < 1 shrinks every config for quick runs)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, steps=5, warmup=2):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="", help="comma list of configs to run (c1,c3,c4,c5); default all")
    args = ap.parse_args()
    import torch
    import loongcollector_b200 as lc
    from loongcollector_b200 import synth
    from oracle import oracle as orc
    sys.path.insert(0, ROOT)
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    eng = lc.Engine(0)
    dev = torch.device("cuda", 0)

    def dput(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    out = []
    # ------------------------------------------------------------------ C1: newline split, 1 Mi x 512 B
    n1 = int((1 << 20) * args.scale)
    buf, off, ln = synth.newline_lines(n1, 512)
    d_buf = dput(buf)
    d_off = torch.empty(n1 + 16, dtype=torch.int32, device=dev)
    d_len = torch.empty(n1 + 16, dtype=torch.int32, device=dev)
    got = {}

    def c1():
        got["n"] = eng.split_lines_dev(d_buf.data_ptr(), buf.size, 10, d_off.data_ptr(), d_len.data_ptr(), n1 + 16)

    dt = timeit(c1)
    assert got["n"] == n1
    assert np.array_equal(d_off[:n1].cpu().numpy().view(np.uint32), off)
    ns = min(n1, 1 << 17)
    t0 = time.perf_counter()
    orc.split_lines(buf[:ns * 512])
    cpu = ns * 512 / (time.perf_counter() - t0) / 1e6
    alg = buf.size + 8 * n1
    out.append({"config": "C1 ProcessorSplitLogStringNative newline split, %d x 512 B" % n1, "MBps": buf.size / dt / 1e6,
                "lines_per_s": n1 / dt, "roofline_frac": alg / dt / 1e9 / peak, "cpu_oracle_1thread_MBps": cpu,
                "ms": dt * 1e3})

    # ------------------------------------------------------------------ C3: Java multiline, ~2 KB records
    n3 = int((1 << 20) * args.scale)
    jb, nlines, nrec = synth.java_stack_records(n3)
    d_j = dput(jb)
    cap = nrec + 1024
    o3 = torch.empty(cap, dtype=torch.int32, device=dev)
    l3 = torch.empty(cap, dtype=torch.int32, device=dev)
    f3 = torch.empty(cap, dtype=torch.uint8, device=dev)
    start = lc.Regex(synth.JAVA_START_PATTERN)

    def c3():
        got["n3"], got["ctr"] = eng.multiline_split_dev(d_j.data_ptr(), jb.size, start, None, None, False, o3.data_ptr(),
                                                        l3.data_ptr(), f3.data_ptr(), cap)

    dt = timeit(c3)
    ns = min(jb.size, 64 << 20)
    cut = int(np.nonzero(jb[:ns] == 10)[0][-1]) + 1
    t0 = time.perf_counter()
    eo, el, ef, ectr = orc.multiline_split(jb[:cut], orc.Regex(synth.JAVA_START_PATTERN), None, None, False)
    cpu = cut / (time.perf_counter() - t0) / 1e6
    # parity on the sample prefix: the same records come first in the GPU output (last one may extend)
    g_off = o3[:len(eo) - 1].cpu().numpy().view(np.uint32)
    g_len = l3[:len(eo) - 1].cpu().numpy().view(np.uint32)
    assert np.array_equal(g_off, eo[:-1]) and np.array_equal(g_len, el[:-1])
    alg = jb.size + 8 * got["n3"] + nlines
    out.append({"config": "C3 ProcessorSplitMultilineLogStringNative Java start-pattern, %d records, %.0f B avg" %
                          (nrec, jb.size / nrec), "MBps": jb.size / dt / 1e6, "records_per_s": got["n3"] / dt,
                "lines_per_s": nlines / dt, "roofline_frac": alg / dt / 1e9 / peak, "cpu_oracle_1thread_MBps": cpu,
                "ms": dt * 1e3, "counters": [int(x) for x in got["ctr"]]})

    # ------------------------------------------------------------------ C4: delimiter -> regex chain, CSV
    n4 = int((8 << 20) * args.scale)
    cb, coff, clen = synth.csv_lines(n4)
    d_c = dput(cb)
    d_co = dput(coff.view(np.int32))
    d_cl = dput(clen.view(np.int32))
    MF = 11
    st4 = torch.empty(n4, dtype=torch.uint8, device=dev)
    nf4 = torch.empty(n4, dtype=torch.int32, device=dev)
    fo4 = torch.empty(n4 * MF, dtype=torch.int32, device=dev)
    fl4 = torch.empty(n4 * MF, dtype=torch.int32, device=dev)
    fd4 = torch.empty(n4 * MF, dtype=torch.int32, device=dev)
    rx = lc.Regex(synth.CSV_URL_PATTERN)
    G = rx.ngroups
    rs = torch.empty(n4, dtype=torch.uint8, device=dev)
    rco = torch.empty(n4 * G, dtype=torch.int32, device=dev)
    rcl = torch.empty(n4 * G, dtype=torch.int32, device=dev)

    def c4():
        eng.delim_parse_dev(d_c.data_ptr(), cb.size, d_co.data_ptr(), d_cl.data_ptr(), n4, b",", ord('"'), 10, True,
                            True, MF, st4.data_ptr(), nf4.data_ptr(), fo4.data_ptr(), fl4.data_ptr(), fd4.data_ptr())
        # field 3 (url) of every line feeds the regex: strided views of the field table are the event table
        uo = fo4.view(n4, MF)[:, 3].contiguous()
        ul = fl4.view(n4, MF)[:, 3].contiguous()
        eng.regex_parse_dev(rx, d_c.data_ptr(), cb.size, uo.data_ptr(), ul.data_ptr(), n4, G, rs.data_ptr(),
                            rco.data_ptr(), rcl.data_ptr())
        got["uo"], got["ul"] = uo, ul

    dt = timeit(c4)
    ns = min(n4, 1 << 17)
    t0 = time.perf_counter()
    est, enf, efo, efl, efd = orc.delim_parse_batch(cb, coff[:ns], clen[:ns], b",", ord('"'), 10, True, True, MF)
    rst, rco_e, rcl_e = orc.regex_parse_batch(orc.Regex(synth.CSV_URL_PATTERN), cb, efo[:, 3].copy(), efl[:, 3].copy(), G)
    cpu = int(clen[:ns].sum() + ns) / (time.perf_counter() - t0) / 1e6
    assert np.array_equal(st4[:ns].cpu().numpy(), est)
    assert np.array_equal(fo4.view(n4, MF)[:ns].cpu().numpy().view(np.uint32), efo)
    assert np.array_equal(rs[:ns].cpu().numpy(), rst)
    assert np.array_equal(rco.view(n4, G)[:ns].cpu().numpy().view(np.uint32), rco_e)
    alg = cb.size + n4 * (8 + 12 * MF + 5) + int(got["ul"].sum().item()) + n4 * (8 + 8 * G + 1)
    out.append({"config": "C4 ProcessorParseDelimiterNative -> ProcessorParseRegexNative chain, %d CSV lines" % n4,
                "MBps": cb.size / dt / 1e6, "lines_per_s": n4 / dt, "roofline_frac": alg / dt / 1e9 / peak,
                "cpu_oracle_1thread_MBps": cpu, "ms": dt * 1e3})
    # ------------------------------------------------------------------ C5: Zipf 64 B - 8 KB, nginx + apache patterns
    n5 = int((1 << 20) * args.scale)
    zb, zoff, zlen, zkind = synth.zipf_mixed_lines(n5)
    d_z = dput(zb)
    pats = [(synth.NGINX_PATTERN, ~zkind), (synth.APACHE_PATTERN, zkind)]
    work = []
    for pat, sel in pats:
        r = lc.Regex(pat)
        o_, l_ = np.ascontiguousarray(zoff[sel]), np.ascontiguousarray(zlen[sel])
        m = o_.size
        work.append((r, dput(o_.view(np.int32)), dput(l_.view(np.int32)), m,
                     torch.empty(m, dtype=torch.uint8, device=dev),
                     torch.empty(m * r.ngroups, dtype=torch.int32, device=dev),
                     torch.empty(m * r.ngroups, dtype=torch.int32, device=dev), pat, o_, l_))

    def c5():
        for r, do, dl, m, s_, co_, cl_, _, _, _ in work:
            eng.regex_parse_dev(r, d_z.data_ptr(), zb.size, do.data_ptr(), dl.data_ptr(), m, r.ngroups, s_.data_ptr(),
                                co_.data_ptr(), cl_.data_ptr())

    dt = timeit(c5, steps=3, warmup=1)
    alg = 0
    t_cpu = 0.0
    cpu_bytes = 0
    for r, do, dl, m, s_, co_, cl_, pat, o_, l_ in work:
        alg += int(l_.astype(np.int64).sum()) + m * (8 + 8 * r.ngroups + 1)
        ns = min(m, 4096)
        t0 = time.perf_counter()
        est, eco, ecl = orc.regex_parse_batch(orc.Regex(pat), zb, o_[:ns], l_[:ns], r.ngroups)
        t_cpu += time.perf_counter() - t0
        cpu_bytes += int(l_[:ns].astype(np.int64).sum())
        assert np.array_equal(s_[:ns].cpu().numpy(), est)
        assert np.array_equal(co_.view(m, r.ngroups)[:ns].cpu().numpy().view(np.uint32), eco)
        assert np.array_equal(cl_.view(m, r.ngroups)[:ns].cpu().numpy().view(np.uint32), ecl)
    out.append({"config": "C5 mixed nginx+apache regex, Zipf 64 B-8 KB lines, %d lines (%.0f B avg)" %
                          (n5, zb.size / n5), "MBps": zb.size / dt / 1e6, "lines_per_s": n5 / dt,
                "roofline_frac": alg / dt / 1e9 / peak, "cpu_oracle_1thread_MBps": cpu_bytes / t_cpu / 1e6,
                "ms": dt * 1e3})
    for o in out:
        print(json.dumps(o))
    eng.close()


if __name__ == "__main__":
    main()
