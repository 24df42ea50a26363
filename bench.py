#!/usr/bin/env python3
"""bench.py -- regex-parsed log throughput of the B200 engine on BASELINE.json's headline config.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the reference-side CPU arm: oracle restatement)

Workload (config.workload = "C2"): ProcessorParseRegexNative, nginx access-log regex with 10 capture
groups (docs/cn/plugins/processor/native/processor-parse-regex-native.md:49), 4 Mi lines x 256 B per GPU
(weak scaling: every rank parses its own shard, no collective on the data path).  One "step" = one pass
of the regex-parse hot path over the whole batch.

  value   input MB/s (1 MB = 1e6 B of log bytes, newline included) with the batch resident in HBM,
          CUDA-event timed on the engine's stream, max over ranks.
  e2e     the same metric through the host-pointer C-ABI call (lc_regex_parse): pinned host arena ->
          H2D -> kernel -> D2H of status + capture tables, every step.
  roofline.achieved = algorithmic bytes (read line bytes + 8 B line table + 8*G B captures + 1 B status per
          line) / mean device time of the regex kernel launch.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "regex_parsed_log_MBps"
UNIT = "MB/s"
LINES_PER_GPU = 4 * 1024 * 1024
LINE_BYTES = 256


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lines", type=int, default=LINES_PER_GPU, help="lines per GPU (default: the C2 size)")
    ap.add_argument("--cpu-sample-lines", type=int, default=131072)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu summary, if any."""
    p = os.path.join(ROOT, "profiles", "summary.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f).get("regex_kernel", {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def job_throughput(bytes_local, ms_local, world, device="cpu"):
    """Whole-job MB/s under weak scaling: every rank parses its own shard, the job finishes when the slowest
    rank does -> sum of the shard bytes / max over ranks of the step time.  Uses the default process group
    (NCCL on GPUs, gloo in the CPU tests); no data-path collective, just this 2-number reduction."""
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    b = torch.tensor([float(bytes_local)], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    return float(b.item()) / (ms * 1e-3) / 1e6, ms


def shard_seed(rank):
    """Shards are independent event groups: rank r generates (and owns) its own lines."""
    return 20260922 + rank


def make_workload(n_lines, seed):
    from loongcollector_b200 import synth
    return synth.nginx_lines(n_lines, seed=seed, line_bytes=LINE_BYTES)


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_regex_parse(buf, off, ln, threads, repeat=1):
    """The CPU path (oracle restatement, PCRE2 interpretive matcher, one matcher per thread -- mirrors
    process_thread_count / mReg[threadNo], ProcessorParseRegexNative.cpp:64-67).  Returns seconds."""
    from oracle import oracle as orc
    from loongcollector_b200 import synth
    rx = orc.Regex(synth.NGINX_PATTERN)
    L = orc.lib()
    n = off.size
    G = rx.ngroups
    status = np.zeros(n, np.uint8)
    co = np.zeros((n, G), np.uint32)
    cl = np.zeros((n, G), np.uint32)
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    matchers = [rx.new_matcher() for _ in range(threads)]

    def work(t):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b <= a:
            return
        for _ in range(repeat):
            L.orc_regex_parse_batch(matchers[t], buf.ctypes.data_as(ctypes.c_void_p),
                                    off[a:b].ctypes.data_as(ctypes.c_void_p), ln[a:b].ctypes.data_as(ctypes.c_void_p),
                                    b - a, len(synth.NGINX_KEYS), status[a:b].ctypes.data_as(ctypes.c_void_p),
                                    co[a:b].ctypes.data_as(ctypes.c_void_p), cl[a:b].ctypes.data_as(ctypes.c_void_p))

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    dt = time.perf_counter() - t0
    for m in matchers:
        L.orc_matcher_free(m)
    return dt, status


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_thread = 16384
    n = min(args.lines, cores * per_thread)
    buf, off, ln = make_workload(n, 20260922)
    in_bytes = int(n) * LINE_BYTES
    for _ in range(args.warmup):
        cpu_regex_parse(buf, off, ln, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_regex_parse(buf, off, ln, cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = in_bytes / dt / 1e6
    sample = "%d lines x %d B per step on %d threads (PCRE2 10.42 interpretive restatement of " \
             "ProcessorParseRegexNative; boost.regex is not installable here)" % (n, LINE_BYTES, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2: ProcessorParseRegexNative nginx 10-group regex, 256 B lines", "lines_per_step": n},
        "lines_per_s": n / dt,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import loongcollector_b200 as lc
    from loongcollector_b200 import synth

    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = lc.Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    rx = lc.Regex(synth.NGINX_PATTERN)
    G = rx.ngroups
    nkeys = len(synth.NGINX_KEYS)

    n = args.lines
    buf, off, ln = make_workload(n, shard_seed(rank))
    in_bytes = int(buf.size)
    # pinned host arena (the SourceBuffer stand-in) + pinned result tables
    L = lc.lib()

    def pinned(nbytes, dtype):
        p = L.lc_host_alloc(max(int(nbytes), 16))
        if not p:
            raise RuntimeError("lc_host_alloc failed")
        arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(int(nbytes),))
        return arr.view(dtype), p

    h_buf, p1 = pinned(buf.size, np.uint8)
    h_buf[:] = buf
    h_off, p2 = pinned(off.size * 4, np.uint32)
    h_off[:] = off
    h_len, p3 = pinned(ln.size * 4, np.uint32)
    h_len[:] = ln
    h_status, p4 = pinned(n, np.uint8)
    h_co, p5 = pinned(n * G * 4, np.uint32)
    h_cl, p6 = pinned(n * G * 4, np.uint32)

    with torch.cuda.stream(stream):
        d_buf = torch.empty(buf.size + 16, dtype=torch.uint8, device=dev)
        d_buf[:buf.size].copy_(torch.from_numpy(h_buf), non_blocking=True)
        d_off = torch.from_numpy(h_off.view(np.int32)).to(dev, non_blocking=True)
        d_len = torch.from_numpy(h_len.view(np.int32)).to(dev, non_blocking=True)
        d_status = torch.empty(n, dtype=torch.uint8, device=dev)
        d_co = torch.empty(n * G, dtype=torch.int32, device=dev)
        d_cl = torch.empty(n * G, dtype=torch.int32, device=dev)
    stream.synchronize()

    def step_dev():
        eng.regex_parse_dev(rx, d_buf.data_ptr(), in_bytes, d_off.data_ptr(), d_len.data_ptr(), n, nkeys,
                            d_status.data_ptr(), d_co.data_ptr(), d_cl.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        t_start.record(stream)
        for k in range(args.steps):
            ev[k][0].record(stream)
            step_dev()
            ev[k][1].record(stream)
        t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = eng.launches - launches0
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    # parity spot check of what the timed kernel produced (status histogram vs the generator's bad fraction)
    st = d_status.cpu().numpy()
    ok_lines = int((st == 0).sum())

    value, ms_per_step = job_throughput(in_bytes, total_ms / args.steps, world, dev)
    _, kern_ms = job_throughput(in_bytes, kern_ms, world, dev)

    # ---- end-to-end through the host-pointer C-ABI (pinned host arena in, result tables out)
    e2e = None
    if not args.no_e2e:
        def step_host():
            rc = L.lc_regex_parse(eng._h, rx._h, h_buf.ctypes.data_as(ctypes.c_void_p), in_bytes,
                                  h_off.ctypes.data_as(ctypes.c_void_p), h_len.ctypes.data_as(ctypes.c_void_p), n,
                                  nkeys, h_status.ctypes.data_as(ctypes.c_void_p), h_co.ctypes.data_as(ctypes.c_void_p),
                                  h_cl.ctypes.data_as(ctypes.c_void_p))
            if rc != 0:
                raise RuntimeError(L.lc_last_error().decode())

        step_host()
        barrier()
        e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            step_host()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e_steps
        e2e_value, e2e_ms = job_throughput(in_bytes, dt * 1e3, world, dev)
        dt = e2e_ms * 1e-3
        assert np.array_equal(h_status, st), "host-API result differs from device-API result"
        e2e = {"value": e2e_value, "unit": UNIT,
               "h2d_bytes_per_step": int(in_bytes + 8 * n), "d2h_bytes_per_step": int(n * (1 + 8 * G)),
               "ms_per_step": dt * 1e3, "steps": e_steps}

    # ---- CPU baseline (rank 0, N == 1 only): oracle port on a bounded sample, 1 thread
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = min(n, args.cpu_sample_lines)
        dt, cst = cpu_regex_parse(buf, off[:ns], ln[:ns], 1)
        assert np.array_equal(cst, st[:ns]), "GPU status differs from the CPU oracle on the sample"
        cpu = {"value": ns * LINE_BYTES / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d lines x %d B, 1 thread, PCRE2 10.42 interpretive (oracle restatement of "
                         "ProcessorParseRegexNative); host has %d cores" % (ns, LINE_BYTES, os.cpu_count() or 0)}

    peak, peak_src = hbm_peak()
    alg_bytes = int(ln.astype(np.int64).sum()) + n * (8 + 8 * G + 1)
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2: ProcessorParseRegexNative nginx 10-group regex, %d lines x %d B per GPU" %
                                   (n, LINE_BYTES), "lines_per_gpu": n, "line_bytes": LINE_BYTES,
                       "l2": "inputs (%.0f MB per GPU) larger than L2, no flush" % (in_bytes / 1e6),
                       "regex_tables": rx.info},
            "lines_per_s": n * world / (ms_per_step * 1e-3),
            "matched_lines_fraction": ok_lines / n,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profile_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }))
    eng.close()
    for p in (p1, p2, p3, p4, p5, p6):
        L.lc_host_free(p)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
