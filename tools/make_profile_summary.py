#!/usr/bin/env python3
"""Builds profiles/summary.json from the per-config `ncu --set full` summaries (profiles/<tag>_full_cN.json, written by
profiles/summarize.py from the captures of tools/profile_r02.sh):  python tools/make_profile_summary.py <tag> [cN=<tag of a later capture> ...]

Per config: the kernels of one step, their per-launch duration and DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum), the sum per step, and a hash of the kernel sources the capture was taken on -- bench.py reports
`roofline.traffic` from here and marks it stale when the sources have changed since."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["loongcollector_b200/csrc/lc_kernels.cu", "loongcollector_b200/csrc/lc_exec.cuh",
           "loongcollector_b200/csrc/lc_scan.cuh", "loongcollector_b200/csrc/lc_tables.h"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TUNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
STEP_KERNELS = {  # kernels of one step per config and how many launches of each a step holds
    "c1": {"split_mask_kernel": 1, "split_scan_kernel": 1, "split_emit_kernel": 1},
    "c2": {"regex_tdfa_staged_kernel": 1},
    "c3": {"split_mask_kernel": 1, "split_scan_kernel": 1, "split_emit_kernel": 1, "ml_pass_kernel<1>": 1,
           "ml_pass_kernel<2>": 1, "ml_pass_kernel<3>": 1},
    "c4": {"delim_tiled_kernel": 1, "regex_tdfa_staged_kernel": 1},
    "c5": {"regex_tdfa_multi_kernel": 8},
}


def main():
    tag = sys.argv[1]
    override = dict(a.split("=", 1) for a in sys.argv[2:])
    h = hashlib.sha256()
    for name in SOURCES:
        with open(os.path.join(ROOT, name), "rb") as f:
            h.update(f.read())
    out = {}
    for cfg, kernels in STEP_KERNELS.items():
        ctag = override.get(cfg, tag)
        p = os.path.join(ROOT, "profiles", "%s_full_%s.json" % (ctag, cfg))
        if not os.path.exists(p):
            continue
        recs = json.load(open(p))
        per = {}
        for r in recs:
            for kname in kernels:
                if kname in r["kernel"] and kname not in per:
                    rd, ru = r["dram__bytes_read.sum"]
                    wr, wu = r["dram__bytes_write.sum"]
                    t, tu = r["gpu__time_duration.sum"]
                    per[kname] = {"dram_bytes_per_launch": float(rd) * UNIT[ru] + float(wr) * UNIT[wu],
                                  "us_per_launch_under_ncu": float(t) * TUNIT[tu], "launches_per_step": kernels[kname]}
        if not per:
            continue
        out[cfg] = {"kernel": " + ".join(per), "kernels": per,
                    "dram_bytes_per_step": sum(v["dram_bytes_per_launch"] * v["launches_per_step"] for v in per.values()),
                    "capture": "profiles/%s_full_%s.json (ncu --set full --clock-control none)" % (ctag, cfg),
                    "sources": SOURCES, "sources_sha16": h.hexdigest()[:16]}
    with open(os.path.join(ROOT, "profiles", "summary.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: (v["kernel"], v["dram_bytes_per_step"]) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
