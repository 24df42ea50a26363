#!/usr/bin/env python3
"""Turns an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the small JSON/markdown summaries committed
under profiles/.  Usage: python profiles/summarize.py <rep> <kernel-substring> <tag> [output directory]"""
import csv
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__f_wavefronts.sum", "l1tex__m_xbar2l1tex_read_sectors.sum"]


def main():
    rep, kern, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kern not in d.get("Kernel Name", ""):
            continue
        rec = {"kernel": d["Kernel Name"][:80]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                rec[h] = (r[i], units[i])
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                rec.setdefault("stalls_per_issue", {})[
                    h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = r[i]
        out.append(rec)
    here = sys.argv[4] if len(sys.argv) > 4 else os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, tag + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps([{k: v for k, v in r.items() if k in ("kernel", "gpu__time_duration.sum")} for r in out]))


if __name__ == "__main__":
    main()
