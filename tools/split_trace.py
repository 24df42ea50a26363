#!/usr/bin/env python
"""Summarise a LC_B200_SPLIT_TRACE dump (debug aid): per-phase durations of split_kernel tiles, in microseconds.
usage: split_trace.py <trace.bin>"""
import sys
import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t = t[t[:, 0] > 0]
names = ["ticket", "load+mask", "block scan", "look-back", "stores", "probe/end"]
d = np.diff(t[:, :7], axis=1) / 1e3
print("tiles %d, span %.1f us, tile lifetime mean %.2f us" % (len(t), (t[:, 6].max() - t[:, 0].min()) / 1e3,
                                                             (t[:, 6] - t[:, 0]).mean() / 1e3))
for k, n in enumerate(names):
    c = d[:, k]
    print("  %-11s mean %6.2f  p50 %6.2f  p90 %6.2f  max %7.2f" % (n, c.mean(), np.median(c), np.percentile(c, 90), c.max()))
first = t[:296]
rest = t[296:]
if len(rest):
    print("  first wave lifetime %.2f us, later tiles %.2f us" % ((first[:, 6] - first[:, 0]).mean() / 1e3,
                                                                  (rest[:, 6] - rest[:, 0]).mean() / 1e3))
    print("  later tiles look-back mean %.2f us" % (np.diff(rest[:, 3:5], axis=1).mean() / 1e3))
