#!/usr/bin/env python
"""Summarise a LC_B200_SPLIT_TRACE dump of split_pipe_kernel (debug aid), microseconds per iteration phase."""
import sys
import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t = t[t[:, 0] > 0]
names = ["wait bytes", "masks+fetch", "barrier A", "scan+walk (B)", "emit"]
d = np.diff(t[:, :6], axis=1) / 1e3
print("tiles %d, span %.1f us, iteration mean %.2f us, walk rounds mean %.2f max %d" %
      (len(t), (t[:, 5].max() - t[:, 0].min()) / 1e3, (t[:, 5] - t[:, 0]).mean() / 1e3, t[:, 6].mean(), t[:, 6].max()))
for k, n in enumerate(names):
    c = d[:, k]
    print("  %-13s mean %6.2f  p50 %6.2f  p90 %6.2f  max %7.2f" % (n, c.mean(), np.median(c), np.percentile(c, 90), c.max()))
