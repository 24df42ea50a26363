#!/usr/bin/env python
"""Summarise a LC_B200_SPLIT_TRACE dump of split_ws_kernel (debug aid), microseconds.
per tile: 0 iteration start, 1 bytes landed, 2 masks done (arrive full), 3 emit called, 4 prefix there (done), 5 emitted,
6 scanner passed full, 7 walk done"""
import sys
import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t = t[(t > 0).all(axis=1)]
walk_ns = t[:, 7] & 0xFFFFFFFF
rounds = (t[:, 7] >> 32) & 0xFF
repolls = t[:, 7] >> 40
t[:, 7] = t[:, 6]  # (scan time is not separated any more)
us = lambda a: a / 1e3
def line(n, c):
    print("  %-34s mean %6.2f  p50 %6.2f  p90 %6.2f  max %7.2f" % (n, c.mean(), np.median(c), np.percentile(c, 90), c.max()))
print("tiles %d, span %.1f us" % (len(t), us(t[:, 5].max() - t[:, 0].min())))
line("data: wait bytes", us(t[:, 1] - t[:, 0]))
line("data: masks -> arrive", us(t[:, 2] - t[:, 1]))
line("data: wait for the prefix", us(t[:, 4] - t[:, 3]))
line("data: emit", us(t[:, 5] - t[:, 4]))
line("masks done -> emit called (2 iter)", us(t[:, 3] - t[:, 2]))
line("scanner: full passed after arrive", us(t[:, 6] - t[:, 2]))
line("scanner: walk", us(walk_ns))
line("scanner: walk rounds", rounds.astype(float))
line("scanner: walk re-polls", repolls.astype(float))
line("full passed -> data sees prefix", us(t[:, 4] - t[:, 6]))
blk_iter = np.sort(t[:, 0])
