"""Reads a dump of scripts/phase_dump.py: per-phase times over CTAs and iterations."""
import sys
import numpy as np
CTAS, ITERS, SLOTS = 320, 64, 16
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(CTAS, ITERS, SLOTS).astype(np.int64)
live = np.nonzero(a[:, 3, 0])[0]
print("CTAs stamped:", len(live), "SMs:", len(set(a[live, 0, 7])))
its = range(3, 12)
t0 = a[live][:, :, 0]
names = ["tiles", "cta-sync", "cluster+ll", "gather", "canon+step", "sync"]
for k in range(6):
    d = (a[live][:, 3:12, k + 1] - a[live][:, 3:12, k]).astype(float)
    print("%-11s mean %7.0f  min %7.0f  max %7.0f ns (over CTAs x iterations)" % (names[k], d.mean(), d.min(), d.max()))
per = (a[live][:, 4:12, 0] - a[live][:, 3:11, 0]).astype(float)
print("period      mean %7.0f ns" % per.mean())
# skew of the iteration start and of the tile-phase end across CTAs
start = a[live][:, 3:12, 0]; end = a[live][:, 3:12, 1]
print("start skew  mean %7.0f ns   tile-end skew mean %7.0f ns (max - min over CTAs)" %
      ((start.max(0) - start.min(0)).mean(), (end.max(0) - end.min(0)).mean()))
g_end = a[live][:, 3:12, 4]
print("gather-end skew mean %.0f ns; last tile end -> first gather end: %.0f ns; -> last gather end %.0f ns" %
      ((g_end.max(0) - g_end.min(0)).mean(), (g_end.min(0) - end.max(0)).mean(), (g_end.max(0) - end.max(0)).mean()))
# cluster leaders (CTA % cluster == 0), all stamps relative to the END of the slowest CTA's tile phase
if len(sys.argv) > 2:
    C = int(sys.argv[2])
    lead = [c for c in live if c % C == 0]
    A = a[lead][:, 3:12, :].astype(float)
    last_tile_end = a[live][:, 3:12, 1].max(0).astype(float)
    labels = {1: "tiles end", 2: "cta sync", 8: "reduce entry", 3: "cluster reduced + LL stored", 4: "gathered",
              9: "totals broadcast (cluster barrier 2)", 10: "canonical", 11: "reduce returned", 12: "step entry",
              13: "stepped", 14: "state out", 5: "stamp 5", 6: "iteration end"}
    print("leaders: mean / min / max ns after the slowest CTA's tile end")
    for k in (1, 2, 8, 3, 4, 9, 10, 11, 12, 13, 14, 5, 6):
        d = A[:, :, k] - last_tile_end[None, :]
        print("  %-38s %7.0f %7.0f %7.0f" % (labels[k], d.mean(), d.min(), d.max()))
    nxt = a[live][:, 4:13, 0].astype(float) - last_tile_end[None, :]
    print("  %-38s %7.0f %7.0f %7.0f" % ("next iteration start (all CTAs)", nxt.mean(), nxt.min(), nxt.max()))
