import numpy as np, sys
CTAS, ITERS, SLOTS = 320, 64, 16
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(CTAS, ITERS, SLOTS).astype(np.int64)
live = np.nonzero(a[:, 3, 0])[0]
t0=a[live][:,3:30,0]; t1=a[live][:,3:30,1]
end=(t1-t1.min(0)); dur=(t1-t0)
print(sys.argv[1], 'CTAs', len(live), 'tile phase mean %.1f us; end spread mean %.2f us; per-CTA mean lateness: first half %.2f second half %.2f; period %.1f' % (dur.mean()/1e3, end.max(0).mean()/1e3, end[:len(live)//2].mean()/1e3, end[len(live)//2:].mean()/1e3, (a[live][0,4:30,0]-a[live][0,3:29,0]).mean()/1e3))
