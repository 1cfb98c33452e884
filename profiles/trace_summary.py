"""Summarise a BF_TRACE_FILE dump (build with -DBF_EXP_TRACE): per-CTA %globaltimer at the trace points of step T/2.
slots: 9 step start, 10 after begin_gather, 11 units done, 0 arrive sync A, 1 released, 12 after xbarrier 0,
       2 arrive sync B, 3 released, 13 after xbarrier 1, 4 arrive sync C1, 5 released, 6 arrive sync C2, 7 released, 14 after xbarrier 2"""
import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(1024, 16)
used = a[:, 9] > 0
a = a[used].astype(np.float64)
t0 = a[:, 9].min()
order = [9, 10, 11, 0, 1, 12, 2, 3, 13, 4, 5, 6, 7, 14]
names = ["start", "gather-setup", "units done", "arrive A", "released A", "xbar0", "arrive B", "released B", "xbar1", "arrive C1", "released C1",
         "arrive C2", "released C2", "xbar2"]
print(f"{used.sum()} CTAs; microseconds since the first CTA started the step: min / median / max")
for k, n in zip(order, names):
    v = (a[:, k] - t0) * 1e-3
    v = v[a[:, k] > 0]
    if v.size:
        print(f"  {n:14s} {v.min():8.1f} {np.median(v):8.1f} {v.max():8.1f}")
