#!/usr/bin/env python3
"""DomainTreeGen over 100 gene trees on the 100-leaf species fixture (bounded check; no oracle involved)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sagephy_b200 as sg
ctx = sg.Context(0)
host = os.path.join(ROOT, "tests", "golden", "species100.pruned.tree")
tmp = tempfile.mkdtemp(); gd = os.path.join(tmp, "genes"); os.makedirs(gd)
gt = ctx.run("GuestTreeGen", ["-s", "77", "-db", "none", host, "0.2", "0.2", "0.1", "g"], n=100, rng=sg.RNG_PHILOX)
for i in range(100):
    open(os.path.join(gd, f"gene{i:03d}.tree"), "wb").write(gt.replicate(1, i))
for n in (1000, 20000):
    t0 = time.time()
    d = ctx.run("DomainTreeGen", ["-s", "9", "-ig", "0.5", "-is", "0.5", host, gd, "0.3", "0.3", "0.3", "d"], n=n, rng=sg.RNG_PHILOX, keep_on_device=True)
    st = d.rep_status
    print("DomainTreeGen 100 genes n=%d" % n, {int(k): int((st == k).sum()) for k in set(st.tolist())}, "mean leaves %.1f" % d.sampled_leaves.mean(), "attempts %.2f" % d.attempts.mean(),
          "sec %.2f" % (time.time() - t0), {k: round(v, 1) for k, v in d.timings().items()})
