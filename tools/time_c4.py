#!/usr/bin/env python3
"""Timing only (no oracle): BASELINE config C4 shape — GuestTreeGen with the three recipient-bias modes on a 1000-leaf species
tree, optionally DomainTreeGen over 100 gene trees. usage: python tools/time_c4.py [n_guest] [n_domain]"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sagephy_b200 as sg

n_guest = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n_domain = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = sg.Context(0)
tmp = tempfile.mkdtemp()
b = ctx.run("HostTreeGen", ["-s", "1", "-min", "1000", "-max", "1000", "-a", "100000", "1.0", "7.0", "0.05", "sp"], n=64, rng=sg.RNG_PHILOX)
host = os.path.join(tmp, "species1000.pruned.tree")
open(host, "wb").write(b.replicate(1, 0))
for db in (["-db", "exponential", "-dbr", "1.0"], ["-db", "simple"], ["-db", "none"]):
    args = ["-s", "5"] + db + ["-max", "100000", "-maxper", "100000", host, "0.3", "0.3", "0.3", "g"]
    for rep in range(2):
        t0 = time.time()
        g = ctx.run("GuestTreeGen", args, n=n_guest, rng=sg.RNG_PHILOX, keep_on_device=True, offset=rep * n_guest)
        dt = time.time() - t0
        st = g.rep_status
    print("GuestTreeGen", " ".join(db), f"n={n_guest}", "ok", int((st == 0).sum()), "mean leaves", round(float(g.sampled_leaves.mean()), 1), "sec", round(dt, 3),
          {k: round(v, 1) for k, v in g.timings().items() if k.endswith("_ms")}, "bytes", int(g.summary().out_bytes.sum()), flush=True)
    g.free()
if n_domain:
    gd = os.path.join(tmp, "genes"); os.makedirs(gd)
    gt = ctx.run("GuestTreeGen", ["-s", "77", "-db", "none", host, "0.2", "0.2", "0.1", "g"], n=100, rng=sg.RNG_PHILOX)
    for i in range(100):
        open(os.path.join(gd, f"gene{i:03d}.tree"), "wb").write(gt.replicate(1, i))
    dargs = ["-s", "9", "-ig", "0.5", "-is", "0.5", "-max", "100000", "-maxper", "100000", host, gd, "0.3", "0.3", "0.3", "d"]
    t0 = time.time()
    d = ctx.run("DomainTreeGen", dargs, n=n_domain, rng=sg.RNG_PHILOX, keep_on_device=True)
    st = d.rep_status
    print("DomainTreeGen 100 gene trees", f"n={n_domain}", "ok", int((st == 0).sum()), "mean leaves", round(float(d.sampled_leaves.mean()), 1), "sec", round(time.time() - t0, 2),
          {k: round(v, 1) for k, v in d.timings().items() if k.endswith("_ms")}, "bytes", int(d.summary().out_bytes.sum()), flush=True)
