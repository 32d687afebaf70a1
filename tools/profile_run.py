#!/usr/bin/env python3
"""Small driver for ncu captures: one pass of a BASELINE config through the public API.
usage: python tools/profile_run.py <c2|c3> <replicates>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sagephy_b200 as sg

cfg, n = sys.argv[1], int(sys.argv[2])
ctx = sg.Context(0)
if cfg == "c2":
    app, args = "HostTreeGen", ["-s", "20260101", "-min", "100", "-max", "100", "-p", "0.8", "-a", "10000", "1.0", "5.0", "0.05", "bench"]
else:
    host = open(os.path.join(ROOT, "tests", "golden", "species100.pruned.tree")).read()
    app, args = "GuestTreeGen", ["-s", "20260101", host, "0.3", "0.3", "0.3", "bench"]
for rep in range(2):
    b = ctx.run(app, args, n=n, rng=sg.RNG_PHILOX, offset=rep * n, keep_on_device=True)
    t = b.timings(); s = b.summary()
    print({k: round(v, 3) for k, v in t.items()}, "ok", s.n_ok, "failed", s.n_failed, "bytes", int(s.out_bytes.sum()))
    b.free()
