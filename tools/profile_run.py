#!/usr/bin/env python3
"""Small driver for ncu captures: one pass of a BASELINE config through the public API.
usage: python tools/profile_run.py <c2|c3|c5ln|c5gamma|chain> <replicates>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sagephy_b200 as sg

cfg, n = sys.argv[1], int(sys.argv[2])
ctx = sg.Context(0)
if cfg == "chain":
    # C2 trees -> BranchRelaxer on the device (no host round trip), 2 lognormal draws per tree
    import time
    gen = ["-s", "20260101", "-min", "100", "-max", "100", "-p", "0.8", "-a", "10000", "1.0", "5.0", "0.05", "bench"]
    for rep in range(2):
        t0 = time.perf_counter()
        src = ctx.run("HostTreeGen", gen, n=n, rng=sg.RNG_PHILOX, offset=rep * n, keep_on_device=True, flags=sg.FLAG_KEEP_TREES, stream_mask=2)
        t1 = time.perf_counter()
        for model in (["IIDLogNormal", "0", "0.25"], ["IIDGamma", "2", "0.5"]):
            t2 = time.perf_counter()
            out = ctx.relax_on_batch(src, ["-s", "7", "chained"] + model, n=2 * n, rng=sg.RNG_PHILOX, keep_on_device=True)
            t3 = time.perf_counter()
            t = out.timings(); s = out.summary()
            print(model[0], "generate_s", round(t1 - t0, 3), "relax_s", round(t3 - t2, 3), {k: round(v, 2) for k, v in t.items()}, "ok", s.n_ok, "failed", s.n_failed, "bytes", int(s.out_bytes.sum()))
            out.free()
        src.free()
    sys.exit(0)
if cfg == "c2":
    app, args = "HostTreeGen", ["-s", "20260101", "-min", "100", "-max", "100", "-p", "0.8", "-a", "10000", "1.0", "5.0", "0.05", "bench"]
elif cfg in ("c5ln", "c5gamma"):
    tree = open(os.path.join(ROOT, "tests", "golden", "species100.pruned.tree")).read().strip()
    app, args = "BranchRelaxer", ["-s", "20260101", tree] + (["IIDLogNormal", "0", "0.25"] if cfg == "c5ln" else ["IIDGamma", "2", "0.5"])
else:
    host = open(os.path.join(ROOT, "tests", "golden", "species100.pruned.tree")).read()
    extra = sys.argv[3:]          # e.g. -db none, -rt 0.0
    app, args = "GuestTreeGen", ["-s", "20260101"] + extra + [host, "0.3", "0.3", "0.3", "bench"]
for rep in range(2):
    b = ctx.run(app, args, n=n, rng=sg.RNG_PHILOX, offset=rep * n, keep_on_device=True)
    t = b.timings(); s = b.summary()
    print({k: round(v, 3) for k, v in t.items()}, "ok", s.n_ok, "failed", s.n_failed, "bytes", int(s.out_bytes.sum()))
    b.free()
