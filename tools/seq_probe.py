import sys, time
sys.path.insert(0, ".")
import sagephy_b200 as sg
from bench import C2_ARGS
import os
ROOT = "."
host = open(os.path.join(ROOT, "tests", "golden", "species100.pruned.tree")).read()
ctx = sg.Context(0)
def run(app, args, n, i):
    t = time.perf_counter(); b = ctx.run(app, args, n=n, rng=sg.RNG_PHILOX, offset=i * n, keep_on_device=True); dt = time.perf_counter() - t
    tm = b.timings(); print(app, i, "wall %.1f" % (dt * 1e3), "dev total %.1f" % tm["total_ms"], "phases %.1f" % sum(v for k, v in tm.items() if k.endswith("_ms") and k not in ("total_ms", "d2h_ms")), flush=True); b.free()
for i in range(2): run("HostTreeGen", C2_ARGS, 1000000, i)
c3 = ["-s", "20260101", "-db", "none", host, "0.3", "0.3", "0.3", "bench"]
for i in (10000, 10001, 10002, 0, 1, 2): run("GuestTreeGen", c3, 800000, i)
