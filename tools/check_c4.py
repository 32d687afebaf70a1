#!/usr/bin/env python3
"""BASELINE config C4 shape on one GPU: 1000-leaf species tree, distance-biased GuestTreeGen, then DomainTreeGen over 100 gene
trees; spot-checks replay parity against the oracle and prints timings. Also exercises the native CLI."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sagephy_b200 as sg
from oracle_util import oracle_run

ctx = sg.Context(0)
tmp = tempfile.mkdtemp()
# 1000-leaf host (P(N=1000) ~ 3.7e-4 -> ~2700 attempts)
t0 = time.time()
b = ctx.run("HostTreeGen", ["-s", "1", "-min", "1000", "-max", "1000", "-a", "100000", "1.0", "7.0", "0.05", "sp"], n=64, rng=sg.RNG_PHILOX)
print("HostTreeGen 1000 leaves: status", set(b.rep_status.tolist()), "attempts mean", b.attempts.mean(), "sec", round(time.time() - t0, 2), b.timings())
host = os.path.join(tmp, "species1000.pruned.tree")
open(host, "wb").write(b.replicate(1, 0))
for db in (["-db", "exponential", "-dbr", "1.0"], ["-db", "simple"], ["-db", "none"]):
    args = ["-s", "5"] + db + ["-max", "100000", "-maxper", "100000", host, "0.3", "0.3", "0.3", "g"]
    t0 = time.time()
    g = ctx.run("GuestTreeGen", args, n=20000, rng=sg.RNG_PHILOX, keep_on_device=True)
    st = g.rep_status
    print("GuestTreeGen", db, "n=20000 status counts", {int(k): int((st == k).sum()) for k in set(st.tolist())}, "mean leaves", g.sampled_leaves.mean(),
          "sec", round(time.time() - t0, 2), {k: round(v, 1) for k, v in g.timings().items()})
    # replay parity on 2 replicates (the oracle's exact-decimal recipient sampling over 1000 arcs is slow)
    r = ctx.run("GuestTreeGen", args, n=2, rng=sg.RNG_MT_REPLAY)
    for i in range(2):
        a2 = list(args); a2[1] = str(5 + i)
        want = oracle_run("GuestTreeGen", a2)["files"]
        for sfx, data in r.files(i).items():
            assert data == want["g" + sfx], (db, i, sfx)
    print("   replay parity OK on 2 replicates")
# DomainTreeGen over 100 gene trees
gd = os.path.join(tmp, "genes"); os.makedirs(gd)
gt = ctx.run("GuestTreeGen", ["-s", "77", "-db", "none", host, "0.2", "0.2", "0.1", "g"], n=100, rng=sg.RNG_PHILOX)
for i in range(100):
    open(os.path.join(gd, f"gene{i:03d}.tree"), "wb").write(gt.replicate(1, i))
dargs = ["-s", "9", "-ig", "0.5", "-is", "0.5", "-max", "100000", "-maxper", "100000", host, gd, "0.3", "0.3", "0.3", "d"]
t0 = time.time()
d = ctx.run("DomainTreeGen", dargs, n=5000, rng=sg.RNG_PHILOX, keep_on_device=True)
st = d.rep_status
print("DomainTreeGen 100 gene trees n=5000 status", {int(k): int((st == k).sum()) for k in set(st.tolist())}, "mean leaves", d.sampled_leaves.mean(), "sec", round(time.time() - t0, 2),
      {k: round(v, 1) for k, v in d.timings().items()})
r = ctx.run("DomainTreeGen", dargs, n=1, rng=sg.RNG_MT_REPLAY)
want = oracle_run("DomainTreeGen", dargs)["files"]
for sfx, data in r.files(0).items():
    assert data == want["d" + sfx], sfx
print("   DomainTreeGen replay parity OK on 1 replicate")
# native CLI writes the reference's files
cwd = os.path.join(tmp, "cli"); os.makedirs(cwd)
subprocess.check_call([os.path.join(ROOT, "sagephy_b200", "sagephy"), "HostTreeGen", "-s", "3", "-min", "5", "-max", "12", "1.0", "3.0", "0.3", "species"], cwd=cwd)
want = oracle_run("HostTreeGen", ["-s", "3", "-min", "5", "-max", "12", "1.0", "3.0", "0.3", "species"])["files"]
for name, data in want.items():
    assert open(os.path.join(cwd, name), "rb").read() == data, name
subprocess.check_call([os.path.join(ROOT, "sagephy_b200", "sagephy"), "GuestTreeGen", "-n", "3", "-s", "4", os.path.join(cwd, "species.pruned.tree"), "0.3", "0.2", "0.4", "gene"], cwd=cwd)
print("CLI files:", sorted(os.listdir(cwd)))
print("ALL OK")
