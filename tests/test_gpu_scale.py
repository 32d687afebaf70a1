"""GPU parity tests at the scale of the BASELINE configs (C2, C3, C4, C5) and the pins that do not depend on the oracle.

Everything goes through the C ABI; the CPU oracle (a process pool over the box's host cores) is the checker only.
Bar: byte-exact in MT seed-replay mode; two-sample KS / chi-square p > 0.01 over 10^5 replicates in Philox mode
(BASELINE.json north_star: leaf counts, tree heights, per-type event frequencies)."""
import os
import re

import numpy as np
import pytest
from scipy import stats

import sagephy_b200 as sg
from oracle_util import oracle_batch_parallel, oracle_mt_words, oracle_run, oracle_run_many
from test_gpu_parity import HOST8, assert_replay_parity, chi2_two_sample_ok, ks_ok, with_seed

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPECIES100 = os.path.join(ROOT, "tests", "golden", "species100.pruned.tree")


def assert_same_files(got, want, prefix, what):
    assert set(prefix + k for k in got) == set(want), (what, sorted(got), sorted(want))
    for sfx, data in got.items():
        exp = want[prefix + sfx]
        if data != exp:
            k = next((j for j in range(min(len(data), len(exp))) if data[j] != exp[j]), min(len(data), len(exp)))
            raise AssertionError(f"{what} stream {sfx}: first difference at byte {k}:\n gpu   : {data[max(0, k - 60):k + 60]!r}\n oracle: {exp[max(0, k - 60):k + 60]!r}")


def replay_parity_pool(ctx, app, args, n, prefix, no_recipient_is_error=False):
    """assert_replay_parity with the oracle runs spread over the host cores."""
    si = (args.index("-s") if "-s" in args else args.index("--seed")) + 1
    seed0 = int(args[si])
    b = ctx.run(app, args, n=n, rng=sg.RNG_MT_REPLAY)
    wants = oracle_run_many([(app, with_seed(args, seed0 + i)) for i in range(n)])
    status, attempts = b.rep_status, b.attempts
    streams = {s: (b.stream(s), b.offsets(s)) for s in range(8) if b.stream_suffix(s)}
    for i, want in enumerate(wants):
        if want["status"] == 1:
            assert status[i] == 1, (i, status[i])
            continue
        if want["status"] == 2:
            assert status[i] == 4, (i, status[i], want["stderr"][:200])
            continue
        if no_recipient_is_error and status[i] == 4:
            # the reference carries on here with a stale arc left behind by an earlier transfer of the same epoch (Epoch.transferedToArc
            # is only updated on success, GuestTreeInHostTreeCreator.java:194-195); the product refuses the replicate instead
            assert want["no_recipient"] > 0, i
            continue
        assert status[i] == 0 and attempts[i] == want["attempts"], (app, i, status[i], attempts[i], want["attempts"])
        got = {b.stream_suffix(s): data[int(off[i]):int(off[i + 1])] for s, (data, off) in streams.items()}
        assert_same_files(got, want["files"], prefix, f"{app} replicate {i}")
    return b, wants


# ---- pins that do not involve the oracle -------------------------------------------------------------------------------------------
def test_philox4x32_10_known_answers_on_device(ctx):
    # Random123 kat_vectors (philox4x32, 10 rounds): counter words, key words -> output words
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for variant in (0, 1):              # 0: key schedule computed per block, 1: precomputed round keys (the find kernel's form)
        ctr = np.array([k[0] for k in kat], dtype=np.uint32); key = np.array([k[1] for k in kat], dtype=np.uint32)
        out = ctx.philox(ctr, key, variant)
        assert out.tolist() == [list(k[2]) for k in kat], variant
    # the two forms agree with each other and with the host build on random inputs
    rng = np.random.default_rng(1)
    ctr = rng.integers(0, 2 ** 32, (5000, 4), dtype=np.uint64).astype(np.uint32); key = rng.integers(0, 2 ** 32, (5000, 2), dtype=np.uint64).astype(np.uint32)
    a, b2, h = ctx.philox(ctr, key, 0), ctx.philox(ctr, key, 1), sg.philox_host(ctr, key)
    assert (a == b2).all() and (a == h).all()


@pytest.mark.parametrize("seed", [(1 << 70) + 5, 10 ** 50 + 12345, -(10 ** 40), 2 ** 127 - 2, 2 ** 120 + 2 ** 64 - 2, -(2 ** 63) - 1])
def test_big_integer_seeds_replay(ctx, seed):
    # math/PRNG.java:29-37,58-60: any decimal BigInteger, low 16 bytes; replicate i == run with -s seed+i in big-integer arithmetic
    for rep in (0, 3):
        assert (ctx.mt_words(seed, 700, replicate=rep) == oracle_mt_words(seed + rep, 700)).all()
    args = ["-s", str(seed), "-min", "3", "-max", "20", "-p", "0.8", "1.0", "2.5", "0.4", "o"]
    b = ctx.run("HostTreeGen", args, n=6, rng=sg.RNG_MT_REPLAY)
    for i in range(6):
        want = oracle_run("HostTreeGen", with_seed(args, seed + i))
        assert b.rep_status[i] == 0 and b.attempts[i] == want["attempts"]
        assert_same_files(b.files(i), want["files"], "o", f"seed {seed} replicate {i}")
        assert f"-s, {seed + i}, ".encode() in b.files(i)[".pruned.info"]
    r = ctx.run("BranchRelaxer", ["-s", str(seed), "-x", "((A:0.5,B:0.5)X:0.3,C:0.8)R:0.1;", "IIDLogNormal", "0", "0.25"], n=4, rng=sg.RNG_MT_REPLAY)
    assert r.stdout == "".join(oracle_run("BranchRelaxer", ["-s", str(seed + i), "-x", "((A:0.5,B:0.5)X:0.3,C:0.8)R:0.1;", "IIDLogNormal", "0", "0.25"])["stdout"] for i in range(4))


def test_partially_tagged_host_tree_has_no_recipient(ctx):
    # A host tree that carries meta tags on only some vertices gives its vertices different HOST tags (io/NewickTreeReader.java:215-225):
    # Epoch.sampleArc then finds no recipient and returns -5 (topology/Epoch.java:256-275), on which the reference dies (or spins).
    # The product must end such replicates with a model error and never index anything with -5.
    host = "((A:0.4,B:0.4):0.6,(C:0.7[x=1],D:0.7[x=1]):0.3):0.2;"
    for db in ("none", "simple"):
        args = ["-s", "1", "-db", db, "-rt", "0.6", host, "0.5", "0.3", "1.5", "o"]
        b, wants = replay_parity_pool(ctx, "GuestTreeGen", args, 48, "o", no_recipient_is_error=True)
        n_err = int((b.rep_status == 4).sum())
        assert 0 < n_err < 48 and sum(1 for w in wants if w["status"] == 2) > 0, n_err
        p = ctx.run("GuestTreeGen", args, n=4000, rng=sg.RNG_PHILOX, stream_mask=0b0010)
        assert set(np.unique(p.rep_status)) <= {0, 4} and (p.rep_status == 4).any() and (p.rep_status == 0).any()


# ---- C2: HostTreeGen at the benchmark's parameters -----------------------------------------------------------------------------------
def test_c2_replay_subset_of_1000(ctx):
    # MT-replay subset of the C2 workload, per-replicate seed s + i; ~276 attempts per accepted tree
    args = ["-s", "20260101", "-min", "100", "-max", "100", "-p", "0.8", "-a", "10000", "1.0", "5.0", "0.05", "c2"]
    b, wants = replay_parity_pool(ctx, "HostTreeGen", args, 1024, "c2")
    assert (b.sampled_leaves[b.rep_status == 0] == 100).all() and (b.rep_status == 0).sum() >= 1000


def test_c2_philox_matches_oracle_distributions(ctx):
    # same workload in throughput mode against 4096 oracle replicates: attempts (geometric, mean 1/P(accept) + 1), unpruned size,
    # pruned-root height, total branch time, births / deaths / unsampled leaves
    args = ["-s", "20260101", "-min", "100", "-max", "100", "-p", "0.8", "-a", "10000", "1.0", "5.0", "0.05", "c2"]
    n = 100000
    b = ctx.run("HostTreeGen", args, n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    ok = b.rep_status == 0
    assert ok.mean() > 0.999
    st = oracle_batch_parallel("HostTreeGen", args, 0, 4096)
    so = st[:, 1] == 0
    assert ks_ok(b.attempts[ok].astype(float), st[so, 0])
    assert ks_ok(b.root_times[ok], st[so, 5]) and ks_ok(b.total_times[ok], st[so, 6])
    ev = b.event_counts[ok]
    for e in (3, 4, 5):
        assert chi2_two_sample_ok(ev[:, e].astype(float), st[so, 9 + e]), e


# ---- C3: GuestTreeGen on the 100-leaf species tree -------------------------------------------------------------------------------------
@pytest.mark.parametrize("db", ["none", "simple"])
def test_c3_guest_tree_gen_ks_at_1e5(ctx, db):
    # 10^5 Philox replicates on the GPU against 8192 oracle replicates (the oracle needs ~25 ms per tree on this host tree)
    n, n_ref = 100000, 8192
    args = ["-s", "424242", "-db", db, SPECIES100, "0.3", "0.3", "0.3", "x"]
    b = ctx.run("GuestTreeGen", args, n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    ok = b.rep_status == 0
    assert ok.all()
    st = oracle_batch_parallel("GuestTreeGen", args, 0, n_ref)
    assert (st[:, 1] == 0).all()
    assert chi2_two_sample_ok(b.sampled_leaves.astype(float), st[:, 4])
    assert ks_ok(b.root_times, st[:, 5])                       # height of the pruned tree's root
    assert ks_ok(b.total_times, st[:, 6])
    ev = b.event_counts
    for e in (0, 2, 3, 4, 5, 6, 7, 8):
        assert chi2_two_sample_ok(ev[:, e].astype(float), st[:, 9 + e]), (db, e)
    assert chi2_two_sample_ok(b.attempts.astype(float), st[:, 0])


def test_guest_and_domain_small_host_ks_at_1e5(ctx, tmp_path):
    # 10^5 replicates on the 8-leaf host, with the pruned-root height among the statistics (both apps)
    n = 100000
    args = ["-s", "777", "-db", "simple", HOST8, "0.5", "0.4", "0.7", "x"]
    b = ctx.run("GuestTreeGen", args, n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    st = oracle_batch_parallel("GuestTreeGen", args, 0, n)
    assert (b.rep_status == 0).all() and (st[:, 1] == 0).all()
    assert chi2_two_sample_ok(b.sampled_leaves.astype(float), st[:, 4]) and ks_ok(b.root_times, st[:, 5]) and ks_ok(b.total_times, st[:, 6])
    for e in (0, 4, 5, 6, 7, 8):
        assert chi2_two_sample_ok(b.event_counts[:, e].astype(float), st[:, 9 + e]), e
    d = tmp_path / "genes"; d.mkdir()
    for i in range(1, 5):
        (d / f"gene{i}.tree").write_bytes(oracle_run("GuestTreeGen", ["-s", str(i), "-db", "none", HOST8, "0.4", "0.2", "0.2", "g"])["files"]["g.pruned.tree"])
    dargs = ["-s", "31", "-ig", "0.6", "-is", "0.5", HOST8, str(d), "0.4", "0.3", "0.8", "x"]
    b = ctx.run("DomainTreeGen", dargs, n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    st = oracle_batch_parallel("DomainTreeGen", dargs, 0, n)
    assert (b.rep_status == 0).all() and (st[:, 1] == 0).all()
    assert chi2_two_sample_ok(b.sampled_leaves.astype(float), st[:, 4]) and ks_ok(b.root_times, st[:, 5]) and ks_ok(b.total_times, st[:, 6])
    for e in (1, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16):
        assert chi2_two_sample_ok(b.event_counts[:, e].astype(float), st[:, 9 + e]), e
    assert chi2_two_sample_ok(b.attempts.astype(float), st[:, 0])


# ---- C4: 1000-leaf species tree, distance-biased transfers, DomainTreeGen over 100 gene trees --------------------------------------------
@pytest.fixture(scope="module")
def species1000(tmp_path_factory):
    d = tmp_path_factory.mktemp("c4")
    r = oracle_run("HostTreeGen", ["-s", "1", "-min", "1000", "-max", "1000", "-a", "100000", "1.0", "7.0", "0.05", "species1000"])
    assert r["status"] == 0
    p = d / "species1000.pruned.tree"
    p.write_bytes(r["files"]["species1000.pruned.tree"])
    return str(p)


@pytest.mark.parametrize("db", [["-db", "exponential", "-dbr", "1.0"], ["-db", "simple"], ["-db", "none"]], ids=["exponential", "simple", "none"])
def test_c4_guest_tree_gen_replay_on_1000_leaf_host(ctx, species1000, db):
    args = ["-s", "5"] + db + ["-max", "100000", "-maxper", "100000", species1000, "0.3", "0.3", "0.3", "g"]
    b, _ = replay_parity_pool(ctx, "GuestTreeGen", args, 6, "g")
    assert (b.rep_status == 0).all() and b.sampled_leaves.mean() > 300


def test_c4_domain_tree_gen_over_100_gene_trees_replay(ctx, species1000, tmp_path):
    gd = tmp_path / "genes"; gd.mkdir()
    genes = oracle_run_many([("GuestTreeGen", ["-s", str(77 + i), "-db", "none", species1000, "0.2", "0.2", "0.1", "g"]) for i in range(100)])
    for i, g in enumerate(genes):
        (gd / f"gene{i:03d}.tree").write_bytes(g["files"]["g.pruned.tree"])
    dargs = ["-s", "9", "-ig", "0.5", "-is", "0.5", "-max", "100000", "-maxper", "100000", species1000, str(gd), "0.3", "0.3", "0.3", "d"]
    b, _ = replay_parity_pool(ctx, "DomainTreeGen", dargs, 3, "d")
    assert (b.rep_status == 0).all()
    # throughput mode at this scale: every replicate valid, emitted trees well formed
    p = ctx.run("DomainTreeGen", dargs, n=256, rng=sg.RNG_PHILOX, stream_mask=0b0011)
    assert (p.rep_status == 0).all()
    t = p.replicate(1, 0).decode(); bare = re.sub(r"\[[^\]]*\]", "", t)
    assert bare.count("(") == bare.count(")") and bare.rstrip().endswith(";")


# ---- C5: BranchRelaxer over 199-vertex trees ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2_pruned_tree():
    r = oracle_run("HostTreeGen", ["-s", "7", "-min", "100", "-max", "100", "-p", "0.8", "1.0", "5.0", "0.05", "c2"])
    t = r["files"]["c2.pruned.tree"].decode().strip()
    assert re.sub(r"\[[^\]]*\]", "", t).count(",") == 99
    return t


@pytest.mark.parametrize("model", [["IIDLogNormal", "0", "0.25"], ["IIDGamma", "2", "0.5"], ["ACTK98", "1.0", "0.5"]], ids=["lognormal", "gamma", "actk98"])
def test_c5_branch_relaxer_replay_on_199_vertex_tree(ctx, c2_pruned_tree, model):
    for extra in ([], ["-x", "-innms"]):
        args = ["-s", "500"] + extra + ["-o", "rel.tree", c2_pruned_tree] + model
        n = 12
        b = ctx.run("BranchRelaxer", args, n=n, rng=sg.RNG_MT_REPLAY)
        wants = oracle_run_many([("BranchRelaxer", with_seed(args, 500 + i)) for i in range(n)])
        for i, want in enumerate(wants):
            if want["status"] == 2:
                assert b.rep_status[i] == 4
                continue
            assert b.rep_status[i] == 0 and b.attempts[i] == want["attempts"]
            assert b.files(i)[""] == want["files"]["rel.tree"] and b.files(i)[".info"] == want["files"]["rel.tree.info"]


LEN_RE = re.compile(r":([0-9][0-9.eE+-]*)")


@pytest.mark.parametrize("model", [["IIDLogNormal", "0", "0.25"], ["IIDGamma", "2", "0.5"], ["IIDNormal", "1.0", "0.01"], ["IIDExponential", "2.0"]],
                         ids=["lognormal", "gamma", "normal", "exponential"])
def test_c5_branch_relaxer_philox_against_the_oracle(ctx, c2_pruned_tree, model):
    # throughput mode against the ORACLE's relaxed lengths (parsed from its Newick output), vertex by vertex in emission order
    n, n_ref = 20000, 3000
    args = ["-s", "99", c2_pruned_tree] + model
    b = ctx.run("BranchRelaxer", args, n=n, rng=sg.RNG_PHILOX)
    ok = b.rep_status == 0
    assert ok.mean() > 0.99
    rel = b.values(1).reshape(n, 199)[ok]
    wants = oracle_run_many([("BranchRelaxer", with_seed(args, 99 + i)) for i in range(n_ref)])
    ref = np.array([[float(x) for x in LEN_RE.findall(w["stdout"])] for w in wants if w["status"] == 0 and w["stdout"]])
    assert ref.shape[1] == 199 and ref.shape[0] > 0.99 * n_ref
    for x in (0, 57, 120, 197):
        assert ks_ok(rel[:, x], ref[:, x]), (model, x)
    # all vertices pooled: rate = relaxed / original length
    orig = np.array([float(x) for x in LEN_RE.findall(re.sub(r"\[[^\]]*\]", "", c2_pruned_tree))])
    assert orig.shape == (199,) and (orig > 0).all()
    assert ks_ok((rel / orig)[::3, ::7].ravel(), (ref / orig)[:, ::7].ravel())


# ---- multi-GPU front end of the C ABI / CLI: shards by index range == one device, byte for byte ----------------------------
def test_app_run_multi_equals_single_device(ctx, tmp_path):
    import subprocess
    import torch
    ndev = torch.cuda.device_count()
    devices = [0, 1 % ndev, 0]                    # three shards; on a one-GPU box all three contexts share device 0
    m = sg.MultiContext(devices)
    try:
        for app, args, rng in (("HostTreeGen", ["-s", "11", "-min", "5", "-max", "40", "-p", "0.8", "1.0", "3.0", "0.3", str(tmp_path / "m")], sg.RNG_PHILOX),
                               ("GuestTreeGen", ["-s", "12", "-db", "simple", HOST8, "0.5", "0.4", "0.7", str(tmp_path / "g")], sg.RNG_MT_REPLAY)):
            n = 1000
            shards, summary = m.run(app, args, n=n, rng=rng, offset=7)
            one = ctx.run(app, args, n=n, rng=rng, offset=7)
            assert [b.n for b in shards] == [333, 333, 334]
            for s in range(8):
                assert b"".join(b.stream(s) for b in shards) == one.stream(s), (app, s)
            assert np.concatenate([b.attempts for b in shards]).tolist() == one.attempts.tolist()
            assert (summary.as_vector() == one.summary().as_vector()).all()
            # packed files of the shards == packed files of the single batch (same .idx, run-wide offsets)
            m.write_files(shards, sg.LAYOUT_PACKED, first_index=7)
            multi_files = {p.name: p.read_bytes() for p in tmp_path.iterdir() if p.is_file()}
            for p in tmp_path.iterdir():
                if p.is_file():
                    p.unlink()
            one.write_files(layout=sg.LAYOUT_PACKED, first_index=7)
            single_files = {p.name: p.read_bytes() for p in tmp_path.iterdir() if p.is_file()}
            assert multi_files == single_files and len(single_files) >= 8
            for p in tmp_path.iterdir():
                if p.is_file():
                    p.unlink()
    finally:
        m.close()
    # the native CLI: `--gpus 2` (two shards; SAGEPHY_DEVICES maps both to device 0 when the box has one GPU) == one device
    exe = os.path.join(ROOT, "sagephy_b200", "sagephy")
    core = ["HostTreeGen", "-n", "301", "-rng", "mt-replay", "-s", "3", "-min", "5", "-max", "12", "1.0", "3.0", "0.3", "species"]
    d1, d2 = tmp_path / "one", tmp_path / "two"; d1.mkdir(); d2.mkdir()
    subprocess.check_call([exe] + core, cwd=d1)
    env = dict(os.environ)
    if ndev < 2:
        env["SAGEPHY_DEVICES"] = "0,0"
    subprocess.check_call([exe] + core + ["--gpus", "2"], cwd=d2, env=env)
    f1 = {p.name: p.read_bytes() for p in d1.iterdir()}; f2 = {p.name: p.read_bytes() for p in d2.iterdir()}
    assert f1 == f2 and "species.pruned.tree.idx" in f1


# ---- IIDRK11 ------------------------------------------------------------------------------------------------------
RK_HOST = "((A:0.4[PARAMS=(2.0,0.5)],B:0.4[PARAMS=(3.0,0.4)]):0.6[PARAMS=(1.5,0.7)],(C:0.7[PARAMS=(2.5,0.3)],(D:0.2[PARAMS=(4,0.25)],E:0.2[PARAMS=(2,0.6)]):0.5[PARAMS=(1.0,1.0)]):0.3[PARAMS=(2.0,0.5)]);"
RK_HOST_PLAIN = "((A:0.4,B:0.4):0.6,(C:0.7,(D:0.2,E:0.2):0.5):0.3):0.2;"


def test_iidrk11_replay_and_distribution(ctx, tmp_path):
    # guest trees without transfers from the oracle's GuestTreeGen on the same species tree, their .leafmap as the guest-to-host map
    # (apps/SaGePhy/IIDRasmussenKellis11RateModel.java; BranchRelaxerParameters.java:108-113)
    n_checked = 0
    for gs in range(1, 9):
        g = oracle_run("GuestTreeGen", ["-s", str(gs), "-db", "none", "-min", "3", RK_HOST_PLAIN, "0.5", "0.3", "0.0", "g"])["files"]
        gt = tmp_path / f"g{gs}.tree"; gt.write_bytes(g["g.pruned.tree"]); gm = tmp_path / f"g{gs}.map"; gm.write_bytes(g["g.pruned.leafmap"])
        for extra in ([], ["-x", "-innms"]):
            args = ["-s", "40"] + extra + ["-o", "rk.tree", str(gt), "IIDRK11", RK_HOST, str(gm), "1.5"]
            n = 12
            b = ctx.run("BranchRelaxer", args, n=n, rng=sg.RNG_MT_REPLAY)
            for i in range(n):
                want = oracle_run("BranchRelaxer", with_seed(args, 40 + i))
                if want["status"] == 2:
                    assert b.rep_status[i] == 4
                    continue
                assert b.rep_status[i] == 0 and b.attempts[i] == want["attempts"], (gs, i, b.rep_status[i])
                assert b.files(i)[""] == want["files"]["rk.tree"] and b.files(i)[".info"] == want["files"]["rk.tree.info"], (gs, extra, i)
                n_checked += 1
    assert n_checked > 150
    # retries: a tight rate window
    args = ["-s", "7", "-min", "0.3", "-max", "3.0", "-a", "40", "-o", "rk.tree", str(tmp_path / "g1.tree"), "IIDRK11", RK_HOST, str(tmp_path / "g1.map"), "1.0"]
    b = ctx.run("BranchRelaxer", args, n=24, rng=sg.RNG_MT_REPLAY)
    for i in range(24):
        want = oracle_run("BranchRelaxer", with_seed(args, 7 + i))
        if want["status"] == 1:
            assert b.rep_status[i] == 1
        elif want["status"] == 0:
            assert b.rep_status[i] == 0 and b.attempts[i] == want["attempts"] and b.files(i)[""] == want["files"]["rk.tree"]
    # constructor errors of the reference refuse the run
    (tmp_path / "short.map").write_text("G0_3 C\n")
    for bad in (["-s", "1", str(tmp_path / "g1.tree"), "IIDRK11", RK_HOST, str(tmp_path / "short.map"), "1.0"],
                ["-s", "1", str(tmp_path / "g1.tree"), "IIDRK11", RK_HOST_PLAIN, str(tmp_path / "g1.map"), "1.0"]):
        assert oracle_run("BranchRelaxer", bad)["status"] == 2
        with pytest.raises(sg.SagephyError):
            ctx.run("BranchRelaxer", bad, n=2, rng=sg.RNG_MT_REPLAY)
    # throughput mode: same conditional distributions as the MT oracle
    args = ["-s", "99", str(tmp_path / "g2.tree"), "IIDRK11", RK_HOST, str(tmp_path / "g2.map"), "1.5"]
    n = 20000
    p = ctx.run("BranchRelaxer", args, n=n, rng=sg.RNG_PHILOX)
    ok = p.rep_status == 0
    assert ok.mean() > 0.99 and set(np.unique(p.rep_status)) <= {0, 4}
    wants = oracle_run_many([("BranchRelaxer", with_seed(args, 99 + i)) for i in range(3000)])
    ref = np.array([[float(x) for x in LEN_RE.findall(w["stdout"])] for w in wants if w["status"] == 0 and w["stdout"]])
    nv = ref.shape[1]
    rel = p.values(1).reshape(n, nv)[ok]
    for x in range(0, nv, max(1, nv // 5)):
        if ref[:, x].std() > 0:
            assert ks_ok(rel[:, x], ref[:, x]), x


# ---- sub-batch pipeline (find of sub-batch k+1 on a second stream next to build / label / emit of k) ------------------------------------
def test_pipelined_run_equals_serial_run(ctx, monkeypatch):
    args = ["-s", "77", "-min", "6", "-max", "14", "-p", "0.8", "1.0", "2.5", "0.3", "x"]
    n = 300000                                   # three sub-batches of 131072
    piped = ctx.run("HostTreeGen", args, n=n, rng=sg.RNG_PHILOX, offset=5, flags=sg.FLAG_PIPELINE)
    assert piped.timings()["sub_batches"] == 3
    serial = ctx.run("HostTreeGen", args, n=n, rng=sg.RNG_PHILOX, offset=5)
    assert serial.timings()["sub_batches"] == 1
    assert (piped.rep_status == 0).all()
    for s in range(4):
        assert piped.stream_bytes(s) == serial.stream_bytes(s) and piped.stream_bytes(s) > 0
        assert (piped.offsets(s) == serial.offsets(s)).all()
        assert piped.stream(s) == serial.stream(s), s
    assert (piped.summary().as_vector() == serial.summary().as_vector()).all()
    assert (piped.attempts == serial.attempts).all()
    piped.free()
    # an output estimate that turns out too small (forced here) falls back to the unpipelined path: same bytes again
    monkeypatch.setenv("SGP_OUTPUT_MARGIN", "0.6")
    again = ctx.run("HostTreeGen", args, n=n, rng=sg.RNG_PHILOX, offset=5, flags=sg.FLAG_PIPELINE)
    monkeypatch.delenv("SGP_OUTPUT_MARGIN")
    assert again.timings()["sub_batches"] == 1
    for s in range(4):
        assert again.stream(s) == serial.stream(s), s


# ---- general engine, one-pass mode (accepted trees copied into a bump-allocated arena by the find kernel) --------------------------------
def test_one_pass_general_engine_equals_two_pass(ctx, monkeypatch, tmp_path):
    """Batches of >= 8192 replicates skip the build kernel: the find kernel's warps copy every accepted tree into an arena sized from a
    pilot run. Same bytes as the two-pass mode (SGP_TWO_PASS), also when the arena estimate is too small (forced: the engine then
    falls back to the build kernel), in both RNG modes; and replicates of the one-pass run are the oracle's, byte for byte."""
    from oracle_util import oracle_run
    n = 9000
    cases = [("GuestTreeGen", ["-s", "61", "-db", "simple", HOST8, "0.5", "0.3", "0.8", "g"])]
    d = tmp_path / "genes"; d.mkdir()
    for i in range(1, 4):
        f = oracle_run("GuestTreeGen", ["-s", str(i), "-db", "none", HOST8, "0.4", "0.2", "0.2", "g"])["files"]
        (d / f"gene{i}.tree").write_bytes(f["g.pruned.tree"])
    cases.append(("DomainTreeGen", ["-s", "62", "-ig", "0.6", HOST8, str(d), "0.4", "0.3", "0.7", "o"]))
    for app, args in cases:
        for rng in (sg.RNG_MT_REPLAY, sg.RNG_PHILOX):
            one = ctx.run(app, args, n=n, rng=rng)
            assert one.timings()["build_ms"] < 0.05                     # no build kernel ran
            monkeypatch.setenv("SGP_TWO_PASS", "1")
            two = ctx.run(app, args, n=n, rng=rng)
            monkeypatch.delenv("SGP_TWO_PASS")
            assert two.timings()["build_ms"] > 0.05
            monkeypatch.setenv("SGP_ARENA_MARGIN", "0.4")
            small = ctx.run(app, args, n=n, rng=rng)
            monkeypatch.delenv("SGP_ARENA_MARGIN")
            assert small.timings()["build_ms"] > 0.05                   # the arena was too small: rebuilt
            assert (one.rep_status == two.rep_status).all() and (one.attempts == two.attempts).all()
            for s in range(8):
                assert one.stream_bytes(s) == two.stream_bytes(s) == small.stream_bytes(s)
                assert bytes(one.stream(s)) == bytes(two.stream(s)), (app, rng, s)
                assert bytes(small.stream(s)) == bytes(two.stream(s)), (app, rng, s)
            if rng == sg.RNG_MT_REPLAY:
                seed0 = int(args[1])
                for i in (0, 1, n // 2, n - 1):
                    a2 = list(args); a2[1] = str(seed0 + i)
                    want = oracle_run(app, a2)
                    if want["status"] != 0:
                        continue
                    prefix = args[-1]
                    for sfx, data in one.files(i).items():
                        assert data == want["files"][prefix + sfx], (app, i, sfx)
            one.free(); two.free(); small.free()
    # chained stages read the arena a one-pass run leaves (acceptance order, arena-capacity stride): same relaxed trees and the same
    # DomainTreeGen forest as from a two-pass source
    gargs = ["-s", "64", "-db", "none", HOST8, "0.4", "0.2", "0.3", str(tmp_path / "gt")]
    src1 = ctx.run("GuestTreeGen", gargs, n=n, rng=sg.RNG_PHILOX, flags=sg.FLAG_KEEP_TREES, keep_on_device=True)
    monkeypatch.setenv("SGP_TWO_PASS", "1")
    src2 = ctx.run("GuestTreeGen", gargs, n=n, rng=sg.RNG_PHILOX, flags=sg.FLAG_KEEP_TREES, keep_on_device=True)
    monkeypatch.delenv("SGP_TWO_PASS")
    assert src1.timings()["build_ms"] < 0.05 < src2.timings()["build_ms"]
    rargs = ["-s", "9", "x", "IIDLogNormal", "0", "0.25"]
    r1 = ctx.relax_on_batch(src1, rargs, which=sg.STREAM_PRUNED_TREE, rng=sg.RNG_PHILOX); r2 = ctx.relax_on_batch(src2, rargs, which=sg.STREAM_PRUNED_TREE, rng=sg.RNG_PHILOX)
    assert (r1.rep_status == r2.rep_status).all() and bytes(r1.stream(0)) == bytes(r2.stream(0))
    dargs = ["-s", "3", "-ig", "0.6", HOST8, "chained", "0.4", "0.3", "0.9", "o"]
    d1 = ctx.domain_on_batch(src1, dargs, first=10, count=6, n=4, rng=sg.RNG_MT_REPLAY); d2 = ctx.domain_on_batch(src2, dargs, first=10, count=6, n=4, rng=sg.RNG_MT_REPLAY)
    for s in range(8):
        assert bytes(d1.stream(s)) == bytes(d2.stream(s)), s
    for b in (r1, r2, d1, d2, src1, src2):
        b.free()
    # a small batch goes one-pass too once its configuration has run on the context (no pilot needed): same bytes as the first run
    args = ["-s", "63", "-db", "none", HOST8, "0.5", "0.3", "0.8", "g"]
    first = ctx.run("GuestTreeGen", args, n=300, rng=sg.RNG_MT_REPLAY)
    again = ctx.run("GuestTreeGen", args, n=300, rng=sg.RNG_MT_REPLAY)
    assert first.timings()["build_ms"] > 0.05 and again.timings()["build_ms"] < 0.05
    for s in range(8):
        assert bytes(first.stream(s)) == bytes(again.stream(s))
    first.free(); again.free()


def test_pipelined_caller_gets_the_bytes_of_separate_runs(ctx):
    # The end-to-end pattern of bench.py: chunk i's outputs are copied to pinned host memory on the copy stream while chunk i+1
    # computes and chunk i-1's buffers have just gone back to the engine's free list (large buffers are recycled between runs,
    # csrc/engine.cu BlockCache). Every chunk must arrive as the bytes of a separate, synchronous run of the same replicates.
    import hashlib
    import torch
    args = ["-s", "20260101", "-min", "100", "-max", "100", "-p", "0.8", "-a", "10000", "1.0", "5.0", "0.05", "bench"]
    n = 4000                                   # ~45 MB per Newick stream: above the recycling threshold
    want = []
    for i in range(4):
        b = ctx.run("HostTreeGen", args, n=n, rng=sg.RNG_PHILOX, offset=i * n, keep_on_device=True)
        want.append([hashlib.sha256(b.stream(s)).hexdigest() for s in range(4)])
        assert b.stream_bytes(0) > 32 << 20
        b.free()
    cap = 64 << 20
    pinned = [[torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(4)] for _ in range(2)]
    sizes = [[0] * 4 for _ in range(2)]
    got = [None] * 4
    prev = None
    for i in range(4):
        b = ctx.run("HostTreeGen", args, n=n, rng=sg.RNG_PHILOX, offset=i * n, keep_on_device=True)
        if prev is not None:
            ctx.wait_copies()
            got[i - 1] = [hashlib.sha256(pinned[(i - 1) & 1][s][:sizes[(i - 1) & 1][s]].numpy().tobytes()).hexdigest() for s in range(4)]
            prev.free()
        for s in range(4):
            sizes[i & 1][s] = b.copy_stream_into_async(s, pinned[i & 1][s].data_ptr(), cap)
        prev = b
    ctx.wait_copies()
    got[3] = [hashlib.sha256(pinned[1][s][:sizes[1][s]].numpy().tobytes()).hexdigest() for s in range(4)]
    prev.free()
    assert got == want
