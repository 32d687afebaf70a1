"""GPU parity tests: everything goes through the C ABI (libsagephy_b200.so) and is compared with the CPU oracle.

Bar: byte-exact text in MT seed-replay mode (replicate i == reference-style run with -s seed+i); two-sample KS p > 0.01
between Philox throughput mode and the MT oracle for the continuous statistics, chi-square for the discrete ones."""
import math
import os
import re

import numpy as np
import pytest
from scipy import stats

import sagephy_b200 as sg
from oracle_util import oracle_batch, oracle_double_to_string, oracle_mt_words, oracle_run, oracle

pytestmark = pytest.mark.gpu


def seed_index(args):
    return (args.index("-s") if "-s" in args else args.index("--seed")) + 1


def with_seed(args, seed):
    a = list(args)
    a[seed_index(a)] = str(seed)
    return a


def assert_replay_parity(ctx, app, args, n, prefix):
    seed0 = int(args[seed_index(args)])
    b = ctx.run(app, args, n=n, rng=sg.RNG_MT_REPLAY)
    status = b.rep_status
    for i in range(n):
        want = oracle_run(app, with_seed(args, seed0 + i))
        if want["status"] == 1:
            assert status[i] == 1, (i, status[i])
            assert b.replicate(3, i) == b"Failed to produce valid pruned tree within max allowed attempts.\n"
            continue
        assert status[i] == 0, (app, args, i, status[i])
        assert b.attempts[i] == want["attempts"]
        got = b.files(i)
        assert set(prefix + k for k in got) == set(want["files"]), (sorted(got), sorted(want["files"]))
        for sfx, data in got.items():
            exp = want["files"][prefix + sfx]
            if data != exp:
                k = next((j for j in range(min(len(data), len(exp))) if data[j] != exp[j]), min(len(data), len(exp)))
                raise AssertionError(f"{app} {args} replicate {i} stream {sfx}: first difference at byte {k}:\n"
                                     f" gpu   : {data[max(0, k - 60):k + 60]!r}\n oracle: {exp[max(0, k - 60):k + 60]!r}")
    return b


# ---- device-side primitives ------------------------------------------------------------------------------------------------
def test_device_double_to_string_matches_oracle(ctx):
    rng = np.random.default_rng(3)
    vals = np.concatenate([
        rng.random(4000), -np.log1p(-rng.random(4000)) / 5.05, np.round(rng.random(2000), 8), rng.random(2000) * 1e-4, rng.random(2000) * 1e8,
        rng.random(1000) * 10.0 ** rng.integers(-300, 300, 1000), 2.0 ** np.arange(-1074, 1024, 7.0),
        np.array([0.0, 1.0, 100.0, 1e7, 9999999.999999998, 1e-3, 9.999999999999999e-4, 1e21, 1e22, 1e23, 4.9e-324, 1.7976931348623157e308, 1e-64, 1234567.0])])
    got = ctx.double_to_string(vals)
    for v, s in zip(vals, got):
        assert s == oracle_double_to_string(float(v)), (v, s)


def test_device_math_is_bit_exact_with_oracle(ctx):
    L = oracle()
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.random(20000), rng.random(5000) * 10.0 ** rng.integers(-100, 100, 5000), 1.0 - rng.random(5000) * 1e-9])
    for fn, f in ((0, L.oracle_log), (1, L.oracle_log10)):
        got = ctx.math(fn, x)
        want = np.array([f(float(v)) for v in x])
        assert (got.view(np.uint64) == want.view(np.uint64)).all()
    y = (rng.random(20000) - 0.5) * 80
    assert (ctx.math(2, y).view(np.uint64) == np.array([L.oracle_exp(float(v)) for v in y]).view(np.uint64)).all()
    z = np.concatenate([x, -x[:1000], rng.random(5000) * 40])
    assert (ctx.math(3, z).view(np.uint64) == np.array([L.oracle_round_sig(float(v), 8) for v in z]).view(np.uint64)).all()


def test_device_pow_and_quantiles_are_bit_exact_with_oracle(ctx):
    L = oracle()
    rng = np.random.default_rng(5)
    x = rng.random(20000) * 10 + 1e-3
    assert (ctx.math(5, x).view(np.uint64) == np.array([L.oracle_pow(float(v), 0.37) for v in x]).view(np.uint64)).all()
    y = (rng.random(20000) - 0.7) * 30
    assert (ctx.math(8, y).view(np.uint64) == np.array([L.oracle_pow(2.718281828459045, float(v)) for v in y]).view(np.uint64)).all()
    p = np.concatenate([rng.random(20000), rng.random(3000) * 0.03, 1 - rng.random(3000) * 0.03])
    assert (ctx.math(4, p).view(np.uint64) == np.array([L.oracle_normal_quantile(float(v)) for v in p]).view(np.uint64)).all()
    q = 2e-6 + rng.random(20000) * (0.999998 - 2e-6)
    for fn, k in ((6, 4.0), (7, 0.25)):
        got = ctx.math(fn, q)
        want = np.array([L.oracle_gamma_quantile(float(v), k / 2, 2.0) for v in q])
        assert (got.view(np.uint64) == want.view(np.uint64)).all()


def test_throughput_mode_log_is_accurate(ctx):
    # log_unit (Philox path only): 0 < x <= 1, normal doubles. Table-driven, so the error is absolute near x = 1:
    # |error| <= max(2 ulp, 3e-16) against libm — far below anything an exponential waiting time can resolve.
    rng = np.random.default_rng(6)
    x = np.concatenate([1.0 - rng.random(50000) * (1 - 2.0 ** -52), rng.random(20000), 2.0 ** -rng.integers(0, 52, 2000).astype(float),
                        1.0 - 2.0 ** -rng.integers(1, 53, 500).astype(float), np.array([1.0, 2.0 ** -52, 0.5, 0.7071067811865476, 0.7071067811865475, 1.0 - 2.0 ** -53])])
    got, want = ctx.math(9, x), np.log(x)
    err = np.abs(got - want) / np.maximum(2.0 * np.spacing(np.abs(want)), 3e-16)
    assert err.max() <= 1.0, (err.max(), x[err.argmax()])
    assert (got[x < 1.0] < 0).all() and np.isfinite(got).all()


@pytest.mark.parametrize("seed", [0, 1, 20260101, 2 ** 62, -1, -300])
def test_device_mt19937_stream(ctx, seed):
    assert (ctx.mt_words(seed, 3000) == oracle_mt_words(seed, 3000)).all()


# ---- seed-replay parity: HostTreeGen -------------------------------------------------------------------------------------------
HOST_CASES = [
    ["-s", "1", "-min", "3", "-max", "4", "-p", "0.8", "1.0", "2.0", "0.5", "o"],
    ["-s", "10", "1.0", "3.0", "0.3", "o"],                                                        # defaults
    ["-s", "20", "-min", "10", "-max", "40", "-p", "0.6", "2.0", "1.5", "0.5", "o"],
    ["-s", "30", "-bi", "-min", "4", "-max", "60", "1.0", "3.0", "0.2", "o"],                      # forced bifurcation, minper = 1
    ["-s", "40", "-nox", "-min", "2", "-max", "30", "1.0", "2.5", "0.4", "o"],                     # no tags
    ["-s", "50", "-stem", "0.25", "-vp", "Sp", "-min", "2", "-max", "30", "1.0", "2.5", "0.4", "o"],
    ["-s", "60", "-min", "0", "-max", "50", "0.7", "1.0", "2.0", "o"],                             # empty pruned trees occur
    ["-s", "70", "-min", "30", "-max", "30", "-a", "40", "1.0", "3.0", "0.3", "o"],                # some replicates hit max attempts
    ["-s", "80", "-p", "0.35", "-min", "1", "-max", "20", "1.3", "2.0", "0.0", "o"],
]


@pytest.mark.parametrize("args", HOST_CASES, ids=[" ".join(a[2:-1]) for a in HOST_CASES])
def test_host_tree_gen_replay_is_byte_exact(ctx, args):
    assert_replay_parity(ctx, "HostTreeGen", args, 24, "o")


def test_host_tree_gen_readme_config_replay(ctx):
    # BASELINE config C1: HostTreeGen -min 100 -max 100 1.0 5.0 0.05 (about 289 attempts per accepted tree)
    b = assert_replay_parity(ctx, "HostTreeGen", ["-s", "7", "-min", "100", "-max", "100", "1.0", "5.0", "0.05", "o"], 12, "o")
    assert (b.sampled_leaves == 100).all()


def test_host_tree_gen_quiet_and_usage(ctx):
    args = ["-s", "5", "-q", "-min", "3", "-max", "20", "1.0", "2.0", "0.3"]
    b = ctx.run("HostTreeGen", args, n=3, rng=sg.RNG_MT_REPLAY)
    want = "".join(oracle_run("HostTreeGen", with_seed(args, 5 + i))["stdout"] for i in range(3))
    assert b.stdout == want
    u = ctx.run("HostTreeGen", ["1.0", "2.0"], n=1)
    assert u.status_code == sg.STATUS_USAGE and u.stdout.startswith("Usage:")
    with pytest.raises(sg.SagephyError):
        ctx.run("HostTreeGen", ["-bogus", "1.0", "2.0", "0.3", "o"], n=1)


# ---- seed-replay parity: GuestTreeGen --------------------------------------------------------------------------------------------
HOST5 = "((A:0.4,B:0.4):0.6,(C:0.7,(D:0.2,E:0.2):0.5):0.3):0.2;"
HOST8 = "(((a:0.1,b:0.1):0.5,(c:0.3,d:0.3):0.3):0.4,((e:0.2,f:0.2):0.6,(g:0.5,h:0.5):0.3):0.2):0.1;"
GUEST_CASES = [
    ["-s", "1", "-db", "none", HOST5, "0.5", "0.4", "0.0", "o"],                        # no transfers
    ["-s", "2", "-db", "none", HOST5, "0.5", "0.4", "0.6", "o"],                        # uniform recipient
    ["-s", "3", HOST5, "0.5", "0.4", "0.6", "o"],                                       # code default: -db simple
    ["-s", "4", "-db", "exponential", "-dbr", "1.5", HOST8, "0.4", "0.3", "0.8", "o"],
    ["-s", "5", "-rt", "1.0", "-db", "none", HOST8, "0.4", "0.3", "1.2", "o"],          # replacing only
    ["-s", "6", "-rt", "0.0", HOST8, "0.4", "0.3", "1.0", "o"],                         # additive only
    ["-s", "7", "-gb", "-gbc", "2.1", HOST8, "0.4", "0.3", "0.5", "o"],                 # gene birth sampling
    ["-s", "8", "-nox", "-p", "0.7", HOST8, "0.6", "0.2", "0.4", "o"],
    ["-s", "9", "-stem", "0.3", "-min", "3", "-max", "40", "-minper", "0", "-maxper", "6", HOST8, "0.8", "0.5", "0.5", "o"],
    ["-s", "10", "-vp", "Gene", "-a", "3", "-min", "12", "-max", "14", HOST5, "0.5", "0.4", "0.3", "o"],   # many failures
    ["--seed", "11", "-vph", "-mpr", "--distance-bias", "none", "--replacing-transfers", "0.3", "--leaf-sampling-probability", "0.9", HOST8, "0.5", "0.4", "0.6", "o"],   # long names, host-suffixed vertex names, -mpr (accepted and ignored)
]


@pytest.mark.parametrize("args", GUEST_CASES, ids=[" ".join(a[2:a.index(HOST5) if HOST5 in a else a.index(HOST8)]) or "plain" for a in GUEST_CASES])
def test_guest_tree_gen_replay_is_byte_exact(ctx, args):
    assert_replay_parity(ctx, "GuestTreeGen", args, 24, "o")


def test_guest_tree_gen_victim_search_through_arc_lists(ctx, monkeypatch):
    """Large hosts find the lineage a replacing transfer replaces through per-arc vertex lists instead of a scan of the pool
    (switched on by host size); forced on here, the replay output must still be the oracle's, byte for byte."""
    monkeypatch.setenv("SGP_ARC_LISTS", "1")
    for args in (["-s", "21", "-db", "none", "-rt", "1.0", HOST8, "0.5", "0.3", "0.9", "g"],
                 ["-s", "22", "-db", "simple", "-rt", "0.7", HOST8, "0.6", "0.2", "1.2", "g"]):
        assert_replay_parity(ctx, "GuestTreeGen", args, 24, "g").free()


def test_guest_tree_gen_large_trees_rerun_with_bigger_pools(ctx):
    # trees beyond the first scratch-pool size (1024 vertices) are re-run with 8x the pool: same bytes as the oracle
    args = ["-s", "41", "-db", "none", "-max", "100000", "-maxper", "100000", HOST8, "5.0", "0.2", "0.3", "o"]
    b = assert_replay_parity(ctx, "GuestTreeGen", args, 6, "o")
    assert b.sampled_leaves.max() > 600


def test_guest_tree_gen_quiet_prints_the_pruned_tree(ctx):
    args = ["-s", "15", "-q", "-db", "none", HOST5, "0.5", "0.4", "0.5"]            # GuestTreeGen.java:59-64
    b = ctx.run("GuestTreeGen", args, n=4, rng=sg.RNG_MT_REPLAY)
    assert b.stdout == "".join(oracle_run("GuestTreeGen", with_seed(args, 15 + i))["stdout"] for i in range(4))


def test_guest_tree_gen_on_generated_100_leaf_host(ctx, tmp_path):
    # BASELINE config C3 shape: host = pruned tree of a C1 run, dup/loss/trans 0.3, additive + replacing transfers
    host = oracle_run("HostTreeGen", ["-s", "7", "-min", "100", "-max", "100", "1.0", "5.0", "0.05", "sp"])["files"]["sp.pruned.tree"]
    p = tmp_path / "species.pruned.tree"
    p.write_bytes(host)
    for db in ("none", "simple", "exponential"):
        assert_replay_parity(ctx, "GuestTreeGen", ["-s", "100", "-db", db, str(p), "0.3", "0.3", "0.3", "o"], 6, "o")
    # the host tree given as a Newick string: the "Arguments:" line of both .info files then embeds ~12 KB of text
    assert_replay_parity(ctx, "GuestTreeGen", ["-s", "200", "-db", "none", host.decode().strip(), "0.3", "0.3", "0.3", "o"], 5, "o")


def test_label_kernels_agree(ctx, monkeypatch):
    # The warp-per-tree (shared-memory, level-synchronous) and thread-per-tree (serial, for oversized trees) label kernels must
    # produce the same bytes; SGP_LABEL_THREAD_KERNEL routes every tree through the latter.
    monkeypatch.setenv("SGP_LABEL_THREAD_KERNEL", "1")
    assert_replay_parity(ctx, "HostTreeGen", HOST_CASES[2], 24, "o")
    assert_replay_parity(ctx, "HostTreeGen", HOST_CASES[5], 24, "o")
    assert_replay_parity(ctx, "GuestTreeGen", GUEST_CASES[4], 24, "o")
    args = ["-s", "11", "-min", "20", "-max", "60", "-p", "0.7", "1.0", "4.0", "0.5", "o"]
    slow = ctx.run("HostTreeGen", args, n=2000, rng=sg.RNG_PHILOX)
    monkeypatch.delenv("SGP_LABEL_THREAD_KERNEL")
    fast = ctx.run("HostTreeGen", args, n=2000, rng=sg.RNG_PHILOX)
    for st in range(4):
        assert bytes(slow.stream(st)) == bytes(fast.stream(st))
    # trees beyond the shared-memory stage (> 2 340 vertices) are labelled in place by the same warp algorithm with global work arrays:
    # SGP_LABEL_GLOBAL_KERNEL routes every tree through that kernel
    monkeypatch.setenv("SGP_LABEL_GLOBAL_KERNEL", "1")
    assert_replay_parity(ctx, "HostTreeGen", HOST_CASES[2], 24, "o")
    assert_replay_parity(ctx, "GuestTreeGen", GUEST_CASES[4], 24, "o")
    assert_replay_parity(ctx, "GuestTreeGen", GUEST_CASES[7], 24, "o")
    inplace = ctx.run("HostTreeGen", args, n=2000, rng=sg.RNG_PHILOX)
    monkeypatch.delenv("SGP_LABEL_GLOBAL_KERNEL")
    for st in range(4):
        assert bytes(inplace.stream(st)) == bytes(fast.stream(st))


def test_output_layouts_match_separate_reference_runs(ctx, tmp_path):
    # per-replicate layout == what N reference runs with prefixes <prefix>_<i> leave behind; the packed
    # layout + tools/unpack_batch.py gives the same files again. "-a 40" makes some replicates fail (failure note only).
    import importlib.util
    spec = importlib.util.spec_from_file_location("unpack_batch", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "unpack_batch.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    for app, core in (("HostTreeGen", ["-min", "30", "-max", "30", "-a", "40", "1.0", "3.0", "0.3"]), ("GuestTreeGen", ["-a", "3", "-min", "12", "-max", "14", HOST5, "0.5", "0.4", "0.3"])):
        n, first = 10, 5
        want = {}
        for i in range(n):
            w = oracle_run(app, ["-s", str(70 + i)] + core + [f"o_{first + i}"])
            want.update({k: v for k, v in w["files"].items()})
        for layout in (sg.LAYOUT_PER_REPLICATE, sg.LAYOUT_PACKED):
            d = tmp_path / f"{app}_{layout}"; d.mkdir()
            b = ctx.run(app, ["-s", "70"] + core + [str(d / "o")], n=n, rng=sg.RNG_MT_REPLAY)
            b.write_files(layout=layout, first_index=first)
            if layout == sg.LAYOUT_PACKED:
                assert mod.unpack(str(d / "o"), out_dir=str(d / "split")) > 0
                got_dir = d / "split"
            else:
                got_dir = d
            got = {p.name: p.read_bytes() for p in got_dir.iterdir() if p.is_file() and not p.name.endswith(".idx") and p.name != "o" and "_" in p.name}
            assert set(got) == set(want), (sorted(set(got) ^ set(want)))
            for k in want:
                # the "Arguments:" line echoes the output prefix, which differs by construction (tmp dir vs bare name)
                strip = lambda t: b"\n".join(l for l in t.split(b"\n") if not l.startswith(b"Arguments:"))
                assert strip(got[k]) == strip(want[k]), (app, layout, k)


def test_leaf_sizes_file_option(ctx, tmp_path):
    # -sizes: the wanted number of sampled leaves is drawn uniformly from a one-column file before the retry loop starts
    # (GuestTreeMachina.java:76) and every attempt is tested against it exactly (:159-163)
    f = tmp_path / "sizes.txt"
    f.write_text("3\n5\n8\n5\n12\n")
    assert_replay_parity(ctx, "HostTreeGen", ["-s", "90", "-sizes", str(f), "-min", "2", "-max", "20", "-p", "0.9", "1.0", "2.5", "0.3", "o"], 24, "o")
    assert_replay_parity(ctx, "HostTreeGen", ["-s", "91", "-bi", "-sizes", str(f), "-min", "2", "-max", "20", "1.0", "2.5", "0.3", "o"], 16, "o")
    assert_replay_parity(ctx, "GuestTreeGen", ["-s", "92", "-sizes", str(f), "-min", "2", "-max", "30", "-minper", "0", HOST5, "0.5", "0.4", "0.3", "o"], 16, "o")
    # Philox fast path: every accepted tree has one of the listed sizes, in the listed proportions (5 appears twice)
    n = 20000
    b = ctx.run("HostTreeGen", ["-s", "93", "-sizes", str(f), "-min", "2", "-max", "20", "-p", "0.9", "1.0", "2.5", "0.3", "o"], n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    assert (b.rep_status == 0).all()
    k, c = np.unique(b.sampled_leaves, return_counts=True)
    assert list(k) == [3, 5, 8, 12]
    assert stats.chisquare(c, f_exp=n * np.array([0.2, 0.4, 0.2, 0.2])).pvalue > 0.001


def test_native_cli_writes_the_reference_files(tmp_path):
    # `sagephy <App> ...` is the stand-in for `java -jar sagephy.jar <App> ...` (apps/SaGePhyStarter.java:75-103)
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sagephy_b200", "sagephy")
    core = ["-s", "3", "-min", "5", "-max", "12", "1.0", "3.0", "0.3", "species"]
    subprocess.check_call([exe, "HostTreeGen"] + core, cwd=tmp_path)
    want = oracle_run("HostTreeGen", core)["files"]
    assert sorted(p.name for p in tmp_path.iterdir()) == sorted(want)
    for name, data in want.items():
        assert (tmp_path / name).read_bytes() == data, name
    out = subprocess.run([exe, "HostTreeGen", "-q"] + core[:-1], cwd=tmp_path, capture_output=True, text=True, check=True).stdout
    assert out == oracle_run("HostTreeGen", ["-q"] + core[:-1])["stdout"]
    usage = subprocess.run([exe, "GuestTreeGen", "x"], cwd=tmp_path, capture_output=True, text=True)
    assert usage.returncode == 0 and usage.stdout.startswith("Usage:")
    d = tmp_path / "many"; d.mkdir()
    subprocess.check_call([exe, "GuestTreeGen", "-n", "3", "-rng", "mt-replay", "--out-layout", "per-replicate", "-s", "4", "-db", "none", HOST5, "0.3", "0.2", "0.4", "gene"], cwd=d)
    for i in range(3):
        w = oracle_run("GuestTreeGen", ["-s", str(4 + i), "-db", "none", HOST5, "0.3", "0.2", "0.4", f"gene_{i}"])["files"]
        for name, data in w.items():
            got = (d / name).read_bytes()
            strip = lambda t: b"\n".join(l for l in t.split(b"\n") if not l.startswith(b"Arguments:"))
            assert strip(got) == strip(data), name


def test_hybrid_host_graph_ends_like_the_reference(ctx, tmp_path):
    # the reference validates the GML network and then dies on its first guest vertex (GuestVertex.java:100 gets the
    # null GuestTreeInHybridGraphCreator.java:272 passes) - stack trace, exit 0, no files; so does the drop-in, for any -n
    import subprocess
    from test_hybrid_cpu import ALLO, BAD_GRAPHS, TAIL
    (tmp_path / "net.gml").write_text(ALLO)
    args = ["-s", "4", "-hybrid", "0.2", "2.0", "0.5", "net.gml", "0.3", "0.2", "0.0", "out"]
    want = oracle_run("GuestTreeGen", ["-s", "4", "-hybrid", "0.2", "2.0", "0.5", str(tmp_path / "net.gml"), "0.3", "0.2", "0.0", "out"])
    assert want["files"] == {} and want["stderr"] == "java.lang.NullPointerException" + TAIL
    for n in (1, 1000):
        b = ctx.run("GuestTreeGen", ["-s", "4", "-hybrid", "0.2", "2.0", "0.5", str(tmp_path / "net.gml"), "0.3", "0.2", "0.0", "out"], n=n, check=False)
        assert b.status_code not in (sg.STATUS_OK, sg.STATUS_USAGE) and b.stderr == want["stderr"] and b.stdout == ""
    bad = ["-s", "4", "-hybrid", "0.2", "2.0", "0.5", BAD_GRAPHS["not_connected"], "0.3", "0.2", "0.0", "out"]
    assert ctx.run("GuestTreeGen", bad, n=5, check=False).stderr == oracle_run("GuestTreeGen", bad)["stderr"]
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sagephy_b200", "sagephy")
    for extra in ([], ["-n", "100"], ["-n", "100", "--gpus", "2"]):
        r = subprocess.run([exe, "GuestTreeGen"] + extra + args, cwd=tmp_path, capture_output=True, text=True, env=dict(os.environ, SAGEPHY_DEVICES="0,0"))
        assert r.returncode == 0 and r.stdout == "" and r.stderr == want["stderr"], (extra, r)
        assert sorted(p.name for p in tmp_path.iterdir()) == ["net.gml"]


# ---- Philox throughput mode: distributions ---------------------------------------------------------------------------------------
def ks_ok(a, b):
    return stats.ks_2samp(a, b).pvalue > 0.01


def chi2_two_sample_ok(a, b, max_bins=40):
    lo, hi = int(min(a.min(), b.min())), int(max(a.max(), b.max()))
    edges = np.unique(np.quantile(np.concatenate([a, b]), np.linspace(0, 1, max_bins + 1)).astype(int))
    edges = np.concatenate([[lo - 1], edges[(edges > lo) & (edges < hi)], [hi]])
    ha, _ = np.histogram(a, bins=edges + 0.5)
    hb, _ = np.histogram(b, bins=edges + 0.5)
    keep = (ha + hb) > 10
    if keep.sum() < 2:
        return True
    return stats.chi2_contingency(np.vstack([ha[keep], hb[keep]]))[1] > 0.01


def test_host_tree_gen_philox_matches_oracle_distributions(ctx):
    n = 100000
    args = ["-s", "20260101", "-min", "2", "-max", "1000", "-p", "0.8", "1.0", "3.0", "0.5", "x"]
    b = ctx.run("HostTreeGen", args, n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    assert (b.rep_status == 0).all()
    sec, st = oracle_batch("HostTreeGen", args, 0, n)
    assert chi2_two_sample_ok(b.sampled_leaves.astype(float), st[:, 4])
    assert ks_ok(b.root_times, st[:, 5])                       # height of the pruned tree's root
    assert ks_ok(b.total_times, st[:, 6])
    ev = b.event_counts
    for e in (4, 5, 3):                                         # duplications (births), losses, unsampled leaves
        assert chi2_two_sample_ok(ev[:, e].astype(float), st[:, 9 + e]), e
    assert chi2_two_sample_ok(b.attempts.astype(float), st[:, 0])
    # closed form: P(2 <= K <= 1000) for the rho-thinned process is not needed here; check E[attempts] agreement instead
    assert abs(b.attempts.mean() - st[:, 0].mean()) < 0.02


def test_host_tree_gen_philox_acceptance_rate_closed_form(ctx):
    # P(N = 100 | lam 5, mu 0.05, T 1) = 3.4593e-3 -> 289.08 trees per accept; Attempts = trees + 1
    n = 4000
    b = ctx.run("HostTreeGen", ["-s", "1", "-min", "100", "-max", "100", "1.0", "5.0", "0.05", "x"], n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    assert (b.rep_status == 0).all() and (b.sampled_leaves == 100).all()
    trees = b.attempts - 1
    assert abs(trees.mean() - 289.08) < 4.5 * 289.08 / math.sqrt(n)
    # results do not depend on how replicates are sharded: same global ids -> same bytes
    c = ctx.run("HostTreeGen", ["-s", "1", "-min", "100", "-max", "100", "1.0", "5.0", "0.05", "x"], n=100, offset=1000, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    off = b.offsets(1)
    assert c.stream(1) == b.stream(1)[int(off[1000]):int(off[1100])]


def test_guest_tree_gen_philox_matches_oracle_distributions(ctx):
    n = 30000
    for db in ("none", "simple"):
        args = ["-s", "777", "-db", db, HOST8, "0.5", "0.4", "0.7", "x"]
        b = ctx.run("GuestTreeGen", args, n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
        ok = b.rep_status == 0
        assert ok.all()
        sec, st = oracle_batch("GuestTreeGen", args, 0, n)
        assert chi2_two_sample_ok(b.sampled_leaves.astype(float), st[:, 4])
        assert ks_ok(b.total_times, st[:, 6])
        ev = b.event_counts
        for e in (0, 4, 5, 6, 7, 8):
            assert chi2_two_sample_ok(ev[:, e].astype(float), st[:, 9 + e]), (db, e)
        assert chi2_two_sample_ok(b.attempts.astype(float), st[:, 0])


# ---- size-independent properties at larger scale ---------------------------------------------------------------------------------
def test_emitted_trees_are_well_formed_at_scale(ctx):
    n = 20000
    b = ctx.run("HostTreeGen", ["-s", "9", "-min", "100", "-max", "100", "-p", "0.8", "1.0", "5.0", "0.05", "x"], n=n, rng=sg.RNG_PHILOX)
    assert (b.rep_status == 0).all()
    pruned, off = b.stream(1), b.offsets(1)
    assert off[-1] == len(pruned) and (np.diff(off.astype(np.int64)) > 0).all()
    assert pruned.count(b"\n") == n and pruned.count(b"[NAME=PrunedTree];\n") == n
    assert pruned.count(b"VERTEXTYPE=Leaf]") == 100 * n          # exactly 100 sampled leaves in every pruned tree
    for i in (0, 1, n // 2, n - 1):
        t = pruned[int(off[i]):int(off[i + 1])].decode()
        bare = re.sub(r"\[[^\]]*\]", "", t)
        assert bare.count("(") == bare.count(")") == 99 and bare.count(",") == 99
        ids = sorted(int(x) for x in re.findall(r"\[ID=(\d+) ", t))
        assert ids == list(range(199))
        info = b.replicate(3, i).decode()
        assert "No. of vertices:\t199\n" in info and "No. of extant leaves:\t100\n" in info
    s = b.summary()
    assert s.n_ok == n and s.leaf_hist[100] == n and s.out_bytes[1] == len(pruned)


# ---- BranchRelaxer ---------------------------------------------------------------------------------------------------------------
RELAX_TREE = "((A:0.5,B:0.5)X:0.3,(C:0.2,D:0.2,E:0.2)Y:0.6)R:0.1;"          # multifurcating, named interior vertices, stem
RELAX_TREE_NOSTEM = "(((a:0.1,b:0.1):0.5,(c:0.3,d:0.3):0.3):0.4,((e:0.2,f:0.2):0.6,(g:0.5,h:0.5):0.3):0.2);"
RELAX_CASES = [
    ["-s", "1", "-x", "-innms", "-o", "rel.tree", RELAX_TREE, "IIDLogNormal", "0", "0.25"],
    ["-s", "2", "-o", "rel.tree", RELAX_TREE_NOSTEM, "IIDGamma", "2", "0.5"],
    ["-s", "3", "-x", "-o", "rel.tree", RELAX_TREE, "IIDGamma", "0.1", "3.0"],
    ["-s", "4", "-o", "rel.tree", RELAX_TREE_NOSTEM, "IIDNormal", "1.0", "0.09"],          # negative rates force retries
    ["-s", "5", "-x", "-o", "rel.tree", RELAX_TREE, "IIDUniform", "0.5", "1.5"],
    ["-s", "6", "-o", "rel.tree", RELAX_TREE, "IIDExponential", "2.0"],
    ["-s", "7", "-x", "-o", "rel.tree", RELAX_TREE_NOSTEM, "Constant", "1.7"],
    ["-s", "8", "-min", "0.8", "-max", "1.3", "-a", "50", "-o", "rel.tree", RELAX_TREE, "IIDLogNormal", "0", "0.04"],   # tight window: many attempts, some failures
    ["-s", "9", RELAX_TREE_NOSTEM, "iidlognormal", "0.1", "0.5"],                              # stdout, case-insensitive model name
    # autocorrelated models: rates follow the parent's rate along RTree.getTopologicalOrdering()
    ["-s", "10", "-x", "-o", "rel.tree", RELAX_TREE, "ACTK98", "1.0", "0.5"],
    ["-s", "11", "-o", "rel.tree", RELAX_TREE_NOSTEM, "ACTK98", "0.7", "2.0"],
    ["-s", "12", "-x", "-o", "rel.tree", RELAX_TREE, "ACRY07", "1.0", "0.5"],                  # stem: two draws at the root
    ["-s", "13", "-o", "rel.tree", RELAX_TREE_NOSTEM, "acry07", "1.3", "1.5"],                 # no stem: root keeps the start rate
    ["-s", "14", "-x", "-o", "rel.tree", RELAX_TREE, "ACABY02", "1.0"],
    ["-s", "15", "-x", "-innms", "-o", "rel.tree", RELAX_TREE, "ACLBPL07", "1.0", "1.0", "2.0", "0.5"],   # CIR, 200 Gaussian steps per branch
    ["-s", "16", "-o", "rel.tree", RELAX_TREE_NOSTEM, "ACLBPL07", "0.5", "1.2", "4.0", "1.5"],
    ["-s", "17", "-min", "0.6", "-max", "1.6", "-a", "30", "-o", "rel.tree", RELAX_TREE, "ACTK98", "1.0", "0.3"],   # retries, some failures
    ["-s", "18", "-o", "rel.tree", "((a:0.0,b:0.1):0.2,c:0.3):0.1;", "ACRY07", "1.0", "0.5"],   # zero-length branch: the reference throws
]


@pytest.mark.parametrize("args", RELAX_CASES, ids=[" ".join(a[2:]).replace(RELAX_TREE, "T1").replace(RELAX_TREE_NOSTEM, "T2") for a in RELAX_CASES])
def test_branch_relaxer_replay_is_byte_exact(ctx, args):
    n = 16
    seed0 = int(args[1])
    b = ctx.run("BranchRelaxer", args, n=n, rng=sg.RNG_MT_REPLAY)
    has_out = "-o" in args
    stdout = ""
    for i in range(n):
        want = oracle_run("BranchRelaxer", with_seed(args, seed0 + i))
        if want["status"] == 1:
            assert b.rep_status[i] == 1
            continue
        if want["status"] == 2:          # exception in the reference (stack trace, no output)
            assert b.rep_status[i] == 4, (i, b.rep_status[i], want["stderr"][:120])
            continue
        assert b.rep_status[i] == 0 and b.attempts[i] == want["attempts"], (i, b.rep_status[i], b.attempts[i], want["attempts"])
        if has_out:
            got = b.files(i)
            assert got[""] == want["files"]["rel.tree"], (got[""], want["files"]["rel.tree"])
            assert got[".info"] == want["files"]["rel.tree.info"]
        else:
            stdout += want["stdout"]
    if not has_out:
        assert b.stdout == stdout


def test_branch_relaxer_samples_from_file_and_multi_tree(ctx, tmp_path):
    f = tmp_path / "rates.txt"
    f.write_text("0.5\n1.0\n1.5\n2.0\n0.25\n")
    args = ["-s", "12", "-x", "-o", "rel.tree", RELAX_TREE, "IIDSamplesFromFile", str(f)]
    b = ctx.run("BranchRelaxer", args, n=8, rng=sg.RNG_MT_REPLAY)
    for i in range(8):
        want = oracle_run("BranchRelaxer", with_seed(args, 12 + i))["files"]
        assert b.files(i)[""] == want["rel.tree"] and b.files(i)[".info"] == want["rel.tree.info"]
    # multi-tree input: replicate i relaxes tree i % T with seed + i
    trees = [RELAX_TREE, RELAX_TREE_NOSTEM, "(x:1.0,y:2.0,z:3.0):0.5;"]
    m = tmp_path / "trees.nwk"
    m.write_text("\n".join(trees) + "\n")
    b = ctx.run("BranchRelaxer", ["-s", "40", "-x", str(m), "IIDLogNormal", "0", "0.25"], n=9, rng=sg.RNG_MT_REPLAY, flags=sg.FLAG_MULTI_TREE_INPUT)
    want = "".join(oracle_run("BranchRelaxer", ["-s", str(40 + i), "-x", trees[i % 3], "IIDLogNormal", "0", "0.25"])["stdout"] for i in range(9))
    assert b.stdout == want


def test_branch_relaxer_philox_matches_reference_distributions(ctx):
    n = 20000
    nv = 15      # vertices of RELAX_TREE_NOSTEM
    for model, extra, fn in (("IIDLogNormal", ["0", "0.25"], lambda u: None), ("IIDGamma", ["2", "0.5"], None), ("IIDNormal", ["1.0", "0.01"], None),
                             ("IIDUniform", ["0.5", "1.5"], None), ("IIDExponential", ["2.0"], None)):
        args = ["-s", "99", RELAX_TREE_NOSTEM, model] + extra
        b = ctx.run("BranchRelaxer", args, n=n, rng=sg.RNG_PHILOX)
        st = b.rep_status
        if model == "IIDGamma":
            # the reference aborts a run whose uniform leaves [2e-6, 0.999998] (ChiSquared.java:24): ~4e-6 per branch
            assert set(np.unique(st)) <= {0, 4} and (st == 4).sum() < 12
        else:
            assert (st == 0).all()
        rates = b.values(0).reshape(n, nv)[st == 0]
        r = ctx.run("BranchRelaxer", args, n=4000, rng=sg.RNG_MT_REPLAY)          # reference stream (MT) through the same quantile code
        ref = r.values(0).reshape(4000, nv)
        assert ks_ok(rates[:, 3], ref[:, 3]) and ks_ok(rates[:, 10], ref[:, 7])
        rel, orig = b.values(1).reshape(n, nv)[st == 0], np.array([0.1, 0.1, 0.5, 0.3, 0.3, 0.3, 0.4, 0.2, 0.2, 0.6, 0.5, 0.5, 0.3, 0.2, 0.0])
        assert np.array_equal(rel, rates * orig)
    # closed form for the gamma model: mean k*theta, variance k*theta^2
    b = ctx.run("BranchRelaxer", ["-s", "5", RELAX_TREE_NOSTEM, "IIDGamma", "2", "0.5"], n=n, rng=sg.RNG_PHILOX)
    g = b.values(0).reshape(n, nv)[b.rep_status == 0].ravel()
    assert abs(g.mean() - 1.0) < 0.01 and abs(g.var() - 0.5) < 0.02
    assert stats.kstest(g[::7], "gamma", args=(2, 0, 0.5)).pvalue > 0.001


def test_branch_relaxer_autocorrelated_models(ctx):
    # strict tree validation of the models with lengthsMustBeUltrametric() (io/PrIMENewickTreeVerifier.java:37-46)
    for tree in ("(a:0.1,a:0.1):0.2;", "((a:0.1,b:0.1),c:0.2):0.1;"):
        want = oracle_run("BranchRelaxer", ["-s", "1", tree, "ACTK98", "1.0", "0.5"])
        assert want["status"] == 2 and "NewickIOException" not in want["stdout"]
        with pytest.raises(sg.SagephyError):
            ctx.run("BranchRelaxer", ["-s", "1", tree, "ACTK98", "1.0", "0.5"], n=2, rng=sg.RNG_MT_REPLAY)
        ok = ctx.run("BranchRelaxer", ["-s", "1", tree, "ACABY02", "1.0"], n=2, rng=sg.RNG_MT_REPLAY)      # not strict
        assert ok.stdout == "".join(oracle_run("BranchRelaxer", ["-s", str(1 + i), tree, "ACABY02", "1.0"])["stdout"] for i in range(2))
    with pytest.raises(sg.SagephyError):
        ctx.run("BranchRelaxer", ["-s", "1", RELAX_TREE, "ACLBPL07", "1.0", "0.1", "0.1", "1.0"], n=1)      # 2*theta*mu < sigma^2
    # Philox mode draws from the same conditional distributions as the MT stream (same code, other uniforms)
    n, nv = 20000, 15
    for model in (["ACTK98", "1.0", "0.5"], ["ACRY07", "1.0", "0.5"], ["ACABY02", "1.0"], ["ACLBPL07", "1.0", "1.0", "2.0", "0.5"]):
        args = ["-s", "77", RELAX_TREE_NOSTEM] + model
        b = ctx.run("BranchRelaxer", args, n=n, rng=sg.RNG_PHILOX)
        r = ctx.run("BranchRelaxer", args, n=4000, rng=sg.RNG_MT_REPLAY)
        assert (b.rep_status == 0).all() and (r.rep_status == 0).all()
        got, ref = b.values(0).reshape(n, nv), r.values(0).reshape(4000, nv)
        for x in (0, 6, 13):
            assert ks_ok(got[:, x], ref[:, x]), (model, x)
        # parent-child dependence is there: vertex 0 (leaf a) follows vertex 2 (its parent)
        if model[0] != "ACLBPL07":
            assert abs(np.corrcoef(np.log(got[:, 0]), np.log(got[:, 2]))[0, 1] - np.corrcoef(np.log(ref[:, 0]), np.log(ref[:, 2]))[0, 1]) < 0.05
        assert np.array_equal(b.values(1).reshape(n, nv), got * np.array([0.1, 0.1, 0.5, 0.3, 0.3, 0.3, 0.4, 0.2, 0.2, 0.6, 0.5, 0.5, 0.3, 0.2, 0.0]))


# ---- chained stage: generator batch -> BranchRelaxer without a host round trip --------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("which", [sg.STREAM_PRUNED_TREE, sg.STREAM_UNPRUNED_TREE], ids=["pruned", "unpruned"])
def test_chained_relaxer_equals_reference_on_the_emitted_newick(ctx, which):
    # The chained stage must give what the reference gives when BranchRelaxer is launched on each emitted tree file.
    # "-min 0" on the generator makes some pruned trees empty (";"): those replicates must fail, not crash.
    gen_args = ["-s", "21", "-min", "0", "-max", "40", "-minper", "0", HOST8, "0.5", "0.4", "0.5", "g"]
    n = 24
    src = ctx.run("GuestTreeGen", gen_args, n=n, rng=sg.RNG_MT_REPLAY, flags=sg.FLAG_KEEP_TREES)
    for model in (["IIDLogNormal", "0", "0.25"], ["IIDGamma", "2", "0.5"], ["IIDUniform", "0.5", "1.5"], ["ACTK98", "1.0", "0.5"], ["ACLBPL07", "1.0", "1.0", "2.0", "0.5"]):
        for extra in ([], ["-x", "-innms"]):
            args = ["-s", "300"] + extra + ["chained-input"] + model
            out = ctx.relax_on_batch(src, args, which=which, rng=sg.RNG_MT_REPLAY)
            assert out.n == n
            n_checked = 0
            for i in range(n):
                text = src.replicate(which, i).decode()
                if src.rep_status[i] != 0 or text.strip() in ("", ";"):
                    assert out.rep_status[i] != 0
                    continue
                want = oracle_run("BranchRelaxer", ["-s", str(300 + i)] + extra + [text.strip()] + model)
                if want["status"] != 0 or not want["stdout"]:
                    assert out.rep_status[i] != 0, (i, want["stderr"][:200])
                    continue
                assert out.rep_status[i] == 0, (i, out.rep_status[i])
                assert out.replicate(0, i).decode() == want["stdout"], (which, model, extra, i)
                n_checked += 1
            assert n_checked >= n // 2
            out.free()
    src.free()


@pytest.mark.gpu
def test_chained_relaxer_equals_text_path_in_throughput_mode(ctx):
    # Philox mode: chaining on the device and re-parsing the emitted text on the host must give the same bytes, also with
    # several draws per tree (replicate i relaxes tree i mod n_src).
    n = 500
    src = ctx.run("HostTreeGen", ["-s", "5", "-min", "4", "-max", "30", "-p", "0.8", "1.0", "3.0", "0.3", "h"], n=n, rng=sg.RNG_PHILOX, flags=sg.FLAG_KEEP_TREES)
    assert (src.rep_status == 0).all()
    trees = "".join(src.replicate(sg.STREAM_PRUNED_TREE, i).decode() for i in range(n))
    for args in (["-s", "9", "-o", "r", "x", "IIDLogNormal", "0", "0.25"], ["-s", "9", "-x", "x", "IIDGamma", "2", "0.5"]):
        chained = ctx.relax_on_batch(src, args, which=sg.STREAM_PRUNED_TREE, n=3 * n, rng=sg.RNG_PHILOX)
        targs = list(args); targs[targs.index("x")] = trees
        text = ctx.run("BranchRelaxer", targs, n=3 * n, rng=sg.RNG_PHILOX, flags=sg.FLAG_MULTI_TREE_INPUT)
        assert (chained.rep_status == text.rep_status).all()
        assert bytes(chained.stream(0)) == bytes(text.stream(0))
        np.testing.assert_array_equal(chained.values(1), text.values(1))
        chained.free(); text.free()
    # a batch made without FLAG_KEEP_TREES cannot be chained
    plain = ctx.run("HostTreeGen", ["-s", "5", "1.0", "3.0", "0.3", "h"], n=4, rng=sg.RNG_PHILOX)
    with pytest.raises(sg.SagephyError):
        ctx.relax_on_batch(plain, ["x", "IIDLogNormal", "0", "0.25"])


# ---- DomainTreeGen ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gene_dir(tmp_path_factory):
    """Gene trees for the domain tests: pruned GuestTreeGen trees (oracle) on HOST8, as files gene1.tree..gene4.tree."""
    d = tmp_path_factory.mktemp("genes")
    for i in range(1, 5):
        f = oracle_run("GuestTreeGen", ["-s", str(i), "-db", "none", HOST8, "0.4", "0.2", "0.2", "g"])["files"]
        (d / f"gene{i}.tree").write_bytes(f["g.pruned.tree"])
    return str(d)


DOMAIN_CASES = [
    ["-s", "1"],                                               # defaults: -db none, -ig 0.5, -is 0.5, -rt 0.5
    ["-s", "2", "-ig", "0.9", "-is", "0.7", "-all"],            # README example shape: mostly inter-gene, all gene trees required
    ["-s", "3", "-rt", "1.0", "-ig", "0.3"],                     # replacing only
    ["-s", "4", "-rt", "0.0", "-is", "1.0"],                     # additive, always inter-species (many swaps inside one species)
    ["-s", "5", "-db", "simple", "-ig", "0.2"],                  # distance-biased intra-gene recipients
    ["-s", "6", "-db", "exponential", "-dbr", "2.0", "-ig", "0.0"],
    ["-s", "7", "-dbs", "-dbc", "2.0", "-p", "0.8"],             # domain birth sampling
    ["-s", "8", "-nox", "-vp", "Dom", "-min", "3", "-max", "60", "-maxper", "40"],
]


@pytest.mark.parametrize("opts", DOMAIN_CASES, ids=[" ".join(o[2:]) or "defaults" for o in DOMAIN_CASES])
def test_domain_tree_gen_replay_is_byte_exact(ctx, gene_dir, opts):
    args = opts + [HOST8, gene_dir, "0.4", "0.3", "0.9", "o"]
    assert_replay_parity(ctx, "DomainTreeGen", args, 24, "o")


def test_domain_tree_gen_philox_matches_oracle_distributions(ctx, gene_dir):
    n = 20000
    args = ["-s", "31", "-ig", "0.6", "-is", "0.5", HOST8, gene_dir, "0.4", "0.3", "0.8", "x"]
    b = ctx.run("DomainTreeGen", args, n=n, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    assert (b.rep_status == 0).all()
    sec, st = oracle_batch("DomainTreeGen", args, 0, n)
    assert chi2_two_sample_ok(b.sampled_leaves.astype(float), st[:, 4])
    assert ks_ok(b.total_times, st[:, 6])
    ev = b.event_counts
    for e in (1, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16):
        assert chi2_two_sample_ok(ev[:, e].astype(float), st[:, 9 + e]), e
    assert chi2_two_sample_ok(b.attempts.astype(float), st[:, 0])


# ---- chained stage: GuestTreeGen batch -> DomainTreeGen without files -----------------------------------------------------------------
@pytest.mark.parametrize("src_opts", [[], ["-nox"]], ids=["tagged", "nox"])
def test_chained_domain_tree_gen_equals_the_run_on_the_emitted_files(ctx, tmp_path, src_opts):
    """DomainTreeGen over the pruned trees of a resident GuestTreeGen batch (sgp_domain_run_on_batch: vertex records read on the
    device, no Newick text) == DomainTreeGen over a directory holding those trees as the files SGP_LAYOUT_PER_REPLICATE writes,
    which in turn == the CPU oracle on that directory (replay mode, byte for byte)."""
    n_genes = 12                                  # two-digit indices: the forest order is the sorted file-name order, not the index order
    gargs = ["-s", "40", "-db", "none"] + src_opts + [HOST8, "0.4", "0.2", "0.2", str(tmp_path / "gt")]
    src = ctx.run("GuestTreeGen", gargs, n=n_genes, rng=sg.RNG_MT_REPLAY, flags=sg.FLAG_KEEP_TREES)
    assert (src.rep_status == 0).all()
    d = tmp_path / "genes"; d.mkdir()
    for i in range(n_genes):
        (d / f"gt_{100 + i}.pruned.tree").write_bytes(src.replicate(sg.STREAM_PRUNED_TREE, i))
    for opts in (["-s", "3", "-ig", "0.6"], ["-s", "4", "-db", "simple", "-rt", "1.0", "-all", "-a", "200"]):
        args = opts + [HOST8, str(d), "0.4", "0.3", "0.9", "o"]
        files = ctx.run("DomainTreeGen", args, n=6, rng=sg.RNG_MT_REPLAY)
        chained = ctx.domain_on_batch(src, args, name_first_index=100, n=6, rng=sg.RNG_MT_REPLAY)
        assert (chained.rep_status == files.rep_status).all()
        for s in range(8):
            assert bytes(chained.stream(s)) == bytes(files.stream(s)), (opts, s)
        # and the file-based run is the oracle's (replicate 0)
        want = oracle_run("DomainTreeGen", args)
        if want["status"] == 0:
            got = files.files(0)
            for sfx, data in got.items():
                assert data == want["files"]["o" + sfx], sfx
        files.free(); chained.free()
    # a sub-range of the batch as the forest
    part = ctx.domain_on_batch(src, ["-s", "5", HOST8, "x", "0.4", "0.3", "0.5", "o"], first=2, count=5, name_first_index=2, n=3, rng=sg.RNG_MT_REPLAY)
    assert (part.rep_status == 0).all() and part.stream_bytes(sg.STREAM_PRUNED_TREE) > 0
    part.free(); src.free()
