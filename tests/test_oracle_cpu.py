"""CPU tests: pin the oracle (the reference has no tests of its own, so the pins are closed forms, independent
re-implementations and the worked example of SURVEY.md Appendix D) and check the C-ABI library loads."""
import ctypes as C
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle_util import (oracle, oracle_batch, oracle_double_to_string, oracle_files, oracle_mt_doubles, oracle_mt_words,
                         oracle_run)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ---- RNG -----------------------------------------------------------------------------------------------------------------
def java_seed_key(seed: int) -> np.ndarray:
    """BigInteger(seed).toByteArray() -> PRNG.get16bytes -> 4 big-endian words, with Python's own big integers."""
    nbytes = max(1, (seed.bit_length() + 8) // 8) if seed >= 0 else max(1, ((-seed - 1).bit_length() + 8) // 8)
    raw = seed.to_bytes(nbytes, "big", signed=True)
    # BigInteger.toByteArray is minimal: strip redundant sign bytes
    while len(raw) > 1 and ((raw[0] == 0 and raw[1] < 0x80) or (raw[0] == 0xFF and raw[1] >= 0x80)):
        raw = raw[1:]
    sixteen = (b"\0" * 16 + raw)[-16:]
    return np.frombuffer(sixteen, dtype=">u4").astype(np.uint32)


BIG_SEEDS = [0, 1, 12345, 2 ** 31, 2 ** 63 - 1, 2 ** 63, 2 ** 64 - 1, (1 << 70) + 5, -1, -129, -(1 << 40), -(1 << 63), -(1 << 63) - 1,
             2 ** 119 - 1, 2 ** 119, 2 ** 120 + 2 ** 64 - 1, 2 ** 120 + 2 ** 64, 2 ** 127 - 1, 2 ** 127, 2 ** 128 + 77, -(2 ** 119), -(2 ** 119) - 1,
             -(2 ** 120) - 2 ** 64, -(2 ** 127), -(2 ** 127) - 1, 10 ** 50 + 12345, -(10 ** 50), 10 ** 80 - 1, 10 ** 38 - 1,
             123456789012345678901234567890123456789999999999999999999, -(10 ** 60 + 5)]


@pytest.mark.parametrize("seed", BIG_SEEDS)
def test_mt_seeding_matches_numpy_init_by_array(seed):
    # PRNG.get16bytes + MersenneTwisterRNG2 ctor == MT19937 init_by_array over 4 big-endian words
    key = java_seed_key(seed)
    bg = np.random.MT19937()
    bg._legacy_seeding(key)
    want = bg.random_raw(2000).astype(np.uint32)
    got = oracle_mt_words(seed, 2000)
    assert (want == got).all()


@pytest.mark.parametrize("seed", BIG_SEEDS)
def test_library_seed_arithmetic_matches_big_integers(seed):
    # product host/device seed code (rng.cuh seed_words / put_seed, compiled for the host) against Python's arbitrary-precision
    # integers: replicate i of a batch is the reference run with -s (seed + i) (math/PRNG.java:29-37,58-60)
    import sagephy_b200 as sg
    for i in (0, 1, 7, 255, 256, 10 ** 6, 2 ** 40 + 3, 9 * 10 ** 18, 2 ** 63 - 1):
        key, text = sg.seed_probe(seed, i)
        assert text == str(seed + i), (seed, i)
        assert key == java_seed_key(seed + i).tolist(), (seed, i)
    key, text = sg.seed_probe("+" + str(abs(seed)), 0)
    assert text == str(abs(seed))
    for bad in ("", "-", "12x", "1.5", " 7"):
        with pytest.raises(sg.SagephyError):
            sg.seed_probe(bad, 0)


def test_oracle_seed_offset_is_big_integer_addition():
    # the oracle's batch entry point adds the replicate index in decimal-string arithmetic; check it through the MT stream
    for seed in (10 ** 50 - 3, -(10 ** 30) + 2, -5, 2 ** 64 - 2):
        for i in (0, 3, 5, 7):
            assert (oracle_mt_words(seed + i, 8) == oracle_mt_words(str(seed + i), 8)).all()
    a = oracle_run("HostTreeGen", ["-s", str(10 ** 40 + 2), "-min", "2", "-max", "9", "1.0", "2.0", "0.3", "x"])
    _, st = oracle_batch("HostTreeGen", ["-s", str(10 ** 40), "-min", "2", "-max", "9", "1.0", "2.0", "0.3", "x"], 2, 1)
    assert int(st[0, 0]) == a["attempts"]


def test_philox4x32_10_random123_known_answers():
    # Random123 kat_vectors (philox4x32, 10 rounds) through the host build of the library's generator, both key-schedule forms,
    # and against an independent big-integer restatement of the published round function
    import sagephy_b200 as sg
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]

    def ref(c, k):
        c = list(c); k = list(k)
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
            k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
        return c
    for c, k, want in kat:
        assert ref(c, k) == list(want)
        for variant in (0, 1):
            assert sg.philox_host([c], [k], variant)[0].tolist() == list(want)
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2 ** 32, (200, 4), dtype=np.uint64); key = rng.integers(0, 2 ** 32, (200, 2), dtype=np.uint64)
    got = sg.philox_host(ctr.astype(np.uint32), key.astype(np.uint32), 1)
    assert got.tolist() == [ref([int(x) for x in c], [int(x) for x in k]) for c, k in zip(ctr, key)]


def test_java_next_double_layout():
    w = oracle_mt_words(77, 2000).astype(np.uint64)
    d = oracle_mt_doubles(77, 1000)
    want = (((w[0::2] >> np.uint64(6)) << np.uint64(27)) + (w[1::2] >> np.uint64(5))).astype(np.float64) * 2.0 ** -53
    assert (d == want).all()
    assert ((d >= 0) & (d < 1)).all()


def test_java_next_int_bounded():
    L = oracle()
    for bound in (1, 2, 3, 7, 16, 100, 1000, 2 ** 30 + 1):
        out = np.empty(5000, dtype=np.int32)
        L.oracle_mt_ints(b"5", bound, 5000, out.ctypes.data)
        assert out.min() >= 0 and out.max() < bound
        if bound <= 16:
            assert len(set(out.tolist())) == bound
    # power of two path: (n * next(31)) >> 31
    w = oracle_mt_words(5, 10)
    out = np.empty(10, dtype=np.int32)
    L.oracle_mt_ints(b"5", 16, 10, out.ctypes.data)
    assert (out == ((16 * (w.astype(np.int64) >> 1)) >> 31)).all()


# ---- Double.toString / fdlibm ----------------------------------------------------------------------------------------------
JAVA_DOUBLE_STRINGS = {   # values documented for java.lang.Double.toString (JDK >= 19 shortest-repr behaviour)
    1.0: "1.0", 100.0: "100.0", 1e7: "1.0E7", 9999999.0: "9999999.0", 0.001: "0.001", 0.0001: "1.0E-4", 1e-5: "1.0E-5",
    1e23: "1.0E23", 2e23: "2.0E23", 4.9e-324: "4.9E-324", 1.7976931348623157e308: "1.7976931348623157E308", 0.1: "0.1",
    0.3: "0.3", 1.0 / 3.0: "0.3333333333333333", 2.0 / 3.0: "0.6666666666666666", 123456789.0: "1.23456789E8", 0.5: "0.5",
    1e-64: "1.0E-64", 2.2250738585072014e-308: "2.2250738585072014E-308", 1e21: "1.0E21", 12345.678: "12345.678",
    8.41e21: "8.41E21", 5e-324 * 3: "1.5E-323", 0.0: "0.0", -2.5: "-2.5", float("inf"): "Infinity", float("nan"): "NaN",
}


def test_double_to_string_known_values():
    for v, s in JAVA_DOUBLE_STRINGS.items():
        assert oracle_double_to_string(v) == s, (v, oracle_double_to_string(v), s)
    assert oracle_double_to_string(-0.0) == "-0.0"


def test_double_to_string_round_trips_and_is_shortest():
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.random(3000), -np.log1p(-rng.random(3000)) / 5.05, np.round(rng.random(2000), 8),
                           rng.random(1000) * 10.0 ** rng.integers(-30, 30, 1000)])
    for v in vals:
        s = oracle_double_to_string(float(v))
        assert float(s.replace("E", "e")) == v
        # Python's repr is also the shortest round-tripping decimal: same digit string
        digits = re.sub(r"[.\-]|E.*$", "", s).strip("0")
        pd = re.sub(r"[.\-]|e.*$", "", repr(float(v))).strip("0")
        assert digits == pd, (v, s, repr(float(v)))


def test_fdlibm_restatements_within_one_ulp():
    L = oracle()
    rng = np.random.default_rng(2)
    for x in np.concatenate([rng.random(5000), rng.random(2000) * 10.0 ** rng.integers(-200, 200, 2000)]):
        x = float(x)
        if x <= 0:
            continue
        a, b = L.oracle_log(x), math.log(x)
        assert abs(a - b) <= abs(math.ulp(b)), x
        a, b = L.oracle_log10(x), math.log10(x)
        assert abs(a - b) <= 2 * abs(math.ulp(b)) + 1e-300, x
    for y in (rng.random(3000) - 0.5) * 60:
        a, b = L.oracle_exp(float(y)), math.exp(float(y))
        assert abs(a - b) <= math.ulp(b)
    assert L.oracle_log(1.0) == 0.0 and L.oracle_exp(0.0) == 1.0 and L.oracle_log10(1000.0) == 3.0 and L.oracle_log10(1e22) == 22.0


def test_round_to_significant_figures():
    L = oracle()
    assert L.oracle_round_sig(0.123456789, 8) == 0.12345679
    assert L.oracle_round_sig(0.1085868017185789 + 0.24064737132287545, 8) == 0.34923417
    assert L.oracle_round_sig(31.875615123, 8) == 31.875615
    assert L.oracle_round_sig(0.0, 8) == 0.0
    assert L.oracle_round_sig(1.0, 8) == 1.0
    assert L.oracle_round_sig(-0.55233891499, 8) == -0.55233891


# ---- golden: SURVEY.md Appendix D (independent hand restatement by the surveyor) -----------------------------------------
def test_appendix_d_worked_example():
    args = ["-s", "1", "-min", "3", "-max", "4", "-p", "0.8", "1.0", "2.0", "0.5", "out"]
    r = oracle_run("HostTreeGen", args)
    f = r["files"]
    with open(os.path.join(GOLDEN, "survey_appendix_d.txt")) as fh:
        gold = dict(line.rstrip("\n").split("\t", 1) for line in fh if "\t" in line)
    assert f["out.unpruned.tree"].decode().strip() == gold["unpruned.tree"]
    assert f["out.pruned.tree"].decode().strip() == gold["pruned.tree"]
    info = f["out.unpruned.info"].decode()
    assert "Attempts:\t9\n" in info and "No. of vertices:\t11\n" in info and "No. of extant leaves:\t6\n" in info
    assert "No. of births:\t5\n" in info and "No. of deaths:\t0\n" in info and "Total branch time:\t3.1875615\n" in info
    assert "Birth ML estimate:\t1.5685972\n" in info and "Death ML estimate:\t0.0\n" in info
    pinfo = f["out.pruned.info"].decode()
    assert "No. of vertices:\t7\n" in pinfo and "Total branch time:\t2.585425\n" in pinfo and "ML estimate" not in pinfo
    assert info.startswith("# HOSTTREEGEN\nArguments:\t[-s, 1, -min, 3, -max, 4, -p, 0.8, 1.0, 2.0, 0.5, out]\n")


# ---- closed-form known answers ----------------------------------------------------------------------
def bd_pmf(n, lam, mu, T):
    E = math.exp((lam - mu) * T)
    beta = lam * (E - 1) / (lam * E - mu)
    alpha = mu * (E - 1) / (lam * E - mu)
    return alpha if n == 0 else (1 - alpha) * (1 - beta) * beta ** (n - 1)


def test_leaf_count_distribution_matches_birth_death_closed_form():
    lam, mu, T, n = 2.0, 0.5, 1.0, 20000
    sec, st = oracle_batch("HostTreeGen", ["-s", "100", "-min", "0", "-max", "100000", str(T), str(lam), str(mu), "x"], 0, n)
    leaves = st[:, 4].astype(int)
    assert (st[:, 1] == 0).all() and (st[:, 0] == 2).all()          # every first attempt is accepted; Attempts = trees + 1
    kmax = 15
    obs = np.array([(leaves == k).sum() for k in range(kmax)] + [(leaves >= kmax).sum()], dtype=float)
    p = np.array([bd_pmf(k, lam, mu, T) for k in range(kmax)]); p = np.append(p, 1 - p.sum())
    chi2 = ((obs - n * p) ** 2 / (n * p)).sum()
    assert chi2 < 45.0, chi2          # 15 dof: P(chi2 > 45) ~ 8e-5
    # E[#losses] = mu (E - 1) / (lam - mu)
    E = math.exp((lam - mu) * T)
    losses = st[:, 9 + 5].mean()
    assert abs(losses - mu * (E - 1) / (lam - mu)) < 0.05


def test_acceptance_rate_of_readme_example():
    # P(N = 100 | lam 5, mu 0.05, T 1) = 3.4593e-3  ->  289.08 trees per accepted tree (Attempts = trees + 1)
    sec, st = oracle_batch("HostTreeGen", ["-s", "500", "-min", "100", "-max", "100", "1.0", "5.0", "0.05", "x"], 0, 60)
    trees = st[:, 0] - 1
    assert (st[:, 4] == 100).all() and (st[:, 3] == 199).all()
    assert abs(trees.mean() - 289.08) < 4.5 * 289.08 / math.sqrt(60)


# ---- structural invariants of the emitted files ----------------------------------------------------------------------------------
def newick_leaves(s):
    return re.findall(r"[(,]([A-Za-z]+\d+(?:_\d+)?):", s)


def test_host_tree_gen_files_are_consistent():
    f = oracle_files("HostTreeGen", ["-s", "42", "-min", "20", "-max", "30", "-p", "0.7", "1.0", "4.0", "0.4", "o"])
    un, pr = f["o.unpruned.tree"].decode(), f["o.pruned.tree"].decode()
    assert un.endswith("[NAME=UnprunedTree];\n") and pr.endswith("[NAME=PrunedTree];\n")
    n_leaf = pr.count("VERTEXTYPE=Leaf]")
    assert 20 <= n_leaf <= 30 and "UnsampledLeaf" not in pr and "Loss" not in pr
    ids_pr = sorted(int(x) for x in re.findall(r"\[ID=(\d+) ", pr))
    assert ids_pr == list(range(2 * n_leaf - 1))                  # survivors are numbered 0..2L-2 (PruningHelper.java:25-65)
    ids_un = sorted(int(x) for x in re.findall(r"\[ID=(\d+) ", un))
    assert ids_un == list(range(len(ids_un)))
    bare = re.sub(r"\[[^\]]*\]", "", un)
    assert bare.count("(") == bare.count(")") == (len(ids_un) - 1) // 2


def test_guest_tree_gen_all_bias_modes_run_and_are_deterministic():
    host = "((A:0.4,B:0.4):0.6,(C:0.7,(D:0.2,E:0.2):0.5):0.3):0.2;"
    for extra in (["-db", "none"], ["-db", "simple"], ["-db", "exponential", "-dbr", "1.5"], ["-gb", "-gbc", "2.1"], ["-rt", "1.0"], ["-nox"]):
        args = ["-s", "9"] + extra + [host, "0.6", "0.4", "0.8", "g"]
        a, b = oracle_run("GuestTreeGen", args), oracle_run("GuestTreeGen", args)
        assert a["status"] == 0, a["stderr"]
        assert a["files"] == b["files"]
        names = set(a["files"])
        assert {"g.unpruned.tree", "g.pruned.tree", "g.unpruned.info", "g.pruned.info"} <= names
        assert ("g.pruned.guest2host" in names) == ("-nox" not in extra)
        if "-nox" not in extra:
            g2h = a["files"]["g.unpruned.guest2host"].decode()
            assert g2h.startswith("# GUEST-TO-HOST MAP\nHost tree:\t(")
            assert "DISCTYPE=RBTreeEpochDiscretiser NMIN=1 NMAX=1 DELTAT=NaN NROOT=1];" in g2h
            lm = a["files"]["g.pruned.leafmap"].decode().strip().split("\n")
            assert all(re.fullmatch(r"G\d+_\d+\t[A-E]", row) for row in lm)


def test_replacing_transfers_produce_replacing_losses():
    host = "((A:0.4,B:0.4):0.6,(C:0.7,(D:0.2,E:0.2):0.5):0.3):0.2;"
    tot_rt = tot_rl = 0
    for seed in range(1, 30):
        info = oracle_files("GuestTreeGen", ["-s", str(seed), "-rt", "1.0", "-db", "none", "-max", "100000", "-maxper", "100000", host, "0.5", "0.3", "1.0", "g"])["g.unpruned.info"].decode()
        tot_rt += int(re.search(r"replacing transfers:\t(\d+)", info).group(1))
        tot_rl += int(re.search(r"replacing losses:\t(\d+)", info).group(1))
        assert int(re.search(r"additive transfers:\t(\d+)", info).group(1)) >= 0
    assert tot_rt > 0 and tot_rl > 0 and tot_rl <= tot_rt      # a victim may later be detached together with its ancestor's subtree


def test_max_attempts_failure_and_usage():
    r = oracle_run("HostTreeGen", ["-s", "1", "-min", "500", "-max", "500", "-a", "5", "1.0", "1.0", "0.5", "f"])
    assert r["status"] == 1 and r["files"] == {"f.info": b"Failed to produce valid pruned tree within max allowed attempts.\n"}
    assert r["attempts"] == 6
    assert oracle_run("HostTreeGen", ["1.0", "1.0"])["stdout"].startswith("Usage:")
    assert oracle_run("GuestTreeGen", ["-h"])["stdout"].startswith("Usage:")


# ---- the product library: loads on a CPU-only box, exports its ABI, and refuses to run without a GPU ---------------------------------
def test_c_abi_library_exports_every_declared_symbol():
    import sagephy_b200 as sg
    lib = sg.load_library()
    header = open(os.path.join(ROOT, "include", "sagephy_b200.h")).read()
    declared = set(re.findall(r"\b(sgp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(sg.EXPORTED_SYMBOLS), declared ^ set(sg.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_no_cpu_fallback_without_device():
    import torch
    import sagephy_b200 as sg
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sg.SagephyError):
        sg.Context(0)


def test_unpack_batch_tool_splits_packed_streams(tmp_path):
    # tools/unpack_batch.py: packed stream + .idx -> one file per replicate, named like separate reference runs
    import importlib.util
    spec = importlib.util.spec_from_file_location("unpack_batch", os.path.join(ROOT, "tools", "unpack_batch.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    prefix = str(tmp_path / "run")
    parts = [b"(a:1,b:2);\n", b"", b"((x:1,y:1):1,z:2);\n"]
    notes = [b"# info 0\n", b"Failed to produce valid pruned tree within max allowed attempts.\n", b"# info 2\n"]
    for sfx, chunks in ((".pruned.tree", parts), (".pruned.info", notes)):
        with open(prefix + sfx, "wb") as f:
            f.write(b"".join(chunks))
        with open(prefix + sfx + ".idx", "w") as f:
            f.write("#sagephy-idx app=HostTreeGen n=3 first=10\n")
            off = 0
            for i, c in enumerate(chunks):
                f.write(f"{off}\t{len(c)}\t{1 if i == 1 else 0}\n"); off += len(c)
    assert mod.unpack(prefix) == 5
    assert (tmp_path / "run_10.pruned.tree").read_bytes() == parts[0] and (tmp_path / "run_12.pruned.tree").read_bytes() == parts[2]
    assert not (tmp_path / "run_11.pruned.tree").exists()
    assert (tmp_path / "run_11.info").read_bytes() == notes[1]          # HostTreeGen's failure note goes to <prefix>.info
    assert (tmp_path / "run_12.pruned.info").read_bytes() == notes[2]


def test_autocorrelated_rate_models_have_the_documented_means():
    # closed-form pins for the oracle's restatement of the AC models (apps/SaGePhy/AC*RateModel.java):
    #   ACTK98 (corrected):  E[r(x) | r(parent)] = r(parent)              -> E[r] = start rate everywhere
    #   ACRY07:              same martingale property for mid-point and vertex rates
    #   ACABY02:             r(x) ~ Exp(mean r(parent))                   -> E[r] = start rate
    #   ACLBPL07 (CIR):      mean reverts to mu: E[r_t] = mu + (r0 - mu) e^{-theta t}
    import re
    tree = "((a:0.3,b:0.3):0.4,c:0.7):0.2;"           # vertex ids (post-order): a=0 b=1 ab=2 c=3 root=4
    n = 4000

    def rates(model_args):
        out = np.empty((n, 5))
        for i in range(n):
            r = oracle_run("BranchRelaxer", ["-s", str(1000 + i), "-o", "/x", tree] + model_args)
            info = r["files"]["/x.info"].decode()
            out[i] = [float(v) for v in re.search(r"Rates:\t\[(.*)\]", info).group(1).split(", ")]
        return out

    for args, tol in ((["ACTK98", "1.5", "0.5"], 0.04), (["ACRY07", "1.5", "0.5"], 0.04), (["ACABY02", "1.5"], 0.12)):
        r = rates(args)
        for x in (0, 2, 3):
            assert abs(r[:, x].mean() - 1.5) < tol * 1.5 * (3 if args[0] == "ACABY02" and x == 0 else 1), (args, x, r[:, x].mean())
    if True:
        start, mu, theta, sigma = 0.5, 1.2, 4.0, 0.5
        r = rates(["ACLBPL07", str(start), str(mu), str(theta), str(sigma)])
        # branch rate = time average of the instantaneous rate over the branch; stem [0, 0.2], then c on [0.2, 0.9]
        def avg(t0, t1):
            return mu + (start - mu) * (math.exp(-theta * t0) - math.exp(-theta * t1)) / (theta * (t1 - t0))
        assert abs(r[:, 4].mean() - avg(0.0, 0.2)) < 0.02, (r[:, 4].mean(), avg(0.0, 0.2))
        assert abs(r[:, 3].mean() - avg(0.2, 0.9)) < 0.02, (r[:, 3].mean(), avg(0.2, 0.9))


def test_domain_tree_gen_on_one_gene_tree_matches_guest_tree_gen(tmp_path):
    # Two independent restatements (DomainTreeInHostTreeCreator vs GuestTreeInHostTreeCreator) of the same process: with a single
    # gene tree and intra-gene transfers only, a domain tree evolves inside that gene tree exactly as a gene tree evolves inside a
    # species tree (both species-blind here: one HOST tag). The random streams are consumed differently (first_guest draw, three
    # extra uniforms per transfer, recipient before staying child), so only distributions can agree.
    from scipy import stats
    species = "((A:0.6,B:0.6):0.4,C:1.0):0.1;"
    # zero stem: GuestTreeGen discards the stem of its host tree, DomainTreeGen lets the domain evolve along the gene tree's stem
    gene = "(((a:0.1,b:0.1):0.5,(c:0.3,d:0.3):0.3):0.4,((e:0.2,f:0.2):0.6,(g:0.5,h:0.5):0.3):0.2):0.0;"
    gd = tmp_path / "genes"; gd.mkdir()
    (gd / "g0.tree").write_text(gene + "\n")
    n = 4000
    common = ["-min", "1", "-max", "100000", "-minper", "0", "-maxper", "100000", "-db", "none"]
    _, g = oracle_batch("GuestTreeGen", ["-s", "1"] + common + [gene, "0.6", "0.4", "0.5", "x"], 0, n)
    _, d = oracle_batch("DomainTreeGen", ["-s", "900001"] + common + ["-ig", "0.0", "-is", "0.0", species, str(gd), "0.6", "0.4", "0.5", "x"], 0, n)
    assert (g[:, 1] == 0).all() and (d[:, 1] == 0).all()
    for col in (0, 4, 6, 7):       # attempts, sampled leaves, total branch time, vertices created
        assert stats.ks_2samp(g[:, col], d[:, col]).pvalue > 0.001, (col, g[:, col].mean(), d[:, col].mean())


def test_committed_bench_line_follows_the_contract():
    # profiles/r01_bench_final.json is the line `python bench.py` printed on a B200 at the final commit of the round
    import json
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["config"]["workload"].startswith("C2") and d["scaling"] == "weak" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"]) and 0 < d["roofline"]["frac"] < 1
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert not (set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
    r = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_reference_arm.json")))
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"] and r["e2e"]["h2d_bytes_per_step"] == 0


def test_round2_bench_line_carries_every_baseline_config():
    # profiles/r02_bench.json: the line `python bench.py` printed on a B200 at the end of round 2
    import json
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "e2e", "gpu_launches", "roofline", "roofline_emit", "cpu_baseline", "clocks", "workloads"):
        assert k in d, k
    assert d["config"]["workload"].startswith("C2") and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9 and d["roofline"]["traffic"] is not None
    assert {"link_gbs_alone", "link_gbs_concurrent_per_gpu_mean", "frac_of_link"} <= set(d["e2e"])
    assert not (set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
    want = {"C3_db_none", "C3_db_simple", "C4_guest_db_exponential", "C4_guest_db_simple", "C4_guest_db_none", "C4_domain_100_genes", "C5_lognormal", "C5_gamma"}
    assert want <= set(d["workloads"])
    for name in want:
        w = d["workloads"][name]
        assert w["value"] > 0 and w["e2e"]["value"] > 0 and w["e2e"]["d2h_bytes_per_step"] > 0, name
        assert {"bound", "achieved", "peak", "frac"} <= set(w["roofline"]) and w["cpu_baseline"]["value"] > 0, name
    assert "chained" in d["workloads"]["C4_domain_100_genes"]


def test_launch_list_summary_tool():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), os.path.join(ROOT, "profiles", "r01_launches_bench_c2.csv")],
                         capture_output=True, text=True, check=True).stdout
    first = out.splitlines()[1]
    assert "k_bd_find_single" in first and float(first.split()[-1].rstrip("%")) > 50      # the sampler dominates the step


def _random_ultrametric_newick(rng, n_leaves):
    """Random binary ultrametric tree (coalescent-like heights), Newick with branch lengths and a stem."""
    nodes = [(f"L{i}", 0.0) for i in range(n_leaves)]
    t = 0.0
    while len(nodes) > 1:
        t += float(rng.exponential(1.0 / (len(nodes) * (len(nodes) - 1))))
        i = int(rng.integers(len(nodes) - 1))            # adjacent pairs keep the tree planar-consistent with its leaf order
        (a, ta), (b, tb) = nodes[i], nodes[i + 1]
        nodes[i:i + 2] = [(f"({a}:{t - ta!r},{b}:{t - tb!r})", t)]
    return nodes[0][0] + ":0.05;"


def test_host_tables_of_the_product_without_a_gpu():
    # K0 (host_model.cpp) through sgp_probe_host_model: the LCA-time matrix against a naive ancestor walk, and the epoch
    # tables against their defining properties (RBTreeEpochDiscretiser.java:146-221): epochs partition time in ascending order,
    # epoch e holds the arcs that are contemporaries during it, a vertex's epoch starts at the vertex's time.
    lib = C.CDLL(os.path.join(ROOT, "sagephy_b200", "libsagephy_b200.so"))
    f = lib.sgp_probe_host_model
    f.restype = C.c_int
    rng = np.random.default_rng(12)
    for n_leaves in (2, 3, 8, 37, 150):
        nw = _random_ultrametric_newick(rng, n_leaves)
        nV, nE = 2 * n_leaves - 1, n_leaves
        i32 = lambda k: np.zeros(k, dtype=np.int32)
        parent, left, right, v2e = i32(nV), i32(nV), i32(nV), i32(nV)
        vtime, ep_lo, ep_up, ep_off = np.zeros(nV), np.zeros(nE), np.zeros(nE), i32(nE + 1)
        ep_arcs, lca = i32(nE * (nE + 1) // 2 + 8), np.zeros(nV * nV)
        nv_o, ne_o = C.c_int32(), C.c_int32()
        err = C.create_string_buffer(256)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = f(nw.encode(), C.c_double(0.0), C.byref(nv_o), C.byref(ne_o), P(parent), P(left), P(right), P(v2e), P(vtime), C.c_int32(nV),
               P(ep_lo), P(ep_up), P(ep_off), C.c_int32(nE), P(ep_arcs), C.c_int32(len(ep_arcs)), P(lca), C.c_int64(nV * nV), err, C.c_int32(256))
        assert rc == 0, err.value
        assert nv_o.value == nV and ne_o.value == nE
        root = int(np.where(parent < 0)[0][0])
        # naive LCA
        def anc(v):
            out = [v]
            while parent[out[-1]] >= 0:
                out.append(int(parent[out[-1]]))
            return out
        chains = [anc(v) for v in range(nV)]
        for x in range(nV):
            sx = set(chains[x])
            for y in range(x, nV):
                a = next(v for v in chains[y] if v in sx)
                assert lca[x * nV + y] == vtime[a] == lca[y * nV + x], (n_leaves, x, y)
        # epochs
        assert (np.diff(ep_up) >= 0).all() and (ep_lo[1:] == ep_up[:-1]).all() and ep_lo[0] == 0.0
        assert ep_off[0] == 0 and (np.diff(ep_off) == np.arange(nE, 0, -1)).all()            # one arc fewer per epoch, 1 arc (the stem) at the top
        for v in range(nV):
            assert vtime[v] == ep_lo[v2e[v]]
        for e in range(nE):
            arcs = ep_arcs[ep_off[e]:ep_off[e + 1]]
            for x in arcs:          # arc x = (x, parent(x)) spans the epoch
                top = vtime[parent[x]] if parent[x] >= 0 else ep_up[-1]
                assert vtime[x] <= ep_lo[e] and top >= ep_up[e] - 1e-12, (e, x)
        assert list(ep_arcs[ep_off[nE - 1]:ep_off[nE]]) == [root]
    # a tree without a stem keeps working for GuestTreeGen-style reading (stem override 0 -> 1e-64) ...
    rc = f(b"((a:1,b:1):1,c:2);", C.c_double(0.0), C.byref(nv_o), C.byref(ne_o), *([None] * 5), C.c_int32(0), None, None, None, C.c_int32(0), None, C.c_int32(0), None, C.c_int64(0), err, C.c_int32(256))
    assert rc == 0 and nv_o.value == 5
    # ... and is refused when the stem is taken from the tree (DomainTreeGen's gene trees) and is missing there
    rc = f(b"((a:1,b:1):1,c:2);", C.c_double(float("nan")), C.byref(nv_o), C.byref(ne_o), *([None] * 5), C.c_int32(0), None, None, None, C.c_int32(0), None, C.c_int32(0), None, C.c_int64(0), err, C.c_int32(256))
    assert rc == 3 and b"top time" in err.value


# ---- IIDRK11: closed forms of the host-specific gamma rates ------------------------------------------------------
RK_HOST = "((A:0.4[PARAMS=(2.0,0.5)],B:0.4[PARAMS=(3.0,0.4)]):0.6[PARAMS=(1.5,0.7)],(C:0.7[PARAMS=(2.5,0.3)],(D:0.2[PARAMS=(4,0.25)],E:0.2[PARAMS=(2,0.6)]):0.5[PARAMS=(1.0,1.0)]):0.3[PARAMS=(2.0,0.5)]);"


def test_iidrk11_oracle_closed_forms(tmp_path):
    # a guest tree that IS the host tree (one gene per species, guest vertices at the host vertices' times): the branch above guest
    # vertex x passes over exactly the host arc above sigma(x), so rate(x) ~ scale * Gamma(k, theta) of that arc
    # (apps/SaGePhy/IIDRasmussenKellis11RateModel.java:110-157); the root arc without PARAMS defaults to Gamma(1000, 0.001) (:77-79)
    guest = "((a:0.4,b:0.4):0.6,(c:0.7,(d:0.2,e:0.2):0.5):0.3);"
    m = tmp_path / "map.txt"; m.write_text("a A\nb\tB\nc C\nd  D\ne E\n\nignored line\n")
    scale = 1.5
    par = [(2.0, 0.5), (3.0, 0.4), (1.5, 0.7), (2.5, 0.3), (4, 0.25), (2, 0.6), (1.0, 1.0), (2.0, 0.5), (1000, 0.001)]   # vertex ids = post-order
    n = 3000
    rates = np.empty((n, 9))
    for i in range(n):
        r = oracle_run("BranchRelaxer", ["-s", str(100 + i), "-o", "x", guest, "IIDRK11", RK_HOST, str(m), str(scale)])
        if r["status"] == 2:        # a uniform outside [2e-6, 0.999998] aborts the reference run (math/ChiSquared.java:24)
            rates[i] = np.nan; continue
        line = [l for l in r["files"]["x.info"].decode().split("\n") if l.startswith("Rates:")][0]
        rates[i] = [float(v) for v in line.split("[")[1].rstrip("]").split(", ")]
        assert "Model:\tIIDRasmussenKellis11Rates\nModel parameters:\n" in r["files"]["x.info"].decode()
    rates = rates[~np.isnan(rates[:, 0])]
    assert len(rates) > 0.99 * n
    for x, (k, th) in enumerate(par):
        want = scale * k * th
        assert abs(rates[:, x].mean() - want) < 5 * scale * math.sqrt(k) * th / math.sqrt(len(rates)), (x, rates[:, x].mean(), want)
    # a guest branch that passes over two host arcs mixes their rates with the time shares as weights: guest "x" (mapped to D) joins
    # "y" (mapped to E) at time 0.2 -> sigma = their parent; the branch above that duplication-free vertex runs from 0.2 to the guest
    # root at 1.0 and passes over the host arcs above (D,E) [0.2, 0.7] and above its parent [0.7, 1.0]
    guest2 = "((x:0.2,y:0.2):0.8,c:1.0);"
    m2 = tmp_path / "map2.txt"; m2.write_text("x D\ny E\nc C\n")
    got = []
    for i in range(2000):
        r = oracle_run("BranchRelaxer", ["-s", str(i), "-o", "x", guest2, "IIDRK11", RK_HOST, str(m2), "1.0"])
        if r["status"] == 0:
            line = [l for l in r["files"]["x.info"].decode().split("\n") if l.startswith("Rates:")][0]
            got.append([float(v) for v in line.split("[")[1].rstrip("]").split(", ")])
    got = np.array(got)
    want = (0.5 / 0.8) * 1.0 * 1.0 + (0.3 / 0.8) * 2.0 * 0.5
    assert abs(got[:, 2].mean() - want) < 0.06, (got[:, 2].mean(), want)
    # constructor errors of the reference: leaf missing from the map, unknown host leaf, no PARAMS anywhere, time scales that disagree
    for bad_map, host, gtree in (("a A\nb B\nc C\nd D\n", RK_HOST, guest), ("a A\nb B\nc C\nd D\ne Z\n", RK_HOST, guest),
                                 ("a A\nb B\nc C\nd D\ne E\n", "((A:0.4,B:0.4):0.6,(C:0.7,(D:0.2,E:0.2):0.5):0.3);", guest),
                                 ("a A\nb B\nc C\nd D\ne E\n", RK_HOST, "((a:0.4,b:0.4):0.7,(c:0.8,(d:0.2,e:0.2):0.6):0.3);")):
        mm = tmp_path / "bad.txt"; mm.write_text(bad_map)
        assert oracle_run("BranchRelaxer", ["-s", "1", gtree, "IIDRK11", host, str(mm), "1.0"])["status"] == 2


def test_quantile_restatements_against_scipy_within_the_algorithms_tolerances():
    """Independent pin of the two quantile restatements BranchRelaxer's rates go through (the device versions are bit-identical to
    the oracle's, tests/test_gpu_parity.py): Voutier's rational approximation of the normal quantile (math/Normal.java:48-86; the
    paper states an absolute error below 1.16e-4 in the central region 0.025 <= p <= 0.975 and a relative error below 2.5e-5 in the
    tails) and Best & Roberts' AS 91 chi-square quantile with AS 32 / Lanczos underneath (math/ChiSquared.java:23-84,
    math/Gamma.java:127-204; about six significant figures)."""
    from scipy import stats
    from oracle_util import oracle
    L = oracle()
    p = np.concatenate([np.linspace(1e-6, 0.025, 400), np.linspace(0.025, 0.975, 2000), np.linspace(0.975, 1 - 1e-6, 400)])
    got = np.array([L.oracle_normal_quantile(float(v)) for v in p]); want = stats.norm.ppf(p)
    central = (p >= 0.025) & (p <= 0.975)
    assert np.max(np.abs(got[central] - want[central])) < 1.2e-4          # measured: 1.16e-4, the published bound
    assert np.max(np.abs(got[~central] / want[~central] - 1.0)) < 3e-5
    assert abs(L.oracle_normal_quantile(0.5)) < 1.2e-4 and L.oracle_normal_quantile(0.3) < 0 < L.oracle_normal_quantile(0.7)
    q = np.linspace(2e-6, 0.999998, 3000)     # ChiSquared.quantile throws outside [2e-6, 0.999998]
    for k, theta in ((2.0, 0.5), (0.25, 2.0), (4.0, 1.0), (50.0, 0.1)):
        got = np.array([L.oracle_gamma_quantile(float(v), k, theta) for v in q]); want = stats.gamma.ppf(q, a=k, scale=theta)
        assert np.max(np.abs(got / want - 1.0)) < 1e-6, (k, theta, float(np.max(np.abs(got / want - 1.0))))      # measured: <= 2e-7
