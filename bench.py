#!/usr/bin/env python3
"""Benchmark of the accelerated SaGePhy hot path (see BASELINE.json / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                       (CPU reference arm: the reference algorithm on host cores)

Headline line = BASELINE config C2: one "step" is one pass of the whole path (sample -> label/prune -> emit, outputs resident
in HBM) over one batch of synthetic HostTreeGen replicates (T=1, birth 5, death 0.05, 100-leaf window, leaf sampling 0.8).
The same JSON line carries a `workloads` object with the other BASELINE configs, each measured the same way on a bounded
batch: C3 (GuestTreeGen on a 100-leaf species tree, -db none / simple), C4 (GuestTreeGen with the three recipient-bias modes on
a 1000-leaf species tree; DomainTreeGen over 100 gene trees) and C5 (BranchRelaxer lognormal / gamma over 199-vertex trees,
chained on the device behind the generator).
Replicates are sharded by index range across ranks (no data-path collective); the only collective is the NCCL
all-reduce(sum) of the summary histograms / counters.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C2: HostTreeGen -min 100 -max 100 -p 0.8 -a 10000 1.0 5.0 0.05, Philox throughput mode, 4 output files/replicate"
C2_ARGS = ["-s", "20260101", "-min", "100", "-max", "100", "-p", "0.8", "-a", "10000", "1.0", "5.0", "0.05", "bench"]
SPECIES100 = os.path.join(ROOT, "tests", "golden", "species100.pruned.tree")
METRIC, UNIT = "accepted trees/sec", "trees/s"

# Algorithmic instruction mix of one vertex expansion of the sampler (DESIGN.md "K1 roofline"), in warp-level instructions:
#   FP64 pipe  : 1-u, table-driven log (r = m/c - 1, degree-6 log1p, exponent and table terms: 10), times 1/(lambda+mu),
#                start - bt, bits -> second uniform, 3 compares                                                        = 17
#   IMAD.WIDE  : Philox4x32-10, two 32x32->64 products per round                                                       = 20
#   other      : 20 three-input XORs of Philox, 4 for the two 52-bit mantissas, table index + shared-memory load (4),
#                counters / leaf count / stack / selects                                                               = 44
# FP64 and IMAD issue at half rate on B200 (measured below), so the bound is the sum of the three dispatch times.
ALG_FP64_INST_PER_VERTEX = 17.0
ALG_IMAD_INST_PER_VERTEX = 20.0
ALG_ALU_INST_PER_VERTEX = 44.0
# DRAM traffic per replicate of the dominant kernels comes from the ncu capture summarised in this file (written by
# tools/ncu_traffic.py from the `ncu --set full` CSV it names); absent file -> traffic is reported as null.
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")


def clocks_sampler(stop, rows, gpu_index):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(gpu_index)],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            if out:
                rows.append([x.strip() for x in out.split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def bind_to_gpu_numa_node(gpu_index):
    """Pins this rank's threads to the CPU cores of the NUMA node its GPU hangs off, so that the pinned host buffers of the
    end-to-end pipeline are allocated (first touch) next to the GPU's PCIe root port. Returns (node or None, reason)."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)], capture_output=True, text=True, timeout=5).stdout.strip()
        bus = out.lower()
        if not bus:
            return None, "nvidia-smi returned no pci.bus_id"
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]                      # nvidia-smi prints an 8-digit domain, sysfs uses 4
        path = f"/sys/bus/pci/devices/{bus}/numa_node"
        if not os.path.exists(path):
            return None, f"{path} does not exist (device not visible in this container's sysfs)"
        node = int(open(path).read().strip())
        if node < 0:
            nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] if os.path.isdir("/sys/devices/system/node") else []
            return None, f"sysfs reports numa_node = -1 for {bus} ({len(nodes)} NUMA node(s) visible: the platform exposes no GPU affinity)"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return None, f"no allowed CPU on node {node}"
        os.sched_setaffinity(0, allowed)
        return node, f"bound to {len(allowed)} CPUs of node {node}"
    except Exception as e:          # noqa: BLE001 - best effort, the reason is reported
        return None, f"{type(e).__name__}: {e}"


def summarise_clocks(rows):
    if not rows:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
    mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in rows:
        for k, nm in enumerate(names):
            if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                reasons.add(nm)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(rows)}


# ---- CPU side: the reference algorithm (C++ oracle port of the Java path) ------------------------------------------------------------
def _cpu_worker(job):
    app, args, first, n = job
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_util import oracle_batch
    sec, st = oracle_batch(app, args, first, n)
    return int((st[:, 1] == 0).sum()), sec


def cpu_reference(app, args, n_trees, procs):
    """Times the reference algorithm on `procs` host processes over disjoint replicate ranges. Returns (trees/s, ok, wall)."""
    from concurrent.futures import ProcessPoolExecutor
    per = max(1, n_trees // procs)
    t0 = time.perf_counter()
    import multiprocessing
    with ProcessPoolExecutor(procs, mp_context=multiprocessing.get_context("spawn")) as ex:
        res = list(ex.map(_cpu_worker, [(app, args, i * per, per) for i in range(procs)]))
    wall = time.perf_counter() - t0
    ok = sum(r[0] for r in res)
    return ok / wall, ok, wall


def cpu_single(app, args, n):
    """Single-thread sample of the oracle port: (trees/s, description)."""
    ok, sec = _cpu_worker((app, args, 0, n))
    return (ok / sec if sec > 0 else None), f"{n} replicates of the same arguments (MT19937 stream), single thread, C++ restatement of the Java path"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, cores)
    per_step = max(procs * 64, 256)         # ~1-2 s of work on all host cores per step
    for _ in range(args.warmup):
        cpu_reference("HostTreeGen", C2_ARGS, max(procs, per_step // 8), procs)
    # the process pool's start-up is outside the timed region: only the workers' own loop time counts (max over workers per step)
    t_sum = 0.0; ok = 0
    for _ in range(args.steps):
        from concurrent.futures import ProcessPoolExecutor
        per = max(1, per_step // procs)
        import multiprocessing
        with ProcessPoolExecutor(procs, mp_context=multiprocessing.get_context("spawn")) as ex:
            res = list(ex.map(_cpu_worker, [("HostTreeGen", C2_ARGS, i * per, per) for i in range(procs)]))
        ok += sum(r[0] for r in res); t_sum += max(r[1] for r in res)
    value = ok / t_sum
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_sum / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "replicates_per_step": per_step, "note": "reference-algorithm CPU restatement (C++ oracle, MT19937 stream, one process per host core, "
                       "time = slowest worker's loop); the Java jar cannot run here (no JVM)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": f"{per_step} replicates per step x {args.steps} steps of the same workload"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


# ---- GPU side ----------------------------------------------------------------------------------------------------------------------
class Harness:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import sagephy_b200 as sg
        self.torch, self.dist, self.sg = torch, dist, sg
        self.world = int(os.environ.get("WORLD_SIZE", "1")); self.rank = int(os.environ.get("RANK", "0")); self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
        torch.cuda.set_device(self.local)
        self.numa_node, self.numa_reason = bind_to_gpu_numa_node(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.ctx = sg.Context(self.local)
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def allsum(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return [float(x) for x in t.cpu()]

    def merge(self, summary):
        # NCCL all-reduce(sum) of the int64 summary vector: the only collective on the path
        from sagephy_b200.sharding import allreduce_summary
        return self.sg.Summary.from_vector(allreduce_summary(summary.as_vector(), device="cuda"))

    def run(self, app, argv, n, step, **kw):
        # global replicate ids: step-major, then rank-major index ranges
        off = (step * self.world + self.rank) * n
        return self.ctx.run(app, argv, n=n, rng=self.sg.RNG_PHILOX, offset=off, **kw)

    # device-resident throughput: W warm-up steps, K timed steps between barrier + synchronize, max over ranks
    def timed_device(self, make_batch, steps, warm):
        for i in range(warm):
            b = make_batch(10_000 + i); b.free()
        self.barrier()
        t0 = time.perf_counter()
        tim = {}; launches = 0; h2d = 0; total = None
        for i in range(steps):
            b = make_batch(i)
            t = b.timings()
            for k, v in t.items():
                if k.endswith("_ms"):
                    tim[k] = tim.get(k, 0.0) + v
            launches += int(t["kernel_launches"]); h2d += int(t["h2d_bytes"])
            s = b.summary().as_vector()
            total = s if total is None else total + s
            b.free()
        self.barrier()
        wall = time.perf_counter() - t0
        return {"wall": wall, "tim": tim, "launches": launches, "h2d": h2d, "summary": self.sg.Summary.from_vector(total)}

    # end to end through the public API with host buffers: argv in, every output stream's bytes + offsets out to pinned host
    # memory. Double-buffered: the device->host copies of chunk i run on the context's copy stream while chunk i+1 computes.
    def timed_e2e(self, make_batch, ne, e_steps, probe_step=50_000):
        torch = self.torch
        b = make_batch(probe_step); sizes = [b.stream_bytes(s) for s in range(8)]; h2d_probe = int(b.timings()["h2d_bytes"]); b.free()
        pinned = [[torch.empty(int(sz * 1.10) + 65536, dtype=torch.uint8, pin_memory=True) if sz else None for sz in sizes] for _ in range(2)]
        offs = [[torch.empty(ne + 1, dtype=torch.int64, pin_memory=True) for _ in range(8)] for _ in range(2)]
        warm_prev = None
        for i in range(2):          # untimed: the two sets of device buffers the pipeline alternates between come into being here
            b = make_batch(59_000 + i)
            if warm_prev is not None:
                warm_prev.free()
            warm_prev = b
        warm_prev.free()
        self.barrier(); t0 = time.perf_counter(); e_ok = 0; d2h = 0; h2d = 0; prev = None
        for i in range(e_steps):
            b = make_batch(60_000 + i)
            if prev is not None:
                self.ctx.wait_copies(); prev.free()
            # the small status fetch goes first: issued after the bulk copies it would queue behind them on the copy engine
            st = b.rep_status; d2h += st.nbytes * 3
            e_ok += int((st == 0).sum()); h2d += int(b.timings()["h2d_bytes"])
            buf, ob = pinned[i & 1], offs[i & 1]
            for s in range(8):
                if buf[s] is not None and b.stream_bytes(s):
                    d2h += b.copy_stream_into_async(s, buf[s].data_ptr(), buf[s].numel())
                    b.copy_offsets_into_async(s, ob[s].data_ptr()); d2h += 8 * (ne + 1)
            prev = b
        self.ctx.wait_copies(); prev.free()
        self.barrier(); e_wall = time.perf_counter() - t0
        wall_max, = self.allmax([e_wall]); ok_sum, d2h_sum, h2d_sum = self.allsum([e_ok, d2h, h2d])
        del pinned, offs
        try:          # give the pinned pages back to the OS: torch caches them, and eight ranks times every workload's buffers add up
            torch._C._host_emptyCache()
        except Exception:
            pass
        return {"value": ok_sum / wall_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d_sum / e_steps / self.world), "d2h_bytes_per_step": int(d2h_sum / e_steps / self.world),
                "replicates_per_gpu_per_step": ne, "steps": e_steps, "d2h_gbs_per_gpu": d2h_sum / self.world / wall_max * 1e-9}


def hbm_peak():
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        try:
            return float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    try:
        return json.load(open(TRAFFIC_FILE))
    except Exception:
        return None


def sampler_peak_gverts(peaks, world):
    # issue roofline of the sampler: time of one warp-vertex = its FP64, IMAD and other instructions at their measured
    # issue rates (the ALU rate follows from the IMAD+SHF+LOP3 mix probe: 3/R_mixed = 1/R_imad + 2/R_alu)
    r_fp64 = peaks["dfma_ginst_per_s"]; r_imad = peaks["imad_ginst_per_s"]; r_mixed = peaks["mixed_int_ginst_per_s"]
    r_alu = 2.0 / max(3.0 / r_mixed - 1.0 / r_imad, 1e-9)
    t_vertex_warp = ALG_FP64_INST_PER_VERTEX / r_fp64 + ALG_IMAD_INST_PER_VERTEX / r_imad + ALG_ALU_INST_PER_VERTEX / r_alu     # ns per warp of 32 vertices
    return 32.0 / t_vertex_warp * world, r_alu


def pcie_probe(h, nbytes=1 << 30, reps=3):
    """Device->host link rate with the e2e leg's own mechanism (cudaMemcpyAsync into pinned memory): this rank alone (rank 0 while the
    others idle) and all ranks at once. GB/s per GPU."""
    torch = h.torch
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)

    def one():
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        return nbytes * reps / (time.perf_counter() - t0) * 1e-9
    one()
    h.barrier()
    alone = one() if h.rank == 0 else 0.0
    h.barrier()
    conc = one()
    h.barrier()
    alone_max, = h.allmax([alone]); conc_sum, = h.allsum([conc]); conc_min = -h.allmax([-conc])[0]
    del src, dst
    return {"link_gbs_alone": alone_max, "link_gbs_concurrent_per_gpu_mean": conc_sum / h.world, "link_gbs_concurrent_per_gpu_min": conc_min,
            "link_gbs_concurrent_aggregate": conc_sum, "bytes_per_copy": nbytes}


def extra_workloads(h, peaks, K, W):
    """BASELINE configs C3, C4, C5 on bounded batches; every rank runs its shard, values are whole-job aggregates."""
    sg = h.sg; world = h.world
    out = {}
    hbm, hbm_src = hbm_peak()
    peak_gv, _ = sampler_peak_gverts(peaks, world)
    tmp = tempfile.mkdtemp(prefix=f"sgp_bench_r{h.rank}_")
    steps = max(1, min(K, 2)); warm = max(W, 3)          # (the first two runs of a configuration size and allocate the large buffers)

    def general(name, app, argv, n, ne, e_steps, cpu_n, note):
        d = h.timed_device(lambda i: h.run(app, argv, n, i, keep_on_device=True), steps, warm)
        wall, = h.allmax([d["wall"]])
        tot = h.merge(d["summary"])
        fb_ms, = h.allmax([d["tim"]["find_ms"] + d["tim"]["build_ms"]])
        ew_ms, = h.allmax([d["tim"]["emit_write_ms"]])
        verts = float(tot.event_counts.sum())            # vertices of the accepted trees
        sampled = float(tot.vertices_sampled)            # vertices created by the find pass (all attempts)
        out_bytes = float(np.sum(tot.out_bytes))
        e = h.timed_e2e(lambda i: h.run(app, argv, ne, i, keep_on_device=True), ne, e_steps)
        line = {"workload": note, "value": tot.n_ok / wall, "unit": UNIT, "ms_per_step": 1e3 * wall / steps, "steps": steps, "warmup": warm, "replicates_per_gpu_per_step": n,
                "n_failed": int(tot.n_failed), "phase_ms_per_step": {k: v / steps for k, v in d["tim"].items()}, "gpu_launches": d["launches"],
                "vertices_per_tree": verts / max(tot.n_ok, 1), "attempts_per_tree": tot.attempts_total / max(tot.n_ok, 1), "out_bytes_per_tree": out_bytes / max(tot.n_ok, 1),
                "e2e": e,
                "roofline": {"bound": "issue", "kernel": "k_find_general + k_build_general", "achieved": (sampled + verts) / (fb_ms * 1e-3) * 1e-9, "peak": peak_gv, "unit": "Gvertex/s",
                             "frac": (sampled + verts) / (fb_ms * 1e-3) * 1e-9 / peak_gv, "traffic": None,
                             "note": "peak = issue-rate model of the birth-death vertex (a DTL vertex does at least that work); achieved = vertices created by find + build / their time"},
                "roofline_emit": {"bound": "hbm", "kernel": "k_write_tree + k_write_rows + k_emit_info<write>", "achieved": (out_bytes + verts * 128) / (ew_ms * 1e-3) * 1e-9 / 1.0, "peak": hbm * world, "peak_source": hbm_src, "unit": "GB/s",
                                  "frac": (out_bytes + verts * 128) / (ew_ms * 1e-3) * 1e-9 / (hbm * world), "traffic": None,
                                  "note": "algorithmic bytes = text written once + 32-byte emission records read once (Newick piece and .guest2host/.leafmap row per vertex and view)"}}
        if h.rank == 0 and world == 1 and cpu_n:
            v, s = cpu_single(app, argv, cpu_n)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": s}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "measured at N = 1 only"}
        out[name] = line

    # ---- C3: GuestTreeGen batches on the 100-leaf species tree, 8*10^5 per step (81 GB of text; 10^6 would not leave room for the records next to it)
    for db in ("none", "simple"):
        general(f"C3_db_{db}", "GuestTreeGen", ["-s", "20260103", "-db", db, SPECIES100, "0.3", "0.3", "0.3", "bench"], h.args.c3_replicates, 16384, 6, 120,
                f"C3: GuestTreeGen -db {db} 0.3 0.3 0.3 on the 100-leaf species tree (tests/golden/species100.pruned.tree), -rt 0.5, 8 output files/replicate")

    # ---- C4: 1000-leaf species tree (made by the product itself), three recipient-bias modes, then DomainTreeGen over 100 gene trees
    hb = h.ctx.run("HostTreeGen", ["-s", "1", "-min", "1000", "-max", "1000", "-a", "100000", "1.0", "7.0", "0.05", "sp"], n=64, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    ok = np.nonzero(hb.rep_status == 0)[0]
    host1000 = os.path.join(tmp, "species1000.pruned.tree")
    open(host1000, "wb").write(hb.replicate(1, int(ok[0])))
    hb.free()
    for db in (["-db", "exponential", "-dbr", "1.0"], ["-db", "simple"], ["-db", "none"]):
        general(f"C4_guest_db_{db[1]}", "GuestTreeGen", ["-s", "5"] + db + ["-max", "100000", "-maxper", "100000", host1000, "0.3", "0.3", "0.3", "g"], h.args.c4_replicates, 4096, 4, 2,
                f"C4: GuestTreeGen {' '.join(db)} 0.3 0.3 0.3 on a 1000-leaf species tree (HostTreeGen -min 1000 -max 1000 1.0 7.0 0.05)")
    gd = os.path.join(tmp, "genes"); os.makedirs(gd)
    gt = h.ctx.run("GuestTreeGen", ["-s", "77", "-db", "none", host1000, "0.2", "0.2", "0.1", "g"], n=100, rng=sg.RNG_PHILOX, stream_mask=0b0010)
    for i in range(100):
        open(os.path.join(gd, f"gene{i:03d}.tree"), "wb").write(gt.replicate(1, i))
    gt.free()
    general("C4_domain_100_genes", "DomainTreeGen", ["-s", "9", "-ig", "0.5", "-is", "0.5", "-max", "100000", "-maxper", "100000", host1000, gd, "0.3", "0.3", "0.3", "d"],
            h.args.c4_domain_replicates, 256, 2, 2, "C4: DomainTreeGen -ig 0.5 -is 0.5 0.3 0.3 0.3 over 100 gene trees (GuestTreeGen -db none 0.2 0.2 0.1) of the 1000-leaf species tree")

    # the same forest handed over on the device (sgp_domain_run_on_batch): ingest kernels + K0 tables + the run, no gene-tree files
    gk = h.ctx.run("GuestTreeGen", ["-s", "77", "-db", "none", host1000, "0.2", "0.2", "0.1", os.path.join(tmp, "gene")], n=100, rng=sg.RNG_PHILOX, stream_mask=0b0010, flags=sg.FLAG_KEEP_TREES, keep_on_device=True)
    dargs = ["-s", "9", "-ig", "0.5", "-is", "0.5", "-max", "100000", "-maxper", "100000", host1000, "chained", "0.3", "0.3", "0.3", "d"]
    # one untimed run first, like every other workload: the first run of a configuration sizes its vertex arena with a pilot pass, which
    # for this latency-bound workload costs as much as the run itself
    h.ctx.domain_on_batch(gk, dargs, n=h.args.c4_domain_replicates, rng=sg.RNG_PHILOX, offset=(h.world + h.rank) * h.args.c4_domain_replicates, keep_on_device=True).free()
    h.barrier(); h.torch.cuda.synchronize(); t0 = time.perf_counter()
    cb = h.ctx.domain_on_batch(gk, dargs, n=h.args.c4_domain_replicates, rng=sg.RNG_PHILOX, offset=h.rank * h.args.c4_domain_replicates, keep_on_device=True)
    h.torch.cuda.synchronize(); chained_s = time.perf_counter() - t0
    chained_max, = h.allmax([chained_s]); chained_ok, = h.allsum([int(cb.summary().n_ok)])
    out["C4_domain_100_genes"]["chained"] = {"seconds_per_step": chained_max, "value": chained_ok / chained_max, "unit": UNIT,
                                             "note": "forest = 100 pruned trees of a resident GuestTreeGen batch (sgp_domain_run_on_batch); includes ingest and the host-side epoch tables"}
    cb.free(); gk.free()

    # ---- C5: BranchRelaxer over 199-vertex trees, chained on the device behind the generator (10 draws per C2 tree)
    n_src = h.args.c5_source_trees; draws = 10; n_rel = n_src * draws
    t0 = time.perf_counter()
    src = h.run("HostTreeGen", C2_ARGS, n_src, 70_000, keep_on_device=True, flags=sg.FLAG_KEEP_TREES, stream_mask=0b0010)
    h.torch.cuda.synchronize(); gen_s = time.perf_counter() - t0
    for model, margs in (("lognormal", ["IIDLogNormal", "0", "0.25"]), ("gamma", ["IIDGamma", "2", "0.5"])):
        rargs = ["-s", "31", "chained-C2-pruned-trees"] + margs

        def relax(i, n=n_rel):
            return h.ctx.relax_on_batch(src, rargs, which=sg.STREAM_PRUNED_TREE, n=n, rng=sg.RNG_PHILOX, offset=(i * world + h.rank) * n, keep_on_device=True)
        d = h.timed_device(relax, max(steps, 2), warm)
        ksteps = max(steps, 2)
        wall, = h.allmax([d["wall"]]); ok_sum, br_sum, ob_sum = h.allsum([d["summary"].n_ok, d["summary"].vertices_sampled, d["summary"].out_bytes[0]])
        ap_ms, ew_ms = h.allmax([d["tim"]["find_ms"], d["tim"]["emit_write_ms"]])
        e = h.timed_e2e(lambda i: relax(i, n_src), n_src, 6)
        line = {"workload": f"C5: BranchRelaxer {' '.join(margs)} over the pruned trees (199 vertices) of a resident C2 batch, {draws} draws per tree, chained on the device",
                "value": ok_sum / wall, "unit": "relaxed trees/s", "ms_per_step": 1e3 * wall / ksteps, "steps": ksteps, "warmup": warm, "relaxed_trees_per_gpu_per_step": n_rel,
                "source_trees_per_gpu": n_src, "pipeline_value": (ok_sum / ksteps) / (wall / ksteps + gen_s),
                "pipeline_note": "relaxed trees/s when the time to generate the source C2 batch is charged to one step",
                "phase_ms_per_step": {k: v / ksteps for k, v in d["tim"].items()}, "gpu_launches": d["launches"], "branches_per_step": br_sum / ksteps,
                "e2e": e,
                "roofline": {"bound": "hbm", "kernel": "k_relax_apply (rates + lengths)", "achieved": br_sum * 24.0 / (ap_ms * 1e-3) * 1e-9, "peak": hbm * world, "peak_source": hbm_src, "unit": "GB/s",
                             "frac": br_sum * 24.0 / (ap_ms * 1e-3) * 1e-9 / (hbm * world), "traffic": None, "algorithmic_bytes_per_branch": 24,
                             "note": "8 B original length read + 8 B rate + 8 B relaxed length written per branch; the gamma quantile (AS 91) is FP64-bound, not HBM-bound"},
                "roofline_emit": {"bound": "hbm", "kernel": "k_relax_emit<write>", "achieved": (ob_sum + br_sum * 16.0) / (ew_ms * 1e-3) * 1e-9, "peak": hbm * world, "unit": "GB/s",
                                  "frac": (ob_sum + br_sum * 16.0) / (ew_ms * 1e-3) * 1e-9 / (hbm * world), "traffic": None}}
        if h.rank == 0 and world == 1:
            # the oracle relaxes one emitted 199-vertex tree per run, as a reference launch would
            tree = src.replicate(1, int(np.nonzero(src.rep_status == 0)[0][0])).decode().strip()
            v, s = cpu_single("BranchRelaxer", ["-s", "31", tree] + margs, 400)
            line["cpu_baseline"] = {"value": v, "unit": "relaxed trees/s", "cores": 1, "kind": "port", "sample": s}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "relaxed trees/s", "cores": 1, "kind": "port", "sample": "measured at N = 1 only"}
        out[f"C5_{model}"] = line
    src.free()
    return out


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version line to stdout when NCCL_DEBUG is
    set in the environment), so file descriptor 1 is pointed at stderr for the life of the process and the line goes to a private
    duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n"); out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--replicates", type=int, default=1_000_000, help="replicates per GPU per step (C2 headline)")
    ap.add_argument("--e2e-replicates", type=int, default=131072)
    ap.add_argument("--cpu-sample", type=int, default=600, help="replicates timed on one host core for cpu_baseline")
    ap.add_argument("--workloads", default="all", help="'all' (C2 headline + C3/C4/C5 in `workloads`) or 'c2' (headline only)")
    ap.add_argument("--c3-replicates", type=int, default=800_000)
    ap.add_argument("--c4-replicates", type=int, default=32768)
    ap.add_argument("--c4-domain-replicates", type=int, default=4096)
    ap.add_argument("--c5-source-trees", type=int, default=200_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    h = Harness(args)
    sg, torch, world, rank = h.sg, h.torch, h.world, h.rank
    n = args.replicates
    W = max(args.warmup, 3); K = args.steps

    # roofline denominators measured live (FP64 / INT issue rates)
    peaks = h.ctx.issue_peaks()
    for i in range(W):
        b = h.run("HostTreeGen", C2_ARGS, n, 1000 + i, keep_on_device=True); h.merge(b.summary()); b.free()

    stop = threading.Event(); rows = []
    th = threading.Thread(target=clocks_sampler, args=(stop, rows, h.local), daemon=True); th.start()
    h.barrier()
    t0 = time.perf_counter()
    tim = {"find_ms": 0.0, "build_ms": 0.0, "label_ms": 0.0, "emit_size_ms": 0.0, "emit_write_ms": 0.0, "total_ms": 0.0}
    launches = 0; total = None
    for i in range(K):
        b = h.run("HostTreeGen", C2_ARGS, n, i, keep_on_device=True)
        t = b.timings()
        for k in tim:
            tim[k] += t[k]
        launches += int(t["kernel_launches"])
        s = h.merge(b.summary())
        total = s if total is None else sg.Summary.from_vector(total.as_vector() + s.as_vector())
        b.free()
    h.barrier()
    wall = time.perf_counter() - t0
    stop.set(); th.join(timeout=2)
    # max over ranks of the timed region
    wall_max, dev_ms, find_ms, emitw_ms = h.allmax([wall, tim["total_ms"], tim["find_ms"], tim["emit_write_ms"]])
    trees = int(total.n_ok)
    value = trees / wall_max

    ne = args.e2e_replicates
    e_steps = max(12, min(4 * K, 32))
    e2e = h.timed_e2e(lambda i: h.run("HostTreeGen", C2_ARGS, ne, i, keep_on_device=True), ne, e_steps)
    link = pcie_probe(h)
    e2e.update({"host_numa_node": h.numa_node, "host_numa_binding": h.numa_reason, **link,
                "frac_of_link": e2e["d2h_gbs_per_gpu"] / max(link["link_gbs_concurrent_per_gpu_mean"], 1e-9)})

    workloads = extra_workloads(h, peaks, K, W) if args.workloads == "all" else {}

    if rank == 0:
        hbm, hbm_src = hbm_peak()
        traffic = ncu_traffic()
        # algorithmic vertices = what the sequential retry loop expands: (attempts up to and including the accepted one) x
        # mean vertices per attempt; attempts are i.i.d., so the mean over all completed attempts is unbiased
        useful_attempts = float(total.attempts_total) - trees
        verts_per_attempt = float(total.vertices_sampled) / max(float(total.attempts_completed), 1.0)
        verts = useful_attempts * verts_per_attempt
        verts_executed = float(total.vertices_sampled + total.vertices_abandoned)
        gverts = verts / (find_ms * 1e-3) * 1e-9
        peak_gverts, r_alu = sampler_peak_gverts(peaks, world)
        # plain issue-slot bound: every algorithmic warp instruction takes one issue slot of one of the 4 schedulers per SM
        sm_count = torch.cuda.get_device_properties(h.local).multi_processor_count
        clk = summarise_clocks(rows)
        issue_slots = sm_count * 4 * (clk["sm_mhz"] or 1965.0) * 1e6
        alg_inst = ALG_FP64_INST_PER_VERTEX + ALG_IMAD_INST_PER_VERTEX + ALG_ALU_INST_PER_VERTEX
        peak_plain = issue_slots * 32.0 / alg_inst * 1e-9 * world
        out_bytes = float(np.sum(total.out_bytes))
        emit_alg_bytes = out_bytes + 2.0 * float(total.event_counts.sum()) * 32      # text written once + one 32-byte emission record per Newick piece (two views) read once
        emit_gbs = emit_alg_bytes / (emitw_ms * 1e-3) * 1e-9
        cpu_val, cpu_sample = None, "measured at N = 1 only"
        if world == 1:
            cpu_val, cpu_sample = cpu_single("HostTreeGen", C2_ARGS, args.cpu_sample)
        tr_find = traffic["kernels"].get("find") if traffic else None
        tr_emit = traffic["kernels"].get("emit_write") if traffic else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * wall_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "replicates_per_gpu_per_step": n, "timing": "wall clock between barrier+cuda synchronize, max over ranks; "
                       "device CUDA-event time in device_ms_per_step", "l2": "inputs are generated on the fly; outputs (>=20 GB/step) exceed L2",
                       "rng": "philox4x32-10"},
            "device_ms_per_step": dev_ms / K,
            "phase_ms_per_step": {k: v / K for k, v in tim.items()},
            "gpu_launches": launches,
            "attempts_per_tree": float(total.attempts_total) / max(trees, 1), "vertices_per_tree": verts / max(trees, 1),
            "vertices_executed_per_tree": verts_executed / max(trees, 1),
            "out_bytes_per_tree": out_bytes / max(trees, 1), "n_failed": int(total.n_failed),
            "e2e": e2e,
            "roofline": {"bound": "issue", "kernel": "k_bd_find", "achieved": gverts, "peak": peak_gverts, "unit": "Gvertex/s", "frac": gverts / peak_gverts,
                         "peak_plain_issue": peak_plain, "frac_plain_issue": gverts / peak_plain,
                         "traffic": tr_find["dram_bytes_per_replicate"] * n if tr_find else None,
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, scaled per replicate)", "traffic_source": traffic["source"] if traffic else None,
                         "alg_inst_per_vertex": {"fp64": ALG_FP64_INST_PER_VERTEX, "imad_wide": ALG_IMAD_INST_PER_VERTEX, "alu": ALG_ALU_INST_PER_VERTEX},
                         "measured_issue_peaks_ginst_per_s": dict(peaks, alu_derived_ginst_per_s=r_alu), "share_of_step": find_ms / dev_ms},
            "roofline_emit": {"bound": "hbm", "kernel": "k_write_tree<unpruned> + k_write_tree<pruned> + k_emit_info<write>", "achieved": emit_gbs, "peak": hbm * world, "peak_source": hbm_src, "unit": "GB/s",
                              "frac": emit_gbs / (hbm * world), "traffic": tr_emit["dram_bytes_per_replicate"] * n if tr_emit else None, "algorithmic_bytes": emit_alg_bytes / K / world,
                              "traffic_unit": "bytes per step over the write launches (ncu, scaled per replicate)", "traffic_source": traffic["source"] if traffic else None,
                              "share_of_step": emitw_ms / dev_ms},
            "post_sampling_dram_bytes_per_replicate": ({k: (v["dram_bytes_per_replicate"]) for k, v in traffic["kernels"].items() if k != "find"} if traffic else None),
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port", "sample": cpu_sample},
            "clocks": clk,
            "workloads": workloads,
        }
        emit_line(line)
    if world > 1:
        h.dist.destroy_process_group()


if __name__ == "__main__":
    main()
