#!/usr/bin/env python3
"""Benchmark of the accelerated SaGePhy hot path (see BASELINE.json / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                       (CPU reference arm: the reference algorithm on host cores)

One "step" = one pass of the whole path (sample -> label/prune -> emit, outputs resident in HBM) over one batch of
synthetic replicates of BASELINE config C2: HostTreeGen, T=1, birth 5, death 0.05, 100-leaf window, leaf sampling 0.8.
Replicates are sharded by index range across ranks (no data-path collective); the only collective is the NCCL
all-reduce(sum) of the summary histograms.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C2: HostTreeGen -min 100 -max 100 -p 0.8 -a 10000 1.0 5.0 0.05, Philox throughput mode, 4 output files/replicate"
C2_ARGS = ["-s", "20260101", "-min", "100", "-max", "100", "-p", "0.8", "-a", "10000", "1.0", "5.0", "0.05", "bench"]
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
# DRAM traffic per replicate from the `ncu --set full` capture profiles/r01_c2_full_metrics.csv (200 000 replicates per launch):
# dram__bytes_read.sum + dram__bytes_write.sum of k_bd_find_single, and of the four k_emit<write> launches together.
NCU_FIND_DRAM_BYTES_PER_REPLICATE = (0.008465 + 0.092016) * 1e9 / 200000
NCU_EMIT_WRITE_DRAM_BYTES_PER_REPLICATE = (5.014102 + 2.810351 + 4.534508 + 2.139986 + 0.050887 + 0.015014 + 0.054069 + 0.010760) * 1e9 / 200000


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
    end-to-end pipeline are allocated (first touch) next to the GPU's PCIe root port. Best effort; returns the node or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
        if bus is None:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)], capture_output=True, text=True, timeout=5).stdout.strip()
            bus = out
        bus = str(bus).lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]                      # nvidia-smi prints an 8-digit domain, sysfs uses 4
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


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


def cpu_reference(n_trees, procs):
    """Times the reference algorithm (CPU oracle, C++ port of the Java code path) on `procs` host processes."""
    from concurrent.futures import ProcessPoolExecutor
    per = max(1, n_trees // procs)
    t0 = time.perf_counter()
    with ProcessPoolExecutor(procs) as ex:
        res = list(ex.map(_cpu_worker, [(i * per, per) for i in range(procs)]))
    wall = time.perf_counter() - t0
    ok = sum(r[0] for r in res)
    return ok / wall, ok, wall, sum(r[1] for r in res)


def _cpu_worker(job):
    first, n = job
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_util import oracle_batch
    sec, st = oracle_batch("HostTreeGen", C2_ARGS, first, n)
    return int((st[:, 1] == 0).sum()), float(st[:, 7].sum())


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, cores)
    per_step = max(procs * 64, 256)         # ~1-2 s of work on all host cores per step
    for _ in range(args.warmup):
        cpu_reference(max(procs, per_step // 8), procs)
    t0 = time.perf_counter(); ok = 0
    for _ in range(args.steps):
        v, k, wall, verts = cpu_reference(per_step, procs)
        ok += k
    wall = time.perf_counter() - t0
    value = ok / wall
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "replicates_per_step": per_step, "note": "reference-algorithm CPU restatement (C++ oracle, MT19937 stream, one process per host core); the Java jar cannot run here (no JVM)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": f"{per_step} replicates per step x {args.steps} steps of the same workload"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--replicates", type=int, default=1_000_000, help="replicates per GPU per step")
    ap.add_argument("--e2e-replicates", type=int, default=131072)
    ap.add_argument("--cpu-sample", type=int, default=600, help="replicates timed on one host core for cpu_baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import sagephy_b200 as sg
    from sagephy_b200.sharding import allreduce_summary

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sg.Context(local)
    n = args.replicates
    W = max(args.warmup, 3); K = args.steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, n_rep, keep=True):
        # global replicate ids: step-major, then rank-major index ranges
        off = (i * world + rank) * n_rep
        return ctx.run("HostTreeGen", C2_ARGS, n=n_rep, rng=sg.RNG_PHILOX, offset=off, keep_on_device=keep)

    def merge(summary):
        # NCCL all-reduce(sum) of the int64 summary vector: the only collective on the path
        return sg.Summary.from_vector(allreduce_summary(summary.as_vector(), device="cuda"))

    # roofline denominators measured live (FP64 / INT issue rates)
    peaks = ctx.issue_peaks()
    for i in range(W):
        b = step(1000 + i, n); merge(b.summary()); b.free()

    stop = threading.Event(); rows = []
    th = threading.Thread(target=clocks_sampler, args=(stop, rows, local), daemon=True); th.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    tim = {"find_ms": 0.0, "build_ms": 0.0, "label_ms": 0.0, "emit_size_ms": 0.0, "emit_write_ms": 0.0, "total_ms": 0.0}
    launches = 0; total = None; node_bytes = 0
    for i in range(K):
        b = step(i, n)
        t = b.timings()
        for k in tim:
            tim[k] += t[k]
        launches += int(t["kernel_launches"])
        s = merge(b.summary())
        total = s if total is None else sg.Summary.from_vector(total.as_vector() + s.as_vector())
        b.free()
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    stop.set(); th.join(timeout=2)
    # max over ranks of the timed region
    tt = torch.tensor([wall, tim["total_ms"], tim["find_ms"], tim["emit_write_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall_max, dev_ms, find_ms, emitw_ms = [float(x) for x in tt.cpu()]
    trees = int(total.n_ok)
    value = trees / wall_max

    # end-to-end through the public API with host buffers: argv in, every output file's bytes + offsets out to pinned host
    # memory. Double-buffered: the device->host copies of chunk i run on the context's copy stream while chunk i+1 computes.
    ne = args.e2e_replicates
    b = step(2000, ne); sizes = [b.stream_bytes(s) for s in range(8)]; b.free()
    pinned = [[torch.empty(int(sz * 1.05) + 4096, dtype=torch.uint8, pin_memory=True) if sz else None for sz in sizes] for _ in range(2)]
    offs = [[torch.empty(ne + 1, dtype=torch.int64, pin_memory=True) for _ in range(8)] for _ in range(2)]
    e_steps = max(12, min(4 * K, 32))
    barrier(); t0 = time.perf_counter(); e_ok = 0; d2h = 0; prev = None
    for i in range(e_steps):
        b = step(3000 + i, ne)
        if prev is not None:
            ctx.wait_copies(); prev.free()
        # the small status fetch goes first: issued after the bulk copies it would queue behind them on the copy engine
        st = b.rep_status; d2h += st.nbytes * 3
        e_ok += int((st == 0).sum())
        buf, ob = pinned[i & 1], offs[i & 1]
        for s in range(8):
            if buf[s] is not None and b.stream_bytes(s):
                d2h += b.copy_stream_into_async(s, buf[s].data_ptr(), buf[s].numel())
                b.copy_offsets_into_async(s, ob[s].data_ptr()); d2h += 8 * (ne + 1)
        prev = b
    ctx.wait_copies(); prev.free()
    barrier(); e_wall = time.perf_counter() - t0
    ee = torch.tensor([e_wall], dtype=torch.float64, device="cuda"); eo = torch.tensor([e_ok, d2h], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ee, op=dist.ReduceOp.MAX); dist.all_reduce(eo)
    e2e_value = float(eo[0]) / float(ee[0])
    h2d_per_step = 4096          # argv-derived tables (host-tree tables, constant strings, Java math tables are resident per context)

    if rank == 0:
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback"
        if os.path.exists(peaks_file):
            try:
                hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"]); hbm_src = "measured"
            except Exception:
                pass
        # algorithmic vertices = what the sequential retry loop expands: (attempts up to and including the accepted one) x
        # mean vertices per attempt; attempts are i.i.d., so the mean over all completed attempts is unbiased
        useful_attempts = float(total.attempts_total) - trees
        verts_per_attempt = float(total.vertices_sampled) / max(float(total.attempts_completed), 1.0)
        verts = useful_attempts * verts_per_attempt
        verts_executed = float(total.vertices_sampled + total.vertices_abandoned)
        gverts = verts / (find_ms * 1e-3) * 1e-9
        # issue roofline of the sampler: time of one warp-vertex = its FP64, IMAD and other instructions at their measured
        # issue rates (the ALU rate follows from the IMAD+SHF+LOP3 mix probe: 3/R_mixed = 1/R_imad + 2/R_alu)
        r_fp64 = peaks["dfma_ginst_per_s"]; r_imad = peaks["imad_ginst_per_s"]; r_mixed = peaks["mixed_int_ginst_per_s"]
        r_alu = 2.0 / max(3.0 / r_mixed - 1.0 / r_imad, 1e-9)
        t_vertex_warp = ALG_FP64_INST_PER_VERTEX / r_fp64 + ALG_IMAD_INST_PER_VERTEX / r_imad + ALG_ALU_INST_PER_VERTEX / r_alu     # ns per warp of 32 vertices
        peak_gverts = 32.0 / t_vertex_warp * world
        out_bytes = float(np.sum(total.out_bytes))
        emit_alg_bytes = out_bytes + float(total.event_counts.sum()) * (80 + 16)      # node records + order arrays read once, text written once
        emit_gbs = emit_alg_bytes / (emitw_ms * 1e-3) * 1e-9
        cpu_n = args.cpu_sample
        cpu_val, cpu_sample = None, "measured at N = 1 only"
        if world == 1:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle_util import oracle_batch
            sec, st = oracle_batch("HostTreeGen", C2_ARGS, 0, cpu_n)
            cpu_val = float((st[:, 1] == 0).sum()) / sec
            cpu_sample = f"{cpu_n} replicates of the same workload (MT19937 stream), single thread, C++ restatement of the Java path"
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
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": int(float(eo[1]) / e_steps / world),
                    "replicates_per_gpu_per_step": ne, "steps": e_steps, "host_numa_node": numa_node},
            "roofline": {"bound": "issue", "kernel": "k_bd_find", "achieved": gverts, "peak": peak_gverts, "unit": "Gvertex/s", "frac": gverts / peak_gverts,
                         "traffic": NCU_FIND_DRAM_BYTES_PER_REPLICATE * n, "traffic_unit": "bytes per launch (ncu, scaled from a 200k-replicate capture)",
                         "alg_inst_per_vertex": {"fp64": ALG_FP64_INST_PER_VERTEX, "imad_wide": ALG_IMAD_INST_PER_VERTEX, "alu": ALG_ALU_INST_PER_VERTEX},
                         "measured_issue_peaks_ginst_per_s": dict(peaks, alu_derived_ginst_per_s=r_alu), "share_of_step": find_ms / dev_ms},
            "roofline_emit": {"bound": "hbm", "kernel": "k_emit<write>", "achieved": emit_gbs, "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
                              "frac": emit_gbs / hbm_peak, "traffic": NCU_EMIT_WRITE_DRAM_BYTES_PER_REPLICATE * n, "algorithmic_bytes": emit_alg_bytes / K / world,
                              "traffic_unit": "bytes per step over the four write launches (ncu, scaled from a 200k-replicate capture)", "share_of_step": emitw_ms / dev_ms},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port", "sample": cpu_sample},
            "clocks": summarise_clocks(rows),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
