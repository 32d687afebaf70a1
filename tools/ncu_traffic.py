#!/usr/bin/env python3
"""DRAM traffic per replicate of the C2 kernels, from an `ncu --set full` capture exported with `--page raw --csv`.

    python tools/ncu_traffic.py <raw.csv> <replicates in the profiled run> <out.json>

bench.py reads the JSON (profiles/r02_traffic.json) and scales it to the replicates of a step for `roofline.traffic`.
Kernel groups: find = k_bd_find*, build = k_bd_build, label = k_label_prune* + k_info_sums, finish = k_finish_records,
emit_write = k_write_tree* + k_write_rows* + k_emit_info<true>. Launches of one profiled pass are summed per group."""
import csv, json, sys

src, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name"); ri = hdr.index("dram__bytes_read.sum"); wi = hdr.index("dram__bytes_write.sum"); ti = hdr.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tscale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}
groups = {"find": ("k_bd_find",), "build": ("k_bd_build",), "label": ("k_label_prune", "k_info_sums"), "finish": ("k_finish_records",),
          "emit_write": ("k_write_tree", "k_write_rows", "k_emit_info<(bool)1>", "k_emit_info<true>", "k_emit_info<1>")}
acc = {g: {"dram_bytes": 0.0, "ms": 0.0, "launches": 0} for g in groups}
passes = {}
for r in rows[2:]:
    if len(r) <= max(ki, ri, wi, ti):
        continue
    name = r[ki]
    for g, pats in groups.items():
        if any(p in name for p in pats):
            def num(x):
                try:
                    v = float(x.replace(",", ""))
                    return v if v == v else None
                except ValueError:
                    return None
            rd, wr = num(r[ri]), num(r[wi])
            b = (rd * scale[units[ri]] + wr * scale[units[wi]]) if rd is not None and wr is not None else float("nan")
            acc[g]["dram_bytes"] += b; acc[g]["ms"] += float(r[ti].replace(",", "")) * tscale[units[ti]]; acc[g]["launches"] += 1
            passes[g] = passes.get(g, 0) + (1 if any(p in name for p in pats[:1]) else 0)
res = {"source": src, "replicates_profiled": n, "kernels": {}}
# the capture may hold several passes of the pipeline (profile_run.py runs two): normalise by the number of find/build launches seen
n_pass = max(1, acc["build"]["launches"] or acc["find"]["launches"] or 1)
for g, a in acc.items():
    if a["launches"]:
        res["kernels"][g] = {"dram_bytes_per_replicate": (a["dram_bytes"] / n_pass / n) if a["dram_bytes"] == a["dram_bytes"] else None, "ms_per_replicate_under_ncu": a["ms"] / n_pass / n, "launches_per_pass": a["launches"] / n_pass}
res["passes_in_capture"] = n_pass
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
