"""Breakdown of the end-to-end pipeline: compute wall time per chunk, copy time alone, pipelined."""
import sys, time, torch
sys.path.insert(0, ".")
import sagephy_b200 as sg
from bench import C2_ARGS
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
ctx = sg.Context(0)
def step(i): return ctx.run("HostTreeGen", C2_ARGS, n=ne, rng=sg.RNG_PHILOX, offset=i * ne, keep_on_device=True)
b = step(0); sizes = [b.stream_bytes(s) for s in range(8)]; b.free()
pinned = [[torch.empty(int(sz * 1.05) + 4096, dtype=torch.uint8, pin_memory=True) if sz else None for sz in sizes] for _ in range(2)]
offs = [[torch.empty(ne + 1, dtype=torch.int64, pin_memory=True) for _ in range(8)] for _ in range(2)]
for i in range(2): step(1 + i).free()
t = time.perf_counter(); bs = []
for i in range(4):
    b = step(10 + i); print("  timings", {k: round(v, 1) for k, v in b.timings().items()}); b.free()
print("compute wall ms/chunk", (time.perf_counter() - t) / 4 * 1e3)
b = step(20)
for rep in range(3):
    t = time.perf_counter(); tot = 0
    for s in range(8):
        if pinned[0][s] is not None and b.stream_bytes(s):
            tot += b.copy_stream_into_async(s, pinned[0][s].data_ptr(), pinned[0][s].numel()); b.copy_offsets_into_async(s, offs[0][s].data_ptr())
    ctx.wait_copies(); dt = time.perf_counter() - t
    print("copy alone ms", dt * 1e3, "GB/s", tot / dt / 1e9)
t = time.perf_counter(); st = b.rep_status; print("rep_status ms", (time.perf_counter() - t) * 1e3)
t = time.perf_counter(); b.free(); print("free ms", (time.perf_counter() - t) * 1e3)

# the pipelined loop of bench.py's e2e leg with host timers around every phase (where do the bubbles on the copy engine come from?)
import collections
ph = collections.defaultdict(float); prev = None; n_steps = 10
torch.cuda.synchronize(); T0 = time.perf_counter()
for i in range(n_steps):
    t = time.perf_counter(); b = step(30 + i); ph["compute (blocking)"] += time.perf_counter() - t
    for k, v in b.timings().items():
        if k.endswith("_ms"): ph["dev " + k] += v * 1e-3
    if prev is not None:
        t = time.perf_counter(); ctx.wait_copies(); ph["wait_copies"] += time.perf_counter() - t
        t = time.perf_counter(); prev.free(); ph["free"] += time.perf_counter() - t
    t = time.perf_counter(); st = b.rep_status; ph["rep_status"] += time.perf_counter() - t
    t = time.perf_counter(); tot = 0
    for s in range(8):
        if pinned[i & 1][s] is not None and b.stream_bytes(s):
            tot += b.copy_stream_into_async(s, pinned[i & 1][s].data_ptr(), pinned[i & 1][s].numel()); b.copy_offsets_into_async(s, offs[i & 1][s].data_ptr())
    ph["enqueue copies"] += time.perf_counter() - t
    prev = b
ctx.wait_copies(); prev.free()
wall = time.perf_counter() - T0
print("pipelined ms/chunk", wall / n_steps * 1e3, "GB/s", tot * n_steps / wall / 1e9, {k: round(v / n_steps * 1e3, 2) for k, v in ph.items()})
