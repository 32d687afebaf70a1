import sys; sys.path.insert(0, "/root/repo")
import sagephy_b200 as sg
from bench import C2_ARGS
ctx = sg.Context(0)
for rep in range(2):
    b = ctx.run("HostTreeGen", C2_ARGS, n=1000000, rng=sg.RNG_PHILOX, keep_on_device=True, stream_mask=2)
    s = b.summary(); t = b.timings()
    print("find_ms", round(t["find_ms"], 1), "sampled", s.vertices_sampled, "abandoned", s.vertices_abandoned, "completed", s.attempts_completed, "attempts", s.attempts_total, "ratio", round(s.vertices_abandoned / (s.vertices_sampled + s.vertices_abandoned), 4))
    b.free()
