"""CPU tests of the multi-GPU host logic: index-range sharding and the gloo all-reduce of summary vectors (world size 2)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from sagephy_b200.sharding import allreduce_summary, shard_range


def test_shard_ranges_partition_the_replicates():
    for n in (0, 1, 7, 1000, 10 ** 6 + 3):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in parts) == n
            assert parts[0][0] == 0
            for (o0, c0), (o1, _) in zip(parts, parts[1:]):
                assert o0 + c0 == o1
            counts = [c for _, c in parts]
            assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sagephy_b200 as sg
    off, cnt = shard_range(1001, rank, world)
    # a rank-local summary as the library would return it for its shard (n_ok = replicates of the shard, ...)
    ev = np.arange(17, dtype=np.int64) * (rank + 1); hist = np.zeros(1025, dtype=np.int64); hist[100] = cnt
    s = sg.Summary(cnt, rank, 276 * cnt, 45000 * cnt, 5000 * cnt, 300 * cnt, ev, hist, np.full(8, 10 * cnt, dtype=np.int64))
    merged = sg.Summary.from_vector(allreduce_summary(s.as_vector()))
    q.put((rank, off, cnt, merged.n_ok, merged.n_failed, int(merged.leaf_hist[100]), int(merged.event_counts[3]), int(merged.out_bytes.sum())))
    dist.destroy_process_group()


def test_gloo_allreduce_of_summaries_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, o0, c0, *m0), (r1, o1, c1, *m1) = res
    assert (o0, c0, o1, c1) == (0, 501, 501, 500)
    assert m0 == m1 == [1001, 1, 1001, 3 * (1 + 2), 80 * 1001]
