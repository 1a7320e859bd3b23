"""N>1 plumbing on CPU: world_size-2 gloo group, the all-gather of the ragged peak table and the rank segments."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from r4w_b200.dist import all_gather_table, all_reduce_power, segment_for_rank
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 7 * 20000 + 123                       # ragged: 7 snapshots over 2 ranks + a tail
        first, n = segment_for_rank(total, 20000, rank, world)
        s0, ns = first // 20000, n // 20000
        # fake peak table whose content encodes the global snapshot index, so order after the gather is checkable
        local = np.zeros((ns, 3, 6))
        for s in range(ns):
            local[s, :, 0] = [3, 25, 8]
            local[s, :, 2] = (s0 + s) * 100 + np.arange(3)
        full = all_gather_table(local)
        psum, cnt = all_reduce_power(1000.0 * (rank + 1), n)          # avg-power line of a time-sharded render
        q.put((rank, first, n, full, psum, cnt))
    finally:
        dist.destroy_process_group()


def test_all_gather_peak_table_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, n0, t0, p0, c0), (r1, f1, n1, t1, p1, c1) = out
    assert p0 == p1 == 3000.0 and c0 == c1 == 7 * 20000 + 123
    assert f0 == 0 and f0 + n0 == f1 and f1 + n1 == 7 * 20000 + 123
    assert n0 == 4 * 20000 and n1 == 3 * 20000 + 123
    assert t0.shape == (7, 3, 6) and np.array_equal(t0, t1)
    assert np.array_equal(t0[:, 0, 2], np.arange(7) * 100.0) and np.array_equal(t0[:, 2, 2], np.arange(7) * 100.0 + 2)


def test_all_gather_single_process_is_identity():
    from r4w_b200.dist import all_gather_table
    a = np.arange(24, dtype=np.float64).reshape(2, 2, 6)
    assert np.array_equal(all_gather_table(a), a)
    from r4w_b200.dist import all_reduce_power
    assert all_reduce_power(12.5, 100) == (12.5, 100)


def test_segments_partition_the_file():
    """time sharding (bench.py, weak and --strong): the ranks' segments tile [0, total) exactly, every inner edge is a multiple
    of the alignment (a 4 ms snapshot), sizes differ by at most one unit, the ragged tail goes to the last rank"""
    from r4w_b200.dist import segment_for_rank, snapshots_for_rank
    for total, align in ((3_000_000_000, 20000), (100_000_000, 20000), (35_500, 5000), (19_999, 20000), (0, 5000)):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            sizes = []
            for r in range(world):
                first, n = segment_for_rank(total, align, r, world)
                assert first == pos and n >= 0
                if r < world - 1:
                    assert (first + n) % align == 0
                pos += n
                sizes.append(n)
            assert pos == total
            assert max(sizes[:-1] or [0]) - min(sizes[:-1] or [0]) <= align
    for n_snap, world in ((150_000, 8), (5000, 3), (7, 8), (0, 2)):
        got = [snapshots_for_rank(n_snap, r, world) for r in range(world)]
        assert got[0][0] == 0 and sum(c for _, c in got) == n_snap
        assert all(got[r][0] + got[r][1] == got[r + 1][0] for r in range(world - 1))
    import pytest
    with pytest.raises(ValueError):
        segment_for_rank(100, 10, 2, 2)
