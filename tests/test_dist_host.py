"""Host-side logic of the multi-GPU path on CPU: block-aligned row split, and the descriptor
exchange over a world_size-2 gloo group (the device side needs GPUs: tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import problems as pr
from ccqppy_b200.dist import shard_rows, batch_range


def test_split_elementwise_is_even():
    rows = pr.box_table(32768).rows
    r = shard_rows(rows, 32768, 8)
    assert r == [(i * 4096, (i + 1) * 4096) for i in range(8)]
    assert shard_rows(rows, 32768, 1) == [(0, 32768)]


def test_split_respects_norm_blocks():
    tab = pr.sphere3_table(16384)       # 5461 discs + 1 identity entry
    for world in (2, 3, 4, 8):
        r = shard_rows(tab.rows, tab.n, world)
        assert r[0][0] == 0 and r[-1][1] == tab.n
        for (a0, a1), (b0, b1) in zip(r[:-1], r[1:]):
            assert a1 == b0 and a1 > a0
        for r0, r1 in r[:-1]:
            assert r1 % 3 == 0                      # never cuts a disc
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 4


def test_split_mixed_and_impossible():
    tab = pr.mixed_table(300)
    r = shard_rows(tab.rows, 300, 4)
    for r0, r1 in r:
        for kind, off, dim, _ in tab.blocks:
            if kind == pr.SPHERE and dim > 1:
                assert not (off < r0 < off + dim) and not (off < r1 < off + dim)
    with pytest.raises(ValueError):
        shard_rows(pr.sphere_table(100).rows, 100, 2)     # one whole-vector sphere cannot be split


def test_batch_split_is_contiguous_and_even():
    for batch, world in ((65536, 8), (10, 4), (7, 8), (1, 2), (100, 3)):
        r = [batch_range(batch, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ccqppy_b200.dist import exchange_descriptors
    mine = bytes([rank + 1]) * 128
    alld = exchange_descriptors(mine)
    q.put((rank, alld == bytes([1]) * 128 + bytes([2]) * 128))
    dist.barrier()
    dist.destroy_process_group()


def test_descriptor_exchange_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, True), (1, True)]


def test_split_properties_on_random_tables():
    """Any table / world size: the ranges tile [0, n) in order, and no norm-type block is cut."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(seed=st.integers(0, 2 ** 31 - 1), world=st.integers(1, 8))
    def check(seed, world):
        rng = np.random.default_rng(seed)
        t = pr.Table()
        for _ in range(int(rng.integers(1, 60))):
            kind = int(rng.integers(0, 7))
            dim = int(rng.integers(1, 40))
            if kind == pr.IDENTITY: t.add(kind, dim)
            elif kind in (pr.LOWER, pr.UPPER): t.add(kind, dim, 0.0)
            elif kind == pr.BOX: t.add(kind, dim, -1.0, 1.0)
            else: t.add(kind, dim, 1.0)
        try:
            r = shard_rows(t.rows, t.n, world)
        except ValueError:
            return          # fewer cut points than ranks: refused, not mis-split
        assert r[0][0] == 0 and r[-1][1] == t.n and len(r) == world
        assert all(a[1] == b[0] and a[1] > a[0] for a, b in zip(r[:-1], r[1:])) and r[-1][1] > r[-1][0]
        cuts = {a[1] for a in r[:-1]}
        for kind, off, dim, _ in t.rows:
            if kind in (pr.SPHERE, pr.CONE_REF, pr.SOC) and dim > 1:
                assert not any(off < c < off + dim for c in cuts)
    check()
