"""Host-side sharding / reduction / gather logic on CPU with the gloo backend, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rmt_app_b200 import ensemble


def test_partition_covers_range_without_overlap():
    for B in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            edges = [ensemble.partition(B, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_shard_sweep_slices_every_key():
    B = 10
    sw = {"temperature": np.arange(B, dtype=float), "concentration": np.arange(B*3, dtype=float).reshape(B, 3)}
    parts = [ensemble.shard_sweep(sw, B, 3, r) for r in range(3)]
    np.testing.assert_array_equal(np.concatenate([p[0]["temperature"] for p in parts]), sw["temperature"])
    np.testing.assert_array_equal(np.concatenate([p[0]["concentration"] for p in parts]), sw["concentration"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        obj = rng.uniform(1.0, 2.0, B)
        obj[[3, B - 2]] = 0.25                         # tie on the minimum: smallest index wins
        outlet = rng.normal(size=(8, B))
        lo, hi = ensemble.partition(B, world, rank)
        loc = obj[lo:hi]
        s, mn, am = float(loc.sum()), float(loc.min()), int(lo + loc.argmin())
        gs, gmn, gam = ensemble.reduce_objective(s, mn, am, group=None)
        full = ensemble.all_gather_rows(torch.from_numpy(outlet[:, lo:hi].copy()), B)
        fobj = ensemble.all_gather_rows(torch.from_numpy(loc.copy()), B)
        ok = (abs(gs - obj.sum()) < 1e-9 and gmn == 0.25 and gam == 3 and
              np.array_equal(full.numpy(), outlet) and np.array_equal(fobj.numpy(), obj))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 11), (3, 64)])
def test_reduce_and_gather_under_gloo(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res


def test_single_process_passthrough():
    assert ensemble.world_info() == (0, 1)
    assert ensemble.reduce_objective(3.0, 1.0, 5) == (3.0, 1.0, 5)
    t = torch.arange(6.0).view(2, 3)
    assert ensemble.all_gather_rows(t, 3) is t


def _pack_worker(rank, world, port, B, nrows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(1)
        rows = rng.normal(size=(nrows, B))
        status = rng.integers(0, 3, B).astype(np.int32)
        obj = np.abs(rows[-1])
        first = min(2, B - 1)
        obj[[first, B - 1]] = 0.0                          # tie on the minimum: the smallest global index wins
        rows[-1] = obj
        lay = ensemble.PackLayout(B, world, nrows)
        lo, hi = ensemble.partition(B, world, rank)
        buf = torch.full((lay.length,), 7.0, dtype=torch.float64)          # stale content must not leak
        r, st, tail = lay.views(buf, rank)
        assert r.shape == (nrows, hi - lo) and st.shape == (hi - lo,) and tail.shape == (4,)
        r.copy_(torch.from_numpy(rows[:, lo:hi].copy()))
        st.copy_(torch.from_numpy(status[lo:hi].copy()))
        loc = obj[lo:hi]
        if hi > lo:
            tail.copy_(torch.tensor([loc.sum(), loc.min(), float(lo + loc.argmin()), float((status[lo:hi] != 0).sum())],
                                    dtype=torch.float64))
        else:
            tail.copy_(torch.tensor([0.0, float("inf"), -1.0, 0.0], dtype=torch.float64))
        g = ensemble.all_gather_packed(buf, world)                         # the ONE collective
        assert tuple(g.shape) == (world, lay.length)
        full, fst, tails = lay.unpack(g.numpy())
        s, mn, am, bad = ensemble.fold_tails(tails)
        ok = (np.array_equal(full, rows) and np.array_equal(fst, status) and abs(s - obj.sum()) < 1e-9
              and mn == 0.0 and am == first and bad == int((status != 0).sum()))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B,nrows", [(2, 11, 9), (3, 64, 3), (3, 2, 9)])
def test_packed_single_collective_under_gloo(world, B, nrows):
    """The cross-GPU step of a sharded ensemble: every rank's outlets / objectives / status / reduction tail in ONE
    buffer, ONE all-gather, unpacked on the host — uneven shards, odd shard sizes (int32 status packing) and a rank
    without any reactor (B < world)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pack_worker, args=(r, world, port, B, nrows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res


def test_pack_layout_single_rank_and_tail_folding():
    lay = ensemble.PackLayout(5, 1, 2)
    assert lay.sizes == [5] and lay.length == 2*5 + 3 + 4
    buf = torch.zeros(lay.length, dtype=torch.float64)
    r, st, tail = lay.views(buf, 0)
    r.copy_(torch.arange(10.0).view(2, 5)); st.copy_(torch.tensor([0, 1, 0, 2, 0], dtype=torch.int32))
    tail.copy_(torch.tensor([3.0, 0.5, 4.0, 2.0], dtype=torch.float64))
    full, fst, tails = lay.unpack(ensemble.all_gather_packed(buf, 1).numpy())
    np.testing.assert_array_equal(full, np.arange(10.0).reshape(2, 5))
    np.testing.assert_array_equal(fst, [0, 1, 0, 2, 0])
    assert ensemble.fold_tails(tails) == (3.0, 0.5, 4, 2)
    assert ensemble.fold_tails([[0.0, np.inf, -1.0, 0.0], [2.0, 1.0, 9.0, 1.0], [2.5, 1.0, 7.0, 0.0]]) == (4.5, 1.0, 7, 1)


def test_device_unpack_equals_host_unpack(monkeypatch):
    """PackLayout.unpack_device (scatter + transpose by tensor copies, one D2H into pinned workspace memory) against the
    host-side unpack on the same gathered buffer — equal and uneven shards, a rank without reactors,
    instance-major and row-major.  CPU tensors stand in for the device here; the stream synchronisation is stubbed."""
    import torch
    from rmt_app_b200.ensemble import PackLayout

    class _Ws:
        def get(self, name, shape, dtype, device=None, pinned=False):
            return torch.empty(shape, dtype=dtype)

    class _Stream:
        def synchronize(self):
            pass
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: _Stream())
    rng = np.random.default_rng(0)
    for B, world, nrows in ((64, 4, 3), (66, 4, 3), (35, 5, 2), (40, 8, 9), (3, 4, 2)):
        lay = PackLayout(B, world, nrows)
        g = torch.zeros((world, lay.length), dtype=torch.float64)
        for r in range(world):
            rows, st, tail = lay.views(g[r], r)
            rows.copy_(torch.from_numpy(rng.normal(size=tuple(rows.shape))))
            st.copy_(torch.from_numpy(rng.integers(0, 4, st.shape[0]).astype(np.int32)))
            tail.copy_(torch.tensor([1.0 + r, 2.0, 3.0, float(r)]))
        full, status, tails = lay.unpack(g.numpy())
        a, b, c = lay.unpack_device(g, _Ws(), transpose=True)
        assert np.array_equal(a, full.T) and np.array_equal(b, status) and np.array_equal(c, tails)
        a, b, c = lay.unpack_device(g, _Ws(), transpose=False)
        assert np.array_equal(a, full) and np.array_equal(b, status)
