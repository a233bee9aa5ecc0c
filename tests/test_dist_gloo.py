"""Host-side multi-rank logic (mufasa_b200/dist.py) on CPU: world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mufasa_b200 import dist as D


def test_partition_rows_covers_everything():
    for ny in (1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            edges = [D.partition_rows(ny, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == ny
            for (a0, a1), (b0, b1) in zip(edges[:-1], edges[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_block_cyclic_partition_covers_everything_evenly():
    for ny in (1, 7, 64, 1024, 1027):
        for world in (1, 2, 3, 8):
            rows = [D.partition_row_blocks(ny, world, r) for r in range(world)]
            allr = np.sort(np.concatenate(rows))
            assert np.array_equal(allr, np.arange(ny))
            for r in rows:
                assert np.all(np.diff(r) > 0)                       # ascending: local order == global order
            if ny % (D.ROW_BLOCK * world) == 0:
                assert len({len(r) for r in rows}) == 1
    assert np.array_equal(D.owned_rows(10, 1, 0, "cyclic"), np.arange(10))
    assert np.array_equal(D.owned_rows(20, 2, 1, "slabs"), np.arange(10, 20))
    cube = np.arange(2 * 20 * 3).reshape(2, 20, 3)
    rows = D.partition_row_blocks(20, 2, 1)
    assert np.array_equal(D.take_rows(cube, rows), cube[:, rows, :]) and D.take_rows(cube, rows).flags["C_CONTIGUOUS"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ny, nx, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y0, y1 = D.partition_rows(ny, world, rank)
        full = torch.arange(3 * ny * nx, dtype=torch.float64).reshape(3, ny, nx)
        got = D.gather_maps(full[:, y0:y1, :].contiguous(), ny)
        if rank == 0:
            assert torch.equal(got, full)
        else:
            assert got is None
        # even slabs take the permute path (no per-rank copies); leading dims are kept
        ne = 4 * world
        full_e = torch.arange(2 * 3 * ne * nx, dtype=torch.float64).reshape(2, 3, ne, nx)
        e0, e1 = D.partition_rows(ne, world, rank)
        got = D.gather_maps(full_e[:, :, e0:e1, :].contiguous(), ne)
        assert (rank == 0 and torch.equal(got, full_e)) or (rank != 0 and got is None)
        # block-cyclic rows (blocks of ROW_BLOCK rows dealt round-robin), ragged last block
        nc = 2 * D.ROW_BLOCK * world + 3
        full_c = torch.arange(3 * nc * nx, dtype=torch.float64).reshape(3, nc, nx)
        mine = torch.from_numpy(D.owned_rows(nc, world, rank, "cyclic"))
        got = D.gather_rows(full_c[:, mine, :].contiguous(), nc, partition="cyclic")
        assert (rank == 0 and torch.equal(got, full_c)) or (rank != 0 and got is None)
        # global arg-max of the mask sizes: largest count, ties -> lowest GLOBAL index
        counts = torch.tensor([[5, 40], [7, 3]])            # rank 0: (count 5 at pixel 40); rank 1: (7 at 3)
        c, i, owner = D.global_best(counts[rank].clone())
        assert (c, i, owner) == (7, 3, 1)
        tie = torch.tensor([[9, 100], [9, 64]])
        c, i, owner = D.global_best(tie[rank].clone())
        assert (c, i, owner) == (9, 64, 1)
        none = torch.tensor([[0, 0], [0, 17]])
        c, i, owner = D.global_best(none[rank].clone())
        assert c == 0
        q.put((rank, "ok"))
    except Exception as e:   # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gather_and_global_best_world2():
    world, ny, nx = 2, 7, 5          # uneven slabs: 4 + 3 rows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ny, nx, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_passthrough():
    t = torch.zeros(2, 4, 4)
    assert D.gather_maps(t, 4) is t
    assert D.global_best(torch.tensor([3, 11])) == (3, 11, 0)


def test_vstats_on_device_equal_numpy():
    """dist.global_vstats (torch sort + numpy's interpolation rules) == multi_v_fit.get_vstats (np.median,
    np.percentile) bit for bit: the velocity window of a partitioned run is the single-GPU run's."""
    from mufasa_b200 import multi_v_fit as mvf
    rng = np.random.default_rng(2)
    for n in (1, 2, 3, 4, 17, 18, 1000, 1001, 65536):
        a = rng.normal(size=(2, n)) * 2
        a[0, : n // 7] = 0
        if n > 3:
            a[1, n // 3] = np.nan
        b = a.copy()
        b[b == 0] = np.nan
        assert tuple(mvf.get_vstats(b)) == tuple(D.global_vstats(a.reshape(2, 1, n), "cpu")), n
    x = torch.from_numpy(rng.normal(size=4097))
    assert D.percentiles_like_numpy(x, [0, 1, 50, 99, 100]) == list(np.percentile(x.numpy(), [0, 1, 50, 99, 100]))
