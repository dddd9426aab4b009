"""CPU: the multi-GPU tiler (pdgn_b200.dist).  Partition logic as unit tests, and the all_gather path with
world_size 2 over gloo, the tile computation substituted by the CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("shape", [(1000, 1000), (7, 5), (1, 9), (8, 8), (33, 2)])
def test_tiles_partition_the_pair_grid(world, shape):
    from pdgn_b200 import dist as pd
    n_rows, n_cols = shape
    cover = np.zeros(shape, dtype=np.int32)
    pr, pc = pd.rank_grid(world)
    assert pr * pc == world and pr <= pc
    mr, mc = pd.max_tile(world, n_rows, n_cols)
    for r in range(world):
        (r0, r1), (c0, c1) = pd.tile_of(r, world, n_rows, n_cols)
        assert 0 <= r0 <= r1 <= n_rows and 0 <= c0 <= c1 <= n_cols
        assert r1 - r0 <= mr and c1 - c0 <= mc
        cover[r0:r1, c0:c1] += 1
    assert np.all(cover == 1)


def test_rank_grid_named_configs():
    from pdgn_b200 import dist as pd
    assert [pd.rank_grid(w) for w in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    (r0, r1), (c0, c1) = pd.tile_of(5, 8, 1000, 1000)
    assert (r1 - r0, c1 - c0) == (500, 250)


def _worker(rank, world, port, na, nb, npts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cpu as ocpu
        from pdgn_b200 import dist as pd
        rng = np.random.default_rng(0)
        A = torch.from_numpy(rng.uniform(-1, 1, (na, npts, 3)).astype(np.float32))
        B = torch.from_numpy(rng.uniform(-1, 1, (nb, npts, 3)).astype(np.float32))

        def cpu_tile(a, b, rows, cols):
            return torch.from_numpy(ocpu.cd_allpairs(a[rows[0]:rows[1]].numpy(), b[cols[0]:cols[1]].numpy()))

        assert pd.is_distributed()
        full = pd.pairwise_cd(A, B, compute_tile=cpu_tile)
        ref = torch.from_numpy(ocpu.cd_allpairs(A.numpy(), B.numpy()))
        q.put((rank, bool(torch.equal(full, ref)), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("na,nb", [(5, 7), (4, 4), (1, 3)])
def test_all_gather_path_world2_gloo(na, nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + na * 13 + nb) % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, na, nb, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok and shape == (na, nb) for _, ok, shape in res)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [1000, 7, 1, 64, 33])
def test_symmetric_plan_covers_every_unordered_block_pair_once_and_is_balanced(world, n):
    from pdgn_b200 import dist as pd
    bounds, owner = pd.sym_plan(world, n)
    t = len(bounds)
    assert bounds[0][0] == 0 and bounds[-1][1] == n and all(bounds[i][1] == bounds[i + 1][0] for i in range(t - 1))
    seen = sorted(ij for tiles in owner for ij in tiles)
    assert seen == [(i, j) for i in range(t) for j in range(i, t)]
    cover = np.zeros((n, n), dtype=np.int32)
    for tiles in owner:
        for i, j in tiles:
            (r0, r1), (c0, c1) = bounds[i], bounds[j]
            cover[r0:r1, c0:c1] += 1
            if i != j:
                cover[c0:c1, r0:r1] += 1
    assert np.all(cover == 1)
    if n >= 64 * world:  # enough blocks: no rank carries more than ~15 % over the mean
        size = [hi - lo for lo, hi in bounds]
        load = [sum(size[i] * size[j] * (0.5 if i == j else 1.0) for i, j in tiles) for tiles in owner]
        assert max(load) <= 1.15 * (sum(load) / world)


def _worker_sym(rank, world, port, n, npts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cpu as ocpu
        from pdgn_b200 import dist as pd
        rng = np.random.default_rng(3)
        A = torch.from_numpy(rng.uniform(-1, 1, (n, npts, 3)).astype(np.float32))
        calls = []

        def cpu_tile(a, b, rows, cols):
            calls.append((rows, cols))
            return torch.from_numpy(ocpu.cd_allpairs(a[rows[0]:rows[1]].numpy(), b[cols[0]:cols[1]].numpy()))

        full = pd.pairwise_cd_symmetric(A, compute_tile=cpu_tile)
        ref = torch.from_numpy(ocpu.cd_allpairs(A.numpy(), A.numpy()))
        pairs = sum((r[1] - r[0]) * (c[1] - c[0]) for r, c in calls)
        q.put((rank, bool(torch.equal(full, ref)), tuple(full.shape), pairs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 16, 2])
def test_symmetric_all_gather_path_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n * 31 + 7) % 400
    procs = [ctx.Process(target=_worker_sym, args=(r, 2, port, n, 48, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and shape == (n, n) for _, ok, shape, _ in res)
    # together the ranks evaluated the upper triangle (diagonal blocks in full): fewer pairs than the n*n of plain tiling
    assert sum(r[3] for r in res) < n * n or n <= 2
