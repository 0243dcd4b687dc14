"""CPU, world_size 2 over gloo: the shard / pose all-gather logic of pcd_reg_hregnet_b200/dist.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcd_reg_hregnet_b200 import dist as hd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = hd.shard_range(n_pairs, rank, world)
    ids = torch.arange(lo, hi, dtype=torch.float32)
    R = ids[:, None, None] + torch.arange(9, dtype=torch.float32).view(1, 3, 3)     # pose "of pair id"
    t = ids[:, None] * 10 + torch.arange(3, dtype=torch.float32)
    if n_pairs % world == 0 and rank == 0:
        # the model path: the pose kernel left the packed [R | t] rows next to R -- the gather sends those as they are
        R.hrn_pose12 = hd.pack_pose(R, t)
    Rg, tg = hd.gather_poses(R, t, n_pairs=n_pairs)
    q.put((rank, Rg.clone(), tg.clone()))
    dist.barrier()
    dist.destroy_process_group()


def _run_once(n_pairs, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    [p.start() for p in ps]
    try:
        res = [q.get(timeout=120) for _ in range(world)]
        [p.join(60) for p in ps]
    finally:
        for p in ps:
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in ps), [p.exitcode for p in ps]
    return res


def _run(n_pairs, world=2, attempts=3):
    """The rendezvous port is picked by binding port 0 and releasing it, and gloo's store teardown can race between
    the ranks: neither is what this test is about, so a failed launch is retried; the results are never relaxed."""
    for a in range(attempts):
        try:
            return _run_once(n_pairs, world)
        except Exception:                                              # noqa: BLE001 -- queue.Empty, AssertionError
            if a == attempts - 1:
                raise


def test_shard_ranges_cover_everything():
    for n, w in ((256, 8), (32, 2), (7, 2), (5, 4), (1, 2)):
        spans = [hd.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_gather_poses_even_and_ragged():
    for n_pairs in (8, 7):
        for rank, Rg, tg in _run(n_pairs):
            ids = torch.arange(n_pairs, dtype=torch.float32)
            assert torch.equal(Rg, ids[:, None, None] + torch.arange(9, dtype=torch.float32).view(1, 3, 3))
            assert torch.equal(tg, ids[:, None] * 10 + torch.arange(3, dtype=torch.float32))
