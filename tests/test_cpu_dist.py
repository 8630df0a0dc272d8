"""CPU, world_size 2, gloo: the N>1 host logic — tile assignment and the tower merge (all-gather of
tower records + the reference's 30 m greedy duplicate rule in (rank, detection) order)."""
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
    from pointcloudhookup_b200 import dist as pdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    def tower(x, y, label):
        return {"center": np.array([x, y, 100.0]), "extent": np.array([10.0, 12.0, 30.0]), "rotation": np.eye(3),
                "height": 30.0, "width": 12.0, "north_angle": 17.0, "label": label}
    # rank 0: two towers; rank 1: one duplicate of rank 0's second tower (tile seam) + one new; a third
    # rank-1 tower duplicates rank 1's own new one
    mine = [tower(0.0, 0.0, 3), tower(350.0, 0.0, 9)] if rank == 0 else \
           [tower(355.0, 5.0, 1), tower(700.0, 0.0, 4), tower(710.0, 0.0, 6)]
    merged = pdist.merge_towers(mine, 30.0)
    empty = pdist.merge_towers([], 30.0)
    pdist.FAST_CAP = 2          # rank 1 now exceeds the fixed block: the announced second gather must give the same list
    merged2 = pdist.merge_towers(mine, 30.0)
    assert [(t["rank"], t["label"]) for t in merged2] == [(t["rank"], t["label"]) for t in merged]
    q.put((rank, [(t["rank"], t["label"], t["center"].tolist()) for t in merged], len(empty),
           pdist.tile_for_rank(rank, 50)))
    dist.destroy_process_group()


def test_tower_merge_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, m0, e0, s0), (r1, m1, e1, s1) = out
    assert m0 == m1                                   # every rank holds the same merged list
    assert [(a, b) for a, b, _ in m0] == [(0, 3), (0, 9), (1, 4)]
    assert e0 == e1 == 0
    assert s0 == 0.0 and s1 == 50 * 350.0


def test_merge_without_process_group():
    from pointcloudhookup_b200 import dist as pdist
    t = [{"center": np.array([0.0, 0, 0]), "extent": np.ones(3), "rotation": np.eye(3), "height": 1.0, "width": 1.0,
          "north_angle": 0.0, "label": 1},
         {"center": np.array([10.0, 0, 0]), "extent": np.ones(3), "rotation": np.eye(3), "height": 1.0, "width": 1.0,
          "north_angle": 0.0, "label": 2}]
    assert [x["label"] for x in pdist.merge_towers(t, 30.0)] == [1]
    assert [x["label"] for x in pdist.merge_towers(t, 5.0)] == [1, 2]
