"""CPU: the halo-exchange protocol of tiles.tile_dbscan (DBSCAN row) — host logic, communication
patterns and the two-phase split — with an oracle-backed clusterer in place of the CUDA kernels.  Parity
definition: N tiles + halo == scikit-learn DBSCAN on the concatenated cloud (test/zzzzz.py:79-84), label for label."""
import os
import socket
import threading

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

EPS, MINPTS = 8.0, 20


def _reference(tiles):
    from sklearn.cluster import DBSCAN
    allp = np.concatenate(tiles)
    return DBSCAN(eps=EPS, min_samples=MINPTS, algorithm="ball_tree").fit(allp).labels_.astype(np.int32)


def test_merge_local_clusters_joins_by_shared_points_and_orders_by_first_core():
    from pointcloudhookup_b200 import tiles as tl
    BIG = tl.I64_MAX
    # rank 0: clusters 0 (cores from gid 5), 1 (from gid 40); rank 1: clusters 0 (halo only, = rank 0's 1), 1 (own, gid 120),
    # 2 (own from gid 100 AND shares a point with rank 0's cluster 0), 3 (outer-band halo only: dropped)
    tables = [np.array([5, 40]), np.array([BIG, 120, 100, BIG])]
    pairs = [np.array([[1, 1, 0]]),                      # rank 0: my cluster 1 is rank 1's cluster 0 (a point I sent)
             np.array([[0, 2, 0]])]                      # rank 1: my cluster 2 is rank 0's cluster 0 (a point I sent)
    maps, k = tl.merge_local_clusters(pairs, tables)
    assert k == 3
    assert maps[0].tolist() == [0, 1]                   # {r0c0, r1c2} first core 5 -> 0; {r0c1, r1c0} first core 40 -> 1
    assert maps[1].tolist() == [1, 2, 0, -1]            # r1c1 first core 120 -> 2; r1c3 has no own core anywhere -> -1
    maps, k = tl.merge_local_clusters([np.zeros((0, 3), np.int64)], [np.zeros(0, np.int64)])
    assert k == 0 and maps[0].tolist() == []


@pytest.mark.parametrize("world", [1, 2, 4])
def test_tile_dbscan_threads_equal_sklearn_on_the_concatenation(world):
    import halo_oracle as ho
    from pointcloudhookup_b200 import tiles as tl
    tiles = ho.corridor_candidates(7, world)
    ref = _reference(tiles)
    group = tl.ThreadComm.Group(world)
    out, errs = [None] * world, []

    def run(r):
        try:
            out[r] = tl.tile_dbscan(torch.from_numpy(tiles[r]), (1.0, 0.0), EPS, MINPTS, tl.ThreadComm(group, r), ho.OracleClusterer())
        except BaseException as e:      # a failing rank must not leave the others at the barrier
            errs.append(e)
            group.barrier.abort()
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    got = np.concatenate([o.labels.numpy() for o in out])
    assert np.array_equal(got, ref)                      # identical labels, not only the same partition
    k = int(ref.max()) + 1
    assert all(o.n_clusters == k for o in out)
    if world > 1:
        assert sum(sum(o.sent) for o in out) > 0 and any(len(set(ref[ref >= 0])) for o in out)
        # at least one cluster spans a cut (the conductor line / the blobs on the cuts)
        offs = np.concatenate([[0], np.cumsum([len(t) for t in tiles])])
        spans = [len({int(np.searchsorted(offs, i, side="right")) for i in np.nonzero(ref == c)[0]}) for c in range(k)]
        assert max(spans) >= 2
    # per-cluster table: all-reduced, identical on every rank, equal to the table of the concatenation
    allp = np.concatenate(tiles)
    for c in range(k):
        m = ref == c
        assert out[0].stats["count"][c] == m.sum()
        assert np.array_equal(out[0].stats["min"][c], allp[m].min(0)) and np.array_equal(out[0].stats["max"][c], allp[m].max(0))
        assert np.allclose(out[0].stats["sum"][c], allp[m].astype(np.float64).sum(0), rtol=1e-12)
    for o in out[1:]:
        assert o.stats.tobytes() == out[0].stats.tobytes()


def test_tiles_closer_than_two_eps_are_refused():
    import halo_oracle as ho
    from pointcloudhookup_b200 import tiles as tl
    tiles = ho.corridor_candidates(3, 3, per_tile=600, tile_len=10.0)     # tile 0 and tile 2 are 10 m apart
    group = tl.ThreadComm.Group(3)
    errs = []

    def run(r):
        try:
            tl.tile_dbscan(torch.from_numpy(tiles[r]), (1.0, 0.0), EPS, MINPTS, tl.ThreadComm(group, r), ho.OracleClusterer())
        except ValueError as e:
            errs.append(str(e))
    th = [threading.Thread(target=run, args=(r,)) for r in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert len(errs) == 3 and all("2*eps" in e for e in errs)


def test_a_failure_in_phase_one_on_one_rank_raises_on_every_rank():
    """The rank whose local clustering fails keeps to the exchange protocol, so its neighbours are not left waiting
    for labels that never come; the failure travels with the pair gather and every rank raises."""
    import halo_oracle as ho
    from pointcloudhookup_b200 import tiles as tl
    tiles = ho.corridor_candidates(5, 3, per_tile=600)

    class Failing(ho.OracleClusterer):
        def cores(self, P, eps, min_samples):
            raise ValueError("DBSCAN cell grid does not fit the packed key")
    group = tl.ThreadComm.Group(3)
    errs = [None] * 3

    def run(r):
        try:
            tl.tile_dbscan(torch.from_numpy(tiles[r]), (1.0, 0.0), EPS, MINPTS, tl.ThreadComm(group, r),
                           Failing() if r == 1 else ho.OracleClusterer())
        except (ValueError, RuntimeError) as e:
            errs[r] = e
    th = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(3)]
    [t.start() for t in th]
    [t.join(timeout=60) for t in th]
    assert not any(t.is_alive() for t in th), "a rank is still waiting for its failed neighbour"
    assert isinstance(errs[1], ValueError) and "packed key" in str(errs[1])
    assert all(isinstance(errs[r], RuntimeError) and "[1]" in str(errs[r]) for r in (0, 2))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import halo_oracle as ho
    from pointcloudhookup_b200 import tiles as tl
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = ho.corridor_candidates(11, world)
    comm = tl.TorchComm()
    res = tl.tile_dbscan(torch.from_numpy(tiles[rank]), (1.0, 0.0), EPS, MINPTS, comm, ho.OracleClusterer())
    q.put((rank, res.labels.numpy(), res.n_clusters, res.stats.tobytes(), res.sent, res.halo, comm.bytes_p2p))
    dist.destroy_process_group()


def test_tile_dbscan_world2_gloo():
    """Two processes over gloo: point-to-point halo send/recv, all-gathers of equivalences, all-reduced table."""
    import halo_oracle as ho
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    tiles = ho.corridor_candidates(11, 2)
    ref = _reference(tiles)
    assert np.array_equal(np.concatenate([out[0][1], out[1][1]]), ref)
    assert out[0][2] == out[1][2] == int(ref.max()) + 1
    assert out[0][3] == out[1][3]
    assert out[0][4][1] == out[1][5][0] > 0 and out[1][4][0] == out[0][5][1] > 0     # sent right == received from left, and back
    assert out[0][6] == 16 * sum(out[0][4]) + 4 * sum(out[0][5])          # halo out (16 B / point) + labels echoed back (4 B / point)


def test_merge_local_clusters_random_graphs_against_a_flood_fill():
    """Random cluster tables and shared-point pairs over 2..5 ranks: the union-find result equals a flood fill, global ids
    rank the components by their smallest own core index, components without any own core index get -1."""
    from pointcloudhookup_b200 import tiles as tl
    rng = np.random.default_rng(11)
    for _ in range(60):
        W = int(rng.integers(2, 6))
        sizes = [int(rng.integers(0, 7)) for _ in range(W)]
        keys = rng.permutation(10_000)[: sum(sizes)].astype(np.int64)
        tables, k = [], 0
        for s in sizes:
            t = keys[k: k + s].copy()
            k += s
            t[rng.uniform(size=s) < 0.2] = tl.I64_MAX                       # clusters seen only in the halo
            tables.append(t)
        pairs = []
        for r in range(W):
            rows = []
            for q in (r - 1, r + 1):
                if 0 <= q < W and sizes[r] and sizes[q]:
                    for _ in range(int(rng.integers(0, 4))):
                        rows.append((q, int(rng.integers(sizes[r])), int(rng.integers(sizes[q]))))
            pairs.append(np.array(rows, dtype=np.int64).reshape(-1, 3))
        maps, n_global = tl.merge_local_clusters(pairs, tables)
        # flood fill over (rank, local id) nodes
        nodes = [(r, a) for r in range(W) for a in range(sizes[r])]
        adj = {v: set() for v in nodes}
        for r, pr in enumerate(pairs):
            for q, a, b in pr.tolist():
                adj[(r, a)].add((q, b))
                adj[(q, b)].add((r, a))
        comp, comps = {}, []
        for v in nodes:
            if v in comp:
                continue
            stack, members = [v], []
            comp[v] = len(comps)
            while stack:
                u = stack.pop()
                members.append(u)
                for w in adj[u]:
                    if w not in comp:
                        comp[w] = len(comps)
                        stack.append(w)
            comps.append(members)
        comp_key = [min(int(tables[r][a]) for r, a in m) for m in comps]
        live = sorted((key, c) for c, key in enumerate(comp_key) if key < tl.I64_MAX)
        gid = {c: g for g, (_, c) in enumerate(live)}
        assert n_global == len(live)
        for r, a in nodes:
            assert maps[r][a] == gid.get(comp[(r, a)], -1)


def test_halo_widths_cover_eps_with_slack():
    from pointcloudhookup_b200 import tiles as tl
    for eps in (0.5, 8.0, 30.0):
        e1, e2 = tl.halo_widths(eps)
        assert eps < e1 < eps * 1.001 + 1e-5 and e2 == 2 * e1
