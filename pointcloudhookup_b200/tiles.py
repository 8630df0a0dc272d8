"""Whole-corridor clustering over spatial tiles with a halo exchange (DBSCAN row;
BASELINE.json configs[3]/[4]).

Parity definition: the un-chunked variant of the reference (test/zzzzz.py:79-84, ONE DBSCAN over the whole
filtered cloud) applied to the concatenation of all tiles' candidate points.  N ranks, each holding one
tile (its candidates in one common frame), must produce the labels that a single GPU produces on the
concatenated array — not merely up to a permutation: clusters are numbered by their smallest core index
in the concatenation, like scikit-learn does.

Protocol (one process per GPU; `Comm` hides torch.distributed, so the same code runs on NCCL, on gloo and
in-process over threads for a rank that holds several tiles):
  1. all-gather per rank: candidate count and the extent [smin, smax] of the candidates along the
     corridor axis  (24 B per rank).
  2. halo: every rank SENDS its candidates within 2*eps of a neighbour's extent to that neighbour
     (point-to-point, NCCL send/recv over NVLink: 16 B per halo point).  2*eps, not eps: a halo point within
     eps of the cut needs ITS whole eps-neighbourhood to have an exact core flag.
  3. local phase 1 (pch_dbscan_cores) on [left halo | own | right halo]: core flags + local cluster ids of
     the core points.  A point flagged core locally IS core globally (a truncated neighbourhood can only
     under-count), own points and halo points within eps of the cut are exact.
  4. the way back (NCCL send/recv, 4 B per halo point): every rank returns its local cluster ids of the halo points
     to their owners.  A point that is core on both sides joins the two local clusters; each rank reduces its
     thousands of shared points to the handful of distinct (my cluster, neighbour's cluster) pairs.  all-gather:
     those pairs and, per local cluster, the smallest global id of its OWN core points.  Union-find on every rank
     (identical input -> identical result), global ids = rank of the component's smallest core point.
  5. local phase 2 (pch_dbscan_finish): core labels rewritten with global ids, border points take the
     smallest GLOBAL id among adjacent clusters, per-cluster count / AABB / sums over own points only.
  6. all-reduce of the per-cluster table (count, sums: sum; AABB: min / max): K x 56 B.
Every true core-core edge is seen by the rank that owns one endpoint (the other endpoint is in its halo
and exact there), so the union of the local components is exactly the global clustering.
"""
from __future__ import annotations

import dataclasses
import math
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

I64_MAX = np.iinfo(np.int64).max
STATS_DTYPE = np.dtype([("count", "<i8"), ("min", "<f4", 3), ("max", "<f4", 3), ("sum", "<f8", 3)])


# -------------------------------------------------------------------------------------------------
# communication: the three patterns the protocol needs
# -------------------------------------------------------------------------------------------------
class Comm:
    rank: int = 0
    world: int = 1

    def all_gather_np(self, arr: np.ndarray) -> List[np.ndarray]:
        """Every rank's 1-D array (sizes may differ), in rank order, on every rank."""
        raise NotImplementedError

    def all_gather_fixed(self, arr: np.ndarray) -> List[np.ndarray]:
        """The same for arrays whose size is the same on every rank (one collective instead of two)."""
        return self.all_gather_np(arr)

    def neighbour_exchange(self, to_left: Optional[torch.Tensor], to_right: Optional[torch.Tensor]):
        """Send (k,4) float32 payloads to rank-1 / rank+1, receive theirs: (from_left, from_right)."""
        raise NotImplementedError

    def neighbour_echo(self, back_left: Optional[torch.Tensor], back_right: Optional[torch.Tensor], n_left: int, n_right: int):
        """The way back: int32 tensors for the points I RECEIVED go to the neighbour they came from; I get the
        neighbours' tensors for the n_left / n_right points I sent them (sizes are known on both sides)."""
        raise NotImplementedError

    def all_reduce_stats(self, stats: np.ndarray) -> np.ndarray:
        out = stats.copy()
        parts = self.all_gather_fixed(stats.view(np.uint8).reshape(-1))
        rows = [p.view(STATS_DTYPE) for p in parts]
        out["count"] = np.sum([p["count"] for p in rows], axis=0)
        out["sum"] = np.sum([p["sum"] for p in rows], axis=0)          # rank order: the same bits on every rank
        out["min"] = np.min([p["min"] for p in rows], axis=0)
        out["max"] = np.max([p["max"] for p in rows], axis=0)
        return out


class SoloComm(Comm):
    def all_gather_np(self, arr):
        return [np.asarray(arr)]

    def neighbour_exchange(self, to_left, to_right):
        return None, None

    def neighbour_echo(self, back_left, back_right, n_left, n_right):
        return None, None


class TorchComm(Comm):
    """torch.distributed: NCCL with device tensors (halo payloads travel GPU to GPU), gloo with host tensors."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.nccl = dist.get_backend() == "nccl"
        self.device = torch.device(device if device is not None else
                                   (f"cuda:{torch.cuda.current_device()}" if self.nccl else "cpu"))
        self.bytes_p2p = 0
        self.bytes_gather = 0

    def _gather_fixed(self, t: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        try:
            self.dist.all_gather_into_tensor(out, t)
        except (RuntimeError, NotImplementedError, AttributeError):
            parts = [torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(parts, t)
            out = torch.stack(parts)
        return out

    def all_gather_np(self, arr):
        a = np.ascontiguousarray(arr)
        raw = a.view(np.uint8).reshape(-1)
        sizes = self._gather_fixed(torch.tensor([raw.size], dtype=torch.int64, device=self.device)).cpu().numpy().reshape(-1)
        mx = int(sizes.max())
        pad = np.zeros(max(mx, 1), dtype=np.uint8)
        pad[: raw.size] = raw
        got = self._gather_fixed(torch.from_numpy(pad).to(self.device)).cpu().numpy()
        self.bytes_gather += int(sizes.sum())
        return [got[r, : int(sizes[r])].copy().view(a.dtype) for r in range(self.world)]

    def all_gather_fixed(self, arr):
        a = np.ascontiguousarray(arr)
        raw = a.view(np.uint8).reshape(-1)
        if raw.size == 0:
            return [a.copy() for _ in range(self.world)]
        got = self._gather_fixed(torch.from_numpy(raw.copy()).to(self.device)).cpu().numpy()
        self.bytes_gather += int(raw.size) * self.world
        return [got[r].copy().view(a.dtype) for r in range(self.world)]

    def neighbour_exchange(self, to_left, to_right):
        dist = self.dist
        dev = self.device
        kl = 0 if to_left is None else int(to_left.shape[0])
        kr = 0 if to_right is None else int(to_right.shape[0])
        sizes = self._gather_fixed(torch.tensor([kl, kr], dtype=torch.int64, device=dev)).cpu().numpy()
        ops, from_left, from_right = [], None, None
        r, W = self.rank, self.world
        if r > 0 and int(sizes[r - 1, 1]) > 0:
            from_left = torch.empty((int(sizes[r - 1, 1]), 4), dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, from_left, r - 1))
        if r + 1 < W and int(sizes[r + 1, 0]) > 0:
            from_right = torch.empty((int(sizes[r + 1, 0]), 4), dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, from_right, r + 1))
        if r > 0 and kl > 0:
            ops.append(dist.P2POp(dist.isend, to_left.to(dev).contiguous(), r - 1))
        if r + 1 < W and kr > 0:
            ops.append(dist.P2POp(dist.isend, to_right.to(dev).contiguous(), r + 1))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.bytes_p2p += 16 * (kl + kr)
        return from_left, from_right

    def neighbour_echo(self, back_left, back_right, n_left, n_right):
        dist, dev = self.dist, self.device
        r, W = self.rank, self.world
        ops, got_left, got_right = [], None, None
        if r > 0 and n_left > 0:
            got_left = torch.empty(n_left, dtype=torch.int32, device=dev)
            ops.append(dist.P2POp(dist.irecv, got_left, r - 1))
        if r + 1 < W and n_right > 0:
            got_right = torch.empty(n_right, dtype=torch.int32, device=dev)
            ops.append(dist.P2POp(dist.irecv, got_right, r + 1))
        if r > 0 and back_left is not None and back_left.numel():
            ops.append(dist.P2POp(dist.isend, back_left.to(dev).contiguous(), r - 1))
            self.bytes_p2p += 4 * back_left.numel()
        if r + 1 < W and back_right is not None and back_right.numel():
            ops.append(dist.P2POp(dist.isend, back_right.to(dev).contiguous(), r + 1))
            self.bytes_p2p += 4 * back_right.numel()
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return got_left, got_right


class ThreadComm(Comm):
    """W ranks as threads of one process (a GPU that holds several tiles; the CPU tests): the collectives are
    rendezvous on a barrier.  Kernels of different ranks never wait on each other — only host threads do."""

    class Group:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, group: "ThreadComm.Group", rank: int):
        self.g, self.rank, self.world = group, rank, group.world

    def _exchange(self, value):
        g = self.g
        g.slots[self.rank] = value
        g.barrier.wait()
        got = list(g.slots)
        g.barrier.wait()
        return got

    def all_gather_np(self, arr):
        return [np.asarray(a).copy() for a in self._exchange(np.ascontiguousarray(arr))]

    def neighbour_exchange(self, to_left, to_right):
        got = self._exchange((to_left, to_right))
        r = self.rank
        from_left = got[r - 1][1] if r > 0 else None
        from_right = got[r + 1][0] if r + 1 < self.world else None
        fix = lambda t: None if t is None or t.shape[0] == 0 else t.clone()
        return fix(from_left), fix(from_right)

    def neighbour_echo(self, back_left, back_right, n_left, n_right):
        got = self._exchange((back_left, back_right))
        r = self.rank
        from_left = got[r - 1][1] if (r > 0 and n_left) else None          # the left neighbour's labels for what I sent it
        from_right = got[r + 1][0] if (r + 1 < self.world and n_right) else None
        fix = lambda t: None if t is None else t.clone()
        return fix(from_left), fix(from_right)


# -------------------------------------------------------------------------------------------------
# the per-rank kernels behind an interface (the CPU tests plug in an oracle-backed one)
# -------------------------------------------------------------------------------------------------
class DeviceClusterer:
    """libpch_b200 kernels on the current CUDA device."""

    def __init__(self):
        from . import _native, device as dv
        self.dv, self.lib, self.check = dv, _native.lib(), _native.check
        dv._require_cuda()
        self._ws = None

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def extent(self, P: torch.Tensor, axis) -> Tuple[float, float]:
        if P.shape[0] == 0:
            return math.inf, -math.inf
        mm = torch.empty(2, dtype=torch.float64, device=P.device)
        self.check(self.lib.pch_axis_extent(P.data_ptr(), P.shape[0], float(axis[0]), float(axis[1]), mm.data_ptr(),
                                            self._stream()), "pch_axis_extent")
        lo, hi = mm.cpu().numpy()
        return float(lo), float(hi)

    def band_pair(self, P: torch.Tensor, axis, hi_left: Optional[float], lo_right: Optional[float]):
        """(indices with s <= hi_left, indices with s >= lo_right), int32 ascending; None = no neighbour on that side.
        Both selections are launched before the ONE device->host read of their two counts."""
        n = P.shape[0]
        dev = P.device
        empty = torch.zeros(0, dtype=torch.int32, device=dev)
        if n == 0 or (hi_left is None and lo_right is None):
            return empty, empty
        st = self._stream()
        cnt = torch.zeros(2, dtype=torch.int64, device=dev)
        outs = [None, None]
        for k, (lo, hi) in enumerate(((-math.inf, hi_left), (lo_right, math.inf))):
            if (k == 0 and hi_left is None) or (k == 1 and lo_right is None):
                continue
            mask = torch.empty(n + 16, dtype=torch.uint8, device=dev)[:n]
            self.check(self.lib.pch_axis_band_mask(P.data_ptr(), n, float(axis[0]), float(axis[1]), float(lo), float(hi), None,
                                                   mask.data_ptr(), st), "pch_axis_band_mask")
            src = torch.empty(n, dtype=torch.int32, device=dev)
            wsb = self.lib.pch_compact_workspace_bytes(n)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            self.check(self.lib.pch_compact_points(P.data_ptr(), None, mask.data_ptr(), n, None, 0.0, None, src.data_ptr(), None,
                                                   cnt[k:].data_ptr(), ws.data_ptr(), wsb, st), "pch_compact_points")
            outs[k] = src
        c = cnt.cpu().numpy()
        return (outs[0][: int(c[0])] if outs[0] is not None else empty,
                outs[1][: int(c[1])] if outs[1] is not None else empty)

    def cores(self, P: torch.Tensor, eps: float, min_samples: int):
        G = P.shape[0]
        dev = P.device
        labels = torch.empty(G, dtype=torch.int32, device=dev)
        cap = max(4096, G // 256)
        wsb = self.lib.pch_dbscan_fused_workspace_bytes(G, G, cap)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        self.check(self.lib.pch_dbscan_cores(P.data_ptr(), G, G, float(eps), int(min_samples), labels.data_ptr(), cap,
                                             ws.data_ptr(), wsb, self._stream()), "pch_dbscan_cores")
        sc = ws[:256].cpu().numpy()
        if int(sc[208 + 24: 208 + 28].view(np.int32)[0]) != 0:
            raise ValueError("DBSCAN cell grid does not fit the packed key")
        if int(sc[:4].view(np.int32)[0]):
            raise RuntimeError("device look-back spin limit hit in dbscan")
        k = int(sc[136:144].view(np.int64)[0])
        self._ws = (P, G, float(eps), int(min_samples), cap, ws, wsb)
        return labels, k

    def shared_report(self, labels: torch.Tensor, pos: torch.Tensor, extra: torch.Tensor, lo: int, hi: int, base: int, k: int):
        """ONE device->host read for everything the exchange needs from phase 1: the labels at the local positions
        `pos` (int32), the int32 payload `extra` (the senders' indices of the halo I received), and the table of the
        smallest OWN core index (base + i - lo over i in [lo, hi)) per local cluster."""
        dev = labels.device
        table = torch.full((max(k, 1),), I64_MAX, dtype=torch.int64, device=dev)
        if k and hi > lo:
            self.check(self.lib.pch_label_min_index(labels.data_ptr(), lo, hi, int(base), k, table.data_ptr(), self._stream()),
                       "pch_label_min_index")
        lab = labels[pos.long()].to(torch.int64) if pos.numel() else torch.zeros(0, dtype=torch.int64, device=dev)
        both = torch.cat([table[:k], lab, extra.to(torch.int64)]).cpu().numpy()
        return both[k: k + pos.numel()].astype(np.int32), both[k + pos.numel():], both[:k].copy()

    def finish(self, label_map: np.ndarray, n_global: int, own_lo: int, own_hi: int):
        P, G, eps, min_samples, cap, ws, wsb = self._ws
        dev = P.device
        m = torch.from_numpy(np.ascontiguousarray(label_map, dtype=np.int32)).to(dev) if len(label_map) else \
            torch.zeros(1, dtype=torch.int32, device=dev)
        labels = torch.empty(G, dtype=torch.int32, device=dev)
        item = STATS_DTYPE.itemsize
        stats = torch.zeros(max(n_global, 1) * item, dtype=torch.uint8, device=dev)
        acc = torch.empty(self.lib.pch_dbscan_acc_bytes(n_global), dtype=torch.uint8, device=dev)
        self.check(self.lib.pch_dbscan_finish(P.data_ptr(), G, G, eps, min_samples, m.data_ptr(), int(n_global), int(own_lo),
                                              int(own_hi), labels.data_ptr(), stats.data_ptr(), acc.data_ptr(), cap,
                                              ws.data_ptr(), wsb, self._stream()), "pch_dbscan_finish")
        host = stats[: n_global * item].cpu().numpy().view(STATS_DTYPE).copy()
        self._ws = None
        return labels, host


# -------------------------------------------------------------------------------------------------
# host logic shared by every rank: which local clusters are the same global cluster
# -------------------------------------------------------------------------------------------------
def merge_local_clusters(pairs: Sequence[np.ndarray], tables: Sequence[np.ndarray]):
    """pairs[r]: int64 (p_r, 3) rows (neighbour rank q, local cluster id on rank r, local cluster id on rank q): the two
    clusters contain the same core point;  tables[r]: int64 [K_r] smallest global id of rank r's OWN core points per
    local cluster (I64_MAX = none).  Returns (maps, n_global): maps[r][local id] = global id (-1: a cluster without
    any own core point anywhere, i.e. seen only in the outer halo band).  Global ids rank the clusters by their
    smallest core point, the order scikit-learn numbers them in on the concatenated cloud."""
    W = len(tables)
    node_off = np.concatenate([[0], np.cumsum([len(t) for t in tables])]).astype(np.int64)
    n_nodes = int(node_off[-1])
    key = np.concatenate([np.asarray(t, dtype=np.int64) for t in tables]) if n_nodes else np.zeros(0, np.int64)
    parent = list(range(n_nodes))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for r, pr in enumerate(pairs):
        for q, a, b in np.asarray(pr, dtype=np.int64).reshape(-1, 3).tolist():
            ra, rb = find(int(node_off[r]) + a), find(int(node_off[q]) + b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
    roots = np.array([find(i) for i in range(n_nodes)], dtype=np.int64)
    comp_key = np.full(n_nodes, I64_MAX, dtype=np.int64)
    np.minimum.at(comp_key, roots, key)
    live = np.nonzero((roots == np.arange(n_nodes)) & (comp_key < I64_MAX))[0]
    order = live[np.argsort(comp_key[live], kind="stable")]
    gid_of_root = np.full(n_nodes, -1, dtype=np.int64)
    gid_of_root[order] = np.arange(len(order))
    node_gid = gid_of_root[roots] if n_nodes else np.zeros(0, np.int64)
    maps = [node_gid[node_off[r]: node_off[r + 1]].astype(np.int32) for r in range(W)]
    return maps, int(len(order))


@dataclasses.dataclass
class TileDbscanResult:
    labels: torch.Tensor          # int32 [G_own]: global cluster ids of this rank's candidates
    n_clusters: int               # K, the same on every rank
    stats: np.ndarray             # STATS_DTYPE [K], reduced over all ranks
    offset: int                   # global id of this rank's first candidate
    counts: List[int]             # candidates per rank
    halo: Tuple[int, int]         # halo points received from the left / right neighbour
    sent: Tuple[int, int]         # halo points sent to the left / right neighbour


def halo_widths(eps: float) -> Tuple[float, float]:
    """(zone in which a neighbour's points can be within eps of mine, halo width) with a little slack for the
    rounding of the projection: supersets are harmless, a missed point is not."""
    e1 = float(eps) * (1.0 + 1e-6) + 1e-6
    return e1, 2.0 * e1


class _Trace:
    """PCH_TRACE_TILES=1: wall time of every step of the protocol on rank 0 (synchronising the device at each mark)."""

    def __init__(self, rank):
        import os
        import time
        self.on = bool(os.environ.get("PCH_TRACE_TILES")) and rank == 0
        self.time = time
        self.t = time.perf_counter()
        self.rows = []

    def mark(self, name):
        if not self.on:
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        now = self.time.perf_counter()
        self.rows.append(f"{name} {1e3 * (now - self.t):.2f}")
        self.t = now

    def done(self, head):
        if self.on:
            print(f"[pch tiles] {head}: " + " | ".join(self.rows) + " ms", flush=True)


def tile_dbscan(P_own: torch.Tensor, axis: Sequence[float], eps: float, min_samples: int, comm: Comm,
                clusterer=None) -> TileDbscanResult:
    """One whole-corridor DBSCAN over the tiles of all ranks; see the module docstring."""
    clu = clusterer if clusterer is not None else DeviceClusterer()
    tr = _Trace(comm.rank)
    r, W = comm.rank, comm.world
    ax = np.asarray(axis, dtype=np.float64)
    ax = ax / np.linalg.norm(ax)
    G = int(P_own.shape[0])
    smin, smax = clu.extent(P_own, ax)
    tr.mark("extent")
    meta = comm.all_gather_fixed(np.array([G, smin, smax], dtype=np.float64))
    tr.mark("gather meta")
    counts = [int(m[0]) for m in meta]
    lo_s = [float(m[1]) for m in meta]
    hi_s = [float(m[2]) for m in meta]
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    _, e2 = halo_widths(eps)
    for i in range(W):
        for j in range(i + 2, W):
            if counts[i] and counts[j] and not (hi_s[i] + e2 < lo_s[j] or hi_s[j] + e2 < lo_s[i]):
                raise ValueError(f"tiles {i} and {j} are closer than 2*eps along the axis: only neighbouring tiles may touch")
    dev = P_own.device
    left = r > 0 and counts[r - 1] > 0 and G > 0
    right = r + 1 < W and counts[r + 1] > 0 and G > 0
    idx_l, idx_r = clu.band_pair(P_own, ax, hi_s[r - 1] + e2 if left else None, lo_s[r + 1] - e2 if right else None)

    def payload(idx):
        if idx.numel() == 0:
            return None
        rows = P_own.index_select(0, idx.long())
        return torch.cat([rows, idx.to(torch.int32).view(torch.float32).unsqueeze(1)], dim=1).contiguous()

    tr.mark("bands")
    from_left, from_right = comm.neighbour_exchange(payload(idx_l), payload(idx_r))
    tr.mark("halo p2p")
    nL = 0 if from_left is None else int(from_left.shape[0])
    nR = 0 if from_right is None else int(from_right.shape[0])
    parts = []
    if nL:
        parts.append(from_left.to(dev)[:, :3])
    parts.append(P_own)
    if nR:
        parts.append(from_right.to(dev)[:, :3])
    n_local = nL + G + nR
    empty_i = torch.zeros(0, dtype=torch.int32, device=dev)
    pair_rows, table = np.zeros((0, 3), np.int64), np.zeros(0, np.int64)
    labels_core, k_local, failure = None, 0, None
    if n_local:
        P_local = torch.cat(parts).contiguous() if len(parts) > 1 else P_own.contiguous()
        try:
            labels_core, k_local = clu.cores(P_local, eps, min_samples)
        except (ValueError, RuntimeError) as e:
            # a rank that stopped here would leave its neighbours waiting in the echo: it keeps to the protocol with
            # "no cluster anywhere" and reports the failure in the gather below, where EVERY rank raises
            failure = e
            labels_core, k_local = torch.full((n_local,), -1, dtype=torch.int32, device=dev), 0
        tr.mark("cores")
    # the way back: what I think of the halo points goes to their owners, device to device; I learn what the
    # neighbours think of the points I sent them.  A point that is core on both sides joins the two local clusters.
    echo_l, echo_r = comm.neighbour_echo(labels_core[:nL].contiguous() if nL else None,
                                         labels_core[nL + G:].contiguous() if nR else None,
                                         int(idx_l.numel()), int(idx_r.numel()))
    tr.mark("echo p2p")
    if n_local and failure is None:
        n_sl, n_sr = (0 if echo_l is None else int(echo_l.numel())), (0 if echo_r is None else int(echo_r.numel()))
        pos = torch.cat([idx_l[:n_sl] + nL, idx_r[:n_sr] + nL]) if (n_sl + n_sr) else empty_i
        extra = torch.cat([t.to(dev) for t in (echo_l, echo_r) if t is not None]) if (n_sl + n_sr) else empty_i
        mine, theirs, table = clu.shared_report(labels_core, pos, extra, nL, nL + G, int(offs[r]), k_local)
        rows = []
        for side, lo_, hi_ in ((r - 1, 0, n_sl), (r + 1, n_sl, n_sl + n_sr)):
            a, b = mine[lo_:hi_].astype(np.int64), theirs[lo_:hi_].astype(np.int64)
            ok = (a >= 0) & (b >= 0)
            if ok.any():
                u = np.unique(a[ok] * (1 << 32) + b[ok])        # thousands of shared points, a handful of cluster pairs
                rows.append(np.stack([np.full(len(u), side, np.int64), u >> 32, u & 0xffffffff], axis=1))
        if rows:
            pair_rows = np.concatenate(rows)
    tr.mark("report")
    # one variable-size all-gather carries both the cluster pairs and the per-cluster tables
    packed = np.concatenate([[len(pair_rows), len(table), int(failure is not None)], pair_rows.reshape(-1), table]).astype(np.int64)
    got = comm.all_gather_np(packed)
    tr.mark("gather pairs")
    failed = [q for q, g in enumerate(got) if int(g[2])]
    if failure is not None:
        raise failure
    if failed:
        raise RuntimeError(f"whole-corridor DBSCAN: phase 1 failed on rank(s) {failed}")
    entries = [g[3: 3 + 3 * int(g[0])].reshape(-1, 3) for g in got]
    tables = [g[3 + 3 * int(g[0]): 3 + 3 * int(g[0]) + int(g[1])] for g in got]
    maps, n_global = merge_local_clusters(entries, tables)
    tr.mark("merge")
    if n_local:
        labels_all, stats_own = clu.finish(maps[r], n_global, nL, nL + G)
        labels = labels_all[nL: nL + G]
    else:
        labels = torch.zeros(0, dtype=torch.int32, device=dev)
        stats_own = np.zeros(n_global, dtype=STATS_DTYPE)
    if len(stats_own) != n_global:
        stats_own = np.zeros(n_global, dtype=STATS_DTYPE)
    empty = stats_own["count"] == 0
    stats_own["min"][empty] = np.inf
    stats_own["max"][empty] = -np.inf
    tr.mark("finish")
    stats = comm.all_reduce_stats(stats_own)
    tr.mark("reduce stats")
    tr.done(f"tile_dbscan G={G} halo={nL}+{nR} K={n_global}")
    return TileDbscanResult(labels, n_global, stats, int(offs[r]), counts, (nL, nR), (int(idx_l.numel()), int(idx_r.numel())))
