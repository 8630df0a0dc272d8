"""The fused hot path: raw LAS records in HBM -> voxel downsample -> (LAS re-quantisation, float32
read-back) -> centroid / height filter -> chunked DBSCAN -> per-cluster reduction -> tower list.

This is pyGUI_towers_test.py:344-368 (downsample_and_extract) without the LAS file between the two
stages: the intermediate ``output/point_2.las`` is reproduced arithmetically (means -> int32 lattice
-> float64 -> float32) inside the reduce kernel, so results are identical to the two-step run.
"""
from __future__ import annotations

import dataclasses
import os
import queue
import threading
from typing import List, Optional

import numpy as np
import torch

from . import device as dv
from . import towers as tw


@dataclasses.dataclass
class PipelineResult:
    n_points: int
    n_voxels: int
    n_candidates: int
    n_clusters: int
    towers: List[dict]
    stages: Optional[tw.TowerStages] = None
    voxel_plan: Optional[dict] = None
    db_plan: Optional[dict] = None


class nvtx_range:
    """NVTX range around a pipeline stage (shows up in Nsight Systems / ncu --nvtx; free when no tool is attached)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


def run_pipeline(dl: dv.DeviceLas, voxel_size: float = 0.1, chunk_size: int = 500000, eps: float = 8.0,
                 min_points: int = 80, ground: str = "percentile", box: str = "aabb", want_points: bool = False,
                 keep_stages: bool = False, **tower_kw) -> PipelineResult:
    with nvtx_range("pch.voxel_downsample"):
        vres = dv.voxel_downsample(dl, voxel_size, chunk_size, want=("f32", "z32") if ground == "percentile" else ("f32",))
    if vres.count == 0:
        return PipelineResult(dl.n, 0, 0, 0, [])
    with nvtx_range("pch.ground+dbscan"):
        stages = tw.run_stages(vres.f32, eps, min_points, ground, zcol=vres.z32)
    with nvtx_range("pch.towers"):
        towers = tw.select_towers(stages, box=box, want_points=want_points, **tower_kw)
    return PipelineResult(dl.n, vres.count, int(stages.filtered.shape[0]), stages.n_clusters, towers,
                          stages if keep_stages else None, vres.plan, stages.db_plan)


_acquire_staging = dv.acquire_staging      # one bounded pinned pool for every host->device path (device.py)
_release_staging = dv.release_staging


host_threads = dv.host_threads
FEED_TIMEOUT_S = 120.0      # a slice that has not arrived by then is a bug or a dead feeder, never a wait worth keeping


AUTO_PACK_MIN_THREADS = 12   # measured on the B200 hosts: 16 threads gather 3.4 GB in 25 ms (PCIe copy: 64 ms), 8 threads tie, 4 lose


def resolve_pack(pack: str, threads: int = 0) -> str:
    """"auto" -> "xyz" when this process has enough host threads for the gather to outrun the whole-record copy."""
    if pack != "auto":
        return pack
    return "xyz" if (threads or host_threads()) >= AUTO_PACK_MIN_THREADS else "none"


def slice_plan(n: int, rec_len: int, chunk_size: int, slice_chunks: int, pack: str, raw_every: int, pinned: bool):
    """How one tile crosses PCIe: (chunk size used, [(first, last) point of each slice], [device record length of
    each slice]).  Slices are whole chunks (chunks are independent, so the voxel stage can run per slice); a slice
    start must be 16-byte aligned in both layouts (34 B records and the 12 B stream), else the tile is one slice;
    with pack="xyz" every raw_every-th slice of a PINNED source is shipped as whole records instead of gathered."""
    if pack not in ("none", "xyz"):
        raise ValueError(f"unknown pack mode {pack!r}")
    n = int(n)
    cs = max(1, min(int(chunk_size), max(n, 1)))
    per_slice = cs * max(1, int(slice_chunks))
    if (cs * 12) % 16 or (cs * int(rec_len)) % 16:
        per_slice = max(n, 1)
    if not pinned:
        raw_every = 0
    bounds = [(lo, min(lo + per_slice, n)) for lo in range(0, n, per_slice)]
    lens = [int(rec_len) if (pack == "none" or (raw_every > 0 and i % raw_every == raw_every - 1)) else 12
            for i in range(len(bounds))]
    return cs, bounds, lens


class _TileFeed:
    """One tile's host -> device transfer plan: slices of whole chunks, each either gathered to the 12-byte X,Y,Z
    stream (pch_host_pack_xyz) or shipped as whole records, each with its own device buffer and CUDA event."""

    def __init__(self, host_records, n, rec_len, chunk_size, slice_chunks, pack, raw_every, device, slot=0, stage=None):
        self.n, self.rec_len, self.pack, self.slot, self.device = int(n), int(rec_len), pack, slot, device
        self.stage = stage          # pinned staging for the gathered stream (pack="xyz"); set before feed()
        if pack not in ("none", "xyz"):
            raise ValueError(f"unknown pack mode {pack!r}")
        if isinstance(host_records, torch.Tensor):
            host = host_records.view(torch.uint8).reshape(-1)[: self.n * rec_len]
        else:
            host = torch.from_numpy(np.asarray(host_records).view(np.uint8).reshape(-1)[: self.n * rec_len])
        if pack == "none" and not host.is_pinned():
            host = host.pin_memory()
        self.host = host
        self.cs, self.bounds, self.lens = slice_plan(self.n, rec_len, chunk_size, slice_chunks, pack, raw_every,
                                                     host.is_pinned())
        self.bufs = []
        for (lo, hi), ln in zip(self.bounds, self.lens):
            b = torch.empty(dv.padded_bytes(hi - lo, ln), dtype=torch.uint8, device=device)
            b[(hi - lo) * ln:].zero_()
            self.bufs.append(b)
        self.zeroed = torch.cuda.Event()
        self.zeroed.record(torch.cuda.current_stream(device))
        self.events = [torch.cuda.Event() for _ in self.bounds]
        self.ready = [threading.Event() for _ in self.bounds]
        self.failure = []

    def feed(self, copy_stream, threads):
        """Gather (pack="xyz") and enqueue the copies, slice by slice.  Runs on a worker thread for pack="xyz"
        (ctypes releases the GIL during the gather)."""
        try:
            torch.cuda.set_device(self.device)
            copy_stream.wait_event(self.zeroed)
            stage = self.stage
            lib = dv._native.lib()
            for i, (lo, hi) in enumerate(self.bounds):
                if self.lens[i] == 12 and self.pack == "xyz":
                    dv.check(lib.pch_host_pack_xyz(self.host.data_ptr() + lo * self.rec_len, hi - lo, self.rec_len,
                                                   stage.data_ptr() + lo * 12, threads), "pch_host_pack_xyz")
                    src = stage[lo * 12: hi * 12]
                else:
                    src = self.host[lo * self.rec_len: hi * self.rec_len]
                with torch.cuda.stream(copy_stream):
                    self.bufs[i][: src.numel()].copy_(src, non_blocking=True)
                    self.events[i].record(copy_stream)
                self.ready[i].set()
        except BaseException as e:     # surface the failure on the consuming thread
            self.fail(e)

    def fail(self, e):
        self.failure.append(e)
        for r in self.ready:
            r.set()

    def wait_copied(self):
        if self.events:
            self.events[-1].synchronize()


def _consume(feed: _TileFeed, scales, offsets, voxel_size, kw) -> PipelineResult:
    """Voxel stage per slice as the slices land, then the ground / tower stages on the whole tile."""
    kw = dict(kw)
    device, n = feed.device, feed.n
    main = torch.cuda.current_stream(device)
    ground = kw.pop("ground", "percentile")
    sink = dv.VoxelSink(max(n, 1), device, want_z=(ground == "percentile"))
    for i, (lo, hi) in enumerate(feed.bounds):
        if not feed.ready[i].wait(timeout=FEED_TIMEOUT_S):
            raise RuntimeError(f"host feed stalled: slice {i} not delivered within {FEED_TIMEOUT_S} s")
        if feed.failure:
            raise feed.failure[0]
        main.wait_event(feed.events[i])
        view = dv.DeviceLas(feed.bufs[i], hi - lo, feed.lens[i], np.asarray(scales, dtype=np.float64),
                            np.asarray(offsets, dtype=np.float64))
        dv.voxel_downsample(view, voxel_size, feed.cs, want=(), sink=sink)
    total = sink.count
    if total == 0:
        return PipelineResult(n, 0, 0, 0, [])
    f32 = sink.f32[:total]
    eps = kw.pop("eps", 8.0)
    min_points = kw.pop("min_points", 80)
    box = kw.pop("box", "aabb")
    want_points = kw.pop("want_points", False)
    stages = tw.run_stages(f32, eps, min_points, ground, zcol=sink.z32[:total] if sink.z32 is not None else None)
    towers = tw.select_towers(stages, box=box, want_points=want_points, **kw)
    return PipelineResult(n, total, int(stages.filtered.shape[0]), stages.n_clusters, towers, None, None, stages.db_plan)


def run_pipeline_from_host(host_records, n: int, rec_len: int, scales, offsets, voxel_size: float = 0.1,
                           chunk_size: int = 500000, slice_chunks: int = 10, device=None, pack: str = "none",
                           threads: int = 0, raw_every: int = 0, **kw) -> PipelineResult:
    """End-to-end entry for HOST record buffers.  The transfer is cut into slices of whole chunks on a copy
    stream, and the voxel stage of slice i runs while slice i+1 is still in flight (chunks are independent,
    so the result is identical to one big call).  The tower stage follows on the concatenated float32 cloud.

    pack="none": `host_records` is a (preferably pinned) uint8 tensor and whole records cross PCIe.
    pack="xyz":  `host_records` is any host uint8 tensor / numpy array (pageable is fine); a worker thread
                 gathers the 12 X,Y,Z bytes of each record into pinned staging (pch_host_pack_xyz, host
                 threads, no arithmetic) slice by slice, so a 34-byte record crosses PCIe as 12 bytes and the
                 gather of slice i+1 overlaps the copy and the voxel stage of slice i.
    raw_every=k (with pack="xyz" and a pinned source): every k-th slice skips the gather and crosses PCIe as
                 whole records while the host threads are busy with the slices around it, so the DMA engine
                 and the host cores both carry part of the stream.
    pack="auto": "xyz" when this process has at least AUTO_PACK_MIN_THREADS host threads, else "none"."""
    dv._require_cuda()
    pack = resolve_pack(pack, threads)
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    copy_stream = torch.cuda.Stream(device=device)
    stage = _acquire_staging(n * 12) if pack == "xyz" else None
    feed = _TileFeed(host_records, n, rec_len, chunk_size, slice_chunks, pack, raw_every, device, stage=stage)
    nt = threads or host_threads()
    worker = None
    if pack == "xyz":
        worker = threading.Thread(target=feed.feed, args=(copy_stream, nt), name="pch-host-pack", daemon=True)
        worker.start()
    else:
        feed.feed(copy_stream, nt)
    try:
        return _consume(feed, scales, offsets, voxel_size, kw)
    finally:
        if worker is not None:
            worker.join()
        feed.wait_copied()
        _release_staging(stage)


def run_tiles_from_host(tiles, scales, offsets, voxel_size: float = 0.1, chunk_size: int = 500000,
                        slice_chunks: int = 10, device=None, pack: str = "xyz", threads: int = 0, raw_every: int = 0,
                        **kw):
    """A corridor as a STREAM of tiles on one GPU (more tiles than GPUs): generator yielding one
    PipelineResult per tile of `tiles` (an iterable of (host_records, n_points, rec_len)), in order.  Each
    tile is processed exactly like run_pipeline_from_host; in addition, while tile k is in its ground / tower
    stages the records of tile k+1 are already being gathered and copied (one tile of look-ahead, two pinned
    staging buffers), so in steady state the device never waits for PCIe and the host never waits for the
    device."""
    dv._require_cuda()
    pack = resolve_pack(pack, threads)
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    copy_stream = torch.cuda.Stream(device=device)
    nt = threads or host_threads()
    jobs: "queue.Queue" = queue.Queue()

    stages = {}                         # staging slot -> pinned buffer (owned by this call until the end)

    def feeder():
        last = {}                       # staging slot -> the feed that last used it
        while True:
            f = jobs.get()
            if f is None:
                return
            try:
                prev = last.get(f.slot)
                if prev is not None:
                    prev.wait_copied()      # its copies have finished reading this staging buffer
                if pack == "xyz":
                    buf = stages.get(f.slot)
                    if buf is None or buf.numel() < f.n * 12:
                        stages.pop(f.slot, None)
                        _release_staging(buf)
                        buf = stages[f.slot] = _acquire_staging(f.n * 12)
                    f.stage = buf
                f.feed(copy_stream, nt)
                last[f.slot] = f
            except BaseException as e:      # never leave the consumer waiting on a feed that will not come
                f.fail(e)

    worker = threading.Thread(target=feeder, name="pch-tile-feeder", daemon=True)
    worker.start()
    try:
        it = iter(tiles)

        def prep(tile, slot):
            if tile is None:
                return None
            rec, n, rec_len = tile
            f = _TileFeed(rec, n, rec_len, chunk_size, slice_chunks, pack, raw_every, device, slot)
            jobs.put(f)
            return f

        slot = 0
        cur = prep(next(it, None), slot)
        while cur is not None:
            slot ^= 1
            nxt = prep(next(it, None), slot)     # queued now: gathered / copied while `cur` is processed
            yield _consume(cur, scales, offsets, voxel_size, kw)
            cur = nxt
    finally:
        jobs.put(None)
        worker.join()
        torch.cuda.current_stream(device).synchronize()
        copy_stream.synchronize()
        for buf in stages.values():
            _release_staging(buf)


# ---------------------------------------------------------------------------------------------
# whole-corridor mode: spatial tiles over several GPUs with a DBSCAN halo exchange (tiles.py)
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass
class TiledResult:
    n_points: int                 # input points of this rank
    n_voxels: int
    n_candidates: int             # this rank's candidates
    n_clusters: int               # global
    towers: List[dict]            # the same list on every rank
    labels: Optional[torch.Tensor] = None      # int32 [n_candidates]: global cluster ids
    candidates: Optional[torch.Tensor] = None  # (n_candidates,3) float32 in the common frame
    origin: Optional[np.ndarray] = None        # float32[3] common frame origin
    halo: Optional[dict] = None                # exchange sizes (points sent / received, bytes)


def corridor_axis(azimuth_deg: float):
    """Unit vector (east, north) of a corridor whose azimuth is measured clockwise from north."""
    a = np.radians(float(azimuth_deg))
    return float(np.sin(a)), float(np.cos(a))


def run_pipeline_tiled(tiles: List[dv.DeviceLas], comm, axis, voxel_size: float = 0.1, chunk_size: int = 500000,
                       eps: float = 8.0, min_points: int = 80, keep: bool = False, clusterer=None,
                       ground: str = "percentile", cell: float = 2.0, hag: float = 3.0, per_tile=None, origin=None,
                       **tower_kw) -> TiledResult:
    """This rank's consecutive corridor tiles -> candidates per tile -> ONE DBSCAN over the tiles of all ranks
    (halo exchange with the neighbouring ranks, tiles.tile_dbscan) -> towers from the all-reduced per-cluster
    table (box = AABB rule of test/008.py:302-319).  Equals `dbscan_chunked(chunk=G)` on the concatenation of
    all ranks' candidates, label for label.
    Ground removal stays per tile: ground="percentile" is the reference's height filter with the tile's OWN centroid and
    percentile, i.e. the reference run on that tile's LAS (utils/tower_extraction.py:57-93); ground="grid" is north_star's
    grid min-z / height-above-ground model of the tile (cells from the tile's own minimum).  The kept points of every
    tile are then expressed in ONE float32 frame, so that all ranks compute distances on identical coordinates.
    `per_tile(dl, voxel_result)` is called after each tile's voxel stage
    (the geoid / CRS conversion of configs[4] hooks in here).  `origin` (float32[3]): a frame every rank already knows
    (e.g. the project's LAS header offset plus a nominal height) saves the collective that otherwise agrees on one."""
    from . import tiles as tl
    dv._require_cuda()
    device = tiles[0].device if tiles else torch.device("cuda", torch.cuda.current_device())
    if ground not in ("percentile", "grid"):
        raise ValueError(f"unknown ground mode {ground!r}")
    tr = tl._Trace(comm.rank)
    # phase A — every rank, every tile, no dependency between ranks: voxel stage, and what defines the tile's frame
    stage, n_vox, n_pts = [], 0, 0
    for dl in tiles:
        with nvtx_range("pch.tile.voxel_downsample"):
            vres = dv.voxel_downsample(dl, voxel_size, chunk_size, want=("f32", "z32") if ground == "percentile" else ("f32",))
        if per_tile is not None:
            with nvtx_range("pch.tile.per_tile"):
                per_tile(dl, vres)
        n_vox += vres.count
        n_pts += dl.n
        if vres.count == 0:
            continue
        raw = vres.f32
        if ground == "grid":
            stage.append({"raw": raw, "mm": dv.f32_minmax_dev(raw)})           # bounding box: no read-back yet
        else:
            _, cen_dev, _, _, mask, _ = tw._ground_filter_percentile(raw, want_mask=True, zcol=vres.z32)
            stage.append({"raw": raw, "mask": mask, "cen": cen_dev})
    # the common frame: the centre of the bounding box (grid) / the centroid (percentile) of rank 0's first tile
    mm_all = None
    if stage and ground == "grid":
        mm_all = torch.stack([t["mm"] for t in stage]).cpu().numpy()           # ONE read-back for all tiles of this rank
        mine = ((mm_all[0, :3] + mm_all[0, 3:]) * np.float32(0.5)).astype(np.float32)
    elif stage:
        mine = stage[0]["cen"].cpu().numpy().astype(np.float32)
    else:
        mine = np.full(3, np.nan, dtype=np.float32)
    tr.mark("voxel stage of all tiles")
    if origin is None:
        origin = comm.all_gather_fixed(mine)[0].astype(np.float32)
    else:
        origin = np.asarray(origin, dtype=np.float32).reshape(3)
    tr.mark("gather frame")
    if not np.all(np.isfinite(origin)):
        raise ValueError("rank 0 holds no points: no common frame")
    origin_dev = torch.from_numpy(origin.copy()).to(device)
    # phase B — ground removal relative to the common frame; the counts of all tiles are read back together
    outs, cnts = [], []
    for i, t in enumerate(stage):
        raw = t["raw"]
        if ground == "grid":
            mn, mx = mm_all[i, 0:2] - origin[:2], mm_all[i, 3:5] - origin[:2]   # float32, like the kernels' shift
            nx, ny = dv.grid_shape(mn, mx, cell, raw.shape[0])
            out, cnt, _ = dv.compact_points_grid(raw, origin_dev, mn, nx, ny, cell, hag, sync=False)
            outs.append(out)
            cnts.append(cnt)
        else:
            cand, g, _, _ = dv.compact_points(raw, None, 0.0, origin_dev, keep_mask=t["mask"])
            outs.append(cand)
    if cnts:
        got = torch.cat(cnts).cpu().numpy()
        outs = [o[: int(c)] for o, c in zip(outs, got)]
    P_own = torch.cat(outs).contiguous() if len(outs) > 1 else (outs[0].contiguous() if outs else
                                                                torch.zeros((0, 3), dtype=torch.float32, device=device))
    tr.mark("ground + compaction")
    with nvtx_range("pch.tile_dbscan(halo)"):
        res = tl.tile_dbscan(P_own, axis, eps, min_points, comm, clusterer)
    tr.mark("tile_dbscan")
    stages = tw.TowerStages(None, origin, np.float32("nan"), 3.0, P_own, res.labels, res.n_clusters, res.stats)
    towers = tw.select_towers(stages, box="aabb", want_points=False, **tower_kw)
    tr.mark("towers")
    tr.done(f"run_pipeline_tiled tiles={len(tiles)}")
    halo = {"received": res.halo, "sent": res.sent, "p2p_bytes": 16 * sum(res.sent), "counts": res.counts}
    return TiledResult(n_pts, n_vox, int(P_own.shape[0]), res.n_clusters, towers, res.labels if keep else None,
                       P_own if keep else None, origin, halo)
