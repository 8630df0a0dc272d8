"""The fused hot path: raw LAS records in HBM -> voxel downsample -> (LAS re-quantisation, float32
read-back) -> centroid / height filter -> chunked DBSCAN -> per-cluster reduction -> tower list.

This is pyGUI_towers_test.py:344-368 (downsample_and_extract) without the LAS file between the two
stages: the intermediate ``output/point_2.las`` is reproduced arithmetically (means -> int32 lattice
-> float64 -> float32) inside the reduce kernel, so results are identical to the two-step run.
"""
from __future__ import annotations

import dataclasses
import os
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


def run_pipeline(dl: dv.DeviceLas, voxel_size: float = 0.1, chunk_size: int = 500000, eps: float = 8.0,
                 min_points: int = 80, ground: str = "percentile", box: str = "aabb", want_points: bool = False,
                 keep_stages: bool = False, **tower_kw) -> PipelineResult:
    vres = dv.voxel_downsample(dl, voxel_size, chunk_size, want=("f32", "z32") if ground == "percentile" else ("f32",))
    if vres.count == 0:
        return PipelineResult(dl.n, 0, 0, 0, [])
    stages = tw.run_stages(vres.f32, eps, min_points, ground, zcol=vres.z32)
    towers = tw.select_towers(stages, box=box, want_points=want_points, **tower_kw)
    return PipelineResult(dl.n, vres.count, int(stages.filtered.shape[0]), stages.n_clusters, towers,
                          stages if keep_stages else None, vres.plan, stages.db_plan)


_STAGING = {}   # (device index, nbytes) -> pinned staging tensor, reused across calls (pinning costs ~0.3 s/GB)


def _staging(nbytes: int) -> torch.Tensor:
    key = int(nbytes)
    buf = _STAGING.get(key)
    if buf is None:
        _STAGING.clear()      # one live staging buffer: a new size replaces the old one
        buf = torch.empty(key, dtype=torch.uint8, pin_memory=True)
        _STAGING[key] = buf
    return buf


host_threads = dv.host_threads


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
                 and the host cores both carry part of the stream."""
    dv._require_cuda()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    cs = max(1, min(int(chunk_size), max(n, 1)))
    per_slice = cs * max(1, int(slice_chunks))
    if pack not in ("none", "xyz"):
        raise ValueError(f"unknown pack mode {pack!r}")
    if (cs * 12) % 16 or (cs * rec_len) % 16:  # slice starts must stay 16-byte aligned in both layouts
        per_slice = n
    if isinstance(host_records, torch.Tensor):
        host = host_records.view(torch.uint8).reshape(-1)[: n * rec_len]
    else:
        host = torch.from_numpy(np.asarray(host_records).view(np.uint8).reshape(-1)[: n * rec_len])
    if pack == "none" and not host.is_pinned():
        host = host.pin_memory()
    if not host.is_pinned():
        raw_every = 0
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    copy_stream.wait_stream(main)
    bounds = [(lo, min(lo + per_slice, n)) for lo in range(0, n, per_slice)]
    # record length of each slice on the device: whole records, or the gathered 12-byte stream
    lens = [rec_len if (pack == "none" or (raw_every > 0 and i % raw_every == raw_every - 1)) else 12
            for i in range(len(bounds))]
    bufs = []
    for (lo, hi), ln in zip(bounds, lens):
        b = torch.empty(dv.padded_bytes(hi - lo, ln), dtype=torch.uint8, device=device)
        b[(hi - lo) * ln:].zero_()
        bufs.append(b)
    copy_stream.wait_stream(main)
    events = [torch.cuda.Event() for _ in bounds]
    ready = [threading.Event() for _ in bounds]
    failure = []

    def feed():
        # runs on a worker thread for pack="xyz" (ctypes releases the GIL during the gather) and inline otherwise
        try:
            torch.cuda.set_device(device)
            stage = _staging(n * 12) if pack == "xyz" else None
            lib = dv._native.lib()
            nt = threads or host_threads()
            for i, (lo, hi) in enumerate(bounds):
                if lens[i] == 12 and pack == "xyz":
                    dv.check(lib.pch_host_pack_xyz(host.data_ptr() + lo * rec_len, hi - lo, rec_len,
                                                   stage.data_ptr() + lo * 12, nt), "pch_host_pack_xyz")
                    src = stage[lo * 12: hi * 12]
                else:
                    src = host[lo * rec_len: hi * rec_len]
                with torch.cuda.stream(copy_stream):
                    bufs[i][: src.numel()].copy_(src, non_blocking=True)
                    events[i].record(copy_stream)
                ready[i].set()
        except BaseException as e:     # surface the failure on the calling thread
            failure.append(e)
            for r in ready:
                r.set()

    worker = None
    if pack == "xyz":
        worker = threading.Thread(target=feed, name="pch-host-pack", daemon=True)
        worker.start()
    else:
        feed()
    ground = kw.pop("ground", "percentile")
    sink = dv.VoxelSink(n, device, want_z=(ground == "percentile"))
    for i, (lo, hi) in enumerate(bounds):
        ready[i].wait()
        if failure:
            raise failure[0]
        main.wait_event(events[i])
        view = dv.DeviceLas(bufs[i], hi - lo, lens[i], np.asarray(scales, dtype=np.float64),
                            np.asarray(offsets, dtype=np.float64))
        dv.voxel_downsample(view, voxel_size, cs, want=(), sink=sink)
    if worker is not None:
        worker.join()
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
