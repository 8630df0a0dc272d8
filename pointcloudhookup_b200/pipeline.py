"""The fused hot path: raw LAS records in HBM -> voxel downsample -> (LAS re-quantisation, float32
read-back) -> centroid / height filter -> chunked DBSCAN -> per-cluster reduction -> tower list.

This is pyGUI_towers_test.py:344-368 (downsample_and_extract) without the LAS file between the two
stages: the intermediate ``output/point_2.las`` is reproduced arithmetically (means -> int32 lattice
-> float64 -> float32) inside the reduce kernel, so results are identical to the two-step run.
"""
from __future__ import annotations

import dataclasses
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
    vres = dv.voxel_downsample(dl, voxel_size, chunk_size, want=("f32",))
    if vres.count == 0:
        return PipelineResult(dl.n, 0, 0, 0, [])
    stages = tw.run_stages(vres.f32, eps, min_points, ground)
    towers = tw.select_towers(stages, box=box, want_points=want_points, **tower_kw)
    return PipelineResult(dl.n, vres.count, int(stages.filtered.shape[0]), stages.n_clusters, towers,
                          stages if keep_stages else None, vres.plan, stages.db_plan)


def run_pipeline_from_host(host_records: torch.Tensor, n: int, rec_len: int, scales, offsets, voxel_size: float = 0.1,
                           chunk_size: int = 500000, slice_chunks: int = 20, device=None, **kw) -> PipelineResult:
    """End-to-end entry for HOST record buffers (pinned uint8 tensor): the H2D copy is cut into slices of
    whole chunks on a copy stream, and the voxel stage of slice i runs while slice i+1 is still in flight
    (chunks are independent, so the result is identical to one big call).  The tower stage follows on the
    concatenated float32 cloud."""
    dv._require_cuda()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    cs = max(1, min(int(chunk_size), max(n, 1)))
    per_slice = cs * max(1, int(slice_chunks))
    if cs % 8:  # slice starts must stay 16-byte aligned (rec_len is even for every LAS format we stream)
        per_slice = n
    host = host_records.view(torch.uint8).reshape(-1)[: n * rec_len]
    if not host.is_pinned():
        host = host.pin_memory()
    dev = torch.empty(dv.padded_bytes(n, rec_len), dtype=torch.uint8, device=device)
    dev[n * rec_len:].zero_()
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    copy_stream.wait_stream(main)
    events, bounds = [], []
    with torch.cuda.stream(copy_stream):
        for lo in range(0, n, per_slice):
            hi = min(lo + per_slice, n)
            dev[lo * rec_len: hi * rec_len].copy_(host[lo * rec_len: hi * rec_len], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            events.append(ev)
            bounds.append((lo, hi))
    parts, total = [], 0
    for ev, (lo, hi) in zip(events, bounds):
        main.wait_event(ev)
        view = dv.DeviceLas(dev[lo * rec_len:], hi - lo, rec_len, np.asarray(scales, dtype=np.float64),
                            np.asarray(offsets, dtype=np.float64))
        v = dv.voxel_downsample(view, voxel_size, cs, want=("f32",))
        parts.append(v.f32)
        total += v.count
    if total == 0:
        return PipelineResult(n, 0, 0, 0, [])
    f32 = parts[0] if len(parts) == 1 else torch.cat(parts)
    del parts
    eps = kw.pop("eps", 8.0)
    min_points = kw.pop("min_points", 80)
    ground = kw.pop("ground", "percentile")
    box = kw.pop("box", "aabb")
    want_points = kw.pop("want_points", False)
    stages = tw.run_stages(f32, eps, min_points, ground)
    towers = tw.select_towers(stages, box=box, want_points=want_points, **kw)
    return PipelineResult(n, total, int(stages.filtered.shape[0]), stages.n_clusters, towers, None, None, stages.db_plan)
