"""The fused hot path: raw LAS records in HBM -> voxel downsample -> (LAS re-quantisation, float32
read-back) -> centroid / height filter -> chunked DBSCAN -> per-cluster reduction -> tower list.

This is pyGUI_towers_test.py:344-368 (downsample_and_extract) without the LAS file between the two
stages: the intermediate ``output/point_2.las`` is reproduced arithmetically (means -> int32 lattice
-> float64 -> float32) inside the reduce kernel, so results are identical to the two-step run.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

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
