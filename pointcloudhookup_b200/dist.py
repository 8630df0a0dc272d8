"""Multi-GPU: one process per GPU, one spatial tile (its own LAS) per rank.

The path shards by tile with NO data-path collective: each rank runs the whole pipeline on its tile
(parity definition, SURVEY.md §8e: the reference run independently on every tile's LAS file).  The only
exchange is the final tower merge: an all-gather of the K' tower records (128 B each) followed by the
reference's own greedy duplicate rule (utils/tower_extraction.py:153-162, 30 m) applied in canonical
(rank, detection order) on every rank, so all ranks hold the same list.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch
import torch.distributed as dist

REC = 20  # center(3) extent(3) rotation(9) height width north_angle label rank


def pack_towers(towers: List[dict], rank: int) -> np.ndarray:
    out = np.zeros((len(towers), REC), dtype=np.float64)
    for i, t in enumerate(towers):
        out[i, 0:3] = t["center"]
        out[i, 3:6] = t["extent"]
        out[i, 6:15] = np.asarray(t["rotation"], dtype=np.float64).reshape(-1)
        out[i, 15], out[i, 16], out[i, 17] = t["height"], t["width"], t["north_angle"]
        out[i, 18], out[i, 19] = t.get("label", -1), rank
    return out


def unpack_towers(arr: np.ndarray) -> List[dict]:
    return [{"center": r[0:3].copy(), "extent": r[3:6].copy(), "rotation": r[6:15].reshape(3, 3).copy(),
             "height": float(r[15]), "width": float(r[16]), "north_angle": float(r[17]), "label": int(r[18]),
             "rank": int(r[19])} for r in arr]


def dedup_towers(towers: List[dict], duplicate_threshold: float = 30.0) -> List[dict]:
    kept, centres = [], []
    for t in towers:
        if any(np.linalg.norm(t["center"] - c) < duplicate_threshold for c in centres):
            continue
        kept.append(t)
        centres.append(t["center"])
    return kept


FAST_CAP = 255   # towers per rank that travel in the single fixed-size collective


def _gather_rows(block: torch.Tensor, world: int) -> torch.Tensor:
    """all-gather of one equally shaped block per rank -> (world, *block.shape)."""
    out = torch.empty((world,) + tuple(block.shape), dtype=block.dtype, device=block.device)
    try:
        dist.all_gather_into_tensor(out, block)
    except (RuntimeError, NotImplementedError, AttributeError):   # backend without the flat variant
        parts = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(parts, block)
        out = torch.stack(parts)
    return out


def merge_towers(towers: List[dict], duplicate_threshold: float = 30.0, device=None) -> List[dict]:
    """All ranks contribute their tile's towers; every rank returns the same merged list.

    One collective and one device->host read per call: every rank sends a fixed (FAST_CAP+1, REC) block whose
    first row carries its tower count (a corridor tile has a few dozen towers at most).  Only if some rank has
    more than FAST_CAP towers — announced in that same block, so all ranks agree — a second, max-sized gather
    follows."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dedup_towers(unpack_towers(pack_towers(towers, 0)), duplicate_threshold)
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = pack_towers(towers, rank)
    k = mine.shape[0]
    block = np.zeros((FAST_CAP + 1, REC), dtype=np.float64)
    block[0, 0] = k
    if k <= FAST_CAP:
        block[1: k + 1] = mine
    host = _gather_rows(torch.from_numpy(block).to(device), world).cpu().numpy()
    counts = [int(host[r, 0, 0]) for r in range(world)]
    cmax = max(counts)
    if cmax == 0:
        return []
    if cmax > FAST_CAP:
        padded = np.zeros((cmax, REC), dtype=np.float64)
        padded[:k] = mine
        big = _gather_rows(torch.from_numpy(padded).to(device), world).cpu().numpy()
        rows = [big[r, : counts[r]] for r in range(world)]
    else:
        rows = [host[r, 1: counts[r] + 1] for r in range(world)]
    allt = []
    for r in range(world):
        allt += unpack_towers(rows[r])
    return dedup_towers(allt, duplicate_threshold)


def tile_for_rank(rank: int, towers_per_tile: int):
    """Along-axis origin of a rank's tile in the synthetic corridor (tiles abut along the axis)."""
    from . import synth
    return rank * towers_per_tile * synth.SPAN
