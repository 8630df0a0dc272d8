"""Tiled RANSAC ground removal on the device.

Reference: test/main_ground.py:77-115 `remove_ground_tiled_ransac(points, tile_size=10.0, **kwargs)` with
`remove_ground_ransac(points, distance_threshold=0.1, max_iterations=1000)` (:8-32) per tile — same names,
same defaults, same return order `(non_ground_points, ground_points)`, same row order (tile by tile in the
order of the two loops, `points[tile_mask]` order inside a tile), and the same omissions: points at or beyond
the last np.arange edge and tiles with fewer than 10 points are in neither array.

The reference draws its samples from an unseeded generator, so two runs of it differ; here the draws come
from a counter-based generator keyed by (seed, tile, trial), so a run is reproducible on any GPU, and a test
can hand in the very triples scikit-learn draws for a given `random_state` (`triples=`) and obtain
scikit-learn's own inlier masks.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _native
from . import device as dv
from ._native import check

MIN_TILE_POINTS = 10                                     # test/main_ground.py:101
TILE_DTYPE = np.dtype([("n_points", "<i4"), ("n_trials", "<i4"), ("n_inliers", "<i4"), ("status", "<i4"),
                       ("anchor", "<f8", 3), ("slope", "<f8", 2), ("score", "<f8")])
assert TILE_DTYPE.itemsize == 64


def arange_edges3(lo: float, hi: float, step: float):
    """(len(np.arange(lo, hi, step)), [edges[0], edges[1], edges[1] - edges[0]]): what pch_ransac_tile_words needs to
    evaluate np.arange's own fill rule edges[i] = edges[0] + i * (edges[1] - edges[0])."""
    e = np.arange(lo, hi, step)
    if len(e) < 2:
        return len(e), None
    return len(e), (C.c_double * 3)(float(e[0]), float(e[1]), float(e[1] - e[0]))


def remove_ground_tiled_ransac(points, tile_size: float = 10.0, distance_threshold: float = 0.1, max_iterations: int = 1000, *,
                               seed: int = 0, triples=None, stop_probability: float = 0.99, return_tiles: bool = False):
    """points: (n,3) float64, numpy (results come back as numpy) or a CUDA tensor (results stay on the device).
    triples: optional (n_tiles, max_iterations, 3) int32 sample rows per tile and trial (tile = i*(len(y_edges)-1)+j).
    Returns (non_ground_points, ground_points) [, per-tile records (TILE_DTYPE)]."""
    dv._require_cuda()
    lib = _native.lib()
    as_numpy = not isinstance(points, torch.Tensor)
    if as_numpy:
        P = torch.from_numpy(np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))).cuda()
    else:
        P = points.to(torch.float64).reshape(-1, 3).contiguous()
    dev = P.device
    n = int(P.shape[0])
    if n == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")   # np.min(points[:, :2])
    st = dv._stream()

    def result(non_ground, ground, tiles):
        if as_numpy:
            non_ground, ground = non_ground.cpu().numpy(), ground.cpu().numpy()
        return (non_ground, ground, tiles) if return_tiles else (non_ground, ground)

    empty = lambda: torch.zeros((0, 3), dtype=torch.float64, device=dev)
    mm_dev = torch.empty(4, dtype=torch.float64, device=dev)
    scratch = torch.empty(4, dtype=torch.int64, device=dev)
    check(lib.pch_xy_minmax_f64(P.data_ptr(), n, mm_dev.data_ptr(), scratch.data_ptr(), st), "pch_xy_minmax_f64")
    mm = mm_dev.cpu().numpy()
    nex, ex3 = arange_edges3(mm[0], mm[2], tile_size)
    ney, ey3 = arange_edges3(mm[1], mm[3], tile_size)
    if nex < 2 or ney < 2:
        return result(empty(), empty(), np.zeros(0, dtype=TILE_DTYPE))
    n_tiles = (nex - 1) * (ney - 1)
    words = torch.empty(n, dtype=torch.int64, device=dev)
    check(lib.pch_ransac_tile_words(P.data_ptr(), n, ex3, nex, ey3, ney, words.data_ptr(), st), "pch_ransac_tile_words")
    # the words are in index order already: a stable sort on the tile bits alone keeps `points[tile_mask]` order
    sw = dv.sort_u64_segmented(words, n, 32, 32 + max(1, int(n_tiles).bit_length()))
    bounds = torch.empty(n_tiles + 2, dtype=torch.int64, device=dev)
    check(lib.pch_word_bounds(sw.data_ptr(), n, n_tiles + 1, bounds.data_ptr(), st), "pch_word_bounds")
    m = int(bounds[n_tiles].item())                        # rows that lie in some tile
    Q = torch.empty((max(m, 1), 3), dtype=torch.float64, device=dev)
    check(lib.pch_gather_rows_f64(P.data_ptr(), sw.data_ptr(), m, Q.data_ptr(), st), "pch_gather_rows_f64")
    tri_ptr = None
    if triples is not None:
        tri = np.ascontiguousarray(np.asarray(triples, dtype=np.int32))
        if tri.shape != (n_tiles, max_iterations, 3):
            raise ValueError(f"triples must have shape {(n_tiles, max_iterations, 3)}, got {tri.shape}")
        tri_dev = torch.from_numpy(tri).to(dev)
        tri_ptr = tri_dev.data_ptr()
    flags = torch.empty(max(m, 1), dtype=torch.uint8, device=dev)
    rec = torch.empty(n_tiles * TILE_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    check(lib.pch_ransac_tiles(Q.data_ptr(), bounds.data_ptr(), n_tiles, float(distance_threshold), int(max_iterations),
                               float(stop_probability), int(seed) & ((1 << 64) - 1), tri_ptr, MIN_TILE_POINTS,
                               flags.data_ptr(), rec.data_ptr(), st), "pch_ransac_tiles")
    tiles = rec.cpu().numpy().view(TILE_DTYPE).copy()
    if np.any(tiles["status"] == 2):
        raise ValueError("RANSAC could not find a valid consensus set. All `max_trials` iterations were skipped because "
                         "each randomly chosen sub-sample failed the passing criteria.")
    ok = tiles["status"] == 0
    g_cnt = np.where(ok, tiles["n_inliers"], 0).astype(np.int64)
    o_cnt = np.where(ok, tiles["n_points"] - tiles["n_inliers"], 0).astype(np.int64)
    g_off = np.concatenate([[0], np.cumsum(g_cnt)]).astype(np.int64)
    o_off = np.concatenate([[0], np.cumsum(o_cnt)]).astype(np.int64)
    ground = torch.empty((max(int(g_off[-1]), 1), 3), dtype=torch.float64, device=dev)
    other = torch.empty((max(int(o_off[-1]), 1), 3), dtype=torch.float64, device=dev)
    offs = torch.from_numpy(np.stack([g_off[:-1], o_off[:-1]])).to(dev)
    check(lib.pch_ransac_split(Q.data_ptr(), flags.data_ptr(), bounds.data_ptr(), n_tiles, offs[0].data_ptr(), offs[1].data_ptr(),
                               ground.data_ptr(), other.data_ptr(), st), "pch_ransac_split")
    return result(other[: int(o_off[-1])], ground[: int(g_off[-1])], tiles)
