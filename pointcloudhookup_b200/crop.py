"""Per-tower box crop and preview subsample on the device.

* test/kuangxuan.py:58-79: for each detected tower, an asymmetric axis-aligned box around its centre and
  ``tower_points = points[mask]`` with six inclusive float64 compares.  `crop_towers` does this for ALL
  towers in one streaming pass over the raw records (pch_las_box_crop), sorts the emitted
  (tower, point index) words and gathers, so every tower's array equals the reference's ``points[mask]``
  (same points, same order) without N x #towers mask passes.
* pyGUI_towers_test.py:174-177 / ui/vtk_widget.py:115-118: ``np.random.choice(len(xyz), k, replace=False)``
  preview subsample.  `preview_subsample` draws k distinct points with a keyed bijection of [0, n)
  evaluated on the device (or evenly spaced for seed=None); the reference's draw comes from numpy's unseeded
  global RNG, so only "k distinct points of the cloud" is reproducible, not the draw itself.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _native
from . import device as dv
from ._native import check, d3


def kuangxuan_bounds(tower: dict) -> np.ndarray:
    """xmin,ymin,zmin,xmax,ymax,zmax of test/kuangxuan.py:63-71 for a tower dict with x,y,z,width,height."""
    w, h = tower["width"], tower["height"]
    cx, cy, cz = tower["x"], tower["y"], tower["z"]
    return np.array([cx - w / 1, cy - w / 2, cz - h / 1, cx + w / 0.6, cy + w / 1, cz + h * 2], dtype=np.float64)


def crop_boxes(dl: dv.DeviceLas, boxes) -> List[torch.Tensor]:
    """[points[mask_b] for b in boxes] as (n_b,3) float64 device tensors; boxes: (T,6) xmin,ymin,zmin,xmax,ymax,zmax."""
    dv._require_cuda()
    lib = _native.lib()
    boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64).reshape(-1, 6))
    T = boxes.shape[0]
    dev = dl.device
    empty = lambda: torch.zeros((0, 3), dtype=torch.float64, device=dev)
    if T == 0:
        return []
    if dl.n == 0:
        return [empty() for _ in range(T)]
    st = dv._stream()
    sc, of = d3(dl.scales), d3(dl.offsets)
    boxes_dev = torch.from_numpy(boxes).to(dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    cap = max(1 << 16, dl.n // 16)
    while True:
        words = torch.empty(cap, dtype=torch.int64, device=dev)
        check(lib.pch_las_box_crop(dl.rec.data_ptr(), dl.n, dl.rec_len, sc, of, boxes_dev.data_ptr(), T,
                                   words.data_ptr(), cap, total.data_ptr(), st), "pch_las_box_crop")
        m = int(total.item())
        if m <= cap:
            break
        cap = m
    if m == 0:
        return [empty() for _ in range(T)]
    words = words[:m]
    bits = 32 + max(1, int(T - 1).bit_length())
    sw = dv.sort_u64_segmented(words, m, 0, bits)
    bounds = torch.empty(T + 1, dtype=torch.int64, device=dev)
    check(lib.pch_word_bounds(sw.data_ptr(), m, T, bounds.data_ptr(), st), "pch_word_bounds")
    pts = torch.empty((m, 3), dtype=torch.float64, device=dev)
    check(lib.pch_las_gather_f64(dl.rec.data_ptr(), dl.n, dl.rec_len, sc, of, sw.data_ptr(), m, pts.data_ptr(), st),
          "pch_las_gather_f64")
    b = bounds.cpu().numpy()
    return [pts[int(b[k]): int(b[k + 1])] for k in range(T)]


def crop_towers(dl: dv.DeviceLas, tower_data: Sequence[dict]) -> List[torch.Tensor]:
    """test/kuangxuan.py:58-79 for every tower of `tower_data` (dicts with x,y,z,width,height)."""
    return crop_boxes(dl, np.stack([kuangxuan_bounds(t) for t in tower_data]) if len(tower_data) else np.zeros((0, 6)))


def sample_indices(n: int, k: int, seed: Optional[int], device) -> torch.Tensor:
    """k distinct indices of [0, n) (int64 device tensor): evenly spaced for seed None/0, else the keyed bijection."""
    dv._require_cuda()
    k = min(int(k), int(n))
    words = torch.empty(k, dtype=torch.int64, device=device)
    check(_native.lib().pch_sample_indices(int(n), k, int(seed or 0) & 0xFFFFFFFFFFFFFFFF, words.data_ptr(),
                                           dv._stream()), "pch_sample_indices")
    return words


def preview_subsample(dl: dv.DeviceLas, k: int = 200000, seed: Optional[int] = None) -> torch.Tensor:
    """`xyz[np.random.choice(len(xyz), k, replace=False)]` when len(xyz) > k, else the whole cloud
    (pyGUI_towers_test.py:174-179): (min(k,n),3) float64 device tensor."""
    if dl.n <= k:
        return dv.decode_xyz(dl, torch.float64)
    words = sample_indices(dl.n, k, seed, dl.device)
    out = torch.empty((k, 3), dtype=torch.float64, device=dl.device)
    check(_native.lib().pch_las_gather_f64(dl.rec.data_ptr(), dl.n, dl.rec_len, d3(dl.scales), d3(dl.offsets),
                                           words.data_ptr(), k, out.data_ptr(), dv._stream()), "pch_las_gather_f64")
    return out
