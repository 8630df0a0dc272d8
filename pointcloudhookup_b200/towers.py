"""Tower extraction pipeline (utils/tower_extraction.py:57-218) on the device.

Per-point work (float32 cast, centroid, shift, order statistics, height filter, chunked DBSCAN,
per-cluster reduction) is CUDA (device.py -> libpch_b200.so).  What stays on the host is the
reference's O(#clusters) control flow, reproduced literally: numpy's percentile index/lerp
arithmetic on two scalars, ``set(all_labels) - {-1}`` iteration order, the size filter, the 30 m
duplicate check, the north angle, and (box="obb") the trimesh-style oriented box of each cluster.
"""
from __future__ import annotations

import dataclasses
import os
import threading
from typing import Callable, List, Optional

import numpy as np
import torch

from . import device as dv

DBSCAN_CHUNK = 50000  # utils/tower_extraction.py:96


# ---------------------------------------------------------------------------------------------
# numpy.percentile(float32 array, q) index / interpolation arithmetic, restated with the same
# public numpy operations numpy 2.x executes (lib/_function_base_impl.py: percentile, _quantile,
# _get_indexes, _get_gamma, _lerp) so dtype promotion and rounding are identical.
# ---------------------------------------------------------------------------------------------
def percentile_ranks_f32(n: int, q: float):
    """(rank_prev, rank_next, gamma) of np.percentile(a_f32, q) for len(a) == n."""
    qa = np.true_divide(q, np.float32(100), out=...)
    vi = np.asanyarray((n - 1) * qa)
    prev = np.asanyarray(np.floor(vi))
    nxt = np.asanyarray(prev + 1)
    if (vi >= n - 1).any():
        prev[...] = -1
        nxt[...] = -1
    if (vi < 0).any():
        prev[...] = 0
        nxt[...] = 0
    prev_i = prev.astype(np.intp)
    next_i = nxt.astype(np.intp)
    gamma = np.asanyarray(vi - prev_i)
    gamma = np.asanyarray(gamma, dtype=vi.dtype)
    return int(prev_i) % n, int(next_i) % n, gamma


def percentile_lerp_f32(a: np.float32, b: np.float32, gamma):
    a = np.asanyarray(a, dtype=np.float32)
    b = np.asanyarray(b, dtype=np.float32)
    t = gamma
    diff = np.subtract(b, a)
    lerp = np.asanyarray(np.add(a, diff * t))
    np.subtract(b, diff * (1 - t), out=lerp, where=t >= 0.5, casting="unsafe", dtype=type(lerp.dtype))
    return lerp[()]


@dataclasses.dataclass
class TowerStages:
    """Device-resident intermediates of one extract_towers run (kept for parity tests)."""
    raw: torch.Tensor               # (M,3) float32 raw_points
    centroid: np.ndarray            # float32[3]
    base: np.float32
    offset_used: float
    filtered: torch.Tensor          # (G,3) float32 = points[z > thr]
    labels: torch.Tensor            # (G,) int32 all_labels
    n_clusters: int
    stats: np.ndarray               # per-label count / AABB / sums (host)
    mask: Optional[torch.Tensor] = None
    db_plan: Optional[dict] = None


_SIDE_STREAMS = {}


def _side_stream(device) -> torch.cuda.Stream:
    key = (threading.get_ident(), torch.device(device).index)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def ground_filter_percentile(raw: torch.Tensor, pct: float = 25, offset: float = 3.0, min_keep: int = 1000,
                             fallback_offset: float = 1.0, want_mask: bool = False,
                             zcol: Optional[torch.Tensor] = None):
    """Stages A+B: (filtered, centroid (device), base, offset used, mask)."""
    return _ground_filter_percentile(raw, pct, offset, min_keep, fallback_offset, want_mask, zcol)[:5]


def _ground_filter_percentile(raw: torch.Tensor, pct: float = 25, offset: float = 3.0, min_keep: int = 1000,
                              fallback_offset: float = 1.0, want_mask: bool = False,
                              zcol: Optional[torch.Tensor] = None):
    """Stages A+B: centroid, shift, percentile threshold, compaction; also returns the host centroid.

    `points = raw - centroid` is a float32 subtract per element and x -> float32(x - c) is monotone, so the
    order statistics of the shifted z column are the shifted order statistics of the RAW z column: the radix
    select runs on the raw column (`zcol`, emitted by the voxel reduce, or extracted here) on a side stream
    while the bit-exact sequential-sum centroid is still being evaluated, and the shifted column is never
    materialised (the compaction derives the keep flag from the cloud itself)."""
    m = raw.shape[0]
    main = torch.cuda.current_stream(raw.device)
    side = _side_stream(raw.device)
    side.wait_stream(main)
    r0, r1, gamma = percentile_ranks_f32(m, pct)
    with torch.cuda.stream(side):
        zraw = zcol if zcol is not None else dv.f32_column(raw, 2)
        two_dev = dv.select_f32(zraw, r0, r1)
    if os.environ.get("PCH_TRACE"):
        cen_dev, _, st = dv.f32_centroid(raw, want_stats=True)
        print(f"[pch] centroid tiles via maps / via real adds per column: {st.tolist()}", flush=True)
    else:
        cen_dev, _ = dv.f32_centroid(raw)
    main.wait_stream(side)
    host = torch.cat([cen_dev, two_dev]).cpu().numpy()          # one D2H: centroid + the two raw order statistics
    two = host[3:5] - host[2]                                   # float32 - float32 = the shifted order statistics
    base = percentile_lerp_f32(two[0], two[1], gamma)
    thr = base + offset                      # np.float32 + python float -> float32 (NEP 50)
    filtered, g, _, mask = dv.compact_points(raw, None, float(thr), cen_dev, want_mask=want_mask)
    used = offset
    if g < min_keep:
        thr = base + fallback_offset
        filtered, g, _, mask = dv.compact_points(raw, None, float(thr), cen_dev, want_mask=want_mask)
        used = fallback_offset
    return filtered, cen_dev, np.float32(base), used, mask, host[:3].copy()


def ground_filter_grid(raw: torch.Tensor, cell: float = 2.0, hag: float = 3.0, want_mask: bool = False):
    """north_star grid min-z mode: keep points more than `hag` above their cell's lowest point."""
    return _ground_filter_grid(raw, cell, hag, want_mask)[:5]


def _ground_filter_grid(raw: torch.Tensor, cell: float = 2.0, hag: float = 3.0, want_mask: bool = False):
    """Centroid (main stream) and the raw cloud's bounding box (side stream) are read back together; the grid origin
    is the shifted minimum float32(min_raw - c) (x -> float32(x - c) is monotone, so it equals the minimum of the
    shifted cloud the oracle takes), and both grid kernels shift on the fly: no shifted copy of the cloud exists."""
    main = torch.cuda.current_stream(raw.device)
    side = _side_stream(raw.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        mm_dev = dv.f32_minmax_dev(raw)
    cen_dev, _ = dv.f32_centroid(raw)
    main.wait_stream(side)
    host = torch.cat([cen_dev, mm_dev]).cpu().numpy()          # one D2H: centroid + raw bounding box
    cen = host[:3].copy()
    mn = host[3:5] - cen[:2]                                    # float32 - float32, like the kernels' __fsub_rn
    mx = host[6:8] - cen[:2]
    nx, ny = dv.grid_shape(mn, mx, cell, raw.shape[0])
    filtered, g, mask = dv.compact_points_grid(raw, cen_dev, mn, nx, ny, cell, hag, want_mask=want_mask)
    return filtered, cen_dev, np.float32("nan"), hag, mask, cen


def run_stages(raw: torch.Tensor, eps: float = 8.0, min_points: int = 80, ground: str = "percentile",
               want_mask: bool = False, zcol: Optional[torch.Tensor] = None, **ground_kw) -> TowerStages:
    if ground == "percentile":
        filtered, cen_dev, base, used, mask, cen = _ground_filter_percentile(raw, want_mask=want_mask, zcol=zcol,
                                                                             **ground_kw)
    elif ground == "grid":
        filtered, cen_dev, base, used, mask, cen = _ground_filter_grid(raw, want_mask=want_mask, **ground_kw)
    else:
        raise ValueError(f"unknown ground mode {ground!r}")
    db = dv.dbscan_chunked(filtered, eps, min_points, DBSCAN_CHUNK)
    return TowerStages(raw, cen, base, used, filtered, db.labels, db.n_clusters, db.stats, mask, db.plan)


# ---------------------------------------------------------------------------------------------
# host epilogue: O(#clusters)
# ---------------------------------------------------------------------------------------------
def canonical_box_axes(ax0, ax2):
    """Sign convention of the device boxes: the long axis points into x >= 0 (first non-zero of x, y, z), the face
    normal into z >= 0 (first non-zero of z, y, x), axis 1 completes a right-handed frame.  (The box itself does not
    depend on the signs; the reference's north angle, read from the first column, does.)"""
    def positive(v, order):
        v = np.asarray(v, dtype=np.float64)
        for k in order:
            if abs(v[k]) > 1e-12:
                return v if v[k] > 0 else -v
        return v
    a0 = positive(ax0, (0, 1, 2))
    a2 = positive(ax2, (2, 1, 0))
    return a0, np.cross(a2, a0), a2


def north_angle_of(rotation: np.ndarray) -> float:
    x_axis = rotation[:, 0]
    h = np.array([x_axis[0], x_axis[1], 0])
    if np.linalg.norm(h) > 1e-6:
        h = h / np.linalg.norm(h)
    else:
        h = np.array([1, 0, 0])
    a = np.degrees(np.arctan2(h[1], h[0]))
    if a < 0:
        a += 360
    return (90 - a) % 360


def label_iteration_order(n_clusters: int, present: np.ndarray) -> List[int]:
    """Iteration order of ``set(all_labels) - {-1}`` (utils/tower_extraction.py:125,131): CPython
    hash-table order of np.int32 keys.  Built literally from the labels that occur."""
    # python ints hash like np.int32 (hash(v) == v), so the table layout and iteration order are the same as for
    # the reference's set of numpy scalars, at a fraction of the cost
    s = set(np.asarray(present, dtype=np.int32).tolist()) - {-1}
    return list(s)


_ORDER_CACHE = {}


def _label_order_array(n_clusters: int, present: np.ndarray) -> np.ndarray:
    """label_iteration_order as an int64 array; memoised for the usual case "every label 0..K-1 occurs" (the set is
    still built literally once per K, so a CPython whose table layout differs is followed, not assumed)."""
    full = len(present) == n_clusters
    if full and n_clusters in _ORDER_CACHE:
        return _ORDER_CACHE[n_clusters]
    arr = np.asarray(label_iteration_order(n_clusters, present), dtype=np.int64)
    if full:
        if len(_ORDER_CACHE) > 64:
            _ORDER_CACHE.clear()
        _ORDER_CACHE[n_clusters] = arr
    return arr


def merge_adjacent_clusters(stats: np.ndarray, n_clusters: int, merge_threshold: float = 6.0):
    """The cluster post-processing of test/tttt.py:93-175 on the per-cluster reduction instead of on the points:
    clusters whose centres lie within `merge_threshold` of each other (inclusive, like KDTree.query_radius) are
    joined transitively.  Returns (component id per original label [K], merged stats [C]) where component ids are
    ranked by their first member in the reference's set() iteration order — the reference's new labels are
    max(label)+1+id — and the merged rows carry count = sum, AABB = union, coordinate sums = sum, i.e. exactly what
    re-reducing the relabelled points would give.

    Centres are float64 sum/count here and float32 np.mean (a sequential float32 sum) in the reference: they agree
    to ~1e-3 m on 50 000-point clusters, so only a pair whose distance is within that of the threshold can differ
    (tolerance-level parity; the component structure is otherwise identical)."""
    K = int(n_clusters)
    if K == 0:
        return np.zeros(0, dtype=np.int64), stats[:0].copy()
    present = np.nonzero(stats["count"][:K] > 0)[0].astype(np.int32)
    order = _label_order_array(K, present)                     # valid_labels in the reference's iteration order
    centres = stats["sum"][order] / stats["count"][order][:, None].astype(np.float64)
    n = len(order)
    parent = np.arange(n)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    # all pairs within the threshold: sort along x and sweep (K is a few thousand at most)
    xs = np.argsort(centres[:, 0], kind="stable")
    cx = centres[xs, 0]
    hi = np.searchsorted(cx, cx + merge_threshold, side="right")
    thr2 = float(merge_threshold) ** 2
    for a in range(n):
        b = np.arange(a + 1, hi[a])
        if b.size == 0:
            continue
        d = centres[xs[b]] - centres[xs[a]]
        close = b[np.einsum("ij,ij->i", d, d) <= thr2]
        for j in close:
            ra, rb = find(xs[a]), find(xs[j])
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
    roots = np.array([find(i) for i in range(n)])
    # component id = rank of the component's first member (index order = the reference's `for i in range(len(...))`)
    first_seen, comp_of_root = {}, np.empty(n, dtype=np.int64)
    for i in range(n):
        r = roots[i]
        if r not in first_seen:
            first_seen[r] = len(first_seen)
        comp_of_root[i] = first_seen[r]
    C = len(first_seen)
    comp = np.full(K, -1, dtype=np.int64)
    comp[order] = comp_of_root
    merged = np.zeros(C, dtype=stats.dtype)
    merged["min"] = np.inf
    merged["max"] = -np.inf
    np.add.at(merged["count"], comp_of_root, stats["count"][order])
    np.add.at(merged["sum"], comp_of_root, stats["sum"][order])
    np.minimum.at(merged["min"], comp_of_root, stats["min"][order])
    np.maximum.at(merged["max"], comp_of_root, stats["max"][order])
    return comp, merged


def select_towers(stages: TowerStages, aspect_ratio_threshold=0.8, min_height=15.0, max_width=50.0, min_width=8,
                  duplicate_threshold=30.0, box: str = "obb", log: Optional[Callable[[str], None]] = None,
                  progress: Optional[Callable[[int], None]] = None, want_points: bool = True,
                  merge_threshold: Optional[float] = None):
    """Stages D+E.  Returns the reference's tower dict list (+ 'label').  merge_threshold (default off) first joins
    clusters whose centres are that close (the variant of test/tttt.py:93-175); the merged clusters carry the labels
    K, K+1, ... like the reference's max(label)+1 numbering."""
    from . import obb as _obb
    centroid = stages.centroid
    stats = stages.stats
    K = stages.n_clusters
    members, label_base, orig_counts = None, 0, stats["count"][:K]
    if merge_threshold is not None and K:
        comp, stats = merge_adjacent_clusters(stats, K, float(merge_threshold))
        members = [np.nonzero(comp == c)[0] for c in range(len(stats))]
        label_base, K = K, len(stats)
    present = np.nonzero(stats["count"][:K] > 0)[0].astype(np.int32)
    # every label 0..K-1 carries at least its head core point, so `present` is all of them; the
    # set() is still built from the values to reproduce the reference's order
    if members is None:
        order = _label_order_array(K, present)
    else:   # the reference iterates set(merged_labels) - {-1}: the set of the NEW label values, built literally
        order = np.asarray(label_iteration_order(K, present + label_base), dtype=np.int64) - label_base
    if box == "aabb" and K:
        # vectorised size filter (test/008.py:302-319 arithmetic in float32, like the reference's numpy):
        # only labels that pass reach the python loop below, in the same set() order
        ext_all = (stats["max"][:K] - stats["min"][:K]).astype(np.float64)
        h_all = ext_all[:, 2]
        w_all = np.maximum(ext_all[:, 0], ext_all[:, 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            ok = (w_all > 0) & (h_all > min_height) & (min_width < w_all) & (w_all < max_width) & \
                 (h_all / w_all > aspect_ratio_threshold)
        order = order[ok[order]]
    order = [int(l) for l in order]
    towers, centres = [], []
    grouped = None   # cluster-major copy of the labelled points, built on the device on first use (O(G), not O(G*K))

    def cluster_rows():
        nonlocal grouped
        if grouped is None:
            grouped = dv.cluster_major_points(stages.filtered, stages.labels, orig_counts)
        return grouped

    def cluster_points(label):
        rows, off = cluster_rows()
        if members is None:
            return rows[int(off[label]): int(off[label + 1])].cpu().numpy()
        return np.concatenate([rows[int(off[m]): int(off[m + 1])].cpu().numpy() for m in members[label]])

    # box="obb": every cluster whose diameter could pass the size filter gets its oriented box ON THE DEVICE, all of
    # them in one launch (one CTA per cluster, pch_obb.cu); nothing but the 144-byte results crosses PCIe
    device_boxes = {}
    if box == "obb" and K and members is None:
        diag = np.linalg.norm((stats["max"][:K] - stats["min"][:K]).astype(np.float64), axis=1)
        cand = [l for l in order if diag[l] > min_height and stats["count"][l] >= 4]
        if cand:
            rows, off = cluster_rows()
            res = dv.obb_batch(rows, np.array([[off[l], off[l + 1]] for l in cand], dtype=np.int64))
            # the size filter for all device boxes at once (same comparisons as the per-label code below); only the
            # boxes that pass, and the clusters the kernel handed back, reach the python loop — in the same set() order
            ext_d = res["extents"]
            h_d, w_d = ext_d[:, 2], np.maximum(ext_d[:, 0], ext_d[:, 1])
            with np.errstate(divide="ignore", invalid="ignore"):
                passes = (h_d > min_height) & (min_width < w_d) & (w_d < max_width) & (h_d / w_d > aspect_ratio_threshold)
            go = passes | (res["status"] != 0)
            device_boxes = {l: r for l, r, g in zip(cand, res, go) if g}
            alive = set(device_boxes)
            order = [l for l in order if l in alive]

    for li, label in enumerate(order):
        try:
            st = stats[label]
            if box == "aabb":
                mn, mx = st["min"], st["max"]
                ext = (mx - mn).astype(np.float64)
                ctr = ((mn + mx) / 2).astype(np.float64)
                rot = np.eye(3)
                height, width = ext[2], max(ext[0], ext[1])
                if width <= 0:
                    continue
                cp = None
            else:
                # a box can only pass the size filter if the cluster's diameter allows it: the OBB's
                # extents are bounded by the AABB diagonal.  Skip hopeless clusters without a hull.
                diag = float(np.linalg.norm((st["max"] - st["min"]).astype(np.float64)))
                if diag <= min_height:
                    continue
                cp = None
                r = device_boxes.get(label)
                if r is not None and int(r["status"]) == 0:
                    a0, a1, a2 = canonical_box_axes(r["rotation"][:, 0], r["rotation"][:, 2])
                    rot = np.column_stack((a0, a1, a2))
                    ext = np.array(r["extents"], dtype=np.float64)
                    ctr = np.array(r["center"], dtype=np.float64)
                else:
                    # "obb_trimesh" / "obb_ordered" (trimesh's own thinned search, host Qhull) and the clusters the
                    # device kernel turned away (degenerate: the host raises like trimesh does; capacity)
                    cp = cluster_points(label)
                    if box == "obb":      # a cluster the kernel handed back: the same exhaustive search on the host
                        tr, ext = _obb.min_volume_box_faces(cp)
                        a0, a1, a2 = canonical_box_axes(tr[:3, 0], tr[:3, 2])
                        tr[:3, :3] = np.column_stack((a0, a1, a2))
                    else:
                        tr, ext = _obb.bounding_box_oriented(cp, ordered=(box == "obb_ordered"))
                    ctr, rot = tr[:3, 3], tr[:3, :3]
                height, width = ext[2], max(ext[0], ext[1])
            aspect = height / width
            if not (height > min_height and min_width < width < max_width and aspect > aspect_ratio_threshold):
                continue
            centre = ctr + centroid
            dup = False
            for c in centres:
                d = np.linalg.norm(centre - c)
                if d < duplicate_threshold:
                    dup = True
                    if log:
                        log(f"⚠️ 跳过重复杆塔{label} (中心距: {d:.1f}m)")
                    break
            if dup:
                continue
            if cp is None and want_points:
                cp = cluster_points(label)
            towers.append({"label": int(label) + label_base, "center": centre, "rotation": rot, "extent": ext,
                           "height": height, "width": width, "north_angle": north_angle_of(rot), "points": cp})
            centres.append(centre)
            if log:
                log(f"✅ 杆塔{label}: {height:.1f}m高 | {width:.1f}m宽 | 中心坐标{centre}")
            if progress:
                progress(75 + int(15 * (li + 1) / max(1, len(order))))
        except Exception as e:  # the reference skips a failing cluster (utils/tower_extraction.py:213-215)
            if log:
                log(f"⚠️ 簇{label} 处理失败: {str(e)}")
            continue
    return towers
