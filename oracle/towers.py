"""Oracle tower extraction: utils/tower_extraction.py::extract_towers restated stage by stage.

Stage A  :57-76    read, stack, astype(float32), centroid = np.mean(axis=0), subtract
Stage B  :79-93    percentile height filter (oracle.ground.percentile_keep_mask)
Stage C  :96-122   DBSCAN on consecutive 50 000-point chunks, labels offset per chunk
Stage D  :125-147  per label in ``set(all_labels) - {-1}`` ORDER: box, size filter
Stage E  :150-218  centre, duplicate check against accepted centres, north angle
The real numpy and the real scikit-learn DBSCAN are called (they are the reference's own
dependencies).  The box is trimesh's hull-face search over every face normal (``box="obb"``,
oracle.obb.min_volume_box_all_faces), its literal thinned form (``box="obb_trimesh"``; both PARITY UNPINNED) or,
with ``box="aabb"``, the axis-aligned variant of test/008.py:302-319 (exactly restatable).
"""
import numpy as np
from sklearn.cluster import DBSCAN

from . import ground, las_io, obb

DBSCAN_CHUNK = 50000


def stage_a(las):
    x, y, z = las_io.scaled(las)
    raw = np.stack([x, y, z], axis=1).astype(np.float32)
    centroid = np.mean(raw, axis=0)
    return raw, centroid, raw - centroid


def stage_c(filtered, eps=8.0, min_points=80, chunk_size=DBSCAN_CHUNK, n_jobs=-1):
    all_labels = np.full(len(filtered), -1, dtype=np.int32)
    current = 0
    for i in range(0, len(filtered), chunk_size):
        chunk = filtered[i:i + chunk_size]
        lab = DBSCAN(eps=eps, min_samples=min_points, n_jobs=n_jobs, algorithm="ball_tree").fit(chunk).labels_
        lab[lab != -1] += current
        all_labels[i:i + chunk_size] = lab
        current = np.max(lab) + 1 if np.any(lab != -1) else current
    return all_labels


def north_angle_of(rotation):
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


def cluster_box(cluster_points, box):
    """(extents (3,), centre in the centred frame (3,), rotation (3,3))."""
    if box == "aabb":
        mn = np.min(cluster_points, axis=0)
        mx = np.max(cluster_points, axis=0)
        return (mx - mn).astype(np.float64), ((mn + mx) / 2).astype(np.float64), np.eye(3)
    if box == "obb":      # every hull-face normal (what the device kernel evaluates)
        tr, ext, _ = obb.min_volume_box_all_faces(cluster_points)
        return ext, tr[:3, 3], tr[:3, :3]
    # "obb_trimesh" / "obb_ordered": trimesh's own thinned search, literally
    tr, ext = obb.bounding_box_oriented(cluster_points, ordered=(box == "obb_ordered"))
    return ext, tr[:3, 3], tr[:3, :3]


def extract_towers_arrays(las, eps=8.0, min_points=80, aspect_ratio_threshold=0.8, min_height=15.0,
                          max_width=50.0, min_width=8, duplicate_threshold=30.0, box="obb",
                          n_jobs=-1, intermediates=None):
    raw, centroid, points = stage_a(las)
    z = points[:, 2]
    mask, base, used = ground.percentile_keep_mask(z)
    filtered = points[mask]
    all_labels = stage_c(filtered, eps, min_points, n_jobs=n_jobs)
    unique_labels = set(all_labels) - {-1}
    if intermediates is not None:
        intermediates.update(raw=raw, centroid=centroid, base=base, offset_used=used, mask=mask,
                             filtered=filtered, labels=all_labels,
                             label_order=[int(v) for v in unique_labels])
    towers, centres = [], []
    for label in unique_labels:
        try:
            cp = filtered[all_labels == label]
            ext, ctr, rot = cluster_box(cp, box)
            height = ext[2]
            width = max(ext[0], ext[1])
            if box == "aabb" and width <= 0:
                continue
            aspect = height / width
            if not (height > min_height and min_width < width < max_width and aspect > aspect_ratio_threshold):
                continue
            centre = ctr + centroid
            if any(np.linalg.norm(centre - c) < duplicate_threshold for c in centres):
                continue
            towers.append({"label": int(label), "center": centre, "rotation": rot, "extent": ext,
                           "height": height, "width": width, "north_angle": north_angle_of(rot),
                           "points": cp})
            centres.append(centre)
        except Exception:
            continue
    return towers


def merge_adjacent_clusters(filtered, all_labels, merge_threshold=6.0):
    """test/tttt.py:93-175, literally: cluster centres = np.mean(cluster_points, axis=0) in set() order, sklearn KDTree
    radius query, union by cluster size, new labels max+1.. in order of first member.  Returns merged_labels."""
    from sklearn.neighbors import KDTree
    unique_labels = set(all_labels) - {-1}
    if not unique_labels:
        return all_labels
    cluster_centers, label_to_index, valid_labels = [], {}, []
    for label in unique_labels:
        cluster_points = filtered[all_labels == label]
        if len(cluster_points) > 0:
            cluster_centers.append(np.mean(cluster_points, axis=0))
            label_to_index[label] = len(cluster_centers) - 1
            valid_labels.append(label)
    if not cluster_centers:
        return all_labels
    cluster_centers = np.array(cluster_centers)
    tree = KDTree(cluster_centers)
    neighbors = tree.query_radius(cluster_centers, r=merge_threshold)
    parent = list(range(len(cluster_centers)))

    def find(x):
        if parent[x] != x:
            parent[x] = find(parent[x])
        return parent[x]

    def union(x, y):
        root_x, root_y = find(x), find(y)
        if root_x != root_y:
            size_x = np.sum(all_labels == valid_labels[x])
            size_y = np.sum(all_labels == valid_labels[y])
            if size_x > size_y:
                parent[root_y] = root_x
            else:
                parent[root_x] = root_y

    for i, neighbor_indices in enumerate(neighbors):
        for j in neighbor_indices:
            if i < j:
                union(i, j)
    new_labels = {}
    current_max_label = max(unique_labels) + 1
    for i in range(len(cluster_centers)):
        root = find(i)
        if root not in new_labels:
            new_labels[root] = current_max_label
            current_max_label += 1
    merged_labels = all_labels.copy()
    for label in valid_labels:
        merged_labels[all_labels == label] = new_labels[find(label_to_index[label])]
    return merged_labels


def extract_towers(input_las_path, **kw):
    return extract_towers_arrays(las_io.read_las(input_las_path), **kw)
