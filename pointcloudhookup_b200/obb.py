"""Host epilogue: minimum-volume oriented box of ONE cluster, the role trimesh plays in the reference
(``trimesh.PointCloud(cluster_points).bounding_box_oriented``, utils/tower_extraction.py:137-139).

This is O(#clusters) work on clusters that survived the device-side AABB pre-filter, exactly where
the reference calls into trimesh/Qhull; scipy's ConvexHull is the same Qhull.  trimesh is absent and
unpinned in the reference, so the algorithm is restated from its documented behaviour (SURVEY.md
A.5): hull -> unique face normals (hemisphere-folded, spherical angles rounded to 1 decimal) ->
for each, rotate normal to +Z, min-area edge-aligned rectangle of the projected hull -> smallest
volume.  Extents stay in (rect long, rect short, along-normal) order unless ``ordered``.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import ConvexHull

_TOL = np.finfo(np.float64).resolution * 100


def _fold_to_hemisphere(normals: np.ndarray) -> np.ndarray:
    n = np.array(normals, dtype=np.float64)
    negative = n < -_TOL
    zero = ~(negative | (n > _TOL))
    sign = np.ones(len(n))
    sign[negative[:, 2]] = -1.0
    sign[zero[:, 2] & negative[:, 1]] = -1.0
    sign[zero[:, 2] & zero[:, 1] & negative[:, 0]] = -1.0
    return n * sign[:, None]


def _to_plane_matrix(theta: float, phi: float) -> np.ndarray:
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    spherical = np.array([[ct, -st, 0.0], [st, ct, 0.0], [0.0, 0.0, 1.0]]) @ \
        np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    out = np.eye(4)
    out[:3, :3] = spherical.T
    return out


def _planar(theta: float, offset=(0.0, 0.0)) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    t = np.eye(3)
    t[0, :2] = [c, s]
    t[1, :2] = [-s, c]
    t[:2, 2] = offset
    return t


def min_area_rectangle(xy: np.ndarray):
    hull = ConvexHull(xy, qhull_options="QbB")
    seg = hull.points[hull.simplices]
    verts = hull.points[hull.vertices]
    edge = seg[:, 1] - seg[:, 0]
    length = np.sqrt(np.dot(edge ** 2, [1, 1]))
    keep = length > 1e-10
    edge = edge[keep] / length[keep].reshape((-1, 1))
    perp = np.fliplr(edge) * [-1.0, 1.0]
    px = np.dot(edge, verts.T)
    py = np.dot(perp, verts.T)
    bounds = np.column_stack((px.min(axis=1), py.min(axis=1), px.max(axis=1), py.max(axis=1)))
    extents = np.diff(bounds.reshape((-1, 2, 2)), axis=1).reshape((-1, 2))
    best = np.prod(extents, axis=1).argmin()
    rect = extents[best]
    offset = -bounds[best][:2] - rect * 0.5
    theta = np.arctan2(*edge[best][::-1])
    t = _planar(theta, offset)
    if rect[0] < rect[1]:
        t = _planar(np.pi / 2) @ t
        rect = np.roll(rect, 1)
    return t, rect


def oriented_bounds(points: np.ndarray, angle_digits: int = 1, ordered: bool = False):
    pts = np.asarray(points, dtype=np.float64)
    hull = ConvexHull(pts, qhull_options="QbB Pp Qt")
    verts = pts[hull.vertices]
    # face normals from the triangles in the original frame: with "QbB" Qhull's `equations` describe the points
    # scaled to a unit cube (every axis by a different factor), not the input
    tri = pts[hull.simplices]
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    area2 = np.linalg.norm(fn, axis=1)
    tri, fn = tri[area2 > 0], fn[area2 > 0] / area2[area2 > 0][:, None]
    fn[np.einsum("ij,ij->i", fn, tri[:, 0] - verts.mean(axis=0)) < 0] *= -1.0
    hemi = _fold_to_hemisphere(fn)
    angles = np.column_stack((np.arctan2(hemi[:, 1], hemi[:, 0]), np.arccos(np.clip(hemi[:, 2], -1.0, 1.0))))
    _, first = np.unique(np.round(angles * 10 ** angle_digits).astype(np.int64), axis=0, return_index=True)
    best_volume, best = np.inf, None
    for i in first:
        to_plane = _to_plane_matrix(angles[i, 0], angles[i, 1])
        proj = verts @ to_plane[:3, :3].T + to_plane[:3, 3]
        thick = np.ptp(proj[:, 2])
        rot2, rect = min_area_rectangle(proj[:, :2])
        vol = np.prod(rect) * thick
        if vol < best_volume:
            best_volume = vol
            ext = np.append(rect, thick)
            rz = np.eye(4)
            rz[:2, :2] = rot2[:2, :2]
            best = (to_plane.copy(), rz)
    to_origin = best[1] @ best[0]
    moved = verts @ to_origin[:3, :3].T + to_origin[:3, 3]
    to_origin[:3, 3] = -(moved.min(axis=0) + np.ptp(moved, axis=0) * 0.5)
    if ordered:
        order = ext.argsort()
        ext = ext[order]
        flip = np.eye(4)
        flip[:3, :3] = -np.eye(3)[order]
        flip[:3, :3] *= np.linalg.det(flip[:3, :3])
        to_origin = flip @ to_origin
    return to_origin, ext


def bounding_box_oriented(points: np.ndarray, ordered: bool = False):
    """(transform box->world 4x4, extents (3,)) — the ``obb.transform`` / ``obb.extents`` the
    reference reads (utils/tower_extraction.py:139,151,165)."""
    to_origin, ext = oriented_bounds(points, ordered=ordered)
    return np.linalg.inv(to_origin), ext


def min_volume_box_faces(points: np.ndarray):
    """Host stand-in for the device kernel (pch_obb.cu) on the clusters it hands back (hulls beyond its capacity): the
    SAME search — every hull-face normal; for each, every SILHOUETTE edge of the hull (the two faces that share it look to
    opposite sides of the normal = an edge of the projected hull) as rectangle direction — so that a tower list does not
    depend on where a box was computed.  Qhull raises for flat input like trimesh.
    -> (transform box->world 4x4, extents (long, short, along-normal))."""
    pts = np.asarray(points, dtype=np.float64)
    hull = ConvexHull(pts, qhull_options="Pp Qt")
    verts = pts[hull.vertices]
    simp, nb = hull.simplices, hull.neighbors
    tri = pts[simp]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    ln = np.linalg.norm(nrm, axis=1)
    ok = ln > 0
    nrm[ok] /= ln[ok][:, None]
    nrm[np.einsum("ij,ij->i", nrm, tri[:, 0] - verts.mean(axis=0)) < 0] *= -1.0
    # edge k of simplex i is opposite vertex k; its neighbour facet is neighbors[i][k]
    f_of = np.repeat(np.arange(len(simp)), 3)
    g_of = nb.reshape(-1)
    ev = np.stack([pts[simp[:, [1, 2, 0]].reshape(-1)], pts[simp[:, [2, 0, 1]].reshape(-1)]], axis=1)
    evec = ev[:, 1] - ev[:, 0]
    uniq = np.unique(np.round(nrm[ok], 12), axis=0, return_index=True)[1]
    best = (np.inf, None, None)
    for n in nrm[ok][np.sort(uniq)]:
        side = nrm @ n
        sil = (side[f_of] >= 0) & (side[g_of] < 0)
        u = evec[sil] - np.outer(evec[sil] @ n, n)
        ul = np.linalg.norm(u, axis=1)
        u = u[ul > 1e-10] / ul[ul > 1e-10][:, None]
        if len(u) == 0:
            continue
        v = np.cross(n, u)
        pu, pv = verts @ u.T, verts @ v.T
        vol = np.ptp(pu, axis=0) * np.ptp(pv, axis=0) * np.ptp(verts @ n)
        k = int(np.argmin(vol))
        if vol[k] < best[0]:
            best = (float(vol[k]), n, u[k])
    _, n, u = best
    v = np.cross(n, u)
    ext = np.array([np.ptp(verts @ u), np.ptp(verts @ v), np.ptp(verts @ n)])
    if ext[0] < ext[1]:
        u, v = v, -u
        ext[[0, 1]] = ext[[1, 0]]
    rot = np.column_stack((u, v, n))
    loc = verts @ rot
    t = np.eye(4)
    t[:3, :3] = rot
    t[:3, 3] = rot @ (loc.min(axis=0) + 0.5 * np.ptp(loc, axis=0))
    return t, ext
