"""Oracle voxel-grid downsample: open3d ``PointCloud.voxel_down_sample`` as the reference calls it.

Follows ui/import_PC.py:8-13 (process_chunk) and :45-65 (chunk loop, concat, re-quantising write);
ui/Sampling.py:10-18,21-80 is the same algorithm.  open3d is absent (PARITY UNPINNED); restated
from its public behaviour:
    origin = min_bound - 0.5*v ; idx = floor((p - origin)/v) per axis (float64 subtract, true divide)
    per voxel: running float64 sum in INPUT order, mean = sum / count
open3d emits voxels in unordered_map iteration order (unspecified); the canonical order used by
this oracle and by the CUDA path is lexicographic (ix, iy, iz), ix most significant.
"""
import numpy as np

from . import las_io


def voxel_indices(points, voxel_size):
    """(n,3) int64 voxel indices exactly as open3d computes them for this chunk."""
    if voxel_size <= 0:
        raise ValueError("voxel_size <= 0")
    p = np.asarray(points, dtype=np.float64)
    origin = p.min(axis=0) - voxel_size * 0.5
    ref = (p - origin) / voxel_size
    return np.floor(ref).astype(np.int64)


def voxel_down_sample(points, voxel_size, return_groups=False):
    """Canonically ordered (m,3) float64 voxel means of one chunk."""
    p = np.asarray(points, dtype=np.float64)
    if p.shape[0] == 0:
        out = np.zeros((0, 3))
        return (out, np.zeros((0, 3), np.int64), np.zeros(0, np.int64)) if return_groups else out
    idx = voxel_indices(p, voxel_size)
    ext = idx.max(axis=0) + 1
    if float(ext[0]) * float(ext[1]) * float(ext[2]) < 2**62:
        key = (idx[:, 0] * ext[1] + idx[:, 1]) * ext[2] + idx[:, 2]
        ukey, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        uidx = np.stack([ukey // (ext[1] * ext[2]), (ukey // ext[2]) % ext[1], ukey % ext[2]], axis=1)
    else:  # index range too wide to pack: row-wise unique is the same lexicographic (ix, iy, iz) order
        uidx, inv, cnt = np.unique(idx, axis=0, return_inverse=True, return_counts=True)
        inv = inv.reshape(-1)
    m = uidx.shape[0]
    means = np.empty((m, 3))
    for a in range(3):
        # bincount adds weights one by one in input order: the in-order float64 running sum
        means[:, a] = np.bincount(inv, weights=p[:, a], minlength=m) / cnt.astype(np.float64)
    if return_groups:
        return means, uidx, cnt
    return means


def downsample_las_arrays(las, voxel_size=0.1, chunk_size=1000000):
    """The chunk loop of run_voxel_downsampling on a decoded LAS dict; returns the concatenated
    float64 means (M,3) and the per-chunk output counts."""
    n = las["n"]
    outs, counts = [], []
    for start in range(0, n, chunk_size):
        end = min(start + chunk_size, n)
        x, y, z = las_io.scaled(las, start, end)
        pts = np.vstack((x, y, z)).T
        d = voxel_down_sample(pts, voxel_size)
        outs.append(d)
        counts.append(d.shape[0])
    final = np.vstack(outs) if outs else np.zeros((0, 3))
    return final, np.array(counts, dtype=np.int64)


def run_voxel_downsampling(input_path, output_path, voxel_size=0.1, chunk_size=1000000):
    las = las_io.read_las(input_path)
    final, counts = downsample_las_arrays(las, voxel_size, chunk_size)
    las_io.write_las(output_path, las, final[:, 0], final[:, 1], final[:, 2])
    return final, counts
