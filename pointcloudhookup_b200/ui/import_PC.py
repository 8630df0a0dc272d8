"""Drop-in for the reference's ui/import_PC.py (the version the live GUI calls,
pyGUI_towers_test.py:351-358): same names, positional order, defaults, callbacks and errors; the
laspy decode, open3d voxel grid and laspy re-quantising write run on the GPU."""
import os
from typing import Callable

import numpy as np
import torch

from .. import device as dv
from .. import las as _las


def process_chunk(points_chunk, voxel_size):
    """Voxel-downsample one (n,3) point block (ui/import_PC.py:8-13).  Returns (m,3) float64; open3d's
    output order is unspecified, ours is lexicographic by voxel index."""
    pts = np.ascontiguousarray(np.asarray(points_chunk).astype(np.float64))
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError("points_chunk must have shape (n, 3)")
    dev = torch.from_numpy(pts).cuda()
    res = dv.voxel_downsample_points(dev, float(voxel_size))
    return res.mean.cpu().numpy()


def _downsample_file(input_path, output_path, voxel_size, chunk_size, progress_callback=None, log_callback=None):
    hdr, rec = _las.read_raw(input_path)
    total_points = hdr.point_count
    if log_callback:
        log_callback(f"📂 原始点数: {total_points}")
        log_callback(f"✨ 开始下采样（voxel_size={voxel_size}, chunk_size={chunk_size}）")
    dl = dv.upload_records_xyz(rec, total_points, hdr.record_length, hdr.scales, hdr.offsets)
    res = dv.voxel_downsample(dl, float(voxel_size), int(chunk_size), want=("lattice",))
    for i, start in enumerate(range(0, total_points, int(chunk_size))):
        end = min(start + int(chunk_size), total_points)
        if log_callback:
            log_callback(f"✅ 已完成第{i+1}块：{end - start} 点")
        if progress_callback:
            progress_callback(int((end / total_points) * 100))
    out_hdr = _las.new_header_like(hdr)
    recs, mm = dv.encode_records(res.lattice, out_hdr.record_length)
    mmh = mm.cpu().numpy()
    _las.write_raw(output_path, out_hdr, recs.cpu().numpy(), mmh[:3] if res.count else None, mmh[3:] if res.count else None)
    return res.count


def run_voxel_downsampling(
    input_path: str,
    output_path: str,
    voxel_size: float = 0.1,
    chunk_size: int = 1000000,
    progress_callback: Callable[[int], None] = None,
    log_callback: Callable[[str], None] = None
):
    if not os.path.exists(input_path):
        raise FileNotFoundError(f"输入文件不存在: {os.path.abspath(input_path)}")

    os.makedirs(os.path.dirname(output_path), exist_ok=True)

    count = _downsample_file(input_path, output_path, voxel_size, chunk_size, progress_callback, log_callback)

    if log_callback:
        log_callback(f"✅ 下采样完成，输出点数: {count}")
        log_callback(f"📁 保存至：{output_path}")
