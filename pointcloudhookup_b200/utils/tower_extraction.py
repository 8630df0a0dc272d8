"""Drop-in for the reference's utils/tower_extraction.py (identical copy: ui/ui/tower_extraction.py).

Same signature, defaults, callback milestones, side effects (./output_towers/tower_<label>.las,
towers_info.xlsx when an Excel writer is installed) and error behaviour (never raises; returns
what it has).  The LAS decode, float32 cast, centroid, percentile filter, chunked DBSCAN, the
per-cluster reduction and the per-tower LAS encode run on the GPU (..towers, ..device).
Extra keyword-only options (defaults reproduce the reference): ``box`` ("obb" | "obb_ordered" |
"aabb" — test/008.py:302-319), ``ground`` ("percentile" | "grid" — the north_star grid min-z mode),
``merge_threshold`` (metres; joins clusters whose centres are that close before the box test — the
post-processing of test/tttt.py:93-175; None = off).
"""
import math  # noqa: F401  (kept: the reference module exposes the same top-level names)
import time
from pathlib import Path

import numpy as np
import torch

from .. import device as dv
from .. import las as _las
from .. import towers as _tw


def extract_towers(
        input_las_path,
        progress_callback=None,
        log_callback=None,
        eps=8.0,
        min_points=80,
        aspect_ratio_threshold=0.8,
        min_height=15.0,
        max_width=50.0,
        min_width=8,
        duplicate_threshold=30.0,
        *, box="obb", ground="percentile", write_outputs=True, merge_threshold=None
):
    tower_obbs = []
    tower_info_list = []

    def log(msg):
        if log_callback:
            log_callback(msg)
        else:
            print(msg)

    def progress(value):
        if progress_callback:
            progress_callback(value)

    output_dir = Path("output_towers")
    if write_outputs:
        output_dir.mkdir(exist_ok=True)

    # ---- read + float32 cast + centroid (utils/tower_extraction.py:57-76)
    try:
        log("📂 读取点云文件...")
        progress(5)
        hdr, rec = _las.read_raw(str(input_las_path))
        dl = dv.upload_records_xyz(rec, hdr.point_count, hdr.record_length, hdr.scales, hdr.offsets)
        raw = dv.decode_xyz(dl, torch.float32)
        if raw.shape[0] == 0:
            raise ValueError("empty point cloud")
        header_info = {"scales": hdr.scales, "offsets": hdr.offsets, "point_format": hdr.point_format,
                       "version": hdr.version, "header": hdr}
        log(f"✅ 点云读取完成，总点数: {raw.shape[0]}")
    except Exception as e:
        log(f"⚠️ 文件读取失败: {str(e)}")
        return tower_obbs

    # ---- height filter (:79-93)
    try:
        log("🔍 执行高度过滤...")
        progress(10)
        if ground == "percentile":
            filtered, cen_dev, base, used, _ = _tw.ground_filter_percentile(raw)
        else:
            filtered, cen_dev, base, used, _ = _tw.ground_filter_grid(raw)
        centroid = cen_dev.cpu().numpy()
        header_info["centroid"] = centroid
        log(f"✅ 高度过滤完成，保留点数: {filtered.shape[0]}")
        if used != 3.0 and ground == "percentile":
            log("⚠️ 过滤后点数太少，尝试降低过滤阈值")
    except Exception as e:
        log(f"⚠️ 高度过滤失败: {str(e)}")
        return tower_obbs

    # ---- chunked DBSCAN (:96-122): all chunks in one batched device pass
    log("\n=== 开始聚类处理 ===")
    progress(20)
    n_f = filtered.shape[0]
    n_chunks = -(-n_f // _tw.DBSCAN_CHUNK) if n_f else 0
    try:
        db = dv.dbscan_chunked(filtered, eps, min_points, _tw.DBSCAN_CHUNK)
        for i in range(n_chunks):
            size = min(_tw.DBSCAN_CHUNK, n_f - i * _tw.DBSCAN_CHUNK)
            log(f"处理分块 {i + 1}/{n_chunks} ({size}点)")
            progress(20 + int(50 * (i + 1) / n_chunks))
    except Exception as e:
        log(f"⚠️ 分块聚类失败: {str(e)}")
        return tower_obbs

    stages = _tw.TowerStages(raw, centroid, base, used, filtered, db.labels, db.n_clusters, db.stats)

    # ---- detection + dedup (:125-218)
    log(f"\n=== 开始杆塔检测（候选簇：{db.n_clusters}个） ===")
    progress(75)
    towers = _tw.select_towers(stages, aspect_ratio_threshold, min_height, max_width, min_width, duplicate_threshold,
                               box=box, log=log, progress=progress, merge_threshold=merge_threshold)
    for t in towers:
        label = t.pop("label")
        tower_obbs.append(t)
        c = t["center"]
        tower_info_list.append({"ID": f"tower_{label}", "经度": c[0], "纬度": c[1], "海拔高度": c[2],
                                "杆塔高度": t["height"], "北方向偏角": t["north_angle"], "宽度": t["width"],
                                "长宽比": t["height"] / t["width"]})
        if write_outputs:
            original_points = t["points"] + centroid
            _save_tower_las(original_points, None, header_info, output_dir / f"tower_{label}.las", log)

    # ---- Excel (:221-233)
    if tower_info_list:
        if write_outputs:
            try:
                import pandas as pd
                pd.DataFrame(tower_info_list).to_excel("towers_info.xlsx", index=False)
                log("\n✅ 杆塔信息已保存到: towers_info.xlsx")
                log(f"检测到杆塔数量: {len(tower_obbs)}个")
            except Exception as e:
                log(f"⚠️ 保存Excel失败: {str(e)}")
    else:
        log("\n⚠️ 未检测到任何杆塔，不生成Excel文件")

    log("\n=== 清理内存 ===")
    progress(100)
    log("✅ 杆塔提取完成")
    return tower_obbs


def _save_tower_las(points, colors, header_info, output_path, log_callback=None):
    """utils/tower_extraction.py:243-262: new header with the source's format/version/scales/offsets,
    las.x/y/z = points (float64) -> int32 lattice, write.  Quantise + encode run on the GPU."""
    try:
        src = header_info.get("header")
        if src is None:
            pf = header_info["point_format"]
            src = _las.LasHeader(version=tuple(header_info["version"]), point_format=int(getattr(pf, "id", pf)),
                                 scales=np.asarray(header_info["scales"], float),
                                 offsets=np.asarray(header_info["offsets"], float))
        hdr = _las.new_header_like(src)
        pts = np.ascontiguousarray(np.stack([points[:, 0].astype(np.float64), points[:, 1].astype(np.float64),
                                             points[:, 2].astype(np.float64)], axis=1))
        lat = dv.quantise(torch.from_numpy(pts).cuda(), hdr.scales, hdr.offsets)
        recs, mm = dv.encode_records(lat, hdr.record_length)
        mmh = mm.cpu().numpy()
        _las.write_raw(str(output_path), hdr, recs.cpu().numpy(), mmh[:3] if len(pts) else None,
                       mmh[3:] if len(pts) else None)
        if log_callback:
            log_callback(f"保存成功：{output_path}")
    except Exception as e:
        if log_callback:
            log_callback(f"⚠️ 保存失败 {output_path}: {str(e)}")


def obb_wireframe(center, rotation, extent):
    """8 corners + 12 edges of an oriented box in Open3D's LineSet.create_from_oriented_bounding_box
    vertex order — what utils/tower_extraction.py:265-279 builds with open3d for display."""
    c = np.asarray(center, dtype=np.float64)
    r = np.asarray(rotation, dtype=np.float64)
    e = np.asarray(extent, dtype=np.float64)
    x, y, z = r[:, 0] * e[0] * 0.5, r[:, 1] * e[1] * 0.5, r[:, 2] * e[2] * 0.5
    pts = np.array([c - x - y - z, c + x - y - z, c - x + y - z, c - x - y + z,
                    c + x + y + z, c - x + y + z, c + x - y + z, c + x + y - z])
    lines = np.array([[0, 1], [1, 7], [7, 2], [2, 0], [3, 6], [6, 4], [4, 5], [5, 3], [0, 3], [1, 6], [7, 4], [2, 5]])
    return pts, lines


class _LineSet:
    """Minimal stand-in for open3d.geometry.LineSet (points, lines, colors) so callers that only read
    these attributes keep working without open3d."""

    def __init__(self, points, lines, color):
        self.points = points
        self.lines = lines
        self.colors = np.tile(np.asarray(color, dtype=np.float64), (len(lines), 1))


def create_obb_geometries(tower_obbs):
    geometries = []
    for tower in tower_obbs:
        try:
            pts, lines = obb_wireframe(tower['center'], tower['rotation'], tower['extent'])
            geometries.append(_LineSet(pts, lines, [1, 0, 0]))
        except Exception:
            continue
    return geometries


def extract_towers_optimized(*args, **kwargs):
    return extract_towers(*args, **kwargs)


if __name__ == "__main__":
    import sys
    start_time = time.time()
    try:
        extract_towers(input_las_path=sys.argv[1] if len(sys.argv) > 1 else "output/point_2.las")
    except Exception as e:
        print(f"⚠️ 程序错误: {str(e)}")
    finally:
        print(f"\n总运行时间: {time.time() - start_time:.1f}秒")
