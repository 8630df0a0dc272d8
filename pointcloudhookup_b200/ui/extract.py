"""Drop-in for the reference's ui/extract.py (called by pyGUI_towers_test.py:528-533).

The per-point part — reading the whole LAS into an (N,3) float64 array (ui/extract.py:114-115,
361-362) — is the GPU decode kernel.  The display boxes are O(#towers) host geometry and keep the
reference's formulas: the "kuangxuan" (box-select) asymmetric box of ui/extract.py:7-38 /
test/kuangxuan.py:69-71, the symmetric variant (:149-160), and the scaled oriented box (:345-420)
whose wireframe is built in numpy in Open3D's corner/edge order instead of through open3d.
"""
import os

import numpy as np
import torch

from .. import device as dv
from .. import las as _las
from ..utils.tower_extraction import obb_wireframe

_EDGES = ((0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7))


def _read_points_f64(las_path):
    if not os.path.exists(las_path):
        raise FileNotFoundError(f"未找到文件: {las_path}")
    hdr, rec = _las.read_raw(las_path)
    dl = dv.upload_records_xyz(rec, hdr.point_count, hdr.record_length, hdr.scales, hdr.offsets)
    return dv.decode_xyz(dl, torch.float64).cpu().numpy()


def create_bbox_using_kuangxuan_method(center, width, height,
                                       x_left_factor=1.0, x_right_factor=1.67,
                                       y_down_factor=0.5, y_up_factor=1.0,
                                       z_down_factor=1.0, z_up_factor=2.0):
    """Asymmetric axis-aligned box around a tower centre; returns (min_coords, max_coords)."""
    cx, cy, cz = center
    lo = np.array([cx - width * x_left_factor, cy - width * y_down_factor, cz - height * z_down_factor])
    hi = np.array([cx + width * x_right_factor, cy + width * y_up_factor, cz + height * z_up_factor])
    return lo, hi


def create_bbox_lineset_from_bounds(min_coords, max_coords, color=(1.0, 0.0, 0.0)):
    """24 points (12 edges x 2 end points) of the axis-aligned box wireframe, plus the colour."""
    x0, y0, z0 = min_coords
    x1, y1, z1 = max_coords
    corners = [[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0],
               [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]]
    pts = [corners[i] for edge in _EDGES for i in edge]
    return np.array(pts), color


def _default_kuangxuan_params():
    return {"x_left_factor": 1.0, "x_right_factor": 1.67, "y_down_factor": 0.5, "y_up_factor": 1.0,
            "z_down_factor": 1.0, "z_up_factor": 2.0}


def _tower_bounds(tower_info, bbox_method, bbox_params):
    center = tower_info['center']
    ext = np.array(tower_info['extent'])
    width = max(ext[0], ext[1])
    height = ext[2]
    if bbox_method == "kuangxuan":
        lo, hi = create_bbox_using_kuangxuan_method(center, width, height, **bbox_params)
    elif bbox_method == "symmetric":
        half = np.array([width * bbox_params.get("x_scale", 2.0), width * bbox_params.get("y_scale", 2.0),
                         height * bbox_params.get("z_scale", 1.5)]) / 2
        lo, hi = center - half, center + half
    else:
        raise ValueError(f"未知的包围盒方法: {bbox_method}")
    return center, width, height, lo, hi


def extract_and_visualize_towers_kuangxuan(las_path: str, tower_obbs: list,
                                           bbox_method: str = "kuangxuan",
                                           bbox_params: dict = None,
                                           line_color: tuple = (1.0, 0.0, 0.0)):
    if bbox_params is None:
        bbox_params = _default_kuangxuan_params()
    full_pcd = _read_points_f64(las_path)
    tower_geometries = []
    print(f"🔧 开始处理 {len(tower_obbs)} 个杆塔，使用方法: {bbox_method}")
    print(f"📊 包围盒参数: {bbox_params}")
    for i, tower_info in enumerate(tower_obbs):
        try:
            center, width, height, lo, hi = _tower_bounds(tower_info, bbox_method, bbox_params)
            if bbox_method == "kuangxuan":
                size = hi - lo
                print(f"📏 杆塔{i}: 原始宽度{width:.1f}m, 高度{height:.1f}m")
                print(f"📐 杆塔{i}: kuangxuan方法 -> X:{size[0]:.1f}m, Y:{size[1]:.1f}m, Z:{size[2]:.1f}m")
            box_pts, color = create_bbox_lineset_from_bounds(lo, hi, line_color)
            tower_geometries.append((box_pts, color))
            print(f"✅ 杆塔{i}处理成功，中心：{center}")
        except Exception as e:
            print(f"⚠️ 杆塔{i}可视化失败: {str(e)}")
            continue
    print(f"✅ 成功处理 {len(tower_geometries)} 个杆塔几何体")
    return full_pcd, tower_geometries


def create_enhanced_tower_boxes_kuangxuan(tower_obbs: list,
                                          bbox_method: str = "kuangxuan",
                                          bbox_params: dict = None,
                                          add_center_marker: bool = True,
                                          add_height_indicator: bool = True):
    if bbox_params is None:
        bbox_params = _default_kuangxuan_params()
    out = []
    for tower_info in tower_obbs:
        try:
            center, width, height, lo, hi = _tower_bounds(tower_info, bbox_method, bbox_params)
            main_pts, _ = create_bbox_lineset_from_bounds(lo, hi, (1.0, 0.0, 0.0))
            out.append((main_pts, (1.0, 0.0, 0.0)))
            if add_center_marker:
                m = min(width, height) * 0.1 / 2
                pts, _ = create_bbox_lineset_from_bounds(center - np.array([m, m, m]), center + np.array([m, m, m]),
                                                         (1.0, 1.0, 0.0))
                out.append((pts, (1.0, 1.0, 0.0)))
            if add_height_indicator:
                out.append((np.array([np.array([center[0], center[1], lo[2]]),
                                      np.array([center[0], center[1], hi[2]])]), (0.0, 1.0, 0.0)))
        except Exception:
            continue
    return out


BBOX_PRESETS = {
    "kuangxuan_original": {"method": "kuangxuan", "params": _default_kuangxuan_params()},
    "kuangxuan_conservative": {"method": "kuangxuan", "params": {
        "x_left_factor": 0.8, "x_right_factor": 1.2, "y_down_factor": 0.4, "y_up_factor": 0.8,
        "z_down_factor": 0.5, "z_up_factor": 1.5}},
    "kuangxuan_aggressive": {"method": "kuangxuan", "params": {
        "x_left_factor": 1.5, "x_right_factor": 2.0, "y_down_factor": 0.8, "y_up_factor": 1.5,
        "z_down_factor": 1.5, "z_up_factor": 3.0}},
    "symmetric_moderate": {"method": "symmetric", "params": {"x_scale": 2.0, "y_scale": 2.0, "z_scale": 1.5}},
    "symmetric_large": {"method": "symmetric", "params": {"x_scale": 3.0, "y_scale": 3.0, "z_scale": 2.0}},
}


def get_bbox_preset(preset_name: str):
    preset = BBOX_PRESETS.get(preset_name, BBOX_PRESETS["kuangxuan_original"])
    return preset["method"], preset["params"]


def visualize_towers_with_point_cloud_kuangxuan(las_path: str, tower_obbs: list,
                                                preset_name: str = "kuangxuan_original",
                                                output_path: str = None):
    try:
        method, params = get_bbox_preset(preset_name)
        full_pcd, geoms = extract_and_visualize_towers_kuangxuan(las_path, tower_obbs, method, params)
        if output_path:
            print(f"💾 结果将保存到: {output_path}")
        return full_pcd, geoms
    except Exception as e:
        print(f"❌ 可视化失败: {str(e)}")
        return None, []


def extract_and_visualize_towers_original(las_path: str, tower_obbs: list,
                                          scale_factors: list = None,
                                          line_color: tuple = (1.0, 0.0, 0.0),
                                          adaptive_scaling: bool = True):
    if scale_factors is None:
        scale_factors = [2.8, 2.8, 4.5]
    full_pcd = _read_points_f64(las_path)
    geoms = []
    print(f"🔧 开始处理 {len(tower_obbs)} 个杆塔，使用放大因子: {scale_factors}")
    for i, tower_info in enumerate(tower_obbs):
        try:
            ext = np.array(tower_info['extent'])
            if adaptive_scaling:
                h = ext[2]
                scale = [3.2, 3.2, 5.0] if h < 20 else ([3.0, 3.0, 4.8] if h < 40 else [2.8, 2.8, 4.5])
                print(f"📏 杆塔{i}: 高度{h:.1f}m, 自适应缩放{scale}")
            else:
                scale = scale_factors
                print(f"📏 杆塔{i}: 固定缩放{scale_factors}")
            big = ext * np.array(scale)
            print(f"📐 杆塔{i}: 原始尺寸{ext} -> 增强尺寸{big}")
            corners, lines = obb_wireframe(tower_info['center'], tower_info['rotation'], big)
            pts = [corners[j] for line in lines for j in line]
            geoms.append((np.array(pts), line_color))
            print(f"✅ 杆塔{i}处理成功，中心：{tower_info['center']}")
        except Exception as e:
            print(f"⚠️ 杆塔{i}可视化失败: {str(e)}")
            continue
    print(f"✅ 成功处理 {len(geoms)} 个杆塔几何体")
    return full_pcd, geoms


def extract_and_visualize_towers(las_path: str, tower_obbs: list,
                                 scale_factors: list = None,
                                 line_color: tuple = (1.0, 0.0, 0.0),
                                 adaptive_scaling: bool = True,
                                 use_kuangxuan_method: bool = True,
                                 kuangxuan_preset: str = "kuangxuan_original"):
    if use_kuangxuan_method:
        method, params = get_bbox_preset(kuangxuan_preset)
        return extract_and_visualize_towers_kuangxuan(las_path, tower_obbs, method, params, line_color)
    return extract_and_visualize_towers_original(las_path, tower_obbs, scale_factors, line_color, adaptive_scaling)
