"""Qt-free mirror of the numeric part of the reference's utils/table_match_gim.py (the
immediate caller of the geoid/CRS path): ``haversine`` (:17-34),
``convert_pointcloud_ellipsoid_to_orthometric`` (:37-142) and ``match_towers`` (:145-196).

The per-tower conversions run as ONE batched device pass (inverse Gauss-Krueger + geoid shift) and the
GIM x point-cloud distance matrix is one kernel; the reference's control flow (first point-cloud tower
within both thresholds wins, per GIM tower) is reproduced on the host.  The Qt table builders
(match_from_gim_tower_list, correct_from_gim_tower_list, :225-463) are GUI code and stay in the
reference; they call match_towers, which is what this module replaces.

`transformer` may be any object with ``.transform(x, y) -> (lon, lat)`` (a pyproj Transformer) or None,
in which case the package's own EPSG:4547 -> EPSG:4326 kernel is used.
"""
import math

import numpy as np

from .. import geo as _geo
from .. import _native
from .elevation_converter import ElevationConverter


def haversine(lat1, lon1, lat2, lon2):
    """Great-circle distance in metres, R = 6371 km (scalar; utils/table_match_gim.py:17-34)."""
    R = 6371.0
    lat1, lon1, lat2, lon2 = map(math.radians, [lat1, lon1, lat2, lon2])
    dlat = lat2 - lat1
    dlon = lon2 - lon1
    a = math.sin(dlat / 2) ** 2 + math.cos(lat1) * math.cos(lat2) * math.sin(dlon / 2) ** 2
    c = 2 * math.atan2(math.sqrt(a), math.sqrt(1 - a))
    return R * c * 1000


def haversine_matrix(lat1, lon1, lat2, lon2):
    """(n1, n2) distance matrix on the device."""
    import torch
    from .. import device as dv
    dv._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    a = [_geo._as_dev_f64(v, dev) for v in (lat1, lon1, lat2, lon2)]
    n1, n2 = a[0].numel(), a[2].numel()
    out = torch.empty((n1, n2), dtype=torch.float64, device=dev)
    _native.check(_native.lib().pch_haversine_matrix(a[0].data_ptr(), a[1].data_ptr(), n1, a[2].data_ptr(),
                                                     a[3].data_ptr(), n2, out.data_ptr(), dv._stream()),
                  "pch_haversine_matrix")
    return out


def convert_pointcloud_ellipsoid_to_orthometric(pointcloud_towers, transformer=None, region_n_value=25.0):
    print("🔄 开始将点云杆塔高程从椭球高转换为正高...")
    print(f"📍 原始点云杆塔数量: {len(pointcloud_towers)}")
    elev_converter = ElevationConverter(region_n_value=region_n_value)
    converted_towers = []
    if not pointcloud_towers:
        return converted_towers
    centers = np.array([np.asarray(t['center'], dtype=np.float64)[:3] for t in pointcloud_towers])
    if transformer is None:
        lon, lat = _geo.gk_inverse(centers[:, 0], centers[:, 1])
        lon, lat = lon.cpu().numpy(), lat.cpu().numpy()
    else:
        lon, lat = transformer.transform(centers[:, 0], centers[:, 1])
        lon, lat = np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)
    ortho = elev_converter.convert_batch(lat, lon, centers[:, 2])
    for i, tower in enumerate(pointcloud_towers):
        h = float(centers[i, 2])
        H = float(ortho[i])
        converted_towers.append({
            'id': f"PC-{i + 1}",
            'converted_center': [float(lon[i]), float(lat[i]), H],
            'height': tower.get('height', 0),
            'north_angle': tower.get('north_angle', 0),
            'original_center': tower['center'],
            'ellipsoid_height': h,
            'orthometric_height': H,
            'n_value': h - H,
            'height_conversion_applied': True,
        })
        print(f"📊 杆塔{i + 1}: 椭球高 {h:.2f}m → 正高 {H:.2f}m (N={h - H:.2f}m)")
    print(f"✅ 点云杆塔高程转换完成，共处理 {len(converted_towers)} 个杆塔")
    return converted_towers


def match_towers(gim_list, pointcloud_towers, transformer=None, distance_threshold=50, height_threshold=100,
                 region_n_value=25.0):
    """[(gim_index, pc_index)], converted towers — first point-cloud tower within both thresholds wins."""
    print("🔍 开始杆塔匹配（在匹配阶段进行高程转换）...")
    converted_towers = convert_pointcloud_ellipsoid_to_orthometric(pointcloud_towers, transformer, region_n_value)
    matched_rows = []
    if not gim_list or not converted_towers:
        return matched_rows, converted_towers
    g_lat = np.array([t.get("lat", 0) for t in gim_list], dtype=np.float64)
    g_lon = np.array([t.get("lng", 0) for t in gim_list], dtype=np.float64)
    g_h = np.array([t.get("h", 0) for t in gim_list], dtype=np.float64)
    p_lon = np.array([t['converted_center'][0] for t in converted_towers])
    p_lat = np.array([t['converted_center'][1] for t in converted_towers])
    p_h = np.array([t['converted_center'][2] for t in converted_towers])
    dist = haversine_matrix(g_lat, g_lon, p_lat, p_lon).cpu().numpy()
    ok = (dist <= distance_threshold) & (np.abs(g_h[:, None] - p_h[None, :]) <= height_threshold)
    for i in range(len(gim_list)):
        js = np.nonzero(ok[i])[0]
        if js.size:
            matched_rows.append((i, int(js[0])))
            print(f"  ✅ 匹配成功！GIM杆塔{i + 1} ↔ 点云杆塔{int(js[0]) + 1}")
    print(f"🎉 匹配完成，共找到 {len(matched_rows)} 对匹配的杆塔")
    return matched_rows, converted_towers
