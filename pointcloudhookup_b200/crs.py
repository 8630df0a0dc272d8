"""Importable counterpart of the reference's crs.py script (which runs at import and exit(1)s when a
hard-coded Windows path is missing, crs.py:6-13).  Two conversions, both on the GPU:

  ellipsoid_to_orthometric_egm96(lon, lat, h)  crs.py:25-35 pipeline: unitconvert deg->rad +
      vgridshift(egm96_15.gtx, multiplier=-1)  ->  H = h - N_EGM96
  cgcs2000_gk114_to_wgs84(x, y)  Transformer.from_crs("EPSG:4547","EPSG:4326",always_xy=True)
      (utils/table_match_gim.py:232; test/005test.py:37)  ->  (lon, lat) degrees
"""
import numpy as np

from . import geo as _geo

EGM96_GRID = "egm96_15.gtx"


def _grid(grid_path=None):
    path = grid_path or _geo.find_grid_file(EGM96_GRID)
    if path is None:
        raise FileNotFoundError(f"未找到 geoid 网格文件 {EGM96_GRID}")
    return _geo.load_grid(path)


def ellipsoid_to_orthometric_egm96(lon, lat, h, grid_path=None):
    out = _geo.geoid_shift(_grid(grid_path), lat, lon, h, multiplier=-1.0).cpu().numpy()
    return out if np.ndim(lon) else float(out[0])


def cgcs2000_gk114_to_wgs84(x, y):
    lon, lat = _geo.gk_inverse(x, y, _geo.EPSG4547)
    lon, lat = lon.cpu().numpy(), lat.cpu().numpy()
    return (lon, lat) if np.ndim(x) else (float(lon[0]), float(lat[0]))


def convert_las_coordinates(input_path, grid_path=None, multiplier=-1.0):
    """Per-point EPSG:4547 -> 4326 (+ orthometric height) of a whole LAS (test/005test.py:8-106):
    returns an (n,3) float64 array [lon, lat, H]."""
    from . import device as dv, las as _las
    hdr, rec = _las.read_raw(input_path)
    dl = dv.upload_records_xyz(rec, hdr.point_count, hdr.record_length, hdr.scales, hdr.offsets)
    return _geo.las_to_geodetic(dl, _grid(grid_path), multiplier, _geo.EPSG4547).cpu().numpy()


# the four towers of crs.py:16-21
CRS_PY_TOWERS = {
    "编号": ["P142", "P143", "P144", "P145"],
    "纬度": [28.379743, 28.376914, 28.373484, 28.369953],
    "经度": [113.363246, 113.364204, 113.365366, 113.366563],
    "椭球高": [104.03, 70.52, 69.68, 67.15],
}

if __name__ == "__main__":
    d = CRS_PY_TOWERS
    H = ellipsoid_to_orthometric_egm96(np.array(d["经度"]), np.array(d["纬度"]), np.array(d["椭球高"]))
    print("\n=== 高程转换结果 ===")
    print("编号    纬度        经度         椭球高    正高      N值")
    for name, la, lo, h, hh in zip(d["编号"], d["纬度"], d["经度"], d["椭球高"], H):
        print(f"{name}  {la:.6f}  {lo:.6f}  {h:7.2f}  {round(hh, 3):8.3f}  {h - round(hh, 3):7.3f}")
