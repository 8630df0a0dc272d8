"""Drop-in for the reference's utils/elevation_converter.py.

Same class/method surface and fallback: the transformer is PROJ's
``+proj=vgridshift +grids=egm08_25.gtx +multiplier=1`` (utils/elevation_converter.py:29-31), i.e.
``h + N``; when the grid cannot be found everything falls back to ``h - region_n_value``
(:33-35,50-55).  The grid lookup and bilinear interpolation run on the GPU (..geo); convert_batch is
one kernel launch over the whole array instead of a Python loop (:65-68) with identical results.
Extra keyword options: ``grid_name`` (default the reference's "egm08_25.gtx"; pass "egm96_15.gtx" or a
path to use the shipped EGM96 grid) and ``multiplier`` (default +1 like the reference; crs.py uses -1).
"""
import os

import numpy as np

from .. import geo as _geo


class ElevationConverter:
    """椭球高到正高的转换器"""

    def __init__(self, region_n_value=25.0, grid_name="egm08_25.gtx", multiplier=1.0):
        self.region_n_value = region_n_value
        self.grid_name = grid_name
        self.multiplier = multiplier
        self.transformer = None
        self.init_transformer()

    def init_transformer(self):
        try:
            path = _geo.find_grid_file(self.grid_name)
            if path is None:
                raise FileNotFoundError(f"geoid grid {self.grid_name} not found")
            self.transformer = _geo.load_grid(path)
            print("✅ EGM2008转换器初始化成功")
        except Exception as e:
            print(f"⚠️ EGM2008转换器初始化失败，将使用经验值: {str(e)}")
            self.transformer = None

    def ellipsoid_to_orthometric(self, lat, lon, ellipsoid_height):
        try:
            if self.transformer:
                out = _geo.geoid_shift(self.transformer, [lat], [lon], [ellipsoid_height], self.multiplier)
                return float(out.cpu().numpy()[0])
            else:
                return ellipsoid_height - self.region_n_value
        except Exception as e:
            print(f"高程转换失败，使用经验值: {str(e)}")
            return ellipsoid_height - self.region_n_value

    def convert_batch(self, lat_array, lon_array, ellipsoid_heights):
        lat = np.asarray(lat_array, dtype=np.float64).reshape(-1)
        lon = np.asarray(lon_array, dtype=np.float64).reshape(-1)
        h = np.asarray(ellipsoid_heights, dtype=np.float64).reshape(-1)
        n = min(len(lat), len(lon), len(h))  # zip() semantics of the reference loop
        if n == 0:
            return np.array([])
        try:
            if self.transformer:
                out = _geo.geoid_shift(self.transformer, lat[:n], lon[:n], h[:n], self.multiplier)
                return out.cpu().numpy()
        except Exception as e:
            print(f"高程转换失败，使用经验值: {str(e)}")
        return h[:n] - self.region_n_value


def convert_elevation(lat, lon, ellipsoid_height, region_n_value=25.0):
    converter = ElevationConverter(region_n_value)
    return converter.ellipsoid_to_orthometric(lat, lon, ellipsoid_height)
