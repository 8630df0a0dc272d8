"""Oracle for the tower-level match: utils/table_match_gim.py:17-34 (haversine),
:37-142 (per-tower CRS + geoid conversion) and :145-196 (match_towers' greedy double loop), restated with
the oracle's CRS/geoid functions.  Pinned by tests/golden/reference_run.json["match"], produced by the
unmodified reference function."""
import math

import numpy as np

from . import crs, geoid


def haversine(lat1, lon1, lat2, lon2):
    R = 6371.0
    lat1, lon1, lat2, lon2 = map(math.radians, [lat1, lon1, lat2, lon2])
    a = math.sin((lat2 - lat1) / 2) ** 2 + math.cos(lat1) * math.cos(lat2) * math.sin((lon2 - lon1) / 2) ** 2
    return R * 2 * math.atan2(math.sqrt(a), math.sqrt(1 - a)) * 1000


def match_towers(gim_list, pc_towers, grid=None, distance_threshold=50, height_threshold=100, region_n_value=25.0):
    conv = []
    for t in pc_towers:
        lon, lat = crs.gk_inverse(t["center"][0], t["center"][1])
        H = float(geoid.ellipsoid_to_orthometric(grid, float(lat), float(lon), t["center"][2], region_n_value))
        conv.append([float(lon), float(lat), H])
    matched = []
    for i, g in enumerate(gim_list):
        for j, c in enumerate(conv):
            if haversine(g.get("lat", 0), g.get("lng", 0), c[1], c[0]) <= distance_threshold and \
                    abs(g.get("h", 0) - c[2]) <= height_threshold:
                matched.append((i, j))
                break
    return matched, conv
