"""Oracle geoid shift: PROJ ``+proj=vgridshift`` on a GTX grid, as the reference calls it.

Call sites: utils/elevation_converter.py:29-31,48 (``+multiplier=1`` -> h + N; falls back to
h - region_n_value when the grid cannot be loaded, :33-35,50-55) and crs.py:25-35
(``+multiplier=-1`` -> H = h - N, rounded to 3 dp).  pyproj/PROJ are absent and the reference
records no expected outputs -> PARITY UNPINNED.  Restated per SURVEY.md Appendix A.6: bilinear
interpolation in float64 of the four float32 nodes, longitude wrapped into [ll_lon, ll_lon+360),
east neighbour of the last column wraps to column 0 on a global grid, north neighbour clamps.
"""
import struct
import numpy as np

NODATA = -88.8888


def read_gtx(path):
    with open(path, "rb") as f:
        raw = f.read()
    ll_lat, ll_lon, dlat, dlon = struct.unpack_from(">4d", raw, 0)
    rows, cols = struct.unpack_from(">2i", raw, 32)
    g = np.frombuffer(raw, dtype=">f4", count=rows * cols, offset=40).reshape(rows, cols).astype(np.float32)
    return {"ll_lat": ll_lat, "ll_lon": ll_lon, "dlat": dlat, "dlon": dlon, "rows": rows, "cols": cols, "grid": g}


def grid_from_npz(path):
    """The shipped egm2008_simulated_0.25deg.npz (lat(721), lon(1441), geoid(721,1441) f64) as a
    GTX-like float32 grid (rows south->north)."""
    d = np.load(path)
    lat, lon, g = d["lat"], d["lon"], d["geoid"]
    if lat[0] > lat[-1]:
        lat, g = lat[::-1], g[::-1]
    return {"ll_lat": float(lat[0]), "ll_lon": float(lon[0]), "dlat": float(lat[1] - lat[0]),
            "dlon": float(lon[1] - lon[0]), "rows": g.shape[0], "cols": g.shape[1],
            "grid": np.ascontiguousarray(g, dtype=np.float32)}


def geoid_height(grid, lat, lon):
    """N(lat, lon) in float64; NaN outside the grid or where all four corners are nodata."""
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    g, rows, cols = grid["grid"], grid["rows"], grid["cols"]
    span = cols * grid["dlon"]
    is_global = span >= 360.0 - 1e-9
    dl = lon - grid["ll_lon"]
    dl = dl - np.floor(dl / 360.0) * 360.0
    gx = dl / grid["dlon"]
    gy = (lat - grid["ll_lat"]) / grid["dlat"]
    ix = np.floor(gx).astype(np.int64)
    iy = np.floor(gy).astype(np.int64)
    fx = gx - ix
    fy = gy - iy
    bad = (gy < 0) | (gy > rows - 1) | ((not is_global) & (gx > cols - 1))
    iy = np.clip(iy, 0, rows - 1)
    ix = np.clip(ix, 0, cols - 1) if not is_global else ix % cols
    ix2 = ix + 1
    ix2 = np.where(ix2 >= cols, 0 if is_global else cols - 1, ix2)
    iy2 = np.minimum(iy + 1, rows - 1)
    g00 = g[iy, ix].astype(np.float64); g01 = g[iy, ix2].astype(np.float64)
    g10 = g[iy2, ix].astype(np.float64); g11 = g[iy2, ix2].astype(np.float64)
    w00 = (1.0 - fx) * (1.0 - fy); w01 = fx * (1.0 - fy); w10 = (1.0 - fx) * fy; w11 = fx * fy
    # PROJ's nodata rule (grids.cpp, vertical grid value): corners equal to the nodata value are left out, the
    # sum is divided by the total weight of the valid corners, and only a cell without any valid corner has no
    # value (PROJ: HUGE_VAL and an error; NaN here).
    nd = np.float32(NODATA)
    n = np.zeros(np.broadcast(gx, gy).shape, dtype=np.float64)
    tw = np.zeros_like(n)
    nw = np.zeros(n.shape, dtype=np.int64)
    for w, gv, fv in ((w00, g00, g[iy, ix]), (w01, g01, g[iy, ix2]), (w10, g10, g[iy2, ix]), (w11, g11, g[iy2, ix2])):
        ok = fv != nd
        n = np.where(ok, n + w * gv, n)
        tw = np.where(ok, tw + w, tw)
        nw = nw + ok
    with np.errstate(divide="ignore", invalid="ignore"):
        n = np.where(nw == 4, n, n / tw)
    n = np.where(bad | (nw == 0), np.nan, n)
    return n


def vgridshift(grid, lon, lat, h, multiplier=-1.0):
    """PROJ forward: z_out = z_in + multiplier * N (PROJ's default multiplier is -1)."""
    return np.asarray(h, dtype=np.float64) + multiplier * geoid_height(grid, lat, lon)


def ellipsoid_to_orthometric(grid, lat, lon, h, region_n_value=25.0):
    """ElevationConverter.ellipsoid_to_orthometric semantics (multiplier=+1; no grid -> h - 25)."""
    if grid is None:
        return np.asarray(h, dtype=np.float64) - region_n_value
    return vgridshift(grid, lon, lat, h, multiplier=1.0)
