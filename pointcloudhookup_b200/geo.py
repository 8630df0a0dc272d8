"""Geoid grids and CRS constants (host set-up) + device calls for the per-point conversions.

Host work here is O(1) per grid / per projection: parsing the GTX header, byte-swapping the grid
once, computing the Krueger series constants.  Every per-point evaluation is a CUDA kernel
(pch_geoid_shift, pch_gk_inverse, pch_las_geodetic).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import os
import struct
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native, device as dv
from ._native import GeoidGrid, TmParams, check

_HERE = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(_HERE, "data")


@dataclasses.dataclass
class HostGrid:
    ll_lat: float
    ll_lon: float
    dlat: float
    dlon: float
    grid: np.ndarray  # (rows, cols) float32, rows south -> north

    @property
    def rows(self):
        return self.grid.shape[0]

    @property
    def cols(self):
        return self.grid.shape[1]

    @property
    def is_global(self):
        return self.cols * self.dlon >= 360.0 - 1e-9


def read_gtx(path: str) -> HostGrid:
    """NOAA/PROJ GTX: 4 BE float64 (ll_lat, ll_lon, dlat, dlon), 2 BE int32 (rows, cols), BE float32 grid."""
    with open(path, "rb") as f:
        raw = f.read()
    ll_lat, ll_lon, dlat, dlon = struct.unpack_from(">4d", raw, 0)
    rows, cols = struct.unpack_from(">2i", raw, 32)
    if rows < 2 or cols < 2 or len(raw) < 40 + rows * cols * 4:
        raise ValueError(f"{path}: not a valid GTX grid")
    g = np.frombuffer(raw, dtype=">f4", count=rows * cols, offset=40).reshape(rows, cols).astype(np.float32)
    return HostGrid(ll_lat, ll_lon, dlat, dlon, g)


def read_npz_grid(path: str) -> HostGrid:
    """lat/lon/geoid arrays as shipped in egm2008_simulated_0.25deg.npz, cast to float32 nodes."""
    d = np.load(path)
    lat, lon, g = d["lat"], d["lon"], d["geoid"]
    if lat[0] > lat[-1]:
        lat, g = lat[::-1], g[::-1]
    return HostGrid(float(lat[0]), float(lon[0]), float(lat[1] - lat[0]), float(lon[1] - lon[0]),
                    np.ascontiguousarray(g, dtype=np.float32))


@dataclasses.dataclass
class DeviceGrid:
    desc: GeoidGrid
    data: torch.Tensor  # float32 [rows*pitch]
    host: HostGrid


def upload_grid(hg: HostGrid, device=None) -> DeviceGrid:
    dv._require_cuda()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    pitch = (hg.cols + 3) // 4 * 4
    padded = np.zeros((hg.rows, pitch), dtype=np.float32)
    padded[:, :hg.cols] = hg.grid
    data = torch.from_numpy(padded.reshape(-1)).to(device)
    desc = GeoidGrid(hg.ll_lat, hg.ll_lon, hg.dlat, hg.dlon, hg.rows, hg.cols, pitch, int(hg.is_global))
    return DeviceGrid(desc, data, hg)


_grid_cache = {}


def find_grid_file(name: str) -> Optional[str]:
    """Look for a geoid grid the way PROJ would (PROJ_LIB / PROJ_DATA) plus this package's data dir
    and the current directory."""
    cands = [name, os.path.join(DATA_DIR, name)]
    for env in ("PCH_GEOID_DIR", "PROJ_DATA", "PROJ_LIB"):
        if os.environ.get(env):
            cands.append(os.path.join(os.environ[env], name))
    for c in cands:
        if os.path.isfile(c):
            return c
    return None


def load_grid(path: str, device=None) -> DeviceGrid:
    key = (os.path.abspath(path), str(device))
    if key not in _grid_cache:
        hg = read_npz_grid(path) if path.endswith(".npz") else read_gtx(path)
        _grid_cache[key] = upload_grid(hg, device)
    return _grid_cache[key]


# ---------------------------------------------------------------------------------------------
def tm_params(a: float = 6378137.0, inv_f: float = 298.257222101, lon0_deg: float = 114.0, k0: float = 1.0,
              fe: float = 500000.0, fn: float = 0.0) -> TmParams:
    """Krueger n-series constants; defaults = EPSG:4547 (CGCS2000 / 3-degree Gauss-Kruger CM 114E)."""
    f = 1.0 / inv_f
    n = f / (2.0 - f)
    n2, n3, n4, n5, n6 = n * n, n ** 3, n ** 4, n ** 5, n ** 6
    rect = a / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0)
    beta = [
        n / 2.0 - 2.0 * n2 / 3.0 + 37.0 * n3 / 96.0 - n4 / 360.0 - 81.0 * n5 / 512.0 + 96199.0 * n6 / 604800.0,
        n2 / 48.0 + n3 / 15.0 - 437.0 * n4 / 1440.0 + 46.0 * n5 / 105.0 - 1118711.0 * n6 / 3870720.0,
        17.0 * n3 / 480.0 - 37.0 * n4 / 840.0 - 209.0 * n5 / 4480.0 + 5569.0 * n6 / 90720.0,
        4397.0 * n4 / 161280.0 - 11.0 * n5 / 504.0 - 830251.0 * n6 / 7257600.0,
        4583.0 * n5 / 161280.0 - 108847.0 * n6 / 3991680.0,
        20648693.0 * n6 / 638668800.0,
    ]
    return TmParams(rect, (C.c_double * 6)(*beta), math.sqrt(f * (2.0 - f)), lon0_deg, k0, fe, fn)


EPSG4547 = tm_params()


def _as_dev_f64(a, device) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))).to(device)


def gk_inverse(x, y, tm: TmParams = EPSG4547, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(lon_deg, lat_deg) device tensors for projected (x=easting, y=northing)."""
    dv._require_cuda()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    xd, yd = _as_dev_f64(x, device), _as_dev_f64(y, device)
    lon, lat = torch.empty_like(xd), torch.empty_like(xd)
    check(_native.lib().pch_gk_inverse(xd.data_ptr(), yd.data_ptr(), xd.numel(), C.byref(tm), lon.data_ptr(),
                                       lat.data_ptr(), dv._stream()), "pch_gk_inverse")
    return lon, lat


def geoid_shift(grid: DeviceGrid, lat, lon, h, multiplier: float = -1.0, want_n: bool = False):
    """PROJ vgridshift forward on device arrays: h + multiplier*N (and optionally N)."""
    device = grid.data.device
    la, lo, hh = _as_dev_f64(lat, device), _as_dev_f64(lon, device), _as_dev_f64(h, device)
    out = torch.empty_like(hh)
    nn = torch.empty_like(hh) if want_n else None
    check(_native.lib().pch_geoid_shift(la.data_ptr(), lo.data_ptr(), hh.data_ptr(), hh.numel(), grid.data.data_ptr(),
                                        C.byref(grid.desc), float(multiplier), out.data_ptr(), dv._ptr(nn),
                                        dv._stream()), "pch_geoid_shift")
    return (out, nn) if want_n else out


def grid_window(grid: DeviceGrid, lat_min, lat_max, lon_min, lon_max, max_bytes=64 * 1024):
    """Grid window (row0, col0, rows, cols) covering the bounding box (+1 node margin), columns
    aligned to 4 floats; (0,0,0,0) when it would wrap the seam or exceed `max_bytes`."""
    d = grid.desc
    r0 = int(math.floor((lat_min - d.ll_lat) / d.dlat)) - 1
    r1 = int(math.floor((lat_max - d.ll_lat) / d.dlat)) + 2
    dl0 = (lon_min - d.ll_lon) % 360.0
    dl1 = (lon_max - d.ll_lon) % 360.0
    if dl1 < dl0:
        return 0, 0, 0, 0
    c0 = int(math.floor(dl0 / d.dlon)) - 1
    c1 = int(math.floor(dl1 / d.dlon)) + 2
    r0, c0 = max(r0, 0), max(c0, 0) // 4 * 4
    r1 = min(r1, d.rows - 1)
    c1 = min((c1 + 4) // 4 * 4, d.pitch)
    rows, cols = r1 - r0 + 1, c1 - c0
    if rows <= 0 or cols <= 0 or rows * cols * 4 > max_bytes:
        return 0, 0, 0, 0
    return r0, c0, rows, cols


def las_to_geodetic(dl: dv.DeviceLas, grid: DeviceGrid, multiplier: float = -1.0, tm: Optional[TmParams] = EPSG4547,
                    window="auto", chunk_minmax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(n,3) float64 [lon, lat, h + multiplier*N] for every LAS point, one fused kernel.  `chunk_minmax`: the
    (n_chunks,6) lattice extrema a voxel stage already computed for these records (VoxelResult.chunk_minmax); the
    grid window is then derived from it instead of from another pass over the records."""
    out = torch.empty((dl.n, 3), dtype=torch.float64, device=dl.device)
    if dl.n == 0:
        return out
    if window == "auto":
        if chunk_minmax is not None:
            mm = torch.cat([chunk_minmax[:, :3].amin(0), chunk_minmax[:, 3:].amax(0)]).cpu().numpy().astype(np.float64)
        else:
            mm = dv.chunk_minmax(dl, dl.n).cpu().numpy()[0].astype(np.float64)
        xs = mm[[0, 3]] * dl.scales[0] + dl.offsets[0]
        ys = mm[[1, 4]] * dl.scales[1] + dl.offsets[1]
        cx = np.array([xs[0], xs[0], xs[1], xs[1], xs.mean(), xs.mean(), xs[0], xs[1]])
        cy = np.array([ys[0], ys[1], ys[0], ys[1], ys[0], ys[1], ys.mean(), ys.mean()])
        if tm is not None:
            lo, la = gk_inverse(cx, cy, tm, dl.device)
            lo, la = lo.cpu().numpy(), la.cpu().numpy()
        else:
            lo, la = cx, cy
        window = grid_window(grid, la.min(), la.max(), lo.min(), lo.max())
    elif window is None:
        window = (0, 0, 0, 0)
    r0, c0, rows, cols = window
    check(_native.lib().pch_las_geodetic(dl.rec.data_ptr(), dl.n, dl.rec_len, _native.d3(dl.scales),
                                         _native.d3(dl.offsets), C.byref(tm) if tm is not None else None,
                                         grid.data.data_ptr(), C.byref(grid.desc), r0, c0, rows, cols,
                                         float(multiplier), out.data_ptr(), dv._stream()), "pch_las_geodetic")
    return out
