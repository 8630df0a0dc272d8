"""Deterministic synthetic transmission-corridor LAS generator.

The reference ships no data (``*.las`` is git-ignored) so every test/bench input is made here:
LAS 1.2, PDRF 3 (34-byte records), scale 0.001 m, offsets (437000, 3139000, 0) — the EPSG:4547
neighbourhood of the reference's recorded run (test/kuangxuan.py:29-33) — a 60 m wide corridor at
azimuth 17 degrees, one lattice tower every 350 m, points emitted in flight (along-axis) order.
Host numpy only; this is workload generation, not part of the measured path.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, Optional, Tuple

import numpy as np

from . import las as _las

SCALES = np.array([0.001, 0.001, 0.001])
OFFSETS = np.array([437000.0, 3139000.0, 0.0])
ORIGIN_EN = (437500.0, 3139500.0)
AZIMUTH_DEG = 17.0
SPAN = 350.0
HALF_WIDTH = 30.0

PDRF3 = np.dtype([("X", "<i4"), ("Y", "<i4"), ("Z", "<i4"), ("intensity", "<u2"), ("flags", "u1"),
                  ("classification", "u1"), ("scan_angle", "i1"), ("user_data", "u1"),
                  ("point_source_id", "<u2"), ("gps_time", "<f8"), ("red", "<u2"), ("green", "<u2"),
                  ("blue", "<u2")])
assert PDRF3.itemsize == 34


def terrain_height(s: np.ndarray, terrain: str) -> np.ndarray:
    if terrain == "flat":
        return np.full_like(s, 80.0, dtype=np.float64)
    if terrain == "hilly":
        return 80.0 + 40.0 * np.sin(2 * np.pi * s / 3000.0) + 15.0 * np.sin(2 * np.pi * s / 700.0 + 1.0)
    raise ValueError(f"unknown terrain {terrain!r}")


def _tower_params(n_towers: int, seed: int):
    rng = np.random.default_rng([seed, 0x70])
    # virtual towers at both ends so every block has two spans to hang conductors from
    heights = rng.uniform(25.0, 45.0, n_towers + 2)
    bases = rng.uniform(8.0, 14.0, n_towers + 2)
    return heights, bases


def _tower_points(rng, n: int, height: float, base: float):
    """Local (ds, dt, dz) samples of a square-base lattice tower: 4 tapered legs, X-bracing on the
    4 faces, 3 cross-arms."""
    kind = rng.random(n)
    v = rng.random(n)
    ds = np.empty(n)
    dt = np.empty(n)
    dz = np.empty(n)
    hw = lambda vv: 0.5 * base * (1.0 - 0.9 * vv)
    legs = kind < 0.4
    brace = (kind >= 0.4) & (kind < 0.8)
    arms = kind >= 0.8
    # legs
    k = int(legs.sum())
    corner = rng.integers(0, 4, k)
    sx = np.where(corner & 1, 1.0, -1.0)
    sy = np.where(corner & 2, 1.0, -1.0)
    w = hw(v[legs])
    ds[legs], dt[legs], dz[legs] = sx * w, sy * w, v[legs] * height
    # X bracing: 6 panels per face, a diagonal from one leg at panel bottom to the other at panel top
    k = int(brace.sum())
    panels = 6
    p = rng.integers(0, panels, k)
    face = rng.integers(0, 4, k)
    flip = rng.integers(0, 2, k) * 2.0 - 1.0
    u = v[brace]
    vv = (p + u) / panels
    w = hw(vv)
    lateral = flip * (2.0 * u - 1.0) * w
    fixed = np.where(face & 1, 1.0, -1.0) * w
    along_s = (face & 2) == 0
    ds[brace] = np.where(along_s, lateral, fixed)
    dt[brace] = np.where(along_s, fixed, lateral)
    dz[brace] = vv * height
    # cross-arms across the corridor at 0.6/0.75/0.9 H (8 m reach), earth-wire peak arm at H (4 m)
    k = int(arms.sum())
    level = rng.integers(0, 4, k)
    frac = np.array([0.6, 0.75, 0.9, 1.0])[level]
    reach = np.array([8.0, 8.0, 8.0, 4.0])[level]
    ds[arms] = rng.normal(0.0, 0.15, k)
    dt[arms] = (2.0 * v[arms] - 1.0) * reach
    dz[arms] = frac * height
    noise = rng.normal(0.0, 0.03, (3, n))
    return ds + noise[0], dt + noise[1], dz + noise[2]


def corridor_blocks(n_points: int, n_towers: int = 5, terrain: str = "flat", seed: int = 1,
                    fractions: Tuple[float, float, float, float] = (0.80, 0.08, 0.07, 0.05),
                    s_origin: float = 0.0) -> Iterator[Dict[str, np.ndarray]]:
    """Yield one block per tower span: dict(X, Y, Z int32 lattice coordinates, cls uint8), points
    sorted along the axis inside the block; concatenating the blocks gives the file order."""
    if n_points < 0 or n_towers < 1:
        raise ValueError("n_points >= 0 and n_towers >= 1 required")
    heights, bases = _tower_params(n_towers, seed)
    az = math.radians(AZIMUTH_DEG)
    sa, ca = math.sin(az), math.cos(az)
    per = n_points // n_towers
    fg, fv, fc, ft = fractions
    for k in range(n_towers):
        nb = per if k < n_towers - 1 else n_points - per * (n_towers - 1)
        rng = np.random.default_rng([seed, 1, k])
        n_t = int(round(nb * ft))
        n_c = int(round(nb * fc))
        n_v = int(round(nb * fv))
        n_g = nb - n_t - n_c - n_v
        s0, s1 = s_origin + SPAN * k, s_origin + SPAN * (k + 1)
        st = 0.5 * (s0 + s1)
        parts_s, parts_t, parts_z, parts_c = [], [], [], []
        # ground + low vegetation
        for n, cls in ((n_g, 2), (n_v, 3)):
            s = rng.uniform(s0, s1, n)
            t = rng.uniform(-HALF_WIDTH, HALF_WIDTH, n)
            z = terrain_height(s, terrain)
            z = z + (rng.normal(0.0, 0.05, n) if cls == 2 else rng.uniform(0.0, 2.5, n))
            parts_s.append(s); parts_t.append(t); parts_z.append(z)
            parts_c.append(np.full(n, cls, np.uint8))
        # conductors: 3 phases + 2 earth wires, parabola-approximated catenaries, sag 2 % of span
        if n_c:
            s = rng.uniform(s0, s1, n_c)
            wire = rng.integers(0, 5, n_c)
            t_w = np.array([-8.0, 0.0, 8.0, -4.0, 4.0])[wire]
            f_w = np.array([0.75, 0.75, 0.75, 1.0, 1.0])[wire]
            left = s < st
            sa_ = np.where(left, st - SPAN, st)          # span start tower position
            ka = np.where(left, k, k + 1)                # index into heights (shifted by +1)
            kb = ka + 1
            u = (s - sa_) / SPAN
            za = terrain_height(sa_, terrain) + f_w * heights[ka]
            zb = terrain_height(sa_ + SPAN, terrain) + f_w * heights[kb]
            z = za + (zb - za) * u - 4.0 * (0.02 * SPAN) * u * (1.0 - u) + rng.normal(0.0, 0.02, n_c)
            parts_s.append(s); parts_t.append(t_w + rng.normal(0.0, 0.02, n_c)); parts_z.append(z)
            parts_c.append(np.full(n_c, 14, np.uint8))
        if n_t:
            ds, dt, dz = _tower_points(rng, n_t, heights[k + 1], bases[k + 1])
            parts_s.append(st + ds); parts_t.append(dt)
            parts_z.append(terrain_height(np.array([st]), terrain)[0] + dz)
            parts_c.append(np.full(n_t, 15, np.uint8))
        s = np.concatenate(parts_s); t = np.concatenate(parts_t)
        z = np.concatenate(parts_z); c = np.concatenate(parts_c)
        order = np.argsort(s, kind="stable")
        s, t, z, c = s[order], t[order], z[order], c[order]
        e = ORIGIN_EN[0] + s * sa + t * ca
        nn = ORIGIN_EN[1] + s * ca - t * sa
        X = np.rint((e - OFFSETS[0]) / SCALES[0]).astype(np.int32)
        Y = np.rint((nn - OFFSETS[1]) / SCALES[1]).astype(np.int32)
        Z = np.rint((z - OFFSETS[2]) / SCALES[2]).astype(np.int32)
        yield {"X": X, "Y": Y, "Z": Z, "cls": c}


def tower_ground_truth(n_towers: int, terrain: str = "flat", seed: int = 1, s_origin: float = 0.0) -> np.ndarray:
    """(n_towers, 5) array: easting, northing, base z, height, base width of every generated tower."""
    heights, bases = _tower_params(n_towers, seed)
    az = math.radians(AZIMUTH_DEG)
    st = s_origin + SPAN * (np.arange(n_towers) + 0.5)
    e = ORIGIN_EN[0] + st * math.sin(az)
    n = ORIGIN_EN[1] + st * math.cos(az)
    return np.stack([e, n, terrain_height(st, terrain), heights[1:-1], bases[1:-1]], axis=1)


def corridor_records(n_points: int, n_towers: int = 5, terrain: str = "flat", seed: int = 1,
                     fractions=(0.80, 0.08, 0.07, 0.05), s_origin: float = 0.0,
                     out: Optional[np.ndarray] = None) -> np.ndarray:
    """Raw PDRF-3 records (structured array, 34 B each) of the whole corridor in file order.
    ``out`` may be a preallocated (e.g. pinned) structured/uint8 buffer of n_points*34 bytes."""
    if out is None:
        rec = np.zeros(n_points, dtype=PDRF3)
    else:
        rec = out.view(np.uint8).reshape(-1)[: n_points * 34].view(PDRF3)
        rec[:] = np.zeros((), dtype=PDRF3)
    pos = 0
    for blk in corridor_blocks(n_points, n_towers, terrain, seed, fractions, s_origin):
        n = blk["X"].size
        r = rec[pos:pos + n]
        r["X"], r["Y"], r["Z"] = blk["X"], blk["Y"], blk["Z"]
        r["classification"] = blk["cls"]
        r["intensity"] = (blk["Z"] & 0x3FF).astype(np.uint16) + 100
        r["flags"] = 0x09  # return 1 of 1
        r["gps_time"] = (pos + np.arange(n, dtype=np.float64)) * 1e-5
        pos += n
    assert pos == n_points
    return rec


def corridor_header(records: np.ndarray) -> _las.LasHeader:
    h = _las.LasHeader(version=(1, 2), point_format=3, record_length=34, header_size=227,
                       offset_to_point_data=227, n_vlr=0, point_count=int(records.size),
                       scales=SCALES.copy(), offsets=OFFSETS.copy())
    return h


def write_corridor_las(path: str, n_points: int, n_towers: int = 5, terrain: str = "flat", seed: int = 1,
                       fractions=(0.80, 0.08, 0.07, 0.05)) -> _las.LasHeader:
    rec = corridor_records(n_points, n_towers, terrain, seed, fractions)
    h = corridor_header(rec)
    if n_points:
        mn = np.array([rec["X"].min(), rec["Y"].min(), rec["Z"].min()])
        mx = np.array([rec["X"].max(), rec["Y"].max(), rec["Z"].max()])
    else:
        mn = mx = None
    _las.write_raw(path, h, rec.view(np.uint8), mn, mx)
    return _las.read_header(path)
