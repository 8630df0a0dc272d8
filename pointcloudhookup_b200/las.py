"""Host-side LAS container handling (header parse / header build / raw record I/O).

Only the *container* is handled on the host: the per-point arithmetic the reference gets from
``laspy`` (``las.x = X*scale+offset``, ``las.x = arr`` -> ``round((arr-offset)/scale)``) runs in the
CUDA kernels of ``csrc/`` on the raw record bytes.  This module never decodes a coordinate.

Reference call sites this replaces (all through the absent third-party ``laspy``):
  ui/import_PC.py:28,35-40,61-65      laspy.read / LasHeader / LasData.write
  utils/tower_extraction.py:60-70     laspy.open().read(), header scales/offsets/point_format/version
  utils/tower_extraction.py:243-257   _save_tower_las
Layout facts are those of the public LAS 1.2/1.4 specification.
"""
from __future__ import annotations

import dataclasses
import os
import struct
from typing import Optional, Tuple

import numpy as np

# nominal record length per point data record format (LAS 1.4 R15 table)
PDRF_LENGTH = {0: 20, 1: 28, 2: 26, 3: 34, 4: 57, 5: 63, 6: 30, 7: 36, 8: 38, 9: 59, 10: 67}


class LasError(ValueError):
    pass


@dataclasses.dataclass
class LasHeader:
    version: Tuple[int, int] = (1, 2)
    point_format: int = 3
    record_length: int = 34
    header_size: int = 227
    offset_to_point_data: int = 227
    n_vlr: int = 0
    point_count: int = 0
    scales: np.ndarray = dataclasses.field(default_factory=lambda: np.array([0.01, 0.01, 0.01]))
    offsets: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(3))
    maxs: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(3))
    mins: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(3))

    def copy(self) -> "LasHeader":
        return dataclasses.replace(self, scales=np.array(self.scales, dtype=np.float64),
                                   offsets=np.array(self.offsets, dtype=np.float64),
                                   maxs=np.array(self.maxs, dtype=np.float64),
                                   mins=np.array(self.mins, dtype=np.float64))


def parse_header(buf: bytes) -> LasHeader:
    if len(buf) < 227 or buf[0:4] != b"LASF":
        raise LasError("not a LAS file (missing LASF signature)")
    vmaj, vmin = buf[24], buf[25]
    header_size, = struct.unpack_from("<H", buf, 94)
    off_pts, n_vlr = struct.unpack_from("<II", buf, 96)
    pfmt_raw = buf[104]
    if pfmt_raw & 0x80 or pfmt_raw & 0x40:
        raise LasError("LAZ-compressed point data is not supported (LAS only)")
    pfmt = pfmt_raw & 0x3F
    rec_len, = struct.unpack_from("<H", buf, 105)
    legacy_count, = struct.unpack_from("<I", buf, 107)
    scales = np.array(struct.unpack_from("<3d", buf, 131))
    offsets = np.array(struct.unpack_from("<3d", buf, 155))
    mx, mnx, my, mny, mz, mnz = struct.unpack_from("<6d", buf, 179)
    count = legacy_count
    if (vmaj, vmin) >= (1, 4) and len(buf) >= 255:
        c64, = struct.unpack_from("<Q", buf, 247)
        if c64:
            count = c64
    if pfmt not in PDRF_LENGTH:
        raise LasError(f"unsupported point data record format {pfmt}")
    if rec_len < PDRF_LENGTH[pfmt]:
        raise LasError(f"record length {rec_len} shorter than PDRF {pfmt} minimum")
    return LasHeader(version=(vmaj, vmin), point_format=pfmt, record_length=rec_len,
                     header_size=header_size, offset_to_point_data=off_pts, n_vlr=n_vlr,
                     point_count=int(count), scales=scales, offsets=offsets,
                     maxs=np.array([mx, my, mz]), mins=np.array([mnx, mny, mnz]))


def read_header(path: str) -> LasHeader:
    with open(path, "rb") as f:
        return parse_header(f.read(375))


def read_raw(path: str, mmap: bool = True) -> Tuple[LasHeader, np.ndarray]:
    """Return (header, uint8 array of shape (point_count*record_length,)) — the undecoded records.

    The array starts at ``offset_to_point_data`` so that record ``i`` begins at byte
    ``i*record_length`` (the alignment contract of ``pch_las_*`` in include/pch_b200.h).
    """
    if not os.path.exists(path):
        raise FileNotFoundError(f"输入文件不存在: {os.path.abspath(path)}")
    hdr = read_header(path)
    nbytes = hdr.point_count * hdr.record_length
    fsize = os.path.getsize(path)
    if hdr.offset_to_point_data + nbytes > fsize:
        raise LasError("LAS file truncated: header promises more point records than the file holds")
    if nbytes == 0:
        return hdr, np.zeros(0, dtype=np.uint8)
    if mmap:
        rec = np.memmap(path, dtype=np.uint8, mode="r", offset=hdr.offset_to_point_data, shape=(nbytes,))
    else:
        with open(path, "rb") as f:
            f.seek(hdr.offset_to_point_data)
            rec = np.frombuffer(f.read(nbytes), dtype=np.uint8)
    return hdr, rec


def new_header_like(src: LasHeader) -> LasHeader:
    """What ``laspy.LasHeader(point_format=..., version=...)`` + copied scales/offsets produces
    (ui/import_PC.py:35-40): same format/version/scales/offsets, no VLRs, nominal record length."""
    vmaj, vmin = src.version
    hsize = 375 if (vmaj, vmin) >= (1, 4) else (235 if (vmaj, vmin) == (1, 3) else 227)
    return LasHeader(version=src.version, point_format=src.point_format,
                     record_length=PDRF_LENGTH[src.point_format], header_size=hsize,
                     offset_to_point_data=hsize, n_vlr=0, point_count=0,
                     scales=np.array(src.scales, dtype=np.float64),
                     offsets=np.array(src.offsets, dtype=np.float64))


def header_bytes(h: LasHeader) -> bytes:
    vmaj, vmin = h.version
    buf = bytearray(h.header_size)
    buf[0:4] = b"LASF"
    buf[24], buf[25] = vmaj, vmin
    buf[26:58] = b"OTHER".ljust(32, b"\0")
    buf[58:90] = b"pointcloudhookup_b200".ljust(32, b"\0")
    struct.pack_into("<HH", buf, 90, 1, 2026)
    struct.pack_into("<H", buf, 94, h.header_size)
    struct.pack_into("<II", buf, 96, h.offset_to_point_data, h.n_vlr)
    buf[104] = h.point_format
    struct.pack_into("<H", buf, 105, h.record_length)
    legacy_ok = h.point_count < 2**32 and not (h.point_format >= 6)
    struct.pack_into("<I", buf, 107, h.point_count if legacy_ok else 0)
    if legacy_ok:
        struct.pack_into("<I", buf, 111, h.point_count)  # points by return [0]
    struct.pack_into("<3d", buf, 131, *[float(v) for v in h.scales])
    struct.pack_into("<3d", buf, 155, *[float(v) for v in h.offsets])
    struct.pack_into("<6d", buf, 179, float(h.maxs[0]), float(h.mins[0]), float(h.maxs[1]),
                     float(h.mins[1]), float(h.maxs[2]), float(h.mins[2]))
    if (vmaj, vmin) >= (1, 4):
        struct.pack_into("<Q", buf, 247, h.point_count)
        struct.pack_into("<Q", buf, 255, h.point_count)
    return bytes(buf)


def write_raw(path: str, header: LasHeader, records: np.ndarray,
              lattice_min: Optional[np.ndarray] = None, lattice_max: Optional[np.ndarray] = None) -> None:
    """Write header + already-encoded records.  ``lattice_min/max`` are the int32 X/Y/Z extrema the
    encode kernel reduced; the header min/max are ``X*scale+offset`` of those (laspy recomputes the
    header bounds on write the same way)."""
    h = header.copy()
    rec = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1)
    if rec.size % h.record_length:
        raise LasError("record buffer is not a whole number of records")
    h.point_count = rec.size // h.record_length
    if h.point_count and lattice_min is not None:
        h.mins = np.asarray(lattice_min, dtype=np.float64) * h.scales + h.offsets
        h.maxs = np.asarray(lattice_max, dtype=np.float64) * h.scales + h.offsets
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(header_bytes(h))
        pad = h.offset_to_point_data - h.header_size
        if pad > 0:
            f.write(b"\0" * pad)
        f.write(rec.tobytes() if not isinstance(rec, np.memmap) else bytes(rec))
