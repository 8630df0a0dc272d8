"""Oracle LAS decode/encode: what ``laspy.read(p).x`` and ``las.x = arr; las.write(p)`` mean.

Follows ui/import_PC.py:28,35-40,47-48,61-65 and utils/tower_extraction.py:60-62,243-257.
laspy itself is absent (PARITY UNPINNED); semantics per LAS 1.2/1.4 spec: x = X*scale + offset in
float64 (multiply then add), X = round-half-even((x - offset)/scale) -> int32.
Independent of pointcloudhookup_b200.las on purpose (numpy structured views, no shared code).
"""
import os
import struct
import numpy as np

_NOMINAL = {0: 20, 1: 28, 2: 26, 3: 34, 6: 30, 7: 36, 8: 38}


def read_las(path):
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:4] == b"LASF"
    ver = (raw[24], raw[25])
    hsize = struct.unpack_from("<H", raw, 94)[0]
    off = struct.unpack_from("<I", raw, 96)[0]
    pfmt = raw[104] & 0x3F
    rlen = struct.unpack_from("<H", raw, 105)[0]
    n = struct.unpack_from("<I", raw, 107)[0]
    if ver >= (1, 4):
        n64 = struct.unpack_from("<Q", raw, 247)[0]
        n = n64 or n
    scales = np.array(struct.unpack_from("<3d", raw, 131))
    offsets = np.array(struct.unpack_from("<3d", raw, 155))
    dt = np.dtype({"names": ["X", "Y", "Z"], "formats": ["<i4", "<i4", "<i4"], "offsets": [0, 4, 8],
                   "itemsize": rlen})
    pts = np.frombuffer(raw, dtype=dt, count=n, offset=off)
    X, Y, Z = (np.ascontiguousarray(pts[k]) for k in "XYZ")
    return {"version": ver, "point_format": pfmt, "record_length": rlen, "scales": scales,
            "offsets": offsets, "X": X, "Y": Y, "Z": Z, "n": int(n), "header_size": hsize}


def scaled(las, lo=0, hi=None):
    """(x, y, z) float64 of records [lo:hi): one multiply then one add, like laspy's scaled view."""
    s, o = las["scales"], las["offsets"]
    return tuple(las[k][lo:hi].astype(np.float64) * s[i] + o[i] for i, k in enumerate("XYZ"))


def quantise(arr, scale, offset):
    """``las.x = arr``: round-half-even of (arr - offset)/scale to int32."""
    q = np.round((np.asarray(arr, dtype=np.float64) - offset) / scale)
    if q.size and (q.min() < -2**31 or q.max() > 2**31 - 1):
        raise OverflowError("coordinate does not fit the int32 LAS lattice")
    return q.astype(np.int32)


def write_las(path, like, x, y, z):
    """New file with like's point_format/version/scales/offsets, X/Y/Z set, all other dims zero."""
    s, o = like["scales"], like["offsets"]
    X, Y, Z = quantise(x, s[0], o[0]), quantise(y, s[1], o[1]), quantise(z, s[2], o[2])
    pfmt, ver = like["point_format"], like["version"]
    rlen = _NOMINAL[pfmt]
    hsize = 375 if ver >= (1, 4) else (235 if ver == (1, 3) else 227)
    n = X.size
    hdr = bytearray(hsize)
    hdr[0:4] = b"LASF"
    hdr[24], hdr[25] = ver
    struct.pack_into("<H", hdr, 94, hsize)
    struct.pack_into("<II", hdr, 96, hsize, 0)
    hdr[104] = pfmt
    struct.pack_into("<H", hdr, 105, rlen)
    if pfmt < 6 and n < 2**32:
        struct.pack_into("<I", hdr, 107, n)
    struct.pack_into("<3d", hdr, 131, *s)
    struct.pack_into("<3d", hdr, 155, *o)
    if n:
        mm = []
        for i, A in enumerate((X, Y, Z)):
            mm += [A.max() * s[i] + o[i], A.min() * s[i] + o[i]]
        struct.pack_into("<6d", hdr, 179, *mm)
    if ver >= (1, 4):
        struct.pack_into("<Q", hdr, 247, n)
    rec = np.zeros(n, dtype=np.dtype({"names": ["X", "Y", "Z"], "formats": ["<i4"] * 3,
                                      "offsets": [0, 4, 8], "itemsize": rlen}))
    rec["X"], rec["Y"], rec["Z"] = X, Y, Z
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(rec.tobytes())
