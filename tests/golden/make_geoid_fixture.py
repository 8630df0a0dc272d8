"""Crop the reference's shipped EGM96 grid (egm96_15.gtx, a public NOAA/PROJ data file) to the
20-35N / 105-120E window around the reference's recorded towers and store it as a small regional
GTX so the GPU box (which has no /root/reference) can test real geoid values.
    python tests/golden/make_geoid_fixture.py
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import geoid  # noqa: E402

g = geoid.read_gtx("/root/reference/egm96_15.gtx")
r0, r1 = int((20 + 90) / 0.25), int((35 + 90) / 0.25)
c0, c1 = int((105 + 180) / 0.25), int((120 + 180) / 0.25)
sub = g["grid"][r0:r1 + 1, c0:c1 + 1]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "egm96_crop_20N35N_105E120E.gtx")
with open(out, "wb") as f:
    f.write(struct.pack(">4d", -90 + r0 * 0.25, -180 + c0 * 0.25, 0.25, 0.25))
    f.write(struct.pack(">2i", sub.shape[0], sub.shape[1]))
    f.write(sub.astype(">f4").tobytes())
print(out, sub.shape, os.path.getsize(out))
# known values on the full grid for the four reference towers (elevation_conversion.py:148-153)
pts = [(28.379751, 113.363246), (28.373584, 113.365316), (28.369979, 113.366579), (28.376940, 113.364167)]
full = [float(geoid.geoid_height(g, la, lo)) for la, lo in pts]
crop = geoid.read_gtx(out)
assert np.allclose(full, [float(geoid.geoid_height(crop, la, lo)) for la, lo in pts], rtol=0, atol=1e-11)
print(full)
