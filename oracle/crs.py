"""Oracle EPSG:4547 -> EPSG:4326: ``Transformer.from_crs("EPSG:4547","EPSG:4326",always_xy=True)``.

Call sites: utils/table_match_gim.py:72-75,232,346 (per tower); test/005test.py:37,55 (per point).
pyproj is absent; restated as PROJ's extended transverse Mercator inverse (Krueger series to n^6,
SURVEY.md Appendix A.7): CGCS2000 ellipsoid a=6378137, 1/f=298.257222101, lon0=114E, k0=1,
FE=500000, FN=0; CGCS2000->WGS84 is a null datum step.  PINNED by the reference's four
known-answer towers (test/kuangxuan.py:29-33 <-> elevation_conversion.py:148-153, 6 decimals).
"""
import numpy as np

A = 6378137.0
INV_F = 298.257222101
LON0 = 114.0
K0 = 1.0
FE = 500000.0
FN = 0.0


def _consts():
    f = 1.0 / INV_F
    n = f / (2.0 - f)
    n2, n3, n4, n5, n6 = n * n, n ** 3, n ** 4, n ** 5, n ** 6
    Ar = A / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0)
    beta = [
        n / 2.0 - 2.0 * n2 / 3.0 + 37.0 * n3 / 96.0 - n4 / 360.0 - 81.0 * n5 / 512.0 + 96199.0 * n6 / 604800.0,
        n2 / 48.0 + n3 / 15.0 - 437.0 * n4 / 1440.0 + 46.0 * n5 / 105.0 - 1118711.0 * n6 / 3870720.0,
        17.0 * n3 / 480.0 - 37.0 * n4 / 840.0 - 209.0 * n5 / 4480.0 + 5569.0 * n6 / 90720.0,
        4397.0 * n4 / 161280.0 - 11.0 * n5 / 504.0 - 830251.0 * n6 / 7257600.0,
        4583.0 * n5 / 161280.0 - 108847.0 * n6 / 3991680.0,
        20648693.0 * n6 / 638668800.0,
    ]
    e2 = f * (2.0 - f)
    return Ar, beta, np.sqrt(e2)


def gk_inverse(x, y):
    """(easting, northing) -> (lon_deg, lat_deg), float64."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    Ar, beta, e = _consts()
    xi = (y - FN) / (K0 * Ar)
    eta = (x - FE) / (K0 * Ar)
    xip, etap = xi.copy(), eta.copy()
    for j, b in enumerate(beta, start=1):
        xip = xip - b * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
        etap = etap - b * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
    sinh_e = np.sinh(etap)
    cos_x = np.cos(xip)
    taup = np.sin(xip) / np.sqrt(sinh_e * sinh_e + cos_x * cos_x)
    lam = np.arctan2(sinh_e, cos_x)
    tau = taup.copy()
    for _ in range(5):
        sigma = np.sinh(e * np.arctanh(e * tau / np.sqrt(1.0 + tau * tau)))
        taui = tau * np.sqrt(1.0 + sigma * sigma) - sigma * np.sqrt(1.0 + tau * tau)
        dtau = (taup - taui) / np.sqrt(1.0 + taui * taui) * (1.0 + (1.0 - e * e) * tau * tau) / ((1.0 - e * e) * np.sqrt(1.0 + tau * tau))
        tau = tau + dtau
    lat = np.arctan(tau)
    return LON0 + np.degrees(lam), np.degrees(lat)
