"""Generate tests/golden/reference_run.json by running the UNMODIFIED reference modules.

    python tests/golden/make_golden.py            (needs /root/reference; run in the build container)

The reference (pure Python) cannot be imported as is: laspy, open3d, trimesh and pyproj are not
installed and cannot be installed offline.  This script registers thin stand-ins for exactly those
four third-party imports — backed by the oracle's restatements of the libraries' arithmetic — and
then imports and runs the reference's OWN files from /root/reference:

    ui/import_PC.py::run_voxel_downsampling, process_chunk
    ui/Sampling.py::voxel_downsample_open3d
    utils/tower_extraction.py::extract_towers          (real numpy percentile, real sklearn DBSCAN)
    utils/elevation_converter.py::ElevationConverter   (fallback path and grid path)

So the reference's control flow — chunk loops, label offsets, set()-order iteration, size filter,
duplicate check, north angle, callback milestones, fallbacks — is pinned by the real code; the
third-party numerics behind the shims stay "parity unpinned" (see oracle/__init__.py).
Outputs are stored as sha256 digests of the exact arrays plus the small float results.
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import geoid as o_geoid, las_io, obb as o_obb, voxel as o_voxel  # noqa: E402
from pointcloudhookup_b200 import synth  # noqa: E402  (input generator only)


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------ shims
def install_shims(geoid_grid=None):
    laspy = types.ModuleType("laspy")

    class _Header:
        def __init__(self, point_format=3, version=(1, 2)):
            self.point_format = point_format
            self.version = version
            self.offsets = np.zeros(3)
            self.scales = np.array([0.01, 0.01, 0.01])

    class _Points:
        def __init__(self, las, lo=0, hi=None):
            self._las, self._lo, self._hi = las, lo, hi

        def __len__(self):
            n = self._las["n"]
            return len(range(*slice(self._lo, self._hi).indices(n)))

        def __getitem__(self, s):
            assert isinstance(s, slice) and s.step in (None, 1)
            return _Points(self._las, s.start or 0, s.stop)

        x = property(lambda self: las_io.scaled(self._las, self._lo, self._hi)[0])
        y = property(lambda self: las_io.scaled(self._las, self._lo, self._hi)[1])
        z = property(lambda self: las_io.scaled(self._las, self._lo, self._hi)[2])

    class LasData:
        def __init__(self, header, las=None):
            self.header = header
            self._las = las
            self._xyz = {}
            self.points = _Points(las) if las is not None else None

        def _get(self, i):
            return las_io.scaled(self._las)[i]

        x = property(lambda self: self._get(0), lambda self, v: self._xyz.__setitem__("x", np.asarray(v)))
        y = property(lambda self: self._get(1), lambda self, v: self._xyz.__setitem__("y", np.asarray(v)))
        z = property(lambda self: self._get(2), lambda self, v: self._xyz.__setitem__("z", np.asarray(v)))

        def write(self, path):
            like = {"scales": np.asarray(self.header.scales, float), "offsets": np.asarray(self.header.offsets, float),
                    "point_format": self.header.point_format, "version": tuple(self.header.version)}
            las_io.write_las(str(path), like, self._xyz["x"], self._xyz["y"], self._xyz["z"])

    def read(path):
        las = las_io.read_las(str(path))
        h = _Header(las["point_format"], las["version"])
        h.offsets, h.scales = las["offsets"], las["scales"]
        return LasData(h, las)

    class _Open:
        def __init__(self, path):
            self.path = path

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def read(self):
            return read(self.path)

    laspy.read, laspy.open, laspy.LasHeader, laspy.LasData = read, _Open, _Header, LasData

    o3d = types.ModuleType("open3d")
    o3d.geometry = types.ModuleType("open3d.geometry")
    o3d.utility = types.ModuleType("open3d.utility")
    o3d.utility.Vector3dVector = lambda a: np.asarray(a, dtype=np.float64)

    class PointCloud:
        def __init__(self):
            self.points = np.zeros((0, 3))

        def voxel_down_sample(self, v):
            out = PointCloud()
            out.points = o_voxel.voxel_down_sample(self.points, v)   # canonical order
            return out
    o3d.geometry.PointCloud = PointCloud

    trimesh = types.ModuleType("trimesh")

    class _Box:
        def __init__(self, transform, extents):
            self.transform, self.extents = transform, extents

    class TPointCloud:
        def __init__(self, pts):
            self.vertices = np.asarray(pts)

        @property
        def bounding_box_oriented(self):
            # trimesh's hull-face search without its 0.1 rad thinning in Qhull's facet order (which no other hull
            # construction can reproduce): every face normal is a candidate.  oracle/obb.py explains.
            t, ext, _ = o_obb.min_volume_box_all_faces(self.vertices)
            return _Box(t, ext)
    trimesh.PointCloud = TPointCloud

    pyproj = types.ModuleType("pyproj")
    pyproj.datadir = types.ModuleType("pyproj.datadir")
    pyproj.datadir.get_data_dir = lambda: "/nonexistent/proj"

    class _Pipeline:
        def __init__(self, mult):
            self.mult = mult

        def transform(self, lon, lat, h):
            return lon, lat, float(o_geoid.vgridshift(geoid_grid, lon, lat, h, self.mult))

    class Transformer:
        @staticmethod
        def from_pipeline(s):
            if geoid_grid is None:
                raise RuntimeError("grid egm08_25.gtx not found")   # what PROJ does without the file
            mult = float(s.split("+multiplier=")[1].split()[0])
            return _Pipeline(mult)
    pyproj.Transformer = Transformer

    for name, mod in (("laspy", laspy), ("open3d", o3d), ("open3d.geometry", o3d.geometry),
                      ("open3d.utility", o3d.utility), ("trimesh", trimesh), ("pyproj", pyproj),
                      ("pyproj.datadir", pyproj.datadir)):
        sys.modules[name] = mod


def load_ref(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------ cases
CASES = {
    # name: (n_points, n_towers, terrain, seed, fractions, voxel, chunk)
    "dense_wires": (120000, 2, "flat", 21, (0.80, 0.08, 0.07, 0.05), 0.1, 50000),
    "towers": (400000, 2, "flat", 22, (0.86, 0.085, 0.005, 0.05), 0.1, 150000),
    "hilly": (150000, 2, "hilly", 23, (0.86, 0.085, 0.005, 0.05), 0.25, 1000000),
}


def run_case(name, cfg, workdir):
    n, towers, terrain, seed, fr, voxel, chunk = cfg
    src = os.path.join(workdir, f"{name}.las")
    synth.write_corridor_las(src, n, towers, terrain, seed, fr)
    out = {"config": {"n": n, "towers": towers, "terrain": terrain, "seed": seed, "fractions": list(fr),
                      "voxel": voxel, "chunk": chunk}}
    with open(src, "rb") as f:
        out["input_sha256"] = hashlib.sha256(f.read()).hexdigest()

    imp = load_ref("ui/import_PC.py", "ref_import_PC")
    logs, prog = [], []
    dst = os.path.join(workdir, "output", f"{name}_ds.las")
    imp.run_voxel_downsampling(src, dst, voxel_size=voxel, chunk_size=chunk, progress_callback=prog.append,
                               log_callback=logs.append)
    ds = las_io.read_las(dst)
    out["downsample"] = {"count": ds["n"], "xyz_sha256": digest(np.stack([ds["X"], ds["Y"], ds["Z"]], 1)),
                         "progress": prog, "n_logs": len(logs), "first_rows": np.stack([ds["X"], ds["Y"], ds["Z"]], 1)[:4].tolist()}
    try:
        imp.run_voxel_downsampling(os.path.join(workdir, "missing.las"), dst)
        out["downsample"]["missing_raises"] = None
    except Exception as e:
        out["downsample"]["missing_raises"] = type(e).__name__

    smp = load_ref("ui/Sampling.py", "ref_Sampling")
    dst2 = os.path.join(workdir, "output", f"{name}_ds2.las")
    smp.voxel_downsample_open3d(src, dst2, voxel, chunk)
    ds2 = las_io.read_las(dst2)
    out["sampling_same_as_import_pc"] = bool(np.array_equal(ds2["X"], ds["X"]) and np.array_equal(ds2["Z"], ds["Z"]))
    smp.voxel_downsample_open3d(os.path.join(workdir, "missing.las"), dst2, voxel, chunk)  # must not raise

    pts = np.stack(las_io.scaled(las_io.read_las(src), 0, 5000), 1)
    pc = imp.process_chunk(pts, voxel)
    out["process_chunk"] = {"count": int(pc.shape[0]), "sha256": digest(pc)}

    tex = load_ref("utils/tower_extraction.py", "ref_tower_extraction")
    rec_labels = []
    real = tex.DBSCAN

    class RecordingDBSCAN(real):
        def fit(self, X, *a, **k):
            r = super().fit(X, *a, **k)
            rec_labels.append(r.labels_.copy())
            return r
    tex.DBSCAN = RecordingDBSCAN
    logs, prog = [], []
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        res = tex.extract_towers(dst, progress_callback=prog.append, log_callback=logs.append)
    finally:
        os.chdir(cwd)
    # rebuild all_labels the way the reference does (per-chunk offset)
    cur, parts = 0, []
    for lab in rec_labels:
        lab = lab.copy()
        lab[lab != -1] += cur
        cur = lab.max() + 1 if (lab != -1).any() else cur
        parts.append(lab)
    all_labels = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    out["towers"] = {
        "n_filtered": int(all_labels.size), "labels_sha256": digest(all_labels), "n_clusters": int(cur),
        "progress": prog, "count": len(res),
        "list": [{"center": t["center"].tolist(), "extent": np.asarray(t["extent"]).tolist(),
                  "rotation": np.asarray(t["rotation"]).tolist(), "height": float(t["height"]),
                  "width": float(t["width"]), "north_angle": float(t["north_angle"]),
                  "n_points": int(len(t["points"])), "points_sha256": digest(t["points"])} for t in res],
        "tower_files": sorted(f for f in os.listdir(os.path.join(workdir, "output_towers"))) if os.path.isdir(os.path.join(workdir, "output_towers")) else [],
    }
    for f in out["towers"]["tower_files"]:
        os.remove(os.path.join(workdir, "output_towers", f))
    return out


def run_elevation():
    out = {}
    towers = [(28.379751, 113.363246, 131.46), (28.373584, 113.365316, 87.77),
              (28.369979, 113.366579, 80.06), (28.376940, 113.364167, 82.56)]  # elevation_conversion.py:148-153
    install_shims(None)
    ec = load_ref("utils/elevation_converter.py", "ref_elev_nogrid")
    conv = ec.ElevationConverter()
    out["fallback"] = [conv.ellipsoid_to_orthometric(*t) for t in towers]
    out["fallback_batch"] = conv.convert_batch(*zip(*towers)).tolist()
    out["convert_elevation"] = ec.convert_elevation(*towers[0], region_n_value=20.0)
    grid = o_geoid.read_gtx(os.path.join(REF, "egm96_15.gtx"))
    install_shims(grid)
    ec2 = load_ref("utils/elevation_converter.py", "ref_elev_grid")
    conv2 = ec2.ElevationConverter()
    out["grid_egm96_plus1"] = conv2.convert_batch(*zip(*towers)).tolist()
    out["towers"] = [list(t) for t in towers]
    return out


def run_match():
    """utils/table_match_gim.py::match_towers (unmodified) — PyQt5 is GUI-only there and is stubbed; pyproj's
    Transformer.from_crs is served by the oracle's inverse Gauss-Krueger."""
    import importlib
    from oracle import crs as o_crs
    for name in ("PyQt5", "PyQt5.QtWidgets", "PyQt5.QtCore", "PyQt5.QtGui"):
        m = types.ModuleType(name)
        m.__getattr__ = lambda attr: type(attr, (), {})      # any Qt class name resolves to a dummy class
        sys.modules[name] = m
    install_shims(None)

    class _T:
        def transform(self, x, y):
            lon, lat = o_crs.gk_inverse(x, y)
            return (float(lon), float(lat)) if np.ndim(x) == 0 else (lon, lat)
    sys.modules["pyproj"].Transformer.from_crs = staticmethod(lambda *a, **k: _T())
    # the module does `from utils.elevation_converter import ElevationConverter`
    utils_pkg = types.ModuleType("utils")
    utils_pkg.__path__ = [os.path.join(REF, "utils")]
    sys.modules["utils"] = utils_pkg
    tm = load_ref("utils/table_match_gim.py", "ref_table_match_gim")
    # point-cloud towers: the reference's own recorded run (test/kuangxuan.py:29-33)
    pc = [{"center": np.array([437587.898, 3140691.58, 131.45735]), "height": 17.4, "north_angle": 12.0},
          {"center": np.array([437787.178, 3140006.96, 87.7722064]), "height": 29.8, "north_angle": 15.5},
          {"center": np.array([437908.948, 3139606.82, 80.0563301]), "height": 21.8, "north_angle": 16.1},
          {"center": np.array([437676.583, 3140379.50, 82.5588932]), "height": 21.0, "north_angle": 14.9}]
    # GIM towers: lat/lng/h in the layout of test/data1.py, placed 0-70 m away from the cloud towers
    gim = [{"lat": 28.379751, "lng": 113.363246, "h": 106.0, "r": 10.0, "properties": {"杆塔编号": "P142"}},
           {"lat": 28.376940 + 3.0e-4, "lng": 113.364167, "h": 57.9, "r": 11.0, "properties": {"杆塔编号": "P143"}},
           {"lat": 28.373584, "lng": 113.365316 + 7.5e-4, "h": 62.0, "r": 12.0, "properties": {"杆塔编号": "P144"}},
           {"lat": 28.369979, "lng": 113.366579, "h": 300.0, "r": 13.0, "properties": {"杆塔编号": "P145"}},
           {"lat": 28.40, "lng": 113.40, "h": 50.0, "r": 0.0, "properties": {"杆塔编号": "far"}}]
    matched, conv = tm.match_towers(gim, pc, tm.Transformer.from_crs("EPSG:4547", "EPSG:4326", always_xy=True))
    hav = [[tm.haversine(g["lat"], g["lng"], c["converted_center"][1], c["converted_center"][0]) for c in conv] for g in gim]
    return {"pc": [{"center": t["center"].tolist(), "height": t["height"], "north_angle": t["north_angle"]} for t in pc],
            "gim": gim, "matched": [list(m) for m in matched],
            "converted": [{k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in c.items()} for c in conv],
            "haversine": hav}


def main():
    install_shims(None)
    result = {"generator": "tests/golden/make_golden.py", "numpy": np.__version__}
    import sklearn
    result["sklearn"] = sklearn.__version__
    with tempfile.TemporaryDirectory() as wd:
        for name, cfg in CASES.items():
            print("case", name, flush=True)
            result[name] = run_case(name, cfg, wd)
    result["elevation"] = run_elevation()
    result["match"] = run_match()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_run.json"), "w") as f:
        json.dump(result, f, indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else list(v.keys())) for k, v in result.items()}, indent=1))


if __name__ == "__main__":
    main()
