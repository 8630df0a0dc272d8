"""Generate tests/golden/variants_run.json by EXECUTING THE REFERENCE'S OWN SOURCE LINES for the two scratch-script
stages that widen the hot path:

    python tests/golden/make_golden_variants.py            (needs /root/reference; run in the build container)

  test/kuangxuan.py:28-45   the tower list parsed from the reference's logged detections
  test/kuangxuan.py:60-79   per-tower box + `tower_points = points[mask]`
  test/tttt.py:93-175       cluster post-processing: merge adjacent clusters (centres, KDTree radius query, union-find,
                            relabel)
  test/main_ground.py:8-32, 77-115   remove_ground_ransac / remove_ground_tiled_ransac, with the REAL scikit-learn
                            RANSACRegressor; the only intervention is random_state = 1000 + call number, so that the draws
                            can be replayed (the reference passes none and is not reproducible run to run)

Neither file can be imported (kuangxuan.py opens a Windows path and an open3d window at import; tttt.py lacks its
own imports), so the line ranges are read from /root/reference at generation time, dedented and exec'd in a
namespace that supplies only what the surrounding script would have (numpy, re, sklearn's KDTree, the inputs).  No
reference source is copied into this repository: only seeded inputs' parameters and output digests are stored.
"""
import json
import os
import re
import sys
import textwrap

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from variant_inputs import crop_inputs, digest, merge_inputs, terrain_cloud  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "variants_run.json")


def ref_lines(rel, lo, hi):
    with open(os.path.join(REF, rel), encoding="utf-8") as f:
        lines = f.read().split("\n")
    return textwrap.dedent("\n".join(lines[lo - 1: hi]))


def main():
    out = {"source": {"kuangxuan_parse": "test/kuangxuan.py:28-45", "kuangxuan_crop": "test/kuangxuan.py:60-79",
                      "tttt_merge": "test/tttt.py:93-175"}}
    # ---- tower list exactly as the reference parses it from its own log
    ns = {"re": re}
    exec(ref_lines("test/kuangxuan.py", 28, 45), ns)
    tower_data = ns["tower_data"]
    out["tower_data"] = tower_data
    # ---- crop: the loop body runs once per tower; collect tower_points after each iteration
    body = ref_lines("test/kuangxuan.py", 61, 79)
    crops = []
    for seed in (101, 102):
        points = crop_inputs(tower_data, seed)
        per_tower = []
        for tower in tower_data:
            ns = {"tower": tower, "points": points}
            exec(body, ns)
            per_tower.append({"count": int(len(ns["tower_points"])), "sha256": digest(ns["tower_points"]),
                              "bounds": [ns["x_min"], ns["y_min"], ns["z_min"], ns["x_max"], ns["y_max"], ns["z_max"]]})
        crops.append({"seed": seed, "n": int(len(points)), "points_sha256": digest(points), "towers": per_tower})
    out["crop"] = crops
    # ---- merge of adjacent clusters
    from sklearn.neighbors import KDTree
    merges = []
    block = ref_lines("test/tttt.py", 93, 175)
    for seed, thr in ((201, 6.0), (202, 6.0), (203, 12.0), (204, 0.5)):
        filtered_points, all_labels = merge_inputs(seed)
        logs = []
        ns = {"np": np, "KDTree": KDTree, "all_labels": all_labels, "filtered_points": filtered_points,
              "merge_threshold": thr, "log": logs.append, "progress": lambda v: None}
        exec(block, ns)
        merged = np.asarray(ns["merged_labels"])
        merges.append({"seed": seed, "merge_threshold": thr, "n": int(len(all_labels)),
                       "labels_sha256": digest(all_labels), "points_sha256": digest(filtered_points),
                       "merged_sha256": digest(merged.astype(np.int64)),
                       "n_merged_clusters": int(len(set(merged.tolist()) - {-1})),
                       "iteration_order": [int(v) for v in ns["unique_labels"]]})
    out["merge"] = merges
    # ---- tiled RANSAC ground: the reference's two functions, scikit-learn's estimator seeded per call
    from sklearn.linear_model import RANSACRegressor as RealRansac
    out["source"]["ransac"] = "test/main_ground.py:8-32,77-115"
    code = ref_lines("test/main_ground.py", 8, 32) + "\n\n" + ref_lines("test/main_ground.py", 77, 115)
    runs = []
    for seed, kw, tile, iters in ((5, dict(nx_m=38.0, ny_m=27.0), 10.0, 120),
                                  (6, dict(nx_m=31.0, ny_m=29.0, origin=(-14.2, -9.9)), 7.5, 80)):
        points = terrain_cloud(seed, **kw)
        sizes = []

        class Recording(RealRansac):
            def fit(self, X, y, **k):
                sizes.append(int(len(X)))
                return super().fit(X, y, **k)

        def seeded(**k):
            return Recording(random_state=1000 + len(sizes), **k)

        ns = {"np": np, "RANSACRegressor": seeded}
        exec(code, ns)
        non_ground, ground = ns["remove_ground_tiled_ransac"](points, tile_size=tile, distance_threshold=0.1, max_iterations=iters)
        runs.append({"seed": seed, "cloud": kw, "tile_size": tile, "distance_threshold": 0.1, "max_iterations": iters,
                     "n": int(len(points)), "points_sha256": digest(points), "fit_sizes": sizes,
                     "non_ground": {"rows": int(len(non_ground)), "sha256": digest(non_ground)},
                     "ground": {"rows": int(len(ground)), "sha256": digest(ground)}})
    out["ransac"] = runs
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
