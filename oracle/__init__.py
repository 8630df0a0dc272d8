"""CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy / scipy / scikit-learn restatement of pointcloudhookup's per-point LAS hot path, stage by
stage, used only as the parity checker (tests/, __graft_entry__.smoke()) and as the timed CPU
baseline (bench.py cpu_baseline / --impl reference).  Nothing under pointcloudhookup_b200/ may
import this package: the product path is the CUDA library and fails loudly without it.

Pinning status:
  * numpy percentile and scikit-learn DBSCAN are the reference's own dependencies and are called
    for real (numpy 2.3.5 / scikit-learn 1.9.0 in this image)            -> pinned by the library.
  * the reference's own control flow (chunking, label offsets, set()-order dedup, north angle,
    progress milestones) is pinned by tests/golden/*.npz, produced by importing the UNMODIFIED
    reference modules from /root/reference with shims for its absent third-party imports
    (tests/golden/make_golden.py).
  * EPSG:4547->4326 is pinned by the reference's four known-answer towers
    (test/kuangxuan.py:29-33 <-> elevation_conversion.py:148-153).
  * laspy / open3d / trimesh / pyproj(vgridshift) arithmetic is restated from the libraries'
    public behaviour; the libraries are absent from /root/reference and from this image and the
    reference records no expected values for them                            -> PARITY UNPINNED.
"""
