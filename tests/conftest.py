import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def make_las_dict(rec, scales, offsets, point_format=3, version=(1, 2)):
    """Oracle-style decoded dict from a structured PDRF3 record array."""
    return {"version": version, "point_format": point_format, "record_length": rec.dtype.itemsize,
            "scales": np.asarray(scales, float), "offsets": np.asarray(offsets, float),
            "X": np.ascontiguousarray(rec["X"]), "Y": np.ascontiguousarray(rec["Y"]),
            "Z": np.ascontiguousarray(rec["Z"]), "n": int(rec.size), "header_size": 227}
