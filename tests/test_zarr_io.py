"""CPU: the hand-written zarr-v2 store (p2igan_b200/zarr_io.py) against the zarr v2 storage spec fields that
zarr.open_group / the reference's experiments read (scripts/infer.py:171-186,249-260)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "p2i-gan-benchmark_b200"))
from p2igan_b200 import zarr_io  # noqa: E402


def test_group_layout_and_roundtrip(tmp_path):
    g = zarr_io.open_group(str(tmp_path / "out.zarr"), {"model_name": "p2igan", "passes": 1})
    a = np.random.RandomState(0).rand(20, 1, 8, 12).astype(np.float32)
    zarr_io.write_array(g, "event_01", a)
    zarr_io.write_array(g, "event_02", a[:5] * 2)
    assert json.load(open(os.path.join(g, ".zgroup"))) == {"zarr_format": 2}
    assert json.load(open(os.path.join(g, ".zattrs")))["model_name"] == "p2igan"
    meta = json.load(open(os.path.join(g, "event_01", ".zarray")))
    assert meta == {"chunks": [20, 1, 8, 12], "compressor": None, "dtype": "<f4", "fill_value": 0.0, "filters": None,
                    "order": "C", "shape": [20, 1, 8, 12], "zarr_format": 2}
    # chunk key of the single chunk of a 4-D array in zarr v2 with the default '.' separator
    assert os.path.getsize(os.path.join(g, "event_01", "0.0.0.0")) == a.nbytes
    assert np.array_equal(zarr_io.read_array(g, "event_01"), a)
    assert list(zarr_io.list_arrays(g)) == ["event_01", "event_02"]
    zarr_io.write_array(g, "event_01", a + 1)                      # overwrite=True semantics
    assert np.array_equal(zarr_io.read_array(g, "event_01"), a + 1)
    with pytest.raises(FileExistsError):
        zarr_io.open_group(g, {})
    zarr_io.open_group(g, {}, overwrite=True)
    assert list(zarr_io.list_arrays(g)) == []
