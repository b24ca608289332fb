"""Minimal zarr-v2 directory store for inference outputs.

scripts/infer.py:171-186,249-260 writes one float32 dataset ``event_XX`` per test event (a single chunk = the whole
array) into a zarr group whose attributes record the run.  zarr itself is not part of this image, so this module writes
the same on-disk layout by hand -- zarr format 2, C order, little-endian float32, one uncompressed chunk per array
(``"compressor": null``) -- which ``zarr.open_group`` / ``xarray.open_zarr`` and the reference's ``experiments/io.py``
read unchanged.  Pure host-side IO; nothing here touches the GPU.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Iterable

import numpy as np


def _write_json(path: str, obj: Any) -> None:
    with open(path, "w", encoding="utf-8") as f:
        json.dump(obj, f, indent=4, sort_keys=True)


def open_group(path: str, attrs: Dict[str, Any], overwrite: bool = False) -> str:
    """Create the group directory (``.zgroup`` + ``.zattrs``).  Like infer.py:160-168, refuses to clobber an existing output."""
    if os.path.exists(path):
        if not overwrite:
            raise FileExistsError(f"Output already exists: {path}")
        import shutil
        shutil.rmtree(path)
    os.makedirs(path)
    _write_json(os.path.join(path, ".zgroup"), {"zarr_format": 2})
    _write_json(os.path.join(path, ".zattrs"), attrs)
    return path


def write_array(group: str, name: str, data: np.ndarray) -> None:
    """``group.create_dataset(name, shape=data.shape, chunks=data.shape, dtype='float32', overwrite=True)[:] = data``."""
    a = np.ascontiguousarray(data, dtype="<f4")
    d = os.path.join(group, name)
    os.makedirs(d, exist_ok=True)
    _write_json(os.path.join(d, ".zarray"), {
        "chunks": list(a.shape), "compressor": None, "dtype": "<f4", "fill_value": 0.0, "filters": None, "order": "C",
        "shape": list(a.shape), "zarr_format": 2})
    chunk = ".".join("0" for _ in a.shape) or "0"
    with open(os.path.join(d, chunk), "wb") as f:
        f.write(a.tobytes(order="C"))


def read_array(group: str, name: str) -> np.ndarray:
    """Inverse of ``write_array`` for single-chunk, uncompressed arrays (used by the tests and for multi-pass averaging)."""
    d = os.path.join(group, name)
    meta = json.load(open(os.path.join(d, ".zarray"), encoding="utf-8"))
    if meta["compressor"] is not None or meta["filters"] is not None or meta["chunks"] != meta["shape"]:
        raise ValueError(f"{d}: only single-chunk uncompressed arrays are supported")
    chunk = ".".join("0" for _ in meta["shape"]) or "0"
    raw = np.fromfile(os.path.join(d, chunk), dtype=np.dtype(meta["dtype"]))
    return raw.reshape(meta["shape"], order=meta["order"])


def list_arrays(group: str) -> Iterable[str]:
    return sorted(n for n in os.listdir(group) if os.path.exists(os.path.join(group, n, ".zarray")))
