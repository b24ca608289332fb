"""ctypes binding of libp2i_sm100a.so, generated from ``include/p2i_b200.h``.

The prototypes are parsed from the header at import time, so the Python side cannot drift from
the C ABI.  There is no fallback: if the library is missing, every op raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
import re
import struct
from typing import Dict, List, Tuple

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(os.path.dirname(PKG_DIR))
HEADER = os.path.join(REPO_ROOT, "include", "p2i_b200.h")
# P2I_LIB_PATH: an alternative build of the same library (e.g. the -DHALO_PROF diagnostics build, tools/halo_prof.sh)
LIB_PATH = os.environ.get("P2I_LIB_PATH") or os.path.join(PKG_DIR, "libp2i_sm100a.so")

_CTYPE = {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong, "size_t": ctypes.c_size_t}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """{name: (return type, [(ctype string, arg name), ...])} for every `p2i_*` prototype."""
    text = open(path, "r", encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?m)^\s*((?:const\s+)?[\w ]+?\**)\s*(p2i_\w+)\s*\(([^;{]*?)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"^(.*?)(\w+)$", a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


def _to_ctype(t: str):
    if "*" in t:
        return ctypes.c_void_p
    t = t.replace("const", "").strip()
    return _CTYPE[t]


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU or PyTorch fallback.")
        dll = ctypes.CDLL(LIB_PATH)
        for name, (ret, args) in self.protos.items():
            fn = getattr(dll, name)           # AttributeError here == header/library drift
            fn.argtypes = [_to_ctype(t) for t, _ in args]
            fn.restype = ctypes.c_char_p if "char" in ret else (ctypes.c_longlong if "long long" in ret else ctypes.c_int)
        self._dll = dll
        return dll

    def call(self, name: str, *args):
        dll = self.load()
        rc = getattr(dll, name)(*args)
        if rc != 0:
            msg = dll.p2i_last_error()
            raise RuntimeError(f"{name} failed ({rc}): {msg.decode() if msg else '?'}")

    def launch_count(self) -> int:
        return int(self.load().p2i_launch_count())


LIB = _Lib()


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("p2igan_b200 runs on CUDA (sm_100a) tensors only; got a tensor on %s. "
                               "There is no CPU path." % t.device)


def pack_do_grad_table(entries) -> bytes:
    """Pack P2iDoGrad structs: 6 pointers + 2 ints (56 bytes each)."""
    out = b""
    for W, D, Dd, g, dW, dD, ch in entries:
        out += struct.pack("<QQQQQQii", W, D, Dd, g, dW, dD, ch, 0)
    return out


def pack_do_table(entries) -> bytes:
    """Pack P2iDoLayer structs: 5 pointers + 2 ints (48 bytes each)."""
    out = b""
    for W, D, Dd, o, ot, ch in entries:
        out += struct.pack("<QQQQQii", W, D, Dd, o, ot, ch, 0)
    return out
