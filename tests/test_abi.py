"""CPU: the C-ABI library loads and exports every symbol that include/p2i_b200.h declares (no compute calls),
and the product path fails loudly without CUDA."""
import ctypes
import os

import pytest
import torch


def test_library_exports_every_declared_symbol():
    from p2igan_b200._lib import LIB, LIB_PATH, parse_header
    protos = parse_header()
    assert len(protos) >= 40
    assert os.path.exists(LIB_PATH), "build with `python __graft_entry__.py build`"
    dll = ctypes.CDLL(LIB_PATH)
    missing = [n for n in protos if not hasattr(dll, n)]
    assert not missing, missing
    LIB.load()
    assert LIB.load().p2i_abi_version() == 1
    assert LIB.launch_count() >= 0


def test_no_cpu_fallback():
    import synth
    from p2igan_b200 import build_discriminator, build_generator
    cfg = synth.make_cfg(32, 32)
    frames, masked, masks = synth.make_batch(1, 16, 32, 32, 12, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            build_generator(cfg)(masked, masks)
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            build_discriminator(cfg)(frames)
    from p2igan_b200.losses import ReconstructionLoss
    with pytest.raises(RuntimeError, match="CUDA"):
        ReconstructionLoss(0.05)(frames, frames)


def test_argument_errors_do_not_need_a_gpu():
    from p2igan_b200._lib import LIB
    dll = LIB.load()
    rc = dll.p2i_conv2d_igemm_fwd(None, None, None, None, None, None, 1, 16, 16, 64, 64, 3, 0, None)
    assert rc == -1 and b"null pointer" in dll.p2i_last_error()
    rc = dll.p2i_pyramid_fwd(ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), 1, 12, 12, None)
    assert rc == -1 and b"multiples of 8" in dll.p2i_last_error()
