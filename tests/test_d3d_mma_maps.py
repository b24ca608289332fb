"""Host-side check of the lane-level index maps of csrc/d3d_first_mma.cu (the warp-level tensor-core kernels of d3d.0,
p2igan.py:139-142): tools/emulate_d3d_mma.py recomputes every shared-memory / global address and every mma.sync / ldmatrix fragment
slot with the kernel's formulas and the PTX fragment definition; the result must be torch's conv3d and its autograd.  The CUDA
kernels themselves are compared with torch on the GPU in test_gpu_discriminator.py::test_d3d_first_layer_forward_and_weight_gradient."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))


@pytest.mark.parametrize("B,T,H,W", [(1, 2, 4, 32), (2, 3, 8, 32)])
def test_fragment_maps_reproduce_conv3d(B, T, H, W):
    import emulate_d3d_mma as E
    e_fwd, e_dw, e_db, holes = E.run(B, T, H, W, seed=B + T)
    assert holes == 0                      # every output slot is written exactly by the staged stores
    assert e_fwd < 1e-4 and e_dw < 1e-3 and e_db < 1e-3
