"""GPU: the branch-free float32 square root and reciprocal inside the EM loop (common.cuh: sqrt_rn_normal,
div_rn_normal) are bit-identical to the IEEE routines (`__fsqrt_rn`, `__fdiv_rn`) they replace, on the range the
mixture variances (>= reg_covar = 1e-6) and standard deviations live in -- sklearn computes 1 / sqrt(var) with two
correctly rounded float32 operations (_gaussian_mixture.py:323-385), so must the kernel."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(x):
    from paa_b200 import _lib
    lib = _lib.load()
    xd = torch.from_numpy(x).cuda()
    out = torch.empty((4, x.shape[0]), dtype=torch.float32, device="cuda")
    _lib.check(lib.paa_selftest_roots(xd.data_ptr(), x.shape[0], out.data_ptr(),
                                      _lib.stream_handle(xd.device)), "paa_selftest_roots")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_fast_roots_are_bit_identical_to_the_ieee_routines():
    rng = np.random.default_rng(7)
    n = 1 << 22
    logs = rng.uniform(np.log(1e-8), np.log(1e16), size=n)
    x = np.exp(logs).astype(np.float32)
    # plus every float32 in a few dense neighbourhoods (powers of two, the 1e-6 floor, 1.0)
    dense = []
    for c in (1e-6, 1.0, 2.0, 4.0, 0.25, 1e-3, 7.3):
        b = np.float32(c).view(np.uint32)
        dense.append((np.arange(-50000, 50000, dtype=np.int64) + int(b)).astype(np.uint32).view(np.float32))
    x = np.concatenate([x] + dense)
    out = _run(x)
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32)), "sqrt_rn_normal != __fsqrt_rn"
    assert np.array_equal(out[2].view(np.uint32), out[3].view(np.uint32)), "div_rn_normal(1, x) != __fdiv_rn(1, x)"
    # and both agree with numpy's correctly rounded float32 results
    assert np.array_equal(out[0], np.sqrt(x))
    assert np.array_equal(out[2], (np.float32(1.0) / x).astype(np.float32))
