"""GPU numerics self-test: exact-decode shortcuts == IEEE division for EVERY unorm value, and the kernels'
deterministic log2 == the oracle's, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle as O

pytestmark = pytest.mark.gpu


def test_exact_unorm_decode_and_log2(ctx):
    rng = np.random.default_rng(11)
    x = np.concatenate([np.float32(2.0) ** np.arange(-60, 60, dtype=np.float32), rng.uniform(1e-8, 1e8, 20000).astype(np.float32),
                        np.exp(rng.uniform(-40, 40, 20000)).astype(np.float32), rng.uniform(-2.0, 2.0, 400000).astype(np.float32),
                        rng.uniform(0.5, 1.5, 400000).astype(np.float32), (rng.integers(1, 1 << 20, 200000) / 1024.0).astype(np.float32), np.array([0.0, -1.0, np.inf, 1e-42], np.float32)])
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    mism = (C.c_uint32 * 4)()
    ctx.check(ctx.lib.bt_selftest_numerics(ctx._h, x.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)), len(x), mism))
    assert list(mism) == [0, 0, 0, 0], list(mism)
    exp = np.array([O.lib().bto_log2f(float(v)) for v in x], dtype=np.float32)
    same = (out.view(np.uint32) == exp.view(np.uint32)) | (np.isnan(out) & np.isnan(exp))
    assert same.all(), (x[~same][:5], out[~same][:5], exp[~same][:5])
