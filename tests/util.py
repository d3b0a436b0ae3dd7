"""Shared helpers for the parity tests."""
import math

import numpy as np

# north_star tolerance: losses and gradients within 1e-5 relative in fp32.
# "Relative" is taken against the quantity's scale (its max magnitude, floored by
# the natural scale given by the caller: ln K for a JSD map, the upstream
# gradient magnitude for a gradient) because maps/gradients are differences of
# O(1) terms: the reference's own fp32 run sits 2e-7..1.6e-6 from its fp64 run
# by this measure (recorded in DESIGN.md), so 1e-5 is ~10x the noise floor.
RTOL_FP32 = 1e-5


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(float(np.abs(b).max()), floor, 1e-300)
    return float(np.abs(a - b).max()) / scale


def assert_close(a, b, rtol=RTOL_FP32, floor=0.0, what=""):
    e = rel_err(a, b, floor)
    assert e <= rtol, f"{what}: scaled error {e:.3e} > {rtol:.1e}"
    return e


def cases(golden, prefix):
    return sorted({k.split("/")[0] for k in golden.files if k.startswith(prefix)})


def lnK(K):
    return math.log(K)
