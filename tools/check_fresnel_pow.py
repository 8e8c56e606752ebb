#!/usr/bin/env python
"""fresnel_schlick (main.c:126-129) evaluates (float)pow(1.0 - (double)u, 5.0);
the device evaluates x2 = x*x; (float)(x2*x2*x) in binary64 (rt_device.cuh:
path_launch).  u = clamp(n.v, 0, 1) is a binary32 in [0, 1]: there are
1 065 353 217 of them, so the equivalence can be checked EXHAUSTIVELY against
this image's libm (numpy's float64 power calls the same glibc pow).

    python tools/check_fresnel_pow.py            # all binary32 in [0, 1]  (~1-2 min)
    python tools/check_fresnel_pow.py --quick    # 1/64 of them + the ends
"""
import sys

import numpy as np


def mismatches(bits):
    u = bits.view(np.float32).astype(np.float64)
    x = 1.0 - u
    want = np.power(x, 5.0).astype(np.float32)
    x2 = x * x
    got = (x2 * x2 * x).astype(np.float32)
    return int((want.view(np.uint32) != got.view(np.uint32)).sum())


def main():
    quick = "--quick" in sys.argv
    one = np.float32(1.0).view(np.uint32)          # 0x3f800000: bits of every float in [0, 1] are 0 .. this
    step = 1 << 24
    bad = total = 0
    for lo in range(0, int(one) + 1, step):
        hi = min(lo + step, int(one) + 1)
        bits = np.arange(lo, hi, 64 if quick else 1, dtype=np.uint32)
        bad += mismatches(bits)
        total += len(bits)
    bad += mismatches(np.array([0, 1, int(one) - 1, int(one)], np.uint32))
    print(f"{total} values of u in [0,1], {bad} mismatches between pow(1-u, 5) and (x*x)*(x*x)*x after rounding to binary32")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
