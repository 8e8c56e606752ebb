"""The reference compares binary32 values with double literals (C promotes the
float): normalize()'s `norm < 1e-5 && norm > -1e-5` (vector.c:132) and
iszerof()'s `f < 0.0001 && f > -0.0001` (vector.c:81).  The device compares in
binary32 against the largest float below the literal (rt_device.cuh: unit3,
near_zero).  Equivalence = no binary32 lies between that float and the literal;
checked on the floats around each boundary."""
import numpy as np


def neighbours(x, k=8):
    out = [np.float32(x)]
    for _ in range(k):
        out.append(np.nextafter(out[-1], np.float32(np.inf)))
    lo = np.float32(x)
    for _ in range(k):
        lo = np.nextafter(lo, np.float32(-np.inf))
        out.append(lo)
    return np.array(out, np.float32)


def check(literal, device_threshold_hex):
    thr = np.float32(float.fromhex(device_threshold_hex))
    assert float(thr) < literal < float(np.nextafter(thr, np.float32(np.inf)))      # literal sits strictly between two floats
    for sign in (1.0, -1.0):
        f = neighbours(literal) * np.float32(sign)
        ref = (f.astype(np.float64) < literal) & (f.astype(np.float64) > -literal)  # the C expression, promoted
        dev = (f <= thr) & (f >= -thr)                                              # the device expression
        assert np.array_equal(ref, dev)


def test_normalize_guard():
    check(1e-5, "0x1.4f8b58p-17")


def test_iszero_guard():
    check(0.0001, "0x1.a36e2ep-14")


def test_device_source_uses_these_thresholds():
    import os

    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ray_tracing_b200", "csrc", "rt_device.cuh")).read()
    assert "n <= 0x1.4f8b58p-17f && n >= -0x1.4f8b58p-17f" in src
    assert "f <= 0x1.a36e2ep-14f && f >= -0x1.a36e2ep-14f" in src
