"""CPU proof of the device's sphere-root shortcut (rt_device.cuh: sphere_entry).

The reference evaluates both roots of the ray/sphere quadratic in binary64 and
rounds them to binary32 (scene.c:114-127).  The CUDA path divides only the root
it returns: a numerator x = -b -+ sqrt(discr) below -2^-100 is taken as negative
without dividing.  numpy's float64 sqrt/divide and float32 casts are the same
correctly rounded IEEE operations the GPU executes, so both algorithms can be run
side by side here on far more (and far nastier) operands than any frame contains.
What must agree is what trace_ray sees: whether the sphere is accepted
(`t >= 0 && t < nearest_t`, scene.c:168) and, if so, the bits of t.
"""
import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)


def literal(nb, sq, a2):
    with np.errstate(all="ignore"):
        tm = ((nb - sq) / a2).astype(np.float32)
        tp = ((nb + sq) / a2).astype(np.float32)
    t = np.where(tm < 0, tp, tm)
    hit = ~((tm < 0) & (tp < 0))
    return hit, t


def shortcut(nb, sq, a2):
    lim = np.float64(-2.0 ** -100)
    with np.errstate(all="ignore"):
        xm, xp = nb - sq, nb + sq
        tm = (xm / a2).astype(np.float32)
        tp = (xp / a2).astype(np.float32)
    # first trip: the minus root is divided unless its numerator is clearly negative
    m_neg = (xm < lim) | (tm < 0)
    p_neg = (xp < lim) | (tp < 0)
    hit = ~(m_neg & p_neg)
    t = np.where(m_neg, tp, tm)
    return hit, t


def accepted(hit, t):
    with np.errstate(invalid="ignore"):
        return hit & (t >= 0) & (t < FLT_MAX)


def check(b, discr, a):
    b = b.astype(np.float32)
    discr = discr.astype(np.float32)
    a = a.astype(np.float32)
    keep = discr > 0                                # scene.c:117: discr > 0 or no intersection
    b, discr, a = b[keep], discr[keep], a[keep]
    nb = (-b).astype(np.float64)
    with np.errstate(all="ignore"):
        sq = np.sqrt(discr.astype(np.float64))
        a2 = (np.float32(2.0) * a).astype(np.float64)
    h0, t0 = literal(nb, sq, a2)
    h1, t1 = shortcut(nb, sq, a2)
    a0, a1 = accepted(h0, t0), accepted(h1, t1)
    assert np.array_equal(a0, a1)
    assert np.array_equal(t0[a0].view(np.uint32), t1[a1].view(np.uint32))
    return int(a0.sum()), len(b)


def test_random_operands_in_the_usual_range():
    rng = np.random.default_rng(1)
    n = 4_000_000
    b = rng.normal(0, 20, n)
    discr = np.abs(rng.normal(0, 400, n)) * 10.0 ** rng.integers(-8, 3, n)
    a = 1.0 + rng.integers(-4, 5, n) * 2.0 ** -23        # d.d of a normalised direction
    acc, tot = check(b, discr, a)
    assert acc > tot // 10


def test_numerators_that_almost_cancel():
    """-b ~ +-sqrt(discr): the returned root is tiny, zero or barely negative."""
    rng = np.random.default_rng(2)
    n = 4_000_000
    s = np.abs(rng.normal(0, 5, n)).astype(np.float32) + np.float32(1e-3)
    discr = (s.astype(np.float64) ** 2).astype(np.float32)
    ulps = rng.integers(-3, 4, n)
    b = np.nextafter(s, np.where(ulps > 0, np.float32(np.inf), np.float32(-np.inf)))
    b = np.where(ulps == 0, s, b) * rng.choice([-1.0, 1.0], n).astype(np.float32)
    a = 1.0 + rng.integers(-4, 5, n) * 2.0 ** -23
    check(b, discr, a)


def test_extreme_magnitudes_and_degenerate_directions():
    rng = np.random.default_rng(3)
    n = 2_000_000
    eb = rng.integers(-149, 128, n).astype(np.float64)
    ed = rng.integers(-149, 128, n).astype(np.float64)
    b = (rng.uniform(1, 2, n) * 2.0 ** eb * rng.choice([-1.0, 1.0], n))
    discr = rng.uniform(1, 2, n) * 2.0 ** ed
    # a = d.d for d = normalize(direction) (scene.c:158): ~1; <= 3e-10 when normalize() left a
    # direction shorter than 1e-5 alone; 0 when the squared norm overflowed (d = v/inf); NaN for
    # non-finite directions.  It is never large: the shortcut needs 2a < 2^50 (-2^-100/2a must not
    # round to -0), so a few absurdly large values are in the list, +inf -- unreachable -- is not.
    a = rng.choice([1.0, 1.0 - 2.0 ** -23, 1.0 + 2.0 ** -22, 1e-10, 1e-30, 0.0, 4.0, 1e6, 2.0 ** 48, np.nan], n)
    check(b, discr, a)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3e38, -3e38, 1.0, -1.0])
    bb, dd, aa = np.meshgrid(special, special, np.array([1.0, 0.0, 1e-10, 2.0 ** 48, np.nan]), indexing="ij")
    check(bb.ravel(), dd.ravel(), aa.ravel())
