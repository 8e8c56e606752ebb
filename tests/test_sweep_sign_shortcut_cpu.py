"""CPU check of the light-sample sweep's sign shortcut (rt_device.cuh:
sample_faces_surface).

pixel() keeps a light sample iff dot(random_direction(), n) > 0 (main.c:193-194),
random_direction() = normalize(random_vector()) (vector.c:99-111).  The device
decides from the un-normalised vector whenever s^2 > tau^2 |rv|^2 with
s = dot(rv, n), tau^2 = 4e-12, and falls back to the literal normalise-then-dot
otherwise.  numpy float32 arithmetic is the same IEEE arithmetic (no contraction),
so the claim "whenever the shortcut fires it agrees with the literal test" can be
hammered here with vectors built to be nearly perpendicular to the normal.
"""
import numpy as np

TAU2 = np.float32(4e-12)
F = np.float32


def literal(rv, n):
    x, y, z = rv[:, 0], rv[:, 1], rv[:, 2]
    norm = np.sqrt((x * x + y * y) + z * z)              # vector.c:113-127 (sqrtf == (float)sqrt((double)s))
    guard = F(float.fromhex("0x1.4f8b58p-17"))            # 1e-5f: (double)n < 1e-5 <=> n <= 1e-5f
    keep = (norm <= guard) & (norm >= -guard)
    with np.errstate(all="ignore"):
        rd = np.where(keep[:, None], rv, rv / norm[:, None])
    return ((rd[:, 0] * n[:, 0] + rd[:, 1] * n[:, 1]) + rd[:, 2] * n[:, 2]) > 0


def shortcut(rv, n):
    x, y, z = rv[:, 0], rv[:, 1], rv[:, 2]
    s = (x * n[:, 0] + y * n[:, 1]) + z * n[:, 2]
    n2 = (x * x + y * y) + z * z
    fires = s * s > TAU2 * n2
    return fires, s > 0


def unit(v):
    v = v.astype(np.float32)
    norm = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    return v / norm[:, None]


def normals(rng, n):
    """Sphere normals (normalised in binary32) and the cubes' axis normals."""
    sph = unit(rng.normal(size=(n, 3)))
    axis = np.zeros((n, 3), np.float32)
    axis[np.arange(n), rng.integers(0, 3, n)] = rng.choice([-1.0, 1.0], n)
    return np.where((rng.random(n) < 0.5)[:, None], sph, axis).astype(np.float32)


def check(rv, n):
    rv = rv.astype(np.float32)
    fires, sign = shortcut(rv, n)
    want = literal(rv, n)
    assert np.array_equal(sign[fires], want[fires])
    return float(fires.mean())


def test_random_vectors_as_the_generator_makes_them():
    rng = np.random.default_rng(5)
    n = 4_000_000
    rv = (rng.random((n, 3)).astype(np.float32) * F(2) - F(1))       # vector.c:99-106
    assert check(rv, normals(rng, n)) > 0.999


def test_vectors_nearly_perpendicular_to_the_normal():
    rng = np.random.default_rng(6)
    n = 4_000_000
    nn = normals(rng, n)
    v = rng.uniform(-1, 1, (n, 3))
    perp = v - (v * nn).sum(1, keepdims=True) * nn                   # in binary64
    eps = rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-9, -4, n)
    rv = perp + eps[:, None] * nn * np.linalg.norm(perp, axis=1, keepdims=True)
    frac = check(rv, nn)
    assert 0.2 < frac < 0.99          # both sides of the threshold are exercised


def test_tiny_and_huge_vectors():
    rng = np.random.default_rng(7)
    n = 1_000_000
    nn = normals(rng, n)
    rv = rng.uniform(-1, 1, (n, 3)) * 10.0 ** rng.integers(-18, 3, (n, 1))
    check(rv, nn)
