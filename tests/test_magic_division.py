"""Index math without integer division (rt_render.cu: div_magic; rt_api.cu:
magic_tiles_x, magic_cells_per_col): n / d == (n * ceil(2^40 / d)) >> 40 as long
as n * d < 2^40 (exactness) and n / d < 2^24 (the 64-bit product does not wrap).
The kernels use it with d = tiles per row / cells per column and n = a tile or
cell index of the same launch, so the quotient is a tile row / column index
(4K: d <= 480, n < 2^18).  Checked at and around every multiple of d."""
import numpy as np


def magic(d):
    return ((1 << 40) + d - 1) // d


def div_magic(n, m):
    return (n.astype(np.uint64) * np.uint64(m)) >> np.uint64(40)


def test_magic_division_is_exact_below_the_limit():
    rng = np.random.default_rng(9)
    ds = list(range(1, 2049)) + [int(x) for x in rng.integers(2049, 1 << 16, 400)]
    for d in ds:
        m = magic(d)
        limit = min((1 << 40) // d, d << 24, 1 << 32)   # n * d < 2^40, n / d < 2^24, n fits the kernel's unsigned
        ks = rng.integers(0, max(limit // d, 1), 64).astype(np.uint64)
        n = np.concatenate([ks * np.uint64(d), ks * np.uint64(d) + np.uint64(d - 1), (ks + np.uint64(1)) * np.uint64(d),
                            rng.integers(0, limit, 256).astype(np.uint64), np.array([0, limit - 1], np.uint64)])
        n = n[n < limit]
        assert np.array_equal(div_magic(n, m), n // np.uint64(d)), d


def test_frame_sizes_stay_below_the_limit():
    # largest products the launches form: tile index * tiles_x and cell index * cells_per_col
    for w, h in ((3840, 2160), (7680, 4320), (16384, 16384)):
        tiles_x, tiles_y = (w + 7) // 8, (h + 3) // 4
        assert tiles_x * tiles_y * tiles_x < 1 << 40
        assert w * w < 1 << 40
