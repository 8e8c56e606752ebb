"""Scene descriptions in the reference's scene-file grammar (scene.c:206-609).

* ``builtin_scene_text(k)`` -- the three scenes the reference ships
  (scene_0/1/2.txt: 9/7/3 objects), re-emitted from compact tables so the
  benchmark does not depend on /root/reference being present.  tests/ check
  that they parse to the same `Object` records as the reference's files.
* ``synthetic_spheres_text(n, seed)`` -- BASELINE.json config 5: n spheres in
  the reference grammar with its quirks respected (>= 3 blanks after `albedo`
  and `metallic`, scene.c:280,320; no exponents; no comments), object 0 an
  emissive sphere so the light-sampling path is exercised.

Values are kept as decimal *strings*: the reference builds floats digit by
digit in binary32 (scene.c:441-461), so the text is the ground truth.
"""
from __future__ import annotations

import numpy as np

# (kind, geometry, albedo, roughness, reflectance, metallic, emission_power, emission_color)
_CUBE, _SPHERE = "cube", "sphere"

_SCENE_0 = [
    (_CUBE, ("0 0 0", "3 5 0.1"), "1 0.3 0.3", "1", "0", "1", "0", "0 0 0"),
    (_CUBE, ("3 0 0", "3 5 0.1"), "1 0.3 0.3", "0.5", "0", "1", "0", "0 0 0"),
    (_CUBE, ("6 0 0", "3 5 0.1"), "1 0.3 0.3", "0", "0", "1", "0", "0 0 0"),
    (_CUBE, ("0 -0.1 0", "9 0.1 9"), "0.4 0.3 0.9", "1", "0", "0", "0", "0 0 0"),
    (_CUBE, ("5 0 6", "1 1 1"), "1 0 0", "1", "0", "0", "0", "0 0 0"),
    (_CUBE, ("4 0 5", "1 1 1"), "1 0 1", "0", "1", "0", "0", "0 0 0"),
    (_SPHERE, ("3 1 3", "1"), "1 0.4 0", "1", "0", "0", "0", "0 0 0"),
    (_SPHERE, ("5 1 3", "1"), "0 1 0", "0", "1", "0", "0", "0 0 0"),
    (_SPHERE, ("3 5 3", "1"), "1 0.4 0", "1", "0", "0", "5", "1 1 1"),
]

_SCENE_1 = [
    (_CUBE, ("0 0 0", "3 0.1 3"), "1 0.3 0.3", "1", "0", "0", "0", "0 0 0"),
    (_CUBE, ("0 5 0", "3 0.1 3"), "0.3 1 0.3", "1", "0", "0", "0", "0 0 0"),
    (_CUBE, ("0 0 0", "0.1 5 3"), "0.3 0.3 1", "0", "1", "1", "0", "0 0 0"),
    (_CUBE, ("3 0 0", "0.1 5 3"), "0.3 1 1", "0", "1", "1", "0", "0 0 0"),
    (_CUBE, ("0 0 0", "3 5 0.1"), "1 0.3 1", "1", "1", "0", "0", "0 0 0"),
    (_CUBE, ("1 4.9 1", "1 0.1 1"), "1 1 0.3", "0", "1", "0", "1", "1 1 1"),
    (_SPHERE, ("1.5 1 1.5", "1"), "0 1 0", "0", "1", "1", "0", "0 0 0"),
]

_SCENE_2 = [
    (_SPHERE, ("-3 0 0", "1"), "0.2 0.5 1", "0", "1", "0", "0", "0 0 0"),
    (_SPHERE, ("0 0 0", "1"), "0.2 0.5 1", "0", "0", "0", "0", "0 0 0"),
    (_SPHERE, ("3 0 0", "1"), "0.5 0.2 1", "0", "0", "1", "0", "0 0 0"),
]

BUILTIN = {0: _SCENE_0, 1: _SCENE_1, 2: _SCENE_2}


def _emit(obj) -> str:
    kind, geom, albedo, rough, refl, metal, epow, ecol = obj
    lines = [kind]
    lines.append(f"  emission_color {{{ecol}}}")
    lines.append(f"  emission_power {epow}")
    lines.append(f"  metallic    {metal}")      # >= 3 blanks: the parser skips 11 chars
    lines.append(f"  reflectance {refl}")
    lines.append(f"  roughness   {rough}")
    lines.append(f"  albedo    {{{albedo}}}")   # >= 3 blanks: the parser skips 9 chars
    if kind == _CUBE:
        lines.append(f"  origin {{{geom[0]}}}")
        lines.append(f"  size   {{{geom[1]}}}")
    else:
        lines.append(f"  center {{{geom[0]}}}")
        lines.append(f"  radius {geom[1]}")
    return "\n".join(lines) + "\n"


def builtin_scene_text(k: int) -> str:
    return "\n".join(_emit(o) for o in BUILTIN[k])


def _dec(x: float, places: int = 3) -> str:
    """Plain decimal, no exponent (the grammar has none)."""
    s = f"{x:.{places}f}"
    return s


def synthetic_spheres_text(n: int = 100_000, seed: int = 20261017, extent=((-50.0, 50.0), (0.0, 20.0), (-50.0, 50.0))) -> str:
    """SURVEY.md 8(d) config 5 generator (numpy default_rng, seed 20261017)."""
    rng = np.random.default_rng(seed)
    out = []
    # object 0: the light
    out.append(_emit((_SPHERE, ("0 40 0", "5"), "1 1 1", "1", "0", "0", "5", "1 1 1")))
    m = n - 1
    cx = rng.uniform(extent[0][0], extent[0][1], m)
    cy = rng.uniform(extent[1][0], extent[1][1], m)
    cz = rng.uniform(extent[2][0], extent[2][1], m)
    rad = rng.uniform(0.1, 0.5, m)
    alb = rng.uniform(0.0, 1.0, (m, 3))
    rough = rng.choice(["0", "0.5", "1"], m)
    metal = np.where(rng.uniform(0, 1, m) < 0.2, "1", "0")
    refl = rng.uniform(0.0, 1.0, m)
    for i in range(m):
        out.append(
            _emit(
                (
                    _SPHERE,
                    (f"{_dec(cx[i])} {_dec(cy[i])} {_dec(cz[i])}", _dec(rad[i])),
                    f"{_dec(alb[i, 0])} {_dec(alb[i, 1])} {_dec(alb[i, 2])}",
                    rough[i],
                    _dec(refl[i]),
                    metal[i],
                    "0",
                    "0 0 0",
                )
            )
        )
    return "".join(out)


def procedural_skybox(size: int = 256, seed: int = 11) -> np.ndarray:
    """Deterministic stand-in cubemap (6, size, size, 3) u8 for runs where the
    reference's JPEGs are not staged: per-face gradients plus value noise so
    neighbouring texels differ (the lookup is nearest-texel)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / max(size - 1, 1)
    faces = np.zeros((6, size, size, 3), np.uint8)
    for f in range(6):
        base = np.stack([0.35 + 0.4 * x * (f + 1) / 6.0, 0.45 + 0.4 * y * (6 - f) / 6.0, 0.6 + 0.3 * (x + y) * 0.5], axis=-1)
        noise = rng.integers(0, 24, size=(size, size, 3))
        faces[f] = np.clip(base * 220 + noise, 0, 255).astype(np.uint8)
    return faces


FACE_FILES = ("front.jpg", "back.jpg", "left.jpg", "right.jpg", "top.jpg", "bottom.jpg")  # CubeFace order
