"""Product scene parser (ray_tracing_b200/csrc/scene_parse.c) against the
reference's parse_scene_file (scene.c:611-624) and the golden object dumps."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from ray_tracing_b200 import host, scenes


def canon(o):
    o = np.array(o, dtype=host.OBJECT_DTYPE, copy=True)
    o["geom"][o["type"] == 1, 4:] = 0      # the reference leaves the sphere's union tail unset (scene.c:222)
    return o


def test_builtin_scenes_match_golden_dumps():
    gold = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    for k, n in ((0, 9), (1, 7), (2, 3)):
        mine = host.parse_scene_string(scenes.builtin_scene_text(k))
        assert len(mine) == n
        assert canon(mine).tobytes() == gold[f"scene{k}_objects"].tobytes()


def test_float_construction_known_answers():
    # SURVEY.md 8(c) parser KAT: digit-by-digit binary32 accumulation, not strtof
    o = host.parse_scene_string("sphere radius 0.3 roughness 0.7 reflectance 0.123456789")
    assert float.hex(float(o[0]["geom"][3])) == "0x1.3333340000000p-2"
    assert float.hex(float(o[0]["roughness"])) == "0x1.6666660000000p-1"
    assert float.hex(float(o[0]["reflectance"])) == "0x1.f9add40000000p-4"
    d = host.parse_scene_string("cube")
    assert float.hex(float(d[0]["reflectance"])) == "0x1.99999a0000000p-3"
    assert d[0]["geom"].tolist() == [0, 0, 0, 1, 1, 1]
    assert np.allclose(d[0]["albedo"], [0.44, 0.68, 0.84])


def test_keyword_skip_quirks():
    # `albedo` advances 9 chars, `metallic` 11 (scene.c:280,320)
    assert host.parse_scene_string("sphere albedo {1 0 0}") is None
    assert host.parse_scene_string("sphere albedo    {1 0 0}") is not None
    assert host.parse_scene_string("sphere metallic 0.5") is None          # parsed as 5 -> range error
    o = host.parse_scene_string("sphere metallic    0.5")
    assert o[0]["metallic"] == 0.5


ERRORS = [
    "banana",
    "sphere radius",
    "sphere radius -",
    "sphere radius 1.",
    "sphere center {1 2}",
    "sphere center {1 2 3",
    "sphere center 1 2 3",
    "cube radius 1",
    "cube center {0 0 0}",
    "sphere origin {0 0 0}",
    "sphere size {1 1 1}",
    "cube size {-1 1 1}",
    "sphere roughness 2",
    "sphere reflectance 1.5",
    "sphere albedo    {2 0 0}",
    "sphere emission_color {0 0 -1}",
    "sphere radius 1 x",
]

VALID = [
    "",
    "   \n\t\r\n",
    "sphere",
    "cube",
    "sphere radius 2 radius 3",
    "sphere\ncube\nsphere",
    "cube origin {-1.5 2.25 -0.001} size {0 0 0}",
    "sphere center {  1\n 2\t3  } radius 0.5 emission_power -3.75",
    "sphere emission_power 12345678.9",
    "spherecube",                         # prefix match, no separator needed
    "sphereradius 2",
    "cube size {1 1 1}sphere",
    "sphere radius 0.1234567890123456789",
    "sphere radius -0",
]


@pytest.mark.parametrize("text", VALID + ERRORS)
def test_against_reference_parser(ref_pixel, tmp_path, text):
    p = tmp_path / "s.txt"
    p.write_text(text)
    want = ref_pixel.parse_scene_file_objects(str(p))
    got = host.parse_scene_file(str(p))
    if want is None:
        assert got is None, text
    else:
        assert got is not None, text
        assert canon(got).tobytes() == canon(want).tobytes(), text


def _fuzz_scene(rng):
    """Grammar-directed generator: mostly valid objects, ~15 % corrupted tokens."""
    num = lambda: rng.choice(["0", "1", "0.5", "0.25", "12.75", "-3", "-0.001", "7.", "x", "0.3333333", "100"])
    vec = lambda: "{" + " ".join(num() for _ in range(int(rng.choice([3, 3, 3, 3, 2])))) + "}"
    props = {"albedo   ": vec, "roughness": num, "reflectance": num, "metallic   ": num, "emission_power": num,
             "emission_color": vec, "radius": num, "center": vec, "origin": vec, "size": vec}
    sphere_ok = ["albedo   ", "roughness", "reflectance", "metallic   ", "emission_power", "emission_color", "radius", "center"]
    cube_ok = ["albedo   ", "roughness", "reflectance", "metallic   ", "emission_power", "emission_color", "origin", "size"]
    out = []
    for _ in range(int(rng.integers(0, 5))):
        kind = str(rng.choice(["sphere", "cube"]))
        out.append(kind)
        for _ in range(int(rng.integers(0, 6))):
            pool = list(props) if rng.uniform() < 0.1 else (sphere_ok if kind == "sphere" else cube_ok)
            k = str(rng.choice(pool))
            sep = str(rng.choice([" ", "  ", "\n", "\t", ""]))
            out.append(k + sep + props[k]())
    return str(rng.choice([" ", "\n", "\n\n"])).join(out)


def test_fuzzed_files_against_reference(ref_pixel, tmp_path):
    rng = np.random.default_rng(42)
    agree_ok = agree_err = 0
    for i in range(600):
        text = _fuzz_scene(rng)
        p = tmp_path / f"f{i}.txt"
        p.write_text(text)
        want = ref_pixel.parse_scene_file_objects(str(p))
        got = host.parse_scene_file(str(p))
        assert (want is None) == (got is None), repr(text)
        if want is not None:
            agree_ok += 1
            assert canon(got).tobytes() == canon(want).tobytes(), repr(text)
        else:
            agree_err += 1
    assert agree_ok > 100 and agree_err > 100


def test_valid_generated_scenes_against_reference(ref_pixel, tmp_path):
    text = scenes.synthetic_spheres_text(500, seed=3)
    p = tmp_path / "gen.txt"
    p.write_text(text)
    want = ref_pixel.parse_scene_file_objects(str(p))
    got = host.parse_scene_file(str(p))
    assert len(got) == 500 and canon(got).tobytes() == canon(want).tobytes()
    assert got[0]["emission_power"] == 5.0 and got[0]["geom"][1] == 40.0


def test_capacity_and_large_variant(tmp_path):
    text = "sphere\n" * 1030
    o = host.parse_scene_string(text)
    assert len(o) == 1024                      # scene.c:602-605: the rest is dropped with a warning
    big = host.parse_scene_string_large(text)
    assert len(big) == 1030
    assert canon(big[:1024]).tobytes() == canon(o).tobytes()
    with pytest.raises(host.RtError):
        host.parse_scene_string_large("sphere radius x")
    assert host.parse_scene_file(str(tmp_path / "missing.txt")) is None


def test_partial_count_on_error():
    ok, objs = host.parse_scene_string_partial("sphere\ncube\nsphere radius x")
    assert not ok and len(objs) == 2           # scene.c:208: num_objects keeps what was parsed


def test_streaming_file_parser_equals_in_memory_parse(tmp_path):
    """rt_parse_scene_file_large reads through a 1 MiB window cut at object
    boundaries; the records must equal those of the in-memory
    parse of the same text, across several windows, and errors keep their line."""
    from ray_tracing_b200 import host, scenes

    text = scenes.synthetic_spheres_text(20000, seed=3)            # ~3.6 MB: four windows
    cubes = "".join(f"cube origin {{{i} 0 {i}}} size {{1 2 3}}\n   albedo   {{0.5 0.25 1}}  roughness 0.5\n" for i in range(3000))
    text = text[: len(text) // 2].rsplit("sphere", 1)[0] + cubes + text[len(text) // 2:].split("sphere", 1)[1].join(["sphere", ""])
    path = tmp_path / "big.txt"
    path.write_text(text)
    a = host.parse_scene_file_large(str(path))
    b = host.parse_scene_string_large(text)
    assert len(a) == len(b) > 20000
    assert np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))
    bad = tmp_path / "bad.txt"
    lines = text.split("\n")
    lines[len(lines) * 3 // 4] = "sphere radius x"
    bad.write_text("\n".join(lines))
    with pytest.raises(host.RtError):
        host.parse_scene_file_large(str(bad))
