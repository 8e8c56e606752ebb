import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    """CPU restatement oracle (oracle/rt_oracle.c)."""
    from oracle.bindings import Port, build

    build(port=True, ref=False)
    return Port()


def _ref(variant):
    from oracle import bindings

    if not bindings.ref_available(variant):
        pytest.skip(f"oracle/_ref/libref_{variant}.so not built (needs /root/reference at build time)")
    return bindings.Ref(variant)


@pytest.fixture(scope="session")
def ref_pixel():
    """Unmodified reference, RNG re-keyed per pixel (the parity target)."""
    return _ref("pixel")


@pytest.fixture(scope="session")
def ref_stream():
    """Unmodified reference as shipped (per-thread RNG stream)."""
    return _ref("stream")


@pytest.fixture(scope="session")
def ref_big():
    return _ref("pixel_big")


@pytest.fixture(scope="session")
def small_sky():
    from oracle.bindings import procedural_skybox

    return procedural_skybox(64, seed=7)


@pytest.fixture(scope="session")
def real_sky():
    """The reference's skybox decoded by the reference's own loader (stb_image);
    falls back to a 512^2 procedural cubemap when the staged JPEGs are absent."""
    from oracle import bindings

    jpg = os.path.join(bindings.ASSETS, "skybox", "front.jpg")
    if bindings.ref_available("stream") and os.path.exists(jpg):
        r = bindings.Ref("stream")
        return r.load_skybox()
    return bindings.procedural_skybox(512, seed=3)


@pytest.fixture(scope="session")
def builtin_objects():
    from ray_tracing_b200 import host, scenes

    return {k: host.parse_scene_string(scenes.builtin_scene_text(k)) for k in (0, 1, 2)}


@pytest.fixture(scope="session")
def renderer():
    """The CUDA path through the C ABI.  Fails loudly without the library."""
    from ray_tracing_b200 import host

    r = host.Renderer(num_gpus=1)
    yield r
    r.close()


def random_scene(n, seed, spheres_only=False, extent=6.0, emissive=True):
    """Random Object records (reference layout) with values that are exact in
    the scene-file grammar's float construction (3 decimals)."""
    from ray_tracing_b200.host import OBJECT_DTYPE

    rng = np.random.default_rng(seed)
    o = np.zeros(n, OBJECT_DTYPE)
    o["type"] = 1 if spheres_only else rng.integers(0, 2, n)
    o["geom"][:, :3] = np.round(rng.uniform(-extent, extent, (n, 3)), 3)
    sph = o["type"] == 1
    o["geom"][sph, 3] = np.round(rng.uniform(0.1, 0.9, sph.sum()), 3)
    o["geom"][~sph, 3:] = np.round(rng.uniform(0.2, 1.5, ((~sph).sum(), 3)), 3)
    o["albedo"] = np.round(rng.uniform(0, 1, (n, 3)), 3)
    o["roughness"] = rng.choice([0.0, 0.5, 1.0], n)
    o["reflectance"] = np.round(rng.uniform(0, 1, n), 3)
    o["metallic"] = rng.choice([0.0, 0.0, 0.0, 1.0], n)
    o["emission_power"] = 0
    o["emission_color"] = 0
    if emissive and n > 2:
        k = n // 2
        o["emission_power"][k] = 4.0
        o["emission_color"][k] = 1.0
    return o
