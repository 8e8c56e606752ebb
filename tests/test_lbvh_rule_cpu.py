"""CPU proof of the LBVH padding rule (ray_tracing_b200/csrc/rt_lbvh_rule.h).

tests/lbvh_sim.c builds the tree the device builds (binary32 Morton keys, Karras
hierarchy, padded boxes from rt_lbvh_rule.h) and walks it with the device's slab
arithmetic -- including the fma(plane, inv, -(o*inv)) form -- for the rays a frame
really traces (camera rays, light samples, bounces), comparing every hit with the
reference's O(N) scan as restated by the oracle (scene.c:156-190).  The device
walk itself is compared with the oracle on the GPU (tests/test_gpu_parity.py);
this test is what says the RULE is sound, on far more rays than a frame holds and
with (nearly) axis-parallel directions mixed in.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIM_DIR = os.path.join(ROOT, "build", "sim")
SIM = os.path.join(SIM_DIR, "lbvh_sim")


@pytest.fixture(scope="module")
def sim():
    from oracle import bindings

    bindings.build(port=True, ref=False)
    os.makedirs(SIM_DIR, exist_ok=True)
    src = os.path.join(ROOT, "tests", "lbvh_sim.c")
    subprocess.run(["gcc", "-std=c11", "-O2", "-fopenmp", "-ffp-contract=off", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), "-o", SIM, src,
                    "-L" + os.path.join(ROOT, "oracle"), "-lrt_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lm", "-lpthread"], check=True)
    return SIM


def _scene(n, seed, name):
    from ray_tracing_b200 import host, scenes

    path = os.path.join(SIM_DIR, name)
    host.parse_scene_string_large(scenes.synthetic_spheres_text(n, seed=seed)).tofile(path)
    return path


def _run(sim, scene, w, h, gens, cam=None, env=None, extra=()):
    e = dict(os.environ, **(env or {}))
    args = [sim, scene, str(w), str(h), str(gens)] + ([str(v) for v in cam] if cam else []) + list(extra)
    p = subprocess.run(args, capture_output=True, text=True, env=e)
    assert p.returncode in (0, 1), p.stderr
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("topology", ["sah", "karras"])
def test_static_rule_matches_linear_scan(sim, topology):
    """Both topologies the product can build (host SAH: bvh_sah.c, the default; device Karras)."""
    scene = _scene(20000, 5, "spheres20k.bin")
    # the device's slab form (fma), with zero / tiny direction components mixed into the secondary rays
    r = _run(sim, scene, 96, 54, 2, env={"SIM_FMA": "1", "SIM_PACK": "1", "SIM_AXIS": "1", "SIM_TOPOLOGY": topology})
    assert r["mismatches"] == 0 and r["rays"] > 12000, r
    assert r["nodes_per_ray"] < 60 and r["deepest_stack"] <= 32, r


def test_sah_topology_visits_fewer_nodes(sim):
    """What the host builder is for: the same hits for fewer node visits (light samples in any-hit
    mode, as the device walks them).  BASELINE config 5 (100 000 spheres): 44.4 -> 39.5 nodes per ray."""
    scene = _scene(20000, 5, "spheres20k.bin")
    env = {"SIM_FMA": "1", "SIM_PACK": "1", "SIM_ANYHIT": "1"}
    a = _run(sim, scene, 96, 54, 2, env=dict(env, SIM_TOPOLOGY="karras"))
    b = _run(sim, scene, 96, 54, 2, env=dict(env, SIM_TOPOLOGY="sah"))
    assert a["mismatches"] == 0 and b["mismatches"] == 0 and a["rays"] == b["rays"]
    assert b["nodes_per_ray"] < 0.95 * a["nodes_per_ray"], (a, b)


def test_static_rule_far_camera_after_refit(sim):
    scene = _scene(20000, 5, "spheres20k.bin")
    # camera 600 away looking at the cloud: rt_api.cu re-pads for 1.05 x the distance to the farthest corner
    r = _run(sim, scene, 128, 72, 1, cam=(400, 300, 420, -1, -0.72, -1), env={"SIM_FMA": "1", "SIM_PACK": "1", "SIM_DMAX": "760"})
    assert r["mismatches"] == 0, r


def test_mixed_cubes_and_spheres(sim):
    from conftest import random_scene

    objs = random_scene(3000, seed=3, extent=30.0)
    path = os.path.join(SIM_DIR, "mixed3000.bin")
    np.ascontiguousarray(objs).tofile(path)
    for topology in ("sah", "karras"):
        r = _run(sim, path, 120, 68, 2, cam=(40, 25, 40, -1, -0.6, -1),
                 env={"SIM_FMA": "1", "SIM_PACK": "1", "SIM_AXIS": "1", "SIM_TOPOLOGY": topology})
        assert r["mismatches"] == 0 and r["rays"] > 8000, r


def test_axis_parallel_and_degenerate_directions(sim):
    """The rays of tests/test_gpu_parity.py::test_trace_random_scenes_linear_and_lbvh: one-hot
    directions, zero components, 1e-7-long directions.  (With an infinite reciprocal the fma slab
    form turned inf - inf into a NaN that emptied an unbounded interval: walk_inverse() caps it.)"""
    from conftest import random_scene

    for seed, n, spheres in ((1, 200, True), (2, 1024, False)):
        objs = random_scene(n, seed, spheres_only=spheres)
        path = os.path.join(SIM_DIR, f"rand{n}.bin")
        np.ascontiguousarray(objs).tofile(path)
        rng = np.random.default_rng(seed + 100)
        m = 20000
        rays = np.concatenate([rng.uniform(-8, 8, (m, 3)), rng.normal(size=(m, 3))], axis=1).astype(np.float32)
        tgt = objs["geom"][rng.integers(0, n, m // 2), :3] + rng.normal(scale=0.3, size=(m // 2, 3))
        rays[m // 2:, 3:] = tgt - rays[m // 2:, :3]
        rays[:50, 3:] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 50)] * rng.choice([-1, 1], (50, 1))
        rays[50:60, 3:] = 1e-7
        rays[60:70, 4] = 0.0
        rpath = os.path.join(SIM_DIR, f"rays{n}.bin")
        rays.tofile(rpath)
        r = _run(sim, path, 200, 100, 0, env={"SIM_FMA": "1", "SIM_PACK": "1", "SIM_RAYS": rpath})
        assert r["mismatches"] == 0 and r["rays"] == m, r


def test_rejected_per_node_rule_is_sound_too(sim):
    """The measured-and-rejected alternative (tight boxes widened per visited node):
    kept runnable so the numbers quoted in rt_lbvh_rule.h can be reproduced."""
    scene = _scene(20000, 5, "spheres20k.bin")
    a = _run(sim, scene, 96, 54, 1)
    b = _run(sim, scene, 96, 54, 1, extra=("--dynamic-pad",))
    assert a["mismatches"] == 0 and b["mismatches"] == 0
    assert b["nodes_per_ray"] <= a["nodes_per_ray"]
