"""Host-side BVH topology builder (ray_tracing_b200/csrc/bvh_sah.c), no GPU needed.

The builder only decides the SHAPE of the tree (the reference has none, scene.c:156-173 scans every
object); these tests check that the shape is a valid binary tree in the device build's conventions
(rt_lbvh.cu: hierarchy_kernel), that its depth respects the walk's stack bound for hostile inputs,
and -- through tests/lbvh_sim.c -- that walking it returns the linear scan's hits with fewer node
visits than the Karras tree.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ray_tracing_b200 import host  # noqa: E402


def build(A, B, lib=None):
    lib = lib or host.load_library()
    fn = lib.rt_host_bvh_sah
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_float,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    n = len(A)
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    prim = np.full(n, -7, np.int32)
    children = np.full(2 * max(n - 1, 1), -7, np.int32)
    parent = np.full(2 * n - 1, -7, np.int32)
    depth = C.c_int(-1)
    rc = fn(A.ctypes.data, B.ctypes.data, n, 0.04, 1e-5, 1e-4, prim.ctypes.data, children.ctypes.data,
            parent.ctypes.data, C.byref(depth))
    assert rc == 0
    return prim, children, parent, depth.value


def geometry(n, seed, kinds=(0, 1)):
    """geomA / geomB records in the device layout (rt_host.h): type as int bits in geomB.w"""
    rng = np.random.default_rng(seed)
    A = np.zeros((n, 4), np.float32)
    B = np.zeros((n, 4), np.float32)
    ty = rng.choice(kinds, n).astype(np.int32)
    A[:, :3] = rng.uniform(-20, 20, (n, 3))
    r = rng.uniform(0.1, 0.5, n)
    A[:, 3] = np.where(ty == 1, r * r, 0)
    B[:, :3] = np.where((ty == 0)[:, None], A[:, :3] + rng.uniform(0.1, 2.0, (n, 3)), 0)
    B[:, 3] = ty.view(np.float32)
    return A, B


def check_tree(n, prim, children, parent, depth):
    assert sorted(prim.tolist()) == list(range(n))
    if n == 1:
        assert parent[0] == -1
        return
    assert parent[0] == -1
    seen_leaf = np.zeros(n, bool)
    seen_node = np.zeros(n - 1, bool)
    deepest = 0
    stack = [(0, 0)]
    while stack:
        node, d = stack.pop()
        assert 0 <= node < n - 1 and not seen_node[node]
        seen_node[node] = True
        for side in (0, 1):
            c = int(children[2 * node + side])
            if c < 0:
                s = ~c
                assert 0 <= s < n and not seen_leaf[s]
                seen_leaf[s] = True
                assert parent[(n - 1) + s] == node
                deepest = max(deepest, d + 1)
            else:
                assert parent[c] == node
                assert c > node                      # preorder numbering
                stack.append((c, d + 1))
    assert seen_leaf.all() and seen_node.all()
    assert deepest == depth


@pytest.mark.parametrize("n", [1, 2, 3, 17, 5000, 40000])
def test_topology_is_a_valid_tree(n):
    A, B = geometry(n, seed=n, kinds=(0, 1) if n < 5000 else (0, 1, 5))
    prim, children, parent, depth = build(A, B)
    check_tree(n, prim, children, parent, depth)
    if n >= 5000:
        assert depth <= 2.5 * np.log2(n)


def test_coincident_and_non_finite_centroids():
    A, B = geometry(3000, seed=4, kinds=(1,))
    A[:, :3] = 7.25                                    # every centroid the same: median splits
    prim, children, parent, depth = build(A, B)
    check_tree(3000, prim, children, parent, depth)
    assert depth == 12
    A, B = geometry(3000, seed=5, kinds=(1,))
    A[::7, 0] = np.nan
    A[3::11, 1] = np.inf
    A[5::13, 2] = -np.inf
    prim, children, parent, depth = build(A, B)
    check_tree(3000, prim, children, parent, depth)


def test_depth_is_capped_for_lopsided_scenes():
    """Three chains of spheres at 16^k along the axes: SAH splits keep peeling a few spheres off,
    a tree 35 deep for 186 spheres.  The builder must fall back to median splits before the walk's
    stack bound (RT_BVH_STACK = 64; bvh_sah.c: RT_SAH_MAX_DEPTH = 60): checked with the file built
    for a cap of 20."""
    import subprocess

    pts = []
    for axis in range(3):
        for k in range(-31, 31):
            p = [0.0, 0.0, 0.0]
            p[axis] = 16.0 ** k
            pts.append(p)
    n = len(pts)
    A = np.zeros((n, 4), np.float32)
    B = np.zeros((n, 4), np.float32)
    A[:, :3] = np.array(pts, np.float32)
    A[:, 3] = 1e-12
    B[:, 3] = np.full(n, 1, np.int32).view(np.float32)
    prim, children, parent, depth = build(A, B)
    check_tree(n, prim, children, parent, depth)
    assert 20 < depth <= 60, depth
    out = os.path.join(ROOT, "build", "sim")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libsah_cap20.so")
    subprocess.run(["gcc", "-std=c11", "-O2", "-shared", "-fPIC", "-DRT_SAH_MAX_DEPTH=20", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "ray_tracing_b200", "csrc"), "-o", so,
                    os.path.join(ROOT, "ray_tracing_b200", "csrc", "bvh_sah.c"), "-lm", "-lpthread"], check=True)
    prim, children, parent, capped = build(A, B, C.CDLL(so))
    check_tree(n, prim, children, parent, capped)
    assert capped <= 20 < depth, (capped, depth)
